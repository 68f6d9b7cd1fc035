// Compiles the PRODUCT's smb device code (gym_pcgrl_b200/csrc/pcgrl_smb.cuh, plain __host__ __device__ functions) for
// the host so that tests/test_smb_device_code_on_host.py can check it against the reference's golden vectors without
// a GPU.  Test harness only: the product path itself runs these functions inside k_smb_get_stats.
#include <stdlib.h>

#include "../gym_pcgrl_b200/csrc/pcgrl_smb.cuh"

extern "C" void smb_device_code_get_stats(const uint8_t* maps, int n, int w, int h, int power, int32_t* out, int out_stride) {
  using namespace pcgrl_smb;
  uint32_t* base = (uint32_t*)malloc(sizeof(uint32_t) * workspace_words(power));
  uint32_t solid_words[MAX_H * ROW_WORDS];
  Workspace ws;
  ws.heap = base;
  ws.nodes = base + heap_words(power);
  ws.visited = ws.nodes + node_words(power);
  for (int i = 0; i < n; i++) {
    int32_t st[8];
    get_stats_one(maps + (size_t)i * w * h, w, h, power, solid_words, ws, st);
    for (int k = 0; k < out_stride; k++) out[(size_t)i * out_stride + k] = (k < 8) ? st[k] : 0;
  }
  free(base);
}
