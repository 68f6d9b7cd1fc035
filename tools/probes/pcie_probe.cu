// Micro-probe: cost of scattered small stores / loads from a kernel to mapped pinned host memory (zero-copy)
// versus cudaMemcpyAsync of the full observation.  Build: nvcc -arch=sm_100a -O3 -o pcie_probe pcie_probe.cu
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
#include <chrono>

__global__ void k_scatter(uint8_t* map_h, uint8_t* heat_h, uint8_t* pos_h, double* rew_h, uint8_t* done_h,
                          const int* act_h, int n, int cells, int it, int* sink) {
  int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= n) return;
  int a = act_h[e];                       // 4-byte read from host memory per warp
  if (lane == 0) {
    int cell = (e * 7 + it * 13) % cells;
    if ((e + it) % 3 == 0) { map_h[(size_t)e * cells + cell] = (uint8_t)(a & 1); heat_h[(size_t)e * cells + cell] = (uint8_t)it; }
    pos_h[2 * e] = (uint8_t)(cell & 15); pos_h[2 * e + 1] = (uint8_t)(cell >> 4);
    rew_h[e] = (double)a;
    done_h[e] = (uint8_t)((e + it) % 97 == 0);
  }
  if ((e + it) % 97 == 0) {               // "reset": rewrite the whole map + heat of this env (coalesced)
    for (int i = lane; i < cells; i += 32) { map_h[(size_t)e * cells + i] = (uint8_t)(i & 1); heat_h[(size_t)e * cells + i] = 0; }
  }
  if (a == 12345) sink[0] = 1;
}

__global__ void k_device_only(uint8_t* map_d, int n, int cells, int it) {
  int e = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (e >= n) return;
  if (lane == 0) map_d[(size_t)e * cells + (e * 7 + it) % cells] = (uint8_t)it;
}

int main() {
  const int n = 4096, cells = 256, iters = 200;
  uint8_t *map_h, *heat_h, *pos_h, *done_h; double* rew_h; int* act_h; int* sink;
  cudaHostAlloc(&map_h, (size_t)n * cells, cudaHostAllocMapped); cudaHostAlloc(&heat_h, (size_t)n * cells, cudaHostAllocMapped);
  cudaHostAlloc(&pos_h, 2 * n, cudaHostAllocMapped); cudaHostAlloc(&done_h, n, cudaHostAllocMapped);
  cudaHostAlloc(&rew_h, 8 * n, cudaHostAllocMapped); cudaHostAlloc(&act_h, 4 * n, cudaHostAllocMapped);
  cudaMalloc(&sink, 4);
  uint8_t *map_d, *heat_d; cudaMalloc(&map_d, (size_t)n * cells); cudaMalloc(&heat_d, (size_t)n * cells);
  double* rew_d; cudaMalloc(&rew_d, 8 * n); int* act_d; cudaMalloc(&act_d, 4 * n);
  for (int i = 0; i < n; i++) act_h[i] = i & 3;
  cudaStream_t s; cudaStreamCreate(&s);
  auto now = [] { return std::chrono::high_resolution_clock::now(); };
  auto us = [](auto a, auto b) { return std::chrono::duration<double, std::micro>(b - a).count(); };
  for (int rep = 0; rep < 2; rep++) {
    // (1) zero-copy: kernel reads actions from host memory and writes deltas to host memory, then sync
    auto t0 = now();
    for (int it = 0; it < iters; it++) { act_h[it % n] = it & 3; k_scatter<<<n / 4, 128, 0, s>>>(map_h, heat_h, pos_h, rew_h, done_h, act_h, n, cells, it, sink); cudaStreamSynchronize(s); }
    auto t1 = now();
    // (2) copies: H2D actions, kernel, D2H map+heat+rew(+small), sync
    for (int it = 0; it < iters; it++) {
      cudaMemcpyAsync(act_d, act_h, 4 * n, cudaMemcpyHostToDevice, s);
      k_device_only<<<n / 4, 128, 0, s>>>(map_d, n, cells, it);
      cudaMemcpyAsync(map_h, map_d, (size_t)n * cells, cudaMemcpyDeviceToHost, s);
      cudaMemcpyAsync(heat_h, heat_d, (size_t)n * cells, cudaMemcpyDeviceToHost, s);
      cudaMemcpyAsync(rew_h, rew_d, 8 * n, cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
    }
    auto t2 = now();
    // (3) small-record copy: H2D actions, kernel, one 64 KB D2H, sync
    for (int it = 0; it < iters; it++) {
      cudaMemcpyAsync(act_d, act_h, 4 * n, cudaMemcpyHostToDevice, s);
      k_device_only<<<n / 4, 128, 0, s>>>(map_d, n, cells, it);
      cudaMemcpyAsync(map_h, map_d, 16 * n, cudaMemcpyDeviceToHost, s);
      cudaStreamSynchronize(s);
    }
    auto t3 = now();
    // (4) kernel + sync only
    for (int it = 0; it < iters; it++) { k_device_only<<<n / 4, 128, 0, s>>>(map_d, n, cells, it); cudaStreamSynchronize(s); }
    auto t4 = now();
    printf("rep %d: zero-copy %.1f us/step | full copies %.1f us/step | 64KB record copy %.1f us/step | kernel+sync %.1f us/step\n",
           rep, us(t0, t1) / iters, us(t1, t2) / iters, us(t2, t3) / iters, us(t3, t4) / iters);
  }
  printf("check %d %d\n", (int)map_h[7], (int)done_h[0]);
  return 0;
}
