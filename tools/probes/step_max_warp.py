"""Which warp sets the duration of a single-step launch?  -DPCGRL_PROFILE build (prebuilt build/libpcgrl_profile.so if present):
per launch, the slowest warp of each class (no change / changed / auto-reset, clock64 from kernel entry to the end of the
epilogue) next to the launch's own duration (CUDA events)."""
import os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
import numpy as np, torch
from gym_pcgrl_b200 import _native, build as B
so = os.path.join(ROOT, "build", "libpcgrl_profile.so")
if not os.path.exists(so):
    os.makedirs(os.path.dirname(so), exist_ok=True)
    subprocess.check_call(["nvcc"] + B.NVCC_FLAGS + ["-DPCGRL_PROFILE", "-o", so] + B.SOURCES, cwd=B.CSRC)
_native.LIB_PATH = so
import bench
n = 4096
env = bench.make_env(n, "cuda:0", 0)
env._ensure_buffers()
env._tens["status"] = torch.zeros(2 * 48, dtype=torch.int32, device="cuda")
env._cbufs.status = env._tens["status"].data_ptr()
env.reset()
steps = 400
acts = torch.from_numpy(bench.host_actions(env, 512 + steps, n, 5)).cuda()
for t in range(512):
    env.step(acts[t])
torch.cuda.synchronize()
rows = []
st64 = env._tens["status"].view(torch.int64)
for t in range(steps):
    st64.zero_()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    env.step(acts[512 + t])
    e1.record()
    torch.cuda.synchronize()
    full = st64.cpu().numpy()
    mx = [int(full[16 + 8 * c + 5]) for c in range(3)]      # slowest warp (cycles) per class in this launch
    cnt = [int(full[16 + 8 * c]) for c in range(3)]
    rows.append((e0.elapsed_time(e1) * 1e3, mx[0], mx[1], mx[2], cnt[2]))
a = np.array(rows, dtype=np.float64)
clk = 1.965e3   # cycles per us at the sampled SM clock
print("launches %d: duration median %.1f us (p10 %.1f, p90 %.1f)" % (len(a), np.median(a[:, 0]), np.percentile(a[:, 0], 10), np.percentile(a[:, 0], 90)))
for c, nm in enumerate(["no change", "changed", "auto-reset"]):
    v = a[:, 1 + c] / clk
    print("slowest %-10s warp per launch: median %.1f us (p10 %.1f, p90 %.1f)" % (nm, np.median(v), np.percentile(v, 10), np.percentile(v, 90)))
slow = np.maximum(np.maximum(a[:, 1], a[:, 2]), a[:, 3]) / clk
print("slowest warp of the launch: median %.1f us; launch duration - slowest warp: median %.1f us" % (np.median(slow), np.median(a[:, 0] - slow)))
print("launches whose slowest warp is an auto-reset: %.0f %%; resets per launch: mean %.1f" % (100.0 * np.mean(a[:, 3] >= np.maximum(a[:, 1], a[:, 2])), a[:, 4].mean()))
print("correlation(duration, slowest warp) = %.2f" % np.corrcoef(a[:, 0], slow)[0, 1])
