"""200 pcgrl_step_host calls in one transport mode (argv[1]: delta | direct) -- run under
`ncu --metrics gpu__time_duration.sum -k regex:k_rollout` to get the step kernel's duration in that mode."""
import os, sys, types
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
import bench
from gym_pcgrl_b200 import HostStepIO
mode = sys.argv[1]
n = 4096
env = bench.make_env(n, "cuda:0", 0)
env.reset()
B = bench.Bench(types.SimpleNamespace(), 0, 1, 0)
B.preroll(env, 512, 77)
io = HostStepIO(env, with_obs=True, with_info=False, mode=mode)
acts = torch.from_numpy(bench.host_actions(env, 200, n, 7)).pin_memory()
base, stride = acts.data_ptr(), acts.stride(0) * 4
for t in range(200):
    io.struct.actions = base + t * stride
    env.step_host(io)
torch.cuda.synchronize()
