import sys, torch
sys.path.insert(0, "/root/repo")
from gym_pcgrl_b200.models import ActorCritic
from gym_pcgrl_b200.policy_native import NativePolicy
net = ActorCritic("FullyConvPolicyBigMap", (14, 14, 1), 392).cuda()
pol = NativePolicy(net)
obs = torch.randint(0, 2, (4096, 14, 14, 1), dtype=torch.uint8, device="cuda")
for _ in range(3): pol(obs)
torch.cuda.synchronize()
from torch.profiler import profile, ProfilerActivity
with profile(activities=[ProfilerActivity.CUDA]) as prof:
    for _ in range(5): pol(obs)
    torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=12, max_name_column_width=70))
