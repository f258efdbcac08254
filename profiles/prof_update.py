import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from torch.profiler import profile, ProfilerActivity
import minigrid_rl_b200 as mg
from minigrid_rl_b200 import ppo
n, T = 65536, 128
env = mg.DeviceEnv(mg.EnvConfig.for_task("GTO"), num_envs=n, seed=42, layout="hwc148")
cfg = mg.PPOConfig(n_steps=T, batch_size=n * T // 32, update_tf32=len(sys.argv) > 1)
eng = mg.RolloutEngine(env, mg.Policy("cuda", seed=1), cfg, seed=1, keep_terminal_frames=False)
eng.collect(); eng.compute_advantages()
B = eng.buf
idx = torch.randperm(n * T, device="cuda")[:cfg.batch_size]
t, i = idx // n, idx % n
def one():
    eng.updater.minibatch(None, None, None, B["actions"][t, i], B["values"][t, i], B["logp"][t, i], B["adv"][t, i], B["ret"][t, i],
                          samples=(B, t, i))
one(); one(); torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    one(); torch.cuda.synchronize()
print(prof.key_averages().table(sort_by="cuda_time_total", row_limit=60, max_name_column_width=70))
