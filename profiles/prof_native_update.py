#!/usr/bin/env python
"""Launch sequence of the hand-written optimizer step for ncu: one rollout of 65 536 x 128 GTO environments, advantages, then
STEPS optimizer steps (mgrl_ppo_gradients + mgrl_ppo_apply) on 262 144-sample minibatches.
   ncu --metrics gpu__time_duration.sum --clock-control none --csv -k regex:'gemm|wgrad|gru|loss|conv1|prep|assemble|grad|adam|pack|moments' ..."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import minigrid_rl_b200 as mg  # noqa: E402

n = int(os.environ.get("PROF_ENVS", "65536"))
T = int(os.environ.get("PROF_T", "128"))
steps = int(os.environ.get("STEPS", "3"))
tf32 = os.environ.get("PROF_TF32", "1") == "1"
dev = torch.device("cuda", 0)
env = mg.DeviceEnv(mg.EnvConfig.for_task("GTO"), num_envs=n, seed=42, layout="hwc148")
cfg = mg.PPOConfig(n_steps=T, batch_size=n * T // 32, update_tf32=tf32)
eng = mg.RolloutEngine(env, mg.Policy(dev, seed=1), cfg, seed=1, keep_terminal_frames=False)
eng.collect(); eng.compute_advantages()
up = eng.updater
total = n * T
idx32 = torch.randperm(total, device=dev).to(torch.int32)
sums = up.moments(eng.buf, idx32, cfg.batch_size)
view = up.view(eng.buf)
ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps + 1)]
ev[0].record()
for k in range(steps):
    up.minibatch_native(view, idx32[k * cfg.batch_size:(k + 1) * cfg.batch_size], sums[k])
    ev[k + 1].record()
torch.cuda.synchronize()
print("ms per optimizer step:", [round(ev[k].elapsed_time(ev[k + 1]), 3) for k in range(steps)])
