#!/usr/bin/env python
"""End-to-end learning check of the device-resident PPO path (what ppo.py:159 `model.learn` does in the reference):
   python profiles/train_task.py [TASK] [ENVS] [ITERS] [BATCH]
Prints per iteration the statistics of the episodes that finished during the rollout (stochastic policy) and, at the end,
a deterministic evaluation.  Reference hyper-parameters (hydra_configs/algorithm/ppo.yaml), horizon 128."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import minigrid_rl_b200 as mg

task = sys.argv[1] if len(sys.argv) > 1 else "GTG"
n = int(sys.argv[2]) if len(sys.argv) > 2 else 4096
iters = int(sys.argv[3]) if len(sys.argv) > 3 else 40
batch = int(sys.argv[4]) if len(sys.argv) > 4 else 8192
T = 128
env = mg.DeviceEnv(mg.EnvConfig.for_task(task), num_envs=n, seed=42, layout="hwc148")
cfg = mg.PPOConfig(n_steps=T, batch_size=batch, n_epochs=4, total_timesteps=n * T * iters)
eng = mg.RolloutEngine(env, mg.Policy("cuda", seed=42), cfg, seed=42)
t0 = time.time()
for it in range(iters):
    eng.collect()
    b = eng.buf
    done = b["ep_len"] != 0
    r = b["rewards"][done]
    stats = (int(done.sum()), float(r.mean()) if r.numel() else 0.0, float((r > 0).float().mean()) if r.numel() else 0.0,
             float(b["ep_len"][done].float().mean()) if r.numel() else 0.0)
    eng.bootstrap_truncated(); eng.compute_advantages()
    eng.updater.set_progress(1.0 - it / iters)
    eng.update(); eng.shift()
    if it % max(1, iters // 20) == 0 or it == iters - 1:
        print(f"iter {it:3d} frames {(it + 1) * n * T:>10d} episodes {stats[0]:7d} mean_reward {stats[1]:.3f} "
              f"success {stats[2]:.3f} mean_len {stats[3]:.1f}  ({time.time() - t0:.0f} s)", flush=True)
print("deterministic evaluation:", eng.evaluate(2, deterministic=True))
print("stochastic evaluation:", eng.evaluate(2, deterministic=False))
print("errors", env.error_flags())
