#!/usr/bin/env python
"""Fixed launch sequence for ncu (deterministic -s/-c): reset, WARM rollout launches, then the launches to capture.
   python profiles/prof_target.py many   -> rollout_kernel launches: 3 warm-up + 2 (T=128 steps per launch)
   python profiles/prof_target.py gae    -> gae_kernel launches: 3 warm-up + 2 ([128, 65536] buffers)
   python profiles/prof_target.py step   -> step_kernel launches: 384 warm-up + 128 (one step per launch)
   python profiles/prof_target.py ppo    -> 2 rollouts of 32 steps: policy_forward_kernel + step_kernel per step"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import minigrid_rl_b200 as mg  # noqa: E402

mode = sys.argv[1] if len(sys.argv) > 1 else "many"
n = int(os.environ.get("PROF_ENVS", "65536"))
T = 128
dev = torch.device("cuda", 0)
env = mg.DeviceEnv(mg.EnvConfig.for_task("GTO"), num_envs=n, seed=42, layout="hwc148")
env.reset()
g = torch.Generator(device=dev).manual_seed(1234)
actions = torch.randint(0, 7, (T, n), dtype=torch.uint8, device=dev, generator=g)
u8 = dict(dtype=torch.uint8, device=dev)
image = torch.empty((T, n, 148), **u8)
dirs = torch.empty((T, n), **u8); mis = torch.empty((T, n), **u8)
rew = torch.empty((T, n), dtype=torch.float32, device=dev)
term = torch.empty((T, n), **u8); trunc = torch.empty((T, n), **u8); eplen = torch.empty((T, n), **u8)
if mode == "ppo":      # policy_forward_kernel + step_kernel, two launches per rollout step
    del image
    eng = mg.RolloutEngine(env, mg.Policy(dev, seed=1), mg.PPOConfig(n_steps=32, batch_size=n * 32 // 8), seed=1,
                           keep_terminal_frames=False)
    for _ in range(2):
        eng.collect()
elif mode == "gae":    # gae_kernel: 3 warm-up + 2 launches on bench-sized [T, N] buffers
    r = torch.rand(T, n, device=dev); v = torch.randn(T, n, device=dev)
    es = (torch.rand(T, n, device=dev) < 0.14).to(torch.uint8)
    lv = torch.randn(n, device=dev); ld = torch.zeros(n, dtype=torch.uint8, device=dev)
    for _ in range(5):
        mg.vec_env.gae(r, v, es, lv, ld, 0.81, 0.945)
elif mode == "many":
    for _ in range(5):
        env.step_many(actions, image, dirs, mis, rew, term, trunc, eplen)
else:
    for _ in range(4):
        for t in range(T):
            env.step(actions[t], image[t], dirs[t], mis[t], rew[t], term[t], trunc[t], eplen[t])
torch.cuda.synchronize()
print("errors", env.error_flags())
