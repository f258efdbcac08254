#!/usr/bin/env python
"""bench.py's PPO leg, segment by segment (device time): pack / collect / truncation bootstrap / GAE / update."""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import minigrid_rl_b200 as mg
n, T = 65536, 128
env = mg.DeviceEnv(mg.EnvConfig.for_task("GTO"), num_envs=n, seed=42, layout="hwc148")
cfg = mg.PPOConfig(n_steps=T, batch_size=n * T // 32, n_epochs=4, update_tf32=True)
eng = mg.RolloutEngine(env, mg.Policy("cuda", seed=42), cfg, seed=42)
eng.iteration(1.0); torch.cuda.synchronize()
for it in range(3):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(6)]
    e[0].record(); eng.policy.packed(); e[1].record(); eng.collect(); e[2].record(); nb = eng.bootstrap_truncated(); e[3].record()
    eng.compute_advantages(); e[4].record(); eng.update(); eng.shift(); e[5].record(); torch.cuda.synchronize()
    done = int((eng.buf["ep_len"] != 0).sum())
    print("it %d: pack %.1f collect %.1f bootstrap %.1f (%d samples) gae %.2f update %.1f ms; episodes finished %d" %
          ((it,) + tuple(e[k].elapsed_time(e[k + 1]) for k in range(2)) + (e[2].elapsed_time(e[3]), nb, e[3].elapsed_time(e[4]),
           e[4].elapsed_time(e[5]), done)))
# is the update bound by the host issuing its ~150 launches per minibatch?
import time
torch.cuda.synchronize(); t0 = time.perf_counter(); eng.update(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print("update: host issue %.1f ms, wall %.1f ms" % (1e3 * (t1 - t0), 1e3 * (t2 - t0)))
