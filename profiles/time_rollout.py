#!/usr/bin/env python
"""Where a policy-in-the-loop rollout spends its time: collect() wall vs device time, host cost of the two ctypes calls."""
import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import minigrid_rl_b200 as mg
n, T = 65536, 128
env = mg.DeviceEnv(mg.EnvConfig.for_task("GTO"), num_envs=n, seed=42, layout="hwc148")
eng = mg.RolloutEngine(env, mg.Policy("cuda", seed=1), mg.PPOConfig(n_steps=T, batch_size=n * T // 32), seed=1)
eng.collect(); torch.cuda.synchronize()
for rep in range(2):
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter(); a.record(); eng.collect(); b.record(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print("collect: device %.1f ms, host issue %.1f ms, wall %.1f ms" % (a.elapsed_time(b), 1e3 * (t1 - t0), 1e3 * (t2 - t0)))
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record(); nb = eng.bootstrap_truncated(); b.record(); torch.cuda.synchronize(); print("bootstrap %d samples: %.1f ms" % (nb, a.elapsed_time(b)))
a.record(); eng.compute_advantages(); b.record(); torch.cuda.synchronize(); print("gae: %.2f ms" % a.elapsed_time(b))
# the bench's sequence: an update between rollouts (weights change -> pack() + fragment packing inside collect)
eng.update(); eng.shift(); torch.cuda.synchronize()
for rep in range(2):
    e = [torch.cuda.Event(enable_timing=True) for _ in range(5)]
    e[0].record(); eng.policy.packed(); e[1].record(); eng.collect(); e[2].record(); eng.bootstrap_truncated(); e[3].record()
    eng.compute_advantages(); e[4].record(); torch.cuda.synchronize()
    print("after update: pack %.1f collect %.1f bootstrap %.1f gae %.2f ms" % tuple(e[k].elapsed_time(e[k + 1]) for k in range(4)))
    eng.update(); eng.shift(); torch.cuda.synchronize()
