#!/usr/bin/env python
"""Where a PPO minibatch update spends its time (gather / forward / backward / optimizer) and the per-step cost of the
rollout's two kernels.  python profiles/time_update.py [tf32]"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import minigrid_rl_b200 as mg
from minigrid_rl_b200 import ppo
strict = not (len(sys.argv) > 1 and sys.argv[1] == "tf32")
n, T = 65536, 128
env = mg.DeviceEnv(mg.EnvConfig.for_task("GTO"), num_envs=n, seed=42, layout="hwc148")
cfg = mg.PPOConfig(n_steps=T, batch_size=n * T // 32)
eng = mg.RolloutEngine(env, mg.Policy("cuda", seed=1, fp32_strict=strict), cfg, seed=1)
def ev(): return torch.cuda.Event(enable_timing=True)
eng.collect(); torch.cuda.synchronize()
a, b_ = ev(), ev(); a.record(); eng.collect(); b_.record(); torch.cuda.synchronize()
print("rollout ms", a.elapsed_time(b_), "per step us", 1000 * a.elapsed_time(b_) / T)
# policy kernel alone / env kernel alone
B = eng.buf
a.record()
for t in range(T):
    eng.policy.forward_rollout(B["frames"], B["dirs"], B["mission"][t + 3], t + 3, eng.prev_age, eng.prev_done, B["age"][t], B["values"][t], B["actions"][t], B["logp"][t])
b_.record(); torch.cuda.synchronize(); print("policy kernel us/step", 1000 * a.elapsed_time(b_) / T)
eng.compute_advantages()
idx = torch.randperm(n * T, device="cuda")[:cfg.batch_size]
t, i = idx // n, idx % n
for rep in range(2):
    e = [ev() for _ in range(5)]
    e[0].record(); image, onehot, mrow = ppo.gather_minibatch(B, t, i)
    args = (B["actions"][t, i], B["values"][t, i], B["logp"][t, i], B["adv"][t, i], B["ret"][t, i])
    e[1].record(); loss, _ = ppo.ppo_minibatch_loss(eng.policy, cfg, image, onehot, mrow, *args)
    e[2].record(); eng.updater.opt.zero_grad(); loss.backward()
    e[3].record(); torch.nn.utils.clip_grad_norm_(eng.updater.params, cfg.max_grad_norm); eng.updater.opt.step()
    e[4].record(); torch.cuda.synchronize()
    print("strict" if strict else "tf32", "gather %.2f fwd %.2f bwd %.2f opt %.2f ms" % tuple(e[k].elapsed_time(e[k + 1]) for k in range(4)))
