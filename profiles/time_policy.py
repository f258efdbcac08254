import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
import minigrid_rl_b200 as mg
n, T = 65536, 32
env = mg.DeviceEnv(mg.EnvConfig.for_task("GTO"), num_envs=n, seed=42, layout="hwc148")
eng = mg.RolloutEngine(env, mg.Policy("cuda", seed=1), mg.PPOConfig(n_steps=T, batch_size=n), seed=1, keep_terminal_frames=False)
eng.collect(); torch.cuda.synchronize()
B = eng.buf
a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
a.record()
for rep in range(3):
    for t in range(T):
        eng.policy.forward_rollout(B["frames"], B["dirs"], B["mission"][t + 3], t + 3, eng.prev_age, eng.prev_done, B["age"][t], B["values"][t], B["actions"][t], B["logp"][t])
b.record(); torch.cuda.synchronize(); print("simt", os.environ.get("MGRL_POLICY_SIMT"), "occ", os.environ.get("MGRL_TC_OCC"), "policy kernel us/step %.1f" % (1000 * a.elapsed_time(b) / (3 * T)))
