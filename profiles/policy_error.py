#!/usr/bin/env python
"""Forward-kernel error on a rollout against the float64 evaluation of the oracle network (the yardstick of the 1e-5 bar of
tests/test_gpu_policy.py) and, for reference, the error of the oracle's own float32 evaluation against the same yardstick:
max |got - want| / max |want| per step for logits and values.
   python profiles/policy_error.py            (tensor-core kernel; MGRL_POLICY_SIMT=1 for the CUDA-core kernel)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from tests.test_gpu_policy import make_engine, oracle_stacks
from oracle import policy_oracle as po

n, T = 640, 20
eng, o = make_engine(n, T)
o64 = po.double_copy(o)
eng.collect()
b = eng.buf
stacks = oracle_stacks(b, T, n)
logits = torch.zeros((n, 7), device="cuda"); val = torch.zeros(n, device="cuda")
age = torch.zeros(n, dtype=torch.uint8, device="cuda")
worst = [0.0, 0.0, 0.0, 0.0]
for t in range(T + 1):
    prev_age = None if t == 0 else b["age"][t - 1]
    prev_done = None if t == 0 else b["ep_len"][t - 1]
    eng.policy.forward_rollout(b["frames"], b["dirs"], b["mission"][t + 3], t + 3, prev_age, prev_done, age, val,
                               logits=logits, seed=77, env_id_base=1000, step=t)
    img, d, mis = stacks[t]
    obs = {"direction": torch.from_numpy(d), "image": torch.from_numpy(img), "mission": torch.from_numpy(mis)}
    with torch.no_grad():
        lo, vo = o64(obs)
        l32, v32 = o(obs)
    el = float((logits.cpu().double() - lo).abs().max()) / float(lo.abs().max())
    ev = float((val.cpu().double() - vo).abs().max()) / float(vo.abs().max())
    el32 = float((l32.double() - lo).abs().max()) / float(lo.abs().max())
    ev32 = float((v32.double() - vo).abs().max()) / float(vo.abs().max())
    worst = [max(worst[0], el), max(worst[1], ev), max(worst[2], el32), max(worst[3], ev32)]
    if t < 3 or t == T:
        print("t %2d kernel: logits %.2e values %.2e | torch-CPU fp32: logits %.2e values %.2e (max |v| %.3f)"
              % (t, el, ev, el32, ev32, float(vo.abs().max())))
print("simt" if os.environ.get("MGRL_POLICY_SIMT") == "1" else "tensor",
      "worst vs float64: kernel logits %.2e values %.2e | torch-CPU fp32 logits %.2e values %.2e" % tuple(worst))
