#!/usr/bin/env python
"""Forward-kernel error against the torch-CPU fp32 oracle on a rollout (the measurement behind the tolerances of
tests/test_gpu_policy.py): max |got - want| / max |want| per step for logits and values.
   python profiles/policy_error.py            (tensor-core kernel; MGRL_POLICY_SIMT=1 for the CUDA-core kernel)"""
import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch
from tests.test_gpu_policy import make_engine, oracle_stacks

n, T = 640, 20
eng, o = make_engine(n, T)
eng.collect()
b = eng.buf
stacks = oracle_stacks(b, T, n)
logits = torch.zeros((n, 7), device="cuda"); val = torch.zeros(n, device="cuda")
age = torch.zeros(n, dtype=torch.uint8, device="cuda")
worst = [0.0, 0.0]
for t in range(T + 1):
    prev_age = None if t == 0 else b["age"][t - 1]
    prev_done = None if t == 0 else b["ep_len"][t - 1]
    eng.policy.forward_rollout(b["frames"], b["dirs"], b["mission"][t + 3], t + 3, prev_age, prev_done, age, val,
                               logits=logits, seed=77, env_id_base=1000, step=t)
    img, d, mis = stacks[t]
    with torch.no_grad():
        lo, vo = o({"direction": torch.from_numpy(d), "image": torch.from_numpy(img), "mission": torch.from_numpy(mis)})
    el = float((logits.cpu() - lo).abs().max()) / float(lo.abs().max())
    ev = float((val.cpu() - vo).abs().max()) / float(vo.abs().max())
    worst = [max(worst[0], el), max(worst[1], ev)]
    if t < 3 or t == T:
        print("t %2d logits %.2e values %.2e (max |v| %.3f)" % (t, el, ev, float(vo.abs().max())))
print("simt" if os.environ.get("MGRL_POLICY_SIMT") == "1" else "tensor", "worst logits %.2e values %.2e" % tuple(worst))
