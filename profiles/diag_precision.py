import os, sys
sys.path.insert(0, '/root/repo')
import torch, numpy as np
from tests.test_gpu_policy import make_engine, oracle_stacks
from oracle import policy_oracle as po
n, T = 640, 6
eng, o = make_engine(n, T)
o64 = po.double_copy(o)
eng.collect()
b = eng.buf
stacks = oracle_stacks(b, T, n)
pol = eng.policy
def run(tag):
    logits = torch.zeros((n, 7), device="cuda"); val = torch.zeros(n, device="cuda")
    age = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for t in (0, 3):
        prev_age = None if t == 0 else b["age"][t - 1]
        prev_done = None if t == 0 else b["ep_len"][t - 1]
        pol.forward_rollout(b["frames"], b["dirs"], b["mission"][t + 3], t + 3, prev_age, prev_done, age, val, logits=logits)
        img, d, mis = stacks[t]
        obs = {"direction": torch.from_numpy(d), "image": torch.from_numpy(img), "mission": torch.from_numpy(mis)}
        with torch.no_grad():
            lo, vo = o64(obs)
        print(tag, "t", t, "logits abs %.2e rel %.2e | values abs %.2e rel %.2e" % (
            float((logits.cpu().double()-lo).abs().max()), float((logits.cpu().double()-lo).abs().max()/lo.abs().max()),
            float((val.cpu().double()-vo).abs().max()), float((val.cpu().double()-vo).abs().max()/vo.abs().max())))
run("tc kernel, cuDNN fp32 LUT   ")
# LUT from the float64 GRU
packed = pol.packed()
seq = pol.sequences.cpu()
with torch.no_grad():
    _, h = o64.gru(o64.embedding(seq))
off, size = __import__("minigrid_rl_b200").policy.WEIGHT_LAYOUT["LUT"]
lut32 = packed[off:off+size].clone()
lut64 = h[-1].float().reshape(-1).cuda()
print("LUT fp32(cuDNN) vs fp64: max abs %.2e" % float((lut32 - lut64).abs().max()))
packed[off:off+size] = lut64
run("tc kernel, float64 LUT      ")
pol.tensor_cores = False
run("fp32 CUDA-core kernel, f64 LUT")
packed[off:off+size] = lut32
run("fp32 CUDA-core kernel, cuDNN LUT")
