"""Time of the first extractor stage of the update (262 144 samples off a [132, 65536, 148] frame buffer): mma.sync kernel,
tcgen05 kernel (two-term split / one TF32 pass).  Run under gpurun."""
import ctypes as C
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from minigrid_rl_b200 import _native as nat  # noqa: E402

n, T, B = 65536, 128, 262144
dev = "cuda"
frames = torch.randint(0, 11, (T + 4, n, 148), dtype=torch.uint8, device=dev)
g = torch.Generator(device=dev).manual_seed(0)
order = os.environ.get("ORDER", "random")
sel = torch.randperm(T * n, device=dev, generator=g)[:B] if order == "random" else torch.arange(B, device=dev)
t, i = (sel // n).to(torch.int32), (sel % n).to(torch.int32)
age = torch.randint(0, 4, (B,), dtype=torch.uint8, device=dev)
w1 = torch.randn(16, 48, device=dev) * 0.3
b1 = torch.randn(16, device=dev) * 0.1
pooled = torch.empty(B, 9, 16, device=dev)
arg = torch.empty(B, 9, 16, dtype=torch.uint8, device=dev)
s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
for mode in (None, "1", "2"):
    if mode is None:
        os.environ.pop("MGRL_CONV1_TC5", None)
    else:
        os.environ["MGRL_CONV1_TC5"] = mode
    call = lambda: nat.check(nat.lib().mgrl_conv1_pool_forward(p(frames), n, p(t), p(i), p(age), B, p(w1), p(b1), p(pooled), p(arg), s), "fwd")  # noqa: E731
    for _ in range(3):
        call()
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
    ev[0].record()
    for _ in range(20):
        call()
    ev[1].record()
    torch.cuda.synchronize()
    us = ev[0].elapsed_time(ev[1]) * 1000 / 20
    print(f"[{order} samples] conv1 forward, {B} samples, mode {mode or 'mma.sync (two-term)'}: {us:.1f} us  ({(B * 592 + B * 144 * 5) / us / 1e3:.0f} GB/s)")
