import os, sys, time
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch, torch.nn.functional as F
torch.backends.cudnn.allow_tf32 = False
B = 262144
dev = "cuda"
def timeit(name, fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); print("%-40s %.2f ms" % (name, a.elapsed_time(b) / n))
lut = torch.randn(296, 128, device=dev, requires_grad=True)
mrow = torch.randint(0, 296, (B,), device=dev)
g = torch.randn(B, 128, device=dev)
timeit("lut[mrow] fwd+bwd", lambda: lut[mrow].backward(g))
timeit("F.embedding fwd+bwd", lambda: F.embedding(mrow, lut).backward(g))
timeit("one_hot matmul fwd+bwd", lambda: (F.one_hot(mrow, 296).float() @ lut).backward(g))
x = torch.randint(0, 11, (B, 12, 7, 7), device=dev, dtype=torch.uint8)
w1 = torch.randn(16, 12, 2, 2, device=dev, requires_grad=True); b1 = torch.zeros(16, device=dev, requires_grad=True)
w2 = torch.randn(32, 16, 2, 2, device=dev, requires_grad=True); b2 = torch.zeros(32, device=dev, requires_grad=True)
w3 = torch.randn(64, 32, 2, 2, device=dev, requires_grad=True); b3 = torch.zeros(64, device=dev, requires_grad=True)
def conv_cudnn():
    h = F.max_pool2d(F.relu(F.conv2d(x.float() / 255, w1, b1)), 2)
    h = F.relu(F.conv2d(h, w2, b2)); h = F.relu(F.conv2d(h, w3, b3)).flatten(1)
    return h
go = torch.randn(B, 64, device=dev)
timeit("conv stack cudnn fwd", lambda: conv_cudnn())
timeit("conv stack cudnn fwd+bwd", lambda: conv_cudnn().backward(go))
def conv_unfold():
    xf = x.float() / 255                                           # [B,12,7,7]
    p = xf.unfold(2, 2, 1).unfold(3, 2, 1)                          # [B,12,6,6,2,2]
    p = p.permute(0, 2, 3, 1, 4, 5).reshape(B * 36, 48)
    h = torch.relu(p @ w1.reshape(16, 48).t() + b1).view(B, 6, 6, 16)
    h = h.view(B, 3, 2, 3, 2, 16).amax(dim=(2, 4))                   # maxpool -> [B,3,3,16]
    p = h.unfold(1, 2, 1).unfold(2, 2, 1)                           # [B,2,2,16,2,2]
    p = p.reshape(B * 4, 64)
    h = torch.relu(p @ w2.permute(0, 1, 2, 3).reshape(32, 64).t() + b2).view(B, 128)   # (oh,ow,c2)
    h = torch.relu(h @ w3.permute(0, 2, 3, 1).reshape(64, 128).t() + b3)
    return h
timeit("conv stack unfold fwd", lambda: conv_unfold())
timeit("conv stack unfold fwd+bwd", lambda: conv_unfold().backward(go))
xl = x.float().contiguous(memory_format=torch.channels_last)
def conv_cl():
    h = F.max_pool2d(F.relu(F.conv2d(xl / 255, w1, b1)), 2)
    h = F.relu(F.conv2d(h, w2, b2)); h = F.relu(F.conv2d(h, w3, b3)).flatten(1)
    return h
timeit("conv stack channels_last fwd+bwd", lambda: conv_cl().backward(go))
f = torch.randn(B, 208, device=dev)
wp = torch.randn(64, 208, device=dev, requires_grad=True); wq = torch.randn(64, 64, device=dev, requires_grad=True)
timeit("mlp 208-64-64 fwd+bwd", lambda: torch.tanh(F.linear(torch.tanh(F.linear(f, wp)), wq)).sum().backward())
