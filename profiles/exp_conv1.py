import os, sys
R = os.path.dirname(os.path.dirname(os.path.abspath(__file__))); sys.path.insert(0, R)
import torch, torch.nn.functional as F
B = 262144; dev = "cuda"
torch.backends.cuda.matmul.allow_tf32 = True
def timeit(name, fn, n=3):
    fn(); torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(n): fn()
    b.record(); torch.cuda.synchronize(); print("%-52s %.2f ms" % (name, a.elapsed_time(b) / n))
x8 = torch.randint(0, 11, (B, 12, 7, 7), device=dev, dtype=torch.uint8)
w1 = torch.randn(16, 12, 2, 2, device=dev, requires_grad=True); b1 = torch.zeros(16, device=dev, requires_grad=True)
g = torch.randn(B, 16, 3, 3, device=dev)
for tf32 in (False, True):
    torch.backends.cudnn.allow_tf32 = tf32
    def conv1_cudnn():
        h = F.max_pool2d(F.relu(F.conv2d(x8.float() / 255, w1, b1)), 2)
        h.backward(g)
    timeit(f"cudnn conv1+relu+pool fwd+bwd(w only) tf32={tf32}", conv1_cudnn)
    xl = (x8.float() / 255).contiguous(memory_format=torch.channels_last)
    def conv1_cl():
        h = F.max_pool2d(F.relu(F.conv2d(xl, w1, b1)), 2)
        h.backward(g)
    timeit(f"  same, channels_last input precomputed tf32={tf32}", conv1_cl)
def conv1_unfold():
    xf = x8.float() / 255
    p = xf.unfold(2, 2, 1).unfold(3, 2, 1).permute(0, 2, 3, 1, 4, 5).reshape(B * 36, 48)
    h = torch.relu(F.linear(p, w1.reshape(16, 48), b1)).view(B, 3, 2, 3, 2, 16).amax(dim=(2, 4))
    h.backward(g.permute(0, 2, 3, 1))
timeit("unfold conv1+relu+pool fwd+bwd", conv1_unfold)
# half precision patches?
def conv1_unfold_bf16():
    xf = x8.to(torch.bfloat16)
    p = xf.unfold(2, 2, 1).unfold(3, 2, 1).permute(0, 2, 3, 1, 4, 5).reshape(B * 36, 48)
    h = torch.relu(F.linear(p.float() / 255, w1.reshape(16, 48), b1)).view(B, 3, 2, 3, 2, 16).amax(dim=(2, 4))
    h.backward(g.permute(0, 2, 3, 1))
timeit("unfold via bf16 patches (exact for 0..10)", conv1_unfold_bf16)
