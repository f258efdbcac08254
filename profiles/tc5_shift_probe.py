import ctypes as C, numpy as np, torch
lib = C.CDLL("minigrid-rl_b200/_lib/libmgrl.so")
out = torch.zeros(4, 128, 16, device="cuda")
rc = lib.mgrl_debug_tc5_shift_probe(C.c_void_p(out.data_ptr()), None)
torch.cuda.synchronize()
R, K, N = 144, 16, 16
A = np.array([[((r * 5 + k * 3) % 11) - 5 for k in range(K)] for r in range(R)], np.float64)
W = np.array([[((n * 7 + k) % 5) - 2 for k in range(K)] for n in range(N)], np.float64)
got = out.cpu().numpy()
for s, sh in enumerate([0, 1, 7, 8]):
    want = A[sh:sh + 128] @ W.T
    print("shift", sh, "max err", np.abs(got[s] - want).max(), "rc", rc)
