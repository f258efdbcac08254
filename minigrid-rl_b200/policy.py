"""Policy of the reference (CustomPPOPolicy over CustomExtractor) held as plain torch tensors.

Mirrors /root/reference/src/policies.py:21-120 (extractor built from hydra_configs/single.yaml:38-57 with
n_frames_stack = 4), policies.py:227-257 (CustomPPOPolicy, init_weights) and the parts of SB3's
ActorCriticPolicy it inherits (MlpExtractor pi/vf 208->64->64 Tanh, action_net, value_net, Categorical).
Parameter names are SB3's state_dict names so checkpoints can travel both ways.

Two consumers share the same parameter tensors:
  * the rollout: `pack()` lays the weights out for the hand-written CUDA forward kernel
    (csrc/mgrl_policy.cu, `mgrl_policy_forward`), including the mission look-up table;
  * the PPO update: `evaluate()` is the differentiable torch expression of the same network (autograd).
The GRU over the stacked mission tokens is evaluated on the 74 x 4 distinct (mission, frames-in-stack)
sequences only (the mission is constant within an episode), never per observation.
"""
from __future__ import annotations

import contextlib
import ctypes as C
import math
import os

import numpy as np

from . import _native as nat
from .missions import N_MISSIONS, token_table

N_ACTIONS = 7
N_WEIGHTS = 84940
N_FRAGMENTS = 93696            # MGRL_POLICY_FRAGMENTS: tf32 hi/lo mma fragments of the tensor-core forward kernel
FLAG_DETERMINISTIC, FLAG_TENSOR = 1, 2
# float offsets of the packed buffer, identical to csrc/mgrl_policy.cu
_L = {}
_off = 0
for _name, _size in [("W1", 48 * 16), ("B1", 16), ("W2", 64 * 32), ("B2", 32), ("W3", 128 * 64), ("B3", 64),
                     ("WD", 16 * 16), ("BD", 16), ("PI1", 208 * 64), ("PI1B", 64), ("PI2", 64 * 64), ("PI2B", 64),
                     ("VF1", 208 * 64), ("VF1B", 64), ("VF2", 64 * 64), ("VF2B", 64), ("WA", 64 * 8), ("BA", 8),
                     ("WV", 64), ("BV", 4), ("LUT", N_MISSIONS * 4 * 128)]:
    _L[_name] = (_off, _size)
    _off += _size
assert _off == N_WEIGHTS
WEIGHT_LAYOUT = dict(_L)

_PREFIX = "features_extractor.extractors."
SHAPES = {
    _PREFIX + "direction.direction_Linear_0.weight": (16, 16),
    _PREFIX + "direction.direction_Linear_0.bias": (16,),
    _PREFIX + "image.image_Conv2d_0.weight": (16, 12, 2, 2),
    _PREFIX + "image.image_Conv2d_0.bias": (16,),
    _PREFIX + "image.image_Conv2d_3.weight": (32, 16, 2, 2),
    _PREFIX + "image.image_Conv2d_3.bias": (32,),
    _PREFIX + "image.image_Conv2d_5.weight": (64, 32, 2, 2),
    _PREFIX + "image.image_Conv2d_5.bias": (64,),
    _PREFIX + "mission.mission_Embedding_0.weight": (32, 32),
    _PREFIX + "mission.mission_GRU_1.weight_ih_l0": (384, 32),
    _PREFIX + "mission.mission_GRU_1.weight_hh_l0": (384, 128),
    _PREFIX + "mission.mission_GRU_1.bias_ih_l0": (384,),
    _PREFIX + "mission.mission_GRU_1.bias_hh_l0": (384,),
    "mlp_extractor.policy_net.0.weight": (64, 208), "mlp_extractor.policy_net.0.bias": (64,),
    "mlp_extractor.policy_net.2.weight": (64, 64), "mlp_extractor.policy_net.2.bias": (64,),
    "mlp_extractor.value_net.0.weight": (64, 208), "mlp_extractor.value_net.0.bias": (64,),
    "mlp_extractor.value_net.2.weight": (64, 64), "mlp_extractor.value_net.2.bias": (64,),
    "action_net.weight": (7, 64), "action_net.bias": (7,),
    "value_net.weight": (1, 64), "value_net.bias": (1,),
}
N_PARAMS = sum(int(np.prod(s)) for s in SHAPES.values())      # 110 216


_CONV1_POOL = None


def _conv1_pool(torch):
    """autograd wrapper of the hand-written first extractor stage (built lazily so the module imports without torch)"""
    global _CONV1_POOL
    if _CONV1_POOL is not None:
        return _CONV1_POOL

    class Conv1Pool(torch.autograd.Function):
        @staticmethod
        def forward(ctx, w1, b1, frames, t32, i32, age):
            B, n = int(t32.numel()), int(frames.shape[1])
            pooled = torch.empty((B, 9, 16), dtype=torch.float32, device=frames.device)
            arg = torch.empty((B, 9, 16), dtype=torch.uint8, device=frames.device)
            age = age.contiguous()
            s = C.c_void_p(torch.cuda.current_stream(frames.device).cuda_stream)
            p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
            nat.check(nat.lib().mgrl_conv1_pool_forward(p(frames), n, p(t32), p(i32), p(age), B, p(w1.detach().contiguous()),
                                                        p(b1.detach().contiguous()), p(pooled), p(arg), s), "conv1_pool_forward")
            ctx.save_for_backward(frames, t32, i32, age, arg)
            return pooled

        @staticmethod
        def backward(ctx, g):
            frames, t32, i32, age, arg = ctx.saved_tensors
            B, n = int(t32.numel()), int(frames.shape[1])
            dw1 = torch.empty((16, 48), dtype=torch.float32, device=frames.device)
            db1 = torch.empty(16, dtype=torch.float32, device=frames.device)
            s = C.c_void_p(torch.cuda.current_stream(frames.device).cuda_stream)
            p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
            nat.check(nat.lib().mgrl_conv1_pool_backward(p(frames), n, p(t32), p(i32), p(age), B, p(arg), p(g.contiguous()),
                                                         p(dw1), p(db1), s), "conv1_pool_backward")
            return dw1.view(16, 12, 2, 2), db1, None, None, None, None

    _CONV1_POOL = Conv1Pool.apply
    return _CONV1_POOL


_SPLITK_LINEAR = None
_PATCH2 = None
_LUT_ROWS = None
NATIVE_MIN_ROWS = 16384        # minibatches at least this deep take the hand-written update kernels below (tests raise it
                               # to get the pure library formulation of the same network as a reference)


def _lut_rows(torch, lut, rows):
    """lut[rows] whose backward is the hand-written segmented sum mgrl_lut_grad (CUDA, large minibatches); F.embedding
    otherwise."""
    if not lut.is_cuda or rows.numel() < NATIVE_MIN_ROWS or not torch.is_grad_enabled():
        return torch.nn.functional.embedding(rows, lut)
    global _LUT_ROWS
    if _LUT_ROWS is None:
        class LutRows(torch.autograd.Function):
            @staticmethod
            def forward(ctx, lut, rows):
                ctx.save_for_backward(rows)
                ctx.n_rows = int(lut.shape[0])
                return lut.index_select(0, rows)

            @staticmethod
            def backward(ctx, g):
                (rows,) = ctx.saved_tensors
                g = g.contiguous()
                out = torch.empty((ctx.n_rows, 128), dtype=torch.float32, device=g.device)
                s = C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
                nat.check(nat.lib().mgrl_lut_grad(C.c_void_p(g.data_ptr()), C.c_void_p(rows.data_ptr()), int(g.shape[0]), ctx.n_rows,
                                                  C.c_void_p(out.data_ptr()), s), "mgrl_lut_grad")
                return out, None

        _LUT_ROWS = LutRows.apply
    return _LUT_ROWS(lut, rows.long().contiguous())


def _patches2x2(torch, h):
    """h [B,3,3,16] (qh, qw, c) -> conv2 input patches [B*4, 64] in (kh, kw, ci) order: hand-written gather and its adjoint
    (mgrl_patch2x2_forward / backward) on CUDA, the same gather by slicing elsewhere."""
    B = h.shape[0]
    if not h.is_cuda:
        return torch.stack([h[:, kh:kh + 2, kw:kw + 2, :] for kh in (0, 1) for kw in (0, 1)], dim=3).reshape(B * 4, 64)
    global _PATCH2
    if _PATCH2 is None:
        class Patch2x2(torch.autograd.Function):
            @staticmethod
            def forward(ctx, pooled):
                b = int(pooled.shape[0])
                out = torch.empty((b * 4, 64), dtype=torch.float32, device=pooled.device)
                s = C.c_void_p(torch.cuda.current_stream(pooled.device).cuda_stream)
                nat.check(nat.lib().mgrl_patch2x2_forward(C.c_void_p(pooled.data_ptr()), b, C.c_void_p(out.data_ptr()), s),
                          "mgrl_patch2x2_forward")
                return out

            @staticmethod
            def backward(ctx, g):
                g = g.contiguous()
                b = int(g.shape[0]) // 4
                d = torch.empty((b, 3, 3, 16), dtype=torch.float32, device=g.device)
                s = C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)
                nat.check(nat.lib().mgrl_patch2x2_backward(C.c_void_p(g.data_ptr()), b, C.c_void_p(d.data_ptr()), s),
                          "mgrl_patch2x2_backward")
                return d

        _PATCH2 = Patch2x2.apply
    return _PATCH2(h.contiguous().float())


def _linear(torch, x, w, b):
    """F.linear whose weight gradient is a split-K batched GEMM.  The update's minibatches are 10^5..10^6 rows deep and
    the layers at most 208 x 128 wide: as ONE `dz.T @ x` GEMM the weight gradient has a handful of output tiles and the
    library runs it on a few CTAs of a 148-SM GPU (1.4 ms of an 8 ms minibatch); cut into 64 row blocks it is a batched
    GEMM that fills the machine, plus a sum over 64 small partial results."""
    if not x.is_cuda or x.shape[0] < NATIVE_MIN_ROWS or x.shape[0] % 64 != 0 or not torch.is_grad_enabled():
        return torch.nn.functional.linear(x, w, b)
    global _SPLITK_LINEAR
    if _SPLITK_LINEAR is None:
        class SplitKLinear(torch.autograd.Function):
            @staticmethod
            def forward(ctx, x, w, b):
                ctx.save_for_backward(x, w)
                return torch.addmm(b, x, w.t())

            @staticmethod
            def backward(ctx, g):
                x, w = ctx.saved_tensors
                g = g.contiguous()
                m, s = g.shape[0], 64
                dx = g @ w if ctx.needs_input_grad[0] else None
                dw = torch.bmm(g.view(s, m // s, -1).transpose(1, 2), x.view(s, m // s, -1)).sum(0)
                db = torch.empty(g.shape[1], dtype=torch.float32, device=g.device)     # one streaming pass (mgrl_colsum)
                nat.check(nat.lib().mgrl_colsum(C.c_void_p(g.data_ptr()), m, int(g.shape[1]), C.c_void_p(db.data_ptr()),
                                                C.c_void_p(torch.cuda.current_stream(g.device).cuda_stream)), "mgrl_colsum")
                return dx, dw, db

        _SPLITK_LINEAR = SplitKLinear.apply
    return _SPLITK_LINEAR(x.contiguous(), w, b)


class Policy:
    """Parameters + the two evaluation paths.  Works on any torch device for `evaluate` (the update's autograd
    path and the CPU tests of the data-parallel logic); `pack` / `forward_rollout` need CUDA."""

    def __init__(self, device="cuda", seed: int = 0, fp32_strict: bool = True):
        import torch
        self.torch = torch
        self.fp32_strict = fp32_strict      # False = allow TF32 in the library kernels of the update, like ppo.py:29-32
        # rollout forward: split-TF32 tensor-core kernel (fp32-class results); MGRL_POLICY_SIMT=1 selects the fp32
        # CUDA-core kernel it replaced (kept for A/B measurements)
        self.tensor_cores = os.environ.get("MGRL_POLICY_SIMT", "0") != "1"
        self.capturing = False              # set by the updater while it captures its optimizer step into a CUDA graph
        self._leaves = None                 # see fresh_leaves()
        self.device = torch.device(device)
        g = torch.Generator().manual_seed(seed)
        self.params = {}
        for name, shape in SHAPES.items():
            self.params[name] = self._init(name, shape, g).to(self.device).requires_grad_(True)
        # the GRU runs through torch.nn.GRU (cuDNN on the device) on the 296 distinct sequences; the module's
        # parameters ARE the canonical tensors
        self._gru = torch.nn.GRU(32, 128, 1, True, True).to(self.device)
        for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            key = _PREFIX + "mission.mission_GRU_1." + n
            with torch.no_grad():
                getattr(self._gru, n).copy_(self.params[key])
            self.params[key] = getattr(self._gru, n)
        # stacked mission token sequences: row m*4 + age = [zeros * (3-age) frames | tokens(m) * (age+1) frames]
        tt = torch.from_numpy(token_table())
        seq = torch.zeros((N_MISSIONS * 4, 128), dtype=torch.int64)
        for age in range(4):
            for f in range(3 - age, 4):
                seq[age::4, f * 32:(f + 1) * 32] = tt
        self.sequences = seq.to(self.device)
        self._packed = None
        self.flat = None                    # see flatten()

    def flatten(self):
        """Re-home every parameter as a view of ONE flat float32 buffer (order = SHAPES = `parameters()`), the layout the
        hand-written optimizer step (csrc/mgrl_update.cu, MGRL_PPO_PARAMS) reads and writes.  The tensors in `params`
        keep their identity, names and shapes."""
        if self.flat is None:
            torch = self.torch
            flat = torch.empty(N_PARAMS, dtype=torch.float32, device=self.device)
            o = 0
            for v in self.params.values():
                n = v.numel()
                flat[o:o + n].copy_(v.detach().reshape(-1))
                v.data = flat[o:o + n].view(v.shape)
                o += n
            self.flat = flat
        return self.flat

    # ---------------------------------------------------------------- init (policies.py:246-257 + torch defaults)
    def _init(self, name, shape, g):
        torch = self.torch
        if name.endswith("bias") and "GRU" not in name:
            return torch.zeros(shape)
        if "Conv2d" in name:
            w = torch.empty(shape)
            torch.nn.init.orthogonal_(w, gain=math.sqrt(2), generator=g)
            return w
        if "Embedding" in name:
            return torch.randn(shape, generator=g)
        if "GRU" in name:
            k = 1.0 / math.sqrt(128)
            return (torch.rand(shape, generator=g) * 2 - 1) * k
        w = torch.randn(shape, generator=g)                       # every Linear: unit-norm rows, gain ignored
        return w / torch.sqrt(w.pow(2).sum(1, keepdim=True))

    def parameters(self):
        return list(self.params.values())

    def _P(self):
        return self._leaves if self._leaves is not None else self.params

    @contextlib.contextmanager
    def fresh_leaves(self):
        """Evaluate through fresh autograd leaves that alias the parameters (same storage, same order as `parameters()`).
        A parameter caches its AccumulateGrad node together with the stream it was created on; if an earlier evaluation
        made it on the legacy default stream and anything keeps that graph alive, a backward pass inside a stream
        capture would have to make the legacy stream wait on the capturing one, which CUDA refuses.  Fresh leaves have
        no such history; their gradients are taken with autograd.grad."""
        self._leaves = {k: v.detach().requires_grad_(True) for k, v in self.params.items()}
        try:
            yield self._leaves
        finally:
            self._leaves = None

    def state_dict(self):
        return {k: v.detach().clone() for k, v in self.params.items()}

    def load_state_dict(self, sd):
        torch = self.torch
        with torch.no_grad():
            for k, v in self.params.items():
                v.copy_(torch.as_tensor(sd[k]).to(v.device, v.dtype).reshape(v.shape))
        self._packed = None

    # ---------------------------------------------------------------- SB3 checkpoint interop (ppo.py:128-132, 145-150)
    def sb3_state_dict(self):
        """The tensors under every key SB3's `policy.state_dict()` has for CustomPPOPolicy with a shared features
        extractor: SB3 also exposes the extractor as `pi_features_extractor.*` and `vf_features_extractor.*` (aliases)."""
        sd = self.state_dict()
        for k in list(sd):
            if k.startswith("features_extractor."):
                sd["pi_" + k] = sd[k]
                sd["vf_" + k] = sd[k]
        return sd

    def save(self, path: str):
        """torch.save of the SB3-named state dict: the `policy.pth` member of an SB3 zip, loadable with
        `model.policy.load_state_dict(torch.load(path))` in the reference."""
        self.torch.save({k: v.cpu() for k, v in self.sb3_state_dict().items()}, path)

    def load(self, path: str):
        """accepts this class's own files and an SB3 `policy.pth` (aliases and optimizer entries are ignored)"""
        sd = self.torch.load(path, map_location="cpu")
        self.load_state_dict({k: sd[k] for k in self.params})

    def save_sb3_zip(self, path: str, data: dict | None = None):
        """An SB3 model archive (`BaseAlgorithm.save`, used at ppo.py:145-150 through EvalCallback / `model.save`): a zip with
        `policy.pth` (the state dict under SB3's names, extractor aliases included), `pytorch_variables.pth`, `data` (JSON:
        plain hyper-parameters only - SB3 pickles its class objects there, which cannot be produced without SB3),
        `_stable_baselines3_version` and `system_info.txt`.  The reference picks the weights up with
        `model.set_parameters(path, exact_match=False)` or `model.policy.load_state_dict(th.load(<policy.pth>))`."""
        import io
        import json
        import zipfile
        torch = self.torch

        def blob(obj):
            f = io.BytesIO()
            torch.save(obj, f)
            return f.getvalue()

        with zipfile.ZipFile(path, "w", zipfile.ZIP_DEFLATED) as z:
            z.writestr("data", json.dumps({"policy_class_name": "CustomPPOPolicy", "n_parameters": N_PARAMS,
                                           "written_by": "minigrid-rl_b200", **(data or {})}, indent=1))
            z.writestr("policy.pth", blob({k: v.cpu() for k, v in self.sb3_state_dict().items()}))
            z.writestr("pytorch_variables.pth", blob({}))
            z.writestr("_stable_baselines3_version", "2.0.0")
            z.writestr("system_info.txt", "minigrid-rl_b200 (B200 device path); weights only\n")

    def load_sb3_zip(self, path: str):
        """Weights of an SB3 model archive (`PPO.load` / `model.save`, ppo.py:128-132,145-150): reads `policy.pth` from the
        zip; the optimizer state and SB3's pickled `data` are not needed for the forward / update paths here."""
        import io
        import zipfile
        with zipfile.ZipFile(path) as z:
            sd = self.torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu")
        missing = [k for k in self.params if k not in sd]
        if missing:
            raise KeyError(f"{path}: policy.pth lacks {missing[:3]}... (not a CustomPPOPolicy checkpoint?)")
        self.load_state_dict({k: sd[k] for k in self.params})

    def load_oracle(self, oracle_policy):
        """copy the weights of an oracle.policy_oracle.OraclePolicy (tests)"""
        o = oracle_policy
        m = {
            _PREFIX + "direction.direction_Linear_0": o.direction[0], _PREFIX + "image.image_Conv2d_0": o.image[0],
            _PREFIX + "image.image_Conv2d_3": o.image[3], _PREFIX + "image.image_Conv2d_5": o.image[5],
            "mlp_extractor.policy_net.0": o.pi[0], "mlp_extractor.policy_net.2": o.pi[2],
            "mlp_extractor.value_net.0": o.vf[0], "mlp_extractor.value_net.2": o.vf[2],
            "action_net": o.action_net, "value_net": o.value_net,
        }
        sd = {}
        for k, mod in m.items():
            sd[k + ".weight"], sd[k + ".bias"] = mod.weight.detach(), mod.bias.detach()
        sd[_PREFIX + "mission.mission_Embedding_0.weight"] = o.embedding.weight.detach()
        for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            sd[_PREFIX + "mission.mission_GRU_1." + n] = getattr(o.gru, n).detach()
        self.load_state_dict(sd)

    # ---------------------------------------------------------------- mission look-up table
    def mission_lut(self):
        """[74*4, 128]: GRU(Embedding(tokens)) final hidden state of every distinct stacked mission (differentiable)."""
        torch = self.torch
        P = self._P()
        x = torch.nn.functional.embedding(self.sequences, P[_PREFIX + "mission.mission_Embedding_0.weight"])
        if x.is_cuda and self._leaves is None and self.flat is None:   # (flatten_parameters would re-home the weights)
            self._gru.flatten_parameters()
        # cuDNN would run the GRU in TF32 by default; the table feeds the fp32 rollout kernel (parity bar 1e-5)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=not self.fp32_strict):
            if self._leaves is None:
                _, h = self._gru(x)
            else:       # the module run on the aliases of its (flattened) weights
                names = ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0")
                _, h = torch.func.functional_call(self._gru, {n: P[_PREFIX + "mission.mission_GRU_1." + n] for n in names}, (x,))
        return h[-1]

    def mission_lut_f64(self):
        """The table for the ROLLOUT kernel, evaluated in float64 and rounded once to float32.  The 128-step recurrence in
        float32 (cuDNN) is off by up to 6e-6 from the exact table - six times the rest of the forward kernel's error and most
        of the 1e-5 parity budget - while the float64 recurrence costs nothing next to a rollout (296 sequences).  No
        gradient flows through this copy; the update differentiates `mission_lut`."""
        torch = self.torch
        if getattr(self, "_gru64", None) is None:
            self._gru64 = torch.nn.GRU(32, 128, 1, True, True).to(self.device).double()
        with torch.no_grad():
            for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
                getattr(self._gru64, n).copy_(self.params[_PREFIX + "mission.mission_GRU_1." + n])
            x = torch.nn.functional.embedding(self.sequences, self.params[_PREFIX + "mission.mission_Embedding_0.weight"].double())
            _, h = self._gru64(x)
        return h[-1].float()

    def _mission_lut_side_stream(self):
        """The GRU over the 296 sequences is 128 dependent little steps (about 1.5 ms forward + backward, whatever the
        minibatch size).  Run it on a side stream: autograd replays a node's backward on the stream of its forward, so
        both passes overlap with the image branch instead of queueing behind it."""
        torch = self.torch
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(device=self.device)
        main = torch.cuda.current_stream(self.device)
        self._side.wait_stream(main)              # the weights were last written (optimizer step) on the main stream
        with torch.cuda.stream(self._side):
            lut = self.mission_lut()
        main.wait_stream(self._side)
        if not self.capturing:                    # (inside a CUDA graph the allocation belongs to the graph's pool)
            lut.record_stream(main)               # allocated on the side stream, consumed on the main one
        return lut

    # ---------------------------------------------------------------- differentiable evaluation (PPO update)
    def evaluate(self, image_u8, dir_onehot, mission_row, lut=None):
        """image_u8 [B,12,7,7] (channels = frame*3 + c, H = view x, W = view y), dir_onehot [B,16], mission_row [B]
        (= mission*4 + age) -> logits [B,7], values [B]."""
        F = self.torch.nn.functional
        P = self._P()
        lut = self.mission_lut() if lut is None else lut
        with self.torch.backends.cudnn.flags(enabled=True, allow_tf32=not self.fp32_strict):
            return self._evaluate(image_u8, dir_onehot, mission_row, lut, F, P)

    def _evaluate(self, image_u8, dir_onehot, mission_row, lut, F, P):
        torch = self.torch
        B = image_u8.shape[0]
        c = _PREFIX + "image.image_Conv2d_"
        # the three 2x2 convolutions as patch GEMMs (7x7 images: the library convolution's backward is ~8x slower here)
        x = image_u8.float() / 255.0
        p = x.unfold(2, 2, 1).unfold(3, 2, 1).permute(0, 2, 3, 1, 4, 5).reshape(B * 36, 48)          # (ci, kh, kw)
        h = torch.relu(F.linear(p, P[c + "0.weight"].reshape(16, 48), P[c + "0.bias"]))
        h = h.view(B, 3, 2, 3, 2, 16).amax(dim=(2, 4))                                                # MaxPool2d(2): 6x6 -> 3x3
        return self._head(h, dir_onehot, mission_row, lut, F, P)

    def _head(self, h, dir_onehot, mission_row, lut, F, P):
        """everything after the first pooled convolution: h [B,3,3,16] (qh, qw, c)"""
        torch = self.torch
        B = h.shape[0]

        def lin(x, w, b):
            return _linear(torch, x, w, b)

        c = _PREFIX + "image.image_Conv2d_"
        p = _patches2x2(torch, h)                                                                     # (kh, kw, ci)
        h = torch.relu(lin(p, P[c + "3.weight"].permute(0, 2, 3, 1).reshape(32, 64), P[c + "3.bias"])).view(B, 128)  # (oh, ow, c2)
        x = torch.relu(lin(h, P[c + "5.weight"].permute(0, 2, 3, 1).reshape(64, 128), P[c + "5.bias"]))
        d = lin(dir_onehot.float(), P[_PREFIX + "direction.direction_Linear_0.weight"],
                     P[_PREFIX + "direction.direction_Linear_0.bias"])
        f = torch.cat([d, x, _lut_rows(torch, lut, mission_row.long())], dim=1)
        t = torch.tanh
        hp = t(lin(t(lin(f, P["mlp_extractor.policy_net.0.weight"], P["mlp_extractor.policy_net.0.bias"])),
                        P["mlp_extractor.policy_net.2.weight"], P["mlp_extractor.policy_net.2.bias"]))
        hv = t(lin(t(lin(f, P["mlp_extractor.value_net.0.weight"], P["mlp_extractor.value_net.0.bias"])),
                        P["mlp_extractor.value_net.2.weight"], P["mlp_extractor.value_net.2.bias"]))
        return (lin(hp, P["action_net.weight"], P["action_net.bias"]),
                lin(hv, P["value_net.weight"], P["value_net.bias"]).squeeze(-1))

    def evaluate_samples(self, frames, dirs, mission, age, t, i, lut=None):
        """`evaluate` for rollout samples (t, i) without materialising their stacked images: the first stage of the image
        branch (Conv2d(12,16,2) + ReLU + MaxPool2d(2), forward and weight gradient) runs in the hand-written kernels
        `mgrl_conv1_pool_forward/backward` straight off the frame buffer; the rest of the network is `_evaluate`'s."""
        torch = self.torch
        F = torch.nn.functional
        P = self._P()
        c = _PREFIX + "image.image_Conv2d_"
        if lut is None:
            lut = self._mission_lut_side_stream() if frames.is_cuda else self.mission_lut()
        a = age[t, i]
        pooled = _conv1_pool(torch)(P[c + "0.weight"], P[c + "0.bias"], frames, t.to(torch.int32), i.to(torch.int32), a)
        k = torch.arange(4, device=t.device)
        valid = ((3 - k)[None, :] <= a.long()[:, None]).to(torch.uint8)
        d = dirs[(t[:, None] + k[None, :]), i[:, None]].long()
        onehot = (F.one_hot(d, 4).to(torch.uint8) * valid[:, :, None]).view(-1, 16)
        mrow = mission[t + 3, i].long() * 4 + a.long()
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=not self.fp32_strict):
            return self._head(pooled.view(-1, 3, 3, 16), onehot, mrow, lut, F, P)

    # ---------------------------------------------------------------- packed weights for the CUDA forward kernel
    def pack(self):
        torch = self.torch
        P = {k: v.detach() for k, v in self.params.items()}
        # fp32 weights, then (CUDA) the fragment section that mgrl_policy_pack_fragments derives from them
        cuda = self.device.type == "cuda"
        out = torch.zeros(N_WEIGHTS + (N_FRAGMENTS if cuda else 0), dtype=torch.float32, device=self.device)

        def put(name, t):
            o, n = WEIGHT_LAYOUT[name]
            t = t.reshape(-1)
            out[o:o + t.numel()] = t

        # conv weights [co][ci][ky][kx] -> [(ci, ky, kx)][co] ; conv2/3 inputs are ordered (ky, kx, ci)
        put("W1", P[_PREFIX + "image.image_Conv2d_0.weight"].permute(1, 2, 3, 0))
        put("B1", P[_PREFIX + "image.image_Conv2d_0.bias"])
        put("W2", P[_PREFIX + "image.image_Conv2d_3.weight"].permute(2, 3, 1, 0))
        put("B2", P[_PREFIX + "image.image_Conv2d_3.bias"])
        put("W3", P[_PREFIX + "image.image_Conv2d_5.weight"].permute(2, 3, 1, 0))
        put("B3", P[_PREFIX + "image.image_Conv2d_5.bias"])
        put("WD", P[_PREFIX + "direction.direction_Linear_0.weight"].t())
        put("BD", P[_PREFIX + "direction.direction_Linear_0.bias"])
        for tag, net in (("PI", "policy_net"), ("VF", "value_net")):
            put(tag + "1", P[f"mlp_extractor.{net}.0.weight"].t()); put(tag + "1B", P[f"mlp_extractor.{net}.0.bias"])
            put(tag + "2", P[f"mlp_extractor.{net}.2.weight"].t()); put(tag + "2B", P[f"mlp_extractor.{net}.2.bias"])
        wa = torch.zeros((64, 8), device=self.device); wa[:, :7] = P["action_net.weight"].t()
        ba = torch.zeros(8, device=self.device); ba[:7] = P["action_net.bias"]
        put("WA", wa); put("BA", ba)
        put("WV", P["value_net.weight"]); put("BV", P["value_net.bias"])
        put("LUT", self.mission_lut_f64())
        self._packed = out.contiguous()
        if cuda:
            s = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
            nat.check(nat.lib().mgrl_policy_pack_fragments(C.c_void_p(self._packed.data_ptr()), s), "mgrl_policy_pack_fragments")
        return self._packed

    def packed(self):
        return self._packed if self._packed is not None else self.pack()

    def invalidate(self):
        self._packed = None

    def kernel_name(self) -> str:
        return ("policy_forward_tc_kernel (mma.sync m16n8k8 TF32, three-term split, fp32 accumulate)" if self.tensor_cores
                else "policy_forward_kernel (fp32 CUDA cores)")

    def forward_rollout(self, frames, dirs, mission, time_index, prev_age, prev_done, age_out, value, action=None,
                        logp=None, logits=None, start_out=None, seed=0, env_id_base=0, step=0, deterministic=False):
        """launch mgrl_policy_forward on the current stream (frames [B,N,148] u8, dirs [B,N] u8, mission [N] u8)"""
        torch = self.torch
        n = int(mission.shape[0])
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        s = C.c_void_p(torch.cuda.current_stream(frames.device).cuda_stream)
        nat.check(nat.lib().mgrl_policy_forward(
            p(self.packed()), p(frames), p(dirs), p(mission), p(prev_age), p(prev_done), p(age_out), p(start_out),
            p(action), p(logp), p(value), p(logits), n, int(time_index), int(seed), int(env_id_base), int(step),
            (FLAG_DETERMINISTIC if deterministic else 0) | (FLAG_TENSOR if self.tensor_cores else 0), s),
            "mgrl_policy_forward")
