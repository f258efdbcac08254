"""Device-resident PPO rollout engine and update.

Replaces, for the hot path, what /root/reference/src/ppo.py:94-136,159 hands to Stable-Baselines3:
`OnPolicyAlgorithm.collect_rollouts` (policy forward -> env.step -> DictRolloutBuffer.add, truncation
bootstrap), `compute_returns_and_advantage` and `PPO.train` (clipped surrogate, clipped value loss, entropy
bonus, per-minibatch advantage normalisation, grad-norm clip, Adam with `linear_schedule`, ppo.py:35-40).

Layout: nothing leaves the GPU during a rollout.  Observations are kept un-stacked as [T+4, N, 148] u8 frame
records (+ direction, mission id, frames-of-history `age`); the 4-frame stack the policy sees is gathered on
read.  The rollout issues two launches per step: `mgrl_policy_forward` (hand-written fused
forward + sampling) and `mgrl_step`.  The update evaluates the same network through torch autograd
(library kernels) on minibatches gathered from the frame buffer.

Data parallel: one process per GPU, environments sharded by global id; the only exchange is one all-reduce of
the flat gradient (+ the advantage moments of the minibatch, so the normalisation equals the single-process
one on the concatenated minibatch) per optimizer step.
"""
from __future__ import annotations

import os
from dataclasses import dataclass

import numpy as np

from .policy import Policy


@dataclass
class PPOConfig:                      # hydra_configs/algorithm/ppo.yaml:9-40
    n_steps: int = 128                # BASELINE.json configs[1] (reference: horizon 1024)
    batch_size: int = 65536           # reference: 256 at 16 envs; scaled with the env count
    n_epochs: int = 4
    gamma: float = 0.8108071290665859
    gae_lambda: float = 0.9452281119742252
    clip_range: float = 0.1
    clip_range_vf: float | None = 0.08341734780140342
    normalize_advantage: bool = True
    ent_coef: float = 0.045732238989694494
    vf_coef: float = 0.8177283657817492
    max_grad_norm: float = 0.5215982006116593
    initial_learning_rate: float = 3e-4
    final_learning_rate: float = 3e-6
    optim_eps: float = 1e-8           # single.yaml:31
    update_tf32: bool = False         # True: the update's library GEMMs may use TF32 like the reference (ppo.py:29-32)
    native_conv1: bool = True         # first extractor stage of the update in the hand-written kernels (CUDA only)
    cuda_graph: bool = True           # replay the optimizer step (forward + backward + clip + Adam) from a CUDA graph
    native_update: bool = True        # the whole optimizer step in hand-written kernels (csrc/mgrl_update.cu; CUDA only)
    update_tcgen05: bool = True       # with update_tf32: the 208 x 128 MLP GEMMs as tcgen05.mma + TMEM (csrc/mgrl_linear_tc5.cu)
    total_timesteps: float = 2e7


def linear_schedule(initial_value: float, final_value: float):
    """ppo.py:35-40"""
    return lambda progress_remaining: max(progress_remaining * initial_value, final_value)


# ------------------------------------------------------------------------------------------- update (any device)
def gather_minibatch(buf, t, i):
    """Stacked observations of samples (t, i) from the un-stacked frame buffer: image [B,12,7,7] u8 (VecTransposeImage +
    VecFrameStack order), direction one-hot [B,16], mission row [B] = mission*4 + age."""
    import torch
    b = t + 3
    age = buf["age"][t, i].long()
    k = torch.arange(4, device=t.device)
    valid = (3 - k)[None, :] <= age[:, None]                                        # [B,4]
    fr = buf["frames"][(b[:, None] - 3 + k[None, :]), i[:, None]]                     # [B,4,148]
    fr = fr[:, :, :147] * valid[:, :, None].to(torch.uint8)
    image = fr.view(-1, 4, 7, 7, 3).permute(0, 1, 4, 2, 3).reshape(-1, 12, 7, 7)
    d = buf["dirs"][(b[:, None] - 3 + k[None, :]), i[:, None]].long()                 # [B,4]
    onehot = torch.nn.functional.one_hot(d, 4).to(torch.uint8) * valid[:, :, None].to(torch.uint8)
    return image, onehot.view(-1, 16), buf["mission"][b, i].long() * 4 + age


def ppo_minibatch_loss(policy: Policy, cfg: PPOConfig, image, onehot, mrow, actions, old_values, old_logp, adv, returns,
                       adv_stats=None, samples=None):
    """[UPSTREAM] PPO.train, one minibatch.  adv_stats = (mean, std) when they were computed over all ranks.
    samples = (buf, t, i): evaluate straight off the rollout buffer (hand-written first extractor stage) instead of the
    gathered (image, onehot, mrow)."""
    import torch
    F = torch.nn.functional
    if samples is not None:
        buf, t, i = samples
        logits, values = policy.evaluate_samples(buf["frames"], buf["dirs"], buf["mission"], buf["age"], t, i)
    else:
        logits, values = policy.evaluate(image, onehot, mrow)
    logp_all = F.log_softmax(logits, dim=1)
    logp = logp_all.gather(1, actions.long().view(-1, 1)).squeeze(1)
    entropy = -(logp_all.exp() * logp_all).sum(1)
    if cfg.normalize_advantage and adv.numel() > 1:
        mean, std = adv_stats if adv_stats is not None else (adv.mean(), adv.std())
        adv = (adv - mean) / (std + 1e-8)
    ratio = torch.exp(logp - old_logp)
    pl = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - cfg.clip_range, 1 + cfg.clip_range)).mean()
    vp = values if cfg.clip_range_vf is None else old_values + torch.clamp(values - old_values, -cfg.clip_range_vf,
                                                                          cfg.clip_range_vf)
    vl = F.mse_loss(returns, vp)
    el = -entropy.mean()
    return pl + cfg.ent_coef * el + cfg.vf_coef * vl, (pl.detach(), vl.detach(), el.detach())


class Updater:
    """Adam + grad clip + (optional) data-parallel all-reduce over `dist` (torch.distributed, any backend)."""

    def __init__(self, policy: Policy, cfg: PPOConfig, dist=None):
        import torch
        self.torch, self.policy, self.cfg, self.dist = torch, policy, cfg, dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.params = policy.parameters()
        # One optimizer step is ~150 small launches; issued from Python they take longer (5.1 ms) than the GPU needs to run
        # them, so on one GPU the step is captured once into a CUDA graph and replayed (capturable Adam, lr in a tensor).
        # One rank: the whole step is one graph.  NCCL ranks: two graphs around the eager gradient all-reduce.  With the two
        # all-reduces captured as well (MGRL_GRAPH_NCCL=1) the step ran as fast as on one GPU (2 GPUs: update 563 ms,
        # 28.0 M frames/s), but the process then hung in destroy_process_group while the graph that holds the captured
        # collectives was alive; that mode stays opt-in until the teardown is sorted out.
        nccl = dist is not None and self.world > 1 and str(dist.get_backend()) == "nccl"
        self.graph_nccl = nccl and os.environ.get("MGRL_GRAPH_NCCL", "0") == "1"
        self.graphed = bool(cfg.cuda_graph and (self.world == 1 or nccl) and self.params[0].is_cuda and cfg.native_conv1)
        if self.graphed:
            self.opt = torch.optim.Adam(self.params, lr=torch.tensor(float(cfg.initial_learning_rate), device=self.params[0].device),
                                        eps=cfg.optim_eps, capturable=True, fused=True)
        else:
            self.opt = torch.optim.Adam(self.params, lr=cfg.initial_learning_rate, eps=cfg.optim_eps)
        self._static = None
        self.schedule = linear_schedule(cfg.initial_learning_rate, cfg.final_learning_rate)
        self.n_all_reduces = 0
        if dist is not None and self.world > 1:       # identical weights everywhere (rank 0's)
            for p in self.params:
                dist.broadcast(p.data, src=0)
            policy.invalidate()

    def describe(self) -> str:
        return ("first extractor stage hand-written (mgrl_conv1_pool_*), patch gather, table and bias gradients hand-written; "
                "remaining GEMMs / GRU table / loss through torch autograd on library kernels"
                + ("; optimizer step replayed from a CUDA graph" if self.graphed else ""))

    def all_reduces_per_step(self) -> int:
        return 0 if self.world == 1 else 1 + int(self.cfg.normalize_advantage)

    def release(self):
        """drop the captured optimizer step (it must not outlive a process group it captured collectives of)"""
        self._static = None

    def set_progress(self, progress_remaining: float):
        lr = self.schedule(progress_remaining)
        for g in self.opt.param_groups:
            if self.torch.is_tensor(g["lr"]):
                g["lr"].fill_(lr)               # read by the captured Adam step
            else:
                g["lr"] = lr

    def global_adv_stats(self, adv):
        """mean and unbiased std of the minibatch advantages over ALL ranks: one small all-reduce"""
        torch = self.torch
        s = torch.stack([adv.sum(dtype=torch.float64), (adv.double() ** 2).sum(),
                         torch.full((), float(adv.numel()), dtype=torch.float64, device=adv.device)])
        self.dist.all_reduce(s)
        self.n_all_reduces += 1
        n = s[2]
        mean = s[0] / n
        var = (s[1] - n * mean * mean) / (n - 1)
        return mean.float(), var.clamp_min(0).sqrt().float()

    def step(self, loss):
        torch = self.torch
        self.opt.zero_grad(set_to_none=True)
        loss.backward()
        grads = [p.grad if p.grad is not None else torch.zeros_like(p) for p in self.params]
        if self.world > 1:                      # ONE all-reduce of the flat gradient buffer (441 KB)
            flat = torch.cat([g.reshape(-1) for g in grads])
            self.dist.all_reduce(flat)
            self.n_all_reduces += 1
            flat /= self.world
            o = 0
            for p, g in zip(self.params, grads):
                p.grad = flat[o:o + g.numel()].view_as(g)
                o += g.numel()
        torch.nn.utils.clip_grad_norm_(self.params, self.cfg.max_grad_norm)
        self.opt.step()
        self.policy.invalidate()

    def minibatch(self, image, onehot, mrow, actions, old_values, old_logp, adv, returns, samples=None):
        torch = self.torch
        stats = None
        if self.world > 1 and self.cfg.normalize_advantage:
            stats = self.global_adv_stats(adv)
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(self.cfg.update_tf32)
        try:
            loss, parts = ppo_minibatch_loss(self.policy, self.cfg, image, onehot, mrow, actions, old_values, old_logp, adv,
                                             returns, stats, samples)
            self.step(loss)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return loss.detach(), parts


    def gradients(self, buf, t, i):
        """Loss and the gradient of every parameter for the minibatch samples (t, i), through the same evaluation path an
        optimizer step takes (no step is made): what the parity tests compare with the torch-CPU oracle."""
        torch = self.torch
        args = (buf["actions"][t, i], buf["values"][t, i], buf["logp"][t, i], buf["adv"][t, i], buf["ret"][t, i])
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(self.cfg.update_tf32)
        try:
            with self.policy.fresh_leaves() as leaves:
                loss, _ = ppo_minibatch_loss(self.policy, self.cfg, None, None, None, *args, None, (buf, t, i))
                grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        return loss.detach(), {k: (g if g is not None else torch.zeros_like(p))
                               for (k, p), g in zip(self.policy.params.items(), grads)}

    def minibatch_samples(self, buf, t, i, eager: bool = False):
        """One optimizer step on the samples (t, i) of the rollout buffer `buf` (hand-written first stage).  `eager`: do not
        use (or disturb) the captured step - for the trailing partial minibatch of an epoch, whose shape differs.  With
        `graphed`, the step - gathers, forward, loss, backward, gradient clip, Adam - is replayed from CUDA graphs over
        static index tensors: three eager steps first (library workspaces, Adam state), then the capture.  One rank: one
        graph.  NCCL ranks: two graphs (gradients | clip + Adam) around the eager all-reduce of the flat gradient, the
        advantage moments being all-reduced eagerly before the first; MGRL_GRAPH_NCCL=1 captures the collectives too."""
        torch = self.torch
        multi = self.world > 1
        norm = multi and self.cfg.normalize_advantage

        def stats_part(tt, ii):                  # advantage moments over ALL ranks -> static tensor
            if norm:
                m, sd = self.global_adv_stats(buf["adv"][tt, ii])
                st["stats"][0].copy_(m); st["stats"][1].copy_(sd)

        def grads_part(tt, ii):
            args = (buf["actions"][tt, ii], buf["values"][tt, ii], buf["logp"][tt, ii], buf["adv"][tt, ii], buf["ret"][tt, ii])
            stats = (st["stats"][0], st["stats"][1]) if norm else None
            # fresh leaves + autograd.grad, not backward(): see Policy.fresh_leaves
            with self.policy.fresh_leaves() as leaves:
                loss, parts = ppo_minibatch_loss(self.policy, self.cfg, None, None, None, *args, stats, (buf, tt, ii))
                grads = torch.autograd.grad(loss, list(leaves.values()), allow_unused=True)
            grads = [g if g is not None else torch.zeros_like(p) for p, g in zip(self.params, grads)]
            if multi:                            # the flat gradient buffer (441 KB) that is all-reduced
                torch.cat([g.reshape(-1) for g in grads], out=st["flat"])
            else:
                for p, g in zip(self.params, grads):
                    p.grad = g.contiguous()      # (the fused Adam wants the parameter's layout)
            return loss.detach(), parts

        def step_part():
            if multi:
                st["flat"].div_(self.world)
                o = 0
                for p in self.params:
                    p.grad = st["flat"][o:o + p.numel()].view(p.shape)
                    o += p.numel()
            torch.nn.utils.clip_grad_norm_(self.params, self.cfg.max_grad_norm)
            self.opt.step()

        def full(tt, ii):
            stats_part(tt, ii)
            out = grads_part(tt, ii)
            if multi:
                self.dist.all_reduce(st["flat"])     # ONE all-reduce of the gradients per optimizer step
            step_part()
            return out

        if not self.graphed or eager:
            args = (buf["actions"][t, i], buf["values"][t, i], buf["logp"][t, i], buf["adv"][t, i], buf["ret"][t, i])
            return self.minibatch(None, None, None, *args, samples=(buf, t, i))
        st = self._static
        if st is None or st["t"].shape != t.shape or st["frames"] is not buf["frames"]:
            dev = t.device
            st = self._static = {"t": torch.empty_like(t), "i": torch.empty_like(i), "frames": buf["frames"], "graph": None,
                                 "graph_b": None, "warm": 0, "out": None,
                                 "stats": torch.zeros(2, dtype=torch.float32, device=dev),
                                 "flat": torch.zeros(sum(p.numel() for p in self.params), dtype=torch.float32, device=dev)}
        st["t"].copy_(t); st["i"].copy_(i)
        split = multi and not self.graph_nccl
        n_red = self.n_all_reduces + (0 if not multi else 1 + int(self.cfg.normalize_advantage))
        prev = torch.backends.cuda.matmul.allow_tf32
        torch.backends.cuda.matmul.allow_tf32 = bool(self.cfg.update_tf32)
        try:
            if st["graph"] is None and st["warm"] < 3:
                side = torch.cuda.Stream(device=t.device)
                side.wait_stream(torch.cuda.current_stream(t.device))
                with torch.cuda.stream(side):
                    out = full(st["t"], st["i"])
                torch.cuda.current_stream(t.device).wait_stream(side)
                st["warm"] += 1
                self.n_all_reduces = n_red
                self.policy.invalidate()
                return out
            if st["graph"] is None:
                graph = torch.cuda.CUDAGraph()
                self.policy.capturing = True
                try:
                    if split:
                        stats_part(st["t"], st["i"])                 # eager: this step's moments
                        with torch.cuda.graph(graph):
                            st["out"] = grads_part(st["t"], st["i"])
                        st["graph_b"] = torch.cuda.CUDAGraph()
                        with torch.cuda.graph(st["graph_b"], pool=graph.pool()):
                            step_part()
                    else:
                        with torch.cuda.graph(graph):
                            st["out"] = full(st["t"], st["i"])
                finally:
                    self.policy.capturing = False
                st["graph"] = graph
            elif split:
                stats_part(st["t"], st["i"])
            st["graph"].replay()
            if split:
                self.dist.all_reduce(st["flat"])
                st["graph_b"].replay()
        finally:
            torch.backends.cuda.matmul.allow_tf32 = prev
        self.n_all_reduces = n_red
        self.policy.invalidate()
        return st["out"]


class NativeUpdater:
    """The PPO optimizer step in hand-written kernels only (include/mgrl.h `mgrl_ppo_*`, csrc/mgrl_update.cu): forward of
    extractor + MLPs + heads, loss, every gradient, global-norm clip and Adam for one minibatch of rollout samples, issued
    as ~30 launches of this library on the current stream - no cuBLAS / cuDNN / torch kernel in between.  Parameters,
    gradients and Adam moments live in flat buffers (`Policy.flatten`).  Data parallel: per epoch ONE all-reduce of the
    advantage moments of all its minibatches, per optimizer step ONE all-reduce of the flat gradient."""

    def __init__(self, policy: Policy, cfg: PPOConfig, dist=None, max_batch: int | None = None):
        import ctypes as C
        import torch
        from . import _native as nat
        self.torch, self.policy, self.cfg, self.dist, self.nat, self.C = torch, policy, cfg, dist, nat, C
        self.world = dist.get_world_size() if dist is not None else 1
        dev = policy.device
        assert dev.type == "cuda", "the hand-written optimizer step needs CUDA (there is no CPU path)"
        self.P = policy.flatten()
        self.params = policy.parameters()
        if dist is not None and self.world > 1:       # identical weights everywhere (rank 0's)
            dist.broadcast(self.P, src=0)
            policy.invalidate()
        self.G = torch.zeros_like(self.P)
        self.M = torch.zeros_like(self.P)
        self.V = torch.zeros_like(self.P)
        self.seq = policy.sequences.to(torch.uint8).contiguous()
        self.max_batch, self._h = 0, None
        self._ensure(int(max_batch or cfg.batch_size))
        self.hyper = nat.PPOHyper(float(cfg.clip_range), float(-1.0 if cfg.clip_range_vf is None else cfg.clip_range_vf),
                                  float(cfg.ent_coef), float(cfg.vf_coef), int(bool(cfg.normalize_advantage)),
                                  int(not cfg.update_tf32), int(bool(cfg.update_tcgen05 and cfg.update_tf32)))
        self.schedule = linear_schedule(cfg.initial_learning_rate, cfg.final_learning_rate)
        self.lr = float(cfg.initial_learning_rate)
        self.steps = 0                      # Adam step counter
        self.n_all_reduces = 0
        self.graphed = False
        self.stats = torch.zeros(4, dtype=torch.float32, device=dev)
        self.norm = torch.zeros(1, dtype=torch.float32, device=dev)
        self.launches_per_step = 33

    def _ensure(self, batch: int):
        """(re)create the context so that its activation buffers hold `batch` samples (about 6.1 KB per sample)"""
        if batch <= self.max_batch:
            return
        nat, C, torch = self.nat, self.C, self.torch
        if self._h is not None:
            torch.cuda.synchronize(self.policy.device)
            nat.check(nat.lib().mgrl_ppo_destroy(self._h), "mgrl_ppo_destroy")
            self._h = None
        dev = self.policy.device
        h = C.c_void_p()
        index = dev.index if dev.index is not None else torch.cuda.current_device()
        nat.check(nat.lib().mgrl_ppo_create(int(index), int(batch), int(self.seq.shape[0]), C.byref(h)), "mgrl_ppo_create")
        self._h, self.max_batch = h, int(batch)
        p = lambda t: C.c_void_p(t.data_ptr())  # noqa: E731
        nat.check(nat.lib().mgrl_ppo_bind(h, p(self.P), p(self.G), p(self.M), p(self.V), p(self.seq)), "mgrl_ppo_bind")

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h is not None:
            try:
                self.nat.lib().mgrl_ppo_destroy(h)
            except Exception:
                pass

    def describe(self) -> str:
        return ("hand-written end to end (mgrl_ppo_gradients + mgrl_ppo_apply: mma.sync TF32 GEMM kernels with fused epilogues, "
                "register-resident GRU forward/backward, loss, clip + Adam; "
                + ("three-term split, fp32-class" if self.hyper.strict_fp32 else "one TF32 pass like ppo.py:29-32")
                + ("; first convolution (pool-window GEMM, kind::f16) and the 208 x 128 MLP GEMMs (kind::tf32) on tcgen05.mma with TMEM accumulators" if self.hyper.use_tcgen05 else "") + ")")

    def all_reduces_per_step(self) -> int:
        return 0 if self.world == 1 else 1

    def release(self):
        pass

    def set_progress(self, progress_remaining: float):
        self.lr = float(self.schedule(progress_remaining))

    def _stream(self):
        return self.C.c_void_p(self.torch.cuda.current_stream(self.policy.device).cuda_stream)

    def view(self, buf):
        """mgrl_rollout_view of a RolloutEngine buffer dict"""
        names = ("frames", "dirs", "mission", "age", "actions", "values", "logp", "adv", "ret")
        for k in names:
            assert buf[k].is_contiguous(), k
        return self.nat.RolloutView(*[buf[k].data_ptr() for k in names], int(buf["frames"].shape[1]))

    def moments(self, buf, idx32, batch: int):
        """(sum, sum of squares, count) of the advantages of every consecutive minibatch of idx32, summed over all ranks"""
        torch, C = self.torch, self.C
        total = int(idx32.numel())
        n_mb = (total + batch - 1) // batch
        sums = torch.empty((n_mb, 3), dtype=torch.float64, device=idx32.device)
        self.nat.check(self.nat.lib().mgrl_ppo_moments(C.c_void_p(buf["adv"].data_ptr()), C.c_void_p(idx32.data_ptr()), int(batch),
                                                       total, C.c_void_p(sums.data_ptr()), self._stream()), "mgrl_ppo_moments")
        if self.world > 1:
            self.dist.all_reduce(sums)
            self.n_all_reduces += 1
        return sums

    def gradients_native(self, view, idx32, sums_row, logits=None, values=None):
        C = self.C
        self._ensure(int(idx32.numel()))
        p = lambda t: None if t is None else C.c_void_p(t.data_ptr())  # noqa: E731
        self.nat.check(self.nat.lib().mgrl_ppo_gradients(self._h, C.byref(view), p(idx32), int(idx32.numel()), p(sums_row),
                                                         C.byref(self.hyper), p(self.stats), p(logits), p(values), self._stream()),
                       "mgrl_ppo_gradients")

    def apply(self):
        C = self.C
        if self.world > 1:
            self.dist.all_reduce(self.G)                 # ONE all-reduce of the gradients per optimizer step
            self.n_all_reduces += 1
        self.steps += 1
        self.nat.check(self.nat.lib().mgrl_ppo_apply(self._h, self.lr, float(self.cfg.max_grad_norm), 1.0 / self.world, 0.9, 0.999,
                                                     float(self.cfg.optim_eps), self.steps, C.c_void_p(self.norm.data_ptr()),
                                                     self._stream()), "mgrl_ppo_apply")
        self.policy.invalidate()

    def minibatch_native(self, view, idx32, sums_row):
        self.gradients_native(view, idx32, sums_row)
        self.apply()

    def loss_terms(self, batch: int):
        """(loss, policy loss, value loss, -entropy) of the last minibatch, from the kernel's sums"""
        s = self.stats / float(batch)
        return s[0] + self.cfg.ent_coef * s[2] + self.cfg.vf_coef * s[1], (s[0], s[1], s[2])

    def gradients(self, buf, t, i):
        """Loss and the gradient of every parameter for the samples (t, i): what the parity tests compare with the oracle."""
        torch = self.torch
        N = int(buf["frames"].shape[1])
        idx32 = (t.long() * N + i.long()).to(torch.int32).contiguous()
        sums = self.moments(buf, idx32, int(idx32.numel())) if self.cfg.normalize_advantage else None
        self.gradients_native(self.view(buf), idx32, None if sums is None else sums[0])
        loss, _ = self.loss_terms(int(idx32.numel()))
        out, o = {}, 0
        for k, v in self.policy.params.items():
            out[k] = self.G[o:o + v.numel()].view(v.shape).clone()
            o += v.numel()
        return loss.detach().clone(), out

    def debug_buffer(self, name: str, shape):
        """a copy of an internal activation buffer of the last gradients call (tests)"""
        torch, C = self.torch, self.C
        out = torch.empty(int(np.prod(shape)), dtype=torch.float32, device=self.policy.device)
        self.nat.check(self.nat.lib().mgrl_ppo_debug_copy(self._h, name.encode(), C.c_void_p(out.data_ptr()), int(out.numel()),
                                                          self._stream()), "mgrl_ppo_debug_copy")
        return out.view(*shape)

    def mission_table(self):
        torch, C = self.torch, self.C
        out = torch.empty((int(self.seq.shape[0]), 128), dtype=torch.float32, device=self.policy.device)
        self.nat.check(self.nat.lib().mgrl_ppo_mission_table(self._h, C.c_void_p(out.data_ptr()), self._stream()),
                       "mgrl_ppo_mission_table")
        return out


# ------------------------------------------------------------------------------------------- rollout (CUDA)
class RolloutEngine:
    def __init__(self, env, policy: Policy, cfg: PPOConfig, dist=None, seed: int = 0, keep_terminal_frames: bool = True):
        import torch
        from . import vec_env
        assert env.layout == "hwc148", "the rollout engine reads 148-byte HWC frame records"
        self.torch, self.env, self.policy, self.cfg, self.dist = torch, env, policy, cfg, dist
        self.gae = vec_env.gae
        self.N, self.T = env.num_envs, cfg.n_steps
        self.rank = dist.get_rank() if dist is not None else 0
        self.env_id_base = int(env._h.ncfg.env_id_base)
        self.seed = int(seed)
        N, T, dev = self.N, self.T, env.device
        u8 = dict(dtype=torch.uint8, device=dev)
        f32 = dict(dtype=torch.float32, device=dev)
        self.buf = {
            "frames": torch.zeros((T + 4, N, 148), **u8), "dirs": torch.zeros((T + 4, N), **u8),
            "mission": torch.zeros((T + 4, N), **u8), "age": torch.zeros((T + 1, N), **u8),
            "start": torch.zeros((T + 1, N), **u8), "actions": torch.zeros((T, N), **u8),
            "logp": torch.zeros((T, N), **f32), "values": torch.zeros((T + 1, N), **f32),
            "rewards": torch.zeros((T, N), **f32), "term": torch.zeros((T, N), **u8), "trunc": torch.zeros((T, N), **u8),
            "ep_len": torch.zeros((T, N), **u8), "adv": torch.zeros((T, N), **f32), "ret": torch.zeros((T, N), **f32),
        }
        self.term_frames = torch.zeros((T, N, 148), **u8) if keep_terminal_frames else None
        self.term_dirs = torch.zeros((T, N), **u8) if keep_terminal_frames else None
        self.prev_done = torch.ones(N, **u8)
        self.prev_age = torch.zeros(N, **u8)
        self.global_step = 0
        self.world = dist.get_world_size() if dist is not None else 1
        self.num_timesteps = 0            # SB3's counter: env steps of all environments (all ranks) collected so far
        self.records_logits = None        # [T, N, 7] logits of the rollout, kept when records are being collected
        self.native = bool(cfg.native_update and cfg.native_conv1 and dev.type == "cuda")
        self.updater = (NativeUpdater(policy, cfg, dist, max_batch=min(cfg.batch_size, T * N)) if self.native
                        else Updater(policy, cfg, dist))
        self.launches = 0
        self.reset()

    def reset(self):
        b = self.buf
        self.env.reset()
        b["frames"][3].copy_(self.env.image); b["dirs"][3].copy_(self.env.dir); b["mission"][3].copy_(self.env.mission)
        self.prev_done.fill_(1); self.prev_age.zero_()
        self.launches += 1

    def collect(self, deterministic: bool = False):
        """One rollout of T steps: 2 launches per step, everything stays on the device."""
        b, T, env, pol = self.buf, self.T, self.env, self.policy
        pol.packed()
        for t in range(T):
            k = t + 3
            pol.forward_rollout(b["frames"], b["dirs"], b["mission"][k], k, self.prev_age, self.prev_done, b["age"][t],
                                b["values"][t], b["actions"][t], b["logp"][t], start_out=b["start"][t], seed=self.seed,
                                env_id_base=self.env_id_base, step=self.global_step, deterministic=deterministic,
                                logits=None if self.records_logits is None else self.records_logits[t])
            env.step(b["actions"][t], b["frames"][k + 1], b["dirs"][k + 1], b["mission"][k + 1], b["rewards"][t],
                     b["term"][t], b["trunc"][t], b["ep_len"][t],
                     term_image=None if self.term_frames is None else self.term_frames[t],
                     term_dir=None if self.term_dirs is None else self.term_dirs[t])
            self.prev_age, self.prev_done = b["age"][t], b["ep_len"][t]      # ep_len != 0 exactly on done steps
            self.global_step += 1
        pol.forward_rollout(b["frames"], b["dirs"], b["mission"][T + 3], T + 3, self.prev_age, self.prev_done, b["age"][T],
                            b["values"][T], start_out=b["start"][T], seed=self.seed, env_id_base=self.env_id_base,
                            step=self.global_step)
        self.launches += 2 * T + 1
        self.num_timesteps += T * self.N * self.world

    def bootstrap_truncated(self):
        """[UPSTREAM] collect_rollouts: reward += gamma * V(terminal_observation) where truncated and not terminated."""
        torch = self.torch
        b = self.buf
        if self.term_frames is None:
            return 0
        idx = torch.nonzero((b["trunc"] != 0) & (b["term"] == 0))
        if idx.numel() == 0:
            return 0
        self._bootstrap(idx[:, 0], idx[:, 1], self.cfg.gamma)
        return int(idx.shape[0])

    def warm_bootstrap(self):
        """Runs the truncation bootstrap once on two rollout slots with weight 0 (rewards unchanged): truncations are rare, and the
        first one otherwise pays for the lazily loaded indexing kernels (tens of milliseconds) inside a timed iteration."""
        if self.term_frames is not None:
            dev = self.buf["rewards"].device
            self._bootstrap(self.torch.tensor([0, 1], device=dev), self.torch.tensor([0, 0], device=dev), 0.0)

    def _bootstrap(self, t, i, gamma):
        torch = self.torch
        b = self.buf
        # The terminal observation extends the finished episode by one frame: stack = buffer slots of times t-2..t + the
        # terminal frame, the old mission, age = min(age + 1, 3).  Valued by the rollout's own forward kernel (the same
        # numbers the rollout would have produced, and no library kernels whose first call on a new batch size costs
        # tens of milliseconds): the M samples are gathered into a small [4, M, 148] frame buffer.
        k = torch.arange(3, device=t.device)
        hist = b["frames"][(t[:, None] + 1 + k[None, :]), i[:, None]]                       # [M,3,148]
        frames4 = torch.cat([hist, self.term_frames[t, i][:, None]], dim=1).permute(1, 0, 2).contiguous()
        dirs4 = torch.cat([b["dirs"][(t[:, None] + 1 + k[None, :]), i[:, None]], self.term_dirs[t, i][:, None]],
                          dim=1).t().contiguous()
        m = int(t.numel())
        v = torch.empty(m, dtype=torch.float32, device=t.device)
        age_out = torch.empty(m, dtype=torch.uint8, device=t.device)
        self.policy.forward_rollout(frames4, dirs4, b["mission"][t + 3, i].contiguous(), 3, b["age"][t, i].contiguous(),
                                    torch.zeros(m, dtype=torch.uint8, device=t.device), age_out, v)
        self.launches += 1
        b["rewards"][t, i] += gamma * v if gamma != 0.0 else torch.zeros_like(v)      # (warm-up: the same kernels, rewards unchanged)

    def compute_advantages(self):
        b, T = self.buf, self.T
        self.gae(b["rewards"], b["values"][:T].contiguous(), b["start"][:T].contiguous(), b["values"][T].contiguous(),
                 b["start"][T].contiguous(), self.cfg.gamma, self.cfg.gae_lambda, b["adv"], b["ret"])
        self.launches += 1

    def update(self, generator=None):
        """n_epochs passes over the rollout in random minibatches (per-rank permutation of the local shard)."""
        torch = self.torch
        b, T, N, cfg = self.buf, self.T, self.N, self.cfg
        total = T * N
        bs = min(cfg.batch_size, total)
        n_mb = 0
        if self.native:
            up = self.updater
            view = up.view(b)
            for _ in range(cfg.n_epochs):
                idx32 = torch.randperm(total, device=b["adv"].device, generator=generator).to(torch.int32)
                sums = up.moments(b, idx32, bs) if cfg.normalize_advantage else None
                for k, s in enumerate(range(0, total, bs)):      # the trailing partial minibatch is trained on too
                    up.minibatch_native(view, idx32[s:s + bs], None if sums is None else sums[k])
                    n_mb += 1
            self.launches += n_mb * up.launches_per_step
            return n_mb
        for _ in range(cfg.n_epochs):
            perm = torch.randperm(total, device=b["adv"].device, generator=generator)
            for s in range(0, total, bs):        # like RolloutBuffer.get: the trailing partial minibatch is trained on too
                idx = perm[s:s + bs]
                t, i = idx // N, idx % N
                if cfg.native_conv1 and b["frames"].is_cuda:
                    self.updater.minibatch_samples(b, t, i, eager=int(idx.numel()) != bs)
                else:
                    args = (b["actions"][t, i], b["values"][t, i], b["logp"][t, i], b["adv"][t, i], b["ret"][t, i])
                    image, onehot, mrow = gather_minibatch(b, t, i)
                    self.updater.minibatch(image, onehot, mrow, *args)
                n_mb += 1
        return n_mb

    # ---------------------------------------------------------------- evaluation / record collection (ppo.py:161, 174-292)
    def _snapshot(self):
        b = self.buf
        return {"state": self.env.get_state().clone(), "seed": self.env.seed, "prev_age": self.prev_age.clone(),
                "prev_done": self.prev_done.clone(), "global_step": self.global_step, "num_timesteps": self.num_timesteps,
                "head": {k: b[k][0:4].clone() for k in ("frames", "dirs", "mission")}}

    def _restore(self, snap):
        b = self.buf
        self.env.set_state(snap["state"], snap["seed"])
        self.prev_age, self.prev_done = snap["prev_age"], snap["prev_done"]
        self.global_step, self.num_timesteps = snap["global_step"], snap["num_timesteps"]
        for k, v in snap["head"].items():
            b[k][0:4].copy_(v)

    def _fresh_episodes(self, seed):
        """reset for an evaluation: the evaluation's own Philox key, first observation into the buffer"""
        b = self.buf
        self.env.reset(seed)
        b["frames"][3].copy_(self.env.image); b["dirs"][3].copy_(self.env.dir); b["mission"][3].copy_(self.env.mission)
        self.prev_done = self.torch.ones_like(self.prev_done)
        self.prev_age = self.torch.zeros_like(self.prev_age)

    def evaluate(self, n_eval_episodes_per_env: int = 1, deterministic: bool = True, seed: int | None = None,
                 max_rollouts: int = 64):
        """SB3 `evaluate_policy` (ppo.py:161; `test()` ppo.py:174-292) on the device: every environment is reset, then its
        FIRST `n_eval_episodes_per_env` episodes count - whatever their length, so long (timed-out) episodes are not
        under-represented - and nothing of the training state moves: environments, frame history, step counters and the
        environment seed are restored afterwards.  Returns episodes, mean reward, success rate (reward > 0), mean length;
        with several ranks the sums are all-reduced."""
        torch = self.torch
        dev = self.buf["rewards"].device
        snap = self._snapshot()
        self._fresh_episodes(self.seed + 0x5EED if seed is None else seed)
        tot = torch.zeros(4, dtype=torch.float64, device=dev)
        seen = torch.zeros(self.N, dtype=torch.int64, device=dev)
        for _ in range(max_rollouts):
            self.collect(deterministic=deterministic)
            b = self.buf
            done = b["ep_len"] != 0
            order = torch.cumsum(done.long(), dim=0) + seen[None, :]          # 1-based index of each finished episode
            take = done & (order <= n_eval_episodes_per_env)
            r = b["rewards"][take].double()
            tot += torch.stack([take.sum().double(), r.sum(), (r > 0).sum().double(), b["ep_len"][take].double().sum()])
            seen += done.sum(0)
            self.shift()
            if bool((seen >= n_eval_episodes_per_env).all()):
                break
        self._restore(snap)
        if self.dist is not None and self.dist.get_world_size() > 1:
            self.dist.all_reduce(tot)
        n = max(float(tot[0]), 1.0)
        return {"episodes": int(tot[0]), "mean_reward": float(tot[1]) / n, "success_rate": float(tot[2]) / n,
                "mean_length": float(tot[3]) / n}

    def collect_records(self, min_records: int = 1, deterministic: bool = True, seed: int | None = None,
                        max_rollouts: int = 16):
        """The rollout collection of `test()` with cfg.collect_rollouts (ppo.py:176-183, 214-262, 279-289): the policy acts
        (argmax unless `deterministic=False`), and every step of every episode that ENDED WITH A REWARD contributes one
        record `(image [12,7,7] u8, direction [16] u8, mission [128] i64, policy [7] f32)` - the stacked observation the
        policy saw and the action probabilities it produced.  Episodes are taken after a reset and must lie inside one
        rollout window.  Returns a dict of numpy arrays (+ `t`, `env`, `episode_return`); `records_as_list` gives the
        reference's pickled list-of-dicts form.  The training state is restored afterwards."""
        torch = self.torch
        dev = self.buf["rewards"].device
        snap = self._snapshot()
        self._fresh_episodes(self.seed + 0xC011EC7 if seed is None else seed)
        self.records_logits = torch.zeros((self.T, self.N, 7), dtype=torch.float32, device=dev)
        out = {k: [] for k in ("image", "direction", "mission", "policy", "t", "env", "episode_return")}
        total = 0
        try:
            for _ in range(max_rollouts):
                self.collect(deterministic=deterministic)
                b, T, N = self.buf, self.T, self.N
                t_end, i_end = torch.nonzero((b["ep_len"] != 0) & (b["rewards"] != 0), as_tuple=True)
                length = b["ep_len"][t_end, i_end].long()
                inside = t_end - length + 1 >= 0                                   # the whole episode is in this window
                t_end, i_end, length = t_end[inside], i_end[inside], length[inside]
                if t_end.numel():
                    # all (t, env) of those episodes: t in [t_end - length + 1, t_end]
                    rep = torch.repeat_interleave(torch.arange(t_end.numel(), device=dev), length)
                    first = torch.cumsum(length, 0) - length
                    t = t_end[rep] - length[rep] + 1 + (torch.arange(rep.numel(), device=dev) - first[rep])
                    i = i_end[rep]
                    image, onehot, mrow = gather_minibatch(b, t, i)
                    out["image"].append(image.cpu().numpy()); out["direction"].append(onehot.cpu().numpy())
                    out["mission"].append(self.policy.sequences[mrow].cpu().numpy())
                    out["policy"].append(torch.softmax(self.records_logits[t, i], dim=1).cpu().numpy())
                    out["t"].append(t.cpu().numpy()); out["env"].append(i.cpu().numpy())
                    out["episode_return"].append(b["rewards"][t_end, i_end][rep].cpu().numpy())
                    total += int(rep.numel())
                self.shift()
                if total >= min_records:
                    break
        finally:
            self.records_logits = None
            self._restore(snap)
        empty = {"image": (0, 12, 7, 7), "direction": (0, 16), "mission": (0, 128), "policy": (0, 7)}
        return {k: (np.concatenate(v) if v else np.zeros(empty.get(k, (0,)))) for k, v in out.items()}

    @staticmethod
    def records_as_list(records):
        """ppo.py:279-289: [{'image': ..., 'direction': ..., 'mission': ..., 'policy': ...}] of nested lists"""
        return [{k: records[k][j].tolist() for k in ("image", "direction", "mission", "policy")}
                for j in range(len(records["policy"]))]

    @staticmethod
    def save_records(records, path: str = "rollouts/data.pkl"):
        """the pickle file `test()` writes (ppo.py:267-292) and the distillation path reads"""
        import os
        import pickle
        os.makedirs(os.path.dirname(path) or ".", exist_ok=True)
        with open(path, "wb") as f:
            pickle.dump(RolloutEngine.records_as_list(records), f)

    def shift(self):
        """the last 4 frame slots of this rollout become the first 4 of the next"""
        b, T = self.buf, self.T
        for key in ("frames", "dirs", "mission"):
            b[key][0:4].copy_(b[key][T:T + 4].clone())

    def iteration(self, progress_remaining: float | None = None):
        """One PPO iteration.  Like SB3 (`_update_current_progress_remaining` before every `train()`), the learning rate
        follows `linear_schedule` of 1 - num_timesteps / total_timesteps AFTER the rollout; pass a value to override."""
        self.collect()
        n_boot = self.bootstrap_truncated()
        self.compute_advantages()
        if progress_remaining is None:
            progress_remaining = max(0.0, 1.0 - self.num_timesteps / float(self.cfg.total_timesteps))
        self.updater.set_progress(progress_remaining)
        n_mb = self.update()
        self.shift()
        return {"minibatches": n_mb, "bootstrapped": n_boot}
