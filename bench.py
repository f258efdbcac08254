#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json: env-steps/sec [+ PPO frames/sec] vs host-CPU baseline).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One bench "step" = one rollout pass over one batch: T=128 simulator steps of 65 536 GTO
environments per GPU (BASELINE.json configs[1]), actions = synthetic uniform u8 resident in HBM,
observations written to a [T,N,148] device rollout buffer (1.2 GB, larger than L2).  The headline
issues the rollout as one launch (mgrl_step_many); `per_step_launch` re-times it as T one-step launches.
Prints ONE JSON line (rank 0).  See DESIGN.md §Measurement for how each field is defined.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

N_ENVS = 65536          # per GPU (configs[1])
T_ROLLOUT = 128
TASK = "GTO"
SEED = 42
# HBM-compulsory bytes per env-step of the step kernel when the 140-B state stays on chip
# (L2 / shared memory): action 1 + image record 148 + dir 1 + mission 1 + reward 4 + term 1 + trunc 1 + ep_len 1
BYTES_PER_ENV_STEP = 158
# with the state streamed from/to HBM as well (N*140 B larger than L2): + 2*140
BYTES_PER_ENV_STEP_STREAMED = 158 + 280


def load_traffic(n, T):
    """measured DRAM bytes per launch of step_kernel from the committed ncu summary of this build (None if absent / other
    shape / another build of the kernel source)"""
    import hashlib
    try:
        with open(os.path.join(ROOT, "profiles", "k1_traffic.json")) as f:
            rec = json.load(f)
        src = b"".join(open(os.path.join(ROOT, "minigrid-rl_b200", "csrc", name), "rb").read()
                       for name in ("mgrl_core.cuh", "mgrl_kernels.cu"))
        if (rec["envs"], rec["steps"]) != (n, T) or rec["source_sha16"] != hashlib.sha256(src).hexdigest()[:16]:
            return None
        return float(rec["dram_bytes_read"]) + float(rec["dram_bytes_write"])
    except Exception:
        return None


def load_peaks():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""

    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index: int):
        self.rows, self.proc, self.index = [], None, index

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append((time.perf_counter(), [x.strip() for x in line.split(",")]))

    def count_since(self, t0):
        return sum(1 for t, _ in self.rows if t >= t0)

    def __exit__(self, *a):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except Exception:
                self.proc.kill()

    def summary(self, t0=0.0, t1=float("inf")):
        """median SM clock / reasons over the samples taken in [t0, t1] (while the workload ran)"""
        sm, smax, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for t, r in self.rows:
            if not (t0 <= t <= t1):
                continue
            try:
                sm.append(float(r[0])); smax = max(smax, float(r[1]))
                for name, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(name)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_baseline(n_envs: int, target_seconds: float, nthreads: int, max_steps: int = T_ROLLOUT):
    """The CPU oracle (C port of the reference path) on the host cores: same task, same batch."""
    import numpy as np
    from oracle import oracle as orc
    cfg = orc.make_config(problem="multi", mission=0)
    env = orc.OracleVecEnv(cfg, n_envs, seed=SEED, nthreads=nthreads)
    env.reset()
    rs = np.random.RandomState(0)
    acts = rs.randint(0, 7, size=(8, n_envs)).astype(np.uint8)
    env.step(acts[0], want_term_obs=False)
    steps, t0 = 0, time.perf_counter()
    while True:
        env.step(acts[steps % 8], want_term_obs=False)
        steps += 1
        el = time.perf_counter() - t0
        if el >= target_seconds or steps >= max_steps:
            break
    return n_envs * steps / el, steps, el


def run_reference(args):
    """`--impl reference`: the reference's CPU path.  The reference is pure Python on packages that
    are not installable here (minigrid, gymnasium, stable_baselines3: DESIGN.md), so this arm times
    the C oracle port of it with every host thread; each step is a bounded sample of the workload."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    nthreads = os.cpu_count() or 1
    per_step = []
    total = args.warmup + args.steps
    budget = max(1.0, min(8.0, 120.0 / max(total, 1)))
    sample_steps = 0
    for i in range(total):
        v, sample_steps, el = cpu_baseline(N_ENVS, budget, nthreads)
        if i >= args.warmup:
            per_step.append(v)
    value = statistics.mean(per_step)
    sample = f"{sample_steps} of {T_ROLLOUT} vector steps of {N_ENVS} GTO envs per bench step"
    line = {
        "impl": "reference", "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1000.0 * N_ENVS * T_ROLLOUT / value,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "u8", "data": "synthetic",
        "config": {"workload": f"{TASK} multi-room 11x11 (BASELINE configs[1]), {N_ENVS} envs/GPU, rollout {T_ROLLOUT} steps, "
                               "uniform random u8 actions resident in HBM",
                   "note": "C oracle port of the reference CPU path on the host cores (the reference itself needs minigrid / "
                           "gymnasium / SB3, absent); outputs: un-stacked observation, reward, flags, like the GPU arm's e2e"},
        "cpu_baseline": {"value": value, "unit": "env-steps/s", "cores": nthreads, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": "env-steps/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--envs", type=int, default=N_ENVS)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-ppo", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the PKP 32 768 / ALL 131 072 sub-records")
    ap.add_argument("--ppo-iters", type=int, default=2)
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    import torch
    import minigrid_rl_b200 as mg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product has no CPU path "
                         "(use --impl reference for the CPU baseline arm)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1 and hasattr(os, "sched_setaffinity") and os.environ.get("MGRL_BENCH_PIN", "1") != "0":
        # one slice of the host cores per rank: the launch thread of a rank (and the host threads of its e2e path) do not
        # migrate onto another rank's cores
        cores = sorted(os.sched_getaffinity(0))
        local_world = int(os.environ.get("LOCAL_WORLD_SIZE", str(world)))
        per = max(1, len(cores) // local_world)
        mine = cores[local_rank * per:(local_rank + 1) * per] or cores
        os.sched_setaffinity(0, mine)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    n, T = args.envs, T_ROLLOUT
    if args.warmup < 3:
        print(f"bench.py: warm-up raised from {args.warmup} to 3 (timing rules: W >= 3)", file=sys.stderr)
    W, K = max(args.warmup, 3), args.steps
    peak, peak_src = load_peaks()
    u8 = dict(dtype=torch.uint8, device=dev)

    def barrier():
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, k):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        e0.record()
        for _ in range(k):
            fn()
        e1.record()
        barrier()
        ms = e0.elapsed_time(e1)
        if dist is not None:
            tmax = torch.tensor([ms], device=dev)
            dist.all_reduce(tmax, op=dist.ReduceOp.MAX)
            ms = float(tmax.item())
        return ms

    def make_rollout(task, n_envs):
        """environments + synthetic actions + [T,N,...] output buffers of one K1 workload"""
        e = mg.DeviceEnv(mg.EnvConfig.for_task(task), num_envs=n_envs, seed=SEED, env_id_base=rank * n_envs, layout="hwc148")
        e.reset()
        gen = torch.Generator(device=dev).manual_seed(1234 + rank)
        acts = torch.randint(0, 7, (T, n_envs), dtype=torch.uint8, device=dev, generator=gen)
        bufs = (torch.empty((T, n_envs, 148), **u8), torch.empty((T, n_envs), **u8), torch.empty((T, n_envs), **u8),
                torch.empty((T, n_envs), dtype=torch.float32, device=dev), torch.empty((T, n_envs), **u8),
                torch.empty((T, n_envs), **u8), torch.empty((T, n_envs), **u8))
        return e, acts, bufs

    def k1_record(task, n_envs, k, w):
        """K1 (`mgrl_step_many`, one launch = one rollout of T steps) on `n_envs` environments of `task` per GPU"""
        e, acts, bufs = make_rollout(task, n_envs)
        for _ in range(w):
            e.step_many(acts, *bufs)
        t_ms = timed(lambda: e.step_many(acts, *bufs), k)
        flags = e.error_flags()
        e.close()
        us = 1000.0 * t_ms / k
        gbs = BYTES_PER_ENV_STEP * n_envs * T / (us * 1e-6) / 1e9
        return {"task": task, "envs_per_gpu": n_envs, "value": world * n_envs * T * k / (t_ms / 1000.0), "unit": "env-steps/s",
                "us_per_launch": us, "achieved_gbs": gbs, "frac": gbs / peak, "env_error_flags": flags}

    env, actions, (image, dirs, mis, rew, term, trunc, eplen) = make_rollout(TASK, n)

    def rollout_many():  # the rollout as ONE launch: T steps per launch, state tile resident in shared memory
        env.step_many(actions, image, dirs, mis, rew, term, trunc, eplen)

    def rollout_steps():  # one launch per env step, the way a policy-in-the-loop rollout has to issue it
        for t in range(T):
            env.step(actions[t], image[t], dirs[t], mis[t], rew[t], term[t], trunc[t], eplen[t])

    # W warm-up bench steps also de-synchronise the episode phases (>= 121 env steps)
    with ClockSampler(local_rank) as clocks:
        for _ in range(W):
            rollout_many()
        t_load0 = time.perf_counter()
        ms = timed(rollout_many, K)
        # the timed region is a few milliseconds; keep the identical load running (untimed) until
        # nvidia-smi (100 ms period) has sampled it a few times, so the clock record is meaningful
        while rank == 0 and clocks.proc and clocks.count_since(t_load0) < 8 and time.perf_counter() - t_load0 < 4.0:
            rollout_many()
            torch.cuda.synchronize()
        t_load1 = time.perf_counter()
        barrier()
    clock_summary = clocks.summary(t_load0, t_load1)
    total_env_steps = world * n * T * K
    value = total_env_steps / (ms / 1000.0)
    launches = K
    us_per_launch = 1000.0 * ms / launches
    achieved = BYTES_PER_ENV_STEP * n * T / (us_per_launch * 1e-6) / 1e9

    # the same rollout issued as T one-step launches (what a policy in the loop needs)
    for _ in range(2):
        rollout_steps()
    ms_steps = timed(rollout_steps, K)
    steps_value = world * n * T * K / (ms_steps / 1000.0)
    err = env.error_flags()

    # end to end through the C ABI with HOST buffers, H2D + D2H inside the timed region.
    #  e2e          mgrl_vec_step_frames_host: pinned actions in, un-stacked observation + reward + flags out -- the
    #               output set of the CPU arm (mg_vec_step), so the two arms are like for like;
    #  e2e.stacked  mgrl_vec_step_stacked_host: the full SB3 drop-in (frame stack x4, one-hot direction, int64 mission
    #               tokens, stacked terminal observations), 64-byte records over PCIe and the 1.6 KB observation dict
    #               maintained in place by the library's host threads (MGRL_WIRE=0: mgrl_vec_step_host, the device-side
    #               stack copied out in full, 1761 B per env-step over PCIe)
    e2e = None
    if not args.no_e2e:
        import numpy as np
        acts_h = np.random.RandomState(rank).randint(0, 7, size=(16, n)).astype(np.uint8)
        e2e_steps = 256      # (64 steps are 10 ms: too short next to the host threads' wake-up and the split controller's probes)

        def time_host(step_fn):
            for i in range(16):
                step_fn(acts_h[i % 16])
            barrier()
            t0 = time.perf_counter()
            for i in range(e2e_steps):
                step_fn(acts_h[i % 16])
            torch.cuda.synchronize()
            el = time.perf_counter() - t0
            if dist is not None:
                tm = torch.tensor([el], device=dev)
                dist.all_reduce(tm, op=dist.ReduceOp.MAX)
                el = float(tm.item())
            return world * n * e2e_steps / el

        cfg = mg.EnvConfig.for_task(TASK)
        venv = mg.B200VecEnv(cfg, num_envs=n, seed=SEED, device=local_rank, env_id_base=rank * n, layout="hwc148")
        venv.reset_frames()
        v_frames = time_host(venv.step_frames)
        venv.close()
        venv = mg.B200VecEnv(cfg, num_envs=n, seed=SEED, device=local_rank, env_id_base=rank * n)
        venv.reset()
        v_stacked = time_host(venv.step_arrays)
        # the SB3 protocol itself: VecEnv.step with per-environment info dicts (terminal observations, Monitor episodes)
        e2e_steps, saved = 8, e2e_steps
        v_vecenv = time_host(lambda a: venv.step(a))
        e2e_steps = saved
        venv.close()
        wire = os.environ.get("MGRL_WIRE", "1") != "0"
        e2e = {"value": v_frames, "unit": "env-steps/s",
               "h2d_bytes_per_step": n * 1 * T, "d2h_bytes_per_step": n * (64 if wire else 148 + 1 + 1 + 4 + 1 + 1 + 1) * T,
               "api": "B200VecEnv.step_frames -> mgrl_vec_step_frames_host (pinned numpy in/out, un-stacked observation: "
                      "the outputs of the CPU arm's vector step)"
                      + ("; PCIe wire format = one 64-byte record per environment (a code byte per view cell + the step's "
                         "scalars), expanded into the 148-byte observation records, rewards and flags by the library's host "
                         "threads while later chunks are in flight (format conversion only); when a GPU has few host threads part of "
                         "the batch's images is copied directly instead (148 + 16 B per environment; MGRL_WIRE_DIRECT=auto), "
                         "d2h_bytes_per_step is the all-records figure" if wire else ""),
               "sample": f"{e2e_steps} vector steps of {n} envs per rank",
               "stacked": {"value": v_stacked, "unit": "env-steps/s",
                           "d2h_bytes_per_step": n * (2 * 64 if wire else 4 * 147 + 16 + 128 * 8 + 4 + 1 + 1 + 1 + 147 + 1) * T,
                           "api": ("B200VecEnv.step_arrays -> mgrl_vec_step_stacked_host (SB3 observation dict: 4-frame stack, one-hot "
                                   "direction, int64 mission tokens, stacked terminal observations; two 64-byte records per environment "
                                   "over PCIe, the dict updated in place in pinned host memory by the library's host threads, like "
                                   "VecFrameStack.stacked_obs)" if wire else
                                   "B200VecEnv.step_arrays -> mgrl_vec_step_host (SB3 observation dict: 4-frame stack, "
                                   "one-hot direction, int64 mission tokens; device-side stack copied out in full)")},
               "vecenv_step": {"value": v_vecenv, "unit": "env-steps/s",
                               "api": "B200VecEnv.step (the SB3 VecEnv protocol: stacked observation dict + one Python info dict "
                                      "per finished environment); host-side Python, 8 vector steps"}}

    # PPO frames/sec (SB3 `time/fps`: env frames per wall second over rollout + update), BASELINE.json's second figure:
    # policy-in-the-loop rollout (2 launches per step) + GAE + n_epochs of minibatch updates, gradients all-reduced
    def ppo_record(task, n_envs, iters, with_kernel=False):
        penv = mg.DeviceEnv(mg.EnvConfig.for_task(task), num_envs=n_envs, seed=SEED, env_id_base=rank * n_envs, layout="hwc148")
        pcfg = mg.PPOConfig(n_steps=T, batch_size=n_envs * T // 32, n_epochs=4, update_tf32=True)
        eng = mg.RolloutEngine(penv, mg.Policy(dev, seed=SEED), pcfg, dist=dist if world > 1 else None, seed=SEED)
        for _ in range(2):                                   # warm-up: allocator, layouts in L2, graph capture of the
            eng.iteration(1.0)                               # optimizer step
        eng.warm_bootstrap()                                 # the truncation bootstrap's lazily loaded kernels (truncations are rare)
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        t_roll = t_upd = 0.0
        barrier()
        for it in range(iters):
            ev[0].record()
            h0 = time.perf_counter()
            eng.collect()
            t_host = (time.perf_counter() - h0) * 1e3 if it == 0 else max(t_host, (time.perf_counter() - h0) * 1e3)   # host time to ISSUE the rollout
            if os.environ.get("MGRL_BENCH_DEBUG"):
                dbg = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
                dbg[0].record(); nb = eng.bootstrap_truncated(); dbg[1].record(); eng.compute_advantages(); dbg[2].record()
                torch.cuda.synchronize()
                print(f"[rank {rank}] it {it}: collect {ev[0].elapsed_time(dbg[0]):.2f} ms, bootstrap of {nb} {dbg[0].elapsed_time(dbg[1]):.2f} ms, "
                      f"advantages {dbg[1].elapsed_time(dbg[2]):.2f} ms", file=sys.stderr)
            else:
                eng.bootstrap_truncated(); eng.compute_advantages()
            ev[1].record()
            eng.updater.set_progress(1.0 - it / max(iters, 1))
            n_mb = eng.update()
            eng.shift()
            ev[2].record()
            torch.cuda.synchronize()
            t_roll += ev[0].elapsed_time(ev[1]); t_upd += ev[1].elapsed_time(ev[2])
        barrier()
        if os.environ.get("MGRL_BENCH_DEBUG"):
            print(f"[rank {rank}] {task} {n_envs}: rollout {t_roll / iters:.2f} ms (host issue {t_host:.2f} ms), update {t_upd / iters:.2f} ms", file=sys.stderr)
        # the ranks meet at every optimizer step (gradient all-reduce), so a rank that is late in its rollout shows up as a wait
        # in the other ranks' update: the iteration time is the MAX over ranks of a rank's own rollout + update, not the sum of
        # the two per-phase maxima (which would count that skew twice)
        tt = torch.tensor([t_roll, t_upd, t_roll + t_upd], device=dev)
        if dist is not None:
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
        t_roll, t_upd, t_iter = float(tt[0]), float(tt[1]), float(tt[2])
        frames = world * n_envs * T * iters
        rec = {"value": frames / (t_iter / 1000.0), "unit": "frames/s",
               "rollout_env_steps_per_s": frames / (t_roll / 1000.0), "rollout_ms": t_roll / iters,
               "update_ms": t_upd / iters, "iteration_ms": t_iter / iters, "iterations": iters,
               "config": {"task": task, "n_steps": T, "n_envs_per_gpu": n_envs, "batch_size_per_gpu": pcfg.batch_size,
                          "n_epochs": 4, "minibatches_per_iteration": n_mb, "policy": "CustomPPOPolicy 110216 params, fp32",
                          "rollout": "mgrl_policy_forward + mgrl_step per step (hand-written kernels)",
                          "update": eng.updater.describe(),
                          "all_reduces_per_optimizer_step": eng.updater.all_reduces_per_step()},
               "env_error_flags": penv.error_flags()}
        if with_kernel:
            # the rollout's policy kernel alone (K3): T launches over the collected frames, CUDA events on the launching stream
            B = eng.buf
            kev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
            kev[0].record()
            for t in range(T):
                eng.policy.forward_rollout(B["frames"], B["dirs"], B["mission"][t + 3], t + 3, eng.prev_age, eng.prev_done,
                                           B["age"][t], B["values"][t], B["actions"][t], B["logp"][t])
            kev[1].record()
            torch.cuda.synchronize()
            k3_us = 1000.0 * kev[0].elapsed_time(kev[1]) / T
            k3_flop = 2.0 * 79616 * n_envs   # MACs / observation: conv 27 648 + 8 192 + 8 192, direction 256, MLPs 26 624 + 8 192, heads 512
            rec["policy_kernel"] = {"us_per_launch": k3_us, "observations_per_launch": n_envs,
                                    "kernel": eng.policy.kernel_name(),
                                    "algorithmic_tflops": k3_flop / (k3_us * 1e-6) / 1e12}
        eng.updater.release()            # a captured optimizer step must not outlive the communicator
        penv.close()
        del eng
        torch.cuda.empty_cache()
        return rec

    def gae_record():
        """K4 (`mgrl_gae`) at the headline shape: T x N float32 rollouts, 17 algorithmic bytes per element (reward 4 + value 4
        + episode start 1 read, advantage 4 + return 4 written).  Three buffer sets (3 x 142 MB > L2) are cycled."""
        sets = []
        for _ in range(3):
            sets.append((torch.rand((T, n), device=dev), torch.randn((T, n), device=dev),
                         (torch.rand((T, n), device=dev) < 0.14).to(torch.uint8), torch.randn(n, device=dev),
                         torch.zeros(n, **u8), torch.empty((T, n), device=dev), torch.empty((T, n), device=dev)))
        calls = [0]

        def one():
            r, v, st, lv, ld, adv, ret = sets[calls[0] % 3]
            mg.vec_env.gae(r, v, st, lv, ld, 0.81, 0.945, adv, ret)
            calls[0] += 1

        for _ in range(3):
            one()
        reps = 30
        t_ms = timed(one, reps)
        us = 1000.0 * t_ms / reps
        gbs = (17.0 * T * n + 5.0 * n) / (us * 1e-6) / 1e9
        return {"kernel": "gae_kernel", "elements_per_launch": T * n, "bytes_per_element": 17, "us_per_launch": us,
                "elements_per_s": world * T * n / (us * 1e-6), "achieved_gbs": gbs, "peak_gbs": peak, "frac": gbs / peak,
                "l2": "three 142 MB buffer sets cycled (> 126 MB L2)"}

    ppo_line = None
    gae_line = None
    cfg_lines = {}
    env.close()
    del image, dirs, mis, rew, term, trunc, eplen, actions
    torch.cuda.empty_cache()
    gae_line = gae_record()
    if not args.no_ppo:
        ppo_line = ppo_record(TASK, n, args.ppo_iters, with_kernel=True)
    # BASELINE.json configs[2] (PKP, 262 144 envs over 8 GPUs = 32 768 per GPU) and configs[4] (ALL, 1 048 576 over 8 = 131 072
    # per GPU, NCCL gradient all-reduce) at their per-GPU sizes: K1 alone and the PPO iteration
    if not args.no_configs and n == N_ENVS:
        for key, task, n_c in (("pkp_32768", "PKP", 32768), ("all_131072", "ALL", 131072)):
            rec = k1_record(task, n_c, max(3, K // 2), 3)
            if not args.no_ppo:
                rec["ppo"] = ppo_record(task, n_c, args.ppo_iters)
            cfg_lines[key] = rec

    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        nthreads = os.cpu_count() or 1
        v, s, el = cpu_baseline(n, 12.0, nthreads, max_steps=8 * T)     # eight rollouts: 20-30 core-seconds of CPU work
        cpu = {"value": v, "unit": "env-steps/s", "cores": nthreads, "kind": "port",
               "sample": f"{s} vector steps of {n} GTO envs ({el:.1f} s) through oracle/mg_oracle.c"}

    if rank == 0:
        line = {
            "metric": "env-steps/sec", "value": value, "unit": "env-steps/s", "n_gpus": world, "steps": K,
            "warmup": W, "ms_per_step": ms / K, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "u8", "data": "synthetic",
            "config": {"workload": f"{TASK} multi-room 11x11 (BASELINE configs[1]), {n} envs/GPU, rollout {T} steps, "
                                   "uniform random u8 actions resident in HBM",
                       "l2": "outputs stream into a 1.2 GB [T,N,148] rollout buffer (> 126 MB L2); the 9.2 MB state "
                             "array and the 27.5 MB of prepared layouts are L2-resident by design and are not counted "
                             "in the algorithmic bytes",
                       "path": "mgrl_step_many: one launch = one rollout of T steps (actions known up front)"},
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                         # dram__bytes_read.sum + dram__bytes_write.sum of one launch of this kernel at this shape, read from
                         # the ncu --set full summary of the CURRENT build (profiles/k1_traffic.json, written by
                         # profiles/ncu_summary.py); null when the shape differs or no capture of this build exists
                         "traffic": load_traffic(n, T), "traffic_unit": "bytes/launch",
                         "algorithmic_bytes_per_launch": BYTES_PER_ENV_STEP * n * T,
                         "kernel": ("rollout_kernel<HWC148,see_through,GEN_MULTI_PLAIN> (14 step warps + 10 generator warps per SM, TMA bulk "
                                    "store of the staged records; the instance compiled for multi-room problems without obstacles)" if os.environ.get("MGRL_ROLLOUT", "1") != "0"
                                    else "step_kernel<HWC148,see_through,64,1>"), "peak_source": peak_src,
                         "bytes_per_env_step": BYTES_PER_ENV_STEP, "us_per_launch": us_per_launch,
                         "env_steps_per_launch": n * T},
            "per_step_launch": {"value": steps_value, "unit": "env-steps/s", "ms_per_step": ms_steps / K,
                                "us_per_launch": 1000.0 * ms_steps / (K * T),
                                "achieved_gbs": BYTES_PER_ENV_STEP * steps_value / world / 1e9,
                                "frac": BYTES_PER_ENV_STEP * steps_value / world / 1e9 / peak,
                                "note": "mgrl_step x T: the same kernel with one step per launch, as a rollout with a "
                                        "policy in the loop issues it (latency-bound at this batch size)"},
            "ppo": ppo_line,
            # PPO frames/s over all GPUs: the path that communicates (one gradient all-reduce per optimizer step)
            "ppo_frames_per_s": None if ppo_line is None else ppo_line["value"],
            "gae": gae_line,
            "configs": cfg_lines,
            "cpu_baseline": cpu, "e2e": e2e, "gpu_launches": launches, "clocks": clock_summary,
            "env_error_flags": err,
        }
        print(json.dumps(line))
    if dist is not None:
        torch.cuda.synchronize()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
