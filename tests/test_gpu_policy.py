"""GPU tier for K3 and the rollout engine: the hand-written fused forward kernel (through the C ABI) against the torch-CPU
fp32 oracle that sees SB3-style stacked observations, the sampling rule, the frame-stack gather across episode
boundaries, truncation bootstrap, GAE on the rollout and one PPO iteration.  Tolerance for floating point
(BASELINE.json north_star: 1e-5 relative in fp32): max |got - want| <= 1e-5 * max |want| over the batch, i.e. relative
to the scale of the logits / values (an element-wise ratio is meaningless for logits that cross zero).  `want` is the
oracle network evaluated in FLOAT64 (policy_oracle.double_copy): the oracle's own float32 evaluation is itself up to
~1e-5 away from it, so comparing two float32 evaluations with each other would measure the sum of two rounding errors.
The bar is 1e-5 for every step and every batch size, with no exceptions (profiles/policy_error.py prints the measured
errors of the kernel and of torch-CPU float32 against the same yardstick)."""
import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = pytest.mark.gpu

import minigrid_rl_b200 as mg  # noqa: E402
from minigrid_rl_b200 import policy as pol, ppo  # noqa: E402
from oracle import oracle as orc, policy_oracle as po, sb3_oracle  # noqa: E402

REL = 1e-5
VALUE_SCALE = 0.25      # typical max |V| of a batch under init_reference(seed 5) (printed by profiles/policy_error.py)


def close(got, want, rel=REL):
    got, want = got.double(), want.double()
    return float((got - want).abs().max()) <= rel * float(want.abs().max())


def make_engine(n, T, task="ALL", seed=11, **kw):
    env = mg.DeviceEnv(mg.EnvConfig.for_task(task), num_envs=n, seed=seed, layout="hwc148", env_id_base=1000)
    o = po.OraclePolicy()
    po.init_reference(o, 5)
    p = pol.Policy("cuda", seed=1)
    p.load_oracle(o)
    cfg = ppo.PPOConfig(n_steps=T, batch_size=n * T // 4, **kw)
    return ppo.RolloutEngine(env, p, cfg, seed=77), o


def oracle_stacks(buf, T, n):
    """SB3 stacks rebuilt independently of the kernel's `age`: FrameStack driven by the done flags."""
    tt = mg.token_table()
    fs_img = sb3_oracle.FrameStack(n, (3, 7, 7), np.uint8)
    fs_dir = sb3_oracle.FrameStack(n, (4,), np.uint8)
    fs_mis = sb3_oracle.FrameStack(n, (32,), np.int64)
    frames = buf["frames"].cpu().numpy()[:, :, :147].reshape(T + 4, n, 7, 7, 3).transpose(0, 1, 4, 2, 3)
    dirs, mission, done = buf["dirs"].cpu().numpy(), buf["mission"].cpu().numpy(), buf["ep_len"].cpu().numpy() != 0
    out = []
    for t in range(T + 1):
        b = t + 3
        if t == 0:
            s = (fs_img.reset(frames[b]), fs_dir.reset(sb3_oracle.one_hot_dir(dirs[b])), fs_mis.reset(tt[mission[b]]))
        else:
            z = np.zeros_like(frames[b])
            s = (fs_img.update(frames[b], done[t - 1], z)[0], fs_dir.update(sb3_oracle.one_hot_dir(dirs[b]), done[t - 1],
                 np.zeros((n, 4), np.uint8))[0], fs_mis.update(tt[mission[b]], done[t - 1], np.zeros((n, 32), np.int64))[0])
        out.append(s)
    return out


def test_forward_kernel_matches_oracle_on_a_rollout():
    n, T = 640, 20                       # 640 = 10 CTAs of 64 observations
    eng, o = make_engine(n, T)
    o64 = po.double_copy(o)
    eng.collect()
    b = eng.buf
    stacks = oracle_stacks(b, T, n)
    logits = torch.zeros((n, 7), device="cuda")
    val = torch.zeros(n, device="cuda")
    act = torch.zeros(n, dtype=torch.uint8, device="cuda")
    lp = torch.zeros(n, device="cuda")
    age = torch.zeros(n, dtype=torch.uint8, device="cuda")
    ages_seen = set()
    for t in range(T + 1):
        prev_age = None if t == 0 else b["age"][t - 1]
        prev_done = None if t == 0 else b["ep_len"][t - 1]
        eng.policy.forward_rollout(b["frames"], b["dirs"], b["mission"][t + 3], t + 3, prev_age, prev_done, age, val, act, lp,
                                   logits=logits, seed=77, env_id_base=1000, step=t)
        img, d, mis = stacks[t]
        with torch.no_grad():
            lo, vo = o64({"direction": torch.from_numpy(d), "image": torch.from_numpy(img), "mission": torch.from_numpy(mis)})
        assert close(logits.cpu(), lo), t
        assert close(val.cpu(), vo), t
        assert torch.equal(age, b["age"][t]) and torch.allclose(val, b["values"][t], rtol=0, atol=0)
        if t < T:   # the rollout stored exactly what this call recomputes (same Philox key: seed, env id, step)
            assert torch.equal(act, b["actions"][t]) and torch.equal(lp, b["logp"][t])
        ages_seen |= set(age.cpu().numpy().tolist())
        # sampling rule: inverse CDF of softmax(logits) on the uniform of Philox(seed, env, step)
        u = np.empty(n, np.float32)
        out4 = (C.c_uint32 * 4)()
        for i in range(0, n, 37):
            ctr = (C.c_uint32 * 4)(t, 0x504F4C49, (1000 + i) & 0xFFFFFFFF, 0)
            key = (C.c_uint32 * 2)(77, 0)
            orc.lib().mg_philox4x32_10(ctr, key, out4)
            u_i = np.float32(out4[0] >> 8) * np.float32(1.0 / 16777216.0)
            a_ref, lp_ref = po.sample_inverse_cdf(logits[i:i + 1].cpu().numpy(), np.array([u_i], np.float32))
            p_row = torch.softmax(logits[i].cpu(), 0).numpy()
            c = np.cumsum(p_row)
            near = np.min(np.abs(c - u_i)) < 1e-6          # a uniform that sits on a CDF boundary may round either way
            assert near or int(act[i]) == int(a_ref[0]), (t, i)
            if int(act[i]) == int(a_ref[0]):
                assert abs(float(lp[i]) - float(lp_ref[0])) <= 1e-5 * max(1.0, abs(float(lp_ref[0])))
    assert ages_seen == {0, 1, 2, 3}
    assert int((b["ep_len"] != 0).sum()) > n               # many episode boundaries were crossed
    eng.env.close()


def test_gather_minibatch_equals_sb3_stacks():
    n, T = 256, 16
    eng, _ = make_engine(n, T)
    eng.collect()
    stacks = oracle_stacks(eng.buf, T, n)
    tt = mg.token_table()
    t = torch.arange(T, device="cuda").repeat_interleave(n)
    i = torch.arange(n, device="cuda").repeat(T)
    image, onehot, mrow = ppo.gather_minibatch(eng.buf, t, i)
    want_img = np.concatenate([stacks[k][0] for k in range(T)])
    want_dir = np.concatenate([stacks[k][1] for k in range(T)])
    want_mis = np.concatenate([stacks[k][2] for k in range(T)])
    assert np.array_equal(image.cpu().numpy(), want_img) and np.array_equal(onehot.cpu().numpy(), want_dir)
    seq = eng.policy.sequences.cpu().numpy()
    assert np.array_equal(seq[mrow.cpu().numpy()], want_mis)
    eng.env.close()


def test_rollout_advantages_and_bootstrap_match_oracle():
    n, T = 512, 40
    eng, o = make_engine(n, T, task="GTG")
    # long-lived episodes so that truncations happen: steer away from 'done' by overwriting sampled actions
    b = eng.buf
    for rep in range(4):                    # 160 steps > max_steps 121
        eng.collect()
        eng.shift()
    eng.collect()
    rewards_before = b["rewards"].clone()
    n_boot = eng.bootstrap_truncated()
    trunc_only = ((b["trunc"] != 0) & (b["term"] == 0)).cpu().numpy()
    assert n_boot == int(trunc_only.sum())
    if n_boot:
        stacks = oracle_stacks(b, T, n)
        tfr = eng.term_frames.cpu().numpy()[:, :, :147].reshape(T, n, 7, 7, 3).transpose(0, 1, 4, 2, 3)
        tdir = eng.term_dirs.cpu().numpy()
        tt = mg.token_table()
        for t, i in np.argwhere(trunc_only)[:50]:
            img, d, mis = stacks[t]
            timg = np.concatenate([img[i, 3:], tfr[t, i]], 0)[None]
            td = np.concatenate([d[i, 4:], sb3_oracle.one_hot_dir(tdir[t, i:i + 1])[0]])[None]
            tm = np.concatenate([mis[i, 32:], tt[b["mission"][t + 3, i].item()]])[None]
            with torch.no_grad():
                _, v = o({"direction": torch.from_numpy(td), "image": torch.from_numpy(timg), "mission": torch.from_numpy(tm)})
            want = rewards_before[t, i].item() + np.float32(eng.cfg.gamma) * v.item()
            assert abs(b["rewards"][t, i].item() - want) <= 1e-5 * max(1.0, abs(want)), (t, i)
    eng.compute_advantages()
    wa, wr = orc.gae(b["rewards"].cpu().numpy(), b["values"][:T].cpu().numpy(), b["start"][:T].cpu().numpy(),
                     b["values"][T].cpu().numpy(), b["start"][T].cpu().numpy(), eng.cfg.gamma, eng.cfg.gae_lambda)
    assert np.array_equal(b["adv"].cpu().numpy().view(np.uint32), wa.view(np.uint32))
    assert np.array_equal(b["ret"].cpu().numpy().view(np.uint32), wr.view(np.uint32))
    # episode_start flags = dones of the previous step
    assert np.array_equal(b["start"][1:T + 1].cpu().numpy() != 0, b["ep_len"].cpu().numpy() != 0)
    eng.env.close()


def test_one_ppo_iteration_updates_the_policy_and_matches_cpu_update():
    n, T = 256, 16
    eng, o = make_engine(n, T, n_epochs=1, native_update=False)
    before = {k: v.detach().clone() for k, v in eng.policy.params.items()}
    eng.collect(); eng.bootstrap_truncated(); eng.compute_advantages()
    # the same minibatch through the GPU updater and through the oracle network + torch.optim on the CPU
    b = eng.buf
    t = torch.arange(T, device="cuda").repeat_interleave(n)
    i = torch.arange(n, device="cuda").repeat(T)
    image, onehot, mrow = ppo.gather_minibatch(b, t, i)
    args = (b["actions"][t, i], b["values"][t, i], b["logp"][t, i], b["adv"][t, i], b["ret"][t, i])
    loss, _ = eng.updater.minibatch(image, onehot, mrow, *args)
    stacks = oracle_stacks(b, T, n)
    obs = {"direction": torch.from_numpy(np.concatenate([s[1] for s in stacks[:T]])),
           "image": torch.from_numpy(np.concatenate([s[0] for s in stacks[:T]])),
           "mission": torch.from_numpy(np.concatenate([s[2] for s in stacks[:T]]))}
    cfg = eng.cfg
    lo, _ = po.ppo_loss(o, obs, *[a.cpu() for a in args], cfg.clip_range, cfg.clip_range_vf, cfg.ent_coef, cfg.vf_coef)
    assert abs(loss.item() - lo.item()) <= 1e-4 * max(1.0, abs(lo.item()))
    opt = torch.optim.Adam(o.parameters(), lr=cfg.initial_learning_rate, eps=cfg.optim_eps)
    lo.backward()
    torch.nn.utils.clip_grad_norm_(o.parameters(), cfg.max_grad_norm)
    opt.step()
    k = "mlp_extractor.policy_net.0.weight"
    assert not torch.equal(before[k], eng.policy.params[k])
    assert torch.allclose(eng.policy.params[k].detach().cpu(), o.pi[0].weight.detach(), rtol=1e-3, atol=2e-5)
    assert torch.allclose(eng.policy.params["features_extractor.extractors.image.image_Conv2d_0.weight"].detach().cpu(),
                          o.image[0].weight.detach(), rtol=1e-3, atol=2e-5)
    # and a full iteration runs end to end
    stats = eng.iteration(0.5)
    assert stats["minibatches"] == 4
    for v in eng.policy.params.values():
        assert torch.isfinite(v).all()
    assert eng.env.error_flags() == 0
    eng.env.close()


def test_evaluate_counts_the_first_episodes_of_every_env_and_restores_the_training_state():
    """evaluate_policy semantics (ppo.py:161): after a reset, exactly the first k episodes of every environment count; the
    training state (environments, frame history, counters) is untouched."""
    n, T = 512, 24
    eng, _ = make_engine(n, T)
    eng.collect(); eng.shift()
    before = (eng.env.get_state().clone(), eng.buf["frames"][0:4].clone(), eng.global_step, eng.num_timesteps, eng.env.seed,
              eng.prev_age.clone())
    stats = eng.evaluate(2, deterministic=False)
    assert stats["episodes"] == 2 * n and 0.0 <= stats["success_rate"] <= 1.0 and 1.0 <= stats["mean_length"] <= 121.0
    assert torch.equal(eng.env.get_state(), before[0]) and torch.equal(eng.buf["frames"][0:4], before[1])
    assert (eng.global_step, eng.num_timesteps, eng.env.seed) == before[2:5] and torch.equal(eng.prev_age, before[5])
    # the continued rollout is the one that would have been collected without the evaluation
    eng2, _ = make_engine(n, T)
    eng2.collect(); eng2.shift()
    eng.collect(); eng2.collect()
    for k in ("frames", "actions", "rewards", "ep_len"):
        assert torch.equal(eng.buf[k], eng2.buf[k]), k
    eng.env.close(); eng2.env.close()


def test_collect_records_matches_the_reference_record_format():
    """ppo.py:214-262,279-289: records = every step of the episodes that ended with a reward: the stacked observation the
    policy saw (rebuilt independently with the numpy SB3 wrappers) and its action probabilities (float64 oracle, 1e-5)."""
    n, T = 256, 40
    eng, o = make_engine(n, T, task="GTO")
    o64 = po.double_copy(o)
    rec = eng.collect_records(min_records=50, deterministic=False, seed=123)
    R = len(rec["policy"])
    assert R >= 50 and rec["image"].shape == (R, 12, 7, 7) and rec["direction"].shape == (R, 16)
    assert rec["mission"].shape == (R, 128) and rec["mission"].dtype == np.int64 and rec["policy"].shape == (R, 7)
    assert np.allclose(rec["policy"].sum(1), 1.0, atol=1e-5) and (rec["episode_return"] != 0).all()
    # the same rollout again, with the whole buffer at hand
    eng._fresh_episodes(123)
    eng.records_logits = torch.zeros((T, n, 7), device="cuda")
    eng.collect(deterministic=False)
    b = eng.buf
    stacks = oracle_stacks(b, T, n)
    ep_len, rew = b["ep_len"].cpu().numpy(), b["rewards"].cpu().numpy()
    want = set()
    for t_end, i in zip(*np.nonzero((ep_len != 0) & (rew != 0))):
        if t_end - int(ep_len[t_end, i]) + 1 >= 0:
            want |= {(t, i) for t in range(t_end - int(ep_len[t_end, i]) + 1, t_end + 1)}
    first = rec["t"] < T       # (all records of the first window)
    got = set(zip(rec["t"][:len(want)].tolist(), rec["env"][:len(want)].tolist()))
    assert got == want and first[:len(want)].all()
    for j in range(0, len(want), 7):
        t, i = int(rec["t"][j]), int(rec["env"][j])
        img, d, mis = stacks[t]
        assert np.array_equal(rec["image"][j], img[i]) and np.array_equal(rec["direction"][j], d[i]) and np.array_equal(rec["mission"][j], mis[i])
        with torch.no_grad():
            lo, _ = o64({"direction": torch.from_numpy(d[i:i + 1]), "image": torch.from_numpy(img[i:i + 1]), "mission": torch.from_numpy(mis[i:i + 1])})
        probs = torch.softmax(lo, 1)[0].numpy()
        assert np.abs(rec["policy"][j] - probs).max() <= 1e-5
    lst = eng.records_as_list({k: v[:3] for k, v in rec.items()})
    assert sorted(lst[0]) == ["direction", "image", "mission", "policy"] and len(lst[0]["image"]) == 12 and len(lst[0]["policy"]) == 7
    eng.records_logits = None
    eng.env.close()


def test_checkpoint_written_here_reproduces_the_kernel_logits_in_the_reference_network(tmp_path):
    """f3: SB3 archive -> reference network (oracle, under SB3's names) -> the logits / values the rollout kernel computes."""
    import io
    import zipfile
    from tests.test_policy_cpu import oracle_from_sb3_state_dict
    n, T = 384, 6
    eng, _ = make_engine(n, T)
    eng.iteration()                                     # weights that are not the initial ones
    path = str(tmp_path / "b200.zip")
    eng.policy.save_sb3_zip(path)
    with zipfile.ZipFile(path) as z:
        sd = torch.load(io.BytesIO(z.read("policy.pth")), map_location="cpu")
    o64 = po.double_copy(oracle_from_sb3_state_dict(sd))
    eng.reset()                                         # (oracle_stacks rebuilds the stacks of a rollout that starts at a reset)
    eng.collect()
    b = eng.buf
    stacks = oracle_stacks(b, T, n)
    logits = torch.zeros((n, 7), device="cuda"); val = torch.zeros(n, device="cuda")
    age = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for t in (1, T):
        eng.policy.forward_rollout(b["frames"], b["dirs"], b["mission"][t + 3], t + 3, b["age"][t - 1], b["ep_len"][t - 1], age,
                                   val, logits=logits)
        img, d, mis = stacks[t]
        with torch.no_grad():
            lo, vo = o64({"direction": torch.from_numpy(d), "image": torch.from_numpy(img), "mission": torch.from_numpy(mis)})
        assert close(logits.cpu(), lo) and close(val.cpu(), vo), t
    eng.env.close()


def test_deterministic_evaluation_takes_the_argmax():
    n, T = 512, 12
    eng, o = make_engine(n, T)
    eng.collect(deterministic=True)
    b = eng.buf
    logits = torch.zeros((n, 7), device="cuda"); val = torch.zeros(n, device="cuda")
    age = torch.zeros(n, dtype=torch.uint8, device="cuda")
    for t in (1, 5, T - 1):
        eng.policy.forward_rollout(b["frames"], b["dirs"], b["mission"][t + 3], t + 3, b["age"][t - 1], b["ep_len"][t - 1], age,
                                   val, logits=logits)
        assert torch.equal(b["actions"][t].long(), logits.argmax(1))
        lsm = torch.log_softmax(logits, 1).gather(1, b["actions"][t].long()[:, None])[:, 0]
        assert torch.allclose(b["logp"][t], lsm, rtol=1e-5, atol=1e-6)
    eng.env.close()


def test_native_conv1_stage_matches_library_path():
    """mgrl_conv1_pool_forward / backward (first extractor stage of the update) against the torch expression: same
    logits / values, same gradients for every parameter."""
    n, T = 384, 12
    eng, _ = make_engine(n, T)
    eng.collect(); eng.compute_advantages()
    b, pol_ = eng.buf, eng.policy
    t = torch.arange(T, device="cuda").repeat_interleave(n)
    i = torch.arange(n, device="cuda").repeat(T)
    g = torch.Generator(device="cuda").manual_seed(0)
    sel = torch.randperm(T * n, device="cuda", generator=g)[:3000]          # 3000: a ragged last CTA / chunk
    t, i = t[sel], i[sel]
    image, onehot, mrow = ppo.gather_minibatch(b, t, i)
    la, va = pol_.evaluate(image, onehot, mrow)
    lb, vb = pol_.evaluate_samples(b["frames"], b["dirs"], b["mission"], b["age"], t, i)
    assert close(lb, la, 2e-5) and close(vb, va, 2e-5)
    wl, wv = torch.randn_like(la), torch.randn_like(va)
    params = pol_.parameters()
    ga = torch.autograd.grad((la * wl).sum() + (va * wv).sum(), params, allow_unused=True)
    gb = torch.autograd.grad((lb * wl).sum() + (vb * wv).sum(), params, allow_unused=True)
    for name, x, y in zip(pol_.params, ga, gb):
        assert (x is None) == (y is None), name
        if x is not None:
            assert close(y, x, 2e-4), name
    # and a whole iteration through the native stage
    stats = eng.iteration(1.0)
    assert stats["minibatches"] == 16 and eng.env.error_flags() == 0
    eng.env.close()


@pytest.mark.parametrize("n", [1, 33, 100])
def test_forward_kernel_ragged_batches(n):
    """batches that do not fill the last CTA (32 observations) of the policy kernel"""
    eng, o = make_engine(n, 8)
    o64 = po.double_copy(o)
    eng.collect()
    b = eng.buf
    stacks = oracle_stacks(b, 8, n)
    # the yardstick of the relative bar is the scale of the value head over a batch (a single observation whose value
    # happens to sit near zero has no scale of its own): max |V| of the 100-env rollout of the same policy
    for t in (0, 3, 8):
        img, d, mis = stacks[t]
        with torch.no_grad():
            lo, vo = o64({"direction": torch.from_numpy(d), "image": torch.from_numpy(img), "mission": torch.from_numpy(mis)})
        got = b["values"][t].cpu().double()
        assert float((got - vo).abs().max()) <= REL * max(float(vo.abs().max()), VALUE_SCALE), (n, t)
    eng.env.close()


def test_update_kernels_match_torch_ops():
    """The small hand-written kernels of the update, one by one, against the torch expression of the same thing:
    mgrl_colsum (bias gradients), mgrl_lut_grad (mission-table gradient), mgrl_patch2x2_forward / backward (conv2 patches)."""
    from minigrid_rl_b200 import _native as nat
    lib, s = nat.lib(), C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    g = torch.Generator(device="cuda").manual_seed(3)
    for rows, cols in [(1, 1), (1000, 7), (70001, 16), (262144, 32), (50000, 64), (33333, 128), (5, 100)]:
        x = torch.randn((rows, cols), device="cuda", generator=g)
        out = torch.full((cols,), 7.0, device="cuda")
        nat.check(lib.mgrl_colsum(p(x), rows, cols, p(out), s), "colsum")
        want = x.double().sum(0)
        assert float((out.double() - want).abs().max()) <= 2e-5 * max(1.0, float(x.abs().sum(0).max())), (rows, cols)
    for batch in (1, 37, 4096, 100000):
        d = torch.randn((batch, 128), device="cuda", generator=g)
        rows = torch.randint(0, 296, (batch,), device="cuda", generator=g)
        if batch == 100000:
            rows = rows % 5 + 90                        # a handful of hot rows, like one task's missions
        out = torch.full((296, 128), 3.0, device="cuda")
        nat.check(lib.mgrl_lut_grad(p(d), p(rows), batch, 296, p(out), s), "lut_grad")
        want = torch.zeros((296, 128), dtype=torch.float64, device="cuda").index_add_(0, rows, d.double())
        assert float((out.double() - want).abs().max()) <= 1e-5 * max(1.0, float(want.abs().max())), batch
    for batch in (1, 77, 5000):
        h = torch.randn((batch, 3, 3, 16), device="cuda", generator=g)
        patches = torch.empty((batch * 4, 64), device="cuda")
        nat.check(lib.mgrl_patch2x2_forward(p(h), batch, p(patches), s), "patch fwd")
        want = torch.stack([h[:, kh:kh + 2, kw:kw + 2, :] for kh in (0, 1) for kw in (0, 1)], dim=3).reshape(batch * 4, 64)
        assert torch.equal(patches, want)
        gp = torch.randn((batch * 4, 64), device="cuda", generator=g)
        dh = torch.empty((batch, 3, 3, 16), device="cuda")
        nat.check(lib.mgrl_patch2x2_backward(p(gp), batch, p(dh), s), "patch bwd")
        hh = h.clone().requires_grad_(True)
        w = torch.stack([hh[:, kh:kh + 2, kw:kw + 2, :] for kh in (0, 1) for kw in (0, 1)], dim=3).reshape(batch * 4, 64)
        (w * gp).sum().backward()
        assert torch.allclose(dh, hh.grad, rtol=1e-6, atol=1e-6)


def test_large_minibatch_update_path_matches_library_formulation(monkeypatch):
    """Minibatches of >= 16 384 samples take split-K weight gradients, mgrl_colsum and mgrl_lut_grad: same outputs and the
    same gradient for every parameter as the library formulation of the network on the same samples."""
    n, T = 1024, 32                                   # 32 768 samples
    eng, _ = make_engine(n, T)
    eng.collect()
    b, pol_ = eng.buf, eng.policy
    t = torch.arange(T, device="cuda").repeat_interleave(n)
    i = torch.arange(n, device="cuda").repeat(T)
    wl = torch.randn((n * T, 7), device="cuda") / (n * T)
    wv = torch.randn(n * T, device="cuda") / (n * T)
    params = pol_.parameters()

    def grads():
        lo, va = pol_.evaluate_samples(b["frames"], b["dirs"], b["mission"], b["age"], t, i)
        return lo.detach(), va.detach(), torch.autograd.grad((lo * wl).sum() + (va * wv).sum(), params, allow_unused=True)

    la, va, ga = grads()
    monkeypatch.setattr(pol, "NATIVE_MIN_ROWS", 1 << 60)
    lb, vb, gb = grads()
    assert close(la, lb, 2e-6) and close(va, vb, 2e-6)
    for name, x, y in zip(pol_.params, ga, gb):
        assert (x is None) == (y is None), name
        if x is not None:
            assert close(x, y, 2e-4), name
    eng.env.close()


def test_graph_replayed_update_equals_eager_update():
    """The optimizer step replayed from a CUDA graph (three eager steps, one capture, replays) leaves the same parameters
    as the same minibatches stepped eagerly."""
    out = []
    for graph in (False, True):
        eng, _ = make_engine(512, 32, n_epochs=1, cuda_graph=graph, native_update=False)
        eng.cfg.batch_size = 2048                      # 8 optimizer steps
        assert eng.updater.graphed == graph
        eng.collect(); eng.compute_advantages()
        gen = torch.Generator(device="cuda").manual_seed(5)
        assert eng.update(generator=gen) == 8
        out.append({k: v.detach().clone() for k, v in eng.policy.params.items()})
        eng.env.close()
    for k in out[0]:
        assert torch.allclose(out[0][k], out[1][k], rtol=2e-3, atol=2e-5), k


def _oracle_minibatch_grads(o, eng, t, i, stacks):
    """loss and gradients of the torch-CPU oracle network for the samples (t, i) of the engine's rollout"""
    b, cfg = eng.buf, eng.cfg
    tc, ic = t.cpu().numpy(), i.cpu().numpy()
    obs = {"image": torch.from_numpy(np.stack([stacks[a][0][e] for a, e in zip(tc, ic)])),
           "direction": torch.from_numpy(np.stack([stacks[a][1][e] for a, e in zip(tc, ic)])),
           "mission": torch.from_numpy(np.stack([stacks[a][2][e] for a, e in zip(tc, ic)]))}
    dt = next(o.parameters()).dtype
    args = [b[k][t, i].cpu() for k in ("actions", "values", "logp", "adv", "ret")]
    args = [args[0]] + [a.to(dt) for a in args[1:]]
    for p_ in o.parameters():
        p_.grad = None
    loss, _ = po.ppo_loss(o, obs, *args, cfg.clip_range, cfg.clip_range_vf, cfg.ent_coef, cfg.vf_coef)
    loss.backward()
    pre = "features_extractor.extractors."
    mods = {pre + "direction.direction_Linear_0": o.direction[0], pre + "image.image_Conv2d_0": o.image[0],
            pre + "image.image_Conv2d_3": o.image[3], pre + "image.image_Conv2d_5": o.image[5],
            "mlp_extractor.policy_net.0": o.pi[0], "mlp_extractor.policy_net.2": o.pi[2],
            "mlp_extractor.value_net.0": o.vf[0], "mlp_extractor.value_net.2": o.vf[2],
            "action_net": o.action_net, "value_net": o.value_net}
    g = {}
    for k, m in mods.items():
        g[k + ".weight"], g[k + ".bias"] = m.weight.grad, m.bias.grad
    g[pre + "mission.mission_Embedding_0.weight"] = o.embedding.weight.grad
    for nme in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
        g[pre + "mission.mission_GRU_1." + nme] = getattr(o.gru, nme).grad
    return loss.detach(), g


@pytest.mark.parametrize("native", [True, False], ids=["handwritten", "autograd"])
@pytest.mark.parametrize("tf32,bound,bound_gru", [(False, 1e-4, 1e-3), (True, 5e-3, 5e-3)], ids=["fp32", "tf32"])
def test_bench_sized_update_gradients_match_the_cpu_oracle(tf32, bound, bound_gru, native):
    """The update path that bench.py runs - minibatches of >= 16 384 samples (hand-written first stage, patch gather,
    table and bias gradients, split weight gradients) - against the torch-CPU oracle network evaluated in FLOAT64, which
    runs the GRU over every stacked mission: PPO loss and the gradient of EVERY parameter, max |got - want| <= bound *
    max |want| per tensor.  fp32: 1e-4, and 1e-3 for the mission branch (Embedding + GRU), whose gradient goes back
    through the 128-step recurrence in float32 (measured 2.7e-4).  update_tf32=True (the reference's own setting,
    ppo.py:29-32: TF32 matmuls): 5e-3, the error of 10-bit-mantissa products in the GEMMs (measured 1.8e-3)."""
    n, T = 512, 32                                      # 16 384 samples = one minibatch of the native path
    eng, o = make_engine(n, T, update_tf32=tf32, native_update=native)
    assert isinstance(eng.updater, ppo.NativeUpdater) == native
    assert n * T >= pol.NATIVE_MIN_ROWS
    eng.collect(); eng.bootstrap_truncated(); eng.compute_advantages()
    t = torch.arange(T, device="cuda").repeat_interleave(n)
    i = torch.arange(n, device="cuda").repeat(T)
    loss, got = eng.updater.gradients(eng.buf, t, i)
    stacks = oracle_stacks(eng.buf, T, n)
    lo, want = _oracle_minibatch_grads(po.double_copy(o), eng, t, i, stacks)
    assert abs(loss.item() - lo.item()) <= max(bound, 1e-4) * max(1.0, abs(lo.item()))
    worst = {}
    for k, w in want.items():
        err = float((got[k].cpu().double() - w).abs().max()) / max(float(w.abs().max()), 1e-12)
        group = "mission branch" if "mission" in k else "other"
        worst[group] = max(worst.get(group, 0.0), err)
        assert err <= (bound_gru if "mission" in k else bound), (k, err)
    print(f"update gradients vs float64 CPU oracle (tf32={tf32}, hand-written step={native}): worst relative error {worst}")
    eng.env.close()


def test_graph_replayed_bench_sized_steps_match_cpu_adam():
    """Five optimizer steps of the captured step (three eager, capture, replay) on one 16 384-sample minibatch against five
    steps of the torch-CPU oracle + torch.optim.Adam + clip_grad_norm_ on the same samples."""
    n, T = 512, 32
    eng, o = make_engine(n, T, native_update=False)
    assert eng.updater.graphed
    eng.collect(); eng.bootstrap_truncated(); eng.compute_advantages()
    t = torch.arange(T, device="cuda").repeat_interleave(n)
    i = torch.arange(n, device="cuda").repeat(T)
    stacks = oracle_stacks(eng.buf, T, n)
    cfg = eng.cfg
    opt = torch.optim.Adam(o.parameters(), lr=cfg.initial_learning_rate, eps=cfg.optim_eps)
    eng.updater.set_progress(1.0)
    for _ in range(5):
        eng.updater.minibatch_samples(eng.buf, t, i)
        _oracle_minibatch_grads(o, eng, t, i, stacks)
        torch.nn.utils.clip_grad_norm_(o.parameters(), cfg.max_grad_norm)
        opt.step()
    assert eng.updater._static is not None and eng.updater._static["graph"] is not None
    P = eng.policy.params
    pairs = {"mlp_extractor.policy_net.0.weight": o.pi[0].weight, "mlp_extractor.value_net.2.bias": o.vf[2].bias,
             "features_extractor.extractors.image.image_Conv2d_0.weight": o.image[0].weight,
             "features_extractor.extractors.image.image_Conv2d_5.bias": o.image[5].bias,
             "features_extractor.extractors.mission.mission_GRU_1.weight_hh_l0": o.gru.weight_hh_l0,
             "action_net.weight": o.action_net.weight}
    for k, w in pairs.items():
        assert torch.allclose(P[k].detach().cpu(), w.detach(), rtol=1e-3, atol=2e-5), k
    eng.env.close()


# ------------------------------------------------------------------------------ the hand-written optimizer step (K5)
def test_handwritten_mission_table_matches_the_float64_gru():
    """gru_fwd_kernel (W_hh in registers, two sequences per CTA) against torch's GRU evaluated in float64 on the 296 stacked
    mission sequences: the table the update differentiates."""
    eng, o = make_engine(64, 4)
    assert isinstance(eng.updater, ppo.NativeUpdater)
    got = eng.updater.mission_table().cpu().double()
    want = eng.policy.mission_lut_f64().cpu().double()          # float64 recurrence, rounded once
    err = float((got - want).abs().max()) / float(want.abs().max())
    print(f"mission table vs float64 GRU: {err:.2e}")
    assert err <= 1e-5
    eng.env.close()


@pytest.mark.parametrize("tf32,bound", [(False, 1e-5), (True, 3e-3)], ids=["fp32", "tf32"])
def test_handwritten_update_forward_matches_the_cpu_oracle(tf32, bound):
    """logits and values that the hand-written update computes for a minibatch (rows_gemm kernels + loss_kernel heads)
    against the float64 oracle network on the SB3 stacks: 1e-5 with the three-term split, 3e-3 with one TF32 pass."""
    n, T = 256, 16
    eng, o = make_engine(n, T, update_tf32=tf32)
    eng.collect(); eng.bootstrap_truncated(); eng.compute_advantages()
    t = torch.arange(T, device="cuda").repeat_interleave(n)
    i = torch.arange(n, device="cuda").repeat(T)
    up = eng.updater
    idx32 = (t * n + i).to(torch.int32)
    logits = torch.empty((n * T, 7), device="cuda"); values = torch.empty(n * T, device="cuda")
    sums = up.moments(eng.buf, idx32, n * T)
    up.gradients_native(up.view(eng.buf), idx32, sums[0], logits, values)
    stacks = oracle_stacks(eng.buf, T, n)
    obs = {"direction": torch.from_numpy(np.concatenate([s[1] for s in stacks[:T]])),
           "image": torch.from_numpy(np.concatenate([s[0] for s in stacks[:T]])),
           "mission": torch.from_numpy(np.concatenate([s[2] for s in stacks[:T]]))}
    with torch.no_grad():
        wl, wv = po.double_copy(o)(obs)
    el = float((logits.cpu().double() - wl).abs().max()) / float(wl.abs().max())
    ev = float((values.cpu().double() - wv).abs().max()) / float(wv.abs().max())
    print(f"hand-written update forward vs float64 oracle (tf32={tf32}): logits {el:.2e} values {ev:.2e}")
    assert el <= bound and ev <= bound
    # the minibatch advantage moments (double sums) against numpy
    a = eng.buf["adv"].cpu().numpy().astype(np.float64).reshape(-1)
    assert np.allclose(sums.cpu().numpy()[0], [a.sum(), (a * a).sum(), a.size], rtol=1e-12)
    eng.env.close()


def test_handwritten_steps_match_cpu_adam():
    """Five optimizer steps of mgrl_ppo_gradients + mgrl_ppo_apply on one 16 384-sample minibatch against five steps of the
    torch-CPU oracle + clip_grad_norm_ + torch.optim.Adam on the same samples."""
    n, T = 512, 32
    eng, o = make_engine(n, T)
    up = eng.updater
    assert isinstance(up, ppo.NativeUpdater)
    eng.collect(); eng.bootstrap_truncated(); eng.compute_advantages()
    t = torch.arange(T, device="cuda").repeat_interleave(n)
    i = torch.arange(n, device="cuda").repeat(T)
    idx32 = (t * n + i).to(torch.int32)
    stacks = oracle_stacks(eng.buf, T, n)
    cfg = eng.cfg
    opt = torch.optim.Adam(o.parameters(), lr=cfg.initial_learning_rate, eps=cfg.optim_eps)
    up.set_progress(1.0)
    view = up.view(eng.buf)
    sums = up.moments(eng.buf, idx32, n * T)
    for _ in range(5):
        up.minibatch_native(view, idx32, sums[0])
        _oracle_minibatch_grads(o, eng, t, i, stacks)
        torch.nn.utils.clip_grad_norm_(o.parameters(), cfg.max_grad_norm)
        opt.step()
    P = eng.policy.params
    pairs = {"mlp_extractor.policy_net.0.weight": o.pi[0].weight, "mlp_extractor.value_net.2.bias": o.vf[2].bias,
             "features_extractor.extractors.image.image_Conv2d_0.weight": o.image[0].weight,
             "features_extractor.extractors.image.image_Conv2d_3.weight": o.image[3].weight,
             "features_extractor.extractors.image.image_Conv2d_5.bias": o.image[5].bias,
             "features_extractor.extractors.direction.direction_Linear_0.weight": o.direction[0].weight,
             "features_extractor.extractors.mission.mission_Embedding_0.weight": o.embedding.weight,
             "features_extractor.extractors.mission.mission_GRU_1.weight_hh_l0": o.gru.weight_hh_l0,
             "features_extractor.extractors.mission.mission_GRU_1.weight_ih_l0": o.gru.weight_ih_l0,
             "action_net.weight": o.action_net.weight, "value_net.bias": o.value_net.bias}
    for k, w in pairs.items():
        assert torch.allclose(P[k].detach().cpu(), w.detach(), rtol=1e-3, atol=2e-5), k
    eng.env.close()


def test_handwritten_iteration_runs_ragged_minibatches_and_keeps_parameter_views():
    """A full iteration through RolloutEngine.update with a batch size that leaves a trailing partial minibatch; the
    parameters stay views of the flat buffer (state-dict names and shapes unchanged) and the rollout kernel picks the new
    weights up."""
    n, T = 200, 12                                      # 2400 samples, minibatches of 1000 -> 1000, 1000, 400
    eng, _ = make_engine(n, T, n_epochs=2)
    eng.cfg.batch_size = 1000
    before = eng.policy.flat.clone()
    stats = eng.iteration(0.5)
    assert stats["minibatches"] == 6
    assert eng.updater.steps == 6
    assert not torch.equal(before, eng.policy.flat)
    o = 0
    for k, v in eng.policy.params.items():
        assert v.shape == torch.Size(pol.SHAPES[k]) and torch.isfinite(v).all(), k
        assert v.data_ptr() == eng.policy.flat.data_ptr() + 4 * o, k
        o += v.numel()
    eng.collect()                                       # repacks the weights for the forward kernel
    assert torch.isfinite(eng.buf["values"]).all()
    assert eng.env.error_flags() == 0
    eng.env.close()


def test_tcgen05_mlp_gemms_match_the_mma_sync_path():
    """csrc/mgrl_linear_tc5.cu (tcgen05.mma kind::tf32, operands in shared-memory core matrices, accumulator in TMEM, epilogue
    through tcgen05.ld) against the mma.sync kernels for the same minibatch with one TF32 pass: first-layer activations, the
    feature gradient and every parameter gradient agree to TF32 rounding, and both sit within the TF32 bound of the float64
    oracle.  A ragged batch (not a multiple of the 128-row tile) exercises the zero-filled tail."""
    n, T = 250, 13                                       # 3250 samples = 25 tiles + 50 rows
    outs = []
    for tc5 in (False, True):
        eng, o = make_engine(n, T, update_tf32=True, update_tcgen05=tc5)
        assert bool(eng.updater.hyper.use_tcgen05) == tc5
        eng.collect(); eng.bootstrap_truncated(); eng.compute_advantages()
        t = torch.arange(T, device="cuda").repeat_interleave(n)
        i = torch.arange(n, device="cuda").repeat(T)
        loss, grads = eng.updater.gradients(eng.buf, t, i)
        a1 = eng.updater.debug_buffer("a1", (n * T, 128)).clone()
        df = eng.updater.debug_buffer("df", (n * T, 208)).clone()
        outs.append((loss, grads, a1, df, eng, o, t, i))
    (l0, g0, a0, d0, *_), (l1, g1, a1, d1, eng, o, t, i) = outs
    assert close(a1, a0, 2e-3) and close(d1, d0, 4e-3)
    assert abs(l0.item() - l1.item()) <= 2e-3 * max(1.0, abs(l0.item()))
    stacks = oracle_stacks(eng.buf, T, n)
    lo, want = _oracle_minibatch_grads(po.double_copy(o), eng, t, i, stacks)
    for k, w in want.items():
        err = float((g1[k].cpu().double() - w).abs().max()) / max(float(w.abs().max()), 1e-12)
        assert err <= 5e-3, (k, err)
    for e in (outs[0][4], outs[1][4]):
        e.env.close()


@pytest.mark.parametrize("batch", [3000, 13, 1, 40000])
def test_tcgen05_conv1_stage_matches_the_mma_sync_kernel_and_float64(batch, monkeypatch):
    """mgrl_conv1_tc5.cu (im2col-free tcgen05 implicit GEMM: pixels converted once into shared-memory planes, the four taps
    of the 2x2 kernel as row-shifted descriptors, accumulators in TMEM) against (a) a float64 evaluation of
    Conv2d(12,16,2) + ReLU + MaxPool2d(2) (policies.py:59, single.yaml:44-47) on the gathered stacks and (b) the mma.sync
    kernel it replaces.  Two-term split: <= 2e-6 of max |pooled| (bytes are exact in TF32, the weights carry hi + lo);
    one TF32 pass: <= 2e-3.  The arg-max byte must name a position whose value is the window's maximum (near-ties may
    resolve differently between kernels; exact ties resolve to the first position in both)."""
    import ctypes as C
    from minigrid_rl_b200 import _native as nat
    n, T = 384, 12
    eng, _ = make_engine(n, T)
    eng.collect()
    b, pol_ = eng.buf, eng.policy
    g = torch.Generator(device="cuda").manual_seed(1)
    reps = (batch + T * n - 1) // (T * n)
    sel = torch.cat([torch.randperm(T * n, device="cuda", generator=g) for _ in range(reps)])[:batch]
    t, i = (sel // n).to(torch.int32), (sel % n).to(torch.int32)
    age = b["age"][t.long(), i.long()].contiguous()
    P = pol_._P()
    c = "features_extractor.extractors.image.image_Conv2d_"
    w1 = (P[c + "0.weight"].detach() * 1.7).contiguous()         # (scaled: fresh weights leave few outputs positive)
    b1 = (P[c + "0.bias"].detach() + 0.05).contiguous()
    s = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731

    def run(mode):
        if mode is None:
            monkeypatch.delenv("MGRL_CONV1_TC5", raising=False)
        else:
            monkeypatch.setenv("MGRL_CONV1_TC5", str(mode))
        pooled = torch.full((batch, 9, 16), float("nan"), device="cuda")
        arg = torch.full((batch, 9, 16), 255, dtype=torch.uint8, device="cuda")
        nat.check(nat.lib().mgrl_conv1_pool_forward(p(b["frames"]), n, p(t), p(i), p(age), batch, p(w1), p(b1), p(pooled), p(arg), s),
                  "conv1_pool_forward")
        torch.cuda.synchronize()
        return pooled, arg

    # float64 reference on the gathered stacks
    k = torch.arange(4, device="cuda")
    fr = b["frames"][(t.long()[:, None] + k[None, :]), i.long()[:, None]][..., :147].reshape(batch, 4, 7, 7, 3)
    valid = ((3 - k)[None, :] <= age.long()[:, None])
    fr = fr * valid[:, :, None, None, None]
    img = fr.permute(0, 1, 4, 2, 3).reshape(batch, 12, 7, 7).double() / 255.0
    conv = torch.nn.functional.conv2d(img, w1.double(), b1.double())
    want, widx = torch.nn.functional.max_pool2d(torch.relu(conv), 2, return_indices=True)
    want = want.permute(0, 2, 3, 1).reshape(batch, 9, 16)
    pre = torch.nn.functional.unfold(conv.reshape(batch * 16, 1, 6, 6), 2, stride=2).reshape(batch, 16, 4, 9).permute(0, 3, 1, 2)  # [B,9,16,4]
    scale = float(want.abs().max())
    assert scale > 0.1
    ref_tc, _ = run(None)
    for mode, bound in ((1, 2e-6), (2, 2e-3)):
        got, arg = run(mode)
        assert not torch.isnan(got).any() and int(arg.max()) < 8
        assert float((got.double() - want).abs().max()) <= bound * scale, (mode, float((got.double() - want).abs().max()) / scale)
        pos, positive = (arg & 3).long(), (arg & 4) != 0
        picked = pre.gather(3, pos[..., None])[..., 0]
        assert float((pre.max(dim=3).values - picked).max()) <= 2 * bound * scale, mode          # the named position holds the maximum
        assert bool(((picked > bound * scale) <= positive).all()) and bool(((picked < -bound * scale) <= ~positive).all()), mode
        if mode == 1:
            assert float((got - ref_tc).abs().max()) <= 2e-6 * scale
    eng.env.close()
