"""CPU tier for the policy / PPO host logic: the product's differentiable network (mission GRU as a 296-row look-up
table) against the torch-CPU oracle that runs the GRU over the full stacked token sequence, the PPO minibatch loss,
the packed weight layout, and the data-parallel update over gloo with world_size 2."""
import os
import re
import socket

import numpy as np
import pytest
import torch

import minigrid_rl_b200 as mg
from minigrid_rl_b200 import policy as pol, ppo
from oracle import policy_oracle as po

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def random_stacked_obs(B, seed):
    rs = np.random.RandomState(seed)
    tt = mg.token_table()
    m, age = rs.randint(0, 74, B), rs.randint(0, 4, B)
    img = rs.randint(0, 11, (B, 12, 7, 7)).astype(np.uint8)
    d = np.zeros((B, 16), np.uint8)
    mis = np.zeros((B, 128), np.int64)
    for i in range(B):
        img[i, :(3 - age[i]) * 3] = 0
        for f in range(3 - age[i], 4):
            d[i, f * 4 + rs.randint(4)] = 1
            mis[i, f * 32:(f + 1) * 32] = tt[m[i]]
    return torch.from_numpy(img), torch.from_numpy(d), torch.from_numpy(mis), torch.from_numpy(m * 4 + age)


def make_pair(seed=3):
    o = po.OraclePolicy()
    po.init_reference(o, seed)
    p = pol.Policy("cpu", seed=1)
    p.load_oracle(o)
    return o, p


def test_parameter_count_and_names():
    assert pol.N_PARAMS == 110216
    o, p = make_pair()
    assert sum(x.numel() for x in o.parameters()) == pol.N_PARAMS
    assert "features_extractor.extractors.image.image_Conv2d_0.weight" in p.state_dict()
    assert "mlp_extractor.policy_net.0.weight" in p.state_dict() and "action_net.weight" in p.state_dict()


def test_sb3_named_checkpoint_round_trip(tmp_path):
    o, p = make_pair()
    path = str(tmp_path / "policy.pth")
    p.save(path)
    sd = torch.load(path)
    assert "pi_features_extractor.extractors.image.image_Conv2d_0.weight" in sd       # SB3's aliases of the shared extractor
    assert "vf_features_extractor.extractors.mission.mission_GRU_1.weight_hh_l0" in sd
    q = pol.Policy("cpu", seed=99)
    q.load(path)
    img, d, mis, mrow = random_stacked_obs(32, 2)
    a, b = p.evaluate(img, d, mrow), q.evaluate(img, d, mrow)
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1])


def oracle_from_sb3_state_dict(sd):
    """An OraclePolicy carrying the weights of an SB3 `policy.state_dict()` (the names CustomPPOPolicy has in the reference:
    policies.py:59 for the extractor, SB3's mlp_extractor / action_net / value_net)."""
    o = po.OraclePolicy()
    pre = "features_extractor.extractors."
    m = {pre + "direction.direction_Linear_0": o.direction[0], pre + "image.image_Conv2d_0": o.image[0],
         pre + "image.image_Conv2d_3": o.image[3], pre + "image.image_Conv2d_5": o.image[5],
         "mlp_extractor.policy_net.0": o.pi[0], "mlp_extractor.policy_net.2": o.pi[2],
         "mlp_extractor.value_net.0": o.vf[0], "mlp_extractor.value_net.2": o.vf[2],
         "action_net": o.action_net, "value_net": o.value_net}
    with torch.no_grad():
        for k, mod in m.items():
            mod.weight.copy_(sd[k + ".weight"]); mod.bias.copy_(sd[k + ".bias"])
        o.embedding.weight.copy_(sd[pre + "mission.mission_Embedding_0.weight"])
        for n in ("weight_ih_l0", "weight_hh_l0", "bias_ih_l0", "bias_hh_l0"):
            getattr(o.gru, n).copy_(sd[pre + "mission.mission_GRU_1." + n])
    return o


def test_sb3_zip_archive_round_trip(tmp_path):
    """ppo.py:128-132,145-150: a model archive written here is read the way SB3's load_from_zip_file reads one (zip members
    `data` JSON + `*.pth` through torch.load), its policy.pth loads into the reference network under SB3's names and gives
    the same logits / values; and an archive loads back into a fresh Policy."""
    import io
    import json
    import zipfile
    _, p = make_pair(seed=8)
    path = str(tmp_path / "model.zip")
    p.save_sb3_zip(path, data={"gamma": 0.81})
    with zipfile.ZipFile(path) as z:
        names = set(z.namelist())
        assert {"data", "policy.pth", "pytorch_variables.pth", "_stable_baselines3_version"} <= names
        assert json.loads(z.read("data").decode())["gamma"] == 0.81
        params = {n[:-4]: torch.load(io.BytesIO(z.read(n)), map_location="cpu") for n in names if n.endswith(".pth")}
    o = oracle_from_sb3_state_dict(params["policy"])
    img, d, mis, mrow = random_stacked_obs(48, 4)
    with torch.no_grad():
        lo, vo = o({"image": img, "direction": d, "mission": mis})
        lp, vp = p.evaluate(img, d, mrow)
    assert torch.allclose(lp, lo, rtol=0, atol=2e-6) and torch.allclose(vp, vo, rtol=0, atol=2e-6)
    q = pol.Policy("cpu", seed=123)
    q.load_sb3_zip(path)
    for k in p.params:
        assert torch.equal(p.params[k].detach(), q.params[k].detach()), k
    with pytest.raises(KeyError):
        bad = str(tmp_path / "bad.zip")
        with zipfile.ZipFile(bad, "w") as z:
            f = io.BytesIO(); torch.save({"x": torch.zeros(1)}, f); z.writestr("policy.pth", f.getvalue())
        q.load_sb3_zip(bad)


def test_lut_network_equals_full_gru_oracle():
    o, p = make_pair()
    img, d, mis, mrow = random_stacked_obs(256, 0)
    lo, vo = o({"direction": d, "image": img, "mission": mis})
    lp, vp = p.evaluate(img, d, mrow)
    assert torch.allclose(lo, lp, rtol=1e-5, atol=1e-6) and torch.allclose(vo, vp, rtol=1e-5, atol=1e-6)


def test_reference_init_statistics():
    p = pol.Policy("cpu", seed=7)
    for k, v in p.state_dict().items():
        if k.endswith("weight") and v.dim() == 2 and "GRU" not in k and "Embedding" not in k:
            assert torch.allclose(v.pow(2).sum(1), torch.ones(v.shape[0]), atol=1e-5), k   # policies.py:253-255
        if k.endswith("bias") and "GRU" not in k:
            assert not v.any()
    w = p.params["features_extractor.extractors.image.image_Conv2d_0.weight"].detach().reshape(16, -1)
    assert torch.allclose(w @ w.t(), 2 * torch.eye(16), atol=1e-4)                         # orthogonal, gain sqrt(2)


def test_ppo_loss_and_gradients_match_oracle():
    o, p = make_pair()
    cfg = ppo.PPOConfig()
    B = 128
    img, d, mis, mrow = random_stacked_obs(B, 1)
    g = torch.Generator().manual_seed(0)
    actions = torch.randint(0, 7, (B,), generator=g).to(torch.uint8)
    old_v, old_lp = torch.randn(B, generator=g), -torch.rand(B, generator=g) * 2
    adv, ret = torch.randn(B, generator=g), torch.randn(B, generator=g)
    lo, _ = po.ppo_loss(o, {"direction": d, "image": img, "mission": mis}, actions, old_v, old_lp, adv, ret, cfg.clip_range,
                        cfg.clip_range_vf, cfg.ent_coef, cfg.vf_coef)
    lp, _ = ppo.ppo_minibatch_loss(p, cfg, img, d, mrow, actions, old_v, old_lp, adv, ret)
    assert torch.allclose(lo, lp, rtol=1e-5, atol=1e-6)
    lo.backward(); lp.backward()
    assert torch.allclose(o.gru.weight_hh_l0.grad, p.params[pol._PREFIX + "mission.mission_GRU_1.weight_hh_l0"].grad,
                          rtol=1e-4, atol=1e-6)
    assert torch.allclose(o.image[0].weight.grad, p.params[pol._PREFIX + "image.image_Conv2d_0.weight"].grad,
                          rtol=1e-4, atol=1e-6)
    assert torch.allclose(o.embedding.weight.grad, p.params[pol._PREFIX + "mission.mission_Embedding_0.weight"].grad,
                          rtol=1e-4, atol=1e-6)


def test_packed_weight_layout_matches_header_and_kernel_source():
    hdr = open(os.path.join(ROOT, "include", "mgrl.h")).read()
    assert int(re.search(r"#define MGRL_POLICY_WEIGHTS (\d+)", hdr).group(1)) == pol.N_WEIGHTS
    assert int(re.search(r"#define MGRL_POLICY_FRAGMENTS (\d+)", hdr).group(1)) == pol.N_FRAGMENTS
    src = open(os.path.join(ROOT, "minigrid-rl_b200", "csrc", "mgrl_policy_layout.cuh")).read()
    assert "N_WEIGHTS == MGRL_POLICY_WEIGHTS" in src
    assert "F_END * 4 == MGRL_POLICY_FRAGMENTS" in open(os.path.join(ROOT, "minigrid-rl_b200", "csrc", "mgrl_policy_tc.cu")).read()
    assert pol.N_WEIGHTS % 4 == 0                   # the fragment section behind the weights is 16-byte aligned
    off = 0
    for name, (o, n) in pol.WEIGHT_LAYOUT.items():
        assert o == off and o % 4 == 0, name        # 16-byte aligned rows for the float4 weight loads
        off += n


def test_linear_schedule_and_gather_shapes():
    f = ppo.linear_schedule(3e-4, 3e-6)             # ppo.py:35-40
    assert f(1.0) == 3e-4 and f(0.5) == 1.5e-4 and f(0.0) == 3e-6
    T, N = 6, 5
    rs = np.random.RandomState(0)
    buf = {"frames": torch.from_numpy(rs.randint(0, 11, (T + 4, N, 148)).astype(np.uint8)),
           "dirs": torch.from_numpy(rs.randint(0, 4, (T + 4, N)).astype(np.uint8)),
           "mission": torch.from_numpy(rs.randint(0, 74, (T + 4, N)).astype(np.uint8)),
           "age": torch.from_numpy(rs.randint(0, 4, (T + 1, N)).astype(np.uint8))}
    t, i = torch.tensor([0, 3, 5]), torch.tensor([4, 0, 2])
    image, onehot, mrow = ppo.gather_minibatch(buf, t, i)
    assert image.shape == (3, 12, 7, 7) and onehot.shape == (3, 16)
    for s in range(3):
        a = int(buf["age"][t[s], i[s]])
        for f in range(4):
            fr = buf["frames"][t[s] + f, i[s], :147].view(7, 7, 3).permute(2, 0, 1)
            want = fr if (3 - f) <= a else torch.zeros_like(fr)
            assert torch.equal(image[s, f * 3:(f + 1) * 3], want)
            want_d = torch.zeros(4, dtype=torch.uint8)
            if (3 - f) <= a:
                want_d[int(buf["dirs"][t[s] + f, i[s]])] = 1
            assert torch.equal(onehot[s, f * 4:(f + 1) * 4], want_d)
        assert int(mrow[s]) == int(buf["mission"][t[s] + 3, i[s]]) * 4 + a


# ----------------------------------------------------------------------------- data parallel over gloo, world_size 2
def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _dp_worker(rank, world, port, out):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(1)
    cfg = ppo.PPOConfig()
    p = pol.Policy("cpu", seed=100 + rank)            # different initial weights: the updater must broadcast rank 0's
    up = ppo.Updater(p, cfg, dist)
    img, d, mis, mrow = random_stacked_obs(64, 5)
    g = torch.Generator().manual_seed(1)
    actions = torch.randint(0, 7, (64,), generator=g).to(torch.uint8)
    old_v, old_lp, adv, ret = (torch.randn(64, generator=g), -torch.rand(64, generator=g), torch.randn(64, generator=g),
                               torch.randn(64, generator=g))
    sl = slice(rank * 32, (rank + 1) * 32)            # each rank holds half of the global minibatch
    for _ in range(2):
        up.minibatch(img[sl], d[sl], mrow[sl], actions[sl], old_v[sl], old_lp[sl], adv[sl], ret[sl])
    torch.save({k: v.detach() for k, v in p.params.items()}, out + f".{rank}")
    if rank == 0:
        torch.save({"n_all_reduces": up.n_all_reduces}, out + ".meta")
    dist.destroy_process_group()


def test_data_parallel_update_equals_single_process(tmp_path):
    import torch.multiprocessing as mp
    out = str(tmp_path / "dp")
    mp.spawn(_dp_worker, args=(2, _free_port(), out), nprocs=2, join=True)
    a, b = torch.load(out + ".0"), torch.load(out + ".1")
    for k in a:
        assert torch.equal(a[k], b[k]), k             # ranks stay bit-identical
    # single process on the concatenated minibatch
    cfg = ppo.PPOConfig()
    p = pol.Policy("cpu", seed=100)
    up = ppo.Updater(p, cfg, None)
    img, d, mis, mrow = random_stacked_obs(64, 5)
    g = torch.Generator().manual_seed(1)
    actions = torch.randint(0, 7, (64,), generator=g).to(torch.uint8)
    old_v, old_lp, adv, ret = (torch.randn(64, generator=g), -torch.rand(64, generator=g), torch.randn(64, generator=g),
                               torch.randn(64, generator=g))
    for _ in range(2):
        up.minibatch(img, d, mrow, actions, old_v, old_lp, adv, ret)
    for k in a:
        assert torch.allclose(a[k], p.params[k].detach(), rtol=2e-4, atol=2e-6), k
    assert torch.load(out + ".meta")["n_all_reduces"] == 4   # per optimizer step: advantage moments + the flat gradient
