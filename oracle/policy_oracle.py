"""torch-CPU fp32 restatement of the reference policy and PPO loss (TEST INFRASTRUCTURE, NOT THE PRODUCT).

Follows /root/reference/src/policies.py:21-120 (CustomExtractor built from
hydra_configs/single.yaml:38-57 with n_frames_stack = 4), policies.py:227-257 (CustomPPOPolicy,
init_weights) and, for what the reference takes from Stable-Baselines3 without vendoring it
([UPSTREAM]): preprocess_obs (image / 255), MlpExtractor (pi/vf 208->64->64 Tanh:
SB3's default net_arch, the YAML one is popped and unused, policies.py:235), action_net / value_net,
CategoricalDistribution, and PPO.train's loss (clipped surrogate, clipped value loss, entropy bonus,
per-minibatch advantage normalisation).

Only tests/ may import this module.  It computes the GRU over the full 128-token stacked mission
(no look-up table), which is what the product's mission LUT is checked against.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F


class OraclePolicy(nn.Module):
    def __init__(self, n_frames_stack: int = 4):
        super().__init__()
        k = n_frames_stack
        self.direction = nn.Sequential(nn.Linear(4 * k, 16))                                  # single.yaml:41-43
        self.image = nn.Sequential(nn.Conv2d(3 * k, 16, (2, 2)), nn.ReLU(), nn.MaxPool2d(2),   # single.yaml:44-53
                                   nn.Conv2d(16, 32, (2, 2)), nn.ReLU(),
                                   nn.Conv2d(32, 64, (2, 2)), nn.ReLU(), nn.Flatten())
        self.embedding = nn.Embedding(32, 32)                                                  # single.yaml:54-57
        self.gru = nn.GRU(32, 128, 1, True, True)
        self.pi = nn.Sequential(nn.Linear(208, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh())  # [UPSTREAM] MlpExtractor
        self.vf = nn.Sequential(nn.Linear(208, 64), nn.Tanh(), nn.Linear(64, 64), nn.Tanh())
        self.action_net = nn.Linear(64, 7)
        self.value_net = nn.Linear(64, 1)

    def features(self, obs):
        """obs: direction (B,16) u8, image (B,12,7,7) u8, mission (B,128) i64; concat order direction, image, mission
        (policies.py:83-102 iterates the ModuleDict, whose keys follow the sorted observation space)."""
        dt = self.direction[0].weight.dtype                     # float32; float64 for the `.double()` copy (tests)
        d = self.direction(obs["direction"].to(dt))
        im = self.image(obs["image"].to(dt) / 255.0)            # [UPSTREAM] preprocess_obs, normalize_images=true
        _, h = self.gru(self.embedding(obs["mission"].long()))  # policies.py:84-91
        return torch.cat([d, im, h[-1]], dim=1)

    def forward(self, obs):
        f = self.features(obs)
        return self.action_net(self.pi(f)), self.value_net(self.vf(f)).squeeze(-1)


def double_copy(policy: OraclePolicy) -> OraclePolicy:
    """The same network evaluated in float64: the yardstick of the 1e-5 bar (the float32 evaluation above carries
    its own rounding error of the same size as the kernels')."""
    import copy
    return copy.deepcopy(policy).double()


def init_reference(policy: OraclePolicy, seed: int) -> None:
    """CustomPPOPolicy.init_weights (policies.py:246-257) applied the way SB3 applies it: Conv2d orthogonal
    (gain sqrt(2) inside the features extractor), every Linear N(0,1) rows scaled to unit L2 norm, biases 0;
    Embedding / GRU keep PyTorch defaults."""
    torch.manual_seed(seed)
    for m in policy.modules():
        if isinstance(m, (nn.Embedding, nn.GRU)):
            m.reset_parameters()
        if isinstance(m, nn.Conv2d):
            nn.init.orthogonal_(m.weight, gain=float(np.sqrt(2)))
            m.bias.data.fill_(0.0)
        if isinstance(m, nn.Linear):
            m.weight.data.normal_(0, 1)
            m.weight.data *= 1 / torch.sqrt(m.weight.data.pow(2).sum(1, keepdim=True))
            m.bias.data.fill_(0.0)


def evaluate_actions(policy: OraclePolicy, obs, actions):
    logits, values = policy(obs)
    logp_all = F.log_softmax(logits, dim=1)            # [UPSTREAM] Categorical(logits=...)
    logp = logp_all.gather(1, actions.long().view(-1, 1)).squeeze(1)
    entropy = -(logp_all.exp() * logp_all).sum(1)
    return values, logp, entropy


def ppo_loss(policy: OraclePolicy, obs, actions, old_values, old_logp, advantages, returns, clip_range, clip_range_vf,
             ent_coef, vf_coef, normalize_advantage=True):
    """[UPSTREAM] PPO.train, one minibatch."""
    values, logp, entropy = evaluate_actions(policy, obs, actions)
    adv = advantages
    if normalize_advantage and adv.numel() > 1:
        adv = (adv - adv.mean()) / (adv.std() + 1e-8)
    ratio = torch.exp(logp - old_logp)
    pl = -torch.min(adv * ratio, adv * torch.clamp(ratio, 1 - clip_range, 1 + clip_range)).mean()
    vp = values if clip_range_vf is None else old_values + torch.clamp(values - old_values, -clip_range_vf, clip_range_vf)
    vl = F.mse_loss(returns, vp)
    el = -entropy.mean()
    return pl + ent_coef * el + vf_coef * vl, (pl, vl, el)


def sample_inverse_cdf(logits: np.ndarray, u: np.ndarray):
    """Categorical sampling as the product defines it: fp32 softmax, inverse CDF on one uniform per row.
    Returns (action, log-prob of the action)."""
    lg = torch.from_numpy(np.ascontiguousarray(logits, np.float32))
    lsm = F.log_softmax(lg, dim=1)
    p = lsm.exp().numpy()
    c = np.cumsum(p.astype(np.float32), axis=1, dtype=np.float32)
    a = (u[:, None] >= c).sum(1).clip(0, logits.shape[1] - 1)
    return a.astype(np.uint8), lsm.numpy()[np.arange(len(a)), a]
