"""CPU restatement of the reference's VMM / Neural-EM hot path (torch, fp32 or fp64).

TEST INFRASTRUCTURE (see oracle/__init__.py) - the checker, never the product.

Every function cites the reference lines it restates (paths relative to
/root/reference).  Parameters are passed as a plain dict keyed by the
reference's own ``state_dict`` names (train_nem.py:171-185, 310-326), so the
same weights drive the reference, this oracle and the CUDA path.

The restatement is written from the math, using the same
torch primitives in the same order as the reference so that on CPU/fp32 it is
bit-identical to it (asserted in tests/test_oracle_vs_reference.py).  Run with
``dtype=torch.float64`` it is the "truth" used to size tolerances.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import numpy as np
import torch
import torch.nn.functional as F

Params = Dict[str, torch.Tensor]

ENC1_W = 1024   # train_nem.py:171  Linear(input_dim, 1024)
ENC2_W = 4096   # train_nem.py:175  Linear(1024*m, 4096)
DEC1_W = 4096   # train_nem.py:181  Linear(H, 4096)
DEC3_IN = 1024  # train_nem.py:183,185  Linear(4096, 1024*m) -> Linear(1024, input_dim)


# ----------------------------------------------------------------------------
# init (train_nem.py:188-217, tools/NEM_utilities.py:9-36)
# ----------------------------------------------------------------------------
def _norm_pdf(x: np.ndarray, scale: float) -> np.ndarray:
    # scipy.stats.norm.pdf(x, loc=0, scale=s) = exp(-(x/s)^2/2) / (sqrt(2 pi) s)
    y = x / scale
    return np.exp(-y * y / 2.0) / (math.sqrt(2.0 * math.pi) * scale)


def compute_norm(center: int, m: int) -> torch.Tensor:
    """tools/NEM_utilities.py:9-17 - Gaussian (sigma .15) over linspace(-.5,.5,m)
    rolled so that its mode sits on view `center`, normalised to sum 1. (m,1) fp32."""
    init_center = int(m / 2)
    num_list = np.linspace(-0.5, 0.5, m, endpoint=True)
    rolled = np.roll(num_list, init_center - center)
    dist = _norm_pdf(rolled, 0.15)
    return torch.Tensor(dist / np.sum(dist)).unsqueeze(1)


def bin_uniform(B: int, k: int, m: int) -> Tuple[torch.Tensor, torch.Tensor]:
    """tools/NEM_utilities.py:29-36."""
    gamma = torch.zeros([B, k, m, 1])
    gamma_prior = torch.zeros([B, k, m, 1])
    stride = int(m / k)
    for i in range(k):
        gamma[:, i, i * stride, :] = 1
        gamma_prior[:, i] = compute_norm(i * stride, m)
    return gamma, gamma_prior


def init_state(B: int, k: int, m: int, input_dim: int, hidden: int, init_type: str = "bin_uniform",
               generator: Optional[torch.Generator] = None):
    """train_nem.py:188-217. Returns (h, pred, gamma, gamma_prior), fp32 CPU."""
    h = torch.zeros(B * k, hidden)
    pred = torch.ones([B, k, m, input_dim]) * 0.0          # pred_init = 0.0 (train_nem.py:156)
    gamma_prior, gamma = torch.ones([B, k, m, 1]) / m, torch.ones([B, k, m, 1])
    if k > 1:
        if init_type == "rand":
            gamma = torch.rand([B, k, m, 1], generator=generator)
            gamma = gamma / gamma.sum(1).unsqueeze(1)
            gamma_prior = gamma
        elif init_type == "bin_uniform":
            gamma, gamma_prior = bin_uniform(B, k, m)
        else:
            raise ValueError(f"oracle does not restate init_type={init_type!r}")
    return h, pred, gamma, gamma_prior


# ----------------------------------------------------------------------------
# train-mode dropout with the CUDA library's stateless mask (NOT torch's Philox stream: bit-parity with
# nn.Dropout's RNG is not a goal, SURVEY.md 7 hard part 6).  Host restatement of
# multi_view_sort_b200/csrc/stages.cuh::dropout_hash so the GPU test can check train-mode forward AND
# backward against the oracle with identical masks.
# ----------------------------------------------------------------------------
def dropout_mask(p_drop: float, seed: int, shape, dtype=torch.float32) -> torch.Tensor:
    """multi_view_sort_b200/csrc/stages.cuh::dropout_scale: one 32-bit hash per group of four consecutive
    elements, one byte per element, kept iff byte >= round(p * 256)."""
    n = int(np.prod(shape))
    assert n % 4 == 0
    group = np.arange(n // 4, dtype=np.uint64)
    seed = seed & ((1 << 64) - 1)
    with np.errstate(over="ignore"):
        x = (group & np.uint64(0xFFFFFFFF)).astype(np.uint32) ^ np.uint32(seed & 0xFFFFFFFF)
        y = (group >> np.uint64(32)).astype(np.uint32) ^ np.uint32(seed >> 32)
        x = x ^ (y * np.uint32(0x9E3779B9))
        x = x ^ (x >> np.uint32(16)); x = x * np.uint32(0x7feb352d)
        x = x ^ (x >> np.uint32(15)); x = x * np.uint32(0x846ca68b)
        x = x ^ (x >> np.uint32(16))
    bytes4 = np.stack([(x >> np.uint32(8 * q)) & np.uint32(255) for q in range(4)], axis=1).reshape(-1)
    keep = bytes4 >= np.uint32(int(p_drop * 256.0 + 0.5))
    return torch.from_numpy(keep.astype(np.float64) / (1.0 - p_drop)).to(dtype).reshape(shape)


def _stage_seed(seed: int, stage: int, it: int) -> int:
    return (seed + (stage << 56) + (it << 40)) & ((1 << 64) - 1)


# ----------------------------------------------------------------------------
# one EM iteration (train_nem.py:219-294)
# ----------------------------------------------------------------------------
def run_inner_rnn(p: Params, masked: torch.Tensor, h_old: torch.Tensor, b: int, k: int, m: int, drop=None,
                  dec1_mask: Optional[torch.Tensor] = None, dec1_pre: Optional[list] = None):
    """train_nem.py:219-245.  drop=None: eval mode (Dropout = identity); drop=(p, seed, iteration): the
    Dropout(0.5) layers of enc1 / enc2 / dec1 (train_nem.py:172,175,181) with the library's masks.

    ReLU-kink diagnostics (tests/test_em_gpu.py::test_bf16_dec1_relu_flips): `dec1_pre` collects dec1's
    pre-activation; `dec1_mask` (0/1, same shape) replaces ReLU by `pre * mask`, i.e. evaluates the network with
    ANOTHER implementation's decision about which dec1 units are active - the two differ only for units whose
    pre-activation is within that implementation's forward rounding error of zero."""
    x = masked.view(b * k * m, -1)
    a = F.elu(F.layer_norm(F.linear(x, p["enc1.0.weight"], p["enc1.0.bias"]), (ENC1_W,),
                           p["enc1.1.weight"], p["enc1.1.bias"]))
    if drop:
        a = a * dropout_mask(drop[0], _stage_seed(drop[1], 1, drop[2]), a.shape, a.dtype)
    a = a.view(b * k, -1)
    e = F.elu(F.layer_norm(F.linear(a, p["enc2.0.weight"], p["enc2.0.bias"]), (ENC2_W,),
                           p["enc2.1.weight"], p["enc2.1.bias"]))
    if drop:
        e = e * dropout_mask(drop[0], _stage_seed(drop[1], 2, drop[2]), e.shape, e.dtype)
    h = torch._VF.gru_cell(e, h_old, p["rnn_nem.weight_ih"], p["rnn_nem.weight_hh"],
                           p["rnn_nem.bias_ih"], p["rnn_nem.bias_hh"])
    if dec1_mask is None and dec1_pre is None:
        u = F.relu(F.layer_norm(F.linear(h, p["dec1.0.weight"], p["dec1.0.bias"]), (DEC1_W,),
                                p["dec1.1.weight"], p["dec1.1.bias"]))
    else:
        pre = F.layer_norm(F.linear(h, p["dec1.0.weight"], p["dec1.0.bias"]), (DEC1_W,),
                           p["dec1.1.weight"], p["dec1.1.bias"])
        if dec1_pre is not None:
            dec1_pre.append(pre.detach())
        u = F.relu(pre) if dec1_mask is None else pre * dec1_mask.to(pre.dtype)
    if drop:
        u = u * dropout_mask(drop[0], _stage_seed(drop[1], 3, drop[2]), u.shape, u.dtype)
    v = F.layer_norm(F.linear(u, p["dec2.0.weight"], p["dec2.0.bias"]), (DEC3_IN * m,),
                     p["dec2.1.weight"], p["dec2.1.bias"])
    z = v.view(b * k * m, -1)
    pred = F.linear(z, p["dec3.weight"], p["dec3.bias"]).view(b, k, m, -1)
    return pred, h


def compute_em_probabilities(pred: torch.Tensor, t_data: torch.Tensor, sigma: float, epsilon: float = 1e-10):
    """train_nem.py:247-263 ('fc' features, bn_init_sigma == 0)."""
    probs = (1 / math.sqrt(2 * np.pi * sigma ** 2)) * torch.exp(-(t_data - pred) ** 2 / (2 * sigma ** 2))
    probs = probs.sum(-1).unsqueeze(-1)
    if epsilon > 0:
        probs = probs + epsilon
    return probs


def e_step(pred: torch.Tensor, t_data: torch.Tensor, sigma: float):
    """train_nem.py:265-270."""
    probs = compute_em_probabilities(pred, t_data, sigma)
    return probs / probs.sum(1).unsqueeze(1)


def em_iteration(p: Params, x: torch.Tensor, state, sigma: float, drop=None, dec1_mask=None, dec1_pre=None):
    """train_nem.py:272-294.  x: (B,1,M,D); state = (h, pred, gamma)."""
    h_old, pred_old, gamma_old = state
    b, k, m = gamma_old.shape[0], gamma_old.shape[1], gamma_old.shape[2]
    delta = x - pred_old
    masked = delta * gamma_old
    pred, h = run_inner_rnn(p, masked, h_old, b, k, m, drop, dec1_mask, dec1_pre)
    gamma = e_step(pred, x, sigma)
    return h, pred, gamma


# ----------------------------------------------------------------------------
# outer loss (tools/NEM_utilities.py:76-128)
# ----------------------------------------------------------------------------
def kl_loss_gamma(prior: torch.Tensor, pred: torch.Tensor):
    """tools/NEM_utilities.py:85-89: KLDiv(batchmean) of log_softmax over the VIEW axis."""
    logp = F.log_softmax(pred.view(-1, pred.shape[2]), dim=1)
    return F.kl_div(logp, prior.view(-1, prior.shape[2]), reduction="batchmean")


def gamma_entropy_loss(gamma: torch.Tensor):
    """tools/NEM_utilities.py:91-95."""
    entropy = torch.sum(gamma * torch.log(gamma + 1e-6)) / gamma.shape[0] / gamma.shape[2] * -1.0
    sigma = torch.sum(torch.sqrt(torch.sum(((gamma - gamma.mean(2).unsqueeze(2)) ** 2), 2))) / gamma.shape[0] / gamma.shape[1]
    return entropy - sigma


def compute_outer_loss(mu, gamma, target, stop_gamma_grad: int, gamma_prior, if_gamma_prior: int):
    """tools/NEM_utilities.py:98-128 with prior == 0 (compute_prior, :66-72)."""
    w = [2, 0.5] if if_gamma_prior == 0 else [2, 1]
    intra_p = (mu - target) ** 2 / 1.0 + math.log(1.0)                      # gaussian_MSE_loss(mu, 1, t)
    if if_gamma_prior == 0:
        inter_p = math.log(1.0) + (1.0 + (mu - 0.0) ** 2) / 2.0 - 0.5       # kl_loss_gaussian(0, mu, 1, 1)
    elif if_gamma_prior == 1:
        inter_p = kl_loss_gamma(gamma_prior, gamma)
    else:
        inter_p = gamma_entropy_loss(gamma) + kl_loss_gamma(gamma_prior, gamma)
    B, M, D = mu.shape[0], mu.shape[2], mu.shape[3]
    g = gamma.detach() if stop_gamma_grad else gamma
    intra = torch.sum(intra_p * g) / (B * M * D)
    if if_gamma_prior == 0:
        inter = torch.sum(inter_p * (1.0 - g)) / (B * M * D)
    else:
        inter = inter_p
    total = intra * w[0] + w[1] * inter
    return total, intra, inter


# ----------------------------------------------------------------------------
# head (train_nem.py:328-380)
# ----------------------------------------------------------------------------
def head_forward(p: Params, h: torch.Tensor, gamma: torch.Tensor, k: int, if_sort: int = 1,
                 if_pooling: int = 0, save_feature: int = 0, drop=None):
    """train_nem.py:328-380 in eval mode.  h: (B*k, H).  Returns logits (B,C), or the
    4096-d descriptor (output of the 2nd Linear, pre-ReLU) when save_feature=1."""
    scores = torch.sigmoid(F.linear(h, p["att.0.weight"], p["att.0.bias"]))      # (B*k,1)
    s = scores.view(-1, k)
    masked = h.view(-1, k, h.shape[1]) * s.unsqueeze(2)
    if if_pooling == 1:
        feat, _ = torch.max(masked, 1)
    elif if_sort == 1:
        _, idx = torch.sort(s, dim=1, descending=True)
        feat = torch.gather(masked, 1, idx.unsqueeze(2).expand_as(masked)).reshape(masked.shape[0], -1)
    elif if_sort == 2:
        sig = torch.sqrt(torch.sum((gamma - gamma.mean(2).unsqueeze(2)) ** 2, 2)).squeeze(-1)
        _, idx = torch.sort(sig, dim=1, descending=True)
        feat = torch.gather(masked, 1, idx.unsqueeze(2).expand_as(masked)).reshape(masked.shape[0], -1)
    else:
        feat = masked.reshape(masked.shape[0], -1)
    t1 = F.relu(F.linear(feat, p["net.0.weight"], p["net.0.bias"]))
    if drop:                                          # VGG classifier Dropout after each hidden ReLU
        t1 = t1 * dropout_mask(drop[0], drop[1] + 1, t1.shape, t1.dtype)
    t3 = F.linear(t1, p["net.3.weight"], p["net.3.bias"])
    if save_feature:
        return t3
    t3r = F.relu(t3)
    if drop:
        t3r = t3r * dropout_mask(drop[0], drop[1] + 2, t3r.shape, t3r.dtype)
    return F.linear(t3r, p["net.6.weight"], p["net.6.bias"])


# ----------------------------------------------------------------------------
# the whole step (tools/Trainer.py:160-203)
# ----------------------------------------------------------------------------
def run_em(p: Params, x: torch.Tensor, k: int, iters: int, sigma: float = 0.1,
           init_type: str = "bin_uniform", if_gamma_prior: int = 1, stop_gamma_grad: int = 0,
           gamma0: Optional[torch.Tensor] = None, gamma_prior: Optional[torch.Tensor] = None,
           keep_all: bool = False, drop=None, dec1_masks=None, dec1_pre: Optional[list] = None,
           warm_up: bool = False, losses: Optional[list] = None):
    """T iterations + the last iteration's outer loss.  x: (B,M,D).
    Returns dict(h, gamma, pred, nem_loss, intra, inter[, hist]).  dec1_masks / dec1_pre: per-iteration
    ReLU-kink diagnostics, see run_inner_rnn.  warm_up: tools/Trainer.py:184-185 (`if_decoder_warm_up == 1 and
    epoch < 10`): every iteration is fed gamma_prior instead of the running responsibilities.  losses: collects
    compute_outer_loss of EVERY iteration (Trainer.py:191-195) as (total, intra, inter) tuples."""
    B, M, D = x.shape
    hidden = p["rnn_nem.weight_hh"].shape[1]
    h, pred, g0, gp = init_state(B, k, M, D, hidden, init_type)
    if gamma0 is not None:
        g0 = gamma0
    if gamma_prior is not None:
        gp = gamma_prior
    dt, dev = x.dtype, x.device          # dev != cpu: bench.py's eager-GPU comparator; the state is built on the host and
    state = (h.to(dev, dt), pred.to(dev, dt), g0.to(dev, dt))   # copied every step, as the reference does (train_nem.py:217)
    gp = gp.to(dev, dt)
    xin = x.unsqueeze(1)
    hist = []
    for it in range(iters):
        st_in = (state[0], state[1], gp) if warm_up else state
        state = em_iteration(p, xin, st_in, sigma, (drop[0], drop[1], it) if drop else None,
                             dec1_masks[it] if dec1_masks is not None else None, dec1_pre)
        if losses is not None:
            losses.append(compute_outer_loss(state[1], state[2], xin, stop_gamma_grad, gp, if_gamma_prior))
        if keep_all:
            hist.append(state)
    h, pred, gamma = state
    nem_loss, intra, inter = compute_outer_loss(pred, gamma, xin, stop_gamma_grad, gp, if_gamma_prior)
    out = dict(h=h, gamma=gamma, pred=pred, nem_loss=nem_loss, intra=intra, inter=inter)
    if keep_all:
        out["hist"] = hist
    return out


def train_step(p_nem: Params, p_head: Params, x: torch.Tensor, labels: Optional[torch.Tensor], k: int,
               iters: int, sigma: float = 0.1, if_gamma_prior: int = 1, stop_gamma_grad: int = 0,
               if_sort: int = 1, if_pooling: int = 0, init_type: str = "bin_uniform", drop_nem=None, drop_head=None):
    """Trainer.py:160-203: EM loop, head, loss = CE + nem_loss_T."""
    r = run_em(p_nem, x, k, iters, sigma, init_type, if_gamma_prior, stop_gamma_grad, drop=drop_nem)
    out = head_forward(p_head, r["h"], r["gamma"], k, if_sort, if_pooling, 0, drop_head)
    r["out"] = out
    if labels is not None:
        r["closs"] = F.cross_entropy(out, labels)
        r["loss"] = r["closs"] + r["nem_loss"]
    return r


# ----------------------------------------------------------------------------
# seeded parameters & inputs (SURVEY.md §8d seed protocol; no reference needed)
# ----------------------------------------------------------------------------
def make_inputs(B: int, M: int, D: int, nclasses: int = 40, seed: int = 10, scale: float = 1.0):
    """x = relu(randn(B,M,D)) * scale, labels = randint(0,nclasses) from Generator(seed)."""
    g = torch.Generator().manual_seed(seed)
    x = torch.relu(torch.randn(B, M, D, generator=g)) * scale
    y = torch.randint(0, nclasses, (B,), generator=g)
    return x, y


def cast_params(p: Params, dtype) -> Params:
    return {k: v.detach().to(dtype).clone().requires_grad_(v.is_floating_point()) for k, v in p.items()}
