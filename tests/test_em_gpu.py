"""GPU parity of the fused EM loop (vmm_em_forward / vmm_em_backward through the autograd
Function) against the CPU oracle on the same seeded inputs and weights.

Tolerances are north_star's: fp32 mode 1e-4 relative on outputs and gradients, bf16 mode
1e-2, with "relative" = max|a-b| / max|b| per tensor (helpers.rel_err).  A probe loss
  L = nem_loss + <h_T, Wh> + <gamma_T, Wg>
drives every backward path (through h, through gamma, through the loss) without the head.
"""
import json
import os

import pytest
import torch

from oracle import nem_oracle as O
from oracle.make_golden import seeded_models
from helpers import load_golden, rel_err, params_of

pytestmark = pytest.mark.gpu
OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")


def _probe(cfg):
    g = torch.Generator().manual_seed(77)
    wh = torch.randn(cfg["B"] * cfg["K"], cfg["H"], generator=g) * 0.01
    wg = torch.randn(cfg["B"], cfg["K"], cfg["M"], 1, generator=g)
    return wh, wg


def _oracle(cfg, nem, x, dtype, dec1_masks=None, dec1_pre=None):
    pn = params_of(nem, dtype=dtype)
    r = O.run_em(pn, x.to(dtype), cfg["K"], cfg["T"], cfg["sigma"], "bin_uniform", cfg["if_gamma_prior"], cfg["stop_gamma_grad"],
                 dec1_masks=dec1_masks, dec1_pre=dec1_pre)
    wh, wg = _probe(cfg)
    loss = r["nem_loss"] + (r["h"] * wh.to(dtype)).sum() + (r["gamma"] * wg.to(dtype)).sum()
    loss.backward()
    return r, pn


# dec1 is the ONE ReLU of the path (train_nem.py:181).  A unit whose pre-activation lies within the forward rounding
# error of zero has no defined derivative at the parity bar: two fp32-accurate evaluations (the reference on another
# BLAS, its own fp64 run, this library) may disagree on its sign, and ONE such unit moves dec1's bias / LayerNorm-beta /
# weight gradients by ~1e-2 of their max-norm at batch 8 (a column sum over only B*K*T = 240 terms), whatever the
# precision of the backward pass.  Measured (tools/relu_flips.py, profiles/r02_relu_flips.txt): at cfg1 with
# bf16-rounded features exactly one unit, |pre| = 4.0e-7, flips - in the fp32 preset too, whose gradients are then
# 1.2e-2 off the plain oracle and 1.8e-5 off the oracle that takes the library's decision for that unit.
KINK = 1e-5          # |pre| / max|pre| below which a ReLU decision is inside the reference's own fp32 rounding of |pre|
                     # (measured flips: 1e-7 .. 1e-6; the fp32-mode forward bar is 1e-4)


def _resolve_relu_kinks(cfg, nem, x, ref, pn, ref_pre, nem_c, max_units=1e-5):
    """Compare dec1's ReLU decisions of the CUDA run with the oracle's.  Every disagreement must sit on the kink
    (|pre_oracle| <= KINK * max|pre|) and be rare; if there is one, the oracle is re-evaluated with the library's
    decision for exactly those units.  Returns (ref, pn, number of such units)."""
    from helpers import cuda_dec1_pre
    pre_c = cuda_dec1_pre(nem_c, cfg["T"])
    masks, n_flip, n_all = [], 0, 0
    for a, b in zip(pre_c, ref_pre):
        flipped = (a > 0) != (b > 0)
        n_flip += int(flipped.sum())
        n_all += b.numel()
        if flipped.any():
            assert b[flipped].abs().max().item() <= KINK * b.abs().max().item(), "a ReLU decision differs AWAY from the kink"
        masks.append(((b > 0) ^ flipped).float())
    assert n_flip <= max(2, max_units * n_all), (n_flip, n_all)
    if n_flip:
        ref, pn = _oracle(cfg, nem.cpu(), x, torch.float32, dec1_masks=masks)
    return ref, pn, n_flip


def _cuda(cfg, nem, x, precision, xdtype):
    nem = nem.cuda().eval()          # parity is defined in eval mode (dropout off) with gradients enabled
    nem.precision = precision
    nem.keep_workspace = True        # stage-level inspection (helpers.cuda_dec1_pre)
    h, gamma, nem_loss, intra, inter = nem.run_em(x.cuda().to(xdtype), cfg["T"])
    wh, wg = _probe(cfg)
    loss = nem_loss + (h * wh.cuda()).sum() + (gamma * wg.cuda()).sum()
    loss.backward()
    torch.cuda.synchronize()
    return dict(h=h, gamma=gamma, nem_loss=nem_loss, intra=intra, inter=inter), nem


def _report(name, tag, cfg, res, nem_cuda, ref, ref_params, truth=None, truth_params=None):
    rows = {}
    for k in ("h", "gamma", "nem_loss", "intra", "inter"):
        rows[k] = rel_err(res[k].reshape(-1), ref[k].reshape(-1))
    named = dict(nem_cuda.named_parameters())
    for n, p in ref_params.items():
        if p.grad is None:
            assert named[n].grad is None, f"{n} must keep grad=None like the reference"
            continue
        rows["grad/" + n] = rel_err(named[n].grad, p.grad)
    ref_vs_truth = {}
    if truth is not None:
        for k in ("h", "gamma", "nem_loss"):
            ref_vs_truth[k] = rel_err(ref[k].reshape(-1), truth[k].reshape(-1))
        for n, p in truth_params.items():
            if p.grad is not None:
                ref_vs_truth["grad/" + n] = rel_err(ref_params[n].grad, p.grad)
    os.makedirs(OUT, exist_ok=True)
    with open(os.path.join(OUT, f"em_parity_{name}_{tag}.json"), "w") as f:
        json.dump(dict(cuda_vs_oracle=rows, oracle_fp32_vs_fp64=ref_vs_truth), f, indent=1)
    return rows


CASES = ["peaky_b6_k4_m12_t4", "prior0_b5_k2_m12_t3", "cfg1_b8_k3_m12_t10"]


def _tolerances(rows, base, ref_vs_truth=None, slack=4.0):
    """north_star's tolerance `base`, widened per tensor to `slack` x the reference's OWN fp32
    rounding noise (its distance from an fp64 run of the same math) where that noise is larger:
    two fp32-accurate implementations cannot agree better than that.  Only the deliberately
    ill-conditioned 'peaky' case (sigma = 0.02, gradients amplified by 1/sigma^2 = 2500) needs it."""
    out = {}
    for k in rows:
        noise = (ref_vs_truth or {}).get(k, 0.0)
        out[k] = max(base, slack * noise)
    return out


@pytest.mark.parametrize("name", CASES)
def test_run_em_fp32_mode(name):
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    ref, pn = _oracle(cfg, nem, x, torch.float32)
    truth, pt = _oracle(cfg, nem, x, torch.float64)
    res, nem_c = _cuda(cfg, nem, x, "fp32", torch.float32)
    rows = _report(name, "fp32", cfg, res, nem_c, ref, pn, truth, pt)
    noise = json.load(open(os.path.join(OUT, f"em_parity_{name}_fp32.json")))["oracle_fp32_vs_fp64"]
    tol = _tolerances(rows, 1e-4, noise)
    print(name, "fp32 worst", max(rows.items(), key=lambda kv: kv[1] / tol[kv[0]]))
    bad = {k: (v, tol[k]) for k, v in rows.items() if not v < tol[k]}
    assert not bad, f"fp32-mode parity (value, tolerance): {bad}"


@pytest.mark.parametrize("name", CASES)
def test_run_em_bf16_features_mode(name):
    """bf16 features: the oracle consumes the bf16-rounded features up-cast to fp32.  Flat 1e-2 in
    max-norm on every output and every gradient tensor (widened only for the ill-conditioned 'peaky' case, by the
    reference's own fp32-vs-fp64 noise); ReLU decisions on the kink are taken from the library (see KINK above)."""
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    xb = x.bfloat16()
    ref_pre = []
    ref, pn = _oracle(cfg, nem, xb.float(), torch.float32, dec1_pre=ref_pre)
    truth, pt = _oracle(cfg, nem, xb.float(), torch.float64)
    ref_plain, pn_plain = ref, pn
    res, nem_c = _cuda(cfg, nem, xb, "bf16", torch.bfloat16)
    ref, pn, n_kink = _resolve_relu_kinks(cfg, nem, xb.float(), ref, pn, ref_pre, nem_c)
    rows = _report(name, "bf16", cfg, res, nem_c, ref, pn, truth, pt)
    report = json.load(open(os.path.join(OUT, f"em_parity_{name}_bf16.json")))
    report["relu_units_on_the_kink"] = n_kink
    named = dict(nem_c.named_parameters())
    report["cuda_vs_plain_oracle"] = {"grad/" + n: rel_err(named[n].grad, p.grad) for n, p in pn_plain.items() if p.grad is not None}
    json.dump(report, open(os.path.join(OUT, f"em_parity_{name}_bf16.json"), "w"), indent=1)
    tol = _tolerances(rows, 1e-2, report["oracle_fp32_vs_fp64"] if name.startswith("peaky") else None, slack=20.0)
    print(name, "bf16 worst", max(rows.items(), key=lambda kv: kv[1] / tol[kv[0]]), "kink units", n_kink)
    bad = {k: (v, tol[k]) for k, v in rows.items() if not v < tol[k]}
    assert not bad, f"bf16-features parity (value, tolerance): {bad}"


def test_dec1_relu_kink_is_the_only_dec1_error():
    """The proof behind KINK: the fp32 preset (1e-4 bar) on the SAME bf16-rounded features as the test above.  Its
    forward is fp32-class, yet dec1's gradients are ~1e-2 off the plain oracle whenever a unit sits on the kink - and
    within 1e-4 of the oracle that takes the library's decision for those (asserted to be on the kink) units."""
    name = "cfg1_b8_k3_m12_t10"
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    xb = x.bfloat16()
    ref_pre = []
    ref, pn = _oracle(cfg, nem, xb.float(), torch.float32, dec1_pre=ref_pre)
    res, nem_c = _cuda(cfg, nem, xb, "fp32", torch.bfloat16)
    named = dict(nem_c.named_parameters())
    plain = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
    ref, pn, n_kink = _resolve_relu_kinks(cfg, nem, xb.float(), ref, pn, ref_pre, nem_c)
    fixed = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
    print("kink units", n_kink, "worst vs plain oracle", max(plain.values()), "vs kink-resolved oracle", max(fixed.values()))
    assert max(fixed.values()) < 1e-4, fixed
    for k in ("h", "gamma", "nem_loss"):
        assert rel_err(res[k].reshape(-1), ref[k].reshape(-1)) < 1e-4, k
    if n_kink == 0:
        assert max(plain.values()) < 1e-4, plain


@pytest.mark.parametrize("name", ["cfg1_b8_k3_m12_t10", "prior0_b5_k2_m12_t3"])
def test_run_em_bf16_margin(name):
    """The margin of the 'bf16' preset on the well-conditioned cases: with row-scaled fp16 operands in every
    data-gradient contraction of the BPTT chain every gradient tensor is within 7e-3 in max-norm (what is left is the
    bf16 rounding of the weight-gradient contractions, worst on matrices summed over few rows)."""
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    xb = x.bfloat16()
    ref_pre = []
    ref, pn = _oracle(cfg, nem, xb.float(), torch.float32, dec1_pre=ref_pre)
    res, nem_c = _cuda(cfg, nem, xb, "bf16", torch.bfloat16)
    ref, pn, _ = _resolve_relu_kinks(cfg, nem, xb.float(), ref, pn, ref_pre, nem_c)
    for k in ("h", "gamma", "nem_loss"):
        assert rel_err(res[k].reshape(-1), ref[k].reshape(-1)) < 1e-3, k
    named = dict(nem_c.named_parameters())
    bad = {}
    for n, p in pn.items():
        if p.grad is None:
            assert named[n].grad is None
            continue
        e = rel_err(named[n].grad, p.grad)
        if not e < 7e-3:
            bad[n] = e
    assert not bad, bad


@pytest.mark.parametrize("name", ["cfg1_b8_k3_m12_t10", "prior0_b5_k2_m12_t3"])
def test_run_em_bf16_b8_variant(name):
    """'bf16_b8' (round 1's headline: single bf16 planes in the data-gradient contractions as well) sits AT the 1e-2 bar
    on the well-conditioned cases (measured 6e-3 ... 1.03e-2 depending on the rounding of the forward pass: rnn_nem.bias_ih
    on prior0) and has no margin on 'peaky' - which is why it is no longer the default.  Pinned at 1.2e-2: an A/B variant,
    not a parity claim; the 'bf16' preset's claim is test_run_em_bf16_features_mode / test_run_em_bf16_margin."""
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    xb = x.bfloat16()
    ref_pre = []
    ref, pn = _oracle(cfg, nem, xb.float(), torch.float32, dec1_pre=ref_pre)
    res, nem_c = _cuda(cfg, nem, xb, "bf16_b8", torch.bfloat16)
    ref, pn, _ = _resolve_relu_kinks(cfg, nem, xb.float(), ref, pn, ref_pre, nem_c)
    named = dict(nem_c.named_parameters())
    bad = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None and not rel_err(named[n].grad, p.grad) < 1.2e-2}
    assert not bad, bad


def test_run_em_bf16_fast_forward():
    """Single-pass bf16 (inference grade): forward state within 5e-2 of the reference at cfg1."""
    name = "cfg1_b8_k3_m12_t10"
    z, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    nem = nem.cuda().eval()
    nem.precision = "bf16_fast"
    with torch.no_grad():
        h, gamma, nem_loss, _, _ = nem.run_em(x.cuda().bfloat16(), cfg["T"])
    assert rel_err(h, z["h"]) < 5e-2
    assert rel_err(gamma.reshape(z["gamma"].shape), z["gamma"]) < 1e-2
    assert abs(nem_loss.item() - float(z["nem_loss"])) < 1e-3 * abs(float(z["nem_loss"]))


@pytest.mark.parametrize("name", ["cfg1_b8_k3_m12_t10", "prior0_b5_k2_m12_t3", "peaky_b6_k4_m12_t4"])
def test_bf16_inference_single_pass(name):
    """Forward-only calls of the bf16-features preset run every contraction in ONE fp16 pass when sigma >= 0.07
    (fp16 pair otherwise - the 'peaky' case): descriptors h_T, responsibilities and the 4096-d head descriptor within
    north_star's 1e-2 of the oracle on the same bf16-rounded features, identical argmax class predictions."""
    from multi_view_sort_b200.functional import resolve_precision, fwd_modes, FWD_CONTRACTIONS, FWD_SINGLE
    _, cfg = load_golden(name)
    nem, head = seeded_models(cfg)
    x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    xb = x.bfloat16()
    pn, ph = params_of(nem), params_of(head)
    with torch.no_grad():
        r = O.run_em(pn, xb.float(), cfg["K"], cfg["T"], cfg["sigma"], "bin_uniform", cfg["if_gamma_prior"], cfg["stop_gamma_grad"])
        ref_logits = O.head_forward(ph, r["h"], r["gamma"], cfg["K"], cfg["if_sort"], cfg["if_pooling"])
    nem, head = nem.cuda().eval(), head.cuda().eval()
    nem.precision = head.precision = "bf16"
    all_single = fwd_modes(**{n: FWD_SINGLE for n in FWD_CONTRACTIONS})
    assert (resolve_precision(nem, training=False)[4] == all_single) == (cfg["sigma"] >= 0.07)
    with torch.no_grad():
        h, gamma, nem_loss, _, _ = nem.run_em(xb.cuda(), cfg["T"])
        logits = head(h, gamma)
    tol = 4e-2 if name.startswith("peaky") else 1e-2            # see test_train_step_fp32_mode_matches_reference_golden
    assert rel_err(h, r["h"]) < 1e-2
    assert rel_err(gamma.reshape(r["gamma"].shape), r["gamma"]) < 1e-2
    assert abs(nem_loss.item() - r["nem_loss"].item()) < 1e-2 * abs(r["nem_loss"].item())
    assert rel_err(logits, ref_logits) < tol
    assert (logits.argmax(1).cpu() == ref_logits.argmax(1)).all(), "argmax class predictions differ"


def test_run_em_matches_reference_golden():
    """CUDA fp32-mode forward against the frozen outputs of the UNMODIFIED reference."""
    name = "cfg1_b8_k3_m12_t10"
    z, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    nem = nem.cuda().eval()
    with torch.no_grad():
        h, gamma, nem_loss, intra, inter = nem.run_em(x.cuda(), cfg["T"])
    assert rel_err(h, z["h"]) < 1e-4
    assert rel_err(gamma.reshape(z["gamma"].shape), z["gamma"]) < 1e-4
    assert abs(nem_loss.item() - float(z["nem_loss"])) < 1e-4 * abs(float(z["nem_loss"]))
    assert abs(intra.item() - float(z["intra"])) < 1e-4 * abs(float(z["intra"]))
    assert abs(inter.item() - float(z["inter"])) < 1e-4 * abs(float(z["inter"]))


def test_inference_mode_equals_training_forward():
    """no_grad (2-slot ring workspace) and training forward produce identical outputs."""
    _, cfg = load_golden("peaky_b6_k4_m12_t4")
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    nem = nem.cuda().eval()
    a = nem.run_em(x.cuda(), cfg["T"])
    with torch.no_grad():
        b = nem.run_em(x.cuda(), cfg["T"])
    for u, v in zip(a, b):
        assert torch.equal(u, v)
    # and the no_grad path really uses the small workspace (2 iteration slots, not T)
    import ctypes as C
    from multi_view_sort_b200.functional import make_config, _bytes
    from multi_view_sort_b200._lib import lib
    small = _bytes(lib().vmm_em_workspace_bytes, make_config(nem, cfg["B"], cfg["T"], False, False))
    big = _bytes(lib().vmm_em_workspace_bytes, make_config(nem, cfg["B"], cfg["T"], False, True))
    assert small < big
    torch.cuda.reset_peak_memory_stats()
    base = torch.cuda.memory_allocated()
    with torch.no_grad():
        nem.run_em(x.cuda(), cfg["T"])
    assert torch.cuda.max_memory_allocated() - base < big


HEAD_CASES = ["cfg1_b8_k3_m12_t10", "peaky_b6_k4_m12_t4", "prior0_b5_k2_m12_t3"]


@pytest.mark.parametrize("name", HEAD_CASES)
def test_train_step_fp32_mode_matches_reference_golden(name):
    """The whole step of tools/Trainer.py:160-219 (EM loop -> head -> CE + nem_loss -> backward) against
    the frozen outputs and gradient fingerprints of the UNMODIFIED reference (tests/golden)."""
    from helpers import grad_summary
    from multi_view_sort_b200.functional import cross_entropy
    z, cfg = load_golden(name)
    nem, head = seeded_models(cfg)
    x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    nem, head = nem.cuda().eval(), head.cuda().eval()
    h, gamma, nem_loss, intra, inter = nem.run_em(x.cuda(), cfg["T"])
    logits = head(h, gamma)
    closs = cross_entropy(logits, y.cuda())
    loss = closs + nem_loss
    loss.backward()
    torch.cuda.synchronize()
    # 'peaky' is deliberately ill-conditioned (sigma = 0.02): the reference's own fp32 run is ~1e-2 away
    # from an fp64 run of the same math on gradients (see test_run_em_fp32_mode); allow 4x that noise.
    tol = 4e-2 if name.startswith("peaky") else 1e-4
    assert rel_err(logits, z["logits"]) < tol
    assert (logits.argmax(1).cpu().numpy() == z["logits"].argmax(1)).all(), "argmax class predictions differ"
    assert abs(closs.item() - float(z["closs"])) < tol * abs(float(z["closs"]))
    assert abs(loss.item() - float(z["loss"])) < tol * abs(float(z["loss"]))
    head.save_fea = 1
    with torch.no_grad():
        desc = head(h.detach(), gamma.detach())
    assert rel_err(desc, z["descriptor"]) < tol
    worst = {}
    for prefix, mod in (("nem.", nem), ("head.", head)):
        for n, p in mod.named_parameters():
            g = z["grad/" + prefix + n]
            if g.size == 1:                                   # NaN marker: the reference leaves grad = None
                assert p.grad is None, n
                continue
            mine = grad_summary(p.grad)
            # fingerprint = [norm, absmax, sum, 64 strided samples]; compare norm/absmax relatively and the
            # samples against the tensor's absmax
            assert abs(mine[0] - g[0]) < 10 * tol * g[0], (n, "norm", mine[0], g[0])
            e = max(abs(mine[1] - g[1]) / g[1], float(abs(mine[3:] - g[3:]).max() / g[1]))
            worst[prefix + n] = e
    # att.0.bias is ONE scalar: a signed sum of B*K terms, so its relative error carries the cancellation
    bad = {k: v for k, v in worst.items() if not v < (5 * tol if k.endswith("att.0.bias") else tol)}
    assert not bad, f"gradient fingerprints vs reference: {bad}"


def test_head_options_match_oracle():
    """if_sort 0/2 and max-pooling heads, and the save_feature descriptor, against the oracle."""
    _, cfg = load_golden("peaky_b6_k4_m12_t4")
    nem, _ = seeded_models(cfg)
    from multi_view_sort_b200 import Multi_View_Net
    x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    pn = params_of(nem, requires_grad=False)
    with torch.no_grad():
        r = O.run_em(pn, x, cfg["K"], 2, cfg["sigma"])
    for if_sort, if_pool in ((0, 0), (2, 0), (1, 1)):
        torch.manual_seed(3)
        head = Multi_View_Net(None, "vggm", nclasses=cfg["nclasses"], k=cfg["K"], rnn_hidden_size=cfg["H"],
                              if_sort=if_sort, if_pooling=if_pool)
        ph = params_of(head)
        hh = r["h"].clone().requires_grad_(True)
        out = O.head_forward(ph, hh, r["gamma"], cfg["K"], if_sort, if_pool, 0)
        w = torch.randn(out.shape, generator=torch.Generator().manual_seed(5))
        (out * w).sum().backward()
        head = head.cuda().eval()
        hc = r["h"].cuda().requires_grad_(True)
        outc = head(hc, r["gamma"].cuda())
        (outc * w.cuda()).sum().backward()
        assert rel_err(outc, out) < 1e-4, (if_sort, if_pool)
        assert rel_err(hc.grad, hh.grad) < 1e-4, (if_sort, if_pool)
        named = dict(head.named_parameters())
        for n, p in ph.items():
            assert rel_err(named[n].grad, p.grad) < 1e-4, (if_sort, if_pool, n)


def test_cfg3_80_views_matches_reference_golden():
    """Reduced BASELINE configs[2]: 80 views x 4096-d, K = 5 (NEM(view_num=80): 700 M parameters, LayerNorm over
    81,920 columns, enc2 K-dim 81,920) - full training step against the frozen reference outputs."""
    from helpers import grad_summary
    from multi_view_sort_b200.functional import cross_entropy
    z, cfg = load_golden("cfg3_b2_k5_m80_t3")
    nem, head = seeded_models(cfg)
    x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    nem, head = nem.cuda().eval(), head.cuda().eval()
    h, gamma, nem_loss, intra, inter = nem.run_em(x.cuda(), cfg["T"])
    logits = head(h, gamma)
    loss = cross_entropy(logits, y.cuda()) + nem_loss
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(h, z["h"]) < 1e-4
    assert rel_err(gamma.reshape(z["gamma"].shape), z["gamma"]) < 1e-4
    assert rel_err(logits, z["logits"]) < 1e-4
    assert (logits.argmax(1).cpu().numpy() == z["logits"].argmax(1)).all()
    assert abs(loss.item() - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    worst = {}
    for prefix, mod in (("nem.", nem), ("head.", head)):
        for n, p in mod.named_parameters():
            g = z["grad/" + prefix + n]
            if g.size == 1:
                assert p.grad is None, n
                continue
            mine = grad_summary(p.grad)
            worst[prefix + n] = max(abs(mine[1] - g[1]) / g[1], float(abs(mine[3:] - g[3:]).max() / g[1]))
    bad = {k: v for k, v in worst.items() if not v < (5e-4 if k.endswith("att.0.bias") else 1e-4)}
    assert not bad, bad


def test_cfg3_full_k_and_t_matches_reference_golden():
    """BASELINE configs[2] at its real K = 5 and T = 20 (80 views x 4096-d; one object - what the CPU reference finishes
    in a minute): full training step against the frozen reference outputs, fp32 preset at 1e-4; the bf16 preset's
    forward (descriptors, responsibilities, argmax) at 1e-2 on the same object with bf16-rounded features is checked
    against the oracle."""
    from helpers import grad_summary
    from multi_view_sort_b200.functional import cross_entropy
    z, cfg = load_golden("cfg3_b1_k5_m80_t20")
    nem, head = seeded_models(cfg)
    x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    nem, head = nem.cuda().eval(), head.cuda().eval()
    h, gamma, nem_loss, intra, inter = nem.run_em(x.cuda(), cfg["T"])
    logits = head(h, gamma)
    loss = cross_entropy(logits, y.cuda()) + nem_loss
    loss.backward()
    torch.cuda.synchronize()
    assert rel_err(h, z["h"]) < 1e-4
    assert rel_err(gamma.reshape(z["gamma"].shape), z["gamma"]) < 1e-4
    assert rel_err(logits, z["logits"]) < 1e-4
    assert (logits.argmax(1).cpu().numpy() == z["logits"].argmax(1)).all()
    assert abs(loss.item() - float(z["loss"])) < 1e-4 * abs(float(z["loss"]))
    worst = {}
    for prefix, mod in (("nem.", nem), ("head.", head)):
        for n, p in mod.named_parameters():
            g = z["grad/" + prefix + n]
            if g.size == 1:
                assert p.grad is None, n
                continue
            mine = grad_summary(p.grad)
            worst[prefix + n] = max(abs(mine[1] - g[1]) / g[1], float(abs(mine[3:] - g[3:]).max() / g[1]))
    bad = {k: v for k, v in worst.items() if not v < (5e-4 if k.endswith("att.0.bias") else 1e-4)}
    assert not bad, bad
    # bf16 preset, forward-only: single-pass contractions, 2-slot ring at M = 80
    nem.precision = head.precision = "bf16"
    with torch.no_grad():
        hb, gb, lb, _, _ = nem.run_em(x.cuda().bfloat16(), cfg["T"])
        r = O.run_em(params_of(nem, requires_grad=False, device="cpu"), x.bfloat16().float(), cfg["K"], cfg["T"], cfg["sigma"])
    assert rel_err(hb, r["h"]) < 1e-2 and rel_err(gb, r["gamma"]) < 1e-2
    assert abs(lb.item() - r["nem_loss"].item()) < 1e-2 * abs(r["nem_loss"].item())


def test_step_level_forward_matches_oracle_iterations():
    """The reference's own driving pattern (tools/Trainer.py:304-336): init_state, then T x
    nem(input, state) under no_grad - every intermediate state against the oracle's."""
    _, cfg = load_golden("prior0_b5_k2_m12_t3")
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    pn = params_of(nem, requires_grad=False)
    with torch.no_grad():
        ref = O.run_em(pn, x, cfg["K"], cfg["T"], cfg["sigma"], keep_all=True)["hist"]
    nem = nem.cuda().eval()
    xc = x.cuda()
    with torch.no_grad():
        init = nem.init_state(xc.shape, xc)
        state = (init[0], init[1], init[2])
        for t in range(cfg["T"]):
            state = nem(xc.unsqueeze(1), state)
            h, pred, gamma = state
            assert rel_err(h, ref[t][0]) < 1e-4, t
            assert rel_err(pred, ref[t][1]) < 1e-4, t
            assert rel_err(gamma, ref[t][2]) < 1e-4, t


@pytest.mark.parametrize("name", ["prior0_b5_k2_m12_t3", "peaky_b6_k4_m12_t4"])
def test_step_level_training_loop_matches_oracle(name):
    """The reference's TRAINING loop, unchanged (tools/Trainer.py:160-219): init_state, T x `state = nem(input, state)`
    with gradients enabled, compute_outer_loss in torch on the returned (pred, gamma) of EVERY iteration, backward
    through autograd.  NEM.forward is one autograd node per iteration (vmm_em_step_forward / _backward); gradients flow
    through h, pred and gamma between the nodes.  Every parameter gradient against the oracle's, and against the fused
    run_em path."""
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    wh, wg = _probe(cfg)
    wl = [0.3, 0.5, 1.0, 0.7, 0.2][:cfg["T"]]                  # a gradient into every iteration's loss, not only the last
    pn = params_of(nem)
    losses = []
    r = O.run_em(pn, x, cfg["K"], cfg["T"], cfg["sigma"], "bin_uniform", cfg["if_gamma_prior"], cfg["stop_gamma_grad"], losses=losses)
    ref_loss = sum(w * l[0] for w, l in zip(wl, losses)) + (r["h"] * wh).sum() + (r["gamma"] * wg).sum()
    ref_loss.backward()
    tol = 4e-2 if name.startswith("peaky") else 1e-4           # see test_train_step_fp32_mode_matches_reference_golden
    nem = nem.cuda().eval()
    xc = x.cuda()
    init = nem.init_state(xc.shape, xc)
    state, gamma_prior = (init[0], init[1], init[2]), init[3]
    total = 0.0
    for t in range(cfg["T"]):
        state = nem(xc.unsqueeze(1), state)
        assert state[0].requires_grad and state[1].requires_grad and state[2].requires_grad
        nl, _, _ = O.compute_outer_loss(state[1], state[2], xc.unsqueeze(1), cfg["stop_gamma_grad"], gamma_prior, cfg["if_gamma_prior"])
        assert abs(nl.item() - losses[t][0].item()) < max(tol, 1e-4) * abs(losses[t][0].item()), t
        total = total + wl[t] * nl
    h, pred, gamma = state
    (total + (h * wh.cuda()).sum() + (gamma * wg.cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert rel_err(h, r["h"]) < tol and rel_err(gamma, r["gamma"]) < tol and rel_err(pred, r["pred"]) < tol
    named = dict(nem.named_parameters())
    errs = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
    assert all(named[n].grad is None for n, p in pn.items() if p.grad is None)
    bad = {n: v for n, v in errs.items() if not v < tol}
    assert not bad, bad


def test_warm_up_and_total_loss_match_oracle():
    """-if_decoder_warm_up 1 (tools/Trainer.py:184-185: every iteration is fed gamma_prior) and -if_total_loss 1
    (:203: the mean of all T NEM losses, detached) on the fused path: outputs, every iteration's loss and every
    gradient against the oracle."""
    _, cfg = load_golden("prior0_b5_k2_m12_t3")
    cfg = dict(cfg, if_gamma_prior=1, stop_gamma_grad=0)
    nem, _ = seeded_models(cfg)
    nem.if_total_loss = 1
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    wh, wg = _probe(cfg)
    pn = params_of(nem)
    losses = []
    r = O.run_em(pn, x, cfg["K"], cfg["T"], cfg["sigma"], "bin_uniform", 1, 0, warm_up=True, losses=losses)
    (r["nem_loss"] + (r["h"] * wh).sum() + (r["gamma"] * wg).sum()).backward()
    nem = nem.cuda().eval()
    h, gamma, nem_loss, intra, inter = nem.run_em(x.cuda(), cfg["T"], warm_up=True)
    (nem_loss + (h * wh.cuda()).sum() + (gamma * wg.cuda()).sum()).backward()
    torch.cuda.synchronize()
    assert rel_err(h, r["h"]) < 1e-4 and rel_err(gamma, r["gamma"]) < 1e-4
    il = nem.iteration_losses.cpu()
    assert il.shape == (cfg["T"], 3)
    for t in range(cfg["T"]):
        for j in range(3):
            assert abs(il[t, j].item() - losses[t][j].item()) < 1e-4 * abs(losses[t][j].item()) + 1e-9, (t, j)
    assert abs(nem_loss.item() - losses[-1][0].item()) < 1e-4 * abs(losses[-1][0].item())
    named = dict(nem.named_parameters())
    errs = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
    bad = {n: v for n, v in errs.items() if not v < 1e-4}
    assert not bad, bad
    # and without warm-up the result differs (the flag is live)
    with torch.no_grad():
        h2, _, _, _, _ = nem.run_em(x.cuda(), cfg["T"])
    assert rel_err(h2, r["h"]) > 1e-3


def test_train_mode_dropout():
    """Train mode applies the reference's Dropout(0.5) (train_nem.py:172,175,181 and the VGG classifier) with a
    stateless hashed mask.  Bit-parity with torch's Philox stream is not a goal (hard part 6);
    instead the oracle applies the SAME masks (oracle.nem_oracle.dropout_mask restates the hash) and the whole
    training step - forward and every gradient - must match it: this pins that forward and backward kernels
    agree on the mask and on the 1/(1-p) scaling.  Also: eval mode is unaffected, a fixed seed is
    deterministic, another seed differs, and the kept fraction is ~1/2."""
    from multi_view_sort_b200.functional import cross_entropy
    _, cfg = load_golden("prior0_b5_k2_m12_t3")
    nem, head = seeded_models(cfg)
    x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    seed_n, seed_h = 1234, 777
    pn, ph = params_of(nem), params_of(head)
    ref = O.train_step(pn, ph, x, y, cfg["K"], cfg["T"], cfg["sigma"], cfg["if_gamma_prior"], cfg["stop_gamma_grad"],
                       cfg["if_sort"], cfg["if_pooling"], drop_nem=(0.5, seed_n), drop_head=(0.5, seed_h))
    ref["loss"].backward()
    nem, head = nem.cuda(), head.cuda()
    xc, yc = x.cuda(), y.cuda()

    def step():
        h, gamma, nem_loss, _, _ = nem.run_em(xc, cfg["T"])
        logits = head(h, gamma)
        return h, gamma, logits, cross_entropy(logits, yc) + nem_loss

    close = lambda a, b: abs(a - b) <= 2e-6 * abs(b)       # the CE reduction uses fp32 atomics: order jitter only
    nem.eval(); head.eval()
    with torch.no_grad():
        e1, e2 = step()[3].item(), step()[3].item()
    assert close(e1, e2)
    nem.train(); head.train()
    nem.dropout_seed, head.dropout_seed = seed_n, seed_h
    h, gamma, logits, loss = step()
    loss.backward()
    torch.cuda.synchronize()
    assert abs(loss.item() - e1) > 1e-3                      # dropout is really on
    assert rel_err(h, ref["h"]) < 1e-4 and rel_err(logits, ref["out"]) < 1e-4
    assert abs(loss.item() - ref["loss"].item()) < 1e-4 * abs(ref["loss"].item())
    for prefix, mod, pp in (("nem.", nem, pn), ("head.", head, ph)):
        named = dict(mod.named_parameters())
        for n, p in pp.items():
            if p.grad is None:
                continue
            tol = 5e-4 if n == "att.0.bias" else 1e-4
            assert rel_err(named[n].grad, p.grad) < tol, (prefix + n, rel_err(named[n].grad, p.grad))
    with torch.no_grad():
        t2 = step()[3].item()
        nem.dropout_seed = 99
        t3 = step()[3].item()
    assert close(loss.item(), t2) and not close(t2, t3)
    from csrc_probe import kept_fraction
    assert abs(kept_fraction(0.5, 1234, 1 << 16) - 0.5) < 0.01


@pytest.mark.parametrize("B,K,M,D,H,T,prior_mode", [(1, 2, 12, 256, 128, 2, 1), (3, 6, 12, 512, 256, 1, 2), (2, 1, 12, 256, 128, 2, 1),
                                                     (5, 5, 12, 264, 136, 2, 1), (7, 3, 4, 128, 64, 3, 0)])
def test_edge_shapes_match_oracle(B, K, M, D, H, T, prior_mode):
    """Ragged / minimal shapes: a single object, K = 1 (gamma == 1, train_nem.py:201-202), K = 6 of 12 views,
    dims that are not multiples of the tile sizes, one iteration, 4 views, all three outer-loss modes."""
    from multi_view_sort_b200 import NEM
    torch.manual_seed(10)
    nem = NEM(view_num=M, k=K, input_dim=D, rnn_hidden_size=H, iter_num=T, init_type="bin_uniform",
              gamma_prior_loss=prior_mode, e_sigma=0.15).eval()
    cfg = dict(B=B, K=K, M=M, D=D, H=H, T=T, sigma=0.15, if_gamma_prior=prior_mode, stop_gamma_grad=0, nclasses=40, scale=0.5)
    x, _ = O.make_inputs(B, M, D, 40, seed=3, scale=0.5)
    ref, pn = _oracle(cfg, nem, x, torch.float32)
    res, nem_c = _cuda(cfg, nem, x, "fp32", torch.float32)
    rows = {k: rel_err(res[k].reshape(-1), ref[k].reshape(-1)) for k in ("h", "gamma", "nem_loss", "intra", "inter")
            if ref[k].abs().max() > 0}
    named = dict(nem_c.named_parameters())
    for n, p in pn.items():
        if p.grad is not None and p.grad.abs().max() > 0:
            rows["grad/" + n] = rel_err(named[n].grad, p.grad)
    bad = {k: v for k, v in rows.items() if not v < 2e-4}
    assert not bad, bad


def test_rand_init_per_object_gamma():
    """init_type='rand' (the NEM class default, train_nem.py:205-208): per-object random responsibilities that
    double as the prior - passed explicitly so both sides see the same draw."""
    from multi_view_sort_b200 import NEM
    B, K, M, D, H, T = 4, 3, 12, 256, 128, 2
    torch.manual_seed(10)
    nem = NEM(view_num=M, k=K, input_dim=D, rnn_hidden_size=H, iter_num=T, init_type="rand", gamma_prior_loss=1, e_sigma=0.15).eval()
    x, _ = O.make_inputs(B, M, D, 40, seed=10, scale=0.5)
    g0 = torch.rand(B, K, M, 1, generator=torch.Generator().manual_seed(9))
    g0 = g0 / g0.sum(1).unsqueeze(1)
    pn = params_of(nem)
    r = O.run_em(pn, x, K, T, 0.15, "rand", 1, 0, gamma0=g0, gamma_prior=g0, keep_all=True)
    # Precondition of a tight gradient comparison: no dec1 pre-activation sits within the forward error of the
    # ReLU kink (with seed 4 one unit has |y| = 2.6e-7 and ANY two fp32 implementations may disagree on its
    # derivative, which moves every upstream gradient by ~3e-3 - see DESIGN.md 6).
    import torch.nn.functional as F
    with torch.no_grad():
        margin = min(F.layer_norm(F.linear(h_, pn["dec1.0.weight"], pn["dec1.0.bias"]), (4096,), pn["dec1.1.weight"],
                                  pn["dec1.1.bias"]).abs().min().item() for h_, _, _ in r["hist"])
    assert margin > 2e-5, margin
    wh, wg = torch.randn(B * K, H, generator=torch.Generator().manual_seed(1)) * 0.01, torch.randn(B, K, M, 1, generator=torch.Generator().manual_seed(2))
    (r["nem_loss"] + (r["h"] * wh).sum() + (r["gamma"] * wg).sum()).backward()
    nem = nem.cuda()
    h, gamma, nem_loss, _, _ = nem.run_em(x.cuda(), T, gamma0=g0.cuda(), gamma_prior=g0.cuda())
    (nem_loss + (h * wh.cuda()).sum() + (gamma * wg.cuda()).sum()).backward()
    assert rel_err(h, r["h"]) < 1e-4 and rel_err(gamma, r["gamma"]) < 1e-4
    assert abs(nem_loss.item() - r["nem_loss"].item()) < 1e-4 * abs(r["nem_loss"].item())
    named = dict(nem.named_parameters())
    errs = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
    bad = {n: v for n, v in errs.items() if not v < 2e-4}
    assert not bad, (bad, errs)
    # and the module's own init_state produces a valid per-object draw
    st = nem.init_state(x.shape, x)
    assert st[2].shape == (B, K, M, 1) and torch.allclose(st[2].sum(1), torch.ones(B, M, 1, device=st[2].device), atol=1e-6)


def test_full_size_properties_configs1():
    """BASELINE configs[1] at full size (4096 objects x 12 views x 4096-d, K=3, T=10; the oracle needs 511 GB of
    autograd state there) through size-independent properties: responsibilities sum to one over K, objects are
    independent (a permuted batch gives permuted outputs; a sub-batch gives the same rows), gradients of a split
    batch add up to the full-batch gradient, and the outer loss equals the mean of the per-half losses."""
    from multi_view_sort_b200 import NEM
    B, K, M, D, H, T = 4096, 3, 12, 4096, 1024, 10
    torch.manual_seed(10)
    nem = NEM(view_num=M, k=K, input_dim=D, rnn_hidden_size=H, iter_num=T, init_type="bin_uniform", gamma_prior_loss=1,
              e_sigma=0.1, precision="bf16").cuda().eval()
    g = torch.Generator(device="cuda").manual_seed(1)
    x = torch.relu(torch.randn(B, M, D, device="cuda", generator=g)).bfloat16()
    with torch.no_grad():
        h, gamma, loss, intra, inter = nem.run_em(x)
        assert torch.isfinite(h).all() and torch.isfinite(gamma).all() and torch.isfinite(loss)
        assert (gamma.sum(1) - 1).abs().max().item() < 1e-5
        assert gamma.min().item() > 0
        perm = torch.randperm(B, device="cuda", generator=g)
        h2, gamma2, loss2, _, _ = nem.run_em(x[perm])
        assert torch.equal(gamma2, gamma[perm])                         # bit-identical: no cross-object arithmetic
        assert torch.equal(h2.view(B, K, H), h.view(B, K, H)[perm])
        assert abs(loss2.item() - loss.item()) < 1e-5 * abs(loss.item())
        hs, gs, la, _, _ = nem.run_em(x[:1000])
        _, _, lb, _, _ = nem.run_em(x[1000:])
        assert torch.equal(gs, gamma[:1000]) and torch.equal(hs, h[:K * 1000])
        assert abs((1000 * la.item() + 3096 * lb.item()) / B - loss.item()) < 1e-5 * abs(loss.item())
        # ... and pinned to the ORACLE: objects are independent in the forward pass, so the first and the last 8 objects
        # of the 4096-batch (the CTA-pair tiles, the 10.4-wave schedule, the 2-slot inference ring at full size) must
        # reproduce the CPU oracle run on those 16 objects alone: north_star's 1e-2 for bf16 features, 1e-4 for fp32
        idx = torch.cat([torch.arange(0, 8), torch.arange(B - 8, B)])
        pn = params_of(nem, requires_grad=False, device="cpu")
        r = O.run_em(pn, x[idx.cuda()].float().cpu(), K, T, 0.1, "bin_uniform", 1, 0)
        assert rel_err(h.view(B, K, H)[idx.cuda()], r["h"].view(16, K, H)) < 1e-2
        assert rel_err(gamma[idx.cuda()], r["gamma"]) < 1e-2
        nem.precision = "fp32"
        hf, gf, _, _, _ = nem.run_em(x.float())
        assert rel_err(hf.view(B, K, H)[idx.cuda()], r["h"].view(16, K, H)) < 1e-4
        assert rel_err(gf[idx.cuda()], r["gamma"]) < 1e-4
        del hf, gf
        nem.precision = "bf16"
    # gradients: full batch == weighted sum of two halves (bf16 gradient planes: compare in relative L2)
    wh = torch.randn(B * K, H, device="cuda", generator=g) * 1e-3
    def grads(xs, whs, weight):
        for p in nem.parameters():
            p.grad = None
        hh, gg, ll, _, _ = nem.run_em(xs)
        ((ll + (hh * whs).sum() / xs.shape[0]) * weight).backward()
        return {n: p.grad.clone() for n, p in nem.named_parameters() if p.grad is not None}
    full = grads(x, wh, 1.0)
    a = grads(x[:2048], wh[:K * 2048], 0.5)
    b = grads(x[2048:], wh[K * 2048:], 0.5)
    for n in full:
        s = a[n] + b[n]
        rel = ((s - full[n]).norm() / full[n].norm().clamp_min(1e-30)).item()
        assert rel < 2e-3, (n, rel)
