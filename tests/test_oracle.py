"""CPU tests of the oracle: (1) bit-equality with the imported reference where
/root/reference exists, (2) agreement with the committed goldens everywhere,
(3) structural properties of the restated math."""
import numpy as np
import pytest
import torch

from oracle import nem_oracle as O
from oracle import ref_import
from oracle.make_golden import CASES, seeded_models
from helpers import load_golden, rel_err, grad_summary, params_of

SMALL = ["peaky_b6_k4_m12_t4", "prior0_b5_k2_m12_t3"]


def _oracle_step(name, dtype=torch.float32):
    z, cfg = load_golden(name)
    nem, head = seeded_models(cfg)
    pn, ph = params_of(nem, dtype=dtype), params_of(head, dtype=dtype)
    x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    r = O.train_step(pn, ph, x.to(dtype), y, cfg["K"], cfg["T"], cfg["sigma"], cfg["if_gamma_prior"],
                     cfg["stop_gamma_grad"], cfg["if_sort"], cfg["if_pooling"])
    r["loss"].backward()
    return z, cfg, r, pn, ph


@pytest.mark.parametrize("name", SMALL + ["cfg1_b8_k3_m12_t10"])
def test_oracle_matches_golden(name):
    """The restatement reproduces the frozen outputs of the reference (bit-exact when the
    torch build matches the one that wrote the fixture; 1e-6 otherwise)."""
    z, cfg, r, pn, ph = _oracle_step(name)
    same_torch = str(z["torch_version"]) == torch.__version__
    tol = 0.0 if same_torch else 1e-6
    assert rel_err(r["h"], z["h"]) <= tol
    assert rel_err(r["gamma"].reshape(z["gamma"].shape), z["gamma"]) <= tol
    assert rel_err(r["out"], z["logits"]) <= tol
    for key in ("nem_loss", "intra", "inter", "closs", "loss"):
        assert abs(r[key].item() - float(z[key])) <= tol * abs(float(z[key])) + (0 if same_torch else 1e-7)
    assert (r["out"].argmax(1).numpy() == z["logits"].argmax(1)).all()
    for n, p in pn.items():
        g = z["grad/nem." + n]
        if np.isnan(g).all():
            assert p.grad is None, n        # after_rnn.* never receives a gradient (train_nem.py:179)
        else:
            assert rel_err(grad_summary(p.grad), g) <= max(tol, 1e-12), n
    for n, p in ph.items():
        assert rel_err(grad_summary(p.grad), z["grad/head." + n]) <= max(tol, 1e-12), n


@pytest.mark.skipif(not ref_import.available(), reason="/root/reference not mounted")
@pytest.mark.parametrize("name", SMALL)
def test_oracle_bit_equals_reference(name):
    """Run the UNMODIFIED reference modules and the restatement on the same weights and
    inputs: every output and every parameter gradient must be bit-identical on CPU."""
    cfg = CASES[name]
    nem, head = seeded_models(cfg)
    rnem, rhead = ref_import.build_reference_models(
        k=cfg["K"], view_num=cfg["M"], input_dim=cfg["D"], rnn_hidden_size=cfg["H"], iter_num=cfg["T"],
        nclasses=cfg["nclasses"], e_sigma=cfg["sigma"], gamma_prior_loss=cfg["if_gamma_prior"],
        stop_gamma_grad=cfg["stop_gamma_grad"], if_sort=cfg["if_sort"], if_pooling=cfg["if_pooling"])
    rnem.load_state_dict(nem.state_dict()); rhead.load_state_dict(head.state_dict())
    rnem.eval(); rhead.eval()
    x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    r = ref_import.reference_train_step(rnem, rhead, x, y, cfg["T"])
    r["loss"].backward()
    pn, ph = params_of(nem), params_of(head)
    o = O.train_step(pn, ph, x, y, cfg["K"], cfg["T"], cfg["sigma"], cfg["if_gamma_prior"],
                     cfg["stop_gamma_grad"], cfg["if_sort"], cfg["if_pooling"])
    o["loss"].backward()
    for key in ("h", "pred", "gamma", "out", "nem_loss", "intra", "inter", "closs", "loss"):
        assert torch.equal(r[key], o[key]), key
    for n, p in rnem.named_parameters():
        if p.grad is None:
            assert pn[n].grad is None
        else:
            assert torch.equal(p.grad, pn[n].grad), n
    for n, p in rhead.named_parameters():
        assert torch.equal(p.grad, ph[n].grad), n


def test_reference_init_tables():
    """bin_uniform one-hot positions and prior normalisation (tools/NEM_utilities.py:29-36)."""
    for k, m in [(3, 12), (4, 12), (5, 80), (2, 12), (6, 12)]:
        g, p = O.bin_uniform(2, k, m)
        stride = m // k
        for i in range(k):
            assert g[0, i, :, 0].argmax().item() == i * stride and g[0, i].sum().item() == 1
            assert abs(p[0, i].sum().item() - 1) < 1e-6
            # Reference quirk, reproduced on purpose: np.roll(grid, m//2 - center) moves the mode
            # AWAY from `center` (to 2*(m//2) - center), tools/NEM_utilities.py:10-13.
            mode = (2 * (m // 2) - i * stride) % m
            assert abs(p[0, i, mode, 0].item() - p[0, i].max().item()) < 1e-7
    if ref_import.available():
        ref = ref_import.load()
        g2, p2 = ref.bin_uniform(2, 3, 12)
        g, p = O.bin_uniform(2, 3, 12)
        assert torch.equal(g, g2) and torch.equal(p, p2)


def test_gamma_sums_to_one_and_object_permutation():
    """gamma sums to 1 over K; permuting the objects of a batch permutes the outputs."""
    name = "peaky_b6_k4_m12_t4"
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    pn = params_of(nem, requires_grad=False)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    with torch.no_grad():
        r = O.run_em(pn, x, cfg["K"], 2, cfg["sigma"])
        perm = torch.randperm(cfg["B"], generator=torch.Generator().manual_seed(0))
        r2 = O.run_em(pn, x[perm], cfg["K"], 2, cfg["sigma"])
    assert torch.allclose(r["gamma"].sum(1), torch.ones_like(r["gamma"].sum(1)), atol=1e-6)
    assert rel_err(r2["gamma"], r["gamma"][perm]) < 1e-5
    H = cfg["H"]
    assert rel_err(r2["h"].view(cfg["B"], cfg["K"], H), r["h"].view(cfg["B"], cfg["K"], H)[perm]) < 1e-5
