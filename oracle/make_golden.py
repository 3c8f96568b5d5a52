"""Freeze outputs of the UNMODIFIED reference (imported from /root/reference) into
tests/golden/*.npz.  TEST INFRASTRUCTURE; run in the dev container only:

    python -m oracle.make_golden

Weights are NOT stored (129 M+ parameters): they are re-created from the seed recipe
`seeded_models()` below (torch CPU RNG, seed 10, module construction order NEM -> head),
which needs only the product package's parameter containers, and are loaded into the
reference modules through their own `load_state_dict` - which also pins state_dict
compatibility.  Each fixture records the torch version that produced it.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import nem_oracle as O  # noqa: E402

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# name -> config.  cfg1 is BASELINE.json configs[0]; the others are reduced parity cases.
CASES = {
    # BASELINE configs[0]: 12 views x 4096-d, K=3, 10 iters, batch 8
    "cfg1_b8_k3_m12_t10": dict(B=8, K=3, M=12, D=4096, H=1024, T=10, sigma=0.1, scale=1.0, nclasses=40,
                               if_gamma_prior=1, stop_gamma_grad=0, if_sort=1, if_pooling=0),
    # responsibilities far from uniform (narrow sigma, few dims) so that E-step bugs cannot hide
    "peaky_b6_k4_m12_t4": dict(B=6, K=4, M=12, D=256, H=512, T=4, sigma=0.02, scale=0.25, nclasses=10,
                               if_gamma_prior=1, stop_gamma_grad=0, if_sort=1, if_pooling=0),
    # NEM-class defaults for the loss (gamma_prior_loss=0), stop_gamma_grad, max-pool head
    "prior0_b5_k2_m12_t3": dict(B=5, K=2, M=12, D=512, H=256, T=3, sigma=0.1, scale=0.5, nclasses=40,
                                if_gamma_prior=0, stop_gamma_grad=1, if_sort=1, if_pooling=1),
    # reduced BASELINE configs[2]: 80 views, K=5 (NEM(view_num=80): 700 M parameters)
    "cfg3_b2_k5_m80_t3": dict(B=2, K=5, M=80, D=4096, H=1024, T=3, sigma=0.1, scale=1.0, nclasses=40,
                              if_gamma_prior=1, stop_gamma_grad=0, if_sort=1, if_pooling=0),
    # BASELINE configs[2] at its real K and T (80 views, K=5, 20 iterations), one object
    "cfg3_b1_k5_m80_t20": dict(B=1, K=5, M=80, D=4096, H=1024, T=20, sigma=0.1, scale=1.0, nclasses=40,
                               if_gamma_prior=1, stop_gamma_grad=0, if_sort=1, if_pooling=0),
}


def seeded_models(cfg, seed=10):
    """The weight recipe shared by the golden generator, the tests and bench.py."""
    from multi_view_sort_b200 import NEM, Multi_View_Net
    torch.manual_seed(seed)
    nem = NEM(nclasses=cfg["nclasses"], view_num=cfg["M"], k=cfg["K"], input_dim=cfg["D"],
              rnn_hidden_size=cfg["H"], iter_num=cfg["T"], stop_gamma_grad=cfg["stop_gamma_grad"],
              init_type="bin_uniform", gamma_prior_loss=cfg["if_gamma_prior"], e_sigma=cfg["sigma"])
    head = Multi_View_Net(None, "vggm", nclasses=cfg["nclasses"], k=cfg["K"], rnn_hidden_size=cfg["H"],
                          if_sort=cfg["if_sort"], if_pooling=cfg["if_pooling"])
    return nem, head


def grad_summary(g: torch.Tensor, nsample=64):
    g = g.detach().double().reshape(-1)
    idx = (torch.arange(nsample, dtype=torch.int64) * (g.numel() - 1)) // (nsample - 1)
    return np.concatenate([[g.norm().item(), g.abs().max().item(), g.sum().item()], g[idx].numpy()])


def main():
    from oracle import ref_import
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    only = sys.argv[1:]
    for name, cfg in CASES.items():
        if only and name not in only:
            continue
        nem_mine, head_mine = seeded_models(cfg)
        torch.manual_seed(1234)   # the reference's own init is overwritten below
        rnem, rhead = ref_import.build_reference_models(
            k=cfg["K"], view_num=cfg["M"], input_dim=cfg["D"], rnn_hidden_size=cfg["H"], iter_num=cfg["T"],
            nclasses=cfg["nclasses"], e_sigma=cfg["sigma"], gamma_prior_loss=cfg["if_gamma_prior"],
            stop_gamma_grad=cfg["stop_gamma_grad"], if_sort=cfg["if_sort"], if_pooling=cfg["if_pooling"])
        rnem.load_state_dict(nem_mine.state_dict())      # strict: keys and shapes must match the reference
        rhead.load_state_dict(head_mine.state_dict())
        rnem.eval(); rhead.eval()
        x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
        r = ref_import.reference_train_step(rnem, rhead, x, y, cfg["T"])
        r["loss"].backward()
        rhead.save_fea = 1
        with torch.no_grad():
            desc = rhead(r["h"].detach(), r["gamma"].detach())
        out = dict(
            h=r["h"].detach().numpy(), gamma=r["gamma"].detach().numpy().reshape(cfg["B"], cfg["K"], cfg["M"]),
            pred_rowsum=r["pred"].detach().double().sum(-1).numpy(),
            nem_loss=r["nem_loss"].item(), intra=r["intra"].item(), inter=r["inter"].item(),
            closs=r["closs"].item(), loss=r["loss"].item(), logits=r["out"].detach().numpy(),
            descriptor=desc.numpy(), labels=y.numpy(),
            torch_version=np.array(torch.__version__), cfg=np.array(repr(cfg)),
        )
        for pname, prm in list(rnem.named_parameters()) + [("head." + n, p) for n, p in rhead.named_parameters()]:
            key = "grad/" + (pname if pname.startswith("head.") else "nem." + pname)
            if prm.grad is None:
                out[key] = np.array([np.nan])
            else:
                out[key] = grad_summary(prm.grad)
        path = os.path.join(GOLDEN_DIR, name + ".npz")
        np.savez_compressed(path, **out)
        print(name, "loss", out["loss"], "gamma range", out["gamma"].min(), out["gamma"].max(),
              "->", path, os.path.getsize(path), "bytes")


if __name__ == "__main__":
    main()
