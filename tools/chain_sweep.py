"""Parity of the bf16-features preset as a function of the forward accumulation-chain bound (64-wide k-blocks)."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nem_oracle as O
from oracle.make_golden import seeded_models
from helpers import load_golden, rel_err
from test_em_gpu import _oracle, _cuda

for name in ["cfg1_b8_k3_m12_t10", "prior0_b5_k2_m12_t3", "peaky_b6_k4_m12_t4"]:
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    xq = x.bfloat16().float()
    ref, pn = _oracle(cfg, nem, xq, torch.float32)
    for chain in (16, 32, 64, 0):
        for p in nem.parameters():
            p.grad = None
        res, nem_c = _cuda(cfg, nem, x, (2, 1, 1, chain, 0, 2), torch.bfloat16)
        named = dict(nem_c.named_parameters())
        errs = {k: rel_err(res[k].reshape(-1), ref[k].reshape(-1)) for k in ("h", "gamma", "nem_loss")}
        g = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
        gm = max(g, key=g.get)
        print(f"{name[:6]} chain {chain:3d} h {errs['h']:.1e} gamma {errs['gamma']:.1e} loss {errs['nem_loss']:.1e} grad max {g[gm]:.2e} ({gm}) "
              f"median {sorted(g.values())[len(g)//2]:.2e}", flush=True)
        nem = nem.cpu()
