"""Measure CUDA-vs-oracle error for every (forward planes, backward planes) combination.
Writes gpurun_out/precision_sweep.json.  Run on the GPU box."""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nem_oracle as O
from oracle.make_golden import seeded_models
from helpers import load_golden, rel_err, params_of
from test_em_gpu import _oracle, _cuda

out = {}
for name in ["cfg1_b8_k3_m12_t10", "prior0_b5_k2_m12_t3"]:
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    ref, pn = _oracle(cfg, nem, x, torch.float32)
    for combo in [(2, 1, 3, 4), (2, 1, 2, 16), (2, 1, 1, 16), (2, 1, 2, 0), (3, 0, 2, 16), (3, 0, 1, 16), (1, 1, 1, 16), (1, 0, 1, 0)]:
        for p in nem.parameters():
            p.grad = None
        res, nem_c = _cuda(cfg, nem, x, combo, torch.float32)
        named = dict(nem_c.named_parameters())
        errs = {k: rel_err(res[k].reshape(-1), ref[k].reshape(-1)) for k in ("h", "gamma", "nem_loss")}
        g = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
        errs["grad_max"] = max(g.values()); errs["grad_argmax"] = max(g, key=g.get)
        errs["grad_median"] = sorted(g.values())[len(g) // 2]
        out[f"{name}/{combo}"] = errs
        print(name, combo, errs, flush=True)
        nem = nem.cpu()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", "precision_sweep.json"), "w"), indent=1)
