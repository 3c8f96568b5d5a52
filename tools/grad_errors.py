"""Per-tensor gradient errors (max-norm and relative L2) of a preset against the oracle, cfg1 shape at batch B."""
import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nem_oracle as O
from oracle.make_golden import seeded_models
from helpers import load_golden, rel_err
from test_em_gpu import _oracle, _cuda
B = int(sys.argv[1]); presets = sys.argv[2:]
_, cfg = load_golden("cfg1_b8_k3_m12_t10"); cfg = dict(cfg, B=B)
nem, _ = seeded_models(cfg)
x, _ = O.make_inputs(B, cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
xb = x.bfloat16()
ref, pn = _oracle(cfg, nem, xb.float(), torch.float32)
rows = {}
for preset in presets:
    for p in nem.parameters(): p.grad = None
    pr = eval(preset) if preset.startswith("(") else preset
    res, nem_c = _cuda(cfg, nem, xb, pr, torch.bfloat16)
    named = dict(nem_c.named_parameters())
    for n, p in pn.items():
        if p.grad is None: continue
        l2 = ((named[n].grad.double().cpu() - p.grad.double()).norm() / p.grad.double().norm()).item()
        rows.setdefault(n, []).append((rel_err(named[n].grad, p.grad), l2))
    nem = nem.cpu()
print("tensor".ljust(22) + "".join(f"{p[:22]:>26s}" for p in presets))
for n, v in rows.items():
    print(n.ljust(22) + "".join(f"   max {a:.1e} l2 {b:.1e}" for a, b in v))
