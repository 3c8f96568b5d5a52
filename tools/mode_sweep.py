"""Per-contraction forward precision sweep: CUDA-vs-oracle error when ONE forward contraction (or a named
combination) drops from the fp16 pair to fewer tensor-core passes, and with the E-step backward fed from the
saved delta.  Writes gpurun_out/mode_sweep.json.  Run on the GPU box.

    python tools/mode_sweep.py [bf16|fp32] [combo ...]      combo = name:mode,name:mode,...[+d1|+d2]
"""
import json, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nem_oracle as O
from oracle.make_golden import seeded_models
from helpers import load_golden, rel_err
from test_em_gpu import _oracle, _cuda
from multi_view_sort_b200.functional import FWD_CONTRACTIONS, fwd_modes

which = sys.argv[1] if len(sys.argv) > 1 else "bf16"
base = (2, 1, 1, 16) if which == "bf16" else (2, 1, 2, 16)
xdt = torch.bfloat16 if which == "bf16" else torch.float32
MODE = {"s": 1, "a": 2, "w": 3, "f": 0}

combos = []
if len(sys.argv) > 2:
    for spec in sys.argv[2:]:
        delta = 0
        if "+d" in spec:
            spec, dd = spec.split("+d")
            delta = int(dd)
        kw, extra = {}, 0
        for item in filter(None, spec.split(",")):
            n, m = item.split(":")
            if n == "wgrad": extra |= 1 << 18          # weight-gradient contractions on the leading plane only
            elif n == "dgrad": extra |= 1 << 19        # data-gradient contractions on the leading plane only
            else: kw[n] = MODE[m]
        combos.append((sys.argv[2:][len(combos)], fwd_modes(**kw) | extra, delta))
else:
    combos.append(("base", 0, 0))
    combos.append(("delta_f16", 0, 1))
    combos.append(("delta_f32", 0, 2))
    for n in FWD_CONTRACTIONS:
        combos.append((n + ":s", fwd_modes(**{n: 1}), 0))
    for n in ("enc2", "dec2", "dec3", "ebwd", "g1"):
        combos.append((n + ":a", fwd_modes(**{n: 2}), 0))
        combos.append((n + ":w", fwd_modes(**{n: 3}), 0))

out = {}
for name in ["cfg1_b8_k3_m12_t10", "prior0_b5_k2_m12_t3", "peaky_b6_k4_m12_t4"]:
    _, cfg = load_golden(name)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    xq = x.to(xdt).float()
    ref, pn = _oracle(cfg, nem, xq, torch.float32)
    for label, modes, delta in combos:
        for p in nem.parameters():
            p.grad = None
        res, nem_c = _cuda(cfg, nem, x, base + (modes, delta), xdt)
        named = dict(nem_c.named_parameters())
        errs = {k: rel_err(res[k].reshape(-1), ref[k].reshape(-1)) for k in ("h", "gamma", "nem_loss")}
        g = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
        errs["grad_max"] = max(g.values()); errs["grad_argmax"] = max(g, key=g.get)
        errs["grad_median"] = sorted(g.values())[len(g) // 2]
        out[f"{name}/{label}"] = errs
        print(f"{name[:6]:6s} {label:14s} h {errs['h']:.1e} gamma {errs['gamma']:.1e} loss {errs['nem_loss']:.1e} "
              f"grad max {errs['grad_max']:.2e} ({errs['grad_argmax']}) median {errs['grad_median']:.2e}", flush=True)
        nem = nem.cpu()
os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
json.dump(out, open(os.path.join(ROOT, "gpurun_out", f"mode_sweep_{which}.json"), "w"), indent=1)
