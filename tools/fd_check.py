import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nem_oracle as O
from oracle.make_golden import seeded_models
from helpers import load_golden
from multi_view_sort_b200.functional import cross_entropy
_, cfg = load_golden("prior0_b5_k2_m12_t3")
nem, head = seeded_models(cfg)
x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
nem, head = nem.cuda(), head.cuda()
nem.precision = head.precision = "fp32_strict"
xc, yc = x.cuda(), y.cuda()
def loss_fn():
    h, gamma, nem_loss, _, _ = nem.run_em(xc, cfg["T"])
    return cross_entropy(head(h, gamma), yc) + nem_loss
for mode in ("eval", "train", "train_nem_only", "train_head_only"):
    nem.train(mode in ("train", "train_nem_only")); head.train(mode in ("train", "train_head_only"))
    nem.dropout_seed = head.dropout_seed = 1234
    for p in list(nem.parameters()) + list(head.parameters()): p.grad = None
    loss_fn().backward()
    g = torch.Generator().manual_seed(1)
    for mod, pname in ((nem,"dec1.0.bias"), (nem,"enc2.1.weight"), (nem,"enc1.0.bias"), (nem,"dec2.0.bias"), (nem, "dec3.bias"), (head,"net.3.bias"), (head,"net.0.bias"), (head, "net.6.bias")):
        prm = dict(mod.named_parameters())[pname]
        d = torch.randn(prm.shape, generator=g).cuda()
        analytic = (prm.grad * d).sum().item()
        for eps in (1e-2, 3e-3):
            with torch.no_grad():
                prm.add_(eps * d); lp = loss_fn().item()
                prm.sub_(2 * eps * d); lm = loss_fn().item()
                prm.add_(eps * d)
            print(f"{mode:16s} {pname:14s} eps {eps:.0e} numeric {(lp-lm)/(2*eps):+.5f} analytic {analytic:+.5f}", flush=True)
