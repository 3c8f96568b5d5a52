import sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nem_oracle as O
from helpers import rel_err, params_of
from multi_view_sort_b200 import NEM
B, K, M, D, H = 4, 3, 12, 256, 128
for init in ("onehot", "rand", "uniform"):
    for T in (1, 2):
        for prec in ("fp32", "fp32_strict", (1,0,1,0)):
            torch.manual_seed(10)
            nem = NEM(view_num=M, k=K, input_dim=D, rnn_hidden_size=H, iter_num=T, init_type="rand", gamma_prior_loss=1, e_sigma=0.15).eval()
            x, _ = O.make_inputs(B, M, D, 40, seed=4, scale=0.5)
            if init == "rand":
                g0 = torch.rand(B, K, M, 1, generator=torch.Generator().manual_seed(9)); g0 = g0 / g0.sum(1).unsqueeze(1)
            elif init == "uniform":
                g0 = torch.full((B, K, M, 1), 1.0 / K)
            else:
                g0 = O.bin_uniform(B, K, M)[0]
            gp = torch.rand(B, K, M, 1, generator=torch.Generator().manual_seed(8)); gp = gp / gp.sum(2).unsqueeze(2)
            pn = params_of(nem)
            r = O.run_em(pn, x, K, T, 0.15, "rand", 1, 0, gamma0=g0, gamma_prior=gp)
            wh = torch.randn(B * K, H, generator=torch.Generator().manual_seed(1)) * 0.01
            (r["nem_loss"] + (r["h"] * wh).sum()).backward()
            nem = nem.cuda(); nem.precision = prec
            h, gamma, nem_loss, _, _ = nem.run_em(x.cuda(), T, gamma0=g0.cuda(), gamma_prior=gp.cuda())
            (nem_loss + (h * wh.cuda()).sum()).backward()
            named = dict(nem.named_parameters())
            errs = {n: rel_err(named[n].grad, p.grad) for n, p in pn.items() if p.grad is not None}
            top = sorted(errs.items(), key=lambda kv: -kv[1])[:5]
            print(init, "T", T, prec, "h", f"{rel_err(h, r['h']):.1e}", " ".join(f"{n}={v:.1e}" for n, v in top), flush=True)
