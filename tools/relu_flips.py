"""Where does the bf16 preset's dec1 gradient error come from?  ("What's weak" #1.)

For every preset: run the CUDA training forward/backward at the cfg1 shape (batch B), read dec1's pre-activation
out of the library's workspace (raw contraction output g4 + LayerNorm row statistics), compare the ReLU mask with
the oracle's, and report the gradient error against (a) the plain oracle and (b) the oracle evaluated with the
CUDA run's mask (`dec1_masks=`): if the dec1 error is a derivative decision at the kink, (b) collapses to the level of
the other tensors.

    python tools/relu_flips.py B preset [preset ...]
        preset = a name, a tuple literal "(2,1,1,16,0,2)", or "bf16:enc2:a,dec2:w,dec3:s" = the bf16 preset's planes with
        per-contraction forward modes (s single pass, a activation pair x weight hi, w activation hi x weight pair, f full)
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
from oracle import nem_oracle as O
from oracle.make_golden import seeded_models
from helpers import load_golden, rel_err, params_of, cuda_dec1_pre
from test_em_gpu import _probe


def main():
    B = int(sys.argv[1]); presets = sys.argv[2:]
    _, cfg = load_golden("cfg1_b8_k3_m12_t10"); cfg = dict(cfg, B=B)
    nem, _ = seeded_models(cfg)
    x, _ = O.make_inputs(B, cfg["M"], cfg["D"], cfg["nclasses"], seed=10, scale=cfg["scale"])
    xb = x.bfloat16()
    wh, wg = _probe(cfg)

    def oracle(masks=None):
        pn = params_of(nem.cpu(), dtype=torch.float32)
        pre = []
        r = O.run_em(pn, xb.float(), cfg["K"], cfg["T"], cfg["sigma"], "bin_uniform", cfg["if_gamma_prior"],
                     cfg["stop_gamma_grad"], dec1_masks=masks, dec1_pre=pre)
        (r["nem_loss"] + (r["h"] * wh).sum() + (r["gamma"] * wg).sum()).backward()
        return r, pn, pre

    ref, pn, ref_pre = oracle()
    report = {}
    for preset in presets:
        for p in nem.parameters():
            p.grad = None
        if preset.startswith("bf16:") or preset.startswith("bf16h:"):
            from multi_view_sort_b200.functional import fwd_modes
            code = {"s": 1, "a": 2, "w": 3, "f": 0}
            kw = {it.split(":")[0]: code[it.split(":")[1]] for it in preset.split(":", 1)[1].split(",") if it}
            pr = (2, 1, 1, 16, fwd_modes(**kw), 2) + ((1,) if preset.startswith("bf16h:") else ())
        else:
            pr = eval(preset) if preset.startswith("(") else preset
        nc = nem.cuda().eval()
        nc.precision = pr
        nc.keep_workspace = True
        h, gamma, nem_loss, _, _ = nc.run_em(xb.cuda(), cfg["T"])
        (nem_loss + (h * wh.cuda()).sum() + (gamma * wg.cuda()).sum()).backward()
        torch.cuda.synchronize()
        pre = cuda_dec1_pre(nc, cfg["T"])
        grads = {n: p.grad.detach().cpu().clone() for n, p in nc.named_parameters() if p.grad is not None}
        flips = [int(((a > 0) != (b > 0)).sum()) for a, b in zip(pre, ref_pre)]
        pre_err = max((a - b).abs().max().item() for a, b in zip(pre, ref_pre))
        kink = [float(b[(a > 0) != (b > 0)].abs().max()) if f else 0.0 for a, b, f in zip(pre, ref_pre, flips)]
        pmax = max(b.abs().max().item() for b in ref_pre)
        allk = torch.cat([b[(a > 0) != (b > 0)].abs() for a, b in zip(pre, ref_pre)]) / pmax
        hist = {f"<{t:g}": int((allk < t).sum()) for t in (1e-6, 1e-5, 1e-4, 1e-3, 1e-2)}
        masks = [(a > 0).float() for a in pre]
        _, pm, _ = oracle(masks)
        e_plain = {n: rel_err(grads[n], p.grad) for n, p in pn.items() if p.grad is not None}
        e_mask = {n: rel_err(grads[n], p.grad) for n, p in pm.items() if p.grad is not None}
        report[preset] = dict(flips=flips, flip_hist_rel=hist, pre_err=pre_err, kink_max=max(kink), h=rel_err(h, ref["h"]),
                              gamma=rel_err(gamma.reshape(-1), ref["gamma"].reshape(-1)), plain=e_plain, forced=e_mask)
        worst_p, worst_m = max(e_plain, key=e_plain.get), max(e_mask, key=e_mask.get)
        print(f"B={B} {preset}: flips per iteration {flips} (largest |pre| of a flipped unit {max(kink):.1e}; max |pre_cuda - pre_oracle| "
              f"{pre_err:.1e}); h {report[preset]['h']:.1e} gamma {report[preset]['gamma']:.1e}; flipped units by |pre|/max|pre| {hist} of {sum(b.numel() for b in ref_pre)}")
        print(f"   vs plain oracle : worst {e_plain[worst_p]:.2e} ({worst_p}); dec1.0.weight {e_plain['dec1.0.weight']:.2e} "
              f"dec1.0.bias {e_plain['dec1.0.bias']:.2e} dec1.1.bias {e_plain['dec1.1.bias']:.2e} dec1.1.weight {e_plain['dec1.1.weight']:.2e}")
        print(f"   vs CUDA's mask  : worst {e_mask[worst_m]:.2e} ({worst_m}); dec1.0.weight {e_mask['dec1.0.weight']:.2e} "
              f"dec1.0.bias {e_mask['dec1.0.bias']:.2e} dec1.1.bias {e_mask['dec1.1.bias']:.2e} dec1.1.weight {e_mask['dec1.1.weight']:.2e}", flush=True)
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    json.dump(report, open(os.path.join(ROOT, "gpurun_out", f"relu_flips_b{B}.json"), "w"), indent=1)


if __name__ == "__main__":
    main()
