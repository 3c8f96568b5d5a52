"""Accuracy of the E-step epilogue's chunk bodies against an fp64 restatement (fast chunk vs the general one)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import ops
torch.manual_seed(0)
def rnd(shape, seed, scale=1.0):
    return (torch.randn(shape, generator=torch.Generator().manual_seed(seed)) * scale).cuda()
for (B, K, Mv, D, zk, xdt) in [(3, 3, 12, 512, 256, torch.bfloat16), (16, 3, 12, 4096, 1024, torch.float32), (2, 5, 80, 1000, 1024, torch.bfloat16)]:
    sigma = 0.25
    rows = B * K * Mv
    z, w3, b3 = rnd((rows, zk), 30), rnd((D, zk), 31, 1.0 / 32), rnd((D,), 32, 0.1)
    x = torch.relu(rnd((B * Mv, D), 33)).to(xdt)
    zp, wp = ops.split_planes(z, 2, True), ops.split_planes(w3, 2, True)
    zq = zp[0].double() + zp[1].double() / 2048
    wq = wp[0].double() + wp[1].double() / 2048
    pred = zq @ wq.t() + b3.double()
    xr = x.double().view(B, 1, Mv, D).expand(B, K, Mv, D).reshape(rows, D)
    d = xr - pred
    e = torch.exp(-d * d / (2 * sigma * sigma))
    nb = D // 128 if D % 128 == 0 else None
    ref_e, ref_sq = e.sum(1), (d * d).sum(1)
    if D % 128 == 0:      # per (row, 128-column block) partials
        blk_e, blk_sq = e.view(rows, D // 128, 128).sum(2), (d * d).view(rows, D // 128, 128).sum(2)
        a = ops.estep_forward(zp, wp, b3, x, K, Mv, sigma, want_sq=True)
        b = ops.estep_forward_save(zp, wp, b3, x, K, Mv, sigma, delta_dtype=torch.float16)
        print("  per-block: fast e %.1e sq %.1e | slow e %.1e sq %.1e | fast vs slow e %.1e sq %.1e" % (
            ((a[0].double() - blk_e).abs().max() / blk_e.max()).item(), ((a[1].double() - blk_sq).abs().max() / blk_sq.max()).item(),
            ((b[0].double() - blk_e).abs().max() / blk_e.max()).item(), ((b[1].double() - blk_sq).abs().max() / blk_sq.max()).item(),
            ((a[0] - b[0]).abs().max() / b[0].max()).item(), ((a[1] - b[1]).abs().max() / b[1].max()).item()))
    def rel(a, b): return ((a.double() - b).abs().max() / b.abs().max()).item()
    res = {}
    pe, psq, _ = ops.estep_forward(zp, wp, b3, x, K, Mv, sigma, want_sq=True); res["fast  no-save sq"] = (rel(pe.sum(1), ref_e), rel(psq.sum(1), ref_sq))
    pe, psq, _ = ops.estep_forward(zp, wp, b3, x, K, Mv, sigma, want_sq=True, want_pp=True); res["slow  no-save sq+pp"] = (rel(pe.sum(1), ref_e), rel(psq.sum(1), ref_sq))
    pe, psq, dl = ops.estep_forward_save(zp, wp, b3, x, K, Mv, sigma, delta_dtype=torch.float16); res["slow  save fp16"] = (rel(pe.sum(1), ref_e), rel(psq.sum(1), ref_sq))
    pe, psq, dl = ops.estep_forward_save(zp, wp, b3, x, K, Mv, sigma, delta_dtype=torch.float32); res["fast  save fp32"] = (rel(pe.sum(1), ref_e), rel(psq.sum(1), ref_sq), rel(dl, d))
    print((B, K, Mv, D, zk, str(xdt)), {k: tuple(f"{v:.1e}" for v in vs) for k, vs in res.items()}, flush=True)
