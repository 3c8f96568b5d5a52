"""A/B timing of operand layouts for the data-gradient contractions: B operand (the weight) MN-major (read in place, as the
backward pass does) vs K-major (a transposed copy), same M, N, K (sustained: 20 back-to-back launches)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import ops

def timeit(fn, reps=20, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(reps): fn()
    b.record(); torch.cuda.synchronize()
    return a.elapsed_time(b) / reps

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
K, M = 3, 12
R, RM = B * K, B * K * M
shapes = [("dgrad dpred.W3", RM, 1024, 4096), ("dgrad dv5.W5", R, 4096, 12288), ("dgrad dv2.W2", R, 12288, 4096),
          ("dgrad dG1.W13", RM, 1024, 1024), ("dgrad dgi.Wih", R, 4096, 3072)]
for name, m, n, k in shapes:
    A = torch.randn(m, k, device="cuda").bfloat16()
    out = torch.zeros(m, n, device="cuda")
    res = []
    for bmn in (True, False):
        Bm = torch.randn((k, n) if bmn else (n, k), device="cuda").bfloat16()
        ms = timeit(lambda: ops.gemm([(A, Bm)], False, bmn, out=out))
        res.append((ms, 2.0 * m * n * k / ms / 1e9))
    print(f"{name:16s} M={m:7d} N={n:6d} K={k:6d} | B MN-major {res[0][0]:7.3f} ms {res[0][1]:6.0f} TF/s | B K-major {res[1][0]:7.3f} ms {res[1][1]:6.0f} TF/s", flush=True)
