"""A/B timing of the contraction kernel as single CTAs vs CTA pairs (vmm_gemm_force_ncta) against cuBLAS, on the
shapes of one training step at batch B (sustained: 20 back-to-back launches per measurement)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import ops
from multi_view_sort_b200._lib import lib

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
shapes = [  # name, M, N, K, a_mn, b_mn, accumulate
    ("enc2 fwd", R, 4096, 12288, False, False, False),
    ("dec2 fwd", R, 12288, 4096, False, False, False),
    ("dec3 fwd (store)", RM, 4096, 1024, False, False, False),
    ("dgrad dv5.W5", R, 4096, 12288, False, True, False),
    ("wgrad dv5^T.u", 12288, 4096, R, True, True, True),
    ("wgrad dgi^T.e", 3072, 4096, R, True, True, True),
    ("wgrad dpred^T.z", 4096, 1024, RM, True, True, True),
    ("wgrad dg1^T.z", 1024, 1024, RM, True, True, True),
]
L = lib()
for name, m, n, k, amn, bmn, acc in shapes:
    A = torch.randn((k, m) if amn else (m, k), device="cuda").bfloat16()
    Bm = torch.randn((k, n) if bmn else (n, k), device="cuda").bfloat16()
    out = torch.zeros(m, n, device="cuda")
    row = []
    for ncta in (1, 2):
        L.vmm_gemm_force_ncta(ncta)
        ms = timeit(lambda: ops.gemm([(A, Bm)], amn, bmn, out=out, accumulate=acc))
        row.append((ms, 2.0 * m * n * k / ms / 1e9))
    L.vmm_gemm_force_ncta(0)
    ref = timeit(lambda: torch.matmul(A.t() if amn else A, Bm if bmn else Bm.t()))
    print(f"{name:18s} M={m:7d} N={n:6d} K={k:7d} | 1cta {row[0][0]:7.3f} ms {row[0][1]:7.0f} TF/s | pair {row[1][0]:7.3f} ms {row[1][1]:7.0f} TF/s"
          f" | cuBLAS {ref:7.3f} ms {2.0*m*n*k/ref/1e9:7.0f} TF/s", flush=True)
