"""Does tensor-core fp32 accumulation limit accuracy at long K?  3-plane (fp32-class operands) GEMM
error vs K, with one accumulation chain vs split into short chains (split-K + fp32 atomics)."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import ops
torch.manual_seed(0)
M, N = 256, 512
for K in (256, 1024, 4096, 16384, 65536):
    A = torch.randn(M, K, device="cuda"); B = torch.randn(N, K, device="cuda") / K ** 0.5
    ref = A.double() @ B.double().t()
    fp32 = (A @ B.t()).double()
    row = [f"K={K:6d} torch-fp32 {((fp32-ref).abs().max()/ref.abs().max()).item():.2e}"]
    for n in (2, 3):
        pairs = ops.plane_pairs(ops.split_planes(A, n), ops.split_planes(B, n))
        for splits in (1, max(1, K // 1024), max(1, K // 256)):
            out = torch.zeros(M, N, device="cuda")
            ops.gemm(pairs, out=out, accumulate=True, splits=splits)
            torch.cuda.synchronize()
            e = ((out.double() - ref).abs().max() / ref.abs().max()).item()
            b = ((out.double() - ref).mean() / ref.abs().mean()).item()
            row.append(f"p{n}/s{splits}: {e:.2e} (bias {b:+.1e})")
    print("  ".join(row), flush=True)
# positive operands make the RZ bias visible
for K in (1024, 16384):
    A = torch.rand(M, K, device="cuda"); B = torch.rand(N, K, device="cuda") / K
    ref = A.double() @ B.double().t()
    for splits in (1, max(1, K // 1024)):
        out = torch.zeros(M, N, device="cuda")
        ops.gemm(ops.plane_pairs(ops.split_planes(A, 3), ops.split_planes(B, 3)), out=out, accumulate=True, splits=splits)
        torch.cuda.synchronize()
        rel = (out.double() - ref) / ref
        print(f"positive K={K} splits={splits}: mean rel {rel.mean().item():+.2e} max {rel.abs().max().item():.2e}; torch fp32 mean rel {(((A@B.t()).double()-ref)/ref).mean().item():+.2e}", flush=True)
