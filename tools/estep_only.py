"""Only the fused dec3 + E-step launch (single fp16 pass, fp32 delta save), for ncu captures and timing."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import ops
B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
K, M, D, zk = 3, 12, 4096, 1024
RM = B * K * M
z = torch.randn(RM, zk, device="cuda"); w3 = torch.randn(D, zk, device="cuda") / 32; b3 = torch.zeros(D, device="cuda")
x = torch.relu(torch.randn(B * M, D, device="cuda")).bfloat16()
zp, wp = ops.split_planes(z, 1, True), ops.split_planes(w3, 1, True)
buf = ops.estep_forward_save(zp, wp, b3, x, K, M, 0.1, delta_dtype=torch.float32, want_sq=False)
buf_sq = ops.estep_forward_save(zp, wp, b3, x, K, M, 0.1, delta_dtype=torch.float32, want_sq=True)
# "save fp32": what iterations 0..T-2 of a training step launch (fast chunk); "+ loss sums": the last iteration's launch
variants = {"save fp32": lambda: ops.estep_forward_save(zp, wp, b3, x, K, M, 0.1, delta_dtype=torch.float32, want_sq=False, out=buf)}
if len(sys.argv) > 2:
    variants["save fp32 + loss sums"] = lambda: ops.estep_forward_save(zp, wp, b3, x, K, M, 0.1, delta_dtype=torch.float32, out=buf_sq)
    variants["save fp16"] = lambda: ops.estep_forward_save(zp, wp, b3, x, K, M, 0.1, delta_dtype=torch.float16)
    variants["no save"] = lambda: ops.estep_forward(zp, wp, b3, x, K, M, 0.1, want_sq=False)
    out = torch.empty(RM, D, device="cuda")
    variants["plain store"] = lambda: ops.gemm([(zp[0], wp[0])], out=out)
for name, run in variants.items():
    for _ in range(3): run()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(10): run()
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 10
    print(f"estep {name:22s} B={B}: {ms:.3f} ms  {2.0*RM*D*zk/ms/1e9:.0f} TF/s")
