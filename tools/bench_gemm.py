"""Time the tcgen05 contraction on the shapes of the VMM path (CUDA events, L2-cold inputs)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import ops

def timeit(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); fn(); b.record(); torch.cuda.synchronize(); ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts)//2]

def main():
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
    K, M = 3, 12
    R, RM = B*K, B*K*M
    shapes = {  # name: (M, N, K, a_mn, b_mn)
        "enc1r  z.W13^T": (RM, 1024, 1024, False, False),
        "enc2   a.W2^T": (R, 4096, 12288, False, False),
        "gru_ih e.Wih^T": (R, 3072, 4096, False, False),
        "dec2   u.W5^T": (R, 12288, 4096, False, False),
        "dgrad  dG5.W5": (R, 4096, 12288, False, True),
        "wgrad  dG5^T.u": (12288, 4096, R, True, True),
        "wgrad  dG1^T.z": (1024, 1024, RM, True, True),
    }
    res = {}
    for name, (m, n, k, amn, bmn) in shapes.items():
        A = torch.randn((k, m) if amn else (m, k), device="cuda").bfloat16()
        Bm = torch.randn((k, n) if bmn else (n, k), device="cuda").bfloat16()
        acc = name.startswith("wgrad")
        out = torch.zeros(m, n, device="cuda")
        ms = timeit(lambda: ops.gemm([(A, Bm)], amn, bmn, out=out, accumulate=acc))
        tf = 2.0*m*n*k/ms/1e9
        ref = timeit(lambda: torch.matmul(A.t() if amn else A, Bm if bmn else Bm.t()))
        res[name] = dict(ms=round(ms, 4), tflops=round(tf, 1), cublas_ms=round(ref, 4), cublas_tflops=round(2.0*m*n*k/ref/1e9, 1))
        print(f"{name:18s} M={m:7d} N={n:6d} K={k:7d}  {ms:8.3f} ms {tf:8.1f} TF/s | cuBLAS {ref:8.3f} ms {2.0*m*n*k/ref/1e9:8.1f} TF/s", flush=True)
    # fused E-step
    zk, D = 1024, 4096
    z = torch.randn(RM, zk, device="cuda"); w3 = torch.randn(D, zk, device="cuda")/32; b3 = torch.zeros(D, device="cuda")
    x = torch.relu(torch.randn(B*M, D, device="cuda"))
    for planes, xd in ((1, torch.bfloat16), (2, torch.float32)):
        zp, wp = ops.split_planes(z, planes), ops.split_planes(w3, planes)
        xx = x.to(xd)
        ms = timeit(lambda: ops.estep_forward(zp, wp, b3, xx, K, M, 0.1))
        print(f"estep fwd planes={planes}: {ms:.3f} ms  {2.0*RM*D*zk/ms/1e9:.1f} TF/s(alg)", flush=True)
        al = torch.randn(RM, device="cuda"); be = torch.randn(RM, device="cuda")
        ms = timeit(lambda: ops.estep_backward(zp, wp, b3, xx, K, M, 0.1, al, be))
        print(f"estep bwd planes={planes}: {ms:.3f} ms  {2.0*RM*D*zk/ms/1e9:.1f} TF/s(alg)", flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(res, open("gpurun_out/bench_gemm.json", "w"), indent=1)

if __name__ == "__main__":
    main()
