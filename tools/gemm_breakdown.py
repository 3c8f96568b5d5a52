"""Per-shape breakdown of the contraction launches of one training step (device time by CUDA events)."""
import ctypes as C, sys, os, collections
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multi_view_sort_b200._lib import lib
from multi_view_sort_b200.functional import cross_entropy

class Rec(C.Structure):
    _fields_ = [("M", C.c_int), ("N", C.c_int), ("k_total", C.c_longlong), ("nseg", C.c_int), ("chain_kb", C.c_int),
                ("chains", C.c_int), ("splits", C.c_int), ("epilogue", C.c_int), ("a_mn", C.c_int), ("b_mn", C.c_int), ("ms", C.c_float)]

B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
if len(sys.argv) > 3: lib().vmm_gemm_force_ncta(int(sys.argv[3]))   # 1 = single CTAs, 2 = CTA pairs
dev = torch.device("cuda")
nem, head = bench.build_models(prec, dev)
x = torch.relu(torch.randn(B, 12, 4096, device=dev)).to(torch.bfloat16 if prec.startswith("bf16") else torch.float32)
y = torch.randint(0, 40, (B,), device=dev)
def step():
    for p in list(nem.parameters()) + list(head.parameters()): p.grad = None
    h, gamma, nem_loss, _, _ = nem.run_em(x, 10)
    (cross_entropy(head(h, gamma), y) + nem_loss).backward()
for _ in range(2): step()
torch.cuda.synchronize()
L = lib(); L.vmm_gemm_profile(1)
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); step(); e1.record(); torch.cuda.synchronize()
buf = (Rec * 4096)()
n = L.vmm_gemm_profile_records(buf, 4096)
L.vmm_gemm_profile(0)
agg = collections.OrderedDict()
for r in buf[:n]:
    key = (r.M, r.N, r.k_total, r.nseg, r.chain_kb, r.splits, "store red estep ebwd".split()[r.epilogue], r.a_mn, r.b_mn)
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += r.ms
tot = sum(v[1] for v in agg.values())
print(f"step {e0.elapsed_time(e1):.1f} ms, contraction launches {n}, total {tot:.1f} ms")
print(f"{'M':>7} {'N':>6} {'Ktot':>7} seg chain spl epi   amn bmn  cnt   ms_tot  share  TF/s(exec)")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    M, N, Kt, ns, ch, sp, ep, am, bm = k
    tf = 2.0 * M * N * Kt * c / (ms * 1e-3) / 1e12
    print(f"{M:7d} {N:6d} {Kt:7d} {ns:3d} {ch:5d} {sp:3d} {ep:5s} {am:3d} {bm:3d} {c:4d} {ms:8.2f} {100*ms/tot:5.1f}% {tf:8.0f}")
