"""Kernel-time table of one forward-only (no_grad) EM + head call - the configs[4] inference path."""
import sys, os, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16384
K = int(sys.argv[2]) if len(sys.argv) > 2 else 3
T = int(sys.argv[3]) if len(sys.argv) > 3 else 10
dev = torch.device("cuda")
nem, head = bench.build_models("bf16", dev, K=K, T=T)
nem.eval(); head.eval()
x = torch.relu(torch.randn(B, 12, 4096, device=dev)).bfloat16()
with torch.no_grad():
    for _ in range(2): head(*nem.run_em(x, T)[:2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); head(*nem.run_em(x, T)[:2]); e1.record(); torch.cuda.synchronize()
    print(f"B={B} K={K} T={T}: {e0.elapsed_time(e1):.2f} ms per call, {B / e0.elapsed_time(e1) * 1e3:.0f} objects/s")
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        head(*nem.run_em(x, T)[:2]); torch.cuda.synchronize()
agg = collections.OrderedDict(); tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", ""); name = re.sub(r"\(.*", "", name); name = re.sub(r"^void ", "", name)
        t = (ev.device_time if hasattr(ev, "device_time") else ev.cuda_time) / 1e3
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += t; tot += t
print(f"total device kernel time {tot:.1f} ms")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:20]:
    print(f"{ms:9.2f} ms {100*ms/tot:5.1f}%  x{c:4d}  {k[:100]}")
