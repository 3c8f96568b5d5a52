"""Kernel-time table of one training step (torch.profiler / CUPTI), to see the non-contraction share."""
import sys, os, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from multi_view_sort_b200.functional import cross_entropy
from torch.profiler import profile, ProfilerActivity
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4096
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda")
nem, head = bench.build_models(prec, dev)
opt = torch.optim.Adam(list(nem.parameters()) + list(head.parameters()), lr=5e-5, weight_decay=1e-4, fused=True)
x = torch.relu(torch.randn(B, 12, 4096, device=dev)).to(torch.bfloat16 if prec.startswith("bf16") else torch.float32)
y = torch.randint(0, 40, (B,), device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    h, gamma, nem_loss, _, _ = nem.run_em(x, 10)
    (cross_entropy(head(h, gamma), y) + nem_loss).backward()
    opt.step()
for _ in range(3): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.OrderedDict(); tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = ev.name.replace("(anonymous namespace)::", ""); name = re.sub(r"\(.*", "", name); name = re.sub(r"^void ", "", name)
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ev.device_time / 1e3 if hasattr(ev, "device_time") else ev.cuda_time / 1e3
        tot += ev.device_time / 1e3 if hasattr(ev, "device_time") else ev.cuda_time / 1e3
print(f"B={B} {prec}: total device kernel time {tot:.1f} ms")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:28]:
    print(f"{ms:9.2f} ms {100*ms/tot:5.1f}%  x{c:4d}  {k[:100]}")
# idle time between kernels of the profiled step (CUPTI timestamps): what a dependent-launch scheme could hide at most
evs = sorted([(ev.time_range.start, ev.time_range.end, ev.name) for ev in prof.events()
              if ev.device_type == torch.autograd.DeviceType.CUDA and ev.time_range.end > ev.time_range.start], key=lambda t: t[0])
if evs:
    span = (evs[-1][1] - evs[0][0]) / 1e3
    gaps = [(evs[i + 1][0] - max(e[1] for e in evs[max(0, i - 3):i + 1])) / 1e3 for i in range(len(evs) - 1)]
    pos = [g for g in gaps if g > 0]
    print(f"span first kernel start .. last kernel end {span:.1f} ms; idle between kernels {sum(pos):.2f} ms over {len(pos)} gaps "
          f"(median {sorted(pos)[len(pos) // 2] * 1e3:.1f} us, max {max(pos) * 1e3:.0f} us)")
    big = sorted(((g, evs[i][2], evs[i + 1][2]) for i, g in enumerate(gaps) if g > 0.02), reverse=True)[:8]
    for g, a, b in big:
        print(f"   {g * 1e3:7.0f} us between {a[:50]} -> {b[:50]}")
