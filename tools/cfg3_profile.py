"""Kernel-time table of one configs[2] training step (80 views, K=5, T=20) at batch B."""
import sys, os, collections, re
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from torch.profiler import profile, ProfilerActivity
from multi_view_sort_b200 import NEM, Multi_View_Net
from multi_view_sort_b200.functional import cross_entropy
B = int(sys.argv[1]) if len(sys.argv) > 1 else 128
dev = torch.device("cuda")
torch.manual_seed(10)
nem = NEM(nclasses=40, view_num=80, k=5, input_dim=4096, rnn_hidden_size=1024, iter_num=20, init_type="bin_uniform",
          gamma_prior_loss=1, e_sigma=0.1, precision="bf16").to(dev)
head = Multi_View_Net(None, "vggm", nclasses=40, k=5, rnn_hidden_size=1024, precision="bf16").to(dev)
x = torch.relu(torch.randn(B, 80, 4096, device=dev)).bfloat16()
y = torch.randint(0, 40, (B,), device=dev)
def step():
    for p in list(nem.parameters()) + list(head.parameters()): p.grad = None
    h, gamma, nem_loss, _, _ = nem.run_em(x, 20)
    (cross_entropy(head(h, gamma), y) + nem_loss).backward()
for _ in range(2): step()
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
    step(); torch.cuda.synchronize()
agg = collections.OrderedDict(); tot = 0.0
for ev in prof.events():
    if ev.device_type == torch.autograd.DeviceType.CUDA:
        name = re.sub(r"^void ", "", re.sub(r"\(.*", "", ev.name.replace("(anonymous namespace)::", "")))
        a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += ev.device_time / 1e3; tot += ev.device_time / 1e3
print(f"cfg3 B={B}: total device kernel time {tot:.1f} ms")
for k, (c, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1])[:18]:
    print(f"{ms:9.2f} ms {100*ms/tot:5.1f}%  x{c:4d}  {k[:100]}")
