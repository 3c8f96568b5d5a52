"""BASELINE configs[2]: 80 views x 4096-d, K = 5, T = 20 (NEM(view_num=80): 700 M parameters): training-step time
and objects/s at batch B (default 256), bf16 features.  Reference-algorithm FLOPs: 814.1 GFLOP/object fwd+bwd."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import NEM, Multi_View_Net
from multi_view_sort_b200.functional import cross_entropy
B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
prec = sys.argv[2] if len(sys.argv) > 2 else "bf16"
dev = torch.device("cuda")
torch.manual_seed(10)
nem = NEM(nclasses=40, view_num=80, k=5, input_dim=4096, rnn_hidden_size=1024, iter_num=20, init_type="bin_uniform",
          gamma_prior_loss=1, e_sigma=0.1, precision=prec).to(dev)
head = Multi_View_Net(None, "vggm", nclasses=40, k=5, rnn_hidden_size=1024, precision=prec).to(dev)
opt = torch.optim.Adam(list(nem.parameters()) + list(head.parameters()), lr=5e-5, weight_decay=1e-4, fused=True)
x = torch.relu(torch.randn(B, 80, 4096, device=dev)).to(torch.bfloat16 if prec.startswith("bf16") else torch.float32)
y = torch.randint(0, 40, (B,), device=dev)
def step():
    opt.zero_grad(set_to_none=True)
    h, gamma, nem_loss, _, _ = nem.run_em(x, 20)
    (cross_entropy(head(h, gamma), y) + nem_loss).backward()
    opt.step()
for _ in range(2): step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): step()
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 3
rec = dict(config="configs[2]: 80 views, K=5, T=20", precision=prec, batch=B, ms_per_step=ms, objects_per_s=B / ms * 1e3,
           alg_tflops=B / ms * 1e3 * 814.1e9 / 1e12, frac_of_bf16_peak=B / ms * 1e3 * 814.1e9 / 1e12 / 1413.0,
           peak_mem_gb=torch.cuda.max_memory_allocated() / 1e9)
print(json.dumps(rec))
os.makedirs("gpurun_out", exist_ok=True)
json.dump(rec, open("gpurun_out/cfg3_bench.json", "w"), indent=1)
