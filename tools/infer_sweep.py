"""BASELINE configs[4]: inference throughput sweep - forward-only EM + head over synthetic objects,
12 views, K in {2,3,4,6}, iters in {3,10}.  Objects are generated on the device in chunks (1 M objects
of fp32 features are 197 GB).  Writes gpurun_out/infer_sweep.json."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import NEM, Multi_View_Net

chunk = int(sys.argv[1]) if len(sys.argv) > 1 else 8192
nchunks = int(sys.argv[2]) if len(sys.argv) > 2 else 4
precisions = sys.argv[3].split(",") if len(sys.argv) > 3 else ["bf16", "bf16_fast"]
dev = torch.device("cuda")
out = []
for prec in precisions:
    for K in (2, 3, 4, 6):
        torch.manual_seed(10)
        nem = NEM(view_num=12, k=K, input_dim=4096, rnn_hidden_size=1024, init_type="bin_uniform", gamma_prior_loss=1,
                  e_sigma=0.1, precision=prec).to(dev).eval()
        head = Multi_View_Net(None, "vggm", nclasses=40, k=K, rnn_hidden_size=1024, precision=prec).to(dev).eval()
        for T in (3, 10):
            x = torch.relu(torch.randn(chunk, 12, 4096, device=dev)).bfloat16()
            with torch.no_grad():
                for _ in range(2):
                    h, g, *_ = nem.run_em(x, T); head(h, g)
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(nchunks):
                    h, g, *_ = nem.run_em(x, T)
                    pred = head(h, g).argmax(1)
                e1.record(); torch.cuda.synchronize()
            ms = e0.elapsed_time(e1)
            rate = chunk * nchunks / (ms * 1e-3)
            gflop = T * 2 * K * (16777216 * 12 + 19922944) / 1e9          # reference-algorithm forward GEMM FLOPs / object
            rec = dict(precision=prec, K=K, T=T, objects=chunk * nchunks, objects_per_s=rate,
                       alg_tflops=rate * gflop / 1e3, frac_of_bf16_peak=rate * gflop / 1e3 / 1413.0)
            out.append(rec); print(rec, flush=True)
        del nem, head
        torch.cuda.empty_cache()
os.makedirs("gpurun_out", exist_ok=True)
json.dump(out, open("gpurun_out/infer_sweep.json", "w"), indent=1)
