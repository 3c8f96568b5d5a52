"""Small end-to-end exercise of every kernel family (pair GEMM in all layouts, fused E-step with delta save, streaming
E-step backward, EM forward/backward, head) for compute-sanitizer runs:
    compute-sanitizer --tool memcheck python tools/sanitize_small.py"""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from multi_view_sort_b200 import NEM, Multi_View_Net, ops
from multi_view_sort_b200.functional import cross_entropy
torch.manual_seed(0)
for amn in (False, True):
    for bmn in (False, True):
        A = torch.randn((200, 384) if amn else (384, 200), device="cuda").bfloat16()
        B = torch.randn((200, 520) if bmn else (520, 200), device="cuda").bfloat16()
        out = ops.gemm([(A, B)], amn, bmn)
        ref = (A.t() if amn else A).float() @ (B if bmn else B.t()).float()
        assert torch.allclose(out, ref, atol=1e-2, rtol=1e-3), (amn, bmn)
nem = NEM(nclasses=10, view_num=12, k=3, input_dim=256, rnn_hidden_size=1024, iter_num=2, init_type="bin_uniform",
          gamma_prior_loss=1, e_sigma=0.1, precision="bf16").cuda()
head = Multi_View_Net(None, "vggm", nclasses=10, k=3, rnn_hidden_size=1024, precision="bf16").cuda()
x = torch.relu(torch.randn(7, 12, 256, device="cuda")).bfloat16()
y = torch.randint(0, 10, (7,), device="cuda")
h, gamma, nem_loss, _, _ = nem.run_em(x, 2)
(cross_entropy(head(h, gamma), y) + nem_loss).backward()
with torch.no_grad():
    nem.eval(); head.eval()
    h, gamma, *_ = nem.run_em(x, 2)
    head(h, gamma)
torch.cuda.synchronize()
print("sanitize_small: ok", float(nem_loss))
