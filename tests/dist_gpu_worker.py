"""Worker of tests/test_dist_gpu.py (launched by torch.distributed.run, one rank per GPU, NCCL): the CUDA training step
on this rank's shard + GradBucket.allreduce_mean() must equal the 1-rank full-batch gradients."""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from multi_view_sort_b200 import NEM, Multi_View_Net                      # noqa: E402
from multi_view_sort_b200.dist import GradBucket, loss_weight, shard_range  # noqa: E402
from multi_view_sort_b200.functional import cross_entropy                  # noqa: E402


def main():
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    backend = os.environ.get("VMM_TEST_BACKEND", "nccl")
    if backend != "nccl":
        local = 0                                                        # single-GPU box: both ranks compute on cuda:0
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if backend == "nccl":
        dist.init_process_group("nccl", device_id=dev)
    else:
        dist.init_process_group(backend)
    cfg = dict(K=2, M=12, D=512, H=256, T=3, sigma=0.1, nclasses=10)
    worst = 0.0
    for B, attach in ((8, True), (8, False), (7, True)):                 # equal shards (both bucket paths), ragged shards
        torch.manual_seed(10)
        nem = NEM(nclasses=cfg["nclasses"], view_num=cfg["M"], k=cfg["K"], input_dim=cfg["D"], rnn_hidden_size=cfg["H"],
                  iter_num=cfg["T"], init_type="bin_uniform", gamma_prior_loss=1, e_sigma=cfg["sigma"], precision="fp32").to(dev).eval()
        head = Multi_View_Net(None, "vggm", nclasses=cfg["nclasses"], k=cfg["K"], rnn_hidden_size=cfg["H"], precision="fp32").to(dev).eval()
        g = torch.Generator().manual_seed(10)
        x = torch.relu(torch.randn(B, cfg["M"], cfg["D"], generator=g)).to(dev)
        y = torch.randint(0, cfg["nclasses"], (B,), generator=g).to(dev)
        params = list(nem.parameters()) + list(head.parameters())

        def step(xs, ys, weight):
            for p in params:
                p.grad = None
            h, gamma, nem_loss, _, _ = nem.run_em(xs, cfg["T"])
            loss = cross_entropy(head(h, gamma), ys) + nem_loss
            (loss * weight).backward()

        step(x, y, 1.0)                                                  # 1-rank full batch
        full = {n: p.grad.clone() for n, p in list(nem.named_parameters()) + list(head.named_parameters()) if p.grad is not None}
        bucket = GradBucket(params)
        if attach:
            bucket.attach(nem, head)
        b, e = shard_range(B, rank, world)
        step(x[b:e], y[b:e], loss_weight(e - b, B, world))
        nbytes = bucket.allreduce_mean()
        torch.cuda.synchronize()
        assert nbytes > 0
        for n, p in list(nem.named_parameters()) + list(head.named_parameters()):
            if n in full:
                err = ((p.grad - full[n]).abs().max() / full[n].abs().max().clamp_min(1e-30)).item()
                worst = max(worst, err)
                assert err < 2e-5, (B, attach, n, err)
            else:
                assert p.grad is None, n                                 # after_rnn.*: stays None like the reference
        if attach:                                                       # the gradients really live in the arenas
            a = nem.__dict__["_vmm_grad_arena"].flat
            assert any(a.data_ptr() <= p.grad.data_ptr() < a.data_ptr() + a.numel() * 4 for p in nem.parameters() if p.grad is not None)
    if rank == 0:
        print(f"DIST_GPU_OK world={world} backend={backend} worst_rel_err={worst:.2e}")
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
