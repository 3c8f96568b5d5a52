"""N = 2 on real GPUs: two CUDA ranks (torchrun) on the halves of a batch + GradBucket.allreduce_mean() reproduce the 1-rank
full-batch gradients of the whole training step (fp32 preset, 2e-5) - equal and ragged shards, with the gradients written
straight into the flat arenas (attach) and through the packing path.  With 2+ GPUs the ranks sit on different devices
and exchange over NCCL; on a single-GPU box both ranks run their CUDA step on cuda:0 and exchange the CUDA buffers over
gloo (NCCL refuses two ranks on one device) - the same GradBucket code path minus the transport."""
import os
import socket
import subprocess
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def test_two_cuda_ranks_equal_one_rank_full_batch():
    backend = "nccl" if torch.cuda.device_count() >= 2 else "gloo"
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), os.path.join(HERE, "dist_gpu_worker.py")]
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600,
                       env=dict(os.environ, VMM_TEST_BACKEND=backend))
    assert r.returncode == 0 and f"DIST_GPU_OK world=2 backend={backend}" in r.stdout, r.stdout[-3000:]


def test_grad_arena_single_rank():
    """Without a process group: attach() makes the backward write into the flat arenas and install views as .grad;
    the values equal the plain path's, and a second backward without zero_grad falls back to accumulation."""
    from multi_view_sort_b200 import NEM, Multi_View_Net
    from multi_view_sort_b200.dist import GradBucket
    from multi_view_sort_b200.functional import cross_entropy
    torch.manual_seed(10)
    nem = NEM(nclasses=10, view_num=12, k=2, input_dim=256, rnn_hidden_size=128, iter_num=2, init_type="bin_uniform",
              gamma_prior_loss=1, e_sigma=0.15).cuda().eval()
    head = Multi_View_Net(None, "vggm", nclasses=10, k=2, rnn_hidden_size=128).cuda().eval()
    x = torch.relu(torch.randn(4, 12, 256, generator=torch.Generator().manual_seed(1))).cuda()
    y = torch.randint(0, 10, (4,), generator=torch.Generator().manual_seed(2)).cuda()
    params = list(nem.parameters()) + list(head.parameters())

    def step():
        h, gamma, nem_loss, _, _ = nem.run_em(x, 2)
        (cross_entropy(head(h, gamma), y) + nem_loss).backward()

    step()
    plain = [None if p.grad is None else p.grad.clone() for p in params]
    for p in params:
        p.grad = None
    bucket = GradBucket(params).attach(nem, head)
    step()
    assert bucket.allreduce_mean() > 0                       # world 1: nothing to exchange, but the arena path is walked
    arena = nem.__dict__["_vmm_grad_arena"].flat
    for p, g in zip(params, plain):
        if g is None:
            assert p.grad is None
        else:                                                # two runs: equal up to the order of the reduce-adds
            assert ((p.grad - g).abs().max() / g.abs().max().clamp_min(1e-30)).item() < 1e-5
    inside = [p for p in nem.parameters() if p.grad is not None and arena.data_ptr() <= p.grad.data_ptr() < arena.data_ptr() + arena.numel() * 4]
    assert len(inside) == 22
    step()                                                   # .grad is set: plain accumulation, no aliasing
    for p, g in zip(params, plain):
        if g is not None:
            assert ((p.grad - 2 * g).abs().max() / g.abs().max().clamp_min(1e-30)).item() < 1e-4


def test_cli_trains_under_torchrun(tmp_path):
    """The reference's command line under torchrun on 2 GPUs: DistributedSampler shards, gradient arenas attached by the
    trainer, NCCL exchange, rank-0 logging / checkpoints, all-reduced validation counters.  Needs 2 GPUs (NCCL)."""
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import glob
    import json
    sys.path.insert(0, HERE)
    from test_features import make_store
    train_root = make_store(str(tmp_path / "train"), "MNet10", per_class=6, M=12, D=4096, seed=1)
    val_root = make_store(str(tmp_path / "val"), "MNet10", per_class=3, M=12, D=4096, seed=2)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", str(_free_port()), "-m", "multi_view_sort_b200.train_nem", "-name", "run2", "-dataset", "MNet10",
           "-cluster_n", "3", "-iter", "2", "-bs", "4", "-train_path", train_root, "-val_path", val_root, "-precision", "bf16",
           "-epoch", "2"]
    env = dict(os.environ, PYTHONPATH=os.path.dirname(HERE) + os.pathsep + os.environ.get("PYTHONPATH", ""))
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, timeout=600, cwd=str(tmp_path), env=env)
    assert r.returncode == 0, r.stdout[-3000:]
    run = tmp_path / "run2"
    assert json.load(open(run / "config.json"))["cluster_n"] == 3
    assert glob.glob(str(run / "NEM" / "model-*.pth")) and glob.glob(str(run / "MV_C_Net" / "model-*.pth"))
    scal = [json.loads(l) for l in open(run / "scalars.jsonl")]
    losses = [s["value"] for s in scal if s["tag"] == "train/train_loss"]
    assert len(losses) == 2 * 3 and all(v == v and abs(v) < 1e4 for v in losses)      # 24 objects / (2 ranks x 4) = 3 steps per epoch
    val = [s["value"] for s in scal if s["tag"] == "val/val_overall_acc"]
    assert len(val) == 2 and all(0.0 <= v <= 1.0 for v in val)
