"""world_size-2 gloo tests (CPU) of the N > 1 host logic: object sharding and the flat
gradient all-reduce.  The compute on each rank is the CPU oracle (the CUDA path needs a GPU),
which is exactly what proves the loss-scaling rule: the average of the rank gradients equals the
full-batch gradient, including ragged shards."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from multi_view_sort_b200.dist import GradBucket, loss_weight, shard_range


def test_shard_range_covers_everything():
    for n in (1, 7, 8, 4096, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                b, e = shard_range(n, r, world)
                assert 0 <= b <= e <= n
                seen += list(range(b, e))
            assert seen == list(range(n))
            sizes = [shard_range(n, r, world)[1] - shard_range(n, r, world)[0] for r in range(world)]
            assert max(sizes) - min(sizes) <= 1


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, B, out):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from oracle import nem_oracle as O
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    torch.set_num_threads(2)
    cfg = dict(K=2, M=12, D=64, H=64, T=2, sigma=0.2, nclasses=10)
    from multi_view_sort_b200 import NEM, Multi_View_Net
    torch.manual_seed(10)
    nem = NEM(nclasses=cfg["nclasses"], view_num=cfg["M"], k=cfg["K"], input_dim=cfg["D"], rnn_hidden_size=cfg["H"],
              iter_num=cfg["T"], init_type="bin_uniform", gamma_prior_loss=1, e_sigma=cfg["sigma"])
    head = Multi_View_Net(None, "vggm", nclasses=cfg["nclasses"], k=cfg["K"], rnn_hidden_size=cfg["H"])
    x, y = O.make_inputs(B, cfg["M"], cfg["D"], cfg["nclasses"], seed=10)
    b, e = shard_range(B, rank, world)

    def step(xs, ys, weight):
        for p in list(nem.parameters()) + list(head.parameters()):
            p.grad = None
        pn = {k: v for k, v in nem.named_parameters()}
        ph = {k: v for k, v in head.named_parameters()}
        r = O.train_step(pn, ph, xs, ys, cfg["K"], cfg["T"], cfg["sigma"], 1, 0, 1, 0)
        (r["loss"] * weight).backward()

    step(x[b:e], y[b:e], loss_weight(e - b, B, world))
    bucket = GradBucket(list(nem.parameters()) + list(head.parameters()))
    nbytes = bucket.allreduce_mean()
    sharded = {n: p.grad.clone() for n, p in list(nem.named_parameters()) + list(head.named_parameters()) if p.grad is not None}
    none_kept = [n for n, p in nem.named_parameters() if p.grad is None]
    step(x, y, 1.0)                                   # full batch on every rank
    worst = 0.0
    for n, p in list(nem.named_parameters()) + list(head.named_parameters()):
        if p.grad is None:
            continue
        worst = max(worst, ((sharded[n] - p.grad).abs().max() / p.grad.abs().max().clamp_min(1e-30)).item())
    if rank == 0:
        torch.save(dict(worst=worst, nbytes=nbytes, none_kept=none_kept), out)
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("B", [8, 7])
def test_two_rank_sharded_step_equals_full_batch(tmp_path, B):
    out = str(tmp_path / "res.pt")
    mp.spawn(_worker, args=(2, _free_port(), B, out), nprocs=2, join=True)
    res = torch.load(out)
    assert res["worst"] < 2e-5, res          # fp32 re-association only
    assert res["nbytes"] > 0
    assert sorted(res["none_kept"]) == ["after_rnn.0.bias", "after_rnn.0.weight"]   # stay None, never zero-filled
