"""GPU test of the trainer / CLI / descriptor export on a tiny synthetic feature store in the reference's layout."""
import glob
import json
import os

import numpy as np
import pytest
import torch

from test_features import make_store

pytestmark = pytest.mark.gpu


def test_train_then_export(tmp_path, monkeypatch):
    from multi_view_sort_b200 import train_nem as cli
    train_root = make_store(str(tmp_path / "train"), "MNet10", per_class=4, M=12, D=4096, seed=1)
    val_root = make_store(str(tmp_path / "val"), "MNet10", per_class=2, M=12, D=4096, seed=2)
    monkeypatch.chdir(tmp_path)
    common = ["-name", "run", "-dataset", "MNet10", "-cluster_n", "3", "-iter", "2", "-bs", "8", "-train_path", train_root,
              "-val_path", val_root, "-precision", "bf16"]
    cli.main(common + ["-epoch", "2"])
    assert json.load(open("run/config.json"))["cluster_n"] == 3
    assert glob.glob("run/NEM/model-*.pth") and glob.glob("run/MV_C_Net/model-*.pth")     # best-accuracy checkpoint
    scal = [json.loads(l) for l in open("run/scalars.jsonl")] if os.path.exists("run/scalars.jsonl") else None
    if scal is not None:
        tags = {s["tag"] for s in scal}
        assert {"params/lr", "train/train_loss", "train/c_loss", "train/nem_loss", "train/intra_loss", "train/inter_loss",
                "train/train_overall_acc", "val/val_mean_class_acc", "val/val_overall_acc", "val/val_loss"} <= tags
        assert all(np.isfinite(s["value"]) for s in scal)
    assert os.path.exists("3_each_class_acc.npy")
    # test mode: load the checkpoint and export one 4096-d descriptor per object
    cli.main(common + ["-train_or_test_mode", "test", "-feature_out", str(tmp_path / "desc")])
    files = sorted(glob.glob(str(tmp_path / "desc" / "*" / "*" / "*.npy")))
    assert len(files) == 8
    d0 = np.load(files[0])
    assert d0.shape == (1, 4096) and np.isfinite(d0).all()
    # the exported descriptor equals a direct call on the same object
    from multi_view_sort_b200.train_nem import build_models, parser
    args = parser.parse_args(common)
    nem, head = build_models(args, 10, save_feature=1)
    nem.load("run", None); head.load("run", None)
    nem, head = nem.cuda().eval(), head.cuda().eval()
    rel = os.path.join(*files[0].split("/")[-3:])
    x = torch.from_numpy(np.load(os.path.join(val_root, rel))).cuda()
    with torch.no_grad():
        h, gamma, *_ = nem.run_em(x)
        d1 = head(h, gamma).cpu().numpy()
    assert np.abs(d1 - d0).max() <= 2e-3 * np.abs(d0).max()       # bf16 preset, batch 1 vs batch 8 reduction order


def test_device_prefetcher_order_and_values():
    """DevicePrefetcher yields the loader's batches in order, on the device, bit-identical, passing non-tensors through."""
    from multi_view_sort_b200.features import DevicePrefetcher
    g = torch.Generator().manual_seed(3)
    batches = [(torch.randint(0, 10, (4,), generator=g), torch.randn(4, 12, 64, generator=g).pin_memory(), [f"p{i}"] * 4)
               for i in range(5)]
    seen = 0
    for i, (y, x, paths) in enumerate(DevicePrefetcher(batches)):
        assert x.is_cuda and y.is_cuda and paths == batches[i][2]
        x = x * 1.0                                  # consume on the compute stream
        assert torch.equal(x.cpu(), batches[i][1]) and torch.equal(y.cpu(), batches[i][0])
        seen += 1
    assert seen == 5 and len(DevicePrefetcher(batches)) == 5


def test_training_updates_the_prepared_weights_and_learns():
    """The operand planes of the weights are cached per parameter version: after every optimizer step (fused Adam
    updates in place) they must be rebuilt, and a few steps on one fixed batch must reduce the loss."""
    from multi_view_sort_b200 import NEM, Multi_View_Net
    from multi_view_sort_b200.functional import cross_entropy
    torch.manual_seed(0)
    nem = NEM(nclasses=10, view_num=12, k=3, input_dim=256, rnn_hidden_size=1024, iter_num=2, init_type="bin_uniform",
              gamma_prior_loss=1, e_sigma=0.1, precision="bf16").cuda().eval()
    head = Multi_View_Net(None, "vggm", nclasses=10, k=3, rnn_hidden_size=1024, precision="bf16").cuda().eval()
    opt = torch.optim.Adam(list(nem.parameters()) + list(head.parameters()), lr=2e-4, fused=True)
    g = torch.Generator().manual_seed(1)
    x = torch.relu(torch.randn(16, 12, 256, generator=g)).cuda().bfloat16()
    y = torch.randint(0, 10, (16,), generator=g).cuda()
    def eval_out():
        with torch.no_grad():
            h, gamma, *_ = nem.run_em(x, 2)
            return h.clone(), head(h, gamma).clone()
    losses, prev = [], eval_out()
    for step in range(12):
        opt.zero_grad(set_to_none=True)
        h, gamma, nem_loss, _, _ = nem.run_em(x, 2)
        closs = cross_entropy(head(h, gamma), y)
        (closs + nem_loss).backward()
        opt.step()
        losses.append(closs.item())
        if step < 3:                                  # a forward-only call right after the step sees the NEW weights
            cur = eval_out()
            assert not torch.equal(cur[0], prev[0]) and not torch.equal(cur[1], prev[1]), "stale weight planes"
            prev = cur
    assert all(torch.isfinite(torch.tensor(losses)))
    assert losses[-1] < 0.7 * losses[0], losses
