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


def test_trainer_trajectory_matches_oracle_adam(tmp_path):
    """f1 as a trajectory: ModelNetTrainer.train_nem_mvcnn (fp32 preset, eval-mode modules so that dropout is off)
    on a tiny in-memory loader for 3 optimizer steps against the oracle's train_step + torch.optim.Adam on the CPU
    with the same weights, data and hyper-parameters: the logged losses within 1e-4 (first step) / 1e-3 (later steps), the parameters after
    the last step close to the oracle's in relative L2 of the update.  Also pins the
    -if_total_loss 1 scalar (closs + mean of the T NEM losses, Trainer.py:203)."""
    from oracle import nem_oracle as O
    from helpers import params_of
    from multi_view_sort_b200 import NEM, Multi_View_Net
    from multi_view_sort_b200.trainer import ModelNetTrainer
    import torch.nn as nn
    cfg = dict(B=6, K=2, M=12, D=512, H=256, T=3, sigma=0.1, nclasses=10)
    for total_loss in (0, 1):
        torch.manual_seed(10)
        nem = NEM(nclasses=cfg["nclasses"], view_num=cfg["M"], k=cfg["K"], input_dim=cfg["D"], rnn_hidden_size=cfg["H"],
                  iter_num=cfg["T"], init_type="bin_uniform", gamma_prior_loss=1, e_sigma=cfg["sigma"], if_total_loss=total_loss,
                  precision="fp32")
        head = Multi_View_Net(None, "vggm", nclasses=cfg["nclasses"], k=cfg["K"], rnn_hidden_size=cfg["H"], precision="fp32")
        pn, ph = params_of(nem), params_of(head)
        start = {k: v.detach().clone() for k, v in list(pn.items()) + [("head." + k, v) for k, v in ph.items()]}
        batches = []
        for i in range(3):
            x, y = O.make_inputs(cfg["B"], cfg["M"], cfg["D"], cfg["nclasses"], seed=20 + i, scale=0.5)
            batches.append((y, x, ["obj"] * cfg["B"]))
        # ---- oracle trajectory (CPU, torch autograd + Adam) ----
        opt_ref = torch.optim.Adam(list(pn.values()) + list(ph.values()), lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999))
        ref_losses = []
        for y, x, _ in batches:
            opt_ref.zero_grad()
            losses = []
            r = O.run_em(pn, x, cfg["K"], cfg["T"], cfg["sigma"], "bin_uniform", 1, 0, losses=losses)
            out = O.head_forward(ph, r["h"], r["gamma"], cfg["K"], 1, 0)
            closs = torch.nn.functional.cross_entropy(out, y)
            loss = closs + (r["nem_loss"] if total_loss == 0 else torch.stack([l[0] for l in losses]).detach().mean())
            loss.backward()
            opt_ref.step()
            ref_losses.append((loss.item(), closs.item(), r["nem_loss"].item()))
        # ---- this package's trainer ----
        # the trainer calls .train(); parity is defined with dropout off
        nem.train = lambda mode=True: nem
        head.train = lambda mode=True: head
        nn.Module.eval(nem); nn.Module.eval(head)
        nem.training = head.training = False
        opt = torch.optim.Adam(list(nem.parameters()) + list(head.parameters()), lr=1e-3, weight_decay=1e-4, betas=(0.9, 0.999))
        log_dir = str(tmp_path / f"traj{total_loss}")
        os.makedirs(log_dir)
        trainer = ModelNetTrainer((nem, head), batches, [], opt, nn.CrossEntropyLoss(), None, log_dir, num_views=cfg["M"],
                                  class_num=cfg["nclasses"])
        cwd = os.getcwd()
        os.chdir(tmp_path)
        try:
            trainer.train_nem_mvcnn(1)
        finally:
            os.chdir(cwd)
        sc = trainer.writer.scalars if trainer.writer else None
        assert sc is not None
        got = list(zip([v[2] for v in sc["train/train_loss"]], [v[2] for v in sc["train/c_loss"]], [v[2] for v in sc["train/nem_loss"]]))
        assert len(got) == 3
        for step, (a, b) in enumerate(zip(got, ref_losses)):
            # step 0 is a pure forward pass: 1e-4.  From the second step on the parameters differ by what Adam makes of
            # gradient entries within rounding of zero (a +-lr step either way, see below), which the losses pick up:
            # measured 1.3e-4 on the NEM loss of the third step at lr = 1e-3; 1e-3 still pins the trajectory
            tol = 1e-4 if step == 0 else 1e-3
            for j in range(3):
                assert abs(a[j] - b[j]) < tol * abs(b[j]), (total_loss, step, j, a, b)
        mine = {k: v.detach().cpu() for k, v in nem.state_dict().items()}
        mine.update({"head." + k: v.detach().cpu() for k, v in head.state_dict().items()})
        ref_end = dict(pn)
        ref_end.update({"head." + k: v for k, v in ph.items()})
        for k, v in ref_end.items():
            moved = (v.detach() - start[k]).norm().item()
            if moved == 0:
                continue
            # Adam's first updates are ~lr * sign(g): an entry whose gradient is within rounding of zero may move the
            # other way, so the parameters are compared in relative L2 of the update, not entry-wise
            err = (mine[k] - v.detach()).norm().item() / moved
            assert err < 5e-2, (total_loss, k, err)
