"""Import the UNMODIFIED reference modules from /root/reference on CPU.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Works only in the dev container
where ``/root/reference`` is mounted; raises ``ReferenceUnavailable`` elsewhere
(the GPU box has no reference tree - tests there use the committed goldens).

No reference code is copied: we only arrange for ``import train_nem`` to work
on a CPU-only box with a modern Python (recipe from SURVEY.md §8c):
  * stub ``tensorboardX`` / ``skimage`` (absent, tools/Trainer.py:9,
    tools/ImgDataset.py:6),
  * alias ``collections.Iterable`` (tools/Trainer.py:11, removed in py3.10),
  * make ``.cuda()`` the identity when no GPU is visible
    (models/MVCNN.py:10-11 call it at import; train_nem.py:217;
    tools/NEM_utilities.py:72).
"""
from __future__ import annotations

import collections
import collections.abc
import os
import sys
import types

REFERENCE_ROOT = os.environ.get("VMM_REFERENCE_ROOT", "/root/reference")


class ReferenceUnavailable(RuntimeError):
    pass


_cached = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "train_nem.py"))


def load():
    """Returns a namespace with NEM, Multi_View_Net, SVCNN, compute_outer_loss,
    compute_prior, bin_uniform, compute_norm - the reference's own objects."""
    global _cached
    if _cached is not None:
        return _cached
    if not available():
        raise ReferenceUnavailable(f"{REFERENCE_ROOT}/train_nem.py not found")
    import torch

    if "tensorboardX" not in sys.modules:
        m = types.ModuleType("tensorboardX")
        m.SummaryWriter = type("SummaryWriter", (), {"__init__": lambda self, *a, **k: None})
        sys.modules["tensorboardX"] = m
    for name in ("skimage", "skimage.io", "skimage.transform"):
        sys.modules.setdefault(name, types.ModuleType(name))
    if not hasattr(collections, "Iterable"):
        collections.Iterable = collections.abc.Iterable
    if not torch.cuda.is_available():
        torch.Tensor.cuda = lambda self, *a, **k: self
        torch.nn.Module.cuda = lambda self, *a, **k: self

    if REFERENCE_ROOT not in sys.path:
        sys.path.insert(0, REFERENCE_ROOT)
    argv = sys.argv
    sys.argv = [argv[0] if argv else "oracle"]
    try:
        import train_nem  # noqa: the reference's own module
        from tools import NEM_utilities
        from models.MVCNN import SVCNN
    finally:
        sys.argv = argv
    ns = types.SimpleNamespace(
        NEM=train_nem.NEM,
        Multi_View_Net=train_nem.Multi_View_Net,
        SVCNN=SVCNN,
        compute_outer_loss=NEM_utilities.compute_outer_loss,
        compute_prior=NEM_utilities.compute_prior,
        bin_uniform=NEM_utilities.bin_uniform,
        compute_norm=NEM_utilities.compute_norm,
        train_nem=train_nem,
    )
    _cached = ns
    return ns


def build_reference_models(k=3, view_num=12, input_dim=4096, rnn_hidden_size=1024,
                           iter_num=10, nclasses=40, e_sigma=0.1, init_type="bin_uniform",
                           gamma_prior_loss=1, stop_gamma_grad=0, if_sort=1, if_pooling=0,
                           save_feature=0):
    """Constructs the reference NEM and Multi_View_Net the way train_nem.py:408-434 does
    (pretraining=False: no network)."""
    ref = load()
    nem = ref.NEM(nclasses=nclasses, view_num=view_num, k=k, batch_size=1, input_dim=input_dim,
                  input_type="fc", rnn_hidden_size=rnn_hidden_size, iter_num=iter_num,
                  stop_gamma_grad=stop_gamma_grad, init_type=init_type,
                  gamma_prior_loss=gamma_prior_loss, e_sigma=e_sigma)
    cnet = ref.SVCNN("svcnn", nclasses=nclasses, pretraining=False, cnn_name="vggm")
    head = ref.Multi_View_Net(cnet, cnn_name="vggm", nclasses=nclasses, k=k,
                              rnn_hidden_size=rnn_hidden_size, if_sort=if_sort,
                              if_pooling=if_pooling, save_feature=save_feature)
    return nem, head


def reference_train_step(nem, head, x, labels, iter_num, loss_fn=None):
    """The reference's per-batch loop, tools/Trainer.py:160-203, driven with the
    reference's own modules.  x: (B,M,D) float32 on CPU.  Returns dict of tensors;
    `loss` carries the autograd graph."""
    import torch

    ref = load()
    if loss_fn is None:
        loss_fn = torch.nn.CrossEntropyLoss()
    init_state = nem.init_state(x.shape, x)
    state = (init_state[0], init_state[1], init_state[2])
    gamma_prior = init_state[3]
    prior = ref.compute_prior(nem.type)
    in_data = x.unsqueeze(0).expand(iter_num, -1, -1, -1).float()
    nem_losses, intra_losses, inter_losses = [], [], []
    gamma = gamma_prior
    for j in range(iter_num):
        inp = in_data[j].unsqueeze(1)
        state = nem(inp, state)
        theta, pred_f, gamma = state
        nem_loss, intra_loss, inter_loss = ref.compute_outer_loss(
            pred_f, gamma, inp, prior, nem.stop_gamma_grad, gamma_prior, nem.if_gamma_prior)
        nem_losses.append(nem_loss)
        intra_losses.append(intra_loss)
        inter_losses.append(inter_loss)
    out = head(state[0], gamma)
    res = dict(h=state[0], pred=state[1], gamma=gamma, nem_loss=nem_losses[-1],
               intra=intra_losses[-1], inter=inter_losses[-1], out=out)
    if labels is not None and head.save_fea == 0:
        closs = loss_fn(out, labels)
        res["closs"] = closs
        res["loss"] = closs + nem_losses[-1]
    return res
