"""Stage-1 hand-off: dump the fc7 features of a trained single-view CNN into the feature store the
VMM path reads.

The reference describes this step (README.md:27-37: ``python mvcnn_save_feature.py -name xxx -cnn_name vggm``) but the
script it ships (mvcnn_save_feature.py:89-108) dumps the CONV maps ``(N, V, 512, 7, 7)`` through
``MVCNN.get_features`` (models/MVCNN.py:98-102), one object per batch, into a hard-coded directory; the fc7 dumper that
produced ``fc_features_vggm_pretrained_imagenet/`` (train_nem.py:103-104) is not in the repository.  This module is that
missing piece, written against the two interfaces on either side of it:

* upstream, the reference's stage-1 model: any module with ``net_1`` (convolutional trunk) and ``net_2`` (the VGG
  classifier ``Sequential`` whose children ``'0' .. '6'`` are Linear, ReLU, Dropout, Linear, ReLU, Dropout, Linear -
  models/MVCNN.py:55-62; ``Multi_View_Net`` borrows its ``'3'`` and ``'6'``, train_nem.py:309-324), fed by a loader that
  yields ``(class_id, images[N, V, C, H, W], paths)`` like ``MultiviewImgDataset`` (tools/ImgDataset.py:71-90);
* downstream, ``features.KmeanImgDataset`` (train_nem.py:77-141): one ``.npy`` per object holding ``(1, V, 4096)``
  float32 under ``<root>/<class>/fc_features_vggm_pretrained_imagenet/<object>.npy``.

The CNN itself stays torch code (image models are out of scope); what is new is batching (the reference
writes one object per forward pass), the choice of tap, and writing the packed shard (features.PackedFeatures) in the
same pass so that the trainer never has to open 12k small files.
"""
from __future__ import annotations

import json
import os
from typing import Callable, Iterable, List, Optional, Sequence

import numpy as np
import torch

from .features import CLASSNAMES, CONV_SUBDIR, FC_SUBDIR, _to_bf16_bits

# children of net_2 that make up each tap (models/MVCNN.py:55-62: torchvision's VGG classifier)
TAPS = {"fc6": ("0", "1"), "fc7": ("0", "1", "2", "3", "4"), "fc7_pre": ("0", "1", "2", "3")}


class Fc7Extractor(torch.nn.Module):
    """images ``(N*V, C, H, W)`` -> ``(N, V, D)``: ``net_1``, flatten, then ``net_2`` up to the tap.  ``tap``: 'fc7'
    (after the second Linear's ReLU: non-negative 4096-d features, what the VMM is trained on), 'fc7_pre' (before that
    ReLU), 'fc6', or 'conv' (the reference's own ``get_features``: the trunk's maps, unflattened)."""

    def __init__(self, model, num_views: int = 12, tap: str = "fc7"):
        super().__init__()
        if tap != "conv" and tap not in TAPS:
            raise KeyError(f"unknown tap {tap!r} (choose from {sorted(TAPS) + ['conv']})")
        if not hasattr(model, "net_1") or (tap != "conv" and not hasattr(model, "net_2")):
            raise KeyError("the stage-1 model must expose net_1 and net_2 (SVCNN / MVCNN with a VGG-style backbone)")
        self.net_1 = model.net_1
        self.tap, self.num_views = tap, num_views
        if tap != "conv":
            children = dict(model.net_2.named_children())
            missing = [k for k in TAPS[tap] if k not in children]
            if missing:
                raise KeyError(f"net_2 has no children {missing}: not a VGG-style classifier")
            self.head = torch.nn.Sequential(*[children[k] for k in TAPS[tap]])

    @torch.no_grad()
    def forward(self, images: torch.Tensor) -> torch.Tensor:
        assert images.dim() == 4 and images.shape[0] % self.num_views == 0, "images must be (N*V, C, H, W)"
        y = self.net_1(images)
        n = images.shape[0] // self.num_views
        if self.tap == "conv":
            return y.view(n, self.num_views, *y.shape[1:])
        return self.head(y.reshape(y.shape[0], -1)).view(n, self.num_views, -1)


def object_name(path: str) -> str:
    """'.../airplane/test/airplane_0627.001.png' -> 'airplane_0627' (mvcnn_save_feature.py:99-100)."""
    return path.split('/')[-1].split('.')[0]


def _first_view_paths(paths, n: int) -> List[str]:
    """The loader's third item after default collation: V sequences of N paths (data[2][view][object],
    mvcnn_save_feature.py:99) - or already a flat list of N first-view paths."""
    if len(paths) > 0 and isinstance(paths[0], (list, tuple)):
        paths = paths[0]
    assert len(paths) == n, f"{len(paths)} paths for {n} objects"
    return [str(p) for p in paths]


class FeatureWriter:
    """Writes objects in the layout ``KmeanImgDataset`` reads, and optionally a packed shard in the same pass.

    ``write(class_ids, feats, names)``: feats ``(N, V, D)`` (any float dtype; stored as float32 ``(1, V, D)`` per
    object, like the reference's files).  ``close()`` finishes the packed shard (``<prefix>.features.npy`` (N, V, D)
    float32 or bf16 bit patterns, ``.labels.npy``, ``.json``)."""

    def __init__(self, out_root: str, dataset: str = "MNet40", fea_type: Optional[str] = "fc",
                 packed_prefix: Optional[str] = None, packed_dtype: str = "float32", expect: Optional[int] = None):
        if dataset not in CLASSNAMES:
            raise KeyError(f"unknown dataset {dataset!r}")
        self.root, self.dataset, self.classnames = out_root, dataset, CLASSNAMES[dataset]
        self.sub = FC_SUBDIR if fea_type == "fc" else CONV_SUBDIR
        self.packed_prefix, self.packed_dtype, self.expect = packed_prefix, packed_dtype, expect
        self.paths: List[str] = []
        self.labels: List[int] = []
        self._mm = None
        self._chunks: List[np.ndarray] = []

    def _pack(self, a: np.ndarray):
        if self.packed_prefix is None:
            return
        b = _to_bf16_bits(a) if self.packed_dtype == "bfloat16" else a.astype(np.float32)
        if self.expect is not None:
            if self._mm is None:
                self._mm = np.lib.format.open_memmap(self.packed_prefix + ".features.npy", mode="w+", dtype=b.dtype,
                                                     shape=(self.expect,) + a.shape[1:])
            lo = len(self.labels) - a.shape[0]
            self._mm[lo:lo + a.shape[0]] = b
        else:
            self._chunks.append(b)

    def write(self, class_ids, feats: torch.Tensor, names: Sequence[str]) -> List[str]:
        a = feats.detach().to(torch.float32).cpu().numpy()
        ids = [int(c) for c in (class_ids.tolist() if torch.is_tensor(class_ids) else class_ids)]
        assert a.shape[0] == len(ids) == len(names)
        out = []
        for i, (cid, name) in enumerate(zip(ids, names)):
            d = os.path.join(self.root, self.classnames[cid], self.sub)
            os.makedirs(d, exist_ok=True)
            path = os.path.join(d, name + ".npy")
            np.save(path, a[i:i + 1])                        # (1, V, D): the reference's per-object array
            out.append(path)
        self.paths.extend(out)
        self.labels.extend(ids)
        if a.ndim == 3:
            self._pack(a)
        return out

    def close(self) -> Optional[str]:
        if self.packed_prefix is None or not self.labels:
            return None
        if self._mm is not None:
            assert len(self.labels) == self.expect, f"wrote {len(self.labels)} objects, expected {self.expect}"
            self._mm.flush()
            shape = self._mm.shape
        else:
            feats = np.concatenate(self._chunks, axis=0)
            np.save(self.packed_prefix + ".features.npy", feats)
            shape = feats.shape
        np.save(self.packed_prefix + ".labels.npy", np.asarray(self.labels, dtype=np.int64))
        with open(self.packed_prefix + ".json", "w") as f:
            json.dump({"dtype": self.packed_dtype, "n": int(shape[0]), "views": int(shape[1]), "dim": int(shape[2]),
                       "dataset": self.dataset, "paths": self.paths}, f)
        return self.packed_prefix


def dump_features(extractor: Callable[[torch.Tensor], torch.Tensor], loader: Iterable, writer: FeatureWriter,
                  device: Optional[torch.device] = None, log_every: int = 0) -> int:
    """Runs ``extractor`` over ``loader`` (items ``(class_id[N], images[N, V, C, H, W], paths)``) and writes every object.
    Any batch size (the reference's script assumes 1: ``classnames[data[0]]``, mvcnn_save_feature.py:94).  Returns the
    number of objects written."""
    seen = 0
    for step, data in enumerate(loader):
        class_ids, images, paths = data[0], data[1], data[2]
        n, v = images.shape[0], images.shape[1]
        x = images.reshape(n * v, *images.shape[2:])
        if device is not None:
            x = x.to(device, non_blocking=True)
        feats = extractor(x)
        names = [object_name(p) for p in _first_view_paths(paths, n)]
        writer.write(class_ids, feats, names)
        seen += n
        if log_every and (step + 1) % log_every == 0:
            print(f"save_feature: {seen} objects")
    return seen


def main(argv=None):
    """``python -m multi_view_sort_b200.save_feature -weights <state_dict.pth> -val_path <.../*/test> -out_dir <root>``:
    needs the reference's stage-1 code importable (``models.MVCNN.SVCNN``, ``tools.ImgDataset.MultiviewImgDataset`` -
    image models and the image reader are not part of this package) and a fine-tuned VGG-M checkpoint."""
    import argparse
    ap = argparse.ArgumentParser(description="fc7 feature dump for the VMM feature store")
    ap.add_argument("-name", default="MVCNN")
    ap.add_argument("-cnn_name", default="vggm")
    ap.add_argument("-num_views", type=int, default=12)
    ap.add_argument("-bs", type=int, default=8, help="objects per forward pass")
    ap.add_argument("-weights", required=True, help="checkpoint of the stage-1 model (Model.save format)")
    ap.add_argument("-val_path", required=True, help=".../*/train or .../*/test, as in the reference")
    ap.add_argument("-out_dir", required=True)
    ap.add_argument("-tap", default="fc7", choices=sorted(TAPS) + ["conv"])
    ap.add_argument("-packed", default=None, help="also write a packed shard with this prefix")
    ap.add_argument("-packed_dtype", default="float32", choices=["float32", "bfloat16"])
    args = ap.parse_args(argv)
    try:
        from models.MVCNN import SVCNN                          # the reference's stage-1 code, on PYTHONPATH
        from tools.ImgDataset import MultiviewImgDataset
    except ImportError as e:
        raise SystemExit(f"save_feature needs the reference's stage-1 modules on PYTHONPATH ({e})")
    dev = torch.device("cuda" if torch.cuda.is_available() else "cpu")
    cnet = SVCNN(args.name, nclasses=40, pretraining=False, cnn_name=args.cnn_name)
    cnet.load_state_dict(torch.load(args.weights, map_location="cpu"))
    ext = Fc7Extractor(cnet, args.num_views, args.tap).to(dev).eval()
    ds = MultiviewImgDataset(args.val_path, scale_aug=False, rot_aug=False, num_views=args.num_views, test_mode=True,
                             shuffle=False)
    loader = torch.utils.data.DataLoader(ds, batch_size=args.bs, shuffle=False, num_workers=0)
    w = FeatureWriter(args.out_dir, fea_type="fc" if args.tap != "conv" else None, packed_prefix=args.packed,
                      packed_dtype=args.packed_dtype, expect=len(ds) if args.packed else None)
    n = dump_features(ext, loader, w, dev, log_every=50)
    w.close()
    print(f"save_feature: wrote {n} objects under {args.out_dir}")


if __name__ == "__main__":
    main()
