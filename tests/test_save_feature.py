"""CPU tests of the stage-1 hand-off: the fc7 tap of a VGG-style stage-1 model, and the writer
whose files the reference's own dataset class reads back bit-for-bit."""
import os

import numpy as np
import pytest
import torch
import torch.nn as nn

from oracle import ref_import
from multi_view_sort_b200.features import CLASSNAMES, FC_SUBDIR, KmeanImgDataset, PackedFeatures, pack_shard
from multi_view_sort_b200.save_feature import Fc7Extractor, FeatureWriter, dump_features, object_name

V, C, HW, FEAT = 4, 3, 8, 32


class TinyStage1(nn.Module):
    """Same shape as the reference's SVCNN with a VGG backbone (models/MVCNN.py:55-62): a trunk and the torchvision
    classifier layout '0'..'6'."""

    def __init__(self, nclasses=10):
        super().__init__()
        self.net_1 = nn.Sequential(nn.Conv2d(C, 4, 3, padding=1), nn.ReLU(), nn.MaxPool2d(2))
        self.net_2 = nn.Sequential(nn.Linear(4 * (HW // 2) ** 2, FEAT), nn.ReLU(True), nn.Dropout(), nn.Linear(FEAT, FEAT),
                                   nn.ReLU(True), nn.Dropout(), nn.Linear(FEAT, nclasses))


def fake_loader(n_objects, bs, dataset="MNet10", seed=0):
    """Items like MultiviewImgDataset + default collation: (class_id[N], images[N, V, C, H, W], V tuples of N paths)."""
    g = torch.Generator().manual_seed(seed)
    names = CLASSNAMES[dataset]
    objs = [(i % 4, f"/data/{names[i % 4]}/test/{names[i % 4]}_{i:04d}") for i in range(n_objects)]
    for lo in range(0, n_objects, bs):
        part = objs[lo:lo + bs]
        ids = torch.tensor([c for c, _ in part])
        imgs = torch.randn(len(part), V, C, HW, HW, generator=g)
        paths = [tuple(f"{stem}.{v + 1:03d}.png" for _, stem in part) for v in range(V)]
        yield ids, imgs, paths


def test_taps_match_manual_computation():
    torch.manual_seed(0)
    m = TinyStage1().eval()
    x = torch.randn(2 * V, C, HW, HW)
    with torch.no_grad():
        trunk = m.net_1(x)
        flat = trunk.view(trunk.shape[0], -1)
        fc6 = torch.relu(m.net_2[0](flat))
        pre7 = m.net_2[3](fc6)
    assert torch.equal(Fc7Extractor(m, V, "fc6").eval()(x), fc6.view(2, V, FEAT))
    assert torch.equal(Fc7Extractor(m, V, "fc7_pre").eval()(x), pre7.view(2, V, FEAT))
    f7 = Fc7Extractor(m, V, "fc7").eval()(x)
    assert torch.equal(f7, torch.relu(pre7).view(2, V, FEAT)) and (f7 >= 0).all()
    # 'conv' = the reference's MVCNN.get_features (models/MVCNN.py:98-102): the trunk's maps regrouped per object
    assert torch.equal(Fc7Extractor(m, V, "conv").eval()(x), trunk.view(2, V, *trunk.shape[1:]))
    with pytest.raises(KeyError):
        Fc7Extractor(m, V, "fc8")
    with pytest.raises(KeyError):
        Fc7Extractor(nn.Linear(2, 2), V)


def test_object_name_follows_the_reference():
    assert object_name("/a/b/airplane/test/airplane_0627.001.png") == "airplane_0627"


@pytest.mark.parametrize("packed_dtype", ["float32", "bfloat16"])
def test_dump_is_what_the_feature_store_reads(tmp_path, packed_dtype):
    torch.manual_seed(1)
    m = TinyStage1().eval()
    ext = Fc7Extractor(m, V, "fc7").eval()
    root = str(tmp_path / "store")
    w = FeatureWriter(root, dataset="MNet10", fea_type="fc", packed_prefix=str(tmp_path / "shard"), packed_dtype=packed_dtype,
                      expect=10 if packed_dtype == "float32" else None)          # both packing paths (memmap / in memory)
    n = dump_features(ext, fake_loader(10, 3), w)
    assert n == 10 and w.close() is not None
    # the per-object files: read back by this package's dataset class and, in the dev container, by the reference's own
    ds = KmeanImgDataset(root + "/", fea_type="fc", dataset="MNet10", shuffle=False)
    assert len(ds) == 10
    want = {}
    for ids, imgs, paths in fake_loader(10, 3):
        f = ext(imgs.reshape(-1, C, HW, HW))
        for i in range(len(ids)):
            want[object_name(paths[0][i])] = (int(ids[i]), f[i])
    for cid, im, path in ds:
        assert path.split("/")[-2] == FC_SUBDIR and np.load(path).shape == (1, V, FEAT)
        c, f = want[object_name(path)]
        assert cid == c and torch.equal(im, f)
    if ref_import.available():
        ref = ref_import.load().train_nem.KmeanImgDataset(root + "/", fea_type="fc", dataset="MNet10", shuffle=False)
        assert ref.filepaths == ds.filepaths
        for a, b in zip(ds, ref):
            assert a[0] == b[0] and torch.equal(a[1], b[1])
    # the packed shard written in the same pass == features.pack_shard of the same objects (in write order)
    pk = PackedFeatures(str(tmp_path / "shard"), pin=False)
    assert len(pk) == 10 and list(pk.paths) == w.paths
    y, x = pk.batch(0, 10)
    for j, path in enumerate(w.paths):
        c, f = want[object_name(path)]
        assert y[j].item() == c
        assert torch.equal(x[j], f if packed_dtype == "float32" else f.bfloat16())
    ref_prefix = pack_shard(ds, str(tmp_path / "shard2"), packed_dtype)
    a, b = np.load(ref_prefix + ".features.npy"), np.load(str(tmp_path / "shard") + ".features.npy")
    order = [w.paths.index(p) for p in ds.filepaths]
    assert a.dtype == b.dtype and np.array_equal(a, b[order])


def test_conv_tap_against_the_reference_model():
    """In the dev container: Fc7Extractor(tap='conv') on the reference's MVCNN == its own get_features; the fc7 tap runs
    on its SVCNN (VGG classifier, children '0'..'6')."""
    if not ref_import.available():
        pytest.skip("reference tree not mounted")
    ref_import.load()
    from models.MVCNN import MVCNN
    torch.manual_seed(2)
    sv = ref_import.load().SVCNN("s", nclasses=40, pretraining=False, cnn_name="vggm")
    sv.net_1 = nn.Sequential(nn.Conv2d(3, 512, 3, stride=32), nn.AdaptiveAvgPool2d(7))      # stand-in trunk with VGG's 512x7x7 output
    mv = MVCNN("m", sv, nclasses=40, cnn_name="vggm", num_views=V).eval()
    x = torch.randn(2 * V, 3, 64, 64)
    with torch.no_grad():
        want = mv.get_features(x)
    assert torch.equal(Fc7Extractor(mv, V, "conv").eval()(x), want)
    f7 = Fc7Extractor(sv, V, "fc7").eval()(x)
    assert f7.shape == (2, V, 4096) and (f7 >= 0).all()
    with torch.no_grad():
        flat = want.reshape(2 * V, -1)
        manual = torch.relu(sv.net_2[3](torch.relu(sv.net_2[0](flat))))
    assert torch.equal(f7, manual.view(2, V, 4096))
