"""CPU tests of the 'next' rows either side of the hot path: the feature-store reader (the reference's
on-disk format), the packed shard format, and the CLI flag surface (f1)."""
import os

import numpy as np
import pytest
import torch

from oracle import ref_import
from multi_view_sort_b200.features import CLASSNAMES, FC_SUBDIR, KmeanImgDataset, PackedFeatures, pack_shard


def make_store(root, dataset="MNet10", per_class=3, M=12, D=64, seed=0):
    rng = np.random.RandomState(seed)
    for c in CLASSNAMES[dataset][:4]:
        d = os.path.join(root, c, FC_SUBDIR)
        os.makedirs(d)
        for i in range(per_class):
            np.save(os.path.join(d, f"{c}_{i:04d}.npy"), np.maximum(rng.randn(1, M, D), 0).astype(np.float32))
    return root + "/"


def test_dataset_reads_reference_layout(tmp_path):
    root = make_store(str(tmp_path))
    ds = KmeanImgDataset(root, fea_type="fc", dataset="MNet10", shuffle=False)
    assert len(ds) == 12
    cid, im, path = ds[5]
    assert im.shape == (12, 64) and im.dtype == torch.float32
    assert CLASSNAMES["MNet10"][cid] == path.split("/")[-3]
    if ref_import.available():
        ref = ref_import.load().train_nem.KmeanImgDataset(root, fea_type="fc", dataset="MNet10", shuffle=False)
        assert ref.filepaths == ds.filepaths
        for i in range(len(ds)):
            a, b = ds[i], ref[i]
            assert a[0] == b[0] and a[2] == b[2] and torch.equal(a[1], b[1])
        np.random.seed(10)
        s1 = KmeanImgDataset(root, fea_type="fc", dataset="MNet10").filepaths
        np.random.seed(10)
        s2 = ref_import.load().train_nem.KmeanImgDataset(root, fea_type="fc", dataset="MNet10").filepaths
        assert s1 == s2                                    # same shuffle under the same numpy seed


@pytest.mark.parametrize("dtype", ["float32", "bfloat16"])
def test_packed_shard_roundtrip(tmp_path, dtype):
    root = make_store(str(tmp_path / "store"))
    ds = KmeanImgDataset(root, fea_type="fc", dataset="MNet10", shuffle=False)
    prefix = pack_shard(ds, str(tmp_path / "shard"), dtype)
    pk = PackedFeatures(prefix, pin=False)
    assert len(pk) == len(ds)
    y, x = pk.batch(2, 9)
    assert x.shape == (7, 12, 64) and y.shape == (7,)
    for j, i in enumerate(range(2, 9)):
        cid, im, _ = ds[i]
        assert y[j].item() == cid
        if dtype == "float32":
            assert torch.equal(x[j], im)
        else:
            assert x.dtype == torch.bfloat16 and torch.equal(x[j], im.bfloat16())   # RNE, same as torch's cast
    seen = sum(xb.shape[0] for _, xb in pk.iterate(5, shuffle=True, seed=1))
    assert seen == len(ds)


def test_cli_flags_match_reference():
    from multi_view_sort_b200 import train_nem as mine
    opts = {a.option_strings[0]: a.default for a in mine.parser._actions if a.option_strings and a.option_strings[0] != "-h"}
    for flag, default in [("-cluster_n", 4), ("-iter_num", 15), ("-bs", 32), ("-lr", 5e-5), ("-e_sigma", 0.1),
                          ("-init_type", "bin_uniform"), ("-if_gamma_prior", 1), ("-if_sort", 1), ("-epoch", 70)]:
        assert opts[flag] == default, flag
    ns = mine.parser.parse_args(["-name", "x", "-cluster_n", "3", "-iter", "10"])      # README's abbreviated form
    assert ns.cluster_n == 3 and ns.iter_num == 10
    if ref_import.available():
        ref = ref_import.load().train_nem.parser
        ref_opts = {a.option_strings[0]: a.default for a in ref._actions if a.option_strings and a.option_strings[0] != "-h"}
        for k, v in ref_opts.items():
            assert k in opts, f"reference flag {k} missing"
            if k in ("-train_path", "-val_path"):          # the reference defaults to its authors' private mounts; here: required
                assert opts[k] == ""
                continue
            assert opts[k] == v, (k, opts[k], v)
        ref_types = {a.option_strings[0]: a.type for a in ref._actions if a.option_strings and a.option_strings[0] != "-h"}
        my_types = {a.option_strings[0]: a.type for a in mine.parser._actions if a.option_strings and a.option_strings[0] != "-h"}
        for k, t in ref_types.items():
            assert my_types[k] == t, (k, my_types[k], t)


def test_precision_rules_of_the_bf16_preset():
    """Host logic of functional.resolve_precision: the sigma rule picks single-pass dec3 for training and
    all-single-pass contractions for forward-only calls iff sigma >= 0.07; other presets are never touched."""
    from types import SimpleNamespace as NS
    from multi_view_sort_b200.functional import (DEC3_SINGLE_MIN_SIGMA, FWD_CONTRACTIONS, FWD_SINGLE, PRECISIONS, fwd_modes,
                                                 precision_planes, resolve_precision)
    assert precision_planes((1, 0, 1, 0)) == (1, 0, 1, 0, 0, 0, 0)
    assert precision_planes("fp32") == PRECISIONS["fp32"] + (0,)
    dec3 = fwd_modes(dec3=FWD_SINGLE)
    every = fwd_modes(**{n: FWD_SINGLE for n in FWD_CONTRACTIONS})
    assert dec3 == 1 << 14 and every == 0x15555
    hi, lo = NS(precision="bf16", e_sigma=0.1), NS(precision="bf16", e_sigma=0.02)
    assert resolve_precision(hi, True)[4] == dec3 and resolve_precision(hi, False)[4] == every
    assert resolve_precision(lo, True)[4] == 0 and resolve_precision(lo, False)[4] == 0
    assert resolve_precision(NS(precision="bf16", e_sigma=DEC3_SINGLE_MIN_SIGMA), True)[4] == dec3
    for name in ("fp32", "fp32_strict", "bf16_fast"):
        m = NS(precision=name, e_sigma=0.1)
        assert resolve_precision(m, True) == resolve_precision(m, False) == precision_planes(name)
