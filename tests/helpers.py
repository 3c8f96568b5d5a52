"""Shared test helpers: golden loading, error metrics, seeded models."""
import os

import numpy as np
import torch

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    cfg = eval(str(z["cfg"]))          # repr(dict) of plain ints/floats written by oracle/make_golden.py
    return z, cfg


def rel_err(a, b):
    """max |a-b| / max |b|  - the 'relative' error of north_star's tolerance, tensor-wise."""
    a = torch.as_tensor(np.asarray(a)).double() if not torch.is_tensor(a) else a.detach().double().cpu()
    b = torch.as_tensor(np.asarray(b)).double() if not torch.is_tensor(b) else b.detach().double().cpu()
    denom = b.abs().max().item()
    return (a - b).abs().max().item() / (denom if denom > 0 else 1.0)


def grad_summary(g, nsample=64):
    g = g.detach().double().reshape(-1).cpu()
    idx = (torch.arange(nsample, dtype=torch.int64) * (g.numel() - 1)) // (nsample - 1)
    return np.concatenate([[g.norm().item(), g.abs().max().item(), g.sum().item()], g[idx].numpy()])


def params_of(module, requires_grad=True, dtype=None, device=None):
    out = {}
    for k, v in module.state_dict().items():
        t = v.detach().clone()
        if dtype is not None:
            t = t.to(dtype)
        if device is not None:
            t = t.to(device)
        out[k] = t.requires_grad_(requires_grad)
    return out


def cuda_dec1_pre(nem, T):
    """dec1's pre-activation (the input of the ONE ReLU of the path, train_nem.py:181) of every EM iteration of the last
    CUDA training forward of `nem` (needs nem.keep_workspace = True), recomputed from the library's workspace - the raw
    contraction output g4 and the LayerNorm row statistics - with the kernels' own formula."""
    import ctypes as C
    from multi_view_sort_b200._lib import lib
    cfg, ws = nem.__dict__["_vmm_last_workspace"]
    L = lib()
    L.vmm_em_workspace_ptr.restype = C.c_void_p
    R, W = cfg.B * cfg.K, 4096
    base = ws.data_ptr()
    named = dict(nem.named_parameters())
    b4, g, be = named["dec1.0.bias"].detach(), named["dec1.1.weight"].detach(), named["dec1.1.bias"].detach()
    wsf = ws.view(torch.float32)
    out = []
    for t in range(T):
        def view(name, n):
            p = L.vmm_em_workspace_ptr(C.byref(cfg), C.c_void_p(base), name.encode(), t)
            assert p, name
            off = (p - base) // 4
            return wsf[off:off + n]
        g4 = view("g4", R * W).view(R, W)
        mean, rstd = view("mean4", R), view("rstd4", R)
        xh = ((g4 + b4) - mean[:, None]) * rstd[:, None]
        out.append(torch.addcmul(be, xh, g).cpu())
    return out
