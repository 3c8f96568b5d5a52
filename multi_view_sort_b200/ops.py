"""Stage-level Python wrappers over the C ABI (device tensors in, device tensors out).

These are thin: they validate dtypes/contiguity, pass raw pointers and the current CUDA
stream, and raise on error.  The fused module-level path (functional.py) uses the
orchestrated entry points instead; the wrappers here exist for the stage-level parity tests
and for tools.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Sequence, Tuple

import torch

from ._lib import GemmSeg, check, lib, planes_c, ptr, stream_ptr


SEG_A_F16, SEG_B_F16, SEG_RESCALE11 = 1, 2, 4


def split_planes(t: torch.Tensor, planes: int, fp16_pair: bool = False):
    """fp32 tensor -> tuple of operand planes: `planes` successive bf16 residuals, or (fp16_pair) the
    fp16 pair (hi, (v - hi) * 2^11) - the dtype of the returned tensors tells the format."""
    assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and 1 <= planes <= (2 if fp16_pair else 3)
    out = tuple(torch.empty_like(t, dtype=torch.float16 if fp16_pair else torch.bfloat16) for _ in range(planes))
    pc = planes_c(out)
    check(lib().vmm_split_planes(ptr(t), C.c_longlong(t.numel()), C.byref(pc), stream_ptr()))
    return out


def _segs(pairs: Sequence[Tuple[torch.Tensor, torch.Tensor]], a_mn: bool, b_mn: bool):
    arr = (GemmSeg * len(pairs))()
    M = N = None
    for i, pair in enumerate(pairs):
        a, b = pair[0], pair[1]
        flags = pair[2] if len(pair) > 2 else 0
        flags |= (SEG_A_F16 if a.dtype == torch.float16 else 0) | (SEG_B_F16 if b.dtype == torch.float16 else 0)
        assert a.dtype in (torch.bfloat16, torch.float16) and b.dtype in (torch.bfloat16, torch.float16) and a.is_cuda and b.is_cuda
        assert a.dim() == 2 and b.dim() == 2 and a.stride(1) == 1 and b.stride(1) == 1
        ka, m = (a.shape[0], a.shape[1]) if a_mn else (a.shape[1], a.shape[0])
        kb, n = (b.shape[0], b.shape[1]) if b_mn else (b.shape[1], b.shape[0])
        assert ka == kb, "K mismatch"
        assert M in (None, m) and N in (None, n)
        M, N = m, n
        arr[i] = GemmSeg(a.data_ptr(), a.stride(0), b.data_ptr(), b.stride(0), ka, flags)
    return arr, M, N


def gemm(pairs, a_mn=False, b_mn=False, bias=None, out=None, accumulate=False, splits=0, chain=0, check_kernel=False,
         row_scale=None):
    """C[M,N] (+)= sum_i A_i B_i^T.  A_i is (M,K) when a_mn is False, (K,M) when True; same for B.
    row_scale (float[M]): output row m is multiplied by row_scale[m] (vmm_gemm_rowscaled)."""
    arr, M, N = _segs(pairs, a_mn, b_mn)
    if out is None:
        assert not accumulate
        out = torch.empty(M, N, device=pairs[0][0].device, dtype=torch.float32)
    assert out.dtype == torch.float32 and out.stride(1) == 1 and out.shape == (M, N)
    if check_kernel:
        check(lib().vmm_gemm_check(M, N, len(pairs), arr, int(a_mn), int(b_mn), ptr(out), C.c_longlong(out.stride(0)),
                                   ptr(bias), int(accumulate), stream_ptr()))
    elif row_scale is not None:
        assert bias is None and row_scale.dtype == torch.float32 and row_scale.numel() == M and not a_mn
        check(lib().vmm_gemm_rowscaled(M, N, len(pairs), arr, int(a_mn), int(b_mn), ptr(out), C.c_longlong(out.stride(0)),
                                       ptr(row_scale), int(accumulate), int(splits), int(chain), stream_ptr()))
    else:
        check(lib().vmm_gemm_bf16(M, N, len(pairs), arr, int(a_mn), int(b_mn), ptr(out), C.c_longlong(out.stride(0)),
                                  ptr(bias), int(accumulate), int(splits), int(chain), stream_ptr()))
    return out


def plane_pairs(a_planes, b_planes):
    """Segment list (least significant first) for operands given as planes - mirrors make_segs() in gemm.cu."""
    af, bf = a_planes[0].dtype == torch.float16, b_planes[0].dtype == torch.float16
    out = []
    if not af and not bf:
        order = max(len(a_planes), len(b_planes))
        for tot in range(order - 1, -1, -1):
            for i in range(len(a_planes)):
                j = tot - i
                if 0 <= j < len(b_planes):
                    out.append((a_planes[i], b_planes[j]))
        return out
    if af and len(a_planes) > 1:
        out.append((a_planes[1], b_planes[0]))
    if bf and len(b_planes) > 1:
        out.append((a_planes[0], b_planes[1]))
    scaled = len(out)
    na, nb = (1 if af else len(a_planes)), (1 if bf else len(b_planes))
    for tot in range(na + nb - 2, -1, -1):
        for i in range(na):
            j = tot - i
            if 0 <= j < nb:
                out.append((a_planes[i], b_planes[j], SEG_RESCALE11 if (len(out) == scaled and scaled > 0) else 0))
    return out


def estep_num_blocks(D: int) -> int:
    return lib().vmm_estep_num_blocks(int(D))


def estep_forward(z_planes, w3_planes, b3, x, k_latent, m_views, sigma, want_sq=True, want_pp=False):
    """Fused dec3 + E-step.  z planes (rows, zk), W3 planes (D, zk), x (B*M, D) float or bf16.
    Returns the per-(row, half-tile) partial sums (e, sq, pp); sum over dim 1 for the row totals."""
    rows, zk = z_planes[0].shape
    D = w3_planes[0].shape[0]
    nb = estep_num_blocks(D)
    dev = z_planes[0].device
    pe = torch.empty(rows, nb, device=dev)
    psq = torch.empty(rows, nb, device=dev) if want_sq else None
    ppp = torch.empty(rows, nb, device=dev) if want_pp else None
    assert x.dtype in (torch.float32, torch.bfloat16) and x.stride(-1) == 1
    zc, wc = planes_c(z_planes), planes_c(w3_planes)
    check(lib().vmm_estep_forward(rows, D, zk, C.byref(zc), C.byref(wc), ptr(b3), ptr(x), int(x.dtype == torch.bfloat16),
                                  C.c_longlong(x.stride(-2)), int(k_latent), int(m_views), C.c_float(sigma), ptr(pe),
                                  ptr(psq), ptr(ppp), stream_ptr()))
    return pe, psq, ppp


def estep_backward(z_planes, w3_planes, b3, x, k_latent, m_views, sigma, alpha, beta, kappa=None, d_b3=None,
                   out_planes=None):
    rows, zk = z_planes[0].shape
    D = w3_planes[0].shape[0]
    n_out = len(z_planes) if out_planes is None else out_planes
    outs = tuple(torch.empty(rows, D, device=z_planes[0].device, dtype=torch.bfloat16) for _ in range(n_out))
    zc, wc, oc = planes_c(z_planes), planes_c(w3_planes), planes_c(outs)
    check(lib().vmm_estep_backward(rows, D, zk, C.byref(zc), C.byref(wc), ptr(b3), ptr(x), int(x.dtype == torch.bfloat16),
                                   C.c_longlong(x.stride(-2)), int(k_latent), int(m_views), C.c_float(sigma), ptr(alpha),
                                   ptr(beta), ptr(kappa), C.byref(oc), ptr(d_b3), stream_ptr()))
    return outs


def estep_forward_save(z_planes, w3_planes, b3, x, k_latent, m_views, sigma, delta_dtype=torch.float16, want_sq=True, out=None):
    """estep_forward that also writes delta = x - pred (rows, D) as fp16 (saturating) or fp32.
    Returns (part_e, part_sq, delta).  want_sq=False: what every EM iteration but the last runs (no loss sums);
    out: reuse a (part_e, part_sq, delta) triple."""
    rows, zk = z_planes[0].shape
    D = w3_planes[0].shape[0]
    nb = estep_num_blocks(D)
    dev = z_planes[0].device
    if out is not None:
        pe, psq, delta = out
    else:
        pe, psq = torch.empty(rows, nb, device=dev), (torch.empty(rows, nb, device=dev) if want_sq else None)
        delta = torch.empty(rows, D, device=dev, dtype=delta_dtype)
    zc, wc = planes_c(z_planes), planes_c(w3_planes)
    check(lib().vmm_estep_forward_save(rows, D, zk, C.byref(zc), C.byref(wc), ptr(b3), ptr(x), int(x.dtype == torch.bfloat16),
                                       C.c_longlong(x.stride(-2)), int(k_latent), int(m_views), C.c_float(sigma), ptr(pe),
                                       ptr(psq), None, ptr(delta), 1 if delta_dtype == torch.float16 else 2, stream_ptr()))
    return pe, psq, delta


def estep_dpred(delta, x, k_latent, m_views, sigma, alpha, beta, kappa=None, d_b3=None, out_planes=1):
    """Streaming E-step backward from the saved delta: same outputs as estep_backward."""
    rows, D = delta.shape
    assert delta.dtype in (torch.float16, torch.float32) and delta.is_contiguous()
    outs = tuple(torch.empty(rows, D, device=delta.device, dtype=torch.bfloat16) for _ in range(out_planes))
    oc = planes_c(outs)
    check(lib().vmm_estep_dpred(rows, D, ptr(delta), 1 if delta.dtype == torch.float16 else 2, ptr(x),
                                int(x.dtype == torch.bfloat16), C.c_longlong(x.stride(-2)), int(k_latent), int(m_views),
                                C.c_float(sigma), ptr(alpha), ptr(beta), ptr(kappa), C.byref(oc), ptr(d_b3), stream_ptr()))
    return outs
