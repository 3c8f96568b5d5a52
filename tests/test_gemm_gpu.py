"""GPU parity of the tcgen05 contraction and the fused E-step epilogues, through the C ABI.

Checked against (a) fp64 torch matmul of the same bf16-rounded operands and (b) the
library's plain one-thread-per-output CUDA kernel.  bf16 x bf16 products are exact in
fp32, so the only difference is fp32 accumulation order: tolerance 2e-5 relative to max|C|.
"""
import pytest
import torch

pytestmark = pytest.mark.gpu


def _rand(shape, seed, scale=1.0):
    g = torch.Generator().manual_seed(seed)
    return (torch.randn(*shape, generator=g) * scale).cuda()


def _relmax(a, b):
    return ((a.double() - b.double()).abs().max() / b.double().abs().max().clamp_min(1e-30)).item()


SHAPES = [(128, 256, 64), (128, 256, 256), (256, 512, 1024), (384, 768, 136), (200, 1000, 72), (1024, 40, 4096),
          (5, 3072, 1024), (3000, 1024, 1024)]


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, False), (True, True)])
@pytest.mark.parametrize("M,N,K", SHAPES)
def test_gemm_layouts(M, N, K, a_mn, b_mn):
    from multi_view_sort_b200 import ops
    Mp, Np = (M + 7) // 8 * 8, (N + 7) // 8 * 8      # pitches must be multiples of 8 elements
    A = _rand((K, Mp) if a_mn else (M, K), 1).bfloat16()
    B = _rand((K, Np) if b_mn else (N, K), 2).bfloat16()
    Av = A[:, :M] if a_mn else A
    Bv = B[:, :N] if b_mn else B
    bias = _rand((N,), 3) if N % 4 == 0 else None
    ref = (Av.double().t() if a_mn else Av.double()) @ (Bv.double() if b_mn else Bv.double().t())
    if bias is not None:
        ref = ref + bias.double()
    out = ops.gemm([(Av, Bv)], a_mn, b_mn, bias=bias)
    chk = ops.gemm([(Av, Bv)], a_mn, b_mn, bias=bias, check_kernel=True)
    torch.cuda.synchronize()
    assert _relmax(chk, ref) < 2e-5, "cross-check kernel disagrees with torch"
    err = _relmax(out, ref)
    assert err < 2e-5, f"tcgen05 GEMM M={M} N={N} K={K} a_mn={a_mn} b_mn={b_mn}: rel err {err:.3e}"


@pytest.mark.parametrize("splits", [0, 1, 3, 7])
def test_gemm_accumulate_splitk(splits):
    from multi_view_sort_b200 import ops
    M, N, K = 256, 512, 4096 + 64
    A = _rand((K, M), 4).bfloat16()            # weight-gradient layout: both operands MN-major
    B = _rand((K, N), 5).bfloat16()
    C0 = _rand((M, N), 6)
    ref = C0.double() + A.double().t() @ B.double()
    out = C0.clone()
    ops.gemm([(A, B)], True, True, out=out, accumulate=True, splits=splits)
    torch.cuda.synchronize()
    assert _relmax(out, ref) < 2e-5


def test_gemm_split_precision_matches_fp32():
    """2 planes (3 passes) reproduce an fp32 contraction to ~1e-5; 3 planes (6 passes) with
    bounded accumulation chains match cuBLAS-sgemm-class accuracy.  The chain bound matters: the
    tensor core truncates at every accumulate step, so one long chain is biased low."""
    from multi_view_sort_b200 import ops
    M, N, K = 512, 768, 16384
    g = torch.Generator().manual_seed(7)
    A = torch.rand(M, K, generator=g).cuda()                 # positive operands expose the truncation bias
    B = (torch.rand(N, K, generator=g) / K).cuda()
    ref = A.double() @ B.double().t()
    errs = {}
    for n, chain in ((1, 0), (2, 0), (3, 0), (3, 4)):
        out = ops.gemm(ops.plane_pairs(ops.split_planes(A, n), ops.split_planes(B, n)), chain=chain)
        torch.cuda.synchronize()
        errs[(n, chain)] = _relmax(out, ref)
    fp32 = _relmax(A @ B.t(), ref)
    print(f"rel err by (planes, chain) {errs}; torch fp32 {fp32:.3e}")
    assert errs[(1, 0)] < 2e-2 and errs[(2, 0)] < 2e-4
    assert errs[(3, 4)] < 2e-6 and errs[(3, 4)] < errs[(3, 0)] / 10


@pytest.mark.parametrize("chain", [1, 4, 16])
def test_gemm_chained_accumulation_layouts(chain):
    """Bounded chains (epilogue read-modify-write) in store mode with a ragged shape, bias and 6 segments."""
    from multi_view_sort_b200 import ops
    M, N, K = 200, 520, 1160
    A, B = _rand((M, K), 31), _rand((N, K), 32, 0.03)
    bias = _rand((N,), 33)
    ref = A.double() @ B.double().t() + bias.double()
    out = ops.gemm(ops.plane_pairs(ops.split_planes(A, 3), ops.split_planes(B, 3)), bias=bias, chain=chain)
    torch.cuda.synchronize()
    assert _relmax(out, ref) < 2e-6
    # weight-gradient layout, accumulate + split-K + chains
    At, Bt = A.t().contiguous(), B.t().contiguous()
    C0 = _rand((M, N), 34)
    acc = C0.clone()
    ops.gemm(ops.plane_pairs(ops.split_planes(At, 3), ops.split_planes(Bt, 3)), True, True, out=acc, accumulate=True,
             splits=3, chain=chain)
    torch.cuda.synchronize()
    assert _relmax(acc, C0.double() + A.double() @ B.double().t()) < 2e-6


@pytest.mark.parametrize("N", [40, 72, 160, 1000])
def test_gemm_chained_accumulation_clipped_n(N):
    """Accumulation chains on tiles clipped at the N edge (the head's logits GEMM: N = 40 classes, K = 4096, chain 16
    -> 4 chains): a clipped tile issues fewer than four TMA groups per chain, and the reduce-add of chain c must still
    wait for the store of chain c-1 to the same addresses (bulk operations are unordered).  Repeated: the hazard is a
    race, and large M keeps many warps in flight."""
    from multi_view_sort_b200 import ops
    M, K = 4096, 4096
    A, B = _rand((M, K), 41).bfloat16(), _rand((N, K), 42, 0.05).bfloat16()
    bias = _rand((N,), 43)
    ref = A.double() @ B.double().t() + bias.double()
    for chain in (16, 4, 1):
        for _ in range(3):
            out = ops.gemm([(A, B)], bias=bias, chain=chain)
            torch.cuda.synchronize()
            assert _relmax(out, ref) < 2e-5, (N, chain)


def test_gemm_two_logical_segments():
    from multi_view_sort_b200 import ops
    M, N = 640, 1024
    A1, B1 = _rand((M, 4096), 9).bfloat16(), _rand((N, 4096), 10).bfloat16()
    A2, B2 = _rand((M, 1024), 11).bfloat16(), _rand((1024, N), 12).bfloat16().t().contiguous()
    ref = A1.double() @ B1.double().t() + A2.double() @ B2.double().t()
    out = ops.gemm([(A1, B1), (A2, B2)])
    torch.cuda.synchronize()
    assert _relmax(out, ref) < 2e-5


def _estep_ref(z, w3, b3, x, K, Mv, sigma):
    rows = z.shape[0]
    B = rows // (K * Mv)
    pred = (z.double() @ w3.double().t() + b3.double()).view(B, K, Mv, -1)
    d = x.double().view(B, 1, Mv, -1) - pred
    e = torch.exp(-d * d / (2 * sigma * sigma))
    return pred, d, e


@pytest.mark.parametrize("planes,xdt", [(1, torch.bfloat16), (2, torch.float32), (1, torch.float32), (3, torch.float32), (2, torch.bfloat16)])
@pytest.mark.parametrize("B,K,Mv,D,zk", [(3, 3, 12, 512, 256), (16, 4, 12, 4096, 1024), (2, 5, 80, 1000, 1024)])
def test_estep_fused_forward_backward(B, K, Mv, D, zk, planes, xdt):
    from multi_view_sort_b200 import ops
    sigma = 0.25
    rows = B * K * Mv
    z = _rand((rows, zk), 20)
    w3 = _rand((D, zk), 21, 1.0 / 32)
    b3 = _rand((D,), 22, 0.1)
    x = torch.relu(_rand((B * Mv, D), 23)).to(xdt)
    zp, wp = ops.split_planes(z, planes), ops.split_planes(w3, planes)
    # what the kernel multiplies: the sum of its planes
    zq = sum(t.double() for t in zp)
    wq = sum(t.double() for t in wp)
    pred, d, e = _estep_ref(zq, wq, b3, x.float(), K, Mv, sigma)
    pe, psq, ppp = ops.estep_forward(zp, wp, b3, x, K, Mv, sigma, want_sq=True, want_pp=True)
    torch.cuda.synchronize()
    tol = 1e-4 if planes == 1 else (2e-5 if planes == 2 else 1e-5)
    assert _relmax(pe.sum(1), e.sum(-1).view(-1)) < tol
    assert _relmax(psq.sum(1), (d * d).sum(-1).view(-1)) < tol
    assert _relmax(ppp.sum(1), (pred * pred).sum(-1).view(-1)) < tol
    alpha, beta, kappa = _rand((rows,), 24), _rand((rows,), 25, 0.1), _rand((rows,), 26, 0.05)
    dref = d * (alpha.double().view(B, K, Mv, 1) * e - beta.double().view(B, K, Mv, 1)) + kappa.double().view(B, K, Mv, 1) * pred
    d_b3 = torch.full((D,), 0.5, device="cuda")
    outs = ops.estep_backward(zp, wp, b3, x, K, Mv, sigma, alpha, beta, kappa, d_b3=d_b3)
    torch.cuda.synchronize()
    got = sum(t.double() for t in outs)
    err = _relmax(got.view(-1), dref.reshape(-1))
    assert err < {1: 6e-3, 2: 3e-5, 3: 1e-5}[planes], f"dpred rel err {err:.3e}"
    # fused dec3-bias gradient: exact fp32 column sums, accumulated into the caller's buffer
    assert _relmax(d_b3 - 0.5, dref.reshape(rows, D).sum(0)) < tol


@pytest.mark.parametrize("ddt,out_planes,xdt,use_kappa", [(torch.float16, 1, torch.bfloat16, False), (torch.float32, 2, torch.float32, True),
                                                          (torch.float32, 3, torch.float32, False), (torch.float16, 2, torch.float32, True)])
@pytest.mark.parametrize("B,K,Mv,D,zk", [(3, 3, 12, 512, 256), (16, 3, 12, 4096, 1024), (2, 5, 80, 1000, 1024)])
def test_estep_saved_delta_backward(B, K, Mv, D, zk, ddt, out_planes, xdt, use_kappa):
    """The forward epilogue's saved delta = x - pred and the streaming backward over it reproduce the
    recompute path (vmm_estep_backward): bit-identical planes with fp32 delta, fp16-rounding close with fp16."""
    from multi_view_sort_b200 import ops
    sigma = 0.25
    rows = B * K * Mv
    z = _rand((rows, zk), 30)
    w3 = _rand((D, zk), 31, 1.0 / 32)
    b3 = _rand((D,), 32, 0.1)
    x = torch.relu(_rand((B * Mv, D), 33)).to(xdt)
    zp, wp = ops.split_planes(z, 2, True), ops.split_planes(w3, 2, True)
    pe0, psq0, _ = ops.estep_forward(zp, wp, b3, x, K, Mv, sigma, want_sq=True)
    pe, psq, delta = ops.estep_forward_save(zp, wp, b3, x, K, Mv, sigma, delta_dtype=ddt)
    torch.cuda.synchronize()
    # saving delta does not change the E-step sums: bit-identical when both launches run the same chunk body (fp32 delta:
    # the fast chunk), equal up to fp32 summation order when the fp16-delta launch takes the general body
    if ddt == torch.float32:
        assert torch.equal(pe, pe0) and torch.equal(psq, psq0)
    else:
        assert _relmax(pe, pe0) < 2e-6 and _relmax(psq, psq0) < 2e-6
    zq = zp[0].double() + zp[1].double() / 2048
    wq = wp[0].double() + wp[1].double() / 2048
    pred, d, e = _estep_ref(zq, wq, b3, x.float(), K, Mv, sigma)
    assert _relmax(delta.double().view(-1), d.reshape(-1)) < (1e-3 if ddt == torch.float16 else 2e-6)
    alpha, beta = _rand((rows,), 34), _rand((rows,), 35, 0.1)
    kappa = _rand((rows,), 36, 0.05) if use_kappa else None
    db_a = torch.zeros(D, device="cuda")
    db_b = torch.zeros(D, device="cuda")
    want = ops.estep_backward(zp, wp, b3, x, K, Mv, sigma, alpha, beta, kappa, d_b3=db_a, out_planes=out_planes)
    got = ops.estep_dpred(delta, x, K, Mv, sigma, alpha, beta, kappa, d_b3=db_b, out_planes=out_planes)
    torch.cuda.synchronize()
    w_sum, g_sum = sum(t.double() for t in want), sum(t.double() for t in got)
    # fp16 delta: one bf16 ulp of the largest entry when the output is a single plane
    tol = (8e-3 if out_planes == 1 else 3e-3) if ddt == torch.float16 else {2: 3e-5, 3: 2e-6}[out_planes]
    assert _relmax(g_sum.view(-1), w_sum.view(-1)) < tol
    assert _relmax(db_b, db_a) < (2e-3 if ddt == torch.float16 else 1e-5)


def test_gemm_fp16_pair_planes():
    """fp16 hi/lo' pair (lo' = lo * 2^11) with the accumulator rescale (tcgen05.mma scale_input_d = 11):
    3 passes give ~22 mantissa bits.  (kind::f16 rejects mixed bf16 x fp16 operands - illegal
    instruction, measured - so gradient contractions use bf16 planes on both sides.)"""
    from multi_view_sort_b200 import ops
    M, N, K = 384, 640, 2048
    A, B = _rand((M, K), 41), _rand((N, K), 42, 0.02)
    ref = A.double() @ B.double().t()
    errs = {}
    out = ops.gemm(ops.plane_pairs(ops.split_planes(A, 1, True), ops.split_planes(B, 1, True)))
    errs["f16x1"] = _relmax(out, ref)
    pairs = ops.plane_pairs(ops.split_planes(A, 2, True), ops.split_planes(B, 2, True))
    assert len(pairs) == 3
    chk = ops.gemm(pairs, check_kernel=True)
    errs["f16pair_check_kernel"] = _relmax(chk, ref)
    for chain in (0, 4):
        out = ops.gemm(pairs, chain=chain)
        errs[f"f16pair_chain{chain}"] = _relmax(out, ref)
    # weight-gradient layout (both MN-major), accumulate + split-K
    At, Bt = A.t().contiguous(), B.t().contiguous()
    acc = torch.zeros(M, N, device="cuda")
    ops.gemm(ops.plane_pairs(ops.split_planes(At, 2, True), ops.split_planes(Bt, 2, True)), True, True, out=acc,
             accumulate=True, splits=4, chain=4)
    errs["mn_f16pair"] = _relmax(acc, ref)
    # an fp16 pair against a single exact fp16 plane (bf16 features up-cast to fp16)
    Ab = A.bfloat16().float()
    out = ops.gemm(ops.plane_pairs(ops.split_planes(Ab, 1, True), ops.split_planes(B, 2, True)), chain=4)
    errs["f16x1_f16pair"] = _relmax(out, Ab.double() @ B.double().t())
    torch.cuda.synchronize()
    print(errs)
    assert errs["f16x1"] < 2e-3
    assert errs["f16pair_check_kernel"] < 2e-6, "reference semantics of the fp16 pair are wrong"
    assert errs["f16pair_chain4"] < 2e-6 and errs["mn_f16pair"] < 2e-6 and errs["f16x1_f16pair"] < 2e-6


@pytest.mark.parametrize("a_mn,b_mn", [(False, False), (False, True), (True, True)])
@pytest.mark.parametrize("M,N,K", [(256, 512, 1024), (384, 768, 136), (3000, 1024, 1024), (1024, 40, 4096)])
def test_cta_pair_kernel_is_bit_identical_to_single_cta(M, N, K, a_mn, b_mn):
    """The CTA-pair execution shape (cluster of 2, tcgen05 cta_group::2, the default for M > 128) accumulates every
    output element over the same K order as the single-CTA kernel: results are bit-identical, including odd numbers
    of 128-row blocks (the pair's second CTA runs on an all-out-of-bounds tile) and chained accumulation."""
    from multi_view_sort_b200 import ops
    from multi_view_sort_b200._lib import lib
    A = _rand((K, M) if a_mn else (M, K), 51).bfloat16()
    B = _rand((K, N) if b_mn else (N, K), 52).bfloat16()
    outs = []
    try:
        for ncta in (1, 2):
            lib().vmm_gemm_force_ncta(ncta)
            outs.append((ops.gemm([(A, B)], a_mn, b_mn), ops.gemm([(A, B), (A, B)], a_mn, b_mn, chain=4)))
    finally:
        lib().vmm_gemm_force_ncta(0)
    torch.cuda.synchronize()
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


@pytest.mark.parametrize("accumulate", [False, True])
def test_gemm_row_scaled_fp16_operand(accumulate):
    """vmm_gemm_rowscaled: A rows stored as fp16(a * s_m) with per-row powers of two spanning 40 binades, the
    contraction un-scales them in its epilogue - the result matches the unscaled product to fp16 operand precision
    although most rows of A would underflow or overflow fp16 without the scale."""
    from multi_view_sort_b200 import ops
    M, N, K = 300, 520, 1024
    A = _rand((M, K), 61)
    A = A * torch.exp2(torch.randint(-30, 10, (M, 1), device="cuda").float())          # rows of very different magnitude
    W = _rand((K, N), 62, 0.05)                                                        # weight stored [n=K... ] MN-major
    amax = A.abs().amax(1, keepdim=True)
    s = torch.exp2(12 - torch.ceil(torch.log2(amax)))
    Ah = (A * s).half()
    assert torch.isfinite(Ah).all()
    Wh = W.half()
    C0 = _rand((M, N), 63) if accumulate else None
    out = C0.clone() if accumulate else None
    out = ops.gemm([(Ah, Wh)], False, True, out=out, accumulate=accumulate, row_scale=(1.0 / s).reshape(-1).contiguous())
    torch.cuda.synchronize()
    ref = (Ah.double() / s.double()) @ Wh.double()
    if accumulate:
        ref = ref + C0.double()
    err = ((out.double() - ref).abs() / ref.abs().amax(1, keepdim=True).clamp_min(1e-300)).max().item()
    assert err < 2e-5, err                                                             # per ROW: every row keeps its precision
    assert _relmax(out.double() - (C0.double() if accumulate else 0), A.double() @ W.double()) < 2e-3   # vs the fp32 operands


@pytest.mark.parametrize("M,N,K", [(2560, 2304, 2048), (12288, 4096, 1024), (2560, 2304 + 40, 2048 + 72)])
@pytest.mark.parametrize("mode", ["store_bias", "store_chain4", "accumulate", "mn_store", "planes2", "rowscale"])
def test_gemm_tail_split(M, N, K, mode):
    """Shapes whose tile count leaves a partial last wave (90 tiles / 768 tiles on 74 CTA pairs): the tail tiles are
    split along K (GemmParams::tail_*), by chain ranges when the launch has enough accumulation chains and by shorter
    chains otherwise; stores add into a zeroed rectangle of C, the bias is added once.  Against fp32 matmul of the
    same bf16 operands."""
    from multi_view_sort_b200 import ops
    from multi_view_sort_b200._lib import lib
    lib().vmm_gemm_tail_split(1)              # every launch (the default confines it to the backward entry points)
    try:
        _tail_split_case(ops, M, N, K, mode)
    finally:
        lib().vmm_gemm_tail_split(-1)


def _tail_split_case(ops, M, N, K, mode):
    A, B = _rand((M, K), 70).bfloat16(), _rand((N, K), 71, 0.05).bfloat16()
    ref = A.double() @ B.double().t()
    if mode == "store_bias":
        bias = _rand((N,), 72)
        out = torch.full((M, N), 7.0, device="cuda")          # stale contents must not leak into the zeroed rectangle
        ops.gemm([(A, B)], bias=bias, out=out)
        ref = ref + bias.double()
    elif mode == "store_chain4":
        out = ops.gemm([(A, B)], chain=4)
    elif mode == "accumulate":
        out = _rand((M, N), 73)
        ref = ref + out.double()
        ops.gemm([(A, B)], out=out, accumulate=True)
    elif mode == "mn_store":
        out = ops.gemm([(A.t().contiguous(), B.t().contiguous())], True, True)
    elif mode == "planes2":
        A32, B32 = _rand((M, K), 70), _rand((N, K), 71, 0.05)
        ap, bp = ops.split_planes(A32, 2), ops.split_planes(B32, 2)
        out = ops.gemm(ops.plane_pairs(ap, bp), chain=16)
        ref = A32.double() @ B32.double().t()
        torch.cuda.synchronize()
        assert _relmax(out, ref) < 3e-5
        return
    else:
        rs = torch.rand(M, device="cuda") + 0.5
        out = ops.gemm([(A, B)], row_scale=rs)
        ref = ref * rs.double().view(-1, 1)
    torch.cuda.synchronize()
    assert _relmax(out, ref) < 2e-5, mode
