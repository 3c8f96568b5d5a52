"""torch.autograd bindings of the orchestrated C-ABI entry points.

`run_em` is the fused path: every EM iteration, the E-step and the last iteration's outer
loss in ONE autograd node whose forward is `vmm_em_forward` and whose backward is
`vmm_em_backward` (hand-written BPTT).  torch only provides device memory, the stream and the
autograd plumbing; no torch op touches the data.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional

import torch

from ._lib import check, lib, ptr, stream_ptr


# ---- ctypes mirrors of include/vmm_b200.h ----------------------------------------------------
class EMConfig(C.Structure):
    _fields_ = [("B", C.c_int), ("K", C.c_int), ("M", C.c_int), ("D", C.c_int), ("H", C.c_int), ("T", C.c_int),
                ("planes", C.c_int), ("planes_bwd", C.c_int), ("fwd_fp16", C.c_int), ("x_is_bf16", C.c_int), ("if_gamma_prior", C.c_int), ("stop_gamma_grad", C.c_int),
                ("training", C.c_int), ("accum_chain", C.c_int), ("sigma", C.c_float), ("ln_eps", C.c_float),
                ("dropout_p", C.c_float), ("dropout_seed", C.c_ulonglong), ("fwd_modes", C.c_uint), ("save_delta", C.c_int), ("dgrad_fp16", C.c_int),
                ("warm_up", C.c_int), ("loss_all", C.c_int)]


EM_PARAM_FIELDS = [
    ("enc1_w", "enc1.0.weight"), ("enc1_b", "enc1.0.bias"), ("enc1_ln_g", "enc1.1.weight"), ("enc1_ln_b", "enc1.1.bias"),
    ("enc2_w", "enc2.0.weight"), ("enc2_b", "enc2.0.bias"), ("enc2_ln_g", "enc2.1.weight"), ("enc2_ln_b", "enc2.1.bias"),
    ("gru_w_ih", "rnn_nem.weight_ih"), ("gru_w_hh", "rnn_nem.weight_hh"), ("gru_b_ih", "rnn_nem.bias_ih"),
    ("gru_b_hh", "rnn_nem.bias_hh"),
    ("dec1_w", "dec1.0.weight"), ("dec1_b", "dec1.0.bias"), ("dec1_ln_g", "dec1.1.weight"), ("dec1_ln_b", "dec1.1.bias"),
    ("dec2_w", "dec2.0.weight"), ("dec2_b", "dec2.0.bias"), ("dec2_ln_g", "dec2.1.weight"), ("dec2_ln_b", "dec2.1.bias"),
    ("dec3_w", "dec3.weight"), ("dec3_b", "dec3.bias"),
]


class EMPointers(C.Structure):      # vmm_em_params and vmm_em_grads share this layout
    _fields_ = [(f, C.c_void_p) for f, _ in EM_PARAM_FIELDS]


def _pointers(tensors) -> EMPointers:
    s = EMPointers()
    for (f, _), t in zip(EM_PARAM_FIELDS, tensors):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), f
        setattr(s, f, t.data_ptr())
    return s


def em_param_tensors(module):
    """The 22 live parameters of a NEM module in the C struct's order (after_rnn.* excluded:
    the reference never applies it, train_nem.py:179 vs 232-235, so it must keep grad=None)."""
    named = dict(module.named_parameters())
    return [named[k] for _, k in EM_PARAM_FIELDS]


def _bytes(fn, cfg) -> int:
    fn.restype = C.c_longlong
    n = fn(C.byref(cfg))
    if n < 0:
        check(int(n))
    return int(n)


# precision name -> (forward operand planes, forward planes are fp16 pairs, bf16 planes of the gradient contractions,
# accumulation-chain bound in 64-wide K blocks, per-contraction forward modes, delta storage).  Measured on B200
# against the reference (tests/test_em_gpu.py, tools/precision_sweep.py, tools/mode_sweep.py): the E-step amplifies
# forward errors by 1/sigma^2 and dec1's ReLU turns them into gradient flips, so a forward pass that a backward pass
# follows needs fp32-class operands (an fp16 hi/lo pair = 22 mantissa bits in 3 tensor-core passes) and short
# accumulation chains, while the gradient contractions get by with fewer bits.  resolve_precision() below applies
# the sigma rule of the "bf16" preset on top of this table.
PRECISIONS = {
    "fp32": (2, 1, 2, 16, 0, 2),          # fp32-class forward, 2 bf16 planes backward        (tolerance 1e-4)
    "fp32_strict": (2, 1, 3, 4, 0, 2),    # + 3 bf16 planes backward, 256-element chains
    # "bf16 features" mode (1e-2 bar): fp32-class forward, ONE tensor-core pass per gradient contraction.  The
    # data-gradient contractions of the BPTT chain read row-scaled fp16 operands (the gradient, and the fp16 hi plane of
    # the weight): 11 mantissa bits instead of 8 where rounding errors accumulate, so every gradient tensor stays within
    # ~4e-3 of the reference (2.5x inside the bar, also on the ill-conditioned sigma = 0.02 case); the weight-gradient
    # contractions (reductions over rows) read bf16 planes.
    "bf16": (2, 1, 1, 16, 0, 2, 1),
    "bf16_h": (2, 1, 1, 16, 1 << 14, 2, 1),      # the same with dec3 single-pass regardless of sigma (A/B)
    # round 1's headline: bf16 planes in the data-gradient contractions too - 4 % faster, gradients within 6e-3 .. 1.1e-2
    "bf16_b8": (2, 1, 1, 16, 0, 2),
    "bf16_fast": (1, 0, 1, 0, 0, 1),      # single-pass bf16 everywhere: inference grade (h ~2e-2)
    "fp32_bf16planes": (3, 0, 2, 16, 0, 2),   # the same accuracy from 3 bf16 residual planes (6 forward passes)
    # ---- A/B variants (tools/) ----
    # E-step backward by recomputing pred (one more dec3 contraction per iteration) instead of streaming over the
    # saved fp32 delta = x - pred: numerically identical, 12 % slower, 0.59 MB/object-iteration less workspace
    "bf16_recompute": (2, 1, 1, 16, 0, 0),
    "fp32_recompute": (2, 1, 2, 16, 0, 0),
    "bf16_delta16": (2, 1, 1, 16, 0, 1),         # fp16 delta: half the delta traffic, gradients ~2e-3 further off
    "bf16_pair_dec3": (2, 1, 1, 16, 0, 2),       # "bf16" without the sigma rule: dec3 always on the fp16 pair
    "bf16_chain32": (2, 1, 1, 32, 1 << 14, 2),   # timing probes of the chain bound (dec3 single-pass): -1 % / -2.4 %
    "bf16_chain0": (2, 1, 1, 0, 1 << 14, 2),     #   of the step time, but chain 64 / 0 fail the sigma = 0.02 case
}


# forward contraction ids of vmm_em_config.fwd_modes (2 bits each) and the VMM_FWD_* / VMM_DELTA_* values
FWD_CONTRACTIONS = ("xw1", "g1", "enc2", "gi", "gh", "dec1", "dec2", "dec3", "ebwd")
FWD_FULL, FWD_SINGLE, FWD_ACT_PAIR, FWD_W_PAIR = 0, 1, 2, 3
DELTA_NONE, DELTA_F16, DELTA_F32 = 0, 1, 2


def fwd_modes(**kw) -> int:
    """fwd_modes(enc2=FWD_SINGLE, dec2=FWD_ACT_PAIR, ...) -> the packed bit field."""
    m = 0
    for name, v in kw.items():
        m |= (int(v) & 3) << (2 * FWD_CONTRACTIONS.index(name))
    return m


def precision_planes(precision):
    """-> (forward planes, fp16 pair?, backward bf16 planes, chain, fwd_modes, save_delta, dgrad_fp16); shorter
    tuples are padded with zeros = every forward contraction at full planes, E-step backward by recompute, bf16
    operands in every gradient contraction."""
    t = tuple(int(v) for v in precision) if isinstance(precision, (tuple, list)) else PRECISIONS[precision]
    return t + (0, 0, 0)[len(t) - 4:] if len(t) < 7 else t


# dec3 feeds ONLY the E-step (sum_d exp(-(x - pred)^2 / 2 sigma^2)) and the loss: its rounding errors average over
# the D feature columns instead of entering the recurrent state, so it is the one forward contraction that
# tolerates a single fp16 pass - as long as the E-step's gain 1/sigma^2 is moderate.  Measured with exact
# gradient contractions (tools/mode_sweep.py fp32): single-pass dec3 adds 5e-5 (cfg1) / 2e-4 (prior0) of gradient
# error at sigma = 0.1 and 2e-1 at sigma = 0.02, i.e. ~2e-4 * (0.1 / sigma)^4.  The bf16-features preset (bar 1e-2)
# therefore runs dec3 in one pass when sigma >= DEC3_SINGLE_MIN_SIGMA (contribution <= 1e-3) and keeps the fp16
# pair below it; every other forward contraction needs the pair (single pass: 5e-2 .. 1e-1 gradient error).
DEC3_SINGLE_MIN_SIGMA = 0.07


def resolve_precision(module, training=True):
    """precision_planes(module.precision) with the sigma rule of the "bf16" preset applied.  Training: dec3 in one
    fp16 pass.  Forward-only calls (no gradient can follow): EVERY contraction in one fp16 pass - without a
    backward pass there is no ReLU-flip / BPTT amplification to protect, and the descriptors stay within 3e-3 of
    the reference (bar 1e-2, tests/test_em_gpu.py::test_bf16_inference_single_pass) at ~2.3x the throughput."""
    t = precision_planes(module.precision)
    if module.precision in ("bf16", "bf16_b8") and float(module.e_sigma) >= DEC3_SINGLE_MIN_SIGMA:
        if training:
            modes = t[4] | fwd_modes(dec3=FWD_SINGLE)
        else:
            modes = fwd_modes(**{n: FWD_SINGLE for n in FWD_CONTRACTIONS})
        t = t[:4] + (modes,) + t[5:]
    return t


def dropout_args(module):
    """(p, seed) of this call: p = 0.5 in train mode (the reference's nn.Dropout(0.5)), 0 in eval mode.  The seed
    is `module.dropout_seed` when set (tests), else derived from torch's seed and a per-module call counter."""
    if not module.training:
        return 0.0, 0
    p = float(getattr(module, "dropout_p", 0.5))
    fixed = getattr(module, "dropout_seed", None)
    if fixed is not None:
        return p, int(fixed) & 0xFFFFFFFFFF
    n = module.__dict__.get("_vmm_dropout_calls", 0)
    module.__dict__["_vmm_dropout_calls"] = n + 1
    rank = torch.distributed.get_rank() if torch.distributed.is_available() and torch.distributed.is_initialized() else 0
    # torch.initial_seed() is the same on every rank (train_nem seeds 10 everywhere): mix the rank in, or all ranks
    # would drop the same units of their local objects
    return p, ((torch.initial_seed() * 1000003) ^ (n * 7919 + 12345) ^ (rank * 0x9E3779B1)) & 0xFFFFFFFFFF


def make_config(module, B, T, x_is_bf16, training, warm_up=False, loss_all=False) -> EMConfig:
    planes, f16, planes_bwd, chain, modes, save_delta, dgrad_fp16 = resolve_precision(module, training)
    drop_p, drop_seed = dropout_args(module)
    return EMConfig(warm_up=int(bool(warm_up)), loss_all=int(bool(loss_all)), fwd_modes=modes, save_delta=save_delta, dgrad_fp16=dgrad_fp16, B=B, K=module.k, M=module.m, D=module.input_dim, H=module.rnn_hidden_size, T=T, planes=planes,
                    planes_bwd=planes_bwd, fwd_fp16=f16,
                    x_is_bf16=int(x_is_bf16), if_gamma_prior=int(module.if_gamma_prior),
                    stop_gamma_grad=int(module.stop_gamma_grad), training=int(training),
                    accum_chain=chain, sigma=float(module.e_sigma),
                    ln_eps=float(module.enc1[1].eps), dropout_p=drop_p, dropout_seed=drop_seed)


class _WeightCache:
    """Operand planes of the weights (+ the W1*W3 fold).  Forward-only calls reuse them while no parameter changed
    (data pointer + autograd version).  Training calls ALWAYS rebuild them, and a backward pass marks them stale:
    fused optimizers (torch.optim.Adam(fused=True), the multi-tensor CUDA kernels) update parameters in place
    WITHOUT bumping the version counter (measured), so a version-keyed cache would keep serving the planes of the
    previous step.  Rebuilding costs ~0.35 ms per step (1.3 GB of streaming) at configs[1]."""

    def __init__(self):
        self.key = None
        self.buf = None

    def invalidate(self):
        self.key = None

    def get(self, cfg: EMConfig, params, force=False):
        key = (cfg.planes, cfg.planes_bwd, cfg.fwd_fp16, cfg.accum_chain, cfg.M, cfg.D, cfg.H) + tuple((p.data_ptr(), p._version) for p in params)
        if force or key != self.key:
            nbytes = _bytes(lib().vmm_em_weights_bytes, cfg)
            if self.buf is None or self.buf.numel() < nbytes or self.buf.device != params[0].device:
                self.buf = torch.empty(nbytes, dtype=torch.uint8, device=params[0].device)
            pp = _pointers(params)
            check(lib().vmm_em_prepare(C.byref(cfg), C.byref(pp), ptr(self.buf), None, stream_ptr()))
            self.key = key
        return self.buf


def _weight_cache(module) -> _WeightCache:
    wc = module.__dict__.get("_vmm_weight_cache")
    if wc is None:
        wc = _WeightCache()
        module.__dict__["_vmm_weight_cache"] = wc
    return wc


class GradArena:
    """One flat fp32 buffer that a module's hand-written backward writes its parameter gradients into (instead of one
    zeros_like per parameter): the gradients autograd hands to the parameters are then VIEWS of one contiguous buffer,
    which dist.GradBucket all-reduces in place - no packing copy.  Attached to a module by GradBucket; used only when
    every parameter's .grad is None (optimizer.zero_grad(set_to_none=True), the default), because accumulating into an
    existing .grad that aliases the arena would add the buffer to itself."""

    def __init__(self, params):
        self.params = list(params)
        self.sizes = [p.numel() for p in self.params]
        self.flat = None

    def views_nozero(self):
        out, off = [], 0
        for p, k in zip(self.params, self.sizes):
            out.append(self.flat[off:off + k].view(p.shape))
            off += k
        return out

    def views(self):
        dev = self.params[0].device
        n = sum(self.sizes)
        if self.flat is None or self.flat.device != dev or self.flat.numel() != n:
            self.flat = torch.empty(n, dtype=torch.float32, device=dev)
        self.flat.zero_()
        out, off = [], 0
        for p, k in zip(self.params, self.sizes):
            out.append(self.flat[off:off + k].view(p.shape))
            off += k
        return out


def _grad_buffers(module, params):
    """-> (zeroed gradient buffers for `params`, direct).  direct: the buffers are views of the module's GradArena (one
    is attached and every .grad is None); the backward then installs them as the parameters' .grad itself and returns
    None to autograd, so that .grad provably aliases the arena (autograd may clone a returned gradient)."""
    arena = module.__dict__.get("_vmm_grad_arena")
    if (arena is not None and len(arena.params) == len(params) and all(a is p for a, p in zip(arena.params, params))
            and all(p.grad is None for p in params)):
        return arena.views(), True
    return [torch.zeros_like(p) for p in params], False


def _install(params, grads):
    for p, g in zip(params, grads):
        if p.requires_grad:
            p.grad = g
    return [None] * len(grads)


def _after_backward(module):
    cb = module.__dict__.get("_vmm_after_backward")
    if cb is not None:
        cb()


class _EMFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, T, training, flags, x, gamma0, prior, *params):
        B = x.shape[0]
        dev = x.device
        cfg = make_config(module, B, T, x.dtype == torch.bfloat16, training, warm_up=flags[0], loss_all=flags[1])
        with torch.cuda.device(dev):
            weights = _weight_cache(module).get(cfg, params, force=bool(training))
            ws = torch.empty(_bytes(lib().vmm_em_workspace_bytes, cfg), dtype=torch.uint8, device=dev)
            scratch = torch.empty(_bytes(lib().vmm_em_scratch_bytes, cfg), dtype=torch.uint8, device=dev)
            h = torch.empty(B * cfg.K, cfg.H, device=dev)
            gamma = torch.empty(B, cfg.K, cfg.M, 1, device=dev)
            loss = torch.empty(3 * T if cfg.loss_all else 3, device=dev)
            pp = _pointers(params)
            check(lib().vmm_em_forward(C.byref(cfg), C.byref(pp), ptr(weights), ptr(x), ptr(gamma0), ptr(prior), ptr(ws),
                                       ptr(scratch), ptr(h), ptr(gamma), ptr(loss), stream_ptr()))
        ctx.cfg = cfg
        ctx.module = module
        ctx.nparams = len(params)
        if module.__dict__.get("keep_workspace"):           # tests / tools: stage-level inspection through vmm_em_workspace_ptr
            module.__dict__["_vmm_last_workspace"] = (cfg, ws)
        if training:
            ctx.save_for_backward(x, gamma0, prior, ws, weights, *params)
        return h, gamma, loss

    @staticmethod
    def backward(ctx, d_h, d_gamma, d_loss):
        x, gamma0, prior, ws, weights = ctx.saved_tensors[:5]
        params = ctx.saved_tensors[5:]
        cfg = ctx.cfg
        dev = x.device
        with torch.cuda.device(dev):
            scratch = torch.empty(_bytes(lib().vmm_em_scratch_bytes, cfg), dtype=torch.uint8, device=dev)
            live = em_param_tensors(ctx.module)
            grads, direct = _grad_buffers(ctx.module, live)
            gp, pp = _pointers(grads), _pointers(params)
            d_h = d_h.contiguous().float() if d_h is not None else None
            d_gamma = d_gamma.contiguous().float() if d_gamma is not None else None
            # with loss_all only the LAST iteration's loss carries gradient (the reference detaches the others)
            d_loss = d_loss[-3:].contiguous().float() if d_loss is not None else torch.zeros(3, device=dev)
            check(lib().vmm_em_backward(C.byref(cfg), C.byref(pp), ptr(weights), ptr(x), ptr(gamma0), ptr(prior), ptr(ws),
                                        ptr(scratch), ptr(d_h), ptr(d_gamma), ptr(d_loss), C.byref(gp), stream_ptr()))
        _weight_cache(ctx.module).invalidate()          # an optimizer step follows: the planes are stale from here on
        if direct:
            grads = _install(live, grads)
        _after_backward(ctx.module)
        return (None, None, None, None, None, None, None, *grads)


def run_em(module, x: torch.Tensor, iters: int, gamma0: Optional[torch.Tensor] = None,
           gamma_prior: Optional[torch.Tensor] = None, warm_up: bool = False):
    """All EM iterations + last-iteration outer loss.  Returns (h_T, gamma_T, nem_loss, intra, inter).

    warm_up: the decoder warm-up of tools/Trainer.py:184-185 (every iteration's residual is weighted by gamma_prior).
    With module.if_total_loss == 1 the outer loss of every iteration is also evaluated and left (detached, (T, 3) =
    nem/intra/inter per iteration) in `module.iteration_losses` for the trainer's logged scalar (Trainer.py:203)."""
    if not x.is_cuda:
        raise RuntimeError("multi_view_sort_b200 has no CPU path: x must be a CUDA tensor")
    assert x.dim() == 3 and x.shape[1] == module.m and x.shape[2] == module.input_dim, "x must be (B, M, D)"
    if x.dtype not in (torch.float32, torch.bfloat16):
        x = x.float()
    x = x.contiguous()
    B = x.shape[0]
    dev = x.device
    if gamma0 is None or (gamma_prior is None and module.if_gamma_prior != 0):
        assert x.shape[1] == module.m, "number of view dose not match"
        g0, gp = module.init_gamma(B, dev)               # init_state's gamma / gamma_prior without its h and pred buffers
        gamma0 = g0 if gamma0 is None else gamma0
        gamma_prior = gp if gamma_prior is None else gamma_prior
    gamma0 = gamma0.to(dev, torch.float32).reshape(B, module.k, module.m).contiguous()
    if gamma_prior is None:
        gamma_prior = gamma0
    gamma_prior = gamma_prior.to(dev, torch.float32).reshape(B, module.k, module.m).contiguous()
    params = em_param_tensors(module)
    # keep every iteration's activations only when a backward pass can follow
    training = torch.is_grad_enabled() and any(p.requires_grad for p in params)
    loss_all = int(getattr(module, "if_total_loss", 0)) == 1
    h, gamma, loss = _EMFunction.apply(module, int(iters), training, (bool(warm_up), loss_all), x, gamma0, gamma_prior, *params)
    if loss_all:
        module.__dict__["iteration_losses"] = loss.detach().view(int(iters), 3)
        loss = loss[-3:]
    return h, gamma, loss[0], loss[1], loss[2]


class _EMStepFunction(torch.autograd.Function):
    """ONE EM iteration NEM.forward(x, (h, pred, gamma)) as an autograd node: vmm_em_step_forward / vmm_em_step_backward.
    Chained T times by the caller it reproduces autograd through the reference's own training loop
    (tools/Trainer.py:180-198); NEM.run_em does the same work in one node and far less memory."""

    @staticmethod
    def forward(ctx, module, training, x, h, pred, gamma, *params):
        B = x.shape[0]
        K, M, D, H = module.k, module.m, module.input_dim, module.rnn_hidden_size
        dev = x.device
        cfg = make_config(module, B, 1, x.dtype == torch.bfloat16, training)
        cfg.dgrad_fp16 = 0
        if training:
            cfg.save_delta = DELTA_F32
        with torch.cuda.device(dev):
            weights = _weight_cache(module).get(cfg, params)      # stale after any backward (invalidate()): rebuilt once per step
            ws = torch.empty(_bytes(lib().vmm_em_workspace_bytes, cfg), dtype=torch.uint8, device=dev)
            scratch = torch.empty(_bytes(lib().vmm_em_step_scratch_bytes, cfg), dtype=torch.uint8, device=dev)
            h_out = torch.empty(B * K, H, device=dev)
            pred_out = torch.empty(B, K, M, D, device=dev)
            gamma_out = torch.empty(B, K, M, 1, device=dev)
            pp = _pointers(params)
            check(lib().vmm_em_step_forward(C.byref(cfg), C.byref(pp), ptr(weights), ptr(x), ptr(h), ptr(pred), ptr(gamma),
                                            ptr(ws), ptr(scratch), ptr(h_out), ptr(pred_out), ptr(gamma_out), stream_ptr()))
        ctx.cfg, ctx.module = cfg, module
        if training:
            ctx.save_for_backward(x, h, pred, gamma, ws, scratch, weights, *params)
        return h_out, pred_out, gamma_out

    @staticmethod
    def backward(ctx, d_h, d_pred, d_gamma):
        x, h, pred, gamma, ws, step_scratch, weights = ctx.saved_tensors[:7]
        params = ctx.saved_tensors[7:]
        cfg, dev = ctx.cfg, x.device
        need_h, need_pred, need_gamma = ctx.needs_input_grad[3:6]
        with torch.cuda.device(dev):
            scratch = torch.empty(_bytes(lib().vmm_em_scratch_bytes, cfg), dtype=torch.uint8, device=dev)
            grads = [torch.zeros_like(p) for p in params]
            gp, pp = _pointers(grads), _pointers(params)
            d_h = d_h.contiguous().float() if d_h is not None else None
            d_pred = d_pred.contiguous().float() if d_pred is not None else None
            d_gamma = d_gamma.contiguous().float() if d_gamma is not None else None
            d_h_in = torch.empty_like(h) if need_h else None
            d_pred_in = torch.empty_like(pred) if (need_pred or need_gamma) else None
            d_gamma_in = torch.empty_like(gamma) if need_gamma else None
            check(lib().vmm_em_step_backward(C.byref(cfg), C.byref(pp), ptr(weights), ptr(x), ptr(h), ptr(pred), ptr(gamma),
                                             ptr(ws), ptr(step_scratch), ptr(scratch), ptr(d_h), ptr(d_pred), ptr(d_gamma),
                                             C.byref(gp), ptr(d_h_in), ptr(d_pred_in), ptr(d_gamma_in), stream_ptr()))
        _weight_cache(ctx.module).invalidate()          # an optimizer step follows: the planes are stale from here on
        return (None, None, None, d_h_in, d_pred_in if need_pred else None, d_gamma_in, *grads)


def em_step(module, x, state):
    """NEM.forward(x, (h, pred, gamma)) -> (h', pred', gamma'): one EM iteration from an arbitrary state
    (train_nem.py:272-294).  Differentiable w.r.t. the parameters and the incoming state, so the reference's own
    loops - training (tools/Trainer.py:180-219), validation (:322-336), export (:433-446) - run unchanged on this
    module.  It materialises pred (B,K,M,D) per call and keeps every iteration's activations for autograd, as the
    reference does; NEM.run_em is the fused, memory-lean path."""
    h, pred, gamma = state
    if not x.is_cuda:
        raise RuntimeError("multi_view_sort_b200 has no CPU path: x must be a CUDA tensor")
    B = x.shape[0]
    K, M, D, H = module.k, module.m, module.input_dim, module.rnn_hidden_size
    xx = x.reshape(B, M, D)
    if xx.dtype not in (torch.float32, torch.bfloat16):
        xx = xx.float()
    xx = xx.contiguous()
    dev = xx.device
    params = em_param_tensors(module)
    h_in = h.to(dev, torch.float32).reshape(B * K, H).contiguous()
    pred_in = pred.to(dev, torch.float32).reshape(B, K, M, D).contiguous()
    gamma_in = gamma.to(dev, torch.float32).reshape(B, K, M, 1).contiguous()
    training = torch.is_grad_enabled() and (any(p.requires_grad for p in params) or h_in.requires_grad
                                            or pred_in.requires_grad or gamma_in.requires_grad)
    return _EMStepFunction.apply(module, training, xx, h_in, pred_in, gamma_in, *params)


# ---- head ---------------------------------------------------------------------------------------
class HeadConfig(C.Structure):
    _fields_ = [("B", C.c_int), ("K", C.c_int), ("H", C.c_int), ("C", C.c_int), ("W", C.c_int), ("M", C.c_int),
                ("planes", C.c_int), ("planes_bwd", C.c_int), ("fwd_fp16", C.c_int), ("accum_chain", C.c_int),
                ("if_sort", C.c_int), ("if_pooling", C.c_int), ("dropout_p", C.c_float), ("dropout_seed", C.c_ulonglong)]


HEAD_PARAM_FIELDS = [("att_w", "att.0.weight"), ("att_b", "att.0.bias"), ("w0", "net.0.weight"), ("b0", "net.0.bias"),
                     ("w3", "net.3.weight"), ("b3", "net.3.bias"), ("w6", "net.6.weight"), ("b6", "net.6.bias")]


class HeadPointers(C.Structure):
    _fields_ = [(f, C.c_void_p) for f, _ in HEAD_PARAM_FIELDS]


def _head_pointers(tensors) -> HeadPointers:
    s = HeadPointers()
    for (f, _), t in zip(HEAD_PARAM_FIELDS, tensors):
        assert t.is_cuda and t.dtype == torch.float32 and t.is_contiguous(), f
        setattr(s, f, t.data_ptr())
    return s


def head_param_tensors(module):
    named = dict(module.named_parameters())
    return [named[k] for _, k in HEAD_PARAM_FIELDS]


def make_head_config(module, B, M) -> HeadConfig:
    planes, f16, planes_bwd, chain = precision_planes(module.precision)[:4]
    drop_p, drop_seed = dropout_args(module)
    return HeadConfig(B=B, K=module.k, H=module.rnn_hidden_size, C=module.nclasses, W=module.net[3].in_features, M=M,
                      planes=planes, planes_bwd=planes_bwd, fwd_fp16=f16, accum_chain=chain, if_sort=int(module.if_sort),
                      if_pooling=int(module.if_pooling), dropout_p=drop_p, dropout_seed=drop_seed)


class _HeadWeightCache:
    """Same policy as _WeightCache: rebuilt on every call that can be followed by a backward pass."""

    def __init__(self):
        self.key, self.buf = None, None

    def invalidate(self):
        self.key = None

    def get(self, cfg, params, force=False):
        key = (cfg.planes, cfg.planes_bwd, cfg.fwd_fp16, cfg.K, cfg.H, cfg.C, cfg.W, cfg.if_pooling) + tuple((p.data_ptr(), p._version) for p in params)
        if force or key != self.key:
            nbytes = _bytes(lib().vmm_head_weights_bytes, cfg)
            if self.buf is None or self.buf.numel() < nbytes or self.buf.device != params[0].device:
                self.buf = torch.empty(nbytes, dtype=torch.uint8, device=params[0].device)
            pp = _head_pointers(params)
            check(lib().vmm_head_prepare(C.byref(cfg), C.byref(pp), ptr(self.buf), stream_ptr()))
            self.key = key
        return self.buf


class _HeadFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, module, want_descriptor, can_backward, h, gamma, *params):
        B = h.shape[0] // module.k
        dev = h.device
        M = gamma.shape[2] if gamma is not None else 1
        cfg = make_head_config(module, B, M)
        with torch.cuda.device(dev):
            wc = module.__dict__.setdefault("_vmm_weight_cache", _HeadWeightCache())
            weights = wc.get(cfg, params, force=can_backward)
            ws = torch.empty(_bytes(lib().vmm_head_workspace_bytes, cfg), dtype=torch.uint8, device=dev)
            out = torch.empty(B, cfg.W if want_descriptor else cfg.C, device=dev)
            pp = _head_pointers(params)
            check(lib().vmm_head_forward(C.byref(cfg), C.byref(pp), ptr(weights), ptr(h), ptr(gamma), ptr(ws),
                                         None if want_descriptor else ptr(out), ptr(out) if want_descriptor else None,
                                         stream_ptr()))
        ctx.cfg, ctx.want_descriptor, ctx.module = cfg, want_descriptor, module
        ctx.save_for_backward(h, ws, weights, *params)
        return out

    @staticmethod
    def backward(ctx, d_out):
        h, ws, weights = ctx.saved_tensors[:3]
        params = ctx.saved_tensors[3:]
        cfg, dev = ctx.cfg, h.device
        with torch.cuda.device(dev):
            scratch = torch.empty(_bytes(lib().vmm_head_scratch_bytes, cfg), dtype=torch.uint8, device=dev)
            live = head_param_tensors(ctx.module)
            grads, direct = _grad_buffers(ctx.module, live)
            gp, pp = _head_pointers(grads), _head_pointers(params)
            d_h = torch.empty_like(h)
            d_out = d_out.contiguous().float()
            check(lib().vmm_head_backward(C.byref(cfg), C.byref(pp), ptr(weights), ptr(h), ptr(ws), ptr(scratch),
                                          None if ctx.want_descriptor else ptr(d_out),
                                          ptr(d_out) if ctx.want_descriptor else None, C.byref(gp), ptr(d_h), stream_ptr()))
        ctx.module.__dict__["_vmm_weight_cache"].invalidate()
        if direct:
            grads = _install(live, grads)
        _after_backward(ctx.module)
        return (None, None, None, d_h, None, *grads)


def head_forward(module, h, gamma):
    """Multi_View_Net.forward: logits (B, nclasses), or the 4096-d descriptor when save_fea == 1."""
    if not h.is_cuda:
        raise RuntimeError("multi_view_sort_b200 has no CPU path: h must be a CUDA tensor")
    h = h.contiguous().float()
    g = None
    if gamma is not None:
        B = h.shape[0] // module.k
        g = gamma.detach().to(h.device, torch.float32).reshape(B, module.k, -1).contiguous()
    params = head_param_tensors(module)
    # grad mode is off inside Function.forward: decide here whether a backward pass (and an optimizer step) can follow
    can_backward = torch.is_grad_enabled() and (h.requires_grad or any(p.requires_grad for p in params))
    return _HeadFunction.apply(module, bool(module.save_fea), can_backward, h, g, *params)


class _CrossEntropyFunction(torch.autograd.Function):
    @staticmethod
    def forward(ctx, logits, labels, scale):
        B, Cn = logits.shape
        with torch.cuda.device(logits.device):
            loss = torch.zeros(1, device=logits.device)
            d_logits = torch.empty_like(logits)
            check(lib().vmm_cross_entropy(B, Cn, ptr(logits), ptr(labels), C.c_float(scale), ptr(loss), ptr(d_logits), None,
                                          stream_ptr()))
        ctx.save_for_backward(d_logits)
        return loss[0]

    @staticmethod
    def backward(ctx, g):
        (d_logits,) = ctx.saved_tensors
        return d_logits * g, None, None


def cross_entropy(logits, labels, scale=None):
    """Mean softmax cross-entropy (nn.CrossEntropyLoss semantics); `scale` overrides 1/B, e.g. 1/B_global
    when the batch is sharded over ranks."""
    logits = logits.contiguous().float()
    labels = labels.to(logits.device, torch.int64).contiguous()
    return _CrossEntropyFunction.apply(logits, labels, float(1.0 / logits.shape[0] if scale is None else scale))
