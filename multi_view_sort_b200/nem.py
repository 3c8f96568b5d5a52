"""Drop-in replacements for the reference's `NEM` and `Multi_View_Net` modules
(train_nem.py:145-294, 296-380), computed by the sm_100a CUDA library.

Same constructor arguments, attributes, `state_dict` keys and tensor layouts as the
reference, so reference checkpoints load and the reference trainer's attribute reads
(tools/Trainer.py:117-130, 276-283) work unchanged.  Parameters live in ordinary torch
modules (device memory + optimizer plumbing); all arithmetic of the hot path is done by
`libvmm_b200.so` through `torch.autograd.Function`s in `functional.py`.
"""
from __future__ import annotations

import torch
import torch.nn as nn

from .model import Model
from . import nem_utilities as U

ENC1_WIDTH = 1024   # train_nem.py:171
ENC2_WIDTH = 4096   # train_nem.py:175
DEC1_WIDTH = 4096   # train_nem.py:181
DEC3_IN = 1024      # train_nem.py:183-185
HEAD_WIDTH = 4096   # VGG-11 classifier width, models/MVCNN.py:55-56


class NEM(Model):
    """Neural-EM latent-view decomposition (train_nem.py:145-294).

    Extra keyword (not in the reference): ``precision`` - a preset of functional.PRECISIONS.  ``"fp32"``: every
    forward contraction on fp16 hi/lo operand pairs (3 tensor-core passes, fp32-class), gradient contractions on two
    bf16 planes - matches the reference's fp32 results to ~1e-5 (1e-4 bar).  ``"bf16"`` (bf16 features, 1e-2 bar): the
    same forward except a single-pass dec3, single-plane bf16 gradient contractions; forward-only calls run every
    contraction in one fp16 pass.  See DESIGN.md 6.
    """

    def __init__(self, nclasses=40, view_num=12, layer_norm=1, k=3, batch_size=32, input_dim=4096,
                 input_type='fc', rnn_hidden_size=1024, iter_num=10, stop_gamma_grad=0, init_type='rand',
                 no_sig=0, bn_init_sigma=0, gamma_prior_loss=0, if_decoder_warm_up=0, if_total_loss=0,
                 arch=0, e_sigma=0.1, precision="fp32"):
        super().__init__('NEM')
        self.nclasses = nclasses
        self.view_dist = 'gaussian'
        self.e_sigma = e_sigma
        self.bn_init_sigma = bn_init_sigma
        self.total_sigma = 0
        self.k = k
        self.b = batch_size
        self.m = view_num
        self.pred_init = 0.0
        self.rnn_hidden_size = rnn_hidden_size
        self.input_dim = input_dim
        self.type = input_type
        self.iter_num = iter_num
        self.stop_gamma_grad = stop_gamma_grad
        self.init_type = init_type
        self.no_sig = no_sig
        self.if_gamma_prior = gamma_prior_loss
        self.if_decoder_warm_up = if_decoder_warm_up
        self.if_total_loss = if_total_loss
        self.arch = arch if self.k > 1 else 0
        self.precision = precision
        if input_type != 'fc':
            raise KeyError("only 'fc' (vector) view features are supported by the CUDA path")

        # Same modules, same construction order as train_nem.py:171-185 -> identical
        # state_dict keys and identical default init under the same torch seed.
        self.enc1 = nn.Sequential(nn.Linear(input_dim, ENC1_WIDTH), nn.LayerNorm(ENC1_WIDTH), nn.ELU(), nn.Dropout(0.5))
        self.enc2 = nn.Sequential(nn.Linear(ENC1_WIDTH * view_num, ENC2_WIDTH), nn.LayerNorm(ENC2_WIDTH), nn.ELU(), nn.Dropout(0.5))
        self.rnn_nem = nn.GRUCell(ENC2_WIDTH, rnn_hidden_size)
        self.after_rnn = nn.Sequential(nn.LayerNorm(rnn_hidden_size))      # constructed, never applied (train_nem.py:179 vs 232-235)
        self.dec1 = nn.Sequential(nn.Linear(rnn_hidden_size, DEC1_WIDTH), nn.LayerNorm(DEC1_WIDTH), nn.ReLU(), nn.Dropout(0.5))
        self.dec2 = nn.Sequential(nn.Linear(DEC1_WIDTH, DEC3_IN * view_num), nn.LayerNorm(DEC3_IN * view_num))
        self.dec3 = nn.Linear(DEC3_IN, input_dim)

    # -- state ---------------------------------------------------------------------------
    def view_tables(self):
        """(K,M) initial responsibilities and (K,M) view prior for this module's init_type."""
        if self.k == 1:
            return torch.ones(1, self.m), torch.ones(1, self.m) / self.m
        return U.view_tables(self.k, self.m, self.init_type)

    def init_state(self, x_shape, x):
        """train_nem.py:188-217: (h, pred, gamma, gamma_prior) on the device of the parameters."""
        B, m, k = x_shape[0], x_shape[1], self.k
        assert m == self.m, "number of view dose not match"
        dev = self.dec3.weight.device
        h = torch.zeros(B * k, self.rnn_hidden_size, device=dev)
        pred = torch.full((B, k, m, self.input_dim), float(self.pred_init), device=dev)
        if k > 1 and self.init_type == 'rand':
            gamma = torch.rand([B, k, m, 1])
            gamma = gamma / gamma.sum(1).unsqueeze(1)
            gamma_prior = gamma
            return h, pred, gamma.to(dev), gamma_prior.to(dev)
        g, p = self.view_tables()
        gamma = g.view(1, k, m, 1).repeat(B, 1, 1, 1).to(dev)
        gamma_prior = p.view(1, k, m, 1).repeat(B, 1, 1, 1).to(dev)
        return h, pred, gamma, gamma_prior

    def init_gamma(self, B, dev):
        """The (gamma, gamma_prior) of init_state as (B, K, M) float32 tensors on `dev`, without the h and pred buffers
        the fused loop never reads (pred alone is B*K*M*D floats: 2.4 GB at configs[1]).  For the deterministic init
        types the (K, M) tables live on the device and are expanded there: no host->device copy per call."""
        k, m = self.k, self.m
        if k > 1 and self.init_type == 'rand':                       # one host draw per call, like the reference
            gamma = torch.rand([B, k, m, 1])
            gamma = (gamma / gamma.sum(1).unsqueeze(1)).to(dev).view(B, k, m)
            return gamma, gamma
        if k == 1 or self.init_type == 'bin_uniform':
            key = (k, m, self.init_type, str(dev))
            cached = self.__dict__.get("_vmm_view_tables")
            if cached is None or cached[0] != key:
                g, p = self.view_tables()
                cached = (key, g.to(dev), p.to(dev))
                self.__dict__["_vmm_view_tables"] = cached
            g, p = cached[1], cached[2]
        else:                                                        # 'bin_rand': python's RNG, drawn per call
            g, p = (t.to(dev) for t in self.view_tables())
        return g.view(1, k, m).expand(B, k, m).contiguous(), p.view(1, k, m).expand(B, k, m).contiguous()

    # -- compute ---------------------------------------------------------------------------
    def invalidate_weights(self):
        """Drop the cached operand planes of the weights.  Needed only after changing parameter VALUES in a way
        autograd cannot see AND without a backward pass through this module in between (e.g. a fused optimizer
        stepping on externally computed gradients, or writes through .data): training calls always rebuild the
        planes and every backward pass marks them stale (functional._WeightCache)."""
        wc = self.__dict__.get("_vmm_weight_cache")
        if wc is not None:
            wc.invalidate()

    def run_em(self, x, iters=None, gamma0=None, gamma_prior=None, labels=None, warm_up=False):
        """Fused path: all `iters` EM iterations and the last iteration's outer loss in one
        autograd node.  x: (B, M, D) on the GPU (fp32, or bf16 when precision='bf16').
        Returns (h_T (B*K,H), gamma_T (B,K,M,1), nem_loss, intra, inter)  - the values the
        reference trainer reads after its loop (tools/Trainer.py:180-203)."""
        from .functional import run_em
        return run_em(self, x, self.iter_num if iters is None else iters, gamma0, gamma_prior, warm_up=warm_up)

    def forward(self, x, state):
        """Step-level drop-in (train_nem.py:272-294): x (B,1,M,D), state=(h,pred,gamma) ->
        (h', pred', gamma').  One EM iteration; differentiable w.r.t. parameters and state (functional.em_step), so the
        reference's trainer loop (tools/Trainer.py:180-219) trains on this module unchanged."""
        from .functional import em_step
        if x.shape[0] != self.b:
            self.b = x.shape[0]
        return em_step(self, x, state)


class Multi_View_Net(Model):
    """Latent-view scoring, sort/concat (or max-pool) and the 3-layer classifier
    (train_nem.py:296-380).  `model` may be an object with a `net_2` Sequential whose
    modules '3' (Linear 4096x4096) is borrowed, as the reference borrows SVCNN's VGG
    classifier; pass None to create the layers here."""

    def __init__(self, model, cnn_name, nclasses=40, k=3, rnn_hidden_size=1024, if_sort=1, if_pooling=0,
                 save_feature=0, precision="fp32"):
        super().__init__('MV_C_Net')
        self.nclasses = nclasses
        self.k = k
        self.cnn_name = cnn_name
        self.rnn_hidden_size = rnn_hidden_size
        self.if_sort = if_sort
        self.if_pooling = if_pooling
        self.save_fea = save_feature
        self.precision = precision
        if not (cnn_name.startswith('vgg')):
            raise KeyError('the CUDA head implements the VGG-style classifier (cnn_name vgg*)')
        in_w = rnn_hidden_size * (k if if_pooling == 0 else 1)
        first = nn.Linear(in_w, HEAD_WIDTH)
        if model is not None and hasattr(model, 'net_2'):
            mid = model.net_2._modules['3']
            last = model.net_2._modules['6']
        else:
            mid = nn.Linear(HEAD_WIDTH, HEAD_WIDTH)
            nn.init.normal_(mid.weight, 0, 0.01)       # torchvision VGG init of classifier Linears
            nn.init.constant_(mid.bias, 0)
            last = nn.Linear(HEAD_WIDTH, nclasses)
        # keys net.0 / net.3 / net.6 as in the reference (VGG classifier indices)
        self.net = nn.Sequential(first, nn.ReLU(True), nn.Dropout(), mid, nn.ReLU(True), nn.Dropout(), last)
        self.att = nn.Sequential(nn.Linear(rnn_hidden_size, 1), nn.Sigmoid())

    def invalidate_weights(self):
        """Drop the cached operand planes of the weights.  Needed only after changing parameter VALUES in a way
        autograd cannot see AND without a backward pass through this module in between (e.g. a fused optimizer
        stepping on externally computed gradients, or writes through .data): training calls always rebuild the
        planes and every backward pass marks them stale (functional._WeightCache)."""
        wc = self.__dict__.get("_vmm_weight_cache")
        if wc is not None:
            wc.invalidate()

    def forward(self, input, gamma):
        """input: (B*K, H) last hidden state, gamma: (B,K,M,1) -> logits (B, nclasses), or
        the 4096-d descriptor (output of the 2nd Linear) when save_feature=1."""
        from .functional import head_forward
        return head_forward(self, input, gamma)
