// The EM loop of one batch, forward and backward, as a fixed sequence of tcgen05 contractions
// and streaming stages enqueued on one stream (no host synchronisation, graph-capturable).
//
// Reference: NEM.forward / run_inner_rnn / e_step (train_nem.py:219-294), the trainer's loop and
// loss composition (tools/Trainer.py:160-203), compute_outer_loss (tools/NEM_utilities.py:98-128).
//
// Algebra used (DESIGN.md "dec3 -> residual -> enc1 fold"): enc1's Linear consumes
// gamma*(x - pred) and pred = dec3(z) is itself a bare Linear, so
//     W1 (gamma (x - W3 z - b3)) + b1 = gamma (W1 x - (W1 W3) z - W1 b3) + b1 .
// W1 x is computed once per batch, W13 = W1 W3 and w1b3 = W1 b3 once per optimizer step; pred is
// never materialised: the only consumer left is the E-step, fused into the dec3 contraction.
#include <string.h>

#include <cuda_fp16.h>

#include "internal.h"
#include "stages.cuh"

namespace vmm {
namespace {

constexpr int E1 = 1024;   // enc1 width, dec3 input width (train_nem.py:171,185)
constexpr int E2 = 4096;   // enc2 / dec1 width (train_nem.py:175,181)

struct Dims {
  int B, K, M, D, H, T, planes, pb;   // pb: operand planes of the gradient contractions
  int chain;                          // accumulation chain length in k-blocks (0 = unlimited)
  int f16;                            // forward operands are fp16 pairs
  int chain_b;                        // chain bound of the gradient contractions (single bf16 plane: none needed)
  unsigned modes;                     // VMM_FWD_* of each forward contraction, 2 bits each
  int save_delta;                     // VMM_DELTA_*: keep delta = x - pred of every iteration for the backward pass
  int dgh;                            // row-scaled fp16 operands in the LayerNorm-fed data-gradient contractions
  long long R, RM, BM, ZW;   // B*K, B*K*M, B*M, 1024*M
  int nblk;
  explicit Dims(const vmm_em_config& c)
      : B(c.B), K(c.K), M(c.M), D(c.D), H(c.H), T(c.T), planes(c.planes), pb(c.planes_bwd), chain(c.accum_chain), f16(c.fwd_fp16), chain_b(c.planes_bwd >= 2 ? c.accum_chain : 0), modes(c.fwd_modes), save_delta(c.training ? c.save_delta : 0), dgh(c.training && c.dgrad_fp16 && c.fwd_fp16 && enc1_backward_fused_ok(1024, c.K) ? 1 : 0), R(1ll * c.B * c.K),
        RM(1ll * c.B * c.K * c.M), BM(1ll * c.B * c.M), ZW(1ll * E1 * c.M), nblk(2 * ((c.D + 255) / 256)) {}
};

struct Weights {
  Op w1, w2, wih, whh, w4, w5, w3, w13;
  float* w13f;
  float* w1b3;
  long long bytes;
  Weights(const vmm_em_config& c, void* base) {
    Dims d(c);
    Bump b(base);
    w1 = b.op(1ll * E1 * d.D, d.planes, d.f16, d.pb);
    w2 = b.op(1ll * E2 * d.ZW, d.planes, d.f16, d.pb);
    wih = b.op(3ll * d.H * E2, d.planes, d.f16, d.pb);
    whh = b.op(3ll * d.H * d.H, d.planes, d.f16, d.pb);
    w4 = b.op(1ll * E2 * d.H, d.planes, d.f16, d.pb);
    w5 = b.op(d.ZW * E2, d.planes, d.f16, d.pb);
    w3 = b.op(1ll * d.D * E1, d.planes, d.f16, d.pb);
    w13 = b.op(1ll * E1 * E1, d.planes, d.f16, d.pb);
    w13f = b.take<float>(1ll * E1 * E1);
    w1b3 = b.take<float>(E1);
    bytes = b.off + 256;
  }
};

struct IterBufs {
  float *g1, *mean1, *rstd1;
  Op a;
  float *g2, *mean2, *rstd2;
  Op e;
  float *gi, *gh, *h;
  Op hp;
  float *g4, *mean4, *rstd4;
  Op u;
  float *g5, *mean5, *rstd5;
  Op z;
  float *gamma, *inv_s;
  void* delta;       // x - pred of this iteration [RM, D] (fp16 or fp32), only with save_delta in training
  float* dmax;       // max_d |x - pred| per row [RM], only with dgrad_fp16 in training
};

struct Workspace {
  Op x;              // operand planes of the features
  float* xw1;        // [B*M, 1024]
  float* h0;         // zeros [R, H]
  float *row_sq, *row_pp;
  int nslots;
  long long bytes;
  long long slot_bytes, slot_base;
  char* base;
  Dims d;
  Workspace(const vmm_em_config& c, void* wsbase) : base(static_cast<char*>(wsbase)), d(c) {
    Bump b(wsbase);
    if (!c.x_is_bf16) x = b.op(d.BM * d.D, d.planes, d.f16, d.pb);
    else if (d.f16) x.f = b.planes(d.BM * d.D, 1, VMM_FMT_F16PAIR);   // bf16 features up-cast to one exact fp16 plane
    xw1 = b.take<float>(d.BM * E1);
    h0 = b.take<float>(d.R * d.H);
    row_sq = b.take<float>(d.RM);
    row_pp = b.take<float>(d.RM);
    b.off = (b.off + 255) & ~255ll;
    slot_base = b.off;
    Bump s(nullptr);
    layout_slot(s, nullptr);
    slot_bytes = (s.off + 255) & ~255ll;
    nslots = c.training ? c.T : 2;
    bytes = slot_base + slot_bytes * nslots + 256;
  }
  void layout_slot(Bump& b, IterBufs* o) const {
    IterBufs t;
    t.g1 = b.take<float>(d.RM * E1);
    t.mean1 = b.take<float>(d.RM);
    t.rstd1 = b.take<float>(d.RM);
    t.a = b.op(d.RM * E1, d.planes, d.f16, d.pb);
    t.g2 = b.take<float>(d.R * E2);
    t.mean2 = b.take<float>(d.R);
    t.rstd2 = b.take<float>(d.R);
    t.e = b.op(d.R * E2, d.planes, d.f16, d.pb);
    t.gi = b.take<float>(d.R * 3 * d.H);
    t.gh = b.take<float>(d.R * 3 * d.H);
    t.h = b.take<float>(d.R * d.H);
    t.hp = b.op(d.R * d.H, d.planes, d.f16, d.pb);
    t.g4 = b.take<float>(d.R * E2);
    t.mean4 = b.take<float>(d.R);
    t.rstd4 = b.take<float>(d.R);
    t.u = b.op(d.R * E2, d.planes, d.f16, d.pb);
    t.g5 = b.take<float>(d.R * d.ZW);
    t.mean5 = b.take<float>(d.R);
    t.rstd5 = b.take<float>(d.R);
    t.z = b.op(d.RM * E1, d.planes, d.f16, d.pb);
    t.gamma = b.take<float>(d.RM);
    t.inv_s = b.take<float>(d.BM);
    t.delta = nullptr;
    if (d.save_delta == VMM_DELTA_F16) t.delta = b.take<__half>(d.RM * d.D);
    else if (d.save_delta == VMM_DELTA_F32) t.delta = b.take<float>(d.RM * d.D);
    t.dmax = (d.dgh && d.save_delta) ? b.take<float>(d.RM) : nullptr;
    if (o) *o = t;
  }
  IterBufs slot(int iter) const {
    const int s = (nslots >= d.T) ? iter : (iter & 1);
    Bump b(base ? base + slot_base + slot_bytes * s : nullptr);
    IterBufs t;
    layout_slot(b, &t);
    return t;
  }
};

struct Scratch {
  float *part_e, *part_sq, *part_pp, *part_dmax;
  // backward
  float *dh, *dh_prev, *dgamma, *alpha, *beta, *kappa, *c1, *c2;
  Planes dpred, dv5, dv4, dgi, dgh, dv2, dg1, dxw1p, dw13p;
  float *dz, *du, *de, *da, *dxw1, *dw13, *dw1b3;
  RowF16 dv5h, dv4h, dv2h, dg1h, dgih, dghh, dpredh;     // row-scaled fp16 copies (dgrad_fp16)
  float *dpred_gmax = nullptr, *dpred_inv_cols = nullptr;                    // dpredh carries ONE scale: the bound of max|dpred| and 1/scale per column
  long long bytes;
  Scratch(const vmm_em_config& c, void* base) {
    Dims d(c);
    Bump b(base);
    part_e = b.take<float>(d.RM * d.nblk);
    part_sq = b.take<float>(d.RM * d.nblk);
    part_pp = b.take<float>(d.RM * d.nblk);
    part_dmax = d.dgh ? b.take<float>(d.RM * d.nblk) : nullptr;
    if (c.training) {
      dh = b.take<float>(d.R * d.H);
      dh_prev = b.take<float>(d.R * d.H);
      dgamma = b.take<float>(d.RM);
      alpha = b.take<float>(d.RM);
      beta = b.take<float>(d.RM);
      kappa = b.take<float>(d.RM);
      c1 = b.take<float>(d.RM);
      c2 = b.take<float>(d.RM);
      dpred = b.planes(d.RM * d.D, d.pb);
      dz = b.take<float>(d.RM * E1);
      dv5 = b.planes(d.R * d.ZW, d.pb);
      du = b.take<float>(d.R * E2);
      dv4 = b.planes(d.R * E2, d.pb);
      dgi = b.planes(d.R * 3 * d.H, d.pb);
      dgh = b.planes(d.R * 3 * d.H, d.pb);
      de = b.take<float>(d.R * E2);
      dv2 = b.planes(d.R * E2, d.pb);
      da = b.take<float>(d.R * d.ZW);
      dg1 = b.planes(d.RM * E1, d.pb);
      dxw1 = b.take<float>(d.BM * E1);
      dxw1p = b.planes(d.BM * E1, d.pb);
      dw13 = b.take<float>(1ll * E1 * E1);
      dw13p = b.planes(1ll * E1 * E1, d.pb);
      dw1b3 = b.take<float>(E1);
      if (d.dgh) {
        dv5h.p = b.take<__half>(d.R * d.ZW); dv5h.inv_scale = b.take<float>(d.R);
        dv4h.p = b.take<__half>(d.R * E2);   dv4h.inv_scale = b.take<float>(d.R);
        dv2h.p = b.take<__half>(d.R * E2);   dv2h.inv_scale = b.take<float>(d.R);
        dg1h.p = b.take<__half>(d.RM * E1);  dg1h.inv_scale = b.take<float>(d.RM);
        if (d.save_delta) {
          dpredh.p = b.take<__half>(d.RM * d.D); dpredh.inv_scale = b.take<float>(d.RM);
          dpred_gmax = b.take<float>(1); dpred_inv_cols = b.take<float>(d.D);
        }
        if (d.H <= 1024) {
          dgih.p = b.take<__half>(d.R * 3 * d.H); dgih.inv_scale = b.take<float>(d.R);
          dghh.p = b.take<__half>(d.R * 3 * d.H); dghh.inv_scale = b.take<float>(d.R);
        }
      }
    }
    bytes = b.off + 256;
  }
};

int check_cfg(const vmm_em_config* c) {
  VMM_REQUIRE(c != nullptr, "null config");
  VMM_REQUIRE(c->B > 0 && c->K > 0 && c->M > 0 && c->D > 0 && c->H > 0 && c->T > 0, "non-positive dimension");
  VMM_REQUIRE(c->D % 8 == 0 && c->H % 8 == 0, "D and H must be multiples of 8");
  VMM_REQUIRE(c->planes >= 1 && c->planes <= 3, "planes must be 1, 2 or 3");
  VMM_REQUIRE(c->planes_bwd >= 1 && c->planes_bwd <= 3, "planes_bwd must be 1, 2 or 3");
  VMM_REQUIRE(c->fwd_fp16 ? c->planes <= 2 : c->planes_bwd <= c->planes, "fp16 pairs have at most 2 planes; bf16 planes_bwd <= planes");
  VMM_REQUIRE(c->if_gamma_prior >= 0 && c->if_gamma_prior <= 2, "if_gamma_prior must be 0, 1 or 2");
  VMM_REQUIRE(c->sigma > 0.f && c->ln_eps > 0.f, "sigma and ln_eps must be positive");
  VMM_REQUIRE(c->accum_chain >= 0, "accum_chain must be >= 0");
  VMM_REQUIRE(c->dropout_p >= 0.f && c->dropout_p < 1.f, "dropout_p must be in [0, 1)");
  VMM_REQUIRE(c->save_delta >= VMM_DELTA_NONE && c->save_delta <= VMM_DELTA_F32, "save_delta must be 0, 1 or 2");
  VMM_REQUIRE(!c->dgrad_fp16 || c->fwd_fp16, "dgrad_fp16 needs fwd_fp16 (the fp16 hi planes of the weights)");
  VMM_REQUIRE((c->fwd_modes >> 20) == 0, "fwd_modes has 9 two-bit fields and 2 gradient-contraction bits");
  VMM_REQUIRE((c->warm_up == 0 || c->warm_up == 1) && (c->loss_all == 0 || c->loss_all == 1), "warm_up / loss_all must be 0 or 1");
  VMM_REQUIRE(1ll * c->B * c->K * c->M < (1ll << 31) / 2, "too many rows for 32-bit indexing; split the batch");
  return 0;
}

// A/B bits 18, 19 of vmm_em_config.fwd_modes: with planes_bwd >= 2, run only the weight-gradient (18) or only the
// data-gradient (19) contractions on the leading plane - separates the two error sources of single-plane gradients
// (tools/mode_sweep.py).  Set per call by em_backward (thread-local: the library has no other per-call state).
static thread_local unsigned g_bwd_bits = 0;

// forward contraction ids (bit positions of vmm_em_config.fwd_modes / 2)
enum { FC_XW1 = 0, FC_G1 = 1, FC_ENC2 = 2, FC_GI = 3, FC_GH = 4, FC_DEC1 = 5, FC_DEC2 = 6, FC_DEC3 = 7, FC_EBWD = 8 };

inline Planes sel_act(const Dims& d, int fc, const Planes& a) {
  const unsigned m = (d.modes >> (2 * fc)) & 3u;
  return (m == VMM_FWD_SINGLE || m == VMM_FWD_W_PAIR) ? a.first(1) : a;
}
inline Planes sel_w(const Dims& d, int fc, const Planes& w) {
  const unsigned m = (d.modes >> (2 * fc)) & 3u;
  return (m == VMM_FWD_SINGLE || m == VMM_FWD_ACT_PAIR) ? w.first(1) : w;
}

// The operand planes a streaming stage has to WRITE for activation `o`: in training everything (forward planes +
// the bf16 set of the gradient contractions); in forward-only runs no gradient set, and only the leading forward
// plane when every contraction that reads the activation runs single-pass.
inline Op fwd_out(const Dims& d, bool training, const Op& o, int fc_a, int fc_b = -1) {
  if (training) return o;
  Op r = o;
  r.b = Planes();
  auto single = [&](int fc) {
    const unsigned m = (d.modes >> (2 * fc)) & 3u;
    return m == VMM_FWD_SINGLE || m == VMM_FWD_W_PAIR;
  };
  if (single(fc_a) && (fc_b < 0 || single(fc_b))) r.f = o.f.first(1);
  return r;
}

// y[rows, n] = a[rows, k] * w[n, k]^T        (nn.Linear forward; bias is added by the consumer stage)
int linear_fwd(const Dims& d, int fc, const Planes& a_all, const Planes& w_all, long long rows, int n, long long k, float* y,
               cudaStream_t st) {
  const Planes a = sel_act(d, fc, a_all), w = sel_w(d, fc, w_all);
  vmm_gemm_seg s[9];
  const int ns = make_segs(s, a, k, w, k, k);
  return gemm_bf16(static_cast<int>(rows), n, ns, s, false, false, y, n, nullptr, false, 1, d.chain, st);
}
// dx[rows, k] (+)= dy[rows, n] * w[n, k]     (data gradient; w is read N-major, no transposed copy)
int linear_dgrad(const Planes& dy_all, const Planes& w_all, long long rows, int n, long long k, float* dx, bool accumulate,
                 int chain, cudaStream_t st) {
  const Planes dy = g_bwd_bits & VMM_BWD_DGRAD_SINGLE ? dy_all.first(1) : dy_all;
  const Planes w = g_bwd_bits & VMM_BWD_DGRAD_SINGLE ? w_all.first(1) : w_all;
  vmm_gemm_seg s[9];
  const int ns = make_segs(s, dy, n, w, k, n);
  return gemm_bf16(static_cast<int>(rows), static_cast<int>(k), ns, s, false, true, dx, k, nullptr, accumulate, 1, chain, st);
}
// dx (+)= diag(inv_scale) * (dyh[rows, n] * w_hi[n, k]) : the same with a row-scaled fp16 gradient and the fp16 hi
// plane of the weight (one tensor-core pass, 11 mantissa bits on both sides)
int linear_dgrad_h(const RowF16& dyh, const Planes& w_f, long long rows, int n, long long k, float* dx, bool accumulate,
                   cudaStream_t st) {
  Planes a;
  a.n = 1;
  a.fmt = VMM_FMT_F16PAIR;
  a.p[0] = reinterpret_cast<__nv_bfloat16*>(dyh.p);
  vmm_gemm_seg s[9];
  const int ns = make_segs(s, a, n, w_f.first(1), k, n);
  return gemm_bf16(static_cast<int>(rows), static_cast<int>(k), ns, s, false, true, dx, k, nullptr, accumulate, 1, 0, st,
                   dyh.inv_scale);
}
// dw[n, k] += dy[rows, n]^T * x[rows, k]     (weight gradient; both operands read MN-major, split-K)
int linear_wgrad(const Planes& dy_all, const Planes& x_all, long long rows, int n, long long k, float* dw, int chain,
                 cudaStream_t st) {
  const Planes dy = g_bwd_bits & VMM_BWD_WGRAD_SINGLE ? dy_all.first(1) : dy_all;
  const Planes x = g_bwd_bits & VMM_BWD_WGRAD_SINGLE ? x_all.first(1) : x_all;
  vmm_gemm_seg s[9];
  const int ns = make_segs(s, dy, n, x, k, rows);
  return gemm_bf16(n, static_cast<int>(k), ns, s, true, true, dw, k, nullptr, true, 0, chain, st);
}

// stage ids for the dropout stream
enum { ST_ENC1 = 1, ST_ENC2 = 2, ST_DEC1 = 3, ST_DEC2 = 4 };

LnDesc ln_desc(long long rows, long long width, const float* C, const float* bias, const float* g, const float* b,
               float eps, const vmm_em_config* c = nullptr, int stage = 0, int iter = 0) {
  LnDesc d;
  memset(&d, 0, sizeof(d));
  d.rows = static_cast<int>(rows);
  d.width = static_cast<int>(width);
  d.C = C;
  d.ldc = width;
  d.bias = bias;
  d.ln_g = g;
  d.ln_b = b;
  d.eps = eps;
  if (c && c->dropout_p > 0.f && stage != ST_DEC2) {      // dec2 has no Dropout (train_nem.py:183)
    d.drop_p = c->dropout_p;
    d.drop_seed = c->dropout_seed + (static_cast<unsigned long long>(stage) << 56) + (static_cast<unsigned long long>(iter) << 40);
  }
  return d;
}

}  // namespace

// ============================================================================================
int em_prepare(const vmm_em_config* c, const vmm_em_params* p, void* weights, void* scratch, cudaStream_t st) {
  VMM_TRY(check_cfg(c));
  VMM_REQUIRE(p && weights, "null argument");
  (void)scratch;
  Dims d(*c);
  Weights w(*c, weights);
  VMM_TRY(split_op(p->enc1_w, 1ll * E1 * d.D, w.w1, st));
  VMM_TRY(split_op(p->enc2_w, 1ll * E2 * d.ZW, w.w2, st));
  VMM_TRY(split_op(p->gru_w_ih, 3ll * d.H * E2, w.wih, st));
  VMM_TRY(split_op(p->gru_w_hh, 3ll * d.H * d.H, w.whh, st));
  VMM_TRY(split_op(p->dec1_w, 1ll * E2 * d.H, w.w4, st));
  VMM_TRY(split_op(p->dec2_w, d.ZW * E2, w.w5, st));
  VMM_TRY(split_op(p->dec3_w, 1ll * d.D * E1, w.w3, st));
  // W13[i, j] = sum_d W1[i, d] W3[d, j]:  A = W1 (K-major over d), B[n=j][k=d] = W3[d][j] (N-major)
  {
    vmm_gemm_seg s[9];
    const int ns = make_segs(s, w.w1.f, d.D, w.w3.f, E1, d.D);
    VMM_TRY(gemm_bf16(E1, E1, ns, s, false, true, w.w13f, E1, nullptr, false, 1, d.chain, st));
    VMM_TRY(split_op(w.w13f, 1ll * E1 * E1, w.w13, st));
  }
  VMM_TRY(gemv_rows(E1, d.D, p->enc1_w, p->dec3_b, w.w1b3, st));
  return 0;
}

int em_forward(const vmm_em_config* c, const vmm_em_params* p, const void* weights, const void* x, const float* gamma0,
               const float* prior, void* workspace, void* scratch, float* h_out, float* gamma_out, float* loss_out,
               cudaStream_t st) {
  VMM_TRY(check_cfg(c));
  VMM_REQUIRE(p && weights && x && gamma0 && workspace && scratch, "null argument");
  VMM_REQUIRE(!loss_out || c->if_gamma_prior == 0 || prior, "prior is required for if_gamma_prior 1/2");
  VMM_REQUIRE(!c->warm_up || prior, "warm_up weights every residual by `prior`");
  Dims d(*c);
  Weights w(*c, const_cast<void*>(weights));
  Workspace ws(*c, workspace);
  Scratch sc(*c, scratch);
  const float eps = c->ln_eps;
  const bool tr = c->training != 0;

  // operand planes of the features and W1 x (once per batch)
  Op xp = ws.x;
  if (c->x_is_bf16) {
    xp.b = Planes();
    xp.b.n = 1;
    xp.b.p[0] = static_cast<__nv_bfloat16*>(const_cast<void*>(x));   // bf16 features are their own (exact) plane
    if (d.f16) VMM_TRY(bf16_to_f16(x, d.BM * d.D, xp.f.p[0], st));
    else xp.f = xp.b;
  } else {
    VMM_TRY(split_op(static_cast<const float*>(x), d.BM * d.D, xp, st));
  }
  VMM_TRY(linear_fwd(d, FC_XW1, xp.f, w.w1.f, d.BM, E1, d.D, ws.xw1, st));
  VMM_CHECK_CUDA(cudaMemsetAsync(ws.h0, 0, sizeof(float) * d.R * d.H, st));

  for (int t = 0; t < d.T; ++t) {
    const IterBufs cur = ws.slot(t);
    const IterBufs prev = ws.slot(t > 0 ? t - 1 : 0);
    const bool last = (t == d.T - 1);
    const float* gam_prev = c->warm_up ? prior : (t > 0 ? prev.gamma : gamma0);
    // --- enc1 on the responsibility-weighted residual (train_nem.py:280-284, 226)
    if (t > 0) VMM_TRY(linear_fwd(d, FC_G1, prev.z.f, w.w13.f, d.RM, E1, E1, cur.g1, st));
    {
      LnDesc ln = ln_desc(d.RM, E1, t > 0 ? cur.g1 : nullptr, p->enc1_b, p->enc1_ln_g, p->enc1_ln_b, eps, c, ST_ENC1, t);
      ln.xw1 = ws.xw1;
      ln.w1b3 = w.w1b3;
      ln.gamma = gam_prev;
      ln.km = d.K * d.M;
      ln.mv = d.M;
      VMM_TRY(ln_forward(LN_ENC1, ACT_ELU, ln, fwd_out(d, tr, cur.a, FC_ENC2), cur.mean1, cur.rstd1, st));
    }
    // --- enc2 (train_nem.py:228-229)
    VMM_TRY(linear_fwd(d, FC_ENC2, cur.a.f, w.w2.f, d.R, E2, d.ZW, cur.g2, st));
    VMM_TRY(ln_forward(LN_BIAS, ACT_ELU, ln_desc(d.R, E2, cur.g2, p->enc2_b, p->enc2_ln_g, p->enc2_ln_b, eps, c, ST_ENC2, t),
                       fwd_out(d, tr, cur.e, FC_GI), cur.mean2, cur.rstd2, st));
    // --- GRU (train_nem.py:232)
    VMM_TRY(linear_fwd(d, FC_GI, cur.e.f, w.wih.f, d.R, 3 * d.H, E2, cur.gi, st));
    if (t > 0) VMM_TRY(linear_fwd(d, FC_GH, prev.hp.f, w.whh.f, d.R, 3 * d.H, d.H, cur.gh, st));
    else VMM_CHECK_CUDA(cudaMemsetAsync(cur.gh, 0, sizeof(float) * d.R * 3 * d.H, st));
    VMM_TRY(gru_forward(static_cast<int>(d.R), d.H, cur.gi, cur.gh, p->gru_b_ih, p->gru_b_hh, t > 0 ? prev.h : ws.h0, cur.h,
                        fwd_out(d, tr, cur.hp, FC_GH, FC_DEC1), st));
    // --- dec1, dec2 (train_nem.py:238-240)
    VMM_TRY(linear_fwd(d, FC_DEC1, cur.hp.f, w.w4.f, d.R, E2, d.H, cur.g4, st));
    VMM_TRY(ln_forward(LN_BIAS, ACT_RELU, ln_desc(d.R, E2, cur.g4, p->dec1_b, p->dec1_ln_g, p->dec1_ln_b, eps, c, ST_DEC1, t),
                       fwd_out(d, tr, cur.u, FC_DEC2), cur.mean4, cur.rstd4, st));
    VMM_TRY(linear_fwd(d, FC_DEC2, cur.u.f, w.w5.f, d.R, static_cast<int>(d.ZW), E2, cur.g5, st));
    VMM_TRY(ln_forward(LN_BIAS, ACT_NONE, ln_desc(d.R, d.ZW, cur.g5, p->dec2_b, p->dec2_ln_g, p->dec2_ln_b, eps),
                       fwd_out(d, tr, cur.z, FC_DEC3, FC_G1), cur.mean5, cur.rstd5, st));
    // --- dec3 fused with the E-step (train_nem.py:241-270)
    const bool want_loss = loss_out && (last || c->loss_all);     // the row sums of the LAST iteration stay in ws.row_* for backward
    const bool want_sq = last || want_loss;
    const bool want_pp = want_sq && c->if_gamma_prior == 0;
    VMM_TRY(estep_forward(static_cast<int>(d.RM), d.D, E1, sel_act(d, FC_DEC3, cur.z.f), sel_w(d, FC_DEC3, w.w3.f), p->dec3_b, x,
                          c->x_is_bf16, d.D, d.K, d.M, c->sigma, sc.part_e, want_sq ? sc.part_sq : nullptr,
                          want_pp ? sc.part_pp : nullptr, cur.delta, d.save_delta, st, cur.dmax ? sc.part_dmax : nullptr));
    if (cur.dmax) VMM_TRY(row_max_partials(static_cast<int>(d.RM), d.nblk, sc.part_dmax, cur.dmax, st));
    VMM_TRY(gamma_forward(d.B, d.K, d.M, d.nblk, c->sigma, sc.part_e, want_sq ? sc.part_sq : nullptr,
                          want_pp ? sc.part_pp : nullptr, cur.gamma, cur.inv_s, ws.row_sq, ws.row_pp, st));
    if (want_loss)
      VMM_TRY(loss_forward(d.B, d.K, d.M, d.D, c->if_gamma_prior, cur.gamma, prior, ws.row_sq, ws.row_pp,
                           loss_out + (c->loss_all ? 3 * t : 0), st));
    if (last) {
      if (h_out) VMM_CHECK_CUDA(cudaMemcpyAsync(h_out, cur.h, sizeof(float) * d.R * d.H, cudaMemcpyDeviceToDevice, st));
      if (gamma_out)
        VMM_CHECK_CUDA(cudaMemcpyAsync(gamma_out, cur.gamma, sizeof(float) * d.RM, cudaMemcpyDeviceToDevice, st));
    }
  }
  return 0;
}

// ============================================================================================
// Step-level API: ONE EM iteration from an arbitrary state, exactly NEM.forward(x, (h, pred, gamma))
// (train_nem.py:272-294).  Unlike the fused loop it must materialise `pred` (the reference returns it),
// so enc1 runs on the explicit residual instead of the W1*W3 fold.  Forward only.
struct StepScratch {
  Op masked, hin;
  float* part_e;
  long long bytes;
  StepScratch(const vmm_em_config& c, void* base) {
    Dims d(c);
    Bump b(base);
    masked = b.op(d.RM * d.D, d.planes, d.f16, d.pb);
    hin = b.op(d.R * d.H, d.planes, d.f16, d.pb);
    part_e = b.take<float>(d.RM * d.nblk);
    bytes = b.off + 256;
  }
};

int em_step_forward(const vmm_em_config* c, const vmm_em_params* p, const void* weights, const void* x, const float* h_in,
                    const float* pred_in, const float* gamma_in, void* workspace, void* scratch, float* h_out,
                    float* pred_out, float* gamma_out, cudaStream_t st) {
  VMM_TRY(check_cfg(c));
  VMM_REQUIRE(p && weights && x && h_in && pred_in && gamma_in && workspace && scratch && h_out && pred_out && gamma_out,
              "null argument");
  VMM_REQUIRE(!c->training || c->save_delta == VMM_DELTA_F32, "the step-level API trains with save_delta = VMM_DELTA_F32");
  vmm_em_config c1 = *c;
  c1.T = 1;                      // one iteration: one workspace slot (kept for vmm_em_step_backward when training)
  Dims d(c1);
  Weights w(c1, const_cast<void*>(weights));
  Workspace ws(c1, workspace);
  StepScratch sc(c1, scratch);
  const IterBufs cur = ws.slot(0);
  const float eps = c->ln_eps;
  const bool tr = c->training != 0;
  VMM_TRY(residual_op(d.RM, d.D, d.K * d.M, d.M, x, c->x_is_bf16, pred_in, gamma_in, sc.masked, st));
  VMM_TRY(linear_fwd(d, FC_XW1, sc.masked.f, w.w1.f, d.RM, E1, d.D, cur.g1, st));
  VMM_TRY(ln_forward(LN_BIAS, ACT_ELU, ln_desc(d.RM, E1, cur.g1, p->enc1_b, p->enc1_ln_g, p->enc1_ln_b, eps, c, ST_ENC1, 0), cur.a,
                     cur.mean1, cur.rstd1, st));
  VMM_TRY(linear_fwd(d, FC_ENC2, cur.a.f, w.w2.f, d.R, E2, d.ZW, cur.g2, st));
  VMM_TRY(ln_forward(LN_BIAS, ACT_ELU, ln_desc(d.R, E2, cur.g2, p->enc2_b, p->enc2_ln_g, p->enc2_ln_b, eps, c, ST_ENC2, 0), cur.e,
                     cur.mean2, cur.rstd2, st));
  VMM_TRY(linear_fwd(d, FC_GI, cur.e.f, w.wih.f, d.R, 3 * d.H, E2, cur.gi, st));
  VMM_TRY(split_op(h_in, d.R * d.H, sc.hin, st));
  VMM_TRY(linear_fwd(d, FC_GH, sc.hin.f, w.whh.f, d.R, 3 * d.H, d.H, cur.gh, st));
  VMM_TRY(gru_forward(static_cast<int>(d.R), d.H, cur.gi, cur.gh, p->gru_b_ih, p->gru_b_hh, h_in, h_out, cur.hp, st));
  VMM_TRY(linear_fwd(d, FC_DEC1, cur.hp.f, w.w4.f, d.R, E2, d.H, cur.g4, st));
  VMM_TRY(ln_forward(LN_BIAS, ACT_RELU, ln_desc(d.R, E2, cur.g4, p->dec1_b, p->dec1_ln_g, p->dec1_ln_b, eps, c, ST_DEC1, 0), cur.u,
                     cur.mean4, cur.rstd4, st));
  VMM_TRY(linear_fwd(d, FC_DEC2, cur.u.f, w.w5.f, d.R, static_cast<int>(d.ZW), E2, cur.g5, st));
  VMM_TRY(ln_forward(LN_BIAS, ACT_NONE, ln_desc(d.R, d.ZW, cur.g5, p->dec2_b, p->dec2_ln_g, p->dec2_ln_b, eps), cur.z, cur.mean5,
                     cur.rstd5, st));
  {
    const Planes z = sel_act(d, FC_DEC3, cur.z.f), w3 = sel_w(d, FC_DEC3, w.w3.f);
    vmm_gemm_seg s[9];
    const int ns = make_segs(s, z, E1, w3, E1, E1);
    VMM_TRY(gemm_bf16(static_cast<int>(d.RM), d.D, ns, s, false, false, pred_out, d.D, p->dec3_b, false, 1, d.chain, st));
    VMM_TRY(estep_forward(static_cast<int>(d.RM), d.D, E1, z, w3, p->dec3_b, x, c->x_is_bf16, d.D, d.K, d.M, c->sigma, sc.part_e,
                          nullptr, nullptr, tr ? cur.delta : nullptr, tr ? d.save_delta : VMM_DELTA_NONE, st));
  }
  VMM_TRY(gamma_forward(d.B, d.K, d.M, d.nblk, c->sigma, sc.part_e, nullptr, nullptr, cur.gamma, cur.inv_s, nullptr, nullptr, st));
  VMM_CHECK_CUDA(cudaMemcpyAsync(gamma_out, cur.gamma, sizeof(float) * d.RM, cudaMemcpyDeviceToDevice, st));
  return 0;
}

// Backward of ONE vmm_em_step_forward call (cfg->training = 1): the building block of autograd through the reference's
// own loop `for j: state = nem(input, state)` (tools/Trainer.py:180-198).  Upstream gradients d_h_out / d_pred_out /
// d_gamma_out (any may be NULL = zero) -> parameter gradients (+=) and d_h_in / d_pred_in / d_gamma_in.
int em_step_backward(const vmm_em_config* c, const vmm_em_params* p, const void* weights, const void* x, const float* h_in,
                     const float* pred_in, const float* gamma_in, const void* workspace, const void* step_scratch,
                     void* scratch, const float* d_h_out, const float* d_pred_out, const float* d_gamma_out,
                     const vmm_em_grads* g, float* d_h_in, float* d_pred_in, float* d_gamma_in, cudaStream_t st) {
  VMM_TRY(check_cfg(c));
  VMM_REQUIRE(c->training && c->save_delta == VMM_DELTA_F32, "vmm_em_step_backward needs a step forward with training = 1, save_delta = F32");
  VMM_REQUIRE(p && weights && x && h_in && pred_in && gamma_in && workspace && step_scratch && scratch && g, "null argument");
  VMM_REQUIRE(!d_gamma_in || d_pred_in, "d_gamma_in needs the d_pred_in buffer (it holds dL/dmasked on the way)");
  vmm_em_config c1 = *c;
  c1.T = 1;
  Dims d(c1);
  Weights w(c1, const_cast<void*>(weights));
  Workspace ws(c1, const_cast<void*>(workspace));
  StepScratch ss(c1, const_cast<void*>(step_scratch));
  Scratch sc(c1, scratch);
  const IterBufs cur = ws.slot(0);
  const float eps = c->ln_eps;
  const int R = static_cast<int>(d.R);
  g_bwd_bits = 0;

  if (d_h_out) VMM_CHECK_CUDA(cudaMemcpyAsync(sc.dh, d_h_out, sizeof(float) * d.R * d.H, cudaMemcpyDeviceToDevice, st));
  else VMM_CHECK_CUDA(cudaMemsetAsync(sc.dh, 0, sizeof(float) * d.R * d.H, st));
  if (d_gamma_out) VMM_CHECK_CUDA(cudaMemcpyAsync(sc.dgamma, d_gamma_out, sizeof(float) * d.RM, cudaMemcpyDeviceToDevice, st));
  else VMM_CHECK_CUDA(cudaMemsetAsync(sc.dgamma, 0, sizeof(float) * d.RM, st));
  VMM_CHECK_CUDA(cudaMemsetAsync(sc.beta, 0, sizeof(float) * d.RM, st));
  // --- E-step (train_nem.py:247-270) and the caller's explicit dL/dpred
  VMM_TRY(gamma_backward(d.B, d.K, d.M, c->sigma, cur.gamma, cur.inv_s, sc.dgamma, sc.alpha, st));
  VMM_TRY(estep_dpred(static_cast<int>(d.RM), d.D, cur.delta, d.save_delta, x, c->x_is_bf16, d.D, d.K, d.M, c->sigma, sc.alpha,
                      sc.beta, nullptr, sc.dpred, g->dec3_b, st, nullptr, nullptr, nullptr, d_pred_out));
  // --- dec3, dec2, dec1
  VMM_TRY(linear_wgrad(sc.dpred, cur.z.b, d.RM, d.D, E1, g->dec3_w, d.chain_b, st));
  VMM_TRY(linear_dgrad(sc.dpred, w.w3.b, d.RM, d.D, E1, sc.dz, false, d.chain_b, st));
  {
    LnDesc ln = ln_desc(d.R, d.ZW, cur.g5, p->dec2_b, p->dec2_ln_g, p->dec2_ln_b, eps);
    VMM_TRY(ln_backward_rows(LN_BIAS, ACT_NONE, ln, cur.mean5, cur.rstd5, sc.dz, d.ZW, sc.dv5, sc.c1, sc.c2, nullptr, nullptr,
                             Planes(), st));
    VMM_TRY(ln_backward_cols(LN_BIAS, ACT_NONE, ln, cur.mean5, cur.rstd5, sc.c1, sc.c2, sc.dz, d.ZW, g->dec2_ln_g, g->dec2_ln_b,
                             g->dec2_b, nullptr, st));
  }
  VMM_TRY(linear_dgrad(sc.dv5, w.w5.b, d.R, static_cast<int>(d.ZW), E2, sc.du, false, d.chain_b, st));
  VMM_TRY(linear_wgrad(sc.dv5, cur.u.b, d.R, static_cast<int>(d.ZW), E2, g->dec2_w, d.chain_b, st));
  {
    LnDesc ln = ln_desc(d.R, E2, cur.g4, p->dec1_b, p->dec1_ln_g, p->dec1_ln_b, eps, c, ST_DEC1, 0);
    VMM_TRY(ln_backward_rows(LN_BIAS, ACT_RELU, ln, cur.mean4, cur.rstd4, sc.du, E2, sc.dv4, sc.c1, sc.c2, nullptr, nullptr,
                             Planes(), st));
    VMM_TRY(ln_backward_cols(LN_BIAS, ACT_RELU, ln, cur.mean4, cur.rstd4, sc.c1, sc.c2, sc.du, E2, g->dec1_ln_g, g->dec1_ln_b,
                             g->dec1_b, nullptr, st));
  }
  VMM_TRY(linear_dgrad(sc.dv4, w.w4.b, d.R, E2, d.H, sc.dh, true, d.chain_b, st));
  VMM_TRY(linear_wgrad(sc.dv4, cur.hp.b, d.R, E2, d.H, g->dec1_w, d.chain_b, st));
  // --- GRU
  VMM_TRY(gru_backward(R, d.H, cur.gi, cur.gh, p->gru_b_ih, p->gru_b_hh, h_in, sc.dh, sc.dgi, sc.dgh, sc.dh_prev, g->gru_b_ih,
                       g->gru_b_hh, st));
  VMM_TRY(linear_dgrad(sc.dgi, w.wih.b, d.R, 3 * d.H, E2, sc.de, false, d.chain_b, st));
  VMM_TRY(linear_wgrad(sc.dgi, cur.e.b, d.R, 3 * d.H, E2, g->gru_w_ih, d.chain_b, st));
  VMM_TRY(linear_dgrad(sc.dgh, w.whh.b, d.R, 3 * d.H, d.H, sc.dh_prev, true, d.chain_b, st));
  VMM_TRY(linear_wgrad(sc.dgh, ss.hin.b, d.R, 3 * d.H, d.H, g->gru_w_hh, d.chain_b, st));
  if (d_h_in) VMM_CHECK_CUDA(cudaMemcpyAsync(d_h_in, sc.dh_prev, sizeof(float) * d.R * d.H, cudaMemcpyDeviceToDevice, st));
  // --- enc2
  {
    LnDesc ln = ln_desc(d.R, E2, cur.g2, p->enc2_b, p->enc2_ln_g, p->enc2_ln_b, eps, c, ST_ENC2, 0);
    VMM_TRY(ln_backward_rows(LN_BIAS, ACT_ELU, ln, cur.mean2, cur.rstd2, sc.de, E2, sc.dv2, sc.c1, sc.c2, nullptr, nullptr,
                             Planes(), st));
    VMM_TRY(ln_backward_cols(LN_BIAS, ACT_ELU, ln, cur.mean2, cur.rstd2, sc.c1, sc.c2, sc.de, E2, g->enc2_ln_g, g->enc2_ln_b,
                             g->enc2_b, nullptr, st));
  }
  VMM_TRY(linear_dgrad(sc.dv2, w.w2.b, d.R, E2, d.ZW, sc.da, false, d.chain_b, st));
  VMM_TRY(linear_wgrad(sc.dv2, cur.a.b, d.R, E2, d.ZW, g->enc2_w, d.chain_b, st));
  // --- enc1 on the explicit residual: G1 = masked W1^T, no fold
  {
    LnDesc ln = ln_desc(d.RM, E1, cur.g1, p->enc1_b, p->enc1_ln_g, p->enc1_ln_b, eps, c, ST_ENC1, 0);
    VMM_TRY(ln_backward_rows(LN_BIAS, ACT_ELU, ln, cur.mean1, cur.rstd1, sc.da, E1, sc.dg1, sc.c1, sc.c2, nullptr, nullptr,
                             Planes(), st));
    VMM_TRY(ln_backward_cols(LN_BIAS, ACT_ELU, ln, cur.mean1, cur.rstd1, sc.c1, sc.c2, sc.da, E1, g->enc1_ln_g, g->enc1_ln_b,
                             g->enc1_b, nullptr, st));
  }
  VMM_TRY(linear_wgrad(sc.dg1, ss.masked.b, d.RM, E1, d.D, g->enc1_w, d.chain_b, st));
  // --- residual: dL/dmasked = dv1 W1 -> dL/dpred_in = -gamma dL/dmasked, dL/dgamma_in = <dL/dmasked, x - pred_in>
  if (d_pred_in) {
    VMM_TRY(linear_dgrad(sc.dg1, w.w1.b, d.RM, E1, d.D, d_pred_in, false, d.chain_b, st));
    VMM_TRY(residual_backward(d.RM, d.D, d.K * d.M, d.M, x, c->x_is_bf16, pred_in, gamma_in, d_pred_in, d_gamma_in, st));
  }
  return 0;
}

int em_backward(const vmm_em_config* c, const vmm_em_params* p, const void* weights, const void* x, const float* gamma0,
                const float* prior, const void* workspace, void* scratch, const float* d_h, const float* d_gamma,
                const float* d_loss, const vmm_em_grads* g, cudaStream_t st) {
  VMM_TRY(check_cfg(c));
  VMM_REQUIRE(c->training, "vmm_em_backward needs a forward run with training = 1");
  VMM_REQUIRE(p && weights && x && gamma0 && workspace && scratch && d_loss && g, "null argument");
  Dims d(*c);
  Weights w(*c, const_cast<void*>(weights));
  Workspace ws(*c, const_cast<void*>(workspace));
  Scratch sc(*c, scratch);
  const float eps = c->ln_eps;
  const int R = static_cast<int>(d.R);
  g_bwd_bits = c->fwd_modes & (VMM_BWD_WGRAD_SINGLE | VMM_BWD_DGRAD_SINGLE);

  Op xp = ws.x;
  if (c->x_is_bf16) {
    xp.b = Planes();
    xp.b.n = 1;
    xp.b.p[0] = static_cast<__nv_bfloat16*>(const_cast<void*>(x));
  }
  if (d_h) VMM_CHECK_CUDA(cudaMemcpyAsync(sc.dh, d_h, sizeof(float) * d.R * d.H, cudaMemcpyDeviceToDevice, st));
  else VMM_CHECK_CUDA(cudaMemsetAsync(sc.dh, 0, sizeof(float) * d.R * d.H, st));
  if (d_gamma) VMM_CHECK_CUDA(cudaMemcpyAsync(sc.dgamma, d_gamma, sizeof(float) * d.RM, cudaMemcpyDeviceToDevice, st));
  else VMM_CHECK_CUDA(cudaMemsetAsync(sc.dgamma, 0, sizeof(float) * d.RM, st));
  VMM_CHECK_CUDA(cudaMemsetAsync(sc.dxw1, 0, sizeof(float) * d.BM * E1, st));
  VMM_CHECK_CUDA(cudaMemsetAsync(sc.dw13, 0, sizeof(float) * E1 * E1, st));
  VMM_CHECK_CUDA(cudaMemsetAsync(sc.dw1b3, 0, sizeof(float) * E1, st));

  float* dh = sc.dh;
  float* dh_prev = sc.dh_prev;
  for (int t = d.T - 1; t >= 0; --t) {
    const IterBufs cur = ws.slot(t);
    const IterBufs prev = ws.slot(t > 0 ? t - 1 : 0);
    const bool last = (t == d.T - 1);
    const float* gam_prev = c->warm_up ? prior : (t > 0 ? prev.gamma : gamma0);
    const bool gam_grad = t > 0 && !c->warm_up;      // the residual's weight is the previous iteration's gamma: it gets a gradient
    const float* kappa = nullptr;
    bool dv5_h = false;
    // --- loss and E-step: per-row scalars of dL/dpred
    if (last) {
      VMM_TRY(loss_backward(d.B, d.K, d.M, d.D, c->if_gamma_prior, c->stop_gamma_grad, cur.gamma, prior, ws.row_sq, ws.row_pp,
                            d_loss, sc.dgamma, sc.beta, sc.kappa, st));
      if (c->if_gamma_prior == 0) kappa = sc.kappa;
    } else {
      VMM_CHECK_CUDA(cudaMemsetAsync(sc.beta, 0, sizeof(float) * d.RM, st));
    }
    VMM_TRY(gamma_backward(d.B, d.K, d.M, c->sigma, cur.gamma, cur.inv_s, sc.dgamma, sc.alpha, st));
    const bool dpred_h = d.dgh && d.save_delta && !kappa && sc.dpredh.p && cur.dmax;
    if (d.save_delta)
      VMM_TRY(estep_dpred(static_cast<int>(d.RM), d.D, cur.delta, d.save_delta, x, c->x_is_bf16, d.D, d.K, d.M, c->sigma,
                          sc.alpha, sc.beta, kappa, dpred_h ? Planes() : sc.dpred, g->dec3_b, st, dpred_h ? sc.dpredh.p : nullptr,
                          dpred_h ? sc.dpredh.inv_scale : nullptr, dpred_h ? cur.dmax : nullptr, nullptr,
                          dpred_h ? sc.dpred_gmax : nullptr, dpred_h ? sc.dpred_inv_cols : nullptr));
    else
      VMM_TRY(estep_backward(static_cast<int>(d.RM), d.D, E1, sel_act(d, FC_EBWD, cur.z.f), sel_w(d, FC_EBWD, w.w3.f), p->dec3_b, x,
                             c->x_is_bf16, d.D, d.K, d.M, c->sigma, sc.alpha, sc.beta, kappa, sc.dpred, g->dec3_b, st));
    // --- dec3: weight gradient; dz = dpred W3 (+ dG1_{t+1} W13, the path through the next residual)
    if (dpred_h) {
      // dW3 += dpred^T z with the globally scaled fp16 dpred and the fp16 hi plane of z: output row d scaled by 1/S
      Planes a;
      a.n = 1;
      a.fmt = VMM_FMT_F16PAIR;
      a.p[0] = reinterpret_cast<__nv_bfloat16*>(sc.dpredh.p);
      vmm_gemm_seg s[9];
      const int ns = make_segs(s, a, d.D, cur.z.f.first(1), E1, d.RM);
      VMM_TRY(gemm_bf16(d.D, E1, ns, s, true, true, g->dec3_w, E1, nullptr, true, 0, 0, st, sc.dpred_inv_cols));
    } else {
      VMM_TRY(linear_wgrad(sc.dpred, cur.z.b, d.RM, d.D, E1, g->dec3_w, d.chain_b, st));
    }
    {
      vmm_gemm_seg s[18];
      const bool one = (g_bwd_bits & VMM_BWD_DGRAD_SINGLE) != 0;
      int ns = make_segs(s, one ? sc.dpred.first(1) : sc.dpred, d.D, one ? w.w3.b.first(1) : w.w3.b, E1, d.D);
      if (!last && !d.dgh) ns += make_segs(s + ns, one ? sc.dg1.first(1) : sc.dg1, E1, one ? w.w13.b.first(1) : w.w13.b, E1, E1);
      if (dpred_h) VMM_TRY(linear_dgrad_h(sc.dpredh, w.w3.f, d.RM, d.D, E1, sc.dz, false, st));
      else VMM_TRY(gemm_bf16(static_cast<int>(d.RM), E1, ns, s, false, true, sc.dz, E1, nullptr, false, 1, d.chain_b, st));
      if (!last && d.dgh) VMM_TRY(linear_dgrad_h(sc.dg1h, w.w13.f, d.RM, E1, E1, sc.dz, true, st));   // own row scales: own launch
    }
    // --- dec2: LayerNorm(1024*M) backward, dgrad, wgrad
    {
      LnDesc ln = ln_desc(d.R, d.ZW, cur.g5, p->dec2_b, p->dec2_ln_g, p->dec2_ln_b, eps);
      const bool wide = d.dgh && d.ZW != E1 && d.ZW % 512 == 0 && d.ZW / 16 <= 1024;    // the kernel that can write dv5h
      dv5_h = wide;
      VMM_TRY(ln_backward_rows(LN_BIAS, ACT_NONE, ln, cur.mean5, cur.rstd5, sc.dz, d.ZW, sc.dv5, sc.c1, sc.c2, nullptr, nullptr,
                               Planes(), st, wide ? sc.dv5h : RowF16()));
      VMM_TRY(ln_backward_cols(LN_BIAS, ACT_NONE, ln, cur.mean5, cur.rstd5, sc.c1, sc.c2, sc.dz, d.ZW, g->dec2_ln_g,
                               g->dec2_ln_b, g->dec2_b, nullptr, st));
    }
    if (dv5_h) VMM_TRY(linear_dgrad_h(sc.dv5h, w.w5.f, d.R, static_cast<int>(d.ZW), E2, sc.du, false, st));
    else VMM_TRY(linear_dgrad(sc.dv5, w.w5.b, d.R, static_cast<int>(d.ZW), E2, sc.du, false, d.chain_b, st));
    VMM_TRY(linear_wgrad(sc.dv5, cur.u.b, d.R, static_cast<int>(d.ZW), E2, g->dec2_w, d.chain_b, st));
    // --- dec1
    {
      LnDesc ln = ln_desc(d.R, E2, cur.g4, p->dec1_b, p->dec1_ln_g, p->dec1_ln_b, eps, c, ST_DEC1, t);
      if (lnw_backward_fused_ok(E2)) {
        VMM_TRY(lnw_backward_fused(ACT_RELU, ln, cur.mean4, cur.rstd4, sc.du, E2, sc.dv4, g->dec1_ln_g, g->dec1_ln_b,
                                   g->dec1_b, st, d.dgh ? sc.dv4h : RowF16()));
      } else {
        VMM_TRY(ln_backward_rows(LN_BIAS, ACT_RELU, ln, cur.mean4, cur.rstd4, sc.du, E2, sc.dv4, sc.c1, sc.c2, nullptr, nullptr,
                                 Planes(), st));
        VMM_TRY(ln_backward_cols(LN_BIAS, ACT_RELU, ln, cur.mean4, cur.rstd4, sc.c1, sc.c2, sc.du, E2, g->dec1_ln_g,
                                 g->dec1_ln_b, g->dec1_b, nullptr, st));
      }
    }
    if (d.dgh && lnw_backward_fused_ok(E2)) VMM_TRY(linear_dgrad_h(sc.dv4h, w.w4.f, d.R, E2, d.H, dh, true, st));
    else VMM_TRY(linear_dgrad(sc.dv4, w.w4.b, d.R, E2, d.H, dh, true, d.chain_b, st));          // dh += dG4 W4
    VMM_TRY(linear_wgrad(sc.dv4, cur.hp.b, d.R, E2, d.H, g->dec1_w, d.chain_b, st));
    // --- GRU
    const bool gru_h = d.dgh && sc.dgih.p != nullptr;
    VMM_TRY(gru_backward(R, d.H, cur.gi, cur.gh, p->gru_b_ih, p->gru_b_hh, t > 0 ? prev.h : ws.h0, dh, sc.dgi, sc.dgh,
                         dh_prev, g->gru_b_ih, g->gru_b_hh, st, gru_h ? sc.dgih : RowF16(), gru_h ? sc.dghh : RowF16()));
    if (gru_h) VMM_TRY(linear_dgrad_h(sc.dgih, w.wih.f, d.R, 3 * d.H, E2, sc.de, false, st));
    else VMM_TRY(linear_dgrad(sc.dgi, w.wih.b, d.R, 3 * d.H, E2, sc.de, false, d.chain_b, st));
    VMM_TRY(linear_wgrad(sc.dgi, cur.e.b, d.R, 3 * d.H, E2, g->gru_w_ih, d.chain_b, st));
    if (t > 0) {
      if (gru_h) VMM_TRY(linear_dgrad_h(sc.dghh, w.whh.f, d.R, 3 * d.H, d.H, dh_prev, true, st));
      else VMM_TRY(linear_dgrad(sc.dgh, w.whh.b, d.R, 3 * d.H, d.H, dh_prev, true, d.chain_b, st));   // dh_{t-1} += dgh Whh
      VMM_TRY(linear_wgrad(sc.dgh, prev.hp.b, d.R, 3 * d.H, d.H, g->gru_w_hh, d.chain_b, st));
    }
    {
      float* tmp = dh;
      dh = dh_prev;
      dh_prev = tmp;
    }
    // --- enc2
    {
      LnDesc ln = ln_desc(d.R, E2, cur.g2, p->enc2_b, p->enc2_ln_g, p->enc2_ln_b, eps, c, ST_ENC2, t);
      if (lnw_backward_fused_ok(E2)) {
        VMM_TRY(lnw_backward_fused(ACT_ELU, ln, cur.mean2, cur.rstd2, sc.de, E2, sc.dv2, g->enc2_ln_g, g->enc2_ln_b,
                                   g->enc2_b, st, d.dgh ? sc.dv2h : RowF16()));
      } else {
        VMM_TRY(ln_backward_rows(LN_BIAS, ACT_ELU, ln, cur.mean2, cur.rstd2, sc.de, E2, sc.dv2, sc.c1, sc.c2, nullptr, nullptr,
                                 Planes(), st));
        VMM_TRY(ln_backward_cols(LN_BIAS, ACT_ELU, ln, cur.mean2, cur.rstd2, sc.c1, sc.c2, sc.de, E2, g->enc2_ln_g,
                                 g->enc2_ln_b, g->enc2_b, nullptr, st));
      }
    }
    if (d.dgh && lnw_backward_fused_ok(E2)) VMM_TRY(linear_dgrad_h(sc.dv2h, w.w2.f, d.R, E2, d.ZW, sc.da, false, st));
    else VMM_TRY(linear_dgrad(sc.dv2, w.w2.b, d.R, E2, d.ZW, sc.da, false, d.chain_b, st));
    VMM_TRY(linear_wgrad(sc.dv2, cur.a.b, d.R, E2, d.ZW, g->enc2_w, d.chain_b, st));
    // --- enc1 + residual: dgamma_{t-1}, dXW1, dG1 (-> dz_{t-1} and dW13)
    {
      LnDesc ln = ln_desc(d.RM, E1, t > 0 ? cur.g1 : nullptr, p->enc1_b, p->enc1_ln_g, p->enc1_ln_b, eps, c, ST_ENC1, t);
      ln.xw1 = ws.xw1;
      ln.w1b3 = w.w1b3;
      ln.gamma = gam_prev;
      ln.km = d.K * d.M;
      ln.mv = d.M;
      if (enc1_backward_fused_ok(E1, d.K)) {
        VMM_TRY(enc1_backward_fused(ACT_ELU, ln, cur.mean1, cur.rstd1, sc.da, E1, gam_grad ? sc.dgamma : nullptr, sc.dxw1,
                                    t > 0 ? sc.dg1 : Planes(), g->enc1_ln_g, g->enc1_ln_b, g->enc1_b,
                                    t > 0 ? sc.dw1b3 : nullptr, st, (t > 0 && d.dgh) ? sc.dg1h : RowF16()));
      } else {
        VMM_TRY(ln_backward_rows(LN_ENC1, ACT_ELU, ln, cur.mean1, cur.rstd1, sc.da, E1, Planes(), sc.c1, sc.c2,
                                 gam_grad ? sc.dgamma : nullptr, sc.dxw1, t > 0 ? sc.dg1 : Planes(), st));
        VMM_TRY(ln_backward_cols(LN_ENC1, ACT_ELU, ln, cur.mean1, cur.rstd1, sc.c1, sc.c2, sc.da, E1, g->enc1_ln_g,
                                 g->enc1_ln_b, g->enc1_b, t > 0 ? sc.dw1b3 : nullptr, st));
      }
    }
    if (t > 0) VMM_TRY(linear_wgrad(sc.dg1, prev.z.b, d.RM, E1, E1, sc.dw13, d.chain_b, st));
    if (t > 0 && c->warm_up) VMM_CHECK_CUDA(cudaMemsetAsync(sc.dgamma, 0, sizeof(float) * d.RM, st));   // gamma_{t-1} feeds nothing
  }
  // --- parameters behind the fold: XW1 = x W1^T, W13 = W1 W3, w1b3 = W1 b3
  VMM_TRY(split_planes(sc.dxw1, d.BM * E1, sc.dxw1p, st));
  VMM_TRY(linear_wgrad(sc.dxw1p, xp.b, d.BM, E1, d.D, g->enc1_w, d.chain_b, st));                       // dW1 += dXW1^T x
  VMM_TRY(split_planes(sc.dw13, 1ll * E1 * E1, sc.dw13p, st));
  {
    vmm_gemm_seg s[9];
    // dW1[i, d] += sum_j dW13[i, j] W3[d, j]
    int ns = make_segs(s, sc.dw13p, E1, w.w3.b, E1, E1);
    VMM_TRY(gemm_bf16(E1, d.D, ns, s, false, false, g->enc1_w, d.D, nullptr, true, 0, d.chain_b, st));
    // dW3[d, j] += sum_i W1[i, d] dW13[i, j]
    ns = make_segs(s, w.w1.b, d.D, sc.dw13p, E1, E1);
    VMM_TRY(gemm_bf16(d.D, E1, ns, s, true, true, g->dec3_w, E1, nullptr, true, 0, d.chain_b, st));
  }
  VMM_TRY(rank1_update(E1, d.D, sc.dw1b3, p->dec3_b, g->enc1_w, st));                       // dW1 += dw1b3 b3^T
  VMM_TRY(gemv_cols_acc(E1, d.D, p->enc1_w, sc.dw1b3, g->dec3_b, st));                      // db3 += W1^T dw1b3
  return 0;
}

}  // namespace vmm

// ============================== extern "C" ==================================================
extern "C" {

long long vmm_em_weights_bytes(const vmm_em_config* cfg) {
  if (vmm::check_cfg(cfg) != 0) return VMM_E_INVALID;
  return vmm::Weights(*cfg, nullptr).bytes;
}
long long vmm_em_workspace_bytes(const vmm_em_config* cfg) {
  if (vmm::check_cfg(cfg) != 0) return VMM_E_INVALID;
  return vmm::Workspace(*cfg, nullptr).bytes;
}
long long vmm_em_scratch_bytes(const vmm_em_config* cfg) {
  if (vmm::check_cfg(cfg) != 0) return VMM_E_INVALID;
  return vmm::Scratch(*cfg, nullptr).bytes;
}
int vmm_em_prepare(const vmm_em_config* cfg, const vmm_em_params* params, void* weights, void* scratch, void* stream) {
  return vmm::em_prepare(cfg, params, weights, scratch, static_cast<cudaStream_t>(stream));
}
int vmm_em_forward(const vmm_em_config* cfg, const vmm_em_params* params, const void* weights, const void* x,
                   const float* gamma0, const float* prior, void* workspace, void* scratch, float* h_out, float* gamma_out,
                   float* loss_out, void* stream) {
  return vmm::em_forward(cfg, params, weights, x, gamma0, prior, workspace, scratch, h_out, gamma_out, loss_out,
                         static_cast<cudaStream_t>(stream));
}
int vmm_em_backward(const vmm_em_config* cfg, const vmm_em_params* params, const void* weights, const void* x,
                    const float* gamma0, const float* prior, const void* workspace, void* scratch, const float* d_h,
                    const float* d_gamma, const float* d_loss, const vmm_em_grads* grads, void* stream) {
  vmm::BackwardScope backward;
  return vmm::em_backward(cfg, params, weights, x, gamma0, prior, workspace, scratch, d_h, d_gamma, d_loss, grads,
                          static_cast<cudaStream_t>(stream));
}
long long vmm_em_step_scratch_bytes(const vmm_em_config* cfg) {
  if (vmm::check_cfg(cfg) != 0) return VMM_E_INVALID;
  return vmm::StepScratch(*cfg, nullptr).bytes;
}
int vmm_em_step_forward(const vmm_em_config* cfg, const vmm_em_params* params, const void* weights, const void* x,
                        const float* h_in, const float* pred_in, const float* gamma_in, void* workspace, void* scratch,
                        float* h_out, float* pred_out, float* gamma_out, void* stream) {
  return vmm::em_step_forward(cfg, params, weights, x, h_in, pred_in, gamma_in, workspace, scratch, h_out, pred_out,
                              gamma_out, static_cast<cudaStream_t>(stream));
}
int vmm_em_step_backward(const vmm_em_config* cfg, const vmm_em_params* params, const void* weights, const void* x,
                         const float* h_in, const float* pred_in, const float* gamma_in, const void* workspace,
                         const void* step_scratch, void* scratch, const float* d_h_out, const float* d_pred_out,
                         const float* d_gamma_out, const vmm_em_grads* grads, float* d_h_in, float* d_pred_in, float* d_gamma_in,
                         void* stream) {
  vmm::BackwardScope backward;
  return vmm::em_step_backward(cfg, params, weights, x, h_in, pred_in, gamma_in, workspace, step_scratch, scratch, d_h_out,
                               d_pred_out, d_gamma_out, grads, d_h_in, d_pred_in, d_gamma_in, static_cast<cudaStream_t>(stream));
}
const void* vmm_em_workspace_ptr(const vmm_em_config* cfg, const void* workspace, const char* name, int iter) {
  if (vmm::check_cfg(cfg) != 0 || !name || iter < 0 || iter >= cfg->T) return nullptr;
  vmm::Workspace ws(*cfg, const_cast<void*>(workspace));
  const vmm::IterBufs b = ws.slot(iter);
  struct { const char* n; const void* p; } tab[] = {
      {"xw1", ws.xw1}, {"g1", b.g1}, {"a_hi", b.a.f.p[0]}, {"g2", b.g2}, {"e_hi", b.e.f.p[0]}, {"gi", b.gi}, {"gh", b.gh},
      {"h", b.h}, {"h_hi", b.hp.f.p[0]}, {"g4", b.g4}, {"u_hi", b.u.f.p[0]}, {"g5", b.g5}, {"z_hi", b.z.f.p[0]}, {"z_lo", b.z.f.p[1]},
      {"gamma", b.gamma}, {"row_sq", ws.row_sq}, {"mean4", b.mean4}, {"rstd4", b.rstd4}, {"mean2", b.mean2}, {"rstd2", b.rstd2},
      {"delta", b.delta}};
  for (auto& e : tab)
    if (strcmp(e.n, name) == 0) return e.p;
  return nullptr;
}

}  // extern "C"
