// Memory-bound stages of the EM iteration that sit between the tensor-core contractions:
// LayerNorm + activation (and the enc1 "residual x responsibility" prologue), the GRU gate
// math, the responsibility normalisation, the outer loss, and all of their backward passes.
// Every kernel reads the fp32 output of a contraction and writes the bf16 operand planes the
// next contraction consumes.
#pragma once
#include <cuda_fp16.h>

#include "internal.h"

namespace vmm {

enum { ACT_NONE = 0, ACT_ELU = 1, ACT_RELU = 2 };

// Row-scaled fp16 copy of a gradient tensor for the data-gradient contractions ("fp16 gradient" mode): row r holds
// fp16(v * s_r) with s_r the power of two that puts max|v_r| into [2^11, 2^12) (1 for an all-zero row) and
// inv_scale[r] = 1 / s_r, which the contraction's epilogue multiplies back in (gemm_bf16(..., row_scale)).  Exact
// per-row maxima: no overflow by construction; elements 2^-26 below their row's maximum flush to zero.
struct RowF16 {
  __half* p = nullptr;
  float* inv_scale = nullptr;
};
enum { LN_BIAS = 0, LN_ENC1 = 1 };

// One LayerNorm(+activation) stage.  The pre-normalisation value of element (row, j) is
//   LN_BIAS : C[row, j] + bias[j]                                  (enc2, dec1, dec2)
//   LN_ENC1 : gamma[row] * (xw1[xrow, j] - C[row, j] - w1b3[j]) + bias[j]
//             with xrow = (row / km) * mv + row % mv; C == nullptr on the first EM iteration
//             (pred_0 = 0, train_nem.py:156,197) where the value is gamma * xw1 + bias.
//             This is enc1.Linear applied to (x - pred) * gamma (train_nem.py:280-284, 171)
//             with pred = dec3(z) folded in:  W1 pred = (W1 W3) z + W1 b3.
struct LnDesc {
  int rows, width;
  const float* C;
  long long ldc;
  const float* bias;
  const float* ln_g;
  const float* ln_b;
  const float* xw1;
  const float* w1b3;
  const float* gamma;
  int km, mv;
  float eps;
  // train-mode Dropout after the activation (train_nem.py:172,175,181): see dropout_scale below; kept values
  // are scaled by 1/(1-p).  The mask is recomputed, never stored.  drop_p == 0: eval mode.
  float drop_p;
  unsigned long long drop_seed;
};

// Stateless dropout decisions shared by forward and backward kernels: ONE 32-bit hash per group of four
// consecutive elements (all widths are multiples of 4), one byte per element; an element is kept iff its
// byte >= round(p * 256) (p is quantised to 1/256; 0.5 is exact).
__host__ __device__ inline unsigned dropout_bits(unsigned long long seed, unsigned long long group) {
  unsigned x = static_cast<unsigned>(group) ^ static_cast<unsigned>(seed);
  const unsigned y = static_cast<unsigned>(group >> 32) ^ static_cast<unsigned>(seed >> 32);
  x ^= y * 0x9E3779B9u;
  x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
  return x;
}
__host__ __device__ inline unsigned dropout_thr8(float p) { return static_cast<unsigned>(p * 256.f + 0.5f); }
// scale of element `idx` (its group is idx / 4)
__host__ __device__ inline float dropout_scale(float p, unsigned long long seed, unsigned long long idx) {
  if (p <= 0.f) return 1.f;
  const unsigned bits = dropout_bits(seed, idx >> 2);
  return ((bits >> (8 * (idx & 3))) & 255u) >= dropout_thr8(p) ? 1.f / (1.f - p) : 0.f;
}
// scales of the four elements of group idx / 4 (idx a multiple of 4)
__host__ __device__ inline void dropout_scale4(float p, unsigned long long seed, unsigned long long idx, float (&s)[4]) {
  const unsigned bits = dropout_bits(seed, idx >> 2), thr = dropout_thr8(p);
  const float k = 1.f / (1.f - p);
  s[0] = (bits & 255u) >= thr ? k : 0.f;
  s[1] = ((bits >> 8) & 255u) >= thr ? k : 0.f;
  s[2] = ((bits >> 16) & 255u) >= thr ? k : 0.f;
  s[3] = (bits >> 24) >= thr ? k : 0.f;
}

int ln_forward(int mode, int act, const LnDesc& d, const Op& out, float* mean, float* rstd, cudaStream_t stream);

// Backward through activation + LayerNorm (+ the enc1 prologue).  d_out is the gradient w.r.t.
// the stage output [rows, width] (fp32).  Produces dv = dL/d(pre-normalisation value) as operand
// planes, the per-row LayerNorm reductions c1/c2 (consumed by ln_backward_cols), and for LN_ENC1
// also dgamma[row] (nullable), dxw1 += gamma*dv summed over the K latent rows that share a feature
// row, and dG1 = -gamma*dv as operand planes (nullable).
int ln_backward_rows(int mode, int act, const LnDesc& d, const float* mean, const float* rstd, const float* d_out,
                     long long ldd, const Planes& dv, float* c1, float* c2, float* dgamma, float* dxw1, const Planes& dg1,
                     cudaStream_t stream, RowF16 dvh = RowF16());   // dvh: only the wide LN_BIAS fast path writes it

// Column reductions of the same stage, accumulated (+=) into fp32 parameter gradients:
// d_ln_g += sum_rows dy*xhat, d_ln_b += sum_rows dy, d_bias += sum_rows dv, and for LN_ENC1
// d_w1b3 += sum_rows (-gamma*dv) (nullable).
int ln_backward_cols(int mode, int act, const LnDesc& d, const float* mean, const float* rstd, const float* c1,
                     const float* c2, const float* d_out, long long ldd, float* d_ln_g, float* d_ln_b, float* d_bias,
                     float* d_w1b3, cudaStream_t stream);

// LN_BIAS stages of width 4096 (enc2, dec1): ln_backward_rows + ln_backward_cols in ONE pass over C and d_out.
inline bool lnw_backward_fused_ok(int width) { return width == 4096; }
int lnw_backward_fused(int act, const LnDesc& d, const float* mean, const float* rstd, const float* d_out, long long ldd,
                       const Planes& dv, float* d_ln_g, float* d_ln_b, float* d_bias, cudaStream_t stream,
                       RowF16 dvh = RowF16());

// enc1 (LN_ENC1, width 1024, K <= 6): ln_backward_rows + ln_backward_cols in ONE pass over G1 and d_out.
inline bool enc1_backward_fused_ok(int width, int k_latent) { return width == 1024 && k_latent >= 1 && k_latent <= 6; }
int enc1_backward_fused(int act, const LnDesc& d, const float* mean, const float* rstd, const float* d_out, long long ldd,
                        float* dgamma, float* dxw1, const Planes& dg1, float* d_ln_g, float* d_ln_b, float* d_bias,
                        float* d_w1b3, cudaStream_t stream, RowF16 dg1h = RowF16());

// GRUCell gate math (train_nem.py:177,232; torch.nn.GRUCell, gate order r,z,n).
// gi = e Wih^T, gh = h Whh^T are the raw contraction outputs [rows, 3H] (biases added here).
int gru_forward(int rows, int H, const float* gi, const float* gh, const float* b_ih, const float* b_hh,
                const float* h_prev, float* h_new, const Op& hp, cudaStream_t stream);
// dh is the total gradient w.r.t. h_new.  Writes dgi/dgh operand planes [rows,3H], the direct
// path dh_prev = dh * z (the Whh contraction adds the rest), and accumulates the bias gradients.
int gru_backward(int rows, int H, const float* gi, const float* gh, const float* b_ih, const float* b_hh,
                 const float* h_prev, const float* dh, const Planes& dgi, const Planes& dgh, float* dh_prev,
                 float* d_b_ih, float* d_b_hh, cudaStream_t stream, RowF16 dgih = RowF16(), RowF16 dghh = RowF16());

// Responsibilities from the fused E-step partial sums (train_nem.py:253-270):
//   p = coef * sum_d exp(..) + 1e-10,  gamma = p / sum_k p.   Also folds the partial sums of
//   (x-pred)^2 and pred^2 into per-row totals when given.
int gamma_forward(int B, int K, int Mv, int nblk, float sigma, const float* part_e, const float* part_sq,
                  const float* part_pp, float* gamma, float* inv_s, float* row_sq, float* row_pp, cudaStream_t stream);
// dgamma -> alpha[row] = dL/dp[row] * coef / sigma^2 (the per-row scalar of the E-step backward).
int gamma_backward(int B, int K, int Mv, float sigma, const float* gamma, const float* inv_s, const float* dgamma,
                   float* alpha, cudaStream_t stream);

int row_max_partials(int rows, int nblk, const float* part, float* out, cudaStream_t stream);   // out[r] = max_j part[r, j]

// Outer loss of the last iteration (tools/NEM_utilities.py:98-128).  out[0..2] = total, intra, inter.
int loss_forward(int B, int K, int Mv, int D, int if_gamma_prior, const float* gamma, const float* prior,
                 const float* row_sq, const float* row_pp, float* out3, cudaStream_t stream);
// Gradients of the loss: dgamma += ..., beta/kappa = per-row scalars of dL/dpred (see EPI_EBWD).
// g3 = upstream gradients of (total, intra, inter) on the device.
int loss_backward(int B, int K, int Mv, int D, int if_gamma_prior, int stop_gamma_grad, const float* gamma,
                  const float* prior, const float* row_sq, const float* row_pp, const float* g3, float* dgamma,
                  float* beta, float* kappa, cudaStream_t stream);

// Step-level API: operand planes of the responsibility-weighted residual (x - pred) * gamma
// (train_nem.py:280-284).  x: [B*Mv, D] float or bf16; pred: [rows, D]; gamma: [rows].
int residual_op(long long rows, int D, int km, int mv, const void* x, int x_is_bf16, const float* pred,
                const float* gamma, const Op& out, cudaStream_t stream);

// Its backward: io = dL/dmasked [rows, D] in, dL/dpred = -gamma * dL/dmasked out; dgamma[r] = <dL/dmasked_r, x - pred_r>
// (nullable).
int residual_backward(long long rows, int D, int km, int mv, const void* x, int x_is_bf16, const float* pred,
                      const float* gamma, float* io, float* dgamma, cudaStream_t stream);

// Small dense helpers on fp32 parameters.
int gemv_rows(int n, int k, const float* W, const float* v, float* out, cudaStream_t stream);            // out[i] = W[i,:].v
int rank1_update(int n, int k, const float* u, const float* v, float* W, cudaStream_t stream);            // W[i,j] += u[i] v[j]
int gemv_cols_acc(int n, int k, const float* W, const float* u, float* out, cudaStream_t stream);         // out[j] += sum_i u[i] W[i,j]
int colsum_acc(int rows, int width, const float* src, long long ld, float* out, cudaStream_t stream);     // out[j] += sum_r src[r,j]
int act_planes(int act, long long rows, int width, const float* src, long long ld, const Op& out, float drop_p,
               unsigned long long drop_seed, cudaStream_t stream);                    // planes(dropout(act(src)))
int act_backward(int act, long long n, const float* pre, float* grad, float drop_p, unsigned long long drop_seed,
                 cudaStream_t stream);                                                // grad *= mask * act'(pre)

}  // namespace vmm
