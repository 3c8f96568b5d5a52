// Internal C++ interfaces between the translation units of libvmm_b200.so.
#pragma once
#include <cuda_fp16.h>

#include "common.h"

namespace vmm {

// Operand planes: successive bf16 residuals of one matrix (see vmm_planes in the C header).
struct Planes {
  __nv_bfloat16* p[3] = {nullptr, nullptr, nullptr};   // fp16 bits when fmt == VMM_FMT_F16PAIR
  int n = 0;
  int fmt = VMM_FMT_BF16;
  Planes first(int k) const {          // the k most significant planes
    Planes r;
    r.n = k < n ? k : n;
    r.fmt = fmt;
    for (int i = 0; i < r.n; ++i) r.p[i] = p[i];
    return r;
  }
};
// A forward tensor that is consumed both by forward contractions (f: bf16 residual planes or an fp16
// pair) and by gradient contractions (b: bf16 planes; kind::f16 cannot mix bf16 and fp16 operands).
// With bf16 forward planes, b aliases the leading planes of f.
struct Op {
  Planes f, b;
};
inline Planes planes_from_c(const vmm_planes* c) {
  Planes r;
  if (c) {
    r.n = c->n;
    r.fmt = c->fmt;
    for (int i = 0; i < 3; ++i) r.p[i] = i < c->n ? static_cast<__nv_bfloat16*>(c->p[i]) : nullptr;
  }
  return r;
}
// Bump allocator over a caller-owned buffer (base == nullptr: size query only); 256-byte aligned.
struct Bump {
  char* base;
  long long off = 0;
  explicit Bump(void* b) : base(static_cast<char*>(b)) {}
  template <typename T>
  T* take(long long count) {
    off = (off + 255) & ~255ll;
    T* p = base ? reinterpret_cast<T*>(base + off) : nullptr;
    off += count * static_cast<long long>(sizeof(T));
    return p;
  }
  Planes planes(long long count, int n, int fmt = VMM_FMT_BF16) {
    Planes p;
    p.n = n;
    p.fmt = fmt;
    for (int i = 0; i < n; ++i) p.p[i] = take<__nv_bfloat16>(count);
    return p;
  }
  // forward operand with n_f planes (fp16 pair when fwd_fp16) + n_b bf16 planes for the gradient contractions
  Op op(long long count, int n_f, int fwd_fp16, int n_b) {
    Op o;
    if (fwd_fp16) {
      o.f = planes(count, n_f, VMM_FMT_F16PAIR);
      o.b = planes(count, n_b, VMM_FMT_BF16);
    } else {
      o.f = planes(count, n_f > n_b ? n_f : n_b, VMM_FMT_BF16);
      o.b = o.f.first(n_b);
      o.f = o.f.first(n_f);
    }
    return o;
  }
};

// Cross terms A_i B_j with i + j < max(na, nb), least significant first.  Returns the count.
int make_segs(vmm_gemm_seg* s, const Planes& a, long long lda, const Planes& b, long long ldb, long long k);

// Backward passes run inside a BackwardScope: contraction launches of the calling thread may then split the tiles of a
// partial last wave along K (gemm.cu: plan_tail_split).  Forward launches never do: their results stay bit-identical
// under a permutation of the batch and from run to run (the backward pass accumulates with atomics anyway).
struct BackwardScope {
  BackwardScope();
  ~BackwardScope();
  BackwardScope(const BackwardScope&) = delete;
  BackwardScope& operator=(const BackwardScope&) = delete;
};
int gemm_bf16(int M, int N, int nseg, const vmm_gemm_seg* segs, bool a_mn, bool b_mn, float* C, long long ldc,
              const float* bias, bool accumulate, int splits, int chain_kb, cudaStream_t stream,
              const float* row_scale = nullptr);   // row_scale[m] multiplies output row m (un-scales a row-scaled A)
int estep_forward(int rows, int D, int zk, const Planes& z, const Planes& w3, const float* b3, const void* x,
                  int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma, float* part_e, float* part_sq,
                  float* part_pp, void* delta, int delta_fmt, cudaStream_t stream, float* part_dmax = nullptr);
int estep_backward(int rows, int D, int zk, const Planes& z, const Planes& w3, const float* b3, const void* x,
                   int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma, const float* alpha,
                   const float* beta, const float* kappa, const Planes& dpred, float* colsum, cudaStream_t stream);
// Streaming E-step backward from the saved delta = x - pred (stages.cu).
int estep_dpred(int rows, int D, const void* delta, int delta_fmt, const void* x, int x_is_bf16, long long ldx,
                int k_latent, int m_views, float sigma, const float* alpha, const float* beta, const float* kappa,
                const Planes& dpred, float* colsum, cudaStream_t stream, __half* dpred_h = nullptr,
                float* inv_scale = nullptr, const float* row_dmax = nullptr,    // + row-scaled fp16 copy (kappa == nullptr only)
                const float* add = nullptr,                                     // + an explicit dL/dpred term [rows, D]
                float* gmax = nullptr, float* inv_scale_cols = nullptr);        // ONE scale for the tensor: fp16 copy only, no bf16 plane
int split_planes(const float* src, long long n, const Planes& out, cudaStream_t stream);
int split_op(const float* src, long long n, const Op& out, cudaStream_t stream);   // both operand sets
int bf16_to_f16(const void* src_bf16, long long n, void* dst_f16, cudaStream_t stream);

}  // namespace vmm
