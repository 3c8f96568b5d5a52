// Internal C++ interfaces between the translation units of libvmm_b200.so.
#pragma once
#include "common.h"

namespace vmm {

// Operand planes: successive bf16 residuals of one matrix (see vmm_planes in the C header).
struct Planes {
  __nv_bfloat16* p[3] = {nullptr, nullptr, nullptr};
  int n = 0;
  Planes first(int k) const {          // the k most significant planes
    Planes r;
    r.n = k < n ? k : n;
    for (int i = 0; i < r.n; ++i) r.p[i] = p[i];
    return r;
  }
};
inline Planes planes_from_c(const vmm_planes* c) {
  Planes r;
  if (c) {
    r.n = c->n;
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
  Planes planes(long long count, int n) {
    Planes p;
    p.n = n;
    for (int i = 0; i < n; ++i) p.p[i] = take<__nv_bfloat16>(count);
    return p;
  }
};

// Cross terms A_i B_j with i + j < max(na, nb), least significant first.  Returns the count.
int make_segs(vmm_gemm_seg* s, const Planes& a, long long lda, const Planes& b, long long ldb, long long k);

int gemm_bf16(int M, int N, int nseg, const vmm_gemm_seg* segs, bool a_mn, bool b_mn, float* C, long long ldc,
              const float* bias, bool accumulate, int splits, int chain_kb, cudaStream_t stream);
int estep_forward(int rows, int D, int zk, const Planes& z, const Planes& w3, const float* b3, const void* x,
                  int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma, float* part_e, float* part_sq,
                  float* part_pp, cudaStream_t stream);
int estep_backward(int rows, int D, int zk, const Planes& z, const Planes& w3, const float* b3, const void* x,
                   int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma, const float* alpha,
                   const float* beta, const float* kappa, const Planes& dpred, float* colsum, cudaStream_t stream);
int split_planes(const float* src, long long n, const Planes& out, cudaStream_t stream);

}  // namespace vmm
