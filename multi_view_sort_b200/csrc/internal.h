// Internal C++ interfaces between the translation units of libvmm_b200.so.
#pragma once
#include "common.h"

namespace vmm {

int gemm_bf16(int M, int N, int nseg, const vmm_gemm_seg* segs, bool a_mn, bool b_mn, float* C, long long ldc,
              const float* bias, bool accumulate, int splits, cudaStream_t stream);
int estep_forward(int rows, int D, int zk, int planes, const void* z_hi, const void* z_lo, const void* w3_hi,
                  const void* w3_lo, const float* b3, const void* x, int x_is_bf16, long long ldx, int k_latent,
                  int m_views, float sigma, float* part_e, float* part_sq, float* part_pp, cudaStream_t stream);
int estep_backward(int rows, int D, int zk, int planes, const void* z_hi, const void* z_lo, const void* w3_hi,
                   const void* w3_lo, const float* b3, const void* x, int x_is_bf16, long long ldx, int k_latent,
                   int m_views, float sigma, const float* alpha, const float* beta, const float* kappa,
                   void* dpred_hi, void* dpred_lo, float* colsum, cudaStream_t stream);
int split_planes(const float* src, long long n, void* hi, void* lo, cudaStream_t stream);

}  // namespace vmm
