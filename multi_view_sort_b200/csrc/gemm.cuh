// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//
//     D[M,N] = sum_seg  A_seg[M,K_seg] * B_seg[N,K_seg]^T        (bf16 in, fp32 accumulate in TMEM)
//
// The K loop runs over up to VMM_MAX_SEG "segments", each with its own pair of TMA tensor
// maps.  Segments are how the library gets fp32-class accuracy out of bf16 tensor cores
// (error-compensated split: A = A_hi + A_lo, B = B_hi + B_lo -> segments (A_lo,B_hi),
// (A_hi,B_lo), (A_hi,B_hi)) and how two contractions that land in the same output are
// fused into one launch.
//
// One CTA per SM, 192 threads: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer
// (one elected lane), warps 2..5 = epilogue (each owns the 32 TMEM lanes it may read).
// Tiles are 128 x BN with BLOCK_K = 64 (one 128-byte swizzle row of bf16), a STAGES-deep
// smem ring fed by TMA, and two TMEM accumulators so the epilogue of tile i overlaps the
// MMAs of tile i+1.
//
// Epilogues (template EPI):
//   EPI_STORE : D (+bias) -> fp32 C, coalesced through a warp-private smem transpose
//   EPI_RED   : C += D with vector atomics (split-K and gradient accumulation)
//   EPI_ESTEP : never writes D.  D is pred = z W3^T; the epilogue adds b3, subtracts the
//               staged feature tile x, and reduces exp(-(x-pred)^2 / 2 sigma^2), (x-pred)^2
//               and pred^2 over the tile's columns into per-(row, n-block) partial sums -
//               the reference's compute_em_probabilities (train_nem.py:247-263) and the
//               row sums its outer loss needs (tools/NEM_utilities.py:98-128).
//   EPI_EBWD  : backward of the above: recomputes the pred tile and emits
//               dL/dpred = delta*(alpha_r*E - beta_r) + kappa_r*pred in bf16 operand planes.
#pragma once
#include "ptx.cuh"

namespace vmm {

constexpr int VMM_MAX_SEG = 8;
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int GEMM_THREADS = 192;
constexpr int GROUP_M = 16;
constexpr int EPI_STAGE_BYTES = 32 * 36 * 4;   // per epilogue warp

enum { EPI_STORE = 0, EPI_RED = 1, EPI_ESTEP = 2, EPI_EBWD = 3 };

struct alignas(64) GemmParams {
  CUtensorMap tma_a[VMM_MAX_SEG];
  CUtensorMap tma_b[VMM_MAX_SEG];
  int kb_end[VMM_MAX_SEG];   // running total of k-blocks at the end of each segment
  int nseg;
  int total_kb;
  int M, N;
  int num_m_blk, num_n_blk;
  int splits;
  int vec_ok;                // C rows are 16-byte aligned -> float4 path
  float* C;
  long long ldc;
  const float* bias;         // [N] or null (EPI_STORE); b3 for ESTEP/EBWD
  // ---- E-step epilogues ----
  const void* x;             // features [B*Mv, N] (float or bf16), row pitch ldx
  long long ldx;
  int km;                    // K_latent * Mv
  int mv;                    // views per object
  float exp_scale;           // -log2(e) / (2 sigma^2)
  float* part_e;             // [M, num_n_blk]  sum_d exp(...)
  float* part_sq;            // [M, num_n_blk]  sum_d (x-pred)^2   (nullable)
  float* part_pp;            // [M, num_n_blk]  sum_d pred^2       (nullable)
  const float* row_alpha;    // EBWD per-row scalars
  const float* row_beta;
  const float* row_kappa;    // nullable
  __nv_bfloat16* out_hi;     // EBWD output planes [M, ldo]
  __nv_bfloat16* out_lo;     // nullable
  long long ldo;
};

template <int BN>
struct GemmCfg {
  static constexpr int STAGES = (BN == 256) ? 4 : 6;
  static constexpr int A_BYTES = BM * BK * 2;
  static constexpr int B_BYTES = BN * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;
  static constexpr int TMEM_COLS = 2 * BN;                 // 512 or 256: power of two
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + 4 * EPI_STAGE_BYTES + BAR_BYTES;
};

struct WorkItem {
  int m_blk, n_blk, kb0, kb1;
};

__device__ __forceinline__ WorkItem decode_work(const GemmParams& p, int work) {
  WorkItem w;
  const int tile = work / p.splits;
  const int split = work - tile * p.splits;
  const int per_group = GROUP_M * p.num_n_blk;
  const int group = tile / per_group;
  const int first_m = group * GROUP_M;
  const int gsize = min(p.num_m_blk - first_m, GROUP_M);
  const int in_group = tile - group * per_group;
  w.m_blk = first_m + in_group % gsize;
  w.n_blk = in_group / gsize;
  w.kb0 = static_cast<int>((static_cast<long long>(p.total_kb) * split) / p.splits);
  w.kb1 = static_cast<int>((static_cast<long long>(p.total_kb) * (split + 1)) / p.splits);
  return w;
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t hi16) { return __uint_as_float(hi16 << 16); }

// Features of one 32x32 chunk -> warp-private smem tile (pitch 36 floats), coalesced.
template <typename XT>
__device__ __forceinline__ void stage_x_chunk(const GemmParams& p, float* st, int row0, int col0, int lane) {
  if constexpr (sizeof(XT) == 4) {
    const float* x = static_cast<const float*>(p.x);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = j * 4 + (lane >> 3), cc = (lane & 7) * 4;
      const int grow = row0 + r, gcol = col0 + cc;
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < p.M && gcol < p.N) {
        const long long xr = static_cast<long long>(grow / p.km) * p.mv + grow % p.mv;
        v = __ldg(reinterpret_cast<const float4*>(x + xr * p.ldx + gcol));
      }
      *reinterpret_cast<float4*>(st + r * 36 + cc) = v;
    }
  } else {
    const __nv_bfloat16* x = static_cast<const __nv_bfloat16*>(p.x);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = j * 8 + (lane >> 2), cc = (lane & 3) * 8;
      const int grow = row0 + r, gcol = col0 + cc;
      uint4 v = make_uint4(0u, 0u, 0u, 0u);
      if (grow < p.M && gcol < p.N) {
        const long long xr = static_cast<long long>(grow / p.km) * p.mv + grow % p.mv;
        v = __ldg(reinterpret_cast<const uint4*>(x + xr * p.ldx + gcol));
      }
      float4 lo4 = make_float4(bf16_bits_to_float(v.x & 0xffffu), bf16_bits_to_float(v.x >> 16),
                               bf16_bits_to_float(v.y & 0xffffu), bf16_bits_to_float(v.y >> 16));
      float4 hi4 = make_float4(bf16_bits_to_float(v.z & 0xffffu), bf16_bits_to_float(v.z >> 16),
                               bf16_bits_to_float(v.w & 0xffffu), bf16_bits_to_float(v.w >> 16));
      *reinterpret_cast<float4*>(st + r * 36 + cc) = lo4;
      *reinterpret_cast<float4*>(st + r * 36 + cc + 4) = hi4;
    }
  }
}

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}

template <int BN, bool A_MN, bool B_MN, int EPI, typename XT, int OPL>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~static_cast<uintptr_t>(1023));
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + 4 * EPI_STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int total_work = p.num_m_blk * p.num_n_blk * p.splits;

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < p.nseg; ++s) {
      tma_prefetch_desc(&p.tma_a[s]);
      tma_prefetch_desc(&p.tma_b[s]);
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 4);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
        const WorkItem w = decode_work(p, work);
        int seg = 0;
        for (int g = w.kb0; g < w.kb1; ++g) {
          while (g >= p.kb_end[seg]) ++seg;
          const int kb = g - (seg ? p.kb_end[seg - 1] : 0);
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if constexpr (!A_MN) {
            tma_load_2d(sa, &p.tma_a[seg], &full_bar[stage], kb * BK, w.m_blk * BM);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c)
              tma_load_2d(sa + c * (64 * BK * 2), &p.tma_a[seg], &full_bar[stage], w.m_blk * BM + c * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            tma_load_2d(sb, &p.tma_b[seg], &full_bar[stage], kb * BK, w.n_blk * BN);
          } else {
#pragma unroll
            for (int c = 0; c < BN / 64; ++c)
              tma_load_2d(sb + c * (64 * BK * 2), &p.tma_b[seg], &full_bar[stage], w.n_blk * BN + c * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16(BM, BN, A_MN ? 1 : 0, B_MN ? 1 : 0);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const WorkItem w = decode_work(p, work);
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
      for (int g = w.kb0; g < w.kb1; ++g) {
        mbar_wait(&full_bar[stage], phase);
        tc_fence_after();
        if (lane == 0) {
          const uint32_t a_base = smem_u32(smem + stage * Cfg::STAGE_BYTES);
          const uint32_t b_base = a_base + Cfg::A_BYTES;
#pragma unroll
          for (int k = 0; k < BK / 16; ++k) {
            // K-major SW128: rows at 128 B pitch, 8-row groups 1024 B apart; +32 B per K=16 step.
            // MN-major SW128: 64-wide MN chunks 8 KB apart (LBO), 8-k groups 1024 B apart (SBO);
            //                 +2048 B per K=16 step.
            const uint64_t adesc = A_MN ? umma_smem_desc_sw128(a_base + k * 2048, 64 * BK * 2, 1024)
                                        : umma_smem_desc_sw128(a_base + k * 32, 16, 1024);
            const uint64_t bdesc = B_MN ? umma_smem_desc_sw128(b_base + k * 2048, 64 * BK * 2, 1024)
                                        : umma_smem_desc_sw128(b_base + k * 32, 16, 1024);
            umma_bf16_ss(tmem_d, adesc, bdesc, idesc, (g > w.kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (g == w.kb1 - 1) umma_commit(&tfull_bar[acc]);
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  } else {
    // ===================== epilogue warps =====================
    const int lane_grp = warp & 3;                                   // TMEM lanes [32*lane_grp, +32)
    float* st = reinterpret_cast<float*>(epi_smem + (warp - 2) * EPI_STAGE_BYTES);
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = blockIdx.x; work < total_work; work += gridDim.x) {
      const WorkItem w = decode_work(p, work);
      const int row0 = w.m_blk * BM + lane_grp * 32;                 // first row of this warp
      const int n0 = w.n_blk * BN;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) + static_cast<uint32_t>(acc * BN);

      float sum_e = 0.f, sum_sq = 0.f, sum_pp = 0.f;
      float alpha = 0.f, beta = 0.f, kappa = 0.f;
      if constexpr (EPI == EPI_EBWD) {
        const int grow = row0 + lane;
        if (grow < p.M) {
          alpha = p.row_alpha[grow];
          beta = p.row_beta[grow];
          if (p.row_kappa) kappa = p.row_kappa[grow];
        }
      }

#pragma unroll 1
      for (int c = 0; c < BN / 32; ++c) {
        const int col0 = n0 + c * 32;
        if (col0 >= p.N) break;                                      // warp-uniform
        uint32_t v[32];
        if constexpr (EPI == EPI_ESTEP || EPI == EPI_EBWD) stage_x_chunk<XT>(p, st, row0, col0, lane);
        tmem_ld_32x32b_x32(taddr0 + static_cast<uint32_t>(c * 32), v);
        tmem_wait_ld();

        if constexpr (EPI == EPI_STORE || EPI == EPI_RED) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            *reinterpret_cast<float4*>(st + lane * 36 + j * 4) =
                make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]), __uint_as_float(v[4 * j + 2]),
                            __uint_as_float(v[4 * j + 3]));
          __syncwarp();
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const int r = j * 4 + (lane >> 3), cc = (lane & 7) * 4;
            const int grow = row0 + r, gcol = col0 + cc;
            if (grow < p.M && gcol < p.N) {
              float4 val = *reinterpret_cast<const float4*>(st + r * 36 + cc);
              float* dst = p.C + static_cast<long long>(grow) * p.ldc + gcol;
              if (p.vec_ok && gcol + 3 < p.N) {
                if constexpr (EPI == EPI_STORE) {
                  if (p.bias) {
                    const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + gcol));
                    val.x += b.x; val.y += b.y; val.z += b.z; val.w += b.w;
                  }
                  *reinterpret_cast<float4*>(dst) = val;
                } else {
                  atomicAdd(reinterpret_cast<float4*>(dst), val);
                }
              } else {
                const float e[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  if (gcol + q < p.N) {
                    if constexpr (EPI == EPI_STORE) dst[q] = e[q] + (p.bias ? __ldg(p.bias + gcol + q) : 0.f);
                    else atomicAdd(dst + q, e[q]);
                  }
              }
            }
          }
          __syncwarp();
        } else {
          // x chunk was staged by this warp; make it visible, then each lane reads its own row.
          __syncwarp();
          float out[32];
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const float4 xv = *reinterpret_cast<const float4*>(st + lane * 36 + j * 4);
            const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
            float4 b4 = make_float4(0.f, 0.f, 0.f, 0.f);
            if (col0 + j * 4 < p.N) b4 = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 4));
            const float bs[4] = {b4.x, b4.y, b4.z, b4.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const bool ok = (col0 + j * 4 + q) < p.N;
              const float pred = __uint_as_float(v[4 * j + q]) + bs[q];
              const float d = xs[q] - pred;
              const float e = exp2f(d * d * p.exp_scale);
              if constexpr (EPI == EPI_ESTEP) {
                if (ok) { sum_e += e; sum_sq += d * d; sum_pp += pred * pred; }
              } else {
                out[4 * j + q] = ok ? (d * (alpha * e - beta) + kappa * pred) : 0.f;
              }
            }
          }
          if constexpr (EPI == EPI_EBWD) {
            __syncwarp();                                            // everyone has consumed its x row
            uint32_t* stw = reinterpret_cast<uint32_t*>(st);
#pragma unroll
            for (int pl = 0; pl < OPL; ++pl) {
              uint32_t w16[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) {
                float a = out[2 * q], b = out[2 * q + 1];
                if (pl == 1) {
                  a -= __bfloat162float(__float2bfloat16_rn(a));
                  b -= __bfloat162float(__float2bfloat16_rn(b));
                }
                w16[q] = pack_bf16x2(a, b);
              }
#pragma unroll
              for (int q = 0; q < 4; ++q)
                *reinterpret_cast<uint4*>(stw + lane * 20 + q * 4) = make_uint4(w16[4 * q], w16[4 * q + 1], w16[4 * q + 2], w16[4 * q + 3]);
              __syncwarp();
              __nv_bfloat16* outp = pl == 0 ? p.out_hi : p.out_lo;
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int r = j * 8 + (lane >> 2), q = lane & 3;
                const int grow = row0 + r, gcol = col0 + q * 8;
                if (grow < p.M && gcol < p.N)
                  *reinterpret_cast<uint4*>(outp + static_cast<long long>(grow) * p.ldo + gcol) =
                      *reinterpret_cast<const uint4*>(stw + r * 20 + q * 4);
              }
              __syncwarp();
            }
          } else {
            __syncwarp();                                            // before the next chunk overwrites st
          }
        }
      }
      if constexpr (EPI == EPI_ESTEP) {
        const int grow = row0 + lane;
        if (grow < p.M) {
          const long long o = static_cast<long long>(grow) * p.num_n_blk + w.n_blk;
          p.part_e[o] = sum_e;
          if (p.part_sq) p.part_sq[o] = sum_sq;
          if (p.part_pp) p.part_pp[o] = sum_pp;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tempty_bar[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace vmm
