// Persistent, warp-specialised tcgen05 GEMM for sm_100a:
//
//     D[M,N] = sum_seg  A_seg[M,K_seg] * B_seg[N,K_seg]^T        (bf16 in, fp32 accumulate in TMEM)
//
// The K loop runs over up to VMM_MAX_SEG "segments", each with its own pair of TMA tensor
// maps.  Segments are how the library gets fp32-class accuracy out of bf16 tensor cores
// (error-compensated split: A = A_0 + A_1 (+ A_2) in successive bf16 residuals, likewise B, and
// the segments are the cross terms A_i B_j with i + j below the split order) and how two
// contractions that land in the same output are fused into one launch.
//
// NCTA = 2 (the default for M > 128): CTAs run as pairs (cluster of 2, tcgen05 cta_group::2).  A pair owns a
// 256 x BN tile: each CTA loads its own 128 rows of A and HALF of the B tile, the leader CTA issues M = 256 MMAs
// that read both CTAs' shared memory, and each CTA drains its own 128 accumulator rows from its own TMEM.
// Per-SM L2->smem traffic drops from 48 KB to 32 KB per 64-wide k-block; the single-CTA kernel ran AT the
// chip's L2 throughput cap (measured: 1.13 PFLOP/s x 48 KB per 128x256x64 block = 12.9 TB/s).
//
// One CTA per SM, 320 threads: warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer
// (one elected lane), warps 2..9 = epilogue (two warps per 32-lane TMEM group, each taking half
// of the tile's columns).
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
//               row sums its outer loss needs (tools/NEM_utilities.py:98-128).  In training it can
//               also write delta = x - pred (fp16 or fp32; template OPL = 2 / 3) so that the
//               backward pass streams over delta instead of recomputing the contraction.
//   EPI_EBWD  : backward of the above: recomputes the pred tile and emits
//               dL/dpred = delta*(alpha_r*E - beta_r) + kappa_r*pred in bf16 operand planes.
#pragma once
#include <cuda_fp16.h>

#include "ptx.cuh"

namespace vmm {

constexpr int VMM_MAX_SEG = 12;
constexpr int BM = 128;
constexpr int BK = 64;
constexpr int EPI_WARPS = 8;
constexpr int GEMM_THREADS = 64 + 32 * EPI_WARPS;
#ifndef VMM_GROUP_M
#define VMM_GROUP_M 32   // m-blocks per rasterisation group (measured 8 / 16 / 32: 171.3 / 170.6 / 169.4 ms of contractions per step)
#endif
constexpr int GROUP_M = VMM_GROUP_M;
constexpr int EPI_STAGE_BYTES = 32 * 32 * 4;   // per epilogue warp (XOR-swizzled 32x32 fp32 tile)

enum { EPI_STORE = 0, EPI_RED = 1, EPI_ESTEP = 2, EPI_EBWD = 3 };

struct alignas(64) GemmParams {
  CUtensorMap tma_a[VMM_MAX_SEG];
  CUtensorMap tma_b[VMM_MAX_SEG];
  CUtensorMap tma_out;       // ESTEP with fp32 delta: the delta tensor, written with TMA stores from the staging tiles
  int seg_kb[VMM_MAX_SEG];   // k-blocks (of BK) in each segment
  unsigned seg_flags[VMM_MAX_SEG];   // VMM_SEG_* bits of each segment
  int nseg;
  int chain_kb;              // k-blocks per segment accumulated in TMEM before the epilogue drains them
  int total_chains;          // ceil(max seg_kb / chain_kb)
  int M, N;
  int num_m_blk, num_n_blk;
  int splits;
  // Tail split: the last (tiles mod workers) output tiles - the partial wave that would otherwise leave most CTAs idle
  // for a whole tile time - are split `tail_splits` ways along K (work items >= tail_first), each part covering a range
  // of accumulation chains of `tail_chain_kb` k-blocks and adding into C (EPI_STORE: into a zeroed tile; the part that
  // holds chain 0 adds the bias).  0 / 1 = off.
  int tail_first, tail_splits, tail_chain_kb, tail_total_chains;
  int total_work;            // work items of the launch (tiles * splits, or full tiles + tail parts)
  int vec_ok;                // C rows are 16-byte aligned -> float4 path
  int tma_c;                 // C is written through tma_out (TMA store / reduce-add of the staging tiles)
  float* C;
  long long ldc;
  const float* bias;         // [N] or null (EPI_STORE); b3 for ESTEP/EBWD
  const float* row_scale;    // [M] or null (EPI_STORE / EPI_RED): D row m is multiplied by row_scale[m] (un-scales a
                             // row-scaled fp16 A operand)
  // ---- E-step epilogues ----
  const void* x;             // features [B*Mv, N] (float or bf16), row pitch ldx
  long long ldx;
  int km;                    // K_latent * Mv
  int mv;                    // views per object
  float exp_scale;           // -log2(e) / (2 sigma^2)
  float* part_e;             // [M, 2*num_n_blk]  sum_d exp(...) per 128-column half tile
  float* part_sq;            // [M, 2*num_n_blk]  sum_d (x-pred)^2   (nullable)
  float* part_pp;            // [M, 2*num_n_blk]  sum_d pred^2       (nullable)
  float* part_dmax;          // [M, 2*num_n_blk]  max_d |x-pred|     (nullable; scales the fp16 copy of dL/dpred)
  const float* row_alpha;    // EBWD per-row scalars
  const float* row_beta;
  const float* row_kappa;    // nullable
  __nv_bfloat16* out_p[3];   // EBWD output planes [M, ldo]: successive bf16 residuals (hi, mid, lo)
  long long ldo;
  float* colsum;             // EBWD: [N] += column sums of dL/dpred (dec3 bias gradient), nullable
  void* delta_out;           // ESTEP with OPL 2 (fp16) / 3 (fp32): delta = x - pred, [M, ldo]
};

// Staging tiles per epilogue warp: the E-step epilogue that writes an fp32 delta keeps the feature tile apart from the
// delta tile (the TMA store of chunk c reads its tile while chunk c+1 is staged) and pays with one pipeline stage
// (5 x 32 KB for a pair).  A third tile (two delta tiles used in turn, 4 stages) measured no better: 1.53 vs 1.50 ms
// for the B = 4096 launch - the epilogue's cost is its shared-memory traffic (16 KB per 32x32 chunk against the 32 KB
// per k-block the tensor core reads), not the wait for the TMA engine.  EPI_BUFS = 3 still works (delta_tile_acquire).
#ifndef VMM_ESTEP_BUFS
#define VMM_ESTEP_BUFS 2   // staging tiles per epilogue warp of the fp32-delta E-step (A/B builds: 3 = two delta tiles used in turn)
#endif
__host__ __device__ constexpr int epi_bufs(int epi, int opl) { return (epi == 2 /* EPI_ESTEP */ && opl == 3) ? VMM_ESTEP_BUFS : 1; }

template <int BN, int NCTA = 1, int EPI_BUFS = 1>
struct GemmCfg {
  static constexpr int A_BYTES = BM * BK * 2;              // this CTA's 128 rows of A
  static constexpr int B_ROWS = BN / NCTA;                 // N rows of the B tile held by this CTA
  static constexpr int B_BYTES = B_ROWS * BK * 2;
  static constexpr int STAGE_BYTES = A_BYTES + B_BYTES;    // 48 KB, or 32 KB per CTA of a pair
  static constexpr int EPI_BYTES = EPI_WARPS * EPI_STAGE_BYTES * EPI_BUFS;
  static constexpr int STAGES = (196608 + EPI_WARPS * EPI_STAGE_BYTES - EPI_BYTES) / STAGE_BYTES;   // 4 (3), or 6 (5) for a pair
  static constexpr int TMEM_COLS = 2 * BN;                 // 512 or 256: power of two
  static constexpr int BAR_BYTES = 256;
  static constexpr int SMEM_BYTES = 1024 + STAGES * STAGE_BYTES + EPI_BYTES + BAR_BYTES;
};

// One work item = one output tile x one contiguous range of accumulation chains [c0, c1).
// A chain is what the tensor core accumulates in TMEM before the epilogue drains it: the MMA
// datapath truncates (round-toward-zero) at every accumulate step, so fp32-class modes keep
// chains short and let the epilogue do the long-range accumulation in round-to-nearest fp32
// (first chain stores, later chains add: same CTA, same thread, same address - no atomics).
struct WorkItem {
  int m_blk, n_blk, c0, c1;
  int chain_kb;              // k-blocks per accumulation chain of this item
  int part;                  // 1: one of several K parts of a tile (every chain adds into C)
};

// `ncta` consecutive m-blocks form one scheduling unit (the two CTAs of a pair take one each).
__device__ __forceinline__ WorkItem decode_work(const GemmParams& p, int work, int ncta, int rank) {
  WorkItem w;
  int tile, split, nsplit, total_chains;
  if (p.tail_splits > 1 && work >= p.tail_first) {
    const int t = work - p.tail_first;
    tile = p.tail_first + t / p.tail_splits;      // (splits == 1 whenever a tail split is planned)
    split = t - (t / p.tail_splits) * p.tail_splits;
    nsplit = p.tail_splits;
    total_chains = p.tail_total_chains;
    w.chain_kb = p.tail_chain_kb;
    w.part = 1;
  } else {
    tile = work / p.splits;
    split = work - tile * p.splits;
    nsplit = p.splits;
    total_chains = p.total_chains;
    w.chain_kb = p.chain_kb;
    w.part = p.splits > 1 ? 1 : 0;
  }
  const int units = (p.num_m_blk + ncta - 1) / ncta;
  const int group_m = GROUP_M / ncta;
  const int per_group = group_m * p.num_n_blk;
  const int group = tile / per_group;
  const int first_m = group * group_m;
  const int gsize = min(units - first_m, group_m);
  const int in_group = tile - group * per_group;
  w.m_blk = (first_m + in_group % gsize) * ncta + rank;
  w.n_blk = in_group / gsize;
  w.c0 = static_cast<int>((static_cast<long long>(total_chains) * split) / nsplit);
  w.c1 = static_cast<int>((static_cast<long long>(total_chains) * (split + 1)) / nsplit);
  return w;
}

__device__ __forceinline__ float bf16_bits_to_float(uint32_t hi16) { return __uint_as_float(hi16 << 16); }

__device__ __forceinline__ uint32_t pack_bf16x2(float a, float b) {
  __nv_bfloat162 t = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&t);
}
__device__ __forceinline__ float ex2_approx(float x) {          // one MUFU; flushes results below 2^-126 to zero
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ uint32_t pack_f16x2_sat(float a, float b) {
  __half2 t = __floats2half2_rn(fminf(fmaxf(a, -65504.f), 65504.f), fminf(fmaxf(b, -65504.f), 65504.f));
  return *reinterpret_cast<uint32_t*>(&t);
}

// Warp-private staging tile: 32 rows x 8 float4, XOR-swizzled so that both the row-per-lane
// accesses and the 4-rows-per-instruction coalesced accesses are bank-conflict free.
__device__ __forceinline__ int st_slot(int r, int c4) { return r * 8 + (c4 ^ (r & 7)); }

// Features of one 32x32 chunk, fetched coalesced into registers (8 x 16 bytes per lane for
// fp32, 4 x 16 for bf16).  Row r of the tile reads feature row (r / (K*Mv)) * Mv + r % Mv.
template <typename XT>
struct XRegs {
  uint4 v[sizeof(XT) == 4 ? 8 : 4];
  float4 b[sizeof(XT) == 4 ? 1 : 2];   // b3 of the columns this lane stages (the same columns in every row it fetches)
};

// Feature-row pointers of the rows a lane fetches (constant over the chunks of one tile; the integer
// divisions of the row map are paid once per tile, not once per chunk).
template <typename XT>
struct XRows {
  const XT* p[sizeof(XT) == 4 ? 8 : 4];
};
template <typename XT>
__device__ __forceinline__ void x_rows(const GemmParams& p, XRows<XT>& xr, int row0, int lane) {
  constexpr int NI = sizeof(XT) == 4 ? 8 : 4;
  constexpr int RPI = 32 / NI, LPR = 32 / RPI, EPL = 16 / sizeof(XT);
  const XT* x = static_cast<const XT*>(p.x);
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    const int grow = row0 + j * RPI + lane / LPR;
    xr.p[j] = nullptr;
    if (grow < p.M) xr.p[j] = x + (static_cast<long long>(grow / p.km) * p.mv + grow % p.mv) * p.ldx + (lane % LPR) * EPL;
  }
}
template <typename XT, bool kBias>
__device__ __forceinline__ void x_fetch(const GemmParams& p, XRegs<XT>& xr, const XRows<XT>& rows, int col0, int lane) {
  constexpr int NI = sizeof(XT) == 4 ? 8 : 4;
  constexpr int RPI = 32 / NI, LPR = 32 / RPI, EPL = 16 / sizeof(XT);
  const bool col_ok = col0 + (lane % LPR) * EPL < p.N;
#pragma unroll
  for (int j = 0; j < NI; ++j) {
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (rows.p[j] && col_ok) v = __ldg(reinterpret_cast<const uint4*>(rows.p[j] + col0));
    xr.v[j] = v;
  }
  if constexpr (kBias) {
#pragma unroll
    for (int i = 0; i < EPL / 4; ++i) {
      float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
      if (col_ok) b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + (lane % LPR) * EPL + i * 4));
      xr.b[i] = b;
    }
  }
}

template <typename XT>
__device__ __forceinline__ void x_stage(const XRegs<XT>& xr, float4* st, int lane) {
  if constexpr (sizeof(XT) == 4) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = j * 4 + (lane >> 3), c4 = lane & 7;
      st[st_slot(r, c4)] = make_float4(__uint_as_float(xr.v[j].x), __uint_as_float(xr.v[j].y),
                                       __uint_as_float(xr.v[j].z), __uint_as_float(xr.v[j].w));
    }
  } else {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = j * 8 + (lane >> 2), c4 = (lane & 3) * 2;
      const uint4 v = xr.v[j];
      st[st_slot(r, c4)] = make_float4(bf16_bits_to_float(v.x & 0xffffu), bf16_bits_to_float(v.x >> 16),
                                       bf16_bits_to_float(v.y & 0xffffu), bf16_bits_to_float(v.y >> 16));
      st[st_slot(r, c4 + 1)] = make_float4(bf16_bits_to_float(v.z & 0xffffu), bf16_bits_to_float(v.z >> 16),
                                           bf16_bits_to_float(v.w & 0xffffu), bf16_bits_to_float(v.w >> 16));
    }
  }
}

// The same with b3 folded in: the staged tile holds x - b3, so that the E-step's delta = (x - b3) - acc needs no bias in
// the row-per-lane layout (where every lane would need all 32 bias values of the chunk).
template <typename XT>
__device__ __forceinline__ void x_stage_sub(const XRegs<XT>& xr, float4* st, int lane) {
  if constexpr (sizeof(XT) == 4) {
    const float4 b = xr.b[0];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = j * 4 + (lane >> 3), c4 = lane & 7;
      st[st_slot(r, c4)] = make_float4(__uint_as_float(xr.v[j].x) - b.x, __uint_as_float(xr.v[j].y) - b.y,
                                       __uint_as_float(xr.v[j].z) - b.z, __uint_as_float(xr.v[j].w) - b.w);
    }
  } else {
    const float4 b0 = xr.b[0], b1 = xr.b[1];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int r = j * 8 + (lane >> 2), c4 = (lane & 3) * 2;
      const uint4 v = xr.v[j];
      st[st_slot(r, c4)] = make_float4(bf16_bits_to_float(v.x & 0xffffu) - b0.x, bf16_bits_to_float(v.x >> 16) - b0.y,
                                       bf16_bits_to_float(v.y & 0xffffu) - b0.z, bf16_bits_to_float(v.y >> 16) - b0.w);
      st[st_slot(r, c4 + 1)] = make_float4(bf16_bits_to_float(v.z & 0xffffu) - b1.x, bf16_bits_to_float(v.z >> 16) - b1.y,
                                           bf16_bits_to_float(v.w & 0xffffu) - b1.z, bf16_bits_to_float(v.w >> 16) - b1.w);
    }
  }
}

// Before delta tile `cur` (0 / 1) of this warp is written again: the TMA store that last read it must be done reading.
// The two tiles are used in turn, so that store is normally the one before the most recent (`last` = the tile of the most
// recent store, -1 none); after a chunk clipped away at the N edge it can be the most recent one.
__device__ __forceinline__ void delta_tile_acquire(int cur, int last, int lane) {
  if (last < 0) return;
  if (lane == 0) {
    if (last == cur) tma_store_wait_read();
    else tma_store_wait_read_keep1();
  }
  __syncwarp();
}

// E-step chunk, fast form (all 32 columns inside N; kSq: also the sum of delta^2 the last iteration's loss needs; a launch
// that wants the sum of pred^2 takes the general form below): per element one
// subtraction, two multiplications, one ex2 and one addition (+ one max with kDmax); delta goes straight from the
// registers into the warp's delta tile, which one TMA store writes while the warp moves on.  kSave: 0 none, 2 fp32.
template <int kSave, bool kDmax, bool kSq>
__device__ __forceinline__ void estep_fast_chunk(const GemmParams& p, const uint32_t (&v)[32], const float4* stx, float4* sto,
                                                 int lane, int row0, int col0, int cur_sto, int& last_sto, float& sum_e,
                                                 float& sum_sq, float& dmax) {
  if constexpr (kSave == 2) delta_tile_acquire(cur_sto, last_sto, lane);
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const float4 xv = stx[st_slot(lane, j)];
    const float d0 = xv.x - __uint_as_float(v[4 * j]), d1 = xv.y - __uint_as_float(v[4 * j + 1]);
    const float d2 = xv.z - __uint_as_float(v[4 * j + 2]), d3 = xv.w - __uint_as_float(v[4 * j + 3]);
    const float dd0 = d0 * d0, dd1 = d1 * d1, dd2 = d2 * d2, dd3 = d3 * d3;
    s[0] += ex2_approx(dd0 * p.exp_scale);
    s[1] += ex2_approx(dd1 * p.exp_scale);
    s[2] += ex2_approx(dd2 * p.exp_scale);
    s[3] += ex2_approx(dd3 * p.exp_scale);
    if constexpr (kSq) { q[0] += dd0; q[1] += dd1; q[2] += dd2; q[3] += dd3; }
    if constexpr (kDmax) dmax = fmaxf(fmaxf(dmax, fmaxf(fabsf(d0), fabsf(d1))), fmaxf(fabsf(d2), fabsf(d3)));
    if constexpr (kSave == 2) sto[st_slot(lane, j)] = make_float4(d0, d1, d2, d3);
  }
  sum_e += (s[0] + s[1]) + (s[2] + s[3]);
  if constexpr (kSq) sum_sq += (q[0] + q[1]) + (q[2] + q[3]);
  if constexpr (kSave == 2) {
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      tma_store_2d(&p.tma_out, sto, col0, row0);
      tma_store_commit();
    }
    last_sto = cur_sto;
  } else {
    __syncwarp();                                    // every lane has read its row before the tile is staged again
  }
}

// ---- per-chunk epilogue bodies ---------------------------------------------------------------
// STORE / RED: accumulator chunk -> staging (row per lane) -> coalesced 4-rows-per-instruction.
template <int EPI>
__device__ __forceinline__ void epi_store_chunk(const GemmParams& p, const uint32_t (&v)[32], float4* st, int row0,
                                                int col0, int lane, bool add_to_c, bool first_chain, bool& st_busy,
                                                int groups_per_chain) {
  if (p.tma_c) {
    // TMA path: scale / bias in the row-per-lane layout, stage the 32x32 tile (its XOR swizzle IS the SWIZZLE_128B box
    // layout) and let ONE bulk tensor store - or reduce-add, for accumulation chains, split-K and gradient
    // accumulation - write it: no read-modify-write loads, no per-element guards (the hardware clips the box).
    if (st_busy) {
      // the staging tile is free once the TMA engine has read it; an accumulation chain also needs the operation that
      // last wrote THESE global addresses - this warp's group of one chain earlier, i.e. `groups_per_chain` groups ago
      // (4 chunks per chain, fewer on a tile clipped at the N edge: chunks past N issue nothing) - to have completed,
      // bulk operations being unordered among themselves
      if (lane == 0) {
        tma_store_wait_read();
        if (add_to_c) tma_store_wait_keep(groups_per_chain - 1);
      }
      __syncwarp();
    }
    float rs = 1.f;
    if (p.row_scale && row0 + lane < p.M) rs = __ldg(p.row_scale + row0 + lane);
    const bool add_bias = (EPI == EPI_STORE) && first_chain && p.bias != nullptr;   // once per output element
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      float4 val = make_float4(__uint_as_float(v[4 * j]) * rs, __uint_as_float(v[4 * j + 1]) * rs,
                               __uint_as_float(v[4 * j + 2]) * rs, __uint_as_float(v[4 * j + 3]) * rs);
      if (add_bias && col0 + j * 4 < p.N) {
        const float4 b = __ldg(reinterpret_cast<const float4*>(p.bias + col0 + j * 4));
        val.x += b.x; val.y += b.y; val.z += b.z; val.w += b.w;
      }
      st[st_slot(lane, j)] = val;
    }
    fence_proxy_async();
    __syncwarp();
    if (lane == 0) {
      if (EPI == EPI_RED || add_to_c) tma_reduce_add_2d(&p.tma_out, st, col0, row0);
      else tma_store_2d(&p.tma_out, st, col0, row0);
      tma_store_commit();
    }
    st_busy = true;
    return;
  }
  // For the read-modify-write chains, issue all eight loads of the running sum before anything
  // depends on them (one exposed L2 round trip per chunk instead of eight).
  float4 old[8];
  const bool rmw = (EPI == EPI_STORE) && add_to_c && p.vec_ok;
  if (rmw) {
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int r = j * 4 + (lane >> 3), c4 = lane & 7;
      const int grow = row0 + r, gcol = col0 + c4 * 4;
      old[j] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (grow < p.M && gcol + 3 < p.N)
        old[j] = *reinterpret_cast<const float4*>(p.C + static_cast<long long>(grow) * p.ldc + gcol);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j)
    st[st_slot(lane, j)] = make_float4(__uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                                       __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int r = j * 4 + (lane >> 3), c4 = lane & 7;
    const int grow = row0 + r, gcol = col0 + c4 * 4;
    if (grow < p.M && gcol < p.N) {
      float4 val = st[st_slot(r, c4)];
      if (p.row_scale) {
        const float rs = __ldg(p.row_scale + grow);
        val.x *= rs; val.y *= rs; val.z *= rs; val.w *= rs;
      }
      float* dst = p.C + static_cast<long long>(grow) * p.ldc + gcol;
      if (p.vec_ok && gcol + 3 < p.N) {
        if constexpr (EPI == EPI_STORE) {
          if (add_to_c) {
            val.x += old[j].x; val.y += old[j].y; val.z += old[j].z; val.w += old[j].w;
          } else if (p.bias) {
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
            if constexpr (EPI == EPI_STORE) dst[q] = e[q] + (add_to_c ? dst[q] : (p.bias ? __ldg(p.bias + gcol + q) : 0.f));
            else atomicAdd(dst + q, e[q]);
          }
      }
    }
  }
  __syncwarp();
}

template <int BN, bool A_MN, bool B_MN, int EPI, typename XT, int OPL, int NCTA>
__global__ void __launch_bounds__(GEMM_THREADS, 1) gemm_tcgen05_kernel(const __grid_constant__ GemmParams p) {
  using Cfg = GemmCfg<BN, NCTA, epi_bufs(EPI, OPL)>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr bool kFused = (EPI == EPI_ESTEP || EPI == EPI_EBWD);
  constexpr int kSave = (EPI == EPI_ESTEP) ? OPL - 1 : 0;     // 0: no delta output, 1: fp16, 2: fp32
  extern __shared__ uint8_t smem_raw[];
  // 1024-byte alignment (SWIZZLE_128B atoms) by pointer arithmetic on the __shared__ array - NOT through an integer
  // round trip, which makes the compiler lose the address space and emit generic LD/ST for all staging traffic
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* epi_smem = smem + STAGES * Cfg::STAGE_BYTES;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(epi_smem + Cfg::EPI_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int rank = NCTA == 2 ? static_cast<int>(cluster_ctarank()) : 0;          // 0 = leader of the pair
  const int worker = NCTA == 2 ? static_cast<int>(cluster_id_x()) : static_cast<int>(blockIdx.x);
  const int nworkers = NCTA == 2 ? static_cast<int>(cluster_count_x()) : static_cast<int>(gridDim.x);
  const int total_work = p.total_work;

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
      mbar_init(&tempty_bar[a], EPI_WARPS * NCTA);        // the leader's barrier collects both CTAs' epilogue warps
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    if constexpr (NCTA == 2) {
      tmem_alloc_pair(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, Cfg::TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();            // the peer's barriers are initialised before anyone signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      // pair: completion bytes of BOTH CTAs are counted on the leader's barrier
      const uint32_t full0 = NCTA == 2 ? mapa_u32(smem_u32(&full_bar[0]), 0) : 0u;
      auto load = [&](void* dst, const CUtensorMap* m, int st, int c0, int c1) {
        if constexpr (NCTA == 2) tma_load_2d_pair(dst, m, full0 + 8u * st, c0, c1);
        else tma_load_2d(dst, m, &full_bar[st], c0, c1);
      };
      for (int work = worker; work < total_work; work += nworkers) {
        const WorkItem w = decode_work(p, work, NCTA, rank);
        const int n0 = w.n_blk * BN + rank * Cfg::B_ROWS;       // first N row of this CTA's part of the B tile
        for (int ch = w.c0; ch < w.c1; ++ch)
        for (int seg = 0; seg < p.nseg; ++seg) {
        const int k_lo = ch * w.chain_kb, k_hi = min(k_lo + w.chain_kb, p.seg_kb[seg]);
        for (int kb = k_lo; kb < k_hi; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1u);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[stage], Cfg::STAGE_BYTES * NCTA);
          uint8_t* sa = smem + stage * Cfg::STAGE_BYTES;
          uint8_t* sb = sa + Cfg::A_BYTES;
          if constexpr (!A_MN) {
            load(sa, &p.tma_a[seg], stage, kb * BK, w.m_blk * BM);
          } else {
#pragma unroll
            for (int c = 0; c < BM / 64; ++c) load(sa + c * (64 * BK * 2), &p.tma_a[seg], stage, w.m_blk * BM + c * 64, kb * BK);
          }
          if constexpr (!B_MN) {
            load(sb, &p.tma_b[seg], stage, kb * BK, n0);
          } else {
#pragma unroll
            for (int c = 0; c < Cfg::B_ROWS / 64; ++c) load(sb + c * (64 * BK * 2), &p.tma_b[seg], stage, n0 + c * 64, kb * BK);
          }
          if (++stage == STAGES) { stage = 0; phase ^= 1u; }
        }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (the leader CTA of a pair) =====================
    if (rank == 0) {
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int work = worker; work < total_work; work += nworkers) {
      const WorkItem w = decode_work(p, work, NCTA, rank);
      for (int ch = w.c0; ch < w.c1; ++ch) {
      mbar_wait(&tempty_bar[acc], acc_phase ^ 1u);
      tc_fence_after();
      const uint32_t tmem_d = tmem_base + static_cast<uint32_t>(acc * BN);
      bool first = true;
      for (int seg = 0; seg < p.nseg; ++seg) {
      const int k_lo = ch * w.chain_kb, k_hi = min(k_lo + w.chain_kb, p.seg_kb[seg]);
      const unsigned fl = p.seg_flags[seg];
      const uint32_t idesc = umma_idesc_16b(BM * NCTA, BN, A_MN ? 1 : 0, B_MN ? 1 : 0, fl & 1u, (fl >> 1) & 1u);
      bool rescale = (fl & 4u) != 0;                     // fold the scaled low-order terms accumulated so far
      for (int kb = k_lo; kb < k_hi; ++kb) {
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
            if constexpr (NCTA == 2) {
              if (rescale && !first && k == 0) umma_f16_ss_rescale11_pair(tmem_d, adesc, bdesc, idesc);
              else umma_bf16_ss_pair(tmem_d, adesc, bdesc, idesc, (!first || k > 0) ? 1u : 0u);
            } else {
              if (rescale && !first && k == 0) umma_f16_ss_rescale11(tmem_d, adesc, bdesc, idesc);
              else umma_bf16_ss(tmem_d, adesc, bdesc, idesc, (!first || k > 0) ? 1u : 0u);
            }
          }
          if constexpr (NCTA == 2) umma_commit_pair(&empty_bar[stage]);     // frees the stage in both CTAs
          else umma_commit(&empty_bar[stage]);
        }
        first = false;
        rescale = false;
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1u; }
      }
      }
      if (lane == 0) {                                   // arrives once every MMA of this chain has completed
        if constexpr (NCTA == 2) umma_commit_pair(&tfull_bar[acc]);
        else umma_commit(&tfull_bar[acc]);
      }
      __syncwarp();
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      }
    }
    }
  } else {
    // ===================== epilogue warps =====================
    // Warp w may read TMEM lanes [32*(w%4), +32).  Two warps share each lane group and split
    // the tile's columns in halves; each walks its half in 32-column chunks with the TMEM
    // load of chunk c+1 and the feature fetch of chunk c+1 in flight while chunk c is processed.
    constexpr int NCH = BN / 2 / 32;
    static_assert(NCH == 4, "epi_store_chunk (tma_store_wait_keep) handles up to 4 chunks per accumulation chain and warp");
    const int ew = warp - 2;
    const int lane_grp = warp & 3;
    const int col_half = ew >> 2;
    float4* st = reinterpret_cast<float4*>(epi_smem + ew * EPI_STAGE_BYTES);      // feature tile / the only tile
    // delta tiles (fp32-delta E-step only; with a single tile per warp everything aliases st)
    constexpr int kEpiBufs = Cfg::EPI_BYTES / (EPI_WARPS * EPI_STAGE_BYTES);
    float4* sto0 = reinterpret_cast<float4*>(epi_smem + (kEpiBufs >= 2 ? EPI_WARPS * EPI_STAGE_BYTES : 0) + ew * EPI_STAGE_BYTES);
    float4* sto1 = kEpiBufs >= 3 ? reinterpret_cast<float4*>(epi_smem + 2 * EPI_WARPS * EPI_STAGE_BYTES + ew * EPI_STAGE_BYTES) : sto0;
    int acc = 0;
    uint32_t acc_phase = 0;
    bool st_busy = false;                      // a TMA store of this warp's staging tile may still be reading it
    int last_sto = -1;                         // fp32-delta E-step: delta tile (0 / 1) of the most recent TMA store, -1 none
    const uint32_t tempty0 = NCTA == 2 ? mapa_u32(smem_u32(&tempty_bar[0]), 0) : 0u;   // the leader's barriers
    constexpr bool kXBias = (EPI == EPI_ESTEP) && kSave != 1;    // the fast E-step chunk stages x - b3
    constexpr int XPF = sizeof(XT) == 2 ? 2 : 1;                 // feature chunks in flight ahead of the one being processed
    for (int work = worker; work < total_work; work += nworkers) {
      const WorkItem w = decode_work(p, work, NCTA, rank);
      const int row0 = w.m_blk * BM + lane_grp * 32;                 // first row of this warp
      const int nbase = w.n_blk * BN + col_half * (BN / 2);
      const int groups_per_chain = min(NCH, max(0, (p.N - nbase + 31) / 32));   // chunks of this warp that lie inside N
      for (int ch = w.c0; ch < w.c1; ++ch) {
      const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(lane_grp * 32) << 16) +
                              static_cast<uint32_t>(acc * BN + col_half * (BN / 2));

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
      XRegs<XT> xa, xb;
      XRows<XT> xrows;
      float4 mybias = make_float4(0.f, 0.f, 0.f, 0.f);               // b3 of this warp's 128 columns, 4 per lane
      if constexpr (kFused) {
        x_rows<XT>(p, xrows, row0, lane);
        x_fetch<XT, kXBias>(p, xa, xrows, nbase, lane);              // overlaps the wait for the MMAs
        if constexpr (XPF == 2) x_fetch<XT, kXBias>(p, xb, xrows, nbase + 32, lane);
        if (nbase + lane * 4 < p.N) mybias = __ldg(reinterpret_cast<const float4*>(p.bias + nbase + lane * 4));
      }
      const bool want_sq = kFused && (EPI == EPI_ESTEP) && p.part_sq != nullptr;   // last iteration only
      const bool want_dmax = kFused && (EPI == EPI_ESTEP) && p.part_dmax != nullptr;
      float dmax = 0.f;

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();

      uint32_t va[32], vb[32];
      tmem_ld_32x32b_x32(taddr0, va);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int col0 = nbase + c * 32;
        uint32_t (&v)[32] = (c & 1) ? vb : va;
        uint32_t (&vn)[32] = (c & 1) ? va : vb;
        tmem_wait_ld();
        if (c + 1 < NCH) tmem_ld_32x32b_x32(taddr0 + static_cast<uint32_t>((c + 1) * 32), vn);
        if (col0 < p.N) {                                            // warp-uniform
          if constexpr (!kFused) {
            epi_store_chunk<EPI>(p, v, st, row0, col0, lane, ch > 0 || w.part != 0, ch == 0, st_busy, groups_per_chain);
          } else {
            XRegs<XT>& xc = (XPF == 2 && (c & 1)) ? xb : xa;
            const int csto = kEpiBufs >= 3 ? (c & 1) : 0;          // delta tile of this chunk
            float4* sto = csto ? sto1 : sto0;
            bool fast = false;                                       // warp-uniform
            if constexpr (kXBias) fast = (col0 + 32 <= p.N) && p.part_pp == nullptr;
            if (fast) x_stage_sub<XT>(xc, st, lane);
            else x_stage<XT>(xc, st, lane);
            if (c + XPF < NCH) x_fetch<XT, kXBias>(p, xc, xrows, col0 + 32 * XPF, lane);
            __syncwarp();
            if (fast) {
              if constexpr (kXBias) {
                if (want_sq) {
                  if (want_dmax) estep_fast_chunk<kSave, true, true>(p, v, st, sto, lane, row0, col0, csto, last_sto, sum_e, sum_sq, dmax);
                  else estep_fast_chunk<kSave, false, true>(p, v, st, sto, lane, row0, col0, csto, last_sto, sum_e, sum_sq, dmax);
                } else {
                  if (want_dmax) estep_fast_chunk<kSave, true, false>(p, v, st, sto, lane, row0, col0, csto, last_sto, sum_e, sum_sq, dmax);
                  else estep_fast_chunk<kSave, false, false>(p, v, st, sto, lane, row0, col0, csto, last_sto, sum_e, sum_sq, dmax);
                }
              }
            } else {
            float out[32];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              if (col0 + j * 4 < p.N) {                              // warp-uniform; N is a multiple of 4
                const float4 xv = st[st_slot(lane, j)];
                const float xs[4] = {xv.x, xv.y, xv.z, xv.w};
                const int src = c * 8 + j;                           // lane holding this float4 of the bias
                const float bs[4] = {__shfl_sync(0xffffffffu, mybias.x, src), __shfl_sync(0xffffffffu, mybias.y, src),
                                     __shfl_sync(0xffffffffu, mybias.z, src), __shfl_sync(0xffffffffu, mybias.w, src)};
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                  const float pred = __uint_as_float(v[4 * j + q]) + bs[q];
                  const float d = xs[q] - pred;
                  const float dd = d * d;
                  const float e = ex2_approx(dd * p.exp_scale);
                  if constexpr (EPI == EPI_ESTEP) {
                    sum_e += e;
                    if (want_sq) { sum_sq += dd; sum_pp = fmaf(pred, pred, sum_pp); }
                    if (want_dmax) dmax = fmaxf(dmax, fabsf(d));
                    if constexpr (kSave != 0) out[4 * j + q] = d;
                  } else {
                    out[4 * j + q] = fmaf(d, fmaf(alpha, e, -beta), kappa * pred);
                  }
                }
              } else {
#pragma unroll
                for (int q = 0; q < 4; ++q) out[4 * j + q] = 0.f;
              }
            }
            __syncwarp();                                            // all lanes have consumed their x row
            if constexpr (kSave == 1) {
              // delta as fp16: 64 B per row -> [32 rows][4 x 16 B], swizzled by (row>>1)&3, then coalesced rows
              uint4* stw = reinterpret_cast<uint4*>(st);
              uint32_t w16[16];
#pragma unroll
              for (int q = 0; q < 16; ++q) w16[q] = pack_f16x2_sat(out[2 * q], out[2 * q + 1]);
#pragma unroll
              for (int q = 0; q < 4; ++q)
                stw[lane * 4 + (q ^ ((lane >> 1) & 3))] = make_uint4(w16[4 * q], w16[4 * q + 1], w16[4 * q + 2], w16[4 * q + 3]);
              __syncwarp();
              __half* outp = static_cast<__half*>(p.delta_out);
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const int r = j * 8 + (lane >> 2), q = lane & 3;
                const int grow = row0 + r, gcol = col0 + q * 8;
                if (grow < p.M && gcol < p.N)
                  *reinterpret_cast<uint4*>(outp + static_cast<long long>(grow) * p.ldo + gcol) = stw[r * 4 + (q ^ ((r >> 1) & 3))];
              }
              __syncwarp();
            } else if constexpr (kSave == 2) {
              // fp32 delta: the swizzled staging tile IS a SWIZZLE_128B box - one TMA store writes it (clipped at the
              // tensor's edges) while the warp goes on; the tile is reused only after the TMA engine has read it
              delta_tile_acquire(csto, last_sto, lane);
#pragma unroll
              for (int j = 0; j < 8; ++j)
                sto[st_slot(lane, j)] = make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
              fence_proxy_async();
              __syncwarp();
              if (lane == 0) {
                tma_store_2d(&p.tma_out, sto, col0, row0);
                tma_store_commit();
              }
              last_sto = csto;
            }
            if constexpr (EPI == EPI_EBWD) {
              if (p.colsum) {
                // exact fp32 column sums of this warp's 32x32 block (bias gradient of dec3)
#pragma unroll
                for (int j = 0; j < 8; ++j)
                  st[st_slot(lane, j)] = make_float4(out[4 * j], out[4 * j + 1], out[4 * j + 2], out[4 * j + 3]);
                __syncwarp();
                const float* stf = reinterpret_cast<const float*>(st);
                float cs = 0.f;
#pragma unroll
                for (int r = 0; r < 32; ++r) cs += stf[st_slot(r, lane >> 2) * 4 + (lane & 3)];
                if (col0 + lane < p.N) atomicAdd(p.colsum + col0 + lane, cs);
                __syncwarp();
              }
              // bf16 planes: 64 B per row -> [32 rows][4 x 16 B], swizzled by (row>>1)&3
              uint4* stw = reinterpret_cast<uint4*>(st);
#pragma unroll
              for (int pl = 0; pl < OPL; ++pl) {
                uint32_t w16[16];
#pragma unroll
                for (int q = 0; q < 16; ++q) {
                  w16[q] = pack_bf16x2(out[2 * q], out[2 * q + 1]);
                  if (pl + 1 < OPL) {                                // residual for the next plane
                    out[2 * q] -= bf16_bits_to_float(w16[q] & 0xffffu);
                    out[2 * q + 1] -= bf16_bits_to_float(w16[q] >> 16);
                  }
                }
#pragma unroll
                for (int q = 0; q < 4; ++q)
                  stw[lane * 4 + (q ^ ((lane >> 1) & 3))] = make_uint4(w16[4 * q], w16[4 * q + 1], w16[4 * q + 2], w16[4 * q + 3]);
                __syncwarp();
                __nv_bfloat16* outp = p.out_p[pl];
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                  const int r = j * 8 + (lane >> 2), q = lane & 3;
                  const int grow = row0 + r, gcol = col0 + q * 8;
                  if (grow < p.M && gcol < p.N)
                    *reinterpret_cast<uint4*>(outp + static_cast<long long>(grow) * p.ldo + gcol) = stw[r * 4 + (q ^ ((r >> 1) & 3))];
                }
                __syncwarp();
              }
            }
            }   // slow path
          }
        }
      }
      if constexpr (EPI == EPI_ESTEP) {
        const int grow = row0 + lane;
        if (grow < p.M) {
          const long long o = static_cast<long long>(grow) * (2 * p.num_n_blk) + 2 * w.n_blk + col_half;
          p.part_e[o] = sum_e;
          if (p.part_sq) p.part_sq[o] = sum_sq;
          if (p.part_pp) p.part_pp[o] = sum_pp;
          if (p.part_dmax) p.part_dmax[o] = dmax;
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if constexpr (NCTA == 2) mbar_arrive_cluster(tempty0 + 8u * acc);
        else mbar_arrive(&tempty_bar[acc]);
      }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1u;
      }
    }
    if (kSave == 2 || (!kFused && p.tma_c)) {
      if (lane == 0) tma_store_wait_all();               // every TMA store of this warp has landed before the CTA exits
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if constexpr (NCTA == 2) cluster_sync_all();            // neither CTA frees TMEM / exits while its peer still reads it
  if (warp == 1) {
    tc_fence_after();
    if constexpr (NCTA == 2) tmem_dealloc_pair(tmem_base, Cfg::TMEM_COLS);
    else tmem_dealloc(tmem_base, Cfg::TMEM_COLS);
  }
}

}  // namespace vmm
