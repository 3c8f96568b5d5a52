// Host side of the tcgen05 GEMM: TMA tensor-map construction, tile/split planning, launch,
// and the extern "C" entry points vmm_gemm_bf16 / vmm_estep_forward / vmm_estep_backward.
#include <stdlib.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <vector>

#include <cuda_fp16.h>

#include "common.h"
#include "gemm.cuh"
#include "internal.h"

namespace vmm {

// ---- error text / device attributes -----------------------------------------------------
static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
int device_sm_count() {
  static int cached[64] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
  if (cached[dev] == 0) {
    int n = 0;
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
    cached[dev] = n;
  }
  return cached[dev];
}

// ---- launch counter and per-launch GEMM timing (bench.py's roofline) ------------------------
static std::atomic<long long> g_launches{0};
void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

// When enabled, every tcgen05 GEMM launch is bracketed by two CUDA events on its own stream;
// vmm_gemm_profile_read() sums the elapsed times.  Events are recorded, never synchronised on,
// inside the timed region.
struct GemmProfile {
  std::mutex mu;
  bool enabled = false;
  std::vector<cudaEvent_t> ev;     // pairs
  std::vector<double> flops;       // executed tensor-core FLOPs of each launch
  std::vector<vmm_gemm_record> recs;
  size_t used = 0;
};
static GemmProfile g_prof;

// ---- TMA descriptors ----------------------------------------------------------------------
typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static PFN_tmapEncodeTiled get_encode_fn() {
  static PFN_tmapEncodeTiled fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<PFN_tmapEncodeTiled>(p);
  });
  return fn;
}

// 2-D bf16 tensor [outer][inner] with `pitch` elements between outer rows; 128-byte swizzle;
// out-of-bounds elements read as zero (this is what pads ragged M, N and K edges).
static int make_tmap_bf16(CUtensorMap* m, const void* ptr, long long inner, long long outer, long long pitch,
                          int box_inner, int box_outer) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return VMM_E_CUDA;
  }
  VMM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0, "operand pointer %p is not 16-byte aligned", ptr);
  VMM_REQUIRE(pitch % 8 == 0, "operand pitch %lld is not a multiple of 8 bf16 elements", pitch);
  VMM_REQUIRE(inner > 0 && outer > 0, "empty operand (%lld x %lld)", inner, outer);
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch) * 2};
  cuuint32_t box[2] = {static_cast<cuuint32_t>(box_inner), static_cast<cuuint32_t>(box_outer)};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (inner %lld outer %lld pitch %lld)", static_cast<int>(r),
              inner, outer, pitch);
    return VMM_E_CUDA;
  }
  return 0;
}

// fp32 tensor [outer][inner] written by the epilogues with TMA stores: 32 x 32 boxes (128-byte rows), 128-byte swizzle
// (the layout of the warps' staging tiles); parts of a box outside the tensor are clipped by the hardware.
static int make_tmap_f32_store(CUtensorMap* m, void* ptr, long long inner, long long outer, long long pitch) {
  PFN_tmapEncodeTiled enc = get_encode_fn();
  if (!enc) {
    set_error("cuTensorMapEncodeTiled not available from the driver");
    return VMM_E_CUDA;
  }
  VMM_REQUIRE((reinterpret_cast<uintptr_t>(ptr) & 15) == 0 && pitch % 4 == 0 && inner > 0 && outer > 0, "bad fp32 store tensor");
  cuuint64_t dims[2] = {static_cast<cuuint64_t>(inner), static_cast<cuuint64_t>(outer)};
  cuuint64_t strides[1] = {static_cast<cuuint64_t>(pitch) * 4};
  cuuint32_t box[2] = {32, 32};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, ptr, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                   CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (fp32 store) failed with CUresult %d", static_cast<int>(r));
    return VMM_E_CUDA;
  }
  return 0;
}

// ---- launch -------------------------------------------------------------------------------
// NCTA = 2: CTA pairs (cluster of 2, tcgen05 cta_group::2), one pair per 256 x 256 output tile.
static std::atomic<int> g_force_ncta{0};   // 0 = choose (pairs whenever M > 128); 1 / 2 = forced (tests, A/B timing)

template <int BN, bool A_MN, bool B_MN, int EPI, typename XT, int OPL, int NCTA>
static int launch_one(const GemmParams& p, cudaStream_t stream) {
  using Cfg = GemmCfg<BN, NCTA, epi_bufs(EPI, OPL)>;
  auto kern = gemm_tcgen05_kernel<BN, A_MN, B_MN, EPI, XT, OPL, NCTA>;
  // the dynamic shared-memory opt-in is a per-device attribute of the function: set it once per (kernel, device)
  static std::atomic<unsigned long long> opted_in{0ull};
  int dev = 0;
  VMM_CHECK_CUDA(cudaGetDevice(&dev));
  const unsigned long long bit = 1ull << (dev & 63);
  if (!(opted_in.load(std::memory_order_acquire) & bit)) {
    const cudaError_t attr_err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::SMEM_BYTES);
    if (attr_err != cudaSuccess) {
      set_error("cudaFuncSetAttribute(smem=%d) failed: %s", Cfg::SMEM_BYTES, cudaGetErrorString(attr_err));
      return VMM_E_CUDA;
    }
    opted_in.fetch_or(bit, std::memory_order_release);
  }
  const int total_work = p.total_work;
  const int workers = device_sm_count() / NCTA;
  const int grid = (total_work < workers ? total_work : workers) * NCTA;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (g_prof.enabled) {
    std::lock_guard<std::mutex> lk(g_prof.mu);
    if (g_prof.used + 2 > g_prof.ev.size()) {
      for (int i = 0; i < 256; ++i) {
        cudaEvent_t e;
        if (cudaEventCreate(&e) != cudaSuccess) break;
        g_prof.ev.push_back(e);
      }
    }
    if (g_prof.used + 2 <= g_prof.ev.size()) {
      e0 = g_prof.ev[g_prof.used];
      e1 = g_prof.ev[g_prof.used + 1];
      g_prof.used += 2;
      double k = 0;
      for (int s = 0; s < p.nseg; ++s) k += static_cast<double>(p.seg_kb[s]) * BK;
      g_prof.flops.push_back(2.0 * p.M * p.N * k);
      vmm_gemm_record r;
      r.M = p.M; r.N = p.N; r.k_total = static_cast<long long>(k); r.nseg = p.nseg; r.chain_kb = p.chain_kb;
      r.chains = p.total_chains; r.splits = p.splits; r.epilogue = EPI; r.a_mn = A_MN; r.b_mn = B_MN; r.ms = 0.f;
      g_prof.recs.push_back(r);
    }
  }
  if (e0) cudaEventRecord(e0, stream);
  if constexpr (NCTA == 2) {
    cudaLaunchConfig_t cfg;
    memset(&cfg, 0, sizeof(cfg));
    cfg.gridDim = dim3(grid);
    cfg.blockDim = dim3(GEMM_THREADS);
    cfg.dynamicSmemBytes = Cfg::SMEM_BYTES;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    VMM_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, p));
  } else {
    kern<<<grid, GEMM_THREADS, Cfg::SMEM_BYTES, stream>>>(p);
  }
  if (e1) cudaEventRecord(e1, stream);
  VMM_LAUNCHED();
  return 0;
}

static int pick_ncta(const GemmParams& p) {
  const int f = g_force_ncta.load(std::memory_order_relaxed);
  if (f == 1 || f == 2) return f;
  return p.num_m_blk >= 2 ? 2 : 1;
}

template <int EPI, typename XT, int OPL>
static int launch_major(const GemmParams& p, bool a_mn, bool b_mn, cudaStream_t s) {
  const bool pair = pick_ncta(p) == 2;
  if constexpr (EPI == EPI_ESTEP || EPI == EPI_EBWD) {
    if (pair) return launch_one<256, false, false, EPI, XT, OPL, 2>(p, s);
    return launch_one<256, false, false, EPI, XT, OPL, 1>(p, s);
  } else {
    if (pair) {
      if (!a_mn && !b_mn) return launch_one<256, false, false, EPI, XT, OPL, 2>(p, s);
      if (!a_mn && b_mn) return launch_one<256, false, true, EPI, XT, OPL, 2>(p, s);
      if (a_mn && !b_mn) return launch_one<256, true, false, EPI, XT, OPL, 2>(p, s);
      return launch_one<256, true, true, EPI, XT, OPL, 2>(p, s);
    }
    if (!a_mn && !b_mn) return launch_one<256, false, false, EPI, XT, OPL, 1>(p, s);
    if (!a_mn && b_mn) return launch_one<256, false, true, EPI, XT, OPL, 1>(p, s);
    if (a_mn && !b_mn) return launch_one<256, true, false, EPI, XT, OPL, 1>(p, s);
    return launch_one<256, true, true, EPI, XT, OPL, 1>(p, s);
  }
}

static int fill_common(GemmParams& p, int M, int N, int nseg, const vmm_gemm_seg* segs, bool a_mn, bool b_mn,
                       int splits_req, bool allow_split, int chain_req) {
  constexpr int BN = 256;
  VMM_REQUIRE(M > 0 && N > 0, "empty problem M=%d N=%d", M, N);
  VMM_REQUIRE(nseg >= 1 && nseg <= VMM_MAX_SEG, "nseg=%d outside [1,%d]", nseg, VMM_MAX_SEG);
  memset(&p, 0, sizeof(p));
  p.M = M;
  p.N = N;
  p.nseg = nseg;
  p.num_m_blk = ceil_div(M, BM);
  p.num_n_blk = ceil_div(N, BN);
  const int ncta = pick_ncta(p);
  int nkb_max = 0;
  for (int s = 0; s < nseg; ++s) {
    const vmm_gemm_seg& g = segs[s];
    VMM_REQUIRE(g.k > 0 && g.a && g.b, "segment %d is empty", s);
    if (!a_mn) VMM_TRY(make_tmap_bf16(&p.tma_a[s], g.a, g.k, M, g.lda, BK, BM));
    else       VMM_TRY(make_tmap_bf16(&p.tma_a[s], g.a, M, g.k, g.lda, 64, BK));
    if (!b_mn) VMM_TRY(make_tmap_bf16(&p.tma_b[s], g.b, g.k, N, g.ldb, BK, BN / ncta));   // a pair's CTAs load half the tile each
    else       VMM_TRY(make_tmap_bf16(&p.tma_b[s], g.b, N, g.k, g.ldb, 64, BK));
    p.seg_kb[s] = ceil_div(g.k, BK);
    p.seg_flags[s] = g.flags;
    if (p.seg_kb[s] > nkb_max) nkb_max = p.seg_kb[s];
  }
  // Accumulation chains and split-K.  chain_req > 0 bounds the k-blocks one TMEM accumulation may
  // span (fp32-class modes); otherwise a chain is as long as the parallel split allows.
  int chain = (chain_req > 0 && chain_req < nkb_max) ? chain_req : nkb_max;
  int want = 1;
  if (allow_split) {
    if (splits_req > 0) {
      want = splits_req;
    } else {
      const int tiles = ceil_div(p.num_m_blk, ncta) * p.num_n_blk;
      const int sms = device_sm_count() / ncta;          // workers: CTAs, or CTA pairs
      if (tiles < sms) want = sms / tiles;               // fill the machine when output tiles are few
      const int max_by_k = nkb_max / 8 > 0 ? nkb_max / 8 : 1;   // keep >= 8 k-blocks per split
      if (want > max_by_k) want = max_by_k;
    }
    if (want > nkb_max) want = nkb_max;
    if (want < 1) want = 1;
    const int by_split = ceil_div(nkb_max, want);
    if (by_split < chain) chain = by_split;
  }
  p.chain_kb = chain;
  p.total_chains = ceil_div(nkb_max, chain);
  p.splits = want < p.total_chains ? want : p.total_chains;
  p.total_work = ceil_div(p.num_m_blk, ncta) * p.num_n_blk * p.splits;
  return 0;
}

// (m unit, n block) of output tile `tile` in the kernel's rasterisation (decode_work): groups of GROUP_M / ncta m units,
// m fastest inside a group.
static void tile_coords(const GemmParams& p, int tile, int ncta, int* m_unit, int* n_blk, int* group_first, int* group_size) {
  const int units = ceil_div(p.num_m_blk, ncta);
  const int group_m = GROUP_M / ncta;
  const int per_group = group_m * p.num_n_blk;
  const int group = tile / per_group;
  const int first_m = group * group_m;
  const int gsize = units - first_m < group_m ? units - first_m : group_m;
  const int in_group = tile - group * per_group;
  *m_unit = first_m + in_group % gsize;
  *n_blk = in_group / gsize;
  *group_first = first_m;
  *group_size = gsize;
}

// Tail split (GemmParams::tail_*): when the tile count leaves a partial last wave that at most half of the workers would
// work on, split those tiles along K so that the whole machine shares them (backward passes only, see BackwardScope:
// measured -1 % of the step; the forward contractions gain nothing - they sit at the power cap - and would lose their
// bit-reproducibility).  EPI_STORE: the parts add into C, so the
// rectangle of C that holds the tail tiles is zeroed first (it may also cover earlier tiles of the same rows, which
// plain-store over the zeros) - returned through `zero_*` for the caller to memset.  Off for launches that already
// split K, for the non-TMA store path, and when fewer than 8 k-blocks per part would be left.
// -1: inside backward passes only (default); 0: never; 1: every launch (tests, A/B timing; VMM_TAIL_SPLIT overrides)
static std::atomic<int> g_tail_mode{getenv("VMM_TAIL_SPLIT") ? (getenv("VMM_TAIL_SPLIT")[0] == '0' ? 0 : 1) : -1};
static thread_local int g_backward_depth = 0;
BackwardScope::BackwardScope() { ++g_backward_depth; }
BackwardScope::~BackwardScope() { --g_backward_depth; }
static void plan_tail_split(GemmParams& p, int nkb_max, bool store, int* zero_row0, int* zero_rows, int* zero_col0) {
  *zero_rows = 0;
  const int mode = g_tail_mode.load(std::memory_order_relaxed);
  if (mode == 0 || (mode < 0 && g_backward_depth == 0) || p.splits != 1 || (store && !p.tma_c)) return;
  const int ncta = pick_ncta(p);
  const int workers = device_sm_count() / ncta;
  const int tiles = ceil_div(p.num_m_blk, ncta) * p.num_n_blk;
  const int tail = tiles % workers;
  if (tiles <= workers || tail == 0 || tail * 2 > workers) return;
  int s = workers / tail;
  if (s > 8) s = 8;
  if (s > nkb_max / 8) s = nkb_max / 8;
  if (s < 2) return;
  const int first = tiles - tail;
  if (store) {
    // all tail tiles must lie in ONE rasterisation group: then they are the tiles (m in group, n >= n0) plus part of
    // column n0, and the rectangle rows(group) x columns [n0 * BN, N) covers them
    int m0, n0, gf0, gs0, m1, n1, gf1, gs1;
    tile_coords(p, first, ncta, &m0, &n0, &gf0, &gs0);
    tile_coords(p, tiles - 1, ncta, &m1, &n1, &gf1, &gs1);
    if (gf0 != gf1) return;
    *zero_row0 = gf0 * ncta * BM;
    const int row_end = (gf0 + gs0) * ncta * BM < p.M ? (gf0 + gs0) * ncta * BM : p.M;
    *zero_rows = row_end - *zero_row0;
    *zero_col0 = n0 * 256;
  }
  p.tail_first = first;
  p.tail_splits = s;
  if (p.total_chains >= s) {
    p.tail_chain_kb = p.chain_kb;
    p.tail_total_chains = p.total_chains;
  } else {
    p.tail_chain_kb = ceil_div(nkb_max, s);
    p.tail_total_chains = ceil_div(nkb_max, p.tail_chain_kb);
    if (p.tail_total_chains < s) p.tail_splits = p.tail_total_chains;
  }
  p.total_work = first + tail * p.tail_splits;
}

int gemm_bf16(int M, int N, int nseg, const vmm_gemm_seg* segs, bool a_mn, bool b_mn, float* C, long long ldc,
              const float* bias, bool accumulate, int splits, int chain_kb, cudaStream_t stream, const float* row_scale) {
  GemmParams p;
  VMM_TRY(fill_common(p, M, N, nseg, segs, a_mn, b_mn, splits, accumulate, chain_kb));
  VMM_REQUIRE(C != nullptr && ldc >= N, "bad C / ldc");
  VMM_REQUIRE(!(accumulate && bias), "bias cannot be combined with accumulate");
  VMM_REQUIRE(!(row_scale && bias), "row_scale cannot be combined with bias");
  p.C = C;
  p.ldc = ldc;
  p.bias = bias;
  p.row_scale = row_scale;
  p.vec_ok = ((reinterpret_cast<uintptr_t>(C) & 15) == 0 && ldc % 4 == 0 &&
              (!bias || (reinterpret_cast<uintptr_t>(bias) & 15) == 0)) ? 1 : 0;
  static const bool tma_epilogue = !(getenv("VMM_TMA_EPILOGUE") && getenv("VMM_TMA_EPILOGUE")[0] == '0');   // A/B switch
  p.tma_c = (tma_epilogue && p.vec_ok && N % 4 == 0) ? 1 : 0;
  if (p.tma_c) VMM_TRY(make_tmap_f32_store(&p.tma_out, C, N, M, ldc));
  int nkb_max = 0;
  for (int s = 0; s < nseg; ++s) nkb_max = p.seg_kb[s] > nkb_max ? p.seg_kb[s] : nkb_max;
  int zr0 = 0, zrows = 0, zc0 = 0;
  plan_tail_split(p, nkb_max, !accumulate, &zr0, &zrows, &zc0);
  if (accumulate) return launch_major<EPI_RED, float, 1>(p, a_mn, b_mn, stream);
  if (zrows > 0)
    VMM_CHECK_CUDA(cudaMemset2DAsync(C + static_cast<long long>(zr0) * ldc + zc0, static_cast<size_t>(ldc) * 4, 0,
                                     static_cast<size_t>(N - zc0) * 4, static_cast<size_t>(zrows), stream));
  return launch_major<EPI_STORE, float, 1>(p, a_mn, b_mn, stream);
}

// Segment list of A (planes) x B (planes), least significant terms first.
//  bf16 x bf16 residual planes : cross terms A_i B_j with i + j < max(na, nb).
//  fp16 pairs (hi, lo' = lo * 2^11): the scaled cross terms (lo'_a hi_b, hi_a lo'_b) are accumulated
//  first, then folded in by rescaling the accumulator by 2^-11 at the first MMA of the main terms.
//  Mixed (bf16 gradient planes x fp16-pair weights/activations): (a_0 lo'_b) scaled, then a_i hi_b.
int make_segs(vmm_gemm_seg* s, const Planes& a, long long lda, const Planes& b, long long ldb, long long k) {
  const bool af = a.fmt == VMM_FMT_F16PAIR, bf = b.fmt == VMM_FMT_F16PAIR;
  const unsigned ff = (af ? VMM_SEG_A_F16 : 0u) | (bf ? VMM_SEG_B_F16 : 0u);
  int n = 0;
  if (!af && !bf) {
    const int order = a.n > b.n ? a.n : b.n;
    for (int sum = order - 1; sum >= 0; --sum)
      for (int i = 0; i < a.n; ++i) {
        const int j = sum - i;
        if (j < 0 || j >= b.n) continue;
        s[n++] = {a.p[i], lda, b.p[j], ldb, k, 0u};
      }
    return n;
  }
  // scaled group
  if (af && a.n > 1) s[n++] = {a.p[1], lda, b.p[0], ldb, k, ff};
  if (bf && b.n > 1) s[n++] = {a.p[0], lda, b.p[1], ldb, k, ff};
  const int scaled = n;
  // main group: residual planes of the bf16 side (if any) against the hi plane of the fp16 side
  const int na = af ? 1 : a.n, nb = bf ? 1 : b.n;
  for (int sum = na + nb - 2; sum >= 0; --sum)
    for (int i = 0; i < na; ++i) {
      const int j = sum - i;
      if (j < 0 || j >= nb) continue;
      s[n] = {a.p[i], lda, b.p[j], ldb, k, ff};
      if (n == scaled && scaled > 0) s[n].flags |= VMM_SEG_RESCALE11;
      ++n;
    }
  return n;
}

static int estep_common(GemmParams& p, int rows, int D, int zk, const Planes& z, const Planes& w3, const float* b3,
                        const void* x, long long ldx, int k_latent, int m_views, float sigma) {
  VMM_REQUIRE(z.n >= 1 && z.n <= 3 && w3.n >= 1 && w3.n <= 3, "planes must be 1, 2 or 3");
  for (int i = 0; i < z.n; ++i) VMM_REQUIRE(z.p[i] != nullptr, "null z plane");
  for (int i = 0; i < w3.n; ++i) VMM_REQUIRE(w3.p[i] != nullptr, "null W3 plane");
  VMM_REQUIRE(b3 && x, "null operand");
  VMM_REQUIRE(D % 8 == 0 && zk % 8 == 0, "D and zk must be multiples of 8");
  VMM_REQUIRE((reinterpret_cast<uintptr_t>(x) & 15) == 0 && (reinterpret_cast<uintptr_t>(b3) & 15) == 0 &&
                  ldx % 8 == 0, "x / b3 must be 16-byte aligned, ldx a multiple of 8");
  VMM_REQUIRE(sigma > 0.f && k_latent > 0 && m_views > 0 && rows % (k_latent * m_views) == 0, "bad E-step geometry");
  vmm_gemm_seg segs[9];
  const int nseg = make_segs(segs, z, zk, w3, zk, zk);
  VMM_TRY(fill_common(p, rows, D, nseg, segs, false, false, 1, false, 0));
  p.bias = b3;
  p.x = x;
  p.ldx = ldx;
  p.km = k_latent * m_views;
  p.mv = m_views;
  p.exp_scale = -1.4426950408889634f / (2.f * sigma * sigma);
  return 0;
}

int estep_forward(int rows, int D, int zk, const Planes& z, const Planes& w3, const float* b3, const void* x,
                  int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma, float* part_e, float* part_sq,
                  float* part_pp, void* delta, int delta_fmt, cudaStream_t stream, float* part_dmax) {
  GemmParams p;
  VMM_TRY(estep_common(p, rows, D, zk, z, w3, b3, x, ldx, k_latent, m_views, sigma));
  VMM_REQUIRE(part_e != nullptr, "part_e is null");
  p.part_e = part_e;
  p.part_sq = part_sq;
  p.part_pp = part_pp;
  p.part_dmax = part_dmax;
  if (!delta) delta_fmt = VMM_DELTA_NONE;
  VMM_REQUIRE(delta_fmt >= VMM_DELTA_NONE && delta_fmt <= VMM_DELTA_F32, "delta_fmt must be 0, 1 or 2");
  VMM_REQUIRE((reinterpret_cast<uintptr_t>(delta) & 15) == 0, "delta must be 16-byte aligned");
  p.delta_out = delta;
  p.ldo = D;
  if (delta && delta_fmt == VMM_DELTA_F32) VMM_TRY(make_tmap_f32_store(&p.tma_out, delta, D, rows, D));
  // the kernel's OPL template slot selects the delta output of the forward epilogue: 1 none, 2 fp16, 3 fp32
  if (x_is_bf16) {
    if (delta_fmt == VMM_DELTA_F16) return launch_major<EPI_ESTEP, __nv_bfloat16, 2>(p, false, false, stream);
    if (delta_fmt == VMM_DELTA_F32) return launch_major<EPI_ESTEP, __nv_bfloat16, 3>(p, false, false, stream);
    return launch_major<EPI_ESTEP, __nv_bfloat16, 1>(p, false, false, stream);
  }
  if (delta_fmt == VMM_DELTA_F16) return launch_major<EPI_ESTEP, float, 2>(p, false, false, stream);
  if (delta_fmt == VMM_DELTA_F32) return launch_major<EPI_ESTEP, float, 3>(p, false, false, stream);
  return launch_major<EPI_ESTEP, float, 1>(p, false, false, stream);
}

int estep_backward(int rows, int D, int zk, const Planes& z, const Planes& w3, const float* b3, const void* x,
                   int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma, const float* alpha,
                   const float* beta, const float* kappa, const Planes& dpred, float* colsum, cudaStream_t stream) {
  GemmParams p;
  VMM_TRY(estep_common(p, rows, D, zk, z, w3, b3, x, ldx, k_latent, m_views, sigma));
  VMM_REQUIRE(alpha && beta && dpred.n >= 1 && dpred.n <= 3, "null E-step backward argument");
  for (int i = 0; i < dpred.n; ++i) {
    VMM_REQUIRE(dpred.p[i] && (reinterpret_cast<uintptr_t>(dpred.p[i]) & 15) == 0, "dpred planes must be 16-byte aligned");
    p.out_p[i] = dpred.p[i];
  }
  p.row_alpha = alpha;
  p.row_beta = beta;
  p.row_kappa = kappa;
  p.ldo = D;
  p.colsum = colsum;
  if (x_is_bf16) {
    if (dpred.n == 1) return launch_major<EPI_EBWD, __nv_bfloat16, 1>(p, false, false, stream);
    if (dpred.n == 2) return launch_major<EPI_EBWD, __nv_bfloat16, 2>(p, false, false, stream);
    return launch_major<EPI_EBWD, __nv_bfloat16, 3>(p, false, false, stream);
  }
  if (dpred.n == 1) return launch_major<EPI_EBWD, float, 1>(p, false, false, stream);
  if (dpred.n == 2) return launch_major<EPI_EBWD, float, 2>(p, false, false, stream);
  return launch_major<EPI_EBWD, float, 3>(p, false, false, stream);
}

// ---- plain CUDA cross-check (tests only) ---------------------------------------------------
__global__ void gemm_check_kernel(int M, int N, int nseg, vmm_gemm_seg s0, vmm_gemm_seg s1, vmm_gemm_seg s2,
                                  vmm_gemm_seg s3, int a_mn, int b_mn, float* C, long long ldc, const float* bias,
                                  int accumulate) {
  const int n = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = blockIdx.y;
  if (n >= N || m >= M) return;
  const vmm_gemm_seg segs[4] = {s0, s1, s2, s3};
  float acc = 0.f;
  for (int s = 0; s < nseg; ++s) {
    const __nv_bfloat16* a = static_cast<const __nv_bfloat16*>(segs[s].a);
    const __nv_bfloat16* b = static_cast<const __nv_bfloat16*>(segs[s].b);
    if (segs[s].flags & VMM_SEG_RESCALE11) acc *= (1.0f / 2048.f);
    for (long long k = 0; k < segs[s].k; ++k) {
      const __nv_bfloat16 ar = a_mn ? a[k * segs[s].lda + m] : a[m * segs[s].lda + k];
      const __nv_bfloat16 br = b_mn ? b[k * segs[s].ldb + n] : b[n * segs[s].ldb + k];
      const float av = (segs[s].flags & VMM_SEG_A_F16) ? __half2float(*reinterpret_cast<const __half*>(&ar)) : __bfloat162float(ar);
      const float bv = (segs[s].flags & VMM_SEG_B_F16) ? __half2float(*reinterpret_cast<const __half*>(&br)) : __bfloat162float(br);
      acc = fmaf(av, bv, acc);
    }
  }
  float* dst = C + static_cast<long long>(m) * ldc + n;
  if (accumulate) *dst += acc;
  else *dst = acc + (bias ? bias[n] : 0.f);
}

__global__ void split_planes_kernel(const float* __restrict__ src, long long n, __nv_bfloat16* __restrict__ p0,
                                    __nv_bfloat16* __restrict__ p1, __nv_bfloat16* __restrict__ p2, int fmt) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    float v = src[i];
    if (fmt == VMM_FMT_F16PAIR) {
      const __half h = __float2half_rn(fminf(fmaxf(v, -65504.f), 65504.f));
      reinterpret_cast<__half*>(p0)[i] = h;
      if (p1) reinterpret_cast<__half*>(p1)[i] =
          __float2half_rn(fminf(fmaxf((v - __half2float(h)) * 2048.f, -65504.f), 65504.f));
      continue;
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(v);
    p0[i] = h;
    if (p1) {
      v -= __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(v);
      p1[i] = m;
      if (p2) p2[i] = __float2bfloat16_rn(v - __bfloat162float(m));
    }
  }
}

int split_planes(const float* src, long long n, const Planes& out, cudaStream_t stream) {
  if (n <= 0) return 0;
  VMM_REQUIRE(src && out.n >= 1 && out.n <= 3 && out.p[0], "split_planes: bad argument");
  VMM_REQUIRE(out.fmt == VMM_FMT_BF16 || (out.fmt == VMM_FMT_F16PAIR && out.n <= 2), "split_planes: bad plane format");
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  split_planes_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(src, n, out.p[0], out.n > 1 ? out.p[1] : nullptr,
                                                                    out.n > 2 ? out.p[2] : nullptr, out.fmt);
  VMM_LAUNCHED();
  return 0;
}

int split_op(const float* src, long long n, const Op& out, cudaStream_t stream) {
  VMM_TRY(split_planes(src, n, out.f, stream));
  if (out.b.n > 0 && out.b.p[0] != out.f.p[0]) VMM_TRY(split_planes(src, n, out.b, stream));
  return 0;
}

__global__ void bf16_to_f16_kernel(const __nv_bfloat16* __restrict__ src, long long n, __half* __restrict__ dst) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride)
    dst[i] = __float2half_rn(fminf(fmaxf(__bfloat162float(src[i]), -65504.f), 65504.f));
}
int bf16_to_f16(const void* src, long long n, void* dst, cudaStream_t stream) {
  if (n <= 0) return 0;
  long long blocks = (n + 255) / 256;
  const long long cap = static_cast<long long>(device_sm_count()) * 16;
  if (blocks > cap) blocks = cap;
  bf16_to_f16_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const __nv_bfloat16*>(src), n,
                                                                   static_cast<__half*>(dst));
  VMM_LAUNCHED();
  return 0;
}

}  // namespace vmm

// ============================== extern "C" ==================================================
extern "C" {

int vmm_version(void) { return VMM_ABI_VERSION; }
long long vmm_launch_count(void) { return vmm::g_launches.load(); }

int vmm_gemm_force_ncta(int ncta) {
  vmm::g_force_ncta.store(ncta == 1 || ncta == 2 ? ncta : 0);
  return 0;
}

int vmm_gemm_tail_split(int mode) {
  vmm::g_tail_mode.store(mode > 0 ? 1 : (mode == 0 ? 0 : -1));
  return 0;
}

int vmm_gemm_profile(int enable) {
  std::lock_guard<std::mutex> lk(vmm::g_prof.mu);
  vmm::g_prof.enabled = enable != 0;
  vmm::g_prof.used = 0;
  vmm::g_prof.flops.clear();
  vmm::g_prof.recs.clear();
  return 0;
}
int vmm_gemm_profile_records(vmm_gemm_record* out, int max_records) {
  std::lock_guard<std::mutex> lk(vmm::g_prof.mu);
  const size_t n = vmm::g_prof.used / 2;
  int w = 0;
  for (size_t i = 0; i < n && w < max_records; ++i) {
    float t = 0.f;
    if (cudaEventSynchronize(vmm::g_prof.ev[2 * i + 1]) != cudaSuccess ||
        cudaEventElapsedTime(&t, vmm::g_prof.ev[2 * i], vmm::g_prof.ev[2 * i + 1]) != cudaSuccess)
      return VMM_E_CUDA;
    out[w] = vmm::g_prof.recs[i];
    out[w].ms = t;
    ++w;
  }
  return w;
}
int vmm_gemm_profile_read(double* total_ms, double* executed_flops, long long* launches) {
  std::lock_guard<std::mutex> lk(vmm::g_prof.mu);
  double ms = 0, fl = 0;
  const size_t n = vmm::g_prof.used / 2;
  for (size_t i = 0; i < n; ++i) {
    float t = 0.f;
    if (cudaEventSynchronize(vmm::g_prof.ev[2 * i + 1]) != cudaSuccess ||
        cudaEventElapsedTime(&t, vmm::g_prof.ev[2 * i], vmm::g_prof.ev[2 * i + 1]) != cudaSuccess) {
      vmm::set_error("vmm_gemm_profile_read: event query failed");
      return VMM_E_CUDA;
    }
    ms += t;
    fl += vmm::g_prof.flops[i];
  }
  if (total_ms) *total_ms = ms;
  if (executed_flops) *executed_flops = fl;
  if (launches) *launches = static_cast<long long>(n);
  vmm::g_prof.used = 0;
  vmm::g_prof.flops.clear();
  vmm::g_prof.recs.clear();
  return 0;
}
const char* vmm_last_error(void) { return vmm::g_err; }

int vmm_gemm_bf16(int M, int N, int nseg, const vmm_gemm_seg* segs, int a_mn_major, int b_mn_major, float* C,
                  long long ldc, const float* bias, int accumulate, int splits, int chain_kblocks, void* stream) {
  return vmm::gemm_bf16(M, N, nseg, segs, a_mn_major != 0, b_mn_major != 0, C, ldc, bias, accumulate != 0, splits,
                        chain_kblocks, static_cast<cudaStream_t>(stream), nullptr);
}

int vmm_gemm_rowscaled(int M, int N, int nseg, const vmm_gemm_seg* segs, int a_mn_major, int b_mn_major, float* C,
                       long long ldc, const float* row_scale, int accumulate, int splits, int chain_kblocks, void* stream) {
  return vmm::gemm_bf16(M, N, nseg, segs, a_mn_major != 0, b_mn_major != 0, C, ldc, nullptr, accumulate != 0, splits,
                        chain_kblocks, static_cast<cudaStream_t>(stream), row_scale);
}

int vmm_gemm_check(int M, int N, int nseg, const vmm_gemm_seg* segs, int a_mn_major, int b_mn_major, float* C,
                   long long ldc, const float* bias, int accumulate, void* stream) {
  if (nseg < 1 || nseg > 4 || !segs || !C) {
    vmm::set_error("vmm_gemm_check: nseg must be in [1,4]");
    return VMM_E_INVALID;
  }
  vmm_gemm_seg s[4] = {segs[0], segs[nseg > 1 ? 1 : 0], segs[nseg > 2 ? 2 : 0], segs[nseg > 3 ? 3 : 0]};
  dim3 grid((N + 127) / 128, M);
  vmm::gemm_check_kernel<<<grid, 128, 0, static_cast<cudaStream_t>(stream)>>>(M, N, nseg, s[0], s[1], s[2], s[3],
                                                                               a_mn_major, b_mn_major, C, ldc, bias,
                                                                               accumulate);
  VMM_LAUNCHED();
  return 0;
}

int vmm_split_planes(const float* src, long long n, const vmm_planes* out, void* stream) {
  return vmm::split_planes(src, n, vmm::planes_from_c(out), static_cast<cudaStream_t>(stream));
}

int vmm_estep_num_blocks(int D) { return 2 * ((D + 255) / 256); }

int vmm_estep_forward(int rows, int D, int zk, const vmm_planes* z, const vmm_planes* w3, const float* b3, const void* x,
                      int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma, float* part_e, float* part_sq,
                      float* part_pp, void* stream) {
  return vmm::estep_forward(rows, D, zk, vmm::planes_from_c(z), vmm::planes_from_c(w3), b3, x, x_is_bf16, ldx, k_latent,
                            m_views, sigma, part_e, part_sq, part_pp, nullptr, VMM_DELTA_NONE,
                            static_cast<cudaStream_t>(stream));
}

int vmm_estep_forward_save(int rows, int D, int zk, const vmm_planes* z, const vmm_planes* w3, const float* b3,
                           const void* x, int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma,
                           float* part_e, float* part_sq, float* part_pp, void* delta, int delta_fmt, void* stream) {
  return vmm::estep_forward(rows, D, zk, vmm::planes_from_c(z), vmm::planes_from_c(w3), b3, x, x_is_bf16, ldx, k_latent,
                            m_views, sigma, part_e, part_sq, part_pp, delta, delta_fmt, static_cast<cudaStream_t>(stream));
}

int vmm_estep_dpred(int rows, int D, const void* delta, int delta_fmt, const void* x, int x_is_bf16, long long ldx,
                    int k_latent, int m_views, float sigma, const float* alpha, const float* beta, const float* kappa,
                    const vmm_planes* dpred, float* d_b3, void* stream) {
  return vmm::estep_dpred(rows, D, delta, delta_fmt, x, x_is_bf16, ldx, k_latent, m_views, sigma, alpha, beta, kappa,
                          vmm::planes_from_c(dpred), d_b3, static_cast<cudaStream_t>(stream));
}

int vmm_estep_backward(int rows, int D, int zk, const vmm_planes* z, const vmm_planes* w3, const float* b3, const void* x,
                       int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma, const float* alpha,
                       const float* beta, const float* kappa, const vmm_planes* dpred, float* d_b3, void* stream) {
  return vmm::estep_backward(rows, D, zk, vmm::planes_from_c(z), vmm::planes_from_c(w3), b3, x, x_is_bf16, ldx, k_latent,
                             m_views, sigma, alpha, beta, kappa, vmm::planes_from_c(dpred), d_b3,
                             static_cast<cudaStream_t>(stream));
}

}  // extern "C"
