// The classification / descriptor head: latent-view scoring, sort + concat (or max-pool) over
// the K latent views, and the 3-layer classifier (reference: Multi_View_Net.forward,
// train_nem.py:328-380, eval-mode semantics), with its hand-written backward; plus the softmax
// cross-entropy the trainer applies to the logits (tools/Trainer.py:202).
#include <cuda_fp16.h>
#include <math.h>
#include <string.h>

#include "internal.h"
#include "stages.cuh"

namespace vmm {
namespace {

struct HDims {
  int B, K, H, C, W, M, planes, pb, chain, f16, chain_b;
  int FW;          // classifier input width: K*H (concat) or H (max-pool)
  int Cp;          // class count padded to the 8-element pitch TMA needs
  long long R;
  explicit HDims(const vmm_head_config& c)
      : B(c.B), K(c.K), H(c.H), C(c.C), W(c.W), M(c.M), planes(c.planes), pb(c.planes_bwd), chain(c.accum_chain), f16(c.fwd_fp16), chain_b(c.planes_bwd >= 2 ? c.accum_chain : 0),
        FW(c.if_pooling ? c.H : c.K * c.H), Cp((c.C + 7) / 8 * 8), R(1ll * c.B * c.K) {}
};

struct HWeights {
  Op w0, w3, w6;
  long long bytes;
  HWeights(const vmm_head_config& c, void* base) {
    HDims d(c);
    Bump b(base);
    w0 = b.op(1ll * d.W * d.FW, d.planes, d.f16, d.pb);
    w3 = b.op(1ll * d.W * d.W, d.planes, d.f16, d.pb);
    w6 = b.op(1ll * d.C * d.W, d.planes, d.f16, d.pb);
    bytes = b.off + 256;
  }
};

struct HWorkspace {
  float* score;    // [B*K] sigmoid scores
  int* order;      // [B*K] latent index placed in slot j (concat) ...
  int* argmax;     // [B*H] winning latent per column (max-pool)
  Op feat;         // [B, FW]
  float* t0;       // [B, W]  Linear0 output (bias added)
  Op t1;           // relu(t0)
  float* t3;       // [B, W]  Linear3 output = the retrieval descriptor
  Op t3r;          // relu(t3)
  long long bytes;
  HWorkspace(const vmm_head_config& c, void* base) {
    HDims d(c);
    Bump b(base);
    score = b.take<float>(d.R);
    order = b.take<int>(d.R);
    argmax = b.take<int>(1ll * d.B * d.H);
    feat = b.op(1ll * d.B * d.FW, d.planes, d.f16, d.pb);
    t0 = b.take<float>(1ll * d.B * d.W);
    t1 = b.op(1ll * d.B * d.W, d.planes, d.f16, d.pb);
    t3 = b.take<float>(1ll * d.B * d.W);
    t3r = b.op(1ll * d.B * d.W, d.planes, d.f16, d.pb);
    bytes = b.off + 256;
  }
};

struct HScratch {
  Planes dl, dt3, dt0;
  float *dt3f, *dt0f, *dfeat, *ds_pre;
  long long bytes;
  HScratch(const vmm_head_config& c, void* base) {
    HDims d(c);
    Bump b(base);
    dl = b.planes(1ll * d.B * d.Cp, d.pb);
    dt3f = b.take<float>(1ll * d.B * d.W);
    dt3 = b.planes(1ll * d.B * d.W, d.pb);
    dt0f = b.take<float>(1ll * d.B * d.W);
    dt0 = b.planes(1ll * d.B * d.W, d.pb);
    dfeat = b.take<float>(1ll * d.B * d.FW);
    ds_pre = b.take<float>(d.R);
    bytes = b.off + 256;
  }
};

int check_head(const vmm_head_config* c) {
  VMM_REQUIRE(c != nullptr, "null head config");
  VMM_REQUIRE(c->B > 0 && c->K > 0 && c->K <= 32 && c->H > 0 && c->C > 0 && c->W > 0, "bad head dimension");
  VMM_REQUIRE(c->H % 8 == 0 && c->W % 8 == 0, "H and W must be multiples of 8");
  VMM_REQUIRE(c->planes >= 1 && c->planes <= 3 && c->planes_bwd >= 1 && c->planes_bwd <= 3 &&
                  (c->fwd_fp16 ? c->planes <= 2 : c->planes_bwd <= c->planes), "bad plane counts");
  VMM_REQUIRE(c->if_sort >= 0 && c->if_sort <= 2, "if_sort must be 0, 1 or 2");
  VMM_REQUIRE(c->if_sort != 2 || c->M > 0, "if_sort == 2 needs M");
  return 0;
}

__device__ __forceinline__ float wsum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ void put_planes(const Planes& P, long long idx, float v) {
  if (P.fmt == VMM_FMT_F16PAIR) {
    const float c = fminf(fmaxf(v, -65504.f), 65504.f);
    const __half h = __float2half_rn(c);
    reinterpret_cast<__half*>(P.p[0])[idx] = h;
    if (P.n > 1) reinterpret_cast<__half*>(P.p[1])[idx] = __float2half_rn((c - __half2float(h)) * 2048.f);
    return;
  }
#pragma unroll
  for (int i = 0; i < 3; ++i)
    if (i < P.n) {
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      P.p[i][idx] = h;
      v -= __bfloat162float(h);
    }
}
__device__ __forceinline__ void put_op(const Op& o, long long idx, float v) {
  put_planes(o.f, idx, v);
  if (o.b.n > 0 && o.b.p[0] != o.f.p[0]) put_planes(o.b, idx, v);
}

// One block (256 threads) per object: scores, ordering, and the classifier input.
__global__ void __launch_bounds__(256) head_score_kernel(int K, int H, int M, int if_sort, int if_pooling,
                                                         const float* __restrict__ h, const float* __restrict__ gamma,
                                                         const float* __restrict__ att_w, const float* __restrict__ att_b,
                                                         float* __restrict__ score, int* __restrict__ order,
                                                         int* __restrict__ argmax, Op feat) {
  __shared__ float s_score[32], s_key[32], red[8];
  __shared__ int s_order[32];
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int k = 0; k < K; ++k) {
    const float* hr = h + (static_cast<long long>(b) * K + k) * H;
    float a = 0.f;
    for (int j = tid; j < H; j += 256) a = fmaf(hr[j], att_w[j], a);
    a = wsum(a);
    __syncthreads();
    if (lane == 0) red[warp] = a;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      const float s = 1.0f / (1.0f + expf(-(t + att_b[0])));
      s_score[k] = s;
      float key = s;
      if (if_sort == 2) {                              // spread of the responsibilities over the views
        const float* g = gamma + (static_cast<long long>(b) * K + k) * M;
        float mean = 0.f;
        for (int m = 0; m < M; ++m) mean += g[m];
        mean /= M;
        float ss = 0.f;
        for (int m = 0; m < M; ++m) ss = fmaf(g[m] - mean, g[m] - mean, ss);
        key = sqrtf(ss);
      }
      s_key[k] = key;
    }
  }
  __syncthreads();
  if (tid == 0) {
    for (int k = 0; k < K; ++k) s_order[k] = k;
    if (if_sort != 0 && !if_pooling) {                 // descending, stable (lower index first on ties)
      for (int i = 1; i < K; ++i) {
        const int v = s_order[i];
        int j = i - 1;
        while (j >= 0 && s_key[s_order[j]] < s_key[v]) { s_order[j + 1] = s_order[j]; --j; }
        s_order[j + 1] = v;
      }
    }
    for (int k = 0; k < K; ++k) {
      score[static_cast<long long>(b) * K + k] = s_score[k];
      order[static_cast<long long>(b) * K + k] = s_order[k];
    }
  }
  __syncthreads();
  if (!if_pooling) {
    for (int i = tid; i < K * H; i += 256) {
      const int slot = i / H, j = i % H, k = s_order[slot];
      put_op(feat, static_cast<long long>(b) * K * H + i, s_score[k] * h[(static_cast<long long>(b) * K + k) * H + j]);
    }
  } else {
    for (int j = tid; j < H; j += 256) {
      float best = -INFINITY;
      int bk = 0;
      for (int k = 0; k < K; ++k) {
        const float v = s_score[k] * h[(static_cast<long long>(b) * K + k) * H + j];
        if (v > best) { best = v; bk = k; }
      }
      argmax[static_cast<long long>(b) * H + j] = bk;
      put_op(feat, static_cast<long long>(b) * H + j, best);
    }
  }
}

// One block per object: d_h and the pre-sigmoid score gradient.
__global__ void __launch_bounds__(256) head_score_backward_kernel(int K, int H, int if_pooling,
                                                                  const float* __restrict__ h,
                                                                  const float* __restrict__ att_w,
                                                                  const float* __restrict__ score,
                                                                  const int* __restrict__ order,
                                                                  const int* __restrict__ argmax,
                                                                  const float* __restrict__ dfeat,
                                                                  float* __restrict__ d_h, float* __restrict__ ds_pre) {
  __shared__ float red[8];
  __shared__ float s_ds;
  const int b = blockIdx.x, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  for (int k = 0; k < K; ++k) {
    int slot = 0;
    if (!if_pooling)
      for (int j = 0; j < K; ++j)
        if (order[static_cast<long long>(b) * K + j] == k) slot = j;
    const float* hr = h + (static_cast<long long>(b) * K + k) * H;
    const float s = score[static_cast<long long>(b) * K + k];
    float a = 0.f;
    for (int j = tid; j < H; j += 256) {
      float g;
      if (!if_pooling) g = dfeat[static_cast<long long>(b) * K * H + static_cast<long long>(slot) * H + j];
      else g = (argmax[static_cast<long long>(b) * H + j] == k) ? dfeat[static_cast<long long>(b) * H + j] : 0.f;
      a = fmaf(g, hr[j], a);
    }
    a = wsum(a);
    __syncthreads();
    if (lane == 0) red[warp] = a;
    __syncthreads();
    if (tid == 0) {
      float t = 0.f;
      for (int w = 0; w < 8; ++w) t += red[w];
      s_ds = t * s * (1.f - s);
      ds_pre[static_cast<long long>(b) * K + k] = s_ds;
    }
    __syncthreads();
    const float dsp = s_ds;
    for (int j = tid; j < H; j += 256) {
      float g;
      if (!if_pooling) g = dfeat[static_cast<long long>(b) * K * H + static_cast<long long>(slot) * H + j];
      else g = (argmax[static_cast<long long>(b) * H + j] == k) ? dfeat[static_cast<long long>(b) * H + j] : 0.f;
      d_h[(static_cast<long long>(b) * K + k) * H + j] = fmaf(s, g, dsp * att_w[j]);
    }
  }
}

__global__ void __launch_bounds__(256) sum_acc_kernel(long long n, const float* __restrict__ src, float* __restrict__ out) {
  __shared__ float red[8];
  float a = 0.f;
  for (long long i = threadIdx.x; i < n; i += 256) a += src[i];
  a = wsum(a);
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = a;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    out[0] += t;
  }
}
__global__ void __launch_bounds__(256) add_kernel(long long n, const float* __restrict__ src, float* __restrict__ dst) {
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) dst[i] += src[i];
}

// Softmax cross-entropy, mean over the batch scaled by `scale` (1/B_global for sharded batches).
// One warp per object.  loss_sum += scale * sum_b CE_b ; d_logits = scale * (softmax - onehot).
__global__ void __launch_bounds__(128) cross_entropy_kernel(int B, int C, const float* __restrict__ logits,
                                                            const long long* __restrict__ labels, float scale,
                                                            float* __restrict__ loss_sum, float* __restrict__ d_logits,
                                                            int* __restrict__ pred) {
  const int b = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (b >= B) return;
  const float* x = logits + static_cast<long long>(b) * C;
  float mx = -INFINITY;
  int arg = 0;
  for (int c = lane; c < C; c += 32)
    if (x[c] > mx) { mx = x[c]; arg = c; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    const float om = __shfl_xor_sync(0xffffffffu, mx, o);
    const int oa = __shfl_xor_sync(0xffffffffu, arg, o);
    if (om > mx || (om == mx && oa < arg)) { mx = om; arg = oa; }
  }
  float se = 0.f;
  for (int c = lane; c < C; c += 32) se += expf(x[c] - mx);
  se = wsum(se);
  const float lse = mx + logf(se);
  const int y = static_cast<int>(labels[b]);
  if (d_logits)
    for (int c = lane; c < C; c += 32)
      d_logits[static_cast<long long>(b) * C + c] = scale * (expf(x[c] - lse) - (c == y ? 1.f : 0.f));
  if (lane == 0) {
    atomicAdd(loss_sum, scale * (lse - x[y]));
    if (pred) pred[b] = arg;
  }
}

int lin_fwd(const Planes& a, const Planes& w, long long rows, int n, long long k, float* y, const float* bias, int chain,
            cudaStream_t st) {
  vmm_gemm_seg s[9];
  const int ns = make_segs(s, a, k, w, k, k);
  return gemm_bf16(static_cast<int>(rows), n, ns, s, false, false, y, n, bias, false, 1, chain, st);
}
int lin_dgrad(const Planes& dy, long long ldy, const Planes& w, long long rows, int n, long long k, float* dx, int chain,
              cudaStream_t st) {
  vmm_gemm_seg s[9];
  const int ns = make_segs(s, dy, ldy, w, k, n);
  return gemm_bf16(static_cast<int>(rows), static_cast<int>(k), ns, s, false, true, dx, k, nullptr, false, 1, chain, st);
}
int lin_wgrad(const Planes& dy, long long ldy, const Planes& x, long long rows, int n, long long k, float* dw, int chain,
              cudaStream_t st) {
  vmm_gemm_seg s[9];
  const int ns = make_segs(s, dy, ldy, x, k, rows);
  return gemm_bf16(n, static_cast<int>(k), ns, s, true, true, dw, k, nullptr, true, 0, chain, st);
}

// d_logits [B, C] fp32 -> operand planes with row pitch Cp (zero padded)
__global__ void __launch_bounds__(256) pad_planes_kernel(int B, int C, int Cp, const float* __restrict__ src, Planes out) {
  const long long n = static_cast<long long>(B) * Cp;
  const long long stride = static_cast<long long>(gridDim.x) * blockDim.x;
  for (long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x; i < n; i += stride) {
    const int b = static_cast<int>(i / Cp), c = static_cast<int>(i % Cp);
    put_planes(out, i, c < C ? src[static_cast<long long>(b) * C + c] : 0.f);
  }
}

unsigned blocks_for(long long n) {
  long long b = (n + 255) / 256;
  const long long cap = static_cast<long long>(device_sm_count()) * 16;
  if (b > cap) b = cap;
  if (b < 1) b = 1;
  return static_cast<unsigned>(b);
}

}  // namespace

int head_prepare(const vmm_head_config* c, const vmm_head_params* p, void* weights, cudaStream_t st) {
  VMM_TRY(check_head(c));
  VMM_REQUIRE(p && weights, "null argument");
  HDims d(*c);
  HWeights w(*c, weights);
  VMM_TRY(split_op(p->w0, 1ll * d.W * d.FW, w.w0, st));
  VMM_TRY(split_op(p->w3, 1ll * d.W * d.W, w.w3, st));
  VMM_TRY(split_op(p->w6, 1ll * d.C * d.W, w.w6, st));
  return 0;
}

int head_forward(const vmm_head_config* c, const vmm_head_params* p, const void* weights, const float* h,
                 const float* gamma, void* workspace, float* logits, float* descriptor, cudaStream_t st) {
  VMM_TRY(check_head(c));
  VMM_REQUIRE(p && weights && h && workspace && (logits || descriptor), "null argument");
  VMM_REQUIRE(c->if_sort != 2 || c->if_pooling || gamma, "if_sort == 2 needs gamma");
  HDims d(*c);
  HWeights w(*c, const_cast<void*>(weights));
  HWorkspace ws(*c, workspace);
  head_score_kernel<<<d.B, 256, 0, st>>>(d.K, d.H, d.M, c->if_sort, c->if_pooling, h, gamma, p->att_w, p->att_b, ws.score,
                                         ws.order, ws.argmax, ws.feat);
  VMM_LAUNCHED();
  VMM_TRY(lin_fwd(ws.feat.f, w.w0.f, d.B, d.W, d.FW, ws.t0, p->b0, d.chain, st));
  VMM_TRY(act_planes(ACT_RELU, d.B, d.W, ws.t0, d.W, ws.t1, c->dropout_p, c->dropout_seed + 1, st));
  VMM_TRY(lin_fwd(ws.t1.f, w.w3.f, d.B, d.W, d.W, ws.t3, p->b3, d.chain, st));
  if (descriptor) VMM_CHECK_CUDA(cudaMemcpyAsync(descriptor, ws.t3, sizeof(float) * d.B * d.W, cudaMemcpyDeviceToDevice, st));
  if (logits) {
    VMM_TRY(act_planes(ACT_RELU, d.B, d.W, ws.t3, d.W, ws.t3r, c->dropout_p, c->dropout_seed + 2, st));
    VMM_TRY(lin_fwd(ws.t3r.f, w.w6.f, d.B, d.C, d.W, logits, p->b6, d.chain, st));
  }
  return 0;
}

int head_backward(const vmm_head_config* c, const vmm_head_params* p, const void* weights, const float* h,
                  const void* workspace, void* scratch, const float* d_logits, const float* d_descriptor,
                  const vmm_head_grads* g, float* d_h, cudaStream_t st) {
  VMM_TRY(check_head(c));
  VMM_REQUIRE(p && weights && h && workspace && scratch && g && d_h && (d_logits || d_descriptor), "null argument");
  HDims d(*c);
  HWeights w(*c, const_cast<void*>(weights));
  HWorkspace ws(*c, const_cast<void*>(workspace));
  HScratch sc(*c, scratch);
  const long long BW = 1ll * d.B * d.W;
  if (d_logits) {
    pad_planes_kernel<<<blocks_for(1ll * d.B * d.Cp), 256, 0, st>>>(d.B, d.C, d.Cp, d_logits, sc.dl);
    VMM_LAUNCHED();
    VMM_TRY(colsum_acc(d.B, d.C, d_logits, d.C, g->b6, st));
    VMM_TRY(lin_wgrad(sc.dl, d.Cp, ws.t3r.b, d.B, d.C, d.W, g->w6, d.chain_b, st));
    VMM_TRY(lin_dgrad(sc.dl, d.Cp, w.w6.b, d.B, d.C, d.W, sc.dt3f, d.chain_b, st));
    VMM_TRY(act_backward(ACT_RELU, BW, ws.t3, sc.dt3f, c->dropout_p, c->dropout_seed + 2, st));
    if (d_descriptor) {
      add_kernel<<<blocks_for(BW), 256, 0, st>>>(BW, d_descriptor, sc.dt3f);
      VMM_LAUNCHED();
    }
  } else {
    VMM_CHECK_CUDA(cudaMemcpyAsync(sc.dt3f, d_descriptor, sizeof(float) * BW, cudaMemcpyDeviceToDevice, st));
  }
  VMM_TRY(split_planes(sc.dt3f, BW, sc.dt3, st));
  VMM_TRY(colsum_acc(d.B, d.W, sc.dt3f, d.W, g->b3, st));
  VMM_TRY(lin_wgrad(sc.dt3, d.W, ws.t1.b, d.B, d.W, d.W, g->w3, d.chain_b, st));
  VMM_TRY(lin_dgrad(sc.dt3, d.W, w.w3.b, d.B, d.W, d.W, sc.dt0f, d.chain_b, st));
  VMM_TRY(act_backward(ACT_RELU, BW, ws.t0, sc.dt0f, c->dropout_p, c->dropout_seed + 1, st));
  VMM_TRY(split_planes(sc.dt0f, BW, sc.dt0, st));
  VMM_TRY(colsum_acc(d.B, d.W, sc.dt0f, d.W, g->b0, st));
  VMM_TRY(lin_wgrad(sc.dt0, d.W, ws.feat.b, d.B, d.W, d.FW, g->w0, d.chain_b, st));
  VMM_TRY(lin_dgrad(sc.dt0, d.W, w.w0.b, d.B, d.W, d.FW, sc.dfeat, d.chain_b, st));
  head_score_backward_kernel<<<d.B, 256, 0, st>>>(d.K, d.H, c->if_pooling, h, p->att_w, ws.score, ws.order, ws.argmax,
                                                  sc.dfeat, d_h, sc.ds_pre);
  VMM_LAUNCHED();
  VMM_TRY(gemv_cols_acc(static_cast<int>(d.R), d.H, h, sc.ds_pre, g->att_w, st));
  sum_acc_kernel<<<1, 256, 0, st>>>(d.R, sc.ds_pre, g->att_b);
  VMM_LAUNCHED();
  return 0;
}

int cross_entropy(int B, int C, const float* logits, const long long* labels, float scale, float* loss_sum,
                  float* d_logits, int* pred, cudaStream_t st) {
  VMM_REQUIRE(B > 0 && C > 0 && logits && labels && loss_sum, "cross_entropy: bad argument");
  cross_entropy_kernel<<<(B + 3) / 4, 128, 0, st>>>(B, C, logits, labels, scale, loss_sum, d_logits, pred);
  VMM_LAUNCHED();
  return 0;
}

}  // namespace vmm

extern "C" {

long long vmm_head_weights_bytes(const vmm_head_config* cfg) {
  if (vmm::check_head(cfg) != 0) return VMM_E_INVALID;
  return vmm::HWeights(*cfg, nullptr).bytes;
}
long long vmm_head_workspace_bytes(const vmm_head_config* cfg) {
  if (vmm::check_head(cfg) != 0) return VMM_E_INVALID;
  return vmm::HWorkspace(*cfg, nullptr).bytes;
}
long long vmm_head_scratch_bytes(const vmm_head_config* cfg) {
  if (vmm::check_head(cfg) != 0) return VMM_E_INVALID;
  return vmm::HScratch(*cfg, nullptr).bytes;
}
int vmm_head_prepare(const vmm_head_config* cfg, const vmm_head_params* params, void* weights, void* stream) {
  return vmm::head_prepare(cfg, params, weights, static_cast<cudaStream_t>(stream));
}
int vmm_head_forward(const vmm_head_config* cfg, const vmm_head_params* params, const void* weights, const float* h,
                     const float* gamma, void* workspace, float* logits, float* descriptor, void* stream) {
  return vmm::head_forward(cfg, params, weights, h, gamma, workspace, logits, descriptor, static_cast<cudaStream_t>(stream));
}
int vmm_head_backward(const vmm_head_config* cfg, const vmm_head_params* params, const void* weights, const float* h,
                      const void* workspace, void* scratch, const float* d_logits, const float* d_descriptor,
                      const vmm_head_grads* grads, float* d_h, void* stream) {
  vmm::BackwardScope backward;
  return vmm::head_backward(cfg, params, weights, h, workspace, scratch, d_logits, d_descriptor, grads, d_h,
                            static_cast<cudaStream_t>(stream));
}
int vmm_cross_entropy(int B, int C, const float* logits, const long long* labels, float scale, float* loss_sum,
                      float* d_logits, int* pred, void* stream) {
  return vmm::cross_entropy(B, C, logits, labels, scale, loss_sum, d_logits, pred, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
