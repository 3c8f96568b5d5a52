/* vmm_b200.h - C ABI of libvmm_b200.so, the sm_100a implementation of the VMM / Neural-EM
 * latent-view decomposition hot path of hjjpku/multi_view_sort.
 *
 * The reference has no FFI: its boundary is the Python call surface the trainer uses
 * (SURVEY.md §8b).  Each entry point below names the reference code it replaces
 * (paths relative to the reference root).  Conventions:
 *   - plain C types only: device pointers, sizes, a cudaStream_t passed as void*;
 *   - the CALLER owns every buffer, including workspaces (query the *_bytes functions);
 *   - return 0 on success or a negative VMM_E_* code; never throws, never aborts;
 *     vmm_last_error() returns a per-thread message for the last failure;
 *   - no global mutable state besides a per-device attribute cache: safe to call from
 *     several host threads, one stream per call;
 *   - all work is enqueued on `stream`; nothing synchronises the device.
 *
 * "Operand planes": matrices consumed by the tensor cores are 16-bit.  VMM_FMT_BF16: one plane is plain
 * bf16; with n planes every operand is stored as n successive bf16 residuals and every contraction sums
 * the cross terms A_i * B_j with i + j < n (3 passes for n = 2, 6 for n = 3) with fp32 accumulation in
 * TMEM; n = 3 reproduces the reference's fp32 arithmetic.  VMM_FMT_F16PAIR: an fp16 hi/lo pair gives the
 * same accuracy in 3 passes (see below) and is what the forward pass of the fp32 / bf16 presets uses.
 */
#ifndef VMM_B200_H_
#define VMM_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VMM_ABI_VERSION 3

#if defined(__GNUC__)
#define VMM_API __attribute__((visibility("default")))
#else
#define VMM_API
#endif

#define VMM_E_INVALID (-1)     /* bad argument / unsupported shape */
#define VMM_E_CUDA (-2)        /* a CUDA runtime/driver call failed */
#define VMM_E_UNSUPPORTED (-3) /* valid request the library does not implement */
#define VMM_E_WORKSPACE (-4)   /* caller-provided workspace too small */

#define VMM_MAX_SEGMENTS 12
#define VMM_MAX_PLANES 3

/* bf16 operand planes of one matrix: p[0] = bf16(v), p[1] = bf16(v - p[0]), p[2] = bf16(v - p[0] - p[1]).
 * n = 1 (plain bf16), 2 (~16 mantissa bits, 3 tensor-core passes per contraction) or
 * 3 (~24 bits, 6 passes: fp32-class). */
typedef struct vmm_planes {
  void* p[VMM_MAX_PLANES];
  int n;
  int fmt; /* VMM_FMT_BF16 or VMM_FMT_F16PAIR */
} vmm_planes;

/* VMM_FMT_BF16: p[i] are successive bf16 residuals (above).
 * VMM_FMT_F16PAIR: p[0] = fp16(v), p[1] = fp16((v - p[0]) * 2^11): 22 mantissa bits in two planes.
 * A contraction accumulates the scaled cross terms first and folds them into the main product
 * with tcgen05.mma's scale_input_d = 11 (D = A*B + D * 2^-11): 3 passes for fp32-class operands
 * instead of 6.  Range is fp16's (|v| < 65504, saturating): used for forward operands
 * (features, activations, weights); gradients stay bf16. */
#define VMM_FMT_BF16 0
#define VMM_FMT_F16PAIR 1

/* vmm_gemm_seg.flags */
#define VMM_SEG_A_F16 1       /* A of this segment is fp16 (else bf16) */
#define VMM_SEG_B_F16 2       /* B of this segment is fp16 (else bf16) */
#define VMM_SEG_RESCALE11 4   /* scale what the chain has accumulated so far by 2^-11 before this segment */

VMM_API int vmm_version(void);
VMM_API const char* vmm_last_error(void);
/* Number of CUDA kernels this library has launched in this process (all threads). */
VMM_API long long vmm_launch_count(void);
/* Measurement aid: while enabled, every tensor-core contraction launch is bracketed by CUDA
 * events on its stream; _read() waits for them, returns the summed device time, the executed
 * tensor-core FLOPs (all split passes) and the launch count since the last call, and resets. */
typedef struct vmm_gemm_record {
  int M, N;
  long long k_total; /* sum of the K of all segments (split passes included) */
  int nseg, chain_kb, chains, splits, epilogue, a_mn, b_mn;
  float ms;
} vmm_gemm_record;
VMM_API int vmm_gemm_profile(int enable);
/* Copies up to max_records per-launch records (with device times) without resetting; returns the count. */
VMM_API int vmm_gemm_profile_records(vmm_gemm_record* out, int max_records);
VMM_API int vmm_gemm_profile_read(double* total_ms, double* executed_flops, long long* launches);

/* ---------------------------------------------------------------------------------------
 * Tensor-core contraction (replaces every nn.Linear / GRUCell matmul on the path:
 * train_nem.py:171-185, 226-241; cuBLAS sgemm in the reference).
 *
 *   C[M,N] (+)= sum_s A_s[M,K_s] * B_s[N,K_s]^T  (+ bias[N])
 *
 * a_mn_major == 0: A_s is stored [M][K_s] (K contiguous, row pitch lda elements);
 * a_mn_major == 1: A_s is stored [K_s][M] (M contiguous, row pitch lda) - the layout a
 * weight-gradient contraction sees.  Same for B with N.  All pointers bf16, 16-byte
 * aligned, pitches multiples of 8 elements.  accumulate != 0 adds into C with atomics
 * (required when splits > 1).  splits <= 0 lets the library choose.
 * chain_kblocks > 0 bounds how many 64-wide K blocks (per segment) the tensor core may
 * accumulate in TMEM before the epilogue folds the partial sum into C in round-to-nearest
 * fp32: the MMA datapath truncates at every accumulate step (measured: -4.4e-9 relative per
 * K element), so fp32-class results need short chains (4 = 256 K elements matches cuBLAS
 * sgemm accuracy).  0 = unlimited.
 * ------------------------------------------------------------------------------------- */
typedef struct vmm_gemm_seg {
  const void* a; /* bf16 */
  long long lda;
  const void* b; /* bf16 */
  long long ldb;
  long long k;
  unsigned flags; /* VMM_SEG_* */
} vmm_gemm_seg;

VMM_API int vmm_gemm_bf16(int M, int N, int nseg, const vmm_gemm_seg* segs, int a_mn_major, int b_mn_major,
                  float* C, long long ldc, const float* bias, int accumulate, int splits, int chain_kblocks,
                  void* stream);

/* Same contraction with output row m multiplied by row_scale[m] (float[M], device) before it is stored / accumulated:
 * un-scales an A operand whose rows were scaled by powers of two to fit fp16 (the data-gradient contractions of the
 * "fp16 gradient" mode: dy rows are stored as fp16(dy * s_m), row_scale = 1 / s_m).  A must not be MN-major. */
VMM_API int vmm_gemm_rowscaled(int M, int N, int nseg, const vmm_gemm_seg* segs, int a_mn_major, int b_mn_major,
                               float* C, long long ldc, const float* row_scale, int accumulate, int splits,
                               int chain_kblocks, void* stream);

/* Execution shape of the contraction kernels: 0 = library's choice (CTA pairs - cluster of 2, tcgen05
 * cta_group::2, 256 x 256 tile per pair - whenever M > 128), 1 = always one CTA per tile, 2 = always pairs.
 * Process-wide; for tests and A/B timing. */
VMM_API int vmm_gemm_force_ncta(int ncta);

/* Tail split of the contraction kernels: the output tiles of a partial last wave are split along K so that every CTA
 * shares them (their parts add into C; a store zeroes that part of C first).  mode -1 = only inside the backward entry
 * points (default: forward results stay bit-identical under batch permutation and from run to run), 0 = never,
 * 1 = every launch.  Process-wide; for tests and A/B timing. */
VMM_API int vmm_gemm_tail_split(int mode);

/* Same contract computed by a plain one-thread-per-output CUDA kernel (no tensor cores):
 * the on-device cross-check used by the GPU tests.  Not used by the product path. */
VMM_API int vmm_gemm_check(int M, int N, int nseg, const vmm_gemm_seg* segs, int a_mn_major, int b_mn_major,
                   float* C, long long ldc, const float* bias, int accumulate, void* stream);

/* fp32 -> bf16 operand planes. */
VMM_API int vmm_split_planes(const float* src, long long n, const vmm_planes* out, void* stream);

/* ---------------------------------------------------------------------------------------
 * E-step fused into the dec3 contraction (replaces NEM.dec3 + compute_em_probabilities,
 * train_nem.py:185, 241-263, and the row sums of compute_outer_loss,
 * tools/NEM_utilities.py:98-128).  pred = z W3^T + b3 is never written to memory.
 *
 *   rows = B*K*M (row r = (b*K + k)*M + m), z planes [rows, zk], W3 planes [D, zk].
 *   x: features [B*M, D] (float, or bf16 when x_is_bf16), row pitch ldx.
 *   part_e [rows, nblk]: sum over the n-block's columns of exp(-(x-pred)^2 / (2 sigma^2))
 *   part_sq / part_pp (nullable): same for (x-pred)^2 and pred^2.
 *   nblk = vmm_estep_num_blocks(D).
 * ------------------------------------------------------------------------------------- */
VMM_API int vmm_estep_num_blocks(int D);
VMM_API int vmm_estep_forward(int rows, int D, int zk, const vmm_planes* z, const vmm_planes* w3, const float* b3,
                              const void* x, int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma,
                              float* part_e, float* part_sq, float* part_pp, void* stream);

/* Backward of the fused E-step: recomputes pred tile by tile and writes
 *   dpred[r,d] = delta*(alpha[r]*exp(-delta^2/(2 sigma^2)) - beta[r]) + kappa[r]*pred,  delta = x - pred
 * as bf16 operand planes [rows, D] (dpred->n planes; kappa may be NULL).
 * d_b3 (nullable): [D] += exact fp32 column sums of dpred - the dec3 bias gradient. */
VMM_API int vmm_estep_backward(int rows, int D, int zk, const vmm_planes* z, const vmm_planes* w3, const float* b3,
                               const void* x, int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma,
                               const float* alpha, const float* beta, const float* kappa, const vmm_planes* dpred,
                               float* d_b3, void* stream);

/* The same pair with delta = x - pred kept between them instead of recomputed: _save is vmm_estep_forward
 * that also writes delta [rows, D] (delta_fmt VMM_DELTA_F16: fp16, saturating; VMM_DELTA_F32: float);
 * _dpred is the streaming backward that reads it (one pass: 2-4 B read + 2 B per plane written per
 * element instead of a rows x D x zk contraction).  Same outputs as vmm_estep_backward. */
VMM_API int vmm_estep_forward_save(int rows, int D, int zk, const vmm_planes* z, const vmm_planes* w3, const float* b3,
                                   const void* x, int x_is_bf16, long long ldx, int k_latent, int m_views, float sigma,
                                   float* part_e, float* part_sq, float* part_pp, void* delta, int delta_fmt,
                                   void* stream);
VMM_API int vmm_estep_dpred(int rows, int D, const void* delta, int delta_fmt, const void* x, int x_is_bf16,
                            long long ldx, int k_latent, int m_views, float sigma, const float* alpha,
                            const float* beta, const float* kappa, const vmm_planes* dpred, float* d_b3, void* stream);

/* ---------------------------------------------------------------------------------------
 * The whole EM loop of one batch (replaces the per-batch loop of tools/Trainer.py:160-203:
 * NEM.init_state, T x NEM.forward (train_nem.py:272-294) and compute_outer_loss of the last
 * iteration), and its hand-written backward (replaces autograd's BPTT through the T
 * iterations, Trainer.py:219).
 *
 * Geometry: B objects, K latent views (-cluster_n), M views, D feature dim, H GRU hidden,
 * T iterations (-iter_num).  Row orders are the reference's: h is [B*K, H], gamma is
 * [B, K, M], x is [B, M, D].
 * ------------------------------------------------------------------------------------- */
typedef struct vmm_em_config {
  int B, K, M, D, H, T;
  int planes;          /* operand planes of the forward contractions (1..3; 1..2 with fwd_fp16)   */
  int planes_bwd;      /* bf16 operand planes of the gradient contractions (1..3)               */
  int fwd_fp16;        /* forward operands are VMM_FMT_F16PAIR instead of bf16 residual planes  */
  int x_is_bf16;       /* features are bf16 ("bf16 features" mode) instead of float             */
  int if_gamma_prior;  /* 0,1,2 as in tools/NEM_utilities.py:98-106                             */
  int stop_gamma_grad; /* detach gamma inside the intra loss (NEM_utilities.py:118-123)         */
  int training;        /* 1: keep every iteration's activations for vmm_em_backward             */
  int accum_chain;     /* chain_kblocks of every contraction (see vmm_gemm_bf16); 0 = unlimited */
  float sigma;         /* e_sigma (train_nem.py:150)                                            */
  float ln_eps;        /* nn.LayerNorm eps, 1e-5                                                */
  float dropout_p;     /* train-mode Dropout after enc1/enc2/dec1 (train_nem.py:172,175,181); 0 = eval */
  unsigned long long dropout_seed; /* masks are a pure function of (seed, stage, iteration, element)    */
  /* Per-contraction operand selection of the forward pass, 2 bits each (VMM_FWD_*), contraction i at bits
   * [2i, 2i+1]: 0 x*W1, 1 z*W13 (enc1 through the fold), 2 enc2, 3 GRU input, 4 GRU hidden, 5 dec1, 6 dec2,
   * 7 dec3 + E-step, 8 the pred recompute of the E-step backward (save_delta == 0 only).  0 everywhere =
   * every contraction uses all `planes`. */
  unsigned fwd_modes;
  /* 0: the E-step backward recomputes pred = z W3^T (a full dec3 contraction per iteration);
   * VMM_DELTA_F16 / VMM_DELTA_F32: the forward E-step epilogue also writes delta = x - pred [B*K*M, D] of
   * every iteration into the workspace (training only) and the backward is a streaming pass over it. */
  int save_delta;
  /* 1 (needs fwd_fp16; training): every data-gradient contraction of the BPTT chain (dpred*W3, dG1*W13, dv5*W5, dv4*W4,
   * dgi*Wih, dgh*Whh, dv2*W2) reads a row-scaled fp16 copy of the gradient (exact row maxima, or for dpred the bound
   * 0.607 sigma |alpha| + max|delta| |beta|) and the fp16 hi plane of the weight instead of single bf16 planes: 11
   * mantissa bits instead of 8 where rounding errors accumulate, at the same number of tensor-core passes.  The
   * weight-gradient contractions (reductions over rows) keep bf16 operands. */
  int dgrad_fp16;
  /* 1: decoder warm-up (tools/Trainer.py:184-185, `-if_decoder_warm_up 1`, epochs < 10): EVERY iteration's residual is
   * weighted by `prior` (NEM.init_state's gamma_prior) instead of the previous iteration's responsibilities (and
   * gamma0 is not read).  The responsibilities then feed only the outer loss and the head. */
  int warm_up;
  /* 1: the outer loss of EVERY iteration is computed (the reference evaluates compute_outer_loss T times,
   * Trainer.py:191-195; `-if_total_loss 1` logs their mean, :203): loss_out holds 3*T floats, iteration t at
   * [3t, 3t+2].  Gradients flow from the LAST iteration's loss only, as in the reference (its list of losses is
   * rebuilt with torch.Tensor(...), which detaches it). */
  int loss_all;
} vmm_em_config;

#define VMM_FWD_FULL 0u      /* all planes of both operands                                   */
#define VMM_FWD_SINGLE 1u    /* leading plane of both operands: one tensor-core pass          */
#define VMM_FWD_ACT_PAIR 2u  /* all planes of the activation x leading plane of the weight    */
#define VMM_FWD_W_PAIR 3u    /* leading plane of the activation x all planes of the weight    */
/* A/B bits of fwd_modes (planes_bwd >= 2 only): leading plane only in the weight- / data-gradient contractions */
#define VMM_BWD_WGRAD_SINGLE (1u << 18)
#define VMM_BWD_DGRAD_SINGLE (1u << 19)
#define VMM_DELTA_NONE 0
#define VMM_DELTA_F16 1
#define VMM_DELTA_F32 2

/* fp32 parameters in the reference's own state_dict layout (torch Linear [out, in]; GRU gate
 * order r, z, n) - train_nem.py:171-185.  after_rnn.* is never used by the reference's
 * forward (train_nem.py:179 vs 232-235) and is not part of this struct. */
typedef struct vmm_em_params {
  const float *enc1_w, *enc1_b, *enc1_ln_g, *enc1_ln_b; /* [1024,D] [1024] [1024] [1024]        */
  const float *enc2_w, *enc2_b, *enc2_ln_g, *enc2_ln_b; /* [4096,1024*M] [4096] ...             */
  const float *gru_w_ih, *gru_w_hh, *gru_b_ih, *gru_b_hh; /* [3H,4096] [3H,H] [3H] [3H]         */
  const float *dec1_w, *dec1_b, *dec1_ln_g, *dec1_ln_b; /* [4096,H] [4096] ...                  */
  const float *dec2_w, *dec2_b, *dec2_ln_g, *dec2_ln_b; /* [1024*M,4096] [1024*M] ...           */
  const float *dec3_w, *dec3_b;                         /* [D,1024] [D]                         */
} vmm_em_params;

/* Same fields, non-const: gradients are ACCUMULATED (+=) into these buffers. */
typedef struct vmm_em_grads {
  float *enc1_w, *enc1_b, *enc1_ln_g, *enc1_ln_b;
  float *enc2_w, *enc2_b, *enc2_ln_g, *enc2_ln_b;
  float *gru_w_ih, *gru_w_hh, *gru_b_ih, *gru_b_hh;
  float *dec1_w, *dec1_b, *dec1_ln_g, *dec1_ln_b;
  float *dec2_w, *dec2_b, *dec2_ln_g, *dec2_ln_b;
  float *dec3_w, *dec3_b;
} vmm_em_grads;

/* Buffer sizes (bytes) the caller must provide. */
VMM_API long long vmm_em_weights_bytes(const vmm_em_config* cfg);   /* prepared bf16 weight planes   */
VMM_API long long vmm_em_workspace_bytes(const vmm_em_config* cfg); /* activations kept for backward */
VMM_API long long vmm_em_scratch_bytes(const vmm_em_config* cfg);   /* transient, forward or backward */

/* Once per set of parameter values (per optimizer step): bf16 operand planes of every weight,
 * W13 = enc1.W * dec3.W and w1b3 = enc1.W * dec3.b (the dec3 -> residual -> enc1 fold, see
 * DESIGN.md) into `weights`. */
VMM_API int vmm_em_prepare(const vmm_em_config* cfg, const vmm_em_params* params, void* weights, void* scratch,
                           void* stream);

/* Forward.  x: [B,M,D]; gamma0, prior: [B,K,M] float (NEM.init_state's gamma / gamma_prior).
 * Outputs: h_out [B*K,H], gamma_out [B,K,M], loss_out[3] = (nem_loss, intra, inter) of the
 * last iteration (3*T floats with cfg->loss_all), all float on the device. */
VMM_API int vmm_em_forward(const vmm_em_config* cfg, const vmm_em_params* params, const void* weights, const void* x,
                           const float* gamma0, const float* prior, void* workspace, void* scratch, float* h_out,
                           float* gamma_out, float* loss_out, void* stream);

/* Backward of vmm_em_forward (cfg->training must have been 1).  d_h [B*K,H] and d_gamma
 * [B,K,M] (either may be NULL = zero) are the gradients w.r.t. h_out / gamma_out; d_loss[3]
 * (device) the gradients w.r.t. loss_out.  Parameter gradients are accumulated into `grads`. */
VMM_API int vmm_em_backward(const vmm_em_config* cfg, const vmm_em_params* params, const void* weights, const void* x,
                            const float* gamma0, const float* prior, const void* workspace, void* scratch,
                            const float* d_h, const float* d_gamma, const float* d_loss, const vmm_em_grads* grads,
                            void* stream);

/* Step-level API: ONE EM iteration from an arbitrary state - exactly NEM.forward(x, (h, pred, gamma))
 * of the reference (train_nem.py:272-294), for callers that drive the loop themselves
 * (tools/Trainer.py:180-198 training, :322-336 validation, :433-446 descriptor export).  It materialises pred
 * [B,K,M,D] (the reference returns it).  Workspace sized with T = 1 (and cfg->training as in this call),
 * scratch with vmm_em_step_scratch_bytes.  With cfg->training = 1 (needs save_delta = VMM_DELTA_F32) the call keeps what
 * vmm_em_step_backward needs in workspace and scratch. */
VMM_API long long vmm_em_step_scratch_bytes(const vmm_em_config* cfg);
VMM_API int vmm_em_step_forward(const vmm_em_config* cfg, const vmm_em_params* params, const void* weights,
                                const void* x, const float* h_in, const float* pred_in, const float* gamma_in,
                                void* workspace, void* scratch, float* h_out, float* pred_out, float* gamma_out,
                                void* stream);

/* Backward of ONE vmm_em_step_forward call made with cfg->training = 1 and save_delta = VMM_DELTA_F32 (workspace sized with
 * T = 1, training = 1; it and step_scratch must be kept unchanged between the two calls): what autograd needs to train
 * through the reference's own loop `for j: state = nem(input, state)` followed by compute_outer_loss on (pred, gamma) of
 * every iteration (tools/Trainer.py:180-198, 219).  d_h_out [B*K,H], d_pred_out [B,K,M,D], d_gamma_out [B,K,M] are the
 * upstream gradients (NULL = zero); parameter gradients are accumulated into `grads`; d_h_in / d_pred_in / d_gamma_in
 * receive the gradients w.r.t. the input state (NULL = not wanted; d_gamma_in needs d_pred_in).  `scratch` is transient,
 * sized with vmm_em_scratch_bytes. */
VMM_API int vmm_em_step_backward(const vmm_em_config* cfg, const vmm_em_params* params, const void* weights, const void* x,
                                 const float* h_in, const float* pred_in, const float* gamma_in, const void* workspace,
                                 const void* step_scratch, void* scratch, const float* d_h_out, const float* d_pred_out,
                                 const float* d_gamma_out, const vmm_em_grads* grads, float* d_h_in, float* d_pred_in,
                                 float* d_gamma_in, void* stream);

/* Address of a named intermediate of iteration `iter` inside `workspace` (for stage-level
 * parity tests and debugging); returns NULL for an unknown name.  Names: "xw1", "g1", "a_hi",
 * "g2", "e_hi", "gi", "gh", "h", "h_hi", "g4", "u_hi", "g5", "z_hi", "z_lo", "gamma", "row_sq", "mean2", "rstd2", "mean4",
 * "rstd4" (LayerNorm row statistics of enc2 / dec1), "delta" (saved x - pred, training with save_delta only). */
VMM_API const void* vmm_em_workspace_ptr(const vmm_em_config* cfg, const void* workspace, const char* name, int iter);

/* ---------------------------------------------------------------------------------------
 * Head: latent-view scoring, sort + concat (or max-pool) over K, 3-layer classifier
 * (replaces Multi_View_Net.forward, train_nem.py:328-380, eval-mode semantics) and its
 * backward.  h is the last hidden state [B*K, H]; gamma [B,K,M] is read only for if_sort == 2.
 * ------------------------------------------------------------------------------------- */
typedef struct vmm_head_config {
  int B, K, H, C, W, M; /* objects, latent views, hidden, classes, classifier width (4096), views */
  int planes, planes_bwd, fwd_fp16, accum_chain;
  int if_sort;          /* 0 unsorted concat, 1 sort by score, 2 sort by responsibility spread    */
  int if_pooling;       /* 1: max over the K latent views instead of concat                       */
  float dropout_p;      /* train-mode Dropout after the two hidden ReLUs (VGG classifier); 0 = eval */
  unsigned long long dropout_seed;
} vmm_head_config;

/* att.0.{weight,bias}, net.0.*, net.3.*, net.6.* of the reference's state_dict (train_nem.py:310,326). */
typedef struct vmm_head_params {
  const float *att_w, *att_b; /* [H] [1]                 */
  const float *w0, *b0;       /* [W, K*H or H] [W]       */
  const float *w3, *b3;       /* [W, W] [W]              */
  const float *w6, *b6;       /* [C, W] [C]              */
} vmm_head_params;
typedef struct vmm_head_grads {
  float *att_w, *att_b, *w0, *b0, *w3, *b3, *w6, *b6; /* accumulated (+=) */
} vmm_head_grads;

VMM_API long long vmm_head_weights_bytes(const vmm_head_config* cfg);
VMM_API long long vmm_head_workspace_bytes(const vmm_head_config* cfg);
VMM_API long long vmm_head_scratch_bytes(const vmm_head_config* cfg);
VMM_API int vmm_head_prepare(const vmm_head_config* cfg, const vmm_head_params* params, void* weights, void* stream);
/* logits [B,C] and/or descriptor [B,W] (the output of the 2nd Linear, pre-ReLU: what save_feature=1
 * returns, train_nem.py:357-361); either may be NULL. */
VMM_API int vmm_head_forward(const vmm_head_config* cfg, const vmm_head_params* params, const void* weights,
                             const float* h, const float* gamma, void* workspace, float* logits, float* descriptor,
                             void* stream);
VMM_API int vmm_head_backward(const vmm_head_config* cfg, const vmm_head_params* params, const void* weights,
                              const float* h, const void* workspace, void* scratch, const float* d_logits,
                              const float* d_descriptor, const vmm_head_grads* grads, float* d_h, void* stream);

/* Softmax cross-entropy of the logits (replaces nn.CrossEntropyLoss at tools/Trainer.py:202):
 * loss_sum[0] += scale * sum_b CE_b, d_logits = scale * (softmax - onehot) (nullable),
 * pred[b] = argmax (nullable).  scale = 1/B for the batch mean (1/B_global when sharded). */
VMM_API int vmm_cross_entropy(int B, int C, const float* logits, const long long* labels, float scale, float* loss_sum,
                              float* d_logits, int* pred, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VMM_B200_H_ */
