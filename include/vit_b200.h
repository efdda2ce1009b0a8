/* vit_b200 -- C ABI of the B200 (sm_100a) kernels behind the ViT encoder step of ViskaWei/VIT.
 *
 * This is the drop-in boundary: plain pointers and sizes, no torch types.  Every entry point
 *   - takes DEVICE pointers that the caller (PyTorch) allocated and owns; the library never
 *     allocates device memory and never synchronises the host,
 *   - takes the CUDA stream explicitly (`stream` is a cudaStream_t passed as void*), so it can be
 *     called from the Python main thread, from the autograd engine's worker thread and under CUDA
 *     graph capture,
 *   - returns 0 on success or a negative VITB200_ERR_* code (no exceptions cross the ABI);
 *     vitb200_strerror() / vitb200_last_cuda_error() give the text.
 *
 * Layouts: row-major contiguous.  Activations are [rows = B*T, H]; weights are in nn.Linear layout
 * [out, in]; q/k/v are addressed in place inside a [B*T, ld] matrix (ld = 3H for the fused QKV
 * projection) with head h at columns [h*d, (h+1)*d).
 *
 * `dtype` selects the activation / GEMM-operand type: VITB200_F32 (reference default precision
 * '32', src/basemodule.py:233) or VITB200_BF16 (Lightning 'bf16-mixed' == torch.autocast(bf16):
 * bf16 operands, fp32 accumulation, fp32 residual stream / LayerNorm / softmax / loss).
 * Buffers typed `void*` hold that type; buffers typed `float*` are always fp32.
 *
 * Dropout masks are never stored.  They are a pure function of (seed, step, site, element index)
 * through Philox4x32-7; `rng` points at two DEVICE uint64 {seed, step} so that a captured CUDA graph
 * sees a new step on every replay.  `site` distinguishes the dropout call sites of one step.
 *
 * Each entry point names the reference code it replaces (paths relative to the reference repo;
 * HF = transformers/models/vit/modeling_vit.py, the third-party file the reference builds on).
 */
#ifndef VIT_B200_H
#define VIT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define VITB200_VERSION 100

#define VITB200_F32 0
#define VITB200_BF16 1

#define VITB200_OK 0
#define VITB200_ERR_CUDA (-1)
#define VITB200_ERR_SHAPE (-2)  /* unsupported shape (e.g. H % 4 != 0, head_dim not in {8,16,32,64,128}) */
#define VITB200_ERR_ARG (-3)    /* null / inconsistent argument */
#define VITB200_ERR_ALIGN (-4)  /* pointer not 16-byte aligned */
#define VITB200_ERR_DEVICE (-5) /* not an sm_100 device */

#define VITB200_ACT_NONE 0
#define VITB200_ACT_GELU 1 /* erf GELU, src/models/builder.py:246 hidden_act='gelu' */

#define VITB200_LOSS_MSE 0 /* nn.MSELoss  (src/models/specvit.py:53; note 'mae' selects this) */
#define VITB200_LOSS_L1 1  /* nn.L1Loss   (src/models/specvit.py:53, loss name contains 'l1') */
#define VITB200_LOSS_CE 2  /* nn.CrossEntropyLoss (src/models/specvit.py:48) */
#define VITB200_LOSS_GIVEN 3 /* backward only: `labels` holds d(objective)/d(logits) [B,C] f32 (caller-side loss) */

/* dropout sites of one step */
#define VITB200_SITE_EMB 0u
#define VITB200_SITE_ATTN(l) (1u + 3u * (uint32_t)(l))
#define VITB200_SITE_PROJ(l) (2u + 3u * (uint32_t)(l))
#define VITB200_SITE_MLP(l) (3u + 3u * (uint32_t)(l))

const char* vitb200_strerror(int rc);
const char* vitb200_last_cuda_error(void);
int vitb200_version(void);
/* Checks that `device` is compute capability 10.x and sets kernel attributes once. */
int vitb200_init(int device);

/* ---- patch embedding ----------------------------------------------------------------------
 * Replaces SlidingWindowTokenizer.forward / Conv1DPatchTokenizer.forward
 * (src/models/tokenization.py:43-50, 66-69) + SpectraEmbeddings.forward (src/models/embedding.py:79-100):
 *   z[b,0,:]   = drop(cls + pos[0])
 *   z[b,1+p,:] = drop(x[b, p*S : p*S+P] . w^T + bias + pos[1+p])       p < n_valid
 *   z[b,1+p,:] = drop(bias + pos[1+p])                                 n_valid <= p < Np (zero-padded windows)
 * x [B,L] f32; w [H,P] (dtype); bias [H], cls [H], pos [Np+1,H] or NULL: f32; z [B,Np+1,H] f32.
 * In BF16 mode x and w are rounded to bf16 and the projection output is rounded to bf16 before the
 * fp32 adds, as autocast does. */
int vitb200_patch_embed_fwd(const float* x, const void* w, const float* bias, const float* cls, const float* pos,
                            float* z, int B, int L, int P, int S, int Np, int n_valid, int H, float p_drop,
                            const uint64_t* rng, uint32_t site, int dtype, void* stream);
/* Backward of the above.  dz [B,Np+1,H] f32 is the gradient w.r.t. z.  Writes (or accumulates when
 * accumulate != 0) dw [H,P], dbias [H], dcls [H], dpos [Np+1,H] (NULL when no learned positions), all f32.
 * ws: vitb200_patch_embed_bwd_ws_bytes() bytes of scratch whose first 4096 bytes were zeroed once. */
size_t vitb200_patch_embed_bwd_ws_bytes(int B, int Np, int P, int H);
int vitb200_patch_embed_bwd(const float* dz, const float* x, float* dw, float* dbias, float* dcls, float* dpos,
                            int B, int L, int P, int S, int Np, int n_valid, int H, float p_drop,
                            const uint64_t* rng, uint32_t site, int accumulate, int dtype, void* ws, void* stream);

/* ---- residual add + LayerNorm ---------------------------------------------------------------
 * Replaces nn.LayerNorm before/after/final (HF:333,340,455; eps = 1e-12, src/models/builder.py:250)
 * fused with the preceding residual add and hidden dropout (HF:267-268,310-312,337):
 *   z_out = z_in + drop(delta)          (delta NULL: z_out is not written, LN reads z_in)
 *   u     = LN(z_out) * gamma + beta ;  mean/rstd saved for backward
 * z_in, z_out [M,H] f32; delta, u [M,H] (dtype); mean, rstd [M] f32.
 * cls_T > 0: LayerNorm only rows r with r % cls_T == 0 (the CLS rows consumed by the head,
 * src/models/specvit.py:78); u/mean/rstd are then compact, indexed r / cls_T. */
int vitb200_add_ln_fwd(const float* z_in, const void* delta, float* z_out, void* u, float* mean, float* rstd,
                       const float* gamma, const float* beta, int M, int H, int cls_T, float eps, float p_drop,
                       const uint64_t* rng, uint32_t site, int dtype, void* stream);
/* Backward: dz = dres + LN'(du) ; ddelta = dropmask * dz (dtype) ; dgamma, dbeta reduced over rows in a
 * fixed order (deterministic).  du (dtype) is compact when cls_T > 0.  dres, ddelta may be NULL.
 * ws: vitb200_add_ln_bwd_ws_bytes() bytes, first 4096 zeroed once. */
size_t vitb200_add_ln_bwd_ws_bytes(int M, int H);
int vitb200_add_ln_bwd(const void* du, const float* z, const float* mean, const float* rstd, const float* gamma,
                       const float* dres, float* dz, void* ddelta, float* dgamma, float* dbeta, int M, int H,
                       int cls_T, float p_drop, const uint64_t* rng, uint32_t site, int accumulate, int dtype,
                       void* ws, void* stream);

/* ---- Linear ---------------------------------------------------------------------------------
 * Replaces nn.Linear forward for query/key/value (HF:228-230, fused as one [3H,H] weight), attention
 * output.dense (HF:266), intermediate.dense + GELU (HF:297-298), output.dense (HF:309), and
 * LinearPreprocessor (src/models/layers.py:62-63):
 *   y = x . w^T + bias ; act == GELU: y holds the pre-activation and y_act = gelu(y)
 * x [M,K], w [N,K], y, y_act [M,N] (dtype); bias [N] f32 or NULL. */
int vitb200_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* y_act, int M, int N, int K,
                       int act, int dtype, void* stream);
/* dx[M,K] = dy[M,N] . w[N,K], optionally multiplied by gelu'(pre_act[M,K]) (the GELU feeding this
 * Linear's input). */
int vitb200_linear_dgrad(const void* dy, const void* w, const void* pre_act, void* dx, int M, int N, int K,
                         int dtype, void* stream);
/* dw[N,K] (+)= dy^T . x ; dbias[N] (+)= column sums of dy.  Split over rows with a fixed-order second
 * stage (deterministic).  ws: vitb200_linear_wgrad_ws_bytes() bytes, first 4096 zeroed once. */
size_t vitb200_linear_wgrad_ws_bytes(int M, int N, int K);
int vitb200_linear_wgrad(const void* dy, const void* x, float* dw, float* dbias, int M, int N, int K,
                         int accumulate, int dtype, void* ws, void* stream);

/* The bf16 Linear entry points above run on the tcgen05 tensor cores (TMA-staged tiles, TMEM accumulator,
 * fused epilogue; csrc/gemm_tc.cu) whenever N % 8 == 0, K % 8 == 0 and the pointers are 16-byte aligned; fp32 mode
 * and odd shapes use the SIMT kernel.  The tensor-core kernels are also exported directly (bf16 only), and the
 * automatic choice can be overridden for A/B tests: mode 0 = automatic, 1 = SIMT only; returns the old mode. */
int vitb200_tc_supported(int M, int N, int K);
int vitb200_tc_linear_fwd(const void* x, const void* w, const float* bias, void* y, void* y_act, int M, int N, int K,
                          int act, void* stream);
int vitb200_tc_linear_dgrad(const void* dy, const void* w, const void* pre_act, void* dx, int M, int N, int K,
                            void* stream);
size_t vitb200_tc_linear_wgrad_ws_bytes(int M, int N, int K);
int vitb200_tc_linear_wgrad(const void* dy, const void* x, float* dw, float* dbias, int M, int N, int K,
                            int accumulate, void* ws, void* stream);
int vitb200_set_gemm_mode(int mode);

/* ---- fused row-chain kernels (bf16, hidden_size 32 or 64) ------------------------------------------------------
 * At the configured shape (H=32) a step is launch/latency bound, so every row-wise op between two attention
 * calls runs in ONE kernel: a CTA owns 128 token rows, all weights of the layer are TMA-staged in shared memory,
 * the four GEMMs run on tcgen05 with the accumulator in TMEM, and each epilogue (bias, dropout, residual,
 * LayerNorm, GELU) writes the next GEMM's A operand straight back into (swizzled) shared memory:
 *   embed : x -> patch GEMM -> +bias,+pos,CLS,dropout -> z0 -> LN1 -> QKV GEMM                 (embedding.py:79-100, HF:333)
 *   layer : ctx -> out-proj -> dropout,+res -> LN2 -> MLP-up,GELU -> MLP-down -> dropout,+res
 *           -> (LN1 + QKV of the next layer | final LN of the CLS rows)                         (HF:266-346,455)
 * Same math, same dropout masks, same saved-for-backward tensors as the unfused entry points above. */
typedef struct {
  int B, L, P, S, Np, n_valid, H;   /* spectrum length L, patch P, stride S, Np patches, H hidden; T = Np+1 */
  float eps, p_drop;
  const uint64_t* rng;
  const float* x;                   /* [B, L] f32 */
  const void* w_p;                  /* [H, P] bf16 */
  const float *b_p, *cls, *pos;     /* [H], [H], [T,H] or NULL */
  const float *ln_g, *ln_b;         /* LN1 of layer 0 */
  const void* w_qkv;                /* [3H, H] bf16 */
  const float* b_qkv;               /* [3H] */
  float* z0;                        /* [B*T, H] f32 */
  void* u;                          /* [B*T, H] bf16 */
  float *mean, *rstd;               /* [B*T] */
  void* qkv;                        /* [B*T, 3H] bf16 */
} vitb200_embed_fwd_args;

typedef struct {
  int B, T, H, last;                /* last != 0: final LayerNorm of the CLS rows instead of LN1 + QKV */
  float eps, p_drop;
  const uint64_t* rng;
  uint32_t site_proj, site_mlp;
  const void* ctx;                  /* [B*T, H] bf16 attention output */
  const float* z_in;                /* [B*T, H] f32 residual stream entering the layer */
  const void *w_o, *w_1, *w_2, *w_qkv;          /* bf16 [H,H], [4H,H], [H,4H], [3H,H] (next layer; unused if last) */
  const float *b_o, *ln2_g, *ln2_b, *b_1, *b_2, *lnn_g, *lnn_b, *b_qkv;
  float* hmid; void* u2; float *mean2, *rstd2;  /* saved for backward */
  void *a, *m;                      /* [B*T, 4H] bf16 pre / post GELU */
  float* z_out;                     /* [B*T, H] f32 */
  void* u_next;                     /* [B*T, H] bf16, or [B, H] (CLS rows) if last */
  float *mean_n, *rstd_n;           /* [B*T], or [B] if last */
  void* qkv_next;                   /* [B*T, 3H] bf16 (unused if last) */
} vitb200_layer_fwd_args;

int vitb200_fused_supported(int H, int P);
int vitb200_fused_embed_fwd(const vitb200_embed_fwd_args* args, void* stream);
int vitb200_fused_layer_fwd(const vitb200_layer_fwd_args* args, void* stream);

/* Fused backward (bf16, H = 32): autograd of the layer kernel above, split around attention backward.
 *   upper : dz (grad of the layer output) -> MLP-down dgrad/wgrad -> gelu' -> MLP-up dgrad/wgrad -> LN2 backward
 *           (+ residual) -> dh ; out-proj dgrad/wgrad -> dctx
 *   lower : dqkv (from attention backward) -> QKV dgrad/wgrad -> LN1 backward (+ dh) -> dz of the layer input
 * Persistent CTAs (grid = vitb200_fused_bwd_grid(M)) loop over 128-row tiles; dgrad and wgrad both run on tcgen05
 * from the SAME shared-memory tiles (K-major view for dgrad, MN-major view for wgrad), weight gradients accumulate
 * in TMEM across tiles, bias / LayerNorm gradients in registers.  Each CTA writes ONE partial gradient set to
 * gpart[cta][n_opt] (same offsets as the gradient arena); vitb200_grad_reduce sums the CTA partials in CTA order
 * (deterministic) into the gradient arena. */
typedef struct {
  int B, T, H;
  float p_drop;
  const uint64_t* rng;
  uint32_t site_proj, site_mlp;
  const float* dz;                   /* [B*T, H] f32, or NULL when dz_cls is given */
  const float* dz_cls;               /* [B, H] f32: gradient of the CLS rows only (all other rows are zero), or NULL */
  const void *m, *a, *u2, *ctx;      /* saved activations (bf16) */
  const float *hmid, *mean2, *rstd2, *ln2_g;
  const void *w_2, *w_1, *w_o;       /* bf16 [H,4H], [4H,H], [H,H] */
  float* dh;                         /* [B*T, H] f32 out */
  void* dctx;                        /* [B*T, H] bf16 out */
  float* gpart;
  int n_opt, off_w2, off_b2, off_w1, off_b1, off_ln2g, off_ln2b, off_wo, off_bo;
} vitb200_layer_bwd_upper_args;

typedef struct {
  int B, T, H;
  const void *dqkv, *u;              /* [B*T, 3H], [B*T, H] bf16 */
  const float *z, *mean1, *rstd1, *ln1_g, *dh;
  const void* w_qkv;                 /* bf16 [3H, H] */
  float* dz;                         /* [B*T, H] f32 out: grad of the layer input */
  float* gpart;
  int n_opt, off_wqkv, off_bqkv, off_ln1g, off_ln1b;
} vitb200_layer_bwd_lower_args;

/* embedding backward (no learned positions, P % 16 == 0): dz0 -> dropout' -> dW_p (tcgen05, MN-major views of the
 * gradient tile and of the re-built patch tile), db_p, dcls; partials go to gpart like the layer kernels. */
typedef struct {
  int B, L, P, S, Np, n_valid, H;
  float p_drop;
  const uint64_t* rng;
  const float* dz0;                  /* [B*T, H] f32 */
  const float* x;                    /* [B, L] f32 */
  float* gpart;
  int n_opt, off_wp, off_bp, off_cls;
} vitb200_embed_bwd_args;
int vitb200_fused_embed_bwd_supported(int H, int P, int learned_pos);
int vitb200_fused_embed_bwd(const vitb200_embed_bwd_args* args, void* stream);

int vitb200_fused_bwd_supported(int H);
int vitb200_fused_bwd_grid(int M);
int vitb200_fused_layer_bwd_upper(const vitb200_layer_bwd_upper_args* args, void* stream);
int vitb200_fused_layer_bwd_lower(const vitb200_layer_bwd_lower_args* args, void* stream);
/* grad[i] = sum_{s < slots} gpart[s*stride + i] for i in [start, end), slots summed in order. */
int vitb200_grad_reduce(const float* gpart, int slots, size_t stride, size_t start, size_t end, float* grad,
                        void* stream);

/* ---- whole-network kernels (bf16, hidden 32, 2 heads of 16, 4H MLP, T = Np + 1 <= 129) --------------------------------
 * The configured model (configs/exp/att_clp/baseline.yaml, configs/config.yaml) is so small that a training step is a
 * chain of dependent launches, not bandwidth or math.  Samples are independent until the gradient reduction, so ONE
 * persistent kernel per direction gives every CTA whole samples and runs the complete network on them:
 *   forward : spectrum -> patch GEMM, CLS, (+pos), dropout -> [LN1, QKV, attention (S = QK^T, softmax, PV), out-proj,
 *             dropout, +res, LN2, MLP-up, GELU, MLP-down, dropout, +res] x layers -> final LN of the CLS row -> head ->
 *             logits, loss     (src/models/embedding.py:79-100, HF:328-346, HF:454-455, src/models/specvit.py:78-89)
 *   backward: autograd of the above for the same sample: head, final LN, every layer (MLP, LN2, out-proj, attention,
 *             QKV, LN1), embedding; parameter gradients accumulate per CTA and leave as ONE partial set per CTA
 *             (gpart[cta][n_opt], summed in CTA order by vitb200_clip_adamw_fused).
 * 128 token rows of a sample run on tcgen05 (TMEM lane = row, all weights TMA-staged in shared memory, activations
 * handed from GEMM to GEMM through swizzled shared-memory tiles); the 129th token runs beside them on a 17th warp.
 * Same math, same dropout masks and the same saved-for-backward tensors as the per-op entry points, so either
 * direction can be mixed with them.
 * Parameters are addressed inside the flat arena: `params` (fp32 master) and `shadow` (bf16 GEMM operands) with element
 * offsets; layer l's block starts at off_layer0 + l * layer_stride and o_* are offsets inside a block (q, k, v weights
 * and biases adjacent, see vit_b200/arena.py).  Saved activations are layer-major: z [layers+1, B*T, H] f32,
 * hmid [layers, B*T, H] f32, u / u2 / ctx [layers, B*T, H], qkv [layers, B*T, 3H], a / m [layers, B*T, 4H] (bf16),
 * stats [4*layers + 2, B*T] f32 (mean1, rstd1, mean2, rstd2 per layer; then mean / rstd of the final LN of the CLS rows,
 * indexed by sample), lse [layers, B, heads, T] f32. */
typedef struct {
  int B, L, P, S, Np, n_valid;       /* spectrum length L, patch P, stride S, Np patches (n_valid real windows) */
  int layers, C, loss_kind, cluster; /* encoder layers, labels, VITB200_LOSS_*, CTAs per sample (1, or 2 = one head each) */
  int cls_only;                      /* != 0: the caller only needs logits / loss / gradients, not every token's last hidden
                                      * state: the last layer then runs for the CLS row alone (specvit.py:78 reads nothing
                                      * else); z[layers] and the last layer's saved rows are valid for the CLS row only */
  float eps, p_hidden, p_attn;       /* dropout probabilities (0 in eval mode) */
  const uint64_t* rng;
  const float* x;                    /* [B, L] f32 */
  const void* labels;                /* f32 [B*C] (MSE / L1) or int64 [B] (CE); NULL: logits only */
  const float* params; const void* shadow;
  int off_cls, off_pos, off_wp, off_bp;          /* off_pos < 0: no learned positions */
  int off_layer0, layer_stride;
  int o_ln1g, o_ln1b, o_wqkv, o_bqkv, o_wo, o_bo, o_ln2g, o_ln2b, o_w1, o_b1, o_w2, o_b2;
  int off_lnfg, off_lnfb, off_wh, off_bh;
  const float *rope_cos, *rope_sin;  /* [T, 8] f32 or NULL */
  float *z, *hmid; void *u, *u2, *qkv, *ctx, *a, *m;
  float *stats, *lse;
  void* s_cls;                       /* [B, H] bf16: final-LayerNorm'd CLS rows */
  float *logits, *loss;              /* [B, C], [1] */
  void* ws;                          /* vitb200_mega_ws_bytes() bytes, zeroed once by the caller */
  /* Device-resident dataset mode (src/dataloader/base.py:219-245 hands the model rows of an in-RAM tensor; here the
   * tensor lives in HBM and the kernels pick the rows themselves).  rows == NULL: sample b reads x[b], labels[b].
   * Otherwise x / labels are the WHOLE dataset and sample b of this step reads row rows[pos * B + b], where
   * pos = rng[1] - rows_base[0] is the number of steps taken since the host stored rows_base (the optimizer kernel
   * advances rng[1]), so one captured launch walks through an epoch's permutation.  loss_log (optional): the step's
   * loss is also stored at loss_log[pos]. */
  const int64_t* rows; const uint64_t* rows_base; float* loss_log;
  /* defer_loss != 0 (training step: vitb200_mega_bwd with the same argument block is the next launch): forward only
   * stores its per-CTA loss terms in ws; the backward kernel sums them and writes loss / loss_log.  Saves the fence +
   * ticket at the end of the forward kernel, which the backward kernel's start is waiting for. */
  int defer_loss;
} vitb200_mega_fwd_args;
int vitb200_mega_supported(int H, int heads, int T, int P, int C, int layers, int rope);
size_t vitb200_mega_ws_bytes(void);
size_t vitb200_mega_fwd_smem_bytes(int layers);
int vitb200_mega_grid(int B, int cluster);
int vitb200_mega_fwd(const vitb200_mega_fwd_args* args, void* stream);
/* Backward of the above for ONE sample per CTA (pair): B * cluster <= 148.  `f` is the forward call's argument block
 * (same buffers, dropout probabilities and rng: the masks are regenerated).  labels / loss_kind as in forward, or
 * VITB200_LOSS_GIVEN with labels = d(objective)/d(logits) [B, C] f32; gloss: DEVICE scalar d(objective)/d(loss) or NULL.
 * gpart [B, n_opt] f32 receives the sample's gradient of EVERY optimised parameter at the parameter's arena offset
 * (alignment gaps are not written); dz0 [B*T, H] f32 (optional) receives d loss / d(embedding output). */
typedef struct {
  vitb200_mega_fwd_args f;
  const void* labels;
  const float* gloss;
  int loss_kind, n_opt;
  float* gpart;
  float* dz0;
  /* Streamed gradient groups (optional, NULL = off): DEVICE counters [layers + 2], zeroed once by the caller.  Every CTA
   * adds 1 to counter g when its partial gradients of group g are in gpart: g = 1 + layers (final LayerNorm + head) and
   * g = 1 + l (encoder layer l) as the layers finish, top first; g = 0 (everything in front of layer 0: cls token,
   * positions, patch projection) at the end.  vitb200_clip_adamw_fused_streamed consumes and resets them. */
  unsigned int* done;
} vitb200_mega_bwd_args;
int vitb200_mega_bwd_supported(int H, int heads, int T, int P, int C, int layers, int B, int cluster);
size_t vitb200_mega_bwd_smem_bytes(int layers);
int vitb200_mega_bwd(const vitb200_mega_bwd_args* args, void* stream);

/* ---- multi-head self-attention ----------------------------------------------------------------
 * Replaces ViTSelfAttention.forward's SDPA / eager attention (HF:171-196,232-249) and
 * ViTSelfAttentionWithRoPE.forward (src/models/vit_with_rope.py:43-84; RoPE = src/models/rope.py:60-98):
 *   P = softmax(q' k'^T * scale) ; ctx = drop(P) v ; lse = logsumexp of the scaled scores
 * q,k,v: rows b*T+t of [B*T, ld] (dtype); ctx [B*T, heads*d] (dtype); lse [B,heads,T] f32.
 * rope_cos/rope_sin [T, d/2] f32 or NULL (q' = rope(q), k' = rope(k)).  d in {8,16,32,64,128}. */
int vitb200_attn_fwd(const void* q, const void* k, const void* v, int ld, void* ctx, float* lse,
                     const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d, float scale,
                     float p_drop, const uint64_t* rng, uint32_t site, int dtype, void* stream);
/* Backward (flash style: P is recomputed from q, k, lse).  dsum [B,heads,T] f32 scratch.
 * dq, dk, dv: rows of [B*T, ld_d] (dtype). */
int vitb200_attn_bwd(const void* q, const void* k, const void* v, int ld, const void* ctx, const void* dctx,
                     const float* lse, float* dsum, void* dq, void* dk, void* dv, int ld_d, const float* rope_cos,
                     const float* rope_sin, int B, int T, int heads, int d, float scale, float p_drop,
                     const uint64_t* rng, uint32_t site, int dtype, void* stream);
/* tcgen05 versions (bf16, head_dim 16 or 32, T <= 176 so that all keys of a (sample, head) fit one TMEM tile):
 * S = Q K^T and O = P V (and dP, dQ, dK, dV in backward) run on the tensor cores with TMEM accumulators, Q/K/V tiles
 * are TMA-loaded in place from the fused [B*T, 3H] QKV buffer (`qkv` = q pointer; k, v at +H, +2H; ld = 3H), softmax
 * is done in registers.  vitb200_attn_fwd / _bwd above choose them automatically when the layout allows. */
int vitb200_set_attn_mode(int mode); /* 0 = automatic, 1 = SIMT only (A/B tests); returns the old mode */
int vitb200_attn_tc_supported(int T, int d, int ld, int H);
int vitb200_attn_tc_fwd(const void* qkv, void* ctx, float* lse, const float* rope_cos, const float* rope_sin, int B,
                        int T, int heads, int d, float scale, float p_drop, const uint64_t* rng, uint32_t site,
                        void* stream);
int vitb200_attn_tc_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                        const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d, float scale,
                        float p_drop, const uint64_t* rng, uint32_t site, void* stream);

/* Per-key-block tcgen05 attention BACKWARD for any sequence length (bf16, d in {16, 32}, fused [B*T, 3H] QKV buffer): the
 * single-tile backward kernel launched once per block of 128 keys, dQ accumulated over the launches in fp32.  Kept as the
 * independent cross-check of vitb200_attn_flash_bwd (bit-identical results); the engine uses the flash kernels.
 * ws: vitb200_attn_tc_blocked_ws_bytes(...) bytes, 16-byte aligned. */
int vitb200_attn_tc_blocked_supported(int T, int d, int ld, int H);
size_t vitb200_attn_tc_blocked_ws_bytes(int B, int T, int heads, int d);
int vitb200_attn_tc_blocked_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                                const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d,
                                float scale, float p_drop, const uint64_t* rng, uint32_t site, void* ws, void* stream);
/* Flash attention forward for ANY sequence length in ONE launch (bf16, d in {16, 32, 64}, fused [B*T, 3H] QKV buffer): the
 * key/value loop runs inside the kernel (TMA double-buffered K/V blocks of 128 keys, S = QK^T and P~V on tcgen05, online
 * softmax with the running max / sum and the output row in registers).  No workspace, no partial outputs, no merge
 * kernel.  Same dropout masks, RoPE and lse output as vitb200_attn_tc_fwd.  (HF:232-249, vit_with_rope.py:43-84) */
int vitb200_attn_flash_supported(int T, int d, int ld, int H);
int vitb200_attn_flash_fwd(const void* qkv, void* ctx, float* lse, const float* rope_cos, const float* rope_sin, int B,
                           int T, int heads, int d, float scale, float p_drop, const uint64_t* rng, uint32_t site,
                           void* stream);
/* Backward of the above in ONE launch for any T (+ one small reduction launch): one CTA per (block of 128 keys, head,
 * sample) keeps dK / dV in tensor memory while it walks over the query tiles; dQ leaves as fp32 partials per key block
 * (ws) and is summed in block order.  dqkv [B*T, 3H] bf16 receives dq | dk | dv.  Same dropout masks as forward. */
size_t vitb200_attn_flash_bwd_ws_bytes(int B, int T, int heads, int d);
int vitb200_attn_flash_bwd(const void* qkv, const void* ctx, const void* dctx, const float* lse, void* dqkv,
                           const float* rope_cos, const float* rope_sin, int B, int T, int heads, int d, float scale,
                           float p_drop, const uint64_t* rng, uint32_t site, void* ws, void* stream);
/* probs[B,heads,T,T] f32 = softmax probabilities before dropout (what the eager / RoPE attention returns
 * as `attention_probs`, src/models/vit_with_rope.py:84); for hooks / output_attentions only. */
int vitb200_attn_probs(const void* q, const void* k, int ld, const float* lse, float* probs, const float* rope_cos,
                       const float* rope_sin, int B, int T, int heads, int d, float scale, int dtype, void* stream);

/* ---- head + loss ---------------------------------------------------------------------------------
 * Replaces MyViT.forward's head and loss (src/models/specvit.py:78-89):
 *   logits = s . w^T + bias ; loss = MSE | L1 (mean over B*C) | CrossEntropy (mean over B)
 * s [B,H] (dtype) = final-LayerNorm'd CLS rows; w [C,H] (dtype); bias [C] f32; labels: f32 [B*C] for
 * MSE/L1, int64 [B] for CE (NULL: logits only); logits [B,C] f32; loss [1] f32. */
int vitb200_head_loss_fwd(const void* s, const void* w, const float* bias, const void* labels, float* logits,
                          float* loss, int B, int H, int C, int loss_kind, int dtype, void* stream);
/* gloss: DEVICE scalar d(objective)/d(loss) or NULL (=1).  ds [B,H] (dtype); dw [C,H], dbias [C] f32. */
int vitb200_head_loss_bwd(const void* s, const void* w, const float* logits, const void* labels,
                          const float* gloss, void* ds, float* dw, float* dbias, int B, int H, int C, int loss_kind,
                          int accumulate, int dtype, void* stream);

/* Single-launch variants for small heads (H <= 128, C <= 4): logits + loss in one kernel; head backward fused with the
 * final-LayerNorm backward of the CLS rows (HF:455): z = residual stream after the last layer (row b at
 * z + b*z_row_stride), mean/rstd [B] from the forward; outputs dz_cls [B,H] f32 (gradient of the CLS rows of the last
 * layer's output; every other row's gradient is zero), dgamma/dbeta of vit.layernorm, dw/dbias of the head. */
int vitb200_head_fused_supported(int H, int C);
int vitb200_head_fused_fwd(const void* s, const void* w, const float* bias, const void* labels, float* logits,
                           float* loss, int B, int H, int C, int loss_kind, int dtype, void* stream);
int vitb200_head_fused_bwd(const void* s, const void* w, const float* logits, const void* labels, const float* gloss,
                           const float* z, size_t z_row_stride, const float* mean, const float* rstd, const float* gamma,
                           float* dz_cls, float* dgamma, float* dbeta, float* dw, float* dbias, int B, int H, int C,
                           int loss_kind, int accumulate, int dtype, void* stream);
/* forward + backward of the head in one launch (training steps: dloss = 1).  Same results as the two calls above. */
int vitb200_head_fused_fwd_bwd(const void* s, const void* w, const float* bias, const void* labels, float* logits,
                               float* loss, const float* z, size_t z_row_stride, const float* mean, const float* rstd,
                               const float* gamma, float* dz_cls, float* dgamma, float* dbeta, float* dw, float* dbias,
                               int B, int H, int C, int loss_kind, int dtype, void* stream);

/* ---- gradient clipping + AdamW ---------------------------------------------------------------------
 * Replaces Lightning's gradient_clip_val -> torch.nn.utils.clip_grad_norm_ (src/basemodule.py:244) and
 * torch.optim.AdamW.step (src/opt/optimizer.py:108) over ONE flat parameter arena.
 * hyper (DEVICE, 8 floats): {lr, beta1, beta2, eps, weight_decay, max_norm (<=0: no clipping), grad_scale, 0}
 * state (DEVICE, 8 floats): {step, grad_norm, clip_coef, bias_corr1, bias_corr2, extra squared norm (see
 * vitb200_sumsq_accum), 0, peer time-out flag}
 * vitb200_grad_norm: grad_norm = ||grad_scale * g||_2, clip_coef = min(1, max_norm/(norm+1e-6)), step += 1.
 * vitb200_adamw: p, m, v updated in place with g * grad_scale * clip_coef; if shadow != NULL it receives
 * the bf16 copy of the new parameters (the GEMM operands of BF16 mode).  rng != NULL: rng[1] += 1. */
size_t vitb200_grad_norm_ws_bytes(size_t n);
int vitb200_grad_norm(const float* g, size_t n, const float* hyper, float* state, void* ws, void* stream);
int vitb200_adamw(float* p, const float* g, float* m, float* v, void* shadow, size_t n, const float* hyper,
                  const float* state, uint64_t* rng, void* stream);
/* acc[0] += sum_i g[i]^2 (deterministic).  For parameters that live outside the flat arena (a trainable preprocessor
 * matrix, src/models/layers.py:51-60): accumulate into state[5] before vitb200_clip_adamw_fused*, which adds it to the
 * arena's squared norm -- torch's clip_grad_norm_ runs over ALL parameters (src/basemodule.py:244) -- and clears it;
 * afterwards vitb200_adamw updates those parameters with the same clip coefficient and bias corrections (state[2..4]).
 * ws: vitb200_grad_norm_ws_bytes(n) bytes, first 4096 zeroed once. */
int vitb200_sumsq_accum(const float* g, size_t n, float* acc, void* ws, void* stream);
/* vitb200_clip_adamw_fused: the whole optimizer tail in ONE launch (the configured model has 40 353 parameters: the
 * tail is three launch latencies, not bandwidth).  If slots > 0, the gradient of elements [red_start, red_end) is first
 * formed as a fixed-order sum of `slots` partial arenas (gpart + s*stride; the per-CTA partials of the fused backward
 * kernels) and written to g; then norm -> clip_coef -> AdamW exactly as the two calls above.  ws: at least
 * vitb200_clip_adamw_fused_ws_bytes() bytes, zeroed once by the caller (grid-barrier ticket / epoch + block partials). */
size_t vitb200_clip_adamw_fused_ws_bytes(void);
int vitb200_clip_adamw_fused(float* p, float* g, float* m, float* v, void* shadow, size_t n, const float* hyper,
                             float* state, uint64_t* rng, const float* gpart, int slots, size_t stride,
                             size_t red_start, size_t red_end, void* ws, void* stream);

/* Streamed variant of the tail (single GPU: peer_bufs == NULL, world == 1; data parallel otherwise, as _dp below): instead
 * of waiting for the backward kernel to finish, every block waits only for the gradient groups its slice of the arena
 * belongs to (counters raised by vitb200_mega_bwd, see vitb200_mega_bwd_args.done), so the partial-sum reduction -- and
 * in a data-parallel run the NVLink exchange -- of the upper layers' gradients runs WHILE the backward kernel is still
 * working on the lower layers (the bucketed overlap of DDP, without a collective launch).  Must directly follow the
 * vitb200_mega_bwd launch that raises the counters, on the same stream; it resets the counters to zero.
 * gs->lo / hi: arena element range [lo, hi) of each group (multiples of 4; together they cover [0, n)). */
#define VITB200_MAX_GROUPS 20
typedef struct {
  unsigned int* done;     /* DEVICE counters [n_groups] */
  unsigned int expect;    /* arrivals per group = CTAs of the backward launch */
  int n_groups;
  unsigned int lo[VITB200_MAX_GROUPS], hi[VITB200_MAX_GROUPS];
} vitb200_grad_stream;
int vitb200_clip_adamw_fused_streamed(float* p, float* g, float* m, float* v, void* shadow, size_t n, const float* hyper,
                                      float* state, uint64_t* rng, const float* gpart, int slots, size_t stride,
                                      size_t red_start, size_t red_end, void* ws, const vitb200_grad_stream* gs,
                                      void* const* peer_bufs, int rank, int world, void* stream);

/* Data-parallel optimizer tail: the DDP gradient all-reduce (implicit in the reference: strategy='ddp',
 * src/hardware_utils.py:86-95) is fused INTO the kernel.  Each rank owns one exchange buffer
 * (vitb200_peer_buffer_bytes(n) bytes: two launch parities x 8 source ranks x n slots of {value, launch tag}) allocated
 * with vitb200_peer_alloc, which also returns a 64-byte CUDA IPC handle; the ranks of one node exchange handles (host side)
 * and map each other's buffers with vitb200_peer_open (NVLink / NVSwitch peer access).  peer_bufs is a DEVICE array of
 * `world` (2 .. 8) pointers in rank order (entry `rank` = the local buffer).  Per launch: every locally reduced gradient
 * element is PUSHED, tagged with the launch number, into every rank's buffer with one 8-byte store (value and tag travel
 * together: no flag, no fence, no acknowledgement); each rank polls its own buffer until all ranks' tags of an element
 * match and sums the values in rank order (bit-identical on every rank), then norm -> clip -> AdamW as
 * vitb200_clip_adamw_fused.  Every rank must launch it the same number of times.  grad_scale (hyper[6]) carries the
 * 1/world mean. */
size_t vitb200_peer_buffer_bytes(size_t n);
int vitb200_peer_alloc(size_t bytes, void** ptr, unsigned char* handle64);
int vitb200_peer_open(const unsigned char* handle64, void** ptr);
int vitb200_peer_close(void* ptr);
int vitb200_peer_free(void* ptr);
int vitb200_clip_adamw_fused_dp(float* p, float* g, float* m, float* v, void* shadow, size_t n, const float* hyper,
                                float* state, uint64_t* rng, const float* gpart, int slots, size_t stride,
                                size_t red_start, size_t red_end, void* ws, void* const* peer_bufs, int rank, int world,
                                void* stream);

/* shadow[i] = bf16(p[i]) (after load_state_dict / an external optimizer touched the fp32 arena) */
int vitb200_cast_bf16(const float* p, void* shadow, size_t n, void* stream);

/* ---- elementwise helpers of the modular (hook-friendly, eval-only) module path ------------------------
 * y = gelu(x) (HF:298 intermediate_act_fn as a separate module call); out = z + delta (HF:337,311). */
int vitb200_gelu_fwd(const void* x, void* y, size_t n, int dtype, void* stream);
int vitb200_residual_add(const float* z, const void* delta, float* out, size_t n, int dtype, void* stream);

/* ---- either side of the step: input stage, dataset hand-off, evaluation metrics --------
 * dst[i] = float(src_bf16[i]).  The input preprocessor (LinearPreprocessor / PrefilledLinear.forward,
 * src/models/preprocessor.py:107-108, src/models/layers.py:62-63; PrefilledAttention 2-D branch,
 * src/models/attention.py:81-82) is vitb200_linear_fwd on [B, D_in] x [D_out, D_in]^T; in BF16 mode its bf16 output is
 * widened with this call into the fp32 pixel buffer the embedding kernels read. */
int vitb200_cast_f32(const void* src, float* dst, size_t n, void* stream);
/* The same Linear specialised for the preprocessor's shape (M = batch rows << N, K: the 4096 x 4096 ZCA matrix is 32 MB of
 * bf16 that every step has to stream): bf16 x [M,K] and w [N,K] on tcgen05 with the contraction split across CTAs so
 * that one wave of <= 148 CTAs pulls the matrix at HBM rate; deterministic fixed-order sum of the splits; output
 * y [M,N] fp32 holding bf16-rounded values (autocast's bf16 output, widened) written straight into the pixel buffer.
 * N % 8 == 0, K % 8 == 0, 16-byte aligned pointers.  ws: vitb200_tc_prelinear_ws_bytes() bytes, first 4096 zeroed once. */
size_t vitb200_tc_prelinear_ws_bytes(int M, int N, int K);
int vitb200_tc_prelinear_fwd(const void* x, const void* w, const float* bias, float* y, int M, int N, int K, void* ws,
                             void* stream);
/* Low-rank ZCA (src/models/preprocessor.py:40-72: P = Vr diag(1/sqrt(lam_r + eps)) Vr^T + s_perp (I - Vr Vr^T), applied as
 * a dense Linear by src/models/layers.py:62-63) in factored form while the matrix is frozen:
 *     y = s_perp x + ((x Vr) o g) Vr^T + bias,   g = 1/sqrt(lam_r + eps) - s_perp
 * two skinny products in ONE launch (one thread-block cluster per spectrum), 2 r D values of Vr instead of D^2 of P.
 * x [B, D] f32; vr [D, R] bf16 (dtype BF16: bf16 operands, fp32 accumulation, bf16-rounded fp32 output) or f32; R in
 * {32, 64}, columns beyond the true rank zero; g [R] f32; bias [D] f32 or NULL; y [B, D] f32. */
int vitb200_zca_lowrank_supported(int D, int R);
int vitb200_zca_lowrank_fwd(const float* x, const void* vr, const float* g, float s_perp, const float* bias, float* y,
                            int B, int D, int R, int dtype, void* stream);
/* Gradient of the patch embedding w.r.t. its input pixels -- what autograd hands to a TRAINABLE preprocessor
 * (PrefilledLinear.freeze(False), src/models/layers.py:51-60):
 *   dx[b, l] = sum_{n : n*S <= l < n*S+P, n < n_valid} sum_h drop'(dz[b, 1+n, h]) * w[h, l - n*S]
 * dz [B, Np+1, H] f32 = gradient w.r.t. the embedding output (same tensor vitb200_patch_embed_bwd takes, same dropout
 * site); w [H, P] (dtype); dx [B, L] f32, fully written. */
int vitb200_patch_embed_dgrad(const float* dz, const void* w, float* dx, int B, int L, int P, int S, int Np,
                              int n_valid, int H, float p_drop, const uint64_t* rng, uint32_t site, int dtype,
                              void* stream);
/* Batch assembly from a DEVICE-resident dataset (replaces DataLoader collation + H2D of the in-RAM tensors of
 * src/dataloader/base.py:219-245,299-300 and the noise injection of src/vit.py:86-88):
 *   x[i, :] = flux[idx[i], :] (+ N(0,1) * error[idx[i], :] * noise_level when noise_level > 0 and error != NULL)
 *   y[i]    = labels[idx[i]]   (rows of label_bytes raw bytes: fp32 [C] or int64; labels/y may both be NULL)
 * flux, error [n_rows, L] f32 (L % 4 == 0); idx [B] int64 on the device (NULL: rows 0..B-1); rng = {seed, step}. */
int vitb200_gather_batch(const float* flux, const float* error, const void* labels, const int64_t* idx, float* x,
                         void* y, int B, int L, int label_bytes, long long n_rows, float noise_level,
                         const uint64_t* rng, void* stream);
/* Running evaluation metrics on the device (replaces the per-batch torchmetrics updates of src/vit.py:94-125:
 * MeanAbsoluteError, MeanSquaredError, R2Score, Accuracy).  acc: DEVICE doubles, zeroed by the caller at epoch start,
 * at least 2 + 4*C entries:  acc[0] += B, acc[1] += B*loss[0] (loss may be NULL),
 *   regression     (is_cls = 0, labels f32 [B, C], C <= 16): acc[2+4c+{0,1,2,3}] += sum_b {|e|, e^2, y, y^2}, e = logits - y
 *   classification (is_cls = 1, labels int64 [B]):           acc[2] += #(argmax_c logits[b, c] == labels[b])
 * One CTA, fixed summation order (deterministic). */
int vitb200_eval_metrics_accum(const float* logits, const void* labels, const float* loss, double* acc, int B, int C,
                               int is_cls, void* stream);

/* ---- test support ----------------------------------------------------------------------------------
 * mask[i] = 1 if element i of dropout site `site` is kept.  For the attention site the element index
 * is ((b*heads+h)*T + i) * round_up(T,4) + j; for every other site it is the linear index. */
int vitb200_dropout_mask(uint8_t* mask, size_t n, float p_drop, const uint64_t* rng, uint32_t site, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* VIT_B200_H */
