/*
 * coponerf_b200 -- C-ABI of the B200-native CoPoNeRF render path.
 *
 * The reference (cvlab-kaist/CoPoNeRF) is pure Python over PyTorch and has no FFI of its
 * own; the "interface each entry point replaces" is therefore the Python call site in
 * models/CoPoNeRF.py that the host-side mirror (coponerf_b200/model.py) routes here.
 * Citations are relative to the reference root.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer to a caller-owned, contiguous buffer (fp32 unless
 *     the name says otherwise); the library allocates no device memory. Its only process-wide state: a per-device
 *     pool of internal streams / events for the chunk lanes of cpn_render_rays (created on first use with lanes > 1,
 *     released by cpn_shutdown) and the optional profiling hooks (cpn_prof_*, cpn_gemm_tc_trace). cpn_render_rays is
 *     therefore single-threaded per device: one host thread (and one caller stream) at a time, as the reference's
 *     callers are;
 *   - every call is asynchronous on `stream` (a cudaStream_t passed as void*): no host
 *     sync, no host copies;
 *   - return value: 0 on success, a negative cpn_status otherwise; cpn_last_error()
 *     returns a per-thread message.
 */
#ifndef COPONERF_B200_H
#define COPONERF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
  CPN_OK = 0,
  CPN_ERR_ARG = -1,        /* bad shape / null pointer / unsupported size */
  CPN_ERR_WORKSPACE = -2,  /* workspace too small */
  CPN_ERR_CUDA = -3        /* a CUDA runtime call or launch failed */
} cpn_status;

#define CPN_FLAG_EARLY_V 8    /* form V = latent_value(...) per sample (one GEMM over all sample rows) and let the attention
                               * read it; default: the attention reads out the hidden layer and the (linear) folded
                               * latent_value runs once per ray ("late readout") */
#define CPN_FLAG_NO_BILINEAR 16 /* keep key_map_2 / query_embed_2 / query_repeat_embed_2 as three 128 x 128 layers (default:
                                 * both attention logits are evaluated as bilinear forms, one layer per round) */
#define CPN_N_LEVELS 4      /* feature maps per view: 3 refined ResNet levels + conv_map */
#define CPN_FEAT_DIM 832    /* 256*3 + 64, models/CoPoNeRF.py:68 */
#define CPN_LATENT 416      /* latent_dim // 2, models/CoPoNeRF.py:74 */
#define CPN_HIDDEN 128      /* hidden_dim, models/CoPoNeRF.py:78 */
#define CPN_PAIR_CONSTS_FLOATS 320
#define CPN_FLAG_SIMT_ONLY 1  /* run the big 1x1 convs on the fp32 CUDA-core GEMM (cross-check path) */
#define CPN_FLAG_F16X3 2      /* tensor-core GEMMs with three fp16 MMAs per product (default: fp16 + 2 fp8) */
#define CPN_FLAG_NO_GFOLD 32  /* keep latent_value -> encode_latent -> query_repeat_embed as a per-ray chain behind a separate
                               * round-1 readout (default: folded into one 128 x 1664 map applied per sample row next to
                               * key_map, one combined readout with weights w2 + 2 w1 for z = R2 + 2 R1) */
#define CPN_FLAG_FULL_H1 64   /* write the hidden-layer image with its value plane (4 bytes per element; default on the default
                               * path: 3 bytes, the key GEMM derives the plane in shared memory). Same bits either way */
#define CPN_FLAG_NO_FOLD 4    /* keep query_encode_latent_2, latent_value and key_map as three GEMMs (default: the
                               * activation-free query_encode_latent_2 is folded into the other two at pack time) */

int cpn_version(void);
int cpn_shutdown(void);                /* destroys the lane streams / events of every device; re-created on demand */
const char* cpn_last_error(void);
size_t cpn_sizeof_render_args(void);   /* ABI check for FFI bindings */

/* ---- one-time weight repack ------------------------------------------------------------
 * Replaces: nn.Module parameter storage of the render-path layers
 * (models/CoPoNeRF.py:71-104, models/lightfield.py:87-116). `src` holds the fp32
 * state_dict tensors back to back in the order of cpn_weight_names(); `dst` receives the
 * kernel layouts (transposed fp32 copies and split-bf16 tensor-core tiles). */
size_t cpn_packed_weights_bytes(void);
size_t cpn_raw_weights_floats(void);
int cpn_n_weight_tensors(void);
const char* cpn_weight_name(int i);    /* state_dict key */
size_t cpn_weight_numel(int i);
int cpn_pack_weights(const float* src, void* dst, void* stream);

/* ---- per-pair setup ----------------------------------------------------------------------
 * cpn_pack_features: NCHW (2B, C, h, w) -> channels-last (2B, h, w, C), one call per level.
 * Replaces the implicit layout F.grid_sample reads (models/CoPoNeRF.py:312,370). */
int cpn_pack_features(const float* nchw, float* nhwc, int n_img, int C, int h, int w, void* stream);

/* cpn_pair_setup: relative poses and intrinsics used by every ray of a pair.
 * Replaces models/CoPoNeRF.py:239-244,259-261,325-332,572-575.
 *   ctx_c2w (B,2,4,4) ctx_K (B,2,4,4) qry_c2w (B,4,4) qry_K (B,4,4) rel_pose (B,4,4)
 *   consts  (B, CPN_PAIR_CONSTS_FLOATS)  out */
int cpn_pair_setup(const float* ctx_c2w, const float* ctx_K, const float* qry_c2w, const float* qry_K,
                   const float* rel_pose, int B, int H, int val, float* consts, void* stream);

/* cpn_pair_prologue: flow upsampling and the cycle-consistency mask.
 * Replaces models/CoPoNeRF.py:230-236 (+ the F.interpolate of utils.flow2kps, utils.py:55).
 *   flow0, flow1 (B,2,fh,fw)  ->  up_flow2 (B,2,256,256) UNSCALED bilinear upsample of flow1,
 *   mask_padded2 (B,256,256) uint8.  rgb_w = context rgb.shape[-2]. */
int cpn_pair_prologue(const float* flow0, const float* flow1, int B, int fh, int fw, int rgb_w,
                      float* up_flow2, uint8_t* mask_padded2, void* stream);

/* ---- the per-ray hot path ---------------------------------------------------------------
 * Replaces the body of CoPoNeRF.forward() from models/CoPoNeRF.py:246 to :566:
 * plucker rays, project_rays (models/epipolar.py:175-253), sample positions, primary and
 * secondary bilinear gathers, fp64 triangulation (utils_training/geometry.py:98-162), the
 * per-sample encoder, two rounds of 2S-way softmax attention, phi (models/lightfield.py:
 * 131-167) and the auxiliary depth / correspondence outputs. */
typedef struct {
  int32_t B;          /* stereo pairs */
  int32_t N;          /* target rays per pair */
  int32_t S;          /* samples per epipolar line (npoints): 32, 64, 96 or 128 */
  int32_t H, W;       /* context image size (model.H, model.W) */
  int32_t flow_h;     /* height of flow[1] (for flow2kps' scale 256/flow_h) */
  int32_t chunk_rays; /* rays processed per internal pass (workspace is sized from it) */
  int32_t flags;      /* CPN_FLAG_* */
  int32_t lanes;      /* chunks in flight on internal streams, 1..4 (workspace scales with it) */
  int32_t reserved;
  /* inputs */
  const float* feat[CPN_N_LEVELS];   /* channels-last maps (2B, h_l, w_l, C_l) */
  int32_t feat_h[CPN_N_LEVELS], feat_w[CPN_N_LEVELS], feat_c[CPN_N_LEVELS];
  const float* pair_consts;          /* (B, CPN_PAIR_CONSTS_FLOATS) from cpn_pair_setup */
  const float* uv;                   /* (B, N, 2) pixel (x, y) */
  const float* interval;             /* (S) = linspace(0, 1, S) */
  const void* weights;               /* from cpn_pack_weights */
  const float* up_flow2;             /* (B,2,256,256) from cpn_pair_prologue */
  const uint8_t* mask_padded2;       /* (B,256,256) */
  /* outputs (shapes as in the reference's out_dict) */
  float* rgb;            /* (B, N, 3) */
  float* valid_mask;     /* (B, N) */
  float* depth_ray;      /* (B, N) */
  float* at_wt;          /* (2B, N, S) round-1 attention weights */
  int64_t* at_wt_max;    /* (2B, N) */
  float* pixel_val;      /* (2B, N, S, 2) */
  float* coords;         /* (2B, N, 9) */
  float* T_to_C1_pts;    /* (B, N, 2) */
  float* T_to_C2_pts;    /* (B, N, 2) */
  float* C2_pts_to_C1;   /* (B, N, 2) */
  uint8_t* mask_c2;      /* (B, N) */
  uint8_t* matchability_cycle_mask; /* (B, N) */
  void* workspace;
  size_t workspace_bytes;
} cpn_render_args;

size_t cpn_render_workspace_bytes(int B, int N, int chunk_rays, int S, int lanes);   /* enough for every flag combination */
/* the same for one flag combination: the default path (flags 0) leaves out the buffers only the A/B paths read */
size_t cpn_render_workspace_bytes_for(int B, int N, int chunk_rays, int S, int lanes, int flags);
int cpn_render_rays(const cpn_render_args* args, void* stream);
/* number of kernels one cpn_render_rays call launches (for bench.py's gpu_launches) */
int cpn_render_launch_count(const cpn_render_args* args);

/* ---- closing stage of the per-pair cost aggregation ---------------------------------------------------
 * Replaces models/aggregation.py:527,539,549-561 (the three correlation_token calls, interpolate4d x3, their mean,
 * soft_argmax x2 and unnormalise_and_convert_mapping_to_flow x2 of UFC.forward):
 *   src[l], trg[l]  refined token features of pyramid level l, (B, sizes[l]^2, C) fp32 ("B (H W) C")
 *   lin             linspace(-1, 1, out), (out) fp32
 *   c               (B, out, out, out, out) = UFC's third return value [b, hs, ws, ht, wt]
 *   flow, flow_flip, flow_t_to_s, flow_s_to_t   (B, 2, out, out) = UFC's flow tuple, in that order */
typedef struct {
  int32_t B, C, out;
  int32_t sizes[3];
  const float* src[3];
  const float* trg[3];
  const float* lin;
  float* c;
  float* flow;
  float* flow_flip;
  float* flow_t_to_s;
  float* flow_s_to_t;
  void* workspace;
  size_t workspace_bytes;
} cpn_ufc_tail_args;
size_t cpn_ufc_tail_workspace_bytes(int B, int C, int out, const int* sizes);
int cpn_ufc_tail(const cpn_ufc_tail_args* args, void* stream);

/* ---- centre-pivot 4-D convolution block ----------------------------------------------------------------------
 * Replaces models/conv4d.py:57-135 (Conv4d, with its MaxPool4d for stride > 1) and, with norm_relu, the
 * GroupNorm(1, Co) + ReLU that follow it in Encoder4D (conv4d.py:138-163).
 *   x (B, Ci, Hq, Hq, Hs, Hs); wq, ws (Co, Ci, k, k) = query_conv / supp_conv weights; bq, bs (Co);
 *   gamma, beta (Co) = GroupNorm affine; y (B, Co, oq, oq, os, os), o = (H + 2 pad - k) / stride + 1; Co is 8 or 32. */
typedef struct {
  int32_t B, Ci, Co, Hq, Hs, k, stride, pad, norm_relu, reserved;
  const float* x;
  const float* wq;
  const float* bq;
  const float* ws;
  const float* bs;
  const float* gamma;
  const float* beta;
  float* y;
  void* workspace;
  size_t workspace_bytes;
} cpn_conv4d_args;
size_t cpn_conv4d_workspace_bytes(int B, int Hq, int Hs, int k, int stride, int pad);
int cpn_conv4d(const cpn_conv4d_args* args, void* stream);

/* ---- linear attention ------------------------------------------------------------------------------------------
 * Replaces LinearAttention.forward (models/aggregation.py:84-117): q (N, L, H, D), k (N, S, H, D), v (N, S, H, Dv)
 * -> out (N, L, H, Dv), elu(x)+1 feature map, D = 32. */
size_t cpn_linear_attention_workspace_bytes(int N, int H, int Dv);
int cpn_linear_attention(const float* q, const float* k, const float* v, int N, int L, int S, int H, int D, int Dv,
                         float* out, void* workspace, size_t workspace_bytes, void* stream);

/* ---- small operators of the cost aggregation (models/aggregation.py) ----------------------------------------
 * Tokens are (B, L = n*n, C) row-major; correlation volumes (B, H, hs, hs, q, q).
 * cpn_layernorm        nn.LayerNorm(C), eps 1e-5 (aggregation.py:257-263)
 * cpn_corr_to_tokens   'B H Hs Ws Ht Wt -> B (H Ht Wt) Hs Ws', bilinear (align_corners) to n x n, '-> B (Hs Ws) C'
 *                      (aggregation.py:283-285,291-293); written at column col0 of rows of length ld
 * cpn_tokens_to_corr   the inverse path (aggregation.py:298-300)
 * cpn_transpose_pq     batched (P x Q) -> (Q x P); 'B H Hs Ws Ht Wt -> B H Ht Wt Hs Ws' with P = Q = hs*hs (:344-346)
 * cpn_dwconv_gelu      DWConv 3x3 + nn.GELU on the n x n token map (aggregation.py:18-29,186-187)
 * cpn_resample_tokens  mode 0 interpolate2d_token (:58-63), 1 einops mean-pool (:316-319), 2 einops repeat (:327-332)
 * cpn_cross_attention  softmax(corr, -1) @ trg_v and softmax(corr, -2)^T @ src_v per head (:314,324-325)
 * cpn_correlation      cosine correlation of token features (:70-80), out (B, L, L) */
int cpn_layernorm(const float* x, const float* gamma, const float* beta, float* y, int tokens, int C, void* stream);
int cpn_corr_to_tokens(const float* corr, float* tok, int B, int H, int hs, int q, int n, int ld, int col0, void* stream);
int cpn_tokens_to_corr(const float* tok, float* corr, int B, int H, int hs, int q, int n, void* stream);
int cpn_transpose_pq(const float* in, float* out, int batch, int P, int Q, void* stream);
int cpn_dwconv_gelu(const float* x, const float* w, const float* bias, float* y, int B, int n, int C, void* stream);
int cpn_resample_tokens(const float* x, float* y, int B, int n, int m_or_pool, int C, int mode, void* stream);
int cpn_cross_attention(const float* corr, const float* src_v, const float* trg_v, float* src_attn, float* trg_attn, int B,
                        int H, int S, int T, int D, void* stream);
size_t cpn_correlation_workspace_bytes(int B, int L, int C);
int cpn_correlation(const float* src, const float* trg, float* out, int B, int L, int C, void* workspace,
                    size_t workspace_bytes, void* stream);

/* ---- per-pair pose features and pose head ---------------------------------------
 * cpn_dual_softmax   P = softmax(c, -1) * softmax(c, -2) for c (B, L, L): attn_fundamental_1 of CrossAttention.forward
 *                    (models/backbone.py:282-291); attn_fundamental_2 is its transpose and is never formed. L % 4 == 0.
 * cpn_gemm_tn        C[Mi, Nj] = act(A^T B + bias), A (L, Mi) with row stride lda, B (L, Nj) with row stride ldb; the L
 *                    range is split over CTAs and reduced in a fixed order. Replaces the v^T (attn v) products of
 *                    backbone.py:311-312 and proj_fundamental applied to the transposed matrix (:318-324).
 * cpn_linear_skinny  y[M, N] = act(x[M, K] W[N, K]^T + bias) for M <= 8 rows and long K, W as stored in the state_dict
 *                    (pose_regressor[0], models/CoPoNeRF.py:34: 134 144 -> 512).
 * cpn_pose_head      pose_regressor[2:], [:, :128], rotation_regressor, translation_regressor, r6d2mat and the 4 x 4
 *                    assembly of models/CoPoNeRF.py:34-52,106-128,198-204: h0 (B, 512) -> rel_pose (B, 4, 4).
 * act: 0 none, 1 ReLU, 2 exact GELU. */
size_t cpn_dual_softmax_workspace_bytes(int B, int L);
int cpn_dual_softmax(const float* c, float* P, int B, int L, void* workspace, size_t workspace_bytes, void* stream);
size_t cpn_gemm_tn_workspace_bytes(int Mi, int Nj, int L);
int cpn_gemm_tn(const float* A, int lda, const float* B, int ldb, const float* bias, float* C, int ldc, int Mi, int Nj,
                int L, int act, void* workspace, size_t workspace_bytes, void* stream);
size_t cpn_linear_skinny_workspace_bytes(int M, int N, int K);
int cpn_linear_skinny(const float* x, const float* W, const float* bias, float* y, int M, int N, int K, int act,
                      void* workspace, size_t workspace_bytes, void* stream);
typedef struct {
  int32_t B, reserved;
  const float* h0;                       /* (B, 512) = relu(pose_regressor[0](pose_feat)) */
  const float *w2, *b2, *w4, *b4;        /* pose_regressor.2 (256, 512), pose_regressor.4 (256, 256) */
  const float *rw1, *rb1, *rw3, *rb3, *rw5, *rb5;   /* rotation_regressor.{1,3,5} */
  const float *tw1, *tb1, *tw3, *tb3, *tw5, *tb5;   /* translation_regressor.{1,3,5} */
  float* rel_pose;                       /* (B, 4, 4) out */
} cpn_pose_head_args;
int cpn_pose_head(const cpn_pose_head_args* args, void* stream);

/* ---- device timing of the dominant kernel (the query_encode_latent GEMM), for roofline reports.
 * Between cpn_prof_begin and cpn_prof_end every launch of that kernel by cpn_render_rays is bracketed
 * by CUDA events on the caller's stream. cpn_prof_end waits for them and returns the summed duration.
 * One session per process at a time; not for production use. */
int cpn_prof_begin(int max_launches);
int cpn_prof_end(float* total_ms, int* launches);

/* The epipolar feature gather alone (for unit tests): F.grid_sample(bilinear, align_corners=False) of the four channels-last
 * maps in a->feat at the coordinates of `rowaux` (rows, 8) = [gx, gy of the primary branch ('border' padding, view v of the
 * row), gx, gy of the secondary branch ('zeros' padding, view 1 - v), 4 unused], rows = B * nr * 2 * S ordered
 * ((b * nr + n) * 2 + v) * S + s  (models/CoPoNeRF.py:312,370). Only B, S and the feat* fields of `a` are read.
 * form 0: out = fp32 rows of 848 (835 used; row index enc_row = (row / 128) * 256 + branch * 128 + row % 128);
 * form 1 / 2: out = the operand image (K = 864) of the f16x3 / f16+f8 scheme; form 3: the f16+f8 image in compact
 * 12 KB blocks (fp16 head + remainder plane, no value plane: what cpn_render_rays feeds the encoder GEMM by default);
 * `taps` = cpn_gather_rows_taps_bytes(rows) bytes of scratch (forms 1-3). Columns 832.. (tanh of the 3-D point, written by the sampling kernel) are left untouched. */
size_t cpn_gather_rows_taps_bytes(int rows);
int cpn_gather_rows(const cpn_render_args* a, int nr, const float* rowaux, void* out, int form, void* taps, void* stream);

/* Phase trace of the tensor-core GEMM (profiling only): while a buffer is set, every launch of the single-CTA kernel
 * writes 8 x uint64 per CTA for its first cap_ctas CTAs: globaltimer ns at CTA start, setup done, first operand stage
 * landed, last MMA issued, accumulators ready, epilogue done, CTA end, and the SM id. Pass NULL to switch it off. */
int cpn_gemm_tc_trace(unsigned long long* device_buf, int cap_ctas);

/* ---- building blocks exported for unit tests ------------------------------------------
 * C[M,N] = act(A[M,K] * W[N,K]^T + bias[N]); fp32 row-major, lda/ldc in floats.
 * `wt` is W transposed to [K,N] (fp32, as produced by cpn_pack_weights for SIMT layers). */
int cpn_gemm_simt(const float* A, int lda, const float* wt, const float* bias, float* C, int ldc,
                  int M, int N, int K, int relu /* 0 none, 1 ReLU, 2 exact GELU */, void* stream);

/* The same GEMM with the k range split over CTAs and reduced in a fixed order, for shapes whose 64 x 64 tiling would leave
 * most SMs idle (token layers of the cost aggregation: M = 256 / 1024 tokens, K up to 2304). act: 0 none, 1 ReLU, 2 GELU. */
size_t cpn_gemm_simt_splitk_workspace_bytes(int M, int N, int K);
int cpn_gemm_simt_splitk(const float* A, int lda, const float* wt, const float* bias, float* C, int ldc, int M, int N,
                         int K, int act, void* workspace, size_t workspace_bytes, void* stream);

/* Tensor-core GEMM of one packed layer (0 query_encode_latent, 1 query_encode_latent_2, 2 latent_value,
 * 3 key_map, 4 key_map_2, 5 query_embed_2, 6 query_repeat_embed_2, 7 latent_value o query_encode_latent_2,
 * 8 key_map o query_encode_latent_2, both with K = 1664, 9 [key_map_2 ; query_repeat_embed_2]^T query_embed_2 with
 * N = 256): C[M, N_layer] = act(A[M, K_layer] * W^T + b),
 * operands split into an fp16 head plus corrections (fp8 on the tensor cores' fp8 path by default: e5m2 activation planes, e4m3 weight planes; fp16 with CPN_TC_F16X3)
 * and accumulated in fp32 on tcgen05.
 * `packed` is the blob from cpn_pack_weights. mode bit CPN_TC_A_IMAGE: A is an "operand image" (128-row tiles,
 * per 32-wide k-chunk a 16 KB block [hi|lo][4][128][8] of fp16) instead of fp32 row-major; CPN_TC_OUT_IMAGE: C is
 * written as such an image with `out_kchunks` k-chunks per tile, source tile t landing in image tile t / out_div at
 * k offset (t % out_div) * N_layer (how the two branches of a sample row are concatenated). */
#define CPN_TC_A_IMAGE 1
#define CPN_TC_OUT_IMAGE 2
#define CPN_TC_F16X3 4   /* three fp16 MMAs per product; default is fp16 + two fp8 correction MMAs */
#define CPN_TC_PAIR 16   /* experiment: cta_group::2 CTA pairs, each SM stages half of every weight tile (f8 scheme,
                          * M % 512 == 0); correct, but measured 20-28 % slower than independent CTAs on B200 */
#define CPN_TC_OUT_CB16 32  /* fp32 output column-blocked: [row tile of 128][16-column block][row][16] (N = 128 layers) */
#define CPN_TC_OUT_ROWDOT 64 /* set by cpn_gemm_tc_rowdot */
#define CPN_TC_A_IMAGE3 1024  /* A is a COMPACT operand image: 12 KB blocks [fp16 head | remainder plane] without the value plane,
                              * which the persistent kernel derives in shared memory (e5m2 of the fp16 head) */
#define CPN_TC_OUT_IMAGE3 2048 /* the output image is written in that compact form (25 % fewer bytes) */
#define CPN_TC_PERSIST2 4096  /* experiment: the persistent kernel runs the two 128-row sub-tiles of a tile one after the other and
                              * drains one while the other accumulates (measured 1.39 vs 1.33 ms on query_encode_latent) */
#define CPN_TC_WS 8192        /* experiment: weight-stationary MMAs (tcgen05.mma.ws, the weight block is fetched once for both
                              * sub-tiles; N tiles of 64 / 128 / 256): same results, no measured gain */
#define CPN_TC_PPAIR 512     /* operand-image GEMMs, f8 scheme, M % 512 == 0: persistent cta_group::2 CTA pairs */
#define CPN_TC_NO_PERSIST 256 /* operand-image GEMMs: one tile per CTA (the first kernel) instead of the persistent kernel */
#define CPN_TC_OUT_KG 128    /* set by cpn_gemm_tc_kg */
#define CPN_TC_CLUSTER 8 /* experiment: the N-tile CTAs of a row tile form a cluster and multicast the A operand
                          * (halves L2 reads, but measured 8-18 % slower than independent CTAs on B200) */
int cpn_gemm_tc(const void* packed, int layer, const void* A, int lda, void* C, int ldc, int M, int relu, int mode,
                int out_div, int out_kchunks, void* stream);
/* The same GEMM for a 128-wide layer whose output is only needed dotted with another (M, 128) matrix, row by row:
 * out[m] = <act(A W^T + b)[m, :], dotv[m, :]> / div, with dotv in the CB16 layout (CPN_TC_OUT_CB16). Replaces
 * key_map_2 / query_repeat_embed_2 followed by the einsum('bijk,bijk->bjk') / 11.31 of models/CoPoNeRF.py:450,474. */
int cpn_gemm_tc_rowdot(const void* packed, int layer, const void* A, int lda, const float* dotv_cb16, float* out, int M,
                       int relu, int mode, float div, void* stream);

/* Layer 10 = [key_map ; G] o query_encode_latent_2 over the hidden-layer operand image (K = 1664, N = 256), M sample rows:
 *   logits[m] = (<relu(WKF h_m + bKF), dotv[m, :]> + rowadd[m]) / div      round-1 logit (models/CoPoNeRF.py:408,450)
 *   gh[m, :]  = G h_m + g0   (M rounded up to 128, 128) fp32 column-blocked [row tile of 128][8 blocks][128 rows][16]:
 *               the per-row term of the round-2 query bias. G folds latent_value,
 *               encode_latent and the z_embed columns of query_repeat_embed (models/CoPoNeRF.py:404,463-472, all linear),
 *               so that query_repeat_embed's per-ray bias is sum_rows w1[m] gh[m, :] and the round-1 readout is never formed.
 * dotv is CB16 with `dot_blocks` 16-column blocks per row tile, of which blocks 0..7 are used. */
int cpn_gemm_tc_kg(const void* packed, const void* h1_image, const float* dotv_cb16, int dot_blocks, const float* rowadd,
                   float div, float* logits, float* gh, int M, int mode, void* stream);

/* ---- generic Linear on the tcgen05 kernel: the token layers of the cost aggregation (q / k / v projections, MLPs,
 * proj_feat: models/aggregation.py:269-340,509-520) and any other nn.Linear with N % 128 == 0, K % 8 == 0.
 *   cpn_linear_tc_pack   W (N, K) fp32 as stored in the state_dict -> split, pre-scaled tensor-core tiles (once per weight)
 *   cpn_linear_tc        y[M, N] = act(x[M, K] W^T + bias), x / y fp32 row-major (ldx, ldy in floats), bias may be NULL;
 *                        act 0 none, 1 ReLU, 2 exact GELU; mode 0 = fp16 + two fp8 corrections, CPN_TC_F16X3 = three fp16 MMAs
 *                        (1e-5 of fp64 at K = 2304, against 4e-5 for the default scheme). */
size_t cpn_linear_tc_packed_bytes(int N, int K);
int cpn_linear_tc_pack(const float* w, int N, int K, void* packed, void* stream);
int cpn_linear_tc(const void* packed, int N, int K, const float* x, int ldx, const float* bias, float* y, int ldy, int M,
                  int act, int mode, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* COPONERF_B200_H */
