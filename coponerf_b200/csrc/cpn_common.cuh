// Internal declarations shared by the coponerf_b200 kernels.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include "../../include/coponerf_b200.h"

#define CPN_KA 848      // 835 encoder inputs padded to a multiple of 16 (zero columns), fp32 row form
#define CPN_KA_IMG 864  // ... padded to a multiple of 32, operand-image form (27 k-chunks)

// ---- packed fp32 weight blob (offsets in floats) -------------------------------------------
// Every matrix is stored transposed, [K][N], so consecutive threads read consecutive outputs.
namespace pw {
constexpr size_t W1T = 0;                        // [848][832]  query_encode_latent (rows >= 835 are zero)
constexpr size_t B1 = W1T + 848 * 832;           // [832]
constexpr size_t W2T = B1 + 832;                 // [832][416]  query_encode_latent_2
constexpr size_t B2 = W2T + 832 * 416;           // [416]
constexpr size_t WVT = B2 + 416;                 // [832][416]  latent_value
constexpr size_t BV = WVT + 832 * 416;           // [416]
constexpr size_t WKT = BV + 416;                 // [832][128]  key_map
constexpr size_t BK = WKT + 832 * 128;           // [128]
constexpr size_t WK2T = BK + 128;                // [128][128]  key_map_2
constexpr size_t BK2 = WK2T + 128 * 128;
constexpr size_t WQT = BK2 + 128;                // [16][128]   query_embed
constexpr size_t BQ = WQT + 16 * 128;
constexpr size_t WQ2T = BQ + 128;                // [128][128]  query_embed_2
constexpr size_t BQ2 = WQ2T + 128 * 128;
constexpr size_t WQRA_T = BQ2 + 128;             // [128][128]  query_repeat_embed, z_embed input channels
constexpr size_t WQRB_T = WQRA_T + 128 * 128;    // [16][128]   query_repeat_embed, local_coords input channels
constexpr size_t BQR = WQRB_T + 16 * 128;
constexpr size_t WQR2T = BQR + 128;              // [128][128]  query_repeat_embed_2
constexpr size_t BQR2 = WQR2T + 128 * 128;
constexpr size_t WET = BQR2 + 128;               // [416][128]  encode_latent
constexpr size_t BE = WET + 416 * 128;
constexpr size_t PHI_INT = BE + 128;             // [18][128]   phi.lin_in
constexpr size_t PHI_BIN = PHI_INT + 18 * 128;
constexpr size_t PHI_ZT = PHI_BIN + 128;         // 3 x [832][128] phi.lin_z.i
constexpr size_t PHI_BZ = PHI_ZT + 3 * 832 * 128;  // 3 x [128]
constexpr size_t PHI_F0T = PHI_BZ + 3 * 128;     // 3 x [128][128] phi.blocks.i.fc_0
constexpr size_t PHI_B0 = PHI_F0T + 3 * 128 * 128;
constexpr size_t PHI_F1T = PHI_B0 + 3 * 128;     // 3 x [128][128] phi.blocks.i.fc_1
constexpr size_t PHI_B1 = PHI_F1T + 3 * 128 * 128;
constexpr size_t PHI_OUT = PHI_B1 + 3 * 128;     // [3][128]    phi.lin_out (not transposed)
constexpr size_t PHI_BOUT = PHI_OUT + 3 * 128;   // [3] (+1 pad)
// query_encode_latent_2 folded into latent_value / key_map (no nonlinearity between them, CoPoNeRF.py:393-408):
// row-major (N, K = 1664) like the state_dict tensors; K index = branch * 832 + k of the 832-wide hidden layer
constexpr size_t WVF = PHI_BOUT + 4;             // [416][1664]
constexpr size_t BVF = WVF + 416 * 1664;         // [416]
constexpr size_t WKF = BVF + 416;                // [128][1664]
// Round-2 query bias as a linear function of the round-1 readout (no activation between latent_value, encode_latent and the
// z_embed half of query_repeat_embed, CoPoNeRF.py:463-473):  rbias = G hbar1 + g0 with G = Wqr[:, :128] We WVF (128 x 1664),
// g0 = Wqr[:, :128] (We bVF + be) + bqr.  The softmax weights of a ray sum to one, so  rbias = sum_rows w1 (G h + g0):
// G is stacked under WKF (one 256-row layer over the hidden image) and the 1664-wide round-1 readout is never formed.
constexpr size_t WG = WKF + 128 * 1664;          // [128][1664], directly after WKF: [WKF ; G] is one (256, 1664) matrix
constexpr size_t BKF = WG + 128 * 1664;          // [128]
constexpr size_t BG = BKF + 128;                 // [128] = g0, directly after BKF
// attention logits as bilinear forms (key_map_2, query_embed_2 and query_repeat_embed_2 have no activation,
// CoPoNeRF.py:408,446,473): <Wk2 k + bk2, Wq2 q + bq2> = k^T (WM q + BM) + (WS . q + CS), with
// WM = Wk2^T Wq2, BM = Wk2^T bq2, WS = Wq2^T bk2, CS = bk2 . bq2; the same with query_repeat_embed_2 for round 2. Both
// WM / BM pairs are stacked into one 256-row layer, so one GEMM over the coordinate embedding serves both rounds.
constexpr size_t WM12 = BG + 128;                // [256][128] row-major (N, K): rows 0-127 round 1, 128-255 round 2
constexpr size_t BM12 = WM12 + 256 * 128;        // [256]
constexpr size_t WS1 = BM12 + 256;               // [128] + CS1 at [128] (+3 pad)
constexpr size_t WS2 = WS1 + 132;
constexpr size_t M1D = (WS2 + 132 + 1) / 2 * 2;               // scratch of the G fold: Wqr[:, :128] We as (128, 416) doubles
constexpr size_t FP32_END = M1D + 2 * 128 * 416;
}  // namespace pw

// ---- per-pair constants (floats), written by cpn_pair_setup ---------------------------------
namespace pc {
constexpr int Q_C2W = 0;        // [2][16] query cam2world expressed in each context frame
constexpr int KQ = 32;          // [16]    query intrinsics
constexpr int KC = 48;          // [2][16] context intrinsics
constexpr int KN = 80;          // [2][9]  context intrinsics, rows 0-1 divided by H
constexpr int IDEN = 98;        // [2][16] inv(c2w_v) @ c2w_v
constexpr int T_OWN = 130;      // [2][16] sample point -> its own view's frame
constexpr int T_OTHER = 162;    // [2][16] sample point -> the other view's frame
constexpr int INV_QC2W = 194;   // [16]    inverse(query cam2world)
constexpr int INV_KQ3 = 210;    // [9]     inverse(query K[:3,:3])
constexpr int REL_FLIP = 219;   // [16] outputs rel_pose_flip, gt_rel_pose, gt_rel_pose_flip
constexpr int GT_REL = 235;
constexpr int GT_REL_FLIP = 251;
constexpr int END = 267;
static_assert(END <= CPN_PAIR_CONSTS_FLOATS, "pair consts overflow");
}  // namespace pc

// ---- per-row scratch written by the sample kernel ---------------------------------------------
// row = ((b * N_chunk + n) * 2 + v) * S + s : the 2S rows of a ray are contiguous.
#define CPN_ROWAUX 8  // grid_prim.xy, grid_sec.xy, clamp(pt).xyz, pad

// Row of the encoder-input matrix A for (sample row, branch): tiles of 128 sample rows, the 128 primary rows of
// a tile followed by its 128 secondary rows, so that one 128-row GEMM tile is one (tile, branch).
__host__ __device__ __forceinline__ size_t enc_row(size_t row, int branch) {
  return (row >> 7) * 256 + (size_t)branch * 128 + (row & 127);
}

// Byte offset of the 16-byte group holding k .. k+7 (k % 8 == 0) of row r (0..127) in image tile t of an
// operand image with `kchunks` 32-wide k-chunks per tile; the lo half sits 8192 bytes further.
__host__ __device__ __forceinline__ size_t act_img_off(size_t t, int kchunks, int k, int r) {
  return (t * kchunks + (k >> 5)) * 16384 + (size_t)((k & 31) >> 3) * 2048 + (size_t)r * 16;
}

void cpn_set_error(const char* fmt, ...);
#define CPN_CHECK_CUDA(expr)                                                              \
  do {                                                                                    \
    cudaError_t _e = (expr);                                                              \
    if (_e != cudaSuccess) {                                                              \
      cpn_set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
      return CPN_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)
#define CPN_CHECK_LAUNCH(name)                                                            \
  do {                                                                                    \
    cudaError_t _e = cudaGetLastError();                                                  \
    if (_e != cudaSuccess) {                                                              \
      cpn_set_error("launch of %s failed: %s", name, cudaGetErrorString(_e));            \
      return CPN_ERR_CUDA;                                                                \
    }                                                                                     \
  } while (0)

// launchers implemented across the .cu files (all asynchronous on `st`)
int launch_ray_setup(const cpn_render_args& a, int ray0, int nr, float* seg, cudaStream_t st);
// a_image: 0 fp32 rows of CPN_KA; 1 operand image (K = CPN_KA_IMG), f16x3 scheme; 2 operand image, f8 scheme
int launch_sample(const cpn_render_args& a, int ray0, int nr, const float* seg, float* rowaux, float* local16,
                  float* A, int a_image, cudaStream_t st);
// taps: (rows, 2 branches, 4 levels, 8) floats of scratch for the precomputed bilinear taps (operand-image forms only)
int launch_gather(const cpn_render_args& a, int ray0, int nr, const float* rowaux, float* A, int a_image, cudaStream_t st,
                  float* taps = nullptr);
// remap256: output row m goes to row (m / 256) * 128 + m % 128 at column offset ((m / 128) & 1) * N (undoes enc_row)
int launch_gemm_simt(const float* A, int lda, const float* wt, const float* bias, const float* rowbias,
                     int rows_per_bias, float* C, int ldc, int M, int N, int K, int relu, cudaStream_t st,
                     int remap256 = 0, float out_div = 0.f);   // out_div != 0: result divided by it
// 16 -> 128 ReLU layer (+ per-ray bias) written as the operand image (K = 128) of the tensor-core layer that follows
// Optional extras: sdot1 / sdot2 (132 floats each: a 128-vector and a constant) give s1[row] = <out, vec> + const
// (and s2); dotv (CB16, as written by the tensor-core GEMM) turns the kernel into a row-dot producer:
// lg[row] = (<out, dotv[row]> + rowadd[row]) / div and no image is written (img may be null).
int launch_mlp16_image(const float* x, const float* wt, const float* bias, const float* rowbias, int rows_per_bias, int M,
                       void* img, int f8, cudaStream_t st, const float* sdot1 = nullptr, float* s1 = nullptr,
                       const float* sdot2 = nullptr, float* s2 = nullptr, const float* dotv = nullptr,
                       const float* rowadd = nullptr, float div = 1.f, float* lg = nullptr, int dot_blocks = CPN_HIDDEN / 16,
                       int dot_block0 = 0);   // dotv rows are [row tile][dot_blocks][128][16]; this layer's columns start at dot_block0
// logits != nullptr: one precomputed logit per sample row (key / qemb unused); else <key, qemb> / 11.31 is computed here
int launch_attn1(const cpn_render_args& a, int ray0, int nr, const float* key, const float* qemb, const float* value,
                 const float* rowaux, float* r1, float* wp, cudaStream_t st, const float* logits = nullptr,
                 float* wts_out = nullptr,    // value == nullptr: no readout here, the weights go to wts_out
                 const float* gh = nullptr, float* rbias = nullptr);   // gh (R, 128): rbias[ray] = sum_rows w1 gh[row]
// z_all: (B, N, 416) latent of every ray of the image
int launch_attn2(const cpn_render_args& a, int ray0, int nr, const float* q2, const float* qemb, const float* value,
                 const float* r1, float* z_all, cudaStream_t st, const float* logits = nullptr, float* wts_out = nullptr,
                 const float* w1 = nullptr);   // w1: wts_out = w2 + 2 w1 (combined readout weight)
// late readout: hbar (rays, 1664) = sum over a ray's rows of w * [h_p ; h_s] read from the hidden-layer operand image
// out_N > 0: output rows are indexed by the ray's position in the whole image (b * out_N + out_ray0 + n) instead of the chunk
// chunk_bytes: 16384, or 12288 for a compact hidden image (CPN_TC_OUT_IMAGE3; the readout never reads the value plane)
int launch_readout_image(const cpn_render_args& a, int nr, const void* h1_image, const float* wts, float* hbar, int f8,
                         cudaStream_t st, int out_N = 0, int out_ray0 = 0, int chunk_bytes = 16384);
int launch_park_r1(const cpn_render_args& a, int ray0, int nr, const float* r1, float* z_all, cudaStream_t st);
int launch_finish_z(const cpn_render_args& a, const float* r2_all, float* z_all, cudaStream_t st);
int launch_finish_z_bias(const cpn_render_args& a, const float* r2_all, const float* bias, float* z_all, cudaStream_t st);
int launch_phi(const cpn_render_args& a, const float* z_all, cudaStream_t st);
int launch_ray_epilogue(const cpn_render_args& a, int ray0, int nr, const float* wp, const float* seg, cudaStream_t st);

// tensor-core path (gemm_tc.cu)
// "Operand image" of an activation matrix [rows x K]: tiles of 128 rows; per tile and 32-wide k-chunk one
// 16 KB block [hi | lo] x [4 groups of 8 k][128 rows][8 fp16] -- exactly what the MMA reads from shared memory
// (K-major, no swizzle), so a consumer stages it with one bulk copy. Block index = tile * (K / 32) + k-chunk.
// Two precision schemes share the block size: "f16x3" stores [hi fp16 8 KB | lo fp16 8 KB]; the default "f8"
// scheme stores [hi fp16 8 KB | e5m2(lo * 2^10) 4 KB | e5m2(hi) 4 KB] with the 8-bit planes as
// [2 groups of 16 k][128 rows][16 bytes] (tc_common.cuh).
constexpr int ACT_BK = 32;
constexpr int ACT_CHUNK_BYTES = 2 * (ACT_BK / 8) * 128 * 16;
constexpr int ACT_LO = 8192;      // f16x3: fp16 lo plane
constexpr int ACT_LO8 = 8192;     // f8: e5m2 remainder plane
constexpr int ACT_X8 = 12288;     // f8: e5m2 value plane
constexpr int CPN_TC_LAYERS = 11;
size_t cpn_packed_fp32_floats();
size_t cpn_tc_weights_bytes();
int cpn_pack_tc_weights(const float* raw, const float* packed_fp32, void* dst, cudaStream_t st);
// layer 9 fused with the 16 -> 128 ReLU layer in front of it (gemm_tc.cu)
int launch_gemm_tc_mlp16(const void* packed, const float* x16, const float* wt, const float* bias, const float* sd1, float* s1,
                         const float* sd2, float* s2, float* qm_cb16, int M, int mode, cudaStream_t st);
// generic Linear on the tensor-core kernel (include/coponerf_b200.h: cpn_linear_tc); the accumulators are scaled by out_mul
int launch_linear_tc(const void* packed, int N, int K, const float* x, int ldx, const float* bias, float* y, int ldy, int M,
                     int act, int mode, float out_mul, cudaStream_t stream);
// layer: 0 query_encode_latent, 1 query_encode_latent_2, 2 latent_value, 3 key_map, 4 key_map_2,
//        5 query_embed_2, 6 query_repeat_embed_2, 7 latent_value o query_encode_latent_2 (K = 1664: the hidden
//        layer of the primary branch then of the secondary one), 8 key_map o query_encode_latent_2.  mode: CPN_TC_A_IMAGE | CPN_TC_OUT_IMAGE.
//        | CPN_TC_F16X3 (three fp16 MMAs per product instead of fp16 + two fp8 corrections).
// CPN_TC_OUT_CB16: fp32 output as [row tile][16-col block][128][16]; CPN_TC_OUT_ROWDOT: C[row] = <out row, dotv row> / dot_div
int launch_gemm_tc(const void* packed, int layer, const void* A, int lda, void* C, int ldc, int M, int relu, int mode,
                   int out_div, int out_kchunks, cudaStream_t st, const float* dotv = nullptr, float dot_div = 1.f,
                   const float* dot_rowadd = nullptr, int dot_blocks = 0, int dot_block0 = 0, float* c2 = nullptr);   // row-dot output: C[row] = (<out row, dotv row> + dot_rowadd[row]) / dot_div
