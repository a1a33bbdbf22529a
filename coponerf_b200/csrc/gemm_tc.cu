// Tensor-core GEMM for the 1x1 convolutions of the per-sample encoder and attention MLPs
// (models/CoPoNeRF.py:387-408,446,473): C[M, N] = act(A[M, K] * W^T + b).
//
// fp32 parity on fp16 tensor cores: every operand is split x = hi + lo into two fp16 values (error 2^-22 |x|)
// and the product is accumulated in fp32 as  hi*hi + hi*lo + lo*hi  (three tcgen05.mma per k-step; the dropped
// lo*lo term is 2^-22 relative). Weights are pre-split, pre-scaled by a per-layer power of two (so the lo halves
// stay out of the fp16 subnormal range) and pre-tiled into the exact shared-memory image the MMA reads, so a
// plain bulk copy (TMA engine, one instruction per stage) moves them.
//
// The activation operand comes in one of two forms:
//   fp32 row-major      producer warps load it coalesced, split it and write the K-major core-matrix layout;
//   "operand image"     already split and tiled by the epilogue of the layer that produced it (ACT_* in
//                       cpn_common.cuh): the TMA warp bulk-copies 16 KB per 128-row tile and k-chunk, no
//                       conversion work at all.
// and the output is written either as fp32 row-major or as the operand image of the next layer.
//
// CTA = 256 rows (two UMMA M=128 sub-tiles sharing every weight stage: halves the L2 -> SM weight traffic per
// flop) x one NT-column tile; the two accumulators sit in TMEM columns [0, NT) and [256, 256 + NT). 10 warps:
//   warp 0   lane 0: bulk-copies weight tiles (and activation images) into the stage ring
//   warp 1   allocates TMEM; lane 0 issues the MMAs and commits stage-free / accumulator-ready barriers
//   warp 2-9 fp32-A mode: produce the A operand (32 rows per warp, two k-chunks of loads in flight), then
//            the epilogue: TMEM -> registers -> scale, bias, ReLU -> global (fp32 or hi/lo image)
#include <stdlib.h>
#include "cpn_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 256;            // rows per CTA: 2 sub-tiles of UMMA M = 128
constexpr int BK = ACT_BK;         // k per stage (32): 2 MMA k-steps of 16
constexpr int STAGES = 3;
constexpr int NT_MAX = 256;        // widest N tile (832 = 4 x 208, 416 = 2 x 208; 256 for layer 10: both accumulators fill TMEM)
constexpr int A_LBO = 128 * 16;                 // bytes between 8-wide k-chunks of a 128-row A sub-tile
constexpr int A_HALF = (BK / 8) * A_LBO;        // hi (or lo) half of one sub-tile stage = 8 KB
constexpr int A_SUB = 2 * A_HALF;               // one sub-tile stage = ACT_CHUNK_BYTES
constexpr int W_STAGE_MAX = 2 * (BK / 8) * NT_MAX * 16;
constexpr int STAGE_BYTES = 2 * A_SUB + W_STAGE_MAX;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;
constexpr int TMEM_COLS = 512;
constexpr int NUM_THREADS = 320;
static_assert(A_SUB == ACT_CHUNK_BYTES, "operand image chunk must equal one A sub-tile stage");

struct TcLayer {
  int out, in;     // weight shape
  int kpad;        // K padded to a multiple of BK (zero weights beyond `in`)
  int nt;          // N tile
  size_t raw;      // offset of the weight in the raw state_dict blob (floats); folded layers: in the packed fp32 section
  size_t bias;     // offset of the fp32 bias in the packed fp32 section
  bool folded;     // weight computed by cpn_pack_weights (pw::WVF / pw::WKF) instead of read from the state_dict
};
// raw blob offsets (floats), see weights.cu kTensors
constexpr size_t RAW_W1 = 0;
constexpr size_t RAW_W2 = RAW_W1 + 832 * 835 + 832;
constexpr size_t RAW_WV = RAW_W2 + 416 * 832 + 416;
constexpr size_t RAW_WK = RAW_WV + 416 * 832 + 416;
constexpr size_t RAW_WK2 = RAW_WK + 128 * 832 + 128;
constexpr size_t RAW_WQ = RAW_WK2 + 128 * 128 + 128;
constexpr size_t RAW_WQ2 = RAW_WQ + 128 * 16 + 128;
constexpr size_t RAW_WQR = RAW_WQ2 + 128 * 128 + 128;
constexpr size_t RAW_WQR2 = RAW_WQR + 128 * 144 + 128;
const TcLayer kLayers[CPN_TC_LAYERS] = {
    {832, 835, 864, 208, RAW_W1, pw::B1, false},      // 0 query_encode_latent
    {416, 832, 832, 208, RAW_W2, pw::B2, false},      // 1 query_encode_latent_2
    {416, 832, 832, 208, RAW_WV, pw::BV, false},      // 2 latent_value
    {128, 832, 832, 128, RAW_WK, pw::BK, false},      // 3 key_map
    {128, 128, 128, 128, RAW_WK2, pw::BK2, false},    // 4 key_map_2
    {128, 128, 128, 128, RAW_WQ2, pw::BQ2, false},    // 5 query_embed_2
    {128, 128, 128, 128, RAW_WQR2, pw::BQR2, false},  // 6 query_repeat_embed_2
    {416, 1664, 1664, 208, pw::WVF, pw::BVF, true},   // 7 latent_value o query_encode_latent_2 (both branches)
    {128, 1664, 1664, 128, pw::WKF, pw::BKF, true},   // 8 key_map o query_encode_latent_2
    {256, 128, 128, 256, pw::WM12, pw::BM12, true},   // 9 [key_map_2 ; query_repeat_embed_2]^T query_embed_2 (bilinear logits), one 256-column tile
    {256, 1664, 1664, 256, pw::WKF, pw::BKF, true},   // 10 [key_map ; G] o query_encode_latent_2: key hidden layer and the
                                                      //    per-row term of the round-2 query bias (cpn_common.cuh, pw::WG)
};
constexpr size_t TC_HEADER_BYTES = 256;   // floats [0..15] 1/scale per layer, [16..31] scale, uints [32..47] absmax bits
static_assert(CPN_TC_LAYERS <= 16, "header slots");

size_t layer_bytes(int l) { return (size_t)kLayers[l].out * kLayers[l].kpad * 4; }  // 4 bytes per weight either scheme
size_t scheme_bytes() {
  size_t n = 0;
  for (int i = 0; i < CPN_TC_LAYERS; ++i) n += layer_bytes(i);
  return n;
}
// scheme 0: f16x3 tiles [w_hi | w_lo] fp16; scheme 1: f8 tiles [w_hi fp16 | e4m3(w_hi 2^-10) | e4m3(w_lo)];
// scheme 2: the f8 planes again in half tiles of NT / 2 rows (one per CTA of a cta_group::2 pair)
size_t layer_offset(int l, int scheme) {
  size_t off = TC_HEADER_BYTES + (size_t)scheme * scheme_bytes();
  for (int i = 0; i < l; ++i) off += layer_bytes(i);
  return off;
}

// ---------------------------------------------------------------------------------------------- packing
__global__ void absmax_kernel(const float* __restrict__ w, size_t n, unsigned int* __restrict__ out) {
  float m = 0.f;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    m = fmaxf(m, fabsf(w[i]));
  for (int o = 16; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  if ((threadIdx.x & 31) == 0) atomicMax(out, __float_as_uint(m));
}

// power-of-two scale that brings max|w| into [2^14, 2^15) (tc_common.cuh: TC_WEIGHT_LOG2)
__device__ __forceinline__ float layer_scale(unsigned int absmax_bits) {
  float m = __uint_as_float(absmax_bits);
  if (!(m > 0.f) || isinf(m)) return 1.f;
  int e;
  frexpf(m, &e);              // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, TC_WEIGHT_LOG2 - e);
}

// dst tile (nt, kc), 128 * NT bytes: f16x3 [hi | lo] x [4 k-groups][NT rows][8 halves];
// f8 [hi as before | e4m3(hi 2^-10) | e4m3(lo)] with the byte planes as [2 k-groups of 16][NT rows][16 bytes]
__global__ void pack_tc_kernel(const float* __restrict__ w, int out, int in, int kpad, int NT, const unsigned int* absmax,
                               __half* __restrict__ dst, unsigned char* __restrict__ dst8,
                               unsigned char* __restrict__ dstp, float* __restrict__ header, int layer) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)out * kpad;
  float scale = layer_scale(*absmax);
  if (i == 0) {
    header[layer] = 1.f / scale;
    header[16 + layer] = scale;
  }
  if (i >= total) return;
  int k = (int)(i % kpad), n = (int)(i / kpad);
  int nt = n / NT, nl = n % NT, kc = k / BK, c = (k % BK) / 8, e = k % 8;
  int kchunks = kpad / BK;
  float x = (k < in) ? w[(size_t)n * in + k] * scale : 0.f;
  __half hi = __float2half_rn(x);
  __half lo = __float2half_rn(x - __half2float(hi));
  size_t half_elems = (size_t)(BK / 8) * NT * 8;
  size_t tile = ((size_t)nt * kchunks + kc) * 2 * half_elems;
  size_t off = ((size_t)c * NT + nl) * 8 + e;
  dst[tile + off] = hi;
  dst[tile + half_elems + off] = lo;
  // f8 scheme tile
  unsigned char* t8 = dst8 + tile * 2;             // same tile size in bytes
  reinterpret_cast<__half*>(t8)[off] = hi;
  const float lo_exact = x - __half2float(hi);
  size_t off8 = ((size_t)((k % BK) / 16) * NT + nl) * 16 + (k % 16);
  const unsigned char w8 = __nv_cvt_float_to_fp8(__half2float(hi) * F8_W_SCALE, __NV_SATFINITE, __NV_E4M3);
  const unsigned char wl8 = __nv_cvt_float_to_fp8(lo_exact, __NV_SATFINITE, __NV_E4M3);
  t8[half_elems * 2 + off8] = w8;
  t8[half_elems * 3 + off8] = wl8;
  // pair tiles: (nt, half, kc) with NH = NT / 2 rows: [hi: 4 groups x NH x 16 B | w8: 2 x NH x 16 | w_lo8: 2 x NH x 16]
  if (!dstp) return;
  const int NH = NT / 2, hf = nl / NH, nh = nl % NH;
  unsigned char* tp = dstp + (((size_t)nt * 2 + hf) * kchunks + kc) * (size_t)NH * 128;
  reinterpret_cast<__half*>(tp)[((size_t)c * NH + nh) * 8 + e] = hi;
  const size_t offp = ((size_t)((k % BK) / 16) * NH + nh) * 16 + (k % 16);
  tp[(size_t)NH * 64 + offp] = w8;
  tp[(size_t)NH * 96 + offp] = wl8;
}

// ---------------------------------------------------------------------------------------------- the GEMM
struct GemmArgs {
  const void* A;                 // fp32 row-major (lda) or operand image
  int lda, kreal, M;
  void* C;                       // fp32 row-major (ldc) or operand image
  int ldc, N, relu;
  int out_div;                   // image output: source tile t lands in image tile t / out_div at k offset (t % out_div) * N
  int out_kchunks;               // k-chunks per tile of the output image
  const unsigned char* wtiles;   // this layer's tiles
  const float* bias;
  const float* inv_scale;        // header[layer]
  int kchunks, NT;
  uint32_t idesc;
  int f8;                        // 1: fp16 + two fp8 correction MMAs (e5m2 activations x e4m3 weights), 0: three fp16 MMAs
  int out_kind;                  // fp32 output: 0 row-major, 2 column-blocked (CB16), 3 per-row dot with `dotv`,
                                 // 4 N tile 0 as kind 3 (with ReLU), the other N tiles row-major into C2 without ReLU
  float* C2;                     // out_kind 4: fp32 (M, N - NT) row-major
  int ws;                        // persistent kernel: weight-stationary MMAs (NT = 64 / 128 / 256, fp16 + fp8 scheme)
  int w3;                        // persistent kernel, compact A: the e4m3(w_hi 2^-10) weight plane is derived on chip too
  unsigned long long* dbg;       // phase timestamps of the first `dbg_cap` CTAs (cpn_gemm_tc_trace), else null
  int dbg_cap;
  // MLP16 producer (gemm_tc_kernel<false, false, 1, true>): A = relu(x16 Wt + b) computed by the producer warps
  const float *mlp_x, *mlp_wt, *mlp_b, *mlp_sd1, *mlp_sd2;   // x (M, 16); Wt [16][128]; b [128]; sd1 / sd2: [128] vector + constant
  float *mlp_s1, *mlp_s2;                                    // s1[row] = <A row, sd1> + sd1[128] (and s2), nullable
  int a_chunk;                   // bytes per (tile, k-chunk) block of the A image: 16384, or 12288 for the compact form without the
                                 // value plane (the persistent kernel derives it in shared memory)
  int out_chunk;                 // the same for an image output
  float out_mul;                 // the accumulators are multiplied by inv_scale * out_mul before the bias (1 except the tail GEMM)
  int dbg_skip;                  // trace runs only (CPN_TC_DBG_SKIP): 1 no image stores, 2 no split / conversions, 4 no TMEM loads,
                                 // persistent kernel: 8 no MMAs, 16 no activation copies, 32 no weight copies
  const float* dotv;             // CB16 matrix the rows are dotted with (out_kind 3); C then holds one float per row
  float dot_div;
  const float* dot_rowadd;       // optional per-row term added to the dot product before the division
  int dot_blocks, dot_block0;    // dotv is [row tile][dot_blocks 16-column blocks][128][16]; this layer starts at dot_block0
};

// Drain one 128-row accumulator sub-tile: TMEM -> registers -> scale, bias, ReLU -> fp32 rows or the operand image
// of the next layer. Warp quadrant q owns TMEM lanes 32 q .. 32 q + 31.
// Columns [c_lo, c_hi) of the tile (multiples of 16); returns this thread's partial row-dot (kinds 3 / 4) and, with
// finish_dot, also writes the finished logit (callers that split the columns over several warps combine the partials).
template <bool OUT_IMAGE>
__device__ __forceinline__ float drain_subtile(const GemmArgs& g, uint32_t tmem, int m0, int n_tile, int esub, int q,
                                               int lane, int c_lo, int c_hi, bool finish_dot) {
  const int NT = g.NT;
  const int rloc = q * 32 + lane;                  // row inside the 128-row sub-tile
  const int row = m0 + esub * 128 + rloc;
  const float inv = *g.inv_scale * g.out_mul;
  const int n0 = n_tile * NT;
  // out_kind 4 (layer 10, one 256-column tile): columns 0-127 (key hidden layer) are dotted like kind 3, columns 128-255
  // (G h + g0) leave column-blocked (8 blocks of 16 per 128-row tile)
  const bool kg = g.out_kind == 4;
  float* const crow = reinterpret_cast<float*>(g.C);
  const int ldc = g.ldc, ccol0 = n0;
  const uint32_t tsrc = tmem + esub * 256 + ((uint32_t)(q * 32) << 16);
  unsigned char* img = nullptr;   // this thread's row inside the output image tile
  int kbase = 0;                  // k of the next layer that column n0 of this tile maps to
  if (OUT_IMAGE) {
    int t = m0 / 128 + esub;
    img = reinterpret_cast<unsigned char*>(g.C) + (size_t)(t / g.out_div) * g.out_kchunks * g.out_chunk + rloc * 16;
    kbase = (t % g.out_div) * g.N + n0;
  }
  // CB16: fp32 [row tile][16-column block][128 rows][16], so that a thread's 16 columns are 64 contiguous bytes and the
  // 32 rows of a warp 2 KB: coalesced for this epilogue and for a consumer that owns one row per thread
  const size_t cb_row = ((size_t)(m0 / 128 + esub) * (g.N / 16)) * 128 * 16 + (size_t)rloc * 16;
  float dot = 0.f;
  for (int c0 = c_lo; c0 < c_hi; c0 += 16) {
    const int kind = kg ? (c0 < CPN_HIDDEN ? 3 : 0) : g.out_kind;
    const int act = kg ? (c0 < CPN_HIDDEN ? 1 : 0) : g.relu;   // 0 none, 1 ReLU, 2 exact GELU
    float v[16];
    if (g.dbg_skip & 4) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = (float)(c0 + j);
    } else {
      tmem_ld16(tsrc + c0, v);
    }
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      float4 b = *reinterpret_cast<const float4*>(g.bias + n0 + c0 + j);
      v[j] = v[j] * inv + b.x;
      v[j + 1] = v[j + 1] * inv + b.y;
      v[j + 2] = v[j + 2] * inv + b.z;
      v[j + 3] = v[j + 3] * inv + b.w;
    }
    if (act == 1) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = fmaxf(v[j], 0.f);
    } else if (act == 2) {
#pragma unroll
      for (int j = 0; j < 16; ++j) v[j] = 0.5f * v[j] * (1.f + erff(v[j] * 0.70710678118654752440f));
    }
    if (OUT_IMAGE) {
      // 16 consecutive k of the next layer; lanes are consecutive rows -> every store instruction writes 512 B runs
      const int k = kbase + c0;
      unsigned char* chunk = img + (size_t)(k / BK) * g.out_chunk;
      if (g.f8) {
        uint2 h[4];
        uint4 l8, x8;
        if (g.dbg_skip & 2) {   // trace runs: raw bits instead of the split
          h[0] = make_uint2(__float_as_uint(v[0]), __float_as_uint(v[1]));
          h[1] = make_uint2(__float_as_uint(v[2]), __float_as_uint(v[3]));
          h[2] = make_uint2(__float_as_uint(v[4]), __float_as_uint(v[5]));
          h[3] = make_uint2(__float_as_uint(v[6]), __float_as_uint(v[7]));
          l8 = make_uint4(__float_as_uint(v[8]), __float_as_uint(v[9]), __float_as_uint(v[10]), __float_as_uint(v[11]));
          x8 = make_uint4(__float_as_uint(v[12]), __float_as_uint(v[13]), __float_as_uint(v[14]), __float_as_uint(v[15]));
        } else {
          split4_f8(make_float4(v[0], v[1], v[2], v[3]), h[0], l8.x, x8.x);
          split4_f8(make_float4(v[4], v[5], v[6], v[7]), h[1], l8.y, x8.y);
          split4_f8(make_float4(v[8], v[9], v[10], v[11]), h[2], l8.z, x8.z);
          split4_f8(make_float4(v[12], v[13], v[14], v[15]), h[3], l8.w, x8.w);
        }
        unsigned char* ph = chunk + ((k % BK) / 8) * A_LBO;
        if (!(g.dbg_skip & 1) || (h[0].x == 0x12345678u && l8.y == 0x9abcdef0u)) {
          *reinterpret_cast<uint4*>(ph) = make_uint4(h[0].x, h[0].y, h[1].x, h[1].y);
          *reinterpret_cast<uint4*>(ph + A_LBO) = make_uint4(h[2].x, h[2].y, h[3].x, h[3].y);
          *reinterpret_cast<uint4*>(chunk + ACT_LO8 + ((k % BK) / 16) * A_LBO) = l8;
          if (g.out_chunk == ACT_CHUNK_BYTES) *reinterpret_cast<uint4*>(chunk + ACT_X8 + ((k % BK) / 16) * A_LBO) = x8;
        }
      } else {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          unsigned char* p = chunk + (((k + h * 8) % BK) / 8) * A_LBO;
          uint4 hi, lo;
          split8(make_float4(v[h * 8], v[h * 8 + 1], v[h * 8 + 2], v[h * 8 + 3]),
                 make_float4(v[h * 8 + 4], v[h * 8 + 5], v[h * 8 + 6], v[h * 8 + 7]), hi, lo);
          *reinterpret_cast<uint4*>(p) = hi;
          *reinterpret_cast<uint4*>(p + A_HALF) = lo;
        }
      }
    } else if (kind == 2) {
      float* out = reinterpret_cast<float*>(g.C) + cb_row + (size_t)((n0 + c0) / 16) * 128 * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else if (kind == 3) {
      const float4* qv = reinterpret_cast<const float4*>(
          g.dotv + (((size_t)(m0 / 128 + esub) * g.dot_blocks + g.dot_block0 + (n0 + c0) / 16) * 128 + rloc) * 16);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float4 x = __ldg(qv + j);
        dot = fmaf(v[4 * j + 3], x.w, fmaf(v[4 * j + 2], x.z, fmaf(v[4 * j + 1], x.y, fmaf(v[4 * j], x.x, dot))));
      }
    } else if (kg) {
      // G h + g0 leaves column-blocked like CB16 with 8 blocks per row tile: 64 contiguous bytes per thread, 2 KB per warp
      float* out = g.C2 + (((size_t)(m0 / 128 + esub) * (CPN_HIDDEN / 16) + (c0 - CPN_HIDDEN) / 16) * 128 + rloc) * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    } else if (row < g.M) {
      float* out = crow + (size_t)row * ldc + ccol0 + c0;
#pragma unroll
      for (int j = 0; j < 16; j += 4) *reinterpret_cast<float4*>(out + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
    }
  }
  if (finish_dot && !OUT_IMAGE && (kg || g.out_kind == 3) && row < g.M)
    reinterpret_cast<float*>(g.C)[row] = (g.dot_rowadd ? dot + g.dot_rowadd[row] : dot) / g.dot_div;
  return dot;
}

// CLUSTER (> 1, operand-image A only): the CTAs of the N tiles of one 256-row tile form a cluster; each loads
// 1/CLUSTER of every A stage and multicasts it to all of them, so the shared A operand is read from L2 once per
// cluster instead of once per N tile (the hypothesis then was an L2 -> SM bandwidth bound; round 2 measured otherwise).
// MLP16: the A operand is relu(x16 Wt + b) (query_embed, 16 -> 128, CoPoNeRF.py:446) computed by the producer warps from the
// 64-byte local_coords rows, so the coordinate embedding never exists in memory (it used to be an operand image written by one
// kernel and read back by this one: 2 x 134 MB per chunk and a launch); the scalar logit terms <q1, WS> + CS come out of it too.
template <bool A_IMAGE, bool OUT_IMAGE, int CLUSTER, bool MLP16 = false>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(GemmArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(16) float mlp_ws[MLP16 ? 16 : 1][CPN_HIDDEN];
  __shared__ __align__(16) float mlp_v[MLP16 ? 3 : 1][CPN_HIDDEN];   // bias, sd1, sd2
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 1];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m0 = blockIdx.y * BM;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full_a = smem_u32(&bars[0]), full_w = smem_u32(&bars[STAGES]), empty = smem_u32(&bars[2 * STAGES]),
                 accum = smem_u32(&bars[3 * STAGES]);
  const int NT = g.NT;
  const uint32_t w_half = (uint32_t)(BK / 8) * NT * 16;      // bytes of the hi (or lo) part of a weight tile
  const bool sub1_valid = (m0 + 128) < g.M;                  // does the second 128-row sub-tile hold any row?
  // optional phase trace: [0] CTA start, [1] setup done, [2] first stage landed, [3] last MMA issued, [4] accumulators
  // ready (seen by an epilogue warp), [5] epilogue warp done, [6] CTA end, [7] smid
  const int cta_lin = blockIdx.y * gridDim.x + blockIdx.x;
  unsigned long long* const dbg = (g.dbg && cta_lin < g.dbg_cap) ? g.dbg + (size_t)cta_lin * 8 : nullptr;
  auto stamp = [&](int i) {
    if (dbg) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[i] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_a + 8 * s, 256);
      mbar_init(full_w + 8 * s, 1);
      mbar_init(empty + 8 * s, CLUSTER);   // one commit from the MMA warp of every CTA that reads this stage
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
  if (MLP16) {
    for (int i = threadIdx.x; i < 16 * CPN_HIDDEN; i += NUM_THREADS) mlp_ws[i / CPN_HIDDEN][i % CPN_HIDDEN] = g.mlp_wt[i];
    for (int i = threadIdx.x; i < CPN_HIDDEN; i += NUM_THREADS) {
      mlp_v[0][i] = g.mlp_b[i];
      mlp_v[1][i] = g.mlp_sd1 ? g.mlp_sd1[i] : 0.f;
      mlp_v[2][i] = g.mlp_sd2 ? g.mlp_sd2[i] : 0.f;
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync();   // every CTA's barriers are initialised before any remote arrive / copy
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    stamp(1);
    if (dbg) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      dbg[7] = smid;
    }
  }
  const uint32_t crank = CLUSTER > 1 ? cluster_ctarank() : 0;
  constexpr uint16_t cmask = (uint16_t)((1u << CLUSTER) - 1);

  if (warp == 0) {
    if (lane == 0) {
      const unsigned char* wsrc = g.wtiles + (size_t)n_tile * g.kchunks * 2 * w_half;
      const unsigned char* asrc = reinterpret_cast<const unsigned char*>(g.A);
      const size_t tile0 = (size_t)(m0 / 128);
      for (int i = 0; i < g.kchunks; ++i) {
        int s = i % STAGES;
        uint32_t u = i / STAGES;
        mbar_wait(empty + 8 * s, (u & 1) ^ 1);
        uint32_t stage = smem0 + s * STAGE_BYTES;
        uint32_t bytes = 2 * w_half + (A_IMAGE ? (sub1_valid ? 2 : 1) * A_SUB : 0);
        mbar_arrive_expect_tx(full_w + 8 * s, bytes);
        bulk_g2s(stage + 2 * A_SUB, wsrc + (size_t)i * 2 * w_half, 2 * w_half, full_w + 8 * s);
        if (A_IMAGE && CLUSTER > 1) {
          // this CTA's slice of the A stage (both sub-tiles are contiguous 16 KB blocks), multicast to the cluster
          constexpr uint32_t SLICE = 2 * A_SUB / CLUSTER;
          const uint32_t off = crank * SLICE, sub = off / A_SUB, in_sub = off % A_SUB;
          if (sub == 0 || sub1_valid)
            bulk_g2s_multicast(stage + off, asrc + ((tile0 + sub) * g.kchunks + i) * ACT_CHUNK_BYTES + in_sub, SLICE,
                               full_w + 8 * s, cmask);
        } else if (A_IMAGE) {
          bulk_g2s(stage, asrc + (tile0 * g.kchunks + i) * ACT_CHUNK_BYTES, A_SUB, full_w + 8 * s);
          if (sub1_valid)
            bulk_g2s(stage + A_SUB, asrc + ((tile0 + 1) * g.kchunks + i) * ACT_CHUNK_BYTES, A_SUB, full_w + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < g.kchunks; ++i) {
        int s = i % STAGES;
        uint32_t u = i / STAGES;
        if (!A_IMAGE) mbar_wait(full_a + 8 * s, u & 1);
        mbar_wait(full_w + 8 * s, u & 1);
        if (i == 0) stamp(2);
        tcgen05_fence_after();
        uint32_t stage = smem0 + s * STAGE_BYTES;
        uint32_t b_hi = stage + 2 * A_SUB, b_lo = b_hi + w_half;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (sub == 1 && !sub1_valid) break;
          uint32_t a_hi = stage + sub * A_SUB, a_lo = a_hi + A_HALF;
          uint32_t d = tmem + sub * 256;
          if (g.f8) {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j)
              mma_f16_ss(d, make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128), make_desc(b_hi + j * 2 * NT * 16, NT * 16, 128),
                         g.idesc, (i | j) != 0);
            // 8-bit planes: K = 32 per instruction = two 16-byte core matrices along k
            mma_f8_ss(d, make_desc(a_hi + ACT_LO8, A_LBO, 128), make_desc(b_hi + w_half, NT * 16, 128), g.idesc | IDESC_A_E5M2, 1);
            mma_f8_ss(d, make_desc(a_hi + ACT_X8, A_LBO, 128), make_desc(b_hi + w_half + w_half / 2, NT * 16, 128),
                      g.idesc | IDESC_A_E5M2, 1);
          } else {
#pragma unroll
            for (int j = 0; j < BK / 16; ++j) {
              uint64_t da_hi = make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128);
              uint64_t da_lo = make_desc(a_lo + j * 2 * A_LBO, A_LBO, 128);
              uint64_t db_hi = make_desc(b_hi + j * 2 * NT * 16, NT * 16, 128);
              uint64_t db_lo = make_desc(b_lo + j * 2 * NT * 16, NT * 16, 128);
              mma_f16_ss(d, da_hi, db_hi, g.idesc, (i | j) != 0);
              mma_f16_ss(d, da_hi, db_lo, g.idesc, 1);
              mma_f16_ss(d, da_lo, db_hi, g.idesc, 1);
            }
          }
        }
        // the stage is free once these MMAs have read it (in every CTA of the cluster, for the shared A slices)
        if (CLUSTER > 1) mma_commit_multicast(empty + 8 * s, cmask); else mma_commit(empty + 8 * s);
      }
      mma_commit(accum);
      stamp(3);
    }
  } else {
    const int pwarp = warp - 2;               // 0..7: rows [32 * pwarp, 32 * pwarp + 32) of the CTA tile
    if (!A_IMAGE) {
      // ---- A producer. lane -> (k-chunk c = lane / 8, row r_in = lane % 8): 8 consecutive lanes store 128
      // contiguous bytes. Loads of chunks i + 1 and i + 2 are in flight while chunk i is converted.
      const float* Ag = reinterpret_cast<const float*>(g.A);
      const int r_in = lane & 7, c = lane >> 3;
      const int sub = pwarp >> 2, rbase = (pwarp & 3) * 32;
      float4 buf[3][4][2];
      auto load_chunk = [&](int i, float4 (&v)[4][2]) {
        const int k = i * BK + c * 8;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          int row = m0 + pwarp * 32 + it * 8 + r_in;
          if (row < g.M && k < g.kreal) {
            const float4* p = reinterpret_cast<const float4*>(Ag + (size_t)row * g.lda + k);
            v[it][0] = __ldg(p);
            v[it][1] = __ldg(p + 1);
          } else {
            v[it][0] = v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
          }
        }
      };
      auto store_chunk = [&](int i, float4 (&v)[4][2]) {
        int s = i % STAGES;
        uint32_t u = i / STAGES;
        mbar_wait(empty + 8 * s, (u & 1) ^ 1);
        unsigned char* a_hi = smem + s * STAGE_BYTES + sub * A_SUB;
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          int r = rbase + it * 8 + r_in;
          if (g.f8) {   // this thread's 8 k: one 16-byte fp16 group, half of a 16-byte 8-bit group in each byte plane
            uint2 h0, h1, l8, x8;
            split4_f8(v[it][0], h0, l8.x, x8.x);
            split4_f8(v[it][1], h1, l8.y, x8.y);
            *reinterpret_cast<uint4*>(a_hi + c * A_LBO + r * 16) = make_uint4(h0.x, h0.y, h1.x, h1.y);
            *reinterpret_cast<uint2*>(a_hi + ACT_LO8 + (c >> 1) * A_LBO + r * 16 + (c & 1) * 8) = l8;
            *reinterpret_cast<uint2*>(a_hi + ACT_X8 + (c >> 1) * A_LBO + r * 16 + (c & 1) * 8) = x8;
          } else {
            uint4 hi, lo;
            split8(v[it][0], v[it][1], hi, lo);
            *reinterpret_cast<uint4*>(a_hi + c * A_LBO + r * 16) = hi;
            *reinterpret_cast<uint4*>(a_hi + A_HALF + c * A_LBO + r * 16) = lo;
          }
        }
        fence_proxy_async_smem();
        mbar_arrive(full_a + 8 * s);
      };
      if (MLP16) {
        // this thread: rows pwarp * 32 + it * 8 + r_in (it = 0..3), outputs k = 32 i + 8 c .. + 7 of every k-chunk i
        float xin[4][16], d1[4] = {0.f, 0.f, 0.f, 0.f}, d2[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          const int row = m0 + pwarp * 32 + it * 8 + r_in;
#pragma unroll
          for (int j = 0; j < 16; j += 4) {
            const float4 v = row < g.M ? __ldg(reinterpret_cast<const float4*>(g.mlp_x + (size_t)row * 16 + j))
                                       : make_float4(0.f, 0.f, 0.f, 0.f);
            xin[it][j] = v.x; xin[it][j + 1] = v.y; xin[it][j + 2] = v.z; xin[it][j + 3] = v.w;
          }
        }
        for (int i = 0; i < g.kchunks; ++i) {
          const int k0 = i * BK + c * 8;
          float acc[4][8];
#pragma unroll
          for (int it = 0; it < 4; ++it)
#pragma unroll
            for (int e = 0; e < 8; ++e) acc[it][e] = mlp_v[0][k0 + e];
#pragma unroll
          for (int j = 0; j < 16; ++j) {
            const float4 w0 = *reinterpret_cast<const float4*>(&mlp_ws[j][k0]), w1 = *reinterpret_cast<const float4*>(&mlp_ws[j][k0 + 4]);
            const float wv[8] = {w0.x, w0.y, w0.z, w0.w, w1.x, w1.y, w1.z, w1.w};
#pragma unroll
            for (int it = 0; it < 4; ++it)
#pragma unroll
              for (int e = 0; e < 8; ++e) acc[it][e] = fmaf(xin[it][j], wv[e], acc[it][e]);
          }
#pragma unroll
          for (int it = 0; it < 4; ++it) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              acc[it][e] = fmaxf(acc[it][e], 0.f);
              d1[it] = fmaf(acc[it][e], mlp_v[1][k0 + e], d1[it]);
              d2[it] = fmaf(acc[it][e], mlp_v[2][k0 + e], d2[it]);
            }
            buf[0][it][0] = make_float4(acc[it][0], acc[it][1], acc[it][2], acc[it][3]);
            buf[0][it][1] = make_float4(acc[it][4], acc[it][5], acc[it][6], acc[it][7]);
          }
          store_chunk(i, buf[0]);
        }
        // the four lanes that share a row (k-groups c = 0..3) add their partial dot products in a fixed order
#pragma unroll
        for (int it = 0; it < 4; ++it) {
          d1[it] += __shfl_xor_sync(0xffffffffu, d1[it], 8);
          d1[it] += __shfl_xor_sync(0xffffffffu, d1[it], 16);
          d2[it] += __shfl_xor_sync(0xffffffffu, d2[it], 8);
          d2[it] += __shfl_xor_sync(0xffffffffu, d2[it], 16);
          const int row = m0 + pwarp * 32 + it * 8 + r_in;
          if (c == 0 && row < g.M) {
            if (g.mlp_s1) g.mlp_s1[row] = d1[it] + g.mlp_sd1[CPN_HIDDEN];
            if (g.mlp_s2) g.mlp_s2[row] = d2[it] + g.mlp_sd2[CPN_HIDDEN];
          }
        }
      } else {
      if (0 < g.kchunks) load_chunk(0, buf[0]);
      if (1 < g.kchunks) load_chunk(1, buf[1]);
      int i = 0;
      for (; i + 2 < g.kchunks; i += 3) {   // rotate three register buffers without copies
        load_chunk(i + 2, buf[2]);
        store_chunk(i, buf[0]);
        if (i + 3 < g.kchunks) load_chunk(i + 3, buf[0]);
        store_chunk(i + 1, buf[1]);
        if (i + 4 < g.kchunks) load_chunk(i + 4, buf[1]);
        store_chunk(i + 2, buf[2]);
      }
      if (i < g.kchunks) store_chunk(i, buf[0]);
      if (i + 1 < g.kchunks) store_chunk(i + 1, buf[1]);
      }
    }
    // ---- epilogue: warp w may touch TMEM lanes 32 * (w % 4) .. + 31; warps 2-5 drain sub-tile 0, 6-9 sub-tile 1
    const int esub = pwarp >> 2, q = warp & 3;
    if (esub == 0 || sub1_valid) {
      mbar_wait(accum, 0);
      if (threadIdx.x == 64) stamp(4);
      tcgen05_fence_after();
      drain_subtile<OUT_IMAGE>(g, tmem, m0, n_tile, esub, q, lane, 0, g.NT, true);
      if (threadIdx.x == 64) stamp(5);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CLUSTER > 1) cluster_sync();   // no CTA leaves while a peer may still signal its barriers
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
  if (threadIdx.x == 32) stamp(6);
}

// Row-dot epilogue of the persistent kernel for the 128-wide dot part (kinds 3 / 4), two column parts per row: the 4 x 16
// values of `dotv` a thread needs are fetched BEFORE the accumulators are ready (they do not depend on them), so the drain is
// register work only. The first version loaded them inside the column loop, behind tcgen05.wait::ld: one exposed L2 / HBM
// round trip per 16 columns, 11-13 us of drain per tile for layer 10 (profiles/r2_kg_trace_*.json).
struct DotPrefetch {
  float4 v[4][4];
};
__device__ __forceinline__ void dot_prefetch(const GemmArgs& g, int m0, int esub, int rloc, int part, DotPrefetch& p) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const float4* qv = reinterpret_cast<const float4*>(
        g.dotv + (((size_t)(m0 / 128 + esub) * g.dot_blocks + g.dot_block0 + part * 4 + i) * 128 + rloc) * 16);
#pragma unroll
    for (int j = 0; j < 4; ++j) p.v[i][j] = __ldg(qv + j);
  }
}
// columns [64 part, 64 part + 64) of the dot part; for kind 4 also columns 128 + [64 part, 64 part + 64) -> C2 (column-blocked)
__device__ __forceinline__ float drain_dot2(const GemmArgs& g, uint32_t tmem, int m0, int esub, int q, int lane, int part,
                                            const DotPrefetch& p) {
  const int rloc = q * 32 + lane;
  const float inv = *g.inv_scale * g.out_mul;
  const uint32_t tsrc = tmem + esub * 256 + ((uint32_t)(q * 32) << 16);
  const float floor_ = (g.out_kind == 4 || g.relu) ? 0.f : -INFINITY;   // ReLU on the dot part, or none (key_map_2 / query_repeat_embed_2)
  float dot = 0.f;
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int c0 = (part * 4 + i) * 16;
    float v[16];
    tmem_ld16(tsrc + c0, v);
#pragma unroll
    for (int j = 0; j < 16; j += 4) {
      const float4 b = *reinterpret_cast<const float4*>(g.bias + c0 + j);
      v[j] = fmaxf(v[j] * inv + b.x, floor_);
      v[j + 1] = fmaxf(v[j + 1] * inv + b.y, floor_);
      v[j + 2] = fmaxf(v[j + 2] * inv + b.z, floor_);
      v[j + 3] = fmaxf(v[j + 3] * inv + b.w, floor_);
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float4 x = p.v[i][j];
      dot = fmaf(v[4 * j + 3], x.w, fmaf(v[4 * j + 2], x.z, fmaf(v[4 * j + 1], x.y, fmaf(v[4 * j], x.x, dot))));
    }
  }
  if (g.out_kind == 4) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int c0 = CPN_HIDDEN + (part * 4 + i) * 16;
      float v[16];
      tmem_ld16(tsrc + c0, v);
      float* out = g.C2 + (((size_t)(m0 / 128 + esub) * (CPN_HIDDEN / 16) + part * 4 + i) * 128 + rloc) * 16;
#pragma unroll
      for (int j = 0; j < 16; j += 4) {
        const float4 b = *reinterpret_cast<const float4*>(g.bias + c0 + j);
        *reinterpret_cast<float4*>(out + j) =
            make_float4(v[j] * inv + b.x, v[j + 1] * inv + b.y, v[j + 2] * inv + b.z, v[j + 3] * inv + b.w);
      }
    }
  }
  return dot;
}

// ---- persistent version (operand-image A) -----------------------------------------------------------------------
// One CTA per SM walks the (row tile, N tile) list with a stride of gridDim.x. The operand ring runs across tiles, so
// the copies of tile j + 1 are in flight while tile j is drained, and TMEM / barriers are set up once per launch. A phase
// trace of the one-tile-per-CTA kernel (profiles/r2_gemm1_trace.json) showed 15.9 us of MMA issue per tile against 6.5 us
// of drain by 8 warps (instruction-latency bound), 1.3 us waiting for the first stage and 0.7 us between CTAs: 36 % of an
// SM's time. Here 24 epilogue warps (three per 128-row sub-tile and TMEM lane quadrant, each a third of the columns) drain
// a tile, and the MMA warp restarts as soon as they have read the accumulators (accum_empty).
// A deeper operand ring does not help: with separate rings for the activation stages (4 x 32 KB) and the weight stages
// (3 x 32 KB, a second producer warp) the query_encode_latent GEMM went from 1.33 to 1.44 ms (main loop 16.9 -> 18.1 us per
// tile, profiles/r2_gemm1_trace_splitring.json). What the loop is bound by: DESIGN.md section 4, main-loop attribution.
// EPI_WARPS = 8 * PARTS: 16 leaves registers for a co-resident CTA of another kernel (the gather / readout of the other chunk
// lane: a persistent grid does not block the dispatch of later kernels the way a long CTA queue does).
// CL = 2: the two CTAs of a cluster work on neighbouring N tiles of the SAME row tile (tiles 2p and 2p + 1); each fetches one
// of the two 128-row activation sub-tiles of every stage and multicasts it to both, so an SM reads 42.6 KB instead of 58.6 KB
// per k-chunk from L2. Every CTA issues its own MMAs; the only coupling is that a stage is refilled when BOTH have consumed it
// (tcgen05.commit multicast to both `empty` barriers, count 2).
template <bool OUT_IMAGE, int P_EPI_WARPS, int CL>
__global__ void __launch_bounds__((2 + P_EPI_WARPS) * 32, 1) gemm_tc_persist_kernel(GemmArgs g, int ntiles_n, int ntiles) {
  constexpr int PARTS = P_EPI_WARPS / 8;
  const uint32_t crank = CL > 1 ? cluster_ctarank() : 0;
  // tile list of this CTA: t = t_first, t_first + t_stride, ... (CL = 2: the cluster takes tile pairs, this CTA tile 2p + rank)
  const int t_first = CL > 1 ? (int)(blockIdx.x / CL) * CL + (int)crank : (int)blockIdx.x, t_stride = (int)gridDim.x;
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 2];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sdot[2][PARTS][128];     // row-dot partials of the column parts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[STAGES]), accum_full = smem_u32(&bars[2 * STAGES]),
                 accum_empty = smem_u32(&bars[2 * STAGES + 1]);
  // compact A image (12 KB blocks: fp16 head + remainder plane): the value plane e5m2(head) of a stage is derived in shared
  // memory by the epilogue warps, which are idle during the MMA loop; the MMA thread then waits on `ready` instead of `full`
  const uint32_t ready = smem_u32(&bars[2 * STAGES + 2]);
  const bool a3 = g.a_chunk != ACT_CHUNK_BYTES, w3 = a3 && g.w3;
  const uint32_t a_bytes = a3 ? (uint32_t)ACT_X8 : (uint32_t)A_SUB;
  const int NT = g.NT;
  const uint32_t w_half = (uint32_t)(BK / 8) * NT * 16;
  auto stamp = [&](int tile_no, int i) {   // optional phase trace, 8 x u64 per (CTA, tile) slot
    if (g.dbg) {
      const int slot = tile_no * gridDim.x + blockIdx.x;
      if (slot < g.dbg_cap) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g.dbg[(size_t)slot * 8 + i] = t;
      }
    }
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full + 8 * s, 1);
      mbar_init(empty + 8 * s, CL);   // one commit from the MMA thread of every CTA that receives this stage's activations
    }
    mbar_init(accum_full, 1);
    mbar_init(accum_empty, P_EPI_WARPS);
    for (int s = 0; s < STAGES; ++s) mbar_init(ready + 8 * s, P_EPI_WARPS);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();   // every CTA's barriers exist before a peer multicasts into them
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      const unsigned char* asrc = reinterpret_cast<const unsigned char*>(g.A);
      uint32_t it = 0;
      for (int t = t_first; t < ntiles; t += t_stride) {
        const int n_tile = t % ntiles_n, m0 = (t / ntiles_n) * BM;
        const bool sub1_valid = (m0 + 128) < g.M;
        const unsigned char* wsrc = g.wtiles + (size_t)n_tile * g.kchunks * 2 * w_half;
        const size_t tile0 = (size_t)(m0 / 128);
        for (int i = 0; i < g.kchunks; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t u = it / STAGES;
          mbar_wait(empty + 8 * s, (u & 1) ^ 1);
          const uint32_t stage = smem0 + s * STAGE_BYTES;
          if (g.dbg_skip & 48) {   // trace runs: 16 no activation copies, 32 no weight copies (the MMAs then read stale stages)
            const bool la = !(g.dbg_skip & 16), lw = !(g.dbg_skip & 32);
            mbar_arrive_expect_tx(full + 8 * s, (lw ? 2 * w_half : 0) + (la ? (sub1_valid ? 2 : 1) * a_bytes : 0));
            if (lw) bulk_g2s(stage + 2 * A_SUB, wsrc + (size_t)i * 2 * w_half, 2 * w_half, full + 8 * s);
            if (la) {
              bulk_g2s(stage, asrc + (tile0 * g.kchunks + i) * (size_t)g.a_chunk, a_bytes, full + 8 * s);
              if (sub1_valid) bulk_g2s(stage + A_SUB, asrc + ((tile0 + 1) * g.kchunks + i) * (size_t)g.a_chunk, a_bytes, full + 8 * s);
            }
            continue;
          }
          mbar_arrive_expect_tx(full + 8 * s, (w3 ? w_half + w_half / 2 : 2 * w_half) + (sub1_valid ? 2 : 1) * a_bytes);
          if (w3) {   // fp16 plane and e4m3(w_lo) plane; the plane between them is derived from the first on chip
            bulk_g2s(stage + 2 * A_SUB, wsrc + (size_t)i * 2 * w_half, w_half, full + 8 * s);
            bulk_g2s(stage + 2 * A_SUB + w_half + w_half / 2, wsrc + (size_t)i * 2 * w_half + w_half + w_half / 2, w_half / 2, full + 8 * s);
          } else {
            bulk_g2s(stage + 2 * A_SUB, wsrc + (size_t)i * 2 * w_half, 2 * w_half, full + 8 * s);
          }
          if (CL > 1) {   // this CTA's sub-tile of the shared row tile, delivered to both CTAs (and both `full` barriers)
            if (crank == 0 || sub1_valid)
              bulk_g2s_multicast(stage + crank * A_SUB, asrc + ((tile0 + crank) * g.kchunks + i) * ACT_CHUNK_BYTES, A_SUB,
                                 full + 8 * s, (uint16_t)((1u << CL) - 1));
          } else {
            bulk_g2s(stage, asrc + (tile0 * g.kchunks + i) * (size_t)g.a_chunk, a_bytes, full + 8 * s);
            if (sub1_valid)
              bulk_g2s(stage + A_SUB, asrc + ((tile0 + 1) * g.kchunks + i) * (size_t)g.a_chunk, a_bytes, full + 8 * s);
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t it = 0, tcount = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
        const int m0 = (t / ntiles_n) * BM;
        const bool sub1_valid = (m0 + 128) < g.M;
        mbar_wait(accum_empty, (tcount & 1) ^ 1);     // the epilogue warps have read the previous tile out of TMEM
        tcgen05_fence_after();
        stamp(tcount, 0);
        for (int i = 0; i < g.kchunks; ++i, ++it) {
          const int s = it % STAGES;
          const uint32_t u = it / STAGES;
          mbar_wait((a3 ? ready : full) + 8 * s, u & 1);
          if (i == 0) stamp(tcount, 1);
          tcgen05_fence_after();
          const uint32_t stage = smem0 + s * STAGE_BYTES;
          const uint32_t b_hi = stage + 2 * A_SUB, b_lo = b_hi + w_half;
          if (g.ws && sub1_valid && !(g.dbg_skip & 8)) {
            // weight-stationary order: every weight block is fetched once and used by both sub-tiles
            const uint32_t a0 = stage, a1 = stage + A_SUB, d0 = tmem, d1 = tmem + 256, idf8 = g.idesc | IDESC_A_E5M2;
            const uint64_t b0 = make_desc(b_hi, NT * 16, 128), b1 = make_desc(b_hi + 2 * NT * 16, NT * 16, 128);
            const uint64_t b2 = make_desc(b_hi + w_half, NT * 16, 128), b3 = make_desc(b_hi + w_half + w_half / 2, NT * 16, 128);
            const uint32_t acc0 = i != 0;
            CPN_MMA_WS("f16", "b0", "fill", d0, make_desc(a0, A_LBO, 128), b0, g.idesc, acc0);
            CPN_MMA_WS("f16", "b0", "lastuse", d1, make_desc(a1, A_LBO, 128), b0, g.idesc, acc0);
            CPN_MMA_WS("f16", "b1", "fill", d0, make_desc(a0 + 2 * A_LBO, A_LBO, 128), b1, g.idesc, 1u);
            CPN_MMA_WS("f16", "b1", "lastuse", d1, make_desc(a1 + 2 * A_LBO, A_LBO, 128), b1, g.idesc, 1u);
            CPN_MMA_WS("f8f6f4", "b2", "fill", d0, make_desc(a0 + ACT_LO8, A_LBO, 128), b2, idf8, 1u);
            CPN_MMA_WS("f8f6f4", "b2", "lastuse", d1, make_desc(a1 + ACT_LO8, A_LBO, 128), b2, idf8, 1u);
            CPN_MMA_WS("f8f6f4", "b3", "fill", d0, make_desc(a0 + ACT_X8, A_LBO, 128), b3, idf8, 1u);
            CPN_MMA_WS("f8f6f4", "b3", "lastuse", d1, make_desc(a1 + ACT_X8, A_LBO, 128), b3, idf8, 1u);
          } else
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            if (sub == 1 && !sub1_valid) break;
            const uint32_t a_hi = stage + sub * A_SUB, a_lo = a_hi + A_HALF;
            const uint32_t d = tmem + sub * 256;
            if (g.dbg_skip & 8) break;   // trace runs: no MMAs, the commit below frees the stage at once
            if (g.f8) {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j)
                mma_f16_ss(d, make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128), make_desc(b_hi + j * 2 * NT * 16, NT * 16, 128),
                           g.idesc, (i | j) != 0);
              mma_f8_ss(d, make_desc(a_hi + ACT_LO8, A_LBO, 128), make_desc(b_hi + w_half, NT * 16, 128), g.idesc | IDESC_A_E5M2, 1);
              mma_f8_ss(d, make_desc(a_hi + ACT_X8, A_LBO, 128), make_desc(b_hi + w_half + w_half / 2, NT * 16, 128),
                        g.idesc | IDESC_A_E5M2, 1);
            } else {
#pragma unroll
              for (int j = 0; j < BK / 16; ++j) {
                const uint64_t da_hi = make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128);
                const uint64_t da_lo = make_desc(a_lo + j * 2 * A_LBO, A_LBO, 128);
                const uint64_t db_hi = make_desc(b_hi + j * 2 * NT * 16, NT * 16, 128);
                const uint64_t db_lo = make_desc(b_lo + j * 2 * NT * 16, NT * 16, 128);
                mma_f16_ss(d, da_hi, db_hi, g.idesc, (i | j) != 0);
                mma_f16_ss(d, da_hi, db_lo, g.idesc, 1);
                mma_f16_ss(d, da_lo, db_hi, g.idesc, 1);
              }
            }
          }
          if (CL > 1) mma_commit_multicast(empty + 8 * s, (uint16_t)((1u << CL) - 1)); else mma_commit(empty + 8 * s);
        }
        mma_commit(accum_full);
        stamp(tcount, 2);
      }
    }
  } else {
    // epilogue warp -> (TMEM lane quadrant q = warp % 4, sub-tile, column part): warps 2.. give every quadrant
    // 2 * PARTS warps, r = (warp - 2) / 4 -> sub = r / PARTS, part = r % PARTS
    const int q = warp & 3, r = (warp - 2) >> 2, esub = r / PARTS, part = r % PARTS;
    const int niter = NT / 16, c_lo = (part * niter / PARTS) * 16, c_hi = ((part + 1) * niter / PARTS) * 16;
    const bool dotkind = !OUT_IMAGE && (g.out_kind == 3 || g.out_kind == 4);
    // compact A image: derive the value plane(s) of the stage that holds global k-chunk number `c`, then release it to the MMA thread
    const int idx = (warp - 2) * 32 + lane, csub = idx >> 8, cg = (idx >> 7) & 1, cr = idx & 127;
    auto convert_stage = [&](uint32_t c, bool sub1) {
      const bool conv = idx < 512 && (csub == 0 || sub1);
      const uint32_t cit = c;
      const int s = cit % STAGES;
      mbar_wait(full + 8 * s, (cit / STAGES) & 1);
      if (conv) {
        unsigned char* st_ = smem + s * STAGE_BYTES + csub * A_SUB;
        const uint4 h0 = *reinterpret_cast<const uint4*>(st_ + (2 * cg) * A_LBO + cr * 16);
        const uint4 h1 = *reinterpret_cast<const uint4*>(st_ + (2 * cg + 1) * A_LBO + cr * 16);
        const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        uint32_t o[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const uint32_t lo2 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&hw[2 * k2]), __NV_SATFINITE, __NV_E5M2);
          const uint32_t hi2 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&hw[2 * k2 + 1]), __NV_SATFINITE, __NV_E5M2);
          o[k2] = lo2 | (hi2 << 16);
        }
        *reinterpret_cast<uint4*>(st_ + ACT_X8 + cg * A_LBO + cr * 16) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      if (w3 && idx < 2 * NT) {
        // weight plane e4m3(w_hi 2^-10) of this stage: item -> (16-k group, weight row); bit-identical to the plane
        // pack_tc_kernel stores (the product is exact in fp16 wherever e4m3 does not round it to zero anyway)
        const int wg = idx >= NT, wn = idx - wg * NT;
        unsigned char* wb = smem + s * STAGE_BYTES + 2 * A_SUB;
        const uint4 h0 = *reinterpret_cast<const uint4*>(wb + ((2 * wg) * NT + wn) * 16);
        const uint4 h1 = *reinterpret_cast<const uint4*>(wb + ((2 * wg + 1) * NT + wn) * 16);
        const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
        const __half2 sc = __float2half2_rn(F8_W_SCALE);
        uint32_t o[4];
#pragma unroll
        for (int k2 = 0; k2 < 4; ++k2) {
          const __half2 p0 = __hmul2(*reinterpret_cast<const __half2*>(&hw[2 * k2]), sc);
          const __half2 p1 = __hmul2(*reinterpret_cast<const __half2*>(&hw[2 * k2 + 1]), sc);
          const uint32_t lo2 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&p0), __NV_SATFINITE, __NV_E4M3);
          const uint32_t hi2 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&p1), __NV_SATFINITE, __NV_E4M3);
          o[k2] = lo2 | (hi2 << 16);
        }
        *reinterpret_cast<uint4*>(wb + w_half + (wg * NT + wn) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
      }
      fence_proxy_async_smem();
      __syncwarp();
      if (lane == 0) mbar_arrive(ready + 8 * s);
    };
    int pre = 0;
    uint32_t tcount = 0, cit = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
      const int n_tile = t % ntiles_n, m0 = (t / ntiles_n) * BM;
      const bool valid = esub == 0 || (m0 + 128) < g.M;
      constexpr bool PREFETCH = !OUT_IMAGE && PARTS == 2;   // dot kinds run on the 16-warp kernel (launcher)
      DotPrefetch pf;
      if (PREFETCH && dotkind && valid) dot_prefetch(g, m0, esub, q * 32 + lane, part, pf);
      if (a3) {
        // value plane of every stage of this tile: thread -> (sub-tile, 16-k group, row): two 16-byte fp16 groups in, one
        // 16-byte e5m2 group out (the same conversion, on the same fp16 values, as split4_f8 writes into a full image).
        // The first `pre` stages were converted at the end of the previous tile's epilogue (below).
        for (int i = pre; i < g.kchunks; ++i, ++cit) convert_stage(cit, (m0 + 128) < g.M);
        pre = 0;
      }
      mbar_wait(accum_full, tcount & 1);
      tcgen05_fence_after();
      if (threadIdx.x == 64) stamp(tcount, 3);
      float dot = 0.f;
      if (valid) {
        if (PREFETCH && dotkind) dot = drain_dot2(g, tmem, m0, esub, q, lane, part, pf);
        else dot = drain_subtile<OUT_IMAGE>(g, tmem, m0, n_tile, esub, q, lane, c_lo, c_hi, false);
      }
      // every TMEM read of this warp has completed (tcgen05.wait::ld inside tmem_ld16): hand the accumulators back
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accum_empty);
      if (a3 && t + (int)gridDim.x < ntiles) {
        // the MMA thread may start the next tile now and its first stages landed during the drain: convert them before the rest
        // of this epilogue (it used to wait 4 us per tile for them, profiles/r2_kg_weight_stationary.log: ring_wait_at_tile_start)
        const int m0n = ((t + (int)gridDim.x) / ntiles_n) * BM;
        pre = g.kchunks < STAGES ? g.kchunks : STAGES;
        for (int i = 0; i < pre; ++i, ++cit) convert_stage(cit, (m0n + 128) < g.M);
      }
      if (dotkind) {   // the three column parts of a row meet in shared memory, summed in part order
        const int rloc = q * 32 + lane;
        sdot[esub][part][rloc] = dot;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + esub * 4 + q), "n"(32 * PARTS) : "memory");
        if (part == 0 && valid) {
          const int row = m0 + esub * 128 + rloc;
          if (row < g.M) {
            float tot = sdot[esub][0][rloc];
#pragma unroll
            for (int pp = 1; pp < PARTS; ++pp) tot += sdot[esub][pp][rloc];
            reinterpret_cast<float*>(g.C)[row] = (g.dot_rowadd ? tot + g.dot_rowadd[row] : tot) / g.dot_div;
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + esub * 4 + q), "n"(32 * PARTS) : "memory");   // sdot is free for the next tile
      }
      if (threadIdx.x == 64) stamp(tcount, 4);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (CL > 1) cluster_sync();   // no CTA leaves while a peer may still multicast into its stages or signal its barriers
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
  if (g.dbg && threadIdx.x == 0) {
    unsigned smid;
    asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
    if ((int)blockIdx.x < g.dbg_cap) g.dbg[(size_t)blockIdx.x * 8 + 7] = smid;
  }
}

// ---- persistent version 2: sub-tile pipelining (experiment, opt-in with CPN_TC_PERSIST2=1) ---------------------------------------------------------
// Main-loop attribution of the kernel above on the query_encode_latent GEMM (profiles/r2_gemm1_mainloop_attribution.log,
// per 256-row tile): MMAs alone 13.9 us (the tensor pipe at its full rate at the power-capped 1.6 GHz), copies alone 11.5 us,
// together 16.8 us, then 4.7 us of drain and 1.6 us of hand-over during which the tensor pipe idles: 13.9 of 23.2 us = 60 %.
// Both 208-column accumulators of a tile are ready at the same moment and TMEM (512 columns) has no room for a second pair.
// Here the two 128-row sub-tiles of a tile run one after the other, each over the whole K: while the MMA thread accumulates
// sub-tile 1 in TMEM columns 256.., eight epilogue warps drain sub-tile 0 from columns 0.., and the other eight drain sub-tile 1
// during sub-tile 0 of the next tile. The price is that the weight tile is staged once per 128 rows instead of once per 256
// (it is L2-resident: weights alone stream at 139 GB/s per SM against 80 GB/s for the activations, same log); a stage is
// one activation block plus the weight block (42.6 KB at NT = 208), so the ring holds five stages instead of three.
// Two extra warps derive the value planes of compact operand images (the epilogue warps are no longer idle during the loop).
constexpr int P2_EPI_WARPS = 16, P2_CONV_WARPS = 2;
constexpr int P2_THREADS = (2 + P2_EPI_WARPS + P2_CONV_WARPS) * 32;
constexpr int P2_MAX_STAGES = 8;
constexpr int P2_SMEM_BYTES = 220 * 1024;

template <bool OUT_IMAGE>
__global__ void __launch_bounds__(P2_THREADS, 1) gemm_tc_persist2_kernel(GemmArgs g, int ntiles_n, int ntiles) {
  constexpr int PARTS = P2_EPI_WARPS / 8;   // column parts per sub-tile
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[3 * P2_MAX_STAGES + 4];
  __shared__ uint32_t tmem_base_s;
  __shared__ float sdot[2][PARTS][128];     // row-dot partials of the column parts
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full = smem_u32(&bars[0]), empty = smem_u32(&bars[P2_MAX_STAGES]), ready = smem_u32(&bars[2 * P2_MAX_STAGES]),
                 accum_full = smem_u32(&bars[3 * P2_MAX_STAGES]), accum_empty = smem_u32(&bars[3 * P2_MAX_STAGES + 2]);
  const bool a3 = g.a_chunk != ACT_CHUNK_BYTES, w3 = a3 && g.w3;
  const uint32_t a_bytes = a3 ? (uint32_t)ACT_X8 : (uint32_t)A_SUB;
  const int NT = g.NT;
  const uint32_t w_half = (uint32_t)(BK / 8) * NT * 16;
  const uint32_t stage_bytes = A_SUB + 2 * w_half;
  const uint32_t nstages = min((uint32_t)P2_MAX_STAGES, (uint32_t)P2_SMEM_BYTES / stage_bytes);
  auto stamp = [&](int tile_no, int sub, int i) {   // optional phase trace, 8 x u64 per (CTA, tile, sub-tile) slot
    if (g.dbg) {
      const int slot = (tile_no * 2 + sub) * gridDim.x + blockIdx.x;
      if (slot < g.dbg_cap) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g.dbg[(size_t)slot * 8 + i] = t;
      }
    }
  };
  if (threadIdx.x == 0) {
    for (uint32_t s = 0; s < nstages; ++s) {
      mbar_init(full + 8 * s, 1);
      mbar_init(empty + 8 * s, 1);
      mbar_init(ready + 8 * s, P2_CONV_WARPS);
    }
    for (int b = 0; b < 2; ++b) {
      mbar_init(accum_full + 8 * b, 1);
      mbar_init(accum_empty + 8 * b, P2_EPI_WARPS / 2);
    }
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      const unsigned char* asrc = reinterpret_cast<const unsigned char*>(g.A);
      uint32_t s = 0, u = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int n_tile = t % ntiles_n, m0 = (t / ntiles_n) * BM;
        const int nsub = (m0 + 128) < g.M ? 2 : 1;
        const unsigned char* wsrc = g.wtiles + (size_t)n_tile * g.kchunks * 2 * w_half;
        for (int sub = 0; sub < nsub; ++sub) {
          const unsigned char* arow = asrc + (size_t)(m0 / 128 + sub) * g.kchunks * (size_t)g.a_chunk;
          for (int i = 0; i < g.kchunks; ++i) {
            mbar_wait(empty + 8 * s, (u & 1) ^ 1);
            const uint32_t stage = smem0 + s * stage_bytes;
            const bool la = !(g.dbg_skip & 16), lw = !(g.dbg_skip & 32);   // trace runs: no activation / no weight copies
            const uint32_t wb = w3 ? w_half + w_half / 2 : 2 * w_half;
            mbar_arrive_expect_tx(full + 8 * s, (lw ? wb : 0) + (la ? a_bytes : 0));
            if (lw) {
              if (w3) {   // fp16 plane and e4m3(w_lo) plane; the plane between them is derived from the first on chip
                bulk_g2s(stage + A_SUB, wsrc + (size_t)i * 2 * w_half, w_half, full + 8 * s);
                bulk_g2s(stage + A_SUB + w_half + w_half / 2, wsrc + (size_t)i * 2 * w_half + w_half + w_half / 2, w_half / 2, full + 8 * s);
              } else {
                bulk_g2s(stage + A_SUB, wsrc + (size_t)i * 2 * w_half, 2 * w_half, full + 8 * s);
              }
            }
            if (la) bulk_g2s(stage, arow + (size_t)i * g.a_chunk, a_bytes, full + 8 * s);
            if (++s == nstages) { s = 0; ++u; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      uint32_t s = 0, u = 0, cnt0 = 0, cnt1 = 0;
      int tcount = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
        const int m0 = (t / ntiles_n) * BM;
        const int nsub = (m0 + 128) < g.M ? 2 : 1;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          if (sub >= nsub) break;
          uint32_t& cnt = sub ? cnt1 : cnt0;
          mbar_wait(accum_empty + 8 * sub, (cnt & 1) ^ 1);   // the epilogue warps of this sub-tile slot have read its last contents
          tcgen05_fence_after();
          stamp(tcount, sub, 0);
          const uint32_t d = tmem + sub * 256;
          for (int i = 0; i < g.kchunks; ++i) {
            mbar_wait((a3 ? ready : full) + 8 * s, u & 1);
            if (i == 0) stamp(tcount, sub, 1);
            tcgen05_fence_after();
            const uint32_t a_hi = smem0 + s * stage_bytes, a_lo = a_hi + A_HALF;
            const uint32_t b_hi = a_hi + A_SUB, b_lo = b_hi + w_half;
            if (!(g.dbg_skip & 8)) {
              if (g.f8) {
#pragma unroll
                for (int j = 0; j < BK / 16; ++j)
                  mma_f16_ss(d, make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128), make_desc(b_hi + j * 2 * NT * 16, NT * 16, 128),
                             g.idesc, (i | j) != 0);
                mma_f8_ss(d, make_desc(a_hi + ACT_LO8, A_LBO, 128), make_desc(b_hi + w_half, NT * 16, 128), g.idesc | IDESC_A_E5M2, 1);
                mma_f8_ss(d, make_desc(a_hi + ACT_X8, A_LBO, 128), make_desc(b_hi + w_half + w_half / 2, NT * 16, 128),
                          g.idesc | IDESC_A_E5M2, 1);
              } else {
#pragma unroll
                for (int j = 0; j < BK / 16; ++j) {
                  const uint64_t da_hi = make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128);
                  const uint64_t da_lo = make_desc(a_lo + j * 2 * A_LBO, A_LBO, 128);
                  const uint64_t db_hi = make_desc(b_hi + j * 2 * NT * 16, NT * 16, 128);
                  const uint64_t db_lo = make_desc(b_lo + j * 2 * NT * 16, NT * 16, 128);
                  mma_f16_ss(d, da_hi, db_hi, g.idesc, (i | j) != 0);
                  mma_f16_ss(d, da_hi, db_lo, g.idesc, 1);
                  mma_f16_ss(d, da_lo, db_hi, g.idesc, 1);
                }
              }
            }
            mma_commit(empty + 8 * s);
            if (++s == nstages) { s = 0; ++u; }
          }
          mma_commit(accum_full + 8 * sub);
          stamp(tcount, sub, 2);
          ++cnt;
        }
      }
    }
  } else if (warp >= 2 + P2_EPI_WARPS) {
    if (a3) {
      // value plane e5m2(head) of every activation stage (and, w3, the weight plane e4m3(w_hi 2^-10)): item -> (16-k group,
      // row): two 16-byte fp16 groups in, one 16-byte fp8 group out; the same conversions, on the same fp16 values, as
      // split4_f8 / pack_tc_kernel write into full images
      const int cidx = (warp - 2 - P2_EPI_WARPS) * 32 + lane;
      uint32_t s = 0, u = 0;
      for (int t = blockIdx.x; t < ntiles; t += gridDim.x) {
        const int m0 = (t / ntiles_n) * BM;
        const int nsub = (m0 + 128) < g.M ? 2 : 1;
        for (int it = 0; it < nsub * g.kchunks; ++it) {
          mbar_wait(full + 8 * s, u & 1);
          unsigned char* st_ = smem + s * stage_bytes;
          for (int j = cidx; j < 256; j += P2_CONV_WARPS * 32) {
            const int cg = j >> 7, cr = j & 127;
            const uint4 h0 = *reinterpret_cast<const uint4*>(st_ + (2 * cg) * A_LBO + cr * 16);
            const uint4 h1 = *reinterpret_cast<const uint4*>(st_ + (2 * cg + 1) * A_LBO + cr * 16);
            const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
            uint32_t o[4];
#pragma unroll
            for (int k2 = 0; k2 < 4; ++k2) {
              const uint32_t lo2 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&hw[2 * k2]), __NV_SATFINITE, __NV_E5M2);
              const uint32_t hi2 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&hw[2 * k2 + 1]), __NV_SATFINITE, __NV_E5M2);
              o[k2] = lo2 | (hi2 << 16);
            }
            *reinterpret_cast<uint4*>(st_ + ACT_X8 + cg * A_LBO + cr * 16) = make_uint4(o[0], o[1], o[2], o[3]);
          }
          if (w3) {
            unsigned char* wb = st_ + A_SUB;
            const __half2 sc = __float2half2_rn(F8_W_SCALE);
            for (int j = cidx; j < 2 * NT; j += P2_CONV_WARPS * 32) {
              const int wg = j >= NT, wn = j - wg * NT;
              const uint4 h0 = *reinterpret_cast<const uint4*>(wb + ((2 * wg) * NT + wn) * 16);
              const uint4 h1 = *reinterpret_cast<const uint4*>(wb + ((2 * wg + 1) * NT + wn) * 16);
              const uint32_t hw[8] = {h0.x, h0.y, h0.z, h0.w, h1.x, h1.y, h1.z, h1.w};
              uint32_t o[4];
#pragma unroll
              for (int k2 = 0; k2 < 4; ++k2) {
                const __half2 p0 = __hmul2(*reinterpret_cast<const __half2*>(&hw[2 * k2]), sc);
                const __half2 p1 = __hmul2(*reinterpret_cast<const __half2*>(&hw[2 * k2 + 1]), sc);
                const uint32_t lo2 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&p0), __NV_SATFINITE, __NV_E4M3);
                const uint32_t hi2 = __nv_cvt_halfraw2_to_fp8x2(*reinterpret_cast<const __half2_raw*>(&p1), __NV_SATFINITE, __NV_E4M3);
                o[k2] = lo2 | (hi2 << 16);
              }
              *reinterpret_cast<uint4*>(wb + w_half + (wg * NT + wn) * 16) = make_uint4(o[0], o[1], o[2], o[3]);
            }
          }
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) mbar_arrive(ready + 8 * s);
          if (++s == nstages) { s = 0; ++u; }
        }
      }
    }
  } else {
    // epilogue warp -> (TMEM lane quadrant q = warp % 4, sub-tile slot, column part): r = (warp - 2) / 4 -> esub = r / PARTS
    const int q = warp & 3, r = (warp - 2) >> 2, esub = r / PARTS, part = r % PARTS;
    const int niter = NT / 16, c_lo = (part * niter / PARTS) * 16, c_hi = ((part + 1) * niter / PARTS) * 16;
    const bool dotkind = !OUT_IMAGE && (g.out_kind == 3 || g.out_kind == 4);
    uint32_t cnt = 0;
    int tcount = 0;
    for (int t = blockIdx.x; t < ntiles; t += gridDim.x, ++tcount) {
      const int n_tile = t % ntiles_n, m0 = (t / ntiles_n) * BM;
      if (esub == 1 && (m0 + 128) >= g.M) continue;   // no second sub-tile: every warp of this slot skips together
      DotPrefetch pf;
      if (!OUT_IMAGE && dotkind) dot_prefetch(g, m0, esub, q * 32 + lane, part, pf);
      mbar_wait(accum_full + 8 * esub, cnt & 1);
      ++cnt;
      tcgen05_fence_after();
      if (lane == 0 && q == 2 && part == 0) stamp(tcount, esub, 3);   // warps 2 / 10
      float dot = 0.f;
      if (!OUT_IMAGE && dotkind) dot = drain_dot2(g, tmem, m0, esub, q, lane, part, pf);
      else dot = drain_subtile<OUT_IMAGE>(g, tmem, m0, n_tile, esub, q, lane, c_lo, c_hi, false);
      // every TMEM read of this warp has completed (tcgen05.wait::ld inside tmem_ld16): hand the accumulators back
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(accum_empty + 8 * esub);
      if (dotkind) {   // the column parts of a row meet in shared memory, summed in part order
        const int rloc = q * 32 + lane;
        sdot[esub][part][rloc] = dot;
        asm volatile("bar.sync %0, %1;" ::"r"(1 + esub * 4 + q), "n"(32 * PARTS) : "memory");
        if (part == 0) {
          const int row = m0 + esub * 128 + rloc;
          if (row < g.M) {
            float tot = sdot[esub][0][rloc];
#pragma unroll
            for (int pp = 1; pp < PARTS; ++pp) tot += sdot[esub][pp][rloc];
            reinterpret_cast<float*>(g.C)[row] = (g.dot_rowadd ? tot + g.dot_rowadd[row] : tot) / g.dot_div;
          }
        }
        asm volatile("bar.sync %0, %1;" ::"r"(1 + esub * 4 + q), "n"(32 * PARTS) : "memory");   // sdot is free for the next tile
      }
      if (lane == 0 && q == 2 && part == 0) stamp(tcount, esub, 4);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

// ---- CTA-pair version (cta_group::2, operand-image A, f8 scheme) -------------------------------------------------
// Two CTAs of a cluster own 2 x 256 rows and the same N tile. Every MMA spans both (M = 256: 128 rows from each
// CTA's A stage) and reads half of the weight tile from each CTA's shared memory, so each SM stages only NT / 2
// weight rows per k-chunk: 45.3 KB instead of 58.6 KB cross the L2 -> SM fabric per chunk and SM (believed to be the bound of
// this GEMM), and the smaller stage allows a 4-deep ring. The leader CTA's MMA thread issues for the pair; the
// peer relays "my stage has landed" to the leader with a remote mbarrier arrive; tcgen05.commit multicasts
// "stage free" / "accumulators ready" to both CTAs.
// All waits are plain (CTA-scope) mbarrier waits: the operands travel through the async proxy (bulk copies in, tensor-core
// reads) and are ordered by the barrier completions themselves. The first version spun on try_wait.acquire.cluster, which
// ptxas lowers to a loop around CCTL.IVALL (an L1 invalidate per spin): 61 % of that kernel's issue samples
// (profiles/r2_ncu_pair_kernel_sass_hotspots.txt) and the reason it measured slower than independent CTAs.
constexpr int PSTAGES = 4;
constexpr int PW_STAGE_MAX = (NT_MAX / 2) * 128;                 // bytes of half a weight tile per k-chunk
constexpr int PSTAGE_BYTES = 2 * A_SUB + PW_STAGE_MAX;
constexpr int PSMEM_BYTES = PSTAGES * PSTAGE_BYTES + 256;

template <bool OUT_IMAGE>
__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_pair_kernel(GemmArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[3 * PSTAGES + 1];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_tile = blockIdx.y, m0 = (blockIdx.z * 2 + (int)rank) * BM;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full = smem_u32(&bars[0]), peer_full = smem_u32(&bars[PSTAGES]), empty = smem_u32(&bars[2 * PSTAGES]),
                 accum = smem_u32(&bars[3 * PSTAGES]);
  const int NT = g.NT, NH = NT / 2;
  const uint32_t wh = (uint32_t)(BK / 8) * NH * 16;          // fp16 plane of this CTA's half tile; 8-bit planes wh / 2 each
  const int cta_lin = (blockIdx.z * gridDim.y + blockIdx.y) * 2 + blockIdx.x;
  unsigned long long* const dbg = (g.dbg && cta_lin < g.dbg_cap) ? g.dbg + (size_t)cta_lin * 8 : nullptr;
  auto stamp = [&](int i) {   // same slots as gemm_tc_kernel
    if (dbg) {
      unsigned long long t;
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
      dbg[i] = t;
    }
  };
  if (threadIdx.x == 0) stamp(0);

  if (threadIdx.x == 0) {
    for (int s = 0; s < PSTAGES; ++s) {
      mbar_init(full + 8 * s, 1);
      mbar_init(peer_full + 8 * s, 1);
      mbar_init(empty + 8 * s, 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(&tmem_base_s), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;
  if (threadIdx.x == 0) {
    stamp(1);
    if (dbg) {
      unsigned smid;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      dbg[7] = smid;
    }
  }

  if (warp == 0) {
    if (lane == 0) {
      // this CTA's half of the weight tile and its own two A sub-tiles
      const unsigned char* wsrc = g.wtiles + ((size_t)n_tile * 2 + rank) * g.kchunks * 2 * wh;
      const unsigned char* asrc = reinterpret_cast<const unsigned char*>(g.A);
      const size_t tile0 = (size_t)(m0 / 128);
      for (int i = 0; i < g.kchunks; ++i) {
        int s = i % PSTAGES;
        uint32_t u = i / PSTAGES;
        mbar_wait(empty + 8 * s, (u & 1) ^ 1);
        uint32_t stage = smem0 + s * PSTAGE_BYTES;
        mbar_arrive_expect_tx(full + 8 * s, 2 * wh + 2 * A_SUB);
        bulk_g2s(stage + 2 * A_SUB, wsrc + (size_t)i * 2 * wh, 2 * wh, full + 8 * s);
        bulk_g2s(stage, asrc + (tile0 * g.kchunks + i) * ACT_CHUNK_BYTES, A_SUB, full + 8 * s);
        bulk_g2s(stage + A_SUB, asrc + ((tile0 + 1) * g.kchunks + i) * ACT_CHUNK_BYTES, A_SUB, full + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 1) {
      // peer: tell the leader when a stage of this CTA has landed
      for (int i = 0; i < g.kchunks; ++i) {
        int s = i % PSTAGES;
        uint32_t u = i / PSTAGES;
        mbar_wait(full + 8 * s, u & 1);
        mbar_arrive_remote(peer_full + 8 * s, 0);
      }
    } else if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(256, NT);
      for (int i = 0; i < g.kchunks; ++i) {
        int s = i % PSTAGES;
        uint32_t u = i / PSTAGES;
        mbar_wait(full + 8 * s, u & 1);
        mbar_wait(peer_full + 8 * s, u & 1);
        if (i == 0) stamp(2);
        tcgen05_fence_after();
        uint32_t stage = smem0 + s * PSTAGE_BYTES;
        uint32_t b_hi = stage + 2 * A_SUB;
#pragma unroll
        for (int sub = 0; sub < 2; ++sub) {
          uint32_t a_hi = stage + sub * A_SUB;
          uint32_t d = tmem + sub * 256;
#pragma unroll
          for (int j = 0; j < BK / 16; ++j)
            mma_f16_ss_pair(d, make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128), make_desc(b_hi + j * 2 * NH * 16, NH * 16, 128),
                            idesc, (i | j) != 0);
          mma_f8_ss_pair(d, make_desc(a_hi + ACT_LO8, A_LBO, 128), make_desc(b_hi + wh, NH * 16, 128), idesc | IDESC_A_E5M2, 1);
          mma_f8_ss_pair(d, make_desc(a_hi + ACT_X8, A_LBO, 128), make_desc(b_hi + wh + wh / 2, NH * 16, 128), idesc | IDESC_A_E5M2, 1);
        }
        mma_commit_pair(empty + 8 * s, 3);   // both CTAs may refill the stage once these MMAs have read it
      }
      mma_commit_pair(accum, 3);
      stamp(3);
    }
  } else {
    const int pwarp = warp - 2, esub = pwarp >> 2, q = warp & 3;
    mbar_wait(accum, 0);
    if (threadIdx.x == 64) stamp(4);
    tcgen05_fence_after();
    drain_subtile<OUT_IMAGE>(g, tmem, m0, n_tile, esub, q, lane, 0, g.NT, true);
    if (threadIdx.x == 64) stamp(5);
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc_pair(tmem, TMEM_COLS);
  if (threadIdx.x == 32) stamp(6);
}

// ---- persistent CTA-pair version (cta_group::2, operand-image A, f8 scheme) --------------------------------------
// The pair kernel above spends 9.95 us of its 31.6 us per tile in the MMA loop (profiles/r2_gemm1_trace_pair.json; the
// independent-CTA kernel: 15.9 of 25.1 us): every MMA spans the two SMs, each stages half of the weight tile, so a k-chunk
// costs 45.3 KB of L2 -> SM traffic per SM instead of 58.6 KB and the ring holds four stages. What made it slower overall was
// the per-tile overhead of a cluster (TMEM allocation, three cluster barriers, 8 epilogue warps). Here a cluster of two CTAs
// is persistent: it walks the (512-row tile, N tile) list, the rings run across tiles, 16 epilogue warps per CTA drain, and the
// leader's MMA thread restarts when the epilogue warps of BOTH CTAs have read their accumulators (remote mbarrier arrives).
constexpr int PP_EPI_WARPS = 16;
constexpr int PP_THREADS = (2 + PP_EPI_WARPS) * 32;

template <bool OUT_IMAGE>
__global__ void __launch_bounds__(PP_THREADS, 1) gemm_tc_ppair_kernel(GemmArgs g, int ntiles_n, int nptiles) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[3 * PSTAGES + 2];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int cluster_id = blockIdx.x >> 1, nclusters = gridDim.x >> 1;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full = smem_u32(&bars[0]), peer_full = smem_u32(&bars[PSTAGES]), empty = smem_u32(&bars[2 * PSTAGES]),
                 accum_full = smem_u32(&bars[3 * PSTAGES]), accum_empty = smem_u32(&bars[3 * PSTAGES + 1]);
  const int NT = g.NT, NH = NT / 2;
  const uint32_t wh = (uint32_t)(BK / 8) * NH * 16;          // fp16 plane of this CTA's half tile; byte planes wh / 2 each
  auto stamp = [&](int tile_no, int i) {
    if (g.dbg) {
      const int slot = tile_no * gridDim.x + blockIdx.x;
      if (slot < g.dbg_cap) {
        unsigned long long t;
        asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
        g.dbg[(size_t)slot * 8 + i] = t;
      }
    }
  };
  if (threadIdx.x == 0) {
    for (int s = 0; s < PSTAGES; ++s) {
      mbar_init(full + 8 * s, 1);
      mbar_init(peer_full + 8 * s, 1);
      mbar_init(empty + 8 * s, 1);
    }
    mbar_init(accum_full, 1);
    mbar_init(accum_empty, 2 * PP_EPI_WARPS);   // the leader's: every epilogue warp of both CTAs arrives on it
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc_pair(smem_u32(&tmem_base_s), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      const unsigned char* asrc = reinterpret_cast<const unsigned char*>(g.A);
      uint32_t it = 0;
      for (int pt = cluster_id; pt < nptiles; pt += nclusters) {
        const int n_tile = pt % ntiles_n, m0 = ((pt / ntiles_n) * 2 + (int)rank) * BM;
        const unsigned char* wsrc = g.wtiles + ((size_t)n_tile * 2 + rank) * g.kchunks * 2 * wh;
        const size_t tile0 = (size_t)(m0 / 128);
        for (int i = 0; i < g.kchunks; ++i, ++it) {
          const int s = it % PSTAGES;
          const uint32_t u = it / PSTAGES;
          mbar_wait(empty + 8 * s, (u & 1) ^ 1);
          const uint32_t stage = smem0 + s * PSTAGE_BYTES;
          mbar_arrive_expect_tx(full + 8 * s, 2 * wh + 2 * A_SUB);
          bulk_g2s(stage + 2 * A_SUB, wsrc + (size_t)i * 2 * wh, 2 * wh, full + 8 * s);
          bulk_g2s(stage, asrc + (tile0 * g.kchunks + i) * ACT_CHUNK_BYTES, A_SUB, full + 8 * s);
          bulk_g2s(stage + A_SUB, asrc + ((tile0 + 1) * g.kchunks + i) * ACT_CHUNK_BYTES, A_SUB, full + 8 * s);
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && rank == 1) {
      // peer: tell the leader when a stage of this CTA has landed
      uint32_t it = 0;
      for (int pt = cluster_id; pt < nptiles; pt += nclusters)
        for (int i = 0; i < g.kchunks; ++i, ++it) {
          const int s = it % PSTAGES;
          mbar_wait(full + 8 * s, (it / PSTAGES) & 1);
          mbar_arrive_remote(peer_full + 8 * s, 0);
        }
    } else if (lane == 0) {
      const uint32_t idesc = make_idesc_f16(256, NT);
      uint32_t it = 0, tcount = 0;
      for (int pt = cluster_id; pt < nptiles; pt += nclusters, ++tcount) {
        mbar_wait(accum_empty, (tcount & 1) ^ 1);     // both CTAs' epilogue warps have read the previous tile out of TMEM
        tcgen05_fence_after();
        stamp(tcount, 0);
        for (int i = 0; i < g.kchunks; ++i, ++it) {
          const int s = it % PSTAGES;
          const uint32_t u = it / PSTAGES;
          mbar_wait(full + 8 * s, u & 1);
          mbar_wait(peer_full + 8 * s, u & 1);
          if (i == 0) stamp(tcount, 1);
          tcgen05_fence_after();
          const uint32_t stage = smem0 + s * PSTAGE_BYTES;
          const uint32_t b_hi = stage + 2 * A_SUB;
#pragma unroll
          for (int sub = 0; sub < 2; ++sub) {
            const uint32_t a_hi = stage + sub * A_SUB;
            const uint32_t d = tmem + sub * 256;
#pragma unroll
            for (int j = 0; j < BK / 16; ++j)
              mma_f16_ss_pair(d, make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128), make_desc(b_hi + j * 2 * NH * 16, NH * 16, 128),
                              idesc, (i | j) != 0);
            mma_f8_ss_pair(d, make_desc(a_hi + ACT_LO8, A_LBO, 128), make_desc(b_hi + wh, NH * 16, 128), idesc | IDESC_A_E5M2, 1);
            mma_f8_ss_pair(d, make_desc(a_hi + ACT_X8, A_LBO, 128), make_desc(b_hi + wh + wh / 2, NH * 16, 128),
                           idesc | IDESC_A_E5M2, 1);
          }
          mma_commit_pair(empty + 8 * s, 3);   // both CTAs may refill the stage once these MMAs have read it
        }
        mma_commit_pair(accum_full, 3);
        stamp(tcount, 2);
      }
    }
  } else {
    // epilogue warps 2..17: quadrant q = warp % 4, r = (warp - 2) / 4 -> sub-tile r / 2, column part r % 2
    const int q = warp & 3, r = (warp - 2) >> 2, esub = r >> 1, part = r & 1;
    const int niter = NT / 16, c_lo = (part * niter / 2) * 16, c_hi = ((part + 1) * niter / 2) * 16;
    uint32_t tcount = 0;
    for (int pt = cluster_id; pt < nptiles; pt += nclusters, ++tcount) {
      const int n_tile = pt % ntiles_n, m0 = ((pt / ntiles_n) * 2 + (int)rank) * BM;
      mbar_wait(accum_full, tcount & 1);
      tcgen05_fence_after();
      if (threadIdx.x == 64) stamp(tcount, 3);
      drain_subtile<OUT_IMAGE>(g, tmem, m0, n_tile, esub, q, lane, c_lo, c_hi, false);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (rank == 0) mbar_arrive(accum_empty);
        else mbar_arrive_remote(accum_empty, 0);
      }
      if (threadIdx.x == 64) stamp(tcount, 4);
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  cluster_sync();
  if (warp == 1) tmem_dealloc_pair(tmem, TMEM_COLS);
}

}  // namespace

// phase trace of the single-CTA kernel (profiling hook like cpn_prof_begin; not thread-safe): while set, every
// gemm_tc_kernel launch writes 8 x u64 per CTA for its first `cap` CTAs
static unsigned long long* g_tc_dbg = nullptr;
static int g_tc_dbg_cap = 0;
extern "C" int cpn_gemm_tc_trace(unsigned long long* device_buf, int cap_ctas) {
  g_tc_dbg = device_buf;
  g_tc_dbg_cap = device_buf ? cap_ctas : 0;
  return CPN_OK;
}

size_t cpn_tc_weights_bytes() { return TC_HEADER_BYTES + 3 * scheme_bytes(); }

int cpn_pack_tc_weights(const float* raw, const float* packed_fp32, void* dst_v, cudaStream_t st) {
  unsigned char* dst = reinterpret_cast<unsigned char*>(dst_v);
  float* header = reinterpret_cast<float*>(dst);
  unsigned int* absmax = reinterpret_cast<unsigned int*>(dst) + 32;
  CPN_CHECK_CUDA(cudaMemsetAsync(dst, 0, cpn_tc_weights_bytes(), st));
  for (int l = 0; l < CPN_TC_LAYERS; ++l) {
    const TcLayer& L = kLayers[l];
    size_t n = (size_t)L.out * L.in;
    const float* w = (L.folded ? packed_fp32 : raw) + L.raw;
    absmax_kernel<<<64, 256, 0, st>>>(w, n, absmax + l);
    CPN_CHECK_LAUNCH("absmax_kernel");
    size_t total = (size_t)L.out * L.kpad;
    pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, L.out, L.in, L.kpad, L.nt, absmax + l,
                                                                    reinterpret_cast<__half*>(dst + layer_offset(l, 0)),
                                                                    dst + layer_offset(l, 1), dst + layer_offset(l, 2), header, l);
    CPN_CHECK_LAUNCH("pack_tc_kernel");
  }
  return CPN_OK;
}

int launch_gemm_tc(const void* packed, int layer, const void* A, int lda, void* C, int ldc, int M, int relu, int mode,
                   int out_div, int out_kchunks, cudaStream_t st, const float* dotv, float dot_div, const float* dot_rowadd, int dot_blocks,
                   int dot_block0, float* c2) {
  const bool a_img = mode & CPN_TC_A_IMAGE, o_img = mode & CPN_TC_OUT_IMAGE;
  if (!packed || !A || !C || layer < 0 || layer >= CPN_TC_LAYERS || M < 0 || (!a_img && (lda & 3)) || (!o_img && !(mode & (CPN_TC_OUT_ROWDOT | CPN_TC_OUT_CB16 | CPN_TC_OUT_KG)) && (ldc & 3)) ||
      (o_img && (out_div < 1 || out_kchunks < 1))) {
    cpn_set_error("gemm_tc: bad argument (layer=%d M=%d lda=%d ldc=%d mode=%d)", layer, M, lda, ldc, mode);
    return CPN_ERR_ARG;
  }
  if (M == 0) return CPN_OK;
  const TcLayer& L = kLayers[layer];
  const unsigned char* tcw = reinterpret_cast<const unsigned char*>(packed) + cpn_packed_fp32_floats() * sizeof(float);
  GemmArgs g;
  g.A = A;
  g.lda = lda;
  g.kreal = (L.in + 7) / 8 * 8;   // columns of A that exist (835 -> 840: the zero pad up to lda = 848 is read)
  if (!a_img && g.kreal > lda) {
    cpn_set_error("gemm_tc: lda=%d smaller than the layer's K=%d rounded up to 8", lda, L.in);
    return CPN_ERR_ARG;
  }
  g.M = M;
  g.C = C;
  g.ldc = ldc;
  g.N = L.out;
  g.relu = relu;
  g.out_div = out_div;
  g.out_kchunks = out_kchunks;
  g.f8 = (mode & CPN_TC_F16X3) ? 0 : 1;
  g.out_kind = (mode & CPN_TC_OUT_KG) ? 4 : ((mode & CPN_TC_OUT_ROWDOT) ? 3 : ((mode & CPN_TC_OUT_CB16) ? 2 : 0));
  g.C2 = c2;
  g.a_chunk = (mode & CPN_TC_A_IMAGE3) ? ACT_X8 : ACT_CHUNK_BYTES;       // 12288: [fp16 head 8 KB | remainder plane 4 KB]
  g.out_chunk = (mode & CPN_TC_OUT_IMAGE3) ? ACT_X8 : ACT_CHUNK_BYTES;
  {
    static int w3_env = -1, ws_env = -1;
    if (w3_env < 0) {   // CPN_TC_W3=1: derive the e4m3(w_hi) weight plane on chip next to a compact activation image (measured slower)
      const char* e = getenv("CPN_TC_W3");
      w3_env = (e && atoi(e) != 0) ? 1 : 0;
      e = getenv("CPN_TC_WS");   // CPN_TC_WS=1: weight-stationary MMAs where the N tile allows them (A/B runs)
      ws_env = (e && atoi(e) != 0) ? 1 : 0;
    }
    g.w3 = w3_env;
    g.ws = (ws_env || (mode & CPN_TC_WS)) && g.f8 && (L.nt == 64 || L.nt == 128 || L.nt == 256);
  }
  if ((mode & (CPN_TC_A_IMAGE3 | CPN_TC_OUT_IMAGE3)) &&
      (!g.f8 || (mode & (CPN_TC_CLUSTER | CPN_TC_NO_PERSIST | CPN_TC_PAIR | CPN_TC_PPAIR)) || ((mode & CPN_TC_A_IMAGE3) && !a_img) ||
       ((mode & CPN_TC_OUT_IMAGE3) && (!o_img || !a_img)))) {
    cpn_set_error("gemm_tc: compact operand images need the fp16 + fp8 scheme and the persistent single-CTA kernel");
    return CPN_ERR_ARG;
  }
  g.out_mul = 1.f;
  g.dbg = g_tc_dbg;
  g.dbg_cap = g_tc_dbg_cap;
  g.dbg_skip = 0;
  if (g_tc_dbg) {   // cost attribution of the epilogue, trace runs only
    const char* e = getenv("CPN_TC_DBG_SKIP");
    if (e) g.dbg_skip = atoi(e);
  }
  g.dotv = dotv;
  g.dot_div = dot_div;
  g.dot_rowadd = dot_rowadd;
  g.dot_blocks = dot_blocks > 0 ? dot_blocks : L.out / 16;
  g.dot_block0 = dot_block0;
  if (g.out_kind && (o_img || (g.out_kind == 3 && (L.out != L.nt || !dotv)) ||
                     (g.out_kind == 4 && (L.out != L.nt || L.nt != 2 * CPN_HIDDEN || !dotv || !c2)))) {
    cpn_set_error("gemm_tc: CB16 / row-dot outputs are fp32; the row-dot needs a single-N-tile layer (two for the KG form)");
    return CPN_ERR_ARG;
  }
  g.wtiles = tcw + layer_offset(layer, g.f8);
  g.bias = reinterpret_cast<const float*>(packed) + L.bias;
  g.inv_scale = reinterpret_cast<const float*>(tcw) + layer;
  g.kchunks = L.kpad / BK;
  g.NT = L.nt;
  g.idesc = make_idesc_f16(128, L.nt);
  dim3 grid(L.out / L.nt, (M + BM - 1) / BM);
  const int ntiles = L.out / L.nt;
  if (a_img && g.f8 && (mode & CPN_TC_PAIR) && (M % (2 * BM)) == 0) {
    // CTA pairs (cta_group::2), opt-in: measured 20-28 % slower than independent CTAs; needs whole
    // 512-row pairs, other shapes take the single-CTA kernel below
    g.wtiles = tcw + layer_offset(layer, 2);
    void (*pk)(GemmArgs) = o_img ? gemm_tc_pair_kernel<true> : gemm_tc_pair_kernel<false>;
    CPN_CHECK_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM_BYTES));
    cudaLaunchConfig_t pc = {};
    pc.gridDim = dim3(2, ntiles, M / (2 * BM));
    pc.blockDim = dim3(NUM_THREADS);
    pc.dynamicSmemBytes = PSMEM_BYTES;
    pc.stream = st;
    cudaLaunchAttribute pa[1];
    pa[0].id = cudaLaunchAttributeClusterDimension;
    pa[0].val.clusterDim.x = 2;
    pa[0].val.clusterDim.y = 1;
    pa[0].val.clusterDim.z = 1;
    pc.attrs = pa;
    pc.numAttrs = 1;
    CPN_CHECK_CUDA(cudaLaunchKernelEx(&pc, pk, g));
    return CPN_OK;
  }
  static int n_sm = 0;   // SMs of the current device (all devices of a box are the same part)
  if (n_sm == 0) {
    int dev = 0;
    CPN_CHECK_CUDA(cudaGetDevice(&dev));
    CPN_CHECK_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  }
  static int ppair_env = -1;   // CPN_TC_PPAIR=1: persistent CTA pairs wherever they apply (A/B runs)
  if (ppair_env < 0) {
    const char* e = getenv("CPN_TC_PPAIR");
    ppair_env = (e && atoi(e) != 0) ? 1 : 0;
  }
  if (ppair_env && !(mode & (CPN_TC_A_IMAGE3 | CPN_TC_OUT_IMAGE3))) mode |= CPN_TC_PPAIR;
  if (a_img && g.f8 && (mode & CPN_TC_PPAIR) && (M % (2 * BM)) == 0 && (g.out_kind == 0 || g.out_kind == 2) &&
      !(mode & (CPN_TC_CLUSTER | CPN_TC_NO_PERSIST))) {
    // persistent CTA pairs: one cluster of two per SM pair walks the (512-row tile, N tile) list
    g.wtiles = tcw + layer_offset(layer, 2);
    void (*pk)(GemmArgs, int, int) = o_img ? gemm_tc_ppair_kernel<true> : gemm_tc_ppair_kernel<false>;
    CPN_CHECK_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, PSMEM_BYTES));
    const int nptiles = ntiles * (M / (2 * BM));
    const int nclusters = nptiles < n_sm / 2 ? nptiles : n_sm / 2;
    cudaLaunchConfig_t pc = {};
    pc.gridDim = dim3(2 * nclusters);
    pc.blockDim = dim3(PP_THREADS);
    pc.dynamicSmemBytes = PSMEM_BYTES;
    pc.stream = st;
    cudaLaunchAttribute pa[1];
    pa[0].id = cudaLaunchAttributeClusterDimension;
    pa[0].val.clusterDim.x = 2;
    pa[0].val.clusterDim.y = 1;
    pa[0].val.clusterDim.z = 1;
    pc.attrs = pa;
    pc.numAttrs = 1;
    CPN_CHECK_CUDA(cudaLaunchKernelEx(&pc, pk, g, ntiles, nptiles));
    return CPN_OK;
  }
  if (a_img && !(mode & (CPN_TC_CLUSTER | CPN_TC_NO_PERSIST))) {
    static int epi_warps = 0;   // CPN_TC_EPI_WARPS = 16 | 24 (A/B runs)
    if (epi_warps == 0) {
      const char* e = getenv("CPN_TC_EPI_WARPS");
      epi_warps = (e && atoi(e) == 24) ? 24 : 16;
    }
    const bool dots = g.out_kind == 3 || g.out_kind == 4;   // their prefetching epilogue is written for two column parts
    static int cl2 = -1;   // CPN_TC_CLUSTER2=1: clusters of two CTAs share the activation stages by multicast (A/B runs)
    if (cl2 < 0) {
      const char* e = getenv("CPN_TC_CLUSTER2");
      cl2 = (e && atoi(e) != 0) ? 1 : 0;
    }
    const int total = ntiles * (int)grid.y;
    if (cl2 && !dots && (ntiles % 2) == 0 && total >= 2 && !(mode & (CPN_TC_A_IMAGE3 | CPN_TC_OUT_IMAGE3))) {
      void (*pk2)(GemmArgs, int, int) = o_img ? gemm_tc_persist_kernel<true, 16, 2> : gemm_tc_persist_kernel<false, 16, 2>;
      CPN_CHECK_CUDA(cudaFuncSetAttribute(pk2, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
      cudaLaunchConfig_t pc = {};
      int ctas = total < n_sm ? total : n_sm;
      ctas &= ~1;
      pc.gridDim = dim3(ctas);
      pc.blockDim = dim3((2 + 16) * 32);
      pc.dynamicSmemBytes = SMEM_BYTES;
      pc.stream = st;
      cudaLaunchAttribute pa[1];
      pa[0].id = cudaLaunchAttributeClusterDimension;
      pa[0].val.clusterDim.x = 2;
      pa[0].val.clusterDim.y = 1;
      pa[0].val.clusterDim.z = 1;
      pc.attrs = pa;
      pc.numAttrs = 1;
      CPN_CHECK_CUDA(cudaLaunchKernelEx(&pc, pk2, g, ntiles, total));
      return CPN_OK;
    }
    static int persist2 = -1;   // CPN_TC_PERSIST2=1: the sub-tile pipelined kernel (measured slower: 1.39 vs 1.33 ms, comment above it)
    if (persist2 < 0) {
      const char* e = getenv("CPN_TC_PERSIST2");
      persist2 = (e && atoi(e) != 0) ? 1 : 0;
    }
    if ((persist2 || (mode & CPN_TC_PERSIST2)) && epi_warps != 24) {
      void (*pk2)(GemmArgs, int, int) = o_img ? gemm_tc_persist2_kernel<true> : gemm_tc_persist2_kernel<false>;
      CPN_CHECK_CUDA(cudaFuncSetAttribute(pk2, cudaFuncAttributeMaxDynamicSharedMemorySize, P2_SMEM_BYTES));
      pk2<<<total < n_sm ? total : n_sm, P2_THREADS, P2_SMEM_BYTES, st>>>(g, ntiles, total);
      CPN_CHECK_LAUNCH("gemm_tc_persist2_kernel");
      return CPN_OK;
    }
    void (*pk)(GemmArgs, int, int) =
        (epi_warps == 24 && !dots) ? (o_img ? gemm_tc_persist_kernel<true, 24, 1> : gemm_tc_persist_kernel<false, 24, 1>)
                        : (o_img ? gemm_tc_persist_kernel<true, 16, 1> : gemm_tc_persist_kernel<false, 16, 1>);
    CPN_CHECK_CUDA(cudaFuncSetAttribute(pk, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
    pk<<<total < n_sm ? total : n_sm, (2 + ((epi_warps == 24 && !dots) ? 24 : 16)) * 32, SMEM_BYTES, st>>>(g, ntiles, total);
    CPN_CHECK_LAUNCH("gemm_tc_persist_kernel");
    return CPN_OK;
  }
  const int cluster = (a_img && (mode & CPN_TC_CLUSTER)) ? ntiles : 1;   // 4, 2 or 1; opt-in: measured slower
  void (*kern)(GemmArgs);
  if (cluster == 4) kern = o_img ? gemm_tc_kernel<true, true, 4> : gemm_tc_kernel<true, false, 4>;
  else if (cluster == 2) kern = o_img ? gemm_tc_kernel<true, true, 2> : gemm_tc_kernel<true, false, 2>;
  else kern = a_img ? (o_img ? gemm_tc_kernel<true, true, 1> : gemm_tc_kernel<true, false, 1>)
                    : (o_img ? gemm_tc_kernel<false, true, 1> : gemm_tc_kernel<false, false, 1>);
  CPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NUM_THREADS);
  cfg.dynamicSmemBytes = SMEM_BYTES;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  CPN_CHECK_CUDA(cudaLaunchKernelEx(&cfg, kern, g));
  return CPN_OK;
}

extern "C" int cpn_gemm_tc(const void* packed, int layer, const void* A, int lda, void* C, int ldc, int M, int relu,
                           int mode, int out_div, int out_kchunks, void* stream) {
  return launch_gemm_tc(packed, layer, A, lda, C, ldc, M, relu, mode & ~(CPN_TC_OUT_ROWDOT | CPN_TC_OUT_KG), out_div, out_kchunks,
                        (cudaStream_t)stream, nullptr, 1.f, nullptr, 0, 0);
}

extern "C" int cpn_gemm_tc_rowdot(const void* packed, int layer, const void* A, int lda, const float* dotv_cb16, float* out,
                                  int M, int relu, int mode, float div, void* stream) {
  return launch_gemm_tc(packed, layer, A, lda, out, 0, M, relu, (mode & ~CPN_TC_OUT_KG) | CPN_TC_OUT_ROWDOT, 1, 1,
                        (cudaStream_t)stream, dotv_cb16, div, nullptr, 0, 0);
}

// layer 9 with the 16 -> 128 ReLU layer in front of it computed by the producer warps (query_embed, CoPoNeRF.py:446):
// Qm (CB16, 16 blocks) = [WM1 ; WM2] relu(Wq x16 + bq) + [BM1 ; BM2], s1 / s2 = <relu(.), WS> + CS
int launch_gemm_tc_mlp16(const void* packed, const float* x16, const float* wt, const float* bias, const float* sd1, float* s1,
                         const float* sd2, float* s2, float* qm_cb16, int M, int mode, cudaStream_t st) {
  if (!packed || !x16 || !wt || !bias || !qm_cb16 || M < 0) {
    cpn_set_error("gemm_tc_mlp16: bad argument");
    return CPN_ERR_ARG;
  }
  if (M == 0) return CPN_OK;
  const TcLayer& L = kLayers[9];
  const unsigned char* tcw = reinterpret_cast<const unsigned char*>(packed) + cpn_packed_fp32_floats() * sizeof(float);
  GemmArgs g = {};
  g.M = M;
  g.C = qm_cb16;
  g.N = L.out;
  g.out_div = g.out_kchunks = 1;
  g.f8 = (mode & CPN_TC_F16X3) ? 0 : 1;
  g.out_kind = 2;
  g.wtiles = tcw + layer_offset(9, g.f8);
  g.bias = reinterpret_cast<const float*>(packed) + L.bias;
  g.inv_scale = reinterpret_cast<const float*>(tcw) + 9;
  g.kchunks = L.kpad / BK;
  g.NT = L.nt;
  g.idesc = make_idesc_f16(128, L.nt);
  g.dot_div = 1.f;
  g.a_chunk = g.out_chunk = ACT_CHUNK_BYTES;
  g.out_mul = 1.f;
  g.mlp_x = x16; g.mlp_wt = wt; g.mlp_b = bias; g.mlp_sd1 = sd1; g.mlp_sd2 = sd2; g.mlp_s1 = s1; g.mlp_s2 = s2;
  if ((s1 && !sd1) || (s2 && !sd2) || L.out != L.nt) {
    cpn_set_error("gemm_tc_mlp16: inconsistent arguments");
    return CPN_ERR_ARG;
  }
  void (*kern)(GemmArgs) = gemm_tc_kernel<false, false, 1, true>;
  CPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  kern<<<dim3(1, (M + BM - 1) / BM), NUM_THREADS, SMEM_BYTES, st>>>(g);
  CPN_CHECK_LAUNCH("gemm_tc_kernel (mlp16)");
  return CPN_OK;
}

extern "C" int cpn_gemm_tc_kg(const void* packed, const void* h1_image, const float* dotv_cb16, int dot_blocks,
                              const float* rowadd, float div, float* logits, float* gh, int M, int mode, void* stream) {
  return launch_gemm_tc(packed, 10, h1_image, 0, logits, 0, M, 1, (mode & (CPN_TC_F16X3 | CPN_TC_A_IMAGE3 | CPN_TC_NO_PERSIST | CPN_TC_PERSIST2 | CPN_TC_WS)) | CPN_TC_A_IMAGE | CPN_TC_OUT_KG, 1, 1,
                        (cudaStream_t)stream, dotv_cb16, div, rowadd, dot_blocks, 0, gh);
}


// ---- generic Linear on the tensor-core kernel (token layers of the cost aggregation, models/aggregation.py:269-340) --------
// y[M, N] = act(x[M, K] W[N, K]^T + b) for any weight with N a multiple of 128 and K a multiple of 8: the weight is packed once
// (cpn_linear_tc_pack: per-layer power-of-two scale, split tiles of both schemes), activations stay fp32 row-major (producer
// warps split them inside the kernel).
namespace {
size_t lin_bias_bytes(int N) { return ((size_t)N * 4 + 255) / 256 * 256; }
size_t lin_tiles_bytes(int N, int K) { return (size_t)N * ((K + BK - 1) / BK * BK) * 4; }
}  // namespace

extern "C" size_t cpn_linear_tc_packed_bytes(int N, int K) {
  if (N <= 0 || K <= 0 || (N % 128) != 0) return 0;
  return TC_HEADER_BYTES + lin_bias_bytes(N) + 2 * lin_tiles_bytes(N, K);
}

extern "C" int cpn_linear_tc_pack(const float* w, int N, int K, void* packed, void* stream) {
  if (!w || !packed || N <= 0 || K <= 0 || (N % 128) != 0) {
    cpn_set_error("cpn_linear_tc_pack: N must be a positive multiple of 128 (N=%d K=%d)", N, K);
    return CPN_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  unsigned char* dst = reinterpret_cast<unsigned char*>(packed);
  CPN_CHECK_CUDA(cudaMemsetAsync(dst, 0, cpn_linear_tc_packed_bytes(N, K), st));
  unsigned int* absmax = reinterpret_cast<unsigned int*>(dst) + 32;
  const int kpad = (K + BK - 1) / BK * BK;
  absmax_kernel<<<64, 256, 0, st>>>(w, (size_t)N * K, absmax);
  CPN_CHECK_LAUNCH("absmax_kernel");
  unsigned char* t0 = dst + TC_HEADER_BYTES + lin_bias_bytes(N);
  const size_t total = (size_t)N * kpad;
  pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(w, N, K, kpad, 128, absmax, reinterpret_cast<__half*>(t0),
                                                                  t0 + lin_tiles_bytes(N, K), nullptr,
                                                                  reinterpret_cast<float*>(dst), 0);
  CPN_CHECK_LAUNCH("pack_tc_kernel");
  return CPN_OK;
}

extern "C" int cpn_linear_tc(const void* packed, int N, int K, const float* x, int ldx, const float* bias, float* y, int ldy,
                             int M, int act, int mode, void* stream) {
  return launch_linear_tc(packed, N, K, x, ldx, bias, y, ldy, M, act, mode, 1.f, (cudaStream_t)stream);
}

int launch_linear_tc(const void* packed, int N, int K, const float* x, int ldx, const float* bias, float* y, int ldy, int M,
                     int act, int mode, float out_mul, cudaStream_t stream) {
  if (!packed || !x || !y || N <= 0 || K <= 0 || (N % 128) != 0 || (K & 7) || (ldx & 3) || (ldy & 3) || ldx < K || M < 0 ||
      act < 0 || act > 2) {
    cpn_set_error("cpn_linear_tc: bad argument (N=%d K=%d ldx=%d ldy=%d M=%d act=%d): N %% 128 == 0, K %% 8 == 0", N, K, ldx, ldy,
                  M, act);
    return CPN_ERR_ARG;
  }
  if (M == 0) return CPN_OK;
  const unsigned char* p = reinterpret_cast<const unsigned char*>(packed);
  GemmArgs g = {};
  g.A = x;
  g.lda = ldx;
  g.kreal = K;
  g.M = M;
  g.C = y;
  g.ldc = ldy;
  g.N = N;
  g.relu = act;
  g.out_div = 1;
  g.out_kchunks = 1;
  g.f8 = (mode & CPN_TC_F16X3) ? 0 : 1;
  g.out_kind = 0;
  g.wtiles = p + TC_HEADER_BYTES + lin_bias_bytes(N) + (g.f8 ? lin_tiles_bytes(N, K) : 0);
  g.bias = bias ? bias : reinterpret_cast<const float*>(p + TC_HEADER_BYTES);   // zeros
  g.inv_scale = reinterpret_cast<const float*>(p);
  g.kchunks = (K + BK - 1) / BK;
  g.NT = 128;
  g.idesc = make_idesc_f16(128, 128);
  g.dot_div = 1.f;
  g.a_chunk = g.out_chunk = ACT_CHUNK_BYTES;
  g.out_mul = out_mul;
  void (*kern)(GemmArgs) = gemm_tc_kernel<false, false, 1>;
  CPN_CHECK_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  dim3 grid(N / 128, (M + BM - 1) / BM);
  kern<<<grid, NUM_THREADS, SMEM_BYTES, stream>>>(g);
  CPN_CHECK_LAUNCH("gemm_tc_kernel (linear)");
  return CPN_OK;
}
