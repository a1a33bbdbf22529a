// Epipolar-line feature gather: F.grid_sample(bilinear, align_corners=False) of the four
// channels-last feature maps, primary ('border' padding, models/CoPoNeRF.py:312) and secondary
// ('zeros' padding at the reprojected coordinates, models/CoPoNeRF.py:370).
#include <stdlib.h>
#include "cpn_common.cuh"
#include "tc_common.cuh"

namespace {

struct Taps {
  int off[4];   // element offset of each tap's channel vector inside one image of the level (or -1)
  float w[4];   // bilinear weights in PyTorch's nw, ne, sw, se order
};

// grid_sampler_compute_source_index + bilinear weights (ATen/native/GridSampler.h)
__device__ __forceinline__ Taps make_taps(float gx, float gy, int h, int w, int C, bool border) {
  float ix = ((gx + 1.f) * (float)w - 1.f) / 2.f;
  float iy = ((gy + 1.f) * (float)h - 1.f) / 2.f;
  if (border) {
    ix = fminf((float)(w - 1), fmaxf(ix, 0.f));
    iy = fminf((float)(h - 1), fmaxf(iy, 0.f));
  }
  float x0 = floorf(ix), y0 = floorf(iy);
  float x1 = x0 + 1.f, y1 = y0 + 1.f;
  Taps t;
  t.w[0] = (x1 - ix) * (y1 - iy);
  t.w[1] = (ix - x0) * (y1 - iy);
  t.w[2] = (x1 - ix) * (iy - y0);
  t.w[3] = (ix - x0) * (iy - y0);
  float xs[4] = {x0, x1, x0, x1}, ys[4] = {y0, y0, y1, y1};
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    bool in = xs[k] >= 0.f && xs[k] <= (float)(w - 1) && ys[k] >= 0.f && ys[k] <= (float)(h - 1);
    t.off[k] = in ? ((int)ys[k] * w + (int)xs[k]) * C : -1;
  }
  return t;
}

// One warp per (row, branch). Row = ((b*nr + n)*2 + v)*S + s; branch 0 reads view v at the sample
// position, branch 1 reads view 1-v at the reprojected position.
__global__ void __launch_bounds__(256) gather_kernel(cpn_render_args a, int nr, const float* __restrict__ rowaux,
                                                     float* __restrict__ A) {
  long long wid = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  long long nrows = (long long)a.B * nr * 2 * a.S;
  if (wid >= nrows * 2) return;
  long long row = wid >> 1;
  int branch = (int)(wid & 1);
  int v = (int)((row / a.S) & 1);
  int b = (int)(row / ((long long)2 * a.S * nr));
  int img = b * 2 + (branch ? 1 - v : v);
  const float* ra = rowaux + (size_t)row * CPN_ROWAUX;
  float gx = ra[branch * 2 + 0], gy = ra[branch * 2 + 1];
  float* out = A + enc_row((size_t)row, branch) * CPN_KA;
  int col = 0;
#pragma unroll
  for (int l = 0; l < CPN_N_LEVELS; ++l) {
    int h = a.feat_h[l], w = a.feat_w[l], C = a.feat_c[l];
    Taps t = make_taps(gx, gy, h, w, C, branch == 0);
    const float* base = a.feat[l] + (size_t)img * h * w * C;
    for (int c = lane * 4; c < C; c += 128) {
      float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        if (t.off[k] >= 0) {
          float4 f = __ldg(reinterpret_cast<const float4*>(base + t.off[k] + c));
          acc.x += f.x * t.w[k];
          acc.y += f.y * t.w[k];
          acc.z += f.z * t.w[k];
          acc.w += f.w * t.w[k];
        }
      }
      *reinterpret_cast<float4*>(out + col + c) = acc;
    }
    col += C;
  }
}

// Operand-image form of the same gather. CTA = 8 warps = 8 consecutive sample rows of one branch; each warp blends
// its row (lanes along channels: every tap is a coalesced 512-byte read), splits it into fp16 hi/lo and parks it in
// shared memory; the CTA then writes the image, where the same 8-channel group of 8 consecutive rows is 128
// contiguous bytes (full-sector, coalesced stores).
constexpr int GI_GROUPS = 4;   // 8-row groups per CTA
constexpr int GI_ROWS = 8, GI_PITCH = CPN_FEAT_DIM + 24;  // halves per staged row: 1712 B = 428 words, 428 % 32 = 12 -> the 8 rows' 16-byte reads hit distinct banks
constexpr int GI_B8 = CPN_FEAT_DIM + 16;                  // bytes per 8-bit plane of a staged row

// make_taps with out-of-range taps turned into (offset 0, weight 0): the blend needs no branches
__device__ __forceinline__ Taps make_taps_safe(float gx, float gy, int h, int w, int C, bool border) {
  Taps t = make_taps(gx, gy, h, w, C, border);
#pragma unroll
  for (int k = 0; k < 4; ++k)
    if (t.off[k] < 0) {
      t.off[k] = 0;
      t.w[k] = 0.f;
    }
  return t;
}

// Bilinear taps of every (sample row, branch, level), one thread each: 4 element offsets + 4 weights = 32 bytes.
// In gather_image_kernel the 32 lanes of a warp share a row, so computing the taps there repeats the same ~100
// instructions per level in every lane (40 % of that issue-bound kernel's instructions); here they are computed once
// and the gather fetches them with two uniform 16-byte loads per level. Same arithmetic, same bits.
__global__ void __launch_bounds__(256) taps_kernel(cpn_render_args a, int nr, const float* __restrict__ rowaux,
                                                   int4* __restrict__ taps) {
  const unsigned nrows = (unsigned)(a.B * nr * 2 * a.S);
  const unsigned i = blockIdx.x * 256 + threadIdx.x;
  if (i >= nrows * 2 * CPN_N_LEVELS) return;
  const int l = i % CPN_N_LEVELS, branch = (i / CPN_N_LEVELS) & 1;
  const unsigned row = i / (2 * CPN_N_LEVELS);
  const float2 g = *reinterpret_cast<const float2*>(rowaux + (size_t)row * CPN_ROWAUX + branch * 2);
  int h = a.feat_h[0], w = a.feat_w[0], C = a.feat_c[0];
#pragma unroll
  for (int k = 1; k < CPN_N_LEVELS; ++k)
    if (l == k) {
      h = a.feat_h[k];
      w = a.feat_w[k];
      C = a.feat_c[k];
    }
  const Taps t = make_taps_safe(g.x, g.y, h, w, C, branch == 0);
  taps[(size_t)i * 2] = make_int4(t.off[0], t.off[1], t.off[2], t.off[3]);
  taps[(size_t)i * 2 + 1] = make_int4(__float_as_int(t.w[0]), __float_as_int(t.w[1]), __float_as_int(t.w[2]), __float_as_int(t.w[3]));
}

// X8 = false: compact image (12 KB blocks, no value plane: the consuming GEMM derives e5m2(head) on chip)
template <bool F8, bool X8>
__global__ void __launch_bounds__(256) gather_image_kernel(cpn_render_args a, int nr, const int4* __restrict__ taps,
                                                           unsigned char* __restrict__ img) {
  // f16x3: [hi | lo][row][channel] fp16. f8: plane 0 = fp16 hi; plane 1 holds the two byte planes back to back,
  // e5m2((x - hi) * 2^10) in bytes [0, 832) and e5m2(hi) in bytes [848, 1680) of each row (tc_common.cuh).
  __shared__ __align__(16) __half sh[2][GI_ROWS][GI_PITCH];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int branch = blockIdx.y;
  const unsigned nrows = (unsigned)(a.B * nr * 2 * a.S);       // < 2^31 (checked by the launcher)
  // a CTA walks GI_GROUPS groups of 8 rows: the per-thread index arithmetic of the copy-out (and the level geometry) is set up
  // once per CTA instead of once per 8 rows (a third of the kernel's instructions was integer address work, profiles/)
#pragma unroll 1
  for (int grp = 0; grp < GI_GROUPS; ++grp) {
  const unsigned row0 = (blockIdx.x * GI_GROUPS + grp) * GI_ROWS, row = row0 + warp;
  if (row0 >= nrows) break;
  if (row < nrows) {
    const unsigned S = (unsigned)a.S;
    const int v = (int)((row / S) & 1u);
    const int b = (int)(row / (2u * S * (unsigned)nr));
    const int im = b * 2 + (branch ? 1 - v : v);
    const int4* trow = taps + ((size_t)row * 2 + branch) * CPN_N_LEVELS * 2;
    __half* sh_hi = &sh[0][warp][0];
    unsigned char* sh_b = reinterpret_cast<unsigned char*>(&sh[1][warp][0]);
    int col = 0;
#pragma unroll
    for (int l = 0; l < CPN_N_LEVELS; ++l) {
      const int h = a.feat_h[l], w = a.feat_w[l], C = a.feat_c[l];
      Taps t;   // precomputed by taps_kernel; the address is the same in every lane (one broadcast transaction)
      {
        const int4 o = __ldg(trow + 2 * l), wv = __ldg(trow + 2 * l + 1);
        t.off[0] = o.x; t.off[1] = o.y; t.off[2] = o.z; t.off[3] = o.w;
        t.w[0] = __int_as_float(wv.x); t.w[1] = __int_as_float(wv.y); t.w[2] = __int_as_float(wv.z); t.w[3] = __int_as_float(wv.w);
      }
      const float* base = a.feat[l] + (size_t)im * h * w * C + lane * 4;
      for (int c = lane * 4; c < C; c += 128, base += 128) {
        const float4 f0 = __ldg(reinterpret_cast<const float4*>(base + t.off[0]));
        const float4 f1 = __ldg(reinterpret_cast<const float4*>(base + t.off[1]));
        const float4 f2 = __ldg(reinterpret_cast<const float4*>(base + t.off[2]));
        const float4 f3 = __ldg(reinterpret_cast<const float4*>(base + t.off[3]));
        float4 acc;   // same order as the fp32 kernel: ((f0 w0 + f1 w1) + f2 w2) + f3 w3
        acc.x = fmaf(f3.x, t.w[3], fmaf(f2.x, t.w[2], fmaf(f1.x, t.w[1], f0.x * t.w[0])));
        acc.y = fmaf(f3.y, t.w[3], fmaf(f2.y, t.w[2], fmaf(f1.y, t.w[1], f0.y * t.w[0])));
        acc.z = fmaf(f3.z, t.w[3], fmaf(f2.z, t.w[2], fmaf(f1.z, t.w[1], f0.z * t.w[0])));
        acc.w = fmaf(f3.w, t.w[3], fmaf(f2.w, t.w[2], fmaf(f1.w, t.w[1], f0.w * t.w[0])));
        if (F8) {
          uint2 hi;
          uint32_t l8, x8;
          tc::split4_f8(acc, hi, l8, x8);
          *reinterpret_cast<uint2*>(sh_hi + col + c) = hi;
          *reinterpret_cast<uint32_t*>(sh_b + col + c) = l8;
          if (X8) *reinterpret_cast<uint32_t*>(sh_b + GI_B8 + col + c) = x8;
        } else {
          uint2 hi, lo;
          tc::split2(acc.x, acc.y, hi.x, lo.x);
          tc::split2(acc.z, acc.w, hi.y, lo.y);
          *reinterpret_cast<uint2*>(sh_hi + col + c) = hi;
          *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(sh_b) + col + c) = lo;
        }
      }
      col += C;
    }
  }
  __syncthreads();
  // 208 16-byte groups per row (104 fp16-hi groups of 8 channels + 2 x 52 byte-plane groups of 16 channels, or
  // 104 + 104 fp16 groups); thread -> (group, row): 8 threads write 128 contiguous bytes
  const int rr = threadIdx.x & 7;
  if (row0 + rr < nrows) {
    constexpr int NG8 = CPN_FEAT_DIM / 8, NG16 = CPN_FEAT_DIM / 16;
    constexpr int CHUNK = (F8 && !X8) ? ACT_X8 : ACT_CHUNK_BYTES;
    unsigned char* tile = img + ((size_t)(row0 >> 7) * 2 + branch) * ((size_t)(CPN_KA_IMG / ACT_BK) * CHUNK) +
                          ((row0 & 127) + rr) * 16;
    const unsigned char* s_hi = reinterpret_cast<const unsigned char*>(&sh[0][rr][0]);
    const unsigned char* s_b = reinterpret_cast<const unsigned char*>(&sh[1][rr][0]);
#pragma unroll
    for (int it = 0; it < (NG8 + 31) / 32; ++it) {   // plane 0: fp16 hi
      const int gi = it * 32 + (threadIdx.x >> 3);
      if (gi < NG8) {
        const int k = gi * 8;
        *reinterpret_cast<uint4*>(tile + (k >> 5) * CHUNK + ((k & 31) >> 3) * 2048) =
            *reinterpret_cast<const uint4*>(s_hi + gi * 16);
      }
    }
    if (F8) {
      constexpr int NPL = X8 ? 2 : 1;
#pragma unroll
      for (int it = 0; it < (NPL * NG16 + 31) / 32; ++it) {
        const int item = it * 32 + (threadIdx.x >> 3);
        if (item < NPL * NG16) {
          const int pl = item >= NG16, gi = item - pl * NG16, k = gi * 16;
          *reinterpret_cast<uint4*>(tile + (k >> 5) * CHUNK + (pl ? ACT_X8 : ACT_LO8) + ((k & 31) >> 4) * 2048) =
              *reinterpret_cast<const uint4*>(s_b + pl * GI_B8 + gi * 16);
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < (NG8 + 31) / 32; ++it) {
        const int gi = it * 32 + (threadIdx.x >> 3);
        if (gi < NG8) {
          const int k = gi * 8;
          *reinterpret_cast<uint4*>(tile + (k >> 5) * CHUNK + ACT_LO + ((k & 31) >> 3) * 2048) =
              *reinterpret_cast<const uint4*>(s_b + gi * 16);
        }
      }
    }
  }
  __syncthreads();   // the staged rows are free for the next group
  }
}

// Sequential-row variant of the operand-image gather (experiment, opt-in with CPN_GATHER_SEQ=1: it removes 58 % of the tap
// loads and measured 590 us per chunk against 595 us -- the kernel is bound by instruction issue of the blend + split +
// staging work, not by the tap loads). Consecutive samples of an epipolar line fall into the same texel cell
// several times in a row at the coarse levels (a 16 x 16 map: 4-8 samples per cell; 32 x 32: 2-4; 64 x 64: 1-2), and the
// kernel above re-fetches the four taps for every sample: 112 of its ~190 L1 wavefronts per (row, branch) are tap loads
// (the L1 data pipe was its limit, 76 % of the LSU-wavefront peak). Here a CTA owns 16 consecutive rows of one branch and a
// WARP owns a (level, half of the channels) unit for all of them: it walks the rows in order and keeps the four tap
// vectors in registers, reloading one only when its texel changes. Same taps, same blend order, same bits. Level 3
// (64 channels at full resolution: a new cell for every sample) is fetched directly, two rows per warp pass.
constexpr int GS_ROWS = 16;
constexpr int GS_SMEM = 2 * GS_ROWS * GI_PITCH * (int)sizeof(__half);

template <bool F8>
__device__ __forceinline__ void gs_store(float4 acc, __half* sh_hi, unsigned char* sh_b, int col) {
  if (F8) {
    uint2 hi;
    uint32_t l8, x8;
    tc::split4_f8(acc, hi, l8, x8);
    *reinterpret_cast<uint2*>(sh_hi + col) = hi;
    *reinterpret_cast<uint32_t*>(sh_b + col) = l8;
    *reinterpret_cast<uint32_t*>(sh_b + GI_B8 + col) = x8;
  } else {
    uint2 hi, lo;
    tc::split2(acc.x, acc.y, hi.x, lo.x);
    tc::split2(acc.z, acc.w, hi.y, lo.y);
    *reinterpret_cast<uint2*>(sh_hi + col) = hi;
    *reinterpret_cast<uint2*>(reinterpret_cast<__half*>(sh_b) + col) = lo;
  }
}
__device__ __forceinline__ float4 gs_blend(const float4& f0, const float4& f1, const float4& f2, const float4& f3, const int4& wv) {
  const float w0 = __int_as_float(wv.x), w1 = __int_as_float(wv.y), w2 = __int_as_float(wv.z), w3 = __int_as_float(wv.w);
  float4 acc;   // same order as the fp32 kernel: ((f0 w0 + f1 w1) + f2 w2) + f3 w3
  acc.x = fmaf(f3.x, w3, fmaf(f2.x, w2, fmaf(f1.x, w1, f0.x * w0)));
  acc.y = fmaf(f3.y, w3, fmaf(f2.y, w2, fmaf(f1.y, w1, f0.y * w0)));
  acc.z = fmaf(f3.z, w3, fmaf(f2.z, w2, fmaf(f1.z, w1, f0.z * w0)));
  acc.w = fmaf(f3.w, w3, fmaf(f2.w, w2, fmaf(f1.w, w1, f0.w * w0)));
  return acc;
}

template <bool F8>
__global__ void __launch_bounds__(256) gather_image_seq_kernel(cpn_render_args a, int nr, const int4* __restrict__ taps,
                                                               unsigned char* __restrict__ img) {
  extern __shared__ __align__(16) unsigned char gs_smem[];
  __half(*sh)[GS_ROWS][GI_PITCH] = reinterpret_cast<__half(*)[GS_ROWS][GI_PITCH]>(gs_smem);   // [plane][row][channel]
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int branch = blockIdx.y;
  const unsigned nrows = (unsigned)(a.B * nr * 2 * a.S);
  const unsigned row0 = blockIdx.x * GS_ROWS;
  // the 16 rows share (pair, view): S is a multiple of 16 (checked by the launcher)
  const unsigned S = (unsigned)a.S;
  const int v = (int)((row0 / S) & 1u), b = (int)(row0 / (2u * S * (unsigned)nr));
  const int im = b * 2 + (branch ? 1 - v : v);
  if (warp < 6) {
    const int l = warp >> 1, cbase = (warp & 1) * 128;          // levels 0-2 have 256 channels: two warps each
    const int h = a.feat_h[l], w = a.feat_w[l], C = a.feat_c[l];
    int col0 = 0;
    for (int k = 0; k < l; ++k) col0 += a.feat_c[k];
    const int col = col0 + cbase + lane * 4;
    if (cbase + lane * 4 < C) {
      const float* base = a.feat[l] + (size_t)im * h * w * C + cbase + lane * 4;
      int oc0 = -1, oc1 = -1, oc2 = -1, oc3 = -1;
      float4 f0 = make_float4(0.f, 0.f, 0.f, 0.f), f1 = f0, f2 = f0, f3 = f0;
#pragma unroll 4
      for (int r = 0; r < GS_ROWS; ++r) {
        const unsigned row = row0 + r;
        if (row >= nrows) break;
        const int4* trow = taps + (((size_t)row * 2 + branch) * CPN_N_LEVELS + l) * 2;
        const int4 o = __ldg(trow), wv = __ldg(trow + 1);
        if (o.x != oc0) { f0 = __ldg(reinterpret_cast<const float4*>(base + o.x)); oc0 = o.x; }
        if (o.y != oc1) { f1 = __ldg(reinterpret_cast<const float4*>(base + o.y)); oc1 = o.y; }
        if (o.z != oc2) { f2 = __ldg(reinterpret_cast<const float4*>(base + o.z)); oc2 = o.z; }
        if (o.w != oc3) { f3 = __ldg(reinterpret_cast<const float4*>(base + o.w)); oc3 = o.w; }
        gs_store<F8>(gs_blend(f0, f1, f2, f3, wv), &sh[0][r][0], reinterpret_cast<unsigned char*>(&sh[1][r][0]), col);
      }
    }
  } else {
    // level 3: (warp - 6) takes rows 0-7 / 8-15, the two half-warps two rows at a time, lanes along the 64 channels
    const int l = CPN_N_LEVELS - 1;
    const int h = a.feat_h[l], w = a.feat_w[l], C = a.feat_c[l];
    int col0 = 0;
    for (int k = 0; k < l; ++k) col0 += a.feat_c[k];
    const int half = lane >> 4, lc = (lane & 15) * 4;
    for (int c = lc; c < C; c += 64) {
      const float* base = a.feat[l] + (size_t)im * h * w * C + c;
#pragma unroll
      for (int it = 0; it < GS_ROWS / 4; ++it) {
        const int r = (warp - 6) * (GS_ROWS / 2) + it * 2 + half;
        const unsigned row = row0 + r;
        if (row >= nrows) continue;
        const int4* trow = taps + (((size_t)row * 2 + branch) * CPN_N_LEVELS + l) * 2;
        const int4 o = __ldg(trow), wv = __ldg(trow + 1);
        const float4 f0 = __ldg(reinterpret_cast<const float4*>(base + o.x));
        const float4 f1 = __ldg(reinterpret_cast<const float4*>(base + o.y));
        const float4 f2 = __ldg(reinterpret_cast<const float4*>(base + o.z));
        const float4 f3 = __ldg(reinterpret_cast<const float4*>(base + o.w));
        gs_store<F8>(gs_blend(f0, f1, f2, f3, wv), &sh[0][r][0], reinterpret_cast<unsigned char*>(&sh[1][r][0]), col0 + c);
      }
    }
  }
  __syncthreads();
  // 208 16-byte groups per row; thread -> (group, row): 16 threads write 256 contiguous bytes of the image
  const int rr = threadIdx.x & (GS_ROWS - 1);
  if (row0 + rr < nrows) {
    constexpr int NG8 = CPN_FEAT_DIM / 8, NG16 = CPN_FEAT_DIM / 16, GPI = 256 / GS_ROWS;   // groups per pass
    unsigned char* tile = img + ((size_t)(row0 >> 7) * 2 + branch) * ((size_t)(CPN_KA_IMG / ACT_BK) * ACT_CHUNK_BYTES) +
                          ((row0 & 127) + rr) * 16;
    const unsigned char* s_hi = reinterpret_cast<const unsigned char*>(&sh[0][rr][0]);
    const unsigned char* s_b = reinterpret_cast<const unsigned char*>(&sh[1][rr][0]);
#pragma unroll
    for (int it = 0; it < (NG8 + GPI - 1) / GPI; ++it) {   // plane 0: fp16 hi
      const int gi = it * GPI + (threadIdx.x / GS_ROWS);
      if (gi < NG8) {
        const int k = gi * 8;
        *reinterpret_cast<uint4*>(tile + (k >> 5) * ACT_CHUNK_BYTES + ((k & 31) >> 3) * 2048) =
            *reinterpret_cast<const uint4*>(s_hi + gi * 16);
      }
    }
    if (F8) {
#pragma unroll
      for (int it = 0; it < (2 * NG16 + GPI - 1) / GPI; ++it) {
        const int item = it * GPI + (threadIdx.x / GS_ROWS);
        if (item < 2 * NG16) {
          const int pl = item >= NG16, gi = item - pl * NG16, k = gi * 16;
          *reinterpret_cast<uint4*>(tile + (k >> 5) * ACT_CHUNK_BYTES + (pl ? ACT_X8 : ACT_LO8) + ((k & 31) >> 4) * 2048) =
              *reinterpret_cast<const uint4*>(s_b + pl * GI_B8 + gi * 16);
        }
      }
    } else {
#pragma unroll
      for (int it = 0; it < (NG8 + GPI - 1) / GPI; ++it) {
        const int gi = it * GPI + (threadIdx.x / GS_ROWS);
        if (gi < NG8) {
          const int k = gi * 8;
          *reinterpret_cast<uint4*>(tile + (k >> 5) * ACT_CHUNK_BYTES + ACT_LO + ((k & 31) >> 3) * 2048) =
              *reinterpret_cast<const uint4*>(s_b + gi * 16);
        }
      }
    }
  }
}

__global__ void nchw_to_nhwc_kernel(const float* __restrict__ src, float* __restrict__ dst, int C, int HW) {
  __shared__ float tile[32][33];
  int img = blockIdx.z;
  int p0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float* s = src + (size_t)img * C * HW;
  float* d = dst + (size_t)img * C * HW;
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int c = c0 + i, p = p0 + threadIdx.x;
    if (c < C && p < HW) tile[i][threadIdx.x] = s[(size_t)c * HW + p];
  }
  __syncthreads();
  for (int i = threadIdx.y; i < 32; i += blockDim.y) {
    int p = p0 + i, c = c0 + threadIdx.x;
    if (c < C && p < HW) d[(size_t)p * C + c] = tile[threadIdx.x][i];
  }
}

}  // namespace

int launch_gather(const cpn_render_args& a, int ray0, int nr, const float* rowaux, float* A, int a_image, cudaStream_t st,
                  float* taps) {
  (void)ray0;
  long long warps = (long long)a.B * nr * 2 * a.S * 2;
  long long blocks = (warps * 32 + 255) / 256;
  if (a_image) {
    long long rows = (long long)a.B * nr * 2 * a.S;
    if (rows >= (1ll << 31) || (a.S % GI_ROWS) != 0) {
      cpn_set_error("gather: %lld sample rows per chunk / S=%d unsupported (S must be a multiple of 8)", rows, a.S);
      return CPN_ERR_ARG;
    }
    if (!taps) {
      cpn_set_error("gather: the operand-image gather needs the tap buffer");
      return CPN_ERR_ARG;
    }
    int4* tp = reinterpret_cast<int4*>(taps);
    taps_kernel<<<(unsigned)((rows * 2 * CPN_N_LEVELS + 255) / 256), 256, 0, st>>>(a, nr, rowaux, tp);
    CPN_CHECK_LAUNCH("taps_kernel");
    static int per_row = -1;   // CPN_GATHER_SEQ=1 selects the sequential-row kernel (measured equal: 590 vs 595 us per chunk)
    if (per_row < 0) {
      const char* e = getenv("CPN_GATHER_SEQ");
      per_row = (e && atoi(e) != 0) ? 0 : 1;
    }
    const bool seq_ok = !per_row && a_image != 3 && (a.S % GS_ROWS) == 0 && a.feat_c[0] == 256 && a.feat_c[1] == 256 && a.feat_c[2] == 256 &&
                        a.feat_c[3] <= 64;
    if (seq_ok) {
      dim3 grid((unsigned)((rows + GS_ROWS - 1) / GS_ROWS), 2);
      void (*k)(cpn_render_args, int, const int4*, unsigned char*) =
          a_image == 2 ? gather_image_seq_kernel<true> : gather_image_seq_kernel<false>;
      CPN_CHECK_CUDA(cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, GS_SMEM));
      k<<<grid, 256, GS_SMEM, st>>>(a, nr, tp, reinterpret_cast<unsigned char*>(A));
      CPN_CHECK_LAUNCH("gather_image_seq_kernel");
      return CPN_OK;
    }
    dim3 grid((unsigned)((rows + GI_ROWS * GI_GROUPS - 1) / (GI_ROWS * GI_GROUPS)), 2);
    if (a_image == 3)
      gather_image_kernel<true, false><<<grid, 256, 0, st>>>(a, nr, tp, reinterpret_cast<unsigned char*>(A));
    else if (a_image == 2)
      gather_image_kernel<true, true><<<grid, 256, 0, st>>>(a, nr, tp, reinterpret_cast<unsigned char*>(A));
    else
      gather_image_kernel<false, true><<<grid, 256, 0, st>>>(a, nr, tp, reinterpret_cast<unsigned char*>(A));
  } else {
    gather_kernel<<<(unsigned)blocks, 256, 0, st>>>(a, nr, rowaux, A);
  }
  CPN_CHECK_LAUNCH("gather_kernel");
  return CPN_OK;
}

extern "C" int cpn_pack_features(const float* nchw, float* nhwc, int n_img, int C, int h, int w, void* stream) {
  if (!nchw || !nhwc || n_img <= 0 || C <= 0 || h <= 0 || w <= 0 || (C % 4) != 0) {
    cpn_set_error("cpn_pack_features: bad argument (C must be a multiple of 4)");
    return CPN_ERR_ARG;
  }
  dim3 grid((h * w + 31) / 32, (C + 31) / 32, n_img), block(32, 8);
  nchw_to_nhwc_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(nchw, nhwc, C, h * w);
  CPN_CHECK_LAUNCH("nchw_to_nhwc_kernel");
  return CPN_OK;
}

extern "C" size_t cpn_gather_rows_taps_bytes(int rows) { return (size_t)rows * 2 * CPN_N_LEVELS * 8 * sizeof(float); }

extern "C" int cpn_gather_rows(const cpn_render_args* a, int nr, const float* rowaux, void* out, int form, void* taps,
                               void* stream) {
  if (!a || !rowaux || !out || nr <= 0 || a->B <= 0 || a->S <= 0 || form < 0 || form > 3) {
    cpn_set_error("cpn_gather_rows: bad argument");
    return CPN_ERR_ARG;
  }
  for (int l = 0; l < CPN_N_LEVELS; ++l)
    if (!a->feat[l] || a->feat_h[l] <= 0 || a->feat_w[l] <= 0 || a->feat_c[l] <= 0 || (a->feat_c[l] & 3)) {
      cpn_set_error("cpn_gather_rows: bad feature level %d", l);
      return CPN_ERR_ARG;
    }
  return launch_gather(*a, 0, nr, rowaux, reinterpret_cast<float*>(out), form, (cudaStream_t)stream,
                       reinterpret_cast<float*>(taps));
}
