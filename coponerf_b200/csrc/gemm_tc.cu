// Tensor-core GEMM for the four large 1x1 convolutions of the per-sample encoder
// (models/CoPoNeRF.py:387-408): query_encode_latent, query_encode_latent_2, latent_value, key_map.
//
// fp32 parity on fp16 tensor cores: every operand is split x = hi + lo into two fp16 values (error 2^-22 |x|)
// and the product is accumulated in fp32 as  hi*hi + hi*lo + lo*hi  (three tcgen05.mma per k-step; the dropped
// lo*lo term is 2^-22 relative). Weights are pre-split, pre-scaled by a per-layer power of two (so the lo halves
// stay out of the fp16 subnormal range) and pre-tiled into the exact shared-memory image the MMA reads, so a
// plain bulk copy (TMA engine, one instruction per stage) moves them. The activation operand is fp32 in global
// memory: producer warps load it coalesced, split it and write the K-major core-matrix layout themselves.
//
// CTA = one 128-row x NT-column output tile. 6 warps:
//   warp 0   lane 0: bulk-copies the weight tile of each k-chunk into the stage ring
//   warp 1   allocates TMEM; lane 0 issues the MMAs and commits stage-free / accumulator-ready barriers
//   warp 2-5 produce the A operand (fp32 -> hi/lo fp16, K-major, no swizzle), then run the epilogue
//            (TMEM -> registers -> scale, bias, ReLU -> global)
#include "cpn_common.cuh"
#include "tc_common.cuh"

namespace {

using namespace tc;

constexpr int BM = 128;            // rows per CTA (UMMA M)
constexpr int BK = 32;             // k per stage: 2 MMA k-steps of 16
constexpr int STAGES = 4;
constexpr int NT_MAX = 208;        // widest N tile (832 = 4 x 208, 416 = 2 x 208)
constexpr int A_LBO = BM * 16 + 32;             // bytes between 8-wide k-chunks of A (+32: conflict-free STS.128)
constexpr int A_HALF = (BK / 8) * A_LBO;        // hi (or lo) half of one A stage
constexpr int W_STAGE_MAX = 2 * (BK / 8) * NT_MAX * 16;
constexpr int STAGE_BYTES = 2 * A_HALF + W_STAGE_MAX;
constexpr int SMEM_BYTES = STAGES * STAGE_BYTES + 256;
constexpr int TMEM_COLS = 256;
constexpr int NUM_THREADS = 192;

struct TcLayer {
  int tensor;      // index into the raw state_dict blob (weights.cu order)
  int out, in;     // weight shape
  int kpad;        // K padded to a multiple of BK (zero weights beyond `in`)
  int nt;          // N tile
  size_t bias;     // offset of the fp32 bias in the packed fp32 section
};
// raw blob offsets (floats) of the four weights, see weights.cu kTensors
constexpr size_t RAW_W1 = 0;
constexpr size_t RAW_W2 = RAW_W1 + 832 * 835 + 832;
constexpr size_t RAW_WV = RAW_W2 + 416 * 832 + 416;
constexpr size_t RAW_WK = RAW_WV + 416 * 832 + 416;
const TcLayer kLayers[4] = {
    {0, 832, 835, 864, 208, pw::B1},
    {2, 416, 832, 832, 208, pw::B2},
    {4, 416, 832, 832, 208, pw::BV},
    {6, 128, 832, 832, 128, pw::BK},
};
const size_t kRawOff[4] = {RAW_W1, RAW_W2, RAW_WV, RAW_WK};
constexpr size_t TC_HEADER_BYTES = 256;   // [0..3] 1/scale per layer, [4..7] scale, [8..11] absmax bits

size_t layer_bytes(int l) { return (size_t)kLayers[l].out * kLayers[l].kpad * 4; }  // hi + lo fp16
size_t layer_offset(int l) {
  size_t off = TC_HEADER_BYTES;
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

// power-of-two scale that brings max|w| to about 2^10
__device__ __forceinline__ float layer_scale(unsigned int absmax_bits) {
  float m = __uint_as_float(absmax_bits);
  if (!(m > 0.f) || isinf(m)) return 1.f;
  int e;
  frexpf(m, &e);              // m = f * 2^e, f in [0.5, 1)
  return ldexpf(1.f, 10 - e);
}

// dst tile (nt, kc): [hi | lo] x [BK/8 k-chunks][NT rows][8 halves]
__global__ void pack_tc_kernel(const float* __restrict__ w, int out, int in, int kpad, int NT, const unsigned int* absmax,
                               __half* __restrict__ dst, float* __restrict__ header, int layer) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  size_t total = (size_t)out * kpad;
  float scale = layer_scale(*absmax);
  if (i == 0) {
    header[layer] = 1.f / scale;
    header[4 + layer] = scale;
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
}

// ---------------------------------------------------------------------------------------------- the GEMM
struct GemmArgs {
  const float* A;
  int lda, kreal, M;
  float* C;
  int ldc, N, relu;
  const unsigned char* wtiles;   // this layer's tiles
  const float* bias;
  const float* inv_scale;        // header[layer]
  int kchunks, NT;
  uint32_t idesc;
};

__global__ void __launch_bounds__(NUM_THREADS, 1) gemm_tc_kernel(GemmArgs g) {
  extern __shared__ __align__(128) unsigned char smem[];
  __shared__ __align__(8) uint64_t bars[3 * STAGES + 1];
  __shared__ uint32_t tmem_base_s;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int n_tile = blockIdx.x, m0 = blockIdx.y * BM;
  const uint32_t smem0 = smem_u32(smem);
  const uint32_t full_a = smem_u32(&bars[0]), full_w = smem_u32(&bars[STAGES]), empty = smem_u32(&bars[2 * STAGES]),
                 accum = smem_u32(&bars[3 * STAGES]);
  const int NT = g.NT;
  const uint32_t w_half = (uint32_t)(BK / 8) * NT * 16;      // bytes of the hi (or lo) part of a weight tile

  if (threadIdx.x == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(full_a + 8 * s, 128);
      mbar_init(full_w + 8 * s, 1);
      mbar_init(empty + 8 * s, 1);
    }
    mbar_init(accum, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(smem_u32(&tmem_base_s), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = tmem_base_s;

  if (warp == 0) {
    if (lane == 0) {
      const unsigned char* src = g.wtiles + (size_t)n_tile * g.kchunks * 2 * w_half;
      for (int i = 0; i < g.kchunks; ++i) {
        int s = i % STAGES;
        uint32_t u = i / STAGES;
        mbar_wait(empty + 8 * s, (u & 1) ^ 1);
        mbar_arrive_expect_tx(full_w + 8 * s, 2 * w_half);
        bulk_g2s(smem0 + s * STAGE_BYTES + 2 * A_HALF, src + (size_t)i * 2 * w_half, 2 * w_half, full_w + 8 * s);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      for (int i = 0; i < g.kchunks; ++i) {
        int s = i % STAGES;
        uint32_t u = i / STAGES;
        mbar_wait(full_a + 8 * s, u & 1);
        mbar_wait(full_w + 8 * s, u & 1);
        tcgen05_fence_after();
        uint32_t a_hi = smem0 + s * STAGE_BYTES, a_lo = a_hi + A_HALF;
        uint32_t b_hi = a_hi + 2 * A_HALF, b_lo = b_hi + w_half;
#pragma unroll
        for (int j = 0; j < BK / 16; ++j) {
          uint64_t da_hi = make_desc(a_hi + j * 2 * A_LBO, A_LBO, 128);
          uint64_t da_lo = make_desc(a_lo + j * 2 * A_LBO, A_LBO, 128);
          uint64_t db_hi = make_desc(b_hi + j * 2 * NT * 16, NT * 16, 128);
          uint64_t db_lo = make_desc(b_lo + j * 2 * NT * 16, NT * 16, 128);
          mma_f16_ss(tmem, da_hi, db_hi, g.idesc, (i | j) != 0);
          mma_f16_ss(tmem, da_hi, db_lo, g.idesc, 1);
          mma_f16_ss(tmem, da_lo, db_hi, g.idesc, 1);
        }
        mma_commit(empty + 8 * s);   // the stage is free once these MMAs have read it
      }
      mma_commit(accum);
    }
  } else {
    // ---- A producer: 4 warps, each covers 32 rows of the tile in 4 passes of 8 rows x 32 k
    const int wq = warp - 2;
    const int r_in = lane >> 2, c = lane & 3;
    for (int i = 0; i < g.kchunks; ++i) {
      int s = i % STAGES;
      uint32_t u = i / STAGES;
      const int k = i * BK + c * 8;
      float4 v[4][2];
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        int row = m0 + wq * 32 + it * 8 + r_in;
        if (row < g.M && k < g.kreal) {
          const float4* p = reinterpret_cast<const float4*>(g.A + (size_t)row * g.lda + k);
          v[it][0] = __ldg(p);
          v[it][1] = __ldg(p + 1);
        } else {
          v[it][0] = v[it][1] = make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
      mbar_wait(empty + 8 * s, (u & 1) ^ 1);
      unsigned char* a_hi = smem + s * STAGE_BYTES;
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        int r = wq * 32 + it * 8 + r_in;
        uint4 hi, lo;
        split8(v[it][0], v[it][1], hi, lo);
        *reinterpret_cast<uint4*>(a_hi + c * A_LBO + r * 16) = hi;
        *reinterpret_cast<uint4*>(a_hi + A_HALF + c * A_LBO + r * 16) = lo;
      }
      fence_proxy_async_smem();
      mbar_arrive(full_a + 8 * s);
    }
    // ---- epilogue: warp w may touch TMEM lanes 32 * (w % 4) .. + 31
    mbar_wait(accum, 0);
    tcgen05_fence_after();
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const float inv = *g.inv_scale;
    const int n0 = n_tile * NT;
    for (int c0 = 0; c0 < NT; c0 += 16) {
      float v[16];
      tmem_ld16(tmem + ((uint32_t)(q * 32) << 16) + c0, v);
      if (row < g.M) {
        float* out = g.C + (size_t)row * g.ldc + n0 + c0;
#pragma unroll
        for (int j = 0; j < 16; j += 4) {
          float4 b = *reinterpret_cast<const float4*>(g.bias + n0 + c0 + j);
          float4 o = make_float4(v[j] * inv + b.x, v[j + 1] * inv + b.y, v[j + 2] * inv + b.z, v[j + 3] * inv + b.w);
          if (g.relu) {
            o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f);
          }
          *reinterpret_cast<float4*>(out + j) = o;
        }
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, TMEM_COLS);
}

}  // namespace

size_t cpn_tc_weights_bytes() { return layer_offset(4); }

int cpn_pack_tc_weights(const float* raw, void* dst_v, cudaStream_t st) {
  unsigned char* dst = reinterpret_cast<unsigned char*>(dst_v);
  float* header = reinterpret_cast<float*>(dst);
  unsigned int* absmax = reinterpret_cast<unsigned int*>(dst) + 8;
  CPN_CHECK_CUDA(cudaMemsetAsync(dst, 0, cpn_tc_weights_bytes(), st));
  for (int l = 0; l < 4; ++l) {
    const TcLayer& L = kLayers[l];
    size_t n = (size_t)L.out * L.in;
    absmax_kernel<<<64, 256, 0, st>>>(raw + kRawOff[l], n, absmax + l);
    CPN_CHECK_LAUNCH("absmax_kernel");
    size_t total = (size_t)L.out * L.kpad;
    pack_tc_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(raw + kRawOff[l], L.out, L.in, L.kpad, L.nt, absmax + l,
                                                                    reinterpret_cast<__half*>(dst + layer_offset(l)), header, l);
    CPN_CHECK_LAUNCH("pack_tc_kernel");
  }
  return CPN_OK;
}

int launch_gemm_tc(const void* packed, int layer, const float* A, int lda, float* C, int ldc, int M, int relu,
                   cudaStream_t st) {
  if (!packed || !A || !C || layer < 0 || layer > 3 || M < 0 || (lda & 3) || (ldc & 3)) {
    cpn_set_error("gemm_tc: bad argument (layer=%d M=%d lda=%d ldc=%d)", layer, M, lda, ldc);
    return CPN_ERR_ARG;
  }
  if (M == 0) return CPN_OK;
  CPN_CHECK_CUDA(cudaFuncSetAttribute(gemm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM_BYTES));
  const TcLayer& L = kLayers[layer];
  const unsigned char* tcw = reinterpret_cast<const unsigned char*>(packed) + cpn_packed_fp32_floats() * sizeof(float);
  GemmArgs g;
  g.A = A;
  g.lda = lda;
  g.kreal = L.in < lda ? ((L.in + 7) / 8 * 8) : lda;   // columns of A that exist (the zero pad of 835 -> 848 is read)
  if (g.kreal > lda) g.kreal = lda;
  g.M = M;
  g.C = C;
  g.ldc = ldc;
  g.N = L.out;
  g.relu = relu;
  g.wtiles = tcw + layer_offset(layer);
  g.bias = reinterpret_cast<const float*>(packed) + L.bias;
  g.inv_scale = reinterpret_cast<const float*>(tcw) + layer;
  g.kchunks = L.kpad / BK;
  g.NT = L.nt;
  g.idesc = make_idesc_f16(BM, L.nt);
  dim3 grid(L.out / L.nt, (M + BM - 1) / BM);
  gemm_tc_kernel<<<grid, NUM_THREADS, SMEM_BYTES, st>>>(g);
  CPN_CHECK_LAUNCH("gemm_tc_kernel");
  return CPN_OK;
}

extern "C" int cpn_gemm_tc(const void* packed, int layer, const float* A, int lda, float* C, int ldc, int M, int relu,
                           void* stream) {
  return launch_gemm_tc(packed, layer, A, lda, C, ldc, M, relu, (cudaStream_t)stream);
}
