// One-time repack of the render-path state_dict tensors into the kernel layouts (cpn_common.cuh, namespace pw).
// Replaces the nn.Module parameter storage of models/CoPoNeRF.py:71-104 and models/lightfield.py:87-116.
#include "cpn_common.cuh"

namespace {

struct WTensor {
  const char* name;  // state_dict key
  int out, in;       // weight (out, in[,1,1]); bias: (out, 1)
};

// Order of the raw blob handed to cpn_pack_weights.
const WTensor kTensors[] = {
    {"query_encode_latent.weight", 832, 835},   {"query_encode_latent.bias", 832, 1},
    {"query_encode_latent_2.weight", 416, 832}, {"query_encode_latent_2.bias", 416, 1},
    {"latent_value.weight", 416, 832},          {"latent_value.bias", 416, 1},
    {"key_map.weight", 128, 832},               {"key_map.bias", 128, 1},
    {"key_map_2.weight", 128, 128},             {"key_map_2.bias", 128, 1},
    {"query_embed.weight", 128, 16},            {"query_embed.bias", 128, 1},
    {"query_embed_2.weight", 128, 128},         {"query_embed_2.bias", 128, 1},
    {"query_repeat_embed.weight", 128, 144},    {"query_repeat_embed.bias", 128, 1},
    {"query_repeat_embed_2.weight", 128, 128},  {"query_repeat_embed_2.bias", 128, 1},
    {"encode_latent.weight", 128, 416},         {"encode_latent.bias", 128, 1},
    {"phi.lin_in.weight", 128, 18},             {"phi.lin_in.bias", 128, 1},
    {"phi.lin_z.0.weight", 128, 832},           {"phi.lin_z.0.bias", 128, 1},
    {"phi.lin_z.1.weight", 128, 832},           {"phi.lin_z.1.bias", 128, 1},
    {"phi.lin_z.2.weight", 128, 832},           {"phi.lin_z.2.bias", 128, 1},
    {"phi.blocks.0.fc_0.weight", 128, 128},     {"phi.blocks.0.fc_0.bias", 128, 1},
    {"phi.blocks.0.fc_1.weight", 128, 128},     {"phi.blocks.0.fc_1.bias", 128, 1},
    {"phi.blocks.1.fc_0.weight", 128, 128},     {"phi.blocks.1.fc_0.bias", 128, 1},
    {"phi.blocks.1.fc_1.weight", 128, 128},     {"phi.blocks.1.fc_1.bias", 128, 1},
    {"phi.blocks.2.fc_0.weight", 128, 128},     {"phi.blocks.2.fc_0.bias", 128, 1},
    {"phi.blocks.2.fc_1.weight", 128, 128},     {"phi.blocks.2.fc_1.bias", 128, 1},
    {"phi.lin_out.weight", 3, 128},             {"phi.lin_out.bias", 3, 1},
};
constexpr int kNumTensors = sizeof(kTensors) / sizeof(kTensors[0]);

size_t tensor_offset(int i) {
  size_t off = 0;
  for (int j = 0; j < i; ++j) off += (size_t)kTensors[j].out * kTensors[j].in;
  return off;
}

// dst[k * N + n] = (k < K) ? src[n * src_ld + col0 + k] : 0    for k in [0, Kpad), n in [0, N)
__global__ void transpose_pack_kernel(const float* __restrict__ src, int src_ld, int col0, int K, int Kpad, int N,
                                      float* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= Kpad * N) return;
  int k = i / N, n = i % N;
  dst[i] = (k < K) ? src[(size_t)n * src_ld + col0 + k] : 0.f;
}

// Folded layer: WF[n][b * 832 + k] = sum_j Wx[n][b * 416 + j] * W2[j][k] (b = primary / secondary branch) and
// bF[n] = bx[n] + sum_b sum_j Wx[n][b * 416 + j] * b2[j]: x -> Wx [W2 h_p + b2 ; W2 h_s + b2] + bx as one layer on
// [h_p ; h_s]. Accumulated in double, so the folded weights are the correctly rounded products.
__global__ void fold_kernel(const float* __restrict__ Wx, const float* __restrict__ bx, const float* __restrict__ W2,
                            const float* __restrict__ b2, int N, float* __restrict__ WF, float* __restrict__ bF) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N * 1664) return;
  const int n = i / 1664, col = i % 1664, b = col / 832, k = col % 832;
  const float* wx = Wx + (size_t)n * 832 + b * 416;
  double acc = 0.0;
  for (int j = 0; j < 416; ++j) acc += (double)wx[j] * (double)W2[(size_t)j * 832 + k];
  WF[i] = (float)acc;
  if (col == 0) {
    double bb = (double)bx[n];
    for (int j = 0; j < 832; ++j) bb += (double)Wx[(size_t)n * 832 + j] * (double)b2[j % 416];
    bF[n] = (float)bb;
  }
}

// Bilinear form of a logit <Wa x + ba, Wq q + bq> = x^T (WM q + BM) + (WS . q + CS):
// WM[n][k] = sum_j Wa[j][n] Wq[j][k], BM[n] = sum_j Wa[j][n] bq[j], WS[k] = sum_j ba[j] Wq[j][k], CS = ba . bq (double).
__global__ void bilinear_fold_kernel(const float* __restrict__ Wa, const float* __restrict__ ba, const float* __restrict__ Wq,
                                     const float* __restrict__ bq, float* __restrict__ WM, float* __restrict__ BM,
                                     float* __restrict__ WS) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < 128 * 128) {
    const int n = i / 128, k = i % 128;
    double acc = 0.0;
    for (int j = 0; j < 128; ++j) acc += (double)Wa[j * 128 + n] * (double)Wq[j * 128 + k];
    WM[i] = (float)acc;
  } else if (i < 128 * 128 + 128) {
    const int n = i - 128 * 128;
    double acc = 0.0, acc2 = 0.0;
    for (int j = 0; j < 128; ++j) {
      acc += (double)Wa[j * 128 + n] * (double)bq[j];
      acc2 += (double)ba[j] * (double)Wq[j * 128 + n];
    }
    BM[n] = (float)acc;
    WS[n] = (float)acc2;
  } else if (i == 128 * 128 + 128) {
    double acc = 0.0;
    for (int j = 0; j < 128; ++j) acc += (double)ba[j] * (double)bq[j];
    WS[128] = (float)acc;
  }
}

// G fold, stage 1: M1 = Wqr[:, :128] We (128 x 416) in double.  Wqr (128, 144) = query_repeat_embed.weight, whose first
// 128 input channels see encode_latent's output (CoPoNeRF.py:467-472); We (128, 416) = encode_latent.weight.
__global__ void gfold_m1_kernel(const float* __restrict__ Wqr, const float* __restrict__ We, double* __restrict__ M1) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 416) return;
  const int n = i / 416, j = i % 416;
  double acc = 0.0;
  for (int m = 0; m < 128; ++m) acc += (double)Wqr[n * 144 + m] * (double)We[m * 416 + j];
  M1[i] = acc;
}
// stage 2: G = M1 WVF (128 x 1664), g0 = M1 bVF + Wqr[:, :128] be + bqr, with WVF / bVF the folded latent_value.
__global__ void gfold_kernel(const double* __restrict__ M1, const float* __restrict__ WVF, const float* __restrict__ bVF,
                             const float* __restrict__ Wqr, const float* __restrict__ be, const float* __restrict__ bqr,
                             float* __restrict__ G, float* __restrict__ g0) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= 128 * 1664) return;
  const int n = i / 1664, col = i % 1664;
  double acc = 0.0;
  for (int j = 0; j < 416; ++j) acc += M1[n * 416 + j] * (double)WVF[(size_t)j * 1664 + col];
  G[i] = (float)acc;
  if (col == 0) {
    double b = (double)bqr[n];
    for (int j = 0; j < 416; ++j) b += M1[n * 416 + j] * (double)bVF[j];
    for (int m = 0; m < 128; ++m) b += (double)Wqr[n * 144 + m] * (double)be[m];
    g0[n] = (float)b;
  }
}

struct Job {
  int tensor, col0, K, Kpad;  // K == 0: plain copy of the whole tensor
  size_t dst;
};

}  // namespace

extern "C" int cpn_n_weight_tensors(void) { return kNumTensors; }
extern "C" const char* cpn_weight_name(int i) { return (i >= 0 && i < kNumTensors) ? kTensors[i].name : nullptr; }
extern "C" size_t cpn_weight_numel(int i) {
  return (i >= 0 && i < kNumTensors) ? (size_t)kTensors[i].out * kTensors[i].in : 0;
}
extern "C" size_t cpn_raw_weights_floats(void) { return tensor_offset(kNumTensors); }
extern "C" size_t cpn_packed_weights_bytes(void) { return cpn_packed_fp32_floats() * sizeof(float) + cpn_tc_weights_bytes(); }

size_t cpn_packed_fp32_floats() { return (pw::FP32_END + 63) / 64 * 64; }

extern "C" int cpn_pack_weights(const float* src, void* dst_v, void* stream) {
  if (!src || !dst_v) {
    cpn_set_error("cpn_pack_weights: null pointer");
    return CPN_ERR_ARG;
  }
  cudaStream_t st = (cudaStream_t)stream;
  float* dst = reinterpret_cast<float*>(dst_v);
  const Job jobs[] = {
      {0, 0, 835, CPN_KA, pw::W1T},   {1, 0, 0, 0, pw::B1},
      {2, 0, 832, 832, pw::W2T},      {3, 0, 0, 0, pw::B2},
      {4, 0, 832, 832, pw::WVT},      {5, 0, 0, 0, pw::BV},
      {6, 0, 832, 832, pw::WKT},      {7, 0, 0, 0, pw::BK},
      {8, 0, 128, 128, pw::WK2T},     {9, 0, 0, 0, pw::BK2},
      {10, 0, 16, 16, pw::WQT},       {11, 0, 0, 0, pw::BQ},
      {12, 0, 128, 128, pw::WQ2T},    {13, 0, 0, 0, pw::BQ2},
      {14, 0, 128, 128, pw::WQRA_T},  {14, 128, 16, 16, pw::WQRB_T},
      {15, 0, 0, 0, pw::BQR},
      {16, 0, 128, 128, pw::WQR2T},   {17, 0, 0, 0, pw::BQR2},
      {18, 0, 416, 416, pw::WET},     {19, 0, 0, 0, pw::BE},
      {20, 0, 18, 18, pw::PHI_INT},   {21, 0, 0, 0, pw::PHI_BIN},
      {22, 0, 832, 832, pw::PHI_ZT},  {23, 0, 0, 0, pw::PHI_BZ},
      {24, 0, 832, 832, pw::PHI_ZT + 832 * 128},      {25, 0, 0, 0, pw::PHI_BZ + 128},
      {26, 0, 832, 832, pw::PHI_ZT + 2 * 832 * 128},  {27, 0, 0, 0, pw::PHI_BZ + 256},
      {28, 0, 128, 128, pw::PHI_F0T},                 {29, 0, 0, 0, pw::PHI_B0},
      {30, 0, 128, 128, pw::PHI_F1T},                 {31, 0, 0, 0, pw::PHI_B1},
      {32, 0, 128, 128, pw::PHI_F0T + 128 * 128},     {33, 0, 0, 0, pw::PHI_B0 + 128},
      {34, 0, 128, 128, pw::PHI_F1T + 128 * 128},     {35, 0, 0, 0, pw::PHI_B1 + 128},
      {36, 0, 128, 128, pw::PHI_F0T + 2 * 128 * 128}, {37, 0, 0, 0, pw::PHI_B0 + 256},
      {38, 0, 128, 128, pw::PHI_F1T + 2 * 128 * 128}, {39, 0, 0, 0, pw::PHI_B1 + 256},
      {40, 0, 0, 0, pw::PHI_OUT},                     {41, 0, 0, 0, pw::PHI_BOUT},
  };
  CPN_CHECK_CUDA(cudaMemsetAsync(dst, 0, cpn_packed_fp32_floats() * sizeof(float), st));
  for (const Job& j : jobs) {
    const WTensor& t = kTensors[j.tensor];
    const float* s = src + tensor_offset(j.tensor);
    if (j.K == 0) {
      CPN_CHECK_CUDA(cudaMemcpyAsync(dst + j.dst, s, (size_t)t.out * t.in * sizeof(float), cudaMemcpyDeviceToDevice, st));
    } else {
      int total = j.Kpad * t.out;
      transpose_pack_kernel<<<(total + 255) / 256, 256, 0, st>>>(s, t.in, j.col0, j.K, j.Kpad, t.out, dst + j.dst);
      CPN_CHECK_LAUNCH("transpose_pack_kernel");
    }
  }
  // query_encode_latent_2 folded into latent_value and key_map
  const float *W2 = src + tensor_offset(2), *b2 = src + tensor_offset(3);
  fold_kernel<<<(416 * 1664 + 255) / 256, 256, 0, st>>>(src + tensor_offset(4), src + tensor_offset(5), W2, b2, 416,
                                                        dst + pw::WVF, dst + pw::BVF);
  CPN_CHECK_LAUNCH("fold_kernel");
  fold_kernel<<<(128 * 1664 + 255) / 256, 256, 0, st>>>(src + tensor_offset(6), src + tensor_offset(7), W2, b2, 128,
                                                        dst + pw::WKF, dst + pw::BKF);
  CPN_CHECK_LAUNCH("fold_kernel");
  // round-2 query bias as a linear map of the hidden layer (cpn_common.cuh, pw::WG): needs WVF / BVF from above
  {
    double* M1 = reinterpret_cast<double*>(dst + pw::M1D);
    gfold_m1_kernel<<<(128 * 416 + 255) / 256, 256, 0, st>>>(src + tensor_offset(14), src + tensor_offset(18), M1);
    CPN_CHECK_LAUNCH("gfold_m1_kernel");
    gfold_kernel<<<(128 * 1664 + 255) / 256, 256, 0, st>>>(M1, dst + pw::WVF, dst + pw::BVF, src + tensor_offset(14),
                                                           src + tensor_offset(19), src + tensor_offset(15), dst + pw::WG,
                                                           dst + pw::BG);
    CPN_CHECK_LAUNCH("gfold_kernel");
  }
  // attention logits as bilinear forms: key_map_2 (tensors 8, 9) and query_repeat_embed_2 (16, 17) against query_embed_2 (12, 13)
  const float *Wq2 = src + tensor_offset(12), *bq2 = src + tensor_offset(13);
  bilinear_fold_kernel<<<(128 * 128 + 129 + 255) / 256, 256, 0, st>>>(src + tensor_offset(8), src + tensor_offset(9), Wq2, bq2,
                                                                      dst + pw::WM12, dst + pw::BM12, dst + pw::WS1);
  CPN_CHECK_LAUNCH("bilinear_fold_kernel");
  bilinear_fold_kernel<<<(128 * 128 + 129 + 255) / 256, 256, 0, st>>>(src + tensor_offset(16), src + tensor_offset(17), Wq2, bq2,
                                                                      dst + pw::WM12 + 128 * 128, dst + pw::BM12 + 128, dst + pw::WS2);
  CPN_CHECK_LAUNCH("bilinear_fold_kernel");
  return cpn_pack_tc_weights(src, dst, reinterpret_cast<char*>(dst_v) + cpn_packed_fp32_floats() * sizeof(float), st);
}
