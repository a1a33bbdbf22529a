// cpn_render_rays: the per-ray stage of CoPoNeRF.forward() (models/CoPoNeRF.py:246-566) as one
// asynchronous sequence of kernels per chunk of rays. Also the library's error string and version.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>
#include <mutex>
#include "cpn_common.cuh"

static thread_local char g_err[512] = "";

void cpn_set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}

extern "C" const char* cpn_last_error(void) { return g_err; }
extern "C" int cpn_version(void) { return 200; }   // bumped whenever an entry point or struct changes (_lib.ABI_VERSION)
extern "C" size_t cpn_sizeof_render_args(void) { return sizeof(cpn_render_args); }

// ---- optional device timing of the dominant kernel (bench.py's roofline) -------------------
// Off unless cpn_prof_begin() was called; events are recorded on the caller's stream around every
// launch of the query_encode_latent GEMM. Not thread-safe: one profiling session per process.
static cudaEvent_t* g_prof_ev = nullptr;
static int g_prof_cap = 0, g_prof_n = 0;

extern "C" int cpn_prof_begin(int max_launches) {
  if (g_prof_ev || max_launches <= 0) {
    cpn_set_error("cpn_prof_begin: session already open or bad capacity");
    return CPN_ERR_ARG;
  }
  g_prof_ev = new cudaEvent_t[2 * (size_t)max_launches];
  for (int i = 0; i < 2 * max_launches; ++i) CPN_CHECK_CUDA(cudaEventCreate(&g_prof_ev[i]));
  g_prof_cap = max_launches;
  g_prof_n = 0;
  return CPN_OK;
}

// Waits for the recorded events, returns the summed duration and the number of launches, closes the session.
extern "C" int cpn_prof_end(float* total_ms, int* launches) {
  if (!g_prof_ev) {
    cpn_set_error("cpn_prof_end: no session");
    return CPN_ERR_ARG;
  }
  float tot = 0.f;
  for (int i = 0; i < g_prof_n; ++i) {
    float ms = 0.f;
    CPN_CHECK_CUDA(cudaEventSynchronize(g_prof_ev[2 * i + 1]));
    CPN_CHECK_CUDA(cudaEventElapsedTime(&ms, g_prof_ev[2 * i], g_prof_ev[2 * i + 1]));
    tot += ms;
  }
  if (total_ms) *total_ms = tot;
  if (launches) *launches = g_prof_n;
  for (int i = 0; i < 2 * g_prof_cap; ++i) cudaEventDestroy(g_prof_ev[i]);
  delete[] g_prof_ev;
  g_prof_ev = nullptr;
  g_prof_cap = g_prof_n = 0;
  return CPN_OK;
}

namespace {

struct ProfScope {  // records an event pair around the enclosed launches when a session is open
  cudaStream_t st;
  bool on;
  explicit ProfScope(cudaStream_t s) : st(s), on(g_prof_ev && g_prof_n < g_prof_cap) {
    if (on) cudaEventRecord(g_prof_ev[2 * g_prof_n], st);
  }
  ~ProfScope() {
    if (on) {
      cudaEventRecord(g_prof_ev[2 * g_prof_n + 1], st);
      ++g_prof_n;
    }
  }
};

// Workspace carve-up for one chunk of `nr` rays of each of B pairs (R = B * nr * 2 * S sample rows).
struct Workspace {
  float *seg, *rowaux, *local16, *A, *H1, *E, *V, *K1, *Kk, *Q1, *Qe, *r1, *wp, *zemb, *rbias, *lg1, *lg2, *wt1, *wt2,
      *hbar, *s1, *s2, *Qm, *taps;
  size_t bytes;
};

// `flags` selects the code path (CPN_FLAG_*): buffers only an alternative path reads are not carved for the default one
Workspace carve(void* base, int B, int nr, int S, int flags) {
  Workspace w;
  const bool simt = flags & CPN_FLAG_SIMT_ONLY, nofold = simt || (flags & CPN_FLAG_NO_FOLD);
  const bool early_v = nofold || (flags & CPN_FLAG_EARLY_V), nobil = nofold || (flags & CPN_FLAG_NO_BILINEAR);
  const bool nogfold = early_v || nobil || (flags & CPN_FLAG_NO_GFOLD);
  size_t rays = (size_t)B * nr, R = (rays * 2 * S + 127) / 128 * 128;   // whole 128-row tiles
  size_t off = 0;
  auto take = [&](size_t floats) {
    float* p = base ? reinterpret_cast<float*>(reinterpret_cast<char*>(base) + off) : nullptr;
    off += (floats * sizeof(float) + 255) / 256 * 256;
    return p;
  };
  w.seg = take(rays * 2 * 6);
  w.rowaux = take(R * CPN_ROWAUX);
  w.local16 = take(R * 16);
  w.taps = take(R * 2 * CPN_N_LEVELS * 8);   // bilinear taps of every (row, branch, level)
  w.A = take(R * 2 * CPN_KA_IMG);   // fp32 rows of CPN_KA, or the operand image with K = CPN_KA_IMG
  w.H1 = take(R * 2 * CPN_FEAT_DIM);
  w.E = take(nofold ? R * CPN_FEAT_DIM : 0);      // query_encode_latent_2 output: unfolded paths only
  w.V = take(early_v ? R * CPN_LATENT : 0);       // per-sample values: early-V / unfolded paths only
  w.K1 = take(R * CPN_HIDDEN);                    // key hidden layer image, or G h + g0 of the default path
  w.Kk = take(nobil ? R * CPN_HIDDEN : 0);        // key_map_2 / query_repeat_embed_2 outputs: three-layer logits only
  w.Q1 = take(R * CPN_HIDDEN);
  w.Qe = take(nobil ? R * CPN_HIDDEN : 0);
  w.r1 = take(rays * CPN_LATENT);
  w.wp = take(rays * 4);
  w.zemb = take(rays * CPN_HIDDEN);
  w.rbias = take(rays * CPN_HIDDEN);
  w.lg1 = take(R);   // attention logits of round 1 / round 2, one per sample row
  w.lg2 = take(R);
  w.wt1 = take(R);   // late readout: softmax weights of round 1 / round 2, weighted hidden layer of round 1
  w.wt2 = take(R);
  w.hbar = take(nogfold ? (rays + 127) / 128 * 128 * 2 * CPN_FEAT_DIM : 0);   // round-1 readout image (per-ray chain path)
  w.s1 = take(R);    // bilinear logits: per-row scalar terms of round 1 / round 2
  w.s2 = take(R);
  w.Qm = take(R * 2 * CPN_HIDDEN);   // [WM q + BM] of both rounds, column-blocked: [row tile][16 blocks][128][16]
  w.bytes = off;
  return w;
}

int check_args(const cpn_render_args* a) {
  if (!a) {
    cpn_set_error("cpn_render_rays: null args");
    return CPN_ERR_ARG;
  }
  if (a->B <= 0 || a->N < 0 || a->H <= 0 || a->W <= 0 || a->chunk_rays <= 0 || a->flow_h <= 0) {
    cpn_set_error("cpn_render_rays: bad sizes B=%d N=%d H=%d W=%d chunk_rays=%d flow_h=%d", a->B, a->N, a->H, a->W,
                  a->chunk_rays, a->flow_h);
    return CPN_ERR_ARG;
  }
  if (a->S <= 0 || a->S > 128 || (a->S % 32) != 0) {
    cpn_set_error("cpn_render_rays: S=%d unsupported (S must be a multiple of 32, S <= 128)", a->S);
    return CPN_ERR_ARG;
  }
  int csum = 0;
  for (int l = 0; l < CPN_N_LEVELS; ++l) {
    if (!a->feat[l] || a->feat_h[l] <= 0 || a->feat_w[l] <= 0 || a->feat_c[l] <= 0 || (a->feat_c[l] & 3)) {
      cpn_set_error("cpn_render_rays: bad feature level %d", l);
      return CPN_ERR_ARG;
    }
    csum += a->feat_c[l];
  }
  if (csum != CPN_FEAT_DIM) {
    cpn_set_error("cpn_render_rays: feature channels sum to %d, expected %d", csum, CPN_FEAT_DIM);
    return CPN_ERR_ARG;
  }
  const void* ptrs[] = {a->pair_consts, a->uv, a->interval, a->weights, a->up_flow2, a->mask_padded2, a->rgb,
                        a->valid_mask, a->depth_ray, a->at_wt, a->at_wt_max, a->pixel_val, a->coords, a->T_to_C1_pts,
                        a->T_to_C2_pts, a->C2_pts_to_C1, a->mask_c2, a->matchability_cycle_mask, a->workspace};
  for (const void* p : ptrs)
    if (!p) {
      cpn_set_error("cpn_render_rays: null pointer in args");
      return CPN_ERR_ARG;
    }
  return CPN_OK;
}

#define CPN_TRY(expr)            \
  do {                           \
    int _s = (expr);             \
    if (_s != CPN_OK) return _s; \
  } while (0)

// ---- chunk lanes -------------------------------------------------------------------------------------------
// Chunks of rays are independent, so they are issued round-robin on a few internal streams ("lanes"), each with
// its own workspace: the gather / attention / phi kernels of one chunk (LSU- and latency-bound, no shared memory)
// then run underneath the tensor-core GEMMs of another. Lanes fork from and join back into the caller's stream
// through events; the pool is created on first use, per device.
constexpr int MAX_LANES = 4, MAX_DEVICES = 64;
struct LanePool {
  bool ready = false;
  cudaStream_t stream[MAX_LANES];
  cudaEvent_t done[MAX_LANES];
  cudaEvent_t fork;
};
LanePool g_pool[MAX_DEVICES];
std::mutex g_pool_mutex;   // creation / destruction only; a device's pool is used by one host thread at a time (header)

int get_pool(LanePool** out) {
  int dev = 0;
  CPN_CHECK_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= MAX_DEVICES) {
    cpn_set_error("device index %d out of range", dev);
    return CPN_ERR_ARG;
  }
  LanePool& p = g_pool[dev];
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  if (!p.ready) {
    for (int i = 0; i < MAX_LANES; ++i) {
      CPN_CHECK_CUDA(cudaStreamCreateWithFlags(&p.stream[i], cudaStreamNonBlocking));
      CPN_CHECK_CUDA(cudaEventCreateWithFlags(&p.done[i], cudaEventDisableTiming));
    }
    CPN_CHECK_CUDA(cudaEventCreateWithFlags(&p.fork, cudaEventDisableTiming));
    p.ready = true;
  }
  *out = &p;
  return CPN_OK;
}

}  // namespace

// Releases the library's only persistent resources: the per-device lane streams / events (created by the first
// cpn_render_rays call with lanes > 1). Safe to call at any time no render is in flight; they are re-created on demand.
extern "C" int cpn_shutdown(void) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  int cur = 0;
  cudaGetDevice(&cur);
  for (int d = 0; d < MAX_DEVICES; ++d) {
    LanePool& p = g_pool[d];
    if (!p.ready) continue;
    cudaSetDevice(d);
    for (int i = 0; i < MAX_LANES; ++i) {
      cudaStreamSynchronize(p.stream[i]);
      cudaStreamDestroy(p.stream[i]);
      cudaEventDestroy(p.done[i]);
    }
    cudaEventDestroy(p.fork);
    p.ready = false;
  }
  cudaSetDevice(cur);
  return CPN_OK;
}

namespace {

bool use_tc(const cpn_render_args& a) { return !(a.flags & CPN_FLAG_SIMT_ONLY); }
// encoder-input form for sample / gather: 0 fp32 rows, 1 operand image (f16x3), 2 operand image (f8 scheme)
int a_form(const cpn_render_args& a) { return !use_tc(a) ? 0 : ((a.flags & CPN_FLAG_F16X3) ? 1 : 2); }
int tc_scheme(const cpn_render_args& a) { return (a.flags & CPN_FLAG_F16X3) ? CPN_TC_F16X3 : 0; }

// fp32 CUDA-core layer: C = act(A * Wt + bias)
int dense_simt(const cpn_render_args& a, const float* A, int lda, size_t wt, size_t bias, float* C, int ldc, int M, int N,
               int K, int relu, cudaStream_t st, int remap256 = 0) {
  const float* W = reinterpret_cast<const float*>(a.weights);
  return launch_gemm_simt(A, lda, W + wt, W + bias, nullptr, 1, C, ldc, M, N, K, relu, st, remap256);
}

}  // namespace

// image-level part of the workspace: the per-ray latent z of every ray (phi runs once over the whole image)
// image-level buffers: z of every ray; for the late readout also the round-2 readout of every ray as an operand image
// (whole 128-ray tiles x 52 k-chunks) and the result of the one GEMM that turns it into R2
size_t z_bytes(int B, int N) { return ((size_t)B * N * CPN_LATENT * sizeof(float) + 255) / 256 * 256; }
size_t hbar_all_bytes(int B, int N) {
  return ((size_t)B * N + 127) / 128 * (2 * CPN_FEAT_DIM / ACT_BK) * (size_t)ACT_CHUNK_BYTES;
}
size_t image_bytes(int B, int N) { return 2 * z_bytes(B, N) + hbar_all_bytes(B, N); }

extern "C" size_t cpn_render_workspace_bytes_for(int B, int N, int chunk_rays, int S, int lanes, int flags) {
  if (B <= 0 || N < 0 || chunk_rays <= 0 || S <= 0 || lanes < 1 || lanes > MAX_LANES) return 0;
  int chunk = chunk_rays < N ? chunk_rays : (N > 0 ? N : 1);
  return image_bytes(B, N) + carve(nullptr, B, chunk, S, flags).bytes * lanes;
}
// every path: the largest carve-up (all alternative-path buffers)
extern "C" size_t cpn_render_workspace_bytes(int B, int N, int chunk_rays, int S, int lanes) {
  return cpn_render_workspace_bytes_for(B, N, chunk_rays, S, lanes, CPN_FLAG_SIMT_ONLY);
}

extern "C" int cpn_render_launch_count(const cpn_render_args* a) {
  if (!a || a->chunk_rays <= 0) return 0;
  int chunks = (a->N + a->chunk_rays - 1) / a->chunk_rays;
  const bool unfolded = (a->flags & CPN_FLAG_NO_FOLD) || (a->flags & CPN_FLAG_SIMT_ONLY);
  const bool late = !unfolded && !(a->flags & CPN_FLAG_EARLY_V);
  int per_chunk = (unfolded ? 17 : (late ? 19 : 16)) + ((a->flags & CPN_FLAG_SIMT_ONLY) ? 0 : 1);   // + taps_kernel
  if (!unfolded && !(a->flags & CPN_FLAG_NO_BILINEAR)) per_chunk -= 2;   // one GEMM over the coordinate embedding instead of three 128 x 128 layers
  if (!unfolded && !(a->flags & CPN_FLAG_NO_BILINEAR)) per_chunk -= 1;   // query_embed inside the layer-9 GEMM
  if (late && !(a->flags & (CPN_FLAG_NO_BILINEAR | CPN_FLAG_NO_GFOLD))) per_chunk -= 5;   // no round-1 readout + per-ray GEMM, encode_latent, z half of query_repeat_embed, park
  return chunks * per_chunk + (late ? 3 : 1);
}

namespace {
int render_chunk(const cpn_render_args& a, const Workspace& w, float* z_all, float* hbar_all, int ray0, int nr,
                 cudaStream_t st) {
    const float* W = reinterpret_cast<const float*>(a.weights);
    int rays = a.B * nr;
    int R = rays * 2 * a.S;
    CPN_TRY(launch_ray_setup(a, ray0, nr, w.seg, st));
    // Experiment, off by default (CPN_COMPACT_A=1): the encoder input image written compact (3 bytes per element: fp16 head +
    // remainder plane), the persistent GEMM derives the value plane e5m2(head) in shared memory. Measured 68 us faster in the
    // gather and 90-160 us slower in the GEMM, whose three-stage ring then waits for a conversion pass per stage; the GEMM is
    // not short of L2 -> SM bandwidth (copies alone run at 20 TB/s, profiles/r2_gemm1_mainloop_attribution.log). The hidden
    // image IS written compact (h1c below): its consumer, layer 10, has the slack.
    static int compact_a = -1;
    if (compact_a < 0) {
      const char* e = getenv("CPN_COMPACT_A");
      compact_a = (e && atoi(e) != 0) ? 1 : 0;
    }
    const bool ac = compact_a && a_form(a) == 2 && !(a.flags & CPN_FLAG_FULL_H1);
    const int aform = ac ? 3 : a_form(a);
    CPN_TRY(launch_sample(a, ray0, nr, w.seg, w.rowaux, w.local16, w.A, aform, st));
    CPN_TRY(launch_gather(a, ray0, nr, w.rowaux, w.A, aform, st, w.taps));
    const int Rp = (R + 127) / 128 * 128;
    const int KC832 = CPN_FEAT_DIM / ACT_BK, KC128 = CPN_HIDDEN / ACT_BK;
    const int sch = tc_scheme(a);
    // late readout (attention.cu): the attention reads out the hidden layer and the folded latent_value runs per ray
    const bool late_v = use_tc(a) && !(a.flags & (CPN_FLAG_NO_FOLD | CPN_FLAG_EARLY_V));
    const bool bilinear = use_tc(a) && !(a.flags & (CPN_FLAG_NO_FOLD | CPN_FLAG_NO_BILINEAR));
    // G fold (cpn_common.cuh, pw::WG): the per-ray bias of round 2 comes from a 128-wide per-row term next to the key
    // hidden layer, so there is no round-1 readout, and one readout with weights w2 + 2 w1 gives z
    const bool gfold = late_v && bilinear && !(a.flags & CPN_FLAG_NO_GFOLD);
    if (use_tc(a)) {
      // per-sample encoder, CoPoNeRF.py:387-397: (835 -> 832 ReLU -> 416) for the primary and the secondary rows.
      // Activations travel between the tensor-core layers as fp16 hi/lo operand images (cpn_common.cuh).
      // default path: the hidden image is only read by layer 10 (which derives the value plane on chip) and by the readout
      // (which never reads it), so it is written compact: 3 bytes per element instead of 4
      const bool h1c = gfold && a_form(a) == 2 && !(a.flags & CPN_FLAG_FULL_H1);
      {
        ProfScope prof(st);
        CPN_TRY(launch_gemm_tc(a.weights, 0, w.A, 0, w.H1, 0, 2 * Rp, 1,
                               CPN_TC_A_IMAGE | CPN_TC_OUT_IMAGE | sch | (h1c ? CPN_TC_OUT_IMAGE3 : 0) | (ac ? CPN_TC_A_IMAGE3 : 0), 1,
                               KC832, st));
      }
      if (a.flags & CPN_FLAG_NO_FOLD) {
        // the (tile, primary) and (tile, secondary) results land side by side: E image rows = sample rows, K = 832
        CPN_TRY(launch_gemm_tc(a.weights, 1, w.H1, 0, w.E, 0, 2 * Rp, 0, CPN_TC_A_IMAGE | CPN_TC_OUT_IMAGE | sch, 2, KC832, st));
        // value and key, CoPoNeRF.py:404-408
        CPN_TRY(launch_gemm_tc(a.weights, 2, w.E, 0, w.V, CPN_LATENT, R, 0, CPN_TC_A_IMAGE | sch, 1, 1, st));
        CPN_TRY(launch_gemm_tc(a.weights, 3, w.E, 0, w.K1, 0, R, 1, CPN_TC_A_IMAGE | CPN_TC_OUT_IMAGE | sch, 1, KC128, st));
      } else {
        // query_encode_latent_2 has no activation (CoPoNeRF.py:393-397), so it is folded into latent_value and
        // key_map at pack time: V = WVF [h_p ; h_s], K1 = relu(WKF [h_p ; h_s]) with K = 1664. The H1 image is
        // already that operand: a sample tile's primary and secondary hidden tiles are adjacent, 2 x 26 k-chunks.
        // 21 % fewer MACs than the three layers, and the E image is never written or read.
        if (!late_v)
          CPN_TRY(launch_gemm_tc(a.weights, 7, w.H1, 0, w.V, CPN_LATENT, R, 0, CPN_TC_A_IMAGE | sch, 1, 1, st));
      }
      if (bilinear) {
        // key_map_2, query_embed_2 and query_repeat_embed_2 have no activation, so both logits are bilinear forms of the
        // 128-wide hidden vectors (cpn_common.cuh, pw::WM12): one 128 -> 256 layer on the coordinate embedding for both rounds
        // instead of one on it and one on each key / repeat-query, and the key hidden layer is dotted in the epilogue
        // of the folded key_map GEMM without ever being stored.
        static int split_q = -1;   // CPN_SPLIT_QUERY_EMBED=1: query_embed as its own kernel writing an operand image (A/B runs)
        if (split_q < 0) {
          const char* e = getenv("CPN_SPLIT_QUERY_EMBED");
          split_q = (e && atoi(e) != 0) ? 1 : 0;
        }
        if (split_q) {
          CPN_TRY(launch_mlp16_image(w.local16, W + pw::WQT, W + pw::BQ, nullptr, 1, R, w.Q1, a_form(a) == 2, st, W + pw::WS1,
                                     w.s1, W + pw::WS2, w.s2));
          CPN_TRY(launch_gemm_tc(a.weights, 9, w.Q1, 0, w.Qm, 0, R, 0, CPN_TC_A_IMAGE | sch | CPN_TC_OUT_CB16, 1, 1, st));
        } else {   // query_embed computed by the producer warps of the layer-9 GEMM: the embedding never exists in memory
          CPN_TRY(launch_gemm_tc_mlp16(a.weights, w.local16, W + pw::WQT, W + pw::BQ, W + pw::WS1, w.s1, W + pw::WS2, w.s2, w.Qm, R,
                                       sch, st));
        }
        if (gfold)   // w.K1 (R, 128) receives G h + g0
          CPN_TRY(launch_gemm_tc(a.weights, 10, w.H1, 0, w.lg1, 0, R, 1,
                                 CPN_TC_A_IMAGE | sch | CPN_TC_OUT_KG | (h1c ? CPN_TC_A_IMAGE3 : 0), 1, 1, st, w.Qm,
                                 11.31f, w.s1, 2 * CPN_HIDDEN / 16, 0, w.K1));
        else
          CPN_TRY(launch_gemm_tc(a.weights, 8, w.H1, 0, w.lg1, 0, R, 1, CPN_TC_A_IMAGE | sch | CPN_TC_OUT_ROWDOT, 1, 1, st, w.Qm,
                                 11.31f, w.s1, 2 * CPN_HIDDEN / 16, 0));
      } else {
        if (!(a.flags & CPN_FLAG_NO_FOLD))
          CPN_TRY(launch_gemm_tc(a.weights, 8, w.H1, 0, w.K1, 0, R, 1, CPN_TC_A_IMAGE | CPN_TC_OUT_IMAGE | sch, 1, KC128, st));
        // coordinate embedding, CoPoNeRF.py:446; written column-blocked so the two logit epilogues read it coalesced
        CPN_TRY(launch_mlp16_image(w.local16, W + pw::WQT, W + pw::BQ, nullptr, 1, R, w.Q1, a_form(a) == 2, st));
        CPN_TRY(launch_gemm_tc(a.weights, 5, w.Q1, 0, w.Qe, 0, R, 0, CPN_TC_A_IMAGE | sch | CPN_TC_OUT_CB16, 1, 1, st));
        // key_map_2 with the round-1 logits <K, Q> / 11.31 (CoPoNeRF.py:450) as its epilogue: K itself is never stored
        CPN_TRY(launch_gemm_tc(a.weights, 4, w.K1, 0, w.lg1, 0, R, 0, CPN_TC_A_IMAGE | sch | CPN_TC_OUT_ROWDOT, 1, 1, st, w.Qe,
                               11.31f));
      }
    } else {
      {
        ProfScope prof(st);
        CPN_TRY(dense_simt(a, w.A, CPN_KA, pw::W1T, pw::B1, w.H1, CPN_FEAT_DIM, 2 * Rp, CPN_FEAT_DIM, CPN_KA, 1, st));
      }
      CPN_TRY(dense_simt(a, w.H1, CPN_FEAT_DIM, pw::W2T, pw::B2, w.E, CPN_FEAT_DIM, 2 * Rp, CPN_LATENT, CPN_FEAT_DIM, 0, st, 1));
      CPN_TRY(dense_simt(a, w.E, CPN_FEAT_DIM, pw::WVT, pw::BV, w.V, CPN_LATENT, R, CPN_LATENT, CPN_FEAT_DIM, 0, st));
      CPN_TRY(dense_simt(a, w.E, CPN_FEAT_DIM, pw::WKT, pw::BK, w.K1, CPN_HIDDEN, R, CPN_HIDDEN, CPN_FEAT_DIM, 1, st));
      CPN_TRY(dense_simt(a, w.K1, CPN_HIDDEN, pw::WK2T, pw::BK2, w.Kk, CPN_HIDDEN, R, CPN_HIDDEN, CPN_HIDDEN, 0, st));
      CPN_TRY(dense_simt(a, w.local16, 16, pw::WQT, pw::BQ, w.Q1, CPN_HIDDEN, R, CPN_HIDDEN, 16, 1, st));
      CPN_TRY(dense_simt(a, w.Q1, CPN_HIDDEN, pw::WQ2T, pw::BQ2, w.Qe, CPN_HIDDEN, R, CPN_HIDDEN, CPN_HIDDEN, 0, st));
    }
    if (gfold) {
      CPN_TRY(launch_attn1(a, ray0, nr, w.Kk, w.Qe, nullptr, w.rowaux, w.r1, w.wp, st, w.lg1, w.wt1, w.K1, w.rbias));
    } else if (late_v) {
      CPN_TRY(launch_attn1(a, ray0, nr, w.Kk, w.Qe, nullptr, w.rowaux, w.r1, w.wp, st, w.lg1, w.wt1));
      CPN_TRY(launch_readout_image(a, nr, w.H1, w.wt1, w.hbar, a_form(a) == 2, st));
      CPN_TRY(launch_gemm_tc(a.weights, 7, w.hbar, 0, w.r1, CPN_LATENT, rays, 0, CPN_TC_A_IMAGE | sch, 1, 1, st));
    } else {
      CPN_TRY(launch_attn1(a, ray0, nr, w.Kk, w.Qe, w.V, w.rowaux, w.r1, w.wp, st, use_tc(a) ? w.lg1 : nullptr));
    }
    // round 2, CoPoNeRF.py:467-473: query_repeat_embed(cat(encode_latent(R1), local_coords)); the z_embed
    // channels are the same for every sample of a ray, so they enter as a per-ray bias.
    if (!gfold) {
      CPN_TRY(dense_simt(a, w.r1, CPN_LATENT, pw::WET, pw::BE, w.zemb, CPN_HIDDEN, rays, CPN_HIDDEN, CPN_LATENT, 0, st));
      CPN_TRY(dense_simt(a, w.zemb, CPN_HIDDEN, pw::WQRA_T, pw::BQR, w.rbias, CPN_HIDDEN, rays, CPN_HIDDEN, CPN_HIDDEN, 0, st));
    }
    if (use_tc(a)) {   // query_repeat_embed_2 with the round-2 logits <Q2, Q> / 11.31 (CoPoNeRF.py:474) as its epilogue
      if (bilinear) {   // the repeat-query hidden layer is dotted with WM2 q + BM2 (second half of w.Qm) where it is produced
        CPN_TRY(launch_mlp16_image(w.local16, W + pw::WQRB_T, nullptr, w.rbias, 2 * a.S, R, nullptr, a_form(a) == 2, st,
                                   nullptr, nullptr, nullptr, nullptr, w.Qm, w.s2, 11.31f, w.lg2, 2 * CPN_HIDDEN / 16,
                                   CPN_HIDDEN / 16));
      } else {
        CPN_TRY(launch_mlp16_image(w.local16, W + pw::WQRB_T, nullptr, w.rbias, 2 * a.S, R, w.K1, a_form(a) == 2, st));
        CPN_TRY(launch_gemm_tc(a.weights, 6, w.K1, 0, w.lg2, 0, R, 0, CPN_TC_A_IMAGE | tc_scheme(a) | CPN_TC_OUT_ROWDOT, 1, 1,
                               st, w.Qe, 11.31f));
      }
    } else {
      CPN_TRY(launch_gemm_simt(w.local16, 16, W + pw::WQRB_T, nullptr, w.rbias, 2 * a.S, w.K1, CPN_HIDDEN, R, CPN_HIDDEN,
                               16, 1, st));
      CPN_TRY(dense_simt(a, w.K1, CPN_HIDDEN, pw::WQR2T, pw::BQR2, w.Kk, CPN_HIDDEN, R, CPN_HIDDEN, CPN_HIDDEN, 0, st));
    }
    if (gfold) {
      // combined weight w2 + 2 w1 -> one readout into the image-level operand image; z is finished per image
      CPN_TRY(launch_attn2(a, ray0, nr, w.Kk, w.Qe, nullptr, w.r1, z_all, st, w.lg2, w.wt2, w.wt1));
      const bool h1c = a_form(a) == 2 && !(a.flags & CPN_FLAG_FULL_H1);
      CPN_TRY(launch_readout_image(a, nr, w.H1, w.wt2, hbar_all, a_form(a) == 2, st, a.N, ray0, h1c ? ACT_X8 : ACT_CHUNK_BYTES));
    } else if (late_v) {
      CPN_TRY(launch_attn2(a, ray0, nr, w.Kk, w.Qe, nullptr, w.r1, z_all, st, w.lg2, w.wt2));
      // the round-2 readout lands in the image-level operand image; its GEMM runs once per image (finish_image)
      CPN_TRY(launch_readout_image(a, nr, w.H1, w.wt2, hbar_all, a_form(a) == 2, st, a.N, ray0));
      CPN_TRY(launch_park_r1(a, ray0, nr, w.r1, z_all, st));
    } else {
      CPN_TRY(launch_attn2(a, ray0, nr, w.Kk, w.Qe, w.V, w.r1, z_all, st, use_tc(a) ? w.lg2 : nullptr));
    }
    CPN_TRY(launch_ray_epilogue(a, ray0, nr, w.wp, w.seg, st));
    return CPN_OK;
}
}  // namespace

extern "C" int cpn_render_rays(const cpn_render_args* args, void* stream) {
  CPN_TRY(check_args(args));
  const cpn_render_args& a = *args;
  cudaStream_t st = (cudaStream_t)stream;
  if (a.N == 0) return CPN_OK;
  const int chunk = a.chunk_rays < a.N ? a.chunk_rays : a.N;
  const int nchunks = (a.N + chunk - 1) / chunk;
  int lanes = a.lanes < 1 ? 1 : a.lanes;
  if (lanes > MAX_LANES) lanes = MAX_LANES;
  if (lanes > nchunks) lanes = nchunks;
  const size_t lane_bytes = carve(nullptr, a.B, chunk, a.S, a.flags).bytes, img_bytes = image_bytes(a.B, a.N);
  if (img_bytes + lane_bytes * lanes > a.workspace_bytes) {
    cpn_set_error("cpn_render_rays: workspace of %zu bytes needed (%d lanes), %zu given", img_bytes + lane_bytes * lanes,
                  lanes, a.workspace_bytes);
    return CPN_ERR_WORKSPACE;
  }
  float* z_all = reinterpret_cast<float*>(a.workspace);
  float* r2_all = reinterpret_cast<float*>(reinterpret_cast<char*>(a.workspace) + z_bytes(a.B, a.N));
  float* hbar_all = reinterpret_cast<float*>(reinterpret_cast<char*>(a.workspace) + 2 * z_bytes(a.B, a.N));
  const bool late_v = use_tc(a) && !(a.flags & (CPN_FLAG_NO_FOLD | CPN_FLAG_EARLY_V));
  const bool gfold = late_v && !(a.flags & (CPN_FLAG_NO_BILINEAR | CPN_FLAG_NO_GFOLD));
  // late readout: R2 of every ray from one GEMM over the image, z = (R2 + R1) + R1, then the light-field decoder
  auto finish_image = [&]() -> int {
    if (late_v) {
      CPN_TRY(launch_gemm_tc(a.weights, 7, hbar_all, 0, r2_all, CPN_LATENT, a.B * a.N, 0, CPN_TC_A_IMAGE | tc_scheme(a), 1, 1, st));
      if (gfold)   // the operand image held sum (w2 + 2 w1) h: z = WVF hbar + 3 b
        CPN_TRY(launch_finish_z_bias(a, r2_all, reinterpret_cast<const float*>(a.weights) + pw::BVF, z_all, st));
      else
        CPN_TRY(launch_finish_z(a, r2_all, z_all, st));
    }
    return launch_phi(a, z_all, st);
  };
  char* lane_base = reinterpret_cast<char*>(a.workspace) + img_bytes;
  if (lanes == 1) {
    Workspace w = carve(lane_base, a.B, chunk, a.S, a.flags);
    for (int ray0 = 0; ray0 < a.N; ray0 += chunk)
      CPN_TRY(render_chunk(a, w, z_all, hbar_all, ray0, (a.N - ray0) < chunk ? (a.N - ray0) : chunk, st));
    return finish_image();
  }
  LanePool* pool = nullptr;
  CPN_TRY(get_pool(&pool));
  CPN_CHECK_CUDA(cudaEventRecord(pool->fork, st));
  Workspace w[MAX_LANES];
  for (int l = 0; l < lanes; ++l) {
    w[l] = carve(lane_base + l * lane_bytes, a.B, chunk, a.S, a.flags);
    CPN_CHECK_CUDA(cudaStreamWaitEvent(pool->stream[l], pool->fork, 0));
  }
  int status = CPN_OK;
  for (int c = 0; c < nchunks && status == CPN_OK; ++c) {
    int ray0 = c * chunk;
    status = render_chunk(a, w[c % lanes], z_all, hbar_all, ray0, (a.N - ray0) < chunk ? (a.N - ray0) : chunk,
                          pool->stream[c % lanes]);
  }
  for (int l = 0; l < lanes; ++l) {   // always join, also after an error, so the caller's stream stays ordered
    cudaEventRecord(pool->done[l], pool->stream[l]);
    cudaStreamWaitEvent(st, pool->done[l], 0);
  }
  if (status != CPN_OK) return status;
  return finish_image();   // the light-field decoder runs once over every ray of the image
}
