// Per-pair setup, per-ray epipolar segment and per-sample geometry.
// Compiled with -fmad=false: the reference evaluates these expressions op by op in fp32
// (and fp64 for the triangulation), so no contraction into FMAs here.
#include <math.h>
#include "cpn_common.cuh"
#include "tc_common.cuh"

namespace {

// ---------------------------------------------------------------- small dense helpers
__device__ void mat4_mul(const double* a, const double* b, double* c) {
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) {
      double s = 0.0;
      for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
      c[i * 4 + j] = s;
    }
}

// Gauss-Jordan with partial pivoting, n <= 4.
__device__ void mat_inv(const double* a, double* out, int n) {
  double m[4][8];
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) {
      m[i][j] = a[i * n + j];
      m[i][n + j] = (i == j) ? 1.0 : 0.0;
    }
  for (int c = 0; c < n; ++c) {
    int p = c;
    for (int r = c + 1; r < n; ++r)
      if (fabs(m[r][c]) > fabs(m[p][c])) p = r;
    if (p != c)
      for (int j = 0; j < 2 * n; ++j) {
        double t = m[c][j];
        m[c][j] = m[p][j];
        m[p][j] = t;
      }
    double d = 1.0 / m[c][c];
    for (int j = 0; j < 2 * n; ++j) m[c][j] *= d;
    for (int r = 0; r < n; ++r)
      if (r != c) {
        double f = m[r][c];
        if (f != 0.0)
          for (int j = 0; j < 2 * n; ++j) m[r][j] -= f * m[c][j];
      }
  }
  for (int i = 0; i < n; ++i)
    for (int j = 0; j < n; ++j) out[i * n + j] = m[i][n + j];
}

__device__ void load16(const float* src, double* dst) {
  for (int i = 0; i < 16; ++i) dst[i] = (double)src[i];
}
__device__ void store16(const double* src, float* dst) {
  for (int i = 0; i < 16; ++i) dst[i] = (float)src[i];
}

// models/CoPoNeRF.py:239-244,259-261,325-332,572-575 and utils.pose_inverse_4x4 (utils.py:111-138)
__global__ void pair_setup_kernel(const float* __restrict__ ctx_c2w, const float* __restrict__ ctx_K,
                                  const float* __restrict__ qry_c2w, const float* __restrict__ qry_K,
                                  const float* __restrict__ rel_pose, int B, int H, int val,
                                  float* __restrict__ consts) {
  int b = blockIdx.x * blockDim.x + threadIdx.x;
  if (b >= B) return;
  float* o = consts + (size_t)b * CPN_PAIR_CONSTS_FLOATS;
  double c0[16], c1[16], qc[16], rel[16], i0[16], i1[16], t[16], u[16], flip[16];
  load16(ctx_c2w + (size_t)b * 32, c0);
  load16(ctx_c2w + (size_t)b * 32 + 16, c1);
  load16(qry_c2w + (size_t)b * 16, qc);
  load16(rel_pose + (size_t)b * 16, rel);
  mat_inv(c0, i0, 4);
  mat_inv(c1, i1, 4);
  // pose_inverse_4x4: [R^T | -R^T t]
  for (int i = 0; i < 16; ++i) flip[i] = 0.0;
  for (int i = 0; i < 3; ++i) {
    double s = 0.0;
    for (int j = 0; j < 3; ++j) {
      flip[i * 4 + j] = rel[j * 4 + i];
      s += -rel[j * 4 + i] * rel[j * 4 + 3];
    }
    flip[i * 4 + 3] = s;
  }
  flip[15] = 1.0;
  // query camera in each context frame
  mat4_mul(i0, qc, t);
  store16(t, o + pc::Q_C2W);
  if (val) mat4_mul(flip, t, u); else mat4_mul(i1, qc, u);
  store16(u, o + pc::Q_C2W + 16);
  for (int i = 0; i < 16; ++i) o[pc::KQ + i] = qry_K[(size_t)b * 16 + i];
  for (int v = 0; v < 2; ++v) {
    const float* K = ctx_K + (size_t)b * 32 + v * 16;
    for (int i = 0; i < 16; ++i) o[pc::KC + v * 16 + i] = K[i];
    for (int r = 0; r < 3; ++r)
      for (int c = 0; c < 3; ++c) o[pc::KN + v * 9 + r * 3 + c] = (r < 2) ? K[r * 4 + c] / (float)H : K[r * 4 + c];
  }
  double id0[16], id1[16];
  mat4_mul(i0, c0, id0);
  mat4_mul(i1, c1, id1);
  store16(id0, o + pc::IDEN);
  store16(id1, o + pc::IDEN + 16);
  store16(id0, o + pc::T_OWN);
  store16(id1, o + pc::T_OWN + 16);
  if (val) {
    store16(flip, o + pc::T_OTHER);      // view-0 samples -> view-1 frame
    store16(rel, o + pc::T_OTHER + 16);  // view-1 samples -> view-0 frame
  } else {
    mat4_mul(i1, c0, t);
    store16(t, o + pc::T_OTHER);
    mat4_mul(i0, c1, t);
    store16(t, o + pc::T_OTHER + 16);
  }
  mat_inv(qc, t, 4);
  store16(t, o + pc::INV_QC2W);
  double k3[9], k3i[9];
  for (int r = 0; r < 3; ++r)
    for (int c = 0; c < 3; ++c) k3[r * 3 + c] = (double)qry_K[(size_t)b * 16 + r * 4 + c];
  mat_inv(k3, k3i, 3);
  for (int i = 0; i < 9; ++i) o[pc::INV_KQ3 + i] = (float)k3i[i];
  store16(flip, o + pc::REL_FLIP);
  mat4_mul(i0, c1, t);
  store16(t, o + pc::GT_REL);
  mat4_mul(i1, c0, t);
  mat_inv(t, u, 4);
  store16(u, o + pc::GT_REL_FLIP);
}

// ---------------------------------------------------------------- per-pair prologue
// F.interpolate(x, 256, mode='bilinear') at one output position (align_corners=False).
__device__ __forceinline__ void interp_coef(int dst, int in, float scale, int& i0, int& i1, float& l0, float& l1) {
  float src = scale * ((float)dst + 0.5f) - 0.5f;
  if (src < 0.f) src = 0.f;
  i0 = (int)src;
  i1 = i0 + ((i0 < in - 1) ? 1 : 0);
  l1 = src - (float)i0;
  l0 = 1.f - l1;
}
__device__ __forceinline__ float upsample_at(const float* __restrict__ f, int fh, int fw, int y, int x) {
  int y0, y1, x0, x1;
  float ly0, ly1, lx0, lx1;
  interp_coef(y, fh, (float)fh / 256.f, y0, y1, ly0, ly1);
  interp_coef(x, fw, (float)fw / 256.f, x0, x1, lx0, lx1);
  return ly0 * (lx0 * f[y0 * fw + x0] + lx1 * f[y0 * fw + x1]) + ly1 * (lx0 * f[y1 * fw + x0] + lx1 * f[y1 * fw + x1]);
}

// models/CoPoNeRF.py:230-236 with utils.warp (utils.py:642-670) and get_gt_correspondence_mask (utils.py:576-602)
__global__ void pair_prologue_kernel(const float* __restrict__ flow0, const float* __restrict__ flow1, int B, int fh,
                                     int fw, float scale, float* __restrict__ up_flow2, uint8_t* __restrict__ mask2) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * 65536) return;
  int b = idx >> 16, y = (idx >> 8) & 255, x = idx & 255;
  const float* f0x = flow0 + (size_t)b * 2 * fh * fw;
  const float* f0y = f0x + fh * fw;
  const float* f1x = flow1 + (size_t)b * 2 * fh * fw;
  const float* f1y = f1x + fh * fw;
  float rx = upsample_at(f1x, fh, fw, y, x), ry = upsample_at(f1y, fh, fw, y, x);
  up_flow2[((size_t)b * 2 + 0) * 65536 + y * 256 + x] = rx;
  up_flow2[((size_t)b * 2 + 1) * 65536 + y * 256 + x] = ry;
  float u2x = rx * scale, u2y = ry * scale;
  // warp(up_flow, up_flow2): grid_sample(up_flow, grid + up_flow2), zeros padding, align_corners=False
  float gx = 2.0f * ((float)x + u2x) / 255.f - 1.0f;
  float gy = 2.0f * ((float)y + u2y) / 255.f - 1.0f;
  float ix = ((gx + 1.f) * 256.f - 1.f) / 2.f, iy = ((gy + 1.f) * 256.f - 1.f) / 2.f;
  float fx0 = floorf(ix), fy0 = floorf(iy);
  float wx = 0.f, wy = 0.f;
  for (int t = 0; t < 4; ++t) {
    float tx = fx0 + (float)(t & 1), ty = fy0 + (float)(t >> 1);
    float wgt = ((t & 1) ? (ix - fx0) : (fx0 + 1.f - ix)) * ((t >> 1) ? (iy - fy0) : (fy0 + 1.f - iy));
    if (tx >= 0.f && tx <= 255.f && ty >= 0.f && ty <= 255.f) {
      int xi = (int)tx, yi = (int)ty;
      wx += upsample_at(f0x, fh, fw, yi, xi) * scale * wgt;
      wy += upsample_at(f0y, fh, fw, yi, xi) * scale * wgt;
    }
  }
  float ex = u2x + wx, ey = u2y + wy;
  bool cyc = sqrtf(ex * ex + ey * ey) <= 10.f;
  float mx = u2x + (float)x, my = u2y + (float)y;
  bool inside = mx >= 0.f && mx <= 255.f && my >= 0.f && my <= 255.f;
  mask2[(size_t)b * 65536 + y * 256 + x] = (cyc && inside) ? 1 : 0;
}

// ---------------------------------------------------------------- epipolar segment (models/epipolar.py)
struct Hit {
  float t, x, y;
  bool valid;
};

__device__ __forceinline__ bool in_bounds(float x, float y) {
  const float eps = 1e-6f;
  return (x >= -eps) && (y >= -eps) && (x <= 1.f + eps) && (y <= 1.f + eps);
}

// epipolar.py:74-122
__device__ __forceinline__ Hit edge_hit(const float* Kn, const float* o, const float* d, int dim, float value) {
  int od = 1 - dim;
  float fs = Kn[dim * 3 + dim], fo = Kn[od * 3 + od], cs = Kn[dim * 3 + 2], co = Kn[od * 3 + 2];
  float os = o[dim], oo = o[od], ds = d[dim], dd = d[od], oz = o[2], dz = d[2];
  float c = (value - cs) / fs;
  float t = (c * oz - os) / (ds - c * dz);
  float other = co + fo * (oo * (c * dz - ds) + dd * (os - c * oz)) / (dz * os - ds * oz);
  float same = 1.0f * value;
  Hit h;
  h.t = t;
  h.x = dim == 0 ? same : other;
  h.y = dim == 0 ? other : same;
  float z = o[2] + t * d[2];
  h.valid = in_bounds(h.x, h.y) && (z > -1e-6f);
  return h;
}

// epipolar.py:23-26,152-162
__device__ __forceinline__ void point_projection(const float* Kn, const float* p, float& x, float& y, bool& valid) {
  float den = p[2] + 1e-8f;
  float q0 = p[0] / den, q1 = p[1] / den, q2 = p[2] / den;
  x = Kn[0] * q0 + Kn[1] * q1 + Kn[2] * q2;
  y = Kn[3] * q0 + Kn[4] * q1 + Kn[5] * q2;
  valid = in_bounds(x, y) && (p[2] > -1e-6f);
}

__device__ __forceinline__ float zero_nonfinite(float v) { return (isnan(v) || isinf(v)) ? 0.f : v; }

// One thread per (pair b, ray n, view v). geometry.plucker_embedding (geometry.py:236-245),
// project_rays (epipolar.py:175-253), start/end sanitising (CoPoNeRF.py:279-285).
__global__ void ray_setup_kernel(cpn_render_args a, int ray0, int nr, float* __restrict__ seg) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * nr * 2) return;
  int v = idx & 1, nl = (idx >> 1) % nr, b = (idx >> 1) / nr;
  int n = ray0 + nl;
  const float* cst = a.pair_consts + (size_t)b * CPN_PAIR_CONSTS_FLOATS;
  const float* c2w = cst + pc::Q_C2W + v * 16;
  const float* Kq = cst + pc::KQ;
  float ux = a.uv[((size_t)b * a.N + n) * 2 + 0], uy = a.uv[((size_t)b * a.N + n) * 2 + 1];
  float xl = (ux - Kq[2]) / Kq[0] * 1.0f, yl = (uy - Kq[6]) / Kq[5] * 1.0f;
  float w[3], org[3], dir[3];
  for (int i = 0; i < 3; ++i) {
    w[i] = c2w[i * 4 + 0] * xl + c2w[i * 4 + 1] * yl + c2w[i * 4 + 2] * 1.0f + c2w[i * 4 + 3] * 1.0f;
    org[i] = c2w[i * 4 + 3];
  }
  float r0 = w[0] - org[0], r1 = w[1] - org[1], r2 = w[2] - org[2];
  float nrm = fmaxf(sqrtf(r0 * r0 + r1 * r1 + r2 * r2), 1e-12f);
  dir[0] = r0 / nrm;
  dir[1] = r1 / nrm;
  dir[2] = r2 / nrm;
  float* co = a.coords + (((size_t)(b * 2 + v)) * a.N + n) * 9;
  co[0] = dir[0];
  co[1] = dir[1];
  co[2] = dir[2];
  co[3] = org[1] * dir[2] - org[2] * dir[1];
  co[4] = org[2] * dir[0] - org[0] * dir[2];
  co[5] = org[0] * dir[1] - org[1] * dir[0];
  co[6] = org[0];
  co[7] = org[1];
  co[8] = org[2];

  const float* Kn = cst + pc::KN + v * 9;
  Hit h[4] = {edge_hit(Kn, org, dir, 0, 0.f), edge_hit(Kn, org, dir, 0, 1.f), edge_hit(Kn, org, dir, 1, 0.f),
              edge_hit(Kn, org, dir, 1, 1.f)};
  // epipolar.py:125-149: invalid hits get the lowest priority; the first index wins ties
  int smin = 0, smax = 0;
  float tmin = h[0].valid ? h[0].t : INFINITY, tmax = h[0].valid ? h[0].t : -INFINITY;
  for (int i = 1; i < 4; ++i) {
    float ti = h[i].valid ? h[i].t : INFINITY, ta = h[i].valid ? h[i].t : -INFINITY;
    if (ti < tmin) { tmin = ti; smin = i; }
    if (ta > tmax) { tmax = ta; smax = i; }
  }
  const float eps = 1e-6f;
  bool depth_zero = org[2] < eps;
  bool at_cam = sqrtf(org[0] * org[0] + org[1] * org[1] + org[2] * org[2]) < eps;
  float zx, zy, ix, iy;
  bool zok, iok;
  point_projection(Kn, at_cam ? dir : org, zx, zy, zok);
  if (depth_zero && !at_cam) zok = false;
  point_projection(Kn, dir, ix, iy, iok);
  float x0 = zok ? zx : h[smin].x, y0 = zok ? zy : h[smin].y;
  float x1 = iok ? ix : h[smax].x, y1 = iok ? iy : h[smax].y;
  bool overlaps = (zok || h[smin].valid) && (iok || h[smax].valid);
  float* sg = seg + (((size_t)b * nr + nl) * 2 + v) * 6;
  sg[0] = zero_nonfinite((x0 - 0.5f) * 2.f);
  sg[1] = zero_nonfinite((y0 - 0.5f) * 2.f);
  sg[2] = zero_nonfinite((x1 - 0.5f) * 2.f);
  sg[3] = zero_nonfinite((y1 - 0.5f) * 2.f);
  sg[4] = overlaps ? 1.f : 0.f;
  sg[5] = 0.f;
}

// ---------------------------------------------------------------- per-sample geometry
__device__ __forceinline__ void cross3(const double* a, const double* b, double* c) {
  c[0] = a[1] * b[2] - a[2] * b[1];
  c[1] = a[2] * b[0] - a[0] * b[2];
  c[2] = a[0] * b[1] - a[1] * b[0];
}
__device__ __forceinline__ float nan_to_num(float v) {
  if (isnan(v)) return 0.f;
  if (isinf(v)) return v > 0.f ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return v;
}
__device__ __forceinline__ void transform_point(const float* T, const float* p, float* o) {
  for (int i = 0; i < 3; ++i) o[i] = ((p[0] * T[i * 4 + 0] + p[1] * T[i * 4 + 1]) + p[2] * T[i * 4 + 2]) + 1.0f * T[i * 4 + 3];
}

// One thread per (b, n, v, s). CoPoNeRF.py:304-309 (sample positions), :420 / geometry.py:98-162
// (fp64 closest point), :336-367 (reprojection into the other view), :384-394 (tanh point codes),
// :411-445 (local_coords).
__global__ void sample_kernel(cpn_render_args a, int ray0, int nr, const float* __restrict__ seg,
                              float* __restrict__ rowaux, float* __restrict__ local16, float* __restrict__ A, int a_image) {
  long long row = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  long long nrows = (long long)a.B * nr * 2 * a.S;
  if (row >= nrows) return;
  int S = a.S;
  int s = (int)(row % S);
  int v = (int)((row / S) & 1);
  int nl = (int)((row / (2 * S)) % nr);
  int b = (int)(row / ((long long)2 * S * nr));
  int n = ray0 + nl;
  const float* cst = a.pair_consts + (size_t)b * CPN_PAIR_CONSTS_FLOATS;
  const float* sg = seg + (((size_t)b * nr + nl) * 2 + v) * 6;
  float t = a.interval[s];
  float pvx = sg[0] + (sg[2] - sg[0]) * t, pvy = sg[1] + (sg[3] - sg[1]) * t;
  size_t m = (size_t)b * 2 + v;
  float* pv = a.pixel_val + ((m * a.N + n) * S + s) * 2;
  pv[0] = pvx;
  pv[1] = pvy;

  // context pixel ray through the sample (geometry.py:100-109)
  const float* Kc = cst + pc::KC + v * 16;
  float px = (pvx + 1.f) / 2.f * (float)(a.W - 1), py = (pvy + 1.f) / 2.f * (float)(a.H - 1);
  float xl = (px - Kc[2]) / Kc[0] * 1.0f, yl = (py - Kc[6]) / Kc[5] * 1.0f;
  const float* Id = cst + pc::IDEN + v * 16;
  float cw[3], cp[3], l2f[3];
  for (int i = 0; i < 3; ++i) {
    cw[i] = Id[i * 4 + 0] * xl + Id[i * 4 + 1] * yl + Id[i * 4 + 2] * 1.0f + Id[i * 4 + 3] * 1.0f;
    cp[i] = Id[i * 4 + 3];
  }
  {
    float r0 = cw[0] - cp[0], r1 = cw[1] - cp[1], r2 = cw[2] - cp[2];
    float nrm = fmaxf(sqrtf(r0 * r0 + r1 * r1 + r2 * r2), 1e-12f);
    l2f[0] = r0 / nrm;
    l2f[1] = r1 / nrm;
    l2f[2] = r2 / nrm;
  }
  float m2f[3] = {cp[1] * l2f[2] - cp[2] * l2f[1], cp[2] * l2f[0] - cp[0] * l2f[2], cp[0] * l2f[1] - cp[1] * l2f[0]};
  const float* co = a.coords + ((m * a.N) + n) * 9;  // query plucker ray + origin
  double l1[3] = {co[0], co[1], co[2]}, m1[3] = {co[3], co[4], co[5]};
  double l2[3] = {l2f[0], l2f[1], l2f[2]}, m2[3] = {m2f[0], m2f[1], m2f[2]};
  double c12[3], c2c[3], mc[3];
  cross3(l1, l2, c12);
  cross3(l2, c12, c2c);
  cross3(m1, c2c, mc);
  double dot = (m2[0] * c12[0] + m2[1] * c12[1]) + m2[2] * c12[2];
  double nn = sqrt((c12[0] * c12[0] + c12[1] * c12[1]) + c12[2] * c12[2]);
  double den = nn * nn + 1e-12;
  float pt[3];
  for (int i = 0; i < 3; ++i) {
    double p = (-mc[i] + dot * l1[i]) / den;
    if (isnan(p) || isinf(p)) p = 0.0;
    pt[i] = (float)p;
  }

  float own[3], oth[3];
  transform_point(cst + pc::T_OWN + v * 16, pt, own);
  transform_point(cst + pc::T_OTHER + v * 16, pt, oth);
  // geometry.project (geometry.py:374-393) with the other view's intrinsics, then
  // utils.normalize_for_grid_sample (utils.py:242-245)
  const float* Ko = cst + pc::KC + (1 - v) * 16;
  float gx = Ko[0] * oth[0] / (oth[2] + 1e-12f) + Ko[2];
  float gy = Ko[5] * oth[1] / (oth[2] + 1e-12f) + Ko[6];
  if (isnan(gx) || isinf(gx)) gx = 1e10f;
  if (isnan(gy) || isinf(gy)) gy = 1e10f;
  gx = (gx / (float)(a.W - 1)) * 2.f - 1.f;
  gy = (gy / (float)(a.H - 1)) * 2.f - 1.f;

  float* ra = rowaux + (size_t)row * CPN_ROWAUX;
  ra[0] = pvx;
  ra[1] = pvy;
  ra[2] = gx;
  ra[3] = gy;
  ra[4] = fminf(fmaxf(pt[0], -100.f), 100.f);
  ra[5] = fminf(fmaxf(pt[1], -100.f), 100.f);
  ra[6] = fminf(fmaxf(pt[2], -100.f), 100.f);
  ra[7] = 0.f;

  // tanh point codes go straight into the encoder input rows (columns 832..834, zero padding after)
  float tp[3], ts[3];
  for (int i = 0; i < 3; ++i) {
    tp[i] = tanhf(nan_to_num(own[i]) / 5.f);
    ts[i] = tanhf(nan_to_num(oth[i]) / 5.f);
  }
  if (a_image) {
    // operand image: k-chunk 26 holds k = 832..863; group 0 = the three codes + five zeros, groups 1-3 zeros
    unsigned char* img = reinterpret_cast<unsigned char*>(A);
    const int r = (int)(row & 127);
    for (int br = 0; br < 2; ++br) {
      const float* t3 = br ? ts : tp;
      // a_image 3: compact f8 image, 12 KB blocks without the value plane (the consuming GEMM derives it on chip)
      const size_t chunk = a_image == 3 ? (size_t)ACT_X8 : (size_t)ACT_CHUNK_BYTES;
      unsigned char* p = img + (((size_t)(row >> 7) * 2 + br) * (CPN_KA_IMG / ACT_BK) + CPN_FEAT_DIM / ACT_BK) * chunk + (size_t)r * 16;
      const uint4 zero = make_uint4(0, 0, 0, 0);
      if (a_image >= 2) {   // f8 scheme: fp16 hi in group 0; the 8-bit planes have two 16-k groups
        uint2 hi;
        uint32_t l8, x8;
        tc::split4_f8(make_float4(t3[0], t3[1], t3[2], 0.f), hi, l8, x8);
        *reinterpret_cast<uint4*>(p) = make_uint4(hi.x, hi.y, 0, 0);
        for (int gq = 1; gq < 4; ++gq) *reinterpret_cast<uint4*>(p + gq * 2048) = zero;
        *reinterpret_cast<uint4*>(p + ACT_LO8) = make_uint4(l8, 0, 0, 0);
        *reinterpret_cast<uint4*>(p + ACT_LO8 + 2048) = zero;
        if (a_image == 2) {
          *reinterpret_cast<uint4*>(p + ACT_X8) = make_uint4(x8, 0, 0, 0);
          *reinterpret_cast<uint4*>(p + ACT_X8 + 2048) = zero;
        }
      } else {
        uint4 hi = make_uint4(0, 0, 0, 0), lo = make_uint4(0, 0, 0, 0);
        tc::split2(t3[0], t3[1], hi.x, lo.x);
        tc::split2(t3[2], 0.f, hi.y, lo.y);
        *reinterpret_cast<uint4*>(p) = hi;
        *reinterpret_cast<uint4*>(p + ACT_LO) = lo;
        for (int gq = 1; gq < 4; ++gq) {
          *reinterpret_cast<uint4*>(p + gq * 2048) = zero;
          *reinterpret_cast<uint4*>(p + gq * 2048 + ACT_LO) = zero;
        }
      }
    }
  } else {
    float* Ap = A + enc_row((size_t)row, 0) * CPN_KA + CPN_FEAT_DIM;
    float* As = A + enc_row((size_t)row, 1) * CPN_KA + CPN_FEAT_DIM;
    for (int i = 0; i < 3; ++i) {
      Ap[i] = tp[i];
      As[i] = ts[i];
    }
    for (int i = 3; i < CPN_KA - CPN_FEAT_DIM; ++i) {
      Ap[i] = 0.f;
      As[i] = 0.f;
    }
  }

  // local_coords (CoPoNeRF.py:411-445): [cam ray dir, 0 0 0, query ray dir, tanh(d / {1,10,100,1000}), query origin]
  float cr0 = (px - Kc[2]) / Kc[0] * 1.0f, cr1 = (py - Kc[6]) / Kc[5] * 1.0f, cr2 = 1.0f;
  float cn = fmaxf(sqrtf(cr0 * cr0 + cr1 * cr1 + cr2 * cr2), 1e-12f);
  float d0 = pt[0] - co[6], d1 = pt[1] - co[7], d2 = pt[2] - co[8];
  float depth = sqrtf(d0 * d0 + d1 * d1 + d2 * d2);
  if (isnan(depth) || isinf(depth)) depth = 1000000.f;
  float* lc = local16 + (size_t)row * 16;
  lc[0] = cr0 / cn;
  lc[1] = cr1 / cn;
  lc[2] = cr2 / cn;
  lc[3] = 0.f;
  lc[4] = 0.f;
  lc[5] = 0.f;
  lc[6] = co[0];
  lc[7] = co[1];
  lc[8] = co[2];
  lc[9] = tanhf(depth);
  lc[10] = tanhf(depth / 10.f);
  lc[11] = tanhf(depth / 100.f);
  lc[12] = tanhf(depth / 1000.f);
  lc[13] = co[6];
  lc[14] = co[7];
  lc[15] = co[8];
}


// ---------------------------------------------------------------- per-ray auxiliary outputs
// Tensor.long() on the host reference is an x86 cvttss2si: out-of-range and NaN inputs give INT64_MIN.
__device__ __forceinline__ long long float_to_long(float f) {
  if (isnan(f) || f >= 9223372036854775808.f || f < -9223372036854775808.f) return (long long)0x8000000000000000ull;
  return (long long)f;
}

// utils.batch_project_to_other_img (utils.py:140-170) for one pixel.
__device__ __forceinline__ void project_to_other(const float* invK3, const float* T, const float* K4, float u, float v,
                                                 float depth, float& ox, float& oy) {
  float h[3], g[4], p3[3], p2[3];
  for (int i = 0; i < 3; ++i) h[i] = ((u * invK3[i * 3 + 0] + v * invK3[i * 3 + 1]) + 1.0f * invK3[i * 3 + 2]) * depth;
  for (int i = 0; i < 4; ++i) g[i] = ((h[0] * T[i * 4 + 0] + h[1] * T[i * 4 + 1]) + h[2] * T[i * 4 + 2]) + 1.0f * T[i * 4 + 3];
  for (int i = 0; i < 3; ++i) p3[i] = g[i] / (g[3] + 1e-6f);
  for (int i = 0; i < 3; ++i) p2[i] = (p3[0] * K4[i * 4 + 0] + p3[1] * K4[i * 4 + 1]) + p3[2] * K4[i * 4 + 2];
  ox = p2[0] / (p2[2] + 1e-6f);
  oy = p2[1] / (p2[2] + 1e-6f);
}

// One thread per (pair, ray). models/CoPoNeRF.py:499-542, utils.flow2kps (utils.py:52-69).
__global__ void ray_epilogue_kernel(cpn_render_args a, int ray0, int nr, const float* __restrict__ wp,
                                    const float* __restrict__ seg) {
  int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= a.B * nr) return;
  int b = idx / nr, n = ray0 + idx % nr;
  const float* cst = a.pair_consts + (size_t)b * CPN_PAIR_CONSTS_FLOATS;
  const float* inv = cst + pc::INV_QC2W;
  float p0 = wp[(size_t)idx * 4 + 0], p1 = wp[(size_t)idx * 4 + 1], p2 = wp[(size_t)idx * 4 + 2];
  float depth = ((inv[8] * p0 + inv[9] * p1) + inv[10] * p2) + inv[11] * 1.0f;
  size_t o = (size_t)b * a.N + n;
  float u = a.uv[o * 2 + 0], v = a.uv[o * 2 + 1];
  float c1x, c1y, c2x, c2y;
  project_to_other(cst + pc::INV_KQ3, cst + pc::Q_C2W, cst + pc::KC, u, v, depth, c1x, c1y);
  project_to_other(cst + pc::INV_KQ3, cst + pc::Q_C2W + 16, cst + pc::KC + 16, u, v, depth, c2x, c2y);
  a.T_to_C1_pts[o * 2 + 0] = c1x;
  a.T_to_C1_pts[o * 2 + 1] = c1y;
  a.T_to_C2_pts[o * 2 + 0] = c2x;
  a.T_to_C2_pts[o * 2 + 1] = c2y;
  long long rx = float_to_long(c2x), ry = float_to_long(c2y);
  int kx = (int)(rx < 0 ? 0 : (rx > 255 ? 255 : rx)), ky = (int)(ry < 0 ? 0 : (ry > 255 ? 255 : ry));
  a.mask_c2[o] = (rx >= 0 && rx < 256 && ry >= 0 && ry < 256) ? 1 : 0;
  a.matchability_cycle_mask[o] = a.mask_padded2[(size_t)b * 65536 + ky * 256 + kx];
  float fs = (float)(256.0 / (double)a.flow_h);
  a.C2_pts_to_C1[o * 2 + 0] = (float)kx + a.up_flow2[((size_t)b * 2 + 0) * 65536 + ky * 256 + kx] * fs;
  a.C2_pts_to_C1[o * 2 + 1] = (float)ky + a.up_flow2[((size_t)b * 2 + 1) * 65536 + ky * 256 + kx] * fs;
  a.depth_ray[o] = depth < 0.f ? 0.f : (depth > 10.f ? 10.f : depth);
  // valid = any view has an epipolar segment inside its image (CoPoNeRF.py:562); phi whites out the rest
  const float* sg = seg + (size_t)idx * 2 * 6;
  a.valid_mask[o] = (sg[4] != 0.f || sg[6 + 4] != 0.f) ? 1.f : 0.f;
}

}  // namespace

int launch_ray_setup(const cpn_render_args& a, int ray0, int nr, float* seg, cudaStream_t st) {
  int total = a.B * nr * 2;
  ray_setup_kernel<<<(total + 127) / 128, 128, 0, st>>>(a, ray0, nr, seg);
  CPN_CHECK_LAUNCH("ray_setup_kernel");
  return CPN_OK;
}

int launch_sample(const cpn_render_args& a, int ray0, int nr, const float* seg, float* rowaux, float* local16,
                  float* A, int a_image, cudaStream_t st) {
  long long rows = (long long)a.B * nr * 2 * a.S;
  sample_kernel<<<(unsigned)((rows + 127) / 128), 128, 0, st>>>(a, ray0, nr, seg, rowaux, local16, A, a_image);
  CPN_CHECK_LAUNCH("sample_kernel");
  return CPN_OK;
}

int launch_ray_epilogue(const cpn_render_args& a, int ray0, int nr, const float* wp, const float* seg, cudaStream_t st) {
  int total = a.B * nr;
  ray_epilogue_kernel<<<(total + 127) / 128, 128, 0, st>>>(a, ray0, nr, wp, seg);
  CPN_CHECK_LAUNCH("ray_epilogue_kernel");
  return CPN_OK;
}

extern "C" int cpn_pair_setup(const float* ctx_c2w, const float* ctx_K, const float* qry_c2w, const float* qry_K,
                              const float* rel_pose, int B, int H, int val, float* consts, void* stream) {
  if (!ctx_c2w || !ctx_K || !qry_c2w || !qry_K || !rel_pose || !consts || B <= 0 || H <= 0) {
    cpn_set_error("cpn_pair_setup: bad argument");
    return CPN_ERR_ARG;
  }
  pair_setup_kernel<<<(B + 31) / 32, 32, 0, (cudaStream_t)stream>>>(ctx_c2w, ctx_K, qry_c2w, qry_K, rel_pose, B, H, val,
                                                                    consts);
  CPN_CHECK_LAUNCH("pair_setup_kernel");
  return CPN_OK;
}

extern "C" int cpn_pair_prologue(const float* flow0, const float* flow1, int B, int fh, int fw, int rgb_w,
                                 float* up_flow2, uint8_t* mask_padded2, void* stream) {
  if (!flow0 || !flow1 || !up_flow2 || !mask_padded2 || B <= 0 || fh <= 0 || fw <= 0 || rgb_w <= 0) {
    cpn_set_error("cpn_pair_prologue: bad argument");
    return CPN_ERR_ARG;
  }
  float scale = (float)(256.0 / (double)rgb_w);
  int total = B * 65536;
  pair_prologue_kernel<<<(total + 255) / 256, 256, 0, (cudaStream_t)stream>>>(flow0, flow1, B, fh, fw, scale, up_flow2,
                                                                              mask_padded2);
  CPN_CHECK_LAUNCH("pair_prologue_kernel");
  return CPN_OK;
}
