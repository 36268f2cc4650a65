// Per-frame conditioning producers of the generator ("next" row f2 of the scope table, + the matting line of f1):
//   draw2 op 0        Module2/data/umlvdfw_test_dataset.py:34-41     68 filled discs -> landmark map in {-1,+1}
//   cal_motion256     Module2/data/umlvdfw_test_dataset.py:67-81     scipy griddata(linear): Delaunay + barycentric
//   kp_to_map_some    Module2/models/geomcgt_ifw_test_model.py:12-44 68 binary key-point maps (netF input)
//   photo matting     Module2/models/geomcgt_ifw_test_model.py:280,292
// The reference runs these on ONE CPU core per frame (cal_motion256 alone is ~5 ms, i.e. a 200 frames/s ceiling in front
// of a generator that renders thousands).  Here a whole batch of frames is one launch each and only the 68x2 landmark
// coordinates travel.  All of it is byte/HBM-write bound except the triangulation, which is a few MFLOP of fp64 per frame.
//
// Arithmetic follows the reference's op order with explicit round-to-nearest intrinsics (no FMA contraction), so the
// motion field is bit-identical to oracle/cond_oracle.py, which in turn is bit-identical to scipy on the golden cases.
#include "common.cuh"

#include "../../include/ap_cond.h"

namespace ap {

void launches_add(int n);

namespace {

constexpr int NLM = AP_COND_LANDMARKS;
constexpr int NS = NLM + 4;  // sites of the triangulation: landmarks + the 4 distinct corners of cal_motion256's `edges`
constexpr int MAXT = AP_COND_MAX_TRIANGLES;
constexpr double INCIRCLE_TOL = 1e-9;
constexpr double INSIDE_TOL = 1e-9;

struct TriRec { int i, j, k, key; };

__device__ __forceinline__ double dmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double dadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double dsub(double a, double b) { return __dsub_rn(a, b); }

// sites 0..67 = landmarks (float32 widened, as np.concatenate with the int64 `edges` array does), 68..71 = corners
__device__ __forceinline__ void load_sites(const float* lm, double* sx, double* sy) {
  const int t = threadIdx.x;
  if (t < NLM) {
    const float2 v = reinterpret_cast<const float2*>(lm)[t];
    sx[t] = (double)v.x;
    sy[t] = (double)v.y;
  } else if (t < NS) {
    const int c = t - NLM;  // (0,0) (255,255) (0,255) (255,0)
    sx[t] = (c == 1 || c == 3) ? 255.0 : 0.0;
    sy[t] = (c == 1 || c == 2) ? 255.0 : 0.0;
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Delaunay by exhaustion: one thread per site triple i < j < k (C(72,3) = 59,640 per frame, unranked from the flat thread
// index so that every CTA carries the same work); the triple is kept when no other site lies strictly inside its
// circumcircle.  In general position the survivors are THE Delaunay triangulation (2n-2-h triangles); co-circular sites
// keep every triangulation of their cell, the rasteriser then picks by smallest key.
// ------------------------------------------------------------------------------------------------------------------
constexpr int NTRIPLES = NS * (NS - 1) * (NS - 2) / 6;

__device__ __forceinline__ int triples_before(int i) {  // triples whose first index is < i
  const int r = NS - i;
  return NTRIPLES - r * (r - 1) * (r - 2) / 6;
}

// fp32 screen of the in-circle determinant: +1 = certainly inside, -1 = certainly outside, 0 = too close to call (the
// caller repeats the test in fp64).  M bounds every product that enters the determinant, so the fp32 rounding error of
// `det` is below ~1e-6 * M; a margin of 1e-5 * M therefore never contradicts the fp64 decision det64 > 1e-9 * mag64
// (mag64 <= M).  The fp64 pipe was the bound of this kernel (73 % active, profiles/r01_run11_cond_ncu.txt): with the
// screen it only sees nearly co-circular quadruples.
__device__ __forceinline__ int incircle_screen(float ax, float ay, float bx, float by, float cx, float cy, float px,
                                               float py, float sgn) {
  const float adx = ax - px, ady = ay - py, bdx = bx - px, bdy = by - py, cdx = cx - px, cdy = cy - py;
  const float a2 = adx * adx + ady * ady, b2 = bdx * bdx + bdy * bdy, c2 = cdx * cdx + cdy * cdy;
  const float u1 = bdy * c2, u2 = b2 * cdy, u3 = bdx * c2, u4 = b2 * cdx, u5 = bdx * cdy, u6 = bdy * cdx;
  const float det = (adx * (u1 - u2) - ady * (u3 - u4) + a2 * (u5 - u6)) * sgn;
  const float M = fabsf(adx) * (fabsf(u1) + fabsf(u2)) + fabsf(ady) * (fabsf(u3) + fabsf(u4)) + a2 * (fabsf(u5) + fabsf(u6));
  const float thr = 1e-5f * M;
  return det > thr ? 1 : (det < -thr ? -1 : 0);
}

// exact in-circle decision of the oracle (float64, its op order): site p strictly inside the circumcircle of (a, b, c)
__device__ __forceinline__ bool incircle_exact(double ax, double ay, double bx, double by, double cx, double cy, double px,
                                               double py, double sgn) {
  const double adx = dsub(ax, px), ady = dsub(ay, py);
  const double bdx = dsub(bx, px), bdy = dsub(by, py);
  const double cdx = dsub(cx, px), cdy = dsub(cy, py);
  const double a2 = dadd(dmul(adx, adx), dmul(ady, ady));
  const double b2 = dadd(dmul(bdx, bdx), dmul(bdy, bdy));
  const double c2 = dadd(dmul(cdx, cdx), dmul(cdy, cdy));
  const double t1 = dmul(adx, dsub(dmul(bdy, c2), dmul(b2, cdy)));
  const double t2 = dmul(ady, dsub(dmul(bdx, c2), dmul(b2, cdx)));
  const double t3 = dmul(a2, dsub(dmul(bdx, cdy), dmul(bdy, cdx)));
  const double det = dmul(dadd(dsub(t1, t2), t3), sgn);
  const double mag = dadd(dadd(fabs(t1), fabs(t2)), fabs(t3));
  return det > dmul(INCIRCLE_TOL, mag);
}

// Two phases per CTA of 128 triples.  A: every thread tests its triple against the <= 12 index-neighbours (+-1, +-2) of
// its three vertices -- landmark indices run along facial contours, so these are the sites most likely to sit inside the
// circumcircle: 99 % of the triples are rejected here after 2 tests on average, and no lane waits long for another.
// B: the few survivors are queued in shared memory and each is checked against ALL sites by a whole warp, one site per
// lane (a serial scan per thread left 12 of 32 lanes active on average: profiles/r01_run11_cond_ncu.txt).
__global__ void __launch_bounds__(128) delaunay_kernel(const float* __restrict__ lm_dst, int* __restrict__ counts,
                                                       TriRec* __restrict__ tris) {
  __shared__ double sx[NS], sy[NS];
  __shared__ float fx[NS], fy[NS];
  __shared__ int before[NS];
  __shared__ TriRec queue[128];  // key field: orientation sign of the triple
  __shared__ int n_queue;
  const int f = blockIdx.y;
  load_sites(lm_dst + (size_t)f * NLM * 2, sx, sy);
  if (threadIdx.x < NS) {
    before[threadIdx.x] = triples_before(threadIdx.x);
    fx[threadIdx.x] = (float)sx[threadIdx.x];  // exact: the sites are float32 values (or the corners 0 / 255)
    fy[threadIdx.x] = (float)sy[threadIdx.x];
  }
  if (threadIdx.x == 0) n_queue = 0;
  __syncthreads();
  const int t = blockIdx.x * 128 + threadIdx.x;
  if (t < NTRIPLES) {
    int lo = 0, hi = NS - 3;  // first index: the largest i with before[i] <= t
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (before[mid] <= t) lo = mid; else hi = mid - 1;
    }
    const int i = lo, m = NS - 1 - i, r = t - before[i];  // r ranks the pair (jj < kk) among the m sites after i
    const float b2m = (float)(2 * m - 1);
    int jj = (int)((b2m - sqrtf(fmaxf(b2m * b2m - 8.f * (float)r, 0.f))) * 0.5f);
    jj = max(0, min(jj, m - 2));
    while ((jj + 1) * m - (jj + 1) * (jj + 2) / 2 <= r) ++jj;
    while (jj * m - jj * (jj + 1) / 2 > r) --jj;
    const int kk = r - (jj * m - jj * (jj + 1) / 2) + jj + 1;
    const int j = i + 1 + jj, k = i + 1 + kk;
    const double ax = sx[i], ay = sy[i], bx = sx[j], by = sy[j], cx = sx[k], cy = sy[k];
    const double e1x = dsub(bx, ax), e1y = dsub(by, ay), e2x = dsub(cx, ax), e2y = dsub(cy, ay);
    const double orient = dsub(dmul(e1x, e2y), dmul(e1y, e2x));
    const double span = fmax(fmax(fabs(e1x), fabs(e1y)), fmax(fabs(e2x), fabs(e2y)));
    bool alive = fabs(orient) > dmul(1e-12, fmax(dmul(span, span), 1e-300));  // not collinear / repeated sites
    const double sgn = orient > 0.0 ? 1.0 : -1.0;
    const float fax = fx[i], fay = fy[i], fbx = fx[j], fby = fy[j], fcx = fx[k], fcy = fy[k], fsgn = (float)sgn;
#pragma unroll 1
    for (int q = 0; q < 12 && alive; ++q) {
      const int off = 1 + q / 6, r6 = q - (q / 6) * 6;
      const int base = r6 < 2 ? i : (r6 < 4 ? j : k);
      int d = base + ((r6 & 1) ? -off : off);
      d = d < 0 ? d + NS : (d >= NS ? d - NS : d);
      if (d == i || d == j || d == k) continue;
      const int screen = incircle_screen(fax, fay, fbx, fby, fcx, fcy, fx[d], fy[d], fsgn);
      if (screen > 0 || (screen == 0 && incircle_exact(ax, ay, bx, by, cx, cy, sx[d], sy[d], sgn))) alive = false;
    }
    if (alive) queue[atomicAdd(&n_queue, 1)] = TriRec{i, j, k, sgn > 0.0 ? 1 : -1};
  }
  __syncthreads();
  const int lane = threadIdx.x & 31, nq = n_queue;
  for (int s = threadIdx.x >> 5; s < nq; s += 4) {
    const TriRec tr = queue[s];
    const double sgn = (double)tr.key;
    bool inside = false;
    for (int d = lane; d < NS && !inside; d += 32) {
      if (d == tr.i || d == tr.j || d == tr.k) continue;
      const int screen = incircle_screen(fx[tr.i], fy[tr.i], fx[tr.j], fy[tr.j], fx[tr.k], fy[tr.k], fx[d], fy[d], (float)tr.key);
      inside = screen > 0 || (screen == 0 && incircle_exact(sx[tr.i], sy[tr.i], sx[tr.j], sy[tr.j], sx[tr.k], sy[tr.k],
                                                             sx[d], sy[d], sgn));
    }
    if (!__any_sync(0xffffffffu, inside) && lane == 0) {
      const int slot = atomicAdd(&counts[f], 1);
      if (slot < MAXT) tris[(size_t)f * MAXT + slot] = TriRec{tr.i, tr.j, tr.k, (tr.i * NS + tr.j) * NS + tr.k};
    }
  }
}

// ------------------------------------------------------------------------------------------------------------------
// Rasteriser: one CTA per 32x8-pixel tile of one frame.  The tile first culls the frame's triangle list against its own
// rectangle (a barycentric coordinate is affine, so "negative at all four tile corners" rejects exactly), keeping the
// precomputed 2x2 edge matrix of the survivors in shared memory; every pixel then walks the short list, takes the
// containing triangle with the smallest key (deterministic under co-circularity, equal to the oracle's "first in
// lexicographic order") and interpolates the source position with LinearNDInterpolator's barycentric formula in fp64.
// ------------------------------------------------------------------------------------------------------------------
struct TileTri {
  double r2x, r2y, m00, m01, m10, m11, det;
  double lim;  // -1e-6 * det^2: "barycentric coordinate >= -1e-6" without the division (numerator * det >= lim)
  int i, j, k, key;
};

// Division-free screen: the three barycentric numerators times det.  c_e >= -1e-6 <=> n_e * det >= -1e-6 * det^2.  A
// strict superset of the exact test below (tolerance 1e-9, rounding ~1e-13), at a seventh of its cost: the two fp64
// divisions of the exact formula are only paid for the one or two triangles that really contain the pixel.
__device__ __forceinline__ void bary_numerators(const TileTri& t, double qx, double qy, double& n0, double& n1, double& n2) {
  const double dx = qx - t.r2x, dy = qy - t.r2y;
  const double a0 = t.m11 * dx - t.m01 * dy, a1 = t.m00 * dy - t.m10 * dx;
  n0 = a0 * t.det; n1 = a1 * t.det; n2 = (t.det - a0 - a1) * t.det;
}

__device__ __forceinline__ void bary(const TileTri& t, double qx, double qy, double& c0, double& c1, double& c2) {
  const double dx = dsub(qx, t.r2x), dy = dsub(qy, t.r2y);
  c0 = __ddiv_rn(dsub(dmul(t.m11, dx), dmul(t.m01, dy)), t.det);
  c1 = __ddiv_rn(dadd(dmul(-t.m10, dx), dmul(t.m00, dy)), t.det);
  c2 = dsub(dsub(1.0, c0), c1);
}

__global__ void __launch_bounds__(256) motion_raster_kernel(const float* __restrict__ lm_src, int src_stride,
                                                            const float* __restrict__ lm_dst,
                                                            const int* __restrict__ counts,
                                                            const TriRec* __restrict__ tris, float* __restrict__ motion) {
  __shared__ double dsx[NS], dsy[NS], vsx[NS], vsy[NS];
  __shared__ TileTri list[MAXT];
  __shared__ int n_list;
  const int f = blockIdx.z;
  load_sites(lm_dst + (size_t)f * NLM * 2, dsx, dsy);
  load_sites(lm_src + (size_t)f * src_stride, vsx, vsy);
  if (threadIdx.x == 0) n_list = 0;
  __syncthreads();
  const int x0 = blockIdx.x * 32, y0 = blockIdx.y * 8;
  const int ntri = min(counts[f], MAXT);
  for (int t = threadIdx.x; t < ntri; t += 256) {
    const TriRec r = tris[(size_t)f * MAXT + t];
    TileTri tt;
    tt.r2x = dsx[r.k]; tt.r2y = dsy[r.k];
    tt.m00 = dsub(dsx[r.i], tt.r2x); tt.m01 = dsub(dsx[r.j], tt.r2x);
    tt.m10 = dsub(dsy[r.i], tt.r2y); tt.m11 = dsub(dsy[r.j], tt.r2y);
    tt.det = dsub(dmul(tt.m00, tt.m11), dmul(tt.m01, tt.m10));
    tt.lim = -1e-6 * tt.det * tt.det;
    tt.i = r.i; tt.j = r.j; tt.k = r.k; tt.key = r.key;
    double mx0 = -1e300, mx1 = -1e300, mx2 = -1e300;
#pragma unroll
    for (int c = 0; c < 4; ++c) {
      double n0, n1, n2;
      bary_numerators(tt, (double)(x0 + (c & 1) * 31), (double)(y0 + (c >> 1) * 7), n0, n1, n2);
      mx0 = fmax(mx0, n0); mx1 = fmax(mx1, n1); mx2 = fmax(mx2, n2);
    }
    // 1e-6 of slack on the cull: it only has to be conservative, the per-pixel test below decides
    if (mx0 >= tt.lim && mx1 >= tt.lim && mx2 >= tt.lim) list[atomicAdd(&n_list, 1)] = tt;
  }
  __syncthreads();
  const int n = n_list;
  const int x = x0 + (threadIdx.x & 31), y = y0 + (threadIdx.x >> 5);
  const double qx = (double)x, qy = (double)y;
  int best = -1, best_key = 0x7fffffff;
  double b0 = 0.0, b1 = 0.0, b2 = 0.0;
  for (int t = 0; t < n; ++t) {
    double c0, c1, c2;
    bary_numerators(list[t], qx, qy, c0, c1, c2);
    const double lim = list[t].lim;
    if (!(c0 >= lim && c1 >= lim && c2 >= lim) || list[t].key >= best_key) continue;
    bary(list[t], qx, qy, c0, c1, c2);
    if (c0 >= -INSIDE_TOL && c1 >= -INSIDE_TOL && c2 >= -INSIDE_TOL && list[t].key < best_key) {
      best = t; best_key = list[t].key; b0 = c0; b1 = c1; b2 = c2;
    }
  }
  float2 o;
  if (best >= 0) {
    const int i = list[best].i, j = list[best].j, k = list[best].k;
    const double vx = dadd(dadd(dmul(b0, vsx[i]), dmul(b1, vsx[j])), dmul(b2, vsx[k]));
    const double vy = dadd(dadd(dmul(b0, vsy[i]), dmul(b1, vsy[j])), dmul(b2, vsy[k]));
    // map_xy.astype(float32) / 127.5 - 1, in float32
    o.x = __fsub_rn(__fdiv_rn(__double2float_rn(vx), 127.5f), 1.0f);
    o.y = __fsub_rn(__fdiv_rn(__double2float_rn(vy), 127.5f), 1.0f);
  } else {
    o.x = o.y = __int_as_float(0x7fc00000);  // outside the hull: griddata's fill value
  }
  reinterpret_cast<float2*>(motion)[((size_t)f * 256 + y) * 256 + x] = o;
}

// ------------------------------------------------------------------------------------------------------------------
// draw2 op 0: a thread owns 4 consecutive pixels of a row and tests them against all discs (centres in shared memory).
// ------------------------------------------------------------------------------------------------------------------
struct DrawP {
  const float* lands;
  float* out;
  int n_points, size, radius;
  unsigned long long hw;  // 16 nibbles: half-width of the filled span at row offset |dy| (OpenCV midpoint circle)
};

__global__ void __launch_bounds__(256) draw_kernel(const DrawP p) {
  __shared__ int cx[256], cy[256];
  const int f = blockIdx.y;
  for (int l = threadIdx.x; l < p.n_points; l += 256) {
    const float2 v = reinterpret_cast<const float2*>(p.lands)[(size_t)f * p.n_points + l];
    const bool ok = fabsf(v.x) < 1e6f && fabsf(v.y) < 1e6f;  // NaN / far-away points never touch the canvas
    cx[l] = ok ? (int)rintf(v.x) : -(1 << 29);                // np.round: half to even
    cy[l] = ok ? (int)rintf(v.y) : -(1 << 29);
  }
  __syncthreads();
  const int quad = blockIdx.x * 256 + threadIdx.x;
  if (quad * 4 >= p.size * p.size) return;
  const int y = (quad * 4) / p.size, x = (quad * 4) - y * p.size;
  unsigned hit = 0;
  for (int l = 0; l < p.n_points; ++l) {
    const int ady = abs(y - cy[l]);
    if (ady <= p.radius) {
      const int h = (int)((p.hw >> (4 * ady)) & 15ull), d0 = x - cx[l];
#pragma unroll
      for (int e = 0; e < 4; ++e) hit |= (abs(d0 + e) <= h ? 1u : 0u) << e;
    }
  }
  float4 o;  // uint8 255 / 255. * 2 - 1 = +1, 0 -> -1
  o.x = (hit & 1u) ? 1.f : -1.f; o.y = (hit & 2u) ? 1.f : -1.f; o.z = (hit & 4u) ? 1.f : -1.f; o.w = (hit & 8u) ? 1.f : -1.f;
  reinterpret_cast<float4*>(p.out)[(size_t)f * (p.size * p.size / 4) + quad] = o;
}

// ------------------------------------------------------------------------------------------------------------------
// kp_to_map (binary): plane (frame, point), a thread owns 4 consecutive pixels; the distance test runs in fp64 as numpy's
// int64 grid minus a float32 scalar does.
// ------------------------------------------------------------------------------------------------------------------
// Plane (t, k) = blockIdx.y lands at out + t * image_stride + k * size * size (image_stride = K * size * size: the
// contiguous [T,K,size,size] tensor; larger: one half of a channel concatenation).  kps_image_stride = 0: the same K
// points for every image.  pre78: the points are first scaled by 7/8 in fp32, (x * 7) / 8, as the caller of the
// reference does (geomcgt_ifw_test_model.py:65-66: lm.numpy() * 7 / 8 on a float32 array).
__device__ __forceinline__ float2 kp_point(const float* kps, int t, int k, int kps_image_stride, int pre78) {
  float2 kp = reinterpret_cast<const float2*>(kps)[(size_t)t * kps_image_stride + k];
  if (pre78) {
    kp.x = __fdiv_rn(__fmul_rn(kp.x, 7.f), 8.f);
    kp.y = __fdiv_rn(__fmul_rn(kp.y, 7.f), 8.f);
  }
  return kp;
}

__global__ void __launch_bounds__(256) kp_kernel(const float* __restrict__ kps, float* __restrict__ out, int size,
                                                 double r2, int K, size_t image_stride, int kps_image_stride, int pre78) {
  const int plane = blockIdx.y;
  const int t = plane / K, k = plane - t * K;
  const float2 kp = kp_point(kps, t, k, kps_image_stride, pre78);
  const bool empty = (kp.x == -1.f) || (kp.y == -1.f);
  const int quad = blockIdx.x * 256 + threadIdx.x;
  if (quad * 4 >= size * size) return;
  const int y = (quad * 4) / size, x = (quad * 4) - y * size;
  const double dy = dsub((double)y, (double)kp.y), dy2 = dmul(dy, dy);
  float v[4];
#pragma unroll
  for (int e = 0; e < 4; ++e) {
    const double dx = dsub((double)(x + e), (double)kp.x);
    v[e] = (!empty && dadd(dmul(dx, dx), dy2) <= r2) ? 1.f : 0.f;
  }
  reinterpret_cast<float4*>(out + (size_t)t * image_stride + (size_t)k * size * size)[quad] = make_float4(v[0], v[1], v[2], v[3]);
}

// A box that CONTAINS the disc of every key point (floor / ceil of centre -+ radius, clipped to the map; empty for a missing
// point or a disc off the map), written as (y0, y1, x0, x1) at bbox[t * box_image_stride + box_off + k]; the box areas of
// an image are summed into area[t] (integer adds: order-independent).  For the zero-skipping first conv of the flow network.
__global__ void kp_box_kernel(const float* __restrict__ kps, int T, int K, int size, float radius, int kps_image_stride,
                              int pre78, int4* __restrict__ bbox, int box_image_stride, int box_off, int* __restrict__ area) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= T * K) return;
  const int t = i / K, k = i - t * K;
  const float2 kp = kp_point(kps, t, k, kps_image_stride, pre78);
  int4 b = make_int4(size, -1, size, -1);
  if (!(kp.x == -1.f || kp.y == -1.f) && kp.x == kp.x && kp.y == kp.y) {
    const float fy0 = floorf(kp.y - radius) - 1.f, fy1 = ceilf(kp.y + radius) + 1.f;
    const float fx0 = floorf(kp.x - radius) - 1.f, fx1 = ceilf(kp.x + radius) + 1.f;
    const int y0 = (int)fmaxf(fy0, 0.f), y1 = (int)fminf(fy1, (float)(size - 1));
    const int x0 = (int)fmaxf(fx0, 0.f), x1 = (int)fminf(fx1, (float)(size - 1));
    if (y1 >= y0 && x1 >= x0) {
      b = make_int4(y0, y1, x0, x1);
      atomicAdd(&area[t], (y1 - y0 + 1) * (x1 - x0 + 1));
    }
  }
  bbox[(size_t)t * box_image_stride + box_off + k] = b;
}

// ------------------------------------------------------------------------------------------------------------------
// photo matting, in the reference's op order:  ((x/2 + 0.5) * m + 1 - m) * 2 - 1,  m = (matte > 0.5)
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) matte_kernel(const float* __restrict__ real_A, const float* __restrict__ matte,
                                                    float* __restrict__ out, float* __restrict__ mask, int C, int HW) {
  const int b = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;
  if (pix >= HW) return;
  const float m = matte[(size_t)b * HW + pix] > 0.5f ? 1.f : 0.f;
  if (mask) mask[(size_t)b * HW + pix] = m;
  if (out) {
    for (int c = 0; c < C; ++c) {
      const size_t off = ((size_t)b * C + c) * HW + pix;
      const float a = __fadd_rn(__fdiv_rn(real_A[off], 2.f), 0.5f);
      const float s = __fsub_rn(__fadd_rn(__fmul_rn(a, m), 1.f), m);
      out[off] = __fsub_rn(__fmul_rn(s, 2.f), 1.f);
    }
  }
}

// OpenCV's filled circle (imgproc/drawing.cpp, integer midpoint walk over the first octant): half-width per row offset
void circle_halfwidths(int radius, int* hw) {
  for (int i = 0; i < 16; ++i) hw[i] = 0;
  int dx = radius, dy = 0, err = 0, inc = 1, dec = 2 * radius - 1;
  while (dx >= dy) {
    if (dx > hw[dy]) hw[dy] = dx;
    if (dy > hw[dx]) hw[dx] = dy;
    ++dy;
    err += inc;
    inc += 2;
    if (err > 0) { err -= dec; --dx; dec -= 2; }
  }
}

size_t motion_counts_bytes(int T) { return (((size_t)T * sizeof(int)) + 255) / 256 * 256; }

}  // namespace
}  // namespace ap

using namespace ap;

extern "C" int ap_cond_draw_landmarks(int device, int T, int n_points, int size, int radius, const float* lands, float* out,
                                      void* cuda_stream) {
  AP_REQUIRE(T >= 1 && T <= 65535 && lands && out, AP_ERR_INVALID, "draw_landmarks: bad argument");
  AP_REQUIRE(n_points >= 1 && n_points <= 256, AP_ERR_INVALID, "draw_landmarks: n_points must be in [1,256]");
  AP_REQUIRE(size >= 4 && size % 4 == 0 && size <= 4096, AP_ERR_INVALID, "draw_landmarks: size must be a multiple of 4");
  AP_REQUIRE(radius >= 0 && radius <= 15, AP_ERR_INVALID, "draw_landmarks: radius must be in [0,15]");
  AP_CUDA(cudaSetDevice(device));
  DrawP p;
  p.lands = lands; p.out = out; p.n_points = n_points; p.size = size; p.radius = radius;
  int hw[16];
  circle_halfwidths(radius, hw);
  p.hw = 0;
  for (int i = 0; i < 16; ++i) p.hw |= (unsigned long long)hw[i] << (4 * i);
  dim3 grid((size * size / 4 + 255) / 256, T);
  draw_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(p);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

extern "C" int ap_cond_motion256_workspace_bytes(int T, size_t* bytes) {
  AP_REQUIRE(T >= 1 && bytes, AP_ERR_INVALID, "motion256_workspace_bytes: bad argument");
  *bytes = motion_counts_bytes(T) + (size_t)T * MAXT * sizeof(TriRec);
  return AP_OK;
}

extern "C" int ap_cond_motion256(int device, int T, const float* lm_src, int src_per_frame, const float* lm_dst,
                                 float* motion, void* workspace, size_t workspace_bytes, int32_t* tri_count,
                                 void* cuda_stream) {
  AP_REQUIRE(T >= 1 && lm_src && lm_dst && motion && workspace, AP_ERR_INVALID, "motion256: bad argument");
  AP_REQUIRE(T <= 65535, AP_ERR_INVALID, "motion256: at most 65535 frames per call");
  size_t need = 0;
  ap_cond_motion256_workspace_bytes(T, &need);
  AP_REQUIRE(workspace_bytes >= need, AP_ERR_INVALID, "motion256: workspace of %zu bytes, %zu needed", workspace_bytes, need);
  AP_REQUIRE(((uintptr_t)workspace & 15) == 0, AP_ERR_INVALID, "motion256: workspace must be 16-byte aligned");
  AP_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  int* counts = (int*)workspace;
  TriRec* tris = (TriRec*)((char*)workspace + motion_counts_bytes(T));
  AP_CUDA(cudaMemsetAsync(counts, 0, (size_t)T * sizeof(int), st));
  delaunay_kernel<<<dim3((NTRIPLES + 127) / 128, T), 128, 0, st>>>(lm_dst, counts, tris);
  AP_CUDA(cudaGetLastError());
  motion_raster_kernel<<<dim3(256 / 32, 256 / 8, T), 256, 0, st>>>(lm_src, src_per_frame ? NLM * 2 : 0, lm_dst, counts, tris,
                                                                  motion);
  AP_CUDA(cudaGetLastError());
  if (tri_count) AP_CUDA(cudaMemcpyAsync(tri_count, counts, (size_t)T * sizeof(int), cudaMemcpyDeviceToDevice, st));
  launches_add(2);
  return AP_OK;
}

extern "C" int ap_cond_kp_to_map(int device, int T, int K, int size, float radius, const float* kps, float* out,
                                 void* cuda_stream) {
  AP_REQUIRE(T >= 1 && K >= 1 && kps && out, AP_ERR_INVALID, "kp_to_map: bad argument");
  AP_REQUIRE((long long)T * K <= 65535, AP_ERR_INVALID, "kp_to_map: at most 65535 maps per call");
  AP_REQUIRE(size >= 4 && size % 4 == 0 && size <= 4096, AP_ERR_INVALID, "kp_to_map: size must be a multiple of 4");
  AP_REQUIRE(radius >= 0.f, AP_ERR_INVALID, "kp_to_map: negative radius");
  AP_CUDA(cudaSetDevice(device));
  dim3 grid((size * size / 4 + 255) / 256, T * K);
  kp_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(kps, out, size, (double)radius * (double)radius, K,
                                                         (size_t)K * size * size, K, 0);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

namespace ap {
// Key-point maps of one half of the flow network's input, written straight into the concatenated [T, C, size, size]
// operand at channel `coff`, plus the boxes of their discs (flownet.cu: ap_flow_warp_landmarks).  2 launches.
int launch_kp_half(const float* kps, int per_frame, int T, int K, int size, float radius, float* out, int C, int coff,
                   int4* bbox, int* area, cudaStream_t st) {
  AP_REQUIRE((long long)T * K <= 65535 && size % 4 == 0, AP_ERR_INVALID, "kp maps: at most 65535 maps per call, size %% 4 == 0");
  dim3 grid((size * size / 4 + 255) / 256, T * K);
  kp_kernel<<<grid, 256, 0, st>>>(kps, out + (size_t)coff * size * size, size, (double)radius * (double)radius, K,
                                  (size_t)C * size * size, per_frame ? K : 0, 1);
  AP_CUDA(cudaGetLastError());
  kp_box_kernel<<<(T * K + 255) / 256, 256, 0, st>>>(kps, T, K, size, radius, per_frame ? K : 0, 1, bbox, C, coff, area);
  AP_CUDA(cudaGetLastError());
  launches_add(2);
  return AP_OK;
}
}  // namespace ap

extern "C" int ap_cond_matte_photo(int device, int B, int C, int HW, const float* real_A, const float* matte, float* out,
                                   float* mask, void* cuda_stream) {
  AP_REQUIRE(B >= 1 && B <= 65535 && C >= 1 && HW >= 1 && matte, AP_ERR_INVALID, "matte_photo: bad argument");
  AP_REQUIRE(out || mask, AP_ERR_INVALID, "matte_photo: no output requested");
  AP_REQUIRE(!out || real_A, AP_ERR_INVALID, "matte_photo: real_A missing");
  AP_CUDA(cudaSetDevice(device));
  dim3 grid((HW + 255) / 256, B);
  matte_kernel<<<grid, 256, 0, (cudaStream_t)cuda_stream>>>(real_A, matte, out, mask, C, HW);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}
