// HBM-bound kernels of the generator: InstanceNorm apply (+ReLU, +residual, +halo, +bf16 hi/lo split),
// the fused double feature warp, weight packing and the debug tap reader.
#include <stdlib.h>

#include "common.cuh"
#include "apply_device.cuh"

namespace ap {

void launches_add(int n);

// ------------------------------------------------------------------------------------------------
// apply: y = IN(raw) [+ IN(raw2)] [+ bias] [+ res_in], optional ReLU; -> res_out (fp32) and/or dst.
// grid (pixel chunks, B).  TPP = C/4 threads per pixel: a thread keeps ONE channel quad (its mean / rstd
// live in registers) and walks pixels, four at a time so that 4-12 independent 16-byte loads are in
// flight per thread.  H and W are powers of two (64/128/256): no integer division anywhere.
// ------------------------------------------------------------------------------------------------

// raw conv output is dead after the apply pass: do not let it displace the operands of the next conv in L2
__device__ __forceinline__ float4 ld_stream_evict_first(const float* p) {
  float4 r;
  uint64_t pol;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
  asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.f32 {%0, %1, %2, %3}, [%4], %5;"
               : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
               : "l"(p), "l"(pol));
  return r;
}

__device__ __forceinline__ float4 ld_stream(const float* p) {
  float4 r;
  asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p));
  return r;
}

// MODE 0: IN (or bias) [+ReLU];  1: + InstanceNorm'ed second operand (ResnetBlock2 shortcut);  2: + fp32 residual
// stream;  3: + residual read from the block's input activation (fp32, or bf16 hi + lo = 16 mantissa bits).
// 85 registers at most (3 CTAs = 24 warps per SM, each thread with 4-12 independent 16-byte loads in flight) and
// CTAs of only 8 pixel passes, so that the grid is several balanced waves instead of 1.7 fat ones.
// UNROLL = pixel passes whose loads are in flight together.  4: two rounds per CTA.  8 (MODE 0 only, "deep"): the whole
// CTA's 8 passes are issued before anything else, and the statistics are finalised UNDER those loads -- a MODE 0 CTA
// moves half the bytes of a MODE 1/2 one at the same latency chain (stats -> round 1 -> round 2), which left it at
// 3.6 TB/s where the two-operand modes reach 5-6.6 TB/s (profiles/r01_run11_clip_launches.txt).
template <int TPP, int MODE, int UNROLL>
__global__ void __launch_bounds__(256, 3) apply_kernel(const ApplyP p) {
  constexpr int NY = 256 / TPP;  // pixels per pass
  constexpr int APPLY_PIX_PER_CTA = 8 * NY;
  const int n = blockIdx.y;
  const int cq = threadIdx.x % TPP, py = threadIdx.x / TPP;
  const int c = cq * 4;
  const int HW = p.H * p.W;
  const int logW = 31 - __clz(p.W);
  float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f}, mean2[4] = {0.f, 0.f, 0.f, 0.f}, rstd2[4] = {0.f, 0.f, 0.f, 0.f};
  const int ns = p.src_shared ? 0 : n;  // clip mode: one source image for every destination image (MODE 0 only)
  auto finalise_stats = [&]() {
    const double inv_n = 1.0 / (double)HW;
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (p.stats) stats_to_affine(p.stats, ns, p.stat_C, p.stat_coff, c + e, inv_n, &mean[e], &rstd[e]);
      else { mean[e] = p.bias ? -p.bias[c + e] : 0.f; rstd[e] = 1.f; }
      if (MODE == 1) stats_to_affine(p.stats2, n, p.stat2_C, p.stat2_coff, c + e, inv_n, &mean2[e], &rstd2[e]);
    }
  };
  if (UNROLL != 8) finalise_stats();
  Dst d{p.fmt, p.d0, p.d1, p.dC, p.dcoff, p.dpad, p.H, p.W, p.halo_reflect};
  const int pix0 = blockIdx.x * APPLY_PIX_PER_CTA + py;
  const float* raw = p.raw + ((size_t)ns * HW) * p.raw_C + p.raw_coff + c;
  const float* raw2 = (MODE == 1) ? p.raw2 + ((size_t)n * HW) * p.raw2_C + p.raw2_coff + c : nullptr;
  const float* rin = (MODE == 2) ? p.res_in + ((size_t)n * HW) * p.C + c : nullptr;
  float* rout = p.res_out ? p.res_out + ((size_t)n * HW) * p.C + c : nullptr;
  const int rWp = p.W + 2 * p.res_pad;
  const size_t rbase = (size_t)n * (p.H + 2 * p.res_pad);
#pragma unroll 1
  for (int k0 = 0; k0 < APPLY_PIX_PER_CTA; k0 += UNROLL * NY) {
    float4 v[UNROLL], u[UNROLL];
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int pix = pix0 + k0 + j * NY;
      v[j] = p.l2_hints ? ld_stream_evict_first(raw + (size_t)pix * p.raw_C) : ld_stream(raw + (size_t)pix * p.raw_C);
      if (MODE == 1) u[j] = p.l2_hints ? ld_stream_evict_first(raw2 + (size_t)pix * p.raw2_C) : ld_stream(raw2 + (size_t)pix * p.raw2_C);
      if (MODE == 2) u[j] = ld_stream(rin + (size_t)pix * p.C);
      if (MODE == 3) {
        const size_t roff = ((rbase + (pix >> logW) + p.res_pad) * rWp + (pix & (p.W - 1)) + p.res_pad) * p.res_C + c;
        if (p.res_fmt == FMT_F32) {
          u[j] = *reinterpret_cast<const float4*>(reinterpret_cast<const float*>(p.res_p0) + roff);
        } else {
          const uint2 h = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.res_p0) + roff);
          const uint2 l = *reinterpret_cast<const uint2*>(reinterpret_cast<const __nv_bfloat16*>(p.res_p1) + roff);
          // bf16 -> fp32 is a 16-bit shift
          u[j].x = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
          u[j].y = __uint_as_float(h.x & 0xffff0000u) + __uint_as_float(l.x & 0xffff0000u);
          u[j].z = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
          u[j].w = __uint_as_float(h.y & 0xffff0000u) + __uint_as_float(l.y & 0xffff0000u);
        }
      }
    }
    if (UNROLL == 8) finalise_stats();  // one round per CTA: the statistics' latency hides under the data loads
#pragma unroll
    for (int j = 0; j < UNROLL; ++j) {
      const int pix = pix0 + k0 + j * NY;
      float4 o;
      o.x = (v[j].x - mean[0]) * rstd[0];
      o.y = (v[j].y - mean[1]) * rstd[1];
      o.z = (v[j].z - mean[2]) * rstd[2];
      o.w = (v[j].w - mean[3]) * rstd[3];
      if (p.relu) { o.x = fmaxf(o.x, 0.f); o.y = fmaxf(o.y, 0.f); o.z = fmaxf(o.z, 0.f); o.w = fmaxf(o.w, 0.f); }
      if (MODE == 1) {
        o.x += (u[j].x - mean2[0]) * rstd2[0];
        o.y += (u[j].y - mean2[1]) * rstd2[1];
        o.z += (u[j].z - mean2[2]) * rstd2[2];
        o.w += (u[j].w - mean2[3]) * rstd2[3];
      }
      if (MODE >= 2) { o.x += u[j].x; o.y += u[j].y; o.z += u[j].z; o.w += u[j].w; }
      if (rout) *reinterpret_cast<float4*>(rout + (size_t)pix * p.C) = o;
      if (p.fmt >= 0) store_pixel(d, n, pix >> logW, pix & (p.W - 1), c, o);
    }
  }
}

// An SM has ONE L1 / shared-memory split at a time.  The persistent convs configure it for (almost) all shared memory;
// a kernel that prefers another split may not become co-resident with them; asking for the same carve-out is an option
// for the side-stream kernels.
bool carveout_enabled() {  // AP_NETG_CARVEOUT=1 (A/B): measured 1-2 % SLOWER overall -- the elementwise kernels lose their L1
  static int v = -1;      // and the side-stream kernels overlap the convs about as well without it -- so it is off by default
  if (v < 0) { const char* e = getenv("AP_NETG_CARVEOUT"); v = (e && e[0] == '1'); }
  return v != 0;
}
template <typename K>
static void prefer_max_shared(K kernel) {
  if (carveout_enabled()) cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
}

template <int TPP>
static void apply_carveouts() {
  prefer_max_shared(apply_kernel<TPP, 0, 4>);
  prefer_max_shared(apply_kernel<TPP, 0, 8>);
  prefer_max_shared(apply_kernel<TPP, 1, 4>);
  prefer_max_shared(apply_kernel<TPP, 2, 4>);
  prefer_max_shared(apply_kernel<TPP, 3, 4>);
}

template <int TPP>
static void launch_apply_tpp(const ApplyP& p, dim3 grid, cudaStream_t st) {
  apply_carveouts<TPP>();  // no-op unless AP_NETG_CARVEOUT=1; function attributes are per device, so not cached
  static int deep = -1;  // AP_NETG_APPLY_DEEP=0: the two-round MODE 0 kernel (A/B)
  if (deep < 0) { const char* e = getenv("AP_NETG_APPLY_DEEP"); deep = !(e && e[0] == '0'); }
  if (p.raw2) apply_kernel<TPP, 1, 4><<<grid, 256, 0, st>>>(p);
  else if (p.res_in) apply_kernel<TPP, 2, 4><<<grid, 256, 0, st>>>(p);
  else if (p.res_fmt >= 0) apply_kernel<TPP, 3, 4><<<grid, 256, 0, st>>>(p);
  else if (deep) apply_kernel<TPP, 0, 8><<<grid, 256, 0, st>>>(p);
  else apply_kernel<TPP, 0, 4><<<grid, 256, 0, st>>>(p);
}

static int g_apply_hints = -1;

int launch_apply(const ApplyP& p_in, cudaStream_t st) {
  if (g_apply_hints < 0) {
    const char* e = getenv("AP_NETG_L2_HINTS");
    g_apply_hints = e ? ((atoi(e) >> 1) & 1) : 0;
  }
  ApplyP p = p_in;
  p.l2_hints = g_apply_hints;
  AP_REQUIRE(p.C == 16 || p.C == 128 || p.C == 256, AP_ERR_INVALID, "apply: C=%d (16, 128 or 256)", p.C);
  AP_REQUIRE(p.raw_C % 4 == 0 && p.raw_coff % 4 == 0, AP_ERR_INVALID, "apply: raw channel layout not 16B aligned");
  AP_REQUIRE(p.fmt < 0 || (p.dC % 4 == 0 && p.dcoff % 4 == 0), AP_ERR_INVALID, "apply: dst channel layout");
  AP_REQUIRE((p.raw2 != nullptr) + (p.res_in != nullptr) + (p.res_fmt >= 0) <= 1, AP_ERR_INVALID,
             "apply: shortcut operand and residual stream are exclusive");
  AP_REQUIRE(p.res_fmt < 0 || p.res_fmt == FMT_F32 || p.res_fmt == FMT_BF16X2, AP_ERR_INVALID, "apply: residual format");
  AP_REQUIRE(!p.src_shared || (!p.raw2 && !p.res_in && p.res_fmt < 0 && !p.res_out), AP_ERR_INVALID,
             "apply: a shared source image goes with the plain InstanceNorm mode only");
  const int ppc = 8 * (256 / (p.C / 4));  // pixels per CTA (matches APPLY_PIX_PER_CTA in the kernel)
  AP_REQUIRE((p.W & (p.W - 1)) == 0 && (p.H * p.W) % ppc == 0, AP_ERR_INVALID, "apply: %dx%d", p.H, p.W);
  dim3 grid(p.H * p.W / ppc, p.B);
  if (p.C == 256) launch_apply_tpp<64>(p, grid, st);
  else if (p.C == 128) launch_apply_tpp<32>(p, grid, st);
  else launch_apply_tpp<4>(p, grid, st);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

// ------------------------------------------------------------------------------------------------
// double feature warp (networks.py:1298-1313 + intrinsic_flow_models/modules.py:596-625).
// Arithmetic follows oracle/netg_oracle.py::double_feature_warping_closed_form in
// PyTorch's op order; __f*_rn intrinsics keep nvcc from contracting the coordinate math into FMAs.
// A CTA owns 64 consecutive pixels of one image.  Phase 1: one thread per (pixel, warp kind) computes
// the four bilinear tap offsets and weights ONCE (level > 0 also interpolates the 256x256 conditioning
// maps, align_corners=True) into shared memory.  Phase 2: all threads gather with a (pixel, channel quad)
// mapping -- 16-byte loads, the producer's InstanceNorm + ReLU applied on the fly -- and write both
// halves of the channel concat.
// ------------------------------------------------------------------------------------------------
struct Lerp { int i0, i1; float l0, l1; };

__device__ __forceinline__ Lerp src_index_ac_true(int dst, float scale, int in_size) {
  // upsample_bilinear2d(align_corners=True): src = scale*dst; i0 = (int)src; l1 = src - i0
  const float real = __fmul_rn(scale, (float)dst);
  Lerp r;
  r.i0 = (int)real;
  r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
  r.l1 = fminf(fmaxf(__fsub_rn(real, (float)r.i0), 0.f), 1.f);
  r.l0 = __fsub_rn(1.f, r.l1);
  return r;
}

__device__ __forceinline__ float bilerp(const float* __restrict__ plane, int stride_x, int W, const Lerp& ly,
                                        const Lerp& lx) {
  const float v00 = plane[((size_t)ly.i0 * W + lx.i0) * stride_x];
  const float v01 = plane[((size_t)ly.i0 * W + lx.i1) * stride_x];
  const float v10 = plane[((size_t)ly.i1 * W + lx.i0) * stride_x];
  const float v11 = plane[((size_t)ly.i1 * W + lx.i1) * stride_x];
  const float top = __fadd_rn(__fmul_rn(lx.l0, v00), __fmul_rn(lx.l1, v01));
  const float bot = __fadd_rn(__fmul_rn(lx.l0, v10), __fmul_rn(lx.l1, v11));
  return __fadd_rn(__fmul_rn(ly.l0, top), __fmul_rn(ly.l1, bot));
}

__device__ __forceinline__ float unnormalize_ac_false(float g, int S) {
  // grid_sampler_unnormalize(align_corners=False): ((g + 1) * S - 1) / 2
  return __fdiv_rn(__fsub_rn(__fmul_rn(__fadd_rn(g, 1.f), (float)S), 1.f), 2.f);
}

struct Taps4 { int off[4]; float w[4]; };  // pixel offsets (y*S+x) or -1, weights in order nw, ne, sw, se

__device__ __forceinline__ Taps4 make_taps4(float ix, float iy, int S) {
  const float fx = floorf(ix), fy = floorf(iy);
  const int x0 = (int)fx, y0 = (int)fy;
  // ATen's CPU grid sampler (the oracle): w = x - floor(x), e = 1 - w; nw = s*e, ne = s*w, sw = n*e, se = n*w
  const float wx = __fsub_rn(ix, fx), wy = __fsub_rn(iy, fy);
  const float ex = __fsub_rn(1.f, wx), ey = __fsub_rn(1.f, wy);
  Taps4 t;
  t.w[0] = __fmul_rn(ex, ey); t.w[1] = __fmul_rn(wx, ey); t.w[2] = __fmul_rn(ex, wy); t.w[3] = __fmul_rn(wx, wy);
  const bool x0ok = (x0 >= 0) && (x0 < S), x1ok = (x0 + 1 >= 0) && (x0 + 1 < S);
  const bool y0ok = (y0 >= 0) && (y0 < S), y1ok = (y0 + 1 >= 0) && (y0 + 1 < S);
  t.off[0] = (x0ok && y0ok) ? y0 * S + x0 : -1;
  t.off[1] = (x1ok && y0ok) ? y0 * S + x0 + 1 : -1;
  t.off[2] = (x0ok && y1ok) ? (y0 + 1) * S + x0 : -1;
  t.off[3] = (x1ok && y1ok) ? (y0 + 1) * S + x0 + 1 : -1;
  return t;
}

constexpr int WARP_PIX = 64;  // pixels per CTA

template <int TPP>
__global__ void __launch_bounds__(256) warp_kernel(const WarpP p) {
  __shared__ int s_off[2][4][WARP_PIX];
  __shared__ float s_w[2][4][WARP_PIX];
  __shared__ int s_keep[WARP_PIX];
  const int n = blockIdx.y;
  const int tid = threadIdx.x;
  const int S = p.S;
  constexpr int C = TPP * 4;
  const int pix_base = blockIdx.x * WARP_PIX;

  // ---- phase 1: tap tables ----
  if (tid < 2 * WARP_PIX) {
    const int kind = tid >> 6, lp = tid & (WARP_PIX - 1);  // kind 0: motion warp, 1: flow warp (+ mask)
    const int pix = pix_base + lp;
    const int logS = 31 - __clz(S);
    const int i = pix >> logS, j = pix & (S - 1);
    const float* mo = p.io->motion + (size_t)n * 256 * 256 * 2;
    const float* fl = p.io->flow + (size_t)n * 2 * 256 * 256;
    const float* ms = p.io->ifmask + (size_t)n * 256 * 256;
    Lerp ly{}, lx{};
    if (p.level > 0) {
      const float scale = 255.f / (float)(S - 1);  // fp32((in-1)/(out-1))
      ly = src_index_ac_true(i, scale, 256);
      lx = src_index_ac_true(j, scale, 256);
    }
    Taps4 t;
    if (kind == 0) {
      // motion warp: F.grid_sample(x, motion) with default align_corners=False
      float mx, my;
      if (p.level == 0) { mx = mo[(size_t)pix * 2 + 0]; my = mo[(size_t)pix * 2 + 1]; }
      else { mx = bilerp(mo + 0, 2, 256, ly, lx); my = bilerp(mo + 1, 2, 256, ly, lx); }
      t = make_taps4(unnormalize_ac_false(mx, S), unnormalize_ac_false(my, S), S);
    } else {
      float fx, fy, mk;
      if (p.level == 0) { fx = fl[pix]; fy = fl[256 * 256 + pix]; mk = ms[pix]; }
      else {
        const float fs = (p.level == 1) ? 0.5f : 0.25f;  // flow / 2**level before the resize (exact)
        // scaling by a power of two commutes exactly with the interpolation arithmetic
        fx = bilerp(fl, 1, 256, ly, lx) * fs;
        fy = bilerp(fl + 256 * 256, 1, 256, ly, lx) * fs;
        mk = bilerp(ms, 1, 256, ly, lx);
      }
      // flow warp: grid = 2*(j + f)/(S-1) - 1, then the same un-normalisation
      const float gx = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fadd_rn((float)j, fx)), (float)(S - 1)), 1.f);
      const float gy = __fsub_rn(__fdiv_rn(__fmul_rn(2.f, __fadd_rn((float)i, fy)), (float)(S - 1)), 1.f);
      t = make_taps4(unnormalize_ac_false(gx, S), unnormalize_ac_false(gy, S), S);
      s_keep[lp] = (mk > 0.5f) ? 1 : 0;  // torch.where(mask > 0.5, out, -1)
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) { s_off[kind][k][lp] = t.off[k]; s_w[kind][k][lp] = t.w[k]; }
  }
  // ---- per-thread channel quad: mean / rstd of the producer's InstanceNorm ----
  const int cq = tid % TPP, py = tid / TPP;
  const int c = cq * 4;
  const int ns = p.src_shared ? 0 : n;  // clip mode: every frame warps the same photo features
  float mean[4], rstd[4];
  {
    const double inv_n = 1.0 / (double)(S * S);
#pragma unroll
    for (int e = 0; e < 4; ++e) stats_to_affine(p.stats, ns, p.stat_C, p.stat_coff, c + e, inv_n, &mean[e], &rstd[e]);
  }
  __syncthreads();

  // ---- phase 2: gather ----
  constexpr int NY = 256 / TPP;
  const float* src = p.raw + (size_t)ns * S * S * p.raw_C + p.raw_coff + c;
  const int logS = 31 - __clz(S);
  Dst d{p.fmt, p.d0, p.d1, p.dC, p.dcoff, p.dpad, S, S, 0};
#pragma unroll 1
  for (int lp = py; lp < WARP_PIX; lp += NY) {
    // All 8 gathers are issued back to back, unconditionally: a tap outside the image reads pixel 0 with weight 0
    // (bilinear 'zeros' padding).  With the loads under `if (off >= 0)` every one of them was waited for in turn.
    float4 val[2][4];
    float wt[2][4];
#pragma unroll
    for (int kind = 0; kind < 2; ++kind)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const int off = s_off[kind][k][lp];
        wt[kind][k] = (off >= 0) ? s_w[kind][k][lp] : 0.f;
        val[kind][k] = __ldg(reinterpret_cast<const float4*>(src + (size_t)(off >= 0 ? off : 0) * p.raw_C));
      }
    float4 a = make_float4(0.f, 0.f, 0.f, 0.f), b = a;
#pragma unroll
    for (int kind = 0; kind < 2; ++kind)
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float4 v = val[kind][k];
        v.x = fmaxf((v.x - mean[0]) * rstd[0], 0.f); v.y = fmaxf((v.y - mean[1]) * rstd[1], 0.f);
        v.z = fmaxf((v.z - mean[2]) * rstd[2], 0.f); v.w = fmaxf((v.w - mean[3]) * rstd[3], 0.f);
        float4& acc = kind == 0 ? a : b;
        acc.x = fmaf(v.x, wt[kind][k], acc.x); acc.y = fmaf(v.y, wt[kind][k], acc.y);
        acc.z = fmaf(v.z, wt[kind][k], acc.z); acc.w = fmaf(v.w, wt[kind][k], acc.w);
      }
    if (!s_keep[lp]) b = make_float4(-1.f, -1.f, -1.f, -1.f);
    const int pix = pix_base + lp;
    const int i = pix >> logS, j = pix & (S - 1);
    store_pixel(d, n, i, j, c, a);
    store_pixel(d, n, i, j, C + c, b);
  }
}

int launch_warp(const WarpP& p, cudaStream_t st) {
  AP_REQUIRE(p.C == 32 || p.C == 64 || p.C == 128, AP_ERR_INVALID, "warp: C=%d", p.C);
  AP_REQUIRE((p.S & (p.S - 1)) == 0 && (p.S * p.S) % WARP_PIX == 0, AP_ERR_INVALID, "warp: S=%d", p.S);
  dim3 grid(p.S * p.S / WARP_PIX, p.B);
  prefer_max_shared(warp_kernel<8>);  // no-op unless AP_NETG_CARVEOUT=1
  prefer_max_shared(warp_kernel<16>);
  prefer_max_shared(warp_kernel<32>);
  if (p.C == 32) warp_kernel<8><<<grid, 256, 0, st>>>(p);
  else if (p.C == 64) warp_kernel<16><<<grid, 256, 0, st>>>(p);
  else warp_kernel<32><<<grid, 256, 0, st>>>(p);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

// ------------------------------------------------------------------------------------------------
// pointer table of the caller's tensors (IoPtrs, common.cuh)
// ------------------------------------------------------------------------------------------------
__global__ void set_io_kernel(IoPtrs* dst, const IoPtrs v) { *dst = v; }

int launch_set_io(IoPtrs* dst, const IoPtrs& v, cudaStream_t st) {
  set_io_kernel<<<1, 1, 0, st>>>(dst, v);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

// ------------------------------------------------------------------------------------------------
// weight packing (runs once per load_state_dict)
// ------------------------------------------------------------------------------------------------
__global__ void pack_weights_kernel(const float* __restrict__ src, int Cout, int Cin, int k, int transposed,
                                    float* dst_simt, int simt_C, int simt_coff, __nv_bfloat16* dst_hi,
                                    __nv_bfloat16* dst_lo) {
  const size_t total = (size_t)Cout * Cin * k * k;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    // i enumerates (slab, co, ci)
    const int ci = (int)(i % Cin);
    const int co = (int)((i / Cin) % Cout);
    const int slab = (int)(i / ((size_t)Cin * Cout));
    const size_t s = transposed ? (((size_t)ci * Cout + co) * k * k + slab)   // ConvTranspose2d [Cin,Cout,kh,kw]
                                : (((size_t)co * Cin + ci) * k * k + slab);   // Conv2d [Cout,Cin,kh,kw]
    const float w = src[s];
    if (dst_simt) dst_simt[((size_t)slab * Cin + ci) * simt_C + simt_coff + co] = w;
    if (dst_hi) {
      const __nv_bfloat16 h = __float2bfloat16_rn(w);
      dst_hi[i] = h;  // [slab][Cout][Cin]
      if (dst_lo) dst_lo[i] = __float2bfloat16_rn(w - __bfloat162float(h));
    }
  }
}

int launch_pack_weights(const float* src, int Cout, int Cin, int k, int transposed, float* dst_simt, int simt_C,
                        int simt_coff, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo, cudaStream_t st) {
  const size_t total = (size_t)Cout * Cin * k * k;
  const int blocks = (int)((total + 255) / 256 > 1184 ? 1184 : (total + 255) / 256);
  pack_weights_kernel<<<blocks, 256, 0, st>>>(src, Cout, Cin, k, transposed, dst_simt, simt_C, simt_coff, dst_hi,
                                              dst_lo);
  AP_CUDA(cudaGetLastError());
  return AP_OK;
}

struct PackT {
  PhasePack pk;
  int ntaps, tdy[4], tdx[4];
};

// kernel row used by output phase p (0/1) at input offset d (0/1):  p=0: d=0 -> k=1;  p=1: d=1 -> k=0, d=0 -> k=2
__device__ __forceinline__ int convT_k(int p, int d) { return p == 0 ? (d == 0 ? 1 : -1) : (d == 1 ? 0 : 2); }

__global__ void pack_convT_phases_kernel(const float* __restrict__ src, int Cin, int Cout, PackT q, __nv_bfloat16* dst_hi,
                                         __nv_bfloat16* dst_lo) {
  const int N = q.pk.nph * Cout;
  const size_t total = (size_t)q.ntaps * N * Cin;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int ci = (int)(i % Cin);
    const int n = (int)((i / Cin) % N);
    const int t = (int)(i / ((size_t)Cin * N));
    const int ph = n / Cout, co = n - ph * Cout;
    const int ky = convT_k(q.pk.py[ph], q.tdy[t]), kx = convT_k(q.pk.px[ph], q.tdx[t]);
    const float w = (ky >= 0 && kx >= 0) ? src[((size_t)ci * Cout + co) * 9 + ky * 3 + kx] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    dst_hi[i] = h;
    if (dst_lo) dst_lo[i] = __float2bfloat16_rn(w - __bfloat162float(h));
  }
}

int launch_pack_convT_phases(const float* src, int Cin, int Cout, const PhasePack& pk, int ntaps, const int* tdy,
                             const int* tdx, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo, cudaStream_t st) {
  PackT q{};
  q.pk = pk; q.ntaps = ntaps;
  for (int t = 0; t < ntaps; ++t) { q.tdy[t] = tdy[t]; q.tdx[t] = tdx[t]; }
  pack_convT_phases_kernel<<<592, 256, 0, st>>>(src, Cin, Cout, q, dst_hi, dst_lo);
  AP_CUDA(cudaGetLastError());
  return AP_OK;
}

__global__ void pack_out_weights_kernel(const float* __restrict__ src, int onc, float* dst) {
  // src [onc][64][7][7] -> dst [onc][49][64]
  const int total = onc * 49 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i % 64, t = (i / 64) % 49, o = i / (64 * 49);
    dst[i] = src[((size_t)o * 64 + c) * 49 + t];
  }
}

int launch_pack_out_weights(const float* src, int onc, float* dst, cudaStream_t st) {
  pack_out_weights_kernel<<<(onc * 49 * 64 + 255) / 256, 256, 0, st>>>(src, onc, dst);
  AP_CUDA(cudaGetLastError());
  return AP_OK;
}

// ------------------------------------------------------------------------------------------------
// debug: activation -> NCHW fp32;  NCHW fp32 -> activation (for ap_conv2d_debug)
// ------------------------------------------------------------------------------------------------
__global__ void read_kernel(const ReadP p) {
  const size_t total = (size_t)p.B * p.C * p.H * p.W;
  const double inv_n = 1.0 / (double)(p.H * p.W);
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int x = (int)(i % p.W);
    const int y = (int)((i / p.W) % p.H);
    const int c = (int)((i / ((size_t)p.W * p.H)) % p.C);
    const int n = (int)(i / ((size_t)p.W * p.H * p.C));
    const size_t pix = ((size_t)n * (p.H + 2 * p.spad) + y + p.spad) * (p.W + 2 * p.spad) + x + p.spad;
    const size_t off = pix * p.sC + p.scoff + c;
    float v;
    if (p.fmt == FMT_F32) v = reinterpret_cast<const float*>(p.p0)[off];
    else {
      v = __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.p0)[off]);
      if (p.fmt == FMT_BF16X2) v += __bfloat162float(reinterpret_cast<const __nv_bfloat16*>(p.p1)[off]);
    }
    if (p.stats) {
      float m, r;
      stats_to_affine(p.stats, n, p.stat_C, p.stat_coff, c, inv_n, &m, &r);
      v = (v - m) * r;
      if (p.relu) v = fmaxf(v, 0.f);
    }
    p.dst[i] = v;
  }
}

__global__ void stats_to_double_kernel(const stat_t* __restrict__ src, double* __restrict__ dst, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    dst[i] = (double)src[i] * STAT_INV_SCALE;
}

int launch_stats_to_double(const stat_t* src, double* dst, size_t n, cudaStream_t st) {
  stats_to_double_kernel<<<64, 256, 0, st>>>(src, dst, n);
  AP_CUDA(cudaGetLastError());
  return AP_OK;
}

int launch_read(const ReadP& p, cudaStream_t st) {
  read_kernel<<<1184, 256, 0, st>>>(p);
  AP_CUDA(cudaGetLastError());
  return AP_OK;
}

__global__ void nchw_to_act_kernel(const float* __restrict__ src, Act d) {
  const size_t total = (size_t)d.B * d.C * d.H * d.W;
  const int Hp = d.H + 2 * d.pad, Wp = d.W + 2 * d.pad;
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % d.C);
    const int x = (int)((i / d.C) % d.W);
    const int y = (int)((i / ((size_t)d.C * d.W)) % d.H);
    const int n = (int)(i / ((size_t)d.C * d.W * d.H));
    const float v = src[(((size_t)n * d.C + c) * d.H + y) * d.W + x];
    // interior + reflected halo (pad 1), same rule as store_pixel
    int ys[2] = {y + d.pad, -1}, xs[2] = {x + d.pad, -1};
    if (d.pad == 1) {
      ys[1] = (y == 1) ? 0 : ((y == d.H - 2) ? d.H + 1 : -1);
      xs[1] = (x == 1) ? 0 : ((x == d.W - 2) ? d.W + 1 : -1);
    }
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b) {
        if (ys[a] < 0 || xs[b] < 0) continue;
        const size_t off = (((size_t)n * Hp + ys[a]) * Wp + xs[b]) * d.C + c;
        if (d.fmt == FMT_F32) reinterpret_cast<float*>(d.p0)[off] = v;
        else {
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          reinterpret_cast<__nv_bfloat16*>(d.p0)[off] = h;
          if (d.fmt == FMT_BF16X2)
            reinterpret_cast<__nv_bfloat16*>(d.p1)[off] = __float2bfloat16_rn(v - __bfloat162float(h));
        }
      }
  }
}

int launch_nchw_to_act(const float* src, const Act& dst, cudaStream_t st) {
  nchw_to_act_kernel<<<1184, 256, 0, st>>>(src, dst);
  AP_CUDA(cudaGetLastError());
  return AP_OK;
}

}  // namespace ap
