// The intrinsic-flow network netF (`FlowUnet`, Module2/intrinsic_flow_models/networks.py:510-644) and the post-processing of
// `flow_network_warp` (Module2/models/geomcgt_ifw_test_model.py:62-76) on the GPU: include/ap_flow.h, SURVEY.md §8 row f3.
//
// First version on CUDA cores (fp32 FMA): the U-Net's 4x4 stride-2 convolutions live on 112/56/28/14/7-pixel grids that do
// not tile into the 128-pixel M tiles of the tcgen05 kernels, and the configuration the released checkpoint uses is not known
// (train_opt.json does not ship), so the network is built from ONE parametric implicit-GEMM kernel:
//   * `fconv_kernel`: 64 pixels x 64 output channels per CTA, pixels indexed flat over (image, y, x) so that any grid size
//     works; tap list (7x7, 3x3 s2, 4x4 s2, the four 2x2 phases of ConvTranspose2d(k4,s2,p1) as blockIdx.z); NHWC or NCHW
//     input.  The normalisation and activation that precede a conv in the reference are applied WHILE ITS OPERAND IS LOADED:
//     a = act(x * scale[n][c] + shift[n][c]) with act in {LeakyReLU 0.1, LeakyReLU 0.2, ReLU}; zero padding applies to the
//     activated tensor as in the reference.  Raw conv outputs are what is stored; a skip concatenation is a channel offset.
//   * BatchNorm (eval): scale/shift come from the checkpoint (gamma / sqrt(var + eps), beta - mean * scale; a conv bias
//     folds into shift).  InstanceNorm: `fstats_kernel` reduces each raw output plane in a fixed order (deterministic).
//   * The reference's in-place activations are part of the semantics (networks.py:523,552-568): the skip half of a
//     concatenation is LeakyReLU_0.2(x), and the parent's in-place ReLU then acts on it -- here simply the activation the
//     consuming conv applies to those channels.
//   * `fpost_kernel` / `fresize_kernel`: x2 bilinear up-sampling (align_corners=False), arg-max, mask, rescale and the
//     align_corners=True resize to 256x256, in PyTorch's op order.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "../../include/ap_flow.h"
#include "common.cuh"
#include "umma.cuh"

namespace ap {

void launches_add(int n);
int64_t launches_get();
int launch_kp_half(const float* kps, int per_frame, int T, int K, int size, float radius, float* out, int C, int coff,
                   int4* bbox, int* area, cudaStream_t st);

enum { FACT_NONE = 0, FACT_LRELU01 = 1, FACT_LRELU02 = 2, FACT_RELU = 3 };

struct FTap {
  int8_t dy, dx;
  uint8_t slab;
};

struct FConvP {
  const float* in;     // NHWC [B,Hin,Win,in_C] (channels in_coff .. in_coff+Cin) or NCHW [B,Cin,Hin,Win]
  int in_nchw, in_C, in_coff;
  const float* scale;  // [B][in_C] per-image per-channel affine of the operand, or null (identity)
  const float* shift;
  int act;
  const float* w;      // [slab][Cin][CoutP]
  const float* bias;   // [Cout] or null
  float* out;          // NHWC [B,Hout,Wout,out_C] at channel out_coff
  int out_C, out_coff;
  int B, Hin, Win, Cin, Hv, Wv, stride, Cout, CoutP;
  int os, Hout, Wout;
  int nphase, ntaps;   // blockIdx.z = phase: output pixel (y*os + py, x*os + px)
  const int* gate_area;  // sparse-operand gate (first conv): image n is left to fconv_sparse_kernel when gate_area[n] <= gate_thresh
  int gate_thresh;
  int8_t py[4], px[4];
  FTap taps[4][49];
};

__device__ __forceinline__ float fact(float v, int act) {
  if (act == FACT_LRELU01) return v > 0.f ? v : 0.1f * v;
  if (act == FACT_LRELU02) return v > 0.f ? v : 0.2f * v;
  if (act == FACT_RELU) return fmaxf(v, 0.f);
  return v;
}

__global__ void __launch_bounds__(256) fconv_kernel(const __grid_constant__ FConvP p) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][64];
  const int tid = threadIdx.x;
  const int HW = p.Hv * p.Wv;
  const int M = p.B * HW;
  const int m0 = blockIdx.x * 64, n0 = blockIdx.y * 64, ph = blockIdx.z;
  const int tx = tid & 15, ty = tid >> 4;
  const int Ktot = p.ntaps * p.Cin;
  const FTap* taps = p.taps[ph];
  if (p.gate_area) {  // every image this tile touches went to the sparse kernel: nothing to do
    const int ia = m0 / HW, ib = min(m0 + 63, M - 1) / HW;
    bool any = false;
    for (int i = ia; i <= ib; ++i) any |= p.gate_area[i] > p.gate_thresh;
    if (!any) return;
  }

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  // operand loader: thread -> (pixel lm, k offsets kk0, kk0 + 4, ...)
  const int lm = tid & 63;
  const int m = m0 + lm;
  const bool m_ok = m < M;
  const int img = m_ok ? m / HW : 0;
  const int pix = m_ok ? m - img * HW : 0;
  const int vy = pix / p.Wv, vx = pix - vy * p.Wv;

  for (int k0 = 0; k0 < Ktot; k0 += 16) {
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const int kk = (tid >> 6) + 4 * i;
      const int k = k0 + kk;
      float v = 0.f;
      if (m_ok && k < Ktot) {
        const int t = k / p.Cin;
        const int c = k - t * p.Cin;
        const int iy = vy * p.stride + taps[t].dy, ix = vx * p.stride + taps[t].dx;
        if (iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win) {
          if (p.in_nchw) v = p.in[((size_t)(img * p.Cin + c) * p.Hin + iy) * p.Win + ix];
          else v = p.in[((size_t)(img * p.Hin + iy) * p.Win + ix) * p.in_C + p.in_coff + c];
          if (p.scale) v = fmaf(v, p.scale[(size_t)img * p.in_C + p.in_coff + c], p.shift[(size_t)img * p.in_C + p.in_coff + c]);
          v = fact(v, p.act);
        }
      }
      As[kk][lm] = v;
    }
    {
      const int kk = tid >> 4;
      const int k = k0 + kk;
      const int col = (tid & 15) * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < Ktot && n0 + col < p.CoutP) {
        const int t = k / p.Cin;
        const int c = k - t * p.Cin;
        w = *reinterpret_cast<const float4*>(p.w + ((size_t)taps[t].slab * p.Cin + c) * p.CoutP + n0 + col);
      }
      *reinterpret_cast<float4*>(&Bs[kk][col]) = w;
    }
    __syncthreads();
#pragma unroll
    for (int k = 0; k < 16; ++k) {
      const float4 a = *reinterpret_cast<const float4*>(&As[k][ty * 4]);
      const float4 b = *reinterpret_cast<const float4*>(&Bs[k][tx * 4]);
      const float av[4] = {a.x, a.y, a.z, a.w};
      const float bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int mm = m0 + ty * 4 + i;
    if (mm >= M) continue;
    const int im = mm / HW, pm = mm - im * HW;
    if (p.gate_area && p.gate_area[im] <= p.gate_thresh) continue;
    const int y = pm / p.Wv, x = pm - y * p.Wv;
    const int oy = y * p.os + p.py[ph], ox = x * p.os + p.px[ph];
    float* dst = p.out + ((size_t)(im * p.Hout + oy) * p.Wout + ox) * p.out_C + p.out_coff;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const int n = n0 + tx * 4 + j;
      if (n < p.Cout) dst[n] = acc[i][j] + (p.bias ? p.bias[n] : 0.f);
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// Register-tiled version of the same implicit GEMM for NHWC operands with Cin % 16 == 0 (every layer but the first):
// 128 pixels x BN output channels per CTA, 8 x (BN/16) accumulators per thread, K walked in 16-channel steps that never
// straddle a tap (so the tap geometry is per step, not per element), operands fetched as 16-byte vectors -- four threads
// read the 64 contiguous bytes of one pixel -- and double-buffered through registers into shared memory (one barrier per
// step).  Accumulation order per output element is k ascending, as in fconv_kernel: the results are independent of the
// tile a pixel falls into, hence of the batch size.
// ---------------------------------------------------------------------------------------------------------------
constexpr int FT_BM = 128, FT_BK = 16, FT_LDA = FT_BM + 4;

template <int BN>
__global__ void __launch_bounds__(256, 2) fconv_tiled_kernel(const __grid_constant__ FConvP p) {
  constexpr int TN = BN / 16;            // columns per thread: 8 (two groups of 4), 4 or 1
  constexpr int NBV = (FT_BK * BN / 4 + 255) / 256;  // 16-byte weight vectors per thread and step
  __shared__ __align__(16) float As[2][FT_BK][FT_LDA];
  __shared__ __align__(16) float Bs[2][FT_BK][BN];
  const int tid = threadIdx.x;
  const int HW = p.Hv * p.Wv;
  const int M = p.B * HW;
  const int m0 = blockIdx.x * FT_BM, n0 = blockIdx.y * BN, ph = blockIdx.z;
  const int tx = tid & 15, ty = tid >> 4;
  const FTap* taps = p.taps[ph];
  const int nsteps = p.ntaps * (p.Cin / FT_BK);
  const int steps_per_tap = p.Cin / FT_BK;

  // operand loader: pixels lm0 = tid >> 2 and lm0 + 64, channel quad q = tid & 3 of the step's 16 channels
  const int q = tid & 3;
  int l_img[2], l_y[2], l_x[2];
  bool l_ok[2];
#pragma unroll
  for (int r = 0; r < 2; ++r) {
    const int m = m0 + (tid >> 2) + 64 * r;
    l_ok[r] = m < M;
    const int mm = l_ok[r] ? m : 0;
    l_img[r] = mm / HW;
    const int pix = mm - l_img[r] * HW;
    l_y[r] = pix / p.Wv;
    l_x[r] = pix - l_y[r] * p.Wv;
  }

  // activation as one select: LeakyReLU slope (1 = identity, 0 = ReLU)
  const float slope = p.act == FACT_LRELU01 ? 0.1f : p.act == FACT_LRELU02 ? 0.2f : p.act == FACT_RELU ? 0.f : 1.f;
  float4 ra[2], rb[NBV];
  bool r_ok[2];
  // fetch(step): issue the global loads of a step into registers and nothing else -- the operand transform waits until
  // stash(), after the FMAs of the current step, so that the load latency sits under them
  auto fetch = [&](int step) {
    const int t = step / steps_per_tap;
    const int cb = (step - t * steps_per_tap) * FT_BK;
    const int dy = taps[t].dy, dx = taps[t].dx, slab = taps[t].slab;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      const int iy = l_y[r] * p.stride + dy, ix = l_x[r] * p.stride + dx;
      r_ok[r] = l_ok[r] && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
      ra[r] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (r_ok[r])
        ra[r] = *reinterpret_cast<const float4*>(p.in + ((size_t)(l_img[r] * p.Hin + iy) * p.Win + ix) * p.in_C + p.in_coff + cb + 4 * q);
    }
#pragma unroll
    for (int i = 0; i < NBV; ++i) {
      const int e = tid + 256 * i;               // vector index: row kk = e / (BN/4), column (e % (BN/4)) * 4
      const int kk = e / (BN / 4), col = (e - kk * (BN / 4)) * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (kk < FT_BK && n0 + col < p.CoutP)
        w = *reinterpret_cast<const float4*>(p.w + ((size_t)slab * p.Cin + cb + kk) * p.CoutP + n0 + col);
      rb[i] = w;
    }
  };
  auto stash = [&](int buf, int step) {
    const int c = (step % steps_per_tap) * FT_BK + 4 * q;
#pragma unroll
    for (int r = 0; r < 2; ++r) {
      float4 v = ra[r];
      if (r_ok[r]) {  // zero padding applies to the activated tensor: out-of-range taps stay 0
        if (p.scale) {
          const float4 sc = *reinterpret_cast<const float4*>(p.scale + (size_t)l_img[r] * p.in_C + p.in_coff + c);
          const float4 sh = *reinterpret_cast<const float4*>(p.shift + (size_t)l_img[r] * p.in_C + p.in_coff + c);
          v.x = fmaf(v.x, sc.x, sh.x); v.y = fmaf(v.y, sc.y, sh.y); v.z = fmaf(v.z, sc.z, sh.z); v.w = fmaf(v.w, sc.w, sh.w);
        }
        v.x = v.x > 0.f ? v.x : v.x * slope; v.y = v.y > 0.f ? v.y : v.y * slope;
        v.z = v.z > 0.f ? v.z : v.z * slope; v.w = v.w > 0.f ? v.w : v.w * slope;
      }
      const int lm = (tid >> 2) + 64 * r;
      As[buf][4 * q + 0][lm] = v.x;
      As[buf][4 * q + 1][lm] = v.y;
      As[buf][4 * q + 2][lm] = v.z;
      As[buf][4 * q + 3][lm] = v.w;
    }
#pragma unroll
    for (int i = 0; i < NBV; ++i) {
      const int e = tid + 256 * i;
      const int kk = e / (BN / 4), col = (e - kk * (BN / 4)) * 4;
      if (kk < FT_BK) *reinterpret_cast<float4*>(&Bs[buf][kk][col]) = rb[i];
    }
  };

  float acc[8][TN];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

  fetch(0);
  stash(0, 0);
  __syncthreads();
  for (int step = 0; step < nsteps; ++step) {
    const int buf = step & 1;
    if (step + 1 < nsteps) fetch(step + 1);
#pragma unroll
    for (int k = 0; k < FT_BK; ++k) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float bv[TN];
      if constexpr (TN == 1) {
        bv[0] = Bs[buf][k][tx];
      } else {
        const float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
        bv[0] = b0.x; bv[1] = b0.y; bv[2] = b0.z; bv[3] = b0.w;
        if constexpr (TN == 8) {
          const float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
          bv[4] = b1.x; bv[5] = b1.y; bv[6] = b1.z; bv[7] = b1.w;
        }
      }
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (step + 1 < nsteps) stash(buf ^ 1, step + 1);
    __syncthreads();
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int mm = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (mm >= M) continue;
    const int im = mm / HW, pm = mm - im * HW;
    const int y = pm / p.Wv, x = pm - y * p.Wv;
    const int oy = y * p.os + p.py[ph], ox = x * p.os + p.px[ph];
    float* dst = p.out + ((size_t)(im * p.Hout + oy) * p.Wout + ox) * p.out_C + p.out_coff;
    if constexpr (TN == 1) {
      const int n = n0 + tx;
      if (n < p.Cout) dst[n] = acc[i][0] + (p.bias ? p.bias[n] : 0.f);
    } else {
#pragma unroll
      for (int g = 0; g < TN / 4; ++g) {
        const int n = n0 + 64 * g + tx * 4;
        if (n + 3 < p.Cout) {
          float4 o = make_float4(acc[i][4 * g], acc[i][4 * g + 1], acc[i][4 * g + 2], acc[i][4 * g + 3]);
          if (p.bias) { o.x += p.bias[n]; o.y += p.bias[n + 1]; o.z += p.bias[n + 2]; o.w += p.bias[n + 3]; }
          *reinterpret_cast<float4*>(dst + n) = o;
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j)
            if (n + j < p.Cout) dst[n + j] = acc[i][4 * g + j] + (p.bias ? p.bias[n + j] : 0.f);
        }
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------------------------
// The first conv of netF reads 2 x 68 key-point maps: one disc of <= 49 pixels per 224 x 224 plane, i.e. 99.9 % zeros,
// and is 70 % of the dense network's FLOPs.  `fbbox_kernel` finds the bounding box of the non-zeros of every (image,
// channel) plane; `fconv_sparse_kernel` gives each output pixel only the (channel, tap) pairs whose input can be non-zero.
// A product with an exact zero contributes exactly nothing, so this is the same sum as the dense conv with the zero terms
// left out (k ascending, taps in raster order).  Nothing is assumed about the caller's tensor: an image whose boxes cover
// more than 1/8 of the volume stays with the dense kernel (per-image decision on the device, no host synchronisation; per
// image so that a frame's result does not depend on its batch neighbours).
// ---------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) fbbox_kernel(const float* __restrict__ x, int H, int W, int C, int4* __restrict__ bbox,
                                                    int* __restrict__ area) {
  const int plane = blockIdx.x;  // n * C + c
  const float* src = x + (size_t)plane * H * W;
  int y0 = H, y1 = -1, x0 = W, x1 = -1;
  for (int i = threadIdx.x; i < H * W; i += 256) {
    if (src[i] != 0.f) {
      const int yy = i / W, xx = i - yy * W;
      y0 = min(y0, yy); y1 = max(y1, yy); x0 = min(x0, xx); x1 = max(x1, xx);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    y0 = min(y0, __shfl_xor_sync(0xffffffffu, y0, o)); y1 = max(y1, __shfl_xor_sync(0xffffffffu, y1, o));
    x0 = min(x0, __shfl_xor_sync(0xffffffffu, x0, o)); x1 = max(x1, __shfl_xor_sync(0xffffffffu, x1, o));
  }
  __shared__ int4 sb[8];
  if ((threadIdx.x & 31) == 0) sb[threadIdx.x >> 5] = make_int4(y0, y1, x0, x1);
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int k = 1; k < 8; ++k) {
      y0 = min(y0, sb[k].x); y1 = max(y1, sb[k].y); x0 = min(x0, sb[k].z); x1 = max(x1, sb[k].w);
    }
    bbox[plane] = make_int4(y0, y1, x0, x1);  // empty plane: y1 < y0
    if (y1 >= y0) atomicAdd(&area[plane / C], (y1 - y0 + 1) * (x1 - x0 + 1));  // integer sum: order-independent
  }
}

struct FSparseP {
  const float* in;     // NCHW [B,Cin,H,W]
  const int4* bbox;    // [B*Cin]
  const int* area;     // [B]
  int thresh;
  const float* w;      // [k*k][Cin][CoutP]
  const float* bias;
  float* out;          // NHWC [B,H,W,out_C] at channel out_coff
  int out_C, out_coff;
  int B, H, W, Cin, Cout, CoutP, k, pad;
};
constexpr int FS_MAXC = 512;

// grid (tiles of 16 x 16 output pixels, groups of CG output channels, images); a thread owns one pixel and CG channels
template <int CG>
__global__ void __launch_bounds__(256) fconv_sparse_kernel(const __grid_constant__ FSparseP p) {
  const int n = blockIdx.z;
  if (p.area[n] > p.thresh) return;  // dense image: fconv_kernel's
  __shared__ int4 sb[FS_MAXC];
  for (int c = threadIdx.x; c < p.Cin; c += 256) sb[c] = p.bbox[(size_t)n * p.Cin + c];
  __syncthreads();
  const int tiles_x = (p.W + 15) / 16;
  const int ty0 = (blockIdx.x / tiles_x) * 16, tx0 = (blockIdx.x % tiles_x) * 16;
  const int y = ty0 + (threadIdx.x >> 4), x = tx0 + (threadIdx.x & 15);
  const int cg = blockIdx.y * CG;
  if (y >= p.H || x >= p.W) return;
  float acc[CG];
#pragma unroll
  for (int j = 0; j < CG; ++j) acc[j] = 0.f;
  // input rows of output row y: y - pad .. y - pad + k - 1
  const int wy0 = y - p.pad, wy1 = y - p.pad + p.k - 1, wx0 = x - p.pad, wx1 = x - p.pad + p.k - 1;
  const int ry0 = ty0 - p.pad, ry1 = ty0 + 15 - p.pad + p.k - 1, rx0 = tx0 - p.pad, rx1 = tx0 + 15 - p.pad + p.k - 1;
  for (int c = 0; c < p.Cin; ++c) {
    const int4 bb = sb[c];
    if (bb.y < ry0 || bb.x > ry1 || bb.w < rx0 || bb.z > rx1) continue;  // uniform over the CTA (covers empty planes)
    const int ia = max(wy0, bb.x), ib = min(wy1, bb.y), ja = max(wx0, bb.z), jb = min(wx1, bb.w);
    const float* src = p.in + ((size_t)n * p.Cin + c) * p.H * p.W;
    for (int iy = ia; iy <= ib; ++iy)
      for (int ix = ja; ix <= jb; ++ix) {
        const float v = src[iy * p.W + ix];
        if (v != 0.f) {
          const int slab = (iy - wy0) * p.k + (ix - wx0);
          const float* wp = p.w + ((size_t)slab * p.Cin + c) * p.CoutP + cg;
#pragma unroll
          for (int g = 0; g < CG / 4; ++g)
            if (cg + 4 * g < p.CoutP) {
              const float4 w4 = __ldg(reinterpret_cast<const float4*>(wp + 4 * g));
              acc[4 * g + 0] = fmaf(v, w4.x, acc[4 * g + 0]);
              acc[4 * g + 1] = fmaf(v, w4.y, acc[4 * g + 1]);
              acc[4 * g + 2] = fmaf(v, w4.z, acc[4 * g + 2]);
              acc[4 * g + 3] = fmaf(v, w4.w, acc[4 * g + 3]);
            }
        }
      }
  }
  float* dst = p.out + ((size_t)(n * p.H + y) * p.W + x) * p.out_C + p.out_coff;
  if ((p.out_C & 3) == 0 && (p.out_coff & 3) == 0) {
#pragma unroll
    for (int g = 0; g < CG / 4; ++g) {
      const int c = cg + 4 * g;
      if (c + 3 < p.Cout) {
        float4 o = make_float4(acc[4 * g], acc[4 * g + 1], acc[4 * g + 2], acc[4 * g + 3]);
        if (p.bias) { o.x += p.bias[c]; o.y += p.bias[c + 1]; o.z += p.bias[c + 2]; o.w += p.bias[c + 3]; }
        *reinterpret_cast<float4*>(dst + c) = o;
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j)
          if (c + j < p.Cout) dst[c + j] = acc[4 * g + j] + (p.bias ? p.bias[c + j] : 0.f);
      }
    }
  } else {
#pragma unroll
    for (int j = 0; j < CG; ++j)
      if (cg + j < p.Cout) dst[cg + j] = acc[j] + (p.bias ? p.bias[cg + j] : 0.f);
  }
}

// InstanceNorm2d(affine=False) of a raw NHWC plane as an operand transform: scale = rstd, shift = -mean * rstd.
// One CTA per (image, 32-channel group): 8 pixel lanes x 32 channels, then the 8 partial sums in a fixed order.
__global__ void __launch_bounds__(256) fstats_kernel(const float* __restrict__ x, int HW, int C, int coff, int nC, float* scale,
                                                     float* shift) {
  __shared__ double ssum[8][32], ssq[8][32];
  const int n = blockIdx.y, c = blockIdx.x * 32 + (threadIdx.x & 31), r = threadIdx.x >> 5;
  double s = 0.0, q = 0.0;
  if (c < nC) {
    const float* src = x + (size_t)n * HW * C + coff + c;
    for (int p = r; p < HW; p += 8) {
      const double v = (double)src[(size_t)p * C];
      s += v;
      q += v * v;
    }
  }
  ssum[r][threadIdx.x & 31] = s;
  ssq[r][threadIdx.x & 31] = q;
  __syncthreads();
  if (r == 0 && c < nC) {
    for (int k = 1; k < 8; ++k) { s += ssum[k][threadIdx.x]; q += ssq[k][threadIdx.x]; }
    const double mean = s / HW;
    double var = q / HW - mean * mean;
    if (var < 0.0) var = 0.0;
    const float rstd = 1.0f / sqrtf((float)var + 1e-5f);
    scale[(size_t)n * C + coff + c] = rstd;
    shift[(size_t)n * C + coff + c] = -(float)mean * rstd;
  }
}

__global__ void fbroadcast_kernel(const float* __restrict__ sc, const float* __restrict__ sh, int C, int coff, int nC, int B,
                                  float* scale, float* shift) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * nC) return;
  const int n = i / nC, c = i - n * nC;
  scale[(size_t)n * C + coff + c] = sc[c];
  shift[(size_t)n * C + coff + c] = sh[c];
}

// ---------------------------------------------------------------------------------------------------------------
// Tensor-core version of the same implicit GEMM (tcgen05.mma kind::f16, fp32 accumulators in TMEM) for the layers with
// Cin % 32 == 0 and Cout in {<= 16, 64, multiples of 128} -- every conv of the U-Net proper at nf >= 32, and the heads.
// fp32 accuracy comes from the scheme the generator's convs use: operands split as x = hi + lo (bf16 each), three products
// A_hi*W_hi + A_hi*W_lo + A_lo*W_hi.
//
// One CTA per (image, 128-pixel tile, BN-channel tile, phase).  The 4x4 stride-2 convs and the four 2x2 phases of the
// transposed conv live on 112/56/28/14/7-pixel grids, so TMA boxes do not fit; the A operand is BUILT: sixteen builder
// warps (a warp instruction reads two whole pixel rows of a chunk, 2 x 256 contiguous bytes) read the raw fp32 NHWC
// activations of one (tap, 64-channel chunk), apply the producer's normalisation and activation, split into hi/lo and
// write the 128-byte-swizzled K-major tile tcgen05.mma reads (the same layout the stem kernel of the generator builds).
// Weights are packed once, at load time, as pre-swizzled [slab][chunk][plane][Cout][64] bf16 images, so a stage's W tile
// is two plain bulk copies.
// Warp roles (704 threads): warp 0 = TMEM allocation + MMA issuer, warp 1 = W loader, warps 2..17 = A builders,
// warps 18..21 = epilogue (tcgen05.ld -> 128 contiguous bytes per pixel row -> global, bias added).
// A ring of NS stages; one `full` barrier per stage collects the 512 builder arrivals and the W bytes, one `empty`
// barrier (tcgen05.commit) releases the stage to both producers.
// ---------------------------------------------------------------------------------------------------------------
struct FUmmaP {
  FConvP c;              // geometry, operand transform, taps, output placement (c.w unused)
  const uint8_t* wimg;   // [slab][Cin/kc][2 planes][wrows][64 k] bf16, rows 128-byte swizzled in groups of 8
  int tiles_per_img;     // ceil(Hv * Wv / 128)
  int kc;                // channels per K chunk: 64, or 32 when Cin is only a multiple of 32 (half-filled rows, 2 k-steps)
  int wrows;             // rows of a weight plane: Cout, padded to 16 for the 5-channel heads
};

constexpr int FU_BWARPS = 16;                           // builder warps
constexpr int FU_THREADS = 64 + 32 * FU_BWARPS + 128;   // MMA warp, W loader, builders, 4 epilogue warps
constexpr int FU_RPT = 128 / FU_BWARPS / 2;             // rows per builder thread and chunk
constexpr int FU_APLANE = 128 * 128;     // [128 rows x 64 k] bf16
constexpr int FU_ASTAGE = 2 * FU_APLANE; // hi + lo
constexpr int FU_MAXC = 1024;            // operand channels whose scale / shift are staged in shared memory

template <int BN>
struct FUCfg {
  static constexpr int NS = BN >= 256 ? 2 : (BN == 128 ? 3 : 4);
  static constexpr int TCOLS = BN < 32 ? 32 : BN;   // TMEM columns (power of two >= 32)
  static constexpr int WPLANE = BN * 128;
  static constexpr int STAGE = FU_ASTAGE + 2 * WPLANE;
  static constexpr size_t SMEM = 1024 + (size_t)NS * STAGE + 2 * FU_MAXC * 4 + 128;
};

template <int BN>
__global__ void __launch_bounds__(FU_THREADS, 1) fconv_umma_kernel(const __grid_constant__ FUmmaP pp) {
  using Cfg = FUCfg<BN>;
  constexpr int NS = Cfg::NS;
  const FConvP& p = pp.c;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: NS stages {A hi | A lo | W hi | W lo} | scale[FU_MAXC] | shift[FU_MAXC] | barriers
  float* s_scale = reinterpret_cast<float*>(sgen + (size_t)NS * Cfg::STAGE);
  float* s_shift = s_scale + FU_MAXC;
  const uint32_t bars = smem_base + NS * Cfg::STAGE + 2 * FU_MAXC * 4;
  // full[s] +8s, empty[s] +32+8s, tfull +64, tmem pointer +72
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sgen + (size_t)NS * Cfg::STAGE + 2 * FU_MAXC * 4 + 72);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int HW = p.Hv * p.Wv;
  const int img = blockIdx.x / pp.tiles_per_img;
  const int pix0 = (blockIdx.x - img * pp.tiles_per_img) * 128;
  const int n0 = blockIdx.y * BN;
  const int ph = blockIdx.z;
  const int kc = pp.kc;
  const int ncb = p.Cin / kc;
  const int nchunks = p.ntaps * ncb;

  if (warp == 0) {
    if (lane == 0) {
      for (int s = 0; s < NS; ++s) {
        mbar_init(bars + 8 * s, 32 * FU_BWARPS + 1);  // full: every builder thread + the W loader's expect_tx arrival
        mbar_init(bars + 32 + 8 * s, 1);    // empty: one tcgen05.commit
      }
      mbar_init(bars + 64, 1);              // accumulator complete
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)Cfg::TCOLS)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (p.scale) {  // the tile lies in ONE image: its per-channel affine is staged once
    for (int c = threadIdx.x; c < p.Cin; c += FU_THREADS) {
      s_scale[c] = p.scale[(size_t)img * p.in_C + p.in_coff + c];
      s_shift[c] = p.shift[(size_t)img * p.in_C + p.in_coff + c];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t accum = 0;
      const int ksteps = kc >> 4;
      for (int s = 0; s < nchunks; ++s) {
        const int st = s % NS;
        mbar_wait(bars + 8 * st, (uint32_t)(s / NS) & 1u);
        tc_fence_after();
        const uint32_t sa = smem_base + st * Cfg::STAGE;
        const uint64_t a_hi = make_sw128_desc(sa), a_lo = make_sw128_desc(sa + FU_APLANE);
        const uint64_t w_hi = make_sw128_desc(sa + FU_ASTAGE), w_lo = make_sw128_desc(sa + FU_ASTAGE + Cfg::WPLANE);
        for (int k = 0; k < ksteps; ++k) {
          const uint64_t o = (uint64_t)(k * 2);
          umma_bf16(tmem_base, a_hi + o, w_hi + o, idesc, accum);
          accum = 1;
          umma_bf16(tmem_base, a_hi + o, w_lo + o, idesc, 1);
          umma_bf16(tmem_base, a_lo + o, w_hi + o, idesc, 1);
        }
        umma_commit(bars + 32 + 8 * st);
      }
      umma_commit(bars + 64);
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== W loader =====================
    if (lane == 0) {
      const size_t plane_bytes = (size_t)pp.wrows * 128;
      for (int s = 0; s < nchunks; ++s) {
        const int st = s % NS;
        mbar_wait(bars + 32 + 8 * st, ((uint32_t)(s / NS) & 1u) ^ 1u);
        const int t = s / ncb, cb = s - t * ncb;
        const int slab = p.taps[ph][t].slab;
        const uint8_t* src = pp.wimg + ((size_t)slab * ncb + cb) * 2 * plane_bytes + (size_t)(n0 >> 3) * 1024;
        const uint32_t dst = smem_base + st * Cfg::STAGE + FU_ASTAGE;
        mbar_expect_tx(bars + 8 * st, 2u * Cfg::WPLANE);
        bulk_load(dst, src, Cfg::WPLANE, bars + 8 * st);
        bulk_load(dst + Cfg::WPLANE, src + plane_bytes, Cfg::WPLANE, bars + 8 * st);
      }
    }
    __syncwarp();
  } else if (warp < 2 + FU_BWARPS) {
    // ===================== A builders =====================
    // A warp instruction reads two whole pixel rows of the chunk (2 x 256 contiguous bytes): lane -> (row parity, 16-byte
    // column), builder warp w owns rows 8w .. 8w+7, FU_RPT rows per thread.  Sixteen warps: the conversion (normalise,
    // activation, hi/lo split: ~8 instructions per element) is what bounds this kernel, not the MMAs.
    const int bw = warp - 2;                 // 0..FU_BWARPS-1
    const int col = lane & 15;               // channels 4*col .. 4*col+3 of the chunk
    const int rsub = lane >> 4;
    const float slope = p.act == FACT_LRELU01 ? 0.1f : p.act == FACT_LRELU02 ? 0.2f : p.act == FACT_RELU ? 0.f : 1.f;
    const bool has_affine = p.scale != nullptr;
    const float* in_img = p.in + (size_t)img * p.Hin * p.Win * p.in_C + p.in_coff + 4 * col;
    const bool col_ok = 4 * col < kc;        // kc = 32: only the first half of a row carries data (the MMAs read 2 k-steps)
    int ry[FU_RPT], rx[FU_RPT];              // input coordinates of tap (0,0) for this thread's rows; ry < -64: no such pixel
    int goff[FU_RPT];                        // element offset of that pixel in the image; soff: byte offset of the row's 8 bytes
    int soff[FU_RPT];                        // in the K-major SWIZZLE_128B tile (16-byte group col/2, half col&1)
#pragma unroll
    for (int i = 0; i < FU_RPT; ++i) {
      const int row = bw * (2 * FU_RPT) + 2 * i + rsub;
      const int pix = pix0 + row;
      const int vy = pix / p.Wv;
      ry[i] = (pix < HW && col_ok) ? vy * p.stride : -100000;
      rx[i] = (pix - vy * p.Wv) * p.stride;
      goff[i] = ry[i] >= 0 ? (ry[i] * p.Win + rx[i]) * p.in_C : 0;
      soff[i] = (row >> 3) * 1024 + (row & 7) * 128 + (((col >> 1) ^ (row & 7)) << 4) + (col & 1) * 8;
    }

    float4 ra[FU_RPT], rb[FU_RPT];
    uint32_t oka = 0, okb = 0;
    // chunk counters advanced incrementally (no divisions in the loop): the fetch side runs one chunk ahead of the build side
    int f_t = 0, f_cb = 0;
    int b_st = 0, b_cb = 0;
    uint32_t b_par = 1u;                     // parity to wait for on empty[stage]: first use of a stage passes
    auto fetch = [&](float4 (&r)[FU_RPT], uint32_t& ok) {
      const int dy = p.taps[ph][f_t].dy, dx = p.taps[ph][f_t].dx;
      const int delta = (dy * p.Win + dx) * p.in_C + f_cb * kc;
      ok = 0;
#pragma unroll
      for (int i = 0; i < FU_RPT; ++i) {
        const int iy = ry[i] + dy, ix = rx[i] + dx;
        if ((unsigned)iy < (unsigned)p.Hin && (unsigned)ix < (unsigned)p.Win) {
          ok |= 1u << i;
          r[i] = *reinterpret_cast<const float4*>(in_img + (goff[i] + delta));
        }
      }
      if (++f_cb == ncb) { f_cb = 0; ++f_t; }
    };
    auto build = [&](const float4 (&r)[FU_RPT], uint32_t ok) {
      float4 sc = make_float4(1.f, 1.f, 1.f, 1.f), sh = make_float4(0.f, 0.f, 0.f, 0.f);
      if (has_affine && col_ok) {
        sc = *reinterpret_cast<const float4*>(s_scale + b_cb * kc + 4 * col);
        sh = *reinterpret_cast<const float4*>(s_shift + b_cb * kc + 4 * col);
      }
      mbar_wait(bars + 32 + 8 * b_st, b_par);
      uint8_t* slot = sgen + (size_t)b_st * Cfg::STAGE;
#pragma unroll
      for (int i = 0; i < FU_RPT; ++i) {
        uint2 hi = make_uint2(0u, 0u), lo = make_uint2(0u, 0u);
        if (ok & (1u << i)) {   // zero padding applies to the activated tensor: out-of-range taps stay 0
          float v[4] = {fmaf(r[i].x, sc.x, sh.x), fmaf(r[i].y, sc.y, sh.y), fmaf(r[i].z, sc.z, sh.z), fmaf(r[i].w, sc.w, sh.w)};
          uint32_t hw[2], lw[2];
#pragma unroll
          for (int e = 0; e < 2; ++e) {
            // LeakyReLU with 0 <= slope <= 1 is max(v, slope * v); slope = 1: identity
            const float a = fmaxf(v[2 * e], v[2 * e] * slope), b = fmaxf(v[2 * e + 1], v[2 * e + 1] * slope);
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(a, b);   // one packed conversion per pair
            const uint32_t hbits = *reinterpret_cast<const uint32_t*>(&h2);
            const float ra_ = a - __uint_as_float(hbits << 16), rb_ = b - __uint_as_float(hbits & 0xffff0000u);
            const __nv_bfloat162 l2 = __floats2bfloat162_rn(ra_, rb_);
            hw[e] = hbits;
            lw[e] = *reinterpret_cast<const uint32_t*>(&l2);
          }
          hi = make_uint2(hw[0], hw[1]);
          lo = make_uint2(lw[0], lw[1]);
        }
        if (col_ok) {
          *reinterpret_cast<uint2*>(slot + soff[i]) = hi;
          *reinterpret_cast<uint2*>(slot + FU_APLANE + soff[i]) = lo;
        }
      }
      fence_proxy_async();
      mbar_arrive(bars + 8 * b_st);
      if (++b_cb == ncb) b_cb = 0;
      if (++b_st == NS) { b_st = 0; b_par ^= 1u; }
    };
    fetch(ra, oka);
    for (int s = 0; s < nchunks; s += 2) {
      if (s + 1 < nchunks) fetch(rb, okb);   // the next chunk's loads fly while this one is converted
      build(ra, oka);
      if (s + 1 < nchunks) {
        if (s + 2 < nchunks) fetch(ra, oka);
        build(rb, okb);
      }
    }
  } else {
    // ===================== epilogue (last four warps): TMEM lane quarter warp & 3 =====================
    const int q = warp & 3;
    const int row = q * 32 + lane;
    const int pix = pix0 + row;
    const bool row_ok = pix < HW;
    const int vy = row_ok ? pix / p.Wv : 0, vx = row_ok ? pix - (pix / p.Wv) * p.Wv : 0;
    const int oy = vy * p.os + p.py[ph], ox = vx * p.os + p.px[ph];
    float* dst = p.out + ((size_t)(img * p.Hout + oy) * p.Wout + ox) * p.out_C + p.out_coff + n0;
    mbar_wait(bars + 64, 0);
    tc_fence_after();
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
      if (row_ok) {
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (c0 + 4 * j >= BN || n0 + c0 + 4 * j >= p.CoutP) continue;   // narrow heads: 5 channels in an 8-channel buffer
          float4 o = make_float4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
          if (p.bias) {
            const float4 b = *reinterpret_cast<const float4*>(p.bias + n0 + c0 + 4 * j);
            o.x += b.x; o.y += b.y; o.z += b.z; o.w += b.w;
          }
          *reinterpret_cast<float4*>(dst + c0 + 4 * j) = o;
        }
      }
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)Cfg::TCOLS) : "memory");
  }
}

// upsample_bilinear2d(align_corners=False), scale factor 2: src = 0.5 * (dst + 0.5) - 0.5, clamped at 0
struct FLerp { int i0, i1; float l0, l1; };
__device__ __forceinline__ FLerp flerp_half(int dst, int in_size) {
  float real = __fsub_rn(__fmul_rn(0.5f, __fadd_rn((float)dst, 0.5f)), 0.5f);
  if (real < 0.f) real = 0.f;
  FLerp r;
  r.i0 = (int)real;
  r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
  r.l1 = __fsub_rn(real, (float)r.i0);
  r.l0 = __fsub_rn(1.f, r.l1);
  return r;
}
__device__ __forceinline__ FLerp flerp_ac(int dst, float scale, int in_size) {  // align_corners=True: src = scale * dst
  const float real = __fmul_rn(scale, (float)dst);
  FLerp r;
  r.i0 = (int)real;
  r.i1 = r.i0 + ((r.i0 < in_size - 1) ? 1 : 0);
  r.l1 = fminf(fmaxf(__fsub_rn(real, (float)r.i0), 0.f), 1.f);
  r.l0 = __fsub_rn(1.f, r.l1);
  return r;
}
__device__ __forceinline__ float fbilerp(float v00, float v01, float v10, float v11, const FLerp& ly, const FLerp& lx) {
  const float top = __fadd_rn(__fmul_rn(lx.l0, v00), __fmul_rn(lx.l1, v01));
  const float bot = __fadd_rn(__fmul_rn(lx.l0, v10), __fmul_rn(lx.l1, v11));
  return __fadd_rn(__fmul_rn(ly.l0, top), __fmul_rn(ly.l1, bot));
}

// heads [B,s,s,8] (channels 0,1 = flow, 2..4 = visibility) -> flow_out [B,2,R,R], vis_out [B,3,R,R] (optional) and the
// masked, rescaled flow fm [B,2,R,R] / mask mk [B,1,R,R] of flow_network_warp (geomcgt_ifw_test_model.py:69-73), R = 2s
__global__ void fpost_kernel(const float* __restrict__ heads, int B, int s, float* flow_out, float* vis_out, float* fm, float* mk) {
  const int R = 2 * s;
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= B * R * R) return;
  const int x = i % R, y = (i / R) % R, n = i / (R * R);
  const FLerp ly = flerp_half(y, s), lx = flerp_half(x, s);
  const float* h = heads + (size_t)n * s * s * 8;
  float v[5];
#pragma unroll
  for (int c = 0; c < 5; ++c)
    v[c] = fbilerp(h[((size_t)ly.i0 * s + lx.i0) * 8 + c], h[((size_t)ly.i0 * s + lx.i1) * 8 + c],
                   h[((size_t)ly.i1 * s + lx.i0) * 8 + c], h[((size_t)ly.i1 * s + lx.i1) * 8 + c], ly, lx);
  const size_t plane = (size_t)R * R, o = (size_t)y * R + x;
  if (flow_out) { flow_out[((size_t)n * 2 + 0) * plane + o] = v[0]; flow_out[((size_t)n * 2 + 1) * plane + o] = v[1]; }
  if (vis_out) {
    vis_out[((size_t)n * 3 + 0) * plane + o] = v[2];
    vis_out[((size_t)n * 3 + 1) * plane + o] = v[3];
    vis_out[((size_t)n * 3 + 2) * plane + o] = v[4];
  }
  // argmax over the three visibility classes (first maximum wins, like torch.argmax); mask = class < 2
  int am = 0;
  float best = v[2];
  if (v[3] > best) { best = v[3]; am = 1; }
  if (v[4] > best) { am = 2; }
  const float mask = am < 2 ? 1.f : 0.f;
  if (fm) {
    // flow_out * 20. * mask, then / 7 * 8
    fm[((size_t)n * 2 + 0) * plane + o] = __fmul_rn(__fdiv_rn(__fmul_rn(__fmul_rn(v[0], 20.f), mask), 7.f), 8.f);
    fm[((size_t)n * 2 + 1) * plane + o] = __fmul_rn(__fdiv_rn(__fmul_rn(__fmul_rn(v[1], 20.f), mask), 7.f), 8.f);
    mk[(size_t)n * plane + o] = mask;
  }
}

// F.interpolate(x, size=(256,256), mode='bilinear', align_corners=True) of NCHW planes
__global__ void fresize_kernel(const float* __restrict__ src, int planes, int R, float* dst) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= planes * 65536) return;
  const int x = i & 255, y = (i >> 8) & 255, pl = i >> 16;
  const float scale = (float)(R - 1) / 255.f;
  const FLerp ly = flerp_ac(y, scale, R), lx = flerp_ac(x, scale, R);
  const float* s = src + (size_t)pl * R * R;
  dst[i] = fbilerp(s[(size_t)ly.i0 * R + lx.i0], s[(size_t)ly.i0 * R + lx.i1], s[(size_t)ly.i1 * R + lx.i0],
                   s[(size_t)ly.i1 * R + lx.i1], ly, lx);
}

}  // namespace ap

using namespace ap;

// =================================================================================================
// host side
// =================================================================================================
namespace {

// host-side bf16 round-to-nearest-even (what __float2bfloat16_rn does on the device)
inline uint16_t f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u);  // NaN stays NaN
  return (uint16_t)((u + 0x7fffu + ((u >> 16) & 1u)) >> 16);
}
inline float bf16_to_f32(uint16_t h) {
  const uint32_t u = (uint32_t)h << 16;
  float f;
  memcpy(&f, &u, 4);
  return f;
}

struct FLayerW {
  float* w = nullptr;     // [slab][Cin][CoutP]
  float* bias = nullptr;  // [Cout] applied in the conv epilogue, or null
  float* nsc = nullptr;   // BatchNorm: per-channel scale / shift of the OUTPUT of this conv (bias folded in)
  float* nsh = nullptr;
  int cout = 0, cin = 0, k = 0, coutp = 0;
  bool transposed = false;
  uint8_t* wimg = nullptr;  // tensor-core path: pre-swizzled bf16 hi/lo image (fconv_umma_kernel), or null
  int bn = 0, kc = 0, wrows = 0;  // its N tile, channels per K chunk, rows per weight plane
};

struct FBuf {
  float* p = nullptr;
  float* scale = nullptr;  // [B][C]
  float* shift = nullptr;
  int H = 0, C = 0;
};

}  // namespace

struct ap_flow {
  int input_nc = 136, nf = 16, start_scale = 2, num_scale = 4, norm = 0, max_nf = 512, size = 224, device = 0;
  int ndown = 0;                      // stride-2 convs of conv_downsample
  std::vector<int> outer, inner, sz;  // per level: channels, input size of the level
  std::map<std::string, FLayerW> w;
  std::vector<void*> owned;
  bool loaded = false;
  // workspace of the current batch size
  int planB = 0;
  std::vector<void*> ws;
  std::vector<FBuf> cds;   // conv_downsample outputs
  std::vector<FBuf> cat;   // cat[l] (l >= 1): [d_{l-1} | u_l]; cat[0] = u_0
  FBuf dinner, heads;
  float *fm = nullptr, *mk = nullptr;
  float* kpbuf = nullptr;  // [B, input_nc, size, size] key-point maps made by ap_flow_warp_landmarks (allocated on first use)
  int kpbufB = 0;
  int4* bbox = nullptr;  // [B * input_nc] non-zero bounding boxes of the operand planes of the first conv
  int* area = nullptr;   // [B] their summed areas
  int sparse = 1, tiled = 1;  // AP_FLOW_SPARSE / AP_FLOW_TILED = 0: the one generic kernel everywhere (A/B, tests)
  int umma = 1;               // AP_FLOW_UMMA = 0: no tensor-core convs (fp32 FFMA kernels only)
  int bn256 = 1;              // AP_FLOW_BN256 = 0: N tile of 128 even where Cout % 256 == 0 (A/B: 256 halves the operand builds)
  int64_t last_launches = 0;
};

static std::string level_prefix(int l) {
  std::string s = "unet_block.";
  for (int i = 0; i < l; ++i) s += "submodule.";
  return s;
}

static void flow_free_ws(ap_flow* h) {
  for (void* p : h->ws) cudaFree(p);
  h->ws.clear();
  h->cds.clear();
  h->cat.clear();
  h->planB = 0;
  h->kpbuf = nullptr;
  h->kpbufB = 0;
}

static int flow_alloc(ap_flow* h, size_t bytes, void** p) {
  AP_CUDA(cudaMalloc(p, bytes));
  h->ws.push_back(*p);
  return AP_OK;
}

static int flow_buf(ap_flow* h, int B, int H, int C, FBuf* b) {
  b->H = H; b->C = C;
  AP_TRY(flow_alloc(h, (size_t)B * H * H * C * 4, (void**)&b->p));
  AP_TRY(flow_alloc(h, (size_t)B * C * 4, (void**)&b->scale));
  AP_TRY(flow_alloc(h, (size_t)B * C * 4, (void**)&b->shift));
  return AP_OK;
}

// Workspace for batches of up to B images.  Every buffer is image-major, so a smaller batch simply uses the front of it:
// the ragged last batch of a clip (733 = 11 x 64 + 29) must not cost a device synchronisation and gigabytes of
// cudaFree / cudaMalloc twice per clip (measured: 36 ms per batch instead of 6).
static int flow_plan(ap_flow* h, int B) {
  if (h->planB >= B) return AP_OK;
  AP_CUDA(cudaDeviceSynchronize());
  flow_free_ws(h);
  const int L = h->num_scale;
  int s = h->size, c = h->nf;
  h->cds.resize(h->ndown + 1);
  AP_TRY(flow_buf(h, B, s, c, &h->cds[0]));
  for (int i = 0; i < h->ndown; ++i) {
    s /= 2; c *= 2;
    AP_TRY(flow_buf(h, B, s, c, &h->cds[i + 1]));
  }
  h->cat.resize(L);
  for (int l = 0; l < L; ++l) AP_TRY(flow_buf(h, B, h->sz[l], l == 0 ? h->outer[0] : 2 * h->outer[l], &h->cat[l]));
  AP_TRY(flow_buf(h, B, h->sz[L - 1] / 2, h->inner[L - 1], &h->dinner));
  AP_TRY(flow_buf(h, B, h->sz[0], 8, &h->heads));
  const int R = 2 * h->sz[0];
  AP_TRY(flow_alloc(h, (size_t)B * 2 * R * R * 4, (void**)&h->fm));
  AP_TRY(flow_alloc(h, (size_t)B * R * R * 4, (void**)&h->mk));
  AP_TRY(flow_alloc(h, (size_t)B * h->input_nc * sizeof(int4), (void**)&h->bbox));
  AP_TRY(flow_alloc(h, (size_t)B * sizeof(int), (void**)&h->area));
  h->planB = B;
  return AP_OK;
}

// taps of a k x k conv with padding `pad`
static void taps_conv(FConvP* p, int k, int pad) {
  p->nphase = 1; p->ntaps = k * k; p->os = 1; p->py[0] = p->px[0] = 0;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) p->taps[0][ky * k + kx] = FTap{(int8_t)(ky - pad), (int8_t)(kx - pad), (uint8_t)(ky * k + kx)};
}
// ConvTranspose2d(k4, s2, p1): output (2y + py, 2x + px) = sum over input rows y + dy with kernel row ky:
//   py = 0: (ky 1, dy 0), (ky 3, dy -1);   py = 1: (ky 2, dy 0), (ky 0, dy +1); same along x
static void taps_convT4(FConvP* p) {
  static const int KY[2][2] = {{1, 3}, {2, 0}}, DY[2][2] = {{0, -1}, {0, 1}};
  p->nphase = 4; p->ntaps = 4; p->os = 2;
  for (int ph = 0; ph < 4; ++ph) {
    const int py = ph >> 1, px = ph & 1;
    p->py[ph] = (int8_t)py; p->px[ph] = (int8_t)px;
    for (int a = 0; a < 2; ++a)
      for (int b = 0; b < 2; ++b)
        p->taps[ph][a * 2 + b] = FTap{(int8_t)DY[py][a], (int8_t)DY[px][b], (uint8_t)(KY[py][a] * 4 + KY[px][b])};
  }
}

static int flow_conv(ap_flow* h, int B, const float* in, int in_nchw, int in_C, int in_coff, const float* scale,
                     const float* shift, int act, int Hin, const FLayerW& w, int stride, int transposed4, int pad, float* out,
                     int out_C, int out_coff, int Hout, cudaStream_t st, const int* gate_area = nullptr, int gate_thresh = 0) {
  FConvP p;
  memset(&p, 0, sizeof(p));
  p.in = in; p.in_nchw = in_nchw; p.in_C = in_C; p.in_coff = in_coff; p.scale = scale; p.shift = shift; p.act = act;
  p.w = w.w; p.bias = w.bias; p.out = out; p.out_C = out_C; p.out_coff = out_coff;
  p.B = B; p.Hin = Hin; p.Win = Hin; p.Cin = w.cin; p.Cout = w.cout; p.CoutP = w.coutp;
  p.Hout = Hout; p.Wout = Hout;
  if (transposed4) {
    taps_convT4(&p);
    p.Hv = Hin; p.Wv = Hin; p.stride = 1;
  } else {
    taps_conv(&p, w.k, pad);
    p.Hv = Hout; p.Wv = Hout; p.stride = stride;
  }
  const int M = B * p.Hv * p.Wv;
  if (gate_area) { p.gate_area = gate_area; p.gate_thresh = gate_thresh; }
  const bool aligned4 = in_C % 4 == 0 && in_coff % 4 == 0 && out_C % 4 == 0 && out_coff % 4 == 0;
  if (h->umma && w.wimg && !in_nchw && !gate_area && aligned4 && w.cin <= FU_MAXC) {
    FUmmaP up;
    up.c = p;
    up.wimg = w.wimg;
    up.tiles_per_img = (p.Hv * p.Wv + 127) / 128;
    up.kc = w.kc;
    up.wrows = w.wrows;
    const dim3 grid((unsigned)(B * up.tiles_per_img), (unsigned)((w.cout + w.bn - 1) / w.bn), (unsigned)p.nphase);
    if (w.bn == 256) fconv_umma_kernel<256><<<grid, FU_THREADS, FUCfg<256>::SMEM, st>>>(up);
    else if (w.bn == 128) fconv_umma_kernel<128><<<grid, FU_THREADS, FUCfg<128>::SMEM, st>>>(up);
    else if (w.bn == 64) fconv_umma_kernel<64><<<grid, FU_THREADS, FUCfg<64>::SMEM, st>>>(up);
    else fconv_umma_kernel<16><<<grid, FU_THREADS, FUCfg<16>::SMEM, st>>>(up);
  } else if (h->tiled && !in_nchw && !gate_area && w.cin % FT_BK == 0 && in_C % 4 == 0 && in_coff % 4 == 0 && out_C % 4 == 0 &&
      out_coff % 4 == 0) {
    (void)aligned4;
    const int mt = (M + FT_BM - 1) / FT_BM;
    if (w.cout > 64) fconv_tiled_kernel<128><<<dim3(mt, (w.cout + 127) / 128, p.nphase), 256, 0, st>>>(p);
    else if (w.cout > 16) fconv_tiled_kernel<64><<<dim3(mt, 1, p.nphase), 256, 0, st>>>(p);
    else fconv_tiled_kernel<16><<<dim3(mt, 1, p.nphase), 256, 0, st>>>(p);
  } else {
    dim3 grid((M + 63) / 64, (w.cout + 63) / 64, p.nphase);
    fconv_kernel<<<grid, 256, 0, st>>>(p);
  }
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

// The first conv on a sparse NCHW operand (stride 1): boxes of the non-zeros, the sparse kernel for the images that are
// sparse and the generic kernel, gated per image, for those that are not.
static int flow_conv_sparse(ap_flow* h, int B, const float* in, int Hin, const FLayerW& w, int pad, float* out, int out_C,
                            int Hout, cudaStream_t st, bool have_boxes) {
  if (!have_boxes) {  // (ap_flow_warp_landmarks knows the boxes of its discs from the coordinates)
    AP_CUDA(cudaMemsetAsync(h->area, 0, (size_t)B * sizeof(int), st));
    fbbox_kernel<<<B * w.cin, 256, 0, st>>>(in, Hin, Hin, w.cin, h->bbox, h->area);
    AP_CUDA(cudaGetLastError());
    launches_add(1);
  }
  FSparseP sp;
  memset(&sp, 0, sizeof(sp));
  sp.in = in; sp.bbox = h->bbox; sp.area = h->area;
  sp.thresh = (int)((long long)Hin * Hin * w.cin / 8);
  sp.w = w.w; sp.bias = w.bias; sp.out = out; sp.out_C = out_C; sp.out_coff = 0;
  sp.B = B; sp.H = Hin; sp.W = Hin; sp.Cin = w.cin; sp.Cout = w.cout; sp.CoutP = w.coutp; sp.k = w.k; sp.pad = pad;
  const int tiles = (Hin + 15) / 16;
  if (w.cout > 16) fconv_sparse_kernel<32><<<dim3(tiles * tiles, (w.cout + 31) / 32, B), 256, 0, st>>>(sp);
  else fconv_sparse_kernel<16><<<dim3(tiles * tiles, 1, B), 256, 0, st>>>(sp);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return flow_conv(h, B, in, 1, w.cin, 0, nullptr, nullptr, FACT_NONE, Hin, w, 1, 0, pad, out, out_C, 0, Hout, st, h->area,
                   sp.thresh);
}

// operand transform of buffer channels [coff, coff + nC): the producing conv's normalisation
static int flow_norm(ap_flow* h, int B, const FBuf& b, int coff, int nC, const FLayerW& w, cudaStream_t st) {
  if (h->norm == 0) {
    fbroadcast_kernel<<<(B * nC + 255) / 256, 256, 0, st>>>(w.nsc, w.nsh, b.C, coff, nC, B, b.scale, b.shift);
  } else {
    fstats_kernel<<<dim3((nC + 31) / 32, B), 256, 0, st>>>(b.p, b.H * b.H, b.C, coff, nC, b.scale, b.shift);
  }
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

extern "C" {

int ap_flow_create(ap_flow** handle, int input_nc, int nf, int start_scale, int num_scale, int norm, int max_nf, int size,
                   int device) {
  AP_REQUIRE(handle != nullptr, AP_ERR_INVALID, "null handle pointer");
  AP_REQUIRE(input_nc >= 1 && nf >= 4 && nf % 4 == 0 && num_scale >= 1 && num_scale <= 8 && (norm == 0 || norm == 1) &&
                 max_nf >= nf && size >= 8,
             AP_ERR_INVALID, "FlowUnet(input_nc=%d, nf=%d, num_scale=%d, norm=%d, max_nf=%d, size=%d)", input_nc, nf, num_scale,
             norm, max_nf, size);
  AP_REQUIRE(start_scale == 1 || start_scale == 2 || start_scale == 4 || start_scale == 8, AP_ERR_INVALID, "start_scale=%d",
             start_scale);
  int ndev = 0;
  AP_CUDA(cudaGetDeviceCount(&ndev));
  AP_REQUIRE(device >= 0 && device < ndev, AP_ERR_INVALID, "device %d of %d", device, ndev);
  ap_flow* h = new ap_flow();
  h->input_nc = input_nc; h->nf = nf; h->start_scale = start_scale; h->num_scale = num_scale; h->norm = norm;
  h->max_nf = max_nf; h->size = size; h->device = device;
  { const char* e = getenv("AP_FLOW_SPARSE"); if (e && e[0] == '0') h->sparse = 0; }
  { const char* e = getenv("AP_FLOW_TILED"); if (e && e[0] == '0') h->tiled = 0; }
  { const char* e = getenv("AP_FLOW_UMMA"); if (e && e[0] == '0') h->umma = 0; }
  { const char* e = getenv("AP_FLOW_BN256"); if (e) h->bn256 = e[0] == '1'; }
  if (h->umma) {  // function attributes are per device: set them for this handle's device
    int prev = 0;
    cudaGetDevice(&prev);
    cudaError_t e1 = cudaSetDevice(device);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(fconv_umma_kernel<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUCfg<64>::SMEM);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(fconv_umma_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUCfg<128>::SMEM);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(fconv_umma_kernel<16>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUCfg<16>::SMEM);
    if (e1 == cudaSuccess) e1 = cudaFuncSetAttribute(fconv_umma_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)FUCfg<256>::SMEM);
    cudaSetDevice(prev);
    if (e1 != cudaSuccess) {
      delete h;
      set_error("FlowUnet: tcgen05 kernels unavailable on device %d: %s", device, cudaGetErrorString(e1));
      cudaGetLastError();
      return AP_ERR_UNSUPPORTED;
    }
  }
  int s = size, nc = nf;
  for (int sc = start_scale; sc > 1; sc /= 2) {
    s = (s + 2 - 3) / 2 + 1;
    nc *= 2;
    ++h->ndown;
  }
  for (int l = 0; l < num_scale; ++l) {
    if (s % 2 != 0 || s < 2) {
      delete h;
      set_error("FlowUnet: level %d works on %dx%d, which a 4x4 stride-2 conv and its transpose do not round-trip (the reference "
                "fails in torch.cat for this configuration)", l, s, s);
      return AP_ERR_UNSUPPORTED;
    }
    long long o = (long long)nc << l, i = (long long)nc << (l + 1);
    h->outer.push_back((int)(o < max_nf ? o : max_nf));
    h->inner.push_back((int)(i < max_nf ? i : max_nf));
    h->sz.push_back(s);
    s /= 2;
  }
  *handle = h;
  return AP_OK;
}

int ap_flow_destroy(ap_flow* h) {
  if (!h) return AP_OK;
  cudaSetDevice(h->device);
  cudaDeviceSynchronize();
  flow_free_ws(h);
  for (void* p : h->owned) cudaFree(p);
  delete h;
  return AP_OK;
}

int ap_flow_load_weights(ap_flow* h, int n, const char* const* names, const float* const* ptrs, const int64_t* shapes,
                         int on_device, void* cuda_stream) {
  AP_REQUIRE(h && names && ptrs && shapes, AP_ERR_INVALID, "null argument");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  std::map<std::string, int> idx;
  for (int i = 0; i < n; ++i) idx[names[i]] = i;
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear();
  h->w.clear();
  h->loaded = false;
  const bool inorm = h->norm == 1;
  std::vector<void*> staging;
  auto host_copy = [&](const std::string& key, size_t elems, std::vector<float>* out) -> int {
    auto it = idx.find(key);
    AP_REQUIRE(it != idx.end(), AP_ERR_INVALID, "missing key %s", key.c_str());
    const int64_t* sh = shapes + 4 * it->second;
    AP_REQUIRE((size_t)(sh[0] * sh[1] * sh[2] * sh[3]) == elems, AP_ERR_INVALID, "size mismatch for %s", key.c_str());
    out->resize(elems);
    if (on_device) AP_CUDA(cudaMemcpy(out->data(), ptrs[it->second], elems * 4, cudaMemcpyDeviceToHost));
    else memcpy(out->data(), ptrs[it->second], elems * 4);
    return AP_OK;
  };
  auto upload = [&](const std::vector<float>& v, float** dst) -> int {
    AP_CUDA(cudaMalloc((void**)dst, v.size() * 4));
    h->owned.push_back(*dst);
    AP_CUDA(cudaMemcpy(*dst, v.data(), v.size() * 4, cudaMemcpyHostToDevice));
    return AP_OK;
  };
  // tensor-core operand image of a packed fp32 weight [slab][cin][coutp] (fconv_umma_kernel), where the layer qualifies:
  // [slab][cin/kc][plane hi|lo][wrows][64 k] bf16, 8-row groups of 1024 bytes, the 16-byte groups of a row XOR-swizzled with
  // the row index (the K-major SWIZZLE_128B layout tcgen05.mma reads)
  auto make_wimg = [&](const std::vector<float>& packed, int nslab, int cin, int cout, int coutp, FLayerW* lw) -> int {
    const int kc = cin % 64 == 0 ? 64 : (cin % 32 == 0 ? 32 : 0);
    const int bn = (h->bn256 && cout % 256 == 0) ? 256 : (cout % 128 == 0 ? 128 : (cout == 64 ? 64 : (cout <= 16 ? 16 : 0)));
    if (kc == 0 || bn == 0 || cin > FU_MAXC) return AP_OK;
    const int wrows = bn == 16 ? 16 : cout;
    const int ncb = cin / kc;
    std::vector<uint16_t> img((size_t)nslab * ncb * 2 * wrows * 64, 0);
    for (int sl = 0; sl < nslab; ++sl)
      for (int ci = 0; ci < cin; ++ci)
        for (int co = 0; co < cout; ++co) {
          const float wv = packed[((size_t)sl * cin + ci) * coutp + co];
          const uint16_t hi = f32_to_bf16(wv);
          const uint16_t lo = f32_to_bf16(wv - bf16_to_f32(hi));
          const int cb = ci / kc, kk = ci - cb * kc;
          const size_t row = (size_t)(co >> 3) * 1024 + (size_t)(co & 7) * 128 + (size_t)(((kk >> 3) ^ (co & 7)) << 4) + (size_t)(kk & 7) * 2;
          const size_t base = (((size_t)sl * ncb + cb) * 2) * (size_t)wrows * 128;
          img[(base + row) / 2] = hi;
          img[(base + (size_t)wrows * 128 + row) / 2] = lo;
        }
    AP_CUDA(cudaMalloc((void**)&lw->wimg, img.size() * 2));
    h->owned.push_back(lw->wimg);
    AP_CUDA(cudaMemcpy(lw->wimg, img.data(), img.size() * 2, cudaMemcpyHostToDevice));
    lw->bn = bn; lw->kc = kc; lw->wrows = wrows;
    return AP_OK;
  };
  // conv `key` (Cout x Cin x k x k, or transposed Cin x Cout x k x k); `normkey` = its BatchNorm ("" = no norm after it);
  // bias_mode: 0 none, 1 epilogue bias (no norm follows), 2 bias in the checkpoint folds into the norm (or cancels)
  auto add = [&](const std::string& name, const std::string& key, int cout, int cin, int k, bool transposed,
                 const std::string& normkey, bool has_bias) -> int {
    FLayerW lw;
    lw.cout = cout; lw.cin = cin; lw.k = k; lw.transposed = transposed; lw.coutp = (cout + 3) / 4 * 4;
    std::vector<float> wt, packed((size_t)k * k * cin * lw.coutp, 0.f), bias;
    AP_TRY(host_copy(key + ".weight", (size_t)cout * cin * k * k, &wt));
    for (int co = 0; co < cout; ++co)
      for (int ci = 0; ci < cin; ++ci)
        for (int sl = 0; sl < k * k; ++sl) {
          const size_t s_ = transposed ? (((size_t)ci * cout + co) * k * k + sl) : (((size_t)co * cin + ci) * k * k + sl);
          packed[((size_t)sl * cin + ci) * lw.coutp + co] = wt[s_];
        }
    AP_TRY(upload(packed, &lw.w));
    if (h->umma) AP_TRY(make_wimg(packed, k * k, cin, cout, lw.coutp, &lw));
    if (has_bias) AP_TRY(host_copy(key + ".bias", cout, &bias));
    if (normkey.empty()) {
      if (has_bias) { bias.resize(lw.coutp, 0.f); AP_TRY(upload(bias, &lw.bias)); }  // vector reads of the epilogues
    } else if (!inorm) {
      // BatchNorm2d in eval mode: y = x * (gamma * invstd) + (beta - mean * gamma * invstd); a conv bias folds into the shift
      std::vector<float> g, b, mu, var, sc(cout), sh(cout);
      AP_TRY(host_copy(normkey + ".weight", cout, &g));
      AP_TRY(host_copy(normkey + ".bias", cout, &b));
      AP_TRY(host_copy(normkey + ".running_mean", cout, &mu));
      AP_TRY(host_copy(normkey + ".running_var", cout, &var));
      for (int c = 0; c < cout; ++c) {
        const float invstd = 1.0f / sqrtf(var[c] + 1e-5f);
        sc[c] = g[c] * invstd;
        sh[c] = b[c] - mu[c] * sc[c] + (has_bias ? bias[c] * sc[c] : 0.f);
      }
      AP_TRY(upload(sc, &lw.nsc));
      AP_TRY(upload(sh, &lw.nsh));
    }  // InstanceNorm: the bias cancels, the statistics come from the data
    h->w[name] = lw;
    return AP_OK;
  };
  int nc = h->nf;
  AP_TRY(add("cds0", "conv_downsample.0", h->nf, h->input_nc, 7, false, "conv_downsample.1", inorm));
  for (int i = 0; i < h->ndown; ++i) {
    AP_TRY(add("cds" + std::to_string(i + 1), "conv_downsample." + std::to_string(3 * (i + 1)), 2 * nc, nc, 3, false,
               "conv_downsample." + std::to_string(3 * (i + 1) + 1), inorm));
    nc *= 2;
  }
  const int L = h->num_scale;
  for (int l = 0; l < L; ++l) {
    const std::string pre = level_prefix(l);
    const bool outermost = l == 0, innermost = l == L - 1;
    const std::string dk = pre + "down." + std::to_string(outermost ? 0 : 1);
    AP_TRY(add("down" + std::to_string(l), dk, h->inner[l], h->outer[l], 4, false,
               innermost ? "" : pre + "down." + std::to_string(outermost ? 1 : 2), inorm));
    AP_TRY(add("up" + std::to_string(l), pre + "up.1", h->outer[l], innermost ? h->inner[l] : 2 * h->inner[l], 4, true,
               pre + "up.2", outermost ? true : inorm));
  }
  // the two heads the caller uses (flow at the outermost level, visibility) as ONE conv with 5 output channels
  {
    std::vector<float> wf, wv, bf, bv;
    const int cin = h->outer[0];
    AP_TRY(host_copy("unet_block.predict_flow.1.weight", (size_t)2 * cin * 9, &wf));
    AP_TRY(host_copy("unet_block.predict_flow.1.bias", 2, &bf));
    AP_TRY(host_copy("predict_vis.1.weight", (size_t)3 * cin * 9, &wv));
    AP_TRY(host_copy("predict_vis.1.bias", 3, &bv));
    FLayerW lw;
    lw.cout = 5; lw.cin = cin; lw.k = 3; lw.coutp = 8;
    std::vector<float> packed((size_t)9 * cin * 8, 0.f), bias(8, 0.f);
    for (int co = 0; co < 5; ++co) {
      const std::vector<float>& src = co < 2 ? wf : wv;
      const int c_ = co < 2 ? co : co - 2;
      for (int ci = 0; ci < cin; ++ci)
        for (int sl = 0; sl < 9; ++sl) packed[((size_t)sl * cin + ci) * 8 + co] = src[((size_t)c_ * cin + ci) * 9 + sl];
      bias[co] = co < 2 ? bf[co] : bv[co - 2];
    }
    AP_TRY(upload(packed, &lw.w));
    AP_TRY(upload(bias, &lw.bias));
    if (h->umma) AP_TRY(make_wimg(packed, 9, cin, 5, 8, &lw));
    h->w["heads"] = lw;
  }
  (void)st;
  flow_free_ws(h);
  h->loaded = true;
  return AP_OK;
}

}  // extern "C"

// the network on key-point maps already on the device; have_boxes: h->bbox / h->area describe them already
static int flow_forward_impl(ap_flow* h, int B, const float* kp_maps, bool have_boxes, float* flow_out, float* vis_out,
                             float* iw_flow, float* if_mask, cudaStream_t st, int64_t before) {
  const int L = h->num_scale;
  // conv_downsample: conv7x7 + norm + LeakyReLU(0.1), then 3x3 stride-2 convs (networks.py:601-612)
  if (h->sparse && h->input_nc <= FS_MAXC && h->size == (h->size + 2 * 3 - 7) + 1) {
    AP_TRY(flow_conv_sparse(h, B, kp_maps, h->size, h->w.at("cds0"), 3, h->cds[0].p, h->cds[0].C, h->size, st, have_boxes));
  } else {
    AP_TRY(flow_conv(h, B, kp_maps, 1, h->input_nc, 0, nullptr, nullptr, FACT_NONE, h->size, h->w.at("cds0"), 1, 0, 3,
                     h->cds[0].p, h->cds[0].C, 0, h->size, st));
  }
  AP_TRY(flow_norm(h, B, h->cds[0], 0, h->cds[0].C, h->w.at("cds0"), st));
  for (int i = 0; i < h->ndown; ++i) {
    const FBuf& src = h->cds[i];
    const FBuf& dst = h->cds[i + 1];
    const FLayerW& w = h->w.at("cds" + std::to_string(i + 1));
    AP_TRY(flow_conv(h, B, src.p, 0, src.C, 0, src.scale, src.shift, FACT_LRELU01, src.H, w, 2, 0, 1, dst.p, dst.C, 0, dst.H, st));
    AP_TRY(flow_norm(h, B, dst, 0, dst.C, w, st));
  }
  // encoder: level l reads x_l (x_0 = LeakyReLU_0.1(norm(.)); x_l = LeakyReLU_0.2(norm(d_{l-1})) -- the in-place activation
  // at the head of `down`, networks.py:523) and writes d_l into the first half of the next level's concatenation buffer
  for (int l = 0; l < L; ++l) {
    const FBuf& src = l == 0 ? h->cds[h->ndown] : h->cat[l];
    const FLayerW& w = h->w.at("down" + std::to_string(l));
    const bool innermost = l == L - 1;
    const FBuf& dst = innermost ? h->dinner : h->cat[l + 1];
    AP_TRY(flow_conv(h, B, src.p, 0, src.C, 0, src.scale, src.shift, l == 0 ? FACT_LRELU01 : FACT_LRELU02, src.H, w, 2, 0, 1,
                     dst.p, dst.C, 0, dst.H, st));
    if (!innermost) AP_TRY(flow_norm(h, B, dst, 0, w.cout, w, st));
  }
  // decoder: u_l = norm(ConvT(ReLU(.))) of the whole concatenation of the level below (or of the innermost d)
  for (int l = L - 1; l >= 0; --l) {
    const bool innermost = l == L - 1;
    const FBuf& src = innermost ? h->dinner : h->cat[l + 1];
    const FLayerW& w = h->w.at("up" + std::to_string(l));
    const FBuf& dst = h->cat[l];
    const int coff = l == 0 ? 0 : h->outer[l];
    AP_TRY(flow_conv(h, B, src.p, 0, src.C, 0, innermost ? nullptr : src.scale, innermost ? nullptr : src.shift, FACT_RELU,
                     src.H, w, 1, 1, 0, dst.p, dst.C, coff, dst.H, st));
    AP_TRY(flow_norm(h, B, dst, coff, w.cout, w, st));
  }
  // heads on LeakyReLU_0.1(u_0): predict_flow of the outermost block + predict_vis (networks.py:571-573, 622-625)
  AP_TRY(flow_conv(h, B, h->cat[0].p, 0, h->cat[0].C, 0, h->cat[0].scale, h->cat[0].shift, FACT_LRELU01, h->cat[0].H,
                   h->w.at("heads"), 1, 0, 1, h->heads.p, 8, 0, h->heads.H, st));
  const int s = h->sz[0], R = 2 * s;
  fpost_kernel<<<(B * R * R + 255) / 256, 256, 0, st>>>(h->heads.p, B, s, flow_out, vis_out, iw_flow ? h->fm : nullptr, h->mk);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  if (iw_flow) {
    fresize_kernel<<<(B * 2 * 65536 + 255) / 256, 256, 0, st>>>(h->fm, B * 2, R, iw_flow);
    fresize_kernel<<<(B * 65536 + 255) / 256, 256, 0, st>>>(h->mk, B, R, if_mask);
    AP_CUDA(cudaGetLastError());
    launches_add(2);
  }
  h->last_launches = launches_get() - before;
  return AP_OK;
}

extern "C" {

int ap_flow_forward(ap_flow* h, int B, const float* kp_maps, float* flow_out, float* vis_out, float* iw_flow, float* if_mask,
                    void* cuda_stream) {
  AP_REQUIRE(h != nullptr && h->loaded, AP_ERR_STATE, "forward before load_weights");
  AP_REQUIRE(B >= 1 && kp_maps, AP_ERR_INVALID, "bad argument");
  AP_REQUIRE((iw_flow == nullptr) == (if_mask == nullptr), AP_ERR_INVALID, "iw_flow and if_mask come together");
  AP_CUDA(cudaSetDevice(h->device));
  AP_TRY(flow_plan(h, B));
  return flow_forward_impl(h, B, kp_maps, false, flow_out, vis_out, iw_flow, if_mask, (cudaStream_t)cuda_stream, launches_get());
}

int ap_flow_warp_landmarks(ap_flow* h, int B, const float* lm1, int lm1_per_frame, const float* lm2, float* iw_flow,
                           float* if_mask, void* cuda_stream) {
  AP_REQUIRE(h != nullptr && h->loaded, AP_ERR_STATE, "forward before load_weights");
  AP_REQUIRE(B >= 1 && lm1 && lm2 && iw_flow && if_mask, AP_ERR_INVALID, "bad argument");
  AP_REQUIRE(h->input_nc % 2 == 0, AP_ERR_UNSUPPORTED, "input_nc = %d is not two sets of key points", h->input_nc);
  AP_CUDA(cudaSetDevice(h->device));
  AP_TRY(flow_plan(h, B));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  if (h->kpbufB < B) {  // the workspace owns the buffer (flow_free_ws releases it); sized like the plan, for its largest batch
    AP_CUDA(cudaStreamSynchronize(st));
    AP_TRY(flow_alloc(h, (size_t)h->planB * h->input_nc * h->size * h->size * 4, (void**)&h->kpbuf));
    h->kpbufB = h->planB;
  }
  const int64_t before = launches_get();
  const int K = h->input_nc / 2;
  const bool boxes = h->sparse && h->input_nc <= FS_MAXC;
  AP_CUDA(cudaMemsetAsync(h->area, 0, (size_t)B * sizeof(int), st));
  AP_TRY(launch_kp_half(lm1, lm1_per_frame, B, K, h->size, 4.f, h->kpbuf, h->input_nc, 0, h->bbox, h->area, st));
  AP_TRY(launch_kp_half(lm2, 1, B, K, h->size, 4.f, h->kpbuf, h->input_nc, K, h->bbox, h->area, st));
  return flow_forward_impl(h, B, h->kpbuf, boxes, nullptr, nullptr, iw_flow, if_mask, st, before);
}

int ap_flow_last_launch_count(ap_flow* h, int64_t* count) {
  AP_REQUIRE(h && count, AP_ERR_INVALID, "null argument");
  *count = h->last_launches;
  return AP_OK;
}

}  // extern "C"
