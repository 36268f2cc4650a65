// Landmark branch `model_landmark_trans` (Module2/models/networks.py:1280-1282), applied to land1 and
// land2 (networks.py:1331-1332):  Conv3x3 1->8 +IN+ReLU ; Conv3x3 s2 8->16 +IN+ReLU ; Conv3x3 s2 16->16 +IN.
//
// 0.13 GFLOP per frame but 13.6 MB of traffic: HBM-bound, far too thin for tensor cores (Cout <= 16).
// One direct-convolution kernel template, one thread per output pixel holding all COUT accumulators;
// both landmark maps go through the same launch as a batch of 2B images.  The InstanceNorm + ReLU of
// the PREVIOUS layer is applied while the taps are loaded (raw conv outputs + per-(n,c) sums are what
// travels through HBM, each tensor written once and read once); zero padding is applied after the
// normalisation, as in the reference.  Biases cancel in the affine-less InstanceNorm.
#include <string.h>

#include "common.cuh"

namespace ap {

void launches_add(int n);
bool carveout_enabled();

// The weights travel BY VALUE in the kernel parameters ([9][CIN][COUT] fp32, at most 9 KB): with the tap /
// channel loops fully unrolled every FMA reads its weight straight from the constant bank, no load at all.
template <int CIN, int COUT>
struct LandP {
  const IoPtrs* io;   // CIN == 1: io->land1 [B1,1,256,256], io->land2 [B2,1,256,256] of the caller
  const float* in0;   // CIN > 1: raw NHWC [B1+B2,Hin,Win,CIN]
  const stat_t* in_stats;  // [B1+B2][CIN][2] (CIN > 1), fixed point
  float* out;         // raw NHWC [B1+B2,Hout,Wout,COUT]
  stat_t* out_stats;  // [B1+B2][COUT][2], fixed point
  int B, Hin, Hout;   // B = B1, the number of land1 maps (the batch size, or 1 in clip mode: one source landmark map)
  float w[9 * CIN * COUT];
};

template <int CIN, int COUT, int STRIDE>
__global__ void __launch_bounds__(256) land_conv_kernel(const __grid_constant__ LandP<CIN, COUT> p) {
  __shared__ float s_mean[CIN], s_rstd[CIN];
  __shared__ float s_red[8][2 * COUT];
  const int tid = threadIdx.x;
  const int HWo = p.Hout * p.Hout;
  const int gpix = blockIdx.x * 256 + tid;  // Hout*Hout is a multiple of 256: one image per CTA
  const int n = gpix / HWo;
  const int pix = gpix - n * HWo;
  const int oy = pix / p.Hout, ox = pix - oy * p.Hout;
  if (CIN > 1 && tid < CIN) {
    const double inv = 1.0 / (double)(p.Hin * p.Hin);
    const double su = (double)p.in_stats[((size_t)n * CIN + tid) * 2 + 0] * STAT_INV_SCALE;
    const double sq = (double)p.in_stats[((size_t)n * CIN + tid) * 2 + 1] * STAT_INV_SCALE;
    const double m = su * inv;
    double var = sq * inv - m * m;
    if (var < 0.0) var = 0.0;
    s_mean[tid] = (float)m;
    s_rstd[tid] = 1.0f / sqrtf((float)var + 1e-5f);
  }
  __syncthreads();

  float acc[COUT];
#pragma unroll
  for (int o = 0; o < COUT; ++o) acc[o] = 0.f;

#pragma unroll
  for (int ky = 0; ky < 3; ++ky) {
    const int iy = oy * STRIDE + ky - 1;
    if (iy < 0 || iy >= p.Hin) continue;
#pragma unroll
    for (int kx = 0; kx < 3; ++kx) {
      const int ix = ox * STRIDE + kx - 1;
      if (ix < 0 || ix >= p.Hin) continue;
      const float* wt = p.w + (ky * 3 + kx) * CIN * COUT;
      if (CIN == 1) {
        const float* src = (n < p.B) ? p.io->land1 + (size_t)n * p.Hin * p.Hin : p.io->land2 + (size_t)(n - p.B) * p.Hin * p.Hin;
        const float v = __ldg(src + iy * p.Hin + ix);
#pragma unroll
        for (int o = 0; o < COUT; ++o) acc[o] = fmaf(v, wt[o], acc[o]);
      } else {
        const float4* src = reinterpret_cast<const float4*>(p.in0 + ((size_t)(n * p.Hin + iy) * p.Hin + ix) * CIN);
#pragma unroll
        for (int c4 = 0; c4 < CIN / 4; ++c4) {
          const float4 r = __ldg(src + c4);
          float v[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
          for (int e = 0; e < 4; ++e) {
            const int c = c4 * 4 + e;
            const float a = fmaxf((v[e] - s_mean[c]) * s_rstd[c], 0.f);  // IN + ReLU of the previous layer
#pragma unroll
            for (int o = 0; o < COUT; ++o) acc[o] = fmaf(a, wt[c * COUT + o], acc[o]);
          }
        }
      }
    }
  }
  float4* dst = reinterpret_cast<float4*>(p.out + (size_t)gpix * COUT);
#pragma unroll
  for (int o4 = 0; o4 < COUT / 4; ++o4) dst[o4] = make_float4(acc[o4 * 4], acc[o4 * 4 + 1], acc[o4 * 4 + 2], acc[o4 * 4 + 3]);

  // per-channel sum / sum of squares of the CTA's 256 pixels in a fixed order (warp shuffle tree, then the 8 warps in
  // warp order), rounded to the fixed-point grid once and added to the totals with an order-independent integer atomic
#pragma unroll
  for (int o = 0; o < COUT; ++o) {
    float s = acc[o], q = acc[o] * acc[o];
#pragma unroll
    for (int d = 16; d >= 1; d >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, d);
      q += __shfl_xor_sync(0xffffffffu, q, d);
    }
    if ((tid & 31) == 0) {
      s_red[tid >> 5][o] = s;
      s_red[tid >> 5][COUT + o] = q;
    }
  }
  __syncthreads();
  if (tid < 2 * COUT) {
    const int which = tid / COUT, c = tid - which * COUT;
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += s_red[w][tid];
    stat_add(p.out_stats + ((size_t)n * COUT + c) * 2 + which, stat_fix(t));
  }
}

// land1 [B1,1,256,256], land2 [B2,1,256,256] -> raw [B1+B2,64,64,16] + stats; r0/r1 are workspace raws (with stats).
// w0/w1/w2 are HOST arrays in [slab][Cin][Cout] order.
int launch_landmark_branch(const IoPtrs* io, const float* w0, const float* w1, const float* w2,
                           const Raw& r0, const Raw& r1, const Raw& r2, int B1, int B2, cudaStream_t st) {
  const int NB = B1 + B2;
  AP_REQUIRE(r0.B == NB && r1.B == NB && r2.B == NB, AP_ERR_INVALID, "landmark: workspace batch");
  if (carveout_enabled()) {  // co-residency with the stem kernel's all-shared-memory SM configuration (see elementwise.cu)
    cudaFuncSetAttribute(land_conv_kernel<1, 8, 1>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(land_conv_kernel<8, 16, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    cudaFuncSetAttribute(land_conv_kernel<16, 16, 2>, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
  }
  {
    LandP<1, 8> a{io, nullptr, nullptr, r0.p, r0.stats, B1, 256, 256, {}};
    memcpy(a.w, w0, sizeof(a.w));
    land_conv_kernel<1, 8, 1><<<NB * 65536 / 256, 256, 0, st>>>(a);
    AP_CUDA(cudaGetLastError());
  }
  {
    LandP<8, 16> b{nullptr, r0.p, r0.stats, r1.p, r1.stats, B1, 256, 128, {}};
    memcpy(b.w, w1, sizeof(b.w));
    land_conv_kernel<8, 16, 2><<<NB * 16384 / 256, 256, 0, st>>>(b);
    AP_CUDA(cudaGetLastError());
  }
  {
    LandP<16, 16> c{nullptr, r1.p, r1.stats, r2.p, r2.stats, B1, 128, 64, {}};
    memcpy(c.w, w2, sizeof(c.w));
    land_conv_kernel<16, 16, 2><<<NB * 4096 / 256, 256, 0, st>>>(c);
    AP_CUDA(cudaGetLastError());
  }
  launches_add(3);
  return AP_OK;
}

}  // namespace ap
