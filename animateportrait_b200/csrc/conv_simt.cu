// CUDA-core (fp32 FMA) convolution kernels.
//
//  * conv_simt_kernel: generic tap-list implicit GEMM (64 pixels x 64 channels per CTA, 4x4 per
//    thread).  It is the arithmetic reference for the tcgen05 path (AP_PREC_FP32_SIMT) and the
//    production kernel for the layers that are too thin for tensor cores: the 7x7 stems on the
//    3-channel photo (networks.py:1218-1243, three stems fused into one Cout=160 problem) and the
//    landmark branch (networks.py:1280-1282).
//  * out_conv_kernel: RefPad3 + Conv7x7 64->output_nc + bias + tanh (networks.py:1277-1279) with the
//    preceding InstanceNorm+ReLU applied while the input tile is staged in shared memory.
//
// Every conv that feeds an affine-less InstanceNorm drops its bias (it cancels exactly; SURVEY.md §8
// a14) and emits per-(n,c) sum / sum-of-squares (fixed point, common.cuh) via one atomic per channel per CTA.
#include "common.cuh"

namespace ap {

__device__ __forceinline__ int reflect_idx(int i, int n) {
  if (i < 0) i = -i;
  if (i >= n) i = 2 * (n - 1) - i;
  return i;
}

template <bool VEC>
__global__ void __launch_bounds__(256) conv_simt_kernel(const SimtConvP p) {
  __shared__ __align__(16) float As[16][68];
  __shared__ __align__(16) float Bs[16][64];
  __shared__ float red[2][16][64];

  const ConvGeom& g = p.g;
  const int tid = threadIdx.x;
  const int HW = g.Hv * g.Wv;
  const int m0 = blockIdx.x * 64;
  const int n0 = blockIdx.y * 64;
  const int img = m0 / HW;
  const int tx = tid & 15, ty = tid >> 4;
  const int Ktot = g.taps.n * g.Cin;

  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;

  const int lm = VEC ? (tid >> 2) : (tid & 63);
  const int pix = m0 - img * HW + lm;
  const int vy = pix / g.Wv, vx = pix - vy * g.Wv;

  for (int k0 = 0; k0 < Ktot; k0 += 16) {
    // ---- A tile: 64 pixels x 16 k ----
    if (VEC) {
      const int t = k0 / g.Cin;
      const int c = k0 - t * g.Cin + (tid & 3) * 4;
      int iy = vy * g.stride + g.taps.dy[t];
      int ix = vx * g.stride + g.taps.dx[t];
      bool ok = true;
      if (g.reflect) {
        iy = reflect_idx(iy, g.Hin);
        ix = reflect_idx(ix, g.Win);
      } else {
        ok = (iy >= 0) && (iy < g.Hin) && (ix >= 0) && (ix < g.Win);
      }
      float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
      if (ok)
        v = *reinterpret_cast<const float4*>(p.in + ((size_t)(img * g.Hin + iy) * g.Win + ix) * p.in_C +
                                             p.in_coff + c);
      const int kb = (tid & 3) * 4;
      As[kb + 0][lm] = v.x;
      As[kb + 1][lm] = v.y;
      As[kb + 2][lm] = v.z;
      As[kb + 3][lm] = v.w;
    } else {
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        const int kk = (tid >> 6) + 4 * i;
        const int k = k0 + kk;
        float v = 0.f;
        if (k < Ktot) {
          const int t = k / g.Cin;
          const int c = k - t * g.Cin;
          int iy = vy * g.stride + g.taps.dy[t];
          int ix = vx * g.stride + g.taps.dx[t];
          bool ok = true;
          if (g.reflect) {
            iy = reflect_idx(iy, g.Hin);
            ix = reflect_idx(ix, g.Win);
          } else {
            ok = (iy >= 0) && (iy < g.Hin) && (ix >= 0) && (ix < g.Win);
          }
          if (ok) {
            if (p.in_nchw)
              v = p.in[((size_t)(img * p.in_C + p.in_coff + c) * g.Hin + iy) * g.Win + ix];
            else
              v = p.in[((size_t)(img * g.Hin + iy) * g.Win + ix) * p.in_C + p.in_coff + c];
          }
        }
        As[kk][lm] = v;
      }
    }
    // ---- W tile: 16 k x 64 couts ----
    {
      const int kk = tid >> 4;
      const int k = k0 + kk;
      const int col = (tid & 15) * 4;
      float4 w = make_float4(0.f, 0.f, 0.f, 0.f);
      if (k < Ktot && n0 + col < g.Cout) {
        const int t = k / g.Cin;
        const int c = k - t * g.Cin;
        w = *reinterpret_cast<const float4*>(p.wpk + ((size_t)g.taps.slab[t] * g.Cin + c) * g.Cout + n0 + col);
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

  // ---- epilogue: scatter raw output, per-channel statistics ----
  const bool col_ok = (n0 + tx * 4) < g.Cout;
  float s[4] = {0.f, 0.f, 0.f, 0.f}, q[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int pm = m0 - img * HW + ty * 4 + i;
    const int y = pm / g.Wv, x = pm - y * g.Wv;
    const int oy = y * g.os + g.py, ox = x * g.os + g.px;
    if (col_ok) {
      float4 v = make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]);
      *reinterpret_cast<float4*>(p.out + ((size_t)(img * g.Hout + oy) * g.Wout + ox) * p.out_C + p.out_coff +
                                 n0 + tx * 4) = v;
    }
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      s[j] += acc[i][j];
      q[j] += acc[i][j] * acc[i][j];
    }
  }
  if (p.stats != nullptr) {
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      red[0][ty][tx * 4 + j] = s[j];
      red[1][ty][tx * 4 + j] = q[j];
    }
    __syncthreads();
    if (tid < 128) {
      const int which = tid >> 6, c = tid & 63;
      if (n0 + c < g.Cout) {
        float t = 0.f;
#pragma unroll
        for (int r = 0; r < 16; ++r) t += red[which][r][c];
        stat_add(p.stats + ((size_t)(img * p.stat_C + p.stat_coff + n0 + c)) * 2 + which, stat_fix(t));
      }
    }
  }
}

static int64_t g_launches = 0;
int64_t launches_get() { return g_launches; }
void launches_add(int n) { g_launches += n; }

int launch_conv_simt(const SimtConvP& p, cudaStream_t st) {
  const ConvGeom& g = p.g;
  AP_REQUIRE((g.Hv * g.Wv) % 64 == 0, AP_ERR_INVALID, "conv_simt: Hv*Wv=%d not a multiple of 64", g.Hv * g.Wv);
  AP_REQUIRE(g.Cout % 4 == 0 && p.out_C % 4 == 0 && p.out_coff % 4 == 0, AP_ERR_INVALID,
             "conv_simt: Cout/out_C/out_coff must be multiples of 4");
  dim3 grid((unsigned)((size_t)g.B * g.Hv * g.Wv / 64), (unsigned)((g.Cout + 63) / 64));
  const bool vec = !p.in_nchw && (g.Cin % 16 == 0) && (p.in_C % 4 == 0) && (p.in_coff % 4 == 0);
  if (vec)
    conv_simt_kernel<true><<<grid, 256, 0, st>>>(p);
  else
    conv_simt_kernel<false><<<grid, 256, 0, st>>>(p);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

// --------------------------------------------------------------------------------------------
// Output stage: IN+ReLU (on the fly) -> RefPad3 -> Conv7x7 64->onc -> +bias -> tanh -> NCHW
// (networks.py:1277-1279).  0.41 GFLOP per frame against 17 MB of input: HBM-bound, N = onc is far too
// thin for tensor cores.  CTA = 32x32 output pixels, 128 threads; each thread owns one column of
// 8 output rows so that a 14-row input column and the 7 ky weights are loaded once per (kx, channel
// quad) and reused for 8 x 7 x 4 FMAs.  The 38x38 input patch is staged 8 channels at a time with the
// preceding InstanceNorm + ReLU applied (4 channels at a time measured slower: half-sector global reads);
// pixel stride 12 floats keeps the float4 reads conflict-free.
// --------------------------------------------------------------------------------------------
constexpr int OC_T = 32, OC_P = OC_T + 6, OC_CG = 8, OC_CS = 12, OC_ROWS = 8;

template <int ONC>
__global__ void __launch_bounds__(128) out_conv_kernel(const OutConvP p) {
  extern __shared__ __align__(16) float sm[];
  float* tile = sm;                           // [38*38][OC_CS]
  float* wsm = tile + OC_P * OC_P * OC_CS;    // [ONC][49][64]
  float* mean = wsm + ONC * 49 * 64;          // [64]
  float* rstd = mean + 64;                    // [64]
  const int tid = threadIdx.x;
  const int n = blockIdx.z;
  const int y0 = blockIdx.y * OC_T, x0 = blockIdx.x * OC_T;
  const int S = 256;
  for (int i = tid; i < ONC * 49 * 64; i += 128) wsm[i] = p.w[i];
  if (tid < 64) {
    const double inv = 1.0 / (double)(S * S);
    const double su = (double)p.stats[((size_t)n * 64 + tid) * 2 + 0] * STAT_INV_SCALE;
    const double sq = (double)p.stats[((size_t)n * 64 + tid) * 2 + 1] * STAT_INV_SCALE;
    const double m = su * inv;
    double var = sq * inv - m * m;
    if (var < 0.0) var = 0.0;
    mean[tid] = (float)m;
    rstd[tid] = (float)(1.0 / sqrt(var + 1e-5));
  }
  const int tx = tid & 31, tg = tid >> 5;
  float acc[ONC][OC_ROWS];
#pragma unroll
  for (int o = 0; o < ONC; ++o)
#pragma unroll
    for (int r = 0; r < OC_ROWS; ++r) acc[o][r] = 0.f;

  for (int c0 = 0; c0 < 64; c0 += OC_CG) {
    __syncthreads();  // previous group fully consumed (and, first time, mean/rstd/weights visible)
    for (int i = tid; i < OC_P * OC_P * (OC_CG / 4); i += 128) {
      const int pp = i / (OC_CG / 4), cq = (i % (OC_CG / 4)) * 4;
      const int py = pp / OC_P, px = pp - py * OC_P;
      const int iy = reflect_idx(y0 + py - 3, S), ix = reflect_idx(x0 + px - 3, S);
      float4 v = __ldg(reinterpret_cast<const float4*>(p.raw + ((size_t)(n * S + iy) * S + ix) * 64 + c0 + cq));
      v.x = fmaxf((v.x - mean[c0 + cq + 0]) * rstd[c0 + cq + 0], 0.f);
      v.y = fmaxf((v.y - mean[c0 + cq + 1]) * rstd[c0 + cq + 1], 0.f);
      v.z = fmaxf((v.z - mean[c0 + cq + 2]) * rstd[c0 + cq + 2], 0.f);
      v.w = fmaxf((v.w - mean[c0 + cq + 3]) * rstd[c0 + cq + 3], 0.f);
      *reinterpret_cast<float4*>(tile + pp * OC_CS + cq) = v;
    }
    __syncthreads();
#pragma unroll 1
    for (int kx = 0; kx < 7; ++kx) {
#pragma unroll
      for (int cq = 0; cq < OC_CG; cq += 4) {
        float4 a[OC_ROWS + 6];
        const float* ip = tile + ((tg * OC_ROWS) * OC_P + tx + kx) * OC_CS + cq;
#pragma unroll
        for (int r = 0; r < OC_ROWS + 6; ++r) a[r] = *reinterpret_cast<const float4*>(ip + r * OC_P * OC_CS);
#pragma unroll
        for (int o = 0; o < ONC; ++o) {
#pragma unroll
          for (int ky = 0; ky < 7; ++ky) {
            const float4 w = *reinterpret_cast<const float4*>(wsm + (o * 49 + ky * 7 + kx) * 64 + c0 + cq);
#pragma unroll
            for (int r = 0; r < OC_ROWS; ++r) {
              acc[o][r] = fmaf(a[r + ky].x, w.x, acc[o][r]);
              acc[o][r] = fmaf(a[r + ky].y, w.y, acc[o][r]);
              acc[o][r] = fmaf(a[r + ky].z, w.z, acc[o][r]);
              acc[o][r] = fmaf(a[r + ky].w, w.w, acc[o][r]);
            }
          }
        }
      }
    }
  }
  float* outp = p.io->out;
#pragma unroll
  for (int o = 0; o < ONC; ++o)
#pragma unroll
    for (int r = 0; r < OC_ROWS; ++r)
      outp[((size_t)(n * ONC + o) * S + y0 + tg * OC_ROWS + r) * S + x0 + tx] = tanhf(acc[o][r] + p.bias[o]);
}

int launch_out_conv(const OutConvP& p, cudaStream_t st) {
  dim3 grid(256 / OC_T, 256 / OC_T, p.B);
  const size_t smem = (size_t)(OC_P * OC_P * OC_CS + p.onc * 49 * 64 + 128) * sizeof(float);
  if (p.onc == 1) {
    AP_CUDA(cudaFuncSetAttribute(out_conv_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    out_conv_kernel<1><<<grid, 128, smem, st>>>(p);
  } else if (p.onc == 3) {
    AP_CUDA(cudaFuncSetAttribute(out_conv_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    out_conv_kernel<3><<<grid, 128, smem, st>>>(p);
  } else {
    set_error("out_conv: output_nc=%d unsupported", p.onc);
    return AP_ERR_UNSUPPORTED;
  }
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

// taps of a k x k conv with padding `pad` (coordinates in the un-haloed input); extra_origin shifts
// them when the consumer addresses a haloed view.
ConvTaps make_taps_conv(int k, int pad, int extra_origin) {
  ConvTaps t{};
  t.n = k * k;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) {
      const int i = ky * k + kx;
      t.dy[i] = (int8_t)(ky - pad + extra_origin);
      t.dx[i] = (int8_t)(kx - pad + extra_origin);
      t.slab[i] = (uint8_t)i;
    }
  return t;
}

// ConvTranspose2d(k=3, s=2, p=1, op=1): out[2i-1+ky, 2j-1+kx] += in[i,j] * W[ky,kx]  (SURVEY.md A.4).
// Output phase (py,px) gathers: py=0 -> ky=1 at di=0;  py=1 -> ky=0 at di=+1 and ky=2 at di=0.
ConvTaps make_taps_convT_phase(int py, int px) {
  ConvTaps t{};
  int kys[2], dys[2], nky, kxs[2], dxs[2], nkx;
  if (py == 0) { nky = 1; kys[0] = 1; dys[0] = 0; } else { nky = 2; kys[0] = 0; dys[0] = 1; kys[1] = 2; dys[1] = 0; }
  if (px == 0) { nkx = 1; kxs[0] = 1; dxs[0] = 0; } else { nkx = 2; kxs[0] = 0; dxs[0] = 1; kxs[1] = 2; dxs[1] = 0; }
  t.n = 0;
  for (int a = 0; a < nky; ++a)
    for (int b = 0; b < nkx; ++b) {
      t.dy[t.n] = (int8_t)dys[a];
      t.dx[t.n] = (int8_t)dxs[b];
      t.slab[t.n] = (uint8_t)(kys[a] * 3 + kxs[b]);
      ++t.n;
    }
  return t;
}

}  // namespace ap
