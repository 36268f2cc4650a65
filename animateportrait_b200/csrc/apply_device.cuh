// Device-side pieces of the InstanceNorm apply pass shared by elementwise.cu (stand-alone kernels) and conv_umma.cu
// (apply warps fused into the trunk conv kernel).
#pragma once
#include "common.cuh"

namespace ap {

// ------------------------------------------------------------------------------------------------
// destination writer shared by apply / warp: 4 consecutive channels of one pixel, any ActFmt,
// optional reflected halo (pad == 1).
// ------------------------------------------------------------------------------------------------
struct Dst {
  int fmt; void* d0; void* d1; int C, coff, pad, H, W, halo_reflect;
};

__device__ __forceinline__ void store4(const Dst& d, size_t pix_index, int c, float4 v) {
  const size_t off = pix_index * d.C + d.coff + c;
  if (d.fmt == FMT_F32) {
    *reinterpret_cast<float4*>(reinterpret_cast<float*>(d.d0) + off) = v;
  } else {
    const __nv_bfloat16 h0 = __float2bfloat16_rn(v.x), h1 = __float2bfloat16_rn(v.y),
                        h2 = __float2bfloat16_rn(v.z), h3 = __float2bfloat16_rn(v.w);
    __nv_bfloat162 a, b;
    a.x = h0; a.y = h1; b.x = h2; b.y = h3;
    uint2 pk;
    pk.x = *reinterpret_cast<uint32_t*>(&a);
    pk.y = *reinterpret_cast<uint32_t*>(&b);
    *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.d0) + off) = pk;
    if (d.fmt == FMT_BF16X2) {
      a.x = __float2bfloat16_rn(v.x - __bfloat162float(h0));
      a.y = __float2bfloat16_rn(v.y - __bfloat162float(h1));
      b.x = __float2bfloat16_rn(v.z - __bfloat162float(h2));
      b.y = __float2bfloat16_rn(v.w - __bfloat162float(h3));
      pk.x = *reinterpret_cast<uint32_t*>(&a);
      pk.y = *reinterpret_cast<uint32_t*>(&b);
      *reinterpret_cast<uint2*>(reinterpret_cast<__nv_bfloat16*>(d.d1) + off) = pk;
    }
  }
}

// writes pixel (n,y,x) and, with halo_reflect, its mirror images in the halo ring
__device__ __forceinline__ void store_pixel(const Dst& d, int n, int y, int x, int c, float4 v) {
  const int Hp = d.H + 2 * d.pad, Wp = d.W + 2 * d.pad;
  const size_t base = (size_t)n * Hp;
  store4(d, (base + y + d.pad) * Wp + x + d.pad, c, v);
  if (d.halo_reflect && d.pad == 1) {
    const int ry = (y == 1) ? 0 : ((y == d.H - 2) ? d.H + 1 : -1);
    const int rx = (x == 1) ? 0 : ((x == d.W - 2) ? d.W + 1 : -1);
    if (ry >= 0) store4(d, (base + ry) * Wp + x + 1, c, v);
    if (rx >= 0) store4(d, (base + y + 1) * Wp + rx, c, v);
    if (ry >= 0 && rx >= 0) store4(d, (base + ry) * Wp + rx, c, v);
  }
}

__device__ __forceinline__ float4 ld_coherent(const float* p) {
  float4 r;
  asm volatile("ld.global.L1::no_allocate.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w) : "l"(p) : "memory");
  return r;
}

constexpr int AF_THREADS = 256, AF_TPP = 64, AF_NY = AF_THREADS / AF_TPP, AF_PIX = 8 * AF_NY;

// `tid` in [0, AF_THREADS): the 256 threads that run this role (a whole CTA of apply_flags_kernel, or the apply warps
// of a conv CTA); items first_item, first_item + item_stride, ... ; named barrier 2 synchronises exactly these threads.
// UNROLL = pixel passes whose loads are in flight together (4 or 8; 8 for the fused role, whose 8 warps per SM must
// sustain the whole HBM stream).
template <int MODE, int UNROLL = 4>
__device__ __forceinline__ void apply_flag_items(const ApplyP& p, int tid, int first_item, int item_stride) {
  const int cq = tid % AF_TPP, py = tid / AF_TPP;
  const int c = cq * 4;
  const int HW = p.H * p.W;
  const int logW = 31 - __clz(p.W);
  const int items_per_img = HW / AF_PIX;
  const int total = items_per_img * p.B;
  Dst d{p.fmt, p.d0, p.d1, p.dC, p.dcoff, p.dpad, p.H, p.W, p.halo_reflect};
  float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f}, mean2[4] = {0.f, 0.f, 0.f, 0.f}, rstd2[4] = {0.f, 0.f, 0.f, 0.f};
  int cur = -1;
  for (int item = first_item; item < total; item += item_stride) {
    const int n = item / items_per_img;
    const int pix0 = (item - n * items_per_img) * AF_PIX + py;
    if (n != cur) {
      if (tid == 0) {
        if (p.wait0.flags) flag_wait(p.wait0.flags + n, p.wait0.expected);
        if (p.wait1.flags) flag_wait(p.wait1.flags + n, p.wait1.expected);
      }
      asm volatile("bar.sync 2, 256;" ::: "memory");
      const double inv_n = 1.0 / (double)HW;
#pragma unroll
      for (int e = 0; e < 4; ++e) {
        if (p.stats) stats_to_affine(p.stats, n, p.stat_C, p.stat_coff, c + e, inv_n, &mean[e], &rstd[e]);
        else { mean[e] = p.bias ? -p.bias[c + e] : 0.f; rstd[e] = 1.f; }
        if (MODE == 1) stats_to_affine(p.stats2, n, p.stat2_C, p.stat2_coff, c + e, inv_n, &mean2[e], &rstd2[e]);
      }
      cur = n;
    }
    const float* raw = p.raw + ((size_t)n * HW) * p.raw_C + p.raw_coff + c;
    const float* raw2 = (MODE == 1) ? p.raw2 + ((size_t)n * HW) * p.raw2_C + p.raw2_coff + c : nullptr;
    const float* rin = (MODE == 2) ? p.res_in + ((size_t)n * HW) * p.C + c : nullptr;
    float* rout = p.res_out ? p.res_out + ((size_t)n * HW) * p.C + c : nullptr;
#pragma unroll 1
    for (int k0 = 0; k0 < AF_PIX; k0 += UNROLL * AF_NY) {
      float4 v[UNROLL], u[UNROLL];
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const int pix = pix0 + k0 + j * AF_NY;
        v[j] = ld_coherent(raw + (size_t)pix * p.raw_C);
        if (MODE == 1) u[j] = ld_coherent(raw2 + (size_t)pix * p.raw2_C);
        if (MODE == 2) u[j] = ld_coherent(rin + (size_t)pix * p.C);
      }
#pragma unroll
      for (int j = 0; j < UNROLL; ++j) {
        const int pix = pix0 + k0 + j * AF_NY;
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
        if (MODE == 2) { o.x += u[j].x; o.y += u[j].y; o.z += u[j].z; o.w += u[j].w; }
        if (rout) *reinterpret_cast<float4*>(rout + (size_t)pix * p.C) = o;
        if (p.fmt >= 0) store_pixel(d, n, pix >> logW, pix & (p.W - 1), c, o);
      }
    }
    if (p.done_flags) {
      __threadfence();
      asm volatile("bar.sync 2, 256;" ::: "memory");
      if (tid == 0) flag_add(p.done_flags + n, 1u);
    }
  }
}

}  // namespace ap
