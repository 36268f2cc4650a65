// Destination writer of the InstanceNorm apply pass and the double warp (elementwise.cu).
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

}  // namespace ap
