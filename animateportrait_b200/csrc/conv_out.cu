// Output stage on tensor cores:  IN + ReLU (of model3.3's raw output, applied on the fly) -> ReflectionPad2d(3)
// -> Conv2d(64 -> output_nc, 7x7) -> + bias -> tanh -> NCHW   (Module2/models/networks.py:1277-1279).
//
// N = output_nc is far too thin for an implicit GEMM over pixels x output channels (the CUDA-core kernel in
// conv_simt.cu needs 3136 FMAs per output pixel and is bound by shared-memory bandwidth).  The contraction over the 64
// input channels is factored out instead:
//     t[p][j] = sum_c a[p][c] * W[c][j]         j = ky*7 + kx  (49 taps, padded to N = 64), p = input pixel
//     out[y][x] = bias + sum_{ky,kx} t[(y+ky-3, x+kx-3)][ky*7+kx]
// The first line is a [pixels x 64] x [64 x 64] GEMM on tcgen05 (A built in shared memory by "builder" warps that
// normalise, ReLU and split the raw fp32 activations into bf16 hi/lo; three products, fp32 accumulation in TMEM);
// the second is a 49-term shifted sum per output pixel out of shared memory.  Per output pixel that is ~100
// shared-memory instructions instead of 3136 FMAs, and the kernel becomes bound by reading the 64-channel input.
//
// Work item = (16x16 output tile, output channel o).  Its 22x22 patch of input positions (reflection applied when
// the patch is gathered) is 484 GEMM rows = 4 M-tiles of 128.  Warp roles (672 threads):
//   warp 0        TMEM allocator, weight loader (bulk copy of the pre-swizzled images), MMA issuer (one lane)
//   warps 1-16    builders: group g = 0/1 (8 warps each) owns A slot g and builds M-tiles g and g+2 of every item
//   warps 17-20   epilogue: tcgen05.ld -> t[484][49] in shared memory -> shifted sums -> bias, tanh, store
// Two accumulator sets (2 x 256 TMEM columns) let the builders/MMAs of item i+1 run under the epilogue of item i.
#include "common.cuh"
#include "umma.cuh"

namespace ap {

void launches_add(int n);

constexpr int OT = 16;                   // output tile edge
constexpr int OP = OT + 6;               // patch edge
constexpr int OPOS = OP * OP;            // 484 patch positions
constexpr int O_WPLANE = 64 * 128;       // [64 rows j][64 k] bf16, 128B-swizzled
constexpr int O_WIMG = 2 * O_WPLANE;     // hi + lo planes of one output channel
constexpr int O_APLANE = 128 * 128;      // [128 rows][64 k] bf16
constexpr int O_ASLOT = 2 * O_APLANE;    // hi + lo
constexpr int O_TSTRIDE = 49;            // odd: conflict-free for lane-strided rows
constexpr int O_TBYTES = ((OPOS * O_TSTRIDE * 4 + 127) / 128) * 128;
constexpr int O_BUILD_WARPS = 8;           // builder warps per group (2 groups)
constexpr int O_THREADS = 32 + 2 * 32 * O_BUILD_WARPS + 128;
constexpr int O_MAX_ONC = 3;
constexpr size_t O_SMEM = 1024 + O_MAX_ONC * O_WIMG + 2 * O_ASLOT + O_TBYTES + 512 + 256;

struct OutUmmaP {
  const float* raw;       // [B,256,256,64] raw output of model3.3 (pre-InstanceNorm)
  const stat_t* stats;    // [B][64][2] fixed point
  const uint8_t* wimg;    // onc x O_WIMG pre-swizzled weight images
  const float* bias;      // [onc]
  const IoPtrs* io;       // io->out: [B,onc,256,256] of the caller (may be peer memory: plain stores)
  int B, onc, items;      // items = B * 256 tiles * onc
};

__device__ __forceinline__ int o_reflect(int i) {
  if (i < 0) i = -i;
  if (i > 255) i = 510 - i;
  return i;
}

__global__ void __launch_bounds__(O_THREADS, 1) out_umma_kernel(const __grid_constant__ OutUmmaP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sW = smem_base;
  const uint32_t sA = smem_base + O_MAX_ONC * O_WIMG;
  uint8_t* gA = sgen + O_MAX_ONC * O_WIMG;
  float* tbuf = reinterpret_cast<float*>(sgen + O_MAX_ONC * O_WIMG + 2 * O_ASLOT);
  const uint32_t bars = smem_base + O_MAX_ONC * O_WIMG + 2 * O_ASLOT + O_TBYTES + 512;
  // a_full[g] +0,+8   a_empty[g] +16,+24   tfull[a] +32,+40   tempty[a] +48,+56   wbar +64   tmem ptr +80
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sgen + O_MAX_ONC * O_WIMG + 2 * O_ASLOT + O_TBYTES + 512 + 80);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int G = gridDim.x;
  const int i_begin = (int)(((long long)blockIdx.x * p.items) / G);
  const int i_end = (int)(((long long)(blockIdx.x + 1) * p.items) / G);
  const int nitems = i_end - i_begin;

  if (warp == 0) {
    if (lane == 0) {
      for (int g = 0; g < 2; ++g) {
        mbar_init(bars + 8 * g, 32 * O_BUILD_WARPS);  // a_full: every builder thread of the group arrives
        mbar_init(bars + 16 + 8 * g, 1);    // a_empty: one tcgen05.commit
        mbar_init(bars + 32 + 8 * g, 1);    // tfull: one tcgen05.commit
        mbar_init(bars + 48 + 8 * g, 4);    // tempty: one arrive per epilogue warp
      }
      mbar_init(bars + 64, 1);
      asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncwarp();
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"(512u)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== weight loader + MMA issuer =====================
    if (lane == 0 && nitems > 0) {
      mbar_expect_tx(bars + 64, (uint32_t)(p.onc * O_WIMG));
      for (int o = 0; o < p.onc; ++o) bulk_load(sW + o * O_WIMG, p.wimg + (size_t)o * O_WIMG, O_WIMG, bars + 64);
      mbar_wait(bars + 64, 0);
      // instruction descriptor: c=f32, a=b=bf16, K-major, N=64, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(64 >> 3) << 17) | ((128u >> 4) << 24);
      for (int it = 0; it < nitems; ++it) {
        const int o = (i_begin + it) % p.onc;
        const uint32_t acc = (uint32_t)it & 1u;
        mbar_wait(bars + 48 + 8 * acc, (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint64_t w_hi = make_sw128_desc(sW + o * O_WIMG);
        const uint64_t w_lo = make_sw128_desc(sW + o * O_WIMG + O_WPLANE);
#pragma unroll
        for (int m = 0; m < 4; ++m) {
          const int g = m & 1;
          const uint32_t use = (uint32_t)it * 2u + (uint32_t)(m >> 1);  // how often slot g has been filled before
          mbar_wait(bars + 8 * g, use & 1u);
          tc_fence_after();
          const uint32_t d = tmem_base + acc * 256u + (uint32_t)m * 64u;
          const uint64_t a_hi = make_sw128_desc(sA + g * O_ASLOT);
          const uint64_t a_lo = make_sw128_desc(sA + g * O_ASLOT + O_APLANE);
#pragma unroll
          for (int k = 0; k < 4; ++k) {
            const uint64_t ko = (uint64_t)(k * 2);
            umma_bf16(d, a_hi + ko, w_hi + ko, idesc, k > 0 ? 1u : 0u);
            umma_bf16(d, a_hi + ko, w_lo + ko, idesc, 1);
            umma_bf16(d, a_lo + ko, w_hi + ko, idesc, 1);
          }
          umma_commit(bars + 16 + 8 * g);
        }
        umma_commit(bars + 32 + 8 * acc);
      }
    }
    __syncwarp();
  } else if (warp <= 2 * O_BUILD_WARPS) {
    // ===================== A builders: group g (4 warps) owns slot g =====================
    // A warp builds 32 GEMM rows (= patch positions) of the M-tile.  One load instruction covers two whole rows
    // (lane>>4 picks the row, lane&15 the channel quad): fully coalesced 256-byte reads, and every thread keeps the
    // SAME 4 channels for all its rows, so their mean / rstd live in registers.
    constexpr int RPW = 128 / O_BUILD_WARPS;  // GEMM rows per builder warp
    const int bt = threadIdx.x - 32;
    const int g = bt / (32 * O_BUILD_WARPS), bw = (bt >> 5) % O_BUILD_WARPS;
    const int q = lane & 15, half = lane >> 4;
    int cur_img = -1;
    float mean[4] = {0.f, 0.f, 0.f, 0.f}, rstd[4] = {1.f, 1.f, 1.f, 1.f};
    for (int it = 0; it < nitems; ++it) {
      const int item = i_begin + it;
      const int tile = item / p.onc;
      const int img = tile >> 8, ty = (tile >> 4) & 15, tx = tile & 15;
      if (img != cur_img) {
#pragma unroll
        for (int e = 0; e < 4; ++e) stats_to_affine(p.stats, img, 64, 0, 4 * q + e, 1.0 / 65536.0, &mean[e], &rstd[e]);
        cur_img = img;
      }
      const float* src_img = p.raw + (size_t)img * 65536 * 64 + 4 * q;
#pragma unroll 1
      for (int mm = 0; mm < 2; ++mm) {
        const int row0 = bw * RPW + half;                  // + 2*i: row inside the M-tile
        const int pos0 = (g + 2 * mm) * 128 + row0;
        const uint32_t use = (uint32_t)it * 2u + (uint32_t)mm;
        float4 v[RPW / 2];
#pragma unroll
        for (int i = 0; i < RPW / 2; ++i) {
          const int pos = pos0 + 2 * i;
          v[i] = make_float4(0.f, 0.f, 0.f, 0.f);
          if (pos < OPOS) {
            const int py = pos / OP, px = pos - py * OP;
            const int sy = o_reflect(ty * OT - 3 + py), sx = o_reflect(tx * OT - 3 + px);
            v[i] = __ldg(reinterpret_cast<const float4*>(src_img + ((size_t)sy * 256 + sx) * 64));
          }
        }
        mbar_wait(bars + 16 + 8 * g, (use & 1u) ^ 1u);  // the MMAs that read this slot last have retired
#pragma unroll
        for (int i = 0; i < RPW / 2; ++i) {
          const int row = row0 + 2 * i;
          const float a0 = fmaxf((v[i].x - mean[0]) * rstd[0], 0.f), a1 = fmaxf((v[i].y - mean[1]) * rstd[1], 0.f);
          const float a2 = fmaxf((v[i].z - mean[2]) * rstd[2], 0.f), a3 = fmaxf((v[i].w - mean[3]) * rstd[3], 0.f);
          const __nv_bfloat16 h0 = __float2bfloat16_rn(a0), h1 = __float2bfloat16_rn(a1);
          const __nv_bfloat16 h2 = __float2bfloat16_rn(a2), h3 = __float2bfloat16_rn(a3);
          const __nv_bfloat16 l0 = __float2bfloat16_rn(a0 - __bfloat162float(h0));
          const __nv_bfloat16 l1 = __float2bfloat16_rn(a1 - __bfloat162float(h1));
          const __nv_bfloat16 l2 = __float2bfloat16_rn(a2 - __bfloat162float(h2));
          const __nv_bfloat16 l3 = __float2bfloat16_rn(a3 - __bfloat162float(h3));
          uint2 hi, lo;
          hi.x = (uint32_t)__bfloat16_as_ushort(h0) | ((uint32_t)__bfloat16_as_ushort(h1) << 16);
          hi.y = (uint32_t)__bfloat16_as_ushort(h2) | ((uint32_t)__bfloat16_as_ushort(h3) << 16);
          lo.x = (uint32_t)__bfloat16_as_ushort(l0) | ((uint32_t)__bfloat16_as_ushort(l1) << 16);
          lo.y = (uint32_t)__bfloat16_as_ushort(l2) | ((uint32_t)__bfloat16_as_ushort(l3) << 16);
          // 128B-swizzled K-major row: 16-byte chunk (q>>1) lands at chunk (q>>1) ^ (row & 7); 8 bytes per thread
          uint8_t* dst = gA + g * O_ASLOT + (row >> 3) * 1024 + (row & 7) * 128 + ((((q >> 1) ^ (row & 7))) << 4) + (q & 1) * 8;
          *reinterpret_cast<uint2*>(dst) = hi;
          *reinterpret_cast<uint2*>(dst + O_APLANE) = lo;
        }
        fence_proxy_async();
        mbar_arrive(bars + 8 * g);
      }
    }
  } else {
    // ===================== epilogue (warps 17..20) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int et = (q << 5) | lane;  // a stable 0..127 index (any bijection works for the shifted sums)
    float* out_base = p.io->out;
    for (int it = 0; it < nitems; ++it) {
      const int item = i_begin + it;
      const int tile = item / p.onc, o = item - tile * p.onc;
      const int img = tile >> 8, ty = (tile >> 4) & 15, tx = tile & 15;
      const uint32_t acc = (uint32_t)it & 1u;
      mbar_wait(bars + 32 + 8 * acc, ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
      // part 1: accumulators -> t[pos][49]
#pragma unroll 1
      for (int m = 0; m < 4; ++m) {
        const int pos = m * 128 + q * 32 + lane;
        float* trow = tbuf + pos * O_TSTRIDE;
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u + (uint32_t)(m * 64), v);
        if (pos < OPOS) {
#pragma unroll
          for (int j = 0; j < 32; ++j) trow[j] = v[j];
        }
        tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + acc * 256u + (uint32_t)(m * 64 + 32), v);
        if (pos < OPOS) {
#pragma unroll
          for (int j = 0; j < 17; ++j) trow[32 + j] = v[j];
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 48 + 8 * acc);  // accumulator set free for the item after next
      asm volatile("bar.sync 2, 128;" ::: "memory");    // t complete
      // part 2: shifted 49-term sums, 2 output pixels per thread
      const float b = p.bias[o];
      float* dst = out_base + ((size_t)(img * p.onc + o) * 256 + ty * OT) * 256 + tx * OT;
#pragma unroll
      for (int k = 0; k < 2; ++k) {
        const int oy = (et >> 4) + 8 * k, ox = et & 15;
        const float* t0 = tbuf + (oy * OP + ox) * O_TSTRIDE;
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int ky = 0; ky < 7; ++ky) {
          const float* tr = t0 + ky * OP * O_TSTRIDE + ky * 7;
          s0 += tr[0 * O_TSTRIDE + 0];
          s1 += tr[1 * O_TSTRIDE + 1];
          s2 += tr[2 * O_TSTRIDE + 2];
          s3 += tr[3 * O_TSTRIDE + 3];
          s0 += tr[4 * O_TSTRIDE + 4];
          s1 += tr[5 * O_TSTRIDE + 5];
          s2 += tr[6 * O_TSTRIDE + 6];
        }
        dst[oy * 256 + ox] = tanhf((s0 + s1) + (s2 + s3) + b);
      }
      asm volatile("bar.sync 2, 128;" ::: "memory");    // t may be overwritten
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// weight images: src = Conv2d weight of model3.7 [onc][64][7][7] fp32 -> per output channel a pre-swizzled
// [plane][64 rows j = ky*7+kx][64 k = c] bf16 image (rows 49..63 zero)
// ------------------------------------------------------------------------------------------------
__global__ void pack_out_umma_kernel(const float* __restrict__ src, int onc, uint8_t* __restrict__ img) {
  const int total = onc * 64 * 64;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int c = i & 63, j = (i >> 6) & 63, o = i >> 12;
    const float w = (j < 49) ? src[((size_t)o * 64 + c) * 49 + j] : 0.f;
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
    const size_t off = (size_t)o * O_WIMG + (j >> 3) * 1024 + (j & 7) * 128 + (((c >> 3) ^ (j & 7)) << 4) + (c & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(img + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(img + O_WPLANE + off) = l;
  }
}

size_t out_umma_weight_bytes(int onc) { return (size_t)onc * O_WIMG; }

int launch_pack_out_umma(const float* src, int onc, uint8_t* img, cudaStream_t st) {
  pack_out_umma_kernel<<<(onc * 4096 + 255) / 256, 256, 0, st>>>(src, onc, img);
  AP_CUDA(cudaGetLastError());
  return AP_OK;
}

static int g_out_sms = 0;

int out_umma_init_device() {
  int dev = 0;
  AP_CUDA(cudaGetDevice(&dev));
  AP_CUDA(cudaDeviceGetAttribute(&g_out_sms, cudaDevAttrMultiProcessorCount, dev));
  AP_CUDA(cudaFuncSetAttribute(out_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)O_SMEM));
  return AP_OK;
}

int launch_out_umma(const OutConvP& q, const uint8_t* wimg, cudaStream_t st) {
  AP_REQUIRE(q.onc >= 1 && q.onc <= O_MAX_ONC, AP_ERR_UNSUPPORTED, "out conv: output_nc=%d", q.onc);
  AP_TRY(umma_init());
  OutUmmaP p{};
  p.raw = q.raw; p.stats = q.stats; p.wimg = wimg; p.bias = q.bias; p.io = q.io;
  p.B = q.B; p.onc = q.onc; p.items = q.B * 256 * q.onc;
  const int grid = p.items < g_out_sms ? p.items : g_out_sms;
  out_umma_kernel<<<grid, O_THREADS, O_SMEM, st>>>(p);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

}  // namespace ap
