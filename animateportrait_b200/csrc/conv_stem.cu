// tcgen05 kernel for the three 7x7 stems of the generator, fused into one problem.
//
// Reference: model_tri00 / model_tri10 / model_tri20 = ReflectionPad2d(3) + Conv2d(3 -> 32|64|64, 7x7)
// (Module2/models/networks.py:1218-1221, 1229-1232, 1240-1243), all three applied to the same photo.
// They are one GEMM per 128-pixel tile:  D[128 px, 160 cout] = A[128 px, K] * W[160, K]^T with
// K = 7*7*3 = 147 (padded to 160 = 10 k-steps of 16).  The bias is dropped: each stem feeds an
// affine-less InstanceNorm.
//
// Cin = 3 makes TMA im2col pointless, so the A operand is built in shared memory by four "builder"
// warps straight in the 128-byte-swizzled K-major layout tcgen05.mma reads:
//   * the tile's reflected input patch (8 rows x 70 px x 3 ch) is staged once as packed
//     (bf16 hi | bf16 lo << 16) words: hi = bf16(x), lo = bf16(x - hi);
//   * each builder thread owns one pixel row of A and gathers its 147 taps with compile-time offsets,
//     PRMT-packing hi and lo planes, 16 bytes per store.
// The whole weight matrix (hi and lo planes, pre-swizzled by pack_stem_umma_kernel) is loaded once per
// CTA with bulk copies and stays resident; CTAs are persistent over a contiguous range of tiles.
// fp32-accurate mode runs A_hi*W_hi + A_hi*W_lo + A_lo*W_hi into one TMEM accumulator (NPROD = 3).
// Two TMEM accumulators (2 x 256 columns) let the epilogue of tile t overlap the MMAs of tile t+1.
//
// Warp roles (416 threads): warp 0 = TMEM allocator, weight loader, MMA issuer; warps 1-8 = A builders (two groups
// splitting the K range); warps 9-12 = epilogue (tcgen05.ld -> raw fp32 NHWC store + per-channel sum / sum of squares for the
// following InstanceNorm, accumulated in registers -- as fixed-point integers, common.cuh: stat_t -- across the CTA's
// tiles of one image).
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace ap {

void launches_add(int n);

constexpr int ST_N = 160;                        // 32 + 64 + 64 output channels
constexpr int ST_K = 147;                        // 7*7*3
constexpr int ST_WCHUNK = ST_N * 128;            // one [160 x 64] bf16 K-chunk, SW128: 20480 B
constexpr int ST_WBYTES = 2 * 3 * ST_WCHUNK;     // hi + lo planes, 3 chunks: 122880 B
constexpr int ST_APLANE = 128 * 128;             // one [128 x 64] bf16 A chunk: 16384 B
constexpr int ST_ASLOT = 2 * ST_APLANE;          // hi + lo
constexpr int ST_PW = 70, ST_PH = 8;             // patch: 64+6 columns, 2+6 rows
constexpr int ST_PATCH = 3 * ST_PH * ST_PW;      // words
constexpr int ST_THREADS = 32 + 256 + 128;    // MMA warp, 8 builder warps, 4 epilogue warps
constexpr int ST_EPI = 4 * 2 * 4096;             // two 4 KB staging slabs per epilogue warp
constexpr int ST_RING = 2 * ST_ASLOT;            // two A slots: chunk c lives in slot c & 1
constexpr size_t ST_SMEM = 1024 + ST_WBYTES + ST_RING + ST_EPI + ST_PATCH * 4 + 256;

struct StemP {
  CUtensorMap tmO;       // fp32 NHWC output [B,256,256,160], box {32 ch, 32 px, 1 row, 1 image}
  const IoPtrs* io;      // io->input: [B,3,256,256] NCHW fp32 of the caller
  const uint8_t* wimg;   // ST_WBYTES: [plane][chunk][160 rows][64 k] bf16, 128B-swizzled smem image
  float* out;            // raw NHWC [B,256,256,160]
  stat_t* stats;         // [B][160][2] fixed point
  int tiles;             // B * 512 (tile = 2 rows x 64 px)
  int dbg;               // AP_STEM_DBG timing probe: 1 = no output stores, 2 = no statistics, 4 = builders skip the A build
};

__device__ __forceinline__ int st_reflect(int i) {
  if (i < 0) i = -i;
  if (i > 255) i = 510 - i;
  return i;
}

// builds the 16-byte groups [J0, J1) of K-chunk C of one A row
template <int C, int J0, int J1>
__device__ __forceinline__ void stem_build_chunk(const uint32_t* __restrict__ pbase, uint8_t* slot, int r8, int atom_off) {
  // groups of 8 consecutive k = 16 bytes of the hi plane and 16 bytes of the lo plane
#pragma unroll
  for (int j = J0; j < J1; ++j) {
    uint32_t u[8];
#pragma unroll
    for (int e = 0; e < 8; ++e) {
      const int k = C * 64 + j * 8 + e;
      if (k < ST_K) {
        const int tap = k / 3, ch = k - tap * 3;
        const int ky = tap / 7, kx = tap - ky * 7;
        u[e] = pbase[ch * (ST_PH * ST_PW) + ky * ST_PW + kx];
      } else {
        u[e] = 0u;
      }
    }
    uint4 hi, lo;
    hi.x = __byte_perm(u[0], u[1], 0x5410); lo.x = __byte_perm(u[0], u[1], 0x7632);
    hi.y = __byte_perm(u[2], u[3], 0x5410); lo.y = __byte_perm(u[2], u[3], 0x7632);
    hi.z = __byte_perm(u[4], u[5], 0x5410); lo.z = __byte_perm(u[4], u[5], 0x7632);
    hi.w = __byte_perm(u[6], u[7], 0x5410); lo.w = __byte_perm(u[6], u[7], 0x7632);
    const int off = atom_off + ((j ^ r8) << 4);
    *reinterpret_cast<uint4*>(slot + off) = hi;
    *reinterpret_cast<uint4*>(slot + ST_APLANE + off) = lo;
  }
}

template <int NPROD>
__global__ void __launch_bounds__(ST_THREADS, 1) stem_umma_kernel(const __grid_constant__ StemP p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* sgen = smem_raw + (smem_base - smem_u32(smem_raw));
  // layout: W | A ring (2 slots) | epilogue slabs | patch | barriers
  const uint32_t sW = smem_base;
  const uint32_t sA = smem_base + ST_WBYTES;
  uint8_t* gA = sgen + ST_WBYTES;
  const uint32_t epi_s = smem_base + ST_WBYTES + ST_RING;
  uint8_t* epi_gen = sgen + ST_WBYTES + ST_RING;
  uint32_t* patch = reinterpret_cast<uint32_t*>(sgen + ST_WBYTES + ST_RING + ST_EPI);
  const uint32_t bars = smem_base + ST_WBYTES + ST_RING + ST_EPI + ST_PATCH * 4;
  // full[c] +0..16, empty[c] +24..40, tfull[a] +48,56, tempty[a] +64,72, wbar +80, tmem ptr +96
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(sgen + ST_WBYTES + ST_RING + ST_EPI + ST_PATCH * 4 + 96);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // contiguous tile range of this CTA
  const int G = gridDim.x;
  const int t_begin = (int)(((long long)blockIdx.x * p.tiles) / G);
  const int t_end = (int)(((long long)(blockIdx.x + 1) * p.tiles) / G);
  const int ntiles = t_end - t_begin;

  if (warp == 0) {
    if (lane == 0) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmO) : "memory");
      for (int c = 0; c < 3; ++c) {
        mbar_init(bars + 8 * c, c == 2 ? 256 : 128);  // full: every builder thread of the chunk arrives
        mbar_init(bars + 24 + 8 * c, 1);   // empty: one tcgen05.commit
      }
      for (int a = 0; a < 2; ++a) {
        mbar_init(bars + 48 + 8 * a, 1);   // tfull: one tcgen05.commit
        mbar_init(bars + 64 + 8 * a, 4);   // tempty: one arrive per epilogue warp
      }
      mbar_init(bars + 80, 1);
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
    if (lane == 0 && ntiles > 0) {
      constexpr int NPL = (NPROD == 3) ? 2 : 1;
      mbar_expect_tx(bars + 80, NPL * 3 * ST_WCHUNK);
      for (int i = 0; i < NPL * 3; ++i) bulk_load(sW + i * ST_WCHUNK, p.wimg + (size_t)i * ST_WCHUNK, ST_WCHUNK, bars + 80);
      mbar_wait(bars + 80, 0);
      // instruction descriptor: c=f32, a=b=bf16, K-major, N=160, M=128
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(ST_N >> 3) << 17) | ((128u >> 4) << 24);
      for (int it = 0; it < ntiles; ++it) {
        const int acc = it & 1;
        mbar_wait(bars + 64 + 8 * acc, (((uint32_t)it >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + (uint32_t)acc * 256u;
        uint32_t accum = 0;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
          mbar_wait(bars + 8 * c, (uint32_t)it & 1u);
          tc_fence_after();
          const uint64_t a_hi = make_sw128_desc(sA + (c & 1) * ST_ASLOT);
          const uint64_t a_lo = make_sw128_desc(sA + (c & 1) * ST_ASLOT + ST_APLANE);
          const uint64_t w_hi = make_sw128_desc(sW + c * ST_WCHUNK);
          const uint64_t w_lo = make_sw128_desc(sW + (3 + c) * ST_WCHUNK);
          const int ksteps = (c == 2) ? 2 : 4;
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            umma_bf16(d, a_hi + o, w_hi + o, idesc, accum);
            accum = 1;
            if (NPROD == 3) {
              umma_bf16(d, a_hi + o, w_lo + o, idesc, 1);
              umma_bf16(d, a_lo + o, w_hi + o, idesc, 1);
            }
          }
          umma_commit(bars + 24 + 8 * c);
        }
        umma_commit(bars + 48 + 8 * acc);
      }
    }
  } else if (warp <= 8) {
    // ===================== A builders (2 groups x 128 threads, one A row per thread) =====================
    // group 0 builds K-chunk 0 and the first half of chunk 2, group 1 chunk 1 and the second half of chunk 2
    const int bt = threadIdx.x - 32;
    const int grp = bt >> 7;
    const int row = bt & 127;
    const int yy = row >> 6, xx = row & 63;
    const int r8 = row & 7;
    const int atom_off = (row >> 3) * 1024 + r8 * 128;
    const uint32_t* pbase = patch + yy * ST_PW + xx;
    // per-thread patch elements: idx = bt + 256*k, k < 7 (1680 words); (ch, py, px) are tile-independent
    constexpr int NPRE = (ST_PATCH + 255) / 256;
    int pch[NPRE];  // ch*65536 | py<<8 | px  packed
#pragma unroll
    for (int k = 0; k < NPRE; ++k) {
      const int idx = bt + 256 * k;
      const int ch = idx / (ST_PH * ST_PW);
      const int r = idx - ch * (ST_PH * ST_PW);
      const int py = r / ST_PW, px = r - py * ST_PW;
      pch[k] = (idx < ST_PATCH) ? ((ch << 16) | (py << 8) | px) : -1;
    }
    float pre[NPRE];
    const float* in_base = p.io->input;
    auto prefetch = [&](int t) {
      const int img = t >> 9, rem = t & 511;
      const int y0 = (rem >> 2) * 2, x0 = (rem & 3) * 64;
      const float* src = in_base + (size_t)img * 3 * 65536;
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        pre[k] = 0.f;
        if (pch[k] >= 0) {
          const int ch = pch[k] >> 16, py = (pch[k] >> 8) & 255, px = pch[k] & 255;
          pre[k] = __ldg(src + ch * 65536 + st_reflect(y0 - 3 + py) * 256 + st_reflect(x0 - 3 + px));
        }
      }
    };
    if (ntiles > 0) prefetch(t_begin);
    for (int it = 0; it < ntiles; ++it) {
      if (it > 0) asm volatile("bar.sync 1, 256;" ::: "memory");  // everyone is done reading the old patch
#pragma unroll
      for (int k = 0; k < NPRE; ++k) {
        if (pch[k] >= 0) {
          const float v = pre[k];
          const __nv_bfloat16 h = __float2bfloat16_rn(v);
          const __nv_bfloat16 l = __float2bfloat16_rn(v - __bfloat162float(h));
          patch[bt + 256 * k] = (uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(l) << 16);
        }
      }
      asm volatile("bar.sync 1, 256;" ::: "memory");
      // the next tile's patch loads fly while this tile's chunks are built
      if (it + 1 < ntiles) prefetch(t_begin + it + 1);
      // chunk c is built in slot c & 1: chunk 0 reuses the slot of the previous tile's chunk 2,
      // chunk 1 the slot of the previous tile's chunk 1, chunk 2 the slot of this tile's chunk 0
      const uint32_t par = ((uint32_t)it & 1u) ^ 1u;
      if (grp == 0) {
        mbar_wait(bars + 24 + 16, par);
        if (!(p.dbg & 4)) stem_build_chunk<0, 0, 8>(pbase, gA, r8, atom_off);
        fence_proxy_async();
        mbar_arrive(bars + 0);
        mbar_wait(bars + 24 + 0, (uint32_t)it & 1u);
        if (!(p.dbg & 4)) stem_build_chunk<2, 0, 2>(pbase, gA, r8, atom_off);
      } else {
        mbar_wait(bars + 24 + 8, par);
        if (!(p.dbg & 4)) stem_build_chunk<1, 0, 8>(pbase, gA + ST_ASLOT, r8, atom_off);
        fence_proxy_async();
        mbar_arrive(bars + 8);
        mbar_wait(bars + 24 + 0, (uint32_t)it & 1u);
        if (!(p.dbg & 4)) stem_build_chunk<2, 2, 4>(pbase, gA, r8, atom_off);
      }
      fence_proxy_async();
      mbar_arrive(bars + 16);
    }
  } else {
    // ===================== epilogue (warps 9..12) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row0 = q * 32;
    const int yy = row0 >> 6, xx0 = row0 & 63;
    uint8_t* slab_gen = epi_gen + q * 8192;
    const uint32_t slab_s = epi_s + q * 8192;
    uint32_t blk = 0;
    stat_t ssum[5], ssq[5];  // fixed point: any grouping of the tiles gives the same totals
#pragma unroll
    for (int g = 0; g < 5; ++g) { ssum[g] = 0; ssq[g] = 0; }
    for (int it = 0; it < ntiles; ++it) {
      const int t = t_begin + it;
      const int img = t >> 9, rem = t & 511;
      const int y0 = (rem >> 2) * 2, x0 = (rem & 3) * 64;
      const int acc = it & 1;
      mbar_wait(bars + 48 + 8 * acc, ((uint32_t)it >> 1) & 1u);
      tc_fence_after();
#pragma unroll
      for (int g = 0; g < 5; ++g, ++blk) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)row0 << 16) + (uint32_t)(acc * 256 + g * 32), v);
        const uint32_t sl = (blk & 1u) * 4096;
        if (!(p.dbg & 1)) epi_store_block<1>(v, slab_gen + sl, slab_s + sl, lane, &p.tmO, g * 32, x0 + xx0, y0 + yy, img);
        if (p.dbg & 2) continue;
        float cs, cq;
        slab_colsums(slab_gen + sl, lane, &cs, &cq);
        ssum[g] += stat_fix(cs);
        ssq[g] += stat_fix(cq);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 64 + 8 * acc);
      const bool flush = (it == ntiles - 1) || (((t + 1) >> 9) != img);
      if (flush) {
#pragma unroll
        for (int g = 0; g < 5; ++g) {
          stat_t* st = p.stats + ((size_t)img * ST_N + g * 32 + lane) * 2;
          stat_add(st, ssum[g]);
          stat_add(st + 1, ssq[g]);
          ssum[g] = 0;
          ssq[g] = 0;
        }
      }
    }
    if (lane == 0) bulk_wait<0>();
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// weight image: src Conv2d weight [Cout_s][3][7][7] fp32 of ONE stem -> rows [coff, coff+Cout_s) of the
// pre-swizzled [plane][chunk][160][64] bf16 image.  k = (ky*7 + kx)*3 + ch, zero for k >= 147.
// ------------------------------------------------------------------------------------------------
__global__ void pack_stem_umma_kernel(const float* __restrict__ src, int cout_s, int coff, uint8_t* __restrict__ img) {
  const int total = cout_s * 192;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
    const int n_local = i / 192, kk = i - n_local * 192;
    const int n = coff + n_local;
    float w = 0.f;
    if (kk < ST_K) {
      const int tap = kk / 3, ch = kk - tap * 3;
      w = src[((size_t)n_local * 3 + ch) * 49 + tap];
    }
    const __nv_bfloat16 h = __float2bfloat16_rn(w);
    const __nv_bfloat16 l = __float2bfloat16_rn(w - __bfloat162float(h));
    const int c = kk >> 6, k = kk & 63;
    const size_t off = (size_t)c * ST_WCHUNK + (n >> 3) * 1024 + (n & 7) * 128 + (((k >> 3) ^ (n & 7)) << 4) + (k & 7) * 2;
    *reinterpret_cast<__nv_bfloat16*>(img + off) = h;
    *reinterpret_cast<__nv_bfloat16*>(img + (size_t)3 * ST_WCHUNK + off) = l;
  }
}

int launch_pack_stem_umma(const float* src, int cout_s, int coff, uint8_t* img, cudaStream_t st) {
  pack_stem_umma_kernel<<<(cout_s * 192 + 255) / 256, 256, 0, st>>>(src, cout_s, coff, img);
  AP_CUDA(cudaGetLastError());
  return AP_OK;
}

size_t stem_umma_weight_bytes() { return ST_WBYTES; }

static int g_stem_sms = 0;

int stem_umma_init_device() {
  int dev = 0;
  AP_CUDA(cudaGetDevice(&dev));
  AP_CUDA(cudaDeviceGetAttribute(&g_stem_sms, cudaDevAttrMultiProcessorCount, dev));
  AP_CUDA(cudaFuncSetAttribute(stem_umma_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
  AP_CUDA(cudaFuncSetAttribute(stem_umma_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ST_SMEM));
  return AP_OK;
}

int launch_stem_umma(const IoPtrs* io, const uint8_t* wimg, const Raw& out, int B, int nprod, cudaStream_t st) {
  AP_TRY(umma_init());
  StemP p{};
  AP_TRY(tmap_encode_out(&p.tmO, out.p, B, 256, 256, ST_N, 1, 0, 0));
  p.io = io; p.wimg = wimg; p.out = out.p; p.stats = out.stats; p.tiles = B * 512;
  {
    static int dbg = -1;
    if (dbg < 0) { const char* e = getenv("AP_STEM_DBG"); dbg = e ? atoi(e) : 0; }
    p.dbg = dbg;
  }
  const int grid = p.tiles < g_stem_sms ? p.tiles : g_stem_sms;
  if (nprod == 3)
    stem_umma_kernel<3><<<grid, ST_THREADS, ST_SMEM, st>>>(p);
  else
    stem_umma_kernel<1><<<grid, ST_THREADS, ST_SMEM, st>>>(p);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

}  // namespace ap
