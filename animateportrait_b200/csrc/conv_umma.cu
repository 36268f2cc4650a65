// tcgen05 implicit-GEMM convolution for sm_100a.
//
// One CTA computes a 128-pixel x BN-channel tile of a tap-list convolution (ConvGeom, common.cuh):
//   D[128 px, BN] = sum over taps t, channel chunks c0:  A_t[128 px, 64 ch] * W_t[BN, 64 ch]^T
// * A tiles come straight from the NHWC bf16 activation buffer through a 4-D TMA box
//   {64 ch, TW px, TH rows, 1 image} placed at (c0, x0*stride + dx_t, y0*stride + dy_t, n): no im2col.
//   Reflection padding = halo ring written by the producer; zero padding = TMA out-of-bounds fill on
//   the un-haloed view; stride-2 convs use TMA elementStrides = 2; the four phases of
//   ConvTranspose2d(k3,s2,p1,op1) are tap subsets with a strided output scatter.
// * W tiles come from tap-major, K-major packed weights [slab][Cout][Cin] through a 3-D TMA box.
// * Both land in shared memory in the 128-byte-swizzled K-major layout that tcgen05.mma reads via
//   shared-memory descriptors; accumulation is fp32 in TMEM.
// * NPROD = 3 runs A_hi*W_hi + A_hi*W_lo + A_lo*W_hi (bf16 hi/lo split) into the same accumulator:
//   fp32-accurate results (9.5e-5 end to end); NPROD = 1 is plain bf16.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
//   warps 2..5 = epilogue: tcgen05.ld the accumulator, store raw fp32 NHWC, and reduce per-channel
//   sum / sum-of-squares for the following InstanceNorm with a 31-shuffle butterfly per 32 columns.
#include <cudaTypedefs.h>

#include "common.cuh"
#include "umma.cuh"

namespace ap {

void launches_add(int n);

struct alignas(64) UmmaParams {
  CUtensorMap tmA[2];  // hi, lo
  CUtensorMap tmW[2];
  int ntaps;
  int8_t dy[9], dx[9];
  uint8_t slab[9];
  int kchunks, last_ksteps, cin_off;
  int tiles_x, tiles_y, TW, TH, stride;
  float* out;
  int os, py, px, Hout, Wout, out_C, out_coff;
  double* stats;
  int stat_C, stat_coff;
};

struct UmmaConv {
  UmmaParams p;
  int BN, nprod, stages;
  dim3 grid;
  size_t smem;
};

constexpr int A_TILE_BYTES = 128 * 128;  // 128 pixels x 64 bf16

template <int BN, int NPROD>
struct UmmaCfg {
  static constexpr int W_TILE_BYTES = BN * 128;
  static constexpr int STAGE_BYTES = (NPROD == 3 ? 2 : 1) * (A_TILE_BYTES + W_TILE_BYTES);
  static constexpr int STAGES_RAW = (200 * 1024) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr int EXTRA_BYTES = 256 + 2 * BN * 4;  // barriers + tmem ptr, stat partials
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + EXTRA_BYTES;
};

template <int BN, int NPROD>
__global__ void __launch_bounds__(192, 1) conv_umma_kernel(const __grid_constant__ UmmaParams p) {
  using Cfg = UmmaCfg<BN, NPROD>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t bars = smem_base + STAGES * Cfg::STAGE_BYTES;
  // barrier layout: full[s] at bars + 8*s, empty[s] at bars + 64 + 8*s, tmem_full at bars + 128, tmem ptr at +136
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 136);
  float* s_sum = reinterpret_cast<float*>(smem_gen + STAGES * Cfg::STAGE_BYTES + 256);
  float* s_sq = s_sum + BN;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

  // tile coordinates
  int t = blockIdx.x;
  const int tx = t % p.tiles_x; t /= p.tiles_x;
  const int ty = t % p.tiles_y; t /= p.tiles_y;
  const int img = t;
  const int n0 = blockIdx.y * BN;
  const int iters = p.ntaps * p.kchunks;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW[0]) : "memory");
    if (NPROD == 3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[1]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW[1]) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 64 + 8 * s, 1);
    }
    mbar_init(bars + 128, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                 "r"((uint32_t)BN)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  if (warp >= 2) {
    for (int c = threadIdx.x - 64; c < 2 * BN; c += 128) s_sum[c] = 0.f;
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer =====================
    if (lane == 0) {
      const int x0 = tx * p.TW * p.stride, y0 = ty * p.TH * p.stride;
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(bars + 64 + 8 * s, ph ^ 1u);
        const int tap = it / p.kchunks, chunk = it - tap * p.kchunks;
        const uint32_t full = bars + 8 * s;
        mbar_expect_tx(full, Cfg::STAGE_BYTES);
        const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
        const int ca = p.cin_off + chunk * 64, cx = x0 + p.dx[tap], cy = y0 + p.dy[tap];
        tma_load_4d(sa, &p.tmA[0], full, ca, cx, cy, img);
        if (NPROD == 3) {
          tma_load_4d(sa + A_TILE_BYTES, &p.tmA[1], full, ca, cx, cy, img);
          tma_load_3d(sa + 2 * A_TILE_BYTES, &p.tmW[0], full, chunk * 64, n0, p.slab[tap]);
          tma_load_3d(sa + 2 * A_TILE_BYTES + Cfg::W_TILE_BYTES, &p.tmW[1], full, chunk * 64, n0, p.slab[tap]);
        } else {
          tma_load_3d(sa + A_TILE_BYTES, &p.tmW[0], full, chunk * 64, n0, p.slab[tap]);
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
      // K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((128u >> 4) << 24);
      uint32_t first = 0;  // 0 for the very first MMA of the tile (overwrite), 1 afterwards
      for (int it = 0; it < iters; ++it) {
        const int s = it % STAGES;
        const uint32_t ph = (uint32_t)(it / STAGES) & 1u;
        mbar_wait(bars + 8 * s, ph);
        asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
        const int chunk = it % p.kchunks;
        const int ksteps = (chunk == p.kchunks - 1) ? p.last_ksteps : 4;
        const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
        const uint64_t a_hi = make_sw128_desc(sa);
        if (NPROD == 3) {
          const uint64_t a_lo = make_sw128_desc(sa + A_TILE_BYTES);
          const uint64_t w_hi = make_sw128_desc(sa + 2 * A_TILE_BYTES);
          const uint64_t w_lo = make_sw128_desc(sa + 2 * A_TILE_BYTES + Cfg::W_TILE_BYTES);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t o = (uint64_t)(k * 2);  // +32 bytes (16 bf16) inside the 128-B swizzle row, >>4
            umma_bf16(tmem_base, a_hi + o, w_hi + o, idesc, first);
            first = 1;
            umma_bf16(tmem_base, a_hi + o, w_lo + o, idesc, 1);
            umma_bf16(tmem_base, a_lo + o, w_hi + o, idesc, 1);
          }
        } else {
          const uint64_t w_hi = make_sw128_desc(sa + A_TILE_BYTES);
          for (int k = 0; k < ksteps; ++k) {
            const uint64_t o = (uint64_t)(k * 2);
            umma_bf16(tmem_base, a_hi + o, w_hi + o, idesc, first);
            first = 1;
          }
        }
        umma_commit(bars + 64 + 8 * s);             // frees the smem stage when these MMAs retire
        if (it == iters - 1) umma_commit(bars + 128);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;                          // TMEM lane quarter this warp may access
    const int row = q * 32 + lane;                   // accumulator row = pixel index inside the tile
    const int yy = row / p.TW, xx = row - yy * p.TW;
    const int oy = (ty * p.TH + yy) * p.os + p.py, ox = (tx * p.TW + xx) * p.os + p.px;
    float* orow = p.out + ((size_t)(img * p.Hout + oy) * p.Wout + ox) * p.out_C + p.out_coff + n0;
    mbar_wait(bars + 128, 0);
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      float v[32];
      tmem_ld32(tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)c0, v);
#pragma unroll
      for (int j = 0; j < 32; j += 4)
        *reinterpret_cast<float4*>(orow + c0 + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
      if (p.stats != nullptr) {
        float sq[32];
#pragma unroll
        for (int j = 0; j < 32; ++j) sq[j] = v[j] * v[j];
        const float cs = butterfly_colsum(v, lane);
        const float cq = butterfly_colsum(sq, lane);
        atomicAdd(&s_sum[c0 + lane], cs);
        atomicAdd(&s_sq[c0 + lane], cq);
      }
    }
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
    if (p.stats != nullptr) {
      asm volatile("bar.sync 1, 128;" ::: "memory");
      for (int c = threadIdx.x - 64; c < BN; c += 128) {
        double* st = p.stats + ((size_t)img * p.stat_C + p.stat_coff + n0 + c) * 2;
        atomicAdd(st, (double)s_sum[c]);
        atomicAdd(st + 1, (double)s_sq[c]);
      }
    }
  }
  __syncthreads();
  if (warp == 1) {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"((uint32_t)BN) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;

template <int BN, int NPROD>
static int set_attr() {
  AP_CUDA(cudaFuncSetAttribute(conv_umma_kernel<BN, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)UmmaCfg<BN, NPROD>::SMEM));
  return AP_OK;
}

int umma_init() {
  if (g_encode) return AP_OK;
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult qres;
  AP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
  AP_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, AP_ERR_CUDA,
             "cuTensorMapEncodeTiled not available from the driver");
  AP_TRY((set_attr<64, 1>()));
  AP_TRY((set_attr<128, 1>()));
  AP_TRY((set_attr<256, 1>()));
  AP_TRY((set_attr<64, 3>()));
  AP_TRY((set_attr<128, 3>()));
  AP_TRY((set_attr<256, 3>()));
  g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  return AP_OK;
}

static int encode(CUtensorMap* m, const void* base, int rank, const cuuint64_t* dims, const cuuint64_t* strides,
                  const cuuint32_t* box, const cuuint32_t* estr) {
  CUresult r = g_encode(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), dims, strides,
                        box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AP_REQUIRE(r == CUDA_SUCCESS, AP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d)", (int)r, rank);
  return AP_OK;
}

// `in` must be a bf16 activation. Zero-padded convs address the un-haloed interior (TMA fills
// out-of-bounds with zeros); reflect-padded ones address the haloed buffer (halo = in.pad >= conv pad).
int umma_conv_create(UmmaConv** out, const ConvGeom& g, const Act& in, int in_coff, const __nv_bfloat16* w_hi,
                     const __nv_bfloat16* w_lo, int nprod, float* out_raw, int out_C, int out_coff, double* stats,
                     int stat_C, int stat_coff) {
  AP_TRY(umma_init());
  AP_REQUIRE(in.fmt == FMT_BF16X2 || in.fmt == FMT_BF16, AP_ERR_INVALID, "umma conv needs bf16 activations");
  AP_REQUIRE(nprod == 1 || (nprod == 3 && in.fmt == FMT_BF16X2 && w_lo), AP_ERR_INVALID, "umma conv: nprod/format");
  AP_REQUIRE(g.taps.n >= 1 && g.taps.n <= 9, AP_ERR_INVALID, "umma conv: %d taps", g.taps.n);
  AP_REQUIRE(g.Cout == 64 || g.Cout == 128 || g.Cout == 256, AP_ERR_UNSUPPORTED, "umma conv: Cout=%d", g.Cout);
  AP_REQUIRE(g.Cin % 16 == 0 && (g.Cin % 64 == 0 || in_coff + g.Cin == in.C), AP_ERR_UNSUPPORTED,
             "umma conv: Cin=%d (coff %d of %d) needs a zero-filled K tail", g.Cin, in_coff, in.C);
  AP_REQUIRE(!g.reflect || in.pad >= 1, AP_ERR_INVALID, "umma conv: reflect padding needs a haloed input");
  AP_REQUIRE(out_C % 4 == 0 && out_coff % 4 == 0, AP_ERR_INVALID, "umma conv: output channel layout");
  int TW, TH;
  if (g.Wv % 128 == 0) { TW = 128; TH = 1; }
  else if (g.Wv == 64 && g.Hv % 2 == 0) { TW = 64; TH = 2; }
  else { set_error("umma conv: virtual grid %dx%d not tileable", g.Hv, g.Wv); return AP_ERR_UNSUPPORTED; }
  AP_REQUIRE(TW * g.stride <= 256, AP_ERR_UNSUPPORTED, "umma conv: TMA box too wide");

  UmmaConv* c = new UmmaConv();
  UmmaParams& p = c->p;
  c->BN = g.Cout;
  c->nprod = nprod;
  // activation maps
  const bool padded_view = g.reflect != 0;
  const int Hp = in.H + 2 * in.pad, Wp = in.W + 2 * in.pad;
  const int org = padded_view ? in.pad : 0;
  const cuuint64_t adims[4] = {(cuuint64_t)in.C, (cuuint64_t)(padded_view ? Wp : in.W),
                               (cuuint64_t)(padded_view ? Hp : in.H), (cuuint64_t)in.B};
  const cuuint64_t astr[3] = {(cuuint64_t)in.C * 2, (cuuint64_t)Wp * in.C * 2, (cuuint64_t)Hp * Wp * in.C * 2};
  const cuuint32_t abox[4] = {64, (cuuint32_t)(TW * g.stride), (cuuint32_t)(TH * g.stride), 1};
  const cuuint32_t aes[4] = {1, (cuuint32_t)g.stride, (cuuint32_t)g.stride, 1};
  const size_t view_off = padded_view ? 0 : ((size_t)in.pad * Wp + in.pad) * in.C;
  int rc = encode(&p.tmA[0], reinterpret_cast<const __nv_bfloat16*>(in.p0) + view_off, 4, adims, astr, abox, aes);
  if (rc == AP_OK && nprod == 3)
    rc = encode(&p.tmA[1], reinterpret_cast<const __nv_bfloat16*>(in.p1) + view_off, 4, adims, astr, abox, aes);
  // weight maps [slab][Cout][Cin]
  for (int i = 0; i < g.taps.n; ++i)
    AP_REQUIRE(g.taps.slab[i] < 9, AP_ERR_INVALID, "umma conv: only 3x3 weight slabs are packed for tcgen05");
  const cuuint64_t wdims[3] = {(cuuint64_t)g.Cin, (cuuint64_t)g.Cout, 9};
  const cuuint64_t wstr[2] = {(cuuint64_t)g.Cin * 2, (cuuint64_t)g.Cin * g.Cout * 2};
  const cuuint32_t wbox[3] = {64, (cuuint32_t)c->BN, 1};
  const cuuint32_t wes[3] = {1, 1, 1};
  if (rc == AP_OK) rc = encode(&p.tmW[0], w_hi, 3, wdims, wstr, wbox, wes);
  if (rc == AP_OK && nprod == 3) rc = encode(&p.tmW[1], w_lo, 3, wdims, wstr, wbox, wes);
  if (rc != AP_OK) { delete c; return rc; }

  p.ntaps = g.taps.n;
  for (int i = 0; i < g.taps.n; ++i) {
    p.dy[i] = (int8_t)(g.taps.dy[i] + org);
    p.dx[i] = (int8_t)(g.taps.dx[i] + org);
    p.slab[i] = g.taps.slab[i];
  }
  p.kchunks = (g.Cin + 63) / 64;
  const int tail = g.Cin - (p.kchunks - 1) * 64;
  p.last_ksteps = (tail + 15) / 16;
  p.cin_off = in_coff;
  p.TW = TW; p.TH = TH;
  p.tiles_x = g.Wv / TW; p.tiles_y = g.Hv / TH;
  p.stride = g.stride;
  p.out = out_raw;
  p.os = g.os; p.py = g.py; p.px = g.px; p.Hout = g.Hout; p.Wout = g.Wout;
  p.out_C = out_C; p.out_coff = out_coff;
  p.stats = stats; p.stat_C = stat_C; p.stat_coff = stat_coff;
  c->grid = dim3((unsigned)(p.tiles_x * p.tiles_y * g.B), (unsigned)(g.Cout / c->BN));
  *out = c;
  return AP_OK;
}

void umma_conv_destroy(UmmaConv* c) { delete c; }

int umma_conv_launch(const UmmaConv* c, cudaStream_t st) {
#define AP_UMMA_CASE(BN_, NP_)                                                                          \
  if (c->BN == BN_ && c->nprod == NP_) {                                                                \
    conv_umma_kernel<BN_, NP_><<<c->grid, 192, UmmaCfg<BN_, NP_>::SMEM, st>>>(c->p);                    \
    AP_CUDA(cudaGetLastError());                                                                        \
    launches_add(1);                                                                                    \
    return AP_OK;                                                                                       \
  }
  AP_UMMA_CASE(64, 1)
  AP_UMMA_CASE(128, 1)
  AP_UMMA_CASE(256, 1)
  AP_UMMA_CASE(64, 3)
  AP_UMMA_CASE(128, 3)
  AP_UMMA_CASE(256, 3)
#undef AP_UMMA_CASE
  set_error("umma conv: no kernel for BN=%d nprod=%d", c->BN, c->nprod);
  return AP_ERR_UNSUPPORTED;
}

}  // namespace ap
