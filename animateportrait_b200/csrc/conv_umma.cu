// tcgen05 implicit-GEMM convolution for sm_100a: persistent, warp-specialised, TMEM double-buffered.
//
// One work item = a 128-pixel x bn-channel tile of a tap-list convolution (ConvGeom, common.cuh):
//   D[128 px, bn] = sum over taps t, channel chunks c0:  A_t[128 px, 64 ch] * W_t[bn, 64 ch]^T
// * A tiles come straight from the NHWC bf16 activation buffer through a 4-D TMA box
//   {64 ch, TW px, TH rows, 1 image} placed at (c0, x0*stride + dx_t, y0*stride + dy_t, n): no im2col.
//   Reflection padding = halo ring written by the producer; zero padding = TMA out-of-bounds fill on
//   the un-haloed view; stride-2 convs use TMA elementStrides = 2; the four phases of
//   ConvTranspose2d(k3,s2,p1,op1) are tap subsets with a strided output view.
// * W tiles come from tap-major, K-major packed weights [slab][Cout][Cin] through 3-D TMA boxes of 64 rows.
// * Both land in shared memory in the 128-byte-swizzled K-major layout that tcgen05.mma reads via
//   shared-memory descriptors; accumulation is fp32 in TMEM.
// * NPROD = 3 runs A_hi*W_hi + A_hi*W_lo + A_lo*W_hi (bf16 hi/lo split) into the same accumulator:
//   fp32-accurate results (9.5e-5 end to end); NPROD = 1 is plain bf16.
// * CTAs are persistent (grid = #SMs, one CTA per SM) and walk a static item list.  Items are full
//   tiles (bn = Cout) for as many whole waves as the tile count gives; the tiles of the last, partial
//   wave are split along N into 2 or 4 items so that the tail occupies every SM (512 tiles on 148 SMs:
//   3 waves of full tiles + 136 half tiles instead of a fourth, 46%-full wave).
// * Two 256-column TMEM accumulators: the epilogue of item i overlaps the MMAs of item i+1.
// * Warp roles: warp 0 = TMA producer, warp 1 = TMEM allocator + MMA issuer (one elected lane),
//   warps 2..5 = epilogue: tcgen05.ld 32 columns at a time, stage them in a swizzled 4 KB slab and
//   TMA-store the {32 ch, 32 px} box (coalesced, asynchronous), and read the per-channel sum /
//   sum-of-squares of the block back from the slab for the following InstanceNorm, as fixed-point integers
//   (common.cuh: stat_t): accumulated on chip / with fire-and-forget atomics in any order, bit-identically.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include <mutex>

#include "common.cuh"
#include "umma.cuh"

namespace ap {

void launches_add(int n);

// timing diagnostics are compiled in only with -DAP_UMMA_DIAG (tools/conv_probe.sh builds such a library)
#ifdef AP_UMMA_DIAG
#define AP_DBG(x) (x)
#else
#define AP_DBG(x) 0
#endif

struct alignas(64) UmmaParams {
  CUtensorMap tmA[2];  // hi, lo
  CUtensorMap tmW[2];
  CUtensorMap tmWs[2];  // the same weights in boxes of 32 rows: N-split items narrower than one 64-row box (small batches)
  CUtensorMap tmO[4];  // fp32 output view(s): one, or one per output phase of a phase-packed transposed conv
  int phase_cols;      // 0, or channels per phase: column block c of the tile goes to tmO[c / phase_cols]
  int ntaps;
  int8_t dy[9], dx[9];
  uint8_t slab[9];
  int kchunks, last_ksteps, cin_off;
  int tiles_x, tiles_y, TW, TH, stride;
  int out_coff;
  stat_t* stats;            // InstanceNorm statistics of the output, fixed point (null: none)
  int stat_C, stat_coff;
  // item list in units of tile GROUPS (CG consecutive 128-pixel tiles, one per CTA of the pair):
  // n_full whole groups with bn = Cout, then (groups - n_full) * split N-parts
  int n_full, split, n_items;
  int l2_hints;  // bit 0: raw output stores evict_last (AP_NETG_L2_HINTS)
  int dbg;  // timing diagnostics only (AP_UMMA_DBG): 1 = no TMA loads after the first fill, 2 = no output stores, 4 = no statistics
};

struct UmmaConv {
  UmmaParams p;
  int BN, nprod, cg;
  HaloConv* halo = nullptr;  // the layer runs on the halo-staged kernel instead (conv_halo.cu)
  dim3 grid;
  size_t smem;
};

constexpr int A_TILE_BYTES = 128 * 128;  // 128 pixels x 64 bf16
// Epilogue staging slabs (4 KB) per epilogue warp.  The N <= 128 kernels keep one: that leaves ~17 KB of the SM's
// shared memory unallocated, enough for CTAs of the HBM-bound kernels (warps, applies) of the other encoder
// branches to be co-resident with a persistent conv CTA -- without it the side streams cannot overlap anything.
template <int BN>
struct EpiCfg {
  static constexpr int SLABS = BN >= 256 ? 2 : 1;
  static constexpr int BYTES = 4 * SLABS * 4096;
};

template <int BN, int NPROD, int CG>
struct UmmaCfg {
  static constexpr int PLANES = (NPROD == 3 ? 2 : 1);
  static constexpr int W_ROWS = BN / CG;                     // rows of the weight tile this CTA stages
  static constexpr int W_BOX = W_ROWS < 64 ? W_ROWS : 64;    // rows per TMA box
  static constexpr int W_TILE_BYTES = W_ROWS * 128;
  static constexpr int STAGE_BYTES = PLANES * (A_TILE_BYTES + W_TILE_BYTES);
  static constexpr int EPI_BYTES = EpiCfg<BN>::BYTES;
  static constexpr int STAGES_RAW = (226 * 1024 - 2 * 4 * 4096 - 1024 - 256) / STAGE_BYTES;
  static constexpr int STAGES = STAGES_RAW > 8 ? 8 : STAGES_RAW;
  static constexpr size_t SMEM = 1024 + (size_t)STAGES * STAGE_BYTES + EPI_BYTES + 256 + (CG == 1 && BN <= 128 ? 8192 : 0);
};

struct Item {
  int img, ty, tx, n0, bn;
};

// Local item `li` of execution unit `unit` (CTA, or CTA pair) -> global item index, or -1 when the unit is done.
// STRIDED: item = unit + li * nunits (neighbouring CTAs work on neighbouring tiles).  CONTIGUOUS (used where the
// InstanceNorm statistics are accumulated on chip): the whole-wave part of the item list is dealt in runs, unit u
// owning [u*K, (u+1)*K), so that consecutive items of a unit lie in the same image; tail items stay strided.
template <bool CONTIGUOUS>
__device__ __forceinline__ int item_at(const UmmaParams& p, int li, int unit, int nunits, int K) {
  if (!CONTIGUOUS) {
    const int it = unit + li * nunits;
    return it < p.n_items ? it : -1;
  }
  if (li < K) return unit * K + li;
  const int it = p.n_full + (li - K) * nunits + unit;
  return it < p.n_items ? it : -1;
}

// Per-warp accumulation of the InstanceNorm statistics in shared memory across the items of one image: one
// atomic per (channel, warp, image run) instead of one per (channel, warp, tile).  Atomics on one address
// serialise in L2; layers with few channels and thousands of tiles were bound by exactly that
// (profiles/r01_stat_atomics.md: 128->64 transposed conv 307 -> 138 us without statistics).  The accumulators are
// fixed-point integers like the totals: how the items are grouped into runs does not change a bit of the result.
constexpr int STAT_ACC_COLS = 128;
constexpr int STAT_ACC_BYTES = 4 * 2 * STAT_ACC_COLS * 8;  // 4 epilogue warps x {sum, sumsq} x 128 channels x int64

__device__ __forceinline__ void stat_flush(stat_t* sacc, stat_t* stats, int stat_C, int stat_coff, int img, int ncols, int lane) {
  for (int c = lane; c < ncols; c += 32) {
    stat_t* dst = stats + ((size_t)img * stat_C + stat_coff + c) * 2;
    stat_add(dst, sacc[c]);
    stat_add(dst + 1, sacc[STAT_ACC_COLS + c]);
    sacc[c] = 0;
    sacc[STAT_ACC_COLS + c] = 0;
  }
}

// ---- single-CTA kernel (cta_group::1): one CTA per 128-pixel tile; used for the N <= 128 layers ----
__device__ __forceinline__ Item decode_item1(const UmmaParams& p, int item, int BN) {
  int m, n0 = 0, bn = BN;
  if (item < p.n_full) {
    m = item;
  } else {
    const int r = item - p.n_full;
    m = p.n_full + r / p.split;
    bn = BN / p.split;
    n0 = (r % p.split) * bn;
  }
  Item it;
  it.tx = m % p.tiles_x; m /= p.tiles_x;
  it.ty = m % p.tiles_y; m /= p.tiles_y;
  it.img = m;
  it.n0 = n0;
  it.bn = bn;
  return it;
}

template <int BN, int NPROD>
__global__ void __launch_bounds__(192, 1) conv_umma1_kernel(const __grid_constant__ UmmaParams p) {
  using Cfg = UmmaCfg<BN, NPROD, 1>;
  constexpr int STAGES = Cfg::STAGES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_s = smem_base + STAGES * Cfg::STAGE_BYTES;
  uint8_t* epi_gen = smem_gen + STAGES * Cfg::STAGE_BYTES;
  constexpr int EPI_SLABS = EpiCfg<BN>::SLABS;
  constexpr int EPI_BYTES = EpiCfg<BN>::BYTES;
  const uint32_t bars = epi_s + EPI_BYTES;
  // barriers: full[s] +8s, empty[s] +64+8s, tfull[a] +128+8a, tempty[a] +144+8a, tmem ptr +160
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(epi_gen + EPI_BYTES + 160);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iters = p.ntaps * p.kchunks;
  constexpr bool ACC = BN <= STAT_ACC_COLS;  // statistics accumulated on chip, contiguous item runs
  const int Krun = ACC ? p.n_full / (int)gridDim.x : 0;
  stat_t* sacc_all = reinterpret_cast<stat_t*>(epi_gen + EPI_BYTES + 256);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmO[0]) : "memory");
    if (NPROD == 3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[1]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW[1]) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);
      mbar_init(bars + 64 + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bars + 128 + 8 * a, 1);  // tfull: one tcgen05.commit
      mbar_init(bars + 144 + 8 * a, 4);  // tempty: one arrive per epilogue warp
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
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
    // ===================== TMA producer =====================
    if (lane == 0) {
      uint32_t cnt = 0;
      for (int li = 0;; ++li) {
        const int item = item_at<ACC>(p, li, blockIdx.x, gridDim.x, Krun);
        if (item < 0) break;
        const Item w = decode_item1(p, item, BN);
        const int x0 = w.tx * p.TW * p.stride, y0 = w.ty * p.TH * p.stride;
        const bool small = w.bn < 64;  // one 32-row box
        const int nbox = small ? 1 : (w.bn >> 6);
        const CUtensorMap* mW0 = small ? &p.tmWs[0] : &p.tmW[0];
        const CUtensorMap* mW1 = small ? &p.tmWs[1] : &p.tmW[1];
        const uint32_t tx_bytes = (NPROD == 3 ? 2u : 1u) * (uint32_t)(A_TILE_BYTES + w.bn * 128);
        for (int it = 0; it < iters; ++it, ++cnt) {
          const uint32_t s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1u;
          mbar_wait(bars + 64 + 8 * s, ph ^ 1u);
          const int tap = it / p.kchunks, chunk = it - tap * p.kchunks;
          const uint32_t full = bars + 8 * s;
          mbar_expect_tx(full, tx_bytes);
          const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
          const int ca = p.cin_off + chunk * 64, cx = x0 + p.dx[tap], cy = y0 + p.dy[tap];
          tma_load_4d(sa, &p.tmA[0], full, ca, cx, cy, w.img);
          if (NPROD == 3) {
            tma_load_4d(sa + A_TILE_BYTES, &p.tmA[1], full, ca, cx, cy, w.img);
            for (int b = 0; b < nbox; ++b) {
              tma_load_3d(sa + 2 * A_TILE_BYTES + b * 8192, mW0, full, chunk * 64, w.n0 + 64 * b, p.slab[tap]);
              tma_load_3d(sa + 2 * A_TILE_BYTES + Cfg::W_TILE_BYTES + b * 8192, mW1, full, chunk * 64, w.n0 + 64 * b,
                          p.slab[tap]);
            }
          } else {
            for (int b = 0; b < nbox; ++b)
              tma_load_3d(sa + A_TILE_BYTES + b * 8192, mW0, full, chunk * 64, w.n0 + 64 * b, p.slab[tap]);
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    if (lane == 0) {
      uint32_t cnt = 0, local = 0;
      for (int li = 0;; ++li, ++local) {
        const int item = item_at<ACC>(p, li, blockIdx.x, gridDim.x, Krun);
        if (item < 0) break;
        const Item w = decode_item1(p, item, BN);
        // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
        // K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(w.bn >> 3) << 17) | ((128u >> 4) << 24);
        const uint32_t acc = local & 1u;
        mbar_wait(bars + 144 + 8 * acc, ((local >> 1) & 1u) ^ 1u);  // epilogue has drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 256u;
        uint32_t first = 0;  // 0 for the very first MMA of the item (overwrite), 1 afterwards
        for (int it = 0; it < iters; ++it, ++cnt) {
          const uint32_t s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1u;
          mbar_wait(bars + 8 * s, ph);
          tc_fence_after();
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
              umma_bf16(d, a_hi + o, w_hi + o, idesc, first);
              first = 1;
              umma_bf16(d, a_hi + o, w_lo + o, idesc, 1);
              umma_bf16(d, a_lo + o, w_hi + o, idesc, 1);
            }
          } else {
            const uint64_t w_hi = make_sw128_desc(sa + A_TILE_BYTES);
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t o = (uint64_t)(k * 2);
              umma_bf16(d, a_hi + o, w_hi + o, idesc, first);
              first = 1;
            }
          }
          umma_commit(bars + 64 + 8 * s);  // frees the smem stage when these MMAs retire
        }
        umma_commit(bars + 128 + 8 * acc);  // accumulator complete
      }
    }
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row0 = q * 32;
    const int yy = row0 / p.TW, xx0 = row0 - yy * p.TW;
    uint8_t* slab_gen = epi_gen + q * (EPI_SLABS * 4096);
    const uint32_t slab_s = epi_s + q * (EPI_SLABS * 4096);
    const uint64_t opol = (p.l2_hints & 1) ? l2_policy_evict_last() : 0;  // raw output is re-read by the next kernel
    uint32_t local = 0, blk = 0;
    stat_t* sacc = sacc_all + q * (2 * STAT_ACC_COLS);
    int simg = -1;
    if (ACC) {
      for (int c = lane; c < 2 * STAT_ACC_COLS; c += 32) sacc[c] = 0;
      __syncwarp();
    }
    for (int li = 0;; ++li, ++local) {
      const int item = item_at<ACC>(p, li, blockIdx.x, gridDim.x, Krun);
      if (item < 0) break;
      const Item w = decode_item1(p, item, BN);
      const bool acc_item = ACC && w.bn == BN;  // N-split tail items use direct atomics
      if (ACC && simg >= 0 && (w.img != simg || !acc_item)) {
        stat_flush(sacc, p.stats, p.stat_C, p.stat_coff, simg, BN, lane);
        simg = -1;
      }
      const uint32_t acc = local & 1u;
      mbar_wait(bars + 128 + 8 * acc, (local >> 1) & 1u);
      tc_fence_after();
      const int ox = w.tx * p.TW + xx0, oy = w.ty * p.TH + yy;
      stat_t* strow = p.stats ? p.stats + ((size_t)w.img * p.stat_C + p.stat_coff + w.n0 + lane) * 2 : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < w.bn; c0 += 32, ++blk) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)row0 << 16) + acc * 256u + (uint32_t)c0, v);
        const uint32_t sl = (blk % EPI_SLABS) * 4096;
        epi_store_block<EPI_SLABS - 1>(v, slab_gen + sl, slab_s + sl, lane, &p.tmO[0], p.out_coff + w.n0 + c0, ox, oy, w.img, opol);
        if (strow != nullptr) {
          float cs, cq;
          slab_colsums(slab_gen + sl, lane, &cs, &cq);
          if (acc_item) {
            sacc[c0 + lane] += stat_fix(cs);
            sacc[STAT_ACC_COLS + c0 + lane] += stat_fix(cq);
            simg = w.img;
          } else {
            stat_add(strow + (size_t)c0 * 2, stat_fix(cs));
            stat_add(strow + (size_t)c0 * 2 + 1, stat_fix(cq));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(bars + 144 + 8 * acc);
    }
    if (ACC && simg >= 0) stat_flush(sacc, p.stats, p.stat_C, p.stat_coff, simg, BN, lane);
    if (lane == 0) bulk_wait<0>();  // all output boxes have landed before the CTA exits
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}


// ---- CTA-pair capable kernel ----
__device__ __forceinline__ Item decode_item(const UmmaParams& p, int item, int BN, int CG, int rank) {
  int g, n0 = 0, bn = BN;
  if (item < p.n_full) {
    g = item;
  } else {
    const int r = item - p.n_full;
    g = p.n_full + r / p.split;
    bn = BN / p.split;
    n0 = (r % p.split) * bn;
  }
  int m = g * CG + rank;
  Item it;
  it.tx = m % p.tiles_x; m /= p.tiles_x;
  it.ty = m % p.tiles_y; m /= p.tiles_y;
  it.img = m;
  it.n0 = n0;
  it.bn = bn;
  return it;
}

// CG = 1: one CTA per 128-pixel tile.  CG = 2: a CTA pair (cluster of 2 on one TPC) runs tcgen05.mma.cta_group::2
// with M = 256; each CTA stages its own A tile and half of the weight tile, the leader (cluster rank 0) issues.
// PACKED (phase-packed transposed convs): statistics accumulated on chip over contiguous item runs (the phases fold onto
// the same channels), one staging slab.
template <int BN, int NPROD, int CG, bool PACKED>
__global__ void __launch_bounds__(192, 1) conv_umma_kernel(const __grid_constant__ UmmaParams p) {
  using Cfg = UmmaCfg<BN, NPROD, CG>;
  constexpr int STAGES = Cfg::STAGES;
  constexpr int PLANES = Cfg::PLANES;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;  // SWIZZLE_128B wants 1024-B alignment
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t epi_s = smem_base + STAGES * Cfg::STAGE_BYTES;
  uint8_t* epi_gen = smem_gen + STAGES * Cfg::STAGE_BYTES;
  constexpr int EPI_SLABS = PACKED ? 1 : EpiCfg<BN>::SLABS;
  constexpr int EPI_BYTES = 4 * EPI_SLABS * 4096;
  constexpr bool ACC = PACKED;
  const uint32_t bars = epi_s + EPI_BYTES;
  // barriers: full[s] +8s, empty[s] +64+8s, tfull[a] +128+8a, tempty[a] +144+8a, tmem ptr +160
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(epi_gen + EPI_BYTES + 160);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int iters = p.ntaps * p.kchunks;
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const int unit = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;   // pair (or CTA) index
  const int nunits = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int Krun = ACC ? p.n_full / nunits : 0;
  stat_t* sacc_all = reinterpret_cast<stat_t*>(epi_gen + EPI_BYTES + 256);

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmO[0]) : "memory");
    if (NPROD == 3) {
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[1]) : "memory");
      asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW[1]) : "memory");
    }
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(bars + 8 * s, 1);       // full: the (leader's) producer arrive.expect_tx
      mbar_init(bars + 64 + 8 * s, 1);  // empty: one tcgen05.commit (multicast to both CTAs of a pair)
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bars + 128 + 8 * a, 1);       // tfull: one tcgen05.commit
      mbar_init(bars + 144 + 8 * a, 4 * CG);  // tempty (leader's is used): one arrive per epilogue warp of the pair
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)),
                   "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's barriers must be initialised before anything is signalled remotely
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 0) {
    // ===================== TMA producer (every CTA loads its own A tile and its share of W) =====================
    if (lane == 0) {
      const uint32_t full0 = (CG == 2) ? mapa_rank(bars, 0) : bars;  // full barriers live in the leader
      uint32_t cnt = 0;
      for (int li = 0;; ++li) {
        const int item = item_at<ACC>(p, li, unit, nunits, Krun);
        if (item < 0) break;
        const Item w = decode_item(p, item, BN, CG, rank);
        const int x0 = w.tx * p.TW * p.stride, y0 = w.ty * p.TH * p.stride;
        const int wrows = w.bn / CG;
        const bool small = wrows < Cfg::W_BOX;  // one 32-row box per plane
        const int nbox = small ? 1 : wrows / Cfg::W_BOX;
        const CUtensorMap* mW = small ? p.tmWs : p.tmW;
        const int wrow0 = w.n0 + rank * wrows;
        const uint32_t tx_bytes = (uint32_t)PLANES * (uint32_t)(A_TILE_BYTES + wrows * 128);
        for (int it = 0; it < iters; ++it, ++cnt) {
          const uint32_t s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1u;
          mbar_wait(bars + 64 + 8 * s, ph ^ 1u);
          const int tap = it / p.kchunks, chunk = it - tap * p.kchunks;
          const uint32_t full = full0 + 8 * s;
          if (AP_DBG(p.dbg & 1) && cnt >= (uint32_t)STAGES) {
            if (rank == 0) mbar_arrive(bars + 8 * s);
            continue;
          }
          if (rank == 0) mbar_expect_tx(bars + 8 * s, CG * tx_bytes);
          const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t sw = sa + PLANES * A_TILE_BYTES;
          const int ca = p.cin_off + chunk * 64, cx = x0 + p.dx[tap], cy = y0 + p.dy[tap];
          // issue order: both A planes first, then the W boxes with hi/lo interleaved.  The order matters: identical
          // W requests of different SMs that reach L2 close together are served once (measured: planes-outer order
          // is 12-20 % slower on the L2-fabric-bound layers)
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            if (CG == 2) tma_load_4d_pair(sa + pl * A_TILE_BYTES, &p.tmA[pl], full, ca, cx, cy, w.img);
            else tma_load_4d(sa + pl * A_TILE_BYTES, &p.tmA[pl], full, ca, cx, cy, w.img);
          }
          for (int b = 0; b < nbox; ++b) {
#pragma unroll
            for (int pl = 0; pl < PLANES; ++pl) {
              const uint32_t dst = sw + pl * Cfg::W_TILE_BYTES + b * (Cfg::W_BOX * 128);
              if (CG == 2) tma_load_3d_pair(dst, &mW[pl], full, chunk * 64, wrow0 + Cfg::W_BOX * b, p.slab[tap]);
              else tma_load_3d(dst, &mW[pl], full, chunk * 64, wrow0 + Cfg::W_BOX * b, p.slab[tap]);
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of the pair only) =====================
    if (lane == 0 && rank == 0) {
      uint32_t cnt = 0, local = 0;
      for (int li = 0;; ++li, ++local) {
        const int item = item_at<ACC>(p, li, unit, nunits, Krun);
        if (item < 0) break;
        const Item w = decode_item(p, item, BN, CG, 0);
        // instruction descriptor (cute::UMMA::InstrDescriptor): c=f32 [4,6)=1, a=bf16 [7,10)=1, b=bf16 [10,13)=1,
        // K-major A and B (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29)
        const uint32_t idesc =
            (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(w.bn >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
        const uint32_t acc = local & 1u;
        mbar_wait(bars + 144 + 8 * acc, ((local >> 1) & 1u) ^ 1u);  // epilogues have drained this accumulator
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 256u;
        uint32_t first = 0;  // 0 for the very first MMA of the item (overwrite), 1 afterwards
        for (int it = 0; it < iters; ++it, ++cnt) {
          const uint32_t s = cnt % STAGES;
          const uint32_t ph = (cnt / STAGES) & 1u;
          mbar_wait(bars + 8 * s, ph);
          tc_fence_after();
          const int chunk = it % p.kchunks;
          const int ksteps = (chunk == p.kchunks - 1) ? p.last_ksteps : 4;
          const uint32_t sa = smem_base + s * Cfg::STAGE_BYTES;
          const uint32_t sw = sa + PLANES * A_TILE_BYTES;
          const uint64_t a_hi = make_sw128_desc(sa);
          const uint64_t w_hi = make_sw128_desc(sw);
          if (NPROD == 3) {
            const uint64_t a_lo = make_sw128_desc(sa + A_TILE_BYTES);
            const uint64_t w_lo = make_sw128_desc(sw + Cfg::W_TILE_BYTES);
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t o = (uint64_t)(k * 2);  // +32 bytes (16 bf16) inside the 128-B swizzle row, >>4
              if (CG == 2) {
                umma_bf16_pair(d, a_hi + o, w_hi + o, idesc, first);
                umma_bf16_pair(d, a_hi + o, w_lo + o, idesc, 1);
                umma_bf16_pair(d, a_lo + o, w_hi + o, idesc, 1);
              } else {
                umma_bf16(d, a_hi + o, w_hi + o, idesc, first);
                umma_bf16(d, a_hi + o, w_lo + o, idesc, 1);
                umma_bf16(d, a_lo + o, w_hi + o, idesc, 1);
              }
              first = 1;
            }
          } else {
            for (int k = 0; k < ksteps; ++k) {
              const uint64_t o = (uint64_t)(k * 2);
              if (CG == 2) umma_bf16_pair(d, a_hi + o, w_hi + o, idesc, first);
              else umma_bf16(d, a_hi + o, w_hi + o, idesc, first);
              first = 1;
            }
          }
          // frees the smem stage (in both CTAs of a pair) when these MMAs retire
          if (CG == 2) umma_commit_pair(bars + 64 + 8 * s);
          else umma_commit(bars + 64 + 8 * s);
        }
        // accumulator complete
        if (CG == 2) umma_commit_pair(bars + 128 + 8 * acc);
        else umma_commit(bars + 128 + 8 * acc);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5) =====================
    const int q = warp & 3;  // TMEM lane quarter this warp may access
    const int row0 = q * 32;
    const int yy = row0 / p.TW, xx0 = row0 - yy * p.TW;
    uint8_t* slab_gen = epi_gen + q * (EPI_SLABS * 4096);
    const uint32_t slab_s = epi_s + q * (EPI_SLABS * 4096);
    const uint32_t tempty0 = (CG == 2) ? mapa_rank(bars + 144, 0) : bars + 144;
    const uint64_t opol = (p.l2_hints & 1) ? l2_policy_evict_last() : 0;  // raw output is re-read by the next kernel
    uint32_t local = 0, blk = 0;
    stat_t* sacc = sacc_all + q * (2 * STAT_ACC_COLS);
    int simg = -1;
    if (ACC) {
      for (int c = lane; c < 2 * STAT_ACC_COLS; c += 32) sacc[c] = 0;
      __syncwarp();
    }
    for (int li = 0;; ++li, ++local) {
      const int item = item_at<ACC>(p, li, unit, nunits, Krun);
      if (item < 0) break;
      const Item w = decode_item(p, item, BN, CG, rank);
      const bool acc_item = ACC && p.phase_cols > 0 && p.phase_cols <= STAT_ACC_COLS;
      if (ACC && simg >= 0 && w.img != simg) {
        stat_flush(sacc, p.stats, p.stat_C, p.stat_coff, simg, p.phase_cols, lane);
        simg = -1;
      }
      const uint32_t acc = local & 1u;
      mbar_wait(bars + 128 + 8 * acc, (local >> 1) & 1u);
      tc_fence_after();
      const int ox = w.tx * p.TW + xx0, oy = w.ty * p.TH + yy;
      stat_t* strow = (p.stats && !AP_DBG(p.dbg & 4)) ? p.stats + ((size_t)w.img * p.stat_C + p.stat_coff + lane) * 2 : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < w.bn; c0 += 32, ++blk) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)row0 << 16) + acc * 256u + (uint32_t)c0, v);
        const uint32_t sl = (blk % EPI_SLABS) * 4096;
        // phase-packed transposed conv: columns [ph * phase_cols, (ph+1) * phase_cols) are output phase ph
        int ch = w.n0 + c0, phs = 0;
        if (p.phase_cols) { phs = ch / p.phase_cols; ch -= phs * p.phase_cols; }
        if (!AP_DBG(p.dbg & 2))
          epi_store_block<EPI_SLABS - 1>(v, slab_gen + sl, slab_s + sl, lane, &p.tmO[phs], p.out_coff + ch, ox, oy, w.img, opol);
        if (strow != nullptr) {
          float cs, cq;
          slab_colsums(slab_gen + sl, lane, &cs, &cq);
          if (acc_item) {  // the phases of a packed transposed conv fold onto the same channel
            sacc[ch + lane] += stat_fix(cs);
            sacc[STAT_ACC_COLS + ch + lane] += stat_fix(cq);
            simg = w.img;
          } else {
            stat_add(strow + (size_t)ch * 2, stat_fix(cs));
            stat_add(strow + (size_t)ch * 2 + 1, stat_fix(cq));
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty0 + 8 * acc);
        else mbar_arrive(bars + 144 + 8 * acc);
      }
    }
    if (ACC && simg >= 0) stat_flush(sacc, p.stats, p.stat_C, p.stat_coff, simg, p.phase_cols, lane);
    if (lane == 0) bulk_wait<0>();  // all output boxes have landed before the CTA exits
    __syncwarp();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();  // the peer's shared memory / barriers stay valid until both CTAs are done
  else __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    if (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ------------------------------------------------------------------------------------------------
// host side
// ------------------------------------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled_v12000 g_encode = nullptr;
static int g_sms = 0;

template <int BN, int NPROD, bool PACKED>
constexpr size_t pair_smem() {
  return UmmaCfg<BN, NPROD, 2>::SMEM - (PACKED ? (size_t)(EpiCfg<BN>::BYTES - 4 * 4096) : 0) + (PACKED ? STAT_ACC_BYTES : 0);
}

static int g_dbg = 0;
static int g_l2_hints = 0;
static int g_pair = 1;  // CTA-pair (cta_group::2) kernels unless AP_NETG_CTA_PAIR=0

template <int BN, int NPROD>
static int set_attr() {
  AP_CUDA(cudaFuncSetAttribute(conv_umma1_kernel<BN, NPROD>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)UmmaCfg<BN, NPROD, 1>::SMEM));
  AP_CUDA(cudaFuncSetAttribute(conv_umma_kernel<BN, NPROD, 2, false>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                               (int)pair_smem<BN, NPROD, false>()));
  if (BN == 256)
    AP_CUDA(cudaFuncSetAttribute(conv_umma_kernel<256, NPROD, 2, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                 (int)pair_smem<256, NPROD, true>()));
  return AP_OK;
}

// Function attributes (the opt-in to > 48 KB of dynamic shared memory) belong to a device, not to the process: they are
// set once for EVERY device a handle or a debug call touches (ap_netg_create(device), ap_conv2d_debug(device)).
constexpr int AP_MAX_DEVICES = 64;
static std::mutex g_init_mu;
static bool g_dev_ready[AP_MAX_DEVICES];
static int g_dev_pairs[AP_MAX_DEVICES][3][2];  // resident CTA pairs per device, [BN 64/128/256][nprod 1/3]

template <int BN, int NPROD>
static int count_pairs() {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)(g_sms > 0 ? g_sms : 148) & ~1u, 1, 1);
  cfg.blockDim = dim3(192, 1, 1);
  cfg.dynamicSmemBytes = UmmaCfg<BN, NPROD, 2>::SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_umma_kernel<BN, NPROD, 2, false>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 0;
  }
  return n;
}

int umma_init() {
  std::lock_guard<std::mutex> lock(g_init_mu);
  int dev = 0;
  AP_CUDA(cudaGetDevice(&dev));
  AP_REQUIRE(dev >= 0 && dev < AP_MAX_DEVICES, AP_ERR_UNSUPPORTED, "device ordinal %d", dev);
  if (!g_encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    AP_CUDA(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres));
    AP_REQUIRE(fn != nullptr && qres == cudaDriverEntryPointSuccess, AP_ERR_CUDA,
               "cuTensorMapEncodeTiled not available from the driver");
    AP_CUDA(cudaDeviceGetAttribute(&g_sms, cudaDevAttrMultiProcessorCount, dev));
    const char* e = getenv("AP_NETG_CTA_PAIR");
    g_pair = e ? atoi(e) : 1;  // 0: never, 1: where it wins (Cout = 256), 2: everywhere
    const char* lh = getenv("AP_NETG_L2_HINTS");
    g_l2_hints = lh ? atoi(lh) : 0;
    const char* d = getenv("AP_UMMA_DBG");
    g_dbg = d ? atoi(d) : 0;
    g_encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  if (g_dev_ready[dev]) return AP_OK;
  AP_TRY((set_attr<64, 1>()));
  AP_TRY((set_attr<128, 1>()));
  AP_TRY((set_attr<256, 1>()));
  AP_TRY((set_attr<64, 3>()));
  AP_TRY((set_attr<128, 3>()));
  AP_TRY((set_attr<256, 3>()));
  g_dev_pairs[dev][0][0] = count_pairs<64, 1>();  g_dev_pairs[dev][0][1] = count_pairs<64, 3>();
  g_dev_pairs[dev][1][0] = count_pairs<128, 1>(); g_dev_pairs[dev][1][1] = count_pairs<128, 3>();
  g_dev_pairs[dev][2][0] = count_pairs<256, 1>(); g_dev_pairs[dev][2][1] = count_pairs<256, 3>();
  AP_TRY(stem_umma_init_device());
  AP_TRY(out_umma_init_device());
  AP_TRY(halo_init_device());
  g_dev_ready[dev] = true;
  return AP_OK;
}

int umma_num_sms() { return g_sms; }

// dtype: 0 = bf16, 1 = fp32.  All maps use SWIZZLE_128B (inner box = 128 bytes).
int tmap_encode(CUtensorMap* m, int dtype, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box, const uint32_t* estr) {
  AP_TRY(umma_init());
  cuuint64_t d[5], s[5];
  cuuint32_t b[5], e[5];
  for (int i = 0; i < rank; ++i) { d[i] = dims[i]; b[i] = box[i]; e[i] = estr[i]; }
  for (int i = 0; i + 1 < rank; ++i) s[i] = strides[i];
  CUresult r = g_encode(m, dtype == 0 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32, (cuuint32_t)rank,
                        const_cast<void*>(base), d, s, b, e, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  AP_REQUIRE(r == CUDA_SUCCESS, AP_ERR_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d (rank %d, dtype %d)", (int)r, rank,
             dtype);
  return AP_OK;
}

// fp32 NHWC output view for the epilogue's TMA stores: box {32 ch, 32 px, 1 row, 1 image}; `os` > 1 selects
// the (py, px) phase of a transposed conv as a strided view.
int tmap_encode_out(CUtensorMap* m, float* out, int B, int Hout, int Wout, int C, int os, int py, int px) {
  const uint64_t dims[4] = {(uint64_t)C, (uint64_t)(Wout / os), (uint64_t)(Hout / os), (uint64_t)B};
  const uint64_t str[3] = {(uint64_t)os * C * 4, (uint64_t)os * Wout * C * 4, (uint64_t)Hout * Wout * C * 4};
  const uint32_t box[4] = {32, 32, 1, 1};
  const uint32_t es[4] = {1, 1, 1, 1};
  return tmap_encode(m, 1, out + ((size_t)py * Wout + px) * C, 4, dims, str, box, es);
}

// how many CTA pairs of this kernel can be resident at once on the current device (74 on a full B200; fewer if a TPC
// is fused off); filled in by umma_init
static int max_pairs(int BN, int nprod) {
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= AP_MAX_DEVICES || !g_dev_ready[dev]) return 0;
  return g_dev_pairs[dev][BN == 64 ? 0 : (BN == 128 ? 1 : 2)][nprod == 3 ? 1 : 0];
}

// `in` must be a bf16 activation. Zero-padded convs address the un-haloed interior (TMA fills
// out-of-bounds with zeros); reflect-padded ones address the haloed buffer (halo = in.pad >= conv pad).
static int conv_tiling(const ConvGeom& g, int* TW, int* TH) {
  if (g.Wv % 128 == 0) { *TW = 128; *TH = 1; }
  else if (g.Wv == 64 && g.Hv % 2 == 0) { *TW = 64; *TH = 2; }
  else { set_error("umma conv: virtual grid %dx%d not tileable", g.Hv, g.Wv); return AP_ERR_UNSUPPORTED; }
  return AP_OK;
}

int umma_conv_create(UmmaConv** out, const ConvGeom& g, const Act& in, int in_coff, const __nv_bfloat16* w_hi,
                     const __nv_bfloat16* w_lo, int nprod, float* out_raw, int out_C, int out_coff, stat_t* stats,
                     int stat_C, int stat_coff, const PhasePack* pk) {
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
  AP_TRY(conv_tiling(g, &TW, &TH));
  AP_REQUIRE(TW * g.stride <= 256, AP_ERR_UNSUPPORTED, "umma conv: TMA box too wide");

  UmmaConv* c = new UmmaConv();
  UmmaParams& p = c->p;
  c->BN = g.Cout;
  c->nprod = nprod;
  if (halo_conv_eligible(g, in, nprod, pk != nullptr)) {
    const int rc_h = halo_conv_create(&c->halo, g, in, in_coff, w_hi, w_lo, nprod, out_raw, out_C, out_coff,
                                      (g_dbg & 8) ? nullptr : stats, stat_C, stat_coff);
    if (rc_h != AP_OK) { delete c; return rc_h; }
    *out = c;
    return AP_OK;
  }
  const int ntiles = (g.Wv / TW) * (g.Hv / TH) * g.B;
  const int pairs = g_pair ? max_pairs(g.Cout, nprod) : 0;
  // measured (profiles/r01_cta_pair.md): pairs win for N = 256 with the 3-product operands (-7..-13%), lose for
  // N <= 128 (+9..+18%) and for single-product bf16 (+8%: 512-cycle stages are too short for the pair handshake)
  c->cg = (pairs > 0 && ntiles % 2 == 0 && ((g.Cout == 256 && nprod == 3) || g_pair == 2 || pk)) ? 2 : 1;
  if (pk) {
    AP_REQUIRE(c->cg == 2 && g.Cout == 256 && pk->nph * pk->cols == 256 && pk->cols % 32 == 0 && g.stride == 1, AP_ERR_UNSUPPORTED,
               "phase-packed transposed conv needs the CTA-pair kernel and N = 256");
  }
  const int wrows = g.Cout / c->cg;
  // activation maps
  const bool padded_view = g.reflect != 0;
  const int Hp = in.H + 2 * in.pad, Wp = in.W + 2 * in.pad;
  const int org = padded_view ? in.pad : 0;
  const uint64_t adims[4] = {(uint64_t)in.C, (uint64_t)(padded_view ? Wp : in.W), (uint64_t)(padded_view ? Hp : in.H),
                             (uint64_t)in.B};
  const uint64_t astr[3] = {(uint64_t)in.C * 2, (uint64_t)Wp * in.C * 2, (uint64_t)Hp * Wp * in.C * 2};
  const uint32_t abox[4] = {64, (uint32_t)(TW * g.stride), (uint32_t)(TH * g.stride), 1};
  const uint32_t aes[4] = {1, (uint32_t)g.stride, (uint32_t)g.stride, 1};
  const size_t view_off = padded_view ? 0 : ((size_t)in.pad * Wp + in.pad) * in.C;
  int rc = tmap_encode(&p.tmA[0], 0, reinterpret_cast<const __nv_bfloat16*>(in.p0) + view_off, 4, adims, astr, abox, aes);
  if (rc == AP_OK && nprod == 3)
    rc = tmap_encode(&p.tmA[1], 0, reinterpret_cast<const __nv_bfloat16*>(in.p1) + view_off, 4, adims, astr, abox, aes);
  // weight maps [slab][Cout][Cin], boxes of 64 output channels
  const int wslabs = pk ? g.taps.n : 9;
  for (int i = 0; i < g.taps.n; ++i)
    AP_REQUIRE(g.taps.slab[i] < wslabs, AP_ERR_INVALID, "umma conv: weight slab %d of %d", g.taps.slab[i], wslabs);
  const uint64_t wdims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, (uint64_t)wslabs};
  const uint64_t wstr[2] = {(uint64_t)g.Cin * 2, (uint64_t)g.Cin * g.Cout * 2};
  const uint32_t wbox[3] = {64, (uint32_t)(wrows < 64 ? wrows : 64), 1};
  const uint32_t wes[3] = {1, 1, 1};
  if (rc == AP_OK) rc = tmap_encode(&p.tmW[0], 0, w_hi, 3, wdims, wstr, wbox, wes);
  if (rc == AP_OK && nprod == 3) rc = tmap_encode(&p.tmW[1], 0, w_lo, 3, wdims, wstr, wbox, wes);
  const uint32_t wbox_s[3] = {64, 32, 1};
  if (rc == AP_OK) rc = tmap_encode(&p.tmWs[0], 0, w_hi, 3, wdims, wstr, wbox_s, wes);
  if (rc == AP_OK && nprod == 3) rc = tmap_encode(&p.tmWs[1], 0, w_lo, 3, wdims, wstr, wbox_s, wes);
  p.phase_cols = pk ? pk->cols : 0;
  if (pk) {
    for (int i = 0; i < pk->nph && rc == AP_OK; ++i)
      rc = tmap_encode_out(&p.tmO[i], out_raw, g.B, 2 * g.Hin, 2 * g.Win, out_C, 2, pk->py[i], pk->px[i]);
  } else if (rc == AP_OK) {
    rc = tmap_encode_out(&p.tmO[0], out_raw, g.B, g.Hout, g.Wout, out_C, g.os, g.py, g.px);
  }
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
  p.out_coff = out_coff;
  p.stats = (g_dbg & 8) ? nullptr : stats;  // AP_UMMA_DBG bit 3 (timing probe only): no InstanceNorm statistics at all
  p.stat_C = stat_C; p.stat_coff = stat_coff;
  // item list: whole waves of full tile groups, the remainder split along N so the tail fills the machine
  const int groups = ntiles / c->cg;
  const int G = c->cg == 2 ? pairs : (g_sms > 0 ? g_sms : 148);  // CTAs (CG = 1) or CTA pairs (CG = 2) resident at once
  const int rem = groups % G;
  int split = 1;
  if (rem > 0) {
    while (split < 4 && rem * split * 2 <= G && wrows / (split * 2) >= 32) split *= 2;  // a part is >= one 32-row box per CTA
  }
  p.split = split;
  p.dbg = g_dbg;
  p.l2_hints = g_l2_hints;
  p.n_full = groups - rem;
  p.n_items = p.n_full + rem * split;
  c->grid = dim3((unsigned)((p.n_items < G ? p.n_items : G) * c->cg), 1);
  *out = c;
  return AP_OK;
}

bool umma_pairs_available() { return umma_init() == AP_OK && g_pair != 0 && max_pairs(256, 3) > 0; }

void umma_conv_destroy(UmmaConv* c) {
  if (c && c->halo) halo_conv_destroy(c->halo);
  delete c;
}

template <int BN, int NPROD, int CG>
static int launch_one(const UmmaConv* c, cudaStream_t st) {
  if constexpr (CG == 1) {
    conv_umma1_kernel<BN, NPROD><<<c->grid, 192, UmmaCfg<BN, NPROD, 1>::SMEM, st>>>(c->p);
    AP_CUDA(cudaGetLastError());
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = c->grid;
    cfg.blockDim = dim3(192, 1, 1);
    const bool packed = c->p.phase_cols > 0 && BN == 256;
    cfg.dynamicSmemBytes = packed ? pair_smem<256, NPROD, true>() : pair_smem<BN, NPROD, false>();
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (packed) AP_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<256, NPROD, 2, true>, c->p));
    else AP_CUDA(cudaLaunchKernelEx(&cfg, conv_umma_kernel<BN, NPROD, 2, false>, c->p));
  }
  launches_add(1);
  return AP_OK;
}

int umma_conv_launch(const UmmaConv* c, cudaStream_t st) {
  if (c->halo) return halo_conv_launch(c->halo, st);
#define AP_UMMA_CASE(BN_, NP_)                                                 \
  if (c->BN == BN_ && c->nprod == NP_)                                         \
    return c->cg == 2 ? launch_one<BN_, NP_, 2>(c, st) : launch_one<BN_, NP_, 1>(c, st);
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
