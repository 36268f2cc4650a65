// Halo-staged tcgen05 convolution for the 3x3 stride-1 convs of the 64x64 trunk (ResnetBlock / ResnetBlock2 / merge,
// Module2/models/networks.py:1251, 2341-2414), N = 256.
//
// The tap-shifted kernel of conv_umma.cu fetches every activation element from L2 once per tap: 9 x 16 KB per 128-pixel
// tile and 64-channel chunk.  That traffic is unique per CTA (nothing to de-duplicate) and at tensor speed it alone
// asks for 64 B/clk per SM -- more than the L2 -> SM fabric gives all 148 SMs at once (~42 B/clk per SM), which is
// exactly where the single-product bf16 trunk sat (0.62-0.72 of the tensor peak) and what bounds batch size 1.
// Here a tile is 16 rows x 8 pixels and its haloed patch (18 rows x 10 pixels x 64 channels) is staged ONCE per channel
// chunk by one TMA box {64 ch, 16 px, 18 rows}; the nine taps are nine shared-memory matrix descriptors into the same
// patch: tap (ky, kx) starts (ky * 16 + kx) rows of 128 bytes further on.  An accumulator row m = (y, x) = (m / 8, m % 8)
// reads patch row (y + ky) * 16 + x + kx: eight consecutive pixels of one image row are one 8-row core-matrix group,
// groups are 16 rows = 2048 bytes apart (stride-byte-offset).  The 128-byte swizzle is a function of the ABSOLUTE
// shared-memory address bits (measured on B200: a descriptor that starts kx rows into an atom reads what TMA wrote there
// with base_offset = 0; setting base_offset to the row phase gives wrong results), so any 128-byte-aligned start works
// as long as the group stride is a multiple of the 1024-byte atom -- hence the 16-pixel pitch of the box.  A traffic
// per tile drops from 9 x 16 KB to 36 KB per chunk (4x).
//
// Two rings: A patches (one per channel chunk, consumed by 9 taps) and W tiles (one per tap and chunk).  Warp roles:
// warp 0 = W producer, warp 1 = MMA issuer, warps 2..5 = epilogue (as conv_umma.cu: TMEM -> swizzled slab -> TMA store,
// fixed-point InstanceNorm statistics), warp 6 = A producer (its own thread so that a patch is requested as soon as its
// slot is free, a whole chunk ahead, whatever the W ring is waiting for).  CG = 2: CTA pair, cta_group::2, M = 256.
#include <cudaTypedefs.h>
#include <stdlib.h>

#include "common.cuh"
#include "umma.cuh"

namespace ap {

void launches_add(int n);

constexpr int H_TILE_W = 8, H_TILE_H = 16;          // M = 128 = 16 rows x 8 pixels
constexpr int H_BOX_W = 16, H_BOX_H = 18;           // patch box: pitch 16 pixels (two swizzle atoms), 18 rows
constexpr int H_A_PLANE = H_BOX_W * H_BOX_H * 128;  // 36864 bytes per plane per 64-channel chunk

struct alignas(64) HaloParams {
  CUtensorMap tmA[2];   // hi, lo: box {64 ch, 16 px, 18 rows, 1 image}
  CUtensorMap tmW[2];   // box {64 ch, 64 rows, 1 slab}
  CUtensorMap tmWs[2];  // box {64 ch, 32 rows, 1 slab}: N-split items narrower than a 64-row box
  CUtensorMap tmO;      // fp32 output: box {32 ch, 8 px, 4 rows, 1 image}
  int ntaps;
  int8_t ty[9], tx[9];  // tap offset inside the patch (0..2)
  uint8_t slab[9];
  int kchunks, last_ksteps, cin_off;
  int org;              // patch origin relative to the tile origin: 0 (haloed view: reflection) or -1 (zero padding)
  int tiles_x, tiles_y;
  int out_coff;
  stat_t* stats;
  int stat_C, stat_coff;
  int n_full, split, n_items;
  // W ring: `nw` slots of `wslot` bytes.  Sized per launch from the widest item: when every item is an N-split part
  // (small batches) the same shared memory holds up to 8 short stages instead of 2-4 long ones -- at batch size 1 the
  // conv is bound by the latency of its 36 (tap, chunk) stages, not by bytes
  int nw, wslot;
};

template <int NPROD, int CG>
struct HaloCfg {
  static constexpr int PLANES = NPROD == 3 ? 2 : 1;
  static constexpr int NA = (NPROD == 1 && CG == 2) ? 3 : 2;                    // A patches in flight
  static constexpr int NW = NPROD == 3 ? 2 : (CG == 2 ? 4 : 3);                 // W stages of full-width items
  static constexpr int NW_MAX = 8;
  static constexpr int W_ROWS = 256 / CG;
  static constexpr int W_TILE = W_ROWS * 128;
  static constexpr int A_STAGE = PLANES * H_A_PLANE;
  static constexpr int W_STAGE = PLANES * W_TILE;
  static constexpr int EPI_SLABS = NPROD == 3 ? 1 : 2;
  static constexpr int EPI_BYTES = 4 * EPI_SLABS * 4096;
  static constexpr size_t SMEM = 1024 + (size_t)NA * A_STAGE + (size_t)NW * W_STAGE + EPI_BYTES + 256;
};

struct HItem {
  int img, ty, tx, n0, bn;
};

__device__ __forceinline__ HItem halo_item(const HaloParams& p, int item, int CG, int rank) {
  int g, n0 = 0, bn = 256;
  if (item < p.n_full) {
    g = item;
  } else {
    const int r = item - p.n_full;
    g = p.n_full + r / p.split;
    bn = 256 / p.split;
    n0 = (r % p.split) * bn;
  }
  int m = g * CG + rank;
  HItem it;
  it.tx = m % p.tiles_x; m /= p.tiles_x;
  it.ty = m % p.tiles_y; m /= p.tiles_y;
  it.img = m;
  it.n0 = n0;
  it.bn = bn;
  return it;
}

// K-major SWIZZLE_128B descriptor of a patch view: 8-row groups 2048 bytes apart; the start may sit on any 128-byte row
// (base_offset stays 0: see the header)
__device__ __forceinline__ uint64_t make_halo_desc(uint32_t saddr) {
  uint64_t d = 0;
  d |= (uint64_t)((saddr & 0x3FFFFu) >> 4);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)((uint32_t)(H_BOX_W * 128) >> 4) << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

template <int NPROD, int CG>
__global__ void __launch_bounds__(224, 1) conv_halo_kernel(const __grid_constant__ HaloParams p) {
  using Cfg = HaloCfg<NPROD, CG>;
  constexpr int PLANES = Cfg::PLANES, NA = Cfg::NA;
  const int NW = p.nw;
  extern __shared__ uint8_t smem_raw[];
  const uint32_t smem_base = (smem_u32(smem_raw) + 1023u) & ~1023u;
  uint8_t* smem_gen = smem_raw + (smem_base - smem_u32(smem_raw));
  const uint32_t sA = smem_base;
  const uint32_t sW = smem_base + NA * Cfg::A_STAGE;
  const uint32_t epi_s = sW + Cfg::NW * Cfg::W_STAGE;
  uint8_t* epi_gen = smem_gen + NA * Cfg::A_STAGE + Cfg::NW * Cfg::W_STAGE;
  constexpr int EPI_SLABS = Cfg::EPI_SLABS;
  const uint32_t bars = epi_s + Cfg::EPI_BYTES;
  // barriers: a_full +0, a_empty +24, w_full +48, w_empty +112, tfull +176, tempty +192, tmem pointer +208
  const uint32_t b_afull = bars, b_aempty = bars + 24, b_wfull = bars + 48, b_wempty = bars + 112, b_tfull = bars + 176,
                 b_tempty = bars + 192;
  uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(epi_gen + Cfg::EPI_BYTES + 208);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int rank = (CG == 2) ? (int)cluster_ctarank() : 0;
  const int unit = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int nunits = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmA[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmW[0]) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&p.tmO) : "memory");
    for (int s = 0; s < NA; ++s) {
      mbar_init(b_afull + 8 * s, 1);
      mbar_init(b_aempty + 8 * s, 1);
    }
    for (int s = 0; s < NW; ++s) {
      mbar_init(b_wfull + 8 * s, 1);
      mbar_init(b_wempty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(b_tfull + 8 * a, 1);
      mbar_init(b_tempty + 8 * a, 4 * CG);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(tmem_ptr_smem)), "r"(512u)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr_smem;

  if (warp == 6) {
    // ===================== A producer: one haloed patch per (item, channel chunk) =====================
    if (lane == 0) {
      const uint32_t afull0 = (CG == 2) ? mapa_rank(b_afull, 0) : b_afull;
      uint32_t cnt = 0;
      for (int li = 0;; ++li) {
        const int item = unit + li * nunits;
        if (item >= p.n_items) break;
        const HItem w = halo_item(p, item, CG, rank);
        const int x0 = w.tx * H_TILE_W + p.org, y0 = w.ty * H_TILE_H + p.org;
        for (int ch = 0; ch < p.kchunks; ++ch, ++cnt) {
          const uint32_t s = cnt % NA, ph = (cnt / NA) & 1u;
          mbar_wait(b_aempty + 8 * s, ph ^ 1u);
          if (rank == 0) mbar_expect_tx(b_afull + 8 * s, (uint32_t)(CG * Cfg::A_STAGE));
          const uint32_t dst = sA + s * Cfg::A_STAGE;
#pragma unroll
          for (int pl = 0; pl < PLANES; ++pl) {
            if (CG == 2) tma_load_4d_pair(dst + pl * H_A_PLANE, &p.tmA[pl], afull0 + 8 * s, p.cin_off + ch * 64, x0, y0, w.img);
            else tma_load_4d(dst + pl * H_A_PLANE, &p.tmA[pl], afull0 + 8 * s, p.cin_off + ch * 64, x0, y0, w.img);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 0) {
    // ===================== W producer: one weight tile per (item, channel chunk, tap) =====================
    if (lane == 0) {
      const uint32_t wfull0 = (CG == 2) ? mapa_rank(b_wfull, 0) : b_wfull;
      uint32_t cnt = 0;
      for (int li = 0;; ++li) {
        const int item = unit + li * nunits;
        if (item >= p.n_items) break;
        const HItem w = halo_item(p, item, CG, rank);
        const int wrows = w.bn / CG;
        const bool small = wrows < 64;
        const int nbox = small ? 1 : wrows / 64;
        const CUtensorMap* mW = small ? p.tmWs : p.tmW;
        const int wrow0 = w.n0 + rank * wrows;
        const uint32_t tx_bytes = (uint32_t)(PLANES * wrows * 128);
        for (int ch = 0; ch < p.kchunks; ++ch) {
          for (int t = 0; t < p.ntaps; ++t, ++cnt) {
            const uint32_t s = cnt % NW, ph = (cnt / NW) & 1u;
            mbar_wait(b_wempty + 8 * s, ph ^ 1u);
            if (rank == 0) mbar_expect_tx(b_wfull + 8 * s, CG * tx_bytes);
            const uint32_t dst = sW + s * p.wslot;
            for (int b = 0; b < nbox; ++b) {
#pragma unroll
              for (int pl = 0; pl < PLANES; ++pl) {
                const uint32_t d = dst + pl * (p.wslot / PLANES) + b * 8192;
                if (CG == 2) tma_load_3d_pair(d, &mW[pl], wfull0 + 8 * s, ch * 64, wrow0 + 64 * b, p.slab[t]);
                else tma_load_3d(d, &mW[pl], wfull0 + 8 * s, ch * 64, wrow0 + 64 * b, p.slab[t]);
              }
            }
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA of the pair only) =====================
    if (lane == 0 && rank == 0) {
      uint32_t acnt = 0, wcnt = 0, local = 0;
      for (int li = 0;; ++li, ++local) {
        const int item = unit + li * nunits;
        if (item >= p.n_items) break;
        const HItem w = halo_item(p, item, CG, 0);
        const uint32_t idesc =
            (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(w.bn >> 3) << 17) | ((uint32_t)((128 * CG) >> 4) << 24);
        const uint32_t acc = local & 1u;
        mbar_wait(b_tempty + 8 * acc, ((local >> 1) & 1u) ^ 1u);
        tc_fence_after();
        const uint32_t d = tmem_base + acc * 256u;
        uint32_t first = 0;
        for (int ch = 0; ch < p.kchunks; ++ch, ++acnt) {
          const uint32_t as = acnt % NA, aph = (acnt / NA) & 1u;
          mbar_wait(b_afull + 8 * as, aph);
          tc_fence_after();
          const int ksteps = (ch == p.kchunks - 1) ? p.last_ksteps : 4;
          const uint32_t pa = sA + as * Cfg::A_STAGE;
          for (int t = 0; t < p.ntaps; ++t, ++wcnt) {
            const uint32_t ws = wcnt % NW, wph = (wcnt / NW) & 1u;
            mbar_wait(b_wfull + 8 * ws, wph);
            tc_fence_after();
            const uint32_t va = pa + (uint32_t)((p.ty[t] * H_BOX_W + p.tx[t]) * 128);
            const uint32_t vw = sW + ws * p.wslot;
            const uint64_t a_hi = make_halo_desc(va);
            const uint64_t w_hi = make_sw128_desc(vw);
            if (NPROD == 3) {
              const uint64_t a_lo = make_halo_desc(va + H_A_PLANE);
              const uint64_t w_lo = make_sw128_desc(vw + p.wslot / PLANES);
              for (int k = 0; k < ksteps; ++k) {
                const uint64_t o = (uint64_t)(k * 2);
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
            if (CG == 2) umma_commit_pair(b_wempty + 8 * ws);
            else umma_commit(b_wempty + 8 * ws);
          }
          // the patch is free once the MMAs of its last tap have retired
          if (CG == 2) umma_commit_pair(b_aempty + 8 * as);
          else umma_commit(b_aempty + 8 * as);
        }
        if (CG == 2) umma_commit_pair(b_tfull + 8 * acc);
        else umma_commit(b_tfull + 8 * acc);
      }
    }
    __syncwarp();
  } else {
    // ===================== epilogue (warps 2..5): 4 image rows x 8 pixels per warp =====================
    const int q = warp & 3;
    const int row0 = q * 32;
    uint8_t* slab_gen = epi_gen + q * (EPI_SLABS * 4096);
    const uint32_t slab_s = epi_s + q * (EPI_SLABS * 4096);
    const uint32_t tempty0 = (CG == 2) ? mapa_rank(b_tempty, 0) : b_tempty;
    uint32_t local = 0, blk = 0;
    for (int li = 0;; ++li, ++local) {
      const int item = unit + li * nunits;
      if (item >= p.n_items) break;
      const HItem w = halo_item(p, item, CG, rank);
      const uint32_t acc = local & 1u;
      mbar_wait(b_tfull + 8 * acc, (local >> 1) & 1u);
      tc_fence_after();
      const int ox = w.tx * H_TILE_W, oy = w.ty * H_TILE_H + 4 * q;
      stat_t* strow = p.stats ? p.stats + ((size_t)w.img * p.stat_C + p.stat_coff + w.n0 + lane) * 2 : nullptr;
#pragma unroll 1
      for (int c0 = 0; c0 < w.bn; c0 += 32, ++blk) {
        float v[32];
        tmem_ld32(tmem_base + ((uint32_t)row0 << 16) + acc * 256u + (uint32_t)c0, v);
        const uint32_t sl = (blk % EPI_SLABS) * 4096;
        epi_store_block<EPI_SLABS - 1>(v, slab_gen + sl, slab_s + sl, lane, &p.tmO, p.out_coff + w.n0 + c0, ox, oy, w.img);
        if (strow != nullptr) {
          float cs, cq;
          slab_colsums(slab_gen + sl, lane, &cs, &cq);
          stat_add(strow + (size_t)c0 * 2, stat_fix(cs));
          stat_add(strow + (size_t)c0 * 2 + 1, stat_fix(cq));
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (CG == 2) mbar_arrive_cluster(tempty0 + 8 * acc);
        else mbar_arrive(b_tempty + 8 * acc);
      }
    }
    if (lane == 0) bulk_wait<0>();
    __syncwarp();
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all();
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
struct HaloConv {
  HaloParams p;
  int nprod, cg;
  dim3 grid;
};

static int g_halo_pairs[64][2];  // resident CTA pairs per device, [nprod 1/3]
static bool g_halo_ready[64];

template <int NPROD>
static int halo_count_pairs(int sms) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3((unsigned)sms & ~1u, 1, 1);
  cfg.blockDim = dim3(224, 1, 1);
  cfg.dynamicSmemBytes = HaloCfg<NPROD, 2>::SMEM;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int n = 0;
  if (cudaOccupancyMaxActiveClusters(&n, conv_halo_kernel<NPROD, 2>, &cfg) != cudaSuccess || n <= 0) {
    cudaGetLastError();
    n = 0;
  }
  return n;
}

int halo_init_device() {  // called by umma_init for every device it sees (under its lock)
  int dev = 0;
  AP_CUDA(cudaGetDevice(&dev));
  if (dev < 0 || dev >= 64 || g_halo_ready[dev]) return AP_OK;
  AP_CUDA(cudaFuncSetAttribute(conv_halo_kernel<1, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HaloCfg<1, 1>::SMEM));
  AP_CUDA(cudaFuncSetAttribute(conv_halo_kernel<1, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HaloCfg<1, 2>::SMEM));
  AP_CUDA(cudaFuncSetAttribute(conv_halo_kernel<3, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)HaloCfg<3, 2>::SMEM));
  const int sms = umma_num_sms() > 0 ? umma_num_sms() : 148;
  g_halo_pairs[dev][0] = halo_count_pairs<1>(sms);
  g_halo_pairs[dev][1] = halo_count_pairs<3>(sms);
  g_halo_ready[dev] = true;
  return AP_OK;
}

// AP_NETG_HALO: 0 = tap-shifted kernel everywhere; bit 0: halo kernel for the single-product (bf16) trunk; bit 1: for the
// 3-product trunk as well (tensor-bound: measured 3 % slower than the tap-shifted pair kernel at B = 16 and no faster at
// B = 1, where 432 narrow MMAs per item bound the conv, not bytes).  Default 1.  The choice never depends on the batch
// size: a frame must come out bit-identical whether it is rendered alone or in a batch.
static int halo_mode() {
  static int v = -1;
  if (v < 0) { const char* e = getenv("AP_NETG_HALO"); v = e ? atoi(e) : 1; }
  return v;
}
static int halo_cg_pref() {  // AP_HALO_CG: 1 or 2 for the single-product kernel (A/B); the 3-product kernel needs pairs
  static int v = -1;
  if (v < 0) { const char* e = getenv("AP_HALO_CG"); v = e ? atoi(e) : 2; }
  return v;
}

bool halo_conv_eligible(const ConvGeom& g, const Act& in, int nprod, bool packed) {
  if (packed) return false;
  if (!(halo_mode() & (nprod == 3 ? 2 : 1))) return false;
  if (g.stride != 1 || g.os != 1 || g.Hv != 64 || g.Wv != 64 || g.Cout != 256 || g.taps.n != 9) return false;
  if (in.H != 64 || in.W != 64) return false;
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || !g_halo_ready[dev]) return false;
  if (nprod == 3) return g_halo_pairs[dev][1] > 0;
  return true;
}

int halo_conv_create(HaloConv** out, const ConvGeom& g, const Act& in, int in_coff, const __nv_bfloat16* w_hi,
                     const __nv_bfloat16* w_lo, int nprod, float* out_raw, int out_C, int out_coff, stat_t* stats,
                     int stat_C, int stat_coff) {
  int dev = 0;
  AP_CUDA(cudaGetDevice(&dev));
  HaloConv* c = new HaloConv();
  HaloParams& p = c->p;
  c->nprod = nprod;
  const int pairs = g_halo_pairs[dev][nprod == 3 ? 1 : 0];
  c->cg = (nprod == 3 || (halo_cg_pref() == 2 && pairs > 0)) ? 2 : 1;
  const bool padded_view = g.reflect != 0;
  const int Hp = in.H + 2 * in.pad, Wp = in.W + 2 * in.pad;
  const uint64_t adims[4] = {(uint64_t)in.C, (uint64_t)(padded_view ? Wp : in.W), (uint64_t)(padded_view ? Hp : in.H), (uint64_t)in.B};
  const uint64_t astr[3] = {(uint64_t)in.C * 2, (uint64_t)Wp * in.C * 2, (uint64_t)Hp * Wp * in.C * 2};
  const uint32_t abox[4] = {64, H_BOX_W, H_BOX_H, 1};
  const uint32_t ones4[4] = {1, 1, 1, 1};
  const size_t view_off = padded_view ? 0 : ((size_t)in.pad * Wp + in.pad) * in.C;
  int rc = tmap_encode(&p.tmA[0], 0, reinterpret_cast<const __nv_bfloat16*>(in.p0) + view_off, 4, adims, astr, abox, ones4);
  if (rc == AP_OK && nprod == 3)
    rc = tmap_encode(&p.tmA[1], 0, reinterpret_cast<const __nv_bfloat16*>(in.p1) + view_off, 4, adims, astr, abox, ones4);
  const uint64_t wdims[3] = {(uint64_t)g.Cin, (uint64_t)g.Cout, 9};
  const uint64_t wstr[2] = {(uint64_t)g.Cin * 2, (uint64_t)g.Cin * g.Cout * 2};
  const uint32_t wbox[3] = {64, 64, 1}, wbox_s[3] = {64, 32, 1}, ones3[3] = {1, 1, 1};
  if (rc == AP_OK) rc = tmap_encode(&p.tmW[0], 0, w_hi, 3, wdims, wstr, wbox, ones3);
  if (rc == AP_OK && nprod == 3) rc = tmap_encode(&p.tmW[1], 0, w_lo, 3, wdims, wstr, wbox, ones3);
  if (rc == AP_OK) rc = tmap_encode(&p.tmWs[0], 0, w_hi, 3, wdims, wstr, wbox_s, ones3);
  if (rc == AP_OK && nprod == 3) rc = tmap_encode(&p.tmWs[1], 0, w_lo, 3, wdims, wstr, wbox_s, ones3);
  if (rc == AP_OK) {
    const uint64_t odims[4] = {(uint64_t)out_C, 64, 64, (uint64_t)g.B};
    const uint64_t ostr[3] = {(uint64_t)out_C * 4, (uint64_t)64 * out_C * 4, (uint64_t)64 * 64 * out_C * 4};
    const uint32_t obox[4] = {32, H_TILE_W, 4, 1};
    rc = tmap_encode(&p.tmO, 1, out_raw, 4, odims, ostr, obox, ones4);
  }
  if (rc != AP_OK) { delete c; return rc; }
  p.ntaps = 9;
  for (int i = 0; i < 9; ++i) {
    // g.taps.dy/dx are offsets from the output pixel in the un-haloed input (-1..1); the patch starts one pixel up-left
    p.ty[i] = (int8_t)(g.taps.dy[i] + 1);
    p.tx[i] = (int8_t)(g.taps.dx[i] + 1);
    p.slab[i] = g.taps.slab[i];
  }
  p.kchunks = (g.Cin + 63) / 64;
  p.last_ksteps = (g.Cin - (p.kchunks - 1) * 64 + 15) / 16;
  p.cin_off = in_coff;
  p.org = padded_view ? (in.pad - 1) : -1;
  p.tiles_x = 64 / H_TILE_W;
  p.tiles_y = 64 / H_TILE_H;
  p.out_coff = out_coff;
  p.stats = stats; p.stat_C = stat_C; p.stat_coff = stat_coff;
  const int ntiles = p.tiles_x * p.tiles_y * g.B;
  if (c->cg == 2 && (ntiles % 2 != 0)) c->cg = 1;
  AP_REQUIRE(!(nprod == 3 && c->cg != 2), AP_ERR_UNSUPPORTED, "halo conv: the 3-product kernel needs CTA pairs");
  const int groups = ntiles / c->cg;
  const int sms = umma_num_sms() > 0 ? umma_num_sms() : 148;
  const int G = c->cg == 2 ? pairs : sms;
  const int rem = groups % G;
  const int wrows = 256 / c->cg;
  int split = 1;
  if (rem > 0) {
    while (split < 4 && rem * split * 2 <= G && wrows / (split * 2) >= 32) split *= 2;
  }
  p.split = split;
  p.n_full = groups - rem;
  p.n_items = p.n_full + rem * split;
  {
    const int planes = nprod == 3 ? 2 : 1;
    const int nw_full = nprod == 3 ? 2 : (c->cg == 2 ? 4 : 3);
    const int ring = nw_full * planes * wrows * 128;
    const int widest = p.n_full > 0 ? wrows : wrows / split;       // W rows per CTA of the widest item of this launch
    p.wslot = planes * widest * 128;                               // multiple of 1024 (widest >= 32 rows, planes x 4 KB)
    p.nw = ring / p.wslot < 8 ? ring / p.wslot : 8;
  }
  c->grid = dim3((unsigned)((p.n_items < G ? p.n_items : G) * c->cg), 1);
  *out = c;
  return AP_OK;
}

void halo_conv_destroy(HaloConv* c) { delete c; }

int halo_conv_launch(const HaloConv* c, cudaStream_t st) {
  if (c->cg == 1) {
    AP_REQUIRE(c->nprod == 1, AP_ERR_UNSUPPORTED, "halo conv: single-CTA kernel is single-product only");
    conv_halo_kernel<1, 1><<<c->grid, 224, HaloCfg<1, 1>::SMEM, st>>>(c->p);
    AP_CUDA(cudaGetLastError());
  } else {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = c->grid;
    cfg.blockDim = dim3(224, 1, 1);
    cfg.dynamicSmemBytes = c->nprod == 3 ? HaloCfg<3, 2>::SMEM : HaloCfg<1, 2>::SMEM;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = 2;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    if (c->nprod == 3) AP_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<3, 2>, c->p));
    else AP_CUDA(cudaLaunchKernelEx(&cfg, conv_halo_kernel<1, 2>, c->p));
  }
  launches_add(1);
  return AP_OK;
}

}  // namespace ap
