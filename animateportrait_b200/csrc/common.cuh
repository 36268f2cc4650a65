// Shared declarations of libapnetg (B200 / sm_100a only).
#pragma once
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/ap_netg.h"

namespace ap {

void set_error(const char* fmt, ...);

#define AP_CUDA(call)                                                                      \
  do {                                                                                     \
    cudaError_t _e = (call);                                                               \
    if (_e != cudaSuccess) {                                                               \
      ap::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(_e)); \
      return AP_ERR_CUDA;                                                                  \
    }                                                                                      \
  } while (0)

#define AP_TRY(call)            \
  do {                          \
    int _r = (call);            \
    if (_r != AP_OK) return _r; \
  } while (0)

#define AP_REQUIRE(cond, code, ...) \
  do {                              \
    if (!(cond)) {                  \
      ap::set_error(__VA_ARGS__);   \
      return (code);                \
    }                               \
  } while (0)

// Activation formats in HBM. All activations are NHWC; `pad` is a halo of that many pixels on each
// side of H and W that the PRODUCER fills (reflection) or leaves zero.
enum ActFmt : int {
  FMT_F32 = 0,     // one fp32 plane
  FMT_BF16X2 = 1,  // two bf16 planes: hi = bf16(x), lo = bf16(x - hi)   (fp32-accurate tcgen05 operands)
  FMT_BF16 = 2     // one bf16 plane
};

struct Act {
  int B = 0, H = 0, W = 0, C = 0;  // C = channels of the whole buffer (concats are channel offsets)
  int pad = 0;
  int fmt = FMT_F32;
  void* p0 = nullptr;  // fp32 plane or bf16 hi plane
  void* p1 = nullptr;  // bf16 lo plane (FMT_BF16X2)
  size_t pixels() const { return (size_t)B * (H + 2 * pad) * (W + 2 * pad); }
  size_t elems() const { return pixels() * C; }
};

// InstanceNorm statistics of a producer kernel, deterministic and independent of the batch size: every producer warp
// (or CTA) stores the {sum, sum of squares} of ITS pixels of an image as one partial row -- no two producers share an
// address, nothing is accumulated with atomics -- and takes a ticket per (image, 32-channel block); whoever draws the
// last ticket adds the `np` rows of the block up in row order (fp64) and writes the final [B][C][2] doubles.  Which
// pixels a row covers depends on the image geometry only, never on the batch size, the grid or the timing, so the
// statistics are bit-identical from run to run and between a batch and its frames run one by one -- like the
// reference's CPU instance_norm (Module2/models/networks.py:34).  The ticket counters return to zero by themselves.
struct StatSink {
  double* stats;    // [B][C][2] final {sum, sumsq}; null: no statistics
  float2* part;     // [B][np][C] partial rows
  uint32_t* count;  // [B][C/32] tickets
  int C, coff;      // channels of the statistics tensor, first channel this producer writes (multiple of 32)
  int np;           // partial rows per image that make a complete plane
};

// Raw (pre-InstanceNorm) conv output: fp32 NHWC, no halo, plus per-(n,c) {sum, sumsq} in double.
struct Raw {
  int B = 0, H = 0, W = 0, C = 0;
  float* p = nullptr;
  double* stats = nullptr;     // [B][C][2]
  float2* part = nullptr;      // [B][H*W/32][C] partial rows (upper bound of every producer's np)
  uint32_t* count = nullptr;   // [B][ceil(C/32)]
  StatSink sink(int np, int coff = 0) const { return StatSink{stats, part, count, C, coff, np}; }
};

// Pointers of the tensors a caller hands to one forward.  Kernels that touch caller memory read them from this table
// in device memory (one per plan, refreshed by a one-thread kernel before every forward) instead of taking them as
// launch parameters: the launch sequence of a plan is then identical from call to call and can be replayed as a CUDA
// graph whatever buffers the caller passes.
struct IoPtrs {
  const float* input;
  const float* land1;
  const float* land2;
  const float* motion;
  const float* flow;
  const float* ifmask;
  float* out;
};

constexpr int AP_MAX_TAPS = 49;
struct ConvTaps {
  int n;
  int8_t dy[AP_MAX_TAPS];  // input offset of tap t relative to  y_virtual*stride
  int8_t dx[AP_MAX_TAPS];
  uint8_t slab[AP_MAX_TAPS];  // which [ky*kw+kx] weight slab the tap multiplies
};

// One convolution "problem": a set of taps applied on a virtual output grid Hv x Wv (the input grid
// for the phases of a transposed conv), scattered to out[(y*os+py, x*os+px)].
struct ConvGeom {
  int B;
  int Hin, Win, Cin;   // logical input (without halo)
  int Hv, Wv;          // virtual output grid
  int stride;          // input step per virtual output pixel
  int reflect;         // 1: reflection padding, 0: zero padding
  int Cout;
  int os, py, px;      // output scatter
  int Hout, Wout;
  ConvTaps taps;
};

// ConvTranspose2d(k3,s2,p1,op1) as ONE stride-1 conv over the 2x2 input taps (dy,dx) in {0,1}^2 whose N dimension packs
// `nph` output phases of `cols` channels each (weights of taps a phase does not use are zero): N = 256 keeps the
// CTA-pair kernel efficient where four thin per-phase GEMMs are L2-fabric bound.  The epilogue routes column block
// ph to output pixels (2i + py[ph], 2j + px[ph]).
struct PhasePack {
  int nph, cols;
  int py[4], px[4];
};

struct SimtConvP {
  ConvGeom g;
  const float* in;  // fp32, NHWC (in_C channels per pixel, first channel in_coff) or NCHW
  int in_nchw;
  int in_C, in_coff;
  const float* wpk;  // [slab][Cin][Cout] fp32
  float* out;        // raw NHWC
  int out_C, out_coff;
  double* stats;  // may be null; indexed [(n*stat_C + stat_coff + c)*2]
  int stat_C, stat_coff;
};

struct ApplyP {
  const float* raw; int raw_C, raw_coff;
  const double* stats; int stat_C, stat_coff;  // null -> no normalisation (bias mode)
  const float* bias;                           // null or [C]
  const float* raw2; int raw2_C, raw2_coff;    // optional second InstanceNorm'ed operand (ResnetBlock2 shortcut)
  const double* stats2; int stat2_C, stat2_coff;
  const float* res_in;  // optional fp32 residual stream [B,H,W,C] added to the result
  float* res_out;       // optional: result written here as fp32 [B,H,W,C]
  // optional residual read from an ACTIVATION buffer instead (the block input itself: fp32, or bf16 hi + lo)
  int res_fmt; const void* res_p0; const void* res_p1; int res_C, res_pad;   // res_fmt = -1: absent
  int relu;
  int B, H, W, C;
  // destination (may be absent: fmt = -1)
  int fmt; void* d0; void* d1; int dC, dcoff, dpad;
  int halo_reflect;  // fill the halo ring by reflection (pad==1)
  int l2_hints;      // raw loads evict_first (set by launch_apply from AP_NETG_L2_HINTS bit 1)
  int src_shared;    // 1: `raw`/`stats` hold ONE image that is normalised into every image of the destination (clip mode)
};

struct WarpP {
  const float* raw; int raw_C, raw_coff;       // raw stem/conv output, normalised + ReLU on the fly
  const double* stats; int stat_C, stat_coff;
  const IoPtrs* io;     // motion [B,256,256,2], flow [B,2,256,256], ifmask [B,1,256,256] of the caller
  int B, S, C, level;   // feature size S, feature channels C, pyramid level 0/1/2
  int src_shared;       // 1: `raw`/`stats` hold ONE image that every frame of the batch warps (clip mode)
  int fmt; void* d0; void* d1; int dC, dcoff, dpad;  // output: 2C channels at dcoff
};

struct OutConvP {
  const float* raw; const double* stats;  // model3.3 raw output [B,256,256,64] + stats (IN+ReLU on the fly)
  const float* w;                         // [onc][49][64] fp32
  const float* bias;                      // [onc]
  const IoPtrs* io;                       // io->out: NCHW [B,onc,256,256] of the caller
  int B, onc;
};

struct ReadP {  // debug tap: Act or Raw(+stats) -> NCHW fp32
  int B, H, W, C;      // logical view
  int fmt; const void* p0; const void* p1; int sC, scoff, spad;
  const double* stats; int stat_C, stat_coff; int relu;  // stats != null: normalise
  float* dst;
};

// (mean, rstd) of an affine-less InstanceNorm2d from the producer's {sum, sum of squares} (fp64), eps = 1e-5
// (nn.InstanceNorm2d default, Module2/models/networks.py:34): biased variance over the H*W plane.
#ifdef __CUDACC__
__device__ __forceinline__ void stats_to_affine(const double* st, int n, int stat_C, int stat_coff, int c, double inv_n,
                                                float* mean, float* rstd) {
  const double su = st[((size_t)n * stat_C + stat_coff + c) * 2 + 0];
  const double sq = st[((size_t)n * stat_C + stat_coff + c) * 2 + 1];
  const double m = su * inv_n;
  double var = sq * inv_n - m * m;  // sums are exact enough in double; the reference's own IN is fp32
  if (var < 0.0) var = 0.0;
  *mean = (float)m;
  *rstd = 1.0f / sqrtf((float)var + 1e-5f);
}
#endif

#ifdef __CUDACC__
// One warp; lane L holds {sum, sumsq} over the warp's pixels of channel c0 + L: store partial row `row` of image n.
__device__ __forceinline__ void stat_put(const StatSink& s, int n, int row, int c0, int lane, float cs, float cq) {
  s.part[((size_t)n * s.np + row) * s.C + s.coff + c0 + lane] = make_float2(cs, cq);
}
// One warp, after it has stored its rows of `nblk` consecutive 32-channel blocks starting at channel c0 of image n:
// one ticket per block; the last arriver of a block reduces it in row order.
__device__ __forceinline__ void stat_arrive(const StatSink& s, int n, int c0, int nblk, int lane) {
  __threadfence();
  __syncwarp();
  uint32_t* cnt = s.count + (size_t)n * (s.C >> 5) + ((s.coff + c0) >> 5);
  uint32_t t = 0;
  if (lane < nblk) t = atomicAdd(cnt + lane, 1u);
  uint32_t last = __ballot_sync(0xffffffffu, lane < nblk && t == (uint32_t)(s.np - 1));
  while (last) {
    const int b = __ffs(last) - 1;
    last &= last - 1;
    __threadfence();
    const float2* src = s.part + (size_t)n * s.np * s.C + s.coff + c0 + 32 * b + lane;
    double su = 0.0, sq = 0.0;
#pragma unroll 8
    for (int r = 0; r < s.np; ++r) {
      const float2 v = __ldcg(src + (size_t)r * s.C);
      su += (double)v.x;
      sq += (double)v.y;
    }
    double* dst = s.stats + ((size_t)n * s.C + s.coff + c0 + 32 * b + lane) * 2;
    dst[0] = su;
    dst[1] = sq;
    if (lane == 0) cnt[b] = 0;  // the counters are back at zero when the kernel ends
  }
}
#endif

// ---- launchers (each returns AP_OK / error and counts one launch) ----
int launch_conv_simt(const SimtConvP& p, cudaStream_t st);
int launch_apply(const ApplyP& p, cudaStream_t st);
int launch_set_io(IoPtrs* dst, const IoPtrs& v, cudaStream_t st);  // refreshes a plan's pointer table (one thread)
int launch_warp(const WarpP& p, cudaStream_t st);
int launch_out_conv(const OutConvP& p, cudaStream_t st);
int launch_read(const ReadP& p, cudaStream_t st);
// weight packing: src torch layout -> [slab][Cin][simt_C] fp32 at column simt_coff (simt) and/or
// [slab][Cout][Cin] bf16 hi,lo (umma)
int launch_pack_weights(const float* src, int Cout, int Cin, int k, int transposed, float* dst_simt, int simt_C,
                        int simt_coff, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo, cudaStream_t st);
int64_t launches_get();
void launches_add(int n);
int launch_pack_out_weights(const float* src, int onc, float* dst, cudaStream_t st);
// ConvTranspose2d weight [Cin][Cout][3][3] -> phase-packed [tap t][ph * Cout + co][Cin] bf16 hi / lo (see PhasePack);
// taps are (tdy[t], tdx[t]), phases (pk.py, pk.px)
int launch_pack_convT_phases(const float* src, int Cin, int Cout, const PhasePack& pk, int ntaps, const int* tdy,
                             const int* tdx, __nv_bfloat16* dst_hi, __nv_bfloat16* dst_lo, cudaStream_t st);
int launch_nchw_to_act(const float* src, const Act& dst, cudaStream_t st);  // debug conv helper

// ---- tcgen05 conv ----
struct UmmaConv;  // opaque launch record (tensor maps + params), see conv_umma.cu
// `sink`: where the InstanceNorm statistics of the output go (sink.stats null: none).  Partial row of (tile t of the
// image, epilogue warp q) = (4 t + q) * slot_mul + slot_add: a layer computed by several convs over the same virtual
// grid (the phases of a transposed conv) gives each conv its own slot_add < slot_mul; sink.np must be
// 4 * tiles per image * slot_mul.
int umma_conv_create(UmmaConv** out, const ConvGeom& g, const Act& in, int in_coff,
                     const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, int nprod,
                     float* out_raw, int out_C, int out_coff, const StatSink& sink, int slot_mul = 1, int slot_add = 0,
                     const PhasePack* pk = nullptr);
int umma_conv_stat_rows(const ConvGeom& g);  // 4 * tiles per image of this geometry
bool umma_pairs_available();  // CTA-pair kernels enabled and launchable on this device
void umma_conv_destroy(UmmaConv* c);
int umma_conv_launch(const UmmaConv* c, cudaStream_t st);
int umma_init();  // resolves cuTensorMapEncodeTiled, sets the kernel attributes on the current device
int umma_num_sms();
int stem_umma_init_device();  // per-device kernel attributes of conv_stem.cu / conv_out.cu (called by umma_init)
int out_umma_init_device();
int tmap_encode(CUtensorMap* m, int dtype, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box, const uint32_t* estr);
int tmap_encode_out(CUtensorMap* m, float* out, int B, int Hout, int Wout, int C, int os, int py, int px);

// ---- tcgen05 fused 7x7 stems (conv_stem.cu) ----
size_t stem_umma_weight_bytes();
int launch_pack_stem_umma(const float* src, int cout_s, int coff, uint8_t* img, cudaStream_t st);
int launch_stem_umma(const IoPtrs* io, const uint8_t* wimg, const Raw& out, int B, int nprod, cudaStream_t st);
constexpr int STEM_STAT_ROWS = 512;  // 128 groups of 4 tiles x 4 epilogue warps per image

// ---- tcgen05 output stage (conv_out.cu): IN+ReLU -> RefPad3 -> Conv7x7 64->onc -> bias -> tanh ----
size_t out_umma_weight_bytes(int onc);
int launch_pack_out_umma(const float* src, int onc, uint8_t* img, cudaStream_t st);
int launch_out_umma(const OutConvP& p, const uint8_t* wimg, cudaStream_t st);

// ---- landmark branch (landmark.cu): three direct convs over both landmark maps ----
int launch_landmark_branch(const IoPtrs* io, const float* w0, const float* w1, const float* w2,
                           const Raw& r0, const Raw& r1, const Raw& r2, int B1, int B2, cudaStream_t st);

ConvTaps make_taps_conv(int k, int pad, int extra_origin);
ConvTaps make_taps_convT_phase(int py, int px);

}  // namespace ap
