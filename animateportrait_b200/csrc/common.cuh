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

// InstanceNorm statistics, deterministic and independent of the batch size.  Producers accumulate per-(n,c)
// {sum, sum of squares} as 64-bit FIXED-POINT integers (STAT_FRAC fractional bits): every leaf -- the float sum of one
// channel over the 32 pixels one epilogue warp holds, computed in a fixed order -- is rounded to the grid once, and
// integer addition is associative, so on-chip accumulation and fire-and-forget atomics in ANY order and grouping give
// bit-identical totals: from run to run, for any batch size, grid or stream timing -- like the reference's CPU
// instance_norm (Module2/models/networks.py:34), with no tickets, no partial buffers and no reduction tail.
// Precision: a leaf is off by <= 2^-23; over P leaves of a plane of N pixels the mean and the mean square are off by
// about sqrt(P) * 2^-23 / N  (<= 4e-10 for every layer of this network) -- against eps = 1e-5 inside the square root
// that changes rstd by < 1e-5 relative, whatever the scale of the channel.  Range: |sum of squares| < 2^63 / 2^22 =
// 2.2e12, i.e. an rms of the pre-norm activations below 5,800 at 256x256 (23,000 at 64x64).
typedef long long stat_t;
constexpr int STAT_FRAC = 22;
constexpr double STAT_SCALE = 4194304.0;          // 2^22
constexpr double STAT_INV_SCALE = 1.0 / 4194304.0;

// Raw (pre-InstanceNorm) conv output: fp32 NHWC, no halo, plus per-(n,c) {sum, sumsq} in fixed point.
struct Raw {
  int B = 0, H = 0, W = 0, C = 0;
  float* p = nullptr;
  stat_t* stats = nullptr;  // [B][C][2], zeroed at the start of every forward
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
  stat_t* stats;  // may be null; indexed [(n*stat_C + stat_coff + c)*2]
  int stat_C, stat_coff;
};

struct ApplyP {
  const float* raw; int raw_C, raw_coff;
  const stat_t* stats; int stat_C, stat_coff;  // null -> no normalisation (bias mode)
  const float* bias;                           // null or [C]
  const float* raw2; int raw2_C, raw2_coff;    // optional second InstanceNorm'ed operand (ResnetBlock2 shortcut)
  const stat_t* stats2; int stat2_C, stat2_coff;
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
  const stat_t* stats; int stat_C, stat_coff;
  const IoPtrs* io;     // motion [B,256,256,2], flow [B,2,256,256], ifmask [B,1,256,256] of the caller
  int B, S, C, level;   // feature size S, feature channels C, pyramid level 0/1/2
  int src_shared;       // 1: `raw`/`stats` hold ONE image that every frame of the batch warps (clip mode)
  int fmt; void* d0; void* d1; int dC, dcoff, dpad;  // output: 2C channels at dcoff
};

struct OutConvP {
  const float* raw; const stat_t* stats;  // model3.3 raw output [B,256,256,64] + stats (IN+ReLU on the fly)
  const float* w;                         // [onc][49][64] fp32
  const float* bias;                      // [onc]
  const IoPtrs* io;                       // io->out: NCHW [B,onc,256,256] of the caller
  int B, onc;
};

struct ReadP {  // debug tap: Act or Raw(+stats) -> NCHW fp32
  int B, H, W, C;      // logical view
  int fmt; const void* p0; const void* p1; int sC, scoff, spad;
  const stat_t* stats; int stat_C, stat_coff; int relu;  // stats != null: normalise
  float* dst;
};

// (mean, rstd) of an affine-less InstanceNorm2d from the producer's {sum, sum of squares} (fixed point), eps = 1e-5
// (nn.InstanceNorm2d default, Module2/models/networks.py:34): biased variance over the H*W plane.
#ifdef __CUDACC__
__device__ __forceinline__ void stats_to_affine(const stat_t* st, int n, int stat_C, int stat_coff, int c, double inv_n,
                                                float* mean, float* rstd) {
  const double su = (double)st[((size_t)n * stat_C + stat_coff + c) * 2 + 0] * STAT_INV_SCALE;
  const double sq = (double)st[((size_t)n * stat_C + stat_coff + c) * 2 + 1] * STAT_INV_SCALE;
  const double m = su * inv_n;
  double var = sq * inv_n - m * m;  // sums are exact enough in double; the reference's own IN is fp32
  if (var < 0.0) var = 0.0;
  *mean = (float)m;
  *rstd = 1.0f / sqrtf((float)var + 1e-5f);
}
// leaf of the statistics: a float partial sum on the fixed-point grid
__device__ __forceinline__ stat_t stat_fix(float v) { return __double2ll_rn((double)v * STAT_SCALE); }
// fire-and-forget 64-bit integer atomic (RED.ADD.64): order-independent
__device__ __forceinline__ void stat_add(stat_t* dst, stat_t v) {
  atomicAdd(reinterpret_cast<unsigned long long*>(dst), static_cast<unsigned long long>(v));
}
#endif

// ---- launchers (each returns AP_OK / error and counts one launch) ----
int launch_conv_simt(const SimtConvP& p, cudaStream_t st);
int launch_apply(const ApplyP& p, cudaStream_t st);
int launch_set_io(IoPtrs* dst, const IoPtrs& v, cudaStream_t st);  // refreshes a plan's pointer table (one thread)
int launch_warp(const WarpP& p, cudaStream_t st);
int launch_out_conv(const OutConvP& p, cudaStream_t st);
int launch_read(const ReadP& p, cudaStream_t st);
int launch_stats_to_double(const stat_t* src, double* dst, size_t n, cudaStream_t st);  // debug: fixed point -> double
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
int umma_conv_create(UmmaConv** out, const ConvGeom& g, const Act& in, int in_coff,
                     const __nv_bfloat16* w_hi, const __nv_bfloat16* w_lo, int nprod,
                     float* out_raw, int out_C, int out_coff, stat_t* stats, int stat_C, int stat_coff,
                     const PhasePack* pk = nullptr);
bool umma_pairs_available();  // CTA-pair kernels enabled and launchable on this device
void umma_conv_destroy(UmmaConv* c);
int umma_conv_launch(const UmmaConv* c, cudaStream_t st);
int umma_init();  // resolves cuTensorMapEncodeTiled, sets the kernel attributes on the current device
int umma_num_sms();
// ---- halo-staged trunk conv (conv_halo.cu): 3x3 stride-1 convs on 64x64, N = 256 ----
struct HaloConv;
int halo_init_device();
bool halo_conv_eligible(const ConvGeom& g, const Act& in, int nprod, bool packed);
int halo_conv_create(HaloConv** out, const ConvGeom& g, const Act& in, int in_coff, const __nv_bfloat16* w_hi,
                     const __nv_bfloat16* w_lo, int nprod, float* out_raw, int out_C, int out_coff, stat_t* stats,
                     int stat_C, int stat_coff);
void halo_conv_destroy(HaloConv* c);
int halo_conv_launch(const HaloConv* c, cudaStream_t st);
int stem_umma_init_device();  // per-device kernel attributes of conv_stem.cu / conv_out.cu (called by umma_init)
int out_umma_init_device();
int tmap_encode(CUtensorMap* m, int dtype, const void* base, int rank, const uint64_t* dims, const uint64_t* strides,
                const uint32_t* box, const uint32_t* estr);
int tmap_encode_out(CUtensorMap* m, float* out, int B, int Hout, int Wout, int C, int os, int py, int px);

// ---- tcgen05 fused 7x7 stems (conv_stem.cu) ----
size_t stem_umma_weight_bytes();
int launch_pack_stem_umma(const float* src, int cout_s, int coff, uint8_t* img, cudaStream_t st);
int launch_stem_umma(const IoPtrs* io, const uint8_t* wimg, const Raw& out, int B, int nprod, cudaStream_t st);

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
