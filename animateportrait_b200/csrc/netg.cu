// Handle, weight loading, workspace planning, the forward schedule and the C ABI (include/ap_netg.h)
// of the B200-native generator `ResnetConditionTriGenerator32_full_ifw`
// (reference: Module2/models/networks.py:1190-1340).
//
// The forward schedule is written once (Runner::run) and interpreted in three phases:
//   SIZE  - walk the graph, add up workspace bytes
//   BUILD - same walk with real pointers: creates the TMA tensor maps of every tcgen05 conv, records debug taps
//   EXEC  - same walk, launching kernels on the caller's stream
// so buffer assignment, tensor maps and launches can never disagree.
#include <stdarg.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <string>
#include <vector>

#include "common.cuh"

namespace ap {

static thread_local char g_err[1024] = "";
void set_error(const char* fmt, ...) {
  va_list ap_;
  va_start(ap_, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap_);
  va_end(ap_);
}

struct LayerSpec {
  std::string name;
  int cout, cin, k;
  bool transposed;
};

static std::vector<LayerSpec> layer_specs(int onc) {
  std::vector<LayerSpec> v;
  v.push_back({"model_tri_merge", 256, 768, 3, false});
  v.push_back({"model_tri00.1", 32, 3, 7, false});
  v.push_back({"model_tri01.0", 128, 64, 3, false});
  v.push_back({"model_tri02.0", 256, 128, 3, false});
  v.push_back({"model_tri10.1", 64, 3, 7, false});
  v.push_back({"model_tri11.0", 64, 64, 3, false});
  v.push_back({"model_tri12.0", 256, 128, 3, false});
  v.push_back({"model_tri20.1", 64, 3, 7, false});
  v.push_back({"model_tri21.0", 128, 64, 3, false});
  v.push_back({"model_tri22.0", 128, 128, 3, false});
  for (int i = 0; i < 9; ++i) {
    const std::string b = "model2." + std::to_string(i);
    const bool b2 = (i + 3) % 3 == 0;  // (i + disp) % div == 0, networks.py:1259
    v.push_back({b + ".conv_block.1", 256, b2 ? 288 : 256, 3, false});
    v.push_back({b + ".conv_block.5", 256, 256, 3, false});
    if (b2) v.push_back({b + ".shortcut.0", 256, 288, 3, false});
  }
  v.push_back({"model3.0", 128, 256, 3, true});
  v.push_back({"model3.3", 64, 128, 3, true});
  v.push_back({"model3.7", onc, 64, 7, false});
  v.push_back({"model_landmark_trans.0", 8, 1, 3, false});
  v.push_back({"model_landmark_trans.3", 16, 8, 3, false});
  v.push_back({"model_landmark_trans.6", 16, 16, 3, false});
  return v;
}

struct PackedT {  // phase-packed weights of a transposed conv (common.cuh: PhasePack)
  PhasePack pk;
  int ntaps;
  int tdy[4], tdx[4];
  __nv_bfloat16* hi = nullptr;
  __nv_bfloat16* lo = nullptr;
};

struct LayerW {
  int cout = 0, cin = 0, k = 0;
  std::vector<PackedT> packT;     // transposed convs on the tensor-core path
  float* simt = nullptr;          // [slab][Cin][Cout]
  __nv_bfloat16* hi = nullptr;    // [slab][Cout][Cin]
  __nv_bfloat16* lo = nullptr;
  std::vector<float> host;        // landmark layers: [slab][Cin][Cout] on the host (kernel-parameter weights)
};

struct TapRec {
  int B, H, W, C;
  int fmt; const void* p0; const void* p1; int sC, scoff, spad;
  const stat_t* stats; int stat_C, stat_coff; int relu;
};

constexpr size_t AP_MAX_PLANS = 6;

struct Plan {
  int B = 0;
  // clip mode: ONE photo shared by all B frames.  Everything that depends on the photo alone (the three stems, tri11,
  // tri21, tri22 and their InstanceNorms) is computed for a batch of 1 and the warps read it for every frame.
  bool shared_photo = false;
  uint64_t last_use = 0;  // plan cache is LRU-bounded (AP_MAX_PLANS): ragged tail batches of clips must not pile up arenas
  char* arena = nullptr;
  size_t arena_bytes = 0;
  char* sarena = nullptr;  // InstanceNorm statistics (fixed-point accumulators), zeroed at the start of every forward
  size_t sarena_bytes = 0;
  bool keep_all = false;   // debug taps: every intermediate keeps its own buffer (no reuse of dead buffers)
  IoPtrs* d_io = nullptr;  // pointer table of the caller's tensors, read by the kernels that touch caller memory
  // the launch sequence of a plan does not depend on the call (pointer table): captured once, replayed as a CUDA graph
  cudaGraphExec_t gexec = nullptr;
  cudaStream_t cap_stream = nullptr;
  int64_t graph_launches = 0;
  std::vector<UmmaConv*> convs;
  std::map<std::string, TapRec> taps;
  // staging for ap_netg_forward_host
  // two slots: while the forward of one host call computes, the inputs of the next call upload into the other slot and
  // the frames of the previous one download (ap_netg_forward_host_async)
  struct HostSlot {
    float* h_in[6] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
    float* h_out = nullptr;
    cudaEvent_t ev_in[3] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev_computed = nullptr, ev_done = nullptr;
    bool busy = false;
  };
  HostSlot hs[2];
  uint64_t host_calls = 0;
  cudaStream_t copy_stream = nullptr, d2h_stream = nullptr;
  // independent branches of the graph (encoder branches, landmark branch, ResnetBlock2 shortcuts) run on side
  // streams so that HBM-bound kernels overlap the tensor-bound persistent convs; fork/join events by index
  cudaStream_t side[3] = {nullptr, nullptr, nullptr};
  std::vector<cudaEvent_t> sync_ev;
  ~Plan() {
    for (cudaStream_t s_ : side) if (s_) cudaStreamDestroy(s_);
    for (cudaEvent_t e : sync_ev) cudaEventDestroy(e);
    for (UmmaConv* c : convs) umma_conv_destroy(c);
    if (gexec) cudaGraphExecDestroy(gexec);
    if (cap_stream) cudaStreamDestroy(cap_stream);
    if (d_io) cudaFree(d_io);
    if (arena) cudaFree(arena);
    if (sarena) cudaFree(sarena);
    for (HostSlot& sl : hs) {
      for (float* p : sl.h_in) if (p) cudaFree(p);
      if (sl.h_out) cudaFree(sl.h_out);
      for (cudaEvent_t e : sl.ev_in) if (e) cudaEventDestroy(e);
      if (sl.ev_computed) cudaEventDestroy(sl.ev_computed);
      if (sl.ev_done) cudaEventDestroy(sl.ev_done);
    }
    if (copy_stream) cudaStreamDestroy(copy_stream);
    if (d2h_stream) cudaStreamDestroy(d2h_stream);
  }
};


// Phase-packed weights of one ConvTranspose2d layer.  Cout = 64: all four phases in one N = 256 conv over the four
// input taps.  Cout = 128: phases (0,0),(0,1) need only the two taps of the same input row; (1,0),(1,1) need all four.
static int make_packed_convT(const float* src_dev, int cin, int cout, bool with_lo, cudaStream_t st, std::vector<void*>* owned,
                             std::vector<PackedT>* out) {
  AP_REQUIRE(cout == 64 || cout == 128, AP_ERR_UNSUPPORTED, "phase-packed transposed conv: Cout=%d", cout);
  const int npack = cout == 64 ? 1 : 2;
  for (int k = 0; k < npack; ++k) {
    PackedT pt{};
    pt.pk.cols = cout;
    pt.pk.nph = 256 / cout;
    for (int i = 0; i < pt.pk.nph; ++i) {
      const int ph = (npack == 1) ? i : 2 * k + i;
      pt.pk.py[i] = ph >> 1;
      pt.pk.px[i] = ph & 1;
    }
    pt.ntaps = (npack == 2 && k == 0) ? 2 : 4;
    for (int t = 0; t < pt.ntaps; ++t) { pt.tdy[t] = t >> 1; pt.tdx[t] = t & 1; }
    const size_t pe = (size_t)pt.ntaps * 256 * cin;
    AP_CUDA(cudaMalloc((void**)&pt.hi, pe * 2));
    owned->push_back(pt.hi);
    if (with_lo) {
      AP_CUDA(cudaMalloc((void**)&pt.lo, pe * 2));
      owned->push_back(pt.lo);
    }
    AP_TRY(launch_pack_convT_phases(src_dev, cin, cout, pt.pk, pt.ntaps, pt.tdy, pt.tdx, pt.hi, pt.lo, st));
    out->push_back(pt);
  }
  return AP_OK;
}

static ConvGeom geom_convT_packed(int B, int Hin, int Cin, const PackedT& pt) {
  ConvGeom g{};
  g.B = B; g.Hin = Hin; g.Win = Hin; g.Cin = Cin;
  g.Hv = Hin; g.Wv = Hin; g.stride = 1; g.reflect = 0; g.Cout = 256;
  g.os = 1; g.py = 0; g.px = 0; g.Hout = 2 * Hin; g.Wout = 2 * Hin;
  g.taps.n = pt.ntaps;
  for (int t = 0; t < pt.ntaps; ++t) { g.taps.dy[t] = (int8_t)pt.tdy[t]; g.taps.dx[t] = (int8_t)pt.tdx[t]; g.taps.slab[t] = (uint8_t)t; }
  return g;
}

}  // namespace ap

using namespace ap;

constexpr int AP_NCLASS = 7;
enum { CL_STEM = 0, CL_LAND = 1, CL_TRUNK = 2, CL_STRIDED = 3, CL_APPLY = 4, CL_WARP = 5, CL_OUT = 6 };

struct ap_netg {
  int onc = 1, prec = 0, device = 0;
  bool profiling = false;
  bool overlap = true;  // run independent branches on side streams (AP_NETG_OVERLAP=0 turns it off)
  bool out_umma = true; // tcgen05 output stage (AP_NETG_OUT_UMMA=0: CUDA-core kernel)
  bool convt_packed = true;  // transposed convs as one phase-packed N=256 conv (AP_NETG_CONVT_PACKED=0 or no CTA pairs: four phase convs)
  bool graphs = true;        // replay the forward of a plan as a CUDA graph (AP_NETG_GRAPH=0 / option "graphs": stream launches)
  bool keep_all = false;     // option "keep_intermediates": no buffer reuse, so that every debug tap survives the forward
  std::vector<cudaEvent_t> ev;      // ev[0] = start, ev[i+1] = after launch i
  std::vector<int> ev_class;        // class of launch i
  std::vector<double> ev_flops;
  size_t ev_used = 0;
  bool loaded = false;
  std::map<std::string, LayerW> w;
  float* w_stem = nullptr;   // fused stems [49][3][160] (CUDA-core path)
  uint8_t* w_stem_img = nullptr;  // fused stems, pre-swizzled bf16 hi/lo smem image (tcgen05 path)
  float* w_out = nullptr;    // [onc][49][64]
  uint8_t* w_out_img = nullptr;  // tcgen05 output stage: pre-swizzled bf16 hi/lo images per output channel
  float* b_merge = nullptr;  // [256]
  float* b_out = nullptr;    // [onc]
  std::vector<void*> owned;  // every device allocation holding weights
  std::map<int, Plan*> plans;  // key = 4 * B + 2 * keep_all + shared_photo
  uint64_t use_clock = 0;
  Plan* last_plan = nullptr;
  int64_t last_launches = 0;
};

namespace ap {

enum Phase { PH_SIZE = 0, PH_BUILD = 1, PH_EXEC = 2 };

struct Inputs {
  const float *input, *land1, *land2, *motion, *flow, *ifmask;
  float* out;
};

static inline size_t align_up(size_t v, size_t a) { return (v + a - 1) / a * a; }

static ConvGeom geom_convT_phase_(int B, int Hin, int Cin, int Cout, int py, int px);

struct Runner {
  ap_netg* h;
  Plan* pl;
  Phase ph;
  cudaStream_t st;
  size_t off = 0, peak = 0, soff = 0, conv_i = 0;
  const cudaEvent_t* in_ready = nullptr;  // host-buffer entry point: [0] photo, [1] motion/flow/ifmask, [2] land1/land2

  int wait_input(int k) {
    if (ph == PH_EXEC && in_ready) AP_CUDA(cudaStreamWaitEvent(st, in_ready[k], 0));
    return AP_OK;
  }

  // ---- branch-level concurrency: `st` is the stream the next launches go to ----
  cudaStream_t main_st = nullptr;
  size_t ev_i = 0;
  bool overlap() const { return ph == PH_EXEC && !h->profiling && h->overlap; }
  int next_event(cudaEvent_t* e) {
    if (ev_i >= pl->sync_ev.size()) {
      cudaEvent_t n;
      AP_CUDA(cudaEventCreateWithFlags(&n, cudaEventDisableTiming));
      pl->sync_ev.push_back(n);
    }
    *e = pl->sync_ev[ev_i++];
    return AP_OK;
  }
  // everything launched on `to` from now on is ordered after what has been launched on `from` so far
  int order_after(cudaStream_t to, cudaStream_t from) {
    if (!overlap() || to == from) return AP_OK;
    cudaEvent_t e;
    AP_TRY(next_event(&e));
    AP_CUDA(cudaEventRecord(e, from));
    AP_CUDA(cudaStreamWaitEvent(to, e, 0));
    return AP_OK;
  }
  cudaStream_t side(int k) {
    if (!overlap()) return main_st;
    return pl->side[k];
  }
  void on(cudaStream_t s_) { st = s_; }

  void* alloc(size_t bytes) {
    off = align_up(off, 1024);
    void* p = (ph == PH_SIZE) ? nullptr : (void*)(pl->arena + off);
    off += bytes;
    if (off > peak) peak = off;
    return p;
  }
  void* salloc(size_t bytes) {
    soff = align_up(soff, 256);
    void* p = (ph == PH_SIZE) ? nullptr : (void*)(pl->sarena + soff);
    soff += bytes;
    return p;
  }
  Act act(int B, int H, int W, int C, int pad, int fmt) {
    Act a;
    a.B = B; a.H = H; a.W = W; a.C = C; a.pad = pad; a.fmt = fmt;
    const size_t es = (fmt == FMT_F32) ? 4 : 2;
    a.p0 = alloc(a.elems() * es);
    if (fmt == FMT_BF16X2) a.p1 = alloc(a.elems() * es);
    return a;
  }
  Raw raw(int B, int H, int W, int C, bool stats) {
    Raw r;
    r.B = B; r.H = H; r.W = W; r.C = C;
    r.p = (float*)alloc((size_t)B * H * W * C * sizeof(float));
    if (stats) r.stats = (stat_t*)salloc((size_t)B * C * 2 * sizeof(stat_t));
    return r;
  }
  // a reused data buffer with statistics of its own: the accumulators of every layer are zeroed ONCE per forward
  Raw with_stats(Raw r) {
    r.stats = (stat_t*)salloc((size_t)r.B * r.C * 2 * sizeof(stat_t));
    return r;
  }
  void tap_act(const char* name, const Act& a, int coff, int C) {
    if (ph != PH_BUILD) return;
    pl->taps[name] = TapRec{a.B, a.H, a.W, C, a.fmt, a.p0, a.p1, a.C, coff, a.pad, nullptr, 0, 0, 0};
  }
  void tap_raw(const char* name, const Raw& r, int coff, int C, int relu) {
    if (ph != PH_BUILD) return;
    pl->taps[name] = TapRec{r.B, r.H, r.W, C, FMT_F32, r.p, nullptr, r.C, coff, 0, r.stats, r.C, coff, relu};
  }
  void tap_f32(const char* name, const float* p, int B, int H, int W, int C) {
    if (ph != PH_BUILD) return;
    pl->taps[name] = TapRec{B, H, W, C, FMT_F32, p, nullptr, C, 0, 0, nullptr, 0, 0, 0};
  }

  const LayerW& W(const std::string& n) { return h->w.at(n); }

  // profiling: one event after every launch (stream order => consecutive differences are kernel times)
  int mark(int cls, double flops) {
    if (ph != PH_EXEC || !h->profiling) return AP_OK;
    if (h->ev_used + 1 >= h->ev.size()) {
      cudaEvent_t e;
      AP_CUDA(cudaEventCreate(&e));
      h->ev.push_back(e);
    }
    AP_CUDA(cudaEventRecord(h->ev[h->ev_used + 1], st));
    h->ev_class.push_back(cls);
    h->ev_flops.push_back(flops);
    ++h->ev_used;
    return AP_OK;
  }
  static double conv_flops(const ConvGeom& g) {
    return 2.0 * g.B * g.Hv * g.Wv * (double)g.Cin * g.Cout * g.taps.n;
  }

  // A 3x3 / transposed-conv layer on the tensor-core or the CUDA-core path, by handle precision.
  int conv(const ConvGeom& g, const Act& in, int in_coff, const LayerW& w, const Raw& out, int out_coff) {
    if (h->prec == AP_PREC_FP32_SIMT) {
      if (ph != PH_EXEC) return AP_OK;
      SimtConvP p{};
      p.g = g;
      p.in = (const float*)in.p0; p.in_nchw = 0; p.in_C = in.C; p.in_coff = in_coff;
      p.wpk = w.simt;
      p.out = out.p; p.out_C = out.C; p.out_coff = out_coff;
      p.stats = out.stats; p.stat_C = out.C; p.stat_coff = out_coff;
      AP_TRY(launch_conv_simt(p, st));
      return mark((g.stride == 1 && g.os == 1) ? CL_TRUNK : CL_STRIDED, conv_flops(g));
    }
    if (ph == PH_SIZE) return AP_OK;
    if (ph == PH_BUILD) {
      UmmaConv* c = nullptr;
      AP_TRY(umma_conv_create(&c, g, in, in_coff, w.hi, w.lo, h->prec == AP_PREC_FP32X3 ? 3 : 1, out.p, out.C, out_coff,
                              out.stats, out.C, out_coff));
      pl->convs.push_back(c);
      return AP_OK;
    }
    AP_TRY(umma_conv_launch(pl->convs.at(conv_i++), st));
    return mark((g.stride == 1 && g.os == 1) ? CL_TRUNK : CL_STRIDED, conv_flops(g));
  }
  // ConvTranspose2d(k3,s2,p1,op1) (networks.py:1271-1274): one or two phase-packed N = 256 convs on the CTA-pair
  // kernel, or (CUDA-core mode / AP_NETG_CONVT_PACKED=0 / no CTA pairs on this device) four per-phase convs
  int convT(const Act& in, const LayerW& w, const Raw& out, int Hin, int Cin, int Cout) {
    if (h->prec == AP_PREC_FP32_SIMT || !h->convt_packed || w.packT.empty()) {
      for (int ph_ = 0; ph_ < 4; ++ph_)
        AP_TRY(conv(geom_convT_phase_(pl->B, Hin, Cin, Cout, ph_ >> 1, ph_ & 1), in, 0, w, out, 0));
      return AP_OK;
    }
    if (ph == PH_SIZE) return AP_OK;
    const int npack = (int)w.packT.size();
    for (int k = 0; k < npack; ++k) {
      const PackedT& pt = w.packT[k];
      if (ph == PH_BUILD) {
        const ConvGeom g = geom_convT_packed(pl->B, Hin, Cin, pt);
        UmmaConv* c = nullptr;
        AP_TRY(umma_conv_create(&c, g, in, 0, pt.hi, pt.lo, h->prec == AP_PREC_FP32X3 ? 3 : 1, out.p, out.C, 0, out.stats,
                                out.C, 0, &pt.pk));
        pl->convs.push_back(c);
      } else {
        AP_TRY(umma_conv_launch(pl->convs.at(conv_i++), st));
        AP_TRY(mark(CL_STRIDED, 2.0 * pl->B * Hin * Hin * (double)Cin * Cout * 9 / (double)npack));
      }
    }
    return AP_OK;
  }
  // thin CUDA-core layers of the validation mode (stems): fp32 input, NCHW or NHWC
  int conv_thin(const ConvGeom& g, const float* in, int nchw, int in_C, const float* wpk, const Raw& out) {
    if (ph != PH_EXEC) return AP_OK;
    SimtConvP p{};
    p.g = g;
    p.in = in; p.in_nchw = nchw; p.in_C = in_C; p.in_coff = 0;
    p.wpk = wpk;
    p.out = out.p; p.out_C = out.C; p.out_coff = 0;
    p.stats = out.stats; p.stat_C = out.C; p.stat_coff = 0;
    AP_TRY(launch_conv_simt(p, st));
    return mark(g.taps.n == 49 ? CL_STEM : CL_LAND, conv_flops(g));
  }
  static ApplyP make_apply(const Raw& r, int rcoff, int C, int relu, const Act* dst, int dcoff, int halo, const float* bias,
                           const Raw* r2, const float* res_in, float* res_out, const Act* res_act) {
    ApplyP p{};
    p.raw = r.p; p.raw_C = r.C; p.raw_coff = rcoff;
    p.stats = r.stats; p.stat_C = r.C; p.stat_coff = rcoff;
    p.bias = bias;
    if (r2) { p.raw2 = r2->p; p.raw2_C = r2->C; p.raw2_coff = 0; p.stats2 = r2->stats; p.stat2_C = r2->C; p.stat2_coff = 0; }
    p.res_in = res_in; p.res_out = res_out;
    p.res_fmt = -1;
    if (res_act) { p.res_fmt = res_act->fmt; p.res_p0 = res_act->p0; p.res_p1 = res_act->p1; p.res_C = res_act->C; p.res_pad = res_act->pad; }
    p.relu = relu;
    p.B = r.B; p.H = r.H; p.W = r.W; p.C = C;
    if (dst) { p.fmt = dst->fmt; p.d0 = dst->p0; p.d1 = dst->p1; p.dC = dst->C; p.dcoff = dcoff; p.dpad = dst->pad; }
    else p.fmt = -1;
    p.halo_reflect = halo;
    return p;
  }
  int apply(const Raw& r, int rcoff, int C, int relu, const Act* dst, int dcoff, int halo, const float* bias = nullptr,
            const Raw* r2 = nullptr, const float* res_in = nullptr, float* res_out = nullptr, const Act* res_act = nullptr) {
    if (ph != PH_EXEC) return AP_OK;
    ApplyP p = make_apply(r, rcoff, C, relu, dst, dcoff, halo, bias, r2, res_in, res_out, res_act);
    if (dst && r.B == 1 && dst->B > 1) { p.B = dst->B; p.src_shared = 1; }  // clip mode: one image into every frame's slot
    AP_TRY(launch_apply(p, st));
    return mark(CL_APPLY, 0.0);
  }
  int warp(const Raw& r, int rcoff, int C, int level, const Act& dst, int dcoff) {
    if (ph != PH_EXEC) return AP_OK;
    WarpP p{};
    p.raw = r.p; p.raw_C = r.C; p.raw_coff = rcoff;
    p.stats = r.stats; p.stat_C = r.C; p.stat_coff = rcoff;
    p.io = pl->d_io;
    p.B = pl->B; p.S = r.H; p.C = C; p.level = level;
    p.src_shared = (r.B == 1 && pl->B > 1) ? 1 : 0;  // clip mode: one photo's features, warped per frame
    p.fmt = dst.fmt; p.d0 = dst.p0; p.d1 = dst.p1; p.dC = dst.C; p.dcoff = dcoff; p.dpad = dst.pad;
    AP_TRY(launch_warp(p, st));
    return mark(CL_WARP, 0.0);
  }

  int run(const Inputs& in);
};

static ConvGeom geom_conv(int B, int Hin, int Cin, int Cout, int k, int stride, int pad, int reflect) {
  ConvGeom g{};
  g.B = B; g.Hin = Hin; g.Win = Hin; g.Cin = Cin;
  g.Hv = Hin / stride; g.Wv = Hin / stride;
  g.stride = stride; g.reflect = reflect; g.Cout = Cout;
  g.os = 1; g.py = 0; g.px = 0; g.Hout = g.Hv; g.Wout = g.Wv;
  g.taps = make_taps_conv(k, pad, 0);
  return g;
}
static ConvGeom geom_convT_phase(int B, int Hin, int Cin, int Cout, int py, int px);
static ConvGeom geom_convT_phase_(int B, int Hin, int Cin, int Cout, int py, int px) { return geom_convT_phase(B, Hin, Cin, Cout, py, px); }
static ConvGeom geom_convT_phase(int B, int Hin, int Cin, int Cout, int py, int px) {
  ConvGeom g{};
  g.B = B; g.Hin = Hin; g.Win = Hin; g.Cin = Cin;
  g.Hv = Hin; g.Wv = Hin;
  g.stride = 1; g.reflect = 0; g.Cout = Cout;
  g.os = 2; g.py = py; g.px = px; g.Hout = 2 * Hin; g.Wout = 2 * Hin;
  g.taps = make_taps_convT_phase(py, px);
  return g;
}

int Runner::run(const Inputs& in) {
  const int B = pl->B;
  const int Bp = pl->shared_photo ? 1 : B;  // batch of the photo-only part of the encoder
  const int prec = h->prec;
  const int afmt = (prec == AP_PREC_FP32X3) ? FMT_BF16X2 : (prec == AP_PREC_BF16 ? FMT_BF16 : FMT_F32);
  const int hp = (prec == AP_PREC_FP32_SIMT) ? 0 : 1;  // halo of reflect-padded tensor-core inputs
  const bool keep = pl->keep_all;
  const IoPtrs* io = pl->d_io;

  main_st = st;
  ev_i = 0;
  if (ph == PH_EXEC && pl->sarena_bytes) AP_CUDA(cudaMemsetAsync(pl->sarena, 0, pl->sarena_bytes, st));
  cudaStream_t s1 = side(0), s2 = side(1), s3 = side(2);

  // trunk buffers: Xb[0] = merge output, Xb[i+1] = output of block i.  The inputs of the three ResnetBlock2 carry
  // 2 x 16 extra channels, cat[x, l1, l2] (networks.py:1335); the landmark branch fills them early on a side stream.
  // Tensor-core modes: the residual stream is a separate fp32 buffer (xres) and the operand buffers are
  // reused in place (XL for the 288-channel inputs, X for the others) so that they stay L2-resident.  Reading the
  // residual back from the bf16 hi/lo operand planes instead was measured slower (two 8-byte loads per thread and
  // fresh output lines every block: in_apply 1.34 -> 1.50 ms/step at B=16) and is only used by the CUDA-core mode,
  // whose operands are fp32 anyway.
  // Buffers whose contents are dead are reused (the residual stream ping-pongs between two buffers, the raw conv
  // outputs of all blocks share three, the decoder lives where the encoder was) unless the plan keeps every
  // intermediate for the debug taps (option "keep_intermediates").
  const bool res_from_act = (prec == AP_PREC_FP32_SIMT);
  Act Xb[10];
  if (res_from_act) {
    for (int i = 0; i < 10; ++i) Xb[i] = act(B, 64, 64, (i % 3 == 0 && i < 9) ? 288 : 256, i == 9 ? 0 : hp, afmt);
  } else {
    const Act XL = act(B, 64, 64, 288, hp, afmt), X = act(B, 64, 64, 256, hp, afmt), D0 = act(B, 64, 64, 256, 0, afmt);
    for (int i = 0; i < 9; ++i) Xb[i] = (i % 3 == 0) ? XL : X;
    Xb[9] = D0;
  }
  Act T = act(B, 64, 64, 256, hp, afmt);
  float* xres[10] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  if (!res_from_act) {
    const int nres = keep ? 10 : 2;
    for (int i = 0; i < nres; ++i) xres[i] = (float*)alloc((size_t)B * 64 * 64 * 256 * sizeof(float));
    for (int i = nres; i < 10; ++i) xres[i] = xres[i & 1];
  }
  Raw trunk_raw[3];  // conv_block.1, shortcut.0, conv_block.5 of whichever block is running (data buffers only: the
  if (!keep) for (int i = 0; i < 3; ++i) trunk_raw[i] = raw(B, 64, 64, 256, false);  // statistics are per layer)
  const size_t encoder_begin = off;

  // ---- landmark branch on land1 and land2 as one batch of 2B maps (networks.py:1280-1282, 1331-1332) ----
  {
    const int B1 = Bp;  // land1 is the SOURCE landmark map: it belongs to the photo (one map in clip mode)
    Raw rl0 = raw(B1 + B, 256, 256, 8, true);
    Raw rl1 = raw(B1 + B, 128, 128, 16, true);
    Raw rl2 = raw(B1 + B, 64, 64, 16, true);
    AP_TRY(order_after(s3, main_st));
    on(s3);
    AP_TRY(wait_input(2));
    if (ph == PH_EXEC) {
      AP_TRY(launch_landmark_branch(io, W("model_landmark_trans.0").host.data(), W("model_landmark_trans.3").host.data(),
                                    W("model_landmark_trans.6").host.data(), rl0, rl1, rl2, B1, B, st));
      AP_TRY(mark(CL_LAND, 2.0 * (B1 + B) * 9.0 * (65536.0 * 8 + 16384.0 * 8 * 16 + 4096.0 * 16 * 16)));
    }
    for (int li = 0; li < 2; ++li) {
      Raw v = rl2;  // view of this landmark's part of the batch: [0, B1) = land1, [B1, B1 + B) = land2
      v.B = li == 0 ? B1 : B;
      if (ph != PH_SIZE) {
        v.p = rl2.p + (size_t)li * B1 * 64 * 64 * 16;
        v.stats = rl2.stats + (size_t)li * B1 * 16 * 2;
      }
      for (int k = 0; k < (res_from_act ? 9 : 3); k += 3) AP_TRY(apply(v, 0, 16, 0, &Xb[k], 256 + 16 * li, 1));
      tap_act(li == 0 ? "land1" : "land2", Xb[0], 256 + 16 * li, 16);
    }
    on(main_st);
  }

  // ---- three 7x7 stems fused into one Cout=160 problem on the photo (networks.py:1218-1243) ----
  Raw stem = raw(Bp, 256, 256, 160, true);
  AP_TRY(wait_input(0));
  if (prec == AP_PREC_FP32_SIMT) {
    AP_TRY(conv_thin(geom_conv(Bp, 256, 3, 160, 7, 1, 3, 1), in.input, 1, 3, h->w_stem, stem));
  } else if (ph == PH_EXEC) {
    AP_TRY(launch_stem_umma(io, h->w_stem_img, stem, Bp, prec == AP_PREC_FP32X3 ? 3 : 1, st));
    AP_TRY(mark(CL_STEM, 2.0 * Bp * 65536.0 * 147 * 160));
  }
  tap_raw("tri00", stem, 0, 32, 1);
  tap_raw("tri10", stem, 32, 64, 1);
  tap_raw("tri20", stem, 96, 64, 1);

  AP_TRY(order_after(s1, main_st));

  // ---- branch 1 (main stream): warp L0 -> tri01 -> tri02 ----
  Act W0 = act(B, 256, 256, 64, 0, afmt);
  AP_TRY(wait_input(1));
  AP_TRY(warp(stem, 0, 32, 0, W0, 0));
  tap_act("warp0", W0, 0, 64);
  Raw r01 = raw(B, 128, 128, 128, true);
  AP_TRY(conv(geom_conv(B, 256, 64, 128, 3, 2, 1, 0), W0, 0, W("model_tri01.0"), r01, 0));
  tap_raw("tri01", r01, 0, 128, 1);
  Act A01 = act(B, 128, 128, 128, 0, afmt);
  AP_TRY(apply(r01, 0, 128, 1, &A01, 0, 0));
  Raw r02 = raw(B, 64, 64, 256, true);
  AP_TRY(conv(geom_conv(B, 128, 128, 256, 3, 2, 1, 0), A01, 0, W("model_tri02.0"), r02, 0));
  tap_raw("tri02", r02, 0, 256, 1);
  Act MI = act(B, 64, 64, 768, 0, afmt);  // cat[x1, x2, x3] (networks.py:1330) as channel offsets
  AP_TRY(apply(r02, 0, 256, 1, &MI, 0, 0));

  // ---- stems of branches 2 and 3, normalised once: T12 = [tri10 | tri20] (side stream 1) ----
  on(s1);
  Act T12 = act(Bp, 256, 256, 128, 0, afmt);
  AP_TRY(apply(stem, 32, 128, 1, &T12, 0, 0));
  AP_TRY(order_after(s2, s1));

  // ---- branch 2 (side stream 1): tri11 -> warp L1 -> tri12 ----
  Raw r11 = raw(Bp, 128, 128, 64, true);
  AP_TRY(conv(geom_conv(Bp, 256, 64, 64, 3, 2, 1, 0), T12, 0, W("model_tri11.0"), r11, 0));
  tap_raw("tri11", r11, 0, 64, 1);
  Act W1 = act(B, 128, 128, 128, 0, afmt);
  AP_TRY(wait_input(1));
  AP_TRY(warp(r11, 0, 64, 1, W1, 0));
  tap_act("warp1", W1, 0, 128);
  Raw r12 = raw(B, 64, 64, 256, true);
  AP_TRY(conv(geom_conv(B, 128, 128, 256, 3, 2, 1, 0), W1, 0, W("model_tri12.0"), r12, 0));
  tap_raw("tri12", r12, 0, 256, 1);
  AP_TRY(apply(r12, 0, 256, 1, &MI, 256, 0));

  // ---- branch 3 (side stream 2): tri21 -> tri22 -> warp L2 ----
  on(s2);
  Raw r21 = raw(Bp, 128, 128, 128, true);
  AP_TRY(conv(geom_conv(Bp, 256, 64, 128, 3, 2, 1, 0), T12, 64, W("model_tri21.0"), r21, 0));
  tap_raw("tri21", r21, 0, 128, 1);
  Act A21 = act(Bp, 128, 128, 128, 0, afmt);
  AP_TRY(apply(r21, 0, 128, 1, &A21, 0, 0));
  Raw r22 = raw(Bp, 64, 64, 128, true);
  AP_TRY(conv(geom_conv(Bp, 128, 128, 128, 3, 2, 1, 0), A21, 0, W("model_tri22.0"), r22, 0));
  tap_raw("tri22", r22, 0, 128, 1);
  AP_TRY(wait_input(1));
  AP_TRY(warp(r22, 0, 128, 2, MI, 512));
  tap_act("warp2", MI, 512, 256);
  on(main_st);
  AP_TRY(order_after(main_st, s1));
  AP_TRY(order_after(main_st, s2));

  // ---- merge + 9 residual blocks (networks.py:1251,1330,1333-1337, 2303-2421): stream order; the shortcut conv of a
  // ResnetBlock2 only depends on the block input and runs on a side stream under conv_block.1's apply pass ----
  // model_tri_merge keeps its bias and has no InstanceNorm (networks.py:1251): no statistics, bias mode
  Raw rM = keep ? raw(B, 64, 64, 256, false) : trunk_raw[0];
  AP_TRY(conv(geom_conv(B, 64, 768, 256, 3, 1, 1, 0), MI, 0, W("model_tri_merge"), rM, 0));
  AP_TRY(apply(rM, 0, 256, 0, &Xb[0], 0, 1, h->b_merge, nullptr, nullptr, xres[0], nullptr));
  if (res_from_act) tap_act("merge", Xb[0], 0, 256);
  else tap_f32("merge", xres[0], B, 64, 64, 256);

  AP_TRY(order_after(main_st, s3));

  for (int i = 0; i < 9; ++i) {
    const std::string b = "model2." + std::to_string(i);
    const bool b2 = (i % 3) == 0;
    const Act& src = Xb[i];
    const int cin = b2 ? 288 : 256;
    const Act* dst = &Xb[i + 1];
    const int dst_halo = (i == 8) ? 0 : 1;
    Raw rs;
    Raw r1 = keep ? raw(B, 64, 64, 256, true) : with_stats(trunk_raw[0]);
    AP_TRY(conv(geom_conv(B, 64, cin, 256, 3, 1, 1, 1), src, 0, W(b + ".conv_block.1"), r1, 0));
    if (b2) {
      rs = keep ? raw(B, 64, 64, 256, true) : with_stats(trunk_raw[1]);
      AP_TRY(order_after(s1, main_st));
      on(s1);
      AP_TRY(conv(geom_conv(B, 64, cin, 256, 3, 1, 1, 0), src, 0, W(b + ".shortcut.0"), rs, 0));
      on(main_st);
    }
    AP_TRY(apply(r1, 0, 256, 1, &T, 0, 1));
    Raw r2 = keep ? raw(B, 64, 64, 256, true) : with_stats(trunk_raw[2]);
    AP_TRY(conv(geom_conv(B, 64, 256, 256, 3, 1, 1, 1), T, 0, W(b + ".conv_block.5"), r2, 0));
    if (b2) AP_TRY(order_after(main_st, s1));
    if (b2) AP_TRY(apply(r2, 0, 256, 0, dst, 0, dst_halo, nullptr, &rs, nullptr, xres[i + 1], nullptr));
    else if (res_from_act) AP_TRY(apply(r2, 0, 256, 0, dst, 0, dst_halo, nullptr, nullptr, nullptr, nullptr, &src));
    else AP_TRY(apply(r2, 0, 256, 0, dst, 0, dst_halo, nullptr, nullptr, xres[i], xres[i + 1], nullptr));
    const std::string tn = "block" + std::to_string(i);
    if (res_from_act) tap_act(tn.c_str(), *dst, 0, 256);
    else tap_f32(tn.c_str(), xres[i + 1], B, 64, 64, 256);
  }

  // ---- decoder (networks.py:1268-1279): two ConvT (phase-packed), then the 7x7 output conv.  Every encoder buffer is
  // dead by now (stream order: all of them were consumed before the merge conv finished) ----
  if (!keep) off = encoder_begin;
  Raw ru0 = raw(B, 128, 128, 128, true);
  AP_TRY(convT(Xb[9], W("model3.0"), ru0, 64, 256, 128));
  tap_raw("up0", ru0, 0, 128, 1);
  Act U0 = act(B, 128, 128, 128, 0, afmt);
  AP_TRY(apply(ru0, 0, 128, 1, &U0, 0, 0));
  Raw ru1 = raw(B, 256, 256, 64, true);
  AP_TRY(convT(U0, W("model3.3"), ru1, 128, 128, 64));
  tap_raw("up1", ru1, 0, 64, 1);
  if (ph == PH_EXEC) {
    OutConvP p{};
    p.raw = ru1.p; p.stats = ru1.stats; p.w = h->w_out; p.bias = h->b_out; p.io = io; p.B = B; p.onc = h->onc;
    if (h->w_out_img && h->out_umma) AP_TRY(launch_out_umma(p, h->w_out_img, st));
    else AP_TRY(launch_out_conv(p, st));
    AP_TRY(mark(CL_OUT, 2.0 * B * 256.0 * 256.0 * 64 * 49 * h->onc));
  }
  return AP_OK;
}

static int get_plan(ap_netg* h, int B, bool shared_photo, Plan** out) {
  const int key = B * 4 + (h->keep_all ? 2 : 0) + (shared_photo ? 1 : 0);
  auto it = h->plans.find(key);
  if (it != h->plans.end()) { it->second->last_use = ++h->use_clock; *out = it->second; return AP_OK; }
  // every plan owns a workspace arena (about 0.25 GB per frame of batch in the fp32-accurate mode): keep the most
  // recently used AP_MAX_PLANS batch shapes.  Eviction frees device memory, so it waits for the device first.
  while (h->plans.size() >= AP_MAX_PLANS) {
    auto victim = h->plans.end();
    for (auto p = h->plans.begin(); p != h->plans.end(); ++p)
      if (victim == h->plans.end() || p->second->last_use < victim->second->last_use) victim = p;
    if (cudaDeviceSynchronize() != cudaSuccess) { set_error("device synchronisation before plan eviction failed"); return AP_ERR_CUDA; }
    if (h->last_plan == victim->second) h->last_plan = nullptr;
    delete victim->second;
    h->plans.erase(victim);
  }
  Plan* pl = new Plan();
  pl->B = B;
  pl->shared_photo = shared_photo;
  pl->keep_all = h->keep_all;
  Inputs none{};
  Runner rs{h, pl, PH_SIZE, nullptr};
  int rc = rs.run(none);
  if (rc != AP_OK) { delete pl; return rc; }
  pl->arena_bytes = align_up(rs.peak, 1024);
  pl->sarena_bytes = align_up(rs.soff, 256);
  if (cudaMalloc(&pl->arena, pl->arena_bytes) != cudaSuccess || cudaMalloc(&pl->sarena, pl->sarena_bytes) != cudaSuccess ||
      cudaMalloc(&pl->d_io, sizeof(IoPtrs)) != cudaSuccess) {
    set_error("workspace allocation of %zu bytes for B=%d failed: %s", pl->arena_bytes, B,
              cudaGetErrorString(cudaGetLastError()));
    delete pl;
    return AP_ERR_CUDA;
  }
  // halos of zero-padded / never-written regions must read as zero
  if (cudaMemset(pl->arena, 0, pl->arena_bytes) != cudaSuccess || cudaMemset(pl->d_io, 0, sizeof(IoPtrs)) != cudaSuccess) {
    delete pl;
    set_error("memset failed");
    return AP_ERR_CUDA;
  }
  Runner rb{h, pl, PH_BUILD, nullptr};
  rc = rb.run(none);
  if (rc != AP_OK) { delete pl; return rc; }
  for (cudaStream_t& s_ : pl->side)
    if (cudaStreamCreateWithFlags(&s_, cudaStreamNonBlocking) != cudaSuccess) {
      set_error("side stream creation failed");
      delete pl;
      return AP_ERR_CUDA;
    }
  pl->last_use = ++h->use_clock;
  h->plans[key] = pl;
  *out = pl;
  return AP_OK;
}

// The launch sequence of a plan captured into a CUDA graph (on a stream of the plan: the caller's stream may be the
// legacy default stream, which cannot capture).  Nothing executes during capture.
static int capture_plan(ap_netg* h, Plan* pl) {
  if (!pl->cap_stream) AP_CUDA(cudaStreamCreateWithFlags(&pl->cap_stream, cudaStreamNonBlocking));
  const int64_t before = launches_get();
  AP_CUDA(cudaStreamBeginCapture(pl->cap_stream, cudaStreamCaptureModeRelaxed));
  Inputs none{};
  Runner rx{h, pl, PH_EXEC, pl->cap_stream};
  const int rc = rx.run(none);
  cudaGraph_t graph = nullptr;
  const cudaError_t e = cudaStreamEndCapture(pl->cap_stream, &graph);
  if (rc != AP_OK) {
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
    return rc;
  }
  if (e != cudaSuccess || graph == nullptr) {
    set_error("CUDA graph capture of the forward failed: %s", cudaGetErrorString(e));
    cudaGetLastError();
    return AP_ERR_CUDA;
  }
  const cudaError_t e2 = cudaGraphInstantiate(&pl->gexec, graph, 0);
  cudaGraphDestroy(graph);
  if (e2 != cudaSuccess) {
    pl->gexec = nullptr;
    set_error("CUDA graph instantiation failed: %s", cudaGetErrorString(e2));
    return AP_ERR_CUDA;
  }
  pl->graph_launches = launches_get() - before;
  return AP_OK;
}

}  // namespace ap

// =================================================================================================
// C ABI
// =================================================================================================
extern "C" {

const char* ap_last_error(void) { return ap::g_err; }
const char* ap_version(void) { return "apnetg 0.1 sm_100a"; }

int ap_netg_create(ap_netg** handle, int output_nc, int precision, int device) {
  AP_REQUIRE(handle != nullptr, AP_ERR_INVALID, "null handle pointer");
  AP_REQUIRE(output_nc == 1 || output_nc == 3, AP_ERR_UNSUPPORTED, "output_nc=%d not supported (1 or 3)", output_nc);
  AP_REQUIRE(precision >= 0 && precision <= 2, AP_ERR_INVALID, "precision=%d", precision);
  int ndev = 0;
  AP_CUDA(cudaGetDeviceCount(&ndev));
  AP_REQUIRE(device >= 0 && device < ndev, AP_ERR_INVALID, "device %d of %d", device, ndev);
  AP_CUDA(cudaSetDevice(device));
  cudaDeviceProp prop;
  AP_CUDA(cudaGetDeviceProperties(&prop, device));
  AP_REQUIRE(prop.major == 10, AP_ERR_UNSUPPORTED, "libapnetg is built for sm_100a only; device %d is sm_%d%d", device,
             prop.major, prop.minor);
  if (precision != AP_PREC_FP32_SIMT) AP_TRY(umma_init());
  ap_netg* h = new ap_netg();
  h->onc = output_nc; h->prec = precision; h->device = device;
  const char* ov = getenv("AP_NETG_OVERLAP");
  h->overlap = !(ov && ov[0] == '0');
  const char* cp = getenv("AP_NETG_CONVT_PACKED");
  // the phase-packed form runs on the CTA-pair kernel only: without pairs (AP_NETG_CTA_PAIR=0, or a device that cannot
  // co-schedule them) the four per-phase convs are used
  h->convt_packed = !(cp && cp[0] == '0') && precision != AP_PREC_FP32_SIMT && umma_pairs_available();
  const char* gr = getenv("AP_NETG_GRAPH");
  h->graphs = !(gr && gr[0] == '0');
  const char* ou = getenv("AP_NETG_OUT_UMMA");
  h->out_umma = !(ou && ou[0] == '0');
  *handle = h;
  return AP_OK;
}

static void free_weights(ap_netg* h) {
  for (void* p : h->owned) cudaFree(p);
  h->owned.clear();
  h->w.clear();
  h->w_stem = h->w_out = h->b_merge = h->b_out = nullptr;
  h->w_stem_img = nullptr;
  h->w_out_img = nullptr;
  h->loaded = false;
}

int ap_netg_destroy(ap_netg* h) {
  if (!h) return AP_OK;
  cudaSetDevice(h->device);
  for (auto& kv : h->plans) delete kv.second;
  for (cudaEvent_t e : h->ev) cudaEventDestroy(e);
  free_weights(h);
  delete h;
  return AP_OK;
}

int ap_netg_load_weights(ap_netg* h, int n, const char* const* names, const float* const* ptrs, const int64_t* shapes,
                         int on_device, void* cuda_stream) {
  AP_REQUIRE(h && names && ptrs && shapes, AP_ERR_INVALID, "null argument");
  AP_CUDA(cudaSetDevice(h->device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  std::map<std::string, int> idx;
  for (int i = 0; i < n; ++i) idx[names[i]] = i;
  const std::vector<LayerSpec> specs = layer_specs(h->onc);
  AP_REQUIRE(n == (int)specs.size() * 2, AP_ERR_INVALID, "state_dict has %d tensors, expected %d", n, (int)specs.size() * 2);
  // strict key / shape check before touching anything (load_state_dict(strict=True) semantics)
  for (const LayerSpec& s : specs) {
    auto wi = idx.find(s.name + ".weight"), bi = idx.find(s.name + ".bias");
    AP_REQUIRE(wi != idx.end() && bi != idx.end(), AP_ERR_INVALID, "missing key %s.weight/.bias", s.name.c_str());
    const int64_t* ws = shapes + 4 * wi->second;
    const int64_t d0 = s.transposed ? s.cin : s.cout, d1 = s.transposed ? s.cout : s.cin;
    AP_REQUIRE(ws[0] == d0 && ws[1] == d1 && ws[2] == s.k && ws[3] == s.k, AP_ERR_INVALID,
               "size mismatch for %s.weight: got [%lld,%lld,%lld,%lld], expected [%lld,%lld,%d,%d]", s.name.c_str(),
               (long long)ws[0], (long long)ws[1], (long long)ws[2], (long long)ws[3], (long long)d0, (long long)d1, s.k, s.k);
    AP_REQUIRE(shapes[4 * bi->second] == s.cout, AP_ERR_INVALID, "size mismatch for %s.bias", s.name.c_str());
  }
  free_weights(h);
  auto dalloc = [&](size_t bytes, void** p) -> int {
    AP_CUDA(cudaMalloc(p, bytes));
    h->owned.push_back(*p);
    return AP_OK;
  };
  std::vector<void*> staging;
  auto dev_src = [&](const float* src, size_t elems, const float** out) -> int {
    if (on_device) { *out = src; return AP_OK; }
    void* d = nullptr;
    AP_CUDA(cudaMalloc(&d, elems * sizeof(float)));
    staging.push_back(d);
    AP_CUDA(cudaMemcpyAsync(d, src, elems * sizeof(float), cudaMemcpyHostToDevice, st));
    *out = (const float*)d;
    return AP_OK;
  };
  int rc = AP_OK;
  AP_TRY(dalloc((size_t)49 * 3 * 160 * 4, (void**)&h->w_stem));
  if (h->prec != AP_PREC_FP32_SIMT) AP_TRY(dalloc(stem_umma_weight_bytes(), (void**)&h->w_stem_img));
  AP_TRY(dalloc((size_t)h->onc * 49 * 64 * 4, (void**)&h->w_out));
  if (h->prec != AP_PREC_FP32_SIMT) AP_TRY(dalloc(out_umma_weight_bytes(h->onc), (void**)&h->w_out_img));
  AP_TRY(dalloc(256 * 4, (void**)&h->b_merge));
  AP_TRY(dalloc(h->onc * 4, (void**)&h->b_out));
  for (const LayerSpec& s : specs) {
    const size_t elems = (size_t)s.cout * s.cin * s.k * s.k;
    const float* src = nullptr;
    rc = dev_src(ptrs[idx[s.name + ".weight"]], elems, &src);
    if (rc != AP_OK) break;
    const int stem_off = s.name == "model_tri00.1" ? 0 : (s.name == "model_tri10.1" ? 32 : (s.name == "model_tri20.1" ? 96 : -1));
    if (stem_off >= 0) {
      rc = launch_pack_weights(src, s.cout, 3, 7, 0, h->w_stem, 160, stem_off, nullptr, nullptr, st);
      if (rc == AP_OK && h->w_stem_img) rc = launch_pack_stem_umma(src, s.cout, stem_off, h->w_stem_img, st);
    }
    else if (s.name == "model3.7") {
      rc = launch_pack_out_weights(src, h->onc, h->w_out, st);
      if (rc == AP_OK && h->w_out_img) rc = launch_pack_out_umma(src, h->onc, h->w_out_img, st);
    }
    else {
      LayerW lw;
      lw.cout = s.cout; lw.cin = s.cin; lw.k = s.k;
      const bool thin = s.name.rfind("model_landmark_trans", 0) == 0;
      if (thin) {
        // tiny layers whose weights ride in the kernel parameters: keep a packed host copy
        std::vector<float> tmp(elems);
        const float* hsrc = ptrs[idx[s.name + ".weight"]];
        if (on_device) {
          if (cudaMemcpy(tmp.data(), hsrc, elems * sizeof(float), cudaMemcpyDeviceToHost) != cudaSuccess) {
            set_error("copy of %s to the host failed", s.name.c_str());
            rc = AP_ERR_CUDA;
            break;
          }
          hsrc = tmp.data();
        }
        lw.host.resize(elems);
        for (int sl = 0; sl < 9; ++sl)
          for (int ci = 0; ci < s.cin; ++ci)
            for (int co = 0; co < s.cout; ++co)
              lw.host[((size_t)sl * s.cin + ci) * s.cout + co] = hsrc[((size_t)co * s.cin + ci) * 9 + sl];
        h->w[s.name] = lw;
        continue;
      }
      const bool simt = h->prec == AP_PREC_FP32_SIMT;
      if (simt) rc = dalloc(elems * 4, (void**)&lw.simt);
      if (rc == AP_OK && !simt) rc = dalloc(elems * 2, (void**)&lw.hi);
      if (rc == AP_OK && !simt && h->prec == AP_PREC_FP32X3) rc = dalloc(elems * 2, (void**)&lw.lo);
      if (rc == AP_OK)
        rc = launch_pack_weights(src, s.cout, s.cin, s.k, s.transposed ? 1 : 0, lw.simt, s.cout, 0, lw.hi, lw.lo, st);
      if (rc == AP_OK && s.transposed && !simt)
        rc = make_packed_convT(src, s.cin, s.cout, h->prec == AP_PREC_FP32X3, st, &h->owned, &lw.packT);
      h->w[s.name] = lw;
    }
    if (rc != AP_OK) break;
    const float* bsrc = nullptr;
    if (s.name == "model_tri_merge" || s.name == "model3.7") {
      // the only two biases that reach the output (every other conv feeds an affine-less InstanceNorm)
      float* dst = s.name == "model3.7" ? h->b_out : h->b_merge;
      rc = dev_src(ptrs[idx[s.name + ".bias"]], s.cout, &bsrc);
      if (rc == AP_OK && cudaMemcpyAsync(dst, bsrc, s.cout * 4, cudaMemcpyDeviceToDevice, st) != cudaSuccess) {
        set_error("bias copy failed");
        rc = AP_ERR_CUDA;
      }
      if (rc != AP_OK) break;
    }
  }
  cudaError_t e = cudaStreamSynchronize(st);
  for (void* p : staging) cudaFree(p);
  if (rc != AP_OK) { free_weights(h); return rc; }
  if (e != cudaSuccess) { set_error("weight packing failed: %s", cudaGetErrorString(e)); free_weights(h); return AP_ERR_CUDA; }
  // plans hold tensor maps that point at the old weight buffers
  for (auto& kv : h->plans) delete kv.second;
  h->plans.clear();
  h->last_plan = nullptr;
  h->loaded = true;
  return AP_OK;
}

int ap_netg_workspace_bytes(ap_netg* h, int B, size_t* bytes) {
  AP_REQUIRE(h && bytes && B >= 1, AP_ERR_INVALID, "bad argument");
  Plan pl;
  pl.B = B;
  Inputs none{};
  pl.keep_all = h->keep_all;
  Runner rs{h, &pl, PH_SIZE, nullptr};
  // SIZE phase never dereferences weights
  AP_REQUIRE(h->loaded, AP_ERR_STATE, "load_weights must be called before workspace_bytes");
  AP_TRY(rs.run(none));
  *bytes = align_up(rs.peak, 1024) + align_up(rs.soff, 256);
  return AP_OK;
}

static int forward_impl(ap_netg* h, int B, const Inputs& in, cudaStream_t st, const cudaEvent_t* in_ready,
                        bool shared_photo = false) {
  Plan* pl = nullptr;
  AP_TRY(get_plan(h, B, shared_photo, &pl));
  const int64_t before = launches_get();
  const IoPtrs io{in.input, in.land1, in.land2, in.motion, in.flow, in.ifmask, in.out};
  // CUDA-graph replay: device-pointer calls of the tensor-core modes.  The host-buffer entry point orders its kernels
  // against three upload events of the call and the profiler records an event per launch: both launch on streams.
  const bool graph = h->graphs && !h->profiling && in_ready == nullptr && h->prec != AP_PREC_FP32_SIMT;
  if (graph && pl->gexec == nullptr) AP_TRY(capture_plan(h, pl));
  AP_TRY(launch_set_io(pl->d_io, io, st));
  if (graph) {
    AP_CUDA(cudaGraphLaunch(pl->gexec, st));
    h->last_launches = pl->graph_launches + 1;
    h->last_plan = pl;
    return AP_OK;
  }
  Runner rx{h, pl, PH_EXEC, st};
  rx.in_ready = in_ready;
  if (h->profiling) {
    if (h->ev.empty()) {
      cudaEvent_t e;
      AP_CUDA(cudaEventCreate(&e));
      h->ev.push_back(e);
    }
    h->ev_used = 0;
    h->ev_class.clear();
    h->ev_flops.clear();
    AP_CUDA(cudaEventRecord(h->ev[0], st));
  }
  AP_TRY(rx.run(in));
  h->last_launches = launches_get() - before;
  h->last_plan = pl;
  return AP_OK;
}

int ap_netg_forward(ap_netg* h, int B, const float* input, const float* land1, const float* land2, const float* motion,
                    const float* flow, const float* ifmask, float* out, void* cuda_stream) {
  AP_REQUIRE(h != nullptr, AP_ERR_INVALID, "null handle");
  AP_REQUIRE(h->loaded, AP_ERR_STATE, "forward before load_weights");
  AP_REQUIRE(B >= 1, AP_ERR_INVALID, "B=%d", B);
  AP_REQUIRE(input && land1 && land2 && motion && flow && ifmask && out, AP_ERR_INVALID, "null tensor pointer");
  AP_CUDA(cudaSetDevice(h->device));
  Inputs in{input, land1, land2, motion, flow, ifmask, out};
  return forward_impl(h, B, in, (cudaStream_t)cuda_stream, nullptr);
}

int ap_netg_forward_shared_photo(ap_netg* h, int B, const float* input, const float* land1, const float* land2,
                                 const float* motion, const float* flow, const float* ifmask, float* out,
                                 void* cuda_stream) {
  AP_REQUIRE(h != nullptr, AP_ERR_INVALID, "null handle");
  AP_REQUIRE(h->loaded, AP_ERR_STATE, "forward before load_weights");
  AP_REQUIRE(B >= 1, AP_ERR_INVALID, "B=%d", B);
  AP_REQUIRE(input && land1 && land2 && motion && flow && ifmask && out, AP_ERR_INVALID, "null tensor pointer");
  AP_CUDA(cudaSetDevice(h->device));
  Inputs in{input, land1, land2, motion, flow, ifmask, out};
  return forward_impl(h, B, in, (cudaStream_t)cuda_stream, nullptr, true);
}

// Host-buffer call into staging slot `slot`: uploads on the copy stream in the order the forward consumes them (photo ->
// stems, motion/flow/ifmask -> warps, landmarks -> landmark branch; the compute stream waits per group), forward on `st`,
// download of the frames on a second copy stream.  Returns without waiting; slot.ev_done marks the frames' arrival.
static int host_enqueue(ap_netg* h, Plan* pl, int B, const float* const src[6], float* out, cudaStream_t st, int slot) {
  const size_t px = (size_t)B * 256 * 256;
  const size_t sz[6] = {px * 3, px, px, px * 2, px * 2, px};
  if (!pl->copy_stream) {
    AP_CUDA(cudaStreamCreateWithFlags(&pl->copy_stream, cudaStreamNonBlocking));
    AP_CUDA(cudaStreamCreateWithFlags(&pl->d2h_stream, cudaStreamNonBlocking));
  }
  Plan::HostSlot& sl = pl->hs[slot];
  if (!sl.h_out) {
    for (int i = 0; i < 3; ++i) AP_CUDA(cudaEventCreateWithFlags(&sl.ev_in[i], cudaEventDisableTiming));
    AP_CUDA(cudaEventCreateWithFlags(&sl.ev_computed, cudaEventDisableTiming));
    AP_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
    for (int i = 0; i < 6; ++i) AP_CUDA(cudaMalloc(&sl.h_in[i], sz[i] * sizeof(float)));
    AP_CUDA(cudaMalloc(&sl.h_out, px * h->onc * sizeof(float)));
  }
  if (sl.busy) AP_CUDA(cudaEventSynchronize(sl.ev_done));  // the call that used this slot two calls ago has fully drained
  sl.busy = true;
  const int order[6] = {0, 3, 4, 5, 1, 2};
  for (int k = 0; k < 6; ++k) {
    const int i = order[k];
    AP_CUDA(cudaMemcpyAsync(sl.h_in[i], src[i], sz[i] * sizeof(float), cudaMemcpyHostToDevice, pl->copy_stream));
    if (k == 0) AP_CUDA(cudaEventRecord(sl.ev_in[0], pl->copy_stream));
    if (k == 3) AP_CUDA(cudaEventRecord(sl.ev_in[1], pl->copy_stream));
    if (k == 5) AP_CUDA(cudaEventRecord(sl.ev_in[2], pl->copy_stream));
  }
  Inputs in{sl.h_in[0], sl.h_in[1], sl.h_in[2], sl.h_in[3], sl.h_in[4], sl.h_in[5], sl.h_out};
  AP_TRY(forward_impl(h, B, in, st, sl.ev_in));
  AP_CUDA(cudaEventRecord(sl.ev_computed, st));
  AP_CUDA(cudaStreamWaitEvent(pl->d2h_stream, sl.ev_computed, 0));
  AP_CUDA(cudaMemcpyAsync(out, sl.h_out, px * h->onc * sizeof(float), cudaMemcpyDeviceToHost, pl->d2h_stream));
  AP_CUDA(cudaEventRecord(sl.ev_done, pl->d2h_stream));
  return AP_OK;
}

int ap_netg_forward_host(ap_netg* h, int B, const float* input, const float* land1, const float* land2,
                         const float* motion, const float* flow, const float* ifmask, float* out, void* cuda_stream) {
  AP_REQUIRE(h != nullptr && h->loaded, AP_ERR_STATE, "forward before load_weights");
  AP_REQUIRE(B >= 1 && input && land1 && land2 && motion && flow && ifmask && out, AP_ERR_INVALID, "bad argument");
  AP_CUDA(cudaSetDevice(h->device));
  Plan* pl = nullptr;
  AP_TRY(get_plan(h, B, false, &pl));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const float* src[6] = {input, land1, land2, motion, flow, ifmask};
  if (h->graphs && !h->profiling && h->prec != AP_PREC_FP32_SIMT && B <= 4) {
    // small batches (the reference's own call is batch size 1, Module2/test.py:42): the uploads are tens of
    // microseconds, launch overhead is what matters -- copy in stream order and replay the plan's CUDA graph
    const size_t px = (size_t)B * 256 * 256;
    const size_t sz[6] = {px * 3, px, px, px * 2, px * 2, px};
    Plan::HostSlot& sl = pl->hs[0];
    if (sl.busy) AP_CUDA(cudaEventSynchronize(sl.ev_done));
    if (!sl.h_out) {
      for (int i = 0; i < 3; ++i) AP_CUDA(cudaEventCreateWithFlags(&sl.ev_in[i], cudaEventDisableTiming));
      AP_CUDA(cudaEventCreateWithFlags(&sl.ev_computed, cudaEventDisableTiming));
      AP_CUDA(cudaEventCreateWithFlags(&sl.ev_done, cudaEventDisableTiming));
      for (int i = 0; i < 6; ++i) AP_CUDA(cudaMalloc(&sl.h_in[i], sz[i] * sizeof(float)));
      AP_CUDA(cudaMalloc(&sl.h_out, px * h->onc * sizeof(float)));
    }
    const int order[6] = {0, 3, 4, 5, 1, 2};
    for (int k = 0; k < 6; ++k)
      AP_CUDA(cudaMemcpyAsync(sl.h_in[order[k]], src[order[k]], sz[order[k]] * sizeof(float), cudaMemcpyHostToDevice, st));
    Inputs in{sl.h_in[0], sl.h_in[1], sl.h_in[2], sl.h_in[3], sl.h_in[4], sl.h_in[5], sl.h_out};
    AP_TRY(forward_impl(h, B, in, st, nullptr));
    AP_CUDA(cudaMemcpyAsync(out, sl.h_out, px * h->onc * sizeof(float), cudaMemcpyDeviceToHost, st));
    AP_CUDA(cudaStreamSynchronize(st));
    return AP_OK;
  }
  AP_TRY(host_enqueue(h, pl, B, src, out, st, 0));
  AP_CUDA(cudaEventSynchronize(pl->hs[0].ev_done));
  pl->hs[0].busy = false;
  return AP_OK;
}

int ap_netg_forward_host_async(ap_netg* h, int B, const float* input, const float* land1, const float* land2,
                               const float* motion, const float* flow, const float* ifmask, float* out, void* cuda_stream) {
  AP_REQUIRE(h != nullptr && h->loaded, AP_ERR_STATE, "forward before load_weights");
  AP_REQUIRE(B >= 1 && input && land1 && land2 && motion && flow && ifmask && out, AP_ERR_INVALID, "bad argument");
  AP_CUDA(cudaSetDevice(h->device));
  Plan* pl = nullptr;
  AP_TRY(get_plan(h, B, false, &pl));
  const float* src[6] = {input, land1, land2, motion, flow, ifmask};
  return host_enqueue(h, pl, B, src, out, (cudaStream_t)cuda_stream, (int)(pl->host_calls++ & 1));
}

int ap_netg_host_sync(ap_netg* h) {
  AP_REQUIRE(h != nullptr, AP_ERR_INVALID, "null handle");
  AP_CUDA(cudaSetDevice(h->device));
  for (auto& kv : h->plans)
    for (Plan::HostSlot& sl : kv.second->hs)
      if (sl.busy) {
        AP_CUDA(cudaEventSynchronize(sl.ev_done));
        sl.busy = false;
      }
  return AP_OK;
}

int ap_device_enable_peer_access(int device, int peer_device) {
  int ndev = 0;
  AP_CUDA(cudaGetDeviceCount(&ndev));
  AP_REQUIRE(device >= 0 && device < ndev && peer_device >= 0 && peer_device < ndev, AP_ERR_INVALID, "devices %d, %d of %d",
             device, peer_device, ndev);
  if (device == peer_device) return AP_OK;
  int can = 0;
  AP_CUDA(cudaDeviceCanAccessPeer(&can, device, peer_device));
  AP_REQUIRE(can != 0, AP_ERR_UNSUPPORTED, "device %d cannot access the memory of device %d", device, peer_device);
  int cur = 0;
  AP_CUDA(cudaGetDevice(&cur));
  AP_CUDA(cudaSetDevice(device));
  const cudaError_t e = cudaDeviceEnablePeerAccess(peer_device, 0);
  cudaSetDevice(cur);
  if (e == cudaErrorPeerAccessAlreadyEnabled) { cudaGetLastError(); return AP_OK; }
  if (e != cudaSuccess) { set_error("cudaDeviceEnablePeerAccess(%d -> %d): %s", device, peer_device, cudaGetErrorString(e)); return AP_ERR_CUDA; }
  return AP_OK;
}

// ---- one buffer on the GPU that owns the clip, mapped by every process of the box (CUDA IPC + NVLink peer access) ----
int ap_peer_alloc(int device, size_t bytes, void** ptr, unsigned char* handle64) {
  AP_REQUIRE(ptr && handle64 && bytes > 0, AP_ERR_INVALID, "bad argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handles are 64 bytes");
  int cur = 0;
  AP_CUDA(cudaGetDevice(&cur));
  AP_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  cudaError_t e = cudaMalloc(&p, bytes);
  cudaIpcMemHandle_t hdl;
  if (e == cudaSuccess) e = cudaIpcGetMemHandle(&hdl, p);
  cudaSetDevice(cur);
  if (e != cudaSuccess) {
    if (p) cudaFree(p);
    set_error("peer buffer of %zu bytes on device %d: %s", bytes, device, cudaGetErrorString(e));
    cudaGetLastError();
    return AP_ERR_CUDA;
  }
  memcpy(handle64, &hdl, 64);
  *ptr = p;
  return AP_OK;
}

int ap_peer_open(int device, const unsigned char* handle64, void** ptr) {
  AP_REQUIRE(ptr && handle64, AP_ERR_INVALID, "bad argument");
  cudaIpcMemHandle_t hdl;
  memcpy(&hdl, handle64, 64);
  int cur = 0;
  AP_CUDA(cudaGetDevice(&cur));
  // opened in the context of the ACCESSING device: the owner's memory is mapped into this device's address space and
  // peer access to the owner is enabled on the way (what NCCL does for its own NVLink buffers)
  AP_CUDA(cudaSetDevice(device));
  void* p = nullptr;
  const cudaError_t e = cudaIpcOpenMemHandle(&p, hdl, cudaIpcMemLazyEnablePeerAccess);
  cudaSetDevice(cur);
  if (e != cudaSuccess) {
    set_error("cudaIpcOpenMemHandle on device %d: %s", device, cudaGetErrorString(e));
    cudaGetLastError();
    return AP_ERR_CUDA;
  }
  *ptr = p;
  return AP_OK;
}

int ap_peer_close(void* ptr) {
  if (ptr && cudaIpcCloseMemHandle(ptr) != cudaSuccess) { cudaGetLastError(); set_error("cudaIpcCloseMemHandle failed"); return AP_ERR_CUDA; }
  return AP_OK;
}

int ap_peer_free(void* ptr) {
  if (ptr && cudaFree(ptr) != cudaSuccess) { cudaGetLastError(); set_error("cudaFree of a peer buffer failed"); return AP_ERR_CUDA; }
  return AP_OK;
}

int ap_netg_last_launch_count(ap_netg* h, int64_t* count) {
  AP_REQUIRE(h && count, AP_ERR_INVALID, "null argument");
  *count = h->last_launches;
  return AP_OK;
}

int ap_netg_set_profiling(ap_netg* h, int enable) {
  AP_REQUIRE(h != nullptr, AP_ERR_INVALID, "null handle");
  h->profiling = enable != 0;
  h->ev_used = 0;
  h->ev_class.clear();
  h->ev_flops.clear();
  return AP_OK;
}

int ap_netg_set_option(ap_netg* h, const char* name, int value) {
  AP_REQUIRE(h && name, AP_ERR_INVALID, "null argument");
  const std::string n(name);
  if (n == "graphs") h->graphs = value != 0;
  else if (n == "overlap") {
    // the captured graphs hold the stream structure they were captured with
    if (h->overlap != (value != 0)) {
      AP_CUDA(cudaSetDevice(h->device));
      AP_CUDA(cudaDeviceSynchronize());
      for (auto& kv : h->plans) delete kv.second;
      h->plans.clear();
      h->last_plan = nullptr;
    }
    h->overlap = value != 0;
  } else if (n == "keep_intermediates") h->keep_all = value != 0;
  else { set_error("unknown option '%s' (graphs, overlap, keep_intermediates)", name); return AP_ERR_INVALID; }
  return AP_OK;
}

int ap_netg_get_profile(ap_netg* h, int max_classes, double* ms, int64_t* launches, double* flops, int* n_classes) {
  AP_REQUIRE(h && ms && launches && flops && n_classes, AP_ERR_INVALID, "null argument");
  AP_REQUIRE(h->profiling && h->ev_used > 0, AP_ERR_STATE, "no profiled forward recorded");
  AP_CUDA(cudaSetDevice(h->device));
  AP_CUDA(cudaEventSynchronize(h->ev[h->ev_used]));
  const int nc = max_classes < AP_NCLASS ? max_classes : AP_NCLASS;
  for (int c = 0; c < nc; ++c) { ms[c] = 0.0; launches[c] = 0; flops[c] = 0.0; }
  for (size_t i = 0; i < h->ev_used; ++i) {
    float t = 0.f;
    AP_CUDA(cudaEventElapsedTime(&t, h->ev[i], h->ev[i + 1]));
    const int c = h->ev_class[i];
    if (c < nc) { ms[c] += t; launches[c] += 1; flops[c] += h->ev_flops[i]; }
  }
  *n_classes = nc;
  return AP_OK;
}

int ap_netg_debug_read(ap_netg* h, const char* tap, float* dst, size_t cap, int64_t* shape4, void* cuda_stream) {
  AP_REQUIRE(h && tap && dst, AP_ERR_INVALID, "null argument");
  AP_REQUIRE(h->last_plan != nullptr, AP_ERR_STATE, "debug_read before any forward");
  auto it = h->last_plan->taps.find(tap);
  AP_REQUIRE(it != h->last_plan->taps.end(), AP_ERR_INVALID, "unknown tap '%s'", tap);
  const TapRec& t = it->second;
  const size_t need = (size_t)t.B * t.C * t.H * t.W;
  AP_REQUIRE(cap >= need, AP_ERR_INVALID, "tap '%s' needs %zu elements, buffer has %zu", tap, need, cap);
  if (shape4) { shape4[0] = t.B; shape4[1] = t.C; shape4[2] = t.H; shape4[3] = t.W; }
  AP_CUDA(cudaSetDevice(h->device));
  ReadP p{t.B, t.H, t.W, t.C, t.fmt, t.p0, t.p1, t.sC, t.scoff, t.spad, t.stats, t.stat_C, t.stat_coff, t.relu, dst};
  return launch_read(p, (cudaStream_t)cuda_stream);
}

int ap_conv2d_debug(int impl, int device, int B, int H, int W, int Cin, int Cout, int ksize, int stride, int pad,
                    int pad_mode, int transposed, const float* x, const float* w, float* y, double* stats,
                    void* cuda_stream) {
  AP_REQUIRE(x && w && y, AP_ERR_INVALID, "null argument");
  AP_REQUIRE(H == W, AP_ERR_UNSUPPORTED, "square inputs only");
  AP_REQUIRE(impl >= 0 && impl <= 2, AP_ERR_INVALID, "impl=%d", impl);
  AP_REQUIRE(!transposed || (ksize == 3 && stride == 2 && pad == 1), AP_ERR_UNSUPPORTED, "transposed: k3 s2 p1 op1 only");
  AP_REQUIRE(impl == AP_PREC_FP32_SIMT || (ksize == 3 && pad == 1), AP_ERR_UNSUPPORTED, "tcgen05 path: 3x3 pad 1 only");
  AP_CUDA(cudaSetDevice(device));
  cudaStream_t st = (cudaStream_t)cuda_stream;
  const bool tc = impl != AP_PREC_FP32_SIMT;
  if (tc) AP_TRY(umma_init());
  const int Ho = transposed ? 2 * H : H / stride;
  Act in;
  in.B = B; in.H = H; in.W = W; in.C = Cin;
  in.pad = (tc && pad_mode == 1) ? 1 : 0;
  in.fmt = impl == AP_PREC_FP32X3 ? FMT_BF16X2 : (impl == AP_PREC_BF16 ? FMT_BF16 : FMT_F32);
  const size_t es = in.fmt == FMT_F32 ? 4 : 2;
  std::vector<void*> tmp;
  auto dalloc = [&](size_t bytes, void** p) -> int {
    AP_CUDA(cudaMalloc(p, bytes));
    tmp.push_back(*p);
    AP_CUDA(cudaMemsetAsync(*p, 0, bytes, st));
    return AP_OK;
  };
  int rc = dalloc(in.elems() * es, &in.p0);
  if (rc == AP_OK && in.fmt == FMT_BF16X2) rc = dalloc(in.elems() * es, &in.p1);
  const size_t welems = (size_t)Cout * Cin * ksize * ksize;
  LayerW lw;
  if (rc == AP_OK && !tc) rc = dalloc(welems * 4, (void**)&lw.simt);
  if (rc == AP_OK && tc) rc = dalloc(welems * 2, (void**)&lw.hi);
  if (rc == AP_OK && impl == AP_PREC_FP32X3) rc = dalloc(welems * 2, (void**)&lw.lo);
  Raw out;
  out.B = B; out.H = Ho; out.W = Ho; out.C = Cout;
  if (rc == AP_OK) rc = dalloc((size_t)B * Ho * Ho * Cout * 4, (void**)&out.p);
  if (rc == AP_OK) rc = dalloc((size_t)B * Cout * 2 * 8, (void**)&out.stats);
  std::vector<UmmaConv*> convs;
  if (rc == AP_OK) rc = launch_nchw_to_act(x, in, st);
  if (rc == AP_OK) rc = launch_pack_weights(w, Cout, Cin, ksize, transposed, lw.simt, Cout, 0, lw.hi, lw.lo, st);
  std::vector<PackedT> packT;
  const char* cp = getenv("AP_NETG_CONVT_PACKED");
  const bool packed = transposed && tc && (Cout == 64 || Cout == 128) && !(cp && cp[0] == '0') && umma_pairs_available();
  if (rc == AP_OK && packed) rc = make_packed_convT(w, Cin, Cout, impl == AP_PREC_FP32X3, st, &tmp, &packT);
  if (packed) {
    for (size_t k = 0; k < packT.size() && rc == AP_OK; ++k) {
      UmmaConv* c = nullptr;
      rc = umma_conv_create(&c, geom_convT_packed(B, H, Cin, packT[k]), in, 0, packT[k].hi, packT[k].lo,
                            impl == AP_PREC_FP32X3 ? 3 : 1, out.p, Cout, 0, out.stats, Cout, 0, &packT[k].pk);
      if (rc == AP_OK) { convs.push_back(c); rc = umma_conv_launch(c, st); }
    }
  }
  const int nph = packed ? 0 : (transposed ? 4 : 1);
  for (int ph = 0; ph < nph && rc == AP_OK; ++ph) {
    ConvGeom g = transposed ? geom_convT_phase(B, H, Cin, Cout, ph >> 1, ph & 1)
                            : geom_conv(B, H, Cin, Cout, ksize, stride, pad, pad_mode);
    if (!tc) {
      SimtConvP p{};
      p.g = g; p.in = (const float*)in.p0; p.in_nchw = 0; p.in_C = Cin; p.in_coff = 0; p.wpk = lw.simt;
      p.out = out.p; p.out_C = Cout; p.out_coff = 0; p.stats = out.stats; p.stat_C = Cout; p.stat_coff = 0;
      rc = launch_conv_simt(p, st);
    } else {
      UmmaConv* c = nullptr;
      rc = umma_conv_create(&c, g, in, 0, lw.hi, lw.lo, impl == AP_PREC_FP32X3 ? 3 : 1, out.p, Cout, 0, out.stats, Cout, 0);
      if (rc == AP_OK) { convs.push_back(c); rc = umma_conv_launch(c, st); }
    }
  }
  if (rc == AP_OK) {
    ReadP rp{B, Ho, Ho, Cout, FMT_F32, out.p, nullptr, Cout, 0, 0, nullptr, 0, 0, 0, y};
    rc = launch_read(rp, st);
  }
  if (rc == AP_OK && stats) rc = launch_stats_to_double(out.stats, stats, (size_t)B * Cout * 2, st);
  cudaError_t e = cudaStreamSynchronize(st);
  for (UmmaConv* c : convs) umma_conv_destroy(c);
  for (void* p : tmp) cudaFree(p);
  if (rc == AP_OK && e != cudaSuccess) { set_error("conv2d_debug: %s", cudaGetErrorString(e)); rc = AP_ERR_CUDA; }
  return rc;
}

}  // extern "C"
