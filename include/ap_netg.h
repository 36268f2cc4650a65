/*
 * ap_netg.h -- C ABI of the B200-native Module2 generator (libapnetg.so).
 *
 * This is the drop-in boundary for ONE path of AnimatePortrait: the per-frame generator
 * `ResnetConditionTriGenerator32_full_ifw` (netG string 'resnet_9blocks_rcatland32_full_ifw').
 * The reference has no FFI of its own (it is pure PyTorch); each entry point below names the
 * reference interface it stands in for.  Plain pointers and sizes only, no torch types.
 * The Python host side (animateportrait_b200/netg.py) binds these with ctypes; INTEGRATION.md
 * shows the stub a maintainer of the reference would add.
 *
 * All tensors are fp32.  Device pointers unless the function name ends in _host.
 * Every function returns 0 on success or a negative AP_ERR_* code; ap_last_error() returns a
 * thread-local human-readable message for the last failure.
 */
#ifndef AP_NETG_H
#define AP_NETG_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ap_netg ap_netg;

enum {
  AP_OK = 0,
  AP_ERR_INVALID = -1,     /* bad argument (the reference raises NotImplementedError / assert) */
  AP_ERR_CUDA = -2,        /* a CUDA runtime / driver call failed */
  AP_ERR_UNSUPPORTED = -3, /* configuration outside the supported generator */
  AP_ERR_STATE = -4        /* e.g. forward before load_weights */
};

/* Arithmetic of the conv layers. */
enum {
  AP_PREC_FP32X3 = 0,   /* tcgen05 bf16 hi/lo split, 3 products, fp32 accumulate: fp32-accurate (<=1e-3) */
  AP_PREC_BF16 = 1,     /* tcgen05 bf16 operands, 1 product, fp32 accumulate, bf16 activations */
  AP_PREC_FP32_SIMT = 2 /* CUDA-core fp32 FMA path (validation of the tensor-core path) */
};

/* Replaces: networks.define_G(3, output_nc, 64, 'resnet_9blocks_rcatland32_full_ifw', 'instance',
 * False, ..., div=3, disp=3)  (Module2/models/networks.py:123-201, branch :175-176; call site
 * Module2/models/geomcgt_ifw_test_model.py:207-209) and the constructor networks.py:1196-1296.
 * output_nc in {1,3}. `device` is the CUDA ordinal (the reference's gpu_ids[0]). */
int ap_netg_create(ap_netg** handle, int output_nc, int precision, int device);

/* Replaces: nn.Module destruction. Frees weights and workspaces. */
int ap_netg_destroy(ap_netg* handle);

/* Replaces: net.load_state_dict(state_dict)  (Module2/models/base_model.py:202).
 * `names[i]` are the reference checkpoint keys (e.g. "model2.0.conv_block.1.weight"),
 * `ptrs[i]` fp32 tensors in torch layout (Conv2d [Cout,Cin,kh,kw]; ConvTranspose2d [Cin,Cout,kh,kw]),
 * `shapes` n x 4 int64 (unused trailing dims = 1), on host (on_device=0) or on the handle's device (1).
 * All 74 tensors must be present (strict, like load_state_dict); may be called again to swap weights.
 * Weights are re-packed on the device for the kernels (tap-major, K-major bf16 hi/lo for tcgen05). */
int ap_netg_load_weights(ap_netg* handle, int n, const char* const* names, const float* const* ptrs,
                         const int64_t* shapes, int on_device, void* cuda_stream);

/* Bytes of device workspace the library allocates (and keeps) for batch size B. */
int ap_netg_workspace_bytes(ap_netg* handle, int B, size_t* bytes);

/* Replaces: netG.forward(input, land1, land2, motion, flow, ifmask)  (networks.py:1315-1340; call site
 * geomcgt_ifw_test_model.py:295).  Shapes (contiguous): input [B,3,256,256], land1/land2 [B,1,256,256],
 * motion [B,256,256,2], flow [B,2,256,256], ifmask [B,1,256,256]; out [B,output_nc,256,256].
 * Asynchronous on `cuda_stream` (a cudaStream_t; NULL = legacy default stream). Inputs are not modified. */
int ap_netg_forward(ap_netg* handle, int B, const float* input, const float* land1, const float* land2,
                    const float* motion, const float* flow, const float* ifmask, float* out,
                    void* cuda_stream);

/* Clip form of the same call: ONE photo shared by all B frames of the batch (the reference renders a clip frame by
 * frame from the same photo, Module2/test.py:58-65, and recomputes the photo's features every time).
 * `input` is [1,3,256,256] and `land1` (the SOURCE landmark map, which belongs to the photo) [1,1,256,256]; everything
 * else as ap_netg_forward.  The part of the network that depends on them alone (model_tri00/10/20 stems, model_tri11,
 * model_tri21, model_tri22, model_landmark_trans(land1); networks.py:1318-1331) runs once per call; the three double
 * warps and the ResnetBlock2 inputs read it for every frame.  Same results as ap_netg_forward on B copies of both. */
int ap_netg_forward_shared_photo(ap_netg* handle, int B, const float* input, const float* land1, const float* land2,
                                 const float* motion, const float* flow, const float* ifmask, float* out,
                                 void* cuda_stream);

/* Same call with HOST buffers (pinned or pageable): copies the six inputs to the device, runs the
 * forward, copies the frames back and synchronises the stream.  This is the end-to-end form the
 * reference's per-frame loop (Module2/test.py:58-65) sees: CPU tensors in, CPU frames out. */
int ap_netg_forward_host(ap_netg* handle, int B, const float* input, const float* land1,
                         const float* land2, const float* motion, const float* flow, const float* ifmask,
                         float* out, void* cuda_stream);

/* Pipelined form of ap_netg_forward_host for a clip rendered batch after batch (the reference's loop over frames,
 * Module2/test.py:58-65): returns as soon as the copies and the forward are enqueued.  Two staging slots alternate, so the
 * inputs of call k+1 upload while call k computes and the frames of call k download while call k+1 computes; a third call
 * in flight waits for the first to drain.  The host buffers (pinned for overlap) must stay valid, and `out` unread, until
 * ap_netg_host_sync returns.  Same frames as ap_netg_forward_host. */
int ap_netg_forward_host_async(ap_netg* handle, int B, const float* input, const float* land1, const float* land2,
                               const float* motion, const float* flow, const float* ifmask, float* out,
                               void* cuda_stream);
int ap_netg_host_sync(ap_netg* handle);

/* Output stage after netG, one fused elementwise kernel (device pointers, asynchronous on `cuda_stream`).
 * Replaces: the foreground/background blend of GeomCGTIFWTestModel.forward
 *   mask1 = F.grid_sample(mask, warp_motion, align_corners=True); fake_B = ((fake_B/2+0.5)*mask1 +
 *   (fakeB_static/2+0.5)*(1-mask1))*2-1          (Module2/models/geomcgt_ifw_test_model.py:297-300)
 * and util.tensor2im, ((x+1)/2*255).astype(uint8) as HWC RGB with grayscale tiled to 3 channels
 * (Module2/util/util.py:9-29; call site Module2/util/visualizer.py:16-52 via Module2/test.py:63-65).
 *   fake_B [B,onc,256,256]; mask [B,1,256,256], motion [B,256,256,2], static_B [B,onc,256,256] -- all three or none
 *   (none: no blend, conversion only); blended [B,onc,256,256] fp32 and/or image_u8 [B,256,256,3], either may be NULL. */
int ap_netg_compose(int device, int B, int output_nc, const float* fake_B, const float* mask, const float* motion,
                    const float* static_B, float* blended, uint8_t* image_u8, void* cuda_stream);

/* Execution options of a handle (none changes results beyond what is stated):
 *   "graphs"             1 (default; AP_NETG_GRAPH=0 turns it off): the launch sequence of a batch shape is captured
 *                        once and replayed as a CUDA graph -- the reference's own call is batch size 1
 *                        (Module2/test.py:42), where launch overhead dominates; 0: plain stream launches.  Bit-identical.
 *   "overlap"            1 (default; AP_NETG_OVERLAP=0): independent branches of the graph run concurrently.  Bit-identical.
 *   "keep_intermediates" 0 (default): buffers of dead intermediates are reused; 1: every intermediate keeps its buffer so
 *                        that ap_netg_debug_read can return any tap after the forward. */
int ap_netg_set_option(ap_netg* handle, const char* name, int value);

/* Multi-GPU gather hook: lets kernels running on `device` store into memory of `peer_device` (same
 * process, or another process's buffer mapped with CUDA IPC).  `out` of ap_netg_forward may then be such memory: the
 * output kernel of every rank writes its frames straight into ONE buffer on the GPU that owns the clip -- the gather
 * of the frames costs no copy and no collective.  Replaces nothing in the reference (it is single-GPU,
 * main_end2end_module2.py:108 passes --gpu_ids 0); idempotent. */
int ap_device_enable_peer_access(int device, int peer_device);

/* The gather buffer itself.  ap_peer_alloc (on the rank that owns the clip): `bytes` of device memory on `device` plus
 * the 64-byte CUDA IPC handle other processes of the box open it with.  ap_peer_open (every other rank): maps that
 * memory into the address space of `device`, the GPU whose kernels will store into it, enabling NVLink peer access
 * to the owner on the way.  ap_peer_close / ap_peer_free undo them.  Plain pointers: the Python side wraps them as
 * tensors and passes slices of them as `out` of ap_netg_forward. */
int ap_peer_alloc(int device, size_t bytes, void** ptr, unsigned char* handle64);
int ap_peer_open(int device, const unsigned char* handle64, void** ptr);
int ap_peer_close(void* ptr);
int ap_peer_free(void* ptr);

/* Number of kernels of this library launched (or replayed from the graph) by the most recent forward on this handle. */
int ap_netg_last_launch_count(ap_netg* handle, int64_t* count);

/* Per-kernel-class device timing of the forward (CUDA events on the launch stream, one per launch).
 * While enabled every forward records events; ap_netg_get_profile synchronises and returns, for the
 * most recent forward, per class: total milliseconds, launches and algorithmic FLOPs (2*MACs).
 * Classes: 0 stem7x7 (CUDA cores), 1 landmark convs (CUDA cores), 2 trunk conv 3x3 s1 @64x64,
 * 3 strided / transposed convs, 4 InstanceNorm apply, 5 double warp, 6 output conv 7x7 + tanh.
 * Returns the number of classes written (<= max_classes) through *n_classes. */
int ap_netg_set_profiling(ap_netg* handle, int enable);
int ap_netg_get_profile(ap_netg* handle, int max_classes, double* ms, int64_t* launches, double* flops,
                        int* n_classes);

/* Debug/validation: copy a named intermediate of the most recent forward to `dst` as NCHW fp32.
 * Names follow oracle/netg_oracle.py taps: tri00 warp0 tri01 tri02 tri10 tri11 warp1 tri12 tri20 tri21
 * tri22 warp2 merge land1 land2 block0..block8 up0 up1.  `shape4` receives [B,C,H,W]. */
int ap_netg_debug_read(ap_netg* handle, const char* tap, float* dst, size_t dst_capacity_elems,
                       int64_t* shape4, void* cuda_stream);

/* Debug/validation of one convolution layer through the same kernels the forward uses.
 *   impl: AP_PREC_* (which kernel family); transposed: 0 Conv2d, 1 ConvTranspose2d(k3,s2,p1,op1)
 *   x [B,Cin,H,W] NCHW, w in torch layout, y NCHW raw conv output (no bias), stats [B,Cout,2] doubles
 *   (sum, sum of squares over HxW of y) or NULL.  pad_mode: 0 zero, 1 reflect.  Device pointers. */
int ap_conv2d_debug(int impl, int device, int B, int H, int W, int Cin, int Cout, int ksize, int stride,
                    int pad, int pad_mode, int transposed, const float* x, const float* w, float* y,
                    double* stats, void* cuda_stream);

/* Thread-local message for the last error returned on this thread. */
const char* ap_last_error(void);

/* "apnetg <version> sm_100a" */
const char* ap_version(void);

#ifdef __cplusplus
}
#endif
#endif /* AP_NETG_H */
