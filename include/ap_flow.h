/*
 * ap_flow.h -- C ABI of the B200-native intrinsic-flow network netF (libapnetg.so), SURVEY.md §8 row f3.
 *
 * The reference runs `FlowUnet` (Module2/intrinsic_flow_models/networks.py:577-644) once per frame in
 * `GeomCGTIFWTestModel.set_input` (Module2/models/geomcgt_ifw_test_model.py:274) through `flow_network_warp`
 * (:62-76): 2 x 68 binary key-point maps at 224x224 in, pixel flow + visibility out, then arg-max / mask / rescale /
 * resize to the two 256x256 tensors the generator consumes (`iw_flow`, `real_A_if_mask`).
 * The network's configuration lives in the checkpoint directory's train_opt.json, which does not ship: the entry
 * points are parametric in exactly the arguments `FlowRegressionModel.initialize` passes
 * (Module2/intrinsic_flow_models/flow_regression_model.py:19-38, `which_model == 'unet'`).
 *
 * Plain pointers and sizes only; fp32 device tensors; every function returns 0 or a negative AP_ERR_* code
 * (ap_netg.h) and ap_last_error() holds the message.
 */
#ifndef AP_FLOW_H
#define AP_FLOW_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct ap_flow ap_flow;

/* Replaces: networks.FlowUnet(input_nc, nf, start_scale, num_scale, norm, gpu_ids)
 * (intrinsic_flow_models/networks.py:577-627; built by flow_regression_model.py:19-27 and loaded by
 * geomcgt_ifw_test_model.py:50-60).  norm: 0 = 'batch' (eval mode, running statistics), 1 = 'instance'.
 * `size` is the input height = width (the caller feeds 224); every level must halve exactly, else AP_ERR_UNSUPPORTED
 * (the reference fails in torch.cat for such configurations). */
int ap_flow_create(ap_flow** handle, int input_nc, int nf, int start_scale, int num_scale, int norm, int max_nf, int size,
                   int device);
int ap_flow_destroy(ap_flow* handle);

/* Replaces: netF.load_state_dict(...).  `names[i]` are the reference's state_dict keys ("conv_downsample.0.weight",
 * "unet_block.submodule.down.1.weight", "unet_block.up.2.running_var", "predict_vis.1.bias", ...), `ptrs[i]` fp32 tensors
 * in torch layout, `shapes` n x 4 int64 (unused trailing dims = 1).  Every tensor the configuration needs must be
 * present (the inner levels' `predict_flow` heads are accepted and ignored: forward() of the caller drops them). */
int ap_flow_load_weights(ap_flow* handle, int n, const char* const* names, const float* const* ptrs, const int64_t* shapes,
                         int on_device, void* cuda_stream);

/* Replaces: flow_out, vis_out, _, _ = netF(input_F)  and the rest of flow_network_warp
 * (geomcgt_ifw_test_model.py:67-75).  kp_maps [B,input_nc,size,size] (ap_cond_kp_to_map makes them).
 * Any output may be NULL:  flow_out [B,2,R,R] and vis_out [B,3,R,R] with R = 2 * size / start_scale (the reference
 * up-samples by a hard-coded 2, networks.py:583,641);  iw_flow [B,2,256,256] = resize(flow_out * 20 * mask * 8/7) and
 * if_mask [B,1,256,256] = resize(mask), mask = (argmax(vis_out) < 2), bilinear align_corners=True.
 * Asynchronous on `cuda_stream`. */
int ap_flow_forward(ap_flow* handle, int B, const float* kp_maps, float* flow_out, float* vis_out, float* iw_flow,
                    float* if_mask, void* cuda_stream);

/* Replaces: flow_network_warp(netF, real_A, lm1, lm2) as a whole (geomcgt_ifw_test_model.py:62-76; call site :274) -- the
 * form the per-frame caller uses.  lm1 / lm2: source / target landmarks [B,K,2] (x,y) float32 in the 256x256 frame of the
 * photo, K = input_nc / 2; lm1_per_frame = 0: lm1 is ONE set [K,2] shared by the batch (a clip is rendered from one
 * photo).  The points are scaled by 7/8 in fp32 and drawn as binary discs of radius 4 (kp_to_map_some, :12-44, :65-66)
 * straight into the network's operand -- no key-point tensors, no concatenation, and the zero-skipping first conv takes
 * its boxes from the coordinates -- then the network and the arg-max / mask / rescale / resize tail run as in
 * ap_flow_forward.  `real_A` is only resized and dropped by the reference and has no counterpart here.
 * iw_flow [B,2,256,256], if_mask [B,1,256,256].  Asynchronous on `cuda_stream`. */
int ap_flow_warp_landmarks(ap_flow* handle, int B, const float* lm1, int lm1_per_frame, const float* lm2, float* iw_flow,
                           float* if_mask, void* cuda_stream);

/* Kernels launched by the most recent forward (the "did the CUDA path run" counter of the tests). */
int ap_flow_last_launch_count(ap_flow* handle, int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* AP_FLOW_H */
