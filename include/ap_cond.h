/*
 * ap_cond.h -- C ABI of the per-frame CONDITIONING producers that feed the generator (libapnetg.so).
 *
 * SURVEY.md §8 row f2 (+ the matting line of row f1): what the reference computes on the CPU, one frame at a
 * time, before netG runs -- the landmark disc maps, the Delaunay piecewise-linear motion field, the key-point
 * maps of the flow network and the photo matting.  Here each is one CUDA launch over a whole batch of frames, so
 * the per-frame payload that has to reach the GPU shrinks from 6 planes of 256x256 floats to 68x2 floats.
 * The reference has no FFI (pure Python: numpy + cv2 + scipy); every entry point names the function it replaces.
 * Plain pointers and sizes only.  DEVICE pointers, asynchronous on `cuda_stream` (a cudaStream_t).
 * Return codes and ap_last_error() as in ap_netg.h.
 */
#ifndef AP_COND_H
#define AP_COND_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AP_COND_LANDMARKS 68      /* points per face (dlib 68-point layout) */
#define AP_COND_MAX_TRIANGLES 512 /* per-frame capacity of the Delaunay triangle list (139 in general position) */

/* Replaces: draw2(size, size, lands, radius, thickness, op=0)   (Module2/data/umlvdfw_test_dataset.py:34-41;
 * call sites :147-148 with radius 3 at crop_size 256, 5 at 512) for T frames at once.
 *   lands [T,n_points,2] (x,y) float32;  out [T,1,size,size] float32 in {-1,+1}.
 * Centres are rounded half-to-even (np.round); discs are OpenCV's filled midpoint circle, clipped at the border.
 * size must be a multiple of 4, radius in [0,15]. */
int ap_cond_draw_landmarks(int device, int T, int n_points, int size, int radius, const float* lands, float* out,
                           void* cuda_stream);

/* Replaces: cal_motion256(lm2d0, lm2d)   (Module2/data/umlvdfw_test_dataset.py:67-81; call site :161), i.e.
 * scipy.interpolate.griddata(method='linear') of the source landmark positions over the Delaunay triangulation of
 * the target landmarks + the four image corners, evaluated on the 256x256 pixel lattice, as float32 / 127.5 - 1.
 *   lm_src  [T,68,2] (src_per_frame=1) or [1,68,2] (src_per_frame=0: one photo, many target frames), (x,y) float32
 *   lm_dst  [T,68,2]
 *   motion  [T,256,256,2]  (channel 0 = x, 1 = y: the `warp_motion` grid netG and F.grid_sample take)
 *   workspace: ap_cond_motion256_workspace_bytes(T) bytes of device memory (triangle lists)
 *   tri_count [T] int32 or NULL: number of Delaunay triangles found per frame (> AP_COND_MAX_TRIANGLES = overflow,
 *   the list was truncated).
 * Two launches: exhaustive empty-circumcircle test of all C(72,3) site triples in fp64, then a tiled
 * point-in-triangle rasteriser with barycentric interpolation in fp64.  Pixels no triangle covers get NaN (as griddata). */
int ap_cond_motion256(int device, int T, const float* lm_src, int src_per_frame, const float* lm_dst, float* motion,
                      void* workspace, size_t workspace_bytes, int32_t* tri_count, void* cuda_stream);
int ap_cond_motion256_workspace_bytes(int T, size_t* bytes);

/* Replaces: kp_to_map_some((size,size), kps, mode='binary', radius=4)   (Module2/models/geomcgt_ifw_test_model.py:12-44;
 * call sites :62-63 with kps = landmarks*7/8, size 224).
 *   kps [T,K,2] (x,y) float32;  out [T,K,size,size] float32 in {0,1};  a point with x == -1 or y == -1 gives an
 *   empty map.  The distance test runs in fp64 as numpy does.  size must be a multiple of 4. */
int ap_cond_kp_to_map(int device, int T, int K, int size, float radius, const float* kps, float* out,
                      void* cuda_stream);

/* Replaces: mask = (matte > 0.5).float(); real_A = ((real_A/2+0.5)*mask + 1-mask)*2-1
 * (Module2/models/geomcgt_ifw_test_model.py:280,292) -- frame-invariant, run once per photo.
 *   real_A [B,C,HW] , matte [B,1,HW] -> out [B,C,HW], mask [B,1,HW] (either may be NULL). */
int ap_cond_matte_photo(int device, int B, int C, int HW, const float* real_A, const float* matte, float* out,
                        float* mask, void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* AP_COND_H */
