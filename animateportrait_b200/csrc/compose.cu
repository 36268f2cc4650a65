// Output stage after netG ("next" rows f1 + f4 of the scope table): the foreground/background blend of
// GeomCGTIFWTestModel.forward (Module2/models/geomcgt_ifw_test_model.py:297-300)
//     mask1  = F.grid_sample(mask, warp_motion, align_corners=True)          bilinear, zero padding
//     fake_B = ((fake_B/2 + 0.5) * mask1 + (fakeB_static/2 + 0.5) * (1 - mask1)) * 2 - 1
// and the image conversion of util.tensor2im (Module2/util/util.py:9-29)
//     img = ((x + 1) / 2.0 * 255.0).astype(uint8), HWC, grayscale tiled to 3 channels
// as ONE elementwise kernel: a handful of bytes per pixel instead of three passes over fp32 tensors, a D2H copy of
// fp32 frames and a numpy conversion on the host.  One thread per pixel; arithmetic in the reference's op order
// (__f*_rn keeps nvcc from contracting it into FMAs).
#include "common.cuh"

namespace ap {

void launches_add(int n);

struct ComposeP {
  const float* fake;     // [B,onc,256,256]
  const float* mask;     // [B,1,256,256] or null (no blend)
  const float* motion;   // [B,256,256,2]
  const float* stat;     // [B,onc,256,256] static drawing
  float* blended;        // [B,onc,256,256] or null
  uint8_t* image;        // [B,256,256,3] or null
  int B, onc;
};

__device__ __forceinline__ float mask_tap(const float* m, int y, int x) {
  return (x >= 0 && x < 256 && y >= 0 && y < 256) ? m[y * 256 + x] : 0.f;
}

__global__ void __launch_bounds__(256) compose_kernel(const ComposeP p) {
  const int n = blockIdx.y;
  const int pix = blockIdx.x * 256 + threadIdx.x;  // 65536 pixels per image
  float m1 = 1.f;
  if (p.mask) {
    const float2 g = *reinterpret_cast<const float2*>(p.motion + ((size_t)n * 65536 + pix) * 2);
    // grid_sampler_unnormalize(align_corners=True): ((g + 1) / 2) * (size - 1)
    const float ix = __fmul_rn(__fdiv_rn(__fadd_rn(g.x, 1.f), 2.f), 255.f);
    const float iy = __fmul_rn(__fdiv_rn(__fadd_rn(g.y, 1.f), 2.f), 255.f);
    const float fx = floorf(ix), fy = floorf(iy);
    const int x0 = (int)fx, y0 = (int)fy;
    const float wx = __fsub_rn(ix, fx), wy = __fsub_rn(iy, fy);
    const float ex = __fsub_rn(1.f, wx), ey = __fsub_rn(1.f, wy);
    const float* m = p.mask + (size_t)n * 65536;
    // ATen CPU grid sampler: nw * (ex*ey) + ne * (wx*ey) + sw * (ex*wy) + se * (wx*wy), accumulated in this order
    float acc = __fmul_rn(mask_tap(m, y0, x0), __fmul_rn(ex, ey));
    acc = __fadd_rn(acc, __fmul_rn(mask_tap(m, y0, x0 + 1), __fmul_rn(wx, ey)));
    acc = __fadd_rn(acc, __fmul_rn(mask_tap(m, y0 + 1, x0), __fmul_rn(ex, wy)));
    acc = __fadd_rn(acc, __fmul_rn(mask_tap(m, y0 + 1, x0 + 1), __fmul_rn(wx, wy)));
    m1 = acc;
  }
  uint8_t rgb[3];
  for (int c = 0; c < p.onc; ++c) {
    const size_t off = ((size_t)(n * p.onc + c)) * 65536 + pix;
    float v = p.fake[off];
    if (p.mask) {
      const float a = __fadd_rn(__fdiv_rn(v, 2.f), 0.5f);
      const float b = __fadd_rn(__fdiv_rn(p.stat[off], 2.f), 0.5f);
      const float mix = __fadd_rn(__fmul_rn(a, m1), __fmul_rn(b, __fsub_rn(1.f, m1)));
      v = __fsub_rn(__fmul_rn(mix, 2.f), 1.f);
    }
    if (p.blended) p.blended[off] = v;
    const float u = __fmul_rn(__fdiv_rn(__fadd_rn(v, 1.f), 2.f), 255.f);
    rgb[c] = (uint8_t)(int)u;  // numpy astype(uint8) of an in-range float: truncation
  }
  if (p.image) {
    if (p.onc == 1) { rgb[1] = rgb[0]; rgb[2] = rgb[0]; }
    uint8_t* dst = p.image + ((size_t)n * 65536 + pix) * 3;
    dst[0] = rgb[0]; dst[1] = rgb[1]; dst[2] = rgb[2];
  }
}

int launch_compose(const ComposeP& p, cudaStream_t st) {
  dim3 grid(256, p.B);
  compose_kernel<<<grid, 256, 0, st>>>(p);
  AP_CUDA(cudaGetLastError());
  launches_add(1);
  return AP_OK;
}

}  // namespace ap

extern "C" int ap_netg_compose(int device, int B, int output_nc, const float* fake_B, const float* mask, const float* motion,
                               const float* static_B, float* blended, uint8_t* image_u8, void* cuda_stream) {
  using namespace ap;
  AP_REQUIRE(B >= 1 && (output_nc == 1 || output_nc == 3) && fake_B, AP_ERR_INVALID, "compose: bad argument");
  AP_REQUIRE(blended || image_u8, AP_ERR_INVALID, "compose: no output requested");
  AP_REQUIRE((mask == nullptr) == (motion == nullptr) && (mask == nullptr) == (static_B == nullptr), AP_ERR_INVALID,
             "compose: mask, motion and static frame come together (all or none)");
  AP_CUDA(cudaSetDevice(device));
  ComposeP p{fake_B, mask, motion, static_B, blended, image_u8, B, output_nc};
  return launch_compose(p, (cudaStream_t)cuda_stream);
}
