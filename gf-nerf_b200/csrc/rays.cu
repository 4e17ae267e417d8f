// Ray generation: Cameras.generate_rays for PERSPECTIVE cameras without distortion.
//
// Replaces the ~40 torch ops of nerfstudio/cameras/cameras.py:583-727 (stack / masked_select / broadcast-multiply /
// sum / normalize_with_norm / sqrt), as GF-NeRF's datamanager calls them once per training batch
// (nerfstudio/data/datamanagers/base_datamanager.py:923-948) and the render loop once per frame, with the
// `lookat_directions` the reference adds (:704, 723).  One thread per ray; every output row is written once.
// Arithmetic = torch's fp32 op sequence, bit for bit against the reference's own output for origins, directions,
// lookat directions and the direction norm (tests/golden/ref_rays.npz): separate mul / add / div, except
// torch.linalg.vector_norm, which accumulates the squares with FMAs.
#include "common.cuh"

namespace gf {

__global__ void __launch_bounds__(256)
generate_rays_kernel(int64_t n, const int64_t* __restrict__ cam_idx, const float* __restrict__ coords_yx,
                     const float* __restrict__ c2w, const float* __restrict__ fx, const float* __restrict__ fy,
                     const float* __restrict__ cx, const float* __restrict__ cy, int64_t n_cams,
                     float* __restrict__ origins, float* __restrict__ directions, float* __restrict__ lookat,
                     float* __restrict__ pixel_area, float* __restrict__ dir_norm) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    int64_t c = cam_idx[i];
    c = c < 0 ? 0 : (c >= n_cams ? n_cams - 1 : c);  // torch would raise on a bad index; checked on the host side
    const float2 yx = __ldg(reinterpret_cast<const float2*>(coords_yx) + i);
    const float fxc = __ldg(fx + c), fyc = __ldg(fy + c), cxc = __ldg(cx + c), cyc = __ldg(cy + c);
    const float4 r0 = __ldg(reinterpret_cast<const float4*>(c2w) + 3 * c);
    const float4 r1 = __ldg(reinterpret_cast<const float4*>(c2w) + 3 * c + 1);
    const float4 r2 = __ldg(reinterpret_cast<const float4*>(c2w) + 3 * c + 2);
    const float xm = __fsub_rn(yx.y, cxc), ym = __fsub_rn(yx.x, cyc);
    // :606-608 the pixel and its +1 neighbours in x and in y, in image-plane coordinates
    const float u[3] = {__fdiv_rn(xm, fxc), __fdiv_rn(__fadd_rn(xm, 1.f), fxc), __fdiv_rn(xm, fxc)};
    const float v[3] = {-__fdiv_rn(ym, fyc), -__fdiv_rn(ym, fyc), -__fdiv_rn(__fadd_rn(ym, 1.f), fyc)};
    float d[3][3], nrm0 = 0.f;
#pragma unroll
    for (int k = 0; k < 3; k++) {
      // :697-699 sum(dir * rotation, -1): (u R_r0 + v R_r1) + (-1) R_r2, then normalize_with_norm
      const float w0 = __fadd_rn(__fadd_rn(__fmul_rn(u[k], r0.x), __fmul_rn(v[k], r0.y)), -r0.z);
      const float w1 = __fadd_rn(__fadd_rn(__fmul_rn(u[k], r1.x), __fmul_rn(v[k], r1.y)), -r1.z);
      const float w2 = __fadd_rn(__fadd_rn(__fmul_rn(u[k], r2.x), __fmul_rn(v[k], r2.y)), -r2.z);
      float nrm = __fsqrt_rn(__fmaf_rn(w2, w2, __fmaf_rn(w1, w1, __fmul_rn(w0, w0))));
      nrm = fmaxf(nrm, 8.8817842e-16f);  // _EPS, camera_utils.py:28
      d[k][0] = __fdiv_rn(w0, nrm);
      d[k][1] = __fdiv_rn(w1, nrm);
      d[k][2] = __fdiv_rn(w2, nrm);
      if (k == 0) nrm0 = nrm;
    }
    origins[3 * i] = r0.w;
    origins[3 * i + 1] = r1.w;
    origins[3 * i + 2] = r2.w;
    directions[3 * i] = d[0][0];
    directions[3 * i + 1] = d[0][1];
    directions[3 * i + 2] = d[0][2];
    if (lookat) {
      lookat[3 * i] = r0.z;
      lookat[3 * i + 1] = r1.z;
      lookat[3 * i + 2] = r2.z;
    }
    if (pixel_area) {  // :711-716
      float s[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const float e0 = __fsub_rn(d[0][0], d[k + 1][0]), e1 = __fsub_rn(d[0][1], d[k + 1][1]),
                    e2 = __fsub_rn(d[0][2], d[k + 1][2]);
        s[k] = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(e0, e0), __fmul_rn(e1, e1)), __fmul_rn(e2, e2)));
      }
      pixel_area[i] = __fmul_rn(s[0], s[1]);
    }
    if (dir_norm) dir_norm[i] = nrm0;
  }
}

// Error-map feedback of the focal (block) stage: error = sum_c |gt - pred| per ray (gfnerf/gf_pipeline.py:180-184)
// written to error_map[image, y, x] (nerfstudio/data/utils/dataloaders.py:140-142) -- the abs / sum / index_put chain
// as one pass.  One thread per ray.  Duplicate (image, y, x) triples in one batch store the same pixel more than
// once; like torch's index_put without accumulate, which of the duplicates lands last is unspecified.
__global__ void __launch_bounds__(256)
error_map_update_kernel(int64_t n, const int64_t* __restrict__ idx, const float* __restrict__ pred,
                        const float* __restrict__ gt, int64_t n_images, int64_t height, int64_t width,
                        float* __restrict__ error_map, float* __restrict__ error_out, int* __restrict__ bad) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const float e = __fadd_rn(__fadd_rn(fabsf(__fsub_rn(gt[3 * i], pred[3 * i])),
                                        fabsf(__fsub_rn(gt[3 * i + 1], pred[3 * i + 1]))),
                              fabsf(__fsub_rn(gt[3 * i + 2], pred[3 * i + 2])));
    if (error_out) error_out[i] = e;
    int64_t c = idx[3 * i], y = idx[3 * i + 1], x = idx[3 * i + 2];
    // torch indexing wraps negative indices once and raises beyond that; out-of-range rows are skipped and flagged
    if (c < 0) c += n_images;
    if (y < 0) y += height;
    if (x < 0) x += width;
    if (c < 0 || c >= n_images || y < 0 || y >= height || x < 0 || x >= width) {
      if (bad) *bad = 1;
      continue;
    }
    error_map[(c * height + y) * width + x] = e;
  }
}

}  // namespace gf

using namespace gf;

extern "C" int gf_error_map_update(int64_t n_rays, const int64_t* indices, const float* pred_rgb,
                                   const float* gt_rgb, int64_t n_images, int64_t height, int64_t width,
                                   float* error_map, float* error_out, int32_t* bad_index_flag, void* stream) {
  GF_REQUIRE(n_rays >= 0 && n_images > 0 && height > 0 && width > 0,
             "gf_error_map_update: bad sizes n_rays=%lld map=[%lld,%lld,%lld]", (long long)n_rays,
             (long long)n_images, (long long)height, (long long)width);
  if (n_rays == 0) return GF_OK;
  GF_REQUIRE(indices && pred_rgb && gt_rgb && error_map, "gf_error_map_update: null pointer");
  error_map_update_kernel<<<stride_grid(n_rays, 256, 8, 4), 256, 0, (cudaStream_t)stream>>>(
      n_rays, indices, pred_rgb, gt_rgb, n_images, height, width, error_map, error_out, bad_index_flag);
  return check_launch("error_map_update_kernel");
}

extern "C" int gf_generate_rays(int64_t n_rays, const int64_t* cam_idx, const float* coords_yx, const float* c2w,
                                const float* fx, const float* fy, const float* cx, const float* cy, int64_t n_cams,
                                float* origins, float* directions, float* lookat, float* pixel_area, float* dir_norm,
                                void* stream) {
  GF_REQUIRE(n_rays >= 0 && n_cams > 0, "gf_generate_rays: bad sizes n_rays=%lld n_cams=%lld", (long long)n_rays,
             (long long)n_cams);
  if (n_rays == 0) return GF_OK;
  GF_REQUIRE(cam_idx && coords_yx && c2w && fx && fy && cx && cy && origins && directions,
             "gf_generate_rays: null pointer");
  generate_rays_kernel<<<stride_grid(n_rays, 256, 8, 4), 256, 0, (cudaStream_t)stream>>>(
      n_rays, cam_idx, coords_yx, c2w, fx, fy, cx, cy, n_cams, origins, directions, lookat, pixel_area, dir_norm);
  return check_launch("generate_rays_kernel");
}
