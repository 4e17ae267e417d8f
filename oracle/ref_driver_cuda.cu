/*
 * ref_driver_cuda.cu -- the REFERENCE's own device functions compiled by nvcc for sm_100a and launched on the GPU.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  Same extraction as ref_driver.cpp (oracle/ref_extract.py pulls the
 * __global__ / __device__ bodies of Hash3DAnchored_cuda.cu / PersSampler_cuda.cu out of /root/reference at build
 * time), but here the text is compiled as real CUDA -- nvcc's own -fmad=true contraction, real half2 atomics -- with
 * only the Eigen subset (ref_shim/eigen_subset.h) standing in for the un-vendored library.  `make -C oracle ref_cuda`
 * -> oracle/_ref/libgf_ref_cuda.so; it travels to the GPU box with the snapshot.  Used by the GF_REF_CUDA=1 tests of
 * tests/test_ref_kernels.py: our kernels beside the reference's on the same B200, the bit-level pin the host build
 * (ref_driver.cpp) cannot give.
 *
 * Every pointer is a DEVICE pointer.  The entry points restate only the launch sequences of the reference's torch
 * host functions (grid / block shapes as Utils/Common.h:36-38: 512 threads, ceil-div grid); temporaries come from
 * cudaMalloc.  Return 0 or the cudaError_t.
 */
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "ref_shim/eigen_subset.h"

#include "_ref/ref_kernels.inc"

static_assert(sizeof(TransInfo) == 576 && sizeof(TreeNode) == 128 && sizeof(EdgePool) == 64, "reference struct layout");

#define THREADS 512u
static inline unsigned div_up(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }
#define CK(x)                        \
  do {                               \
    cudaError_t e_ = (x);            \
    if (e_ != cudaSuccess) return (int)e_; \
  } while (0)

extern "C" {

int refcu_hash_forward(int n_points, int n_volumes, void* feat_pool_f16, int* prim_pool, int* feat_local_idx,
                       int* feat_local_size, float* bias_pool, float* points, int64_t* volume_idx, void* out_f16) {
  dim3 block(THREADS, 1, 1), grid(div_up(n_points, THREADS), N_LEVELS, 1);
  Hash3DAnchoredForwardKernel<__half><<<grid, block>>>(n_points, n_volumes, (__half*)feat_pool_f16, prim_pool,
                                                       feat_local_idx, feat_local_size, (Wec3f*)bias_pool,
                                                       (Wec3f*)points, volume_idx, (__half*)out_f16);
  CK(cudaGetLastError());
  return (int)cudaDeviceSynchronize();
}

int refcu_hash_backward(int n_points, int n_volumes, int* prim_pool, int* feat_local_idx, int* feat_local_size,
                        float* bias_pool, float* points, int64_t* volume_idx, void* grad_in_f16, void* grad_out_f16) {
  dim3 block(THREADS, 1, 1), grid(div_up(n_points, THREADS), N_LEVELS, 1);
  Hash3DAnchoredBackwardKernel<__half><<<grid, block>>>(n_points, n_volumes, prim_pool, feat_local_idx,
                                                        feat_local_size, (Wec3f*)bias_pool, (Wec3f*)points,
                                                        volume_idx, (__half*)grad_in_f16, (__half*)grad_out_f16);
  CK(cudaGetLastError());
  return (int)cudaDeviceSynchronize();
}

/* PersSampler::GetSamples, PersSampler_cuda.cu:321-477.  rays_d normalised, noise already x fineness; dense outputs
 * [R,1024,..] zero-filled by the caller; oct_idx_start_end [R,2] (zeroed by the caller) is returned for inspection. */
int refcu_get_samples(int64_t n_rays, float* rays_o, float* rays_d, float* noise, void* tree_nodes, void* transes,
                      uint8_t* search_order, float global_near, float sample_l, int scale_by_dis,
                      int64_t max_oct_per_ray, float* world_pts, float* warp_pts, float* dirs, float* dists, float* ts,
                      int64_t* anchors, int64_t* pts_idx_start_end, float* first_oct_dis, int64_t* oct_idx_start_end) {
  float* bounds = nullptr;
  int64_t *counter = nullptr, *oct_idx = nullptr, *sampled_oct = nullptr;
  float* oct_nf = nullptr;
  CK(cudaMalloc(&bounds, sizeof(float) * 2 * n_rays));
  CK(cudaMalloc(&counter, sizeof(int64_t)));
  CK(cudaMemset(counter, 0, sizeof(int64_t)));
  {
    float* hb = (float*)malloc(sizeof(float) * 2 * n_rays);
    for (int64_t i = 0; i < n_rays; i++) {
      hb[2 * i] = global_near;
      hb[2 * i + 1] = 1e8f;
    }
    CK(cudaMemcpy(bounds, hb, sizeof(float) * 2 * n_rays, cudaMemcpyHostToDevice));
    free(hb);
  }
  dim3 block(THREADS, 1, 1), grid(div_up(n_rays, THREADS), 1, 1);
  FindRayOctreeIntersectionKernel<false><<<grid, block>>>(n_rays, max_oct_per_ray, search_order, (Wec3f*)rays_o,
                                                          (Wec3f*)rays_d, (Wec2f*)bounds, counter,
                                                          (Wec2i64*)oct_idx_start_end, (TreeNode*)tree_nodes, nullptr,
                                                          nullptr, nullptr);
  CK(cudaGetLastError());
  int64_t n_all = 0;
  CK(cudaMemcpy(&n_all, counter, sizeof(int64_t), cudaMemcpyDeviceToHost));
  CK(cudaMalloc(&oct_idx, sizeof(int64_t) * (n_all + 1)));
  CK(cudaMalloc(&oct_nf, sizeof(float) * 2 * (n_all + 1)));
  FindRayOctreeIntersectionKernel<true><<<grid, block>>>(n_rays, max_oct_per_ray, search_order, (Wec3f*)rays_o,
                                                         (Wec3f*)rays_d, (Wec2f*)bounds, counter,
                                                         (Wec2i64*)oct_idx_start_end, (TreeNode*)tree_nodes, oct_idx,
                                                         (Wec2f*)oct_nf, nullptr);
  CK(cudaGetLastError());
  CK(cudaMemset(pts_idx_start_end, 0, sizeof(int64_t) * 2 * n_rays));
  RayMarchKernel<false><<<grid, block>>>(n_rays, sample_l, scale_by_dis != 0, (Wec3f*)rays_o, (Wec3f*)rays_d, noise,
                                         (Wec2i64*)oct_idx_start_end, oct_idx, (Wec2f*)oct_nf, (TreeNode*)tree_nodes,
                                         (TransInfo*)transes, (Wec2i64*)pts_idx_start_end, nullptr, nullptr, nullptr,
                                         nullptr, nullptr, nullptr, nullptr, nullptr);
  CK(cudaGetLastError());
  {  /* :420 pts_idx_start_end[:,0] = cumsum(pts_idx_start_end[:,0]) */
    int64_t* h = (int64_t*)malloc(sizeof(int64_t) * 2 * n_rays);
    CK(cudaMemcpy(h, pts_idx_start_end, sizeof(int64_t) * 2 * n_rays, cudaMemcpyDeviceToHost));
    int64_t run = 0;
    for (int64_t i = 0; i < n_rays; i++) {
      run += h[2 * i];
      h[2 * i] = run;
    }
    CK(cudaMemcpy(pts_idx_start_end, h, sizeof(int64_t) * 2 * n_rays, cudaMemcpyHostToDevice));
    free(h);
  }
  CK(cudaMalloc(&sampled_oct, sizeof(int64_t) * n_rays * MAX_SAMPLE_PER_RAY));
  CK(cudaMemset(sampled_oct, 0xff, sizeof(int64_t) * n_rays * MAX_SAMPLE_PER_RAY));
  RayMarchKernel<true><<<grid, block>>>(n_rays, sample_l, scale_by_dis != 0, (Wec3f*)rays_o, (Wec3f*)rays_d, noise,
                                        (Wec2i64*)oct_idx_start_end, oct_idx, (Wec2f*)oct_nf, (TreeNode*)tree_nodes,
                                        (TransInfo*)transes, (Wec2i64*)pts_idx_start_end, (Wec3f*)world_pts,
                                        (Wec3f*)warp_pts, (Wec3f*)dirs, (Wec3i64*)anchors, dists, ts, sampled_oct,
                                        first_oct_dis);
  CK(cudaGetLastError());
  CK(cudaDeviceSynchronize());
  cudaFree(bounds);
  cudaFree(counter);
  cudaFree(oct_idx);
  cudaFree(oct_nf);
  cudaFree(sampled_oct);
  return 0;
}

/* MarkVistNodeKernel alone (:518-574); adders [n_nodes] = -1, mark = 0 prepared by the caller. */
int refcu_mark_visit(int64_t n_rays, int64_t* pts_idx_start_end, int64_t* oct_indices, float* weights, float* alphas,
                     int64_t* weight_adder, int64_t* alpha_adder, int64_t* visit_mark, int64_t* visit_cnt) {
  MarkVistNodeKernel<<<div_up(n_rays, THREADS), THREADS>>>(n_rays, pts_idx_start_end, oct_indices, weights, alphas,
                                                           weight_adder, alpha_adder, visit_mark, visit_cnt);
  CK(cudaGetLastError());
  return (int)cudaDeviceSynchronize();
}

int refcu_trans_query_frame(int64_t n_pts, void* tree_nodes, int64_t n_nodes, void* transes, int64_t* anchors,
                            float* world_pts, float* out) {
  TransQueryFrameKernel<<<div_up(n_pts, THREADS), THREADS>>>(n_pts, n_nodes, (TreeNode*)tree_nodes,
                                                             (TransInfo*)transes, anchors, (Wec3f*)world_pts,
                                                             (Wec3f*)out);
  CK(cudaGetLastError());
  return (int)cudaDeviceSynchronize();
}

/* PersOctree::MarkInvisibleNodes (:707-742) and the kernel of PersOctree::UpdateBlockIdxs (:746-798), in place on the
 * device node blob */
int refcu_mark_invisible_nodes(int64_t n_nodes, int64_t n_cams, void* tree_nodes, float* intri, float* w2c,
                               float* bounds) {
  MarkInvisibleNodesKernel<<<div_up(n_nodes, THREADS), THREADS>>>(n_nodes, n_cams, (TreeNode*)tree_nodes,
                                                                  (Watrix33f*)intri, (Watrix34f*)w2c, (Wec2f*)bounds);
  CK(cudaGetLastError());
  return (int)cudaDeviceSynchronize();
}

int refcu_set_block_idxs(int64_t n_nodes, int64_t n_blocks, void* tree_nodes, float* centers) {
  SetBlockIdxsNearestKernel<<<div_up(n_nodes, THREADS), THREADS>>>(n_nodes, n_blocks, (TreeNode*)tree_nodes,
                                                                   (Wec3f*)centers);
  CK(cudaGetLastError());
  return (int)cudaDeviceSynchronize();
}

} /* extern "C" */
