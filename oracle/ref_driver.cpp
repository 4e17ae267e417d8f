/*
 * ref_driver.cpp -- runs the REFERENCE's own device functions on the host.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  `make -C oracle ref` (1) extracts the __global__ / __device__ function bodies
 * of gfnerf/bindings/field/Hash3DAnchored_cuda.cu and gfnerf/bindings/PtsSampler/PersSampler_cuda.cu from
 * /root/reference into oracle/_ref/ref_kernels.inc (oracle/ref_extract.py; build output, git-ignored, deleted after
 * the link), (2) compiles this file, which #includes that text unmodified between two shims written for this
 * repository -- ref_shim/cuda_host_shim.h (threadIdx / blockIdx, atomics, __half) and ref_shim/eigen_subset.h (the
 * fixed-size Eigen types the kernels use) -- into oracle/_ref/libgf_ref_host.so.
 *
 * What this pins: the ALGORITHM of every kernel on the hot path -- hash index arithmetic and blend, the DFS over the
 * octree with its child ordering and stack discipline, the march loop with its step rounding, the vote rules -- is
 * the reference's own code here, not a restatement.  tests/test_ref_kernels.py holds oracle/gf_oracle.c to it.
 * What it cannot pin: the last bit of fp32 results.  nvcc contracts mul+add pairs into FMAs where it chooses to
 * (-fmad=true) and g++ where it chooses to (-ffp-contract); the two shared objects built here (contraction off /
 * fast) bracket that, and Eigen's evaluation order inside the shim is our reading of Eigen 3.4 (see its header).
 *
 * The entry points restate only the LAUNCH SEQUENCES of the reference's host functions, which are torch code
 * (tensors, `<<< >>>`): Hash3DAnchoredFunction::forward / backward (Hash3DAnchored_cuda.cu:160-239),
 * PersSampler::GetSamples (PersSampler_cuda.cu:321-477), UpdateOctNodes (:584-677), TransQueryFrame (:895-921),
 * GetPointsAnchors (:923-983), GetEdgeSamples (:497-515).
 */
#include <stdint.h>
#include <stdlib.h>

#include <vector>

#include "ref_shim/cuda_host_shim.h"
#include "ref_shim/eigen_subset.h"

#include "_ref/ref_kernels.inc"

static_assert(sizeof(Wec3f) == 12 && sizeof(Wec2f) == 8 && sizeof(Wec2i64) == 16 && sizeof(Wec3i64) == 24, "vector layout");
static_assert(sizeof(TransInfo) == 576, "TransInfo layout (PersSampler.h:31-38)");
static_assert(sizeof(TreeNode) == 128, "TreeNode layout (PersSampler.h:40-49)");
static_assert(offsetof(TransInfo, weight) == 384 && offsetof(TransInfo, center) == 528 &&
              offsetof(TransInfo, dis_summary) == 544, "TransInfo offsets");
static_assert(offsetof(TreeNode, childs) == 24 && offsetof(TreeNode, is_leaf_node) == 88 &&
              offsetof(TreeNode, trans_idx) == 96 && offsetof(TreeNode, block_idx) == 104, "TreeNode offsets");

#define THREADS 512u /* Utils/Common.h:36 THREAD_CAP */
static inline unsigned div_up(int64_t a, int64_t b) { return (unsigned)((a + b - 1) / b); }

extern "C" {

const char* ref_build_flavour(void) {
#ifdef __FP_FAST_FMAF
  return "g++ host build; fp contraction per -ffp-contract (FMA hardware available)";
#else
  return "g++ host build; no FMA";
#endif
}

/* host threads a launch is spread over (1 = serial and deterministic: the default, what the tests use) */
void ref_set_threads(int n) { gf_ref_threads = n < 1 ? 1 : n; }
int ref_get_threads(void) { return gf_ref_threads; }

/* Hash3DAnchoredFunction::forward, Hash3DAnchored_cuda.cu:160-196.  feat_pool_f16 = feat_pool.to(kFloat16) (:185),
 * out_f16 [n,32] zero-filled by the caller (:182); the caller widens it to fp32 (:195). */
void ref_hash_forward(int n_points, int n_volumes, void* feat_pool_f16, int* prim_pool, int* feat_local_idx,
                      int* feat_local_size, float* bias_pool, float* points, int64_t* volume_idx, void* out_f16) {
  gf_launch(div_up(n_points, THREADS), N_LEVELS, THREADS, [&] {
    Hash3DAnchoredForwardKernel<__half>(n_points, n_volumes, (__half*)feat_pool_f16, prim_pool, feat_local_idx,
                                        feat_local_size, (Wec3f*)bias_pool, (Wec3f*)points, volume_idx,
                                        (__half*)out_f16);
  });
}

/* Hash3DAnchoredFunction::backward, :198-239.  grad_in_f16 = (grad * 128).to(kFloat16) (:219); grad_out_f16
 * [pool_size, 2] zero-filled (:221); the caller returns grad_out.to(fp32) / 128 (:238).  The half2 atomics land in
 * the emulation's serial order (level-major, points ascending) -- one of the orders the GPU may produce. */
void ref_hash_backward(int n_points, int n_volumes, int* prim_pool, int* feat_local_idx, int* feat_local_size,
                       float* bias_pool, float* points, int64_t* volume_idx, void* grad_in_f16, void* grad_out_f16) {
  gf_launch(div_up(n_points, THREADS), N_LEVELS, THREADS, [&] {
    Hash3DAnchoredBackwardKernel<__half>(n_points, n_volumes, prim_pool, feat_local_idx, feat_local_size,
                                         (Wec3f*)bias_pool, (Wec3f*)points, volume_idx, (__half*)grad_in_f16,
                                         (__half*)grad_out_f16);
  });
}

/* PersSampler::GetSamples, PersSampler_cuda.cu:321-477.  rays_d already normalised (:323), noise already
 * multiplied by ray_march_fineness_ (:380-389).  Dense outputs [R,1024,..] zero-filled by the caller (:437-444).
 * pts_idx_start_end [R,2] as the reference returns it; n_oct [R], oct_idx / oct_nf dense [R, max_oct] (optional)
 * expose the leaf lists of the traversal.  The counting pass hands out leaf-list ranges with an atomicAdd in ray
 * order here (any order is a valid GPU outcome; the per-ray lists do not depend on it). */
void ref_get_samples(int64_t n_rays, float* rays_o, float* rays_d, float* noise, void* tree_nodes, void* transes,
                     uint8_t* search_order, float global_near, float sample_l, int scale_by_dis,
                     int64_t max_oct_per_ray, float* world_pts, float* warp_pts, float* dirs, float* dists, float* ts,
                     int64_t* anchors, int64_t* pts_idx_start_end, float* first_oct_dis, int64_t* n_oct,
                     int64_t* oct_idx_dense, float* oct_nf_dense) {
  std::vector<float> bounds(2 * n_rays);
  for (int64_t i = 0; i < n_rays; i++) {
    bounds[2 * i] = global_near;
    bounds[2 * i + 1] = 1e8f;
  }
  int64_t counter = 0;
  std::vector<int64_t> oct_se(2 * n_rays, 0), stack_info(1);
  const unsigned grid = div_up(n_rays, THREADS);
  gf_launch(grid, 1, THREADS, [&] {
    FindRayOctreeIntersectionKernel<false>(n_rays, max_oct_per_ray, search_order, (Wec3f*)rays_o, (Wec3f*)rays_d,
                                           (Wec2f*)bounds.data(), &counter, (Wec2i64*)oct_se.data(),
                                           (TreeNode*)tree_nodes, nullptr, nullptr, stack_info.data());
  });
  std::vector<int64_t> oct_idx(counter > 0 ? counter : 1);
  std::vector<float> oct_nf(2 * (counter > 0 ? counter : 1));
  gf_launch(grid, 1, THREADS, [&] {
    FindRayOctreeIntersectionKernel<true>(n_rays, max_oct_per_ray, search_order, (Wec3f*)rays_o, (Wec3f*)rays_d,
                                          (Wec2f*)bounds.data(), &counter, (Wec2i64*)oct_se.data(),
                                          (TreeNode*)tree_nodes, oct_idx.data(), (Wec2f*)oct_nf.data(),
                                          stack_info.data());
  });
  for (int64_t i = 0; i < n_rays; i++) {
    const int64_t s = oct_se[2 * i], e = oct_se[2 * i + 1];
    if (n_oct) n_oct[i] = e - s;
    for (int64_t k = s; k < e && k - s < max_oct_per_ray; k++) {
      if (oct_idx_dense) oct_idx_dense[i * max_oct_per_ray + (k - s)] = oct_idx[k];
      if (oct_nf_dense) {
        oct_nf_dense[(i * max_oct_per_ray + (k - s)) * 2] = oct_nf[2 * k];
        oct_nf_dense[(i * max_oct_per_ray + (k - s)) * 2 + 1] = oct_nf[2 * k + 1];
      }
    }
  }
  std::vector<int64_t> pts_se(2 * n_rays, 0);
  gf_launch(grid, 1, THREADS, [&] {
    RayMarchKernel<false>(n_rays, sample_l, scale_by_dis != 0, (Wec3f*)rays_o, (Wec3f*)rays_d, noise,
                          (Wec2i64*)oct_se.data(), oct_idx.data(), (Wec2f*)oct_nf.data(), (TreeNode*)tree_nodes,
                          (TransInfo*)transes, (Wec2i64*)pts_se.data(), nullptr, nullptr, nullptr, nullptr, nullptr,
                          nullptr, nullptr, nullptr);
  });
  /* :420 pts_idx_start_end[:,0] = cumsum(pts_idx_start_end[:,0]) */
  int64_t run = 0;
  for (int64_t i = 0; i < n_rays; i++) {
    run += pts_se[2 * i];
    pts_se[2 * i] = run;
  }
  std::vector<int64_t> sampled_oct(n_rays * (int64_t)MAX_SAMPLE_PER_RAY, -1);
  gf_launch(grid, 1, THREADS, [&] {
    RayMarchKernel<true>(n_rays, sample_l, scale_by_dis != 0, (Wec3f*)rays_o, (Wec3f*)rays_d, noise,
                         (Wec2i64*)oct_se.data(), oct_idx.data(), (Wec2f*)oct_nf.data(), (TreeNode*)tree_nodes,
                         (TransInfo*)transes, (Wec2i64*)pts_se.data(), (Wec3f*)world_pts, (Wec3f*)warp_pts,
                         (Wec3f*)dirs, (Wec3i64*)anchors, dists, ts, sampled_oct.data(), first_oct_dis);
  });
  for (int64_t i = 0; i < 2 * n_rays; i++) pts_idx_start_end[i] = pts_se[i];
}

/* PersSampler::UpdateOctNodes up to MarkInvalidNodes, :584-660: MarkVistNodeKernel, then the torch expressions
 *   stats = max(stats, mask * adder); stats += mark * (1 - mask) * adder; clamp(-100, 2^20)   with mask = adder > 0
 * for the weight and the alpha statistics, then MarkInvalidNodes.  In place on tree_nodes / the three stat arrays. */
void ref_update_oct_nodes(int64_t n_rays, int64_t* pts_idx_start_end, int64_t* oct_indices, float* weights,
                          float* alphas, void* tree_nodes, int64_t n_nodes, int64_t* weight_stats,
                          int64_t* alpha_stats, int64_t* visit_cnt) {
  std::vector<int64_t> w_add(n_nodes, -1), a_add(n_nodes, -1), mark(n_nodes, 0);
  gf_launch(div_up(n_rays, THREADS), 1, THREADS, [&] {
    MarkVistNodeKernel(n_rays, pts_idx_start_end, oct_indices, weights, alphas, w_add.data(), a_add.data(),
                       mark.data(), visit_cnt);
  });
  for (int64_t i = 0; i < n_nodes; i++) {
    int64_t* stats[2] = {weight_stats + i, alpha_stats + i};
    const int64_t adder[2] = {w_add[i], a_add[i]};
    for (int k = 0; k < 2; k++) {
      const int64_t m = adder[k] > 0 ? 1 : 0;
      int64_t s = *stats[k];
      s = std::max(s, m * adder[k]);
      s += mark[i] * (1 - m) * adder[k];
      s = std::min<int64_t>(std::max<int64_t>(s, -100), 1 << 20);
      *stats[k] = s;
    }
  }
  gf_launch(div_up(n_nodes, THREADS), 1, THREADS,
            [&] { MarkInvalidNodes(n_nodes, weight_stats, alpha_stats, (TreeNode*)tree_nodes); });
}

/* PersSampler::TransQueryFrame, :895-921.  out [n,3] zero-filled by the caller. */
void ref_trans_query_frame(int64_t n_pts, void* tree_nodes, int64_t n_nodes, void* transes, int64_t* anchors,
                           float* world_pts, float* out) {
  gf_launch(div_up(n_pts, THREADS), 1, THREADS, [&] {
    TransQueryFrameKernel(n_pts, n_nodes, (TreeNode*)tree_nodes, (TransInfo*)transes, anchors, (Wec3f*)world_pts,
                          (Wec3f*)out);
  });
}

/* PersSampler::GetPointsAnchors, :923-983.  t_cur = (t_starts + t_ends) / 2 is the caller's; anchors [R,S] = -1. */
void ref_points_anchors(int64_t n_rays, int64_t n_pts_per_ray, float* rays_o, float* rays_d, float* t_cur,
                        void* tree_nodes, int64_t n_nodes, int64_t* anchors) {
  std::vector<float> ts(2 * n_rays * n_nodes, 0.f);
  const unsigned grid = div_up(n_rays * n_nodes, THREADS);
  gf_launch(grid, 1, THREADS, [&] {
    GetRaysTreeNodesIntersectsKernel(n_rays, n_nodes, (TreeNode*)tree_nodes, (Wec3f*)rays_o, (Wec3f*)rays_d,
                                     ts.data());
  });
  gf_launch(grid, 1, THREADS,
            [&] { GetTreeNodeIdxFromTsKernel(n_rays, n_nodes, n_pts_per_ray, t_cur, ts.data(), anchors); });
}

/* PersSampler::GetEdgeSamples' kernel, :479-495 (the host part draws the random edge indices / coordinates). */
void ref_edge_samples(int64_t n_pts, void* edge_pool, void* transes, int64_t* edge_indices, float* edge_coords,
                      float* out_pts, int64_t* out_idx) {
  gf_launch(div_up(n_pts, THREADS), 1, THREADS, [&] {
    GetEdgeSamplesKernel(n_pts, (EdgePool*)edge_pool, (TransInfo*)transes, edge_indices, (Wec2f*)edge_coords,
                         (Wec3f*)out_pts, out_idx);
  });
}

/* PersOctree::MarkInvisibleNodes, :707-742 */
void ref_mark_invisible_nodes(int64_t n_nodes, int64_t n_cams, void* tree_nodes, float* intri, float* w2c,
                              float* bounds) {
  gf_launch(div_up(n_nodes, THREADS), 1, THREADS, [&] {
    MarkInvisibleNodesKernel(n_nodes, n_cams, (TreeNode*)tree_nodes, (Watrix33f*)intri, (Watrix34f*)w2c,
                             (Wec2f*)bounds);
  });
}

/* PersOctree::UpdateBlockIdxs' kernel, :746-766 */
void ref_set_block_idxs(int64_t n_nodes, int64_t n_blocks, void* tree_nodes, float* centers) {
  gf_launch(div_up(n_nodes, THREADS), 1, THREADS,
            [&] { SetBlockIdxsNearestKernel(n_nodes, n_blocks, (TreeNode*)tree_nodes, (Wec3f*)centers); });
}

int64_t ref_sizeof_edge_pool(void) { return (int64_t)sizeof(EdgePool); }

} /* extern "C" */

/* ---- host member functions of PersOctree, PtsSampler/PersSampler.cpp:154-417, 833-895 ----------------------------
 * The class below declares only the members those two bodies touch (PtsSampler/PersSampler.h:51-95); the bodies
 * themselves are the reference's text (ref_host_fns.inc), compiled against ref_shim/torch_stub.h. */
#include "ref_shim/torch_stub.h"

class PersOctree {
  using Tensor = torch::Tensor;

 public:
  void ProcOctree(bool compact, bool subdivide, bool brute_force);
  void ConstructEdgePool();
  std::vector<TreeNode> tree_nodes_;
  std::vector<EdgePool> edge_pool_;
  Tensor tree_nodes_gpu_, tree_weight_stats_, tree_alpha_stats_, tree_visit_cnt_;
};

#include "_ref/ref_host_fns.inc"

extern "C" {

/* PersOctree::ProcOctree on blobs.  Call with nodes_out == NULL to learn the node count, then again with buffers of
 * that size (the function is deterministic).  Returns the new node count, or -(line) of a failed reference CHECK. */
int64_t ref_proc_octree(const void* nodes_in, int64_t n_in, const int64_t* weight_stats, const int64_t* alpha_stats,
                        const int64_t* visit_cnt, int compact, int subdivide, int brute_force, void* nodes_out,
                        int64_t* weight_out, int64_t* alpha_out, int64_t capacity) {
  PersOctree oc;
  oc.tree_nodes_.resize((size_t)n_in);
  std::memcpy((void*)oc.tree_nodes_.data(), nodes_in, (size_t)n_in * sizeof(TreeNode));
  oc.tree_nodes_gpu_ = torch::from_blob((void*)nodes_in, {n_in * (int64_t)sizeof(TreeNode)}, CPUUInt8);
  oc.tree_weight_stats_ = torch::from_blob((void*)weight_stats, {n_in}, CPUInt64);
  oc.tree_alpha_stats_ = torch::from_blob((void*)alpha_stats, {n_in}, CPUInt64);
  oc.tree_visit_cnt_ = torch::from_blob((void*)visit_cnt, {n_in}, CPUInt64);
  try {
    oc.ProcOctree(compact != 0, subdivide != 0, brute_force != 0);
  } catch (const RefCheckFailure& f) {
    return -(int64_t)f.line;
  }
  const int64_t n = (int64_t)oc.tree_nodes_.size();
  if (nodes_out && n <= capacity) {
    std::memcpy(nodes_out, oc.tree_nodes_gpu_.data_ptr(), (size_t)n * sizeof(TreeNode));
    std::memcpy(weight_out, oc.tree_weight_stats_.data_ptr(), (size_t)n * 8);
    std::memcpy(alpha_out, oc.tree_alpha_stats_.data_ptr(), (size_t)n * 8);
  }
  return n;
}

/* PersOctree::ConstructEdgePool on a node blob; same two-call protocol.  Returns the number of 64-byte entries. */
int64_t ref_construct_edge_pool(const void* nodes_in, int64_t n_in, void* pool_out, int64_t capacity) {
  PersOctree oc;
  oc.tree_nodes_.resize((size_t)n_in);
  std::memcpy((void*)oc.tree_nodes_.data(), nodes_in, (size_t)n_in * sizeof(TreeNode));
  oc.ConstructEdgePool();
  const int64_t n = (int64_t)oc.edge_pool_.size();
  if (pool_out && n <= capacity) std::memcpy(pool_out, oc.edge_pool_.data(), (size_t)n * sizeof(EdgePool));
  return n;
}

} /* extern "C" */
