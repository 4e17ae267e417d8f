/*
 * ref_driver_torch.cpp -- the REFERENCE's PersOctree construction, run on the CPU against this image's real libtorch.
 *
 * TEST INFRASTRUCTURE ONLY (oracle/).  PersOctree::PersOctree / ConstructTreeNode / GetVisiCams / DistanceSummary /
 * PCA / ConstructTrans / ProcOctree / ConstructEdgePool (PtsSampler/PersSampler.cpp:9-417, 516-895) are torch tensor
 * code, so unlike the kernels they are compiled with the real <torch/torch.h>; the only substitutions are
 *   - `kCUDA` -> `kCPU` (one token: every `.to(torch::kCUDA)` and every CUDA* option macro of Utils/Common.h lands
 *     on the CPU device; the arithmetic is the same ATen operator set),
 *   - the fixed-size Eigen subset of ref_shim/eigen_subset.h for the un-vendored Eigen,
 *   - a no-op ScopeWatch / PRINT_VAL.
 * The text itself comes from /root/reference at build time (oracle/ref_extract.py -> _ref/ref_octree_src.inc, deleted
 * after the link).  `make -C oracle ref_octree` -> oracle/_ref/libgf_ref_octree.so.  tests/test_ref_octree.py holds
 * gf_octree_build (csrc/octree_build.cu) and the search-order table to it.
 */
#include <torch/torch.h>

#include <cmath>
#include <cstring>
#include <functional>
#include <iostream>
#include <memory>
#include <tuple>
#include <vector>

#include "ref_shim/eigen_subset.h"

#define kCUDA kCPU
#define PRINT_VAL(x) do { } while (false)
struct ScopeWatch {
  explicit ScopeWatch(const char*) {}
};

#include "_ref/ref_octree_src.inc"

static_assert(sizeof(TransInfo) == 576 && sizeof(TreeNode) == 128 && sizeof(EdgePool) == 64, "reference struct layout");

static torch::Tensor from_f32(const float* p, std::initializer_list<int64_t> shape) {
  return torch::from_blob((void*)p, shape, torch::TensorOptions().dtype(torch::kFloat32)).clone();
}

extern "C" {

/* PersOctree::PersOctree on host arrays (c2w [n,3,4], w2c [n,3,4], intri [n,3,3], bound [n,2]).  torch's generator is
 * seeded with `seed` first (the constructor draws its sample points and first cameras from it).  Two-call protocol:
 * nodes_out == NULL returns the sizes only.  search_order_out: 64 bytes. */
int64_t refoct_build(int64_t max_depth, float bbox_side_len, float split_dist_thres, const float* c2w,
                     const float* w2c, const float* intri, const float* bound, int64_t n_cams, uint64_t seed,
                     void* nodes_out, void* trans_out, uint8_t* search_order_out, int64_t* n_trans) {
  static std::unique_ptr<PersOctree> cache;
  static uint64_t cache_seed = ~0ull;
  try {
    if (!cache || cache_seed != seed || !nodes_out) {
      torch::manual_seed(seed);
      cache = std::make_unique<PersOctree>(max_depth, bbox_side_len, split_dist_thres, from_f32(c2w, {n_cams, 3, 4}),
                                           from_f32(w2c, {n_cams, 3, 4}), from_f32(intri, {n_cams, 3, 3}),
                                           from_f32(bound, {n_cams, 2}));
      cache_seed = seed;
    }
  } catch (const std::exception& e) {
    std::cerr << "refoct_build: " << e.what() << std::endl;
    return -1;
  }
  *n_trans = (int64_t)cache->pers_trans_.size();
  const int64_t n = (int64_t)cache->tree_nodes_.size();
  if (nodes_out) {
    std::memcpy(nodes_out, cache->tree_nodes_.data(), (size_t)n * sizeof(TreeNode));
    std::memcpy(trans_out, cache->pers_trans_.data(), (size_t)*n_trans * sizeof(TransInfo));
    std::memcpy(search_order_out, cache->node_search_order_.data_ptr<uint8_t>(), 64);
    cache.reset();
  }
  return n;
}

/* PersOctree::ConstructTrans on given sample points and cameras; the first virtual camera is torch::randint's draw
 * under `seed`.  out: one 576-byte TransInfo (side_len left 0, as ConstructTrans leaves it to its caller). */
int refoct_construct_trans(const float* rand_pts, int64_t n_pts, const float* c2w, int64_t n_cams, const float* intri0,
                           const float* center, uint64_t seed, void* out) {
  try {
    torch::manual_seed(seed);
    PersOctree* oc = (PersOctree*)::operator new(sizeof(PersOctree));   // ConstructTrans reads no member
    std::memset((void*)oc, 0, sizeof(PersOctree));
    TransInfo t = oc->ConstructTrans(from_f32(rand_pts, {n_pts, 3}), from_f32(c2w, {n_cams, 3, 4}),
                                     from_f32(intri0, {3, 3}), from_f32(center, {3}));
    ::operator delete((void*)oc);
    t.side_len = 0.f;
    std::memcpy(out, &t, sizeof(TransInfo));
    return 0;
  } catch (const std::exception& e) {
    std::cerr << "refoct_construct_trans: " << e.what() << std::endl;
    return -1;
  }
}

} /* extern "C" */
