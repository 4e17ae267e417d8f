// TorchScript custom classes `torch.classes.my_classes.{Hash3DAnchored, PersSampler}` over the C-ABI.
//
// This is the reference's own operator boundary (gfnerf/bindings/hashanchored/bindings.cpp:343-401): the reference's
// Python does `torch.classes.load_library(".../f2nerf-bindings.so")` and then instantiates these two classes
// (gfnerf/hash_3d_anchored.py:13-25, gfnerf/perssampler.py:30-31,103-123).  Loading THIS library instead gives the
// same class names and method signatures, with every method body reduced to pointer extraction + one gf_* call on
// torch's current stream.  No torch type crosses into libgfnerf_b200.so.
//
//  * Hash3DAnchored: the complete surface (ctor, AnchoredQuery with autograd, GetParams, States, LoadStates, Reset,
//    Zero, SetFeatPoolRequireGrad, to, ReleaseResources), state tensors and their order as in
//    field/Hash3DAnchored.cpp:17-200.
//  * PersSampler: InitSampler (octree + leaf transforms built on the host by gf_octree_build, as the reference builds
//    them on the host in PersSampler.cpp:92-152, 516-831), GetSamples, UpdateOctNodes (vote, statistics, pruning and
//    the milestone / compact_freq ProcOctree through gf_octree_proc), UpdateRayMarch, UpdateMode, States, LoadStates,
//    trans_query_frame, the getters.
#include <torch/custom_class.h>
#include <torch/script.h>
#include <torch/cuda.h>
#include <c10/cuda/CUDAGuard.h>
#include <c10/cuda/CUDAStream.h>

#include <cmath>
#include <fstream>
#include <random>

#include "../../include/gfnerf_b200.h"

#define GF_CHECK(call) TORCH_CHECK((call) == GF_OK, "gfnerf_b200: ", gf_last_error())

namespace {

using torch::Tensor;

bool is_prime(uint32_t x) {
  if (x < 4) return x > 1;
  if (x % 2 == 0) return false;
  for (uint32_t i = 3; (uint64_t)i * i <= x; i += 2)
    if (x % i == 0) return false;
  return true;
}

struct Hash3DAnchoredImpl : torch::CustomClassHolder {
  Tensor feat_pool_, prim_pool_, bias_pool_, shadow_, scales_;
  int64_t pool_size_, local_size_, n_volumes_;

  Hash3DAnchoredImpl(int64_t log2_table_size, int64_t n_volumes, double /*learn_rate*/) {
    TORCH_CHECK(torch::cuda::is_available(), "Hash3DAnchored needs a CUDA device: the B200 kernels have no CPU fallback");
    auto f32 = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCUDA);
    pool_size_ = (int64_t(1) << log2_table_size) * GF_N_LEVELS;
    n_volumes_ = n_volumes;
    // Hash3DAnchored.cpp:26  (rand * .2 - 1) * 1e-4
    feat_pool_ = (torch::rand({pool_size_, GF_N_CHANNELS}, f32) * .2f - 1.f) * 1e-4f;
    feat_pool_.requires_grad_(true);
    // :32-55  3 * 16 * n_volumes random primes in [2^28, 2^30)
    std::mt19937 rng(1234);
    std::uniform_int_distribution<uint32_t> dist(1u << 28, (1u << 30) - 1);
    auto prim = torch::empty({GF_N_LEVELS, n_volumes, 3}, torch::kInt32);
    int32_t* p = prim.data_ptr<int32_t>();
    for (int64_t i = 0; i < prim.numel(); i++) {
      uint32_t x = dist(rng) | 1u;
      while (!is_prime(x) || x >= (1u << 30)) x = x + 2 >= (1u << 30) ? (dist(rng) | 1u) : x + 2;
      p[i] = (int32_t)x;
    }
    prim_pool_ = prim.to(torch::kCUDA);
    bias_pool_ = torch::zeros({GF_N_LEVELS * n_volumes, 3}, f32);  // rand_bias is never set (:57-62)
    local_size_ = ((pool_size_ / GF_N_LEVELS) >> 4) << 4;           // :66-70
    scales_ = torch::empty({GF_N_LEVELS}, f32);
    GF_CHECK(gf_hash_level_scales(scales_.data_ptr<float>(), nullptr, stream()));
    shadow_ = torch::empty({pool_size_, GF_N_CHANNELS}, f32.dtype(torch::kFloat16));
  }

  static void* stream() { return (void*)c10::cuda::getCurrentCUDAStream().stream(); }

  void cast_table() {  // Hash3DAnchored_cuda.cu:185: the table is cast to fp16 on every forward
    GF_CHECK(gf_hash_cast_table(feat_pool_.data_ptr<float>(), shadow_.data_ptr(), feat_pool_.numel(), stream()));
  }

  Tensor forward_raw(const Tensor& points, const Tensor& anchors) {
    TORCH_CHECK(points.is_cuda() && anchors.is_cuda(), "gfnerf_b200: tensor is not on a CUDA device (there is no CPU path)");
    TORCH_CHECK(points.dim() == 2 && points.size(1) == 3 && points.scalar_type() == torch::kFloat32, "points: f32 [n,3]");
    TORCH_CHECK(anchors.dim() == 1 && anchors.scalar_type() == torch::kInt64, "anchors: i64 [n]");
    c10::cuda::CUDAGuard guard(points.device());
    auto pts = points.contiguous();
    auto anc = anchors.contiguous();
    cast_table();
    auto out = torch::empty({pts.size(0), GF_HASH_DIM}, pts.options());
    GF_CHECK(gf_hash_forward(pts.size(0), nullptr, (int32_t)n_volumes_, local_size_, shadow_.data_ptr(),
                             prim_pool_.data_ptr<int32_t>(), bias_pool_.data_ptr<float>(), scales_.data_ptr<float>(),
                             pts.data_ptr<float>(), anc.data_ptr<int64_t>(), 1, nullptr, out.data_ptr<float>(),
                             stream()));
    return out;
  }

  Tensor backward_raw(const Tensor& points, const Tensor& anchors, const Tensor& grad_out) {
    c10::cuda::CUDAGuard guard(points.device());
    auto g = grad_out.contiguous().to(torch::kFloat32);
    auto grad_table = torch::zeros_like(feat_pool_);
    GF_CHECK(gf_hash_backward(points.size(0), nullptr, (int32_t)n_volumes_, local_size_, prim_pool_.data_ptr<int32_t>(),
                              bias_pool_.data_ptr<float>(), scales_.data_ptr<float>(), points.data_ptr<float>(),
                              anchors.data_ptr<int64_t>(), 1, g.data_ptr<float>(), 0, grad_table.data_ptr<float>(),
                              stream()));
    return grad_table;
  }

  Tensor AnchoredQuery(const Tensor& points, const Tensor& anchors);

  std::vector<Tensor> GetParams() { return {feat_pool_}; }
  std::vector<Tensor> States() {
    return {feat_pool_.detach(), prim_pool_, bias_pool_, torch::full({1}, n_volumes_, torch::kInt32)};
  }
  int64_t LoadStates(const std::vector<Tensor>& states, int64_t idx) {
    {
      torch::NoGradGuard ng;
      feat_pool_.copy_(states[idx++]);
    }
    prim_pool_ = states[idx++].clone().to(torch::kCUDA).contiguous();
    bias_pool_ = states[idx++].clone().to(torch::kCUDA).contiguous();
    n_volumes_ = states[idx++].item<int64_t>();
    return idx;
  }
  void Reset() {
    torch::NoGradGuard ng;
    feat_pool_.uniform_(-1e-2, 1e-2);  // :171-174
  }
  void Zero() {
    torch::NoGradGuard ng;
    feat_pool_.zero_();
  }
  void SetFeatPoolRequireGrad(bool require_grad) { feat_pool_.requires_grad_(require_grad); }
  void to(std::string device) {  // :180-200: "cpu" parks the table on the host, anything else is cuda
    const bool rg = feat_pool_.requires_grad();
    feat_pool_ = feat_pool_.detach().to(device == "cpu" ? torch::kCPU : torch::kCUDA).requires_grad_(rg);
  }
  void ReleaseResources() {
    feat_pool_ = Tensor();
    prim_pool_ = Tensor();
    bias_pool_ = Tensor();
    shadow_ = Tensor();
  }
};

// Hash3DAnchoredFunction (Hash3DAnchored_cuda.cu:160-239).  Unlike the reference the query points / anchors are saved
// on the autograd node, not on the encoder object, so two forwards before a backward do not corrupt each other.
struct AnchoredQueryFn : torch::autograd::Function<AnchoredQueryFn> {
  static Tensor forward(torch::autograd::AutogradContext* ctx, Tensor feat_pool, Tensor points, Tensor anchors,
                        c10::intrusive_ptr<Hash3DAnchoredImpl> self) {
    auto pts = points.contiguous();
    auto anc = anchors.contiguous();
    ctx->save_for_backward({pts, anc});
    ctx->saved_data["self"] = self;
    return self->forward_raw(pts, anc);
  }
  static torch::autograd::tensor_list backward(torch::autograd::AutogradContext* ctx,
                                                torch::autograd::tensor_list grad_outputs) {
    auto saved = ctx->get_saved_variables();
    auto self = ctx->saved_data["self"].toCustomClass<Hash3DAnchoredImpl>();
    return {self->backward_raw(saved[0], saved[1], grad_outputs[0]), Tensor(), Tensor(), Tensor()};
  }
};

Tensor Hash3DAnchoredImpl::AnchoredQuery(const Tensor& points, const Tensor& anchors) {
  if (feat_pool_.requires_grad() && torch::GradMode::is_enabled())
    return AnchoredQueryFn::apply(feat_pool_, points, anchors, c10::intrusive_ptr<Hash3DAnchoredImpl>::reclaim_copy(this));
  return forward_raw(points, anchors);
}

// ---------------------------------------------------------------------------------------------------------------
struct PersSamplerImpl : torch::CustomClassHolder {
  Tensor tree_nodes_, pers_trans_, visit_cnt_, weight_stats_, alpha_stats_, search_order_;
  Tensor w2c_, intri_, bound_;               // cameras, for MarkInvisibleNodes at the subdivision milestones
  Tensor edge_pool_;                         // 64-byte EdgePool records (built lazily, see GetEdgeSamples)
  std::vector<int64_t> sub_div_milestones_;  // reversed: the next milestone is at the back
  double global_near_ = 0.01, sample_l_ = 1.0 / 256, fineness_ = 1.0, init_fineness_ = 16.0, decay_end_ = 10000.0;
  double sampled_oct_per_ray_ = 512.0;
  int64_t mode_ = 0, max_oct_ = 1024, compact_freq_ = 1000;
  bool scale_by_dis_ = true;

  PersSamplerImpl() {}
  static void* stream() { return (void*)c10::cuda::getCurrentCUDAStream().stream(); }
  int64_t n_nodes() const { return tree_nodes_.numel() / GF_TREE_NODE_BYTES; }
  int64_t n_trans() const { return pers_trans_.numel() / GF_TRANS_INFO_BYTES; }

  // PersSampler::PersSampler + PersOctree::PersOctree (PersSampler.cpp:899-952, 92-152): the octree and the leaf
  // transforms are built on the host by gf_octree_build (csrc/octree_build.cu), as the reference builds them on the host
  void InitSampler(double split_dist_thres, std::vector<int64_t> sub_div_milestones, int64_t compact_freq,
                   int64_t max_oct_intersect_per_ray, double global_near, bool scale_by_dis, int64_t bbox_levels,
                   double sample_l, int64_t max_level, Tensor c2w, Tensor w2c, Tensor intri, Tensor bounds, int64_t mode,
                   double sampled_oct_per_ray, double ray_march_fineness, double ray_march_init_fineness,
                   int64_t ray_march_fineness_decay_end_iter) {
    TORCH_CHECK(torch::cuda::is_available(), "PersSampler needs a CUDA device: the B200 kernels have no CPU fallback");
    Configure(compact_freq, max_oct_intersect_per_ray, global_near, scale_by_dis, sample_l, mode, ray_march_fineness,
              ray_march_init_fineness, ray_march_fineness_decay_end_iter);
    sampled_oct_per_ray_ = (double)sampled_oct_per_ray;
    sub_div_milestones_.assign(sub_div_milestones.rbegin(), sub_div_milestones.rend());  // popped from the back
    auto f32c = torch::TensorOptions().dtype(torch::kFloat32).device(torch::kCPU);
    auto c2w_h = c2w.detach().to(f32c).contiguous(), intri_h = intri.detach().to(f32c).contiguous(),
         bounds_h = bounds.detach().to(f32c).contiguous();
    TORCH_CHECK(c2w_h.dim() == 3 && c2w_h.size(1) == 3 && c2w_h.size(2) == 4, "c2w: f32 [n,3,4]");
    TORCH_CHECK(intri_h.dim() == 3 && intri_h.size(0) == c2w_h.size(0) && bounds_h.size(0) == c2w_h.size(0),
                "intri: f32 [n,3,3], bounds: f32 [n,2]");
    w2c_ = w2c.detach().to(torch::kFloat32).to(torch::kCUDA).contiguous();
    intri_ = intri_h.to(torch::kCUDA);
    bound_ = bounds_h.to(torch::kCUDA);
    void* handle = nullptr;
    int64_t n_nodes = 0, n_trans = 0;
    GF_CHECK(gf_octree_build(max_level, (float)(int64_t(1) << (bbox_levels - 1)), (float)split_dist_thres,
                             c2w_h.data_ptr<float>(), intri_h.data_ptr<float>(), bounds_h.data_ptr<float>(),
                             c2w_h.size(0), /*seed=*/0u, 32 * 32 * 32, 128, &handle, &n_nodes, &n_trans));
    auto nodes = torch::empty({n_nodes * GF_TREE_NODE_BYTES}, torch::kUInt8);
    auto trans = torch::empty({n_trans * GF_TRANS_INFO_BYTES}, torch::kUInt8);
    GF_CHECK(gf_octree_build_fetch(handle, nodes.data_ptr(), trans.data_ptr()));
    LoadStates({nodes, trans, torch::zeros({n_nodes}, torch::kInt64), torch::tensor(sub_div_milestones_, torch::kInt64)}, 0);
  }
  // the scalar arguments of InitSampler (PersSampler.cpp:899-952) without the cameras
  void Configure(int64_t compact_freq, int64_t max_oct_intersect_per_ray, double global_near, bool scale_by_dis,
                 double sample_l, int64_t mode, double ray_march_fineness, double ray_march_init_fineness,
                 int64_t ray_march_fineness_decay_end_iter) {
    compact_freq_ = compact_freq;
    max_oct_ = max_oct_intersect_per_ray;
    global_near_ = global_near;
    scale_by_dis_ = scale_by_dis;
    sample_l_ = sample_l;
    mode_ = mode;
    fineness_ = ray_march_fineness;
    init_fineness_ = ray_march_init_fineness;
    decay_end_ = (double)ray_march_fineness_decay_end_iter;
  }

  std::vector<Tensor> States() {  // PersSampler.cpp:969-979: the milestones still ahead, next one last
    return {tree_nodes_, pers_trans_, visit_cnt_, torch::tensor(sub_div_milestones_, torch::kInt64).to(torch::kCUDA)};
  }
  int64_t LoadStates(const std::vector<Tensor>& states, int64_t idx) {  // PersSampler.cpp:983-1016
    tree_nodes_ = states[idx++].clone().to(torch::kCUDA).contiguous();
    pers_trans_ = states[idx++].clone().to(torch::kCUDA).contiguous();
    visit_cnt_ = states[idx++].clone().to(torch::kCUDA).contiguous();
    edge_pool_ = Tensor();
    auto ms = states[idx++].to(torch::kCPU).to(torch::kInt64).contiguous();
    sub_div_milestones_.assign(ms.data_ptr<int64_t>(), ms.data_ptr<int64_t>() + ms.numel());
    auto i64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA);
    weight_stats_ = torch::full({n_nodes()}, 1000, i64);  // INIT_NODE_STAT
    alpha_stats_ = torch::full({n_nodes()}, 1000, i64);
    auto so = torch::empty({64}, torch::kUInt8);  // children front to back per ray octant (PersSampler.cpp:137-151)
    GF_CHECK(gf_octree_search_order(so.data_ptr<uint8_t>()));
    search_order_ = so.to(torch::kCUDA);
    return idx;
  }

  // PersSampler::GetSamples (PersSampler_cuda.cu:321-477): the reference's eight dense tensors, same order
  std::vector<Tensor> GetSamples(const Tensor& rays_o_raw, const Tensor& rays_d_raw, const Tensor& /*bounds*/) {
    TORCH_CHECK(rays_o_raw.is_cuda() && rays_d_raw.is_cuda(), "gfnerf_b200: tensor is not on a CUDA device");
    TORCH_CHECK(tree_nodes_.defined(), "PersSampler: no octree loaded (LoadStates)");
    c10::cuda::CUDAGuard guard(rays_o_raw.device());
    auto rays_o = rays_o_raw.contiguous().to(torch::kFloat32);
    auto rays_d = (rays_d_raw / torch::linalg_norm(rays_d_raw, 2, {-1}, true)).contiguous().to(torch::kFloat32);  // :323
    const int64_t R = rays_o.size(0), S = GF_MAX_SAMPLE_PER_RAY;
    auto f32 = rays_o.options();
    auto i64 = f32.dtype(torch::kInt64);
    // :380-389  ones in VALIDATE mode, U(0.5, 1.5) in TRAIN mode, times the fineness
    Tensor noise = mode_ == 1 ? torch::ones({S + R + 10}, f32) : torch::rand({S + R + 10}, f32) + 0.5f;
    noise = noise * (float)fineness_;
    auto world = torch::zeros({R, S, 3}, f32), warp = torch::zeros({R, S, 3}, f32), dirs = torch::zeros({R, S, 3}, f32);
    auto anchors = torch::zeros({R, S, 3}, i64);
    auto dists = torch::zeros({R, S}, f32), ts = torch::zeros({R, S}, f32);
    auto start_end = torch::zeros({R, 2}, i64);
    auto first = torch::zeros({R, 1}, f32);
    auto counts = torch::zeros({R}, f32.dtype(torch::kInt32));
    if (R > 0) {
      gf_sampler_out out = {};
      out.world_pts = world.data_ptr<float>();
      out.warp_pts = warp.data_ptr<float>();
      out.dirs = dirs.data_ptr<float>();
      out.dists = dists.data_ptr<float>();
      out.ts = ts.data_ptr<float>();
      out.anchors_i64 = anchors.data_ptr<int64_t>();
      out.pts_idx_start_end = start_end.data_ptr<int64_t>();
      out.counts = counts.data_ptr<int32_t>();
      out.first_oct_dis = first.data_ptr<float>();
      GF_CHECK(gf_sampler_get_samples(R, rays_o.data_ptr<float>(), rays_d.data_ptr<float>(), noise.data_ptr<float>(),
                                      tree_nodes_.data_ptr(), n_nodes(), pers_trans_.data_ptr(), n_trans(),
                                      search_order_.data_ptr<uint8_t>(), (float)global_near_, (float)sample_l_,
                                      scale_by_dis_ ? 1 : 0, max_oct_, &out, stream()));
    }
    return {world, warp, dirs, dists, ts, anchors, start_end, first};
  }

  // the per-step part of PersSampler::UpdateOctNodes (PersSampler_cuda.cu:584-655): votes, stat update, pruning.
  // ProcOctree at the milestones / every compact_freq steps is host work (gfnerf_b200.persoctree.PersOctree.proc_octree).
  void UpdateOctNodes(const Tensor& sampled_anchors, const Tensor& pts_idx_bounds, const Tensor& sampled_weight,
                      const Tensor& sampled_alpha, int64_t iter_step) {
    c10::cuda::CUDAGuard guard(sampled_weight.device());
    const int64_t R = sampled_weight.size(0), S = GF_MAX_SAMPLE_PER_RAY;
    auto se = pts_idx_bounds.select(1, 0);  // [R,2]
    auto counts = (se.select(1, 1) - se.select(1, 0)).to(torch::kInt32).contiguous();
    auto offsets = (torch::arange(R + 1, sampled_weight.options().dtype(torch::kInt64)) * S).to(torch::kInt32);
    auto node = sampled_anchors.reshape({R * S, 3}).select(1, 1).to(torch::kInt32).contiguous();
    auto w = sampled_weight.reshape({-1}).contiguous().to(torch::kFloat32);
    auto a = sampled_alpha.reshape({-1}).contiguous().to(torch::kFloat32);
    auto scratch = torch::empty({3 * n_nodes()}, visit_cnt_.options());
    GF_CHECK(gf_sampler_update_oct_nodes(R, counts.data_ptr<int32_t>(), offsets.data_ptr<int32_t>(),
                                         node.data_ptr<int32_t>(), w.data_ptr<float>(), a.data_ptr<float>(),
                                         tree_nodes_.data_ptr(), n_nodes(), weight_stats_.data_ptr<int64_t>(),
                                         alpha_stats_.data_ptr<int64_t>(), visit_cnt_.data_ptr<int64_t>(),
                                         scratch.data_ptr<int64_t>(), stream()));
    // :657-677 subdivision at the milestones, compaction every compact_freq steps
    while (!sub_div_milestones_.empty() && sub_div_milestones_.back() <= iter_step) {
      ProcOctree(true, true, sub_div_milestones_.back() <= 0);
      MarkInvisibleNodes();
      ProcOctree(true, false, false);
      sub_div_milestones_.pop_back();
    }
    if (compact_freq_ > 0 && iter_step % compact_freq_ == 0) ProcOctree(true, false, false);
  }

  // PersOctree::ProcOctree (PersSampler.cpp:154-417) on the device (gf_octree_proc_device): the node blob and the
  // statistics are rebuilt in HBM; the host reads back the new node count and the error word only
  void ProcOctree(bool compact, bool subdivide, bool brute_force) {
    const int64_t n_in = n_nodes(), cap = subdivide ? 9 * n_in : n_in;
    auto u8 = tree_nodes_.options();
    auto i64 = weight_stats_.options();
    const int64_t scratch_bytes = gf_octree_proc_device_scratch_bytes(n_in);
    auto scratch = torch::empty({scratch_bytes}, u8);
    auto nodes_o = torch::empty({cap * GF_TREE_NODE_BYTES}, u8);
    auto w_o = torch::empty({cap}, i64), a_o = torch::empty({cap}, i64);
    auto res = torch::zeros({2}, i64);   // [0] = nodes written, low word of [1] = error word
    GF_CHECK(gf_octree_proc_device(tree_nodes_.data_ptr(), n_in, weight_stats_.data_ptr<int64_t>(),
                                   alpha_stats_.data_ptr<int64_t>(), visit_cnt_.data_ptr<int64_t>(), compact, subdivide,
                                   brute_force, nodes_o.data_ptr(), w_o.data_ptr<int64_t>(), a_o.data_ptr<int64_t>(), cap,
                                   scratch.data_ptr(), scratch_bytes, res.data_ptr<int64_t>(),
                                   reinterpret_cast<int32_t*>(res.data_ptr<int64_t>() + 1), stream()));
    auto host = res.cpu();
    const int64_t n_out = host.data_ptr<int64_t>()[0], err = host.data_ptr<int64_t>()[1] & 0xffffffffLL;
    TORCH_CHECK(!(err & 1), "gf_octree_proc: the root was pruned (no valid leaf left in the octree)");
    TORCH_CHECK(!(err & 2), "gf_octree_proc: a removed node is still linked (compact = 0 on a pruned tree?)");
    TORCH_CHECK(err == 0, "gf_octree_proc_device: error word ", err);
    tree_nodes_ = nodes_o.slice(0, 0, n_out * GF_TREE_NODE_BYTES).clone();
    weight_stats_ = w_o.slice(0, 0, n_out).clone();
    alpha_stats_ = a_o.slice(0, 0, n_out).clone();
    visit_cnt_ = torch::zeros({n_out}, visit_cnt_.options());
  }

  // PersOctree::MarkInvisibleNodes (MarkInvisibleNodesKernel + CheckVisible, PersSampler_cuda.cu:680-742): a node no
  // camera can see loses its transform -- one kernel over the device node blob (csrc/octree_device.cu)
  void MarkInvisibleNodes() {
    if (!w2c_.defined()) return;  // state loaded without cameras
    c10::cuda::CUDAGuard guard(tree_nodes_.device());
    GF_CHECK(gf_octree_mark_invisible(tree_nodes_.data_ptr(), n_nodes(), w2c_.data_ptr<float>(),
                                      intri_.data_ptr<float>(), bound_.data_ptr<float>(), w2c_.size(0), stream()));
  }

  void UpdateRayMarch(int64_t cur_step) {  // PersSampler.cpp:958-967 (fp32 arithmetic)
    if ((double)cur_step >= decay_end_) {
      fineness_ = 1.0;
    } else {
      const float progress = (float)cur_step / (float)decay_end_;
      fineness_ = std::exp(std::log(1.f) * progress + std::log((float)init_fineness_) * (1.f - progress));
    }
  }
  void UpdateMode(int64_t mode) { mode_ = mode; }

  Tensor TransQueryFrame(const Tensor& world_positions, const Tensor& anchors) {  // :854-922
    c10::cuda::CUDAGuard guard(world_positions.device());
    auto wp = world_positions.contiguous().to(torch::kFloat32);
    auto an = anchors.contiguous().to(torch::kInt64);
    auto out = torch::zeros_like(wp);
    GF_CHECK(gf_sampler_trans_query_frame(wp.size(0), tree_nodes_.data_ptr(), n_nodes(), pers_trans_.data_ptr(),
                                          an.data_ptr<int64_t>(), wp.data_ptr<float>(), out.data_ptr<float>(), stream()));
    return out;
  }

  // PersSampler::GetPointsAnchors (:924-980)
  Tensor GetPointsAnchors(const Tensor& rays_origins, const Tensor& rays_dirs, const Tensor& t_starts,
                          const Tensor& t_ends) {
    c10::cuda::CUDAGuard guard(rays_origins.device());
    const int64_t R = t_starts.size(0), S = t_starts.size(1);
    auto t_cur = ((t_starts.to(torch::kFloat32) + t_ends.to(torch::kFloat32)) / 2.0).reshape({R, S}).contiguous();
    auto o = rays_origins.contiguous().to(torch::kFloat32), d = rays_dirs.contiguous().to(torch::kFloat32);
    auto anchors = torch::empty({R, S, 1}, o.options().dtype(torch::kInt64));
    GF_CHECK(gf_sampler_points_anchors(R, S, o.data_ptr<float>(), d.data_ptr<float>(), t_cur.data_ptr<float>(),
                                       tree_nodes_.data_ptr(), n_nodes(), anchors.data_ptr<int64_t>(), stream()));
    return anchors;
  }

  // PersSampler::GetEdgeSamples (:479-516).  The edge pool (PersSampler.cpp:833-893) is built at the first use after
  // InitSampler / LoadStates and kept while the octree is pruned and subdivided, as the reference keeps the one its
  // constructor built (the records name transforms, which are never removed).
  std::tuple<Tensor, Tensor> GetEdgeSamples(int64_t n_pts) {
    c10::cuda::CUDAGuard guard(tree_nodes_.device());
    if (!edge_pool_.defined()) {
      auto nodes = tree_nodes_.cpu().contiguous();
      int64_t n = 0;
      GF_CHECK(gf_octree_edge_pool(nodes.data_ptr(), n_nodes(), nullptr, 0, &n));
      auto pool = torch::empty({std::max<int64_t>(n, 1) * 64}, torch::kUInt8);
      GF_CHECK(gf_octree_edge_pool(nodes.data_ptr(), n_nodes(), pool.data_ptr(), n, &n));
      edge_pool_ = pool.slice(0, 0, n * 64).to(torch::kCUDA);
    }
    const int64_t n_edges = edge_pool_.numel() / 64;
    TORCH_CHECK(n_edges > 0, "GetEdgeSamples: the octree has no pair of neighbouring valid leaves");
    auto i64 = torch::TensorOptions().dtype(torch::kInt64).device(torch::kCUDA);
    auto edge_idx = torch::randint(0, n_edges, {n_pts}, i64).contiguous();                                  // :498
    auto coords = (torch::rand({n_pts, 2}, i64.dtype(torch::kFloat32)) * 2.f - 1.f).contiguous();            // :499
    auto out_pts = torch::empty({n_pts, 2, 3}, coords.options());
    auto out_idx = torch::empty({n_pts, 2}, i64);
    GF_CHECK(gf_sampler_edge_samples(n_pts, edge_pool_.data_ptr(), n_edges, pers_trans_.data_ptr(),
                                     edge_idx.data_ptr<int64_t>(), coords.data_ptr<float>(), out_pts.data_ptr<float>(),
                                     out_idx.data_ptr<int64_t>(), stream()));
    return std::make_tuple(out_pts, out_idx);
  }

  // QueryTreeNodeCenterKernel (:984-1028): centre of the node each anchor names (zeros for an anchor out of range)
  Tensor QueryTreeNodeCenters(const Tensor& anchors) {
    c10::cuda::CUDAGuard guard(tree_nodes_.device());
    auto idx = anchors.reshape({-1}).to(torch::kInt64).to(tree_nodes_.device());
    auto center = tree_nodes_.view({-1, GF_TREE_NODE_BYTES}).slice(1, 0, 12).contiguous().view(torch::kFloat32);
    auto ok = (idx >= 0) & (idx < n_nodes());
    return center.index_select(0, idx.clamp(0, n_nodes() - 1)) * ok.unsqueeze(1).to(torch::kFloat32);
  }

  // PersOctree::UpdateBlockIdxs + SetBlockIdxsNearestKernel (PersSampler_cuda.cu:746-798): nearest block centre per
  // node (fp32 norm, the first of equal minima), then compact
  void UpdateBlockIdxs(const Tensor& centers) {
    c10::cuda::CUDAGuard guard(tree_nodes_.device());
    auto c = centers.to(tree_nodes_.device()).to(torch::kFloat32).reshape({-1, 3}).contiguous();
    GF_CHECK(gf_octree_set_block_idxs(tree_nodes_.data_ptr(), n_nodes(), c.data_ptr<float>(), c.size(0), stream()));
    ProcOctree(true, false, false);
  }

  // PersSampler::VisOctree (PersSampler.cpp:478-514): the leaves' boxes as an .obj wireframe
  void VisOctree(std::string base_exp_dir) {
    auto nodes = tree_nodes_.cpu().contiguous();
    const uint8_t* p = nodes.data_ptr<uint8_t>();
    const int64_t n = n_nodes();
    std::ofstream f(base_exp_dir + "/octree.obj", std::ios::out);
    TORCH_CHECK(f.good(), "VisOctree: cannot write ", base_exp_dir, "/octree.obj");
    for (int64_t i = 0; i < n; i++) {
      const float* cs = reinterpret_cast<const float*>(p + i * GF_TREE_NODE_BYTES);
      for (int st = 0; st < 8; st++)
        f << "v " << cs[0] + (((st >> 2) & 1) - 0.5f) * cs[3] << " " << cs[1] + (((st >> 1) & 1) - 0.5f) * cs[3] << " "
          << cs[2] + ((st & 1) - 0.5f) * cs[3] << std::endl;
    }
    for (int64_t i = 0; i < n; i++) {
      if (!p[i * GF_TREE_NODE_BYTES + 88]) continue;  // is_leaf_node
      for (int a = 0; a < 8; a++)
        for (int b = a + 1; b < 8; b++) {
          const int st = a ^ b;
          if (st == 1 || st == 2 || st == 4) f << "l " << i * 8 + a + 1 << " " << i * 8 + b + 1 << std::endl;
        }
    }
  }

  // introspection (bindings.cpp:86-299): nested lists, as the reference returns them
  std::vector<int64_t> get_sub_div_milestones_() { return sub_div_milestones_; }
  template <typename T, typename Out>
  std::vector<Out> node_field(int64_t offset) {
    auto nodes = tree_nodes_.cpu().contiguous();
    const uint8_t* p = nodes.data_ptr<uint8_t>();
    std::vector<Out> out((size_t)n_nodes());
    for (int64_t i = 0; i < n_nodes(); i++) {
      T v;
      memcpy(&v, p + i * GF_TREE_NODE_BYTES + offset, sizeof(T));
      out[(size_t)i] = (Out)v;
    }
    return out;
  }
  std::vector<std::vector<double>> get_tree_nodes_center_() {
    auto x = node_field<float, double>(0), y = node_field<float, double>(4), z = node_field<float, double>(8);
    std::vector<std::vector<double>> out(x.size());
    for (size_t i = 0; i < x.size(); i++) out[i] = {x[i], y[i], z[i]};
    return out;
  }
  std::vector<double> get_tree_nodes_side_len_() { return node_field<float, double>(12); }
  std::vector<bool> get_tree_nodes_is_leaf_node_() {
    auto v = node_field<uint8_t, int64_t>(88);
    return std::vector<bool>(v.begin(), v.end());
  }
  std::vector<int64_t> get_tree_nodes_trans_idx_() { return node_field<int64_t, int64_t>(96); }
  std::vector<int64_t> get_tree_nodes_block_idx_() { return node_field<int64_t, int64_t>(104); }
  // (w2xz [n][12][2][4], weight [n][3][12], center [n][3], side_len [n], dis_summary [n])
  std::tuple<std::vector<std::vector<std::vector<std::vector<double>>>>, std::vector<std::vector<std::vector<double>>>,
             std::vector<std::vector<double>>, std::vector<double>, std::vector<double>>
  get_pers_trans_info() {
    auto tr = pers_trans_.cpu().contiguous();
    const uint8_t* p = tr.data_ptr<uint8_t>();
    const int64_t n = n_trans();
    std::vector<std::vector<std::vector<std::vector<double>>>> w2xz((size_t)n);
    std::vector<std::vector<std::vector<double>>> weight((size_t)n);
    std::vector<std::vector<double>> center((size_t)n);
    std::vector<double> side((size_t)n), dis((size_t)n);
    for (int64_t i = 0; i < n; i++) {
      const float* f = reinterpret_cast<const float*>(p + i * GF_TRANS_INFO_BYTES);
      w2xz[(size_t)i].assign(GF_N_PROS, std::vector<std::vector<double>>(2, std::vector<double>(4)));
      for (int k = 0; k < GF_N_PROS; k++)
        for (int a = 0; a < 2; a++)
          for (int b = 0; b < 4; b++) w2xz[(size_t)i][k][a][b] = f[(k * 2 + a) * 4 + b];
      weight[(size_t)i].assign(3, std::vector<double>(GF_N_PROS));
      for (int r = 0; r < 3; r++)
        for (int k = 0; k < GF_N_PROS; k++) weight[(size_t)i][r][k] = f[96 + r * GF_N_PROS + k];
      center[(size_t)i] = {f[132], f[133], f[134]};
      side[(size_t)i] = f[135];
      dis[(size_t)i] = f[136];
    }
    return std::make_tuple(w2xz, weight, center, side, dis);
  }

  int64_t get_compact_freq_() { return compact_freq_; }
  int64_t get_max_oct_intersect_per_ray_() { return max_oct_; }
  double get_global_near_() { return global_near_; }
  double get_sample_l_() { return sample_l_; }
  double get_scale_by_dis_() { return scale_by_dis_ ? 1.0 : 0.0; }   // double_t in the reference (bindings.cpp:272-275)
  int64_t get_mode_() { return mode_; }
  int64_t get_n_volumes_() { return n_trans(); }
  double get_sampled_oct_per_ray_() { return sampled_oct_per_ray_; }
  double get_ray_march_fineness_() { return fineness_; }
};

}  // namespace

TORCH_LIBRARY(my_classes, m) {  // same names as gfnerf/bindings/hashanchored/bindings.cpp:343-401
  m.class_<Hash3DAnchoredImpl>("Hash3DAnchored")
      .def(torch::init<int64_t, int64_t, double>())
      .def("AnchoredQuery", &Hash3DAnchoredImpl::AnchoredQuery)
      .def("LoadStates", &Hash3DAnchoredImpl::LoadStates)
      .def("States", &Hash3DAnchoredImpl::States)
      .def("Reset", &Hash3DAnchoredImpl::Reset)
      .def("Zero", &Hash3DAnchoredImpl::Zero)
      .def("GetParams", &Hash3DAnchoredImpl::GetParams)
      .def("SetFeatPoolRequireGrad", &Hash3DAnchoredImpl::SetFeatPoolRequireGrad)
      .def("ReleaseResources", &Hash3DAnchoredImpl::ReleaseResources)
      .def("to", &Hash3DAnchoredImpl::to);

  m.class_<PersSamplerImpl>("PersSampler")
      .def(torch::init<>())
      .def("InitSampler", &PersSamplerImpl::InitSampler)
      .def("Configure", &PersSamplerImpl::Configure)
      .def("GetSamples", &PersSamplerImpl::GetSamples)
      .def("UpdateOctNodes", &PersSamplerImpl::UpdateOctNodes)
      .def("ProcOctree", &PersSamplerImpl::ProcOctree)
      .def("MarkInvisibleNodes", &PersSamplerImpl::MarkInvisibleNodes)
      .def("UpdateRayMarch", &PersSamplerImpl::UpdateRayMarch)
      .def("UpdateMode", &PersSamplerImpl::UpdateMode)
      .def("States", &PersSamplerImpl::States)
      .def("LoadStates", &PersSamplerImpl::LoadStates)
      .def("get_compact_freq_", &PersSamplerImpl::get_compact_freq_)
      .def("get_max_oct_intersect_per_ray_", &PersSamplerImpl::get_max_oct_intersect_per_ray_)
      .def("get_global_near_", &PersSamplerImpl::get_global_near_)
      .def("get_sample_l_", &PersSamplerImpl::get_sample_l_)
      .def("get_scale_by_dis_", &PersSamplerImpl::get_scale_by_dis_)
      .def("get_mode_", &PersSamplerImpl::get_mode_)
      .def("get_n_volumes_", &PersSamplerImpl::get_n_volumes_)
      .def("get_sampled_oct_per_ray_", &PersSamplerImpl::get_sampled_oct_per_ray_)
      .def("get_ray_march_fineness_", &PersSamplerImpl::get_ray_march_fineness_)
      .def("trans_query_frame", &PersSamplerImpl::TransQueryFrame)
      .def("qurey_tree_nodes_centers", &PersSamplerImpl::QueryTreeNodeCenters)  // (sic) bindings.cpp:376
      .def("get_points_anchors", &PersSamplerImpl::GetPointsAnchors)
      .def("GetEdgeSamples", &PersSamplerImpl::GetEdgeSamples)
      .def("UpdateBlockIdxs", &PersSamplerImpl::UpdateBlockIdxs)
      .def("VisOctree", &PersSamplerImpl::VisOctree)
      .def("get_sub_div_milestones_", &PersSamplerImpl::get_sub_div_milestones_)
      .def("get_tree_nodes_center_", &PersSamplerImpl::get_tree_nodes_center_)
      .def("get_tree_nodes_side_len_", &PersSamplerImpl::get_tree_nodes_side_len_)
      .def("get_tree_nodes_is_leaf_node_", &PersSamplerImpl::get_tree_nodes_is_leaf_node_)
      .def("get_tree_nodes_trans_idx_", &PersSamplerImpl::get_tree_nodes_trans_idx_)
      .def("get_tree_nodes_block_idx_", &PersSamplerImpl::get_tree_nodes_block_idx_)
      .def("get_pers_trans_info", &PersSamplerImpl::get_pers_trans_info);
}
