// Data-parallel gradient exchange over NVLink peer memory, fused with the optimizer (sm_100a, one process per GPU).
//
// What the reference does at this point of a step: nerfstudio wraps the model in DDP (gfnerf/gf_pipeline.py:136-138),
// which all-reduces the gradients of the registered parameters and then every rank runs the same torch.optim.Adam
// step on its replica (nerfstudio/engine/optimizers.py:125-137) -- with the hash table silently left out, because
// feat_pool is not an nn.Parameter.  Round 1 did the straightforward thing for ALL parameters:
// ncclAllReduce of the fp32 table gradient (35.6 MB at log2T = 19) + Adam over the whole table on every rank.  That
// one collective was 0.45 ms of exposed time per 5 ms step at 2 / 4 / 8 GPUs (scaling efficiency 0.91).
//
// Here the table is exchanged ZeRO-style by ONE kernel per rank, gf_peer_reduce_adam, over peer mappings of every
// rank's buffers (cudaIpc*, NVSwitch: every GPU reaches every peer at full NVLink bandwidth):
//
//   rank r owns rows [r, r + 1) * n / world of the table.  For its rows it
//     1. reads the gradient rows of ALL ranks (world - 1 peer loads + 1 local, 16 bytes each, in a fixed rank order,
//        so the sum -- and with it every replica -- is bit-identical and independent of timing),
//     2. applies Adam to its fp32 master rows / moments (which only the owner keeps up to date),
//     3. stores the updated rows as fp16 into the gather table ("shadow") of EVERY rank (world - 1 peer stores).
//   = reduce-scatter + optimizer + all-gather in one pass: (world-1)/world * 36 MB in over NVLink, 18 MB/world *
//   (world-1) out, nothing staged, no fp32 all-gather at all (the forward only ever reads the fp16 shadow).
//   The small parameters (MLP, appearance embedding: ~0.1 MB) go through the same kernel with every rank owning
//   everything (each reads all ranks' gradients and updates its own replica -- same order, same result).
//
// Cross-GPU ordering is two tiny single-CTA kernels (gf_peer_barrier) on the same stream: before the exchange
// ("every rank's gradients are complete", which also ORs the ranks' NaN-gradient flags: the trainer's guard,
// trainer.py:416-426, must take the same decision everywhere) and after it ("every rank has finished reading my
// gradients and writing my shadow").  They are separate launches on purpose: a multi-CTA kernel that spins on a
// remote flag deadlocks as soon as some of its CTAs are not resident (another stream's kernel on the SMs), a
// one-CTA kernel cannot.  Flags live in peer-mapped memory and carry a monotonically increasing epoch, so they never
// need resetting; a wait that exceeds ~2 s raises *d_error instead of hanging the GPU.
#include "common.cuh"

namespace gf {

constexpr int kMaxPeers = 8;

struct PeerPtrs {
  void* p[kMaxPeers];
};

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;\n" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];\n" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long global_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;\n" : "=l"(t));
  return t;
}

// lane p < world: publish (epoch, local flag) in rank p's flag array, slot `rank`; then wait until every rank has
// published this epoch in mine.  d_any_flag = OR of the ranks' flags.
__global__ void peer_barrier_kernel(int world, int rank, uint32_t epoch, PeerPtrs flags, const int32_t* d_local_flag,
                                    int32_t* d_any_flag, int32_t* d_error) {
  const int p = threadIdx.x;
  const uint32_t mine = (d_local_flag && *d_local_flag) ? 1u : 0u;
  uint32_t got = 0u;
  bool ok = true;
  if (p < world) {
    __threadfence_system();   // everything this GPU wrote before (earlier kernels of the stream) is visible system-wide
    st_release_sys(reinterpret_cast<uint32_t*>(flags.p[p]) + rank, 2u * epoch + mine);
    const uint32_t* slot = reinterpret_cast<const uint32_t*>(flags.p[rank]) + p;
    const unsigned long long t0 = global_ns();
    while (true) {
      got = ld_acquire_sys(slot);
      if ((got >> 1) >= epoch) break;   // (a peer may already be an epoch ahead of a slow reader: never behind)
      if (global_ns() - t0 > 2000000000ull) {
        ok = false;
        break;
      }
      __nanosleep(64);
    }
  }
  const unsigned bad = __ballot_sync(0xffffffffu, !ok);
  // a peer that is already in the next epoch cannot tell us its flag of this one; flags only matter at the barrier
  // in front of the exchange, where no rank can be ahead (it would have needed OUR signal of this epoch to leave)
  const unsigned any = __ballot_sync(0xffffffffu, p < world && ok && (got >> 1) == epoch && (got & 1u));
  if (p == 0) {
    if (d_any_flag) *d_any_flag = (any != 0u) ? 1 : 0;
    if (bad && d_error) atomicOr(d_error, 1);
  }
}

// rows [lo4, hi4) (in float4 units) of a flat parameter array: sum the gradient of all ranks, Adam, store fp16 to all
__host__ __device__ inline void peer_bias_corrections(long long t, float lr, float beta1, float beta2, float* step_size,
                                                      float* inv_bc2_sqrt) {
  const double bc1 = 1.0 - pow((double)beta1, (double)t);
  const double bc2 = 1.0 - pow((double)beta2, (double)t);
  *step_size = (float)((double)lr / bc1);
  *inv_bc2_sqrt = (float)(1.0 / sqrt(bc2));
}

template <int WORLD, bool SHADOW>
__global__ void __launch_bounds__(256)
peer_reduce_adam_kernel(int64_t lo4, int64_t hi4, PeerPtrs grads, float4* __restrict__ param, float4* __restrict__ m,
                        float4* __restrict__ v, PeerPtrs shadows, float lr, float beta1, float beta2, float eps,
                        const long long* __restrict__ d_step, float inv_div, const int* __restrict__ skip_flag) {
  if (skip_flag && *skip_flag) return;   // NaN somewhere: no rank updates anything (the gradients are zeroed by the caller)
  float step_size, inv_bc2_sqrt;
  peer_bias_corrections(*d_step + 1, lr, beta1, beta2, &step_size, &inv_bc2_sqrt);
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = lo4 + (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < hi4; i += stride) {
    float4 g[WORLD];
#pragma unroll
    for (int r = 0; r < WORLD; r++) g[r] = __ldcg(reinterpret_cast<const float4*>(grads.p[r]) + i);   // all in flight
    float4 s = g[0];
#pragma unroll
    for (int r = 1; r < WORLD; r++) {   // fixed order: the same sum on whichever rank computes it
      s.x += g[r].x;
      s.y += g[r].y;
      s.z += g[r].z;
      s.w += g[r].w;
    }
    float4 p = param[i], mm = m[i], vv = v[i];
    float* pp = &p.x; float* gp = &s.x; float* mp = &mm.x; float* vp = &vv.x;
#pragma unroll
    for (int k = 0; k < 4; k++) {        // torch.optim.Adam, the arithmetic of adam_kernel (optim.cu)
      const float gk = gp[k] * inv_div;
      mp[k] = mp[k] + (gk - mp[k]) * (1.f - beta1);
      vp[k] = vp[k] * beta2 + (1.f - beta2) * gk * gk;
      const float denom = sqrtf(vp[k]) * inv_bc2_sqrt + eps;
      pp[k] = pp[k] - step_size * (mp[k] / denom);
    }
    param[i] = p; m[i] = mm; v[i] = vv;
    if (SHADOW) {
      __half2 l = __floats2half2_rn(p.x, p.y), h = __floats2half2_rn(p.z, p.w);
      uint2 o;
      o.x = *reinterpret_cast<uint32_t*>(&l);
      o.y = *reinterpret_cast<uint32_t*>(&h);
#pragma unroll
      for (int r = 0; r < WORLD; r++) reinterpret_cast<uint2*>(shadows.p[r])[i] = o;
    }
  }
}

// dst[i] = max over the ranks of src_r[i]: the octree votes (MarkVistNodeKernel's atomicMax adders / marks and the
// visit counts, PersSampler_cuda.cu:518-574) of ALL rays of the step, so that every replica applies the votes one
// process seeing every ray would have produced -- an all-reduce(MAX) as (world - 1) peer loads per element
template <int WORLD>
__global__ void __launch_bounds__(256) peer_max_i64_kernel(int64_t n, PeerPtrs src, long long* __restrict__ dst) {
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    long long v[WORLD];
#pragma unroll
    for (int r = 0; r < WORLD; r++) v[r] = __ldcg(reinterpret_cast<const long long*>(src.p[r]) + i);
    long long m = v[0];
#pragma unroll
    for (int r = 1; r < WORLD; r++) m = v[r] > m ? v[r] : m;
    dst[i] = m;
  }
}

__global__ void peer_count_kernel(long long* d_step, const int* skip_flag) {
  if (!(skip_flag && *skip_flag)) *d_step += 1;
}

template <bool SHADOW>
static int launch_reduce_adam(int world, int grid, cudaStream_t st, int64_t lo4, int64_t hi4, const PeerPtrs& g, float* param,
                              float* m, float* v, const PeerPtrs& sh, float lr, float b1, float b2, float eps,
                              const long long* d_step, float inv_div, const int* skip) {
#define GF_PEER_CASE(W)                                                                                              \
  case W:                                                                                                            \
    peer_reduce_adam_kernel<W, SHADOW><<<grid, 256, 0, st>>>(lo4, hi4, g, (float4*)param, (float4*)m, (float4*)v, sh, \
                                                             lr, b1, b2, eps, d_step, inv_div, skip);                 \
    break;
  switch (world) {
    GF_PEER_CASE(1) GF_PEER_CASE(2) GF_PEER_CASE(3) GF_PEER_CASE(4) GF_PEER_CASE(5) GF_PEER_CASE(6) GF_PEER_CASE(7)
    GF_PEER_CASE(8)
    default:
      set_error("gf_peer_reduce_adam: world size %d not built (1..8)", world);
      return GF_ERR_INVALID;
  }
#undef GF_PEER_CASE
  return check_launch("peer_reduce_adam_kernel");
}

}  // namespace gf

using namespace gf;

extern "C" {

int gf_peer_alloc(int64_t bytes, void** dptr) {
  GF_REQUIRE(bytes > 0 && dptr, "gf_peer_alloc: bad arguments");
  GF_CUDA(cudaMalloc(dptr, (size_t)bytes));
  GF_CUDA(cudaMemset(*dptr, 0, (size_t)bytes));
  return GF_OK;
}

int gf_peer_free(void* dptr) {
  if (dptr) GF_CUDA(cudaFree(dptr));
  return GF_OK;
}

int gf_peer_export(const void* dptr, void* handle64) {
  GF_REQUIRE(dptr && handle64, "gf_peer_export: null pointer");
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t is 64 bytes");
  GF_CUDA(cudaIpcGetMemHandle(reinterpret_cast<cudaIpcMemHandle_t*>(handle64), const_cast<void*>(dptr)));
  return GF_OK;
}

int gf_peer_import(const void* handle64, void** dptr) {
  GF_REQUIRE(handle64 && dptr, "gf_peer_import: null pointer");
  cudaIpcMemHandle_t h;
  memcpy(&h, handle64, sizeof(h));
  GF_CUDA(cudaIpcOpenMemHandle(dptr, h, cudaIpcMemLazyEnablePeerAccess));
  return GF_OK;
}

int gf_peer_close(void* dptr) {
  if (dptr) GF_CUDA(cudaIpcCloseMemHandle(dptr));
  return GF_OK;
}

int gf_peer_barrier(int world, int rank, uint32_t epoch, void* const* flag_ptrs, const int32_t* d_local_flag,
                    int32_t* d_any_flag, int32_t* d_error, void* stream) {
  GF_REQUIRE(world >= 1 && world <= kMaxPeers && rank >= 0 && rank < world && flag_ptrs && epoch > 0 &&
                 epoch < 0x7fffffffu,
             "gf_peer_barrier: bad arguments");
  PeerPtrs f{};
  for (int r = 0; r < world; r++) {
    GF_REQUIRE(flag_ptrs[r] != nullptr, "gf_peer_barrier: null flag array of rank %d", r);
    f.p[r] = flag_ptrs[r];
  }
  peer_barrier_kernel<<<1, 32, 0, (cudaStream_t)stream>>>(world, rank, epoch, f, d_local_flag, d_any_flag, d_error);
  return check_launch("peer_barrier_kernel");
}

int gf_peer_max_i64(int world, int64_t n, void* const* src_ptrs, int64_t* dst, void* stream) {
  GF_REQUIRE(world >= 1 && world <= kMaxPeers && n >= 0 && src_ptrs, "gf_peer_max_i64: bad arguments");
  if (n == 0) return GF_OK;
  GF_REQUIRE(dst != nullptr, "gf_peer_max_i64: null output");
  PeerPtrs p{};
  for (int r = 0; r < world; r++) {
    GF_REQUIRE(src_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(src_ptrs[r]) & 7) == 0,
               "gf_peer_max_i64: source pointer of rank %d is null or not 8-byte aligned", r);
    p.p[r] = src_ptrs[r];
  }
  const int grid = stride_grid(n, 256, 4, 1);
  cudaStream_t st = (cudaStream_t)stream;
#define GF_PEER_CASE(W) \
  case W:               \
    peer_max_i64_kernel<W><<<grid, 256, 0, st>>>(n, p, (long long*)dst); \
    break;
  switch (world) {
    GF_PEER_CASE(1) GF_PEER_CASE(2) GF_PEER_CASE(3) GF_PEER_CASE(4) GF_PEER_CASE(5) GF_PEER_CASE(6) GF_PEER_CASE(7)
    GF_PEER_CASE(8)
    default:
      set_error("gf_peer_max_i64: world size %d not built (1..8)", world);
      return GF_ERR_INVALID;
  }
#undef GF_PEER_CASE
  return check_launch("peer_max_i64_kernel");
}

int gf_peer_reduce_adam(int world, int64_t n, int64_t lo, int64_t hi, void* const* grad_ptrs, float* param,
                        float* exp_avg, float* exp_avg_sq, void* const* shadow_ptrs, float lr, float beta1, float beta2,
                        float eps, int64_t* d_step, float grad_div, const int32_t* skip_flag, void* stream) {
  GF_REQUIRE(world >= 1 && world <= kMaxPeers && n >= 0 && lo >= 0 && lo <= hi && hi <= n && grad_div != 0.f,
             "gf_peer_reduce_adam: bad arguments");
  GF_REQUIRE(lo % 4 == 0 && (hi % 4 == 0), "gf_peer_reduce_adam: the owned range must be a multiple of 4 elements");
  GF_REQUIRE(grad_ptrs && param && exp_avg && exp_avg_sq && d_step, "gf_peer_reduce_adam: null pointer");
  PeerPtrs g{}, sh{};
  for (int r = 0; r < world; r++) {
    GF_REQUIRE(grad_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(grad_ptrs[r]) & 15) == 0,
               "gf_peer_reduce_adam: gradient pointer of rank %d is null or not 16-byte aligned", r);
    g.p[r] = grad_ptrs[r];
    if (shadow_ptrs) {
      GF_REQUIRE(shadow_ptrs[r] != nullptr && (reinterpret_cast<uintptr_t>(shadow_ptrs[r]) & 7) == 0,
                 "gf_peer_reduce_adam: shadow pointer of rank %d is null or not 8-byte aligned", r);
      sh.p[r] = shadow_ptrs[r];
    }
  }
  cudaStream_t st = (cudaStream_t)stream;
  if (hi > lo) {
    const int64_t lo4 = lo / 4, hi4 = hi / 4;
    const int grid = stride_grid(hi4 - lo4, 256, 8, 1);
    int rc = shadow_ptrs
                 ? launch_reduce_adam<true>(world, grid, st, lo4, hi4, g, param, exp_avg, exp_avg_sq, sh, lr, beta1,
                                            beta2, eps, (const long long*)d_step, 1.f / grad_div, skip_flag)
                 : launch_reduce_adam<false>(world, grid, st, lo4, hi4, g, param, exp_avg, exp_avg_sq, sh, lr, beta1,
                                             beta2, eps, (const long long*)d_step, 1.f / grad_div, skip_flag);
    if (rc) return rc;
  }
  peer_count_kernel<<<1, 1, 0, st>>>((long long*)d_step, skip_flag);
  return check_launch("peer_count_kernel");
}

}  // extern "C"
