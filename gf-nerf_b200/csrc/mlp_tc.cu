// Fused field MLP on the 5th-generation tensor cores (tcgen05 + TMEM), sm_100a.
//
// Same function as the register-chained kernels of mlp.cu (the two MLPNetwork stacks of
// gfnerf/nerfacto_field.py:174-179,217-227 + trunc_exp + sigmoid, SH / embedding folded into a per-ray bias),
// re-mapped to the Blackwell execution model:
//
//  * a CTA of 128 threads owns a tile of 128 samples; THREAD r IS ROW r of every activation matrix and owns
//    TMEM lane r (tcgen05.ld/st shape 32x32b: warp w <-> lanes 32w..32w+31);
//  * every layer is D[128 x N] = A[128 x K] . W^T issued by ONE thread as K/16 tcgen05.mma (M = 128,
//    kind::f16, fp32 accumulate in TMEM);
//  * the A operand never touches shared memory: the epilogue of layer l (tcgen05.ld of the fp32 accumulator row,
//    + bias, ReLU, pack to fp16 pairs) writes the next layer's A straight back into TMEM with tcgen05.st and the
//    next MMA reads it from there (TS form, A K-major in TMEM: two fp16 per 32-bit column);
//  * weights (B operands) are staged once per CTA into shared memory as fp16 in the canonical no-swizzle
//    K-major core-matrix layout and addressed through shared-memory matrix descriptors;
//  * completion: tcgen05.commit -> mbarrier; the 128 threads wait on it with try_wait.parity.
//  * TMEM budget per CTA: accumulator 64 columns + A operand 32 columns -> 128 allocated, so four CTAs
//    (four tiles in flight) share an SM's 512 columns; persistent grid of 4 x 148 CTAs.
//
// Per sample the kernel reads 64 B of features + 4 B ray id (+ the ray's 256-B bias row, L1-resident along a ray)
// and writes 16 B; the 2 x 11 392 FLOP of the five layers run on the tensor pipe.
#include "common.cuh"
#include "tcgen05.cuh"

namespace gf {
namespace tc {

constexpr int kH = 64;
// parameter blob offsets (torch nn.Linear layout), see include/gfnerf_b200.h
constexpr int kW0 = 0, kB0 = kW0 + kH * 32, kW1 = kB0 + kH, kB1 = kW1 + 16 * kH, kW2 = kB1 + 16,
              kB2 = kW2 + kH * 63, kW3 = kB2 + kH, kB3 = kW3 + kH * kH, kW4 = kB3 + kH, kB4 = kW4 + 3 * kH;

constexpr int kTile = 128;     // samples per CTA tile = MMA M
constexpr int kThreads = 128;

// B operands in shared memory: [N rows][K] fp16, K-major, no swizzle: element (n, k) at
//   (n / 8) * SBO + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2,   SBO = (K / 8) * 128
constexpr uint32_t kLbo = 128;
constexpr uint32_t sbo_of(int K) { return (uint32_t)(K / 8) * 128u; }
constexpr uint32_t kOffB0 = 0;                                   // N 64, K 32
constexpr uint32_t kOffB1 = kOffB0 + 64 * 32 * 2;                // N 16, K 64
constexpr uint32_t kOffB2 = kOffB1 + 16 * 64 * 2;                // N 64, K 16 (geo columns of the head's layer 0)
constexpr uint32_t kOffB3 = kOffB2 + 64 * 16 * 2;                // N 64, K 64
constexpr uint32_t kOffB4 = kOffB3 + 64 * 64 * 2;                // N 16 (3 real rows), K 64
constexpr uint32_t kOffBias = kOffB4 + 16 * 64 * 2;              // b0[64] | b1[16] | b3[64] | b4[16]  (fp32)
constexpr uint32_t kOffBar = kOffBias + 160 * 4;
constexpr uint32_t kOffTmem = kOffBar + 8;
constexpr uint32_t kFwdSmem = kOffTmem + 8;

constexpr uint32_t kTmemCols = 128;
constexpr uint32_t kColD = 0;    // accumulator, up to 64 fp32 columns
constexpr uint32_t kColA = 64;   // A operand, up to 32 columns (64 fp16)

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}

__device__ __forceinline__ void put_b(unsigned char* smem, uint32_t off, int K, int n, int k, float v) {
  *reinterpret_cast<__half*>(smem + off + (n >> 3) * sbo_of(K) + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) =
      __float2half_rn(v);
}

// fp32 parameter blob -> fp16 B tiles + fp32 biases
__device__ __forceinline__ void stage_weights(const float* __restrict__ p, unsigned char* smem) {
  for (uint32_t i = threadIdx.x; i < kOffBias / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
  __syncthreads();
  for (int i = threadIdx.x; i < kH * 32; i += blockDim.x) put_b(smem, kOffB0, 32, i >> 5, i & 31, __ldg(p + kW0 + i));
  for (int i = threadIdx.x; i < 16 * kH; i += blockDim.x) put_b(smem, kOffB1, 64, i >> 6, i & 63, __ldg(p + kW1 + i));
  // geo column c (1..15) of the head's first layer multiplies h[c] = in2[15 + c]; column 0 (the density logit) is 0
  for (int i = threadIdx.x; i < kH * 15; i += blockDim.x) {
    const int j = i / 15, c = i % 15;
    put_b(smem, kOffB2, 16, j, 1 + c, __ldg(p + kW2 + j * 63 + 16 + c));
  }
  for (int i = threadIdx.x; i < kH * kH; i += blockDim.x) put_b(smem, kOffB3, 64, i >> 6, i & 63, __ldg(p + kW3 + i));
  for (int i = threadIdx.x; i < 3 * kH; i += blockDim.x) put_b(smem, kOffB4, 64, i >> 6, i & 63, __ldg(p + kW4 + i));
  float* bs = reinterpret_cast<float*>(smem + kOffBias);
  for (int i = threadIdx.x; i < 160; i += blockDim.x) {
    float v;
    if (i < 64) v = __ldg(p + kB0 + i);
    else if (i < 80) v = __ldg(p + kB1 + i - 64);
    else if (i < 144) v = __ldg(p + kB3 + i - 80);
    else v = (i - 144) < 3 ? __ldg(p + kB4 + i - 144) : 0.f;
    bs[i] = v;
  }
}

// one layer: D[kColD .. +N) = A[kColA .. +K/2) . B^T, K/16 MMAs, then commit to the mbarrier
template <int N, int K>
__device__ __forceinline__ void issue_layer(uint32_t tmem, uint32_t b_smem, uint32_t bar) {
  constexpr uint32_t idesc = instr_desc(kTile, N);
#pragma unroll
  for (int k = 0; k < K / 16; k++)
    mma_ts(tmem + kColD, tmem + kColA + 8 * k, smem_desc(b_smem + 2 * k * kLbo, kLbo, sbo_of(K)), idesc, k > 0);
  mma_commit(bar);
}

__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// acc (fp32 bits in v[0..32)) + bias[0..32) -> ReLU -> 16 packed fp16 pairs
__device__ __forceinline__ void bias_relu_pack32(const uint32_t (&v)[32], const float* bias, uint32_t* out16) {
#pragma unroll
  for (int q = 0; q < 8; q++) {
    const float4 b = *reinterpret_cast<const float4*>(bias + 4 * q);
    out16[2 * q] = pack_h2(fmaxf(__uint_as_float(v[4 * q]) + b.x, 0.f), fmaxf(__uint_as_float(v[4 * q + 1]) + b.y, 0.f));
    out16[2 * q + 1] =
        pack_h2(fmaxf(__uint_as_float(v[4 * q + 2]) + b.z, 0.f), fmaxf(__uint_as_float(v[4 * q + 3]) + b.w, 0.f));
  }
}

__global__ void __launch_bounds__(kThreads)
mlp_fwd_tc_kernel(int64_t n, const int32_t* __restrict__ d_n_ptr, const float* __restrict__ params,
                  const __half* __restrict__ feat, const int32_t* __restrict__ ray_id,
                  const float* __restrict__ ray_bias, float* __restrict__ sigma, float* __restrict__ rgb) {
  extern __shared__ __align__(128) unsigned char smem[];
  const int tid = threadIdx.x, warp = tid >> 5;
  stage_weights(params, smem);
  const uint32_t bar = smem_u32(smem + kOffBar);
  if (tid == 0) {
    mbar_init(bar, 1);
    fence_mbar_init();
  }
  if (warp == 0) {
    __syncwarp();
    tmem_alloc(smem_u32(smem + kOffTmem), kTmemCols);
  }
  fence_proxy_async();  // the weight tiles were written through the generic proxy; the MMA reads them through the async proxy
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *reinterpret_cast<volatile uint32_t*>(smem + kOffTmem);
  const uint32_t lane_addr = tmem + ((uint32_t)(warp * 32) << 16);  // this warp's TMEM lane quadrant
  const uint32_t sB = smem_u32(smem);
  const float* bs = reinterpret_cast<const float*>(smem + kOffBias);
  if (d_n_ptr) {
    const int64_t dn = *d_n_ptr;
    n = dn < n ? dn : n;
  }
  uint32_t phase = 0;
  const int64_t n_tiles = (n + kTile - 1) / kTile;
  for (int64_t tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
    const int64_t row = tile * kTile + tid;
    const bool valid = row < n;
    // ---- layer 0 input: this sample's 32 fp16 features -> A[.., 0:16 columns) ------------------------------
    {
      uint32_t x[16];
      const uint4* src = reinterpret_cast<const uint4*>(feat + row * 32);
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const uint4 u = valid ? __ldg(src + q) : make_uint4(0u, 0u, 0u, 0u);
        x[4 * q] = u.x; x[4 * q + 1] = u.y; x[4 * q + 2] = u.z; x[4 * q + 3] = u.w;
      }
      tmem_st16(lane_addr + kColA, x);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_layer<64, 32>(tmem, sB + kOffB0, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 0: h1 = relu(acc + b0) -> A (64 fp16 = 32 columns) -----------------------------------------
    {
      uint32_t v[32], a[32];
      tmem_ld32(lane_addr + kColD, v);
      tmem_wait_ld();
      bias_relu_pack32(v, bs, a);
      tmem_ld32(lane_addr + kColD + 32, v);
      tmem_wait_ld();
      bias_relu_pack32(v, bs + 32, a + 16);
      tmem_st32(lane_addr + kColA, a);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_layer<16, 64>(tmem, sB + kOffB1, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 1: h = acc + b1; sigma = exp(h0 + 1); geo features -> A (16 fp16 = 8 columns) --------------
    {
      uint32_t v[16], a[8];
      tmem_ld16(lane_addr + kColD, v);
      tmem_wait_ld();
      float h[16];
#pragma unroll
      for (int j = 0; j < 16; j++) h[j] = __uint_as_float(v[j]) + bs[64 + j];
      if (valid) sigma[row] = expf(h[0] + 1.f);  // trunc_exp(h0 + 1), nerfacto_field.py:499
      h[0] = 0.f;                                 // column 0 of the geo tile has zero weights
#pragma unroll
      for (int j = 0; j < 8; j++) a[j] = pack_h2(h[2 * j], h[2 * j + 1]);
      tmem_st8(lane_addr + kColA, a);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_layer<64, 16>(tmem, sB + kOffB2, bar);
    }
    // the ray's bias row (b2 + W2[:, SH] SH(d) + W2[:, emb] emb) while the MMA runs
    const float* rb = ray_bias + (int64_t)(valid ? __ldg(ray_id + row) : 0) * kH;
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 2: h2 = relu(acc + ray_bias[ray]) -------------------------------------------------------
    {
      uint32_t v[32], a[32];
      tmem_ld32(lane_addr + kColD, v);
      tmem_wait_ld();
      bias_relu_pack32(v, rb, a);
      tmem_ld32(lane_addr + kColD + 32, v);
      tmem_wait_ld();
      bias_relu_pack32(v, rb + 32, a + 16);
      tmem_st32(lane_addr + kColA, a);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_layer<64, 64>(tmem, sB + kOffB3, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 3: h3 = relu(acc + b3) -------------------------------------------------------------------
    {
      uint32_t v[32], a[32];
      tmem_ld32(lane_addr + kColD, v);
      tmem_wait_ld();
      bias_relu_pack32(v, bs + 80, a);
      tmem_ld32(lane_addr + kColD + 32, v);
      tmem_wait_ld();
      bias_relu_pack32(v, bs + 112, a + 16);
      tmem_st32(lane_addr + kColA, a);
    }
    tmem_wait_st();
    tc_fence_before();
    __syncthreads();
    if (tid == 0) {
      tc_fence_after();
      issue_layer<16, 64>(tmem, sB + kOffB4, bar);
    }
    mbar_wait(bar, phase);
    phase ^= 1;
    tc_fence_after();
    // ---- epilogue 4: rgb = sigmoid(acc + b4) ------------------------------------------------------------------
    {
      uint32_t v[4];
      tmem_ld4(lane_addr + kColD, v);
      tmem_wait_ld();
      if (valid) {
        rgb[3 * row] = sigmoidf_(__uint_as_float(v[0]) + bs[144]);
        rgb[3 * row + 1] = sigmoidf_(__uint_as_float(v[1]) + bs[145]);
        rgb[3 * row + 2] = sigmoidf_(__uint_as_float(v[2]) + bs[146]);
      }
    }
    // the next tile's tcgen05.st into A and MMA into D are ordered behind this tile's loads by the
    // fence + __syncthreads at the top of the next iteration
    tc_fence_before();
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem, kTmemCols);
}

}  // namespace tc
}  // namespace gf

using namespace gf;

// launched by gf_mlp_forward (mlp.cu)
int gf_launch_mlp_fwd_tc(int64_t n, const int32_t* d_n_ptr, const float* params, const void* feat_f16,
                         const int32_t* ray_id, const float* ray_bias, float* sigma, float* rgb, cudaStream_t st) {
  const int64_t tiles = div_up(n, tc::kTile);
  const int grid = (int)std::min<int64_t>(tiles, (int64_t)sm_count() * 4);
  // 46 KB of dynamic shared memory (19 KB used) caps residency at four CTAs per SM = 4 x 128 TMEM columns, so a
  // fifth CTA can never sit in tcgen05.alloc waiting for columns
  static_assert(tc::kFwdSmem <= 46 * 1024, "forward smem");
  tc::mlp_fwd_tc_kernel<<<grid, tc::kThreads, 46 * 1024, st>>>(n, d_n_ptr, params, (const __half*)feat_f16, ray_id,
                                                                 ray_bias, sigma, rgb);
  return check_launch("mlp_fwd_tc_kernel");
}
