// Fused field MLP (density net + colour head), forward and backward, for sm_100a.
//
// Replaces the two MLPNetwork stacks of the reference field (gfnerf/mlp.py:25-57, built at
// gfnerf/nerfacto_field.py:174-179, 217-227), trunc_exp(x + 1) (:499, nerfstudio/
// field_components/activations.py:23-38), the tcnn SH degree-4 direction encoding (:152-158,
// 521), the appearance-embedding lookup (:530-537) and the [SH | geo | emb] concat (:540-547),
// plus everything autograd does for them.
//
// Design (DESIGN.md "field MLP"):
//  * the reference round-trips every activation through HBM (five cuBLAS SGEMMs, ~1.2 KB per
//    point each way).  Here a warp carries 16 or 32 points through all five layers in
//    registers: the fp32 accumulator fragment of one m16n8k16 tensor-core MMA is, after
//    ReLU + fp16 packing, exactly the A fragment of the next layer, so nothing but the 64-byte
//    feature row is read and 16 bytes (sigma, rgb) are written per point.
//  * the head's first layer is split: SH(dir) and the appearance embedding are constant along a
//    ray, so W2[:, SH|emb] . [SH; emb] + b2 is evaluated once per RAY in fp32 (ray_bias kernel)
//    and enters the per-point layer as the accumulator's initial value; only the 15 geo
//    features go through the per-point MMA (K = 16 instead of 63).
//  * backward recomputes the forward (no activation is stored), runs the dgrad chain in
//    registers, and stages the fp16 activations / gradients of the CTA's 128 points in shared
//    memory, from where the eight warps accumulate all weight gradients with transposed
//    ldmatrix fragments into register accumulators that live for the whole persistent kernel;
//    one atomicAdd per parameter per CTA at the end.  Bias gradients ride along as an MMA
//    against an all-ones B fragment.  d/d(ray_bias) is reduced per warp with shuffles.
//
// Weights are fp16 in shared memory (converted from the fp32 blob at kernel start), fp32
// accumulation: the "fp16 MLP" precision class of the north star (1e-2 relative).
#include <stdlib.h>

#include "common.cuh"

namespace gf {

constexpr int kH = 64;
// parameter blob offsets (torch nn.Linear layout), see include/gfnerf_b200.h
constexpr int kW0 = 0, kB0 = kW0 + kH * 32, kW1 = kB0 + kH, kB1 = kW1 + 16 * kH, kW2 = kB1 + 16,
              kB2 = kW2 + kH * 63, kW3 = kB2 + kH, kB3 = kW3 + kH * kH, kW4 = kB3 + kH, kB4 = kW4 + 3 * kH,
              kParamCount = kB4 + 3;

// shared-memory weight tiles, fp16, rows padded so that fragment loads are bank-conflict free
constexpr int kS0 = 40, kS1 = 72, kS2 = 24, kS3 = 72, kS4 = 72;  // row strides in halves
constexpr int kOffW0 = 0, kOffW1 = kOffW0 + 64 * kS0, kOffW2 = kOffW1 + 16 * kS1, kOffW3 = kOffW2 + 64 * kS2,
              kOffW4 = kOffW3 + 64 * kS3, kWHalves = kOffW4 + 8 * kS4;
constexpr int kBiasFloats = 64 + 16 + 64 + 4;  // b0 | b1 | b3 | b4
constexpr int kWBytes = kWHalves * 2, kWSmemBytes = kWBytes + kBiasFloats * 4;
static_assert(kWBytes % 16 == 0, "bias block must stay 16-byte aligned");

__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t relu_pack(float lo, float hi) { return pack2(fmaxf(lo, 0.f), fmaxf(hi, 0.f)); }
__device__ __forceinline__ float2 unpack2(uint32_t v) { return __half22float2(*reinterpret_cast<__half2*>(&v)); }

__device__ __forceinline__ void mma16816(float (&c)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};\n"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}

__device__ __forceinline__ void ldsm_x4_trans(uint32_t (&r)[4], const __half* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];\n"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm_x2_trans(uint32_t (&r)[2], const __half* p) {
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(p);
  asm volatile("ldmatrix.sync.aligned.m8n8.x2.trans.shared.b16 {%0,%1}, [%2];\n" : "=r"(r[0]), "=r"(r[1]) : "r"(addr));
}

__device__ __forceinline__ uint32_t lds32(const __half* p) { return *reinterpret_cast<const uint32_t*>(p); }
__device__ __forceinline__ void sts32(__half* p, uint32_t v) { *reinterpret_cast<uint32_t*>(p) = v; }

// fp32 parameter blob -> fp16 shared tiles (+ fp32 biases of the layers whose bias is an accumulator init)
__device__ __forceinline__ void stage_weights(const float* __restrict__ p, __half* ws, float* bs) {
  for (int i = threadIdx.x; i < kWHalves; i += blockDim.x) ws[i] = __float2half_rn(0.f);
  __syncthreads();
  for (int i = threadIdx.x; i < kH * 32; i += blockDim.x) ws[kOffW0 + (i >> 5) * kS0 + (i & 31)] = __float2half_rn(__ldg(p + kW0 + i));
  for (int i = threadIdx.x; i < 16 * kH; i += blockDim.x) ws[kOffW1 + (i >> 6) * kS1 + (i & 63)] = __float2half_rn(__ldg(p + kW1 + i));
  // geo columns of the head's first layer: column c (1..15) multiplies h[c] = in2[15 + c]; column 0 (h[0], the
  // density logit) stays zero
  for (int i = threadIdx.x; i < kH * 15; i += blockDim.x) {
    const int j = i / 15, c = i % 15;
    ws[kOffW2 + j * kS2 + 1 + c] = __float2half_rn(__ldg(p + kW2 + j * 63 + 16 + c));
  }
  for (int i = threadIdx.x; i < kH * kH; i += blockDim.x) ws[kOffW3 + (i >> 6) * kS3 + (i & 63)] = __float2half_rn(__ldg(p + kW3 + i));
  for (int i = threadIdx.x; i < 3 * kH; i += blockDim.x) ws[kOffW4 + (i >> 6) * kS4 + (i & 63)] = __float2half_rn(__ldg(p + kW4 + i));
  for (int i = threadIdx.x; i < kBiasFloats; i += blockDim.x) {
    float v;
    if (i < 64) v = __ldg(p + kB0 + i);
    else if (i < 80) v = __ldg(p + kB1 + i - 64);
    else if (i < 144) v = __ldg(p + kB3 + i - 80);
    else v = (i - 144) < 3 ? __ldg(p + kB4 + i - 144) : 0.f;
    bs[i] = v;
  }
}

// acc[mt][nt] += A[mt] (16 x 16*KB) . W^T  with W = ws[n][k] (row stride S halves)
template <int MT, int NT, int KB, int S>
__device__ __forceinline__ void layer_fwd(const __half* ws, const uint32_t (&a)[MT][KB][4], float (&acc)[MT][NT][4],
                                          int g, int t) {
#pragma unroll
  for (int kb = 0; kb < KB; kb++) {
#pragma unroll
    for (int nt = 0; nt < NT; nt++) {
      const __half* wp = ws + (8 * nt + g) * S + 16 * kb + 2 * t;
      const uint32_t b0 = lds32(wp), b1 = lds32(wp + 8);
#pragma unroll
      for (int mt = 0; mt < MT; mt++) mma16816(acc[mt][nt], a[mt][kb], b0, b1);
    }
  }
}

// acc[nt] += A (16 x 16*KB) . W  with W = ws[k][n] (dgrad: the same tile read transposed)
template <int NT, int KB, int S>
__device__ __forceinline__ void layer_bwd(const __half* ws, const uint32_t (&a)[KB][4], float (&acc)[NT][4], int lane) {
  static_assert(NT % 2 == 0, "n-tiles are loaded in pairs");
  const int i = lane >> 3, r = lane & 7;
#pragma unroll
  for (int kb = 0; kb < KB; kb++) {
#pragma unroll
    for (int np = 0; np < NT / 2; np++) {
      uint32_t b[4];
      ldsm_x4_trans(b, ws + (16 * kb + 8 * (i & 1) + r) * S + 8 * (2 * np + (i >> 1)));
      mma16816(acc[2 * np], a[kb], b[0], b[1]);
      mma16816(acc[2 * np + 1], a[kb], b[2], b[3]);
    }
  }
}

template <int NT>
__device__ __forceinline__ void init_bias(float (&acc)[NT][4], const float* b, int t) {
#pragma unroll
  for (int nt = 0; nt < NT; nt++) {
    const float2 v = *reinterpret_cast<const float2*>(b + 8 * nt + 2 * t);
    acc[nt][0] = v.x; acc[nt][1] = v.y; acc[nt][2] = v.x; acc[nt][3] = v.y;
  }
}

// accumulator fragments of 2*KB n-tiles -> A fragments of KB k-blocks
template <int KB, bool RELU>
__device__ __forceinline__ void acc_to_a(const float (&acc)[2 * KB][4], uint32_t (&a)[KB][4]) {
#pragma unroll
  for (int kb = 0; kb < KB; kb++) {
    if (RELU) {
      a[kb][0] = relu_pack(acc[2 * kb][0], acc[2 * kb][1]);
      a[kb][1] = relu_pack(acc[2 * kb][2], acc[2 * kb][3]);
      a[kb][2] = relu_pack(acc[2 * kb + 1][0], acc[2 * kb + 1][1]);
      a[kb][3] = relu_pack(acc[2 * kb + 1][2], acc[2 * kb + 1][3]);
    } else {
      a[kb][0] = pack2(acc[2 * kb][0], acc[2 * kb][1]);
      a[kb][1] = pack2(acc[2 * kb][2], acc[2 * kb][3]);
      a[kb][2] = pack2(acc[2 * kb + 1][0], acc[2 * kb + 1][1]);
      a[kb][3] = pack2(acc[2 * kb + 1][2], acc[2 * kb + 1][3]);
    }
  }
}

__device__ __forceinline__ float sigmoidf(float x) { return 1.f / (1.f + __expf(-x)); }

// feature rows of one m-tile as A fragments (rows >= n read as zero)
__device__ __forceinline__ void load_feat(const __half* __restrict__ feat, int64_t r_lo, int64_t r_hi, int64_t n, int t,
                                          uint32_t (&a)[2][4]) {
#pragma unroll
  for (int kb = 0; kb < 2; kb++) {
    const uint32_t* lo = reinterpret_cast<const uint32_t*>(feat + r_lo * 32 + 16 * kb + 2 * t);
    const uint32_t* hi = reinterpret_cast<const uint32_t*>(feat + r_hi * 32 + 16 * kb + 2 * t);
    a[kb][0] = r_lo < n ? __ldg(lo) : 0u;
    a[kb][1] = r_hi < n ? __ldg(hi) : 0u;
    a[kb][2] = r_lo < n ? __ldg(lo + 4) : 0u;
    a[kb][3] = r_hi < n ? __ldg(hi + 4) : 0u;
  }
}

__device__ __forceinline__ void init_ray_bias(float (&acc)[8][4], const float* __restrict__ ray_bias, int ray_lo,
                                              int ray_hi, int t) {
#pragma unroll
  for (int nt = 0; nt < 8; nt++) {
    float2 lo = make_float2(0.f, 0.f), hi = make_float2(0.f, 0.f);
    if (ray_lo >= 0) lo = __ldg(reinterpret_cast<const float2*>(ray_bias + (int64_t)ray_lo * kH + 8 * nt + 2 * t));
    if (ray_hi >= 0) hi = __ldg(reinterpret_cast<const float2*>(ray_bias + (int64_t)ray_hi * kH + 8 * nt + 2 * t));
    acc[nt][0] = lo.x; acc[nt][1] = lo.y; acc[nt][2] = hi.x; acc[nt][3] = hi.y;
  }
}

// ---------------------------------------------------------------------------------------------
// per-ray part of the head's first layer
// ---------------------------------------------------------------------------------------------
// tcnn SphericalHarmonics degree 4 on (dir+1)/2, fp16 output (gfnerf/nerfacto_field.py:64-70, 152-158, 521)
__device__ __forceinline__ void sh4(float dx, float dy, float dz, float (&o)[16]) {
  const float x = ((dx + 1.f) * .5f) * 2.f - 1.f, y = ((dy + 1.f) * .5f) * 2.f - 1.f, z = ((dz + 1.f) * .5f) * 2.f - 1.f;
  const float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  o[0] = 0.28209479177387814f;
  o[1] = -0.48860251190291987f * y;
  o[2] = 0.48860251190291987f * z;
  o[3] = -0.48860251190291987f * x;
  o[4] = 1.0925484305920792f * xy;
  o[5] = -1.0925484305920792f * yz;
  o[6] = 0.94617469575755997f * z2 - 0.31539156525251999f;
  o[7] = -1.0925484305920792f * xz;
  o[8] = 0.54627421529603959f * x2 - 0.54627421529603959f * y2;
  o[9] = 0.59004358992664352f * y * (-3.0f * x2 + y2);
  o[10] = 2.8906114426405538f * xy * z;
  o[11] = 0.45704579946446572f * y * (1.0f - 5.0f * z2);
  o[12] = 0.3731763325901154f * z * (5.0f * z2 - 3.0f);
  o[13] = 0.45704579946446572f * x * (1.0f - 5.0f * z2);
  o[14] = 1.4453057213202769f * z * (x2 - y2);
  o[15] = 0.59004358992664352f * x * (-x2 + 3.0f * y2);
#pragma unroll
  for (int i = 0; i < 16; i++) o[i] = __half2float(__float2half_rn(o[i]));
}

constexpr int kRbRays = 32;  // rays per CTA of the ray-bias kernels
constexpr int kRbIn = 48;    // SH(16) | emb(32)

// ray_bias[r][j] = b2[j] + W2[j][0:16] . SH(dir_r) + W2[j][31:63] . emb_r      (fp32)
__global__ void __launch_bounds__(256)
ray_bias_kernel(int64_t n_rays, const float* __restrict__ params, const float* __restrict__ dirs,
                const float* __restrict__ emb, float* __restrict__ ray_bias) {
  __shared__ float s_w[kH][kRbIn + 1];
  __shared__ float s_in[kRbRays][kRbIn + 1];
  for (int i = threadIdx.x; i < kH * kRbIn; i += blockDim.x) {
    const int j = i / kRbIn, k = i % kRbIn;
    s_w[j][k] = __ldg(params + kW2 + j * 63 + (k < 16 ? k : 15 + k));
  }
  const int64_t r0 = (int64_t)blockIdx.x * kRbRays;
  if (threadIdx.x < kRbRays) {
    const int64_t r = r0 + threadIdx.x;
    float o[16];
    if (r < n_rays) sh4(__ldg(dirs + 3 * r), __ldg(dirs + 3 * r + 1), __ldg(dirs + 3 * r + 2), o);
    else
      for (int i = 0; i < 16; i++) o[i] = 0.f;
    for (int i = 0; i < 16; i++) s_in[threadIdx.x][i] = o[i];
  }
  for (int i = threadIdx.x; i < kRbRays * 32; i += blockDim.x) {
    const int rr = i >> 5, k = i & 31;
    const int64_t r = r0 + rr;
    s_in[rr][16 + k] = (emb && r < n_rays) ? __ldg(emb + r * 32 + k) : 0.f;
  }
  __syncthreads();
  for (int i = threadIdx.x; i < kRbRays * kH; i += blockDim.x) {
    const int rr = i >> 6, j = i & 63;
    const int64_t r = r0 + rr;
    if (r >= n_rays) continue;
    float acc = __ldg(params + kB2 + j);
#pragma unroll 8
    for (int k = 0; k < kRbIn; k++) acc = fmaf(s_w[j][k], s_in[rr][k], acc);
    ray_bias[r * kH + j] = acc;
  }
}

// G = d_ray_bias [R,64]:  d_b2 += sum_r G ; d_W2[:, SH|emb] += G^T [SH|emb] ; d_emb = G W2[:, emb]
__global__ void __launch_bounds__(256)
ray_bias_bwd_kernel(int64_t n_rays, const float* __restrict__ params, const float* __restrict__ dirs,
                    const float* __restrict__ emb, const float* __restrict__ g_rb, float* __restrict__ d_params,
                    float* __restrict__ d_emb) {
  __shared__ float s_g[kRbRays][kH + 1];
  __shared__ float s_in[kRbRays][kRbIn + 1];
  const int64_t r0 = (int64_t)blockIdx.x * kRbRays;
  if (threadIdx.x < kRbRays) {
    const int64_t r = r0 + threadIdx.x;
    float o[16];
    if (r < n_rays) sh4(__ldg(dirs + 3 * r), __ldg(dirs + 3 * r + 1), __ldg(dirs + 3 * r + 2), o);
    else
      for (int i = 0; i < 16; i++) o[i] = 0.f;
    for (int i = 0; i < 16; i++) s_in[threadIdx.x][i] = o[i];
  }
  for (int i = threadIdx.x; i < kRbRays * 32; i += blockDim.x) {
    const int rr = i >> 5, k = i & 31;
    const int64_t r = r0 + rr;
    s_in[rr][16 + k] = (emb && r < n_rays) ? __ldg(emb + r * 32 + k) : 0.f;
  }
  for (int i = threadIdx.x; i < kRbRays * kH; i += blockDim.x) {
    const int rr = i >> 6, j = i & 63;
    const int64_t r = r0 + rr;
    s_g[rr][j] = r < n_rays ? __ldg(g_rb + r * kH + j) : 0.f;
  }
  __syncthreads();
  if (d_params) {
    // (j, k) outputs: k < 48 weight columns, k == 48 the bias
    for (int i = threadIdx.x; i < kH * (kRbIn + 1); i += blockDim.x) {
      const int j = i / (kRbIn + 1), k = i % (kRbIn + 1);
      float acc = 0.f;
      if (k < kRbIn) {
#pragma unroll 8
        for (int rr = 0; rr < kRbRays; rr++) acc = fmaf(s_g[rr][j], s_in[rr][k], acc);
        if (k >= 16 && !emb) continue;
        if (acc != 0.f) atomicAdd(d_params + kW2 + j * 63 + (k < 16 ? k : 15 + k), acc);
      } else {
#pragma unroll 8
        for (int rr = 0; rr < kRbRays; rr++) acc += s_g[rr][j];
        if (acc != 0.f) atomicAdd(d_params + kB2 + j, acc);
      }
    }
  }
  if (d_emb) {
    for (int i = threadIdx.x; i < kRbRays * 32; i += blockDim.x) {
      const int rr = i >> 5, k = i & 31;
      const int64_t r = r0 + rr;
      if (r >= n_rays) continue;
      float acc = 0.f;
#pragma unroll 8
      for (int j = 0; j < kH; j++) acc = fmaf(s_g[rr][j], __ldg(params + kW2 + j * 63 + 31 + k), acc);
      d_emb[r * 32 + k] += acc;
    }
  }
}

// ---------------------------------------------------------------------------------------------
// forward
// ---------------------------------------------------------------------------------------------
constexpr int kFwdBlock = 256;

template <int MT>
__global__ void __launch_bounds__(kFwdBlock)
mlp_fwd_kernel(int64_t n, const int32_t* __restrict__ d_n_ptr, const float* __restrict__ params,
               const __half* __restrict__ feat, const int32_t* __restrict__ ray_id,
               const float* __restrict__ ray_bias, float* __restrict__ sigma, float* __restrict__ rgb) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __half* ws = reinterpret_cast<__half*>(smem_raw);
  float* bs = reinterpret_cast<float*>(smem_raw + kWBytes);
  stage_weights(params, ws, bs);
  __syncthreads();
  if (d_n_ptr) {
    const int64_t dn = *d_n_ptr;
    n = dn < n ? dn : n;
  }
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  const int64_t gwarp = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int64_t n_warps = ((int64_t)gridDim.x * blockDim.x) >> 5;
  const int64_t n_tiles = (n + 16 * MT - 1) / (16 * MT);
  for (int64_t tile = gwarp; tile < n_tiles; tile += n_warps) {
    const int64_t p0 = tile * 16 * MT;
    uint32_t a0[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) load_feat(feat, p0 + 16 * mt + g, p0 + 16 * mt + g + 8, n, t, a0[mt]);
    float acc[MT][8][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) init_bias<8>(acc[mt], bs, t);
    layer_fwd<MT, 8, 2, kS0>(ws + kOffW0, a0, acc, g, t);
    uint32_t a1[MT][4][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) acc_to_a<4, true>(acc[mt], a1[mt]);
    float acch[MT][2][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) init_bias<2>(acch[mt], bs + 64, t);
    layer_fwd<MT, 2, 4, kS1>(ws + kOffW1, a1, acch, g, t);
    uint32_t a2[MT][1][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      const int64_t r_lo = p0 + 16 * mt + g, r_hi = r_lo + 8;
      if (t == 0) {  // trunc_exp(h0 + 1)
        if (r_lo < n) sigma[r_lo] = expf(acch[mt][0][0] + 1.f);
        if (r_hi < n) sigma[r_hi] = expf(acch[mt][0][2] + 1.f);
      }
      acc_to_a<1, false>(acch[mt], a2[mt]);
      const int ray_lo = r_lo < n ? __ldg(ray_id + r_lo) : -1, ray_hi = r_hi < n ? __ldg(ray_id + r_hi) : -1;
      init_ray_bias(acc[mt], ray_bias, ray_lo, ray_hi, t);
    }
    layer_fwd<MT, 8, 1, kS2>(ws + kOffW2, a2, acc, g, t);
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      acc_to_a<4, true>(acc[mt], a1[mt]);
      init_bias<8>(acc[mt], bs + 80, t);
    }
    layer_fwd<MT, 8, 4, kS3>(ws + kOffW3, a1, acc, g, t);
    float acco[MT][1][4];
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      acc_to_a<4, true>(acc[mt], a1[mt]);
      init_bias<1>(acco[mt], bs + 144, t);
    }
    layer_fwd<MT, 1, 4, kS4>(ws + kOffW4, a1, acco, g, t);
#pragma unroll
    for (int mt = 0; mt < MT; mt++) {
      const int64_t r_lo = p0 + 16 * mt + g, r_hi = r_lo + 8;
      if (t == 0) {
        if (r_lo < n) { rgb[3 * r_lo] = sigmoidf(acco[mt][0][0]); rgb[3 * r_lo + 1] = sigmoidf(acco[mt][0][1]); }
        if (r_hi < n) { rgb[3 * r_hi] = sigmoidf(acco[mt][0][2]); rgb[3 * r_hi + 1] = sigmoidf(acco[mt][0][3]); }
      } else if (t == 1) {
        if (r_lo < n) rgb[3 * r_lo + 2] = sigmoidf(acco[mt][0][0]);
        if (r_hi < n) rgb[3 * r_hi + 2] = sigmoidf(acco[mt][0][2]);
      }
    }
  }
}

// ---------------------------------------------------------------------------------------------
// backward
// ---------------------------------------------------------------------------------------------
constexpr int kBwdWarps = 8, kBwdBlock = kBwdWarps * 32, kBwdPts = kBwdWarps * 16;
constexpr int kSX = 40, kSA = 72, kSh = 24;  // row strides (halves) of the point-major operand tiles
constexpr int kTX = 0, kTH1 = kTX + kBwdPts * kSX, kTHh = kTH1 + kBwdPts * kSA, kTH2 = kTHh + kBwdPts * kSh,
              kTH3 = kTH2 + kBwdPts * kSA, kTG1 = kTH3 + kBwdPts * kSA, kTGh = kTG1 + kBwdPts * kSA,
              kTG2 = kTGh + kBwdPts * kSh, kTG3 = kTG2 + kBwdPts * kSA, kTGo = kTG3 + kBwdPts * kSA,
              kTileHalves = kTGo + kBwdPts * kSh;
constexpr int kBwdSmemBytes = kWSmemBytes + kTileHalves * 2;
static_assert(kWSmemBytes % 16 == 0, "operand tiles must be 16-byte aligned for ldmatrix");

// A fragments of KB k-blocks -> rows [row0+g], [row0+g+8] of a point-major tile
template <int KB, int S>
__device__ __forceinline__ void store_a(__half* tile, int row0, int g, int t, const uint32_t (&a)[KB][4]) {
  __half* lo = tile + (row0 + g) * S + 2 * t;
  __half* hi = lo + 8 * S;
#pragma unroll
  for (int kb = 0; kb < KB; kb++) {
    sts32(lo + 16 * kb, a[kb][0]);
    sts32(hi + 16 * kb, a[kb][1]);
    sts32(lo + 16 * kb + 8, a[kb][2]);
    sts32(hi + 16 * kb + 8, a[kb][3]);
  }
}

// gradient accumulators masked by relu'(act) (act as packed fp16 A fragments), result as A fragments;
// the masked fp32 values are written back to acc
template <int KB>
__device__ __forceinline__ void mask_to_a(float (&acc)[2 * KB][4], const uint32_t (&act)[KB][4], uint32_t (&a)[KB][4]) {
#pragma unroll
  for (int kb = 0; kb < KB; kb++) {
#pragma unroll
    for (int q = 0; q < 4; q++) {
      const float2 h = unpack2(act[kb][q]);
      float* c = acc[2 * kb + (q >> 1)] + 2 * (q & 1);
      c[0] = h.x > 0.f ? c[0] : 0.f;
      c[1] = h.y > 0.f ? c[1] : 0.f;
      a[kb][q] = pack2(c[0], c[1]);
    }
  }
}

// A^T (grad tile, [pt][m]) fragment of the 16 x 16 block (m0.., k0..)
__device__ __forceinline__ void load_at(const __half* tile, int S, int k0, int m0, int lane, uint32_t (&a)[4]) {
  const int i = lane >> 3, r = lane & 7;
  ldsm_x4_trans(a, tile + (k0 + 8 * (i >> 1) + r) * S + m0 + 8 * (i & 1));
}
// B fragments (act tile, [pt][n]) of two n-tiles n0, n0+8
__device__ __forceinline__ void load_b2(const __half* tile, int S, int k0, int n0, int lane, uint32_t (&b)[4]) {
  const int i = lane >> 3, r = lane & 7;
  ldsm_x4_trans(b, tile + (k0 + 8 * (i & 1) + r) * S + n0 + 8 * (i >> 1));
}
__device__ __forceinline__ void load_b1(const __half* tile, int S, int k0, int n0, int lane, uint32_t (&b)[2]) {
  const int i = (lane >> 3) & 1, r = lane & 7;
  ldsm_x2_trans(b, tile + (k0 + 8 * i + r) * S + n0);
}

template <bool WGRAD>
__global__ void __launch_bounds__(kBwdBlock, 1)
mlp_bwd_kernel(int64_t n, const int32_t* __restrict__ d_n_ptr, const float* __restrict__ params,
               const __half* __restrict__ feat, const int32_t* __restrict__ ray_id,
               const float* __restrict__ ray_bias, const float* __restrict__ d_sigma,
               const float* __restrict__ d_rgb, __half* __restrict__ d_feat, float* __restrict__ d_params,
               float* __restrict__ d_ray_bias, float gscale) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const float inv_gscale = 1.f / gscale;
  __half* ws = reinterpret_cast<__half*>(smem_raw);
  float* bs = reinterpret_cast<float*>(smem_raw + kWBytes);
  __half* tiles = reinterpret_cast<__half*>(smem_raw + kWSmemBytes);
  stage_weights(params, ws, bs);
  __syncthreads();
  if (d_n_ptr) {
    const int64_t dn = *d_n_ptr;
    n = dn < n ? dn : n;
  }
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3, warp = threadIdx.x >> 5;
  const int mt = warp & 3, nh = warp >> 2;  // weight-gradient slice of this warp

  // weight-gradient accumulators (fp32), alive for the whole kernel
  float gW3[4][4], gW0[2][4], gW2[4], gW1[4], gW4[4], gB3[4], gB0[4], gB1[4], gB4[4];
#pragma unroll
  for (int q = 0; q < 4; q++) {
#pragma unroll
    for (int j = 0; j < 4; j++) gW3[j][q] = 0.f;
    gW0[0][q] = gW0[1][q] = gW2[q] = gW1[q] = gW4[q] = gB3[q] = gB0[q] = gB1[q] = gB4[q] = 0.f;
  }

  const int64_t n_ctiles = (n + kBwdPts - 1) / kBwdPts;
  for (int64_t ct = blockIdx.x; ct < n_ctiles; ct += gridDim.x) {
    const int row0 = 16 * warp;
    const int64_t r_lo = ct * kBwdPts + row0 + g, r_hi = r_lo + 8;
    const bool v_lo = r_lo < n, v_hi = r_hi < n;
    // ---- forward recompute ------------------------------------------------------------
    uint32_t ax[1][2][4];
    load_feat(feat, r_lo, r_hi, n, t, ax[0]);
    float acc[1][8][4];
    init_bias<8>(acc[0], bs, t);
    layer_fwd<1, 8, 2, kS0>(ws + kOffW0, ax, acc, g, t);
    uint32_t ah1[1][4][4];
    acc_to_a<4, true>(acc[0], ah1[0]);
    float acch[1][2][4];
    init_bias<2>(acch[0], bs + 64, t);
    layer_fwd<1, 2, 4, kS1>(ws + kOffW1, ah1, acch, g, t);
    const float pre_lo = acch[0][0][0] + 1.f, pre_hi = acch[0][0][2] + 1.f;  // density logit + 1 (lanes t == 0)
    uint32_t ahh[1][1][4];
    acc_to_a<1, false>(acch[0], ahh[0]);
    const int ray_lo = v_lo ? __ldg(ray_id + r_lo) : -1, ray_hi = v_hi ? __ldg(ray_id + r_hi) : -1;
    init_ray_bias(acc[0], ray_bias, ray_lo, ray_hi, t);
    layer_fwd<1, 8, 1, kS2>(ws + kOffW2, ahh, acc, g, t);
    uint32_t ah2[1][4][4];
    acc_to_a<4, true>(acc[0], ah2[0]);
    init_bias<8>(acc[0], bs + 80, t);
    layer_fwd<1, 8, 4, kS3>(ws + kOffW3, ah2, acc, g, t);
    uint32_t ah3[1][4][4];
    acc_to_a<4, true>(acc[0], ah3[0]);
    float acco[1][1][4];
    init_bias<1>(acco[0], bs + 144, t);
    layer_fwd<1, 1, 4, kS4>(ws + kOffW4, ah3, acco, g, t);
    // ---- dgrad chain (all gradients carry the factor gscale while they are fp16) ---------
    // d o = d rgb * s (1 - s); columns 2t, 2t+1 of the 8-wide (3 real) output tile
    float go[4] = {0.f, 0.f, 0.f, 0.f};
    if (t < 2) {
      const int c0 = 2 * t;
      if (v_lo) {
        const float s0 = sigmoidf(acco[0][0][0]);
        go[0] = __ldg(d_rgb + 3 * r_lo + c0) * gscale * s0 * (1.f - s0);
        if (t == 0) {
          const float s1 = sigmoidf(acco[0][0][1]);
          go[1] = __ldg(d_rgb + 3 * r_lo + 1) * gscale * s1 * (1.f - s1);
        }
      }
      if (v_hi) {
        const float s0 = sigmoidf(acco[0][0][2]);
        go[2] = __ldg(d_rgb + 3 * r_hi + c0) * gscale * s0 * (1.f - s0);
        if (t == 0) {
          const float s1 = sigmoidf(acco[0][0][3]);
          go[3] = __ldg(d_rgb + 3 * r_hi + 1) * gscale * s1 * (1.f - s1);
        }
      }
    }
    uint32_t ago[1][4] = {{pack2(go[0], go[1]), pack2(go[2], go[3]), 0u, 0u}};
    // g h3 = g o . W4  (k = 3 real rows of the 8-row tile; rows 8..15 of the k-block are zero)
    float (&ga)[8][4] = acc[0];
#pragma unroll
    for (int nt = 0; nt < 8; nt++) ga[nt][0] = ga[nt][1] = ga[nt][2] = ga[nt][3] = 0.f;
    {
      const int i = lane >> 3, r = lane & 7;
#pragma unroll
      for (int q = 0; q < 2; q++) {
        uint32_t b[4];
        ldsm_x4_trans(b, ws + kOffW4 + r * kS4 + 8 * (4 * q + i));
#pragma unroll
        for (int j = 0; j < 4; j++) mma16816(ga[4 * q + j], ago[0], b[j], 0u);
      }
    }
    uint32_t ag3[4][4];
    mask_to_a<4>(ga, ah3[0], ag3);
    if (WGRAD) {
      store_a<4, kSA>(tiles + kTH3, row0, g, t, ah3[0]);
      store_a<4, kSA>(tiles + kTG3, row0, g, t, ag3);
      store_a<1, kSh>(tiles + kTGo, row0, g, t, ago);
    }
    // g h2 = g h3 . W3
#pragma unroll
    for (int nt = 0; nt < 8; nt++) ga[nt][0] = ga[nt][1] = ga[nt][2] = ga[nt][3] = 0.f;
    layer_bwd<8, 4, kS3>(ws + kOffW3, ag3, ga, lane);
    uint32_t ag2[4][4];
    mask_to_a<4>(ga, ah2[0], ag2);
    if (WGRAD) {
      store_a<4, kSA>(tiles + kTH2, row0, g, t, ah2[0]);
      store_a<4, kSA>(tiles + kTG2, row0, g, t, ag2);
      // d ray_bias[ray] += column sums of g h2 over the rows of that ray
      const int ray0 = __shfl_sync(0xffffffffu, ray_lo, 0);
      const bool uniform = __all_sync(0xffffffffu, ray_lo == ray0 && ray_hi == ray0) && ray0 >= 0;
      if (uniform) {
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          float s0 = ga[nt][0] + ga[nt][2], s1 = ga[nt][1] + ga[nt][3];
#pragma unroll
          for (int off = 4; off < 32; off <<= 1) {
            s0 += __shfl_xor_sync(0xffffffffu, s0, off);
            s1 += __shfl_xor_sync(0xffffffffu, s1, off);
          }
          if (g == 0) {
            float* dst = d_ray_bias + (int64_t)ray0 * kH + 8 * nt + 2 * t;
            if (s0 != 0.f) atomicAdd(dst, s0 * inv_gscale);
            if (s1 != 0.f) atomicAdd(dst + 1, s1 * inv_gscale);
          }
        }
      } else {
#pragma unroll
        for (int nt = 0; nt < 8; nt++) {
          if (ray_lo >= 0) {
            float* dst = d_ray_bias + (int64_t)ray_lo * kH + 8 * nt + 2 * t;
            if (ga[nt][0] != 0.f) atomicAdd(dst, ga[nt][0] * inv_gscale);
            if (ga[nt][1] != 0.f) atomicAdd(dst + 1, ga[nt][1] * inv_gscale);
          }
          if (ray_hi >= 0) {
            float* dst = d_ray_bias + (int64_t)ray_hi * kH + 8 * nt + 2 * t;
            if (ga[nt][2] != 0.f) atomicAdd(dst, ga[nt][2] * inv_gscale);
            if (ga[nt][3] != 0.f) atomicAdd(dst + 1, ga[nt][3] * inv_gscale);
          }
        }
      }
    }
    // g h = [ d sigma * exp(clamp(h0 + 1)) | g h2 . W2[:, geo] ]
    float accg[2][4];
#pragma unroll
    for (int nt = 0; nt < 2; nt++) accg[nt][0] = accg[nt][1] = accg[nt][2] = accg[nt][3] = 0.f;
    layer_bwd<2, 4, kS2>(ws + kOffW2, ag2, accg, lane);
    if (t == 0) {  // _TruncExp.backward: g * exp(clamp(x, -15, 15))
      if (v_lo) accg[0][0] += __ldg(d_sigma + r_lo) * gscale * expf(fminf(fmaxf(pre_lo, -15.f), 15.f));
      if (v_hi) accg[0][2] += __ldg(d_sigma + r_hi) * gscale * expf(fminf(fmaxf(pre_hi, -15.f), 15.f));
    }
    uint32_t agh[1][4];
    acc_to_a<1, false>(accg, agh);
    if (WGRAD) {
      store_a<1, kSh>(tiles + kTHh, row0, g, t, ahh[0]);
      store_a<1, kSh>(tiles + kTGh, row0, g, t, agh);
    }
    // g h1 = g h . W1
#pragma unroll
    for (int nt = 0; nt < 8; nt++) ga[nt][0] = ga[nt][1] = ga[nt][2] = ga[nt][3] = 0.f;
    layer_bwd<8, 1, kS1>(ws + kOffW1, agh, ga, lane);
    uint32_t ag1[4][4];
    mask_to_a<4>(ga, ah1[0], ag1);
    if (WGRAD) {
      store_a<4, kSA>(tiles + kTH1, row0, g, t, ah1[0]);
      store_a<4, kSA>(tiles + kTG1, row0, g, t, ag1);
      store_a<2, kSX>(tiles + kTX, row0, g, t, ax[0]);
    }
    // g x = g h1 . W0, handed to the hash backward as fp16(g * 128) (Hash3DAnchored_cuda.cu:209)
    float accx[4][4];
#pragma unroll
    for (int nt = 0; nt < 4; nt++) accx[nt][0] = accx[nt][1] = accx[nt][2] = accx[nt][3] = 0.f;
    layer_bwd<4, 4, kS0>(ws + kOffW0, ag1, accx, lane);
    const float out_scale = GF_GRAD_SCALE * inv_gscale;
#pragma unroll
    for (int nt = 0; nt < 4; nt++) {
      if (v_lo)
        *reinterpret_cast<uint32_t*>(d_feat + r_lo * 32 + 8 * nt + 2 * t) =
            pack2(accx[nt][0] * out_scale, accx[nt][1] * out_scale);
      if (v_hi)
        *reinterpret_cast<uint32_t*>(d_feat + r_hi * 32 + 8 * nt + 2 * t) =
            pack2(accx[nt][2] * out_scale, accx[nt][3] * out_scale);
    }
    if (WGRAD) {
      // ---- weight gradients over the CTA's 128 points ----------------------------------
      __syncthreads();
      const uint32_t ones = 0x3C003C00u;
#pragma unroll 2
      for (int kb = 0; kb < kBwdPts / 16; kb++) {
        const int k0 = 16 * kb;
        uint32_t a[4], b[4], b1[2];
        load_at(tiles + kTG3, kSA, k0, 16 * mt, lane, a);  // d W3[16mt.., :] = g h3^T h2
        load_b2(tiles + kTH2, kSA, k0, 32 * nh, lane, b);
        mma16816(gW3[0], a, b[0], b[1]);
        mma16816(gW3[1], a, b[2], b[3]);
        load_b2(tiles + kTH2, kSA, k0, 32 * nh + 16, lane, b);
        mma16816(gW3[2], a, b[0], b[1]);
        mma16816(gW3[3], a, b[2], b[3]);
        if (nh == 0) mma16816(gB3, a, ones, ones);
        load_at(tiles + kTG1, kSA, k0, 16 * mt, lane, a);  // d W0 = g h1^T x
        load_b2(tiles + kTX, kSX, k0, 16 * nh, lane, b);
        mma16816(gW0[0], a, b[0], b[1]);
        mma16816(gW0[1], a, b[2], b[3]);
        if (nh == 1) mma16816(gB0, a, ones, ones);
        load_at(tiles + kTG2, kSA, k0, 16 * mt, lane, a);  // d W2[:, geo] = g h2^T h
        load_b1(tiles + kTHh, kSh, k0, 8 * nh, lane, b1);
        mma16816(gW2, a, b1[0], b1[1]);
        load_at(tiles + kTGh, kSh, k0, 0, lane, a);        // d W1 = g h^T h1
        load_b1(tiles + kTH1, kSA, k0, 8 * warp, lane, b1);
        mma16816(gW1, a, b1[0], b1[1]);
        if (warp == 0) mma16816(gB1, a, ones, ones);
        load_at(tiles + kTGo, kSh, k0, 0, lane, a);        // d W4 = g o^T h3
        load_b1(tiles + kTH3, kSA, k0, 8 * warp, lane, b1);
        mma16816(gW4, a, b1[0], b1[1]);
        if (warp == 1) mma16816(gB4, a, ones, ones);
      }
      __syncthreads();
    }
  }
  if (WGRAD) {
    // one atomic per parameter per CTA; fragment element q of tile (m0, n0): row m0+g+8*(q>>1), col n0+2t+(q&1)
    auto flush = [&](const float (&c)[4], float* base, int ld, int m0, int n0, int m_max, int n_lo, int n_hi,
                     int n_shift) {
#pragma unroll
      for (int q = 0; q < 4; q++) {
        const int m = m0 + g + 8 * (q >> 1), nn = n0 + 2 * t + (q & 1);
        if (m < m_max && nn >= n_lo && nn < n_hi && c[q] != 0.f)
          atomicAdd(base + m * ld + nn + n_shift, c[q] * inv_gscale);
      }
    };
    auto flush_bias = [&](const float (&c)[4], float* base, int m0, int m_max) {
      if (t == 0) {
        if (m0 + g < m_max && c[0] != 0.f) atomicAdd(base + m0 + g, c[0] * inv_gscale);
        if (m0 + g + 8 < m_max && c[2] != 0.f) atomicAdd(base + m0 + g + 8, c[2] * inv_gscale);
      }
    };
#pragma unroll
    for (int j = 0; j < 4; j++) flush(gW3[j], d_params + kW3, kH, 16 * mt, 32 * nh + 8 * j, kH, 0, kH, 0);
    if (nh == 0) flush_bias(gB3, d_params + kB3, 16 * mt, kH);
#pragma unroll
    for (int j = 0; j < 2; j++) flush(gW0[j], d_params + kW0, 32, 16 * mt, 16 * nh + 8 * j, kH, 0, 32, 0);
    if (nh == 1) flush_bias(gB0, d_params + kB0, 16 * mt, kH);
    // geo column c (1..15) of the tile is W2 column 15 + c
    flush(gW2, d_params + kW2, 63, 16 * mt, 8 * nh, kH, 1, 16, 15);
    flush(gW1, d_params + kW1, kH, 0, 8 * warp, 16, 0, kH, 0);
    if (warp == 0) flush_bias(gB1, d_params + kB1, 0, 16);
    flush(gW4, d_params + kW4, kH, 0, 8 * warp, 3, 0, kH, 0);
    if (warp == 1) flush_bias(gB4, d_params + kB4, 0, 3);
  }
}

}  // namespace gf

using namespace gf;

int gf_launch_mlp_fwd_tc(int64_t n, const int32_t* d_n_ptr, const float* params, const void* feat_f16,
                         const int32_t* ray_id, const float* ray_bias, float* sigma, float* rgb, cudaStream_t st);

int gf_launch_mlp_bwd_tc(int64_t n, const int32_t* d_n_ptr, const float* params, const void* feat_f16,
                         const int32_t* ray_id, const float* ray_bias, const float* d_sigma, const float* d_rgb,
                         void* d_feat, float* d_params, float* d_ray_bias, float gscale, cudaStream_t st);

extern "C" {

int64_t gf_mlp_param_count(int hidden) { return hidden == kH ? (int64_t)kParamCount : -1; }

int gf_mlp_ray_bias(int64_t n_rays, int hidden, const float* params, const float* ray_dirs, const float* ray_emb,
                    float* ray_bias, void* stream) {
  GF_REQUIRE(hidden == kH, "gf_mlp_ray_bias: hidden width %d is not built (64 only)", hidden);
  GF_REQUIRE(n_rays >= 0, "gf_mlp_ray_bias: bad sizes");
  if (n_rays == 0) return GF_OK;
  GF_REQUIRE(params && ray_dirs && ray_bias, "gf_mlp_ray_bias: null pointer");
  ray_bias_kernel<<<(int)div_up(n_rays, kRbRays), 256, 0, (cudaStream_t)stream>>>(n_rays, params, ray_dirs, ray_emb,
                                                                                 ray_bias);
  return check_launch("ray_bias_kernel");
}

int gf_mlp_ray_bias_backward(int64_t n_rays, int hidden, const float* params, const float* ray_dirs,
                             const float* ray_emb, const float* d_ray_bias, float* d_params, float* d_ray_emb,
                             void* stream) {
  GF_REQUIRE(hidden == kH, "gf_mlp_ray_bias_backward: hidden width %d is not built (64 only)", hidden);
  GF_REQUIRE(n_rays >= 0, "gf_mlp_ray_bias_backward: bad sizes");
  if (n_rays == 0) return GF_OK;
  GF_REQUIRE(params && ray_dirs && d_ray_bias, "gf_mlp_ray_bias_backward: null pointer");
  GF_REQUIRE(!d_ray_emb || ray_emb, "gf_mlp_ray_bias_backward: d_ray_emb without ray_emb");
  ray_bias_bwd_kernel<<<(int)div_up(n_rays, kRbRays), 256, 0, (cudaStream_t)stream>>>(
      n_rays, params, ray_dirs, ray_emb, d_ray_bias, d_params, d_ray_emb);
  return check_launch("ray_bias_bwd_kernel");
}

int gf_mlp_forward(int64_t n, const int32_t* d_n_ptr, int hidden, const float* params, const void* feat_f16,
                   const int32_t* ray_id, const float* ray_bias, float* sigma, float* rgb, void* stream) {
  GF_REQUIRE(hidden == kH, "gf_mlp_forward: hidden width %d is not built (64 only)", hidden);
  GF_REQUIRE(n >= 0, "gf_mlp_forward: bad sizes");
  if (n == 0) return GF_OK;
  GF_REQUIRE(params && feat_f16 && ray_id && ray_bias && sigma && rgb, "gf_mlp_forward: null pointer");
  // GF_MLP_TC=0 selects the mma.sync kernel (profiling A/B only); default: tcgen05 / TMEM kernel (mlp_tc.cu)
  static const int use_tc = [] { const char* e = getenv("GF_MLP_TC"); return e ? atoi(e) : 1; }();
  if (use_tc)
    return gf_launch_mlp_fwd_tc(n, d_n_ptr, params, feat_f16, ray_id, ray_bias, sigma, rgb, (cudaStream_t)stream);
  constexpr int MT = 2;
  const int64_t tiles = div_up(n, 16 * MT);
  const int grid = (int)std::min<int64_t>(div_up(tiles, kFwdBlock / 32), (int64_t)sm_count() * 2);
  mlp_fwd_kernel<MT><<<grid, kFwdBlock, kWSmemBytes, (cudaStream_t)stream>>>(
      n, d_n_ptr, params, (const __half*)feat_f16, ray_id, ray_bias, sigma, rgb);
  return check_launch("mlp_fwd_kernel");
}

int gf_mlp_backward(int64_t n, const int32_t* d_n_ptr, int hidden, const float* params, const void* feat_f16,
                    const int32_t* ray_id, const float* ray_bias, const float* d_sigma, const float* d_rgb,
                    void* d_feat_scaled_f16, float* d_params, float* d_ray_bias, float grad_scale, void* stream) {
  GF_REQUIRE(hidden == kH, "gf_mlp_backward: hidden width %d is not built (64 only)", hidden);
  GF_REQUIRE(n >= 0 && grad_scale > 0.f, "gf_mlp_backward: bad sizes");
  if (n == 0) return GF_OK;
  GF_REQUIRE(params && feat_f16 && ray_id && ray_bias && d_sigma && d_rgb && d_feat_scaled_f16,
             "gf_mlp_backward: null pointer");
  GF_REQUIRE((d_params == nullptr) == (d_ray_bias == nullptr),
             "gf_mlp_backward: d_params and d_ray_bias go together (both NULL = frozen MLP)");
  static const int use_tc = [] { const char* e = getenv("GF_MLP_TC"); return e ? atoi(e) : 1; }();
  if (use_tc)
    return gf_launch_mlp_bwd_tc(n, d_n_ptr, params, feat_f16, ray_id, ray_bias, d_sigma, d_rgb, d_feat_scaled_f16,
                                d_params, d_ray_bias, grad_scale, (cudaStream_t)stream);
  static bool attr_set = false;
  if (!attr_set) {
    GF_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kBwdSmemBytes));
    GF_CUDA(cudaFuncSetAttribute(mlp_bwd_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWSmemBytes));
    attr_set = true;
  }
  const int64_t ctiles = div_up(n, kBwdPts);
  if (d_params) {
    const int grid = (int)std::min<int64_t>(ctiles, (int64_t)sm_count());
    mlp_bwd_kernel<true><<<grid, kBwdBlock, kBwdSmemBytes, (cudaStream_t)stream>>>(
        n, d_n_ptr, params, (const __half*)feat_f16, ray_id, ray_bias, d_sigma, d_rgb, (__half*)d_feat_scaled_f16,
        d_params, d_ray_bias, grad_scale);
  } else {
    const int grid = (int)std::min<int64_t>(ctiles, (int64_t)sm_count() * 2);
    mlp_bwd_kernel<false><<<grid, kBwdBlock, kWSmemBytes, (cudaStream_t)stream>>>(
        n, d_n_ptr, params, (const __half*)feat_f16, ray_id, ray_bias, d_sigma, d_rgb, (__half*)d_feat_scaled_f16,
        nullptr, nullptr, grad_scale);
  }
  return check_launch("mlp_bwd_kernel");
}

}  // extern "C"
