// Device helpers of the Hash3DAnchored encode shared by hash3d.cu (stand-alone gather / scatter) and mlp_tc.cu (the
// scatter fused into the MLP backward).  Arithmetic follows the oracle's FMA convention exactly (oracle/gf_oracle.c);
// line references are to the reference's field/Hash3DAnchored_cuda.cu.
#pragma once
#include "common.cuh"

namespace gf {

template <bool POW2>
__device__ __forceinline__ uint32_t wrap(uint32_t h, uint32_t local_size) {
  if (POW2) return h & (local_size - 1u);
  return h % local_size;
}

struct Cell {
  uint32_t px, py, pz;
  float a, b, c;
};

// :26-46  pt*mul + bias, floor, fractional part.  bias == nullptr: an all-zero bias pool (what the reference always
// has, Hash3DAnchored.cpp:57-62), same arithmetic without the three loads.
__device__ __forceinline__ Cell cell_of(float x, float y, float z, float mul, const float* __restrict__ bias) {
  float p0 = __fmaf_rn(x, mul, bias ? __ldg(bias + 0) : 0.f);
  float p1 = __fmaf_rn(y, mul, bias ? __ldg(bias + 1) : 0.f);
  float p2 = __fmaf_rn(z, mul, bias ? __ldg(bias + 2) : 0.f);
  float f0 = floorf(p0), f1 = floorf(p1), f2 = floorf(p2);
  Cell c;
  c.px = __float2uint_rz(f0);  // saturating, negative/NaN -> 0
  c.py = __float2uint_rz(f1);
  c.pz = __float2uint_rz(f2);
  c.a = __fsub_rn(p0, f0);
  c.b = __fsub_rn(p1, f1);
  c.c = __fsub_rn(p2, f2);
  return c;
}

// :48-55 corner rows, order 000,001,010,011,100,101,110,111 (bits x,y,z)
template <bool POW2>
__device__ __forceinline__ void corners(const Cell& c, uint32_t pa, uint32_t pb, uint32_t pc, uint32_t local_size,
                                        uint32_t (&pos)[8]) {
  uint32_t x0 = c.px * pa, x1 = (c.px + 1u) * pa;
  uint32_t y0 = c.py * pb, y1 = (c.py + 1u) * pb;
  uint32_t z0 = c.pz * pc, z1 = (c.pz + 1u) * pc;
  pos[0] = wrap<POW2>(x0 ^ y0 ^ z0, local_size);
  pos[1] = wrap<POW2>(x0 ^ y0 ^ z1, local_size);
  pos[2] = wrap<POW2>(x0 ^ y1 ^ z0, local_size);
  pos[3] = wrap<POW2>(x0 ^ y1 ^ z1, local_size);
  pos[4] = wrap<POW2>(x1 ^ y0 ^ z0, local_size);
  pos[5] = wrap<POW2>(x1 ^ y0 ^ z1, local_size);
  pos[6] = wrap<POW2>(x1 ^ y1 ^ z0, local_size);
  pos[7] = wrap<POW2>(x1 ^ y1 ^ z1, local_size);
}

// :58-69 trilinear weights, products left to right
__device__ __forceinline__ void weights(const Cell& c, float (&w)[8]) {
  float ia = __fsub_rn(1.f, c.a), ib = __fsub_rn(1.f, c.b), ic = __fsub_rn(1.f, c.c);
  float iaib = __fmul_rn(ia, ib), iab = __fmul_rn(ia, c.b), aib = __fmul_rn(c.a, ib), ab = __fmul_rn(c.a, c.b);
  w[0] = __fmul_rn(iaib, ic);
  w[1] = __fmul_rn(iaib, c.c);
  w[2] = __fmul_rn(iab, ic);
  w[3] = __fmul_rn(iab, c.c);
  w[4] = __fmul_rn(aib, ic);
  w[5] = __fmul_rn(aib, c.c);
  w[6] = __fmul_rn(ab, ic);
  w[7] = __fmul_rn(ab, c.c);
}

// fp16 round trip of a product, :148-151  (__half)(w0 * ws[d])
__device__ __forceinline__ float q16(float v) { return __half2float(__float2half_rn(v)); }

// One level of the backward scatter for the 32 consecutive samples held by a warp (:81-155): recompute cell / corner
// rows / weights, quantise the 8 x 2 contributions like the reference (fp16(g * 128) times w, each product rounded to
// fp16), sum runs of lanes that fall into the same cell of the same volume with a segmented shuffle reduction whose
// depth adapts to the longest run, and let the run's first lane issue the 8 vectorised fp32 reductions.
// g0 / g1: this sample's two gradients of the level, already fp16(g * 128) values.  Must be called by all 32 lanes.
template <bool POW2>
__device__ __forceinline__ void hash_scatter_level(int l, float x, float y, float z, int vol, bool valid, float g0,
                                                   float g1, int lane, int32_t n_volumes, uint32_t local_size,
                                                   const int32_t* __restrict__ prim_pool,
                                                   const float* __restrict__ bias_pool, float scale,
                                                   float* __restrict__ grad_table, bool aggregate) {
  const int tr = (l * n_volumes + vol) * 3;
  const Cell c = cell_of(x, y, z, scale, bias_pool ? bias_pool + tr : nullptr);
  const uint32_t pa = (uint32_t)__ldg(prim_pool + tr), pb = (uint32_t)__ldg(prim_pool + tr + 1),
                 pc = (uint32_t)__ldg(prim_pool + tr + 2);
  uint32_t pos[8];
  corners<POW2>(c, pa, pb, pc, local_size, pos);
  float w[8];
  weights(c, w);
  float c0[8], c1[8];
#pragma unroll
  for (int d = 0; d < 8; d++) {
    c0[d] = q16(__fmul_rn(g0, w[d]));
    c1[d] = q16(__fmul_rn(g1, w[d]));
  }
  bool writer = valid && (g0 != 0.f || g1 != 0.f);  // :147 skip when both are zero
  if (aggregate) {
    // runs of lanes in the same cell of the same volume -> one reduction per run
    const uint32_t ppx = __shfl_up_sync(0xffffffffu, c.px, 1), ppy = __shfl_up_sync(0xffffffffu, c.py, 1),
                   ppz = __shfl_up_sync(0xffffffffu, c.pz, 1);
    const int pvol = __shfl_up_sync(0xffffffffu, vol, 1);
    const bool head = lane == 0 || ppx != c.px || ppy != c.py || ppz != c.pz || pvol != vol;
    const uint32_t heads = __ballot_sync(0xffffffffu, head);
    const uint32_t above = lane == 31 ? 0u : (heads & (0xffffffffu << (lane + 1)));
    const int end = above ? (__ffs(above) - 1) : 32;
    // longest run in the warp: the segmented reduction needs ceil(log2) of it steps -- 5 on the coarse levels
    // (the whole warp in one cell), 0..2 on the fine ones, where most of the samples are
    const int maxrun = (int)__reduce_max_sync(0xffffffffu, head ? (unsigned)(end - lane) : 0u);
    if (maxrun > 1) {  // warp-uniform
      if (!valid) {
#pragma unroll
        for (int d = 0; d < 8; d++) c0[d] = c1[d] = 0.f;
      }
      for (int off = 1; off < maxrun; off <<= 1) {
        const bool take = lane + off < end;
#pragma unroll
        for (int d = 0; d < 8; d++) {
          float o0 = __shfl_down_sync(0xffffffffu, c0[d], off);
          float o1 = __shfl_down_sync(0xffffffffu, c1[d], off);
          if (take) {
            c0[d] += o0;
            c1[d] += o1;
          }
        }
      }
      writer = head && valid;
    }
  }
  if (writer) {
    float2* tab = reinterpret_cast<float2*>(grad_table) + (int64_t)l * local_size;
#pragma unroll
    for (int d = 0; d < 8; d++) {
      const float s0 = c0[d] * (1.f / GF_GRAD_SCALE), s1 = c1[d] * (1.f / GF_GRAD_SCALE);
      if (s0 != 0.f || s1 != 0.f) atomicAdd(tab + pos[d], make_float2(s0, s1));
    }
  }
}

}  // namespace gf
