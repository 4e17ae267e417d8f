// Helpers shared by the tcgen05 field-MLP kernels (mlp_tc.cu: H = 64, mlp_tc128.cu: H = 128): fp16 packing, the
// ReLU-mask bit layout, the no-swizzle core-matrix tile addressing of weights and [sample][feature] tiles.
#pragma once
#include "common.cuh"
#include "tcgen05.cuh"

namespace gf {
namespace tc {

// B operands: [N rows][K] fp16, K-major, no swizzle: element (n, k) at
//   (n / 8) * SBO + (k / 8) * 128 + (n % 8) * 16 + (k % 8) * 2,   SBO = (K / 8) * 128
constexpr uint32_t kLbo = 128;
constexpr uint32_t sbo_of(int K) { return (uint32_t)(K / 8) * 128u; }

__device__ __forceinline__ uint32_t pack_h2(float lo, float hi) {
  __half2 h = __floats2half2_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t relu_h2(uint32_t v) {
  __half2 h = __hmax2(*reinterpret_cast<__half2*>(&v), __float2half2_rn(0.f));
  return *reinterpret_cast<uint32_t*>(&h);
}
// two fp32 -> packed fp16 pair with ReLU in the conversion (cvt.rn.relu.f16x2.f32: first source = upper half)
__device__ __forceinline__ uint32_t pack_relu_h2(float lo, float hi) {
  uint32_t d;
  asm("cvt.rn.relu.f16x2.f32 %0, %1, %2;\n" : "=r"(d) : "f"(hi), "f"(lo));
  return d;
}
// fp32 -> fp16 hi + fp16 lo: hi = the value truncated to 11 significant bits (exact in fp16), lo = the rest
__device__ __forceinline__ float trunc11(float v) { return __uint_as_float(__float_as_uint(v) & 0xFFFFE000u); }
// bits of the ReLU-mask word for the pair (2q, 2q+1) of a thread's 32 columns: chosen so that the backward expands
// them to a half2 AND-mask with one shift + one PRMT in sign-replicating mode (mask_bits_pack32)
__host__ __device__ constexpr uint32_t mask_bits_of_pair(int q) {
  return q < 8 ? ((1u << (7 - q)) | (1u << (23 - q))) : ((1u << (15 - (q - 8))) | (1u << (31 - (q - 8))));
}
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

__device__ __forceinline__ void put_b(unsigned char* smem, uint32_t off, int K, int n, int k, float v) {
  *reinterpret_cast<__half*>(smem + off + (n >> 3) * sbo_of(K) + (k >> 3) * 128 + (n & 7) * 16 + (k & 7) * 2) =
      __float2half_rn(v);
}


// 16-byte chunk j (features 8 j .. 8 j + 7) of row r of a [sample][feature] tile
__device__ __forceinline__ uint4* tile_chunk(unsigned char* buf, uint32_t sbo, int r, int j) {
  return reinterpret_cast<uint4*>(buf + (r >> 3) * sbo + j * 128 + (r & 7) * 16);
}
template <int NCH>
__device__ __forceinline__ void store_chunks(unsigned char* buf, uint32_t sbo, int r, int j0, const uint32_t* a) {
#pragma unroll
  for (int q = 0; q < NCH; q++)
    *tile_chunk(buf, sbo, r, j0 + q) = make_uint4(a[4 * q], a[4 * q + 1], a[4 * q + 2], a[4 * q + 3]);
}
__device__ __forceinline__ void store_ones(unsigned char* buf, uint32_t sbo, int r, int j) {
  *tile_chunk(buf, sbo, r, j) = make_uint4(0x00003C00u, 0u, 0u, 0u);  // feature 8 j = 1.0, the rest 0
}


// gradient row masked by the forward's ReLU-mask word -> 16 packed pairs.  Pair q < 8: its two bits sit at 7 - q and
// 23 - q, so after a left shift by q they are the sign bits of bytes 0 and 2, which PRMT (selector nibble | 8 =
// replicate the byte's sign) spreads over the two halves; pairs 8..15 use bytes 1 and 3.
__device__ __forceinline__ void mask_bits_pack32(const uint32_t (&v)[32], uint32_t mask, uint32_t (&out)[16]) {
#pragma unroll
  for (int q = 0; q < 16; q++) {
    uint32_t sel;   // (__byte_perm only takes 3-bit selectors; the sign-replicate bit needs the PTX form)
    asm("prmt.b32 %0, %1, %2, %3;\n" : "=r"(sel) : "r"(mask << (q & 7)), "r"(0u), "r"(q < 8 ? 0xAA88u : 0xBB99u));
    out[q] = pack_h2(__uint_as_float(v[2 * q]), __uint_as_float(v[2 * q + 1])) & sel;
  }
}

}  // namespace tc
}  // namespace gf
