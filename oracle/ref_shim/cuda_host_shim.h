/* Host stand-ins for the CUDA language features the reference's device code uses, so that the function bodies that
 * oracle/ref_extract.py pulls out of /root/reference compile unmodified with g++ and run one "thread" at a time.
 * TEST INFRASTRUCTURE ONLY (oracle/).  None of the reference's kernels uses shared memory, __syncthreads or warp
 * intrinsics, so a serial loop over (blockIdx, threadIdx) is a faithful execution; atomics become plain
 * read-modify-writes (the emulated launch is single-threaded). */
#ifndef GF_REF_CUDA_HOST_SHIM_H
#define GF_REF_CUDA_HOST_SHIM_H
#include <math.h>
#include <stdint.h>
#include <string.h>

#include <algorithm>
#include <cmath>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline

struct gf_dim3 {
  unsigned x = 1, y = 1, z = 1;
};
static thread_local gf_dim3 threadIdx, blockIdx, blockDim, gridDim;

using std::max;
using std::min;

/* ---- fp16: _Float16 gives IEEE binary16 with round-to-nearest-even conversions, like __float2half_rn ---- */
struct __half {
  _Float16 v;
  __half() = default;
  __half(float f) : v((_Float16)f) {}
  explicit __half(double f) : v((_Float16)(float)f) {}
  operator float() const { return (float)v; }
};
struct __half2 {
  __half x, y;
};

/* atomicAdd(__half2*): two independent fp16 additions, each rounded to fp16 (what HADD2 / the red.f16x2 unit do) */
static inline __half2 atomicAdd(__half2* addr, __half2 val) {
  __half2 old = *addr;
  addr->x = __half((float)old.x + (float)val.x);
  addr->y = __half((float)old.y + (float)val.y);
  return old;
}
static inline unsigned long long atomicAdd(unsigned long long* addr, unsigned long long val) {
  unsigned long long old = *addr;
  *addr = old + val;
  return old;
}
static inline float atomicAdd(float* addr, float val) {
  float old = *addr;
  *addr = old + val;
  return old;
}
static inline long long atomicMax(long long* addr, long long val) {
  long long old = *addr;
  if (val > old) *addr = val;
  return old;
}

/* launch emulation: for every block and thread of a 1-D block / 2-D grid, set the built-ins and call f() */
template <typename F>
static inline void gf_launch(unsigned grid_x, unsigned grid_y, unsigned block_x, F&& f) {
  gridDim.x = grid_x;
  gridDim.y = grid_y;
  blockDim.x = block_x;
  for (unsigned by = 0; by < grid_y; by++)
    for (unsigned bx = 0; bx < grid_x; bx++)
      for (unsigned tx = 0; tx < block_x; tx++) {
        blockIdx.x = bx;
        blockIdx.y = by;
        threadIdx.x = tx;
        f();
      }
}
#endif
