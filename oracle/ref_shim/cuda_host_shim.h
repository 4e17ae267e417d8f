/* Host stand-ins for the CUDA language features the reference's device code uses, so that the function bodies that
 * oracle/ref_extract.py pulls out of /root/reference compile unmodified with g++ and run one "thread" at a time.
 * TEST INFRASTRUCTURE ONLY (oracle/).  None of the reference's kernels uses shared memory, __syncthreads or warp
 * intrinsics, so a loop over (blockIdx, threadIdx) in any order is a faithful execution. */
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

/* Atomics.  A launch normally runs one "thread" after the other (below), where these are plain read-modify-writes in
 * a fixed order; with ref_set_threads(n > 1) the thread instances of a launch are spread over n host threads (the CPU
 * arm of bench.py: the reference's kernels on all host cores), so they are real atomics: compare-and-swap loops for
 * the types without a hardware add.  Relaxed ordering, like the device's. */
/* atomicAdd(__half2*): two independent fp16 additions, each rounded to fp16 (what HADD2 / the red.f16x2 unit do) */
static inline __half2 atomicAdd(__half2* addr, __half2 val) {
  static_assert(sizeof(__half2) == 4, "__half2 is one 32-bit word");
  uint32_t* w = reinterpret_cast<uint32_t*>(addr);
  uint32_t seen = __atomic_load_n(w, __ATOMIC_RELAXED), want;
  __half2 old, upd;
  do {
    memcpy(&old, &seen, 4);
    upd.x = __half((float)old.x + (float)val.x);
    upd.y = __half((float)old.y + (float)val.y);
    memcpy(&want, &upd, 4);
  } while (!__atomic_compare_exchange_n(w, &seen, want, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return old;
}
static inline unsigned long long atomicAdd(unsigned long long* addr, unsigned long long val) {
  return __atomic_fetch_add(addr, val, __ATOMIC_RELAXED);
}
static inline float atomicAdd(float* addr, float val) {
  uint32_t* w = reinterpret_cast<uint32_t*>(addr);
  uint32_t seen = __atomic_load_n(w, __ATOMIC_RELAXED), want;
  float old, upd;
  do {
    memcpy(&old, &seen, 4);
    upd = old + val;
    memcpy(&want, &upd, 4);
  } while (!__atomic_compare_exchange_n(w, &seen, want, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED));
  return old;
}
static inline long long atomicMax(long long* addr, long long val) {
  long long old = __atomic_load_n(addr, __ATOMIC_RELAXED);
  while (val > old && !__atomic_compare_exchange_n(addr, &old, val, true, __ATOMIC_RELAXED, __ATOMIC_RELAXED)) {
  }
  return old;
}

/* launch emulation: for every block and thread of a 1-D block / 2-D grid, set the built-ins and call f().
 * gf_ref_threads <= 1 (the default, what every test uses): one instance after the other, blocks in launch order --
 * deterministic.  > 1: the instances are distributed over that many OpenMP threads (build with -fopenmp; without it
 * the pragma is ignored and the loop stays serial); none of the kernels synchronises within a block, so any
 * interleaving is one the GPU could produce. */
static int gf_ref_threads = 1;
template <typename F>
static inline void gf_launch(unsigned grid_x, unsigned grid_y, unsigned block_x, F&& f) {
  if (gf_ref_threads <= 1) {
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
    return;
  }
  const long long n = (long long)grid_x * grid_y * block_x;
#pragma omp parallel for schedule(dynamic, 32) num_threads(gf_ref_threads)
  for (long long i = 0; i < n; i++) {
    gridDim.x = grid_x;
    gridDim.y = grid_y;
    blockDim.x = block_x;
    const long long b = i / block_x;
    threadIdx.x = (unsigned)(i - b * block_x);
    blockIdx.x = (unsigned)(b % grid_x);
    blockIdx.y = (unsigned)(b / grid_x);
    f();
  }
}
#endif
