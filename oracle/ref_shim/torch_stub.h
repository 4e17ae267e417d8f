/* A stand-in for the sliver of torch that the reference's PersOctree::ProcOctree / ConstructEdgePool touch in their
 * prologue and epilogue (PtsSampler/PersSampler.cpp:156-171, 405-415): byte buffers that are "moved" between
 * devices, viewed as typed pointers, and created from a blob.  TEST INFRASTRUCTURE ONLY (oracle/); everything lives
 * on the host.  Plus the logging / CHECK macros of Utils/Common.h the bodies use. */
#ifndef GF_REF_TORCH_STUB_H
#define GF_REF_TORCH_STUB_H
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>

#include <cstring>
#include <functional>
#include <initializer_list>
#include <memory>
#include <string>
#include <vector>

namespace torch {
enum DeviceStub { kCPU, kCUDA };
struct OptStub {
  int elem_size;
};
struct Tensor {
  std::shared_ptr<std::vector<uint8_t>> buf;
  int elem_size = 1;
  Tensor to(DeviceStub) const { /* a device copy: new storage, same bytes */
    Tensor t;
    t.elem_size = elem_size;
    t.buf = buf ? std::make_shared<std::vector<uint8_t>>(*buf) : nullptr;
    return t;
  }
  Tensor contiguous() const { return *this; }
  void* data_ptr() const { return buf ? (void*)buf->data() : nullptr; }
  template <typename T>
  T* data_ptr() const {
    return (T*)data_ptr();
  }
  int64_t size(int) const { return buf ? (int64_t)(buf->size() / elem_size) : 0; }
};
inline Tensor from_blob(void* p, std::initializer_list<int64_t> shape, OptStub o) {
  Tensor t;
  t.elem_size = o.elem_size;
  int64_t n = 1;
  for (int64_t s : shape) n *= s;
  t.buf = std::make_shared<std::vector<uint8_t>>((size_t)(n * o.elem_size));
  if (n) std::memcpy(t.buf->data(), p, (size_t)(n * o.elem_size));
  return t;
}
inline Tensor zeros(std::initializer_list<int64_t> shape, OptStub o) {
  Tensor t;
  t.elem_size = o.elem_size;
  int64_t n = 1;
  for (int64_t s : shape) n *= s;
  t.buf = std::make_shared<std::vector<uint8_t>>((size_t)(n * o.elem_size), (uint8_t)0);
  return t;
}
}  // namespace torch

#define CPUUInt8 torch::OptStub{1}
#define CPUInt64 torch::OptStub{8}
#define CUDAInt64 torch::OptStub{8}

#define RE_INTER(x, y) reinterpret_cast<x>(y)
#define PRINT_VAL(x) do { } while (false)
struct RefCheckFailure {
  const char* what;
  int line;
};
#define CHECK(c) do { if (!(c)) throw RefCheckFailure{#c, __LINE__}; } while (false)
#define CHECK_EQ(a, b) CHECK((a) == (int64_t)(b))
#define CHECK_LE(a, b) CHECK((a) <= (b))
#define CHECK_LT(a, b) CHECK((a) < (b))
#define CHECK_GE(a, b) CHECK((a) >= (b))
#define CHECK_GT(a, b) CHECK((a) > (b))
struct ScopeWatch {
  explicit ScopeWatch(const char*) {}
};
#endif
