// PersOctree maintenance ON THE DEVICE: ProcOctree (compact + subdivide) over the reference's 128-byte TreeNode blob
// without the node blob ever leaving HBM.
//
// The reference's PersOctree::ProcOctree (gfnerf/bindings/PtsSampler/PersSampler.cpp:154-417) copies the node blob and
// the vote statistics to the host, rebuilds the tree there and uploads it again -- every compact_freq (1000) steps
// and twice at each of the five subdivision milestones; gf_octree_proc (octree_host.cu) restates that host code and
// stays as the byte-level checker.  This kernel produces the SAME bytes (tests/test_octree_device_gpu.py: blobs and
// statistics identical to gf_octree_proc and to the reference's own ProcOctree fixtures) with one single-CTA launch;
// the only thing the host reads back is the new node count (8 bytes), when it wants it.
//
// The host algorithm is sequential; each phase is restated in an order-independent form:
//  A  pruning to a fixed point (:170-205): "unlink invalid leaves from their parents" and "interior nodes without
//     children become leaves" are each a parallel sweep with a barrier in between, repeated while anything changed;
//  B  chain splicing (:207-240): the sequential walk removes exactly the non-root interior nodes with a single child
//     (set D) and hangs the node below each maximal D-chain under the first non-D ancestor, whatever the visiting
//     order -- so: flag D, then every live non-D node whose parent is in D walks up to its new parent, then D dies;
//  C  renumbering of the survivors in their old order (:242-316): a block-wide exclusive scan of the survivor flags;
//  D  subdivision (:318-417): the host recursion numbers nodes depth first, a split leaf directly followed by its
//     eight children.  That numbering is a preorder rank: subtree sizes bottom-up level by level (a split leaf weighs
//     9), then ranks top-down level by level (rank of a child = rank of the parent + 1 + sizes of its earlier
//     siblings); the tree is <= 16 + 10 levels deep, so this is ~50 block barriers.
// One CTA of 1024 threads: the tree has 10^3..10^5 nodes and the work is a few passes over 128-byte records.
#include "common.cuh"

namespace gf {
namespace {

struct TreeNodeDev {   // 128 bytes, the reference's TreeNode (PersSampler.h:31-49)
  float center[3];     // @0
  float side_len;      // @12
  int64_t parent;      // @16
  int64_t childs[8];   // @24
  uint8_t is_leaf;     // @88
  uint8_t pad0[7];
  int64_t trans_idx;   // @96
  int64_t block_idx;   // @104
  uint8_t pad1[16];
};
static_assert(sizeof(TreeNodeDev) == GF_TREE_NODE_BYTES, "TreeNode blob layout");

constexpr int64_t kInitNodeStatDev = 1000;   // INIT_NODE_STAT, PersSampler_cuda.cu:11-17
constexpr int kProcThreads = 1024;

__device__ __forceinline__ void copy_node(TreeNodeDev* dst, const TreeNodeDev* src) {
  const uint4* s = reinterpret_cast<const uint4*>(src);
  uint4* d = reinterpret_cast<uint4*>(dst);
#pragma unroll
  for (int k = 0; k < 8; k++) d[k] = s[k];
}
__device__ __forceinline__ int child_count(const TreeNodeDev& x) {
  int c = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) c += x.childs[k] >= 0;
  return c;
}

// exclusive scan of flag[0..n) into out[0..n), total returned to every thread.  Thread t owns a contiguous chunk.
__device__ int64_t block_exclusive_scan(const int32_t* flag, int32_t* out, int64_t n, int32_t* s_part) {
  const int tid = threadIdx.x, T = blockDim.x;
  const int64_t chunk = (n + T - 1) / T, lo = min((int64_t)tid * chunk, n), hi = min(lo + chunk, n);
  int32_t sum = 0;
  for (int64_t i = lo; i < hi; i++) sum += flag[i];
  s_part[tid] = sum;
  __syncthreads();
  // Hillis-Steele over the T partial sums
  for (int off = 1; off < T; off <<= 1) {
    const int32_t v = tid >= off ? s_part[tid - off] : 0;
    __syncthreads();
    s_part[tid] += v;
    __syncthreads();
  }
  int32_t run = s_part[tid] - sum;   // exclusive prefix of this thread's chunk
  const int64_t total = s_part[T - 1];
  for (int64_t i = lo; i < hi; i++) {
    const int32_t f = flag[i];
    out[i] = run;
    run += f;
  }
  __syncthreads();
  return total;
}

// scratch layout (gf_octree_proc_device_scratch_bytes): nb [n] nodes | cn [n] nodes | cw, ca [n] i64 | 7 x i32 [n]
__global__ void __launch_bounds__(kProcThreads)
octree_proc_kernel(const TreeNodeDev* __restrict__ in, int64_t n, const int64_t* __restrict__ w_in,
                   const int64_t* __restrict__ a_in, const int64_t* __restrict__ visit, int compact, int subdivide,
                   int brute_force, TreeNodeDev* __restrict__ out, int64_t* __restrict__ w_out,
                   int64_t* __restrict__ a_out, int64_t capacity, unsigned char* __restrict__ scratch,
                   int64_t* __restrict__ d_n_out, int32_t* __restrict__ d_error) {
  __shared__ int32_t s_part[kProcThreads];
  __shared__ int s_flag, s_err, s_maxdepth;
  const int tid = threadIdx.x, T = blockDim.x;
  TreeNodeDev* nb = reinterpret_cast<TreeNodeDev*>(scratch);
  TreeNodeDev* cn = nb + n;
  int64_t* cw = reinterpret_cast<int64_t*>(cn + n);
  int64_t* ca = cw + n;
  int32_t* flag = reinterpret_cast<int32_t*>(ca + n);   // survivor flag, later: split flag
  int32_t* new_idx = flag + n;
  int32_t* inv = new_idx + n;
  int32_t* dflag = inv + n;                              // phase B: in D; phase D: depth
  int32_t* size = dflag + n;
  int32_t* rank = size + n;
  if (tid == 0) {
    s_err = 0;
    s_maxdepth = 0;
  }
  for (int64_t u = tid; u < n; u += T) copy_node(nb + u, in + u);
  __syncthreads();

  if (compact) {
    // ---- A: prune to a fixed point (:170-205) ----
    while (true) {
      for (int64_t u = tid; u < n; u += T) {
        const TreeNodeDev& x = nb[u];
        if (x.is_leaf && x.trans_idx < 0 && x.parent >= 0) {
          TreeNodeDev& p = nb[x.parent];
#pragma unroll
          for (int k = 0; k < 8; k++)
            if (p.childs[k] == u) p.childs[k] = -1;
        }
      }
      if (tid == 0) s_flag = 0;
      __syncthreads();
      for (int64_t u = 1 + tid; u < n; u += T) {
        if (child_count(nb[u]) == 0) {
          if (!nb[u].is_leaf) s_flag = 1;
          nb[u].is_leaf = 1;
        }
      }
      __syncthreads();
      const int again = s_flag;
      __syncthreads();
      if (!again) break;
    }
    // ---- B: splice out chains of single-child interior nodes (:207-240) ----
    for (int64_t u = tid; u < n; u += T) dflag[u] = (nb[u].parent >= 0 && child_count(nb[u]) == 1) ? 1 : 0;
    __syncthreads();
    for (int64_t u = tid; u < n; u += T) {
      TreeNodeDev& x = nb[u];
      if ((x.is_leaf && x.trans_idx < 0) || dflag[u]) continue;
      int64_t v = x.parent;
      if (v < 0 || !dflag[v]) continue;
      int64_t top = v;
      while (dflag[v]) {   // the root is never in D, so this ends
        top = v;
        v = nb[v].parent;
      }
      TreeNodeDev& vv = nb[v];
#pragma unroll
      for (int k = 0; k < 8; k++)
        if (vv.childs[k] == top) vv.childs[k] = u;
      x.parent = v;
    }
    __syncthreads();
    for (int64_t u = tid; u < n; u += T)
      if (dflag[u]) {
        nb[u].trans_idx = -1;
        nb[u].is_leaf = 1;
      }
    __syncthreads();
  }

  // ---- C: renumber the survivors (interior nodes and valid leaves) in their old order (:242-316) ----
  for (int64_t u = tid; u < n; u += T) flag[u] = (!nb[u].is_leaf || nb[u].trans_idx >= 0) ? 1 : 0;
  __syncthreads();
  const int64_t m = block_exclusive_scan(flag, new_idx, n, s_part);
  if (flag[0] == 0) {   // the root was pruned: no valid leaf left
    if (tid == 0) {
      *d_n_out = 0;
      atomicOr(d_error, 1);
    }
    return;
  }
  for (int64_t u = tid; u < n; u += T) {
    if (!flag[u]) continue;
    const int64_t i = new_idx[u];
    copy_node(cn + i, nb + u);   // byte copy: padding travels with the node
    TreeNodeDev& x = cn[i];
    if (x.parent >= 0) {
      if (!flag[x.parent]) s_err = 2;   // CHECK_GE(node.parent, 0), PersSampler.cpp:307
      x.parent = flag[x.parent] ? new_idx[x.parent] : -1;
    }
#pragma unroll
    for (int k = 0; k < 8; k++)
      if (x.childs[k] >= 0) {
        if (!flag[x.childs[k]]) s_err = 2;   // CHECK_GE(node.childs[st], 0), :315
        x.childs[k] = flag[x.childs[k]] ? new_idx[x.childs[k]] : -1;
      }
    cw[i] = w_in[u];
    ca[i] = a_in[u];
    inv[i] = (int32_t)u;
  }
  __syncthreads();
  if (s_err) {   // a removed node is still linked (compact = 0 on a pruned tree): the reference aborts here
    if (tid == 0) {
      *d_n_out = 0;
      atomicOr(d_error, s_err);
    }
    return;
  }
  if (!subdivide) {
    if (m > capacity) {
      if (tid == 0) {
        *d_n_out = m;
        atomicOr(d_error, 4);
      }
      return;
    }
    for (int64_t i = tid; i < m; i += T) {
      copy_node(out + i, cn + i);
      w_out[i] = cw[i];
      a_out[i] = ca[i];
    }
    if (tid == 0) *d_n_out = m;
    return;
  }

  // ---- D: subdivision (:318-417) as a preorder ranking ----
  int32_t* split = flag;
  int32_t* depth = dflag;
  for (int64_t i = tid; i < m; i += T) {
    split[i] = (cn[i].is_leaf && (brute_force || visit[inv[i]] > 4)) ? 1 : 0;   // :354
    int d = 0;
    for (int64_t p = cn[i].parent; p >= 0; p = cn[p].parent) d++;
    depth[i] = d;
    atomicMax(&s_maxdepth, d);
    size[i] = split[i] ? 9 : 1;
  }
  __syncthreads();
  const int maxdepth = s_maxdepth;
  for (int d = maxdepth - 1; d >= 0; d--) {   // subtree sizes, deepest parents first
    for (int64_t i = tid; i < m; i += T) {
      if (depth[i] != d || cn[i].is_leaf) continue;
      int32_t s = 1;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int64_t c = cn[i].childs[k];
        if (c >= 0) s += size[c];
      }
      size[i] = s;
    }
    __syncthreads();
  }
  if (tid == 0) rank[0] = 0;
  __syncthreads();
  for (int d = 0; d < maxdepth; d++) {        // preorder ranks, shallowest parents first
    for (int64_t i = tid; i < m; i += T) {
      if (depth[i] != d || cn[i].is_leaf) continue;
      int32_t run = rank[i] + 1;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const int64_t c = cn[i].childs[k];
        if (c >= 0) {
          rank[c] = run;
          run += size[c];
        }
      }
    }
    __syncthreads();
  }
  const int64_t total = size[0];
  if (total > capacity) {
    if (tid == 0) {
      *d_n_out = total;
      atomicOr(d_error, 4);
    }
    return;
  }
  for (int64_t i = tid; i < m; i += T) {
    const TreeNodeDev& s = cn[i];
    const int64_t nu = rank[i];
    TreeNodeDev node;
    {
      uint4* z = reinterpret_cast<uint4*>(&node);
#pragma unroll
      for (int k = 0; k < 8; k++) z[k] = make_uint4(0u, 0u, 0u, 0u);   // the host rebuilds nodes from zeroed memory
    }
    node.center[0] = s.center[0];
    node.center[1] = s.center[1];
    node.center[2] = s.center[2];
    node.side_len = s.side_len;
    node.parent = s.parent >= 0 ? (int64_t)rank[s.parent] : -1;
    node.is_leaf = s.is_leaf;
    node.trans_idx = s.trans_idx;
    node.block_idx = s.block_idx;
#pragma unroll
    for (int k = 0; k < 8; k++) node.childs[k] = (!s.is_leaf && s.childs[k] >= 0) ? (int64_t)rank[s.childs[k]] : s.childs[k];
    int64_t sw = cw[i], sa = ca[i];
    if (split[i]) {
      const float half = __fmul_rn(s.side_len, .5f);
#pragma unroll 1
      for (int st = 0; st < 8; st++) {
        TreeNodeDev ch;
        uint4* z = reinterpret_cast<uint4*>(&ch);
#pragma unroll
        for (int k = 0; k < 8; k++) z[k] = make_uint4(0u, 0u, 0u, 0u);
        ch.center[0] = __fadd_rn(s.center[0], __fmul_rn(half, float((st >> 2) & 1) - .5f));
        ch.center[1] = __fadd_rn(s.center[1], __fmul_rn(half, float((st >> 1) & 1) - .5f));
        ch.center[2] = __fadd_rn(s.center[2], __fmul_rn(half, float(st & 1) - .5f));
        ch.side_len = half;
        ch.parent = nu;
#pragma unroll
        for (int k = 0; k < 8; k++) ch.childs[k] = -1;
        ch.is_leaf = 1;
        ch.trans_idx = s.trans_idx;
        ch.block_idx = 0;   // left unset by the reference (:376)
        copy_node(out + nu + 1 + st, &ch);
        w_out[nu + 1 + st] = sw;
        a_out[nu + 1 + st] = sa;
        node.childs[st] = nu + 1 + st;
      }
      node.is_leaf = 0;
      node.trans_idx = -1;
      sw = sa = kInitNodeStatDev;
    }
    copy_node(out + nu, &node);
    w_out[nu] = sw;
    a_out[nu] = sa;
  }
  if (tid == 0) *d_n_out = total;
}


// ---- MarkInvisibleNodesKernel + CheckVisible (PersSampler_cuda.cu:680-742) -------------------------------------------
// A node no camera can see loses its transform (trans_idx = -1); run at the subdivision milestones between the two
// ProcOctree calls (:657-677).  One thread per node; the cameras go through shared memory in tiles of kCamTile records
// {w2c 3x4, fx, cx, fy, cy, bound0, bound1} read by the whole CTA, so a camera is fetched from HBM once per CTA and
// not once per node; a thread stops testing at its first visible camera (the reference counts them all and compares
// the count with 1), a CTA stops staging when all its nodes are settled.
//
// Arithmetic = the SASS nvcc 12.9 emits for the reference's kernel body at -fmad=true (oracle/_ref/libgf_ref_cuda.so;
// tests/test_octree_device_gpu.py runs the two side by side and asserts identical blobs):
//   radius  = float(double(side_len) * 0.707)            (the literal is a double)
//   cam_i   = fma(c.x, m_i0, c.y * m_i1) + fma(c.z, m_i2, m_i3)
//   norm    = sqrt(fma(x, x, fma(y, y, z * z)))
//   q = radius / -z;  img_x = fx * (x / -z);  img_y = fy * (y / -z);  bias_y = fy * q
//   out of image <=> img_x > fma(fx, q, cx)  ||  fma(fx, q, img_x) < -cx  ||  bias_y + img_y < -cy  ||  img_y > cy + bias_y
// (bias_x never exists as a rounded value: its product is fused into both sums; bias_y does.)
constexpr int kCamTile = 128;
constexpr int kCamRec = 18;
constexpr int kVisThreads = 256;

__device__ __forceinline__ bool camera_sees(const float* __restrict__ c, float px, float py, float pz, float radius) {
  const float x = __fadd_rn(__fmaf_rn(px, c[0], __fmul_rn(py, c[1])), __fmaf_rn(pz, c[2], c[3]));
  const float y = __fadd_rn(__fmaf_rn(px, c[4], __fmul_rn(py, c[5])), __fmaf_rn(pz, c[6], c[7]));
  const float z = __fadd_rn(__fmaf_rn(px, c[8], __fmul_rn(py, c[9])), __fmaf_rn(pz, c[10], c[11]));
  const float fx = c[12], cx = c[13], fy = c[14], cy = c[15], b0 = c[16], b1 = c[17];
  const float nz = -z;
  if (nz < __fadd_rn(b0, -radius) || nz > __fadd_rn(radius, b1)) return false;
  const float norm = __fsqrt_rn(__fmaf_rn(x, x, __fmaf_rn(y, y, __fmul_rn(z, z))));
  if (norm < radius) return true;
  const float q = __fdiv_rn(radius, nz);
  const float bias_y = __fmul_rn(fy, q);
  const float img_x = __fmul_rn(fx, __fdiv_rn(x, nz));
  const float img_y = __fmul_rn(fy, __fdiv_rn(y, nz));
  if (img_x > __fmaf_rn(fx, q, cx) || __fmaf_rn(fx, q, img_x) < -cx) return false;
  if (__fadd_rn(bias_y, img_y) < -cy || img_y > __fadd_rn(cy, bias_y)) return false;
  return true;
}

__global__ void __launch_bounds__(kVisThreads)
mark_invisible_kernel(TreeNodeDev* __restrict__ nodes, int64_t n_nodes, const float* __restrict__ w2c,
                      const float* __restrict__ intri, const float* __restrict__ bounds, int64_t n_cams) {
  __shared__ float cams[kCamTile * kCamRec];
  const int64_t n_round = (n_nodes + kVisThreads - 1) / kVisThreads * kVisThreads;   // whole CTAs stay in the loop
  for (int64_t node = (int64_t)blockIdx.x * kVisThreads + threadIdx.x; node < n_round;
       node += (int64_t)gridDim.x * kVisThreads) {
    const bool live = node < n_nodes;
    float px = 0.f, py = 0.f, pz = 0.f, radius = 0.f;
    if (live) {
      const float4 cs = *reinterpret_cast<const float4*>(nodes + node);   // center xyz, side_len
      px = cs.x, py = cs.y, pz = cs.z;
      radius = (float)((double)cs.w * 0.707);
    }
    bool seen = !live;
    for (int64_t c0 = 0; c0 < n_cams; c0 += kCamTile) {
      const int nc = (int)min((int64_t)kCamTile, n_cams - c0);
      __syncthreads();
      for (int i = threadIdx.x; i < nc * kCamRec; i += kVisThreads) {
        const int cam = i / kCamRec, f = i - cam * kCamRec;
        const int64_t g = c0 + cam;
        float v;
        if (f < 12) v = w2c[g * 12 + f];
        else if (f < 16) v = intri[g * 9 + (f == 12 ? 0 : f == 13 ? 2 : f == 14 ? 4 : 5)];
        else v = bounds[g * 2 + (f - 16)];
        cams[i] = v;
      }
      __syncthreads();
      for (int cam = 0; cam < nc && !seen; cam++) seen = camera_sees(cams + cam * kCamRec, px, py, pz, radius);
      if (__syncthreads_and(seen)) break;
    }
    if (live && !seen) nodes[node].trans_idx = -1;
  }
}

// ---- SetBlockIdxsNearestKernel (PersSampler_cuda.cu:746-766) ---------------------------------------------------------
// block_idx of every node = index of the nearest block centre: fp32 difference and norm (fma(dx, dx, fma(dy, dy,
// dz * dz)), IEEE sqrt -- the reference's SASS), compared as doubles against a running minimum that starts at 1e9 with
// a strict `<` (the first of equal minima wins; -1 if no centre is closer than 1e9).
constexpr int kCenterTile = 512;

__global__ void __launch_bounds__(kVisThreads)
set_block_idxs_kernel(TreeNodeDev* __restrict__ nodes, int64_t n_nodes, const float* __restrict__ centers,
                      int64_t n_blocks) {
  __shared__ float ctr[kCenterTile * 3];
  const int64_t n_round = (n_nodes + kVisThreads - 1) / kVisThreads * kVisThreads;
  for (int64_t node = (int64_t)blockIdx.x * kVisThreads + threadIdx.x; node < n_round;
       node += (int64_t)gridDim.x * kVisThreads) {
    const bool live = node < n_nodes;
    float px = 0.f, py = 0.f, pz = 0.f;
    if (live) {
      const float4 cs = *reinterpret_cast<const float4*>(nodes + node);
      px = cs.x, py = cs.y, pz = cs.z;
    }
    double best = 1e+9;
    int64_t best_idx = -1;
    for (int64_t b0 = 0; b0 < n_blocks; b0 += kCenterTile) {
      const int nb = (int)min((int64_t)kCenterTile, n_blocks - b0);
      __syncthreads();
      for (int i = threadIdx.x; i < nb * 3; i += kVisThreads) ctr[i] = centers[b0 * 3 + i];
      __syncthreads();
      for (int b = 0; b < nb; b++) {
        const float dx = __fadd_rn(px, -ctr[3 * b]), dy = __fadd_rn(py, -ctr[3 * b + 1]),
                    dz = __fadd_rn(pz, -ctr[3 * b + 2]);
        const double d = (double)__fsqrt_rn(__fmaf_rn(dx, dx, __fmaf_rn(dy, dy, __fmul_rn(dz, dz))));
        if (d < best) best = d, best_idx = b0 + b;
      }
    }
    if (live) nodes[node].block_idx = best_idx;
  }
}

}  // namespace
}  // namespace gf

using namespace gf;

extern "C" int64_t gf_octree_proc_device_scratch_bytes(int64_t n_in) {
  return n_in <= 0 ? 0 : n_in * (2 * (int64_t)sizeof(TreeNodeDev) + 2 * 8 + 6 * 4) + 256;
}

extern "C" int gf_octree_proc_device(const void* nodes_in, int64_t n_in, const int64_t* weight_stats_in,
                                     const int64_t* alpha_stats_in, const int64_t* visit_cnt_in, int compact,
                                     int subdivide, int brute_force, void* nodes_out, int64_t* weight_stats_out,
                                     int64_t* alpha_stats_out, int64_t capacity, void* scratch, int64_t scratch_bytes,
                                     int64_t* d_n_out, int32_t* d_error, void* stream) {
  GF_REQUIRE(nodes_in && n_in > 0 && n_in < 0x7fffffffLL / 9 && weight_stats_in && alpha_stats_in && visit_cnt_in,
             "gf_octree_proc_device: null input / empty or oversized tree");
  GF_REQUIRE(nodes_out && weight_stats_out && alpha_stats_out && capacity > 0 && d_n_out && d_error,
             "gf_octree_proc_device: null output");
  GF_REQUIRE(scratch && scratch_bytes >= gf_octree_proc_device_scratch_bytes(n_in),
             "gf_octree_proc_device: scratch of %lld bytes, %lld needed", (long long)scratch_bytes,
             (long long)gf_octree_proc_device_scratch_bytes(n_in));
  GF_REQUIRE((reinterpret_cast<uintptr_t>(nodes_in) & 15) == 0 && (reinterpret_cast<uintptr_t>(nodes_out) & 15) == 0 &&
                 (reinterpret_cast<uintptr_t>(scratch) & 15) == 0,
             "gf_octree_proc_device: node blobs and scratch must be 16-byte aligned");
  octree_proc_kernel<<<1, kProcThreads, 0, (cudaStream_t)stream>>>(
      (const TreeNodeDev*)nodes_in, n_in, weight_stats_in, alpha_stats_in, visit_cnt_in, compact, subdivide,
      brute_force, (TreeNodeDev*)nodes_out, weight_stats_out, alpha_stats_out, capacity, (unsigned char*)scratch,
      d_n_out, d_error);
  return check_launch("octree_proc_kernel");
}

extern "C" int gf_octree_mark_invisible(void* tree_nodes, int64_t n_nodes, const float* w2c, const float* intri,
                                        const float* bounds, int64_t n_cams, void* stream) {
  GF_REQUIRE(tree_nodes && n_nodes > 0, "gf_octree_mark_invisible: null / empty node blob");
  GF_REQUIRE(n_cams >= 0 && (n_cams == 0 || (w2c && intri && bounds)), "gf_octree_mark_invisible: null camera arrays");
  GF_REQUIRE((reinterpret_cast<uintptr_t>(tree_nodes) & 15) == 0,
             "gf_octree_mark_invisible: the node blob must be 16-byte aligned");
  mark_invisible_kernel<<<stride_grid(n_nodes, kVisThreads, 4), kVisThreads, 0, (cudaStream_t)stream>>>(
      (TreeNodeDev*)tree_nodes, n_nodes, w2c, intri, bounds, n_cams);
  return check_launch("mark_invisible_kernel");
}

extern "C" int gf_octree_set_block_idxs(void* tree_nodes, int64_t n_nodes, const float* centers, int64_t n_blocks,
                                        void* stream) {
  GF_REQUIRE(tree_nodes && n_nodes > 0, "gf_octree_set_block_idxs: null / empty node blob");
  GF_REQUIRE(n_blocks >= 0 && (n_blocks == 0 || centers), "gf_octree_set_block_idxs: null centres");
  GF_REQUIRE((reinterpret_cast<uintptr_t>(tree_nodes) & 15) == 0,
             "gf_octree_set_block_idxs: the node blob must be 16-byte aligned");
  set_block_idxs_kernel<<<stride_grid(n_nodes, kVisThreads, 4), kVisThreads, 0, (cudaStream_t)stream>>>(
      (TreeNodeDev*)tree_nodes, n_nodes, centers, n_blocks);
  return check_launch("set_block_idxs_kernel");
}
