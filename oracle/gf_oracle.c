/*
 * gf_oracle.c -- CPU ORACLE for the GF-NeRF per-ray hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This file is a plain-C restatement of the
 * reference's CUDA / torch algorithm for the path; it is the checker the CUDA
 * kernels are compared with, and the CPU baseline bench.py times.  Nothing in
 * the product package (gf-nerf_b200/) may import, link or execute it; only
 * tests/, __graft_entry__.smoke() and bench.py (cpu_baseline / --impl reference)
 * do.
 *
 * HOW PARITY IS PINNED.  The reference ships no golden vectors, KATs or tests
 * for Hash3DAnchored / PersSampler and its own build cannot
 * run here (un-vendored tiny-cuda-nn, a patched Eigen).  But its DEVICE CODE
 * needs neither: `make -C oracle ref` extracts the __global__ / __device__
 * bodies of Hash3DAnchored_cuda.cu and PersSampler_cuda.cu where they lie under
 * /root/reference and compiles them for the host between a CUDA shim and a
 * fixed-size Eigen subset (oracle/ref_driver.cpp, ref_shim/, -> oracle/_ref/).
 * tests/test_ref_kernels.py holds the hash / traversal / march / vote / cold-
 * query functions below to that code: live where /root/reference exists, and
 * through tests/golden/ref_kernels.npz (its outputs) everywhere.  Integer
 * results are identical; the hash blend is identical bit for bit; the march
 * agrees to 1e-4 (fp contraction is the compiler's choice: nvcc's for the
 * reference binary, g++'s for the host build, spelled out here for nvcc).
 * The host side is held to the reference too: PersOctree::ProcOctree and
 * ConstructEdgePool run from the same extraction (stand-in Tensor), and the
 * whole PersOctree constructor (GetVisiCams, DistanceSummary, ConstructTreeNode,
 * PCA, ConstructTrans) compiled against this image's real libtorch on the CPU
 * (oracle/ref_driver_torch.cpp, tests/test_ref_octree.py): same tree, field for
 * field.
 * What stays UNPINNED: the last bit of fp32 results that depend on nvcc's
 * contraction and on Eigen's evaluation order inside the shim (our reading of
 * Eigen 3.4), and the tcnn SH-4 encoding (un-vendored third party).
 * The pieces of the path that are importable Python (get_weights_f2nerf,
 * renderers, MLPNetwork, trunc_exp, CharbonnierLoss, S3IM, Adam) pin the
 * composite / MLP / loss functions below through tests/golden/make_golden.py.
 *
 * Floating-point convention (what "bit-exact" means for indices / node ids /
 * sample counts): fp32 everywhere, IEEE div/sqrt, and the FMA contraction nvcc
 * (-fmad=true, its default and the reference's build) applies to the
 * reference's expression order -- verified for the hash blend by compiling the
 * expression with nvcc 12.9 (DESIGN.md "FMA convention"): a*b + c*d + e*f ...
 * becomes t = c*d; t = fma(a,b,t); t = fma(e,f,t) ...  Fixed-size Eigen products of
 * the reference follow Eigen 3.4's evaluation order (see the PersSampler section).
 * Build with -ffp-contract=off so that only the explicit fmaf() calls fuse.
 *
 * Reference paths below are relative to gfnerf/bindings/ unless they start with
 * nerfstudio/ or gfnerf/.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define N_LEVELS 16
#define N_CHANNELS 2
#define N_PROS 12
#define MAX_STACK_SIZE 48
#define MAX_SAMPLE_PER_RAY 1024
#define TREE_NODE_BYTES 128
#define TRANS_INFO_BYTES 576

typedef _Float16 half_t;

static inline float h2f(half_t h) { return (float)h; }
static inline half_t f2h(float f) { return (half_t)f; } /* round-to-nearest-even */

int orc_num_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
void orc_set_num_threads(int n) {
#ifdef _OPENMP
  omp_set_num_threads(n);
#else
  (void)n;
#endif
}

/* ------------------------------------------------------------------------- */
/* Hash3DAnchored                                                            */
/* ------------------------------------------------------------------------- */

/* host evaluation of the level scales, field/Hash3DAnchored_cuda.cu:28 */
void orc_hash_level_scales(float* scales16) {
  for (int l = 0; l < N_LEVELS; l++)
    scales16[l] = exp2f((10.f - 3.f) * (float)l / (float)(N_LEVELS - 1) + 3.f);
}

/* static_cast<unsigned>(floorf(x)) with the GPU's saturating cvt.rzi.u32.f32
 * (Hash3DAnchored_cuda.cu:44-46; UB on the CPU for x<0, so pinned here). */
static inline uint32_t f2u_sat(float x) {
  if (!(x > 0.f)) return 0u; /* negatives, -0, NaN -> 0 */
  if (x >= 4294967296.f) return 0xffffffffu;
  return (uint32_t)x;
}

typedef struct {
  uint32_t pos[8];
  float w[8];
} hash_cell;

/* First table ROW of level l.  Hash3DAnchored.cpp:66-70 computes feat_local_idx[l] = l * local_size, but the
 * kernels add it to a pointer to SCALARS (`T* feat_pool`, Hash3DAnchored_cuda.cu:38; `T* grad_out`, :105) and then
 * index pos * N_CHANNELS + k: level l starts at row l * local_size / 2, consecutive levels overlap by half a
 * window, rows >= 8.5 * local_size are never used.  PINNED by the reference's own kernel bodies run on the host
 * (oracle/ref_driver.cpp, tests/test_ref_kernels.py).  local_size is even (the reference makes it a multiple of 16). */
static inline int64_t level_base_row(int l, int64_t local_size) { return ((int64_t)l * local_size) >> 1; }

/* index + weight math shared by forward and backward, Hash3DAnchored_cuda.cu:26-69 / 97-140 */
static inline void hash_cell_eval(const float* pt, float mul, const float* bias, const int32_t* prim,
                                  uint32_t local_size, hash_cell* c) {
  float p0 = fmaf(pt[0], mul, bias[0]); /* pt *= mul; pt = pt + bias  (nvcc contracts) */
  float p1 = fmaf(pt[1], mul, bias[1]);
  float p2 = fmaf(pt[2], mul, bias[2]);
  float f0 = floorf(p0), f1 = floorf(p1), f2 = floorf(p2);
  uint32_t px = f2u_sat(f0), py = f2u_sat(f1), pz = f2u_sat(f2);
  uint32_t pa = (uint32_t)prim[0], pb = (uint32_t)prim[1], pc = (uint32_t)prim[2];
  /* order 000,001,010,011,100,101,110,111 with bits (x,y,z) = (a,b,c) */
  for (int d = 0; d < 8; d++) {
    uint32_t dx = (d >> 2) & 1u, dy = (d >> 1) & 1u, dz = d & 1u;
    c->pos[d] = (((px + dx) * pa) ^ ((py + dy) * pb) ^ ((pz + dz) * pc)) % local_size;
  }
  float a = p0 - f0, b = p1 - f1, cc = p2 - f2;
  float ia = 1.f - a, ib = 1.f - b, ic = 1.f - cc;
  c->w[0] = ia * ib * ic;
  c->w[1] = ia * ib * cc;
  c->w[2] = ia * b * ic;
  c->w[3] = ia * b * cc;
  c->w[4] = a * ib * ic;
  c->w[5] = a * ib * cc;
  c->w[6] = a * b * ic;
  c->w[7] = a * b * cc;
}

/*
 * Hash3DAnchoredForwardKernel + the casts of Hash3DAnchoredFunction::forward
 * (field/Hash3DAnchored_cuda.cu:11-79, 185, 195).
 * feat_f32 [16*local_size,2] is rounded to fp16 first (the `.to(kFloat16)`);
 * out [n,32] is the fp16-rounded blend widened to fp32.
 * idx_out (optional) int32 [n,16,8] receives the table row of every corner.
 */
void orc_hash_forward(int64_t n, int32_t n_volumes, int64_t local_size, const float* feat_f32,
                      const int32_t* prim_pool, const float* bias_pool, const float* scales,
                      const float* pts, const int64_t* anchors, float* out, int32_t* idx_out) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    int64_t vol = anchors[i];
    for (int l = 0; l < N_LEVELS; l++) {
      hash_cell c;
      int64_t tr = (int64_t)l * n_volumes + vol;
      hash_cell_eval(pts + 3 * i, scales[l], bias_pool + 3 * tr, prim_pool + 3 * tr,
                     (uint32_t)local_size, &c);
      const float* tab = feat_f32 + (int64_t)l * local_size; /* :38 the pointer is a SCALAR pointer, see level_base_row */
      for (int k = 0; k < N_CHANNELS; k++) {
        float f[8];
        for (int d = 0; d < 8; d++) f[d] = h2f(f2h(tab[(int64_t)c.pos[d] * N_CHANNELS + k]));
        /* nvcc: t = w001*f001; t = fma(w000,f000,t); t = fma(w010,f010,t); ... */
        float t = c.w[1] * f[1];
        t = fmaf(c.w[0], f[0], t);
        for (int d = 2; d < 8; d++) t = fmaf(c.w[d], f[d], t);
        out[i * (N_LEVELS * N_CHANNELS) + l * N_CHANNELS + k] = h2f(f2h(t));
      }
      if (idx_out)
        for (int d = 0; d < 8; d++)
          idx_out[(i * N_LEVELS + l) * 8 + d] = (int32_t)(level_base_row(l, local_size) + c.pos[d]);
    }
  }
}

/*
 * Hash3DAnchoredBackwardKernel + Hash3DAnchoredFunction::backward
 * (field/Hash3DAnchored_cuda.cu:81-155, 198-239).
 * grad_out [n,32] fp32 -> grad_in = fp16(grad_out*128); each corner receives
 * fp16(w_d * grad_in) (:148-151); the reference sums those with half2 atomics
 * (order-dependent fp16 rounding) -- the oracle sums them EXACTLY (fp64) and
 * divides by 128, which is what any summation order approximates.
 * grad_table fp64 [16*local_size,2], zeroed here.
 */
void orc_hash_backward(int64_t n, int32_t n_volumes, int64_t local_size, const int32_t* prim_pool,
                       const float* bias_pool, const float* scales, const float* pts,
                       const int64_t* anchors, const float* grad_out, double* grad_table) {
  memset(grad_table, 0, sizeof(double) * (size_t)(N_LEVELS * local_size * N_CHANNELS));
  /* every addend is a multiple of 2^-31 (fp16 value / 128) and the sums stay far below 2^21, so the fp64
     accumulation is exact and the result does not depend on the order the threads add in */
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    int64_t vol = anchors[i];
    for (int l = 0; l < N_LEVELS; l++) {
      hash_cell c;
      int64_t tr = (int64_t)l * n_volumes + vol;
      hash_cell_eval(pts + 3 * i, scales[l], bias_pool + 3 * tr, prim_pool + 3 * tr,
                     (uint32_t)local_size, &c);
      float g0 = h2f(f2h(grad_out[i * 32 + l * 2 + 0] * 128.f));
      float g1 = h2f(f2h(grad_out[i * 32 + l * 2 + 1] * 128.f));
      if (g0 != 0.f || g1 != 0.f) {
        double* tab = grad_table + (int64_t)l * local_size; /* :105 scalar offset, as in the forward */
        for (int d = 0; d < 8; d++) {
          double a0 = (double)h2f(f2h(g0 * c.w[d])) / 128.0, a1 = (double)h2f(f2h(g1 * c.w[d])) / 128.0;
#pragma omp atomic
          tab[(int64_t)c.pos[d] * 2 + 0] += a0;
#pragma omp atomic
          tab[(int64_t)c.pos[d] * 2 + 1] += a1;
        }
      }
    }
  }
}

/* ------------------------------------------------------------------------- */
/* PersSampler                                                               */
/* ------------------------------------------------------------------------- */

typedef struct {
  float center[3];
  float side_len;
  int64_t parent;
  int64_t childs[8];
  uint8_t is_leaf_node;
  uint8_t pad0[7];
  int64_t trans_idx;
  int64_t block_idx;
  uint8_t pad1[16];
} tree_node; /* PtsSampler/PersSampler.h:40-49, alignas(32) -> 128 B */

typedef struct {
  float w2xz[N_PROS][2][4];
  float weight[3][N_PROS];
  float center[3];
  float side_len;
  float dis_summary;
  uint8_t pad[28];
} trans_info; /* PtsSampler/PersSampler.h:31-38, alignas(32) -> 576 B */

int orc_sizeof_tree_node(void) { return (int)sizeof(tree_node); }
int orc_sizeof_trans_info(void) { return (int)sizeof(trans_info); }

/* child visiting order per ray octant, PtsSampler/PersSampler.cpp:137-151:
 * std::sort with cmp(a,b) = (a&bt)^(st&bt), bt = lowest differing bit of a,b --
 * i.e. a strict order on key(a) = bitrev3(a ^ st), descending. */
void orc_search_order(uint8_t* order64) {
  for (int st = 0; st < 8; st++) {
    int idx[8];
    for (int i = 0; i < 8; i++) idx[i] = i;
    /* insertion sort with the reference comparator (the order is total, so any
       correct sort yields the same permutation) */
    for (int i = 1; i < 8; i++) {
      int v = idx[i], j = i - 1;
      while (j >= 0) {
        int a = v, b = idx[j];
        int bt = (a ^ b) & -(a ^ b);
        int a_before_b = ((a & bt) ^ (st & bt)) != 0;
        if (!a_before_b) break;
        idx[j + 1] = idx[j];
        j--;
      }
      idx[j + 1] = v;
    }
    for (int i = 0; i < 8; i++) order64[st * 8 + i] = (uint8_t)idx[i];
  }
}

/* GetIntersection, PersSampler_cuda.cu:21-51 */
static inline void get_intersection(const float* o, const float* d, const float* c, float side,
                                    float* near, float* far) {
  float tmp[3][2];
  float hf = side * .5f;
  for (int i = 0; i < 3; i++) {
    if (d[i] < 1e-6f && d[i] > -1e-6f) {
      if (o[i] > c[i] - hf && o[i] < c[i] + hf) {
        tmp[i][0] = -1e6f;
        tmp[i][1] = 1e6f;
      } else {
        tmp[i][0] = 1e6f;
        tmp[i][1] = -1e6f;
      }
    } else if (d[i] > 0) {
      tmp[i][0] = (c[i] - hf - o[i]) / d[i];
      tmp[i][1] = (c[i] + hf - o[i]) / d[i];
    } else {
      tmp[i][0] = (c[i] + hf - o[i]) / d[i];
      tmp[i][1] = (c[i] - hf - o[i]) / d[i];
    }
  }
  *near = fmaxf(*near, fmaxf(tmp[0][0], fmaxf(tmp[1][0], tmp[2][0])));
  *far = fminf(*far, fminf(tmp[0][1], fminf(tmp[1][1], tmp[2][1])));
}

/* FindRayOctreeIntersectionKernel, PersSampler_cuda.cu:53-152 (one pass; the
 * count pass and the fill pass of the reference visit the same leaves). */
static int64_t find_ray_octree_intersection(const float* o, const float* d, float overall_near,
                                            float overall_far, const tree_node* nodes,
                                            const uint8_t* search_order, int64_t max_cnt,
                                            int64_t* out_idx, float* out_near_far) {
  int64_t stack_info[MAX_STACK_SIZE];
  int64_t stack_ptr = 0, cnt = 0;
  stack_info[0] = 0;
  stack_info[1] = -1;
  int64_t ray_st = ((int64_t)(d[0] > 0.f) << 2) | ((int64_t)(d[1] > 0.f) << 1) | (int64_t)(d[2] > 0.f);
  const uint8_t* so = search_order + ray_st * 8;
  while (stack_ptr >= 0 && cnt < max_cnt) {
    int64_t u = stack_info[stack_ptr * 2];
    const tree_node* node = nodes + u;
    if (stack_info[stack_ptr * 2 + 1] == -1) {
      float cur_near = overall_near, cur_far = overall_far;
      get_intersection(o, d, node->center, node->side_len, &cur_near, &cur_far);
      if (cur_near < cur_far) {
        int64_t child_ptr = 0;
        while (child_ptr < 8 && node->childs[so[child_ptr]] < 0) child_ptr++;
        if (child_ptr < 8) {
          stack_info[stack_ptr * 2 + 1] = child_ptr;
          stack_ptr++;
          stack_info[stack_ptr * 2] = node->childs[so[child_ptr]];
          stack_info[stack_ptr * 2 + 1] = -1;
        } else {
          if (node->trans_idx >= 0) {
            out_idx[cnt] = u;
            out_near_far[cnt * 2] = cur_near;
            out_near_far[cnt * 2 + 1] = cur_far;
            cnt++;
          }
          stack_ptr--;
        }
      } else {
        stack_ptr--;
      }
    } else {
      int64_t child_ptr = stack_info[stack_ptr * 2 + 1] + 1;
      while (child_ptr < 8 && node->childs[so[child_ptr]] < 0) child_ptr++;
      if (child_ptr < 8) {
        stack_info[stack_ptr * 2 + 1] = child_ptr;
        stack_ptr++;
        stack_info[stack_ptr * 2] = node->childs[so[child_ptr]];
        stack_info[stack_ptr * 2 + 1] = -1;
      } else {
        stack_ptr--;
      }
    }
  }
  return cnt;
}

static inline float norm3(float x, float y, float z) {
  /* Eigen norm(): sqrt of the unrolled redux x^2 + (y^2 + z^2), contracted */
  return sqrtf(fmaf(x, x, fmaf(y, y, z * z)));
}

/*
 * Evaluation order of the reference's fixed-size Eigen expressions (Eigen 3.4.0, scalar device path;
 * restated from the structure of Eigen's Redux.h / ProductEvaluators.h / GeneralProduct.h, from memory --
 * Eigen itself is not in this image; oracle/ref_shim/eigen_subset.h implements the same reading, so the
 * host build of the reference kernels cannot confirm it -- the last-bit part of parity that stays unpinned):
 *   - small products (2x4 * 4x1, 1x2 * 2x3, 3x3 * 3x1, and 3x12 * 12x3 which falls under
 *     EIGEN_GEMM_TO_COEFFBASED_THRESHOLD) are coefficient-based: lhs.row(i).cwiseProduct(rhs.col(j)).sum(),
 *     and a fully unrolled sum() splits its range in halves recursively (redux_novec_unroller):
 *       len 2: a0 + a1;  len 3: a0 + (a1 + a2);  len 4: (a0 + a1) + (a2 + a3);
 *       len 12: ((a0 + (a1 + a2)) + (a3 + (a4 + a5))) + ((a6 + (a7 + a8)) + (a9 + (a10 + a11)))
 *   - 3x12 * 12x1 (depth 12 > 8 with a vector rhs) is a GEMV: sequential accumulation over k from 0.
 * nvcc contraction (-fmad=true) of every `product + product` is  t = second product (rounded); fma(first, t),
 * of `product + sum` is fma(product, sum)  -- the rule verified for the hash blend (DESIGN.md "FMA convention").
 */

/* W_i (2x4) * [xyz;1], PersSampler_cuda.cu:163,178 : (w0 x + w1 y) + (w2 z + w3 * 1) */
static inline void proj_xz(const float w[2][4], const float* p, float* xz) {
  for (int r = 0; r < 2; r++)
    xz[r] = fmaf(w[r][0], p[0], w[r][1] * p[1]) + fmaf(w[r][2], p[2], w[r][3]);
}

/* QueryFrameTransform, PersSampler_cuda.cu:155-170 (weight * transed_vals: GEMV, sequential) */
static inline void query_frame_transform(const trans_info* tr, const float* p, float* out) {
  float v[N_PROS];
  for (int i = 0; i < N_PROS; i++) {
    float xz[2];
    proj_xz(tr->w2xz[i], p, xz);
    v[i] = xz[0] / xz[1];
  }
  for (int r = 0; r < 3; r++) {
    /* w0 v0 + w1 v1 + w2 v2 + ... under nvcc's contraction: the SECOND product is the rounded one,
     * t = w1 v1; t = fma(w0, v0, t); t = fma(w2, v2, t); ...  (read off the SASS of the reference's RayMarchKernel
     * built by nvcc for sm_100a, profiles/r02a_ref_march_sass_notes.md; ours beside it on the B200 is bit-equal) */
    float acc = tr->weight[r][1] * v[1];
    acc = fmaf(tr->weight[r][0], v[0], acc);
    for (int k = 2; k < N_PROS; k++) acc = fmaf(tr->weight[r][k], v[k], acc);
    out[r] = acc;
  }
}

/* a0 + (a1 + a2) with a_k = w[k] * t[k*stride] */
static inline float tri_sum(const float* w, const float* t, int stride) {
  return fmaf(w[0], t[0], fmaf(w[1], t[stride], w[2] * t[2 * stride]));
}

/* QueryFrameTransformJac, PersSampler_cuda.cu:172-188 (weight * transed_jac: coefficient-based, tree sum) */
static inline void query_frame_transform_jac(const trans_info* tr, const float* p, float jac[3][3]) {
  float tj[N_PROS][3];
  for (int i = 0; i < N_PROS; i++) {
    float xz[2];
    proj_xz(tr->w2xz[i], p, xz);
    float dv0 = 1.f / xz[1];
    float dv1 = -xz[0] / (xz[1] * xz[1]);
    for (int c = 0; c < 3; c++) tj[i][c] = fmaf(dv0, tr->w2xz[i][0][c], dv1 * tr->w2xz[i][1][c]);
  }
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) {
      const float* w = tr->weight[r];
      float q0 = tri_sum(w, &tj[0][c], 3), q1 = tri_sum(w + 3, &tj[3][c], 3);
      float q2 = tri_sum(w + 6, &tj[6][c], 3), q3 = tri_sum(w + 9, &tj[9][c], 3);
      jac[r][c] = (q0 + q1) + (q2 + q3);
    }
}

/*
 * PersSampler::GetSamples = FindRayOctreeIntersectionKernel + RayMarchKernel
 * (PersSampler_cuda.cu:53-152, 190-318, 321-477).  rays_d must already be
 * normalised (:323), noise already multiplied by ray_march_fineness_ (:389).
 * Dense reference layout; outputs must be zero-filled by the caller (:437-444).
 * Any output pointer may be NULL.  counts[R] = samples per ray,
 * n_oct[R] = leaves intersected, oct_idx/oct_nf (optional) dense [R,max_oct].
 */
void orc_sampler_get_samples(int64_t n_rays, const float* rays_o, const float* rays_d,
                             const float* noise, const void* tree_nodes_blob,
                             const void* pers_trans_blob, const uint8_t* search_order,
                             float global_near, float sample_l, int scale_by_dis,
                             int64_t max_oct_per_ray, float* world_pts, float* warp_pts,
                             float* dirs, float* dists, float* ts, int64_t* anchors,
                             int32_t* counts, float* first_oct_dis, int32_t* n_oct,
                             int64_t* oct_idx_out, float* oct_nf_out) {
  const tree_node* nodes = (const tree_node*)tree_nodes_blob;
  const trans_info* transes = (const trans_info*)pers_trans_blob;
#pragma omp parallel
  {
    int64_t* oct_idx = (int64_t*)malloc(sizeof(int64_t) * (size_t)max_oct_per_ray);
    float* oct_nf = (float*)malloc(sizeof(float) * 2 * (size_t)max_oct_per_ray);
#pragma omp for schedule(dynamic, 16)
    for (int64_t ray = 0; ray < n_rays; ray++) {
      const float* o = rays_o + 3 * ray;
      const float* d = rays_d + 3 * ray;
      int64_t n_oct_nodes = find_ray_octree_intersection(o, d, global_near, 1e8f, nodes, search_order,
                                                         max_oct_per_ray, oct_idx, oct_nf);
      if (n_oct) n_oct[ray] = (int32_t)n_oct_nodes;
      if (oct_idx_out)
        for (int64_t k = 0; k < n_oct_nodes; k++) {
          oct_idx_out[ray * max_oct_per_ray + k] = oct_idx[k];
          oct_nf_out[(ray * max_oct_per_ray + k) * 2] = oct_nf[2 * k];
          oct_nf_out[(ray * max_oct_per_ray + k) * 2 + 1] = oct_nf[2 * k + 1];
        }
      if (first_oct_dis) first_oct_dis[ray] = n_oct_nodes > 0 ? oct_nf[0] : 1e9f;
      int64_t pts_ptr = 0;
      if (n_oct_nodes > 0) { /* the reference reads OOB here when a ray hits nothing (:244) */
        const float* rn = noise + ray;
        const int64_t base = ray * MAX_SAMPLE_PER_RAY;
        int64_t oct_ptr = 0;
        int64_t cur_oct_idx = oct_idx[0];
        float cur_march_step = 0.f, exp_march_step = 0.f;
        float cur_t = oct_nf[0], cur_far = oct_nf[1], cur_near = oct_nf[0];
        float cur_xyz[3];
        for (int c = 0; c < 3; c++) cur_xyz[c] = fmaf(d[c], cur_t, o[c]);
        int the_first_pts = 1;
        while (pts_ptr < MAX_SAMPLE_PER_RAY && oct_ptr < n_oct_nodes) {
          const tree_node* cur_node = nodes + cur_oct_idx;
          const trans_info* tr = transes + cur_node->trans_idx;
          float cur_radius =
              norm3(o[0] - tr->center[0], o[1] - tr->center[1], o[2] - tr->center[2]) / tr->dis_summary;
          float cur_radius_clip = fmaxf(cur_radius, 1.f);
          float jac[3][3];
          query_frame_transform_jac(tr, cur_xyz, jac);
          float proj[3];
          for (int r = 0; r < 3; r++) proj[r] = fmaf(jac[r][0], d[0], fmaf(jac[r][1], d[1], jac[r][2] * d[2]));   /* len 3: a0 + (a1 + a2) */
          float pn = norm3(proj[0], proj[1], proj[2]) + 1e-6f;
          float exp_march_step_warp = sample_l * rn[pts_ptr];
          exp_march_step = exp_march_step_warp / pn;
          if (scale_by_dis) exp_march_step *= cur_radius_clip;
          cur_march_step = exp_march_step;
          if (!the_first_pts) {
            int64_t s = base + pts_ptr;
            if (world_pts) memcpy(world_pts + 3 * s, cur_xyz, 12);
            if (ts) ts[s] = cur_t;
            if (dirs) memcpy(dirs + 3 * s, d, 12);
            if (warp_pts) query_frame_transform(tr, cur_xyz, warp_pts + 3 * s);
            if (dists) dists[s] = exp_march_step * pn;
            if (anchors) {
              anchors[3 * s + 0] = cur_node->trans_idx;
              anchors[3 * s + 1] = cur_oct_idx;
              anchors[3 * s + 2] = cur_node->block_idx;
            }
            pts_ptr += 1;
          }
          /* nvcc fuses `cur_march_step = exp_march_step * float(ex_march_steps)` into both of its consumers, the loop
           * condition and `cur_t += cur_march_step` (SASS of the reference kernel: FFMA R3 = exp * ex + cur_t;
           * FSETP.GT R3, cur_far; ... MOV cur_t, R3): the position after a crossing is rounded once */
          float next_t = cur_t + cur_march_step;
          while (next_t > cur_far) {
            oct_ptr++;
            if (oct_ptr >= n_oct_nodes) break;
            cur_oct_idx = oct_idx[oct_ptr];
            cur_near = oct_nf[2 * oct_ptr];
            cur_far = oct_nf[2 * oct_ptr + 1];
            int64_t ex_march_steps = (int64_t)ceilf(fmaxf((cur_near - cur_t) / exp_march_step, 1.f));
            cur_march_step = exp_march_step * (float)ex_march_steps;
            next_t = fmaf(exp_march_step, (float)ex_march_steps, cur_t);
          }
          cur_t = next_t;
          for (int c = 0; c < 3; c++) cur_xyz[c] = fmaf(d[c], cur_t, o[c]);
          the_first_pts = 0;
        }
      }
      counts[ray] = (int32_t)pts_ptr;
    }
    free(oct_idx);
    free(oct_nf);
  }
}

/* TransQueryFrameKernel, PersSampler_cuda.cu:854-922 */
void orc_trans_query_frame(int64_t n_pts, const void* tree_nodes_blob, int64_t n_nodes,
                           const void* pers_trans_blob, const int64_t* anchors,
                           const float* world_pts, float* out) {
  const tree_node* nodes = (const tree_node*)tree_nodes_blob;
  const trans_info* transes = (const trans_info*)pers_trans_blob;
  for (int64_t i = 0; i < n_pts; i++) {
    int64_t a = anchors[i];
    if (a >= n_nodes || a < 0) continue;
    const tree_node* nd = nodes + a;
    if (!nd->is_leaf_node) continue;
    if (nd->trans_idx >= 0) {
      query_frame_transform(transes + nd->trans_idx, world_pts + 3 * i, out + 3 * i);
    } else {
      /* (p - c) / (side * 0.5): 0.5 is a double literal in the reference (:913) */
      for (int c = 0; c < 3; c++)
        out[3 * i + c] = (float)((double)(world_pts[3 * i + c] - nd->center[c]) / ((double)nd->side_len * 0.5));
    }
  }
}

/*
 * MarkVistNodeKernel + stat update + MarkInvalidNodes of
 * PersSampler::UpdateOctNodes (PersSampler_cuda.cu:518-655), dense layout:
 * oct_indices/weights/alphas are [R,1024]; counts[R] valid samples per ray.
 */
void orc_update_oct_nodes(int64_t n_rays, const int32_t* counts, const int64_t* oct_indices,
                          const float* weights, const float* alphas, void* tree_nodes_blob,
                          int64_t n_nodes, int64_t* weight_stats, int64_t* alpha_stats,
                          int64_t* visit_cnt) {
  tree_node* nodes = (tree_node*)tree_nodes_blob;
  int64_t* w_add = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_nodes);
  int64_t* a_add = (int64_t*)malloc(sizeof(int64_t) * (size_t)n_nodes);
  int64_t* mark = (int64_t*)calloc((size_t)n_nodes, sizeof(int64_t));
  for (int64_t i = 0; i < n_nodes; i++) w_add[i] = a_add[i] = -1;
#define AMAX(arr, i, v) do { if ((arr)[i] < (v)) (arr)[i] = (v); } while (0)
  for (int64_t ray = 0; ray < n_rays; ray++) {
    int64_t s0 = ray * MAX_SAMPLE_PER_RAY, s1 = s0 + counts[ray];
    if (s0 >= s1) continue;
    float max_w = 0.f, max_a = 0.f;
    for (int64_t s = s0; s < s1; s++) {
      max_w = fmaxf(max_w, weights[s]);
      max_a = fmaxf(max_a, alphas[s]);
    }
    /* REL_*_THRES / ABS_*_THRES are double literals: the product is formed in
       double and fminf narrows it (:543-544) */
    const float w_thres = fminf((float)((double)max_w * 0.1), (float)0.01);
    const float a_thres = fminf((float)((double)max_a * 0.1), (float)0.02);
    float cur_w = 0.f, cur_a = 0.f;
    int64_t cur_oct = -1, cur_cnt = 0;
    for (int64_t s = s0; s < s1; s++) {
      if (cur_oct != oct_indices[s]) {
        if (cur_oct >= 0) {
          AMAX(w_add, cur_oct, (int64_t)(cur_w > w_thres ? 512 : -1));
          AMAX(a_add, cur_oct, (int64_t)(cur_a > a_thres ? 32 : -1));
          AMAX(visit_cnt, cur_oct, cur_cnt);
          mark[cur_oct] = 1;
        }
        cur_oct = oct_indices[s];
        cur_w = 0.f;
        cur_a = 0.f;
        cur_cnt = 0;
      }
      cur_w = fmaxf(cur_w, weights[s]);
      cur_a = fmaxf(cur_a, alphas[s]);
      cur_cnt += 1;
    }
    if (cur_oct >= 0) {
      AMAX(w_add, cur_oct, (int64_t)(cur_w > w_thres ? 512 : -1));
      AMAX(a_add, cur_oct, (int64_t)(cur_a > a_thres ? 32 : -1));
      AMAX(visit_cnt, cur_oct, cur_cnt);
      mark[cur_oct] = 1;
    }
  }
#undef AMAX
  for (int64_t i = 0; i < n_nodes; i++) {
    /* :634-646  stats = max(stats, mask*adder); stats += mark*(1-mask)*adder; clamp */
    int64_t m = w_add[i] > 0;
    int64_t s = weight_stats[i];
    if (m * w_add[i] > s) s = m * w_add[i];
    s += mark[i] * (1 - m) * w_add[i];
    if (s < -100) s = -100;
    if (s > (1 << 20)) s = 1 << 20;
    weight_stats[i] = s;
    m = a_add[i] > 0;
    s = alpha_stats[i];
    if (m * a_add[i] > s) s = m * a_add[i];
    s += mark[i] * (1 - m) * a_add[i];
    if (s < -100) s = -100;
    if (s > (1 << 20)) s = 1 << 20;
    alpha_stats[i] = s;
    if (weight_stats[i] < 0 || alpha_stats[i] < 0) nodes[i].trans_idx = -1; /* MarkInvalidNodes :576-582 */
  }
  free(w_add);
  free(a_add);
  free(mark);
}

/* ------------------------------------------------------------------------- */
/* Compositing                                                               */
/* ------------------------------------------------------------------------- */

static inline float nan_to_num_f(float x) {
  if (isnan(x)) return 0.f;
  if (isinf(x)) return x > 0 ? 3.4028234663852886e38f : -3.4028234663852886e38f;
  return x;
}

/*
 * RaySamples.get_weights_f2nerf (nerfstudio/cameras/rays.py:178-200) +
 * RGBRenderer.combine_rgb (renderers.py:97-110, training mode: no nan_to_num,
 * no clamp, no background), DepthRenderer 'expected' before its clip
 * (:269-280), AccumulationRenderer (:220).  CSR layout: samples of ray r are
 * [offsets[r], offsets[r+1]).  The prefix sum runs in fp64 and is narrowed, so
 * the oracle is the value any fp32 summation order approximates.
 */
void orc_composite_forward(int64_t n_rays, const int32_t* offsets, const float* sigma,
                           const float* delta, const float* rgb, const float* t, float* weights,
                           float* alphas, float* trans, float* out_rgb, float* out_depth,
                           float* out_acc) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t r = 0; r < n_rays; r++) {
    double cum = 0.0, cr = 0, cg = 0, cb = 0, cd = 0, ca = 0;
    for (int32_t s = offsets[r]; s < offsets[r + 1]; s++) {
      float dd = delta[s] * sigma[s];
      float alpha = 1.f - expf(-dd);
      float T = expf(-(float)cum);
      float w = nan_to_num_f(alpha * T);
      if (weights) weights[s] = w;
      if (alphas) alphas[s] = alpha;
      if (trans) trans[s] = T;
      cr += (double)w * rgb[3 * s + 0];
      cg += (double)w * rgb[3 * s + 1];
      cb += (double)w * rgb[3 * s + 2];
      cd += (double)w * t[s];
      ca += (double)w;
      cum += (double)dd;
    }
    out_rgb[3 * r + 0] = (float)cr;
    out_rgb[3 * r + 1] = (float)cg;
    out_rgb[3 * r + 2] = (float)cb;
    if (out_depth) out_depth[r] = (float)(cd / (ca + 1e-10));
    if (out_acc) out_acc[r] = (float)ca;
  }
}

/*
 * Autograd of the above w.r.t. sigma and rgb given g_rgb [R,3] (and g_acc [R],
 * optional): w_i = (1-exp(-dd_i)) * exp(-sum_{j<i} dd_j),
 * dL/ddd_i = gw_i * exp(-dd_i) * T_i - sum_{j>i} gw_j w_j, gw_i = g_rgb.c_i + g_acc.
 * (nan_to_num has zero gradient where it fires; ignored: weights are finite for
 * finite inputs.)
 */
void orc_composite_backward(int64_t n_rays, const int32_t* offsets, const float* sigma,
                            const float* delta, const float* rgb, const float* g_rgb,
                            const float* g_acc, float* d_sigma, float* d_rgb) {
#pragma omp parallel for schedule(dynamic, 64)
  for (int64_t r = 0; r < n_rays; r++) {
    int32_t s0 = offsets[r], s1 = offsets[r + 1];
    double cum = 0.0;
    /* forward sweep to get total, then backward sweep for the suffix sums */
    double total = 0.0;
    for (int32_t s = s0; s < s1; s++) total += (double)(delta[s] * sigma[s]);
    (void)total;
    /* first pass: weights and gw, store suffix in reverse */
    double suffix = 0.0;
    /* need T_i: recompute cum prefix going forward, so do a forward pass storing T */
    int32_t n = s1 - s0;
    double* Tbuf = (double*)malloc(sizeof(double) * (size_t)(n > 0 ? n : 1));
    for (int32_t i = 0; i < n; i++) {
      Tbuf[i] = exp(-cum);
      cum += (double)(delta[s0 + i] * sigma[s0 + i]);
    }
    for (int32_t i = n - 1; i >= 0; i--) {
      int32_t s = s0 + i;
      double dd = (double)(delta[s] * sigma[s]);
      double e = exp(-dd);
      double w = (1.0 - e) * Tbuf[i];
      double gw = (double)g_rgb[3 * r] * rgb[3 * s] + (double)g_rgb[3 * r + 1] * rgb[3 * s + 1] +
                  (double)g_rgb[3 * r + 2] * rgb[3 * s + 2] + (g_acc ? (double)g_acc[r] : 0.0);
      double d_dd = gw * e * Tbuf[i] - suffix;
      d_sigma[s] = (float)(d_dd * (double)delta[s]);
      d_rgb[3 * s + 0] = (float)(w * g_rgb[3 * r + 0]);
      d_rgb[3 * s + 1] = (float)(w * g_rgb[3 * r + 1]);
      d_rgb[3 * s + 2] = (float)(w * g_rgb[3 * r + 2]);
      suffix += gw * w;
    }
    free(Tbuf);
  }
}

/* ------------------------------------------------------------------------- */
/* Field MLP                                                                 */
/* ------------------------------------------------------------------------- */

/*
 * tcnn SphericalHarmonics degree 4 on d01 = (dir+1)/2 (gfnerf/nerfacto_field.py:64-70,
 * 152-158, 521): tiny-cuda-nn (un-vendored; Dockerfile pins v1.6) maps [0,1] back
 * to [-1,1] and evaluates its real-SH table, emitting fp16.  Restated from the
 * published tcnn spherical_harmonics.h; PARITY UNPINNED (the reference tests
 * check the output shape only, tests/field_components/test_encodings.py:124-139).
 */
void orc_sh4(const float* dir_unit, float* out16) {
  float x = ((dir_unit[0] + 1.f) * .5f) * 2.f - 1.f;
  float y = ((dir_unit[1] + 1.f) * .5f) * 2.f - 1.f;
  float z = ((dir_unit[2] + 1.f) * .5f) * 2.f - 1.f;
  float xy = x * y, xz = x * z, yz = y * z, x2 = x * x, y2 = y * y, z2 = z * z;
  float o[16];
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
  for (int i = 0; i < 16; i++) out16[i] = h2f(f2h(o[i]));
}

int64_t orc_mlp_param_count(int H) {
  return (int64_t)H * 32 + H + 16 * H + 16 + (int64_t)H * 63 + H + (int64_t)H * H + H + 3 * H + 3;
}

typedef struct {
  const float *w0, *b0, *w1, *b1, *w2, *b2, *w3, *b3, *w4, *b4;
} mlp_params;

static mlp_params mlp_split(const float* p, int H) {
  mlp_params m;
  m.w0 = p; p += H * 32;
  m.b0 = p; p += H;
  m.w1 = p; p += 16 * H;
  m.b1 = p; p += 16;
  m.w2 = p; p += H * 63;
  m.b2 = p; p += H;
  m.w3 = p; p += H * H;
  m.b3 = p; p += H;
  m.w4 = p; p += 3 * H;
  m.b4 = p;
  return m;
}

static inline void linear(const float* w, const float* b, int out, int in, const float* x, float* y) {
  for (int o = 0; o < out; o++) {
    double acc = b[o];
    for (int i = 0; i < in; i++) acc += (double)w[o * in + i] * (double)x[i];
    y[o] = (float)acc;
  }
}

/*
 * GFNeRFField.get_density + get_outputs, init stage
 * (gfnerf/nerfacto_field.py:437-507, 509-591) with MLPNetwork (gfnerf/mlp.py:45-57)
 * in fp32: h = base(feat); sigma = exp(h0 + 1); rgb = sigmoid(head(cat[SH, h[1:16], emb])).
 * acts (optional) [n, 3H+63] saves hidden activations for the backward.
 */
void orc_mlp_forward(int64_t n, int H, const float* params, const float* feat, const int32_t* ray_id,
                     const float* ray_dirs, const float* ray_emb, float* sigma, float* rgb) {
  mlp_params m = mlp_split(params, H);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    float h1[256], h[16], in2[63], h2[256], h3[256], o[3];
    linear(m.w0, m.b0, H, 32, feat + 32 * i, h1);
    for (int k = 0; k < H; k++) h1[k] = h1[k] > 0 ? h1[k] : 0;
    linear(m.w1, m.b1, 16, H, h1, h);
    sigma[i] = expf(h[0] + 1.f);
    int32_t r = ray_id[i];
    orc_sh4(ray_dirs + 3 * r, in2);
    for (int k = 0; k < 15; k++) in2[16 + k] = h[1 + k];
    for (int k = 0; k < 32; k++) in2[31 + k] = ray_emb ? ray_emb[32 * r + k] : 0.f;
    linear(m.w2, m.b2, H, 63, in2, h2);
    for (int k = 0; k < H; k++) h2[k] = h2[k] > 0 ? h2[k] : 0;
    linear(m.w3, m.b3, H, H, h2, h3);
    for (int k = 0; k < H; k++) h3[k] = h3[k] > 0 ? h3[k] : 0;
    linear(m.w4, m.b4, 3, H, h3, o);
    for (int k = 0; k < 3; k++) rgb[3 * i + k] = 1.f / (1.f + expf(-o[k]));
  }
}

/*
 * Autograd of orc_mlp_forward incl. _TruncExp.backward
 * (nerfstudio/field_components/activations.py:33-36: g*exp(clamp(x,-15,15))).
 * d_feat [n,32] fp32 (unscaled), d_params fp64 [param_count] (zeroed here),
 * d_ray_emb fp64 [R,32] (optional, zeroed by caller).
 */
void orc_mlp_backward(int64_t n, int H, const float* params, const float* feat, const int32_t* ray_id,
                      const float* ray_dirs, const float* ray_emb, const float* d_sigma,
                      const float* d_rgb, float* d_feat, double* d_params, double* d_ray_emb) {
  mlp_params m = mlp_split(params, H);
  int64_t np = orc_mlp_param_count(H);
  memset(d_params, 0, sizeof(double) * (size_t)np);
#pragma omp parallel
  {
  double* acc_p = (double*)calloc((size_t)np, sizeof(double));
  double* gw0 = acc_p;
  double* gb0 = gw0 + H * 32;
  double* gw1 = gb0 + H;
  double* gb1 = gw1 + 16 * H;
  double* gw2 = gb1 + 16;
  double* gb2 = gw2 + H * 63;
  double* gw3 = gb2 + H;
  double* gb3 = gw3 + H * H;
  double* gw4 = gb3 + H;
  double* gb4 = gw4 + 3 * H;
#pragma omp for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    float h1[256], h[16], in2[63], h2[256], h3[256], o[3];
    const float* x = feat + 32 * i;
    linear(m.w0, m.b0, H, 32, x, h1);
    for (int k = 0; k < H; k++) h1[k] = h1[k] > 0 ? h1[k] : 0;
    linear(m.w1, m.b1, 16, H, h1, h);
    int32_t r = ray_id[i];
    orc_sh4(ray_dirs + 3 * r, in2);
    for (int k = 0; k < 15; k++) in2[16 + k] = h[1 + k];
    for (int k = 0; k < 32; k++) in2[31 + k] = ray_emb ? ray_emb[32 * r + k] : 0.f;
    linear(m.w2, m.b2, H, 63, in2, h2);
    for (int k = 0; k < H; k++) h2[k] = h2[k] > 0 ? h2[k] : 0;
    linear(m.w3, m.b3, H, H, h2, h3);
    for (int k = 0; k < H; k++) h3[k] = h3[k] > 0 ? h3[k] : 0;
    linear(m.w4, m.b4, 3, H, h3, o);
    double go[3], gh3[256], gh2[256], gin2[63], gh[16], gh1[256];
    for (int k = 0; k < 3; k++) {
      double s = 1.0 / (1.0 + exp(-(double)o[k]));
      go[k] = (double)d_rgb[3 * i + k] * s * (1.0 - s);
    }
    for (int j = 0; j < H; j++) gh3[j] = 0;
    for (int k = 0; k < 3; k++) {
      gb4[k] += go[k];
      for (int j = 0; j < H; j++) {
        gw4[k * H + j] += go[k] * h3[j];
        gh3[j] += go[k] * m.w4[k * H + j];
      }
    }
    for (int j = 0; j < H; j++) {
      if (!(h3[j] > 0)) gh3[j] = 0;
      gh2[j] = 0;
    }
    for (int k = 0; k < H; k++) {
      gb3[k] += gh3[k];
      for (int j = 0; j < H; j++) {
        gw3[k * H + j] += gh3[k] * h2[j];
        gh2[j] += gh3[k] * m.w3[k * H + j];
      }
    }
    for (int j = 0; j < H; j++)
      if (!(h2[j] > 0)) gh2[j] = 0;
    for (int j = 0; j < 63; j++) gin2[j] = 0;
    for (int k = 0; k < H; k++) {
      gb2[k] += gh2[k];
      for (int j = 0; j < 63; j++) {
        gw2[k * 63 + j] += gh2[k] * in2[j];
        gin2[j] += gh2[k] * m.w2[k * 63 + j];
      }
    }
    if (d_ray_emb)
      for (int k = 0; k < 32; k++) {
#pragma omp atomic
        d_ray_emb[32 * r + k] += gin2[31 + k];
      }
    float pre = h[0] + 1.f;
    float cl = pre < -15.f ? -15.f : (pre > 15.f ? 15.f : pre);
    gh[0] = (double)d_sigma[i] * exp((double)cl);
    for (int k = 0; k < 15; k++) gh[1 + k] = gin2[16 + k];
    for (int j = 0; j < H; j++) gh1[j] = 0;
    for (int k = 0; k < 16; k++) {
      gb1[k] += gh[k];
      for (int j = 0; j < H; j++) {
        gw1[k * H + j] += gh[k] * h1[j];
        gh1[j] += gh[k] * m.w1[k * H + j];
      }
    }
    for (int j = 0; j < H; j++)
      if (!(h1[j] > 0)) gh1[j] = 0;
    double gx[32];
    for (int j = 0; j < 32; j++) gx[j] = 0;
    for (int k = 0; k < H; k++) {
      gb0[k] += gh1[k];
      for (int j = 0; j < 32; j++) {
        gw0[k * 32 + j] += gh1[k] * x[j];
        gx[j] += gh1[k] * m.w0[k * 32 + j];
      }
    }
    for (int j = 0; j < 32; j++) d_feat[32 * i + j] = (float)gx[j];
  }
#pragma omp critical
  for (int64_t k = 0; k < np; k++) d_params[k] += acc_p[k];
  free(acc_p);
  }
}

/* ------------------------------------------------------------------------- */
/* Loss + Adam                                                               */
/* ------------------------------------------------------------------------- */

/* CharbonnierLoss, nerfstudio/model_components/losses.py:73-84 (out_norm 'b') */
double orc_charbonnier(int64_t n_rays, const float* rgb, const float* target, float eps, float* g_rgb) {
  double loss = 0;
  for (int64_t i = 0; i < n_rays * 3; i++) {
    double d = (double)rgb[i] - (double)target[i];
    double s = sqrt(d * d + (double)eps * (double)eps);
    loss += s;
    if (g_rgb) g_rgb[i] = (float)(d / s / (double)n_rays);
  }
  return loss / (double)n_rays;
}

/*
 * S3IM (nerfstudio/model_components/losses.py:713-794): the R rays are laid out `repeat` times (once in order, then
 * randomly permuted: index[n_virtual] is that concatenation) as a virtual image [3, patch_h, n_virtual / patch_h];
 * SSIM with a ksize x ksize Gaussian window (sigma 1.5, :726-734), stride `stride`, zero padding (ksize-1)/2
 * (:737-753); loss = 1 - mean(ssim_map).  g_src (optional, [R,3], ACCUMULATED) = mult * dloss/dsrc.
 * Returns mult * loss.  fp64 accumulation.
 */
double orc_s3im(int64_t n_rays, int64_t n_virtual, const int64_t* index, const float* src, const float* tar,
                int patch_h, int ksize, int stride, double mult, float* g_src) {
  const int64_t W = n_virtual / patch_h;
  const int pad = (ksize - 1) / 2;
  const int64_t oh = (patch_h + 2 * pad - ksize) / stride + 1, ow = (W + 2 * pad - ksize) / stride + 1;
  double g1[16], win[16][16];
  double gs = 0;
  for (int x = 0; x < ksize; x++) {
    /* the reference builds the window in fp32 (torch.Tensor, :727-734) */
    g1[x] = (double)(float)exp(-(double)((x - ksize / 2) * (x - ksize / 2)) / (2.0 * 1.5 * 1.5));
    gs += g1[x];
  }
  float g1f[16];
  for (int x = 0; x < ksize; x++) g1f[x] = (float)(g1[x]) / (float)gs;
  for (int y = 0; y < ksize; y++)
    for (int x = 0; x < ksize; x++) win[y][x] = (double)(float)(g1f[y] * g1f[x]);
  const double C1 = 0.01 * 0.01, C2 = 0.03 * 0.03;
  const double n_map = (double)(3 * oh * ow);
  double total = 0;
  for (int c = 0; c < 3; c++)
    for (int64_t oy = 0; oy < oh; oy++)
      for (int64_t ox = 0; ox < ow; ox++) {
        double mu1 = 0, mu2 = 0, e11 = 0, e22 = 0, e12 = 0;
        for (int ky = 0; ky < ksize; ky++)
          for (int kx = 0; kx < ksize; kx++) {
            int64_t h = oy * stride - pad + ky, w = ox * stride - pad + kx;
            if (h < 0 || h >= patch_h || w < 0 || w >= W) continue;
            int64_t r = index[h * W + w];
            double x = src[3 * r + c], y = tar[3 * r + c], wt = win[ky][kx];
            mu1 += wt * x; mu2 += wt * y; e11 += wt * x * x; e22 += wt * y * y; e12 += wt * x * y;
          }
        double s11 = e11 - mu1 * mu1, s22 = e22 - mu2 * mu2, s12 = e12 - mu1 * mu2;
        double A1 = 2 * mu1 * mu2 + C1, A2 = 2 * s12 + C2, B1 = mu1 * mu1 + mu2 * mu2 + C1, B2 = s11 + s22 + C2;
        total += A1 * A2 / (B1 * B2);
        if (g_src)
          for (int ky = 0; ky < ksize; ky++)
            for (int kx = 0; kx < ksize; kx++) {
              int64_t h = oy * stride - pad + ky, w = ox * stride - pad + kx;
              if (h < 0 || h >= patch_h || w < 0 || w >= W) continue;
              int64_t r = index[h * W + w];
              double x = src[3 * r + c], y = tar[3 * r + c], wt = win[ky][kx];
              double dA1 = 2 * mu2 * wt, dA2 = 2 * (wt * y - mu2 * wt), dB1 = 2 * mu1 * wt,
                     dB2 = 2 * wt * x - 2 * mu1 * wt;
              double d = (dA1 * A2 + A1 * dA2) / (B1 * B2) - A1 * A2 * (dB1 * B2 + B1 * dB2) / (B1 * B2 * B1 * B2);
              g_src[3 * r + c] += (float)(-mult * d / n_map);
            }
      }
  return mult * (1.0 - total / n_map);
}

/* torch.optim.Adam single-tensor step (no amsgrad / weight decay / maximize):
 * m = b1 m + (1-b1) g; v = b2 v + (1-b2) g^2;
 * p -= lr/(1-b1^t) * m / (sqrt(v)/sqrt(1-b2^t) + eps)       (torch/optim/adam.py) */
void orc_adam_step(int64_t n, float* param, const float* grad, float* exp_avg, float* exp_avg_sq,
                   float lr, float beta1, float beta2, float eps, int64_t step) {
  double bc1 = 1.0 - pow((double)beta1, (double)step);
  double bc2 = 1.0 - pow((double)beta2, (double)step);
  double step_size = (double)lr / bc1;
  double bc2_sqrt = sqrt(bc2);
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n; i++) {
    float g = grad[i];
    float m = exp_avg[i] + (g - exp_avg[i]) * (1.f - beta1); /* lerp_ */
    float v = exp_avg_sq[i] * beta2 + (1.f - beta2) * g * g;
    exp_avg[i] = m;
    exp_avg_sq[i] = v;
    double denom = sqrt((double)v) / bc2_sqrt + (double)eps;
    param[i] = (float)((double)param[i] - step_size * ((double)m / denom));
  }
}

/* ---- ray generation ---------------------------------------------------------------------------
 * Cameras.generate_rays for PERSPECTIVE cameras without distortion (reference
 * nerfstudio/cameras/cameras.py:583-727, the path GF-NeRF's datamanager takes): per ray its camera index and pixel
 * coordinates (y, x) -> origin c2w[:, 3], unit direction, lookat direction c2w[:, 2] (GF-NeRF's addition, :704),
 * pixel_area = dx * dy from the two one-pixel-offset directions (:711-716), directions_norm.
 * torch evaluates every op separately in fp32: no contraction; sums over the last dim of 3 left to right. */
static inline float sum3(float a, float b, float c) { return (a + b) + c; }

void orc_generate_rays(int64_t n_rays, const int64_t* cam_idx, const float* coords_yx, const float* c2w,
                       const float* fx, const float* fy, const float* cx, const float* cy, float* origins,
                       float* directions, float* lookat, float* pixel_area, float* dir_norm) {
#pragma omp parallel for schedule(static)
  for (int64_t i = 0; i < n_rays; i++) {
    const int64_t c = cam_idx[i];
    const float y = coords_yx[2 * i], x = coords_yx[2 * i + 1];
    const float* m = c2w + c * 12;
    /* :606-608 image-plane coordinates of the pixel and of its +1 neighbours in x and in y */
    const float cam[3][2] = {{(x - cx[c]) / fx[c], -(y - cy[c]) / fy[c]},
                             {((x - cx[c]) + 1.f) / fx[c], -(y - cy[c]) / fy[c]},
                             {(x - cx[c]) / fx[c], -((y - cy[c]) + 1.f) / fy[c]}};
    float d[3][3];
    for (int k = 0; k < 3; k++) {
      /* :697-699 sum(dir[None, :] * rotation, -1), then normalize_with_norm (camera_utils.py:240-252) */
      float w[3];
      for (int r = 0; r < 3; r++) w[r] = sum3(cam[k][0] * m[4 * r], cam[k][1] * m[4 * r + 1], -1.f * m[4 * r + 2]);
      /* torch.linalg.vector_norm accumulates x*x with FMAs (checked against the reference's output bit for bit) */
      float nrm = sqrtf(fmaf(w[2], w[2], fmaf(w[1], w[1], w[0] * w[0])));
      if (nrm < 8.8817842e-16f) nrm = 8.8817842e-16f; /* _EPS = 4 eps(float64), camera_utils.py:28 */
      for (int r = 0; r < 3; r++) d[k][r] = w[r] / nrm;
      if (k == 0 && dir_norm) dir_norm[i] = nrm;
    }
    for (int r = 0; r < 3; r++) {
      origins[3 * i + r] = m[4 * r + 3];
      directions[3 * i + r] = d[0][r];
      lookat[3 * i + r] = m[4 * r + 2];
    }
    float ex[3], ey[3];
    for (int r = 0; r < 3; r++) {
      ex[r] = d[0][r] - d[1][r];
      ey[r] = d[0][r] - d[2][r];
    }
    const float dx = sqrtf(sum3(ex[0] * ex[0], ex[1] * ex[1], ex[2] * ex[2]));
    const float dy = sqrtf(sum3(ey[0] * ey[0], ey[1] * ey[1], ey[2] * ey[2]));
    pixel_area[i] = dx * dy;
  }
}

/* ---- cold sampler queries --------------------------------------------------------------------
 * PersSampler::GetPointsAnchors (GetRaysTreeNodesIntersectsKernel + GetTreeNodeIdxFromTsKernel,
 * PersSampler_cuda.cu:799-853, 924-980): for every sample its leaf; where several leaves contain t (a sample exactly
 * on a shared face) the reference's stores race -- the largest node index is taken here. */
void orc_points_anchors(int64_t n_rays, int64_t n_pts_per_ray, const float* rays_o, const float* rays_d,
                        const float* t_cur, const void* tree_nodes_blob, int64_t n_nodes, int64_t* anchors) {
  const tree_node* nodes = (const tree_node*)tree_nodes_blob;
  for (int64_t i = 0; i < n_rays * n_pts_per_ray; i++) anchors[i] = -1;
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < n_rays; r++) {
    for (int64_t u = 0; u < n_nodes; u++) {
      if (!nodes[u].is_leaf_node) continue;
      float near = -1e6f, far = 1e6f;
      get_intersection(rays_o + 3 * r, rays_d + 3 * r, nodes[u].center, nodes[u].side_len, &near, &far);
      if (far <= near) continue;
      for (int64_t i = 0; i < n_pts_per_ray; i++) {
        const float t = t_cur[r * n_pts_per_ray + i];
        if (t >= near && t <= far && u > anchors[r * n_pts_per_ray + i]) anchors[r * n_pts_per_ray + i] = u;
      }
    }
  }
}

/* GetEdgeSamplesKernel (:479-495) over 64-byte EdgePool records (PersSampler.h:50-56) */
void orc_edge_samples(int64_t n_pts, const void* edge_pool_blob, const void* pers_trans_blob, const int64_t* edge_idx,
                      const float* edge_coords, float* out_pts, int64_t* out_idx) {
  const trans_info* transes = (const trans_info*)pers_trans_blob;
  for (int64_t i = 0; i < n_pts; i++) {
    const char* e = (const char*)edge_pool_blob + edge_idx[i] * 64;
    int64_t ab[2];
    memcpy(ab, e, 16);
    const float* f = (const float*)(e + 16);
    float p[3];
    for (int k = 0; k < 3; k++) p[k] = fmaf(f[6 + k], edge_coords[2 * i + 1], fmaf(f[3 + k], edge_coords[2 * i], f[k]));
    for (int s = 0; s < 2; s++) {
      query_frame_transform(transes + ab[s], p, out_pts + (2 * i + s) * 3);
      out_idx[2 * i + s] = ab[s];
    }
  }
}

/* ---- octree maintenance between two ProcOctree calls ------------------------------------------
 * CheckVisible + MarkInvisibleNodesKernel (PersSampler_cuda.cu:680-723): a node whose bounding sphere (radius
 * 0.707 * side_len -- a DOUBLE literal, so the product is taken in double and rounded once) is seen by no camera gets
 * trans_idx = -1.  The contraction spelled out below is the one nvcc applies to the reference's expressions (read off
 * the SASS of oracle/_ref/libgf_ref_cuda.so): the 4-term rows of w2c * center.homogeneous() as
 * fma(x, m0, y * m1) + fma(z, m2, m3); the squared norm as fma(x, x, fma(y, y, z * z)); bias_x only ever fused into the
 * two sums that use it, bias_y rounded.  w2c [n_cams,3,4], intri [n_cams,3,3], bounds [n_cams,2], row major. */
static int check_visible(const float* c, float side_len, const float* K, const float* m, const float* bound) {
  const float radius = (float)((double)side_len * 0.707);
  const float x = fmaf(c[0], m[0], c[1] * m[1]) + fmaf(c[2], m[2], m[3]);
  const float y = fmaf(c[0], m[4], c[1] * m[5]) + fmaf(c[2], m[6], m[7]);
  const float z = fmaf(c[0], m[8], c[1] * m[9]) + fmaf(c[2], m[10], m[11]);
  if (-z < bound[0] - radius || -z > bound[1] + radius) return 0;
  if (sqrtf(fmaf(x, x, fmaf(y, y, z * z))) < radius) return 1;
  const float cx = K[2], cy = K[5], fx = K[0], fy = K[4];
  const float q = radius / -z;
  const float bias_y = q * fy;
  const float img_x = x / -z * fx, img_y = y / -z * fy;
  if (fmaf(q, fx, img_x) < -cx || img_x > fmaf(q, fx, cx) || img_y + bias_y < -cy || img_y > cy + bias_y) return 0;
  return 1;
}

void orc_mark_invisible_nodes(int64_t n_nodes, int64_t n_cams, void* tree_nodes_blob, const float* intri,
                              const float* w2c, const float* bounds) {
  tree_node* nodes = (tree_node*)tree_nodes_blob;
#pragma omp parallel for schedule(static)
  for (int64_t u = 0; u < n_nodes; u++) {
    int64_t n_visible = 0;
    for (int64_t k = 0; k < n_cams; k++)
      n_visible += check_visible(nodes[u].center, nodes[u].side_len, intri + 9 * k, w2c + 12 * k, bounds + 2 * k);
    if (n_visible < 1) nodes[u].trans_idx = -1;
  }
}

/* SetBlockIdxsNearestKernel (:746-766): fp32 difference and norm, compared in double against a minimum that starts at
 * 1e9 with a strict `<` (the first of equal minima; -1 when nothing is closer than 1e9). */
void orc_set_block_idxs(int64_t n_nodes, int64_t n_blocks, void* tree_nodes_blob, const float* centers) {
  tree_node* nodes = (tree_node*)tree_nodes_blob;
  for (int64_t u = 0; u < n_nodes; u++) {
    double min_dist = 1e+9;
    int64_t best = -1;
    for (int64_t b = 0; b < n_blocks; b++) {
      const float dx = nodes[u].center[0] - centers[3 * b], dy = nodes[u].center[1] - centers[3 * b + 1],
                  dz = nodes[u].center[2] - centers[3 * b + 2];
      const double d = (double)sqrtf(fmaf(dx, dx, fmaf(dy, dy, dz * dz)));
      if (d < min_dist) min_dist = d, best = b;
    }
    nodes[u].block_idx = best;
  }
}
