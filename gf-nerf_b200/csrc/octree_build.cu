// PersOctree construction on the host: the octree over the camera rig and the perspective-warp transform of every
// valid leaf, as the reference's 128-byte TreeNode / 576-byte TransInfo blobs.
//
// Replaces PersOctree::PersOctree, GetVisiCams, DistanceSummary, ConstructTreeNode, PCA and ConstructTrans of the
// reference (gfnerf/bindings/PtsSampler/PersSampler.cpp:12-26, 45-88, 92-152, 516-610, 613-831) -- host C++ in the
// reference too (torch CPU ops + Eigen; neither is used here).  Same decisions in the same order:
//   * a cell is split while at least N_PROS / 2 cameras see it and their distance summary (geometric mean of the
//     nearest quarter) is below split_dist_thres cell sides; children are created in child-index order, depth first;
//   * a camera sees a cell if one ray of its 128 x ~72 pixel grid crosses the cell inside the camera's [near, far];
//   * a valid leaf gets a transform: six well-spread cameras (farthest-point selection on the unit sphere around the
//     cell, first one drawn at random) are re-aimed at the cell centre, their 12 (x/z, y/z) projections of 32^3
//     random points of the cell go through a PCA, and the three leading components, normalised by the mean inverse
//     Jacobian, are the 3 x 12 mixing weights.
// The random draws replay numpy's legacy RandomState (MT19937) so that this builder and the numpy restatement
// (gfnerf_b200/persoctree.py, the one the committed fixtures were built with) see the same sample points and the
// same first camera; eigenvectors are unique up to sign, which is all a PCA defines.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include "common.cuh"

namespace gf {
namespace {

constexpr int kNPros = GF_N_PROS;  // 12
constexpr int kNVirt = kNPros / 2;

struct TreeNodeB {
  float center[3];
  float side_len;
  int64_t parent;
  int64_t childs[8];
  uint8_t is_leaf;
  uint8_t pad0[7];
  int64_t trans_idx;
  int64_t block_idx;
  uint8_t pad1[16];
};
static_assert(sizeof(TreeNodeB) == GF_TREE_NODE_BYTES, "TreeNode blob layout");
struct TransInfoB {
  float w2xz[kNPros][2][4];
  float weight[3][kNPros];
  float center[3];
  float side_len;
  float dis_summary;
  uint8_t pad[28];
};
static_assert(sizeof(TransInfoB) == GF_TRANS_INFO_BYTES, "TransInfo blob layout");

// numpy.random.RandomState(seed): MT19937, init_genrand seeding, random_sample = 53-bit doubles, randint = masked
// rejection on 32-bit draws
struct NumpyRandomState {
  uint32_t mt[624];
  int idx;
  explicit NumpyRandomState(uint32_t seed) {
    mt[0] = seed;
    for (int i = 1; i < 624; i++) mt[i] = 1812433253u * (mt[i - 1] ^ (mt[i - 1] >> 30)) + (uint32_t)i;
    idx = 624;
  }
  uint32_t next() {
    if (idx >= 624) {
      for (int k = 0; k < 624; k++) {
        const uint32_t y = (mt[k] & 0x80000000u) | (mt[(k + 1) % 624] & 0x7fffffffu);
        mt[k] = mt[(k + 397) % 624] ^ (y >> 1) ^ ((y & 1u) ? 0x9908b0dfu : 0u);
      }
      idx = 0;
    }
    uint32_t y = mt[idx++];
    y ^= y >> 11;
    y ^= (y << 7) & 0x9d2c5680u;
    y ^= (y << 15) & 0xefc60000u;
    y ^= y >> 18;
    return y;
  }
  double rand() {
    const uint32_t a = next() >> 5, b = next() >> 6;
    return (a * 67108864.0 + b) / 9007199254740992.0;
  }
  uint32_t randint(uint32_t n) {  // uniform in [0, n)
    const uint32_t rng = n - 1;
    if (rng == 0) return 0;
    uint32_t mask = rng;
    mask |= mask >> 1;
    mask |= mask >> 2;
    mask |= mask >> 4;
    mask |= mask >> 8;
    mask |= mask >> 16;
    uint32_t v;
    while ((v = next() & mask) > rng) {
    }
    return v;
  }
};

inline float norm3f(const float* v) {
  float s = v[0] * v[0];
  s += v[1] * v[1];
  s += v[2] * v[2];
  return sqrtf(s);
}

// DistanceSummary (:12-26): exp(mean of the log distances below their 25 % quantile)
float distance_summary(const std::vector<float>& dis) {
  if (dis.empty()) return 1e8f;
  std::vector<float> lg(dis.size());
  for (size_t i = 0; i < dis.size(); i++) lg[i] = logf(dis[i]);
  std::vector<double> sorted(lg.begin(), lg.end());
  std::sort(sorted.begin(), sorted.end());
  const double pos = 0.25 * (double)(sorted.size() - 1);
  const size_t lo = (size_t)floor(pos), hi = std::min(lo + 1, sorted.size() - 1);
  const double t = pos - (double)lo, a = sorted[lo], b = sorted[hi];
  const double q = t >= 0.5 ? b - (b - a) * (1.0 - t) : a + (b - a) * t;  // numpy's _lerp
  const float thres = (float)q;
  double sum = 0.0;
  size_t cnt = 0;
  for (float v : lg)
    if (v < thres) {
      sum += v;
      cnt++;
    }
  if (cnt == 0) {
    for (float v : lg) sum += v;
    cnt = lg.size();
  }
  return expf((float)(sum / (double)cnt));
}

// symmetric eigen-decomposition (cyclic Jacobi), n <= 12: eigenvalues in w, eigenvectors in the columns of V
void jacobi_eigh(int n, double* A, double* w, double* V) {
  for (int i = 0; i < n; i++)
    for (int j = 0; j < n; j++) V[i * n + j] = i == j ? 1.0 : 0.0;
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0;
    for (int i = 0; i < n; i++)
      for (int j = i + 1; j < n; j++) off += A[i * n + j] * A[i * n + j];
    if (off < 1e-300) break;
    for (int p = 0; p < n; p++)
      for (int q = p + 1; q < n; q++) {
        const double apq = A[p * n + q];
        if (fabs(apq) < 1e-300) continue;
        const double theta = (A[q * n + q] - A[p * n + p]) / (2.0 * apq);
        const double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
        const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
        for (int k = 0; k < n; k++) {
          const double akp = A[k * n + p], akq = A[k * n + q];
          A[k * n + p] = c * akp - s * akq;
          A[k * n + q] = s * akp + c * akq;
        }
        for (int k = 0; k < n; k++) {
          const double apk = A[p * n + k], aqk = A[q * n + k];
          A[p * n + k] = c * apk - s * aqk;
          A[q * n + k] = s * apk + c * aqk;
        }
        for (int k = 0; k < n; k++) {
          const double vkp = V[k * n + p], vkq = V[k * n + q];
          V[k * n + p] = c * vkp - s * vkq;
          V[k * n + q] = s * vkp + c * vkq;
        }
      }
  }
  for (int i = 0; i < n; i++) w[i] = A[i * n + i];
}

bool inv3(const double m[9], double out[9]) {
  const double c0 = m[4] * m[8] - m[5] * m[7], c1 = m[5] * m[6] - m[3] * m[8], c2 = m[3] * m[7] - m[4] * m[6];
  const double det = m[0] * c0 + m[1] * c1 + m[2] * c2;
  if (det == 0.0 || !std::isfinite(det)) return false;
  const double id = 1.0 / det;
  out[0] = c0 * id;
  out[1] = (m[2] * m[7] - m[1] * m[8]) * id;
  out[2] = (m[1] * m[5] - m[2] * m[4]) * id;
  out[3] = c1 * id;
  out[4] = (m[0] * m[8] - m[2] * m[6]) * id;
  out[5] = (m[2] * m[3] - m[0] * m[5]) * id;
  out[6] = c2 * id;
  out[7] = (m[1] * m[6] - m[0] * m[7]) * id;
  out[8] = (m[0] * m[4] - m[1] * m[3]) * id;
  return true;
}

struct Builder {
  int64_t max_depth;
  float split_dist_thres;
  int64_t n_cams, n_rand_pts;
  const float *c2w, *intri, *bound;  // [n,3,4], [n,3,3], [n,2]
  NumpyRandomState rng;
  std::vector<float> rays_d;         // [n_cams][n_pix][3]
  std::vector<float> cam_pos;        // [n_cams][3]
  int64_t n_pix = 0;
  double half_diag_fov = 0.0;
  std::vector<TreeNodeB> nodes;
  std::vector<TransInfoB> trans;
  const char* error = nullptr;

  Builder(uint32_t seed) : rng(seed) {}

  // GetVisiCams (:51-66): the pixel-grid ray directions of every camera, in world space
  void setup_visibility_rays(int64_t res_w) {
    const float* it = intri;
    const double half_w = it[2], half_h = it[5];
    const float cx = it[2], cy = it[5], fx = it[0], fy = it[4];
    const int64_t res_h = (int64_t)nearbyint((double)res_w / half_w * half_h);
    auto linspace = [](double start, double stop, int64_t num, std::vector<float>& out) {
      out.resize((size_t)num);
      const double step = num > 1 ? (stop - start) / (double)(num - 1) : 0.0;
      for (int64_t k = 0; k < num; k++) out[(size_t)k] = (float)((double)k * step + start);
      if (num > 1) out[(size_t)num - 1] = (float)stop;
    };
    std::vector<float> vi, vj;
    linspace(.5, half_h * 2. - .5, res_h, vi);
    linspace(.5, half_w * 2. - .5, res_w, vj);
    n_pix = res_h * res_w;
    std::vector<float> cam((size_t)n_pix * 3);
    for (int64_t a = 0; a < res_h; a++)
      for (int64_t b = 0; b < res_w; b++) {
        float* c = &cam[(size_t)(a * res_w + b) * 3];
        c[0] = (vj[(size_t)b] - cx) / fx;
        c[1] = -(vi[(size_t)a] - cy) / fy;
        c[2] = -1.f;
      }
    rays_d.resize((size_t)n_cams * n_pix * 3);
    cam_pos.resize((size_t)n_cams * 3);
    for (int64_t n = 0; n < n_cams; n++) {
      const float* m = c2w + n * 12;
      for (int k = 0; k < 3; k++) cam_pos[(size_t)n * 3 + k] = m[4 * k + 3];
      for (int64_t p = 0; p < n_pix; p++) {
        const float* c = &cam[(size_t)p * 3];
        float* d = &rays_d[((size_t)n * n_pix + p) * 3];
        for (int i = 0; i < 3; i++) {
          float s = m[4 * i] * c[0];
          s += m[4 * i + 1] * c[1];
          s += m[4 * i + 2] * c[2];
          d[i] = s;
        }
      }
    }
    const double corner[3] = {half_w / fx, half_h / fy, 1.0};
    half_diag_fov = acos(1.0 / sqrt(corner[0] * corner[0] + corner[1] * corner[1] + 1.0)) + 1e-3;
  }

  // one grid ray of camera `cam` through the box [lo, hi] inside [near, far]?  (:67-86)
  bool camera_sees(int64_t cam, const float lo[3], const float hi[3]) const {
    const float* o = &cam_pos[(size_t)cam * 3];
    const float bn = bound[cam * 2], bf = bound[cam * 2 + 1];
    const float* d = &rays_d[(size_t)cam * n_pix * 3];
    auto fix = [](float v) { return std::isnan(v) ? 0.f : (std::isinf(v) ? (v > 0 ? 1e6f : -1e6f) : v); };
    for (int64_t p = 0; p < n_pix; p++, d += 3) {
      float far = INFINITY, near = -INFINITY;
      for (int k = 0; k < 3; k++) {
        const float a = fix((lo[k] - o[k]) / d[k]), b = fix((hi[k] - o[k]) / d[k]);
        far = fminf(far, fmaxf(a, b));
        near = fmaxf(near, fminf(a, b));
      }
      far = fminf(far, bf);
      near = fmaxf(near, bn);
      if (far > near) return true;
    }
    return false;
  }

  std::vector<int64_t> visible_cams(float side_len, const float center[3], const std::vector<int64_t>& candidates) const {
    std::vector<int64_t> out;
    float lo[3], hi[3];
    for (int k = 0; k < 3; k++) {
      lo[k] = center[k] - side_len * .5f;
      hi[k] = center[k] + side_len * .5f;
    }
    const double radius = (double)side_len * 0.8660254 + 1e-6;
    std::vector<char> seen(candidates.size(), 0);
    // cameras are independent: flags in parallel, gathered in candidate order (same list whatever the thread count)
#pragma omp parallel for schedule(dynamic, 4)
    for (int64_t ci = 0; ci < (int64_t)candidates.size(); ci++) {
      const int64_t cam = candidates[(size_t)ci];
      // conservative view-cone reject (a pure speed-up: a rejected camera has no grid ray through the cell)
      const float* m = c2w + cam * 12;
      double rel[3], dist = 0.0, dot = 0.0;
      for (int k = 0; k < 3; k++) {
        rel[k] = (double)center[k] - (double)cam_pos[(size_t)cam * 3 + k];
        dist += rel[k] * rel[k];
        dot += rel[k] * -(double)m[4 * k + 2];
      }
      dist = sqrt(dist);
      if (dist > radius * 1.0001) {
        const double ang = acos(std::max(-1.0, std::min(1.0, dot / dist))) - asin(std::min(1.0, radius / dist));
        if (ang > half_diag_fov + 1e-4) continue;
      }
      if (camera_sees(cam, lo, hi)) seen[(size_t)ci] = 1;
    }
    for (size_t ci = 0; ci < candidates.size(); ci++)
      if (seen[ci]) out.push_back(candidates[ci]);
    return out;
  }

  // ConstructTrans (:613-831)
  bool construct_trans(const std::vector<float>& rand_pts, const std::vector<int64_t>& cams, const float center[3],
                       TransInfoB& out) {
    const int n_cur = (int)cams.size();
    std::vector<float> pos((size_t)n_cur * 3), dis((size_t)n_cur), normed((size_t)n_cur * 3);
    std::vector<double> axes((size_t)n_cur * 9);
    for (int i = 0; i < n_cur; i++) {
      const float* m = c2w + cams[(size_t)i] * 12;
      double r[9];
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) r[3 * a + b] = m[4 * a + b];
      if (!inv3(r, &axes[(size_t)i * 9])) return false;
      for (int k = 0; k < 9; k++) axes[(size_t)i * 9 + k] = (double)(float)axes[(size_t)i * 9 + k];
      float rel[3];
      for (int k = 0; k < 3; k++) {
        pos[(size_t)i * 3 + k] = m[4 * k + 3];
        rel[k] = m[4 * k + 3] - center[k];
      }
      dis[(size_t)i] = norm3f(rel);
      for (int k = 0; k < 3; k++) normed[(size_t)i * 3 + k] = rel[k] / dis[(size_t)i];
    }
    const float dis_summary = distance_summary(dis);
    // farthest-point selection of the virtual cameras (:652-673), the first one at random
    std::vector<int> good;
    std::vector<char> marks((size_t)n_cur, 0);
    good.push_back((int)rng.randint((uint32_t)n_cur));
    marks[(size_t)good[0]] = 1;
    const int n_pick = std::min(kNVirt, n_cur);
    for (int it = 1; it < n_pick; it++) {
      int candi = -1;
      float best = -INFINITY;
      for (int i = 0; i < n_cur; i++) {
        float cur = 1e8f;
        if (marks[(size_t)i]) {
          cur = -2.f;
        } else {
          for (int j = 0; j < n_cur; j++)
            if (marks[(size_t)j]) {
              float df[3];
              for (int k = 0; k < 3; k++) df[k] = normed[(size_t)j * 3 + k] - normed[(size_t)i * 3 + k];
              cur = fminf(cur, norm3f(df));
            }
        }
        if (cur > best) {  // first maximum
          best = cur;
          candi = i;
        }
      }
      marks[(size_t)candi] = 1;
      good.push_back(candi);
    }
    for (int i = 0; (int)good.size() < kNVirt; i++) good.push_back(good[(size_t)i]);

    double frame[kNPros][2][4];
    for (int k = 0; k < kNVirt; k++) {
      const int g = good[(size_t)k];
      const double d = dis[(size_t)g];
      const double scale = std::min(std::max(d / (double)dis_summary, 1.0), 1e9);
      const double reach = std::min(std::max(d, (double)dis_summary), 1e9);
      double rel[3], gpos[3], ez[3], nrm = 0.0;
      for (int c = 0; c < 3; c++) {
        rel[c] = ((double)pos[(size_t)g * 3 + c] - (double)center[c]) / d * reach;
        gpos[c] = rel[c] + (double)center[c];
        nrm += rel[c] * rel[c];
      }
      nrm = sqrt(nrm);
      for (int c = 0; c < 3; c++) ez[c] = rel[c] / nrm;
      // rotate the camera so that its z axis looks along ez (Eigen::AngleAxisf, :741)
      const double* ax = &axes[(size_t)g * 9];
      const double fz[3] = {ax[6], ax[7], ax[8]};
      double cr[3] = {fz[1] * ez[2] - fz[2] * ez[1], fz[2] * ez[0] - fz[0] * ez[2], fz[0] * ez[1] - fz[1] * ez[0]};
      const double cn = sqrt(cr[0] * cr[0] + cr[1] * cr[1] + cr[2] * cr[2]);
      const double cos_v = std::max(-0.999999, std::min(0.999999, fz[0] * ez[0] + fz[1] * ez[1] + fz[2] * ez[2]));
      const double sin_v = std::max(-0.999999, std::min(0.999999, cn));
      double angle = asin(sin_v);
      if (cos_v < 0) angle = M_PI - angle;
      if (cn > 0)
        for (int c = 0; c < 3; c++) cr[c] /= cn;
      const double K[9] = {0, -cr[2], cr[1], cr[2], 0, -cr[0], -cr[1], cr[0], 0};
      double K2[9], R[9];
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          double s = 0;
          for (int c = 0; c < 3; c++) s += K[3 * a + c] * K[3 * c + b];
          K2[3 * a + b] = s;
        }
      for (int a = 0; a < 9; a++) R[a] = (a % 4 == 0 ? 1.0 : 0.0) + sin(angle) * K[a] + (1 - cos(angle)) * K2[a];
      double ga[9];  // good_axis . R^T
      for (int a = 0; a < 3; a++)
        for (int b = 0; b < 3; b++) {
          double s = 0;
          for (int c = 0; c < 3; c++) s += ax[3 * a + c] * R[3 * b + c];
          ga[3 * a + b] = s;
        }
      const double focal = (double)(float)(intri[0] / intri[2]);
      for (int half = 0; half < 2; half++) {  // x-projection k, y-projection k + 6
        const int p = k + half * kNVirt;
        double xw = 0, zw = 0;
        for (int c = 0; c < 3; c++) {
          frame[p][0][c] = ga[3 * half + c] * focal * scale;
          frame[p][1][c] = ga[6 + c];
          xw += frame[p][0][c] * gpos[c];
          zw += frame[p][1][c] * gpos[c];
        }
        frame[p][0][3] = -xw;
        frame[p][1][3] = -zw;
      }
    }
    // the 12 projections of the sample points and their Jacobians (:779-812), PCA of the projections
    const int64_t n = (int64_t)rand_pts.size() / 3;
    std::vector<double> v((size_t)n * kNPros), dvd((size_t)n * kNPros * 3);
    // The sums over the sample points are taken over kChunks fixed slices, each sequentially, and the slices are
    // combined in order: the result does not depend on the number of threads.
    constexpr int kChunks = 64;
    const int64_t per = (n + kChunks - 1) / kChunks;
    double mean_c[kChunks][kNPros];
    int bad = 0;
#pragma omp parallel for schedule(static) reduction(| : bad)
    for (int ch = 0; ch < kChunks; ch++) {
      for (int p = 0; p < kNPros; p++) mean_c[ch][p] = 0.0;
      for (int64_t i = ch * per; i < std::min(n, (ch + 1) * per); i++) {
        const float* pt = &rand_pts[(size_t)i * 3];
        for (int p = 0; p < kNPros; p++) {
          double t0 = frame[p][0][3], t1 = frame[p][1][3];
          for (int c = 0; c < 3; c++) {
            t0 += frame[p][0][c] * pt[c];
            t1 += frame[p][1][c] * pt[c];
          }
          if (!(t1 < 0)) bad |= 1;  // :795 a sample point behind a virtual camera
          const double da = 1.0 / t1, db = t0 / -(t1 * t1);
          for (int c = 0; c < 3; c++) dvd[((size_t)i * kNPros + p) * 3 + c] = da * frame[p][0][c] + db * frame[p][1][c];
          v[(size_t)i * kNPros + p] = t0 / t1;
          mean_c[ch][p] += t0 / t1;
        }
      }
    }
    if (bad) return false;
    double mean[kNPros] = {0};
    for (int ch = 0; ch < kChunks; ch++)
      for (int p = 0; p < kNPros; p++) mean[p] += mean_c[ch][p];
    for (int p = 0; p < kNPros; p++) mean[p] /= (double)n;
    std::vector<double> cov_c((size_t)kChunks * kNPros * kNPros, 0.0);
#pragma omp parallel for schedule(static)
    for (int ch = 0; ch < kChunks; ch++) {
      double* cc = &cov_c[(size_t)ch * kNPros * kNPros];
      for (int64_t i = ch * per; i < std::min(n, (ch + 1) * per); i++)
        for (int a = 0; a < kNPros; a++) {
          const double ma = v[(size_t)i * kNPros + a] - mean[a];
          for (int b = a; b < kNPros; b++) cc[a * kNPros + b] += ma * (v[(size_t)i * kNPros + b] - mean[b]);
        }
    }
    double cov[kNPros * kNPros] = {0};
    for (int ch = 0; ch < kChunks; ch++)
      for (int k = 0; k < kNPros * kNPros; k++) cov[k] += cov_c[(size_t)ch * kNPros * kNPros + k];
    for (int a = 0; a < kNPros; a++)
      for (int b = a; b < kNPros; b++) {
        cov[a * kNPros + b] /= (double)n;
        cov[b * kNPros + a] = cov[a * kNPros + b];
      }
    double w[kNPros], V[kNPros * kNPros];
    jacobi_eigh(kNPros, cov, w, V);
    int order[kNPros];
    for (int i = 0; i < kNPros; i++) order[i] = i;
    std::stable_sort(order, order + kNPros, [&](int a, int b) { return w[a] > w[b]; });
    double W[3][kNPros];
    for (int r = 0; r < 3; r++)
      for (int p = 0; p < kNPros; p++) W[r][p] = (double)(float)V[p * kNPros + order[r]];
    // mean over the points of 1 / max_k |d v_k / d warp_c|, the inverse Jacobian through the 3 components
    double step_c[kChunks][3];
    int singular = 0;
#pragma omp parallel for schedule(static) reduction(| : singular)
    for (int ch = 0; ch < kChunks; ch++) {
      step_c[ch][0] = step_c[ch][1] = step_c[ch][2] = 0.0;
      for (int64_t i = ch * per; i < std::min(n, (ch + 1) * per); i++) {
        double jac[9];
        for (int r = 0; r < 3; r++)
          for (int c = 0; c < 3; c++) {
            double s = 0;
            for (int p = 0; p < kNPros; p++) s += W[r][p] * dvd[((size_t)i * kNPros + p) * 3 + c];
            jac[3 * r + c] = s;
          }
        double ji[9];
        if (!inv3(jac, ji)) {
          singular |= 1;
          continue;
        }
        for (int c = 0; c < 3; c++) {
          double mx = 0;
          for (int p = 0; p < kNPros; p++) {
            double s = 0;
            for (int q = 0; q < 3; q++) s += dvd[((size_t)i * kNPros + p) * 3 + q] * ji[3 * q + c];
            mx = std::max(mx, fabs(s));
          }
          step_c[ch][c] += 1.0 / mx;
        }
      }
    }
    if (singular) return false;
    double mean_step[3] = {0, 0, 0};
    for (int ch = 0; ch < kChunks; ch++)
      for (int c = 0; c < 3; c++) mean_step[c] += step_c[ch][c];
    memset(&out, 0, sizeof(out));
    for (int r = 0; r < 3; r++) {
      mean_step[r] /= (double)n;
      for (int p = 0; p < kNPros; p++) {
        out.weight[r][p] = (float)(W[r][p] / mean_step[r]);
        if (!std::isfinite(out.weight[r][p])) return false;
      }
    }
    for (int p = 0; p < kNPros; p++)
      for (int a = 0; a < 2; a++)
        for (int c = 0; c < 4; c++) {
          out.w2xz[p][a][c] = (float)frame[p][a][c];
          if (!std::isfinite(out.w2xz[p][a][c])) return false;
        }
    for (int c = 0; c < 3; c++) out.center[c] = center[c];
    out.dis_summary = dis_summary;
    return true;
  }

  // ConstructTreeNode (:516-591)
  void construct(int64_t u, int64_t depth, const float center[3], float side_len, const std::vector<int64_t>& candidates) {
    if (error) return;
    for (int k = 0; k < 3; k++) nodes[(size_t)u].center[k] = center[k];
    nodes[(size_t)u].side_len = side_len;
    if (depth > max_depth) {
      nodes[(size_t)u].is_leaf = 1;
      return;
    }
    const std::vector<int64_t> visi = visible_cams(side_len, center, candidates);
    std::vector<float> cam_dis(visi.size());
    for (size_t i = 0; i < visi.size(); i++) {
      float rel[3];
      for (int k = 0; k < 3; k++) rel[k] = cam_pos[(size_t)visi[i] * 3 + k] - center[k];
      cam_dis[i] = norm3f(rel);
    }
    const float dsum = distance_summary(cam_dis);
    const bool enough = (int)visi.size() >= kNVirt;
    if (enough && dsum < side_len * split_dist_thres) {
      for (int st = 0; st < 8; st++) {
        const int64_t v = (int64_t)nodes.size();
        TreeNodeB ch;
        memset(&ch, 0, sizeof(ch));
        ch.parent = u;
        for (int k = 0; k < 8; k++) ch.childs[k] = -1;
        ch.trans_idx = -1;
        ch.block_idx = -1;
        nodes.push_back(ch);
        nodes[(size_t)u].childs[st] = v;
        const float off[3] = {float((st >> 2) & 1) - .5f, float((st >> 1) & 1) - .5f, float(st & 1) - .5f};
        float cc[3];
        for (int k = 0; k < 3; k++) cc[k] = center[k] + side_len * .5f * off[k];
        construct(v, depth + 1, cc, side_len * .5f, visi);
      }
    } else if (!enough) {
      nodes[(size_t)u].is_leaf = 1;  // a leaf without a transform: too few cameras see it
    } else {
      nodes[(size_t)u].is_leaf = 1;
      nodes[(size_t)u].trans_idx = (int64_t)trans.size();
      std::vector<float> rand_pts((size_t)n_rand_pts * 3);
      for (int64_t i = 0; i < n_rand_pts; i++)
        for (int k = 0; k < 3; k++) rand_pts[(size_t)i * 3 + k] = ((float)rng.rand() - .5f) * side_len + center[k];
      TransInfoB tr;
      if (!construct_trans(rand_pts, visi, center, tr)) {
        error = "ConstructTrans: degenerate transform (sample points behind a virtual camera or a singular Jacobian)";
        return;
      }
      tr.side_len = side_len;
      trans.push_back(tr);
    }
  }
};

}  // namespace
}  // namespace gf

using namespace gf;

extern "C" {

int gf_octree_build(int64_t max_depth, float bbox_side_len, float split_dist_thres, const float* c2w, const float* intri,
                    const float* bound, int64_t n_cams, uint32_t seed, int64_t n_rand_pts, int64_t visi_res_w,
                    void** handle, int64_t* n_nodes, int64_t* n_trans) {
  GF_REQUIRE(c2w && intri && bound && n_cams > 0 && handle && n_nodes && n_trans, "gf_octree_build: null / empty input");
  GF_REQUIRE(max_depth >= 0 && bbox_side_len > 0.f && n_rand_pts > 0 && visi_res_w > 0, "gf_octree_build: bad sizes");
  Builder* b = new Builder(seed);
  b->max_depth = max_depth;
  b->split_dist_thres = split_dist_thres;
  b->n_cams = n_cams;
  b->n_rand_pts = n_rand_pts;
  b->c2w = c2w;
  b->intri = intri;
  b->bound = bound;
  b->setup_visibility_rays(visi_res_w);
  TreeNodeB root;
  memset(&root, 0, sizeof(root));
  root.parent = -1;
  for (int k = 0; k < 8; k++) root.childs[k] = -1;
  root.trans_idx = -1;
  root.block_idx = -1;
  b->nodes.push_back(root);
  std::vector<int64_t> all((size_t)n_cams);
  for (int64_t i = 0; i < n_cams; i++) all[(size_t)i] = i;
  const float zero[3] = {0.f, 0.f, 0.f};
  b->construct(0, 0, zero, bbox_side_len, all);
  if (b->error) {
    set_error("gf_octree_build: %s", b->error);
    delete b;
    return GF_ERR_INVALID;
  }
  b->rays_d.clear();
  b->rays_d.shrink_to_fit();
  b->c2w = b->intri = b->bound = nullptr;  // the caller's arrays are not kept
  *handle = b;
  *n_nodes = (int64_t)b->nodes.size();
  *n_trans = (int64_t)b->trans.size();
  return GF_OK;
}

int gf_octree_build_fetch(void* handle, void* tree_nodes_out, void* pers_trans_out) {
  GF_REQUIRE(handle, "gf_octree_build_fetch: null handle");
  Builder* b = (Builder*)handle;
  if (tree_nodes_out) memcpy(tree_nodes_out, b->nodes.data(), b->nodes.size() * sizeof(TreeNodeB));
  if (pers_trans_out) memcpy(pers_trans_out, b->trans.data(), b->trans.size() * sizeof(TransInfoB));
  delete b;
  return GF_OK;
}

// children in front-to-back order for each ray octant (PersSampler.cpp:137-151): descending bitrev3(child ^ octant)
int gf_octree_search_order(uint8_t* out64) {
  GF_REQUIRE(out64, "gf_octree_search_order: null output");
  auto bitrev3 = [](int v) { return ((v & 1) << 2) | (v & 2) | ((v >> 2) & 1); };
  for (int st = 0; st < 8; st++) {
    int idx[8];
    for (int i = 0; i < 8; i++) idx[i] = i;
    std::stable_sort(idx, idx + 8, [&](int a, int b) { return bitrev3(a ^ st) > bitrev3(b ^ st); });
    for (int i = 0; i < 8; i++) out64[st * 8 + i] = (uint8_t)idx[i];
  }
  return GF_OK;
}

}  // extern "C"
