// PersOctree maintenance on the host: ProcOctree (compact + subdivide) over the reference's 128-byte TreeNode blob.
//
// Replaces PersOctree::ProcOctree (reference gfnerf/bindings/PtsSampler/PersSampler.cpp:154-417), which is host C++ in
// the reference too (it copies the node blob to the CPU, rebuilds it and uploads it again).  Same node order as the
// reference: compaction keeps the survivors in their old order, subdivision renumbers depth first with the eight new
// children of a split leaf directly behind it.  Plain C++ (no Eigen, no torch); the blob layout is
// PersSampler.h:31-49 as verified in SURVEY.md section 8b.
#include <math.h>
#include <string.h>

#include <vector>

#include "common.cuh"

namespace gf {
namespace {

struct TreeNodeBlob {  // 128 bytes, little-endian
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
static_assert(sizeof(TreeNodeBlob) == GF_TREE_NODE_BYTES, "TreeNode blob layout");

constexpr int64_t kInitNodeStat = 1000;  // INIT_NODE_STAT, PersSampler_cuda.cu:11-17

struct Builder {
  const std::vector<TreeNodeBlob>& src;
  const std::vector<int64_t>&w, &a;
  const int64_t* visit;
  const std::vector<int64_t>& inv_idx;
  bool brute_force;
  std::vector<TreeNodeBlob> out;
  std::vector<int64_t> ow, oa;

  // :318-417 depth-first copy; a visited leaf becomes an interior node with eight leaf children
  int64_t rec(int64_t u, int64_t pa) {
    const int64_t new_u = (int64_t)out.size();
    TreeNodeBlob node;
    memset(&node, 0, sizeof(node));
    const TreeNodeBlob& s = src[(size_t)u];
    memcpy(node.center, s.center, sizeof(node.center));
    node.side_len = s.side_len;
    node.parent = pa;
    memcpy(node.childs, s.childs, sizeof(node.childs));
    node.is_leaf = s.is_leaf;
    node.trans_idx = s.trans_idx;
    node.block_idx = s.block_idx;
    out.push_back(node);
    ow.push_back(w[(size_t)u]);
    oa.push_back(a[(size_t)u]);
    if (node.is_leaf) {
      if (!brute_force && visit[inv_idx[(size_t)u]] <= 4) return new_u;  // :354
      for (int st = 0; st < 8; st++) {
        const float off[3] = {float((st >> 2) & 1) - .5f, float((st >> 1) & 1) - .5f, float(st & 1) - .5f};
        TreeNodeBlob ch;
        memset(&ch, 0, sizeof(ch));
        const float half = node.side_len * .5f;
        for (int k = 0; k < 3; k++) ch.center[k] = node.center[k] + half * off[k];
        ch.side_len = half;
        ch.parent = new_u;
        for (int k = 0; k < 8; k++) ch.childs[k] = -1;
        ch.is_leaf = 1;
        ch.trans_idx = node.trans_idx;
        ch.block_idx = 0;  // left unset by the reference (:376)
        out[(size_t)new_u].childs[st] = (int64_t)out.size();
        out.push_back(ch);
        ow.push_back(ow[(size_t)new_u]);
        oa.push_back(oa[(size_t)new_u]);
      }
      out[(size_t)new_u].is_leaf = 0;
      out[(size_t)new_u].trans_idx = -1;
      ow[(size_t)new_u] = kInitNodeStat;
      oa[(size_t)new_u] = kInitNodeStat;
    } else {
      for (int st = 0; st < 8; st++) {
        const int64_t c = out[(size_t)new_u].childs[st];
        if (c >= 0) {
          const int64_t nc = rec(c, new_u);
          out[(size_t)new_u].childs[st] = nc;
        }
      }
    }
    return new_u;
  }
};

}  // namespace
}  // namespace gf

using namespace gf;

extern "C" int gf_octree_proc(const void* nodes_in, int64_t n_in, const int64_t* weight_stats_in,
                              const int64_t* alpha_stats_in, const int64_t* visit_cnt_in, int compact, int subdivide,
                              int brute_force, void* nodes_out, int64_t* weight_stats_out, int64_t* alpha_stats_out,
                              int64_t capacity, int64_t* n_out) {
  GF_REQUIRE(nodes_in && n_in > 0 && weight_stats_in && alpha_stats_in && visit_cnt_in && n_out,
             "gf_octree_proc: null input / empty tree");
  std::vector<TreeNodeBlob> nb((size_t)n_in);
  memcpy(nb.data(), nodes_in, (size_t)n_in * sizeof(TreeNodeBlob));
  const int64_t n = n_in;
  auto has_child = [&](int64_t u) {
    for (int k = 0; k < 8; k++)
      if (nb[(size_t)u].childs[k] >= 0) return true;
    return false;
  };
  // :170-205 drop invalid leaves from their parents; interior nodes left without children become leaves; repeat
  while (compact) {
    for (int64_t u = 0; u < n; u++) {
      TreeNodeBlob& x = nb[(size_t)u];
      if (!x.is_leaf) continue;
      if (x.trans_idx < 0 && x.parent >= 0) {
        TreeNodeBlob& p = nb[(size_t)x.parent];
        for (int k = 0; k < 8; k++)
          if (p.childs[k] == u) p.childs[k] = -1;
      }
    }
    bool update = false;
    for (int64_t u = 1; u < n; u++) {
      if (!has_child(u)) {
        if (!nb[(size_t)u].is_leaf) update = true;
        nb[(size_t)u].is_leaf = 1;
      }
    }
    if (!update) break;
  }
  // :207-240 splice out chains of single-child interior nodes
  if (compact) {
    auto single_child = [&](int64_t x) {
      int found = -1, cnt = 0;
      for (int k = 0; k < 8; k++)
        if (nb[(size_t)x].childs[k] >= 0) {
          if (cnt++ == 0) found = k;
        }
      return cnt == 1 ? found : -1;
    };
    for (int64_t u = 0; u < n; u++) {
      if (nb[(size_t)u].is_leaf && nb[(size_t)u].trans_idx < 0) continue;
      int64_t v = nb[(size_t)u].parent;
      while (v >= 0 && nb[(size_t)v].parent >= 0 && single_child(v) >= 0) {
        const int64_t vv = nb[(size_t)v].parent;
        for (int k = 0; k < 8; k++)
          if (nb[(size_t)vv].childs[k] == v) nb[(size_t)vv].childs[k] = u;
        nb[(size_t)u].parent = vv;
        nb[(size_t)v].trans_idx = -1;
        nb[(size_t)v].is_leaf = 1;
        v = vv;
      }
    }
  }
  // :242-300 renumber the survivors (interior nodes and valid leaves) in their old order
  std::vector<int64_t> new_idx((size_t)n, -1), inv_idx;
  inv_idx.reserve((size_t)n);
  for (int64_t u = 0; u < n; u++) {
    if (!nb[(size_t)u].is_leaf || nb[(size_t)u].trans_idx >= 0) {
      new_idx[(size_t)u] = (int64_t)inv_idx.size();
      inv_idx.push_back(u);
    }
  }
  GF_REQUIRE(new_idx[0] == 0, "gf_octree_proc: the root was pruned (no valid leaf left in the octree)");
  std::vector<TreeNodeBlob> nn(inv_idx.size());
  std::vector<int64_t> nw(inv_idx.size()), na(inv_idx.size());
  for (size_t i = 0; i < inv_idx.size(); i++) {
    nn[i] = nb[(size_t)inv_idx[i]];  // byte copy: padding travels with the node
    // CHECK_GE(node.parent, 0) / CHECK_GE(node.childs[st], 0), PersSampler.cpp:307,315: the reference aborts when a
    // removed node is still linked, which is what ProcOctree without `compact` does on a tree with pruned leaves
    if (nn[i].parent >= 0) {
      nn[i].parent = new_idx[(size_t)nn[i].parent];
      GF_REQUIRE(nn[i].parent >= 0, "gf_octree_proc: the parent of node %lld was removed (compact = 0 on a pruned tree?)",
                 (long long)inv_idx[i]);
    }
    for (int k = 0; k < 8; k++)
      if (nn[i].childs[k] >= 0) {
        nn[i].childs[k] = new_idx[(size_t)nn[i].childs[k]];
        GF_REQUIRE(nn[i].childs[k] >= 0,
                   "gf_octree_proc: a removed node is still linked from node %lld (compact = 0 on a pruned tree?)",
                   (long long)inv_idx[i]);
      }
    nw[i] = weight_stats_in[inv_idx[i]];
    na[i] = alpha_stats_in[inv_idx[i]];
  }
  if (subdivide) {
    Builder b{nn, nw, na, visit_cnt_in, inv_idx, brute_force != 0, {}, {}, {}};
    b.out.reserve(nn.size() * 2);
    b.rec(0, -1);
    nn.swap(b.out);
    nw.swap(b.ow);
    na.swap(b.oa);
  }
  *n_out = (int64_t)nn.size();
  if (!nodes_out) return GF_OK;  // size query
  GF_REQUIRE(weight_stats_out && alpha_stats_out, "gf_octree_proc: null output");
  GF_REQUIRE(capacity >= (int64_t)nn.size(), "gf_octree_proc: output capacity %lld < %lld nodes", (long long)capacity,
             (long long)nn.size());
  memcpy(nodes_out, nn.data(), nn.size() * sizeof(TreeNodeBlob));
  memcpy(weight_stats_out, nw.data(), nw.size() * sizeof(int64_t));
  memcpy(alpha_stats_out, na.data(), na.size() * sizeof(int64_t));
  return GF_OK;
}

// PersOctree::ConstructEdgePool (PersSampler.cpp:833-893): for every pair of valid leaves, the faces of the smaller
// one whose centre lies on (inside, with 1e-4 slack) the larger one -> a 64-byte EdgePool record
// {t_idx_a, t_idx_b, face centre, the two in-face half-edge vectors} (PersSampler.h:50-56).  Host code, O(leaves^2)
// like the reference.  edges_out == NULL: size query.
extern "C" int gf_octree_edge_pool(const void* nodes_in, int64_t n_nodes, void* edges_out, int64_t capacity,
                                   int64_t* n_out) {
  GF_REQUIRE(nodes_in && n_nodes > 0 && n_out, "gf_octree_edge_pool: null input / empty tree");
  struct EdgeB {
    int64_t t_idx_a, t_idx_b;
    float center[3], dir_0[3], dir_1[3];
    uint8_t pad[12];
  };
  static_assert(sizeof(EdgeB) == 64, "EdgePool blob layout");
  const TreeNodeBlob* nodes = (const TreeNodeBlob*)nodes_in;
  std::vector<int64_t> valid;
  for (int64_t i = 0; i < n_nodes; i++)
    if (nodes[i].trans_idx >= 0) valid.push_back(i);
  std::vector<EdgeB> pool;
  auto is_inside = [](const TreeNodeBlob& nd, const float* pt) {
    float m = 0.f;
    for (int k = 0; k < 3; k++) m = fmaxf(m, fabsf((pt[k] - nd.center[k]) / nd.side_len * 2.f));
    return m < (1.f + 1e-4f);
  };
  for (size_t ia = 0; ia < valid.size(); ia++)
    for (size_t ib = ia + 1; ib < valid.size(); ib++) {
      const int64_t a = valid[ia], b = valid[ib];
      int64_t u = a, v = b;
      if (nodes[u].side_len > nodes[v].side_len) std::swap(u, v);
      const float len = nodes[u].side_len * .5f;
      for (int axis = 0; axis < 3; axis++)
        for (int sgn = 0; sgn < 2; sgn++) {  // +x, -x, +y, -y, +z, -z, the reference's order
          float pt[3] = {nodes[u].center[0], nodes[u].center[1], nodes[u].center[2]};
          pt[axis] = sgn == 0 ? pt[axis] + len : pt[axis] - len;
          if (!is_inside(nodes[v], pt)) continue;
          EdgeB e;
          memset(&e, 0, sizeof(e));
          e.t_idx_a = nodes[a].trans_idx;
          e.t_idx_b = nodes[b].trans_idx;
          memcpy(e.center, pt, sizeof(pt));
          const int d0 = axis == 0 ? 1 : 0, d1 = axis == 2 ? 1 : 2;  // the two in-face axes, in x < y < z order
          e.dir_0[d0] = len;
          e.dir_1[d1] = len;
          pool.push_back(e);
        }
    }
  *n_out = (int64_t)pool.size();
  if (!edges_out) return GF_OK;
  GF_REQUIRE(capacity >= (int64_t)pool.size(), "gf_octree_edge_pool: output capacity %lld < %lld edges",
             (long long)capacity, (long long)pool.size());
  memcpy(edges_out, pool.data(), pool.size() * sizeof(EdgeB));
  return GF_OK;
}
