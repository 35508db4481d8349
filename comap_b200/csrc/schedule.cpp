// Tree handling and packed op streams for the mapping / simulation kernels.
//
// The reference walks TreeTemplate<Node> recursively per call (Bio++
// DRHomogeneousTreeLikelihood::computeSubtreeLikelihoodPostfix/Prefix,
// NonHomogeneousSequenceSimulator::evolveInternal; call sites CoETools.cpp:209,
// AnalysisTools.cpp:591-593).  Here the walk is compiled once per tree into flat, chunked
// op streams that each CTA streams through shared memory with TMA bulk copies; every
// record carries the transition / count tables it needs, in the order it needs them.
#include "common.h"
#include <algorithm>
#include <cstring>

namespace cmb {

void build_tree(Tree& t, int n_nodes, const int32_t* parent, const double* brlen) {
  if (n_nodes < 3) fail("cmb_set_tree: need at least 3 nodes");
  if (parent[n_nodes - 1] != -1) fail("cmb_set_tree: the last node must be the root (parent -1)");
  t = Tree();
  t.n_nodes = n_nodes;
  t.B = n_nodes - 1;
  t.parent.assign(parent, parent + n_nodes);
  t.brlen.assign(n_nodes, 0.);
  std::vector<std::vector<int>> kids(n_nodes);
  for (int v = 0; v < n_nodes - 1; v++) {
    if (parent[v] <= v || parent[v] >= n_nodes)
      fail("cmb_set_tree: ids must be in post-order (node %d has parent %d)", v, parent[v]);
    if (!(brlen[v] >= 0.) || !(brlen[v] <= 10000.)) fail("cmb_set_tree: bad branch length at node %d", v);
    kids[parent[v]].push_back(v);
    t.brlen[v] = brlen[v] < 1e-6 ? 1e-6 : brlen[v]; // Bio++ lower bound
  }
  t.leaf_row.assign(n_nodes, -1);
  for (int v = 0; v < n_nodes; v++) {
    if (kids[v].empty()) t.leaf_row[v] = t.n_leaves++;
    else if (kids[v].size() == 1) fail("cmb_set_tree: node %d has a single child", v);
  }
  if (kids[n_nodes - 1].empty()) fail("cmb_set_tree: root is a leaf");

  // binarise: a k-ary node becomes a chain of binary nodes joined by virtual zero-length
  // edges (P = I, no output row); products are taken in the same child order.
  t.bin.resize(n_nodes);
  for (int v = 0; v < n_nodes; v++) {
    t.bin[v].orig = v;
    t.bin[v].branch = (v == n_nodes - 1) ? -1 : v;
    t.bin[v].tip_row = t.leaf_row[v];
  }
  for (int v = 0; v < n_nodes; v++) {
    const auto& k = kids[v];
    if (k.empty()) continue;
    int cur = v;
    for (size_t i = 0; i + 2 < k.size(); i++) {
      BinNode u;
      int uid = (int)t.bin.size();
      t.bin.push_back(u);
      t.bin[cur].left = k[i];
      t.bin[cur].right = uid;
      t.bin[k[i]].parent = cur;
      t.bin[uid].parent = cur;
      cur = uid;
    }
    t.bin[cur].left = k[k.size() - 2];
    t.bin[cur].right = k[k.size() - 1];
    t.bin[k[k.size() - 2]].parent = cur;
    t.bin[k[k.size() - 1]].parent = cur;
  }
  t.bin_root = n_nodes - 1;
  // leaves below (iterative post-order)
  {
    std::vector<std::pair<int, int>> st{{t.bin_root, 0}};
    while (!st.empty()) {
      auto [v, phase] = st.back();
      st.pop_back();
      BinNode& n = t.bin[v];
      if (n.left < 0) { n.leaves = 1; continue; }
      if (phase == 0) {
        st.push_back({v, 1});
        st.push_back({n.left, 0});
        st.push_back({n.right, 0});
      } else n.leaves = t.bin[n.left].leaves + t.bin[n.right].leaves;
    }
  }
  // down order: post-order, larger child first (its message waits on the stack while the
  // smaller subtree is processed -> stack depth <= log2(#leaves))
  {
    int depth = 0;
    std::vector<std::pair<int, int>> st{{t.bin_root, 0}};
    while (!st.empty()) {
      auto [v, phase] = st.back();
      st.pop_back();
      const BinNode& n = t.bin[v];
      if (n.left < 0) continue;
      if (phase == 0) {
        int a = n.left, b = n.right;
        if (t.bin[b].leaves > t.bin[a].leaves) std::swap(a, b);
        st.push_back({v, 1});
        st.push_back({b, 0});
        st.push_back({a, 0});
      } else t.down_order.push_back(v);
    }
    // depth of the message stack
    int sp = 0;
    for (int v : t.down_order) {
      BinNode& n = t.bin[v];
      int a = n.left, b = n.right;
      if (t.bin[b].leaves > t.bin[a].leaves) std::swap(a, b);
      if (t.bin[a].left >= 0) sp--; // pop a's message
      if (v != t.bin_root) {
        const BinNode& p = t.bin[n.parent];
        int pa = p.left, pb = p.right;
        if (t.bin[pb].leaves > t.bin[pa].leaves) std::swap(pa, pb);
        if (pa == v) { sp++; depth = std::max(depth, sp); }
      }
    }
    t.down_depth = depth;
  }
  // up order: pre-order, smaller child first (the larger one's message waits)
  {
    int depth = 0;
    std::vector<int> st;
    int v = t.bin_root;
    while (v >= 0) {
      t.up_order.push_back(v);
      const BinNode& n = t.bin[v];
      int a = n.left, b = n.right;
      if (t.bin[a].leaves > t.bin[b].leaves) std::swap(a, b);
      bool ia = t.bin[a].left >= 0, ib = t.bin[b].left >= 0;
      if (ia) {
        if (ib) { st.push_back(b); depth = std::max(depth, (int)st.size()); }
        v = a;
      } else if (ib) v = b;
      else if (!st.empty()) { v = st.back(); st.pop_back(); }
      else v = -1;
    }
    t.up_depth = depth;
  }
  if (t.down_depth > kMaxStack || t.up_depth > kMaxStack)
    fail("cmb_set_tree: traversal stack depth %d exceeds %d", std::max(t.down_depth, t.up_depth), kMaxStack);
  assign_slots(t, false);
}

// Storage slots of the down partials.  With skip_cherries, a cherry (inner non-root node
// whose two children are leaves) gets none: the up pass recomputes its partial from the
// two tip rows, so it never travels through HBM (about a third of the inner nodes of a
// random tree).
void assign_slots(Tree& t, bool skip_cherries) {
  t.n_slots = 0;
  for (int v : t.down_order) {
    BinNode& n = t.bin[v];
    n.slot = -1;
    n.cherry = false;
    if (v == t.bin_root) continue;
    n.cherry = skip_cherries && t.bin[n.left].left < 0 && t.bin[n.right].left < 0;
    if (!n.cherry) n.slot = t.n_slots++;
  }
}

namespace {

struct Packer {
  OpStream& s;
  uint32_t cap;
  std::vector<unsigned char> cur;
  uint32_t nrec = 0;
  Packer(OpStream& s_, uint32_t cap_) : s(s_), cap(cap_) {}
  void flush() {
    if (cur.empty()) return;
    s.chunk_off.push_back((uint32_t)s.bytes.size());
    s.chunk_bytes.push_back((uint32_t)cur.size());
    s.chunk_nrec.push_back(nrec);
    s.chunk_cap = std::max<uint32_t>(s.chunk_cap, (uint32_t)cur.size());
    s.bytes.insert(s.bytes.end(), cur.begin(), cur.end());
    cur.clear();
    nrec = 0;
  }
  void add(const std::vector<unsigned char>& rec) {
    if (!cur.empty() && cur.size() + rec.size() > cap) flush();
    cur.insert(cur.end(), rec.begin(), rec.end());
    nrec++;
  }
};

void append(std::vector<unsigned char>& rec, const void* p, size_t n) {
  const unsigned char* b = static_cast<const unsigned char*>(p);
  rec.insert(rec.end(), b, b + n);
}
void pad16(std::vector<unsigned char>& rec) { rec.resize((rec.size() + 15) & ~size_t(15), 0); }

// table of `what` (0 = P, 1 = W, 2 = cumP) for bin node v's edge, classes [c0, c0+cb)
void append_table(std::vector<unsigned char>& rec, const Tree& t, const ModelTables& mt, int v,
                  int what, int c0, int cb) {
  const int A = mt.A;
  const size_t AA = (size_t)A * A;
  int br = t.bin[v].branch;
  if (br >= 0) {
    const std::vector<double>& src = what == 0 ? mt.P : what == 1 ? mt.W : mt.cumP;
    append(rec, &src[((size_t)br * mt.C + c0) * AA], sizeof(double) * AA * cb);
  } else { // virtual edge: P = I, W = 0, cumP = step
    std::vector<double> tab(AA * cb, 0.);
    for (int c = 0; c < cb; c++)
      for (int x = 0; x < A; x++)
        for (int y = 0; y < A; y++)
          tab[c * AA + x * A + y] = what == 0 ? (x == y) : what == 1 ? 0. : (y >= x ? 1. : 0.);
    append(rec, tab.data(), sizeof(double) * tab.size());
  }
}

uint32_t chunk_capacity(size_t max_rec) { return (uint32_t)std::max<size_t>(8192, max_rec); }

} // namespace

void build_down_stream(OpStream& s, const Tree& t, const ModelTables& mt, int c0, int cb) {
  s = OpStream();
  std::vector<std::vector<unsigned char>> recs;
  size_t max_rec = 0;
  for (int v : t.down_order) {
    const BinNode& n = t.bin[v];
    int a = n.left, b = n.right;
    if (t.bin[b].leaves > t.bin[a].leaves) std::swap(a, b);
    DownHdr h{};
    bool ta = t.bin[a].left < 0, tb = t.bin[b].left < 0;
    h.flags = (ta ? kDownTipA : 0) | (tb ? kDownTipB : 0);
    h.row_a = ta ? t.bin[a].tip_row : -1;
    h.row_b = tb ? t.bin[b].tip_row : -1;
    h.slot = n.slot;
    bool push = false;
    if (v == t.bin_root) h.flags |= kDownRoot;
    else {
      const BinNode& p = t.bin[n.parent];
      int pa = p.left, pb = p.right;
      if (t.bin[pb].leaves > t.bin[pa].leaves) std::swap(pa, pb);
      push = (pa == v);
    }
    if (push) h.flags |= kDownPush;
    std::vector<unsigned char> rec;
    append(rec, &h, sizeof h);
    append_table(rec, t, mt, b, 0, c0, cb);
    if (ta) append_table(rec, t, mt, a, 0, c0, cb);
    if (push) append_table(rec, t, mt, v, 0, c0, cb);
    pad16(rec);
    max_rec = std::max(max_rec, rec.size());
    recs.push_back(std::move(rec));
  }
  Packer pk(s, chunk_capacity(max_rec));
  for (auto& r : recs) pk.add(r);
  pk.flush();
}

static void up_like_stream(OpStream& s, const Tree& t, const ModelTables& mt, int c0, int cb, bool sim) {
  s = OpStream();
  std::vector<std::vector<unsigned char>> recs;
  size_t max_rec = 0;
  std::vector<int> st;
  for (size_t i = 0; i < t.up_order.size(); i++) {
    int v = t.up_order[i];
    const BinNode& n = t.bin[v];
    int a = n.left, b = n.right;
    if (t.bin[a].leaves > t.bin[b].leaves) std::swap(a, b);
    bool ta = t.bin[a].left < 0, tb = t.bin[b].left < 0;
    UpHdr h{};
    h.flags = (ta ? kUpTipA : 0) | (tb ? kUpTipB : 0);
    if (!ta) {
      h.flags |= kUpTakeA;
      if (!tb) { h.flags |= kUpPush; st.push_back(b); }
    } else if (!tb) h.flags |= kUpTakeB;
    else if (!st.empty()) { h.flags |= kUpPop; st.pop_back(); }
    const bool ca = !sim && t.bin[a].cherry, cb_ = !sim && t.bin[b].cherry;
    if (sim) {
      h.ref_a = ta ? t.bin[a].tip_row : -1;
      h.ref_b = tb ? t.bin[b].tip_row : -1;
      h.out_a = t.bin[a].orig;
      h.out_b = t.bin[b].orig;
      // RNG key: the original parent and the child's rank among its children (id order)
      auto rank_of = [&](int child) {
        if (t.bin[child].orig < 0) return 0;
        const int par = t.parent[t.bin[child].orig];
        int k = 0;
        for (int u = 0; u < t.bin[child].orig; u++) k += t.parent[u] == par;
        return k;
      };
      const int real = t.bin[a].orig >= 0 ? t.bin[a].orig : t.bin[b].orig;
      h.ref_a2 = rank_of(a);
      h.ref_b2 = rank_of(b);
      h.pad2 = real >= 0 ? t.parent[real] : -1;
    } else {
      h.ref_a = ta ? t.bin[a].tip_row : ca ? t.bin[t.bin[a].left].tip_row : t.bin[a].slot;
      h.ref_b = tb ? t.bin[b].tip_row : cb_ ? t.bin[t.bin[b].left].tip_row : t.bin[b].slot;
      h.ref_a2 = ca ? t.bin[t.bin[a].right].tip_row : -1;
      h.ref_b2 = cb_ ? t.bin[t.bin[b].right].tip_row : -1;
      if (ca) h.flags |= kUpCherryA;
      if (cb_) h.flags |= kUpCherryB;
      h.out_a = t.bin[a].branch;
      h.out_b = t.bin[b].branch;
    }
    s.aux.push_back((int32_t)h.flags); s.aux.push_back(h.ref_a); s.aux.push_back(h.ref_b); s.aux.push_back(0);
    s.aux.push_back(h.ref_a2); s.aux.push_back(h.ref_b2); s.aux.push_back(0); s.aux.push_back(0);
    s.n_records++;
    std::vector<unsigned char> rec;
    append(rec, &h, sizeof h);
    if (sim) {
      append_table(rec, t, mt, a, 2, 0, mt.C);
      append_table(rec, t, mt, b, 2, 0, mt.C);
    } else {
      append_table(rec, t, mt, a, 0, c0, cb);
      append_table(rec, t, mt, a, 1, c0, cb);
      append_table(rec, t, mt, b, 0, c0, cb);
      append_table(rec, t, mt, b, 1, c0, cb);
      if (ca) { append_table(rec, t, mt, t.bin[a].left, 0, c0, cb); append_table(rec, t, mt, t.bin[a].right, 0, c0, cb); }
      if (cb_) { append_table(rec, t, mt, t.bin[b].left, 0, c0, cb); append_table(rec, t, mt, t.bin[b].right, 0, c0, cb); }
    }
    pad16(rec);
    max_rec = std::max(max_rec, rec.size());
    recs.push_back(std::move(rec));
  }
  Packer pk(s, chunk_capacity(max_rec));
  for (auto& r : recs) pk.add(r);
  pk.flush();
}

// ---- tensor-core (DMMA m8n8k4) flavour of the up stream, A = 4 (k1_mma.cu)
// One chunk per node (chunk_off / chunk_bytes = the record's offset and size; the producer
// warp copies record n into ring stage n).  A record = UpHdr | F1a[C] | F1b[C] | F3a[C] (child
// a inner) | F3b[C] (child b inner) | raw P of a's two leaves [C] each (a cherry) | same for
// b.  F* are 32-double B-operand fragments, one double per lane (lane = 4 n + k holds B[k][n]):
//   F1: B[y][2x] = P[x][y], B[y][2x+1] = W[x][y]  -> lane (site s, q) gets (P D)[q], (W D)[q]
//   F3: B[x][2y] = P[x][y], B[x][2y+1] = 0        -> lane (site s, q) gets (P^T U)[q]
// The raw 4x4 tables are 128 B = one word per bank pair: any column pick is conflict-free.
static void table_of(const Tree& t, const ModelTables& mt, int v, int what, int c, double (&tab)[16]) {
  const int br = t.bin[v].branch;
  for (int x = 0; x < 4; x++)
    for (int y = 0; y < 4; y++) {
      if (br >= 0) tab[x * 4 + y] = (what == 0 ? mt.P : mt.W)[(((size_t)br * mt.C + c) * 4 + x) * 4 + y];
      else tab[x * 4 + y] = what == 0 ? (x == y ? 1. : 0.) : 0.; // virtual edge: P = I, W = 0
    }
}
void build_up_mma_stream(OpStream& s, const Tree& t, const ModelTables& mt) {
  if (mt.A != 4) fail("internal: the tensor-core up stream is built for A = 4");
  s = OpStream();
  const int C = mt.C;
  std::vector<std::vector<unsigned char>> recs;
  size_t max_rec = 0;
  // Cherries are not nodes of this walk: a cherry child is expanded INSIDE its parent's record (its partial
  // is recomputed from the two tip rows, its message stays in registers and its two leaf branches are
  // contracted there), so a third of the node iterations and the stack traffic of their messages go away.
  // Kinds of a child: inner (stored partial), tip, cherry.  With the smaller child first only
  // tip-tip, tip-cherry, tip-inner, cherry-cherry, cherry-inner and inner-inner occur.
  auto kind_of = [&](int x) { return t.bin[x].left < 0 ? 1 : t.bin[x].cherry ? 2 : 0; };
  std::vector<int> st;
  int depth = 0;
  int v = t.bin_root;
  while (v >= 0) {
    const BinNode& n = t.bin[v];
    int a = n.left, b = n.right;
    if (t.bin[a].leaves > t.bin[b].leaves) std::swap(a, b);
    const int ka = kind_of(a), kb = kind_of(b);
    if (ka == 0 && kb != 0) fail("internal: up order must expand the smaller child first");
    const bool ta = ka == 1, tb = kb == 1, ca = ka == 2, cb = kb == 2;
    UpMmaHdr h{};
    h.flags = (ta ? kUpTipA : 0) | (tb ? kUpTipB : 0) | (ca ? kUpCherryA : 0) | (cb ? kUpCherryB : 0);
    int next = -1;
    int push_level = 0, pop_level = 0xff; // stack slots the consumers use: no run-time stack pointer
    if (ka == 0) { h.flags |= kUpTakeA | kUpPush; push_level = (int)st.size(); st.push_back(b); depth = std::max(depth, (int)st.size()); next = a; }
    else if (kb == 0) { h.flags |= kUpTakeB; next = b; }
    else if (!st.empty()) { h.flags |= kUpPop; next = st.back(); st.pop_back(); pop_level = (int)st.size(); }
    h.kase = (uint32_t)(ka * 3 + kb) | (uint32_t)push_level << 8 | (uint32_t)pop_level << 16;
    h.ref_a = ta ? t.bin[a].tip_row : ca ? t.bin[t.bin[a].left].tip_row : t.bin[a].slot;
    h.ref_b = tb ? t.bin[b].tip_row : cb ? t.bin[t.bin[b].left].tip_row : t.bin[b].slot;
    const int ref_a2 = ca ? t.bin[t.bin[a].right].tip_row : -1;
    const int ref_b2 = cb ? t.bin[t.bin[b].right].tip_row : -1;
    h.out_a = t.bin[a].branch;
    h.out_b = t.bin[b].branch;
    h.out_a1 = ca ? t.bin[t.bin[a].left].branch : -1;
    h.out_a2 = ca ? t.bin[t.bin[a].right].branch : -1;
    h.out_b1 = cb ? t.bin[t.bin[b].left].branch : -1;
    h.out_b2 = cb ? t.bin[t.bin[b].right].branch : -1;
    s.n_records++;
    std::vector<unsigned char> rec;
    append(rec, &h, sizeof h);
    double P[16], W[16], frag[32];
    for (int e = 0; e < 2; e++) { // F1a, F1b: children whose partial goes through the DMMA
      if (e ? tb : ta) continue;
      for (int c = 0; c < C; c++) {
        table_of(t, mt, e ? b : a, 0, c, P);
        table_of(t, mt, e ? b : a, 1, c, W);
        for (int l = 0; l < 32; l++) {
          const int k = l & 3, nn = l >> 2;
          frag[l] = (nn & 1) ? W[(nn >> 1) * 4 + k] : P[(nn >> 1) * 4 + k];
        }
        append(rec, frag, sizeof frag);
      }
    }
    for (int e = 0; e < 2; e++) { // F3a, F3b: messages to non-tip children
      if (e ? tb : ta) continue;
      for (int c = 0; c < C; c++) {
        table_of(t, mt, e ? b : a, 0, c, P);
        for (int l = 0; l < 32; l++) {
          const int k = l & 3, nn = l >> 2;
          frag[l] = (nn & 1) ? 0. : P[k * 4 + (nn >> 1)];
        }
        append(rec, frag, sizeof frag);
      }
    }
    for (int e = 0; e < 2; e++) { // raw 4x4 tables for conflict-free column picks
      const int x = e ? b : a;
      if (e ? tb : ta) {          // tip: P[C], W[C]
        for (int what = 0; what < 2; what++)
          for (int c = 0; c < C; c++) { table_of(t, mt, x, what, c, P); append(rec, P, sizeof P); }
      } else if (e ? cb : ca) {   // cherry: P1[C], P2[C], W1[C], W2[C] of its two leaf edges
        for (int what = 0; what < 2; what++)
          for (int leaf : {t.bin[x].left, t.bin[x].right})
            for (int c = 0; c < C; c++) { table_of(t, mt, leaf, what, c, P); append(rec, P, sizeof P); }
      }
    }
    pad16(rec);
    // Packed stage of this node in the kernel's ring: record | four tip rows (when a child is a
    // tip or a cherry) | partial chunk of each stored child.
    const uint32_t tips_off = (uint32_t)((rec.size() + 127) & ~size_t(127));
    const uint32_t blk_off = tips_off + ((ta || tb || ca || cb) ? 4u * kChunkSites : 0u);
    const uint32_t n_blk = (uint32_t)(ka == 0) + (uint32_t)(kb == 0);
    s.stage_bytes = std::max(s.stage_bytes, blk_off + n_blk * (uint32_t)C * kChunkSites * 32u);
    s.blk_off_max[n_blk] = std::max(s.blk_off_max[n_blk], blk_off);
    h.tips_off = (int32_t)tips_off;
    h.blk_off = (int32_t)blk_off;
    std::memcpy(rec.data(), &h, sizeof h);
    s.aux.push_back((int32_t)h.flags); s.aux.push_back(h.ref_a); s.aux.push_back(h.ref_b); s.aux.push_back((int32_t)tips_off);
    s.aux.push_back(ref_a2); s.aux.push_back(ref_b2); s.aux.push_back((int32_t)blk_off); s.aux.push_back(0);
    max_rec = std::max(max_rec, rec.size());
    recs.push_back(std::move(rec));
    v = next;
  }
  s.stack_depth = depth;
  Packer pk(s, (uint32_t)max_rec);
  for (auto& r : recs) { pk.add(r); pk.flush(); }
}

// Down stream for the tensor-core kernels (A = 4), one chunk per node, post-order with the
// larger child (a) first.  Where the operands of node v come from:
//   a tip, b tip    : two column picks                          (raw P_a[C], P_b[C])
//   a inner, b tip  : a's partial is the running one -> DMMA    (F_a[C], raw P_b[C])
//   a inner, b inner: a's message waits on the stack, b's partial is the running one (F_b[C])
// kDownPush: v is the larger child of a node whose other child is inner too, so v's message
// P_v D_v goes to the stack (F_v[C] follows).  F: B[y][2x] = P[x][y], B[y][2x+1] = 0.
void build_down_mma_stream(OpStream& s, const Tree& t, const ModelTables& mt) {
  if (mt.A != 4) fail("internal: the tensor-core down stream is built for A = 4");
  s = OpStream();
  const int C = mt.C;
  std::vector<std::vector<unsigned char>> recs;
  size_t max_rec = 0;
  auto larger_first = [&](int v, int& a, int& b) {
    a = t.bin[v].left; b = t.bin[v].right;
    if (t.bin[b].leaves > t.bin[a].leaves) std::swap(a, b);
  };
  int depth = 0, sp = 0;
  for (int v : t.down_order) {
    const BinNode& n = t.bin[v];
    int a, b;
    larger_first(v, a, b);
    const bool ta = t.bin[a].left < 0, tb = t.bin[b].left < 0;
    if (ta && !tb) fail("internal: down order must expand the larger child first");
    DownHdr h{};
    h.flags = (ta ? kDownTipA : 0) | (tb ? kDownTipB : 0);
    h.row_a = ta ? t.bin[a].tip_row : -1;
    h.row_b = tb ? t.bin[b].tip_row : -1;
    h.slot = n.slot;
    int pop_level = 0xff, push_level = 0xff; // stack slots in the record: no run-time stack pointer
    if (!ta && !tb) pop_level = --sp; // pops a's message
    bool push = false;
    if (v == t.bin_root) h.flags |= kDownRoot;
    else {
      int pa, pb;
      larger_first(n.parent, pa, pb);
      push = pa == v && t.bin[pb].left >= 0;
    }
    if (push) { h.flags |= kDownPush; push_level = sp; depth = std::max(depth, ++sp); }
    h.flags |= (uint32_t)pop_level << 16 | (uint32_t)push_level << 24;
    s.aux.push_back((int32_t)h.flags); s.aux.push_back(h.row_a); s.aux.push_back(h.row_b); s.aux.push_back(0);
    s.n_records++;
    std::vector<unsigned char> rec;
    append(rec, &h, sizeof h);
    double P[16], frag[32];
    auto add_frag = [&](int node) {
      for (int c = 0; c < C; c++) {
        table_of(t, mt, node, 0, c, P);
        for (int l = 0; l < 32; l++) {
          const int k = l & 3, nn = l >> 2;
          frag[l] = (nn & 1) ? 0. : P[(nn >> 1) * 4 + k];
        }
        append(rec, frag, sizeof frag);
      }
    };
    auto add_raw = [&](int node) {
      for (int c = 0; c < C; c++) {
        table_of(t, mt, node, 0, c, P);
        append(rec, P, sizeof P);
      }
    };
    if (!ta) add_frag(tb ? a : b);
    if (ta) add_raw(a);
    if (tb) add_raw(b);
    if (push) add_frag(v);
    pad16(rec);
    max_rec = std::max(max_rec, rec.size());
    recs.push_back(std::move(rec));
  }
  s.stack_depth = depth;
  Packer pk(s, (uint32_t)max_rec);
  for (auto& r : recs) { pk.add(r); pk.flush(); }
}

// ---- tensor-core (DMMA m8n8k4) streams for proteins, A = 20 (k1_mma20.cu) ------------------------------------
// A 20-vector of a site lives in five registers of the quad of lanes that owns the site: lane q holds states
// 4 kt + q, kt = 0..4 (the A-operand layout of a k-tile).  A 20 x 20 matrix-vector product y = M x is 3 n-tiles x
// 5 k-tiles of DMMA m8n8k4; the B operand of (nt, kt) is one double per lane, lane = 4 n + k -> M[xo(n)][4 kt + k]
// with the output columns permuted, xo(n) = 4 (2 nt + (n & 1)) + (n >> 1), so that the C fragment of a lane (its two
// columns 2 q and 2 q + 1) is states 4 (2 nt) + q and 4 (2 nt + 1) + q: k-tiles 2 nt and 2 nt + 1 of y in the very
// layout the next product reads -- chained products need no shuffle.  15 fragments of 256 B per matrix.
// One record per (node, class): a CTA works on ONE rate class, so a record is the class's tables only.
namespace {
void table_of_a(const Tree& t, const ModelTables& mt, int v, int what, int c, std::vector<double>& tab) {
  const int A = mt.A, br = t.bin[v].branch;
  tab.resize((size_t)A * A);
  for (int x = 0; x < A; x++)
    for (int y = 0; y < A; y++) {
      if (br >= 0) tab[x * A + y] = (what == 0 ? mt.P : mt.W)[(((size_t)br * mt.C + c) * A + x) * A + y];
      else tab[x * A + y] = what == 0 ? (x == y ? 1. : 0.) : 0.; // virtual edge: P = I, W = 0
    }
}
// fragments of y = M x (transposed = false) or y = M^T x (true), M row-major 20 x 20
void append_frags20(std::vector<unsigned char>& rec, const std::vector<double>& M, bool transposed) {
  double frag[32];
  for (int nt = 0; nt < 3; nt++)
    for (int kt = 0; kt < 5; kt++) {
      for (int l = 0; l < 32; l++) {
        const int k = l & 3, n = l >> 2;
        const int xo = 4 * (2 * nt + (n & 1)) + (n >> 1), xi = 4 * kt + k;
        frag[l] = xo < 20 ? (transposed ? M[xi * 20 + xo] : M[xo * 20 + xi]) : 0.;
      }
      append(rec, frag, sizeof frag);
    }
}
// raw table for column picks of a resolved tip, transposed: [state][xo] = M[xo][state]
void append_raw20_t(std::vector<unsigned char>& rec, const std::vector<double>& M) {
  double tab[400];
  for (int st = 0; st < 20; st++)
    for (int xo = 0; xo < 20; xo++) tab[st * 20 + xo] = M[xo * 20 + st];
  append(rec, tab, sizeof tab);
}
} // namespace

// Up stream: pre-order, smaller child first; kinds inner (0) / tip (1) only (cherries keep their stored partial).
// Record (node, class) = UpMmaHdr | child a: inner FP[15] FW[15] FM[15] (S = P D, T = W D, message = P^T U),
// tip PT[400] WT[400] | child b likewise.  Records are laid out node-major: index node * C + class.
void build_up_mma20_stream(OpStream& s, const Tree& t, const ModelTables& mt, int sites_per_cta) {
  if (mt.A != 20) fail("internal: the protein tensor-core up stream is built for A = 20");
  s = OpStream();
  const int C = mt.C;
  std::vector<int> st;
  int depth = 0;
  std::vector<double> P, W;
  Packer pk(s, 1u << 30);
  size_t max_rec = 0;
  for (size_t idx = 0; idx < t.up_order.size(); idx++) {
    const int v = t.up_order[idx];
    const BinNode& n = t.bin[v];
    int a = n.left, b = n.right;
    if (t.bin[a].leaves > t.bin[b].leaves) std::swap(a, b);
    const bool ta = t.bin[a].left < 0, tb = t.bin[b].left < 0;
    if (!ta && tb) fail("internal: up order must expand the smaller child first");
    UpMmaHdr h{};
    h.flags = (ta ? kUpTipA : 0) | (tb ? kUpTipB : 0);
    int push_level = 0, pop_level = 0xff;
    if (!ta) { h.flags |= kUpTakeA | kUpPush; push_level = (int)st.size(); st.push_back(b); depth = std::max(depth, (int)st.size()); }
    else if (!tb) h.flags |= kUpTakeB;
    else if (!st.empty()) { h.flags |= kUpPop; st.pop_back(); pop_level = (int)st.size(); }
    h.kase = (uint32_t)((ta ? 1 : 0) * 3 + (tb ? 1 : 0)) | (uint32_t)push_level << 8 | (uint32_t)pop_level << 16;
    h.ref_a = ta ? t.bin[a].tip_row : t.bin[a].slot;
    h.ref_b = tb ? t.bin[b].tip_row : t.bin[b].slot;
    h.out_a = t.bin[a].branch; h.out_b = t.bin[b].branch;
    h.out_a1 = h.out_a2 = h.out_b1 = h.out_b2 = -1;
    const uint32_t rec_bytes = (uint32_t)(sizeof h + ((ta ? 2 * 3200 : 3 * 3840) + (tb ? 2 * 3200 : 3 * 3840)));
    const uint32_t tips_off = (rec_bytes + 127) & ~127u;
    const uint32_t blk_off = tips_off + ((ta || tb) ? 2u * (uint32_t)sites_per_cta : 0u);
    const uint32_t n_blk = (uint32_t)!ta + (uint32_t)!tb;
    s.stage_bytes = std::max(s.stage_bytes, blk_off + n_blk * (uint32_t)sites_per_cta * 160u);
    h.tips_off = (int32_t)tips_off; h.blk_off = (int32_t)blk_off;
    s.aux.push_back((int32_t)h.flags); s.aux.push_back(h.ref_a); s.aux.push_back(h.ref_b); s.aux.push_back((int32_t)tips_off);
    s.aux.push_back(-1); s.aux.push_back(-1); s.aux.push_back((int32_t)blk_off); s.aux.push_back(0);
    s.n_records++;
    for (int c = 0; c < C; c++) {
      std::vector<unsigned char> rec;
      append(rec, &h, sizeof h);
      for (int e = 0; e < 2; e++) {
        const int x = e ? b : a;
        table_of_a(t, mt, x, 0, c, P);
        table_of_a(t, mt, x, 1, c, W);
        if (e ? tb : ta) { append_raw20_t(rec, P); append_raw20_t(rec, W); }
        else { append_frags20(rec, P, false); append_frags20(rec, W, false); append_frags20(rec, P, true); }
      }
      if (rec.size() != rec_bytes) fail("internal: protein up record size");
      max_rec = std::max(max_rec, rec.size());
      pk.add(rec); pk.flush();
    }
  }
  s.stack_depth = depth;
  s.chunk_cap = (uint32_t)max_rec;
}

// Down stream: post-order, larger child (a) first; record (node, class) = DownHdr | tables:
//   a tip, b tip    : PT_a[400] PT_b[400]
//   a inner, b tip  : FP_a[15] (a's partial is the running one), PT_b[400]
//   a inner, b inner: FP_b[15] (b's partial is the running one; a's message waits on the stack)
//   + FP_v[15] when v's own message goes to the stack (kDownPush)
void build_down_mma20_stream(OpStream& s, const Tree& t, const ModelTables& mt) {
  if (mt.A != 20) fail("internal: the protein tensor-core down stream is built for A = 20");
  s = OpStream();
  const int C = mt.C;
  auto larger_first = [&](int v, int& a, int& b) {
    a = t.bin[v].left; b = t.bin[v].right;
    if (t.bin[b].leaves > t.bin[a].leaves) std::swap(a, b);
  };
  int depth = 0, sp = 0;
  std::vector<double> P;
  Packer pk(s, 1u << 30);
  size_t max_rec = 0;
  for (int v : t.down_order) {
    const BinNode& n = t.bin[v];
    int a, b;
    larger_first(v, a, b);
    const bool ta = t.bin[a].left < 0, tb = t.bin[b].left < 0;
    if (ta && !tb) fail("internal: down order must expand the larger child first");
    DownHdr h{};
    h.flags = (ta ? kDownTipA : 0) | (tb ? kDownTipB : 0);
    h.row_a = ta ? t.bin[a].tip_row : -1;
    h.row_b = tb ? t.bin[b].tip_row : -1;
    h.slot = n.slot;
    int pop_level = 0xff, push_level = 0xff;
    if (!ta && !tb) pop_level = --sp; // pops a's message
    bool push = false;
    if (v == t.bin_root) h.flags |= kDownRoot;
    else {
      int pa, pb;
      larger_first(n.parent, pa, pb);
      push = pa == v && t.bin[pb].left >= 0;
    }
    if (push) { h.flags |= kDownPush; push_level = sp; depth = std::max(depth, ++sp); }
    h.flags |= (uint32_t)pop_level << 16 | (uint32_t)push_level << 24; // stack slots: no run-time stack pointer
    s.aux.push_back((int32_t)h.flags); s.aux.push_back(h.row_a); s.aux.push_back(h.row_b); s.aux.push_back(0);
    s.n_records++;
    for (int c = 0; c < C; c++) {
      std::vector<unsigned char> rec;
      append(rec, &h, sizeof h);
      if (!ta) { table_of_a(t, mt, tb ? a : b, 0, c, P); append_frags20(rec, P, false); }
      if (ta) { table_of_a(t, mt, a, 0, c, P); append_raw20_t(rec, P); }
      if (tb) { table_of_a(t, mt, b, 0, c, P); append_raw20_t(rec, P); }
      if (push) { table_of_a(t, mt, v, 0, c, P); append_frags20(rec, P, false); }
      pad16(rec);
      max_rec = std::max(max_rec, rec.size());
      pk.add(rec); pk.flush();
    }
  }
  s.stack_depth = depth;
  s.chunk_cap = (uint32_t)max_rec;
}

void build_up_stream(OpStream& s, const Tree& t, const ModelTables& mt, int c0, int cb) {
  up_like_stream(s, t, mt, c0, cb, false);
}
void build_sim_stream(OpStream& s, const Tree& t, const ModelTables& mt) {
  up_like_stream(s, t, mt, 0, mt.C, true);
}

} // namespace cmb
