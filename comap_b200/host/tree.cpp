// Newick reader with Bio++'s node numbering: ids in post-order as read (a leaf when it is
// met, a parent after its last child, root last) -- PhylogeneticsApplicationTools::getTree,
// CoMap.cpp:125-126; the order of the rows of output.vectors.file.
#include "bpp.h"
#include <cstdlib>
#include <functional>

namespace host {

namespace {
struct RawNode {
  std::vector<int> kids;
  std::string name;
  double len = 0.;
  bool has_len = false;
};
} // namespace

Tree parse_newick(const std::string& text) {
  std::string s;
  for (char c : text) if (c != '\n' && c != '\r') s += c;
  size_t semi = s.find(';');
  if (semi != std::string::npos) s = s.substr(0, semi);
  s = trim(s);
  if (s.empty()) throw Error("Newick: empty tree");
  std::vector<RawNode> nodes;
  size_t pos = 0;
  // iterative-recursive descent (trees here are a few thousand nodes deep at most)
  std::function<int()> rec = [&]() -> int {
    RawNode n;
    while (pos < s.size() && isspace((unsigned char)s[pos])) pos++;
    if (pos < s.size() && s[pos] == '(') {
      pos++;
      for (;;) {
        n.kids.push_back(rec());
        while (pos < s.size() && isspace((unsigned char)s[pos])) pos++;
        if (pos >= s.size()) throw Error("Newick: unbalanced parentheses");
        if (s[pos] == ',') { pos++; continue; }
        if (s[pos] == ')') { pos++; break; }
        throw Error(std::string("Newick: unexpected character '") + s[pos] + "'");
      }
    }
    // label (leaf name, or bootstrap value of an inner node -- ignored there)
    std::string label;
    if (pos < s.size() && (s[pos] == '\'' || s[pos] == '"')) {
      char q = s[pos++];
      while (pos < s.size() && s[pos] != q) label += s[pos++];
      pos++;
    } else {
      while (pos < s.size() && s[pos] != ':' && s[pos] != ',' && s[pos] != ')' && s[pos] != '(' && s[pos] != '[') label += s[pos++];
    }
    while (pos < s.size() && s[pos] == '[') { // comments / NHX
      size_t e = s.find(']', pos);
      pos = e == std::string::npos ? s.size() : e + 1;
    }
    n.name = trim(label);
    if (pos < s.size() && s[pos] == ':') {
      pos++;
      char* end = nullptr;
      n.len = std::strtod(s.c_str() + pos, &end);
      if (end == s.c_str() + pos) throw Error("Newick: bad branch length");
      pos = (size_t)(end - s.c_str());
      n.has_len = true;
    }
    nodes.push_back(n);
    return (int)nodes.size() - 1;
  };
  int root = rec();
  Tree t;
  // a bifurcating root is removed as DRHomogeneousTreeLikelihood(checkRooted = true) does
  // (CoETools.cpp:124): the two root edges are merged; ids are then re-assigned in post-order
  if (nodes[root].kids.size() == 2) {
    int a = nodes[root].kids[0], b = nodes[root].kids[1];
    if (nodes[a].kids.empty() && !nodes[b].kids.empty()) std::swap(a, b);
    if (nodes[a].kids.empty()) throw Error("tree has only two leaves");
    nodes[b].len += nodes[a].len;
    nodes[a].kids.push_back(b);
    nodes[a].len = 0.;
    root = a;
    t.was_unrooted = true;
  }
  // post-order renumbering from `root`
  std::vector<int> order;
  std::vector<std::pair<int, size_t>> st{{root, 0}};
  while (!st.empty()) {
    auto& [v, k] = st.back();
    if (k < nodes[v].kids.size()) {
      int c = nodes[v].kids[k++];
      st.push_back({c, 0});
    } else {
      order.push_back(v);
      st.pop_back();
    }
  }
  std::vector<int> new_id(nodes.size(), -1);
  for (size_t i = 0; i < order.size(); i++) new_id[order[i]] = (int)i;
  const int n = (int)order.size();
  t.parent.assign(n, -1);
  t.brlen.assign(n, 0.);
  t.name.assign(n, "");
  for (int v : order) {
    int id = new_id[v];
    for (int c : nodes[v].kids) t.parent[new_id[c]] = id;
    t.brlen[id] = v == root ? 0. : nodes[v].len;
    if (nodes[v].kids.empty()) {
      if (nodes[v].name.empty()) throw Error("Newick: a leaf has no name");
      t.name[id] = nodes[v].name;
      t.leaves.push_back(id);
    }
  }
  t.n_root_children = (int)nodes[root].kids.size();
  if (t.leaves.size() < 3) throw Error("tree has fewer than three leaves");
  return t;
}

} // namespace host
