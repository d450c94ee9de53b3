#include "tree.h"

#include <algorithm>
#include <cstdlib>
#include <cstring>

namespace plg {

namespace {

struct Parser {
  const char *s;
  size_t len;
  HostTree t;

  static bool stop(char c) { return c == ',' || c == ':' || c == ';' || c == ')'; }

  // label: every char up to one of , : ; )   (fitch.cpp:137-146, newick.py getLeafName)
  void read_label(size_t &i, int node) {
    size_t a = i;
    while (i < len && !stop(s[i])) ++i;
    t.nodes[node].label.assign(s + a, i - a);
  }

  // ":<number>" then anything up to the next , or ) is ignored (fitch.cpp:186)
  void read_length(size_t &i, int node) {
    if (i < len && s[i] == ':') {
      ++i;
      char *end = nullptr;
      double v = strtod(s + i, &end);
      if (end != s + i) {
        t.nodes[node].brlen = v;
        t.nodes[node].has_len = true;
        i = (size_t)(end - s);
      }
      while (i < len && s[i] != ',' && s[i] != ')' && s[i] != ';') ++i;
    }
  }

  int subtree(size_t &i, int parent, int depth) {
    if (depth > 100000) PLG_FAIL(PLG_ERR_PARSE, "Newick nesting too deep");
    int id = (int)t.nodes.size();
    t.nodes.emplace_back();
    t.nodes[id].parent = parent;
    if (parent >= 0) t.nodes[parent].children.push_back(id);
    if (i < len && s[i] == '(') {
      ++i;
      for (;;) {
        subtree(i, id, depth + 1);
        if (i >= len) PLG_FAIL(PLG_ERR_PARSE, "unbalanced '(' in Newick string");
        if (s[i] == ',') {
          ++i;
          continue;
        }
        if (s[i] == ')') {
          ++i;
          break;
        }
        PLG_FAIL(PLG_ERR_PARSE, "unexpected '%c' at offset %zu in Newick string", s[i], i);
      }
      read_label(i, id);  // internal label / support value (fitch.cpp:168-173)
      read_length(i, id);
    } else {
      read_label(i, id);
      read_length(i, id);
    }
    return id;
  }
};

}  // namespace

HostTree parse_newick(const char *s) {
  if (!s) PLG_FAIL(PLG_ERR_PARSE, "null Newick string");
  Parser p{s, strlen(s), {}};
  size_t i = 0;
  while (i < p.len && (s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
  if (i >= p.len) PLG_FAIL(PLG_ERR_PARSE, "empty Newick string");
  p.t.root = p.subtree(i, -1, 0);
  while (i < p.len && (s[i] == ';' || s[i] == ' ' || s[i] == '\n' || s[i] == '\t' || s[i] == '\r')) ++i;
  if (i < p.len) PLG_FAIL(PLG_ERR_PARSE, "trailing characters at offset %zu in Newick string", i);
  return std::move(p.t);
}

std::vector<int> HostTree::post_order() const {
  std::vector<int> out;
  out.reserve(nodes.size());
  // iterative: (node, next child index)
  std::vector<std::pair<int, size_t>> st;
  st.emplace_back(root, 0);
  while (!st.empty()) {
    auto &top = st.back();
    const HostNode &n = nodes[top.first];
    if (top.second < n.children.size()) {
      int c = n.children[top.second++];
      st.emplace_back(c, 0);
    } else {
      out.push_back(top.first);
      st.pop_back();
    }
  }
  return out;
}

TaxaMap::TaxaMap(int n, const char *const *names, const int32_t *idx) {
  sorted.reserve(n);
  for (int i = 0; i < n; i++) sorted.emplace_back(std::string(names[i]), (int)idx[i]);
  // std::map::insert keeps the first of duplicate keys (fitch.cpp:322)
  std::stable_sort(sorted.begin(), sorted.end(),
                   [](const std::pair<std::string, int> &a, const std::pair<std::string, int> &b) {
                     return a.first < b.first;
                   });
}

int TaxaMap::find(const std::string &label) const {
  auto it = std::lower_bound(sorted.begin(), sorted.end(), label,
                             [](const std::pair<std::string, int> &a, const std::string &b) {
                               return a.first < b;
                             });
  if (it == sorted.end() || it->first != label) return -1;
  return it->second;
}

void append_tree(TreeBatch &b, const HostTree &t, const TaxaMap &taxa, bool with_len) {
  if (b.node_off.empty()) {
    b.node_off.push_back(0);
    b.child_off.push_back(0);
  }
  std::vector<int> po = t.post_order();
  std::vector<int> local(t.nodes.size(), -1);
  int n_inner = 0;
  auto emit_child = [&](int c) {
    const HostNode &cn = t.nodes[c];
    if (cn.children.empty()) {
      int row = taxa.find(cn.label);
      if (row < 0) PLG_FAIL(PLG_ERR_LABEL, "leaf label '%s' is not in the taxa map", cn.label.c_str());
      b.child.push_back(row);
    } else {
      b.child.push_back(-(local[c] + 1));
    }
    if (with_len) b.child_len.push_back(cn.brlen);
  };
  for (int x : po) {
    const HostNode &n = t.nodes[x];
    if (n.children.empty()) continue;
    for (int c : n.children) emit_child(c);
    local[x] = n_inner++;
    b.child_off.push_back((int64_t)b.child.size());
  }
  if (n_inner == 0) {  // single-leaf tree: unary root so the tree owns one op
    emit_child(t.root);
    b.child_off.push_back((int64_t)b.child.size());
    n_inner = 1;
  }
  b.node_off.push_back(b.node_off.back() + n_inner);
}

TreeBatch batch_from_desc(int64_t n_trees, int n_tips, const int32_t *desc, const double *brlen,
                          const int32_t *tip_rows) {
  if (n_tips < 2) PLG_FAIL(PLG_ERR_ARG, "descriptor trees need at least 2 tips");
  TreeBatch b;
  const int ni = n_tips - 1;
  b.node_off.resize(n_trees + 1);
  b.child_off.resize(n_trees * ni + 1);
  b.child.resize((size_t)n_trees * ni * 2);
  if (brlen) b.child_len.assign(brlen, brlen + (size_t)n_trees * ni * 2);
  for (int64_t t = 0; t <= n_trees; t++) b.node_off[t] = t * ni;
  for (int64_t x = 0; x <= n_trees * ni; x++) b.child_off[x] = 2 * x;
  if (tip_rows)
    for (int i = 0; i < n_tips; i++)
      if (tip_rows[i] < 0) PLG_FAIL(PLG_ERR_ARG, "tip row %d is negative", tip_rows[i]);
  std::vector<int64_t> used_by((size_t)n_tips + ni, -1);  // every node is a child at most (hence exactly) once
  for (int64_t t = 0; t < n_trees; t++)
    for (int i = 0; i < ni; i++)
      for (int c = 0; c < 2; c++) {
        int32_t id = desc[((size_t)t * ni + i) * 2 + c];
        if (id < 0 || id >= n_tips + i)
          PLG_FAIL(PLG_ERR_ARG, "descriptor of tree %lld node %d: child id %d is not earlier in post-order",
                   (long long)t, i, id);
        if (used_by[id] == t)
          PLG_FAIL(PLG_ERR_ARG, "descriptor of tree %lld node %d: node %d is a child twice (not a tree)",
                   (long long)t, i, id);
        used_by[id] = t;
        b.child[((size_t)t * ni + i) * 2 + c] = id < n_tips ? (tip_rows ? tip_rows[id] : id) : -(id - n_tips + 1);
      }
  return b;
}

}  // namespace plg
