// Host-side tree model: Newick parsing and flat descriptors.
//
// Mirrors what the reference builds in fitch.cpp:97-193 (PhylogenyParser ->
// Node child lists -> Tree::postOrder) and newick.py:354-414 (parseNewick), but
// as index arrays instead of pointer nodes, because its only consumers are the
// program builders that emit flat op lists for the device.
#pragma once
#include <string>
#include <vector>

#include "common.h"

namespace plg {

struct HostNode {
  int parent = -1;
  std::vector<int> children;  // in Newick string order (fitch.cpp:40 push_back)
  std::string label;
  double brlen = 0.0;  // length of the branch above this node (0 if absent)
  bool has_len = false;
};

struct HostTree {
  std::vector<HostNode> nodes;
  int root = -1;
  int n_leaves() const {
    int c = 0;
    for (auto &n : nodes) c += n.children.empty();
    return c;
  }
  // children-first order, children in list order, then self (fitch.cpp:85-91)
  std::vector<int> post_order() const;
};

// Parses one Newick string.  Inside the reference's domain fence (SURVEY.md 8a:
// strings on which fitch.cpp:148-191 terminates) the child order equals the
// reference's.  Throws Error{PLG_ERR_PARSE}.
HostTree parse_newick(const char *s);

// Flat batch of k-ary trees in CSR form, the common input of the program builders.
// Internal node i of tree t is node_off[t] + i (post-order within the tree; the
// last one is the root).  child[] entries: >= 0 -> alignment row (tip),
// < 0 -> internal node -(x+1) of the same tree.
struct TreeBatch {
  std::vector<int64_t> node_off;   // [n_trees+1] into child_off
  std::vector<int64_t> child_off;  // [n_nodes+1] into child
  std::vector<int32_t> child;
  std::vector<double> child_len;   // same shape as child (likelihood only; may be empty)
  int64_t n_trees() const { return (int64_t)node_off.size() - 1; }
};

// label -> alignment row lookup used when lowering Newick trees (fitch.cpp:312-323)
struct TaxaMap {
  std::vector<std::pair<std::string, int>> sorted;  // by label
  TaxaMap(int n, const char *const *names, const int32_t *idx);
  int find(const std::string &label) const;  // -1 if absent
};

// Appends `t` to `b`.  A tree that is a single leaf gets one unary internal node
// so that every tree has a root op.  Throws Error{PLG_ERR_LABEL}.
void append_tree(TreeBatch &b, const HostTree &t, const TaxaMap &taxa, bool with_len);

// Binary descriptor [n_trees][n_tips-1][2] (+ optional brlen) -> TreeBatch.
TreeBatch batch_from_desc(int64_t n_trees, int n_tips, const int32_t *desc, const double *brlen,
                          const int32_t *tip_rows);

}  // namespace plg
