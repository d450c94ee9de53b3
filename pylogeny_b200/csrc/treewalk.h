// Depth-first "tree walk" programs for batches of binary descriptor trees (plg_*_score_trees).
//
// The level-synchronous programs (fitch.cu / lik.cu) store every internal node's vector in HBM and
// read it back one launch later: 3 vector transfers per node-op, the streaming roofline of
// SURVEY.md 8d.  A tree walk keeps a whole post-order on chip instead: one thread (Fitch) or one
// thread per pattern (likelihood) evaluates ALL n-1 node-ops of a tree for its patterns, the
// running result stays in registers, and only where both children of a node are internal is one
// of them parked on a small stack (<= ceil(log2 n) levels when the child that needs more levels is
// evaluated first -- Sethi-Ullman / Strahler order).  HBM then only supplies the tips.
//
// The stack program of every tree is produced ON THE DEVICE from the raw descriptor
// (plan_walk_kernel, one thread per tree), so the host never touches per-node data: it copies
// desc / brlen up and the scores down.
#pragma once
#include "common.h"

namespace plg {

// One node-op of a walk, in execution order (the root is last).
//   a, b  >= 0: tip row;  -1: the value the previous op left in registers ("top");
//         <= -2: pop stack level (-2 - x)
//   push  >= 0: park the result at that stack level;  -1: it stays in registers for the next op
//   slot  2*i, i = the descriptor entry of this node: a is its first child (branch data at child
//         slot 2*i), b its second (2*i + 1)
struct WalkOp {
  int a, b, push, slot;
};

constexpr int WALK_MAX_LEVELS = 16;  // trees of up to 65,536 tips

// stack levels that suffice for any binary tree of n_tips leaves in Strahler order
inline int walk_levels(int n_tips) {
  int l = 1;
  while ((1 << l) < n_tips) l++;
  return l;
}

// Host check of a descriptor batch (the errors batch_from_desc reports) without building a batch.
void validate_desc(int64_t n_trees, int n_tips, const int32_t *desc, const int32_t *tip_rows, int n_rows,
                   bool rows_matter);

// Device planner: desc[n_trees][n_tips-1][2] (device) -> ops[n_trees][n_tips-1] (device).
// tip_rows (device, may be null) maps tip ids to alignment rows; rows are clamped to
// [0, n_rows-1].  scratch must hold n_trees * (3 * (n_tips-1) + 2) ints.
void launch_plan_walk(int64_t n_trees, int n_tips, const int32_t *d_desc, const int32_t *d_tip_rows, int n_rows,
                      WalkOp *d_ops, int *d_scratch, cudaStream_t st);

}  // namespace plg
