// Fitch parsimony engine: device-resident alignment, op programs, executor.
#pragma once
#include "common.h"
#include "tree.h"

namespace plg {

constexpr int FITCH_MAX_PLANES = 20;
constexpr int FITCH_JOIN3 = -3;  // FitchOp::nchild marker: cost += miss(fitch(c0, c1), c2), nothing stored
constexpr int FITCH_FUSED = -4;  // N = fitch(c0, c1) stored to dst; cost += miss(fitch(N, c2), c3)

// One internal node = one op (16 bytes, read once per block).
struct FitchOp {
  int dst;        // slot to store the node's state set, -1 = do not store (roots, pure cost ops)
  int nchild;     // children folded in list order (fitch.cpp:215-226)
  int child_off;  // into the children[] slot array
  int tree;       // index of the cost accumulator this op adds to
};

// Depth-first SPR search walks for parsimony (fitch_walk.cuh fitch_search_walk_kernel): one group
// = one pruned subtree, its visits in the order a thread runs them.  A visit arrives at a node
// with the partial X of the reduced tree and scores the two insertion edges below it.
//   x_in >= 0: X is a stored operand (first step of a side); -1: the partial the previous visit
//   kept in registers; <= -2: pop stack level (-2 - x_in).  keep 1 / 2: the partial toward d1 / d2
//   continues in registers; push >= 0: the other one is parked at that stack level.
//   tree1 / tree2: cost accumulator of the neighbour inserted into edge x--d1 / x--d2, -1 = none.
struct FitchWalkVisit {
  int x_in, d1, d2, tree1, tree2, keep, push, pad;
};
struct FitchWalkGroup {
  int tv, first, count, pad;
};

// Device-side planning of a full SPR neighbourhood (spr_plan_kernel): what the host uploads
// instead of the visit list.  One SprEdgeRec per directed edge 3*x+j of the base tree, one
// SprSearchRec per (pruned subtree, side).
struct SprEdgeRec {
  int c;          // nb[x][j], -1 = none
  int far_slot;   // operand id of D[c->x], the partial of c's side
  int far_sub;    // internal nodes on c's side = visits of a search that continues into c
  int child_end;  // rooted name of the edge (out_moves)
};
struct SprSearchRec {
  int start, from;       // first node reached and the suppressed node u it is reached from
  int x_in;              // operand id of D[other->u], the partial that enters `start`
  int visit_off;         // first visit of the search
  int uniq_base;         // unique id of the search's first kept move
  int move_cnt;          // 2 per visit
  int dup_flag;          // bit e (e < 2): the e-th move repeats an earlier topology; bit 2: the complement moves
  int pe;                // rooted name of the pruned branch (out_moves)
};

struct FitchProgram;
struct FitchNbProgram;
struct FitchWorkspace {
  DevBuf<unsigned int> scratch;  // internal-node vectors, slot-major
  DevBuf<FitchOp> ops;
  DevBuf<int> children;
  DevBuf<unsigned int> cost;
  DevBuf<int32_t> walk_desc;    // tree walks (treewalk.h): raw descriptors, planned ops, planner scratch
  DevBuf<int4> walk_ops;
  DevBuf<int> walk_scratch;
  DevBuf<int32_t> walk_rows;
  DevBuf<int4> nbw_visits;      // search walks: FitchWalkVisit[] (two int4 each), FitchWalkGroup[]
  DevBuf<int4> nbw_groups;
  DevBuf<int4> spr_edges, spr_searches, spr_moves;  // device-planned SPR neighbourhoods
  PinBuf<int4> spr_moves_host;                      // move table staging (pinned)
};

// Scratch pool shared by every alignment handle on the current device (grow-only; dropping a
// handle must not free tens of GB that the next call would re-allocate).
FitchWorkspace &fitch_workspace();

struct FitchAln {
  int64_t P = 0;        // patterns
  int n_cols = 0;       // alignment rows = tip slots
  int K = 0, K_real = 0;  // planes allocated (template instance) / distinct states seen
  int64_t W4 = 0, words32 = 0;
  int n_wbits = 0;
  DevBuf<unsigned int> tips;      // [n_cols][K][words32]
  DevBuf<unsigned int> wbits;     // [n_wbits][words32]
  DevBuf<unsigned char> chunk_nb; // [W4] weight bit-planes per 128-pattern chunk, 0xFF = all weights 1
  FitchAln(int64_t n_prof, int n_cols, const uint8_t *states, const int64_t *weights);
};

struct FitchProgram {
  HostArr<FitchOp> ops;            // sorted by level (pinned: uploaded asynchronously)
  HostArr<int> children;           // slot ids: < n_cols tip row, else n_cols + scratch slot
  std::vector<int64_t> level_off;  // [n_levels+1] into ops
  int64_t n_scratch_slots = 0;
  int64_t n_costs = 0;
  HostArr<FitchWalkVisit> walk_visits;  // run after the op levels
  HostArr<FitchWalkGroup> walk_groups;  // longest first
  // device-planned SPR neighbourhood: walk_visits stays empty, spr_plan_kernel writes the visits
  HostArr<SprEdgeRec> spr_edges;
  HostArr<SprSearchRec> spr_searches;
  int64_t spr_n_visits = 0, spr_n_unique = 0;
  int64_t spr_lo = 0, spr_hi = 0;  // unique ids scored by this program (accumulator spr_first_acc + id - spr_lo)
  int spr_first_acc = 0;
  double spr_vecs = 0, spr_req_vecs = 0;  // algorithmic / requested vectors of the walk launch
  FitchProgram() = default;
  FitchProgram(const FitchProgram &) = delete;
  FitchProgram &operator=(const FitchProgram &) = delete;
  void clear();
};

void build_full_program(const TreeBatch &b, int64_t t0, int64_t t1, int n_tip_slots, FitchProgram &prog);
// out_moves (device-planned SPR neighbourhoods only): int32 [spr_n_unique][4] on the host, or null
void run_fitch_program(const FitchAln &a, FitchWorkspace &ws, const FitchProgram &prog, int32_t *out_costs,
                       cudaStream_t st, int32_t *out_moves = nullptr);
void score_batch_full(FitchAln &a, const TreeBatch &b, int32_t *out_costs, cudaStream_t st);
// binary descriptor batches by on-chip tree walks (PLG_TREE_MODE=levels selects the level programs)
bool tree_walk_enabled();
void score_desc_walk(FitchAln &a, int64_t n_trees, int n_tips, const int32_t *desc, const int32_t *tip_rows,
                     int32_t *out_costs, cudaStream_t st);

}  // namespace plg

// opaque handle of the C-ABI
struct plg_fitch_aln {
  plg::FitchAln impl;
  plg_fitch_aln(int64_t P, int n_cols, const uint8_t *s, const int64_t *w) : impl(P, n_cols, s, w) {}
};
