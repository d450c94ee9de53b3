// Rearrangement neighbourhoods on the host: unrooted view of a base tree, enumeration of the
// distinct NNI/SPR/TBR neighbours with canonical topology hashes, and the edge-insertion
// programs that score all of them from the base tree's directional partials.
//
// Reference behaviour this stands in for: rearrangement.py:542-594 (iterSPRForBranch,
// _flipOperation, allSPR), :440-511 (move -- branch lengths of the neighbour), and the
// per-neighbour scoring loop landscape.py:724-756.  The reference emits 4(n-3)(n-2) SPR moves of
// which 2(n-3)(2n-7) are distinct topologies; here each distinct topology is
// produced once (first occurrence in enumeration order).
#pragma once
#include <array>
#include <vector>

#include "fitch.h"
#include "lik.h"

namespace plg {

// Unrooted binary tree: tips 0..n-1 (degree 1), internal nodes n..2n-3 (degree 3).
struct UTree {
  int n_tips = 0, n_nodes = 0;
  std::vector<std::array<int, 3>> nb;      // neighbours, -1 = none
  std::vector<std::array<double, 3>> len;  // branch length to that neighbour
  std::vector<int> rooted_child_end;       // per directed edge 3*x+k: id (rooted desc numbering) naming the edge
  std::vector<int> tip_label;              // taxon identity of each tip for topology hashing (default: tip id)
  int slot_of(int x, int y) const {
    for (int k = 0; k < 3; k++)
      if (nb[x][k] == y) return k;
    return -1;
  }
  bool is_tip(int x) const { return x < n_tips; }
};

// Rooted binary descriptor (include/pylogeny_b200.h) -> unrooted tree; the root's two child
// branches merge into one edge whose length is their sum.
UTree utree_from_desc(int n_tips, const int32_t *desc, const double *brlen, const int32_t *tip_labels = nullptr);

struct Move {
  int u, k;        // pruned: the subtree on nb[u][k]'s side of edge {u, nb[u][k]}; u is suppressed
  int x, kx;       // target edge {x, nb[x][kx]}, x on the side nearer to u
  int depth;       // 1 = NNI distance
  int parent;      // index of the move one step nearer to u on the same search path, -1 at depth 1
  int side;        // which of u's two remaining neighbours the path left through (0/1)
  uint64_t hash;   // canonical topology hash of the neighbour
  int unique_id;   // index among distinct topologies, -1 if a duplicate of an earlier move
};

struct Neighbourhood {
  UTree t;
  std::vector<Move> moves;       // every (prune, target) pair in enumeration order
  std::vector<int> unique;       // indices into moves, first occurrence of each topology
  std::vector<int64_t> search_off;  // moves of search s = (pruned u, k, side): [search_off[s], search_off[s+1])
  uint64_t base_hash = 0;
};

// Canonical topology hash of an unrooted tree (sum over edges of a mixed split hash).
uint64_t topology_hash(const UTree &t);
// max_depth < 0: all SPR moves; 1: NNI.
Neighbourhood enumerate_spr(const UTree &t, int max_depth);
// same into a caller-owned neighbourhood whose storage is reused from call to call (a hill climb
// enumerates one neighbourhood per step; a fresh 12 MB move array costs more in page faults than
// the enumeration itself)
void enumerate_spr(const UTree &t, int max_depth, Neighbourhood &out);
// Moves `sel` (unique ids) of the FULL SPR neighbourhood without enumerating it (O(n) per move).
void spr_moves_by_unique_id(const UTree &t, int64_t n_sel, const int64_t *sel, Move *out);
// The neighbour tree of one move, with the reference's branch-length rule (rearrangement.py:466-497).
UTree apply_move(const UTree &t, const Move &m);
// Rooted post-order descriptor of an unrooted tree, rooted on tip 0's pendant edge
// (what rerootToLeaf does for the lowest-order leaf, rearrangement.py:267-331).
void desc_from_utree(const UTree &t, int32_t *desc, double *brlen);

// Edge-insertion programs over moves unique[lo..hi): directional partials of the base tree, one
// partial per search step, one join / 3-way evaluation per distinct neighbour.
enum { LIK_NB_OPS = 0, LIK_NB_WALK = 2 };

struct FitchNbProgram {
  FitchProgram prog;
  std::vector<int> dir_acc;                 // accumulator of directional op 3*x+k (-1 for tips)
  std::vector<std::array<int, 2>> dir_kids; // children directed edges of each directional op
  int64_t first_move_acc = 0;
};
void build_fitch_nb_program(const Neighbourhood &nb, const int32_t *tip_rows, int n_tip_slots, int64_t lo,
                            int64_t hi, FitchNbProgram &out);
// The same neighbours scored by depth-first search walks (FitchWalkVisit / FitchWalkGroup,
// fitch_walk.cuh): no scratch slot per search step, accumulator layout as above.
void build_fitch_nb_walk_program(const Neighbourhood &nb, const int32_t *tip_rows, int n_tip_slots, int64_t lo,
                                 int64_t hi, FitchNbProgram &out);
// Combines the downloaded accumulators into one Fitch length per move in [lo, hi).
void finish_fitch_nb(const Neighbourhood &nb, const FitchNbProgram &p, const int32_t *acc, int64_t lo, int64_t hi,
                     int32_t *out_costs);

// hoist_eval (mode LIK_NB_OPS): the far side of every target branch and every pruned subtree are
// pushed through their branches once (one-operand updates) instead of once per neighbour, so an
// evaluation multiplies one matrix instead of three (the 20-state kernels; the DNA level kernels
// fuse the whole evaluation into the update instead).
// mode: LIK_NB_OPS one CLV update + one 3-way evaluation per neighbour (any state count);
// LIK_NB_WALK depth-first search walks (LikWalkVisit, lik_walkm.cuh; DNA x Gamma-4).
void build_lik_nb_program(const Neighbourhood &nb, const int32_t *tip_rows, int n_rows, int64_t lo, int64_t hi,
                          LikProgram &out, int mode = LIK_NB_OPS, bool hoist_eval = false);

// Device-planned full SPR neighbourhood (parsimony).  The host does O(n) work only: the base tree's
// directional ops, one SprEdgeRec per directed edge, one SprSearchRec per search (pruned subtree
// u--nb[u][k], side) and the de-duplication of the 8(n-3) distance-1 moves (the only ones that can
// repeat).  The search walks themselves -- one FitchWalkVisit per node reached, in the same
// depth-first order and with the same neighbour numbering as enumerate_spr +
// build_fitch_nb_walk_program -- are laid out by spr_plan_kernel (fitch.cu), one thread per search.
struct SprDevicePlan {
  FitchNbProgram base;               // directional level ops + prog.spr_* (what run_fitch_program uploads)
  std::vector<int> order;            // directed edges, dependencies first (finish pass)
  std::vector<int> search_edge;      // per search: pruned directed edge 3*u+k
  std::vector<int64_t> search_uniq;  // per search: unique id of its first kept move; [n_search] = n_unique
  int64_t n_visits = 0, n_unique = 0;
  int64_t lo = 0, hi = 0;            // the unique ids this plan scores (a shard), [0, n_unique) = all
};
// [lo, hi): only searches that hold ids of the range are planned (shards of one neighbourhood).
void prepare_spr_device_plan(const UTree &t, const int32_t *tip_rows, int n_tip_slots, SprDevicePlan &out,
                             int64_t lo = 0, int64_t hi = INT64_MAX);
// Fitch length of the distinct neighbours [lo, hi) from the downloaded accumulators (no move array);
// out_costs is indexed by neighbour.
void finish_fitch_spr_plan(const UTree &t, const SprDevicePlan &p, const int32_t *acc, int32_t *out_costs);

// Likelihood twin: level ops of the base tree + the records spr_lik_plan_kernel expands (full SPR
// neighbourhood, all distinct neighbours; prog.n_trees = their number).
void build_lik_spr_device_plan(const UTree &t, const int32_t *tip_rows, int n_rows, LikProgram &prog,
                               int64_t lo = 0, int64_t hi = INT64_MAX);

// lnL of the distinct neighbours [lo, hi) of an enumerated SPR/NNI neighbourhood (nb_api.cpp).
void score_spr_lik(LikAln &a, Model &m, const Neighbourhood &nb, const int32_t *tip_rows, int64_t lo, int64_t hi,
                   double *out_lnl, cudaStream_t stream);

}  // namespace plg

// ---------------------------------------------------------------------------------------------
// TBR (tree bisection and reconnection).  The reference only names the operator
// (rearrangement.py:27 TYPE_TBR, :88-90 isTBR; allType raises for it, :661-664), so this is the
// textbook definition: remove an edge {u,v}, suppress u and v, reconnect any edge of one part
// with any edge of the other by a new branch (which keeps the removed branch's length; broken
// edges are halved and suppressed ones summed, the reference's SPR rule, rearrangement.py:466-497).
// SPR neighbours are the pairs in which one side keeps its original attachment.
namespace plg {

struct TbrCand {
  int u, k;       // this side is the part containing u after cutting edge {u, nb[u][k]}
  int x, kx;      // reconnect in the middle of edge {x, nb[x][kx]}; x < 0: the original attachment
  int depth, parent, side;
  uint64_t delta;  // topology hash change caused by this side alone
};

struct TbrPair {
  int c1, c2;      // indices into cands (c1 on the smaller-id end's side)
  uint64_t hash;
};

struct TbrNeighbourhood {
  UTree t;
  std::vector<TbrCand> cands;
  std::vector<int64_t> edge_cand_off;  // per undirected edge e: cands of side u in [off[2e], off[2e+1]), side v in [off[2e+1], off[2e+2])
  std::vector<std::array<int, 2>> edges;  // (u, k)
  std::vector<TbrPair> pairs;             // distinct neighbours, grouped by edge
  std::vector<int64_t> edge_pair_off;     // [n_edges + 1]
  uint64_t base_hash = 0;
};

TbrNeighbourhood enumerate_tbr(const UTree &t);
UTree apply_tbr(const TbrNeighbourhood &nb, const TbrPair &p);

struct FitchTbrProgram {
  FitchProgram prog;
  std::vector<int> dir_acc;
  std::vector<std::array<int, 2>> dir_kids;
  int64_t first_pair_acc = 0;
};
// Programs over the pairs of edges [e0, e1).
void build_fitch_tbr_program(const TbrNeighbourhood &nb, const int32_t *tip_rows, int n_tip_slots, int e0, int e1,
                             FitchTbrProgram &out);
void finish_fitch_tbr(const TbrNeighbourhood &nb, const FitchTbrProgram &p, const int32_t *acc, int e0, int e1,
                      int32_t *out_costs);
void build_lik_tbr_program(const TbrNeighbourhood &nb, const int32_t *tip_rows, int n_rows, int e0, int e1,
                           LikProgram &out);
// scratch slots a chunk of edges needs (for memory budgeting)
int64_t tbr_slots_needed(const TbrNeighbourhood &nb, int e0, int e1);
// 20 states x Gamma-4: fused candidate steps + pair grid (csrc/lik_tbr20.cuh); same pairs, same tree ids
void build_lik_tbr20_program(const TbrNeighbourhood &nb, const int32_t *tip_rows, int n_rows, int e0, int e1,
                             LikProgram &prog);
int64_t tbr20_slots_needed(const TbrNeighbourhood &nb, int e0, int e1);

}  // namespace plg
