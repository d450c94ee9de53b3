// Felsenstein-pruning likelihood engine: device-resident tips, CLV op programs, executor.
#pragma once
#include <unordered_map>

#include "common.h"
#include "model.h"
#include "tree.h"

namespace plg {

// dst = CLV( P[pa] * a ) .* ( P[pb] * b ); operand ids: < n_rows tip row, else n_rows + slot
struct LikOp {
  int dst, a, b, pa, pb;
};
// lnL[tree] = sum_p w_p log( (1/R) sum_r sum_k pi_k prod_o (P[po] x_o)_r[k] ) + scalers over the
// operands o in {a, b, c}; c < 0 = absent; pm < 0 = identity (the operand sits at the node itself)
struct LikEval {
  int a, pa, b, pb, c, pc, tree, pad;
};

// DNA x Gamma-4 update with an optional fused evaluation (csrc/lik_dna.cuh clv4_kernel); 48 bytes
struct LikOpF {
  int dst, a, b, pa, pb;             // as LikOp
  int tree;                          // >= 0: fused evaluation writes partial[tree]
  int e_pa, e_b, e_pb, e_c, e_pc;    // evaluation operands: (P[e_pa] N) (P[e_pb] b) (P[e_pc] c)
  int pad;                           // 1: do not store dst (no reader)
};

// Depth-first SPR search walks (csrc/lik_walkm.cuh).  One group = one pruned subtree (both sides
// of the pruning point); its visits are listed in the order one warp executes them.
//   x_in >= 0: the incoming partial is a stored operand (first step of a side); -1: it is the
//   partial the previous visit kept in registers; <= -2: pop it from stack level (-2 - x_in).
//   keep: 0 none, 1 / 2: the partial toward d1 / d2 continues in registers into the next visit;
//   push >= 0: the other partial is parked at that stack level.
struct LikWalkVisit {
  int x_in, p_in, d1, p_t1, p_h1, d2, p_t2, p_h2, tree1, tree2, keep, push, pad0, pad1, pad2, pad3;
};
struct LikWalkGroup {
  int tv, p_v, first, count;
};

// Device-side planning of a full SPR neighbourhood (spr_lik_plan_kernel): per directed edge 3*x+j
// and per (pruned subtree, side); the Fitch records (fitch.h SprEdgeRec / SprSearchRec) plus the
// P-matrix ids the likelihood visit needs.
struct LikSprEdge {
  int c, far_slot, far_sub, child_end;  // as SprEdgeRec
  int p_t, p_h, pad0, pad1;             // P-matrix of the branch x--c, and of its half
};
struct LikSprSearch {
  int start, from, x_in, visit_off;      // as SprSearchRec
  int uniq_base, move_cnt, dup_flag, pe;
  int p_in, tv, pad0, pad1;              // P-matrix of the merged branch other--start; operand id of the pruned subtree
};

// 20 states x Gamma-4 TBR neighbourhoods (csrc/lik_tbr20.cuh)
struct Tbr20Op {
  int a, pa;   // stage 1: X = P[pa] * operand a (standard layout or tip row) ...
  int b;       // ... o raw vector b (standard layout, already through its branch); -1: none
  int dst;     // >= 0: X is stored in the standard layout under this operand id (children read it)
  int p2, b2;  // p2 >= 0: stage 2: Y = (P[p2] * X) o raw b2 (b2 < 0: no join)
  int p3;      // p3 >= 0: stage 3: Z = P[p3] * Y
  int pdst;    // >= 0: pair-layout slot of the last stage's result; bit 30 set: times pi
};
struct Tbr20Tile {
  int a[8], b[8];  // pair slots of the tile's rows (stay side) and columns (pushed side)
  int tree[64];    // [row * 8 + col] -> tree id of the pair, -1: no such neighbour
};

struct LikWorkspace {
  DevBuf<double> clv;      // [slot][R*K][Ppad]
  DevBuf<int> scl;         // [slot][Ppad] per-pattern 2^-256 scaling counters
  DevBuf<double> pmat;     // [pm][R][K*K+1]  (padded stride: conflict-free shared-memory staging)
  DevBuf<double> pfrag;    // [pm][32 lanes][R] (DNA x Gamma-4): (P(t), P(t/2)) as DMMA B fragments (pairfrag4_kernel)
  DevBuf<double> tiptab;   // [pm][ncodes][R*K]  sum over the code's states of P columns
  DevBuf<double> brlen;    // [pm]
  DevBuf<LikOp> ops;
  DevBuf<char> fops;       // LikOpF[] (DNA x Gamma-4 path: updates with fused evaluations)
  DevBuf<LikWalkVisit> walk_visits;
  DevBuf<LikWalkGroup> walk_groups;
  DevBuf<double> walk_stack;   // [warp][level][16][32]
  DevBuf<int> walk_stack_scl;  // [warp][level][32]
  DevBuf<unsigned long long> walk_counter;
  DevBuf<int4> spr_edges, spr_searches, spr_moves;  // device-planned SPR neighbourhoods
  DevBuf<double> spr_acct;                          // byte accounting of the planned walk (profiling)
  DevBuf<int4> eval_rows;                           // LikEvalRow[] (lik_aa.cuh eval20_rows_kernel)
  DevBuf<double> pairvec;      // [pair slot][Ppad][80]: TBR candidates in the pair grid's layout (lik_tbr20.cuh)
  DevBuf<int> pair_scl;        // [pair slot][Ppad]
  DevBuf<Tbr20Op> t20_ops;
  DevBuf<Tbr20Tile> t20_tiles;
  DevBuf<int> t20_tile_off;
  DevBuf<int32_t> tw_desc;     // tree walks (treewalk.h): raw descriptors, planned ops, planner scratch
  DevBuf<int4> tw_ops;
  DevBuf<int> tw_scratch;
  DevBuf<int32_t> tw_rows;
  DevBuf<LikEval> evals;
  DevBuf<double> partial;  // [eval][blocks]
  DevBuf<int> partial_e;   // [eval][slices] exponents of the search walks' (mantissa, exponent) records
  DevBuf<double> lnl;      // [tree]
};

LikWorkspace &lik_workspace();  // shared per device, grow-only (see fitch_workspace)

struct LikAln {
  int K = 0, n_codes = 0, n_rows = 0;
  int64_t P = 0, Ppad = 0;
  DevBuf<unsigned char> codes;  // [n_rows][Ppad]; padding = the any-state code
  DevBuf<double> weights;       // [Ppad]; padding = 0
  DevBuf<unsigned int> masks;   // [n_codes] bit k = state k compatible
  std::vector<unsigned int> h_masks;
  LikAln(int datatype, int n_rows, int64_t P, const uint8_t *codes, const double *weights);
};

struct LikProgram {
  std::vector<LikOp> ops;          // sorted by level
  std::vector<int64_t> level_off;  // [n_levels+1]
  // fused TBR candidate steps (run after the op levels) and the pair grid that replaces `evals`
  std::vector<Tbr20Op> t20_ops;           // sorted by level
  std::vector<int64_t> t20_level_off;
  std::vector<Tbr20Tile> t20_tiles;       // grouped by bisected branch
  std::vector<int> t20_tile_off;          // [n_edges + 1]
  int64_t t20_pair_slots = 0;
  double t20_products = 0, t20_bytes = 0, t20_pair_bytes = 0;  // per pattern: 20x20x4 products, vector bytes moved
  std::vector<LikEval> evals;
  std::vector<LikWalkVisit> walk_visits;  // depth-first search walks, run after the op levels
  std::vector<LikWalkGroup> walk_groups;  // longest first
  int walk_levels = 0;                    // stack levels a walk needs
  // device-planned SPR neighbourhood: walk_visits stays empty, spr_lik_plan_kernel writes the visits
  std::vector<LikSprEdge> spr_edges;
  std::vector<LikSprSearch> spr_searches;
  int64_t spr_n_visits = 0, spr_n_unique = 0;
  int64_t spr_lo = 0, spr_hi = 0;         // unique ids scored by this program (tree id = unique id - spr_lo)
  std::vector<double> brlens;      // unique branch lengths -> P-matrix index
  std::unordered_map<uint64_t, int> pm_index;
  int64_t n_slots = 0, n_trees = 0;
  void clear();
  int pm(double t);
};

void build_full_lik_program(const TreeBatch &b, int64_t t0, int64_t t1, int n_rows, LikProgram &prog);
// out_moves (device-planned SPR neighbourhoods only): int32 [spr_n_unique][4] on the host, or null
void run_lik_program(const LikAln &a, const Model &m, LikWorkspace &ws, const LikProgram &prog, double *out_lnl,
                     cudaStream_t st, int32_t *out_moves = nullptr);
void score_lik_batch_full(LikAln &a, const Model &m, const TreeBatch &b, double *out_lnl, cudaStream_t st);

// Building blocks of the DNA x Gamma-4 level kernels for callers that keep their own device
// state (batched branch-length smoothing, pll_api.cu): everything is passed as device pointers.
//   dna4_pmatrices: P-matrix + tip table of entries idx[0..n) (idx == nullptr: entries 0..n) from
//   brlen[entry]; dna4_updates: n ops, independent of each other; dna4_evals: lnl[ev.tree] for n
//   evaluations (partial: >= n * Ppad / 128 doubles of scratch).
void dna4_pmatrices(const LikAln &a, const Model &m, const double *d_brlen, const int *d_idx, int64_t n,
                    double *d_pmat, double *d_tiptab, cudaStream_t st);
void dna4_updates(const LikAln &a, const Model &m, const LikOpF *d_ops, int64_t n, const double *d_pmat,
                  const double *d_tiptab, double *d_clv, int *d_scl, cudaStream_t st);
void dna4_evals(const LikAln &a, const Model &m, const LikEval *d_evals, int64_t n, const double *d_pmat,
                const double *d_tiptab, const double *d_clv, const int *d_scl, double *d_partial, double *d_lnl,
                cudaStream_t st);
// Batched branch-length smoothing, DNA x Gamma-4 (pll_api.cu): a re-orientation whose result is an end of the
// branch the NEXT action optimises is fused with that action's sum-table pass -- the new vector goes
// from the registers straight into the derivative sums (one 132-byte read per pattern less).
struct Eigen4 {
  double piU[16];   // pi_i U[i][k]
  double Uinv[16];  // [k][i]
  double g[16];     // -eval[k] * rate[r], [r][k]
};
struct ClvOptOp {
  int dst, a, b, pa, pb;       // the update, as LikOp
  int other, branch, tree;     // the branch's other end (operand id), its index, the tree's flag index
  int dslot, pad0, pad1, pad2; // slot of the derivative-sum shares (index in the next step's item list)
};
// n fused ops; dpart[(dslot * 3 + j) * blocks_per_op + block], blocks_per_op = Ppad / 256
void dna4_updates_opt(const LikAln &a, const Model &m, const ClvOptOp *d_ops, int64_t n, const double *d_pmat,
                      const double *d_tiptab, double *d_clv, int *d_scl, const Eigen4 &eg, const double *d_brlen,
                      const int *d_active, double *d_dpart, cudaStream_t st);
// the same for 20 states x Gamma-4 (lik_aa.cuh); P-matrix stride 4 * 401, tip table 23 * 80 per entry,
// partial: >= n * Ppad / 64 doubles
void aa20_pmatrices(const LikAln &a, const Model &m, const double *d_brlen, const int *d_idx, int64_t n,
                    double *d_pmat, double *d_tiptab, cudaStream_t st);
void aa20_updates(const LikAln &a, const Model &m, const LikOp *d_ops, int64_t n, const double *d_pmat,
                  double *d_clv, int *d_scl, cudaStream_t st);
void aa20_evals(const LikAln &a, const Model &m, const LikEval *d_evals, int64_t n, const double *d_pmat,
                const double *d_clv, const int *d_scl, double *d_partial, double *d_lnl, cudaStream_t st);

}  // namespace plg

// opaque handles of the C-ABI
struct plg_lik_aln {
  plg::LikAln impl;
  plg_lik_aln(int dt, int n, int64_t P, const uint8_t *c, const double *w) : impl(dt, n, P, c, w) {}
};
struct plg_model {
  plg::Model impl;
  plg_model(int K, const double *e, const double *f, double a, int n) : impl(K, e, f, a, n) {}
};
