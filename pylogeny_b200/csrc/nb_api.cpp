// C-ABI entry points that score a whole rearrangement neighbourhood in one call
// (replaces the per-neighbour loop landscape.py:724-756 and, for likelihood, the closest
// reference analogue libpllWrapper.c:241-319 getSPR/NNIMovesInDistance).
#include <algorithm>

#include "hostpool.h"
#include "neighbourhood.h"
#include "treewalk.h"

// DNA x Gamma-4 neighbourhood planner: depth-first walks (lik_walkm.cuh walk4m_kernel) ship; the
// level-by-level planner stays selectable as the A/B reference (PLG_LIK_NB_MODE=ops; 64 taxa x 10k
// sites per step on B200: 12.4 ms against 5.6 ms).

#include <chrono>
#include <cstdlib>
#include <cstring>
#include <unordered_map>

using namespace plg;

namespace {


// move_type = kind | radius << 8 (PLG_MOVE_SPR_WITHIN): search depth of the SPR enumeration
int depth_for(int move_type) {
  const int kind = move_type & 0xFF, radius = move_type >> 8;
  if (radius < 0) PLG_FAIL(PLG_ERR_ARG, "negative search radius");
  if (kind == PLG_MOVE_NNI) return 1;
  if (kind == PLG_MOVE_SPR) return radius == 0 ? -1 : radius;
  PLG_FAIL(PLG_ERR_UNSUPPORTED, "move type %d: NNI (2), SPR (1) and TBR (3) neighbourhoods are built in", move_type);
}

void fill_moves(const Neighbourhood &nb, int32_t *out_moves) {
  if (!out_moves) return;
  const UTree &t = nb.t;
  const int64_t M = (int64_t)nb.unique.size();
  const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>(HostPool::get().size(), M / 16384));
  HostPool::get().run(n_thr, [&](int ti, int nt) {
    for (int64_t i = M * ti / nt; i < M * (ti + 1) / nt; i++) {
      const Move &m = nb.moves[nb.unique[i]];
      const int v = t.nb[m.u][m.k];
      const int pe = t.rooted_child_end[3 * m.u + m.k];
      out_moves[4 * i + 0] = pe;                 // child end of the pruned branch (base-tree node id)
      out_moves[4 * i + 1] = pe == v ? 0 : 1;    // 0: the subtree below it moves; 1: its complement moves
      out_moves[4 * i + 2] = t.rooted_child_end[3 * m.x + m.kx];  // child end of the target branch
      out_moves[4 * i + 3] = m.depth;            // 1 = NNI
    }
  });
}

void shard_range(int64_t M, int64_t idx, int64_t cnt, int64_t *lo, int64_t *hi) {
  if (cnt < 1 || idx < 0 || idx >= cnt) PLG_FAIL(PLG_ERR_ARG, "bad shard %lld of %lld", (long long)idx, (long long)cnt);
  *lo = M * idx / cnt;
  *hi = M * (idx + 1) / cnt;
}


void fill_tbr_moves(const TbrNeighbourhood &tb, int32_t *out_moves) {
  if (!out_moves) return;
  const UTree &t = tb.t;
  for (size_t i = 0; i < tb.pairs.size(); i++) {
    const TbrCand &c1 = tb.cands[tb.pairs[i].c1], &c2 = tb.cands[tb.pairs[i].c2];
    out_moves[4 * i + 0] = t.rooted_child_end[3 * c1.u + c1.k];                       // bisected branch
    out_moves[4 * i + 1] = c1.x < 0 ? -1 : t.rooted_child_end[3 * c1.x + c1.kx];      // reconnection edge, part 1
    out_moves[4 * i + 2] = c2.x < 0 ? -1 : t.rooted_child_end[3 * c2.x + c2.kx];      // reconnection edge, part 2
    out_moves[4 * i + 3] = c1.depth + c2.depth;
  }
}

// edges [e0, e1) whose pairs intersect [lo, hi), cut into chunks that fit `max_slots`
template <typename Run>
void for_tbr_chunks(const TbrNeighbourhood &tb, int64_t lo, int64_t hi, int64_t max_slots, bool t20, Run run) {
  const int n_edges = (int)tb.edges.size();
  int e = 0;
  while (e < n_edges && tb.edge_pair_off[e + 1] <= lo) e++;
  while (e < n_edges && tb.edge_pair_off[e] < hi) {
    int e1 = e + 1;
    while (e1 < n_edges && tb.edge_pair_off[e1] < hi &&
           (t20 ? tbr20_slots_needed(tb, e, e1 + 1) : tbr_slots_needed(tb, e, e1 + 1)) <= max_slots) e1++;
    run(e, e1);
    e = e1;
  }
}

int score_tbr_fitch(FitchAln &a, int n_tips, const int32_t *desc, const int32_t *tip_rows, int64_t shard_index,
                    int64_t shard_count, int64_t *n_moves, int32_t *out_moves, int32_t *out_costs, cudaStream_t st) {
  TbrNeighbourhood tb = enumerate_tbr(utree_from_desc(n_tips, desc, nullptr));
  const int64_t M = (int64_t)tb.pairs.size();
  *n_moves = M;
  fill_tbr_moves(tb, out_moves);
  if (!out_costs || M == 0) return PLG_OK;
  int64_t lo, hi;
  shard_range(M, shard_index, shard_count, &lo, &hi);
  size_t free_b = 0, total_b = 0;
  PLG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t vec_bytes = (size_t)a.K * a.words32 * 4;
  const size_t budget = std::min<size_t>((free_b + fitch_workspace().scratch.n * 4) / 2, (size_t)64 << 30);
  const int64_t max_slots = std::max<int64_t>(8 * n_tips, (int64_t)(budget / vec_bytes));
  FitchTbrProgram p;
  std::vector<int32_t> acc, costs;
  for_tbr_chunks(tb, lo, hi, max_slots, false, [&](int e0, int e1) {
    build_fitch_tbr_program(tb, tip_rows, a.n_cols, e0, e1, p);
    acc.resize(p.prog.n_costs);
    run_fitch_program(a, fitch_workspace(), p.prog, acc.data(), st);
    const int64_t p0 = tb.edge_pair_off[e0], p1 = tb.edge_pair_off[e1];
    costs.resize(p1 - p0);
    finish_fitch_tbr(tb, p, acc.data(), e0, e1, costs.data());
    for (int64_t i = std::max(p0, lo); i < std::min(p1, hi); i++) out_costs[i] = costs[i - p0];
  });
  return PLG_OK;
}

int score_tbr_lik(LikAln &a, const Model &m, int n_tips, const int32_t *desc, const double *brlen,
                  const int32_t *tip_rows, int64_t shard_index, int64_t shard_count, int64_t *n_moves,
                  int32_t *out_moves, double *out_lnl, cudaStream_t st) {
  StageTimer tm;
  TbrNeighbourhood tb = enumerate_tbr(utree_from_desc(n_tips, desc, brlen));
  tm.lap("tbr enumerate");
  const int64_t M = (int64_t)tb.pairs.size();
  *n_moves = M;
  fill_tbr_moves(tb, out_moves);
  if (!out_lnl || M == 0) return PLG_OK;
  int64_t lo, hi;
  shard_range(M, shard_index, shard_count, &lo, &hi);
  const size_t slot_bytes = ((size_t)m.R * a.K * 8 + 4) * a.Ppad;
  LikWorkspace &wsp = lik_workspace();
  const size_t held = wsp.clv.n * 8 + wsp.scl.n * 4 + wsp.pairvec.n * 8 + wsp.pair_scl.n * 4;
  // (cudaMemGetInfo is a driver call of 0.1 - 2 ms: skipped once the pools hold what the whole neighbourhood needs)
  const size_t need_all = (size_t)(3 * (int64_t)tb.t.n_nodes + 3 * (int64_t)tb.cands.size() + 9 * (int64_t)tb.t.n_nodes) * slot_bytes;
  size_t free_b = 0, total_b = 0;
  if (need_all > held) PLG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t budget = need_all <= held ? need_all : std::min<size_t>((free_b + held) / 2, (size_t)80 << 30);
  int64_t max_slots = std::max<int64_t>(8 * n_tips, (int64_t)(budget / slot_bytes));
  if (const char *e = std::getenv("PLG_TBR_MAX_SLOTS")) max_slots = std::max<int64_t>(1, atoll(e));  // tests: several chunks of bisections
  tm.lap("tbr moves+mem");
  LikProgram p;
  std::vector<double> lnl;
  // 20 states x Gamma-4: fused candidate steps + pair grid on the FP64 tensor pipe (lik_tbr20.cuh);
  // PLG_TBR20=0 keeps the level program of plain updates and row evaluations (A/B reference)
  const char *t20_env = std::getenv("PLG_TBR20");
  const bool t20 = a.K == 20 && m.R == 4 && !(t20_env && !strcmp(t20_env, "0"));
  for_tbr_chunks(tb, lo, hi, max_slots, t20, [&](int e0, int e1) {
    if (t20) build_lik_tbr20_program(tb, tip_rows, a.n_rows, e0, e1, p);
    else build_lik_tbr_program(tb, tip_rows, a.n_rows, e0, e1, p);
    tm.lap("tbr build");
    const int64_t p0 = tb.edge_pair_off[e0], p1 = tb.edge_pair_off[e1];
    lnl.resize(p1 - p0);
    run_lik_program(a, m, lik_workspace(), p, lnl.data(), st);
    tm.lap("tbr run");
    for (int64_t i = std::max(p0, lo); i < std::min(p1, hi); i++) out_lnl[i] = lnl[i - p0];
  });
  return PLG_OK;
}

}  // namespace

namespace plg {
static int lik_nb_mode() {
  const char *e = std::getenv("PLG_LIK_NB_MODE");
  if (e && !strcmp(e, "ops")) return LIK_NB_OPS;
  return LIK_NB_WALK;
}
// lnL of the distinct neighbours [lo, hi) of an enumerated SPR/NNI neighbourhood, in chunks that
// fit the device (out_lnl is indexed by neighbour, only [lo, hi) is written).
void score_spr_lik(LikAln &a, Model &m, const Neighbourhood &nb, const int32_t *tip_rows, int64_t lo, int64_t hi,
                   double *out_lnl, cudaStream_t stream) {
  size_t free_b = 0, total_b = 0;
  PLG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t slot_bytes = ((size_t)m.R * a.K * 8 + 4) * a.Ppad;
  const size_t held = lik_workspace().clv.n * 8 + lik_workspace().scl.n * 4;
  const size_t budget = std::min<size_t>((free_b + held) / 2, (size_t)80 << 30);
  const int64_t max_slots = std::max<int64_t>(4 * nb.t.n_tips, (int64_t)(budget / slot_bytes));
  LikProgram p;
  StageTimer tm;
  int64_t chunk = hi - lo;
  for (int64_t c0 = lo; c0 < hi;) {
    int64_t c1 = std::min(hi, c0 + chunk);
    const bool dna4 = a.K == 4 && m.R == 4;
    build_lik_nb_program(nb, tip_rows, a.n_rows, c0, c1, p, dna4 ? lik_nb_mode() : LIK_NB_OPS, !dna4);
    if (p.n_slots > max_slots && c1 - c0 > 1) {
      chunk = std::max<int64_t>(1, (c1 - c0) / 2);
      continue;
    }
    tm.lap("  build");
    run_lik_program(a, m, lik_workspace(), p, out_lnl + c0, stream);
    tm.lap("  run");
    c0 = c1;
  }
}
}  // namespace plg

extern "C" int plg_neighbourhood_size(int move_type, int n_tips, const int32_t *desc, int64_t *n_moves) {
  PLG_API_BEGIN
  if (!desc || !n_moves) PLG_FAIL(PLG_ERR_ARG, "null argument");
  if (move_type == PLG_MOVE_TBR) {  // no closed form: enumerate
    *n_moves = (int64_t)enumerate_tbr(utree_from_desc(n_tips, desc, nullptr)).pairs.size();
    return PLG_OK;
  }
  const int depth = depth_for(move_type);
  UTree t = utree_from_desc(n_tips, desc, nullptr);  // validates the descriptor
  if (depth > 1) {  // radius-limited SPR: depends on the tree shape
    *n_moves = (int64_t)enumerate_spr(t, depth).unique.size();
    return PLG_OK;
  }
  // closed forms for an unrooted binary tree (checked against the enumeration in the tests)
  const int64_t n = n_tips;
  *n_moves = n < 4 ? 0 : (depth == 1 ? 2 * (n - 3) : 2 * (n - 3) * (2 * n - 7));
  PLG_API_END
}

extern "C" int plg_neighbourhood_trees(int move_type, int n_tips, const int32_t *desc, const double *brlen,
                                       const int32_t *tip_rows, int64_t n_sel, const int64_t *sel,
                                       int32_t *out_desc, double *out_brlen, uint64_t *out_hash) {
  PLG_API_BEGIN
  if (!desc || (n_sel > 0 && !sel)) PLG_FAIL(PLG_ERR_ARG, "null argument");
  const size_t stride = (size_t)(n_tips - 1) * 2;
  if (move_type == PLG_MOVE_TBR) {
    TbrNeighbourhood tb = enumerate_tbr(utree_from_desc(n_tips, desc, brlen, tip_rows));
    for (int64_t i = 0; i < n_sel; i++) {
      if (sel[i] < 0 || sel[i] >= (int64_t)tb.pairs.size()) PLG_FAIL(PLG_ERR_ARG, "move index out of range");
      if (out_desc) {
        UTree r = apply_tbr(tb, tb.pairs[sel[i]]);
        desc_from_utree(r, out_desc + i * stride, out_brlen ? out_brlen + i * stride : nullptr);
      }
      if (out_hash) out_hash[i] = tb.pairs[sel[i]].hash;
    }
    return PLG_OK;
  }
  if (depth_for(move_type) < 0 && n_sel <= 64 && n_tips >= 4) {  // a few moves (the winner of a climb step)
    const UTree t = utree_from_desc(n_tips, desc, brlen, tip_rows);
    std::vector<Move> mv(n_sel);
    spr_moves_by_unique_id(t, n_sel, sel, mv.data());
    for (int64_t i = 0; i < n_sel; i++) {
      if (out_desc) {
        UTree r = apply_move(t, mv[i]);
        desc_from_utree(r, out_desc + i * stride, out_brlen ? out_brlen + i * stride : nullptr);
      }
      if (out_hash) out_hash[i] = mv[i].hash;
    }
    return PLG_OK;
  }
  static thread_local Neighbourhood nb;  // storage reused from call to call
  enumerate_spr(utree_from_desc(n_tips, desc, brlen, tip_rows), depth_for(move_type), nb);
  for (int64_t i = 0; i < n_sel; i++) {
    if (sel[i] < 0 || sel[i] >= (int64_t)nb.unique.size()) PLG_FAIL(PLG_ERR_ARG, "move index out of range");
    const Move &m = nb.moves[nb.unique[sel[i]]];
    if (out_desc) {
      UTree r = apply_move(nb.t, m);
      desc_from_utree(r, out_desc + i * stride, out_brlen ? out_brlen + i * stride : nullptr);
    }
    if (out_hash) out_hash[i] = m.hash;
  }
  PLG_API_END
}

extern "C" int plg_neighbourhood_all_moves(int move_type, int n_tips, const int32_t *desc, int64_t *n_all,
                                          int32_t *out_moves) {
  PLG_API_BEGIN
  if (!desc || !n_all) PLG_FAIL(PLG_ERR_ARG, "null argument");
  static thread_local Neighbourhood nb;
  enumerate_spr(utree_from_desc(n_tips, desc, nullptr), depth_for(move_type), nb);
  const int64_t A = (int64_t)nb.moves.size();
  *n_all = A;
  if (!out_moves) return PLG_OK;
  std::unordered_map<uint64_t, int> first;  // repeats exist at distance 1 only (enumerate_spr)
  for (int64_t i = 0; i < A; i++)
    if (nb.moves[i].depth == 1 && nb.moves[i].unique_id >= 0) first.emplace(nb.moves[i].hash, nb.moves[i].unique_id);
  const UTree &t = nb.t;
  for (int64_t i = 0; i < A; i++) {
    const Move &m = nb.moves[i];
    const int pe = t.rooted_child_end[3 * m.u + m.k];
    int id = m.unique_id;
    if (id < 0) {
      auto it = first.find(m.hash);
      if (it == first.end()) PLG_FAIL(PLG_ERR_ARG, "internal: repeated move %lld has no first occurrence", (long long)i);
      id = it->second;
    }
    out_moves[5 * i + 0] = pe;
    out_moves[5 * i + 1] = pe == t.nb[m.u][m.k] ? 0 : 1;
    out_moves[5 * i + 2] = t.rooted_child_end[3 * m.x + m.kx];
    out_moves[5 * i + 3] = m.depth;
    out_moves[5 * i + 4] = id;
  }
  PLG_API_END
}

extern "C" int plg_topology_hash(int n_tips, const int32_t *desc, const int32_t *tip_rows, uint64_t *out) {
  PLG_API_BEGIN
  if (!desc || !out) PLG_FAIL(PLG_ERR_ARG, "null argument");
  *out = topology_hash(utree_from_desc(n_tips, desc, nullptr, tip_rows));
  PLG_API_END
}

static int fitch_score_neighbourhood_impl(plg_fitch_aln *aln, int move_type, int n_tips, const int32_t *desc,
                                         const int32_t *tip_rows, int64_t shard_index, int64_t shard_count,
                                         int64_t *n_moves, int32_t *out_moves, int32_t *out_costs, void *stream) {
  {
  if (!aln || !desc || !n_moves) PLG_FAIL(PLG_ERR_ARG, "null argument");
  FitchAln &a = aln->impl;
  for (int x = 0; x < n_tips; x++) {
    int row = tip_rows ? tip_rows[x] : x;
    if (row < 0 || row >= a.n_cols) PLG_FAIL(PLG_ERR_ARG, "tip row %d is outside the alignment", row);
  }
  StageTimer tm;
  if (move_type == PLG_MOVE_TBR)
    return score_tbr_fitch(a, n_tips, desc, tip_rows, shard_index, shard_count, n_moves, out_moves, out_costs,
                           (cudaStream_t)stream);
  const int depth = depth_for(move_type);
  const char *mode_env = std::getenv("PLG_FITCH_NB_MODE");
  const bool walk = !(mode_env && !strcmp(mode_env, "ops")) && n_tips <= (1 << WALK_MAX_LEVELS);
  // Full-radius SPR neighbourhood (whole, or one shard of it): the host prepares O(n) records and
  // the search walks are laid out by spr_plan_kernel (PLG_FITCH_NB_PLAN=host keeps the host planner
  // for A/B tests; NNI and radius-limited searches use it too).
  const char *plan_env = std::getenv("PLG_FITCH_NB_PLAN");
  if (walk && depth < 0 && n_tips >= 4 && !(plan_env && !strcmp(plan_env, "host"))) {
    static thread_local SprDevicePlan plan;  // keeps its pinned staging across calls
    static thread_local HostArr<int32_t> acc;  // pinned: the 1 MB of accumulators comes back at full PCIe rate
    const UTree t = utree_from_desc(n_tips, desc, nullptr);
    const int64_t M = 2 * ((int64_t)n_tips - 3) * (2 * (int64_t)n_tips - 7);  // distinct SPR neighbours
    int64_t lo, hi;
    shard_range(M, shard_index, shard_count, &lo, &hi);
    const bool whole = lo == 0 && hi == M;
    prepare_spr_device_plan(t, tip_rows, a.n_cols, plan, lo, hi);
    tm.lap("prepare");
    if (plan.n_unique != M) PLG_FAIL(PLG_ERR_ARG, "internal: %lld distinct SPR neighbours planned, %lld expected",
                                     (long long)plan.n_unique, (long long)M);
    *n_moves = M;
    if (out_moves && (!out_costs || !whole)) {  // the move table of the WHOLE neighbourhood: host enumeration
      static thread_local Neighbourhood nb0;
      enumerate_spr(t, depth, nb0);
      fill_moves(nb0, out_moves);
    }
    if (!out_costs || hi == lo) return PLG_OK;
    acc.resize(plan.base.prog.n_costs);
    run_fitch_program(a, fitch_workspace(), plan.base.prog, acc.data(), (cudaStream_t)stream, whole ? out_moves : nullptr);
    tm.lap("run");
    finish_fitch_spr_plan(t, plan, acc.data(), out_costs);
    tm.lap("finish");
    return PLG_OK;
  }
  static thread_local Neighbourhood nb;  // storage reused from call to call (one neighbourhood per climb step)
  enumerate_spr(utree_from_desc(n_tips, desc, nullptr), depth, nb);
  tm.lap("enumerate");
  const int64_t M = (int64_t)nb.unique.size();
  *n_moves = M;
  fill_moves(nb, out_moves);
  tm.lap("fill_moves");
  if (!out_costs || M == 0) return PLG_OK;
  int64_t lo, hi;
  shard_range(M, shard_index, shard_count, &lo, &hi);
  size_t free_b = 0, total_b = 0;
  PLG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t vec_bytes = (size_t)a.K * a.words32 * 4;
  const size_t budget = std::min<size_t>((free_b + fitch_workspace().scratch.n * 4) / 2, (size_t)64 << 30);
  const int64_t max_slots = std::max<int64_t>(4 * n_tips, (int64_t)(budget / vec_bytes));
  static thread_local FitchNbProgram p;  // keeps its pinned staging across calls
  std::vector<int32_t> acc;
  // PLG_FITCH_NB_MODE=ops: one level-synchronous op per search step (first formulation);
  // default: depth-first search walks
  int64_t chunk = hi - lo;
  for (int64_t c0 = lo; c0 < hi;) {
    int64_t c1 = std::min(hi, c0 + chunk);
    if (walk) build_fitch_nb_walk_program(nb, tip_rows, a.n_cols, c0, c1, p);
    else build_fitch_nb_program(nb, tip_rows, a.n_cols, c0, c1, p);
    if (p.prog.n_scratch_slots > max_slots && c1 - c0 > 1) {
      chunk = std::max<int64_t>(1, (c1 - c0) / 2);
      continue;
    }
    tm.lap("build");
    acc.resize(p.prog.n_costs);
    run_fitch_program(a, fitch_workspace(), p.prog, acc.data(), (cudaStream_t)stream);
    tm.lap("run");
    finish_fitch_nb(nb, p, acc.data(), c0, c1, out_costs + c0);
    tm.lap("finish");
    c0 = c1;
  }
  }
  return PLG_OK;
}

extern "C" int plg_fitch_score_neighbourhood(plg_fitch_aln *aln, int move_type, int n_tips, const int32_t *desc,
                                             const int32_t *tip_rows, int64_t shard_index, int64_t shard_count,
                                             int64_t *n_moves, int32_t *out_moves, int32_t *out_costs,
                                             void *stream) {
  PLG_API_BEGIN_LOCKED
  return fitch_score_neighbourhood_impl(aln, move_type, n_tips, desc, tip_rows, shard_index, shard_count, n_moves,
                                        out_moves, out_costs, stream);
  PLG_API_END
}

// Same with the costs written to DEVICE memory (out_where == PLG_OUT_DEVICE).  Fitch lengths are
// finished on the host (constant parts per search + device miss counts), so the block goes up again
// from pinned staging -- still one copy less for a caller who feeds a collective.
extern "C" int plg_fitch_score_neighbourhood_out(plg_fitch_aln *aln, int move_type, int n_tips, const int32_t *desc,
                                                 const int32_t *tip_rows, int64_t shard_index, int64_t shard_count,
                                                 int64_t *n_moves, int32_t *out_moves, int32_t *out_costs,
                                                 int out_where, void *stream) {
  PLG_API_BEGIN_LOCKED
  if (out_where == PLG_OUT_HOST || !out_costs)
    return fitch_score_neighbourhood_impl(aln, move_type, n_tips, desc, tip_rows, shard_index, shard_count, n_moves,
                                          out_moves, out_costs, stream);
  if (out_where != PLG_OUT_DEVICE) PLG_FAIL(PLG_ERR_ARG, "out_where: PLG_OUT_HOST or PLG_OUT_DEVICE");
  if (!n_moves) PLG_FAIL(PLG_ERR_ARG, "null argument");
  int64_t M = 0;
  const int rc0 = fitch_score_neighbourhood_impl(aln, move_type, n_tips, desc, tip_rows, shard_index, shard_count, &M,
                                                 out_moves, nullptr, stream);  // size (and the move table) first
  if (rc0 != PLG_OK) return rc0;
  static thread_local HostArr<int32_t> stage;
  stage.resize((size_t)std::max<int64_t>(1, M));
  const int rc = fitch_score_neighbourhood_impl(aln, move_type, n_tips, desc, tip_rows, shard_index, shard_count, n_moves,
                                                nullptr, stage.data(), stream);
  if (rc != PLG_OK) return rc;
  int64_t lo, hi;
  shard_range(M, shard_index, shard_count, &lo, &hi);
  if (hi > lo) {
    PLG_CUDA(cudaMemcpyAsync(out_costs + lo, stage.data() + lo, (size_t)(hi - lo) * sizeof(int32_t), cudaMemcpyHostToDevice,
                             (cudaStream_t)stream));
    PLG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  }
  PLG_API_END
}

// out_dev: out_lnl is device memory.  The device-planned SPR walk writes there directly; the other
// planners (TBR, NNI, radius-limited, 20 states) finish on the host and go up again from staging.
static int lik_score_neighbourhood_impl(plg_lik_aln *aln, plg_model *m, int move_type, int n_tips,
                                        const int32_t *desc, const double *brlen, const int32_t *tip_rows,
                                        int64_t shard_index, int64_t shard_count, int64_t *n_moves,
                                        int32_t *out_moves, double *out_lnl, bool out_dev, void *stream) {
  {
  if (!aln || !m || !desc || !brlen || !n_moves) PLG_FAIL(PLG_ERR_ARG, "null argument");
  LikAln &a = aln->impl;
  for (int x = 0; x < n_tips; x++) {
    int row = tip_rows ? tip_rows[x] : x;
    if (row < 0 || row >= a.n_rows) PLG_FAIL(PLG_ERR_ARG, "tip row %d is outside the alignment", row);
  }
  StageTimer tm;
  static thread_local HostArr<double> stage;  // host-finished planners with a device output
  auto stage_up = [&](int64_t M) {
    int64_t lo, hi;
    shard_range(M, shard_index, shard_count, &lo, &hi);
    if (hi > lo) {
      PLG_CUDA(cudaMemcpyAsync(out_lnl + lo, stage.data() + lo, (size_t)(hi - lo) * 8, cudaMemcpyHostToDevice, (cudaStream_t)stream));
      PLG_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
    }
  };
  if (move_type == PLG_MOVE_TBR) {
    if (!(out_dev && out_lnl))
      return score_tbr_lik(a, m->impl, n_tips, desc, brlen, tip_rows, shard_index, shard_count, n_moves, out_moves,
                           out_lnl, (cudaStream_t)stream);
    int64_t M = 0;
    score_tbr_lik(a, m->impl, n_tips, desc, brlen, tip_rows, shard_index, shard_count, &M, nullptr, nullptr, (cudaStream_t)stream);
    stage.resize((size_t)std::max<int64_t>(1, M));
    score_tbr_lik(a, m->impl, n_tips, desc, brlen, tip_rows, shard_index, shard_count, n_moves, out_moves, stage.data(),
                  (cudaStream_t)stream);
    stage_up(M);
    return PLG_OK;
  }
  const int depth = depth_for(move_type);
  // Full-radius SPR neighbourhood, whole or one shard (DNA x Gamma-4 walks): O(n) host records, the
  // search walks are laid out by spr_lik_plan_kernel (PLG_LIK_NB_PLAN=host keeps the host planner)
  const char *plan_env = std::getenv("PLG_LIK_NB_PLAN");
  if (out_lnl && depth < 0 && n_tips >= 4 && n_tips <= (1 << WALK_MAX_LEVELS) && a.K == 4 && m->impl.R == 4 &&
      lik_nb_mode() == LIK_NB_WALK && !(plan_env && !strcmp(plan_env, "host"))) {
    const size_t slot_bytes = ((size_t)m->impl.R * a.K * 8 + 4) * a.Ppad;
    const size_t held = lik_workspace().clv.n * 8 + lik_workspace().scl.n * 4;
    const size_t need = (size_t)3 * (n_tips - 2) * slot_bytes;
    size_t free_b = 0, total_b = 0;
    if (need > held) PLG_CUDA(cudaMemGetInfo(&free_b, &total_b));  // (a driver call: skipped once the pool is big enough)
    tm.lap("meminfo");
    if (need <= held || need <= std::min<size_t>((free_b + held) / 2, (size_t)80 << 30)) {
      static thread_local LikProgram p;
      const UTree t = utree_from_desc(n_tips, desc, brlen);
      const int64_t M = 2 * ((int64_t)n_tips - 3) * (2 * (int64_t)n_tips - 7);
      int64_t lo, hi;
      shard_range(M, shard_index, shard_count, &lo, &hi);
      const bool whole = lo == 0 && hi == M;
      build_lik_spr_device_plan(t, tip_rows, a.n_rows, p, lo, hi);
      tm.lap("prepare");
      if (p.spr_n_unique != M) PLG_FAIL(PLG_ERR_ARG, "internal: %lld distinct SPR neighbours planned, %lld expected",
                                        (long long)p.spr_n_unique, (long long)M);
      *n_moves = M;
      if (out_moves && !whole) {  // the move table of the WHOLE neighbourhood: host enumeration
        static thread_local Neighbourhood nb0;
        enumerate_spr(t, depth, nb0);
        fill_moves(nb0, out_moves);
      }
      if (hi > lo) {  // (out_lnl is indexed by neighbour: the shard's block starts at lo)
        OutDeviceScope scope(out_dev);
        run_lik_program(a, m->impl, lik_workspace(), p, out_lnl + lo, (cudaStream_t)stream, whole ? out_moves : nullptr);
      }
      tm.lap("run");
      return PLG_OK;
    }
  }
  static thread_local Neighbourhood nb;
  enumerate_spr(utree_from_desc(n_tips, desc, brlen), depth, nb);
  tm.lap("enumerate");
  const int64_t M = (int64_t)nb.unique.size();
  *n_moves = M;
  fill_moves(nb, out_moves);
  if (!out_lnl || M == 0) return PLG_OK;
  int64_t lo, hi;
  shard_range(M, shard_index, shard_count, &lo, &hi);
  if (out_dev) stage.resize((size_t)M);
  score_spr_lik(a, m->impl, nb, tip_rows, lo, hi, out_dev ? stage.data() : out_lnl, (cudaStream_t)stream);
  if (out_dev) stage_up(M);
  tm.lap("score");
  }
  return PLG_OK;
}

extern "C" int plg_lik_score_neighbourhood(plg_lik_aln *aln, plg_model *m, int move_type, int n_tips,
                                           const int32_t *desc, const double *brlen, const int32_t *tip_rows,
                                           int64_t shard_index, int64_t shard_count, int64_t *n_moves,
                                           int32_t *out_moves, double *out_lnl, void *stream) {
  PLG_API_BEGIN_LOCKED
  return lik_score_neighbourhood_impl(aln, m, move_type, n_tips, desc, brlen, tip_rows, shard_index, shard_count, n_moves,
                                      out_moves, out_lnl, false, stream);
  PLG_API_END
}

extern "C" int plg_lik_score_neighbourhood_out(plg_lik_aln *aln, plg_model *m, int move_type, int n_tips,
                                               const int32_t *desc, const double *brlen, const int32_t *tip_rows,
                                               int64_t shard_index, int64_t shard_count, int64_t *n_moves,
                                               int32_t *out_moves, double *out_lnl, int out_where, void *stream) {
  PLG_API_BEGIN_LOCKED
  if (out_where != PLG_OUT_HOST && out_where != PLG_OUT_DEVICE) PLG_FAIL(PLG_ERR_ARG, "out_where: PLG_OUT_HOST or PLG_OUT_DEVICE");
  return lik_score_neighbourhood_impl(aln, m, move_type, n_tips, desc, brlen, tip_rows, shard_index, shard_count, n_moves,
                                      out_moves, out_lnl, out_where == PLG_OUT_DEVICE, stream);
  PLG_API_END
}
