#include "neighbourhood.h"

#include <algorithm>
#include <thread>
#include <unordered_map>

#include "hostpool.h"

namespace plg {

namespace {

inline uint64_t splitmix(uint64_t x) {
  x += 0x9E3779B97F4A7C15ull;
  x = (x ^ (x >> 30)) * 0xBF58476D1CE4E5B9ull;
  x = (x ^ (x >> 27)) * 0x94D049BB133111EBull;
  return x ^ (x >> 31);
}

// Directed edges of an unrooted tree in dependency order.  D[x->y] (the partial of x's side
// seen from y) needs D[a->x], D[b->x] for x's other neighbours a, b.
struct DirOrder {
  std::vector<int> order;  // directed edge ids 3*x+k with x internal, dependencies first
  std::vector<int> level;  // per directed edge id (tips: 0)
  int max_level = 0;
};

DirOrder dir_order(const UTree &t) {
  DirOrder d;
  d.level.assign((size_t)t.n_nodes * 3, 0);
  if (t.n_tips < 3) return d;
  // BFS from tip 0
  std::vector<int> par(t.n_nodes, -1), bfs;
  bfs.reserve(t.n_nodes);
  bfs.push_back(0);
  for (size_t i = 0; i < bfs.size(); i++) {
    int x = bfs[i];
    for (int k = 0; k < 3; k++) {
      int y = t.nb[x][k];
      if (y >= 0 && y != par[x]) {
        par[y] = x;
        bfs.push_back(y);
      }
    }
  }
  // down directions x -> par[x], leaves first
  for (size_t i = bfs.size(); i-- > 1;) {
    int x = bfs[i];
    if (t.is_tip(x)) continue;
    int k = t.slot_of(x, par[x]);
    int lv = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        int c = t.nb[x][j];
        lv = std::max(lv, d.level[3 * c + t.slot_of(c, x)]);
      }
    d.level[3 * x + k] = lv + 1;
    d.order.push_back(3 * x + k);
  }
  // up directions p -> x for internal p, root side first
  for (size_t i = 1; i < bfs.size(); i++) {
    int x = bfs[i], p = par[x];
    if (t.is_tip(p)) continue;  // p == tip 0: D[0->x] is the tip itself
    int k = t.slot_of(p, x);
    int lv = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        int c = t.nb[p][j];
        lv = std::max(lv, d.level[3 * c + t.slot_of(c, p)]);
      }
    d.level[3 * p + k] = lv + 1;
    d.order.push_back(3 * p + k);
  }
  for (int e : d.order) d.max_level = std::max(d.max_level, d.level[e]);
  return d;
}

// XOR of leaf keys on x's side of every directed edge, and their grand total.
std::vector<uint64_t> side_hashes(const UTree &t, const DirOrder &d, uint64_t *all) {
  std::vector<uint64_t> h((size_t)t.n_nodes * 3, 0);
  uint64_t tot = 0;
  for (int x = 0; x < t.n_tips; x++) {
    uint64_t key = splitmix((uint64_t)(t.tip_label.empty() ? x : t.tip_label[x]) + 1);
    tot ^= key;
    for (int k = 0; k < 3; k++) h[3 * x + k] = key;
  }
  for (int e : d.order) {
    int x = e / 3, k = e % 3;
    uint64_t v = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        int c = t.nb[x][j];
        v ^= h[3 * c + t.slot_of(c, x)];
      }
    h[e] = v;
  }
  *all = tot;
  return h;
}

inline uint64_t split_term(uint64_t side, uint64_t all) { return splitmix(std::min(side, side ^ all)); }

}  // namespace

UTree utree_from_desc(int n_tips, const int32_t *desc, const double *brlen, const int32_t *tip_labels) {
  if (n_tips < 2) PLG_FAIL(PLG_ERR_ARG, "a tree needs at least 2 tips");
  const int n = n_tips, ni = n - 1, root = n + ni - 1;
  std::vector<int> parent(2 * n - 1, -1);
  std::vector<double> plen(2 * n - 1, 0.0);
  for (int i = 0; i < ni; i++)
    for (int c = 0; c < 2; c++) {
      int ch = desc[2 * i + c];
      if (ch < 0 || ch >= n + i || parent[ch] >= 0)
        PLG_FAIL(PLG_ERR_ARG, "descriptor node %d: child %d is not a valid earlier node", i, ch);
      parent[ch] = n + i;
      plen[ch] = brlen ? brlen[2 * i + c] : 0.0;
    }
  for (int x = 0; x < root; x++)
    if (parent[x] < 0) PLG_FAIL(PLG_ERR_ARG, "descriptor does not connect node %d", x);
  UTree t;
  t.n_tips = n;
  t.n_nodes = 2 * n - 2;
  t.nb.assign(t.n_nodes, {-1, -1, -1});
  t.len.assign(t.n_nodes, {0.0, 0.0, 0.0});
  t.rooted_child_end.assign((size_t)t.n_nodes * 3, -1);
  t.tip_label.resize(n);
  for (int x = 0; x < n; x++) t.tip_label[x] = tip_labels ? tip_labels[x] : x;
  auto link = [&](int x, int y, double l, int child_end) {
    for (int pass = 0; pass < 2; pass++) {
      int a = pass ? y : x, b = pass ? x : y;
      int k = 0;
      while (k < 3 && t.nb[a][k] >= 0) k++;
      if (k == 3) PLG_FAIL(PLG_ERR_ARG, "descriptor is not binary at node %d", a);
      t.nb[a][k] = b;
      t.len[a][k] = l;
      t.rooted_child_end[3 * a + k] = child_end;
    }
  };
  // internal nodes list their two children first, then the parent side
  for (int i = 0; i < ni - 1; i++)
    for (int c = 0; c < 2; c++) {
      int ch = desc[2 * i + c];
      link(n + i, ch, plen[ch], ch);
    }
  const int c0 = desc[2 * (ni - 1)], c1 = desc[2 * (ni - 1) + 1];
  link(c0, c1, plen[c0] + plen[c1], c0);  // the edge through the root is named by its first child
  return t;
}

uint64_t topology_hash(const UTree &t) {
  DirOrder d = dir_order(t);
  uint64_t all;
  std::vector<uint64_t> h = side_hashes(t, d, &all);
  uint64_t s = 0;
  for (int x = 0; x < t.n_nodes; x++)
    for (int k = 0; k < 3; k++) {
      int y = t.nb[x][k];
      if (y > x) s += split_term(h[3 * x + k], all);
    }
  return s;
}

Neighbourhood enumerate_spr(const UTree &t, int max_depth) {
  Neighbourhood nb;
  enumerate_spr(t, max_depth, nb);
  return nb;
}

void enumerate_spr(const UTree &t, int max_depth, Neighbourhood &nb) {
  nb.t = t;
  DirOrder d = dir_order(t);
  uint64_t all;
  std::vector<uint64_t> h = side_hashes(t, d, &all);
  uint64_t base = 0;
  for (int x = 0; x < t.n_nodes; x++)
    for (int k = 0; k < 3; k++)
      if (t.nb[x][k] > x) base += split_term(h[3 * x + k], all);
  nb.base_hash = base;
  if (t.n_tips < 4) { nb.moves.clear(); nb.unique.clear(); nb.search_off.assign(1, 0); return; }
  // far[3*x+k] = hash of the side beyond edge x -> nb[x][k]; term[] = its split term
  std::vector<uint64_t> far((size_t)t.n_nodes * 3, 0), term((size_t)t.n_nodes * 3, 0);
  for (int x = 0; x < t.n_nodes; x++)
    for (int k = 0; k < 3; k++) {
      const int y = t.nb[x][k];
      if (y < 0) continue;
      far[3 * x + k] = h[3 * y + t.slot_of(y, x)];
      term[3 * x + k] = split_term(far[3 * x + k], all);
    }
  struct Frame { int x, from, parent_move, depth; uint64_t rem, add; };
  // Number of search steps of every (pruned subtree, side): all edges beyond `start` seen from u,
  // i.e. 2*leaves-2 of that side (depth-limited: 2 if start is internal).  Prefix sums give every
  // search its slice of the move array, so the host cores fill it in place.
  std::vector<int> leaves((size_t)t.n_nodes * 3, 0);  // leaves on x's side of edge x -> nb[x][k]
  for (int x = 0; x < t.n_tips; x++) leaves[3 * x] = leaves[3 * x + 1] = leaves[3 * x + 2] = 1;
  for (int e : d.order) {
    const int x = e / 3, k = e % 3;
    int c = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) c += leaves[3 * t.nb[x][j] + t.slot_of(t.nb[x][j], x)];
    leaves[e] = c;
  }
  const int n_int = t.n_nodes - t.n_tips;
  std::vector<int64_t> off((size_t)n_int * 6 + 1, 0);
  for (int u = t.n_tips; u < t.n_nodes; u++)
    for (int k = 0; k < 3; k++)
      for (int side = 0; side < 2; side++) {
        const int ks = (k + 1 + side) % 3, start = t.nb[u][ks];
        int64_t cnt = 0;
        if (!t.is_tip(start)) {
          if (max_depth < 0) cnt = 2 * (int64_t)leaves[3 * start + t.slot_of(start, u)] - 2;
          else if (max_depth == 1) cnt = 2;
          else {  // radius-limited search (getSPRMovesInDistance): count by walking
            std::vector<std::array<int, 3>> st{{start, u, 0}};
            while (!st.empty()) {
              const auto f = st.back();
              st.pop_back();
              if (t.is_tip(f[0]) || f[2] >= max_depth) continue;
              for (int j = 0; j < 3; j++)
                if (t.nb[f[0]][j] != f[1]) { cnt++; st.push_back({t.nb[f[0]][j], f[0], f[2] + 1}); }
            }
          }
        }
        off[(size_t)(u - t.n_tips) * 6 + k * 2 + side + 1] = cnt;
      }
  for (size_t i = 0; i + 1 < off.size(); i++) off[i + 1] += off[i];
  nb.moves.resize(off.back());
  const int n_thr = std::max(1, std::min(HostPool::get().size(), std::max(1, n_int / 8)));
  auto work = [&](int ti) {
    std::vector<Frame> st;
    const int u0 = t.n_tips + (int)((int64_t)n_int * ti / n_thr), u1 = t.n_tips + (int)((int64_t)n_int * (ti + 1) / n_thr);
    for (int u = u0; u < u1; u++)
      for (int k = 0; k < 3; k++) {
        const uint64_t hv = far[3 * u + k];
        for (int side = 0; side < 2; side++) {
          const int ks = (k + 1 + side) % 3;
          int64_t at = off[(size_t)(u - t.n_tips) * 6 + k * 2 + side];
          st.push_back({t.nb[u][ks], u, -1, 0, term[3 * u + ks], 0});
          while (!st.empty()) {
            const Frame f = st.back();
            st.pop_back();
            if (t.is_tip(f.x)) continue;
            if (max_depth >= 0 && f.depth >= max_depth) continue;
            for (int j = 2; j >= 0; j--) {  // (reverse so that the stack pops in neighbour order)
              const int c = t.nb[f.x][j];
              if (c == f.from) continue;
              const uint64_t joined = split_term(far[3 * f.x + j] ^ hv, all);
              Move m;
              m.u = u; m.k = k; m.x = f.x; m.kx = j; m.depth = f.depth + 1; m.parent = f.parent_move;
              m.side = side;
              m.hash = base - f.rem + f.add + joined;
              m.unique_id = -1;
              nb.moves[at] = m;
              st.push_back({c, f.x, (int)at, f.depth + 1, f.rem + term[3 * f.x + j], f.add + joined});
              at++;
            }
          }
        }
      }
  };
  HostPool::get().run(n_thr, [&](int ti, int) { work(ti); });
  nb.search_off = off;
  // Distinct topologies.  Only moves at distance 1 (NNI) repeat: each NNI neighbour arises from
  // 4 of the 8(n-3) distance-1 (prune, target) pairs, which accounts for all of
  // 4(n-3)(n-2) - 2(n-3)(2n-7) = 6(n-3) duplicates; tests check distinctness of every hash.
  // The distance-1 moves are the first two of every search: dedup those serially in enumeration
  // order (marking repeats with unique_id -2), then number everything else in parallel.
  std::unordered_map<uint64_t, int> seen;
  seen.reserve(32 * (size_t)t.n_tips);
  for (size_t s = 0; s + 1 < off.size(); s++)
    for (int64_t i = off[s]; i < off[s + 1] && nb.moves[i].depth == 1; i++)
      if (!seen.emplace(nb.moves[i].hash, (int)i).second) nb.moves[i].unique_id = -2;
  const int64_t n_moves = (int64_t)nb.moves.size();
  const int n_thr2 = (int)std::max<int64_t>(1, std::min<int64_t>(HostPool::get().size(), n_moves / 8192));
  std::vector<int64_t> kept(n_thr2 + 1, 0);
  HostPool::get().run(n_thr2, [&](int ti, int nt) {
    int64_t c = 0;
    for (int64_t i = n_moves * ti / nt; i < n_moves * (ti + 1) / nt; i++) c += nb.moves[i].unique_id != -2;
    kept[ti + 1] = c;
  });
  for (int i = 0; i < n_thr2; i++) kept[i + 1] += kept[i];
  nb.unique.resize(kept[n_thr2]);
  HostPool::get().run(n_thr2, [&](int ti, int nt) {
    int64_t at = kept[ti];
    for (int64_t i = n_moves * ti / nt; i < n_moves * (ti + 1) / nt; i++) {
      if (nb.moves[i].unique_id == -2) { nb.moves[i].unique_id = -1; continue; }
      nb.moves[i].unique_id = (int)at;
      nb.unique[at++] = (int)i;
    }
  });
}

UTree apply_move(const UTree &t, const Move &m) {
  UTree r = t;
  const int u = m.u, v = t.nb[u][m.k];
  const int a = t.nb[u][(m.k + 1) % 3], b = t.nb[u][(m.k + 2) % 3];
  const double la = t.len[u][(m.k + 1) % 3], lb = t.len[u][(m.k + 2) % 3], lv = t.len[u][m.k];
  // suppress u: a -- b with the summed length (rearrangement.py:480-490)
  const int ka = r.slot_of(a, u), kb = r.slot_of(b, u);
  r.nb[a][ka] = b; r.len[a][ka] = la + lb;
  r.nb[b][kb] = a; r.len[b][kb] = la + lb;
  // break the target edge in half and hang u (with the pruned branch) in the middle (:466-478)
  const int x = m.x, c = t.nb[x][m.kx];
  const double half = t.len[x][m.kx] / 2.0;
  const int kx = r.slot_of(x, c), kc = r.slot_of(c, x);
  r.nb[x][kx] = u; r.len[x][kx] = half;
  r.nb[c][kc] = u; r.len[c][kc] = half;
  r.nb[u] = {v, x, c};
  r.len[u] = {lv, half, half};
  for (auto &e : r.rooted_child_end) e = -1;  // names of the base tree no longer apply
  return r;
}

void desc_from_utree(const UTree &t, int32_t *desc, double *brlen) {
  const int n = t.n_tips;
  if (n == 2) {
    desc[0] = 0; desc[1] = 1;
    if (brlen) { brlen[0] = t.len[0][0]; brlen[1] = 0.0; }
    return;
  }
  // iterative post-order from tip 0's neighbour, away from tip 0
  const int w = t.nb[0][0];
  std::vector<int> new_id(t.n_nodes, -1);
  for (int x = 0; x < n; x++) new_id[x] = x;
  struct Frame { int x, from, next; };
  std::vector<Frame> st;
  st.push_back({w, 0, 0});
  int made = 0;
  while (!st.empty()) {
    Frame &f = st.back();
    if (t.is_tip(f.x)) { st.pop_back(); continue; }
    if (f.next < 3) {
      int c = t.nb[f.x][f.next++];
      if (c != f.from) st.push_back({c, f.x, 0});
      continue;
    }
    int slot = 0;
    for (int j = 0; j < 3; j++) {
      int c = t.nb[f.x][j];
      if (c == f.from) continue;
      desc[2 * made + slot] = new_id[c];
      if (brlen) brlen[2 * made + slot] = t.len[f.x][j];
      slot++;
    }
    new_id[f.x] = n + made++;
    st.pop_back();
  }
  desc[2 * made] = 0;
  desc[2 * made + 1] = new_id[w];
  if (brlen) { brlen[2 * made] = t.len[0][0]; brlen[2 * made + 1] = 0.0; }
}

// ---------------------------------------------------------------------------------------------
// edge-insertion programs
// ---------------------------------------------------------------------------------------------
namespace {

// moves needed by unique[lo..hi): the selected ones and every ancestor on their search paths
std::vector<char> needed_moves(const Neighbourhood &nb, int64_t lo, int64_t hi) {
  std::vector<char> need(nb.moves.size(), 0);
  for (int64_t i = lo; i < hi; i++) {
    int m = nb.unique[i];
    while (m >= 0 && !need[m]) {
      need[m] = 1;
      m = nb.moves[m].parent;
    }
  }
  return need;
}

}  // namespace

// Runs fn(t, n_threads) on the host cores.
template <typename F>
static void parallel_for_threads(int n_thr, F fn) {
  HostPool::get().run(n_thr, fn);
}

void build_fitch_nb_program(const Neighbourhood &nb, const int32_t *tip_rows, int n_tip_slots, int64_t lo,
                            int64_t hi, FitchNbProgram &out) {
  const UTree &t = nb.t;
  FitchProgram &prog = out.prog;
  prog.clear();
  DirOrder d = dir_order(t);
  out.dir_acc.assign((size_t)t.n_nodes * 3, -1);
  out.dir_kids.assign((size_t)t.n_nodes * 3, {-1, -1});
  std::vector<int> dir_slot((size_t)t.n_nodes * 3, -1);  // operand id of D[x->.]
  for (int x = 0; x < t.n_tips; x++)
    for (int k = 0; k < 3; k++) dir_slot[3 * x + k] = tip_rows ? tip_rows[x] : x;
  const int64_t n_dir = (int64_t)d.order.size();
  const int64_t n_moves = (int64_t)nb.moves.size();
  // which search steps are needed, and their scratch slots (enumeration order)
  const bool all = (lo == 0 && hi == (int64_t)nb.unique.size());
  std::vector<char> need;
  std::vector<int> n_slot_rank;
  int64_t n_need = n_moves;
  if (!all) {
    need = needed_moves(nb, lo, hi);
    n_slot_rank.resize(n_moves);
    n_need = 0;
    for (int64_t i = 0; i < n_moves; i++) { n_slot_rank[i] = (int)n_need; n_need += need[i]; }
  }
  auto slot_of_move = [&](int64_t i) { return n_tip_slots + (int)(n_dir + (all ? i : n_slot_rank[i])); };
  // level layout: directional ops keep their dependency level, search steps sit at Ld + depth
  int max_depth = 0;
  const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>(HostPool::get().size(), n_moves / 4096));
  std::vector<std::vector<int64_t>> cnt(n_thr);
  parallel_for_threads(n_thr, [&](int ti, int nt) {
    std::vector<int64_t> &c = cnt[ti];
    c.assign(64, 0);
    for (int64_t i = n_moves * ti / nt; i < n_moves * (ti + 1) / nt; i++) {
      if (!all && !need[i]) continue;
      const int dep = nb.moves[i].depth;
      if ((size_t)dep >= c.size()) c.resize(dep + 64, 0);
      c[dep]++;
    }
  });
  for (auto &c : cnt) {
    int top = (int)c.size() - 1;
    while (top > 0 && c[top] == 0) top--;
    max_depth = std::max(max_depth, top);
  }
  const int n_levels = d.max_level + max_depth + 1;
  prog.level_off.assign(n_levels + 1, 0);
  for (int e : d.order) prog.level_off[d.level[e] + 1]++;
  for (auto &c : cnt)
    for (int dep = 1; dep <= max_depth && dep < (int)c.size(); dep++) prog.level_off[d.max_level + dep + 1] += c[dep];
  for (int l = 0; l < n_levels; l++) prog.level_off[l + 1] += prog.level_off[l];
  // per-thread write cursors per depth
  std::vector<std::vector<int64_t>> cur(n_thr, std::vector<int64_t>(max_depth + 1, 0));
  for (int dep = 1; dep <= max_depth; dep++) {
    int64_t at = prog.level_off[d.max_level + dep];
    for (int ti = 0; ti < n_thr; ti++) {
      cur[ti][dep] = at;
      at += dep < (int)cnt[ti].size() ? cnt[ti][dep] : 0;
    }
  }
  const int64_t n_ops = n_dir + n_need;
  prog.ops.resize(n_ops);
  prog.children.resize(2 * n_dir + 4 * n_need);
  FitchOp *ops = prog.ops.data();
  int *kids = prog.children.data();
  // directional ops (serial: 3(n-2) of them)
  {
    std::vector<int64_t> dcur(prog.level_off.begin(), prog.level_off.begin() + d.max_level + 1);
    int64_t slots = 0, accs = 0;
    for (int e : d.order) {
      const int x = e / 3, k = e % 3;
      const int64_t at = dcur[d.level[e]]++;
      FitchOp o;
      o.dst = n_tip_slots + (int)slots++;
      o.nchild = 2;
      o.tree = (int)accs;
      o.child_off = (int)(2 * at);
      out.dir_acc[e] = (int)accs++;
      int s = 0;
      for (int j = 0; j < 3; j++)
        if (j != k) {
          const int c = t.nb[x][j];
          const int ce = 3 * c + t.slot_of(c, x);
          kids[2 * at + s] = dir_slot[ce];
          out.dir_kids[e][s] = t.is_tip(c) ? -1 : ce;
          s++;
        }
      dir_slot[e] = o.dst;
      ops[at] = o;
    }
  }
  const int dump_acc = (int)n_dir;  // own misses of search-step partials (never read)
  out.first_move_acc = n_dir + 1;
  // search steps, fused with the join of the neighbour they realise when that one is selected
  parallel_for_threads(n_thr, [&](int ti, int nt) {
    std::vector<int64_t> &c = cur[ti];
    for (int64_t i = n_moves * ti / nt; i < n_moves * (ti + 1) / nt; i++) {
      if (!all && !need[i]) continue;
      const Move &m = nb.moves[i];
      const int x = m.x, cnode = t.nb[x][m.kx];
      int from, xin;
      if (m.parent < 0) {
        const int other = t.nb[m.u][(m.k + 1 + (1 - m.side)) % 3];
        from = m.u;
        xin = dir_slot[3 * other + t.slot_of(other, m.u)];  // D[other->u] enters x over the merged edge
      } else {
        from = nb.moves[m.parent].x;
        xin = slot_of_move(m.parent);
      }
      int s = -1;
      for (int j = 0; j < 3; j++)
        if (t.nb[x][j] != from && t.nb[x][j] != cnode) s = t.nb[x][j];
      const int64_t at = c[m.depth]++;
      const int64_t koff = 2 * n_dir + 4 * (at - n_dir);
      FitchOp o;
      // a step onto a pendant edge has no further steps below it: its partial is only consumed
      // by its own fused join, so it is never stored
      o.dst = t.is_tip(cnode) ? -1 : slot_of_move(i);
      o.child_off = (int)koff;
      kids[koff] = xin;
      kids[koff + 1] = dir_slot[3 * s + t.slot_of(s, x)];
      const int uid = m.unique_id;
      if (uid >= lo && uid < hi) {
        const int v = t.nb[m.u][m.k];
        o.nchild = FITCH_FUSED;
        o.tree = (int)(out.first_move_acc + (uid - lo));
        kids[koff + 2] = dir_slot[3 * cnode + t.slot_of(cnode, x)];
        kids[koff + 3] = dir_slot[3 * v + t.slot_of(v, m.u)];
      } else {
        o.nchild = 2;
        o.tree = dump_acc;
        kids[koff + 2] = kids[koff + 3] = 0;
        if (o.dst < 0) o.nchild = 1;  // duplicate topology on a pendant edge: nothing to do (unary, no store)
      }
      ops[at] = o;
    }
  });
  prog.n_scratch_slots = n_dir + n_need;
  prog.n_costs = out.first_move_acc + (hi - lo);
}

// Search-walk program (fitch_walk.cuh fitch_search_walk_kernel): the directional ops of the base
// tree as level ops, then one group of visits per pruned subtree.  Search steps come in adjacent
// pairs (2q, 2q+1) -- the two insertion edges at the node a step reaches = one visit.  The visits
// of a search are listed depth-first; where both children continue, the smaller subtree goes
// first and the other partial waits on the stack (depth <= log2 n).  Searches are independent, so
// the host cores lay them out in parallel at offsets known from the subtree sizes.
void build_fitch_nb_walk_program(const Neighbourhood &nb, const int32_t *tip_rows, int n_tip_slots, int64_t lo,
                                 int64_t hi, FitchNbProgram &out) {
  const UTree &t = nb.t;
  FitchProgram &prog = out.prog;
  prog.clear();
  DirOrder d = dir_order(t);
  out.dir_acc.assign((size_t)t.n_nodes * 3, -1);
  out.dir_kids.assign((size_t)t.n_nodes * 3, {-1, -1});
  std::vector<int> dir_slot((size_t)t.n_nodes * 3, -1);
  for (int x = 0; x < t.n_tips; x++)
    for (int k = 0; k < 3; k++) dir_slot[3 * x + k] = tip_rows ? tip_rows[x] : x;
  const int64_t n_dir = (int64_t)d.order.size();
  prog.level_off.assign(d.max_level + 2, 0);
  for (int e : d.order) prog.level_off[d.level[e] + 1]++;
  for (int l = 0; l <= d.max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
  prog.ops.resize(n_dir);
  prog.children.resize(2 * n_dir);
  {
    std::vector<int64_t> dcur(prog.level_off.begin(), prog.level_off.end() - 1);
    int64_t slots = 0, accs = 0;
    for (int e : d.order) {
      const int x = e / 3, k = e % 3;
      const int64_t at = dcur[d.level[e]]++;
      FitchOp o;
      o.dst = n_tip_slots + (int)slots++;
      o.nchild = 2;
      o.tree = (int)accs;
      o.child_off = (int)(2 * at);
      out.dir_acc[e] = (int)accs++;
      int s = 0;
      for (int j = 0; j < 3; j++)
        if (j != k) {
          const int c = t.nb[x][j];
          const int ce = 3 * c + t.slot_of(c, x);
          prog.children[2 * at + s] = dir_slot[ce];
          out.dir_kids[e][s] = t.is_tip(c) ? -1 : ce;
          s++;
        }
      dir_slot[e] = o.dst;
      prog.ops[at] = o;
    }
  }
  out.first_move_acc = n_dir + 1;
  prog.n_scratch_slots = n_dir;
  prog.n_costs = out.first_move_acc + (hi - lo);
  const int64_t n_moves = (int64_t)nb.moves.size(), n_pairs = n_moves / 2;
  if (n_pairs == 0) return;
  const bool all = (lo == 0 && hi == (int64_t)nb.unique.size());
  std::vector<char> need;
  if (!all) need = needed_moves(nb, lo, hi);
  // child[m]: the visit that continues below move m; sub[q]: needed visits in the subtree of visit q.
  // Parent links never leave a search, so the searches are swept in parallel (last visit first).
  static thread_local std::vector<int> child, sub;
  child.resize(n_moves);
  sub.resize(n_pairs);
  int *childw = child.data(), *subw = sub.data();
  const int64_t n_search = (int64_t)nb.search_off.size() - 1;
  const int n_thr0 = (int)std::max<int64_t>(1, std::min<int64_t>(HostPool::get().size(), n_moves / 8192));
  HostPool::get().run(n_thr0, [&](int ti, int nt) {
    for (int64_t sidx = n_search * ti / nt; sidx < n_search * (ti + 1) / nt; sidx++) {
      const int64_t m0 = nb.search_off[sidx], m1 = nb.search_off[sidx + 1];
      for (int64_t m = m0; m < m1; m++) childw[m] = -1;
      for (int64_t q = m0 / 2; q < m1 / 2; q++) subw[q] = 0;
      for (int64_t q = m1 / 2; q-- > m0 / 2;) {
        if (!all && !need[2 * q] && !need[2 * q + 1]) continue;
        subw[q] += 1;
        const int pm = nb.moves[2 * q].parent;
        if (pm >= 0) { childw[pm] = (int)q; subw[pm / 2] += subw[q]; }
      }
    }
  });
  // roots of the needed searches (the first visit of every search), grouped by pruned subtree (u, k)
  struct Root { int q, group; int64_t off; };
  std::vector<Root> roots;
  std::vector<FitchWalkGroup> groups;  // (HostArr does not keep its contents when it grows)
  int cur_u = -1, cur_k = -1;
  int64_t n_visits = 0;
  for (int64_t sidx = 0; sidx < n_search; sidx++) {
    if (nb.search_off[sidx + 1] == nb.search_off[sidx]) continue;
    const int64_t q = nb.search_off[sidx] / 2;
    const Move &r = nb.moves[2 * q];
    if (sub[q] == 0) continue;
    if (r.u != cur_u || r.k != cur_k) {
      const int v = t.nb[r.u][r.k];
      FitchWalkGroup g;
      g.tv = dir_slot[3 * v + t.slot_of(v, r.u)];
      g.first = (int)n_visits;
      g.count = 0;
      g.pad = 0;
      groups.push_back(g);
      cur_u = r.u; cur_k = r.k;
    }
    roots.push_back({(int)q, (int)groups.size() - 1, n_visits});
    groups.back().count += sub[q];
    n_visits += sub[q];
  }
  // longest searches first (the kernel's blocks are scheduled in group order)
  std::stable_sort(groups.begin(), groups.end(),
                   [](const FitchWalkGroup &a, const FitchWalkGroup &b) { return a.count > b.count; });
  prog.walk_groups.resize(groups.size());
  std::copy(groups.begin(), groups.end(), prog.walk_groups.data());
  prog.walk_visits.resize(n_visits);
  FitchWalkVisit *vis = prog.walk_visits.data();
  const int *childp = child.data(), *subp = sub.data();  // (thread_local: the pool's workers must not name them)
  const int64_t n_roots = (int64_t)roots.size();
  const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>(HostPool::get().size(), n_visits / 2048));
  HostPool::get().run(n_thr, [&](int ti, int nt) {
    struct Frame { int q, x_in; };
    std::vector<Frame> st;
    // contiguous blocks of roots with about equal numbers of visits
    const int64_t v0 = n_visits * ti / nt, v1 = n_visits * (ti + 1) / nt;
    for (int64_t ri = 0; ri < n_roots; ri++) {
      if (roots[ri].off < v0 || roots[ri].off >= v1) continue;
      const Move &r = nb.moves[2 * roots[ri].q];
      const int ko = (r.k + 1 + (1 - r.side)) % 3, other = t.nb[r.u][ko];
      int64_t at = roots[ri].off;
      int depth = 0;
      st.push_back({roots[ri].q, dir_slot[3 * other + t.slot_of(other, r.u)]});
      while (!st.empty()) {
        const Frame f = st.back();
        st.pop_back();
        if (f.x_in <= -2) depth--;
        const Move &m1 = nb.moves[2 * f.q], &m2 = nb.moves[2 * f.q + 1];
        const int x = m1.x, c1 = t.nb[x][m1.kx], c2 = t.nb[x][m2.kx];
        FitchWalkVisit w;
        w.x_in = f.x_in;
        w.d1 = dir_slot[3 * c1 + t.slot_of(c1, x)];
        w.d2 = dir_slot[3 * c2 + t.slot_of(c2, x)];
        w.tree1 = (m1.unique_id >= lo && m1.unique_id < hi) ? (int)(out.first_move_acc + (m1.unique_id - lo)) : -1;
        w.tree2 = (m2.unique_id >= lo && m2.unique_id < hi) ? (int)(out.first_move_acc + (m2.unique_id - lo)) : -1;
        w.keep = 0; w.push = -1; w.pad = 0;
        const int ch1 = (all || need[2 * f.q]) ? childp[2 * f.q] : -1;
        const int ch2 = (all || need[2 * f.q + 1]) ? childp[2 * f.q + 1] : -1;
        if (ch1 >= 0 && ch2 >= 0) {
          const bool first1 = subp[ch1] <= subp[ch2];
          w.keep = first1 ? 1 : 2;
          w.push = depth;
          st.push_back({first1 ? ch2 : ch1, -2 - depth});
          st.push_back({first1 ? ch1 : ch2, -1});
          depth++;
        } else if (ch1 >= 0) {
          w.keep = 1;
          st.push_back({ch1, -1});
        } else if (ch2 >= 0) {
          w.keep = 2;
          st.push_back({ch2, -1});
        }
        vis[at++] = w;
      }
    }
  });
}

// Directional ops of the base tree as a level program (shared by the walk planners).
static void emit_dir_program(const UTree &t, const DirOrder &d, const int32_t *tip_rows, int n_tip_slots,
                             FitchNbProgram &out, std::vector<int> &dir_slot) {
  FitchProgram &prog = out.prog;
  prog.clear();
  out.dir_acc.assign((size_t)t.n_nodes * 3, -1);
  out.dir_kids.assign((size_t)t.n_nodes * 3, {-1, -1});
  dir_slot.assign((size_t)t.n_nodes * 3, -1);
  for (int x = 0; x < t.n_tips; x++)
    for (int k = 0; k < 3; k++) dir_slot[3 * x + k] = tip_rows ? tip_rows[x] : x;
  const int64_t n_dir = (int64_t)d.order.size();
  prog.level_off.assign(d.max_level + 2, 0);
  for (int e : d.order) prog.level_off[d.level[e] + 1]++;
  for (int l = 0; l <= d.max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
  prog.ops.resize(n_dir);
  prog.children.resize(2 * n_dir);
  std::vector<int64_t> dcur(prog.level_off.begin(), prog.level_off.end() - 1);
  int64_t slots = 0, accs = 0;
  for (int e : d.order) {
    const int x = e / 3, k = e % 3;
    const int64_t at = dcur[d.level[e]]++;
    FitchOp o;
    o.dst = n_tip_slots + (int)slots++;
    o.nchild = 2;
    o.tree = (int)accs;
    o.child_off = (int)(2 * at);
    out.dir_acc[e] = (int)accs++;
    int s = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        const int c = t.nb[x][j];
        const int ce = 3 * c + t.slot_of(c, x);
        prog.children[2 * at + s] = dir_slot[ce];
        out.dir_kids[e][s] = t.is_tip(c) ? -1 : ce;
        s++;
      }
    dir_slot[e] = o.dst;
    prog.ops[at] = o;
  }
  out.first_move_acc = n_dir + 1;
  prog.n_scratch_slots = n_dir;
}

// What the device planners share: one SprEdgeRec per directed edge, one SprSearchRec per search
// (6 per internal node, enumeration order), one group per pruned subtree with a non-empty search
// (pad = its directed edge 3*u+k; longest first), and the de-duplication of the distance-1 moves.
namespace {
struct SprSearchLayout {
  std::vector<SprEdgeRec> edges;
  std::vector<SprSearchRec> searches;
  std::vector<FitchWalkGroup> groups;
  std::vector<int> search_edge;
  std::vector<int64_t> search_uniq;
  int64_t n_visits = 0, n_unique = 0, dup_internal = 0, n_started = 0, unscored = 0;
};

// [lo, hi): the unique ids a call scores.  Searches that hold none of them get no visits (a shard
// of the neighbourhood plans and walks only its own searches; the two at its borders are walked
// whole and score only their ids inside the range).
void layout_spr_searches(const UTree &t, const DirOrder &d, const std::vector<int> &dir_slot, SprSearchLayout &out,
                         int64_t lo = 0, int64_t hi = INT64_MAX) {
  out.edges.clear(); out.searches.clear(); out.groups.clear(); out.search_edge.clear();
  out.search_uniq.assign(1, 0);
  out.n_visits = out.n_unique = out.dup_internal = out.n_started = out.unscored = 0;
  if (t.n_tips < 4) return;
  // leaves on x's side of edge x -> nb[x][k]
  std::vector<int> leaves((size_t)t.n_nodes * 3, 0);
  for (int x = 0; x < t.n_tips; x++) leaves[3 * x] = leaves[3 * x + 1] = leaves[3 * x + 2] = 1;
  for (int e : d.order) {
    const int x = e / 3, k = e % 3;
    int c = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) c += leaves[3 * t.nb[x][j] + t.slot_of(t.nb[x][j], x)];
    leaves[e] = c;
  }
  out.edges.resize((size_t)t.n_nodes * 3);
  for (int x = 0; x < t.n_nodes; x++)
    for (int j = 0; j < 3; j++) {
      SprEdgeRec r;
      r.c = t.nb[x][j];
      r.far_slot = r.far_sub = 0;
      r.child_end = -1;
      if (r.c >= 0) {
        const int back = 3 * r.c + t.slot_of(r.c, x);
        r.far_slot = dir_slot[back];
        r.far_sub = leaves[back] - 1;
        r.child_end = t.rooted_child_end[3 * x + j];
      }
      out.edges[3 * (size_t)x + j] = r;
    }
  // hashes of the distance-1 moves (as enumerate_spr computes them) for the de-duplication
  uint64_t all;
  std::vector<uint64_t> h = side_hashes(t, d, &all);
  uint64_t base = 0;
  for (int x = 0; x < t.n_nodes; x++)
    for (int k = 0; k < 3; k++)
      if (t.nb[x][k] > x) base += split_term(h[3 * x + k], all);
  auto far = [&](int x, int k) { const int y = t.nb[x][k]; return h[3 * y + t.slot_of(y, x)]; };
  // open-addressing set of the 8(n-3) distance-1 hashes (0 = empty; a hash of 0 is stored as 1)
  size_t cap = 64;
  while (cap < 32 * (size_t)t.n_tips) cap <<= 1;
  static thread_local std::vector<uint64_t> seen;
  seen.assign(cap, 0);
  auto first_time = [&](uint64_t key) {
    if (key == 0) key = 1;
    for (size_t i = splitmix(key) & (cap - 1);; i = (i + 1) & (cap - 1)) {
      if (seen[i] == key) return false;
      if (seen[i] == 0) { seen[i] = key; return true; }
    }
  };
  const int n_int = t.n_nodes - t.n_tips;
  out.searches.reserve((size_t)n_int * 6);
  int64_t move_off = 0, dups = 0, vis_off = 0;
  for (int u = t.n_tips; u < t.n_nodes; u++)
    for (int k = 0; k < 3; k++) {
      int cur_group = -1;
      const uint64_t hv = far(u, k);
      const int v = t.nb[u][k], pe = t.rooted_child_end[3 * u + k];
      for (int side = 0; side < 2; side++) {
        const int ks = (k + 1 + side) % 3, ko = (k + 1 + (1 - side)) % 3;
        const int start = t.nb[u][ks], other = t.nb[u][ko];
        SprSearchRec r;
        r.start = start;
        r.from = u;
        r.x_in = dir_slot[3 * other + t.slot_of(other, u)];
        r.visit_off = (int)vis_off;
        r.uniq_base = (int)(move_off - dups);
        r.move_cnt = t.is_tip(start) ? 0 : 2 * leaves[3 * start + t.slot_of(start, u)] - 2;
        r.dup_flag = pe == v ? 0 : 4;
        r.pe = pe;
        const int n_moves = r.move_cnt;
        if (r.move_cnt > 0) {
          const uint64_t rem = split_term(far(u, ks), all);
          int e = 0, dup_int = 0;
          for (int j = 2; j >= 0; j--) {  // enumerate_spr writes the higher slot's move first
            if (t.nb[start][j] == u) continue;
            const uint64_t joined = split_term(far(start, j) ^ hv, all);
            if (!first_time(base - rem + joined)) {
              r.dup_flag |= 1 << e;
              dups++;
              if (!t.is_tip(t.nb[start][j])) dup_int++;
            }
            e++;
          }
          const int64_t u0 = r.uniq_base, u1 = u0 + n_moves - ((r.dup_flag & 1) + ((r.dup_flag >> 1) & 1));
          // a one-visit search whose two moves both repeat earlier topologies has nothing to do
          // (the host planner drops it as well); neither has a search outside the scored range
          if ((r.move_cnt == 2 && (r.dup_flag & 3) == 3) || u1 <= lo || u0 >= hi) r.move_cnt = 0;
          else {
            out.dup_internal += dup_int;
            out.unscored += (int64_t)n_moves - (std::min(u1, hi) - std::max(u0, lo));  // repeats and ids outside the range
          }
        }
        if (r.move_cnt > 0) {
          if (cur_group < 0) {
            FitchWalkGroup g;
            g.tv = dir_slot[3 * v + t.slot_of(v, u)];
            g.first = r.visit_off;
            g.count = 0;
            g.pad = 3 * u + k;
            out.groups.push_back(g);
            cur_group = (int)out.groups.size() - 1;
          }
          out.groups[cur_group].count += r.move_cnt / 2;
          out.n_started++;
          vis_off += r.move_cnt / 2;
        }
        out.searches.push_back(r);
        out.search_edge.push_back(3 * u + k);
        move_off += n_moves;
        out.search_uniq.push_back(move_off - dups);
      }
    }
  out.n_visits = vis_off;
  out.n_unique = move_off - dups;
  // longest searches first (the kernels take groups in this order)
  std::stable_sort(out.groups.begin(), out.groups.end(),
                   [](const FitchWalkGroup &a, const FitchWalkGroup &b) { return a.count > b.count; });
}
}  // namespace

// A few moves of the full SPR neighbourhood by unique id, without enumerating it: the search
// layout (O(n)) says which search holds the id, and only that search is expanded -- in
// enumerate_spr's order, with its hashes.  What a hill climb needs to materialise the winner.
void spr_moves_by_unique_id(const UTree &t, int64_t n_sel, const int64_t *sel, Move *out) {
  DirOrder d = dir_order(t);
  std::vector<int> dir_slot((size_t)t.n_nodes * 3, 0);
  static thread_local SprSearchLayout lay;
  layout_spr_searches(t, d, dir_slot, lay);
  uint64_t all;
  std::vector<uint64_t> h = side_hashes(t, d, &all);
  uint64_t base = 0;
  for (int x = 0; x < t.n_nodes; x++)
    for (int k = 0; k < 3; k++)
      if (t.nb[x][k] > x) base += split_term(h[3 * x + k], all);
  auto far = [&](int x, int k) { const int y = t.nb[x][k]; return h[3 * y + t.slot_of(y, x)]; };
  struct Frame { int x, from, parent_move, depth; uint64_t rem, add; };
  std::vector<Frame> st;
  std::vector<Move> moves;
  for (int64_t q = 0; q < n_sel; q++) {
    const int64_t uid = sel[q];
    if (uid < 0 || uid >= lay.n_unique) PLG_FAIL(PLG_ERR_ARG, "move index out of range");
    const size_t s = std::upper_bound(lay.search_uniq.begin(), lay.search_uniq.end(), uid) - lay.search_uniq.begin() - 1;
    const int e = lay.search_edge[s], u = e / 3, k = e % 3, side = (int)(s & 1), ks = (k + 1 + side) % 3;
    const uint64_t hv = far(u, k);
    moves.clear();
    st.push_back({t.nb[u][ks], u, -1, 0, split_term(far(u, ks), all), 0});
    while (!st.empty()) {
      const Frame f = st.back();
      st.pop_back();
      if (t.is_tip(f.x)) continue;
      for (int j = 2; j >= 0; j--) {
        const int c = t.nb[f.x][j];
        if (c == f.from) continue;
        const uint64_t joined = split_term(far(f.x, j) ^ hv, all);
        Move m;
        m.u = u; m.k = k; m.x = f.x; m.kx = j; m.depth = f.depth + 1; m.parent = f.parent_move; m.side = side;
        m.hash = base - f.rem + f.add + joined;
        m.unique_id = -1;
        st.push_back({c, f.x, (int)moves.size(), f.depth + 1, f.rem + split_term(far(f.x, j), all), f.add + joined});
        moves.push_back(m);
      }
    }
    // the id-th kept move of the search: only its first two moves can be repeats
    const int dup = lay.searches[s].dup_flag & 3;
    int64_t want = uid - lay.search_uniq[s], at = -1;
    for (size_t i = 0; i < moves.size(); i++) {
      if (i < 2 && (dup >> i & 1)) continue;
      if (want-- == 0) { at = (int64_t)i; break; }
    }
    if (at < 0) PLG_FAIL(PLG_ERR_ARG, "internal: move %lld not found in its search", (long long)uid);
    out[q] = moves[at];
    out[q].unique_id = (int)uid;
  }
}

void prepare_spr_device_plan(const UTree &t, const int32_t *tip_rows, int n_tip_slots, SprDevicePlan &out,
                             int64_t lo, int64_t hi) {
  DirOrder d = dir_order(t);
  std::vector<int> dir_slot;
  emit_dir_program(t, d, tip_rows, n_tip_slots, out.base, dir_slot);
  out.order = d.order;
  FitchProgram &prog = out.base.prog;
  prog.spr_first_acc = (int)out.base.first_move_acc;
  static thread_local SprSearchLayout lay;
  layout_spr_searches(t, d, dir_slot, lay, lo, hi);
  out.search_edge = lay.search_edge;
  out.search_uniq = lay.search_uniq;
  out.n_visits = lay.n_visits;
  out.n_unique = lay.n_unique;
  out.lo = std::max<int64_t>(0, lo);
  out.hi = std::min<int64_t>(hi, lay.n_unique);
  prog.spr_lo = out.lo;
  prog.spr_hi = out.hi;
  prog.spr_edges.resize(lay.edges.size());
  std::copy(lay.edges.begin(), lay.edges.end(), prog.spr_edges.data());
  prog.spr_searches.resize(lay.searches.size());
  std::copy(lay.searches.begin(), lay.searches.end(), prog.spr_searches.data());
  prog.walk_groups.resize(lay.groups.size());
  std::copy(lay.groups.begin(), lay.groups.end(), prog.walk_groups.data());
  prog.spr_n_visits = out.n_visits;
  prog.spr_n_unique = out.n_unique;
  prog.n_costs = out.base.first_move_acc + (out.hi - out.lo);
  // byte accounting of the walk launch (run_fitch_program): a scored neighbour = fold + 3-way join
  // (6 vectors); a repeated distance-1 move still folds when the search continues below it (exact for
  // a whole neighbourhood; in a shard the unscored moves of its two border searches are taken to
  // continue half of the time)
  const bool whole = out.lo == 0 && out.hi == lay.n_unique;
  prog.spr_vecs = 6.0 * (double)(out.hi - out.lo) +
                  (whole ? 3.0 * (double)lay.dup_internal : 1.5 * (double)lay.unscored);
  prog.spr_req_vecs = (double)lay.groups.size() + 2.0 * (double)out.n_visits + (double)lay.n_started;
}

// Likelihood twin of prepare_spr_device_plan: directional CLV updates of the base tree as level
// ops, P-matrix ids of every branch (whole, halved) and of the three merged branches around every
// internal node, and the records spr_lik_plan_kernel (lik.cu) expands into LikWalkVisits.
void build_lik_spr_device_plan(const UTree &t, const int32_t *tip_rows, int n_rows, LikProgram &prog, int64_t lo,
                               int64_t hi) {
  prog.clear();
  DirOrder d = dir_order(t);
  std::vector<int> dir_slot((size_t)t.n_nodes * 3, -1);
  for (int x = 0; x < t.n_tips; x++)
    for (int k = 0; k < 3; k++) dir_slot[3 * x + k] = tip_rows ? tip_rows[x] : x;
  prog.level_off.assign(d.max_level + 2, 0);
  for (int e : d.order) prog.level_off[d.level[e] + 1]++;
  for (int l = 0; l <= d.max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
  std::vector<int64_t> cur(prog.level_off.begin(), prog.level_off.end() - 1);
  prog.ops.resize(d.order.size());
  int64_t slots = 0;
  for (int e : d.order) {
    const int x = e / 3, k = e % 3;
    LikOp o;
    o.dst = n_rows + (int)slots++;
    int s = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        const int c = t.nb[x][j];
        const int id = dir_slot[3 * c + t.slot_of(c, x)];
        const int pm = prog.pm(t.len[x][j]);
        if (s == 0) { o.a = id; o.pa = pm; } else { o.b = id; o.pb = pm; }
        s++;
      }
    dir_slot[e] = o.dst;
    prog.ops[cur[d.level[e]]++] = o;
  }
  prog.n_slots = slots;
  static thread_local SprSearchLayout lay;
  layout_spr_searches(t, d, dir_slot, lay, lo, hi);
  prog.spr_lo = std::max<int64_t>(0, lo);
  prog.spr_hi = std::min<int64_t>(hi, lay.n_unique);
  prog.spr_edges.resize(lay.edges.size());
  for (size_t e = 0; e < lay.edges.size(); e++) {
    const SprEdgeRec &r = lay.edges[e];
    LikSprEdge &o = prog.spr_edges[e];
    o.c = r.c; o.far_slot = r.far_slot; o.far_sub = r.far_sub; o.child_end = r.child_end;
    o.p_t = o.p_h = 0; o.pad0 = o.pad1 = 0;
    if (r.c >= 0) {
      const double len = t.len[e / 3][e % 3];
      o.p_t = prog.pm(len);
      o.p_h = prog.pm(len / 2.0);  // target branch broken in half (rearrangement.py:466-470)
    }
  }
  prog.spr_searches.resize(lay.searches.size());
  for (size_t s = 0; s < lay.searches.size(); s++) {
    const SprSearchRec &r = lay.searches[s];
    LikSprSearch &o = prog.spr_searches[s];
    o.start = r.start; o.from = r.from; o.x_in = r.x_in; o.visit_off = r.visit_off;
    o.uniq_base = r.uniq_base; o.move_cnt = r.move_cnt; o.dup_flag = r.dup_flag; o.pe = r.pe;
    const int e = lay.search_edge[s], u = e / 3, k = e % 3, side = (int)(s & 1);
    const int ks = (k + 1 + side) % 3, ko = (k + 1 + (1 - side)) % 3, v = t.nb[u][k];
    o.p_in = prog.pm(t.len[u][ko] + t.len[u][ks]);  // merged edge other -- start (rearrangement.py:480-490)
    o.tv = dir_slot[3 * v + t.slot_of(v, u)];
    o.pad0 = o.pad1 = 0;
  }
  prog.walk_groups.resize(lay.groups.size());
  for (size_t g = 0; g < lay.groups.size(); g++) {
    const FitchWalkGroup &f = lay.groups[g];
    LikWalkGroup &o = prog.walk_groups[g];
    o.tv = f.tv; o.p_v = prog.pm(t.len[f.pad / 3][f.pad % 3]); o.first = f.first; o.count = f.count;
  }
  prog.spr_n_visits = lay.n_visits;
  prog.spr_n_unique = lay.n_unique;
  prog.n_trees = prog.spr_hi - prog.spr_lo;
  int lv = 1;
  while ((1 << lv) < t.n_tips) lv++;
  prog.walk_levels = lv + 1;  // the walk parks the larger side: depth <= log2(visits of a search) + 1
}

void finish_fitch_spr_plan(const UTree &t, const SprDevicePlan &p, const int32_t *acc, int32_t *out_costs) {
  std::vector<uint32_t> total((size_t)t.n_nodes * 3, 0);
  for (int e : p.order) {
    uint32_t s = (uint32_t)acc[p.base.dir_acc[e]];
    for (int j = 0; j < 2; j++)
      if (p.base.dir_kids[e][j] >= 0) s += total[p.base.dir_kids[e][j]];
    total[e] = s;
  }
  const int64_t n_search = (int64_t)p.search_edge.size();
  const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>(HostPool::get().size(), p.n_unique / 32768));
  HostPool::get().run(n_thr, [&](int ti, int nt) {
    for (int64_t s = n_search * ti / nt; s < n_search * (ti + 1) / nt; s++) {
      const int e = p.search_edge[s], u = e / 3, k = e % 3, v = t.nb[u][k];
      // length(T') = length(pruned subtree) + length(rest) + [their state sets miss at the new edge]
      const uint32_t parts = total[e] + (t.is_tip(v) ? 0u : total[3 * v + t.slot_of(v, u)]);
      for (int64_t i = std::max(p.search_uniq[s], p.lo); i < std::min(p.search_uniq[s + 1], p.hi); i++)
        out_costs[i] = (int32_t)(parts + (uint32_t)acc[p.base.first_move_acc + (i - p.lo)]);
    }
  });
}

void finish_fitch_nb(const Neighbourhood &nb, const FitchNbProgram &p, const int32_t *acc, int64_t lo, int64_t hi,
                     int32_t *out_costs) {
  const UTree &t = nb.t;
  // total Fitch length of every directional partial: own misses + the two partials below it
  DirOrder d = dir_order(t);
  std::vector<uint32_t> total((size_t)t.n_nodes * 3, 0);
  for (int e : d.order) {
    uint32_t s = (uint32_t)acc[p.dir_acc[e]];
    for (int j = 0; j < 2; j++)
      if (p.dir_kids[e][j] >= 0) s += total[p.dir_kids[e][j]];
    total[e] = s;
  }
  // length(T') = length(pruned subtree) + length(rest) + [their state sets miss at the new edge];
  // the first two only depend on the pruned branch
  std::vector<uint32_t> parts((size_t)t.n_nodes * 3, 0);
  for (int u = t.n_tips; u < t.n_nodes; u++)
    for (int k = 0; k < 3; k++) {
      const int v = t.nb[u][k];
      parts[3 * u + k] = total[3 * u + k] + (t.is_tip(v) ? 0u : total[3 * v + t.slot_of(v, u)]);
    }
  const int n_thr = (int)std::max<int64_t>(1, std::min<int64_t>(HostPool::get().size(), (hi - lo) / 16384));
  parallel_for_threads(n_thr, [&](int ti, int nt) {
    for (int64_t i = lo + (hi - lo) * ti / nt; i < lo + (hi - lo) * (ti + 1) / nt; i++) {
      const Move &m = nb.moves[nb.unique[i]];
      out_costs[i - lo] = (int32_t)(parts[3 * m.u + m.k] + (uint32_t)acc[p.first_move_acc + (i - lo)]);
    }
  });
}

void build_lik_nb_program(const Neighbourhood &nb, const int32_t *tip_rows, int n_rows, int64_t lo, int64_t hi,
                          LikProgram &prog, int mode, bool hoist_eval) {
  const UTree &t = nb.t;
  prog.clear();
  DirOrder d = dir_order(t);
  std::vector<int> dir_slot((size_t)t.n_nodes * 3, -1);
  for (int x = 0; x < t.n_tips; x++)
    for (int k = 0; k < 3; k++) dir_slot[3 * x + k] = tip_rows ? tip_rows[x] : x;
  struct Tmp { LikOp op; int level; };
  std::vector<Tmp> tmp;
  int64_t slots = 0;
  for (int e : d.order) {
    const int x = e / 3, k = e % 3;
    Tmp o;
    o.op.dst = n_rows + (int)slots++;
    int s = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        const int c = t.nb[x][j];
        const int id = dir_slot[3 * c + t.slot_of(c, x)];
        const int pm = prog.pm(t.len[x][j]);
        if (s == 0) { o.op.a = id; o.op.pa = pm; } else { o.op.b = id; o.op.pb = pm; }
        s++;
      }
    o.level = d.level[e];
    dir_slot[e] = o.op.dst;
    tmp.push_back(o);
  }
  if (mode == LIK_NB_WALK) {
    // Depth-first walks (lik_walkm.cuh).  Search steps come in adjacent pairs (2q, 2q+1): the two
    // insertion edges at the node a step reaches = one visit.  Visits of one pruned subtree are
    // listed in the order a warp runs them: the partial toward the child with the smaller
    // remaining search continues in registers, the other one is parked on a stack (depth <= log2 n).
    std::vector<char> need = needed_moves(nb, lo, hi);
    const size_t n_moves = nb.moves.size(), n_pairs = n_moves / 2;
    std::vector<int> child(n_moves, -1);   // move -> the visit that continues below it
    std::vector<int> sub(n_pairs, 0);      // needed visits in the subtree of a visit
    for (size_t q = n_pairs; q-- > 0;) {
      if (!need[2 * q] && !need[2 * q + 1]) continue;
      sub[q] += 1;
      const int pm = nb.moves[2 * q].parent;
      if (pm >= 0) { child[pm] = (int)q; sub[pm / 2] += sub[q]; }
    }
    struct Frame { int q, x_in; };
    std::vector<Frame> st;
    int cur_u = -1, cur_k = -1, depth = 0;
    for (size_t q = 0; q < n_pairs; q++) {
      const Move &r = nb.moves[2 * q];
      if (r.parent >= 0 || sub[q] == 0) continue;  // roots of needed searches only
      if (r.u != cur_u || r.k != cur_k) {
        const int v = t.nb[r.u][r.k];
        LikWalkGroup g;
        g.tv = dir_slot[3 * v + t.slot_of(v, r.u)];
        g.p_v = prog.pm(t.len[r.u][r.k]);
        g.first = (int)prog.walk_visits.size();
        g.count = 0;
        prog.walk_groups.push_back(g);
        cur_u = r.u; cur_k = r.k;
      }
      {
        const int ko = (r.k + 1 + (1 - r.side)) % 3;
        const int other = t.nb[r.u][ko];
        st.push_back({(int)q, dir_slot[3 * other + t.slot_of(other, r.u)]});
      }
      while (!st.empty()) {
        const Frame f = st.back();
        st.pop_back();
        if (f.x_in <= -2) depth--;
        const Move &m1 = nb.moves[2 * f.q], &m2 = nb.moves[2 * f.q + 1];
        const int x = m1.x, c1 = t.nb[x][m1.kx], c2 = t.nb[x][m2.kx];
        double lin;
        if (m1.parent < 0) {
          const int ko = (m1.k + 1 + (1 - m1.side)) % 3, ks = (m1.k + 1 + m1.side) % 3;
          lin = t.len[m1.u][ko] + t.len[m1.u][ks];  // merged edge other -- x (rearrangement.py:480-490)
        } else {
          const Move &pm = nb.moves[m1.parent];
          lin = t.len[pm.x][pm.kx];
        }
        const double t1 = t.len[x][m1.kx], t2 = t.len[x][m2.kx];
        LikWalkVisit w;
        w.x_in = f.x_in; w.p_in = prog.pm(lin);
        w.d1 = dir_slot[3 * c1 + t.slot_of(c1, x)]; w.p_t1 = prog.pm(t1); w.p_h1 = prog.pm(t1 / 2.0);
        w.d2 = dir_slot[3 * c2 + t.slot_of(c2, x)]; w.p_t2 = prog.pm(t2); w.p_h2 = prog.pm(t2 / 2.0);
        w.tree1 = (m1.unique_id >= lo && m1.unique_id < hi) ? (int)(m1.unique_id - lo) : -1;
        w.tree2 = (m2.unique_id >= lo && m2.unique_id < hi) ? (int)(m2.unique_id - lo) : -1;
        w.keep = 0; w.push = -1;
        w.pad0 = w.pad1 = w.pad2 = w.pad3 = 0;
        const int ch1 = need[2 * f.q] ? child[2 * f.q] : -1, ch2 = need[2 * f.q + 1] ? child[2 * f.q + 1] : -1;
        if (ch1 >= 0 && ch2 >= 0) {
          const bool first1 = sub[ch1] <= sub[ch2];
          w.keep = first1 ? 1 : 2;
          w.push = depth;
          st.push_back({first1 ? ch2 : ch1, -2 - depth});
          st.push_back({first1 ? ch1 : ch2, -1});
          depth++;
          prog.walk_levels = std::max(prog.walk_levels, depth);
        } else if (ch1 >= 0) {
          w.keep = 1;
          st.push_back({ch1, -1});
        } else if (ch2 >= 0) {
          w.keep = 2;
          st.push_back({ch2, -1});
        }
        if (w.keep == 2) {  // normal form (walk4m_kernel): the partial toward d1 is the one that continues
          std::swap(w.d1, w.d2); std::swap(w.p_t1, w.p_t2); std::swap(w.p_h1, w.p_h2); std::swap(w.tree1, w.tree2);
          w.keep = 1;
        }
        prog.walk_visits.push_back(w);
        prog.walk_groups.back().count++;
      }
    }
    std::stable_sort(prog.walk_groups.begin(), prog.walk_groups.end(),
                     [](const LikWalkGroup &a, const LikWalkGroup &b) { return a.count > b.count; });
    int max_level = 0;
    for (auto &o : tmp) max_level = std::max(max_level, o.level);
    prog.level_off.assign(max_level + 2, 0);
    for (auto &o : tmp) prog.level_off[o.level + 1]++;
    for (int l = 0; l <= max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
    std::vector<int64_t> cur(prog.level_off.begin(), prog.level_off.end() - 1);
    prog.ops.resize(tmp.size());
    for (auto &o : tmp) prog.ops[cur[o.level]++] = o.op;
    prog.n_slots = slots;
    prog.n_trees = hi - lo;
    return;
  }
  std::vector<char> need = needed_moves(nb, lo, hi);
  std::vector<int> n_slot(nb.moves.size(), -1);
  for (size_t i = 0; i < nb.moves.size(); i++) {
    if (!need[i]) continue;
    const Move &m = nb.moves[i];
    const int x = m.x, c = t.nb[x][m.kx];
    int from, xin;
    double lin;
    if (m.parent < 0) {
      const int ko = (m.k + 1 + (1 - m.side)) % 3, ks = (m.k + 1 + m.side) % 3;
      const int other = t.nb[m.u][ko];
      from = m.u;
      xin = dir_slot[3 * other + t.slot_of(other, m.u)];
      lin = t.len[m.u][ko] + t.len[m.u][ks];  // merged edge other -- x (rearrangement.py:480-490)
    } else {
      const Move &pm = nb.moves[m.parent];
      from = pm.x;
      xin = n_slot[m.parent];
      lin = t.len[pm.x][pm.kx];
    }
    int s = -1, ks2 = -1;
    for (int j = 0; j < 3; j++)
      if (t.nb[x][j] != from && t.nb[x][j] != c) { s = t.nb[x][j]; ks2 = j; }
    Tmp o;
    o.op.dst = n_rows + (int)slots++;
    o.op.a = xin; o.op.pa = prog.pm(lin);
    o.op.b = dir_slot[3 * s + t.slot_of(s, x)]; o.op.pb = prog.pm(t.len[x][ks2]);
    o.level = d.max_level + m.depth;
    n_slot[i] = o.op.dst;
    tmp.push_back(o);
  }
  // hoisted operands: H[x, kx] = P_half D[c->x], TVV[u, k] = P_v D[v->u] (directional partials only,
  // so they all sit one level above the base tree's pass)
  std::vector<int> h_slot, tvv_slot;
  if (hoist_eval) { h_slot.assign((size_t)t.n_nodes * 3, -1); tvv_slot.assign((size_t)t.n_nodes * 3, -1); }
  auto hoisted = [&](std::vector<int> &memo, int e, int operand, double len) {
    if (memo[e] < 0) {
      Tmp o;
      o.op.dst = n_rows + (int)slots++;
      o.op.a = operand; o.op.pa = prog.pm(len);
      o.op.b = -1; o.op.pb = -1;
      o.level = d.max_level + 1;
      memo[e] = o.op.dst;
      tmp.push_back(o);
    }
    return memo[e];
  };
  for (int64_t i = lo; i < hi; i++) {
    const int mi = nb.unique[i];
    const Move &m = nb.moves[mi];
    const int v = t.nb[m.u][m.k], c = t.nb[m.x][m.kx];
    const double half = t.len[m.x][m.kx] / 2.0;  // target branch broken in half (rearrangement.py:466-470)
    LikEval ev;
    ev.a = n_slot[mi]; ev.pa = prog.pm(half);
    ev.b = dir_slot[3 * c + t.slot_of(c, m.x)]; ev.pb = prog.pm(half);
    ev.c = dir_slot[3 * v + t.slot_of(v, m.u)]; ev.pc = prog.pm(t.len[m.u][m.k]);
    if (hoist_eval) {
      ev.b = hoisted(h_slot, 3 * m.x + m.kx, ev.b, half); ev.pb = -1;
      ev.c = hoisted(tvv_slot, 3 * m.u + m.k, ev.c, t.len[m.u][m.k]); ev.pc = -1;
    }
    ev.tree = (int)(i - lo);
    ev.pad = 0;
    prog.evals.push_back(ev);
  }
  int max_level = 0;
  for (auto &o : tmp) max_level = std::max(max_level, o.level);
  prog.level_off.assign(max_level + 2, 0);
  for (auto &o : tmp) prog.level_off[o.level + 1]++;
  for (int l = 0; l <= max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
  std::vector<int64_t> cur(prog.level_off.begin(), prog.level_off.end() - 1);
  prog.ops.resize(tmp.size());
  for (auto &o : tmp) prog.ops[cur[o.level]++] = o.op;
  prog.n_slots = slots;
  prog.n_trees = hi - lo;
}

}  // namespace plg

// =============================================================================================
// TBR
// =============================================================================================
namespace plg {

TbrNeighbourhood enumerate_tbr(const UTree &t) {
  TbrNeighbourhood nb;
  nb.t = t;
  DirOrder d = dir_order(t);
  uint64_t all;
  std::vector<uint64_t> h = side_hashes(t, d, &all);
  uint64_t base = 0;
  for (int x = 0; x < t.n_nodes; x++)
    for (int k = 0; k < 3; k++)
      if (t.nb[x][k] > x) base += split_term(h[3 * x + k], all);
  nb.base_hash = base;
  nb.edge_cand_off.push_back(0);
  nb.edge_pair_off.push_back(0);
  if (t.n_tips < 4) return nb;
  std::vector<uint64_t> far((size_t)t.n_nodes * 3, 0), term((size_t)t.n_nodes * 3, 0);
  for (int x = 0; x < t.n_nodes; x++)
    for (int k = 0; k < 3; k++) {
      const int y = t.nb[x][k];
      if (y < 0) continue;
      far[3 * x + k] = h[3 * y + t.slot_of(y, x)];
      term[3 * x + k] = split_term(far[3 * x + k], all);
    }
  struct Frame { int x, from, parent, depth; uint64_t rem, add; };
  std::vector<Frame> st;
  // candidates of the part containing u when edge {u, nb[u][k]} is cut
  auto side_cands = [&](int u, int k) {
    TbrCand orig{u, k, -1, -1, 0, -1, 0, 0};
    nb.cands.push_back(orig);
    if (t.is_tip(u)) return;
    const uint64_t hv = far[3 * u + k];  // the other part moves relative to this one
    for (int side = 0; side < 2; side++) {
      const int ks = (k + 1 + side) % 3;
      st.push_back({t.nb[u][ks], u, -1, 0, term[3 * u + ks], 0});
      while (!st.empty()) {
        const Frame f = st.back();
        st.pop_back();
        if (t.is_tip(f.x)) continue;
        for (int j = 2; j >= 0; j--) {
          const int c = t.nb[f.x][j];
          if (c == f.from) continue;
          const uint64_t joined = split_term(far[3 * f.x + j] ^ hv, all);
          TbrCand m{u, k, f.x, j, f.depth + 1, f.parent, side, joined + f.add - f.rem};
          nb.cands.push_back(m);
          st.push_back({c, f.x, (int)nb.cands.size() - 1, f.depth + 1, f.rem + term[3 * f.x + j], f.add + joined});
        }
      }
    }
  };
  // distinct topologies: open-addressing set of the 64-bit hashes (about n^2 / 2 candidate pairs per
  // internal branch are tried; node-based containers made this loop the step's largest host cost)
  size_t cap = 1024;
  while (cap < (size_t)16 * t.n_tips * t.n_tips * (size_t)std::max(1, t.n_tips / 8)) cap <<= 1;
  cap = std::min<size_t>(cap, (size_t)1 << 28);
  std::vector<uint64_t> table(cap, 0);
  size_t filled = 0;
  auto insert = [&](uint64_t h) {  // true if new
    if (h == 0) h = 0x9E3779B97F4A7C15ull;  // 0 marks an empty cell
    if (2 * (filled + 1) > table.size()) {
      std::vector<uint64_t> old(table.size() * 2, 0);
      old.swap(table);
      for (uint64_t x : old)
        if (x) {
          size_t i = (size_t)(x * 0x9E3779B97F4A7C15ull >> 20) & (table.size() - 1);
          while (table[i]) i = (i + 1) & (table.size() - 1);
          table[i] = x;
        }
    }
    size_t i = (size_t)(h * 0x9E3779B97F4A7C15ull >> 20) & (table.size() - 1);
    while (table[i]) {
      if (table[i] == h) return false;
      i = (i + 1) & (table.size() - 1);
    }
    table[i] = h;
    filled++;
    return true;
  };
  insert(base);
  for (int u = 0; u < t.n_nodes; u++)
    for (int k = 0; k < 3; k++) {
      const int v = t.nb[u][k];
      if (v < u) continue;  // (also skips -1) each undirected edge once, from its smaller end
      nb.edges.push_back({u, k});
      side_cands(u, k);
      nb.edge_cand_off.push_back((int64_t)nb.cands.size());
      side_cands(v, t.slot_of(v, u));
      nb.edge_cand_off.push_back((int64_t)nb.cands.size());
      const size_t ne = nb.edges.size() - 1;
      for (int64_t c1 = nb.edge_cand_off[2 * ne]; c1 < nb.edge_cand_off[2 * ne + 1]; c1++)
        for (int64_t c2 = nb.edge_cand_off[2 * ne + 1]; c2 < nb.edge_cand_off[2 * ne + 2]; c2++) {
          const uint64_t hh = base + nb.cands[c1].delta + nb.cands[c2].delta;
          if (!insert(hh)) continue;
          nb.pairs.push_back({(int)c1, (int)c2, hh});
        }
      nb.edge_pair_off.push_back((int64_t)nb.pairs.size());
    }
  return nb;
}

UTree apply_tbr(const TbrNeighbourhood &nb, const TbrPair &p) {
  const UTree &t = nb.t;
  UTree r = t;
  auto move_side = [&](const TbrCand &c) {
    if (c.x < 0) return;  // original attachment
    const int u = c.u, v = t.nb[u][c.k];
    const int a = t.nb[u][(c.k + 1) % 3], b = t.nb[u][(c.k + 2) % 3];
    const double la = t.len[u][(c.k + 1) % 3], lb = t.len[u][(c.k + 2) % 3], lv = t.len[u][c.k];
    const int ka = r.slot_of(a, u), kb = r.slot_of(b, u);
    r.nb[a][ka] = b; r.len[a][ka] = la + lb;
    r.nb[b][kb] = a; r.len[b][kb] = la + lb;
    const int x = c.x, y = t.nb[x][c.kx];
    const double half = t.len[x][c.kx] / 2.0;
    const int kx = r.slot_of(x, y), ky = r.slot_of(y, x);
    r.nb[x][kx] = u; r.len[x][kx] = half;
    r.nb[y][ky] = u; r.len[y][ky] = half;
    r.nb[u] = {v, x, y};
    r.len[u] = {lv, half, half};
  };
  move_side(nb.cands[p.c1]);
  move_side(nb.cands[p.c2]);
  for (auto &e : r.rooted_child_end) e = -1;
  return r;
}

int64_t tbr_slots_needed(const TbrNeighbourhood &nb, int e0, int e1) {
  // per candidate: search-step partial N, rooted vector R and (likelihood) R pushed through the reconnecting branch
  return 3 * (int64_t)nb.t.n_nodes + 3 * (nb.edge_cand_off[2 * e1] - nb.edge_cand_off[2 * e0]);
}

namespace {

// Shared skeleton: walks directional ops, then the candidates of edges [e0,e1) (search-step
// partial N and rooted vector R per candidate), then the pairs.  The callbacks emit typed ops.
struct TbrLayout {
  DirOrder d;
  std::vector<int> dir_slot;   // operand id of D[x -> .]
  std::vector<int> n_slot, r_slot;  // per candidate (global index), operand ids
};

}  // namespace

void build_fitch_tbr_program(const TbrNeighbourhood &nb, const int32_t *tip_rows, int n_tip_slots, int e0, int e1,
                             FitchTbrProgram &out) {
  const UTree &t = nb.t;
  FitchProgram &prog = out.prog;
  prog.clear();
  DirOrder d = dir_order(t);
  out.dir_acc.assign((size_t)t.n_nodes * 3, -1);
  out.dir_kids.assign((size_t)t.n_nodes * 3, {-1, -1});
  std::vector<int> dir_slot((size_t)t.n_nodes * 3, -1);
  for (int x = 0; x < t.n_tips; x++)
    for (int k = 0; k < 3; k++) dir_slot[3 * x + k] = tip_rows ? tip_rows[x] : x;
  struct Tmp { FitchOp op; int level; int kids[2]; };
  std::vector<Tmp> tmp;
  int64_t slots = 0, accs = 0;
  for (int e : d.order) {
    const int x = e / 3, k = e % 3;
    Tmp o;
    o.op.dst = n_tip_slots + (int)slots++;
    o.op.nchild = 2;
    o.op.tree = (int)accs;
    out.dir_acc[e] = (int)accs++;
    int s = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        const int c = t.nb[x][j];
        const int ce = 3 * c + t.slot_of(c, x);
        o.kids[s] = dir_slot[ce];
        out.dir_kids[e][s] = t.is_tip(c) ? -1 : ce;
        s++;
      }
    o.level = d.level[e];
    dir_slot[e] = o.op.dst;
    tmp.push_back(o);
  }
  const int dump_acc = (int)accs++;
  out.first_pair_acc = accs;
  const int64_t c_lo = nb.edge_cand_off[2 * e0], c_hi = nb.edge_cand_off[2 * e1];
  std::vector<int> n_slot(c_hi - c_lo, -1), r_slot(c_hi - c_lo, -1);
  int max_depth = 0;
  for (int64_t ci = c_lo; ci < c_hi; ci++) {
    const TbrCand &c = nb.cands[ci];
    if (c.x < 0) {  // rooted on its original attachment: the directional set at u (or the tip itself)
      r_slot[ci - c_lo] = dir_slot[3 * c.u + c.k];
      continue;
    }
    const int x = c.x, y = t.nb[x][c.kx];
    int from, xin;
    if (c.parent < 0) {
      const int other = t.nb[c.u][(c.k + 1 + (1 - c.side)) % 3];
      from = c.u;
      xin = dir_slot[3 * other + t.slot_of(other, c.u)];
    } else {
      from = nb.cands[c.parent].x;
      xin = n_slot[c.parent - c_lo];
    }
    int s = -1;
    for (int j = 0; j < 3; j++)
      if (t.nb[x][j] != from && t.nb[x][j] != y) s = t.nb[x][j];
    Tmp o;
    o.op.dst = n_tip_slots + (int)slots++;
    o.op.nchild = 2;
    o.op.tree = dump_acc;
    o.kids[0] = xin;
    o.kids[1] = dir_slot[3 * s + t.slot_of(s, x)];
    o.level = d.max_level + 2 * c.depth - 1;
    n_slot[ci - c_lo] = o.op.dst;
    tmp.push_back(o);
    Tmp q;  // this part rooted in the middle of {x, y}
    q.op.dst = n_tip_slots + (int)slots++;
    q.op.nchild = 2;
    q.op.tree = dump_acc;
    q.kids[0] = o.op.dst;
    q.kids[1] = dir_slot[3 * y + t.slot_of(y, x)];
    q.level = o.level + 1;
    r_slot[ci - c_lo] = q.op.dst;
    tmp.push_back(q);
    max_depth = std::max(max_depth, c.depth);
  }
  const int pair_level = d.max_level + 2 * max_depth + 1;
  const int64_t p_lo = nb.edge_pair_off[e0], p_hi = nb.edge_pair_off[e1];
  for (int64_t pi = p_lo; pi < p_hi; pi++) {
    const TbrPair &p = nb.pairs[pi];
    Tmp o;
    o.op.dst = -1;
    o.op.nchild = 2;
    o.op.tree = (int)(out.first_pair_acc + (pi - p_lo));
    o.kids[0] = r_slot[p.c1 - c_lo];
    o.kids[1] = r_slot[p.c2 - c_lo];
    o.level = pair_level;
    tmp.push_back(o);
  }
  int max_level = 0;
  for (auto &o : tmp) max_level = std::max(max_level, o.level);
  prog.level_off.assign(max_level + 2, 0);
  for (auto &o : tmp) prog.level_off[o.level + 1]++;
  for (int l = 0; l <= max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
  std::vector<int64_t> cur(prog.level_off.begin(), prog.level_off.end() - 1);
  prog.ops.resize(tmp.size());
  prog.children.resize(tmp.size() * 2);
  for (auto &o : tmp) {
    const int64_t at = cur[o.level]++;
    o.op.child_off = (int)(2 * at);
    prog.children[2 * at] = o.kids[0];
    prog.children[2 * at + 1] = o.kids[1];
    prog.ops[at] = o.op;
  }
  prog.n_scratch_slots = slots;
  prog.n_costs = out.first_pair_acc + (p_hi - p_lo);
}

void finish_fitch_tbr(const TbrNeighbourhood &nb, const FitchTbrProgram &p, const int32_t *acc, int e0, int e1,
                      int32_t *out_costs) {
  const UTree &t = nb.t;
  DirOrder d = dir_order(t);
  std::vector<uint32_t> total((size_t)t.n_nodes * 3, 0);
  for (int e : d.order) {
    uint32_t s = (uint32_t)acc[p.dir_acc[e]];
    for (int j = 0; j < 2; j++)
      if (p.dir_kids[e][j] >= 0) s += total[p.dir_kids[e][j]];
    total[e] = s;
  }
  const int64_t p_lo = nb.edge_pair_off[e0];
  for (int e = e0; e < e1; e++) {
    const int u = nb.edges[e][0], k = nb.edges[e][1], v = t.nb[u][k];
    // both parts keep their own Fitch length wherever they are re-rooted (root invariance)
    const uint32_t parts = (t.is_tip(u) ? 0u : total[3 * u + k]) + (t.is_tip(v) ? 0u : total[3 * v + t.slot_of(v, u)]);
    for (int64_t pi = nb.edge_pair_off[e]; pi < nb.edge_pair_off[e + 1]; pi++)
      out_costs[pi - p_lo] = (int32_t)(parts + (uint32_t)acc[p.first_pair_acc + (pi - p_lo)]);
  }
}

void build_lik_tbr_program(const TbrNeighbourhood &nb, const int32_t *tip_rows, int n_rows, int e0, int e1,
                           LikProgram &prog) {
  const UTree &t = nb.t;
  prog.clear();
  DirOrder d = dir_order(t);
  std::vector<int> dir_slot((size_t)t.n_nodes * 3, -1);
  for (int x = 0; x < t.n_tips; x++)
    for (int k = 0; k < 3; k++) dir_slot[3 * x + k] = tip_rows ? tip_rows[x] : x;
  struct Tmp { LikOp op; int level; };
  std::vector<Tmp> tmp;
  int64_t slots = 0;
  for (int e : d.order) {
    const int x = e / 3, k = e % 3;
    Tmp o;
    o.op.dst = n_rows + (int)slots++;
    int s = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        const int c = t.nb[x][j];
        const int id = dir_slot[3 * c + t.slot_of(c, x)];
        const int pm = prog.pm(t.len[x][j]);
        if (s == 0) { o.op.a = id; o.op.pa = pm; } else { o.op.b = id; o.op.pb = pm; }
        s++;
      }
    o.level = d.level[e];
    dir_slot[e] = o.op.dst;
    tmp.push_back(o);
  }
  const int64_t c_lo = nb.edge_cand_off[2 * e0], c_hi = nb.edge_cand_off[2 * e1];
  std::vector<int> n_slot(c_hi - c_lo, -1), r_slot(c_hi - c_lo, -1);
  for (int64_t ci = c_lo; ci < c_hi; ci++) {
    const TbrCand &c = nb.cands[ci];
    if (c.x < 0) { r_slot[ci - c_lo] = dir_slot[3 * c.u + c.k]; continue; }
    const int x = c.x, y = t.nb[x][c.kx];
    int from, xin;
    double lin;
    if (c.parent < 0) {
      const int ko = (c.k + 1 + (1 - c.side)) % 3, ks = (c.k + 1 + c.side) % 3;
      const int other = t.nb[c.u][ko];
      from = c.u;
      xin = dir_slot[3 * other + t.slot_of(other, c.u)];
      lin = t.len[c.u][ko] + t.len[c.u][ks];
    } else {
      const TbrCand &pc = nb.cands[c.parent];
      from = pc.x;
      xin = n_slot[c.parent - c_lo];
      lin = t.len[pc.x][pc.kx];
    }
    int s = -1, ks2 = -1;
    for (int j = 0; j < 3; j++)
      if (t.nb[x][j] != from && t.nb[x][j] != y) { s = t.nb[x][j]; ks2 = j; }
    Tmp o;
    o.op.dst = n_rows + (int)slots++;
    o.op.a = xin; o.op.pa = prog.pm(lin);
    o.op.b = dir_slot[3 * s + t.slot_of(s, x)]; o.op.pb = prog.pm(t.len[x][ks2]);
    o.level = d.max_level + 2 * c.depth - 1;
    n_slot[ci - c_lo] = o.op.dst;
    tmp.push_back(o);
    const double half = t.len[x][c.kx] / 2.0;
    Tmp q;
    q.op.dst = n_rows + (int)slots++;
    q.op.a = o.op.dst; q.op.pa = prog.pm(half);
    q.op.b = dir_slot[3 * y + t.slot_of(y, x)]; q.op.pb = prog.pm(half);
    q.level = o.level + 1;
    r_slot[ci - c_lo] = q.op.dst;
    tmp.push_back(q);
  }
  // The reconnecting branch is the same for every pair of a bisection: push each second-part
  // candidate through it ONCE (Q = P_v R2, a one-operand update) instead of once per pair; a pair
  // is then an elementwise product of two stored vectors.
  int top_level = 0;
  for (auto &o : tmp) top_level = std::max(top_level, o.level);
  std::vector<int> q_slot(c_hi - c_lo, -1);
  const int64_t p_lo = nb.edge_pair_off[e0];
  for (int e = e0; e < e1; e++) {
    const int u = nb.edges[e][0], k = nb.edges[e][1];
    const int pmv = prog.pm(t.len[u][k]);  // the reconnecting branch keeps the removed branch's length
    // ... on the side with fewer candidates (the model is reversible: pi_k a_k (P b)_k sums to the
    // same as pi_k (P a)_k b_k); a pendant branch has ONE candidate on its tip side
    const bool push_second = nb.edge_cand_off[2 * e + 2] - nb.edge_cand_off[2 * e + 1] <=
                             nb.edge_cand_off[2 * e + 1] - nb.edge_cand_off[2 * e];
    for (int64_t pi = nb.edge_pair_off[e]; pi < nb.edge_pair_off[e + 1]; pi++) {
      const TbrPair &p = nb.pairs[pi];
      const int64_t cq = push_second ? p.c2 : p.c1, cr = push_second ? p.c1 : p.c2;
      int &q = q_slot[cq - c_lo];
      if (q < 0) {
        Tmp o;
        o.op.dst = n_rows + (int)slots++;
        o.op.a = r_slot[cq - c_lo]; o.op.pa = pmv;
        o.op.b = -1; o.op.pb = -1;
        o.level = top_level + 1;
        q = o.op.dst;
        tmp.push_back(o);
      }
      LikEval ev;  // pairs are listed c1-major: operand a is the one that stays while c2 runs
      ev.a = push_second ? r_slot[cr - c_lo] : q; ev.pa = -1;
      ev.b = push_second ? q : r_slot[cr - c_lo]; ev.pb = -1;
      ev.c = -1; ev.pc = -1; ev.pad = 0;
      ev.tree = (int)(pi - p_lo);
      prog.evals.push_back(ev);
    }
  }
  int max_level = 0;
  for (auto &o : tmp) max_level = std::max(max_level, o.level);
  prog.level_off.assign(max_level + 2, 0);
  for (auto &o : tmp) prog.level_off[o.level + 1]++;
  for (int l = 0; l <= max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
  std::vector<int64_t> cur(prog.level_off.begin(), prog.level_off.end() - 1);
  prog.ops.resize(tmp.size());
  for (auto &o : tmp) prog.ops[cur[o.level]++] = o.op;
  prog.n_slots = slots;
  prog.n_trees = nb.edge_pair_off[e1] - p_lo;
}

int64_t tbr20_slots_needed(const TbrNeighbourhood &nb, int e0, int e1) {
  // standard layout: D, T, H per directed edge + N per candidate; pair layout: one vector per candidate
  return 9 * (int64_t)nb.t.n_nodes + 2 * (nb.edge_cand_off[2 * e1] - nb.edge_cand_off[2 * e0]);
}

void build_lik_tbr20_program(const TbrNeighbourhood &nb, const int32_t *tip_rows, int n_rows, int e0, int e1,
                             LikProgram &prog) {
  const UTree &t = nb.t;
  prog.clear();
  DirOrder d = dir_order(t);
  std::vector<int> dir_slot((size_t)t.n_nodes * 3, -1);
  for (int x = 0; x < t.n_tips; x++)
    for (int k = 0; k < 3; k++) dir_slot[3 * x + k] = tip_rows ? tip_rows[x] : x;
  int64_t slots = 0;
  {  // base tree: directional vectors, level ops as in build_lik_tbr_program
    struct Tmp { LikOp op; int level; };
    std::vector<Tmp> tmp;
    for (int e : d.order) {
      const int x = e / 3, k = e % 3;
      Tmp o;
      o.op.dst = n_rows + (int)slots++;
      int s = 0;
      for (int j = 0; j < 3; j++)
        if (j != k) {
          const int c = t.nb[x][j];
          const int id = dir_slot[3 * c + t.slot_of(c, x)];
          const int pm = prog.pm(t.len[x][j]);
          if (s == 0) { o.op.a = id; o.op.pa = pm; } else { o.op.b = id; o.op.pb = pm; }
          s++;
        }
      o.level = d.level[e];
      dir_slot[e] = o.op.dst;
      tmp.push_back(o);
    }
    int max_level = 0;
    for (auto &o : tmp) max_level = std::max(max_level, o.level);
    prog.level_off.assign(max_level + 2, 0);
    for (auto &o : tmp) prog.level_off[o.level + 1]++;
    for (int l = 0; l <= max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
    std::vector<int64_t> cur(prog.level_off.begin(), prog.level_off.end() - 1);
    prog.ops.resize(tmp.size());
    for (auto &o : tmp) prog.ops[cur[o.level]++] = o.op;
  }
  struct TOp { Tbr20Op op; int level; };
  std::vector<TOp> ops;
  const double Cb = 4 * 20 * 8.0 + 4.0;
  auto opnd = [&](int id) { return id < 0 ? 0.0 : (id < n_rows ? 1.0 : Cb); };
  auto emit = [&](const Tbr20Op &o, int level) {
    ops.push_back({o, level});
    prog.t20_products += 1.0 + (o.p2 >= 0) + (o.p3 >= 0);
    prog.t20_bytes += opnd(o.a) + opnd(o.b) + opnd(o.b2) + (o.dst >= 0 ? Cb : 0.0) + (o.pdst >= 0 ? Cb : 0.0);
  };
  // T[y -> x] = P(len) D[y -> x], H[y -> x] = P(len / 2) D[y -> x], indexed like dir_slot (3 * y + slot of x),
  // made on first use
  std::vector<int> t_slot((size_t)t.n_nodes * 3, -1), h_slot((size_t)t.n_nodes * 3, -1);
  auto pushed = [&](std::vector<int> &tab, int y, int x, double len) {
    int &s = tab[3 * y + t.slot_of(y, x)];
    if (s < 0) {
      s = n_rows + (int)slots++;
      emit({dir_slot[3 * y + t.slot_of(y, x)], prog.pm(len), -1, s, -1, -1, -1, -1}, 0);
    }
    return s;
  };
  const int64_t c_lo = nb.edge_cand_off[2 * e0], c_hi = nb.edge_cand_off[2 * e1];
  const int64_t p_lo = nb.edge_pair_off[e0];
  // which candidates carry a pair, which side of every bisection is pushed through the reconnecting branch
  std::vector<char> used(c_hi - c_lo, 0), has_child(c_hi - c_lo, 0);
  for (int64_t pi = p_lo; pi < nb.edge_pair_off[e1]; pi++) {
    used[nb.pairs[pi].c1 - c_lo] = 1;
    used[nb.pairs[pi].c2 - c_lo] = 1;
  }
  for (int64_t ci = c_lo; ci < c_hi; ci++)
    if (nb.cands[ci].parent >= 0) has_child[nb.cands[ci].parent - c_lo] = 1;
  std::vector<int> n_slot(c_hi - c_lo, -1), pair_slot(c_hi - c_lo, -1);
  int64_t pslots = 0;
  for (int e = e0; e < e1; e++) {
    const int u = nb.edges[e][0], k = nb.edges[e][1];
    const int pmv = prog.pm(t.len[u][k]);  // the reconnecting branch keeps the removed branch's length
    const int64_t s0 = nb.edge_cand_off[2 * e], s1 = nb.edge_cand_off[2 * e + 1], s2 = nb.edge_cand_off[2 * e + 2];
    const bool push_second = s2 - s1 <= s1 - s0;  // the side with fewer candidates takes the extra product
    for (int64_t ci = s0; ci < s2; ci++) {
      const TbrCand &c = nb.cands[ci];
      const bool push = (ci >= s1) == push_second;
      const bool want_pair = used[ci - c_lo];
      int pdst = -1;
      if (want_pair) {
        pair_slot[ci - c_lo] = (int)pslots++;
        pdst = pair_slot[ci - c_lo] | (push ? 0 : (1 << 30));  // the stay side carries pi
      }
      if (c.x < 0) {  // original attachment: the part as it hangs on the cut branch
        if (want_pair) emit({dir_slot[3 * c.u + c.k], push ? pmv : prog.pm(0.0), -1, -1, -1, -1, -1, pdst}, 0);
        continue;
      }
      const int x = c.x, y = t.nb[x][c.kx];
      int from, xin;
      double lin;
      if (c.parent < 0) {
        const int ko = (c.k + 1 + (1 - c.side)) % 3, ks = (c.k + 1 + c.side) % 3;
        const int other = t.nb[c.u][ko];
        from = c.u;
        xin = dir_slot[3 * other + t.slot_of(other, c.u)];
        lin = t.len[c.u][ko] + t.len[c.u][ks];
      } else {
        const TbrCand &pc = nb.cands[c.parent];
        from = pc.x;
        xin = n_slot[c.parent - c_lo];
        lin = t.len[pc.x][pc.kx];
      }
      int s = -1, ks2 = -1;
      for (int j = 0; j < 3; j++)
        if (t.nb[x][j] != from && t.nb[x][j] != y) { s = t.nb[x][j]; ks2 = j; }
      if (!want_pair && !has_child[ci - c_lo]) continue;
      Tbr20Op o;
      o.a = xin; o.pa = prog.pm(lin);
      o.b = pushed(t_slot, s, x, t.len[x][ks2]);
      o.dst = has_child[ci - c_lo] ? (n_slot[ci - c_lo] = n_rows + (int)slots++) : -1;
      o.p2 = o.b2 = o.p3 = -1;
      o.pdst = pdst;
      if (want_pair) {
        const double half = t.len[x][c.kx] / 2.0;
        o.p2 = prog.pm(half);
        o.b2 = pushed(h_slot, y, x, half);
        o.p3 = push ? pmv : -1;
      }
      emit(o, c.depth);
    }
    // pair grid of this bisection: rows = stay side, columns = pushed side
    std::vector<int64_t> rows, cols;
    for (int64_t ci = s0; ci < s2; ci++)
      if (used[ci - c_lo]) ((ci >= s1) == push_second ? cols : rows).push_back(ci);
    std::vector<int> pos(s2 - s0, -1);
    for (size_t i = 0; i < rows.size(); i++) pos[rows[i] - s0] = (int)i;
    for (size_t i = 0; i < cols.size(); i++) pos[cols[i] - s0] = (int)i;
    const int ntr = (int)(rows.size() + 7) / 8, ntc = (int)(cols.size() + 7) / 8;
    const size_t tile0 = prog.t20_tiles.size();
    prog.t20_tile_off.push_back((int)tile0);
    prog.t20_tiles.resize(tile0 + (size_t)ntr * ntc);
    for (int tr = 0; tr < ntr; tr++)
      for (int tc = 0; tc < ntc; tc++) {
        Tbr20Tile &T = prog.t20_tiles[tile0 + (size_t)tr * ntc + tc];
        for (int i = 0; i < 8; i++) {
          const size_t ri = std::min(rows.size() - 1, (size_t)tr * 8 + i), cj = std::min(cols.size() - 1, (size_t)tc * 8 + i);
          T.a[i] = pair_slot[rows[ri] - c_lo];
          T.b[i] = pair_slot[cols[cj] - c_lo];
        }
        for (int i = 0; i < 64; i++) T.tree[i] = -1;
      }
    for (int64_t pi = nb.edge_pair_off[e]; pi < nb.edge_pair_off[e + 1]; pi++) {
      const TbrPair &p = nb.pairs[pi];
      const int64_t cr = push_second ? p.c1 : p.c2, cc = push_second ? p.c2 : p.c1;
      const int ri = pos[cr - s0], cj = pos[cc - s0];
      prog.t20_tiles[tile0 + (size_t)(ri / 8) * ntc + cj / 8].tree[(ri % 8) * 8 + cj % 8] = (int)(pi - p_lo);
    }
    prog.t20_pair_bytes += (double)ntr * ntc * 16.0 * Cb;
  }
  prog.t20_tile_off.push_back((int)prog.t20_tiles.size());
  int max_level = 0;
  for (auto &o : ops) max_level = std::max(max_level, o.level);
  prog.t20_level_off.assign(max_level + 2, 0);
  for (auto &o : ops) prog.t20_level_off[o.level + 1]++;
  for (int l = 0; l <= max_level; l++) prog.t20_level_off[l + 1] += prog.t20_level_off[l];
  std::vector<int64_t> cur(prog.t20_level_off.begin(), prog.t20_level_off.end() - 1);
  prog.t20_ops.resize(ops.size());
  for (auto &o : ops) prog.t20_ops[cur[o.level]++] = o.op;
  prog.n_slots = slots;
  prog.t20_pair_slots = pslots;
  prog.n_trees = nb.edge_pair_off[e1] - p_lo;
}

}  // namespace plg
