// libpllWrapper compat: the opaque "problem" handle of libpllWrapper.c (:12-19, :97-239) on the
// B200 engine -- new / destroy / getLogLikelihood / getNewickString.
//
// getLogLikelihood in the reference is pllOptimizeBranchLengths(inst, partitions, 3) followed by
// instance->likelihood (libpllWrapper.c:205-209) on a tree loaded with DEFAULT branch lengths
// (pllTreeInitTopologyNewick(..., PLL_TRUE), :152) under pllInitModel defaults (:159).  libpll's
// source is not available offline, so what follows restates the published procedure it implements
// (marked "from memory" there): branch-by-branch Newton-Raphson on log z
// (z = exp(-t)), start z = 0.9, one Newton step per branch per smoothing pass, at most 3 passes,
// z clamped to [1e-15, 1-1e-6], depth-first from the first taxon with on-demand re-orientation of
// the conditional likelihood vectors.  Parity with libpll's number is UNPINNED;
// parity with the CPU oracle at the resulting branch lengths is tested.
//
// Device work: sumtable_kernel (per-pattern eigen-space products of the two CLVs facing a
// branch), deriv_kernel (lnL, d/dlz, d2/dlz2 per pattern, block-reduced), plus the CLV update
// kernels of lik.cu for the re-orientations.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <fstream>
#include <memory>
#include <sstream>
#include <sys/stat.h>
#include <unordered_map>

#include "hostpool.h"
#include "lik.h"
#include "lik_pmatrix.cuh"
#include "neighbourhood.h"

namespace plg {

constexpr double PLL_ZMIN = 1.0e-15, PLL_ZMAX = 1.0 - 1.0e-6, PLL_DEFAULTZ = 0.9, PLL_DELTAZ = 1.0e-5;
constexpr double SM_LOG_MINLH = -256.0 * 0.693147180559945309417232121458176568;  // log(2^-256), folded
constexpr int SM_THREADS = 256;

bool protein_model_table(const char *name, const double **exch, const double **freqs);  // protein_models.cpp

// S[r*K+k][p] = (sum_i pi_i A_r[i] U[i][k]) * (sum_j Uinv[k][j] B_r[j]);  L_p(t) = (1/R) sum S exp(eval_k rate_r t)
template <int K, int R>
__global__ void __launch_bounds__(SM_THREADS)
sumtable_kernel(int a_id, int b_id, int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
                const unsigned int *__restrict__ masks, const ModelDev *__restrict__ model,
                const double *__restrict__ clv, const int *__restrict__ scl, double *__restrict__ S,
                int *__restrict__ sc_out) {
  const int r = threadIdx.x % R;
  const int64_t p = (int64_t)blockIdx.x * (SM_THREADS / R) + threadIdx.x / R;
  double A[K], B[K];
  int sc = 0;
  auto fetch = [&](int id, double(&x)[K]) {
    if (id < n_rows) {
      const unsigned int m = masks[codes[(int64_t)id * Ppad + p]];
#pragma unroll
      for (int k = 0; k < K; k++) x[k] = (m >> k & 1) ? 1.0 : 0.0;
    } else {
      const double *src = clv + ((int64_t)(id - n_rows) * R * K + r * K) * Ppad + p;
#pragma unroll
      for (int k = 0; k < K; k++) x[k] = src[(int64_t)k * Ppad];
      sc += scl[(int64_t)(id - n_rows) * Ppad + p];
    }
  };
  fetch(a_id, A);
  fetch(b_id, B);
#pragma unroll
  for (int i = 0; i < K; i++) A[i] *= model->freqs[i];
  for (int k = 0; k < K; k++) {
    double l = 0.0, rr = 0.0;
#pragma unroll
    for (int i = 0; i < K; i++) {
      l = fma(A[i], model->U[i * K + k], l);
      rr = fma(model->Uinv[k * K + i], B[i], rr);
    }
    S[((int64_t)r * K + k) * Ppad + p] = l * rr;
  }
  if (r == 0) sc_out[p] = sc;
}

struct ExpTable {  // e0 = exp(eval*rate*t), e1 = (eval*rate) e0 * dt/dlz, e2 likewise squared
  double e0[LIK_MAX_CATS * LIK_MAX_STATES], e1[LIK_MAX_CATS * LIK_MAX_STATES], e2[LIK_MAX_CATS * LIK_MAX_STATES];
};

template <int K, int R>
__global__ void __launch_bounds__(SM_THREADS)
deriv_kernel(const double *__restrict__ S, const int *__restrict__ sc, const double *__restrict__ weights,
             const ExpTable *__restrict__ tab, int64_t Ppad, double *__restrict__ partial, int n_blocks) {
  __shared__ double ws[3][SM_THREADS / 32];
  const int r = threadIdx.x % R;
  const int64_t p = (int64_t)blockIdx.x * (SM_THREADS / R) + threadIdx.x / R;
  double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) {
    const double s = S[((int64_t)r * K + k) * Ppad + p];
    l0 = fma(s, tab->e0[r * K + k], l0);
    l1 = fma(s, tab->e1[r * K + k], l1);
    l2 = fma(s, tab->e2[r * K + k], l2);
  }
#pragma unroll
  for (int d = 1; d < R; d <<= 1) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, d);
    l1 += __shfl_xor_sync(0xffffffffu, l1, d);
    l2 += __shfl_xor_sync(0xffffffffu, l2, d);
  }
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  const double w = weights[p];
  if (r == 0 && w != 0.0) {
    const double inv = 1.0 / l0, d1 = l1 * inv;
    v0 = w * (log(l0 * (1.0 / R)) + sc[p] * SM_LOG_MINLH);
    v1 = w * d1;
    v2 = w * (l2 * inv - d1 * d1);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v0 += __shfl_down_sync(0xffffffffu, v0, d);
    v1 += __shfl_down_sync(0xffffffffu, v1, d);
    v2 += __shfl_down_sync(0xffffffffu, v2, d);
  }
  if ((threadIdx.x & 31) == 0) {
    ws[0][threadIdx.x >> 5] = v0; ws[1][threadIdx.x >> 5] = v1; ws[2][threadIdx.x >> 5] = v2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double s = 0.0;
    for (int i = 0; i < SM_THREADS / 32; i++) s += ws[threadIdx.x][i];
    partial[(int64_t)threadIdx.x * n_blocks + blockIdx.x] = s;
  }
}

// ---------------------------------------------------------------------------------------------
struct PllProblem {
  std::shared_ptr<LikAln> aln;   // shared with the alignment cache (read-only after construction)
  std::shared_ptr<Model> model;
  UTree tree;
  std::vector<std::string> tip_names;  // tip id -> label
  std::vector<int32_t> tip_rows;       // tip id -> alignment row
  bool smoothed = false;
  double lnl = 0.0;

  // result of the last plg_pll_moves_compute (libpllWrapper.c:36-93 builds the same triples)
  std::vector<int32_t> mv_kind;
  std::vector<float> mv_lnl;
  std::vector<std::string> mv_newick;

  // scratch (device) for smoothing
  DevBuf<double> S;
  DevBuf<int> Ssc;
  DevBuf<double> partial;
  DevBuf<ExpTable> tab;
  std::vector<double> h_partial;

  int dir_id(int x, int k) const { return x < tree.n_tips ? tip_rows[x] : aln->n_rows + 3 * x + k; }
  int dir_to(int x, int y) const { return dir_id(x, tree.slot_of(x, y)); }
  int64_t n_slots() const { return 3 * (int64_t)tree.n_nodes; }

  void clv_update(int x, int k) {  // D[x -> nb[x][k]] from the two other neighbours' partials
    LikProgram prog;
    LikOp op;
    op.dst = dir_id(x, k);
    int s = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        const int c = tree.nb[x][j];
        const int id = dir_to(c, x), pm = prog.pm(tree.len[x][j]);
        if (s == 0) { op.a = id; op.pa = pm; } else { op.b = id; op.pb = pm; }
        s++;
      }
    prog.ops.push_back(op);
    prog.level_off = {0, 1};
    prog.n_slots = n_slots();
    prog.n_trees = 0;
    run_lik_program(*aln, *model, lik_workspace(), prog, nullptr, 0);
  }

  template <int K, int R>
  void derivs_typed(int a_id, int b_id, bool build_table, double lz, double out[3]) {
    LikWorkspace &ws = lik_workspace();
    const int per_block = SM_THREADS / R;
    const int n_blocks = (int)(aln->Ppad / per_block);
    if (build_table) {
      sumtable_kernel<K, R><<<n_blocks, SM_THREADS>>>(a_id, b_id, aln->n_rows, aln->Ppad, aln->codes.p, aln->masks.p,
                                                      model->d.p, ws.clv.p, ws.scl.p, S.p, Ssc.p);
      count_launch();
    }
    ExpTable h;
    for (int r = 0; r < R; r++)
      for (int k = 0; k < K; k++) {
        const double g = -model->h.eval[k] * model->h.rates[r];  // t = -lz: d/dlz exp(eval*rate*t) = g * exp(.)
        const double e = std::exp(g * lz);
        h.e0[r * K + k] = e; h.e1[r * K + k] = g * e; h.e2[r * K + k] = g * g * e;
      }
    PLG_CUDA(cudaMemcpyAsync(tab.p, &h, sizeof h, cudaMemcpyHostToDevice, 0));
    deriv_kernel<K, R><<<n_blocks, SM_THREADS>>>(S.p, Ssc.p, aln->weights.p, tab.p, aln->Ppad, partial.p, n_blocks);
    count_launch();
    PLG_CUDA(cudaMemcpy(h_partial.data(), partial.p, (size_t)3 * n_blocks * 8, cudaMemcpyDeviceToHost));
    for (int q = 0; q < 3; q++) {
      double s = 0.0;
      for (int b = 0; b < n_blocks; b++) s += h_partial[(size_t)q * n_blocks + b];  // fixed order
      out[q] = s;
    }
  }

  void derivs(int a_id, int b_id, bool build_table, double lz, double out[3]) {
    const int K = aln->K, R = model->R;
    if (K == 4 && R == 4) derivs_typed<4, 4>(a_id, b_id, build_table, lz, out);
    else if (K == 4 && R == 1) derivs_typed<4, 1>(a_id, b_id, build_table, lz, out);
    else if (K == 20 && R == 4) derivs_typed<20, 4>(a_id, b_id, build_table, lz, out);
    else if (K == 20 && R == 1) derivs_typed<20, 1>(a_id, b_id, build_table, lz, out);
    else PLG_FAIL(PLG_ERR_UNSUPPORTED, "kernels exist for 4 or 20 states with 1 or 4 rate categories");
  }

  // One makenewz-style update of branch {x, y}: `maxiter` Newton steps on lz = log z.  Returns the new z.
  double optimise_branch(int x, int y, int maxiter, double *lnl_out) {
    const int kx = tree.slot_of(x, y), ky = tree.slot_of(y, x);
    double z = std::exp(-tree.len[x][kx]);
    z = std::min(std::max(z, PLL_ZMIN), PLL_ZMAX);
    const int a_id = dir_to(x, y), b_id = dir_to(y, x);
    bool first = true;
    double d[3] = {0, 0, 0};
    while (maxiter-- > 0) {
      double zprev = z;
      const double zstep = (1.0 - PLL_ZMAX) * z + PLL_ZMIN;
      for (int guard = 0; guard < 64; guard++) {  // "bad curvature: shorten the branch" loop
        derivs(a_id, b_id, first, std::log(std::max(z, PLL_ZMIN)), d);
        first = false;
        if (d[2] >= 0.0 && z < PLL_ZMAX) zprev = z = 0.37 * z + 0.63;
        else break;
      }
      if (d[2] < 0.0) {
        const double tantmp = -d[1] / d[2];
        if (tantmp < 100.0) {
          z *= std::exp(tantmp);
          if (z < PLL_ZMIN) z = PLL_ZMIN;
          if (z > 0.25 * zprev + 0.75) z = 0.25 * zprev + 0.75;
        } else {
          z = 0.25 * zprev + 0.75;
        }
      }
      if (z > PLL_ZMAX) z = PLL_ZMAX;
      if (std::fabs(z - zprev) <= zstep) break;
    }
    const double t = -std::log(z);
    tree.len[x][kx] = tree.len[y][ky] = t;
    if (lnl_out) *lnl_out = d[0];
    return z;
  }

  // smooth(): optimise the branch above x, then its subtree, then restore x's down view
  void smooth(int x, int from, bool *all_smoothed) {
    const int kf = tree.slot_of(x, from);
    const double z0 = std::exp(-tree.len[x][kf]);
    const double z = optimise_branch(x, from, 1, nullptr);
    if (std::fabs(z - z0) > PLL_DELTAZ) *all_smoothed = false;
    if (tree.is_tip(x)) return;
    for (int j = 0; j < 3; j++) {
      if (j == kf) continue;
      clv_update(x, j);  // view from x toward this child: needs the (just re-optimised) branch above x
      smooth(tree.nb[x][j], x, all_smoothed);
    }
    clv_update(x, kf);  // children's branches changed: refresh x's down view
  }

  double evaluate() {  // lnL across the branch at tip 0 with all down views fresh
    const int w = tree.nb[0][0];
    LikProgram prog;
    LikEval ev;
    ev.a = tip_rows[0]; ev.pa = -1;
    ev.b = dir_to(w, 0); ev.pb = prog.pm(tree.len[0][0]);
    ev.c = -1; ev.pc = -1; ev.tree = 0; ev.pad = 0;
    prog.evals.push_back(ev);
    prog.level_off = {0, 0};
    prog.n_slots = n_slots();
    prog.n_trees = 1;
    double out = 0.0;
    run_lik_program(*aln, *model, lik_workspace(), prog, &out, 0);
    return out;
  }

  void down_pass() {  // every D[x -> parent] with tip 0 as the root, leaves first
    std::vector<int> par(tree.n_nodes, -1), bfs{0};
    for (size_t i = 0; i < bfs.size(); i++)
      for (int k = 0; k < 3; k++) {
        const int y = tree.nb[bfs[i]][k];
        if (y >= 0 && y != par[bfs[i]]) { par[y] = bfs[i]; bfs.push_back(y); }
      }
    for (size_t i = bfs.size(); i-- > 1;)
      if (!tree.is_tip(bfs[i])) clv_update(bfs[i], tree.slot_of(bfs[i], par[bfs[i]]));
  }

  double loglik(int max_passes);

  // the literal per-branch host loop (one device round trip per Newton step): single-category models,
  // and the A/B reference of the lockstep path (PLG_PLL_HANDLE_MODE=host)
  double loglik_host_loop(int max_passes) {
    if (tree.n_tips < 3) PLG_FAIL(PLG_ERR_UNSUPPORTED, "need at least 3 taxa");
    const int per_block = SM_THREADS / model->R;
    const int n_blocks = (int)(aln->Ppad / per_block);
    S.alloc((size_t)model->R * aln->K * aln->Ppad);
    Ssc.alloc(aln->Ppad);
    partial.alloc((size_t)3 * n_blocks);
    tab.alloc(1);
    h_partial.resize((size_t)3 * n_blocks);
    down_pass();
    for (int pass = 0; pass < max_passes; pass++) {  // smoothTree(maxtimes)
      bool all_smoothed = true;
      smooth(tree.nb[0][0], 0, &all_smoothed);
      if (all_smoothed) break;
    }
    lnl = evaluate();
    smoothed = true;
    return lnl;
  }

  void newick_rec(const UTree &tree, int x, int from, std::ostringstream &os) const {
    if (tree.is_tip(x)) { os << tip_names[x]; return; }
    os << "(";
    bool first = true;
    for (int j = 0; j < 3; j++) {
      const int c = tree.nb[x][j];
      if (c == from) continue;
      if (!first) os << ",";
      first = false;
      newick_rec(tree, c, x, os);
      os << ":" << tree.len[x][j];
    }
    os << ")";
  }

  std::string newick() const { return newick_of(tree); }

  std::string newick_of(const UTree &tree) const {  // unrooted, from the inner node next to the first taxon (:227-229)
    std::ostringstream os;
    os.precision(17);
    const int w = tree.nb[0][0];
    os << "(" << tip_names[0] << ":" << tree.len[0][0];
    for (int j = 0; j < 3; j++) {
      const int c = tree.nb[w][j];
      if (c == 0) continue;
      os << ",";
      newick_rec(tree, c, w, os);
      os << ":" << tree.len[w][j];
    }
    os << ");";
    return os.str();
  }
};

// ---------------------------------------------------------------------------------------------
// PHYLIP, relaxed names.  Interleaved (n first lines "name seq...", then blocks of n continuation
// lines) is tried first -- it is what alignment.py writes (one line per taxon is its one-block case);
// a file that does not come out at n x m that way is read as SEQUENTIAL (each taxon's sequence runs
// over as many lines as it needs before the next name).
static bool read_phylip_as(std::istream &in, size_t n, size_t m, bool sequential, std::vector<std::string> &names,
                           std::vector<std::string> &seqs) {
  names.assign(n, "");
  seqs.assign(n, "");
  std::string line;
  size_t row = 0;
  bool header_done = false;
  while (std::getline(in, line)) {
    std::istringstream ls(line);
    std::string tok;
    std::vector<std::string> toks;
    while (ls >> tok) toks.push_back(tok);
    if (toks.empty()) continue;
    size_t from = 0;
    if (sequential) {
      if (row == n) break;
      if (names[row].empty()) { names[row] = toks[0]; from = 1; }
      for (size_t i = from; i < toks.size(); i++) seqs[row] += toks[i];
      if (seqs[row].size() >= m) row++;
      continue;
    }
    if (!header_done) { names[row] = toks[0]; from = 1; }
    for (size_t i = from; i < toks.size(); i++) seqs[row] += toks[i];
    if (++row == n) { row = 0; header_done = true; }
    bool full = header_done;
    for (auto &s : seqs) full = full && s.size() >= m;
    if (full) break;
  }
  for (auto &s : seqs)
    if (s.size() != m) return false;
  return true;
}

static void read_phylip(const char *path, std::vector<std::string> &names, std::vector<std::string> &seqs) {
  std::ifstream in(path);
  if (!in) PLG_FAIL(PLG_ERR_IO, "Alignment file could not be parsed.");
  size_t n = 0, m = 0;
  in >> n >> m;
  if (!in || n == 0 || m == 0) PLG_FAIL(PLG_ERR_IO, "Alignment file could not be parsed.");
  std::string line;
  std::getline(in, line);
  const std::streampos body = in.tellg();
  if (read_phylip_as(in, n, m, false, names, seqs)) return;
  in.clear();
  in.seekg(body);
  if (!read_phylip_as(in, n, m, true, names, seqs)) PLG_FAIL(PLG_ERR_IO, "Alignment file could not be parsed.");
}

// Partition file as pll.py:112-137 writes it: one line "<MODEL>, <name> = <from>-<to>" per partition.
// One partition covering the whole alignment is what the engine scores (createSimpleModel,
// pll.py:106-116 -- the only form landscape.py / scoring.py produce).  Several partitions (createModel,
// pll.py:118-137) or a partial range would need one model per site block with joint branch lengths:
// refused with PLG_ERR_UNSUPPORTED rather than scored under the first line's model.
static std::string read_partition_model(const char *path, long long *range_end) {
  std::ifstream in(path);
  if (!in) PLG_FAIL(PLG_ERR_IO, "Partition file could not be read.");
  std::string line, model;
  int n_parts = 0;
  *range_end = -1;
  while (std::getline(in, line)) {
    if (line.find_first_not_of(" \t\r\n") == std::string::npos) continue;
    const size_t comma = line.find(','), eq = line.find('=');
    if (comma == std::string::npos || eq == std::string::npos || eq < comma)
      PLG_FAIL(PLG_ERR_IO, "Partition file could not be parsed.");
    if (++n_parts > 1)
      PLG_FAIL(PLG_ERR_UNSUPPORTED, "partition file with more than one partition: per-partition models are not implemented "
               "(pll.partitionModel.createModel, pll.py:118-137); use createSimpleModel");
    model = line.substr(0, comma);
    model.erase(std::remove_if(model.begin(), model.end(), [](char c) { return c == ' ' || c == '\t' || c == '\r'; }),
                model.end());
    long long from = 0, to = 0;
    char dash = 0;
    std::istringstream rs(line.substr(eq + 1));
    if (!(rs >> from >> dash >> to) || dash != '-' || from < 1 || to < from)
      PLG_FAIL(PLG_ERR_IO, "Partition file could not be parsed.");
    std::string rest;
    if (rs >> rest) PLG_FAIL(PLG_ERR_UNSUPPORTED, "partition range '%s': only one contiguous range is supported", line.c_str());
    if (from != 1) PLG_FAIL(PLG_ERR_UNSUPPORTED, "partition starts at site %lld: it must cover the whole alignment", from);
    *range_end = to;
  }
  if (model.empty()) PLG_FAIL(PLG_ERR_IO, "Partition file could not be parsed.");
  return model;
}

// "DNA", or a protein model the engine carries (WAG, JTT)
static bool model_is_dna(const std::string &model_name) {
  if (model_name == "DNA") return true;
  const double *e, *f;
  if (!protein_model_table(model_name.c_str(), &e, &f))
    PLG_FAIL(PLG_ERR_UNSUPPORTED, "model '%s': the engine carries DNA (GTR), WAG and JTT (pll.py:102-115 passes the name "
             "through); other matrices go through plg_model_create", model_name.c_str());
  return false;
}

static uint8_t dna_code(char c) {
  switch (c >= 'a' && c <= 'z' ? c - 32 : c) {
    case 'A': return 1; case 'C': return 2; case 'G': return 4; case 'T': case 'U': return 8;
    case 'M': return 3; case 'R': return 5; case 'W': return 9; case 'S': return 6; case 'Y': return 10;
    case 'K': return 12; case 'V': return 7; case 'H': return 11; case 'D': return 13; case 'B': return 14;
    default: return 15;
  }
}
static uint8_t aa_code(char c) {
  static const char *order = "ARNDCQEGHILKMFPSTWYV";
  const char u = c >= 'a' && c <= 'z' ? c - 32 : c;
  const char *at = strchr(order, u);
  if (at && u) return (uint8_t)(at - order);
  if (u == 'B') return 20;
  if (u == 'Z') return 21;
  return 22;
}

// HostTree (bifurcating or trifurcating root, binary elsewhere) -> UTree with tips in Newick order
static UTree utree_from_newick(const HostTree &ht, std::vector<std::string> &tip_names) {
  std::vector<int> po = ht.post_order();
  std::vector<int> id(ht.nodes.size(), -1);
  int n_tips = 0;
  for (int x : po)
    if (ht.nodes[x].children.empty()) { id[x] = n_tips++; tip_names.push_back(ht.nodes[x].label); }
  if (n_tips < 3) PLG_FAIL(PLG_ERR_UNSUPPORTED, "need at least 3 taxa");
  // binarise: resolve every k-ary node left to right with zero-length branches, as a rooted descriptor
  std::vector<int32_t> desc;
  std::vector<double> br;
  for (int x : po) {
    const HostNode &nd = ht.nodes[x];
    if (nd.children.empty()) continue;
    if (nd.children.size() == 1) PLG_FAIL(PLG_ERR_UNSUPPORTED, "unary nodes are not supported in likelihood trees");
    int acc = id[nd.children[0]];
    double acc_len = ht.nodes[nd.children[0]].brlen;
    for (size_t j = 1; j < nd.children.size(); j++) {
      desc.push_back(acc);
      desc.push_back(id[nd.children[j]]);
      br.push_back(acc_len);
      br.push_back(ht.nodes[nd.children[j]].brlen);
      acc = n_tips + (int)desc.size() / 2 - 1;
      acc_len = 0.0;
    }
    id[x] = acc;
  }
  if ((int)desc.size() / 2 != n_tips - 1) PLG_FAIL(PLG_ERR_PARSE, "Tree was not able to be parsed.");
  return utree_from_desc(n_tips, desc.data(), br.data());
}

// ---------------------------------------------------------------------------------------------
// Batched getLogLikelihood: many trees on one alignment, each with the reference's treatment
// (default z = 0.9, up to `passes` smoothing passes, lnL; libpllWrapper.c:152, :205).  The smoothing
// traversal of a tree is a fixed action list -- OPT(x, from): one Newton step on the branch above
// x; CLV(x, k): re-orient x's conditional vector toward neighbour k -- whose LENGTH is the same
// for every binary tree of n tips ((2n-3) OPT + 3(n-2) CLV) although the order differs.  So B
// trees advance in lockstep: step i runs the i-th action of every tree, CLV actions as one
// clv4_kernel launch, OPT actions as sumtable -> opt4_batch_kernel (the whole makenewz step of a
// branch in one block: derivative sums, bad-curvature guard, Newton update) -> P-matrices of the
// changed branches.  Nothing comes back to the host until the scores: the per-tree path pays a
// device round trip per branch (~20 ms per 64-taxon tree); here it is ~1 ms per tree.
// One conditional vector per inner node is enough (libpll's x-pointer orientation): while
// smooth(x) runs, the parent's vector keeps pointing at x and a finished child's points back at x.
// ---------------------------------------------------------------------------------------------
struct OptItem {
  int a_id, b_id, tree, branch;  // operand ids of the branch's two ends; tree of the chunk; global branch index
};

// K = 4: the sum table is NOT stored -- the Newton step normally needs only the derivative sums that came
// with it, and the rare bad-curvature retry recomputes a pattern's 16 table entries from the two operand
// columns (2 x 16 loads instead of 16, but 13 % less traffic per tree on the common path).  K = 20 keeps the
// stored table (80 entries per pattern do not fit in registers beside the operands).
struct OptOperands {
  int n_rows;
  const unsigned char *codes;
  const unsigned int *masks;
  const double *clv;
  const int *scl;
};
template <int K>
__host__ __device__ constexpr bool store_sumtable() { return K != 4; }

template <int K, int R>
__device__ __forceinline__ void newton_step_block(const OptItem it, int item, int64_t Ppad, const double *__restrict__ weights,
                                                  const ModelDev *__restrict__ model, const double *__restrict__ S,
                                                  const int *__restrict__ Ssc, double *__restrict__ brlen,
                                                  int *__restrict__ changed, const double *__restrict__ dpart,
                                                  int blocks_per_item, const OptOperands ops);

// One makenewz step of one branch per tree, ONE launch (was: sum table, Newton step, P-matrices = three):
// every block writes its tile of the branch's sum table and its share of the derivative sums at the
// branch's CURRENT length; the block that finishes last (ticket counter) adds the shares up in a fixed
// order, takes the Newton step -- reading the table back only if the bad-curvature guard has to try
// another length --, stores the new length and recomputes the branch's P-matrix and tip table.  Trees
// that have converged (active == 0) skip the whole step.
template <int K, int R, bool FUSED>
__global__ void __launch_bounds__(SM_THREADS)
opt_step_kernel(const OptItem *__restrict__ items, int blocks_per_item, int n_rows, int64_t Ppad,
                const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
                const ModelDev *__restrict__ model, const double *__restrict__ clv,
                const int *__restrict__ scl, double *__restrict__ S, int *__restrict__ Ssc,
                const double *__restrict__ weights, double *__restrict__ brlen,
                double *__restrict__ dpart, const int *__restrict__ active, int *__restrict__ changed,
                unsigned int *__restrict__ tickets, double *__restrict__ pmat, double *__restrict__ tiptab,
                int pm_stride, int tip_stride) {
  __shared__ double tab[3][R * K];
  __shared__ double red[3][SM_THREADS / 32];
  __shared__ int s_last;
  const int item = blockIdx.x / blocks_per_item;
  const OptItem it = items[item];
  if (!active[it.tree]) return;
  if (threadIdx.x < R * K) {
    const double z = fmin(fmax(exp(-brlen[it.branch]), PLL_ZMIN), PLL_ZMAX);
    const double lz = log(fmax(z, PLL_ZMIN));
    const double g = -model->eval[threadIdx.x % K] * model->rates[threadIdx.x / K];
    const double e = exp(g * lz);
    tab[0][threadIdx.x] = e; tab[1][threadIdx.x] = g * e; tab[2][threadIdx.x] = g * g * e;
  }
  __syncthreads();
  const int r = threadIdx.x % R;
  const int64_t p = (int64_t)(blockIdx.x - item * blocks_per_item) * (SM_THREADS / R) + threadIdx.x / R;
  double A[K], B[K];
  int sc = 0;
  auto fetch = [&](int id, double(&x)[K]) {
    if (id < n_rows) {
      const unsigned int m = masks[codes[(int64_t)id * Ppad + p]];
#pragma unroll
      for (int k = 0; k < K; k++) x[k] = (m >> k & 1) ? 1.0 : 0.0;
    } else {
      const double *src = clv + ((int64_t)(id - n_rows) * R * K + r * K) * Ppad + p;
#pragma unroll
      for (int k = 0; k < K; k++) x[k] = src[(int64_t)k * Ppad];
      sc += scl[(int64_t)(id - n_rows) * Ppad + p];
    }
  };
  fetch(it.a_id, A);
  fetch(it.b_id, B);
#pragma unroll
  for (int i = 0; i < K; i++) A[i] *= model->freqs[i];
  double *So = S + (int64_t)it.tree * R * K * Ppad;
  double l0 = 0.0, l1 = 0.0, l2 = 0.0;
  for (int k = 0; k < K; k++) {
    double l = 0.0, rr = 0.0;
#pragma unroll
    for (int i = 0; i < K; i++) {
      l = fma(A[i], model->U[i * K + k], l);
      rr = fma(model->Uinv[k * K + i], B[i], rr);
    }
    const double sv = l * rr;
    if (store_sumtable<K>()) So[((int64_t)r * K + k) * Ppad + p] = sv;
    l0 = fma(sv, tab[0][r * K + k], l0);
    l1 = fma(sv, tab[1][r * K + k], l1);
    l2 = fma(sv, tab[2][r * K + k], l2);
  }
  if (store_sumtable<K>() && r == 0) Ssc[(int64_t)it.tree * Ppad + p] = sc;
#pragma unroll
  for (int d = 1; d < R; d <<= 1) {
    l0 += __shfl_xor_sync(0xffffffffu, l0, d);
    l1 += __shfl_xor_sync(0xffffffffu, l1, d);
    l2 += __shfl_xor_sync(0xffffffffu, l2, d);
  }
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  const double w = weights[p];
  if (r == 0 && w != 0.0) {
    const double inv = 1.0 / l0, q1 = l1 * inv;
    v0 = w * (log(l0 * (1.0 / R)) + sc * SM_LOG_MINLH);
    v1 = w * q1;
    v2 = w * (l2 * inv - q1 * q1);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v0 += __shfl_down_sync(0xffffffffu, v0, d);
    v1 += __shfl_down_sync(0xffffffffu, v1, d);
    v2 += __shfl_down_sync(0xffffffffu, v2, d);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = v0; red[1][threadIdx.x >> 5] = v1; red[2][threadIdx.x >> 5] = v2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
    for (int i = 0; i < SM_THREADS / 32; i++) t += red[threadIdx.x][i];
    dpart[((int64_t)item * 3 + threadIdx.x) * blocks_per_item + (blockIdx.x - item * blocks_per_item)] = t;
  }
  if constexpr (FUSED) {  // (two-launch form: newton_kernel finishes the step, and this kernel keeps its registers)
  // the last block of this branch to get here finishes the step (threadfence reduction pattern)
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned int ticket = atomicAdd(tickets + item, 1u);
    s_last = ticket == (unsigned int)blocks_per_item - 1;
    if (s_last) tickets[item] = 0;  // ready for the next step
  }
  __syncthreads();
  if (!s_last) return;
  __threadfence();
  newton_step_block<K, R>(it, item, Ppad, weights, model, S, Ssc, brlen, changed, dpart, blocks_per_item,
                          OptOperands{n_rows, codes, masks, clv, scl});
  __syncthreads();
  pmatrix_block<K, R>(model, brlen[it.branch], masks, pmat + (int64_t)it.branch * pm_stride, tiptab + (int64_t)it.branch * tip_stride);
  }
}

// DNA x Gamma-4, two-launch form: thread = pattern (128 per block), as the conditional-vector kernels.  The
// generic kernel above gives a thread one (pattern, category): 8 loads of 64-byte runs per operand and a
// block prologue (branch length -> exp -> log -> exp) in front of them; the launch list of a 256-tree
// batch had it at 1.6 TB/s against 4.6 for clv4_kernel (profiles/launches_r2w.csv).  Here every thread
// issues its column loads first, the first 16 threads make the table meanwhile, and no sum table is
// stored.  The eigensystem arrives BY VALUE (kernel parameters live in the constant bank, so U, U^-1 and pi
// are immediate operands of the FMAs: the first version read them from shared memory and executed ~900
// instructions per pattern tile for ~280 of arithmetic, issue-bound at 2.9 TB/s -- profiles/r3e_*).
#ifndef PLG_OPT4_MINB
#define PLG_OPT4_MINB 4
#endif
constexpr int OPT4_THREADS = 128;
#ifndef PLG_OPT4_TILES
#define PLG_OPT4_TILES 2
#endif
constexpr int OPT4_TILES = PLG_OPT4_TILES;  // (Ppad is a multiple of 256)
// (Eigen4: lik.h)
__global__ void __launch_bounds__(OPT4_THREADS, PLG_OPT4_MINB)
opt_step4_kernel(const OptItem *__restrict__ items, int blocks_per_item, int n_rows, int64_t Ppad,
                 const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks, const Eigen4 eg,
                 const double *__restrict__ clv, const int *__restrict__ scl, const double *__restrict__ weights,
                 const double *__restrict__ brlen, double *__restrict__ dpart, const int *__restrict__ active) {
  __shared__ double tab[3][16];
  __shared__ double red[3][OPT4_THREADS / 32];
  const int item = blockIdx.x / blocks_per_item;
  const OptItem it = items[item];
  if (!active[it.tree]) return;
  if (threadIdx.x < 16) {
    const double z = fmin(fmax(exp(-brlen[it.branch]), PLL_ZMIN), PLL_ZMAX);
    const double lz = log(fmax(z, PLL_ZMIN));
    const double g = eg.g[threadIdx.x];
    const double e = exp(g * lz);
    tab[0][threadIdx.x] = e; tab[1][threadIdx.x] = g * e; tab[2][threadIdx.x] = g * g * e;
  }
  const bool tipA = it.a_id < n_rows, tipB = it.b_id < n_rows;  // (block-uniform)
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  // OPT4_TILES tiles of 128 patterns per block: the prologue above (item, branch length, exp / log / exp) is
  // a chain of dependent global loads and transcendental calls as long as a tile's own loads
#pragma unroll 1
  for (int tl = 0; tl < OPT4_TILES; tl++) {
    const int64_t p = ((int64_t)(blockIdx.x - item * blocks_per_item) * OPT4_TILES + tl) * OPT4_THREADS + threadIdx.x;
    double a[16], b[16];
    int sc = 0;
    if (tipA) {
      const unsigned int m = masks[codes[(int64_t)it.a_id * Ppad + p]];
#pragma unroll
      for (int e = 0; e < 16; e++) a[e] = (m >> (e & 3) & 1) ? 1.0 : 0.0;
    } else {
      const double *src = clv + (int64_t)(it.a_id - n_rows) * 16 * Ppad + p;
#pragma unroll
      for (int e = 0; e < 16; e++) a[e] = __ldg(src + (int64_t)e * Ppad);
      sc += __ldg(scl + (int64_t)(it.a_id - n_rows) * Ppad + p);
    }
    if (tipB) {
      const unsigned int m = masks[codes[(int64_t)it.b_id * Ppad + p]];
#pragma unroll
      for (int e = 0; e < 16; e++) b[e] = (m >> (e & 3) & 1) ? 1.0 : 0.0;
    } else {
      const double *src = clv + (int64_t)(it.b_id - n_rows) * 16 * Ppad + p;
#pragma unroll
      for (int e = 0; e < 16; e++) b[e] = __ldg(src + (int64_t)e * Ppad);
      sc += __ldg(scl + (int64_t)(it.b_id - n_rows) * Ppad + p);
    }
    const double w = weights[p];
    if (tl == 0) __syncthreads();  // the table
    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double l = a[r * 4] * eg.piU[k], rr = eg.Uinv[k * 4] * b[r * 4];
#pragma unroll
        for (int i = 1; i < 4; i++) {
          l = fma(a[r * 4 + i], eg.piU[i * 4 + k], l);
          rr = fma(eg.Uinv[k * 4 + i], b[r * 4 + i], rr);
        }
        const double sv = l * rr;
        l0 = fma(sv, tab[0][r * 4 + k], l0);
        l1 = fma(sv, tab[1][r * 4 + k], l1);
        l2 = fma(sv, tab[2][r * 4 + k], l2);
      }
    }
    if (w != 0.0) {
      const double inv = 1.0 / l0, q1 = l1 * inv;
      v0 += w * (log(l0 * 0.25) + sc * SM_LOG_MINLH);
      v1 += w * q1;
      v2 += w * (l2 * inv - q1 * q1);
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v0 += __shfl_down_sync(0xffffffffu, v0, d);
    v1 += __shfl_down_sync(0xffffffffu, v1, d);
    v2 += __shfl_down_sync(0xffffffffu, v2, d);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = v0; red[1][threadIdx.x >> 5] = v1; red[2][threadIdx.x >> 5] = v2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < OPT4_THREADS / 32; i++) t += red[threadIdx.x][i];
    dpart[((int64_t)item * 3 + threadIdx.x) * blocks_per_item + (blockIdx.x - item * blocks_per_item)] = t;
  }
}

// 20 states x Gamma-4, two-launch form.  The sum table of a branch IS a conditional-vector update with two
// special matrices: sv[r][k] = (sum_i pi_i U[i][k] a[r][i]) * (sum_i Uinv[k][i] b[r][i]) = ((M1 a) o (M2 b))[r][k]
// with M1 = (diag(pi) U)^T and M2 = U^-1 for every category -- so it runs on the FP64 tensor pipe through
// clv20_kernel into a conditional-vector slot (one per tree), scaling counters included.  This kernel then
// folds the table with the branch's exp / g exp / g^2 exp entries: thread = pattern, 80 coalesced loads.
// (The generic (pattern, category)-per-thread kernel read U and U^-1 from memory once per FMA: 254.6 ms of
// a 128-tree batch's 483 ms, profiles/launches_r3l.csv.)
constexpr int DSUM20_THREADS = 128;
__global__ void __launch_bounds__(DSUM20_THREADS)
dsum20_kernel(const OptItem *__restrict__ items, int blocks_per_item, int64_t Ppad, const ModelDev *__restrict__ model,
              const double *__restrict__ S, const int *__restrict__ Ssc, const double *__restrict__ weights,
              const double *__restrict__ brlen, double *__restrict__ dpart, const int *__restrict__ active) {
  __shared__ double tab[3][80];
  __shared__ double red[3][DSUM20_THREADS / 32];
  const int item = blockIdx.x / blocks_per_item;
  const OptItem it = items[item];
  if (!active[it.tree]) return;
  if (threadIdx.x < 80) {
    const double z = fmin(fmax(exp(-brlen[it.branch]), PLL_ZMIN), PLL_ZMAX);
    const double lz = log(fmax(z, PLL_ZMIN));
    const double g = -model->eval[threadIdx.x % 20] * model->rates[threadIdx.x / 20];
    const double e = exp(g * lz);
    tab[0][threadIdx.x] = e; tab[1][threadIdx.x] = g * e; tab[2][threadIdx.x] = g * g * e;
  }
  const int64_t p = (int64_t)(blockIdx.x - item * blocks_per_item) * DSUM20_THREADS + threadIdx.x;
  const double *Sb = S + (int64_t)it.tree * 80 * Ppad + p;
  double sv[80];
#pragma unroll
  for (int e = 0; e < 80; e++) sv[e] = __ldg(Sb + (int64_t)e * Ppad);
  const int sc = __ldg(Ssc + (int64_t)it.tree * Ppad + p);
  const double w = weights[p];
  __syncthreads();
  double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
  for (int e = 0; e < 80; e++) {
    l0 = fma(sv[e], tab[0][e], l0);
    l1 = fma(sv[e], tab[1][e], l1);
    l2 = fma(sv[e], tab[2][e], l2);
  }
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
  if (w != 0.0) {
    const double inv = 1.0 / l0, q1 = l1 * inv;
    v0 = w * (log(l0 * 0.25) + sc * SM_LOG_MINLH);
    v1 = w * q1;
    v2 = w * (l2 * inv - q1 * q1);
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v0 += __shfl_down_sync(0xffffffffu, v0, d);
    v1 += __shfl_down_sync(0xffffffffu, v1, d);
    v2 += __shfl_down_sync(0xffffffffu, v2, d);
  }
  if ((threadIdx.x & 31) == 0) {
    red[0][threadIdx.x >> 5] = v0; red[1][threadIdx.x >> 5] = v1; red[2][threadIdx.x >> 5] = v2;
  }
  __syncthreads();
  if (threadIdx.x < 3) {
    double t = 0.0;
#pragma unroll
    for (int i = 0; i < DSUM20_THREADS / 32; i++) t += red[threadIdx.x][i];
    dpart[((int64_t)item * 3 + threadIdx.x) * blocks_per_item + (blockIdx.x - item * blocks_per_item)] = t;
  }
}

// Second launch of the two-launch form (large batches: one block per tree finishes its branch's step).
template <int K, int R>
__global__ void __launch_bounds__(SM_THREADS)
newton_kernel(const OptItem *__restrict__ items, int blocks_per_item, int64_t Ppad, const unsigned int *__restrict__ masks,
              const ModelDev *__restrict__ model, const double *__restrict__ S, const int *__restrict__ Ssc,
              const double *__restrict__ weights, double *__restrict__ brlen, const double *__restrict__ dpart,
              const int *__restrict__ active, int *__restrict__ changed, double *__restrict__ pmat,
              double *__restrict__ tiptab, int pm_stride, int tip_stride, const OptOperands ops) {
  const OptItem it = items[blockIdx.x];
  if (!active[it.tree]) return;
  newton_step_block<K, R>(it, blockIdx.x, Ppad, weights, model, S, Ssc, brlen, changed, dpart, blocks_per_item, ops);
  __syncthreads();
  pmatrix_block<K, R>(model, brlen[it.branch], masks, pmat + (int64_t)it.branch * pm_stride, tiptab + (int64_t)it.branch * tip_stride);
}

// One makenewz step (optimise_branch with maxiter = 1) by one block, on the tree's sum table.
template <int K, int R>
__device__ __forceinline__ void newton_step_block(const OptItem it, int item, int64_t Ppad, const double *__restrict__ weights,
                                                  const ModelDev *__restrict__ model, const double *__restrict__ S,
                                                  const int *__restrict__ Ssc, double *__restrict__ brlen,
                                                  int *__restrict__ changed, const double *__restrict__ dpart,
                                                  int blocks_per_item, const OptOperands ops) {
  constexpr int RK = R * K;
  __shared__ double tab[3][RK];
  __shared__ double red[3][SM_THREADS / 32];
  __shared__ double sh_lz;
  __shared__ int sh_go;
  const double *Sb = S + (int64_t)it.tree * RK * Ppad;
  const int *sc = Ssc + (int64_t)it.tree * Ppad;
  double z0 = 0.0, z = 0.0, zprev = 0.0, d1 = 0.0, d2 = 0.0;
  if (threadIdx.x == 0) {
    z0 = exp(-brlen[it.branch]);
    z = fmin(fmax(z0, PLL_ZMIN), PLL_ZMAX);
    zprev = z;
  }
  for (int guard = 0; guard < 64; guard++) {  // "bad curvature: shorten the branch" loop
    if (threadIdx.x == 0) sh_lz = log(fmax(z, PLL_ZMIN));
    __syncthreads();
    if (threadIdx.x < RK) {
      const int r = threadIdx.x / K, k = threadIdx.x % K;
      const double g = -model->eval[k] * model->rates[r];  // t = -lz: d/dlz exp(eval*rate*t) = g * exp(.)
      const double e = exp(g * sh_lz);
      tab[0][threadIdx.x] = e; tab[1][threadIdx.x] = g * e; tab[2][threadIdx.x] = g * g * e;
    }
    __syncthreads();
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    if (guard == 0) {  // the sums at the current length came with the table
      const double *dp = dpart + (int64_t)item * 3 * blocks_per_item;
      for (int b = threadIdx.x; b < blocks_per_item; b += SM_THREADS) {
        v0 += dp[b]; v1 += dp[blocks_per_item + b]; v2 += dp[2 * blocks_per_item + b];
      }
    }
    for (int64_t p = threadIdx.x; guard > 0 && p < Ppad; p += SM_THREADS) {
      const double w = weights[p];
      if (w == 0.0) continue;
      double l0 = 0.0, l1 = 0.0, l2 = 0.0;
      int scp = 0;
      if constexpr (store_sumtable<K>()) {
#pragma unroll
        for (int e = 0; e < RK; e++) {
          const double sv = Sb[(int64_t)e * Ppad + p];
          l0 = fma(sv, tab[0][e], l0);
          l1 = fma(sv, tab[1][e], l1);
          l2 = fma(sv, tab[2][e], l2);
        }
        scp = sc[p];
      } else {  // the pattern's table entries again, from the two operand columns (as opt_step_kernel makes them)
        double A[RK], B[RK];
        auto fetch = [&](int id, double(&x)[RK]) {
          if (id < ops.n_rows) {
            const unsigned int m = ops.masks[ops.codes[(int64_t)id * Ppad + p]];
#pragma unroll
            for (int e = 0; e < RK; e++) x[e] = (m >> (e % K) & 1) ? 1.0 : 0.0;
          } else {
            const double *src = ops.clv + (int64_t)(id - ops.n_rows) * RK * Ppad + p;
#pragma unroll
            for (int e = 0; e < RK; e++) x[e] = src[(int64_t)e * Ppad];
            scp += ops.scl[(int64_t)(id - ops.n_rows) * Ppad + p];
          }
        };
        fetch(it.a_id, A);
        fetch(it.b_id, B);
#pragma unroll
        for (int r = 0; r < R; r++)
#pragma unroll
          for (int k = 0; k < K; k++) {
            double l = 0.0, rr = 0.0;
#pragma unroll
            for (int i = 0; i < K; i++) {
              l = fma(A[r * K + i] * model->freqs[i], model->U[i * K + k], l);
              rr = fma(model->Uinv[k * K + i], B[r * K + i], rr);
            }
            const double sv = l * rr;
            l0 = fma(sv, tab[0][r * K + k], l0);
            l1 = fma(sv, tab[1][r * K + k], l1);
            l2 = fma(sv, tab[2][r * K + k], l2);
          }
      }
      const double inv = 1.0 / l0, q1 = l1 * inv;
      v0 += w * (log(l0 * (1.0 / R)) + scp * SM_LOG_MINLH);
      v1 += w * q1;
      v2 += w * (l2 * inv - q1 * q1);
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      v0 += __shfl_down_sync(0xffffffffu, v0, d);
      v1 += __shfl_down_sync(0xffffffffu, v1, d);
      v2 += __shfl_down_sync(0xffffffffu, v2, d);
    }
    if ((threadIdx.x & 31) == 0) {
      red[0][threadIdx.x >> 5] = v0; red[1][threadIdx.x >> 5] = v1; red[2][threadIdx.x >> 5] = v2;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      d1 = d2 = 0.0;
      for (int i = 0; i < SM_THREADS / 32; i++) { d1 += red[1][i]; d2 += red[2][i]; }  // fixed order
      if (d2 >= 0.0 && z < PLL_ZMAX) { zprev = z = 0.37 * z + 0.63; sh_go = 1; }
      else sh_go = 0;
    }
    __syncthreads();
    if (!sh_go) break;
  }
  if (threadIdx.x == 0) {
    if (d2 < 0.0) {
      const double tantmp = -d1 / d2;
      if (tantmp < 100.0) {
        z *= exp(tantmp);
        if (z < PLL_ZMIN) z = PLL_ZMIN;
        if (z > 0.25 * zprev + 0.75) z = 0.25 * zprev + 0.75;
      } else {
        z = 0.25 * zprev + 0.75;
      }
    }
    if (z > PLL_ZMAX) z = PLL_ZMAX;
    brlen[it.branch] = -log(z);
    if (fabs(z - z0) > PLL_DELTAZ) changed[it.tree] = 1;
    __threadfence_block();
  }
}

__global__ void next_pass_kernel(int n, int *__restrict__ active, int *__restrict__ changed) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) { active[t] = active[t] && changed[t]; changed[t] = 0; }  // smoothTree stops once a pass moved nothing
}

// Per tree: the action lists of the down pass and of one smoothing pass (the same for every pass).
struct TreePlan {
  std::vector<LikOp> down;    // n - 2 updates, leaves first
  struct Act { int is_opt; LikOp op; OptItem it; };
  std::vector<Act> acts;      // (2n - 3) OPT + 3(n - 2) CLV, smooth() order
  LikEval eval;
};

static TreePlan plan_tree(const UTree &t, const std::vector<int32_t> &tip_rows, int n_rows, int tree_idx,
                          int64_t slot0, int64_t branch0) {
  TreePlan out;
  const int n = t.n_tips;
  // branch ids: undirected edges numbered by their first appearance (x, k) with x < nb[x][k]
  std::vector<int> branch_of((size_t)t.n_nodes * 3, -1);
  int nb_count = 0;
  for (int x = 0; x < t.n_nodes; x++)
    for (int k = 0; k < 3; k++) {
      const int y = t.nb[x][k];
      if (y > x) { branch_of[3 * x + k] = nb_count; branch_of[3 * y + t.slot_of(y, x)] = nb_count; nb_count++; }
    }
  auto operand = [&](int x) { return x < n ? (int)tip_rows[x] : (int)(n_rows + slot0 + (x - n)); };
  auto pm = [&](int x, int k) { return (int)(branch0 + branch_of[3 * x + k]); };
  auto clv_op = [&](int x, int k) {  // x's vector re-oriented toward nb[x][k]
    LikOp f;
    f.dst = operand(x);
    int s = 0;
    for (int j = 0; j < 3; j++)
      if (j != k) {
        const int c = t.nb[x][j];
        if (s == 0) { f.a = operand(c); f.pa = pm(x, j); } else { f.b = operand(c); f.pb = pm(x, j); }
        s++;
      }
    return f;
  };
  // down pass toward tip 0
  std::vector<int> par(t.n_nodes, -1), bfs{0};
  for (size_t i = 0; i < bfs.size(); i++)
    for (int k = 0; k < 3; k++) {
      const int y = t.nb[bfs[i]][k];
      if (y >= 0 && y != par[bfs[i]]) { par[y] = bfs[i]; bfs.push_back(y); }
    }
  for (size_t i = bfs.size(); i-- > 1;)
    if (!t.is_tip(bfs[i])) out.down.push_back(clv_op(bfs[i], t.slot_of(bfs[i], par[bfs[i]])));
  // smooth(nb[0][0], 0), iteratively
  struct Fr { int x, from, state; };
  std::vector<Fr> st{{t.nb[0][0], 0, 0}};
  while (!st.empty()) {
    Fr &f = st.back();
    const int x = f.x, from = f.from, kf = t.slot_of(x, from);
    if (f.state == 0) {
      TreePlan::Act a;
      a.is_opt = 1;
      a.it.a_id = operand(x); a.it.b_id = operand(from); a.it.tree = tree_idx; a.it.branch = pm(x, kf);
      out.acts.push_back(a);
      if (t.is_tip(x)) { st.pop_back(); continue; }
      f.state = 1;
      continue;
    }
    if (f.state <= 2) {
      int seen = 0, j = 0;
      for (; j < 3; j++)
        if (j != kf && ++seen == f.state) break;
      TreePlan::Act a;
      a.is_opt = 0;
      a.op = clv_op(x, j);
      out.acts.push_back(a);
      f.state++;
      st.push_back({t.nb[x][j], x, 0});  // (invalidates f)
      continue;
    }
    TreePlan::Act a;
    a.is_opt = 0;
    a.op = clv_op(x, kf);
    out.acts.push_back(a);
    st.pop_back();
  }
  const int w = t.nb[0][0];
  out.eval.a = tip_rows[0]; out.eval.pa = -1;
  out.eval.b = operand(w); out.eval.pb = pm(0, 0);
  out.eval.c = -1; out.eval.pc = -1; out.eval.tree = tree_idx; out.eval.pad = 0;
  return out;
}

// lnL and optimised branch lengths (per tree: len[x][k] as in the UTree) of trees [0, n_trees)
// the two model families with batch kernels: how their level kernels are launched and laid out
struct Dna4Family {
  static constexpr int K = 4, R = 4, PM_STRIDE = 68, TIP_STRIDE = 256, EVAL_TILE = 128;
  using Op = LikOpF;
  static Op make(const LikOp &o) {
    Op f;
    f.dst = o.dst; f.a = o.a; f.b = o.b; f.pa = o.pa; f.pb = o.pb;
    f.tree = -1; f.e_pa = f.e_b = f.e_pb = f.e_c = f.e_pc = -1; f.pad = 0;
    return f;
  }
  static void pmatrices(const LikAln &a, const Model &m, const double *bl, const int *idx, int64_t n, double *pm, double *tt,
                        cudaStream_t st) { dna4_pmatrices(a, m, bl, idx, n, pm, tt, st); }
  static void updates(const LikAln &a, const Model &m, const Op *ops, int64_t n, const double *pm, const double *tt,
                      double *clv, int *scl, cudaStream_t st) { dna4_updates(a, m, ops, n, pm, tt, clv, scl, st); }
  static void evals(const LikAln &a, const Model &m, const LikEval *ev, int64_t n, const double *pm, const double *tt,
                    const double *clv, const int *scl, double *partial, double *lnl, cudaStream_t st) {
    dna4_evals(a, m, ev, n, pm, tt, clv, scl, partial, lnl, st);
  }
};
struct Aa20Family {
  static constexpr int K = 20, R = 4, PM_STRIDE = 4 * 401, TIP_STRIDE = 23 * 80, EVAL_TILE = 64;
  using Op = LikOp;
  static Op make(const LikOp &o) { return o; }
  static void pmatrices(const LikAln &a, const Model &m, const double *bl, const int *idx, int64_t n, double *pm, double *tt,
                        cudaStream_t st) { aa20_pmatrices(a, m, bl, idx, n, pm, tt, st); }
  static void updates(const LikAln &a, const Model &m, const Op *ops, int64_t n, const double *pm, const double *,
                      double *clv, int *scl, cudaStream_t st) { aa20_updates(a, m, ops, n, pm, clv, scl, st); }
  static void evals(const LikAln &a, const Model &m, const LikEval *ev, int64_t n, const double *pm, const double *,
                    const double *clv, const int *scl, double *partial, double *lnl, cudaStream_t st) {
    aa20_evals(a, m, ev, n, pm, clv, scl, partial, lnl, st);
  }
};

// Device buffers of one batch.  A caller that scores one tree at a time (the per-handle drop-in path,
// plg_pll_loglik) keeps a Scratch alive between calls so that nothing is allocated per tree.
template <typename F>
struct BatchScratch {
  DevBuf<double> clv, S, pmat, tiptab, brlen, partial, lnl, dpart;
  DevBuf<int> scl, Ssc, idx, flags;
  DevBuf<unsigned int> tickets;
  DevBuf<typename F::Op> d_ops;
  DevBuf<OptItem> d_items;
  DevBuf<LikEval> d_evals;
  DevBuf<ClvOptOp> d_fops;
  DevBuf<LikOp> d_sops;  // 20 states: the Newton steps' sum tables as conditional-vector updates
};

template <typename F>
static void pll_loglik_batch(const LikAln &aln, const Model &model, std::vector<UTree> &trees,
                             const std::vector<std::vector<int32_t>> &tip_rows, int passes, double *out_lnl,
                             BatchScratch<F> *keep = nullptr) {
  constexpr int K = F::K, R = F::R, RK = K * R;
  using Op = typename F::Op;
  const int64_t n_trees = (int64_t)trees.size();
  if (n_trees == 0) return;
  const int n = trees[0].n_tips, n_inner = n - 2, n_br = 2 * n - 3;
  const int64_t Ppad = aln.Ppad;
  size_t free_b = 0, total_b = 0;
  PLG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  const size_t per_tree = ((size_t)n_inner * (RK * 8 + 4) + RK * 8 + 4) * Ppad + (size_t)n_br * (F::PM_STRIDE + F::TIP_STRIDE) * 8;
  int64_t B = std::max<int64_t>(1, std::min<int64_t>(n_trees, (int64_t)(std::min<size_t>(free_b / 2, (size_t)64 << 30) / per_tree)));
  if (const char *e = getenv("PLG_PLL_BATCH_CHUNK")) B = std::max<int64_t>(1, std::min<int64_t>(B, atoll(e)));  // tests: several chunks
  BatchScratch<F> local;
  BatchScratch<F> &sc_ = keep ? *keep : local;
  DevBuf<double> &clv = sc_.clv, &S = sc_.S, &pmat = sc_.pmat, &tiptab = sc_.tiptab, &brlen = sc_.brlen, &partial = sc_.partial,
                 &lnl = sc_.lnl, &dpart = sc_.dpart;
  DevBuf<int> &scl = sc_.scl, &Ssc = sc_.Ssc, &idx = sc_.idx, &flags = sc_.flags;
  DevBuf<unsigned int> &tickets = sc_.tickets;
  DevBuf<Op> &d_ops = sc_.d_ops;
  DevBuf<OptItem> &d_items = sc_.d_items;
  DevBuf<LikEval> &d_evals = sc_.d_evals;
  DevBuf<ClvOptOp> &d_fops = sc_.d_fops;
  DevBuf<LikOp> &d_sops = sc_.d_sops;
  // K = 20: the sum tables are conditional-vector slots B * n_inner + tree of the same buffers (clv20_kernel
  // writes them), and two more P-matrix entries hold M1 = (diag(pi) U)^T, M2 = U^-1 (dsum20_kernel above)
  constexpr bool S_IN_CLV = K == 20 && R == 4;
  clv.alloc((size_t)B * (n_inner + (S_IN_CLV ? 1 : 0)) * RK * Ppad);
  scl.alloc((size_t)B * (n_inner + (S_IN_CLV ? 1 : 0)) * Ppad);
  S.alloc(store_sumtable<K>() && !S_IN_CLV ? (size_t)B * RK * Ppad : 1);
  Ssc.alloc(store_sumtable<K>() && !S_IN_CLV ? (size_t)B * Ppad : 1);
  double *const Sp = S_IN_CLV ? clv.p + (size_t)B * n_inner * RK * Ppad : S.p;
  int *const Sscp = S_IN_CLV ? scl.p + (size_t)B * n_inner * Ppad : Ssc.p;
  pmat.alloc(((size_t)B * n_br + 2) * F::PM_STRIDE);
  const int PM_M1 = (int)(B * n_br), PM_M2 = PM_M1 + 1;
  if (S_IN_CLV) {
    std::vector<double> hm((size_t)2 * F::PM_STRIDE, 0.0);
    for (int r = 0; r < R; r++)
      for (int k = 0; k < K; k++)
        for (int i = 0; i < K; i++) {
          hm[(size_t)r * (K * K + 1) + k * K + i] = model.h.freqs[i] * model.h.U[i * K + k];
          hm[(size_t)F::PM_STRIDE + (size_t)r * (K * K + 1) + k * K + i] = model.h.Uinv[k * K + i];
        }
    PLG_CUDA(cudaMemcpy(pmat.p + (size_t)PM_M1 * F::PM_STRIDE, hm.data(), hm.size() * 8, cudaMemcpyHostToDevice));
  }
  tiptab.alloc((size_t)B * n_br * F::TIP_STRIDE);
  brlen.alloc((size_t)B * n_br);
  const int64_t ebpo = Ppad / F::EVAL_TILE;
  partial.alloc((size_t)B * std::max<int64_t>(1, ebpo) * 2);
  lnl.alloc((size_t)B);
  flags.alloc((size_t)2 * B);
  d_evals.alloc((size_t)B);
  const int n_steps = n_br + 3 * n_inner;
  d_ops.alloc((size_t)B * (n_inner + 3 * n_inner));
  d_items.alloc((size_t)B * n_br);
  idx.alloc((size_t)B * n_br);
  const int bpi = (int)(Ppad / (SM_THREADS / R));
  // blocks per branch of the launch that leaves the derivative-sum shares (DNA, large batches: opt_step4_kernel)
  const int bpi2 = (K == 4 && R == 4) ? (int)(Ppad / (OPT4_THREADS * OPT4_TILES)) : (S_IN_CLV ? (int)(Ppad / DSUM20_THREADS) : bpi);
  Eigen4 eigen4;
  if (K == 4 && R == 4)
    for (int e = 0; e < 16; e++) {
      eigen4.piU[e] = model.h.freqs[e >> 2] * model.h.U[e];  // [i][k]: pi_i U[i][k]
      eigen4.Uinv[e] = model.h.Uinv[e];
      eigen4.g[e] = -model.h.eval[e & 3] * model.h.rates[e >> 2];
    }
  dpart.alloc((size_t)2 * B * 3 * bpi);  // two halves, alternating by step (a fused re-orientation writes the NEXT step's)
  tickets.alloc((size_t)B);
  PLG_CUDA(cudaMemsetAsync(tickets.p, 0, (size_t)B * sizeof(unsigned int), 0));
  std::vector<Op> h_ops;
  std::vector<OptItem> h_items;
  std::vector<int> h_idx, h_flags;
  std::vector<LikEval> h_evals;
  std::vector<double> h_brlen, h_lnl;
  StageTimer tm;
  for (int64_t t0 = 0; t0 < n_trees; t0 += B) {
    const int64_t cnt = std::min(B, n_trees - t0);
    std::vector<TreePlan> plans(cnt);
    for (int64_t i = 0; i < cnt; i++)  // (checked here: the pool's workers must not throw)
      if (trees[t0 + i].n_tips != n) PLG_FAIL(PLG_ERR_ARG, "trees of a batch must have the same number of taxa");
    HostPool::get().run((int)std::min<int64_t>(HostPool::get().size(), std::max<int64_t>(1, cnt / 8)), [&](int ti, int nt) {
      for (int64_t i = cnt * ti / nt; i < cnt * (ti + 1) / nt; i++)
        plans[i] = plan_tree(trees[t0 + i], tip_rows[t0 + i], aln.n_rows, (int)i, i * n_inner, i * n_br);
    });
    // step-major layout: down-pass steps first, then the steps of one smoothing pass
    h_ops.clear(); h_items.clear(); h_idx.clear(); h_evals.clear();
    std::vector<int64_t> down_off(n_inner + 1, 0), clv_off(n_steps + 1, 0), opt_off(n_steps + 1, 0);
    for (int j = 0; j < n_inner; j++) {
      for (int64_t i = 0; i < cnt; i++) h_ops.push_back(F::make(plans[i].down[j]));
      down_off[j + 1] = (int64_t)h_ops.size();
    }
    const int64_t pass_ops0 = (int64_t)h_ops.size();
    // DNA, wide batches: a re-orientation whose result is the b end of the branch the tree's NEXT action
    // optimises runs fused with that action's sum-table pass (lik.h ClvOptOp); such items come first in
    // their step's list and skip the sum-table launch
    const bool can_fuse = K == 4 && R == 4 && cnt > 4 && !getenv("PLG_PLL_NO_FUSE");
    auto fuses = [&](int64_t i, int s) {
      if (!can_fuse || s < 0 || s + 1 >= n_steps) return false;
      const TreePlan::Act &a = plans[i].acts[s], &nx = plans[i].acts[s + 1];
      return !a.is_opt && nx.is_opt && nx.it.b_id == a.op.dst;
    };
    std::vector<ClvOptOp> h_fops;
    std::vector<int64_t> fop_off(n_steps + 1, 0), n_pre(n_steps, 0);
    for (int s = 0; s < n_steps; s++) {
      for (int64_t i = 0; i < cnt; i++)  // items whose sums the previous step's fused launch leaves
        if (plans[i].acts[s].is_opt && fuses(i, s - 1)) { h_items.push_back(plans[i].acts[s].it); h_idx.push_back(plans[i].acts[s].it.branch); }
      n_pre[s] = (int64_t)h_items.size() - opt_off[s];
      int64_t slot = 0;
      for (int64_t i = 0; i < cnt; i++) {
        const TreePlan::Act &a = plans[i].acts[s];
        if (a.is_opt) {
          if (!fuses(i, s - 1)) { h_items.push_back(a.it); h_idx.push_back(a.it.branch); }
        } else if (fuses(i, s)) {
          const OptItem &it = plans[i].acts[s + 1].it;
          h_fops.push_back({a.op.dst, a.op.a, a.op.b, a.op.pa, a.op.pb, it.a_id, it.branch, it.tree, (int)slot++, 0, 0, 0});
        } else {
          h_ops.push_back(F::make(a.op));
        }
      }
      clv_off[s + 1] = (int64_t)h_ops.size() - pass_ops0;
      opt_off[s + 1] = (int64_t)h_items.size();
      fop_off[s + 1] = (int64_t)h_fops.size();
    }
    std::vector<LikOp> h_sops;
    if (S_IN_CLV) {  // (item order: dsum20_kernel and newton_kernel index the tables by tree)
      h_sops.reserve(h_items.size());
      for (const OptItem &it : h_items)
        h_sops.push_back({(int)(aln.n_rows + B * n_inner + it.tree), it.a_id, it.b_id, PM_M1, PM_M2});
    }
    for (int64_t i = 0; i < cnt; i++) h_evals.push_back(plans[i].eval);
    // start from the trees' current lengths (callers load default z = 0.9: useDefaultz, libpllWrapper.c:152;
    // a handle scored twice continues from where the first call left its branches, like the reference)
    h_brlen.assign((size_t)cnt * n_br, 0.0);
    for (int64_t i = 0; i < cnt; i++) {
      const UTree &t = trees[t0 + i];
      int b = 0;
      for (int x = 0; x < t.n_nodes; x++)  // same numbering as plan_tree
        for (int k = 0; k < 3; k++)
          if (t.nb[x][k] > x) h_brlen[(size_t)i * n_br + b++] = t.len[x][k];
    }
    h_flags.assign((size_t)2 * cnt, 0);
    for (int64_t i = 0; i < cnt; i++) h_flags[i] = 1;  // active; changed = 0
    tm.lap("  smooth plan");
    PLG_CUDA(cudaMemcpy(d_ops.p, h_ops.data(), h_ops.size() * sizeof(Op), cudaMemcpyHostToDevice));
    PLG_CUDA(cudaMemcpy(d_items.p, h_items.data(), h_items.size() * sizeof(OptItem), cudaMemcpyHostToDevice));
    d_sops.alloc(std::max<size_t>(1, h_sops.size()));
    if (!h_sops.empty()) PLG_CUDA(cudaMemcpy(d_sops.p, h_sops.data(), h_sops.size() * sizeof(LikOp), cudaMemcpyHostToDevice));
    d_fops.alloc(std::max<size_t>(1, h_fops.size()));
    if (!h_fops.empty()) PLG_CUDA(cudaMemcpy(d_fops.p, h_fops.data(), h_fops.size() * sizeof(ClvOptOp), cudaMemcpyHostToDevice));
    PLG_CUDA(cudaMemcpy(idx.p, h_idx.data(), h_idx.size() * sizeof(int), cudaMemcpyHostToDevice));
    PLG_CUDA(cudaMemcpy(d_evals.p, h_evals.data(), h_evals.size() * sizeof(LikEval), cudaMemcpyHostToDevice));
    PLG_CUDA(cudaMemcpy(brlen.p, h_brlen.data(), h_brlen.size() * 8, cudaMemcpyHostToDevice));
    PLG_CUDA(cudaMemcpy(flags.p, h_flags.data(), h_flags.size() * sizeof(int), cudaMemcpyHostToDevice));
    int *active = flags.p, *changed = flags.p + cnt;
    cudaStream_t st = 0;
    // (20 states: the tensor-pipe sum table + dsum20_kernel + newton_kernel beat the generic fused kernel even for
    // one tree: 36 us per Newton step against ~23)
    const bool fused = cnt <= 4 && !S_IN_CLV;
    tm.lap("  smooth upload");
    F::pmatrices(aln, model, brlen.p, nullptr, cnt * n_br, pmat.p, tiptab.p, st);
    for (int j = 0; j < n_inner; j++)
      F::updates(aln, model, d_ops.p + down_off[j], down_off[j + 1] - down_off[j], pmat.p, tiptab.p, clv.p, scl.p, st);
    for (int pass = 0; pass < passes; pass++) {
      for (int s = 0; s < n_steps; s++) {
        const int64_t nc = clv_off[s + 1] - clv_off[s], no = opt_off[s + 1] - opt_off[s], nf = fop_off[s + 1] - fop_off[s];
        double *dp_cur = dpart.p + (size_t)(s & 1) * B * 3 * bpi, *dp_next = dpart.p + (size_t)((s + 1) & 1) * B * 3 * bpi;
        if (nc) F::updates(aln, model, d_ops.p + pass_ops0 + clv_off[s], nc, pmat.p, tiptab.p, clv.p, scl.p, st);
        if constexpr (K == 4 && R == 4)
          if (nf) dna4_updates_opt(aln, model, d_fops.p + fop_off[s], nf, pmat.p, tiptab.p, clv.p, scl.p, eigen4, brlen.p, active, dp_next, st);
        if (no) {
          const OptItem *its = d_items.p + opt_off[s];
          const int64_t npre = n_pre[s];  // their derivative-sum shares are already in dp_cur
          // a few trees: one launch per step (the last block of a branch finishes it: launch-bound regime);
          // many trees: two launches, no fences or tickets in the wide one
          if (fused)
            opt_step_kernel<K, R, true><<<(unsigned)(no * bpi), SM_THREADS, 0, st>>>(
                its, bpi, aln.n_rows, Ppad, aln.codes.p, aln.masks.p, model.d.p, clv.p, scl.p, Sp, Sscp,
                aln.weights.p, brlen.p, dp_cur, active, changed, tickets.p, pmat.p, tiptab.p, F::PM_STRIDE, F::TIP_STRIDE);
          else if constexpr (K == 4 && R == 4) {
            if (no > npre)
              opt_step4_kernel<<<(unsigned)((no - npre) * bpi2), OPT4_THREADS, 0, st>>>(
                  its + npre, bpi2, aln.n_rows, Ppad, aln.codes.p, aln.masks.p, eigen4, clv.p, scl.p, aln.weights.p, brlen.p,
                  dp_cur + (size_t)npre * 3 * bpi2, active);
          }
          else if constexpr (S_IN_CLV) {
            aa20_updates(aln, model, d_sops.p + opt_off[s], no, pmat.p, clv.p, scl.p, st);
            dsum20_kernel<<<(unsigned)(no * bpi2), DSUM20_THREADS, 0, st>>>(its, bpi2, Ppad, model.d.p, Sp, Sscp, aln.weights.p,
                                                                          brlen.p, dp_cur, active);
          } else
            opt_step_kernel<K, R, false><<<(unsigned)(no * bpi), SM_THREADS, 0, st>>>(
                its, bpi, aln.n_rows, Ppad, aln.codes.p, aln.masks.p, model.d.p, clv.p, scl.p, Sp, Sscp,
                aln.weights.p, brlen.p, dp_cur, active, changed, nullptr, pmat.p, tiptab.p, F::PM_STRIDE, F::TIP_STRIDE);
          count_launch();
          if (!fused) {
            newton_kernel<K, R><<<(unsigned)no, SM_THREADS, 0, st>>>(its, bpi2, Ppad, aln.masks.p, model.d.p, Sp, Sscp,
                                                                      aln.weights.p, brlen.p, dp_cur, active, changed,
                                                                      pmat.p, tiptab.p, F::PM_STRIDE, F::TIP_STRIDE,
                                                                      OptOperands{aln.n_rows, aln.codes.p, aln.masks.p, clv.p, scl.p});
            count_launch();
          }
        }
      }
      next_pass_kernel<<<(unsigned)((cnt + 255) / 256), 256, 0, st>>>((int)cnt, active, changed);
      count_launch();
    }
    F::evals(aln, model, d_evals.p, cnt, pmat.p, tiptab.p, clv.p, scl.p, partial.p, lnl.p, st);
    PLG_CUDA(cudaGetLastError());
    tm.lap("  smooth enqueue");
    h_lnl.resize(cnt);
    PLG_CUDA(cudaMemcpy(h_lnl.data(), lnl.p, (size_t)cnt * 8, cudaMemcpyDeviceToHost));
    PLG_CUDA(cudaMemcpy(h_brlen.data(), brlen.p, (size_t)cnt * n_br * 8, cudaMemcpyDeviceToHost));
    tm.lap("  smooth wait+D2H");
    for (int64_t i = 0; i < cnt; i++) {
      out_lnl[t0 + i] = h_lnl[i];
      UTree &t = trees[t0 + i];
      int b = 0;
      for (int x = 0; x < t.n_nodes; x++)  // same numbering as plan_tree
        for (int k = 0; k < 3; k++) {
          const int y = t.nb[x][k];
          if (y > x) { const double len = h_brlen[(size_t)i * n_br + b++]; t.len[x][k] = len; t.len[y][t.slot_of(y, x)] = len; }
        }
    }
  }
}

}  // namespace plg
#include "pll_smooth1.cuh"
namespace plg {

// DNA x Gamma-4, one tree: the whole smoothing program in one cooperative launch (pll_smooth1.cuh).
// Returns false when the device cannot launch cooperatively: the caller takes the lockstep program instead.
static bool smooth1(const LikAln &aln, const Model &model, UTree &tree, const std::vector<int32_t> &tip_rows,
                    int max_passes, double *out_lnl) {
  struct Scratch {
    DevBuf<double> clv, pmat, tiptab, brlen, out, spill;
    DevBuf<int> scl;
    DevBuf<unsigned long long> ll;
    DevBuf<S1Act> acts;
    int sms = -1, coop = 0;
  };
  static thread_local Scratch keep[8];
  int dev = 0;
  cudaGetDevice(&dev);
  Scratch &sc = keep[dev & 7];
  constexpr size_t S1_DYN_MAX = (size_t)176 << 10;
  if (sc.sms < 0) {
    PLG_CUDA(cudaFuncSetAttribute(smooth1_kernel<false, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S1_DYN_MAX));
    PLG_CUDA(cudaFuncSetAttribute(smooth1_kernel<false, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S1_DYN_MAX));
    PLG_CUDA(cudaFuncSetAttribute(smooth1_kernel<true, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S1_DYN_MAX));
    PLG_CUDA(cudaFuncSetAttribute(smooth1_kernel<true, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)S1_DYN_MAX));
    PLG_CUDA(cudaDeviceGetAttribute(&sc.sms, cudaDevAttrMultiProcessorCount, dev));
    PLG_CUDA(cudaDeviceGetAttribute(&sc.coop, cudaDevAttrCooperativeLaunch, dev));
  }
  const int64_t Ppad = aln.Ppad;
  if (!sc.coop) return false;
  const int n = tree.n_tips, n_inner = n - 2, n_br = 2 * n - 3;
  TreePlan plan = plan_tree(tree, tip_rows, aln.n_rows, 0, 0, 0);
  std::vector<S1Act> acts;
  acts.reserve(plan.down.size() + plan.acts.size() + 1);
  for (const LikOp &o : plan.down) acts.push_back({0, o.dst, o.a, o.b, o.pa, o.pb, -1, 0});
  for (const TreePlan::Act &a : plan.acts) {
    if (a.is_opt) acts.push_back({1, -1, a.it.a_id, a.it.b_id, -1, -1, a.it.branch, 0});
    else acts.push_back({0, a.op.dst, a.op.a, a.op.b, a.op.pa, a.op.pb, -1, 0});
  }
  acts.push_back({1, -1, plan.eval.a, plan.eval.b, -1, -1, plan.eval.pb, 0});  // lnL across the branch at tip 0
  // shared memory of a block: every branch's length and P-matrix -- as many matrices as fit (all of them up to 174
  // taxa; the others go to the block's own global region) -- and the action list when it still fits
  const size_t len_bytes = (size_t)((n_br + 1) & ~1) * 8, acts_bytes = acts.size() * sizeof(S1Act);
  if (len_bytes + 64 * 8 > S1_DYN_MAX) return false;
  const int n_sm_br = (int)std::min<size_t>((size_t)n_br, (S1_DYN_MAX - len_bytes) / (64 * 8));
  const bool fit = n_sm_br == n_br;
  const size_t own_bytes = (size_t)n_sm_br * 64 * 8 + len_bytes;
  int acts_smem = own_bytes + acts_bytes <= S1_DYN_MAX ? 1 : 0;
  const size_t dyn_bytes = own_bytes + (acts_smem ? acts_bytes : 0);
  // co-resident grid; a block takes every gridDim.x-th tile of S1_TILE patterns.  One tile per block when that
  // many blocks are co-resident (a thread then keeps one pattern for the whole launch).
  const void *kern[2][2] = {{(const void *)smooth1_kernel<false, false>, (const void *)smooth1_kernel<false, true>},
                            {(const void *)smooth1_kernel<true, false>, (const void *)smooth1_kernel<true, true>}};
  const int64_t n_tiles = Ppad / S1_TILE;
  int per_sm[2] = {0, 0};
  if (fit) {
    PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], smooth1_kernel<true, true>, S1_THREADS, dyn_bytes));
    PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], smooth1_kernel<false, true>, S1_THREADS, dyn_bytes));
  } else {
    PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[1], smooth1_kernel<true, false>, S1_THREADS, dyn_bytes));
    PLG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm[0], smooth1_kernel<false, false>, S1_THREADS, dyn_bytes));
  }
  if (per_sm[0] < 1) return false;
  int n_blocks = (int)std::min<int64_t>(n_tiles, n_tiles <= (int64_t)per_sm[1] * sc.sms ? per_sm[1] * sc.sms : per_sm[0] * sc.sms);
  if (const char *e = getenv("PLG_S1_MAX_BLOCKS")) n_blocks = std::max(1, std::min(n_blocks, atoi(e)));  // tests: several tiles per block
  const bool one = n_blocks == n_tiles;
  std::vector<double> h_brlen(n_br, 0.0);
  {
    int b = 0;
    for (int x = 0; x < tree.n_nodes; x++)  // same numbering as plan_tree
      for (int k = 0; k < 3; k++)
        if (tree.nb[x][k] > x) h_brlen[b++] = tree.len[x][k];
  }
  sc.clv.alloc((size_t)n_inner * 16 * Ppad);
  sc.scl.alloc((size_t)n_inner * Ppad);
  sc.pmat.alloc((size_t)n_br * 68);
  sc.tiptab.alloc((size_t)n_br * 256);
  sc.brlen.alloc(n_br);
  sc.ll.alloc((size_t)2 * 6 * n_blocks * n_blocks);  // inboxes: [parity][destination][source]
  sc.spill.alloc(std::max<size_t>(1, (size_t)n_blocks * (n_br - n_sm_br) * 64));
  sc.out.alloc(1);
  sc.acts.alloc(acts.size());
  cudaStream_t st = 0;
  PLG_CUDA(cudaMemcpyAsync(sc.acts.p, acts.data(), acts_bytes, cudaMemcpyHostToDevice, st));
  PLG_CUDA(cudaMemcpyAsync(sc.brlen.p, h_brlen.data(), (size_t)n_br * 8, cudaMemcpyHostToDevice, st));
  PLG_CUDA(cudaMemsetAsync(sc.ll.p, 0, (size_t)2 * 6 * n_blocks * n_blocks * 8, st));  // generation 0 = nothing sent yet
  dna4_pmatrices(aln, model, sc.brlen.p, nullptr, n_br, sc.pmat.p, sc.tiptab.p, st);
  const S1Act *d_acts = sc.acts.p;
  int n_down = (int)plan.down.size(), n_pass = (int)plan.acts.size(), passes = max_passes, n_rows = aln.n_rows, nbr = n_br, nsm = n_sm_br;
  double *spill = sc.spill.p;
  int64_t ppad = Ppad;
  const unsigned char *codes = aln.codes.p;
  const unsigned int *masks = aln.masks.p;
  const double *weights = aln.weights.p;
  const ModelDev *md = model.d.p;
  const double *pm = sc.pmat.p;
  double *clv = sc.clv.p, *bl = sc.brlen.p, *out = sc.out.p;
  int *scl = sc.scl.p;
  unsigned long long *ll = sc.ll.p;
  void *args[] = {&d_acts, &n_down, &n_pass, &passes, &n_rows, &nbr, &nsm, &spill, &ppad, &codes, &masks, &weights, &md,
                  &pm, &clv, &scl, &bl, &ll, &out, &acts_smem};
  PLG_CUDA(cudaLaunchCooperativeKernel(kern[one ? 1 : 0][fit ? 1 : 0], dim3(n_blocks), dim3(S1_THREADS), args, dyn_bytes, st));
  count_launch();
  double h_out = 0.0;
  PLG_CUDA(cudaMemcpyAsync(&h_out, sc.out.p, 8, cudaMemcpyDeviceToHost, st));
  PLG_CUDA(cudaMemcpyAsync(h_brlen.data(), sc.brlen.p, (size_t)n_br * 8, cudaMemcpyDeviceToHost, st));
  PLG_CUDA(cudaStreamSynchronize(st));
  *out_lnl = h_out;
  int b = 0;
  for (int x = 0; x < tree.n_nodes; x++)
    for (int k = 0; k < 3; k++) {
      const int y = tree.nb[x][k];
      if (y > x) { const double len = h_brlen[b++]; tree.len[x][k] = len; tree.len[y][tree.slot_of(y, x)] = len; }
    }
  return true;
}

// One handle = one tree.  DNA x Gamma-4: the whole smoothing in ONE cooperative launch (smooth1 above;
// PLG_PLL_HANDLE_MODE=batch keeps the lockstep program as a batch of one -- ~5 n launches per pass --,
// =host the literal per-branch loop with a device round trip per Newton step: the A/B references).
double PllProblem::loglik(int max_passes) {
  if (tree.n_tips < 3) PLG_FAIL(PLG_ERR_UNSUPPORTED, "need at least 3 taxa");
  const char *mode = getenv("PLG_PLL_HANDLE_MODE");
  const bool dna4 = aln->K == 4 && model->R == 4, aa20 = aln->K == 20 && model->R == 4;
  if ((mode && !strcmp(mode, "host")) || !(dna4 || aa20)) return loglik_host_loop(max_passes);
  if (dna4 && max_passes > 0 && !(mode && !strcmp(mode, "batch"))) {
    double out = 0.0;
    if (smooth1(*aln, *model, tree, tip_rows, max_passes, &out)) {
      lnl = out;
      smoothed = true;
      return lnl;
    }
  }
  std::vector<UTree> trees(1, tree);
  std::vector<std::vector<int32_t>> rows(1, tip_rows);
  double out = 0.0;
  if (dna4) {
    static thread_local BatchScratch<Dna4Family> keep[8];
    int dev = 0;
    cudaGetDevice(&dev);
    pll_loglik_batch<Dna4Family>(*aln, *model, trees, rows, max_passes, &out, &keep[dev & 7]);
  } else {
    static thread_local BatchScratch<Aa20Family> keep[8];
    int dev = 0;
    cudaGetDevice(&dev);
    pll_loglik_batch<Aa20Family>(*aln, *model, trees, rows, max_passes, &out, &keep[dev & 7]);
  }
  tree = trees[0];
  lnl = out;
  smoothed = true;
  return lnl;
}

}  // namespace plg

using namespace plg;

struct plg_pll {
  PllProblem p;
};

namespace {
struct PllAlignment {
  std::vector<std::string> names;
  std::unordered_map<std::string, int> row_of;
  std::shared_ptr<LikAln> aln;
  std::shared_ptr<Model> model;
  // identity of the file it was read from
  std::string path, model_name;
  size_t n_sites = 0;
  long long size = 0, mtime_ns = 0, ino = 0;
  int device = 0;
};

std::shared_ptr<PllAlignment> load_alignment(const char *phylip_path, const std::string &model_name, bool dna,
                                             long long range_end) {
  std::shared_ptr<PllAlignment> al(new PllAlignment());
  std::vector<std::string> seqs;
  read_phylip(phylip_path, al->names, seqs);
  const std::vector<std::string> &names = al->names;
  for (size_t i = 0; i < names.size(); i++) al->row_of.emplace(names[i], (int)i);
  // encode + compress identical columns (weights = multiplicities; lnL is unchanged)
  const size_t n = names.size(), m = seqs[0].size();
  if (range_end >= 0 && (size_t)range_end != m)
    PLG_FAIL(PLG_ERR_UNSUPPORTED, "partition covers sites 1-%lld of %zu: it must cover the whole alignment", range_end, m);
  std::vector<uint8_t> codes(n * m);
  for (size_t r = 0; r < n; r++)
    for (size_t s = 0; s < m; s++) codes[r * m + s] = dna ? dna_code(seqs[r][s]) : aa_code(seqs[r][s]);
  std::unordered_map<std::string, int> seen;
  std::vector<size_t> first;
  std::vector<double> w;
  std::string col(n, '\0');
  for (size_t s = 0; s < m; s++) {
    for (size_t r = 0; r < n; r++) col[r] = (char)codes[r * m + s];
    auto ins = seen.emplace(col, (int)first.size());
    if (ins.second) { first.push_back(s); w.push_back(1.0); }
    else w[ins.first->second] += 1.0;
  }
  const size_t P = first.size();
  al->n_sites = m;
  std::vector<uint8_t> packed(n * P);
  for (size_t r = 0; r < n; r++)
    for (size_t q = 0; q < P; q++) packed[r * P + q] = codes[r * m + first[q]];
  al->aln.reset(new LikAln(dna ? PLG_DATA_DNA : PLG_DATA_AA, (int)n, (int64_t)P, packed.data(), w.data()));
  if (dna) {
    // GTR with all six rates 1.0 and empirical base frequencies; ambiguity codes share their
    // count among compatible bases in proportion to the current estimate (8 rounds)
    double f[4] = {0.25, 0.25, 0.25, 0.25};
    for (int round = 0; round < 8; round++) {
      double cnt[4] = {0, 0, 0, 0};
      for (size_t q = 0; q < P; q++)
        for (size_t r = 0; r < n; r++) {
          const unsigned mask = packed[r * P + q];
          if (mask == 15 || mask == 0) continue;
          double tot = 0;
          for (int k = 0; k < 4; k++) if (mask >> k & 1) tot += f[k];
          for (int k = 0; k < 4; k++) if (mask >> k & 1) cnt[k] += w[q] * f[k] / tot;
        }
      double s = cnt[0] + cnt[1] + cnt[2] + cnt[3];
      if (s <= 0) break;
      for (int k = 0; k < 4; k++) f[k] = std::max(cnt[k] / s, 1e-6);
    }
    const double ones[6] = {1, 1, 1, 1, 1, 1};
    al->model.reset(new Model(4, ones, f, 1.0, 4));
  } else {
    const double *exch = nullptr, *freqs = nullptr;
    protein_model_table(model_name.c_str(), &exch, &freqs);  // (known: model_is_dna checked the name)
    al->model.reset(new Model(20, exch, freqs, 1.0, 4));
  }
  return al;
}

// most recently used first; entries keep device memory, so only a few are held
// (PLG_PLL_ALN_CACHE=0 disables the cache).  Callers hold g_api_mutex.
std::vector<std::shared_ptr<PllAlignment>> g_aln_cache;

std::shared_ptr<PllAlignment> cached_alignment(const char *phylip_path, const std::string &model_name, bool dna,
                                               long long range_end) {
  const char *env = getenv("PLG_PLL_ALN_CACHE");
  const size_t keep = env ? (size_t)std::max(0, atoi(env)) : 4;
  struct stat sb;
  if (keep == 0 || stat(phylip_path, &sb) != 0) return load_alignment(phylip_path, model_name, dna, range_end);  // (a missing file fails inside)
  int dev = 0;
  cudaGetDevice(&dev);
  const long long mt = (long long)sb.st_mtim.tv_sec * 1000000000ll + sb.st_mtim.tv_nsec;
  for (size_t i = 0; i < g_aln_cache.size(); i++) {
    const PllAlignment &c = *g_aln_cache[i];
    if (c.path == phylip_path && c.model_name == model_name && c.size == (long long)sb.st_size && c.mtime_ns == mt &&
        c.ino == (long long)sb.st_ino && c.device == dev) {
      std::shared_ptr<PllAlignment> hit = g_aln_cache[i];
      if (range_end >= 0 && range_end != (long long)hit->n_sites)
        PLG_FAIL(PLG_ERR_UNSUPPORTED, "partition covers sites 1-%lld of %zu: it must cover the whole alignment", range_end,
                 hit->n_sites);
      g_aln_cache.erase(g_aln_cache.begin() + i);
      g_aln_cache.insert(g_aln_cache.begin(), hit);
      return hit;
    }
  }
  std::shared_ptr<PllAlignment> al = load_alignment(phylip_path, model_name, dna, range_end);
  al->path = phylip_path; al->model_name = model_name;
  al->size = (long long)sb.st_size; al->mtime_ns = mt; al->ino = (long long)sb.st_ino; al->device = dev;
  g_aln_cache.insert(g_aln_cache.begin(), al);
  if (g_aln_cache.size() > keep) g_aln_cache.resize(keep);
  return al;
}
}  // namespace

extern "C" int plg_pll_new(const char *phylip_path, const char *newick, const char *partition_path, plg_pll **out) {
  PLG_API_BEGIN_LOCKED
  if (!phylip_path || !newick || !partition_path || !out) PLG_FAIL(PLG_ERR_ARG, "null argument");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    PLG_FAIL(PLG_ERR_CUDA, "no CUDA device: pylogeny_b200 has no CPU fallback");
  // The reference creates one instance per TREE and re-parses the alignment each time
  // (libpllWrapper.c:138, scoring.py:41-75).  Here the parsed, pattern-compressed, device-resident
  // alignment and its model are kept per (file identity, partition model, device) and shared by
  // the handles: a hill climb pays the 45 ms of parsing once, not per neighbour.
  long long range_end = -1;
  const std::string model_name = read_partition_model(partition_path, &range_end);
  const bool dna = model_is_dna(model_name);
  std::shared_ptr<PllAlignment> al = cached_alignment(phylip_path, model_name, dna, range_end);
  const std::vector<std::string> &names = al->names;
  std::unique_ptr<plg_pll> h(new plg_pll());
  PllProblem &pr = h->p;
  HostTree ht = parse_newick(newick);
  pr.tree = utree_from_newick(ht, pr.tip_names);
  if ((size_t)pr.tree.n_tips != names.size())
    PLG_FAIL(PLG_ERR_IO, "Could not find correct correspondances.");  // libpllWrapper.c:153-156
  for (auto &nm : pr.tip_names) {
    auto it = al->row_of.find(nm);
    if (it == al->row_of.end()) PLG_FAIL(PLG_ERR_IO, "Could not find correct correspondances.");
    pr.tip_rows.push_back(it->second);
  }
  // default branch lengths (useDefaultz = PLL_TRUE, libpllWrapper.c:152): every z = 0.9
  for (auto &l : pr.tree.len) for (auto &x : l) x = -std::log(PLL_DEFAULTZ);
  pr.aln = al->aln;
  pr.model = al->model;
  *out = h.release();
  PLG_API_END
}

extern "C" int plg_pll_destroy(plg_pll *p) {
  delete p;
  return PLG_OK;
}

extern "C" int plg_pll_loglik(plg_pll *p, double *out) {
  PLG_API_BEGIN_LOCKED
  if (!p || !out) PLG_FAIL(PLG_ERR_ARG, "null argument");
  *out = p->p.loglik(3);  // pllOptimizeBranchLengths(instance, partitions, 3), libpllWrapper.c:205
  PLG_API_END
}

// lnL at the CURRENT branch lengths, no optimisation (pllEvaluateLikelihood, libpllWrapper.c:77-78)
extern "C" int plg_pll_evaluate(plg_pll *p, double *out) {
  PLG_API_BEGIN_LOCKED
  if (!p || !out) PLG_FAIL(PLG_ERR_ARG, "null argument");
  p->p.down_pass();
  *out = p->p.evaluate();
  PLG_API_END
}

extern "C" int plg_pll_newick(plg_pll *p, char *buf, size_t *len) {
  PLG_API_BEGIN
  if (!p || !len) PLG_FAIL(PLG_ERR_ARG, "null argument");
  const std::string s = p->p.newick();
  const size_t need = s.size() + 1;
  if (buf && *len >= need) memcpy(buf, s.c_str(), need);
  const bool ok = buf && *len >= need;
  *len = need;
  if (!ok && buf) PLG_FAIL(PLG_ERR_ARG, "buffer too small: need %zu bytes", need);
  PLG_API_END
}

// get{SPR,NNI,All}MovesInDistance (libpllWrapper.c:241-319): every distinct neighbour of the
// handle's current tree within `maxdist`, its lnL at the current branch lengths (neighbour
// lengths by the reference's own rule, rearrangement.py:466-497) cast to float like :54, and its
// Newick string; best first.  Computed once here, copied out by plg_pll_moves_fetch.
extern "C" int plg_pll_moves_compute(plg_pll *p, int which, int maxdist, int64_t *n_moves, size_t *newick_bytes) {
  PLG_API_BEGIN_LOCKED
  if (!p || !n_moves || !newick_bytes) PLG_FAIL(PLG_ERR_ARG, "null argument");
  if (which != PLG_MOVE_SPR && which != PLG_MOVE_NNI && which != PLG_MOVES_ALL)
    PLG_FAIL(PLG_ERR_ARG, "moves: SPR (1), NNI (2) or all (4)");
  PllProblem &pr = p->p;
  pr.mv_kind.clear(); pr.mv_lnl.clear(); pr.mv_newick.clear();
  *n_moves = 0; *newick_bytes = 0;
  if (maxdist < 1) return PLG_OK;  // mintrav = 1 > maxtrav: libpll's search visits nothing
  Neighbourhood nb = enumerate_spr(pr.tree, which == PLG_MOVE_NNI ? 1 : maxdist);
  const int64_t M = (int64_t)nb.unique.size();
  if (M == 0) return PLG_OK;
  std::vector<double> lnl(M);
  score_spr_lik(*pr.aln, *pr.model, nb, pr.tip_rows.data(), 0, M, lnl.data(), 0);
  std::vector<int64_t> order(M);
  for (int64_t i = 0; i < M; i++) order[i] = i;
  std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return lnl[a] > lnl[b]; });
  size_t bytes = 0;
  for (int64_t i : order) {
    const Move &m = nb.moves[nb.unique[i]];
    const bool nni = which == PLG_MOVE_NNI || (which == PLG_MOVES_ALL && m.depth == 1);
    pr.mv_kind.push_back(nni ? PLG_MOVE_NNI : PLG_MOVE_SPR);
    pr.mv_lnl.push_back((float)lnl[i]);
    pr.mv_newick.push_back(pr.newick_of(apply_move(nb.t, m)));
    bytes += pr.mv_newick.back().size() + 1;
  }
  *n_moves = M;
  *newick_bytes = bytes;
  PLG_API_END
}

extern "C" int plg_pll_moves_fetch(plg_pll *p, int32_t *kinds, float *lnl, char *newick_buf, size_t newick_bytes) {
  PLG_API_BEGIN
  if (!p) PLG_FAIL(PLG_ERR_ARG, "null argument");
  const PllProblem &pr = p->p;
  size_t at = 0;
  for (size_t i = 0; i < pr.mv_kind.size(); i++) {
    if (kinds) kinds[i] = pr.mv_kind[i];
    if (lnl) lnl[i] = pr.mv_lnl[i];
    if (newick_buf) {
      const std::string &s = pr.mv_newick[i];
      if (at + s.size() + 1 > newick_bytes) PLG_FAIL(PLG_ERR_ARG, "buffer too small for the Newick strings");
      memcpy(newick_buf + at, s.c_str(), s.size() + 1);  // NUL-separated
      at += s.size() + 1;
    }
  }
  PLG_API_END
}

// Model and pattern set of a problem, for tests: exch[K(K-1)/2], freqs[K], rates[R], and counts.
extern "C" int plg_pll_info(plg_pll *p, int *n_states, int64_t *n_patterns, double *freqs, double *rates) {
  PLG_API_BEGIN
  if (!p) PLG_FAIL(PLG_ERR_ARG, "null argument");
  if (n_states) *n_states = p->p.aln->K;
  if (n_patterns) *n_patterns = p->p.aln->P;
  if (freqs) for (int k = 0; k < p->p.aln->K; k++) freqs[k] = p->p.model->h.freqs[k];
  if (rates) for (int r = 0; r < p->p.model->R; r++) rates[r] = p->p.model->h.rates[r];
  PLG_API_END
}

// ---- batched getLogLikelihood ----------------------------------------------------------------
namespace {
thread_local std::vector<std::string> g_batch_newicks;  // optimised trees of the last batch on this thread
}

// lnL of every tree after the reference's per-tree treatment (scoring.getLogLikelihood ->
// libpllWrapper new / getLogLikelihood / getNewickString / destroy, scoring.py:41-75): default
// branch lengths z = 0.9, up to `passes` smoothing passes (the reference uses 3), likelihood.
// DNA x Gamma-4 runs the lockstep batch above; other models fall back to one handle per tree.
extern "C" int plg_pll_loglik_batch(const char *phylip_path, const char *partition_path, int64_t n_trees,
                                    const char *const *newicks, int passes, double *out_lnl,
                                    size_t *newick_bytes) {
  PLG_API_BEGIN_LOCKED
  if (!phylip_path || !partition_path || (n_trees > 0 && (!newicks || !out_lnl)))
    PLG_FAIL(PLG_ERR_ARG, "null argument");
  if (passes < 0) PLG_FAIL(PLG_ERR_ARG, "negative number of smoothing passes");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    PLG_FAIL(PLG_ERR_CUDA, "no CUDA device: pylogeny_b200 has no CPU fallback");
  long long range_end = -1;
  const std::string model_name = read_partition_model(partition_path, &range_end);
  const bool dna = model_is_dna(model_name);
  std::shared_ptr<PllAlignment> al = cached_alignment(phylip_path, model_name, dna, range_end);
  std::vector<UTree> trees(n_trees);
  std::vector<std::vector<std::string>> names(n_trees);
  std::vector<std::vector<int32_t>> rows(n_trees);
  for (int64_t i = 0; i < n_trees; i++) {
    if (!newicks[i]) PLG_FAIL(PLG_ERR_ARG, "tree %lld is NULL", (long long)i);
    HostTree ht = parse_newick(newicks[i]);
    trees[i] = utree_from_newick(ht, names[i]);
    if ((size_t)trees[i].n_tips != al->names.size())
      PLG_FAIL(PLG_ERR_IO, "Could not find correct correspondances.");  // libpllWrapper.c:153-156
    for (auto &nm : names[i]) {
      auto it = al->row_of.find(nm);
      if (it == al->row_of.end()) PLG_FAIL(PLG_ERR_IO, "Could not find correct correspondances.");
      rows[i].push_back(it->second);
    }
    for (auto &l : trees[i].len) for (auto &x : l) x = -std::log(PLL_DEFAULTZ);
  }
  g_batch_newicks.assign(n_trees, std::string());
  const bool dna4 = al->aln->K == 4 && al->model->R == 4, aa20 = al->aln->K == 20 && al->model->R == 4;
  if (dna4 || aa20) {
    if (dna4) pll_loglik_batch<Dna4Family>(*al->aln, *al->model, trees, rows, passes, out_lnl);
    else pll_loglik_batch<Aa20Family>(*al->aln, *al->model, trees, rows, passes, out_lnl);
    PllProblem writer;  // (only its Newick writer is used)
    for (int64_t i = 0; i < n_trees; i++) {
      writer.tip_names = names[i];
      g_batch_newicks[i] = writer.newick_of(trees[i]);
    }
  } else {
    for (int64_t i = 0; i < n_trees; i++) {
      PllProblem pr;
      pr.tree = trees[i]; pr.tip_names = names[i]; pr.tip_rows = rows[i];
      pr.aln = al->aln; pr.model = al->model;
      out_lnl[i] = pr.loglik(passes);
      g_batch_newicks[i] = pr.newick();
    }
  }
  if (newick_bytes) {
    size_t b = 0;
    for (auto &s : g_batch_newicks) b += s.size() + 1;
    *newick_bytes = b;
  }
  PLG_API_END
}

// The optimised trees of the last plg_pll_loglik_batch on this thread, NUL-separated, in input order.
extern "C" int plg_pll_batch_newicks(char *buf, size_t bytes) {
  PLG_API_BEGIN
  if (!buf) PLG_FAIL(PLG_ERR_ARG, "null argument");
  size_t at = 0;
  for (auto &s : g_batch_newicks) {
    if (at + s.size() + 1 > bytes) PLG_FAIL(PLG_ERR_ARG, "buffer too small for the Newick strings");
    memcpy(buf + at, s.c_str(), s.size() + 1);
    at += s.size() + 1;
  }
  PLG_API_END
}
