// Felsenstein-pruning log-likelihood on sm_100a (FP64 CUDA cores), level-synchronous op programs.
//
// Replaces the libpll 1.0.2 kernels reached from libpllWrapper.c (pllInitModel :159,
// newview/evaluate behind pllOptimizeBranchLengths :205 and pllEvaluateLikelihood :77):
//   pmatrix_kernel   P(t * r_c) = U exp(L t r_c) U^-1 per unique branch length and Gamma
//                    category, plus the tip lookup table sum_{j in code} P[.][j]
//   clv_update_kernel  x3 = (P_L x1) .* (P_R x2) per pattern and category, with per-pattern
//                    2^-256 rescaling counters (PLL_MINLIKELIHOOD semantics)
//   root_eval_kernel site likelihood at a branch, log, pattern weights, scaler correction
//   reduce_partials_kernel  fixed-order (deterministic) sum of the per-block partials
//
// Data layout in HBM: a conditional likelihood vector is pattern-minor SoA,
// clv[slot][r*K + k][Ppad] doubles (+ int32 scaler[slot][Ppad]), so that the 32 lanes of a
// warp -- 8 consecutive patterns x 4 rate categories -- read 64-byte contiguous runs of four
// planes: every 32-byte sector fetched is fully used.  Tips are 1 byte per pattern.
// 4x4 / 20x20 per-site products are not a dense contraction worth tensor cores
// (BASELINE.json north_star); the roofline that bounds the DNA kernels is HBM.
#include <algorithm>
#include <cmath>
#include <cstring>

#include "lik.h"
#include "lik_pmatrix.cuh"
#include "treewalk.h"

namespace plg {

constexpr double TWO256 = 1.15792089237316195423570985008687907853269984665640564039457584007913129639936e77;
constexpr double MINLH = 1.0 / TWO256;
// log(2^-256) = -256 RN(ln 2) exactly (a power-of-two multiple of the correctly rounded ln 2); spelling it
// log(MINLH) in device code costs a full FP64 log evaluation per site term -- nvcc does not fold it
constexpr double LOG_MINLH = -256.0 * 0.693147180559945309417232121458176568;
constexpr int LIK_THREADS = 256;

// ----------------------------------------------------------------------------
// device
// ----------------------------------------------------------------------------
template <int K, int R>
__global__ void pmatrix_kernel(const ModelDev *__restrict__ model, const double *__restrict__ brlen,
                               const unsigned int *__restrict__ masks, double *__restrict__ pmat,
                               double *__restrict__ tiptab, const int *__restrict__ idx = nullptr) {
  constexpr int KKP = K * K + 1, NC = NCodes<K>::value;
  const int pm = idx ? idx[blockIdx.x] : blockIdx.x;  // (idx: only the entries whose length changed)
  pmatrix_block<K, R>(model, brlen[pm], masks, pmat + (int64_t)pm * R * KKP, tiptab + (int64_t)pm * NC * R * K);
}

// Shared-memory staging of one operand's matrix: tip -> lookup table, inner -> P-matrix.
template <int K, int R>
__device__ __forceinline__ void stage_operand(double *sm, bool is_tip, int pm, const double *__restrict__ pmat,
                                              const double *__restrict__ tiptab) {
  constexpr int KKP = K * K + 1, NC = NCodes<K>::value;
  if (is_tip) {
    const double *src = tiptab + (int64_t)pm * NC * R * K;
    for (int e = threadIdx.x; e < NC * R * K; e += blockDim.x) sm[e] = src[e];
  } else {
    const double *src = pmat + (int64_t)pm * R * KKP;
    for (int e = threadIdx.x; e < R * KKP; e += blockDim.x) sm[e] = src[e];
  }
}

// (P x)[0..K) for this thread's pattern p and category r.
template <int K, int R>
__device__ __forceinline__ void apply_operand(const double *sm, bool is_tip, int id, int n_rows, int64_t p, int r,
                                              int64_t Ppad, const unsigned char *__restrict__ codes,
                                              const double *__restrict__ clv, double (&v)[K]) {
  constexpr int KKP = K * K + 1;
  if (is_tip) {
    const int code = codes[(int64_t)id * Ppad + p];
    const double *row = sm + (code * R + r) * K;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = row[k];
  } else {
    const double *src = clv + ((int64_t)(id - n_rows) * R * K + r * K) * Ppad + p;
    double x[K];
#pragma unroll
    for (int j = 0; j < K; j++) x[j] = __ldg(src + (int64_t)j * Ppad);
    const double *Pm = sm + r * KKP;
#pragma unroll
    for (int k = 0; k < K; k++) {
      double s = 0.0;
#pragma unroll
      for (int j = 0; j < K; j++) s = fma(Pm[k * K + j], x[j], s);
      v[k] = s;
    }
  }
}

template <int K, int R>
__global__ void __launch_bounds__(LIK_THREADS)
clv_update_kernel(const LikOp *__restrict__ ops, int blocks_per_op, int n_rows, int64_t Ppad,
                  const unsigned char *__restrict__ codes, const double *__restrict__ pmat,
                  const double *__restrict__ tiptab, double *__restrict__ clv, int *__restrict__ scl) {
  constexpr int KKP = K * K + 1, NC = NCodes<K>::value;
  constexpr int SM = (NC * R * K > R * KKP) ? NC * R * K : R * KKP;
  __shared__ double smA[SM];
  __shared__ double smB[SM];
  const int op_idx = blockIdx.x / blocks_per_op;
  const int pb = blockIdx.x - op_idx * blocks_per_op;
  const LikOp op = ops[op_idx];
  const bool tipA = op.a < n_rows, tipB = op.b >= 0 && op.b < n_rows;
  stage_operand<K, R>(smA, tipA, op.pa, pmat, tiptab);
  if (op.b >= 0) stage_operand<K, R>(smB, tipB, op.pb, pmat, tiptab);
  __syncthreads();
  const int r = threadIdx.x % R;
  const int64_t p = (int64_t)pb * (LIK_THREADS / R) + threadIdx.x / R;
  // Ppad is a multiple of LIK_THREADS/R, so every thread owns a (possibly padding) pattern
  double va[K], vb[K];
  apply_operand<K, R>(smA, tipA, op.a, n_rows, p, r, Ppad, codes, clv, va);
  int sc = tipA ? 0 : scl[(int64_t)(op.a - n_rows) * Ppad + p];
  if (op.b >= 0) {
    apply_operand<K, R>(smB, tipB, op.b, n_rows, p, r, Ppad, codes, clv, vb);
    if (!tipB) sc += scl[(int64_t)(op.b - n_rows) * Ppad + p];
#pragma unroll
    for (int k = 0; k < K; k++) va[k] *= vb[k];
  }
  double mx = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) mx = fmax(mx, fabs(va[k]));
#pragma unroll
  for (int d = 1; d < R; d <<= 1) mx = fmax(mx, __shfl_xor_sync(0xffffffffu, mx, d));
  const bool rescale = mx < MINLH;  // all R*K entries below 2^-256
  double *dst = clv + ((int64_t)(op.dst - n_rows) * R * K + r * K) * Ppad + p;
#pragma unroll
  for (int k = 0; k < K; k++) dst[(int64_t)k * Ppad] = rescale ? va[k] * TWO256 : va[k];
  if (r == 0) scl[(int64_t)(op.dst - n_rows) * Ppad + p] = sc + (rescale ? 1 : 0);
}

// Operand value at the evaluation point: P x when a matrix is given, x itself when pm < 0.
template <int K, int R>
__device__ __forceinline__ void eval_operand(const double *sm, int id, int pm, int n_rows, int64_t p, int r,
                                             int64_t Ppad, const unsigned char *__restrict__ codes,
                                             const unsigned int *__restrict__ masks, const double *__restrict__ clv,
                                             const int *__restrict__ scl, double (&v)[K], int &sc) {
  const bool tip = id < n_rows;
  if (pm >= 0) {
    apply_operand<K, R>(sm, tip, id, n_rows, p, r, Ppad, codes, clv, v);
  } else if (tip) {
    const unsigned int m = masks[codes[(int64_t)id * Ppad + p]];
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = (m >> k & 1) ? 1.0 : 0.0;
  } else {
    const double *src = clv + ((int64_t)(id - n_rows) * R * K + r * K) * Ppad + p;
#pragma unroll
    for (int k = 0; k < K; k++) v[k] = __ldg(src + (int64_t)k * Ppad);
  }
  if (!tip) sc += scl[(int64_t)(id - n_rows) * Ppad + p];
}

// Site log-likelihood at a node joining up to three subtrees:
//   L_p = (1/R) sum_r sum_k pi_k prod_{o in a,b,c} (P_o x_o)[k]
// Two operands = evaluation across a branch (libpll `evaluate`); three = a subtree inserted
// into the middle of an edge (edge-insertion scoring of an SPR neighbour).
template <int K, int R>
__global__ void __launch_bounds__(LIK_THREADS)
root_eval_kernel(const LikEval *__restrict__ evals, int blocks_per_eval, int n_rows, int64_t Ppad,
                 const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
                 const double *__restrict__ weights, const ModelDev *__restrict__ model,
                 const double *__restrict__ pmat, const double *__restrict__ tiptab,
                 const double *__restrict__ clv, const int *__restrict__ scl, double *__restrict__ partial) {
  constexpr int KKP = K * K + 1, NC = NCodes<K>::value;
  constexpr int SM = (NC * R * K > R * KKP) ? NC * R * K : R * KKP;
  __shared__ double sm[3][SM];
  __shared__ double warp_sums[LIK_THREADS / 32];
  const int ev_idx = blockIdx.x / blocks_per_eval;
  const int pb = blockIdx.x - ev_idx * blocks_per_eval;
  const LikEval ev = evals[ev_idx];
  if (ev.pa >= 0) stage_operand<K, R>(sm[0], ev.a < n_rows, ev.pa, pmat, tiptab);
  if (ev.pb >= 0) stage_operand<K, R>(sm[1], ev.b < n_rows, ev.pb, pmat, tiptab);
  if (ev.c >= 0 && ev.pc >= 0) stage_operand<K, R>(sm[2], ev.c < n_rows, ev.pc, pmat, tiptab);
  __syncthreads();
  const int r = threadIdx.x % R;
  const int64_t p = (int64_t)pb * (LIK_THREADS / R) + threadIdx.x / R;
  double v[K], x[K];
  int sc = 0;
  eval_operand<K, R>(sm[0], ev.a, ev.pa, n_rows, p, r, Ppad, codes, masks, clv, scl, v, sc);
  eval_operand<K, R>(sm[1], ev.b, ev.pb, n_rows, p, r, Ppad, codes, masks, clv, scl, x, sc);
#pragma unroll
  for (int k = 0; k < K; k++) v[k] *= x[k];
  if (ev.c >= 0) {
    eval_operand<K, R>(sm[2], ev.c, ev.pc, n_rows, p, r, Ppad, codes, masks, clv, scl, x, sc);
#pragma unroll
    for (int k = 0; k < K; k++) v[k] *= x[k];
  }
  double term = 0.0;
#pragma unroll
  for (int k = 0; k < K; k++) term = fma(model->freqs[k], v[k], term);
#pragma unroll
  for (int d = 1; d < R; d <<= 1) term += __shfl_xor_sync(0xffffffffu, term, d);
  double val = 0.0;
  if (r == 0) {
    const double w = weights[p];
    if (w != 0.0) val = w * (log(term * (1.0 / R)) + sc * LOG_MINLH);
  }
  // fixed-shape block reduction (deterministic order)
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int w = 0; w < LIK_THREADS / 32; w++) s += warp_sums[w];
    partial[(int64_t)ev.tree * blocks_per_eval + pb] = s;
  }
}

// One warp per tree: strided partial sums, then a fixed shuffle tree (deterministic order).
__global__ void reduce_partials_kernel(int n_trees, int blocks_per_eval, const double *__restrict__ partial,
                                       double *__restrict__ lnl) {
  const int t = blockIdx.x * (blockDim.x / 32) + threadIdx.x / 32;
  if (t >= n_trees) return;
  const int lane = threadIdx.x & 31;
  double s = 0.0;
  for (int b = lane; b < blocks_per_eval; b += 32) s += partial[(int64_t)t * blocks_per_eval + b];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) s += __shfl_down_sync(0xffffffffu, s, d);
  if (lane == 0) lnl[t] = s;
}

}  // namespace plg
#include "lik_dna.cuh"
namespace plg {
constexpr int DNA4_PATTERNS_PER_BLOCK = D4_THREADS;
}
#include "lik_aa.cuh"
#include "lik_tbr20.cuh"
#include "lik_walkm.cuh"
#include "lik_twalk.cuh"
namespace plg {

// ----------------------------------------------------------------------------
// host: alignment
// ----------------------------------------------------------------------------
// largest tip code of the uploaded alignment (range check of LikAln's constructor)
__global__ void max_code_kernel(const uint4 *__restrict__ codes, int64_t n16, unsigned int *__restrict__ out) {
  unsigned int m = 0;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) {
    const uint4 v = __ldg(codes + i);
    m = __vmaxu4(m, __vmaxu4(__vmaxu4(v.x, v.y), __vmaxu4(v.z, v.w)));
  }
  m = max(max(m & 255u, (m >> 8) & 255u), max((m >> 16) & 255u, m >> 24));
  m = __reduce_max_sync(0xffffffffu, m);
  if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

LikAln::LikAln(int datatype, int n_rows_, int64_t P_, const uint8_t *h_codes, const double *h_w) {
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    PLG_FAIL(PLG_ERR_CUDA, "no CUDA device: pylogeny_b200 has no CPU fallback");
  if (datatype != PLG_DATA_DNA && datatype != PLG_DATA_AA) PLG_FAIL(PLG_ERR_ARG, "datatype must be 4 (DNA) or 20 (AA)");
  if (n_rows_ <= 0 || P_ < 0) PLG_FAIL(PLG_ERR_ARG, "alignment needs n_rows > 0 and n_patterns >= 0");
  K = datatype;
  n_rows = n_rows_;
  P = P_;
  if (K == 4) {
    n_codes = 16;
    for (unsigned int c = 0; c < 16; c++) h_masks.push_back(c ? c : 15u);  // code 0 (no state) reads as unknown
  } else {
    n_codes = 23;
    for (int c = 0; c < 20; c++) h_masks.push_back(1u << c);
    h_masks.push_back((1u << 2) | (1u << 3));  // B = N|D   (order ARNDCQEGHILKMFPSTWYV)
    h_masks.push_back((1u << 5) | (1u << 6));  // Z = Q|E
    h_masks.push_back((1u << 20) - 1);          // X, -, ?
  }
  const int64_t tile = LIK_THREADS;  // multiple of LIK_THREADS / R for every supported R
  Ppad = std::max<int64_t>(tile, (P + tile - 1) / tile * tile);
  const uint8_t any = (uint8_t)(K == 4 ? 15 : 22);
  // Rows go up straight from the caller's buffer (a pitched copy: DMA at full rate when it is pinned; a padded
  // host copy of a 128 x 1M alignment cost more than scoring 64 trees on it), padding and the range check
  // happen on the device.
  codes.alloc((size_t)n_rows * Ppad);
  weights.alloc((size_t)Ppad);
  masks.alloc(h_masks.size());
  if (P > 0) PLG_CUDA(cudaMemcpy2D(codes.p, (size_t)Ppad, h_codes, (size_t)P, (size_t)P, (size_t)n_rows, cudaMemcpyHostToDevice));
  if (Ppad > P) PLG_CUDA(cudaMemset2D(codes.p + P, (size_t)Ppad, any, (size_t)(Ppad - P), (size_t)n_rows));
  PLG_CUDA(cudaMemset(weights.p, 0, (size_t)Ppad * 8));
  if (P > 0) PLG_CUDA(cudaMemcpy(weights.p, h_w, (size_t)P * 8, cudaMemcpyHostToDevice));
  PLG_CUDA(cudaMemcpy(masks.p, h_masks.data(), h_masks.size() * 4, cudaMemcpyHostToDevice));
  DevBuf<unsigned int> d_max;
  d_max.alloc(1);
  PLG_CUDA(cudaMemset(d_max.p, 0, sizeof(unsigned int)));
  const int64_t n16 = (int64_t)n_rows * Ppad / 16;  // (Ppad is a multiple of 256)
  max_code_kernel<<<(unsigned)std::min<int64_t>((n16 + 255) / 256, 148 * 8), 256>>>(reinterpret_cast<const uint4 *>(codes.p), n16, d_max.p);
  count_launch();
  PLG_CUDA(cudaGetLastError());
  unsigned int mx = 0;
  PLG_CUDA(cudaMemcpy(&mx, d_max.p, sizeof(unsigned int), cudaMemcpyDeviceToHost));
  if (mx >= (unsigned int)n_codes)  // (the error path alone walks the host copy, to name the place)
    for (int r = 0; r < n_rows; r++)
      for (int64_t p = 0; p < P; p++)
        if (h_codes[(size_t)r * P + p] >= n_codes)
          PLG_FAIL(PLG_ERR_ARG, "tip code %d at row %d pattern %lld is out of range", h_codes[(size_t)r * P + p], r, (long long)p);
}

// ----------------------------------------------------------------------------
// host: programs
// ----------------------------------------------------------------------------
// opt-in to > 48 KB of dynamic shared memory once per device (function attributes are per context)
template <typename F>
static void ensure_dyn_smem(F *kernel, int bytes, bool (&done)[64]) {
  int dev = 0;
  cudaGetDevice(&dev);
  if (done[dev & 63]) return;
  PLG_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
  done[dev & 63] = true;
}

LikWorkspace &lik_workspace() {
  static LikWorkspace *ws[64] = {nullptr};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!ws[dev & 63]) ws[dev & 63] = new LikWorkspace();  // intentionally never freed (process lifetime)
  return *ws[dev & 63];
}

void LikProgram::clear() {
  walk_visits.clear();
  walk_groups.clear();
  walk_levels = 0;
  spr_edges.clear();
  spr_searches.clear();
  spr_n_visits = spr_n_unique = 0;
  spr_lo = spr_hi = 0;
  ops.clear();
  level_off.clear();
  t20_ops.clear();
  t20_level_off.clear();
  t20_tiles.clear();
  t20_tile_off.clear();
  t20_pair_slots = 0;
  t20_products = t20_bytes = t20_pair_bytes = 0;
  evals.clear();
  brlens.clear();
  pm_index.clear();
  n_slots = n_trees = 0;
}

int LikProgram::pm(double t) {
  if (!(t >= 0.0)) t = 0.0;  // negative / NaN lengths are read as zero
  uint64_t bits;
  memcpy(&bits, &t, 8);
  auto it = pm_index.find(bits);
  if (it != pm_index.end()) return it->second;
  int id = (int)brlens.size();
  brlens.push_back(t);
  pm_index.emplace(bits, id);
  return id;
}

// Full re-traversal: every non-root internal node is one CLV update (k-ary nodes become a chain
// of binary updates joined by zero-length branches), the root is one branch evaluation.
void build_full_lik_program(const TreeBatch &b, int64_t t0, int64_t t1, int n_rows, LikProgram &prog) {
  prog.clear();
  if (b.child_len.size() != b.child.size()) PLG_FAIL(PLG_ERR_ARG, "likelihood programs need branch lengths");
  struct Tmp { LikOp op; int level; };
  std::vector<Tmp> tmp;
  std::vector<int> node_id, node_level;
  int64_t slot = 0;
  int max_level = 0;
  for (int64_t t = t0; t < t1; t++) {
    const int64_t base = b.node_off[t], end = b.node_off[t + 1];
    node_id.assign(end - base, -1);
    node_level.assign(end - base, 0);
    for (int64_t x = base; x < end; x++) {
      const int64_t c0 = b.child_off[x], k = b.child_off[x + 1] - c0;
      const bool is_root = (x == end - 1);
      auto operand = [&](int64_t c, int &id, int &lvl) {
        int32_t ch = b.child[c];
        if (ch >= 0) { id = ch; lvl = 0; }
        else { id = node_id[-ch - 1]; lvl = node_level[-ch - 1]; }
      };
      if (k < 1) PLG_FAIL(PLG_ERR_ARG, "internal node without children");
      if (is_root && k == 1) PLG_FAIL(PLG_ERR_UNSUPPORTED, "likelihood of a tree whose root has a single child");
      int acc_id, acc_lvl;
      double acc_len = b.child_len[c0];
      operand(c0, acc_id, acc_lvl);
      auto emit = [&](int cid, int clvl, double clen) {
        Tmp e;
        e.op.dst = n_rows + (int)slot++;
        e.op.a = acc_id; e.op.pa = prog.pm(acc_len);
        e.op.b = cid; e.op.pb = cid >= 0 ? prog.pm(clen) : 0;
        e.level = std::max(acc_lvl, clvl) + 1;
        tmp.push_back(e);
        acc_id = e.op.dst; acc_lvl = e.level; acc_len = 0.0;  // the chain continues AT this node
        max_level = std::max(max_level, e.level);
      };
      if (k == 1) emit(-1, 0, 0.0);  // unary node: CLV = P(t) x
      for (int64_t j = 1; j < k; j++) {
        int cid, clvl;
        operand(c0 + j, cid, clvl);
        const double clen = b.child_len[c0 + j];
        if (is_root && j == k - 1) {
          // site likelihood across the last branch; sum_k pi_k A_k (P B)_k is symmetric in A,B
          // (reversibility), and for a bifurcating root P(ta)A . P(tb)B collapses to P(ta+tb)
          LikEval ev;
          ev.a = acc_id; ev.pa = -1; ev.b = cid; ev.pb = prog.pm(acc_len + clen);
          ev.c = -1; ev.pc = -1; ev.pad = 0;
          if (ev.b < n_rows && ev.a >= n_rows) std::swap(ev.a, ev.b);  // tip on the indicator side
          ev.tree = (int)(t - t0);
          prog.evals.push_back(ev);
          max_level = std::max(max_level, std::max(acc_lvl, clvl));
        } else {
          emit(cid, clvl, clen);
        }
      }
      node_id[x - base] = acc_id;
      node_level[x - base] = acc_lvl;
    }
  }
  prog.level_off.assign(max_level + 2, 0);
  for (auto &e : tmp) prog.level_off[e.level + 1]++;
  for (int l = 0; l <= max_level; l++) prog.level_off[l + 1] += prog.level_off[l];
  std::vector<int64_t> cur(prog.level_off.begin(), prog.level_off.end() - 1);
  prog.ops.resize(tmp.size());
  for (auto &e : tmp) prog.ops[cur[e.level]++] = e.op;
  prog.n_slots = slot;
  prog.n_trees = t1 - t0;
}

template <int K, int R>
static void run_typed(const LikAln &a, const Model &m, LikWorkspace &ws, const LikProgram &prog, cudaStream_t st) {
  constexpr int KKP = K * K + 1, NC = NCodes<K>::value;
  const int n_pm = (int)prog.brlens.size();
  ws.pmat.alloc(std::max<size_t>(1, (size_t)n_pm * R * KKP));
  ws.tiptab.alloc(std::max<size_t>(1, (size_t)n_pm * NC * R * K));
  ws.brlen.alloc(std::max<size_t>(1, n_pm));
  if (n_pm) {
    PLG_CUDA(cudaMemcpyAsync(ws.brlen.p, prog.brlens.data(), (size_t)n_pm * 8, cudaMemcpyHostToDevice, st));
    prof_begin(PROF_PMATRIX, st);
    pmatrix_kernel<K, R><<<n_pm, 128, 0, st>>>(m.d.p, ws.brlen.p, a.masks.p, ws.pmat.p, ws.tiptab.p);
    count_launch();
    prof_end(PROF_PMATRIX, st, (double)n_pm * (R * KKP + NC * R * K) * 8.0, n_pm);
  }
  const int per_block = LIK_THREADS / R;
  const int64_t bpo = a.Ppad / per_block;
  for (size_t l = 0; l + 1 < prog.level_off.size(); l++) {
    const int64_t o0 = prog.level_off[l], cnt = prog.level_off[l + 1] - o0;
    const int64_t max_ops = std::max<int64_t>(1, ((int64_t)1 << 30) / bpo);
    for (int64_t o = 0; o < cnt; o += max_ops) {
      const int64_t c = std::min(max_ops, cnt - o);
      // algorithmic bytes: C + 4 per inner CLV moved (scaler included), 1 per tip code
      double bytes = 0;
      const double Cb = (double)R * K * 8.0 + 4.0;
      for (int64_t q = o0 + o; q < o0 + o + c; q++) {
        const LikOp &op = prog.ops[q];
        bytes += Cb + (op.a < a.n_rows ? 1.0 : Cb) + (op.b < 0 ? 0.0 : (op.b < a.n_rows ? 1.0 : Cb));
      }
      prof_begin(PROF_CLV_UPDATE, st);
      clv_update_kernel<K, R><<<(unsigned)(c * bpo), LIK_THREADS, 0, st>>>(
          ws.ops.p + o0 + o, (int)bpo, a.n_rows, a.Ppad, a.codes.p, ws.pmat.p, ws.tiptab.p, ws.clv.p, ws.scl.p);
      count_launch();
      prof_end(PROF_CLV_UPDATE, st, bytes * (double)a.P, (double)c * (double)a.P);
    }
  }
  const int64_t n_ev = (int64_t)prog.evals.size();
  if (n_ev) {
    const int64_t max_ev = std::max<int64_t>(1, ((int64_t)1 << 30) / bpo);
    for (int64_t e = 0; e < n_ev; e += max_ev) {
      const int64_t c = std::min(max_ev, n_ev - e);
      double bytes = 0;
      const double Cb = (double)R * K * 8.0 + 4.0;
      for (int64_t q = e; q < e + c; q++) {
        const LikEval &ev = prog.evals[q];
        bytes += 8.0 + (ev.a < a.n_rows ? 1.0 : Cb) + (ev.b < a.n_rows ? 1.0 : Cb) +
                 (ev.c < 0 ? 0.0 : (ev.c < a.n_rows ? 1.0 : Cb));
      }
      prof_begin(PROF_ROOT_EVAL, st);
      root_eval_kernel<K, R><<<(unsigned)(c * bpo), LIK_THREADS, 0, st>>>(
          ws.evals.p + e, (int)bpo, a.n_rows, a.Ppad, a.codes.p, a.masks.p, a.weights.p, m.d.p, ws.pmat.p,
          ws.tiptab.p, ws.clv.p, ws.scl.p, ws.partial.p);
      count_launch();
      prof_end(PROF_ROOT_EVAL, st, bytes * (double)a.P, (double)c * (double)a.P);
    }
    reduce_partials_kernel<<<(unsigned)((prog.n_trees + 7) / 8), 256, 0, st>>>((int)prog.n_trees, (int)bpo,
                                                                               ws.partial.p, ws.lnl.p);
    count_launch();
  }
}

// Lays out the search walks of a full SPR neighbourhood on the device (likelihood twin of
// fitch.cu spr_plan_kernel): one thread per search runs the depth-first traversal of the host
// planner (neighbourhood.cpp build_lik_nb_program, mode walk) and writes the LikWalkVisit list the
// walk kernel executes next on the same stream; unique ids follow the host enumeration
// (enumerate_spr).  acct (profiling only): [0] algorithmic, [1] requested bytes per pattern of the
// walk launch, the same per-visit sums run_dna4 takes over a host-built visit list.
constexpr size_t SPR_LIK_PLAN_SMEM_MAX = 200 << 10;
__global__ void __launch_bounds__(64)
spr_lik_plan_kernel(const int4 *__restrict__ edges_g, int n_edge_int4, const int4 *__restrict__ searches,
                    int n_search, int n_rows, int id_lo, int id_hi, int4 *__restrict__ visits,
                    int4 *__restrict__ out_moves, double *__restrict__ acct) {
  extern __shared__ int4 spr_lik_edges_s[];  // the edge records, when they fit (n_edge_int4 > 0)
  const int4 *edges = edges_g;
  if (n_edge_int4 > 0) {
    for (int i = threadIdx.x; i < n_edge_int4; i += blockDim.x) spr_lik_edges_s[i] = __ldg(edges_g + i);
    __syncthreads();
    edges = spr_lik_edges_s;
  }
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n_search) return;
  const int4 r0 = __ldg(searches + 3 * s), r1 = __ldg(searches + 3 * s + 1), r2 = __ldg(searches + 3 * s + 2);
  const int move_cnt = r1.y;
  if (move_cnt == 0) return;
  const int uniq_base = r1.x, dup = r1.z & 3, complement = (r1.z >> 2) & 1, pe = r1.w;
  const int dup0 = dup & 1, dup01 = dup0 + (dup >> 1);
  constexpr double Cb = 4 * 4 * 8.0 + 4.0;
  const double o_tv = r2.y < n_rows ? 1.0 : Cb;
  double bytes = 0.0, req = 0.0;
  constexpr int LV = WALK_MAX_LEVELS + 2;
  int sx[LV], sfrom[LV], spre[LV], sin[LV], sdep[LV], spin[LV];
  int sp = 1, park = 0;
  int64_t at = r0.w;
  sx[0] = r0.x; sfrom[0] = r0.y; spre[0] = 0; sin[0] = r0.z; sdep[0] = 1; spin[0] = r2.x;
  while (sp > 0) {
    sp--;
    const int x = sx[sp], from = sfrom[sp], pre = spre[sp], x_in = sin[sp], dep = sdep[sp], p_in = spin[sp];
    if (x_in <= -2) park--;
    const int4 *ep = edges + 6 * x;
    const int4 e0 = ep[0], e1 = ep[2], e2 = ep[4];
    // the two neighbours other than `from`: j1 = higher slot (move 2*pre), j2 = lower slot (2*pre + 1)
    const int j1 = e2.x == from ? 1 : 2, j2 = e0.x == from ? 1 : 0;
    const int4 E1 = j1 == 1 ? e1 : e2, E2 = j2 == 1 ? e1 : e0;
    const int4 Q1 = ep[2 * j1 + 1], Q2 = ep[2 * j2 + 1];
    int uid1, uid2;
    if (pre == 0) {
      uid1 = dup0 ? -1 : uniq_base;
      uid2 = (dup >> 1) ? -1 : uniq_base + 1 - dup0;
    } else {
      uid1 = uniq_base + 2 * pre - dup01;
      uid2 = uid1 + 1;
    }
    const int sub1 = E1.z, sub2 = E2.z;  // visits below: 0 = the neighbour is a tip
    const int pre2 = pre + 1, pre1 = pre + 1 + sub2;
    int keep = 0, push = -1;
    if (sub1 > 0 && sub2 > 0) {
      const bool first1 = sub1 <= sub2;
      keep = first1 ? 1 : 2;
      push = park;
      sx[sp] = first1 ? E2.x : E1.x; sfrom[sp] = x; spre[sp] = first1 ? pre2 : pre1; sin[sp] = -2 - park;
      sdep[sp] = dep + 1; spin[sp] = first1 ? Q2.x : Q1.x;
      sp++;
      sx[sp] = first1 ? E1.x : E2.x; sfrom[sp] = x; spre[sp] = first1 ? pre1 : pre2; sin[sp] = -1;
      sdep[sp] = dep + 1; spin[sp] = first1 ? Q1.x : Q2.x;
      sp++;
      park++;
    } else if (sub1 > 0) {
      keep = 1;
      sx[sp] = E1.x; sfrom[sp] = x; spre[sp] = pre1; sin[sp] = -1; sdep[sp] = dep + 1; spin[sp] = Q1.x;
      sp++;
    } else if (sub2 > 0) {
      keep = 2;
      sx[sp] = E2.x; sfrom[sp] = x; spre[sp] = pre2; sin[sp] = -1; sdep[sp] = dep + 1; spin[sp] = Q2.x;
      sp++;
    }
    // normal form (walk4m_kernel relies on it): the partial that continues in registers is the one
    // toward d1, so the visit's "2" side is the one that may be parked
    if (out_moves) {
      if (uid1 >= 0) out_moves[uid1] = make_int4(pe, complement, E1.w, dep);
      if (uid2 >= 0) out_moves[uid2] = make_int4(pe, complement, E2.w, dep);
    }
    // trees exist for the scored range of unique ids only (a shard scores [id_lo, id_hi))
    uid1 = uid1 >= id_lo && uid1 < id_hi ? uid1 - id_lo : -1;
    uid2 = uid2 >= id_lo && uid2 < id_hi ? uid2 - id_lo : -1;
    int4 *vp = visits + 4 * at;
    if (keep == 2) {
      vp[0] = make_int4(x_in, p_in, E2.y, Q2.x);
      vp[1] = make_int4(Q2.y, E1.y, Q1.x, Q1.y);
      vp[2] = make_int4(uid2, uid1, 1, push);
    } else {
      vp[0] = make_int4(x_in, p_in, E1.y, Q1.x);
      vp[1] = make_int4(Q1.y, E2.y, Q2.x, Q2.y);
      vp[2] = make_int4(uid1, uid2, keep, push);
    }
    vp[3] = make_int4(0, 0, 0, 0);
    at++;
    if (acct) {
      const double o1 = E1.y < n_rows ? 1.0 : Cb, o2 = E2.y < n_rows ? 1.0 : Cb;
      const double xin = x_in >= 0 ? (x_in < n_rows ? 1.0 : Cb) : Cb;
      const bool n1 = uid1 >= 0 || sub1 > 0, n2 = uid2 >= 0 || sub2 > 0;
      if (n1) bytes += Cb + xin + o2 + (uid1 >= 0 ? Cb + o1 + o_tv + 8.0 : 0.0);
      if (n2) bytes += Cb + xin + o1 + (uid2 >= 0 ? Cb + o2 + o_tv + 8.0 : 0.0);
      req += (x_in >= 0 ? xin : (x_in <= -2 ? Cb : 0.0)) + o1 + o2 + (push >= 0 ? Cb : 0.0) +
             (uid1 >= 0 ? 0.25 : 0.0) + (uid2 >= 0 ? 0.25 : 0.0);
    }
  }
  if (acct) {
    atomicAdd(acct, bytes);
    atomicAdd(acct + 1, req);
  }
}

// B fragments of the DMMA search walk (lik_walkm.cuh): for every branch length t the pair
// (P(t), P(t / 2)) -- a target branch and its half (rearrangement.py:466-470) -- interleaved the way
// mma.m8n8k4 wants its B operand: lane 4 g + q holds (g odd ? P(t/2) : P(t))[r][g >> 1][q] for the four
// categories r, side by side, so a lane fetches a whole fragment with ONE 32-byte load
// (pfrag[pm][lane][r]).  Same arithmetic as pmatrix_kernel, entry by entry.
__global__ void __launch_bounds__(128)
pairfrag4_kernel(const ModelDev *__restrict__ model, const double *__restrict__ brlen, double *__restrict__ pfrag) {
  const int pm = blockIdx.x, lane = threadIdx.x >> 2, r = threadIdx.x & 3;
  const int g = lane >> 2, j = lane & 3, i = g >> 1;
  const double t = (g & 1) ? brlen[pm] * 0.5 : brlen[pm];
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) s += model->U[i * 4 + k] * exp(model->eval[k] * t * model->rates[r]) * model->Uinv[k * 4 + j];
  s = fmax(s, 0.0);
  if (t == 0.0) s = (i == j) ? 1.0 : 0.0;
  pfrag[((int64_t)pm * 32 + lane) * 4 + r] = s;
}

// DNA x Gamma-4 path: thread-per-pattern kernels, evaluations fused into the CLV update that
// produces their first operand whenever the other operands are ready by then.
static int64_t dna4_partial_stride(const LikAln &a) { return a.Ppad / D4_THREADS; }

static void run_dna4(const LikAln &a, const Model &m, LikWorkspace &ws, const LikProgram &prog, cudaStream_t st,
                     int32_t *out_moves) {
  constexpr int K = 4, R = 4, KKP = 17, NC = 16;
  StageTimer tm;
  const int n_pm = (int)prog.brlens.size();
  ws.pmat.alloc(std::max<size_t>(1, (size_t)n_pm * R * KKP));
  ws.tiptab.alloc(std::max<size_t>(1, (size_t)n_pm * NC * R * K));
  ws.brlen.alloc(std::max<size_t>(1, n_pm));
  if (n_pm) {
    PLG_CUDA(cudaMemcpyAsync(ws.brlen.p, prog.brlens.data(), (size_t)n_pm * 8, cudaMemcpyHostToDevice, st));
    prof_begin(PROF_PMATRIX, st);
    pmatrix_kernel<K, R><<<n_pm, 128, 0, st>>>(m.d.p, ws.brlen.p, a.masks.p, ws.pmat.p, ws.tiptab.p);
    count_launch();
    prof_end(PROF_PMATRIX, st, (double)n_pm * (R * KKP + NC * R * K) * 8.0, n_pm);
  }
  tm.lap("    pmatrix");
  // ---- fusion pass (host) ----
  const int64_t n_ops = (int64_t)prog.ops.size();
  std::vector<LikOpF> fops(n_ops);
  std::vector<int> producer(prog.n_slots, -1), op_level(n_ops, 0);
  for (size_t l = 0; l + 1 < prog.level_off.size(); l++)
    for (int64_t o = prog.level_off[l]; o < prog.level_off[l + 1]; o++) op_level[o] = (int)l;
  for (int64_t o = 0; o < n_ops; o++) {
    const LikOp &op = prog.ops[o];
    LikOpF f;
    f.dst = op.dst; f.a = op.a; f.b = op.b; f.pa = op.pa; f.pb = op.pb;
    f.tree = -1; f.e_pa = f.e_b = f.e_pb = f.e_c = f.e_pc = -1; f.pad = 0;
    fops[o] = f;
    producer[op.dst - a.n_rows] = (int)o;
  }
  auto level_of = [&](int id) { return id < a.n_rows ? -1 : op_level[producer[id - a.n_rows]]; };
  std::vector<LikEval> rest;
  for (const LikEval &ev : prog.evals) {
    bool fused = false;
    if (ev.a >= a.n_rows && ev.pa >= 0 && ev.c >= 0 && ev.pb >= 0 && ev.pc >= 0) {
      const int o = producer[ev.a - a.n_rows];
      if (fops[o].tree < 0 && level_of(ev.b) < op_level[o] && level_of(ev.c) < op_level[o]) {
        fops[o].tree = ev.tree; fops[o].e_pa = ev.pa;
        fops[o].e_b = ev.b; fops[o].e_pb = ev.pb; fops[o].e_c = ev.c; fops[o].e_pc = ev.pc;
        fused = true;
      }
    }
    if (!fused) rest.push_back(ev);
  }
  // dead-store elimination: a CLV nobody reads (its only consumer was the evaluation fused into
  // the op that produces it) stays in registers
  {
    std::vector<int> readers(prog.n_slots, 0);
    auto rd = [&](int id) { if (id >= a.n_rows) readers[id - a.n_rows]++; };
    for (const LikOpF &f : fops) { rd(f.a); if (f.b >= 0) rd(f.b); if (f.tree >= 0) { rd(f.e_b); rd(f.e_c); } }
    for (const LikEval &ev : rest) { rd(ev.a); rd(ev.b); if (ev.c >= 0) rd(ev.c); }
    for (LikOpF &f : fops)
      if (f.tree >= 0 && readers[f.dst - a.n_rows] == 0) f.pad = 1;
  }
  ws.fops.alloc(std::max<size_t>(1, (size_t)n_ops * sizeof(LikOpF)));
  if (n_ops)
    PLG_CUDA(cudaMemcpyAsync(ws.fops.p, fops.data(), (size_t)n_ops * sizeof(LikOpF), cudaMemcpyHostToDevice, st));
  if (!rest.empty())
    PLG_CUDA(cudaMemcpyAsync(ws.evals.p, rest.data(), rest.size() * sizeof(LikEval), cudaMemcpyHostToDevice, st));
  tm.lap("    fuse+upload");
  const int64_t bpo = a.Ppad / DNA4_PATTERNS_PER_BLOCK;
  // finest block tiling among the kernels below (search walks reduce per 32-pattern slice)
  const int walk_patterns = WM_PATTERNS;
  const bool device_plan = prog.spr_n_visits > 0;
  const bool walks = device_plan || !prog.walk_visits.empty();
  const int64_t pstride = !walks ? dna4_partial_stride(a) : a.Ppad / walk_patterns;
  if (prog.n_trees) PLG_CUDA(cudaMemsetAsync(ws.partial.p, 0, (size_t)prog.n_trees * pstride * 8, st));
  if (prog.n_trees && walks) {  // (mantissa, exponent) records of the walk: untouched entries read as the finished sum 0
    ws.partial_e.alloc((size_t)prog.n_trees * pstride);
    PLG_CUDA(cudaMemsetAsync(ws.partial_e.p, 0x80, (size_t)prog.n_trees * pstride * 4, st));
  }
  const double Cb = (double)R * K * 8.0 + 4.0;
  auto opnd = [&](int id) { return id < 0 ? 0.0 : (id < a.n_rows ? 1.0 : Cb); };
  const LikOpF *d_ops = reinterpret_cast<const LikOpF *>(ws.fops.p);
  for (size_t l = 0; l + 1 < prog.level_off.size(); l++) {
    const int64_t o0 = prog.level_off[l], cnt = prog.level_off[l + 1] - o0;
    const int64_t max_ops = std::max<int64_t>(1, ((int64_t)1 << 30) / bpo);
    for (int64_t o = 0; o < cnt; o += max_ops) {
      const int64_t c = std::min(max_ops, cnt - o);
      // algorithmic bytes (per node-op): CLV update = result + operands, evaluation =
      // its three operands + weight; C+4 per inner CLV, 1 per tip code.  Requested = what is moved
      // after fusion (the partial is not re-read) and dead-store elimination.
      double bytes = 0, req = 0;
      for (int64_t q = o0 + o; q < o0 + o + c; q++) {
        const LikOpF &f = fops[q];
        bytes += Cb + opnd(f.a) + opnd(f.b);
        req += (f.pad ? 0.0 : Cb) + opnd(f.a) + opnd(f.b);
        if (f.tree >= 0) {
          bytes += Cb + opnd(f.e_b) + opnd(f.e_c) + 8.0;
          req += opnd(f.e_b) + opnd(f.e_c) + 8.0;
        }
      }
      prof_begin(PROF_CLV_UPDATE, st);
      clv4_kernel<<<(unsigned)(c * bpo), D4_THREADS, 0, st>>>(d_ops + o0 + o, (int)bpo, a.n_rows, a.Ppad, a.codes.p,
                                                             a.masks.p, a.weights.p, m.d.p, ws.pmat.p, ws.tiptab.p,
                                                             ws.clv.p, ws.scl.p, ws.partial.p, (int)pstride);
      count_launch();
      prof_end(PROF_CLV_UPDATE, st, bytes * (double)a.P, (double)c * (double)a.P, req * (double)a.P);
    }
  }
  const int64_t n_ev = (int64_t)rest.size();
  const int64_t max_ev = std::max<int64_t>(1, ((int64_t)1 << 30) / bpo);
  for (int64_t e = 0; e < n_ev; e += max_ev) {
    const int64_t c = std::min(max_ev, n_ev - e);
    double bytes = 0;
    for (int64_t q = e; q < e + c; q++) bytes += 8.0 + opnd(rest[q].a) + opnd(rest[q].b) + opnd(rest[q].c);
    prof_begin(PROF_ROOT_EVAL, st);
    const int64_t ebpo = a.Ppad / D4_THREADS;
    eval4_kernel<<<(unsigned)(c * ebpo), D4_THREADS, 0, st>>>(ws.evals.p + e, (int)ebpo, a.n_rows, a.Ppad, a.codes.p,
                                                             a.masks.p, a.weights.p, m.d.p, ws.pmat.p, ws.tiptab.p,
                                                             ws.clv.p, ws.scl.p, ws.partial.p, (int)pstride);
    count_launch();
    prof_end(PROF_ROOT_EVAL, st, bytes * (double)a.P, (double)c * (double)a.P);
  }
  tm.lap("    level launches");
  // depth-first search walks: one persistent launch, a warp per (pattern slice, pruned subtree)
  if (walks) {
    int dev = 0, n_sm = 148;
    PLG_CUDA(cudaGetDevice(&dev));
    PLG_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
    const int n_blocks = n_sm * PLG_WM_MINB;
    const int64_t n_warps = (int64_t)n_blocks * WM_WARPS;
    const int levels = std::max(1, prog.walk_levels);
    ws.walk_visits.alloc(device_plan ? (size_t)prog.spr_n_visits : prog.walk_visits.size());
    ws.walk_groups.alloc(prog.walk_groups.size());
    ws.walk_stack.alloc((size_t)n_warps * levels * 16 * 32);
    ws.walk_stack_scl.alloc((size_t)n_warps * levels * 4 * 32);
    ws.walk_counter.alloc(1);
    PLG_CUDA(cudaMemcpyAsync(ws.walk_groups.p, prog.walk_groups.data(), prog.walk_groups.size() * sizeof(LikWalkGroup),
                             cudaMemcpyHostToDevice, st));
    PLG_CUDA(cudaMemsetAsync(ws.walk_counter.p, 0, sizeof(unsigned long long), st));
    ws.pfrag.alloc((size_t)std::max(1, n_pm) * 128);
    if (n_pm) {
      pairfrag4_kernel<<<n_pm, 128, 0, st>>>(m.d.p, ws.brlen.p, ws.pfrag.p);
      count_launch();
    }
    // accounting: algorithmic = per neighbour one CLV update (result + 2 operands) and one 3-way
    // evaluation (3 operands + weight), as for every other planner; requested = what the walk
    // really loads and stores (far-side operands, first partial of a side, stack traffic, Tv once
    // per search)
    double bytes = 0, req = 0, units = 0;
    const bool acct_on = device_plan && prof_enabled();
    if (device_plan) {
      // the visit list is written by spr_lik_plan_kernel from O(n) records, not uploaded
      ws.spr_edges.alloc(prog.spr_edges.size() * 2);
      ws.spr_searches.alloc(prog.spr_searches.size() * 3);
      if (out_moves) ws.spr_moves.alloc((size_t)prog.spr_n_unique);
      ws.spr_acct.alloc(2);
      PLG_CUDA(cudaMemcpyAsync(ws.spr_edges.p, prog.spr_edges.data(), prog.spr_edges.size() * sizeof(LikSprEdge),
                               cudaMemcpyHostToDevice, st));
      PLG_CUDA(cudaMemcpyAsync(ws.spr_searches.p, prog.spr_searches.data(),
                               prog.spr_searches.size() * sizeof(LikSprSearch), cudaMemcpyHostToDevice, st));
      if (acct_on) PLG_CUDA(cudaMemsetAsync(ws.spr_acct.p, 0, 2 * sizeof(double), st));
      const int n_search = (int)prog.spr_searches.size();
      const size_t edge_bytes = prog.spr_edges.size() * sizeof(LikSprEdge);
      const bool in_smem = edge_bytes <= SPR_LIK_PLAN_SMEM_MAX && !getenv("PLG_SPR_PLAN_GLOBAL");  // (tests: the fallback for huge trees)
      static bool plan_attr_set[64] = {false};
      ensure_dyn_smem(spr_lik_plan_kernel, (int)SPR_LIK_PLAN_SMEM_MAX, plan_attr_set);
      spr_lik_plan_kernel<<<(n_search + 63) / 64, 64, in_smem ? edge_bytes : 0, st>>>(
          ws.spr_edges.p, in_smem ? (int)(edge_bytes / 16) : 0, ws.spr_searches.p, n_search, a.n_rows,
          (int)prog.spr_lo, (int)prog.spr_hi, reinterpret_cast<int4 *>(ws.walk_visits.p),
          out_moves ? ws.spr_moves.p : nullptr,
          acct_on ? ws.spr_acct.p : nullptr);
      count_launch();
      if (out_moves)
        PLG_CUDA(cudaMemcpyAsync(out_moves, ws.spr_moves.p, (size_t)prog.spr_n_unique * sizeof(int4),
                                 cudaMemcpyDeviceToHost, st));
      // (profiling: the accounting sums come back with the results, run_lik_program patches the record)
      units = 2.0 * (double)prog.spr_n_visits;
    } else {
      PLG_CUDA(cudaMemcpyAsync(ws.walk_visits.p, prog.walk_visits.data(),
                               prog.walk_visits.size() * sizeof(LikWalkVisit), cudaMemcpyHostToDevice, st));
    }
    for (size_t gi = 0; !device_plan && gi < prog.walk_groups.size(); gi++) {
      const LikWalkGroup &g = prog.walk_groups[gi];
      req += opnd(g.tv) + 8.0;
      for (int q = g.first; q < g.first + g.count; q++) {
        const LikWalkVisit &v = prog.walk_visits[q];
        const double xin = v.x_in >= 0 ? opnd(v.x_in) : Cb;
        const bool n1 = v.tree1 >= 0 || v.keep == 1 || (v.push >= 0 && v.keep == 2);
        const bool n2 = v.tree2 >= 0 || v.keep == 2 || (v.push >= 0 && v.keep == 1);
        if (n1) bytes += Cb + xin + opnd(v.d2) + (v.tree1 >= 0 ? Cb + opnd(v.d1) + opnd(g.tv) + 8.0 : 0.0);
        if (n2) bytes += Cb + xin + opnd(v.d1) + (v.tree2 >= 0 ? Cb + opnd(v.d2) + opnd(g.tv) + 8.0 : 0.0);
        req += (v.x_in >= 0 ? opnd(v.x_in) : (v.x_in <= -2 ? Cb : 0.0)) + opnd(v.d1) + opnd(v.d2) +
               (v.push >= 0 ? Cb : 0.0) + (v.tree1 >= 0 ? 0.25 : 0.0) + (v.tree2 >= 0 ? 0.25 : 0.0);
        units += 2.0;
      }
    }
    const int64_t n_items = (a.Ppad / walk_patterns) * (int64_t)prog.walk_groups.size();
    prof_begin(PROF_WALK, st);
    {
      const int smem = WM_WARPS * WM_SMEM_DOUBLES_PER_WARP * 8;
      static bool wm_attr_set[64] = {false};
      ensure_dyn_smem(walk4m_kernel, smem, wm_attr_set);
      walk4m_kernel<<<n_blocks, WM_THREADS, smem, st>>>(ws.walk_groups.p, (int)prog.walk_groups.size(), ws.walk_visits.p,
                                                       n_items, a.n_rows, a.Ppad, a.codes.p, a.weights.p, m.d.p,
                                                       ws.pfrag.p, ws.clv.p, ws.scl.p,
                                                       ws.walk_stack.p, ws.walk_stack_scl.p, levels, ws.partial.p,
                                                       ws.partial_e.p,
                                                       (int)pstride, ws.walk_counter.p);
    }
    count_launch();
    prof_end(PROF_WALK, st, bytes * (double)a.P, units * (double)a.P, req * (double)a.P);
    tm.lap("    walk launch");
  }
  if (prog.n_trees) {
    if (walks)
      reduce_me_kernel<<<(unsigned)((prog.n_trees + 7) / 8), 256, 0, st>>>((int)prog.n_trees, (int)pstride, ws.partial.p,
                                                                           ws.partial_e.p, ws.lnl.p);
    else
      reduce_partials_kernel<<<(unsigned)((prog.n_trees + 7) / 8), 256, 0, st>>>((int)prog.n_trees, (int)pstride,
                                                                                 ws.partial.p, ws.lnl.p);
    count_launch();
  }
}

// n independent 20-state updates: the staged kernel (lik_tbr20.cuh clv20s_kernel), or with PLG_A20_STAGED=0 /
// row offsets beyond 32 bits the unstaged clv20_kernel (A/B reference)
static void launch_clv20(const LikAln &a, const LikOp *d_ops, int64_t n, const double *d_pmat, double *d_clv, int *d_scl,
                         cudaStream_t st) {
  const int64_t bpo = a.Ppad / A20_TILE;
  const char *e = getenv("PLG_A20_STAGED");
  const bool staged = !(e && !strcmp(e, "0"));
  if (staged && a.Ppad * 20 <= (int64_t)0x7fffffff) {
    static bool attr_set[64] = {false};
    ensure_dyn_smem(clv20s_kernel, T20_DYN_SMEM, attr_set);
    clv20s_kernel<<<(unsigned)(n * bpo), A20_THREADS, T20_DYN_SMEM, st>>>(d_ops, (int)n, a.n_rows, a.Ppad, a.codes.p,
                                                                         a.masks.p, d_pmat, d_clv, d_scl);
  } else {
    clv20_kernel<<<(unsigned)(n * bpo), A20_THREADS, 0, st>>>(d_ops, (int)bpo, a.n_rows, a.Ppad, a.codes.p, a.masks.p,
                                                             d_pmat, d_clv, d_scl);
  }
}

// 20 states x Gamma-4 path: DMMA kernels (lik_aa.cuh).  Tips enter the tensor-core product as
// indicator columns, so only the P-matrices are needed (the tip tables are still produced by
// pmatrix_kernel for the generic path).
static void run_aa20(const LikAln &a, const Model &m, LikWorkspace &ws, const LikProgram &prog, cudaStream_t st) {
  constexpr int K = 20, R = 4, KKP = 401, NC = 23;
  const int n_pm = (int)prog.brlens.size();
  ws.pmat.alloc(std::max<size_t>(1, (size_t)n_pm * R * KKP));
  ws.tiptab.alloc(std::max<size_t>(1, (size_t)n_pm * NC * R * K));
  ws.brlen.alloc(std::max<size_t>(1, n_pm));
  if (n_pm) {
    PLG_CUDA(cudaMemcpyAsync(ws.brlen.p, prog.brlens.data(), (size_t)n_pm * 8, cudaMemcpyHostToDevice, st));
    prof_begin(PROF_PMATRIX, st);
    pmatrix_kernel<K, R><<<n_pm, 128, 0, st>>>(m.d.p, ws.brlen.p, a.masks.p, ws.pmat.p, ws.tiptab.p);
    count_launch();
    prof_end(PROF_PMATRIX, st, (double)n_pm * (R * KKP + NC * R * K) * 8.0, n_pm);
  }
  const int64_t bpo = a.Ppad / A20_TILE;
  const double Cb = (double)R * K * 8.0 + 4.0;
  auto opnd = [&](int id) { return id < 0 ? 0.0 : (id < a.n_rows ? 1.0 : Cb); };
  for (size_t l = 0; l + 1 < prog.level_off.size(); l++) {
    const int64_t o0 = prog.level_off[l], cnt = prog.level_off[l + 1] - o0;
    const int64_t max_ops = std::max<int64_t>(1, ((int64_t)1 << 30) / bpo);
    for (int64_t o = 0; o < cnt; o += max_ops) {
      const int64_t c = std::min(max_ops, cnt - o);
      double bytes = 0;
      for (int64_t q = o0 + o; q < o0 + o + c; q++) bytes += Cb + opnd(prog.ops[q].a) + opnd(prog.ops[q].b);
      prof_begin(PROF_CLV_UPDATE, st);
      launch_clv20(a, ws.ops.p + o0 + o, c, ws.pmat.p, ws.clv.p, ws.scl.p, st);
      count_launch();
      prof_end(PROF_CLV_UPDATE, st, bytes * (double)a.P, (double)c * (double)a.P);
    }
  }
  if (!prog.t20_ops.empty()) {  // TBR neighbourhood: fused candidate steps, then the pair grid (lik_tbr20.cuh)
    for (size_t l = 0; l + 1 < prog.t20_level_off.size(); l++) {
      const int64_t o0 = prog.t20_level_off[l], cnt = prog.t20_level_off[l + 1] - o0;
      if (!cnt) continue;
      const double share = (double)cnt / (double)prog.t20_ops.size();  // (per-level split of the program totals)
      static bool t20_attr_set[64] = {false};
      ensure_dyn_smem(tbr20_step_kernel, T20_DYN_SMEM, t20_attr_set);
      prof_begin(PROF_CLV_UPDATE, st);
      tbr20_step_kernel<<<(unsigned)(cnt * bpo), A20_THREADS, T20_DYN_SMEM, st>>>(ws.t20_ops.p + o0, (int)cnt, a.n_rows, a.Ppad,
                                                                      a.codes.p, a.masks.p, m.d.p, ws.pmat.p, ws.clv.p,
                                                                      ws.scl.p, ws.pairvec.p, ws.pair_scl.p);
      count_launch();
      // units: node-ops of 2 products each (what FLOP_PER_NODEOP_AA counts)
      prof_end(PROF_CLV_UPDATE, st, share * prog.t20_bytes * (double)a.P, share * 0.5 * prog.t20_products * (double)a.P);
    }
    const int n_edges = (int)prog.t20_tile_off.size() - 1, n_chunks = (int)(a.Ppad / P20_CHUNK);
    const int64_t n_items = (int64_t)prog.t20_tiles.size() * n_chunks;
    if (n_items) {
      prof_begin(PROF_ROOT_EVAL, st);
      pairs20_kernel<<<(unsigned)((n_items + 3) / 4), 128, 0, st>>>(ws.t20_tiles.p, ws.t20_tile_off.p, n_edges, n_chunks,
                                                                   n_items, a.Ppad, a.weights.p, ws.pairvec.p,
                                                                   ws.pair_scl.p, ws.partial.p);
      count_launch();
      prof_end(PROF_ROOT_EVAL, st, prog.t20_pair_bytes * (double)a.P, (double)prog.n_trees * (double)a.P);
    }
    if (prog.n_trees) {
      reduce_partials_kernel<<<(unsigned)((prog.n_trees + 7) / 8), 256, 0, st>>>((int)prog.n_trees, n_chunks,
                                                                                 ws.partial.p, ws.lnl.p);
      count_launch();
    }
    return;
  }
  int64_t n_ev = (int64_t)prog.evals.size();
  // evaluations without matrices whose first operand repeats (TBR pairs): one block keeps that
  // operand's tile in registers for the whole row (PLG_AA_EVAL_ROWS=0: the per-evaluation kernel)
  bool rows_ok = n_ev > 0;
  if (const char *e = getenv("PLG_AA_EVAL_ROWS")) rows_ok = rows_ok && atoi(e) != 0;
  for (int64_t q = 0; rows_ok && q < n_ev; q++)
    rows_ok = prog.evals[q].pa < 0 && prog.evals[q].pb < 0 && prog.evals[q].c < 0;
  if (rows_ok) {
    std::vector<LikEvalRow> rows;
    double bytes = 0, req = 0;
    for (int64_t q = 0; q < n_ev;) {
      int64_t j = q + 1;
      while (j < n_ev && prog.evals[j].a == prog.evals[q].a) j++;
      rows.push_back({prog.evals[q].a, (int)q, (int)(j - q), 0});
      req += opnd(prog.evals[q].a);
      for (int64_t i = q; i < j; i++) {
        bytes += 8.0 + opnd(prog.evals[i].a) + opnd(prog.evals[i].b);
        req += 8.0 + opnd(prog.evals[i].b);
      }
      q = j;
    }
    DevBuf<int4> &d_rows = ws.eval_rows;  // (per-device workspace)
    d_rows.alloc(rows.size());
    PLG_CUDA(cudaMemcpyAsync(d_rows.p, rows.data(), rows.size() * sizeof(LikEvalRow), cudaMemcpyHostToDevice, st));
    prof_begin(PROF_ROOT_EVAL, st);
    eval20_rows_kernel<<<(unsigned)(rows.size() * bpo), A20_THREADS, 0, st>>>(
        reinterpret_cast<const LikEvalRow *>(d_rows.p), (int)rows.size(), ws.evals.p, (int)bpo, a.n_rows, a.Ppad,
        a.codes.p, a.masks.p, a.weights.p, m.d.p, ws.clv.p, ws.scl.p, ws.partial.p);
    count_launch();
    prof_end(PROF_ROOT_EVAL, st, bytes * (double)a.P, (double)n_ev * (double)a.P, req * (double)a.P);
    PLG_CUDA(cudaStreamSynchronize(st));  // `rows` is pageable host memory
    n_ev = 0;
  }
  const int64_t max_ev = std::max<int64_t>(1, ((int64_t)1 << 30) / bpo);
  for (int64_t e = 0; e < n_ev; e += max_ev) {
    const int64_t c = std::min(max_ev, n_ev - e);
    double bytes = 0;
    for (int64_t q = e; q < e + c; q++)
      bytes += 8.0 + opnd(prog.evals[q].a) + opnd(prog.evals[q].b) + opnd(prog.evals[q].c);
    prof_begin(PROF_ROOT_EVAL, st);
    eval20_kernel<<<(unsigned)(c * bpo), A20_THREADS, 0, st>>>(ws.evals.p + e, (int)bpo, a.n_rows, a.Ppad, a.codes.p,
                                                              a.masks.p, a.weights.p, m.d.p, ws.pmat.p, ws.clv.p,
                                                              ws.scl.p, ws.partial.p);
    count_launch();
    prof_end(PROF_ROOT_EVAL, st, bytes * (double)a.P, (double)c * (double)a.P);
  }
  if (prog.n_trees) {
    reduce_partials_kernel<<<(unsigned)((prog.n_trees + 7) / 8), 256, 0, st>>>((int)prog.n_trees, (int)bpo,
                                                                               ws.partial.p, ws.lnl.p);
    count_launch();
  }
}

void run_lik_program(const LikAln &a, const Model &m, LikWorkspace &ws, const LikProgram &prog, double *out_lnl,
                     cudaStream_t st, int32_t *out_moves) {
  if (m.K != a.K) PLG_FAIL(PLG_ERR_ARG, "model has %d states but the alignment has %d", m.K, a.K);
  const size_t RK = (size_t)m.R * a.K;
  ws.clv.alloc(std::max<size_t>(1, (size_t)prog.n_slots * RK * a.Ppad));
  ws.scl.alloc(std::max<size_t>(1, (size_t)prog.n_slots * a.Ppad));
  ws.ops.alloc(std::max<size_t>(1, prog.ops.size()));
  ws.evals.alloc(std::max<size_t>(1, prog.evals.size()));
  const bool dna4 = (a.K == 4 && m.R == 4);
  const int64_t bpo = !dna4 ? a.Ppad / (LIK_THREADS / m.R)
                             : (prog.walk_visits.empty() && prog.spr_n_visits == 0 ? dna4_partial_stride(a)
                                                                                    : a.Ppad / WM_PATTERNS);  // walk slicing
  const bool t20 = !prog.t20_ops.empty();
  ws.partial.alloc(std::max<size_t>(1, (size_t)prog.n_trees * (size_t)(t20 ? a.Ppad / P20_CHUNK : bpo)));
  ws.lnl.alloc(std::max<size_t>(1, (size_t)prog.n_trees));
  if (t20) {
    if (!(a.K == 20 && m.R == 4)) PLG_FAIL(PLG_ERR_ARG, "internal: TBR-20 program on a %d-state, %d-category model", a.K, m.R);
    if (a.Ppad * 20 > (int64_t)0x7fffffff) PLG_FAIL(PLG_ERR_UNSUPPORTED, "TBR-20 kernels index a vector's rows with 32 bits: %lld patterns are too many", (long long)a.P);
    ws.pairvec.alloc(std::max<size_t>(1, (size_t)prog.t20_pair_slots * 80 * a.Ppad));
    ws.pair_scl.alloc(std::max<size_t>(1, (size_t)prog.t20_pair_slots * a.Ppad));
    ws.t20_ops.alloc(prog.t20_ops.size());
    ws.t20_tiles.alloc(std::max<size_t>(1, prog.t20_tiles.size()));
    ws.t20_tile_off.alloc(prog.t20_tile_off.size());
    PLG_CUDA(cudaMemcpyAsync(ws.t20_ops.p, prog.t20_ops.data(), prog.t20_ops.size() * sizeof(Tbr20Op), cudaMemcpyHostToDevice, st));
    if (!prog.t20_tiles.empty())
      PLG_CUDA(cudaMemcpyAsync(ws.t20_tiles.p, prog.t20_tiles.data(), prog.t20_tiles.size() * sizeof(Tbr20Tile),
                               cudaMemcpyHostToDevice, st));
    PLG_CUDA(cudaMemcpyAsync(ws.t20_tile_off.p, prog.t20_tile_off.data(), prog.t20_tile_off.size() * sizeof(int),
                             cudaMemcpyHostToDevice, st));
  }
  if (!dna4 && !prog.ops.empty())
    PLG_CUDA(cudaMemcpyAsync(ws.ops.p, prog.ops.data(), prog.ops.size() * sizeof(LikOp), cudaMemcpyHostToDevice, st));
  if (!dna4 && !prog.evals.empty())
    PLG_CUDA(cudaMemcpyAsync(ws.evals.p, prog.evals.data(), prog.evals.size() * sizeof(LikEval),
                             cudaMemcpyHostToDevice, st));
  PLG_CUDA(cudaMemsetAsync(ws.lnl.p, 0, std::max<size_t>(1, (size_t)prog.n_trees) * 8, st));
  if (dna4) run_dna4(a, m, ws, prog, st, out_moves);
  else if (a.K == 4 && m.R == 1) run_typed<4, 1>(a, m, ws, prog, st);
  else if (a.K == 20 && m.R == 4) run_aa20(a, m, ws, prog, st);
  else if (a.K == 20 && m.R == 1) run_typed<20, 1>(a, m, ws, prog, st);
  else PLG_FAIL(PLG_ERR_UNSUPPORTED, "kernels exist for 4 or 20 states with 1 or 4 rate categories");
  PLG_CUDA(cudaGetLastError());
  if (out_lnl && prog.n_trees) {
    StageTimer tm;
    PLG_CUDA(cudaMemcpyAsync(out_lnl, ws.lnl.p, (size_t)prog.n_trees * 8, out_copy_kind(), st));
    const bool acct = dna4 && prog.spr_n_visits > 0 && prof_enabled();
    double h[2] = {0.0, 0.0};
    if (acct) PLG_CUDA(cudaMemcpyAsync(h, ws.spr_acct.p, sizeof(h), cudaMemcpyDeviceToHost, st));
    PLG_CUDA(cudaStreamSynchronize(st));
    if (acct) {  // byte accounting of the device-planned walk (per pattern) -> its profile record
      const double Cb = 4 * 4 * 8.0 + 4.0;
      for (const LikWalkGroup &g : prog.walk_groups) h[1] += (g.tv < a.n_rows ? 1.0 : Cb) + 8.0;
      prof_patch_last(PROF_WALK, h[0] * (double)a.P, 2.0 * (double)prog.spr_n_visits * (double)a.P, h[1] * (double)a.P);
    }
    tm.lap("    sync+D2H");
  }
}

void dna4_pmatrices(const LikAln &a, const Model &m, const double *d_brlen, const int *d_idx, int64_t n,
                    double *d_pmat, double *d_tiptab, cudaStream_t st) {
  if (n <= 0) return;
  prof_begin(PROF_PMATRIX, st);
  pmatrix_kernel<4, 4><<<(unsigned)n, 128, 0, st>>>(m.d.p, d_brlen, a.masks.p, d_pmat, d_tiptab, d_idx);
  count_launch();
  prof_end(PROF_PMATRIX, st, (double)n * (4 * 17 + 16 * 16) * 8.0, (double)n);
}

void dna4_updates(const LikAln &a, const Model &m, const LikOpF *d_ops, int64_t n, const double *d_pmat,
                  const double *d_tiptab, double *d_clv, int *d_scl, cudaStream_t st) {
  if (n <= 0) return;
  const int64_t bpo = a.Ppad / DNA4_PATTERNS_PER_BLOCK;
  prof_begin(PROF_CLV_UPDATE, st);
  clv4_kernel<<<(unsigned)(n * bpo), D4_THREADS, 0, st>>>(d_ops, (int)bpo, a.n_rows, a.Ppad, a.codes.p, a.masks.p,
                                                         a.weights.p, m.d.p, d_pmat, d_tiptab, d_clv, d_scl, nullptr, 0);
  count_launch();
  // (operand kinds are not known here: counted as inner/inner, 3C + 12 per pattern)
  prof_end(PROF_CLV_UPDATE, st, (double)n * 3.0 * 132.0 * (double)a.P, (double)n * (double)a.P);
}

// Fused re-orientation + sum-table pass (lik.h ClvOptOp).  Block = 128 threads x 2 tiles of 128 patterns: the
// tile / block mapping, the arithmetic and the order of the sums are those of opt_step4_kernel (pll_api.cu),
// so the derivative-sum shares are bit-identical to the unfused pair of launches.
constexpr double PLLF_ZMIN = 1.0e-15, PLLF_ZMAX = 1.0 - 1.0e-6;
__global__ void __launch_bounds__(D4_THREADS, 4)
clvopt4_kernel(const ClvOptOp *__restrict__ ops, int blocks_per_op, int n_rows, int64_t Ppad,
               const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
               const double *__restrict__ weights, const Eigen4 eg, const double *__restrict__ pmat,
               const double *__restrict__ tiptab, double *__restrict__ clv, int *__restrict__ scl,
               const double *__restrict__ brlen, const int *__restrict__ active, double *__restrict__ dpart) {
  static_assert(D4_THREADS == 128, "tile mapping of opt_step4_kernel");
  __shared__ __align__(16) double smA[256];
  __shared__ __align__(16) double smB[256];
  __shared__ double tab[3][16];
  __shared__ double red[3][D4_THREADS / 32];
  const int op_idx = blockIdx.x / blocks_per_op;
  const int blk = blockIdx.x - op_idx * blocks_per_op;
  const ClvOptOp op = ops[op_idx];
  const bool act = active[op.tree] != 0;
  d4_stage(smA, op.a < n_rows, op.pa, pmat, tiptab);
  d4_stage(smB, op.b < n_rows, op.pb, pmat, tiptab);
  if (act && threadIdx.x < 16) {
    const double z = fmin(fmax(exp(-brlen[op.branch]), PLLF_ZMIN), PLLF_ZMAX);
    const double lz = log(fmax(z, PLLF_ZMIN));
    const double g = eg.g[threadIdx.x];
    const double e = exp(g * lz);
    tab[0][threadIdx.x] = e; tab[1][threadIdx.x] = g * e; tab[2][threadIdx.x] = g * g * e;
  }
  __syncthreads();
  const bool tipO = op.other < n_rows;
  double v0 = 0.0, v1 = 0.0, v2 = 0.0;
#pragma unroll 1
  for (int tl = 0; tl < 2; tl++) {
    const int64_t p = ((int64_t)blk * 2 + tl) * D4_THREADS + threadIdx.x;
    double v[16];
    int sc = 0;
    d4_operand<false>(smA, op.a, op.pa, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
    d4_operand<true>(smB, op.b, op.pb, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) mx = fmax(mx, fabs(v[i]));
    if (mx < MINLH) {
#pragma unroll
      for (int i = 0; i < 16; i++) v[i] *= TWO256;
      sc += 1;
    }
    double *dst = clv + (int64_t)(op.dst - n_rows) * 16 * Ppad + p;
#pragma unroll
    for (int i = 0; i < 16; i++) dst[(int64_t)i * Ppad] = v[i];
    scl[(int64_t)(op.dst - n_rows) * Ppad + p] = sc;
    if (!act) continue;
    double a[16];
    if (tipO) {
      const unsigned int m = masks[codes[(int64_t)op.other * Ppad + p]];
#pragma unroll
      for (int e = 0; e < 16; e++) a[e] = (m >> (e & 3) & 1) ? 1.0 : 0.0;
    } else {
      const double *src = clv + (int64_t)(op.other - n_rows) * 16 * Ppad + p;
#pragma unroll
      for (int e = 0; e < 16; e++) a[e] = __ldg(src + (int64_t)e * Ppad);
      sc += __ldg(scl + (int64_t)(op.other - n_rows) * Ppad + p);
    }
    const double w = weights[p];
    double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
    for (int r = 0; r < 4; r++) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double l = a[r * 4] * eg.piU[k], rr = eg.Uinv[k * 4] * v[r * 4];
#pragma unroll
        for (int i = 1; i < 4; i++) {
          l = fma(a[r * 4 + i], eg.piU[i * 4 + k], l);
          rr = fma(eg.Uinv[k * 4 + i], v[r * 4 + i], rr);
        }
        const double sv = l * rr;
        l0 = fma(sv, tab[0][r * 4 + k], l0);
        l1 = fma(sv, tab[1][r * 4 + k], l1);
        l2 = fma(sv, tab[2][r * 4 + k], l2);
      }
    }
    if (w != 0.0) {
      const double inv = 1.0 / l0, q1 = l1 * inv;
      v0 += w * (log(l0 * 0.25) + sc * LOG_MINLH);
      v1 += w * q1;
      v2 += w * (l2 * inv - q1 * q1);
    }
  }
  if (!act) return;
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
    for (int i = 0; i < D4_THREADS / 32; i++) t += red[threadIdx.x][i];
    dpart[((int64_t)op.dslot * 3 + threadIdx.x) * blocks_per_op + blk] = t;
  }
}

void dna4_updates_opt(const LikAln &a, const Model &m, const ClvOptOp *d_ops, int64_t n, const double *d_pmat,
                      const double *d_tiptab, double *d_clv, int *d_scl, const Eigen4 &eg, const double *d_brlen,
                      const int *d_active, double *d_dpart, cudaStream_t st) {
  if (n <= 0) return;
  (void)m;
  const int64_t bpo = a.Ppad / (2 * D4_THREADS);
  prof_begin(PROF_CLV_UPDATE, st);
  clvopt4_kernel<<<(unsigned)(n * bpo), D4_THREADS, 0, st>>>(d_ops, (int)bpo, a.n_rows, a.Ppad, a.codes.p, a.masks.p,
                                                            a.weights.p, eg, d_pmat, d_tiptab, d_clv, d_scl, d_brlen,
                                                            d_active, d_dpart);
  count_launch();
  prof_end(PROF_CLV_UPDATE, st, (double)n * 4.0 * 132.0 * (double)a.P, (double)n * (double)a.P);
}

void dna4_evals(const LikAln &a, const Model &m, const LikEval *d_evals, int64_t n, const double *d_pmat,
                const double *d_tiptab, const double *d_clv, const int *d_scl, double *d_partial, double *d_lnl,
                cudaStream_t st) {
  if (n <= 0) return;
  const int64_t ebpo = a.Ppad / D4_THREADS;
  PLG_CUDA(cudaMemsetAsync(d_partial, 0, (size_t)n * ebpo * 8, st));
  prof_begin(PROF_ROOT_EVAL, st);
  eval4_kernel<<<(unsigned)(n * ebpo), D4_THREADS, 0, st>>>(d_evals, (int)ebpo, a.n_rows, a.Ppad, a.codes.p, a.masks.p,
                                                           a.weights.p, m.d.p, d_pmat, d_tiptab, d_clv, d_scl, d_partial,
                                                           (int)ebpo);
  count_launch();
  prof_end(PROF_ROOT_EVAL, st, (double)n * (2.0 * 132.0 + 8.0) * (double)a.P, (double)n * (double)a.P);
  reduce_partials_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>((int)n, (int)ebpo, d_partial, d_lnl);
  count_launch();
}

void aa20_pmatrices(const LikAln &a, const Model &m, const double *d_brlen, const int *d_idx, int64_t n,
                    double *d_pmat, double *d_tiptab, cudaStream_t st) {
  if (n <= 0) return;
  prof_begin(PROF_PMATRIX, st);
  pmatrix_kernel<20, 4><<<(unsigned)n, 128, 0, st>>>(m.d.p, d_brlen, a.masks.p, d_pmat, d_tiptab, d_idx);
  count_launch();
  prof_end(PROF_PMATRIX, st, (double)n * (4 * 401 + 23 * 80) * 8.0, (double)n);
}

void aa20_updates(const LikAln &a, const Model &m, const LikOp *d_ops, int64_t n, const double *d_pmat,
                  double *d_clv, int *d_scl, cudaStream_t st) {
  if (n <= 0) return;
  prof_begin(PROF_CLV_UPDATE, st);
  launch_clv20(a, d_ops, n, d_pmat, d_clv, d_scl, st);
  count_launch();
  prof_end(PROF_CLV_UPDATE, st, (double)n * 3.0 * 644.0 * (double)a.P, (double)n * (double)a.P);
}

void aa20_evals(const LikAln &a, const Model &m, const LikEval *d_evals, int64_t n, const double *d_pmat,
                const double *d_clv, const int *d_scl, double *d_partial, double *d_lnl, cudaStream_t st) {
  if (n <= 0) return;
  const int64_t bpo = a.Ppad / A20_TILE;
  prof_begin(PROF_ROOT_EVAL, st);
  eval20_kernel<<<(unsigned)(n * bpo), A20_THREADS, 0, st>>>(d_evals, (int)bpo, a.n_rows, a.Ppad, a.codes.p, a.masks.p,
                                                            a.weights.p, m.d.p, d_pmat, d_clv, d_scl, d_partial);
  count_launch();
  prof_end(PROF_ROOT_EVAL, st, (double)n * (2.0 * 644.0 + 8.0) * (double)a.P, (double)n * (double)a.P);
  reduce_partials_kernel<<<(unsigned)((n + 7) / 8), 256, 0, st>>>((int)n, (int)bpo, d_partial, d_lnl);
  count_launch();
}

void score_lik_batch_full(LikAln &a, const Model &m, const TreeBatch &b, double *out_lnl, cudaStream_t st) {
  const int64_t n_trees = b.n_trees();
  if (n_trees == 0) return;
  StageTimer tm;
  size_t free_b = 0, total_b = 0;
  PLG_CUDA(cudaMemGetInfo(&free_b, &total_b));
  tm.lap("meminfo");
  const size_t slot_bytes = ((size_t)m.R * a.K * 8 + 4) * a.Ppad;
  const size_t held = lik_workspace().clv.n * 8 + lik_workspace().scl.n * 4;
  const size_t budget = std::min<size_t>((free_b + held) / 2, (size_t)64 << 30);
  const int64_t max_slots = std::max<int64_t>(1, (int64_t)(budget / slot_bytes));
  LikProgram prog;
  int64_t t0 = 0;
  while (t0 < n_trees) {
    // slots needed by a tree <= its child count (binary: one per non-root internal node)
    int64_t t1 = t0 + 1;
    auto need = [&](int64_t lo, int64_t hi) { return b.child_off[b.node_off[hi]] - b.child_off[b.node_off[lo]]; };
    while (t1 < n_trees && need(t0, t1 + 1) / 2 <= max_slots) t1++;
    build_full_lik_program(b, t0, t1, a.n_rows, prog);
    tm.lap("build_full");
    run_lik_program(a, m, lik_workspace(), prog, out_lnl + t0, st);
    tm.lap("run");
    t0 = t1;
  }
}

// Binary descriptor batches, DNA x Gamma-4: tree walks planned on the device (treewalk.h,
// lik_twalk.cuh).  Host work per call: validate, copy desc / brlen up, copy lnL down.
static bool lik_tree_walk_enabled() {
  const char *e = getenv("PLG_TREE_MODE");
  return !(e && !strcmp(e, "levels"));
}

void score_lik_desc_walk(LikAln &a, const Model &m, int64_t n_trees, int n_tips, const int32_t *desc,
                         const double *brlen, const int32_t *tip_rows, double *out_lnl, cudaStream_t st) {
  if (n_trees == 0) return;
  StageTimer tm;
  validate_desc(n_trees, n_tips, desc, tip_rows, a.n_rows, true);
  tm.lap("validate");
  LikWorkspace &ws = lik_workspace();
  const int ni = n_tips - 1;
  const int64_t tiles = a.Ppad / TW_THREADS;
  int64_t chunk = std::min<int64_t>(n_trees, ((int64_t)1 << 30) / ((int64_t)2 * ni * 512));
  chunk = std::min<int64_t>(chunk, ((int64_t)512 << 20) / (tiles * 8));
  if (const char *e = getenv("PLG_WALK_CHUNK")) chunk = std::min<int64_t>(chunk, atoll(e));  // tests: force several chunks
  chunk = std::max<int64_t>(1, chunk);
  int dev = 0, n_sm = 148;
  PLG_CUDA(cudaGetDevice(&dev));
  PLG_CUDA(cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev));
  const int n_blocks = n_sm * PLG_TW_MINB;
  const int levels = walk_levels(n_tips);
  ws.tw_desc.alloc((size_t)chunk * ni * 2);
  ws.brlen.alloc((size_t)chunk * ni * 2 + 1);
  ws.tw_ops.alloc((size_t)chunk * ni);
  ws.tw_scratch.alloc((size_t)chunk * (3 * (size_t)ni + 2));
  ws.tw_rows.alloc((size_t)n_tips);
  ws.pmat.alloc((size_t)chunk * 2 * ni * 64);
  ws.walk_stack.alloc((size_t)n_blocks * levels * 16 * TW_THREADS);
  ws.walk_stack_scl.alloc((size_t)n_blocks * levels * TW_THREADS);
  ws.walk_counter.alloc(1);
  ws.partial.alloc((size_t)chunk * tiles);
  ws.lnl.alloc((size_t)chunk);
  if (tip_rows)
    PLG_CUDA(cudaMemcpyAsync(ws.tw_rows.p, tip_rows, (size_t)n_tips * sizeof(int32_t), cudaMemcpyHostToDevice, st));
  const double Cb = 132.0;
  for (int64_t t0 = 0; t0 < n_trees; t0 += chunk) {
    const int64_t cnt = std::min(chunk, n_trees - t0);
    const int64_t n_slots = cnt * 2 * ni;
    PLG_CUDA(cudaMemcpyAsync(ws.tw_desc.p, desc + (size_t)t0 * ni * 2, (size_t)n_slots * sizeof(int32_t),
                             cudaMemcpyHostToDevice, st));
    PLG_CUDA(cudaMemcpyAsync(ws.brlen.p, brlen + (size_t)t0 * ni * 2, (size_t)n_slots * sizeof(double),
                             cudaMemcpyHostToDevice, st));
    PLG_CUDA(cudaMemsetAsync(ws.walk_counter.p, 0, sizeof(unsigned long long), st));
    launch_plan_walk(cnt, n_tips, ws.tw_desc.p, tip_rows ? ws.tw_rows.p : nullptr, a.n_rows,
                     reinterpret_cast<WalkOp *>(ws.tw_ops.p), ws.tw_scratch.p, st);
    prof_begin(PROF_PMATRIX, st);
    pmatrix4_slots_kernel<<<(unsigned)((n_slots + 3) / 4), 256, 0, st>>>(m.d.p, ws.brlen.p, n_slots, 2 * ni, ws.pmat.p);
    count_launch();
    prof_end(PROF_PMATRIX, st, (double)n_slots * 64 * 8.0, (double)n_slots);
    prof_begin(PROF_WALK, st);
    static bool tw_attr_set[64] = {false};
    ensure_dyn_smem(twalk4_kernel, TW_SMEM_BYTES, tw_attr_set);
    twalk4_kernel<<<n_blocks, TW_THREADS, TW_SMEM_BYTES, st>>>(reinterpret_cast<const WalkOp *>(ws.tw_ops.p), ni, (int)cnt,
                                                  (unsigned long long)(cnt * tiles), a.Ppad, a.codes.p, a.weights.p,
                                                  m.d.p, ws.pmat.p, ws.walk_stack.p, ws.walk_stack_scl.p, levels,
                                                  ws.partial.p, (int)tiles, ws.walk_counter.p);
    count_launch();
    // algorithmic bytes: every non-root node-op writes its CLV once and it is read
    // once, tips 1 byte, weight 8 bytes; requested: tips + weight (stack traffic stays in L2)
    prof_end(PROF_WALK, st, (double)cnt * (double)a.P * (2.0 * (ni - 1) * Cb + n_tips + 8.0),
             (double)cnt * ni * (double)a.P, (double)cnt * (double)a.P * (n_tips + 8.0));
    reduce_partials_kernel<<<(unsigned)((cnt + 7) / 8), 256, 0, st>>>((int)cnt, (int)tiles, ws.partial.p, ws.lnl.p);
    count_launch();
    PLG_CUDA(cudaGetLastError());
    PLG_CUDA(cudaMemcpyAsync(out_lnl + t0, ws.lnl.p, (size_t)cnt * 8, out_copy_kind(), st));
    PLG_CUDA(cudaStreamSynchronize(st));
    tm.lap("walk chunk");
  }
}

}  // namespace plg

// ----------------------------------------------------------------------------
// C-ABI
// ----------------------------------------------------------------------------
using namespace plg;

extern "C" int plg_lik_aln_create(int datatype, int n_rows, int64_t n_patterns, const uint8_t *codes,
                                  const double *weights, plg_lik_aln **out) {
  PLG_API_BEGIN_LOCKED
  if (!out || (n_patterns > 0 && (!codes || !weights))) PLG_FAIL(PLG_ERR_ARG, "null argument");
  *out = new plg_lik_aln(datatype, n_rows, n_patterns, codes, weights);
  PLG_API_END
}
extern "C" void plg_lik_aln_destroy(plg_lik_aln *aln) { delete aln; }

extern "C" int plg_model_create(int n_states, const double *exch, const double *freqs, double alpha, int n_cat,
                                plg_model **out) {
  PLG_API_BEGIN_LOCKED
  if (!out || !exch || !freqs) PLG_FAIL(PLG_ERR_ARG, "null argument");
  int dev_count = 0;
  if (cudaGetDeviceCount(&dev_count) != cudaSuccess || dev_count == 0)
    PLG_FAIL(PLG_ERR_CUDA, "no CUDA device: pylogeny_b200 has no CPU fallback");
  *out = new plg_model(n_states, exch, freqs, alpha, n_cat);
  PLG_API_END
}
extern "C" void plg_model_destroy(plg_model *m) { delete m; }

extern "C" int plg_model_gamma_rates(const plg_model *m, double *rates) {
  PLG_API_BEGIN
  if (!m || !rates) PLG_FAIL(PLG_ERR_ARG, "null argument");
  for (int i = 0; i < m->impl.R; i++) rates[i] = m->impl.h.rates[i];
  PLG_API_END
}

extern "C" int plg_model_pmatrix(const plg_model *m, double t, double *out) {
  PLG_API_BEGIN
  if (!m || !out) PLG_FAIL(PLG_ERR_ARG, "null argument");
  m->impl.pmatrix(t, out);
  PLG_API_END
}

static void lik_score_trees_impl(plg_lik_aln *aln, plg_model *m, int64_t n_trees, int n_tips, const int32_t *desc,
                                 const double *brlen, const int32_t *tip_rows, double *out_lnl, void *stream);

extern "C" int plg_lik_score_trees(plg_lik_aln *aln, plg_model *m, int64_t n_trees, int n_tips, const int32_t *desc,
                                   const double *brlen, const int32_t *tip_rows, double *out_lnl, void *stream) {
  PLG_API_BEGIN_LOCKED
  lik_score_trees_impl(aln, m, n_trees, n_tips, desc, brlen, tip_rows, out_lnl, stream);
  PLG_API_END
}

extern "C" int plg_lik_score_trees_shard(plg_lik_aln *aln, plg_model *m, int64_t n_trees, int n_tips,
                                         const int32_t *desc, const double *brlen, const int32_t *tip_rows,
                                         int64_t shard_index, int64_t shard_count, double *out_lnl, int out_where,
                                         void *stream) {
  PLG_API_BEGIN_LOCKED
  if (shard_count < 1 || shard_index < 0 || shard_index >= shard_count)
    PLG_FAIL(PLG_ERR_ARG, "bad shard %lld of %lld", (long long)shard_index, (long long)shard_count);
  if (out_where != PLG_OUT_HOST && out_where != PLG_OUT_DEVICE) PLG_FAIL(PLG_ERR_ARG, "out_where: PLG_OUT_HOST or PLG_OUT_DEVICE");
  if (n_tips < 2) PLG_FAIL(PLG_ERR_ARG, "descriptor trees need at least 2 tips");
  const int64_t lo = n_trees * shard_index / shard_count, hi = n_trees * (shard_index + 1) / shard_count;
  const size_t stride = (size_t)(n_tips - 1) * 2;
  OutDeviceScope scope(out_where == PLG_OUT_DEVICE);
  if (hi > lo)
    lik_score_trees_impl(aln, m, hi - lo, n_tips, desc ? desc + lo * stride : nullptr, brlen ? brlen + lo * stride : nullptr,
                         tip_rows, out_lnl ? out_lnl + lo : nullptr, stream);
  PLG_API_END
}

static void lik_score_trees_impl(plg_lik_aln *aln, plg_model *m, int64_t n_trees, int n_tips, const int32_t *desc,
                                 const double *brlen, const int32_t *tip_rows, double *out_lnl, void *stream) {
  {
  if (!aln || !m || !out_lnl || (n_trees > 0 && (!desc || !brlen))) PLG_FAIL(PLG_ERR_ARG, "null argument");
  if (aln->impl.K == 4 && m->impl.K == 4 && m->impl.R == 4 && lik_tree_walk_enabled() && n_tips >= 2 &&
      n_tips <= (1 << WALK_MAX_LEVELS)) {
    score_lik_desc_walk(aln->impl, m->impl, n_trees, n_tips, desc, brlen, tip_rows, out_lnl, (cudaStream_t)stream);
    return;
  }
  TreeBatch b = batch_from_desc(n_trees, n_tips, desc, brlen, tip_rows);
  for (int32_t c : b.child)
    if (c >= aln->impl.n_rows) PLG_FAIL(PLG_ERR_ARG, "tip row %d is outside the alignment (%d rows)", c, aln->impl.n_rows);
  score_lik_batch_full(aln->impl, m->impl, b, out_lnl, (cudaStream_t)stream);
  }
}
