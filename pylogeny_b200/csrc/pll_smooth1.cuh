// Branch-length smoothing of ONE tree in ONE cooperative launch (DNA x Gamma-4).
//
// What the reference asks per tree (libpllWrapper.c:205: pllOptimizeBranchLengths(.., 3) on default
// lengths, then the lnL; scoring.py:41-75 makes one instance per tree, heuristic.py:107-260 calls it
// per neighbour through landscape.vertex.scoreLikelihood) is a chain of ~5 n dependent steps per pass:
// re-orient a conditional vector, or one Newton step on a branch.  As separate launches (the lockstep
// program of pll_loglik_batch with a batch of one) that is ~930 launches per 64-taxon tree and ~10 ms,
// all of it launch gaps.  But the steps depend on each other only THROUGH THE BRANCH LENGTHS: a vector
// update is pattern-local, and a Newton step needs three sums over all patterns.  So one persistent
// grid -- a block = a tile of 128 patterns -- runs the whole program, and the ONLY data that crosses
// blocks are those sums.  The launch is one long dependent chain (~1,300 actions), so what it costs is
// the latency of an action, and the design removes latencies one by one:
//
//  * every block takes the same Newton step from the same sums, so each keeps its own branch lengths
//    and its own copy of ALL the tree's P-matrices in shared memory (64 doubles per branch; a tip
//    operand is the matrix applied to the indicator vector of its code, the same additions as a tip
//    lookup table): nothing is published, no matrix is staged, and a vector update needs no block
//    barrier at all (trees of more than 174 taxa: the matrices that do not fit live in the block's own
//    global region, instantiation FIT = false);
//  * the sums travel as flagged words (the LL protocol of collective libraries): a block stores each
//    half of its three partial sums as one 8-byte word (generation << 32 | half) into every block's
//    inbox, every block polls its own inbox until all words carry the current generation and adds
//    them in a fixed order.  One store -> load round trip per Newton evaluation, no fence, no atomic,
//    no separate barrier (the first version -- atomic counter, spin, then read the partials -- was
//    three round trips; the parity double-buffer is safe because a block can only be one evaluation
//    ahead of the slowest);
//  * the operand an action reads from memory is fetched (cp.async into the thread's own shared-memory
//    slots) while the previous action computes; the other operand is as a rule the column the previous
//    action produced, which stays in registers.
//
// A block covers tiles blockIdx.x, blockIdx.x + gridDim.x, ... , so any alignment length runs on a
// co-resident grid (cooperative launch: co-residency is what makes polling safe).  With several tiles
// per block (instantiation ONE = false) the operands are loaded where they are used and a
// curvature-guard retry recomputes the sum-table columns.
//
// Loads of vectors written inside the kernel are ld.global.cg / cp.async of the thread's own data.
#pragma once

namespace plg {

#ifndef PLG_S1_SPLIT
#define PLG_S1_SPLIT 1
#endif
constexpr int S1_SPLIT = PLG_S1_SPLIT;        // threads per pattern: each keeps 4 / S1_SPLIT rate categories of a column
                                              // (2: 2.51 ms against 2.37 per 64 x 10k tree -- twice the warps at every barrier)
constexpr int S1_TILE = 128;                  // patterns per tile
constexpr int S1_THREADS = S1_TILE * S1_SPLIT;
constexpr int S1_E = 16 / S1_SPLIT;           // column entries per thread
constexpr int S1_R = 4 / S1_SPLIT;            // rate categories per thread
static_assert(S1_SPLIT == 1 || S1_SPLIT == 2, "a pattern's threads are neighbouring lanes");

struct S1Act {
  int kind;              // 0: vector update (dst = a x b through pa, pb); 1: Newton step on `branch` between a and b
  int dst, a, b, pa, pb;
  int branch, pad;
};

// what is fetched ahead of an action (one tile per block): both tip codes (registers), and ONE vector
// operand's column and scaling counter (`which`: 1 = operand a, 2 = operand b, 0 = none), copied
// asynchronously into the thread's own shared-memory slots -- held in registers across the running action
// the column is spilled by the compiler the moment it arrives, which puts the load back on the critical
// path.  When both operands are vectors in memory the second is loaded where it is used.
struct S1Pre {
  int ca, cb, which;
};

// The launch's own vector layout: [node][tile][entry][pattern] -- the entries of a pattern sit at constant
// offsets from one address (the [node][entry][pattern] layout of the other kernels costs a 64-bit multiply-add
// per entry, a sixth of this kernel's instructions).  `half`: which S1_E entries this thread keeps.
template <typename T>
__device__ __forceinline__ T *s1_col(T *clv, int node, int64_t p, int64_t Ppad, int half) {
  return clv + ((int64_t)node * Ppad + (p & ~(int64_t)(S1_TILE - 1))) * 16 + half * (S1_E * S1_TILE) + (p & (S1_TILE - 1));
}

__device__ __forceinline__ void s1_load(double (&x)[S1_E], int &sc, int id, int n_rows, int64_t p, int64_t Ppad, int half,
                                        const double *clv, const int *scl) {
  const double *base = s1_col(clv, id - n_rows, p, Ppad, half);
#pragma unroll
  for (int i = 0; i < S1_E; i++) x[i] = __ldcg(base + i * S1_TILE);
  sc = __ldcg(scl + (int64_t)(id - n_rows) * Ppad + p);
}

__device__ __forceinline__ void s1_cp_async8(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void s1_cp_async4(void *smem, const void *gmem) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 4;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory");
}
__device__ __forceinline__ void s1_cp_wait() { asm volatile("cp.async.wait_all;\n" ::: "memory"); }

// fetch for the action with operands (a, b); `skip`: the vector that will be in registers (or -1);
// px / psc: this thread's slots ([entry][pattern] doubles, one int)
__device__ __forceinline__ void s1_fetch(S1Pre &o, double *px, int *psc, int a, int b, int skip, int n_rows, int64_t p,
                                         int64_t Ppad, int half, const unsigned char *__restrict__ codes, const double *clv,
                                         const int *scl) {
  o.which = 0;
  int vec = -1;
  if (a != skip) {
    if (a < n_rows) o.ca = codes[(int64_t)a * Ppad + p];
    else { vec = a; o.which = 1; }
  }
  if (b != skip) {
    if (b < n_rows) o.cb = codes[(int64_t)b * Ppad + p];
    else if (vec < 0) { vec = b; o.which = 2; }
  }
  if (vec >= 0) {
    const double *base = s1_col(clv, vec - n_rows, p, Ppad, half);
#pragma unroll
    for (int i = 0; i < S1_E; i++) s1_cp_async8(px + i * S1_TILE, base + i * S1_TILE);
    s1_cp_async4(psc, scl + (int64_t)(vec - n_rows) * Ppad + p);
  }
}

// v (*)= P x for this thread's categories of a column in registers (sm: their matrices, [r][k][j])
template <bool MUL>
__device__ __forceinline__ void s1_apply_regs(const double *sm, const double (&x)[S1_E], double (&v)[S1_E]) {
#pragma unroll
  for (int r = 0; r < S1_R; r++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double2 p01 = *reinterpret_cast<const double2 *>(sm + r * 16 + k * 4);
      const double2 p23 = *reinterpret_cast<const double2 *>(sm + r * 16 + k * 4 + 2);
      double s = p01.x * x[r * 4];
      s = fma(p01.y, x[r * 4 + 1], s);
      s = fma(p23.x, x[r * 4 + 2], s);
      s = fma(p23.y, x[r * 4 + 3], s);
      if (MUL) v[r * 4 + k] *= s; else v[r * 4 + k] = s;
    }
}

// all-reduce over the S1_SPLIT neighbouring lanes of a pattern
__device__ __forceinline__ double s1_pair_sum(double x) {
  if (S1_SPLIT == 2) x += __shfl_xor_sync(0xffffffffu, x, 1);
  return x;
}
__device__ __forceinline__ int s1_pair_max(int x) {
  if (S1_SPLIT == 2) x = max(x, __shfl_xor_sync(0xffffffffu, x, 1));
  return x;
}

// The blocks' partial sums, exchanged as flagged words: every block PUSHES its three sums into every block's
// inbox (ll[parity][destination][source][3 sums][2 halves]; each 8-byte word = generation << 32 | half of a
// double) and polls only its own inbox -- lines no other SM reads.  (Polling the senders' slots instead had
// all blocks x 32 lanes spinning on the same ~30 lines: 4.5 us per exchange at 80 blocks.)
// red[q][warp]: this block's per-warp sums; on return tot[Q0..2] hold the grid's sums, added in the same order
// by every block (Q0 = 1: a Newton evaluation does not need the lnL, and every word is instructions on the
// launch's critical path).
template <int Q0>
__device__ __forceinline__ void s1_exchange(unsigned long long *ll, unsigned int gen, int parity,
                                            const double (*red)[S1_THREADS / 32], double *tot) {
  const unsigned int G = gridDim.x;
  unsigned long long *box = ll + (size_t)parity * G * G * 6;
  if (threadIdx.x < G) {
    unsigned long long w[6];
    const unsigned long long tag = (unsigned long long)gen << 32;
#pragma unroll
    for (int q = Q0; q < 3; q++) {
      double m = 0.0;
#pragma unroll
      for (int i = 0; i < S1_THREADS / 32; i++) m += red[q][i];
      const unsigned long long bits = (unsigned long long)__double_as_longlong(m);
      w[2 * q] = tag | (bits & 0xffffffffull);
      w[2 * q + 1] = tag | (bits >> 32);
    }
    for (unsigned int dest = threadIdx.x; dest < G; dest += S1_THREADS) {
      unsigned long long *slot = box + ((size_t)dest * G + blockIdx.x) * 6;
#pragma unroll
      for (int q = Q0; q < 3; q++)
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};\n" ::"l"(slot + 2 * q), "l"(w[2 * q]), "l"(w[2 * q + 1]) : "memory");
    }
  }
  if (threadIdx.x < 32) {  // ONE warp adds the blocks' shares: strided, then a fixed shuffle tree.  (Measured slower, per
                           // 64 x 10k tree against 2.19 ms: every thread polling one record, 3.21; a __nanosleep of
                           // 20 ns between polling rounds, 3.24.)
    double a[3] = {0.0, 0.0, 0.0};
    const unsigned long long *mine = box + (size_t)blockIdx.x * G * 6;
    for (unsigned int b0 = 0; b0 < G; b0 += 128) {  // four records per lane in flight
      unsigned long long w[4][6];
      bool done[4], all;
#pragma unroll
      for (int j = 0; j < 4; j++) done[j] = b0 + threadIdx.x + 32 * j >= G;
      do {
        all = true;
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (!done[j]) {
            const unsigned long long *s = mine + (size_t)(b0 + threadIdx.x + 32 * j) * 6;
#pragma unroll
            for (int q = Q0; q < 3; q++)
              asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(w[j][2 * q]), "=l"(w[j][2 * q + 1]) : "l"(s + 2 * q) : "memory");
          }
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (!done[j]) {
            bool ok = true;
#pragma unroll
            for (int q = 2 * Q0; q < 6; q++) ok = ok && (unsigned int)(w[j][q] >> 32) == gen;
            done[j] = ok;
            all = all && ok;
          }
      } while (!all);
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (b0 + threadIdx.x + 32 * j < G) {
#pragma unroll
          for (int q = Q0; q < 3; q++)
            a[q] += __longlong_as_double((long long)((w[j][2 * q] & 0xffffffffull) | (w[j][2 * q + 1] << 32)));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
#pragma unroll
      for (int q = Q0; q < 3; q++) a[q] += __shfl_down_sync(0xffffffffu, a[q], d);
    }
    if (threadIdx.x == 0) { tot[0] = a[0]; tot[1] = a[1]; tot[2] = a[2]; }
  }
}

// acts[0 .. n_down): the down pass; acts[n_down .. n_down + n_pass): one smoothing pass (run up to
// max_passes times, until a pass moves no branch by more than PLL_DELTAZ in z); acts[n_down + n_pass]:
// the evaluation step (kind 1: its sums at the branch's final length give the lnL).
// pmat: the matrices of the initial lengths (rows of 17, as pmatrix_kernel writes them); ll: 2 x gridDim.x^2 x 6
// words, zero at launch; brlen: initial lengths in, final lengths out; out[0] = lnL.
// Dynamic shared memory: n_br (FIT) or n_sm_br x 64 doubles (matrices), n_br doubles (lengths), then the action list
// when acts_smem != 0; spill_all: gridDim.x x (n_br - n_sm_br) x 64 doubles when not FIT.
template <bool ONE, bool FIT>
__global__ void __launch_bounds__(S1_THREADS)
smooth1_kernel(const S1Act *__restrict__ acts, int n_down, int n_pass, int max_passes, int n_rows, int n_br, int n_sm_br,
               double *spill_all, int64_t Ppad,
               const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
               const double *__restrict__ weights, const ModelDev *__restrict__ model, const double *__restrict__ pmat,
               double *clv, int *scl, double *brlen, unsigned long long *ll, double *out, int acts_smem) {
  __shared__ double mU[16], mUinv[16], mg[16], mpi[4], meval[4], mrate[4];
  __shared__ unsigned int smask[16];
  __shared__ double ex[16];
  __shared__ double tab[3][16];
  __shared__ double red[3][S1_THREADS / 32];
  __shared__ double tot[3];
  __shared__ double pre_x[ONE ? 2 : 1][16][ONE ? S1_TILE : 1];  // fetched columns, [buffer][entry][pattern]
  __shared__ int pre_sc[2][S1_THREADS];
  const int n_tiles = (int)(Ppad / S1_TILE);
  const int pat = threadIdx.x / S1_SPLIT, half = threadIdx.x % S1_SPLIT;  // a pattern's threads are neighbouring lanes
  extern __shared__ __align__(16) unsigned char s1_dyn[];
  // FIT: every branch's matrix is in shared memory.  Otherwise (trees of more than ~174 taxa) the first n_sm_br are,
  // and the rest live in this block's own global region (L2): slower steps on those branches, still one launch.
  double *pm = reinterpret_cast<double *>(s1_dyn);  // [branch][r][k][j]
  double *bl = pm + (size_t)(FIT ? n_br : n_sm_br) * 64;
  double *spill = FIT ? nullptr : spill_all + (size_t)blockIdx.x * (n_br - n_sm_br) * 64;
  auto mat = [&](int b) -> double * {
    if (FIT) return pm + b * 64;
    return b < n_sm_br ? pm + b * 64 : spill + (size_t)(b - n_sm_br) * 64;
  };
  const S1Act *alist = acts;
  if (acts_smem) {
    const int n_all = n_down + n_pass + 1;
    int4 *d = reinterpret_cast<int4 *>(bl + ((n_br + 1) & ~1));
    const int4 *g = reinterpret_cast<const int4 *>(acts);
    for (int e = threadIdx.x; e < n_all * 2; e += S1_THREADS) d[e] = g[e];
    alist = reinterpret_cast<const S1Act *>(d);
  }
  for (int e = threadIdx.x; e < n_br; e += S1_THREADS) bl[e] = brlen[e];
  for (int e = threadIdx.x; e < n_br * 64; e += S1_THREADS) mat(e >> 6)[e & 63] = pmat[(e >> 6) * 68 + ((e >> 4) & 3) * 17 + (e & 15)];
  if (threadIdx.x < 16) {
    mU[threadIdx.x] = model->U[threadIdx.x];          // [i][k]  (LIK_MAX_STATES-strided rows are not used: K = 4 packs them)
    mUinv[threadIdx.x] = model->Uinv[threadIdx.x];    // [k][j]
    mg[threadIdx.x] = -model->eval[threadIdx.x & 3] * model->rates[threadIdx.x >> 2];
    smask[threadIdx.x] = masks[threadIdx.x];
    if (threadIdx.x < 4) {
      mpi[threadIdx.x] = model->freqs[threadIdx.x];
      meval[threadIdx.x] = model->eval[threadIdx.x];
      mrate[threadIdx.x] = model->rates[threadIdx.x];
    }
  }
  __syncthreads();
  int parity = 0;
  unsigned int generation = 0;
  // ONE tile per block (the host picks the instantiation): a thread owns one pattern for the whole launch.
  // Most actions read the vector the previous action wrote (the re-orientation chain along the depth-first
  // path, the Newton step on the branch just re-oriented toward): that column stays in registers (last_v),
  // and the other operand of the next action is fetched while this one computes.
  constexpr bool one = ONE;
  const int64_t p_own = (int64_t)blockIdx.x * S1_TILE + pat;
  const double w_own = half == 0 ? weights[p_own] : 0.0;  // (a pattern's site term is added once)
  int last_dst = -1, last_sc = 0;
  double last_v[S1_E];
#pragma unroll
  for (int i = 0; i < S1_E; i++) last_v[i] = 0.0;

  // operand as it sits at its node (tips: indicator of the code's states), three ways: the column the previous
  // action left in registers, the fetched one (ONE: o holds this action's fetch; idx = 1 / 2 for operand a / b),
  // or loaded here
  auto raw = [&](int id, int idx, const S1Pre &o, const double *px, const int *psc, int64_t p, double (&x)[S1_E], int &sc) {
    if (one && id == last_dst) {
#pragma unroll
      for (int i = 0; i < S1_E; i++) x[i] = last_v[i];
      sc += last_sc;
    } else if (id < n_rows) {
      const int code = one ? (idx == 1 ? o.ca : o.cb) : (int)codes[(int64_t)id * Ppad + p];
      const unsigned int m = smask[code];
#pragma unroll
      for (int i = 0; i < S1_E; i++) x[i] = (m >> (i & 3) & 1) ? 1.0 : 0.0;
    } else if (one && o.which == idx) {
#pragma unroll
      for (int i = 0; i < S1_E; i++) x[i] = px[i * S1_TILE];
      sc += *psc;
    } else {
      int xs;
      s1_load(x, xs, id, n_rows, p, Ppad, half, clv, scl);
      sc += xs;
    }
  };

  // vector update: dst = (P_pa a) o (P_pb b).  Everything it touches is the thread's own or read-only shared
  // memory: no block barrier.
  auto update = [&](const S1Act &op, const S1Pre &o, const double *px, const int *psc) {
    const double *ma = mat(op.pa) + half * (S1_R * 16), *mb = mat(op.pb) + half * (S1_R * 16);
    for (int tile = blockIdx.x; tile < (one ? (int)blockIdx.x + 1 : n_tiles); tile += gridDim.x) {
      const int64_t p = (int64_t)tile * S1_TILE + pat;
      double x[S1_E], v[S1_E];
      int sc = 0;
      raw(op.a, 1, o, px, psc, p, x, sc);
      s1_apply_regs<false>(ma, x, v);
      raw(op.b, 2, o, px, psc, p, x, sc);
      s1_apply_regs<true>(mb, x, v);
      int hi = 0;
#pragma unroll
      for (int i = 0; i < S1_E; i++) hi = max(hi, __double2hiint(v[i]));
      hi = s1_pair_max(hi);
      if (hi < 0x2FF00000) {  // every entry below 2^-256 (values are >= 0): rescale the column, count it
#pragma unroll
        for (int i = 0; i < S1_E; i++) v[i] *= 1.15792089237316195423570985008687907853269984665640564039457584007913129639936e77;
        sc += 1;
      }
      double *dst = s1_col(clv, op.dst - n_rows, p, Ppad, half);
#pragma unroll
      for (int i = 0; i < S1_E; i++) dst[i * S1_TILE] = v[i];
      scl[(int64_t)(op.dst - n_rows) * Ppad + p] = sc;  // (every thread of the pattern: each later reads what it wrote itself)
      if (one) {
#pragma unroll
        for (int i = 0; i < S1_E; i++) last_v[i] = v[i];
        last_sc = sc;
      }
    }
    last_dst = op.dst;
  };

  // this thread's entries of the sum-table column of pattern p for the branch between operands a and b
  auto column = [&](const S1Act &op, const S1Pre &o, const double *px, const int *psc, int64_t p, double (&sv)[S1_E], int &sc) {
    double A[S1_E], B[S1_E];
    sc = 0;
    raw(op.a, 1, o, px, psc, p, A, sc);
    raw(op.b, 2, o, px, psc, p, B, sc);
#pragma unroll
    for (int r = 0; r < S1_R; r++)
#pragma unroll
      for (int k = 0; k < 4; k++) {
        double l = 0.0, rr = 0.0;
#pragma unroll
        for (int i = 0; i < 4; i++) {
          l = fma(A[r * 4 + i] * mpi[i], mU[i * 4 + k], l);
          rr = fma(mUinv[k * 4 + i], B[r * 4 + i], rr);
        }
        sv[r * 4 + k] = l * rr;
      }
  };

  // sums (lnL when `lnl`, d1, d2) over all patterns at log z = lz, identical in every thread of the grid.
  // One tile per block: the caller's column (sv1, svsc1); else the columns are rebuilt per tile.
  auto sums = [&](auto want_lnl, const S1Act &op, const S1Pre &o, const double (&sv1)[S1_E], int svsc1, double lz,
                  double &s0, double &s1, double &s2) {
    constexpr bool lnl = decltype(want_lnl)::value;
    if (threadIdx.x < 16) {
      const double g = mg[threadIdx.x], e = exp(g * lz);
      tab[0][threadIdx.x] = e; tab[1][threadIdx.x] = g * e; tab[2][threadIdx.x] = g * g * e;
    }
    __syncthreads();
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    auto term = [&](const double (&sv)[S1_E], int sc, double w) {
      double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
      for (int e = 0; e < S1_E; e++) {
        l0 = fma(sv[e], tab[0][half * S1_E + e], l0);
        l1 = fma(sv[e], tab[1][half * S1_E + e], l1);
        l2 = fma(sv[e], tab[2][half * S1_E + e], l2);
      }
      l0 = s1_pair_sum(l0); l1 = s1_pair_sum(l1); l2 = s1_pair_sum(l2);
      if (w != 0.0) {
        const double inv = 1.0 / l0, q1 = l1 * inv;
        if (lnl) v0 += w * (log(l0 * 0.25) + sc * SM_LOG_MINLH);  // (a Newton step does not look at the lnL)
        v1 += w * q1;
        v2 += w * (l2 * inv - q1 * q1);
      }
    };
    if (one) {
      term(sv1, svsc1, w_own);
    } else {
      for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
        const int64_t p = (int64_t)tile * S1_TILE + pat;
        double sv[S1_E];
        int sc;
        column(op, o, nullptr, nullptr, p, sv, sc);
        term(sv, sc, half == 0 ? weights[p] : 0.0);
      }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      v0 += __shfl_down_sync(0xffffffffu, v0, d);
      v1 += __shfl_down_sync(0xffffffffu, v1, d);
      v2 += __shfl_down_sync(0xffffffffu, v2, d);
    }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = v0; red[1][threadIdx.x >> 5] = v1; red[2][threadIdx.x >> 5] = v2; }
    __syncthreads();
    generation++;
    s1_exchange<lnl ? 0 : 1>(ll, generation, parity, red, tot);
    __syncthreads();
    s0 = tot[0]; s1 = tot[1]; s2 = tot[2];
    parity ^= 1;
    // (tab / red / tot are next written behind the block barriers of the step that follows)
  };

  // one makenewz step (libpll: optimise the branch with maxiter = 1), as newton_step_block; true: the branch moved
  auto newton = [&](const S1Act &op, const S1Pre &o, const double *px, const int *psc) -> bool {
    double sv[S1_E];
    int svsc = 0;
    if (one) column(op, o, px, psc, p_own, sv, svsc);
    const double z0 = exp(-bl[op.branch]);
    double z = fmin(fmax(z0, PLL_ZMIN), PLL_ZMAX), zprev = z;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int guard = 0; guard < 64; guard++) {  // "bad curvature: shorten the branch"
      if (guard) __syncthreads();  // tab / tot of the previous evaluation
      sums(std::false_type{}, op, o, sv, svsc, log(fmax(z, PLL_ZMIN)), s0, s1, s2);
      if (s2 >= 0.0 && z < PLL_ZMAX) zprev = z = 0.37 * z + 0.63;
      else break;
    }
    if (s2 < 0.0) {
      const double tantmp = -s1 / s2;
      if (tantmp < 100.0) {
        z *= exp(tantmp);
        if (z < PLL_ZMIN) z = PLL_ZMIN;
        if (z > 0.25 * zprev + 0.75) z = 0.25 * zprev + 0.75;
      } else {
        z = 0.25 * zprev + 0.75;
      }
    }
    if (z > PLL_ZMAX) z = PLL_ZMAX;
    const double t_new = -log(z);
    // the moved branch's matrix, by this block for itself (the arithmetic of pmatrix_block, lik_pmatrix.cuh)
    if (threadIdx.x < 16) ex[threadIdx.x] = exp(meval[threadIdx.x & 3] * t_new * mrate[threadIdx.x >> 2]);
    __syncthreads();
    if (threadIdx.x < 64) {
      const int r = threadIdx.x >> 4, i = (threadIdx.x >> 2) & 3, j = threadIdx.x & 3;
      double s = 0.0;
#pragma unroll
      for (int k = 0; k < 4; k++) s += mU[i * 4 + k] * ex[r * 4 + k] * mUinv[k * 4 + j];
      s = fmax(s, 0.0);  // transition probabilities: round-off can leave -1e-17 at t ~ 0
      if (t_new == 0.0) s = (i == j) ? 1.0 : 0.0;  // a zero-length branch is the identity exactly
      mat(op.branch)[threadIdx.x] = s;
    }
    if (threadIdx.x == 0) bl[op.branch] = t_new;
    __syncthreads();
    return fabs(z - z0) > PLL_DELTAZ;
  };

  // The pipeline: `A` runs on the fetch `o` in slot `buf`; the fetch for the next action N is issued before A's work.
  int buf = 0;
  bool fetched = false, changed = false;
  S1Pre o;
  o.ca = o.cb = o.which = 0;
  auto slot_x = [&](int b) { return &pre_x[one ? b : 0][one ? half * S1_E : 0][one ? pat : 0]; };
  auto run = [&](const S1Act &A, bool has_next, const S1Act &N) {
    if (one) {
      if (!fetched) s1_fetch(o, slot_x(buf), &pre_sc[buf][threadIdx.x], A.a, A.b, last_dst, n_rows, p_own, Ppad, half, codes, clv, scl);
      s1_cp_wait();  // this action's fetched column (the thread's own slots: no block barrier)
    }
    S1Pre q;
    q.ca = q.cb = q.which = 0;
    if (one && has_next)
      s1_fetch(q, slot_x(buf ^ 1), &pre_sc[buf ^ 1][threadIdx.x], N.a, N.b, A.kind == 0 ? A.dst : last_dst, n_rows, p_own, Ppad,
               half, codes, clv, scl);
    if (A.kind == 0) update(A, o, slot_x(buf), &pre_sc[buf][threadIdx.x]);
    else if (newton(A, o, slot_x(buf), &pre_sc[buf][threadIdx.x])) changed = true;
    if (one && has_next) o = q;
    buf ^= 1;
    fetched = has_next;
  };

  // (the records are read two actions ahead: a record is the first thing an action needs)
  const int n_acts = n_down + n_pass;
  {
    S1Act A = alist[0], N = alist[1];
    for (int i = 0; i < n_down; i++) {
      const S1Act NN = alist[min(i + 2, n_acts)];
      run(A, true, N);
      A = N; N = NN;
    }
  }
  for (int pass = 0; pass < max_passes; pass++) {
    changed = false;
    // (what follows a pass's last action depends on `changed`: that one starts without a fetch)
    S1Act A = alist[n_down], N = alist[n_down + 1];
    for (int i = 0; i < n_pass; i++) {
      const S1Act NN = alist[min(n_down + i + 2, n_acts)];
      run(A, i + 1 < n_pass, N);
      A = N; N = NN;
    }
    if (!changed) break;
  }
  {  // lnL at the final lengths
    const S1Act op = alist[n_down + n_pass];
    double sv[S1_E];
    int svsc = 0;
    if (one) {
      s1_fetch(o, slot_x(0), &pre_sc[0][threadIdx.x], op.a, op.b, last_dst, n_rows, p_own, Ppad, half, codes, clv, scl);
      s1_cp_wait();
      column(op, o, slot_x(0), &pre_sc[0][threadIdx.x], p_own, sv, svsc);
    }
    const double z = fmin(fmax(exp(-bl[op.branch]), PLL_ZMIN), PLL_ZMAX);
    double s0, s1, s2;
    __syncthreads();  // tab / tot of the last Newton evaluation
    sums(std::true_type{}, op, o, sv, svsc, log(fmax(z, PLL_ZMIN)), s0, s1, s2);
    if (blockIdx.x == 0) {
      for (int e = threadIdx.x; e < n_br; e += S1_THREADS) brlen[e] = bl[e];
      if (threadIdx.x == 0) out[0] = s0;
    }
  }
}

}  // namespace plg
