// Branch-length smoothing of ONE tree in ONE cooperative launch (DNA x Gamma-4).
//
// What the reference asks per tree (libpllWrapper.c:205: pllOptimizeBranchLengths(.., 3) on default
// lengths, then the lnL; scoring.py:41-75 makes one instance per tree, heuristic.py:107-260 calls it
// per neighbour through landscape.vertex.scoreLikelihood) is a chain of ~5 n dependent steps per pass:
// re-orient a conditional vector, or one Newton step on a branch.  As separate launches (the lockstep
// program of pll_loglik_batch with a batch of one) that is ~930 launches per 64-taxon tree and ~10 ms,
// all of it launch gaps.  But the steps depend on each other only THROUGH THE BRANCH LENGTHS: a vector
// update is pattern-local, and a Newton step needs three sums over all patterns.  So one persistent
// grid -- thread = pattern, block = 128 patterns -- runs the whole program, and the ONLY data that
// crosses blocks are those sums:
//
//  * every block takes the same Newton step from the same sums, so each keeps its own branch lengths
//    (shared memory) and its own copy of the matrices it has moved (a private global region; `moved`
//    says which branches read it instead of the launch's initial tables) -- nothing is published;
//  * the sums travel as flagged words (the LL protocol of collective libraries): a block stores each
//    half of its three partial sums as one 8-byte word (generation << 32 | half) into every block's
//    inbox, every block polls its own inbox until all words carry the current generation and adds
//    them in a fixed order.  One store -> load round trip per Newton evaluation, no fence, no atomic,
//    no separate barrier (the first version -- atomic counter, spin, then read the partials -- was
//    three round trips; the parity double-buffer is safe because a block can only be one evaluation
//    ahead of the slowest);
//  * the chain of ~1,300 actions is software-pipelined: while an action computes, the next one's
//    record, matrices and (one tile per block) operands are already in flight, so an action's
//    critical path is its arithmetic and two block barriers, not two L2 round trips.
//
// A block covers tiles blockIdx.x, blockIdx.x + gridDim.x, ... of 128 patterns, so any alignment length
// runs on a co-resident grid (cooperative launch: co-residency is what makes polling safe).  With
// several tiles per block the operands are loaded where they are used and a curvature-guard retry
// recomputes the sum-table columns.
//
// Loads of data written inside the kernel (vectors, counters, private matrices) are ld.global.cg: the
// read-only (.nc) path the other kernels use is not coherent with stores of the same launch.
#pragma once

namespace plg {

#ifndef PLG_S1_THREADS
#define PLG_S1_THREADS 128
#endif
constexpr int S1_THREADS = PLG_S1_THREADS;
static_assert(S1_THREADS >= 64 && 256 % S1_THREADS == 0, "staging assumes 64 <= threads <= 256");
constexpr int S1_TT_PER = 256 / S1_THREADS;   // tip-table entries per thread
constexpr int S1_PRIV = 68 + 256;             // doubles per branch in a block's private region: P, tip table

struct S1Act {
  int kind;              // 0: vector update (dst = a x b through pa, pb); 1: Newton step on `branch` between a and b
  int dst, a, b, pa, pb;
  int branch, pad;
};

// what is fetched ahead of an action (one tile per block): both tip codes (registers), and ONE vector
// operand's column and scaling counter (`which`: 1 = operand a, 2 = operand b, 0 = none), copied
// asynchronously into the thread's own shared-memory slots -- held in registers across the running action
// the column is spilled by the compiler the moment it arrives, which puts the load back on the critical
// path.  The other operand of an action is as a rule the vector the previous action produced, which is in
// registers anyway; when both are vectors in memory the second is loaded where it is used.
struct S1Pre {
  int ca, cb, which;
};

// a matrix / tip table fetched ahead: this thread's elements, filed into the staging buffer later
struct S1Mat {
  double m[S1_TT_PER];
  int how;  // 0: none; 1: tip table; 2: matrix; 3 / 4: tip table / matrix of the branch the running action moves
};

__device__ __forceinline__ void s1_mat_fetch(S1Mat &o, bool tip, int pm, int moving, const unsigned char *moved,
                                             const double *__restrict__ pmat, const double *__restrict__ tiptab,
                                             const double *priv) {
  if (pm == moving) { o.how = tip ? 3 : 4; return; }
  const bool mv = moved[pm] != 0;
  if (tip) {
    const double *src = mv ? priv + (int64_t)pm * S1_PRIV + 68 : tiptab + (int64_t)pm * 256;
#pragma unroll
    for (int q = 0; q < S1_TT_PER; q++) o.m[q] = __ldcg(src + threadIdx.x + q * S1_THREADS);
    o.how = 1;
  } else {
    const double *src = mv ? priv + (int64_t)pm * S1_PRIV : pmat + (int64_t)pm * 68;
    if (threadIdx.x < 64) o.m[0] = __ldcg(src + (threadIdx.x >> 4) * 17 + (threadIdx.x & 15));
    o.how = 2;
  }
}

__device__ __forceinline__ void s1_mat_file(const S1Mat &o, double *sm, const double *cur_pm, const double *cur_tt) {
  if (o.how == 1) {
#pragma unroll
    for (int q = 0; q < S1_TT_PER; q++) sm[threadIdx.x + q * S1_THREADS] = o.m[q];
  } else if (o.how == 2) {
    if (threadIdx.x < 64) sm[threadIdx.x] = o.m[0];
  } else if (o.how == 3) {
#pragma unroll
    for (int q = 0; q < S1_TT_PER; q++) sm[threadIdx.x + q * S1_THREADS] = cur_tt[threadIdx.x + q * S1_THREADS];
  } else if (o.how == 4) {
    if (threadIdx.x < 64) sm[threadIdx.x] = cur_pm[(threadIdx.x >> 4) * 17 + (threadIdx.x & 15)];
  }
}

// The launch's own vector layout: [node][tile][entry][thread] -- the 16 entries of a pattern sit at constant
// offsets from one address (the [node][entry][pattern] layout of the other kernels costs a 64-bit multiply-add
// per entry, a sixth of this kernel's instructions).
template <typename T>
__device__ __forceinline__ T *s1_col(T *clv, int node, int64_t p, int64_t Ppad) {
  return clv + ((int64_t)node * Ppad + (p & ~(int64_t)(S1_THREADS - 1))) * 16 + (p & (S1_THREADS - 1));
}

__device__ __forceinline__ void s1_load(double (&x)[16], int &sc, int id, int n_rows, int64_t p, int64_t Ppad, const double *clv,
                                        const int *scl) {
  const double *base = s1_col(clv, id - n_rows, p, Ppad);
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = __ldcg(base + i * S1_THREADS);
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
// px / psc: this thread's slots ([i][thread] doubles, one int)
__device__ __forceinline__ void s1_fetch(S1Pre &o, double *px, int *psc, int a, int b, int skip, int n_rows, int64_t p,
                                         int64_t Ppad, const unsigned char *__restrict__ codes, const double *clv,
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
    const double *base = s1_col(clv, vec - n_rows, p, Ppad);
#pragma unroll
    for (int i = 0; i < 16; i++) s1_cp_async8(px + i * S1_THREADS, base + i * S1_THREADS);
    s1_cp_async4(psc, scl + (int64_t)(vec - n_rows) * Ppad + p);
  }
}

// v (*)= P x for a column in registers
template <bool MUL>
__device__ __forceinline__ void s1_apply_regs(const double *sm, const double (&x)[16], double (&v)[16]) {
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      double s = sm[r * 16 + k * 4] * x[r * 4];
      s = fma(sm[r * 16 + k * 4 + 1], x[r * 4 + 1], s);
      s = fma(sm[r * 16 + k * 4 + 2], x[r * 4 + 2], s);
      s = fma(sm[r * 16 + k * 4 + 3], x[r * 4 + 3], s);
      if (MUL) v[r * 4 + k] *= s; else v[r * 4 + k] = s;
    }
}

// The blocks' partial sums, exchanged as flagged words: every block PUSHES its three sums into every block's
// inbox (ll[parity][destination][source][3 sums][2 halves]; each 8-byte word = generation << 32 | half of a
// double) and polls only its own inbox -- lines no other SM reads.  (Polling the senders' slots instead had
// all blocks x 32 lanes spinning on the same ~30 lines: 4.5 us per exchange at 80 blocks.)
// red[q][warp]: this block's per-warp sums; on return tot[0..2] hold the grid's sums, added in the same order
// by every block.
__device__ __forceinline__ void s1_exchange(unsigned long long *ll, unsigned int gen, int parity,
                                            const double (*red)[S1_THREADS / 32], double *tot) {
  const unsigned int G = gridDim.x;
  unsigned long long *box = ll + (size_t)parity * G * G * 6;
  if (threadIdx.x < G) {
    unsigned long long w[6];
    const unsigned long long tag = (unsigned long long)gen << 32;
#pragma unroll
    for (int q = 0; q < 3; q++) {
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
      for (int q = 0; q < 3; q++)
        asm volatile("st.relaxed.gpu.global.v2.u64 [%0], {%1, %2};\n" ::"l"(slot + 2 * q), "l"(w[2 * q]), "l"(w[2 * q + 1]) : "memory");
    }
  }
  if (threadIdx.x < 32) {  // one warp adds the blocks' shares: strided, then a fixed shuffle tree
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
            for (int q = 0; q < 3; q++)
              asm volatile("ld.relaxed.gpu.global.v2.u64 {%0, %1}, [%2];\n" : "=l"(w[j][2 * q]), "=l"(w[j][2 * q + 1]) : "l"(s + 2 * q) : "memory");
          }
#pragma unroll
        for (int j = 0; j < 4; j++)
          if (!done[j]) {
            bool ok = true;
#pragma unroll
            for (int q = 0; q < 6; q++) ok = ok && (unsigned int)(w[j][q] >> 32) == gen;
            done[j] = ok;
            all = all && ok;
          }
      } while (!all);
#pragma unroll
      for (int j = 0; j < 4; j++)
        if (b0 + threadIdx.x + 32 * j < G) {
#pragma unroll
          for (int q = 0; q < 3; q++)
            a[q] += __longlong_as_double((long long)((w[j][2 * q] & 0xffffffffull) | (w[j][2 * q + 1] << 32)));
        }
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) {
      a[0] += __shfl_down_sync(0xffffffffu, a[0], d);
      a[1] += __shfl_down_sync(0xffffffffu, a[1], d);
      a[2] += __shfl_down_sync(0xffffffffu, a[2], d);
    }
    if (threadIdx.x == 0) { tot[0] = a[0]; tot[1] = a[1]; tot[2] = a[2]; }
  }
}

// acts[0 .. n_down): the down pass; acts[n_down .. n_down + n_pass): one smoothing pass (run up to
// max_passes times, until a pass moves no branch by more than PLL_DELTAZ in z); acts[n_down + n_pass]:
// the evaluation step (kind 1: its sums at the branch's final length give the lnL).
// pmat / tiptab: the matrices of the initial lengths (read only); priv: gridDim.x x n_br x S1_PRIV doubles;
// ll: 2 x gridDim.x^2 x 6 words, zero at launch; brlen: initial lengths in, final lengths out; out[0] = lnL.
// Dynamic shared memory: n_br doubles (lengths) and n_br bytes (moved), each rounded up to 16 bytes, then the
// action list when acts_smem != 0.
template <bool ONE>
__global__ void __launch_bounds__(S1_THREADS)
smooth1_kernel(const S1Act *__restrict__ acts, int n_down, int n_pass, int max_passes, int n_rows, int n_br, int64_t Ppad,
               const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
               const double *__restrict__ weights, const ModelDev *__restrict__ model, const double *__restrict__ pmat,
               const double *__restrict__ tiptab, double *priv_all, double *clv, int *scl, double *brlen,
               unsigned long long *ll, double *out, int acts_smem) {
  __shared__ __align__(16) double smA[2][256];
  __shared__ __align__(16) double smB[2][256];
  __shared__ double cur_pm[68], cur_tt[256];
  __shared__ double mU[16], mUinv[16], mg[16], mpi[4];
  __shared__ unsigned int smask[16];
  __shared__ double tab[3][16];
  __shared__ double red[3][S1_THREADS / 32];
  __shared__ double tot[3];
  __shared__ double pre_x[ONE ? 2 : 1][16][ONE ? S1_THREADS : 1];  // fetched columns, [buffer][entry][thread]
  __shared__ int pre_sc[2][S1_THREADS];
  const int n_tiles = (int)(Ppad / S1_THREADS);
  extern __shared__ __align__(16) unsigned char s1_dyn[];
  double *bl = reinterpret_cast<double *>(s1_dyn);
  unsigned char *moved = s1_dyn + (((size_t)n_br * 8 + 15) & ~(size_t)15);
  const int moved_bytes = (n_br + 15) & ~15;
  const S1Act *alist = acts;
  if (acts_smem) {
    const int n_all = n_down + n_pass + 1;
    int4 *d = reinterpret_cast<int4 *>(moved + moved_bytes);
    const int4 *g = reinterpret_cast<const int4 *>(acts);
    for (int e = threadIdx.x; e < n_all * 2; e += S1_THREADS) d[e] = g[e];
    alist = reinterpret_cast<const S1Act *>(d);
  }
  for (int e = threadIdx.x; e < n_br; e += S1_THREADS) { bl[e] = brlen[e]; moved[e] = 0; }
  if (threadIdx.x < 16) {
    mU[threadIdx.x] = model->U[threadIdx.x];          // [i][k]  (LIK_MAX_STATES-strided rows are not used: K = 4 packs them)
    mUinv[threadIdx.x] = model->Uinv[threadIdx.x];    // [k][j]
    mg[threadIdx.x] = -model->eval[threadIdx.x & 3] * model->rates[threadIdx.x >> 2];
    smask[threadIdx.x] = masks[threadIdx.x];
    if (threadIdx.x < 4) mpi[threadIdx.x] = model->freqs[threadIdx.x];
  }
  __syncthreads();
  double *priv = priv_all + (size_t)blockIdx.x * n_br * S1_PRIV;
  int parity = 0;
  unsigned int generation = 0;
  // ONE tile per block (the host picks the instantiation): a thread owns one pattern for the whole launch.
  // Most actions read the vector the previous action wrote (the re-orientation chain along the depth-first
  // path, the Newton step on the branch just re-oriented toward): that column stays in registers (last_v),
  // and the other operand of the next action is fetched while this one computes.
  constexpr bool one = ONE;
  const int64_t p_own = (int64_t)blockIdx.x * S1_THREADS + threadIdx.x;
  const double w_own = weights[p_own];
  int last_dst = -1, last_sc = 0;
  double last_v[16];
#pragma unroll
  for (int i = 0; i < 16; i++) last_v[i] = 0.0;

  // One operand of an action, three ways: the column the previous action left in registers, the fetched
  // one (ONE: o holds this action's fetch; idx = 1 / 2 for operand a / b), or loaded here.
  // operand seen through a staged matrix / tip table: v (*)= P x, sc += its scaling counter
  auto through = [&](auto mul, const double *sm, int id, int idx, const S1Pre &o, const double *px, const int *psc, int64_t p,
                     double (&v)[16], int &sc) {
    constexpr bool MUL = decltype(mul)::value;
    if (one && id == last_dst) {
      s1_apply_regs<MUL>(sm, last_v, v);
      sc += last_sc;
    } else if (id < n_rows) {
      const int code = one ? (idx == 1 ? o.ca : o.cb) : (int)codes[(int64_t)id * Ppad + p];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const double t = sm[code * 16 + i];
        if (MUL) v[i] *= t; else v[i] = t;
      }
    } else if (one && o.which == idx) {
      double x[16];
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = px[i * S1_THREADS];
      s1_apply_regs<MUL>(sm, x, v);
      sc += *psc;
    } else {
      double x[16];
      int xs;
      s1_load(x, xs, id, n_rows, p, Ppad, clv, scl);
      s1_apply_regs<MUL>(sm, x, v);
      sc += xs;
    }
  };
  // operand as it sits at its node (tips: indicator of the code's states)
  auto raw = [&](int id, int idx, const S1Pre &o, const double *px, const int *psc, int64_t p, double (&x)[16], int &sc) {
    if (one && id == last_dst) {
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = last_v[i];
      sc += last_sc;
    } else if (id < n_rows) {
      const int code = one ? (idx == 1 ? o.ca : o.cb) : (int)codes[(int64_t)id * Ppad + p];
      const unsigned int m = smask[code];
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = (m >> (i & 3) & 1) ? 1.0 : 0.0;
    } else if (one && o.which == idx) {
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = px[i * S1_THREADS];
      sc += *psc;
    } else {
      int xs;
      s1_load(x, xs, id, n_rows, p, Ppad, clv, scl);
      sc += xs;
    }
  };

  // vector update; sa / sb: the staged matrices
  auto update = [&](const S1Act &op, const double *sa, const double *sb, const S1Pre &o, const double *px, const int *psc) {
    for (int tile = blockIdx.x; tile < (one ? (int)blockIdx.x + 1 : n_tiles); tile += gridDim.x) {
      const int64_t p = (int64_t)tile * S1_THREADS + threadIdx.x;
      double v[16];
      int sc = 0;
      through(std::false_type{}, sa, op.a, 1, o, px, psc, p, v, sc);
      through(std::true_type{}, sb, op.b, 2, o, px, psc, p, v, sc);
      int hi = 0;
#pragma unroll
      for (int i = 0; i < 16; i++) hi = max(hi, __double2hiint(v[i]));
      if (hi < 0x2FF00000) {  // every entry below 2^-256 (values are >= 0): rescale the column, count it
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] *= 1.15792089237316195423570985008687907853269984665640564039457584007913129639936e77;
        sc += 1;
      }
      double *dst = s1_col(clv, op.dst - n_rows, p, Ppad);
#pragma unroll
      for (int i = 0; i < 16; i++) dst[i * S1_THREADS] = v[i];
      scl[(int64_t)(op.dst - n_rows) * Ppad + p] = sc;
      if (one) {
#pragma unroll
        for (int i = 0; i < 16; i++) last_v[i] = v[i];
        last_sc = sc;
      }
    }
    last_dst = op.dst;
  };

  // sum-table column of pattern p for the branch between operands a and b
  auto column = [&](const S1Act &op, const S1Pre &o, const double *px, const int *psc, int64_t p, double (&sv)[16], int &sc) {
    double A[16], B[16];
    sc = 0;
    raw(op.a, 1, o, px, psc, p, A, sc);
    raw(op.b, 2, o, px, psc, p, B, sc);
#pragma unroll
    for (int r = 0; r < 4; r++)
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
  auto sums = [&](const S1Act &op, const S1Pre &o, const double (&sv1)[16], int svsc1, double lz, bool lnl,
                  double &s0, double &s1, double &s2) {
    if (threadIdx.x < 16) {
      const double g = mg[threadIdx.x], e = exp(g * lz);
      tab[0][threadIdx.x] = e; tab[1][threadIdx.x] = g * e; tab[2][threadIdx.x] = g * g * e;
    }
    __syncthreads();
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    auto term = [&](const double (&sv)[16], int sc, double w) {
      double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
      for (int e = 0; e < 16; e++) {
        l0 = fma(sv[e], tab[0][e], l0);
        l1 = fma(sv[e], tab[1][e], l1);
        l2 = fma(sv[e], tab[2][e], l2);
      }
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
        const int64_t p = (int64_t)tile * S1_THREADS + threadIdx.x;
        double sv[16];
        int sc;
        column(op, o, nullptr, nullptr, p, sv, sc);
        term(sv, sc, weights[p]);
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
    s1_exchange(ll, generation, parity, red, tot);
    __syncthreads();
    s0 = tot[0]; s1 = tot[1]; s2 = tot[2];
    parity ^= 1;
    __syncthreads();  // tab / red / tot are reused by the next evaluation
  };

  // one makenewz step (libpll: optimise the branch with maxiter = 1), as newton_step_block; true: the branch moved
  auto newton = [&](const S1Act &op, const S1Pre &o, const double *px, const int *psc) -> bool {
    double sv[16];
    int svsc = 0;
    if (one) column(op, o, px, psc, p_own, sv, svsc);
    const double z0 = exp(-bl[op.branch]);
    double z = fmin(fmax(z0, PLL_ZMIN), PLL_ZMAX), zprev = z;
    double s0 = 0.0, s1 = 0.0, s2 = 0.0;
    for (int guard = 0; guard < 64; guard++) {  // "bad curvature: shorten the branch"
      sums(op, o, sv, svsc, log(fmax(z, PLL_ZMIN)), false, s0, s1, s2);
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
    // the moved branch's matrices: in shared memory for the steps that follow at once, in the block's
    // private region for the later ones
    pmatrix_block<4, 4>(model, t_new, masks, cur_pm, cur_tt);
    __syncthreads();
    double *pv = priv + (int64_t)op.branch * S1_PRIV;
    if (threadIdx.x < 68) pv[threadIdx.x] = cur_pm[threadIdx.x];
#pragma unroll
    for (int q = 0; q < S1_TT_PER; q++) pv[68 + threadIdx.x + q * S1_THREADS] = cur_tt[threadIdx.x + q * S1_THREADS];
    if (threadIdx.x == 0) { bl[op.branch] = t_new; moved[op.branch] = 1; }
    return fabs(z - z0) > PLL_DELTAZ;
  };

  // The pipeline: `A` runs from staging buffer `buf` and (ONE) the fetch `o`; the fetches for the next action
  // N are issued before A's work and filed after it.
  int buf = 0;
  bool staged = false, changed = false;
  S1Pre o;
  o.ca = o.cb = o.which = 0;
  auto slot_x = [&](int b) { return &pre_x[one ? b : 0][0][one ? threadIdx.x : 0]; };
  auto run = [&](const S1Act &A, bool has_next, const S1Act &N) {
    if (!staged) {
      if (A.kind == 0) {
        S1Mat ma, mb;
        ma.how = mb.how = 0;
        s1_mat_fetch(ma, A.a < n_rows, A.pa, -1, moved, pmat, tiptab, priv);
        s1_mat_fetch(mb, A.b < n_rows, A.pb, -1, moved, pmat, tiptab, priv);
        s1_mat_file(ma, smA[buf], cur_pm, cur_tt);
        s1_mat_file(mb, smB[buf], cur_pm, cur_tt);
      }
      if (one) s1_fetch(o, slot_x(buf), &pre_sc[buf][threadIdx.x], A.a, A.b, last_dst, n_rows, p_own, Ppad, codes, clv, scl);
      __syncthreads();
    }
    if (one) s1_cp_wait();  // this action's fetched column (the thread's own slots: no block barrier)
    S1Mat na, nb;
    na.how = nb.how = 0;
    S1Pre q;
    q.ca = q.cb = q.which = 0;
    if (has_next) {
      if (N.kind == 0) {
        const int moving = A.kind == 1 ? A.branch : -1;
        s1_mat_fetch(na, N.a < n_rows, N.pa, moving, moved, pmat, tiptab, priv);
        s1_mat_fetch(nb, N.b < n_rows, N.pb, moving, moved, pmat, tiptab, priv);
      }
      if (one)
        s1_fetch(q, slot_x(buf ^ 1), &pre_sc[buf ^ 1][threadIdx.x], N.a, N.b, A.kind == 0 ? A.dst : last_dst, n_rows, p_own, Ppad,
                 codes, clv, scl);
    }
    if (A.kind == 0) update(A, smA[buf], smB[buf], o, slot_x(buf), &pre_sc[buf][threadIdx.x]);
    else if (newton(A, o, slot_x(buf), &pre_sc[buf][threadIdx.x])) changed = true;
    if (has_next) {
      s1_mat_file(na, smA[buf ^ 1], cur_pm, cur_tt);
      s1_mat_file(nb, smB[buf ^ 1], cur_pm, cur_tt);
      if (one) o = q;
    }
    __syncthreads();  // the staging buffers, cur_pm / cur_tt, bl, moved
    buf ^= 1;
    staged = has_next;
  };

  for (int i = 0; i < n_down; i++) run(alist[i], i + 1 < n_down + n_pass, alist[i + 1]);
  for (int pass = 0; pass < max_passes; pass++) {
    changed = false;
    // (what follows a pass's last action depends on `changed`: that one starts unstaged)
    for (int i = 0; i < n_pass; i++) run(alist[n_down + i], i + 1 < n_pass, alist[n_down + i + 1]);
    if (!changed) break;
  }
  {  // lnL at the final lengths
    const S1Act op = alist[n_down + n_pass];
    double sv[16];
    int svsc = 0;
    if (one) {
      s1_fetch(o, slot_x(0), &pre_sc[0][threadIdx.x], op.a, op.b, last_dst, n_rows, p_own, Ppad, codes, clv, scl);
      s1_cp_wait();
      column(op, o, slot_x(0), &pre_sc[0][threadIdx.x], p_own, sv, svsc);
    }
    const double z = fmin(fmax(exp(-bl[op.branch]), PLL_ZMIN), PLL_ZMAX);
    double s0, s1, s2;
    sums(op, o, sv, svsc, log(fmax(z, PLL_ZMIN)), true, s0, s1, s2);
    if (blockIdx.x == 0) {
      for (int e = threadIdx.x; e < n_br; e += S1_THREADS) brlen[e] = bl[e];
      if (threadIdx.x == 0) out[0] = s0;
    }
  }
}

}  // namespace plg
