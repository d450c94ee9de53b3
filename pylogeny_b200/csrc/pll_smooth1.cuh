// Branch-length smoothing of ONE tree in ONE cooperative launch (DNA x Gamma-4).
//
// What the reference asks per tree (libpllWrapper.c:205: pllOptimizeBranchLengths(.., 3) on default
// lengths, then the lnL; scoring.py:41-75 makes one instance per tree, heuristic.py:107-260 calls it
// per neighbour through landscape.vertex.scoreLikelihood) is a chain of ~5 n dependent steps per pass:
// re-orient a conditional vector, or one Newton step on a branch.  As separate launches (the lockstep
// program of pll_loglik_batch with a batch of one) that is ~930 launches per 64-taxon tree and ~10 ms,
// all of it launch gaps.  But the steps depend on each other only THROUGH THE BRANCH LENGTHS: a vector
// update is pattern-local, and a Newton step needs three sums over all patterns.  So one persistent
// grid -- thread = pattern, block = 128 patterns -- runs the whole program; vector updates need no
// synchronisation at all, a Newton step is: per-thread sum-table column in REGISTERS, block partials,
// one grid barrier, every block adds the partials in the same fixed order and takes the same step
// (the curvature guard re-evaluates and costs one more barrier per retry).  The
// moved branch's P-matrix is recomputed by every block into its own shared memory for immediate use;
// block 0 also publishes it for the later steps (one barrier later at the earliest).
//
// A block covers tiles blockIdx.x, blockIdx.x + gridDim.x, ... of 128 patterns, so any alignment length
// runs on a co-resident grid (a curvature-guard retry recomputes the sum-table columns instead of keeping
// them in registers across the barrier).
//
// Loads of data written inside the kernel (vectors, counters, matrices) are ld.global.cg: the
// read-only (.nc) path the other kernels use is not coherent with stores of the same launch.
#pragma once

namespace plg {

#ifndef PLG_S1_THREADS
#define PLG_S1_THREADS 128
#endif
constexpr int S1_THREADS = PLG_S1_THREADS;

struct S1Act {
  int kind;              // 0: vector update (dst = a x b through pa, pb); 1: Newton step on `branch` between a and b
  int dst, a, b, pa, pb;
  int branch, pad;
};

__device__ __forceinline__ void s1_stage(double *sm, bool tip, int pm, int cur_branch, const double *cur_pm,
                                         const double *cur_tt, const double *pmat, const double *tiptab) {
  if (tip) {
    if (pm == cur_branch) {
      for (int e = threadIdx.x; e < 256; e += S1_THREADS) sm[e] = cur_tt[e];
    } else {
      const double *src = tiptab + (int64_t)pm * 256;
      for (int e = threadIdx.x; e < 256; e += S1_THREADS) sm[e] = __ldcg(src + e);
    }
  } else {
    if (pm == cur_branch) {
      for (int e = threadIdx.x; e < 64; e += S1_THREADS) sm[e] = cur_pm[(e >> 4) * 17 + (e & 15)];
    } else {
      const double *src = pmat + (int64_t)pm * 68;
      for (int e = threadIdx.x; e < 64; e += S1_THREADS) sm[e] = __ldcg(src + (e >> 4) * 17 + (e & 15));
    }
  }
}

// v (*)= operand `id` seen through the staged matrix / tip table (sm); sc += its scaling counter
template <bool MUL>
__device__ __forceinline__ void s1_operand(const double *sm, int id, int n_rows, int64_t p, int64_t Ppad,
                                           const unsigned char *__restrict__ codes, const double *clv, const int *scl,
                                           double (&v)[16], int &sc) {
  if (id < n_rows) {
    const int code = codes[(int64_t)id * Ppad + p];
#pragma unroll
    for (int i = 0; i < 16; i++) {
      const double t = sm[code * 16 + i];
      if (MUL) v[i] *= t; else v[i] = t;
    }
  } else {
    const double *base = clv + (int64_t)(id - n_rows) * 16 * Ppad + p;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = __ldcg(base + (int64_t)i * Ppad);
    sc += __ldcg(scl + (int64_t)(id - n_rows) * Ppad + p);
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
}

// v (*)= P x for a column that is already in registers (the previous action's result, see `fwd` below)
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

// Grid barrier: one monotonic counter, generation g waits for g * gridDim.x arrivals.  Hand-rolled instead of
// cooperative_groups' grid.sync() because the launch runs ~380 of them back to back on a dependent chain
// (the launch is still cooperative: co-residency is what makes spinning safe).
__device__ __forceinline__ void s1_grid_barrier(unsigned int *counter, unsigned int &generation) {
  __syncthreads();
  if (threadIdx.x == 0) {
    generation++;
    const unsigned int target = generation * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned int seen;
    do {
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];\n" : "=r"(seen) : "l"(counter) : "memory");
    } while (seen < target);
  }
  __syncthreads();
}

// operand as it sits at its node (tips: indicator of the code's states)
__device__ __forceinline__ void s1_raw(int id, int n_rows, int64_t p, int64_t Ppad, const unsigned char *__restrict__ codes,
                                       const unsigned int *__restrict__ masks, const double *clv, const int *scl,
                                       double (&v)[16], int &sc) {
  if (id < n_rows) {
    const unsigned int m = masks[codes[(int64_t)id * Ppad + p]];
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = (m >> (i & 3) & 1) ? 1.0 : 0.0;
  } else {
    const double *base = clv + (int64_t)(id - n_rows) * 16 * Ppad + p;
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] = __ldcg(base + (int64_t)i * Ppad);
    sc += __ldcg(scl + (int64_t)(id - n_rows) * Ppad + p);
  }
}

// acts[0 .. n_down): the down pass; acts[n_down .. n_down + n_pass): one smoothing pass (run up to
// max_passes times, until a pass moves no branch by more than PLL_DELTAZ in z); acts[n_down + n_pass]:
// the evaluation step (kind 1: its sums at the branch's final length give the lnL).
// dpart: 2 x gridDim.x x 3 doubles; out[0] = lnL.
__global__ void __launch_bounds__(S1_THREADS)
smooth1_kernel(const S1Act *__restrict__ acts, int n_down, int n_pass, int max_passes, int n_rows, int64_t Ppad,
               const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
               const double *__restrict__ weights, const ModelDev *__restrict__ model, double *pmat, double *tiptab,
               double *clv, int *scl, double *brlen, double *dpart, double *out, unsigned int *bar, int acts_smem) {
  __shared__ __align__(16) double smA[256];
  __shared__ __align__(16) double smB[256];
  __shared__ double cur_pm[68], cur_tt[256];
  __shared__ double mU[16], mUinv[16], mg[16], mpi[4];
  __shared__ double tab[3][16];
  __shared__ double red[3][S1_THREADS / 32];
  __shared__ double tot[3];
  const int n_tiles = (int)(Ppad / S1_THREADS);
  // the action list in shared memory when it fits (acts_smem != 0): every action starts with its record, and a
  // global load there is one more round trip on the dependent chain of ~1,300 actions
  extern __shared__ __align__(16) unsigned char s1_dyn[];
  const S1Act *alist = acts;
  if (acts_smem) {
    const int n_all = n_down + n_pass + 1;
    int4 *d = reinterpret_cast<int4 *>(s1_dyn);
    const int4 *g = reinterpret_cast<const int4 *>(acts);
    for (int e = threadIdx.x; e < n_all * 2; e += S1_THREADS) d[e] = g[e];
    alist = reinterpret_cast<const S1Act *>(s1_dyn);
  }
  if (threadIdx.x < 16) {
    mU[threadIdx.x] = model->U[threadIdx.x];          // [i][k]  (LIK_MAX_STATES-strided rows are not used: K = 4 packs them)
    mUinv[threadIdx.x] = model->Uinv[threadIdx.x];    // [k][j]
    mg[threadIdx.x] = -model->eval[threadIdx.x & 3] * model->rates[threadIdx.x >> 2];
    if (threadIdx.x < 4) mpi[threadIdx.x] = model->freqs[threadIdx.x];
  }
  __syncthreads();
  int cur_branch = -1, parity = 0;
  unsigned int generation = 0;
  // Forwarding: with one tile per block a thread owns ONE pattern for the whole launch, and most actions
  // read the vector the previous action wrote (the re-orientation chain along the depth-first path, the
  // Newton step on the branch just re-oriented toward): that column stays in registers.
  const bool fwd = n_tiles <= (int)gridDim.x;
  int last_dst = -1, last_sc = 0;
  double last_v[16];
#pragma unroll
  for (int i = 0; i < 16; i++) last_v[i] = 0.0;

  auto update = [&](const S1Act &op) {
    s1_stage(smA, op.a < n_rows, op.pa, cur_branch, cur_pm, cur_tt, pmat, tiptab);
    s1_stage(smB, op.b < n_rows, op.pb, cur_branch, cur_pm, cur_tt, pmat, tiptab);
    __syncthreads();
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t p = (int64_t)tile * S1_THREADS + threadIdx.x;
      double v[16];
      int sc = 0;
      if (fwd && op.a == last_dst) { s1_apply_regs<false>(smA, last_v, v); sc += last_sc; }
      else s1_operand<false>(smA, op.a, n_rows, p, Ppad, codes, clv, scl, v, sc);
      if (fwd && op.b == last_dst) { s1_apply_regs<true>(smB, last_v, v); sc += last_sc; }
      else s1_operand<true>(smB, op.b, n_rows, p, Ppad, codes, clv, scl, v, sc);
      int hi = 0;
#pragma unroll
      for (int i = 0; i < 16; i++) hi = max(hi, __double2hiint(v[i]));
      if (hi < 0x2FF00000) {  // every entry below 2^-256 (values are >= 0): rescale the column, count it
#pragma unroll
        for (int i = 0; i < 16; i++) v[i] *= 1.15792089237316195423570985008687907853269984665640564039457584007913129639936e77;
        sc += 1;
      }
      double *dst = clv + (int64_t)(op.dst - n_rows) * 16 * Ppad + p;
#pragma unroll
      for (int i = 0; i < 16; i++) dst[(int64_t)i * Ppad] = v[i];
      scl[(int64_t)(op.dst - n_rows) * Ppad + p] = sc;
      if (fwd) {
#pragma unroll
        for (int i = 0; i < 16; i++) last_v[i] = v[i];
        last_sc = sc;
      }
    }
    last_dst = op.dst;
    __syncthreads();  // smA / smB are restaged by the next step
  };

  // sum-table column of pattern p for the branch between operands a and b
  auto column = [&](int a_id, int b_id, int64_t p, double (&sv)[16], int &sc) {
    double A[16], B[16];
    sc = 0;
    if (fwd && a_id == last_dst) {
#pragma unroll
      for (int i = 0; i < 16; i++) A[i] = last_v[i];
      sc += last_sc;
    } else {
      s1_raw(a_id, n_rows, p, Ppad, codes, masks, clv, scl, A, sc);
    }
    if (fwd && b_id == last_dst) {
#pragma unroll
      for (int i = 0; i < 16; i++) B[i] = last_v[i];
      sc += last_sc;
    } else {
      s1_raw(b_id, n_rows, p, Ppad, codes, masks, clv, scl, B, sc);
    }
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

  // sums (lnL, d1, d2) over all patterns for the branch between a and b at log z = lz, identical in every
  // thread of the grid
  auto sums = [&](int a_id, int b_id, double lz, double &s0, double &s1, double &s2) {
    if (threadIdx.x < 16) {
      const double g = mg[threadIdx.x], e = exp(g * lz);
      tab[0][threadIdx.x] = e; tab[1][threadIdx.x] = g * e; tab[2][threadIdx.x] = g * g * e;
    }
    __syncthreads();
    double v0 = 0.0, v1 = 0.0, v2 = 0.0;
    for (int tile = blockIdx.x; tile < n_tiles; tile += gridDim.x) {
      const int64_t p = (int64_t)tile * S1_THREADS + threadIdx.x;
      const double w = weights[p];
      double sv[16];
      int sc;
      column(a_id, b_id, p, sv, sc);
      double l0 = 0.0, l1 = 0.0, l2 = 0.0;
#pragma unroll
      for (int e = 0; e < 16; e++) {
        l0 = fma(sv[e], tab[0][e], l0);
        l1 = fma(sv[e], tab[1][e], l1);
        l2 = fma(sv[e], tab[2][e], l2);
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
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = v0; red[1][threadIdx.x >> 5] = v1; red[2][threadIdx.x >> 5] = v2; }
    __syncthreads();
    double *dp = dpart + (size_t)parity * gridDim.x * 3;
    if (threadIdx.x < 3) {
      double s = 0.0;
#pragma unroll
      for (int i = 0; i < S1_THREADS / 32; i++) s += red[threadIdx.x][i];
      dp[(size_t)threadIdx.x * gridDim.x + blockIdx.x] = s;
    }
    s1_grid_barrier(bar, generation);
    if (threadIdx.x < 32) {  // one warp adds the blocks' shares: strided, then a fixed shuffle tree
      double a0 = 0.0, a1 = 0.0, a2 = 0.0;
      for (unsigned b = threadIdx.x; b < gridDim.x; b += 32) {
        a0 += __ldcg(dp + b); a1 += __ldcg(dp + gridDim.x + b); a2 += __ldcg(dp + 2 * (size_t)gridDim.x + b);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) {
        a0 += __shfl_down_sync(0xffffffffu, a0, d);
        a1 += __shfl_down_sync(0xffffffffu, a1, d);
        a2 += __shfl_down_sync(0xffffffffu, a2, d);
      }
      if (threadIdx.x == 0) { tot[0] = a0; tot[1] = a1; tot[2] = a2; }
    }
    __syncthreads();
    s0 = tot[0]; s1 = tot[1]; s2 = tot[2];
    parity ^= 1;
    __syncthreads();  // tab / red / tot are reused by the next evaluation
  };

  for (int i = 0; i < n_down; i++) update(alist[i]);
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (int pass = 0; pass < max_passes; pass++) {
    bool changed = false;
    for (int i = 0; i < n_pass; i++) {
      const S1Act op = alist[n_down + i];
      if (op.kind == 0) { update(op); continue; }
      // one makenewz step (libpll: optimise the branch with maxiter = 1), as newton_step_block
      const double z0 = exp(-__ldcg(brlen + op.branch));
      double z = fmin(fmax(z0, PLL_ZMIN), PLL_ZMAX), zprev = z;
      for (int guard = 0; guard < 64; guard++) {  // "bad curvature: shorten the branch"
        sums(op.a, op.b, log(fmax(z, PLL_ZMIN)), s0, s1, s2);
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
      if (fabs(z - z0) > PLL_DELTAZ) changed = true;
      const double t_new = -log(z);
      // the moved branch's matrices: here for the steps that follow, published by block 0 for later
      pmatrix_block<4, 4>(model, t_new, masks, cur_pm, cur_tt);
      __syncthreads();
      cur_branch = op.branch;
      if (blockIdx.x == 0) {
        for (int e = threadIdx.x; e < 68; e += S1_THREADS) pmat[(int64_t)op.branch * 68 + e] = cur_pm[e];
        for (int e = threadIdx.x; e < 256; e += S1_THREADS) tiptab[(int64_t)op.branch * 256 + e] = cur_tt[e];
        if (threadIdx.x == 0) brlen[op.branch] = t_new;
      }
    }
    if (!changed) break;
  }
  {  // lnL at the final lengths
    const S1Act op = alist[n_down + n_pass];
    s1_grid_barrier(bar, generation);  // (block 0's last brlen store)
    const double z = fmin(fmax(exp(-__ldcg(brlen + op.branch)), PLL_ZMIN), PLL_ZMAX);
    sums(op.a, op.b, log(fmax(z, PLL_ZMIN)), s0, s1, s2);
    if (blockIdx.x == 0 && threadIdx.x == 0) out[0] = s0;
  }
}

}  // namespace plg
