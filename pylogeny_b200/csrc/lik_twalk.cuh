// DNA x Gamma-4 tree walk: a block of 128 threads (one pattern each) evaluates the WHOLE post-order
// of a tree for its 128 patterns (treewalk.h).  The running CLV column stays in registers, the
// few parked ones in a per-block stack that is reused by every tree the block scores (so it lives
// in L2), and the two P-matrices of a node-op are staged once per block, double-buffered, while
// the previous node-op computes.  HBM supplies one byte per tip per pattern; what is left is FP64
// arithmetic (2 x 64 FMA per inner operand) -- against 2(n-2)(C+4) bytes per pattern per tree for
// the level-synchronous kernels.
//
// Work items (pattern tile, tree) are drawn tile-major from an atomic counter by a persistent
// grid: neighbouring blocks score different trees on the same tile, so tip codes are L2 hits.
// The per-tile partial sums are written to partial[tree][tile] and reduced in a fixed order.
#pragma once

namespace plg {

constexpr int TW_THREADS = 128;
#ifndef PLG_TW_MINB
#define PLG_TW_MINB 4
#endif

// P(t r_c) of EVERY child slot of every tree straight from the descriptor's branch lengths:
// pmat[slot][r][k][j], 64 doubles per slot.  The two slots of a root hold P(ta + tb) (evaluation
// across the root branch, as build_full_lik_program) and nothing.
__global__ void pmatrix4_slots_kernel(const ModelDev *__restrict__ model, const double *__restrict__ brlen,
                                      int64_t n_slots, int slots_per_tree, double *__restrict__ pmat) {
  const int64_t slot = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 6);
  if (slot >= n_slots) return;
  const int e = threadIdx.x & 63;
  const int q = (int)(slot % slots_per_tree);
  double t = brlen[slot];
  if (q == slots_per_tree - 2) t += brlen[slot + 1];
  if (q == slots_per_tree - 1 || !(t >= 0.0)) t = 0.0;
  const int r = e >> 4, i = (e >> 2) & 3, j = e & 3;
  double s = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) s += model->U[i * 4 + k] * exp(model->eval[k] * t * model->rates[r]) * model->Uinv[k * 4 + j];
  s = fmax(s, 0.0);
  if (t == 0.0) s = (i == j) ? 1.0 : 0.0;
  pmat[slot * 64 + e] = s;
}

// both matrices of a node-op (child slots `slot`, `slot + 1`): P as [r][k][j] and, for tip
// operands, its transpose [r][j][k] (a tip's column of P as two 16-byte reads per category)
__device__ __forceinline__ void tw_stage(double *buf, const double *__restrict__ pm, int slot) {
  const int m = threadIdx.x >> 6, e = threadIdx.x & 63;
  const double x = __ldg(pm + ((int64_t)slot + m) * 64 + e);
  buf[m * 128 + e] = x;
  buf[m * 128 + 64 + (e & 48) + ((e & 3) << 2) + ((e >> 2) & 3)] = x;
}

// v (*)= P x for one operand of a node-op
template <bool MUL>
__device__ __forceinline__ void tw_operand(int id, const double *M, int64_t p, int64_t Ppad,
                                           const unsigned char *__restrict__ codes, const double (&top)[16], int sct,
                                           const double *stk, const int *stks, double (&v)[16], int &sc) {
  if (id >= 0) {
    const int code = codes[(int64_t)id * Ppad + p];
    const int mask = code ? code : 15;
    if (__all_sync(0xffffffffu, (mask & (mask - 1)) == 0)) {  // unambiguous: one column of P
      const double *col = M + 64 + ((__ffs(mask) - 1) << 2);
#pragma unroll
      for (int r = 0; r < 4; r++) {
        const double2 t01 = *reinterpret_cast<const double2 *>(col + r * 16);
        const double2 t23 = *reinterpret_cast<const double2 *>(col + r * 16 + 2);
        if (MUL) { v[r * 4] *= t01.x; v[r * 4 + 1] *= t01.y; v[r * 4 + 2] *= t23.x; v[r * 4 + 3] *= t23.y; }
        else { v[r * 4] = t01.x; v[r * 4 + 1] = t01.y; v[r * 4 + 2] = t23.x; v[r * 4 + 3] = t23.y; }
      }
    } else {
      double x[16];
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = (mask >> (i & 3) & 1) ? 1.0 : 0.0;
      v4_apply_reg<MUL>(M, x, v);
    }
  } else if (id == -1) {
    v4_apply_reg<MUL>(M, top, v);
    sc += sct;
  } else {
    const int l = -2 - id;
    double x[16];
#pragma unroll
    for (int i = 0; i < 16; i++) x[i] = stk[(l * 16 + i) * TW_THREADS];
    sc += stks[l * TW_THREADS];
    v4_apply_reg<MUL>(M, x, v);
  }
}

__global__ void __launch_bounds__(TW_THREADS, PLG_TW_MINB)
twalk4_kernel(const WalkOp *__restrict__ ops, int ni, int n_trees, unsigned long long n_items, int64_t Ppad,
              const unsigned char *__restrict__ codes, const double *__restrict__ weights,
              const ModelDev *__restrict__ model, const double *__restrict__ pmat, double *__restrict__ stack,
              int *__restrict__ stack_scl, int levels, double *__restrict__ partial, int pstride,
              unsigned long long *counter) {
  __shared__ __align__(16) double mats[2][256];  // [buffer][a: P, P^T | b: P, P^T]
  __shared__ unsigned long long s_item;
  __shared__ double wsum[TW_THREADS / 32];
  double *stk = stack + (int64_t)blockIdx.x * levels * 16 * TW_THREADS + threadIdx.x;
  int *stks = stack_scl + (int64_t)blockIdx.x * levels * TW_THREADS + threadIdx.x;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1ull);
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= n_items) break;
    const int64_t tile = (int64_t)(item / (unsigned)n_trees);
    const int tree = (int)(item - (unsigned long long)tile * (unsigned)n_trees);
    const int64_t p = tile * TW_THREADS + threadIdx.x;
    const int4 *op_p = reinterpret_cast<const int4 *>(ops + (int64_t)tree * ni);
    const double *pm = pmat + (int64_t)tree * 2 * ni * 64;
    int4 nxt = __ldg(op_p);
    tw_stage(mats[0], pm, nxt.w);
    const double w = weights[p];
    __syncthreads();
    double top[16];
    int sct = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) top[i] = 0.0;
    double val = 0.0;
    for (int i = 0; i < ni; i++) {
      const int4 op = nxt;
      if (i + 1 < ni) {
        nxt = __ldg(op_p + i + 1);
        tw_stage(mats[(i + 1) & 1], pm, nxt.w);
      }
      const double *MA = mats[i & 1], *MB = mats[i & 1] + 128;
      double v[16];
      int sc = 0;
      if (i + 1 < ni) {
        tw_operand<false>(op.x, MA, p, Ppad, codes, top, sct, stk, stks, v, sc);
        tw_operand<true>(op.y, MB, p, Ppad, codes, top, sct, stk, stks, v, sc);
        sc += wk_rescale(v);
        if (op.z >= 0) {
#pragma unroll
          for (int q = 0; q < 16; q++) stk[(op.z * 16 + q) * TW_THREADS] = v[q];
          stks[op.z * TW_THREADS] = sc;
        } else {
#pragma unroll
          for (int q = 0; q < 16; q++) top[q] = v[q];
          sct = sc;
        }
      } else {
        // root: site likelihood across the root branch, sum_k pi_k a_k (P(ta+tb) b)_k
        tw_operand<false>(op.y, MA, p, Ppad, codes, top, sct, stk, stks, v, sc);
        if (op.x >= 0) {
          const int code = codes[(int64_t)op.x * Ppad + p];
          const int mask = code ? code : 15;
#pragma unroll
          for (int q = 0; q < 16; q++) v[q] = (mask >> (q & 3) & 1) ? v[q] : 0.0;
        } else if (op.x == -1) {
#pragma unroll
          for (int q = 0; q < 16; q++) v[q] *= top[q];
          sc += sct;
        } else {
          const int l = -2 - op.x;
#pragma unroll
          for (int q = 0; q < 16; q++) v[q] *= stk[(l * 16 + q) * TW_THREADS];
          sc += stks[l * TW_THREADS];
        }
        val = v4_site(v, sc, w, model);
      }
      __syncthreads();
    }
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
    if ((threadIdx.x & 31) == 0) wsum[threadIdx.x >> 5] = val;
    __syncthreads();
    if (threadIdx.x == 0) partial[(int64_t)tree * pstride + tile] = (wsum[0] + wsum[1]) + (wsum[2] + wsum[3]);
  }
}

}  // namespace plg
