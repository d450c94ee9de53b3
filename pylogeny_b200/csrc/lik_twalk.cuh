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

// v (*)= P x for the four categories of one pattern: mat = P as [r][k][j] in shared memory, x in registers
template <bool MUL>
__device__ __forceinline__ void v4_apply_reg(const double *mat, const double (&x)[16], double (&v)[16]) {
#pragma unroll
  for (int r = 0; r < 4; r++)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double2 p01 = *reinterpret_cast<const double2 *>(mat + r * 16 + k * 4);
      const double2 p23 = *reinterpret_cast<const double2 *>(mat + r * 16 + k * 4 + 2);
      double s = p01.x * x[r * 4];
      s = fma(p01.y, x[r * 4 + 1], s);
      s = fma(p23.x, x[r * 4 + 2], s);
      s = fma(p23.y, x[r * 4 + 3], s);
      if (MUL) v[r * 4 + k] *= s; else v[r * 4 + k] = s;
    }
}

// weighted site log-likelihood of one pattern from its 16 joined entries e[r * 4 + k]
__device__ __forceinline__ double v4_site(const double (&e)[16], int sc, double w, const ModelDev *__restrict__ model) {
  double term = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) term = fma(model->freqs[k], (e[k] + e[4 + k]) + (e[8 + k] + e[12 + k]), term);
  return w != 0.0 ? w * (log(term * 0.25) + sc * LOG_MINLH) : 0.0;
}

// per pattern: all 16 entries below 2^-256 -> multiply by 2^256 and count (libpll's scaling)
__device__ __forceinline__ int tw_rescale(double (&n)[16]) {
  int hi = 0;
#pragma unroll
  for (int i = 0; i < 16; i++) hi = max(hi, __double2hiint(n[i]));
  if (hi >= 0x2FF00000) return 0;  // high word of 2^-256 (biased exponent 1023 - 256 = 0x2FF)
#pragma unroll
  for (int i = 0; i < 16; i++) n[i] *= TWO256;
  return 1;
}

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

// Both matrices of a node-op (child slots `slot`, `slot + 1`): each thread fetches one element
// early (tw_fetch, while the previous node-op computes) and files it late (tw_file): P as
// [r][k][j] and, for tip operands, its transpose [r][j][k] (a tip's column of P is then two
// 16-byte reads per category).
__device__ __forceinline__ double tw_fetch(const double *__restrict__ pm, int slot) {
  return __ldg(pm + ((int64_t)slot + (threadIdx.x >> 6)) * 64 + (threadIdx.x & 63));
}
__device__ __forceinline__ void tw_file(double *buf, double x) {
  const int m = threadIdx.x >> 6, e = threadIdx.x & 63;
  buf[m * 128 + e] = x;
  buf[m * 128 + 64 + (e & 48) + ((e & 3) << 2) + ((e >> 2) & 3)] = x;
}

constexpr int TW_SMEM_LEVELS = 2;  // stack levels kept in shared memory (the rest: global, L2-resident)

// parked CLV column `l`: shared memory for the lowest levels, the block's global stack above
struct TwStack {
  double *sm;   // [TW_SMEM_LEVELS][16][TW_THREADS] + this thread
  int *sms;     // [TW_SMEM_LEVELS][TW_THREADS] + this thread
  double *gl;   // global [levels][16][TW_THREADS] + this thread
  int *gls;
  __device__ __forceinline__ void load(int l, double (&x)[16], int &sc) const {
    if (l < TW_SMEM_LEVELS) {
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = sm[(l * 16 + i) * TW_THREADS];
      sc += sms[l * TW_THREADS];
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) x[i] = gl[(l * 16 + i) * TW_THREADS];
      sc += gls[l * TW_THREADS];
    }
  }
  __device__ __forceinline__ void store(int l, const double (&x)[16], int sc) const {
    if (l < TW_SMEM_LEVELS) {
#pragma unroll
      for (int i = 0; i < 16; i++) sm[(l * 16 + i) * TW_THREADS] = x[i];
      sms[l * TW_THREADS] = sc;
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++) gl[(l * 16 + i) * TW_THREADS] = x[i];
      gls[l * TW_THREADS] = sc;
    }
  }
};

// v (*)= P x for one operand of a node-op; `code` = the tip's code when id >= 0 (fetched one
// node-op ahead)
template <bool MUL>
__device__ __forceinline__ void tw_operand(int id, int code, const double *M, const double (&top)[16], int sct,
                                           const TwStack &stk, double (&v)[16], int &sc) {
  if (id >= 0) {
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
    double x[16];
    stk.load(-2 - id, x, sc);
    v4_apply_reg<MUL>(M, x, v);
  }
}

constexpr int TW_SMEM_BYTES = (2 * 256 + TW_SMEM_LEVELS * 16 * TW_THREADS) * 8 + TW_SMEM_LEVELS * TW_THREADS * 4;

__global__ void __launch_bounds__(TW_THREADS, PLG_TW_MINB)
twalk4_kernel(const WalkOp *__restrict__ ops, int ni, int n_trees, unsigned long long n_items, int64_t Ppad,
              const unsigned char *__restrict__ codes, const double *__restrict__ weights,
              const ModelDev *__restrict__ model, const double *__restrict__ pmat, double *__restrict__ stack,
              int *__restrict__ stack_scl, int levels, double *__restrict__ partial, int pstride,
              unsigned long long *counter) {
  extern __shared__ __align__(16) double tw_smem[];
  double(*mats)[256] = reinterpret_cast<double(*)[256]>(tw_smem);  // [buffer][a: P, P^T | b: P, P^T]
  __shared__ unsigned long long s_item;
  __shared__ double wsum[TW_THREADS / 32];
  TwStack stk;
  stk.sm = tw_smem + 2 * 256 + threadIdx.x;
  stk.sms = reinterpret_cast<int *>(tw_smem + 2 * 256 + TW_SMEM_LEVELS * 16 * TW_THREADS) + threadIdx.x;
  stk.gl = stack + (int64_t)blockIdx.x * levels * 16 * TW_THREADS + threadIdx.x;
  stk.gls = stack_scl + (int64_t)blockIdx.x * levels * TW_THREADS + threadIdx.x;
  for (;;) {
    if (threadIdx.x == 0) s_item = atomicAdd(counter, 1ull);
    __syncthreads();
    const unsigned long long item = s_item;
    if (item >= n_items) break;
    const int64_t tile = (int64_t)(item / (unsigned)n_trees);
    const int tree = (int)(item - (unsigned long long)tile * (unsigned)n_trees);
    const unsigned char *cp = codes + tile * TW_THREADS + threadIdx.x;
    const int4 *op_p = reinterpret_cast<const int4 *>(ops + (int64_t)tree * ni);
    const double *pm = pmat + (int64_t)tree * 2 * ni * 64;
    // software pipeline: descriptors two node-ops ahead, tip codes and matrix elements one ahead
    int4 op = __ldg(op_p);
    int4 nxt = ni > 1 ? __ldg(op_p + 1) : op;
    tw_file(mats[0], tw_fetch(pm, op.w));
    int ca = op.x >= 0 ? cp[(int64_t)op.x * Ppad] : 0, cb = op.y >= 0 ? cp[(int64_t)op.y * Ppad] : 0;
    const double w = weights[tile * TW_THREADS + threadIdx.x];
    __syncthreads();
    double top[16];
    int sct = 0;
#pragma unroll
    for (int i = 0; i < 16; i++) top[i] = 0.0;
    double val = 0.0;
    for (int i = 0; i < ni; i++) {
      const bool more = i + 1 < ni;
      int4 nxt2 = nxt;
      double melt = 0.0;
      int na = 0, nb = 0;
      if (more) {
        if (i + 2 < ni) nxt2 = __ldg(op_p + i + 2);
        melt = tw_fetch(pm, nxt.w);
        if (nxt.x >= 0) na = cp[(int64_t)nxt.x * Ppad];
        if (nxt.y >= 0) nb = cp[(int64_t)nxt.y * Ppad];
      }
      const double *MA = mats[i & 1], *MB = mats[i & 1] + 128;
      double v[16];
      int sc = 0;
      if (more) {
        tw_operand<false>(op.x, ca, MA, top, sct, stk, v, sc);
        tw_operand<true>(op.y, cb, MB, top, sct, stk, v, sc);
        sc += tw_rescale(v);
        if (op.z >= 0) {
          stk.store(op.z, v, sc);
        } else {
#pragma unroll
          for (int q = 0; q < 16; q++) top[q] = v[q];
          sct = sc;
        }
        tw_file(mats[(i + 1) & 1], melt);
      } else {
        // root: site likelihood across the root branch, sum_k pi_k a_k (P(ta+tb) b)_k
        tw_operand<false>(op.y, cb, MA, top, sct, stk, v, sc);
        if (op.x >= 0) {
          const int mask = ca ? ca : 15;
#pragma unroll
          for (int q = 0; q < 16; q++) v[q] = (mask >> (q & 3) & 1) ? v[q] : 0.0;
        } else if (op.x == -1) {
#pragma unroll
          for (int q = 0; q < 16; q++) v[q] *= top[q];
          sc += sct;
        } else {
          double x[16];
          stk.load(-2 - op.x, x, sc);
#pragma unroll
          for (int q = 0; q < 16; q++) v[q] *= x[q];
        }
        val = v4_site(v, sc, w, model);
      }
      op = nxt; nxt = nxt2; ca = na; cb = nb;
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
