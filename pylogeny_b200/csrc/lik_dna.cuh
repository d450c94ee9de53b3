// DNA (4 states) x Gamma-4 specialisation of the likelihood kernels: one thread per pattern owns
// all 16 entries (4 categories x 4 states) of a CLV column.
//
// Why a second formulation next to the generic (pattern, category)-per-thread kernels of lik.cu:
// ncu on the generic root evaluation (profiles/r1a_*.txt) showed it issue-bound, not HBM-bound --
// 10.5e9 warp instructions per launch, a third of them the FP64 log() executed by warps in which
// only 8 of 32 lanes hold a site sum.  Here a warp covers 32 patterns: 16 coalesced 256-byte loads
// per operand, P-matrix rows read from shared memory as warp-wide broadcasts, one log() per lane.
//
// The CLV update can carry a fused evaluation (edge insertion: the partial N just computed is
// joined with the far-side CLV and the pruned subtree's CLV while still in registers), which
// removes one of the six CLV transfers per SPR neighbour.
#pragma once

namespace plg {

// Block size / residency chosen by the sweep in profiles/tune_d4.sh (B200, 64 taxa x 10k sites):
// 128 x {5,6,7} blocks per SM plateau at ~5.1 TB/s algorithmic; 4 blocks (128 registers) gives 4.6.
#ifndef PLG_D4_THREADS
#define PLG_D4_THREADS 128
#endif
#ifndef PLG_D4_MINB
#define PLG_D4_MINB 6
#endif
constexpr int D4_THREADS = PLG_D4_THREADS;

__device__ __forceinline__ void d4_load16(const double *__restrict__ base, int64_t Ppad, double (&x)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i++) x[i] = __ldg(base + (int64_t)i * Ppad);
}

// stage P[4][16] (global stride 17 per category) or the 16x16 tip table into shared memory
__device__ __forceinline__ void d4_stage(double *sm, bool tip, int pm, const double *__restrict__ pmat,
                                         const double *__restrict__ tiptab) {
  if (tip) {
    const double *src = tiptab + (int64_t)pm * 256;
    for (int e = threadIdx.x; e < 256; e += D4_THREADS) sm[e] = src[e];
  } else {
    const double *src = pmat + (int64_t)pm * 68;
    for (int e = threadIdx.x; e < 64; e += D4_THREADS) sm[e] = src[(e >> 4) * 17 + (e & 15)];
  }
}

// v[r*4+k] (*)= sum_j P[r][k][j] x[r*4+j]; P rows are warp-uniform shared-memory broadcasts
template <bool MUL>
__device__ __forceinline__ void d4_apply(const double *sm, const double (&x)[16], double (&v)[16]) {
#pragma unroll
  for (int r = 0; r < 4; r++)
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

template <bool MUL>
__device__ __forceinline__ void d4_tip(const double *sm, int code, double (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const double2 t = *reinterpret_cast<const double2 *>(sm + code * 16 + i);
    if (MUL) { v[i] *= t.x; v[i + 1] *= t.y; } else { v[i] = t.x; v[i + 1] = t.y; }
  }
}

// One operand of a product: tip -> table row; inner -> P x.  `pm < 0` = identity.
template <bool MUL>
__device__ __forceinline__ void d4_operand(const double *sm, int id, int pm, int n_rows, int64_t p, int64_t Ppad,
                                           const unsigned char *__restrict__ codes,
                                           const unsigned int *__restrict__ masks, const double *__restrict__ clv,
                                           const int *__restrict__ scl, double (&v)[16], int &sc) {
  if (id < n_rows) {
    const int code = codes[(int64_t)id * Ppad + p];
    if (pm >= 0) {
      d4_tip<MUL>(sm, code, v);
    } else {
      const unsigned int m = masks[code];
#pragma unroll
      for (int i = 0; i < 16; i++) {
        const double ind = (m >> (i & 3) & 1) ? 1.0 : 0.0;
        if (MUL) v[i] *= ind; else v[i] = ind;
      }
    }
  } else {
    double x[16];
    d4_load16(clv + (int64_t)(id - n_rows) * 16 * Ppad + p, Ppad, x);
    sc += scl[(int64_t)(id - n_rows) * Ppad + p];
    if (pm >= 0) {
      d4_apply<MUL>(sm, x, v);
    } else {
#pragma unroll
      for (int i = 0; i < 16; i++)
        if (MUL) v[i] *= x[i]; else v[i] = x[i];
    }
  }
}

// w_p * (log site likelihood + scaler correction), summed over the block in a fixed order
__device__ __forceinline__ void d4_site_sum(const double (&v)[16], int sc, int64_t p,
                                            const double *__restrict__ weights,
                                            const ModelDev *__restrict__ model, double *warp_sums,
                                            double *__restrict__ out) {
  double term = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) term = fma(model->freqs[k], (v[k] + v[4 + k]) + (v[8 + k] + v[12 + k]), term);
  const double w = weights[p];
  double val = 0.0;
  if (w != 0.0) val = w * (log(term * 0.25) + sc * LOG_MINLH);
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
  if ((threadIdx.x & 31) == 0) warp_sums[threadIdx.x >> 5] = val;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
#pragma unroll
    for (int i = 0; i < D4_THREADS / 32; i++) s += warp_sums[i];
    *out = s;
  }
}

__global__ void __launch_bounds__(D4_THREADS, PLG_D4_MINB)
clv4_kernel(const LikOpF *__restrict__ ops, int blocks_per_op, int n_rows, int64_t Ppad,
            const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
            const double *__restrict__ weights, const ModelDev *__restrict__ model,
            const double *__restrict__ pmat, const double *__restrict__ tiptab, double *__restrict__ clv,
            int *__restrict__ scl, double *__restrict__ partial, int pstride) {
  __shared__ __align__(16) double smA[256];
  __shared__ __align__(16) double smB[256];
  __shared__ __align__(16) double smE[64 + 256 + 256];
  __shared__ double warp_sums[D4_THREADS / 32];
  const int op_idx = blockIdx.x / blocks_per_op;
  const int pb = blockIdx.x - op_idx * blocks_per_op;
  const LikOpF op = ops[op_idx];
  const int64_t p = (int64_t)pb * D4_THREADS + threadIdx.x;
  d4_stage(smA, op.a < n_rows, op.pa, pmat, tiptab);
  if (op.b >= 0) d4_stage(smB, op.b < n_rows, op.pb, pmat, tiptab);
  if (op.tree >= 0) {
    d4_stage(smE, false, op.e_pa, pmat, tiptab);
    d4_stage(smE + 64, op.e_b < n_rows, op.e_pb, pmat, tiptab);
    d4_stage(smE + 320, op.e_c < n_rows, op.e_pc, pmat, tiptab);
  }
  __syncthreads();
  double v[16];
  int sc = 0;
  d4_operand<false>(smA, op.a, op.pa, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  if (op.b >= 0) d4_operand<true>(smB, op.b, op.pb, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  double mx = 0.0;
#pragma unroll
  for (int i = 0; i < 16; i++) mx = fmax(mx, fabs(v[i]));
  if (mx < MINLH) {  // every entry below 2^-256: rescale the column, count it
#pragma unroll
    for (int i = 0; i < 16; i++) v[i] *= TWO256;
    sc += 1;
  }
  if (!op.pad) {  // pad = 1: nobody reads this CLV (dead-store elimination on the host)
    double *dst = clv + (int64_t)(op.dst - n_rows) * 16 * Ppad + p;
#pragma unroll
    for (int i = 0; i < 16; i++) dst[(int64_t)i * Ppad] = v[i];
    scl[(int64_t)(op.dst - n_rows) * Ppad + p] = sc;
  }
  if (op.tree >= 0) {
    double e[16];
    d4_apply<false>(smE, v, e);
    d4_operand<true>(smE + 64, op.e_b, op.e_pb, n_rows, p, Ppad, codes, masks, clv, scl, e, sc);
    d4_operand<true>(smE + 320, op.e_c, op.e_pc, n_rows, p, Ppad, codes, masks, clv, scl, e, sc);
    d4_site_sum(e, sc, p, weights, model, warp_sums, partial + (int64_t)op.tree * pstride + pb);
  }
}

__global__ void __launch_bounds__(D4_THREADS, PLG_D4_MINB)
eval4_kernel(const LikEval *__restrict__ evals, int blocks_per_eval, int n_rows, int64_t Ppad,
             const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
             const double *__restrict__ weights, const ModelDev *__restrict__ model,
             const double *__restrict__ pmat, const double *__restrict__ tiptab,
             const double *__restrict__ clv, const int *__restrict__ scl, double *__restrict__ partial, int pstride) {
  __shared__ __align__(16) double sm[3][256];
  __shared__ double warp_sums[D4_THREADS / 32];
  const int ev_idx = blockIdx.x / blocks_per_eval;
  const int pb = blockIdx.x - ev_idx * blocks_per_eval;
  const LikEval ev = evals[ev_idx];
  const int64_t p = (int64_t)pb * D4_THREADS + threadIdx.x;
  if (ev.pa >= 0) d4_stage(sm[0], ev.a < n_rows, ev.pa, pmat, tiptab);
  if (ev.pb >= 0) d4_stage(sm[1], ev.b < n_rows, ev.pb, pmat, tiptab);
  if (ev.c >= 0 && ev.pc >= 0) d4_stage(sm[2], ev.c < n_rows, ev.pc, pmat, tiptab);
  __syncthreads();
  double v[16];
  int sc = 0;
  d4_operand<false>(sm[0], ev.a, ev.pa, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  d4_operand<true>(sm[1], ev.b, ev.pb, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  if (ev.c >= 0) d4_operand<true>(sm[2], ev.c, ev.pc, n_rows, p, Ppad, codes, masks, clv, scl, v, sc);
  d4_site_sum(v, sc, p, weights, model, warp_sums, partial + (int64_t)ev.tree * pstride + pb);
}

}  // namespace plg
