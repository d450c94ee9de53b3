// 20-state x Gamma-4 specialisation of the likelihood kernels on the FP64 tensor cores.
//
// For amino-acid data the per-category product V_r = P_r (20x20) * X_r (20 x patterns) IS a dense
// contraction with the same P_r for every pattern, so it maps onto mma.sync.m8n8k4.f64 (DMMA):
// the P_r fragments live in registers for the whole tile, X_r streams through as the B operand.
// (tcgen05 has no FP64 kind; DMMA is the FP64 tensor path on sm_100a.)  The generic
// (pattern, category)-per-thread kernels of lik.cu need one shared-memory load per FMA and
// measured 1.7 TB/s algorithmic / ~7 TFLOP/s (profiles/r1_aa_before.txt); these replace them for
// K = 20, R = 4.
//
// Block = 4 warps = the 4 rate categories of a 64-pattern tile.  Fragment layouts (PTX ISA,
// m8n8k4 .f64): A a0 = A[lane>>2][lane&3]; B b0 = B[lane&3][lane>>2]; C c0,c1 = C[lane>>2][2*(lane&3)+{0,1}].
// States are padded 20 -> 24 rows (three m-tiles); K = 20 = five k-steps.
#pragma once

namespace plg {

#ifndef PLG_A20_UNROLL
#define PLG_A20_UNROLL 1
#endif
constexpr int A20_UNROLL = PLG_A20_UNROLL;
constexpr int A20_TILE = 64;     // patterns per block
constexpr int A20_THREADS = 128;  // 4 warps = 4 rate categories

__device__ __forceinline__ void dmma884(double (&c)[2], double a, double b) {
  asm("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
               : "+d"(c[0]), "+d"(c[1])
               : "d"(a), "d"(b));
}

// A fragments of P[pm][r] (global layout [r][401], row-major 20x20 inside)
__device__ __forceinline__ void a20_load_p(const double *__restrict__ pmat, int pm, int r, int lane, double (&a)[3][5]) {
  const double *P = pmat + ((int64_t)pm * 4 + r) * 401;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 3; mt++) {
    const int row = mt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < 5; ks++) a[mt][ks] = row < 20 ? __ldg(P + row * 20 + ks * 4 + t) : 0.0;
  }
}

// c[mt] (+)= P_r * X_r for the 8 patterns starting at p0; tips enter as 0/1 indicator columns
__device__ __forceinline__ void a20_apply(const double (&a)[3][5], int id, int n_rows, int r, int64_t p0, int lane,
                                          int64_t Ppad, const unsigned char *__restrict__ codes,
                                          const unsigned int *__restrict__ masks, const double *__restrict__ clv,
                                          double (&c)[3][2]) {
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 3; mt++) c[mt][0] = c[mt][1] = 0.0;
  if (id < n_rows) {
    const unsigned int m = masks[codes[(int64_t)id * Ppad + p0 + g]];
#pragma unroll
    for (int ks = 0; ks < 5; ks++) {
      const double b = (m >> (ks * 4 + t) & 1) ? 1.0 : 0.0;
#pragma unroll
      for (int mt = 0; mt < 3; mt++) dmma884(c[mt], a[mt][ks], b);
    }
  } else {
    const double *src = clv + ((int64_t)(id - n_rows) * 80 + r * 20 + t) * Ppad + p0 + g;
    double b[5];
#pragma unroll
    for (int ks = 0; ks < 5; ks++) b[ks] = __ldg(src + (int64_t)ks * 4 * Ppad);
#pragma unroll
    for (int ks = 0; ks < 5; ks++)
#pragma unroll
      for (int mt = 0; mt < 3; mt++) dmma884(c[mt], a[mt][ks], b[ks]);
  }
}

// an operand that sits at the node itself (no P): its values directly in C-fragment layout
__device__ __forceinline__ void a20_raw(int id, int n_rows, int r, int64_t p0, int lane, int64_t Ppad,
                                        const unsigned char *__restrict__ codes,
                                        const unsigned int *__restrict__ masks, const double *__restrict__ clv,
                                        double (&c)[3][2]) {
  const int g = lane >> 2, t = lane & 3;
  if (id < n_rows) {
    const unsigned int m0 = masks[codes[(int64_t)id * Ppad + p0 + 2 * t]];
    const unsigned int m1 = masks[codes[(int64_t)id * Ppad + p0 + 2 * t + 1]];
#pragma unroll
    for (int mt = 0; mt < 3; mt++) {
      const int row = mt * 8 + g;
      c[mt][0] = (row < 20 && (m0 >> row & 1)) ? 1.0 : 0.0;
      c[mt][1] = (row < 20 && (m1 >> row & 1)) ? 1.0 : 0.0;
    }
  } else {
#pragma unroll
    for (int mt = 0; mt < 3; mt++) {
      const int row = mt * 8 + g;
      if (row < 20) {
        const double2 v = *reinterpret_cast<const double2 *>(clv + ((int64_t)(id - n_rows) * 80 + r * 20 + row) * Ppad + p0 + 2 * t);
        c[mt][0] = v.x; c[mt][1] = v.y;
      } else {
        c[mt][0] = c[mt][1] = 0.0;
      }
    }
  }
}

// Largest high word among this lane group's entries of each of its two patterns (values are >= 0, so the
// high words order like the doubles: integer max and 32-bit shuffles instead of FP64 compares)
__device__ __forceinline__ void a20_colmax(const double (&c)[3][2], int g, int &m0, int &m1) {
  m0 = max(__double2hiint(c[0][0]), __double2hiint(c[1][0]));
  m1 = max(__double2hiint(c[0][1]), __double2hiint(c[1][1]));
  if (g < 4) {
    m0 = max(m0, __double2hiint(c[2][0]));
    m1 = max(m1, __double2hiint(c[2][1]));
  }
#pragma unroll
  for (int d = 4; d < 32; d <<= 1) {
    m0 = max(m0, __shfl_xor_sync(0xffffffffu, m0, d));
    m1 = max(m1, __shfl_xor_sync(0xffffffffu, m1, d));
  }
}
constexpr int A20_HI_MINLH = 0x2FF00000;  // high word of 2^-256

// The four categories of a pattern are rescaled together (all 80 entries below 2^-256: libpll's rule), so
// the four warps exchange their column maxima per 8-pattern tile and store the tile already scaled.
// (Round 1 stored first and repaired afterwards: on data where most patterns need a factor at the deeper
// nodes -- 50 random protein sequences sit at 2^-300 per site -- the repair pass re-read and re-wrote
// most of what the kernel had just written and took a third of its time, ncu profiles/r2n_*.)
__global__ void __launch_bounds__(A20_THREADS)
clv20_kernel(const LikOp *__restrict__ ops, int blocks_per_op, int n_rows, int64_t Ppad,
             const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
             const double *__restrict__ pmat, double *__restrict__ clv, int *__restrict__ scl) {
  __shared__ int smax[2][4][8];  // [tile parity][category][pattern of the tile]
  const int op_idx = blockIdx.x / blocks_per_op;
  const int pb = blockIdx.x - op_idx * blocks_per_op;
  const LikOp op = ops[op_idx];
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t pbase = (int64_t)pb * A20_TILE;
  double pa[3][5], pbm[3][5];
  a20_load_p(pmat, op.pa, r, lane, pa);
  if (op.b >= 0) a20_load_p(pmat, op.pb, r, lane, pbm);
  double *dst = clv + ((int64_t)(op.dst - n_rows) * 80 + r * 20) * Ppad;
  const bool booker = r == 0 && g == 0;  // lane t books the counters of patterns 2t, 2t + 1 of every tile
  const int *sa = op.a >= n_rows ? scl + (int64_t)(op.a - n_rows) * Ppad : nullptr;
  const int *sb = op.b >= n_rows ? scl + (int64_t)(op.b - n_rows) * Ppad : nullptr;
#pragma unroll(A20_UNROLL)
  for (int nt = 0; nt < A20_TILE / 8; nt++) {
    const int64_t p0 = pbase + nt * 8;
    int2 ca_ = make_int2(0, 0), cb_ = ca_;
    if (booker) {
      if (sa) ca_ = *reinterpret_cast<const int2 *>(sa + p0 + 2 * t);
      if (sb) cb_ = *reinterpret_cast<const int2 *>(sb + p0 + 2 * t);
    }
    double ca[3][2], cb[3][2];
    a20_apply(pa, op.a, n_rows, r, p0, lane, Ppad, codes, masks, clv, ca);
    if (op.b >= 0) {
      a20_apply(pbm, op.b, n_rows, r, p0, lane, Ppad, codes, masks, clv, cb);
#pragma unroll
      for (int mt = 0; mt < 3; mt++) { ca[mt][0] *= cb[mt][0]; ca[mt][1] *= cb[mt][1]; }
    }
    int m0, m1;
    a20_colmax(ca, g, m0, m1);
    int(*mx)[8] = smax[nt & 1];
    if (g == 0) *reinterpret_cast<int2 *>(&mx[r][2 * t]) = make_int2(m0, m1);
    __syncthreads();
    const int r0 = max(max(mx[0][2 * t], mx[1][2 * t]), max(mx[2][2 * t], mx[3][2 * t])) < A20_HI_MINLH ? 1 : 0;
    const int r1 = max(max(mx[0][2 * t + 1], mx[1][2 * t + 1]), max(mx[2][2 * t + 1], mx[3][2 * t + 1])) < A20_HI_MINLH ? 1 : 0;
    const double f0 = r0 ? TWO256 : 1.0, f1 = r1 ? TWO256 : 1.0;
#pragma unroll
    for (int mt = 0; mt < 3; mt++) {
      const int row = mt * 8 + g;
      if (row < 20)
        *reinterpret_cast<double2 *>(dst + (int64_t)row * Ppad + p0 + 2 * t) = make_double2(ca[mt][0] * f0, ca[mt][1] * f1);
    }
    if (booker)
      *reinterpret_cast<int2 *>(scl + (int64_t)(op.dst - n_rows) * Ppad + p0 + 2 * t) = make_int2(ca_.x + cb_.x + r0, ca_.y + cb_.y + r1);
  }
}

__global__ void __launch_bounds__(A20_THREADS)
eval20_kernel(const LikEval *__restrict__ evals, int blocks_per_eval, int n_rows, int64_t Ppad,
              const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
              const double *__restrict__ weights, const ModelDev *__restrict__ model,
              const double *__restrict__ pmat, const double *__restrict__ clv, const int *__restrict__ scl,
              double *__restrict__ partial) {
  __shared__ double sterm[4][A20_TILE];
  __shared__ double wsum[2];
  const int ev_idx = blockIdx.x / blocks_per_eval;
  const int pb = blockIdx.x - ev_idx * blocks_per_eval;
  const LikEval ev = evals[ev_idx];
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t pbase = (int64_t)pb * A20_TILE;
  double pa[3][5], pbm[3][5], pc[3][5];
  if (ev.pa >= 0) a20_load_p(pmat, ev.pa, r, lane, pa);
  if (ev.pb >= 0) a20_load_p(pmat, ev.pb, r, lane, pbm);
  if (ev.c >= 0 && ev.pc >= 0) a20_load_p(pmat, ev.pc, r, lane, pc);
  double fr[3];
#pragma unroll
  for (int mt = 0; mt < 3; mt++) fr[mt] = (mt * 8 + g) < 20 ? model->freqs[mt * 8 + g] : 0.0;
#pragma unroll(A20_UNROLL)
  for (int nt = 0; nt < A20_TILE / 8; nt++) {
    const int64_t p0 = pbase + nt * 8;
    double v[3][2], x[3][2];
    if (ev.pa >= 0) a20_apply(pa, ev.a, n_rows, r, p0, lane, Ppad, codes, masks, clv, v);
    else a20_raw(ev.a, n_rows, r, p0, lane, Ppad, codes, masks, clv, v);
    if (ev.pb >= 0) a20_apply(pbm, ev.b, n_rows, r, p0, lane, Ppad, codes, masks, clv, x);
    else a20_raw(ev.b, n_rows, r, p0, lane, Ppad, codes, masks, clv, x);
#pragma unroll
    for (int mt = 0; mt < 3; mt++) { v[mt][0] *= x[mt][0]; v[mt][1] *= x[mt][1]; }
    if (ev.c >= 0) {
      if (ev.pc >= 0) a20_apply(pc, ev.c, n_rows, r, p0, lane, Ppad, codes, masks, clv, x);
      else a20_raw(ev.c, n_rows, r, p0, lane, Ppad, codes, masks, clv, x);
#pragma unroll
      for (int mt = 0; mt < 3; mt++) { v[mt][0] *= x[mt][0]; v[mt][1] *= x[mt][1]; }
    }
    double s0 = 0.0, s1 = 0.0;
#pragma unroll
    for (int mt = 0; mt < 3; mt++) { s0 = fma(fr[mt], v[mt][0], s0); s1 = fma(fr[mt], v[mt][1], s1); }
#pragma unroll
    for (int d = 4; d < 32; d <<= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, d);
      s1 += __shfl_xor_sync(0xffffffffu, s1, d);
    }
    if (g == 0) { sterm[r][nt * 8 + 2 * t] = s0; sterm[r][nt * 8 + 2 * t + 1] = s1; }
  }
  __syncthreads();
  double val = 0.0;
  if (threadIdx.x < A20_TILE) {
    const int q = threadIdx.x;
    const int64_t p = pbase + q;
    const double w = weights[p];
    if (w != 0.0) {
      const double term = (sterm[0][q] + sterm[1][q]) + (sterm[2][q] + sterm[3][q]);
      int sc = 0;
      if (ev.a >= n_rows) sc += scl[(int64_t)(ev.a - n_rows) * Ppad + p];
      if (ev.b >= n_rows) sc += scl[(int64_t)(ev.b - n_rows) * Ppad + p];
      if (ev.c >= n_rows) sc += scl[(int64_t)(ev.c - n_rows) * Ppad + p];
      val = w * (log(term * 0.25) + sc * LOG_MINLH);
    }
  }
  if (threadIdx.x < A20_TILE) {  // warps 0 and 1 hold the 64 site values: fixed-order reduction
#pragma unroll
    for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
    if (lane == 0) wsum[r] = val;
  }
  __syncthreads();
  if (threadIdx.x == 0) partial[(int64_t)ev.tree * blocks_per_eval + pb] = wsum[0] + wsum[1];
}

// Rows of evaluations that share their first operand and need no matrix (TBR pairs: one rooted
// vector of part 1 against every candidate of part 2 already pushed through the reconnecting
// branch).  The shared operand's tile stays in registers while the row's second operands stream
// past it: 644 instead of 1288 bytes per pair and pattern.  Blocks are ordered tile-major, so the
// rows in flight at any time read the same pattern tile of the same candidate set and hit in L2.
// Arithmetic and reduction order are those of eval20_kernel (bit-identical results).
struct LikEvalRow {
  int a, first, count, pad;
};

__global__ void __launch_bounds__(A20_THREADS)
eval20_rows_kernel(const LikEvalRow *__restrict__ rows, int n_ev_rows, const LikEval *__restrict__ evals, int tiles,
                   int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
                   const unsigned int *__restrict__ masks, const double *__restrict__ weights,
                   const ModelDev *__restrict__ model, const double *__restrict__ clv, const int *__restrict__ scl,
                   double *__restrict__ partial) {
  __shared__ double sterm[4][A20_TILE];
  __shared__ double wsum[2];
  const int tile = blockIdx.x / n_ev_rows;
  const LikEvalRow rw = rows[blockIdx.x - tile * n_ev_rows];
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t pbase = (int64_t)tile * A20_TILE;
  double fr[3];
#pragma unroll
  for (int mt = 0; mt < 3; mt++) fr[mt] = (mt * 8 + g) < 20 ? model->freqs[mt * 8 + g] : 0.0;
  double av[A20_TILE / 8][3][2];
#pragma unroll
  for (int nt = 0; nt < A20_TILE / 8; nt++) a20_raw(rw.a, n_rows, r, pbase + nt * 8, lane, Ppad, codes, masks, clv, av[nt]);
  double w = 0.0;
  int sa = 0;
  if (threadIdx.x < A20_TILE) {
    w = weights[pbase + threadIdx.x];
    if (rw.a >= n_rows) sa = scl[(int64_t)(rw.a - n_rows) * Ppad + pbase + threadIdx.x];
  }
  for (int e = rw.first; e < rw.first + rw.count; e++) {
    const int b = evals[e].b, tree = evals[e].tree;
#pragma unroll
    for (int nt = 0; nt < A20_TILE / 8; nt++) {
      double x[3][2];
      a20_raw(b, n_rows, r, pbase + nt * 8, lane, Ppad, codes, masks, clv, x);
      double s0 = 0.0, s1 = 0.0;
#pragma unroll
      for (int mt = 0; mt < 3; mt++) {
        s0 = fma(fr[mt], av[nt][mt][0] * x[mt][0], s0);
        s1 = fma(fr[mt], av[nt][mt][1] * x[mt][1], s1);
      }
#pragma unroll
      for (int d = 4; d < 32; d <<= 1) {
        s0 += __shfl_xor_sync(0xffffffffu, s0, d);
        s1 += __shfl_xor_sync(0xffffffffu, s1, d);
      }
      if (g == 0) { sterm[r][nt * 8 + 2 * t] = s0; sterm[r][nt * 8 + 2 * t + 1] = s1; }
    }
    __syncthreads();
    if (threadIdx.x < A20_TILE) {
      const int q = threadIdx.x;
      double val = 0.0;
      if (w != 0.0) {
        const double term = (sterm[0][q] + sterm[1][q]) + (sterm[2][q] + sterm[3][q]);
        int sc = sa;
        if (b >= n_rows) sc += scl[(int64_t)(b - n_rows) * Ppad + pbase + q];
        val = w * (log(term * 0.25) + sc * LOG_MINLH);
      }
#pragma unroll
      for (int d = 16; d > 0; d >>= 1) val += __shfl_down_sync(0xffffffffu, val, d);
      if (lane == 0) wsum[r] = val;
    }
    __syncthreads();
    if (threadIdx.x == 0) partial[(int64_t)tree * tiles + tile] = wsum[0] + wsum[1];
  }
}

}  // namespace plg
