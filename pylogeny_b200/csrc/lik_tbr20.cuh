// 20 states x Gamma-4: TBR neighbourhoods (BASELINE configs[4]) -- candidate steps and pair grid.
//
// A TBR neighbour = (bisected branch e, reconnection edge in part 1, reconnection edge in part 2).
// Per candidate reconnection edge c = {x, y} of one part (reached from the cut through parent
// candidate p, sibling subtree s at x):
//     N(c) = (P_in N(p)) o T[s->x]          partial of the reduced part at x, looking at y
//     R(c) = (P_half N(c)) o H[y->x]        the part rooted in the middle of {x, y}
//     Q(c) = P_v R(c)                       ... pushed through the reconnecting branch (one part only)
// and a neighbour's likelihood is the pair term  sum_{r,k} pi_k R(c1)[r,k] Q(c2)[r,k].
// T[y->x] = P_full D[y->x] and H[y->x] = P_half D[y->x] are products of the BASE tree's directional
// vectors with their own branch: the same for every bisection, so they are made once per directed
// edge instead of once per candidate (2.3 instead of 4.3 20x20x4 products per candidate), and the
// whole candidate -- three chained products, two elementwise joins -- is ONE op here: the chained
// operand goes from the accumulator (C fragment) to the next product's B fragment through a
// 1.5 KB warp-private shared-memory tile instead of through HBM (5 instead of 8 vector transfers).
//
// Pair grid: the pair terms of one bisection are a (stay candidates) x (pushed candidates) matrix
// product per pattern with contraction length 80 -- dense, so it runs on the FP64 tensor pipe as
// 8 x 8 tiles: pairs20_kernel.  For that the candidates' final vectors are written in a second,
// pattern-major layout: pairvec[slot][pattern][80] (r-major, state-minor), stay side pre-multiplied
// by pi.  The four lanes (g, t = 0..3) of a warp then read candidate g's 640 bytes of a pattern as ten
// 64-byte runs (16 bytes per lane and load), which is both the A fragment (row g, k = t) and, for the
// pushed side, the B fragment (k = t, column g) of 20 consecutive DMMAs -- the contraction index is
// merely enumerated as (step, t) -> element 8 (step / 2) + 2 t + (step & 1).  One warp = one tile x 32
// patterns, site terms accumulated in registers.
#pragma once

namespace plg {

// (Tbr20Op, Tbr20Tile: lik.h)
constexpr int T20_TIMES_PI = 1 << 30;

constexpr int P20_CHUNK = 32;  // patterns per warp item of the pair grid
#ifndef PLG_T20_MINB
#define PLG_T20_MINB 3
#endif

// B fragment of operand `id` for the 8 patterns at p0 (tips: 0/1 indicator columns)
__device__ __forceinline__ void t20_load_b(int id, int n_rows, int r, int64_t p0, int lane, int64_t Ppad,
                                           const unsigned char *__restrict__ codes,
                                           const unsigned int *__restrict__ masks, const double *__restrict__ clv,
                                           double (&b)[5]) {
  const int g = lane >> 2, t = lane & 3;
  if (id < n_rows) {
    const unsigned int m = masks[codes[(int64_t)id * Ppad + p0 + g]];
#pragma unroll
    for (int ks = 0; ks < 5; ks++) b[ks] = (m >> (ks * 4 + t) & 1) ? 1.0 : 0.0;
  } else {
    const double *src = clv + ((int64_t)(id - n_rows) * 80 + r * 20 + t) * Ppad + p0 + g;
#pragma unroll
    for (int ks = 0; ks < 5; ks++) b[ks] = __ldg(src + (int64_t)ks * 4 * Ppad);
  }
}
__device__ __forceinline__ void t20_mma(const double (&P)[3][5], const double (&b)[5], double (&c)[3][2]) {
#pragma unroll
  for (int mt = 0; mt < 3; mt++) c[mt][0] = c[mt][1] = 0.0;
#pragma unroll
  for (int ks = 0; ks < 5; ks++)
#pragma unroll
    for (int mt = 0; mt < 3; mt++) dmma884(c[mt], P[mt][ks], b[ks]);
}

// c = P * c: the accumulator tile (24 x 8, C layout) becomes the B operand through shared memory
__device__ __forceinline__ void t20_chain(double (*tile)[8], const double (&P)[3][5], double (&c)[3][2], int lane) {
  const int g = lane >> 2, t = lane & 3;
  __syncwarp();
#pragma unroll
  for (int mt = 0; mt < 3; mt++) *reinterpret_cast<double2 *>(&tile[mt * 8 + g][2 * t]) = make_double2(c[mt][0], c[mt][1]);
  __syncwarp();
  double b[5];
#pragma unroll
  for (int ks = 0; ks < 5; ks++) b[ks] = tile[ks * 4 + t][g];
  t20_mma(P, b, c);
}

#define t20_colmax a20_colmax
constexpr int T20_HI_MINLH = 0x2FF00000;  // high word of 2^-256
constexpr int T20_HI_256 = 256 << 20;     // what a factor 2^256 adds to a high word

#ifndef PLG_T20_UNROLL
#define PLG_T20_UNROLL 8
#endif
#ifndef PLG_T20_STAGES
#define PLG_T20_STAGES 3
#endif
constexpr int T20_STAGES = PLG_T20_STAGES;
constexpr int T20_UNROLL = PLG_T20_UNROLL;
constexpr int T20_STAGE_DOUBLES = 3 * 20 * 8;  // per warp and stage: operands a, b, b2 x 20 states x 8 patterns
constexpr int T20_DYN_SMEM = 4 * T20_STAGES * T20_STAGE_DOUBLES * 8;

__device__ __forceinline__ void t20_cp16(double *smem_dst, const double *gsrc) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"((unsigned)__cvta_generic_to_shared(smem_dst)), "l"(gsrc));
}

// This warp's (category r) rows of the op's stored operands for the 8 patterns at p0 -> one stage.
// 20 rows x 64 bytes per operand = 80 16-byte pieces over 32 lanes; nothing passes through registers,
// so T20_STAGES - 1 tiles are in flight per warp while one is being multiplied.  The per-lane piece
// offsets (global: row * Ppad + part, shared: row * 64 + part * 16 bytes) are computed ONCE per block:
// in the first version this loop's address arithmetic was a third of the kernel's instructions
// (profiles/r2o_tbr20_step_ncu.txt, source page).
// A fragments of P[pm][r] like a20_load_p, but as loads the compiler may not hoist out of the tile loop
__device__ __forceinline__ void t20_load_p_again(const double *__restrict__ pmat, int pm, int r, int lane, double (&a)[3][5]) {
  const double *P = pmat + ((int64_t)pm * 4 + r) * 401;
  const int g = lane >> 2, t = lane & 3;
#pragma unroll
  for (int mt = 0; mt < 3; mt++) {
    const int row = mt * 8 + g;
#pragma unroll
    for (int ks = 0; ks < 5; ks++) {
      a[mt][ks] = 0.0;
      if (row < 20) asm volatile("ld.global.nc.f64 %0, [%1];\n" : "=d"(a[mt][ks]) : "l"(P + row * 20 + ks * 4 + t));
    }
  }
}

struct T20Pieces {
  int g[3];       // element offsets from (operand base + p0)
  unsigned s[3];  // byte offsets inside an operand's 20 x 8 stage block
  int n;          // pieces this lane copies (3 for lanes < 16, else 2)
};
__device__ __forceinline__ T20Pieces t20_pieces(int lane, int64_t Ppad) {
  T20Pieces pc;
#pragma unroll
  for (int k = 0; k < 3; k++) {
    const int i = lane + 32 * k;
    pc.g[k] = (int)((i >> 2) * Ppad + (i & 3) * 2);
    pc.s[k] = (unsigned)((i >> 2) * 64 + (i & 3) * 16);
  }
  pc.n = lane < 16 ? 3 : 2;
  return pc;
}
__device__ __forceinline__ void t20_stage_issue(unsigned stage_s, const double *const (&base)[3], const T20Pieces &pc, int64_t p0) {
#pragma unroll
  for (int o = 0; o < 3; o++) {
    if (!base[o]) continue;  // absent operand, or a tip row (read from the codes)
    const double *src = base[o] + p0;
    const unsigned dst = stage_s + o * (20 * 64);
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + pc.s[0]), "l"(src + pc.g[0]));
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + pc.s[1]), "l"(src + pc.g[1]));
    if (pc.n == 3) asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst + pc.s[2]), "l"(src + pc.g[2]));
  }
  asm volatile("cp.async.commit_group;\n" ::);
}

__global__ void __launch_bounds__(A20_THREADS, PLG_T20_MINB)
tbr20_step_kernel(const Tbr20Op *__restrict__ ops, int n_ops, int n_rows, int64_t Ppad,
                  const unsigned char *__restrict__ codes, const unsigned int *__restrict__ masks,
                  const ModelDev *__restrict__ model, const double *__restrict__ pmat, double *__restrict__ clv,
                  int *__restrict__ scl, double *__restrict__ pairvec, int *__restrict__ pair_scl) {
  extern __shared__ __align__(16) double t20_smem[];
  __shared__ __align__(16) double tr[4][24][8];
  __shared__ int smx[2][2][4][8];  // [tile parity][X | final result][category][pattern of the tile]: column maxima
  // blocks are ordered pattern-tile-major: the blocks in flight work on a few hundred ops of the SAME
  // 64 patterns, so the T / H vectors of the base tree (shared by all bisections: 16 MB per tile at 50
  // taxa) are read from L2, not from HBM
  const int pb = blockIdx.x / n_ops;
  const Tbr20Op op = ops[blockIdx.x - pb * n_ops];
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t pbase = (int64_t)pb * A20_TILE;
  const int pslot = op.pdst >= 0 ? (op.pdst & ~T20_TIMES_PI) : -1;
  double *stages = t20_smem + r * T20_STAGES * T20_STAGE_DOUBLES;
  const unsigned stages_s = (unsigned)__cvta_generic_to_shared(stages);
  const T20Pieces pc = t20_pieces(lane, Ppad);
  const double *obase[3];  // this warp's category rows of the stored operands (null: absent or a tip)
  {
    const int ids[3] = {op.a, op.b, op.b2};
#pragma unroll
    for (int o = 0; o < 3; o++) obase[o] = ids[o] >= n_rows ? clv + ((int64_t)(ids[o] - n_rows) * 80 + r * 20) * Ppad : nullptr;
  }
#pragma unroll
  for (int s = 0; s < T20_STAGES - 1; s++)
    t20_stage_issue(stages_s + s * (T20_STAGE_DOUBLES * 8), obase, pc, pbase + s * 8);
  double pa[3][5], p2[3][5];  // (the third matrix, pushed candidates only, is fetched per tile: L1 hits, 30 registers less)
  a20_load_p(pmat, op.pa, r, lane, pa);
  if (op.p2 >= 0) a20_load_p(pmat, op.p2, r, lane, p2);
  double fr[3];
#pragma unroll
  for (int mt = 0; mt < 3; mt++)
    fr[mt] = (op.pdst >= 0 && (op.pdst & T20_TIMES_PI) && mt * 8 + g < 20) ? model->freqs[mt * 8 + g] : 1.0;
  double *dst = op.dst >= 0 ? clv + ((int64_t)(op.dst - n_rows) * 80 + r * 20) * Ppad : nullptr;
  double *pd = pslot >= 0 ? pairvec + (int64_t)pslot * Ppad * 80 + r * 20 : nullptr;
  // scaling counters: lane (r = 0, g = 0, t) books patterns 2t, 2t + 1 of every tile
  const bool booker = r == 0 && g == 0;
  const int *sa = op.a >= n_rows ? scl + (int64_t)(op.a - n_rows) * Ppad : nullptr;
  const int *sb = op.b >= n_rows ? scl + (int64_t)(op.b - n_rows) * Ppad : nullptr;
  const int *sb2 = op.b2 >= n_rows ? scl + (int64_t)(op.b2 - n_rows) * Ppad : nullptr;
#pragma unroll(T20_UNROLL)
  for (int nt = 0; nt < A20_TILE / 8; nt++) {
    const int64_t p0 = pbase + nt * 8;
    // tile nt + STAGES - 1 goes into the stage tile nt - 1 was read from (every lane is past it: syncwarp)
    __syncwarp();
    if (nt + T20_STAGES - 1 < A20_TILE / 8)
      t20_stage_issue(stages_s + ((nt + T20_STAGES - 1) % T20_STAGES) * (T20_STAGE_DOUBLES * 8), obase, pc,
                      pbase + (nt + T20_STAGES - 1) * 8);
    else
      asm volatile("cp.async.commit_group;\n" ::);  // (keeps the group count uniform)
    int2 ca = make_int2(0, 0), cb = ca, cb2 = ca;
    if (booker) {
      if (sa) ca = *reinterpret_cast<const int2 *>(sa + p0 + 2 * t);
      if (sb) cb = *reinterpret_cast<const int2 *>(sb + p0 + 2 * t);
      if (sb2) cb2 = *reinterpret_cast<const int2 *>(sb2 + p0 + 2 * t);
    }
    asm volatile("cp.async.wait_group %0;\n" ::"n"(T20_STAGES - 1));
    __syncwarp();
    const double(*sg)[20][8] = reinterpret_cast<const double(*)[20][8]>(stages + (nt % T20_STAGES) * T20_STAGE_DOUBLES);
    double x[3][2], c[3][2], bf[5];
    if (op.a < n_rows) {
      const unsigned int m = masks[codes[(int64_t)op.a * Ppad + p0 + g]];
#pragma unroll
      for (int ks = 0; ks < 5; ks++) bf[ks] = (m >> (ks * 4 + t) & 1) ? 1.0 : 0.0;
    } else {
#pragma unroll
      for (int ks = 0; ks < 5; ks++) bf[ks] = sg[0][ks * 4 + t][g];
    }
    t20_mma(pa, bf, x);
    if (op.b >= 0) {
#pragma unroll
      for (int mt = 0; mt < 3; mt++)
        if (mt * 8 + g < 20) {
          const double2 v = *reinterpret_cast<const double2 *>(&sg[1][mt * 8 + g][2 * t]);
          x[mt][0] *= v.x; x[mt][1] *= v.y;
        }
    }
    int mx0, mx1, mo0, mo1;
    t20_colmax(x, g, mx0, mx1);
    mo0 = mx0; mo1 = mx1;
#pragma unroll
    for (int mt = 0; mt < 3; mt++) { c[mt][0] = x[mt][0]; c[mt][1] = x[mt][1]; }
    if (op.p2 >= 0) {
      t20_chain(tr[r], p2, c, lane);
      if (op.b2 >= 0) {
#pragma unroll
        for (int mt = 0; mt < 3; mt++)
          if (mt * 8 + g < 20) {
            const double2 v = *reinterpret_cast<const double2 *>(&sg[2][mt * 8 + g][2 * t]);
            c[mt][0] *= v.x; c[mt][1] *= v.y;
          }
      }
      if (op.p3 >= 0) {
        double p3[3][5];
        t20_load_p_again(pmat, op.p3, r, lane, p3);
        t20_chain(tr[r], p3, c, lane);
      }
      t20_colmax(c, g, mo0, mo1);
    }
    // the four categories of a pattern are scaled together: exchange the column maxima, then store.
    // X below 2^-256 in all 80 entries: one factor 2^256 (libpll's rule); the later stages are linear in
    // X, so the final result inherits it and takes a second one if it is still below the threshold.
    int(*mx)[4][8] = smx[nt & 1];
    if (g == 0) {
      *reinterpret_cast<int2 *>(&mx[0][r][2 * t]) = make_int2(mx0, mx1);
      *reinterpret_cast<int2 *>(&mx[1][r][2 * t]) = make_int2(mo0, mo1);
    }
    __syncthreads();
    int rx[2], ro[2];
#pragma unroll
    for (int e = 0; e < 2; e++) {
      const int q = 2 * t + e;
      const int hx = max(max(mx[0][0][q], mx[0][1][q]), max(mx[0][2][q], mx[0][3][q]));
      const int ho = max(max(mx[1][0][q], mx[1][1][q]), max(mx[1][2][q], mx[1][3][q]));
      rx[e] = hx < T20_HI_MINLH ? 1 : 0;
      ro[e] = rx[e] + (ho + (rx[e] ? T20_HI_256 : 0) < T20_HI_MINLH ? 1 : 0);
    }
    if (dst) {
      const double f0 = rx[0] ? TWO256 : 1.0, f1 = rx[1] ? TWO256 : 1.0;
#pragma unroll
      for (int mt = 0; mt < 3; mt++)
        if (mt * 8 + g < 20)
          *reinterpret_cast<double2 *>(dst + (int64_t)(mt * 8 + g) * Ppad + p0 + 2 * t) = make_double2(x[mt][0] * f0, x[mt][1] * f1);
    }
    if (pd) {
      const double f0 = ro[0] == 0 ? 1.0 : (ro[0] == 1 ? TWO256 : TWO256 * TWO256);
      const double f1 = ro[1] == 0 ? 1.0 : (ro[1] == 1 ? TWO256 : TWO256 * TWO256);
#pragma unroll
      for (int mt = 0; mt < 3; mt++)
        if (mt * 8 + g < 20) {
          pd[(p0 + 2 * t) * 80 + mt * 8 + g] = c[mt][0] * fr[mt] * f0;
          pd[(p0 + 2 * t + 1) * 80 + mt * 8 + g] = c[mt][1] * fr[mt] * f1;
        }
    }
    if (booker) {
      const int s0 = ca.x + cb.x, s1 = ca.y + cb.y;
      if (op.dst >= 0) *reinterpret_cast<int2 *>(scl + (int64_t)(op.dst - n_rows) * Ppad + p0 + 2 * t) = make_int2(s0 + rx[0], s1 + rx[1]);
      if (pslot >= 0)
        *reinterpret_cast<int2 *>(pair_scl + (int64_t)pslot * Ppad + p0 + 2 * t) = make_int2(s0 + cb2.x + ro[0], s1 + cb2.y + ro[1]);
    }
  }
}

// The plain 20-state update dst = (P_a a) o (P_b b) (LikOp) on the same staged pipeline: the operands' rows
// arrive by cp.async into the warp-private ring while the previous tiles are multiplied, blocks are ordered
// pattern-tile-major.  Replaces clv20_kernel (lik_aa.cuh: operand loads in front of every product) behind
// aa20_updates; both are kept bit-compatible (same products, same scaling rule).
__global__ void __launch_bounds__(A20_THREADS, PLG_T20_MINB)
clv20s_kernel(const LikOp *__restrict__ ops, int n_ops, int n_rows, int64_t Ppad, const unsigned char *__restrict__ codes,
              const unsigned int *__restrict__ masks, const double *__restrict__ pmat, double *__restrict__ clv,
              int *__restrict__ scl) {
  extern __shared__ __align__(16) double t20_smem[];
  __shared__ int smx[2][4][8];  // [tile parity][category][pattern of the tile]: column maxima
  const int pb = blockIdx.x / n_ops;
  const LikOp op = ops[blockIdx.x - pb * n_ops];
  const int lane = threadIdx.x & 31, r = threadIdx.x >> 5, g = lane >> 2, t = lane & 3;
  const int64_t pbase = (int64_t)pb * A20_TILE;
  double *stages = t20_smem + r * T20_STAGES * T20_STAGE_DOUBLES;
  const unsigned stages_s = (unsigned)__cvta_generic_to_shared(stages);
  const T20Pieces pc = t20_pieces(lane, Ppad);
  const double *obase[3];
  obase[0] = op.a >= n_rows ? clv + ((int64_t)(op.a - n_rows) * 80 + r * 20) * Ppad : nullptr;
  obase[1] = op.b >= n_rows ? clv + ((int64_t)(op.b - n_rows) * 80 + r * 20) * Ppad : nullptr;
  obase[2] = nullptr;
#pragma unroll
  for (int s = 0; s < T20_STAGES - 1; s++)
    t20_stage_issue(stages_s + s * (T20_STAGE_DOUBLES * 8), obase, pc, pbase + s * 8);
  double pa[3][5], pbm[3][5];
  a20_load_p(pmat, op.pa, r, lane, pa);
  if (op.b >= 0) a20_load_p(pmat, op.pb, r, lane, pbm);
  double *dst = clv + ((int64_t)(op.dst - n_rows) * 80 + r * 20) * Ppad;
  const bool booker = r == 0 && g == 0;
  const int *sa = op.a >= n_rows ? scl + (int64_t)(op.a - n_rows) * Ppad : nullptr;
  const int *sb = op.b >= n_rows ? scl + (int64_t)(op.b - n_rows) * Ppad : nullptr;
#pragma unroll(T20_UNROLL)
  for (int nt = 0; nt < A20_TILE / 8; nt++) {
    const int64_t p0 = pbase + nt * 8;
    __syncwarp();
    if (nt + T20_STAGES - 1 < A20_TILE / 8)
      t20_stage_issue(stages_s + ((nt + T20_STAGES - 1) % T20_STAGES) * (T20_STAGE_DOUBLES * 8), obase, pc,
                      pbase + (nt + T20_STAGES - 1) * 8);
    else
      asm volatile("cp.async.commit_group;\n" ::);
    int2 ca = make_int2(0, 0), cb = ca;
    if (booker) {
      if (sa) ca = *reinterpret_cast<const int2 *>(sa + p0 + 2 * t);
      if (sb) cb = *reinterpret_cast<const int2 *>(sb + p0 + 2 * t);
    }
    asm volatile("cp.async.wait_group %0;\n" ::"n"(T20_STAGES - 1));
    __syncwarp();
    const double(*sg)[20][8] = reinterpret_cast<const double(*)[20][8]>(stages + (nt % T20_STAGES) * T20_STAGE_DOUBLES);
    double x[3][2], bf[5];
    if (op.a < n_rows) {
      const unsigned int m = masks[codes[(int64_t)op.a * Ppad + p0 + g]];
#pragma unroll
      for (int ks = 0; ks < 5; ks++) bf[ks] = (m >> (ks * 4 + t) & 1) ? 1.0 : 0.0;
    } else {
#pragma unroll
      for (int ks = 0; ks < 5; ks++) bf[ks] = sg[0][ks * 4 + t][g];
    }
    t20_mma(pa, bf, x);
    if (op.b >= 0) {
      double y[3][2];
      if (op.b < n_rows) {
        const unsigned int m = masks[codes[(int64_t)op.b * Ppad + p0 + g]];
#pragma unroll
        for (int ks = 0; ks < 5; ks++) bf[ks] = (m >> (ks * 4 + t) & 1) ? 1.0 : 0.0;
      } else {
#pragma unroll
        for (int ks = 0; ks < 5; ks++) bf[ks] = sg[1][ks * 4 + t][g];
      }
      t20_mma(pbm, bf, y);
#pragma unroll
      for (int mt = 0; mt < 3; mt++) { x[mt][0] *= y[mt][0]; x[mt][1] *= y[mt][1]; }
    }
    int m0, m1;
    t20_colmax(x, g, m0, m1);
    int(*mx)[8] = smx[nt & 1];
    if (g == 0) *reinterpret_cast<int2 *>(&mx[r][2 * t]) = make_int2(m0, m1);
    __syncthreads();
    const int r0 = max(max(mx[0][2 * t], mx[1][2 * t]), max(mx[2][2 * t], mx[3][2 * t])) < T20_HI_MINLH ? 1 : 0;
    const int r1 = max(max(mx[0][2 * t + 1], mx[1][2 * t + 1]), max(mx[2][2 * t + 1], mx[3][2 * t + 1])) < T20_HI_MINLH ? 1 : 0;
    const double f0 = r0 ? TWO256 : 1.0, f1 = r1 ? TWO256 : 1.0;
#pragma unroll
    for (int mt = 0; mt < 3; mt++)
      if (mt * 8 + g < 20)
        *reinterpret_cast<double2 *>(dst + (int64_t)(mt * 8 + g) * Ppad + p0 + 2 * t) = make_double2(x[mt][0] * f0, x[mt][1] * f1);
    if (booker)
      *reinterpret_cast<int2 *>(scl + (int64_t)(op.dst - n_rows) * Ppad + p0 + 2 * t) = make_int2(ca.x + cb.x + r0, ca.y + cb.y + r1);
  }
}

// Pair grid (see the header): warp item = (tile, chunk of 32 patterns).  Items of one bisection are
// adjacent -- chunk-major, tiles inside -- so the vectors of a bisection's candidates are re-read
// from L2 (and, within a block, from L1) by the tiles that share them.
__global__ void __launch_bounds__(128)
pairs20_kernel(const Tbr20Tile *__restrict__ tiles, const int *__restrict__ tile_off, int n_edges, int n_chunks,
               int64_t n_items, int64_t Ppad, const double *__restrict__ weights, const double *__restrict__ pairvec,
               const int *__restrict__ pair_scl, double *__restrict__ partial) {
  const int64_t item = (int64_t)blockIdx.x * 4 + (threadIdx.x >> 5);
  if (item >= n_items) return;
  const int lane = threadIdx.x & 31, g = lane >> 2, t = lane & 3;
  // bisection that holds the item: items of edge e start at tile_off[e] * n_chunks
  int lo = 0, hi = n_edges;
  while (hi - lo > 1) {
    const int mid = (lo + hi) >> 1;
    if ((int64_t)tile_off[mid] * n_chunks <= item) lo = mid; else hi = mid;
  }
  const int nt_e = tile_off[lo + 1] - tile_off[lo];
  const int64_t local = item - (int64_t)tile_off[lo] * n_chunks;
  const int chunk = (int)(local / nt_e);
  const Tbr20Tile &T = tiles[tile_off[lo] + (int)(local - (int64_t)chunk * nt_e)];
  const int64_t p0 = (int64_t)chunk * P20_CHUNK;
  const int sa = T.a[g], sb = T.b[g], sb0 = T.b[2 * t], sb1 = T.b[2 * t + 1];
  const double *pa = pairvec + ((int64_t)sa * Ppad + p0) * 80 + t * 2;
  const double *pb = pairvec + ((int64_t)sb * Ppad + p0) * 80 + t * 2;
  const int *qa = pair_scl + (int64_t)sa * Ppad + p0, *qb0 = pair_scl + (int64_t)sb0 * Ppad + p0,
            *qb1 = pair_scl + (int64_t)sb1 * Ppad + p0;
  double acc0 = 0.0, acc1 = 0.0;
#pragma unroll 1
  for (int pp = 0; pp < P20_CHUNK; pp++) {
    double a[20], b[20];
#pragma unroll
    for (int j = 0; j < 20; j += 2) {
      const double2 u = __ldg(reinterpret_cast<const double2 *>(pa + pp * 80 + j * 4));
      const double2 v = __ldg(reinterpret_cast<const double2 *>(pb + pp * 80 + j * 4));
      a[j] = u.x; a[j + 1] = u.y;
      b[j] = v.x; b[j + 1] = v.y;
    }
    double c[2] = {0.0, 0.0}, d[2] = {0.0, 0.0};
#pragma unroll
    for (int j = 0; j < 20; j += 2) {
      dmma884(c, a[j], b[j]);
      dmma884(d, a[j + 1], b[j + 1]);
    }
    const double w = __ldg(weights + p0 + pp);
    if (w != 0.0) {
      const int s = __ldg(qa + pp);
      acc0 += w * (log((c[0] + d[0]) * 0.25) + (s + __ldg(qb0 + pp)) * LOG_MINLH);
      acc1 += w * (log((c[1] + d[1]) * 0.25) + (s + __ldg(qb1 + pp)) * LOG_MINLH);
    }
  }
  const int t0 = T.tree[g * 8 + 2 * t], t1 = T.tree[g * 8 + 2 * t + 1];
  if (t0 >= 0) partial[(int64_t)t0 * n_chunks + chunk] = acc0;
  if (t1 >= 0) partial[(int64_t)t1 * n_chunks + chunk] = acc1;
}

}  // namespace plg
