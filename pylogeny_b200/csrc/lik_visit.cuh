// DNA x Gamma-4 "visit" kernel: one search step of the SPR edge-insertion walk scores BOTH
// neighbours that hang off the node it arrives at, with the four operand CLVs staged ONCE in
// shared memory by the TMA (cp.async.bulk + mbarrier) instead of once per neighbour.
//
// Arriving at node x from the pruned side with the partial X (reduced tree seen through the
// incoming edge), x's two other neighbours c1, c2 give two insertion edges:
//     N1 = (P_in X) o (P_t2 D2)        lnL1 = sum_sites log sum pi (P_h1 N1) o (P_h1 D1) o (P_v Tv)
//     N2 = (P_in X) o (P_t1 D1)        lnL2 = sum_sites log sum pi (P_h2 N2) o (P_h2 D2) o (P_v Tv)
// with D_i = D[c_i -> x], t_i the length of edge {x, c_i}, h_i = t_i / 2, Tv the pruned subtree.
// Shared operands (X, D1, D2, Tv) are read once for two neighbours and N_i is only written when
// the walk continues below c_i: 4 reads + <= 2 writes per two neighbours instead of 2 x (4 + 1).
//
// Each block owns 128 patterns.  Inner operands land in shared memory as 16 rows of 1 KB
// (plane-major, exactly the global layout) through 16 bulk copies per operand signalled on one
// mbarrier; tip operands use that space for their lookup tables instead.  64 KB + 3 KB of
// matrices per block => 3 resident blocks per SM keep ~190 KB of loads in flight per SM while
// the others compute, without holding operands in registers.
#pragma once

namespace plg {

constexpr int V4_THREADS = 128;
constexpr int V4_LAND = 16 * V4_THREADS;                  // doubles per operand landing zone
constexpr int V4_SMEM_BYTES = (4 * V4_LAND + 6 * 64) * 8 + 128;

__device__ __forceinline__ uint32_t v4_smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void v4_stage_mat(double *sm, int pm, const double *__restrict__ pmat) {
  const double *src = pmat + (int64_t)pm * 68;
  for (int e = threadIdx.x; e < 64; e += V4_THREADS) sm[e] = src[(e >> 4) * 17 + (e & 15)];
}
__device__ __forceinline__ void v4_stage_tab(double *sm, int pm, const double *__restrict__ tiptab) {
  const double *src = tiptab + (int64_t)pm * 256;
  for (int e = threadIdx.x; e < 256; e += V4_THREADS) sm[e] = src[e];
}

// v (*)= P x with x read from the landing zone (column threadIdx.x)
template <bool MUL>
__device__ __forceinline__ void v4_apply(const double *mat, const double *land, double (&v)[16]) {
#pragma unroll
  for (int r = 0; r < 4; r++) {
    double x[4];
#pragma unroll
    for (int j = 0; j < 4; j++) x[j] = land[(r * 4 + j) * V4_THREADS + threadIdx.x];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double2 p01 = *reinterpret_cast<const double2 *>(mat + r * 16 + k * 4);
      const double2 p23 = *reinterpret_cast<const double2 *>(mat + r * 16 + k * 4 + 2);
      double s = p01.x * x[0];
      s = fma(p01.y, x[1], s);
      s = fma(p23.x, x[2], s);
      s = fma(p23.y, x[3], s);
      if (MUL) v[r * 4 + k] *= s; else v[r * 4 + k] = s;
    }
  }
}

// same with x in registers
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

template <bool MUL>
__device__ __forceinline__ void v4_tip(const double *tab, int code, double (&v)[16]) {
#pragma unroll
  for (int i = 0; i < 16; i += 2) {
    const double2 t = *reinterpret_cast<const double2 *>(tab + code * 16 + i);
    if (MUL) { v[i] *= t.x; v[i + 1] *= t.y; } else { v[i] = t.x; v[i + 1] = t.y; }
  }
}

__device__ __forceinline__ double v4_site(const double (&e)[16], int sc, double w, const ModelDev *__restrict__ model) {
  double term = 0.0;
#pragma unroll
  for (int k = 0; k < 4; k++) term = fma(model->freqs[k], (e[k] + e[4 + k]) + (e[8 + k] + e[12 + k]), term);
  return w != 0.0 ? w * (log(term * 0.25) + sc * LOG_MINLH) : 0.0;
}

__global__ void __launch_bounds__(V4_THREADS, 3)
visit4_kernel(const LikVisit *__restrict__ visits, int blocks_per_op, int n_rows, int64_t Ppad,
              const unsigned char *__restrict__ codes, const double *__restrict__ weights,
              const ModelDev *__restrict__ model, const double *__restrict__ pmat,
              const double *__restrict__ tiptab, double *__restrict__ clv, int *__restrict__ scl,
              double *__restrict__ partial, int pstride) {
  extern __shared__ __align__(128) unsigned char v4_smem[];
  double *land = reinterpret_cast<double *>(v4_smem);                 // [4][16][128]
  double *mats = land + 4 * V4_LAND;                                  // [6][64]
  uint64_t *mbar = reinterpret_cast<uint64_t *>(mats + 6 * 64);
  double *wsum = reinterpret_cast<double *>(mbar + 1);                // [2][4] (inside the 128-byte tail)
  const int v_idx = blockIdx.x / blocks_per_op;
  const int pb = blockIdx.x - v_idx * blocks_per_op;
  const LikVisit vi = visits[v_idx];
  const int64_t p = (int64_t)pb * V4_THREADS + threadIdx.x;
  const int ids[4] = {vi.x_in, vi.tv, vi.d1, vi.d2};
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;\n" ::"r"(v4_smem_u32(mbar)));
    asm volatile("fence.mbarrier_init.release.cluster;\n" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x < 32) {  // warp 0: arm the barrier, then 16 x 1 KB bulk copies per inner operand
    int n_inner = 0;
#pragma unroll
    for (int o = 0; o < 4; o++) n_inner += ids[o] >= n_rows;
    if (threadIdx.x == 0 && n_inner)
      asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;\n" ::"r"(v4_smem_u32(mbar)),
                   "r"(n_inner * V4_LAND * 8)
                   : "memory");
    __syncwarp();
    if (threadIdx.x < 16) {
#pragma unroll
      for (int o = 0; o < 4; o++)
        if (ids[o] >= n_rows) {
          const double *src = clv + ((int64_t)(ids[o] - n_rows) * 16 + threadIdx.x) * Ppad + (int64_t)pb * V4_THREADS;
          asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];\n" ::"r"(
                           v4_smem_u32(land + o * V4_LAND + threadIdx.x * V4_THREADS)),
                       "l"(src), "r"(V4_THREADS * 8), "r"(v4_smem_u32(mbar))
                       : "memory");
        }
    }
  }
  // matrices (and, for tip operands, their lookup tables in the unused landing zone) by plain loads
  v4_stage_mat(mats + 0 * 64, vi.p_in, pmat);
  v4_stage_mat(mats + 1 * 64, vi.p_v, pmat);
  v4_stage_mat(mats + 2 * 64, vi.p_t1, pmat);
  v4_stage_mat(mats + 3 * 64, vi.p_h1, pmat);
  v4_stage_mat(mats + 4 * 64, vi.p_t2, pmat);
  v4_stage_mat(mats + 5 * 64, vi.p_h2, pmat);
  if (vi.x_in < n_rows) v4_stage_tab(land + 0 * V4_LAND, vi.p_in, tiptab);
  if (vi.tv < n_rows) v4_stage_tab(land + 1 * V4_LAND, vi.p_v, tiptab);
  if (vi.d1 < n_rows) { v4_stage_tab(land + 2 * V4_LAND, vi.p_t1, tiptab); v4_stage_tab(land + 2 * V4_LAND + 256, vi.p_h1, tiptab); }
  if (vi.d2 < n_rows) { v4_stage_tab(land + 3 * V4_LAND, vi.p_t2, tiptab); v4_stage_tab(land + 3 * V4_LAND + 256, vi.p_h2, tiptab); }
  int code[4] = {0, 0, 0, 0}, sc[4] = {0, 0, 0, 0};
#pragma unroll
  for (int o = 0; o < 4; o++) {
    if (ids[o] < n_rows) code[o] = codes[(int64_t)ids[o] * Ppad + p];
    else sc[o] = scl[(int64_t)(ids[o] - n_rows) * Ppad + p];
  }
  const double w = weights[p];
  __syncthreads();
  if (ids[0] >= n_rows || ids[1] >= n_rows || ids[2] >= n_rows || ids[3] >= n_rows) {
    uint32_t done = 0;
    while (!done)
      asm volatile("{\n.reg .pred P1;\nmbarrier.try_wait.parity.shared::cta.b64 P1, [%1], 0;\nselp.b32 %0, 1, 0, P1;\n}\n"
                   : "=r"(done)
                   : "r"(v4_smem_u32(mbar))
                   : "memory");
  }
  double a[16], tvv[16];
  if (vi.x_in < n_rows) v4_tip<false>(land + 0 * V4_LAND, code[0], a);
  else v4_apply<false>(mats + 0 * 64, land + 0 * V4_LAND, a);
  if (vi.tv < n_rows) v4_tip<false>(land + 1 * V4_LAND, code[1], tvv);
  else v4_apply<false>(mats + 1 * 64, land + 1 * V4_LAND, tvv);
  double v1 = 0.0, v2 = 0.0;
#pragma unroll
  for (int mv = 0; mv < 2; mv++) {
    // neighbour `mv`: partial from the OTHER child's CLV over its full branch, far side = own child over half
    const int oth = mv == 0 ? 3 : 2, own = mv == 0 ? 2 : 3;
    const int dst = mv == 0 ? vi.n1 : vi.n2, tree = mv == 0 ? vi.tree1 : vi.tree2;
    if (dst < 0 && tree < 0) continue;
    const double *m_full_oth = mats + (oth == 3 ? 4 : 2) * 64;   // P_t of the other child
    const double *m_half_own = mats + (own == 2 ? 3 : 5) * 64;   // P_h of the own child
    double n[16];
#pragma unroll
    for (int i = 0; i < 16; i++) n[i] = a[i];
    if (ids[oth] < n_rows) v4_tip<true>(land + oth * V4_LAND, code[oth], n);
    else v4_apply<true>(m_full_oth, land + oth * V4_LAND, n);
    int scn = sc[0] + sc[oth];
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) mx = fmax(mx, fabs(n[i]));
    if (mx < MINLH) {
#pragma unroll
      for (int i = 0; i < 16; i++) n[i] *= TWO256;
      scn += 1;
    }
    if (dst >= 0) {
      double *out = clv + (int64_t)(dst - n_rows) * 16 * Ppad + p;
#pragma unroll
      for (int i = 0; i < 16; i++) out[(int64_t)i * Ppad] = n[i];
      scl[(int64_t)(dst - n_rows) * Ppad + p] = scn;
    }
    if (tree >= 0) {
      double e[16];
      v4_apply_reg<false>(m_half_own, n, e);
      if (ids[own] < n_rows) v4_tip<true>(land + own * V4_LAND + 256, code[own], e);
      else v4_apply<true>(m_half_own, land + own * V4_LAND, e);
#pragma unroll
      for (int i = 0; i < 16; i++) e[i] *= tvv[i];
      const double val = v4_site(e, scn + sc[own] + sc[1], w, model);
      if (mv == 0) v1 = val; else v2 = val;
    }
  }
  // two fixed-order block reductions
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v1 += __shfl_down_sync(0xffffffffu, v1, d);
    v2 += __shfl_down_sync(0xffffffffu, v2, d);
  }
  if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = v1; wsum[4 + (threadIdx.x >> 5)] = v2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (vi.tree1 >= 0) partial[(int64_t)vi.tree1 * pstride + pb] = (wsum[0] + wsum[1]) + (wsum[2] + wsum[3]);
    if (vi.tree2 >= 0) partial[(int64_t)vi.tree2 * pstride + pb] = (wsum[4] + wsum[5]) + (wsum[6] + wsum[7]);
  }
}

}  // namespace plg

// ---------------------------------------------------------------------------------------------
// visit4b_kernel: the same pair-fused search step without the 64 KB landing zone.  Thread = one
// pattern; the two vectors shared by both neighbours -- a = P_in X and tvv = P_v Tv -- are parked in
// shared memory by the thread that computed them (conflict-free [entry][thread] layout, no barrier
// needed), so the register file only ever holds the working set of ONE neighbour and 5-6 blocks stay
// resident per SM.  D1 / D2 are read from global twice by the same thread (second read hits L1);
// tip tables are read through the read-only path instead of being staged.
namespace plg {

constexpr int VB_THREADS = 128;
#ifndef PLG_VB_MINB
#define PLG_VB_MINB 5
#endif

template <bool MUL>
__device__ __forceinline__ void vb_operand(const double *mat, int id, int pm_tab, int n_rows, int code, int64_t p,
                                           int64_t Ppad, const double *__restrict__ tiptab,
                                           const double *__restrict__ clv, double (&v)[16]) {
  if (id < n_rows) {
    const double *row = tiptab + (int64_t)pm_tab * 256 + code * 16;
#pragma unroll
    for (int i = 0; i < 16; i += 2) {
      const double2 t = __ldg(reinterpret_cast<const double2 *>(row + i));
      if (MUL) { v[i] *= t.x; v[i + 1] *= t.y; } else { v[i] = t.x; v[i + 1] = t.y; }
    }
  } else {
    double x[16];
    d4_load16(clv + (int64_t)(id - n_rows) * 16 * Ppad + p, Ppad, x);
    v4_apply_reg<MUL>(mat, x, v);
  }
}

__global__ void __launch_bounds__(VB_THREADS, PLG_VB_MINB)
visit4b_kernel(const LikVisit *__restrict__ visits, int blocks_per_op, int n_rows, int64_t Ppad,
               const unsigned char *__restrict__ codes, const double *__restrict__ weights,
               const ModelDev *__restrict__ model, const double *__restrict__ pmat,
               const double *__restrict__ tiptab, double *__restrict__ clv, int *__restrict__ scl,
               double *__restrict__ partial, int pstride) {
  __shared__ __align__(16) double mats[6 * 64];
  __shared__ double park[2][16][VB_THREADS];  // a, tvv
  __shared__ double wsum[8];
  const int v_idx = blockIdx.x / blocks_per_op;
  const int pb = blockIdx.x - v_idx * blocks_per_op;
  const LikVisit vi = visits[v_idx];
  const int64_t p = (int64_t)pb * VB_THREADS + threadIdx.x;
  v4_stage_mat(mats + 0 * 64, vi.p_in, pmat);
  v4_stage_mat(mats + 1 * 64, vi.p_v, pmat);
  v4_stage_mat(mats + 2 * 64, vi.p_t1, pmat);
  v4_stage_mat(mats + 3 * 64, vi.p_h1, pmat);
  v4_stage_mat(mats + 4 * 64, vi.p_t2, pmat);
  v4_stage_mat(mats + 5 * 64, vi.p_h2, pmat);
  const int cx = vi.x_in < n_rows ? codes[(int64_t)vi.x_in * Ppad + p] : 0;
  const int ct = vi.tv < n_rows ? codes[(int64_t)vi.tv * Ppad + p] : 0;
  const int c1 = vi.d1 < n_rows ? codes[(int64_t)vi.d1 * Ppad + p] : 0;
  const int c2 = vi.d2 < n_rows ? codes[(int64_t)vi.d2 * Ppad + p] : 0;
  const int sx = vi.x_in < n_rows ? 0 : scl[(int64_t)(vi.x_in - n_rows) * Ppad + p];
  const int st = vi.tv < n_rows ? 0 : scl[(int64_t)(vi.tv - n_rows) * Ppad + p];
  const int s1 = vi.d1 < n_rows ? 0 : scl[(int64_t)(vi.d1 - n_rows) * Ppad + p];
  const int s2 = vi.d2 < n_rows ? 0 : scl[(int64_t)(vi.d2 - n_rows) * Ppad + p];
  const double w = weights[p];
  __syncthreads();
  {
    double t[16];
    vb_operand<false>(mats + 0 * 64, vi.x_in, vi.p_in, n_rows, cx, p, Ppad, tiptab, clv, t);
#pragma unroll
    for (int i = 0; i < 16; i++) park[0][i][threadIdx.x] = t[i];
    vb_operand<false>(mats + 1 * 64, vi.tv, vi.p_v, n_rows, ct, p, Ppad, tiptab, clv, t);
#pragma unroll
    for (int i = 0; i < 16; i++) park[1][i][threadIdx.x] = t[i];
  }
  double v1 = 0.0, v2 = 0.0;
#pragma unroll
  for (int mv = 0; mv < 2; mv++) {
    const int dst = mv == 0 ? vi.n1 : vi.n2, tree = mv == 0 ? vi.tree1 : vi.tree2;
    if (dst < 0 && tree < 0) continue;
    // neighbour mv: partial over the OTHER child's full branch, far side = own child over half its branch
    const int oth_id = mv == 0 ? vi.d2 : vi.d1, own_id = mv == 0 ? vi.d1 : vi.d2;
    const int oth_pm = mv == 0 ? vi.p_t2 : vi.p_t1, own_pm = mv == 0 ? vi.p_h1 : vi.p_h2;
    const double *m_oth = mats + (mv == 0 ? 4 : 2) * 64, *m_own = mats + (mv == 0 ? 3 : 5) * 64;
    const int oth_code = mv == 0 ? c2 : c1, own_code = mv == 0 ? c1 : c2;
    double n[16];
#pragma unroll
    for (int i = 0; i < 16; i++) n[i] = park[0][i][threadIdx.x];
    vb_operand<true>(m_oth, oth_id, oth_pm, n_rows, oth_code, p, Ppad, tiptab, clv, n);
    int scn = sx + (mv == 0 ? s2 : s1);
    double mx = 0.0;
#pragma unroll
    for (int i = 0; i < 16; i++) mx = fmax(mx, fabs(n[i]));
    if (mx < MINLH) {
#pragma unroll
      for (int i = 0; i < 16; i++) n[i] *= TWO256;
      scn += 1;
    }
    if (dst >= 0) {
      double *out = clv + (int64_t)(dst - n_rows) * 16 * Ppad + p;
#pragma unroll
      for (int i = 0; i < 16; i++) out[(int64_t)i * Ppad] = n[i];
      scl[(int64_t)(dst - n_rows) * Ppad + p] = scn;
    }
    if (tree >= 0) {
      double e[16];
      v4_apply_reg<false>(m_own, n, e);
      vb_operand<true>(m_own, own_id, own_pm, n_rows, own_code, p, Ppad, tiptab, clv, e);
#pragma unroll
      for (int i = 0; i < 16; i++) e[i] *= park[1][i][threadIdx.x];
      const double val = v4_site(e, scn + (mv == 0 ? s1 : s2) + st, w, model);
      if (mv == 0) v1 = val; else v2 = val;
    }
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) {
    v1 += __shfl_down_sync(0xffffffffu, v1, d);
    v2 += __shfl_down_sync(0xffffffffu, v2, d);
  }
  if ((threadIdx.x & 31) == 0) { wsum[threadIdx.x >> 5] = v1; wsum[4 + (threadIdx.x >> 5)] = v2; }
  __syncthreads();
  if (threadIdx.x == 0) {
    if (vi.tree1 >= 0) partial[(int64_t)vi.tree1 * pstride + pb] = (wsum[0] + wsum[1]) + (wsum[2] + wsum[3]);
    if (vi.tree2 >= 0) partial[(int64_t)vi.tree2 * pstride + pb] = (wsum[4] + wsum[5]) + (wsum[6] + wsum[7]);
  }
}

}  // namespace plg
