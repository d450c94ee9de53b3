// Fitch tree walk: one thread carries one VW-word column (128 patterns for DNA) of all K planes
// through the WHOLE post-order of a tree (treewalk.h).  The running state set lives in registers,
// parked sets in the thread's local-memory stack (L1/L2), so HBM/L2 only supplies the tips:
// the kernel is bound by the integer pipes (about 25 bit-ops per node-op per 32 patterns), not by
// the 3*K/8 bytes per node-op that the level-synchronous kernel has to move (SURVEY.md 8d i).
//
// Semantics per node (binary; fitch.cpp:215-226, 262-264): X = a & b; where X is empty the node
// takes a | b and the pattern costs its weight.  Integer sums only => bit-exact, any order.
#pragma once

namespace plg {

template <int K, int VW>
__device__ __forceinline__ void fw_load_tip(const unsigned int *__restrict__ tips, int row, int64_t vec_words,
                                            int64_t words32, int64_t w0, unsigned int (&x)[K][VW]) {
  const unsigned int *src = tips + (int64_t)row * vec_words + w0;
#pragma unroll
  for (int s = 0; s < K; s++) load_vec<VW>(src + s * words32, x[s]);
}

template <int K, int VW>
__global__ void __launch_bounds__(256)
fitch_walk_kernel(const unsigned int *__restrict__ tips, const WalkOp *__restrict__ ops, int ni, int64_t words32,
                  int blocks_per_tree, const unsigned int *__restrict__ wbits,
                  const unsigned char *__restrict__ chunk_nb, unsigned int *__restrict__ cost) {
  const int tree = blockIdx.x / blocks_per_tree;
  const int cb = blockIdx.x - tree * blocks_per_tree;
  const int64_t col = (int64_t)cb * blockDim.x + threadIdx.x;
  const int64_t w0 = col * VW;
  const int64_t vec_words = (int64_t)K * words32;
  unsigned int total = 0;
  if (w0 < words32) {
    unsigned int stack[WALK_MAX_LEVELS][K][VW];
    unsigned int top[K][VW];
#pragma unroll
    for (int s = 0; s < K; s++)
#pragma unroll
      for (int v = 0; v < VW; v++) top[s][v] = 0;
    const unsigned int nb = chunk_nb[w0 >> 2];
    const int4 *op_p = reinterpret_cast<const int4 *>(ops + (int64_t)tree * ni);
    int4 nxt = __ldg(op_p);
    for (int i = 0; i < ni; i++) {
      const int4 op = nxt;
      if (i + 1 < ni) nxt = __ldg(op_p + i + 1);
      unsigned int xa[K][VW], xb[K][VW];
      if (op.x >= 0) {
        fw_load_tip<K, VW>(tips, op.x, vec_words, words32, w0, xa);
      } else if (op.x == -1) {
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) xa[s][v] = top[s][v];
      } else {
        const int l = -2 - op.x;
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) xa[s][v] = stack[l][s][v];
      }
      if (op.y >= 0) {
        fw_load_tip<K, VW>(tips, op.y, vec_words, words32, w0, xb);
      } else if (op.y == -1) {
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) xb[s][v] = top[s][v];
      } else {
        const int l = -2 - op.y;
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) xb[s][v] = stack[l][s][v];
      }
      unsigned int miss[VW];
#pragma unroll
      for (int v = 0; v < VW; v++) {
        unsigned int any = 0;
#pragma unroll
        for (int s = 0; s < K; s++) any |= xa[s][v] & xb[s][v];
        miss[v] = ~any;
      }
#pragma unroll
      for (int s = 0; s < K; s++)
#pragma unroll
        for (int v = 0; v < VW; v++) xa[s][v] = (xa[s][v] & xb[s][v]) | (miss[v] & (xa[s][v] | xb[s][v]));
      if (op.z >= 0) {
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) stack[op.z][s][v] = xa[s][v];
      } else {
#pragma unroll
        for (int s = 0; s < K; s++)
#pragma unroll
          for (int v = 0; v < VW; v++) top[s][v] = xa[s][v];
      }
      if (nb == 0xFFu) {
#pragma unroll
        for (int v = 0; v < VW; v++) total += __popc(miss[v]);
      } else {
        for (unsigned int b = 0; b < nb; b++) {
          unsigned int wv[VW];
          load_vec<VW>(wbits + (int64_t)b * words32 + w0, wv);
          unsigned int c = 0;
#pragma unroll
          for (int v = 0; v < VW; v++) c += __popc(miss[v] & wv[v]);
          total += c << b;
        }
      }
    }
  }
  total = __reduce_add_sync(0xffffffffu, total);
  __shared__ unsigned int warp_sums[8];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (lane == 0) warp_sums[warp] = total;
  __syncthreads();
  if (warp == 0) {
    unsigned int v = lane < (blockDim.x + 31) / 32 ? warp_sums[lane] : 0;
    v = __reduce_add_sync(0xffffffffu, v);
    if (lane == 0 && v) atomicAdd(cost + tree, v);
  }
}

template <int K, int VW>
static void launch_fitch_walk(const FitchAln &a, const WalkOp *d_ops, int64_t n_trees, int ni, unsigned int *d_cost,
                              cudaStream_t st) {
  const int64_t cols = a.words32 / VW;
  int best_t = 256;
  int64_t best_waste = INT64_MAX;
  for (int t = 256; t >= 64; t -= 32) {
    const int64_t blocks = (cols + t - 1) / t, waste = blocks * t - cols;
    if (waste < best_waste) { best_waste = waste; best_t = t; }
  }
  const int64_t bpt = (cols + best_t - 1) / best_t;
  const int64_t max_trees = std::max<int64_t>(1, ((int64_t)1 << 30) / bpt);
  for (int64_t t0 = 0; t0 < n_trees; t0 += max_trees) {
    const int64_t cnt = std::min(max_trees, n_trees - t0);
    prof_begin(PROF_FITCH_OPS, st);
    fitch_walk_kernel<K, VW><<<(unsigned)(cnt * bpt), best_t, 0, st>>>(a.tips.p, d_ops + t0 * ni, ni, a.words32,
                                                                      (int)bpt, a.wbits.p, a.chunk_nb.p, d_cost + t0);
    count_launch();
    // algorithmic bytes: every node-op moves 3 vectors of K_real planes; requested:
    // the tips (n per tree) -- everything else stays on chip
    const double vec = a.K_real * (double)a.P / 8.0;
    prof_end(PROF_FITCH_OPS, st, (double)cnt * ni * 3.0 * vec, (double)cnt * ni * (double)a.P,
             (double)cnt * (ni + 1) * vec);
  }
}

}  // namespace plg
