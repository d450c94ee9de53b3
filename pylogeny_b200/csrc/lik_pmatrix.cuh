// P-matrix + tip lookup table of ONE branch length, as a block-wide device function: shared by
// pmatrix_kernel (lik.cu) and by the fused Newton-step kernel of the branch-length smoother
// (pll_api.cu), which recomputes the matrices of the branch it has just moved.
#pragma once
#include "model.h"

namespace plg {

template <int K>
struct NCodes { static constexpr int value = K == 4 ? 16 : 23; };

// all threads of the block call it; pmat_dst / tiptab_dst point at this branch's entries
template <int K, int R>
__device__ __forceinline__ void pmatrix_block(const ModelDev *__restrict__ model, double t,
                                              const unsigned int *__restrict__ masks, double *__restrict__ pmat_dst,
                                              double *__restrict__ tiptab_dst) {
  constexpr int KK = K * K, KKP = KK + 1, NC = NCodes<K>::value;
  __shared__ double ex[R * K];
  __shared__ double Pm[R * KKP];
  for (int e = threadIdx.x; e < R * K; e += blockDim.x)
    ex[e] = exp(model->eval[e % K] * t * model->rates[e / K]);
  __syncthreads();
  for (int e = threadIdx.x; e < R * KK; e += blockDim.x) {
    const int r = e / KK, i = (e / K) % K, j = e % K;
    double s = 0.0;
#pragma unroll 4
    for (int k = 0; k < K; k++) s += model->U[i * K + k] * ex[r * K + k] * model->Uinv[k * K + j];
    s = fmax(s, 0.0);  // transition probabilities: round-off can leave -1e-17 at t ~ 0
    if (t == 0.0) s = (i == j) ? 1.0 : 0.0;  // a zero-length branch is the identity exactly
    Pm[r * KKP + i * K + j] = s;
    pmat_dst[r * KKP + i * K + j] = s;
  }
  __syncthreads();
  for (int e = threadIdx.x; e < NC * R * K; e += blockDim.x) {
    const int code = e / (R * K), r = (e / K) % R, k = e % K;
    const unsigned int m = masks[code];
    double s = 0.0;
    for (int j = 0; j < K; j++)
      if (m >> j & 1) s += Pm[r * KKP + k * K + j];
    tiptab_dst[e] = s;
  }
}

}  // namespace plg
