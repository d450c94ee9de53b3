// Substitution model on the host: reversible rate matrix eigensystem + discrete Gamma rates.
//
// Stands in for what libpll's pllInitModel sets up at libpllWrapper.c:159
// (GTR / protein eigensystem, Gamma category rates).  Computed once per model on
// the host in double precision and uploaded; the per-branch P-matrices are then
// built on the device (lik.cu: pmatrix_kernel).
#pragma once
#include <vector>

#include "common.h"

namespace plg {

constexpr int LIK_MAX_STATES = 20;
constexpr int LIK_MAX_CATS = 8;

struct ModelDev {  // layout of the device copy
  double eval[LIK_MAX_STATES];
  double U[LIK_MAX_STATES * LIK_MAX_STATES];     // right eigenvectors, row-major [i][k]
  double Uinv[LIK_MAX_STATES * LIK_MAX_STATES];  // row-major [k][j]
  double freqs[LIK_MAX_STATES];
  double rates[LIK_MAX_CATS];
};

struct Model {
  int K = 0, R = 0;
  double alpha = 1.0;
  ModelDev h;
  DevBuf<ModelDev> d;
  Model(int K, const double *exch, const double *freqs, double alpha, int n_cat);
  void pmatrix(double t, double *out) const;  // host P(t), row-major K*K
};

// Yang (1994) discrete Gamma, mean of each of n equal-probability categories.
void gamma_mean_rates(double alpha, int n, double *rates);

}  // namespace plg
