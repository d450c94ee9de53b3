#include "model.h"

#include <cmath>
#include <cstring>

namespace plg {

namespace {

// Regularised lower incomplete gamma P(a, x), Lentz continued fraction above a+1,
// power series below (Abramowitz & Stegun 6.5.29 / 6.5.31).
double reg_lower_gamma(double a, double x) {
  if (x <= 0.0) return 0.0;
  const double lg = std::lgamma(a);
  const double pref = std::exp(a * std::log(x) - x - lg);
  if (x < a + 1.0) {
    double term = 1.0 / a, sum = term, ap = a;
    for (int n = 1; n < 100000; n++) {
      ap += 1.0;
      term *= x / ap;
      sum += term;
      if (std::fabs(term) <= std::fabs(sum) * 1e-17) break;
    }
    return pref * sum;
  }
  const double tiny = 1e-300;
  double b = x + 1.0 - a, c = 1.0 / tiny, d = 1.0 / b, h = d;
  for (int i = 1; i < 100000; i++) {
    const double an = -(double)i * ((double)i - a);
    b += 2.0;
    d = an * d + b;
    if (std::fabs(d) < tiny) d = tiny;
    c = b + an / c;
    if (std::fabs(c) < tiny) c = tiny;
    d = 1.0 / d;
    const double del = d * c;
    h *= del;
    if (std::fabs(del - 1.0) <= 1e-17) break;
  }
  return 1.0 - pref * h;
}

// x such that P(shape, rate*x) = p; monotone bracketed bisection to the last ulp.
double gamma_cut(double p, double shape, double rate) {
  double hi = 1.0;
  while (reg_lower_gamma(shape, rate * hi) < p) hi *= 2.0;
  double lo = 0.0;
  for (int it = 0; it < 2000 && hi - lo > 1e-17 * hi; it++) {
    const double mid = 0.5 * (lo + hi);
    (reg_lower_gamma(shape, rate * mid) < p ? lo : hi) = mid;
  }
  return 0.5 * (lo + hi);
}

}  // namespace

void gamma_mean_rates(double alpha, int n, double *rates) {
  double below = 0.0;
  for (int i = 0; i < n; i++) {
    double upto = 1.0;
    if (i + 1 < n) upto = reg_lower_gamma(alpha + 1.0, alpha * gamma_cut((i + 1.0) / n, alpha, alpha));
    rates[i] = (upto - below) * n;
    below = upto;
  }
}

Model::Model(int K_, const double *exch, const double *freqs, double alpha_, int n_cat) {
  K = K_;
  R = n_cat;
  alpha = alpha_;
  if (K < 2 || K > LIK_MAX_STATES) PLG_FAIL(PLG_ERR_ARG, "n_states must be in 2..%d", LIK_MAX_STATES);
  if (R < 1 || R > LIK_MAX_CATS) PLG_FAIL(PLG_ERR_ARG, "n_cat must be in 1..%d", LIK_MAX_CATS);
  if (!(alpha > 0.0)) PLG_FAIL(PLG_ERR_ARG, "alpha must be positive");
  double fsum = 0;
  for (int i = 0; i < K; i++) {
    if (!(freqs[i] > 0.0)) PLG_FAIL(PLG_ERR_ARG, "state frequencies must be positive");
    fsum += freqs[i];
  }
  memset(&h, 0, sizeof h);
  std::vector<double> pi(K), sq(K);
  for (int i = 0; i < K; i++) {
    pi[i] = freqs[i] / fsum;
    sq[i] = std::sqrt(pi[i]);
    h.freqs[i] = pi[i];
  }
  // symmetric S[i][j] = sqrt(pi_i) s_ij sqrt(pi_j); diagonal from the row sums of Q
  std::vector<double> S((size_t)K * K, 0.0), V((size_t)K * K, 0.0);
  {
    int e = 0;
    for (int i = 0; i < K; i++)
      for (int j = i + 1; j < K; j++, e++) {
        if (!(exch[e] >= 0.0)) PLG_FAIL(PLG_ERR_ARG, "exchangeabilities must be non-negative");
        S[i * K + j] = S[j * K + i] = exch[e] * sq[i] * sq[j];
      }
  }
  double mu = 0.0;  // expected substitutions per unit time before normalisation
  for (int i = 0; i < K; i++) {
    double off = 0.0;  // -Q_ii = sum_j s_ij pi_j
    for (int j = 0; j < K; j++)
      if (j != i) off += S[i * K + j] * sq[j] / sq[i];
    S[i * K + i] = -off;
    mu += pi[i] * off;
  }
  if (!(mu > 0.0)) PLG_FAIL(PLG_ERR_ARG, "rate matrix has no substitutions");
  for (auto &x : S) x /= mu;
  for (int i = 0; i < K; i++) V[i * K + i] = 1.0;
  // cyclic Jacobi: S <- G^T S G, V <- V G until off-diagonal mass vanishes
  for (int sweep = 0; sweep < 100; sweep++) {
    double off = 0.0;
    for (int i = 0; i < K; i++)
      for (int j = i + 1; j < K; j++) off += S[i * K + j] * S[i * K + j];
    if (off < 1e-290) break;
    for (int p = 0; p < K - 1; p++)
      for (int q = p + 1; q < K; q++) {
        const double apq = S[p * K + q];
        if (apq == 0.0) continue;
        const double tau = (S[q * K + q] - S[p * K + p]) / (2.0 * apq);
        const double t = (tau >= 0 ? 1.0 : -1.0) / (std::fabs(tau) + std::sqrt(1.0 + tau * tau));
        const double c = 1.0 / std::sqrt(1.0 + t * t), s = t * c;
        for (int k = 0; k < K; k++) {
          const double x = S[k * K + p], y = S[k * K + q];
          S[k * K + p] = c * x - s * y;
          S[k * K + q] = s * x + c * y;
        }
        for (int k = 0; k < K; k++) {
          const double x = S[p * K + k], y = S[q * K + k];
          S[p * K + k] = c * x - s * y;
          S[q * K + k] = s * x + c * y;
        }
        for (int k = 0; k < K; k++) {
          const double x = V[k * K + p], y = V[k * K + q];
          V[k * K + p] = c * x - s * y;
          V[k * K + q] = s * x + c * y;
        }
      }
  }
  for (int i = 0; i < K; i++) {
    h.eval[i] = S[i * K + i];
    for (int k = 0; k < K; k++) {
      h.U[i * K + k] = V[i * K + k] / sq[i];
      h.Uinv[k * K + i] = V[i * K + k] * sq[i];
    }
  }
  gamma_mean_rates(alpha, R, h.rates);
  d.alloc(1);
  PLG_CUDA(cudaMemcpy(d.p, &h, sizeof h, cudaMemcpyHostToDevice));
}

void Model::pmatrix(double t, double *out) const {
  for (int i = 0; i < K; i++)
    for (int j = 0; j < K; j++) {
      double s = 0.0;
      for (int k = 0; k < K; k++) s += h.U[i * K + k] * std::exp(h.eval[k] * t) * h.Uinv[k * K + j];
      out[i * K + j] = s > 0.0 ? s : 0.0;  // probabilities: clamp round-off negatives (as the device kernel does)
      if (t == 0.0) out[i * K + j] = (i == j) ? 1.0 : 0.0;
    }
}

}  // namespace plg
