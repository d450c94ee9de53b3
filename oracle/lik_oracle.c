/* TEST INFRASTRUCTURE ONLY -- not product code.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() may load this.
 *
 * PARITY UNPINNED.  The reference's likelihood arithmetic lives in libpll 1.0.2
 * (prebuilt `pll-sse3`, fetched by /root/reference/install-pll.sh:2-4, linked at
 * /root/reference/setup.py:26) which is NOT under /root/reference, not installed
 * and not fetchable (no network); no reference test pins a log-likelihood value
 * (SURVEY.md 8c).  This file restates the PUBLISHED algorithms libpll implements
 * for the calls made at /root/reference/pylogeny/libpllWrapper.c:127-166
 * (PLL_GAMMA rate heterogeneity, per-site scaling, GTR / protein partitions) and
 * :205-209 (likelihood evaluation):
 *   - reversible-Q eigendecomposition, P(t) = U exp(L t) U^-1
 *   - Yang (1994) discrete Gamma, `ncat` equiprobable categories, MEAN rates
 *   - Felsenstein (1981) pruning, conditional likelihood vectors per node
 *   - per-site power-of-two rescaling: all R*K entries < 2^-256 -> * 2^256,
 *     counter++ ; lnL += counter * log(2^-256)
 * scalar double precision, obviously-correct loops.  What pins it instead:
 * tests/test_oracle_cpu.py checks Gamma rates against scipy.stats.gamma,
 * P(t) against scipy.linalg.expm, and the pruning lnL against a brute-force
 * sum over all internal-state assignments on tiny trees.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define OR_MAXK 20
#define OR_TWO256 1.15792089237316195423570985008687907853269984665640564039457584007913129639936e77
#define OR_MINLH (1.0 / OR_TWO256)

/* ---- regularised lower incomplete gamma P(a,x): series / continued fraction ---- */
static double inc_gamma_p(double a, double x) {
  if (x <= 0) return 0.0;
  double gln = lgamma(a);
  if (x < a + 1.0) {
    double ap = a, sum = 1.0 / a, del = sum;
    for (int n = 0; n < 10000; n++) {
      ap += 1.0; del *= x / ap; sum += del;
      if (fabs(del) < fabs(sum) * 1e-17) break;
    }
    return sum * exp(-x + a * log(x) - gln);
  } else {
    double b = x + 1.0 - a, c = 1.0 / 1e-300, d = 1.0 / b, h = d;
    for (int i = 1; i < 10000; i++) {
      double an = -i * (i - a);
      b += 2.0;
      d = an * d + b; if (fabs(d) < 1e-300) d = 1e-300;
      c = b + an / c; if (fabs(c) < 1e-300) c = 1e-300;
      d = 1.0 / d;
      double del = d * c;
      h *= del;
      if (fabs(del - 1.0) < 1e-17) break;
    }
    return 1.0 - exp(-x + a * log(x) - gln) * h;
  }
}

/* quantile of Gamma(shape a, rate b) by bisection on P(a, b*x) */
static double gamma_quantile(double p, double a, double b) {
  double lo = 0.0, hi = 1.0;
  while (inc_gamma_p(a, hi * b) < p) hi *= 2.0;
  for (int it = 0; it < 400; it++) {
    double mid = 0.5 * (lo + hi);
    if (inc_gamma_p(a, mid * b) < p) lo = mid; else hi = mid;
    if (hi - lo <= 1e-17 * hi) break;
  }
  return 0.5 * (lo + hi);
}

/* Yang 1994 eq. 10: mean of each of ncat equal-probability slices of
 * Gamma(alpha, alpha): r_i = ncat * [P(alpha+1, b_i*alpha) - P(alpha+1, b_{i-1}*alpha)] */
void oracle_gamma_rates(double alpha, int ncat, double *rates) {
  double prev = 0.0;
  for (int i = 0; i < ncat; i++) {
    double cur = 1.0;
    if (i < ncat - 1) {
      double cut = gamma_quantile((double)(i + 1) / ncat, alpha, alpha);
      cur = inc_gamma_p(alpha + 1.0, cut * alpha);
    }
    rates[i] = (cur - prev) * ncat;
    prev = cur;
  }
}

/* ---- eigensystem of a reversible rate matrix ---- */
/* exch: K(K-1)/2 exchangeabilities, upper triangle row-major; freqs: K.
 * Q_ij = s_ij * pi_j, rows sum to 0, scaled so that -sum_i pi_i Q_ii = 1.
 * B = D^1/2 Q D^-1/2 is symmetric; cyclic Jacobi on B. Outputs: eval[K],
 * U[K][K] (right eigenvectors as columns), Uinv[K][K]. */
void oracle_eigen(int K, const double *exch, const double *freqs, double *eval,
                  double *U, double *Uinv) {
  double Q[OR_MAXK][OR_MAXK], B[OR_MAXK][OR_MAXK], V[OR_MAXK][OR_MAXK];
  int idx = 0;
  for (int i = 0; i < K; i++) for (int j = i + 1; j < K; j++) {
    Q[i][j] = exch[idx] * freqs[j];
    Q[j][i] = exch[idx] * freqs[i];
    idx++;
  }
  double mu = 0;
  for (int i = 0; i < K; i++) {
    double s = 0;
    for (int j = 0; j < K; j++) if (j != i) s += Q[i][j];
    Q[i][i] = -s;
    mu += freqs[i] * s;
  }
  for (int i = 0; i < K; i++) for (int j = 0; j < K; j++) {
    B[i][j] = Q[i][j] / mu * sqrt(freqs[i]) / sqrt(freqs[j]);
    V[i][j] = (i == j);
  }
  for (int i = 0; i < K; i++) for (int j = i + 1; j < K; j++) { /* enforce exact symmetry */
    double m = 0.5 * (B[i][j] + B[j][i]); B[i][j] = B[j][i] = m;
  }
  for (int sweep = 0; sweep < 200; sweep++) {
    double off = 0;
    for (int i = 0; i < K; i++) for (int j = i + 1; j < K; j++) off += B[i][j] * B[i][j];
    if (off < 1e-300) break;
    for (int p = 0; p < K; p++) for (int q = p + 1; q < K; q++) {
      if (fabs(B[p][q]) < 1e-300) continue;
      double theta = (B[q][q] - B[p][p]) / (2.0 * B[p][q]);
      double t = (theta >= 0 ? 1.0 : -1.0) / (fabs(theta) + sqrt(theta * theta + 1.0));
      double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
      for (int k = 0; k < K; k++) { /* columns */
        double bkp = B[k][p], bkq = B[k][q];
        B[k][p] = c * bkp - s * bkq; B[k][q] = s * bkp + c * bkq;
      }
      for (int k = 0; k < K; k++) { /* rows */
        double bpk = B[p][k], bqk = B[q][k];
        B[p][k] = c * bpk - s * bqk; B[q][k] = s * bpk + c * bqk;
      }
      for (int k = 0; k < K; k++) {
        double vkp = V[k][p], vkq = V[k][q];
        V[k][p] = c * vkp - s * vkq; V[k][q] = s * vkp + c * vkq;
      }
    }
  }
  for (int i = 0; i < K; i++) {
    eval[i] = B[i][i];
    for (int j = 0; j < K; j++) {
      U[i * K + j] = V[i][j] / sqrt(freqs[i]);
      Uinv[j * K + i] = V[i][j] * sqrt(freqs[i]);
    }
  }
}

/* P(t) = U diag(exp(eval * t)) Uinv, row-major K x K */
void oracle_pmatrix(int K, const double *eval, const double *U, const double *Uinv,
                    double t, double *P) {
  for (int i = 0; i < K; i++) for (int j = 0; j < K; j++) {
    double s = 0;
    for (int k = 0; k < K; k++) s += U[i * K + k] * exp(eval[k] * t) * Uinv[k * K + j];
    P[i * K + j] = s > 0.0 ? s : 0.0; /* a probability: round-off can leave -1e-17 at t ~ 0 */
    if (t == 0.0) P[i * K + j] = (i == j) ? 1.0 : 0.0; /* zero-length branch: identity exactly */
  }
}

/* A model's state frequencies are a probability distribution: tables that sum to 1 only within their
 * printed digits (PAML's wag.dat: 1 + 3e-6) are normalised before use, as the engine does
 * (pylogeny_b200/csrc/model.cpp Model::Model).  Without it lnL shifts by P * log(sum pi). */
static const double *oracle_unit_freqs(int K, const double *freqs, double *buf) {
  double s = 0;
  for (int i = 0; i < K; i++) s += freqs[i];
  for (int i = 0; i < K; i++) buf[i] = freqs[i] / s;
  return buf;
}

/* ---- Felsenstein pruning over a post-order descriptor ----
 * Tree: n_tips tips (ids 0..n_tips-1 = rows of `codes`), n_inner internal nodes
 * with ids n_tips+i in post-order; desc[2i],desc[2i+1] = children of internal i,
 * brlen[2i],brlen[2i+1] = the two child branch lengths.  Root = last internal
 * node.  codes[row*P + p] indexes `masks` (bit k set = state k compatible):
 * DNA: 16 masks (code == mask); AA: 23 codes (20 states, B, Z, X).
 * Returns lnL = sum_p w_p * ( log( (1/R) sum_r sum_k pi_k clv_root[p][r][k] )
 *                             + scale_p * log(2^-256) ).                      */
double oracle_lnl(int K, int n_tips, int64_t P, const uint8_t *codes,
                  const uint32_t *masks, const double *weights,
                  const double *exch, const double *freqs, double alpha, int ncat,
                  int n_inner, const int32_t *desc, const double *brlen) {
  double eval[OR_MAXK], U[OR_MAXK * OR_MAXK], Uinv[OR_MAXK * OR_MAXK], rates[64], unit[OR_MAXK];
  freqs = oracle_unit_freqs(K, freqs, unit);
  oracle_eigen(K, exch, freqs, eval, U, Uinv);
  oracle_gamma_rates(alpha, ncat, rates);
  const int RK = ncat * K;
  double *clv = (double *)malloc(sizeof(double) * (size_t)n_inner * RK);
  int *scl = (int *)malloc(sizeof(int) * (size_t)n_inner);
  double *pm = (double *)malloc(sizeof(double) * (size_t)n_inner * 2 * ncat * K * K);
  for (int i = 0; i < n_inner; i++) for (int c = 0; c < 2; c++) for (int r = 0; r < ncat; r++)
    oracle_pmatrix(K, eval, U, Uinv, brlen[2 * i + c] * rates[r],
                   pm + (((size_t)i * 2 + c) * ncat + r) * K * K);
  double lnl = 0.0;
  const double log_min = log(OR_MINLH);
  for (int64_t p = 0; p < P; p++) {
    for (int i = 0; i < n_inner; i++) {
      double *out = clv + (size_t)i * RK;
      int sc = 0;
      for (int r = 0; r < ncat; r++) for (int k = 0; k < K; k++) out[r * K + k] = 1.0;
      for (int c = 0; c < 2; c++) {
        int ch = desc[2 * i + c];
        for (int r = 0; r < ncat; r++) {
          const double *Pm = pm + (((size_t)i * 2 + c) * ncat + r) * K * K;
          for (int k = 0; k < K; k++) {
            double s = 0;
            if (ch < n_tips) {
              uint32_t m = masks[codes[(size_t)ch * P + p]];
              for (int j = 0; j < K; j++) if (m >> j & 1) s += Pm[k * K + j];
            } else {
              const double *x = clv + (size_t)(ch - n_tips) * RK + r * K;
              for (int j = 0; j < K; j++) s += Pm[k * K + j] * x[j];
            }
            out[r * K + k] *= s;
          }
        }
        if (ch >= n_tips) sc += scl[ch - n_tips];
      }
      double mx = 0;
      for (int e = 0; e < RK; e++) if (fabs(out[e]) > mx) mx = fabs(out[e]);
      if (mx < OR_MINLH) { for (int e = 0; e < RK; e++) out[e] *= OR_TWO256; sc++; }
      scl[i] = sc;
    }
    const double *root = clv + (size_t)(n_inner - 1) * RK;
    double site = 0;
    for (int r = 0; r < ncat; r++) for (int k = 0; k < K; k++) site += freqs[k] * root[r * K + k];
    site /= ncat;
    lnl += weights[p] * (log(site) + scl[n_inner - 1] * log_min);
  }
  free(clv); free(scl); free(pm);
  return lnl;
}

/* Brute force for tiny trees: sum over all K^n_inner internal-state assignments,
 * no pruning, no scaling -- pins oracle_lnl itself. */
double oracle_lnl_bruteforce(int K, int n_tips, int64_t P, const uint8_t *codes,
                             const uint32_t *masks, const double *weights,
                             const double *exch, const double *freqs, double alpha,
                             int ncat, int n_inner, const int32_t *desc,
                             const double *brlen) {
  double eval[OR_MAXK], U[OR_MAXK * OR_MAXK], Uinv[OR_MAXK * OR_MAXK], rates[64], unit[OR_MAXK];
  freqs = oracle_unit_freqs(K, freqs, unit);
  oracle_eigen(K, exch, freqs, eval, U, Uinv);
  oracle_gamma_rates(alpha, ncat, rates);
  double *pm = (double *)malloc(sizeof(double) * (size_t)n_inner * 2 * K * K);
  int st[16];
  double lnl = 0;
  for (int64_t p = 0; p < P; p++) {
    double site = 0;
    for (int r = 0; r < ncat; r++) {
      for (int i = 0; i < n_inner; i++) for (int c = 0; c < 2; c++)
        oracle_pmatrix(K, eval, U, Uinv, brlen[2 * i + c] * rates[r], pm + ((size_t)i * 2 + c) * K * K);
      long total = 1;
      for (int i = 0; i < n_inner; i++) total *= K;
      for (long a = 0; a < total; a++) {
        long x = a;
        for (int i = 0; i < n_inner; i++) { st[i] = (int)(x % K); x /= K; }
        double pr = freqs[st[n_inner - 1]];
        for (int i = 0; i < n_inner; i++) for (int c = 0; c < 2; c++) {
          int ch = desc[2 * i + c];
          const double *Pm = pm + ((size_t)i * 2 + c) * K * K + st[i] * K;
          if (ch < n_tips) {
            uint32_t m = masks[codes[(size_t)ch * P + p]];
            double s = 0;
            for (int j = 0; j < K; j++) if (m >> j & 1) s += Pm[j];
            pr *= s;
          } else pr *= Pm[st[ch - n_tips]];
        }
        site += pr;
      }
    }
    lnl += weights[p] * log(site / ncat);
  }
  free(pm);
  return lnl;
}
