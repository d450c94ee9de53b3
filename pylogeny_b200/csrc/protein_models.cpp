// Empirical amino-acid model tables (order ARNDCQEGHILKMFPSTWYV).
//
// PROVENANCE.  libpll carries its protein matrices inside the library; they are not in
// /root/reference and no package on this box ships them (SURVEY.md section 7 "Protein model
// tables"), so both tables were typed from memory of the files PAML distributes:
//   WAG  wag.dat    (Whelan & Goldman 2001): 190 exchangeabilities + 20 frequencies
//   JTT  jones.dat  (Jones, Taylor & Thornton 1992): 190 integer counts-derived rates + 20 frequencies
// What IS checked (tests/test_host_cpu.py::test_protein_tables): 190 positive entries each; the
// frequencies sum to 1 within the files' own rounding (1e-6); the resulting Q is reversible with
// stationary distribution = the table's frequencies; the dominant exchanges of either paper come out
// on top (WAG: I-V, F-Y, D-E, N-D, K-R ...; JTT: I-V, D-E, K-R, Q-H, H-Y, F-Y ...); landmark entries
// (WAG s(R,A) = 0.551571, s(D,N) = 5.42942, s(V,I) = 7.8213; JTT s(R,A) = 58, s(E,D) = 767, s(V,I) = 961).
// What is NOT checked: the remaining digits of the other entries against the published files, and
// that libpll 1.0.2's built-in "JTT" equals jones.dat -- RAxML-family code carries the DCMut
// re-derivation of the same counts (Kosiol & Goldman 2005), whose entries differ by a few percent.
// Kernel parity (engine vs oracle) does not depend on the table; any reversible model can be
// supplied through plg_model_create.
#include "common.h"

namespace plg {

// lower triangle of PAML's wag.dat, row by row: s(R,A); s(N,A) s(N,R); ...
static const double WAG_LOWER[190] = {
    0.551571,
    0.509848, 0.635346,
    0.738998, 0.147304, 5.429420,
    1.027040, 0.528191, 0.265256, 0.0302949,
    0.908598, 3.035500, 1.543640, 0.616783, 0.0988179,
    1.582850, 0.439157, 0.947198, 6.174160, 0.021352, 5.469470,
    1.416720, 0.584665, 1.125560, 0.865584, 0.306674, 0.330052, 0.567717,
    0.316954, 2.137150, 3.956290, 0.930676, 0.248972, 4.294110, 0.570025, 0.249410,
    0.193335, 0.186979, 0.554236, 0.039437, 0.170135, 0.113917, 0.127395, 0.0304501, 0.138190,
    0.397915, 0.497671, 0.131528, 0.0848047, 0.384287, 0.869489, 0.154263, 0.0613037, 0.499462, 3.170970,
    0.906265, 5.351420, 3.012010, 0.479855, 0.0740339, 3.894900, 2.584430, 0.373558, 0.890432, 0.323832, 0.257555,
    0.893496, 0.683162, 0.198221, 0.103754, 0.390482, 1.545260, 0.315124, 0.174100, 0.404141, 4.257460, 4.854020,
    0.934276,
    0.210494, 0.102711, 0.0961621, 0.0467304, 0.398020, 0.0999208, 0.0811339, 0.049931, 0.679371, 1.059470, 2.115170,
    0.088836, 1.190630,
    1.438550, 0.679489, 0.195081, 0.423984, 0.109404, 0.933372, 0.682355, 0.243570, 0.696198, 0.0999288, 0.415844,
    0.556896, 0.171329, 0.161444,
    3.370790, 1.224190, 3.974230, 1.071760, 1.407660, 1.028870, 0.704939, 1.341820, 0.740169, 0.319440, 0.344739,
    0.967130, 0.493905, 0.545931, 1.613280,
    2.121110, 0.554413, 2.030060, 0.374866, 0.512984, 0.857928, 0.822765, 0.225833, 0.473307, 1.458160, 0.326622,
    1.386980, 1.516120, 0.171903, 0.795384, 4.378020,
    0.113133, 1.163920, 0.0719167, 0.129767, 0.717070, 0.215737, 0.156557, 0.336983, 0.262569, 0.212483, 0.665309,
    0.137505, 0.515706, 1.529640, 0.139405, 0.523742, 0.110864,
    0.240735, 0.381533, 1.086000, 0.325711, 0.543833, 0.227710, 0.196303, 0.103604, 3.873440, 0.420170, 0.398618,
    0.133264, 0.428437, 6.454280, 0.216046, 0.786993, 0.291148, 2.485390,
    2.006010, 0.251849, 0.196246, 0.152335, 1.002140, 0.301281, 0.588731, 0.187247, 0.118358, 7.821300, 1.800340,
    0.305434, 2.058450, 0.649892, 0.314887, 0.232739, 1.388230, 0.365369, 0.314730};

extern const double WAG_FREQS[20] = {0.0866279, 0.043972,  0.0390894, 0.0570451, 0.0193078, 0.0367281, 0.0580589,
                                     0.0832518, 0.0244313, 0.048466,  0.086209,  0.0620286, 0.0195027, 0.0384319,
                                     0.0457631, 0.0695179, 0.0610127, 0.0143859, 0.0352742, 0.0708956};

// lower triangle of PAML's jones.dat, row by row
static const double JTT_LOWER[190] = {
    58,
    54, 45,
    81, 16, 528,
    56, 113, 34, 10,
    57, 310, 86, 49, 9,
    105, 29, 58, 767, 5, 323,
    179, 137, 81, 130, 59, 26, 119,
    27, 328, 391, 112, 69, 597, 26, 23,
    36, 22, 47, 11, 17, 9, 12, 6, 16,
    30, 38, 12, 7, 23, 72, 9, 6, 56, 229,
    35, 646, 263, 26, 7, 292, 181, 27, 45, 21, 14,
    54, 44, 30, 15, 31, 43, 18, 14, 33, 479, 388, 65,
    15, 5, 10, 4, 78, 4, 5, 5, 40, 89, 248, 4, 43,
    194, 74, 15, 15, 14, 164, 18, 24, 115, 10, 102, 21, 16, 17,
    378, 101, 503, 59, 223, 53, 30, 201, 73, 40, 59, 47, 29, 92, 285,
    475, 64, 232, 38, 42, 51, 32, 33, 46, 245, 25, 103, 226, 12, 118, 477,
    9, 126, 8, 4, 115, 18, 10, 55, 8, 9, 52, 10, 24, 53, 6, 35, 12,
    11, 20, 70, 46, 209, 24, 7, 8, 573, 32, 24, 8, 18, 536, 10, 63, 21, 71,
    298, 17, 16, 31, 62, 20, 45, 47, 11, 961, 180, 14, 323, 62, 23, 38, 112, 25, 16};

extern const double JTT_FREQS[20] = {0.076748, 0.051691, 0.042645, 0.051544, 0.019803, 0.040752, 0.061830,
                                     0.073152, 0.022944, 0.053761, 0.091904, 0.058676, 0.023826, 0.040126,
                                     0.050901, 0.068765, 0.058565, 0.014261, 0.032102, 0.066005};

// upper triangle, row-major (the layout plg_model_create takes)
struct Upper {
  double v[190];
  explicit Upper(const double *lower) {
    int e = 0;
    for (int i = 0; i < 20; i++)
      for (int j = i + 1; j < 20; j++) v[e++] = lower[j * (j - 1) / 2 + i];
  }
};
static const Upper g_wag_upper(WAG_LOWER), g_jtt_upper(JTT_LOWER);
extern const double *const WAG_EXCH_PTR = g_wag_upper.v;
extern const double *const JTT_EXCH_PTR = g_jtt_upper.v;

// Built-in protein model by the name a partition file gives it (pll.py:102-115): 0 = unknown
bool protein_model_table(const char *name, const double **exch, const double **freqs) {
  auto eq = [&](const char *a) {
    for (const char *p = name, *q = a;; p++, q++) {
      const char c = (*p >= 'a' && *p <= 'z') ? *p - 32 : *p;
      if (c != *q) return false;
      if (!c) return true;
    }
  };
  if (eq("WAG")) { *exch = WAG_EXCH_PTR; *freqs = WAG_FREQS; return true; }
  if (eq("JTT")) { *exch = JTT_EXCH_PTR; *freqs = JTT_FREQS; return true; }
  return false;
}

}  // namespace plg

extern "C" int plg_protein_model(const char *name, double *exch, double *freqs) {
  PLG_API_BEGIN
  if (!name || !exch || !freqs) PLG_FAIL(PLG_ERR_ARG, "null argument");
  const double *e = nullptr, *f = nullptr;
  if (!plg::protein_model_table(name, &e, &f))
    PLG_FAIL(PLG_ERR_UNSUPPORTED, "protein model '%s': the engine carries WAG and JTT", name);
  for (int i = 0; i < 190; i++) exch[i] = e[i];
  for (int i = 0; i < 20; i++) freqs[i] = f[i];
  PLG_API_END
}
