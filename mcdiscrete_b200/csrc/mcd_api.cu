// mcd_api.cu -- the C ABI of libmcd.so (include/mcd.h): context, tables, runs, drains, unit hooks.
// No C++ exception crosses the boundary; every entry point returns an mcd_status.
#include <math.h>
#include <stdarg.h>
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

#include <new>
#include <string>
#include <vector>

#include "mcd_internal.h"

namespace {

thread_local std::string g_last_error;

constexpr double kPiEFour = 6.51408491409531e-14;     // parameters.py:7
constexpr double kBohrM = 5.29177210544e-11;          // scipy.constants 'Bohr radius' (CODATA 2022), parameters.py:8
constexpr uint64_t kStatsWords = MCD_HIST_TOTAL + sizeof(mcd_counters) / 8;

struct HostTable {
  float* cdf = nullptr;
  float* Es = nullptr;
  float* Ws = nullptr;
  int* nvalid = nullptr;
  uint32_t* guide = nullptr;
  uint32_t nE = 0, nW = 0;
};

}  // namespace

struct mcd_ctx {
  int device = 0;
  int sm_count = 0;
  mcd_params par;
  mcd_material mat;
  DevParams P;
  HostTable tab[MCD_MAX_TABLES];
  bool owns_tables = true;
  size_t l2_persist_bytes = 0;                // persisting-L2 carve-out (0 = not used)
  float4* sig_poly = nullptr;                 // per-cell quadratics of the cumulative cross-sections
  double sig_poly_err = 0.0;                  // verified max error relative to the total
  float* mott_dev[4] = {nullptr, nullptr, nullptr, nullptr};   // Mott table: Es, s, cdf, om (elastic = Mott)
  float4* sig_hot = nullptr;                  // hot records (Ruth / Ruth+Quinn boundaries, total) on a finer grid
  double sig_hot_err = 0.0;                   // verified max error of the two linear boundaries
  int slot_order[MCD_MAX_SHELLS] = {0};
  // outputs
  mcd_el_row* els = nullptr;
  uint64_t max_el = 0;
  uint32_t* scats = nullptr;
  uint32_t n_chunks = 0;                      // chunk pool size (MCD_CHUNK_ROWS rows each)
  unsigned int* chunk_next = nullptr;
  unsigned int* chunk_count = nullptr;
  unsigned int* nev = nullptr;                // events per electron of the last run
  uint32_t* sorted = nullptr;                 // rows regrouped in trajectory order (mcd_order_scats)
  uint64_t sorted_cap_rows = 0;
  unsigned long long* offsets = nullptr;      // max_el + 1
  unsigned long long* tile_tmp = nullptr;
  unsigned long long* total_dev = nullptr;
  int ordered = 0;                            // 0 no, 1 rows (mcd_order_scats), 2 columns (mcd_order_scats_columns),
                                              // 3 columns enqueued behind the run, not yet verified
  uint64_t col_stride = 0;                    // bytes between the float columns of the columnar layout ON THE DEVICE
  unsigned long long* ticket = nullptr;       // 8 words: [0..3] id tickets of the passes, [4..6] leftover counts
  uint32_t* rq[2] = {nullptr, nullptr};       // leftover queues of the tail hand-over (statistics-only DS runs)
  unsigned int rq_cap = 0;
  unsigned long long* stats = nullptr;
  int n_blocks = 0;
  TrajLaunchInfo li{};
  // last run
  uint64_t last_count = 0;
  bool seg_cached = false;
  // stored rows as runs of consecutive pool rows: (first pool row, rows, prefix of rows before the run)
  struct Run { uint64_t pool_row, rows, before; };
  std::vector<Run> runs;
  uint64_t rows_total = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  TrajTailStream tail;                        // high-priority side stream of the hand-over passes (mcd_tuning)
  bool timed = false;
  std::string err;
};

namespace {

int fail(mcd_ctx* ctx, int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  if (ctx) ctx->err = buf;
  return code;
}

#define CU(ctx, call)                                                                          \
  do {                                                                                         \
    cudaError_t e__ = (call);                                                                  \
    if (e__ != cudaSuccess)                                                                    \
      return fail(ctx, e__ == cudaErrorMemoryAllocation ? MCD_ERR_NOMEM : MCD_ERR_CUDA,        \
                  "%s failed: %s", #call, cudaGetErrorString(e__));                            \
  } while (0)

double gryz_sigma_host(double E, double Es, double ns) {          // crossSections.py:213-242
  if (E <= Es) return 0.0;
  const double U = E / Es;
  return ns * kPiEFour * pow((U - 1.) / (U + 1.), 1.5) *
         (1. + (2. * (1. - (1. / (2. * U))) / 3. * log(2.7 + sqrt(U - 1.)))) / (E * Es);
}

// sigma formulas in float64 (crossSections.py:37-58, 183-208, 247-284) for the host-side order proof
// elastic: ruther_sigma :37-58, ruther_sigma_wDefl :63-84 (Z^2 -> Z(Z+1)), ruther_N_sigma :89-112 and
// ruther_N_sigma_wDefl :115-143 (Nigam screening beta* = 0.48 * 5.43 Z^(2/3) / E)
double ruth_sigma_host(double E, double Z, int elastic) {
  const double zz = (elastic == MCD_ELASTIC_RUTH_VANILLA_WDEFL || elastic == MCD_ELASTIC_RUTH_NIGRAM_WDEFL)
                        ? Z * (Z + 1.) : Z * Z;
  if (elastic == MCD_ELASTIC_RUTH_NIGRAM || elastic == MCD_ELASTIC_RUTH_NIGRAM_WDEFL) {
    const double b = 0.48 * (5.43 * pow(Z, 2. / 3.) / E);
    return kPiEFour * zz / (4. * b * (1. + b) * E * E);
  }
  const double a = 3.4 * pow(Z, 0.666667) / E, Ek = E * 1e-3, r = (Ek + 511.) / (Ek + 1024.);
  return 5.21e-21 * (zz / (Ek * Ek)) * r * r * (4. * M_PI) / (a * (1. + a));
}
double moller_sigma_host(double E, double Wc, double nfree) {
  if (E <= Wc) return 0.;
  const double eps = Wc / E;
  return nfree * kPiEFour / (E * E) * (1. / eps - 1. / (1. - eps) + log(eps / (1. - eps)));
}
double quinn_sigma_host(double E, double Epl, double Ef, double n) {
  const double a = E / Ef, b = Epl / Ef;
  return Epl * log((sqrt(1. + b) - 1.) / (sqrt(a) - sqrt(a - b))) / (2. * kBohrM * E) / n * 1e4;
}

// The kernel evaluates the channels in slot order Ruth, Quinn, Moller, Gryzinski[order[]].  If that is the
// descending-sigma order (what pickFromSigmas' sort produces, scattering.py:48) with a 0.1 % margin at every
// one of 4096 log-spaced energies of [Emin, E0], the per-event order check is skipped.
int order_is_static(const mcd_params& p, const mcd_material& mt, const int* order) {
  const int n = 4096;
  for (int i = 0; i < n; ++i) {
    const double E = p.Emin * pow(p.E0 / p.Emin, (double)i / (n - 1));
    double s[MCD_MAX_CHANNELS];
    s[0] = ruth_sigma_host(E, mt.Z, p.elastic);
    s[1] = quinn_sigma_host(E, mt.plasmon_e, mt.fermi_e, mt.atnd);
    s[2] = moller_sigma_host(E, p.Wc, mt.n_val);
    for (int k = 0; k < mt.n_shells; ++k) s[3 + k] = gryz_sigma_host(E, mt.shell_E[order[k]], mt.shell_n[order[k]]);
    // A shell whose binding energy lies above E has sigma = 0 (crossSections.py:235-236); zeros must be TRAILING (the
    // stable descending sort of pickFromSigmas leaves them last, and a zero-width interval is never picked), every
    // positive value must exceed its successor by the margin.
    bool zeros = false;
    for (int k = 1; k < mt.n_shells + 3; ++k) {
      if (s[k] < 0.0) return 0;
      if (s[k] == 0.0) { zeros = true; continue; }
      if (zeros || !(s[k - 1] > 1.001 * s[k])) return 0;
    }
  }
  return 1;
}

// cumulative inverse mean free paths (1/Angstrom) in slot order at energy E, float64
void cum_host(const mcd_params& p, const mcd_material& mt, const int* order, double E, double* cum) {
  const double scale = mt.atnd * 1e-14;
  double c = ruth_sigma_host(E, mt.Z, p.elastic) * scale;
  cum[0] = c;
  c += quinn_sigma_host(E, mt.plasmon_e, mt.fermi_e, mt.atnd) * scale;
  cum[1] = c;
  c += moller_sigma_host(E, p.Wc, mt.n_val) * scale;
  cum[2] = c;
  for (int k = 0; k < mt.n_shells; ++k) {
    c += gryz_sigma_host(E, mt.shell_E[order[k]], mt.shell_n[order[k]]) * scale;
    cum[3 + k] = c;
  }
}

// Per-cell quadratics of cum_k(E): octaves [2^e, 2^(e+1)) cut into 2^logM cells, coordinate x in [0,1) inside a
// cell; interpolation at the three Chebyshev nodes of [0,1], verified on 7 further points per cell.  logM is raised
// until the worst error is below 1.5e-7 of the total (else the kernel keeps evaluating the formulas).
int build_sig_poly(mcd_ctx* c, const int* order) {
  const mcd_params& p = c->par;
  const mcd_material& mt = c->mat;
  const int nch = mt.n_shells + 3;
  int e_min, e_max;
  frexp(p.Emin * 0.999, &e_min); e_min -= 1;               // 2^e_min <= 0.999 Emin
  frexp(p.E0, &e_max); e_max -= 1;                         // 2^e_max <= E0 < 2^(e_max+1)
  if (e_min < 1 || e_max > 30) return 0;
  const double node[3] = {0.5 - 0.25 * sqrt(3.0), 0.5, 0.5 + 0.25 * sqrt(3.0)};
  for (int logM = 7; logM <= 11; ++logM) {
    const int M = 1 << logM;
    const int ncell = (e_max - e_min + 1) * M;
    std::vector<float4> poly((size_t)ncell * nch);
    double worst = 0.0;
    for (int cell = 0; cell < ncell; ++cell) {
      const double Ea = ldexp(1.0 + (double)(cell % M) / M, e_min + cell / M);
      const double dEc = ldexp(1.0 / M, e_min + cell / M);
      double f[3][MCD_MAX_CHANNELS];
      if (Ea + dEc <= mt.plasmon_e * 1.001) {                // below the Quinn threshold: never visited
        for (int k = 0; k < nch; ++k) poly[(size_t)cell * nch + k] = make_float4(1.f, 0.f, 0.f, 0.f);
        continue;
      }
      for (int i = 0; i < 3; ++i) cum_host(p, mt, order, fmax(Ea + node[i] * dEc, mt.plasmon_e * 1.001), f[i]);
      for (int k = 0; k < nch; ++k) {
        // Newton form through (node_i, f_i) expanded to c0 + c1 x + c2 x^2
        const double d01 = (f[1][k] - f[0][k]) / (node[1] - node[0]);
        const double d12 = (f[2][k] - f[1][k]) / (node[2] - node[1]);
        const double c2 = (d12 - d01) / (node[2] - node[0]);
        const double c1 = d01 - c2 * (node[0] + node[1]);
        const double c0 = f[0][k] - node[0] * (d01 - c2 * node[1]);
        poly[(size_t)cell * nch + k] = make_float4((float)c0, (float)c1, (float)c2, 0.f);
      }
      if (Ea + dEc < p.Emin * 0.999 || Ea > p.E0 * 1.001) continue;       // outside the visited range
      for (int j = 0; j < 7; ++j) {
        const double x = (j + 0.5) / 7.0;
        double ref[MCD_MAX_CHANNELS];
        const double E = Ea + x * dEc;
        if (E <= mt.plasmon_e * 1.001) continue;
        cum_host(p, mt, order, E, ref);
        for (int k = 0; k < nch; ++k) {
          const float4 q = poly[(size_t)cell * nch + k];
          const float xf = (float)x;
          const double got = (double)fmaf(fmaf(q.z, xf, q.y), xf, q.x);
          worst = fmax(worst, fabs(got - ref[k]) / ref[nch - 1]);
        }
      }
    }
    if (getenv("MCD_DEBUG")) fprintf(stderr, "[mcd] sig_poly logM=%d worst=%.3g\n", logM, worst);
    if (worst < 1.5e-7) {
      cudaFree(c->sig_poly); c->sig_poly = nullptr;
      if (cudaMalloc(&c->sig_poly, poly.size() * sizeof(float4)) != cudaSuccess) return 0;
      if (cudaMemcpy(c->sig_poly, poly.data(), poly.size() * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess) return 0;
      DevParams& P = c->P;
      P.sig_poly = c->sig_poly;
      P.sig_shift = 23 - logM;
      P.sig_base = (uint32_t)(((uint32_t)(e_min + 127) << 23) >> P.sig_shift);
      P.sig_ncell = ncell;
      P.sig_nch = nch;
      c->sig_poly_err = worst;
      return 1;
    }
  }
  return 0;
}

// Hot records (DevParams::sig_hot): per cell of a finer grid, linear interpolants of the Ruth and Ruth+Quinn
// boundaries (Chebyshev nodes of [0,1], verified <= 1.5e-6 of the total on 7 further points) and the quadratic of
// the total (verified <= 1.5e-7); the grid is refined until both hold.
int build_sig_hot(mcd_ctx* c, const int* order) {
  const mcd_params& p = c->par;
  const mcd_material& mt = c->mat;
  const int nch = mt.n_shells + 3;
  int e_min, e_max;
  frexp(p.Emin * 0.999, &e_min); e_min -= 1;
  frexp(p.E0, &e_max); e_max -= 1;
  if (e_min < 1 || e_max > 30) return 0;
  const double n2[2] = {0.5 - 0.25 * sqrt(2.0), 0.5 + 0.25 * sqrt(2.0)};
  const double n3[3] = {0.5 - 0.25 * sqrt(3.0), 0.5, 0.5 + 0.25 * sqrt(3.0)};
  for (int logM = 7; logM <= 13; ++logM) {
    const int M = 1 << logM;
    const int ncell = (e_max - e_min + 1) * M;
    std::vector<float4> hot((size_t)ncell * 2);
    double worst_lin = 0.0, worst_tot = 0.0;
    for (int cell = 0; cell < ncell; ++cell) {
      const double Ea = ldexp(1.0 + (double)(cell % M) / M, e_min + cell / M);
      const double dEc = ldexp(1.0 / M, e_min + cell / M);
      const double Elo = mt.plasmon_e * 1.001;
      if (Ea + dEc <= Elo) {                                   // below the Quinn threshold: never visited
        hot[2 * (size_t)cell] = make_float4(0.5f, 0.f, 0.75f, 0.f);
        hot[2 * (size_t)cell + 1] = make_float4(1.f, 0.f, 0.f, 0.f);
        continue;
      }
      double f2[2][MCD_MAX_CHANNELS], f3[3][MCD_MAX_CHANNELS];
      for (int i = 0; i < 2; ++i) cum_host(p, mt, order, fmax(Ea + n2[i] * dEc, Elo), f2[i]);
      for (int i = 0; i < 3; ++i) cum_host(p, mt, order, fmax(Ea + n3[i] * dEc, Elo), f3[i]);
      float lin[2][2];
      for (int k = 0; k < 2; ++k) {
        const double c1 = (f2[1][k] - f2[0][k]) / (n2[1] - n2[0]);
        lin[k][0] = (float)(f2[0][k] - c1 * n2[0]);
        lin[k][1] = (float)c1;
      }
      const int T = nch - 1;
      const double d01 = (f3[1][T] - f3[0][T]) / (n3[1] - n3[0]);
      const double d12 = (f3[2][T] - f3[1][T]) / (n3[2] - n3[1]);
      const double q2 = (d12 - d01) / (n3[2] - n3[0]);
      const double q1 = d01 - q2 * (n3[0] + n3[1]);
      const double q0 = f3[0][T] - n3[0] * (d01 - q2 * n3[1]);
      hot[2 * (size_t)cell] = make_float4(lin[0][0], lin[0][1], lin[1][0], lin[1][1]);
      hot[2 * (size_t)cell + 1] = make_float4((float)q0, (float)q1, (float)q2, 0.f);
      if (Ea + dEc < p.Emin * 0.999 || Ea > p.E0 * 1.001) continue;       // outside the visited range
      for (int j = 0; j < 7; ++j) {
        const double x = (j + 0.5) / 7.0;
        const double E = Ea + x * dEc;
        if (E <= Elo) continue;
        double ref[MCD_MAX_CHANNELS];
        cum_host(p, mt, order, E, ref);
        const float xf = (float)x;
        for (int k = 0; k < 2; ++k)
          worst_lin = fmax(worst_lin, fabs((double)fmaf(lin[k][1], xf, lin[k][0]) - ref[k]) / ref[T]);
        worst_tot = fmax(worst_tot, fabs((double)fmaf(fmaf((float)q2, xf, (float)q1), xf, (float)q0) - ref[T]) / ref[T]);
      }
    }
    if (getenv("MCD_DEBUG")) fprintf(stderr, "[mcd] sig_hot logM=%d worst_lin=%.3g worst_tot=%.3g\n", logM, worst_lin, worst_tot);
    if (worst_lin < 1.5e-6 && worst_tot < 1.5e-7) {
      cudaFree(c->sig_hot); c->sig_hot = nullptr;
      if (cudaMalloc(&c->sig_hot, hot.size() * sizeof(float4)) != cudaSuccess) return 0;
      if (cudaMemcpy(c->sig_hot, hot.data(), hot.size() * sizeof(float4), cudaMemcpyHostToDevice) != cudaSuccess) return 0;
      DevParams& P = c->P;
      P.sig_hot = c->sig_hot;
      P.hot_shift = 23 - logM;
      P.hot_base = (uint32_t)(((uint32_t)(e_min + 127) << 23) >> P.hot_shift);
      P.hot_ncell = ncell;
      P.hot_mask = (1u << P.hot_shift) - 1u;
      P.hot_last = (uint32_t)(ncell - 1);
      P.hot_scale = ldexpf(1.0f, -P.hot_shift);
      c->sig_hot_err = worst_lin;
      if (worst_tot > c->sig_poly_err) c->sig_poly_err = worst_tot;
      return 1;
    }
  }
  return 0;
}

void build_dev_params(mcd_ctx* c) {
  const mcd_params& p = c->par;
  const mcd_material& mt = c->mat;
  DevParams& P = c->P;
  memset(&P, 0, sizeof P);
  DevMat& m = P.m;
  const double scale = mt.atnd * 1e-14;                            // scattering.py:35
  const double Z = (double)mt.Z;
  m.sig_scale_inv = (float)(1.0 / scale);
  const double aZ = 3.4 * pow(Z, 0.666667);                        // crossSections.py:27
  m.aZ = (float)aZ;
  const double zz = (p.elastic == MCD_ELASTIC_RUTH_VANILLA_WDEFL || p.elastic == MCD_ELASTIC_RUTH_NIGRAM_WDEFL)
                        ? Z * (Z + 1.0) : Z * Z;
  if (p.elastic == MCD_ELASTIC_RUTH_NIGRAM || p.elastic == MCD_ELASTIC_RUTH_NIGRAM_WDEFL) {
    // pi e^4 zz / (4 b*/E (1 + b*/E) E^2) = pi e^4 zz / (4 b* (E + b*)),  b* = 0.48 * 5.43 Z^(2/3)
    const double bN = 0.48 * 5.43 * pow(Z, 2.0 / 3.0);
    m.aS = (float)bN;
    m.rN = m.rD = 0.0f;                                            // no relativistic factor: (E/E)^2
    m.cR = (float)(kPiEFour * zz * scale / (4.0 * bN));
  } else {
    m.aS = (float)aZ;
    m.rN = 511000.0f;
    m.rD = 1024000.0f;
    m.cR = (float)(5.21e-21 * zz * 1e6 * 4.0 * M_PI * scale / aZ);
  }
  m.cM = (float)(mt.n_val * kPiEFour * scale);
  m.Wc = (float)p.Wc;
  m.invWc = (float)(1.0 / p.Wc);
  const double b = mt.plasmon_e / mt.fermi_e;
  m.cQ = (float)(mt.plasmon_e * 1e-10 / (2.0 * kBohrM));
  m.Kq = (float)((sqrt(1.0 + b) - 1.0) / b);
  m.inv_Ef = (float)(1.0 / mt.fermi_e);
  m.bq = (float)b;
  m.Ef = (float)mt.fermi_e;
  m.Epl = (float)mt.plasmon_e;
  m.nshell = mt.n_shells;
  // Gryzinski slots: expected descending sigma at E0 (stable) so the in-kernel order check passes
  int order[MCD_MAX_SHELLS];
  double sg[MCD_MAX_SHELLS];
  for (int i = 0; i < mt.n_shells; ++i) {
    order[i] = i;
    sg[i] = gryz_sigma_host(p.E0, mt.shell_E[i], mt.shell_n[i]);
  }
  for (int i = 1; i < mt.n_shells; ++i) {                          // stable insertion sort, descending
    const int o = order[i];
    int j = i;
    while (j > 0 && sg[order[j - 1]] < sg[o]) { order[j] = order[j - 1]; --j; }
    order[j] = o;
  }
  for (int k = 0; k < MCD_MAX_SHELLS; ++k) {
    const int i = k < mt.n_shells ? order[k] : 0;
    const bool on = k < mt.n_shells;
    m.gShell[k] = on ? i : 0;
    m.gE[k] = on ? (float)mt.shell_E[i] : 3.0e38f;                 // unused slots: never ionisable
    m.gInvE[k] = on ? (float)(1.0 / mt.shell_E[i]) : 0.f;
    m.gC[k] = on ? (float)(mt.shell_n[i] * kPiEFour * scale / mt.shell_E[i]) : 0.f;
  }
  const double J = (mt.Z >= 13) ? (9.76 * Z + 58.5 / pow(Z, 0.19)) : 11.5 * Z;   // stoppingPowers.py:134-137
  m.spC = (float)(2.0 * kPiEFour * mt.atnd * 1e-14);
  m.spCZ = (float)(2.0 * kPiEFour * mt.atnd * 1e-14 * Z);
  m.cJ = (float)(1.166 / J);
  m.kJ = (float)(mt.bethe_k * J);
  m.n_val = (float)mt.n_val;
  m.inv_Eval = (float)(1.0 / mt.E_val);
  {
    double N = mt.n_val, C = mt.n_val * log(mt.E_val);
    for (int i = 0; i < mt.n_shells; ++i) { N += mt.shell_n[i]; C += mt.shell_n[i] * log(mt.shell_E[i]); }
    const double inv = exp(-C / N);
    m.spN = (float)N;
    m.inv_Eeff_hi = (float)inv;
    m.inv_Eeff_lo = (float)(inv - (double)m.inv_Eeff_hi);
  }

  P.E0 = (float)p.E0;
  P.Emin = (float)p.Emin;
  P.thickness = (float)p.thickness;
  const double t = p.s_tilt_deg * M_PI / 180.0;                    // multiScatter.py:97-100
  P.d0x = (float)(-sin(t));
  P.d0y = 0.f;
  P.d0z = (float)cos(t);
  P.maxScatt = (int)(p.maxScatt > 0x7fffffff ? 0x7fffffff : p.maxScatt);
  P.mode = p.mode;
  P.bethe = p.bethe;
  P.gryz_per_shell = (p.gryz_table_mode == MCD_GRYZ_PER_SHELL);
  P.seed_lo = (uint32_t)p.seed;
  P.seed_hi = (uint32_t)(p.seed >> 32);
  for (int r = 0; r < 10; ++r) {                                    // Philox key schedule
    P.pkey[2 * r] = P.seed_lo + (uint32_t)r * 0x9E3779B9u;
    P.pkey[2 * r + 1] = P.seed_hi + (uint32_t)r * 0xBB67AE85u;
  }
  // (elastic = Mott: the elastic cross-section is a step function of E that arrives with mcd_mott_upload; the
  // kernel checks the order per event and evaluates the channels directly, no per-cell records)
  P.order_static = (p.mode == MCD_MODE_DS && p.elastic != MCD_ELASTIC_MOTT) ? order_is_static(p, mt, order) : 0;
  memcpy(c->slot_order, order, sizeof(int) * MCD_MAX_SHELLS);
  P.scat_mask = p.scat_mask;
  P.flags = p.flags;
  // histogram axes: energy [Emin,E0], depth [0, 2R], R = E0 / S_Bethe,classical(E0)
  P.hist.e_lo = (float)p.Emin;
  P.hist.e_scale = (float)(MCD_HIST_E_BINS / (p.E0 - p.Emin));
  const double sp0 = 2.0 * kPiEFour * mt.atnd * (Z / p.E0) * log(1.166 * p.E0 / J) * 1e-14;
  P.hist.depth_scale = (float)(MCD_HIST_DEPTH_BINS / (2.0 * p.E0 / sp0));
  // launch tuning defaults (mcd_tuning_set changes them per context); the environment variables are the A/B
  // overrides of scripts/ab_kernel.py and only set the initial values
  auto env_int = [](const char* name, int dflt) { const char* v = getenv(name); return v ? atoi(v) : dflt; };
  P.tune_passes = env_int("MCD_PASSES", 3);
  P.tune_dump_T = env_int("MCD_DUMP_T", 31);
  P.tune_blocks_per_sm = env_int("MCD_BLOCKS_PER_SM", 0);
  P.tune_no_pool = env_int("MCD_NO_POOL", 0) != 0;
  P.tune_tail_priority = env_int("MCD_TAIL_PRIORITY", 1) != 0;
}

void layout_of(uint32_t mask, mcd_scat_layout* L) {
  int w = 2;
  for (int b = 0; b < 16; ++b) L->word_of[b] = -1;
  L->word_of[0] = (mask & MCD_SC_TYPE) ? 1 : -1;
  const uint32_t three = MCD_SC_POSITION | MCD_SC_DIRECTION;
  for (int b = 1; b <= 8; ++b) {
    if (!(mask & (1u << b))) continue;
    L->word_of[b] = w;
    w += ((1u << b) & three) ? 3 : 1;
  }
  L->row_words = (mask & 0x1ffu) ? (uint32_t)w : 0u;
}

void refresh_tables(mcd_ctx* c) {
  for (int s = 0; s < MCD_MAX_TABLES; ++s) {
    DevTable& d = c->P.tab[s];
    const HostTable& h = c->tab[s];
    d.cdf = h.cdf; d.Es = h.Es; d.Ws = h.Ws; d.nvalid = h.nvalid; d.guide = h.guide;
    d.nE = (int)h.nE; d.nW = (int)h.nW;
  }
}

void free_table(HostTable& h) {
  cudaFree(h.cdf); cudaFree(h.Es); cudaFree(h.Ws); cudaFree(h.nvalid); cudaFree(h.guide);
  h = HostTable();
}

int alloc_table(mcd_ctx* c, int slot, const float* Es, uint32_t nE, const float* Ws, uint32_t nW,
                double es_start_hint) {
  if (slot < 0 || slot >= MCD_MAX_TABLES) return fail(c, MCD_ERR_INVALID, "table slot %d out of range", slot);
  if (nE < 2 || nW < 3 || !Es || !Ws) return fail(c, MCD_ERR_INVALID, "table axes too small (nE=%u nW=%u)", nE, nW);
  if (!c->owns_tables) return fail(c, MCD_ERR_STATE, "tables of a cloned context belong to its parent");
  HostTable& h = c->tab[slot];
  free_table(h);
  CU(c, cudaMalloc(&h.cdf, (size_t)nE * (nW - 1) * sizeof(float)));
  CU(c, cudaMalloc(&h.Es, nE * sizeof(float)));
  CU(c, cudaMalloc(&h.Ws, nW * sizeof(float)));
  CU(c, cudaMalloc(&h.nvalid, nE * sizeof(int)));
  CU(c, cudaMalloc(&h.guide, (size_t)nE * (MCD_GUIDE + 1) * sizeof(uint32_t)));
  CU(c, cudaMemcpy(h.Es, Es, nE * sizeof(float), cudaMemcpyHostToDevice));
  CU(c, cudaMemcpy(h.Ws, Ws, nW * sizeof(float), cudaMemcpyHostToDevice));
  h.nE = nE;
  h.nW = nW;
  refresh_tables(c);
  DevTable& d = c->P.tab[slot];
  // axes must be strictly increasing (bisect_left semantics, scattering.py:78,89)
  for (uint32_t i = 1; i < nE; ++i)
    if (!(Es[i] > Es[i - 1])) return fail(c, MCD_ERR_INVALID, "table %d: Es is not strictly increasing at %u", slot, i);
  for (uint32_t i = 1; i < nW; ++i)
    if (!(Ws[i] > Ws[i - 1])) return fail(c, MCD_ERR_INVALID, "table %d: Ws is not strictly increasing at %u", slot, i);
  d.E_lo = Es[0];
  d.E_hi = Es[nE - 1];
  d.inv_dE = (float)((double)(nE - 1) / ((double)Es[nE - 1] - (double)Es[0]));
  // table_row() guesses the row as ceil((E - E_lo) * inv_dE - 0.01) and corrects UPWARD by at most one.  With
  // Es[i] = E_lo + (i + d_i) dE and a float32 guess that is off by e cells, the guess is never above the answer iff
  // max|d_i| + e <= 0.01, and never more than one below it.  Both are bounded here (e <= max|E| 2^-23 / dE); any
  // other axis (non-uniform, or finer than float32 resolves) is searched by a real bisection of Es[].
  {
    const double dE = ((double)Es[nE - 1] - (double)Es[0]) / (double)(nE - 1);
    double worst = 0.0;
    for (uint32_t i = 0; i < nE; ++i) worst = fmax(worst, fabs(((double)Es[i] - (double)Es[0]) / dE - (double)i));
    const double guess_err = fmax(fabs((double)Es[nE - 1]), fabs((double)Es[0])) * 1.2e-7 / dE;
    d.es_uniform = (worst + guess_err < 0.009) ? 1 : 0;
  }
  // Is Es exactly numpy.linspace(start, stop, nE, dtype=float32) for the (start, stop) the reference would use
  // (probTables.py:137-142, 364-379)?  Then the device regenerates grid values instead of loading them.
  d.es_analytic = 0;
  const double starts[2] = {c->par.Emin, es_start_hint};
  for (int k = 0; k < 2 && !d.es_analytic && !getenv("MCD_NO_ANALYTIC_ES"); ++k) {
    const double start = starts[k], stop = c->par.E0;
    if (!(start > 0) || !(stop > start)) continue;
    const double step = (stop - start) / (double)(nE - 1);
    bool same = ((float)stop == Es[nE - 1]);
    for (uint32_t i = 0; same && i + 1 < nE; ++i) same = ((float)((double)i * step + start) == Es[i]);
    if (same) { d.es_analytic = 1; d.es_start = start; d.es_step = step; }
  }
  return MCD_OK;
}

// temporary device buffer for the unit hooks
template <typename T>
struct Tmp {
  T* p = nullptr;
  ~Tmp() { cudaFree(p); }
  cudaError_t alloc(uint64_t n) { return cudaMalloc(&p, (n ? n : 1) * sizeof(T)); }
  cudaError_t up(const T* h, uint64_t n) {
    cudaError_t e = alloc(n);
    if (e != cudaSuccess) return e;
    return cudaMemcpy(p, h, n * sizeof(T), cudaMemcpyHostToDevice);
  }
  cudaError_t down(T* h, uint64_t n) { return cudaMemcpy(h, p, n * sizeof(T), cudaMemcpyDeviceToHost); }
};

int need_ctx(mcd_ctx* c) {
  if (!c) return fail(nullptr, MCD_ERR_INVALID, "null context");
  cudaError_t e = cudaSetDevice(c->device);
  if (e != cudaSuccess) return fail(c, MCD_ERR_CUDA, "cudaSetDevice(%d): %s", c->device, cudaGetErrorString(e));
  return MCD_OK;
}

}  // namespace

extern "C" {

int mcd_abi_version(void) { return MCD_ABI_VERSION; }

const char* mcd_last_error(const mcd_ctx* ctx) { return ctx ? ctx->err.c_str() : g_last_error.c_str(); }

int mcd_device_count(int* n) {
  if (!n) return fail(nullptr, MCD_ERR_INVALID, "null pointer");
  int k = 0;
  cudaError_t e = cudaGetDeviceCount(&k);
  if (e != cudaSuccess || k == 0) {
    *n = 0;
    return fail(nullptr, MCD_ERR_NO_DEVICE, "no usable CUDA device (%s); libmcd has no CPU fallback",
                e == cudaSuccess ? "count = 0" : cudaGetErrorString(e));
  }
  *n = k;
  return MCD_OK;
}

int mcd_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor,
                    uint64_t* total_mem) {
  cudaDeviceProp pr;
  cudaError_t e = cudaGetDeviceProperties(&pr, device);
  if (e != cudaSuccess) return fail(nullptr, MCD_ERR_NO_DEVICE, "cudaGetDeviceProperties(%d): %s", device,
                                    cudaGetErrorString(e));
  if (name && name_len > 0) snprintf(name, name_len, "%s", pr.name);
  if (sm_count) *sm_count = pr.multiProcessorCount;
  if (cc_major) *cc_major = pr.major;
  if (cc_minor) *cc_minor = pr.minor;
  if (total_mem) *total_mem = (uint64_t)pr.totalGlobalMem;
  return MCD_OK;
}

// the side stream of the hand-over passes: highest priority the device offers, non-blocking
static cudaError_t alloc_tail_stream(mcd_ctx* c) {
  int lo = 0, hi = 0;
  cudaError_t e = cudaDeviceGetStreamPriorityRange(&lo, &hi);
  if (e != cudaSuccess) return e;
  if ((e = cudaStreamCreateWithPriority(&c->tail.stream, cudaStreamNonBlocking, hi)) != cudaSuccess) return e;
  if ((e = cudaEventCreateWithFlags(&c->tail.first_done, cudaEventDisableTiming)) != cudaSuccess) return e;
  return cudaEventCreateWithFlags(&c->tail.tail_done, cudaEventDisableTiming);
}

// leftover queues of the tail hand-over: a warp hands over at most 32 electrons, 12 words each
static cudaError_t alloc_queues(mcd_ctx* c) {
  if (!mcd_traj_uses_queue(c->P, c->P.scat_row_words > 0)) return cudaSuccess;
  c->rq_cap = (unsigned int)(c->li.blocks * c->li.threads);
  for (int i = 0; i < 2; ++i) {
    cudaError_t e = cudaMalloc(&c->rq[i], (size_t)c->rq_cap * 12 * sizeof(uint32_t));
    if (e != cudaSuccess) return e;
  }
  return cudaSuccess;
}

int mcd_create(const mcd_params* p, const mcd_material* m, int device, mcd_ctx** out) {
  if (!p || !m || !out) return fail(nullptr, MCD_ERR_INVALID, "null argument");
  *out = nullptr;
  if (m->n_shells < 1 || m->n_shells > MCD_MAX_SHELLS)
    return fail(nullptr, MCD_ERR_INVALID, "n_shells=%d outside 1..%d", m->n_shells, MCD_MAX_SHELLS);
  if (p->mode != MCD_MODE_DS && p->mode != MCD_MODE_CONT) return fail(nullptr, MCD_ERR_INVALID, "bad mode %d", p->mode);
  if (p->bethe < 0 || p->bethe > 2) return fail(nullptr, MCD_ERR_INVALID, "bad Bethe model %d", p->bethe);
  if (p->elastic < MCD_ELASTIC_RUTH_VANILLA || p->elastic > MCD_ELASTIC_MOTT)
    return fail(nullptr, MCD_ERR_INVALID, "bad elastic model %d (Browning is not on the path)", p->elastic);
  if (p->elastic == MCD_ELASTIC_MOTT && p->mode != MCD_MODE_DS)
    return fail(nullptr, MCD_ERR_INVALID, "elastic = Mott exists in DS mode only (the reference's continuous classes "
                "never receive the table: scattering.py:794 raises AttributeError)");
  if (!(p->E0 > p->Emin) || !(p->Emin > 0)) return fail(nullptr, MCD_ERR_INVALID, "need E0 > Emin > 0");
  if (p->mode == MCD_MODE_DS && !(p->Emin > m->plasmon_e))
    return fail(nullptr, MCD_ERR_INVALID, "DS needs Emin > plasmon energy (quinn_sigma, crossSections.py:272-281)");
  if (p->maxScatt < 0) return fail(nullptr, MCD_ERR_INVALID, "maxScatt < 0");
  // the event index of a scats row, of a parked electron and of the regrouping key is a 24-bit field
  if (p->maxScatt >= (1 << 24))
    return fail(nullptr, MCD_ERR_INVALID, "maxScatt = %lld: at most 2^24 - 1 events per electron (24-bit event index)",
                (long long)p->maxScatt);
  int n = 0;
  int rc = mcd_device_count(&n);
  if (rc != MCD_OK) return rc;
  if (device < 0 || device >= n) return fail(nullptr, MCD_ERR_INVALID, "device %d of %d", device, n);
  cudaError_t e = cudaSetDevice(device);
  if (e != cudaSuccess) return fail(nullptr, MCD_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  mcd_ctx* c = new (std::nothrow) mcd_ctx();
  if (!c) return fail(nullptr, MCD_ERR_NOMEM, "out of host memory");
  c->device = device;
  c->par = *p;
  c->mat = *m;
  cudaDeviceProp pr;
  e = cudaGetDeviceProperties(&pr, device);
  if (e != cudaSuccess) { delete c; return fail(nullptr, MCD_ERR_CUDA, "props: %s", cudaGetErrorString(e)); }
  c->sm_count = pr.multiProcessorCount;
  if (getenv("MCD_L2_PERSIST") && pr.persistingL2CacheMaxSize > 0) {
    size_t want = (size_t)atoll(getenv("MCD_L2_PERSIST")) << 20;
    if (want > (size_t)pr.persistingL2CacheMaxSize) want = (size_t)pr.persistingL2CacheMaxSize;
    if (want && cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, want) == cudaSuccess) c->l2_persist_bytes = want;
    cudaGetLastError();
  }
  build_dev_params(c);
  if (getenv("MCD_DEBUG")) fprintf(stderr, "[mcd] order_static=%d\n", c->P.order_static);
  c->P.use_sig_poly = 0;
  if (p->mode == MCD_MODE_DS && c->P.order_static && !getenv("MCD_NO_SIGPOLY"))
    c->P.use_sig_poly = build_sig_poly(c, c->slot_order) && build_sig_hot(c, c->slot_order);
  mcd_scat_layout L;
  layout_of(p->scat_mask, &L);
  c->P.scat_row_words = (int)L.row_words;
  e = mcd_traj_max_blocks(c->P, c->P.scat_row_words > 0, c->sm_count, &c->li);
  if (e != cudaSuccess) {
    std::string msg = cudaGetErrorString(e);
    delete c;
    return fail(nullptr, MCD_ERR_CUDA, "trajectory kernel unusable on device %d (%s): %s -- libmcd is built for "
                "sm_100a only and has no fallback", device, pr.name, msg.c_str());
  }
  c->n_blocks = c->li.blocks;
  if ((e = cudaMalloc(&c->ticket, 8 * sizeof(unsigned long long))) != cudaSuccess || (e = alloc_queues(c)) != cudaSuccess ||
      (e = alloc_tail_stream(c)) != cudaSuccess || (e = cudaMalloc(&c->stats, kStatsWords * 8)) != cudaSuccess ||
      (e = cudaMalloc(&c->chunk_next, sizeof(unsigned int))) != cudaSuccess ||
      (e = cudaMemset(c->stats, 0, kStatsWords * 8)) != cudaSuccess ||
      (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess) {
    std::string msg = cudaGetErrorString(e);
    mcd_destroy(c);
    return fail(nullptr, MCD_ERR_CUDA, "context allocation: %s", msg.c_str());
  }
  *out = c;
  return MCD_OK;
}

int mcd_tuning_get(const mcd_ctx* c, mcd_tuning* out) {
  if (!c || !out) return fail(nullptr, MCD_ERR_INVALID, "null argument");
  out->handover_passes = c->P.tune_passes;
  out->handover_threshold = c->P.tune_dump_T;
  out->blocks_per_sm = c->P.tune_blocks_per_sm;
  out->statistics_kernel = c->P.tune_no_pool;
  out->tail_priority = c->P.tune_tail_priority;
  return MCD_OK;
}

int mcd_tuning_set(mcd_ctx* c, const mcd_tuning* t) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!t) return fail(c, MCD_ERR_INVALID, "null argument");
  if (t->handover_passes < 1 || t->handover_passes > 4 || t->handover_threshold < 0 || t->handover_threshold > 31 ||
      t->blocks_per_sm < 0 || t->statistics_kernel < 0 || t->statistics_kernel > 1 || t->tail_priority < 0 ||
      t->tail_priority > 1)
    return fail(c, MCD_ERR_INVALID, "tuning out of range (passes 1..4, threshold 0..31, blocks_per_sm >= 0, kernel 0|1, "
                "tail_priority 0|1)");
  if (c->els) return fail(c, MCD_ERR_STATE, "mcd_tuning_set must precede mcd_set_capacity");
  c->P.tune_passes = t->handover_passes;
  c->P.tune_dump_T = t->handover_threshold;
  c->P.tune_blocks_per_sm = t->blocks_per_sm;
  c->P.tune_no_pool = t->statistics_kernel;
  c->P.tune_tail_priority = t->tail_priority;
  // the kernel choice and its geometry may have changed: launch info and the hand-over queues follow
  CU(c, mcd_traj_max_blocks(c->P, c->P.scat_row_words > 0, c->sm_count, &c->li));
  c->n_blocks = c->li.blocks;
  for (int i = 0; i < 2; ++i) { cudaFree(c->rq[i]); c->rq[i] = nullptr; }
  CU(c, alloc_queues(c));
  return MCD_OK;
}

int mcd_clone(const mcd_ctx* parent, mcd_ctx** out) {
  if (!parent || !out) return fail(nullptr, MCD_ERR_INVALID, "null argument");
  *out = nullptr;
  cudaError_t e = cudaSetDevice(parent->device);
  if (e != cudaSuccess) return fail(nullptr, MCD_ERR_CUDA, "cudaSetDevice: %s", cudaGetErrorString(e));
  mcd_ctx* c = new (std::nothrow) mcd_ctx();
  if (!c) return fail(nullptr, MCD_ERR_NOMEM, "out of host memory");
  c->device = parent->device;
  c->sm_count = parent->sm_count;
  c->par = parent->par;
  c->mat = parent->mat;
  c->P = parent->P;                      // includes the table descriptors (device pointers)
  for (int s = 0; s < MCD_MAX_TABLES; ++s) c->tab[s] = parent->tab[s];
  c->owns_tables = false;
  c->li = parent->li;
  c->l2_persist_bytes = parent->l2_persist_bytes;
  c->n_blocks = parent->n_blocks;
  if ((e = cudaMalloc(&c->ticket, 8 * sizeof(unsigned long long))) != cudaSuccess || (e = alloc_queues(c)) != cudaSuccess ||
      (e = alloc_tail_stream(c)) != cudaSuccess || (e = cudaMalloc(&c->stats, kStatsWords * 8)) != cudaSuccess ||
      (e = cudaMalloc(&c->chunk_next, sizeof(unsigned int))) != cudaSuccess ||
      (e = cudaMemset(c->stats, 0, kStatsWords * 8)) != cudaSuccess ||
      (e = cudaEventCreate(&c->ev0)) != cudaSuccess || (e = cudaEventCreate(&c->ev1)) != cudaSuccess) {
    std::string msg = cudaGetErrorString(e);
    mcd_destroy(c);
    return fail(nullptr, MCD_ERR_CUDA, "clone allocation: %s", msg.c_str());
  }
  *out = c;
  return MCD_OK;
}

int mcd_destroy(mcd_ctx* c) {
  if (!c) return MCD_OK;
  cudaSetDevice(c->device);
  if (c->owns_tables) for (auto& t : c->tab) free_table(t);
  if (c->owns_tables) { cudaFree(c->sig_poly); cudaFree(c->sig_hot); for (float* q : c->mott_dev) cudaFree(q); }
  cudaFree(c->els); cudaFree(c->scats); cudaFree(c->chunk_next); cudaFree(c->chunk_count); cudaFree(c->ticket);
  cudaFree(c->rq[0]); cudaFree(c->rq[1]);
  cudaFree(c->nev); cudaFree(c->sorted); cudaFree(c->offsets); cudaFree(c->tile_tmp); cudaFree(c->total_dev);
  cudaFree(c->stats);
  if (c->ev0) cudaEventDestroy(c->ev0);
  if (c->ev1) cudaEventDestroy(c->ev1);
  if (c->tail.first_done) cudaEventDestroy(c->tail.first_done);
  if (c->tail.tail_done) cudaEventDestroy(c->tail.tail_done);
  if (c->tail.stream) cudaStreamDestroy(c->tail.stream);
  delete c;
  return MCD_OK;
}

// ---- tables ------------------------------------------------------------------------------------
int mcd_table_build(mcd_ctx* c, int slot, const float* Es, uint32_t nE, const float* Ws, uint32_t nW,
                    double shell_E, double shell_n, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  rc = alloc_table(c, slot, Es, nE, Ws, nW, shell_E);
  if (rc) return rc;
  HostTable& h = c->tab[slot];
  TableBuildArgs a;
  a.kind = slot == MCD_TABLE_MOLLER ? 0 : 1;
  a.Es = h.Es; a.Ws = h.Ws; a.cdf = h.cdf; a.nvalid = h.nvalid;
  a.nE = (int)nE; a.nW = (int)nW;
  a.Ef32 = (float)c->mat.fermi_e;
  a.numEl = a.kind == 0 ? c->mat.n_val : shell_n;
  a.Eb = shell_E;
  if (a.kind == 1 && !(shell_E > 0 && shell_n > 0)) return fail(c, MCD_ERR_INVALID, "Gryzinski table needs shell_E, shell_n");
  int blocks = c->sm_count * 4;
  if (blocks > (int)nE) blocks = (int)nE;
  Tmp<double> scratch;
  CU(c, scratch.alloc((uint64_t)blocks * (nW - 1)));
  a.scratch = scratch.p;
  CU(c, mcd_launch_table_build(a, blocks, (cudaStream_t)stream));
  CU(c, mcd_launch_table_guide(h.cdf, (int)nE, (int)nW, h.nvalid, h.guide, (cudaStream_t)stream));
  CU(c, cudaStreamSynchronize((cudaStream_t)stream));
  return MCD_OK;
}

int mcd_table_upload(mcd_ctx* c, int slot, const float* Es, uint32_t nE, const float* Ws, uint32_t nW,
                     const float* cdf) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!cdf) return fail(c, MCD_ERR_INVALID, "null cdf");
  // Gryzinski grids may start at the shell's binding energy = Ws[0] (probTables.py:125,137-142)
  rc = alloc_table(c, slot, Es, nE, Ws, nW, (double)Ws[0]);
  if (rc) return rc;
  HostTable& h = c->tab[slot];
  CU(c, cudaMemcpy(h.cdf, cdf, (size_t)nE * (nW - 1) * sizeof(float), cudaMemcpyHostToDevice));
  CU(c, mcd_launch_table_nvalid(h.cdf, (int)nE, (int)nW, h.nvalid, 0));
  CU(c, mcd_launch_table_guide(h.cdf, (int)nE, (int)nW, h.nvalid, h.guide, 0));
  CU(c, cudaDeviceSynchronize());
  return MCD_OK;
}

// largest float32 that is <= x: for a float32 v, (x < v) == (round_down(x) < v)
static float f32_round_down(double x) {
  float f = (float)x;
  if ((double)f > x) f = nextafterf(f, -INFINITY);
  return f;
}

int mcd_mott_upload(mcd_ctx* c, uint32_t nE, uint32_t nT, const double* Es, const double* sigmas, const double* cdf,
                    const double* thetas) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (c->par.elastic != MCD_ELASTIC_MOTT) return fail(c, MCD_ERR_STATE, "the context was not created with elastic = Mott");
  if (!c->owns_tables) return fail(c, MCD_ERR_STATE, "tables of a cloned context belong to its parent");
  if (!Es || !sigmas || !cdf || !thetas || nE < 2 || nT < 2) return fail(c, MCD_ERR_INVALID, "Mott table: null or too small");
  for (uint32_t i = 0; i < nE; ++i) {
    if (i && !(Es[i] > Es[i - 1])) return fail(c, MCD_ERR_INVALID, "Mott table: Es not strictly increasing at %u", i);
    if (!(sigmas[i] > 0.0)) return fail(c, MCD_ERR_INVALID, "Mott table: sigma[%u] is not positive", i);
  }
  // bisect_left(Es, E) == nE raises IndexError in the reference (scattering.py:493)
  if (c->par.E0 > Es[nE - 1])
    return fail(c, MCD_ERR_INVALID, "Mott table ends at %g eV but E0 = %g eV", Es[nE - 1], c->par.E0);
  std::vector<float> hEs(nE), hs(nE), hcdf((size_t)nE * nT), hom(nT);
  const double scale = c->mat.atnd * 1e-14;                           // scattering.py:35
  for (uint32_t i = 0; i < nE; ++i) {
    hEs[i] = f32_round_down(Es[i]);
    hs[i] = (float)(sigmas[i] * scale);
    for (uint32_t j = 0; j < nT; ++j) {
      const double v = cdf[(size_t)i * nT + j];
      if (j && v < cdf[(size_t)i * nT + j - 1]) return fail(c, MCD_ERR_INVALID, "Mott table: CDF row %u decreases at %u", i, j);
      hcdf[(size_t)i * nT + j] = f32_round_down(v);
    }
  }
  for (uint32_t j = 0; j < nT; ++j) {
    const double ch = cos(thetas[j] * 0.5 * M_PI / 180.0);              // scattering.py:122: cos(radians(theta/2))^2
    hom[j] = (float)(1.0 - ch * ch);
  }
  for (float*& q : c->mott_dev) { cudaFree(q); q = nullptr; }
  const size_t sz[4] = {nE * sizeof(float), nE * sizeof(float), (size_t)nE * nT * sizeof(float), nT * sizeof(float)};
  const float* src[4] = {hEs.data(), hs.data(), hcdf.data(), hom.data()};
  for (int k = 0; k < 4; ++k) {
    CU(c, cudaMalloc(&c->mott_dev[k], sz[k]));
    CU(c, cudaMemcpy(c->mott_dev[k], src[k], sz[k], cudaMemcpyHostToDevice));
  }
  DevMott& M = c->P.mott;
  M.Es = c->mott_dev[0]; M.s = c->mott_dev[1]; M.cdf = c->mott_dev[2]; M.om = c->mott_dev[3];
  M.nE = (int)nE; M.nT = (int)nT;
  return MCD_OK;
}

int mcd_table_shape(const mcd_ctx* c, int slot, uint32_t* nE, uint32_t* nW) {
  if (!c || slot < 0 || slot >= MCD_MAX_TABLES) return fail(nullptr, MCD_ERR_INVALID, "bad ctx/slot");
  if (nE) *nE = c->tab[slot].nE;
  if (nW) *nW = c->tab[slot].nW;
  return MCD_OK;
}

int mcd_table_download(mcd_ctx* c, int slot, float* out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (slot < 0 || slot >= MCD_MAX_TABLES || !c->tab[slot].cdf || !out)
    return fail(c, MCD_ERR_STATE, "table slot %d not present", slot);
  const HostTable& h = c->tab[slot];
  CU(c, cudaMemcpy(out, h.cdf, (size_t)h.nE * (h.nW - 1) * sizeof(float), cudaMemcpyDeviceToHost));
  return MCD_OK;
}

// ---- runs --------------------------------------------------------------------------------------
int mcd_set_capacity(mcd_ctx* c, uint64_t max_electrons, uint64_t max_scat_rows) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (max_electrons == 0) return fail(c, MCD_ERR_INVALID, "max_electrons = 0");
  if (max_electrons >> 32) return fail(c, MCD_ERR_INVALID, "at most 2^32-1 electrons per run (scats word 0 is 32 bit)");
  cudaFree(c->els); c->els = nullptr;
  cudaFree(c->scats); c->scats = nullptr;
  cudaFree(c->chunk_count); c->chunk_count = nullptr;
  cudaFree(c->nev); c->nev = nullptr;
  cudaFree(c->sorted); c->sorted = nullptr; c->sorted_cap_rows = 0;
  cudaFree(c->offsets); c->offsets = nullptr;
  cudaFree(c->tile_tmp); c->tile_tmp = nullptr;
  cudaFree(c->total_dev); c->total_dev = nullptr;
  c->max_el = 0; c->n_chunks = 0;
  CU(c, cudaMalloc(&c->els, max_electrons * sizeof(mcd_el_row)));
  c->max_el = max_electrons;
  if (max_scat_rows && c->P.scat_row_words) {
    // every warp of the resident grid may leave one partly filled chunk
    const uint64_t warps = (uint64_t)c->n_blocks * (c->li.threads / 32);
    const uint64_t nch = (max_scat_rows + MCD_CHUNK_ROWS - 1) / MCD_CHUNK_ROWS + warps;
    if (nch >> 31) return fail(c, MCD_ERR_INVALID, "scats pool too large");
    CU(c, cudaMalloc(&c->scats, nch * MCD_CHUNK_ROWS * (uint64_t)c->P.scat_row_words * 4));
    CU(c, cudaMalloc(&c->chunk_count, nch * sizeof(unsigned int)));
    CU(c, cudaMalloc(&c->nev, max_electrons * sizeof(unsigned int)));
    c->n_chunks = (uint32_t)nch;
  }
  return MCD_OK;
}

int mcd_scat_layout_get(const mcd_ctx* c, mcd_scat_layout* out) {
  if (!c || !out) return fail(nullptr, MCD_ERR_INVALID, "null argument");
  layout_of(c->par.scat_mask, out);
  return MCD_OK;
}

int mcd_reset_stats(mcd_ctx* c, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  CU(c, cudaMemsetAsync(c->stats, 0, kStatsWords * 8, (cudaStream_t)stream));
  return MCD_OK;
}

int mcd_run(mcd_ctx* c, uint64_t first_id, uint64_t count, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!c->els) return fail(c, MCD_ERR_STATE, "mcd_set_capacity has not been called");
  if (count == 0 || count > c->max_el) return fail(c, MCD_ERR_INVALID, "count %llu outside 1..%llu",
                                                   (unsigned long long)count, (unsigned long long)c->max_el);
  if (c->par.mode == MCD_MODE_DS) {
    if (!c->tab[MCD_TABLE_MOLLER].cdf || !c->tab[MCD_TABLE_GRYZ0].cdf)
      return fail(c, MCD_ERR_STATE, "DS mode needs the Moller table (slot 0) and Gryzinski table (slot 1)");
    if (c->P.gryz_per_shell)
      for (int k = 0; k < c->mat.n_shells; ++k)
        if (!c->tab[MCD_TABLE_GRYZ0 + k].cdf) return fail(c, MCD_ERR_STATE, "per-shell mode: Gryzinski table %d missing", k);
    if (c->par.elastic == MCD_ELASTIC_MOTT && !c->P.mott.nE)
      return fail(c, MCD_ERR_STATE, "elastic = Mott needs the Mott table (mcd_mott_upload)");
  }
  cudaStream_t s = (cudaStream_t)stream;
  CU(c, cudaMemsetAsync(c->ticket, 0, 8 * sizeof(unsigned long long), s));
  CU(c, cudaMemsetAsync(c->chunk_next, 0, sizeof(unsigned int), s));
  if (c->n_chunks) CU(c, cudaMemsetAsync(c->chunk_count, 0, (size_t)c->n_chunks * sizeof(unsigned int), s));
  DevOut O{};
  O.els = c->els;
  O.scats = c->scats;
  O.n_chunks = c->n_chunks;
  O.chunk_next = c->chunk_next;
  O.chunk_count = c->chunk_count;
  O.nev = c->nev;
  O.ticket = c->ticket;
  O.stats = c->stats;
  O.rq[0] = c->rq[0];
  O.rq[1] = c->rq[1];
  O.rq_counts = c->ticket + 4;
  O.rq_cap = c->rq_cap;
  O.dump_T = -1;
  if (c->par.mode == MCD_MODE_DS && c->l2_persist_bytes > 0 && s != nullptr) {
    // keep the Moller guide (the first, L2-sized stage of two thirds of the table searches) resident in L2 while
    // the 1 GB of CDF rows streams through it
    cudaStreamAttrValue attr;
    memset(&attr, 0, sizeof attr);
    const HostTable& h = c->tab[MCD_TABLE_MOLLER];
    size_t bytes = (size_t)h.nE * (MCD_GUIDE + 1) * sizeof(uint32_t);
    if (bytes > c->l2_persist_bytes) bytes = c->l2_persist_bytes;
    attr.accessPolicyWindow.base_ptr = h.guide;
    attr.accessPolicyWindow.num_bytes = bytes;
    attr.accessPolicyWindow.hitRatio = 1.0f;
    attr.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
    attr.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    cudaStreamSetAttribute(s, cudaStreamAttributeAccessPolicyWindow, &attr);      // best effort
    cudaGetLastError();
  }
  CU(c, cudaEventRecord(c->ev0, s));
  CU(c, mcd_launch_traj(c->P, O, first_id, count, c->sm_count, s, &c->li, &c->tail));
  CU(c, cudaEventRecord(c->ev1, s));
  c->last_count = count;
  c->seg_cached = false;
  c->ordered = 0;
  c->timed = true;
  return MCD_OK;
}

int mcd_sync(mcd_ctx* c, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  CU(c, cudaStreamSynchronize((cudaStream_t)stream));
  return MCD_OK;
}

static int cache_segments(mcd_ctx* c) {
  if (c->seg_cached) return MCD_OK;
  c->runs.clear();
  c->rows_total = 0;
  if (c->scats && c->n_chunks) {
    unsigned int used = 0;
    CU(c, cudaMemcpy(&used, c->chunk_next, sizeof used, cudaMemcpyDeviceToHost));
    if (used > c->n_chunks) used = c->n_chunks;
    std::vector<unsigned int> cnt(used);
    if (used) CU(c, cudaMemcpy(cnt.data(), c->chunk_count, used * sizeof(unsigned int), cudaMemcpyDeviceToHost));
    // merge consecutive chunks into runs: a run continues while the previous chunk is full
    bool open = false;
    for (unsigned int k = 0; k < used; ++k) {
      if (cnt[k] == 0) { open = false; continue; }
      if (open) c->runs.back().rows += cnt[k];
      else c->runs.push_back({(uint64_t)k * MCD_CHUNK_ROWS, cnt[k], c->rows_total});
      c->rows_total += cnt[k];
      open = (cnt[k] == MCD_CHUNK_ROWS);
    }
  }
  c->seg_cached = true;
  return MCD_OK;
}

int mcd_row_counts(mcd_ctx* c, uint64_t* n_els, uint64_t* n_scats) {
  int rc = need_ctx(c);
  if (rc) return rc;
  rc = cache_segments(c);
  if (rc) return rc;
  if (n_els) *n_els = c->last_count;
  if (n_scats) *n_scats = c->rows_total;
  return MCD_OK;
}

int mcd_drain_els(mcd_ctx* c, mcd_el_row* host, uint64_t offset, uint64_t n, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!host || offset + n > c->last_count) return fail(c, MCD_ERR_INVALID, "els range outside the last run");
  if (n == 0) return MCD_OK;
  CU(c, cudaMemcpyAsync(host, c->els + offset, n * sizeof(mcd_el_row), cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return MCD_OK;
}

int mcd_drain_scats(mcd_ctx* c, uint32_t* host, uint64_t offset, uint64_t n, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  rc = cache_segments(c);
  if (rc) return rc;
  const uint64_t total = c->rows_total;
  if (!host || offset + n > total) return fail(c, MCD_ERR_INVALID, "scats range outside the stored rows");
  const uint64_t rw = (uint64_t)c->P.scat_row_words;
  uint64_t done = 0;
  // first run containing `offset`
  size_t lo = 0, hi = c->runs.size();
  while (lo < hi) {
    const size_t mid = (lo + hi) / 2;
    if (c->runs[mid].before + c->runs[mid].rows <= offset) lo = mid + 1; else hi = mid;
  }
  for (size_t r = lo; r < c->runs.size() && done < n; ++r) {
    const auto& R = c->runs[r];
    const uint64_t from = offset + done;
    const uint64_t skip = from - R.before;
    const uint64_t k = (R.rows - skip < n - done) ? R.rows - skip : n - done;
    const uint32_t* src = c->scats + (R.pool_row + skip) * rw;
    CU(c, cudaMemcpyAsync(host + done * rw, src, k * rw * 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    done += k;
  }
  return MCD_OK;
}

static int order_impl(mcd_ctx* c, void* stream, bool columns) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!c->scats || !c->nev || !c->last_count) return fail(c, MCD_ERR_STATE, "no scats rows to order");
  cudaStream_t s = (cudaStream_t)stream;
  CU(c, cudaStreamSynchronize(s));
  rc = cache_segments(c);
  if (rc) return rc;
  mcd_counters cn;
  CU(c, cudaMemcpy(&cn, c->stats + MCD_HIST_TOTAL, sizeof cn, cudaMemcpyDeviceToHost));
  if (cn.scat_rows_dropped) return fail(c, MCD_ERR_OVERFLOW, "rows were dropped: cannot regroup an incomplete run");
  if (!c->offsets) {
    CU(c, cudaMalloc(&c->offsets, (c->max_el + 1) * sizeof(unsigned long long)));
    CU(c, cudaMalloc(&c->tile_tmp, mcd_order_tiles(c->max_el) * sizeof(unsigned long long)));
    CU(c, cudaMalloc(&c->total_dev, sizeof(unsigned long long)));
  }
  if (c->sorted_cap_rows < c->rows_total) {
    cudaFree(c->sorted); c->sorted = nullptr; c->sorted_cap_rows = 0;
    const uint64_t cap = c->rows_total + c->rows_total / 8 + 1024;
    CU(c, cudaMalloc(&c->sorted, cap * (uint64_t)c->P.scat_row_words * 4));
    c->sorted_cap_rows = cap;
  }
  unsigned int used = 0;
  CU(c, cudaMemcpy(&used, c->chunk_next, sizeof used, cudaMemcpyDeviceToHost));
  if (used > c->n_chunks) used = c->n_chunks;
  // columnar layout: (row_words - 2) float columns of rows_total entries, each starting on a 256-byte boundary,
  // then the channel bytes; it always fits the row-major buffer (4 (rw-2) + 1 < 4 rw bytes per row, +1024 rows slack)
  c->col_stride = columns ? ((c->rows_total * 4 + 255) / 256) * 256 : 0;
  if (columns && c->P.scat_row_words < 2) return fail(c, MCD_ERR_STATE, "no scats columns");
  CU(c, mcd_launch_order_scats(c->nev, c->last_count, c->tile_tmp, c->offsets, c->total_dev, c->scats, c->chunk_count,
                               used, c->P.scat_row_words, c->sorted, c->col_stride,
                               columns ? (c->col_stride >> 2) : c->sorted_cap_rows, nullptr, c->sm_count, s));
  unsigned long long total = 0;
  CU(c, cudaMemcpyAsync(&total, c->total_dev, sizeof total, cudaMemcpyDeviceToHost, s));
  CU(c, cudaStreamSynchronize(s));
  if (total != c->rows_total)
    return fail(c, MCD_ERR_STATE, "event counts (%llu) and stored rows (%llu) disagree", total,
                (unsigned long long)c->rows_total);
  c->ordered = columns ? 2 : 1;
  return MCD_OK;
}

int mcd_order_scats(mcd_ctx* c, void* stream) { return order_impl(c, stream, false); }
int mcd_order_scats_columns(mcd_ctx* c, void* stream) { return order_impl(c, stream, true); }

int mcd_order_scats_columns_enqueue(mcd_ctx* c, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!c->scats || !c->nev || !c->last_count) return fail(c, MCD_ERR_STATE, "no scats rows to order");
  if (c->P.scat_row_words < 2) return fail(c, MCD_ERR_STATE, "no scats columns");
  cudaStream_t s = (cudaStream_t)stream;
  if (!c->offsets) {
    CU(c, cudaMalloc(&c->offsets, (c->max_el + 1) * sizeof(unsigned long long)));
    CU(c, cudaMalloc(&c->tile_tmp, mcd_order_tiles(c->max_el) * sizeof(unsigned long long)));
    CU(c, cudaMalloc(&c->total_dev, sizeof(unsigned long long)));
  }
  // the row count is not known on the host yet: columns are laid out for the capacity of the chunk pool
  const uint64_t rw = (uint64_t)c->P.scat_row_words;
  const uint64_t cap_rows = (uint64_t)c->n_chunks * MCD_CHUNK_ROWS;
  const uint64_t stride = ((cap_rows * 4 + 255) / 256) * 256;
  const uint64_t need_rows = ((rw - 2) * stride + cap_rows + rw * 4 - 1) / (rw * 4);
  if (c->sorted_cap_rows < need_rows) {
    cudaFree(c->sorted); c->sorted = nullptr; c->sorted_cap_rows = 0;
    CU(c, cudaMalloc(&c->sorted, need_rows * rw * 4));
    c->sorted_cap_rows = need_rows;
  }
  c->col_stride = stride;
  CU(c, mcd_launch_order_scats(c->nev, c->last_count, c->tile_tmp, c->offsets, c->total_dev, c->scats, c->chunk_count,
                               c->n_chunks, c->P.scat_row_words, c->sorted, c->col_stride, c->col_stride >> 2,
                               c->chunk_next, c->sm_count, s));
  c->ordered = 3;
  return MCD_OK;
}

// An enqueued regrouping is checked on first use (the caller has synchronised the stream by then).
static int columns_ready(mcd_ctx* c) {
  if (c->ordered == 2) return MCD_OK;
  if (c->ordered != 3) return fail(c, MCD_ERR_STATE, "mcd_order_scats_columns has not been called for the last run");
  mcd_counters cn;
  unsigned long long total = 0;
  CU(c, cudaMemcpy(&cn, c->stats + MCD_HIST_TOTAL, sizeof cn, cudaMemcpyDeviceToHost));
  CU(c, cudaMemcpy(&total, c->total_dev, sizeof total, cudaMemcpyDeviceToHost));
  if (cn.scat_rows_dropped) return fail(c, MCD_ERR_OVERFLOW, "rows were dropped: cannot regroup an incomplete run");
  if (total != (unsigned long long)cn.scat_rows_stored)
    return fail(c, MCD_ERR_STATE, "event counts (%llu) and stored rows (%llu) disagree", total,
                (unsigned long long)cn.scat_rows_stored);
  c->rows_total = (uint64_t)cn.scat_rows_stored;
  c->ordered = 2;
  return MCD_OK;
}

// host layout of the drained block: columns packed for the actual row count
int mcd_scat_columns_layout(mcd_ctx* c, uint64_t* n_rows, uint64_t* n_float_cols, uint64_t* col_stride,
                            uint64_t* type_offset, uint64_t* total_bytes) {
  int rc = need_ctx(c);
  if (rc) return rc;
  rc = columns_ready(c);
  if (rc) return rc;
  const uint64_t nf = (uint64_t)c->P.scat_row_words - 2;
  const uint64_t hs = ((c->rows_total * 4 + 255) / 256) * 256;
  if (n_rows) *n_rows = c->rows_total;
  if (n_float_cols) *n_float_cols = nf;
  if (col_stride) *col_stride = hs;
  if (type_offset) *type_offset = nf * hs;
  if (total_bytes) *total_bytes = nf * hs + c->rows_total;
  return MCD_OK;
}

int mcd_drain_scats_columns(mcd_ctx* c, void* host, uint64_t capacity_bytes, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  rc = columns_ready(c);
  if (rc) return rc;
  const uint64_t nf = (uint64_t)c->P.scat_row_words - 2, n = c->rows_total;
  const uint64_t hs = ((n * 4 + 255) / 256) * 256;
  if (!host || capacity_bytes < nf * hs + n) return fail(c, MCD_ERR_INVALID, "host buffer smaller than the column block");
  if (!n) return MCD_OK;
  const unsigned char* dev = reinterpret_cast<const unsigned char*>(c->sorted);
  unsigned char* h = static_cast<unsigned char*>(host);
  if (hs == c->col_stride) {                                          // device layout == host layout: one copy
    CU(c, cudaMemcpyAsync(h, dev, nf * hs + n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
    return MCD_OK;
  }
  for (uint64_t j = 0; j < nf; ++j)
    CU(c, cudaMemcpyAsync(h + j * hs, dev + j * c->col_stride, n * 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  CU(c, cudaMemcpyAsync(h + nf * hs, dev + nf * c->col_stride, n, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return MCD_OK;
}

int mcd_drain_scats_ordered(mcd_ctx* c, uint32_t* host, uint64_t offset, uint64_t n, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (c->ordered != 1) return fail(c, MCD_ERR_STATE, "mcd_order_scats has not been called for the last run");
  if (!host || offset + n > c->rows_total) return fail(c, MCD_ERR_INVALID, "scats range outside the stored rows");
  const uint64_t rw = (uint64_t)c->P.scat_row_words;
  if (n) CU(c, cudaMemcpyAsync(host, c->sorted + offset * rw, n * rw * 4, cudaMemcpyDeviceToHost, (cudaStream_t)stream));
  return MCD_OK;
}

int mcd_scat_offsets(mcd_ctx* c, uint64_t* host, uint64_t first, uint64_t n, void* stream) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (c->ordered == 3) {
    rc = columns_ready(c);
    if (rc) return rc;
  }
  if (!c->ordered) return fail(c, MCD_ERR_STATE, "mcd_order_scats has not been called for the last run");
  if (!host || first + n > c->last_count) return fail(c, MCD_ERR_INVALID, "offset range outside the last run");
  CU(c, cudaMemcpyAsync(host, c->offsets + first, (n + 1) * sizeof(unsigned long long), cudaMemcpyDeviceToHost,
                        (cudaStream_t)stream));
  return MCD_OK;
}

int mcd_counters_get(mcd_ctx* c, mcd_counters* out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!out) return fail(c, MCD_ERR_INVALID, "null pointer");
  CU(c, cudaMemcpy(out, c->stats + MCD_HIST_TOTAL, sizeof(mcd_counters), cudaMemcpyDeviceToHost));
  return MCD_OK;
}

int mcd_hist_get(mcd_ctx* c, int64_t* host) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!host) return fail(c, MCD_ERR_INVALID, "null pointer");
  CU(c, cudaMemcpy(host, c->stats, MCD_HIST_TOTAL * 8, cudaMemcpyDeviceToHost));
  return MCD_OK;
}

int mcd_stats_device_ptr(mcd_ctx* c, void** dev_ptr, uint64_t* n_int64) {
  if (!c || !dev_ptr || !n_int64) return fail(nullptr, MCD_ERR_INVALID, "null argument");
  *dev_ptr = c->stats;
  *n_int64 = kStatsWords;
  return MCD_OK;
}

int mcd_stats_merge(mcd_ctx* dst, mcd_ctx* src, void* stream) {
  int rc = need_ctx(dst);
  if (rc) return rc;
  if (!src || src == dst || src->device != dst->device) return fail(dst, MCD_ERR_INVALID, "merge needs a second context on the same device");
  CU(dst, mcd_launch_stats_merge(dst->stats, src->stats, (int)kStatsWords, (cudaStream_t)stream));
  return MCD_OK;
}

int mcd_hist_ranges(const mcd_ctx* c, double* e_lo, double* e_hi, double* depth_max) {
  if (!c) return fail(nullptr, MCD_ERR_INVALID, "null context");
  if (e_lo) *e_lo = c->par.Emin;
  if (e_hi) *e_hi = c->par.E0;
  if (depth_max) *depth_max = (double)MCD_HIST_DEPTH_BINS / (double)c->P.hist.depth_scale;
  return MCD_OK;
}

int mcd_last_kernel_ms(mcd_ctx* c, float* ms) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!ms || !c->timed) return fail(c, MCD_ERR_STATE, "no run has been timed");
  CU(c, cudaEventSynchronize(c->ev1));
  CU(c, cudaEventElapsedTime(ms, c->ev0, c->ev1));
  return MCD_OK;
}

int mcd_launch_info(mcd_ctx* c, int* blocks, int* threads, int* regs, int* blocks_per_sm) {
  if (!c) return fail(nullptr, MCD_ERR_INVALID, "null context");
  if (blocks) *blocks = c->li.blocks;
  if (threads) *threads = c->li.threads;
  if (regs) *regs = c->li.regs;
  if (blocks_per_sm) *blocks_per_sm = c->li.blocks_per_sm;
  return MCD_OK;
}

int mcd_last_launches(mcd_ctx* c, int* launches) {
  if (!c || !launches) return fail(nullptr, MCD_ERR_INVALID, "null argument");
  *launches = c->li.launches;
  return MCD_OK;
}

// ---- unit hooks --------------------------------------------------------------------------------
int mcd_test_sigmas(mcd_ctx* c, const float* E, uint64_t n, float* sigma_out, float* mfp_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  const int nch = c->mat.n_shells + 3;
  Tmp<float> dE, dS, dM;
  CU(c, dE.up(E, n)); CU(c, dS.alloc(n * nch)); CU(c, dM.alloc(n));
  CU(c, mcd_launch_test_sigmas(c->P, dE.p, n, dS.p, dM.p, 0));
  CU(c, dS.down(sigma_out, n * nch)); CU(c, dM.down(mfp_out, n));
  return MCD_OK;
}

int mcd_test_cum(mcd_ctx* c, const float* E, uint64_t n, float* cum_out, int32_t* slot_channel_out,
                 int32_t* uses_poly, double* poly_err) {
  int rc = need_ctx(c);
  if (rc) return rc;
  const int nch = c->mat.n_shells + 3;
  if (n) {
    Tmp<float> dE, dC;
    CU(c, dE.up(E, n)); CU(c, dC.alloc(n * nch));
    CU(c, mcd_launch_test_cum(c->P, dE.p, n, dC.p, 0));
    CU(c, dC.down(cum_out, n * nch));
  }
  if (slot_channel_out) {
    slot_channel_out[0] = 0; slot_channel_out[1] = 2 + c->mat.n_shells; slot_channel_out[2] = 1;
    for (int k = 0; k < c->mat.n_shells; ++k) slot_channel_out[3 + k] = 2 + c->P.m.gShell[k];
  }
  if (uses_poly) *uses_poly = c->P.use_sig_poly;
  if (poly_err) *poly_err = c->sig_poly_err;
  return MCD_OK;
}

int mcd_test_select(mcd_ctx* c, const float* sigma, int nch, const float* u, uint64_t n, int32_t* pick_out,
                    int32_t* path_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (nch < 1 || nch > MCD_MAX_CHANNELS) return fail(c, MCD_ERR_INVALID, "nch=%d outside 1..%d", nch, MCD_MAX_CHANNELS);
  Tmp<float> dS, dU;
  Tmp<int> dP, dQ;
  CU(c, dS.up(sigma, n * nch)); CU(c, dU.up(u, n)); CU(c, dP.alloc(n)); CU(c, dQ.alloc(n));
  CU(c, mcd_launch_test_select(dS.p, nch, dU.p, n, dP.p, dQ.p, 0));
  CU(c, dP.down(pick_out, n));
  if (path_out) CU(c, dQ.down(path_out, n));
  return MCD_OK;
}

int mcd_test_ruth(mcd_ctx* c, const float* E, const float* u, uint64_t n, float* c2_out, float* om_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  Tmp<float> dE, dU, dC, dO;
  CU(c, dE.up(E, n)); CU(c, dU.up(u, n)); CU(c, dC.alloc(n)); CU(c, dO.alloc(n));
  CU(c, mcd_launch_test_ruth(c->P, dE.p, dU.p, n, dC.p, dO.p, 0));
  CU(c, dC.down(c2_out, n)); CU(c, dO.down(om_out, n));
  return MCD_OK;
}

int mcd_test_bincoll(mcd_ctx* c, const float* E, const float* W, uint64_t n, float* c2_out, float* om_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  Tmp<float> dE, dW, dC, dO;
  CU(c, dE.up(E, n)); CU(c, dW.up(W, n)); CU(c, dC.alloc(n)); CU(c, dO.alloc(n));
  CU(c, mcd_launch_test_bincoll(c->P, dE.p, dW.p, n, dC.p, dO.p, 0));
  CU(c, dC.down(c2_out, n)); CU(c, dO.down(om_out, n));
  return MCD_OK;
}

int mcd_test_rotate(mcd_ctx* c, const float* d_in, const float* y_in, const float* c2, const float* om,
                    const float* u_phi, uint64_t n, float* d_out, float* y_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  Tmp<float> dD, dY, dC, dO, dU, oD, oY;
  CU(c, dD.up(d_in, 3 * n)); CU(c, dY.up(y_in, 3 * n)); CU(c, dC.up(c2, n)); CU(c, dO.up(om, n));
  CU(c, dU.up(u_phi, n)); CU(c, oD.alloc(3 * n)); CU(c, oY.alloc(3 * n));
  CU(c, mcd_launch_test_rotate(dD.p, dY.p, dC.p, dO.p, dU.p, n, oD.p, oY.p, 0));
  CU(c, oD.down(d_out, 3 * n)); CU(c, oY.down(y_out, 3 * n));
  return MCD_OK;
}

int mcd_test_stopping(mcd_ctx* c, int bethe, const float* E, uint64_t n, float* sp_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (bethe < 0 || bethe > 2) return fail(c, MCD_ERR_INVALID, "bad Bethe model");
  Tmp<float> dE, dS;
  CU(c, dE.up(E, n)); CU(c, dS.alloc(n));
  CU(c, mcd_launch_test_stopping(c->P, bethe, dE.p, n, dS.p, 0));
  CU(c, dS.down(sp_out, n));
  return MCD_OK;
}

int mcd_test_lookup(mcd_ctx* c, int slot, const float* E, const float* u, uint64_t n, float* W_out,
                    int32_t* iE_out, int32_t* jW_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (slot < 0 || slot >= MCD_MAX_TABLES || !c->tab[slot].cdf) return fail(c, MCD_ERR_STATE, "table slot %d not present", slot);
  Tmp<float> dE, dU, dW;
  Tmp<int> dI, dJ;
  CU(c, dE.up(E, n)); CU(c, dU.up(u, n)); CU(c, dW.alloc(n)); CU(c, dI.alloc(n)); CU(c, dJ.alloc(n));
  CU(c, mcd_launch_test_lookup(c->P, slot, dE.p, dU.p, n, dW.p, dI.p, dJ.p, 0));
  CU(c, dW.down(W_out, n)); CU(c, dI.down(iE_out, n)); CU(c, dJ.down(jW_out, n));
  return MCD_OK;
}

int mcd_test_mott(mcd_ctx* c, const float* E, const float* u, uint64_t n, float* sigma_out, float* c2_out,
                  int32_t* iE_out, int32_t* jT_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (!c->P.mott.nE) return fail(c, MCD_ERR_STATE, "no Mott table (mcd_mott_upload)");
  Tmp<float> dE, dU, dS, dC;
  Tmp<int> dI, dJ;
  CU(c, dE.up(E, n)); CU(c, dU.up(u, n)); CU(c, dS.alloc(n)); CU(c, dC.alloc(n)); CU(c, dI.alloc(n)); CU(c, dJ.alloc(n));
  CU(c, mcd_launch_test_mott(c->P, dE.p, dU.p, n, dS.p, dC.p, dI.p, dJ.p, 0));
  CU(c, dS.down(sigma_out, n)); CU(c, dC.down(c2_out, n)); CU(c, dI.down(iE_out, n)); CU(c, dJ.down(jT_out, n));
  return MCD_OK;
}

int mcd_test_step(mcd_ctx* c, const float* state_in, const float* u, uint64_t n, float* state_out, float* row_out) {
  int rc = need_ctx(c);
  if (rc) return rc;
  if (c->par.mode == MCD_MODE_DS && (!c->tab[MCD_TABLE_MOLLER].cdf || !c->tab[MCD_TABLE_GRYZ0].cdf))
    return fail(c, MCD_ERR_STATE, "DS step needs tables in slots 0 and 1");
  Tmp<float> dS, dU, oS, oR;
  CU(c, dS.up(state_in, 10 * n)); CU(c, dU.up(u, 4 * n)); CU(c, oS.alloc(10 * n)); CU(c, oR.alloc(8 * n));
  CU(c, mcd_launch_test_step(c->P, dS.p, dU.p, n, oS.p, oR.p, 0));
  CU(c, oS.down(state_out, 10 * n)); CU(c, oR.down(row_out, 8 * n));
  return MCD_OK;
}

int mcd_test_philox(uint64_t seed, uint64_t el_id, uint32_t ev_idx, uint32_t raw_out[4], float u_out[4]) {
  int n = 0;
  int rc = mcd_device_count(&n);
  if (rc) return rc;
  Tmp<uint32_t> dR;
  Tmp<float> dU;
  CU(nullptr, dR.alloc(4)); CU(nullptr, dU.alloc(4));
  CU(nullptr, mcd_launch_test_philox(seed, el_id, ev_idx, dR.p, dU.p, 0));
  CU(nullptr, dR.down(raw_out, 4)); CU(nullptr, dU.down(u_out, 4));
  return MCD_OK;
}

}  // extern "C"
