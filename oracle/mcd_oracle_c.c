/* TEST INFRASTRUCTURE ONLY -- plain-C (float64) restatement of the per-electron trajectory loop of
 * elena-pascal/MC-discrete, for fast statistical reference samples and the CPU baseline of bench.py.
 * Nothing under mcdiscrete_b200/ links or loads this file.
 *
 * Parity status: PINNED -- tests/test_oracle_c.py checks it event for event against oracle/mcd_oracle.py,
 * which is itself pinned to golden vectors produced by the reference's own code (tests/golden/).
 *
 * file:line citations are relative to the reference checkout (MC/tools/...).
 * Tables (probTables.py) are passed in as arrays (built by the numpy oracle or any other builder).
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#define ORC_MAX_SHELLS 8
#define ORC_MAX_CH (ORC_MAX_SHELLS + 3)

typedef struct {
  int Z, nshell;
  double n_val, E_val, atnd, Ef, Epl, bethe_k, bohr_r;
  double Es[ORC_MAX_SHELLS], ns[ORC_MAX_SHELLS];
} orc_mat;

typedef struct {
  const float* cdf;   /* nE x (nW-1), zeros = masked (probTables.py:500-507,533) */
  const float* Es;
  const float* Ws;
  int nE, nW;
} orc_tab;

typedef struct {
  double E0, Emin, tilt_deg, thickness, Wc;
  long maxScatt;
  int mode;  /* 0 DS, 1 cont */
  int bethe; /* 0 classical, 1 JL, 2 explicit */
  unsigned long long seed;
  int elastic; /* sigma evaluated for the elastic channel: 0 ruther_sigma, 1 ruther_sigma_wDefl, 2 ruther_N_sigma,
                  3 ruther_N_sigma_wDefl (the label -> function mapping of scattering.py:483-491 / :781-789 is done
                  by the caller, oracle/mcd_oracle.py::elastic_sigma); 4 = 'Mott' (DS only): sigma and polar angle from
                  the table below (scattering.py:493, :588-591, :96-122) */
  int mott_nE, mott_nT;
  const double* mott_Es;    /* nE ascending (extFunctions.py:93-97) */
  const double* mott_sig;   /* nE, cm^2 */
  const double* mott_cdf;   /* nE x nT (probTable, extFunctions.py:99-109) */
  const double* mott_theta; /* nT, degrees */
} orc_par;

static const double PI_EFOUR = 6.51408491409531e-14; /* parameters.py:7 */

/* ---- Philox4x32-10 (same stream definition as the CUDA path) -------------------------------- */
static void philox(uint32_t c[4], uint32_t k0, uint32_t k1) {
  for (int r = 0; r < 10; ++r) {
    uint64_t p0 = (uint64_t)0xD2511F53u * c[0], p1 = (uint64_t)0xCD9E8D57u * c[2];
    uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0, n1 = (uint32_t)p1;
    uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1, n3 = (uint32_t)p0;
    c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
}
static double u01(uint32_t x) { return ((double)(x >> 9) + 0.5) * (1.0 / 8388608.0); }

/* ---- cross sections (crossSections.py) --------------------------------------------------------- */
static double alpha_(double E, double Z) { return 3.4 * pow(Z, 0.666667) / E; }                 /* :14-27 */
static double ruther_sigma(double E, double Z) {                                              /* :37-58 */
  double a = alpha_(E, Z), Ek = E * 1e-3, r = (Ek + 511.) / (Ek + 1024.);
  return 5.21e-21 * (Z * Z / (Ek * Ek)) * (r * r) * (4. * M_PI) / (a * (1. + a));
}
static double elastic_sigma(int kind, double E, double Z) {
  if (kind == 0) return ruther_sigma(E, Z);
  if (kind == 1) return ruther_sigma(E, Z) / (Z * Z) * (Z * (Z + 1.));                       /* :63-84 */
  double b = 0.48 * (5.43 * pow(Z, 2. / 3.) / E);                                             /* :89-143 */
  return PI_EFOUR * (kind == 2 ? Z * Z : (Z + 1.) * Z) / (4 * b * (1 + b) * (E * E));
}
static int bisect_left_d(const double* a, int n, double x) {       /* bisect.bisect_left on a float64 array */
  int lo = 0, hi = n;
  while (lo < hi) {
    int mid = (lo + hi) / 2;
    if (a[mid] < x) lo = mid + 1; else hi = mid;
  }
  return lo;
}
static double moller_sigma(double E, double Wc, double nfree) {                               /* :183-208 */
  if (E <= Wc) return 0.;
  double eps = Wc * 1. / E;
  return nfree * PI_EFOUR * (1. / (E * E)) * (1. / eps - 1. / (1. - eps) + log(eps / (1. - eps)));
}
static double gryz_sigma(double E, double Esi, double nsi) {                                  /* :213-242 */
  if (E <= Esi) return 0.;
  double U = E * 1. / Esi;
  return nsi * PI_EFOUR * pow((U - 1.) / (U + 1.), 1.5) *
         (1. + (2. * (1. - (1. / (2. * U))) / 3. * log(2.7 + pow(U - 1., 0.5)))) / (E * Esi);
}
static double quinn_sigma(double E, double Epl, double Ef, double n, double a0) {             /* :247-284 */
  double E_Ef = E * 1. / Ef, Epl_Ef = Epl * 1. / Ef;
  double tot = Epl * log((pow(1. + Epl_Ef, 0.5) - 1.) / (pow(E_Ef, 0.5) - pow(E_Ef - Epl_Ef, 0.5))) / (2. * a0 * E);
  return (tot / n) * 1e4;
}
static double mean_J(double Z) { return Z >= 13 ? 9.76 * Z + 58.5 / pow(Z, 0.19) : 11.5 * Z; } /* stoppingPowers.py:134 */
static double stopping(const orc_mat* m, int model, double E) {                   /* stoppingPowers.py:112-215 */
  double Z = m->Z, J = mean_J(Z);
  if (model == 0) return 2. * PI_EFOUR * m->atnd * (Z / E) * log(1.166 * E / J) * 1e-14;
  if (model == 1) return 2. * PI_EFOUR * m->atnd * (Z / E) * log(1.166 * (E + m->bethe_k * J) / J) * 1e-14;
  double s = m->n_val * log(E * 1. / m->E_val);
  for (int i = 0; i < m->nshell; ++i) s = s + m->ns[i] * log(E * 1. / m->Es[i]);
  return (2. * PI_EFOUR * m->atnd / E) * 1e-14 * s;
}

/* pickFromSigmas (scattering.py:39-59): stable descending sort, cumsum/sum, bisect_left */
static int pick_from_sigmas(const double* sig, int n, double u) {
  double v[ORC_MAX_CH];
  int id[ORC_MAX_CH];
  for (int i = 0; i < n; ++i) {
    int j = i;
    while (j > 0 && v[j - 1] < sig[i]) { v[j] = v[j - 1]; id[j] = id[j - 1]; --j; }
    v[j] = sig[i]; id[j] = i;
  }
  double tot = 0.;
  for (int i = 0; i < n; ++i) tot += v[i];
  double cum = 0.;
  for (int i = 0; i < n; ++i) {
    cum += v[i];
    if (!(cum / tot < u)) return id[i];
  }
  return id[n - 1];
}

/* pickTable (scattering.py:63-94) with python-bisect probe order and masked-as-not-less semantics */
static double pick_table(const orc_tab* t, double E, double u) {
  int lo = 0, hi = t->nE;
  while (lo < hi) { int mid = (lo + hi) / 2; if ((double)t->Es[mid] < E) lo = mid + 1; else hi = mid; }
  const float* row = t->cdf + (size_t)lo * (size_t)(t->nW - 1);
  int a = 0, b = t->nW - 1;
  while (a < b) {
    int mid = (a + b) / 2;
    if (row[mid] != 0.0f && (double)row[mid] < u) a = mid + 1; else b = mid;
  }
  return (double)t->Ws[a];
}

/* update_direction (electron.py:87-99) + newdir (rotation.py:41-53), numpy-quaternion arithmetic */
static void update_direction(double d[3], double y[3], double c2, double halfPhi) {
  double sT = pow(1. - c2, 0.5), cT = pow(c2, 0.5), sP = sin(halfPhi), cP = pow(1. - pow(sP, 2.), 0.5);
  double pw = cT, pv[3] = {y[0] * sT, y[1] * sT, y[2] * sT};
  double aw = cP, av[3] = {d[0] * sP, d[1] * sP, d[2] * sP};
  double w = aw * pw - (av[0] * pv[0] + av[1] * pv[1] + av[2] * pv[2]);
  double cx = av[1] * pv[2] - av[2] * pv[1], cy = av[2] * pv[0] - av[0] * pv[2], cz = av[0] * pv[1] - av[1] * pv[0];
  double x = aw * pv[0] + pw * av[0] + cx, yy = aw * pv[1] + pw * av[1] + cy, z = aw * pv[2] + pw * av[2] + cz;
  double n = w * w + x * x + yy * yy + z * z;
  double m[3][3] = {{1 - 2 * (yy * yy + z * z) / n, 2 * (x * yy - z * w) / n, 2 * (x * z + yy * w) / n},
                    {2 * (x * yy + z * w) / n, 1 - 2 * (x * x + z * z) / n, 2 * (yy * z - x * w) / n},
                    {2 * (x * z - yy * w) / n, 2 * (yy * z + x * w) / n, 1 - 2 * (x * x + yy * yy) / n}};
  double nd[3], ny[3];
  for (int i = 0; i < 3; ++i) {
    nd[i] = m[i][0] * d[0] + m[i][1] * d[1] + m[i][2] * d[2];
    ny[i] = m[i][0] * y[0] + m[i][1] * y[1] + m[i][2] * y[2];
  }
  double a = sqrt(nd[0] * nd[0] + nd[1] * nd[1] + nd[2] * nd[2]);
  double b = sqrt(ny[0] * ny[0] + ny[1] * ny[1] + ny[2] * ny[2]);
  for (int i = 0; i < 3; ++i) { d[i] = nd[i] / a; y[i] = ny[i] / b; }
}

/* els row: [first outcome, n_rows (1 or 2: SURVEY Q4), final_E, dx,dy,dz, last_x,last_y,last_z] (9 doubles)
 * optional per-event rows (stride 8): [channel, x, y, z, pathl, E_loss, c2, energy] for the first max_ev_rows events */
static void one_electron(const orc_par* p, const orc_mat* m, const orc_tab* tm, const orc_tab* tg, uint64_t id,
                         double* el, int* nev_out, double* ev_rows, long max_ev_rows, long* ev_cursor) {
  double E = p->E0, xyz[3] = {0., 0., 0.}, y[3] = {0., 1., 0.};
  double t = p->tilt_deg * (M_PI / 180.0);              /* multiScatter.py:97-100 (np.radians) */
  double d[3] = {-sin(t), 0., cos(t)};
  int outcome = -1, nrows = 0;
  long nev = 0;
  const int nch = m->nshell + 3;
  while (outcome < 0) {                               /* singleScatter.py:15 / :210 */
    if (nev + 1 > p->maxScatt) {                      /* :19-23 */
      outcome = 4; nrows = 1;
      el[2] = E; memcpy(el + 3, d, 24); memcpy(el + 6, xyz, 24);
      break;
    }
    uint32_t c[4] = {(uint32_t)id, (uint32_t)(id >> 32), (uint32_t)nev, 0u};
    philox(c, (uint32_t)p->seed, (uint32_t)(p->seed >> 32));
    double u0 = u01(c[0]), u1 = u01(c[1]), u2 = u01(c[2]), u3 = u01(c[3]);
    double W, c2, halfPhi, pathl;
    int ch = 0;
    double pos0[3] = {xyz[0], xyz[1], xyz[2]};
    if (p->mode == 0) {
      double sig[ORC_MAX_CH], tot = 0.;
      sig[0] = (p->elastic == 4) ? p->mott_sig[bisect_left_d(p->mott_Es, p->mott_nE, E)]    /* scattering.py:493 */
                                 : elastic_sigma(p->elastic, E, m->Z);
      sig[1] = moller_sigma(E, p->Wc, m->n_val);
      for (int i = 0; i < m->nshell; ++i) sig[2 + i] = gryz_sigma(E, m->Es[i], m->ns[i]);
      sig[nch - 1] = quinn_sigma(E, m->Epl, m->Ef, m->atnd, m->bohr_r);
      for (int i = 0; i < nch; ++i) tot += sig[i];    /* sum(self.sigma.values()) scattering.py:519 */
      double mfp = 1. / (m->atnd * tot) * 1e14;
      ch = pick_from_sigmas(sig, nch, u0);
      pathl = -mfp * log(u1);
    } else {
      double mfp = 1. / (m->atnd * elastic_sigma(p->elastic, E, m->Z)) * 1e14;
      pathl = -mfp * log(u0);
    }
    /* update_xyz (electron.py:64-84) */
    double nz = xyz[2] + pathl * d[2];
    if (nz <= 0.) {
      outcome = 1; nrows = 1; el[2] = E; memcpy(el + 3, d, 24); memcpy(el + 6, xyz, 24);
    } else if (nz >= p->thickness) {
      outcome = 2; nrows = 1; el[2] = E; memcpy(el + 3, d, 24); memcpy(el + 6, xyz, 24);
    } else {
      xyz[0] += pathl * d[0]; xyz[1] += pathl * d[1]; xyz[2] = nz;
    }
    if (p->mode == 0) {
      if (ch == 0 && p->elastic == 4) {               /* Mott: scattering.py:588-591, pickMottTable :96-122 */
        W = 0.;
        const double* row = p->mott_cdf + (size_t)bisect_left_d(p->mott_Es, p->mott_nE, E) * (size_t)p->mott_nT;
        double theta = p->mott_theta[bisect_left_d(row, p->mott_nT, u2)];
        double ch2 = cos(theta * 0.5 * (M_PI / 180.0));        /* np.radians */
        c2 = ch2 * ch2;
      } else if (ch == 0) {                           /* scattering.py:575-577, :125-152 */
        W = 0.;
        double a = alpha_(E, m->Z);
        c2 = 1. - (a * u2 / (1. + a - u2));
      } else if (ch == nch - 1) {                     /* Quinn :643-647 */
        W = m->Epl; c2 = 1.;
      } else {                                        /* Moller :602-609 / Gryz :621-631 (table [0], Q1) */
        W = pick_table(ch == 1 ? tm : tg, E, u2);
        c2 = (W <= (E - m->Ef) * 0.5) ? 0.5 * (sqrt(1 - (W / E)) + 1) : NAN;   /* :156-179 */
      }
      halfPhi = M_PI * u3;
    } else {
      W = pathl * stopping(m, p->bethe, E);           /* scattering.py:820 */
      double a = alpha_(E, m->Z);
      c2 = 1. - (a * u1 / (1. + a - u1));
      halfPhi = M_PI * u2;
    }
    /* update_energy (electron.py:44-59) */
    double En = E - W;
    if (En <= p->Emin) {
      if (outcome < 0) { outcome = 0; nrows = 1; el[2] = E; memcpy(el + 3, d, 24); memcpy(el + 6, xyz, 24); }
      else nrows = 2;                                  /* second 'abs' row (Q4) */
    }
    E = En;
    update_direction(d, y, c2, halfPhi);
    if (ev_rows && *ev_cursor < max_ev_rows) {
      double* r = ev_rows + 8 * (*ev_cursor);
      r[0] = ch; r[1] = pos0[0]; r[2] = pos0[1]; r[3] = pos0[2]; r[4] = pathl; r[5] = W; r[6] = c2; r[7] = E;
      ++*ev_cursor;
    }
    ++nev;
  }
  el[0] = outcome;
  el[1] = nrows;
  *nev_out = (int)nev;
}

/* ---- driver: pthreads pull blocks of 16 electrons from a shared counter ------------------------------- */
#include <pthread.h>

typedef struct {
  const orc_par* p; const orc_mat* m; const orc_tab* tm; const orc_tab* tg;
  uint64_t first, count;
  double* els; int* nev;
  long long* next;
} orc_job;

static void* worker(void* arg) {
  orc_job* j = (orc_job*)arg;
  for (;;) {
    long long b = __atomic_fetch_add(j->next, 16, __ATOMIC_RELAXED);
    if (b >= (long long)j->count) break;
    long long e = b + 16 < (long long)j->count ? b + 16 : (long long)j->count;
    for (long long i = b; i < e; ++i) {
      long dummy = 0;
      one_electron(j->p, j->m, j->tm, j->tg, j->first + (uint64_t)i, j->els + 9 * i, j->nev + i, 0, 0, &dummy);
    }
  }
  return 0;
}

/* Run electrons [first, first+count).  els: count x 9 doubles; nev: count ints.
 * ev_rows (optional, forces a single thread) receives the per-event rows of the first electrons. */
int orc_run(const orc_par* p, const orc_mat* m, const orc_tab* tm, const orc_tab* tg, uint64_t first,
            uint64_t count, double* els, int* nev, int nthreads, double* ev_rows, long max_ev_rows) {
  if (p->mode == 0 && (!tm || !tg)) return -1;
  if (p->elastic == 4 && (p->mode != 0 || !p->mott_Es || !p->mott_sig || !p->mott_cdf || !p->mott_theta)) return -2;
  if (ev_rows) {
    long cursor = 0;
    for (uint64_t i = 0; i < count; ++i)
      one_electron(p, m, tm, tg, first + i, els + 9 * i, nev + i, ev_rows, max_ev_rows, &cursor);
    return 0;
  }
  if (nthreads < 1) nthreads = 1;
  if (nthreads > 1024) nthreads = 1024;
  long long next = 0;
  orc_job job = {p, m, tm, tg, first, count, els, nev, &next};
  pthread_t th[1024];
  int started = 0;
  for (int t = 1; t < nthreads; ++t)
    if (pthread_create(&th[started], 0, worker, &job) == 0) ++started;
  worker(&job);
  for (int t = 0; t < started; ++t) pthread_join(th[t], 0);
  return 0;
}
