// mcd_unit.cu -- K5: per-formula / per-step kernels fed injected uniforms (one thread per vector).
// They call the SAME device functions as the trajectory kernels (mcd_device.cuh), so parity of these
// hooks against the oracle is parity of the hot path's arithmetic.
#include "mcd_internal.h"

namespace {

constexpr int kT = 128;

template <int S>
__device__ void sigmas_insertion(const DevParams& P, float E, float* out /* S+3 */, float* tot) {
  const DevMat& m = P.m;
  const float invE = f_rcp(E);
  out[0] = P.mott.nE ? __ldg(P.mott.s + mott_row(P.mott, E)) : s_ruth(m, E);     // scattering.py:493
  out[1] = s_moller(m, E, invE);
  for (int k = 0; k < S; ++k) out[2 + m.gShell[k]] = s_gryz(E, invE, m.gE[k], m.gInvE[k], m.gC[k]);
  out[2 + S] = s_quinn(m, E, invE);
  // same association as the trajectory kernel: slot order Ruth, Quinn, Moller, Gryz slots
  float t = out[0] + out[2 + S];
  t += out[1];
  for (int k = 0; k < S; ++k) t += out[2 + m.gShell[k]];
  *tot = t;
}

__global__ void k_sigmas(const __grid_constant__ DevParams P, const float* E, uint64_t n, float* sig, float* mfp) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s[MCD_MAX_CHANNELS], tot = 0.f;
  const int S = P.m.nshell;
  switch (S) {
    case 1: sigmas_insertion<1>(P, E[i], s, &tot); break;
    case 2: sigmas_insertion<2>(P, E[i], s, &tot); break;
    case 3: sigmas_insertion<3>(P, E[i], s, &tot); break;
    case 4: sigmas_insertion<4>(P, E[i], s, &tot); break;
    case 5: sigmas_insertion<5>(P, E[i], s, &tot); break;
    case 6: sigmas_insertion<6>(P, E[i], s, &tot); break;
    case 7: sigmas_insertion<7>(P, E[i], s, &tot); break;
    default: sigmas_insertion<8>(P, E[i], s, &tot); break;
  }
  for (int k = 0; k < S + 3; ++k) sig[i * (S + 3) + k] = s[k] * P.m.sig_scale_inv;
  mfp[i] = 1.0f / tot;
}

// the hot path's cumulative inverse mean free paths in slot order (per-cell quadratics when enabled)
__global__ void k_cum(const __grid_constant__ DevParams P, const float* E, uint64_t n, float* cum_out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int nch = P.m.nshell + 3;
  const float e = E[i];
  if (P.use_sig_poly) {
    const uint32_t bits = __float_as_uint(e);
    const uint32_t cell = min((bits >> P.sig_shift) - P.sig_base, (uint32_t)(P.sig_ncell - 1));
    const float x = (float)(bits & ((1u << P.sig_shift) - 1u)) * __uint_as_float((uint32_t)(127 - P.sig_shift) << 23);
    const float4* q = P.sig_poly + (size_t)cell * (size_t)nch;
    for (int k = 2; k < nch - 1; ++k) {                    // the boundaries ds_cold_slot reads
      const float4 c = __ldg(q + k);
      cum_out[i * nch + k] = fmaf(fmaf(c.z, x, c.y), x, c.x);
    }
    // the hot record: Ruth and Ruth+Quinn boundaries (linear) and the total (quadratic), finer cells
    const uint32_t hc = min((bits >> P.hot_shift) - P.hot_base, P.hot_last);
    const float hx = (float)(bits & P.hot_mask) * P.hot_scale;
    const float4 h0 = __ldg(P.sig_hot + 2 * (size_t)hc), h1 = __ldg(P.sig_hot + 2 * (size_t)hc + 1);
    cum_out[i * nch + 0] = fmaf(h0.y, hx, h0.x);
    cum_out[i * nch + 1] = fmaf(h0.w, hx, h0.z);
    cum_out[i * nch + nch - 1] = fmaf(fmaf(h1.z, hx, h1.y), hx, h1.x);
  } else {
    const DevMat& m = P.m;
    const float invE = f_rcp(e);
    float c = s_ruth(m, e);
    cum_out[i * nch + 0] = c;
    c += s_quinn(m, e, invE);
    cum_out[i * nch + 1] = c;
    c += s_moller(m, e, invE);
    cum_out[i * nch + 2] = c;
    for (int k = 0; k < m.nshell; ++k) {
      c += s_gryz(e, invE, m.gE[k], m.gInvE[k], m.gC[k]);
      cum_out[i * nch + 3 + k] = c;
    }
  }
}

// pickFromSigmas on arbitrary sigmas: mirrors the trajectory kernel's decision structure -- the fast
// float32 path is taken only when the values are already in strictly descending order (zeros allowed at
// the tail) and u*total is outside the guard band of every boundary; otherwise the exact float64 path.
__global__ void k_select(const float* sigma, int nch, const float* u, uint64_t n, int* pick, int* path) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s[MCD_MAX_CHANNELS];
  for (int k = 0; k < nch; ++k) s[k] = sigma[i * nch + k];
  float cum = 0.f, tot = 0.f;
  bool ordered = true;
  for (int k = 0; k < nch; ++k) {
    tot += s[k];
    if (k > 0) ordered = ordered && ((s[k - 1] > s[k]) || (s[k] == 0.0f));
  }
  const float X = u[i] * tot, guard = 1e-5f * tot;
  int slot = 0;
  bool safe = ordered;
  for (int k = 0; k < nch - 1; ++k) {
    cum += s[k];
    slot += (cum < X) ? 1 : 0;
    safe = safe && (fabsf(cum - X) > guard);
  }
  safe = safe && (tot - X > guard);
  if (safe) {
    pick[i] = slot;
    if (path) path[i] = 0;
  } else {
    pick[i] = select_exact(s, nch, u[i]);
    if (path) path[i] = 1;
  }
}

__global__ void k_ruth(const __grid_constant__ DevParams P, const float* E, const float* u, uint64_t n, float* c2,
                       float* om) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float o = ruth_om(P.m, f_rcp(E[i]), u[i]);
  om[i] = o;
  c2[i] = 1.0f - o;
}

__global__ void k_bincoll(const __grid_constant__ DevParams P, const float* E, const float* W, uint64_t n,
                          float* c2, float* om) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (W[i] > 0.5f * (E[i] - P.m.Ef)) {            // scattering.py:177-179 -> NaN
    om[i] = __int_as_float(0x7fc00000);
    c2[i] = __int_as_float(0x7fc00000);
    return;
  }
  const float o = bincoll_om(f_rcp(E[i]), W[i]);
  om[i] = o;
  c2[i] = 1.0f - o;
}

__global__ void k_rotate(const float* d, const float* y, const float* c2, const float* om, const float* u,
                         uint64_t n, float* d_out, float* y_out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float dx = d[3 * i], dy = d[3 * i + 1], dz = d[3 * i + 2];
  float yx = y[3 * i], yy = y[3 * i + 1], yz = y[3 * i + 2];
  rotate_dir(dx, dy, dz, yx, yy, yz, c2[i], om[i], u[i]);
  d_out[3 * i] = dx; d_out[3 * i + 1] = dy; d_out[3 * i + 2] = dz;
  y_out[3 * i] = yx; y_out[3 * i + 1] = yy; y_out[3 * i + 2] = yz;
}

__global__ void k_stopping(const __grid_constant__ DevParams P, int bethe, const float* E, uint64_t n, float* sp) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  sp[i] = sp_bethe(P.m, bethe, E[i], f_rcp(E[i]));
}

__global__ void k_lookup(const __grid_constant__ DevParams P, int slot, const float* E, const float* u, uint64_t n,
                         float* W, int* iE, int* jW) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int a = 0, b = 0;
  W[i] = table_lookup(P.tab[slot], E[i], u[i], &a, &b);
  iE[i] = a;
  jW[i] = b;
}

__global__ void k_mott(const __grid_constant__ DevParams P, const float* E, const float* u, uint64_t n, float* sig,
                       float* c2, int* iE, int* jT) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const int r = mott_row(P.mott, E[i]);
  int j = 0;
  const float om = mott_om(P.mott, r, u[i], &j);
  sig[i] = __ldg(P.mott.s + r) * P.m.sig_scale_inv;
  c2[i] = 1.0f - om;
  iE[i] = r;
  jT[i] = j;
}

template <int S>
__device__ void step_one(const DevParams& P, bool on, ElState& e, float4 u, EventRow& row) {
  if (!on) return;
  if (P.mode == MCD_MODE_DS) {
    EventMid mid;
    float tot;
    ds_phase_a<S>(P, e, u, row, mid, tot, ds_hot_fetch(P, e.E));
    ds_phase_b<S>(P, e, u, row, mid, tot);
  } else {
    cont_event(P, e, u, row);
  }
  event_rotate(e, row);
}

__global__ void k_step(const __grid_constant__ DevParams P, const float* st, const float* uu, uint64_t n,
                       float* st_out, float* row_out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool on = i < n;
  ElState e;
  e.E = P.E0; e.x = e.y = e.z = 0.f; e.dx = 0.f; e.dy = 0.f; e.dz = 1.f; e.yx = 0.f; e.yy = 1.f; e.yz = 0.f;
  float4 u = make_float4(0.5f, 0.5f, 0.5f, 0.5f);
  if (on) {
    const float* s = st + 10 * i;
    e.E = s[0]; e.x = s[1]; e.y = s[2]; e.z = s[3]; e.dx = s[4]; e.dy = s[5]; e.dz = s[6];
    e.yx = s[7]; e.yy = s[8]; e.yz = s[9];
    u = make_float4(uu[4 * i], uu[4 * i + 1], uu[4 * i + 2], uu[4 * i + 3]);
  }
  EventRow row;
  switch (P.m.nshell) {
    case 1: step_one<1>(P, on, e, u, row); break;
    case 2: step_one<2>(P, on, e, u, row); break;
    case 3: step_one<3>(P, on, e, u, row); break;
    case 4: step_one<4>(P, on, e, u, row); break;
    case 5: step_one<5>(P, on, e, u, row); break;
    case 6: step_one<6>(P, on, e, u, row); break;
    case 7: step_one<7>(P, on, e, u, row); break;
    default: step_one<8>(P, on, e, u, row); break;
  }
  if (!on) return;
  float* o = st_out + 10 * i;
  o[0] = e.E; o[1] = e.x; o[2] = e.y; o[3] = e.z; o[4] = e.dx; o[5] = e.dy; o[6] = e.dz;
  o[7] = e.yx; o[8] = e.yy; o[9] = e.yz;
  float* r = row_out + 8 * i;
  r[0] = (float)slot_channel(P.m, row.slot); r[1] = row.pathl; r[2] = row.E_loss; r[3] = row.c2; r[4] = row.halfPhi;
  r[5] = row.energy; r[6] = (float)row.outcome; r[7] = (float)row.flags;
}

__global__ void k_philox(uint32_t slo, uint32_t shi, uint32_t elo, uint32_t ehi, uint32_t ev, uint32_t* raw,
                         float* u) {
  const uint4 r = philox4x32_10(make_uint4(elo, ehi, ev, 0u), make_uint2(slo, shi));
  raw[0] = r.x; raw[1] = r.y; raw[2] = r.z; raw[3] = r.w;
  u[0] = u01(r.x); u[1] = u01(r.y); u[2] = u01(r.z); u[3] = u01(r.w);
}

inline int nblk(uint64_t n) { return (int)((n + kT - 1) / kT); }

}  // namespace

#define LAUNCH(k, n, s, ...)                                   \
  do {                                                         \
    if ((n) == 0) return cudaSuccess;                          \
    k<<<nblk(n), kT, 0, s>>>(__VA_ARGS__);                     \
    return cudaGetLastError();                                 \
  } while (0)

cudaError_t mcd_launch_test_sigmas(const DevParams& P, const float* E, uint64_t n, float* sig, float* mfp,
                                   cudaStream_t s) { LAUNCH(k_sigmas, n, s, P, E, n, sig, mfp); }
cudaError_t mcd_launch_test_cum(const DevParams& P, const float* E, uint64_t n, float* cum, cudaStream_t s) {
  LAUNCH(k_cum, n, s, P, E, n, cum);
}
cudaError_t mcd_launch_test_select(const float* sigma, int nch, const float* u, uint64_t n, int* pick, int* path,
                                   cudaStream_t s) { LAUNCH(k_select, n, s, sigma, nch, u, n, pick, path); }
cudaError_t mcd_launch_test_ruth(const DevParams& P, const float* E, const float* u, uint64_t n, float* c2,
                                 float* om, cudaStream_t s) { LAUNCH(k_ruth, n, s, P, E, u, n, c2, om); }
cudaError_t mcd_launch_test_bincoll(const DevParams& P, const float* E, const float* W, uint64_t n, float* c2,
                                    float* om, cudaStream_t s) { LAUNCH(k_bincoll, n, s, P, E, W, n, c2, om); }
cudaError_t mcd_launch_test_rotate(const float* d, const float* y, const float* c2, const float* om,
                                   const float* u, uint64_t n, float* d_out, float* y_out, cudaStream_t s) {
  LAUNCH(k_rotate, n, s, d, y, c2, om, u, n, d_out, y_out);
}
cudaError_t mcd_launch_test_stopping(const DevParams& P, int bethe, const float* E, uint64_t n, float* sp,
                                     cudaStream_t s) { LAUNCH(k_stopping, n, s, P, bethe, E, n, sp); }
cudaError_t mcd_launch_test_lookup(const DevParams& P, int slot, const float* E, const float* u, uint64_t n,
                                   float* W, int* iE, int* jW, cudaStream_t s) {
  LAUNCH(k_lookup, n, s, P, slot, E, u, n, W, iE, jW);
}
cudaError_t mcd_launch_test_mott(const DevParams& P, const float* E, const float* u, uint64_t n, float* sig, float* c2,
                                 int* iE, int* jT, cudaStream_t s) { LAUNCH(k_mott, n, s, P, E, u, n, sig, c2, iE, jT); }
cudaError_t mcd_launch_test_step(const DevParams& P, const float* st, const float* u, uint64_t n, float* st_out,
                                 float* row_out, cudaStream_t s) { LAUNCH(k_step, n, s, P, st, u, n, st_out, row_out); }
cudaError_t mcd_launch_test_philox(uint64_t seed, uint64_t el, uint32_t ev, uint32_t* raw, float* u,
                                   cudaStream_t s) {
  k_philox<<<1, 1, 0, s>>>((uint32_t)seed, (uint32_t)(seed >> 32), (uint32_t)el, (uint32_t)(el >> 32), ev, raw, u);
  return cudaGetLastError();
}
