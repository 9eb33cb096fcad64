// mcd_tables.cu -- K3: build the inverse-CDF energy-loss tables on the device.
//
// Replaces probTable.computeBlock (MC/tools/probTables.py:415-462) with the integrands moller_dCS /
// gryz_dCS (MC/tools/extFunctions.py:159-219): per incident energy E (one table row)
//     f_j    = dsigma/dW(E, W_j)
//     step_j = (W_{j+1} - W_j) (f_j + f_{j+1}) / 2      masked where W_{j+1} > Wmax(E) = (E - Ef)/2
//     CDF_j  = cumsum(step)_j / cumsum(step)_last        stored as float32, masked entries = 0
// One block scans one row at a time (block-wide inclusive scan over up to ~1e6 steps, tiles of
// 256 x 4 steps with a running carry); blocks stride over rows.
//
// Arithmetic follows what the reference actually executes under numpy >= 2 with a float32 E mesh and a
// float32 MASKED W mesh: eps = W/E, epsB = Eb/E, epsB/(epsB+eps), (1+epsB)^-1.5, E^3 and the W
// differences are float32; everything downstream of the first python-scalar-on-masked-array operation
// is float64 (see oracle/mcd_oracle.py::ProbTable._rows_integrand, pinned bit-exactly to the reference).
// The scan is float64, so the only difference to the reference's sequential float64 cumsum is
// reassociation at the 1e-16 level, invisible after the float32 store except on exact rounding ties.
#include "mcd_internal.h"

namespace {

constexpr int kTB = 256;
constexpr int kItems = 4;
constexpr double kPiEFour = 6.51408491409531e-14;   // parameters.py:7

__device__ __forceinline__ double integrand(const TableBuildArgs& a, float E32, float W32, double rowc) {
  const float eps32 = __fdiv_rn(W32, E32);
  const double eps = (double)eps32;
  if (a.kind == 0) {                                            // extFunctions.py:178-182
    if (!(eps32 < 1.0f)) return 0.0;
    const double ome = 1.0 - eps;
    return rowc * (1.0 / (eps * eps) + (1.0 / (ome * ome)) - (1.0 / (eps * ome)));
  }
  // extFunctions.py:213-218
  const float epsB32 = __fdiv_rn((float)a.Eb, E32);
  if (!((eps32 < 1.0f) && (epsB32 < 1.0f))) return 0.0;
  const double expo = (double)__fdiv_rn(epsB32, __fadd_rn(epsB32, eps32));
  const double ome = 1.0 - eps;
  const double epsB = (double)epsB32;
  const double w = (double)W32;
  return (rowc * eps) * pow(ome, expo) *
         ((double)__fsub_rn(1.0f, epsB32) +
          (double)__fmul_rn(4.0f, epsB32) * log(2.7 + sqrt(ome / epsB)) / (3.0 * eps)) / (w * w * w);
}

__device__ __forceinline__ double warp_incl_scan(double v, unsigned lane) {
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const double o = __shfl_up_sync(MCD_FULL_MASK, v, d);
    if (lane >= (unsigned)d) v += o;
  }
  return v;
}

__global__ void __launch_bounds__(kTB) table_build_kernel(const TableBuildArgs a) {
  __shared__ double s_warp[kTB / 32];
  __shared__ double s_carry;
  __shared__ int s_nv;
  const int nsteps = a.nW - 1;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  double* scratch = a.scratch + (size_t)blockIdx.x * (size_t)nsteps;

  for (int row = blockIdx.x; row < a.nE; row += gridDim.x) {
    const float E32 = a.Es[row];
    const float wmax32 = __fmul_rn(__fsub_rn(E32, a.Ef32), 0.5f);          // probTables.py:17-41
    // row constant: numEl pi e^4 (x (1+epsB)^-1.5) / E^3 -- cancels in the normalisation, kept for fidelity
    double rowc = a.numEl * kPiEFour;
    if (a.kind == 0) rowc /= (double)powf(E32, 3.0f);
    else rowc *= (double)powf(__fadd_rn(1.0f, __fdiv_rn((float)a.Eb, E32)), -1.5f);
    if (threadIdx.x == 0) {
      // number of W entries <= wmax (Ws ascending) -> valid steps = that - 1   (probTables.py:439-454)
      int lo = 0, hi = a.nW;
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (a.Ws[mid] > wmax32) hi = mid; else lo = mid + 1;
      }
      s_nv = max(lo - 1, 0);
      s_carry = 0.0;
    }
    __syncthreads();
    const int nv = s_nv;

    for (int tile = 0; tile < nv; tile += kTB * kItems) {
      const int j0 = tile + threadIdx.x * kItems;
      double loc[kItems];
      double fprev = 0.0;
      float wprev = 0.f;
      if (j0 < nv) {
        wprev = a.Ws[j0];
        fprev = integrand(a, E32, wprev, rowc);
      }
      double run = 0.0;
#pragma unroll
      for (int k = 0; k < kItems; ++k) {
        const int j = j0 + k;
        double st = 0.0;
        if (j < nv) {
          const float wn = a.Ws[j + 1];
          const double fn = integrand(a, E32, wn, rowc);
          st = (double)__fsub_rn(wn, wprev) * (fn + fprev) / 2.0;            // np.trapz on the 2-stack
          wprev = wn;
          fprev = fn;
        }
        run += st;
        loc[k] = run;
      }
      const double incl = warp_incl_scan(run, lane);
      if (lane == 31) s_warp[warp] = incl;
      __syncthreads();
      double woff = 0.0;
#pragma unroll
      for (int w = 0; w < kTB / 32; ++w) woff += (w < (int)warp) ? s_warp[w] : 0.0;
      const double base = s_carry + woff + (incl - run);
#pragma unroll
      for (int k = 0; k < kItems; ++k)
        if (j0 + k < nv) scratch[j0 + k] = base + loc[k];
      __syncthreads();
      if (threadIdx.x == kTB - 1) s_carry = base + run;
      __syncthreads();
    }
    const double total = s_carry;
    float* out = a.cdf + (size_t)row * (size_t)nsteps;
    for (int j = threadIdx.x; j < nsteps; j += kTB)
      out[j] = (j < nv) ? (float)(scratch[j] / total) : 0.0f;
    if (threadIdx.x == 0) a.nvalid[row] = nv;
    __syncthreads();
  }
}

// for uploaded tables: nvalid[row] = number of leading non-zero entries (readFromMemory masks zeros,
// probTables.py:533)
__global__ void table_nvalid_kernel(const float* __restrict__ cdf, int nE, int nW, int* __restrict__ nvalid) {
  const int row = blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= nE) return;
  const float* r = cdf + (size_t)row * (size_t)(nW - 1);
  int lo = 0, hi = nW - 1;          // zeros form the tail: first index with r[i] == 0
  while (lo < hi) {
    const int mid = (lo + hi) >> 1;
    if (r[mid] != 0.0f) lo = mid + 1; else hi = mid;
  }
  nvalid[row] = lo;
}

// guide[row][g] = bisect_left(row[0:nvalid], g / MCD_GUIDE) for g = 0..MCD_GUIDE: brackets every lookup
// (table_lookup in mcd_device.cuh) without changing its result.
__global__ void table_guide_kernel(const float* __restrict__ cdf, int nE, int nW, const int* __restrict__ nvalid,
                                   uint32_t* __restrict__ guide) {
  const int row = blockIdx.x;
  const float* r = cdf + (size_t)row * (size_t)(nW - 1);
  const int nv = nvalid[row];
  for (int g = threadIdx.x; g <= MCD_GUIDE; g += blockDim.x) {
    const float t = (float)g * (1.0f / (float)MCD_GUIDE);
    int lo = 0, hi = nv;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (r[mid] < t) lo = mid + 1; else hi = mid;
    }
    guide[(size_t)row * (MCD_GUIDE + 1) + g] = (uint32_t)lo;
  }
}

}  // namespace

cudaError_t mcd_launch_table_guide(const float* cdf, int nE, int nW, const int* nvalid, uint32_t* guide,
                                   cudaStream_t stream) {
  table_guide_kernel<<<nE, 256, 0, stream>>>(cdf, nE, nW, nvalid, guide);
  return cudaGetLastError();
}

cudaError_t mcd_launch_table_build(const TableBuildArgs& a, int blocks, cudaStream_t stream) {
  table_build_kernel<<<blocks, kTB, 0, stream>>>(a);
  return cudaGetLastError();
}

cudaError_t mcd_launch_table_nvalid(const float* cdf, int nE, int nW, int* nvalid, cudaStream_t stream) {
  table_nvalid_kernel<<<(nE + 127) / 128, 128, 0, stream>>>(cdf, nE, nW, nvalid);
  return cudaGetLastError();
}
