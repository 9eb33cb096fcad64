// mcd_bench.cu -- micro-kernels that MEASURE the two compute roofs of the trajectory kernels on the
// device at hand: the FP32 FMA pipe (2 flop/lane/clk) and the MUFU/XU pipe
// (transcendentals, 16 lanes/clk/SM).  Used by bench.py for the roofline denominators; not on the hot path.
#include "mcd_internal.h"

namespace {

constexpr int kIters = 4096;
constexpr int kChains = 8;

__global__ void __launch_bounds__(256) fma_chain(float* out, float a, float b) {
  float x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = (float)(threadIdx.x + i);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) x[i] = fmaf(x[i], a, b);
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;            // keep the chains alive
}

__global__ void __launch_bounds__(256) mufu_chain(float* out, float a) {
  float x[kChains];
#pragma unroll
  for (int i = 0; i < kChains; ++i) x[i] = a + 1e-3f * (float)(threadIdx.x + i);
  for (int it = 0; it < kIters; ++it) {
#pragma unroll
    for (int i = 0; i < kChains; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
  }
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < kChains; ++i) s += x[i];
  if (s == 123.456f) out[0] = s;
}

}  // namespace

extern "C" int mcd_measure_peaks(int device, float* fp32_tflops, float* mufu_tops, float* sm_mhz_est) {
  if (cudaSetDevice(device) != cudaSuccess) return MCD_ERR_NO_DEVICE;
  cudaDeviceProp pr;
  if (cudaGetDeviceProperties(&pr, device) != cudaSuccess) return MCD_ERR_CUDA;
  float* d = nullptr;
  if (cudaMalloc(&d, 16) != cudaSuccess) return MCD_ERR_CUDA;
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0);
  cudaEventCreate(&e1);
  const int blocks = pr.multiProcessorCount * 8;
  const double threads = (double)blocks * 256.0;
  float best_f = 1e30f, best_m = 1e30f;
  for (int rep = 0; rep < 6; ++rep) {
    float ms = 0.f;
    cudaEventRecord(e0);
    fma_chain<<<blocks, 256>>>(d, 0.999f, 0.001f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_f) best_f = ms;
    cudaEventRecord(e0);
    mufu_chain<<<blocks, 256>>>(d, -0.5f);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    cudaEventElapsedTime(&ms, e0, e1);
    if (rep > 0 && ms < best_m) best_m = ms;
  }
  cudaError_t err = cudaGetLastError();
  cudaEventDestroy(e0);
  cudaEventDestroy(e1);
  cudaFree(d);
  if (err != cudaSuccess) return MCD_ERR_CUDA;
  const double ops = threads * kIters * kChains;
  if (fp32_tflops) *fp32_tflops = (float)(2.0 * ops / (best_f * 1e-3) / 1e12);
  if (mufu_tops) *mufu_tops = (float)(ops / (best_m * 1e-3) / 1e12);
  if (sm_mhz_est) *sm_mhz_est = (float)(ops / (best_m * 1e-3) / (16.0 * pr.multiProcessorCount) / 1e6);
  return MCD_OK;
}
