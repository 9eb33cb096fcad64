// mcd_internal.h -- host-side declarations shared by the translation units of libmcd.so
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include "../../include/mcd.h"
#include "mcd_device.cuh"

struct TableBuildArgs {
  int kind;                 // 0 Moller, 1 Gryzinski
  const float* Es;          // device, nE
  const float* Ws;          // device, nW
  float* cdf;               // device, nE x (nW-1)
  int* nvalid;              // device, nE
  double* scratch;          // device, gridDim.x x (nW-1)
  int nE, nW;
  float Ef32;               // float32(Ef): (E - Ef)*0.5 is float32 arithmetic in the reference
  double numEl;             // n_val (Moller) or ns (Gryzinski)
  double Eb;                // shell binding energy (Gryzinski), python float in the reference
};

struct TrajLaunchInfo {
  int blocks, threads, regs, blocks_per_sm;
  int launches;             // kernel launches per mcd_run (tail hand-over passes)
};

// mcd_traj.cu
// tail: optional high-priority side stream for the hand-over passes and the histogram pass (+ two events that order
// it behind the first pass and in front of whatever follows on `stream`); null = everything on `stream`
struct TrajTailStream {
  cudaStream_t stream = nullptr;
  cudaEvent_t first_done = nullptr, tail_done = nullptr;
};
cudaError_t mcd_launch_traj(const DevParams& P, const DevOut& O, unsigned long long first_id,
                            unsigned long long count, int sm_count, cudaStream_t stream, TrajLaunchInfo* info,
                            const TrajTailStream* tail = nullptr);
cudaError_t mcd_traj_max_blocks(const DevParams& P, bool scats, int sm_count, TrajLaunchInfo* info);
cudaError_t mcd_launch_stats_merge(unsigned long long* dst, unsigned long long* src, int n, cudaStream_t stream);
int mcd_traj_uses_queue(const DevParams& P, bool scats);     // 1 if mcd_launch_traj wants the leftover queues

cudaError_t mcd_launch_order_scats(const unsigned int* nev, unsigned long long count, unsigned long long* tile_tmp,
                                   unsigned long long* offsets, unsigned long long* total_dev, const uint32_t* pool,
                                   const unsigned int* chunk_count, unsigned int nchunks, int row_words,
                                   uint32_t* sorted, unsigned long long col_stride, unsigned long long max_rows,
                                   const unsigned int* nchunks_dev, int sm_count, cudaStream_t stream);
// col_stride != 0: columnar output; max_rows: rows the output holds (targets beyond it are skipped); nchunks_dev
// (enqueued columnar regrouping only): device word holding the used chunks
unsigned long long mcd_order_tiles(unsigned long long count);

// mcd_tables.cu
cudaError_t mcd_launch_table_build(const TableBuildArgs& a, int blocks, cudaStream_t stream);
cudaError_t mcd_launch_table_nvalid(const float* cdf, int nE, int nW, int* nvalid, cudaStream_t stream);
cudaError_t mcd_launch_table_guide(const float* cdf, int nE, int nW, const int* nvalid, uint32_t* guide,
                                   cudaStream_t stream);

// mcd_unit.cu
cudaError_t mcd_launch_test_sigmas(const DevParams& P, const float* E, uint64_t n, float* sig, float* mfp,
                                   cudaStream_t s);
cudaError_t mcd_launch_test_cum(const DevParams& P, const float* E, uint64_t n, float* cum, cudaStream_t s);
cudaError_t mcd_launch_test_select(const float* sigma, int nch, const float* u, uint64_t n, int* pick, int* path,
                                   cudaStream_t s);
cudaError_t mcd_launch_test_ruth(const DevParams& P, const float* E, const float* u, uint64_t n, float* c2,
                                 float* om, cudaStream_t s);
cudaError_t mcd_launch_test_bincoll(const DevParams& P, const float* E, const float* W, uint64_t n, float* c2,
                                    float* om, cudaStream_t s);
cudaError_t mcd_launch_test_rotate(const float* d, const float* y, const float* c2, const float* om,
                                   const float* u, uint64_t n, float* d_out, float* y_out, cudaStream_t s);
cudaError_t mcd_launch_test_stopping(const DevParams& P, int bethe, const float* E, uint64_t n, float* sp,
                                     cudaStream_t s);
cudaError_t mcd_launch_test_lookup(const DevParams& P, int slot, const float* E, const float* u, uint64_t n,
                                   float* W, int* iE, int* jW, cudaStream_t s);
cudaError_t mcd_launch_test_mott(const DevParams& P, const float* E, const float* u, uint64_t n, float* sig, float* c2,
                                 int* iE, int* jT, cudaStream_t s);
cudaError_t mcd_launch_test_step(const DevParams& P, const float* st, const float* u, uint64_t n, float* st_out,
                                 float* row_out, cudaStream_t s);
cudaError_t mcd_launch_test_philox(uint64_t seed, uint64_t el, uint32_t ev, uint32_t* raw, float* u,
                                   cudaStream_t s);
