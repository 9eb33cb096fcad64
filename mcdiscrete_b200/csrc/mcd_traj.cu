// mcd_traj.cu -- K1/K2: persistent one-thread-per-electron trajectory kernels (sm_100a).
//
// Replaces the job loop of multiTraj_DS / multiTraj_cont (MC/tools/multiScatter.py:66-116, 171-229) and
// the while-loops of trajectory_DS / trajectory_cont_cl (MC/tools/singleScatter.py:10-41, 205-239).
//
//  * grid = resident blocks only (SMs x occupancy); a lane whose electron ended takes the next id from
//    a global ticket (one warp-aggregated atomic per refill) so warps stay full until ids run out;
//  * RNG = Philox4x32-10, key = seed, counter = (electron id, event index): results do not depend on
//    which lane / block / GPU runs an electron;
//  * outputs: one 32-byte els row per electron at index id - first_id; scats rows are appended by each
//    warp into chunks of MCD_CHUNK_ROWS rows that it takes from a global pool (ballot + popc give every
//    lane its slot; the only atomic is one bump-allocation per chunk), so rows of one event are written
//    contiguously and there is no per-event contention; backscatter histograms and counters accumulate
//    in shared memory and are flushed once per block.
#include <stdlib.h>

#include "mcd_internal.h"

namespace {

#ifndef MCD_THREADS
#define MCD_THREADS 192
#endif
constexpr int kThreads = MCD_THREADS;
#ifndef MCD_MIN_BLOCKS
#define MCD_MIN_BLOCKS 6          // 6 x 192 threads = 36 warps per SM at 56 registers (measured best: r1h notes)
#endif
constexpr int kCntOutcome = 0;       // 5
constexpr int kCntEvents = 5;
constexpr int kCntDone = 6;
constexpr int kCntDouble = 7;
constexpr int kCntNan = 8;
constexpr int kCntExact = 9;
constexpr int kCntChan = 10;         // MCD_MAX_CHANNELS
constexpr int kCntTotal = kCntChan + MCD_MAX_CHANNELS;

__device__ __forceinline__ unsigned lanemask_lt() {
  unsigned m;
  asm("mov.u32 %0, %%lanemask_lt;" : "=r"(m));
  return m;
}

__device__ __forceinline__ void store_scat_row(uint32_t* __restrict__ dst, uint32_t mask, uint32_t rel_id,
                                               uint32_t ev, int channel, const EventRow& r, const ElState& e) {
  int w = 0;
  dst[w++] = rel_id;
  dst[w++] = (ev & 0xffffffu) | ((uint32_t)channel << 24);
  if (mask & MCD_SC_POSITION) {
    dst[w++] = __float_as_uint(r.px); dst[w++] = __float_as_uint(r.py); dst[w++] = __float_as_uint(r.pz);
  }
  if (mask & MCD_SC_PATHL) dst[w++] = __float_as_uint(r.pathl);
  if (mask & MCD_SC_E) dst[w++] = __float_as_uint(r.E);
  if (mask & MCD_SC_E_LOSS) dst[w++] = __float_as_uint(r.E_loss);
  if (mask & MCD_SC_POL_ANGLE) dst[w++] = __float_as_uint(r.c2);
  if (mask & MCD_SC_AZ_ANGLE) dst[w++] = __float_as_uint(r.halfPhi);
  if (mask & MCD_SC_ENERGY) dst[w++] = __float_as_uint(r.energy);
  if (mask & MCD_SC_DIRECTION) {
    dst[w++] = __float_as_uint(e.dx); dst[w++] = __float_as_uint(e.dy); dst[w++] = __float_as_uint(e.dz);
  }
}

template <int S, int MODE, bool SCATS>
__global__ void __launch_bounds__(kThreads, MCD_MIN_BLOCKS)
traj_kernel(const __grid_constant__ DevParams P, const DevOut O, const unsigned long long first_id,
            const unsigned long long count) {
  __shared__ unsigned int s_hist[MCD_HIST_TOTAL];
  __shared__ unsigned long long s_cnt[kCntTotal];
  __shared__ unsigned long long s_rows[2];      // stored, dropped

  for (int i = threadIdx.x; i < MCD_HIST_TOTAL; i += kThreads) s_hist[i] = 0u;
  for (int i = threadIdx.x; i < kCntTotal; i += kThreads) s_cnt[i] = 0ull;
  if (threadIdx.x < 2) s_rows[threadIdx.x] = 0ull;
  __syncthreads();

  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = lanemask_lt();
  constexpr bool want_scats = SCATS;            // compile time: the no-scats kernel carries no row fields
  const bool want_hist = (P.flags & MCD_FLAG_HISTOGRAMS) != 0u;
  const bool want_chan = (P.flags & MCD_FLAG_CHANNEL_COUNTS) != 0u;
  int chunk = -1;                                // this warp's current chunk (warp-uniform)
  unsigned int used = 0u;                        // rows already stored in it
  unsigned int w_stored = 0u, w_dropped = 0u;    // warp-uniform tallies

  bool alive = false, exhausted = false;
  ElState e;
  uint32_t rel_id = 0u;                          // id relative to first_id (a run holds < 2^32 electrons)
  uint32_t ev = 0u;                              // events executed by the current electron

  while (true) {
    // ---- refill dead lanes (multiScatter.py:105-107: a fresh electron at pos0 / dir0) ----------
    const unsigned need = __ballot_sync(MCD_FULL_MASK, !alive && !exhausted);
    if (need) {
      const int leader = __ffs(need) - 1;
      unsigned long long base = 0ull;
      if ((int)lane == leader) base = atomicAdd(O.ticket, (unsigned long long)__popc(need));
      base = __shfl_sync(MCD_FULL_MASK, base, leader);
      if (!alive && !exhausted) {
        const unsigned long long rid = base + (unsigned long long)__popc(need & lt);
        rel_id = (uint32_t)rid;
        if (rid < count) {
          alive = true;
          ev = 0u;
          e.E = P.E0;
          e.x = 0.f; e.y = 0.f; e.z = 0.f;
          e.dx = P.d0x; e.dy = P.d0y; e.dz = P.d0z;
          e.yx = 0.f; e.yy = 1.f; e.yz = 0.f;                    // electron.py:38
        } else {
          exhausted = true;
        }
      }
    }
    if (!__any_sync(MCD_FULL_MASK, alive)) break;

    // ---- one loop iteration of trajectory_DS / trajectory_cont_cl ---------------------------
    // an alive lane either executes one event or, past maxScatt, retires as 'long' (singleScatter.py:19-23)
    const bool has_row = alive && (ev < (uint32_t)P.maxScatt);
    int outcome = (alive && !has_row) ? MCD_OUT_LONG : -1;
    EventRow row;
    row.flags = 0;
    row.slot = 0;
    row.E = e.E;
    if (has_row) {
      const float4 u = event_uniforms(P, first_id + rel_id, ev);
      if (MODE == MCD_MODE_DS) {
        EventMid mid;
        ds_phase_a<S>(P, e, u, row, mid);
        ds_phase_b(P, e, u, row, mid);
      } else {
        cont_event(P, e, u, row);
      }
      outcome = row.outcome;
    }

    // ---- retire: saveOutcomes row (electron.py:105-113), histograms, counters.  Taken BEFORE the rotation of this
    // event: final_E = energy before the event's loss, final_dir = direction before its rotation, last_pos =
    // position as update_xyz left it -----------------------------------------------------------------
    if (outcome >= 0) {
      const bool dbl = (row.flags & MCD_EVF_DOUBLE) != 0;
      const uint32_t nev = ev + (has_row ? 1u : 0u);
      if (want_scats) O.nev[rel_id] = nev;
      float4* dst = reinterpret_cast<float4*>(O.els + rel_id);
      dst[0] = make_float4(__uint_as_float((uint32_t)outcome | (dbl ? 0x100u : 0u)), row.E, e.dx, e.dy);
      dst[1] = make_float4(e.dz, e.x, e.y, e.z);
      atomicAdd(&s_cnt[kCntOutcome + outcome], 1ull);
      atomicAdd(&s_cnt[kCntDone], 1ull);
      atomicAdd(&s_cnt[kCntEvents], (unsigned long long)nev);
      if (dbl) atomicAdd(&s_cnt[kCntDouble], 1ull);
      if (want_hist && outcome == MCD_OUT_BKS) {
        int b = (int)((row.E - P.hist.e_lo) * P.hist.e_scale);
        b = max(0, min(b, MCD_HIST_E_BINS - 1));
        atomicAdd(&s_hist[b], 1u);
        const float pol = acosf(fminf(1.0f, fmaxf(-1.0f, -e.dz))) * 57.29577951308232f;       // from -z normal
        b = max(0, min((int)(pol * 2.0f), MCD_HIST_POLAR_BINS - 1));
        atomicAdd(&s_hist[MCD_HIST_E_BINS + b], 1u);
        const float az = atan2f(e.dy, e.dx) * 57.29577951308232f + 180.0f;
        b = max(0, min((int)az, MCD_HIST_AZ_BINS - 1));
        atomicAdd(&s_hist[MCD_HIST_E_BINS + MCD_HIST_POLAR_BINS + b], 1u);
        b = max(0, min((int)(e.z * P.hist.depth_scale), MCD_HIST_DEPTH_BINS - 1));
        atomicAdd(&s_hist[MCD_HIST_E_BINS + MCD_HIST_POLAR_BINS + MCD_HIST_AZ_BINS + b], 1u);
      }
      alive = false;
    }
    // update_direction of this event (the exit event still rotates and, if requested, records its direction)
    if (has_row) event_rotate(e, row);

    // ---- scats append: warp-local slots in the warp's chunk; a new chunk when it is full --------
    if (want_scats) {
      const unsigned m = __ballot_sync(MCD_FULL_MASK, has_row);
      if (m) {
        const unsigned k = (unsigned)__popc(m);
        const unsigned rank = (unsigned)__popc(m & lt);
        long long slot_row = -1;                                   // row index in the pool, -1 = dropped
        if (chunk >= 0 && used + k <= MCD_CHUNK_ROWS) {
          slot_row = (long long)chunk * MCD_CHUNK_ROWS + used + rank;
          used += k;
          w_stored += k;
        } else {
          const unsigned rem = (chunk >= 0) ? (MCD_CHUNK_ROWS - used) : 0u;
          unsigned int nc = 0u;
          if (lane == 0) {
            if (chunk >= 0) O.chunk_count[chunk] = MCD_CHUNK_ROWS;   // old chunk gets filled below
            nc = atomicAdd(O.chunk_next, 1u);
          }
          nc = __shfl_sync(MCD_FULL_MASK, nc, 0);
          const bool got = nc < O.n_chunks;
          if (rank < rem) slot_row = (long long)chunk * MCD_CHUNK_ROWS + used + rank;
          else if (got) slot_row = (long long)nc * MCD_CHUNK_ROWS + (rank - rem);
          w_stored += got ? k : rem;
          w_dropped += got ? 0u : (k - rem);
          chunk = got ? (int)nc : -1;
          used = got ? (k - rem) : 0u;
        }
        if (has_row && slot_row >= 0)
          store_scat_row(O.scats + (size_t)slot_row * (size_t)P.scat_row_words, P.scat_mask, (uint32_t)rel_id, ev,
                         slot_channel(P.m, row.slot), row, e);
      }
    }
    if (has_row) {
      if (want_chan) atomicAdd(&s_cnt[kCntChan + slot_channel(P.m, row.slot)], 1ull);
      if (row.flags & (MCD_EVF_NANCLAMP | MCD_EVF_EXACT)) {
        if (row.flags & MCD_EVF_NANCLAMP) atomicAdd(&s_cnt[kCntNan], 1ull);
        if (row.flags & MCD_EVF_EXACT) atomicAdd(&s_cnt[kCntExact], 1ull);
      }
      ++ev;
    }

  }

  // ---- close this warp's chunk, flush block statistics ---------------------------------------------
  if (want_scats && lane == 0) {
    if (chunk >= 0) O.chunk_count[chunk] = used;
    if (w_stored) atomicAdd(&s_rows[0], (unsigned long long)w_stored);
    if (w_dropped) atomicAdd(&s_rows[1], (unsigned long long)w_dropped);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MCD_HIST_TOTAL; i += kThreads) {
    const unsigned int v = s_hist[i];
    if (v) atomicAdd(O.stats + i, (unsigned long long)v);
  }
  unsigned long long* cnt = O.stats + MCD_HIST_TOTAL;
  // mcd_counters layout: electrons_done, events, outcome[5], double_records, stored, dropped, nan, exact, channel[]
  if (threadIdx.x == 0) {
    atomicAdd(cnt + 0, s_cnt[kCntDone]);
    atomicAdd(cnt + 1, s_cnt[kCntEvents]);
    for (int i = 0; i < 5; ++i) if (s_cnt[kCntOutcome + i]) atomicAdd(cnt + 2 + i, s_cnt[kCntOutcome + i]);
    if (s_cnt[kCntDouble]) atomicAdd(cnt + 7, s_cnt[kCntDouble]);
    if (want_scats) {
      if (s_rows[0]) atomicAdd(cnt + 8, s_rows[0]);
      if (s_rows[1]) atomicAdd(cnt + 9, s_rows[1]);
    }
    if (s_cnt[kCntNan]) atomicAdd(cnt + 10, s_cnt[kCntNan]);
    if (s_cnt[kCntExact]) atomicAdd(cnt + 11, s_cnt[kCntExact]);
    if (want_chan)
      for (int i = 0; i < MCD_MAX_CHANNELS; ++i)
        if (s_cnt[kCntChan + i]) atomicAdd(cnt + 12 + i, s_cnt[kCntChan + i]);
  }
}

typedef void (*TrajKernel)(const DevParams, const DevOut, const unsigned long long, const unsigned long long);

template <bool SCATS>
TrajKernel pick_kernel_t(const DevParams& P) {
  if (P.mode == MCD_MODE_CONT) return traj_kernel<1, MCD_MODE_CONT, SCATS>;
  switch (P.m.nshell) {
    case 1: return traj_kernel<1, MCD_MODE_DS, SCATS>;
    case 2: return traj_kernel<2, MCD_MODE_DS, SCATS>;
    case 3: return traj_kernel<3, MCD_MODE_DS, SCATS>;
    case 4: return traj_kernel<4, MCD_MODE_DS, SCATS>;
    case 5: return traj_kernel<5, MCD_MODE_DS, SCATS>;
    case 6: return traj_kernel<6, MCD_MODE_DS, SCATS>;
    case 7: return traj_kernel<7, MCD_MODE_DS, SCATS>;
    case 8: return traj_kernel<8, MCD_MODE_DS, SCATS>;
    default: return nullptr;
  }
}

TrajKernel pick_kernel(const DevParams& P, bool scats) {
  return scats ? pick_kernel_t<true>(P) : pick_kernel_t<false>(P);
}

}  // namespace

// ---- regroup the appended scats rows into trajectory order (the reference's row order: electrons in id order,
// events in order; multiScatter.py:107-112 appends one whole trajectory after the other) -------------------------
namespace {

constexpr int kScanT = 256;
constexpr int kScanItems = 8;

// offsets[i] = sum_{j<i} nev[j] (exclusive, 64 bit).  Three small kernels: per-tile sums, scan of the tile sums
// by one block, per-tile exclusive scan plus tile offset.
__global__ void scan_tile_sums(const unsigned int* __restrict__ nev, unsigned long long n,
                               unsigned long long* __restrict__ tile_sum) {
  __shared__ unsigned long long s_w[kScanT / 32];
  const unsigned long long base = (unsigned long long)blockIdx.x * kScanT * kScanItems;
  unsigned long long v = 0;
  for (int k = 0; k < kScanItems; ++k) {
    const unsigned long long i = base + (unsigned long long)k * kScanT + threadIdx.x;
    if (i < n) v += nev[i];
  }
  for (int d = 16; d > 0; d >>= 1) v += __shfl_down_sync(MCD_FULL_MASK, v, d);
  if ((threadIdx.x & 31) == 0) s_w[threadIdx.x >> 5] = v;
  __syncthreads();
  if (threadIdx.x == 0) {
    unsigned long long t = 0;
    for (int w = 0; w < kScanT / 32; ++w) t += s_w[w];
    tile_sum[blockIdx.x] = t;
  }
}

__global__ void scan_tiles(unsigned long long* __restrict__ tile_sum, unsigned int ntiles,
                           unsigned long long* __restrict__ total) {
  // single block, sequential over chunks of blockDim tiles (ntiles is a few thousand at most)
  __shared__ unsigned long long s_v[1024];
  __shared__ unsigned long long s_carry;
  if (threadIdx.x == 0) s_carry = 0;
  __syncthreads();
  for (unsigned int b = 0; b < ntiles; b += blockDim.x) {
    const unsigned int i = b + threadIdx.x;
    const unsigned long long v = (i < ntiles) ? tile_sum[i] : 0ull;
    s_v[threadIdx.x] = v;
    __syncthreads();
    for (unsigned int d = 1; d < blockDim.x; d <<= 1) {              // Hillis-Steele inclusive scan
      const unsigned long long a = (threadIdx.x >= d) ? s_v[threadIdx.x - d] : 0ull;
      __syncthreads();
      s_v[threadIdx.x] += a;
      __syncthreads();
    }
    if (i < ntiles) tile_sum[i] = s_carry + s_v[threadIdx.x] - v;    // exclusive
    __syncthreads();
    if (threadIdx.x == blockDim.x - 1) s_carry += s_v[threadIdx.x];
    __syncthreads();
  }
  if (threadIdx.x == 0) *total = s_carry;
}

__global__ void scan_apply(const unsigned int* __restrict__ nev, unsigned long long n,
                           const unsigned long long* __restrict__ tile_off, unsigned long long* __restrict__ offsets) {
  // one thread handles kScanItems CONSECUTIVE electrons; block-level exclusive scan of the thread sums
  __shared__ unsigned long long s_w[kScanT / 32];
  const unsigned long long base = (unsigned long long)blockIdx.x * kScanT * kScanItems + (unsigned long long)threadIdx.x * kScanItems;
  unsigned int loc[kScanItems];
  unsigned long long sum = 0;
  for (int k = 0; k < kScanItems; ++k) {
    loc[k] = (base + k < n) ? nev[base + k] : 0u;
    sum += loc[k];
  }
  unsigned long long incl = sum;
  const unsigned lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (int d = 1; d < 32; d <<= 1) {
    const unsigned long long o = __shfl_up_sync(MCD_FULL_MASK, incl, d);
    if (lane >= (unsigned)d) incl += o;
  }
  if (lane == 31) s_w[warp] = incl;
  __syncthreads();
  unsigned long long woff = 0;
  for (unsigned w = 0; w < warp; ++w) woff += s_w[w];
  unsigned long long run = tile_off[blockIdx.x] + woff + incl - sum;
  for (int k = 0; k < kScanItems; ++k) {
    if (base + k < n) offsets[base + k] = run;
    run += loc[k];
  }
  if (base <= n && n < base + kScanItems) offsets[n] = run - 0;       // offsets[n] = total (set by the owner of n)
}

// every stored row of the chunk pool goes to sorted[offsets[rel_id] + event].  One block per chunk (1024 rows), one
// row per thread and pass; rows are moved as 8-byte words when the row length allows (row starts are then 8-byte
// aligned in both buffers), so a 24-byte row is three loads and three stores.
__global__ void __launch_bounds__(256) scatter_rows(const uint32_t* __restrict__ pool,
                                                    const unsigned int* __restrict__ chunk_count, unsigned int nchunks,
                                                    int row_words, const unsigned long long* __restrict__ offsets,
                                                    uint32_t* __restrict__ sorted) {
  for (unsigned int c = blockIdx.x; c < nchunks; c += gridDim.x) {
    const unsigned int cnt = chunk_count[c];
    const uint32_t* cbase = pool + (size_t)c * MCD_CHUNK_ROWS * (size_t)row_words;
    if ((row_words & 1) == 0) {
      const int rw2 = row_words >> 1;
      for (unsigned int r = threadIdx.x; r < cnt; r += blockDim.x) {
        const uint2* src = reinterpret_cast<const uint2*>(cbase) + (size_t)r * rw2;
        const uint2 head = __ldcs(src);
        const unsigned long long d = offsets[head.x] + (unsigned long long)(head.y & 0xffffffu);
        uint2* dst = reinterpret_cast<uint2*>(sorted) + d * (unsigned long long)rw2;
        dst[0] = head;
        for (int w = 1; w < rw2; ++w) dst[w] = __ldcs(src + w);
      }
    } else {
      for (unsigned int r = threadIdx.x; r < cnt; r += blockDim.x) {
        const uint32_t* src = cbase + (size_t)r * row_words;
        const uint32_t rel = src[0], w1 = src[1];
        uint32_t* dst = sorted + (offsets[rel] + (unsigned long long)(w1 & 0xffffffu)) * (unsigned long long)row_words;
        dst[0] = rel;
        dst[1] = w1;
        for (int w = 2; w < row_words; ++w) dst[w] = src[w];
      }
    }
  }
}

}  // namespace

cudaError_t mcd_launch_order_scats(const unsigned int* nev, unsigned long long count, unsigned long long* tile_tmp,
                                   unsigned long long* offsets, unsigned long long* total_dev, const uint32_t* pool,
                                   const unsigned int* chunk_count, unsigned int nchunks, int row_words,
                                   uint32_t* sorted, int sm_count, cudaStream_t stream) {
  const unsigned long long per_tile = (unsigned long long)kScanT * kScanItems;
  const unsigned int ntiles = (unsigned int)((count + 1 + per_tile - 1) / per_tile);
  scan_tile_sums<<<ntiles, kScanT, 0, stream>>>(nev, count, tile_tmp);
  scan_tiles<<<1, 1024, 0, stream>>>(tile_tmp, ntiles, total_dev);
  scan_apply<<<ntiles, kScanT, 0, stream>>>(nev, count, tile_tmp, offsets);
  if (sorted && nchunks) {
    const unsigned int grid = nchunks < (unsigned int)(sm_count * 64) ? nchunks : (unsigned int)(sm_count * 64);
    scatter_rows<<<grid, 256, 0, stream>>>(pool, chunk_count, nchunks, row_words, offsets, sorted);
  }
  return cudaGetLastError();
}

unsigned long long mcd_order_tiles(unsigned long long count) {
  const unsigned long long per_tile = (unsigned long long)kScanT * kScanItems;
  return (count + 1 + per_tile - 1) / per_tile;
}

cudaError_t mcd_traj_max_blocks(const DevParams& P, int sm_count, TrajLaunchInfo* info) {
  // both variants have the same launch bounds; the occupancy of the row-writing one is reported
  TrajKernel k = pick_kernel(P, true);
  if (!k) return cudaErrorInvalidValue;
  int per_sm = 0;
  cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, kThreads, 0);
  if (err != cudaSuccess) return err;
  cudaFuncAttributes fa;
  err = cudaFuncGetAttributes(&fa, k);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  if (const char* env = getenv("MCD_BLOCKS_PER_SM")) {      // tuning knob: fewer resident blocks than the maximum
    const int v = atoi(env);
    if (v >= 1 && v < per_sm) per_sm = v;
  }
  info->blocks = per_sm * sm_count;
  info->threads = kThreads;
  info->regs = fa.numRegs;
  info->blocks_per_sm = per_sm;
  return cudaSuccess;
}

cudaError_t mcd_launch_traj(const DevParams& P, const DevOut& O, unsigned long long first_id,
                            unsigned long long count, int sm_count, cudaStream_t stream, TrajLaunchInfo* info) {
  TrajKernel k = pick_kernel(P, (P.scat_row_words > 0) && (O.scats != nullptr));
  if (!k) return cudaErrorInvalidValue;
  TrajLaunchInfo li;
  cudaError_t err = mcd_traj_max_blocks(P, sm_count, &li);
  if (err != cudaSuccess) return err;
  if (info) *info = li;
  k<<<li.blocks, kThreads, 0, stream>>>(P, O, first_id, count);
  return cudaGetLastError();
}
