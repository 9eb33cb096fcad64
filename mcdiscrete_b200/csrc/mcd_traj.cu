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
// the park-pool kernel has its own geometry: 256 threads x 4 blocks = 32 warps per SM at 64 registers (no spills) and
// 148 KB of shared memory, which leaves more L1 than 192 x 6 (r2c / r2d A/B runs: -1.8 % on C2)
#ifndef MCD_POOL_THREADS
#define MCD_POOL_THREADS 256
#endif
#ifndef MCD_POOL_MIN_BLOCKS
#define MCD_POOL_MIN_BLOCKS 4
#endif
constexpr int kPoolThreads = MCD_POOL_THREADS;
constexpr int kCntOutcome = 0;       // 5
constexpr int kCntEvents = 5;
constexpr int kCntDone = 6;
constexpr int kCntDouble = 7;
constexpr int kCntNan = 8;
constexpr int kCntExact = 9;
constexpr int kCntChan = 10;         // MCD_MAX_CHANNELS
constexpr int kCntEventsHi = kCntChan + MCD_MAX_CHANNELS;      // carry word of the event count
constexpr int kCntTotal = kCntEventsHi + 1;

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

// saveOutcomes (electron.py:105-113) of a lane whose trajectory ended: the 32-byte els row (one full sector, two
// 16-byte stores) and the outcome / event counters.  preE = energy before the event's loss; e = direction before the
// event's rotation, position as update_xyz left it.  The backscatter histograms are NOT built here: they
// are a function of the els row alone and are binned by hist_kernel in one pass over the rows after the last launch
// of the run (no acosf / atan2f code, 6 KB of shared memory and 1564-bin flush per block in the trajectory kernels).
__device__ __forceinline__ void retire_lane(const DevOut& O, unsigned int* s_cnt, int outcome, bool dbl,
                                            float preE, const ElState& e, uint32_t rel_id, uint32_t nev) {
  float4* dst = reinterpret_cast<float4*>(O.els + rel_id);
  dst[0] = make_float4(__uint_as_float((uint32_t)outcome | (dbl ? 0x100u : 0u)), preE, e.dx, e.dy);
  dst[1] = make_float4(e.dz, e.x, e.y, e.z);
  // block counters are 32-bit words (native shared-memory atomics; a 64-bit shared atomicAdd is a CAS loop); the
  // event count carries into a second word
  atomicAdd(&s_cnt[kCntOutcome + outcome], 1u);
  atomicAdd(&s_cnt[kCntDone], 1u);
  if (atomicAdd(&s_cnt[kCntEvents], nev) + nev < nev) atomicAdd(&s_cnt[kCntEventsHi], 1u);
  if (dbl) atomicAdd(&s_cnt[kCntDouble], 1u);
}

// block counters -> global stats block (histograms || mcd_counters), once per block
__device__ __forceinline__ void flush_block_stats(const DevOut& O, const unsigned int* s_cnt,
                                                  const unsigned long long* s_rows, bool want_scats, bool want_chan) {
  unsigned long long* cnt = O.stats + MCD_HIST_TOTAL;
  // mcd_counters layout: electrons_done, events, outcome[5], double_records, stored, dropped, nan, exact, channel[]
  if (threadIdx.x == 0) {
    atomicAdd(cnt + 0, (unsigned long long)s_cnt[kCntDone]);
    atomicAdd(cnt + 1, ((unsigned long long)s_cnt[kCntEventsHi] << 32) | s_cnt[kCntEvents]);
    for (int i = 0; i < 5; ++i) if (s_cnt[kCntOutcome + i]) atomicAdd(cnt + 2 + i, (unsigned long long)s_cnt[kCntOutcome + i]);
    if (s_cnt[kCntDouble]) atomicAdd(cnt + 7, (unsigned long long)s_cnt[kCntDouble]);
    if (want_scats) {
      if (s_rows[0]) atomicAdd(cnt + 8, s_rows[0]);
      if (s_rows[1]) atomicAdd(cnt + 9, s_rows[1]);
    }
    if (s_cnt[kCntNan]) atomicAdd(cnt + 10, (unsigned long long)s_cnt[kCntNan]);
    if (s_cnt[kCntExact]) atomicAdd(cnt + 11, (unsigned long long)s_cnt[kCntExact]);
    if (want_chan)
      for (int i = 0; i < MCD_MAX_CHANNELS; ++i)
        if (s_cnt[kCntChan + i]) atomicAdd(cnt + 12 + i, (unsigned long long)s_cnt[kCntChan + i]);
  }
}

// Backscatter histograms of one run from its els rows (final energy 512 bins [Emin, E0], exit polar
// angle from the surface normal 180 bins [0, 90] deg, exit azimuth 360 bins, last interior depth 512 bins [0, 2R]).
// One pass over 32 bytes per electron (HBM-bound: ~5 us per 1e6 electrons), block histograms in shared memory,
// one 64-bit atomic per non-empty bin and block.
__global__ void __launch_bounds__(256) hist_kernel(const DevHistSpec H, const mcd_el_row* __restrict__ els,
                                                   const unsigned long long count, unsigned long long* __restrict__ stats) {
  __shared__ unsigned int s_hist[MCD_HIST_TOTAL];
  for (int i = threadIdx.x; i < MCD_HIST_TOTAL; i += blockDim.x) s_hist[i] = 0u;
  __syncthreads();
  const float4* rows = reinterpret_cast<const float4*>(els);
  for (unsigned long long r = (unsigned long long)blockIdx.x * blockDim.x + threadIdx.x; r < count;
       r += (unsigned long long)gridDim.x * blockDim.x) {
    const float4 a = __ldcs(rows + 2 * r);               // code, final_E, dx, dy
    if ((__float_as_uint(a.x) & 0xffu) != (uint32_t)MCD_OUT_BKS) continue;
    const float4 b = __ldcs(rows + 2 * r + 1);           // dz, last_x, last_y, last_z
    int k = (int)((a.y - H.e_lo) * H.e_scale);
    k = max(0, min(k, MCD_HIST_E_BINS - 1));
    atomicAdd(&s_hist[k], 1u);
    const float pol = acosf(fminf(1.0f, fmaxf(-1.0f, -b.x))) * 57.29577951308232f;       // from the -z normal
    k = max(0, min((int)(pol * 2.0f), MCD_HIST_POLAR_BINS - 1));
    atomicAdd(&s_hist[MCD_HIST_E_BINS + k], 1u);
    const float az = atan2f(a.w, a.z) * 57.29577951308232f + 180.0f;
    k = max(0, min((int)az, MCD_HIST_AZ_BINS - 1));
    atomicAdd(&s_hist[MCD_HIST_E_BINS + MCD_HIST_POLAR_BINS + k], 1u);
    k = max(0, min((int)(b.w * H.depth_scale), MCD_HIST_DEPTH_BINS - 1));
    atomicAdd(&s_hist[MCD_HIST_E_BINS + MCD_HIST_POLAR_BINS + MCD_HIST_AZ_BINS + k], 1u);
  }
  __syncthreads();
  for (int i = threadIdx.x; i < MCD_HIST_TOTAL; i += blockDim.x) {
    const unsigned int v = s_hist[i];
    if (v) atomicAdd(stats + i, (unsigned long long)v);
  }
}

// dst += src; src = 0 (both histogram || counter blocks): the contexts of a pipeline are merged into one block
// before the run's single all-reduce
__global__ void stats_merge_kernel(unsigned long long* __restrict__ dst, unsigned long long* __restrict__ src, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    const unsigned long long v = src[i];
    if (v) { dst[i] += v; src[i] = 0ull; }
  }
}

// POLY = 1: the per-cell cross-section records exist (DevParams::use_sig_poly): no formula code, no flag tests;
// POLY = 0: the S+3 formulas are evaluated every event (channel order not static on [Emin, E0], or elastic = Mott)
template <int S, int MODE, bool SCATS, int POLY>
__global__ void __launch_bounds__(kThreads, MCD_MIN_BLOCKS)
traj_kernel(const __grid_constant__ DevParams P, const DevOut O, const unsigned long long first_id,
            const unsigned long long count) {
  __shared__ unsigned int s_cnt[kCntTotal];
  __shared__ unsigned long long s_rows[2];      // stored, dropped

  for (int i = threadIdx.x; i < kCntTotal; i += kThreads) s_cnt[i] = 0u;
  if (threadIdx.x < 2) s_rows[threadIdx.x] = 0ull;
  __syncthreads();

  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = lanemask_lt();
  constexpr bool want_scats = SCATS;            // compile time: the no-scats kernel carries no row fields
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
      HotRec hr;
      if (MODE == MCD_MODE_DS) hr = ds_hot_fetch<POLY>(P, e.E);    // loads first, Philox hides their latency
      const float4 u = event_uniforms(P, first_id + rel_id, ev);
      if (MODE == MCD_MODE_DS) {
        EventMid mid;
        float tot;
        ds_phase_a<S, POLY>(P, e, u, row, mid, tot, hr);
        ds_phase_b<S>(P, e, u, row, mid, tot);
      } else {
        cont_event(P, e, u, row);
      }
      outcome = row.outcome;
    }

    // ---- retire: saveOutcomes row (electron.py:105-113), histograms, counters.  Taken BEFORE the rotation of this
    // event: final_E = energy before the event's loss, final_dir = direction before its rotation, last_pos =
    // position as update_xyz left it -----------------------------------------------------------------
    if (outcome >= 0) {
      const uint32_t nev = ev + (has_row ? 1u : 0u);
      if (want_scats) O.nev[rel_id] = nev;
      retire_lane(O, s_cnt, outcome, (row.flags & MCD_EVF_DOUBLE) != 0, row.E, e, rel_id, nev);
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
      if (want_chan) atomicAdd(&s_cnt[kCntChan + slot_channel(P.m, row.slot)], 1u);
      if (row.flags & (MCD_EVF_NANCLAMP | MCD_EVF_EXACT)) {
        if (row.flags & MCD_EVF_NANCLAMP) atomicAdd(&s_cnt[kCntNan], 1u);
        if (row.flags & MCD_EVF_EXACT) atomicAdd(&s_cnt[kCntExact], 1u);
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
  flush_block_stats(O, s_cnt, s_rows, want_scats, want_chan);
}

// ---------------------------------------------------------------------------------------------------------------
// K1p: trajectory_DS without row recording (statistics + els rows only): table events go through a warp-private
// PARK POOL in shared memory.
//
// 12 % of the DS events (Moller / Gryzinski) need an inverse-CDF table search = a chain of three dependent memory
// accesses executed by ~4 of the 32 lanes; in traj_kernel 98 % of the warp-iterations pay that chain (40 % of the
// stall samples, 12 % of the warp instructions at 3.4 lanes).  Here a lane whose event picked a table channel PARKS
// its electron (state + the two uniforms it still needs, 14 words) in the warp's pool and takes another electron at
// once.  When 32 electrons are parked the warp runs one BATCH iteration: the lanes swap their current electrons into
// the pool (as 'ready' entries) for the 32 parked ones and all 32 lanes search together -- one converged chain per 32
// table events instead of one per warp-iteration.  Ready entries re-enter lanes as lanes park or retire.  Results do
// not depend on any of this: the uniforms of an event are a function of (seed, electron id, event index).
//
// Pool layout (per warp): slots of kSlotWords words moved with 128-bit shared-memory accesses; slots [0, n_ready)
// hold ready electrons, slots [kPoolSlots - n_pend, kPoolSlots) hold parked ones (a stack growing downwards).
// ---------------------------------------------------------------------------------------------------------------
#ifndef MCD_POOL_SLOTS
#define MCD_POOL_SLOTS 48
#endif
constexpr int kPoolSlots = MCD_POOL_SLOTS;
#ifndef MCD_BATCH_MIN
#define MCD_BATCH_MIN 32             // parked electrons that trigger a batch iteration (<= 32)
#endif
constexpr int kBatchMin = MCD_BATCH_MIN;
constexpr int kSlotWords = 20;       // 16 used: E x y z | dx dy dz yx | yy yz rel_id ev+bits | u.z u.w u.x total; a stride of 20
                                     // words keeps the 128-bit accesses of 8 consecutive slots on distinct banks

// a parked event (table channel): resolve which one (ds_cold_slot), search its table, binary-collision angle
template <int S>
__device__ __forceinline__ void ds_table_event(const DevParams& P, const ElState& e, float ux, float tot, float uz,
                                               int& slot, int& flags, float& W, float& om) {
  if (slot == MCD_SLOT_UNRESOLVED) slot = ds_cold_slot<S>(P, e.E, ux, tot, flags);
  if (slot >= 2) {
    ds_table_loss(P, e, slot_table<S>(P, slot), uz, flags, W, om);
  } else {                                                   // the exact path moved it back to Ruth / Quinn (rare)
    W = (slot == 1) ? P.m.Epl : 0.0f;
    om = (slot == 0) ? ruth_om(P.m, f_rcp(e.E), uz) : 0.0f;
  }
}

// out-of-line copy for the pool-full case (keeps the hot loop's code small); everything travels by value so that
// the caller's slot / flags stay in registers: returns (W, om, slot, flags)
template <int S>
static __device__ __noinline__ float4 table_event_call(const DevParams& P, float E, float ux, float tot, float uz,
                                                       int slot, int flags) {
  ElState t;
  t.E = E;
  float W, om;
  ds_table_event<S>(P, t, ux, tot, uz, slot, flags, W, om);
  return make_float4(W, om, __int_as_float(slot), __int_as_float(flags));
}

__device__ __forceinline__ void pool_put(uint32_t* slot, const ElState& e, uint32_t rel_id, uint32_t evw) {
  uint4* p = reinterpret_cast<uint4*>(slot);
  p[0] = make_uint4(__float_as_uint(e.E), __float_as_uint(e.x), __float_as_uint(e.y), __float_as_uint(e.z));
  p[1] = make_uint4(__float_as_uint(e.dx), __float_as_uint(e.dy), __float_as_uint(e.dz), __float_as_uint(e.yx));
  p[2] = make_uint4(__float_as_uint(e.yy), __float_as_uint(e.yz), rel_id, evw);
}

__device__ __forceinline__ void pool_unpack(const uint4& a, const uint4& b, const uint4& c, ElState& e,
                                            uint32_t& rel_id, uint32_t& evw) {
  e.E = __uint_as_float(a.x); e.x = __uint_as_float(a.y); e.y = __uint_as_float(a.z); e.z = __uint_as_float(a.w);
  e.dx = __uint_as_float(b.x); e.dy = __uint_as_float(b.y); e.dz = __uint_as_float(b.z); e.yx = __uint_as_float(b.w);
  e.yy = __uint_as_float(c.x); e.yz = __uint_as_float(c.y);
  rel_id = c.z;
  evw = c.w;
}

#define MCD_POOL_EACH_STATE(OP) \
  OP(0, e.E) OP(1, e.x) OP(2, e.y) OP(3, e.z) OP(4, e.dx) OP(5, e.dy) OP(6, e.dz) OP(7, e.yx) OP(8, e.yy) OP(9, e.yz)

template <int S, bool RESUME>
__global__ void __launch_bounds__(kPoolThreads, MCD_POOL_MIN_BLOCKS)
traj_kernel_pool(const __grid_constant__ DevParams P, const DevOut O, const unsigned long long first_id,
                 const unsigned long long count_arg) {
  // RESUME pass: the "ids" are the entries of the leftover queue the previous pass wrote; only as many blocks as
  // are needed to give every lane two electrons (lanes + pool) do any work
  const unsigned long long count = RESUME ? *O.rq_in_count : count_arg;
  if (RESUME && (unsigned long long)blockIdx.x * (2ull * kPoolThreads) >= count) return;
  __shared__ unsigned int s_cnt[kCntTotal];
  __shared__ __align__(16) uint32_t s_pool[kPoolThreads / 32][kSlotWords * kPoolSlots];

  for (int i = threadIdx.x; i < kCntTotal; i += kPoolThreads) s_cnt[i] = 0u;
  __syncthreads();

  const unsigned lane = threadIdx.x & 31u;
  const unsigned lt = lanemask_lt();
  const bool want_chan = (P.flags & MCD_FLAG_CHANNEL_COUNTS) != 0u;
  uint32_t* const pool = s_pool[threadIdx.x >> 5];
  // warp-uniform bookkeeping in ONE register (the loop is at the register limit of 6 blocks per SM):
  // bits 0-7 ready entries, bits 8-15 parked entries, bit 16 = the ticket ran past `count`
  uint32_t wst = 0u;
#define n_ready ((int)(wst & 0xffu))
#define n_pend ((int)((wst >> 8) & 0xffu))
#define exhausted ((wst >> 16) != 0u)

  int alive = 0;                     // (int, not bool: the compiler keeps bools of this loop packed in bytes)
  ElState e;
  uint32_t rel_id = 0u, ev = 0u;

  while (true) {
    // ---- refill dead lanes: ready electrons of the pool first, then fresh ones (multiScatter.py:105-107) ------
    const unsigned need = __ballot_sync(MCD_FULL_MASK, !alive);
    if (need) {
      const int nneed = __popc(need);
      const int rank = __popc(need & lt);
      const int take = min(nneed, n_ready);
      if (!alive && rank < take) {                           // (take = 0: no lane)
        const uint4* q = reinterpret_cast<const uint4*>(pool + (n_ready - 1 - rank) * kSlotWords);
        pool_unpack(q[0], q[1], q[2], e, rel_id, ev);
        alive = true;
      }
      wst -= (uint32_t)take;
      const int nreq = nneed - take;
      if (nreq > 0 && !exhausted) {
        const unsigned still = __ballot_sync(MCD_FULL_MASK, !alive);
        const int leader = __ffs(still) - 1;
        unsigned long long base = 0ull;
        if ((int)lane == leader) base = atomicAdd(O.ticket, (unsigned long long)nreq);
        base = __shfl_sync(MCD_FULL_MASK, base, leader);
        if (!alive) {
          const unsigned long long rid = base + (unsigned long long)(rank - take);
          if (rid < count) {
            alive = true;
            if (RESUME) {                                          // an electron the previous pass handed over
              const uint32_t* q = O.rq_in + rid;
#define MCD_LDQ(w, f) f = __uint_as_float(__ldcs(q + (size_t)(w) * O.rq_cap));
              MCD_POOL_EACH_STATE(MCD_LDQ)
#undef MCD_LDQ
              rel_id = __ldcs(q + (size_t)10 * O.rq_cap);
              ev = __ldcs(q + (size_t)11 * O.rq_cap);
            } else {
              rel_id = (uint32_t)rid;
              ev = 0u;
              e.E = P.E0;
              e.x = 0.f; e.y = 0.f; e.z = 0.f;
              e.dx = P.d0x; e.dy = P.d0y; e.dz = P.d0z;
              e.yx = 0.f; e.yy = 1.f; e.yz = 0.f;                  // electron.py:38
            }
          }
        }
        if (base + (unsigned long long)nreq >= count) wst |= 0x10000u;
      }
    }
    // until the ids run out every lane holds an electron here (the refill above filled all of them)
    unsigned am = MCD_FULL_MASK;
    int a = 32;
    // ---- tail: once the ids have run out a warp that is down to dump_T electrons hands them to the next pass
    // (leftover queue in global memory) and leaves; that pass runs them in full warps again.  A trajectory is
    // 7..870 events long, so without this every warp (and its block's slot on the SM) lingers for ~800 iterations
    // at an average of 12 live lanes.  Parked electrons are searched first (forced batch), so the queue only holds
    // plain states. ---------------------------------------------------------------------------------------------
    bool tail = false;
    // batch trigger: while ids are left all 32 lanes are alive, so it is one masked compare of the packed counters
    bool batch = (wst & 0xff00u) >= ((uint32_t)kBatchMin << 8);
    if (exhausted) {
      am = __ballot_sync(MCD_FULL_MASK, alive);
      a = __popc(am);
      if (a == 0 && n_pend == 0) break;
      tail = (a + n_pend <= O.dump_T);
      batch = n_pend > 0 && (n_pend >= min(a, kBatchMin) || tail);
      if (tail && n_pend == 0) {
        const int leader = __ffs(am) - 1;
        unsigned long long base = 0ull;
        if ((int)lane == leader) base = atomicAdd(O.rq_out_count, (unsigned long long)a);
        base = __shfl_sync(MCD_FULL_MASK, base, leader);
        if (alive) {
          uint32_t* q = O.rq_out + (base + (unsigned long long)__popc(am & lt));
#define MCD_STQ(w, f) q[(size_t)(w) * O.rq_cap] = __float_as_uint(f);
          MCD_POOL_EACH_STATE(MCD_STQ)
#undef MCD_STQ
          q[(size_t)10 * O.rq_cap] = rel_id;
          q[(size_t)11 * O.rq_cap] = ev;
        }
        break;
      }
    }

    int done = 0;                                  // this lane completes an event in this iteration
    int outcome = -1, slot = 0, flags = 0;
    // W must be 0 for a lane that completes no event (branch-free energy update below); om / uphi are read under
    // `done` only and are deliberately left without an initialiser (no zeroing instructions per iteration)
    float W = 0.f, om, uphi;
    if (batch) {
      // ---- BATCH iteration: current electrons -> ready entries, parked electrons -> lanes, then search ----------
      // (the two pool regions may overlap: all lanes read their parked entry before any lane stores)
      const int nb = min(n_pend, 32);
      const bool ld = (int)lane < nb;
      // lanes >= nb keep their electron and sit this iteration out (only when fewer than 32 are parked)
      const bool st = alive && (ld || tail);
      const unsigned sm = __ballot_sync(MCD_FULL_MASK, st);
      const uint4* q = reinterpret_cast<const uint4*>(pool + (kPoolSlots - n_pend + (int)lane) * kSlotWords);
      // vector by vector, so that only four temporaries are live: read (all lanes) -> barrier -> store own -> adopt
      uint4* w = reinterpret_cast<uint4*>(pool + (n_ready + __popc(sm & lt)) * kSlotWords);
      uint4 t = make_uint4(0u, 0u, 0u, 0u);
      if (ld) t = q[0];
      __syncwarp();
      if (st) w[0] = make_uint4(__float_as_uint(e.E), __float_as_uint(e.x), __float_as_uint(e.y), __float_as_uint(e.z));
      if (ld) { e.E = __uint_as_float(t.x); e.x = __uint_as_float(t.y); e.y = __uint_as_float(t.z); e.z = __uint_as_float(t.w); }
      if (ld) t = q[1];
      __syncwarp();
      if (st) w[1] = make_uint4(__float_as_uint(e.dx), __float_as_uint(e.dy), __float_as_uint(e.dz), __float_as_uint(e.yx));
      if (ld) { e.dx = __uint_as_float(t.x); e.dy = __uint_as_float(t.y); e.dz = __uint_as_float(t.z); e.yx = __uint_as_float(t.w); }
      if (ld) t = q[2];
      __syncwarp();
      if (st) w[2] = make_uint4(__float_as_uint(e.yy), __float_as_uint(e.yz), rel_id, ev);
      if (ld) { e.yy = __uint_as_float(t.x); e.yz = __uint_as_float(t.y); rel_id = t.z; ev = t.w; }
      if (ld) t = q[3];                                          // u.z, u.w, u.x, total of the parked event
      uphi = __uint_as_float(t.y);
      wst += (uint32_t)__popc(sm) - ((uint32_t)nb << 8);
      alive = ld || (alive && !st);
      __syncwarp();
      if (ld) {
        slot = (int)((ev >> 24) & 15u);
        outcome = (int)((ev >> 28) & 7u) - 1;
        flags = 0;                                 // (exact picks of phase A were counted when they happened)
        ev &= 0xffffffu;
        ds_table_event<S>(P, e, __uint_as_float(t.z), __uint_as_float(t.w), __uint_as_float(t.x), slot, flags, W, om);
        done = true;
        if (flags) {                               // rare: counted here, not after every event of every iteration
          if (flags & MCD_EVF_NANCLAMP) atomicAdd(&s_cnt[kCntNan], 1u);
          if (flags & MCD_EVF_EXACT) atomicAdd(&s_cnt[kCntExact], 1u);
        }
      }
    } else {
      // ---- normal iteration of trajectory_DS: an alive lane executes one event or, past maxScatt, retires as
      // 'long' (singleScatter.py:19-23); a lane whose event needs a table parks its electron -----------------------
      const bool has_row = alive && (ev < (uint32_t)P.maxScatt);
      if (alive && !has_row) outcome = MCD_OUT_LONG;
      bool park = false;
      float uz, ux, tot;                           // read under `park` only
      EventRow row;
      if (has_row) {
        const HotRec hr = ds_hot_fetch<1>(P, e.E);                  // loads first, Philox hides their latency
        const float4 u = event_uniforms(P, first_id + rel_id, ev);
        EventMid mid;
        ds_phase_a<S, 1>(P, e, u, row, mid, tot, hr, &s_cnt[kCntExact]);
        outcome = row.outcome;
        slot = row.slot;
        flags = row.flags;
        uphi = u.w;
        W = mid.W;
        om = mid.om;
        uz = u.z;
        ux = u.x;
        park = mid.tab >= 0;
        done = !park;
      }
      const unsigned pm = __ballot_sync(MCD_FULL_MASK, park);
      if (pm) {
        const int room = kPoolSlots - n_ready - n_pend;
        const int r = __popc(pm & lt);
        const int top = kPoolSlots - n_pend - 1;                 // first free slot of the parked stack
        const int nput = min(__popc(pm), room);
#ifndef MCD_NO_PARK_TAKE
        // a lane that parks takes a ready electron of the pool in the same breath (the refill at the top of the loop
        // then only serves retired lanes and an empty pool): one pool transaction per parked event instead of two
        const int first_ready = n_ready - 1;
        const int ntake = min(nput, n_ready);
        wst += ((uint32_t)nput << 8) - (uint32_t)ntake;          // (before the branches: nothing lives across the call)
#else
        wst += (uint32_t)nput << 8;
#endif
        if (park) {
          if (r < room) {
            uint32_t* w = pool + (top - r) * kSlotWords;
            pool_put(w, e, rel_id, ev | ((uint32_t)slot << 24) | ((uint32_t)(outcome + 1) << 28));
            *reinterpret_cast<uint4*>(w + 12) = make_uint4(__float_as_uint(uz), __float_as_uint(uphi), __float_as_uint(ux), __float_as_uint(tot));
            alive = false;
            outcome = -1;
#ifndef MCD_NO_PARK_TAKE
            if (r < ntake) {
              const uint4* q = reinterpret_cast<const uint4*>(pool + (first_ready - r) * kSlotWords);
              pool_unpack(q[0], q[1], q[2], e, rel_id, ev);
              alive = true;
            }
#endif
          } else {
            // pool full (rare): this lane searches its table at once, like traj_kernel does for every table event
            const float4 r4 = table_event_call<S>(P, e.E, ux, tot, uz, slot, flags);
            W = r4.x;
            om = r4.y;
            slot = __float_as_int(r4.z);
            flags = __float_as_int(r4.w);
            done = true;
            if (flags & MCD_EVF_NANCLAMP) atomicAdd(&s_cnt[kCntNan], 1u);
            if (flags & MCD_EVF_EXACT) atomicAdd(&s_cnt[kCntExact], 1u);
          }
        }
        __syncwarp();
      }
    }

    // ---- energy update (electron.py:44-59), retire (saveOutcomes before the rotation), update_direction ---------
    // (branch-free: W is 0 for a lane that completes no event in this iteration -- dead, parked or sitting out --
    // so its energy is unchanged, and the absorption test is masked by `done`)
    const float preE = e.E;
    const float En = preE - W;
    const bool absorbed = done && (En <= P.Emin);
    const bool dbl = absorbed && (outcome >= 0);    // reference records a second 'abs' row (Q4)
    outcome = (absorbed && outcome < 0) ? MCD_OUT_ABS : outcome;
    e.E = En;
    if (outcome >= 0) {
      retire_lane(O, s_cnt, outcome, dbl, preE, e, rel_id, ev + (done ? 1u : 0u));
      alive = false;
    }
    if (done) {
      rotate_dir(e.dx, e.dy, e.dz, e.yx, e.yy, e.yz, 1.0f - om, om, uphi);
      if (want_chan) atomicAdd(&s_cnt[kCntChan + slot_channel(P.m, slot)], 1u);
      ++ev;
    }
  }

#undef n_ready
#undef n_pend
#undef exhausted
  __syncthreads();
  flush_block_stats(O, s_cnt, nullptr, false, want_chan);
}

typedef void (*TrajKernel)(const DevParams, const DevOut, const unsigned long long, const unsigned long long);

template <bool SCATS, int POLY>
TrajKernel pick_kernel_tp(const DevParams& P) {
  switch (P.m.nshell) {
    case 1: return traj_kernel<1, MCD_MODE_DS, SCATS, POLY>;
    case 2: return traj_kernel<2, MCD_MODE_DS, SCATS, POLY>;
    case 3: return traj_kernel<3, MCD_MODE_DS, SCATS, POLY>;
    case 4: return traj_kernel<4, MCD_MODE_DS, SCATS, POLY>;
    case 5: return traj_kernel<5, MCD_MODE_DS, SCATS, POLY>;
    case 6: return traj_kernel<6, MCD_MODE_DS, SCATS, POLY>;
    case 7: return traj_kernel<7, MCD_MODE_DS, SCATS, POLY>;
    case 8: return traj_kernel<8, MCD_MODE_DS, SCATS, POLY>;
    default: return nullptr;
  }
}

template <bool SCATS>
TrajKernel pick_kernel_t(const DevParams& P) {
  if (P.mode == MCD_MODE_CONT) return traj_kernel<1, MCD_MODE_CONT, SCATS, 0>;
  return P.use_sig_poly ? pick_kernel_tp<SCATS, 1>(P) : pick_kernel_tp<SCATS, 0>(P);
}

template <bool RESUME>
TrajKernel pick_kernel_pool(const DevParams& P) {
  switch (P.m.nshell) {
    case 1: return traj_kernel_pool<1, RESUME>;
    case 2: return traj_kernel_pool<2, RESUME>;
    case 3: return traj_kernel_pool<3, RESUME>;
    case 4: return traj_kernel_pool<4, RESUME>;
    case 5: return traj_kernel_pool<5, RESUME>;
    case 6: return traj_kernel_pool<6, RESUME>;
    case 7: return traj_kernel_pool<7, RESUME>;
    case 8: return traj_kernel_pool<8, RESUME>;
    default: return nullptr;
  }
}

TrajKernel pick_kernel(const DevParams& P, bool scats) {
  if (scats) return pick_kernel_t<true>(P);
  // statistics-only DS runs use the park-pool kernel; it is compiled for the per-cell cross-section records only
  // (runs whose records could not be built -- channel order not static on [Emin, E0] -- take the plain kernel)
  if (P.mode == MCD_MODE_DS && P.use_sig_poly && !P.tune_no_pool) return pick_kernel_pool<false>(P);
  return pick_kernel_t<false>(P);
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

// Regrouping of the appended rows (mcd_order_scats / mcd_order_scats_columns): every stored row of the chunk pool goes
// to position d = offsets[rel_id] + event of the output.
//
// One block per chunk.  The chunk (<= 1024 rows, contiguous) is staged in shared memory with coalesced loads, then
// read back TRANSPOSED: a warp of the trajectory kernel appends the rows of one loop iteration side by side (lane
// order) and all 32 lanes are alive until the ids run out, so rows i, i+32, i+64, ... of a chunk are consecutive
// events of ONE electron and land in consecutive output positions.  Lane t of a warp therefore takes row t*32 + r:
// the 32 stores of one instruction hit one or two 128-byte lines (columns) or one contiguous run of rows (row-major)
// instead of 32 different places (the row-at-a-time version wrote 2.6x its payload and read 3.2x).
// Nothing depends on that structure for correctness: every row still computes its own destination.  Shared-memory
// rows are skewed by one word per 32 rows so that the transposed reads (stride 32 rows) fall on 32 different banks.
//
// COLS = false: packed rows in trajectory order.
// COLS = true: the payload words 2.. of every row go to one float column each (column j = word 2+j at byte offset
// j * col_stride) and the channel index to a byte column.  Words 0/1 (electron id, event index) are NOT stored: in
// trajectory order they follow from the offsets.
// 16-byte asynchronous global -> shared copy (LDGSTS, L2 only: the chunk is read once)
__device__ __forceinline__ void cp_async16(uint32_t* smem_dst, const uint32_t* gmem_src) {
  const unsigned int d = (unsigned int)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"(d), "l"(gmem_src) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

constexpr unsigned int kStageSkew = 4;                     // words of skew per 32 rows (keeps 16-byte units aligned)
__host__ __device__ constexpr unsigned int stage_words(unsigned int rw) { return MCD_CHUNK_ROWS * rw + 32u * kStageSkew; }

// Stage chunk `c` into `buf`: 16-byte units, unit u holds words 4u .. 4u+3 of the chunk; 32 rows = 8 rw units, so a
// unit never straddles a 32-row group and the group's skew moves it as a whole.
__device__ __forceinline__ unsigned int stage_chunk(uint32_t* buf, const uint32_t* __restrict__ pool,
                                                    const unsigned int* __restrict__ chunk_count, unsigned int c,
                                                    unsigned int rw) {
  const unsigned int cnt = min(chunk_count[c], MCD_CHUNK_ROWS);
  const uint32_t* cbase = pool + (size_t)c * MCD_CHUNK_ROWS * (size_t)rw;
  const unsigned int nunits = (cnt * rw + 3u) >> 2;          // (a partial last unit reads inside the chunk's capacity)
  const unsigned int upg = 8u * rw;                          // units per 32-row group
  for (unsigned int u = threadIdx.x; u < nunits; u += 256u)
    cp_async16(buf + 4u * u + kStageSkew * (u / upg), cbase + 4u * u);
  cp_async_commit();
  return cnt;
}

// COLS = true: the payload words 2.. of every row go to one float column each (column j = word 2+j at byte offset
// j * col_stride) and the channel index to a byte column.  Words 0/1 (electron id, event index) are NOT stored: in
// trajectory order they follow from the offsets.
// Double-buffered: while the block scatters chunk k out of one shared-memory buffer, the asynchronous copies of its
// next chunk fill the other one.
template <bool COLS>
__global__ void __launch_bounds__(256) scatter_chunks(const uint32_t* __restrict__ pool,
                                                      const unsigned int* __restrict__ chunk_count, unsigned int nchunks,
                                                      int row_words, const unsigned long long* __restrict__ offsets,
                                                      unsigned char* __restrict__ out, unsigned long long col_stride,
                                                      unsigned long long max_rows,
                                                      const unsigned int* __restrict__ nchunks_dev) {
  extern __shared__ __align__(16) uint32_t s_rows[];       // 2 buffers of stage_words(row_words)
  // max_rows: a run that dropped rows (pool overflow; the caller re-runs it) has event counts beyond the stored
  // rows -- such targets are skipped instead of written past the output
  // enqueued right behind the trajectory kernel (no host round trip): the number of used chunks is still on the device
  if (nchunks_dev) nchunks = min(nchunks, *nchunks_dev);
  unsigned char* type_col = out + (unsigned long long)(row_words - 2) * col_stride;
  uint32_t* sorted = reinterpret_cast<uint32_t*>(out);
  const unsigned int lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  const unsigned int rw = (unsigned int)row_words;
  const unsigned int bufw = stage_words(rw);
  unsigned int c = blockIdx.x;
  if (c >= nchunks) return;
  unsigned int b = 0u;
  unsigned int cnt = stage_chunk(s_rows, pool, chunk_count, c, rw);
  while (true) {
    const unsigned int cn = c + gridDim.x;
    unsigned int cnt_next = 0u;
    if (cn < nchunks) {
      cnt_next = stage_chunk(s_rows + (b ^ 1u) * bufw, pool, chunk_count, cn, rw);
      cp_async_wait<1>();                                  // chunk c has landed, chunk cn is in flight
    } else {
      cp_async_wait<0>();
    }
    __syncthreads();
    const uint32_t* buf = s_rows + b * bufw;
    for (unsigned int r = warp; r < 32u; r += 8u) {
      const unsigned int i = lane * 32u + r;
      if (i >= cnt) continue;
      const uint32_t* src = buf + i * rw + kStageSkew * lane;          // (i >> 5 == lane)
      const uint32_t rel = src[0], w1 = src[1];
      const unsigned long long d = offsets[rel] + (unsigned long long)(w1 & 0xffffffu);
      if (d >= max_rows) continue;
      if (COLS) {
        type_col[d] = (unsigned char)(w1 >> 24);
        for (unsigned int w = 2; w < rw; ++w)
          reinterpret_cast<uint32_t*>(out + (unsigned long long)(w - 2) * col_stride)[d] = src[w];
      } else {
        uint32_t* dst = sorted + d * (unsigned long long)rw;
        for (unsigned int w = 0; w < rw; ++w) dst[w] = src[w];
      }
    }
    if (cn >= nchunks) break;
    __syncthreads();                                       // buffer b may be refilled in the next round
    c = cn;
    cnt = cnt_next;
    b ^= 1u;
  }
}

}  // namespace

cudaError_t mcd_launch_order_scats(const unsigned int* nev, unsigned long long count, unsigned long long* tile_tmp,
                                   unsigned long long* offsets, unsigned long long* total_dev, const uint32_t* pool,
                                   const unsigned int* chunk_count, unsigned int nchunks, int row_words,
                                   uint32_t* sorted, unsigned long long col_stride, unsigned long long max_rows,
                                   const unsigned int* nchunks_dev, int sm_count, cudaStream_t stream) {
  const unsigned long long per_tile = (unsigned long long)kScanT * kScanItems;
  const unsigned int ntiles = (unsigned int)((count + 1 + per_tile - 1) / per_tile);
  scan_tile_sums<<<ntiles, kScanT, 0, stream>>>(nev, count, tile_tmp);
  scan_tiles<<<1, 1024, 0, stream>>>(tile_tmp, ntiles, total_dev);
  scan_apply<<<ntiles, kScanT, 0, stream>>>(nev, count, tile_tmp, offsets);
  if (sorted && nchunks) {
    const unsigned int grid = nchunks < (unsigned int)(sm_count * 64) ? nchunks : (unsigned int)(sm_count * 64);
    const size_t smem = 2 * (size_t)stage_words((unsigned int)row_words) * sizeof(uint32_t);      // 25 .. 116 KB
    auto kern = col_stride ? scatter_chunks<true> : scatter_chunks<false>;
    if (smem > 48 * 1024) {
      cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
      if (e != cudaSuccess) return e;
    }
    kern<<<grid, 256, smem, stream>>>(pool, chunk_count, nchunks, row_words, offsets, reinterpret_cast<unsigned char*>(sorted),
                                      col_stride, max_rows, nchunks_dev);
  }
  return cudaGetLastError();
}

unsigned long long mcd_order_tiles(unsigned long long count) {
  const unsigned long long per_tile = (unsigned long long)kScanT * kScanItems;
  return (count + 1 + per_tile - 1) / per_tile;
}

cudaError_t mcd_traj_max_blocks(const DevParams& P, bool scats, int sm_count, TrajLaunchInfo* info) {
  TrajKernel k = pick_kernel(P, scats);
  if (!k) return cudaErrorInvalidValue;
  const int threads = mcd_traj_uses_queue(P, scats) ? kPoolThreads : kThreads;     // (queue <=> park-pool kernel)
  int per_sm = 0;
  cudaError_t err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k, threads, 0);
  if (err != cudaSuccess) return err;
  cudaFuncAttributes fa;
  err = cudaFuncGetAttributes(&fa, k);
  if (err != cudaSuccess) return err;
  if (per_sm < 1) per_sm = 1;
  if (P.tune_blocks_per_sm >= 1 && P.tune_blocks_per_sm < per_sm) per_sm = P.tune_blocks_per_sm;   // (mcd_tuning)
  // shared-memory carve-out: exactly what the resident blocks need (+1 KB per block that the system reserves), the
  // rest of the 256 KB stays L1 for the cross-section cells and the table windows
  const size_t need = (size_t)per_sm * (fa.sharedSizeBytes + 1024);
  int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
  if (pct > 100) pct = 100;
  cudaFuncSetAttribute(k, cudaFuncAttributePreferredSharedMemoryCarveout, pct);
  info->blocks = per_sm * sm_count;
  info->threads = threads;
  info->regs = fa.numRegs;
  info->blocks_per_sm = per_sm;
  info->launches = 1;
  return cudaSuccess;
}

int mcd_traj_uses_queue(const DevParams& P, bool scats) {
  return (!scats && P.mode == MCD_MODE_DS && P.use_sig_poly && !P.tune_no_pool) ? 1 : 0;
}

cudaError_t mcd_launch_traj(const DevParams& P, const DevOut& O_in, unsigned long long first_id,
                            unsigned long long count, int sm_count, cudaStream_t stream_in, TrajLaunchInfo* info,
                            const TrajTailStream* tail) {
  cudaStream_t stream = stream_in;
  const bool scats = (P.scat_row_words > 0) && (O_in.scats != nullptr);
  TrajKernel k = pick_kernel(P, scats);
  if (!k) return cudaErrorInvalidValue;
  TrajLaunchInfo li;
  cudaError_t err = mcd_traj_max_blocks(P, scats, sm_count, &li);
  if (err != cudaSuccess) return err;
  DevOut O = O_in;
  // tail hand-over (park-pool kernel only): pass 0 runs the fresh ids and dumps what its warps still hold when they
  // are down to dump_T electrons; passes 1..n-1 run the leftover queue of the pass before; the last pass does not
  // dump.  mcd_tuning: handover_passes = 1 (or threshold 0): one launch.
  int passes = 1;
  int dump_T = 0;
  if (mcd_traj_uses_queue(P, scats) && O_in.rq[0] && P.tune_dump_T > 0 && P.tune_passes > 1) {
    passes = P.tune_passes > 4 ? 4 : P.tune_passes;
    dump_T = P.tune_dump_T > 31 ? 31 : P.tune_dump_T;
  }
  li.launches = passes;
  if (info) *info = li;
  for (int p = 0; p < passes; ++p) {
    const bool last = (p == passes - 1);
    O.ticket = O_in.ticket + p;
    O.rq_in = p ? O_in.rq[(p - 1) & 1] : nullptr;
    O.rq_in_count = p ? O_in.rq_counts + (p - 1) : nullptr;
    O.rq_out = last ? nullptr : O_in.rq[p & 1];
    O.rq_out_count = last ? nullptr : O_in.rq_counts + p;
    O.dump_T = last ? -1 : dump_T;
    if (p == 0) {
      k<<<li.blocks, li.threads, 0, stream>>>(P, O, first_id, count);
      // Several runs in flight on several streams: the hand-over passes of THIS run become eligible only when its
      // first pass ends, by which time the first passes of the runs launched after it are already waiting for SM
      // slots and would be served first.  On a high-priority side stream the (short, sparse) tail of a run gets the
      // next free slots instead, so a run completes one main pass after it started rather than `depth` of them.
      if (passes > 1 && tail && tail->stream && P.tune_tail_priority) {
        if ((err = cudaEventRecord(tail->first_done, stream)) != cudaSuccess) return err;
        if ((err = cudaStreamWaitEvent(tail->stream, tail->first_done, 0)) != cudaSuccess) return err;
        stream = tail->stream;
      }
    } else {
      TrajKernel kr = pick_kernel_pool<true>(P);
      // same shared-memory carve-out as the first pass
      cudaFuncAttributes fa;
      if (cudaFuncGetAttributes(&fa, kr) == cudaSuccess) {
        const size_t need = (size_t)li.blocks_per_sm * (fa.sharedSizeBytes + 1024);
        int pct = (int)((need * 100 + 228 * 1024 - 1) / (228 * 1024));
        cudaFuncSetAttribute(kr, cudaFuncAttributePreferredSharedMemoryCarveout, pct > 100 ? 100 : pct);
      }
      kr<<<li.blocks, li.threads, 0, stream>>>(P, O, first_id, count);
    }
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
  }
  if (P.flags & MCD_FLAG_HISTOGRAMS) {
    // the backscatter histograms: one pass over the els rows this run wrote
    unsigned long long nb = (count + 256ull * 8ull - 1ull) / (256ull * 8ull);
    const unsigned long long cap = (unsigned long long)sm_count * 8ull;
    if (nb > cap) nb = cap;
    hist_kernel<<<(unsigned int)(nb ? nb : 1), 256, 0, stream>>>(P.hist, O_in.els, count, O_in.stats);
    err = cudaGetLastError();
    if (err != cudaSuccess) return err;
  }
  if (stream != stream_in) {                               // back to the caller's stream
    if ((err = cudaEventRecord(tail->tail_done, stream)) != cudaSuccess) return err;
    if ((err = cudaStreamWaitEvent(stream_in, tail->tail_done, 0)) != cudaSuccess) return err;
  }
  return cudaSuccess;
}

cudaError_t mcd_launch_stats_merge(unsigned long long* dst, unsigned long long* src, int n, cudaStream_t stream) {
  stats_merge_kernel<<<(n + 255) / 256, 256, 0, stream>>>(dst, src, n);
  return cudaGetLastError();
}
