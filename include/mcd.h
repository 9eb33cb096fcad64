/*
 * mcd.h -- C ABI of libmcd.so: the B200 (sm_100a) implementation of MC-discrete's per-electron
 * trajectory loop.  Plain pointers and sizes only; no C++/torch types cross this boundary.
 *
 * What each group replaces in the reference (paths relative to the reference checkout):
 *   mcd_create / mcd_destroy        the bound arguments of the worker partial: inputPar dict + material
 *                                   (MC/doScatter.py:229-233, MC/tools/material.py:117-146)
 *   mcd_table_*                     probTable.generate/computeBlock/mapToMemory/readFromMemory
 *                                   (MC/tools/probTables.py:415-535) -- rows are built ON DEVICE
 *   mcd_run                         multiTraj_DS / multiTraj_cont job loop + trajectory_DS /
 *                                   trajectory_cont_cl (MC/tools/multiScatter.py:66-116,171-229;
 *                                   MC/tools/singleScatter.py:10-41,205-239)
 *   mcd_drain_els / mcd_drain_scats thingsToSave.addToList + pickle-to-parent + listener
 *                                   (MC/tools/fileTools.py:193-242, MC/doScatter.py:99-119)
 *   mcd_hist_* / mcd_counters       no reference equivalent (reference keeps raw rows only)
 *   mcd_test_*                      per-formula / per-step hooks fed injected uniforms, for parity tests
 *                                   against MC/tools/scattering.py, crossSections.py, stoppingPowers.py,
 *                                   rotation.py, electron.py
 *
 * Conventions
 *   - every function returns 0 (MCD_OK) or a negative mcd_status; mcd_last_error() gives the text.
 *   - one mcd_ctx <-> one CUDA device; a ctx is thread-affine and not re-entrant.
 *   - `stream` arguments are cudaStream_t handles passed as void* (0 = legacy default stream).
 *   - host output buffers are caller-owned; the library never frees caller memory.
 *   - there is NO CPU fallback: without a usable CUDA device every compute entry point fails.
 */
#ifndef MCD_H_
#define MCD_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MCD_ABI_VERSION 1
#define MCD_MAX_SHELLS 8
#define MCD_MAX_CHANNELS (MCD_MAX_SHELLS + 3)

typedef enum {
  MCD_OK = 0,
  MCD_ERR_INVALID = -1,      /* bad argument / unsupported configuration */
  MCD_ERR_CUDA = -2,         /* CUDA runtime error (text in mcd_last_error) */
  MCD_ERR_NO_DEVICE = -3,    /* no usable CUDA device: there is no CPU fallback */
  MCD_ERR_STATE = -4,        /* call sequence error (e.g. DS run without tables) */
  MCD_ERR_OVERFLOW = -5,     /* output buffer too small (rows were dropped and counted) */
  MCD_ERR_NOMEM = -6
} mcd_status;

/* mode / model enums (input.file keys `mode`, `Bethe_model`, `elastic`; MC/tools/fileTools.py:27) */
enum { MCD_MODE_DS = 0, MCD_MODE_CONT = 1 };
enum { MCD_BETHE_CLASSICAL = 0, MCD_BETHE_JL = 1, MCD_BETHE_EXPLICIT = 2 };
/* elastic cross-section EVALUATED for the 'Ruth' channel (crossSections.py:37-58, 63-84, 89-112, 115-143).  The
 * reference's label -> function rules differ between DS (scattering.py:483-491) and continuous mode
 * (scattering.py:781-789, where the _wDefl branches are unreachable); the host mirror applies them
 * (mcdiscrete_b200/engine.py::elastic_kind).  The polar angle is Rutherford_halfPol with the vanilla screening
 * parameter for every model (scattering.py:575-577, :829). */
enum { MCD_ELASTIC_RUTH_VANILLA = 0, MCD_ELASTIC_RUTH_VANILLA_WDEFL = 1, MCD_ELASTIC_RUTH_NIGRAM = 2,
       MCD_ELASTIC_RUTH_NIGRAM_WDEFL = 3,
       /* elastic = 'Mott' (DS mode only; the reference's continuous classes crash on it): the elastic channel is
        * labelled 'Mott', its cross-section is a step function of E read from the table (scattering.py:493) and
        * its polar angle comes from pickMottTable (scattering.py:96-122, :588-591).  Needs mcd_mott_upload. */
       MCD_ELASTIC_MOTT = 4 };
/* tables['Gryz'][0] is read for EVERY shell in the reference (scattering.py:625) */
enum { MCD_GRYZ_OUTERMOST = 0, MCD_GRYZ_PER_SHELL = 1 };
/* table slots: 0 = Moller, 1 + k = Gryzinski table k in the reference's tables['Gryz'] order
 * (reversed name_s, probTables.py:579), so slot 1 is the outermost shell. */
enum { MCD_TABLE_MOLLER = 0, MCD_TABLE_GRYZ0 = 1 };
#define MCD_MAX_TABLES (1 + MCD_MAX_SHELLS)

/* outcome codes (electron.py:52,69,77,80; singleScatter.py:21) */
enum { MCD_OUT_ABS = 0, MCD_OUT_BKS = 1, MCD_OUT_TRSM = 2, MCD_OUT_TRDF = 3, MCD_OUT_LONG = 4 };

/* electron_output columns (fileTools.py:200-212) */
#define MCD_EL_OUTCOME   (1u << 0)
#define MCD_EL_FINAL_E   (1u << 1)
#define MCD_EL_FINAL_DIR (1u << 2)
#define MCD_EL_LAST_POS  (1u << 3)
/* scatter_output columns (scattering.py:527-676, electron.py:59,99) */
#define MCD_SC_TYPE      (1u << 0)
#define MCD_SC_POSITION  (1u << 1)   /* x,y,z  pre-step vertex (DS only) */
#define MCD_SC_PATHL     (1u << 2)
#define MCD_SC_E         (1u << 3)   /* pre-loss energy (DS only) */
#define MCD_SC_E_LOSS    (1u << 4)
#define MCD_SC_POL_ANGLE (1u << 5)   /* cos^2(theta/2) */
#define MCD_SC_AZ_ANGLE  (1u << 6)   /* halfPhi */
#define MCD_SC_ENERGY    (1u << 7)   /* post-loss energy */
#define MCD_SC_DIRECTION (1u << 8)   /* post-rotation dx,dy,dz */

/* flags */
#define MCD_FLAG_HISTOGRAMS     (1u << 0)  /* fused backscatter histograms (energy/polar/azimuth/depth) */
#define MCD_FLAG_CHANNEL_COUNTS (1u << 1)  /* count events per channel */

/* Material constants, computed by the host in float64 exactly as MC/tools/material.py:15-88,151-302. */
typedef struct {
  int32_t Z;
  int32_t n_shells;                       /* core shells, in name_s order (sigma-dict insertion order) */
  double n_val, E_val;                    /* valence electrons / energy */
  double atnd;                            /* atoms per m^3        (material.py:30) */
  double fermi_e, plasmon_e;              /* eV                   (material.py:56,88) */
  double bethe_k;                         /* Joy-Luo k            (material.py:183,220,260,297) */
  double shell_E[MCD_MAX_SHELLS];         /* binding energies eV */
  double shell_n[MCD_MAX_SHELLS];         /* occupancies */
} mcd_material;

/* input.file parameters (inputFile/input.file:3-39, fileTools.py:8-46). */
typedef struct {
  double E0, Emin;                        /* eV */
  double s_tilt_deg;                      /* dir0 = (-sin t, 0, cos t)  (multiScatter.py:97-100) */
  double thickness;                       /* Angstrom */
  double Wc;                              /* eV, Moller cut-off */
  double tol_E, tol_W;                    /* recorded only; the axis search runs on the host */
  int64_t maxScatt;
  int32_t mode, bethe, elastic, gryz_table_mode;
  uint32_t el_mask, scat_mask, flags;
  uint32_t _pad;
  uint64_t seed;                          /* Philox4x32-10 key; counter = (electron id, event index) */
} mcd_params;

/* One els row (32 bytes).  code: bits 0-7 first outcome, bit 8 = the reference would have recorded a
 * second 'abs' row for this electron (bks whose exit-event loss crossed Emin; singleScatter.py:32-41,
 * SURVEY Q4), bits 16-31 unused.  final_E/dir/last_pos are the saveOutcomes snapshot (electron.py:105-113). */
typedef struct {
  uint32_t code;
  float final_E;
  float dx, dy, dz;
  float last_x, last_y, last_z;
} mcd_el_row;

/* A scats row is a packed record of 32-bit words (see mcd_scat_layout):
 *   word 0: electron id relative to first_id of the run          (u32)
 *   word 1: event index (bits 0-23) | channel (bits 24-31)        (u32)
 *           channel = index in sigma insertion order: 0 Ruth, 1 Moller, 2..1+S Gryz<name_s[k]>, 2+S Quinn
 *   then one float per requested column, in MCD_SC_* bit order (POSITION and DIRECTION take 3 words). */
typedef struct {
  uint32_t row_words;                     /* words per row */
  int32_t word_of[16];                    /* first word of column bit b, or -1 when not requested */
} mcd_scat_layout;

typedef struct {
  uint64_t electrons_done;
  uint64_t events;                        /* executed loop iterations == scats rows (incl. dropped) */
  uint64_t outcome[5];                    /* by MCD_OUT_* (first outcome) */
  uint64_t double_records;                /* SURVEY Q4 occurrences */
  uint64_t scat_rows_stored, scat_rows_dropped;
  uint64_t nan_clamped;                   /* table W > (E-Ef)/2 (binaryCollModel NaN) */
  uint64_t select_exact;                  /* channel picks resolved by the exact float64 path */
  uint64_t channel[MCD_MAX_CHANNELS];     /* only with MCD_FLAG_CHANNEL_COUNTS */
} mcd_counters;

/* histogram block: int64 counts, laid out [energy | polar | azimuth | depth | reserved] */
#define MCD_HIST_E_BINS 512
#define MCD_HIST_POLAR_BINS 180
#define MCD_HIST_AZ_BINS 360
#define MCD_HIST_DEPTH_BINS 512
#define MCD_HIST_TOTAL (MCD_HIST_E_BINS + MCD_HIST_POLAR_BINS + MCD_HIST_AZ_BINS + MCD_HIST_DEPTH_BINS)

typedef struct mcd_ctx mcd_ctx;

int mcd_abi_version(void);
const char* mcd_last_error(const mcd_ctx* ctx);           /* ctx may be NULL: last error of this thread */
int mcd_device_count(int* n);
int mcd_device_info(int device, char* name, int name_len, int* sm_count, int* cc_major, int* cc_minor,
                    uint64_t* total_mem);

int mcd_create(const mcd_params* p, const mcd_material* m, int device, mcd_ctx** out);
int mcd_destroy(mcd_ctx* ctx);
/* A second context on the same device that SHARES the parent's tables (read-only) but owns its own output
 * buffers, ticket and statistics block: two contexts on two streams let the host drain batch k while batch k+1
 * runs, and let the next kernel fill the SMs that the previous kernel's tail leaves idle.  Destroy clones
 * before their parent; tables built/uploaded on the parent after cloning are not seen by the clone. */
int mcd_clone(const mcd_ctx* parent, mcd_ctx** out);

/* Launch tuning of a context (defaults are the measured optimum on B200; the A/B tool scripts/ab_kernel.py
 * presets them through MCD_PASSES / MCD_DUMP_T / MCD_BLOCKS_PER_SM / MCD_NO_POOL).  Results never depend on them.
 * Set before mcd_set_capacity; clones inherit the parent's values at mcd_clone time. */
typedef struct mcd_tuning {
  int32_t handover_passes;     /* launches per statistics-only DS run: 1 = no tail hand-over, default 3, at most 4 */
  int32_t handover_threshold;  /* a warp hands its electrons to the next pass when it holds <= this many (0..31, 31) */
  int32_t blocks_per_sm;       /* resident blocks per SM; 0 = as many as fit (default) */
  int32_t statistics_kernel;   /* 0 = park-pool kernel for statistics-only DS runs (default), 1 = the plain kernel */
  int32_t tail_priority;       /* 1 = hand-over passes on a high-priority side stream (default), 0 = on the run's stream */
} mcd_tuning;
int mcd_tuning_get(const mcd_ctx* ctx, mcd_tuning* out);
int mcd_tuning_set(mcd_ctx* ctx, const mcd_tuning* t);

/* ---- tables -------------------------------------------------------------------------- */
/* Build table `slot` on the device from host axes (float32, as found by the host axis search that
 * mirrors probTables.py:280-411).  shell_E/shell_n are ignored for slot 0 (Moller uses Wc/n_val). */
int mcd_table_build(mcd_ctx* ctx, int slot, const float* Es, uint32_t nE, const float* Ws, uint32_t nW,
                    double shell_E, double shell_n, void* stream);
/* Upload a table computed elsewhere (e.g. the reference's *.table memmap, probTables.py:490-535):
 * cdf is nE x (nW-1) float32 row-major, zeros = masked. */
int mcd_table_upload(mcd_ctx* ctx, int slot, const float* Es, uint32_t nE, const float* Ws, uint32_t nW,
                     const float* cdf);
/* The Mott table of `elastic = Mott` (mottTable, extFunctions.py:21-153, as genTables builds it,
 * probTables.py:601-613): Es[nE] ascending in eV, sigmas[nE] in cm^2, cdf = probTable nE x nT row-major (cumulative
 * sums over the angle grid divided by the row sum, float64 like the reference keeps it), thetas[nT] in degrees.
 * Replaces: tables['Mott'] read at scattering.py:466-467, :493 and :591.  E0 must not exceed Es[nE-1] (the
 * reference raises IndexError there).  Only valid on a context created with MCD_ELASTIC_MOTT. */
int mcd_mott_upload(mcd_ctx* ctx, uint32_t nE, uint32_t nT, const double* Es, const double* sigmas, const double* cdf,
                    const double* thetas);
int mcd_table_shape(const mcd_ctx* ctx, int slot, uint32_t* nE, uint32_t* nW);
int mcd_table_download(mcd_ctx* ctx, int slot, float* cdf_out);   /* nE x (nW-1) float32 */

/* ---- runs ---------------------------------------------------------------------------- */
/* Size the device output buffers: one els row per electron of a run, and `max_scat_rows` scats rows
 * (0 = no scats recording).  Must be called before mcd_run. */
int mcd_set_capacity(mcd_ctx* ctx, uint64_t max_electrons, uint64_t max_scat_rows);
int mcd_scat_layout_get(const mcd_ctx* ctx, mcd_scat_layout* out);
/* Clear histograms and counters (they otherwise accumulate over runs). */
int mcd_reset_stats(mcd_ctx* ctx, void* stream);
/* Simulate electrons [first_id, first_id+count) asynchronously on `stream`.  Row buffers are
 * overwritten by each run; drain before the next one. */
int mcd_run(mcd_ctx* ctx, uint64_t first_id, uint64_t count, void* stream);
int mcd_sync(mcd_ctx* ctx, void* stream);
/* After sync: number of els rows (== count of the last run) / stored scats rows. */
int mcd_row_counts(mcd_ctx* ctx, uint64_t* n_els, uint64_t* n_scats);
/* Copy els rows [offset, offset+n) of the last run to host (row i = electron first_id+i). */
int mcd_drain_els(mcd_ctx* ctx, mcd_el_row* host, uint64_t offset, uint64_t n, void* stream);
/* Copy scats rows [offset, offset+n) (append order; use words 0/1 to regroup) to host. */
int mcd_drain_scats(mcd_ctx* ctx, uint32_t* host_words, uint64_t offset, uint64_t n, void* stream);
/* Regroup the stored scats rows of the last run ON THE DEVICE into the reference's row order (electron after
 * electron in id order, events in order; multiScatter.py:107-112): a scan of the per-electron event counts gives
 * row offsets, one pass scatters the appended rows to offsets[id] + event.  Needs a second row buffer of the same
 * size and no dropped rows.  Afterwards mcd_drain_scats_ordered copies rows [offset, offset+n) of the ordered
 * array and mcd_scat_offsets copies offsets[first .. first+n] (n+1 values; electron i owns rows
 * offsets[i] .. offsets[i+1]-1). */
int mcd_order_scats(mcd_ctx* ctx, void* stream);
int mcd_drain_scats_ordered(mcd_ctx* ctx, uint32_t* host_words, uint64_t offset, uint64_t n, void* stream);
int mcd_scat_offsets(mcd_ctx* ctx, uint64_t* host, uint64_t first, uint64_t n, void* stream);
/* Columnar variant of the same regrouping (what the reference's listener finally stores is columns:
 * pd.DataFrame.from_dict(results_d[key]), doScatter.py:112-117): every payload word of a row becomes one float32
 * column and the channel index one uint8 column, all in trajectory order; the electron id and event index words
 * are not stored (they follow from mcd_scat_offsets), so a `type, position, energy` row costs 17 bytes of D2H
 * traffic instead of 24.  Layout of the block copied by mcd_drain_scats_columns: float column j (= row word 2+j of
 * mcd_scat_layout) at byte offset j * col_stride, the uint8 channel column at type_offset; n_rows entries each. */
int mcd_order_scats_columns(mcd_ctx* ctx, void* stream);
/* The same regrouping ENQUEUED on `stream` right behind mcd_run, with no host synchronisation (the used chunk count
 * is read on the device, the columns are laid out for the capacity of the chunk pool): in a pipelined run the
 * regrouping of batch k+1 then overlaps the D2H drain of batch k.  Synchronise the stream before calling
 * mcd_scat_columns_layout / mcd_drain_scats_columns / mcd_scat_offsets, which verify the result (no dropped rows,
 * event counts == stored rows) on first use. */
int mcd_order_scats_columns_enqueue(mcd_ctx* ctx, void* stream);
int mcd_scat_columns_layout(mcd_ctx* ctx, uint64_t* n_rows, uint64_t* n_float_cols, uint64_t* col_stride,
                            uint64_t* type_offset, uint64_t* total_bytes);
int mcd_drain_scats_columns(mcd_ctx* ctx, void* host, uint64_t capacity_bytes, void* stream);
int mcd_counters_get(mcd_ctx* ctx, mcd_counters* out);
int mcd_hist_get(mcd_ctx* ctx, int64_t* host /* MCD_HIST_TOTAL */);
/* Device address of the int64 histogram||counter block, for an in-place NCCL all-reduce by the caller
 * (the only collective of the path).  n_int64 = MCD_HIST_TOTAL + sizeof(mcd_counters)/8. */
int mcd_stats_device_ptr(mcd_ctx* ctx, void** dev_ptr, uint64_t* n_int64);
/* dst's histogram||counter block += src's, and src's block is cleared, on `stream` (which must be ordered after the
 * runs of both contexts): the contexts of a pipeline (mcd_clone) are merged into one block before the single
 * all-reduce of a multi-GPU run. */
int mcd_stats_merge(mcd_ctx* dst, mcd_ctx* src, void* stream);
/* Histogram axis ranges: energy [Emin,E0], polar [0,90] deg, azimuth [-180,180) deg, depth [0,depth_max] A. */
int mcd_hist_ranges(const mcd_ctx* ctx, double* e_lo, double* e_hi, double* depth_max);
/* Timing of the last mcd_run's trajectory kernel in ms (CUDA events on the launch stream). */
int mcd_last_kernel_ms(mcd_ctx* ctx, float* ms);
/* Launch geometry of the trajectory kernel (blocks, threads, registers/thread, blocks/SM). */
int mcd_launch_info(mcd_ctx* ctx, int* blocks, int* threads, int* regs, int* blocks_per_sm);
/* Kernel launches the last mcd_run made: 1, or the passes of the tail hand-over (statistics-only DS runs: warps
 * that are down to a few electrons pass them to a follow-up launch that runs them in full warps). */
int mcd_last_launches(mcd_ctx* ctx, int* launches);

/* ---- roofline denominators ---------------------------------------------------------------------- */
/* Measure the FP32 FMA peak (TFLOP/s, FMA = 2 flop) and the MUFU/XU peak (1e12 transcendental op/s) of
 * `device` with dependent-chain micro-kernels; sm_mhz_est = the SM clock implied by the MUFU rate. */
int mcd_measure_peaks(int device, float* fp32_tflops, float* mufu_tops, float* sm_mhz_est);

/* ---- unit hooks (injected uniforms; all arrays are HOST pointers, n vectors) --------------- */
/* sigma_out: n x (S+3) in sigma insertion order, cm^2; mfp_out: total mean free path in Angstrom. */
int mcd_test_sigmas(mcd_ctx* ctx, const float* E, uint64_t n, float* sigma_out, float* mfp_out);
/* The cumulative inverse mean free paths (1/Angstrom) the hot path uses for the channel pick and the path length,
 * in SLOT order (descending sigma: Ruth, Quinn, Moller, Gryzinski shells): cum_out is n x (S+3);
 * slot_channel_out[S+3] maps slots to sigma insertion indices; *uses_poly = 1 when they come from the per-cell
 * quadratics built at mcd_create (else from the formulas), *poly_err = their verified worst error / total. */
int mcd_test_cum(mcd_ctx* ctx, const float* E, uint64_t n, float* cum_out, int32_t* slot_channel_out,
                 int32_t* uses_poly, double* poly_err);
/* pickFromSigmas on given sigmas (n x nch, insertion order): pick_out = insertion index;
 * path_out (optional) = 0 fast float32 path, 1 exact float64 path. */
int mcd_test_select(mcd_ctx* ctx, const float* sigma, int nch, const float* u, uint64_t n,
                    int32_t* pick_out, int32_t* path_out);
int mcd_test_ruth(mcd_ctx* ctx, const float* E, const float* u, uint64_t n, float* c2_out, float* om_out);
int mcd_test_bincoll(mcd_ctx* ctx, const float* E, const float* W, uint64_t n, float* c2_out, float* om_out);
/* om = 1 - c2 carried separately (float32 cancellation); dirs are n x 3. */
int mcd_test_rotate(mcd_ctx* ctx, const float* d_in, const float* y_in, const float* c2, const float* om,
                    const float* u_phi, uint64_t n, float* d_out, float* y_out);
int mcd_test_stopping(mcd_ctx* ctx, int bethe, const float* E, uint64_t n, float* sp_out);
int mcd_test_lookup(mcd_ctx* ctx, int slot, const float* E, const float* u, uint64_t n, float* W_out,
                    int32_t* iE_out, int32_t* jW_out);
/* Mott hooks: sigma_out = sigmas[bisect_left(Es, E)] in cm^2 (scattering.py:493); c2_out = pickMottTable(E) with
 * the injected uniform u (scattering.py:96-122), iE_out / jT_out = the two bisect_left results. */
int mcd_test_mott(mcd_ctx* ctx, const float* E, const float* u, uint64_t n, float* sigma_out, float* c2_out,
                  int32_t* iE_out, int32_t* jT_out);
/* One loop iteration (after the maxScatt check).  state = [E, x,y,z, dx,dy,dz, yx,yy,yz] (n x 10);
 * u = n x 4 (DS: type,path,channel,phi; cont: path,polar,phi,unused);
 * row_out = n x 8: [channel, pathl, E_loss, c2, halfPhi, energy_after, outcome(-1 none), flags]. */
int mcd_test_step(mcd_ctx* ctx, const float* state_in, const float* u, uint64_t n, float* state_out,
                  float* row_out);
int mcd_test_philox(uint64_t seed, uint64_t el_id, uint32_t ev_idx, uint32_t raw_out[4], float u_out[4]);

#ifdef __cplusplus
}
#endif
#endif /* MCD_H_ */
