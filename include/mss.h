/* mss.h -- C-ABI of the B200-native sliding-window map-sparsification engine (libmss.so).
 *
 * The reference (fishmarch/MS-SLAM @ e4730ec) has no plugin/FFI layer for this path: MapSparsification::Sparsifying
 * (/root/reference/src/MapSparsification.cc:58-171) builds a GUROBI model object by object and calls
 * GRBModel::optimize() (:157).  This header is the boundary that replaces those GUROBI calls: the C++ class
 * ms_slam_b200/host/MapSparsification.{h,cc} keeps the reference's public surface
 * (/root/reference/include/MapSparsification.h:26-70) and calls the entry points below instead of
 *   GRBEnv / GRBEnv::start            (MapSparsification.cc:6,20)        -> mss_create
 *   GRBModel, addVar, addConstr, ...  (MapSparsification.cc:61-153)     -> the mss_window_view snapshot
 *   GRBModel::optimize                (MapSparsification.cc:154-157)    -> mss_solve / mss_solve_batch
 *   GRBVar::get(GRB_DoubleAttr_X)     (MapSparsification.cc:159-166)    -> mss_result.keep_bits
 *   ~GRBEnv                                                             -> mss_destroy
 *
 * Plain C, plain pointers and sizes; no C++/torch types cross this boundary.  All functions return 0 (MSS_OK) or a
 * negative mss_status and never throw.  A handle is single-threaded (one per sparsifier thread, like GRBEnv) and owns
 * one CUDA stream and all device scratch (grown on demand, never shrunk, no allocation on the hot call after warm-up).
 * There is no CPU fallback: without a usable CUDA device mss_create fails with MSS_E_CUDA.
 */
#ifndef MSS_H_
#define MSS_H_

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MSS_VERSION 210            /* 0.2.1: mp_tie (tie-breaks on the caller's ranks), wire flags (one-byte nObs table), dual bound, BoW;
                                      0.2.0: result_memory, persistent device mirror (mss_mirror_*), keyframe compaction,
                                      single-process multi-device handles; 0.1.3: packed transport layouts, components */

#define MSS_GRID_COLS 64           /* FRAME_GRID_COLS, /root/reference/include/Frame.h:45 */
#define MSS_GRID_ROWS 48           /* FRAME_GRID_ROWS, /root/reference/include/Frame.h:44 */
#define MSS_CELL_NONE 0xFFFFu      /* keypoint outside the grid (Frame::PosInGrid false, src/Frame.cc:657-668) */

typedef enum mss_status {
    MSS_OK = 0,
    MSS_E_BADARG = -1,       /* NULL / inconsistent sizes / out-of-range index found in a view */
    MSS_E_CUDA = -2,         /* CUDA runtime error (message in mss_last_error) */
    MSS_E_NCCL = -3,         /* NCCL error or NCCL library not loadable */
    MSS_E_NOMEM = -4,
    MSS_E_NOCONVERGE = -5,   /* round cap hit; the returned selection is still feasible (every row at its best level) */
    MSS_E_INTERNAL = -6
} mss_status;

typedef enum mss_memory {
    MSS_MEM_HOST = 0,        /* pointers are host memory (pageable or pinned); the engine stages them itself */
    MSS_MEM_DEVICE = 1       /* pointers are device memory on the engine's device; consumed in place */
} mss_memory;

/* How a view's arrays are encoded.  SOA is the plain form; PACKED is the transport form for views that travel over
 * PCIe every call (about 0.6x the bytes): FlattenWindow can emit either at the same host cost. */
/* Where the result arrays of a window live (mss_window_view::result_memory). */
typedef enum mss_result_memory {
    MSS_RESULT_SAME = 0,     /* same kind as the view's arrays (the default of a zero-initialised view) */
    MSS_RESULT_HOST = 1,     /* host buffers although the view is device-resident: "inputs in HBM, bitmask back to the host" */
    MSS_RESULT_DEVICE = 2    /* device buffers although the view travels from the host */
} mss_result_memory;

typedef enum mss_layout {
    MSS_LAYOUT_SOA = 0,      /* feat_mp i32 + feat_cell u16, mp_nobs i32, mp_obs_kf i32 */
    MSS_LAYOUT_PACKED = 1,   /* slots u32 = (map point << 12) | cell, mp_nobs16 u16, obs_pairs u32 = (map point << 12) | outside kf */
    MSS_LAYOUT_PACKED16 = 2  /* as PACKED, but the slots of a keyframe are sorted by map-point index and delta-coded as u16 tokens */
} mss_layout;
#define MSS_SLOT_CELL_NONE 0xFFFu      /* packed slot: keypoint outside the grid (MSS_CELL_NONE of the SOA form) */
#define MSS_SLOT_EMPTY 0xFFFFFFFFu     /* packed slot: empty slot / bad map point (feat_mp == -1 of the SOA form) */

typedef struct mss_handle mss_handle;

/* Solver parameters = the reference's yaml keys (MapSparsification.cc:9-12) + the device algorithm's caps. */
typedef struct mss_config {
    int32_t device;            /* CUDA device ordinal */
    int32_t min_points;        /* Sparsification.N           (mnMinNum) */
    float   lambda;            /* Sparsification.Lambda      (mfLambda: cost of one missing point in a keyframe row) */
    float   grid_lambda;       /* Sparsification.GridLambda  (mfGridLambda: cost of one uncovered occupied cell) */
    int32_t max_rounds;        /* cap on propagation+greedy rounds per window (0 -> 256) */
    int32_t all_rule_steps;    /* greedy steps that use the strict conflict-free rule before a deficient row may also nominate its
                                  top-deficit candidates on its own (0 = none, the default: rows nominate from the first greedy step) */
    int32_t max_drop_rounds;   /* cap on reverse-delete rounds (0 -> 16) */
    int32_t stall_den;         /* a propagation round that decides fewer than 1/stall_den of the undecided points is followed by a
                                  greedy step instead of another propagation round (0 -> 8, negative = never) */
} mss_config;

/* One window, flattened (SoA).  Snapshot of the pointer graph Sparsifying() walks:
 *   KeyFrame::mvpMapPoints (include/KeyFrame.h:298) + KeyFrame::mGrid (:307)   -> feat_ptr / feat_mp / feat_cell
 *   MapPoint::nObs (include/MapPoint.h:101), MapPoint::mObservations (:150)    -> mp_nobs / mp_obs_ptr / mp_obs_kf
 *   KeyFrame::GetNumberMPs() of keyframes outside the window (KeyFrame.cc:286) -> okf_total
 * KF table = K window keyframes then H outside keyframes.  Caller-owned, read-only, not retained past the call. */
typedef struct mss_window_view {
    int32_t K;                 /* window keyframes (vpKFs.size()) */
    int32_t H;                 /* outside keyframes observing at least one window map point */
    int32_t M;                 /* map-point table size */
    int32_t F;                 /* feature slots = feat_ptr[K] */
    int32_t O;                 /* observations  = mp_obs_ptr[M] */
    int32_t memory;            /* mss_memory of every pointer below */
    const int32_t*  feat_ptr;  /* [K+1] slot range of each window keyframe (a caller may leave empty slots out altogether) */
    const int32_t*  feat_mp;   /* [F]   map-point table index, -1 = empty slot or bad map point (MapSparsification.cc:70,90) */
    const uint16_t* feat_cell; /* [F]   col*48+row in the 64x48 grid, MSS_CELL_NONE = not in mGrid */
    const int32_t*  mp_nobs;   /* [M]   MapPoint::Observations() */
    const int32_t*  mp_obs_ptr;/* [M+1] */
    const int32_t*  mp_obs_kf; /* [O]   KF-table index of each observation; >= K means outside keyframe (value - K); entries < K
                                        (window keyframes) are ignored and may be left out, as may the lists of map points
                                        that are not variables */
    const int32_t*  okf_total; /* [H]   GetNumberMPs() of each outside keyframe */
    /* ---- MSS_LAYOUT_PACKED replaces feat_mp / feat_cell / mp_nobs / mp_obs_ptr / mp_obs_kf (those five may then be NULL);
     *      feat_ptr and okf_total are used as above.  Requires M <= 2^20, Observations() <= 65535, H <= 4095. ---- */
    int32_t layout;            /* mss_layout; 0 (SOA) for zero-initialised views */
    int32_t n_max_floor;       /* nMaxObservation (MapSparsification.cc:66-76) is at least this; 0 for a whole window.  A component
                                  of a larger window (mss_components) carries the window-wide nMax here, so that its costs are
                                  the ones the whole-window model would use */
    const uint32_t* slots;     /* [F]   (map-point table index << 12) | (col*48+row), low 12 bits MSS_SLOT_CELL_NONE = not in mGrid;
                                        MSS_SLOT_EMPTY = empty slot.  The order of the slots inside a keyframe is free (the
                                        result does not depend on it); sorted by value is the fastest */
    const uint16_t* mp_nobs16; /* [M]   MapPoint::Observations() */
    /* ---- MSS_LAYOUT_PACKED16: slots16 replaces slots; F counts TOKENS and feat_ptr holds token ranges.  The slots of a
     *      keyframe are sorted by map-point table index; a running index starts at 0 in every keyframe.  Token t:
     *        (t >> 12) < 15 : a slot: index += t >> 12; cell = t & 0xFFF (MSS_SLOT_CELL_NONE = not in mGrid)
     *        (t >> 12) == 15: no slot: index += 15 * ((t & 0xFFF) + 1)
     *      (about 1.03 tokens per slot for windows numbered in discovery order: 2.06 bytes per slot instead of 4) ---- */
    const uint16_t* slots16;   /* [F]   tokens */
    const uint32_t* obs_pairs; /* [O]   observations of the window's map points by OUTSIDE keyframes only, in any order:
                                        (map-point table index << 12) | j, j = 0..H-1 the outside keyframe (KF-table index K + j) */
    int32_t result_memory;     /* mss_result_memory of keep_bits / kf_cov / kf_slack of this window's mss_result */
    int32_t nobs8;             /* packed layouts: != 0 -> mp_nobs16 points to a uint8_t array [M] instead (valid when Observations() <= 255
                                  for every map point of the window: one byte less per map point on the wire) */
    const uint32_t* mp_tie;    /* [M] optional (NULL = off): tie-break rank of every map point, lower wins.  Wherever the selection has to
                                  choose between candidates of equal gain it prefers the lower TABLE INDEX by default, so the result
                                  depends on how the caller numbered the table; with mp_tie (e.g. the rank of MapPoint::mnId, SURVEY 8b
                                  "ties by (cost, MP gid)") it is the same for every numbering of the same window.  Same memory kind as
                                  the other arrays. */
} mss_window_view;

/* Result of one window.  keep_bits / kf_cov / kf_slack are caller-allocated (same mss_memory as the view unless the view's
 * result_memory says otherwise) or NULL. */
typedef struct mss_result {
    uint32_t* keep_bits;       /* [(M+31)/32] bit p = 1 keep map point p, 0 = SetBadFlag() it (MapSparsification.cc:162-165);
                                  map points that are not variables of the window are always 1 */
    int32_t*  kf_cov;          /* [K+H] points kept in each keyframe row (with multiplicity), may be NULL */
    int32_t*  kf_slack;        /* [K+H] max(0, need - cov): the ILP's t_k / u_j at their optimum, may be NULL */
    double    objective;       /* F(x) of SURVEY Appendix A.3 = sum c_p x_p + GridLambda*uncovered + Lambda*slack */
    double    dual_bound;      /* lower bound on the ILP optimum proven on device (NaN when not computed) */
    int64_t   sum_cost;        /* sum c_p over kept variables (costs are integers: nMax - nObs_p) */
    int32_t   uncovered_cells; /* occupied cells left without a kept point */
    int32_t   total_slack;     /* sum of kf_slack */
    int32_t   n_max;           /* nMaxObservation (MapSparsification.cc:66-76) */
    int32_t   n_vars;          /* distinct map points that are ILP variables */
    int32_t   n_cells;         /* occupied-valid (keyframe, cell) rows */
    int32_t   nnz;             /* valid grid-listed slots (incidences of the keyframe rows) */
    int32_t   n_kept;          /* variables kept */
    int32_t   rounds;          /* propagation + greedy + drop rounds executed */
    int32_t   status;          /* MSS_OK or MSS_E_NOCONVERGE / MSS_E_BADARG for this window */
    float     time_build_us;   /* device time: snapshot scan + outside-row assembly */
    float     time_solve_us;   /* device time: selection + evaluation + bit packing */
} mss_result;

typedef struct mss_stats {
    int64_t kernel_launches;   /* kernels of this library launched since mss_create */
    int64_t solves;            /* windows solved on this rank since mss_create */
    double  last_device_ms;    /* device time of the last mss_solve[_batch] kernel (CUDA events) */
    double  last_total_ms;     /* host wall time of the last call */
    int64_t last_h2d_bytes;    /* bytes staged host->device by the last call */
    int64_t last_d2h_bytes;
    int64_t device_bytes;      /* device scratch currently owned */
    int32_t grid_ctas;         /* CTAs of the persistent kernel in the last call */
    int32_t sm_count;
    int64_t last_row_entries;  /* work accounting of the last call: CSR / live-list entries read by all row phases after the build */
    int64_t last_var_visits;   /* ... and map points visited by all variable phases after round 1 */
} mss_stats;

int         mss_version(void);
int         mss_create(const mss_config* cfg, mss_handle** out);
void        mss_destroy(mss_handle* h);
const char* mss_last_error(const mss_handle* h);            /* never NULL; "" when no error */
int         mss_set_params(mss_handle* h, int32_t min_points, float lambda, float grid_lambda);

/* Solve one window (synchronous: returns after the result is in the caller's buffers). */
int mss_solve(mss_handle* h, const mss_window_view* view, mss_result* result);

/* Solve nwin independent windows in one launch.  With a communicator attached (mss_comm_init) window w is solved by
 * rank w % nranks and the results of all windows are all-gathered (keep bits + per-row coverage only), so every rank
 * returns all nwin results; views of windows owned by other ranks need only K, H, M filled in. */
int mss_solve_batch(mss_handle* h, int32_t nwin, const mss_window_view* views, mss_result* results);

/* Connected components of one window (any layout / memory kind).  The reference's final flush puts every unsparsified
 * keyframe into ONE model (MapSparsification.cc:38-47); that model decomposes exactly along the components of the graph
 * {keyframe rows, window + outside} x {variables}, so each component is an independent window (give it the window-wide
 * nMax through n_max_floor) and a flush can be solved as a batch / sharded over GPUs.  row_label[K+H] and mp_label[M]
 * (same memory kind as the view) receive dense component ids, numbered in order of the first keyframe row of each
 * component; mp_label is -1 for map points that are not variables.  n_max receives the window-wide nMaxObservation. */
int mss_components(mss_handle* h, const mss_window_view* view, int32_t* row_label, int32_t* mp_label, int32_t* ncomp, int32_t* n_max);

/* Multi-GPU: one process per GPU.  Rank 0 calls mss_comm_unique_id, ships the 128 bytes to the other ranks by any
 * means (torch.distributed broadcast, MPI, a file), then every rank calls mss_comm_init. */
#define MSS_UNIQUE_ID_BYTES 128
int mss_comm_unique_id(void* out_id128);
int mss_comm_init(mss_handle* h, const void* id128, int32_t rank, int32_t nranks);
int mss_comm_destroy(mss_handle* h);

/* Multi-GPU from ONE process (the reference is a single process, src/System.cc:159-160: its sparsifier can drive all the
 * GPUs of the box, it cannot be one rank of many).  A multi handle owns one engine per device and one worker thread per
 * device; window w of a batch is solved by device w % n (host views, host result buffers), all devices at the same time.
 * No collective: the host is the only consumer of the bitmasks. */
typedef struct mss_multi mss_multi;
int  mss_multi_create(const mss_config* cfg, const int32_t* devices /* NULL = 0..n-1; cfg->device is ignored */, int32_t n, mss_multi** out);
void mss_multi_destroy(mss_multi* m);
int  mss_multi_device_count(const mss_multi* m);
const char* mss_multi_last_error(const mss_multi* m);
int  mss_multi_set_params(mss_multi* m, int32_t min_points, float lambda, float grid_lambda);
int  mss_multi_solve_batch(mss_multi* m, int32_t nwin, const mss_window_view* views, mss_result* results);
int  mss_multi_get_stats(const mss_multi* m, int32_t device_index, mss_stats* out);

/* Pinned host memory for views/results that travel every call (optional; any host memory works). */
void* mss_host_alloc(size_t bytes);
void  mss_host_free(void* p);
/* Device memory on the engine's device for callers that keep views resident (MSS_MEM_DEVICE). */
void* mss_device_alloc(mss_handle* h, size_t bytes);
void  mss_device_free(mss_handle* h, void* p);
int   mss_memcpy_h2d(mss_handle* h, void* dst_device, const void* src_host, size_t bytes);
int   mss_memcpy_d2h(mss_handle* h, void* dst_host, const void* src_device, size_t bytes);

/* ---- Persistent device mirror of the keyframe x map-point incidence -------------------------------------------
 * Instead of re-walking the pointer graph for every window (MapSparsification.cc:67-151: one copy of mvpMapPoints and of
 * mGrid per keyframe, one copy of mObservations per variable), the incidence lives in HBM and follows the map through small
 * deltas recorded where the map changes:
 *   KeyFrame::mvpMapPoints[idx]   AddMapPoint / EraseMapPointMatch / ReplaceMapPointMatch (src/KeyFrame.cc:299-309)   -> MSS_MOP_SLOT
 *   MapPoint::mObservations       AddObservation / EraseObservation / SetBadFlag / Replace (src/MapPoint.cc:133-255)    -> MSS_MOP_OBS
 *   MapPoint::nObs, mbBad         the same calls                                                                        -> MSS_MOP_MP
 *   KeyFrame::EraseBadDescriptor  compaction of a sparsified keyframe (src/KeyFrame.cc:311-361)                         -> MSS_MOP_KF_COMPACT
 *   a new keyframe (mvpMapPoints, mGrid; src/LocalMapping.cc ProcessNewKeyFrame)                                        -> mss_mirror_add_keyframe
 * A window is then K keyframe handles (4 K bytes over PCIe instead of the whole flattened view): the view -- map points in
 * discovery order = mnIndexForSparsification (MapSparsification.cc:91-99), outside keyframes ordered by their sort key
 * (KeyFrame::mnId) -- is assembled on the device and solved in place; what comes back is a bitmask over map-point HANDLES.
 * Handles are small dense integers chosen by the caller (keyframes: 0, 1, 2, ... ; map points likewise); the mirror grows on
 * demand.  Layout: keyframe-major, `slots_per_kf` slots per keyframe:
 *   slot_mp[kf][i]   handle of mvpMapPoints[i] or -1          slot_cell[kf][i]  col*48+row of the keypoint, MSS_CELL_NONE = not in mGrid
 *   obs_mp[kf][i]    handle of the map point whose mObservations holds (kf -> i) (left index; right index for a right-only
 *                    observation), or -1.  mObservations is mirrored on its own: it is what pass 3 reads (:125-142) and it is
 *                    not always equal to mvpMapPoints (a keyframe LocalMapping has not processed yet holds tracked points
 *                    that do not observe it back). */
typedef struct mss_mirror mss_mirror;

typedef enum mss_mirror_op_kind {
    MSS_MOP_SLOT = 1,        /* a = keyframe, b = slot index, c = map-point handle or -1 */
    MSS_MOP_OBS = 2,         /* a = keyframe, b = slot index, c = map-point handle or -1 */
    MSS_MOP_MP = 3,          /* a = map point, b = Observations(), c = isBad() ? 1 : 0 */
    MSS_MOP_KF_COMPACT = 4   /* a = keyframe: drop its empty slots, keep order; every kept slot's point observes it at the new index */
} mss_mirror_op_kind;

typedef struct mss_mirror_op { int32_t kind, a, b, c; } mss_mirror_op;

typedef struct mss_mirror_stats {
    int32_t n_keyframes;       /* highest keyframe handle seen + 1 */
    int32_t n_map_points;      /* highest map-point handle seen + 1 */
    int32_t slots_per_kf;
    int32_t reserved_;
    int64_t device_bytes;
    int64_t ops_applied;
    int64_t windows_built;
    double  last_build_ms;     /* device time of the view assembly of the last mss_mirror_solve (CUDA events) */
    double  last_solve_ms;     /* device time of the solve kernel of the last mss_mirror_solve */
    double  last_total_ms;     /* host wall time of the last mss_mirror_solve */
    int64_t last_h2d_bytes, last_d2h_bytes;
} mss_mirror_stats;

/* One window of a mirror solve.  Windows of one call must be independent (no shared map point, no shared outside
 * keyframe that sees variables of both): they are, being disjoint covisibility components. */
typedef struct mss_mirror_window {
    int32_t        K;            /* window keyframes */
    int32_t        n_max_floor;  /* as mss_window_view::n_max_floor */
    const int32_t* kf;           /* [K] keyframe handles in window order (host memory) */
    uint32_t*      del_bits;     /* out, host memory, indexed by map-point handle: [(del_words)] bit h = 1 <=> map point h is a
                                    variable the selection dropped (SetBadFlag() it); only the words [h_lo/32, h_hi/32] are written */
    int32_t        del_words;    /* capacity of del_bits in 32-bit words (>= (mss_mirror_stats.n_map_points + 31) / 32) */
    int32_t        h_lo, h_hi;   /* out: handle range [h_lo, h_hi) of the window's map points (h_lo = h_hi = 0 when it has none) */
    int32_t        M, H, F, O;   /* out: sizes of the view assembled on the device */
    int32_t        n_deleted;    /* out */
    int32_t        apply;        /* in: 1 = also apply the deletion to the mirror itself (what SetBadFlag does to the map,
                                    src/MapPoint.cc:227-255: the points become bad, their slots and observations are cleared),
                                    so that the hand-back does not have to travel back as deltas */
    int32_t*       mp_handle;    /* optional out, host memory, [mp_cap]: map-point handle of every table index (bit order of
                                    mss_result::keep_bits) */
    int32_t        mp_cap;
    int32_t        reserved_;
} mss_mirror_window;

int  mss_mirror_create(mss_handle* h, int32_t slots_per_kf, mss_mirror** out);
void mss_mirror_destroy(mss_mirror* m);
/* new keyframe `kf` (or a full refresh of an existing one): n_slots <= slots_per_kf; cells / slot_mp / obs_mp are host
 * arrays of n_slots (obs_mp may be NULL: no observation yet) */
int  mss_mirror_add_keyframe(mss_mirror* m, int32_t kf, uint32_t sort_key, int32_t n_slots, const uint16_t* cells,
                             const int32_t* slot_mp, const int32_t* obs_mp);
/* many keyframes at once (bulk load of an existing map): handles kf0 .. kf0 + n - 1, arrays of n * slots_per_kf (unused
 * tail slots: -1 / MSS_CELL_NONE), n_slots[n], sort_key[n] */
int  mss_mirror_add_keyframes(mss_mirror* m, int32_t kf0, int32_t n, const uint32_t* sort_key, const int32_t* n_slots,
                              const uint16_t* cells, const int32_t* slot_mp, const int32_t* obs_mp);
/* bulk load of map-point attributes: handles mp0 .. mp0 + n - 1 */
int  mss_mirror_set_map_points(mss_mirror* m, int32_t mp0, int32_t n, const int32_t* nobs, const uint8_t* bad);
/* apply ops in array order (a later op on the same address wins) */
int  mss_mirror_apply(mss_mirror* m, const mss_mirror_op* ops, int32_t n);
/* assemble and solve nwin windows; results[w] as in mss_solve_batch (keep_bits / kf_cov / kf_slack host buffers or NULL,
 * keep_bits in table order = discovery order).  Synchronous. */
int  mss_mirror_solve(mss_mirror* m, int32_t nwin, mss_mirror_window* windows, mss_result* results);
/* debug / tests: copy the view the mirror assembles for one window into caller-sized host arrays (MSS_LAYOUT_PACKED, slots
 * sorted by value inside each keyframe).  Pass NULL arrays to get the sizes (K, H, M, F, O) only. */
int  mss_mirror_build_view(mss_mirror* m, const mss_mirror_window* window, int32_t* sizes5, int32_t* feat_ptr, uint32_t* slots,
                           uint16_t* mp_nobs16, uint32_t* obs_pairs, int32_t* okf_total, int32_t* mp_handle, int32_t* okf_handle);
/* connected components of one window (see mss_components): kf_label[K] (host) receives the component of every window
 * keyframe, numbered in order of first appearance; n_max the window-wide nMaxObservation.  The final flush of the sparsifier
 * solves the components as one batch of independent windows (n_max_floor = n_max). */
int  mss_mirror_components(mss_mirror* m, const mss_mirror_window* window, int32_t* kf_label, int32_t* ncomp, int32_t* n_max);
int  mss_mirror_get_stats(const mss_mirror* m, mss_mirror_stats* out);

/* ---- Compaction of sparsified keyframes on the device ----------------------------------------------------------
 * KeyFrame::EraseBadDescriptor (src/KeyFrame.cc:311-361) keeps, in order, the rows of mDescriptors (32 bytes), mvKeysUn
 * (cv::KeyPoint, 28 bytes), mvuRight and mvDepth whose slot still holds a map point.  mss_compact_keyframes does that in place
 * on device-resident arrays, one CTA per keyframe; any of the four arrays may be NULL.  n_out[nkf] (host) receives the rows
 * left.  With mss_mirror_compact_keyframes the flags come from the mirror (slot_mp[kf][i] >= 0, i.e. after a window's
 * deletion was applied on the device) and the mirror's own rows of these keyframes are compacted as well (MSS_MOP_KF_COMPACT). */
typedef struct mss_kf_payload {
    int32_t n;                 /* rows (keypoints) before the compaction */
    int32_t kf;                /* keyframe handle in the mirror (mss_mirror_compact_keyframes), else -1 */
    const uint8_t* keep;       /* [n] device: 1 = the row survives (mss_compact_keyframes) */
    void*   descriptors;       /* [n][32] device */
    void*   keypoints;         /* [n][28] device */
    float*  uright;            /* [n] device */
    float*  depth;             /* [n] device */
} mss_kf_payload;
int mss_compact_keyframes(mss_handle* h, int32_t nkf, const mss_kf_payload* kfs, int32_t* n_out);
int mss_mirror_compact_keyframes(mss_mirror* m, int32_t nkf, const mss_kf_payload* kfs, int32_t* n_out);

/* ---- BoW re-transform of compacted keyframes + KeyFrameDatabase inverted file ---------------------------------------
 * After the compaction the reference transforms the surviving descriptors of a sparsified keyframe again
 * (mpORBvocabulary->transform(vCurrentDesc, mBowVec, mFeatVec, 4), src/KeyFrame.cc:352-354; DBoW2 TemplatedVocabulary.h:1126-1259)
 * and LoopClosing::DeleteOutdatedInfo adds it to the KeyFrameDatabase (src/LoopClosing.cc:318-329).
 *
 * mss_voc_create: the vocabulary tree as DBoW2's text file lists it (TemplatedVocabulary.h:1338-1418): node 0 is the root,
 * node i > 0 has parent[i] < i, is_leaf[i], a 32-byte ORB descriptor and a weight; children keep their order of appearance,
 * word ids are handed out to the leaves in file order.  TF-IDF weighting with L1 scoring (ORBvoc: "10 6 0 0").
 * mss_bow_transform: per keyframe n descriptors in DEVICE memory (e.g. the rows mss_compact_keyframes left) -> per feature the
 * word and the node `levelsup` levels above the leaves (HOST arrays, may be NULL), the BowVector (distinct words ascending +
 * L1-normalised values, bit-identical to DBoW2's doubles) and the FeatureVector as (node, feature) pairs in map order.
 * add_to_database != 0: the keyframes are appended to the inverted file (KeyFrameDatabase::add).
 * mss_kfdb_common_words: words a query BowVector shares with every database keyframe id < kf_cap (the counting loop of the
 * detection queries, src/KeyFrameDatabase.cc:610-640). */
typedef struct mss_vocabulary mss_vocabulary;
typedef struct mss_bow_keyframe {
    int32_t n;                 /* descriptors (<= 4096) */
    int32_t kf_id;             /* id stored in the inverted file */
    const void* descriptors;   /* [n][32] device */
    int32_t* word;             /* [n] host out, may be NULL */
    int32_t* node;             /* [n] host out, may be NULL */
    int32_t* bow_word;         /* [n] host out: distinct words, ascending */
    double*  bow_value;        /* [n] host out */
    int32_t* n_bow;            /* host out: entries of the BowVector */
    int32_t* fv_node;          /* [n] host out: FeatureVector, pairs sorted by (node, feature) */
    int32_t* fv_feature;       /* [n] host out */
    int32_t* n_fv;             /* host out */
} mss_bow_keyframe;
int  mss_voc_create(mss_handle* h, int32_t n_nodes, int32_t levels, const int32_t* parent, const uint8_t* is_leaf, const uint8_t* descriptors,
                    const double* weight, mss_vocabulary** out);
void mss_voc_destroy(mss_vocabulary* v);
int32_t mss_voc_words(const mss_vocabulary* v);
int  mss_bow_transform(mss_vocabulary* v, int32_t nkf, const mss_bow_keyframe* kfs, int32_t levelsup, int32_t add_to_database);
int  mss_kfdb_common_words(mss_vocabulary* v, int32_t n_query_words, const int32_t* query_words, int32_t kf_cap, int32_t* common);
int32_t mss_kfdb_postings(const mss_vocabulary* v);

/* Lower bound of the reference ILP, proven on the device per window (mss_result.dual_bound; NaN while off -- the default --
 * or when the round cap was hit).  The reference gets its certificate from GUROBI's branch and bound (MIPGap 0.002,
 * src/MapSparsification.cc:153-157); here it comes from the state reached by exact dominance rules + one round of dual
 * ascent (csrc/mss_bound.cuh): objective / dual_bound - 1 is a certified optimality gap.  Costs six short extra phases per
 * window and a copy of the residual lists in device memory.  Also switched on by the environment variable MSS_DUAL_BOUND=1. */
int   mss_set_dual_bound(mss_handle* h, int32_t enable);
int   mss_get_stats(const mss_handle* h, mss_stats* out);
/* The CUDA stream (cudaStream_t) the handle launches on, for callers that time with their own events. */
void* mss_stream(mss_handle* h);

/* Debug: per-phase device timeline of the windows solved on this rank by the next calls.  mss_debug_get_trace copies up
 * to cap_pairs (phase << 24 | FREE points left, ns since the window started) pairs of local window i of the last call
 * and returns how many.  Not part of the reference-facing surface. */
int   mss_debug_trace(mss_handle* h, int32_t enable);
int   mss_debug_get_trace(const mss_handle* h, int32_t local_window, uint32_t* out_pairs, int32_t cap_pairs);

#ifdef __cplusplus
}
#endif
#endif /* MSS_H_ */
