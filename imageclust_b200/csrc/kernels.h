// kernels.h -- host-callable launchers of the sm_100a kernels (internal to the library).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

namespace ic {

// ---- resident per-slot state -------------------------------------------------------------
// A "slot" is a row/column of the distance matrix.  Slot s holds one live cluster
// (the reference's clusters[] entry, clustering.go:11-15) or is retired.
//   ks[s]  = {key, size}: key is the monotone order id (item index for singletons,
//            N + t for the cluster made by merge t) == the reference's slice order;
//            key < 0 when the slot is retired.
//   nn[s*kNNK + j] = j-th smallest cached partner among clusters with a LOWER key:
//            {partner key, distance bits, partner slot, partner size};
//            y == 0xFFFFFFFF for an empty entry.  nn_more[s] != 0: selectable partners
//            may exist beyond the list (rescan the row when the list runs dry).
typedef int2 SlotKS;
typedef uint4 SlotNN;
constexpr int kNNK = 4;  // cached partners per row
constexpr uint32_t kNoPartner = 0xFFFFFFFFu;

// ---- K0 prep (prep.cu) -------------------------------------------------------------------
// column sums of X [n x d] (row stride ldx) in double
cudaError_t launch_colsum(const float* x, int64_t n, int64_t d, int64_t ldx, double* colsum, cudaStream_t s);
// xc = x - mean; hi = tf32(xc); lo = tf32(xc - hi); norm[i] = sum_k (hi+lo)^2 (double).
// hi/lo are [n_pad x d_pad] (zero padded), norms [n_pad].
cudaError_t launch_split(const float* x, int64_t n, int64_t d, int64_t ldx, const double* colsum, int center,
                         float* hi, float* lo, double* norms, int64_t n_pad, int64_t d_pad, cudaStream_t s);
cudaError_t launch_fill(float* dm, int64_t count, float value, cudaStream_t s);
// label block of the combined feature rows (embeddings.go:166-183): x[i][d_img + id] = 1 for the item's labels, else 0
cudaError_t launch_label_block(float* x, int64_t n, int64_t d, int64_t d_img, const int32_t* label_offsets,
                               const int32_t* label_ids, cudaStream_t s);
// ks[s] = {s, 1}, gkey[s] = s (padding up to a multiple of 4: -1)
cudaError_t launch_init_slots(SlotKS* ks, int32_t* gkey, int64_t n, cudaStream_t s);

// ---- K1 Gram / initial Ward distances ----------------------------------------------------
constexpr int kGramBM = 128;  // tile rows   (UMMA M)
constexpr int kGramBN = 256;  // tile cols   (UMMA N)
constexpr int kGramBK = 32;   // fp32 elements per 128-byte swizzle row
struct GramPlan {
    CUtensorMap map_hi;  // box {32 floats, 128 rows} over hi [n_pad x d_pad]
    CUtensorMap map_lo;
    const int2* tiles;   // (row block, col block) list in L2-friendly order
    int n_tiles;
    int k_blocks;        // d_pad / 32
};
// dm[i][j] = dm[j][i] = max(0.5*(norm_i + norm_j) - <x_i, x_j>, 0), diag 0; full square, row stride ld.
// dm holds the rows [row_begin, row_end) only (a rank's row block; 0..n on one GPU), symmetric and at full
// width: the plan's tile list covers every tile whose rows or columns touch them.
cudaError_t launch_gram_tcgen05(const GramPlan& plan, const double* norms, float* dm, int64_t n, int64_t ld,
                                int64_t row_begin, int64_t row_end, int num_sms, cudaStream_t s, int terms = 23);
size_t gram_tcgen05_smem_bytes();

// ---- K1 (int8 Ozaki slices): the same Gram identity with EXACT integer tensor-core arithmetic -----------------
// K0': x_c = x - mean is rounded to a 21-bit fixed-point value per row, v = rint(x_c / q_i), q_i = 2^(e_i - 20), and cut
// into three balanced base-128 digits v = h*2^14 + m*2^7 + l (each in [-64, 64]); norm_i = q_i^2 * sum v^2 (double).
constexpr int kI8Tile = 128;   // tile rows == tile cols (UMMA M = N = 128)
constexpr int kI8BK = 128;     // int8 elements per 128-byte swizzle row
cudaError_t launch_split_i8(const float* x, int64_t n, int64_t d, int64_t ldx, const double* colsum, int center,
                            int8_t* h, int8_t* m, int8_t* l, float* quanta, double* norms, int64_t n_pad, int64_t d_pad,
                            cudaStream_t s);
struct GramI8Plan {
    CUtensorMap map_h, map_m, map_l;  // box {128 bytes, 128 rows} over the slice arrays [n_pad x d_pad] (int8)
    const int2* tiles;                // (row block, col block) of 128 x 128 tiles, L2-friendly order
    int n_tiles;
    int k_blocks;                     // d_pad / 128
};
// six kind::i8 products per k-step into three int32 TMEM accumulators (weights 2^28, 2^21, 2^14); the three dropped
// low-order products are below 2^-21 of the Gram value.  Same output contract as launch_gram_tcgen05.
cudaError_t launch_gram_i8(const GramI8Plan& plan, const double* norms, const float* quanta, float* dm, int64_t n,
                           int64_t ld, int64_t row_begin, int64_t row_end, int num_sms, cudaStream_t s, int debug = 0);
// audit kernel: the reference's own arithmetic (sequential fp32, clustering.go:136-157), bit exact
cudaError_t launch_gram_exact(const float* x, int64_t n, int64_t d, int64_t ldx, float* dm, int64_t ld,
                              int64_t row_begin, int64_t row_end, cudaStream_t s);

// ---- K2 nearest-neighbour sweep ----------------------------------------------------------
// First sweep (keys are the slot indices): nn[s][*] = the kNNK smallest (dm[s][u], u) over columns u < s,
// for the resident rows s in [row_begin, row_end).
cudaError_t launch_nn_sweep(const float* dm, int64_t row_begin, int64_t row_end, int64_t ld, SlotNN* nn,
                            int32_t* nn_more, cudaStream_t s);

// ---- K3 persistent merge loop ------------------------------------------------------------
// The distance matrix is row-block sharded over P "ranks" (SURVEY 8e): rank q owns slots
// [q*C, (q+1)*C), C = ceil(n / P) rounded up to a multiple of 4, i.e. the rows of those slots with ALL their columns.
//   * one GPU:            P = 1.
//   * several GPUs:       one process / GPU / rank; every rank launches the kernel with n_local = 1
//                         and reaches its peers' rows and rank mailboxes through peer-mapped pointers.
//   * virtual shards:     P ranks emulated by ONE cooperative launch on one GPU (n_local = P): same code
//                         path, used to test the sharded protocol on a single device.
constexpr int kMaxRanks = 8;
constexpr int kReqPerBlock = 2;  // row rescans a block may request per iteration
struct LoopState {
    int32_t n;              // slots (== items)
    int32_t n_ranks;        // P
    int32_t rank0;          // first rank this launch runs (multi-GPU: my rank; otherwise 0)
    int32_t n_local;        // ranks this launch runs (1, or P for virtual shards)
    int32_t rows_per_rank;  // C
    uint32_t gen;           // launch generation: part of every mailbox tag (mailboxes are never re-zeroed)
    int64_t ld;
    float* dm_rank[kMaxRanks];   // first row of rank q's row block [C x ld]
    void* rankbox[kMaxRanks];    // rank q's inter-rank mailbox: [2][P] 128-byte records
    // per local rank v (v = rank - rank0): base + v * stride
    SlotKS* ks;        // [n_local][n]   replica of every slot's {key, size}, maintained by block 0 of the rank
    int32_t* gkey;     // [n_local][n4]  keys only (what a row scan streams); padding = -1
    SlotNN* nn;        // [n][kNNK]      slot indexed: only the owner of a slot touches its entries
    int32_t* nn_more;  // [n]            bit 0: partners beyond the list; bit 1: list ran dry (entry 0 = lower bound)
    // merge trace, [n_local][n] each (every rank records the same trace)
    int32_t* tr_key_hi;
    int32_t* tr_key_lo;
    float* tr_dist;
    int32_t* tr_size;
    float* tr_gap;
    void* records;     // [n_local] x { [reader G][2][writer G] 128-byte records }
    int64_t records_stride;
    void* partials;    // [n_local] x { [owner G][2][kReqPerBlock][sender G] 128-byte records }
    int64_t partials_stride;
    int32_t* ctl;      // [n_local][16], see CTL_*
    long long* prof;   // [16] or NULL: SM cycles block 0 of local rank 0 spent per phase
};
struct LoopParams {
    int32_t n_target;    // CalculateOptimalClusters result (clustering.go:220)
    int32_t max_size;    // clustering.go:228
    int32_t max_merges;  // stop after this many merges in this launch (<0: unlimited)
    float near_tie_tol;
    int32_t scan_every;  // rescan requests are published every scan_every-th iteration (batched row scans)
    int32_t debug;       // experiments only (currently unused by the kernels)
    // Reference arithmetic (batched loop only).  exact != 0: every stored pair <= horizon holds the reference's own value
    // (WardDistance of the two fp32 centroids, clustering.go:136-145); pairs above it are Lance-Williams values, known to
    // within eps_filter.  Merges are only taken among values <= safe = horizon / (1 + 2 eps_filter); when the minimum
    // gets there the kernel stops (STOP_HORIZON) and the host raises the horizon.  Members of a batch after the first
    // must be below T (1 - delta_cut): fp32 centroid distances are reducible only up to rounding.
    int32_t refill_at;   // a row whose partner list holds fewer valid entries than this (and may have more partners) is rescanned
    int32_t exact;
    float eps_filter;
    float abs_slack;     // absolute error a never re-evaluated tensor-core Gram value may carry (monitor only; safe includes it)
    double horizon, safe, delta_cut;
};
// ctl[] indices.  N_LIVE and N_MERGES are read at launch (resume) and written at exit.
constexpr int kCtlWords = 32;
enum { CTL_N_LIVE = 0, CTL_N_MERGES = 1, CTL_EXHAUSTED = 2, CTL_ERROR = 3, CTL_NEAR_TIES = 4, CTL_RESCANS = 5,
       CTL_DONE = 6, CTL_BUBBLES = 7, CTL_NEXT_HI = 8, CTL_NEXT_LO = 9, CTL_NEXT_DIST = 10, CTL_STOP = 11,
       // (12 = CTL_ITERS, below) reference-arithmetic bookkeeping of the batched loop and of refine.cu:
       CTL_XQ_OVERFLOW = 13,   // the exact-evaluation queue of an iteration overflowed (the host re-evaluates the new rows)
       CTL_FILTER_VIOL = 14,   // re-evaluated pairs whose stored value was off by more than eps_filter (must stay 0)
       CTL_FILTER_MAXERR = 15, // largest |stored - reference| / reference seen (float bits)
       CTL_N_EXACT = 16,       // pairs evaluated with the reference's arithmetic
       CTL_ORDER_VIOL = 17,    // pairs created inside a batch that came out below a later member of it (must stay 0)
       CTL_N_CUT = 18,         // batches shortened by delta_cut
       CTL_XQ_FIRST_KEY = 19   // key of the first cluster created by the iteration whose queue overflowed
};
// CTL_STOP values
enum { STOP_TARGET = 1, STOP_EXHAUSTED = 2, STOP_MAX_MERGES = 3, STOP_EPOCHS = 4, STOP_ERROR = 5, STOP_HORIZON = 6,
       STOP_XQ = 7, STOP_ORDER = 8, STOP_COMPACT = 9 };
constexpr int kLoopThreads = 512;
// C: rows (slots) per rank, ceil(n / P) rounded up to a multiple of 4
int64_t merge_loop_rows_per_rank(int64_t n, int n_ranks);
// blocks per rank for a problem of n slots over n_ranks ranks, n_local of them in this launch
// `replica`: every block keeps the keys of ALL slots in shared memory (needs merge_loop_replica_fits(n))
bool merge_loop_replica_fits(int64_t n);
cudaError_t merge_loop_grid(int num_sms, int64_t n, int n_ranks, int n_local, int want_blocks, bool replica,
                            int* blocks_per_rank);
size_t merge_loop_smem_bytes(int64_t n, int n_ranks, int blocks_per_rank, bool replica);
size_t merge_loop_records_bytes(int blocks_per_rank);   // per rank
size_t merge_loop_partials_bytes(int blocks_per_rank);  // per rank
size_t merge_loop_rankbox_bytes();                      // per rank (barrier flags + [2][kMaxRanks] records)
cudaError_t launch_merge_loop(const LoopState& st, const LoopParams& p, int blocks_per_rank, bool replica,
                              cudaStream_t s);
// ---- K3b batched merge loop (merge_batch.cu): one GPU, several provably consecutive merges per iteration -----
constexpr int kBatchThreads = 512;
constexpr int kMaxBatch = 512;        // candidate pairs / merges per iteration (<= kBatchThreads: one per thread)
constexpr int kBatchCand = 64;        // candidate pairs a block may publish per iteration
constexpr int kBatchMaxDry = 2048;    // rows rescanned per round of the rescan phase (partial-list buffers)
#ifndef IC_WIN_MIN
#define IC_WIN_MIN 2048
#endif
constexpr int kBatchWinMin = IC_WIN_MIN;    // columns of a row-scan window (at most kBatchMaxWin windows per row)
constexpr int kBatchMaxWin = 128;
constexpr int kBatchMaxBlocks = 160;
constexpr int CTL_ITERS = 12;         // ctl[]: iterations of the batched loop
// Sharded runs (one process per GPU, row-block shards like merge_loop.cu): every rank runs the kernel on its own row
// block; the ranks exchange their candidate pairs once per iteration through peer-mapped "exchange boxes":
// Everything a rank needs for the batch selection is PUSHED into its box by the peers (stores over NVLink are fire and
// forget; a remote load costs ~3 000 cycles):
//   [0, 256)        u64 flags[kMaxRanks]            cross-rank barrier: flags[q] is written by rank q
//   [256, 1280)     summary[src][slot] of 32 bytes   {u64 stopper minimum, u64 head minimum, i32 candidate pairs} of rank
//                                                    src in iteration slot (iteration mod 3), written by src's block 0
//   [1280, 1344)    u64 stop[3], head[3]; i32 cnt[3] this rank's own accumulators (local atomics)
//   [2048, ...)     uint4 cand[src][kBatchXCand][2]  candidate pairs of rank src in the current iteration
constexpr int kXResCap = 512;
constexpr int kBatchXCand = 2048;
constexpr size_t kBatchXSummary = 256, kBatchXAccum = 1280, kBatchXCandBase = 2048;
//   [.., +768)      uint4 rankmin[src][slot][2]      behind the candidates: the smallest head of a rank whose candidates did
//                                                    not fit its region (then only the global minimum is merged)
constexpr size_t kBatchXRankMin = kBatchXCandBase + static_cast<size_t>(kMaxRanks) * kBatchXCand * 32;
constexpr size_t kBatchXBoxBytes = kBatchXRankMin + 1024;
struct BatchState {
    int32_t n;                             // slots (== items until the first compaction, then the live count at the last one)
    int32_t key_base;                      // N: the cluster created by merge t carries the key N + t
    // K4 epochs (compact.cu).  Clusters with key < order_key sit in key order by slot (a row scan of such a cluster only
    // needs the columns before its own); pairs of two clusters with key < mirror_key are stored in both rows.
    int32_t order_key, mirror_key;
    int32_t compact_at;                    // stop (STOP_COMPACT) when the live count is <= this; 0: never
    int32_t n_ranks, rank, rows_per_rank;  // 1, 0, - on one GPU
    uint32_t gen;                          // launch generation (cross-rank barrier sequence numbers)
    float* dm_rank[kMaxRanks];             // sharded: first row of every rank's row block (peer mapped)
    uint8_t* xbox[kMaxRanks];              // sharded: every rank's exchange box (peer mapped)
    int32_t win_cols, n_win;  // filled in by launch_merge_batch
    int64_t ld;
    float* dm;          // [rows x ld] (rows = n, or the rank's row block); lower triangle valid at launch, then a pair lives in
                        // the row of its higher-key cluster
    SlotKS* ks;         // [n4] {key, size}; padding key = -1
    int32_t* gkey;      // [n4]
    SlotNN* nn;         // [n][kNNK]
    int32_t* nn_more;   // [n]
    int32_t* tr_key_hi;
    int32_t* tr_key_lo;
    float* tr_dist;
    int32_t* tr_size;
    float* tr_gap;
    int32_t* ctl;       // [16]
    long long* prof;    // [16] or NULL
    // scratch (zeroed before every launch)
    uint4* hdr;         // [kBatchMaxBlocks] per block: {stopper minimum lo, hi, head minimum lo, hi}
    uint4* blockmin;    // [kBatchMaxBlocks][2] sharded: every block's smallest head as a candidate record
    uint4* cand;        // [n][2] candidate pairs {head lo, head hi, row slot, partner slot} {size, size, partner key, 0}
    int32_t* counters;  // [3][4] dry rows, candidate pairs; per iteration mod 3
    int2* dryq;         // [n]   rows to rescan {slot, key}
    int32_t* lsize;     // [n4]  size of the live cluster in every slot, 0: retired (rebuilt from ks at launch)
    uint4* partials;    // [kBatchMaxDry][kBatchMaxWin][8] partial lists of the window scans
    int32_t* part_cnt;  // [kBatchMaxDry] windows done per row
    uint32_t* bar;      // grid barrier counter
    // reference arithmetic (LoopParams::exact): fp32 centroid of every slot's cluster (replicated on every rank), and
    // the queue of the pairs an iteration wrote at or below the horizon {index of the merge in the batch, column slot or
    // 0x80000000 | index of an earlier merge of the batch (cross term)}
    float* cen;         // [2N x ldc] by cluster key (row N + t: the cluster made by merge t), ldc = d rounded up to 4, zero padded
    int64_t ldc;
    int2* xqm;          // [kMaxBatch][kXResCap] per merge of the batch: {column | cross, Lance-Williams value bits}
    int4* xq;           // [xq_cap] overflow of the merges' queues: {merge, column | cross, value bits, position}
    int32_t xq_cap;
    // one GPU: the partner list of a new cluster is selected from its re-evaluated pairs (everything else in its row is
    // above the horizon), so its row needs no scan: xres[merge][position] = {value bits, partner key, partner slot, size}
    uint4* xres;        // [kMaxBatch][kXResCap]
    int32_t* xhit;      // [kMaxBatch] pairs queued per merge of the current batch
    int32_t* xfar;      // [kMaxBatch] the merge's new row has finite values above the horizon
    // near lists (near.cu); near_meta == nullptr: not in use (Lance-Williams only runs)
    int2* near_meta;    // [n] by slot
    uint2* near_pool;
    int32_t near_pool_cap;
    int32_t* near_cursor;  // [0]: next free pool entry
    int32_t* slot_of_key;  // [2N] slot of every live cluster's key, -1: not alive
    int2* nearq;        // [n] rows whose partner list is re-selected from their near list {slot, key}
};
size_t merge_batch_smem_bytes(int64_t n);
int64_t merge_batch_windows(int64_t n);
cudaError_t merge_batch_grid(int num_sms, int64_t n, int* blocks);  // *blocks = 0: does not fit
cudaError_t launch_merge_batch(const BatchState& st, const LoopParams& p, int blocks, cudaStream_t s);  // n_ranks > 1: sharded
// test hook: P virtual ranks of the sharded kernel on one GPU (one cooperative launch, P * blocks_per_rank blocks)
void merge_batch_fill_windows(BatchState* st);
cudaError_t launch_merge_batch_virtual(const BatchState* d_states, int n_ranks, int64_t n, const LoopParams& p, int blocks_per_rank,
                                       cudaStream_t s);
// ---- reference arithmetic for selected pairs (refine.cu) ---------------------------------------------------------
// cen[k] = x[k] for k < N (zero padded to ldc floats per row): the singleton centroids (clustering.go:19-20); centroids are
// stored by cluster key
cudaError_t launch_init_centroids(const float* x, int64_t n, int64_t d, int64_t ldx, float* cen, int64_t ldc, cudaStream_t s);
struct RefineArgs {
    float* dm;            // resident rows [r_lo, r_hi) x ld
    int64_t ld;
    int32_t n_slots, r_lo, r_hi;
    const SlotKS* ks;     // {key, size} per slot
    const int32_t* gkey;  // [n4] keys, padding -1
    const float* cen;
    int64_t ldc;
    int32_t mirror_key;   // pairs of two clusters with key < mirror_key are stored in both rows (compact.cu)
    int32_t rows_per_rank;
    float* dm_rank[kMaxRanks];  // every rank's row block (mirrored stores); [0] = dm on one GPU
    double lo, hi;        // band: lo < stored value <= hi
    int32_t min_row_key;  // only rows whose cluster key is >= this
    int32_t lower_only;   // keys are still the slot indices: row r only has partners in the columns u < r
    int32_t row0, row1;   // rows collected by this launch (within [r_lo, r_hi))
    int2* q;              // [cap] collected pairs {row slot, column slot}
    int32_t cap;
    int32_t* cnt;         // [2]: pairs collected (may exceed cap: overflow), unused
    int32_t* ctl;         // monitors (CTL_FILTER_*, CTL_N_EXACT)
    int32_t* nn_more;     // rows whose values changed get their partner list queued for a rescan (dry bit)
    float eps_filter, abs_slack;
};
// pairs (r, u), key_u < key_r, both live, whose stored value lies in the band -> q
cudaError_t launch_refine_collect(const RefineArgs& a, cudaStream_t s);
// dm[r][u] = WardDistance(centroid r, centroid u) for the first min(*cnt, cap) collected pairs, one warp per pair
cudaError_t launch_refine_eval(const RefineArgs& a, int num_sms, cudaStream_t s);

// ---- K4 active-cluster compaction (compact.cu) -------------------------------------------------------------------
struct CompactArgs {
    int32_t n_old, n_new, n_new4;           // slots before; live clusters == slots after; rounded up to 4
    const int32_t* oldslot;                 // [n_new4] old slot of every new slot (key order), -1 padding
    const int32_t* newslot;                 // [n_old]  new slot of every old slot, -1: retired
    const SlotKS* ks_old;
    const SlotNN* nn_old;
    const int32_t* nn_more_old;
    SlotKS* ks_new;
    int32_t* gkey_new;
    SlotNN* nn_new;
    int32_t* nn_more_new;
    // matrix: old rows may live on peers (sharded); the new rows [row0, row1) are resident here, first at dm_new
    const float* dm_old[kMaxRanks];
    int32_t rows_per_rank_old;
    int64_t ld_old;
    float* dm_new;
    float* dm_new_rank[kMaxRanks];          // every rank's new row block (the mirror pass reads the transposed tiles)
    int32_t rows_per_rank_new, row_base_new, row0, row1;
    int64_t ld_new;
    int32_t my_rank;                        // real shards: partner lists live with their row's owner; -1 otherwise
    int32_t order_key_old;                  // clusters with a key below this sat in key order by slot in the old layout
    const int2* near_meta_old;              // near lists follow their rows (nullptr: not in use)
    int2* near_meta_new;
    int32_t* slot_of_key;                   // rewritten for the live clusters' new slots
    uint32_t* gmin;                         // mirror pass: atomicMin of the selectable lower-triangle values' bits (or nullptr)
};
// newslot / oldslot from the live keys (ascending); *n_live_out = live clusters found
cudaError_t launch_compact_map(const SlotKS* ks, int32_t n_old, int32_t* keymap, int32_t key_cap, int32_t* newslot,
                               int32_t* oldslot, int32_t n_new4, int32_t* n_live_out, cudaStream_t s);
cudaError_t launch_compact_state(const CompactArgs& a, cudaStream_t s);  // slot tables, partner lists, centroid rows
cudaError_t launch_compact_rows(const CompactArgs& a, cudaStream_t s);   // lower triangle of the new matrix (+ diagonal, padding)
cudaError_t launch_mirror_lower(const CompactArgs& a, cudaStream_t s);   // upper triangle <- lower triangle
// one process (contiguous matrices): both triangles of the new matrix in one pass, 64 x 64 tiles
cudaError_t launch_compact_tiles(const CompactArgs& a, cudaStream_t s);

// ---- near lists (near.cu): every lower-key partner of a row whose stored value is <= horizon ----------------------
// meta[slot] = {offset into the pool, count | kNearFarBit if partners beyond the horizon may exist}; .y == -1: the row has no
// near list (the loop scans the whole row instead)
constexpr int32_t kNearFarBit = 0x40000000, kNearCntMask = 0x3FFFFFFF;
struct NearArgs {
    const float* dm;      // resident rows [r_lo, r_hi) x ld
    int64_t ld;
    int32_t n_slots, r_lo, r_hi;
    const int32_t* gkey;
    int32_t order_key;    // rows of clusters with a key below this only have partners in the columns before their own
    double horizon;
    int2* meta;           // [n_slots]
    uint2* pool;          // {value bits, partner key}
    int32_t pool_cap;
    int32_t* cursor;      // [0]: entries in use after the build, [1]: 1 if the pool held them all
};
cudaError_t launch_slot_of_key_init(int32_t* sok, int32_t n, int32_t cap, cudaStream_t s);  // sok[k] = k for items, -1 above
cudaError_t launch_near_build(const NearArgs& a, cudaStream_t s);  // count sweep, scan, fill sweep
cudaError_t launch_mark_rows_dry(const int32_t* gkey, int32_t* nn_more, int32_t r_lo, int32_t r_hi, cudaStream_t s);
cudaError_t launch_init_lists_dry(SlotNN* nn, int32_t* nn_more, int32_t n, cudaStream_t s);  // no lists: every row rebuilt before use

// device-side barrier of the P single-GPU processes (one lane per peer, flags in the rank mailboxes)
cudaError_t launch_rank_barrier(void* const* rankbox, int n_ranks, int rank, uint64_t seq, cudaStream_t s);

}  // namespace ic
