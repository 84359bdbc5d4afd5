/*
 * imageclust_b200.h -- C ABI of the B200-native size-constrained Ward clustering.
 *
 * Drop-in boundary for ONE reference path:
 *   internal/clustering.PerformClusteringWithConstraints   clustering.go:198-284
 *   internal/clustering.CalculateOptimalClusters           clustering.go:168-186
 * (the reference has no FFI/plugin interface: the Go function is the boundary and
 * its only caller is internal/workflow/workflow.go:89-94).  A cgo shim keeps the
 * Go signatures and calls the entry points below; see INTEGRATION.md.
 *
 * Everything is extern "C", plain pointers and sizes.  All functions return
 * IC_OK (0) or a negative IC_ERR_* code; ic_last_error(ctx) gives the text.
 * There is NO CPU fallback: without a CUDA device ic_create fails.
 *
 * Threading: a context owns one device, one stream and its workspaces; calls on
 * the same context must be serialised by the caller, distinct contexts may run
 * concurrently (the reference function is re-entrant; the shim holds a mutex).
 */
#ifndef IMAGECLUST_B200_H
#define IMAGECLUST_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden */
#endif

#define IC_OK 0
#define IC_ERR_TOO_FEW (-1)  /* totalItems < minSize            clustering.go:169-171 */
#define IC_ERR_UNSAT (-2)    /* ceil(N/max) > floor(N/min)      clustering.go:175-177 */
#define IC_ERR_BAD_ARG (-3)  /* NULL / non-positive sizes / ragged input (Go would panic, :149-151) */
#define IC_ERR_CUDA (-4)     /* CUDA runtime / driver failure */
#define IC_ERR_OOM (-5)      /* problem does not fit this device's HBM */
#define IC_ERR_STATE (-6)    /* staged call made out of order */
#define IC_ERR_TIMEOUT (-7)  /* device-side watchdog tripped (internal error) */
#define IC_ERR_INTERNAL (-9)

/* ic_initial_distances modes */
#define IC_GRAM_TCGEN05_3XTF32 0 /* K1 (first version): TMA + tcgen05 kind::tf32, exact fixed-point slice + residual, fp32 TMEM */
#define IC_GRAM_EXACT_FP32 1     /* SIMT kernel with the reference's own sequential fp32 arithmetic (audit path) */
#define IC_GRAM_TCGEN05_I8 2     /* K1 (default): TMA + tcgen05 kind::i8, three int8 digits of a 22-bit fixed-point row, exact int32
                                    accumulation in TMEM (6 products per k-step at 4x the TF32 rate) */

typedef struct ic_ctx ic_ctx;

typedef struct ic_stats {
    int64_t n_items;       /* N */
    int64_t dim;           /* D */
    int32_t n_target;      /* CalculateOptimalClusters result */
    int32_t n_merges;      /* merges performed */
    int32_t n_final;       /* clusters alive when the loop ended */
    int32_t n_out;         /* clusters in the output map (size >= minSize) */
    int32_t exhausted;     /* loop ended with no admissible pair (clustering.go:222-225) */
    int32_t n_near_ties;   /* merges whose runner-up was within near_tie_tol (relative) */
    int32_t n_rescans;     /* row rescans done by the merge loop */
    int32_t gram_mode;     /* IC_GRAM_* used */
    float near_tie_tol;
    /* device time of each phase, CUDA events on the context's stream, milliseconds */
    float ms_h2d;          /* pinned host -> device copy of X */
    float ms_prep;         /* K0: centring, TF32 hi/lo split, norms */
    float ms_gram;         /* K1: initial distance matrix */
    float ms_nn_init;      /* K2: first nearest-neighbour sweep */
    float ms_loop;         /* K3: persistent merge loop */
    float ms_d2h;          /* merge trace read-back */
    float ms_host;         /* host assembly of the cluster lists (wall clock) */
    float ms_total;        /* wall clock of the whole call */
    int64_t kernel_launches; /* kernels this call launched */
    int64_t h2d_bytes;
    int64_t d2h_bytes;
    int64_t matrix_bytes;  /* bytes of the distance matrix resident in HBM */
    int32_t n_iterations;  /* iterations of the merge loop (batched loop: several merges each) */
    int32_t loop_mode;     /* 1: batched loop (merge_batch.cu), 0: one merge per iteration (merge_loop.cu) */
    /* reference arithmetic (option "exact"): the loop keeps Lance-Williams values and re-evaluates every pair at or below
     * a horizon as WardDistance of the two fp32 centroids, clustering.go:83-86,136-157 */
    int32_t exact;            /* 1: the run used the horizon (merge sequence = the reference's arithmetic) */
    int32_t n_horizon_raises; /* times the horizon was set / raised (each: one sweep of refine.cu) */
    int64_t n_exact;          /* pairs evaluated with the reference's arithmetic */
    int32_t n_filter_viol;    /* re-evaluated pairs whose stored value was off by more than eps_filter: must be 0, else the
                                 horizon's guarantee does not hold and the sequence may differ from the reference's */
    int32_t n_order_viol;     /* pairs created inside a batch that came out below a later pair of it: must be 0 */
    int32_t n_cut;            /* iterations whose batch was shortened by delta_cut */
    float filter_max_err;     /* largest relative error of a stored value seen at re-evaluation */
    double horizon;           /* last horizon */
    float ms_refine;          /* host wall time of the horizon sweeps (inside ms_loop) */
    int32_t n_restarts;       /* 1: an optimistically taken batch failed its order check and the clustering was run again
                                 with delta_cut_fallback (see "delta_cut") */
    int32_t n_compactions;    /* K4: times the live clusters were renumbered densely and the matrix moved (compact.cu) */
    float ms_compact;         /* host wall time of the compactions (inside ms_loop) */
    float ms_loop_kernel;     /* device time of the loop kernel's launches alone (CUDA events around each launch) */
    int32_t loop_launches;    /* launches of the loop kernel (one per segment between horizon raises / compactions) */
} ic_stats;

/* ---- context ---------------------------------------------------------- */
int ic_create(ic_ctx **out, int device);
void ic_destroy(ic_ctx *ctx);
const char *ic_last_error(const ic_ctx *ctx);
/* pinned staging buffer for the flattened [N x D] matrix (the cgo shim copies the
 * Go [][]float32 rows into it; Go pointers never cross the boundary) */
void *ic_pinned_alloc(size_t bytes);
void ic_pinned_free(void *p);
/* knobs: "near_tie_tol" (float, default 1e-5), "center" (0/1, default 1),
 * "gram_mode" (IC_GRAM_*), "loop_blocks" (merge-loop blocks per rank, 0 = auto),
 * "virtual_ranks" (1..8: row-block shards emulated on ONE GPU by one cooperative launch --
 * the same kernel path as the multi-GPU build, for tests), "scan_every" (row rescans are requested every k-th
 * merge-loop iteration, default 4: batching them keeps the scan phase out of most iterations), "loop_mode" (1, default:
 * batched loop -- every iteration takes all merges that are provably the next ones of the reference's sequence, on one
 * GPU or across the ranks of a sharded context; 0: one merge per iteration, also what "virtual_ranks" runs; set it
 * before ic_load, it cannot change once ic_initial_distances has run), "no_replica",
 * "profile_loop", "verbose".
 * Reference arithmetic (none of these may change the result, only the time):
 * "exact" (0/1, default 1: every stored value at or below the horizon is the reference's own WardDistance),
 * "horizon_factor" (> 1, default 1.18) / "horizon_factor_first" (0 = the same), "eps_filter" (3e-5), "delta_cut" (0 =
 * optimistic batches with the order check) / "delta_cut_fallback" (1e-5, used after a failed check), "refill_at" (1/2),
 * "near_lists" (0/1), "abs_slack".
 * K4: "compact" (0/1), "compact_ratio" (0.25..0.9, default 0.7), "compact_min" (no compaction below this many slots,
 * default 4096), "compact_tiles" (1: one-pass tile kernel on an unsharded context; 0: the two-pass path sharded runs use),
 * "mirror_init" (mirror pass after K1), "fast_start" (1: ic_run_resident / ic_cluster_with_constraints skip the first
 * nearest-neighbour sweep when the horizon and near lists are on -- the mirror pass collects the global minimum). */
int ic_set_option(ic_ctx *ctx, const char *name, double value);

/* ---- CalculateOptimalClusters, clustering.go:168-186 (host, exact) ---- */
int ic_optimal_clusters(int64_t total_items, int64_t min_size, int64_t max_size, int64_t *out);

/* ---- PerformClusteringWithConstraints, clustering.go:198-284 ----------
 * x: row-major [n x d] fp32 on the HOST (ldx = row stride in floats, >= d).
 * cluster_offsets: capacity n+1; members: capacity n; ids are 0..n_clusters-1 in
 * the reference's slice order, members in the reference's order (hi ++ lo,
 * clustering.go:31); items of clusters below min_size are absent (:268-271).
 * Returns IC_ERR_TOO_FEW / IC_ERR_UNSAT where the reference returns (nil,false). */
int ic_cluster_with_constraints(ic_ctx *ctx, const float *x, int64_t n, int64_t d, int64_t ldx,
                                int64_t min_size, int64_t max_size, int32_t *cluster_offsets,
                                int32_t *members, int32_t *n_clusters, ic_stats *stats);

/* ---- staged entry points (unit parity with the reference's functions) -- */
/* upload X (host pointer) or adopt a device pointer (copied), then K0 prep */
int ic_load(ic_ctx *ctx, const float *x_host, int64_t n, int64_t d, int64_t ldx);
int ic_load_device(ic_ctx *ctx, const float *x_dev, int64_t n, int64_t d, int64_t ldx);
/* Input formation on the device -- the step right before the path:
 * GenerateLabelVector + CombineEmbeddings, internal/embeddings/embeddings.go:166-183, called per item at
 * internal/workflow/workflow.go:167-168.  Row i of X = image embedding i (d_img floats) ++ a vector over the label
 * set (n_labels floats) holding 1.0 at the index of every label of item i that is in the set, 0.0 elsewhere.
 * label_ids[label_offsets[i] .. label_offsets[i+1]) are item i's labels as indices into the label set (the shim does
 * the map[string]int lookup of embeddings.go:169); -1 stands for a label that is not in the set and is ignored, as the
 * reference ignores it; any other index outside [0, n_labels) is IC_ERR_BAD_ARG.  Only the image block crosses PCIe. */
int ic_load_combined(ic_ctx *ctx, const float *img_host, int64_t n, int64_t d_img, int64_t ld_img,
                     const int32_t *label_offsets, const int32_t *label_ids, int64_t n_labels);
/* the resident X [n x d] back on the host (row stride ld >= d) -- inspection / tests */
int ic_read_x(ic_ctx *ctx, float *out_host, int64_t ld);
/* ComputeInitialDistanceMatrix + WardDistance + DotFloat32, clustering.go:61-73,136-157 */
int ic_initial_distances(ic_ctx *ctx, int mode, int64_t max_size);
/* replace the resident matrix (host [n x n], row stride ld) -- test hook */
int ic_set_matrix(ic_ctx *ctx, const float *m_host, int64_t ld);
/* first NN sweep over the matrix (row caches used by FindClosestClusters) */
int ic_nn_init(ic_ctx *ctx);
/* FindClosestClusters, clustering.go:119-133, on the resident state: returns the
 * keys (slice-order ids) of the pair and its distance; key_hi = -1 if none */
int ic_find_closest(ic_ctx *ctx, int32_t *key_hi, int32_t *key_lo, float *dist);
/* merge loop (clustering.go:220-246: FindClosestClusters + maxSize check +
 * MergeClusters + UpdateDistanceMatrix); stops at n_target clusters, on
 * exhaustion, or after max_merges (<0: no limit) */
int ic_merge_loop(ic_ctx *ctx, int64_t min_size, int64_t max_size, int64_t max_merges);
/* all of the above on resident data + trace read-back + host assembly */
int ic_run_resident(ic_ctx *ctx, int64_t min_size, int64_t max_size, int32_t *cluster_offsets,
                    int32_t *members, int32_t *n_clusters, ic_stats *stats);
/* output assembly, clustering.go:265-280, from the merge trace */
int ic_build_clusters(ic_ctx *ctx, int64_t min_size, int32_t *cluster_offsets, int32_t *members,
                      int32_t *n_clusters);

/* ---- row-block sharding of ONE clustering across the GPUs of a box ----------
 * (BASELINE north_star: "the distance matrix is row-block sharded across the 8 GPUs";
 * the reference has no counterpart: its [][]float32 matrix lives in one process,
 * clustering.go:61-73.)  One process per GPU; rank r keeps the rows of the slots
 * [r*C, (r+1)*C), C = ceil(n / world) rounded up to a multiple of 4, with all their columns.  Every rank makes the
 * same calls with the same X, min/max; every rank gets the complete result.
 *   ic_shard_init(ctx, rank, world)            before ic_load
 *   ic_load(ctx, x, ...)                        the whole X on every rank (replicated, <= 2 GB)
 *   ic_shard_export(ctx, blob)                  CUDA-IPC handles of my row block + rank mailbox
 *   <host all-gathers the world blobs, rank order>   (torch.distributed / MPI / Go net: plumbing)
 *   ic_shard_connect(ctx, blobs)                peer-maps the other ranks' row blocks over NVLink
 *   ic_run_resident(...) / ic_merge_loop(...)   as on one GPU; the ranks' persistent kernels exchange their candidate
 *                                               pairs once per iteration (dozens of merges) through peer memory
 * ic_read_matrix / ic_set_matrix touch only the rows [row_begin,row_end) of ic_shard_rows. */
#define IC_SHARD_HANDLE_BYTES 256
int ic_shard_init(ic_ctx *ctx, int rank, int world);
int ic_shard_export(ic_ctx *ctx, void *handle);
int ic_shard_connect(ic_ctx *ctx, const void *handles);
int ic_shard_rows(ic_ctx *ctx, int64_t *row_begin, int64_t *row_end);

/* ---- inspection -------------------------------------------------------- */
/* distance matrix by slot, host [n x n] row stride ld; dead slots hold stale values */
int ic_read_matrix(ic_ctx *ctx, float *out_host, int64_t ld);
/* per slot: key (-1 = retired) and size */
int ic_read_slots(ic_ctx *ctx, int32_t *key, int32_t *size);
/* merge trace: one entry per merge (capacity >= n): keys of the merged pair, the
 * merge distance, the new size, and the relative gap to the runner-up candidate */
int ic_get_merge_trace(ic_ctx *ctx, int32_t *key_hi, int32_t *key_lo, float *dist, int32_t *size,
                       float *gap, int64_t capacity, int64_t *n_merges);
/* the same trace as a dendrogram in the layout of scipy.cluster.hierarchy.linkage (the reference has no
 * on-disk format): row t = {key_lo, key_hi, sqrt(2 * dist), size} as doubles, z[capacity][4].  Keys are item indices
 * (< n) or n + t' for the cluster made by merge t' -- scipy's own numbering -- and sqrt(2 d) is scipy's Ward height, so
 * an unconstrained run reproduces linkage(X, "ward"); with constraints it is the forest the path built (n_final roots),
 * which a caller can re-cut at other min / max sizes without clustering again. */
int ic_get_linkage(ic_ctx *ctx, double *z, int64_t capacity, int64_t *n_rows);
int ic_get_stats(ic_ctx *ctx, ic_stats *stats);
/* debug (option "profile_loop" = 1): SM cycles block 0 spent in each phase of the merge loop
 * {publish, exchange poll + fold, decision + update, row scans, partial folds, merges, iterations, rescans,
 *  0, bubbles ...} of the last launch */
int ic_get_loop_profile(ic_ctx *ctx, int64_t *out16);
/* debug (profile_loop = 1): SM cycles every block of (local) rank 0 spent waiting for the slowest record of the
 * exchanges of the last launch -- the block with the smallest wait is the one the others waited for */
int ic_get_loop_block_waits(ic_ctx *ctx, int64_t *out, int64_t capacity, int64_t *n_blocks);

/* microbenchmarks used by bench.py's roofline legs: one launch of the named
 * kernel on the resident problem, device time in ms */
int ic_time_kernel(ic_ctx *ctx, const char *which, int repeats, float *ms_each);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* IMAGECLUST_B200_H */
