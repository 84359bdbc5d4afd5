/*
 * oracle/oracle.h -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 * C interface of the CPU oracle for internal/clustering/clustering.go.
 * Loaded only by tests/, __graft_entry__.smoke() and bench.py's CPU legs.
 * PARITY UNPINNED by the reference's own tests (it has none); see ward_literal.c.
 */
#ifndef IMAGECLUST_ORACLE_H
#define IMAGECLUST_ORACLE_H

#ifdef __cplusplus
extern "C" {
#endif

#define ORACLE_ERR_TOO_FEW (-1)  /* clustering.go:169-171 */
#define ORACLE_ERR_UNSAT (-2)    /* clustering.go:175-177 */
#define ORACLE_ERR_BAD_ARG (-3)  /* min/max < 1: Go behaviour implementation defined */
#define ORACLE_ERR_INTERNAL (-9)

/* flags of oracle_fast_cluster */
#define ORACLE_FAST_EAGER 1 /* mask inadmissible pairs when written instead of rejecting lazily */
#define ORACLE_FAST_LW 2    /* Lance-Williams update (double arithmetic, fp32 store) instead of centroids */
#define ORACLE_FAST_LW32 4  /* with ORACLE_FAST_LW: do the recurrence in fp32 */

typedef struct {
    int n_target;      /* CalculateOptimalClusters result */
    int n_merges;      /* merges performed */
    long n_rejections; /* maxSize rejections (clustering.go:228-234); 0 in eager mode */
    int exhausted;     /* 1 if the loop ended through clustering.go:222-225 */
    int n_final;       /* clusters left when the loop ended */
    int n_out;         /* clusters in the output map (size >= minSize) */
} oracle_stats;

/* One entry per merge; arrays need capacity N. key = monotone order id: item
 * index for singletons, N + t for the cluster created by merge t; key order is
 * the reference's slice order. pos_* are slice positions
 * (literal oracle only; -1 from the fast oracle). */
typedef struct {
    int *key_hi;
    int *key_lo;
    int *pos_i;
    int *pos_j;
    float *dist;
    int *size;
    float *gap; /* fast oracle only, may be NULL: (second best - best) / best */
} oracle_trace;

float oracle_dot_f32(const float *a, const float *b, int d);
float oracle_ward_distance(const float *ca, long size_a, const float *cb, long size_b, int d);
void oracle_merge_centroid(const float *ca, int size_a, const float *cb, int size_b, int d, float *out);
int oracle_optimal_clusters(long total, long min_size, long max_size, long *out);
void oracle_find_closest(float *const *m, int n, int *out_i, int *out_j);

int oracle_literal_cluster(const float *x, int n_items, int d, int min_size, int max_size,
                           int *offsets, int *members, int *n_out, oracle_trace *tr,
                           float *init_matrix, float *final_matrix, int *final_keys,
                           oracle_stats *st);

int oracle_fast_cluster(const float *x, int n_items, int d, int min_size, int max_size, int flags,
                        int n_threads, int *offsets, int *members, int *n_out, oracle_trace *tr,
                        oracle_stats *st);

/* as oracle_fast_cluster; init_matrix (N x N, may be NULL) replaces the matrix
 * computed from x -- used to replay the device's own initial distances. */
int oracle_fast_cluster_ex(const float *x, int n_items, int d, int min_size, int max_size, int flags,
                           int n_threads, const float *init_matrix, int *offsets, int *members,
                           int *n_out, oracle_trace *tr, oracle_stats *st);

/* Initial Ward matrix only (ComputeInitialDistanceMatrix, clustering.go:61-73),
 * multi-threaded, full symmetric N x N, bit-identical to the literal path. */
int oracle_initial_matrix(const float *x, int n_items, int d, int n_threads, float *out);

/* CPU restatement of the DEVICE's merge-loop algorithm (ward_device.c): Lance-Williams values above a horizon, the
 * reference's own centroid values below it, batches of consecutive merges.  Must reproduce oracle_fast_cluster(flags=0). */
typedef struct {
    long n_iterations;    /* batches */
    long n_exact;         /* pairs evaluated with the reference's arithmetic (WardDistance of two centroids) */
    long n_raises;        /* horizon raises */
    long n_cut;           /* batches shortened by delta_cut */
    long n_violations;    /* pairs created inside a batch that came out below a later member of it (must be 0) */
    double max_filter_err; /* largest |stored - reference| / reference seen when a pair was re-evaluated */
    double horizon;
} oracle_device_stats;

int oracle_device_cluster(const float *x, int n_items, int d, int min_size, int max_size, const float *init_matrix,
                          double horizon_factor, double eps_filter, double delta_cut, int max_batch, int n_threads,
                          int *offsets, int *members, int *n_out, oracle_trace *tr, oracle_stats *st,
                          oracle_device_stats *ds);

#ifdef __cplusplus
}
#endif
#endif
