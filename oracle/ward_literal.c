/*
 * oracle/ward_literal.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement, operation for operation, of the reference's size-constrained
 * Ward clustering (reference: internal/clustering/clustering.go, Go, single
 * threaded, fp32).  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs may load this file's shared object.
 *
 * PARITY UNPINNED: the reference ships no tests, fixtures or golden vectors
 *  and no Go toolchain exists in this image, so this
 * restatement cannot be checked against the reference's own executable.  It is
 * pinned instead against (a) an independent numpy restatement
 * (oracle/numpy_literal.py), (b) scipy's Ward linkage on unconstrained inputs
 * and (c) hand-computed small cases -- see tests/test_oracle_*.py.
 *
 * Rules followed (each cited at the function that applies it):
 *   - fp32 storage, every fp32 operation rounded separately (no FMA
 *     contraction: compile with -ffp-contract=off; Go/amd64 GOAMD64=v1 emits
 *     separate MULSS/ADDSS);
 *   - sequential dot product in index order       clustering.go:152-155
 *   - int product before the float conversion     clustering.go:142
 *   - strict '<' scan, rows then columns           clustering.go:123-131
 *   - delete larger position first, append new last clustering.go:51-58,100-116,240-241
 *   - members = hi-position members ++ lo-position  clustering.go:31,237
 *   - MaxFloat32 initial minimum and reject marker  clustering.go:120,230-231
 *   - clusters below minSize dropped, dense ids      clustering.go:266-280
 *
 * The data layout deliberately mirrors the reference's slice-of-slices and
 * per-pair temporary (clustering.go:137) so that timing this file is a fair
 * "C restatement of the Go reference" CPU baseline.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include "oracle.h"

/* ---- Cluster (clustering.go:11-15) ------------------------------------ */
typedef struct {
    int *indices;    /* Indices */
    int n_indices;
    int size;        /* Size */
    float *centroid; /* Centroid, D floats */
    int key;         /* instrumentation only: monotone order id (not in reference) */
} lit_cluster;

/* DotFloat32, clustering.go:148-157: sum += a[i]*b[i], i ascending, fp32. */
float oracle_dot_f32(const float *a, const float *b, int d)
{
    float sum = 0.0f;
    for (int i = 0; i < d; i++) {
        float p = a[i] * b[i];
        sum = sum + p;
    }
    return sum;
}

/* WardDistance, clustering.go:136-145. Allocates the diff temp per pair like
 * the reference does (:137). */
float oracle_ward_distance(const float *ca, long size_a, const float *cb, long size_b, int d)
{
    float *diff = (float *)malloc(sizeof(float) * (size_t)(d > 0 ? d : 1));
    for (int i = 0; i < d; i++)
        diff[i] = ca[i] - cb[i];
    float dsq = oracle_dot_f32(diff, diff, d);
    free(diff);
    float num = (float)(size_a * size_b); /* integer product first, :142 */
    float den = (float)(size_a + size_b); /* :143 */
    return (num / den) * dsq;             /* :144 */
}

/* Centroid of a merge, clustering.go:36-40. */
void oracle_merge_centroid(const float *ca, int size_a, const float *cb, int size_b, int d, float *out)
{
    int size = size_a + size_b;
    float fa = (float)size_a, fb = (float)size_b, fs = (float)size;
    for (int i = 0; i < d; i++) {
        float pa = fa * ca[i];
        float pb = fb * cb[i];
        float s = pa + pb;
        out[i] = s / fs;
    }
}

/* CalculateOptimalClusters, clustering.go:168-186.
 * returns 0 ok; ORACLE_ERR_TOO_FEW (:169-171); ORACLE_ERR_UNSAT (:175-177);
 * ORACLE_ERR_BAD_ARG for min/max < 1 where Go's float->int conversion of +-Inf
 * is implementation defined (:173-174) -- the drop-in rejects those up front. */
int oracle_optimal_clusters(long total, long min_size, long max_size, long *out)
{
    if (min_size < 1 || max_size < 1 || total < 0)
        return ORACLE_ERR_BAD_ARG;
    if (total < min_size)
        return ORACLE_ERR_TOO_FEW;
    long lo = (long)ceil((double)total / (double)max_size);
    long hi = (long)floor((double)total / (double)min_size);
    if (lo > hi)
        return ORACLE_ERR_UNSAT;
    long n = lo;
    if (lo < hi)
        n = (lo + hi) / 2;
    *out = n;
    return 0;
}

/* FindClosestClusters, clustering.go:119-133. */
void oracle_find_closest(float *const *m, int n, int *out_i, int *out_j)
{
    float best = FLT_MAX;
    int bi = -1, bj = -1;
    for (int i = 0; i < n; i++) {
        const float *row = m[i];
        for (int j = 0; j < i; j++) {
            if (row[j] < best) {
                best = row[j];
                bi = i;
                bj = j;
            }
        }
    }
    *out_i = bi;
    *out_j = bj;
}

static void free_cluster(lit_cluster *c)
{
    free(c->indices);
    free(c->centroid);
    c->indices = NULL;
    c->centroid = NULL;
}

/* PerformClusteringWithConstraints, clustering.go:198-284. */
int oracle_literal_cluster(const float *x, int n_items, int d, int min_size, int max_size,
                           int *offsets, int *members, int *n_out, oracle_trace *tr,
                           float *init_matrix, float *final_matrix, int *final_keys,
                           oracle_stats *st)
{
    oracle_stats local;
    if (!st)
        st = &local;
    memset(st, 0, sizeof(*st));
    if (n_out)
        *n_out = 0;

    long n_target = 0;
    int rc = oracle_optimal_clusters(n_items, min_size, max_size, &n_target); /* :203 */
    if (rc != 0)
        return rc; /* reference: log + return nil,false (:204-207) */
    st->n_target = (int)n_target;

    int n = n_items;
    /* :211-214 one singleton per item, centroid = copy of the row (:19-20) */
    lit_cluster *cl = (lit_cluster *)calloc((size_t)(n > 0 ? n : 1), sizeof(lit_cluster));
    for (int i = 0; i < n; i++) {
        cl[i].indices = (int *)malloc(sizeof(int));
        cl[i].indices[0] = i;
        cl[i].n_indices = 1;
        cl[i].size = 1;
        cl[i].centroid = (float *)malloc(sizeof(float) * (size_t)(d > 0 ? d : 1));
        memcpy(cl[i].centroid, x + (size_t)i * (size_t)d, sizeof(float) * (size_t)d);
        cl[i].key = i;
    }

    /* ComputeInitialDistanceMatrix, :61-73: one row slice per cluster, both
     * triangles written, diagonal zero. */
    float **m = (float **)calloc((size_t)(n > 0 ? n : 1), sizeof(float *));
    for (int i = 0; i < n; i++) {
        m[i] = (float *)calloc((size_t)n, sizeof(float));
        for (int j = 0; j < i; j++) {
            float dist = oracle_ward_distance(cl[i].centroid, cl[i].size, cl[j].centroid, cl[j].size, d);
            m[i][j] = dist;
            m[j][i] = dist;
        }
    }
    if (init_matrix)
        for (int i = 0; i < n; i++)
            memcpy(init_matrix + (size_t)i * (size_t)n, m[i], sizeof(float) * (size_t)n);

    int t = 0; /* merges done */
    while (n > n_target) { /* :220 */
        int i, j;
        oracle_find_closest(m, n, &i, &j); /* :221 */
        if (i == -1 || j == -1) {          /* :222-225 */
            st->exhausted = 1;
            break;
        }
        if (cl[i].size + cl[j].size > max_size) { /* :228-234 */
            m[i][j] = FLT_MAX;
            m[j][i] = FLT_MAX;
            st->n_rejections++;
            continue;
        }
        float d_ij = m[i][j];

        /* MergeClusters(clusters[i], clusters[j]), :29-47, i is the larger position */
        lit_cluster nc;
        nc.n_indices = cl[i].n_indices + cl[j].n_indices;
        nc.indices = (int *)malloc(sizeof(int) * (size_t)nc.n_indices);
        memcpy(nc.indices, cl[i].indices, sizeof(int) * (size_t)cl[i].n_indices);
        memcpy(nc.indices + cl[i].n_indices, cl[j].indices, sizeof(int) * (size_t)cl[j].n_indices);
        nc.size = cl[i].size + cl[j].size;
        nc.centroid = (float *)malloc(sizeof(float) * (size_t)(d > 0 ? d : 1));
        oracle_merge_centroid(cl[i].centroid, cl[i].size, cl[j].centroid, cl[j].size, d, nc.centroid);
        nc.key = n_items + t;

        if (tr) {
            tr->key_hi[t] = cl[i].key;
            tr->key_lo[t] = cl[j].key;
            tr->pos_i[t] = i;
            tr->pos_j[t] = j;
            tr->dist[t] = d_ij;
            tr->size[t] = nc.size;
        }

        /* RemoveClusters, :51-58 (larger position first), then append, :241 */
        free_cluster(&cl[i]);
        free_cluster(&cl[j]);
        memmove(&cl[i], &cl[i + 1], sizeof(lit_cluster) * (size_t)(n - i - 1));
        memmove(&cl[j], &cl[j + 1], sizeof(lit_cluster) * (size_t)(n - 1 - j - 1));
        cl[n - 2] = nc;

        /* RemoveRowsAndColumns, :100-116 */
        for (int r = 0; r < n; r++) {
            float *row = m[r];
            memmove(row + i, row + i + 1, sizeof(float) * (size_t)(n - i - 1));
            memmove(row + j, row + j + 1, sizeof(float) * (size_t)(n - 1 - j - 1));
        }
        free(m[i]);
        memmove(&m[i], &m[i + 1], sizeof(float *) * (size_t)(n - i - 1));
        free(m[j]);
        memmove(&m[j], &m[j + 1], sizeof(float *) * (size_t)(n - 1 - j - 1));
        n -= 1; /* n is now len(clusters) after the append */

        /* UpdateDistanceMatrix, :81-93: new row from centroids, appended as the
         * last column of every row and as the last row. */
        float *new_row = (float *)calloc((size_t)n, sizeof(float));
        for (int k = 0; k < n - 1; k++)
            new_row[k] = oracle_ward_distance(cl[k].centroid, cl[k].size, nc.centroid, nc.size, d);
        new_row[n - 1] = 0.0f;
        for (int k = 0; k < n - 1; k++) {
            m[k] = (float *)realloc(m[k], sizeof(float) * (size_t)n);
            m[k][n - 1] = new_row[k];
        }
        m[n - 1] = new_row;
        t++;
    }
    st->n_merges = t;
    st->n_final = n;

    if (final_matrix)
        for (int i = 0; i < n; i++)
            memcpy(final_matrix + (size_t)i * (size_t)n, m[i], sizeof(float) * (size_t)n);
    if (final_keys)
        for (int i = 0; i < n; i++)
            final_keys[i] = cl[i].key;

    /* :249-262 oversize split is unreachable (every merge is guarded at :228);
     * assert instead of restating the dead, buggy splitter. */
    for (int i = 0; i < n; i++)
        if (cl[i].size > max_size)
            rc = ORACLE_ERR_INTERNAL;

    /* :265-280 output in slice order, skipping clusters below minSize */
    int cid = 0, pos = 0;
    if (offsets)
        offsets[0] = 0;
    for (int i = 0; i < n; i++) {
        if (cl[i].size < min_size)
            continue;
        if (members)
            memcpy(members + pos, cl[i].indices, sizeof(int) * (size_t)cl[i].n_indices);
        pos += cl[i].n_indices;
        cid++;
        if (offsets)
            offsets[cid] = pos;
    }
    if (n_out)
        *n_out = cid;
    st->n_out = cid;

    for (int i = 0; i < n; i++) {
        free_cluster(&cl[i]);
        free(m[i]);
    }
    free(cl);
    free(m);
    return rc;
}
