/*
 * oracle/ward_fast.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * Same semantics as oracle/ward_literal.c (the operation-for-operation
 * restatement of internal/clustering/clustering.go) but with O(N^2) memory
 * traffic per run instead of O(N^3): a per-row nearest-neighbour cache, slot
 * reuse instead of physical row/column deletion, and OpenMP over the
 * centroid distance evaluations.  Every distance is still produced by the
 * reference's arithmetic (sequential fp32 dot, clustering.go:136-157; centroid
 * merge clustering.go:39), so traces are bit-identical to the literal oracle;
 * tests/test_oracle_fast.py proves that for N up to a few thousand.  It is the
 * parity source at sizes the literal oracle cannot finish.
 *
 * PARITY UNPINNED by the reference's own tests (it has none); see ward_literal.c.
 *
 * Ordering argument: the reference's slice always holds the
 * survivors in original relative order followed by merged clusters in merge
 * order (clustering.go:240-241), so slice position order == order of a
 * monotone key (item index for singletons, N+t for the t-th merge).  The scan
 * at clustering.go:123-131 returns the lexicographically smallest
 * (d, i, j), i > j positions, hence the smallest (d, key_hi, key_lo).
 *
 * Optional modes, used to validate the device algorithm on the CPU:
 *   ORACLE_FAST_EAGER  inadmissible pairs (size sum > maxSize, clustering.go:228)
 *                      are masked when written instead of rejected lazily.
 *   ORACLE_FAST_LW     new row by the Lance-Williams recurrence from the two
 *                      old rows (the device's K3) instead of from centroids.
 */
#include <float.h>
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "oracle.h"

/* squared distance with the reference's rounding: diff rounded to fp32, then
 * sum += diff*diff in index order (clustering.go:138-141,152-155).  Four rows
 * at a time so the four add chains overlap; each chain is still sequential. */
static void dsq_rows4(const float *c0, const float *c1, const float *c2, const float *c3,
                      const float *cn, int d, float out[4])
{
    float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
    for (int i = 0; i < d; i++) {
        float b = cn[i];
        float d0 = c0[i] - b, d1 = c1[i] - b, d2 = c2[i] - b, d3 = c3[i] - b;
        float p0 = d0 * d0, p1 = d1 * d1, p2 = d2 * d2, p3 = d3 * d3;
        s0 = s0 + p0;
        s1 = s1 + p1;
        s2 = s2 + p2;
        s3 = s3 + p3;
    }
    out[0] = s0;
    out[1] = s1;
    out[2] = s2;
    out[3] = s3;
}

static inline float ward_weight(long sa, long sb, float dsq)
{
    float num = (float)(sa * sb); /* clustering.go:142 */
    float den = (float)(sa + sb); /* clustering.go:143 */
    return (num / den) * dsq;     /* clustering.go:144 */
}

/* out[q] = WardDistance(cluster ks[q], new cluster) for q < nk */
static void ward_to_many(const float *cents, int d, const int *ks, int nk, const int *size,
                         const float *cn, long size_new, float *out, int n_threads)
{
    int nblk = (nk + 3) / 4;
    (void)n_threads;
#pragma omp parallel for schedule(static) num_threads(n_threads)
    for (int blk = 0; blk < nblk; blk++) {
        int q0 = blk * 4;
        int idx[4];
        for (int u = 0; u < 4; u++)
            idx[u] = ks[(q0 + u < nk) ? q0 + u : nk - 1];
        float r[4];
        dsq_rows4(cents + (size_t)idx[0] * d, cents + (size_t)idx[1] * d, cents + (size_t)idx[2] * d,
                  cents + (size_t)idx[3] * d, cn, d, r);
        for (int u = 0; u < 4 && q0 + u < nk; u++)
            out[q0 + u] = ward_weight(size[idx[u]], size_new, r[u]);
    }
}

int oracle_initial_matrix(const float *x, int n, int d, int n_threads, float *out)
{
    if (n_threads < 1)
        n_threads = 1;
#pragma omp parallel for schedule(dynamic, 8) num_threads(n_threads)
    for (int i = 0; i < n; i++) {
        const float *ci = x + (size_t)i * d;
        out[(size_t)i * n + i] = 0.0f;
        for (int j0 = 0; j0 < i; j0 += 4) {
            int idx[4];
            for (int u = 0; u < 4; u++)
                idx[u] = (j0 + u < i) ? j0 + u : i - 1;
            float r[4];
            /* WardDistance(clusters[i], clusters[j]): diff = c_i - c_j; the
             * square is sign independent so evaluating c_j - c_i is identical */
            dsq_rows4(x + (size_t)idx[0] * d, x + (size_t)idx[1] * d, x + (size_t)idx[2] * d,
                      x + (size_t)idx[3] * d, ci, d, r);
            for (int u = 0; u < 4 && j0 + u < i; u++) {
                float v = ward_weight(1, 1, r[u]);
                out[(size_t)i * n + (j0 + u)] = v;
                out[(size_t)(j0 + u) * n + i] = v;
            }
        }
    }
    return 0;
}

typedef struct {
    int n, d, max_size, flags;
    float *m;   /* n x n symmetric, slot indexed */
    int *key;   /* -1 when the slot is retired */
    int *size;
    float *nn_d; /* row cache over partners with LOWER key */
    int *nn_key; /* partner key, -1 if the row has no selectable partner */
    int *slot_of_key;
} fast_state;

/* (d1,k1) < (d2,k2) lexicographically */
static inline int cand_less(float d1, int k1, float d2, int k2)
{
    return d1 < d2 || (d1 == d2 && k1 < k2);
}

static void rescan_row(fast_state *s, int r)
{
    const float *row = s->m + (size_t)r * s->n;
    int kr = s->key[r];
    float best = FLT_MAX; /* clustering.go:120: only entries < MaxFloat32 can win */
    int bk = -1;
    for (int u = 0; u < s->n; u++) {
        int ku = s->key[u];
        if (ku < 0 || ku >= kr)
            continue;
        float v = row[u];
        if (v < best || (v == best && bk >= 0 && ku < bk)) {
            best = v;
            bk = ku;
        }
    }
    s->nn_d[r] = best;
    s->nn_key[r] = bk;
}

int oracle_fast_cluster_ex(const float *x, int n, int d, int min_size, int max_size, int flags,
                           int n_threads, const float *init_matrix, int *offsets, int *members,
                           int *n_out, oracle_trace *tr, oracle_stats *st)
{
    oracle_stats local;
    if (!st)
        st = &local;
    memset(st, 0, sizeof(*st));
    if (n_out)
        *n_out = 0;
    if (n_threads < 1)
        n_threads = 1;
    if (flags & ORACLE_FAST_LW)
        flags |= ORACLE_FAST_EAGER; /* MaxFloat32 markers would poison the recurrence */
    const int eager = (flags & ORACLE_FAST_EAGER) != 0;
    const int lw = (flags & ORACLE_FAST_LW) != 0;
    const int lw32 = (flags & ORACLE_FAST_LW32) != 0;

    long n_target = 0;
    int rc = oracle_optimal_clusters(n, min_size, max_size, &n_target);
    if (rc != 0)
        return rc;
    st->n_target = (int)n_target;

    fast_state s;
    s.n = n;
    s.d = d;
    s.max_size = max_size;
    s.flags = flags;
    size_t nn = (size_t)(n > 0 ? n : 1);
    s.m = (float *)malloc(sizeof(float) * nn * nn);
    s.key = (int *)malloc(sizeof(int) * nn);
    s.size = (int *)malloc(sizeof(int) * nn);
    s.nn_d = (float *)malloc(sizeof(float) * nn);
    s.nn_key = (int *)malloc(sizeof(int) * nn);
    s.slot_of_key = (int *)malloc(sizeof(int) * 2 * nn);
    float *cents = NULL;
    if (!lw) {
        cents = (float *)malloc(sizeof(float) * nn * (size_t)(d > 0 ? d : 1));
        memcpy(cents, x, sizeof(float) * (size_t)n * (size_t)d);
    }
    int *child_hi = (int *)malloc(sizeof(int) * nn);
    int *child_lo = (int *)malloc(sizeof(int) * nn);
    int *ks = (int *)malloc(sizeof(int) * nn);
    float *newrow = (float *)malloc(sizeof(float) * nn);
    float *cnew = (float *)malloc(sizeof(float) * (size_t)(d > 0 ? d : 1));
    if (!s.m || !s.key || !s.size || !s.nn_d || !s.nn_key || !s.slot_of_key || !child_hi || !child_lo ||
        !ks || !newrow || !cnew || (!lw && !cents))
        return ORACLE_ERR_INTERNAL;

    if (init_matrix)
        memcpy(s.m, init_matrix, sizeof(float) * (size_t)n * (size_t)n);
    else
        oracle_initial_matrix(x, n, d, n_threads, s.m);
    for (int i = 0; i < n; i++) {
        s.key[i] = i;
        s.size[i] = 1;
        s.slot_of_key[i] = i;
    }
    if (eager && 2 > max_size)
        for (size_t q = 0; q < (size_t)n * n; q++)
            s.m[q] = INFINITY;
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (int i = 0; i < n; i++)
        rescan_row(&s, i);

    int n_live = n, t = 0;
    while (n_live > n_target) { /* clustering.go:220 */
        /* FindClosestClusters (clustering.go:119-133) over the row caches */
        float bd = FLT_MAX, sd = INFINITY; /* best, second best (another row) */
        int bhi = -1, blo = -1;
        for (int r = 0; r < n; r++) {
            if (s.key[r] < 0 || s.nn_key[r] < 0)
                continue;
            float v = s.nn_d[r];
            int khi = s.key[r], klo = s.nn_key[r];
            if (bhi < 0 ? (v < bd) : (v < bd || (v == bd && (khi < bhi || (khi == bhi && klo < blo))))) {
                if (bhi >= 0)
                    sd = bd;
                bd = v;
                bhi = khi;
                blo = klo;
            } else if (v < sd) {
                sd = v;
            }
        }
        if (bhi < 0) { /* clustering.go:222-225 */
            st->exhausted = 1;
            break;
        }
        int a = s.slot_of_key[bhi], b = s.slot_of_key[blo];
        int sa = s.size[a], sb = s.size[b];
        if (sa + sb > max_size) { /* clustering.go:228-234 (never taken in eager mode) */
            s.m[(size_t)a * n + b] = FLT_MAX;
            s.m[(size_t)b * n + a] = FLT_MAX;
            st->n_rejections++;
            rescan_row(&s, a);
            continue;
        }
        if (tr) {
            tr->key_hi[t] = bhi;
            tr->key_lo[t] = blo;
            tr->pos_i[t] = -1;
            tr->pos_j[t] = -1;
            tr->dist[t] = bd;
            tr->size[t] = sa + sb;
            if (tr->gap)
                tr->gap[t] = (sd - bd) / bd;
        }
        child_hi[t] = bhi;
        child_lo[t] = blo;

        /* survivors other than a, b */
        int nk = 0;
        for (int u = 0; u < n; u++)
            if (s.key[u] >= 0 && u != a && u != b)
                ks[nk++] = u;

        int snew = sa + sb;
        if (lw) {
            const float *ra = s.m + (size_t)a * n, *rb = s.m + (size_t)b * n;
            for (int q = 0; q < nk; q++) {
                int k = ks[q];
                int sk = s.size[k];
                float v;
                if (sk + snew > max_size) {
                    v = INFINITY;
                } else if (lw32) {
                    float num = (float)(sa + sk) * ra[k] + (float)(sb + sk) * rb[k] - (float)sk * bd;
                    v = num / (float)(snew + sk);
                } else {
                    double t1 = (double)(sa + sk) * (double)ra[k];
                    double t2 = (double)(sb + sk) * (double)rb[k];
                    double t3 = (double)sk * (double)bd;
                    double num = (t1 + t2) - t3;
                    /* product with the correctly rounded reciprocal: the device keeps a table of 1.0 / size sum */
                    v = (float)(num * (1.0 / (double)(snew + sk)));
                }
                if (!(v >= 0.0f))
                    v = (v != v) ? INFINITY : 0.0f;
                newrow[q] = v;
            }
        } else {
            /* MergeClusters centroid, clustering.go:39: a is the larger position */
            oracle_merge_centroid(cents + (size_t)a * d, sa, cents + (size_t)b * d, sb, d, cnew);
            ward_to_many(cents, d, ks, nk, s.size, cnew, snew, newrow, n_threads);
            if (eager)
                for (int q = 0; q < nk; q++)
                    if (s.size[ks[q]] + snew > max_size)
                        newrow[q] = INFINITY;
            memcpy(cents + (size_t)b * d, cnew, sizeof(float) * (size_t)d);
        }

        /* retire a, reuse slot b for the merged cluster (highest key so far) */
        s.key[a] = -1;
        s.key[b] = n + t;
        s.size[b] = snew;
        s.slot_of_key[n + t] = b;
        float nb = FLT_MAX;
        int nbk = -1;
        for (int q = 0; q < nk; q++) {
            int k = ks[q];
            float v = newrow[q];
            s.m[(size_t)b * n + k] = v;
            s.m[(size_t)k * n + b] = v;
            if (v < nb || (v == nb && nbk >= 0 && s.key[k] < nbk)) {
                nb = v;
                nbk = s.key[k];
            }
        }
        s.nn_d[b] = nb;
        s.nn_key[b] = nbk;
        /* rows whose cached partner died need a rescan */
        for (int q = 0; q < nk; q++) {
            int k = ks[q];
            if (s.nn_key[k] == bhi || s.nn_key[k] == blo)
                rescan_row(&s, k);
        }
        n_live--;
        t++;
    }
    st->n_merges = t;
    st->n_final = n_live;

    /* output assembly, clustering.go:265-280: slice order == key order */
    int *order = ks; /* reuse */
    int nf = 0;
    /* singletons that survived come first (keys < n in index order), then merged by key */
    for (int k = 0; k < n + t; k++) {
        int slot = s.slot_of_key[k];
        if (s.key[slot] == k)
            order[nf++] = k;
    }
    int cid = 0, pos = 0;
    if (offsets)
        offsets[0] = 0;
    int *stack = (int *)malloc(sizeof(int) * (nn + 1));
    for (int q = 0; q < nf; q++) {
        int k = order[q];
        int slot = s.slot_of_key[k];
        if (s.size[slot] > max_size)
            rc = ORACLE_ERR_INTERNAL;
        if (s.size[slot] < min_size)
            continue; /* clustering.go:268-271 */
        /* members(hi) ++ members(lo), clustering.go:31,237 */
        int sp = 0;
        stack[sp++] = k;
        while (sp > 0) {
            int kk = stack[--sp];
            if (kk < n) {
                if (members)
                    members[pos] = kk;
                pos++;
            } else {
                stack[sp++] = child_lo[kk - n];
                stack[sp++] = child_hi[kk - n];
            }
        }
        cid++;
        if (offsets)
            offsets[cid] = pos;
    }
    free(stack);
    if (n_out)
        *n_out = cid;
    st->n_out = cid;

    free(s.m);
    free(s.key);
    free(s.size);
    free(s.nn_d);
    free(s.nn_key);
    free(s.slot_of_key);
    free(cents);
    free(child_hi);
    free(child_lo);
    free(ks);
    free(newrow);
    free(cnew);
    return rc;
}

int oracle_fast_cluster(const float *x, int n, int d, int min_size, int max_size, int flags,
                        int n_threads, int *offsets, int *members, int *n_out, oracle_trace *tr,
                        oracle_stats *st)
{
    return oracle_fast_cluster_ex(x, n, d, min_size, max_size, flags, n_threads, NULL, offsets, members,
                                  n_out, tr, st);
}
