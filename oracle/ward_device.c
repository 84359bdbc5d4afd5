/*
 * oracle/ward_device.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the DEVICE's merge-loop algorithm (imageclust_b200/csrc/merge_batch.cu),
 * used to prove on the CPU -- against the literal restatement of internal/clustering/clustering.go -- that the
 * procedure the kernel implements yields the reference's merge sequence bit for bit:
 *
 *   * distances are kept by Lance-Williams updates (double arithmetic, fp32 store) -- an APPROXIMATION of the
 *     reference's value, which is WardDistance of the two fp32 centroids (clustering.go:83-86,136-145);
 *   * HORIZON: every pair whose stored value is <= horizon holds the reference's own value (centroids are kept and
 *     merged with clustering.go:39's arithmetic, the pair is re-evaluated with the sequential fp32 dot of
 *     clustering.go:148-157); a pair above the horizon is known to within eps_filter.  Decisions are only taken among
 *     values below horizon / (1 + 2 eps_filter); when the minimum gets there the horizon is raised and the band is
 *     re-evaluated;
 *   * BATCH RULE: heads below the stopper T are merged in one iteration (see oracle/batch_rule.py); because the
 *     reference's fp32 values are reducible only up to rounding, members after the first must also be below
 *     T * (1 - delta_cut);
 *   * eager admissibility (clustering.go:228-234) as in ward_fast.c.
 *
 * PARITY UNPINNED by the reference's own tests (it has none); see ward_literal.c.
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

static float dsq_seq(const float *a, const float *b, int d)
{
    float s = 0.0f;
    for (int i = 0; i < d; i++) {
        float df = a[i] - b[i];
        float p = df * df;
        s = s + p;
    }
    return s;
}
static inline float ward_w(long sa, long sb, float dsq)
{
    float num = (float)(sa * sb);
    float den = (float)(sa + sb);
    return (num / den) * dsq;
}
static inline float lw_double(int sa, int sb, int sk, float dka, float dkb, float dab)
{
    double t1 = (double)(sa + sk) * (double)dka;
    double t2 = (double)(sb + sk) * (double)dkb;
    double t3 = (double)sk * (double)dab;
    double num = (t1 + t2) - t3;
    float v = (float)(num * (1.0 / (double)(sa + sb + sk)));
    if (!(v >= 0.0f))
        v = (v != v) ? INFINITY : 0.0f;
    return v;
}

typedef struct {
    float d1, d2; /* smallest and second smallest selectable value over partners with a lower key */
    int k1, k2;   /* their keys (-1: none) */
} top2;

typedef struct {
    int n, d, max_size;
    float *m;
    int *key, *size, *slot_of_key;
    float *cents;
    top2 *tp;
} dstate;

static inline int cless(float d1, int k1, float d2, int k2) { return d1 < d2 || (d1 == d2 && k1 < k2); }

static void rescan2(dstate *s, int r)
{
    const float *row = s->m + (size_t)r * s->n;
    int kr = s->key[r];
    top2 t = {FLT_MAX, FLT_MAX, -1, -1};
    for (int u = 0; u < s->n; u++) {
        int ku = s->key[u];
        if (ku < 0 || ku >= kr)
            continue;
        float v = row[u];
        if (!(v < FLT_MAX))
            continue;
        if (t.k1 < 0 || cless(v, ku, t.d1, t.k1)) {
            t.d2 = t.d1;
            t.k2 = t.k1;
            t.d1 = v;
            t.k1 = ku;
        } else if (t.k2 < 0 || cless(v, ku, t.d2, t.k2)) {
            t.d2 = v;
            t.k2 = ku;
        }
    }
    s->tp[r] = t;
}

typedef struct {
    float d;
    int khi, klo, a, b;
} headrec;
static int head_cmp(const void *x, const void *y)
{
    const headrec *p = (const headrec *)x, *q = (const headrec *)y;
    if (p->d != q->d)
        return p->d < q->d ? -1 : 1;
    return p->khi < q->khi ? -1 : (p->khi > q->khi ? 1 : 0);
}

int oracle_device_cluster(const float *x, int n, int d, int min_size, int max_size, const float *init_matrix,
                          double horizon_factor, double eps_filter, double delta_cut, int max_batch, int n_threads,
                          int *offsets, int *members, int *n_out, oracle_trace *tr, oracle_stats *st,
                          oracle_device_stats *ds)
{
    oracle_stats local;
    oracle_device_stats dlocal;
    if (!st)
        st = &local;
    if (!ds)
        ds = &dlocal;
    memset(st, 0, sizeof(*st));
    memset(ds, 0, sizeof(*ds));
    if (n_out)
        *n_out = 0;
    if (n_threads < 1)
        n_threads = 1;
    if (max_batch < 1)
        max_batch = 1;
    long n_target = 0;
    int rc = oracle_optimal_clusters(n, min_size, max_size, &n_target);
    if (rc != 0)
        return rc;
    st->n_target = (int)n_target;

    dstate s;
    s.n = n;
    s.d = d;
    s.max_size = max_size;
    size_t nn = (size_t)(n > 0 ? n : 1);
    s.m = (float *)malloc(sizeof(float) * nn * nn);
    s.key = (int *)malloc(sizeof(int) * nn);
    s.size = (int *)malloc(sizeof(int) * nn);
    s.slot_of_key = (int *)malloc(sizeof(int) * 2 * nn);
    s.cents = (float *)malloc(sizeof(float) * nn * (size_t)(d > 0 ? d : 1));
    s.tp = (top2 *)malloc(sizeof(top2) * nn);
    int *child_hi = (int *)malloc(sizeof(int) * nn);
    int *child_lo = (int *)malloc(sizeof(int) * nn);
    headrec *heads = (headrec *)malloc(sizeof(headrec) * nn);
    int *first_touch = (int *)malloc(sizeof(int) * nn);
    int *evq = (int *)malloc(sizeof(int) * nn);
    float *cnew = (float *)malloc(sizeof(float) * (size_t)(d > 0 ? d : 1));
    if (!s.m || !s.key || !s.size || !s.slot_of_key || !s.cents || !s.tp || !child_hi || !child_lo || !heads ||
        !first_touch || !evq || !cnew)
        return ORACLE_ERR_INTERNAL;
    memcpy(s.cents, x, sizeof(float) * (size_t)n * (size_t)d);
    if (init_matrix)
        memcpy(s.m, init_matrix, sizeof(float) * (size_t)n * (size_t)n);
    else
        oracle_initial_matrix(x, n, d, n_threads, s.m);
    for (int i = 0; i < n; i++) {
        s.key[i] = i;
        s.size[i] = 1;
        s.slot_of_key[i] = i;
    }
    if (2 > max_size)
        for (size_t q = 0; q < (size_t)n * n; q++)
            s.m[q] = INFINITY;

    /* the band (lo, hi] of stored values is re-evaluated with the reference's arithmetic */
    double horizon = 0.0;
#define REFINE_BAND(lo_, hi_)                                                                        \
    do {                                                                                             \
        long cnt_ = 0;                                                                               \
        double err_ = 0.0;                                                                           \
        _Pragma("omp parallel for schedule(dynamic, 16) num_threads(n_threads) reduction(+ : cnt_) reduction(max : err_)") \
        for (int r_ = 0; r_ < n; r_++) {                                                             \
            if (s.key[r_] < 0)                                                                       \
                continue;                                                                            \
            for (int u_ = 0; u_ < n; u_++) {                                                         \
                if (s.key[u_] < 0 || s.key[u_] >= s.key[r_])                                         \
                    continue;                                                                        \
                float v_ = s.m[(size_t)r_ * n + u_];                                                 \
                if ((double)v_ > (lo_) && (double)v_ <= (hi_)) {                                     \
                    float w_ = ward_w(s.size[r_], s.size[u_],                                        \
                                      dsq_seq(s.cents + (size_t)r_ * d, s.cents + (size_t)u_ * d, d)); \
                    double e_ = fabs((double)w_ - (double)v_) / fmax((double)w_, 1e-30);             \
                    if (e_ > err_)                                                                   \
                        err_ = e_;                                                                   \
                    s.m[(size_t)r_ * n + u_] = w_;                                                   \
                    s.m[(size_t)u_ * n + r_] = w_;                                                   \
                    cnt_++;                                                                          \
                }                                                                                    \
            }                                                                                        \
        }                                                                                            \
        ds->n_exact += cnt_;                                                                         \
        if (err_ > ds->max_filter_err)                                                               \
            ds->max_filter_err = err_;                                                               \
    } while (0)

    {
        float mn = FLT_MAX;
        for (int r = 0; r < n; r++)
            for (int u = 0; u < r; u++)
                if (s.m[(size_t)r * n + u] < mn)
                    mn = s.m[(size_t)r * n + u];
        if (mn < FLT_MAX) {
            horizon = (double)mn * horizon_factor;
            if (init_matrix)
                REFINE_BAND(-1.0, horizon);
        }
    }
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
    for (int i = 0; i < n; i++)
        rescan2(&s, i);

    int n_live = n, t = 0;
    while (n_live > n_target) {
        /* heads of all live rows, in scan order */
        int nh = 0;
        float t_d = FLT_MAX; /* stopper: (t_d, t_k, t_f), t_f = 1 for a second entry ("after the head of the same row") */
        int t_k = 0x7fffffff, t_f = 1;
        for (int r = 0; r < n; r++) {
            if (s.key[r] < 0 || s.tp[r].k1 < 0)
                continue;
            heads[nh].d = s.tp[r].d1;
            heads[nh].khi = s.key[r];
            heads[nh].klo = s.tp[r].k1;
            heads[nh].a = r;
            heads[nh].b = s.slot_of_key[s.tp[r].k1];
            nh++;
            if (s.tp[r].k2 >= 0) {
                float v = s.tp[r].d2;
                if (v < t_d || (v == t_d && s.key[r] < t_k) || (v == t_d && s.key[r] == t_k && 1 < t_f)) {
                    t_d = v;
                    t_k = s.key[r];
                    t_f = 1;
                }
            }
        }
        if (nh == 0) {
            st->exhausted = 1;
            break;
        }
        qsort(heads, (size_t)nh, sizeof(headrec), head_cmp);
        const double safe = horizon / (1.0 + 2.0 * eps_filter);
        if ((double)heads[0].d > safe) { /* the minimum reached the horizon: raise it, re-evaluate the band */
            double nh_ = fmax(horizon, (double)heads[0].d) * horizon_factor;
            REFINE_BAND(horizon, nh_);
            horizon = nh_;
            ds->n_raises++;
#pragma omp parallel for schedule(dynamic, 64) num_threads(n_threads)
            for (int i = 0; i < n; i++)
                if (s.key[i] >= 0)
                    rescan2(&s, i);
            continue;
        }
        /* any head that is not the first head at both of its clusters is a stopper */
        for (int i = 0; i < n; i++)
            first_touch[i] = -1;
        for (int i = 0; i < nh; i++) {
            const headrec *h = &heads[i];
            if (!(h->d < t_d || (h->d == t_d && h->khi < t_k) || (h->d == t_d && h->khi == t_k && 0 < t_f)))
                break; /* at or above the stopper already */
            if (first_touch[h->a] >= 0 || first_touch[h->b] >= 0) {
                t_d = h->d;
                t_k = h->khi;
                t_f = 0;
                break;
            }
            first_touch[h->a] = i;
            first_touch[h->b] = i;
        }
        int m = 0;
        int limit = n_live - (int)n_target;
        if (limit > max_batch)
            limit = max_batch;
        const double cut = (double)t_d * (1.0 - delta_cut);
        for (int i = 0; i < nh && m < limit; i++) {
            const headrec *h = &heads[i];
            if (!(h->d < t_d || (h->d == t_d && h->khi < t_k) || (h->d == t_d && h->khi == t_k && 0 < t_f)))
                break;
            if (i > 0 && !((double)h->d < cut)) {
                ds->n_cut++;
                break;
            }
            if ((double)h->d > safe)
                break;
            m++;
        }
        ds->n_iterations++;
        /* the merges of the batch, one after the other (their inputs are disjoint; cross terms come out of the
         * sequential order by themselves) */
        float last_d = heads[m - 1].d;
        int last_k = heads[m - 1].khi;
        for (int i = 0; i < m; i++) {
            const headrec h = heads[i];
            int a = h.a, b = h.b, sa = s.size[a], sb = s.size[b], snew = sa + sb;
            if (tr) {
                tr->key_hi[t] = h.khi;
                tr->key_lo[t] = h.klo;
                tr->pos_i[t] = -1;
                tr->pos_j[t] = -1;
                tr->dist[t] = h.d;
                tr->size[t] = snew;
                if (tr->gap)
                    tr->gap[t] = 0.0f;
            }
            child_hi[t] = h.khi;
            child_lo[t] = h.klo;
            oracle_merge_centroid(s.cents + (size_t)a * d, sa, s.cents + (size_t)b * d, sb, d, cnew);
            memcpy(s.cents + (size_t)b * d, cnew, sizeof(float) * (size_t)d);
            const float *ra = s.m + (size_t)a * n;
            float *rb = s.m + (size_t)b * n;
            int ne = 0;
            for (int k = 0; k < n; k++) {
                if (s.key[k] < 0 || k == a || k == b)
                    continue;
                int sk = s.size[k];
                float v = sk + snew > max_size ? INFINITY : lw_double(sa, sb, sk, ra[k], rb[k], h.d);
                rb[k] = v;
                s.m[(size_t)k * n + b] = v;
                if ((double)v <= horizon)
                    evq[ne++] = k;
            }
            double err = 0.0;
#pragma omp parallel for schedule(static) num_threads(n_threads) reduction(max : err)
            for (int q = 0; q < ne; q++) {
                int k = evq[q];
                float w = ward_w(s.size[k], snew, dsq_seq(s.cents + (size_t)k * d, cnew, d));
                double e = fabs((double)w - (double)rb[k]) / fmax((double)w, 1e-30);
                if (e > err)
                    err = e;
                rb[k] = w;
                s.m[(size_t)k * n + b] = w;
            }
            ds->n_exact += ne;
            if (err > ds->max_filter_err)
                ds->max_filter_err = err;
            /* did a pair created by this merge come out below a later member of the batch?  (what delta_cut is for) */
            for (int q = 0; q < ne && i + 1 < m; q++) {
                float w = rb[evq[q]];
                if (w < last_d || (w == last_d && n + t < last_k))
                    ds->n_violations++;
            }
            s.key[a] = -1;
            s.key[b] = n + t;
            s.size[b] = snew;
            s.slot_of_key[n + t] = b;
            for (int k = 0; k < n; k++)
                s.m[(size_t)a * n + k] = s.m[(size_t)k * n + a] = INFINITY;
            n_live--;
            t++;
        }
        /* partner lists: the new rows, and the rows that lost one of their two cached partners */
#pragma omp parallel for schedule(dynamic, 16) num_threads(n_threads)
        for (int r = 0; r < n; r++) {
            if (s.key[r] < 0)
                continue;
            int need = s.key[r] >= n + t - m;
            if (!need && s.tp[r].k1 >= 0) {
                int s1 = s.slot_of_key[s.tp[r].k1];
                if (s.key[s1] != s.tp[r].k1)
                    need = 1;
            }
            if (!need && s.tp[r].k2 >= 0) {
                int s2 = s.slot_of_key[s.tp[r].k2];
                if (s.key[s2] != s.tp[r].k2)
                    need = 1;
            }
            if (need)
                rescan2(&s, r);
        }
    }
    st->n_merges = t;
    st->n_final = n_live;
    ds->horizon = horizon;

    /* output assembly, clustering.go:265-280 (as ward_fast.c) */
    int cid = 0, pos = 0;
    if (offsets)
        offsets[0] = 0;
    int *stack = (int *)malloc(sizeof(int) * (nn + 1));
    for (int k = 0; k < n + t; k++) {
        int slot = s.slot_of_key[k];
        if (s.key[slot] != k)
            continue;
        if (s.size[slot] > max_size)
            rc = ORACLE_ERR_INTERNAL;
        if (s.size[slot] < min_size)
            continue;
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
    free(s.slot_of_key);
    free(s.cents);
    free(s.tp);
    free(child_hi);
    free(child_lo);
    free(heads);
    free(first_touch);
    free(evq);
    free(cnew);
    return rc;
}
