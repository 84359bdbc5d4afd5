"""Independent numpy restatement of clustering.go -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

A second, separately written restatement of
``/root/reference/internal/clustering/clustering.go`` used to pin the C oracle
(``ward_literal.c``): two independent restatements that agree bit for bit on
traces, maps and matrices.  Pure Python loops -- small cases only.

PARITY UNPINNED by the reference's own tests (it has none).
"""
from __future__ import annotations

import math

import numpy as np

F32 = np.float32
MAXF32 = np.finfo(np.float32).max


def dot_float32(a, b):
    """DotFloat32, clustering.go:148-157: sequential fp32 accumulation.
    np.add.accumulate is strictly left-to-right (no pairwise blocking)."""
    if len(a) != len(b):
        raise ValueError("DotFloat32: slices have different lengths")  # :149-151 panics
    if len(a) == 0:
        return F32(0)
    prod = (a * b).astype(F32)
    acc = np.add.accumulate(np.concatenate(([F32(0)], prod)), dtype=F32)
    return F32(acc[-1])


def ward_distance(ca, sa, cb, sb):
    """WardDistance, clustering.go:136-145."""
    diff = (ca - cb).astype(F32)
    dsq = dot_float32(diff, diff)
    num = F32(sa * sb)  # integer product first (:142)
    den = F32(sa + sb)
    return F32(F32(num / den) * dsq)


def merge_centroid(ca, sa, cb, sb):
    """clustering.go:39."""
    fa, fb, fs = F32(sa), F32(sb), F32(sa + sb)
    return (((fa * ca).astype(F32) + (fb * cb).astype(F32)).astype(F32) / fs).astype(F32)


def calculate_optimal_clusters(total, min_size, max_size):
    """CalculateOptimalClusters, clustering.go:168-186 -> (n, err or None)."""
    if total < min_size:
        return 0, "too_few"
    lo = int(math.ceil(float(total) / float(max_size)))
    hi = int(math.floor(float(total) / float(min_size)))
    if lo > hi:
        return 0, "unsat"
    n = lo
    if lo < hi:
        n = (lo + hi) // 2
    return n, None


def find_closest(m):
    """FindClosestClusters, clustering.go:119-133 (vectorised per row; the first
    strict minimum in row-major order is what np.argmin returns per row)."""
    best = MAXF32
    bi = bj = -1
    for i in range(len(m)):
        if i == 0:
            continue
        row = np.asarray(m[i][:i], dtype=F32)
        with np.errstate(invalid="ignore"):
            row_cmp = np.where(np.isnan(row), np.inf, row)
        j = int(np.argmin(row_cmp))
        if row_cmp[j] < best:
            best = row_cmp[j]
            bi, bj = i, j
    return bi, bj


def perform_clustering_with_constraints(x, min_size, max_size):
    """PerformClusteringWithConstraints, clustering.go:198-284.
    Returns None on constraint error, else dict(clusters, trace, rejections, exhausted)."""
    x = np.asarray(x, dtype=F32)
    total = len(x)
    n_clusters, err = calculate_optimal_clusters(total, min_size, max_size)
    if err is not None:
        return None
    idx = [[i] for i in range(total)]
    size = [1] * total
    cent = [x[i].copy() for i in range(total)]
    key = list(range(total))
    m = [[F32(0)] * total for _ in range(total)]
    for i in range(total):
        for j in range(i):
            dist = ward_distance(cent[i], size[i], cent[j], size[j])
            m[i][j] = dist
            m[j][i] = dist
    init = np.array(m, dtype=F32).reshape(total, total)
    trace = []
    rejections = 0
    exhausted = False
    while len(idx) > n_clusters:
        i, j = find_closest(m)
        if i == -1 or j == -1:
            exhausted = True
            break
        if size[i] + size[j] > max_size:
            m[i][j] = MAXF32
            m[j][i] = MAXF32
            rejections += 1
            continue
        d_ij = m[i][j]
        new_idx = idx[i] + idx[j]
        new_size = size[i] + size[j]
        new_cent = merge_centroid(cent[i], size[i], cent[j], size[j])
        trace.append((key[i], key[j], i, j, float(d_ij), new_size))
        new_key = total + len(trace) - 1
        for lst in (idx, size, cent, key):
            del lst[i]
            del lst[j]
        for row in m:
            del row[i]
            del row[j]
        del m[i]
        del m[j]
        idx.append(new_idx)
        size.append(new_size)
        cent.append(new_cent)
        key.append(new_key)
        n = len(idx)
        new_row = [ward_distance(cent[k], size[k], new_cent, new_size) for k in range(n - 1)] + [F32(0)]
        for k in range(n - 1):
            m[k].append(new_row[k])
        m.append(new_row)
    clusters = [np.array(c, dtype=np.int32) for c, s in zip(idx, size) if s >= min_size]
    return dict(clusters=clusters, trace=trace, rejections=rejections, exhausted=exhausted,
                init_matrix=init, final_keys=list(key),
                final_matrix=np.array(m, dtype=F32).reshape(len(idx), len(idx)))
