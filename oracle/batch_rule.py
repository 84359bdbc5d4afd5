"""The BATCH RULE of imageclust_b200/csrc/merge_batch.cu, restated in numpy -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The device loop takes several merges per iteration.  This module restates WHICH merges one iteration may take and HOW
the distances of a batch are updated (Lance-Williams rows + the chained cross terms), on a dense matrix, so that the
CPU test suite can check the claim the kernel rests on: the batched procedure produces the merge sequence of the
sequential algorithm (oracle ``fast_cluster`` in Lance-Williams mode, itself pinned against the literal restatement of
/root/reference/internal/clustering/clustering.go:198-284), bit for bit.

Rule: every live row r has a head -- its smallest pair (d, key_r, key_partner) over partners with a
LOWER key -- and a second entry.  Walk the heads in the reference's scan order (clustering.go:119-133: smallest
(d, key_hi, key_lo)); stop at the first pair that touches a cluster of an earlier pair.  That stopper is
T = min(second entry of any row, any head that is not the first head at both of its clusters); all heads below T are
merged, in order.
"""
from __future__ import annotations

import numpy as np

F32 = np.float32
MAXF = np.finfo(np.float32).max


def _lw(sa, sb, sk, dka, dkb, dab):
    """oracle/ward_fast.c LW mode == loop_common.cuh lance_williams(): double arithmetic, product with the correctly
    rounded reciprocal of the size sum, one rounding to fp32, canonical non-negative result."""
    num = (np.float64(sa + sk) * np.float64(dka) + np.float64(sb + sk) * np.float64(dkb)) - np.float64(sk) * np.float64(dab)
    with np.errstate(invalid="ignore", over="ignore"):
        v = F32(num * (np.float64(1.0) / np.float64(sa + sb + sk)))
    if not (v >= 0):
        v = F32(np.inf) if v != v else F32(0.0)
    return F32(v + F32(0.0))


def batched_cluster(init_matrix, n_target: int, max_size: int, max_batch: int = 512):
    """Run the batched procedure on a symmetric fp32 matrix of singleton distances.

    Returns (key_hi, key_lo, dist, size, batch_sizes): the merge trace in the format of the oracle and the number of
    merges each iteration took."""
    m = np.array(init_matrix, dtype=F32, copy=True)
    n = m.shape[0]
    key = np.arange(n, dtype=np.int64)      # monotone order ids: item index, then n + t for merge t
    size = np.ones(n, dtype=np.int64)
    alive = np.ones(n, dtype=bool)
    if max_size < 2:
        m[:] = np.inf                        # 1 + 1 > maxSize: every pair is inadmissible (clustering.go:228)
    tr_hi, tr_lo, tr_d, tr_s, batches = [], [], [], [], []
    n_live, t = n, 0
    while n_live > n_target:
        idx = np.flatnonzero(alive)
        # heads and second entries over partners with a lower key, selectable distances only (< MaxFloat32, :120-124)
        heads, stoppers = [], []
        for r in idx:
            part = idx[key[idx] < key[r]]
            d = m[r, part]
            ok = d < MAXF
            part, d = part[ok], d[ok]
            if len(part) == 0:
                continue
            order = np.lexsort((key[part], d))          # (d, partner key)
            p0 = part[order[0]]
            heads.append((F32(d[order[0]]), int(key[r]), int(key[p0]), int(r), int(p0)))
            if len(order) > 1:
                stoppers.append((F32(d[order[1]]), int(key[r]), 1))   # (d, key_hi) + "after the head of the same row"
        if not heads:
            break                                        # exhausted: no admissible pair (:222-225)
        heads.sort(key=lambda h: (h[0], h[1]))
        first_touch = {}
        for h in heads:                                  # any head that is not the first head at both of its clusters
            for slot in (h[3], h[4]):
                if slot in first_touch:
                    stoppers.append((h[0], h[1], 0))
                else:
                    first_touch[slot] = h
        T = min(stoppers) if stoppers else (F32(np.inf), 0, 0)
        batch = [h for h in heads if (h[0], h[1], 0) < T][: max(1, min(max_batch, n_live - n_target))]
        assert batch, "the global minimum head is always below the stopper"
        batches.append(len(batch))
        merged = set()
        for h in batch:
            merged.update((h[3], h[4]))
        others = np.array([k for k in idx if k not in merged], dtype=np.int64)
        old = m.copy()                                   # every input of a batch is a value from BEFORE the batch
        for i, (d, khi, klo, a, b) in enumerate(batch):
            sa, sb = int(size[a]), int(size[b])
            for k in others:                             # Lance-Williams rows, eager admissibility (:228)
                sk = int(size[k])
                v = F32(np.inf) if sk + sa + sb > max_size else _lw(sa, sb, sk, old[k, a], old[k, b], d)
                m[k, b] = m[b, k] = v
            for j in range(i):                           # cross terms with the earlier merges of the batch
                dj, _, _, aj, bj = batch[j]
                saj, sbj = int(size[aj]), int(size[bj])
                sj = saj + sbj
                # merge j seen from k = a and k = b (values before the batch) ...
                t1 = F32(np.inf) if sa + sj > max_size else _lw(saj, sbj, sa, old[a, aj], old[a, bj], dj)
                t2 = F32(np.inf) if sb + sj > max_size else _lw(saj, sbj, sb, old[b, aj], old[b, bj], dj)
                # ... then this merge seen from k = new_j
                v = F32(np.inf) if sj + sa + sb > max_size else _lw(sa, sb, sj, t1, t2, d)
                m[b, bj] = m[bj, b] = v
        for i, (d, khi, klo, a, b) in enumerate(batch):
            tr_hi.append(khi)
            tr_lo.append(klo)
            tr_d.append(d)
            tr_s.append(int(size[a] + size[b]))
            size[b] = size[a] + size[b]
            key[b] = n + t                               # appended last (:241): the highest key so far
            alive[a] = False
            m[a, :] = np.inf
            m[:, a] = np.inf
            t += 1
            n_live -= 1
    return (np.array(tr_hi, np.int32), np.array(tr_lo, np.int32), np.array(tr_d, F32), np.array(tr_s, np.int32),
            np.array(batches, np.int32))
