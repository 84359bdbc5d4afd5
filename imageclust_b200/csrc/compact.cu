// compact.cu -- K4: active-cluster compaction.
//
// The reference physically deletes two rows and two columns per merge and appends the new cluster last
// (RemoveClusters clustering.go:51-58, RemoveRowsAndColumns :100-116, append :90-93,241): its matrix is always dense and
// in slice order.  The device loop retires slots instead (no per-merge data movement), so between compactions the live
// clusters thin out: every row scan and every Lance-Williams row streams retired columns, and distances of clusters
// created since the matrix was laid out are gathered one sector at a time.  Whenever the live count has fallen to half
// of the slot count the host stops the loop and runs this file:
//   * the live clusters are renumbered densely IN KEY ORDER (== the reference's slice order at that moment);
//   * the matrix moves out of place into a buffer of a quarter of the size, lower triangle from the rows of the
//     higher-key clusters (where a pair lives), then mirrored, so every row holds ALL its partners contiguously;
//   * slot tables, partner lists and the centroid row index follow the renumbering.  Keys never change: the merge trace,
//     tie-breaks and output order are untouched.
// HBM-bound: reads ~n_old * n_live * 2 bytes (every other column of the old rows is live), writes 4 n_live^2.
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {

__global__ void __launch_bounds__(256) compact_scatter_kernel(const SlotKS* __restrict__ ks, int32_t n_old,
                                                              int32_t* __restrict__ keymap) {
    const int32_t s = static_cast<int32_t>(blockIdx.x) * 256 + threadIdx.x;
    if (s >= n_old) return;
    const int32_t key = ks[s].x;
    if (key >= 0) keymap[key] = s;
}

// one block: rank of every present key (ascending) = new slot
__global__ void __launch_bounds__(1024) compact_rank_kernel(const int32_t* __restrict__ keymap, int32_t key_cap,
                                                            int32_t* __restrict__ newslot, int32_t* __restrict__ oldslot,
                                                            int32_t n_new4, int32_t* __restrict__ n_live_out) {
    __shared__ int32_t s_cnt[1024];
    const int tid = threadIdx.x;
    const int32_t per = (key_cap + 1023) / 1024;
    const int32_t k0 = min(key_cap, tid * per), k1 = min(key_cap, k0 + per);
    int32_t c = 0;
    for (int32_t k = k0; k < k1; ++k) c += keymap[k] >= 0 ? 1 : 0;
    s_cnt[tid] = c;
    __syncthreads();
    for (int o = 1; o < 1024; o <<= 1) {  // inclusive scan
        const int32_t v = tid >= o ? s_cnt[tid - o] : 0;
        __syncthreads();
        s_cnt[tid] += v;
        __syncthreads();
    }
    int32_t rank = s_cnt[tid] - c;
    for (int32_t k = k0; k < k1; ++k) {
        const int32_t s = keymap[k];
        if (s >= 0) {
            newslot[s] = rank;
            oldslot[rank] = s;
            ++rank;
        }
    }
    const int32_t total = s_cnt[1023];
    for (int32_t r = total + tid; r < n_new4; r += 1024) oldslot[r] = -1;
    if (tid == 0) *n_live_out = total;
}

__global__ void __launch_bounds__(256) compact_state_kernel(const CompactArgs a) {
    const int32_t s = static_cast<int32_t>(blockIdx.x) * 256 + threadIdx.x;
    if (s >= a.n_new4) return;
    const int32_t o = s < a.n_new ? a.oldslot[s] : -1;
    if (o < 0) {
        a.ks_new[s] = make_int2(-1, 0);
        a.gkey_new[s] = -1;
        if (s < a.n_new) a.nn_more_new[s] = 0;  // (cannot happen: fewer live keys than the host counted)
        return;
    }
    const int2 k = a.ks_old[o];
    a.ks_new[s] = k;
    a.gkey_new[s] = k.x;
    if (a.nn_new == nullptr) return;  // (a further replica of the slot table: the lists were permuted with the first)
    if (a.slot_of_key != nullptr) a.slot_of_key[k.x] = s;
    if (a.near_meta_new != nullptr) a.near_meta_new[s] = a.near_meta_old[o];  // near lists carry keys: the pool stays
    if (a.my_rank >= 0) {  // real shards: only the owner of a row holds its list
        if (s / a.rows_per_rank_new != a.my_rank) return;
        if (o / a.rows_per_rank_old != a.my_rank) {  // the row moves here: its list is rebuilt by a scan before it is used
            if (a.near_meta_new != nullptr) a.near_meta_new[s] = make_int2(0, -1);  // (its near list stayed with the old owner)
            a.nn_new[static_cast<int64_t>(s) * kNNK] = make_uint4(kNoPartner, 0u, kNoPartner, 0u);
#pragma unroll
            for (int e = 1; e < kNNK; ++e) a.nn_new[static_cast<int64_t>(s) * kNNK + e] = make_uint4(kNoPartner, kNoPartner, kNoPartner, 0u);
            a.nn_more_new[s] = 3;
            return;
        }
    }
    int32_t more = a.nn_more_old[o];
#pragma unroll
    for (int e = 0; e < kNNK; ++e) {
        uint4 v = a.nn_old[static_cast<int64_t>(o) * kNNK + e];
        if (v.z != kNoPartner) {
            const int32_t p = a.newslot[v.z];
            if (p < 0) {  // a listed partner that is not alive: rebuild the list before it is used
                v = make_uint4(kNoPartner, v.y, kNoPartner, 0u);
                more |= 3;
            } else {
                v.z = static_cast<uint32_t>(p);
            }
        }
        a.nn_new[static_cast<int64_t>(s) * kNNK + e] = v;
    }
    a.nn_more_new[s] = more;
}

// new row s' <- the live lower-key partners of its cluster, from the cluster's old row (a pair lives in the row of its
// higher-key cluster; lower key == lower new slot).  One block per new row.
//   local source row:  gather the wanted columns (oldslot[u] for u < s'): every other sector of the old row is touched once;
//   remote source row (another rank's memory, NVLink): 4-byte gathers are request-bound there, so the old row is STREAMED with
//   coalesced 16-byte loads -- the columns before the cluster's own old slot if it is older than the last compaction (its
//   partners all sit there), the whole row otherwise -- and the live lower-key columns are scattered into the (local) new row.
__global__ void __launch_bounds__(256) compact_rows_kernel(const CompactArgs a) {
    const int32_t s = a.row0 + static_cast<int32_t>(blockIdx.x);
    if (s >= a.row1) return;
    const int32_t o = a.oldslot[s];
    const int32_t q = o / a.rows_per_rank_old;
    const float* src = a.dm_old[q] + static_cast<int64_t>(o - q * a.rows_per_rank_old) * a.ld_old;
    float* dst = a.dm_new + static_cast<int64_t>(s - a.row_base_new) * a.ld_new;
    // (the mirror pass that follows writes every entry above the diagonal; only the padding columns need a value here)
    for (int32_t u = a.n_new + threadIdx.x; u < a.n_new4; u += 256) dst[u] = INFINITY;
    if (threadIdx.x == 0) dst[s] = 0.0f;
    if (a.my_rank < 0 || q == a.my_rank) {
        for (int32_t u = threadIdx.x; u < s; u += 256) dst[u] = __ldg(src + a.oldslot[u]);
        return;
    }
    const int32_t key_o = a.ks_old[o].x;
    const int32_t n_old4 = (a.n_old + 3) & ~3;
    const int32_t end = key_o < a.order_key_old ? min(n_old4, (o + 3) & ~3) : n_old4;
    for (int32_t c0 = threadIdx.x * 4; c0 < end; c0 += 256 * 4) {
        const float4 v = __ldcg(reinterpret_cast<const float4*>(src + c0));
        const int4 ns = *reinterpret_cast<const int4*>(a.newslot + c0);  // (newslot is padded to a multiple of 4 with -1)
        if (ns.x >= 0 && ns.x < s) dst[ns.x] = v.x;
        if (ns.y >= 0 && ns.y < s) dst[ns.y] = v.y;
        if (ns.z >= 0 && ns.z < s) dst[ns.z] = v.z;
        if (ns.w >= 0 && ns.w < s) dst[ns.w] = v.w;
    }
}

// upper triangle <- lower triangle, 64 x 64 tiles through shared memory (16-byte accesses on both sides); one block per
// tile of the lower triangle, tiles enumerated linearly (no empty blocks)
constexpr int kMT = 64;
__global__ void __launch_bounds__(256) mirror_lower_kernel(const CompactArgs a, int32_t nb) {
    __shared__ float tile[kMT][kMT + 1];
    // linear index -> (bi, bj), bj <= bi
    const int64_t lin = blockIdx.x;
    int32_t bi = static_cast<int32_t>((sqrt(8.0 * static_cast<double>(lin) + 1.0) - 1.0) * 0.5);
    while (static_cast<int64_t>(bi) * (bi + 1) / 2 > lin) --bi;
    while (static_cast<int64_t>(bi + 1) * (bi + 2) / 2 <= lin) ++bi;
    const int32_t bj = static_cast<int32_t>(lin - static_cast<int64_t>(bi) * (bi + 1) / 2);
    if (bi >= nb) return;
    // source rows i in tile bi, columns j in tile bj; destination rows j (must be resident here), columns i
    const int32_t j_lo = bj * kMT, j_hi = j_lo + kMT;
    if (j_hi <= a.row0 || j_lo >= a.row1) return;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;  // 16 float4 across, 16 rows per pass
    uint32_t my_min = 0xFFFFFFFFu;
    __shared__ uint32_t s_wmin[8];
#pragma unroll
    for (int r = ty; r < kMT; r += 16) {
        const int32_t i = bi * kMT + r, j0 = bj * kMT + tx * 4;
        float4 v = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
        if (i < a.n_new && j0 < a.n_new) {  // ld is a multiple of 32: the 16 bytes are inside the row
            const int32_t q = i / a.rows_per_rank_new;
            v = __ldcg(reinterpret_cast<const float4*>(a.dm_new_rank[q] + static_cast<int64_t>(i - q * a.rows_per_rank_new) * a.ld_new + j0));
        }
        tile[r][tx * 4 + 0] = v.x;
        tile[r][tx * 4 + 1] = v.y;
        tile[r][tx * 4 + 2] = v.z;
        tile[r][tx * 4 + 3] = v.w;
        if (a.gmin != nullptr) {  // smallest selectable value of the lower triangle (what the first sweep's minimum head would be)
            const float vs[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (j0 + e < i && i < a.n_new && __float_as_uint(vs[e]) < kMaxFloatBits) my_min = min(my_min, __float_as_uint(vs[e]));
        }
    }
    if (a.gmin != nullptr) {  // (one word for 1.2 M blocks: the atomic only where the block beats the running minimum)
        my_min = __reduce_min_sync(0xffffffffu, my_min);
        if ((threadIdx.x & 31) == 0) s_wmin[threadIdx.x >> 5] = my_min;
    }
    __syncthreads();
    if (a.gmin != nullptr && threadIdx.x == 0) {
        uint32_t bmin = s_wmin[0];
#pragma unroll
        for (int w = 1; w < 8; ++w) bmin = min(bmin, s_wmin[w]);
        if (bmin < __ldcg(a.gmin)) atomicMin(a.gmin, bmin);
    }
#pragma unroll
    for (int r = ty; r < kMT; r += 16) {
        const int32_t j = bj * kMT + r, i0 = bi * kMT + tx * 4;  // write dm[j][i0..i0+3] = dm[i0..i0+3][j]
        if (j >= a.n_new || j < a.row0 || j >= a.row1 || i0 >= a.n_new) continue;
        float* dst = a.dm_new + static_cast<int64_t>(j - a.row_base_new) * a.ld_new + i0;
        const float v[4] = {tile[tx * 4 + 0][r], tile[tx * 4 + 1][r], tile[tx * 4 + 2][r], tile[tx * 4 + 3][r]};
        if (i0 > j && i0 + 3 < a.n_new) {
            *reinterpret_cast<float4*>(dst) = make_float4(v[0], v[1], v[2], v[3]);
        } else {
#pragma unroll
            for (int e = 0; e < 4; ++e)
                if (i0 + e > j && i0 + e < a.n_new) dst[e] = v[e];
        }
    }
}


// One pass for an unsharded matrix (one process: contiguous old and new matrices): 64 x 64 tiles of the NEW matrix.  The
// value of the pair (new slots i > j) is the old entry [oldslot[i]][oldslot[j]] (a pair lives in the row of its higher-key
// cluster, and the new numbering is the key order); the block gathers the tile (a row's 64 columns are ~90 consecutive old
// columns for clusters older than the last compaction: sector-efficient; tiles are enumerated row-tile major, so the blocks
// in flight read neighbouring pieces of the same 64 old rows) and writes it twice -- as is and transposed through shared
// memory, coalesced 256-byte rows both ways.  Bytes: the live lower triangle read once (in sectors), the whole new matrix
// written once; compact_rows + mirror_lower read the old rows in full, write the lower triangle, read it again and write
// the upper one (config C, nine compactions: 195 GB -> ~70 GB).
constexpr int kCT = 64;   // tile edge (128 x 128 tiles measured slower at config C: 37 ms for the nine compactions against 19 ms)
constexpr int kCTRows = 256 / kCT;  // rows per pass of the 256 threads
__global__ void __launch_bounds__(256) compact_tiles_kernel(const CompactArgs a, int32_t nb) {
    extern __shared__ float ct_smem[];
    float (*const tile)[kCT + 1] = reinterpret_cast<float (*)[kCT + 1]>(ct_smem);
    int32_t* const s_oi = reinterpret_cast<int32_t*>(ct_smem + kCT * (kCT + 1));
    int32_t* const s_oj = s_oi + kCT;
    const int64_t lin = blockIdx.x;
    int32_t bi = static_cast<int32_t>((sqrt(8.0 * static_cast<double>(lin) + 1.0) - 1.0) * 0.5);
    while (static_cast<int64_t>(bi) * (bi + 1) / 2 > lin) --bi;
    while (static_cast<int64_t>(bi + 1) * (bi + 2) / 2 <= lin) ++bi;
    const int32_t bj = static_cast<int32_t>(lin - static_cast<int64_t>(bi) * (bi + 1) / 2);
    if (bi >= nb) return;
    const int tid = threadIdx.x;
    if (tid < kCT) {
        const int32_t i = bi * kCT + tid;
        s_oi[tid] = i < a.n_new ? a.oldslot[i] : -1;
    } else if (tid < 2 * kCT) {
        const int32_t j = bj * kCT + tid - kCT;
        s_oj[tid - kCT] = j < a.n_new ? a.oldslot[j] : -1;
    }
    __syncthreads();
    const float* const old = a.dm_old[0];
    const int tx = tid % kCT, ty = tid / kCT;  // kCT columns across, kCTRows rows per pass
    const int32_t oj = s_oj[tx], j = bj * kCT + tx;
#pragma unroll 16  // (sixteen independent 4-byte loads in flight per thread)
    for (int r = ty; r < kCT; r += kCTRows) {
        const int32_t i = bi * kCT + r, oi = s_oi[r];
        // padding (rows / columns beyond the live count) holds +inf, the diagonal 0 (what compact_rows + mirror_lower leave)
        float v = INFINITY;
        if (oi >= 0 && oj >= 0) {
            if (i == j)
                v = 0.0f;
            else if (i > j)
                v = __ldcg(old + static_cast<int64_t>(oi) * a.ld_old + oj);
            else  // (diagonal tiles only: the pair's value is in the row of the higher new slot)
                v = __ldcg(old + static_cast<int64_t>(oj) * a.ld_old + oi);
        }
        tile[r][tx] = v;
    }
    __syncthreads();
    // tile (bi, bj) as is: rows i, columns j (every column up to the row stride gets a value: ld_new is a multiple of 32 and
    // the last column block covers [n_new, nb * 64) -- clipped to the stride)
    for (int r = ty; r < kCT; r += kCTRows) {
        const int32_t i = bi * kCT + r;
        if (i < a.n_new4 && j < a.ld_new) a.dm_new[static_cast<int64_t>(i) * a.ld_new + j] = tile[r][tx];
    }
    if (bi == bj) return;
    // transposed: rows j of tile bj, columns i of tile bi
    const int32_t i_t = bi * kCT + tx;
    for (int r = ty; r < kCT; r += kCTRows) {
        const int32_t jr = bj * kCT + r;
        if (jr < a.n_new4 && i_t < a.ld_new) a.dm_new[static_cast<int64_t>(jr) * a.ld_new + i_t] = tile[tx][r];
    }
}

}  // namespace

cudaError_t launch_compact_map(const SlotKS* ks, int32_t n_old, int32_t* keymap, int32_t key_cap, int32_t* newslot,
                               int32_t* oldslot, int32_t n_new4, int32_t* n_live_out, cudaStream_t s) {
    cudaError_t e = cudaMemsetAsync(keymap, 0xFF, sizeof(int32_t) * static_cast<size_t>(key_cap), s);
    if (e != cudaSuccess) return e;
    e = cudaMemsetAsync(newslot, 0xFF, sizeof(int32_t) * static_cast<size_t>((n_old + 3) / 4 * 4), s);
    if (e != cudaSuccess) return e;
    compact_scatter_kernel<<<static_cast<unsigned>((n_old + 255) / 256), 256, 0, s>>>(ks, n_old, keymap);
    compact_rank_kernel<<<1, 1024, 0, s>>>(keymap, key_cap, newslot, oldslot, n_new4, n_live_out);
    return cudaGetLastError();
}

cudaError_t launch_compact_state(const CompactArgs& a, cudaStream_t s) {
    compact_state_kernel<<<static_cast<unsigned>((a.n_new4 + 255) / 256), 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_compact_rows(const CompactArgs& a, cudaStream_t s) {
    if (a.row1 <= a.row0) return cudaSuccess;
    compact_rows_kernel<<<static_cast<unsigned>(a.row1 - a.row0), 256, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_compact_tiles(const CompactArgs& a, cudaStream_t s) {
    if (a.n_new <= 0) return cudaSuccess;
    const int64_t nb = (a.n_new4 + kCT - 1) / kCT;
    const int64_t tiles = nb * (nb + 1) / 2;
    const size_t smem = sizeof(float) * kCT * (kCT + 1) + sizeof(int32_t) * 2 * kCT;
    cudaError_t e = cudaFuncSetAttribute(compact_tiles_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    compact_tiles_kernel<<<static_cast<unsigned>(tiles), 256, smem, s>>>(a, static_cast<int32_t>(nb));
    return cudaGetLastError();
}

cudaError_t launch_mirror_lower(const CompactArgs& a, cudaStream_t s) {
    if (a.n_new <= 1 || a.row1 <= a.row0) return cudaSuccess;
    const int64_t nb = (a.n_new + kMT - 1) / kMT;
    const int64_t tiles = nb * (nb + 1) / 2;
    mirror_lower_kernel<<<static_cast<unsigned>(tiles), 256, 0, s>>>(a, static_cast<int32_t>(nb));
    return cudaGetLastError();
}

}  // namespace ic
