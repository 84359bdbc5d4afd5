// near.cu -- per-row NEAR LISTS: every partner of a row whose stored value is at or below the horizon.
//
// FindClosestClusters (clustering.go:119-133) only ever picks pairs near the current minimum, and on the device every such
// pair is at or below the horizon.  A row's cached partner list (its 4 smallest lower-key partners)
// therefore never needs more than the row's pairs <= horizon -- a few hundred of the row's 10^5 columns on the benchmark
// mixtures -- and pairs between two unchanged clusters never change.  So instead of re-reading a whole row (and the keys)
// every time a partner list runs low, the loop re-selects from the row's near list: {value bits, partner key} of every
// lower-key partner <= horizon, built by one sweep of the matrix whenever the horizon is set or raised (this file), and
// from the re-evaluated pairs of a new cluster (merge_batch.cu).  Entries carry KEYS, liveness comes from slot_of_key[], so
// compactions leave the pool untouched.  A row whose partners beyond the horizon may exist carries a flag: its list then
// ends in a bound at the horizon.
// HBM-bound sweep: 4 bytes per live lower-triangle pair per build (+ the keys).
#include <algorithm>

#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {
constexpr int kNT = 256;

__global__ void __launch_bounds__(256) slot_of_key_init_kernel(int32_t* __restrict__ sok, int32_t n, int32_t cap) {
    const int32_t i = static_cast<int32_t>(blockIdx.x) * 256 + threadIdx.x;
    if (i < cap) sok[i] = i < n ? i : -1;
}

// One block per resident row, ONE sweep: the row's pairs at or below the horizon are stashed in shared memory, the block
// then reserves a contiguous segment of the pool (one atomic on the cursor) and copies the stash.  A row with more than
// kNearStash such pairs sweeps a second time into its segment (rare).  If the pool is exhausted the row gets no near list
// (meta.y = -1: the loop scans it instead) -- so do all later ones, and the loop's own appends.
// (round 2 until here: count sweep, scan kernel, fill sweep -- twice the bytes.)
constexpr int kNearStash = 2048;
__global__ void __launch_bounds__(kNT) near_sweep_kernel(const __grid_constant__ NearArgs a) {
    const int32_t r = a.r_lo + static_cast<int32_t>(blockIdx.x);
    if (r >= a.r_hi) return;
    const int32_t key_r = a.gkey[r];
    __shared__ int32_t s_cnt, s_far, s_base;
    __shared__ uint2 s_stash[kNearStash];
    if (threadIdx.x == 0) s_cnt = s_far = 0;
    __syncthreads();
    if (key_r < 0) {
        if (threadIdx.x == 0) a.meta[r] = make_int2(0, 0);
        return;
    }
    const float* row = a.dm + static_cast<int64_t>(r - a.r_lo) * a.ld;
    const int32_t n4 = (a.n_slots + 3) & ~3;
    const int32_t u_end = key_r < a.order_key ? min(n4, (r + 3) & ~3) : n4;  // older than the last compaction: key order == slot order
    const float hi = static_cast<float>(a.horizon);
    const int lane = threadIdx.x & 31;
    for (int pass = 0; pass < 2; ++pass) {
        for (int32_t u0 = threadIdx.x * 4; u0 < ((u_end + kNT * 4 - 1) / (kNT * 4)) * (kNT * 4); u0 += kNT * 4) {
            float4 v = make_float4(INFINITY, INFINITY, INFINITY, INFINITY);
            int4 k = make_int4(-1, -1, -1, -1);
            if (u0 < u_end) {
                v = pass == 0 ? ld_stream_f4(reinterpret_cast<const float4*>(row + u0)) : __ldcg(reinterpret_cast<const float4*>(row + u0));
                k = __ldg(reinterpret_cast<const int4*>(a.gkey + u0));
            }
            const float vs[4] = {v.x, v.y, v.z, v.w};
            const int32_t ks4[4] = {k.x, k.y, k.z, k.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                const bool part = ks4[e] >= 0 && ks4[e] < key_r;
                const bool hit = part && vs[e] <= hi;
                const bool far = part && !hit && __float_as_uint(vs[e]) < kMaxFloatBits;
                const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                if (pass == 0 && __any_sync(0xffffffffu, far) && lane == 0) s_far = 1;
                if (mask == 0u) continue;
                int32_t pos = 0;
                const int leader = __ffs(mask) - 1;
                if (lane == leader) pos = atomicAdd(&s_cnt, __popc(mask));
                pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(mask & ((1u << lane) - 1u));
                if (hit) {
                    const uint2 ent = make_uint2(__float_as_uint(vs[e]), static_cast<uint32_t>(ks4[e]));
                    if (pass == 0) {
                        if (pos < kNearStash) s_stash[pos] = ent;
                    } else {
                        a.pool[static_cast<int64_t>(s_base) + pos] = ent;
                    }
                }
            }
        }
        if (pass == 1) return;
        __syncthreads();
        const int32_t cnt = s_cnt;
        if (threadIdx.x == 0) {
            const int32_t base = cnt > 0 ? atomicAdd(a.cursor, cnt) : 0;
            const bool fits = static_cast<int64_t>(base) + cnt <= static_cast<int64_t>(a.pool_cap) && base >= 0;
            s_base = fits ? base : -1;
            a.meta[r] = fits ? make_int2(base, cnt | (s_far ? kNearFarBit : 0)) : make_int2(0, -1);
            s_cnt = 0;  // (positions of the second sweep, if there is one)
        }
        __syncthreads();
        if (s_base < 0 || cnt == 0) return;
        if (cnt <= kNearStash) {
            for (int32_t i = threadIdx.x; i < cnt; i += kNT) a.pool[static_cast<int64_t>(s_base) + i] = s_stash[i];
            return;
        }
    }
}

// every live resident row gets its partner list rebuilt before it is used again (after a horizon raise the bounds are stale)
__global__ void __launch_bounds__(256) mark_rows_dry_kernel(const int32_t* __restrict__ gkey, int32_t* __restrict__ nn_more,
                                                            int32_t r_lo, int32_t r_hi) {
    const int32_t r = r_lo + static_cast<int32_t>(blockIdx.x) * 256 + threadIdx.x;
    if (r < r_hi && gkey[r] >= 0) nn_more[r] |= 3;
}
// no partner lists yet: every row is rebuilt (from its near list) before it is used -- the start without a first sweep
__global__ void __launch_bounds__(256) init_lists_dry_kernel(uint4* __restrict__ nn, int32_t* __restrict__ nn_more, int32_t n) {
    const int32_t r = static_cast<int32_t>(blockIdx.x) * 256 + threadIdx.x;
    if (r >= n) return;
#pragma unroll
    for (int x = 0; x < kNNK; ++x) nn[static_cast<int64_t>(r) * kNNK + x] = make_uint4(kNoPartner, kNoPartner, kNoPartner, 0u);
    nn_more[r] = 3;  // kMoreBit | kDryBit
}
}  // namespace

cudaError_t launch_init_lists_dry(SlotNN* nn, int32_t* nn_more, int32_t n, cudaStream_t s) {
    if (n <= 0) return cudaSuccess;
    init_lists_dry_kernel<<<static_cast<unsigned>((n + 255) / 256), 256, 0, s>>>(reinterpret_cast<uint4*>(nn), nn_more, n);
    return cudaGetLastError();
}

cudaError_t launch_slot_of_key_init(int32_t* sok, int32_t n, int32_t cap, cudaStream_t s) {
    if (cap <= 0) return cudaSuccess;
    slot_of_key_init_kernel<<<static_cast<unsigned>((cap + 255) / 256), 256, 0, s>>>(sok, n, cap);
    return cudaGetLastError();
}

cudaError_t launch_near_build(const NearArgs& a, cudaStream_t s) {
    if (a.r_hi <= a.r_lo) return cudaSuccess;
    const unsigned rows = static_cast<unsigned>(a.r_hi - a.r_lo);
    cudaError_t e = cudaMemsetAsync(a.cursor, 0, sizeof(int32_t) * 2, s);  // the pool is rebuilt from scratch
    if (e != cudaSuccess) return e;
    near_sweep_kernel<<<rows, kNT, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_mark_rows_dry(const int32_t* gkey, int32_t* nn_more, int32_t r_lo, int32_t r_hi, cudaStream_t s) {
    if (r_hi <= r_lo) return cudaSuccess;
    mark_rows_dry_kernel<<<static_cast<unsigned>((r_hi - r_lo + 255) / 256), 256, 0, s>>>(gkey, nn_more, r_lo, r_hi);
    return cudaGetLastError();
}

}  // namespace ic
