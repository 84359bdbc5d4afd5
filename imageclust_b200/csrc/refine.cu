// refine.cu -- re-evaluation of a band of stored distances with the reference's own arithmetic.
//
// The reference's matrix entry for two clusters is always WardDistance(centroid, centroid) in sequential fp32
// (clustering.go:83-86,136-157).  The device keeps tensor-core Gram values (K1) and Lance-Williams values (K3b), which
// agree with it to a few 1e-6 -- not enough to order the near-ties of 1e5 merges the way the reference does.  The merge
// loop therefore only decides among values at or below a HORIZON, and every stored value at or below the horizon is the
// reference's own.  This file establishes that invariant for a band (lo, hi] of stored values: when the horizon is first
// set (initial matrix), when it is raised, and when the loop's own queue overflowed.
//   refine_collect_kernel   HBM-bound sweep of the resident rows: pairs (r, u), key_u < key_r, both live, lo < value <= hi
//   refine_eval_kernel      one warp per collected pair: exact.cuh
// Re-evaluating a pair that already holds the reference's value is harmless (the value is a pure function of the two
// centroids and sizes), so bands may overlap.
#include <algorithm>

#include "common.cuh"
#include "exact.cuh"
#include "kernels.h"

namespace ic {

namespace {
constexpr int kColT = 256;

__global__ void __launch_bounds__(256) init_centroids_kernel(const float* __restrict__ x, int64_t n, int64_t d, int64_t ldx,
                                                             float* __restrict__ cen, int64_t ldc) {
    const int64_t total = n * ldc;
    for (int64_t i = static_cast<int64_t>(blockIdx.x) * blockDim.x + threadIdx.x; i < total;
         i += static_cast<int64_t>(gridDim.x) * blockDim.x) {
        const int64_t r = i / ldc, c = i - r * ldc;
        cen[i] = c < d ? x[r * ldx + c] : 0.0f;
    }
}

// one block per row; 16-byte loads of the row and of the keys.  The columns of the row's pairs in the band are stashed in
// shared memory, then ONE contiguous range of the queue is reserved for the row and filled from the stash -- consecutive
// queue entries share their row, so the evaluation kernel runs eight summation chains per warp and reads the row's centroid
// once per eight pairs.  A row with more than kStash pairs in the band is swept a second time (rare; the second sweep of
// every row, which this replaces, did not stay in L2: 57 GB of DRAM reads for the three sweeps of config C).
constexpr int kStash = 2048;
__global__ void __launch_bounds__(kColT) refine_collect_kernel(const __grid_constant__ RefineArgs a) {
    const int32_t r = a.row0 + static_cast<int32_t>(blockIdx.x);
    if (r >= a.row1) return;
    const int32_t key_r = a.gkey[r];
    if (key_r < 0 || key_r < a.min_row_key) return;
    const float* row = a.dm + static_cast<int64_t>(r - a.r_lo) * a.ld;
    const int32_t n4 = (a.n_slots + 3) & ~3;
    const int32_t u_end = a.lower_only ? min(n4, (r + 3) & ~3) : n4;
    const int lane = threadIdx.x & 31;
    const float lo = static_cast<float>(a.lo), hi = static_cast<float>(a.hi);
    const bool lo_open = a.lo < 0.0;  // band starts below every value
    __shared__ int32_t s_cnt, s_base, s_total;
    __shared__ int32_t s_stash[kStash];
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    const int32_t u_pad = ((u_end + kColT * 4 - 1) / (kColT * 4)) * (kColT * 4);
    for (int pass = 0; pass < 2; ++pass) {
        for (int32_t u0 = threadIdx.x * 4; u0 < u_pad; u0 += kColT * 4) {
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
                const bool hit = ks4[e] >= 0 && ks4[e] < key_r && (lo_open || vs[e] > lo) && vs[e] <= hi;
                const uint32_t mask = __ballot_sync(0xffffffffu, hit);
                if (mask == 0u) continue;
                int32_t pos = 0;
                const int leader = __ffs(mask) - 1;
                if (lane == leader) pos = atomicAdd(&s_cnt, __popc(mask));
                pos = __shfl_sync(0xffffffffu, pos, leader) + __popc(mask & ((1u << lane) - 1u));
                if (hit) {
                    if (pass == 0) {
                        if (pos < kStash) s_stash[pos] = u0 + e;
                    } else if (s_base + pos < a.cap) {
                        a.q[s_base + pos] = make_int2(r, u0 + e);
                    }
                }
            }
        }
        if (pass == 1) return;
        __syncthreads();
        if (threadIdx.x == 0) {
            s_total = s_cnt;
            s_base = s_cnt > 0 ? atomicAdd(a.cnt, s_cnt) : 0;
            s_cnt = 0;
        }
        __syncthreads();
        if (s_total == 0) return;  // (uniform) nothing of this row lies in the band
        if (s_total <= kStash) {
            for (int32_t i = threadIdx.x; i < s_total; i += kColT)
                if (s_base + i < a.cap) a.q[s_base + i] = make_int2(r, s_stash[i]);
            return;
        }
    }
}

__global__ void __launch_bounds__(256) refine_eval_kernel(const __grid_constant__ RefineArgs a) {
    __shared__ __align__(16) float s_buf[8][kExGroupR * kExStrideR];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int32_t total = min(__ldg(a.cnt), a.cap);
    const int32_t gw = static_cast<int32_t>(blockIdx.x) * 8 + warp, GW = static_cast<int32_t>(gridDim.x) * 8;
    const int d4 = static_cast<int>(a.ldc);
    int32_t n_done = 0;
    float max_err = 0.0f;
    // the collect kernel appends a row's pairs in runs: consecutive entries mostly share their row, whose centroid (and the
    // latency of the summation chain) is then shared by up to kExGroupR evaluations
    for (int32_t base = gw * kExGroupR; base < total; base += GW * kExGroupR) {
        const int32_t cnt = min(kExGroupR, total - base);
        int2 p = make_int2(-1, -1);
        if (lane < cnt) p = a.q[base + lane];
        int32_t done = 0;
        while (done < cnt) {
            const int32_t r0 = __shfl_sync(0xffffffffu, p.x, done);
            const uint32_t same = __ballot_sync(0xffffffffu, lane >= done && lane < cnt && p.x == r0) >> done;
            const int32_t run = __ffs(~same) - 1;  // entries done .. done + run - 1 share the row
            const int32_t u = __shfl_sync(0xffffffffu, p.y, min(done + lane, 31));
            const int2 kr = a.ks[r0];
            int2 ku = make_int2(0, 0);
            const float* pb = nullptr;
            if (lane < run) {
                ku = a.ks[u];
                pb = a.cen + static_cast<int64_t>(ku.x) * a.ldc;  // centroids are stored by key
            }
            const float dsq = warp_exact_dsq_group<kExGroupR, kExChunkR>(a.cen + static_cast<int64_t>(kr.x) * a.ldc, pb, run, d4, s_buf[warp]);
            if (lane < run) {
                const float w = ward_weight(kr.y, ku.y, dsq);
                float* dst = a.dm + static_cast<int64_t>(r0 - a.r_lo) * a.ld + u;
                const float stored = *dst;
                max_err = fmaxf(max_err, exact_monitor(a.ctl, stored, w, a.eps_filter, a.abs_slack));
                if (__float_as_uint(stored) != __float_as_uint(w)) {
                    *dst = w;
                    if (kr.x < a.mirror_key) {  // both clusters are older than the last compaction: the pair is stored in both rows
                        const int32_t q = u / a.rows_per_rank;
                        a.dm_rank[q][static_cast<int64_t>(u - q * a.rows_per_rank) * a.ld + r0] = w;
                    }
                    atomicOr(a.nn_more + r0, 3);  // kMoreBit | kDryBit: the row's partner list is rebuilt before it is used
                }
            }
            if (lane == 0) n_done += run;
            done += run;
        }
    }
    if (lane == 0 && n_done > 0) atomicAdd(a.ctl + CTL_N_EXACT, n_done);
    exact_monitor_flush(a.ctl, max_err);
}
}  // namespace

cudaError_t launch_init_centroids(const float* x, int64_t n, int64_t d, int64_t ldx, float* cen, int64_t ldc, cudaStream_t s) {
    if (n <= 0 || ldc <= 0) return cudaSuccess;
    const int64_t total = n * ldc;
    const unsigned blocks = static_cast<unsigned>(std::min<int64_t>((total + 255) / 256, 148 * 16));
    init_centroids_kernel<<<blocks, 256, 0, s>>>(x, n, d, ldx, cen, ldc);
    return cudaGetLastError();
}

cudaError_t launch_refine_collect(const RefineArgs& a, cudaStream_t s) {
    if (a.row1 <= a.row0) return cudaSuccess;
    refine_collect_kernel<<<static_cast<unsigned>(a.row1 - a.row0), kColT, 0, s>>>(a);
    return cudaGetLastError();
}

cudaError_t launch_refine_eval(const RefineArgs& a, int num_sms, cudaStream_t s) {
    refine_eval_kernel<<<static_cast<unsigned>(num_sms * 8), 256, 0, s>>>(a);
    return cudaGetLastError();
}

}  // namespace ic
