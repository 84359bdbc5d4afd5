// exact.cuh -- WardDistance with the reference's OWN arithmetic for one pair of clusters, one warp per pair.
//
// The reference recomputes every distance of a new cluster from fp32 centroids (clustering.go:83-86):
//   WardDistance (:136-145)   diff[i] = a[i] - b[i];  dsq = DotFloat32(diff, diff);  (float32(na*nb) / float32(na+nb)) * dsq
//   DotFloat32   (:148-157)   sum += a[i]*b[i], i ascending, fp32, every multiply and add rounded separately (no FMA on
//                             amd64 with the default GOAMD64=v1)
//   MergeClusters (:36-40)    centroid[i] = (float32(na)*ca[i] + float32(nb)*cb[i]) / float32(na+nb)
// The merge loop keeps Lance-Williams values (cheap, off by a few 1e-6) and calls this for the pairs that can decide a
// merge.  Centroids are stored by cluster KEY (row k < N: item k; row N + t: the cluster made by
// merge t, written once by the iteration that creates it), so nothing moves when slots are reused or renumbered.
//
// The 32 lanes load both rows coalesced and form the squared differences in parallel; the sum itself is the
// reference's sequential chain (D dependent fp32 adds, ~4 cycles each: 8.4 k cycles at D = 2048) run by lane 0 from
// shared memory, while the loads of the next chunk are in flight.
#pragma once
#include "common.cuh"
#include "kernels.h"

namespace ic {

constexpr int kExChunk = 256;  // floats of every row staged per step (merge loop; refine.cu: 128)

// centroid[i] = (float32(na)*ca[i] + float32(nb)*cb[i]) / float32(na+nb), clustering.go:39
IC_DEVINL float ex_merge1(float fa, float a, float fb, float b, float fs) {
    return __fdiv_rn(__fadd_rn(__fmul_rn(fa, a), __fmul_rn(fb, b)), fs);
}

// The reference's sequential sum (clustering.go:152-155) over n4 float4 of shared memory: sum = fl(sum + p_i), i ascending.
// The adds are one dependent chain (4 cycles each); the shared-memory loads are not, so four float4 are requested a batch
// ahead -- issued just in time, every float4 added ~30 cycles of load latency to 16 cycles of adds (round 2's capture:
// the chain's FADDs waited on the short scoreboard, ~1.8 k cycles per 128-float chunk instead of ~0.55 k).
IC_DEVINL float ex_chain_sum(const float4* __restrict__ src, int n4, float sum) {
    float4 c0 = make_float4(0.f, 0.f, 0.f, 0.f), c1 = c0, c2 = c0, c3 = c0;
    if (n4 > 0) c0 = src[0];
    if (n4 > 1) c1 = src[1];
    if (n4 > 2) c2 = src[2];
    if (n4 > 3) c3 = src[3];
    int i = 0;
    for (; i + 4 <= n4; i += 4) {
        float4 d0 = c0, d1 = c1, d2 = c2, d3 = c3;
        if (i + 4 < n4) c0 = src[i + 4];
        if (i + 5 < n4) c1 = src[i + 5];
        if (i + 6 < n4) c2 = src[i + 6];
        if (i + 7 < n4) c3 = src[i + 7];
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, d0.x), d0.y), d0.z), d0.w);
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, d1.x), d1.y), d1.z), d1.w);
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, d2.x), d2.y), d2.z), d2.w);
        sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, d3.x), d3.y), d3.z), d3.w);
    }
    // tail (n4 % 4 float4, already loaded): +0 terms must NOT be added for real data, so only the valid ones
    if (i < n4) sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, c0.x), c0.y), c0.z), c0.w);
    if (i + 1 < n4) sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, c1.x), c1.y), c1.z), c1.w);
    if (i + 2 < n4) sum = __fadd_rn(__fadd_rn(__fadd_rn(__fadd_rn(sum, c2.x), c2.y), c2.z), c2.w);
    return sum;
}

// dsq = DotFloat32(diff, diff), diff = A - B over d4 floats (rows are zero padded to a multiple of 4: +0 terms leave the
// sum unchanged).  Whole warp; sbuf = this warp's kExChunk floats of shared memory.  Rows are read with ld.global.cg: the
// centroid store is written by other SMs in the phase before, L1 must not serve it.
IC_DEVINL float warp_exact_dsq(const float* __restrict__ pa, const float* __restrict__ pb, int d4, float* sbuf) {
    const int lane = threadIdx.x & 31;
    constexpr int Q = kExChunk / 128;
    float4 a[Q], b[Q];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    auto issue = [&](int c0) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int e = c0 + 4 * (lane + 32 * q);
            const bool ok = e < d4;
            a[q] = ok ? __ldcg(reinterpret_cast<const float4*>(pa + e)) : z;
            b[q] = ok ? __ldcg(reinterpret_cast<const float4*>(pb + e)) : z;
        }
    };
    float sum = 0.0f;
    issue(0);
    for (int c0 = 0; c0 < d4; c0 += kExChunk) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const float dx = __fsub_rn(a[q].x, b[q].x), dy = __fsub_rn(a[q].y, b[q].y), dz = __fsub_rn(a[q].z, b[q].z),
                        dw = __fsub_rn(a[q].w, b[q].w);
            reinterpret_cast<float4*>(sbuf)[lane + 32 * q] =
                make_float4(__fmul_rn(dx, dx), __fmul_rn(dy, dy), __fmul_rn(dz, dz), __fmul_rn(dw, dw));
        }
        __syncwarp();
        if (c0 + kExChunk < d4) issue(c0 + kExChunk);  // in flight while lane 0 runs the chain
        if (lane == 0) sum = ex_chain_sum(reinterpret_cast<const float4*>(sbuf), (min(kExChunk, d4 - c0)) >> 2, sum);
        __syncwarp();
    }
    return __shfl_sync(0xffffffffu, sum, 0);
}

// Up to G pairs that share their first cluster, one warp: the squared differences of every pair are staged side by side
// (CH floats per step) and lanes 0 .. np-1 run the np sequential chains in lockstep -- the chain latency (the floor of one
// evaluation) is paid once per group, row A is read once, and the fp32 add pipe (one warp instruction per 2 cycles and
// scheduler whatever the number of active lanes) issues 1/np of the instructions.  `pb` is lane k's second row (lanes >= np:
// ignored); returns pair k's dsq on lane k.  sbuf = G * (CH + 4) floats of this warp (the stride keeps the lanes' 16-byte
// reads on distinct banks).  The merge loop's exact phase has ~10 groups per SM and iteration: G = 4, CH = 256 (more, shorter
// groups); refine.cu has millions: G = 8, CH = 128.
template <int G, int CH>
IC_DEVINL float warp_exact_dsq_group(const float* __restrict__ pa, const float* pb, int np, int d4, float* sbuf) {
    const int lane = threadIdx.x & 31;
    constexpr int Q = CH / 128;
    constexpr int kStride = CH + 4;
    const float* pbk[G];
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const unsigned long long v = __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(pb), k);
        pbk[k] = reinterpret_cast<const float*>(v);
    }
    float4 a[Q], b[G][Q];
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    auto issue = [&](int c0) {
#pragma unroll
        for (int q = 0; q < Q; ++q) {
            const int e = c0 + 4 * (lane + 32 * q);
            const bool ok = e < d4;
            a[q] = ok ? __ldcg(reinterpret_cast<const float4*>(pa + e)) : z;
#pragma unroll
            for (int k = 0; k < G; ++k) b[k][q] = (ok && k < np) ? __ldcg(reinterpret_cast<const float4*>(pbk[k] + e)) : z;
        }
    };
    float sum = 0.0f;
    issue(0);
    for (int c0 = 0; c0 < d4; c0 += CH) {
#pragma unroll
        for (int k = 0; k < G; ++k) {
            if (k < np) {
#pragma unroll
                for (int q = 0; q < Q; ++q) {
                    const float dx = __fsub_rn(a[q].x, b[k][q].x), dy = __fsub_rn(a[q].y, b[k][q].y), dz = __fsub_rn(a[q].z, b[k][q].z),
                                dw = __fsub_rn(a[q].w, b[k][q].w);
                    reinterpret_cast<float4*>(sbuf + k * kStride)[lane + 32 * q] =
                        make_float4(__fmul_rn(dx, dx), __fmul_rn(dy, dy), __fmul_rn(dz, dz), __fmul_rn(dw, dw));
                }
            }
        }
        __syncwarp();
        if (c0 + CH < d4) issue(c0 + CH);  // in flight while the chains run
        if (lane < np) sum = ex_chain_sum(reinterpret_cast<const float4*>(sbuf + lane * kStride), (min(CH, d4 - c0)) >> 2, sum);
        __syncwarp();
    }
    return sum;
}
constexpr int kExGroup = 4;                    // merge loop
constexpr int kExStride = kExChunk + 4;

// The same evaluation with the rows staged by cp.async (merge loop): a ring of S stages of CH floats of the G + 1 rows per
// warp, S - 1 chunks in flight while the chains of the current chunk run.  With register staging (above) one chunk is in
// flight, and a warp that has a single group per iteration -- the merge loop's case -- waits a full memory latency per chunk:
// round 2 measured 42 k cycles per iteration for a chain of 8.4 k.  Every lane copies, squares (in place, over the second
// rows) and publishes its own 16 bytes of every row; lanes 0 .. np-1 then run the chains.
constexpr int kExAsyncCH = 128;
constexpr int kExAsyncRow = kExAsyncCH + 4;                         // floats (the pad keeps the chains' 16-byte reads on distinct banks)
constexpr int kExAsyncStage = (kExGroup + 1) * kExAsyncRow;         // floats
template <int S>  // stages of the ring: S * kExAsyncStage floats (2 640 bytes each)
IC_DEVINL float warp_exact_dsq_group_async(const float* __restrict__ pa, const float* pb, int np, int d4, float* sbuf) {
    constexpr int G = kExGroup, CH = kExAsyncCH;
    static_assert(CH == 128, "one 16-byte copy per lane, row and chunk");
    const int lane = threadIdx.x & 31;
    const float* pbk[G];
#pragma unroll
    for (int k = 0; k < G; ++k) {
        const unsigned long long v = __shfl_sync(0xffffffffu, reinterpret_cast<unsigned long long>(pb), k);
        pbk[k] = reinterpret_cast<const float*>(v);
    }
    const uint32_t sbase = static_cast<uint32_t>(__cvta_generic_to_shared(sbuf)) + static_cast<uint32_t>(lane) * 16u;
    const int nch = (d4 + CH - 1) / CH;
    auto issue = [&](int c) {  // chunk c into stage c % S; always commits one group
        const int e = c * CH + 4 * lane;
        if (c < nch && e < d4) {
            const uint32_t dst = sbase + static_cast<uint32_t>((c % S) * kExAsyncStage * 4);
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(pa + e) : "memory");
#pragma unroll
            for (int k = 0; k < G; ++k)
                if (k < np)
                    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst + static_cast<uint32_t>((k + 1) * kExAsyncRow * 4)), "l"(pbk[k] + e) : "memory");
        }
        asm volatile("cp.async.commit_group;" ::: "memory");
    };
#pragma unroll
    for (int c = 0; c < S - 1; ++c) issue(c);
    float sum = 0.0f;
    for (int c = 0; c < nch; ++c) {
        issue(c + S - 1);  // into the stage of chunk c - 1, whose chains finished before the __syncwarp that ended the last trip
        asm volatile("cp.async.wait_group %0;" ::"n"(S - 1) : "memory");  // this lane's copies of chunk c have landed
        float* stg = sbuf + (c % S) * kExAsyncStage;
        if (c * CH + 4 * lane < d4) {
            const float4 a = reinterpret_cast<const float4*>(stg)[lane];
#pragma unroll
            for (int k = 0; k < G; ++k) {
                if (k < np) {
                    float4* q = reinterpret_cast<float4*>(stg + (k + 1) * kExAsyncRow) + lane;
                    const float4 b = *q;
                    const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y), dz = __fsub_rn(a.z, b.z), dw = __fsub_rn(a.w, b.w);
                    *q = make_float4(__fmul_rn(dx, dx), __fmul_rn(dy, dy), __fmul_rn(dz, dz), __fmul_rn(dw, dw));
                }
            }
        }
        __syncwarp();
        if (lane < np) sum = ex_chain_sum(reinterpret_cast<const float4*>(stg + (lane + 1) * kExAsyncRow), (min(CH, d4 - c * CH)) >> 2, sum);
        __syncwarp();
    }
    asm volatile("cp.async.wait_group 0;" ::: "memory");
    return sum;
}
constexpr int kExGroupR = 8, kExChunkR = 128;  // refine.cu
constexpr int kExStrideR = kExChunkR + 4;

// (float32(na * nb) / float32(na + nb)) * dsq: integer product first (clustering.go:142-144)
IC_DEVINL float ward_weight(int na, int nb, float dsq) {
    const float num = __ll2float_rn(static_cast<long long>(na) * static_cast<long long>(nb));
    const float den = __ll2float_rn(static_cast<long long>(na) + static_cast<long long>(nb));
    return __fmul_rn(__fdiv_rn(num, den), dsq);
}

// bookkeeping of one re-evaluation: how far off was the stored value?  Returns the relative error (beyond what a tensor-core
// Gram value may be off by); a violation of the filter tolerance is counted right away (rare), the running maximum is kept by
// the caller and published once (exact_monitor_flush): one atomic per evaluation on one address serialises the whole phase.
IC_DEVINL float exact_monitor(int32_t* ctl, float stored, float w, float eps_filter, float abs_slack) {
    if (!(stored < __uint_as_float(kMaxFloatBits))) return 0.0f;
    const float excess = fmaxf(fabsf(stored - w) - abs_slack, 0.0f);
    const float err = excess / fmaxf(w, 1e-30f);
    if (err > eps_filter) atomicAdd(ctl + CTL_FILTER_VIOL, 1);
    return (err == err) ? fminf(err, 1e30f) : 1e30f;
}
IC_DEVINL void exact_monitor_flush(int32_t* ctl, float max_err) {  // whole warp
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) max_err = fmaxf(max_err, __shfl_xor_sync(0xffffffffu, max_err, o));
    if ((threadIdx.x & 31) == 0 && max_err > 0.0f) atomicMax(ctl + CTL_FILTER_MAXERR, static_cast<int32_t>(__float_as_uint(max_err)));
}

}  // namespace ic
