// gram_tcgen05.cu -- K1: initial Ward distance matrix as a tensor-core Gram GEMM.
//
// Replaces ComputeInitialDistanceMatrix + WardDistance + DotFloat32
// (clustering.go:61-73, 136-157) for singleton clusters, where the Ward weight
// |a||b|/(|a|+|b|) is exactly 1/2:
//      d(i,j) = 0.5*||x_i - x_j||^2 = 0.5*(||x_i||^2 + ||x_j||^2) - <x_i, x_j>.
// The contraction <x_i, x_j> runs on the 5th-gen tensor cores:
//   * operands: the split x = s1 + rt written by K0 (prep.cu): s1 a short
//     fixed-point slice, rt the TF32 residual; both K-major fp32 containers,
//     fetched by TMA (cp.async.bulk.tensor, 128-byte swizzle) into a shared-memory ring;
//   * four tcgen05.mma kind::tf32 per k-step, issued by one thread, into TWO fp32
//     TMEM accumulators: s1*s1 alone (every partial sum is exactly representable, so
//     the tensor core's truncating accumulation loses nothing) and
//     rt*s1 + s1*rt + rt*rt (small: its truncation error is ~1e-9 of the result).
//     A plain 3xTF32 split into ONE accumulator measured 2.6e-4 relative error on the
//     distances at D=2048 (experiments/gram_error.py): the accumulator truncates ~1 ulp
//     of its magnitude per MMA, which the Gram identity's cancellation amplifies;
//   * epilogue warps: tcgen05.ld of both accumulators -> sum, norms, half, clamp in
//     double -> fp32, coalesced stores of BOTH triangles (direct rows through a
//     shared-memory transpose, mirrored rows straight from registers: lane = row
//     makes them contiguous).
// Only tiles that touch the lower triangle are computed.  Persistent kernel, one
// CTA per SM, static tile list in an L2-friendly order (built on the host).
//
// Roofline: tensor pipe.  Algorithmic flops = 2*D per unordered pair;
// the 4-product split issues 4x that on the pipe.
#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {
constexpr int BM = kGramBM, BN = kGramBN, BK = kGramBK;
constexpr int kStages = 2;
constexpr int kABytes = BM * BK * 4;   // 16 KB
constexpr int kBBytes = BN * BK * 4;   // 32 KB
constexpr int kStageBytes = 2 * kABytes + 2 * kBBytes;  // hi+lo of A and B: 96 KB
constexpr int kTmemCols = 512;                          // 2 accumulators x 256 fp32 columns
constexpr int kThreads = 256;
constexpr int kEpiWarp0 = 4;
constexpr int kStageTile = 32 * 33;  // floats, per epilogue warp

// kind::tf32 instruction descriptor (UMMA::InstrDescriptor): c=F32 [4,6), a=TF32 [7,10),
// b=TF32 [10,13), a/b K-major (bits 15,16 = 0), N>>3 at [17,23), M>>4 at [24,29).
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | (static_cast<uint32_t>(BN >> 3) << 17) |
                            (static_cast<uint32_t>(BM >> 4) << 24);

struct __align__(8) SmemTail {
    double norm_a[BM];
    double norm_b[BN];
    float stage[4][kStageTile];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tfull[1];
    uint64_t tempty[1];
    uint32_t tmem_slot;
};
}  // namespace

size_t gram_tcgen05_smem_bytes() { return 1024 + static_cast<size_t>(kStages) * kStageBytes + sizeof(SmemTail); }

__global__ void __launch_bounds__(kThreads, 1)
gram_tcgen05_kernel(const __grid_constant__ CUtensorMap map_hi, const __grid_constant__ CUtensorMap map_lo,
                    const int2* __restrict__ tiles, int n_tiles, int k_blocks, const double* __restrict__ norms,
                    float* __restrict__ dm, int64_t n, int64_t ld, int64_t row_begin, int64_t row_end, int terms) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    SmemTail* tail = reinterpret_cast<SmemTail*>(smem + static_cast<size_t>(kStages) * kStageBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_hi);
        tma_prefetch_desc(&map_lo);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&tail->full[s], 1);
            mbar_init(&tail->empty[s], 1);
        }
        mbar_init(&tail->tfull[0], 1);
        mbar_init(&tail->tempty[0], 4);  // one arrive per epilogue warp
        mbar_fence_init();
    }
    if (warp == 2) {
        tmem_alloc(&tail->tmem_slot, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tail->tmem_slot;

    if (warp == 0) {
        // ===== TMA producer =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                const int2 tile = tiles[t];
                const int row0 = tile.x * BM, col0 = tile.y * BN;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&tail->empty[stage], phase ^ 1u);
                    uint8_t* sb = smem + static_cast<size_t>(stage) * kStageBytes;
                    mbar_arrive_expect_tx(&tail->full[stage], kStageBytes);
                    const int kc = kb * BK;
                    tma_load_2d(sb, &map_hi, &tail->full[stage], kc, row0);
                    tma_load_2d(sb + kABytes, &map_lo, &tail->full[stage], kc, row0);
                    tma_load_2d(sb + 2 * kABytes, &map_hi, &tail->full[stage], kc, col0);
                    tma_load_2d(sb + 2 * kABytes + kABytes, &map_hi, &tail->full[stage], kc, col0 + 128);
                    tma_load_2d(sb + 2 * kABytes + kBBytes, &map_lo, &tail->full[stage], kc, col0);
                    tma_load_2d(sb + 2 * kABytes + kBBytes + kABytes, &map_lo, &tail->full[stage], kc, col0 + 128);
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
            }
        }
    } else if (warp == 1) {
        // ===== MMA issuer (one thread) =====
        if (lane == 0) {
            int stage = 0;
            uint32_t phase = 0;
            uint32_t acc_phase = 0;
            const uint32_t d_exact = tmem_base;        // s1*s1
            const uint32_t d_small = tmem_base + BN;   // rt*s1 + s1*rt + rt*rt
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&tail->tempty[0], acc_phase ^ 1u);  // epilogue drained both accumulators
                tc_fence_after();
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&tail->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa_hi = smem_u32(smem + static_cast<size_t>(stage) * kStageBytes);
                    const uint32_t sa_lo = sa_hi + kABytes;
                    const uint32_t sb_hi = sa_hi + 2 * kABytes;
                    const uint32_t sb_lo = sb_hi + kBBytes;
#pragma unroll
                    for (int k = 0; k < BK / 8; ++k) {  // UMMA_K = 8 for tf32 = 32 bytes along K
                        const uint32_t off = static_cast<uint32_t>(k) * 32u;
                        const uint64_t a_hi = umma_desc_k_sw128(sa_hi + off);
                        const uint64_t a_lo = umma_desc_k_sw128(sa_lo + off);
                        const uint64_t b_hi = umma_desc_k_sw128(sb_hi + off);
                        const uint64_t b_lo = umma_desc_k_sw128(sb_lo + off);
                        const uint32_t first = (kb | k) != 0 ? 1u : 0u;
                        uint32_t accum = first;
                        if (terms & 1) { umma_tf32(d_small, a_lo, b_hi, kIdesc, accum); accum = 1u; }
                        if (terms & 2) { umma_tf32(d_small, a_hi, b_lo, kIdesc, accum); accum = 1u; }
                        if (terms & 16) { umma_tf32(d_small, a_lo, b_lo, kIdesc, accum); accum = 1u; }
                        if (terms & 4) umma_tf32(d_exact, a_hi, b_hi, kIdesc, first);
                    }
                    umma_commit(&tail->empty[stage]);  // frees the smem stage when these MMAs retire
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(&tail->tfull[0]);  // both accumulators complete
                acc_phase ^= 1u;
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===== epilogue: TMEM -> registers -> Ward distance -> both triangles =====
        const int ew = warp - kEpiWarp0;      // TMEM lane quadrant of this warp
        const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..127
        float* stg = tail->stage[ew];
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int2 tile = tiles[t];
            const int64_t row0 = static_cast<int64_t>(tile.x) * BM, col0 = static_cast<int64_t>(tile.y) * BN;
            asm volatile("bar.sync 1, 128;" ::: "memory");  // previous tile's norm reads are done
            tail->norm_a[et] = norms[row0 + et];
            tail->norm_b[et] = norms[col0 + et];
            tail->norm_b[et + 128] = norms[col0 + et + 128];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            mbar_wait(&tail->tfull[0], acc_phase);
            tc_fence_after();
            const int64_t gi = row0 + ew * 32 + lane;  // this thread's matrix row
            const double ni = tail->norm_a[ew * 32 + lane];
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
            for (int c = 0; c < BN / 32; ++c) {
                uint32_t r[32], r2[32];
                tmem_ld_32x32(taddr0 + c * 32, r);        // exact part
                tmem_ld_32x32(taddr0 + BN + c * 32, r2);  // small part
                tmem_ld_wait();
                const int64_t gj0 = col0 + c * 32;
                if (gj0 > row0 + BM - 1) continue;  // chunk entirely above the diagonal (warp uniform)
                float v[32];
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const double g = static_cast<double>(__uint_as_float(r[q])) + static_cast<double>(__uint_as_float(r2[q]));
                    const double tq = 0.5 * (ni + tail->norm_b[c * 32 + q]) - g;
                    float f = static_cast<float>(tq);
                    v[q] = (tq < 0.0) ? 0.0f : f;  // clamp the cancellation residue; NaN stays NaN
                    if (terms & 8) v[q] = static_cast<float>(g);  // debug: the raw Gram value
                }
                // mirrored entries dm[j][i]: for a fixed column j the 32 lanes hold consecutive i
#pragma unroll
                for (int q = 0; q < 32; ++q) {
                    const int64_t gj = gj0 + q;
                    if (gj < gi && gi < n && gj >= row_begin && gj < row_end) __stcs(dm + (gj - row_begin) * ld + gi, v[q]);
                }
                // direct entries dm[i][j]: transpose through shared memory so lanes hold consecutive j
#pragma unroll
                for (int q = 0; q < 32; ++q) stg[lane * 33 + q] = v[q];
                __syncwarp();
                const int64_t gj = gj0 + lane;
#pragma unroll 4
                for (int rr = 0; rr < 32; ++rr) {
                    const int64_t gr = row0 + ew * 32 + rr;
                    const float val = stg[rr * 33 + lane];
                    if (gr < row_end && gr >= row_begin && gj <= gr)
                        __stcs(dm + (gr - row_begin) * ld + gj, gj == gr ? 0.0f : val);
                }
                __syncwarp();
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&tail->tempty[0]);
            acc_phase ^= 1u;
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

cudaError_t launch_gram_tcgen05(const GramPlan& plan, const double* norms, float* dm, int64_t n, int64_t ld,
                                int64_t row_begin, int64_t row_end, int num_sms, cudaStream_t s, int terms) {
    if (plan.n_tiles == 0) return cudaSuccess;
    const size_t smem = gram_tcgen05_smem_bytes();
    cudaError_t e = cudaFuncSetAttribute(gram_tcgen05_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    const int grid = plan.n_tiles < num_sms ? plan.n_tiles : num_sms;
    gram_tcgen05_kernel<<<grid, kThreads, smem, s>>>(plan.map_hi, plan.map_lo, plan.tiles, plan.n_tiles,
                                                     plan.k_blocks, norms, dm, n, ld, row_begin, row_end, terms);
    return cudaGetLastError();
}

}  // namespace ic
