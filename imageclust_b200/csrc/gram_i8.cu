// gram_i8.cu -- K1 (int8 path): initial Ward distance matrix as an EXACT integer tensor-core Gram GEMM.
//
// Replaces ComputeInitialDistanceMatrix + WardDistance + DotFloat32 (clustering.go:61-73, 136-157) for
// singleton clusters: d(i,j) = 0.5*(||x_i||^2 + ||x_j||^2) - <x_i, x_j>.
//
// Why integers: the tensor core's fp32 accumulation truncates (profiles/r01_gram_accuracy.md), so an fp32-accurate
// Gram value needs products whose partial sums are exactly representable.  K0' (prep.cu: split_i8_kernel) rounds
// every centred row to a 22-bit fixed-point value v = h*2^14 + m*2^7 + l with balanced base-128 digits; the digit
// products are accumulated by tcgen05.mma kind::i8 in 32-bit integers -- exactly -- at 4x the TF32 rate with a
// quarter of the operand bytes:
//      <v_i, v_j> = 2^28 <h,h'> + 2^21 (<h,m'> + <m,h'>) + 2^14 (<h,l'> + <m,m'> + <l,h'>) + [2^7 (<m,l'>+<l,m'>) + <l,l'>]
// Six MMAs per k-step into THREE int32 TMEM accumulators (one per weight); the bracket is dropped when D >= 1024
// (its products are zero-mean noise of relative size ~ 2^-13 / sqrt(D) on a within-cluster distance); shorter rows
// also accumulate the 2^7 products in a fourth accumulator.  The epilogue combines the accumulators in 64-bit integers, scales by
// the two rows' quanta (powers of two: exact), applies the norms in double and stores fp32.
//
// Layout: 128 x 128 output tiles, K-blocks of 128 int8 (one 128-byte swizzle row), 2-stage 96 KB TMA ring
// (cp.async.bulk.tensor), one MMA-issuing thread, 4 epilogue warps (tcgen05.ld 32x32b.x16), persistent, one CTA per SM.
// Roofline: tensor pipe (int8); algorithmic flops = 2*D per unordered pair, the pipe executes 6x that in int8 ops.
#include "common.cuh"
#include "kernels.h"

namespace ic {

namespace {
constexpr int T = kI8Tile;
constexpr int kStages = 2;
constexpr int kTileBytes = T * kI8BK;          // 16 KB: one digit of one operand
constexpr int kStageBytes = 6 * kTileBytes;    // A{h,m,l} + B{h,m,l}
constexpr int kTmemCols = 512;                 // 3 accumulators x 128 columns (allocation is a power of two)
constexpr int kEpiWarps = 8;                   // two per TMEM lane quadrant: each drains one half of the columns
constexpr int kThreads = (4 + kEpiWarps) * 32;
constexpr int kEpiWarp0 = 4;
constexpr int kChunk = 16;                     // accumulator columns per epilogue step

// kind::i8 instruction descriptor: c = S32 (2) at [4,6), a = b = signed 8 bit (1) at [7,10) / [10,13), K-major,
// N>>3 at [17,23), M>>4 at [24,29)
constexpr uint32_t kIdesc = (2u << 4) | (1u << 7) | (1u << 10) | (static_cast<uint32_t>(T >> 3) << 17) |
                            (static_cast<uint32_t>(T >> 4) << 24);

struct __align__(8) SmemTail {
    double norm_a[T];
    double norm_b[T];
    float qa[T];
    float qb[T];
    float stage[kEpiWarps][32 * (kChunk + 1)];
    uint64_t full[kStages];
    uint64_t empty[kStages];
    uint64_t tfull[1];
    uint64_t tempty[1];
    uint32_t tmem_slot;
};

IC_DEVINL void tmem_ld_32x16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
          "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
}  // namespace

static size_t gram_i8_smem_bytes() { return 1024 + static_cast<size_t>(kStages) * kStageBytes + sizeof(SmemTail); }

// kLow: also accumulate the 2^7 products <m,l'> + <l,m'> (a fourth accumulator).  Needed when D is small: the dropped
// products are noise of relative size ~ 1/sqrt(D) on the distance (measured 3.8e-5 at D = 96 with six products).
template <bool kLow>
__global__ void __launch_bounds__(kThreads, 1)
gram_i8_kernel(const __grid_constant__ CUtensorMap map_h, const __grid_constant__ CUtensorMap map_m,
               const __grid_constant__ CUtensorMap map_l, const int2* __restrict__ tiles, int n_tiles, int k_blocks,
               const double* __restrict__ norms, const float* __restrict__ quanta, float* __restrict__ dm, int64_t n,
               int64_t ld, int64_t row_begin, int64_t row_end, int debug) {
    // flags: bit 3 = store the lower triangle only (the batched merge loop never reads dm[j][i], j < i: half the stores);
    // experiments only, wrong results: bit 0 = no TMA loads, bit 1 = no MMAs, bit 2 = no epilogue stores
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    SmemTail* tail = reinterpret_cast<SmemTail*>(smem + static_cast<size_t>(kStages) * kStageBytes);

    const int warp = threadIdx.x >> 5;
    const int lane = threadIdx.x & 31;

    if (warp == 0 && lane == 0) {
        tma_prefetch_desc(&map_h);
        tma_prefetch_desc(&map_m);
        tma_prefetch_desc(&map_l);
    }
    if (warp == 1 && lane == 0) {
        for (int s = 0; s < kStages; ++s) {
            mbar_init(&tail->full[s], 1);
            mbar_init(&tail->empty[s], 1);
        }
        mbar_init(&tail->tfull[0], 1);
        mbar_init(&tail->tempty[0], kEpiWarps);  // one arrive per epilogue warp
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
                const int row0 = tile.x * T, col0 = tile.y * T;
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&tail->empty[stage], phase ^ 1u);
                    uint8_t* sb = smem + static_cast<size_t>(stage) * kStageBytes;
                    if (debug & 1) {
                        mbar_arrive(&tail->full[stage]);
                        if (++stage == kStages) {
                            stage = 0;
                            phase ^= 1u;
                        }
                        continue;
                    }
                    mbar_arrive_expect_tx(&tail->full[stage], kStageBytes);
                    const int kc = kb * kI8BK;
                    tma_load_2d(sb + 0 * kTileBytes, &map_h, &tail->full[stage], kc, row0);
                    tma_load_2d(sb + 1 * kTileBytes, &map_m, &tail->full[stage], kc, row0);
                    tma_load_2d(sb + 2 * kTileBytes, &map_l, &tail->full[stage], kc, row0);
                    tma_load_2d(sb + 3 * kTileBytes, &map_h, &tail->full[stage], kc, col0);
                    tma_load_2d(sb + 4 * kTileBytes, &map_m, &tail->full[stage], kc, col0);
                    tma_load_2d(sb + 5 * kTileBytes, &map_l, &tail->full[stage], kc, col0);
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
            const uint32_t d28 = tmem_base;           // <h,h'>
            const uint32_t d21 = tmem_base + T;       // <h,m'> + <m,h'>
            const uint32_t d14 = tmem_base + 2 * T;   // <h,l'> + <m,m'> + <l,h'>
            const uint32_t d7 = tmem_base + 3 * T;    // <m,l'> + <l,m'>   (kLow only)
            for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
                mbar_wait(&tail->tempty[0], acc_phase ^ 1u);  // epilogue drained the accumulators
                tc_fence_after();
                for (int kb = 0; kb < k_blocks; ++kb) {
                    mbar_wait(&tail->full[stage], phase);
                    tc_fence_after();
                    const uint32_t sa = smem_u32(smem + static_cast<size_t>(stage) * kStageBytes);
                    const uint32_t sbb = sa + 3 * kTileBytes;
#pragma unroll
                    for (int k = 0; k < ((debug & 2) ? 0 : kI8BK / 32); ++k) {  // UMMA_K = 32 for 8-bit operands = 32 bytes along K
                        const uint32_t off = static_cast<uint32_t>(k) * 32u;
                        const uint64_t a_h = umma_desc_k_sw128(sa + 0 * kTileBytes + off);
                        const uint64_t a_m = umma_desc_k_sw128(sa + 1 * kTileBytes + off);
                        const uint64_t a_l = umma_desc_k_sw128(sa + 2 * kTileBytes + off);
                        const uint64_t b_h = umma_desc_k_sw128(sbb + 0 * kTileBytes + off);
                        const uint64_t b_m = umma_desc_k_sw128(sbb + 1 * kTileBytes + off);
                        const uint64_t b_l = umma_desc_k_sw128(sbb + 2 * kTileBytes + off);
                        const uint32_t acc = (kb | k) != 0 ? 1u : 0u;
                        umma_i8(d28, a_h, b_h, kIdesc, acc);
                        umma_i8(d21, a_h, b_m, kIdesc, acc);
                        umma_i8(d21, a_m, b_h, kIdesc, 1u);
                        umma_i8(d14, a_h, b_l, kIdesc, acc);
                        umma_i8(d14, a_m, b_m, kIdesc, 1u);
                        umma_i8(d14, a_l, b_h, kIdesc, 1u);
                        if (kLow) {
                            umma_i8(d7, a_m, b_l, kIdesc, acc);
                            umma_i8(d7, a_l, b_m, kIdesc, 1u);
                        }
                    }
                    umma_commit(&tail->empty[stage]);  // frees the smem stage when these MMAs retire
                    if (++stage == kStages) {
                        stage = 0;
                        phase ^= 1u;
                    }
                }
                umma_commit(&tail->tfull[0]);  // the three accumulators are complete
                acc_phase ^= 1u;
            }
        }
    } else if (warp >= kEpiWarp0) {
        // ===== epilogue: TMEM -> registers -> Ward distance -> both triangles =====
        const int ew = warp & 3;                      // TMEM lane quadrant of this warp (hardware: warp id % 4)
        const int eh = (warp - kEpiWarp0) >> 2;       // which half of the tile's columns
        const int et = threadIdx.x - kEpiWarp0 * 32;  // 0..255
        float* stg = tail->stage[warp - kEpiWarp0];
        uint32_t acc_phase = 0;
        for (int t = blockIdx.x; t < n_tiles; t += gridDim.x) {
            const int2 tile = tiles[t];
            const int64_t row0 = static_cast<int64_t>(tile.x) * T, col0 = static_cast<int64_t>(tile.y) * T;
            asm volatile("bar.sync 1, 256;" ::: "memory");  // previous tile's norm reads are done
            if (et < T) {
                tail->norm_a[et] = norms[row0 + et];
                tail->qa[et] = quanta[row0 + et];
            } else {
                tail->norm_b[et - T] = norms[col0 + et - T];
                tail->qb[et - T] = quanta[col0 + et - T];
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            mbar_wait(&tail->tfull[0], acc_phase);
            tc_fence_after();
            const int64_t gi = row0 + ew * 32 + lane;  // this thread's matrix row
            const double ni = tail->norm_a[ew * 32 + lane];
            const double qi = static_cast<double>(tail->qa[ew * 32 + lane]);
            const uint32_t taddr0 = tmem_base + (static_cast<uint32_t>(ew * 32) << 16);
#pragma unroll 1
            for (int c = eh * (T / kChunk / 2); c < (eh + 1) * (T / kChunk / 2); ++c) {
                uint32_t r28[kChunk], r21[kChunk], r14[kChunk], r7[kLow ? kChunk : 1];
                tmem_ld_32x16(taddr0 + c * kChunk, r28);
                tmem_ld_32x16(taddr0 + T + c * kChunk, r21);
                tmem_ld_32x16(taddr0 + 2 * T + c * kChunk, r14);
                if (kLow) tmem_ld_32x16(taddr0 + 3 * T + c * kChunk, reinterpret_cast<uint32_t(&)[kChunk]>(r7));
                tmem_ld_wait();
                const int64_t gj0 = col0 + c * kChunk;
                if (gj0 > row0 + T - 1 || (debug & 4)) continue;  // chunk entirely above the diagonal (warp uniform)
                float v[kChunk];
#pragma unroll
                for (int q = 0; q < kChunk; ++q) {
                    // exact: every term is a multiple of 2^7 and the total is below 2^53
                    long long tot = (static_cast<long long>(static_cast<int32_t>(r28[q])) << 28) +
                                    (static_cast<long long>(static_cast<int32_t>(r21[q])) << 21) +
                                    (static_cast<long long>(static_cast<int32_t>(r14[q])) << 14);
                    if (kLow) tot += static_cast<long long>(static_cast<int32_t>(r7[kLow ? q : 0])) << 7;
                    const double g = static_cast<double>(tot) * (qi * static_cast<double>(tail->qb[c * kChunk + q]));
                    const double tq = 0.5 * (ni + tail->norm_b[c * kChunk + q]) - g;
                    const float f = static_cast<float>(tq);
                    v[q] = (tq < 0.0) ? 0.0f : f;  // clamp the rounding residue of (near-)duplicates; NaN stays NaN
                }
                // mirrored entries dm[j][i]: for a fixed column j the 32 lanes hold consecutive i
#pragma unroll
                for (int q = 0; q < ((debug & 8) ? 0 : kChunk); ++q) {
                    const int64_t gj = gj0 + q;
                    if (gj < gi && gi < n && gj >= row_begin && gj < row_end) __stcs(dm + (gj - row_begin) * ld + gi, v[q]);
                }
                // direct entries dm[i][j]: transpose through shared memory so that lanes hold consecutive j
#pragma unroll
                for (int q = 0; q < kChunk; ++q) stg[lane * (kChunk + 1) + q] = v[q];
                __syncwarp();
                // a chunk wholly below the diagonal and inside the matrix: 16-byte stores, 8 rows per instruction
                const int64_t wr0 = row0 + ew * 32;  // first row of this warp
                if (gj0 + kChunk - 1 < wr0 && wr0 >= row_begin && wr0 + 32 <= row_end && wr0 + 32 <= n) {
                    const int c4 = (lane & 3) * 4;
#pragma unroll
                    for (int rr = 0; rr < 32; rr += 8) {
                        const int rloc = rr + (lane >> 2);
                        const float* sp = stg + rloc * (kChunk + 1) + c4;
                        __stcs(reinterpret_cast<float4*>(dm + (wr0 + rloc - row_begin) * ld + gj0 + c4),
                               make_float4(sp[0], sp[1], sp[2], sp[3]));
                    }
                    __syncwarp();
                    continue;
                }
                const int64_t gj = gj0 + (lane & (kChunk - 1));
#pragma unroll 4
                for (int rr = 0; rr < 32; rr += 2) {
                    const int rloc = rr + (lane >> 4);
                    const int64_t gr = row0 + ew * 32 + rloc;
                    const float val = stg[rloc * (kChunk + 1) + (lane & (kChunk - 1))];
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

cudaError_t launch_gram_i8(const GramI8Plan& plan, const double* norms, const float* quanta, float* dm, int64_t n,
                           int64_t ld, int64_t row_begin, int64_t row_end, int num_sms, cudaStream_t s, int debug) {
    if (plan.n_tiles == 0) return cudaSuccess;
    const size_t smem = gram_i8_smem_bytes();
    const int grid = plan.n_tiles < num_sms ? plan.n_tiles : num_sms;
    const bool low = plan.k_blocks * kI8BK < 1024;  // short rows: keep the 2^7 products as well
    cudaError_t e = low ? cudaFuncSetAttribute(gram_i8_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem))
                        : cudaFuncSetAttribute(gram_i8_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, static_cast<int>(smem));
    if (e != cudaSuccess) return e;
    if (low)
        gram_i8_kernel<true><<<grid, kThreads, smem, s>>>(plan.map_h, plan.map_m, plan.map_l, plan.tiles, plan.n_tiles,
                                                          plan.k_blocks, norms, quanta, dm, n, ld, row_begin, row_end, debug);
    else
        gram_i8_kernel<false><<<grid, kThreads, smem, s>>>(plan.map_h, plan.map_m, plan.map_l, plan.tiles, plan.n_tiles,
                                                           plan.k_blocks, norms, quanta, dm, n, ld, row_begin, row_end, debug);
    return cudaGetLastError();
}

}  // namespace ic
