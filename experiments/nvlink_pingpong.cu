// Experiment (2 GPUs, one process): latency of the primitives the sharded merge loop uses between ranks.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o experiments/nvlink_pingpong experiments/nvlink_pingpong.cu
// Each GPU runs one block; a "ball" (a counter) is written into the PEER's memory and polled in LOCAL memory.
//   mode 0: st.volatile to the peer                       mode 1: fence.acq_rel.sys + st.volatile
//   mode 2: bulk asynchronous store (TMA) of 16 bytes     mode 3: mode 0 + acquire fence.sys after every poll
//   mode 4: remote LOAD latency (dependent ld.volatile from the peer); modes 5/6: the same with ld.cg / ld.ca on the
//           same two words (a cached copy on this GPU would show as a short latency)
#include <cstdint>
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return static_cast<uint32_t>(__cvta_generic_to_shared(p)); }

__global__ void pingpong(volatile uint32_t* mine, uint32_t* peer, int me, int rounds, int mode, long long* cycles) {
    __shared__ __align__(16) uint32_t stage[4];
    if (threadIdx.x != 0) return;
    const long long t0 = clock64();
    if (mode == 5 || mode == 6) {  // are .cg / default loads of PEER memory cached on this GPU?  (same two words, re-read)
        uint32_t v = 0;
        for (int r = 0; r < rounds; ++r) {
            uint32_t x;
            if (mode == 5)
                asm volatile("ld.global.cg.u32 %0, [%1];" : "=r"(x) : "l"(peer + (v & 1)) : "memory");
            else
                asm volatile("ld.global.ca.u32 %0, [%1];" : "=r"(x) : "l"(peer + (v & 1)) : "memory");
            v += x + 1;
        }
        cycles[0] = clock64() - t0;
        cycles[1] = v;
        return;
    }
    if (mode == 4) {  // dependent remote loads
        uint32_t v = 0;
        for (int r = 0; r < rounds; ++r) {
            uint32_t x;
            asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(x) : "l"(peer + (v & 1)) : "memory");
            v += x + 1;
        }
        cycles[0] = clock64() - t0;
        cycles[1] = v;
        return;
    }
    for (int r = 1; r <= rounds; ++r) {
        const uint32_t ball = 2 * r - (me == 0 ? 1 : 0);  // GPU 0 serves first
        if (me == 1) {  // wait for the serve
            uint32_t spins = 0;
            while (*mine < static_cast<uint32_t>(2 * r - 1)) if (++spins > (1u << 26)) { cycles[0] = -1; return; }
            if (mode == 3) asm volatile("fence.acq_rel.sys;" ::: "memory");
        }
        if (mode == 1) asm volatile("fence.acq_rel.sys;" ::: "memory");
        if (mode == 2) {
            stage[0] = ball;
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 16;" ::"l"(__cvta_generic_to_global(peer)),
                         "r"(smem_u32(stage)) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
        } else {
            asm volatile("st.volatile.global.u32 [%0], %1;" ::"l"(peer), "r"(ball) : "memory");
        }
        if (me == 0) {  // wait for the return
            uint32_t spins = 0;
            while (*mine < static_cast<uint32_t>(2 * r)) if (++spins > (1u << 26)) { cycles[0] = -1; return; }
            if (mode == 3) asm volatile("fence.acq_rel.sys;" ::: "memory");
        }
    }
    cycles[0] = clock64() - t0;
}

int main() {
    int n = 0;
    CK(cudaGetDeviceCount(&n));
    if (n < 2) { printf("needs 2 GPUs\n"); return 0; }
    uint32_t* box[2];
    long long* cyc[2];
    cudaStream_t st[2];
    for (int d = 0; d < 2; ++d) {
        CK(cudaSetDevice(d));
        CK(cudaDeviceEnablePeerAccess(1 - d, 0));
        CK(cudaMalloc(&box[d], 256));
        CK(cudaMalloc(&cyc[d], 16));
        CK(cudaStreamCreate(&st[d]));
    }
    const int rounds = 20000;
    const char* names[] = {"st.volatile", "fence.sys + st.volatile", "bulk async store (TMA)", "st.volatile + acquire fence.sys", "dependent remote ld.volatile", "dependent remote ld.cg (same words)", "dependent remote ld.ca (same words)"};
    for (int mode = 0; mode < 7; ++mode) {
        for (int d = 0; d < 2; ++d) {
            CK(cudaSetDevice(d));
            CK(cudaMemset(box[d], 0, 256));
            CK(cudaDeviceSynchronize());
        }
        for (int d = 0; d < 2; ++d) {
            CK(cudaSetDevice(d));
            if (mode >= 4 && d == 1) continue;
            pingpong<<<1, 32, 0, st[d]>>>(box[d], box[1 - d], d, rounds, mode, cyc[d]);
        }
        long long h[2] = {0, 0};
        for (int d = 0; d < 2; ++d) {
            CK(cudaSetDevice(d));
            CK(cudaDeviceSynchronize());
        }
        CK(cudaSetDevice(0));
        CK(cudaMemcpy(h, cyc[0], 16, cudaMemcpyDeviceToHost));
        if (mode >= 4)
            printf("%-34s %8.0f cycles per load\n", names[mode], double(h[0]) / rounds);
        else
            printf("%-34s %8.0f cycles per round trip (2 one-way hops)\n", names[mode], double(h[0]) / rounds);
    }
    return 0;
}
