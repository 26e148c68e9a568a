// microbenchmark 2: P producer warps share the copies of a stage (each lane one copy), 1 consumer warp releases
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(160, 1) k(const char *src, size_t src_bytes, int copy_bytes, int copies, int slots, int stages, int P) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t full[8], empty[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < slots; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&full[s])), "r"(P));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&empty[s])), "r"(1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t slot_bytes = (uint32_t)copy_bytes * copies;
    if (warp < P) {
        uint32_t slot = 0, use = 0;
        size_t span = (size_t)32 << 20;
        size_t pos = (size_t)blockIdx.x * 7919 * 1024;
        for (int st = 0; st < stages; st++) {
            if (use > 0) {
                uint32_t ph = (use & 1u) ^ 1u;
                asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(su32(&empty[slot])), "r"(ph) : "memory");
            }
            // this warp's share of the bytes
            int mine = 0;
            for (int c = warp * 32 + lane; c < copies; c += 32 * P) mine++;
            int tot = mine;
            for (int o = 16; o; o >>= 1) tot += __shfl_xor_sync(0xffffffffu, tot, o);
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&full[slot])), "r"((uint32_t)tot * copy_bytes) : "memory");
            __syncwarp();
            for (int c = warp * 32 + lane; c < copies; c += 32 * P) {
                size_t off = (pos + (size_t)c * copy_bytes) % (span - copy_bytes) / 128 * 128;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(su32(sm + (size_t)slot * slot_bytes + (size_t)c * copy_bytes)), "l"(src + off), "r"(copy_bytes), "r"(su32(&full[slot])) : "memory");
            }
            pos += slot_bytes;
            if (++slot == (uint32_t)slots) { slot = 0; use++; }
        }
    } else if (warp == P) {
        uint32_t slot = 0, ph = 0;
        for (int st = 0; st < stages; st++) {
            asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(su32(&full[slot])), "r"(ph) : "memory");
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(su32(&empty[slot])) : "memory");
            if (++slot == (uint32_t)slots) { slot = 0; ph ^= 1u; }
        }
    }
}
int main() {
    const size_t bytes = (size_t)1 << 30;
    char *src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    for (int P : {1, 2, 4})
        for (int copy_bytes : {800, 1600, 3200}) {
            int slots = 2, copies = 100 * 1024 / copy_bytes, stages = 400;
            size_t smem = (size_t)copies * copy_bytes * slots;
            k<<<148, 160, smem>>>(src, bytes, copy_bytes, copies, slots, 20, P);
            cudaEventRecord(e0);
            k<<<148, 160, smem>>>(src, bytes, copy_bytes, copies, slots, stages, P);
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            double tot = (double)148 * stages * copies * copy_bytes;
            printf("P=%d copy %6d B x %3d /stage: %7.3f ms  %6.2f TB/s  (%5.1f B/clk/SM, %5.1f clk/copy) %s\n", P, copy_bytes, copies, ms,
                   tot / ms * 1e-9, tot / (ms * 1e-3) / 148 / 1.965e9, ms * 1e-3 * 1.965e9 / ((double)stages * copies), cudaGetErrorString(cudaGetLastError()));
        }
    return 0;
}
