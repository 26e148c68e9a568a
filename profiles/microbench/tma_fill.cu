// microbenchmark: aggregate fill rate of cp.async.bulk (global -> shared) per SM, 148 persistent CTAs,
// as a function of bytes per copy, copies per stage and ring depth; consumers release immediately.
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t su32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(64, 1) k(const char *src, size_t src_bytes, int copy_bytes, int copies, int slots, int stages, int mode, int imode) {
    extern __shared__ __align__(128) unsigned char sm[];
    __shared__ __align__(8) uint64_t full[8], empty[8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (threadIdx.x == 0) {
        for (int s = 0; s < slots; s++) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(su32(&empty[s])), "r"(1));
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    const uint32_t slot_bytes = (uint32_t)copy_bytes * copies;
    if (warp == 0) {
        // producer: lane l issues copies l, l+32, ...
        uint32_t slot = 0, use = 0;
        // mode 0: every CTA streams its own contiguous region (DRAM); mode 1: all CTAs re-read the first 32 MB (L2 hits);
        // mode 2: pseudo-random rows of the whole buffer
        size_t region = src_bytes / gridDim.x / 128 * 128;
        size_t base = mode == 0 ? (size_t)blockIdx.x * region : 0;
        size_t span = mode == 0 ? region : (mode == 1 ? (size_t)32 << 20 : src_bytes);
        size_t pos = (size_t)blockIdx.x * 7919 * 1024;
        for (int st = 0; st < stages; st++) {
            if (use > 0) {
                uint32_t ph = (use & 1u) ^ 1u;
                asm volatile("{ .reg .pred p; W: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1; @p bra D; bra W; D: }" ::"r"(su32(&empty[slot])), "r"(ph) : "memory");
            }
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(su32(&full[slot])), "r"(slot_bytes) : "memory");
            __syncwarp();
            if (imode == 0) {
            for (int c = lane; c < copies; c += 32) {
                size_t off;
                if (mode == 2) { off = ((pos + (size_t)c * 2654435761ull * 128) % (span - copy_bytes)) / 128 * 128; }
                else off = (pos + (size_t)c * copy_bytes) % (span - copy_bytes) / 128 * 128;
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(su32(sm + (size_t)slot * slot_bytes + (size_t)c * copy_bytes)), "l"(src + base + off), "r"(copy_bytes), "r"(su32(&full[slot])) : "memory");
            }
            } else {
            // offsets computed by the lanes in parallel, then issued by one lane / by the converged warp
            for (int c0 = 0; c0 < copies; c0 += 32) {
                const int c = c0 + lane;
                size_t off;
                if (mode == 2) { off = ((pos + (size_t)c * 2654435761ull * 128) % (span - copy_bytes)) / 128 * 128; }
                else off = (pos + (size_t)c * copy_bytes) % (span - copy_bytes) / 128 * 128;
                const int nn = min(32, copies - c0);
                for (int jj = 0; jj < nn; jj++) {
                    const unsigned long long o = __shfl_sync(0xffffffffu, (unsigned long long)off, jj);
                    if (imode == 1) {
                        if (lane == 0)
                        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(su32(sm + (size_t)slot * slot_bytes + (size_t)(c0 + jj) * copy_bytes)), "l"(src + base + o), "r"(copy_bytes), "r"(su32(&full[slot])) : "memory");
                    } else {
                        asm volatile("{ .reg .pred p; elect.sync _|p, 0xffffffff; @p cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3]; }"
                             ::"r"(su32(sm + (size_t)slot * slot_bytes + (size_t)(c0 + jj) * copy_bytes)), "l"(src + base + o), "r"(copy_bytes), "r"(su32(&full[slot])) : "memory");
                    }
                }
            }
            }
            pos += (mode == 2) ? 1000003ull * 128 : slot_bytes;
            if (++slot == (uint32_t)slots) { slot = 0; use++; }
        }
    } else {
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
    const char *mname[3] = {"stream(DRAM)", "reuse 32MB (L2)", "random"};
    for (int imode = 0; imode < 3; imode++)
    for (int mode = 1; mode < 3; mode++)
        for (int copy_bytes : {800, 1600, 3200})
            for (int slots : {2}) {
                for (int stage_kb : {100}) {
                    int copies = stage_kb * 1024 / copy_bytes; if (copies < 1) continue;
                    if ((size_t)copies * copy_bytes * slots > 216 * 1024) continue;
                    int stages = 400;
                    size_t smem = (size_t)copies * copy_bytes * slots;
                    k<<<148, 64, smem>>>(src, bytes, copy_bytes, copies, slots, 20, mode, imode);
                    cudaEventRecord(e0);
                    k<<<148, 64, smem>>>(src, bytes, copy_bytes, copies, slots, stages, mode, imode);
                    cudaEventRecord(e1); cudaEventSynchronize(e1);
                    float ms; cudaEventElapsedTime(&ms, e0, e1);
                    double tot = (double)148 * stages * copies * copy_bytes;
                    printf("imode %d %-16s copy %6d B x %3d /stage, %d slots: %7.3f ms  %6.2f TB/s  (%5.1f GB/s/SM, %5.1f B/clk/SM) %s\n", imode, mname[mode], copy_bytes, copies, slots, ms,
                           tot / ms * 1e-9, tot / ms * 1e-6 / 148, tot / (ms * 1e-3) / 148 / 1.965e9, cudaGetErrorString(cudaGetLastError()));
                }
            }
    return 0;
}
