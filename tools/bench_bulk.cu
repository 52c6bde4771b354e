// Micro-benchmark: per-SM throughput of cp.async.bulk (global -> shared) as a function of copy size, copies per stage,
// pipeline depth, issuing pattern (one thread vs one lane per copy) and source footprint (HBM vs L2 resident).
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_bulk tools/bench_bulk.cu && /tmp/bench_bulk
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void wait(uint64_t* bar, uint32_t par) {
    asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(bar)), "r"(par) : "memory");
}
// lanes: 0 = every copy of a stage issued by thread 0; 1 = copy c issued by lane c of warp 0
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, size_t src_bytes, int copy_bytes, int ncopies, int slot_bytes, int iters,
                                            int depth, int dshift, int lanes, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[16];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x >= 32) return;
    const int lane = threadIdx.x;
    // lean loop: power-of-two depth and footprint, no divisions (the uniform datapath makes them very slow)
    const uint32_t dmask = depth - 1, fmask = (uint32_t)(src_bytes >> 12) - 1;   // footprint in 4 KB pages
    uint32_t page = blockIdx.x * 7919u + lane * 16u;   // copy c reads 64 KB further on
    const uint32_t sm0 = s32(sm), bar0 = s32(bar);
    long long t0 = clock64();
    for (int it = 0; it < iters + depth; ++it) {
        const uint32_t s = it & dmask;
        const uint32_t b = bar0 + s * 8;
        if (it >= depth) wait(bar + s, ((it >> dshift) - 1) & 1);
        if (it < iters) {
            if (lane == 0) asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(b), "r"(copy_bytes * ncopies) : "memory");
            __syncwarp();
            const uint32_t dst = sm0 + s * slot_bytes;
            if (lanes) {
                if (lane < ncopies)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + lane * copy_bytes),
                                 "l"(src + ((size_t)(page & fmask) << 12)), "r"(copy_bytes), "r"(b) : "memory");
            } else if (lane == 0) {
                for (int c = 0; c < ncopies; ++c)
                    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(dst + c * copy_bytes),
                                 "l"(src + ((size_t)((page + c * 16u) & fmask) << 12)), "r"(copy_bytes), "r"(b) : "memory");
            }
            page += ncopies * 16u;
            __syncwarp();
        }
    }
    if (lane == 0) cycles[blockIdx.x] = clock64() - t0;
}
int main() {
    size_t bytes = 1ull << 30;
    uint8_t* src; cudaMalloc(&src, bytes + (1 << 20)); cudaMemset(src, 1, bytes + (1 << 20));
    long long* cyc; cudaMallocManaged(&cyc, 148 * 8);
    const int SMEM = 212 * 1024;
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
    struct C { int bytes, n; } cases[] = {{2080, 5}, {2080, 16}, {4160, 4}, {8320, 2}, {16640, 1}, {16640, 2}, {33280, 1}, {6144, 1}, {24576, 1}, {49152, 1}, {1024, 16}};
    for (size_t footprint : {(size_t)1 << 30, (size_t)32 << 20})
        for (auto& c : cases)
            for (int lanes : {0, 1})
                for (int depth : {2, 4, 8, 16}) {
                    const int slot = ((c.bytes * c.n + 1023) / 1024) * 1024;
                    if (slot * depth > SMEM) continue;
                    if (lanes && c.n == 1) continue;
                    const int iters = 2000;
                    int dshift = 0; while ((1 << dshift) < depth) ++dshift;
                    k<<<148, 128, SMEM>>>(src, footprint, c.bytes, c.n, slot, iters, depth, dshift, lanes, cyc);
                    cudaError_t e = cudaDeviceSynchronize();
                    if (e != cudaSuccess) { printf("%d x %d: %s\n", c.n, c.bytes, cudaGetErrorString(e)); return 1; }
                    double mx = 0; for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
                    printf("src %4zu MB  %2d x %5d B  %s  depth %2d: %8.1f cycles/stage %7.1f cycles/copy %6.2f B/cycle/SM\n", footprint >> 20, c.n, c.bytes,
                           lanes ? "lane-per-copy" : "one thread   ", depth, mx / iters, mx / iters / c.n, (double)c.bytes * c.n * iters / mx);
                }
    return 0;
}
