// Micro-benchmark: per-SM throughput of cp.async.bulk (global -> shared) as a function of copy size and alignment.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/bench_bulk tools/bench_bulk.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
__device__ __forceinline__ uint32_t s32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__global__ void __launch_bounds__(128, 1) k(const uint8_t* src, size_t src_bytes, int copy_bytes, int ncopies, int src_misalign, int dst_stride,
                                            int iters, int depth, long long* cycles) {
    extern __shared__ __align__(1024) uint8_t sm[];
    __shared__ uint64_t bar[8];
    if (threadIdx.x == 0) {
        for (int i = 0; i < depth; ++i) asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(s32(bar + i)));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x != 0) return;
    const size_t stage_src = (size_t)ncopies * 4096 * 4;
    long long t0 = clock64();
    // keep `depth` stages in flight
    for (int it = 0; it < iters + depth; ++it) {
        const int s = it % depth;
        if (it >= depth) {
            uint32_t par = ((it / depth) - 1) & 1;
            asm volatile("{\n\t.reg .pred p;\n\tW:\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t@p bra D;\n\tbra W;\n\tD:\n\t}" ::"r"(s32(bar + s)), "r"(par) : "memory");
        }
        if (it < iters) {
            asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(s32(bar + s)), "r"(copy_bytes * ncopies) : "memory");
            size_t base = ((size_t)blockIdx.x * 7919 + (size_t)it * stage_src) % (src_bytes - stage_src - 65536);
            base &= ~(size_t)4095;
            for (int c = 0; c < ncopies; ++c)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(s32(sm + (size_t)s * 65536 + (size_t)c * dst_stride)),
                             "l"(src + base + (size_t)c * 16384 + src_misalign), "r"(copy_bytes), "r"(s32(bar + s)) : "memory");
        }
    }
    cycles[blockIdx.x] = clock64() - t0;
}
int main() {
    size_t bytes = 1ull << 30;
    uint8_t* src; cudaMalloc(&src, bytes); cudaMemset(src, 1, bytes);
    long long* cyc; cudaMallocManaged(&cyc, 148 * 8);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 3 * 65536);
    struct C { const char* name; int bytes, n, mis, dstride; } cases[] = {
        {"1 x 49152 aligned", 49152, 1, 0, 0},        {"8 x 1056 src+16 dst 2080-stride", 1056, 8, 16, 2080},
        {"8 x 1056 src aligned dst 2080-stride", 1056, 8, 0, 2080}, {"8 x 1152 aligned dst 2304-stride", 1152, 8, 0, 2304},
        {"8 x 2080 src+16 dst 2080", 2080, 8, 16, 2080}, {"8 x 2304 aligned", 2304, 8, 0, 2304}, {"8 x 2304 src+112", 2304, 8, 112, 2304},
        {"1 x 16384 aligned", 16384, 1, 0, 0}, {"4 x 12288 aligned", 12288, 4, 0, 12288}, {"48 x 1024 aligned", 1024, 48, 0, 1024}};
    for (auto& c : cases)
        for (int depth : {1, 3}) {
            const int iters = 400;
            k<<<148, 128, 3 * 65536>>>(src, bytes, c.bytes, c.n, c.mis, c.dstride, iters, depth, cyc);
            cudaError_t e = cudaDeviceSynchronize();
            if (e != cudaSuccess) { printf("%s: %s\n", c.name, cudaGetErrorString(e)); return 1; }
            double mx = 0; for (int i = 0; i < 148; ++i) mx = cyc[i] > mx ? cyc[i] : mx;
            printf("%-40s depth %d: %8.1f cycles/stage  %6.2f B/cycle/SM\n", c.name, depth, mx / iters, (double)c.bytes * c.n * iters / mx);
        }
    return 0;
}
