// tma_probe.cu — how fast can one SM pull L2-resident data into shared memory with 1-D bulk copies
// (cp.async.bulk.shared::cluster.global, SASS UBLKCP)?  One persistent CTA per SM; a ring of NST
// stages of STAGE bytes, each refilled by NSPLIT bulk copies; consumers only wait on the mbarrier
// (no LDS), so the number is the pure copy rate.  Prints bytes/clk/SM and TB/s for a few shapes.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tma_probe tma_probe.cu
#include <cstdio>
#include <cstdlib>
#include <cuda_runtime.h>

__device__ __forceinline__ unsigned smem_u32(const void *p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long *bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(unsigned long long *bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_wait(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nW: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D;\nbra W;\nD:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// spin on the non-blocking test instead of the (suspending) try_wait
__device__ __forceinline__ void mbar_wait_test(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nW2: mbarrier.test_wait.parity.shared::cta.b64 p, [%0], %1;\n@p bra D2;\nbra W2;\nD2:\n}\n" ::"r"(smem_u32(bar)), "r"(parity) : "memory");
}
// try_wait with an explicit (short) suspend-time hint in ns
__device__ __forceinline__ void mbar_wait_hint(unsigned long long *bar, unsigned parity) {
    asm volatile("{\n.reg .pred p;\nW3: mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1, %2;\n@p bra D3;\nbra W3;\nD3:\n}\n" ::"r"(smem_u32(bar)), "r"(parity), "r"(20u) : "memory");
}
__device__ __forceinline__ void bulk_g2s(void *dst, const void *src, unsigned bytes, unsigned long long *bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

__global__ void __launch_bounds__(128, 1)
probe(const unsigned char *src, size_t region, int stage_bytes, int nst, int nsplit, int iters, int shared_src,
      long long *clk_out, int wait_mode) {
    extern __shared__ __align__(128) unsigned char sm[];
    unsigned long long *full = reinterpret_cast<unsigned long long *>(sm);
    unsigned char *ring = sm + 128;
    if (threadIdx.x == 0) {
        for (int s = 0; s < nst; ++s) mbar_init(&full[s], 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    // every CTA walks its own slice of the region (shared_src: all CTAs read the same addresses)
    const size_t per_cta = region / gridDim.x / stage_bytes * stage_bytes;
    const unsigned char *base = src + (shared_src ? 0 : (size_t)blockIdx.x * per_cta);
    const size_t span = shared_src ? region / stage_bytes * stage_bytes : per_cta;
    const int part = stage_bytes / nsplit;
    long long t0 = clock64();
    if (threadIdx.x == 0) {
        size_t off = 0;
        for (int i = 0; i < iters + nst; ++i) {
            const int s = i % nst;
            if (i >= nst) {
                const unsigned par = (unsigned)(((i - nst) / nst) & 1);
                if (wait_mode == 0) mbar_wait(&full[s], par);
                else if (wait_mode == 1) mbar_wait_test(&full[s], par);
                else mbar_wait_hint(&full[s], par);
            }
            if (i < iters) {
                mbar_expect_tx(&full[s], (unsigned)stage_bytes);
                for (int q = 0; q < nsplit; ++q)
                    bulk_g2s(ring + (size_t)s * stage_bytes + (size_t)q * part, base + off + (size_t)q * part, (unsigned)part, &full[s]);
                off += stage_bytes;
                if (off + stage_bytes > span) off = 0;
            }
        }
    }
    __syncthreads();
    long long t1 = clock64();
    if (threadIdx.x == 0) clk_out[blockIdx.x] = t1 - t0;
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    const size_t region = (size_t)96 << 20;      // L2-resident after the first pass (126 MB L2)
    unsigned char *src;
    long long *clk;
    cudaMalloc(&src, region);
    cudaMemset(src, 1, region);
    cudaMalloc(&clk, sms * sizeof(long long));
    cudaFuncSetAttribute(probe, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
    struct Cfg { int stage, nst, nsplit, shared, ctas; } cfgs[] = {
        {53248, 4, 2, 0, 0}, {53248, 4, 1, 0, 0}, {53248, 4, 8, 0, 0}, {53248, 4, 26, 0, 0},
        {26624, 8, 1, 0, 0}, {13312, 16, 1, 0, 0}, {4096, 32, 1, 0, 0}, {16384, 12, 1, 0, 0},
        {53248, 2, 2, 0, 0}, {53248, 4, 2, 1, 0}, {53248, 4, 2, 0, 16}, {53248, 4, 2, 0, 74},
    };
    for (int wait_mode = 0; wait_mode < 3; ++wait_mode)
    for (auto c : cfgs) {
        const int ctas = c.ctas ? c.ctas : sms;
        const int iters = (int)(((size_t)64 << 20) / c.stage);     // 64 MB per CTA
        const size_t smem = 128 + (size_t)c.stage * c.nst;
        float best = 1e30f;
        long long h[256];
        for (int rep = 0; rep < 3; ++rep) {
            cudaEvent_t e0, e1;
            cudaEventCreate(&e0); cudaEventCreate(&e1);
            cudaEventRecord(e0);
            probe<<<ctas, 128, smem>>>(src, region, c.stage, c.nst, c.nsplit, iters, c.shared, clk, wait_mode);
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        cudaMemcpy(h, clk, ctas * sizeof(long long), cudaMemcpyDeviceToHost);
        double mean = 0; for (int i = 0; i < ctas; ++i) mean += (double)h[i]; mean /= ctas;
        const double bytes = (double)iters * c.stage;
        cudaError_t e = cudaGetLastError();
        printf("wait %s: stage %6d B x %2d stages, %2d copies/stage, %s, %3d CTAs: %6.1f B/clk/SM, chip %5.2f TB/s  (%s)\n",
               wait_mode == 0 ? "try_wait " : wait_mode == 1 ? "test_wait" : "try+hint ", c.stage, c.nst, c.nsplit, c.shared ? "same addresses" : "own slices   ", ctas, bytes / mean,
               bytes * ctas / (best * 1e-3) / 1e12, cudaGetErrorString(e));
    }
    return 0;
}
