// dmma_probe.cu — what DMMA rate does the ORG inner-loop SHAPE allow, with nothing else in the way?
// One CTA of W warps per SM (or two of W/2), every warp: k-steps of 4 fragment loads from shared
// memory (conflict-free layout of org.cu) + 4 DMMA m8n8k4 into NACC accumulator pairs; the
// fragments of the next step are loaded while the current one multiplies (ping-pong).
//   NACC = 4: every k-step accumulates into the same 4 tiles (ORG today: dependent at distance 4)
//   NACC = 8: consecutive k-steps alternate between two sets (two degrees interleaved)
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_probe dmma_probe.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &d0, double &d1, double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(d0), "+d"(d1) : "d"(a), "d"(b));
}

template <int NACC, int EXTRA_DFMA>
__global__ void __launch_bounds__(512, 1) probe(int iters, int ksteps, double *sink) {
    extern __shared__ double sm[];
    const int OS = 68, BS = 34, ROWS = 48;
    double *Y = sm, *B = sm + ROWS * OS;
    for (int i = threadIdx.x; i < ROWS * (OS + 2 * BS); i += blockDim.x) sm[i] = 1e-3 * (i % 97);
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, g = lane >> 2, t4 = lane & 3;
    const int wm = (warp >> 2) & 3, wn = warp & 3;
    const double *Yw = Y + t4 * OS + wm * 16 + g, *Bw = B + t4 * 2 * BS + wn * 8 + g;
    double T[NACC / 4][2][2][2];
    double h[4] = {1.0, 1.1, 1.2, 1.3}, p[4] = {0, 0, 0, 0};
#pragma unroll
    for (int s = 0; s < NACC / 4; ++s)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) { T[s][a][b][0] = 0.0; T[s][a][b][1] = 0.0; }
    auto load = [&](int ks, double (&af)[2], double (&bf)[2]) {
        const int r = (ks % (ROWS / 4)) * 4;
        af[0] = Yw[r * OS]; af[1] = Yw[r * OS + 8];
        bf[0] = Bw[r * 2 * BS]; bf[1] = Bw[r * 2 * BS + BS];
    };
    for (int it = 0; it < iters; ++it) {
        double a0[2], b0[2], a1[2], b1[2];
        load(0, a0, b0);
#pragma unroll 1
        for (int ks = 0; ks < ksteps; ks += 2) {
            load(ks + 1, a1, b1);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt) dmma(T[0][pl][mt][0], T[0][pl][mt][1], a0[mt], b0[pl]);
            load(ks + 2, a0, b0);
#pragma unroll
            for (int pl = 0; pl < 2; ++pl)
#pragma unroll
                for (int mt = 0; mt < 2; ++mt)
                    dmma(T[NACC / 4 - 1][pl][mt][0], T[NACC / 4 - 1][pl][mt][1], a1[mt], b1[pl]);
        }
        // the per-degree scalar work of ORG: EXTRA_DFMA dependent-ish FP64 instructions
#pragma unroll
        for (int e = 0; e < EXTRA_DFMA; ++e) p[e & 3] = fma(h[e & 3], T[0][e & 1][(e >> 1) & 1][(e >> 2) & 1], p[e & 3]);
    }
    double s = p[0] + p[1] + p[2] + p[3];
#pragma unroll
    for (int q = 0; q < NACC / 4; ++q)
#pragma unroll
        for (int a = 0; a < 2; ++a)
#pragma unroll
            for (int b = 0; b < 2; ++b) s += T[q][a][b][0] + T[q][a][b][1];
    if (s == 123.456) sink[0] = s;
}

template <int NACC, int EXTRA>
void run(const char *name, int threads, int ctas_per_sm, int ksteps, int sms, double *sink) {
    const int iters = 20000 / ksteps * 6;
    const size_t smem = 48 * (68 + 68) * 8;
    cudaFuncSetAttribute(probe<NACC, EXTRA>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    float best = 1e30f;
    for (int rep = 0; rep < 3; ++rep) {
        cudaEvent_t e0, e1;
        cudaEventCreate(&e0); cudaEventCreate(&e1);
        cudaEventRecord(e0);
        probe<NACC, EXTRA><<<sms * ctas_per_sm, threads, smem>>>(iters, ksteps, sink);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    const double dmmas = (double)sms * ctas_per_sm * (threads / 32) * (double)iters * ksteps * 4.0;
    printf("%-34s %4d thr x %d CTA/SM, %2d k-steps/degree: %6.2f TFLOP/s (DMMA only; + %d FP64/degree)  %s\n", name, threads,
           ctas_per_sm, ksteps, dmmas * 512.0 / (best * 1e-3) / 1e12, EXTRA, cudaGetErrorString(cudaGetLastError()));
}

int main() {
    int dev = 0, sms = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    double *sink;
    cudaMalloc(&sink, 64);
    for (int ks : {2, 6, 12, 48}) {
        run<4, 0>("same 4 accumulators", 512, 1, ks, sms, sink);
        run<8, 0>("two sets alternating", 512, 1, ks, sms, sink);
        run<4, 28>("same 4 acc + degree scalar work", 512, 1, ks, sms, sink);
        run<8, 28>("two sets + degree scalar work", 512, 1, ks, sms, sink);
        run<4, 0>("same 4 accumulators", 256, 2, ks, sms, sink);
        run<4, 0>("same 4 accumulators", 256, 1, ks, sms, sink);
        run<8, 0>("two sets alternating", 256, 1, ks, sms, sink);
        run<4, 0>("same 4 accumulators", 128, 1, ks, sms, sink);
    }
    return 0;
}
