// Micro-benchmark: FP64 vector FMA (DFMA) vs FP64 tensor-core MMA (mma.sync.m8n8k4.f64) vs both interleaved, on sm_100a.
// Question (north star, item f): can the contraction part of the pair evaluators run on the tensor pipe *beside* the
// vector FP64 pipe that evaluates the kernel powers?  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double &c0, double &c1, double a, double b)
{
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};" : "+d"(c0), "+d"(c1) : "d"(a), "d"(b));
}

template <int MODE> __global__ void __launch_bounds__(256) k(double *out, int iters)
{
    double f[8], c[8][2];
    const double a = 1.0000001 + threadIdx.x * 1e-9, b = 0.9999999;
#pragma unroll
    for (int q = 0; q < 8; q++) { f[q] = q + threadIdx.x * 1e-9; c[q][0] = q; c[q][1] = -q; }
    for (int i = 0; i < iters; i++) {
        if (MODE == 0 || MODE == 2) {
#pragma unroll
            for (int q = 0; q < 8; q++) f[q] = fma(f[q], a, b);
        }
        if (MODE == 1 || MODE == 2) {
#pragma unroll
            for (int q = 0; q < 8; q++) dmma(c[q][0], c[q][1], a, b);
        }
    }
    double s = 0.;
#pragma unroll
    for (int q = 0; q < 8; q++) s += f[q] + c[q][0] + c[q][1];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}

template <int MODE> double run(int blocks, int iters, double *d)
{
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, 256>>>(d, 100);
    float best = 1e30f;
    for (int rep = 0; rep < 3; rep++) {
        cudaEventRecord(e0);
        k<MODE><<<blocks, 256>>>(d, iters);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        best = ms < best ? ms : best;
    }
    return best;
}

int main()
{
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    const int blocks = sms * 8, iters = 20000;
    double *d; cudaMalloc(&d, (size_t)blocks * 256 * sizeof(double));
    const double threads = (double)blocks * 256, warps = threads / 32;
    const double t0 = run<0>(blocks, iters, d), t1 = run<1>(blocks, iters, d), t2 = run<2>(blocks, iters, d);
    const double fl_v = 2. * 8 * iters * threads;                // DFMA: 2 flops per lane
    const double fl_t = 2. * 8 * 8 * 4 * 8 * iters * warps;      // m8n8k4: 256 FMA per warp instruction, 8 per iteration
    printf("DFMA only : %.3f ms  %.2f TFLOP/s\n", t0, fl_v / (t0 * 1e-3) / 1e12);
    printf("DMMA only : %.3f ms  %.2f TFLOP/s\n", t1, fl_t / (t1 * 1e-3) / 1e12);
    printf("both      : %.3f ms  (sum of the two alone %.3f ms, max %.3f ms): %s\n", t2, t0 + t1, t0 > t1 ? t0 : t1,
           t2 < 0.8 * (t0 + t1) ? "the pipes overlap" : "no overlap: one shared FP64 datapath");
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
