// Pipe-rate micro-benchmarks used to ground DESIGN.md's cost model (DFMA vs MUFU vs SHFL vs LDS on B200).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o build/ubench tools/ubench.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

template <int MODE>
__global__ void k(double *out, int iters, double seed)
{
    double a0 = seed + threadIdx.x, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    float f0 = (float)a0, f1 = f0 + 1, f2 = f0 + 2, f3 = f0 + 3, f4 = f0 + 4, f5 = f0 + 5, f6 = f0 + 6, f7 = f0 + 7;
    __shared__ double sm[1024];
    sm[threadIdx.x % 1024] = a0;
    __syncthreads();
    for (int i = 0; i < iters; ++i) {
        if (MODE == 0) {  // DFMA
            a0 = fma(a0, 1.0000001, 1e-9); a1 = fma(a1, 1.0000001, 1e-9); a2 = fma(a2, 1.0000001, 1e-9); a3 = fma(a3, 1.0000001, 1e-9);
            a4 = fma(a4, 1.0000001, 1e-9); a5 = fma(a5, 1.0000001, 1e-9); a6 = fma(a6, 1.0000001, 1e-9); a7 = fma(a7, 1.0000001, 1e-9);
        } else if (MODE == 1) {  // FFMA
            f0 = fmaf(f0, 1.0000001f, 1e-9f); f1 = fmaf(f1, 1.0000001f, 1e-9f); f2 = fmaf(f2, 1.0000001f, 1e-9f); f3 = fmaf(f3, 1.0000001f, 1e-9f);
            f4 = fmaf(f4, 1.0000001f, 1e-9f); f5 = fmaf(f5, 1.0000001f, 1e-9f); f6 = fmaf(f6, 1.0000001f, 1e-9f); f7 = fmaf(f7, 1.0000001f, 1e-9f);
        } else if (MODE == 2) {  // MUFU ex2
            f0 = exp2f(f0) ; f1 = exp2f(f1); f2 = exp2f(f2); f3 = exp2f(f3); f4 = exp2f(f4); f5 = exp2f(f5); f6 = exp2f(f6); f7 = exp2f(f7);
            f0 = __fmaf_rn(f0, 0.f, 0.5f); f1 = __fmaf_rn(f1, 0.f, 0.5f); f2 = __fmaf_rn(f2, 0.f, 0.5f); f3 = __fmaf_rn(f3, 0.f, 0.5f);
            f4 = __fmaf_rn(f4, 0.f, 0.5f); f5 = __fmaf_rn(f5, 0.f, 0.5f); f6 = __fmaf_rn(f6, 0.f, 0.5f); f7 = __fmaf_rn(f7, 0.f, 0.5f);
        } else if (MODE == 3) {  // SHFL of doubles (2 SHFL each)
            a0 = __shfl_up_sync(0xffffffffu, a0, 1); a1 = __shfl_up_sync(0xffffffffu, a1, 1); a2 = __shfl_up_sync(0xffffffffu, a2, 1); a3 = __shfl_up_sync(0xffffffffu, a3, 1);
            a4 = __shfl_up_sync(0xffffffffu, a4, 1); a5 = __shfl_up_sync(0xffffffffu, a5, 1); a6 = __shfl_up_sync(0xffffffffu, a6, 1); a7 = __shfl_up_sync(0xffffffffu, a7, 1);
        } else if (MODE == 4) {  // LDS.64
            int j = (threadIdx.x + i) & 1023;
            a0 += sm[j]; a1 += sm[(j + 32) & 1023]; a2 += sm[(j + 64) & 1023]; a3 += sm[(j + 96) & 1023];
            a4 += sm[(j + 128) & 1023]; a5 += sm[(j + 160) & 1023]; a6 += sm[(j + 192) & 1023]; a7 += sm[(j + 224) & 1023];
        } else if (MODE == 5) {  // F2F f64->f32 and back
            f0 = (float)a0; a0 = (double)f0 + 1.0; f1 = (float)a1; a1 = (double)f1 + 1.0; f2 = (float)a2; a2 = (double)f2 + 1.0; f3 = (float)a3; a3 = (double)f3 + 1.0;
        } else if (MODE == 6) {  // dependent DFMA chain (latency)
            a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9);
            a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9); a0 = fma(a0, 1.0000001, 1e-9);
        } else if (MODE == 7) {  // dependent shuffle chain (latency)
            f0 = __shfl_up_sync(0xffffffffu, f0, 1); f0 = __shfl_up_sync(0xffffffffu, f0, 1); f0 = __shfl_up_sync(0xffffffffu, f0, 1); f0 = __shfl_up_sync(0xffffffffu, f0, 1);
            f0 = __shfl_up_sync(0xffffffffu, f0, 1); f0 = __shfl_up_sync(0xffffffffu, f0, 1); f0 = __shfl_up_sync(0xffffffffu, f0, 1); f0 = __shfl_up_sync(0xffffffffu, f0, 1);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7 + f0 + f1 + f2 + f3 + f4 + f5 + f6 + f7;
}

template <int MODE>
void run(const char *name, int ops_per_iter, int threads, int blocks_per_sm)
{
    int dev = 0, sms = 0, khz = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, dev);
    int blocks = sms * blocks_per_sm, iters = 20000;
    double *out;
    cudaMalloc(&out, sizeof(double) * blocks * threads);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    k<MODE><<<blocks, threads>>>(out, 100, 1.0);
    cudaEventRecord(e0);
    k<MODE><<<blocks, threads>>>(out, iters, 1.0);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    double ops = (double)blocks * threads * iters * ops_per_iter;
    double per_clk_sm = ops / (ms * 1e-3) / sms / (khz * 1e3);
    printf("%-28s threads=%4d blk/sm=%d  %.3f ms  %.2f Gop/s  %.1f ops/clk/SM (at max clock %d MHz)  %.1f cyc/iter/warp-ish\n", name, threads,
           blocks_per_sm, ms, ops / ms * 1e-6, per_clk_sm, khz / 1000, (ms * 1e-3) * (khz * 1e3) / iters);
    cudaFree(out);
}

int main()
{
    cudaDeviceProp p;
    cudaGetDeviceProperties(&p, 0);
    printf("device %s sm_%d%d SMs=%d L2=%d MB smem/SM=%zu KB\n", p.name, p.major, p.minor, p.multiProcessorCount, p.l2CacheSize >> 20,
           p.sharedMemPerMultiprocessor >> 10);
    run<0>("DFMA x8 indep", 8, 512, 2);
    run<0>("DFMA x8 indep (1 warp/SMSP)", 8, 128, 1);
    run<1>("FFMA x8 indep", 8, 512, 2);
    run<2>("MUFU.EX2 x8 (+8 FFMA)", 8, 512, 2);
    run<3>("SHFL f64 x8 (16 SHFL)", 16, 512, 2);
    run<4>("LDS.64 x8 (+8 DADD)", 8, 512, 2);
    run<5>("F2F 64<->32 x8", 8, 512, 2);
    run<6>("DFMA dependent x8 (1 warp)", 8, 32, 1);
    run<7>("SHFL dependent x8 (1 warp)", 8, 32, 1);
    return 0;
}
