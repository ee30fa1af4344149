// hmma_rate.cu — latency and throughput of legacy warp-level mma.sync (m16n8k16, f16 / bf16 in, f32 accumulate) on sm_100a.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o hmma_rate hmma_rate.cu && ./hmma_rate
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

template <int CHAINS, bool F16>
__global__ void __launch_bounds__(1024) k(float* out, int iters, long long* cycles) {
    float c[CHAINS][4];
    for (int i = 0; i < CHAINS; ++i)
        for (int j = 0; j < 4; ++j) c[i][j] = threadIdx.x * 0.001f + i;
    uint32_t a[4] = {0x3c003c00u + threadIdx.x, 0x3c003c00u, 0x3c003c00u, 0x3c003c00u}, b[2] = {0x3c003c00u, 0x3c003c00u};
    __syncthreads();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < CHAINS; ++i) {
            if (F16)
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b[0]), "r"(b[1]));
        }
    }
    const long long t1 = clock64();
    float s = 0;
    for (int i = 0; i < CHAINS; ++i)
        for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cycles = t1 - t0;
}

template <int CHAINS, bool F16>
void run(int warps_per_sm, int sms) {
    float* out;
    long long* cyc;
    cudaMalloc(&out, sizeof(float) * 1024 * sms);
    cudaMalloc(&cyc, sizeof(long long));
    const int iters = 2000;
    k<CHAINS, F16><<<sms, warps_per_sm * 32>>>(out, iters, cyc);
    cudaDeviceSynchronize();
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    k<CHAINS, F16><<<sms, warps_per_sm * 32>>>(out, iters, cyc);
    cudaEventRecord(e1);
    cudaDeviceSynchronize();
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    long long h; cudaMemcpy(&h, cyc, sizeof(h), cudaMemcpyDeviceToHost);
    const double per_warp = (double)h / ((double)iters * CHAINS);                       // cycles per HMMA as one warp sees it
    const double per_smsp = (double)h / ((double)iters * CHAINS * (warps_per_sm / 4.0)); // cycles per HMMA per scheduler
    const double tflops = 2.0 * 16 * 8 * 16 * (double)iters * CHAINS * warps_per_sm * sms / (ms * 1e-3) / 1e12;
    printf("%s chains=%d warps/SM=%2d: %.1f cyc/HMMA/warp, %.2f cyc/HMMA/SMSP, %.1f TFLOP/s (%d SMs, %.3f ms, err=%s)\n", F16 ? "f16 " : "bf16", CHAINS,
           warps_per_sm, per_warp, warps_per_sm >= 4 ? per_smsp : 0.0, tflops, sms, ms, cudaGetErrorString(cudaGetLastError()));
    cudaFree(out); cudaFree(cyc);
}

int main() {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    run<1, true>(1, sms);     // dependent chain: latency
    run<2, true>(1, sms);
    run<4, true>(1, sms);
    run<8, true>(1, sms);
    run<1, true>(4, sms);
    run<4, true>(4, sms);
    run<8, true>(4, sms);
    run<4, true>(8, sms);
    run<4, true>(16, sms);
    run<8, true>(16, sms);
    run<4, true>(32, sms);
    run<4, false>(16, sms);
    run<8, false>(32, sms);
    return 0;
}
