// Throughput of the MUFU ops the epilogues use (tanh.approx, ex2.approx, rcp.approx) and of packed FFMA2, per SM.
// nvcc -gencode arch=compute_100a,code=sm_100a -o mufu_rate mufu_rate.cu && ./mufu_rate
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void k(float* out, int iters) {
    float a[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) a[i] = 0.001f * (threadIdx.x + i);
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (OP == 0) asm volatile("tanh.approx.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 1) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 2) asm volatile("rcp.approx.ftz.f32 %0, %0;" : "+f"(a[i]));
            if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %0, %0;" : "+f"(a[i]));
            if (OP == 5) asm volatile("tanh.approx.bf16x2 %0, %0;" : "+r"(*reinterpret_cast<unsigned*>(&a[i])));
            if (OP == 6) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(*reinterpret_cast<unsigned*>(&a[i])));
            if (OP == 7) asm volatile("tanh.approx.f16x2 %0, %0;" : "+r"(*reinterpret_cast<unsigned*>(&a[i])));
            if (OP == 8) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(*reinterpret_cast<unsigned*>(&a[i])));
        }
        if (OP == 4) {
#pragma unroll
            for (int i = 0; i < 8; i += 2)
                asm volatile("{.reg .b64 r; mov.b64 r, {%0, %1}; fma.rn.f32x2 r, r, r, r; mov.b64 {%0, %1}, r;}" : "+f"(a[i]), "+f"(a[i + 1]));
        }
    }
    float s = 0;
    for (int i = 0; i < 8; ++i) s += a[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
}
template <int OP>
void run(const char* name, int per_iter) {
    float* out; cudaMalloc(&out, 148 * 1024 * 4);
    const int iters = 4096;
    k<OP><<<148, 1024>>>(out, 16);
    cudaEvent_t s, e; cudaEventCreate(&s); cudaEventCreate(&e);
    cudaEventRecord(s); k<OP><<<148, 1024>>>(out, iters); cudaEventRecord(e); cudaEventSynchronize(e);
    float ms; cudaEventElapsedTime(&ms, s, e);
    int clk; cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    double lane_ops = 1024.0 * iters * per_iter;                  // per SM
    double cycles = ms * 1e-3 * clk * 1e3;
    printf("%-22s %7.1f lane-ops / clk / SM   (%.3f ms)\n", name, lane_ops / cycles, ms);
    cudaFree(out);
}
int main() {
    run<0>("tanh.approx.f32", 8); run<1>("ex2.approx.ftz.f32", 8); run<2>("rcp.approx.ftz.f32", 8); run<3>("fma.rn.f32", 8); run<4>("fma.rn.f32x2 (pairs)", 8);
    run<5>("tanh.approx.bf16x2 (x2)", 16); run<6>("ex2.approx.bf16x2 (x2)", 16); run<7>("tanh.approx.f16x2 (x2)", 16); run<8>("ex2.approx.f16x2 (x2)", 16);
    return 0;
}
