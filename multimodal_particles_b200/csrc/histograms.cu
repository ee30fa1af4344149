// histograms.cu — validation histograms of a generated batch, accumulated on the device so that
// only a few KB of int64 counts cross NVLink (all-reduce) instead of the particle clouds.
// Observables that need no clustering (mp/data/particle_clouds/jets.py:90-107): per-particle
// histograms of the continuous features, token frequencies, particle multiplicity per jet.
#include "mmb_internal.h"

namespace mmb {

// counts layout: [Dc][bins] feature histograms | [S] token counts | [max_mult + 1] multiplicities
__global__ void __launch_bounds__(128)
validation_histograms_kernel(const float* __restrict__ x, const uint8_t* __restrict__ k, const uint8_t* __restrict__ mask,
                             int B, int N, int Dc, int S, int bins, float lo, float scale, int max_mult,
                             unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int hist[];  // Dc*bins + S + max_mult + 1
    const int size = Dc * bins + S + max_mult + 1;
    for (int i = threadIdx.x; i < size; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    for (int jet = blockIdx.x; jet < B; jet += gridDim.x) {
        int live_here = 0;
        for (int n = threadIdx.x; n < N; n += blockDim.x) {
            const size_t p = (size_t)jet * N + n;
            if (mask[p]) {
                ++live_here;
                for (int c = 0; c < Dc; ++c) {
                    int b = (int)floorf((x[p * Dc + c] - lo) * scale);
                    b = b < 0 ? 0 : (b > bins - 1 ? bins - 1 : b);
                    atomicAdd(&hist[c * bins + b], 1u);
                }
                atomicAdd(&hist[Dc * bins + k[p]], 1u);
            }
        }
        // block-wide sum of live_here
        __shared__ int part[4];
        const int w = __reduce_add_sync(0xffffffffu, live_here);
        if ((threadIdx.x & 31) == 0) part[threadIdx.x >> 5] = w;
        __syncthreads();
        if (threadIdx.x == 0) {
            int m = part[0] + part[1] + part[2] + part[3];
            m = m > max_mult ? max_mult : m;
            atomicAdd(&hist[Dc * bins + S + m], 1u);
        }
        __syncthreads();
    }
    __syncthreads();
    for (int i = threadIdx.x; i < size; i += blockDim.x)
        if (hist[i]) atomicAdd(&counts[i], (unsigned long long)hist[i]);
}

int launch_validation_histograms(const float* x, const uint8_t* k, const uint8_t* mask, int B, int N, int Dc, int S,
                                 int bins, float lo, float hi, int max_mult, unsigned long long* counts, cudaStream_t stream) {
    const int size = Dc * bins + S + max_mult + 1;
    const int grid = B < 148 * 4 ? B : 148 * 4;
    validation_histograms_kernel<<<grid, 128, size * sizeof(unsigned int), stream>>>(
        x, k, mask, B, N, Dc, S, bins, lo, (float)bins / (hi - lo), max_mult, counts);
    return cuda_ok(cudaGetLastError(), "validation_histograms launch");
}

}  // namespace mmb
