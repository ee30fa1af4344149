// histograms.cu — validation histograms of a generated batch, accumulated on the device so that
// only a few KB of int64 counts cross NVLink (all-reduce) instead of the particle clouds.
// Observables that need no clustering (mp/data/particle_clouds/jets.py:90-107): per-particle
// histograms of the continuous features, token frequencies, particle multiplicity per jet.
#include "mmb_device.cuh"
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

// ---- post-processing + jet observables (particles.py:85-89,124-156; utils.py:310-337; jets.py:90-107,138-141) ----------
// One warp per jet, lane l serves the slots l, l+32, ...: 28 B per particle (x r+w 24, token 1, mask 1, flavor/charge 2)
// plus one 44-byte row per jet; the seven jet sums are warp shuffles.
__device__ __forceinline__ float wsum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__global__ void __launch_bounds__(128) jet_observables_kernel(const float* __restrict__ x, const uint8_t* __restrict__ k,
                                                              const uint8_t* __restrict__ mask, float3 mean, float3 sd, int B, int N,
                                                              float* __restrict__ x_phys, int8_t* __restrict__ fc, float* __restrict__ jets) {
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    float px = 0, py = 0, pz = 0, e = 0, mult = 0, q0 = 0, q1 = 0;
    for (int n = lane; n < N; n += 32) {
        const size_t p = (size_t)b * N + n;
        const float m = mask[p] ? 1.0f : 0.0f;
        const float pt = (x[p * 3] * sd.x + mean.x) * m, eta = (x[p * 3 + 1] * sd.y + mean.y) * m, phi = (x[p * 3 + 2] * sd.z + mean.z) * m;
        const int tok = k[p];
        // tokens_to_physics: 0 photon, 1 neutral hadron, 2/3 charged hadron -/+, 4/5 electron -/+, 6/7 muon -/+
        const int flavor = m != 0.0f ? (tok < 2 ? tok : 1 + (tok >> 1)) : 0;
        const int charge = m != 0.0f ? (tok < 2 ? 0 : ((tok & 1) ? 1 : -1)) : 0;
        if (x_phys) { x_phys[p * 3] = pt; x_phys[p * 3 + 1] = eta; x_phys[p * 3 + 2] = phi; }
        if (fc) { fc[p * 2] = (int8_t)flavor; fc[p * 2 + 1] = (int8_t)charge; }
        px += pt * cosf(phi); py += pt * sinf(phi); pz += pt * sinhf(eta); e += pt * coshf(eta);
        mult += m; q0 += (float)charge; q1 += (float)charge * pt;
    }
    px = wsum(px); py = wsum(py); pz = wsum(pz); e = wsum(e); mult = wsum(mult); q0 = wsum(q0); q1 = wsum(q1);
    if (lane == 0 && jets) {
        float* o = jets + (size_t)b * MMB_JET_OBS;
        const float pt = sqrtf(fmaxf(px * px + py * py, 0.0f));
        o[MMB_JET_PX] = px; o[MMB_JET_PY] = py; o[MMB_JET_PZ] = pz; o[MMB_JET_E] = e; o[MMB_JET_PT] = pt;
        o[MMB_JET_M] = sqrtf(fmaxf(e * e - px * px - py * py - pz * pz, 0.0f));
        o[MMB_JET_ETA] = 0.5f * logf((pt + pz) / (pt - pz));
        o[MMB_JET_PHI] = atan2f(py, px);
        o[MMB_JET_MULT] = mult; o[MMB_JET_QTOTAL] = q0; o[MMB_JET_QJET] = q1 / pt;
    }
}

int launch_jet_observables(const float* x, const uint8_t* k, const uint8_t* mask, const float* mean, const float* sd, int B, int N,
                           float* x_phys, int8_t* fc, float* jets, cudaStream_t stream) {
    const float3 m = mean ? make_float3(mean[0], mean[1], mean[2]) : make_float3(0.0f, 0.0f, 0.0f);
    const float3 s = sd ? make_float3(sd[0], sd[1], sd[2]) : make_float3(1.0f, 1.0f, 1.0f);
    jet_observables_kernel<<<(B + 3) / 4, 128, 0, stream>>>(x, k, mask, m, s, B, N, x_phys, fc, jets);
    return cuda_ok(cudaGetLastError(), "jet_observables launch");
}

// ---- source state on the device (utils.py:222-286; particles.py:65-69,111-113) -------------------------------------
// One warp per jet.  The integer side (multiplicity, flavor, charge -> token) is a pure function of the Philox words and
// matches oracle/mmb_oracle.c bit for bit; the normals go through Box-Muller with the fast log / sincos.
__global__ void __launch_bounds__(128) sample_source_kernel(float* __restrict__ x, uint8_t* __restrict__ k, uint8_t* __restrict__ mask,
                                                            int B, int N, float scale, float c0, float c1, float c2, float c3,
                                                            const float* __restrict__ mult_cdf, uint64_t seed, uint64_t jet_offset) {
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const uint64_t jet = jet_offset + (uint64_t)b;
    int mult = N;
    if (mult_cdf) {
        const float u = u01(philox_block(seed, jet, 11, 0, 0).x);
        int cnt = 0;   // number of cdf entries <= u  = first index with u < cdf[i]
        for (int i = lane; i <= N; i += 32) cnt += (u < __ldg(mult_cdf + i)) ? 0 : 1;
#pragma unroll
        for (int o = 16; o; o >>= 1) cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
        mult = cnt < N ? cnt : N;
    }
    for (int n = lane; n < N; n += 32) {
        const size_t p = (size_t)b * N + n;
        const bool live = n < mult;
        float z0 = 0.0f, z1 = 0.0f, z2 = 0.0f;
        int tok = 0;
        if (live) {
            const uint4 r = philox_block(seed, jet, 9, 0, n);
            {
                const float u1 = ((float)(r.x >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(r.y >> 8) * (1.0f / 16777216.0f);
                const float rad = sqrtf(-2.0f * __logf(u1));
                float sn, cs;
                __sincosf(6.283185307179586f * u2, &sn, &cs);
                z0 = rad * cs; z1 = rad * sn;
            }
            {
                const float u1 = ((float)(r.z >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(r.w >> 8) * (1.0f / 16777216.0f);
                z2 = sqrtf(-2.0f * __logf(u1)) * __cosf(6.283185307179586f * u2);
            }
            const uint4 q = philox_block(seed, jet, 10, 0, n);
            const float uf = u01(q.x);
            const int flavor = uf < c0 ? 0 : uf < c1 ? 1 : uf < c2 ? 2 : uf < c3 ? 3 : 4;
            tok = flavor < 2 ? flavor : 2 * flavor - 2 + (int)(q.y >> 31);   // charged: token pair (-, +)
        }
        x[p * 3] = z0 * scale; x[p * 3 + 1] = z1 * scale; x[p * 3 + 2] = z2 * scale;
        k[p] = (uint8_t)tok;
        mask[p] = live ? 1 : 0;
    }
}

int launch_sample_source(float* x, uint8_t* k, uint8_t* mask, int B, int N, float scale, const float* cat_probs, const float* mult_cdf,
                         uint64_t seed, uint64_t jet_offset, cudaStream_t stream) {
    float c[4], acc = 0.0f;
    for (int i = 0; i < 4; ++i) { acc = acc + cat_probs[i]; c[i] = acc; }
    sample_source_kernel<<<(B + 3) / 4, 128, 0, stream>>>(x, k, mask, B, N, scale, c[0], c[1], c[2], c[3], mult_cdf, seed, jet_offset);
    return cuda_ok(cudaGetLastError(), "sample_source launch");
}

}  // namespace mmb
