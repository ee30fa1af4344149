// histograms.cu — validation histograms of a generated batch, accumulated on the device so that
// only a few KB of int64 counts cross NVLink (all-reduce) instead of the particle clouds.
// Observables that need no clustering (mp/data/particle_clouds/jets.py:90-107): per-particle
// histograms of the continuous features, token frequencies, particle multiplicity per jet.
#include "mmb_device.cuh"
#include "mmb_internal.h"

namespace mmb {

// counts layout: [Dc][bins] feature histograms | [S] token counts | [max_mult + 1] multiplicities
// One WARP per jet (grid-stride), no block barrier inside the pass.  Feature bins go to a histogram private to the warp (shared-
// memory atomics only collide inside one warp), token counts are ballots accumulated in a register of lane s, the multiplicity
// is a popcount; the block merges its warps once at the end and adds to the global int64 counts.  14 B per particle read.
constexpr int kHistWarps = 8;
__global__ void __launch_bounds__(kHistWarps * 32)
validation_histograms_kernel(const float* __restrict__ x, const uint8_t* __restrict__ k, const uint8_t* __restrict__ mask,
                             int B, int N, int Dc, int S, int bins, float lo, float scale, int max_mult,
                             unsigned long long* __restrict__ counts) {
    extern __shared__ unsigned int hist[];  // kHistWarps x (Dc*bins) | S + max_mult + 1
    const int nf = Dc * bins, lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    unsigned int* mine = hist + warp * nf;
    unsigned int* shared_tail = hist + kHistWarps * nf;   // [S] tokens, [max_mult + 1] multiplicities
    for (int i = threadIdx.x; i < kHistWarps * nf + S + max_mult + 1; i += blockDim.x) hist[i] = 0u;
    __syncthreads();
    unsigned int tok_count = 0;   // lane s counts token s
    for (int jet = blockIdx.x * kHistWarps + warp; jet < B; jet += gridDim.x * kHistWarps) {
        int live_here = 0;
        for (int n0 = 0; n0 < N; n0 += 32) {
            const int n = n0 + lane;
            const size_t p = (size_t)jet * N + n;
            const bool live = n < N && mask[p] != 0;
            int tok = -1;
            if (live) {
                tok = k[p];
                for (int c = 0; c < Dc; ++c) {
                    int b = (int)floorf((x[p * Dc + c] - lo) * scale);
                    b = b < 0 ? 0 : (b > bins - 1 ? bins - 1 : b);
                    atomicAdd(&mine[c * bins + b], 1u);
                }
            }
            live_here += __popc(__ballot_sync(0xffffffffu, live));
            if (S <= 32) {
                for (int s = 0; s < S; ++s) {
                    const unsigned int votes = __popc(__ballot_sync(0xffffffffu, tok == s));
                    if (lane == s) tok_count += votes;
                }
            } else if (live) {
                atomicAdd(&shared_tail[tok], 1u);
            }
        }
        if (lane == 0) atomicAdd(&shared_tail[S + (live_here > max_mult ? max_mult : live_here)], 1u);
    }
    if (S <= 32 && lane < S && tok_count) atomicAdd(&shared_tail[lane], tok_count);
    __syncthreads();
    for (int i = threadIdx.x; i < nf; i += blockDim.x) {
        unsigned int a = 0;
        for (int w = 0; w < kHistWarps; ++w) a += hist[w * nf + i];
        if (a) atomicAdd(&counts[i], (unsigned long long)a);
    }
    for (int i = threadIdx.x; i < S + max_mult + 1; i += blockDim.x)
        if (shared_tail[i]) atomicAdd(&counts[nf + i], (unsigned long long)shared_tail[i]);
}

int launch_validation_histograms(const float* x, const uint8_t* k, const uint8_t* mask, int B, int N, int Dc, int S,
                                 int bins, float lo, float hi, int max_mult, unsigned long long* counts, cudaStream_t stream) {
    const size_t smem = ((size_t)kHistWarps * Dc * bins + S + max_mult + 1) * sizeof(unsigned int);
    if (smem > 200 * 1024) return fail(MMB_EUNSUPPORTED, "validation histograms: %d x %d bins do not fit in shared memory", Dc, bins);
    if (smem > 48 * 1024)
        if (int rc = cuda_ok(cudaFuncSetAttribute(validation_histograms_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem), "histogram smem"))
            return rc;
    const int jets_per_block_pass = kHistWarps, want = (B + jets_per_block_pass - 1) / jets_per_block_pass;
    const int grid = want < 148 * 8 ? (want > 0 ? want : 1) : 148 * 8;
    validation_histograms_kernel<<<grid, kHistWarps * 32, smem, stream>>>(
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
