// bridge_update.cu — the fused, HBM-bound hybrid update.
//
// One pass over the state: absorbing birth -> Euler step on (pT, eta, phi) -> telegraph jump on the
// token, all from registers.  Algorithmic traffic per particle-step (Dc=3, S=8): x r+w 24 B, v 12 B,
// logits 32 B, one jump uniform 4 B, token r+w 2 B, mask r 1 B = 75 B (84 B with the absorbing
// logit, its uniform and the mask write).  No data is reused, so there is nothing to stage in
// shared memory: each thread owns FOUR consecutive particles, which makes every global access a
// 16-byte vector (x/v: 3 x float4, logits: S x float4, u: float4) or a 4-byte word (tokens, masks),
// fully coalesced, with all loads issued before the first use.
#include "mmb_device.cuh"
#include "mmb_internal.h"

namespace mmb {

template <typename T>
__device__ __forceinline__ T ld_stream(const T* p) { return __ldcs(p); }

__device__ __forceinline__ float f4c(const float4& q, int c) { return c == 0 ? q.x : c == 1 ? q.y : c == 2 ? q.z : q.w; }

template <int S>
__global__ void __launch_bounds__(256)
bridge_update_vec4_kernel(float* __restrict__ x, uint8_t* __restrict__ k, uint8_t* __restrict__ mask,
                          const float* __restrict__ v, const float* __restrict__ logits,
                          const float* __restrict__ absorb, const float* __restrict__ uj,
                          const float* __restrict__ ua, StepScalars sc, size_t groups, int flags) {
    const size_t g = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (g >= groups) return;
    const bool do_euler = !(flags & MMB_FLAG_NO_EULER), do_jump = !(flags & MMB_FLAG_NO_JUMP),
               do_birth = (flags & MMB_FLAG_ABSORBING) != 0;

    // ---- all loads first (independent; ~300 B in flight per thread)
    float4 xr[3], vr[3], lr[S], ur = make_float4(2.f, 2.f, 2.f, 2.f), ar, uar;
    uchar4 m4 = reinterpret_cast<const uchar4*>(mask)[g];
    uchar4 k4 = make_uchar4(0, 0, 0, 0);
    if (do_euler) {
#pragma unroll
        for (int i = 0; i < 3; ++i) {
            xr[i] = reinterpret_cast<const float4*>(x)[g * 3 + i];
            vr[i] = ld_stream(reinterpret_cast<const float4*>(v) + g * 3 + i);
        }
    }
    if (do_jump) {
        k4 = reinterpret_cast<const uchar4*>(k)[g];
        ur = ld_stream(reinterpret_cast<const float4*>(uj) + g);
#pragma unroll
        for (int i = 0; i < S; ++i) lr[i] = ld_stream(reinterpret_cast<const float4*>(logits) + g * S + i);
    }
    if (do_birth) {
        ar = ld_stream(reinterpret_cast<const float4*>(absorb) + g);
        uar = ld_stream(reinterpret_cast<const float4*>(ua) + g);
    }

    // ---- birth (bridges.py:260-286)
    int m[4] = {m4.x, m4.y, m4.z, m4.w};
    if (do_birth) {
        const float a[4] = {ar.x, ar.y, ar.z, ar.w}, u[4] = {uar.x, uar.y, uar.z, uar.w};
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = absorbing_birth(m[j], a[j], u[j], sc);
        reinterpret_cast<uchar4*>(mask)[g] = make_uchar4(m[0], m[1], m[2], m[3]);
    }
    // ---- Euler (bridges.py:38-45): 12 floats = particles 0..3 x (3 features), particle j owns 3j..3j+2
    if (do_euler) {
        float xs[12] = {xr[0].x, xr[0].y, xr[0].z, xr[0].w, xr[1].x, xr[1].y, xr[1].z, xr[1].w,
                        xr[2].x, xr[2].y, xr[2].z, xr[2].w};
        const float vs[12] = {vr[0].x, vr[0].y, vr[0].z, vr[0].w, vr[1].x, vr[1].y, vr[1].z, vr[1].w,
                              vr[2].x, vr[2].y, vr[2].z, vr[2].w};
#pragma unroll
        for (int i = 0; i < 12; ++i) xs[i] = euler(xs[i], vs[i], sc.dt, (float)m[i / 3]);
#pragma unroll
        for (int i = 0; i < 3; ++i)
            reinterpret_cast<float4*>(x)[g * 3 + i] = make_float4(xs[4 * i], xs[4 * i + 1], xs[4 * i + 2], xs[4 * i + 3]);
    }
    // ---- telegraph jump (bridges.py:106-132,179-201)
    if (do_jump) {
        const int kk[4] = {k4.x, k4.y, k4.z, k4.w};
        const float u[4] = {ur.x, ur.y, ur.z, ur.w};
        int nk[4];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            float lg[S];
#pragma unroll
            for (int s = 0; s < S; ++s) lg[s] = f4c(lr[(j * S + s) >> 2], (j * S + s) & 3);
            nk[j] = telegraph_jump<S>(lg, kk[j], u[j], sc) * m[j];
        }
        reinterpret_cast<uchar4*>(k)[g] = make_uchar4(nk[0], nk[1], nk[2], nk[3]);
    }
}

// any Dc <= 8, S <= 32, any alignment: one particle per thread
__global__ void __launch_bounds__(256)
bridge_update_generic_kernel(float* __restrict__ x, uint8_t* __restrict__ k, uint8_t* __restrict__ mask,
                             const float* __restrict__ v, const float* __restrict__ logits,
                             const float* __restrict__ absorb, const float* __restrict__ uj,
                             const float* __restrict__ ua, StepScalars sc, size_t first, size_t count,
                             int Dc, int S, int flags) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= count) return;
    const size_t p = first + i;
    int m = mask[p];
    if (flags & MMB_FLAG_ABSORBING) {
        m = absorbing_birth(m, absorb[p], ua[p], sc);
        mask[p] = (uint8_t)m;
    }
    if (!(flags & MMB_FLAG_NO_EULER))
        for (int c = 0; c < Dc; ++c) x[p * Dc + c] = euler(x[p * Dc + c], v[p * Dc + c], sc.dt, (float)m);
    if (!(flags & MMB_FLAG_NO_JUMP))
        k[p] = (uint8_t)(telegraph_jump_rt(logits + p * S, S, k[p], uj[p], sc) * m);
}

static bool aligned16(const void* p) { return p == nullptr || (reinterpret_cast<uintptr_t>(p) & 15u) == 0; }

int launch_bridge_update(float* x, uint8_t* k, uint8_t* mask, const float* v, const float* logits,
                         const float* absorb, const float* uj, const float* ua, StepScalars sc,
                         size_t P, int Dc, int S, int flags, cudaStream_t stream) {
    size_t done = 0;
    const bool vec_ok = Dc == 3 && (S == 8 || S == 4 || (flags & MMB_FLAG_NO_JUMP)) && aligned16(x) && aligned16(v) &&
                        aligned16(logits) && aligned16(absorb) && aligned16(uj) && aligned16(ua) &&
                        (reinterpret_cast<uintptr_t>(k) & 3u) == 0 && (reinterpret_cast<uintptr_t>(mask) & 3u) == 0;
    if (vec_ok && P >= 4) {
        const size_t groups = P / 4;
        const unsigned grid = (unsigned)((groups + 255) / 256);
        if (S == 4 && !(flags & MMB_FLAG_NO_JUMP))
            bridge_update_vec4_kernel<4><<<grid, 256, 0, stream>>>(x, k, mask, v, logits, absorb, uj, ua, sc, groups, flags);
        else
            bridge_update_vec4_kernel<8><<<grid, 256, 0, stream>>>(x, k, mask, v, logits, absorb, uj, ua, sc, groups, flags);
        done = groups * 4;
    }
    if (done < P) {
        const size_t count = P - done;
        const unsigned grid = (unsigned)((count + 255) / 256);
        bridge_update_generic_kernel<<<grid, 256, 0, stream>>>(x, k, mask, v, logits, absorb, uj, ua, sc, done, count, Dc, S, flags);
    }
    return cuda_ok(cudaGetLastError(), "bridge_update launch");
}

// ---- uniforms exactly as the generation kernels draw them (tests) ------------------------------
__global__ void philox_uniforms_kernel(float* u, uint64_t seed, uint64_t jet_offset, int n_steps, int B, int N) {
    const size_t total = (size_t)n_steps * B * N;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int n = (int)(i % N);
        const int b = (int)((i / N) % B);
        const int s = (int)(i / ((size_t)N * B));
        u[i] = philox_uniform(seed, jet_offset + (uint64_t)b, 0, s, n);
    }
}

int launch_philox_uniforms(float* u, uint64_t seed, uint64_t jet_offset, int n_steps, int B, int N, cudaStream_t stream) {
    const size_t total = (size_t)n_steps * B * N;
    if (total == 0) return MMB_OK;
    const unsigned grid = (unsigned)((total + 255) / 256 < 148 * 8 ? (total + 255) / 256 : 148 * 8);
    philox_uniforms_kernel<<<grid, 256, 0, stream>>>(u, seed, jet_offset, n_steps, B, N);
    return cuda_ok(cudaGetLastError(), "philox_uniforms launch");
}

}  // namespace mmb
