// trans.cu — trans-dimensional jump diffusion on the GPU: TransdimensionalEPiC.forward and the JumpSampler loop.
//
// Reference (mp/ = /root/reference/multimodal_particles/):
//   network      mp/models/generative/transdimensional/transdimensional_model.py:245-426 (through EpsilonPrecond :124-133)
//   batch view   mp/models/generative/transdimensional/structure.py:226-250
//   rate         mp/models/generative/diffusion/noising.py:123-216
//   sampler      mp/models/generative/transdimensional/sampler.py:157-324, adjust_st_batch mp/data/particle_clouds/jets_dataloader.py:433-478
//
// One evaluation = a short train of kernels on the caller's stream:
//   trans_colstats / trans_tokens   tokens = argmax of the batch-axis softmax of the one-hot block (structure.py:231-232), prefix mask
//   trans_time                      EPiC time embedding + the temb_proj terms of both stacks (per jet, or per step in the sampler)
//   EPiC trunk                      epic_tc.cu / epic_fp32.cu, with the last local hidden
//   transformer stack 1             absorb_head_tc.cu (tcgen05): nearest-particle logits + x0-dimension logits (per-jet head)
//   trans_rate                      birth rate from the x0-dimension logits; nearest particle by inverse CDF on one uniform
//   transformer stack 2             vector weights + the 2S+1 statistics of the new particle
//   trans_auto                      mean / std of the particle a birth adds
//   trans_sampler_update            (sampler only) Euler-Maruyama step, centre-of-mass removal, birth — one HBM-bound pass
// >99 % of the arithmetic is in the two stacks (144.5 MFLOP per jet-evaluation).
#include <math.h>

#include <vector>

#include "mmb_device.cuh"
#include "mmb_internal.h"

namespace mmb {

struct TransHeads {
    MmbTransDims d;
    int device, sm_count;
    TfStack s1, s2;
    float* time_wT;   // device: (1 + 2 n_blocks) x { W^T [C][C], b [C] }: temb_net, stack-1 temb_proj, stack-2 temb_proj
    float* logfact;   // device: log(k!) for k < 3 R + 2
};

namespace {

constexpr int kC = 128;

// ---- tokens: batch-axis softmax (structure.py:231-232) ----------------------------------------------------------------
// Column statistics over the batch in a fixed order (the oracle's): the batch is cut into chunks of 128 jets; inside a chunk
// eight partial sums over b mod 8 (ascending b) are combined by a balanced tree; chunk sums are added in ascending order.
// Grid (column groups of 32, chunks); warp w of a block owns the jets b = w (mod 8) of its chunk.
constexpr int kChunk = 128;
__global__ void __launch_bounds__(256) trans_colmax_kernel(const float* __restrict__ onehot, int B, int NS, float* __restrict__ pmax) {
    __shared__ float s_red[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, b0 = blockIdx.y * kChunk, b1 = min(B, b0 + kChunk);
    float mx = -INFINITY;
    if (c < NS)
        for (int b = b0 + w; b < b1; b += 8) mx = fmaxf(mx, __ldg(onehot + (size_t)b * NS + c));
    s_red[w][lane] = mx;
    __syncthreads();
    if (w == 0 && c < NS) {
#pragma unroll
        for (int i = 1; i < 8; ++i) mx = fmaxf(mx, s_red[i][lane]);
        pmax[(size_t)blockIdx.y * NS + c] = mx;
    }
}
__global__ void __launch_bounds__(256) trans_colsum_kernel(const float* __restrict__ onehot, int B, int NS, const float* __restrict__ M,
                                                           float* __restrict__ psum) {
    __shared__ float s_red[8][32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const int c = blockIdx.x * 32 + lane, b0 = blockIdx.y * kChunk, b1 = min(B, b0 + kChunk);
    float mx, part = 0.0f;
    if (c < NS) {
        mx = __ldg(M + c);
        float e[kChunk / 8];   // the loads and exponentials are independent; only the adds are ordered
#pragma unroll
        for (int i = 0; i < kChunk / 8; ++i) {
            const int b = b0 + w + 8 * i;
            e[i] = b < b1 ? expf_exact_dn(__fadd_rn(__ldg(onehot + (size_t)b * NS + c), -mx)) : -1.0f;
        }
#pragma unroll
        for (int i = 0; i < kChunk / 8; ++i)
            if (e[i] >= 0.0f) part = __fadd_rn(part, e[i]);
    }
    s_red[w][lane] = part;
    __syncthreads();
    if (w == 0 && c < NS)
        psum[(size_t)blockIdx.y * NS + c] =
            __fadd_rn(__fadd_rn(__fadd_rn(s_red[0][lane], s_red[1][lane]), __fadd_rn(s_red[2][lane], s_red[3][lane])),
                      __fadd_rn(__fadd_rn(s_red[4][lane], s_red[5][lane]), __fadd_rn(s_red[6][lane], s_red[7][lane])));
}
// one warp per column: the chunk partials are fetched in parallel, then added in ascending chunk order (the oracle's order)
__global__ void __launch_bounds__(128) trans_colcombine_kernel(const float* __restrict__ part, int n_chunks, int NS, int is_max,
                                                               float* __restrict__ out) {
    const int lane = threadIdx.x & 31, c = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (c >= NS) return;
    float acc = is_max ? -INFINITY : 0.0f;
    for (int k0 = 0; k0 < n_chunks; k0 += 32) {
        const float v = k0 + lane < n_chunks ? __ldg(part + (size_t)(k0 + lane) * NS + c) : (is_max ? -INFINITY : 0.0f);
        const int cnt = min(32, n_chunks - k0);
        for (int i = 0; i < cnt; ++i) {
            const float vi = __shfl_sync(0xffffffffu, v, i);
            acc = is_max ? fmaxf(acc, vi) : __fadd_rn(acc, vi);
        }
    }
    if (lane == 0) out[c] = acc;
}

__global__ void trans_tokens_kernel(const float* __restrict__ onehot, const float* __restrict__ M, const float* __restrict__ Z,
                                    const int32_t* __restrict__ dims, int B, int N, int S, uint8_t* __restrict__ k,
                                    uint8_t* __restrict__ mask) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= (size_t)B * N) return;
    const int n = (int)(i % N), b = (int)(i / N);
    int best = 0;
    float bp = -1.0f;
    for (int s = 0; s < S; ++s) {
        const int c = n * S + s;
        const float pr = __fdiv_rn(expf_exact_dn(__fadd_rn(__ldg(onehot + i * S + s), -__ldg(M + c))), __ldg(Z + c));
        if (pr > bp) { bp = pr; best = s; }
    }
    k[i] = (uint8_t)best;
    mask[i] = n < dims[b] ? 1 : 0;
}

// ---- time terms (utils.py:183-198; gsdm.py:8-26,58; transdimensional_model.py:288-290) ---------------------------------
// A block serves kTimeJets jets: thread (o, half) owns output channel o of half of them, so every weight row is read once per
// block and the activations are read from shared memory four jets at a time.
constexpr int kTimeJets = 16;   // 8192 jets = 512 blocks (32 jets per block: 256 blocks of 8 warps left most schedulers with two warps, 144 us)
__global__ void __launch_bounds__(2 * kC) trans_time_kernel(const float* __restrict__ wT, int nblk, const float* __restrict__ ts, int B, int T,
                                                            float* __restrict__ temb_epic, float* __restrict__ tb1, float* __restrict__ tb2) {
    __shared__ __align__(16) float s_in[kC][kTimeJets];   // [channel][jet]
    constexpr int JT = kTimeJets / 2;
    const int o = threadIdx.x & (kC - 1), jh = threadIdx.x >> 7, j0 = blockIdx.x * kTimeJets;
    const int half = kC / 2;
    for (int jj = 0; jj < JT; ++jj) {
        const int j = jh * JT + jj, b = j0 + j;
        float val = 0.0f;
        if (b < B) {
            const float t = __ldg(ts + b);
            const float fe = (float)(9.210340371976184 / (double)(half - 1));   // ln(10000) / (half - 1)
            const int q = o < half ? o : o - half;
            const float a = (t * 1000.0f) * expf((float)q * -fe);
            val = o < half ? sinf(a) : cosf(a);
            if (o < T) {   // EPiC embedding: [cos(t f), sin(t f)], f_i = exp(-ln(1e4) i / (T/2))
                const int h2 = T / 2, i2 = o < h2 ? o : o - h2;
                const float f = expf(-9.210340371976184f * (float)i2 / (float)h2);
                temb_epic[(size_t)b * T + o] = (T % 2 && o == T - 1) ? 0.0f : (o < h2 ? cosf(t * f) : sinf(t * f));
            }
        }
        s_in[o][j] = val;
    }
    __syncthreads();
    const size_t mat = (size_t)kC * kC + kC;
    float acc[JT];
    auto gemv = [&](const float* W) {
        const float bias = __ldg(W + (size_t)kC * kC + o);
#pragma unroll
        for (int j = 0; j < JT; ++j) acc[j] = bias;
#pragma unroll 8
        for (int c = 0; c < kC; ++c) {
            const float w = __ldg(W + (size_t)c * kC + o);
#pragma unroll
            for (int j = 0; j < JT; j += 4) {
                const float4 m4 = *reinterpret_cast<const float4*>(&s_in[c][jh * JT + j]);
                acc[j] = fmaf(w, m4.x, acc[j]); acc[j + 1] = fmaf(w, m4.y, acc[j + 1]);
                acc[j + 2] = fmaf(w, m4.z, acc[j + 2]); acc[j + 3] = fmaf(w, m4.w, acc[j + 3]);
            }
        }
    };
    gemv(wT);   // temb = temb_net(emb); act = swish(temb)
    __syncthreads();
#pragma unroll
    for (int j = 0; j < JT; ++j) s_in[o][jh * JT + j] = acc[j] / (1.0f + expf(-acc[j]));
    __syncthreads();
    for (int m = 0; m < 2 * nblk; ++m) {
        gemv(wT + (size_t)(1 + m) * mat);
        float* out = m < nblk ? tb1 : tb2;
        const int blk = m < nblk ? m : m - nblk;
#pragma unroll
        for (int j = 0; j < JT; ++j)
            if (j0 + jh * JT + j < B) out[((size_t)(j0 + jh * JT + j) * nblk + blk) * kC + o] = acc[j];
    }
}

// ---- rate + nearest particle (noising.py:166-216; transdimensional_model.py:313-339) -----------------------------------
__device__ __forceinline__ float fr_rate(const MmbForwardRate& fr, float t) {
    return fr.kind == 1 ? fr.scalar : fr.scalar * (t > fr.rate_cut_t ? 1.0f : 0.0f) + fr.offset;
}
__device__ __forceinline__ float fr_integral(const MmbForwardRate& fr, float t) {
    return fr.kind == 1 ? fr.scalar * t : (t - fr.rate_cut_t) * fr.scalar * (t > fr.rate_cut_t ? 1.0f : 0.0f) + fr.offset * t;
}
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// one warp per jet
__global__ void __launch_bounds__(128) trans_rate_kernel(const float* __restrict__ x0_logits, const float* __restrict__ near_logits,
                                                         const int32_t* __restrict__ dims, const float* __restrict__ ts, int ts_stride,
                                                         const int32_t* __restrict__ nearest_in, const float* __restrict__ u_nearest,
                                                         MmbForwardRate fr, const float* __restrict__ logfact, int B, int N, int R, int direct,
                                                         float* __restrict__ rate, int32_t* __restrict__ nearest_out) {
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const float t = __ldg(ts + (size_t)b * ts_stride);
    const int d = dims[b];
    const float* lg = x0_logits + (size_t)b * R;
    const float I = fr_integral(fr, t), logI = logf(I);
    float mx = -INFINITY;
    if (direct) {   // rate_use_x0_pred = False: rate = softplus(rate logit) * forward_rate(t)   (transdimensional_model.py:328-332)
        const float a = lg[0];
        if (lane == 0) rate[b] = (a > 20.0f ? a : log1pf(expf(a))) * fr_rate(fr, t);   // torch softplus: beta 1, threshold 20
    }
    for (int i = d - 1 + lane; i < R; i += 32) mx = fmaxf(mx, lg[i]);
    mx = warp_max(mx);
    float z = 0.0f;
    for (int i = d - 1 + lane; i < R; i += 32) z += expf_exact(lg[i] - mx);
    z = warp_sum(z);
    float acc = 0.0f;
    for (int i = d - 1 + lane; i < R; i += 32) {
        const float prob = expf_exact(lg[i] - mx) / z;
        float ratio;
        if (d > 1) {
            ratio = fmaxf((1.0f / I) * (float)((i + 1) - d), 0.0f);
        } else {   // x_t has one particle: ratio of Poisson probabilities with the truncated normaliser (noising.py:199-212)
            auto logp = [&](int k) { return (k == 0 ? 0.0f : (float)k * logI) - I - __ldg(logfact + k); };
            // the Poisson pmf is log-concave with its mode at floor(I): the window maximum sits at the mode clamped into the
            // window, and past the mode the terms only fall, so the sum stops once they drop below e^-25 of it
            const int mode = min(max((int)I, i), i + 2 * R - 1);
            const float m2 = logp(mode);
            float se = 0.0f;
            for (int j = 0; j < 2 * R; ++j) {
                const float a = logp(i + j) - m2;
                se += expf(a);
                if (i + j > mode && a < -25.0f) break;
            }
            const float dim1 = m2 + logf(se);
            const float dim2 = i == 0 ? -1000.0f : logp(i - 1);
            ratio = expf(dim2 - dim1);
        }
        acc += ratio * prob;
    }
    acc = warp_sum(acc);
    if (lane == 0 && !direct) rate[b] = fr_rate(fr, t) * acc;
    // nearest particle: given, or multinomial(softmax(near_atom_logits)) over ALL N slots by inverse CDF — the scan is
    // sequential in fp32 so that the chosen index is bit-identical to the oracle's
    if (lane == 0 && nearest_out) {
        int near;
        if (nearest_in) near = nearest_in[b];
        else {
            const float* nl = near_logits + (size_t)b * N;
            float m = -INFINITY, zz = 0.0f, c = 0.0f;
            for (int n = 0; n < N; ++n) m = nl[n] > m ? nl[n] : m;
            for (int n = 0; n < N; ++n) zz = __fadd_rn(zz, expf_exact(__fadd_rn(nl[n], -m)));
            const float u = u_nearest[b];
            near = N - 1;
            for (int n = 0; n < N; ++n) {
                c = __fadd_rn(c, __fdiv_rn(expf_exact(__fadd_rn(nl[n], -m)), zz));
                if (u < c) { near = n; break; }
            }
        }
        nearest_out[b] = near;
    }
}

// ---- mean / std of the particle a birth adds (transdimensional_model.py:369-424); one warp per jet ------------------------
__global__ void __launch_bounds__(128) trans_auto_kernel(const float* __restrict__ x, const uint8_t* __restrict__ mask,
                                                         const int32_t* __restrict__ nearest, const float* __restrict__ vec_w,
                                                         const float* __restrict__ post_auto, const int32_t* __restrict__ dims,
                                                         int B, int N, int S, float* __restrict__ new_mean, float* __restrict__ new_std,
                                                         float* __restrict__ auto_mean, float* __restrict__ auto_std) {
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int F = 3 + S, near = nearest[b];
    const float* xb = x + (size_t)b * N * 3;
    const float xa0 = xb[near * 3], xa1 = xb[near * 3 + 1], xa2 = xb[near * 3 + 2];
    float p0 = 0.0f, p1 = 0.0f, p2 = 0.0f;
    for (int n = lane; n < N; n += 32) {
        const float m = mask[(size_t)b * N + n] ? 1.0f : 0.0f, w = vec_w[(size_t)b * N + n];
        const float d0 = (xa0 - xb[n * 3]) * m, d1 = (xa1 - xb[n * 3 + 1]) * m, d2 = (xa2 - xb[n * 3 + 2]) * m;
        const float inv = 1.0f / (sqrtf((d0 * d0 + d1 * d1) + d2 * d2) + 1e-3f);
        p0 = fmaf(w, d0 * inv, p0); p1 = fmaf(w, d1 * inv, p1); p2 = fmaf(w, d2 * inv, p2);
    }
    p0 = xa0 + warp_sum(p0); p1 = xa1 + warp_sum(p1); p2 = xa2 + warp_sum(p2);
    const float* pa = post_auto + (size_t)b * (2 * S + 1);
    const int slot = dims[b];
    for (int i = lane; i < F; i += 32) {
        const float mean = i == 0 ? p0 : i == 1 ? p1 : i == 2 ? p2 : pa[1 + (i - 3)];
        const float sd = i < 3 ? pa[0] : pa[1 + S + (i - 3)];
        if (new_mean) { new_mean[(size_t)b * F + i] = mean; new_std[(size_t)b * F + i] = sd; }
        if (auto_mean && slot < N) {   // get_next_dim_added_mask (structure.py:175-184): only the slot a birth fills is non-zero
            const size_t at = (size_t)b * N * F + (i < 3 ? (size_t)slot * 3 + i : (size_t)N * 3 + (size_t)slot * S + (i - 3));
            auto_mean[at] = mean;
            auto_std[at] = sd;
        }
    }
}

// ---- one sampler update (sampler.py:221-255 + adjust_st_batch) -----------------------------------------------------------
// Philox streams of the sampler: 2,3,4 = diffusion noise of a particle (12 normals, 11 used); 5 = per-jet uniforms
// (word 0: nearest particle, word 1: birth); 6,7,8 = the new particle's 11 normals.
__device__ __forceinline__ void box_muller(uint32_t a, uint32_t b, float& n0, float& n1) {
    const float u1 = ((float)(a >> 8) + 1.0f) * (1.0f / 16777216.0f);   // (0, 1]
    const float u2 = (float)(b >> 8) * (1.0f / 16777216.0f);
    const float r = sqrtf(-2.0f * __logf(u1));
    float s, c;
    __sincosf(6.283185307179586f * u2, &s, &c);
    n0 = r * c;
    n1 = r * s;
}
__device__ __forceinline__ void philox_normals12(uint64_t seed, uint64_t jet, int stream0, int step, int idx, float (&z)[12]) {
#pragma unroll
    for (int q = 0; q < 3; ++q) {
        const uint4 r = philox_block(seed, jet, stream0 + q, step, idx);
        box_muller(r.x, r.y, z[4 * q], z[4 * q + 1]);
        box_muller(r.z, r.w, z[4 * q + 2], z[4 * q + 3]);
    }
}
__device__ __forceinline__ float nan_to_num(float a) {   // torch.nan_to_num: NaN -> 0, +-inf -> +-FLT_MAX; finite values take one compare
    return fabsf(a) <= 3.4028234664e38f ? a : (a != a ? 0.0f : copysignf(3.4028234664e38f, a));
}
__device__ __forceinline__ float softplus(float a) { return a > 20.0f ? a : log1pf(expf(a)); }

// The one-hot block of the update is purely element-wise (sampler.py:221-231 on the one-hot slots): one thread per particle slot
// over a flat grid, 96 B per live slot (one-hot r+w 64, logits 32), two Philox blocks for its eight normals.  It runs BEFORE the
// continuous kernel below (which updates dims and writes the one-hot row of a newborn particle into slot `dims`).
template <int S, bool CORR>   // CORR: corrector row (device coefficients, predictor-step mask)
__global__ void __launch_bounds__(256) trans_sampler_onehot_kernel(float* __restrict__ onehot, const int32_t* __restrict__ dims,
                                                                   const float* __restrict__ logits, float c_decay, float cs, float c_noise,
                                                                   const float* __restrict__ coef, const int32_t* __restrict__ mask_dims,
                                                                   const float* __restrict__ z_diff, uint64_t seed, uint64_t jet_offset,
                                                                   int step, int B, int N) {
    constexpr int F = 3 + S;
    const size_t pi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (pi >= (size_t)B * N) return;
    const int b = (int)(pi / N), n = (int)(pi % N);
    if (n >= __ldg(dims + b)) return;
    if constexpr (CORR) {
        if (n >= __ldg(mask_dims + b)) return;   // corrector rows: the predictor step's mask (sampler.py:219, 277-279)
        c_decay = __ldg(coef); cs = __ldg(coef + 1); c_noise = __ldg(coef + 2);
    }
    const bool noisy = c_noise != 0.0f;
    float oh[S], lg[S], z[8];
    if constexpr (S % 4 == 0) {
#pragma unroll
        for (int s2 = 0; s2 < S; s2 += 4) {
            const float4 a = *reinterpret_cast<const float4*>(onehot + pi * S + s2);
            const float4 l = __ldg(reinterpret_cast<const float4*>(logits + pi * S + s2));
            oh[s2] = a.x; oh[s2 + 1] = a.y; oh[s2 + 2] = a.z; oh[s2 + 3] = a.w;
            lg[s2] = l.x; lg[s2 + 1] = l.y; lg[s2 + 2] = l.z; lg[s2 + 3] = l.w;
        }
    } else {
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2) { oh[s2] = onehot[pi * S + s2]; lg[s2] = __ldg(logits + pi * S + s2); }
    }
    if (noisy) {
        if (z_diff) {
            const float* zb = z_diff + (size_t)b * N * F;
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) z[s2] = __ldg(zb + (size_t)N * 3 + n * S + s2);
        } else {
            const uint64_t jet = jet_offset + (uint64_t)b;
            const uint4 r0 = philox_block(seed, jet, 3, step, n);
            box_muller(r0.x, r0.y, z[0], z[1]);
            box_muller(r0.z, r0.w, z[2], z[3]);
            if constexpr (S > 4) {
                const uint4 r1 = philox_block(seed, jet, 4, step, n);
                box_muller(r1.x, r1.y, z[4], z[5]);
                box_muller(r1.z, r1.w, z[6], z[7]);
            }
        }
    }
#pragma unroll
    for (int s2 = 0; s2 < S; ++s2) {
        float a = c_decay * oh[s2] + cs * lg[s2];
        if (noisy) a = a + c_noise * z[s2];
        oh[s2] = nan_to_num(a);
    }
    if constexpr (S % 4 == 0) {
#pragma unroll
        for (int s2 = 0; s2 < S; s2 += 4) *reinterpret_cast<float4*>(onehot + pi * S + s2) = make_float4(oh[s2], oh[s2 + 1], oh[s2 + 2], oh[s2 + 3]);
    } else {
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2) onehot[pi * S + s2] = oh[s2];
    }
}

// Continuous block, one WARP per jet: lane l serves the particle slots l, l+32, l+64, l+96, so the three centre-of-mass reductions
// are warp shuffles and no block-wide barrier sits in the pass; four jets per 128-thread block.  (The one-hot block is the kernel above.)  Dead slots (n >= dims) hold zeros —
// the sampler's invariant, asserted by the reference in adjust_st_batch (jets_dataloader.py:451-452) — and stay zero under the
// update, so they are neither read nor written: the pass moves 133 B per LIVE particle-step.
template <int S, int JUMP_MODE>
__global__ void __launch_bounds__(128) trans_sampler_update_kernel(float* __restrict__ x, float* __restrict__ onehot, int32_t* __restrict__ dims,
                                                                   const float* __restrict__ v, const float* __restrict__ logits,
                                                                   const float* __restrict__ rate, const float* __restrict__ new_mean,
                                                                   const float* __restrict__ new_std, float c_decay, float c_score,
                                                                   float c_noise, float inv_std, float jump_dt,
                                                                   const float* __restrict__ coef, const int32_t* __restrict__ mask_dims,
                                                                   float death_prob,
                                                                   const float* __restrict__ z_diff, const float* __restrict__ u_jump,
                                                                   const float* __restrict__ u_death,
                                                                   const float* __restrict__ z_new, uint64_t seed, uint64_t jet_offset,
                                                                   int step, int B, int N) {
    constexpr int jump_mode = JUMP_MODE;
    constexpr bool CORR = JUMP_MODE != 1;   // corrector rows take device coefficients and the predictor step's mask
    // jump_mode: 1 = predictor row (birth); 0 = corrector row without jumps (one centre-of-mass removal, sampler.py:280-282);
    //            2 = corrector row with the jump corrector (birth and death, sampler.py:285-312)
    constexpr int F = 3 + S, SL = 4;   // up to 128 slots = 4 per lane
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int dim = dims[b];
    const int upd = CORR ? min(__ldg(mask_dims + b), dim) : dim;   // slots that take the increment (stale mask of a corrector row)
    const uint64_t jet = jet_offset + (uint64_t)b;
    bool born = false, dies = false;
    if (jump_mode != 0) {
        const bool drawn = !u_jump || (jump_mode == 2 && !u_death);
        const uint4 ur = drawn ? philox_block(seed, jet, 5, step, 0) : make_uint4(0, 0, 0, 0);
        const float uj = u_jump ? __ldg(u_jump + b) : u01(ur.y);
        born = (uj < __ldg(rate + b) * jump_dt) && dim < N;
        if (jump_mode == 2) {
            const float ud = u_death ? __ldg(u_death + b) : u01(ur.z);
            dies = (ud < death_prob) && dim > 1;
            if (born && dies) born = dies = false;   // the newborn is the particle delete_dims removes: the state is as before
        }
    }
    const int new_dim = born ? dim + 1 : (dies ? dim - 1 : dim);
    const float* zb = z_diff ? z_diff + (size_t)b * N * F : nullptr;
    float cs = -(c_score * inv_std);   // c_score * -(inv_std * D): one rounding apart from the reference order
    if constexpr (CORR) { c_decay = __ldg(coef); cs = __ldg(coef + 1); c_noise = __ldg(coef + 2); }
    const bool noisy = c_noise != 0.0f;
    // ---- continuous block: Euler-Maruyama with the noise centred over the live particles, then centre-of-mass removal,
    //      the birth, and a second centre-of-mass removal (sampler.py:221-255, jets_dataloader.py:433-478)
    float xs[SL][3], zc[SL][3];
#pragma unroll
    for (int q = 0; q < SL; ++q) {
        const int n = lane + 32 * q;
        const size_t pi = (size_t)b * N + n;
#pragma unroll
        for (int c = 0; c < 3; ++c) { xs[q][c] = 0.0f; zc[q][c] = 0.0f; }
        if (n < dim) {
            float vs[3];
#pragma unroll
            for (int c = 0; c < 3; ++c) { xs[q][c] = x[pi * 3 + c]; vs[c] = __ldg(v + pi * 3 + c); }
            if (noisy) {
                if (zb) {
#pragma unroll
                    for (int c = 0; c < 3; ++c) zc[q][c] = __ldg(zb + n * 3 + c);
                } else {
                    const uint4 r0 = philox_block(seed, jet, 2, step, n);
                    float spare;
                    box_muller(r0.x, r0.y, zc[q][0], zc[q][1]);
                    box_muller(r0.z, r0.w, zc[q][2], spare);
                }
            }
            if (n < upd) {
#pragma unroll
                for (int c = 0; c < 3; ++c) xs[q][c] = c_decay * xs[q][c] + cs * vs[c];
            }
        }
    }
    const float inv_dim = 1.0f / (float)dim;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        if (noisy) {
            const float zm = warp_sum((zc[0][c] + zc[1][c]) + (zc[2][c] + zc[3][c])) * inv_dim;
#pragma unroll
            for (int q = 0; q < SL; ++q)
                if (lane + 32 * q < upd) xs[q][c] = xs[q][c] + c_noise * (zc[q][c] - zm);
        }
#pragma unroll
        for (int q = 0; q < SL; ++q) xs[q][c] = nan_to_num(xs[q][c]);
        const float xm = warp_sum((xs[0][c] + xs[1][c]) + (xs[2][c] + xs[3][c])) * inv_dim;
#pragma unroll
        for (int q = 0; q < SL; ++q)
            if (lane + 32 * q < dim) xs[q][c] -= xm;
    }
    if (born && lane == (dim & 31)) {   // the new particle takes slot `dim` (sampler.py:238-255)
        float zn[12];
        if (z_new) {
#pragma unroll
            for (int i = 0; i < F; ++i) zn[i] = __ldg(z_new + (size_t)b * F + i);
        } else {
            philox_normals12(seed, jet, 6, step, 0, zn);
        }
        const float* nm = new_mean + (size_t)b * F;
        const float* ns = new_std + (size_t)b * F;
        float nx[3], no[S];
#pragma unroll
        for (int c = 0; c < 3; ++c) nx[c] = nan_to_num(__ldg(nm + c) + zn[c] * softplus(__ldg(ns + c)));
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2) no[s2] = nan_to_num(__ldg(nm + 3 + s2) + zn[3 + s2] * softplus(__ldg(ns + 3 + s2)));
#pragma unroll
        for (int q = 0; q < SL; ++q)
            if (q == (dim >> 5)) { xs[q][0] = nx[0]; xs[q][1] = nx[1]; xs[q][2] = nx[2]; }
        const size_t pi = (size_t)b * N + dim;
#pragma unroll
        for (int s2 = 0; s2 < S; ++s2) onehot[pi * S + s2] = no[s2];
    }
    if (dies) {   // delete_dims removes the last live particle (sampler.py:305-308); its slot goes back to zero
        if (lane == (new_dim & 31)) {
#pragma unroll
            for (int q = 0; q < SL; ++q)
                if (q == (new_dim >> 5)) { xs[q][0] = 0.0f; xs[q][1] = 0.0f; xs[q][2] = 0.0f; }
            const size_t pi = (size_t)b * N + new_dim;
#pragma unroll
            for (int c = 0; c < 3; ++c) x[pi * 3 + c] = 0.0f;
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) onehot[pi * S + s2] = 0.0f;
        }
    }
    const float inv_new = 1.0f / (float)new_dim;
    if (jump_mode != 0)
#pragma unroll
    for (int c = 0; c < 3; ++c) {   // second adjust_st_batch with the new multiplicity (it runs whether or not a particle was born)
        const float xm = warp_sum((xs[0][c] + xs[1][c]) + (xs[2][c] + xs[3][c])) * inv_new;
#pragma unroll
        for (int q = 0; q < SL; ++q)
            if (lane + 32 * q < new_dim) xs[q][c] -= xm;
    }
#pragma unroll
    for (int q = 0; q < SL; ++q) {
        const int n = lane + 32 * q;
        if (n < new_dim) {
            const size_t pi = (size_t)b * N + n;
#pragma unroll
            for (int c = 0; c < 3; ++c) x[pi * 3 + c] = xs[q][c];
        }
    }
    if (lane == 0 && new_dim != dim) dims[b] = new_dim;
}

// ---- Langevin corrector step size (sampler.py:264-274): grad_norm = mean_b |score_b|, noise_norm = mean_b |noise_b| over the
// whole batch, step = (snr noise_norm / grad_norm)^2 2 alpha.  One warp per jet writes its two norms (the noise is the same
// Philox draw the update kernels regenerate, centred over the live particles); one block reduces them in a fixed order and
// leaves coef = {1, -step inv_std, sqrt(2 step) or 0} for the update kernels.
template <int S>
__global__ void __launch_bounds__(128) trans_corrector_norms_kernel(const int32_t* __restrict__ dims, const float* __restrict__ v,
                                                                    const float* __restrict__ logits, float inv_std,
                                                                    const float* __restrict__ z_diff, uint64_t seed, uint64_t jet_offset,
                                                                    int step, int B, int N, float* __restrict__ norm_s,
                                                                    float* __restrict__ norm_n) {
    constexpr int F = 3 + S, SL = 4;
    const int lane = threadIdx.x & 31, b = blockIdx.x * 4 + (threadIdx.x >> 5);
    if (b >= B) return;
    const int dim = dims[b];
    const uint64_t jet = jet_offset + (uint64_t)b;
    const float* zb = z_diff ? z_diff + (size_t)b * N * F : nullptr;
    float zc[SL][3], g2 = 0.0f, n2 = 0.0f;
#pragma unroll
    for (int q = 0; q < SL; ++q) {
        const int n = lane + 32 * q;
        const size_t pi = (size_t)b * N + n;
#pragma unroll
        for (int c = 0; c < 3; ++c) zc[q][c] = 0.0f;
        if (n < dim) {
            float zo[8];
            if (zb) {
#pragma unroll
                for (int c = 0; c < 3; ++c) zc[q][c] = __ldg(zb + n * 3 + c);
#pragma unroll
                for (int s2 = 0; s2 < S; ++s2) zo[s2] = __ldg(zb + (size_t)N * 3 + n * S + s2);
            } else {
                const uint4 r0 = philox_block(seed, jet, 2, step, n);
                float spare;
                box_muller(r0.x, r0.y, zc[q][0], zc[q][1]);
                box_muller(r0.z, r0.w, zc[q][2], spare);
                const uint4 r1 = philox_block(seed, jet, 3, step, n);
                box_muller(r1.x, r1.y, zo[0], zo[1]);
                box_muller(r1.z, r1.w, zo[2], zo[3]);
                if constexpr (S > 4) {
                    const uint4 r2 = philox_block(seed, jet, 4, step, n);
                    box_muller(r2.x, r2.y, zo[4], zo[5]);
                    box_muller(r2.z, r2.w, zo[6], zo[7]);
                }
            }
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) n2 += zo[s2] * zo[s2];
#pragma unroll
            for (int c = 0; c < 3; ++c) { const float sc = inv_std * __ldg(v + pi * 3 + c); g2 += sc * sc; }
#pragma unroll
            for (int s2 = 0; s2 < S; ++s2) { const float sc = inv_std * __ldg(logits + pi * S + s2); g2 += sc * sc; }
        }
    }
    const float inv_dim = 1.0f / (float)dim;
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        const float zm = warp_sum((zc[0][c] + zc[1][c]) + (zc[2][c] + zc[3][c])) * inv_dim;
#pragma unroll
        for (int q = 0; q < SL; ++q)
            if (lane + 32 * q < dim) { const float a = zc[q][c] - zm; n2 += a * a; }
    }
    g2 = warp_sum(g2);
    n2 = warp_sum(n2);
    if (lane == 0) { norm_s[b] = sqrtf(g2); norm_n[b] = sqrtf(n2); }
}

__global__ void __launch_bounds__(256) trans_corrector_coef_kernel(const float* __restrict__ norm_s, const float* __restrict__ norm_n, int B,
                                                                   float snr, float alpha, float inv_std, int noise_on,
                                                                   float* __restrict__ coef) {
    __shared__ double sg[256], sn[256];
    double g = 0.0, n = 0.0;
    for (int b = threadIdx.x; b < B; b += 256) { g += (double)norm_s[b]; n += (double)norm_n[b]; }
    sg[threadIdx.x] = g;
    sn[threadIdx.x] = n;
    __syncthreads();
    for (int w = 128; w > 0; w >>= 1) {
        if (threadIdx.x < w) { sg[threadIdx.x] += sg[threadIdx.x + w]; sn[threadIdx.x] += sn[threadIdx.x + w]; }
        __syncthreads();
    }
    if (threadIdx.x == 0) {
        const float grad_norm = (float)(sg[0] / B), noise_norm = (float)(sn[0] / B);
        const float r = snr * noise_norm / grad_norm;
        const float step = r * r * 2.0f * alpha;
        coef[0] = 1.0f;
        coef[1] = -(step * inv_std);
        coef[2] = noise_on ? sqrtf(2.0f * step) : 0.0f;
        coef[3] = step;
    }
}

// broadcast ts[step] for the sampler's per-jet uniform of the nearest particle when drawn in-kernel
__global__ void trans_philox_u_near_kernel(float* __restrict__ u, uint64_t seed, uint64_t jet_offset, int step, int B) {
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < B) u[b] = u01(philox_block(seed, jet_offset + (uint64_t)b, 5, step, 0).x);
}

inline size_t align64(size_t n) { return (n + 63) & ~(size_t)63; }

// workspace carve-up shared by the forward and the sampler (in floats, every region 256-byte aligned)
struct TransWs {
    size_t k, mask, M, Z, pmax, psum, means, temb, tb1, tb2, v, logits, hidden, near_logits, vec_w, x0_logits, post_auto, nearest, new_mean, new_std, rate,
        u_near, ts, norm_s, norm_n, coef, mask_dims, pack, pack_ints, total;
    TransWs(const MmbEpicDims& e, const MmbTransDims& d, int B, int N, int n_time) {
        const size_t P = (size_t)B * N, F = 3 + d.vocab_size;
        size_t at = 0;
        auto take = [&](size_t floats) { const size_t o = at; at += align64(floats); return o; };
        k = take((P + 3) / 4); mask = take((P + 3) / 4);
        M = take((size_t)N * d.vocab_size); Z = take((size_t)N * d.vocab_size);
        const size_t chunks = ((size_t)B + kChunk - 1) / kChunk;
        pmax = take(chunks * N * d.vocab_size); psum = take(chunks * N * d.vocab_size);
        means = take((size_t)B * kC);
        temb = take((size_t)n_time * e.dim_time_emb);
        tb1 = take((size_t)n_time * d.n_blocks * kC); tb2 = take((size_t)n_time * d.n_blocks * kC);
        v = take(P * 3); logits = take(P * d.vocab_size); hidden = take(P * e.dim_hidden_local);
        near_logits = take(P); vec_w = take(P); x0_logits = take((size_t)B * d.max_particles);
        post_auto = take((size_t)B * (2 * d.vocab_size + 1)); nearest = take(B); new_mean = take((size_t)B * F);
        new_std = take((size_t)B * F); rate = take(B); u_near = take(B); ts = take(n_time);
        norm_s = take(B); norm_n = take(B); coef = take(4); mask_dims = take(B);
        pack_ints = tf_pack_scratch_ints(B); pack = take(pack_ints);   // jet packing of the transformer stacks (int32)
        total = at;
    }
};

}  // namespace

// ---- host -------------------------------------------------------------------------------------------------------------
// outputs of post_rate_proj: max_num_particles x0-dimension logits, or one rate logit with encoder.rate_use_x0_pred = False
// (transdimensional_model.py:185-188)
static int rate_dim(const MmbTransDims& d) { return d.rate_direct ? 1 : d.max_particles; }

static size_t trans_floats(const MmbTransDims& d) {
    const size_t C = d.transformer_dim, lin = C * C + C, H = d.hidden, S = d.vocab_size, R = rate_dim(d);
    const size_t block = 6 * C + 6 * lin;
    return lin + 2 * (size_t)d.n_blocks * lin + (C * (H + S) + C) + d.n_blocks * block + lin + (R * C + R) + (C + 1) +
           (C * (H + S + 3) + C) + d.n_blocks * block + (C + 1) + lin + ((2 * S + 1) * C + (2 * S + 1));
}

static bool trans_dims_ok(const MmbTransDims& d) {
    return d.transformer_dim == kC && d.n_heads == 2 && d.n_blocks >= 1 && d.n_blocks <= 4 && d.hidden >= 1 && d.vocab_size >= 1 &&
           d.hidden + d.vocab_size + 3 <= 32 && d.max_particles >= 1 && d.max_particles <= 128 && 2 * d.vocab_size + 1 <= 128;
}

static void trans_destroy(TransHeads* h) {
    if (!h) return;
    tf_stack_free(&h->s1);
    tf_stack_free(&h->s2);
    if (h->time_wT) cudaFree(h->time_wT);
    if (h->logfact) cudaFree(h->logfact);
    delete h;
}

static int trans_create(const MmbTransDims* dims, const float* W, size_t n_floats, int device, TransHeads** out) {
    const MmbTransDims d = *dims;
    if (!trans_dims_ok(d))
        return fail(MMB_EUNSUPPORTED, "trans heads are built for transformer_dim=128, n_heads=2, 1..4 blocks, hidden+vocab+3<=32, <=128 particles");
    if (n_floats != trans_floats(d)) return fail(MMB_EINVAL, "trans heads blob has %zu floats, layout wants %zu", n_floats, trans_floats(d));
    const int C = kC, H = d.hidden, S = d.vocab_size, R = rate_dim(d), nb = d.n_blocks, PA = 2 * S + 1;
    const size_t lin = (size_t)C * C + C, block = 6 * (size_t)C + 6 * lin;
    const float* temb_net = W;
    const float* s1 = W + lin + 2 * (size_t)nb * lin;
    const float* s1_blocks = s1 + (size_t)C * (H + S) + C;
    const float* pre_rate = s1_blocks + (size_t)nb * block;
    const float* post_rate = pre_rate + lin;
    const float* near_w = post_rate + (size_t)R * C + R;
    const float* s2 = near_w + C + 1;
    const float* s2_blocks = s2 + (size_t)C * (H + S + 3) + C;
    const float* vecw = s2_blocks + (size_t)nb * block;
    const float* pre_auto = vecw + C + 1;
    const float* post_auto = pre_auto + lin;
    // fold mean -> Linear(C->C) -> Linear(C->n): W = post pre, b = post pre_b + post_b (the mean commutes with the first Linear)
    auto fold = [&](const float* pre, const float* post, int n_out, std::vector<float>& Wf, std::vector<float>& bf) {
        Wf.assign((size_t)n_out * C, 0.0f);
        bf.assign(n_out, 0.0f);
        for (int o = 0; o < n_out; ++o) {
            for (int c = 0; c < C; ++c) {
                double acc = 0;
                for (int j = 0; j < C; ++j) acc += (double)post[(size_t)o * C + j] * pre[(size_t)j * C + c];
                Wf[(size_t)o * C + c] = (float)acc;
            }
            double acc = post[(size_t)n_out * C + o];
            for (int j = 0; j < C; ++j) acc += (double)post[(size_t)o * C + j] * pre[(size_t)C * C + j];
            bf[o] = (float)acc;
        }
    };
    std::vector<float> w1, b1, w2, b2;
    fold(pre_rate, post_rate, R, w1, b1);
    fold(pre_auto, post_auto, PA, w2, b2);
    // time matrices, transposed for coalesced reads
    const int n_mat = 1 + 2 * nb;
    std::vector<float> tw((size_t)n_mat * lin);
    for (int m = 0; m < n_mat; ++m) {
        const float* src = m == 0 ? temb_net : W + lin + (size_t)(m - 1) * lin;
        float* dst = tw.data() + (size_t)m * lin;
        for (int o = 0; o < C; ++o)
            for (int c = 0; c < C; ++c) dst[(size_t)c * C + o] = src[(size_t)o * C + c];
        for (int o = 0; o < C; ++o) dst[(size_t)C * C + o] = src[(size_t)C * C + o];
    }
    std::vector<float> lf((size_t)3 * d.max_particles + 2);
    for (size_t k = 0; k < lf.size(); ++k) lf[k] = (float)lgamma((double)k + 1.0);

    int prev = 0;
    if (int rc = cuda_ok(cudaGetDevice(&prev), "cudaGetDevice")) return rc;
    if (int rc = cuda_ok(cudaSetDevice(device), "cudaSetDevice")) return rc;
    TransHeads* h = new TransHeads{d, device, 148, TfStack{}, TfStack{}, nullptr, nullptr};
    int rc = cuda_ok(cudaDeviceGetAttribute(&h->sm_count, cudaDevAttrMultiProcessorCount, device), "sm count");
    if (!rc) rc = tf_stack_build(&h->s1, s1, H + S, s1_blocks, nb, near_w, near_w[C], w1.data(), b1.data(), R);
    if (!rc) rc = tf_stack_build(&h->s2, s2, H + S + 3, s2_blocks, nb, vecw, vecw[C], w2.data(), b2.data(), PA);
    if (!rc) rc = cuda_ok(cudaMalloc(&h->time_wT, tw.size() * 4), "cudaMalloc time weights");
    if (!rc) rc = cuda_ok(cudaMemcpy(h->time_wT, tw.data(), tw.size() * 4, cudaMemcpyHostToDevice), "time weights upload");
    if (!rc) rc = cuda_ok(cudaMalloc(&h->logfact, lf.size() * 4), "cudaMalloc log-factorials");
    if (!rc) rc = cuda_ok(cudaMemcpy(h->logfact, lf.data(), lf.size() * 4, cudaMemcpyHostToDevice), "log-factorials upload");
    cudaSetDevice(prev);
    if (rc) { trans_destroy(h); return rc; }
    *out = h;
    return MMB_OK;
}

static int check_pair(const EpicModel* m, const TransHeads* h, int N) {
    if (m->dims.dim_hidden_local != h->d.hidden || m->dims.vocab_size != h->d.vocab_size || m->dims.dim_continuous != 3)
        return fail(MMB_EINVAL, "trunk (H=%d, S=%d, Dc=%d) does not match the trans heads (H=%d, S=%d, Dc=3)", m->dims.dim_hidden_local,
                    m->dims.vocab_size, m->dims.dim_continuous, h->d.hidden, h->d.vocab_size);
    if (m->dims.disc_head_hidden != 0) return fail(MMB_EINVAL, "the trans trunk is created without a discrete head (fc_layer is never applied)");
    if (m->dims.dim_context != 0) return fail(MMB_EUNSUPPORTED, "the trans trunk takes no context features (its batches carry none: return_type 'list')");
    if (N < 1 || N > h->d.max_particles) return fail(MMB_EINVAL, "N=%d outside 1..max_particles=%d", N, h->d.max_particles);
    if (m->dims.dim_time_emb > kC) return fail(MMB_EUNSUPPORTED, "time embedding wider than 128");
    return MMB_OK;
}

// everything of one evaluation after the time terms; ts / tb pointers with stride 0 = one time for all jets
static int trans_eval(const EpicModel* m, const TransHeads* h, const float* x, const float* onehot, const int32_t* dims,
                      const float* ts, int ts_stride, const float* temb, const float* tb1, const float* tb2, int time_stride,
                      const int32_t* nearest_in, const float* u_nearest, const MmbForwardRate& fr, int B, int N, float* ws,
                      const TransWs& L, float* x0_logits_out, float* near_logits_out, float* auto_mean, float* auto_std,
                      int precision, cudaStream_t s) {
    const int S = h->d.vocab_size, H = h->d.hidden, R = rate_dim(h->d), nb = h->d.n_blocks, T = m->dims.dim_time_emb;
    const bool direct = h->d.rate_direct != 0;
    const size_t P = (size_t)B * N;
    uint8_t* k = reinterpret_cast<uint8_t*>(ws + L.k);
    uint8_t* mask = reinterpret_cast<uint8_t*>(ws + L.mask);
    int32_t* nearest = reinterpret_cast<int32_t*>(ws + L.nearest);
    // rate_use_x0_pred = False: the head's single output stays in the workspace and the caller's x0_dim_logits are zeros
    // (transdimensional_model.py:326-327)
    float* x0l = (x0_logits_out && !direct) ? x0_logits_out : ws + L.x0_logits;
    if (x0_logits_out && direct)
        if (int rc = cuda_ok(cudaMemsetAsync(x0_logits_out, 0, (size_t)B * h->d.max_particles * sizeof(float), s), "x0 logits reset")) return rc;
    float* nl = near_logits_out ? near_logits_out : ws + L.near_logits;
    {
        const int NS = N * S, chunks = (B + kChunk - 1) / kChunk;
        const dim3 grid((NS + 31) / 32, chunks);
        trans_colmax_kernel<<<grid, 256, 0, s>>>(onehot, B, NS, ws + L.pmax);
        trans_colcombine_kernel<<<(NS + 3) / 4, 128, 0, s>>>(ws + L.pmax, chunks, NS, 1, ws + L.M);
        trans_colsum_kernel<<<grid, 256, 0, s>>>(onehot, B, NS, ws + L.M, ws + L.psum);
        trans_colcombine_kernel<<<(NS + 3) / 4, 128, 0, s>>>(ws + L.psum, chunks, NS, 0, ws + L.Z);
    }
    trans_tokens_kernel<<<(unsigned)((P + 255) / 256), 256, 0, s>>>(onehot, ws + L.M, ws + L.Z, dims, B, N, S, k, mask);
    if (int rc = cuda_ok(cudaGetLastError(), "trans tokens launch")) return rc;
    int rc = mmb_epic_forward(reinterpret_cast<const MmbEpicModel*>(m), x, k, mask, temb, time_stride ? T : 0, B, N, ws + L.v, ws + L.logits,
                              ws + L.hidden, precision, s);
    if (rc) return rc;
    TfStackIO io{};
    io.mode = 1; io.H = H; io.S = S; io.hidden = ws + L.hidden; io.mask = mask; io.onehot = onehot;
    io.tbias = tb1; io.tbias_stride = time_stride ? nb * kC : 0; io.dot_out = nl; io.jet_out = ws + L.means;
    io.pack_scratch = reinterpret_cast<int32_t*>(ws + L.pack); io.pack_scratch_ints = L.pack_ints;
    if ((rc = launch_tf_stack(&h->s1, h->sm_count, io, B, N, s))) return rc;
    if ((rc = launch_jet_head(&h->s1, ws + L.means, B, x0l, s))) return rc;
    trans_rate_kernel<<<(B + 3) / 4, 128, 0, s>>>(x0l, nl, dims, ts, ts_stride, nearest_in, u_nearest, fr, h->logfact, B, N, R, direct ? 1 : 0,
                                                  ws + L.rate, nearest);
    if ((rc = cuda_ok(cudaGetLastError(), "trans rate launch"))) return rc;
    io.mode = 2; io.x = x; io.nearest = nearest; io.tbias = tb2; io.dot_out = ws + L.vec_w; io.jet_out = ws + L.means;
    if ((rc = launch_tf_stack(&h->s2, h->sm_count, io, B, N, s))) return rc;
    if ((rc = launch_jet_head(&h->s2, ws + L.means, B, ws + L.post_auto, s))) return rc;
    trans_auto_kernel<<<(B + 3) / 4, 128, 0, s>>>(x, mask, nearest, ws + L.vec_w, ws + L.post_auto, dims, B, N, S, ws + L.new_mean,
                                                 ws + L.new_std, auto_mean, auto_std);
    return cuda_ok(cudaGetLastError(), "trans auto launch");
}

struct CorrectorArgs {   // a corrector row: device coefficients, the predictor step's mask, jump corrector
    const float* coef = nullptr;
    const int32_t* mask_dims = nullptr;
    int jump_mode = 1;
    float death_prob = 0.0f;
    const float* u_death = nullptr;
};

static int launch_sampler_update(float* x, float* onehot, int32_t* dims, const float* v, const float* logits, const float* rate,
                                 const float* new_mean, const float* new_std, float c_decay, float c_score, float c_noise, float inv_std,
                                 float jump_dt, const float* z_diff, const float* u_jump, const float* z_new, uint64_t seed,
                                 uint64_t jet_offset, int step, int B, int N, int S, cudaStream_t s, const CorrectorArgs& ca = CorrectorArgs()) {
    if (N > 128 || N < 1) return fail(MMB_EUNSUPPORTED, "sampler update handles 1..128 particle slots per jet");
#define MMB_UPD2(SV, MODE)                                                                                                            \
    trans_sampler_onehot_kernel<SV, MODE != 1><<<(unsigned)(((size_t)B * N + 255) / 256), 256, 0, s>>>(                                \
        onehot, dims, logits, c_decay, -(c_score * inv_std), c_noise, ca.coef, ca.mask_dims, z_diff, seed, jet_offset, step, B, N);     \
    trans_sampler_update_kernel<SV, MODE><<<(B + 3) / 4, 128, 0, s>>>(x, onehot, dims, v, logits, rate, new_mean, new_std, c_decay,      \
                                                                      c_score, c_noise, inv_std, jump_dt, ca.coef, ca.mask_dims,        \
                                                                      ca.death_prob, z_diff, u_jump, ca.u_death, z_new, seed,          \
                                                                      jet_offset, step, B, N)
#define MMB_UPD(SV)                                                                                                                  \
    do {                                                                                                                             \
        if (ca.jump_mode == 1) { MMB_UPD2(SV, 1); } else if (ca.jump_mode == 0) { MMB_UPD2(SV, 0); } else { MMB_UPD2(SV, 2); }          \
    } while (0)
    if (ca.jump_mode != 1 && (!ca.coef || !ca.mask_dims)) return fail(MMB_EINVAL, "corrector update without coefficients / mask");
    switch (S) {
        case 4: MMB_UPD(4); break;
        case 5: MMB_UPD(5); break;
        case 6: MMB_UPD(6); break;
        case 8: MMB_UPD(8); break;
        default: return fail(MMB_EUNSUPPORTED, "sampler update is built for vocab sizes 4, 5, 6, 8 (got %d)", S);
    }
#undef MMB_UPD2
#undef MMB_UPD
    return cuda_ok(cudaGetLastError(), "sampler update launch");
}

// Langevin corrector row: batch norms -> step size on the device -> the same two update kernels with device coefficients
static int launch_corrector_update(float* x, float* onehot, int32_t* dims, const int32_t* mask_dims, const float* v, const float* logits,
                                   const float* rate, const float* new_mean, const float* new_std, float alpha, int noise_on, float inv_std,
                                   float snr, float jump_dt, int jump_corrector, float death_prob, const float* z_diff, const float* u_jump,
                                   const float* u_death, const float* z_new, uint64_t seed, uint64_t jet_offset, int step, int B, int N, int S,
                                   float* norm_s, float* norm_n, float* coef, cudaStream_t s) {
    if (N > 128 || N < 1) return fail(MMB_EUNSUPPORTED, "sampler update handles 1..128 particle slots per jet");
#define MMB_NORMS(SV)                                                                                                                  \
    trans_corrector_norms_kernel<SV><<<(B + 3) / 4, 128, 0, s>>>(dims, v, logits, inv_std, z_diff, seed, jet_offset, step, B, N, norm_s, norm_n)
    switch (S) {
        case 4: MMB_NORMS(4); break;
        case 5: MMB_NORMS(5); break;
        case 6: MMB_NORMS(6); break;
        case 8: MMB_NORMS(8); break;
        default: return fail(MMB_EUNSUPPORTED, "sampler update is built for vocab sizes 4, 5, 6, 8 (got %d)", S);
    }
#undef MMB_NORMS
    trans_corrector_coef_kernel<<<1, 256, 0, s>>>(norm_s, norm_n, B, snr, alpha, inv_std, noise_on, coef);
    if (int rc = cuda_ok(cudaGetLastError(), "corrector norms launch")) return rc;
    CorrectorArgs ca;
    ca.coef = coef;
    ca.mask_dims = mask_dims;
    ca.jump_mode = jump_corrector ? 2 : 0;
    ca.death_prob = death_prob;
    ca.u_death = u_death;
    // c_noise only tells the kernels whether noise is drawn at all; the value comes from coef
    return launch_sampler_update(x, onehot, dims, v, logits, rate, new_mean, new_std, 1.0f, 0.0f, noise_on ? 1.0f : 0.0f, inv_std, jump_dt,
                                 z_diff, u_jump, z_new, seed, jet_offset, step, B, N, S, s, ca);
}

}  // namespace mmb

using namespace mmb;

extern "C" {

size_t mmb_trans_packed_floats(const MmbTransDims* dims) { return dims ? trans_floats(*dims) : 0; }

int mmb_trans_create(const MmbTransDims* dims, const float* packed, size_t n_floats, int device, MmbTransHeads** out) {
    if (!dims || !packed || !out) return fail(MMB_EINVAL, "mmb_trans_create: null argument");
    TransHeads* h = nullptr;
    try {   // the one-time packing uses std::vector: nothing may throw across the C ABI
        if (int rc = trans_create(dims, packed, n_floats, device, &h)) return rc;
    } catch (...) {
        return fail(MMB_ENOMEM, "mmb_trans_create: out of host memory");
    }
    *out = reinterpret_cast<MmbTransHeads*>(h);
    return MMB_OK;
}

void mmb_trans_destroy(MmbTransHeads* heads) { trans_destroy(reinterpret_cast<TransHeads*>(heads)); }

size_t mmb_trans_forward_workspace_bytes(const MmbEpicModel* trunk, const MmbTransHeads* heads, int B, int N) {
    if (!trunk || !heads || B < 0 || N < 0) return 0;
    return TransWs(reinterpret_cast<const EpicModel*>(trunk)->dims, reinterpret_cast<const TransHeads*>(heads)->d, B, N, B).total * sizeof(float);
}

int mmb_trans_forward(const MmbEpicModel* trunk, const MmbTransHeads* heads, const float* x, const float* onehot, const int32_t* dims,
                      const float* ts, const int32_t* nearest_in, const float* u_nearest, const MmbForwardRate* forward_rate, int B, int N,
                      float* d_xt, float* rate, float* auto_mean, float* auto_std, float* x0_dim_logits, float* near_atom_logits,
                      int32_t* nearest_out, void* workspace, size_t workspace_bytes, int precision, void* stream) {
    const EpicModel* m = reinterpret_cast<const EpicModel*>(trunk);
    const TransHeads* h = reinterpret_cast<const TransHeads*>(heads);
    if (!m || !h || !x || !onehot || !dims || !ts || !forward_rate || !d_xt || !rate || !x0_dim_logits || !near_atom_logits || !workspace)
        return fail(MMB_EINVAL, "mmb_trans_forward: null argument");
    if (!nearest_in && !u_nearest) return fail(MMB_EINVAL, "mmb_trans_forward: give nearest_in or u_nearest");
    if ((auto_mean == nullptr) != (auto_std == nullptr)) return fail(MMB_EINVAL, "mmb_trans_forward: auto_mean and auto_std go together");
    if (B < 0) return fail(MMB_EINVAL, "mmb_trans_forward: negative size");
    if (int rc = check_pair(m, h, N)) return rc;
    const TransWs L(m->dims, h->d, B, N, B);
    if (workspace_bytes < L.total * sizeof(float)) return fail(MMB_ENOMEM, "mmb_trans_forward: workspace %zu B < %zu B", workspace_bytes, L.total * sizeof(float));
    if (B == 0) return MMB_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* ws = static_cast<float*>(workspace);
    const int S = h->d.vocab_size, F = 3 + S, T = m->dims.dim_time_emb;
    trans_time_kernel<<<(B + kTimeJets - 1) / kTimeJets, 2 * kC, 0, s>>>(h->time_wT, h->d.n_blocks, ts, B, T, ws + L.temb, ws + L.tb1, ws + L.tb2);
    if (int rc = cuda_ok(cudaGetLastError(), "trans time launch")) return rc;
    if (auto_mean) {
        if (int rc = cuda_ok(cudaMemsetAsync(auto_mean, 0, (size_t)B * N * F * sizeof(float), s), "auto_mean clear")) return rc;
        if (int rc = cuda_ok(cudaMemsetAsync(auto_std, 0, (size_t)B * N * F * sizeof(float), s), "auto_std clear")) return rc;
    }
    if (int rc = trans_eval(m, h, x, onehot, dims, ts, 1, ws + L.temb, ws + L.tb1, ws + L.tb2, 1, nearest_in, u_nearest, *forward_rate, B, N,
                            ws, L, x0_dim_logits, near_atom_logits, auto_mean, auto_std, precision, s))
        return rc;
    // D_xt = [all continuous slots | all one-hot slots] per jet (transdimensional_model.py:277-280)
    int rc = cuda_ok(cudaMemcpy2DAsync(d_xt, (size_t)N * F * 4, ws + L.v, (size_t)N * 3 * 4, (size_t)N * 3 * 4, B, cudaMemcpyDeviceToDevice, s), "D_xt v");
    if (!rc) rc = cuda_ok(cudaMemcpy2DAsync(d_xt + (size_t)N * 3, (size_t)N * F * 4, ws + L.logits, (size_t)N * S * 4, (size_t)N * S * 4, B,
                                            cudaMemcpyDeviceToDevice, s), "D_xt logits");
    if (!rc) rc = cuda_ok(cudaMemcpyAsync(rate, ws + L.rate, (size_t)B * 4, cudaMemcpyDeviceToDevice, s), "rate");
    if (!rc && nearest_out) rc = cuda_ok(cudaMemcpyAsync(nearest_out, ws + L.nearest, (size_t)B * 4, cudaMemcpyDeviceToDevice, s), "nearest");
    return rc;
}

size_t mmb_trans_sample_workspace_bytes(const MmbEpicModel* trunk, const MmbTransHeads* heads, int B, int N) {
    if (!trunk || !heads || B < 0 || N < 0) return 0;
    // time tables for up to 4096 steps
    return TransWs(reinterpret_cast<const EpicModel*>(trunk)->dims, reinterpret_cast<const TransHeads*>(heads)->d, B, N, 4096).total * sizeof(float);
}

int mmb_trans_sample(const MmbEpicModel* trunk, const MmbTransHeads* heads, float* x, float* onehot, int32_t* dims,
                     const MmbJumpSchedule* sch, const MmbForwardRate* forward_rate, const float* z_diff, const float* u_near,
                     const float* u_jump, const float* z_new, const float* u_death, uint64_t seed, uint64_t jet_offset, int B, int N,
                     void* workspace, size_t workspace_bytes, int precision, void* stream) {
    const EpicModel* m = reinterpret_cast<const EpicModel*>(trunk);
    const TransHeads* h = reinterpret_cast<const TransHeads*>(heads);
    if (!m || !h || !x || !onehot || !dims || !sch || !forward_rate || !workspace) return fail(MMB_EINVAL, "mmb_trans_sample: null argument");
    if (!sch->ts || !sch->c_decay || !sch->c_score || !sch->c_noise || !sch->inv_std) return fail(MMB_EINVAL, "mmb_trans_sample: incomplete schedule");
    const int n = sch->n_steps;
    if (B < 0 || n < 0 || n > 4096) return fail(MMB_EINVAL, "mmb_trans_sample: 0..4096 steps");
    const bool injected = z_diff || u_near || u_jump || z_new || u_death;
    if (injected && !(z_diff && u_near && u_jump && z_new)) return fail(MMB_EINVAL, "mmb_trans_sample: inject all four noise arrays or none");
    bool correctors = false;
    for (int i = 0; sch->kind && i < n; ++i) {
        if (sch->kind[i] > 1) return fail(MMB_EINVAL, "mmb_trans_sample: row kind %d", (int)sch->kind[i]);
        correctors = correctors || sch->kind[i] == 1;
    }
    if (correctors && sch->jump_corrector) {
        if (!sch->death_prob) return fail(MMB_EINVAL, "mmb_trans_sample: jump corrector needs death_prob");
        if (injected && !u_death) return fail(MMB_EINVAL, "mmb_trans_sample: jump corrector with injected noise needs u_death");
    }
    if (int rc = check_pair(m, h, N)) return rc;
    const TransWs L(m->dims, h->d, B, N, 4096);
    if (workspace_bytes < L.total * sizeof(float)) return fail(MMB_ENOMEM, "mmb_trans_sample: workspace too small");
    if (B == 0 || n == 0) return MMB_OK;
    cudaStream_t s = static_cast<cudaStream_t>(stream);
    float* ws = static_cast<float*>(workspace);
    const int S = h->d.vocab_size, F = 3 + S, T = m->dims.dim_time_emb, nb = h->d.n_blocks;
    // all jets share ts: the time terms of every step are computed once, as a batch of n "jets"
    float* ts_dev = ws + L.ts;
    int32_t* mask_dims = reinterpret_cast<int32_t*>(ws + L.mask_dims);
    if (correctors)   // a schedule that opens with corrector rows takes the current dims as its mask
        if (int rc = cuda_ok(cudaMemcpyAsync(mask_dims, dims, (size_t)B * 4, cudaMemcpyDeviceToDevice, s), "mask dims")) return rc;
    if (int rc = cuda_ok(cudaMemcpyAsync(ts_dev, sch->ts, (size_t)n * 4, cudaMemcpyHostToDevice, s), "schedule upload")) return rc;
    trans_time_kernel<<<(n + kTimeJets - 1) / kTimeJets, 2 * kC, 0, s>>>(h->time_wT, nb, ts_dev, n, T, ws + L.temb, ws + L.tb1, ws + L.tb2);
    if (int rc = cuda_ok(cudaGetLastError(), "trans time launch")) return rc;
    for (int i = 0; i < n; ++i) {
        const float* un = u_near ? u_near + (size_t)i * B : ws + L.u_near;
        if (!u_near) {
            trans_philox_u_near_kernel<<<(B + 255) / 256, 256, 0, s>>>(ws + L.u_near, seed, jet_offset, i, B);
            if (int rc = cuda_ok(cudaGetLastError(), "u_near launch")) return rc;
        }
        int rc = trans_eval(m, h, x, onehot, dims, ts_dev + i, 0, ws + L.temb + (size_t)i * T, ws + L.tb1 + (size_t)i * nb * kC,
                            ws + L.tb2 + (size_t)i * nb * kC, 0, nullptr, un, *forward_rate, B, N, ws, L, nullptr, nullptr, nullptr, nullptr,
                            precision, s);
        const float* zd = z_diff ? z_diff + (size_t)i * B * N * F : nullptr;
        const float* uj = u_jump ? u_jump + (size_t)i * B : nullptr;
        const float* zn = z_new ? z_new + (size_t)i * B * F : nullptr;
        if (!rc && sch->kind && sch->kind[i] == 1) {
            rc = launch_corrector_update(x, onehot, dims, mask_dims, ws + L.v, ws + L.logits, ws + L.rate, ws + L.new_mean, ws + L.new_std,
                                         sch->c_score[i], sch->c_noise[i] != 0.0f, sch->inv_std[i], sch->corrector_snr, sch->jump_dt,
                                         sch->jump_corrector, sch->death_prob ? sch->death_prob[i] : 0.0f, zd, uj,
                                         u_death ? u_death + (size_t)i * B : nullptr, zn, seed, jet_offset, i, B, N, S, ws + L.norm_s,
                                         ws + L.norm_n, ws + L.coef, s);
        } else if (!rc) {
            // corrector rows reuse the live mask of their predictor step (sampler.py:219): keep the dims it was built from
            if (correctors) rc = cuda_ok(cudaMemcpyAsync(mask_dims, dims, (size_t)B * 4, cudaMemcpyDeviceToDevice, s), "mask dims");
            if (!rc)
                rc = launch_sampler_update(x, onehot, dims, ws + L.v, ws + L.logits, ws + L.rate, ws + L.new_mean, ws + L.new_std,
                                           sch->c_decay[i], sch->c_score[i], sch->c_noise[i], sch->inv_std[i], sch->jump_dt, zd, uj, zn, seed,
                                           jet_offset, i, B, N, S, s);
        }
        if (rc) return rc;
    }
    return MMB_OK;
}

int mmb_trans_sampler_update(float* x, float* onehot, int32_t* dims, const float* v, const float* logits, const float* rate,
                             const float* new_mean, const float* new_std, float c_decay, float c_score, float c_noise, float inv_std,
                             float jump_dt, const float* z_diff, const float* u_jump, const float* z_new, uint64_t seed,
                             uint64_t jet_offset, int step, int B, int N, int S, void* stream) {
    if (!x || !onehot || !dims || !v || !logits || !rate || !new_mean || !new_std) return fail(MMB_EINVAL, "mmb_trans_sampler_update: null argument");
    if (B < 0) return fail(MMB_EINVAL, "mmb_trans_sampler_update: negative size");
    if (B == 0) return MMB_OK;
    return launch_sampler_update(x, onehot, dims, v, logits, rate, new_mean, new_std, c_decay, c_score, c_noise, inv_std, jump_dt, z_diff,
                                 u_jump, z_new, seed, jet_offset, step, B, N, S, static_cast<cudaStream_t>(stream));
}

int mmb_trans_corrector_update(float* x, float* onehot, int32_t* dims, const int32_t* mask_dims, const float* v, const float* logits,
                               const float* rate, const float* new_mean, const float* new_std, float alpha, int noise_on, float inv_std,
                               float corrector_snr, float jump_dt, int jump_corrector, float death_prob, const float* z_diff,
                               const float* u_jump, const float* u_death, const float* z_new, uint64_t seed, uint64_t jet_offset, int step,
                               int B, int N, int S, float* scratch, void* stream) {
    if (!x || !onehot || !dims || !v || !logits || !scratch) return fail(MMB_EINVAL, "mmb_trans_corrector_update: null argument");
    if (jump_corrector && (!rate || !new_mean || !new_std)) return fail(MMB_EINVAL, "mmb_trans_corrector_update: jump corrector needs rate, new_mean, new_std");
    if (B < 0) return fail(MMB_EINVAL, "mmb_trans_corrector_update: negative size");
    if (B == 0) return MMB_OK;
    return launch_corrector_update(x, onehot, dims, mask_dims ? mask_dims : dims, v, logits, rate, new_mean, new_std, alpha, noise_on, inv_std,
                                   corrector_snr, jump_dt, jump_corrector, death_prob, z_diff, u_jump, u_death, z_new, seed, jet_offset, step, B,
                                   N, S, scratch, scratch + B, scratch + 2 * (size_t)B, static_cast<cudaStream_t>(stream));
}

}  // extern "C"
