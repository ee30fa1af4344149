// epic_mma.cu — the EPiC generation loop as register-resident warp-level tensor-core chains (mma.sync m16n8k16), sm_100a.
//
// Why a second tensor-core engine next to epic_tc.cu (tcgen05): the default EPiC is a chain of K = N = 16 GEMMs.  On tcgen05
// every link of that chain is a round trip  st.shared A tile -> fence.proxy.async -> CTA barrier -> tcgen05.mma -> commit ->
// mbarrier wait -> tcgen05.ld  (~500 cycles, seven per solver step) and the per-jet global MLP runs on one warp while the
// others wait; ncu of the round-1 kernel: tensor pipe 12 %, issue slots 52 %, barrier stalls first.  With warp-level MMA the
// accumulator fragment of one layer IS the A fragment of the next (row g / g+8, column pairs 2t, 2t+1 in both layouts), so a
// warp carries 32 particles of one jet through the whole network without leaving its registers:
//
//   * warp = 32 consecutive particles of ONE jet = two m16 tiles (a warp whose upper 16 rows are all dead runs the one-tile
//     copy of the step loop); a jet with m live particles occupies ceil(last_live / 32) warps and no others — no dead warps,
//     no tile packing constraints, any N up to 256;
//   * per-particle Linear = 4 mma.sync per warp (2 m-tiles x 2 n-tiles, K = 16), weights as pre-swizzled B fragments in
//     shared memory (one 8-byte load per lane per tile); bias = accumulator initialisation; leaky-ReLU / residual / skip on
//     the fp32 accumulator; pack to 16-bit pairs = next A fragment;
//   * masked sum pooling: fp32 column sums over the thread's four rows, then a three-stage xor butterfly over the row groups
//     — every lane ends with the sums of ITS fragment columns, which is precisely an A fragment with 16 identical rows;
//   * so the per-jet global MLP is tensor-core work too: M = 16 (identical rows), K = 16..48, N = 16, accumulators stay in
//     the fragment layout, and the per-jet bias of fc_local1 comes out as exactly the accumulator initialisation the
//     per-particle GEMM needs.  Jets that span several warps exchange 16 partial sums through shared memory (named barrier
//     per jet, ping-pong buffers) and then each warp runs the same global MLP redundantly — identical bits in every warp;
//   * the hybrid update runs in the "owner" layout (lane = particle): logits / velocity go through a 1.5 KB per-warp staging
//     tile, the first A fragment comes back through the same tile with ldmatrix.
//
// Operand types (template): fp16 (11-bit mantissa; sums scaled by 2^-7 so that nothing overflows) or bf16 with the weights
// split W = hi + lo as in epic_tc.cu.  fp32 accumulate, fp32 residual stream / skip / pooling / update in both.
//
// Reference semantics: SURVEY.md §A.2 (mp/models/architectures/epic.py:136-241, utils.py:112-172,
// mp/models/generative/multimodal_bridge_matching.py:90-113,199-216, bridges.py:38-45,106-132,179-201).
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include <type_traits>
#include <vector>

#include "mmb_device.cuh"
#include "mmb_internal.h"

namespace mmb {
namespace {

constexpr int kH = 16;          // hidden width this engine is built for
#ifndef MMB_MMA_MINB
#define MMB_MMA_MINB 3
#endif
constexpr int kW = 8;           // warps per CTA
constexpr int kMaxCls = 8;      // a jet spans at most this many warps (N <= 256)
constexpr int kMaxL = 4;
constexpr int kStageBytes = 1536;   // per warp: A tile [32 rows][48 B]  /  logits [32][8] f32 + velocity [32][4] f32
constexpr int kSkipBytes = 2048;    // per warp: fp32 skip connection, [4][32 lanes] float4
constexpr int kPoolFloats = 32;     // per warp: two (ping-pong) slots of 16 partial column sums
constexpr float kSumScale = 1.0f / 128.0f;   // pooled sums enter the fp16 GEMMs scaled by 2^-7 (weights by 2^7): exact, overflow-safe

// ---- image: B-fragment tiles (256 B each: lane l holds {b0, b1}) + fp32 bias table --------------------------------------------------
struct MmaLayout {
    int L, G, GT, T, skip, S, Sh, Dc;
    // tile indices
    int t_local0, t_g0, t_g1, t_g2, t_layer0, layer_tiles, o_g1, o_g2, o_l1g, o_l1, o_l2, t_out, t_h2, n_tiles;
    int lo_tiles;     // 0 or n_tiles: the low-order halves follow the high-order ones
    // float offsets in the bias table
    int b_g1, b_g2, b_layer0, layer_floats, ob_g2, ob_l2, b_out, b_h2, n_floats;
    __host__ __device__ int tile_layer(int l) const { return t_layer0 + l * layer_tiles; }
    __host__ __device__ size_t tile_bytes() const { return (size_t)(n_tiles + lo_tiles) * 256; }
    __host__ __device__ size_t image_bytes() const { return tile_bytes() + (size_t)n_floats * 4; }
};

MmaLayout make_layout(const MmbEpicDims& d, bool lo) {
    MmaLayout m{};
    m.L = d.num_blocks; m.G = d.dim_hidden_glob; m.GT = (d.dim_hidden_glob + 15) / 16; m.T = d.dim_time_emb;
    m.skip = d.skip_connection; m.S = d.vocab_size; m.Sh = d.disc_head_hidden; m.Dc = d.dim_continuous;
    int t = 0;
    auto take = [&](int n) { int r = t; t += n; return r; };
    m.t_local0 = take(2);
    m.t_g0 = take(4);                 // [mean | sum] x [n0, n1]
    m.t_g1 = take(2);
    m.t_g2 = take(2 * m.GT);
    m.t_layer0 = t;
    {
        int p = 0;
        auto tk = [&](int n) { int r = p; p += n; return r; };
        m.o_g1 = tk((2 + m.GT) * 2);  // k-steps [mean, sum, xg...] x [n0, n1]
        m.o_g2 = tk(2 * m.GT);
        m.o_l1g = tk(m.GT * 2);       // k-steps over xg x [n0, n1]
        m.o_l1 = tk(2);
        m.o_l2 = tk(2);
        m.layer_tiles = p;
    }
    t += m.layer_tiles * m.L;
    m.t_out = take(2);                // n-tile 0: head pre-activation (or raw logits), n-tile 1: velocity
    m.t_h2 = take(1);
    m.n_tiles = t;
    m.lo_tiles = lo ? t : 0;
    int o = 0;
    auto takef = [&](int n) { int r = o; o += n; return r; };
    m.b_g1 = takef(16); m.b_g2 = takef(16 * m.GT);
    m.b_layer0 = o;
    {
        int p = 0;
        auto tk = [&](int n) { int r = p; p += n; return r; };
        m.ob_g2 = tk(16 * m.GT); m.ob_l2 = tk(16);
        m.layer_floats = p;
    }
    o += m.layer_floats * m.L;
    m.b_out = takef(16);              // [0,8): n-tile 0 bias, [8,16): n-tile 1 bias
    m.b_h2 = takef(8);
    m.n_floats = (o + 3) & ~3;
    return m;
}

// ---- small PTX wrappers ---------------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <bool F16>
__device__ __forceinline__ uint32_t pack2(float lo, float hi) {
    if constexpr (F16) {
        __half2 h = __floats2half2_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    } else {
        __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
        return *reinterpret_cast<uint32_t*>(&h);
    }
}
template <bool F16>
__device__ __forceinline__ float round16(float a) {
    if constexpr (F16) return __half2float(__float2half_rn(a));
    else return __bfloat162float(__float2bfloat16_rn(a));
}

// D += A * B, m16n8k16, fp32 accumulate
template <bool F16>
__device__ __forceinline__ void mma(float (&c)[4], const uint32_t (&a)[4], const uint2 b) {
    if constexpr (F16)
        asm("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
    else
        asm("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
            : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3]) : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b.x), "r"(b.y));
}

__device__ __forceinline__ void ldmatrix_x4(uint32_t (&a)[4], uint32_t addr) {
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(a[0]), "=r"(a[1]), "=r"(a[2]), "=r"(a[3]) : "r"(addr));
}
__device__ __forceinline__ void jet_bar(int id, int threads) { asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory"); }

__device__ __forceinline__ float lrelu_f(float a) { return fmaxf(a, 0.01f * a); }
__device__ __forceinline__ void lrelu4(float (&c)[4]) {
    float t0, t1, t2, t3;
    fmul2(t0, t1, c[0], c[1], 0.01f, 0.01f);
    fmul2(t2, t3, c[2], c[3], 0.01f, 0.01f);
    c[0] = fmaxf(c[0], t0); c[1] = fmaxf(c[1], t1); c[2] = fmaxf(c[2], t2); c[3] = fmaxf(c[3], t3);
}
__device__ __forceinline__ float selu_f(float a) {
    const float scale = 1.0507009873554804934193349852946f;
    const float alpha_scale = 1.0507009873554804934193349852946f * 1.6732632423543772848170429916717f;
    return a > 0.0f ? scale * a : alpha_scale * (__expf(a) - 1.0f);
}

// the categorical jump of the tensor-core engines (same Form B rule as mmb::telegraph_jump, fast intrinsics; epic_tc.cu)
template <int S>
__device__ __forceinline__ int jump_fast(const float (&lg)[S], int k, float u, const StepScalars& sc) {
    float mx = lg[0];
#pragma unroll
    for (int s = 1; s < S; ++s) mx = fmaxf(mx, lg[s]);
    float e[S], z = 0.0f, ek = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        e[s] = __expf(lg[s] - mx);
        z += e[s];
        ek = (k == s) ? e[s] : ek;
    }
    const float zinv = __fdividef(1.0f, z);
    const float base = (1.0f + sc.cc * ek * zinv) * sc.dt, slope = sc.bc * zinv * sc.dt;
    float lam[S], Lam = 0.0f;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        lam[s] = fmaf(e[s], slope, base);
        Lam += lam[s];
    }
    const float E = __expf(-Lam);
    float c = 0.0f;
    int below = 0;
#pragma unroll
    for (int s = 0; s < S; ++s) {
        c = fmaf(lam[s], E, c);
        below += (u >= c) ? 1 : 0;
    }
    return below < S ? below : k;
}

struct MmaParams {
    const uint8_t* image;
    MmaLayout lay;
    float* x;
    uint8_t* k;
    const uint8_t* mask;
    const float* step_tab;   // [n_steps][4] (bc, cc, sp, t)
    int n_steps;
    float dt;
    const float* u_jump;     // [n_steps, B, N] or null
    uint64_t seed, jet_offset;
    int B, N;
    const float* tvec;       // [n_steps][2 + 2L][16] per-step time vectors (prologue kernel)
    const int32_t* counts;   // [1 + kMaxCls]: jets per class (class = warps the jet spans; 0 = empty jet)
    const int32_t* lists;    // [kMaxCls][B]: jets of class c at row c - 1
    const int32_t* jet_cnt;  // [B] live particles
};

// A column vector held in the fragment layout: v[j][b] = element 8 j + 2 t + b  (same in every row group g)
template <bool F16>
__device__ __forceinline__ void vec_frag(uint32_t (&a)[4], const float (&v)[2][2], float scale = 1.0f) {
    a[0] = a[1] = pack2<F16>(v[0][0] * scale, v[0][1] * scale);
    a[2] = a[3] = pack2<F16>(v[1][0] * scale, v[1][1] * scale);
}
// the rounding residual of the same vector (second term of a hi + lo split of the A operand)
template <bool F16>
__device__ __forceinline__ void vec_frag_lo(uint32_t (&a)[4], const float (&v)[2][2], float scale = 1.0f) {
    float r[2][2];
#pragma unroll
    for (int j = 0; j < 2; ++j)
#pragma unroll
        for (int b = 0; b < 2; ++b) r[j][b] = v[j][b] * scale - round16<F16>(v[j][b] * scale);
    a[0] = a[1] = pack2<F16>(r[0][0], r[0][1]);
    a[2] = a[3] = pack2<F16>(r[1][0], r[1][1]);
}

template <int DC, int S, int SH, int GT, bool F16, bool WLO>
__global__ void __launch_bounds__(kW * 32, MMB_MMA_MINB) epic_mma_generate_kernel(const MmaParams p) {
    extern __shared__ __align__(128) uint8_t smem[];
    const MmaLayout& lay = p.lay;
    const int tid = threadIdx.x;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int g = lane >> 2, t = lane & 3;

    // ---- CTA -> (class, jets).  CTAs are ordered by class, widest jets first; q = kW / class jets per CTA.
    int cls = 0, first = 0, n_here = 0;
    {
        int b = blockIdx.x;
        for (int c = kMaxCls; c >= 1; --c) {
            const int n_c = __ldg(p.counts + c);
            const int q = kW / c;
            const int ncta = (n_c + q - 1) / q;
            if (b < ncta) {
                cls = c;
                first = b * q;
                n_here = min(q, n_c - first);
                break;
            }
            b -= ncta;
        }
        if (cls == 0) return;   // past the last CTA of work (the grid is sized for the worst case)
    }
    // ---- carve shared memory
    const uint2* s_tiles = reinterpret_cast<const uint2*>(smem);
    const float* s_bias = reinterpret_cast<const float*>(smem + lay.tile_bytes());
    uint8_t* s_warp = smem + ((lay.image_bytes() + 127) & ~(size_t)127);
    uint8_t* stage = s_warp + warp * kStageBytes;
    float4* skipbuf = reinterpret_cast<float4*>(s_warp + kW * kStageBytes + warp * kSkipBytes);
    float* pool_all = reinterpret_cast<float*>(s_warp + kW * (kStageBytes + kSkipBytes));
    {
        const uint4* src = reinterpret_cast<const uint4*>(p.image);
        uint4* dst = reinterpret_cast<uint4*>(smem);
        const int n16 = (int)(lay.image_bytes() / 16);
        for (int i = tid; i < n16; i += blockDim.x) dst[i] = __ldg(src + i);
    }
    __syncthreads();

    const int js = warp / cls, slice = warp - js * cls;   // jet slot of this warp, its 32-row slice of the jet
    if (js >= n_here) return;                             // idle warp slot (class does not divide kW, or the last CTA of a class)
    const int jet = __ldg(p.lists + (size_t)(cls - 1) * p.B + first + js);
    const int N = p.N;
    const int n = 32 * slice + lane;                      // the particle this lane OWNS (state, update)
    const bool valid = n < N;
    const size_t pidx = (size_t)jet * N + n;
    float xs[DC];
    int kk = 0;
    bool live = false;
    if (valid) live = p.mask[pidx] != 0;
#pragma unroll
    for (int c = 0; c < DC; ++c) xs[c] = live ? p.x[pidx * DC + c] : 0.0f;
    if (live) kk = p.k[pidx];
    const unsigned bal = __ballot_sync(0xffffffffu, live);
    const float inv_cnt = 1.0f / (float)__ldg(p.jet_cnt + jet);

    const int bar_id = 1 + js, bar_threads = 32 * cls;
    float* pool_mine = pool_all + warp * kPoolFloats;
    const float* pool_jet = pool_all + (js * cls) * kPoolFloats;
    const int L = lay.L;
    const uint64_t jet_key = p.jet_offset + (uint64_t)jet;

    auto tile = [&](int idx) { return s_tiles[idx * 32 + lane]; };
    auto tile_lo = [&](int idx) { return s_tiles[(idx + lay.n_tiles) * 32 + lane]; };
    auto bias2 = [&](const float* v16, int j) { return *reinterpret_cast<const float2*>(v16 + 8 * j + 2 * t); };

    auto run = [&](auto nmt_tag) {
        constexpr int NMT = decltype(nmt_tag)::value;
        // row masks of this thread's fragment rows: [mt][hh] = row 16 mt + 8 hh + g
        float mk[NMT][2];
#pragma unroll
        for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
            for (int hh = 0; hh < 2; ++hh) mk[mt][hh] = ((bal >> (16 * mt + 8 * hh + g)) & 1u) ? 1.0f : 0.0f;
        float xl[NMT][2][4];
        uint32_t af[NMT][4];
        uint32_t uq0 = 0, uq1 = 0, uq2 = 0, uq3 = 0;
        int pp = 0;   // ping-pong slot of the pooling exchange

        // D[mt][j] += A[mt] * W(tile idx + j) for both m-tiles and both n-tiles of a 16 -> 16 Linear
        auto gemm16 = [&](float (&acc)[NMT][2][4], const uint32_t (&a)[NMT][4], int idx) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                const uint2 bh = tile(idx + j);
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt) mma<F16>(acc[mt][j], a[mt], bh);
                if constexpr (WLO) {
                    const uint2 bl = tile_lo(idx + j);
#pragma unroll
                    for (int mt = 0; mt < NMT; ++mt) mma<F16>(acc[mt][j], a[mt], bl);
                }
            }
        };
        auto pack_rows = [&](uint32_t (&a)[NMT][4], const float (&v)[NMT][2][4]) {
#pragma unroll
            for (int mt = 0; mt < NMT; ++mt) {
                a[mt][0] = pack2<F16>(v[mt][0][0], v[mt][0][1]);
                a[mt][1] = pack2<F16>(v[mt][0][2], v[mt][0][3]);
                a[mt][2] = pack2<F16>(v[mt][1][0], v[mt][1][1]);
                a[mt][3] = pack2<F16>(v[mt][1][2], v[mt][1][3]);
            }
        };
        // masked column sums of xl over the jet: this warp's rows by adds + xor butterfly, then the jet's other warps
        auto pool = [&](float (&s)[2][2]) {
#pragma unroll
            for (int j = 0; j < 2; ++j) {
                float a0 = xl[0][j][0] + xl[0][j][2], a1 = xl[0][j][1] + xl[0][j][3];
                if constexpr (NMT == 2) {
                    a0 += xl[1][j][0] + xl[1][j][2];
                    a1 += xl[1][j][1] + xl[1][j][3];
                }
                s[j][0] = a0; s[j][1] = a1;
            }
#pragma unroll
            for (int off = 4; off <= 16; off <<= 1) {
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    s[j][0] += __shfl_xor_sync(0xffffffffu, s[j][0], off);
                    s[j][1] += __shfl_xor_sync(0xffffffffu, s[j][1], off);
                }
            }
            if (cls > 1) {
                if (g == 0) *reinterpret_cast<float4*>(pool_mine + pp * 16 + 4 * t) = make_float4(s[0][0], s[0][1], s[1][0], s[1][1]);
                jet_bar(bar_id, bar_threads);
                float4 acc = *reinterpret_cast<const float4*>(pool_jet + pp * 16 + 4 * t);
                for (int w = 1; w < cls; ++w) {
                    const float4 q = *reinterpret_cast<const float4*>(pool_jet + w * kPoolFloats + pp * 16 + 4 * t);
                    acc.x += q.x; acc.y += q.y; acc.z += q.z; acc.w += q.w;
                }
                s[0][0] = acc.x; s[0][1] = acc.y; s[1][0] = acc.z; s[1][1] = acc.w;
                pp ^= 1;
            }
        };
        // one M = 16 GEMM of the per-jet path: c[j] (+)= vec(A) * W(tiles idx .. idx + nj - 1); only elements 0, 1 of c[j] matter
        auto gstep = [&](float (*c)[4], int nj, const uint32_t (&a)[4], int idx) {
#pragma unroll
            for (int j = 0; j < 2 * GT; ++j) {
                if (j < nj) {
                    mma<F16>(c[j], a, tile(idx + j));
                    if constexpr (WLO) mma<F16>(c[j], a, tile_lo(idx + j));
                }
            }
        };

        for (int step = 0; step < p.n_steps; ++step) {
            const float* tv = p.tvec + (size_t)step * (2 + 2 * L) * 16;
            // ---- (a) first A tile: [x_hi, x_lo, onehot(k)] per live particle, through the staging tile
            {   // every lane writes its row every step (dead rows: zeros — the tile doubles as the logits staging area)
                float hi[DC], lo[DC];
#pragma unroll
                for (int c = 0; c < DC; ++c) {
                    hi[c] = round16<F16>(xs[c]);
                    lo[c] = xs[c] - hi[c];
                }
                static_assert(DC == 3, "row packing below is written for three continuous features");
                constexpr uint32_t ONE = F16 ? 0x3C00u : 0x3F80u;
                const uint32_t sel = live ? (ONE << (16 * (kk & 1))) : 0u;
                const int kw = kk >> 1;
                uint32_t w[8];
                w[0] = pack2<F16>(hi[0], hi[1]);     // xs == 0 on dead rows
                w[1] = pack2<F16>(hi[2], lo[0]);
                w[2] = pack2<F16>(lo[1], lo[2]);
#pragma unroll
                for (int i = 0; i < 4; ++i) w[3 + i] = (i < S / 2 && kw == i) ? sel : 0u;
                w[7] = 0u;
                *reinterpret_cast<uint4*>(stage + lane * 48) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(stage + lane * 48 + 16) = make_uint4(w[4], w[5], w[6], w[7]);
            }
            __syncwarp();
#pragma unroll
            for (int mt = 0; mt < NMT; ++mt)
                ldmatrix_x4(af[mt], smem_u32(stage + (16 * mt + (lane & 7) + 8 * ((lane >> 3) & 1)) * 48 + (lane >> 4) * 16));

            // ---- (b) local_0 (+ folded embeddings), leaky-ReLU, mask -> xl; skip copy
            {
                float acc[NMT][2][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float2 b = __ldg(reinterpret_cast<const float2*>(tv + 8 * j + 2 * t));
#pragma unroll
                    for (int mt = 0; mt < NMT; ++mt) { acc[mt][j][0] = acc[mt][j][2] = b.x; acc[mt][j][1] = acc[mt][j][3] = b.y; }
                }
                gemm16(acc, af, lay.t_local0);
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        lrelu4(acc[mt][j]);
                        fmul2(xl[mt][j][0], xl[mt][j][1], acc[mt][j][0], acc[mt][j][1], mk[mt][0], mk[mt][0]);
                        fmul2(xl[mt][j][2], xl[mt][j][3], acc[mt][j][2], acc[mt][j][3], mk[mt][1], mk[mt][1]);
                    }
                if (lay.skip) {
#pragma unroll
                    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                        for (int j = 0; j < 2; ++j)
                            skipbuf[(2 * mt + j) * 32 + lane] = make_float4(xl[mt][j][0], xl[mt][j][1], xl[mt][j][2], xl[mt][j][3]);
                }
                pack_rows(af, xl);
            }
            // ---- (c) pooling + EPiC_Projection globals (epic.py:187-190)
            float ps[2][2];
            pool(ps);
            float xg[2 * GT][2], skg[2 * GT][2];
            {
                float pm[2][2];
#pragma unroll
                for (int j = 0; j < 2; ++j) { pm[j][0] = ps[j][0] * inv_cnt; pm[j][1] = ps[j][1] * inv_cnt; }
                uint32_t a_mean[4], a_sum[4], a_mean_lo[4], a_sum_lo[4];
                vec_frag<F16>(a_mean, pm);
                vec_frag<F16>(a_sum, ps, F16 ? kSumScale : 1.0f);
                vec_frag_lo<F16>(a_mean_lo, pm);
                vec_frag_lo<F16>(a_sum_lo, ps, F16 ? kSumScale : 1.0f);
                float c[2 * GT][4];
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float2 b = __ldg(reinterpret_cast<const float2*>(tv + 16 + 8 * j + 2 * t));
                    c[j][0] = c[j][2] = b.x; c[j][1] = c[j][3] = b.y;
                }
                gstep(c, 2, a_mean, lay.t_g0);
                gstep(c, 2, a_sum, lay.t_g0 + 2);
#pragma unroll
                for (int j = 0; j < 2; ++j) {   // low-order part of the pooled operands against the high-order weights
                    mma<F16>(c[j], a_mean_lo, tile(lay.t_g0 + j));
                    mma<F16>(c[j], a_sum_lo, tile(lay.t_g0 + 2 + j));
                }
                float v[2][2];
#pragma unroll
                for (int j = 0; j < 2; ++j) { v[j][0] = lrelu_f(c[j][0]); v[j][1] = lrelu_f(c[j][1]); }
                uint32_t a[4];
                vec_frag<F16>(a, v);
#pragma unroll
                for (int j = 0; j < 2; ++j) {
                    const float2 b = bias2(s_bias + lay.b_g1, j);
                    c[j][0] = c[j][2] = b.x; c[j][1] = c[j][3] = b.y;
                }
                gstep(c, 2, a, lay.t_g1);
#pragma unroll
                for (int j = 0; j < 2; ++j) { v[j][0] = lrelu_f(c[j][0]); v[j][1] = lrelu_f(c[j][1]); }
                vec_frag<F16>(a, v);
#pragma unroll
                for (int j = 0; j < 2 * GT; ++j) {
                    const float2 b = bias2(s_bias + lay.b_g2, j);
                    c[j][0] = c[j][2] = b.x; c[j][1] = c[j][3] = b.y;
                }
                gstep(c, 2 * GT, a, lay.t_g2);
#pragma unroll
                for (int j = 0; j < 2 * GT; ++j) {
                    xg[j][0] = lrelu_f(c[j][0]); xg[j][1] = lrelu_f(c[j][1]);
                    skg[j][0] = lay.skip ? xg[j][0] : 0.0f; skg[j][1] = lay.skip ? xg[j][1] : 0.0f;
                }
            }
            // ---- EPiC layers (epic.py:217-241, 152-155)
            for (int l = 0; l < L; ++l) {
                const int tl = lay.tile_layer(l);
                const float* bl = s_bias + lay.b_layer0 + l * lay.layer_floats;
                if (l > 0) pool(ps);   // layer 0 pools the same tensor the projection pooled
                float bl1[2][2];
                {
                    float pm[2][2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) { pm[j][0] = ps[j][0] * inv_cnt; pm[j][1] = ps[j][1] * inv_cnt; }
                    uint32_t a_mean[4], a_sum[4], a_lo[4];
                    vec_frag<F16>(a_mean, pm);
                    vec_frag<F16>(a_sum, ps, F16 ? kSumScale : 1.0f);
                    float c[2 * GT][4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float2 b = __ldg(reinterpret_cast<const float2*>(tv + (2 + 2 * l) * 16 + 8 * j + 2 * t));
                        c[j][0] = c[j][2] = b.x; c[j][1] = c[j][3] = b.y;
                    }
                    gstep(c, 2, a_mean, tl + lay.o_g1);
                    gstep(c, 2, a_sum, tl + lay.o_g1 + 2);
                    vec_frag_lo<F16>(a_lo, pm);
#pragma unroll
                    for (int j = 0; j < 2; ++j) mma<F16>(c[j], a_lo, tile(tl + lay.o_g1 + j));
                    vec_frag_lo<F16>(a_lo, ps, F16 ? kSumScale : 1.0f);
#pragma unroll
                    for (int j = 0; j < 2; ++j) mma<F16>(c[j], a_lo, tile(tl + lay.o_g1 + 2 + j));
                    uint32_t a[4];
#pragma unroll
                    for (int gt = 0; gt < GT; ++gt) {   // xg k-steps
                        const float v2[2][2] = {{xg[2 * gt][0], xg[2 * gt][1]}, {xg[2 * gt + 1][0], xg[2 * gt + 1][1]}};
                        vec_frag<F16>(a, v2);
                        gstep(c, 2, a, tl + lay.o_g1 + 4 + 2 * gt);
                    }
                    float v[2][2];
#pragma unroll
                    for (int j = 0; j < 2; ++j) { v[j][0] = lrelu_f(c[j][0]); v[j][1] = lrelu_f(c[j][1]); }
                    vec_frag<F16>(a, v);
                    // fc_global2 + residual (epic.py:231-232)
#pragma unroll
                    for (int j = 0; j < 2 * GT; ++j) {
                        const float2 b = bias2(bl + lay.ob_g2, j);
                        c[j][0] = c[j][2] = b.x + xg[j][0]; c[j][1] = c[j][3] = b.y + xg[j][1];
                    }
                    gstep(c, 2 * GT, a, tl + lay.o_g2);
                    float xm[2 * GT][2];
#pragma unroll
                    for (int j = 0; j < 2 * GT; ++j) {
                        xm[j][0] = lrelu_f(c[j][0]); xm[j][1] = lrelu_f(c[j][1]);
                        xg[j][0] = xm[j][0] + skg[j][0]; xg[j][1] = xm[j][1] + skg[j][1];   // trunk skip (epic.py:155)
                    }
                    // per-jet part of fc_local1: time vector + Wl1[:, H:H+G] xg   (epic.py:233-238)
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float2 b = __ldg(reinterpret_cast<const float2*>(tv + (3 + 2 * l) * 16 + 8 * j + 2 * t));
                        c[j][0] = c[j][2] = b.x; c[j][1] = c[j][3] = b.y;
                    }
#pragma unroll
                    for (int gt = 0; gt < GT; ++gt) {
                        const float v2[2][2] = {{xm[2 * gt][0], xm[2 * gt][1]}, {xm[2 * gt + 1][0], xm[2 * gt + 1][1]}};
                        vec_frag<F16>(a, v2);
                        gstep(c, 2, a, tl + lay.o_l1g + 2 * gt);
                    }
#pragma unroll
                    for (int j = 0; j < 2; ++j) { bl1[j][0] = c[j][0]; bl1[j][1] = c[j][1]; }
                }
                // fc_local1 -> leaky-ReLU -> fc_local2 + residual -> leaky-ReLU, mask, trunk skip
                uint32_t a1[NMT][4];
                {
                    float acc[NMT][2][4];
#pragma unroll
                    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                        for (int j = 0; j < 2; ++j) { acc[mt][j][0] = acc[mt][j][2] = bl1[j][0]; acc[mt][j][1] = acc[mt][j][3] = bl1[j][1]; }
                    gemm16(acc, af, tl + lay.o_l1);
#pragma unroll
                    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                        for (int j = 0; j < 2; ++j) lrelu4(acc[mt][j]);
                    pack_rows(a1, acc);
                }
                {
                    float acc[NMT][2][4];
#pragma unroll
                    for (int j = 0; j < 2; ++j) {
                        const float2 b = bias2(bl + lay.ob_l2, j);
#pragma unroll
                        for (int mt = 0; mt < NMT; ++mt) {
                            fadd2(acc[mt][j][0], acc[mt][j][1], xl[mt][j][0], xl[mt][j][1], b.x, b.y);
                            fadd2(acc[mt][j][2], acc[mt][j][3], xl[mt][j][2], xl[mt][j][3], b.x, b.y);
                        }
                    }
                    gemm16(acc, a1, tl + lay.o_l2);
#pragma unroll
                    for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                        for (int j = 0; j < 2; ++j) {
                            lrelu4(acc[mt][j]);
                            float4 sk = make_float4(0.f, 0.f, 0.f, 0.f);
                            if (lay.skip) sk = skipbuf[(2 * mt + j) * 32 + lane];
                            ffma2(xl[mt][j][0], xl[mt][j][1], acc[mt][j][0], acc[mt][j][1], mk[mt][0], mk[mt][0], sk.x, sk.y);
                            ffma2(xl[mt][j][2], xl[mt][j][3], acc[mt][j][2], acc[mt][j][3], mk[mt][1], mk[mt][1], sk.z, sk.w);
                        }
                    pack_rows(af, xl);
                }
            }
            // ---- output layer (+ head Linear 0 folded in), SELU, head Linear 2   (epic.py:158-162, mbm.py:105-111)
            float hz[NMT][4], hv[NMT][4];
            {
                const float2 bz = bias2(s_bias + lay.b_out, 0), bv = bias2(s_bias + lay.b_out, 1);
                const uint2 wz = tile(lay.t_out), wv = tile(lay.t_out + 1);
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt) {
                    hz[mt][0] = hz[mt][2] = bz.x; hz[mt][1] = hz[mt][3] = bz.y;
                    hv[mt][0] = hv[mt][2] = bv.x; hv[mt][1] = hv[mt][3] = bv.y;
                    mma<F16>(hz[mt], af[mt], wz);
                    mma<F16>(hv[mt], af[mt], wv);
                }
                if constexpr (WLO) {
                    const uint2 wzl = tile_lo(lay.t_out), wvl = tile_lo(lay.t_out + 1);
#pragma unroll
                    for (int mt = 0; mt < NMT; ++mt) { mma<F16>(hz[mt], af[mt], wzl); mma<F16>(hv[mt], af[mt], wvl); }
                }
            }
            if constexpr (SH > 0) {
                const float2 b2 = bias2(s_bias + lay.b_h2, 0);
                const uint2 w2 = tile(lay.t_h2);
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt) {
                    uint32_t az[4];
                    az[0] = pack2<F16>(selu_f(hz[mt][0]), selu_f(hz[mt][1]));
                    az[1] = pack2<F16>(selu_f(hz[mt][2]), selu_f(hz[mt][3]));
                    az[2] = az[3] = 0u;
                    hz[mt][0] = hz[mt][2] = b2.x; hz[mt][1] = hz[mt][3] = b2.y;
                    mma<F16>(hz[mt], az, w2);
                    if constexpr (WLO) mma<F16>(hz[mt], az, tile_lo(lay.t_h2));
                }
            }
            // ---- fragment layout -> owner layout through the staging tile
            __syncwarp();
            {
                float* slog = reinterpret_cast<float*>(stage);
                float* sv = reinterpret_cast<float*>(stage + 1024);
#pragma unroll
                for (int mt = 0; mt < NMT; ++mt)
#pragma unroll
                    for (int hh = 0; hh < 2; ++hh) {
                        const int row = 16 * mt + 8 * hh + g;
                        *reinterpret_cast<float2*>(slog + row * 8 + 2 * t) = make_float2(hz[mt][2 * hh], hz[mt][2 * hh + 1]);
                        if (t < 2) *reinterpret_cast<float2*>(sv + row * 4 + 2 * t) = make_float2(hv[mt][2 * hh], hv[mt][2 * hh + 1]);
                    }
            }
            __syncwarp();
            // ---- hybrid update of the owned particle (bridges.py:38-45,179-201)
            if (!p.u_jump && (step & 3) == 0) {   // one Philox block serves four particles x one step: share it inside the lane quad
                const uint4 blk = philox_block(p.seed, jet_key, 0, step + (n & 3), n >> 2);
                uint32_t a0 = blk.x, a1 = blk.y, a2 = blk.z, a3 = blk.w;
                {
                    const bool hi = (lane & 2) != 0;
                    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, hi ? a0 : a2, 2), r1 = __shfl_xor_sync(0xffffffffu, hi ? a1 : a3, 2);
                    if (hi) { a0 = r0; a1 = r1; } else { a2 = r0; a3 = r1; }
                }
                {
                    const bool hi = (lane & 1) != 0;
                    const uint32_t r0 = __shfl_xor_sync(0xffffffffu, hi ? a0 : a1, 1), r1 = __shfl_xor_sync(0xffffffffu, hi ? a2 : a3, 1);
                    if (hi) { a0 = r0; a2 = r1; } else { a1 = r0; a3 = r1; }
                }
                uq0 = a0; uq1 = a1; uq2 = a2; uq3 = a3;
            }
            if (live) {
                const float4 v4 = *reinterpret_cast<const float4*>(stage + 1024 + lane * 16);
                const float4 l0 = *reinterpret_cast<const float4*>(stage + lane * 32);
                float lg[S];
                lg[0] = l0.x; lg[1] = l0.y; lg[2] = l0.z; lg[3] = l0.w;
                if constexpr (S > 4) {
                    const float4 l1 = *reinterpret_cast<const float4*>(stage + lane * 32 + 16);
                    lg[4] = l1.x; lg[5] = l1.y; lg[6] = l1.z; lg[7] = l1.w;
                }
                xs[0] = fmaf(p.dt, v4.x, xs[0]); xs[1] = fmaf(p.dt, v4.y, xs[1]); xs[2] = fmaf(p.dt, v4.z, xs[2]);
                const StepScalars sc{p.dt, __ldg(p.step_tab + step * 4 + 0), __ldg(p.step_tab + step * 4 + 1), 0.0f};
                const int ph = step & 3;
                const float u = p.u_jump ? __ldg(p.u_jump + ((size_t)step * p.B + jet) * N + n)
                                         : u01(ph == 0 ? uq0 : ph == 1 ? uq1 : ph == 2 ? uq2 : uq3);
                kk = jump_fast<S>(lg, kk, u, sc);
            }
            __syncwarp();   // the staging tile becomes the A tile of the next step
        }
        // ---- final state: live particles as computed, dead ones 0 (x * mask, k * mask)
        if (valid) {
#pragma unroll
            for (int c = 0; c < DC; ++c) p.x[pidx * DC + c] = live ? xs[c] : 0.0f;
            p.k[pidx] = (uint8_t)(live ? kk : 0);
        }
        for (int m = 32 * cls + 32 * slice + lane; m < N; m += 32 * cls) {   // rows past the last live particle have no warp
            const size_t q = (size_t)jet * N + m;
#pragma unroll
            for (int c = 0; c < DC; ++c) p.x[q * DC + c] = 0.0f;
            p.k[q] = 0;
        }
    };
    if ((bal >> 16) != 0u) run(std::integral_constant<int, 2>{});
    else run(std::integral_constant<int, 1>{});
}

// ---- prologue: per-step time vectors + binning of the jets by the number of warps they span ---------------------------------------------
// blocks [0, n_steps): vectors of one step, from the fp32 weights (same for every jet):
//   v0 = local_0 bias + W0[:, T:T+C] a + W0[:, :T] temb      (a = bias of the continuous embedding)
//   v1 = global_0 bias + G0[:, 2H:] temb;  per layer: fc_global1 bias + time part, fc_local1 bias + time part
// blocks [n_steps, ...): one thread per jet: live count, last live index -> class; jets of a class are appended to its list
// (order irrelevant: a jet's result does not depend on where it runs).  Empty jets get the reference's result right here:
// the mean pool divides by zero (epic.py:141), every feature becomes NaN, tokens are multiplied by the mask -> 0.
constexpr int kPrologueThreads = 256;
__global__ void __launch_bounds__(kPrologueThreads) mma_prologue_kernel(const float* __restrict__ W, MmbEpicLayout Lo, MmbEpicDims d,
                                                                        const float* __restrict__ temb, int n_steps, float* __restrict__ tvec,
                                                                        const uint8_t* __restrict__ mask, int B, int N, int32_t* __restrict__ counts,
                                                                        int32_t* __restrict__ lists, int32_t* __restrict__ jet_cnt,
                                                                        float* __restrict__ x, uint8_t* __restrict__ k) {
    const int T = d.dim_time_emb, C = d.dim_cont_emb, D = d.dim_disc_emb, H = d.dim_hidden_local, G = d.dim_hidden_glob, L = d.num_blocks;
    if ((int)blockIdx.x < n_steps) {
        const int step = blockIdx.x, v = threadIdx.x >> 4, o = threadIdx.x & 15;
        if (v >= 2 + 2 * L) return;
        const float* te = temb + (size_t)step * T;
        float acc;
        const float* wt;
        if (v == 0) {
            const float* w0 = W + Lo.local0_w + (size_t)o * (T + C + D);
            acc = W[Lo.local0_b + o];
            for (int c = 0; c < C; ++c) acc = fmaf(w0[T + c], W[Lo.emb_cont_b + c], acc);
            wt = w0;
        } else if (v == 1) {
            acc = W[Lo.global0_b + o];
            wt = W + Lo.global0_w + (size_t)o * (2 * H + T) + 2 * H;
        } else {
            const float* Wl = W + Lo.layer0 + (size_t)((v - 2) >> 1) * Lo.layer_stride;
            if (v & 1) { acc = Wl[Lo.l_l1_b + o]; wt = Wl + Lo.l_l1_w + (size_t)o * (H + G + T) + H + G; }
            else { acc = Wl[Lo.l_g1_b + o]; wt = Wl + Lo.l_g1_w + (size_t)o * (2 * H + G + T) + 2 * H + G; }
        }
        for (int i = 0; i < T; ++i) acc = fmaf(wt[i], te[i], acc);
        tvec[((size_t)step * (2 + 2 * L) + v) * 16 + o] = acc;
        return;
    }
    __shared__ int s_cnt[1 + kMaxCls], s_base[1 + kMaxCls];
    if (threadIdx.x <= kMaxCls) s_cnt[threadIdx.x] = 0;
    __syncthreads();
    const int jet = ((int)blockIdx.x - n_steps) * kPrologueThreads + threadIdx.x;
    int cls = -1, pos = 0;
    if (jet < B) {
        const uint8_t* row = mask + (size_t)jet * N;
        int cnt = 0, last = 0;
        if ((N & 15) == 0 && (reinterpret_cast<uintptr_t>(mask) & 15) == 0) {
            for (int i = 0; i < N; i += 16) {
                const uint4 q = __ldg(reinterpret_cast<const uint4*>(row + i));
                const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
                for (int j = 0; j < 4; ++j)
#pragma unroll
                    for (int b = 0; b < 4; ++b)
                        if ((w[j] >> (8 * b)) & 0xffu) { ++cnt; last = i + 4 * j + b + 1; }
            }
        } else {
            for (int i = 0; i < N; ++i)
                if (row[i]) { ++cnt; last = i + 1; }
        }
        jet_cnt[jet] = cnt;
        cls = (last + 31) / 32;
        pos = atomicAdd(&s_cnt[cls], 1);
    }
    __syncthreads();
    if (threadIdx.x <= kMaxCls) s_base[threadIdx.x] = s_cnt[threadIdx.x] ? atomicAdd(&counts[threadIdx.x], s_cnt[threadIdx.x]) : 0;
    __syncthreads();
    if (cls > 0) lists[(size_t)(cls - 1) * B + s_base[cls] + pos] = jet;
    if (cls == 0) {
        const float nan = __int_as_float(0x7fc00000);
        for (int i = 0; i < N * d.dim_continuous; ++i) x[(size_t)jet * N * d.dim_continuous + i] = nan;
        for (int i = 0; i < N; ++i) k[(size_t)jet * N + i] = 0;
    }
}

template <typename T16>
inline uint16_t to16(float v);
template <>
inline uint16_t to16<__half>(float v) { __half h = __float2half_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
template <>
inline uint16_t to16<__nv_bfloat16>(float v) { __nv_bfloat16 h = __float2bfloat16_rn(v); return *reinterpret_cast<uint16_t*>(&h); }
template <typename T16>
inline float from16(uint16_t b);
template <>
inline float from16<__half>(uint16_t b) { return __half2float(*reinterpret_cast<__half*>(&b)); }
template <>
inline float from16<__nv_bfloat16>(uint16_t b) { return __bfloat162float(*reinterpret_cast<__nv_bfloat16*>(&b)); }

// One k16 x n8 B-fragment tile of y = W x: lane (g, t) holds W[n0 + g][k0 + 2t, +1] and W[n0 + g][k0 + 2t + 8, +9];
// `w(o, k)` returns the (already folded) weight or 0 outside the matrix.
template <typename T16, typename F>
void fill_tile(std::vector<uint16_t>& img, int n_tiles, int idx, bool lo, int n0, int k0, F w) {
    for (int lane = 0; lane < 32; ++lane) {
        const int g = lane >> 2, t = lane & 3;
        for (int e = 0; e < 4; ++e) {
            const int kcol = k0 + 2 * t + (e & 1) + 8 * (e >> 1);
            const double v = w(n0 + g, kcol);
            const uint16_t hi = to16<T16>((float)v);
            img[((size_t)idx * 32 + lane) * 4 + e] = hi;
            if (lo) img[((size_t)(idx + n_tiles) * 32 + lane) * 4 + e] = to16<T16>((float)(v - (double)from16<T16>(hi)));
        }
    }
}

template <typename T16>
int build_image(const MmbEpicDims& d, const MmbEpicLayout& Lo, const float* W, bool lo, float sum_scale, void** out, size_t* out_bytes) {
    const MmaLayout lay = make_layout(d, lo);
    const int Dc = d.dim_continuous, S = d.vocab_size, T = d.dim_time_emb, C = d.dim_cont_emb, D = d.dim_disc_emb, H = kH,
              G = d.dim_hidden_glob, L = d.num_blocks, Sh = d.disc_head_hidden, GT = lay.GT;
    const int K0 = T + C + D;
    std::vector<uint16_t> img((size_t)(lay.n_tiles + lay.lo_tiles) * 128, 0);
    std::vector<float> bias((size_t)lay.n_floats, 0.0f);
    const double inv_scale = 1.0 / (double)sum_scale;
    auto put = [&](int idx, int n0, int k0, auto w) { fill_tile<T16>(img, lay.n_tiles, idx, lo, n0, k0, w); };
    // local_0 with the embeddings folded in: columns [x_hi (Dc) | x_lo (Dc) | onehot (S)]
    auto w_local0 = [&](int o, int kc) -> double {
        if (o >= H) return 0.0;
        const float* w0 = W + Lo.local0_w + (size_t)o * K0;
        if (kc < 2 * Dc) {
            const int j = kc % Dc;
            double acc = 0;
            for (int c = 0; c < C; ++c) acc += (double)w0[T + c] * W[Lo.emb_cont_w + (size_t)c * Dc + j];
            return acc;
        }
        const int s = kc - 2 * Dc;
        if (s >= S) return 0.0;
        double acc = 0;
        for (int dd = 0; dd < D; ++dd) acc += (double)w0[T + C + dd] * W[Lo.emb_disc + (size_t)s * D + dd];
        return acc;
    };
    for (int j = 0; j < 2; ++j) put(lay.t_local0 + j, 8 * j, 0, w_local0);
    // projection globals: global_0 [H][mean H | sum H | T], global_1 [H][H], global_2 [G][H]
    for (int j = 0; j < 2; ++j) {
        put(lay.t_g0 + j, 8 * j, 0, [&](int o, int kc) { return (double)W[Lo.global0_w + (size_t)o * (2 * H + T) + kc]; });
        put(lay.t_g0 + 2 + j, 8 * j, 0, [&](int o, int kc) { return inv_scale * W[Lo.global0_w + (size_t)o * (2 * H + T) + H + kc]; });
        put(lay.t_g1 + j, 8 * j, 0, [&](int o, int kc) { return (double)W[Lo.global1_w + (size_t)o * H + kc]; });
    }
    for (int j = 0; j < 2 * GT; ++j)
        put(lay.t_g2 + j, 8 * j, 0, [&](int o, int kc) { return o < G ? (double)W[Lo.global2_w + (size_t)o * H + kc] : 0.0; });
    for (int i = 0; i < H; ++i) bias[lay.b_g1 + i] = W[Lo.global1_b + i];
    for (int i = 0; i < G; ++i) bias[lay.b_g2 + i] = W[Lo.global2_b + i];
    for (int l = 0; l < L; ++l) {
        const float* Wl = W + Lo.layer0 + (size_t)l * Lo.layer_stride;
        const int tl = lay.tile_layer(l);
        float* bl = bias.data() + lay.b_layer0 + (size_t)l * lay.layer_floats;
        const int Kg = 2 * H + G + T, Kl = H + G + T;
        for (int j = 0; j < 2; ++j) {
            put(tl + lay.o_g1 + j, 8 * j, 0, [&](int o, int kc) { return (double)Wl[Lo.l_g1_w + (size_t)o * Kg + kc]; });
            put(tl + lay.o_g1 + 2 + j, 8 * j, 0, [&](int o, int kc) { return inv_scale * Wl[Lo.l_g1_w + (size_t)o * Kg + H + kc]; });
            for (int gt = 0; gt < GT; ++gt) {
                put(tl + lay.o_g1 + 4 + 2 * gt + j, 8 * j, 16 * gt,
                    [&](int o, int kc) { return kc < G ? (double)Wl[Lo.l_g1_w + (size_t)o * Kg + 2 * H + kc] : 0.0; });
                put(tl + lay.o_l1g + 2 * gt + j, 8 * j, 16 * gt,
                    [&](int o, int kc) { return kc < G ? (double)Wl[Lo.l_l1_w + (size_t)o * Kl + H + kc] : 0.0; });
            }
            put(tl + lay.o_l1 + j, 8 * j, 0, [&](int o, int kc) { return (double)Wl[Lo.l_l1_w + (size_t)o * Kl + kc]; });
            put(tl + lay.o_l2 + j, 8 * j, 0, [&](int o, int kc) { return (double)Wl[Lo.l_l2_w + (size_t)o * H + kc]; });
        }
        for (int j = 0; j < 2 * GT; ++j)
            put(tl + lay.o_g2 + j, 8 * j, 0, [&](int o, int kc) { return o < G ? (double)Wl[Lo.l_g2_w + (size_t)o * H + kc] : 0.0; });
        for (int i = 0; i < G; ++i) bl[lay.ob_g2 + i] = Wl[Lo.l_g2_b + i];
        for (int i = 0; i < H; ++i) bl[lay.ob_l2 + i] = Wl[Lo.l_l2_b + i];
    }
    // output layer: n-tile 0 = head pre-activation F1 (W_out z-rows) (or the raw logits without a head), n-tile 1 = velocity rows
    if (Sh) {
        put(lay.t_out, 0, 0, [&](int j, int kc) {
            if (j >= Sh) return 0.0;
            double acc = 0;
            for (int s = 0; s < S; ++s) acc += (double)W[Lo.head0_w + (size_t)j * S + s] * W[Lo.out_w + (size_t)(Dc + s) * H + kc];
            return acc;
        });
        for (int j = 0; j < Sh; ++j) {
            double acc = W[Lo.head0_b + j];
            for (int s = 0; s < S; ++s) acc += (double)W[Lo.head0_w + (size_t)j * S + s] * W[Lo.out_b + Dc + s];
            bias[lay.b_out + j] = (float)acc;
        }
        put(lay.t_h2, 0, 0, [&](int o, int kc) { return (o < S && kc < Sh) ? (double)W[Lo.head2_w + (size_t)o * Sh + kc] : 0.0; });
        for (int o = 0; o < S; ++o) bias[lay.b_h2 + o] = W[Lo.head2_b + o];
    } else {
        put(lay.t_out, 0, 0, [&](int o, int kc) { return o < S ? (double)W[Lo.out_w + (size_t)(Dc + o) * H + kc] : 0.0; });
        for (int o = 0; o < S; ++o) bias[lay.b_out + o] = W[Lo.out_b + Dc + o];
    }
    put(lay.t_out + 1, 0, 0, [&](int o, int kc) { return o < Dc ? (double)W[Lo.out_w + (size_t)o * H + kc] : 0.0; });
    for (int o = 0; o < Dc; ++o) bias[lay.b_out + 8 + o] = W[Lo.out_b + o];

    const size_t nb = img.size() * sizeof(uint16_t), nf = bias.size() * sizeof(float);
    uint8_t* dev = nullptr;
    if (int rc = cuda_ok(cudaMalloc(&dev, nb + nf), "cudaMalloc mma image")) return rc;
    int rc = cuda_ok(cudaMemcpy(dev, img.data(), nb, cudaMemcpyHostToDevice), "mma image upload");
    if (!rc) rc = cuda_ok(cudaMemcpy(dev + nb, bias.data(), nf, cudaMemcpyHostToDevice), "mma image upload");
    if (rc) { cudaFree(dev); return rc; }
    *out = dev;
    *out_bytes = nb + nf;
    return MMB_OK;
}

size_t smem_bytes(const MmaLayout& lay) {
    return ((lay.image_bytes() + 127) & ~(size_t)127) + (size_t)kW * (kStageBytes + kSkipBytes) + (size_t)kW * kPoolFloats * 4;
}

template <int DC, int S, int SH, int GT, bool F16, bool WLO>
int launch_kernel(const MmaParams& p, int grid, cudaStream_t stream) {
    const size_t bytes = smem_bytes(p.lay);
    auto kern = epic_mma_generate_kernel<DC, S, SH, GT, F16, WLO>;
    if (int rc = cuda_ok(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bytes), "mma smem attribute")) return rc;
    kern<<<grid, kW * 32, bytes, stream>>>(p);
    return cuda_ok(cudaGetLastError(), "epic_mma launch");
}

template <bool F16, bool WLO>
int dispatch(const MmbEpicDims& d, const MmaParams& p, int grid, cudaStream_t stream) {
    const int sh = d.disc_head_hidden, gt = p.lay.GT;
    if (d.vocab_size == 8 && sh == 8 && gt == 1) return launch_kernel<3, 8, 8, 1, F16, WLO>(p, grid, stream);
    if (d.vocab_size == 8 && sh == 0 && gt == 1) return launch_kernel<3, 8, 0, 1, F16, WLO>(p, grid, stream);
    if (d.vocab_size == 8 && sh == 8 && gt == 2) return launch_kernel<3, 8, 8, 2, F16, WLO>(p, grid, stream);
    if (d.vocab_size == 8 && sh == 0 && gt == 2) return launch_kernel<3, 8, 0, 2, F16, WLO>(p, grid, stream);
    if (d.vocab_size == 4 && sh == 4 && gt == 1) return launch_kernel<3, 4, 4, 1, F16, WLO>(p, grid, stream);
    if (d.vocab_size == 4 && sh == 0 && gt == 1) return launch_kernel<3, 4, 0, 1, F16, WLO>(p, grid, stream);
    return fail(MMB_EUNSUPPORTED, "warp-MMA engine instantiated for Dc=3, (S, head) in {(8,8),(8,0),(4,4),(4,0)}, G<=16 (S=8: G<=32)");
}

}  // namespace

bool mma_supported(const MmbEpicDims* d, int N) {
    const bool head_ok = d->disc_head_hidden == 0 || d->disc_head_hidden == d->vocab_size;
    const int gt = (d->dim_hidden_glob + 15) / 16;
    return d->dim_hidden_local == kH && d->dim_hidden_glob >= 1 && gt <= (d->vocab_size == 8 ? 2 : 1) && d->dim_time_emb >= 1 &&
           d->num_blocks >= 1 && d->num_blocks <= kMaxL && d->dim_continuous == 3 && (d->vocab_size == 8 || d->vocab_size == 4) &&
           head_ok && N >= 1 && N <= 32 * kMaxCls;
}

int mma_build_images(EpicModel* m, const float* packed_host) {
    if (int rc = build_image<__half>(m->dims, m->layout, packed_host, false, kSumScale, &m->mma_image_f16, &m->mma_image_f16_bytes)) return rc;
    return build_image<__nv_bfloat16>(m->dims, m->layout, packed_host, true, 1.0f, &m->mma_image_bf16, &m->mma_image_bf16_bytes);
}

// scratch (4-byte units): time vectors [n_steps][2 + 2L][16] | counts [16] | jet_cnt [B] | lists [kMaxCls][B]
size_t mma_generate_scratch_floats(const MmbEpicDims* d, int n_steps, int B) {
    return (((size_t)n_steps * (2 + 2 * d->num_blocks) * 16 + 3) & ~(size_t)3) + 16 + (size_t)(B > 0 ? B : 0) * (1 + kMaxCls) + 16;
}

int launch_generate_mma(const EpicModel* m, float* x, uint8_t* k, const uint8_t* mask, const float* dev_table, float* scratch,
                        int n_steps, float dt, const float* u_jump, uint64_t seed, uint64_t jet_offset,
                        int B, int N, bool f16, cudaStream_t stream) {
    if (B == 0 || n_steps == 0) return MMB_OK;
    const void* image = f16 ? m->mma_image_f16 : m->mma_image_bf16;
    if (!image) return fail(MMB_EUNSUPPORTED, "warp-MMA engine: no operand image for this model");
    MmaParams p{};
    p.image = static_cast<const uint8_t*>(image);
    p.lay = make_layout(m->dims, !f16);
    p.x = x; p.k = k; p.mask = mask;
    p.step_tab = dev_table;
    p.n_steps = n_steps; p.dt = dt; p.u_jump = u_jump; p.seed = seed; p.jet_offset = jet_offset;
    p.B = B; p.N = N;
    const size_t tv_floats = ((size_t)n_steps * (2 + 2 * m->dims.num_blocks) * 16 + 3) & ~(size_t)3;
    int32_t* counts = reinterpret_cast<int32_t*>(scratch + tv_floats);
    int32_t* jet_cnt = counts + 16;
    int32_t* lists = jet_cnt + B;
    if (int rc = cuda_ok(cudaMemsetAsync(counts, 0, 16 * sizeof(int32_t), stream), "mma counters")) return rc;
    const int bin_blocks = (B + kPrologueThreads - 1) / kPrologueThreads;
    mma_prologue_kernel<<<n_steps + bin_blocks, kPrologueThreads, 0, stream>>>(m->w, m->layout, m->dims, dev_table + (size_t)n_steps * 4, n_steps,
                                                                                scratch, mask, B, N, counts, lists, jet_cnt, x, k);
    if (int rc = cuda_ok(cudaGetLastError(), "mma prologue launch")) return rc;
    p.tvec = scratch; p.counts = counts; p.lists = lists; p.jet_cnt = jet_cnt;
    // worst-case number of CTAs: every class rounds up once, the widest class of this N packs the fewest jets per CTA
    const int ncls = (N + 31) / 32;
    const int q_min = kW / ncls > 0 ? kW / ncls : 1;
    const int grid = (B + q_min - 1) / q_min + ncls;
    return f16 ? dispatch<true, false>(m->dims, p, grid, stream) : dispatch<false, true>(m->dims, p, grid, stream);
}

}  // namespace mmb
