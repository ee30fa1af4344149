// epic_wide_tc.cu — EPiC with 128 hidden units per particle on tcgen05 tensor cores (sm_100a).
//
// The class default of the reference's EPiCNetwork is dim_hidden_local = 128, num_blocks = 6 (mp/models/architectures/epic.py:
// 99-101); at that width a network evaluation is real GEMM work: per EPiC layer two [N particles x 128] x [128 x 128] products
// (fc_local1, fc_local2: epic.py:233-238), 0.41 MFLOP per particle and evaluation at L = 6.  This file is the trunk for such
// models (`mmb_epic_forward` / the step loop of `mmb_generate` with MMB_PREC_BF16); the H = 16 engines live in epic_tc.cu /
// epic_mma.cu.
//
// One CTA (512 threads) carries TWO jets at a time ("tiles" A and B, 128 rows each = TMEM lanes), persistent over pairs of jets:
//   * per tile the residual stream X [128 x 128] fp32 and one accumulator ACC [128 x 128] live in TMEM (2 x 256 = all 512 columns);
//     thread (r, cq) serves row r, columns [32 cq, 32 cq + 32) of whichever tile is in its epilogue;
//   * every Linear over the particles is 8 tcgen05.mma M128 x N128 x K16 with bf16 operands; the weights stream L2 -> shared
//     memory by 1-D TMA into a two-slot ring, one matrix ahead, and each matrix serves both tiles;
//   * the two tiles are staggered: while the CUDA cores run the epilogue of one (bias, leaky-ReLU, mask, skip, pooling sums,
//     bf16 operand for the next GEMM, fp32 residual back to TMEM with tcgen05.st), the tensor core runs the GEMM of the other;
//   * fc_local2 accumulates straight onto X (the residual add is the accumulate flag), its bias rides on one more K-step
//     against a ones tile; fc_local1's bias is per jet (its global and context columns) and is added in the epilogue;
//   * local_0 with the embeddings folded in is ONE K-step: the operand row is [x_hi, x_lo, onehot(k)] (two bf16 per feature);
//   * the per-jet global path (masked mean / sum pooling -> global MLPs, epic.py:136-143, 187-190, 226-232) runs on the CUDA
//     cores for both jets at once, so each weight is fetched from L2 once per pair, while the fc_local1 GEMMs are in flight;
//   * the trunk skip (epic.py:148-155) is kept per tile as fp16 in shared memory.
// Numerics: bf16 operands, fp32 accumulation, fp32 residual / pooling / global path (bf16 weights for its three wide
// matrices).  Checked against the fp32 kernel with the tolerance written in tests/test_gpu_wide.py.
#include <cuda_bf16.h>
#include <cuda_fp16.h>

#include <stdlib.h>

#include <vector>

#include "mmb_device.cuh"
#include "mmb_internal.h"

namespace mmb {
namespace {

constexpr int kH = 128;
constexpr int kSlot = 36864;        // bytes per streamed matrix: 32 KB weight tile + 4 KB bias tile
constexpr int kThreads = 512;
constexpr int kCW = 32;             // accumulator columns per thread
constexpr int kMaxL = 8, kMaxG = 32, kMaxTX = 96, kMaxT = 64;
constexpr int kJ = 2;               // jets (tiles) in flight per CTA

// ---- PTX wrappers (same conventions as absorb_head_tc.cu) ---------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
    uint32_t ok;
    do {
        asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
                     : "=r"(ok) : "r"(bar), "r"(parity) : "memory");
    } while (!ok);
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void* src, uint32_t bytes, uint32_t bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}
__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
__device__ __forceinline__ void fence_proxy_async() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tmem_alloc(uint32_t slot_smem, uint32_t ncols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(slot_smem), "r"(ncols) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
    return pred != 0;
}
__device__ __forceinline__ void umma(uint32_t d_tmem, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
                 ::"r"(d_tmem), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate) : "memory");
}
__device__ __forceinline__ void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
// 32 consecutive fp32 columns of this thread's TMEM lane: load / store
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, float (&v)[32]) {
    uint32_t r[32];
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]), "=r"(r[9]),
          "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]),
          "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]),
          "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr) : "memory");
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
}
__device__ __forceinline__ void tmem_st32(uint32_t taddr, const float (&v)[32]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
        ::"r"(taddr), "r"(__float_as_uint(v[0])), "r"(__float_as_uint(v[1])), "r"(__float_as_uint(v[2])), "r"(__float_as_uint(v[3])),
          "r"(__float_as_uint(v[4])), "r"(__float_as_uint(v[5])), "r"(__float_as_uint(v[6])), "r"(__float_as_uint(v[7])),
          "r"(__float_as_uint(v[8])), "r"(__float_as_uint(v[9])), "r"(__float_as_uint(v[10])), "r"(__float_as_uint(v[11])),
          "r"(__float_as_uint(v[12])), "r"(__float_as_uint(v[13])), "r"(__float_as_uint(v[14])), "r"(__float_as_uint(v[15])),
          "r"(__float_as_uint(v[16])), "r"(__float_as_uint(v[17])), "r"(__float_as_uint(v[18])), "r"(__float_as_uint(v[19])),
          "r"(__float_as_uint(v[20])), "r"(__float_as_uint(v[21])), "r"(__float_as_uint(v[22])), "r"(__float_as_uint(v[23])),
          "r"(__float_as_uint(v[24])), "r"(__float_as_uint(v[25])), "r"(__float_as_uint(v[26])), "r"(__float_as_uint(v[27])),
          "r"(__float_as_uint(v[28])), "r"(__float_as_uint(v[29])), "r"(__float_as_uint(v[30])), "r"(__float_as_uint(v[31]))
        : "memory");
    asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ uint64_t smem_desc(uint32_t saddr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return (uint64_t)((saddr & 0x3FFFFu) >> 4) | ((uint64_t)((lbo_bytes >> 4) & 0x3FFFu) << 16) |
           ((uint64_t)((sbo_bytes >> 4) & 0x3FFFu) << 32) | ((uint64_t)1 << 46);
}
__host__ __device__ constexpr uint32_t instr_desc(int M, int N) {
    return (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
__device__ __forceinline__ uint32_t pack_bf16(float lo, float hi) {
    __nv_bfloat162 h = __floats2bfloat162_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ uint32_t pack_f16(float lo, float hi) {
    __half2 h = __floats2half2_rn(lo, hi);
    return *reinterpret_cast<uint32_t*>(&h);
}
// [128 rows x 128 k] bf16 tile, K-major canonical: 8-row groups 2048 B apart (SBO), 16-byte k-chunks 128 B apart (LBO)
__device__ __forceinline__ void store_row(uint8_t* tile, int row, int col, const float (&v)[32]) {
    uint8_t* p = tile + (row >> 3) * 2048 + (row & 7) * 16 + (col >> 3) * 128;
#pragma unroll
    for (int c = 0; c < 4; ++c)
        *reinterpret_cast<uint4*>(p + c * 128) = make_uint4(pack_bf16(v[8 * c], v[8 * c + 1]), pack_bf16(v[8 * c + 2], v[8 * c + 3]),
                                                            pack_bf16(v[8 * c + 4], v[8 * c + 5]), pack_bf16(v[8 * c + 6], v[8 * c + 7]));
}
// skip tile: fp16 [128 rows][128 cols], 256 B per row, 16-byte chunks XOR-swizzled by the row so that a quarter-warp (8 rows,
// same chunk number) touches 8 different bank groups
__device__ __forceinline__ uint4* skip_chunk(uint8_t* tile, int row, int chunk) {
    return reinterpret_cast<uint4*>(tile + row * 256 + ((chunk ^ (row & 7)) << 4));
}
// column sums over the 32 rows of a warp (recursive halving, 31 shuffles): on return lane l holds in vals[0] the total of
// column halving_index(l)
__device__ __forceinline__ void warp_halving_sum32(float (&vals)[32], int lane) {
    int off = 16;
#pragma unroll
    for (int cnt = 32; cnt > 1; cnt >>= 1, off >>= 1) {
        const bool upper = (lane & off) != 0;
#pragma unroll
        for (int i = 0; i < cnt / 2; ++i) {
            const float keep = upper ? vals[cnt / 2 + i] : vals[i];
            const float send = upper ? vals[i] : vals[cnt / 2 + i];
            vals[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
        }
    }
}
__device__ __forceinline__ int halving_index32(int lane) {
    int idx = 0, off = 16;
#pragma unroll
    for (int cnt = 32; cnt > 1; cnt >>= 1, off >>= 1) idx += (lane & off) ? cnt / 2 : 0;
    return idx;
}
__device__ __forceinline__ float lrelu(float a) { return a > 0.0f ? a : 0.01f * a; }
__device__ __forceinline__ float selu(float a) {
    const float alpha = 1.6732632423543772848170429916717f, scale = 1.0507009873554804934193349852946f;
    return scale * (a > 0.0f ? a : alpha * (__expf(a) - 1.0f));
}

__host__ __device__ inline int round8(int v) { return (v + 7) & ~7; }

// fp32 side table (floats) and bf16 wide matrices of the global path; every offset in elements of its array
struct WideLayout {
    int L, G, T, X, S, Sh, Dc, skip;
    int Tp, TXp, Gp;            // padded widths (multiples of 8, zero weights in the padding)
    // fp32 table
    int c0, w0t;                // [128], [128][Tp]: local_0 bias (+ folded embedding bias), its time columns
    int b_g0, b_g1, g2, b_g2;   // [128], [128], [Gp][128], [Gp]
    int layer0, layer_stride;   // per layer: b_lg1 [128] | Wg2 [Gp][128] | b_lg2 [Gp] | Wl1g [128][TXp + Gp] | b_l1 [128]
    int o_blg1, o_wg2, o_blg2, o_wl1g, o_bl1;
    int head0, b_head0, head2, b_head2;   // [16][16], [16], [16][16], [16]
    int tab_floats;
    // bf16 matrices
    int g0, g1;                 // [128][256 + TXp], [128][128]
    int wg1_0, wg1_stride;      // per layer [128][256 + TXp + Gp], columns [mean | sum | ctx | xg]
    int big_elems;
    int n_seq;                  // streamed matrices per evaluation: local_0, L x (fc_local1, fc_local2), output
    __host__ __device__ int K0() const { return 256 + TXp; }
    __host__ __device__ int K1() const { return 256 + TXp + Gp; }
    __host__ __device__ int K2() const { return TXp + Gp; }
};

WideLayout make_layout(const MmbEpicDims& d) {
    WideLayout w{};
    w.L = d.num_blocks; w.G = d.dim_hidden_glob; w.T = d.dim_time_emb; w.X = d.dim_context; w.S = d.vocab_size; w.Sh = d.disc_head_hidden;
    w.Dc = d.dim_continuous; w.skip = d.skip_connection;
    w.Tp = round8(w.T); w.TXp = round8(w.T + w.X); w.Gp = round8(w.G);
    int o = 0;
    auto take = [&](int n) { const int at = o; o += (n + 3) & ~3; return at; };
    w.c0 = take(128); w.w0t = take(128 * w.Tp);
    w.b_g0 = take(128); w.b_g1 = take(128); w.g2 = take(w.Gp * 128); w.b_g2 = take(w.Gp);
    w.layer0 = o;
    {
        int p = 0;
        auto tk = [&](int n) { const int at = p; p += (n + 3) & ~3; return at; };
        w.o_blg1 = tk(128); w.o_wg2 = tk(w.Gp * 128); w.o_blg2 = tk(w.Gp); w.o_wl1g = tk(128 * w.K2()); w.o_bl1 = tk(128);
        w.layer_stride = p;
    }
    o += w.layer_stride * w.L;
    w.head0 = take(256); w.b_head0 = take(16); w.head2 = take(256); w.b_head2 = take(16);
    w.tab_floats = o;
    int b = 0;
    auto tb = [&](int n) { const int at = b; b += (n + 7) & ~7; return at; };
    w.g0 = tb(128 * w.K0()); w.g1 = tb(128 * 128);
    w.wg1_0 = b; w.wg1_stride = (128 * w.K1() + 7) & ~7;
    b += w.wg1_stride * w.L;
    w.big_elems = b;
    w.n_seq = 2 + 2 * w.L;
    return w;
}

struct WideParams {
    const uint8_t* image;            // n_seq slots of kSlot bytes
    const float* tab;                // WideLayout fp32 table
    const __nv_bfloat16* big;        // WideLayout bf16 matrices
    WideLayout lay;
    const float* x;                  // [B,N,3]
    const uint8_t* k;                // [B,N]
    const uint8_t* mask;             // [B,N]
    const float* temb;               // [B or 1][T + X]
    int temb_stride;
    int B, N;
    float* v_out;                    // [B,N,3]
    float* logits_out;               // [B,N,S]
    float* hidden_out;               // [B,N,128] or null
    long long* trace;                // debug: clock64() stamps of the first pair of CTA 0 (tools/wide_trace.py); null in production
};

// dynamic shared memory
constexpr int kOffRing = 0;
constexpr int kOffA = 2 * kSlot;                    // two operand tiles
constexpr int kOffSkip = kOffA + 2 * 32768;         // two fp16 skip tiles
constexpr int kOffOnes = kOffSkip + 2 * 32768;
constexpr int kOffVec = kOffOnes + 4096;            // per-jet vectors (floats), see Vec
struct Vec {   // float offsets inside the vector area
    static constexpr int in_ld = 256 + kMaxTX + kMaxG;          // [mean 128 | sum 128 | ctx TXp | xg Gp]
    static constexpr int in = 0;                                 // [kJ][in_ld]
    static constexpr int g1 = in + kJ * in_ld;                   // [kJ][128]
    static constexpr int g0 = g1 + kJ * 128;                     // [kJ][128]
    static constexpr int bl1 = g0 + kJ * 128;                    // [kJ][128]
    static constexpr int tv0 = bl1 + kJ * 128;                   // [kJ][128]
    static constexpr int cx = tv0 + kJ * 128;                    // [kJ][kMaxTX + kMaxG]: [ctx | xm]
    static constexpr int skg = cx + kJ * (kMaxTX + kMaxG);       // [kJ][kMaxG]
    static constexpr int part = skg + kJ * kMaxG;                // [kJ][16 warps][32]
    static constexpr int head = part + kJ * 16 * 32;             // head0 [16][16] | b0 [16] | head2 [16][16] | b2 [16] | b_out... (copied once)
    static constexpr int floats = head + 256 + 16 + 256 + 16;
};
constexpr int kSmemBytes = kOffVec + Vec::floats * 4;
static_assert(kSmemBytes <= 227 * 1024, "shared memory budget");

__global__ void __launch_bounds__(kThreads, 1) epic_wide_kernel(const WideParams p) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t s_tmem_slot;
    __shared__ __align__(8) uint64_t s_bars[4];      // full[0], full[1], mma[A], mma[B]
    __shared__ int s_cnt[kJ];
    const WideLayout& ly = p.lay;
    const int tid = threadIdx.x, r = tid & 127, cq = tid >> 7, warp = __shfl_sync(0xffffffffu, tid >> 5, 0), lane = tid & 31;
    const int qq = warp & 3;
    const int L = ly.L, G = ly.G, Gp = ly.Gp, TXp = ly.TXp, TX = ly.T + ly.X, S = ly.S, Sh = ly.Sh;
    const bool skip_on = ly.skip != 0;
    uint8_t* sOnes = smem + kOffOnes;
    float* sv = reinterpret_cast<float*>(smem + kOffVec);
    auto sA = [&](int t) { return smem + kOffA + t * 32768; };
    auto sSkip = [&](int t) { return smem + kOffSkip + t * 32768; };

    for (int i = tid; i < 256; i += kThreads) reinterpret_cast<uint4*>(sOnes)[i] = make_uint4(0x3F803F80u, 0x3F803F80u, 0x3F803F80u, 0x3F803F80u);
    for (int i = tid; i < 256 + 16 + 256 + 16; i += kThreads) sv[Vec::head + i] = __ldg(p.tab + ly.head0 + i);
    const uint32_t bar_full0 = smem_u32(&s_bars[0]), bar_full1 = smem_u32(&s_bars[1]);
    const uint32_t bar_mma[2] = {smem_u32(&s_bars[2]), smem_u32(&s_bars[3])};
    if (tid == 0) { mbar_init(bar_full0, 1); mbar_init(bar_full1, 1); mbar_init(bar_mma[0], 1); mbar_init(bar_mma[1], 1); }
    if (warp == 0) tmem_alloc(smem_u32(&s_tmem_slot), 512);
    fence_barrier_init();
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = __shfl_sync(0xffffffffu, s_tmem_slot, 0);
    const uint32_t lane_off = ((uint32_t)(qq * 32) << 16);
    const int col0 = cq * kCW;
    const uint32_t dX[2] = {tmem, tmem + 256}, dACC[2] = {tmem + 128, tmem + 384};

    const int n_pairs = (p.B + 1) / 2;
    const int my_pairs = n_pairs > (int)blockIdx.x ? (n_pairs - 1 - (int)blockIdx.x) / (int)gridDim.x + 1 : 0;
    const uint32_t total_mats = (uint32_t)my_pairs * ly.n_seq;
    const uint32_t wbase = smem_u32(smem + kOffRing);
    auto issue_load = [&](uint32_t i) {   // thread 0 only
        const uint32_t bar = (i & 1) ? bar_full1 : bar_full0;
        mbar_expect_tx(bar, kSlot);
        bulk_g2s(wbase + (i & 1) * kSlot, p.image + (size_t)(i % ly.n_seq) * kSlot, kSlot, bar);
    };
    if (tid == 0) {
        if (total_mats > 0) issue_load(0);
        if (total_mats > 1) issue_load(1);
    }
    uint32_t wseq = 0;                 // streamed matrix the next GEMM phase consumes
    uint32_t mma_phase[2] = {0, 0};
    constexpr uint32_t idesc128 = instr_desc(128, 128), idesc16 = instr_desc(128, 16);
    const uint64_t ones_desc = smem_desc(smem_u32(sOnes), 128, 256);

    // D (+)= A[128 x 16 nk] W^T with streamed matrix `ws` (+ bias K-step); commit to the tile's barrier.  One elected lane of warp 0.
    auto gemm = [&](int t, uint32_t ws, uint32_t d, int nk, uint32_t idesc, bool accumulate, bool bias) {
        if (warp == 0) {
            const uint32_t w = __shfl_sync(0xffffffffu, ws, 0);
            if (elect_one()) {
                const uint32_t wb = wbase + (w & 1) * kSlot;
                mbar_wait((w & 1) ? bar_full1 : bar_full0, (w >> 1) & 1);
                tc_fence_after();
                const uint64_t ad = smem_desc(smem_u32(sA(t)), 128, 2048), wd = smem_desc(wb, 128, 2048);
#pragma unroll 8
                for (int j = 0; j < nk; ++j) umma(d, ad + (uint64_t)(j * 16), wd + (uint64_t)(j * 16), idesc, (accumulate || j > 0) ? 1u : 0u);
                if (bias) umma(d, ones_desc, smem_desc(wb + 32768, 128, 256), idesc, 1u);
                umma_commit(bar_mma[t]);
            }
            __syncwarp();
        }
    };
    auto wait_tile = [&](int t) {
        mbar_wait(bar_mma[t], mma_phase[t]); mma_phase[t] ^= 1;
        tc_fence_after();
    };
    // matrix `wseq` has served both tiles: its ring slot is free -> prefetch the matrix two ahead
    auto matrix_done = [&]() {
        if (tid == 0 && wseq + 2 < total_mats) issue_load(wseq + 2);
        ++wseq;
    };
    auto publish = [&]() {   // operand tile / TMEM writes of all threads -> visible to the MMA proxy, then the block barrier
        tc_fence_before();
        fence_proxy_async();
        __syncthreads();
    };

    // ---- global-path helpers: out[j][o] for both jets at once; thread (o = tid >> 2, kq = tid & 3) covers 16-byte weight chunks
    // kq, kq + 4, ... of row o; the four partial sums meet by shuffle.  W row-major [n_out][K], K a multiple of 8.  The weights
    // do not depend on the data, so a stage's chunks are fetched into registers up front (all loads in flight at once: one L2
    // latency per stage instead of one per chunk) — `load_*` may be issued before the barrier that publishes the stage's input.
    struct RowB { uint4 w[12]; };    // bf16 rows: K <= 384
    struct RowF { float4 w[8]; };    // fp32 rows: K <= 128
    auto load_rows_bf16 = [&](const __nv_bfloat16* W, int K, RowB& rw) {
        const uint4* row = reinterpret_cast<const uint4*>(W + (size_t)(tid >> 2) * K);
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int c = (tid & 3) + 4 * i;
            rw.w[i] = c < K / 8 ? __ldg(row + c) : make_uint4(0u, 0u, 0u, 0u);
        }
    };
    auto dot_rows_bf16 = [&](const RowB& rw, int K, const float* in, int in_ld, float (&acc)[kJ]) {
#pragma unroll
        for (int j = 0; j < kJ; ++j) acc[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < 12; ++i) {
            const int c = (tid & 3) + 4 * i;
            if (c < K / 8) {
                const uint32_t ww[4] = {rw.w[i].x, rw.w[i].y, rw.w[i].z, rw.w[i].w};
#pragma unroll
                for (int j = 0; j < kJ; ++j) {
                    const float4 a = *reinterpret_cast<const float4*>(in + j * in_ld + 8 * c), b = *reinterpret_cast<const float4*>(in + j * in_ld + 8 * c + 4);
                    const float xs[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                        acc[j] = fmaf(__uint_as_float(ww[e] << 16), xs[2 * e], acc[j]);
                        acc[j] = fmaf(__uint_as_float(ww[e] & 0xffff0000u), xs[2 * e + 1], acc[j]);
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kJ; ++j) {
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
        }
    };
    auto load_rows_f32 = [&](const float* W, int K, int n_out, RowF& rw) {
        const float4* row = reinterpret_cast<const float4*>(W + (size_t)(tid >> 2) * K);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = (tid & 3) + 4 * i;
            rw.w[i] = ((tid >> 2) < n_out && c < K / 4) ? __ldg(row + c) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
    };
    auto dot_rows_f32 = [&](const RowF& rw, int K, const float* in, int in_ld, float (&acc)[kJ]) {
#pragma unroll
        for (int j = 0; j < kJ; ++j) acc[j] = 0.0f;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            const int c = (tid & 3) + 4 * i;
            if (c < K / 4) {
#pragma unroll
                for (int j = 0; j < kJ; ++j) {
                    const float4 a = *reinterpret_cast<const float4*>(in + j * in_ld + 4 * c);
                    acc[j] = fmaf(rw.w[i].x, a.x, acc[j]); acc[j] = fmaf(rw.w[i].y, a.y, acc[j]);
                    acc[j] = fmaf(rw.w[i].z, a.z, acc[j]); acc[j] = fmaf(rw.w[i].w, a.w, acc[j]);
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kJ; ++j) {
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 1);
            acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], 2);
        }
    };
    float* s_in = sv + Vec::in;
    // pooled sums of both tiles -> [mean | sum] of the jets' input vectors (epic.py:136-143)
    auto pool_to_input = [&]() {
        if (tid < kJ * 128) {
            const int j = tid >> 7, c = tid & 127;
            const float* part = sv + Vec::part + j * 16 * 32 + (c >> 5) * 4 * 32 + (c & 31);   // warps 4 cq .. 4 cq + 3 hold column c of rows 32 qq ..
            const float s = (part[0] + part[32]) + (part[64] + part[96]);
            s_in[j * Vec::in_ld + c] = s / (float)s_cnt[j];     // 0 / 0 = NaN for an empty jet, as the reference (epic.py:141)
            s_in[j * Vec::in_ld + 128 + c] = s;
        }
    };

#define WIDE_TRACE(id) do { if (p.trace && blockIdx.x == 0 && pi == 0 && tid == 0) p.trace[id] = clock64(); } while (0)
    for (int pi = 0; pi < my_pairs; ++pi) {
        const int pair = (int)blockIdx.x + pi * (int)gridDim.x;
        WIDE_TRACE(0);
        const int jet0 = 2 * pair;
        const bool has_b = jet0 + 1 < p.B;
        bool live[2] = {false, false};
        // ---- load the pair: masks, first operand rows [x_hi, x_lo, onehot(k)], context vectors --------------------------------
#pragma unroll
        for (int t = 0; t < kJ; ++t) {
            const int jet = jet0 + t;
            const bool on = jet < p.B;
            live[t] = on && r < p.N && p.mask[(size_t)jet * p.N + r] != 0;
            const int cnt = __syncthreads_count(cq == 0 && live[t]);
            if (tid == 0) s_cnt[t] = cnt;
            if (cq == 0) {
                uint32_t w[8] = {0, 0, 0, 0, 0, 0, 0, 0};
                if (live[t]) {
                    const float* xr = p.x + ((size_t)jet * p.N + r) * 3;
                    const float x0 = xr[0], x1 = xr[1], x2 = xr[2];
                    const float h0 = __bfloat162float(__float2bfloat16(x0)), h1 = __bfloat162float(__float2bfloat16(x1)),
                                h2 = __bfloat162float(__float2bfloat16(x2));
                    const int kk = p.k[(size_t)jet * p.N + r];
                    w[0] = pack_bf16(h0, h1); w[1] = pack_bf16(h2, x0 - h0); w[2] = pack_bf16(x1 - h1, x2 - h2);
                    const int pos = 6 + kk;   // onehot column
                    w[pos >> 1] |= 0x3F80u << (16 * (pos & 1));
                }
                uint8_t* q = sA(t) + (r >> 3) * 2048 + (r & 7) * 16;
                *reinterpret_cast<uint4*>(q) = make_uint4(w[0], w[1], w[2], w[3]);
                *reinterpret_cast<uint4*>(q + 128) = make_uint4(w[4], w[5], w[6], w[7]);
            } else if (cq == 1) {   // context vector [time embedding | embedded context], zero padded
                for (int i = r; i < kMaxTX; i += 128) {
                    const float v = (on && i < TX) ? __ldg(p.temb + (size_t)jet * p.temb_stride + i) : 0.0f;
                    s_in[t * Vec::in_ld + 256 + i] = v;
                    sv[Vec::cx + t * (kMaxTX + kMaxG) + i] = v;
                }
            }
        }
        publish();
        WIDE_TRACE(1);
        // ---- local_0 (one K-step) for both tiles; meanwhile the per-jet time vector of its bias
        gemm(0, wseq, dX[0], 1, idesc128, false, false);
        if (has_b) gemm(1, wseq, dX[1], 1, idesc128, false, false);
        {
            float acc[kJ];
            RowF rw;
            load_rows_f32(p.tab + ly.w0t, ly.Tp, 128, rw);
            dot_rows_f32(rw, ly.Tp, s_in + 256, Vec::in_ld, acc);
            if ((tid & 3) == 0) {
                const int o = tid >> 2;
#pragma unroll
                for (int j = 0; j < kJ; ++j) sv[Vec::tv0 + j * 128 + o] = acc[j] + __ldg(p.tab + ly.c0 + o);
            }
        }
        __syncthreads();
        // epilogue of a particle Linear: bias -> leaky-ReLU (-> mask, + skip) -> bf16 operand tile (and fp32 X, pooling sums)
        // kind 0: local_0 (X = lrelu(X + tv0) * mask; defines the skip); 1: fc_local1 (A = lrelu(ACC + bl1)); 2: fc_local2
        // (X = lrelu(X) * mask + skip)
        auto epilogue = [&](int t, int kind, bool last) {
            wait_tile(t);
            const bool warp_live = __any_sync(0xffffffffu, live[t]);
            float v[32];
            if (warp_live) {
                tmem_ld32((kind == 1 ? dACC[t] : dX[t]) + lane_off + col0, v);
                if (kind != 2) {
                    const float* b = sv + (kind == 0 ? Vec::tv0 : Vec::bl1) + t * 128 + col0;
#pragma unroll
                    for (int j = 0; j < 32; j += 4) {
                        const float4 b4 = *reinterpret_cast<const float4*>(b + j);
                        v[j] += b4.x; v[j + 1] += b4.y; v[j + 2] += b4.z; v[j + 3] += b4.w;
                    }
                }
#pragma unroll
                for (int j = 0; j < 32; ++j) v[j] = lrelu(v[j]);
                if (kind != 1) {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = live[t] ? v[j] : 0.0f;   // select: a dead row may hold anything
                    if (skip_on) {
                        if (kind == 0) {
#pragma unroll
                            for (int c = 0; c < 4; ++c)
                                *skip_chunk(sSkip(t), r, 4 * cq + c) = make_uint4(pack_f16(v[8 * c], v[8 * c + 1]), pack_f16(v[8 * c + 2], v[8 * c + 3]),
                                                                                  pack_f16(v[8 * c + 4], v[8 * c + 5]), pack_f16(v[8 * c + 6], v[8 * c + 7]));
                        } else {
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                const uint4 s4 = *skip_chunk(sSkip(t), r, 4 * cq + c);
                                const uint32_t sw[4] = {s4.x, s4.y, s4.z, s4.w};
#pragma unroll
                                for (int i = 0; i < 4; ++i) {
                                    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&sw[i]));
                                    v[8 * c + 2 * i] += f.x; v[8 * c + 2 * i + 1] += f.y;
                                }
                            }
                        }
                    }
                    tmem_st32(dX[t] + lane_off + col0, v);
                    if (last && p.hidden_out && live[t]) {   // the last local hidden (EPiCWrapper.forward(output_hidden_local=True), epic.py:159-160)
                        float* h = p.hidden_out + ((size_t)(jet0 + t) * p.N + r) * kH + col0;
#pragma unroll
                        for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(h + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    }
                }
                store_row(sA(t), r, col0, v);
                if (kind != 1) {   // masked column sums for the next pooling (dead rows are zero already)
                    warp_halving_sum32(v, lane);
                    sv[Vec::part + t * 16 * 32 + warp * 32 + halving_index32(lane)] = v[0];
                }
            } else if (kind != 1) {
                sv[Vec::part + t * 16 * 32 + warp * 32 + lane] = 0.0f;
            }
            if (kind != 1 && last && p.hidden_out && !live[t] && jet0 + t < p.B && r < p.N) {
                // x * mask: zero, except in a jet without particles, whose mean pool is 0 / 0 (epic.py:141) and NaN * 0 = NaN
                const float z = s_cnt[t] == 0 ? __int_as_float(0x7fc00000) : 0.0f;
                float* h = p.hidden_out + ((size_t)(jet0 + t) * p.N + r) * kH + col0;
#pragma unroll
                for (int j = 0; j < 32; j += 4) *reinterpret_cast<float4*>(h + j) = make_float4(z, z, z, z);
            }
        };
        WIDE_TRACE(2);
        epilogue(0, 0, false);
        WIDE_TRACE(3);
        publish();
        if (L > 0) gemm(0, wseq + 1, dACC[0], 8, idesc128, false, false);           // fc_local1 of layer 0, tile A
        if (has_b) epilogue(1, 0, false);
        publish();
        WIDE_TRACE(4);
        matrix_done();                                                               // local_0 served both tiles
        if (has_b && L > 0) gemm(1, wseq, dACC[1], 8, idesc128, false, false);
        // ---- EPiC_Projection globals (epic.py:187-190)
        {
            float acc[kJ];
            RowB rb;
            load_rows_bf16(p.big + ly.g0, ly.K0(), rb);
            pool_to_input();
            __syncthreads();
            dot_rows_bf16(rb, ly.K0(), s_in, Vec::in_ld, acc);
            load_rows_bf16(p.big + ly.g1, 128, rb);
            RowF rf;
            load_rows_f32(p.tab + ly.g2, 128, Gp, rf);
            if ((tid & 3) == 0)
#pragma unroll
                for (int j = 0; j < kJ; ++j) sv[Vec::g0 + j * 128 + (tid >> 2)] = lrelu(acc[j] + __ldg(p.tab + ly.b_g0 + (tid >> 2)));
            __syncthreads();
            dot_rows_bf16(rb, 128, sv + Vec::g0, 128, acc);
            if ((tid & 3) == 0)
#pragma unroll
                for (int j = 0; j < kJ; ++j) sv[Vec::g1 + j * 128 + (tid >> 2)] = lrelu(acc[j] + __ldg(p.tab + ly.b_g1 + (tid >> 2)));
            __syncthreads();
            dot_rows_f32(rf, 128, sv + Vec::g1, 128, acc);
            if ((tid & 3) == 0 && (tid >> 2) < Gp) {
                const int o = tid >> 2;
#pragma unroll
                for (int j = 0; j < kJ; ++j) {
                    const float xg = o < G ? lrelu(acc[j] + __ldg(p.tab + ly.b_g2 + o)) : 0.0f;
                    s_in[j * Vec::in_ld + 256 + TXp + o] = xg;
                    sv[Vec::skg + j * kMaxG + o] = skip_on ? xg : 0.0f;
                }
            }
            __syncthreads();
        }
        WIDE_TRACE(5);
        // ---- EPiC layers (epic.py:217-241, 152-155)
        for (int l = 0; l < L; ++l) {
            if (l == 1) WIDE_TRACE(6);
            const float* tl = p.tab + ly.layer0 + (size_t)l * ly.layer_stride;
            {   // per-jet path: fc_global1 -> fc_global2 (+ residual) -> the per-jet part of fc_local1
                float acc[kJ];
                RowB rb;
                RowF rf2, rf3;
                load_rows_bf16(p.big + ly.wg1_0 + (size_t)l * ly.wg1_stride, ly.K1(), rb);
                load_rows_f32(tl + ly.o_wg2, 128, Gp, rf2);
                load_rows_f32(tl + ly.o_wl1g, ly.K2(), 128, rf3);
                if (l > 0) {
                    pool_to_input();
                    __syncthreads();
                }
                dot_rows_bf16(rb, ly.K1(), s_in, Vec::in_ld, acc);
                if ((tid & 3) == 0)
#pragma unroll
                    for (int j = 0; j < kJ; ++j) sv[Vec::g1 + j * 128 + (tid >> 2)] = lrelu(acc[j] + __ldg(tl + ly.o_blg1 + (tid >> 2)));
                __syncthreads();
                dot_rows_f32(rf2, 128, sv + Vec::g1, 128, acc);
                if ((tid & 3) == 0 && (tid >> 2) < Gp) {
                    const int o = tid >> 2;
#pragma unroll
                    for (int j = 0; j < kJ; ++j) {
                        const float xm = o < G ? lrelu(acc[j] + __ldg(tl + ly.o_blg2 + o) + s_in[j * Vec::in_ld + 256 + TXp + o]) : 0.0f;
                        sv[Vec::cx + j * (kMaxTX + kMaxG) + TXp + o] = xm;                               // fc_local1 sees the layer's own output
                        s_in[j * Vec::in_ld + 256 + TXp + o] = xm + sv[Vec::skg + j * kMaxG + o];        // the next layer the skipped one (epic.py:155)
                    }
                }
                __syncthreads();
                dot_rows_f32(rf3, ly.K2(), sv + Vec::cx, kMaxTX + kMaxG, acc);
                if ((tid & 3) == 0)
#pragma unroll
                    for (int j = 0; j < kJ; ++j) sv[Vec::bl1 + j * 128 + (tid >> 2)] = acc[j] + __ldg(tl + ly.o_bl1 + (tid >> 2));
                __syncthreads();
            }
            const bool last = l == L - 1;
            if (l == 1) WIDE_TRACE(7);
            epilogue(0, 1, false);
            if (l == 1) WIDE_TRACE(8);
            publish();
            gemm(0, wseq + 1, dX[0], 8, idesc128, true, true);                      // fc_local2 accumulates onto X
            if (l == 1) WIDE_TRACE(9);
            if (has_b) epilogue(1, 1, false);
            if (l == 1) WIDE_TRACE(10);
            publish();
            matrix_done();                                                           // fc_local1 served both tiles
            if (has_b) gemm(1, wseq, dX[1], 8, idesc128, true, true);
            if (l == 1) WIDE_TRACE(11);
            epilogue(0, 2, last);
            if (l == 1) WIDE_TRACE(12);
            publish();
            gemm(0, wseq + 1, dACC[0], 8, last ? idesc16 : idesc128, false, last);  // next fc_local1, or the output layer
            if (has_b) epilogue(1, 2, last);
            publish();
            matrix_done();                                                           // fc_local2 served both tiles
            if (has_b) gemm(1, wseq, dACC[1], 8, last ? idesc16 : idesc128, false, last);
            if (l == 1) WIDE_TRACE(13);
        }
        WIDE_TRACE(14);
        // ---- output layer + discrete head (epic.py:158-162, mbm.py:105-113): h = out(x) * mask, logits = head(h[3:])
#pragma unroll
        for (int t = 0; t < kJ; ++t) {
            if (t == 1 && !has_b) continue;
            wait_tile(t);
            if (cq == 0) {   // whole warps: tcgen05.ld is warp-collective
                float v[32];
                tmem_ld32(dACC[t] + lane_off, v);
                if (r < p.N) {
                    const size_t pidx = (size_t)(jet0 + t) * p.N + r;
                    float h[16];
                    const float z = s_cnt[t] == 0 ? __int_as_float(0x7fc00000) : 0.0f;   // h * mask; NaN * 0 in a jet without particles
#pragma unroll
                    for (int i = 0; i < 16; ++i) h[i] = live[t] ? v[i] : z;
#pragma unroll
                    for (int c = 0; c < 3; ++c) p.v_out[pidx * 3 + c] = h[c];
                    const float* hd = sv + Vec::head;
                    float z1[16], lg[8];
#pragma unroll
                    for (int o = 0; o < 16; ++o) {   // head Linear 0 + SELU (rows >= Sh are zero in the table and unused)
                        float a = hd[256 + o];
#pragma unroll
                        for (int s = 0; s < 8; ++s) a = fmaf(hd[o * 16 + s], h[3 + s], a);
                        z1[o] = selu(a);
                    }
#pragma unroll
                    for (int o = 0; o < 8; ++o) {
                        float a = hd[256 + 16 + 256 + o];
#pragma unroll
                        for (int s = 0; s < 16; ++s) a = fmaf(hd[256 + 16 + o * 16 + s], z1[s], a);
                        lg[o] = Sh ? a : h[3 + o];
                    }
#pragma unroll
                    for (int o = 0; o < 8; ++o)
                        if (o < S) p.logits_out[pidx * S + o] = lg[o];
                }
            }
        }
        tc_fence_before();
        __syncthreads();
        WIDE_TRACE(15);
        matrix_done();   // output layer served both tiles
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem, 512);
}

__device__ long long g_wide_trace[32];

// element (row o, k) of a K-major tile with 8-row groups `sbo` bytes apart -> bf16 index
inline size_t tile_index(int o, int k, int sbo) { return ((size_t)(o / 8) * sbo + (size_t)(k / 8) * 128 + (o % 8) * 16 + (k % 8) * 2) / 2; }

}  // namespace

struct WideImage {
    void* image = nullptr;
    float* tab = nullptr;
    __nv_bfloat16* big = nullptr;
    WideLayout lay;
};

bool wide_supported(const MmbEpicDims* d, int N) {
    return d->dim_hidden_local == kH && d->dim_hidden_glob >= 1 && d->dim_hidden_glob <= kMaxG && d->dim_time_emb >= 1 && d->dim_time_emb <= kMaxT &&
           d->dim_time_emb + d->dim_context <= kMaxTX && d->dim_context >= 0 && d->num_blocks >= 1 && d->num_blocks <= kMaxL &&
           d->dim_continuous == 3 && d->vocab_size >= 1 && d->vocab_size <= 8 && d->disc_head_hidden >= 0 && d->disc_head_hidden <= 16 &&
           N >= 1 && N <= 128;
}

void wide_free_image(EpicModel* m) {
    WideImage* w = static_cast<WideImage*>(m->wide);
    if (!w) return;
    if (w->image) cudaFree(w->image);
    if (w->tab) cudaFree(w->tab);
    if (w->big) cudaFree(w->big);
    delete w;
    m->wide = nullptr;
}

int wide_build_image(EpicModel* m, const float* W) {
    const MmbEpicDims& d = m->dims;
    const MmbEpicLayout& Lo = m->layout;
    const WideLayout ly = make_layout(d);
    const int T = d.dim_time_emb, X = d.dim_context, TX = T + X, C = d.dim_cont_emb, D = d.dim_disc_emb, G = d.dim_hidden_glob, L = d.num_blocks,
              S = d.vocab_size, Sh = d.disc_head_hidden, Dc = d.dim_continuous, H = kH;
    const int K0 = T + C + D;
    std::vector<__nv_bfloat16> img((size_t)ly.n_seq * kSlot / 2, __float2bfloat16(0.0f));
    std::vector<float> tab((size_t)ly.tab_floats, 0.0f);
    std::vector<__nv_bfloat16> big((size_t)ly.big_elems, __float2bfloat16(0.0f));
    auto put = [&](int slot, int o, int k, double v) { img[(size_t)slot * kSlot / 2 + tile_index(o, k, 2048)] = __float2bfloat16((float)v); };
    auto put_bias = [&](int slot, int o, float b) {
        const __nv_bfloat16 hi = __float2bfloat16(b);
        img[(size_t)slot * kSlot / 2 + 32768 / 2 + tile_index(o, 0, 256)] = hi;
        img[(size_t)slot * kSlot / 2 + 32768 / 2 + tile_index(o, 1, 256)] = __float2bfloat16(b - __bfloat162float(hi));
    };
    // local_0 with the embeddings folded in: operand columns [x_hi (3) | x_lo (3) | onehot (S)]
    for (int o = 0; o < H; ++o) {
        const float* w0 = W + Lo.local0_w + (size_t)o * K0;
        for (int j = 0; j < Dc; ++j) {
            double acc = 0;
            for (int c = 0; c < C; ++c) acc += (double)w0[T + c] * W[Lo.emb_cont_w + (size_t)c * Dc + j];
            put(0, o, j, acc);
            put(0, o, Dc + j, acc);
        }
        for (int s = 0; s < S; ++s) {
            double acc = 0;
            for (int dd = 0; dd < D; ++dd) acc += (double)w0[T + C + dd] * W[Lo.emb_disc + (size_t)s * D + dd];
            put(0, o, 2 * Dc + s, acc);
        }
        double c0 = W[Lo.local0_b + o];
        for (int c = 0; c < C; ++c) c0 += (double)w0[T + c] * W[Lo.emb_cont_b + c];
        tab[ly.c0 + o] = (float)c0;
        for (int t = 0; t < T; ++t) tab[ly.w0t + (size_t)o * ly.Tp + t] = w0[t];
    }
    // projection globals: global_0 [H][mean | sum | ctx], global_1, global_2
    for (int o = 0; o < H; ++o) {
        const float* g0 = W + Lo.global0_w + (size_t)o * (2 * H + TX);
        for (int k = 0; k < 2 * H + TX; ++k) big[ly.g0 + (size_t)o * ly.K0() + k] = __float2bfloat16(g0[k]);
        for (int k = 0; k < H; ++k) big[ly.g1 + (size_t)o * 128 + k] = __float2bfloat16(W[Lo.global1_w + (size_t)o * H + k]);
        tab[ly.b_g0 + o] = W[Lo.global0_b + o];
        tab[ly.b_g1 + o] = W[Lo.global1_b + o];
    }
    for (int o = 0; o < G; ++o) {
        for (int k = 0; k < H; ++k) tab[ly.g2 + (size_t)o * 128 + k] = W[Lo.global2_w + (size_t)o * H + k];
        tab[ly.b_g2 + o] = W[Lo.global2_b + o];
    }
    for (int l = 0; l < L; ++l) {
        const float* Wl = W + Lo.layer0 + (size_t)l * Lo.layer_stride;
        float* tl = tab.data() + ly.layer0 + (size_t)l * ly.layer_stride;
        __nv_bfloat16* wg1 = big.data() + ly.wg1_0 + (size_t)l * ly.wg1_stride;
        const int Kg = 2 * H + G + TX, Kl = H + G + TX;
        for (int o = 0; o < H; ++o) {
            const float* g1 = Wl + Lo.l_g1_w + (size_t)o * Kg;       // reference columns [mean | sum | xg | ctx] -> [mean | sum | ctx | xg]
            for (int k = 0; k < 2 * H; ++k) wg1[(size_t)o * ly.K1() + k] = __float2bfloat16(g1[k]);
            for (int k = 0; k < TX; ++k) wg1[(size_t)o * ly.K1() + 256 + k] = __float2bfloat16(g1[2 * H + G + k]);
            for (int k = 0; k < G; ++k) wg1[(size_t)o * ly.K1() + 256 + ly.TXp + k] = __float2bfloat16(g1[2 * H + k]);
            tl[ly.o_blg1 + o] = Wl[Lo.l_g1_b + o];
            const float* l1 = Wl + Lo.l_l1_w + (size_t)o * Kl;       // [local H | xg | ctx]: the local part streams as matrix 1 + 2l
            for (int k = 0; k < H; ++k) put(1 + 2 * l, o, k, l1[k]);
            for (int k = 0; k < TX; ++k) tl[ly.o_wl1g + (size_t)o * ly.K2() + k] = l1[H + G + k];
            for (int k = 0; k < G; ++k) tl[ly.o_wl1g + (size_t)o * ly.K2() + ly.TXp + k] = l1[H + k];
            tl[ly.o_bl1 + o] = Wl[Lo.l_l1_b + o];
            for (int k = 0; k < H; ++k) put(2 + 2 * l, o, k, Wl[Lo.l_l2_w + (size_t)o * H + k]);
            put_bias(2 + 2 * l, o, Wl[Lo.l_l2_b + o]);
        }
        for (int o = 0; o < G; ++o) {
            for (int k = 0; k < H; ++k) tl[ly.o_wg2 + (size_t)o * 128 + k] = Wl[Lo.l_g2_w + (size_t)o * H + k];
            tl[ly.o_blg2 + o] = Wl[Lo.l_g2_b + o];
        }
    }
    for (int o = 0; o < Dc + S; ++o) {
        for (int k = 0; k < H; ++k) put(1 + 2 * L, o, k, W[Lo.out_w + (size_t)o * H + k]);
        put_bias(1 + 2 * L, o, W[Lo.out_b + o]);
    }
    if (Sh) {
        for (int o = 0; o < Sh; ++o) {
            for (int s = 0; s < S; ++s) tab[ly.head0 + o * 16 + s] = W[Lo.head0_w + (size_t)o * S + s];
            tab[ly.b_head0 + o] = W[Lo.head0_b + o];
        }
        for (int o = 0; o < S; ++o) {
            for (int s = 0; s < Sh; ++s) tab[ly.head2 + o * 16 + s] = W[Lo.head2_w + (size_t)o * Sh + s];
            tab[ly.b_head2 + o] = W[Lo.head2_b + o];
        }
    }
    WideImage* w = new WideImage();
    w->lay = ly;
    m->wide = w;
    int rc = cuda_ok(cudaMalloc(&w->image, img.size() * 2), "cudaMalloc wide image");
    if (!rc) rc = cuda_ok(cudaMalloc(&w->tab, tab.size() * 4), "cudaMalloc wide table");
    if (!rc) rc = cuda_ok(cudaMalloc(&w->big, big.size() * 2), "cudaMalloc wide global matrices");
    if (!rc) rc = cuda_ok(cudaMemcpy(w->image, img.data(), img.size() * 2, cudaMemcpyHostToDevice), "wide image upload");
    if (!rc) rc = cuda_ok(cudaMemcpy(w->tab, tab.data(), tab.size() * 4, cudaMemcpyHostToDevice), "wide table upload");
    if (!rc) rc = cuda_ok(cudaMemcpy(w->big, big.data(), big.size() * 2, cudaMemcpyHostToDevice), "wide matrices upload");
    if (rc) wide_free_image(m);
    return rc;
}

int wide_read_trace(long long* out, int n) {
    cudaDeviceSynchronize();
    cudaMemcpyFromSymbol(out, g_wide_trace, sizeof(long long) * (n < 32 ? n : 32));
    return n < 32 ? n : 32;
}

int launch_epic_forward_wide(const EpicModel* m, const float* x, const uint8_t* k, const uint8_t* mask, const float* temb, int temb_stride,
                             int B, int N, float* v_out, float* logits_out, float* hidden_out, cudaStream_t stream) {
    const WideImage* w = static_cast<const WideImage*>(m->wide);
    if (!w) return fail(MMB_EUNSUPPORTED, "wide EPiC trunk: no operand image for this model");
    if (B == 0 || N == 0) return MMB_OK;
    if (N > 128) return fail(MMB_EUNSUPPORTED, "wide EPiC trunk handles up to 128 particle slots per jet (got %d)", N);
    WideParams p{};
    p.image = static_cast<const uint8_t*>(w->image); p.tab = w->tab; p.big = w->big; p.lay = w->lay;
    p.x = x; p.k = k; p.mask = mask; p.temb = temb; p.temb_stride = temb_stride; p.B = B; p.N = N;
    p.v_out = v_out; p.logits_out = logits_out; p.hidden_out = hidden_out;
    static const bool trace_on = [] { const char* e = getenv("MMB_WIDE_TRACE"); return e && e[0] == '1'; }();   // debug knob
    if (trace_on) cudaGetSymbolAddress(reinterpret_cast<void**>(&p.trace), g_wide_trace);
    if (int rc = cuda_ok(cudaFuncSetAttribute(epic_wide_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmemBytes), "wide smem attribute"))
        return rc;
    const int pairs = (B + 1) / 2;
    const int grid = pairs < m->sm_count ? pairs : m->sm_count;
    epic_wide_kernel<<<grid, kThreads, kSmemBytes, stream>>>(p);
    return cuda_ok(cudaGetLastError(), "epic_wide launch");
}

}  // namespace mmb
