/*
 * mmb_oracle.c — CPU restatement of the Multimodal-Bridges generation hot path.
 * TEST INFRASTRUCTURE ONLY (see mmb_oracle.h).  Parity: PINNED by tests/golden/*.npz.
 *
 * Build: oracle/Makefile (gcc -O2 -ffp-contract=off -mfma -fopenmp).  -ffp-contract=off matters:
 * every fused multiply-add below is an explicit fmaf() and every other product/sum is a separately
 * rounded IEEE fp32 operation, so the fp32 CUDA path (which spells the same sequence with
 * __fmaf_rn/__fmul_rn/__fadd_rn/__fdiv_rn) is bit-identical to this file, not merely close.
 *
 * Reference lines followed (mp/ = /root/reference/multimodal_particles/):
 *   time grid / loop        mp/models/generative/multimodal_bridge_matching.py:199-216
 *   sinusoidal embedding    mp/models/architectures/utils.py:183-198
 *   input embeddings        mp/models/architectures/utils.py:112-172
 *   EPiC trunk              mp/models/architectures/epic.py:136-241
 *   discrete head           mp/models/generative/multimodal_bridge_matching.py:90-113
 *   Euler step              mp/models/generative/bridges.py:38-45
 *   telegraph rate + jump   mp/models/generative/bridges.py:106-132,179-201
 *   absorbing step          mp/models/generative/bridges.py:218-231,251-286
 *
 * Order-of-operations choices (free within fp32 rounding of the reference; fixed here and in the
 * kernels): dot products start from the bias and accumulate with fmaf in ascending input index,
 * except that the per-jet part of a per-particle layer's input (time embedding, global vector) is
 * accumulated first; masked sums use the 128-lane tree of tree_sum().
 */
#include "mmb_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------ */
/* exp in IEEE fp32 operations only (Cody-Waite reduction + degree-5 polynomial on r^2 term), so
 * CPU and GPU agree bit-for-bit; |rel err| < 2 ulp on [-87, 88]. */
static inline float bits_to_float(uint32_t u) { float f; memcpy(&f, &u, 4); return f; }

float mmbo_expf(float x) {
    if (x != x) return x;
    if (x < -87.0f) return 0.0f;
    if (x > 88.0f) return INFINITY;
    float n = rintf(x * 1.44269504f);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r) + 1.0f;
    int e = (int)n;
    return y * bits_to_float((uint32_t)(e + 127) << 23);
}

/* same function with gradual underflow: results below 2^-126 come out as denormals (down to exp(-104) ~ 2^-150 -> 0),
 * as torch's exp does.  Only the batch-axis softmax of the trans-dimensional token rule needs this range. */
float mmbo_expf_dn(float x) {
    if (!(x < -87.0f)) return mmbo_expf(x);
    if (x < -104.0f) return 0.0f;
    float n = rintf(x * 1.44269504f);
    float r = fmaf(n, -0.693359375f, x);
    r = fmaf(n, 2.12194440e-4f, r);
    float p = 1.9875691500e-4f;
    p = fmaf(p, r, 1.3981999507e-3f);
    p = fmaf(p, r, 8.3334519073e-3f);
    p = fmaf(p, r, 4.1665795894e-2f);
    p = fmaf(p, r, 1.6666665459e-1f);
    p = fmaf(p, r, 5.0000001201e-1f);
    float r2 = r * r;
    float y = fmaf(p, r2, r) + 1.0f;
    int e = (int)n + 64;                                          /* y * 2^(e) is a normal number, exactly */
    return (y * bits_to_float((uint32_t)(e + 127) << 23)) * 5.42101086e-20f;   /* * 2^-64: one rounding into the denormals */
}

static inline float lrelu(float a) { return a > 0.0f ? a : a * 0.01f; }           /* F.leaky_relu */
static inline float selu(float a) {                                               /* nn.SELU */
    const float scale = 1.0507009873554804934193349852946f;
    const float alpha_scale = 1.0507009873554804934193349852946f * 1.6732632423543772848170429916717f;
    return a > 0.0f ? scale * a : alpha_scale * (mmbo_expf(a) - 1.0f);
}

/* ------------------------------------------------------------------------------------------ */
/* Philox4x32-10 (Salmon et al. 2011).  counter = (n>>2, step, jet_lo, jet_hi*4+stream), key = seed */
static inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        uint32_t n1 = (uint32_t)p1;
        uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

static inline float philox_uniform(uint64_t seed, uint64_t jet, int stream_id, int step, int n) {
    uint32_t c[4] = {(uint32_t)(n >> 2), (uint32_t)step, (uint32_t)jet,
                     (uint32_t)(jet >> 32) * 4u + (uint32_t)stream_id};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    return (float)(c[n & 3] >> 8) * (1.0f / 16777216.0f);
}

void mmbo_philox_uniforms(float* u, uint64_t seed, uint64_t jet_offset, int stream_id,
                          int n_steps, int B, int N) {
    for (int s = 0; s < n_steps; ++s)
        for (int b = 0; b < B; ++b)
            for (int n = 0; n < N; ++n)
                u[((size_t)s * B + b) * N + n] = philox_uniform(seed, jet_offset + (uint64_t)b, stream_id, s, n);
}

/* ------------------------------------------------------------------------------------------ */
void mmbo_step_table(int num_timesteps, float time_eps, int S, float gamma, float gamma_absorb, int T,
                     float* t, float* temb, float* bc, float* cc, float* sp, float* dt) {
    /* torch.linspace(0, 1-eps, n) fp32: start + i*step for the first half, end - (n-1-i)*step after */
    const int n = num_timesteps;
    const float end = (float)(1.0 - (double)time_eps);
    const float step = (end - 0.0f) / (float)(n - 1);
    float* grid = (float*)malloc(sizeof(float) * (size_t)n);
    for (int i = 0; i < n; ++i)
        grid[i] = (i < n / 2) ? 0.0f + step * (float)i : end - step * (float)(n - 1 - i);
    *dt = (grid[n - 1] - grid[0]) / (float)(n - 1);
    const int half = T / 2;
    for (int i = 1; i < n; ++i) {
        const float ti = grid[i];
        t[i - 1] = ti;
        for (int j = 0; j < half; ++j) {
            float f = expf((float)(-log(10000.0)) * (float)j / (float)half);
            float a = ti * f;
            temb[(size_t)(i - 1) * T + j] = cosf(a);
            temb[(size_t)(i - 1) * T + half + j] = sinf(a);
        }
        if (T % 2) temb[(size_t)(i - 1) * T + T - 1] = 0.0f;
        float w = expf((float)(-(double)S * (double)gamma) * (1.0f - ti));
        bc[i - 1] = (w * (float)S) / (1.0f - w);
        cc[i - 1] = w;
        if (sp) {
            float e1 = expf(-gamma_absorb * ti);
            float num = 1.0f - expf(gamma_absorb * (ti - 1.0f));
            float den = 1.0f - expf(-gamma_absorb);
            sp[i - 1] = e1 * num / den;
        }
    }
    free(grid);
}

/* ------------------------------------------------------------------------------------------ */
/* masked-sum order shared with the fp32 kernel: 128 lanes, lane = n mod 128 accumulates its
 * particles in ascending order from +0; xor-butterfly inside each 32-lane warp; warps in order. */
static float tree_sum(const float* vals, int stride, int N) {
    float lane[128];
    for (int i = 0; i < 128; ++i) lane[i] = 0.0f;
    for (int n = 0; n < N; ++n) lane[n & 127] = lane[n & 127] + vals[(size_t)n * stride];
    for (int off = 16; off >= 1; off >>= 1) {
        float nxt[128];
        for (int i = 0; i < 128; ++i) nxt[i] = lane[i] + lane[i ^ off];
        memcpy(lane, nxt, sizeof(lane));
    }
    return ((lane[0] + lane[32]) + lane[64]) + lane[96];
}

static inline float dot_from(float acc, const float* w, const float* in, int n) {
    for (int i = 0; i < n; ++i) acc = fmaf(w[i], in[i], acc);
    return acc;
}

typedef struct {
    float *xl, *skipl, *l1, *masked; /* [N][H] */
} Scratch;

static void epic_forward_jet(const MmbEpicDims* d, const MmbEpicLayout* Lo, const float* W,
                             const float* x, const uint8_t* k, const uint8_t* mask, const float* temb, int N,
                             float* v_out, float* logits_out, float* hidden_out, Scratch* sc) {
    const int Dc = d->dim_continuous, S = d->vocab_size, T = d->dim_time_emb, C = d->dim_cont_emb,
              D = d->dim_disc_emb, H = d->dim_hidden_local, G = d->dim_hidden_glob, L = d->num_blocks,
              Sh = d->disc_head_hidden;
    const int K0 = T + C + D;
    /* `temb` holds the reference's per-jet `context` vector [time embedding T | embedded context X] (utils.py:166-170);
     * only its time part is among the particle features (utils.py:139-145) */
    const int TX = T + d->dim_context;
    float *xl = sc->xl, *skipl = sc->skipl, *l1 = sc->l1, *masked = sc->masked;
    float pj[256], pool[768], g0[256], g1[256], xg[256], skipg[256], emb[256], sum[256];

    /* ---- InputEmbeddings + EPiC_Projection.local_0 (utils.py:133-172, epic.py:186) */
    for (int o = 0; o < H; ++o) pj[o] = dot_from(W[Lo->local0_b + o], W + Lo->local0_w + (size_t)o * K0, temb, T);
    float cnt = 0.0f;
    for (int n = 0; n < N; ++n) {
        const float m = (float)mask[n];
        cnt += m;
        if (mask[n]) {
            for (int c = 0; c < C; ++c)
                emb[c] = dot_from(W[Lo->emb_cont_b + c], W + Lo->emb_cont_w + (size_t)c * Dc, x + (size_t)n * Dc, Dc);
            const float* e = W + Lo->emb_disc + (size_t)k[n] * D;
            for (int o = 0; o < H; ++o) {
                const float* w = W + Lo->local0_w + (size_t)o * K0;
                float acc = dot_from(pj[o], w + T, emb, C);
                acc = dot_from(acc, w + T + C, e, D);
                xl[(size_t)n * H + o] = lrelu(acc);
            }
        } else {
            for (int o = 0; o < H; ++o) xl[(size_t)n * H + o] = lrelu(W[Lo->local0_b + o]);
        }
        for (int o = 0; o < H; ++o) masked[(size_t)n * H + o] = xl[(size_t)n * H + o] * m;
    }
    /* ---- meansum_pool + global_0..2 (epic.py:136-143,187-190) */
    for (int o = 0; o < H; ++o) {
        sum[o] = tree_sum(masked + o, H, N);
        pool[o] = sum[o] / cnt;
        pool[H + o] = sum[o];
    }
    for (int i = 0; i < TX; ++i) pool[2 * H + i] = temb[i];
    for (int o = 0; o < H; ++o)
        g0[o] = lrelu(dot_from(W[Lo->global0_b + o], W + Lo->global0_w + (size_t)o * (2 * H + TX), pool, 2 * H + TX));
    for (int o = 0; o < H; ++o)
        g1[o] = lrelu(dot_from(W[Lo->global1_b + o], W + Lo->global1_w + (size_t)o * H, g0, H));
    for (int o = 0; o < G; ++o)
        xg[o] = lrelu(dot_from(W[Lo->global2_b + o], W + Lo->global2_w + (size_t)o * H, g1, H));
    /* x_local * mask; skips (epic.py:148-149,191) */
    memcpy(xl, masked, sizeof(float) * (size_t)N * H);
    if (d->skip_connection) {
        memcpy(skipl, xl, sizeof(float) * (size_t)N * H);
        memcpy(skipg, xg, sizeof(float) * (size_t)G);
    }
    /* ---- EPiC layers (epic.py:217-241, 152-155) */
    for (int l = 0; l < L; ++l) {
        const float* Wl = W + Lo->layer0 + (size_t)l * Lo->layer_stride;
        for (int n = 0; n < N; ++n) {
            const float m = (float)mask[n];
            for (int o = 0; o < H; ++o) masked[(size_t)n * H + o] = xl[(size_t)n * H + o] * m;
        }
        for (int o = 0; o < H; ++o) {
            sum[o] = tree_sum(masked + o, H, N);
            pool[o] = sum[o] / cnt;
            pool[H + o] = sum[o];
        }
        for (int i = 0; i < G; ++i) pool[2 * H + i] = xg[i];
        for (int i = 0; i < TX; ++i) pool[2 * H + G + i] = temb[i];
        const int Kg = 2 * H + G + TX;
        for (int o = 0; o < H; ++o)
            g1[o] = lrelu(dot_from(Wl[Lo->l_g1_b + o], Wl + Lo->l_g1_w + (size_t)o * Kg, pool, Kg));
        for (int o = 0; o < G; ++o)
            g0[o] = lrelu(dot_from(Wl[Lo->l_g2_b + o], Wl + Lo->l_g2_w + (size_t)o * H, g1, H) + xg[o]);
        memcpy(xg, g0, sizeof(float) * (size_t)G);
        const int Kl = H + G + TX;
        for (int o = 0; o < H; ++o) {
            const float* w = Wl + Lo->l_l1_w + (size_t)o * Kl;
            float acc = dot_from(Wl[Lo->l_l1_b + o], w + H, xg, G);
            pj[o] = dot_from(acc, w + H + G, temb, TX);
        }
        for (int n = 0; n < N; ++n) {
            const float m = (float)mask[n];
            const float* xn = xl + (size_t)n * H;
            for (int o = 0; o < H; ++o)
                l1[(size_t)n * H + o] = lrelu(dot_from(pj[o], Wl + Lo->l_l1_w + (size_t)o * Kl, xn, H));
        }
        for (int n = 0; n < N; ++n) {
            const float m = (float)mask[n];
            float* xn = xl + (size_t)n * H;
            float nw[256];
            for (int o = 0; o < H; ++o) {
                float acc = dot_from(Wl[Lo->l_l2_b + o], Wl + Lo->l_l2_w + (size_t)o * H, l1 + (size_t)n * H, H);
                nw[o] = lrelu(acc + xn[o]) * m;
            }
            for (int o = 0; o < H; ++o) xn[o] = d->skip_connection ? nw[o] + skipl[(size_t)n * H + o] : nw[o];
        }
        if (d->skip_connection)
            for (int o = 0; o < G; ++o) xg[o] = xg[o] + skipg[o];
    }
    /* ---- output layer, heads (epic.py:158-162, mbm.py:105-113) */
    for (int n = 0; n < N; ++n) {
        const float m = (float)mask[n];
        const float* xn = xl + (size_t)n * H;
        float h[64], z1[256];
        for (int o = 0; o < Dc + S; ++o)
            h[o] = dot_from(W[Lo->out_b + o], W + Lo->out_w + (size_t)o * H, xn, H) * m;
        for (int c = 0; c < Dc; ++c) v_out[(size_t)n * Dc + c] = h[c];
        if (Sh) {
            for (int o = 0; o < Sh; ++o)
                z1[o] = selu(dot_from(W[Lo->head0_b + o], W + Lo->head0_w + (size_t)o * S, h + Dc, S));
            for (int o = 0; o < S; ++o)
                logits_out[(size_t)n * S + o] = dot_from(W[Lo->head2_b + o], W + Lo->head2_w + (size_t)o * Sh, z1, Sh);
        } else {
            for (int o = 0; o < S; ++o) logits_out[(size_t)n * S + o] = h[Dc + o];
        }
        if (hidden_out)
            for (int o = 0; o < H; ++o) hidden_out[(size_t)n * H + o] = xn[o];
    }
}

static int dims_ok(const MmbEpicDims* d) {
    return d->dim_hidden_local <= 256 && d->dim_hidden_glob <= 256 && d->dim_time_emb + d->dim_context <= 256 && d->dim_context >= 0 &&
           d->dim_cont_emb <= 256 && d->disc_head_hidden <= 256 && d->dim_continuous + d->vocab_size <= 64 &&
           d->vocab_size <= 32;
}

static Scratch scratch_new(int N, int H) {
    Scratch s;
    size_t n = (size_t)N * H;
    s.xl = (float*)malloc(sizeof(float) * n);
    s.skipl = (float*)malloc(sizeof(float) * n);
    s.l1 = (float*)malloc(sizeof(float) * n);
    s.masked = (float*)malloc(sizeof(float) * n);
    return s;
}
static void scratch_free(Scratch* s) { free(s->xl); free(s->skipl); free(s->l1); free(s->masked); }

void mmbo_epic_forward(const MmbEpicDims* dims, const float* packed,
                       const float* x, const uint8_t* k, const uint8_t* mask,
                       const float* temb, int temb_stride, int B, int N,
                       float* v_out, float* logits_out, float* hidden_out) {
    if (!dims_ok(dims)) return;
    const MmbEpicLayout Lo = mmb_epic_layout(dims);
    const int Dc = dims->dim_continuous, S = dims->vocab_size, H = dims->dim_hidden_local;
#pragma omp parallel
    {
        Scratch sc = scratch_new(N, H);
#pragma omp for schedule(static)
        for (int b = 0; b < B; ++b)
            epic_forward_jet(dims, &Lo, packed, x + (size_t)b * N * Dc, k + (size_t)b * N, mask + (size_t)b * N,
                             temb + (size_t)b * temb_stride, N, v_out + (size_t)b * N * Dc,
                             logits_out + (size_t)b * N * S, hidden_out ? hidden_out + (size_t)b * N * H : NULL, &sc);
        scratch_free(&sc);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* The fused hybrid update for one particle (bridges.py:260-286, 38-45, 106-132, 179-201).
 * Jump: S independent Poisson(lam_s) draws gated by "at most one jump in total" are, exactly, a
 * categorical with P(select s) = lam_s * exp(-Lam), Lam = sum_s lam_s, "none selected" = stay and
 * "s = k selected" = stay (the self slot is retained so the thresholds do not depend on k beyond
 * q_k).  One uniform decides: first s with u < c_s, c_s = sequential fp32 prefix sum. */
static inline void update_particle(float* x, uint8_t* k, uint8_t* mask, const float* v, const float* logits,
                                   const float* absorb_logit, float u_jump, const float* u_absorb,
                                   float dt, float bc, float cc, float sp, int Dc, int S, int flags) {
    uint8_t m = *mask;
    if (flags & MMB_FLAG_ABSORBING) {
        float sg = 1.0f / (1.0f + mmbo_expf(-absorb_logit[0]));
        float p = dt * (sp * sg);
        p = p > 1.0f ? 1.0f : p; /* torch.clamp(max=1): NaN stays NaN -> (u < NaN) false */
        uint8_t born = (*u_absorb < p) ? 1 : 0;
        m = (m == 1) ? 1 : born;
        *mask = m;
    }
    const float mf = (float)m;
    if (!(flags & MMB_FLAG_NO_EULER))
        for (int c = 0; c < Dc; ++c) x[c] = (x[c] + dt * v[c]) * mf;
    if (flags & MMB_FLAG_NO_JUMP) return;

    float mx = logits[0];
    for (int s = 1; s < S; ++s) mx = logits[s] > mx ? logits[s] : mx;
    float e[32], z = 0.0f;
    for (int s = 0; s < S; ++s) { e[s] = mmbo_expf(logits[s] - mx); z = z + e[s]; }
    const int kk = *k;
    const float zinv = 1.0f / z; /* q_s = e_s * (1/z): one IEEE division per particle */
    const float ck = cc * (e[kk] * zinv);
    float lam[32], Lam = 0.0f;
    for (int s = 0; s < S; ++s) {
        float q = e[s] * zinv;
        float rate = (1.0f + bc * q) + ck;
        lam[s] = rate * dt;
        Lam = Lam + lam[s];
    }
    const float E = mmbo_expf(-Lam);
    int nk = kk;
    float c = 0.0f;
    for (int s = 0; s < S; ++s) {
        c = c + lam[s] * E;
        if (u_jump < c) { nk = s; break; }
    }
    *k = (uint8_t)(nk * (int)m);
}

void mmbo_bridge_update(float* x, uint8_t* k, uint8_t* mask,
                        const float* v, const float* logits, const float* absorb_logit,
                        const float* u_jump, const float* u_absorb,
                        float dt, float bc, float cc, float sp,
                        int B, int N, int Dc, int S, int flags) {
    const size_t P = (size_t)B * N;
#pragma omp parallel for schedule(static)
    for (size_t p = 0; p < P; ++p)
        update_particle(x + p * Dc, k + p, mask + p, v ? v + p * Dc : NULL, logits ? logits + p * S : NULL,
                        absorb_logit ? absorb_logit + p : NULL, u_jump ? u_jump[p] : 2.0f, u_absorb ? u_absorb + p : NULL,
                        dt, bc, cc, sp, Dc, S, flags);
}

/* ------------------------------------------------------------------------------------------ */
/* absorbing-rate transformer head (absorbing_flows.py:94-131).  h is kept as [N][C] (particle-major);
 * the reference's [B,C,N] layout is a transpose of it. */
size_t mmbo_absorb_head_floats(int H, int C, int n_blocks) {
    size_t lin = (size_t)C * C + C, nrm = 2 * (size_t)C;
    return (size_t)C * (H + 2) + C + (size_t)n_blocks * (3 * nrm + 6 * lin) + lin + C + 1;
}

static inline float swishf(float a) { return a * (1.0f / (1.0f + mmbo_expf(-a))); }

/* GroupNorm(32 groups, eps 1e-6, affine) over [N][C] (gsdm.py:34-35) */
static void group_norm(const float* in, float* out, const float* g, const float* b, int N, int C) {
    const int gs = C / 32;
    for (int grp = 0; grp < 32; ++grp) {
        double s = 0.0, q = 0.0;
        for (int n = 0; n < N; ++n)
            for (int c = grp * gs; c < (grp + 1) * gs; ++c) { double v = in[(size_t)n * C + c]; s += v; q += v * v; }
        const double cnt = (double)N * gs, mean = s / cnt;
        double var = q / cnt - mean * mean;
        if (var < 0) var = 0;
        const float rstd = (float)(1.0 / sqrt(var + 1e-6)), mf = (float)mean;
        for (int n = 0; n < N; ++n)
            for (int c = grp * gs; c < (grp + 1) * gs; ++c)
                out[(size_t)n * C + c] = (in[(size_t)n * C + c] - mf) * rstd * g[c] + b[c];
    }
}

static void linear_rows(const float* in, float* out, const float* W, const float* b, int N, int Cin, int Cout) {
    for (int n = 0; n < N; ++n)
        for (int o = 0; o < Cout; ++o)
            out[(size_t)n * Cout + o] = dot_from(b[o], W + (size_t)o * Cin, in + (size_t)n * Cin, Cin);
}

typedef struct { float *t1, *t2, *q, *kk, *vv, *sc; } TfScratch;
static TfScratch tf_scratch_new(int N, int C) {
    TfScratch s;
    s.t1 = (float*)malloc(sizeof(float) * (size_t)N * C);
    s.t2 = (float*)malloc(sizeof(float) * (size_t)N * C);
    s.q = (float*)malloc(sizeof(float) * (size_t)N * C);
    s.kk = (float*)malloc(sizeof(float) * (size_t)N * C);
    s.vv = (float*)malloc(sizeof(float) * (size_t)N * C);
    s.sc = (float*)malloc(sizeof(float) * (size_t)N);
    return s;
}
static void tf_scratch_free(TfScratch* s) { free(s->t1); free(s->t2); free(s->q); free(s->kk); free(s->vv); free(s->sc); }

/* n_blocks x (ResnetBlock, AttnBlock) on x [N][C] in place (gsdm.py:54-66,142-168); tb [n_blocks][C] is the
 * temb_proj(swish(temb)) term of each ResnetBlock.  Returns the blob pointer after the last block. */
static const float* tf_blocks(const float* p, float* x, const float* tb_all, int N, int C, int n_heads, int n_blocks, TfScratch* s) {
    const size_t lin = (size_t)C * C + C;
    float *t1 = s->t1, *t2 = s->t2, *q = s->q, *kk = s->kk, *vv = s->vv, *sc = s->sc;
    const int dh = C / n_heads;
    const float scale = 1.0f / sqrtf((float)dh);
    for (int blk = 0; blk < n_blocks; ++blk) {
        const float *n1g = p, *n1b = p + C, *c1 = p + 2 * C, *n2g = c1 + lin, *n2b = n2g + C, *c2 = n2b + C,
                    *ng = c2 + lin, *nb = ng + C, *wq = nb + C, *wk = wq + lin, *wv = wk + lin, *wo = wv + lin;
        p = wo + lin;
        const float* tb = tb_all + (size_t)blk * C;
        /* ResnetBlock (gsdm.py:54-66) */
        group_norm(x, t1, n1g, n1b, N, C);
        for (size_t i = 0; i < (size_t)N * C; ++i) t1[i] = swishf(t1[i]);
        linear_rows(t1, t2, c1, c1 + (size_t)C * C, N, C, C);
        for (int n = 0; n < N; ++n)
            for (int c = 0; c < C; ++c) t2[(size_t)n * C + c] += tb[c];
        group_norm(t2, t1, n2g, n2b, N, C);
        for (size_t i = 0; i < (size_t)N * C; ++i) t1[i] = swishf(t1[i]);
        linear_rows(t1, t2, c2, c2 + (size_t)C * C, N, C, C);
        for (size_t i = 0; i < (size_t)N * C; ++i) x[i] += t2[i];
        /* AttnBlock (gsdm.py:142-168): all N slots attend to all N slots, no padding mask */
        group_norm(x, t1, ng, nb, N, C);
        linear_rows(t1, q, wq, wq + (size_t)C * C, N, C, C);
        linear_rows(t1, kk, wk, wk + (size_t)C * C, N, C, C);
        linear_rows(t1, vv, wv, wv + (size_t)C * C, N, C, C);
        for (int h = 0; h < n_heads; ++h)
            for (int qi = 0; qi < N; ++qi) {
                float mx = -INFINITY;
                for (int ki = 0; ki < N; ++ki) {
                    float a = dot_from(0.0f, kk + (size_t)ki * C + h * dh, q + (size_t)qi * C + h * dh, dh) * scale;
                    sc[ki] = a;
                    mx = a > mx ? a : mx;
                }
                float z = 0.0f;
                for (int ki = 0; ki < N; ++ki) { sc[ki] = mmbo_expf(sc[ki] - mx); z += sc[ki]; }
                const float zi = 1.0f / z;
                for (int d = 0; d < dh; ++d) {
                    float acc = 0.0f;
                    for (int ki = 0; ki < N; ++ki) acc = fmaf(vv[(size_t)ki * C + h * dh + d], sc[ki] * zi, acc);
                    t2[(size_t)qi * C + h * dh + d] = acc;
                }
            }
        linear_rows(t2, t1, wo, wo + (size_t)C * C, N, C, C);
        for (size_t i = 0; i < (size_t)N * C; ++i) x[i] += t1[i];
    }
    return p;
}

void mmbo_absorb_head(const float* W, int H, int C, int n_heads, int n_blocks,
                      const float* hidden, const uint8_t* mask, const float* tbias, int tbias_stride,
                      int B, int N, float* logit_out) {
    const size_t lin = (size_t)C * C + C;
#pragma omp parallel
    {
        float* x = (float*)malloc(sizeof(float) * (size_t)N * C);
        TfScratch s = tf_scratch_new(N, C);
        float* in0 = (float*)malloc(sizeof(float) * (size_t)(H + 2));
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            const float* p = W;
            /* transformer_1_proj_in on cat[x_local_last, one_hot(mask)] (absorbing_flows.py:113-118) */
            for (int n = 0; n < N; ++n) {
                for (int i = 0; i < H; ++i) in0[i] = hidden[((size_t)b * N + n) * H + i];
                in0[H] = mask[(size_t)b * N + n] ? 0.0f : 1.0f;
                in0[H + 1] = mask[(size_t)b * N + n] ? 1.0f : 0.0f;
                for (int o = 0; o < C; ++o)
                    x[(size_t)n * C + o] = dot_from(p[(size_t)C * (H + 2) + o], p + (size_t)o * (H + 2), in0, H + 2);
            }
            p += (size_t)C * (H + 2) + C;
            p = tf_blocks(p, x, tbias + (size_t)b * tbias_stride, N, C, n_heads, n_blocks, &s);
            /* pre_rate_proj, post_rate_proj (absorbing_flows.py:127-131) */
            linear_rows(x, s.t1, p, p + (size_t)C * C, N, C, C);
            p += lin;
            for (int n = 0; n < N; ++n) logit_out[(size_t)b * N + n] = dot_from(p[C], p, s.t1 + (size_t)n * C, C);
        }
        free(x); free(in0); tf_scratch_free(&s);
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Trans-dimensional jump diffusion: TransdimensionalEPiC.forward + JumpSampler.sample
 * (mp/models/generative/transdimensional/transdimensional_model.py:245-426, sampler.py:157-324,
 *  structure.py:226-250, mp/models/generative/diffusion/noising.py:15-39,123-216,
 *  mp/data/particle_clouds/jets_dataloader.py:380-478). */
/* outputs of post_rate_proj: max_num_particles x0-dimension logits, or one rate logit with rate_use_x0_pred = False (model :185-188) */
static int trans_rate_dim(const MmbTransDims* d) { return d->rate_direct ? 1 : d->max_particles; }

size_t mmbo_trans_floats(const MmbTransDims* d) {
    const size_t C = d->transformer_dim, lin = C * C + C, nrm = 2 * C, H = d->hidden, S = d->vocab_size, R = (size_t)trans_rate_dim(d);
    const size_t block = 3 * nrm + 6 * lin;
    return lin + 2 * (size_t)d->n_blocks * lin
         + (C * (H + S) + C) + d->n_blocks * block + lin + (R * C + R) + (C + 1)
         + (C * (H + S + 3) + C) + d->n_blocks * block + (C + 1) + lin + ((2 * S + 1) * C + (2 * S + 1));
}

/* tokens = argmax_s softmax_BATCH(onehot)[b,n,s]: F.softmax without dim on a 3-D tensor normalises over
 * dim 0 (structure.py:231-232).  Column sums in the order the kernels use: chunks of 128 jets; inside a chunk 8
 * partial sums over b mod 8 (ascending b) combined by a balanced tree; chunk sums added in ascending order. */
void mmbo_trans_tokens(const float* onehot, int B, int N, int S, uint8_t* k) {
    const int NS = N * S;
    float* M = (float*)malloc(sizeof(float) * (size_t)NS);
    float* Z = (float*)malloc(sizeof(float) * (size_t)NS);
    for (int c = 0; c < NS; ++c) {
        float mx = -INFINITY;
        for (int b = 0; b < B; ++b) { float a = onehot[(size_t)b * NS + c]; mx = a > mx ? a : mx; }
        float z = 0.0f;
        for (int b0 = 0; b0 < B; b0 += 128) {
            float part[8] = {0, 0, 0, 0, 0, 0, 0, 0};
            for (int b = b0; b < B && b < b0 + 128; ++b) part[b & 7] = part[b & 7] + mmbo_expf_dn(onehot[(size_t)b * NS + c] - mx);
            z = z + (((part[0] + part[1]) + (part[2] + part[3])) + ((part[4] + part[5]) + (part[6] + part[7])));
        }
        M[c] = mx;
        Z[c] = z;
    }
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n) {
            int best = 0; float bp = -1.0f;
            for (int s2 = 0; s2 < S; ++s2) {
                const int c = n * S + s2;
                const float pr = mmbo_expf_dn(onehot[(size_t)b * NS + c] - M[c]) / Z[c];
                if (pr > bp) { bp = pr; best = s2; }
            }
            k[(size_t)b * N + n] = (uint8_t)best;
        }
    free(M); free(Z);
}

/* per-jet time terms: EPiC sinusoidal embedding of ts (utils.py:183-198) and, for both stacks,
 * temb_proj_blk(swish(temb_net(get_timestep_embedding(1000 ts, C)))) (gsdm.py:8-26,58; model :288-290) */
void mmbo_trans_time_terms(const MmbTransDims* d, const float* W, const float* ts, int B, int T,
                           float* temb_epic /*[B][T]*/, float* tb1 /*[B][n_blocks][C]*/, float* tb2) {
    const int C = d->transformer_dim, half = C / 2, nb = d->n_blocks;
    const size_t lin = (size_t)C * C + C;
    float* e = (float*)malloc(sizeof(float) * C);
    float* t = (float*)malloc(sizeof(float) * C);
    for (int b = 0; b < B; ++b) {
        const int h2 = T / 2;
        for (int j = 0; j < h2; ++j) {
            float f = expf((float)(-log(10000.0)) * (float)j / (float)h2);
            float a = ts[b] * f;
            temb_epic[(size_t)b * T + j] = cosf(a);
            temb_epic[(size_t)b * T + h2 + j] = sinf(a);
        }
        if (T % 2) temb_epic[(size_t)b * T + T - 1] = 0.0f;
        const float tt = ts[b] * 1000.0f;
        const float fe = (float)(log(10000.0) / (double)(half - 1));
        for (int j = 0; j < half; ++j) {
            float f = expf((float)j * -fe);
            e[j] = sinf(tt * f);
            e[half + j] = cosf(tt * f);
        }
        if (C % 2) e[C - 1] = 0.0f;
        for (int o = 0; o < C; ++o) t[o] = swishf(dot_from(W[(size_t)C * C + o], W + (size_t)o * C, e, C));
        for (int st = 0; st < 2; ++st)
            for (int blk = 0; blk < nb; ++blk) {
                const float* P = W + lin + ((size_t)st * nb + blk) * lin;
                float* out = (st ? tb2 : tb1) + ((size_t)b * nb + blk) * C;
                for (int o = 0; o < C; ++o) out[o] = dot_from(P[(size_t)C * C + o], P + (size_t)o * C, t, C);
            }
    }
    free(e); free(t);
}

static float fr_rate(const MmbForwardRate* fr, float t) {
    if (fr->kind == 1) return fr->scalar;
    return fr->scalar * (t > fr->rate_cut_t ? 1.0f : 0.0f) + fr->offset;
}
static float fr_integral(const MmbForwardRate* fr, float t) {
    if (fr->kind == 1) return fr->scalar * t;
    return (t - fr->rate_cut_t) * fr->scalar * (t > fr->rate_cut_t ? 1.0f : 0.0f) + fr->offset * t;
}
static float poisson_logp(float k, float lam) { return (k == 0.0f ? 0.0f : k * logf(lam)) - lam - lgammaf(k + 1.0f); }

/* get_rate_using_x0_pred (noising.py:166-216) for one jet */
float mmbo_trans_rate(const float* logits, int R, int xt_dim, const MmbForwardRate* fr, float t) {
    const float I = fr_integral(fr, t);
    float mx = -INFINITY, z = 0.0f, acc = 0.0f;
    for (int i = xt_dim - 1; i < R; ++i) mx = logits[i] > mx ? logits[i] : mx;
    for (int i = xt_dim - 1; i < R; ++i) z += mmbo_expf(logits[i] - mx);
    for (int i = xt_dim - 1; i < R; ++i) {
        const float prob = mmbo_expf(logits[i] - mx) / z;
        float ratio;
        if (xt_dim > 1) {
            ratio = (1.0f / I) * (float)((i + 1) - xt_dim);
            if (ratio < 0.0f) ratio = 0.0f;
        } else {
            float m2 = -INFINITY;
            const int trunc = 2 * R;
            for (int j = 0; j < trunc; ++j) { float lp = poisson_logp((float)(i + j), I); m2 = lp > m2 ? lp : m2; }
            float se = 0.0f;
            for (int j = 0; j < trunc; ++j) se += expf(poisson_logp((float)(i + j), I) - m2);
            const float dim1 = m2 + logf(se);
            const float dim2 = i == 0 ? -1000.0f : poisson_logp((float)(i - 1 > 0 ? i - 1 : 0), I);
            ratio = expf(dim2 - dim1);
        }
        acc += ratio * prob;
    }
    return fr_rate(fr, t) * acc;
}

static inline float softplusf(float a) { return a > 20.0f ? a : log1pf(expf(a)); }

void mmbo_trans_forward(const MmbEpicDims* ed, const float* epacked, const MmbTransDims* d, const float* W,
                        const float* x, const float* onehot, const int32_t* dims, const float* ts,
                        const int32_t* nearest_in, const float* u_nearest, const MmbForwardRate* fr, int B, int N,
                        float* d_xt, float* rate, float* auto_mean, float* auto_std, float* x0_dim_logits,
                        float* near_atom_logits, int32_t* nearest_out, float* new_mean, float* new_std) {
    const int C = d->transformer_dim, H = d->hidden, S = d->vocab_size, R = trans_rate_dim(d), nb = d->n_blocks, T = ed->dim_time_emb;
    const int F = 3 + S;
    const size_t lin = (size_t)C * C + C, block = 6 * (size_t)C + 6 * lin;
    uint8_t* k = (uint8_t*)malloc((size_t)B * N);
    uint8_t* mask = (uint8_t*)malloc((size_t)B * N);
    float* temb = (float*)malloc(sizeof(float) * (size_t)B * T);
    float* tb1 = (float*)malloc(sizeof(float) * (size_t)B * nb * C);
    float* tb2 = (float*)malloc(sizeof(float) * (size_t)B * nb * C);
    float* v = (float*)malloc(sizeof(float) * (size_t)B * N * 3);
    float* lg = (float*)malloc(sizeof(float) * (size_t)B * N * S);
    float* hid = (float*)malloc(sizeof(float) * (size_t)B * N * H);
    mmbo_trans_tokens(onehot, B, N, S, k);
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n) mask[(size_t)b * N + n] = n < dims[b];
    mmbo_trans_time_terms(d, W, ts, B, T, temb, tb1, tb2);
    mmbo_epic_forward(ed, epacked, x, k, mask, temb, T, B, N, v, lg, hid);
    /* D_xt: all continuous slots, then all one-hot slots (model :277-280) */
    for (int b = 0; b < B; ++b) {
        memcpy(d_xt + (size_t)b * N * F, v + (size_t)b * N * 3, sizeof(float) * (size_t)N * 3);
        memcpy(d_xt + (size_t)b * N * F + (size_t)N * 3, lg + (size_t)b * N * S, sizeof(float) * (size_t)N * S);
    }
    const float* s1 = W + lin + 2 * (size_t)nb * lin;
    const size_t in1 = H + S, in2 = H + S + 3;
    const float* s1_blocks = s1 + (size_t)C * in1 + C;
    const float* pre_rate = s1_blocks + (size_t)nb * block;
    const float* post_rate = pre_rate + lin;
    const float* near_w = post_rate + (size_t)R * C + R;
    const float* s2 = near_w + C + 1;
    const float* s2_blocks = s2 + (size_t)C * in2 + C;
    const float* vecw = s2_blocks + (size_t)nb * block;
    const float* pre_auto = vecw + C + 1;
    const float* post_auto = pre_auto + lin;
    const int PA = 2 * S + 1;
#pragma omp parallel
    {
        float* h = (float*)malloc(sizeof(float) * (size_t)N * C);
        float* in = (float*)malloc(sizeof(float) * (in2 + 1));
        float* emb = (float*)malloc(sizeof(float) * (size_t)C);
        float* pa = (float*)malloc(sizeof(float) * (size_t)PA);
        float* vw = (float*)malloc(sizeof(float) * (size_t)N);
        TfScratch s = tf_scratch_new(N, C);
#pragma omp for schedule(dynamic, 1)
        for (int b = 0; b < B; ++b) {
            const float* xb = x + (size_t)b * N * 3;
            const float* ob = onehot + (size_t)b * N * S;
            const float* hb = hid + (size_t)b * N * H;
            const uint8_t* mb = mask + (size_t)b * N;
            /* ---- rate / nearest-particle stack (model :295-334): inputs NOT masked */
            for (int n = 0; n < N; ++n) {
                for (int i = 0; i < H; ++i) in[i] = hb[(size_t)n * H + i];
                for (int i = 0; i < S; ++i) in[H + i] = ob[(size_t)n * S + i];
                for (int o = 0; o < C; ++o) h[(size_t)n * C + o] = dot_from(s1[(size_t)C * in1 + o], s1 + (size_t)o * in1, in, (int)in1);
            }
            tf_blocks(s1_blocks, h, tb1 + (size_t)b * nb * C, N, C, d->n_heads, nb, &s);
            linear_rows(h, s.t1, pre_rate, pre_rate + (size_t)C * C, N, C, C);
            for (int c = 0; c < C; ++c) {
                float acc = 0.0f;
                for (int n = 0; n < N; ++n) acc += s.t1[(size_t)n * C + c];
                emb[c] = acc / (float)N;
            }
            if (d->rate_direct) {   /* model :326-332: x0_dim_logits = 0, rate = softplus(rate logit) * forward_rate(t) */
                float* xl = x0_dim_logits + (size_t)b * d->max_particles;
                for (int r = 0; r < d->max_particles; ++r) xl[r] = 0.0f;
                const float a = dot_from(post_rate[(size_t)C], post_rate, emb, C);
                rate[b] = (a > 20.0f ? a : log1pf(expf(a))) * fr_rate(fr, ts[b]);   /* torch softplus: beta 1, threshold 20 */
            } else {
                float* xl = x0_dim_logits + (size_t)b * R;
                for (int r = 0; r < R; ++r) xl[r] = dot_from(post_rate[(size_t)R * C + r], post_rate + (size_t)r * C, emb, C);
                rate[b] = mmbo_trans_rate(xl, R, dims[b], fr, ts[b]);
            }
            float* nl = near_atom_logits + (size_t)b * N;
            for (int n = 0; n < N; ++n) nl[n] = dot_from(near_w[C], near_w, h + (size_t)n * C, C);
            int near;
            if (nearest_in) near = nearest_in[b];
            else {  /* multinomial(softmax(near_atom_logits)) over ALL N slots, by inverse CDF on one uniform */
                float mx = -INFINITY, z = 0.0f, c = 0.0f;
                for (int n = 0; n < N; ++n) mx = nl[n] > mx ? nl[n] : mx;
                for (int n = 0; n < N; ++n) z += mmbo_expf(nl[n] - mx);
                near = N - 1;
                for (int n = 0; n < N; ++n) {
                    c += mmbo_expf(nl[n] - mx) / z;
                    if (u_nearest[b] < c) { near = n; break; }
                }
            }
            if (nearest_out) nearest_out[b] = near;
            /* ---- vector stack (model :341-409): inputs masked */
            const float xa0 = xb[near * 3], xa1 = xb[near * 3 + 1], xa2 = xb[near * 3 + 2];
            for (int n = 0; n < N; ++n) {
                const float m = mb[n] ? 1.0f : 0.0f;
                const float d0 = xa0 - xb[n * 3], d1 = xa1 - xb[n * 3 + 1], d2 = xa2 - xb[n * 3 + 2];
                for (int i = 0; i < H; ++i) in[i] = hb[(size_t)n * H + i] * m;
                for (int i = 0; i < S; ++i) in[H + i] = ob[(size_t)n * S + i] * m;
                in[H + S] = sqrtf((d0 * d0 + d1 * d1) + d2 * d2) * m;
                in[H + S + 1] = (n == near ? 1.0f : 0.0f) * m;
                in[H + S + 2] = (n == near ? 0.0f : 1.0f) * m;
                for (int o = 0; o < C; ++o) h[(size_t)n * C + o] = dot_from(s2[(size_t)C * in2 + o], s2 + (size_t)o * in2, in, (int)in2);
            }
            tf_blocks(s2_blocks, h, tb2 + (size_t)b * nb * C, N, C, d->n_heads, nb, &s);
            float pm0 = 0.0f, pm1 = 0.0f, pm2 = 0.0f;
            for (int n = 0; n < N; ++n) {
                const float w = dot_from(vecw[C], vecw, h + (size_t)n * C, C);
                const float m = mb[n] ? 1.0f : 0.0f;
                float d0 = (xa0 - xb[n * 3]) * m, d1 = (xa1 - xb[n * 3 + 1]) * m, d2 = (xa2 - xb[n * 3 + 2]) * m;
                const float nr = sqrtf((d0 * d0 + d1 * d1) + d2 * d2) + 1e-3f;
                pm0 += w * (d0 / nr); pm1 += w * (d1 / nr); pm2 += w * (d2 / nr);
            }
            pm0 = xa0 + pm0; pm1 = xa1 + pm1; pm2 = xa2 + pm2;
            linear_rows(h, s.t1, pre_auto, pre_auto + (size_t)C * C, N, C, C);
            for (int c = 0; c < C; ++c) {
                float acc = 0.0f;
                for (int n = 0; n < N; ++n) acc += s.t1[(size_t)n * C + c];
                emb[c] = acc / (float)N;
            }
            for (int r = 0; r < PA; ++r) pa[r] = dot_from(post_auto[(size_t)PA * C + r], post_auto + (size_t)r * C, emb, C);
            float mean11[64], std11[64];
            mean11[0] = pm0; mean11[1] = pm1; mean11[2] = pm2;
            std11[0] = std11[1] = std11[2] = pa[0];
            for (int s2i = 0; s2i < S; ++s2i) { mean11[3 + s2i] = pa[1 + s2i]; std11[3 + s2i] = pa[1 + S + s2i]; }
            if (new_mean) for (int i = 0; i < F; ++i) { new_mean[(size_t)b * F + i] = mean11[i]; new_std[(size_t)b * F + i] = std11[i]; }
            if (auto_mean) {
                float* am = auto_mean + (size_t)b * N * F;
                float* as = auto_std + (size_t)b * N * F;
                memset(am, 0, sizeof(float) * (size_t)N * F);
                memset(as, 0, sizeof(float) * (size_t)N * F);
                const int slot = dims[b];   /* get_next_dim_added_mask (structure.py:175-184): the slot a birth fills */
                if (slot < N) {
                    for (int c = 0; c < 3; ++c) { am[slot * 3 + c] = mean11[c]; as[slot * 3 + c] = std11[c]; }
                    for (int s2i = 0; s2i < S; ++s2i) {
                        am[(size_t)N * 3 + slot * S + s2i] = mean11[3 + s2i];
                        as[(size_t)N * 3 + slot * S + s2i] = std11[3 + s2i];
                    }
                }
            }
        }
        free(h); free(in); free(emb); free(pa); free(vw); tf_scratch_free(&s);
    }
    free(k); free(mask); free(temb); free(tb1); free(tb2); free(v); free(lg); free(hid);
}

static inline float nan_to_num(float a) { return a != a ? 0.0f : (a > 3.4028234664e38f ? 3.4028234664e38f : (a < -3.4028234664e38f ? -3.4028234664e38f : a)); }

/* JetsGraphicalStructure.adjust_st_batch (jets_dataloader.py:433-478): nan_to_num, then remove the mean of the
 * continuous features of the live particles */
static void adjust_jet(float* x, float* oh, int dim, int N, int S) {
    for (int i = 0; i < N * 3; ++i) x[i] = nan_to_num(x[i]);
    for (int i = 0; i < N * S; ++i) oh[i] = nan_to_num(oh[i]);
    const int cnt = dim == 0 ? N : dim;
    for (int c = 0; c < 3; ++c) {
        float acc = 0.0f;
        for (int n = 0; n < N; ++n) acc += x[n * 3 + c];
        const float mean = acc / (float)cnt;
        for (int n = 0; n < cnt; ++n) x[n * 3 + c] = x[n * 3 + c] - mean;
    }
}

/* one JumpSampler update (sampler.py:221-255) with injected noise */
void mmbo_trans_sampler_update(float* x, float* onehot, int32_t* dims, const float* v, const float* logits, const float* rate,
                               const float* new_mean, const float* new_std,
                               float c_decay, float c_score, float c_noise, float inv_std, float jump_dt,
                               const float* z_diff, const float* u_jump, const float* z_new, int B, int N, int S) {
    const int F = 3 + S;
    float* zc = (float*)malloc(sizeof(float) * (size_t)N * 3);
    for (int b = 0; b < B; ++b) {
        float* xb = x + (size_t)b * N * 3;
        float* ob = onehot + (size_t)b * N * S;
        const float* zb = z_diff + (size_t)b * N * F;
        const int dim = dims[b];
        /* noise: delete_dims + adjust_st_batch on the noise batch (sampler.py:224-229) */
        for (int i = 0; i < N * 3; ++i) zc[i] = (i / 3) < dim ? zb[i] : 0.0f;
        for (int c = 0; c < 3; ++c) {
            float acc = 0.0f;
            for (int n = 0; n < N; ++n) acc += zc[n * 3 + c];
            const float mean = acc / (float)dim;
            for (int n = 0; n < dim; ++n) zc[n * 3 + c] = zc[n * 3 + c] - mean;
        }
        for (int n = 0; n < N; ++n) {
            const float m = n < dim ? 1.0f : 0.0f;
            for (int c = 0; c < 3; ++c) {
                const float score = -(inv_std * v[((size_t)b * N + n) * 3 + c]);
                float a = c_decay * xb[n * 3 + c] + m * (c_score * score);
                if (c_noise != 0.0f) a = a + m * (c_noise * zc[n * 3 + c]);
                xb[n * 3 + c] = a;
            }
            for (int s = 0; s < S; ++s) {
                const float score = -(inv_std * logits[((size_t)b * N + n) * S + s]);
                float a = c_decay * ob[n * S + s] + m * (c_score * score);
                if (c_noise != 0.0f) a = a + m * (c_noise * zb[(size_t)N * 3 + n * S + s]);
                ob[n * S + s] = a;
            }
        }
        adjust_jet(xb, ob, dim, N, S);
        /* birth (sampler.py:238-255) */
        if (u_jump[b] < rate[b] * jump_dt && dim < N) {
            for (int c = 0; c < 3; ++c)
                xb[dim * 3 + c] = new_mean[(size_t)b * F + c] + z_new[(size_t)b * F + c] * softplusf(new_std[(size_t)b * F + c]);
            for (int s = 0; s < S; ++s)
                ob[dim * S + s] = new_mean[(size_t)b * F + 3 + s] + z_new[(size_t)b * F + 3 + s] * softplusf(new_std[(size_t)b * F + 3 + s]);
            dims[b] = dim + 1;
        }
        adjust_jet(xb, ob, dims[b], N, S);
    }
    free(zc);
}

/* one Langevin corrector update (sampler.py:258-282) and, with jump_corrector, its birth/death jumps (:285-312).
 * v / logits / rate / new_mean / new_std come from the evaluation at t - dt; alpha = 1 - dt beta(t - dt).
 * mask_dims: the reference reuses the `mask` of the predictor step (sampler.py:219) in every corrector of that step, so
 * particles born since then take no Langevin increment; the noise is centred, and the state re-centred, over the
 * current dims. */
void mmbo_trans_corrector_update(float* x, float* onehot, int32_t* dims, const int32_t* mask_dims, const float* v, const float* logits, const float* rate,
                                 const float* new_mean, const float* new_std, float alpha, int noise_on, float inv_std, float snr,
                                 float jump_dt, int jump_corrector, float death_prob, const float* z_diff, const float* u_jump,
                                 const float* u_death, const float* z_new, int B, int N, int S) {
    const int F = 3 + S;
    float* zc = (float*)malloc(sizeof(float) * (size_t)B * N * F);   /* noise after delete_dims + adjust_st_batch, [x | one-hot] */
    double gsum = 0.0, nsum = 0.0;
    for (int b = 0; b < B; ++b) {
        const float* zb = z_diff + (size_t)b * N * F;
        float* zj = zc + (size_t)b * N * F;
        const int dim = dims[b];
        for (int i = 0; i < N * 3; ++i) zj[i] = (i / 3) < dim ? zb[i] : 0.0f;
        for (int i = 0; i < N * S; ++i) zj[N * 3 + i] = (i / S) < dim ? zb[N * 3 + i] : 0.0f;
        for (int c = 0; c < 3; ++c) {
            float acc = 0.0f;
            for (int n = 0; n < N; ++n) acc += zj[n * 3 + c];
            const float mean = acc / (float)dim;
            for (int n = 0; n < dim; ++n) zj[n * 3 + c] = zj[n * 3 + c] - mean;
        }
        double g2 = 0.0, n2 = 0.0;
        for (int i = 0; i < N * 3; ++i) { const float sc = -(inv_std * v[(size_t)b * N * 3 + i]); g2 += (double)sc * sc; }
        for (int i = 0; i < N * S; ++i) { const float sc = -(inv_std * logits[(size_t)b * N * S + i]); g2 += (double)sc * sc; }
        for (int i = 0; i < N * F; ++i) n2 += (double)zj[i] * zj[i];
        gsum += (double)(float)sqrt(g2);
        nsum += (double)(float)sqrt(n2);
    }
    const float grad_norm = (float)(gsum / B), noise_norm = (float)(nsum / B);
    float r = snr * noise_norm / grad_norm;
    const float step = r * r * 2.0f * alpha;
    const float sq = sqrtf(2.0f * step);
    for (int b = 0; b < B; ++b) {
        float* xb = x + (size_t)b * N * 3;
        float* ob = onehot + (size_t)b * N * S;
        const float* zj = zc + (size_t)b * N * F;
        const int dim = dims[b];
        const int upd = mask_dims[b] < dim ? mask_dims[b] : dim;   /* slots in [dim, mask_dims) are dead: score = noise = 0 there */
        for (int n = 0; n < upd; ++n) {
            for (int c = 0; c < 3; ++c) {
                const float score = -(inv_std * v[((size_t)b * N + n) * 3 + c]);
                const float inc = noise_on ? step * score + sq * zj[n * 3 + c] : step * score;
                xb[n * 3 + c] = xb[n * 3 + c] + inc;
            }
            for (int s = 0; s < S; ++s) {
                const float score = -(inv_std * logits[((size_t)b * N + n) * S + s]);
                const float inc = noise_on ? step * score + sq * zj[N * 3 + n * S + s] : step * score;
                ob[n * S + s] = ob[n * S + s] + inc;
            }
        }
        adjust_jet(xb, ob, dim, N, S);
        if (!jump_corrector) continue;
        const int born = u_jump[b] < rate[b] * jump_dt && dim < N;
        const int dies = u_death[b] < death_prob && dim > 1;
        int nd = dim;
        if (born) {
            for (int c = 0; c < 3; ++c)
                xb[dim * 3 + c] = new_mean[(size_t)b * F + c] + z_new[(size_t)b * F + c] * softplusf(new_std[(size_t)b * F + c]);
            for (int s = 0; s < S; ++s)
                ob[dim * S + s] = new_mean[(size_t)b * F + 3 + s] + z_new[(size_t)b * F + 3 + s] * softplusf(new_std[(size_t)b * F + 3 + s]);
            nd += 1;
        }
        if (dies) nd -= 1;
        for (int n = nd; n < N; ++n) {   /* delete_dims */
            for (int c = 0; c < 3; ++c) xb[n * 3 + c] = 0.0f;
            for (int s = 0; s < S; ++s) ob[n * S + s] = 0.0f;
        }
        dims[b] = nd;
        adjust_jet(xb, ob, nd, N, S);
    }
    free(zc);
}

void mmbo_trans_sample(const MmbEpicDims* ed, const float* epacked, const MmbTransDims* d, const float* W,
                       float* x, float* onehot, int32_t* dims, const MmbJumpSchedule* sch, const MmbForwardRate* fr,
                       const float* z_diff, const float* u_near, const float* u_jump, const float* z_new, const float* u_death,
                       const int32_t* mask_dims_in, int B, int N) {
    const int S = d->vocab_size, F = 3 + S, R = d->max_particles;
    int32_t* mdims = (int32_t*)malloc(sizeof(int32_t) * (size_t)B);
    memcpy(mdims, mask_dims_in ? mask_dims_in : dims, sizeof(int32_t) * (size_t)B);
    float* dx = (float*)malloc(sizeof(float) * (size_t)B * N * F);
    float* v = (float*)malloc(sizeof(float) * (size_t)B * N * 3);
    float* lg = (float*)malloc(sizeof(float) * (size_t)B * N * S);
    float* rate = (float*)malloc(sizeof(float) * (size_t)B);
    float* xl = (float*)malloc(sizeof(float) * (size_t)B * R);
    float* nl = (float*)malloc(sizeof(float) * (size_t)B * N);
    float* nm = (float*)malloc(sizeof(float) * (size_t)B * F);
    float* ns = (float*)malloc(sizeof(float) * (size_t)B * F);
    float* ts = (float*)malloc(sizeof(float) * (size_t)B);
    for (int step = 0; step < sch->n_steps; ++step) {
        for (int b = 0; b < B; ++b) ts[b] = sch->ts[step];
        mmbo_trans_forward(ed, epacked, d, W, x, onehot, dims, ts, NULL, u_near + (size_t)step * B, fr, B, N,
                           dx, rate, NULL, NULL, xl, nl, NULL, nm, ns);
        for (int b = 0; b < B; ++b) {
            memcpy(v + (size_t)b * N * 3, dx + (size_t)b * N * F, sizeof(float) * (size_t)N * 3);
            memcpy(lg + (size_t)b * N * S, dx + (size_t)b * N * F + (size_t)N * 3, sizeof(float) * (size_t)N * S);
        }
        if (sch->kind && sch->kind[step] == 1)
            mmbo_trans_corrector_update(x, onehot, dims, mdims, v, lg, rate, nm, ns, sch->c_score[step], sch->c_noise[step] != 0.0f,
                                        sch->inv_std[step], sch->corrector_snr, sch->jump_dt, sch->jump_corrector,
                                        sch->death_prob ? sch->death_prob[step] : 0.0f, z_diff + (size_t)step * B * N * F,
                                        u_jump + (size_t)step * B, u_death ? u_death + (size_t)step * B : NULL,
                                        z_new + (size_t)step * B * F, B, N, S);
        else {
            memcpy(mdims, dims, sizeof(int32_t) * (size_t)B);
            mmbo_trans_sampler_update(x, onehot, dims, v, lg, rate, nm, ns, sch->c_decay[step], sch->c_score[step], sch->c_noise[step],
                                      sch->inv_std[step], sch->jump_dt, z_diff + (size_t)step * B * N * F, u_jump + (size_t)step * B,
                                      z_new + (size_t)step * B * F, B, N, S);
        }
    }
    free(dx); free(v); free(lg); free(rate); free(xl); free(nl); free(nm); free(ns); free(ts); free(mdims);
}

/* ------------------------------------------------------------------------------------------ */
/* ParticleClouds.postprocess + compute_4mom + JetClassHighLevelFeatures kinematics / charges
 * (mp/data/particle_clouds/particles.py:85-89,124-156; utils.py:310-337; jets.py:90-107,138-141) */
void mmbo_jet_observables(const float* x, const uint8_t* k, const uint8_t* mask, const float* mean, const float* sd, int B, int N,
                          float* x_phys, int8_t* fc, float* jets) {
    for (int b = 0; b < B; ++b) {
        double px = 0, py = 0, pz = 0, e = 0, mult = 0, q0 = 0, q1 = 0;
        for (int n = 0; n < N; ++n) {
            const size_t p = (size_t)b * N + n;
            const float m = mask[p] ? 1.0f : 0.0f;
            float c[3];
            for (int i = 0; i < 3; ++i) c[i] = (x[p * 3 + i] * (sd ? sd[i] : 1.0f) + (mean ? mean[i] : 0.0f)) * m;
            const int tok = k[p];
            const int flavor = m != 0.0f ? (tok < 2 ? tok : 1 + (tok >> 1)) : 0;
            const int charge = m != 0.0f ? (tok < 2 ? 0 : ((tok & 1) ? 1 : -1)) : 0;
            if (x_phys) for (int i = 0; i < 3; ++i) x_phys[p * 3 + i] = c[i];
            if (fc) { fc[p * 2] = (int8_t)flavor; fc[p * 2 + 1] = (int8_t)charge; }
            px += c[0] * cosf(c[2]); py += c[0] * sinf(c[2]); pz += c[0] * sinhf(c[1]); e += c[0] * coshf(c[1]);
            mult += m; q0 += charge; q1 += charge * c[0];
        }
        if (jets) {
            float* o = jets + (size_t)b * MMB_JET_OBS;
            const float fpx = (float)px, fpy = (float)py, fpz = (float)pz, fe = (float)e;
            const float pt = sqrtf(fmaxf(fpx * fpx + fpy * fpy, 0.0f));
            o[MMB_JET_PX] = fpx; o[MMB_JET_PY] = fpy; o[MMB_JET_PZ] = fpz; o[MMB_JET_E] = fe; o[MMB_JET_PT] = pt;
            o[MMB_JET_M] = sqrtf(fmaxf(fe * fe - fpx * fpx - fpy * fpy - fpz * fpz, 0.0f));
            o[MMB_JET_ETA] = 0.5f * logf((pt + fpz) / (pt - fpz));
            o[MMB_JET_PHI] = atan2f(fpy, fpx);
            o[MMB_JET_MULT] = (float)mult; o[MMB_JET_QTOTAL] = (float)q0; o[MMB_JET_QJET] = (float)q1 / pt;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* sample_noise("GaussNoise") + sample_masks + tokens (mp/data/particle_clouds/utils.py:222-307; particles.py:65-69,111-113),
 * driven by the Philox streams of mmb_sample_source.  Integer outputs are exact functions of the Philox words. */
static void philox_words(uint64_t seed, uint64_t jet, int stream_id, int step, int idx, uint32_t out[4]) {
    uint32_t c[4] = {(uint32_t)idx, (uint32_t)step, (uint32_t)jet, (uint32_t)(jet >> 32) * 4u + (uint32_t)stream_id};
    philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
    memcpy(out, c, sizeof(c));
}
void mmbo_sample_source(float* x, uint8_t* k, uint8_t* mask, int B, int N, float scale, const float* cat_probs, const float* mult_cdf,
                        uint64_t seed, uint64_t jet_offset) {
    float c[4], acc = 0.0f;
    for (int i = 0; i < 4; ++i) { acc = acc + cat_probs[i]; c[i] = acc; }
    for (int b = 0; b < B; ++b) {
        const uint64_t jet = jet_offset + (uint64_t)b;
        int mult = N;
        uint32_t r[4], q[4];
        if (mult_cdf) {
            philox_words(seed, jet, 11, 0, 0, r);
            const float u = (float)(r[0] >> 8) * (1.0f / 16777216.0f);
            mult = 0;
            while (mult < N && !(u < mult_cdf[mult])) ++mult;
        }
        for (int n = 0; n < N; ++n) {
            const size_t p = (size_t)b * N + n;
            float z[3] = {0, 0, 0};
            int tok = 0;
            if (n < mult) {
                philox_words(seed, jet, 9, 0, n, r);
                const float u1 = ((float)(r[0] >> 8) + 1.0f) * (1.0f / 16777216.0f), u2 = (float)(r[1] >> 8) * (1.0f / 16777216.0f);
                const float rad = sqrtf(-2.0f * logf(u1));
                z[0] = rad * cosf(6.283185307179586f * u2); z[1] = rad * sinf(6.283185307179586f * u2);
                const float u3 = ((float)(r[2] >> 8) + 1.0f) * (1.0f / 16777216.0f), u4 = (float)(r[3] >> 8) * (1.0f / 16777216.0f);
                z[2] = sqrtf(-2.0f * logf(u3)) * cosf(6.283185307179586f * u4);
                philox_words(seed, jet, 10, 0, n, q);
                const float uf = (float)(q[0] >> 8) * (1.0f / 16777216.0f);
                const int flavor = uf < c[0] ? 0 : uf < c[1] ? 1 : uf < c[2] ? 2 : uf < c[3] ? 3 : 4;
                tok = flavor < 2 ? flavor : 2 * flavor - 2 + (int)(q[1] >> 31);
            }
            for (int i = 0; i < 3; ++i) x[p * 3 + i] = z[i] * scale;
            k[p] = (uint8_t)tok;
            mask[p] = n < mult;
        }
    }
}

/* ------------------------------------------------------------------------------------------ */
/* Forward half of a training / validation step (multimodal_bridge_matching.py:148-197; bridges.py:23-27,99-104,134-177,233-249) */
void mmbo_sample_bridges(const float* x0, const float* x1, const uint8_t* k0, const uint8_t* k1, const float* ts, float sigma, float gamma,
                         int S, const float* z, const float* u, int B, int N, float* xt, uint8_t* kt) {
    const float neg_s_gamma = (float)(-(double)S * (double)gamma);
    const float inv_s = 1.0f / (float)S, ninv_s = -1.0f / (float)S;
    for (int b = 0; b < B; ++b) {
        const float t = ts[b], omt = 1.0f - t;
        const float w1 = mmbo_expf(neg_s_gamma * (1.0f - t)), w0 = mmbo_expf(neg_s_gamma * (t - 0.0f)), w01 = mmbo_expf(neg_s_gamma * 1.0f);
        for (int n = 0; n < N; ++n) {
            const size_t i = (size_t)b * N + n;
            for (int c = 0; c < 3; ++c) {
                const float lin = t * x1[i * 3 + c] + omt * x0[i * 3 + c];
                xt[i * 3 + c] = lin + sigma * z[i * 3 + c];
            }
            const int a0 = k0[i], a1 = k1[i];
            const float p01 = inv_s + w01 * (ninv_s + (a0 == a1 ? 1.0f : 0.0f));
            float tot = 0.0f;
            for (int k = 0; k < S; ++k) {
                const float pa = inv_s + w1 * (ninv_s + (k == a1 ? 1.0f : 0.0f));
                const float pb = inv_s + w0 * (ninv_s + (k == a0 ? 1.0f : 0.0f));
                tot = tot + (pa * pb) / p01;
            }
            int pick = S - 1;
            float c = 0.0f;
            for (int k = 0; k < S; ++k) {
                const float pa = inv_s + w1 * (ninv_s + (k == a1 ? 1.0f : 0.0f));
                const float pb = inv_s + w0 * (ninv_s + (k == a0 ? 1.0f : 0.0f));
                c = c + ((pa * pb) / p01) / tot;
                if (u[i] < c) { pick = k; break; }
            }
            kt[i] = (uint8_t)pick;
        }
    }
}

void mmbo_absorbing_sample(const float* sp, const uint8_t* target_mask, const float* u, int B, int N, uint8_t* mask_t) {
    for (int b = 0; b < B; ++b)
        for (int n = 0; n < N; ++n) {
            const size_t i = (size_t)b * N + n;
            mask_t[i] = (target_mask[i] || u[i] < sp[b]) ? 1 : 0;
        }
}

void mmbo_bridge_losses(const float* v, const float* logits, const float* x0, const float* x1, const uint8_t* k1, const uint8_t* mask,
                        int B, int N, int S, float* out) {
    double mse = 0, ce = 0, cnt = 0;
    for (size_t i = 0; i < (size_t)B * N; ++i) {
        if (!mask[i]) continue;
        for (int c = 0; c < 3; ++c) {
            const float d = v[i * 3 + c] - (x1[i * 3 + c] - x0[i * 3 + c]);
            mse += (double)d * d;
        }
        double mx = -INFINITY, se = 0;
        for (int s = 0; s < S; ++s) mx = logits[i * S + s] > mx ? logits[i * S + s] : mx;
        for (int s = 0; s < S; ++s) se += exp((double)logits[i * S + s] - mx);
        ce += (mx + log(se)) - logits[i * S + k1[i]];
        cnt += 1;
    }
    out[0] = (float)(mse / cnt); out[1] = (float)(ce / cnt); out[2] = (float)cnt;
}

/* ------------------------------------------------------------------------------------------ */
int mmbo_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

void mmbo_generate(const MmbEpicDims* dims, const float* packed, float* x, uint8_t* k, const uint8_t* mask, const float* context,
                   const MmbStepTable* st, const float* u_jump, uint64_t seed, uint64_t jet_offset,
                   int B, int N, int nthreads) {
    if (!dims_ok(dims)) return;
    const MmbEpicLayout Lo = mmb_epic_layout(dims);
    const int Dc = dims->dim_continuous, S = dims->vocab_size, H = dims->dim_hidden_local, T = dims->dim_time_emb,
              X = dims->dim_context;
#ifdef _OPENMP
    if (nthreads <= 0) nthreads = omp_get_max_threads();
#else
    nthreads = 1;
#endif
#pragma omp parallel num_threads(nthreads)
    {
        Scratch sc = scratch_new(N, H);
        float* v = (float*)malloc(sizeof(float) * (size_t)N * Dc);
        float* lg = (float*)malloc(sizeof(float) * (size_t)N * S);
        uint8_t* mk = (uint8_t*)malloc((size_t)N);
#pragma omp for schedule(dynamic, 4)
        for (int b = 0; b < B; ++b) {
            float* xb = x + (size_t)b * N * Dc;
            uint8_t* kb = k + (size_t)b * N;
            memcpy(mk, mask + (size_t)b * N, (size_t)N);
            float ctx[256];   /* [time embedding of the step | the jet's embedded context] (utils.py:133-170) */
            for (int i = 0; i < X; ++i) ctx[T + i] = context[(size_t)b * X + i];
            for (int s = 0; s < st->n_steps; ++s) {
                memcpy(ctx, st->temb + (size_t)s * T, sizeof(float) * (size_t)T);
                epic_forward_jet(dims, &Lo, packed, xb, kb, mk, ctx, N, v, lg, NULL, &sc);
                for (int n = 0; n < N; ++n) {
                    float u = u_jump ? u_jump[((size_t)s * B + b) * N + n]
                                     : philox_uniform(seed, jet_offset + (uint64_t)b, 0, s, n);
                    update_particle(xb + (size_t)n * Dc, kb + n, mk + n, v + (size_t)n * Dc, lg + (size_t)n * S,
                                    NULL, u, NULL, st->dt, st->bc[s], st->cc[s], 0.0f, Dc, S, MMB_FLAG_MULTIMODAL);
                }
            }
        }
        free(v); free(lg); free(mk);
        scratch_free(&sc);
    }
}
