/*
 * mmb_oracle.h — CPU restatement (plain C) of the Multimodal-Bridges generation hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing in the product path (multimodal_particles_b200/, the C-ABI
 * library) may link or call this; only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline
 * legs use it, as the checker (and as the timed CPU baseline), never as the thing shipped.
 *
 * Parity status: PINNED — against outputs of the unmodified reference executed in the build
 * container with injected uniforms (tests/golden/make_golden.py writes tests/golden/*.npz;
 * tests/test_oracle_golden.py checks this file against them).  The reference's own tests hold no
 * golden vectors for this path.
 *
 * Signatures deliberately mirror include/mmbridge.h so that parity tests pass the same buffers to
 * both sides (host pointers here, device pointers there).
 */
#ifndef MMB_ORACLE_H
#define MMB_ORACLE_H

#include "../include/mmbridge.h"

#ifdef __cplusplus
extern "C" {
#endif

float mmbo_expf(float x);
float mmbo_expf_dn(float x);   /* with gradual underflow (trans-dimensional token rule) */

/* host-side step table with libm, following mbm.py:203-211, utils.py:183-198, bridges.py:125-130,
 * 218-231.  The product computes the same table with torch ops; tests compare the two. */
void mmbo_step_table(int num_timesteps, float time_eps, int S, float gamma, float gamma_absorb, int T,
                     float* t, float* temb, float* bc, float* cc, float* sp, float* dt);

void mmbo_epic_forward(const MmbEpicDims* dims, const float* packed,
                       const float* x, const uint8_t* k, const uint8_t* mask,
                       const float* temb, int temb_stride, int B, int N,
                       float* v_out, float* logits_out, float* hidden_out);

void mmbo_bridge_update(float* x, uint8_t* k, uint8_t* mask,
                        const float* v, const float* logits, const float* absorb_logit,
                        const float* u_jump, const float* u_absorb,
                        float dt, float bc, float cc, float sp,
                        int B, int N, int Dc, int S, int flags);

void mmbo_philox_uniforms(float* u, uint64_t seed, uint64_t jet_offset, int stream_id,
                          int n_steps, int B, int N);

/* MultiModalBridgeMatching.simulate_dynamics (mbm.py:199-216); OpenMP over jets (nthreads<=0: all);
 * context [B][dims->dim_context]: the jets' embedded context features (mbm.py:143-144), NULL without */
void mmbo_generate(const MmbEpicDims* dims, const float* packed, float* x, uint8_t* k, const uint8_t* mask, const float* context,
                   const MmbStepTable* steps, const float* u_jump, uint64_t seed, uint64_t jet_offset,
                   int B, int N, int nthreads);

/*
 * Absorbing-rate head of AbsorbingGenerator (mp/models/generative/absorbing/absorbing_flows.py:94-131;
 * gsdm.py:34-66,142-168): Linear(H+2 -> C) on [last local hidden, one_hot(mask)], n_blocks x
 * (ResnetBlock, AttnBlock) over the N particle slots, Linear(C->C), Linear(C->1).
 * Weight blob (fp32, row-major [out][in], each followed by its bias), in this order:
 *   proj_in [C][H+2]+[C];  per block: norm1 g[C] b[C]; conv1 [C][C]+[C]; norm2 g b; conv2 [C][C]+[C];
 *   attn norm g b; q [C][C]+[C]; k; v; proj_out;   then pre_rate [C][C]+[C]; post_rate [1][C]+[1].
 * tbias [B or 1][n_blocks][C] = temb_proj_b(swish(temb_net(timestep_embedding(1000 t)))) computed by the
 * caller (per-step constants at generation time); stride 0 = shared by all jets.
 */
size_t mmbo_absorb_head_floats(int H, int C, int n_blocks);
void mmbo_absorb_head(const float* W, int H, int C, int n_heads, int n_blocks,
                      const float* hidden /*[B,N,H]*/, const uint8_t* mask /*[B,N]*/,
                      const float* tbias, int tbias_stride, int B, int N, float* logit_out /*[B,N]*/);

/*
 * Trans-dimensional jump diffusion (signatures mirror include/mmbridge.h: mmb_trans_*).
 * TransdimensionalEPiC.forward (transdimensional_model.py:245-426) and JumpSampler.sample (sampler.py:157-324,
 * uniform dt, no corrector, no conditioning) with the random draws injected.
 */
size_t mmbo_trans_floats(const MmbTransDims* d);
void mmbo_trans_tokens(const float* onehot, int B, int N, int S, uint8_t* k);
void mmbo_trans_time_terms(const MmbTransDims* d, const float* W, const float* ts, int B, int T,
                           float* temb_epic, float* tb1, float* tb2);
float mmbo_trans_rate(const float* logits, int R, int xt_dim, const MmbForwardRate* fr, float t);
void mmbo_trans_forward(const MmbEpicDims* ed, const float* epacked, const MmbTransDims* d, const float* W,
                        const float* x, const float* onehot, const int32_t* dims, const float* ts,
                        const int32_t* nearest_in, const float* u_nearest, const MmbForwardRate* fr, int B, int N,
                        float* d_xt, float* rate, float* auto_mean, float* auto_std, float* x0_dim_logits,
                        float* near_atom_logits, int32_t* nearest_out, float* new_mean, float* new_std);
void mmbo_trans_sampler_update(float* x, float* onehot, int32_t* dims, const float* v, const float* logits, const float* rate,
                               const float* new_mean, const float* new_std,
                               float c_decay, float c_score, float c_noise, float inv_std, float jump_dt,
                               const float* z_diff, const float* u_jump, const float* z_new, int B, int N, int S);
void mmbo_trans_corrector_update(float* x, float* onehot, int32_t* dims, const int32_t* mask_dims, const float* v, const float* logits, const float* rate,
                                 const float* new_mean, const float* new_std, float alpha, int noise_on, float inv_std, float snr,
                                 float jump_dt, int jump_corrector, float death_prob, const float* z_diff, const float* u_jump,
                                 const float* u_death, const float* z_new, int B, int N, int S);
void mmbo_trans_sample(const MmbEpicDims* ed, const float* epacked, const MmbTransDims* d, const float* W,
                       float* x, float* onehot, int32_t* dims, const MmbJumpSchedule* sch, const MmbForwardRate* fr,
                       const float* z_diff, const float* u_near, const float* u_jump, const float* z_new, const float* u_death,
                       const int32_t* mask_dims_in, int B, int N);

/* post-processing + jet observables (mmb_jet_observables); jet sums accumulated in double */
void mmbo_jet_observables(const float* x, const uint8_t* k, const uint8_t* mask, const float* mean, const float* sd, int B, int N,
                          float* x_phys, int8_t* fc, float* jets);

/* source state from the Philox streams of mmb_sample_source (tokens / masks exact, normals through libm) */
void mmbo_sample_source(float* x, uint8_t* k, uint8_t* mask, int B, int N, float scale, const float* cat_probs, const float* mult_cdf,
                        uint64_t seed, uint64_t jet_offset);

/* forward half of a training / validation step, draws injected (mmb_sample_bridges, mmb_absorbing_sample, mmb_bridge_losses) */
void mmbo_sample_bridges(const float* x0, const float* x1, const uint8_t* k0, const uint8_t* k1, const float* ts, float sigma, float gamma,
                         int S, const float* z, const float* u, int B, int N, float* xt, uint8_t* kt);
void mmbo_absorbing_sample(const float* sp, const uint8_t* target_mask, const float* u, int B, int N, uint8_t* mask_t);
void mmbo_bridge_losses(const float* v, const float* logits, const float* x0, const float* x1, const uint8_t* k1, const uint8_t* mask,
                        int B, int N, int S, float* out);

int mmbo_max_threads(void);

#ifdef __cplusplus
}
#endif
#endif
