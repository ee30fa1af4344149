"""ctypes access to oracle/libmmb_oracle.so — the CPU checker (test infrastructure only)."""
import ctypes
import json
import os
import subprocess
from types import SimpleNamespace

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
ORACLE_DIR = os.path.join(ROOT, "oracle")
ORACLE_SO = os.path.join(ORACLE_DIR, "libmmb_oracle.so")

from multimodal_particles_b200._native import EpicDims  # noqa: E402  (plain ctypes struct)
from multimodal_particles_b200.steptable import CStepTable  # noqa: E402

_lib = None
_fp = ctypes.POINTER(ctypes.c_float)
_u8p = ctypes.POINTER(ctypes.c_uint8)


def lib():
    global _lib
    if _lib is None:
        src_mtime = max(os.path.getmtime(os.path.join(ORACLE_DIR, f)) for f in ("mmb_oracle.c", "mmb_oracle.h"))
        if not os.path.exists(ORACLE_SO) or (os.path.getmtime(ORACLE_SO) < src_mtime and os.access(ORACLE_DIR, os.W_OK)):
            subprocess.run(["make", "-C", ORACLE_DIR], check=True, capture_output=True)
        L = ctypes.CDLL(ORACLE_SO)
        L.mmbo_expf.restype, L.mmbo_expf.argtypes = ctypes.c_float, [ctypes.c_float]
        L.mmbo_max_threads.restype = ctypes.c_int
        _lib = L
    return _lib


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def u8(a):
    return np.ascontiguousarray(a, dtype=np.uint8)


def _p(a, ty=_fp):
    return None if a is None else a.ctypes.data_as(ty)


def epic_forward(dims: EpicDims, packed, x, k, mask, temb, want_hidden=False):
    B, N, Dc = x.shape
    x, k, mask, temb, packed = f32(x), u8(k).reshape(B, N), u8(mask).reshape(B, N), f32(temb), f32(packed)
    v = np.empty((B, N, Dc), np.float32)
    logits = np.empty((B, N, dims.vocab_size), np.float32)
    hidden = np.empty((B, N, dims.dim_hidden_local), np.float32) if want_hidden else None
    width = dims.dim_time_emb + dims.dim_context   # a row: [time embedding | embedded context of the jet]
    stride = 0 if temb.reshape(-1, width).shape[0] == 1 and B != 1 else width
    lib().mmbo_epic_forward(ctypes.byref(dims), _p(packed), _p(x), _p(k, _u8p), _p(mask, _u8p), _p(temb),
                            ctypes.c_int(stride), B, N, _p(v), _p(logits), _p(hidden))
    return (v, logits, hidden) if want_hidden else (v, logits)


def bridge_update(x, k, mask, v, logits, u_jump, dt, bc, cc, absorb_logit=None, u_absorb=None, sp=0.0, flags=0):
    """Returns updated copies (x, k, mask)."""
    B, N, Dc = x.shape
    S = logits.shape[-1] if logits is not None else 1
    x, k, mask = f32(x).copy(), u8(k).reshape(B, N).copy(), u8(mask).reshape(B, N).copy()
    args = [None if a is None else f32(a) for a in (v, logits, absorb_logit, u_jump, u_absorb)]
    lib().mmbo_bridge_update(_p(x), _p(k, _u8p), _p(mask, _u8p), *[_p(a) for a in args],
                             ctypes.c_float(dt), ctypes.c_float(bc), ctypes.c_float(cc), ctypes.c_float(sp),
                             B, N, Dc, S, flags)
    return x, k, mask


def generate(dims: EpicDims, packed, x, k, mask, table, u_jump=None, seed=0, jet_offset=0, nthreads=0, context=None):
    B, N, Dc = x.shape
    x, k, mask, packed = f32(x).copy(), u8(k).reshape(B, N).copy(), u8(mask).reshape(B, N), f32(packed)
    u = None if u_jump is None else f32(u_jump)
    ctx = None if context is None else f32(context).reshape(B, dims.dim_context)
    assert (ctx is None) == (dims.dim_context == 0)
    ct = CStepTable.from_table(table)
    lib().mmbo_generate(ctypes.byref(dims), _p(packed), _p(x), _p(k, _u8p), _p(mask, _u8p), _p(ctx), ctypes.byref(ct), _p(u),
                        ctypes.c_uint64(seed), ctypes.c_uint64(jet_offset), B, N, nthreads)
    return x, k


def philox_uniforms(seed, jet_offset, n_steps, B, N, stream_id=0):
    u = np.empty((n_steps, B, N), np.float32)
    lib().mmbo_philox_uniforms(_p(u), ctypes.c_uint64(seed), ctypes.c_uint64(jet_offset), stream_id, n_steps, B, N)
    return u


def step_table_libm(num_timesteps, time_eps, S, gamma, T, gamma_absorb=None):
    n = num_timesteps - 1
    t, temb, bc, cc, sp = (np.empty(n, np.float32), np.empty((n, T), np.float32), np.empty(n, np.float32),
                           np.empty(n, np.float32), np.empty(n, np.float32))
    dt = ctypes.c_float()
    lib().mmbo_step_table(num_timesteps, ctypes.c_float(time_eps), S, ctypes.c_float(gamma),
                          ctypes.c_float(gamma_absorb or 0.0), T, _p(t), _p(temb), _p(bc), _p(cc),
                          _p(sp) if gamma_absorb is not None else None, ctypes.byref(dt))
    return SimpleNamespace(t=t, temb=temb, bc=bc, cc=cc, sp=sp, dt=dt.value)


# ---- golden fixtures ------------------------------------------------------------------------
def _ns(d):
    return SimpleNamespace(**{k: _ns(v) if isinstance(v, dict) else v for k, v in d.items()})


def load_mbm_golden(path):
    """-> (fixture npz, config namespace, model of this repo with the fixture's weights)."""
    from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import (
        MultimodalBridgeMatchingConfig)
    from multimodal_particles_b200.multimodal_bridge_matching import MultiModalBridgeMatching
    z = np.load(path)
    cfg = MultimodalBridgeMatchingConfig.from_dict(json.loads(str(z["config"])))
    model = MultiModalBridgeMatching(cfg)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model.load_state_dict(sd, strict=True)
    return z, cfg, model


def packed_model(model):
    enc = model.encoder
    head = enc.fc_layer if enc.add_discrete_head else None
    dims = enc.epic.epic_dims(head[0].out_features if head is not None else 0)
    return dims, enc.epic.pack_weights(head).numpy()


# ---- absorbing flow -----------------------------------------------------------------------------
def absorb_head(packed, H, C, n_heads, n_blocks, hidden, mask, tbias):
    """tbias [B or 1, n_blocks, C] -> rate logits [B, N]"""
    B, N, _ = hidden.shape
    hidden, mask, tbias, packed = f32(hidden), u8(mask).reshape(B, N), f32(tbias), f32(packed)
    lib().mmbo_absorb_head_floats.restype = ctypes.c_size_t
    assert packed.size == lib().mmbo_absorb_head_floats(H, C, n_blocks), "head blob size"
    out = np.empty((B, N), np.float32)
    stride = 0 if tbias.shape[0] == 1 and B != 1 else n_blocks * C
    lib().mmbo_absorb_head(_p(packed), H, C, n_heads, n_blocks, _p(hidden), _p(mask, _u8p), _p(tbias), stride, B, N, _p(out))
    return out


def load_absorbing_golden(path):
    from multimodal_particles_b200.absorbing_flows import AbsorbingFlow
    from multimodal_particles_b200.config_classes.absorbing_flows_config import AbsorbingConfig
    z = np.load(path)
    cfg = AbsorbingConfig.from_dict(json.loads(str(z["config"])))
    model = AbsorbingFlow(cfg)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model.load_state_dict(sd, strict=True)
    return z, cfg, model


def absorbing_step(model, trunk, head_blob, x, k, mask, temb, tbias, u_jump, u_absorb, dt, bc, cc, sp):
    """One step of AbsorbingFlow.simulate_dynamics with oracle pieces -> (x, k, mask, heads)."""
    dims, packed = trunk
    g = model.generator
    v, logits, hidden = epic_forward(dims, packed, x, k, mask, temb, want_hidden=True)
    a = absorb_head(head_blob, g.encoder_output_dim_local, g.transformer_dim, g.n_heads, g.n_attn_blocks, hidden, mask, tbias)
    x, k, mask = bridge_update(x, k, mask, v, logits, u_jump, dt, bc, cc, absorb_logit=a, u_absorb=u_absorb, sp=sp, flags=1)
    return x, k, mask, (v, logits, a)


def absorbing_trunk(model):
    g = model.generator
    head = g.discrete_head_mlp if g.add_discrete_head else None
    dims = g.epic.epic_dims(head[0].out_features if head is not None else 0)
    return dims, g.epic.pack_weights(head).numpy()


# ---- trans-dimensional jump diffusion ---------------------------------------------------------------
from multimodal_particles_b200._native import ForwardRate, JumpSchedule, TransDims  # noqa: E402

_i32p = ctypes.POINTER(ctypes.c_int32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)


def load_trans_golden(path):
    from multimodal_particles_b200.config_classes.transdimensional_unconditional_config import TransdimensionalEpicConfig
    from multimodal_particles_b200.transdimensional import TransdimensionalJumpDiffusion
    z = np.load(path)
    cfg = TransdimensionalEpicConfig.from_dict(json.loads(str(z["config"])))
    model = TransdimensionalJumpDiffusion(cfg)
    sd = {k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}
    model.load_state_dict(sd, strict=True)
    return z, cfg, model


def trans_packed(model):
    """-> (EpicDims, trunk blob, TransDims, heads blob) of a TransdimensionalJumpDiffusion of this repo."""
    m = model.net.model
    return m.epic.epic_dims(0), m.epic.pack_weights(None).numpy(), m.trans_dims(), m.pack_heads().numpy()


def trans_tokens(onehot):
    B, N, S = onehot.shape
    k = np.empty((B, N), np.uint8)
    lib().mmbo_trans_tokens(_p(f32(onehot)), B, N, S, _p(k, _u8p))
    return k


def trans_forward(packed, x, onehot, dims, ts, fr: ForwardRate, nearest=None, u_nearest=None):
    edims, eblob, tdims, tblob = packed
    B, N, _ = x.shape
    S, R, F = tdims.vocab_size, tdims.max_particles, 3 + tdims.vocab_size
    lib().mmbo_trans_floats.restype = ctypes.c_size_t
    assert tblob.size == lib().mmbo_trans_floats(ctypes.byref(tdims)), "trans blob size"
    x, onehot, dims, ts = f32(x), f32(onehot), i32(dims), f32(ts)
    out = SimpleNamespace(d_xt=np.empty((B, N * F), np.float32), rate=np.empty(B, np.float32),
                          auto_mean=np.empty((B, N * F), np.float32), auto_std=np.empty((B, N * F), np.float32),
                          x0_dim_logits=np.empty((B, R), np.float32), near_atom_logits=np.empty((B, N), np.float32),
                          nearest=np.empty(B, np.int32), new_mean=np.empty((B, F), np.float32), new_std=np.empty((B, F), np.float32))
    near = None if nearest is None else i32(nearest)
    un = None if u_nearest is None else f32(u_nearest)
    lib().mmbo_trans_forward(ctypes.byref(edims), _p(f32(eblob)), ctypes.byref(tdims), _p(f32(tblob)), _p(x), _p(onehot),
                             _p(dims, _i32p), _p(ts), _p(near, _i32p), _p(un), ctypes.byref(fr), B, N,
                             _p(out.d_xt), _p(out.rate), _p(out.auto_mean), _p(out.auto_std), _p(out.x0_dim_logits),
                             _p(out.near_atom_logits), _p(out.nearest, _i32p), _p(out.new_mean), _p(out.new_std))
    return out


def trans_sampler_update(x, onehot, dims, v, logits, rate, new_mean, new_std, c_decay, c_score, c_noise, inv_std, jump_dt,
                         z_diff, u_jump, z_new):
    """Returns updated copies (x, onehot, dims)."""
    B, N, _ = x.shape
    S = onehot.shape[-1]
    x, onehot, dims = f32(x).copy(), f32(onehot).copy(), i32(dims).copy()
    cf = ctypes.c_float
    lib().mmbo_trans_sampler_update(_p(x), _p(onehot), _p(dims, _i32p), _p(f32(v)), _p(f32(logits)), _p(f32(rate)), _p(f32(new_mean)),
                                    _p(f32(new_std)), cf(c_decay), cf(c_score), cf(c_noise), cf(inv_std), cf(jump_dt),
                                    _p(f32(z_diff)), _p(f32(u_jump)), _p(f32(z_new)), B, N, S)
    return x, onehot, dims


def trans_corrector_update(x, onehot, dims, mask_dims, v, logits, rate, new_mean, new_std, alpha, noise_on, inv_std, snr, jump_dt,
                           jump_corrector, death_prob, z_diff, u_jump, u_death, z_new):
    """Returns updated copies (x, onehot, dims)."""
    B, N, _ = x.shape
    S = onehot.shape[-1]
    x, onehot, dims = f32(x).copy(), f32(onehot).copy(), i32(dims).copy()
    cf = ctypes.c_float
    lib().mmbo_trans_corrector_update(_p(x), _p(onehot), _p(dims, _i32p), _p(i32(mask_dims), _i32p), _p(f32(v)), _p(f32(logits)),
                                      _p(f32(rate)), _p(f32(new_mean)), _p(f32(new_std)), cf(alpha), int(noise_on), cf(inv_std), cf(snr),
                                      cf(jump_dt), int(jump_corrector), cf(death_prob), _p(f32(z_diff)), _p(f32(u_jump)),
                                      _p(f32(u_death)), _p(f32(z_new)), B, N, S)
    return x, onehot, dims


def trans_sample(packed, x, onehot, dims, sched, fr: ForwardRate, z_diff, u_near, u_jump, z_new, u_death=None, mask_dims=None):
    """Returns final copies (x, onehot, dims)."""
    edims, eblob, tdims, tblob = packed
    B, N, _ = x.shape
    x, onehot, dims = f32(x).copy(), f32(onehot).copy(), i32(dims).copy()
    cs = JumpSchedule.from_schedule(sched)
    lib().mmbo_trans_sample(ctypes.byref(edims), _p(f32(eblob)), ctypes.byref(tdims), _p(f32(tblob)), _p(x), _p(onehot),
                            _p(dims, _i32p), ctypes.byref(cs), ctypes.byref(fr), _p(f32(z_diff)), _p(f32(u_near)), _p(f32(u_jump)),
                            _p(f32(z_new)), _p(None if u_death is None else f32(u_death)),
                            _p(None if mask_dims is None else i32(mask_dims), _i32p), B, N)
    return x, onehot, dims


def trans_initial_state(z_init, N, S):
    """x_T ~ N(0,I) -> one centred particle per jet (sampler.py:170-183): the first particle's continuous features
    are x - x = 0, its one-hot block is the raw draw; everything else is deleted."""
    B = z_init.shape[0]
    x = np.zeros((B, N, 3), np.float32)
    oh = np.zeros((B, N, S), np.float32)
    oh[:, 0, :] = z_init[:, N * 3: N * 3 + S]
    return x, oh, np.ones(B, np.int32)


# ---- post-processing + jet observables ------------------------------------------------------------------
def jet_observables(x, k, mask, stats=None):
    B, N, _ = x.shape
    x, k, mask = f32(x), u8(k).reshape(B, N), u8(mask).reshape(B, N)
    mean = None if stats is None else f32(stats["mean"][:3])
    sd = None if stats is None else f32(stats["std"][:3])
    x_phys, fc, jets = np.empty((B, N, 3), np.float32), np.empty((B, N, 2), np.int8), np.empty((B, 11), np.float32)
    lib().mmbo_jet_observables(_p(x), _p(k, _u8p), _p(mask, _u8p), _p(mean), _p(sd), B, N, _p(x_phys),
                               fc.ctypes.data_as(ctypes.POINTER(ctypes.c_int8)), _p(jets))
    return x_phys, fc, jets


# ---- source state ---------------------------------------------------------------------------------------
def sample_source(B, N, scale, cat_probs, mult_cdf, seed, jet_offset):
    x, k, mask = np.empty((B, N, 3), np.float32), np.empty((B, N), np.uint8), np.empty((B, N), np.uint8)
    cdf = None if mult_cdf is None else f32(mult_cdf)
    lib().mmbo_sample_source(_p(x), _p(k, _u8p), _p(mask, _u8p), B, N, ctypes.c_float(scale), _p(f32(cat_probs)), _p(cdf),
                             ctypes.c_uint64(seed), ctypes.c_uint64(jet_offset))
    return x, k, mask


# ---- forward half of a training / validation step -----------------------------------------------------
def sample_bridges(x0, x1, k0, k1, t, sigma, gamma, S, z, u):
    B, N, _ = x0.shape
    xt, kt = np.empty((B, N, 3), np.float32), np.empty((B, N), np.uint8)
    lib().mmbo_sample_bridges(_p(f32(x0)), _p(f32(x1)), _p(u8(k0).reshape(B, N), _u8p), _p(u8(k1).reshape(B, N), _u8p), _p(f32(t)),
                              ctypes.c_float(sigma), ctypes.c_float(gamma), S, _p(f32(z)), _p(f32(u)), B, N, _p(xt), _p(kt, _u8p))
    return xt, kt


def absorbing_sample(sp, target_mask, u):
    B, N = target_mask.shape
    out = np.empty((B, N), np.uint8)
    lib().mmbo_absorbing_sample(_p(f32(sp)), _p(u8(target_mask), _u8p), _p(f32(u)), B, N, _p(out, _u8p))
    return out


def bridge_losses(v, logits, x0, x1, k1, mask):
    B, N, S = logits.shape
    out = np.empty(3, np.float32)
    lib().mmbo_bridge_losses(_p(f32(v)), _p(f32(logits)), _p(f32(x0)), _p(f32(x1)), _p(u8(k1).reshape(B, N), _u8p),
                             _p(u8(mask).reshape(B, N), _u8p), B, N, S, _p(out))
    return out
