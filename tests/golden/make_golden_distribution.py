"""Distribution fixture from the UNMODIFIED reference sampler (container only):

    python tests/golden/make_golden_distribution.py

Unlike every other fixture, nothing is swapped here: ``torch.poisson`` (bridges.py:185) and the rest of
``MultiModalBridgeMatching.simulate_dynamics`` (multimodal_bridge_matching.py:199-216) run exactly as shipped, driven by
torch's own generator.  The reference is run ``N_RUNS`` times on the C2 shape (N = 128, S = 8, 99 solver steps, default
widths, 2048 jets per run) with independent source batches and RNG seeds.  Stream-level reproduction of such a run is
impossible by construction (SURVEY.md §7 "RNG"), so what is committed are the DISTRIBUTIONS of each run — quantile
functions of the particle features, of the per-jet feature sums and of the jet observables the reference's own classes
compute (``ParticleClouds.postprocess`` + ``JetClassHighLevelFeatures``, fastjet substructure disabled), token frequencies
and per-jet flavor multiplicities — plus the weights.  tests/test_gpu_distribution.py generates jets with the CUDA
samplers (one uniform per particle-step instead of S Poisson draws, in-kernel Philox) from the same weights and requires
its distance to the reference runs to lie within the reference's own run-to-run spread (north star: "W1 on pT/eta/phi/jet
mass and on flavor multiplicities within the reference's own seed-to-seed spread").
"""
import json
import os
import sys
import time
from dataclasses import asdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, HERE)
sys.path.insert(0, ROOT)
import make_golden as mg  # noqa: E402,F401  (installs the import shim)

from multimodal_particles.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig  # noqa: E402
from multimodal_particles.data.particle_clouds.jets import JetClassHighLevelFeatures  # noqa: E402
from multimodal_particles.data.particle_clouds.particles import ParticleClouds  # noqa: E402
from multimodal_particles.models.generative.multimodal_bridge_matching import HybridState, MultiModalBridgeMatching  # noqa: E402

from multimodal_particles_b200.databatch import jetclass_like_databatch  # noqa: E402  (synthetic source batch recipe only)

N_RUNS, JETS, NQ = 4, 2048, 257
STATS = {"mean": [1.2, 0.0, 0.0], "std": [0.35, 0.2, 0.2]}   # de-standardisation to a jet-like scale (pT > 0)
JET_OBS = ("pt", "m", "eta", "phi", "Q_total")
WEIGHT_SEED, SOURCE_SEED0, RNG_SEED0 = 0, 9000, 500


def quantiles(a, nq=NQ):
    a = np.asarray(a, np.float64)
    a = a[np.isfinite(a)]
    return np.quantile(a, (np.arange(nq) + 0.5) / nq).astype(np.float32)


def summarise(x, k, mask, prefix, out):
    """x [B,N,3] f32, k [B,N,1] int64, mask [B,N,1] int64 (torch, CPU) -> distribution summaries under ``prefix``."""
    live = mask[..., 0].bool()
    for c in range(3):
        out[f"{prefix}/feat{c}"] = quantiles(x[..., c][live].numpy())
        out[f"{prefix}/jetsum{c}"] = quantiles((x[..., c] * mask[..., 0]).sum(1).numpy())
    tok = k[..., 0][live].numpy()
    out[f"{prefix}/token_freq"] = (np.bincount(tok, minlength=8) / tok.size).astype(np.float32)
    pc = ParticleClouds(dataset=HybridState(None, x.clone(), k.clone(), mask.clone()))
    pc.postprocess(input_continuous="standardize", input_discrete="tokens", stats=STATS)
    JetClassHighLevelFeatures.substructure = lambda self: None
    jets = JetClassHighLevelFeatures(pc)
    for name in JET_OBS:
        out[f"{prefix}/jet_{name}"] = quantiles(getattr(jets, name).numpy())
    assert pc.flavor.shape[-1] == 5   # one-hot (photon, h0, h+-, e+-, mu+-), zero rows on padding
    for f in range(5):   # per-jet multiplicity of each flavor
        out[f"{prefix}/flavor_mult{f}"] = quantiles(((pc.flavor[..., f] == 1) & live).sum(1).numpy())


def main():
    cfg = MultimodalBridgeMatchingConfig.from_yaml("/root/reference/tests/resources/configs_files/config-mbm-test.yaml")
    assert cfg.data.max_num_particles == 128 and cfg.data.vocab_size_features == 8 and cfg.bridge.num_timesteps == 100
    torch.manual_seed(WEIGHT_SEED)
    model = MultiModalBridgeMatching(cfg)
    with torch.no_grad():   # sharpen the random-init heads so tokens and features actually move (as in the trajectory fixtures)
        model.encoder.fc_layer[2].weight.mul_(6.0)
        model.encoder.epic.epic.output_layer.weight_g.mul_(3.0)
    out = dict(config=json.dumps(asdict(cfg)), n_runs=np.int32(N_RUNS), jets=np.int32(JETS), nq=np.int32(NQ),
               stats_mean=np.array(STATS["mean"], np.float32), stats_std=np.array(STATS["std"], np.float32),
               source_seed0=np.int32(SOURCE_SEED0))
    out.update(mg.np_state_dict(model))
    torch.set_num_threads(os.cpu_count() or 1)
    for r in range(N_RUNS):
        batch = jetclass_like_databatch(JETS, 128, generator=torch.Generator().manual_seed(SOURCE_SEED0 + r))
        state = HybridState(None, batch.source_continuous.clone(), batch.source_discrete.clone(), batch.source_mask.clone())
        torch.manual_seed(RNG_SEED0 + r)   # the reference's torch.poisson consumes the global generator
        t0 = time.perf_counter()
        with torch.no_grad():
            final = model.simulate_dynamics(state, batch)   # unmodified: torch.poisson inside TelegraphBridge.solver_step
        dt = time.perf_counter() - t0
        moved = (final.discrete != batch.source_discrete)[batch.source_mask.bool()].float().mean().item()
        summarise(final.continuous, final.discrete, batch.source_mask, f"run{r}", out)
        print(f"run {r}: {JETS} jets in {dt:.1f} s ({JETS / dt:.0f} jets/s, {torch.get_num_threads()} threads), tokens moved {moved:.3f}")
    path = os.path.join(HERE, "mbm_distribution.npz")
    np.savez_compressed(path, **out)
    print(f"mbm_distribution: {os.path.getsize(path) / 1024:.0f} KiB")


if __name__ == "__main__":
    main()
