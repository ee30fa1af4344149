"""CPU: the oracle's sampler (one uniform per particle-step, Form B, Philox) against the distributions of the UNMODIFIED
reference sampler (torch.poisson tau-leaping, bridges.py:185-194) — SURVEY.md §A.4 claims the two are the same law; this pins
it on the C2 shape with the reference's own seed-to-seed spread as the yardstick."""
import numpy as np
import torch

import distribution_lib as dl
import oracle_lib as ol
from multimodal_particles_b200.databatch import jetclass_like_databatch


def load():
    from multimodal_particles_b200.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig
    from multimodal_particles_b200.multimodal_bridge_matching import MultiModalBridgeMatching
    import json
    z = np.load(dl.GOLD)
    cfg = MultimodalBridgeMatchingConfig.from_dict(json.loads(str(z["config"])))
    model = MultiModalBridgeMatching(cfg)
    model.load_state_dict({k[3:]: torch.from_numpy(z[k]) for k in z.files if k.startswith("sd/")}, strict=True)
    return z, cfg, model


def test_fixture_is_self_consistent():
    z = np.load(dl.GOLD)
    runs, keys = dl.reference_runs(z)
    assert len(runs) >= 2 and int(z["jets"]) >= 2048 and "jet_m" in keys and "flavor_mult2" in keys and "feat0" in keys
    for r in runs:
        assert abs(float(r["token_freq"].sum()) - 1.0) < 1e-5
        assert np.all(np.diff(r["jet_pt"]) >= 0) and np.all(r["jet_pt"] > 0)
    # every reference run passes the gate against the others' spread (leave-one-in sanity of the yardstick)
    for r in runs:
        dl.gate(z, r, "reference run")


def test_oracle_sampler_matches_unmodified_reference_in_distribution():
    z, cfg, model = load()
    dims, packed = ol.packed_model(model)
    B = int(z["jets"])
    batch = jetclass_like_databatch(B, 128, generator=torch.Generator().manual_seed(int(z["source_seed0"]) + 100))
    x0, k0, m0 = batch.source_continuous.numpy(), batch.source_discrete[..., 0].numpy(), batch.source_mask[..., 0].numpy()
    x, k = ol.generate(dims, packed, x0, k0, m0, model.step_table(), seed=31, jet_offset=0)
    stats = {"mean": z["stats_mean"].tolist(), "std": z["stats_std"].tolist()}
    _, fc, jets = ol.jet_observables(x, k, m0.astype(np.uint8), stats)
    cand = dl.summarise(x, k, m0, fc[..., 0], jets, int(z["nq"]))
    rows = dl.gate(z, cand, "oracle (Form B, Philox)")
    assert len(rows) >= 17
    # and the gate has teeth: the same sampler with a skewed uniform stream (u^2 favours the low-index thresholds) is rejected
    u = ol.philox_uniforms(31, 0, model.step_table().n_steps, B, 128) ** 2
    xb, kb = ol.generate(dims, packed, x0, k0, m0, model.step_table(), u_jump=u)
    _, fcb, jb = ol.jet_observables(xb, kb, m0.astype(np.uint8), stats)
    try:
        dl.gate(z, dl.summarise(xb, kb, m0, fcb[..., 0], jb, int(z["nq"])), "biased")
    except AssertionError:
        return
    raise AssertionError("a sampler fed with squared uniforms passed the distribution gate")
