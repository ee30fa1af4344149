"""Oracle post-processing / jet observables against the fixture produced by the reference
(tests/golden/make_golden_observables.py: ParticleClouds.postprocess + JetClassHighLevelFeatures)."""
import os

import numpy as np

import oracle_lib as ol

GOLD = os.path.join(os.path.dirname(__file__), "golden", "observables.npz")
KIN = ("px", "py", "pz", "e", "pt", "m", "eta", "phi")


def check_against_fixture(z, x_phys, fc, jets):
    np.testing.assert_allclose(x_phys, z["continuous"], rtol=1e-6, atol=1e-6)
    flavor = np.eye(5, dtype=np.int8)[fc[..., 0]] * z["mask"].astype(np.int8)
    assert np.array_equal(flavor, z["flavor"]) and np.array_equal(fc[..., 1:2], z["charge"])
    assert np.array_equal(np.concatenate([flavor, fc[..., 1:2]], -1), z["discrete"])
    live = z["multiplicity"][:, 0] > 0          # an empty jet is 0/0 in eta and Q_jet on both sides
    for i, name in enumerate(KIN):
        ref, got = z[f"jet_{name}"], jets[:, i]
        if name == "m":   # sqrt of a cancelling difference: compare m^2 on the scale of e^2
            np.testing.assert_allclose(got ** 2, ref ** 2, rtol=0, atol=2e-5 * float((z["jet_e"] ** 2).max()), err_msg=name)
        elif name == "eta":
            np.testing.assert_allclose(got[live], ref[live], rtol=2e-4, atol=2e-4, err_msg=name)
        else:
            np.testing.assert_allclose(got, ref, rtol=2e-5, atol=2e-5, err_msg=name)
    assert np.array_equal(jets[:, 8].astype(np.int32), z["multiplicity"][:, 0])
    np.testing.assert_allclose(jets[:, 9], z["Q_total"], atol=0)
    np.testing.assert_allclose(jets[live, 10], z["Q_jet"][live], rtol=2e-5, atol=2e-6)
    assert np.isnan(jets[~live, 6]).all() and np.isnan(z["jet_eta"][~live]).all()


def test_oracle_observables_match_reference():
    z = np.load(GOLD)
    x_phys, fc, jets = ol.jet_observables(z["x"], z["k"][..., 0], z["mask"][..., 0], {"mean": z["mean"], "std": z["std"]})
    check_against_fixture(z, x_phys, fc, jets)
