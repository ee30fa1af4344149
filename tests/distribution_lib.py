"""Distribution gate against the UNMODIFIED reference sampler (tests/golden/mbm_distribution.npz, produced by
tests/golden/make_golden_distribution.py from /root/reference with torch.poisson untouched).

A candidate sample (oracle port or CUDA kernels; one uniform per particle-step, Philox) is summarised exactly like the
reference runs were — quantile functions of particle features, per-jet sums, jet observables (pT, mass, eta, phi, charge),
per-jet flavor multiplicities, token frequencies — and, for every quantity, its mean 1-D Wasserstein distance to the
reference runs must not exceed ``FACTOR`` x the largest distance between two reference runs (the reference's own
seed-to-seed spread).  W1 between two quantile functions on the same probability grid = mean |Qa - Qb|.
"""
import itertools
import os

import numpy as np

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "mbm_distribution.npz")
FACTOR = 1.5
JET_OBS = ("pt", "m", "eta", "phi", "Q_total")
JET_COLUMNS = ("px", "py", "pz", "e", "pt", "m", "eta", "phi", "multiplicity", "Q_total", "Q_jet")


def quantiles(a, nq):
    a = np.asarray(a, np.float64)
    a = a[np.isfinite(a)]
    return np.quantile(a, (np.arange(nq) + 0.5) / nq).astype(np.float32)


def summarise(x, k, mask, flavor, jets, nq):
    """x [B,N,3], k [B,N], mask [B,N] (numpy), flavor [B,N] int (tokens_to_physics), jets [B,11] -> {quantity: summary}"""
    live = mask.astype(bool)
    out = {}
    for c in range(3):
        out[f"feat{c}"] = quantiles(x[..., c][live], nq)
        out[f"jetsum{c}"] = quantiles((x[..., c] * mask).sum(1), nq)
    tok = k[live]
    out["token_freq"] = (np.bincount(tok, minlength=8) / tok.size).astype(np.float32)
    for name in JET_OBS:
        out[f"jet_{name}"] = quantiles(jets[:, JET_COLUMNS.index(name)], nq)
    for f in range(5):
        out[f"flavor_mult{f}"] = quantiles(((flavor == f) & live).sum(1), nq)
    return out


def distance(a, b):
    return float(np.abs(np.asarray(a, np.float64) - np.asarray(b, np.float64)).mean())


def reference_runs(z):
    n = int(z["n_runs"])
    keys = sorted(k.split("/", 1)[1] for k in z.files if k.startswith("run0/"))
    return [{k: z[f"run{r}/{k}"] for k in keys} for r in range(n)], keys


def gate(z, cand, label=""):
    """-> list of (quantity, candidate distance, reference spread); raises AssertionError outside the spread."""
    runs, keys = reference_runs(z)
    rows, bad = [], []
    for q in keys:
        pair = [distance(a[q], b[q]) for a, b in itertools.combinations(runs, 2)]
        d = float(np.mean([distance(cand[q], r[q]) for r in runs]))
        rows.append((q, d, max(pair), float(np.mean(pair))))
        if d > FACTOR * max(pair) + 1e-9:
            bad.append(rows[-1])
    assert not bad, f"{label}: outside the reference's seed-to-seed spread (quantity, distance, max pair, mean pair): {bad}"
    return rows
