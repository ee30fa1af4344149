"""Forward fixture of TransdimensionalEPiC with ``encoder.rate_use_x0_pred = False`` from the reference (container only):

    python tests/golden/make_golden_trans_direct.py

With that flag ``post_rate_proj`` has ONE output, the birth rate is ``softplus(rate logit) * forward_rate(t)`` and
``x0_dim_logits`` are zeros (transdimensional_model.py:185-188, 326-332).  Same inputs and patches as the forward part of
make_golden_trans.py; no shipped config sets the flag.
"""
import json
import os
import sys
from dataclasses import asdict

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_trans as mt  # noqa: E402
import make_golden as mg  # noqa: E402


def main():
    cfg = mt.TransdimensionalEpicConfig()
    cfg.data.return_type = "list"
    cfg.encoder.rate_use_x0_pred = False
    N = cfg.data.max_num_particles = 16
    S = cfg.data.vocab_size_features
    dm = mt.fake_datamodule(cfg)
    torch.manual_seed(311)
    with mt.quiet():
        model = mt.TransdimensionalJumpDiffusion(cfg, dm)
    net = model.net
    assert net.model.post_rate_proj.weight.shape[0] == 1
    with torch.no_grad():
        net.model.post_rate_proj.weight.mul_(6.0)
        net.model.near_atom_proj.weight.mul_(8.0)
        net.model.vec_weighting_proj.weight.mul_(4.0)
        net.model.post_auto_proj.weight.mul_(4.0)
        net.model.epic.epic.output_layer.weight_g.mul_(3.0)
    gs = dm.graphical_structure
    gs.max_problem_dim = N
    mt.EpsilonPrecond.forward = mt.patched_precond_forward
    g = torch.Generator().manual_seed(312)
    B = 6
    dims = torch.tensor([1, 3, 16, 8, 2, 5])
    m = (torch.arange(N)[None] < dims[:, None]).float().unsqueeze(-1)
    x = torch.randn(B, N, 3, generator=g) * m
    x = x - (x.sum(1, keepdim=True) / dims.view(B, 1, 1)) * m
    oh = torch.randn(B, N, S, generator=g) * m
    ts = torch.tensor([0.05, 0.5, 0.999, 0.2, 0.75, 0.011])
    nearest = torch.tensor([0, 2, 7, 0, 1, 4])
    st = mt.StructuredDataBatch([x.clone(), oh.clone()], dims.clone(), dm.observed, dm.exist, dm.is_onehot, gs)
    with mt.quiet(), torch.no_grad():
        D, rate, (am, asd), x0l, nal = net(st, ts, forward_rate=model.forward_rate, predict="eps", nearest_atom=nearest)
    assert float(x0l.abs().max()) == 0.0 and rate.shape == (B, 1)
    fr = model.forward_rate
    out = dict(config=json.dumps(asdict(cfg)), forward_rate=np.array([0.0, fr.get_scalar(), fr.offset, fr.rate_cut_t], dtype=np.float64))
    out.update({"fwd/x": x.numpy(), "fwd/onehot": oh.numpy(), "fwd/dims": dims.numpy().astype(np.int32), "fwd/ts": ts.numpy(),
                "fwd/nearest": nearest.numpy().astype(np.int32), "fwd/d_xt": D.numpy(), "fwd/rate": rate.view(-1).numpy(),
                "fwd/auto_mean": am.numpy(), "fwd/auto_std": asd.numpy(), "fwd/x0_dim_logits": x0l.numpy(),
                "fwd/near_atom_logits": nal.numpy()})
    out.update(mg.np_state_dict(model))
    path = os.path.join(HERE, "trans_direct.npz")
    np.savez_compressed(path, **out)
    print(f"trans_direct: {os.path.getsize(path) / 1024:.0f} KiB; rates {rate.view(-1).tolist()}")


if __name__ == "__main__":
    main()
