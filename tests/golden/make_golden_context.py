"""Fixtures for models WITH context features, from the unmodified reference (container only):

    python tests/golden/make_golden_context.py

``MultiModalBridgeMatching.forward`` hands ``batch.context_continuous`` / ``batch.context_discrete`` to the encoder
(multimodal_bridge_matching.py:143-144); ``InputEmbeddings`` embeds them and appends them to the time embedding
(architectures/utils.py:155-170); the resulting per-jet ``context`` vector enters ``global_0``, ``fc_global1`` and ``fc_local1``
(architectures/epic.py:187-189, 226-238).  No shipped config switches this on, so the cases are built here:

* ``mbm_ctx``: config-mbm-test.yaml + 2 continuous context features through a Linear (the reference calls that kind
  "Embedding") to width 3;
* ``mbm_ctx_id``: 5 continuous context features, no embedding (identity), odd widths, no skip, no head.

Discrete context features cannot be exercised: the constructor stores the module as ``embedding_context_discrete``
(utils.py:100-106) while ``forward`` looks for ``embedding_discrete_context`` (utils.py:161), so the embedding is never
appended and ``global_0`` fails with a shape error for any ``dim_emb_context_discrete`` > 0 (tried here: "mat1 and mat2 shapes
cannot be multiplied (5x51 and 55x16)").

Same recording as ``make_golden.mbm_case``: whole trajectories with injected jump uniforms, heads at selected steps.
"""
import os
import sys
from collections import namedtuple

import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden as mg  # noqa: E402  (installs the import shim)

from multimodal_particles.config_classes.multimodal_bridge_matching_config import MultimodalBridgeMatchingConfig  # noqa: E402

# the reference's data batches are namedtuples: indexed (mbm.py:211 ``len(batch[0])``) and read by attribute (mbm.py:143-144)
Batch = namedtuple("Batch", ["source_continuous", "context_continuous", "context_discrete"])

REF_CFG = "/root/reference/tests/resources/configs_files/config-mbm-test.yaml"


def main():
    cfg = MultimodalBridgeMatchingConfig.from_yaml(REF_CFG)
    d, e = cfg.data, cfg.encoder
    d.dim_context_continuous = 2
    e.embedding_context_continuous, e.dim_emb_context_continuous = "Embedding", 3
    cfg.bridge.num_timesteps = 40
    B, N, S = 5, 128, 8
    g = torch.Generator().manual_seed(31)
    mask = mg.prefix_masks([128, 1, 45, 70, 17], N)
    x0 = torch.randn(B, N, 3, generator=g) * mask
    k0 = torch.randint(0, S, (B, N, 1), generator=g) * mask
    cc = torch.randn(B, 2, generator=g) * 1.5
    batch = Batch(x0, cc, None)
    mg.mbm_case("mbm_ctx", cfg, x0, k0, mask, seed=131, snap_steps={0, 20, 38}, batch=batch,
                extra=dict(context_continuous=cc.numpy()))

    cfg = MultimodalBridgeMatchingConfig.from_yaml(REF_CFG)
    d, e = cfg.data, cfg.encoder
    d.dim_context_continuous = 5
    e.dim_hidden_glob, e.dim_hidden_local, e.dim_emb_time, e.num_blocks = 19, 24, 14, 3
    e.dim_emb_features_continuous, e.dim_emb_features_discrete = 12, 10
    e.skip_connection, e.add_discrete_head = False, False
    d.max_num_particles, d.vocab_size_features = 37, 5
    cfg.bridge.num_timesteps = 12
    B, N, S = 3, 37, 5
    g = torch.Generator().manual_seed(32)
    mask = mg.prefix_masks([37, 20, 5], N)
    cc = torch.randn(B, 5, generator=g)
    x0 = torch.randn(B, N, 3, generator=g) * mask
    batch = Batch(x0, cc, None)
    mg.mbm_case("mbm_ctx_id", cfg, x0, torch.randint(0, S, (B, N, 1), generator=g) * mask,
                mask, seed=132, snap_steps={0, 5, 10}, batch=batch, extra=dict(context_continuous=cc.numpy()))


if __name__ == "__main__":
    main()
