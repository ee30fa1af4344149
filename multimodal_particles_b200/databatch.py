"""Batch convention of the generation path (mp/data/particle_clouds/jets_dataloader.py:18-28,
239-271): any indexable whose ``batch[0]`` has length B, with attributes ``source_continuous``,
``source_discrete``, ``source_mask`` (and optionally ``context_*``)."""
from collections import namedtuple
from typing import NamedTuple

import torch


class MultimodalDatabatch(NamedTuple):
    source_continuous: torch.Tensor
    source_discrete: torch.Tensor
    source_mask: torch.Tensor
    target_continuous: torch.Tensor
    target_discrete: torch.Tensor
    target_mask: torch.Tensor
    context_continuous: torch.Tensor
    context_discrete: torch.Tensor


ParticleData = namedtuple("ParticleData", ["source_continuous", "source_discrete", "source_mask",
                                           "target_continuous", "target_discrete", "target_mask"])


def random_databatch(config, generator=None) -> ParticleData:
    """Random batch with the config's shapes and the reference's dtypes
    (JetsDataloaderModule.random_databatch, jets_dataloader.py:239-271): uniform features,
    uniform tokens, Bernoulli(1/2) int64 masks; ``target_discrete`` is float there too."""
    d = config.data
    shape = (d.batch_size, d.max_num_particles)
    rand = lambda *s: torch.rand(*s, generator=generator)
    randint = lambda lo, hi, s: torch.randint(lo, hi, s, generator=generator)
    return ParticleData(
        source_continuous=rand(*shape, d.dim_features_continuous),
        source_discrete=randint(0, d.vocab_size_features, (*shape, d.dim_features_discrete)),
        source_mask=randint(0, 2, (*shape, 1)),
        target_continuous=rand(*shape, d.dim_features_continuous),
        target_discrete=rand(*shape, d.dim_features_discrete),
        target_mask=randint(0, 2, (*shape, 1)),
    )


def jetclass_like_databatch(batch_size, max_num_particles=128, dim_continuous=3, vocab_size=8,
                            mean_multiplicity=45.0, std_multiplicity=18.0, generator=None) -> ParticleData:
    """Synthetic JetClass-shaped source batch: multiplicity
    m ~ clamp(round(N(45,18)),1,N) with a prefix mask as ``sample_masks`` builds it
    (mp/data/particle_clouds/utils.py:283-286), Gaussian-noise source features and uniform tokens,
    both zeroed on padding (particles.py:66-69)."""
    g = generator
    n = max_num_particles
    mult = torch.randn(batch_size, generator=g) * std_multiplicity + mean_multiplicity
    mult = mult.round().clamp(1, n).long()
    mask = (torch.arange(n)[None, :] < mult[:, None]).long().unsqueeze(-1)
    cont = torch.randn(batch_size, n, dim_continuous, generator=g) * mask
    disc = torch.randint(0, vocab_size, (batch_size, n, 1), generator=g) * mask
    return ParticleData(cont, disc, mask, cont.clone(), disc.clone().float(), mask.clone())
