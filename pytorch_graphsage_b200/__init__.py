"""
pytorch_graphsage_b200 -- B200-native sample -> gather -> aggregate -> project engine behind the plug-in API of
bkj/pytorch-graphsage (nn_modules.py registries + GSSupervised.forward).

    from pytorch_graphsage_b200 import sampler_lookup, prep_lookup, aggregator_lookup, GSSupervised, set_seeds

Everything numeric runs in libgsage_b200.so (include/gsage_b200.h); importing this package on a machine
without the built library or without a CUDA device succeeds, *using* it raises (there is no CPU fallback).
"""

from ._lib import GsageError, lib, launch_count           # noqa: F401
from .graph import GraphCSR                                # noqa: F401
from .rng import DeviceMT19937, default_rng, set_seeds     # noqa: F401
from .operators import (sampler_lookup, prep_lookup, aggregator_lookup, UniformNeighborSampler,   # noqa: F401
                        SparseUniformNeighborSampler, IdentityPrep, NodeEmbeddingPrep, LinearPrep, MeanAggregator,
                        PoolAggregator, MaxPoolAggregator, MeanPoolAggregator, AttentionAggregator, LSTMAggregator)
from .model import GSSupervised, FeatureTable             # noqa: F401
from .parallel import FusedAdam, FlatGradBucket            # noqa: F401
from . import ops, synth, problem                          # noqa: F401

__all__ = ['sampler_lookup', 'prep_lookup', 'aggregator_lookup', 'GSSupervised', 'FeatureTable', 'GraphCSR',
           'DeviceMT19937', 'default_rng', 'set_seeds', 'ops', 'synth', 'lib', 'launch_count', 'GsageError', 'FusedAdam',
           'FlatGradBucket']
