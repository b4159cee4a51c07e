"""aat_b200 — B200-native tokenization front end of mrsndmn/audio-adaptive-tokenizer.

Host-side mirror of the reference's interface for the hot path (log-mel -> adaptive segment
boundaries -> ragged mean-pool), calling hand-written sm_100a kernels through the C ABI in
``include/aat_b200.h``.  Importing the package is cheap and needs neither torch nor a GPU; the first
compute call loads ``libaat_b200.so`` and fails loudly if it (or a CUDA device) is missing.
"""
from .audio import AudioWaveform
from .tokenizer import AdaptiveAudioAmplitudeTokenizer, PackedBatch

__all__ = ["AudioWaveform", "AdaptiveAudioAmplitudeTokenizer", "PackedBatch", "mean_pool_segments", "DatasetMean"]


def __getattr__(name):  # torch is only needed by the pooling module
    if name in ("mean_pool_segments", "DatasetMean"):
        from . import pooling

        return getattr(pooling, name)
    raise AttributeError(name)
