"""``AudioWaveform`` — the value object of ref:src/aat/audio.py:3-15, kept verbatim in behaviour."""
from __future__ import annotations

import numpy as np


class AudioWaveform:
    def __init__(self, waveform, sampling_rate):
        self.waveform: np.ndarray = waveform
        self.sampling_rate = sampling_rate

        assert len(waveform.shape) == 1, "channel dim is not supported for waveform"

        self.duration_seconds: float = self.waveform.shape[-1] / self.sampling_rate

    def assert_sampling_rate(self, expected_sapmling_rate: int):
        assert self.sampling_rate == expected_sapmling_rate, (
            f"Audio sampling rate mismatch: ausio_sampling_rate={self.sampling_rate}, "
            f"expected_sapmling_rate={expected_sapmling_rate}"
        )
