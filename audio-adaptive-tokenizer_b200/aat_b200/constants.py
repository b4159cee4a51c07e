"""Host-side constant tables of the tokenizer (computed once per tokenizer, in float64).

The reference builds them with ``transformers.audio_utils.mel_filter_bank(..., norm="slaney",
mel_scale="slaney")`` and ``window_function(n_fft, "hann")`` (ref:src/aat/tokenizer.py:41-51;
TF:audio_utils.py:263-332, 356-375, 453-544, 560-620).  They are evaluated here with the same
numpy expression order so that the ``mel_filters`` / ``window_fn`` attributes are bit-identical
(tests/test_host_logic.py checks this against transformers when it is importable), and they are
uploaded to the device as they are.
"""
from __future__ import annotations

import numpy as np


def _hertz_to_mel_slaney(freq):
    min_log_hertz, min_log_mel = 1000.0, 15.0
    logstep = 27.0 / np.log(6.4)
    if isinstance(freq, np.ndarray):
        mels = 3.0 * freq / 200.0
        log_region = freq >= min_log_hertz
        mels[log_region] = min_log_mel + np.log(freq[log_region] / min_log_hertz) * logstep
        return mels
    mels = 3.0 * freq / 200.0
    if freq >= min_log_hertz:
        mels = min_log_mel + np.log(freq / min_log_hertz) * logstep
    return mels


def _mel_to_hertz_slaney(mels):
    min_log_hertz, min_log_mel = 1000.0, 15.0
    logstep = np.log(6.4) / 27.0
    freq = 200.0 * mels / 3.0
    if isinstance(mels, np.ndarray):
        log_region = mels >= min_log_mel
        freq[log_region] = min_log_hertz * np.exp(logstep * (mels[log_region] - min_log_mel))
    elif mels >= min_log_mel:
        freq = min_log_hertz * np.exp(logstep * (mels - min_log_mel))
    return freq


def mel_filter_bank_slaney(num_frequency_bins: int, num_mel_filters: int, min_frequency: float,
                           max_frequency: float, sampling_rate: int) -> np.ndarray:
    """(num_frequency_bins, num_mel_filters) float64 triangular bank, Slaney scale + area norm."""
    if num_frequency_bins < 2:
        raise ValueError(f"Require num_frequency_bins: {num_frequency_bins} >= 2")
    if min_frequency > max_frequency:
        raise ValueError(f"Require min_frequency: {min_frequency} <= max_frequency: {max_frequency}")
    mel_min = _hertz_to_mel_slaney(min_frequency)
    mel_max = _hertz_to_mel_slaney(max_frequency)
    mel_freqs = np.linspace(mel_min, mel_max, num_mel_filters + 2)
    filter_freqs = _mel_to_hertz_slaney(mel_freqs)
    fft_freqs = np.linspace(0, sampling_rate // 2, num_frequency_bins)
    filter_diff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down_slopes = -slopes[:, :-2] / filter_diff[:-1]
    up_slopes = slopes[:, 2:] / filter_diff[1:]
    mel_filters = np.maximum(np.zeros(1), np.minimum(down_slopes, up_slopes))
    enorm = 2.0 / (filter_freqs[2 : num_mel_filters + 2] - filter_freqs[:num_mel_filters])
    mel_filters *= np.expand_dims(enorm, 0)
    return mel_filters


def hann_window_periodic(window_length: int) -> np.ndarray:
    """``window_function(n, "hann")``: periodic Hann, float64."""
    return np.hanning(window_length + 1)[:-1]
