"""ORACLE (test infrastructure, not product code) — faithful CPU port.

This file restates the reference's *own* code for the tokenization front end,
statement by statement, and calls the very same third-party functions the
reference calls (``transformers.audio_utils.spectrogram``,
``scipy.signal.argrelextrema``, ``numpy``, ``torch.Tensor.mean``).  It exists
so that parity tests, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs have a checker and a CPU baseline
that travels to the GPU box (``/root/reference`` does not).

Nothing under ``audio-adaptive-tokenizer_b200/`` may import this module.

Pinned against the live reference by ``tests/golden/make_golden.py`` (run in
the build container with ``/root/reference/src`` importable) and
``tests/test_oracle_golden.py``.  Versions of record: transformers 5.5.0,
scipy 1.18.1, numpy 2.3.5, torch 2.11.0.

Reference sites restated here
  ref:src/aat/audio.py:3-15                    -> AudioWaveform
  ref:src/aat/tokenizer.py:15-53               -> RefTokenizer.__init__
  ref:src/aat/tokenizer.py:55-92               -> find_amplitude_minimas
  ref:src/aat/tokenizer.py:94-105              -> milliseconds_to_frames / pads
  ref:src/aat/tokenizer.py:107-119             -> get_melspec
  ref:src/aat/tokenizer.py:121-139             -> pretokenize
  ref:src/aat/tokenizer.py:141-183             -> process_segments_boarders
  ref:src/aat/tokenizer.py:185-200             -> tokenize
  ref:scripts/mean_hubert_embeddings.py:19-20  -> mean_pool_segments
  ref:scripts/audio_tokenization_melspec.py:40 -> znorm
"""
from __future__ import annotations

from typing import List

import numpy as np


class AudioWaveform:
    """ref:src/aat/audio.py:3-15."""

    def __init__(self, waveform, sampling_rate):
        self.waveform = waveform
        self.sampling_rate = sampling_rate
        assert len(waveform.shape) == 1, "channel dim is not supported for waveform"
        self.duration_seconds = self.waveform.shape[-1] / self.sampling_rate

    def assert_sampling_rate(self, expected):
        assert self.sampling_rate == expected, (
            f"Audio sampling rate mismatch: ausio_sampling_rate={self.sampling_rate}, "
            f"expected_sapmling_rate={expected}"
        )


class RefTokenizer:
    """Port of ``AdaptiveAudioAmplitudeTokenizer`` (ref:src/aat/tokenizer.py:14-200)."""

    def __init__(
        self,
        running_mean_points=12,
        min_segment_duration_milliseconds=125,
        max_segment_duration_milliseconds=1500,
        n_fft=400,
        hop_length=160,
        num_mel_filters=64,
        sampling_rate=16000,
        max_amplitude_for_minima=15,
    ):
        from transformers.audio_utils import mel_filter_bank, window_function

        self.running_mean_points = running_mean_points
        self.max_amplitude_for_minima = max_amplitude_for_minima
        self.n_fft = n_fft
        self.hop_length = hop_length
        self.num_mel_filters = num_mel_filters
        self.sampling_rate = sampling_rate
        self.min_segment_duration_milliseconds = min_segment_duration_milliseconds
        self.max_segment_duration_milliseconds = max_segment_duration_milliseconds
        self.min_segment_frames = self.milliseconds_to_frames(min_segment_duration_milliseconds)
        self.max_segment_frames = self.milliseconds_to_frames(max_segment_duration_milliseconds)
        self.mel_filters = mel_filter_bank(
            num_frequency_bins=1 + self.n_fft // 2,
            num_mel_filters=num_mel_filters,
            min_frequency=0.0,
            max_frequency=8000.0,
            sampling_rate=sampling_rate,
            norm="slaney",
            mel_scale="slaney",
        )
        self.window_fn = window_function(self.n_fft, "hann")

    # ref:src/aat/tokenizer.py:94-95
    def milliseconds_to_frames(self, milliseconds: int) -> int:
        return int(milliseconds * self.sampling_rate / 1000)

    # ref:src/aat/tokenizer.py:102-105
    def right_pad_waveform_with_zeros(self, waveform):
        padded = np.zeros([self.min_segment_frames])
        padded[: waveform.shape[-1]] = waveform
        return padded

    # ref:src/aat/tokenizer.py:107-119
    def get_melspec(self, audio_waveform: np.ndarray) -> np.ndarray:
        from transformers.audio_utils import spectrogram

        return spectrogram(
            audio_waveform,
            self.window_fn,
            frame_length=self.n_fft,
            hop_length=self.hop_length,
            power=2.0,
            mel_filters=self.mel_filters,
            log_mel="log10",
        )

    # ref:src/aat/tokenizer.py:55-92
    def find_amplitude_minimas(self, melspec: np.ndarray, return_intermediates: bool = False):
        from scipy.signal import argrelextrema

        amp = -10 * melspec.mean(axis=0)

        def running_mean(x, N):
            cumsum = np.cumsum(x)
            return (cumsum[N:] - cumsum[:-N]) / float(N)

        rm = running_mean(amp, self.running_mean_points)

        def greater_eps(x1, x2):
            return x1 > x2 + 1e-5

        minimas = argrelextrema(rm, greater_eps)[0]
        minimas = minimas[rm[minimas] > self.max_amplitude_for_minima]
        if return_intermediates:
            return minimas, amp, np.cumsum(amp), rm
        return minimas

    # ref:src/aat/tokenizer.py:121-139
    def pretokenize(self, audio_waveform: np.ndarray, melspec=None):
        if melspec is None:
            melspec = self.get_melspec(audio_waveform)
        minimas = self.find_amplitude_minimas(melspec)
        boarders = (minimas * self.hop_length).tolist() + [audio_waveform.shape[-1]]
        return boarders, melspec

    # ref:src/aat/tokenizer.py:141-183
    def process_segments_boarders(self, audio_waveform: np.ndarray, segments_boarders) -> List[np.ndarray]:
        segments: List[np.ndarray] = []
        prev = 0
        for b in segments_boarders:
            length = b - prev
            if length < self.min_segment_frames:
                continue
            if length > self.max_segment_frames:
                split_sizes = [self.max_segment_frames] * (length // self.max_segment_frames)
                split_sizes = np.cumsum(split_sizes)
                gap = length - split_sizes[-1]
                if gap == 0:
                    split_sizes = split_sizes[:-1]
                elif gap < self.min_segment_frames:
                    split_sizes[-1] = length - self.min_segment_frames
                segments.extend(np.split(audio_waveform[prev:b], split_sizes))
            else:
                segments.append(audio_waveform[prev:b])
            prev = b
        if prev != audio_waveform.shape[-1]:
            segments.append(self.right_pad_waveform_with_zeros(audio_waveform[prev:]))
        return segments

    # ref:src/aat/tokenizer.py:185-200
    def tokenize(self, audio_waveform_sr: AudioWaveform, melspec=None):
        audio_waveform_sr.assert_sampling_rate(self.sampling_rate)
        wf = audio_waveform_sr.waveform
        boarders, melspec = self.pretokenize(wf, melspec=melspec)
        segments = self.process_segments_boarders(wf, boarders)
        assert len(segments) < 300
        assert sum(x.shape[-1] for x in segments) >= wf.shape[-1]
        return [AudioWaveform(s, audio_waveform_sr.sampling_rate) for s in segments], melspec

    # convenience for long-form audio where tokenize()'s <300 assert fires
    # (SURVEY.md §5 "Long-context"): pretokenize + process_segments_boarders.
    def segment_lengths(self, audio_waveform: np.ndarray, melspec=None):
        boarders, melspec = self.pretokenize(audio_waveform, melspec=melspec)
        segs = self.process_segments_boarders(audio_waveform, boarders)
        return [int(s.shape[-1]) for s in segs], boarders, melspec


def znorm(x: np.ndarray) -> np.ndarray:
    """ref:scripts/audio_tokenization_melspec.py:40."""
    return (x - x.mean()) / (x.std() + 1e-6)


def mean_pool_segments(embeddings_list):
    """ref:scripts/mean_hubert_embeddings.py:19-20 — list of [1, n_i, D] -> [1, S, D] fp32."""
    import torch

    mean_embeddings = [x.mean(dim=1, keepdim=True).to(torch.float32) for x in embeddings_list]
    return torch.cat(mean_embeddings, dim=1)


def mean_pool_csr(emb, seg_off):
    """Same pooling, for a packed ``emb [T, D]`` + CSR ``seg_off [S+1]`` layout."""
    import torch

    emb = torch.as_tensor(emb)
    off = [int(v) for v in seg_off]
    parts = [emb[off[i] : off[i + 1]].unsqueeze(0) for i in range(len(off) - 1)]
    if not parts:
        return torch.zeros(1, 0, emb.shape[-1], dtype=torch.float32)
    return mean_pool_segments(parts)


def dataset_mean(pooled_list):
    """Row A9 of SURVEY.md §8a: unweighted mean over all pooled vectors -> [D] (fp64 accumulate)."""
    import torch

    cat = torch.cat([p.reshape(-1, p.shape[-1]) for p in pooled_list], dim=0).to(torch.float64)
    return cat.mean(dim=0).to(torch.float32)
