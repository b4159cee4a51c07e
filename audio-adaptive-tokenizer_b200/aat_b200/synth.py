"""Deterministic synthetic inputs for tests, goldens and bench.

Host-side numpy generators only: this module produces *inputs*, never results.

* ``bursty_speech`` – 16 kHz mono "syllable" audio: Gaussian noise times an
  envelope of voiced bursts U(80,600) ms (Hann-shaped, amplitude U(0.3,1.0))
  separated by pauses U(30,250) ms at a 1e-3 floor.  This is the recipe the
  survey validated to exercise the ragged path of
  ref:src/aat/tokenizer.py:55-183 (50-60 minima / ~33 segments per 16 s).
* ``stationary_noise`` / ``silence`` – the degenerate inputs with no minima.
* ``hubert_frames`` – HuBERT conv-stack output length for a segment of L
  samples, closed form of TF:models/hubert/modeling_hubert.py:675-688.
* ``segment_frame_offsets`` – CSR frame offsets for a list of segment lengths
  under the per-segment-encode convention that
  ref:scripts/mean_hubert_embeddings.py:18-20 consumes.
"""
from __future__ import annotations

import numpy as np

SAMPLING_RATE = 16000


def seed_for(config: int, utterance: int) -> int:
    """Seed convention of SURVEY.md §8d: ``1000*config + utterance``."""
    return 1000 * int(config) + int(utterance)


def bursty_speech(n_samples: int, seed: int, dtype=np.float32) -> np.ndarray:
    rng = np.random.default_rng(seed)
    sr = SAMPLING_RATE
    env = np.full(n_samples, 1e-3, dtype=np.float64)
    pos = 0
    while pos < n_samples:
        burst = int(rng.uniform(80.0, 600.0) * sr / 1000.0)
        amp = rng.uniform(0.3, 1.0)
        shape = np.hanning(burst) * amp + 1e-3
        end = min(n_samples, pos + burst)
        env[pos:end] = shape[: end - pos]
        pos = end + int(rng.uniform(30.0, 250.0) * sr / 1000.0)
    noise = rng.standard_normal(n_samples)
    return (noise * env).astype(dtype)


def stationary_noise(n_samples: int, seed: int, dtype=np.float32) -> np.ndarray:
    return np.random.default_rng(seed).standard_normal(n_samples).astype(dtype)


def silence(n_samples: int, dtype=np.float32) -> np.ndarray:
    return np.zeros(n_samples, dtype=dtype)


def znorm(x: np.ndarray) -> np.ndarray:
    """Call-site normalisation, ref:scripts/audio_tokenization_melspec.py:40."""
    x = np.asarray(x, dtype=np.float64)
    return (x - x.mean()) / (x.std() + 1e-6)


def hubert_frames(n_samples):
    """Frames the HuBERT feature extractor yields for ``n_samples`` samples.

    Kernels (10,3,3,3,3,2,2) / strides (5,2,2,2,2,2,2) collapse to
    ``(n - 400)//320 + 1`` (0 when the segment is shorter than one receptive field).
    """
    n = np.asarray(n_samples, dtype=np.int64)
    return np.maximum((n - 400) // 320 + 1, 0)


def segment_frame_offsets(segment_lengths) -> np.ndarray:
    """CSR offsets (S+1,) int64 in HuBERT-frame units, one entry per segment."""
    n = hubert_frames(np.asarray(segment_lengths, dtype=np.int64))
    off = np.zeros(n.size + 1, dtype=np.int64)
    np.cumsum(n, out=off[1:])
    return off


def mel_frames(n_samples: int, hop_length: int = 160) -> int:
    """``1 + N//hop`` — frame count of the centred STFT, TF:audio_utils.py:778."""
    return 1 + int(n_samples) // int(hop_length)
