"""Deterministic synthetic inputs for tests, goldens and bench.

This module produces *inputs*, never results.  The numpy generators below run on the host (tests, goldens,
the oracle's subset); ``device_bursty_batch`` / ``device_normal`` generate the same recipe ON the device from a
counter-based RNG through the C ABI (``aat_synth_*``), for the benchmark's dataset-scale job.

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
* ``utterance_frame_offsets`` – the same for one whole-utterance encoding
  (convention (ii)).
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


def utterance_frame_offsets(segment_lengths, n_samples: int) -> np.ndarray:
    """CSR offsets (S+1,) int64 over the rows of ONE whole-utterance encoding (convention (ii)): the
    encoder ran once over the ``n_samples`` samples (``T = hubert_frames(n_samples)`` rows) and the segment that starts
    at sample ``s`` starts at row ``min(s // 320, T)`` — the collator's ``// hop_length``
    (ref:src/aat/training/collate.py:340) with the encoder's stride.  The lengths add up to at least ``n_samples``
    (ref:src/aat/tokenizer.py:195), so the last offset is ``T``.  Host-side twin of ``aat_utterance_frame_csr``."""
    lengths = np.asarray(segment_lengths, dtype=np.int64).reshape(-1)
    if int(lengths.sum()) < int(n_samples):
        raise ValueError("segment lengths must cover the utterance (sum >= n_samples)")
    rows = int(hubert_frames(n_samples))
    starts = np.zeros(lengths.size + 1, dtype=np.int64)
    np.cumsum(lengths, out=starts[1:])
    return np.minimum(starts // 320, rows)


def mel_frames(n_samples: int, hop_length: int = 160) -> int:
    """``1 + N//hop`` — frame count of the centred STFT, TF:audio_utils.py:778."""
    return 1 + int(n_samples) // int(hop_length)


# ---------------------------------------------------------------------------------------- on the device
def device_bursty_batch(batch, seed_base: int, utt_index_base: int = 0, out=None):
    """Packed float32 waveform of ``batch`` (a :class:`~aat_b200.tokenizer.PackedBatch`) generated on its device:
    utterance ``b`` is bursty speech from seed ``seed_base + utt_index_base + b`` (Philox; the seed convention of
    :func:`seed_for` when ``seed_base = 1000 * config``).  Enqueued on the current stream; returns ``out``."""
    import ctypes

    import torch

    from . import _cabi

    lib = _cabi.lib()
    if out is None:
        out = torch.empty(batch.total_samples, dtype=torch.float32, device=batch.device)
    if out.dtype != torch.float32 or out.numel() != batch.total_samples or not out.is_contiguous():
        raise TypeError("out must be a contiguous float32 CUDA tensor with the plan's total_samples entries")
    ws = getattr(batch, "_synth_ws", None)
    if ws is None:
        ws = batch._synth_ws = torch.empty(int(lib.aat_synth_workspace_bytes(batch.handle)) + 16, dtype=torch.uint8,
                                           device=batch.device)
    stream = ctypes.c_void_p(torch.cuda.current_stream(batch.device).cuda_stream)
    _cabi.check(lib.aat_synth_waveforms(batch.ctx.handle, batch.handle, int(seed_base), int(utt_index_base),
                                        out.data_ptr(), ws.data_ptr(), stream))
    return out


def device_normal(out, seed: int):
    """Fill the float32 CUDA tensor ``out`` with N(0, 1) (random-init HuBERT-shaped embeddings); element ``i`` is a
    pure function of ``(seed, i)``."""
    import ctypes

    import torch

    from . import _cabi
    from .context import default_context

    if out.dtype != torch.float32 or not out.is_cuda or not out.is_contiguous():
        raise TypeError("out must be a contiguous float32 CUDA tensor")
    ctx = default_context(out.device.index)
    stream = ctypes.c_void_p(torch.cuda.current_stream(out.device).cuda_stream)
    _cabi.check(_cabi.lib().aat_synth_normal(ctx.handle, out.data_ptr(), out.numel(), int(seed), stream))
    return out
