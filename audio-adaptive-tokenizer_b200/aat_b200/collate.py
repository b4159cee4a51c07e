"""GPU versions of the collator-side data movement either side of the hot path.

The reference's ``TokenizedAudioWaveformCollator.__call__`` builds these tensors with a double Python loop
in DataLoader workers (ref:src/aat/training/collate.py:242-253, 291-346; "todo vectorize", :248); here they
are produced on the device from the boundary kernel's outputs, so the training collator can stay
GPU-resident.  All functions enqueue on the current torch stream and return CUDA tensors.
"""
from __future__ import annotations

import ctypes

from . import _cabi
from .tokenizer import PackedBatch


def _stream(device=None):
    import torch

    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def _raise_on_status(status, what):
    import torch

    bad = torch.nonzero(status < 0).flatten()
    if bad.numel():
        code = int(status[bad[0]].item())
        raise _cabi.AatError(code, f"{what}: utterance {int(bad[0])} (the reference raises here: segment longer than "
                                   "the tile, boarders not increasing, or a slice past the padded waveform)")


def normalize_waveforms(batch: PackedBatch, wave, mode: str = "zscore", out_dtype=None, return_stats: bool = False):
    """Per-utterance normalisation of a packed waveform tensor.

    mode="zscore": ``(x - x.mean()) / (x.std() + 1e-6)`` — what the reference applies before ``get_melspec`` /
    ``tokenize`` (ref:scripts/audio_tokenization_melspec.py:40, ref:src/aat/training/collate.py:135-152).
    mode="w2v2": ``(x - x.mean()) / sqrt(x.var() + 1e-7)`` in float32 — the Wav2Vec2 feature extractor's
    ``zero_mean_unit_var_norm`` (ref:src/aat/training/collate.py:301)."""
    import torch

    modes = {"zscore": _cabi.AAT_NORM_ZSCORE, "w2v2": _cabi.AAT_NORM_W2V2}
    if mode not in modes:
        raise ValueError(f"mode must be one of {sorted(modes)}")
    codes = {torch.float32: _cabi.AAT_F32, torch.float64: _cabi.AAT_F64}
    if wave.dtype not in codes or not wave.is_cuda or not wave.is_contiguous() or wave.numel() != batch.total_samples:
        raise TypeError("wave must be a contiguous packed CUDA tensor (float32 or float64) matching the plan")
    if out_dtype is None:
        out_dtype = torch.float32 if mode == "w2v2" else torch.float64
    out = torch.empty(batch.total_samples, dtype=out_dtype, device=wave.device)
    stats = torch.empty(2 * batch.n_utts, dtype=torch.float64, device=wave.device)
    _cabi.check(_cabi.lib().aat_normalize(batch.ctx.handle, batch.handle, wave.data_ptr(), codes[wave.dtype], modes[mode],
                                          out.data_ptr(), codes[out_dtype], stats.data_ptr(), _stream(wave.device)))
    return (out, stats.view(batch.n_utts, 2)) if return_stats else out


def pad_segment_boarders(batch: PackedBatch, s_max=None, check: bool = True):
    """``_make_padded_segments_boarders`` on the device: ``(segments_boarders_padded, attention_mask)``,
    both ``[B, S_max]`` int64.  ``s_max=None`` reads the largest segment count back from the device (one sync).
    ``check`` reads the status back (one sync) and raises when an utterance has more than ``s_max`` segments."""
    import torch

    if s_max is None:
        s_max = int(batch.seg_count.max().item())
    padded = torch.empty((batch.n_utts, s_max), dtype=torch.int64, device=batch.device)
    mask = torch.empty_like(padded)
    status = torch.zeros(batch.n_utts, dtype=torch.int32, device=batch.device)
    _cabi.check(_cabi.lib().aat_pad_segment_boarders(batch.ctx.handle, batch.handle, batch.seg_len.data_ptr(),
                                                     batch.seg_count.data_ptr(), s_max, padded.data_ptr(),
                                                     mask.data_ptr(), status.data_ptr(), _stream(batch.device)))
    if check:
        _raise_on_status(status, "pad_segment_boarders")
    return padded, mask


def scatter_segments(batch: PackedBatch, wave_padded, boarders_padded, max_segment_frames: int, with_mask: bool = True,
                     check: bool = True):
    """``batched_segments [B, S, max_segment_frames]`` (+ ``segments_waveforms_mask``) from the padded
    ``input_values [B, N_max]`` float32 and the padded boarders (ref:src/aat/training/collate.py:321-335)."""
    import torch

    if wave_padded.dtype != torch.float32 or wave_padded.dim() != 2 or not wave_padded.is_contiguous():
        raise TypeError("wave_padded must be a contiguous float32 [B, N_max] CUDA tensor")
    B, s_max = boarders_padded.shape
    out = torch.empty((B, s_max, max_segment_frames), dtype=torch.float32, device=wave_padded.device)
    mask = torch.empty_like(out) if with_mask else None
    status = torch.zeros(B, dtype=torch.int32, device=wave_padded.device)
    _cabi.check(_cabi.lib().aat_scatter_segments(
        batch.ctx.handle, wave_padded.data_ptr(), int(wave_padded.shape[1]), B, boarders_padded.data_ptr(), s_max,
        int(max_segment_frames), out.data_ptr(), mask.data_ptr() if with_mask else None, status.data_ptr(), _stream(wave_padded.device)))
    if check:
        _raise_on_status(status, "scatter_segments")
    return (out, mask) if with_mask else out


def scatter_mel_segments(batch: PackedBatch, boarders_padded, max_segment_frames: int, mel=None, check: bool = True):
    """``batched_segments_melspectrograms [B, S, n_mels, 1 + max_segment_frames // hop]`` from the batch's packed
    log-mel (ref:src/aat/training/collate.py:309-312, 337-342)."""
    import torch

    mel = batch.mel if mel is None else mel
    hop = int(batch.tokenizer.hop_length)
    max_items = 1 + int(max_segment_frames) // hop
    B, s_max = boarders_padded.shape
    out = torch.empty((B, s_max, batch.n_mels, max_items), dtype=torch.float32, device=mel.device)
    status = torch.zeros(B, dtype=torch.int32, device=mel.device)
    _cabi.check(_cabi.lib().aat_scatter_mel_segments(batch.ctx.handle, batch.handle, mel.data_ptr(),
                                                     boarders_padded.data_ptr(), s_max, max_items, out.data_ptr(),
                                                     status.data_ptr(), _stream(mel.device)))
    if check:
        _raise_on_status(status, "scatter_mel_segments")
    return out


def masked_mean_pool(audio_embeds, audio_embeds_attention_mask):
    """Mean over the valid frames of ``audio_embeds [R, L, D]`` under ``audio_embeds_attention_mask [R, L]`` —
    the ``SegmentProjectionEnum.mean`` pooling that ``AslmModel.audio_embeddings_projection`` leaves as
    ``NotImplementedError`` (ref:src/aslm/modeling_aslm.py:258-259).  Returns ``(pooled [R, D] float32,
    row_mask [R] int64)``; rows without a valid frame are zero and masked out, like the CLS branch (:249-254)."""
    import torch

    from .context import default_context
    from .pooling import _torch_dtype_code

    if audio_embeds.dim() != 3 or not audio_embeds.is_cuda:
        raise TypeError("audio_embeds must be a CUDA tensor [R, L, D]")
    emb = audio_embeds.contiguous()
    R, L, D = emb.shape
    mask = audio_embeds_attention_mask.to(device=emb.device, dtype=torch.int64).contiguous()
    if mask.shape != (R, L):
        raise ValueError("mask must have shape [R, L]")
    out = torch.empty((R, D), dtype=torch.float32, device=emb.device)
    row_mask = torch.empty(R, dtype=torch.int64, device=emb.device)
    ctx = default_context(emb.device.index)
    _cabi.check(_cabi.lib().aat_masked_mean_pool(ctx.handle, emb.data_ptr(), _torch_dtype_code(emb.dtype), R, L, D,
                                                 mask.data_ptr(), out.data_ptr(), row_mask.data_ptr(),
                                                 _stream(emb.device)))
    return out, row_mask


def uniform_segment_lengths(n_samples: int, frames_per_segment: int):
    """Uniform segmentation of the collator (ref:src/aat/training/collate.py:141-149): equal segments plus the
    remainder; returns the lengths (host integers — there is no arithmetic to accelerate)."""
    import numpy as np

    num = n_samples // frames_per_segment
    lengths = [frames_per_segment] * num
    if n_samples % frames_per_segment > 0:
        lengths.append(n_samples - sum(lengths))
    return np.asarray(lengths, dtype=np.int64)


def crop_to_words(frames_boarders_raw, word_start, word_end, word_start_idx: int, n_words: int, sampling_rate: int,
                  hop_length: int, running_mean_points: int, n_samples: int, n_mel_frames: int,
                  melspec_overlapping: int = 5):
    """The collator's n-word cropping (ref:src/aat/training/collate.py:169-212) as a pure function of the segment
    lengths: which segments, waveform samples and mel frames survive when ``n_words`` consecutive words starting at
    ``word_start_idx`` (the reference draws it with ``random.randint(0, len(words) - n_words)``, :176) are kept.

    frames_boarders_raw : segment lengths in samples (their sum must be ``n_samples``, the reference asserts it, :172)
    word_start, word_end: per-word times in seconds (``item['word_start']`` / ``item['word_end']``)

    Returns ``(frames_boarders, (wave_lo, wave_hi), (mel_lo, mel_hi), (word_lo, word_hi))``: the boarders of the kept
    segments rebased to the first kept segment (leading zero cut off, :197-199), the half-open waveform slice widened
    by ``melspec_overlapping`` hops each side (:203-206), the mel slice widened by the running-mean window on the left
    (:208-211), and the word range.  Host integers only — there is nothing here for a GPU to accelerate; the outputs
    are what ``scatter_segments`` / ``scatter_mel_segments`` are then fed with."""
    import numpy as np

    raw = np.asarray(frames_boarders_raw, dtype=np.int64)
    if int(raw.sum()) != int(n_samples):
        raise AssertionError("segment lengths must add up to the waveform length (ref:src/aat/training/collate.py:172)")
    boarders = raw.cumsum()
    word_end_idx = word_start_idx + n_words
    start_frame = int(word_start[word_start_idx] * sampling_rate)
    end_frame = int(word_end[word_end_idx - 1] * sampling_rate)
    with_zero = np.insert(boarders, 0, [0])
    first = max(int(np.searchsorted(with_zero, start_frame)) - 1, 0)
    last = int(np.searchsorted(with_zero, end_frame, side="right"))
    if not last < len(with_zero):
        raise AssertionError("the last word ends after the last segment (ref:src/aat/training/collate.py:186)")
    seg_lo, seg_hi = int(with_zero[first]), int(with_zero[last])
    if not (seg_lo <= start_frame and seg_hi >= end_frame):
        raise AssertionError("the kept segments do not cover the kept words (ref:src/aat/training/collate.py:189-192)")
    kept = with_zero[first: last + 1] - seg_lo
    overlap = melspec_overlapping * hop_length
    wave_lo = max(0, seg_lo - overlap)
    wave_hi = min(seg_hi + overlap, int(n_samples))
    mel_lo = max(0, wave_lo // hop_length - running_mean_points - melspec_overlapping)
    mel_hi = min(wave_hi // hop_length + melspec_overlapping, int(n_mel_frames))
    return kept[1:], (wave_lo, wave_hi), (mel_lo, mel_hi), (word_start_idx, word_end_idx)
