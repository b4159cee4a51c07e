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


def normalize_waveforms_padded(batch: PackedBatch, wave, mode: str = "w2v2", n_max=None, with_mask: bool = True):
    """Per-utterance normalisation written straight into the feature extractor's padded layout
    (``audio_processor(waveforms, padding=True, return_tensors="pt")``, ref:src/aat/training/collate.py:301-304):
    returns ``(input_values [B, N_max] float32, attention_mask [B, N_max] int32 or None)`` (the dtypes the feature
    extractor returns)."""
    import torch

    modes = {"zscore": _cabi.AAT_NORM_ZSCORE, "w2v2": _cabi.AAT_NORM_W2V2}
    codes = {torch.float32: _cabi.AAT_F32, torch.float64: _cabi.AAT_F64}
    if wave.dtype not in codes or not wave.is_cuda or not wave.is_contiguous() or wave.numel() != batch.total_samples:
        raise TypeError("wave must be a contiguous packed CUDA tensor (float32 or float64) matching the plan")
    if n_max is None:
        n_max = int(batch.n_samples.max()) if batch.n_utts else 0
    out = torch.empty((batch.n_utts, n_max), dtype=torch.float32, device=wave.device)
    mask = torch.empty((batch.n_utts, n_max), dtype=torch.int32, device=wave.device) if with_mask else None
    _cabi.check(_cabi.lib().aat_normalize_padded(batch.ctx.handle, batch.handle, wave.data_ptr(), codes[wave.dtype],
                                                 modes[mode], out.data_ptr(), int(n_max),
                                                 mask.data_ptr() if with_mask else None, None, _stream(wave.device)))
    return out, mask


def scatter_mel_tiles(batch: PackedBatch, mel, mel_elem_off, mel_frames, mel_row_stride, boarders_padded,
                      max_segment_frames: int, check: bool = True):
    """:func:`scatter_mel_segments` for mel blocks described by three int64 device arrays (first element, columns, row
    stride per utterance) — column slices of the packed log-mel, as the n-word cropping leaves them."""
    import torch

    hop = int(batch.tokenizer.hop_length)
    max_items = 1 + int(max_segment_frames) // hop
    B, s_max = boarders_padded.shape
    out = torch.empty((B, s_max, batch.n_mels, max_items), dtype=torch.float32, device=mel.device)
    status = torch.zeros(B, dtype=torch.int32, device=mel.device)
    _cabi.check(_cabi.lib().aat_scatter_mel_tiles(batch.ctx.handle, B, mel.data_ptr(), mel_elem_off.data_ptr(),
                                                  mel_frames.data_ptr(), mel_row_stride.data_ptr(),
                                                  boarders_padded.data_ptr(), s_max, max_items, out.data_ptr(),
                                                  status.data_ptr(), _stream(mel.device)))
    if check:
        _raise_on_status(status, "scatter_mel_tiles")
    return out


def collate_batch(tokenizer, waveforms, *, audio_encoder_type: str = "hubert", segmentation: str = "adaptive",
                  uniform_segmentation_frames_per_segment=None, word_crops=None, device=None):
    """The audio tensors of ``TokenizedAudioWaveformCollator.__call__`` (ref:src/aat/training/collate.py:255-352),
    built on the GPU from raw waveforms: z-score (:135-152) -> log-mel -> adaptive (or uniform, :141-149) segmentation
    -> optional n-word cropping (:169-212) -> the feature extractor's normalisation and padding (:301-304) -> padded
    boarders (:242-253) -> waveform tiles + mask (:321-335) or log-mel tiles (:337-342).  Text columns are not produced.

    waveforms  : list of 1-D float arrays (float64, as HF ``datasets`` delivers ``item['audio']['array']``)
    audio_encoder_type : "hubert" / "wav2vec2" (waveform tiles) or "efficient_net" (log-mel tiles)
    segmentation : "adaptive" or "uniform" (then ``uniform_segmentation_frames_per_segment`` is required)
    word_crops : optional list with, per item, None or ``dict(word_start=[...], word_end=[...], word_start_idx=int,
                 n_words=int)`` — the draws the reference takes from ``random`` (:122, :176) are the caller's

    Returns a dict with the reference's keys (``segments_boarders_padded``, ``segments_boarders_attention_mask``,
    ``segments_max_frame_len``, ``batched_segments``, ``segments_waveforms_mask``, ``batched_segments_melspectrograms``,
    ``segments_count``) as CUDA tensors, plus ``audio_input_values`` / ``audio_attention_mask`` (the processor's output,
    None for efficient_net) and ``batch`` (the :class:`PackedBatch` holding the packed log-mel).  Raises where the
    reference raises (a segment longer than the tile, a slice past the padded waveform).  Everything but the integer
    cropping arithmetic and the (tiny) boarder tables runs in the library's kernels."""
    import numpy as np
    import torch

    if segmentation not in ("adaptive", "uniform"):
        raise ValueError(f"Unhandled segmentation type: {segmentation}")
    waves = [np.asarray(w) for w in waveforms]
    if any(w.ndim != 1 for w in waves):
        raise AssertionError("channel dim is not supported for waveform")
    B = len(waves)
    lengths = [int(w.shape[0]) for w in waves]
    batch = tokenizer.plan(lengths, device=device)
    dev = batch.device
    raw = batch.pack([torch.from_numpy(np.ascontiguousarray(w, dtype=np.float64)) for w in waves])
    hop = int(tokenizer.hop_length)
    max_frames = int(tokenizer.max_segment_frames)

    # z-score fused into the log-mel kernel's staging (what the collator caches as `melspec`), boundaries
    batch.logmel(raw, znorm_stats=batch.waveform_stats(raw))
    if segmentation == "adaptive":
        batch.boundaries()
        status = batch.status.cpu().numpy()
        if (status < 0).any():
            bad = int(np.argmax(status < 0))
            raise _cabi.AatError(int(status[bad]), f"boundary kernel reported an error for utterance {bad}")
        counts = batch.seg_count.cpu().numpy()
        seg_len = batch.seg_len.cpu().numpy()
        raw_lengths = [seg_len[int(batch.seg_slot_off[b]): int(batch.seg_slot_off[b]) + int(counts[b])] for b in range(B)]
    else:
        if not uniform_segmentation_frames_per_segment:
            raise ValueError("uniform segmentation needs uniform_segmentation_frames_per_segment")
        raw_lengths = [uniform_segment_lengths(n, int(uniform_segmentation_frames_per_segment)) for n in lengths]

    # n-word cropping: integer arithmetic on the host, slices taken on the device
    boarders, wave_rng, mel_rng = [], [], []
    for b in range(B):
        crop = word_crops[b] if word_crops is not None else None
        n_mel = 1 + lengths[b] // hop
        if crop is None:
            boarders.append(np.cumsum(raw_lengths[b]))
            wave_rng.append((0, lengths[b]))
            mel_rng.append((0, n_mel))
        else:
            kept, wr, mr, _ = crop_to_words(raw_lengths[b], crop["word_start"], crop["word_end"], crop["word_start_idx"],
                                            crop["n_words"], int(tokenizer.sampling_rate), hop,
                                            int(tokenizer.running_mean_points), lengths[b], n_mel)
            boarders.append(np.asarray(kept, dtype=np.int64))
            wave_rng.append(wr)
            mel_rng.append(mr)
    s_max = max(len(x) for x in boarders)
    padded_h = np.zeros((B, s_max), dtype=np.int64)
    mask_h = np.zeros((B, s_max), dtype=np.int64)
    for b, sb in enumerate(boarders):
        padded_h[b, : len(sb)] = sb
        mask_h[b, : len(sb)] = 1
    padded = torch.from_numpy(padded_h).to(dev)
    result = {
        "segments_boarders_padded": padded,
        "segments_boarders_attention_mask": torch.from_numpy(mask_h).to(dev),
        "segments_max_frame_len": torch.tensor([int(np.max(x)) for x in raw_lengths], device=dev),
        "segments_count": s_max, "batch": batch,
        "batched_segments": None, "segments_waveforms_mask": None, "batched_segments_melspectrograms": None,
        "audio_input_values": None, "audio_attention_mask": None,
    }

    if audio_encoder_type != "efficient_net":
        cropped = any(r != (0, n) for r, n in zip(wave_rng, lengths))
        if cropped:
            crop_batch = tokenizer.plan([hi - lo for lo, hi in wave_rng], device=device)
            crop_raw = torch.cat([raw[int(batch.wave_off[b]) + lo: int(batch.wave_off[b]) + hi]
                                  for b, (lo, hi) in enumerate(wave_rng)])
        else:
            crop_batch, crop_raw = batch, raw
        values, attention = normalize_waveforms_padded(crop_batch, crop_raw, "w2v2")
        assert values.shape[1] > 0
        segs, seg_mask = scatter_segments(crop_batch, values, padded, max_frames)
        result.update(audio_input_values=values, audio_attention_mask=attention, batched_segments=segs,
                      segments_waveforms_mask=seg_mask)
    else:
        n_mels = batch.n_mels
        off = [n_mels * int(batch.frame_off[b]) + lo for b, (lo, hi) in enumerate(mel_rng)]
        frames = [hi - lo for lo, hi in mel_rng]
        stride = [1 + lengths[b] // hop for b in range(B)]
        as_dev = lambda v: torch.tensor(v, dtype=torch.int64, device=dev)  # noqa: E731
        result["batched_segments_melspectrograms"] = scatter_mel_tiles(batch, batch.mel, as_dev(off), as_dev(frames),
                                                                       as_dev(stride), padded, max_frames)
    return result


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
