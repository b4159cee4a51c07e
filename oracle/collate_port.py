"""ORACLE (test infrastructure, not product code) — port of the collator's ragged -> padded layout and of
the waveform normalisations either side of the hot path.

Pinned (round 2): tests/golden/make_collate_golden.py imports the UNMODIFIED `aat.training.collate` from
/root/reference/src (placeholder modules for the absent `efficientnet_pytorch` and for
`transformers.trainer.ALL_LAYERNORM_LAYERS`, a locally built Wav2Vec2FeatureExtractor instead of the hub download) and
commits what `TokenizedAudioWaveformCollator.__call__` / `_initial_process_segments` return for six batches
(tests/golden/collate_v1.npz).  tests/test_collate_golden.py checks the loops below — a statement-by-statement
restatement of the cited lines — against those fixtures bit for bit, and the CUDA path against the same fixtures.

  ref:src/aat/training/collate.py:242-253   _make_padded_segments_boarders
  ref:src/aat/training/collate.py:309-346   batched_segments / segments_waveforms_mask / melspectrogram tiles
  ref:src/aat/training/collate.py:135,138,152, ref:scripts/audio_tokenization_melspec.py:40   z-score
  TF:models/wav2vec2/feature_extraction_wav2vec2.py:78-95   zero_mean_unit_var_norm (no attention mask branch)

Nothing under ``audio-adaptive-tokenizer_b200/`` may import this module.
"""
from __future__ import annotations

import numpy as np
import torch


def make_padded_segments_boarders(segments_boarders, batch_size):
    max_len = max(len(x) for x in segments_boarders)
    padded = torch.zeros([batch_size, max_len], dtype=torch.long)
    mask = torch.zeros_like(padded)
    for i, sb in enumerate(segments_boarders):
        padded[i, : len(sb)] = torch.tensor(sb, dtype=torch.long)
        mask[i, : len(sb)] = 1
    return padded, mask


def scatter_segments(audio_input_values, segments_boarders_padded, max_segment_waveform_frames, items_melspecs=None,
                     hop_length=160, num_mel_filters=64):
    """Returns (batched_segments, segments_waveforms_mask, batched_segments_melspectrograms)."""
    batch_size, segments_count = segments_boarders_padded.shape
    max_melspec_items = int(1 + np.floor(max_segment_waveform_frames / hop_length))
    mel_tiles = None
    if items_melspecs is not None:
        mel_tiles = torch.zeros([batch_size, segments_count, num_mel_filters, max_melspec_items])
    batched_segments = torch.zeros([batch_size, segments_count, max_segment_waveform_frames])
    segments_waveforms_mask = torch.zeros([batch_size, segments_count, max_segment_waveform_frames])
    for batch_i in range(batch_size):
        prev = 0
        for segment_i in range(segments_count):
            boarder = int(segments_boarders_padded[batch_i, segment_i])
            if segment_i > 0 and boarder == 0:
                continue
            assert prev < boarder
            length = boarder - prev
            current = audio_input_values[batch_i, prev:boarder]
            batched_segments[batch_i, segment_i, :length] = current
            segments_waveforms_mask[batch_i, segment_i, :length] = 1
            if mel_tiles is not None:
                a, b = prev // hop_length, boarder // hop_length
                seg = items_melspecs[batch_i][:, a:b]
                mel_tiles[batch_i, segment_i, :, : seg.shape[1]] = torch.from_numpy(np.ascontiguousarray(seg))
            prev = boarder
    return batched_segments, segments_waveforms_mask, mel_tiles


def znorm(x: np.ndarray) -> np.ndarray:
    return (x - x.mean()) / (x.std() + 1e-6)


def w2v2_norm(x: np.ndarray) -> np.ndarray:
    x = np.asarray(x, dtype=np.float32)
    return (x - x.mean()) / np.sqrt(x.var() + 1e-7)


def masked_mean_pool(audio_embeds, audio_embeds_attention_mask):
    """Definition used for row N4 (the reference raises NotImplementedError at ref:src/aslm/modeling_aslm.py:258-259):
    masked mean over the frame axis, zeros where no frame is valid; row mask as in the CLS branch (:249-254)."""
    m = audio_embeds_attention_mask.to(torch.float64).unsqueeze(-1)
    s = (audio_embeds.to(torch.float64) * m).sum(dim=1)
    n = m.sum(dim=1)
    out = torch.where(n > 0, s / n.clamp(min=1), torch.zeros_like(s))
    return out, (audio_embeds_attention_mask != 0).any(dim=-1).long()


def crop_to_words(waveform, melspec, frames_boarders_raw, words, word_start, word_end, word_start_idx, n_words, sampling_rate,
                  hop_length, running_mean_points):
    """ref:src/aat/training/collate.py:169-212, statement by statement, with the random draw (:176) passed in as
    ``word_start_idx``.  Returns (words, waveform, melspec, frames_boarders) as the collator leaves them."""
    frames_boarders_raw = np.asarray(frames_boarders_raw)
    frames_boarders = frames_boarders_raw.cumsum()
    waveform_num_frames = waveform.shape[-1]
    assert frames_boarders_raw.sum() == waveform_num_frames
    waveform_end_frame = waveform.shape[-1]
    word_end_idx = word_start_idx + n_words
    words = words[word_start_idx:word_end_idx]
    waveform_start_frame = int(word_start[word_start_idx] * sampling_rate)
    waveform_end_frame = int(word_end[word_end_idx - 1] * sampling_rate)
    frames_boarders_with_zero = np.insert(frames_boarders, 0, [0])
    waveform_start_segment_idx = np.searchsorted(frames_boarders_with_zero, waveform_start_frame)
    waveform_start_segment_idx -= 1
    waveform_start_segment_idx = max(waveform_start_segment_idx, 0)
    assert waveform_start_segment_idx >= 0
    waveform_end_segment_idx = np.searchsorted(frames_boarders_with_zero, waveform_end_frame, side='right')
    assert waveform_end_segment_idx < len(frames_boarders_with_zero)
    start_segment_waveform_num = frames_boarders_with_zero[waveform_start_segment_idx]
    assert start_segment_waveform_num <= waveform_start_frame
    end_segment_waveform_num = frames_boarders_with_zero[waveform_end_segment_idx]
    assert end_segment_waveform_num >= waveform_end_frame
    frames_boarders = frames_boarders_with_zero[waveform_start_segment_idx:(waveform_end_segment_idx + 1)]
    frames_boarders = frames_boarders - start_segment_waveform_num
    assert frames_boarders[0] == 0
    frames_boarders = frames_boarders[1:]
    melspec_overlapping = 5
    waveform_frames_overlapping = melspec_overlapping * hop_length
    start_segment_waveform_num = max(0, start_segment_waveform_num - waveform_frames_overlapping)
    end_segment_waveform_num = min(end_segment_waveform_num + waveform_frames_overlapping, waveform.shape[-1])
    waveform = waveform[start_segment_waveform_num:end_segment_waveform_num]
    start_segment_melspec, end_segment_melspec = start_segment_waveform_num // hop_length, end_segment_waveform_num // hop_length
    start_segment_melspec = max(0, start_segment_melspec - running_mean_points - melspec_overlapping)
    end_segment_melspec = min(end_segment_melspec + melspec_overlapping, melspec.shape[-1])
    melspec = melspec[:, start_segment_melspec:end_segment_melspec]
    return words, waveform, melspec, frames_boarders
