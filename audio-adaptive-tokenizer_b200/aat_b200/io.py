"""On-disk formats of the reference around the hot path — wire compatible readers and
writers, plus packed shards so the GPU path is not starved by one-file-per-utterance I/O.

Formats (none is documented by the reference; they are what its scripts write and read):
  mel cache           ``data/libris_melspectrograms/<id>``  = ``torch.save(np.ndarray (n_mels, T) float32)``
                      written at ref:scripts/audio_tokenization_melspec.py:42, read at ref:src/aat/training/collate.py:133
  segment embeddings  ``data/audio_segments_embeddings/<file>`` = ``torch.save(list of tensors [1, n_i, D])``
                      read at ref:scripts/mean_hubert_embeddings.py:18 and ref:src/aat/datasets/hubert_libris.py:15
  pooled embeddings   ``data/audio_segments_embeddings_mean/<file>`` = ``torch.save(tensor [1, S, D] float32)``
                      written at ref:scripts/mean_hubert_embeddings.py:20-23
  segment_frames      HF-dataset column: list of int segment lengths per item (ref:scripts/audio_tokenization.py:37-38)
"""
from __future__ import annotations

import os
from typing import Iterable, List, Sequence, Tuple

import numpy as np


# ------------------------------------------------------------------------------------------------ mel cache
def save_melspec(path: str, melspec: np.ndarray) -> None:
    import torch

    torch.save(np.asarray(melspec, dtype=np.float32), path)


def load_melspec(path: str) -> np.ndarray:
    import torch

    return torch.load(path, weights_only=False)


# ------------------------------------------------------------------------------------------------ embeddings
def pack_segment_embeddings(embeddings_list) -> Tuple["torch.Tensor", np.ndarray]:
    """List of ``[1, n_i, D]`` tensors -> packed ``[T, D]`` tensor + CSR offsets ``[S+1]`` (int64)."""
    import torch

    parts = list(embeddings_list)
    lengths = [int(x.shape[1]) for x in parts]
    off = np.zeros(len(parts) + 1, dtype=np.int64)
    np.cumsum(lengths, out=off[1:])
    if not parts:
        raise ValueError("empty embedding list")
    packed = torch.cat([x.reshape(x.shape[1], x.shape[2]) for x in parts], dim=0)
    return packed, off


def load_segment_embeddings(path: str):
    """``torch.load`` of the reference's per-file list, packed for the pool kernel."""
    import torch

    return pack_segment_embeddings(torch.load(path, map_location="cpu", weights_only=True))


def save_pooled(path: str, pooled) -> None:
    import torch

    t = pooled if isinstance(pooled, torch.Tensor) else torch.as_tensor(pooled)
    if t.dim() == 2:
        t = t.unsqueeze(0)
    torch.save(t.to(torch.float32).cpu(), path)


def pool_embedding_files(source_dir: str, target_dir: str, device="cuda", files_per_launch: int = 64,
                         skip_existing: bool = True) -> int:
    """GPU version of ref:scripts/mean_hubert_embeddings.py:7-23: every file of per-segment embeddings in
    ``source_dir`` becomes a ``[1, S, D]`` float32 file of per-segment means in ``target_dir``.

    Many files are packed into ONE pool launch (their segments are simply concatenated in the CSR), so the
    kernel streams megabytes per launch instead of a few kilobytes per file.  Unlike the reference, which
    deletes ``target_dir`` first, finished files are skipped (the resume behaviour of
    ref:scripts/audio_tokenization_melspec.py:32,36-37).  Returns the number of files written."""
    import torch

    from .pooling import mean_pool_segments

    os.makedirs(target_dir, exist_ok=True)
    names = sorted(os.listdir(source_dir))
    if skip_existing:
        done = set(os.listdir(target_dir))
        names = [n for n in names if n not in done]
    written = 0
    for i0 in range(0, len(names), files_per_launch):
        group = names[i0:i0 + files_per_launch]
        packed, offs = [], []
        for name in group:
            e, o = load_segment_embeddings(os.path.join(source_dir, name))
            packed.append(e)
            offs.append(o)
        dtype = packed[0].dtype
        if any(p.dtype != dtype or p.shape[1] != packed[0].shape[1] for p in packed):
            raise ValueError("files of one launch must share dtype and embedding width")
        rows = np.cumsum([0] + [int(p.shape[0]) for p in packed])
        seg_off = np.concatenate([offs[0]] + [o[1:] + rows[k] for k, o in enumerate(offs) if k > 0])
        emb = torch.cat(packed, dim=0).to(device, non_blocking=True)
        pooled = mean_pool_segments(emb, seg_off)[0].cpu()
        seg_counts = np.cumsum([0] + [o.size - 1 for o in offs])
        for k, name in enumerate(group):
            save_pooled(os.path.join(target_dir, name), pooled[seg_counts[k]:seg_counts[k + 1]].unsqueeze(0).clone())
            written += 1
    return written


def segment_frames_column(segments) -> List[int]:
    """The ``segment_frames`` dataset column (ref:scripts/audio_tokenization.py:37-38) from either the list of
    ``AudioWaveform`` segments ``tokenize`` returns or an array of lengths."""
    out = []
    for s in segments:
        out.append(int(s.waveform.shape[-1]) if hasattr(s, "waveform") else int(s))
    return out


# ------------------------------------------------------------------------------------------------ packed shards
class MelShardWriter:
    """Accumulates many utterances' log-mels (and optionally their segment lengths) and writes ONE file:
    ``{"ids", "frame_off" [B+1], "mel" [n_mels, sum T], "seg_off" [B+1], "segment_frames" [sum S]}``.
    ``mel[:, frame_off[b]:frame_off[b+1]]`` is utterance b's ``(n_mels, T_b)`` spectrogram."""

    def __init__(self, n_mels: int = 64):
        self.n_mels = n_mels
        self.ids: List[str] = []
        self.mels: List[np.ndarray] = []
        self.segs: List[np.ndarray] = []

    def add(self, item_id: str, melspec: np.ndarray, segment_frames: Sequence[int] = ()):
        mel = np.asarray(melspec, dtype=np.float32)
        if mel.ndim != 2 or mel.shape[0] != self.n_mels:
            raise ValueError(f"melspec must have shape ({self.n_mels}, T)")
        self.ids.append(str(item_id))
        self.mels.append(mel)
        self.segs.append(np.asarray(list(segment_frames), dtype=np.int64))

    def write(self, path: str) -> None:
        import torch

        frame_off = np.cumsum([0] + [m.shape[1] for m in self.mels]).astype(np.int64)
        seg_off = np.cumsum([0] + [s.size for s in self.segs]).astype(np.int64)
        torch.save({"ids": list(self.ids), "frame_off": torch.from_numpy(frame_off),
                    "mel": torch.from_numpy(np.concatenate(self.mels, axis=1) if self.mels else np.zeros((self.n_mels, 0), np.float32)),
                    "seg_off": torch.from_numpy(seg_off),
                    "segment_frames": torch.from_numpy(np.concatenate(self.segs) if self.segs else np.zeros(0, np.int64))},
                   path)


class MelShardReader:
    def __init__(self, path: str):
        import torch

        d = torch.load(path, map_location="cpu", weights_only=True)
        self.ids = list(d["ids"])
        self.frame_off = d["frame_off"].numpy()
        self.seg_off = d["seg_off"].numpy()
        self.mel = d["mel"].numpy()
        self.segment_frames_all = d["segment_frames"].numpy()
        self._index = {k: i for i, k in enumerate(self.ids)}

    def __len__(self):
        return len(self.ids)

    def melspec(self, item_id: str) -> np.ndarray:
        b = self._index[str(item_id)]
        return self.mel[:, self.frame_off[b]:self.frame_off[b + 1]]

    def segment_frames(self, item_id: str) -> np.ndarray:
        b = self._index[str(item_id)]
        return self.segment_frames_all[self.seg_off[b]:self.seg_off[b + 1]]
