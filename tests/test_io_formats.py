"""Wire compatibility with the reference's on-disk formats.  CPU only, except the directory
pooling job, which needs the GPU."""
import os

import numpy as np
import pytest
import torch

from aat_b200 import io as aio


def test_mel_cache_round_trip_matches_reference_idiom(tmp_path):
    mel = np.random.default_rng(0).standard_normal((64, 101)).astype(np.float32)
    # what ref:scripts/audio_tokenization_melspec.py:42 writes ...
    ref_path = str(tmp_path / "ref_item")
    torch.save(mel, ref_path)
    got = aio.load_melspec(ref_path)
    assert isinstance(got, np.ndarray) and got.dtype == np.float32 and np.array_equal(got, mel)
    # ... and what ref:src/aat/training/collate.py:133 reads back
    ours = str(tmp_path / "our_item")
    aio.save_melspec(ours, mel)
    assert np.array_equal(torch.load(ours, weights_only=False), mel)


def test_segment_embedding_lists_pack_to_csr(tmp_path):
    g = torch.Generator().manual_seed(0)
    lst = [torch.randn(1, n, 768, generator=g) for n in (6, 74, 13, 1)]
    path = str(tmp_path / "emb")
    torch.save(lst, path)
    packed, off = aio.load_segment_embeddings(path)
    assert packed.shape == (94, 768) and off.tolist() == [0, 6, 80, 93, 94]
    assert torch.equal(packed[6:80], lst[1][0])
    pooled = torch.cat([x.mean(dim=1, keepdim=True).to(torch.float32) for x in lst], dim=1)
    out = str(tmp_path / "pooled")
    aio.save_pooled(out, pooled[0])
    back = torch.load(out, weights_only=True)
    assert back.shape == (1, 4, 768) and back.dtype == torch.float32 and torch.equal(back, pooled)


def test_segment_frames_column_and_shards(tmp_path):
    from aat_b200 import AudioWaveform

    segs = [AudioWaveform(np.zeros(n), 16000) for n in (24000, 8000)]
    assert aio.segment_frames_column(segs) == [24000, 8000]
    assert aio.segment_frames_column(np.array([5, 7])) == [5, 7]
    w = aio.MelShardWriter()
    rng = np.random.default_rng(1)
    mels = {f"id{i}": rng.standard_normal((64, 50 + i)).astype(np.float32) for i in range(5)}
    for i, (k, m) in enumerate(mels.items()):
        w.add(k, m, [2000 + i, 3000])
    path = str(tmp_path / "shard.pt")
    w.write(path)
    r = aio.MelShardReader(path)
    assert len(r) == 5
    for i, (k, m) in enumerate(mels.items()):
        assert np.array_equal(r.melspec(k), m) and r.segment_frames(k).tolist() == [2000 + i, 3000]


@pytest.mark.gpu
def test_pool_embedding_directory_matches_reference_script(tmp_path):
    """ref:scripts/mean_hubert_embeddings.py:16-23 on a small directory; several files share one launch."""
    src, dst = tmp_path / "audio_segments_embeddings", tmp_path / "audio_segments_embeddings_mean"
    os.makedirs(src)
    g = torch.Generator().manual_seed(5)
    want = {}
    for i in range(7):
        lst = [torch.randn(1, int(n), 768, generator=g) for n in torch.randint(6, 75, (3 + i,), generator=g)]
        torch.save(lst, str(src / f"file_{i}.pt"))
        want[f"file_{i}.pt"] = torch.cat([x.mean(dim=1, keepdim=True).to(torch.float32) for x in lst], dim=1)
    assert aio.pool_embedding_files(str(src), str(dst), files_per_launch=3) == 7
    for name, ref in want.items():
        got = torch.load(str(dst / name), weights_only=True)
        assert got.shape == ref.shape and got.dtype == torch.float32
        rel = (got - ref).norm(dim=-1) / ref.norm(dim=-1)
        assert float(rel.max()) <= 1e-5
    assert aio.pool_embedding_files(str(src), str(dst)) == 0  # everything is already there
