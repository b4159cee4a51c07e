"""CPU checks of the collator port (oracle for SURVEY §8f rows N1/N2): the restated loops agree with an
independent vectorised formulation, and the GPU-facing ABI declares the matching entry points."""
import numpy as np
import torch

from oracle import collate_port


def test_padded_boarders_and_scatter_against_vectorised_numpy():
    rng = np.random.default_rng(0)
    B, max_frames, hop, n_mels = 5, 24000, 160, 64
    seg_lengths = [rng.integers(2000, 24000, size=int(rng.integers(1, 9))) for _ in range(B)]
    boarders = [np.cumsum(l) for l in seg_lengths]
    padded, mask = collate_port.make_padded_segments_boarders(boarders, B)
    s_max = max(len(b) for b in boarders)
    assert padded.shape == (B, s_max) and padded.dtype == torch.long
    for b in range(B):
        assert padded[b, : len(boarders[b])].tolist() == boarders[b].tolist()
        assert padded[b, len(boarders[b]):].sum() == 0 and mask[b].sum() == len(boarders[b])
    n_max = max(int(b[-1]) for b in boarders)
    wave = torch.from_numpy(rng.standard_normal((B, n_max)).astype(np.float32))
    mels = [rng.standard_normal((n_mels, 1 + n_max // hop)).astype(np.float32) for _ in range(B)]
    seg, segmask, tiles = collate_port.scatter_segments(wave, padded, max_frames, items_melspecs=mels)
    assert seg.shape == (B, s_max, max_frames) and tiles.shape == (B, s_max, n_mels, 1 + max_frames // hop)
    for b in range(B):
        starts = np.concatenate([[0], boarders[b][:-1]])
        for s, (a, e) in enumerate(zip(starts, boarders[b])):
            assert torch.equal(seg[b, s, : e - a], wave[b, a:e]) and seg[b, s, e - a:].abs().sum() == 0
            assert segmask[b, s].sum() == e - a
            cols = e // hop - a // hop
            assert np.array_equal(tiles[b, s, :, :cols].numpy(), mels[b][:, a // hop: e // hop])
        assert seg[b, len(boarders[b]):].abs().sum() == 0
    # a segment longer than the tile makes the reference (and the port) fail
    try:
        collate_port.scatter_segments(wave, padded, 1000)
        raise AssertionError("expected a shape mismatch")
    except RuntimeError:
        pass


def test_normalisation_ports():
    rng = np.random.default_rng(1)
    x = rng.standard_normal(50000) * 0.3 + 2.0
    z = collate_port.znorm(x)
    assert z.dtype == np.float64 and abs(z.mean()) < 1e-12 and abs(z.std() - 1.0) < 1e-5
    w = collate_port.w2v2_norm(x)
    assert w.dtype == np.float32 and abs(float(w.mean())) < 1e-4 and abs(float(w.std()) - 1.0) < 1e-3


def test_crop_to_words_matches_the_collator_port():
    """the n-word cropping of ref:src/aat/training/collate.py:169-212 (host index arithmetic)."""
    import os
    import sys

    sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "audio-adaptive-tokenizer_b200"))
    from aat_b200 import collate

    rng = np.random.default_rng(5)
    sr, hop, npts = 16000, 160, 12
    checked = 0
    for trial in range(200):
        lengths = rng.integers(2000, 24001, size=int(rng.integers(3, 40)))
        n = int(lengths.sum())
        T = 1 + n // hop
        wave = rng.standard_normal(n).astype(np.float32)
        mel = rng.standard_normal((8, T)).astype(np.float32)
        n_all = int(rng.integers(4, 30))
        cuts = np.sort(rng.uniform(0.0, n / sr, size=2 * n_all))
        word_start, word_end = cuts[0::2], cuts[1::2]
        words = [f"w{i}" for i in range(n_all)]
        n_words = int(rng.integers(1, n_all))
        idx = int(rng.integers(0, n_all - n_words + 1))
        try:
            want = collate_port.crop_to_words(wave, mel, lengths, words, word_start, word_end, idx, n_words, sr, hop, npts)
        except AssertionError:
            with np.testing.assert_raises(AssertionError):
                collate.crop_to_words(lengths, word_start, word_end, idx, n_words, sr, hop, npts, n, T)
            continue
        boarders, (w0, w1), (m0, m1), (a, b) = collate.crop_to_words(lengths, word_start, word_end, idx, n_words, sr, hop,
                                                                    npts, n, T)
        assert words[a:b] == want[0]
        assert np.array_equal(wave[w0:w1], want[1]) and np.array_equal(mel[:, m0:m1], want[2])
        assert np.array_equal(boarders, want[3])
        checked += 1
    assert checked > 150
    with np.testing.assert_raises(AssertionError):
        collate.crop_to_words([1000, 2000], [0.0], [0.1], 0, 1, sr, hop, npts, 2999, 19)
