"""Frame CSR of the whole-utterance encode convention (synthetic embeddings, convention (ii)):
``aat_utterance_frame_csr`` against the oracle's loop, and the pool kernel on that CSR against torch.

The reference has no code for this convention (its scripts pool per-segment encodings,
ref:scripts/mean_hubert_embeddings.py:18-20); the oracle loop in ``oracle/restate.py`` defines it from the collator's
``// hop_length`` idiom (ref:src/aat/training/collate.py:340).  The CPU side of the same definition is checked in
``tests/test_host_logic.py::test_whole_utterance_frame_offsets_match_the_oracle_loop``.
"""
import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _expected_csr(batch, lengths):
    from oracle import restate

    want, base = [], 0
    for b, n in enumerate(lengths):
        _, lens, _ = batch.segments_of(b)
        off = restate.utterance_frame_offsets_loop(lens.tolist(), n)
        want.extend((base + off[:-1]).tolist())
        base += int(off[-1])
    want.append(base)
    return want, base


def test_utterance_frame_csr_matches_the_oracle_and_pools_like_torch():
    import torch

    from aat_b200 import AdaptiveAudioAmplitudeTokenizer, synth

    tok = AdaptiveAudioAmplitudeTokenizer()
    lengths = [160000, 256000, 100, 2080, 31999, 399, 400, 48000]
    waves = [synth.bursty_speech(n, 700 + i) for i, n in enumerate(lengths)]
    batch = tok.plan(lengths)
    batch.logmel(batch.pack([torch.from_numpy(w) for w in waves]))
    batch.boundaries()
    per_segment = batch.seg_off.clone(), batch._csr_totals.clone()
    seg_off, totals = batch.utterance_frame_csr()
    torch.cuda.synchronize()
    assert int(batch.status.min().item()) >= 0
    n_seg = int(batch.n_seg.item())
    want, rows = _expected_csr(batch, lengths)
    assert rows == int(synth.hubert_frames(lengths).sum())
    assert totals.cpu().tolist() == [n_seg, rows]
    assert seg_off[: n_seg + 1].cpu().tolist() == want
    # buffers of its own: the per-segment CSR of the batch is still there
    assert torch.equal(batch.seg_off, per_segment[0]) and torch.equal(batch._csr_totals, per_segment[1])

    dim = 256
    emb = torch.randn(rows, dim, device="cuda", generator=torch.Generator(device="cuda").manual_seed(12))
    out = torch.zeros(batch.total_seg_slots, dim, device="cuda")
    batch.pool(emb, out, csr=(seg_off, totals))
    torch.cuda.synchronize()
    got = out[:n_seg].cpu().numpy()
    ref = np.stack([emb[want[s]: want[s + 1]].double().mean(dim=0).cpu().numpy() if want[s + 1] > want[s]
                    else np.full(dim, np.nan) for s in range(n_seg)])
    empty = np.diff(np.asarray(want)) == 0
    assert empty.any() and not empty.all()  # the 100- and 399-sample utterances have no encoder rows
    assert np.isnan(got[empty]).all()
    np.testing.assert_allclose(got[~empty], ref[~empty], rtol=2e-5, atol=2e-6)

    # the row count taken from the device (emb is an over-allocation)
    big = torch.cat([emb, torch.full((33, dim), 1e30, device="cuda")])
    out2 = torch.empty_like(out)
    batch.pool(big, out2, csr=(seg_off, totals), rows_from_device=True)
    torch.cuda.synchronize()
    assert torch.equal(out2[:n_seg].nan_to_num(nan=-1.0), out[:n_seg].nan_to_num(nan=-1.0))


def test_utterance_frame_csr_of_equal_utterances_is_a_block_layout():
    """Equal lengths: utterance b owns rows [b * T, (b + 1) * T) and segment starts land on ``start // 320``."""
    import torch

    from aat_b200 import AdaptiveAudioAmplitudeTokenizer, synth

    tok = AdaptiveAudioAmplitudeTokenizer()
    lengths = [64000] * 300  # more utterances than threads of a CTA: the prefix of the row counts takes several strides
    batch = tok.plan(lengths)
    batch.logmel(synth.device_bursty_batch(batch, 8000, 0))
    batch.boundaries()
    seg_off, totals = batch.utterance_frame_csr()
    torch.cuda.synchronize()
    T = int(synth.hubert_frames(64000))
    n_seg = int(batch.n_seg.item())
    assert totals.cpu().tolist() == [n_seg, 300 * T]
    utt_first = batch.utt_seg_off.cpu().numpy()
    off = seg_off[: n_seg + 1].cpu().numpy()
    assert np.array_equal(off[utt_first[:-1]], np.arange(300) * T) and off[-1] == 300 * T
    want, rows = _expected_csr(batch, lengths)
    assert rows == 300 * T and off.tolist() == want
