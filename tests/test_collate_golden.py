"""Rows N1 / N2 / N4 of SURVEY.md §8f against outputs of the LIVE reference collator
(tests/golden/collate_v1.npz, written by tests/golden/make_collate_golden.py from the unmodified
``aat.training.collate.TokenizedAudioWaveformCollator``).

CPU part: the oracle's collator port and the product's host-side integer logic (``crop_to_words``,
``uniform_segment_lengths``) reproduce the fixtures.  GPU part (``-m gpu``): ``collate.collate_batch`` — z-score,
log-mel, boundaries, cropping, the feature extractor's normalisation + padding, padded boarders, waveform / log-mel
tiles, all through the C ABI — reproduces them.
"""
import json
import os

import numpy as np
import pytest
import torch

from aat_b200 import collate, synth
from oracle import collate_port

HERE = os.path.dirname(os.path.abspath(__file__))
ARR = np.load(os.path.join(HERE, "golden", "collate_v1.npz"))
INFO = json.load(open(os.path.join(HERE, "golden", "collate_v1.json")))["batches"]
OK_BATCHES = [k for k, v in INFO.items() if v["raises"] is None]


def waves_of(name):
    return [synth.bursty_speech(n, seed).astype(np.float64) + dc for n, seed, dc in INFO[name]["recipes"]]


def n_items(name):
    return len(INFO[name]["recipes"])


def crops_of(name):
    """The reference's draws (recorded by the generator) -> the `word_crops` argument; words sit every 0.4 s."""
    info = INFO[name]
    if not info["n_words"]:
        return None
    draws = info["draws"]
    n_words = draws[0][2]
    crops = []
    k = 1
    for (n, _, _), total in zip(info["recipes"], info["words_per_item"]):
        dur = n / 16000.0
        starts = [round(0.05 + 0.4 * i, 3) for i in range(total)]
        ends = [round(min(s + 0.33, dur - 0.01), 3) for s in starts]
        if total > n_words:
            crops.append(dict(word_start=starts, word_end=ends, word_start_idx=draws[k][2], n_words=n_words))
            k += 1
        else:
            crops.append(None)
    assert k == len(draws)
    return crops


# ------------------------------------------------------------------------------------------------ CPU: oracle + host logic
@pytest.mark.parametrize("name", OK_BATCHES)
def test_port_reproduces_the_reference_collator(name):
    info = INFO[name]
    boarders = [ARR[f"{name}/item{i}/segments_boarders"] for i in range(n_items(name))]
    padded, mask = collate_port.make_padded_segments_boarders(boarders, len(boarders))
    assert np.array_equal(padded.numpy(), ARR[f"{name}/segments_boarders_padded"])
    assert np.array_equal(mask.numpy(), ARR[f"{name}/segments_boarders_attention_mask"])
    mels = [ARR[f"{name}/item{i}/melspec"] for i in range(n_items(name))]
    values = torch.from_numpy(ARR[f"{name}/input_values"])
    seg, seg_mask, tiles = collate_port.scatter_segments(values, padded, 24000, items_melspecs=mels)
    if info["audio_encoder_type"] == "efficient_net":
        assert np.array_equal(tiles.numpy(), ARR[f"{name}/batched_segments_melspectrograms"])
    else:
        assert np.array_equal(seg.numpy(), ARR[f"{name}/batched_segments"])
        assert np.array_equal(seg_mask.numpy(), ARR[f"{name}/segments_waveforms_mask"])


def test_port_normalisations_reproduce_the_feature_extractor_and_the_cached_mel():
    from oracle import ref_port

    name = "adaptive_hubert"
    ref = ref_port.RefTokenizer()
    for i, w in enumerate(waves_of(name)):
        want = ARR[f"{name}/input_values"][i]
        got = collate_port.w2v2_norm(w)
        assert np.array_equal(got, want[: w.size]) and not want[w.size:].any()
        assert np.array_equal(ARR[f"{name}/input_attention_mask"][i], (np.arange(want.size) < w.size).astype(np.int64))
        mel = ref.get_melspec(collate_port.znorm(w))
        assert np.array_equal(mel, ARR[f"{name}/item{i}/melspec"])
        lengths, _, _ = ref.segment_lengths(collate_port.znorm(w))
        assert lengths == ARR[f"{name}/item{i}/segment_lengths_full"].tolist()
        assert np.array_equal(np.cumsum(lengths), ARR[f"{name}/item{i}/segments_boarders"])


def test_live_wav2vec2_feature_extractor_agrees_with_the_fixture():
    """The class the reference's processor wraps is importable here: the fixture is what it returns today."""
    from transformers import Wav2Vec2FeatureExtractor

    fe = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0, do_normalize=True,
                                  return_attention_mask=True)
    name = "uniform_hubert"
    out = fe(waves_of(name), padding=True, return_tensors="np", sampling_rate=16000)
    assert np.array_equal(out.input_values, ARR[f"{name}/input_values"])


def test_word_cropping_and_uniform_lengths_reproduce_the_reference():
    name = "adaptive_nwords"
    info = INFO[name]
    crops = crops_of(name)
    assert any(c is not None for c in crops)
    for i, ((n, _, _), crop) in enumerate(zip(info["recipes"], crops)):
        full = ARR[f"{name}/item{i}/segment_lengths_full"]
        n_mel = 1 + n // 160
        kept, (wlo, whi), (mlo, mhi), _ = collate.crop_to_words(full, crop["word_start"], crop["word_end"],
                                                                crop["word_start_idx"], crop["n_words"], 16000, 160, 12, n, n_mel)
        assert np.array_equal(kept, ARR[f"{name}/item{i}/segments_boarders"])
        assert whi - wlo == info["waveform_lengths"][i]
        assert mhi - mlo == ARR[f"{name}/item{i}/melspec"].shape[1]
        # the oracle's statement-by-statement port agrees as well
        w = waves_of(name)[i]
        mel_full = np.zeros((64, n_mel), dtype=np.float32)
        words = [f"w{j}" for j in range(info["words_per_item"][i])]
        _, w_c, mel_c, fb = collate_port.crop_to_words(w, mel_full, full, words, crop["word_start"], crop["word_end"],
                                                       crop["word_start_idx"], crop["n_words"], 16000, 160, 12)
        assert np.array_equal(fb, kept) and w_c.shape[-1] == whi - wlo and mel_c.shape[1] == mhi - mlo
    for name in ("uniform_hubert", "uniform_efficient_net", "uniform_too_long"):
        info = INFO[name]
        for i, (n, _, _) in enumerate(info["recipes"]):
            lengths = collate.uniform_segment_lengths(n, info["uniform_frames"])
            assert np.array_equal(np.cumsum(lengths), ARR[f"{name}/item{i}/segments_boarders"])
            assert int(lengths.max()) == int(ARR[f"{name}/segments_max_frame_len"][i])


# ------------------------------------------------------------------------------------------------ GPU: the product
@pytest.mark.gpu
@pytest.mark.parametrize("name", OK_BATCHES)
def test_gpu_collator_reproduces_the_reference_collator(name):
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    info = INFO[name]
    tok = AdaptiveAudioAmplitudeTokenizer()
    res = collate.collate_batch(tok, waves_of(name), audio_encoder_type=info["audio_encoder_type"],
                                segmentation=info["segmentation"],
                                uniform_segmentation_frames_per_segment=info["uniform_frames"], word_crops=crops_of(name))
    torch.cuda.synchronize()
    # integer tensors: exact
    assert res["segments_count"] == info["segments_count"]
    assert np.array_equal(res["segments_boarders_padded"].cpu().numpy(), ARR[f"{name}/segments_boarders_padded"])
    assert np.array_equal(res["segments_boarders_attention_mask"].cpu().numpy(), ARR[f"{name}/segments_boarders_attention_mask"])
    assert np.array_equal(res["segments_max_frame_len"].cpu().numpy(), ARR[f"{name}/segments_max_frame_len"])
    if info["audio_encoder_type"] == "efficient_net":
        assert res["batched_segments"] is None and res["segments_waveforms_mask"] is None
        got = res["batched_segments_melspectrograms"].cpu().numpy()
        want = ARR[f"{name}/batched_segments_melspectrograms"]
        assert got.shape == want.shape
        # log-mel of the device-normalised waveform: the float64 statistics are merged in another order than numpy's
        # pairwise sum, so a few elements in ten thousand move by one float32 unit; padding must be exactly zero
        assert np.all(np.abs(got - want) <= 1e-5 * np.maximum(1.0, np.abs(want)))
        assert np.mean(got == want) >= 0.999 and np.array_equal(got == 0.0, want == 0.0)
    else:
        assert res["batched_segments_melspectrograms"] is None
        values = res["audio_input_values"].cpu().numpy()
        want_values = ARR[f"{name}/input_values"]
        assert values.shape == want_values.shape
        assert np.all(np.abs(values - want_values) <= 2e-6 * np.maximum(1.0, np.abs(want_values)))
        assert np.array_equal(res["audio_attention_mask"].cpu().numpy(), ARR[f"{name}/input_attention_mask"])
        assert np.array_equal(res["segments_waveforms_mask"].cpu().numpy(), ARR[f"{name}/segments_waveforms_mask"])
        got, want = res["batched_segments"].cpu().numpy(), ARR[f"{name}/batched_segments"]
        assert np.all(np.abs(got - want) <= 2e-6 * np.maximum(1.0, np.abs(want)))
        assert np.array_equal(got == 0.0, want == 0.0)
    # the scatter itself is exact: fed the reference's own input_values / mels it returns the reference's tiles bit for bit
    batch = res["batch"]
    padded = res["segments_boarders_padded"]
    if info["audio_encoder_type"] == "efficient_net":
        mels = [ARR[f"{name}/item{i}/melspec"] for i in range(n_items(name))]
        flat = torch.cat([torch.from_numpy(m).reshape(-1) for m in mels]).cuda()
        off = np.concatenate([[0], np.cumsum([m.size for m in mels])[:-1]])
        frames = [m.shape[1] for m in mels]
        dev = lambda v: torch.tensor(np.asarray(v), dtype=torch.int64, device="cuda")  # noqa: E731
        tiles = collate.scatter_mel_tiles(batch, flat, dev(off), dev(frames), dev(frames), padded, 24000)
        assert np.array_equal(tiles.cpu().numpy(), ARR[f"{name}/batched_segments_melspectrograms"])
    else:
        values = torch.from_numpy(ARR[f"{name}/input_values"]).cuda()
        segs, seg_mask = collate.scatter_segments(batch, values, padded, 24000)
        assert np.array_equal(segs.cpu().numpy(), ARR[f"{name}/batched_segments"])
        assert np.array_equal(seg_mask.cpu().numpy(), ARR[f"{name}/segments_waveforms_mask"])


@pytest.mark.gpu
def test_gpu_collator_raises_where_the_reference_raises():
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer, _cabi

    name = "uniform_too_long"
    info = INFO[name]
    assert info["raises"] == "RuntimeError"  # a 30 000-sample segment does not fit the 24 000-sample tile (collate.py:333)
    with pytest.raises(_cabi.AatError):
        collate.collate_batch(AdaptiveAudioAmplitudeTokenizer(), waves_of(name), audio_encoder_type="hubert",
                              segmentation="uniform", uniform_segmentation_frames_per_segment=info["uniform_frames"])


@pytest.mark.gpu
def test_gpu_w2v2_normalisation_against_the_live_feature_extractor():
    """aat_normalize(W2V2) / aat_normalize_padded against Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm itself."""
    from transformers import Wav2Vec2FeatureExtractor

    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    rng = np.random.default_rng(3)
    waves = [synth.bursty_speech(n, 60 + i).astype(np.float64) + dc for i, (n, dc) in
             enumerate([(48000, 0.0), (4097, 2.5), (100, 0.0), (70001, -0.1)])]
    waves.append(rng.standard_normal(12345) * 1e-3)
    live = Wav2Vec2FeatureExtractor.zero_mean_unit_var_norm([w.astype(np.float32) for w in waves],
                                                            attention_mask=None)
    tok = AdaptiveAudioAmplitudeTokenizer()
    batch = tok.plan([w.size for w in waves])
    packed = batch.pack([torch.from_numpy(w) for w in waves])
    got = collate.normalize_waveforms(batch, packed.float(), "w2v2")
    padded, mask = collate.normalize_waveforms_padded(batch, packed, "w2v2")
    torch.cuda.synchronize()
    for b, w in enumerate(waves):
        o0, o1 = int(batch.wave_off[b]), int(batch.wave_off[b + 1])
        np.testing.assert_allclose(got[o0:o1].cpu().numpy(), live[b], rtol=2e-6, atol=2e-6)
        np.testing.assert_allclose(padded[b, : w.size].cpu().numpy(), live[b], rtol=2e-6, atol=2e-6)
        assert not padded[b, w.size:].any() and int(mask[b].sum()) == w.size


@pytest.mark.gpu
def test_gpu_mel_tiles_from_column_slices_match_the_port():
    """aat_scatter_mel_tiles on VIEWS (first element, columns, row stride per utterance): the cropped mels of the n-word
    path are column slices of the full ones (ref:src/aat/training/collate.py:208-212) and are not copied."""
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    rng = np.random.default_rng(8)
    tok = AdaptiveAudioAmplitudeTokenizer()
    batch = tok.plan([16000])  # only lends its context and mel count
    full = [rng.standard_normal((64, t)).astype(np.float32) for t in (500, 301, 777)]
    rngs = [(17, 480), (0, 301), (400, 777)]
    views = [m[:, lo:hi] for m, (lo, hi) in zip(full, rngs)]
    boarders = [np.cumsum(rng.integers(2000, 20000, size=k)) for k in (4, 2, 5)]
    # keep every slice inside its view: the last boarder // 160 must not exceed the view's columns
    boarders = [b[b // 160 <= v.shape[1]] for b, v in zip(boarders, views)]
    padded, _ = collate_port.make_padded_segments_boarders(boarders, 3)
    n_max = int(padded.max())
    _, _, want = collate_port.scatter_segments(torch.zeros(3, n_max), padded, 24000, items_melspecs=views)
    flat = torch.cat([torch.from_numpy(m).reshape(-1) for m in full]).cuda()
    base = np.concatenate([[0], np.cumsum([m.size for m in full])[:-1]])
    dev = lambda v: torch.tensor(np.asarray(v), dtype=torch.int64, device="cuda")  # noqa: E731
    got = collate.scatter_mel_tiles(batch, flat, dev([b + lo for b, (lo, _) in zip(base, rngs)]),
                                    dev([hi - lo for lo, hi in rngs]), dev([m.shape[1] for m in full]), padded.cuda(), 24000)
    assert np.array_equal(got.cpu().numpy(), want.numpy())
