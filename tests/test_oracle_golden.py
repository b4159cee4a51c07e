"""The oracle is pinned against outputs of the live reference (tests/golden/, made by
make_golden.py in the build container).  CPU only."""
import hashlib

import numpy as np
import pytest

from oracle import ref_port, restate

MEL_CASES = ["c1_10s", "c2_16s_u0", "c2_16s_znorm", "silence_2s", "edge_100", "edge_201", "edge_1919", "edge_1920",
             "edge_1999", "edge_2000", "edge_2080"]
ALL_CASES = MEL_CASES + ["c2_16s_u1", "c3_20s", "noise_10s", "edge_24000", "edge_24001", "edge_25999", "edge_26000",
                         "edge_48000", "edge_50000"]


def _tok(golden, case):
    info = golden.cases[case]
    if "min_ms" in info:
        return ref_port.RefTokenizer(min_segment_duration_milliseconds=info["min_ms"],
                                     max_segment_duration_milliseconds=info["max_ms"])
    return ref_port.RefTokenizer()


def test_constants_match_reference(golden):
    assert np.array_equal(restate.mel_filter_bank_slaney(), golden.get("const", "mel_filters"))
    assert np.array_equal(restate.hann_periodic(), golden.get("const", "window"))
    tok = ref_port.RefTokenizer()
    assert np.array_equal(tok.mel_filters, golden.get("const", "mel_filters"))
    assert np.array_equal(tok.window_fn, golden.get("const", "window"))


@pytest.mark.parametrize("case", MEL_CASES)
def test_logmel_port_and_restatement_bit_exact(golden, case):
    wave = golden.wave(case)
    ref = golden.get(case, "mel")
    assert np.array_equal(ref_port.RefTokenizer().get_melspec(wave), ref)
    assert np.array_equal(restate.logmel(wave), ref)


def test_logmel_naive_dft_agrees(golden, c_oracle):
    """FFT-independent check: an O(n^2) long-double DFT gives the same float32 log-mel."""
    wave = golden.wave("edge_2080").astype(np.float64)
    got = c_oracle.logmel_naive(wave, golden.get("const", "window"), golden.get("const", "mel_filters"))
    ref = golden.get("edge_2080", "mel")
    assert got.shape == ref.shape
    assert np.mean(got == ref) > 0.999
    np.testing.assert_allclose(got, ref, rtol=0, atol=1e-6)


@pytest.mark.parametrize("case", MEL_CASES)
def test_minima_intermediates_bit_exact(golden, case, c_oracle):
    mel = golden.get(case, "mel")
    ref_min = golden.get(case, "minima")
    tok = ref_port.RefTokenizer()
    m, amp, cs, rm = tok.find_amplitude_minimas(mel, return_intermediates=True)
    assert np.array_equal(m, ref_min)
    for name, arr in (("amp", amp), ("cs", cs), ("rm", rm)):
        assert np.array_equal(arr, golden.get(case, name)), name
    m2, amp2, cs2, rm2 = restate.find_minimas_seq(mel, intermediates=True)
    m3, amp3, cs3, rm3 = c_oracle.find_minimas(mel, intermediates=True)
    for got in (m2, m3):
        assert np.array_equal(got, ref_min)
    for got, name in ((amp2, "amp"), (cs2, "cs"), (rm2, "rm"), (amp3, "amp"), (cs3, "cs"), (rm3, "rm")):
        assert np.array_equal(got, golden.get(case, name)), name


@pytest.mark.parametrize("case", ALL_CASES + ["minmax_16s"])
def test_end_to_end_lengths(golden, case, c_oracle):
    wave = golden.wave(case)
    tok = _tok(golden, case)
    lengths, boarders, mel = tok.segment_lengths(wave)
    assert boarders == golden.get(case, "boarders").tolist()
    assert lengths == golden.get(case, "lengths").tolist()
    assert hashlib.sha256(np.ascontiguousarray(mel).tobytes()).hexdigest() == golden.cases[case]["mel_sha256"]
    # the integer restatements agree with the port on the same boarders
    s2, l2, _ = restate.segment_state_machine(len(wave), boarders, tok.min_segment_frames, tok.max_segment_frames)
    s3, l3, _ = c_oracle.state_machine(len(wave), boarders, tok.min_segment_frames, tok.max_segment_frames)
    assert l2 == lengths and l3.tolist() == lengths and s2 == s3.tolist()


SM_CASES = ["sm_25000", "sm_49000", "sm_48000", "sm_160000", "sm_merge_a", "sm_merge_b", "sm_100", "sm_1999",
            "sm_mm_20000", "sm_mm_12000", "sm_mm_9000", "sm_mm_17000"]


@pytest.mark.parametrize("case", SM_CASES)
def test_state_machine_known_answers(golden, case, c_oracle):
    info = golden.cases[case]
    boarders = golden.get(case, "boarders").tolist()
    ref = golden.get(case, "lengths").tolist()
    mn, mx, n = info["min_segment_frames"], info["max_segment_frames"], info["n_samples"]
    assert restate.segment_state_machine(n, boarders, mn, mx)[1] == ref
    assert c_oracle.state_machine(n, boarders, mn, mx)[1].tolist() == ref


def test_state_machine_survey_table(c_oracle):
    """Known answers listed in SURVEY.md §4 (measured on the live reference by the survey)."""
    table = [
        (2000, 24000, 25000, [25000], [23000, 2000]),
        (2000, 24000, 49000, [49000], [24000, 23000, 2000]),
        (2000, 24000, 48000, [48000], [24000, 24000]),
        (2000, 24000, 32000, [32000], [24000, 8000]),
        (2000, 24000, 160000, [160000], [24000] * 6 + [16000]),
        (2000, 24000, 10000, [1000, 1500, 5000, 9000, 10000], [5000, 4000, 2000]),
        (2000, 24000, 10000, [5000, 9500, 10000], [5000, 4500, 2000]),
        (2000, 24000, 100, [100], [2000]),
        (8000, 4000, 20000, [20000], [4000] * 5),
        (8000, 4000, 12000, [3000, 12000], [4000, 4000, 4000]),
        (8000, 4000, 9000, [9000], [4000, 0, 8000]),
    ]
    for mn, mx, n, boarders, want in table:
        assert restate.segment_state_machine(n, boarders, mn, mx)[1] == want
        assert c_oracle.state_machine(n, boarders, mn, mx)[1].tolist() == want


def test_tail_longer_than_min_raises(c_oracle):
    with pytest.raises(ValueError):
        restate.segment_state_machine(10000, [5000], 2000, 24000)
    with pytest.raises(ValueError):
        c_oracle.state_machine(10000, [5000], 2000, 24000)
    with pytest.raises(ValueError):
        ref_port.RefTokenizer().process_segments_boarders(np.zeros(10000), [5000])


@pytest.mark.parametrize("case", ["c1_10s", "c2_16s_u0", "c3_20s"])
def test_pool_port_matches_golden(golden, case, c_oracle):
    import torch

    info = golden.cases[case]
    off = golden.get(case, "frame_off")
    emb = np.random.default_rng(info["pool_seed"]).standard_normal((int(off[-1]), info["pool_dim"]), dtype=np.float32)
    assert hashlib.sha256(emb.tobytes()).hexdigest() == info["emb_sha256"]
    ref = golden.get(case, "pooled")
    got = ref_port.mean_pool_csr(torch.from_numpy(emb), off).numpy()
    assert np.array_equal(got, ref)
    # the C restatements agree within float32 round-off of a <= 74-term sum
    f64 = c_oracle.mean_pool_f64(emb, off)
    for cand in (c_oracle.mean_pool_f32(emb, off), ref[0]):
        err = np.linalg.norm(cand - f64, axis=1) / np.linalg.norm(f64, axis=1)
        assert err.max() < 1e-6


def test_long_form_boundaries(golden):
    """Config 4: one 30-minute stream (pretokenize + process_segments_boarders; tokenize() asserts < 300)."""
    wave = golden.wave("c4_30min")
    mel = restate.logmel(wave)
    assert hashlib.sha256(mel.tobytes()).hexdigest() == golden.cases["c4_30min"]["mel_sha256"]
    tok = ref_port.RefTokenizer()
    lengths, boarders, _ = tok.segment_lengths(wave, melspec=mel)
    assert np.array_equal(tok.find_amplitude_minimas(mel), golden.get("c4_30min", "minima"))
    assert lengths == golden.get("c4_30min", "lengths").tolist()
    from oracle import c_oracle as co

    assert np.array_equal(co.find_minimas(mel), golden.get("c4_30min", "minima"))
