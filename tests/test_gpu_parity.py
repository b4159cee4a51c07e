"""Parity tests proper: the CUDA path (through the C ABI) against the oracle and the committed goldens.

Bars (BASELINE.json north_star):
  * segment boundaries / offsets: bit-exact when fed the reference's mel frames;
  * log-mel: |a - b| <= 1e-5 * max(1, |ref|) elementwise (float32) — in practice the float64 pipeline
    is bit-identical on > 99.99 % of the elements, which the tests also assert;
  * pooled embeddings: ||a - b||_2 / ||ref||_2 <= 1e-5 per segment vector and
    |a - b| <= 1e-5 * max(|ref|, rms(ref)) elementwise.
"""
import hashlib

import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

MEL_TOL = 1e-5
POOL_TOL = 1e-5


@pytest.fixture(scope="module")
def tok():
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    return AdaptiveAudioAmplitudeTokenizer()


def assert_mel_close(got, ref, min_exact=0.9999):
    assert got.shape == ref.shape and got.dtype == np.float32
    err = np.abs(got.astype(np.float64) - ref.astype(np.float64))
    assert np.all(err <= MEL_TOL * np.maximum(1.0, np.abs(ref))), float(err.max())
    assert np.mean(got == ref) >= min_exact, float(np.mean(got == ref))


def assert_pooled_close(got, ref):
    got = np.asarray(got, dtype=np.float64).reshape(-1, ref.shape[-1])
    ref = np.asarray(ref, dtype=np.float64).reshape(-1, ref.shape[-1])
    assert got.shape == ref.shape
    norm = np.linalg.norm(ref, axis=1)
    rel = np.linalg.norm(got - ref, axis=1) / np.where(norm > 0, norm, 1.0)
    assert rel.max() <= POOL_TOL, float(rel.max())
    rms = np.sqrt(np.mean(ref * ref, axis=1, keepdims=True))
    assert np.all(np.abs(got - ref) <= POOL_TOL * np.maximum(np.abs(ref), rms))


# ----------------------------------------------------------------------------------------- K1 + K2
MEL_CASES = ["c1_10s", "c2_16s_u0", "c2_16s_znorm", "silence_2s", "edge_100", "edge_201", "edge_1919", "edge_1920",
             "edge_1999", "edge_2000", "edge_2080"]


@pytest.mark.parametrize("case", MEL_CASES)
def test_logmel_matches_reference_golden(tok, golden, case):
    got = tok.get_melspec(golden.wave(case))
    assert_mel_close(got, golden.get(case, "mel"))


def test_logmel_other_dtypes_follow_float64_promotion(tok, golden):
    from oracle import restate

    wave = golden.wave("c1_10s")[:40000]
    i16 = np.round(wave * 20000).astype(np.int16)
    assert_mel_close(tok.get_melspec(i16), restate.logmel(i16))
    f16 = wave.astype(np.float16)
    assert_mel_close(tok.get_melspec(f16), restate.logmel(f16))


def test_logmel_argument_errors(tok):
    with pytest.raises(ValueError):
        tok.get_melspec(np.zeros((2, 100)))
    with pytest.raises(ValueError):
        tok.get_melspec(np.zeros(0))


def test_logmel_tiny_inputs_reflect_repeatedly(tok):
    """N <= 200 makes numpy's reflect padding wrap several times."""
    from oracle import restate

    rng = np.random.default_rng(3)
    for n in (1, 2, 3, 7, 100, 159, 160, 161, 199, 200, 201, 399, 400, 401):
        w = rng.standard_normal(n)
        if n == 1:
            continue  # np.pad(mode="reflect") on a single sample is ill-defined across numpy versions
        assert_mel_close(tok.get_melspec(w), restate.logmel(w), min_exact=0.999)


# ----------------------------------------------------------------------------------------- K3
@pytest.mark.parametrize("case", MEL_CASES)
def test_minima_bit_exact_on_reference_mel(tok, golden, case):
    got = tok.find_amplitude_minimas(golden.get(case, "mel"))
    assert got.dtype == np.int64
    assert np.array_equal(got, golden.get(case, "minima"))


@pytest.mark.parametrize("case", MEL_CASES)
def test_segments_bit_exact_on_reference_mel(tok, golden, case):
    wave = golden.wave(case)
    boarders, mel = tok.pretokenize(wave, melspec=golden.get(case, "mel"))
    assert boarders == golden.get(case, "boarders").tolist()
    segs = tok.process_segments_boarders(wave, boarders)
    assert [s.shape[-1] for s in segs] == golden.get(case, "lengths").tolist()
    assert tok.segment_lengths(wave, melspec=golden.get(case, "mel")).tolist() == golden.get(case, "lengths").tolist()


ALL_CASES = MEL_CASES + ["c2_16s_u1", "c3_20s", "noise_10s", "edge_24000", "edge_24001", "edge_25999", "edge_26000",
                         "edge_48000", "edge_50000"]


@pytest.mark.parametrize("case", ALL_CASES)
def test_end_to_end_tokenize_matches_reference(tok, golden, case):
    """Own mel -> own boundaries: must reproduce the reference's segment lengths."""
    from aat_b200 import AudioWaveform

    wave = golden.wave(case)
    segments, mel = tok.tokenize(AudioWaveform(wave, 16000))
    assert [s.waveform.shape[-1] for s in segments] == golden.get(case, "lengths").tolist()
    assert mel.shape == tuple(golden.cases[case]["mel_shape"])
    boarders, _ = tok.pretokenize(wave)
    assert boarders == golden.get(case, "boarders").tolist()
    # segments are views of the caller's array except the zero-padded tail
    total = sum(s.waveform.shape[-1] for s in segments)
    assert total >= wave.shape[-1]
    if total == wave.shape[-1]:
        assert all(np.shares_memory(s.waveform, wave) for s in segments if s.waveform.size)


def test_min_greater_than_max_configuration(golden):
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    t = AdaptiveAudioAmplitudeTokenizer(min_segment_duration_milliseconds=500, max_segment_duration_milliseconds=250)
    wave = golden.wave("minmax_16s")
    assert t.segment_lengths(wave).tolist() == golden.get("minmax_16s", "lengths").tolist()


SM_CASES = ["sm_25000", "sm_49000", "sm_48000", "sm_160000", "sm_merge_a", "sm_merge_b", "sm_100", "sm_1999",
            "sm_mm_20000", "sm_mm_12000", "sm_mm_9000", "sm_mm_17000"]


@pytest.mark.parametrize("case", SM_CASES)
def test_state_machine_known_answers(golden, case):
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    info = golden.cases[case]
    t = AdaptiveAudioAmplitudeTokenizer(
        min_segment_duration_milliseconds=info["min_segment_frames"] / 16, max_segment_duration_milliseconds=info["max_segment_frames"] / 16)
    assert (t.min_segment_frames, t.max_segment_frames) == (info["min_segment_frames"], info["max_segment_frames"])
    wave = np.zeros(info["n_samples"])
    segs = t.process_segments_boarders(wave, golden.get(case, "boarders").tolist())
    assert [s.shape[-1] for s in segs] == golden.get(case, "lengths").tolist()


def test_state_machine_matches_oracle_on_random_boarders(c_oracle):
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    rng = np.random.default_rng(11)
    for min_ms, max_ms in ((125, 1500), (500, 250), (10, 30), (1000, 1000)):
        t = AdaptiveAudioAmplitudeTokenizer(min_segment_duration_milliseconds=min_ms, max_segment_duration_milliseconds=max_ms)
        for _ in range(25):
            n = int(rng.integers(1, 400000))
            k = int(rng.integers(0, 40))
            boarders = np.sort(rng.integers(0, n, size=k)).tolist() + [n]
            want_s, want_l, _ = c_oracle.state_machine(n, boarders, t.min_segment_frames, t.max_segment_frames)
            starts, lengths, _ = t._process_boarders(n, boarders)
            assert lengths.tolist() == want_l.tolist() and starts.tolist() == want_s.tolist()


def test_tail_longer_than_min_raises_like_reference(tok):
    with pytest.raises(ValueError):
        tok.process_segments_boarders(np.zeros(10000), [5000])


def test_tokenize_assertions(tok):
    from aat_b200 import AudioWaveform

    with pytest.raises(AssertionError):
        tok.tokenize(AudioWaveform(np.zeros(16000), 8000))
    # stale intended invariants of ref:src/aat/tokenizer_test.py:18-34: silence has no minima
    boarders, mel = tok.pretokenize(np.zeros(32000))
    assert boarders == [32000] and mel.shape == (64, 201) and np.all(mel == -10.0)
    segs, _ = tok.tokenize(AudioWaveform(np.zeros(32000), 16000))
    assert [s.waveform.shape[-1] for s in segs] == [24000, 8000]


def test_segment_length_invariants_on_bursty_audio(tok):
    """The four inequalities of ref:src/aat/tokenizer_test.py:46-52, on synthetic speech-like audio."""
    from aat_b200 import AudioWaveform, synth

    segs, _ = tok.tokenize(AudioWaveform(synth.bursty_speech(256000, 77), 16000))
    frames = [s.waveform.shape[0] for s in segs]
    assert min(frames) != max(frames)
    assert min(frames) >= tok.min_segment_frames and min(frames) < tok.max_segment_frames * 0.5
    assert max(frames) <= tok.max_segment_frames and max(frames) > tok.min_segment_frames * 2


def test_long_form_30min_stream(tok, golden):
    """Config 4 shape: T_mel = 180 001; the serial float32 chain must survive 44 chunk hand-overs."""
    from oracle import restate

    wave = golden.wave("c4_30min")
    ref_mel = restate.logmel(wave)
    assert hashlib.sha256(ref_mel.tobytes()).hexdigest() == golden.cases["c4_30min"]["mel_sha256"]
    assert np.array_equal(tok.find_amplitude_minimas(ref_mel), golden.get("c4_30min", "minima"))
    assert tok.segment_lengths(wave, melspec=ref_mel).tolist() == golden.get("c4_30min", "lengths").tolist()
    got_mel = tok.get_melspec(wave)
    assert_mel_close(got_mel, ref_mel)
    assert tok.segment_lengths(wave).tolist() == golden.get("c4_30min", "lengths").tolist()


# ----------------------------------------------------------------------------------------- K4
@pytest.mark.parametrize("case", ["c1_10s", "c2_16s_u0", "c3_20s"])
def test_pool_matches_reference_golden(golden, case):
    import torch

    from aat_b200 import mean_pool_segments

    info = golden.cases[case]
    off = golden.get(case, "frame_off")
    emb = np.random.default_rng(info["pool_seed"]).standard_normal((int(off[-1]), info["pool_dim"]), dtype=np.float32)
    ref = golden.get(case, "pooled")
    te = torch.from_numpy(emb)
    # the reference's own calling convention: a list of [1, n_i, D] tensors (host)
    lst = [te[off[i]:off[i + 1]].unsqueeze(0) for i in range(off.size - 1)]
    got = mean_pool_segments(lst)
    assert tuple(got.shape) == ref.shape and got.dtype == torch.float32
    assert_pooled_close(got.numpy(), ref)
    # packed CUDA form
    got2 = mean_pool_segments(te.cuda(), off)
    assert_pooled_close(got2.cpu().numpy(), ref)


def test_pool_host_list_path_copies_each_tensor_once():
    """The reference's list-of-[1, n_i, D] argument on the host goes through aat_host_mean_pool_list: float16 tensors,
    a non-contiguous one, an empty one, and the column sums."""
    import torch

    from aat_b200 import mean_pool_segments
    from oracle import ref_port

    g = torch.Generator().manual_seed(5)
    lens = [7, 74, 1, 33, 0, 12]
    for dtype in (torch.float32, torch.float16):
        lst = [torch.randn(1, n, 768, generator=g).to(dtype) for n in lens]
        lst[1] = torch.randn(1, 768, 74, generator=g).to(dtype).transpose(1, 2)  # non-contiguous view
        keep = [x for x in lst if x.shape[1] > 0]
        want = ref_port.mean_pool_segments(keep).numpy()[0]
        cs = np.zeros(769)
        got = mean_pool_segments(lst, colsum=cs).numpy()[0]
        assert got.shape == (len(lens), 768) and np.all(np.isnan(got[4]))
        live = np.asarray([n > 0 for n in lens])
        tol = 1e-6 if dtype == torch.float32 else 2e-3
        np.testing.assert_allclose(got[live], want, rtol=tol, atol=tol)
        assert cs[768] == len(lens)


def _random_offsets(rng, n_rows, lo, hi):
    cuts = [0]
    while cuts[-1] < n_rows:
        cuts.append(min(n_rows, cuts[-1] + int(rng.integers(lo, hi + 1))))
    return np.asarray(cuts, dtype=np.int64)


@pytest.mark.parametrize("dim", [768, 1024, 64, 200, 2048, 4096])
def test_pool_ragged_against_oracle(c_oracle, dim):
    import torch

    from aat_b200 import mean_pool_segments

    rng = np.random.default_rng(dim)
    n_rows = 6000 if dim <= 1024 else 1500
    emb = rng.standard_normal((n_rows, dim), dtype=np.float32)
    mixes = {"all_min": (6, 6), "all_max": (74, 74), "uniform": (6, 74), "tiny": (1, 3), "giant": (n_rows, n_rows)}
    for name, (lo, hi) in mixes.items():
        off = _random_offsets(rng, n_rows, lo, hi)
        ref = c_oracle.mean_pool_f64(emb, off)
        got = mean_pool_segments(torch.from_numpy(emb).cuda(), off).cpu().numpy()[0]
        assert_pooled_close(got, ref)
    # alternating 6 / 74
    lens = np.tile([6, 74], n_rows // 80 + 1)
    off = np.minimum(np.concatenate([[0], np.cumsum(lens)]), n_rows)
    off = off[: int(np.argmax(off == n_rows)) + 1].astype(np.int64)
    got = mean_pool_segments(torch.from_numpy(emb).cuda(), off).cpu().numpy()[0]
    assert_pooled_close(got, c_oracle.mean_pool_f64(emb, off))
    # Zipf-heavy: mostly minimal segments, a few very long ones (SURVEY §8d, config 3 ragged stress)
    lens = np.minimum(6 * rng.zipf(1.6, size=n_rows), n_rows // 3)
    off = np.minimum(np.concatenate([[0], np.cumsum(lens)]), n_rows)
    off = off[: int(np.argmax(off == n_rows)) + 1].astype(np.int64)
    got = mean_pool_segments(torch.from_numpy(emb).cuda(), off).cpu().numpy()[0]
    assert_pooled_close(got, c_oracle.mean_pool_f64(emb, off))


def test_pool_randomised_carry_paths_against_torch():
    """The cross-CTA carry has many paths: early look at the neighbour, single pieces, aligned groups of 16 CTAs with a
    leader, stragglers before and after the groups, segments that end exactly on a CTA boundary, leading / trailing
    rows outside every segment, a row count taken from the device that is smaller than the allocation.  Sixty random
    layouts (sizes chosen so that the grid is full: 296 CTAs) against a float64 torch reduction."""
    import torch

    from aat_b200 import AdaptiveAudioAmplitudeTokenizer
    from aat_b200.pooling import _pool_device

    rng = np.random.default_rng(2024)
    tok = AdaptiveAudioAmplitudeTokenizer()
    plan = tok.plan([16000])  # lends its pool scratch
    stream = plan._stream()
    for case in range(60):
        dim = int(rng.choice([64, 256, 768, 1024, 2048]))
        rows_per_stage = max(1, (24 * 1024) // (dim * 4))
        n_rows = int(rng.integers(296 * 2 * rows_per_stage, 296 * 6 * rows_per_stage))
        kind = case % 6
        if kind == 0:    # a few giants among ordinary segments
            lens = rng.integers(1, 40, size=n_rows)
            for _ in range(3):
                lens[rng.integers(0, lens.size)] = rng.integers(n_rows // 20, n_rows // 2)
        elif kind == 1:  # one segment over (almost) everything
            lens = np.asarray([n_rows - int(rng.integers(0, 50))])
        elif kind == 2:  # segments about one CTA long: every boundary is cut, some end exactly on it
            per_cta = n_rows / 296.0
            lens = np.maximum(1, rng.normal(per_cta, per_cta * 0.3, size=600).astype(np.int64))
        elif kind == 3:  # segments of 10..40 CTAs: groups with stragglers on both sides
            per_cta = n_rows / 296.0
            lens = (per_cta * rng.integers(10, 40, size=40)).astype(np.int64) + rng.integers(0, 7, size=40)
        elif kind == 4:  # tiny segments
            lens = rng.integers(1, 4, size=n_rows)
        else:            # Zipf
            lens = np.minimum(rng.zipf(1.4, size=n_rows), n_rows // 2)
        first = int(rng.integers(0, 30)) if case % 2 else 0   # rows in front of the first segment
        off = first + np.concatenate([[0], np.cumsum(lens)])
        off = off[off <= n_rows - (int(rng.integers(0, 30)) if case % 3 == 0 else 0)]
        if off.size < 2:
            off = np.asarray([first, n_rows])
        off = off.astype(np.int64)
        S = off.size - 1
        alloc = n_rows + (int(rng.integers(1, 2000)) if case % 4 == 1 else 0)  # allocation larger than the covered rows
        emb = torch.randn(alloc, dim, device="cuda")
        d_off = torch.from_numpy(off).cuda()
        out = torch.full((S, dim), float("nan"), device="cuda")
        totals = torch.tensor([S, int(off[-1])], dtype=torch.int64, device="cuda")
        from_dev = alloc != n_rows
        _pool_device(plan.ctx, emb, d_off, S, totals, out, None, stream, plan=plan.handle, rows_from_device=from_dev)
        torch.cuda.synchronize()
        cs = torch.zeros(alloc + 1, dim, dtype=torch.float64, device="cuda")
        cs[1:] = torch.cumsum(emb.double(), dim=0)
        n = (d_off[1:] - d_off[:-1]).double()
        want = ((cs[d_off[1:]] - cs[d_off[:-1]]) / n[:, None]).float()
        ok = n > 0
        err = (out[ok].double() - want[ok].double()).norm(dim=1) / want[ok].double().norm(dim=1).clamp_min(1e-30)
        assert float(err.max()) <= 1e-5, (case, kind, dim, n_rows, S, float(err.max()))
        assert bool(torch.isnan(out[~ok]).all())


def test_pool_empty_segments_gaps_and_degenerate_shapes(c_oracle):
    import torch

    from aat_b200 import mean_pool_segments

    rng = np.random.default_rng(5)
    emb = rng.standard_normal((300, 768), dtype=np.float32)
    off = np.asarray([0, 0, 10, 10, 10, 150, 300, 300], dtype=np.int64)  # empty segments at the start, middle and end
    got = mean_pool_segments(torch.from_numpy(emb).cuda(), off).cpu().numpy()[0]
    ref = c_oracle.mean_pool_f64(emb, off)
    empty = np.diff(off) == 0
    assert np.all(np.isnan(got[empty])) and np.all(np.isnan(ref[empty]))  # torch: mean of an empty slice is NaN
    assert_pooled_close(got[~empty], ref[~empty])
    # rows outside [off[0], off[S]) belong to no segment
    off2 = np.asarray([20, 60, 200], dtype=np.int64)
    got2 = mean_pool_segments(torch.from_numpy(emb).cuda(), off2).cpu().numpy()[0]
    assert_pooled_close(got2, c_oracle.mean_pool_f64(emb, off2))
    # fewer rows than CTAs
    small = emb[:5]
    off3 = np.asarray([0, 2, 5], dtype=np.int64)
    got3 = mean_pool_segments(torch.from_numpy(small).cuda(), off3).cpu().numpy()[0]
    assert_pooled_close(got3, c_oracle.mean_pool_f64(small, off3))


def test_pool_is_deterministic_and_graph_safe(c_oracle):
    import torch

    from aat_b200 import mean_pool_segments

    rng = np.random.default_rng(9)
    emb = torch.from_numpy(rng.standard_normal((20000, 768), dtype=np.float32)).cuda()
    off = _random_offsets(rng, 20000, 300, 900)  # long segments: every CTA boundary cuts one
    a = mean_pool_segments(emb, off).clone()
    for _ in range(5):
        assert torch.equal(mean_pool_segments(emb, off), a)
    assert_pooled_close(a.cpu().numpy()[0], c_oracle.mean_pool_f64(emb.cpu().numpy(), off))


def test_pool_half_precision_follows_torch(c_oracle):
    import torch

    from aat_b200 import mean_pool_segments
    from oracle import ref_port

    rng = np.random.default_rng(21)
    off = _random_offsets(rng, 3000, 6, 74)
    for dtype in (torch.float16, torch.bfloat16):
        emb = torch.from_numpy(rng.standard_normal((3000, 768), dtype=np.float32)).to(dtype)
        ref = ref_port.mean_pool_csr(emb, off).numpy()[0]
        got = mean_pool_segments(emb.cuda(), off).cpu().numpy()[0]
        # the reference rounds the mean to the input dtype: agree to one unit in the last place of that dtype
        ulp = 2.0 ** -10 if dtype == torch.float16 else 2.0 ** -7
        assert np.all(np.abs(got - ref) <= ulp * np.maximum(np.abs(ref), 2.0 ** -14))
        assert np.mean(got == ref) > 0.95


def test_pool_colsum_and_dataset_mean(c_oracle):
    import torch

    from aat_b200 import mean_pool_segments
    from aat_b200.pooling import DatasetMean

    rng = np.random.default_rng(31)
    dm = DatasetMean(768)
    pooled_all = []
    for b in range(3):
        emb = torch.from_numpy(rng.standard_normal((4000 + 500 * b, 768), dtype=np.float32)).cuda()
        off = _random_offsets(rng, emb.shape[0], 6, 74)
        out = mean_pool_segments(emb, off, colsum=dm.colsum_buffer())
        dm.accumulate()
        pooled_all.append(out[0].double().cpu())
    dm.allreduce()
    cat = torch.cat(pooled_all)
    assert dm.count == cat.shape[0]
    want = cat.mean(dim=0).float().numpy()
    np.testing.assert_allclose(dm.result().cpu().numpy(), want, rtol=1e-6, atol=1e-7)
    # running totals straight from the pool call (accumulate=True) give the same answer
    from aat_b200.pooling import _pool_device
    import ctypes

    dm2 = DatasetMean(768)
    rng = np.random.default_rng(31)
    for b in range(3):
        emb = torch.from_numpy(rng.standard_normal((4000 + 500 * b, 768), dtype=np.float32)).cuda()
        off = torch.from_numpy(_random_offsets(rng, emb.shape[0], 6, 74)).cuda()
        out = torch.empty(off.numel() - 1, 768, device="cuda")
        _pool_device(dm2.ctx, emb, off, off.numel() - 1, None, out, dm2.running_buffer(),
                     ctypes.c_void_p(torch.cuda.current_stream().cuda_stream), accumulate=True)
    assert dm2.count == dm.count
    np.testing.assert_allclose(dm2.result().cpu().numpy(), want, rtol=1e-6, atol=1e-7)


# ----------------------------------------------------------------------------------------- batched path
def test_packed_batch_matches_per_utterance_path(tok, golden):
    import torch

    from aat_b200 import synth

    lengths = [160000, 256000, 100, 2080, 31999, 48000]
    waves = [synth.bursty_speech(n, 500 + i) for i, n in enumerate(lengths)]
    batch = tok.plan(lengths)
    packed = batch.pack([torch.from_numpy(w) for w in waves])
    batch.logmel(packed)
    batch.boundaries()
    batch.frame_csr()
    torch.cuda.synchronize()
    assert int(batch.status.min().item()) >= 0
    total = 0
    for b, w in enumerate(waves):
        mel = tok.get_melspec(w)
        assert np.array_equal(batch.mel_of(b).cpu().numpy(), mel)
        assert np.array_equal(batch.minima_of(b), tok.find_amplitude_minimas(mel))
        starts, lens, tail = batch.segments_of(b)
        assert lens.tolist() == tok.segment_lengths(w).tolist()
        so = int(batch.utt_seg_off[b].item())
        off = batch.seg_off[so: so + len(lens) + 1].cpu().numpy()
        assert np.array_equal(np.diff(off), synth.hubert_frames(lens))
        total += len(lens)
    assert int(batch.n_seg.item()) == total
    # the fused amplitude curve and the one recomputed from the mel give identical segments; the frame CSR
    # built in the boundary kernel's tail equals the stand-alone kernel's
    before = batch.seg_len.clone(), batch.seg_count.clone(), batch.seg_off.clone(), batch.utt_seg_off.clone()
    batch.seg_off.zero_(), batch.n_seg.zero_(), batch.utt_seg_off.zero_()
    batch.boundaries(use_amp=False, with_csr=True)
    torch.cuda.synchronize()
    assert torch.equal(batch.seg_len, before[0]) and torch.equal(batch.seg_count, before[1])
    assert int(batch.n_seg.item()) == total
    assert torch.equal(batch.seg_off[: total + 1], before[2][: total + 1]) and torch.equal(batch.utt_seg_off, before[3])


def test_amplitude_pass_equals_the_fused_epilogue(tok, golden):
    """aat_amplitude = the log-mel kernel's fused epilogue = numpy's -10 * mel.mean(axis=0), bit for bit."""
    import torch

    from aat_b200 import synth

    lengths = [160000, 100, 31999, 256000, 2080]
    batch = tok.plan(lengths)
    wave = synth.device_bursty_batch(batch, 4000, 0)
    batch.logmel(wave, with_amp=True)
    torch.cuda.synchronize()
    fused = batch.amp.clone()
    batch.amp.zero_()
    batch.logmel(wave, with_amp=False)
    batch.amplitude()
    torch.cuda.synchronize()
    assert torch.equal(batch.amp, fused)
    mel = batch.mel_of(3).cpu().numpy()
    o0 = int(batch.frame_off[3])
    assert np.array_equal(fused[o0:o0 + mel.shape[1]].cpu().numpy(), -10 * mel.mean(axis=0))
    # on the reference's own mel
    case = "c1_10s"
    one = tok.plan([golden.cases[case]["n_samples"]])
    ref_mel = golden.get(case, "mel")
    amp = one.amplitude(torch.from_numpy(ref_mel).reshape(-1).cuda()).cpu().numpy()
    assert np.array_equal(amp, -10 * ref_mel.mean(axis=0))


def test_packed_batch_segments_reference_mel(tok, golden):
    """Device path fed the reference's own mel frames: bit-exact offsets."""
    import torch

    cases = ["c1_10s", "c2_16s_u0", "silence_2s", "edge_2080"]
    lengths = [golden.cases[c]["n_samples"] for c in cases]
    batch = tok.plan(lengths)
    mel = torch.cat([torch.from_numpy(golden.get(c, "mel")).reshape(-1) for c in cases]).cuda()
    batch.boundaries(mel=mel)
    torch.cuda.synchronize()
    for b, c in enumerate(cases):
        assert np.array_equal(batch.minima_of(b), golden.get(c, "minima"))
        assert batch.segments_of(b)[1].tolist() == golden.get(c, "lengths").tolist()


def test_pipeline_replays_in_a_cuda_graph(tok):
    import torch

    from aat_b200 import synth

    lengths = [64000] * 8
    batch = tok.plan(lengths)
    wave = batch.pack([torch.from_numpy(synth.bursty_speech(n, 900 + i)) for i, n in enumerate(lengths)])
    batch.logmel(wave), batch.boundaries(), batch.frame_csr()
    torch.cuda.synchronize()
    n_seg = int(batch.n_seg.item())
    n_rows = int(batch.seg_off[n_seg].item())
    emb = torch.randn(n_rows, 768, device="cuda")
    out = torch.empty(batch.total_seg_slots, 768, device="cuda")
    batch.pool(emb, out)
    torch.cuda.synchronize()
    want = out[:n_seg].clone()
    want_len = batch.seg_len.clone()
    g = torch.cuda.CUDAGraph()
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            batch.logmel(wave), batch.boundaries(), batch.pool(emb, out)
    for _ in range(3):
        out.zero_()
        batch.seg_len.zero_()
        g.replay()
        torch.cuda.synchronize()
        assert torch.equal(out[:n_seg], want) and torch.equal(batch.seg_len, want_len)


def test_back_to_back_steps_match_synchronised_steps(tok):
    """The kernels of a step overlap each other's tails (programmatic dependent launch), the log-mel kernel takes its
    tiles from a self-resetting counter and the boundary kernel builds the CSR by look-back over self-clearing words:
    a loop of steps enqueued without any synchronisation, on rotating inputs, must give exactly what the same steps
    give when every kernel is followed by a device synchronisation."""
    import torch

    from aat_b200 import synth
    from aat_b200.pooling import DatasetMean

    lengths = [48000, 160000, 3000, 96000, 256000, 20000, 131072, 64000]
    sets = []
    for k in range(3):
        waves = [synth.bursty_speech(n, 4000 + 17 * k + i) for i, n in enumerate(lengths)]
        sets.append(torch.from_numpy(np.concatenate(waves)).cuda())
    batch = tok.plan(lengths)
    dim = 256
    # reference pass: one kernel at a time
    want = []
    embs = []
    for k, wave in enumerate(sets):
        batch.logmel(wave)
        torch.cuda.synchronize()
        batch.boundaries()
        torch.cuda.synchronize()
        n_seg = int(batch.n_seg.item())
        n_rows = int(batch.seg_off[n_seg].item())
        g = torch.Generator(device="cuda").manual_seed(k)
        emb = torch.randn(n_rows + 64, dim, device="cuda", generator=g)  # a few rows after the last segment
        embs.append(emb)
        out = torch.zeros(batch.total_seg_slots, dim, device="cuda")
        cs = torch.zeros(dim + 1, dtype=torch.float64, device="cuda")
        batch.pool(emb, out, colsum=cs)
        torch.cuda.synchronize()
        want.append((batch.mel.clone(), batch.seg_len.clone(), batch.seg_count.clone(), batch.seg_off[: n_seg + 1].clone(),
                     out[:n_seg].clone(), cs.clone(), n_seg))
    # free-running loop: 12 steps back to back, results copied out on the same stream
    got = []
    outs = [torch.zeros(batch.total_seg_slots, dim, device="cuda") for _ in range(12)]
    css = [torch.zeros(dim + 1, dtype=torch.float64, device="cuda") for _ in range(12)]
    for it in range(12):
        k = it % 3
        batch.logmel(sets[k])
        batch.boundaries()
        batch.pool(embs[k], outs[it], colsum=css[it])
        got.append((batch.mel.clone(), batch.seg_len.clone(), batch.seg_count.clone(), batch.seg_off.clone(), batch.n_seg.clone()))
    torch.cuda.synchronize()
    for it in range(12):
        mel, seg_len, seg_count, seg_off, out, cs, n_seg = want[it % 3]
        g_mel, g_len, g_count, g_off, g_nseg = got[it]
        assert int(g_nseg.item()) == n_seg, it
        assert torch.equal(g_mel, mel) and torch.equal(g_count, seg_count), it
        assert torch.equal(g_off[: n_seg + 1], seg_off), it
        for b in range(len(lengths)):
            o, c = int(batch.seg_slot_off[b]), int(seg_count[b].item())
            assert torch.equal(g_len[o: o + c], seg_len[o: o + c]), (it, b)
        assert torch.equal(outs[it][:n_seg], out), it
        assert torch.equal(css[it], cs), it
    # nothing but the library's kernels on the stream: reduce(i) -> logmel(i + 1) overlap as well
    acc = torch.zeros(dim + 1, dtype=torch.float64, device="cuda")
    out = torch.zeros(batch.total_seg_slots, dim, device="cuda")
    for it in range(9):
        k = it % 3
        batch.logmel(sets[k])
        batch.boundaries()
        batch.pool(embs[k], out, colsum=acc, accumulate=True)
    torch.cuda.synchronize()
    mel, seg_len, seg_count, seg_off, ref_out, cs, n_seg = want[2]
    assert int(batch.n_seg.item()) == n_seg and torch.equal(batch.mel, mel)
    assert torch.equal(batch.seg_off[: n_seg + 1], seg_off) and torch.equal(out[:n_seg], ref_out)
    total = 3.0 * (want[0][5] + want[1][5] + want[2][5])
    assert torch.allclose(acc, total, rtol=1e-12, atol=0.0)


def test_two_plans_on_two_streams_are_independent(tok):
    """Tile counter, completion ticket, look-back words and the pool kernel's cross-CTA scratch belong to the plan: two
    plans driven concurrently from two streams — whole steps, pool included, with long segments so that every CTA
    boundary of the pool kernel cuts one — give what they give alone."""
    import torch

    from aat_b200 import synth

    # long max duration => segments of hundreds of HuBERT frames: the pool kernel's carry path runs at every CTA boundary
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    tok_long = AdaptiveAudioAmplitudeTokenizer(max_segment_duration_milliseconds=20000, max_amplitude_for_minima=1000)
    specs = [(tok, [64000] * 24, 768), (tok_long, [160000, 31999, 256000, 8000, 96000, 48000, 20000, 131072] * 4, 1024)]
    plans, waves, embs, outs, want = [], [], [], [], []
    for k, (t, lengths, dim) in enumerate(specs):
        batch = t.plan(lengths)
        wave = batch.pack([torch.from_numpy(synth.bursty_speech(n, 8800 + 50 * k + i)) for i, n in enumerate(lengths)])
        batch.logmel(wave), batch.boundaries()
        torch.cuda.synchronize()
        n_seg = int(batch.n_seg.item())
        n_rows = int(batch.n_frames.item())
        assert n_rows == int(batch.seg_off[n_seg].item())
        emb = torch.randn(n_rows, dim, device="cuda", generator=torch.Generator(device="cuda").manual_seed(k))
        out = torch.zeros(batch.total_seg_slots, dim, device="cuda")
        batch.pool(emb, out, emb_ready=True)
        torch.cuda.synchronize()
        want.append((batch.mel.clone(), batch.seg_count.clone(), batch.seg_off[: n_seg + 1].clone(), n_seg, out[:n_seg].clone()))
        plans.append(batch), waves.append(wave), embs.append(emb), outs.append(out)
    # the second plan's segments are long enough to be cut by the pool kernel's CTA boundaries
    assert float(np.diff(want[1][2].cpu().numpy()).max()) > 400
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    for rep in range(20):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                plans[k].logmel(waves[k])
                plans[k].boundaries()
                outs[k].zero_()
                plans[k].pool(embs[k], outs[k], emb_ready=(rep % 2 == 0))
    torch.cuda.synchronize()
    for k in (0, 1):
        mel, count, off, n_seg, pooled = want[k]
        assert int(plans[k].n_seg.item()) == n_seg
        assert torch.equal(plans[k].mel, mel) and torch.equal(plans[k].seg_count, count)
        assert torch.equal(plans[k].seg_off[: n_seg + 1], off)
        assert torch.equal(outs[k][:n_seg], pooled)


def test_pool_launches_without_a_plan_are_serialised_across_streams(c_oracle):
    """mean_pool_segments names no plan, so its launches share the context's scratch block; the library orders them
    against each other whatever streams they come from.  Long segments: every CTA boundary needs the scratch."""
    import torch

    from aat_b200 import mean_pool_segments

    rng = np.random.default_rng(77)
    cases = []
    for k in range(2):
        n_rows = 30000 + 5000 * k
        emb = torch.from_numpy(rng.standard_normal((n_rows, 768), dtype=np.float32)).cuda()
        off = _random_offsets(rng, n_rows, 400, 1500)
        want = mean_pool_segments(emb, off).clone()
        cases.append((emb, torch.from_numpy(off).cuda(), off, want))
    assert_pooled_close(cases[0][3].cpu().numpy()[0], c_oracle.mean_pool_f64(cases[0][0].cpu().numpy(), cases[0][2]))
    streams = [torch.cuda.Stream(), torch.cuda.Stream()]
    torch.cuda.synchronize()
    got = [[], []]
    for rep in range(15):
        for k in (0, 1):
            with torch.cuda.stream(streams[k]):
                got[k].append(mean_pool_segments(cases[k][0], cases[k][1]))
    torch.cuda.synchronize()
    for k in (0, 1):
        for g in got[k]:
            assert torch.equal(g, cases[k][3])


def test_pool_flags_row_count_from_the_device_and_early_fetch(tok):
    """AAT_POOL_ROWS_FROM_DEVICE: the embedding tensor is an allocation larger than what the segments cover and the
    covered row count comes from the boundary kernel; AAT_POOL_EMB_READY only changes when rows are requested;
    AAT_POOL_SHARE_SMS halves the grid."""
    import torch

    from aat_b200 import synth

    lengths = [256000] * 16
    batch = tok.plan(lengths)
    wave = synth.device_bursty_batch(batch, 2000, 0)
    batch.logmel(wave), batch.boundaries()
    torch.cuda.synchronize()
    n_seg, n_rows = int(batch.n_seg.item()), int(batch.n_frames.item())
    assert n_rows == int(batch.seg_off[n_seg].item()) and 30 * 16 <= n_seg <= 40 * 16
    big = torch.randn(n_rows + 5000, 768, device="cuda")
    exact = big[:n_rows].clone()
    out = torch.zeros(batch.total_seg_slots, 768, device="cuda")
    want = batch.pool(exact, out)[:n_seg].clone()
    for kw in (dict(rows_from_device=True), dict(emb_ready=True), dict(rows_from_device=True, emb_ready=True)):
        out.zero_()
        src = big if kw.get("rows_from_device") else exact
        # whole steps, so that the pool launch really follows the boundary kernel
        batch.logmel(wave), batch.boundaries()
        got = batch.pool(src, out, **kw)[:n_seg]
        torch.cuda.synchronize()
        assert torch.equal(got, want), kw
    # AAT_POOL_SHARE_SMS: one CTA per SM.  The CTA tile borders move, and with them the order in which a segment's rows
    # are added (rows pair up from the tile's first row): equal to float32 rounding, bit-identical among launches
    # with the flag, same under the other flags.
    shared = None
    for kw in (dict(share_sms=True), dict(share_sms=True, rows_from_device=True), dict(share_sms=True, emb_ready=True)):
        out.zero_()
        src = big if kw.get("rows_from_device") else exact
        batch.logmel(wave), batch.boundaries()
        got = batch.pool(src, out, **kw)[:n_seg].clone()
        torch.cuda.synchronize()
        assert torch.allclose(got, want, rtol=1e-6, atol=1e-7), kw
        shared = got if shared is None else shared
        assert torch.equal(got, shared), kw
    cs = torch.zeros(769, dtype=torch.float64, device="cuda")
    batch.logmel(wave), batch.boundaries()
    batch.pool(big, out, colsum=cs, rows_from_device=True)
    torch.cuda.synchronize()
    assert int(cs[768].item()) == n_seg
    assert torch.allclose(cs[:768], want.double().sum(dim=0), rtol=1e-12, atol=1e-9)


def test_pipelined_steps_match_serial_steps(tok):
    """aat_b200.pipeline.TokenizerPipeline: with 2 or 3 plans on as many streams (boundaries and pool of one batch
    overlapping the log-mel of the next), steps replayed from CUDA graphs when their input buffers come round again,
    every step gives bit for bit what the strictly serial, kernel-by-kernel loop gives, with and without the fused
    z-score; the dataset mean agrees to float64 rounding (the slots' sums are added in another order)."""
    import torch

    from aat_b200 import synth
    from aat_b200.pipeline import TokenizerPipeline

    lengths = [48000, 160000, 3000, 96000, 256000, 20000, 131072, 64000] * 3
    dim = 512
    plan = tok.plan(lengths)
    sets = [synth.device_bursty_batch(plan, 9000, 24 * k).clone() for k in range(5)]
    rows_ub = sum((n + 2000) // 320 + 2 for n in lengths)
    embs = [synth.device_normal(torch.empty(rows_ub, dim, device="cuda"), 70 + k) for k in range(5)]
    first_mel = {}
    for znorm in (False, True):
        results = {}
        # (batches in flight, amplitude curve from the fused epilogue, steps replayed from CUDA graphs)
        variants = {1: (1, True, False), 2: (2, True, True), 3: (3, True, True), -3: (3, False, True), 4: (4, True, False)}
        for depth, (n_slots, fused_amp, graphs) in variants.items():
            # the pool grid of the pipelined variants (one CTA per SM, AAT_POOL_SHARE_SMS) for the serial reference too:
            # the flag moves the CTA tile borders, hence the last bit of segments that straddle one
            pipe = TokenizerPipeline(tok, lengths, dim, depth=n_slots, fused_amp=fused_amp, graphs=graphs, share_sms=True)
            assert pipe.share_sms
            got = []
            for it in range(21):  # every (slot, input set) pair comes round again: captured and replayed CUDA graphs
                k = it % 5
                slot = pipe.submit(sets[k], embs[k], znorm=znorm, rows_from_device=True, inputs_ready=True)
                with torch.cuda.stream(slot.stream):  # copies ordered behind the step, ahead of the slot's reuse
                    got.append((slot.batch.mel.clone(), slot.batch.seg_len.clone(), slot.batch.seg_count.clone(),
                                slot.batch.seg_off.clone(), slot.batch._csr_totals.clone(), slot.out.clone()))
            dm = pipe.dataset_mean()
            mean = dm.result()
            torch.cuda.synchronize()
            results[depth] = (got, dm.acc.clone(), mean.clone())
            with pytest.raises(RuntimeError):  # the sums are folded: another pass needs reset_sums() first
                pipe.submit(sets[0], embs[0], rows_from_device=True, inputs_ready=True)
            pipe.reset_sums()
            pipe.submit(sets[0], embs[0], znorm=znorm, rows_from_device=True, inputs_ready=True)
            torch.cuda.synchronize()
        ref_steps, ref_acc, ref_mean = results[1]
        assert int(ref_acc[dim].item()) == sum(int(g[4][0].item()) for g in ref_steps)
        for depth in (2, 3, -3, 4):
            steps, acc, mean = results[depth]
            for it, (a, b) in enumerate(zip(ref_steps, steps)):
                n_seg = int(a[4][0].item())
                assert torch.equal(a[4], b[4]), (depth, it)
                assert torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]), (depth, it)
                assert torch.equal(a[3][: n_seg + 1], b[3][: n_seg + 1]), (depth, it)
                assert torch.equal(a[5][:n_seg], b[5][:n_seg]), (depth, it)
            assert int(acc[dim].item()) == int(ref_acc[dim].item())
            assert torch.allclose(acc, ref_acc, rtol=1e-13, atol=1e-9) and torch.allclose(mean, ref_mean, rtol=1e-6, atol=1e-9)
        first_mel[znorm] = results[1][0][0][0]
    assert not torch.equal(first_mel[False], first_mel[True])  # the fused z-score really normalised


def test_profile_sampling_records_every_nth_launch(tok):
    import torch

    from aat_b200 import _cabi, synth

    lengths = [32000] * 4
    batch = tok.plan(lengths)
    wave = batch.pack([torch.from_numpy(synth.bursty_speech(n, 60 + i)) for i, n in enumerate(lengths)])
    _cabi.profile_enable(batch.ctx.handle, ("logmel", "boundaries"), every=4)
    for _ in range(10):
        batch.logmel(wave), batch.boundaries()
    torch.cuda.synchronize()
    prof = _cabi.profile_summary(batch.ctx.handle)
    _cabi.profile_enable(batch.ctx.handle, ())
    assert prof["logmel"][0] == 3 and prof["boundaries"][0] == 3 and prof["pool"][0] == 0
    assert prof["logmel"][1] > 0.0


# ----------------------------------------------------------------------------------------- full-size properties
def test_config2_full_size_properties(tok):
    """BASELINE config 2 (64 x 16 s, D = 768): size-independent properties on the whole batch plus
    oracle parity on a sample of utterances."""
    import torch

    from aat_b200 import synth
    from oracle import ref_port

    B, N, D = 64, 256000, 768
    waves = [synth.bursty_speech(N, synth.seed_for(2, i)) for i in range(B)]
    batch = tok.plan([N] * B)
    batch.logmel(batch.pack([torch.from_numpy(w) for w in waves]))
    batch.boundaries()
    batch.frame_csr()
    torch.cuda.synchronize()
    assert int(batch.status.min().item()) >= 0
    n_seg = int(batch.n_seg.item())
    seg_off = batch.seg_off[: n_seg + 1].cpu().numpy()
    assert np.all(np.diff(seg_off) >= 0) and seg_off[0] == 0
    ref = ref_port.RefTokenizer()
    for b in range(B):
        starts, lens, tail = batch.segments_of(b)
        assert lens.sum() >= N and (lens.sum() > N) == tail
        assert lens.min() >= tok.min_segment_frames and lens.max() <= tok.max_segment_frames
        assert np.array_equal(starts[1:], np.cumsum(lens)[:-1])
        if b % 16 == 0:
            want, _, mel = ref.segment_lengths(waves[b])
            assert lens.tolist() == want
            assert_mel_close(batch.mel_of(b).cpu().numpy(), mel)
    # pooling: linearity and the frame-weighted checksum  sum_s n_s * pooled[s] == column sums of E
    n_rows = int(seg_off[-1])
    g = torch.Generator(device="cuda").manual_seed(1)
    e1 = torch.randn(n_rows, D, device="cuda", generator=g)
    e2 = torch.randn(n_rows, D, device="cuda", generator=g)
    out = torch.empty(batch.total_seg_slots, D, device="cuda")
    p1 = batch.pool(e1, out)[:n_seg].clone()
    p2 = batch.pool(e2, out)[:n_seg].clone()
    p3 = batch.pool(2.0 * e1 + e2, out)[:n_seg].clone()
    torch.cuda.synchronize()
    assert torch.allclose(p3, 2.0 * p1 + p2, rtol=1e-5, atol=1e-5)
    counts = torch.from_numpy(np.diff(seg_off)).cuda().double()
    lhs = (p1.double() * counts[:, None]).sum(dim=0)
    rhs = e1.double().sum(dim=0)
    assert torch.allclose(lhs, rhs, rtol=1e-6, atol=1e-3)
    want = ref_port.mean_pool_csr(e1[: seg_off[40]].cpu(), seg_off[:41]).numpy()[0]
    assert_pooled_close(p1[:40].cpu().numpy(), want)


def _full_size_properties(tok, B, N, D, config, oracle_utts):
    """Size-independent properties of a whole BASELINE-sized batch (device-generated audio), vectorised on
    the device, plus bit-exact oracle parity of the segment lengths on a sample of utterances."""
    import torch

    from aat_b200 import synth
    from oracle import ref_port

    batch = tok.plan([N] * B)
    wave = synth.device_bursty_batch(batch, 1000 * config, 0)
    batch.logmel(wave), batch.boundaries()
    torch.cuda.synchronize()
    assert int(batch.status.min().item()) >= 0
    cap = int(batch.seg_slot_off[1])
    lens = batch.seg_len.view(B, cap)
    starts = batch.seg_start.view(B, cap)
    count = batch.seg_count.long()
    live = torch.arange(cap, device="cuda")[None, :] < count[:, None]
    total = (lens * live).sum(dim=1)
    tail = (batch.status & 1).bool()
    assert bool((total >= N).all()) and bool(((total > N) == tail).all())             # sum of lengths >= N, > iff padded tail
    assert int(lens[live].min()) >= tok.min_segment_frames and int(lens[live].max()) <= tok.max_segment_frames
    want_starts = torch.cumsum(lens * live, dim=1) - lens * live                        # starts = exclusive cumsum of lengths
    assert bool((starts[live] == want_starts[live]).all())
    n_seg = int(batch.n_seg.item())
    assert n_seg == int(count.sum().item())
    seg_off = batch.seg_off[: n_seg + 1]
    frames = torch.clamp((lens[live] - 400) // 320 + 1, min=0)                        # row-major == packed order
    assert bool((seg_off[1:] - seg_off[:-1] == frames).all()) and int(seg_off[0]) == 0
    assert int(batch.n_frames.item()) == int(seg_off[-1].item())
    ref = ref_port.RefTokenizer()
    for b in oracle_utts:
        w = wave[b * N:(b + 1) * N].cpu().numpy()
        c, o = int(count[b]), b * cap
        assert batch.seg_len[o:o + c].tolist() == ref.segment_lengths(w)[0], b
    # pooling: the frame-weighted checksum  sum_s n_s * pooled[s] == column sums of E, and a sample against torch
    n_rows = int(seg_off[-1].item())
    emb = synth.device_normal(torch.empty(n_rows, D, device="cuda"), 11)
    out = torch.empty(batch.total_seg_slots, D, device="cuda")
    cs = torch.zeros(D + 1, dtype=torch.float64, device="cuda")
    pooled = batch.pool(emb, out, colsum=cs)[:n_seg]
    torch.cuda.synchronize()
    lhs = (pooled.double() * frames.double()[:, None]).sum(dim=0)
    rhs = emb.double().sum(dim=0)
    assert torch.allclose(lhs, rhs, rtol=1e-6, atol=2e-3 * (n_rows / 5e4) ** 0.5)
    assert int(cs[D].item()) == n_seg and torch.allclose(cs[:D], pooled.double().sum(dim=0), rtol=1e-12, atol=1e-9)
    k = min(n_seg, 64)
    off = seg_off[: k + 1].cpu().numpy()
    want = ref_port.mean_pool_csr(emb[: off[-1]].cpu(), off).numpy()[0]
    assert_pooled_close(pooled[:k].cpu().numpy(), want)


def test_config3_full_size_properties(tok):
    """BASELINE config 3: 256 x 20 s, HuBERT-large 1024-d."""
    _full_size_properties(tok, 256, 320000, 1024, 3, oracle_utts=(0, 101, 255))


def test_config4_full_size_properties(tok):
    """BASELINE config 4: 8 x 30-min streams, 768-d (one stream checked against the oracle: ~4 s of CPU)."""
    _full_size_properties(tok, 8, 28_800_000, 768, 4, oracle_utts=(5,))


# ----------------------------------------------------------------------------------------- N1 / N2
def test_waveform_normalisations(tok):
    import torch

    from aat_b200 import collate, synth
    from oracle import collate_port

    rng = np.random.default_rng(17)
    lengths = [256000, 4095, 4096, 4097, 100, 70001]
    waves = [synth.bursty_speech(n, 600 + i).astype(np.float64) + (0.0 if i % 2 else 3.5) for i, n in enumerate(lengths)]
    batch = tok.plan(lengths)
    packed = batch.pack([torch.from_numpy(w) for w in waves])
    out, stats = collate.normalize_waveforms(batch, packed, "zscore", return_stats=True)
    out32 = collate.normalize_waveforms(batch, packed.float(), "w2v2")
    torch.cuda.synchronize()
    assert out.dtype == torch.float64 and out32.dtype == torch.float32
    for b, w in enumerate(waves):
        o0, o1 = int(batch.wave_off[b]), int(batch.wave_off[b + 1])
        want = collate_port.znorm(w)
        np.testing.assert_allclose(out[o0:o1].cpu().numpy(), want, rtol=1e-12, atol=1e-12)
        assert abs(float(stats[b, 0]) - w.mean()) <= 1e-13 * max(1.0, abs(w.mean()))
        assert abs(float(stats[b, 1]) - w.var()) <= 1e-12 * w.var()
        want32 = collate_port.w2v2_norm(w.astype(np.float32))
        np.testing.assert_allclose(out32[o0:o1].cpu().numpy(), want32, rtol=2e-6, atol=2e-6)


def test_znorm_then_logmel_matches_reference_golden(tok, golden):
    """Device z-score -> K1+K2 reproduces the reference's mel of the numpy-normalised waveform."""
    import torch

    from aat_b200 import collate, synth

    raw = synth.bursty_speech(256000, synth.seed_for(2, 2)).astype(np.float64)
    batch = tok.plan([raw.size])
    normed = collate.normalize_waveforms(batch, torch.from_numpy(raw).cuda(), "zscore")
    batch.logmel(normed)
    batch.boundaries()
    torch.cuda.synchronize()
    assert_mel_close(batch.mel_of(0).cpu().numpy(), golden.get("c2_16s_znorm", "mel"), min_exact=0.999)
    assert batch.segments_of(0)[1].tolist() == golden.get("c2_16s_znorm", "lengths").tolist()


def test_fused_znorm_is_bit_identical_to_the_separate_pass(tok, golden):
    """aat_logmel with znorm_stats: the z-score applied while the samples are staged (reciprocal + one exact-remainder
    correction instead of a division per sample) gives bit for bit the log-mel of the separately normalised float64
    waveform, for float64 and float32 input, including utterances shorter than one frame and a DC offset."""
    import torch

    from aat_b200 import collate, synth

    lengths = [256000, 4097, 100, 70001, 31999, 1]
    waves = [synth.bursty_speech(n, 640 + i).astype(np.float64) * (1.0 + 3.0 * i) + (0.0 if i % 2 else 2.5)
             for i, n in enumerate(lengths)]
    batch = tok.plan(lengths)
    for dtype in (torch.float64, torch.float32):
        packed = batch.pack([torch.from_numpy(w) for w in waves]).to(dtype)
        normed = collate.normalize_waveforms(batch, packed, "zscore")  # float64 out
        batch.logmel(normed)
        batch.boundaries()
        torch.cuda.synchronize()
        want = (batch.mel.clone(), batch.amp.clone(), batch.seg_len.clone(), batch.seg_count.clone())
        stats = batch.waveform_stats(packed)
        batch.mel.zero_(), batch.amp.zero_()
        batch.logmel(packed, znorm_stats=stats)
        batch.boundaries()
        torch.cuda.synchronize()
        assert torch.equal(batch.mel, want[0]) and torch.equal(batch.amp, want[1])
        assert torch.equal(batch.seg_count, want[3])
        for b in range(len(lengths)):
            w = waves[b] if dtype == torch.float64 else waves[b].astype(np.float32).astype(np.float64)
            assert abs(float(stats[b, 0]) - w.mean()) <= 1e-12 * max(1.0, abs(w.mean()))
            assert abs(float(stats[b, 1]) - w.var()) <= 1e-11 * max(w.var(), 1e-300)
    # against the live reference's mel of the numpy-normalised waveform
    raw = synth.bursty_speech(256000, synth.seed_for(2, 2)).astype(np.float64)
    one = tok.plan([raw.size])
    d_raw = torch.from_numpy(raw).cuda()
    one.logmel(d_raw, znorm_stats=one.waveform_stats(d_raw))
    one.boundaries()
    torch.cuda.synchronize()
    assert_mel_close(one.mel_of(0).cpu().numpy(), golden.get("c2_16s_znorm", "mel"), min_exact=0.999)
    assert one.segments_of(0)[1].tolist() == golden.get("c2_16s_znorm", "lengths").tolist()


def test_padded_layout_matches_collator_port(tok):
    import torch

    from aat_b200 import collate, synth
    from oracle import collate_port

    lengths = [64000, 40000, 96000, 2080]
    waves = [synth.bursty_speech(n, 700 + i) for i, n in enumerate(lengths)]
    batch = tok.plan(lengths)
    batch.logmel(batch.pack([torch.from_numpy(w) for w in waves]))
    batch.boundaries()
    torch.cuda.synchronize()
    seg_lengths = [batch.segments_of(b)[1] for b in range(len(lengths))]
    boarders = [np.cumsum(l) for l in seg_lengths]
    want_pad, want_mask = collate_port.make_padded_segments_boarders(boarders, len(lengths))
    got_pad, got_mask = collate.pad_segment_boarders(batch)
    assert torch.equal(got_pad.cpu(), want_pad) and torch.equal(got_mask.cpu(), want_mask)

    # the feature extractor's padded input_values: [B, N_max] float32, long enough for the zero-padded tails
    n_max = max(int(b[-1]) for b in boarders)
    padded_wave = np.zeros((len(lengths), n_max), dtype=np.float32)
    for b, w in enumerate(waves):
        padded_wave[b, : w.size] = collate_port.w2v2_norm(w)
    max_frames = tok.max_segment_frames
    mels = [batch.mel_of(b).cpu().numpy() for b in range(len(lengths))]
    want_seg, want_segmask, want_mel = collate_port.scatter_segments(torch.from_numpy(padded_wave), want_pad, max_frames,
                                                                     items_melspecs=mels)
    got_seg, got_segmask = collate.scatter_segments(batch, torch.from_numpy(padded_wave).cuda(), got_pad, max_frames)
    got_mel = collate.scatter_mel_segments(batch, got_pad, max_frames)
    torch.cuda.synchronize()
    assert torch.equal(got_seg.cpu(), want_seg) and torch.equal(got_segmask.cpu(), want_segmask)
    assert torch.equal(got_mel.cpu(), want_mel)
    # a tile that is too small is what makes the reference raise
    with pytest.raises(Exception):
        collate.scatter_segments(batch, torch.from_numpy(padded_wave).cuda(), got_pad, 1000)


# ----------------------------------------------------------------------------------------- non-default constructor arguments
@pytest.mark.parametrize("kwargs", [
    dict(hop_length=200, num_mel_filters=80),
    dict(hop_length=100, num_mel_filters=40, running_mean_points=5, max_amplitude_for_minima=12),
    dict(hop_length=160, num_mel_filters=128, running_mean_points=30, min_segment_duration_milliseconds=50,
         max_segment_duration_milliseconds=400),
    dict(hop_length=400, num_mel_filters=23, running_mean_points=3, max_amplitude_for_minima=16.5),
    dict(sampling_rate=8000, hop_length=80, num_mel_filters=32),
])
def test_non_default_configurations_match_oracle(kwargs):
    """hop, mel count, running-mean width, gate and durations are free parameters of the reference's
    constructor (ref:src/aat/tokenizer.py:15-24); n_fft stays 400."""
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer, AudioWaveform, synth
    from oracle import ref_port

    import warnings

    tok = AdaptiveAudioAmplitudeTokenizer(**kwargs)
    with warnings.catch_warnings():  # sampling_rate=8000 leaves the filters above 4 kHz empty: transformers warns
        warnings.simplefilter("ignore")
        ref = ref_port.RefTokenizer(**kwargs)
    assert np.array_equal(tok.mel_filters, ref.mel_filters)
    for seed, n in ((1, 96000), (2, 31999), (3, 4000)):
        wave = synth.bursty_speech(n, 4000 + seed)
        mel = tok.get_melspec(wave)
        want_mel = ref.get_melspec(wave)
        assert_mel_close(mel, want_mel, min_exact=0.999)
        assert np.array_equal(tok.find_amplitude_minimas(want_mel), ref.find_amplitude_minimas(want_mel))
        want_lengths, want_boarders, _ = ref.segment_lengths(wave, melspec=want_mel)
        assert tok.segment_lengths(wave, melspec=want_mel).tolist() == want_lengths
        boarders, _ = tok.pretokenize(wave)
        assert boarders == ref.pretokenize(wave)[0]


@pytest.mark.parametrize("kwargs, n, kind", [
    (dict(min_segment_duration_milliseconds=10, max_segment_duration_milliseconds=20), 960000, "speech"),
    (dict(min_segment_duration_milliseconds=10, max_segment_duration_milliseconds=50), 1920000, "silence"),
    (dict(min_segment_duration_milliseconds=30, max_segment_duration_milliseconds=35), 480000, "speech"),
    (dict(min_segment_duration_milliseconds=40, max_segment_duration_milliseconds=25), 320000, "speech"),  # min > max
])
def test_many_segments_per_chunk_and_long_splits(kwargs, n, kind):
    """The boundary kernel queues segments in shared memory (128 per flush) and cuts over-long gaps with a resumable
    np.split: short durations make one 512-frame chunk overflow the queue several times, and a silent stretch makes
    ONE boarder produce thousands of pieces across many flushes."""
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer, synth
    from oracle import ref_port

    tok = AdaptiveAudioAmplitudeTokenizer(**kwargs)
    ref = ref_port.RefTokenizer(**kwargs)
    wave = synth.bursty_speech(n, 77) if kind == "speech" else np.zeros(n, dtype=np.float32)
    if kind == "silence":
        wave[:16000] = synth.bursty_speech(16000, 78)  # a second of sound, then nothing
    want_mel = ref.get_melspec(wave)
    want_lengths, _, _ = ref.segment_lengths(wave, melspec=want_mel)
    got = tok.segment_lengths(wave, melspec=want_mel).tolist()
    assert len(want_lengths) > 800
    assert got == want_lengths


def test_unsupported_fft_length_raises_not_implemented():
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    with pytest.raises(NotImplementedError):
        AdaptiveAudioAmplitudeTokenizer(n_fft=512).get_melspec(np.zeros(4000))


def test_masked_mean_pool_in_padded_layout():
    """mean over valid frames of [B*S, L, D] under the feature mask."""
    import torch

    from aat_b200 import collate
    from oracle import collate_port

    g = torch.Generator().manual_seed(3)
    for dtype, D, tol in ((torch.float32, 768, 1e-6), (torch.float32, 1024, 1e-6), (torch.float16, 768, 2e-3),
                          (torch.bfloat16, 512, 2e-2)):
        R, L = 97, 74
        x = torch.randn(R, L, D, generator=g).to(dtype)
        lengths = torch.randint(0, L + 1, (R,), generator=g)
        mask = (torch.arange(L)[None, :] < lengths[:, None]).long()
        mask[5] = 0                       # a fully padded segment
        mask[7, ::3] = 0                  # a non-prefix mask
        want, want_rows = collate_port.masked_mean_pool(x, mask)
        got, rows = collate.masked_mean_pool(x.cuda(), mask.cuda())
        assert torch.equal(rows.cpu(), want_rows)
        assert torch.allclose(got.cpu().double(), want, rtol=tol, atol=tol)
        assert float(got[5].abs().max()) == 0.0


# ----------------------------------------------------------------------------------------- on-device synthetic inputs
def test_device_generator_is_counter_based_and_exercises_the_ragged_path(tok):
    """aat_synth_*: utterance u is a pure function of (seed, u) whatever batch it is generated in; the recipe yields the
    31-36 segments per 16 s the survey validated for the host generator; embeddings are N(0, 1)."""
    import torch

    from aat_b200 import synth

    B, N = 8, 256000
    batch = tok.plan([N] * B)
    a = synth.device_bursty_batch(batch, 5000, 40).clone()
    b = synth.device_bursty_batch(batch, 5000, 40)
    assert torch.equal(a, b)                                    # deterministic
    c = synth.device_bursty_batch(batch, 5000, 44)
    assert torch.equal(c[: 4 * N], a[4 * N:])                    # utterance 44..47 is the same wherever it is generated
    assert not torch.equal(c[4 * N:], a[: 4 * N])
    ragged = tok.plan([N, 100, 31999, 4097])                     # another plan shape, same utterance
    d = synth.device_bursty_batch(ragged, 5000, 40)
    assert torch.equal(d[:N], a[:N]) and torch.equal(d[N + 100: N + 100 + 31999], a[2 * N: 2 * N + 31999])
    x = a.view(B, N).double()
    assert abs(float(x.mean())) < 1e-3 and 0.05 < float(x.std()) < 0.6 and float(x.abs().max()) < 6.0
    assert float(x.abs().view(B, -1, 160).amax(dim=2).min()) < 0.02   # pauses sit at the 1e-3 floor
    batch.logmel(a), batch.boundaries()
    torch.cuda.synchronize()
    counts = batch.seg_count.cpu().numpy()
    assert counts.min() >= 25 and counts.max() <= 42, counts
    # ... and the host path agrees with the device path on the device-generated audio (bit-exact segment lengths)
    from oracle import ref_port

    ref = ref_port.RefTokenizer()
    w0 = a[:N].cpu().numpy()
    assert batch.segments_of(0)[1].tolist() == ref.segment_lengths(w0)[0]
    e = torch.empty(1_000_003, device="cuda")
    synth.device_normal(e, 9)
    e2 = torch.empty(1_000_003, device="cuda")
    synth.device_normal(e2, 9)
    assert torch.equal(e, e2) and abs(float(e.mean())) < 5e-3 and abs(float(e.std()) - 1.0) < 5e-3
    assert abs(float((e ** 4).mean()) - 3.0) < 0.05                  # Gaussian kurtosis
    synth.device_normal(e2, 10)
    assert not torch.equal(e, e2)


def test_pad_segment_boarders_reports_overflow(tok):
    import torch

    from aat_b200 import _cabi, collate, synth

    batch = tok.plan([256000, 64000])
    batch.logmel(synth.device_bursty_batch(batch, 2000, 0)), batch.boundaries()
    torch.cuda.synchronize()
    s_max = int(batch.seg_count.max().item())
    padded, mask = collate.pad_segment_boarders(batch)
    assert padded.shape == (2, s_max) and int(mask.sum().item()) == int(batch.seg_count.sum().item())
    with pytest.raises(_cabi.AatError):
        collate.pad_segment_boarders(batch, s_max=s_max - 1)


def _two_gpu_worker(rank, world, port, out_path):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    # the tokenizer names its device explicitly while the CURRENT device stays 0 on both ranks until set below
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer, synth
    from aat_b200.pooling import DatasetMean

    tok = AdaptiveAudioAmplitudeTokenizer(device=rank)
    dev = torch.device("cuda", rank)
    batch = tok.plan([256000] * 8)
    dm = DatasetMean(768, device=rank)
    ref = torch.zeros(769, dtype=torch.float64, device=dev)
    out = torch.zeros(batch.total_seg_slots, 768, device=dev)
    emb = torch.empty(8 * 810, 768, device=dev)
    for step in range(3):
        # launched from a thread whose current device is cuda:0 on BOTH ranks: the entry points must run on the plan's device
        wave = synth.device_bursty_batch(batch, 3000, (rank * 3 + step) * 8)
        synth.device_normal(emb, 100 * rank + step)
        batch.logmel(wave), batch.boundaries()
        batch.pool(emb, out, colsum=dm.running_buffer(), accumulate=True, rows_from_device=True)
        torch.cuda.synchronize(dev)
        s = int(batch.n_seg.item())
        ref[:768] += out[:s].double().sum(dim=0)
        ref[768] += s
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    dm.allreduce()
    mean = dm.result()
    parts = [torch.empty_like(ref) for _ in range(world)]
    dist.all_gather(parts, ref)
    want = torch.stack(parts).sum(dim=0)
    rel = float(((dm.acc[:768] - want[:768]).abs().max() / want[:768].abs().max()).item())
    ok = int(dm.acc[768].item()) == int(want[768].item()) and rel <= 1e-12
    ok = ok and torch.allclose(mean, (want[:768] / want[768]).float(), rtol=1e-6, atol=1e-9)
    if rank == 0:
        with open(out_path, "w") as f:
            f.write("ok" if ok else f"mismatch rel={rel}")
    dist.destroy_process_group()


def test_dataset_mean_allreduce_on_two_gpus(tmp_path):
    """A9 at N > 1: DatasetMean.allreduce (NCCL) + result against an independently reduced value (torch float64 sums of
    the pooled vectors, all_gather'ed and added in rank order), to 1e-12.  Also exercises a tokenizer bound to a device
    that is not the thread's current device."""
    import socket

    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs (gpurun --gpus 2)")
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "result.txt")
    mp.spawn(_two_gpu_worker, args=(2, port, out), nprocs=2, join=True)
    assert open(out).read() == "ok"
