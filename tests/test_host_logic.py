"""Host-side logic and the C-ABI surface — no GPU, no compute calls."""
import ctypes
import os
import pickle
import re

import numpy as np
import pytest

import aat_b200
from aat_b200 import _cabi, constants, dist, synth
from aat_b200.audio import AudioWaveform
from aat_b200.tokenizer import AdaptiveAudioAmplitudeTokenizer, _wave_arg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_constants_bit_identical_to_transformers(golden):
    assert np.array_equal(constants.mel_filter_bank_slaney(201, 64, 0.0, 8000.0, 16000), golden.get("const", "mel_filters"))
    assert np.array_equal(constants.hann_window_periodic(400), golden.get("const", "window"))
    tf = pytest.importorskip("transformers.audio_utils")
    for n_mels in (40, 64, 80, 128):
        want = tf.mel_filter_bank(201, n_mels, 0.0, 8000.0, 16000, norm="slaney", mel_scale="slaney")
        assert np.array_equal(constants.mel_filter_bank_slaney(201, n_mels, 0.0, 8000.0, 16000), want)
    assert np.array_equal(constants.hann_window_periodic(400), tf.window_function(400, "hann"))


def test_tokenizer_attributes_match_reference_defaults():
    t = AdaptiveAudioAmplitudeTokenizer()
    assert (t.running_mean_points, t.n_fft, t.hop_length, t.num_mel_filters, t.sampling_rate) == (12, 400, 160, 64, 16000)
    assert (t.min_segment_frames, t.max_segment_frames, t.max_amplitude_for_minima) == (2000, 24000, 15)
    assert t.min_segment_duration_milliseconds == 125 and t.max_segment_duration_milliseconds == 1500
    assert t.mel_filters.shape == (201, 64) and t.mel_filters.dtype == np.float64
    assert t.window_fn.shape == (400,) and t.window_fn[0] == 0.0
    assert int((t.mel_filters != 0).sum()) == 388
    # the min > max configuration the reference really runs (ref:scripts/trainer_train.py:116-122)
    t2 = AdaptiveAudioAmplitudeTokenizer(min_segment_duration_milliseconds=500, max_segment_duration_milliseconds=250)
    assert (t2.min_segment_frames, t2.max_segment_frames) == (8000, 4000)
    assert t.milliseconds_to_frames(125) == 2000


def test_padding_helpers():
    t = AdaptiveAudioAmplitudeTokenizer()
    w = np.arange(1, 6, dtype=np.float32)
    r = t.right_pad_waveform_with_zeros(w)
    l = t.left_pad_waveform_with_zeros(w)
    assert r.shape == (2000,) and r.dtype == np.float64 and r[:5].tolist() == [1, 2, 3, 4, 5] and r[5:].sum() == 0
    assert l[-5:].tolist() == [1, 2, 3, 4, 5] and l[:-5].sum() == 0


def test_audio_waveform_contract():
    w = AudioWaveform(np.zeros(32000), 16000)
    assert w.duration_seconds == 2.0
    w.assert_sampling_rate(16000)
    with pytest.raises(AssertionError):
        w.assert_sampling_rate(8000)
    with pytest.raises(AssertionError):
        AudioWaveform(np.zeros((2, 10)), 16000)


def test_tokenizer_pickles_without_native_state():
    t = AdaptiveAudioAmplitudeTokenizer(min_segment_duration_milliseconds=500)
    t2 = pickle.loads(pickle.dumps(t))
    assert t2.min_segment_frames == 8000 and np.array_equal(t2.mel_filters, t.mel_filters)


def test_wave_argument_promotion():
    a, dt = _wave_arg(np.zeros(10, dtype=np.float32))
    assert dt == _cabi.AAT_F32 and a.dtype == np.float32
    a, dt = _wave_arg(np.arange(10, dtype=np.int16))
    assert dt == _cabi.AAT_F64 and a.dtype == np.float64
    with pytest.raises(ValueError):
        _wave_arg(np.zeros((2, 5)))
    with pytest.raises(ValueError):
        _wave_arg(np.zeros(4, dtype=np.complex64))


def test_library_exports_every_declared_symbol():
    header = open(os.path.join(ROOT, "include", "aat_b200.h")).read()
    header = re.sub(r"/\*.*?\*/", "", header, flags=re.S)
    declared = set(re.findall(r"\b(aat_[a-z0-9_]+)\s*\(", header))
    assert len(declared) >= 26
    assert declared == set(_cabi.SIGNATURES), declared ^ set(_cabi.SIGNATURES)
    lib = ctypes.CDLL(_cabi.LIB_PATH)
    for name in declared:
        assert hasattr(lib, name), name
    assert _cabi.lib().aat_version() == _cabi.ABI_VERSION == int(re.search(r'#define AAT_B200_VERSION (\d+)', open(os.path.join(ROOT, 'include', 'aat_b200.h')).read()).group(1))


def test_config_struct_layout_matches_header():
    assert ctypes.sizeof(_cabi.AatConfig) == 48
    assert _cabi.AatConfig.min_segment_frames.offset == 24 and _cabi.AatConfig.max_amplitude_for_minima.offset == 40


def test_capacity_and_frame_helpers():
    from aat_b200.context import make_config

    cfg = make_config(12, 2000, 24000, 400, 160, 64, 16000, 15)
    lib = _cabi.lib()
    assert lib.aat_num_mel_frames(ctypes.byref(cfg), 160000) == 1001
    assert lib.aat_num_mel_frames(ctypes.byref(cfg), 100) == 1
    assert lib.aat_segment_capacity(ctypes.byref(cfg), 256000) >= 256000 // 2000 + 256000 // 24000 + 2
    assert synth.mel_frames(28_800_000) == 180001


def test_compute_fails_loudly_without_cuda():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    t = AdaptiveAudioAmplitudeTokenizer()
    with pytest.raises(_cabi.AatError) as e:
        t.get_melspec(np.zeros(1000))
    assert e.value.status == _cabi.AAT_ERR_CUDA


def test_unsupported_fft_length_is_reported():
    from aat_b200.context import make_config

    cfg = make_config(12, 2000, 24000, 512, 160, 64, 16000, 15)
    handle = ctypes.c_void_p()
    win = np.zeros(512)
    fb = np.zeros((257, 64))
    status = _cabi.lib().aat_create(0, ctypes.byref(cfg), win.ctypes.data, fb.ctypes.data, ctypes.byref(handle))
    assert status == _cabi.AAT_ERR_UNSUPPORTED
    assert b"n_fft" in _cabi.lib().aat_last_error()


def test_hubert_frame_formula_and_offsets():
    assert synth.hubert_frames([2000, 24000, 160000, 256000, 320000, 28_800_000, 399, 0]).tolist() == \
        [6, 74, 499, 799, 999, 89999, 0, 0]
    off = synth.segment_frame_offsets([2000, 24000, 100])
    assert off.tolist() == [0, 6, 80, 80]


def test_whole_utterance_frame_offsets_match_the_oracle_loop():
    """SURVEY.md §8d convention (ii): host twin of aat_utterance_frame_csr against the oracle's loop, on the oracle
    port's segmentations of real (synthetic) audio and on the edge lengths of the golden set."""
    from oracle import ref_port, restate

    ref = ref_port.RefTokenizer()
    cases = [(n, ref.segment_lengths(synth.bursty_speech(n, 900 + i))[0])
             for i, n in enumerate((100, 399, 400, 719, 720, 1999, 2000, 2080, 24000, 24001, 25999, 50000, 256000))]
    cases += [(32000, [24000, 8000]), (25000, [23000, 2000]), (100, [2000]), (0, [])]
    for n, lengths in cases:
        off = synth.utterance_frame_offsets(lengths, n)
        assert off.dtype == np.int64 and off.shape == (len(lengths) + 1,)
        assert np.array_equal(off, restate.utterance_frame_offsets_loop(lengths, n)), (n, lengths)
        rows = int(synth.hubert_frames(n))
        assert off[0] == 0 and off[-1] == rows and np.all(np.diff(off) >= 0)  # every encoder row in exactly one segment
    # known answers: 16 s, cuts at 2 s and 9.99 s; a padded tail behind the last encoder row is an empty segment
    assert synth.utterance_frame_offsets([32000, 127840, 96160], 256000).tolist() == [0, 100, 499, 799]
    assert synth.utterance_frame_offsets([2000, 2000], 2080).tolist() == [0, 6, 6]
    with pytest.raises(ValueError):
        synth.utterance_frame_offsets([2000], 4000)


def test_sharding_balances_samples():
    rng = np.random.default_rng(0)
    n = rng.integers(16000, 480000, size=257)
    for world in (1, 2, 4, 8):
        shards = dist.shard_by_duration(n, world)
        assert sorted(np.concatenate(shards).tolist()) == list(range(257))
        loads = np.array([n[s].sum() for s in shards])
        assert loads.max() - loads.min() <= n.max()
    assert [dist.shard_range(10, r, 4) for r in range(4)] == [(0, 2), (2, 5), (5, 7), (7, 10)]


def test_synth_is_deterministic():
    a = synth.bursty_speech(16000, 5)
    b = synth.bursty_speech(16000, 5)
    assert a.dtype == np.float32 and np.array_equal(a, b) and not np.array_equal(a, synth.bursty_speech(16000, 6))


def test_pipeline_defaults_and_pool_flag_values():
    """Host-side constants of the pipelined schedule: six batches in flight whatever the shape, and the flag bits of
    ``aat_pool_flags`` as the header declares them (the pipeline passes AAT_POOL_SHARE_SMS when depth > 1)."""
    import re

    from aat_b200 import _cabi
    from aat_b200.pipeline import TokenizerPipeline

    assert TokenizerPipeline.default_depth([256000] * 64) == 6
    assert TokenizerPipeline.default_depth([28800000] * 8) == 6
    assert TokenizerPipeline.default_depth([]) == 6
    header = open(os.path.join(ROOT, "include", "aat_b200.h")).read()
    for name in ("AAT_POOL_ACCUMULATE", "AAT_POOL_EMB_READY", "AAT_POOL_ROWS_FROM_DEVICE", "AAT_POOL_SHARE_SMS"):
        m = re.search(name + r"\s*=\s*(\d+)", header)
        assert m and int(m.group(1)) == getattr(_cabi, name), name

