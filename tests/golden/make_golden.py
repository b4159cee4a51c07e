"""Generate golden vectors from the LIVE reference (run in the build container only).

    python tests/golden/make_golden.py

Imports ``aat.tokenizer`` from /root/reference/src (read-only) with an empty
``statsmodels`` stub on sys.path (the reference imports it at
ref:src/aat/tokenizer.py:7 but only uses it in a commented-out line), runs the
reference on deterministic synthetic inputs (``aat_b200.synth``) and writes

    tests/golden/golden_v1.npz      arrays, keyed ``<case>/<name>``
    tests/golden/MANIFEST.json      case list, library versions, sha256 of long arrays

The reference's own tests hold no golden vectors for this path (SURVEY.md §4,
§8c), so these fixtures — outputs of the reference itself — are what pins the
oracle.  /root/reference does not exist on the GPU box; tests only read the
committed files.
"""
from __future__ import annotations

import hashlib
import json
import os
import sys
import tempfile

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"


def _import_reference():
    stub = tempfile.mkdtemp(prefix="aat_stub_")
    os.makedirs(os.path.join(stub, "statsmodels"))
    for name in ("__init__.py", "api.py"):
        open(os.path.join(stub, "statsmodels", name), "w").close()
    sys.path[:0] = [stub, REF_SRC, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
    from aat.audio import AudioWaveform  # noqa: E402
    from aat.tokenizer import AdaptiveAudioAmplitudeTokenizer  # noqa: E402

    return AdaptiveAudioAmplitudeTokenizer, AudioWaveform


def sha(a: np.ndarray) -> str:
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    Tok, AudioWaveform = _import_reference()
    import scipy
    import torch
    import transformers
    from aat_b200 import synth

    out = {}
    manifest = {"versions": {"numpy": np.__version__, "scipy": scipy.__version__, "torch": torch.__version__,
                             "transformers": transformers.__version__}, "cases": {}}
    tok = Tok()

    def intermediates(t, mel):
        amp = -10 * mel.mean(axis=0)
        cs = np.cumsum(amp)
        n = t.running_mean_points
        rm = (cs[n:] - cs[:-n]) / float(n)
        return amp, cs, rm

    def run_case(name, wave, t=tok, keep_mel=True, keep_wave=False, pool_dim=None, pool_seed=None, meta=None):
        mel = t.get_melspec(wave)
        minima = t.find_amplitude_minimas(mel)
        boarders, _ = t.pretokenize(wave, mel)
        segs = t.process_segments_boarders(wave, boarders)
        lengths = np.asarray([s.shape[-1] for s in segs], dtype=np.int64)
        amp, cs, rm = intermediates(t, mel)
        out[f"{name}/minima"] = minima.astype(np.int64)
        out[f"{name}/boarders"] = np.asarray(boarders, dtype=np.int64)
        out[f"{name}/lengths"] = lengths
        if keep_mel:
            out[f"{name}/mel"] = mel
            out[f"{name}/amp"] = amp
            out[f"{name}/cs"] = cs
            out[f"{name}/rm"] = rm
        if keep_wave:
            out[f"{name}/wave"] = wave
        info = {"n_samples": int(wave.shape[0]), "wave_dtype": str(wave.dtype), "mel_shape": list(mel.shape),
                "mel_sha256": sha(mel), "n_minima": int(minima.size), "n_segments": int(lengths.size),
                "min_segment_frames": int(t.min_segment_frames), "max_segment_frames": int(t.max_segment_frames)}
        if len(segs) < 300:
            segs2, _ = t.tokenize(AudioWaveform(wave, 16000), melspec=mel)
            assert [s.waveform.shape[-1] for s in segs2] == lengths.tolist()
        if pool_dim is not None:
            off = synth.segment_frame_offsets(lengths)
            rng = np.random.default_rng(pool_seed)
            emb = rng.standard_normal((int(off[-1]), pool_dim), dtype=np.float32)
            te = torch.from_numpy(emb)
            lst = [te[off[i]:off[i + 1]].unsqueeze(0) for i in range(off.size - 1)]
            # ref:scripts/mean_hubert_embeddings.py:19-20, verbatim idiom
            mean_embeddings = [x.mean(dim=1, keepdim=True).to(torch.float32) for x in lst]
            pooled = torch.cat(mean_embeddings, dim=1)
            out[f"{name}/pooled"] = pooled.numpy()
            out[f"{name}/frame_off"] = off
            info.update({"pool_dim": pool_dim, "pool_seed": pool_seed, "emb_sha256": sha(emb)})
        if meta:
            info.update(meta)
        manifest["cases"][name] = info
        return mel, lengths

    # config 1: 10 s clip, D=768
    run_case("c1_10s", synth.bursty_speech(160000, synth.seed_for(1, 0)), pool_dim=768, pool_seed=11000,
             meta={"gen": "bursty_speech(160000, 1000)"})
    # config 2: two of the 64 x 16 s utterances
    run_case("c2_16s_u0", synth.bursty_speech(256000, synth.seed_for(2, 0)), pool_dim=768, pool_seed=12000,
             meta={"gen": "bursty_speech(256000, 2000)"})
    run_case("c2_16s_u1", synth.bursty_speech(256000, synth.seed_for(2, 1)), keep_mel=False,
             meta={"gen": "bursty_speech(256000, 2001)"})
    # float64 z-normalised input, as ref:scripts/audio_tokenization_melspec.py:40 feeds get_melspec
    w = synth.bursty_speech(256000, synth.seed_for(2, 2)).astype(np.float64)
    run_case("c2_16s_znorm", (w - w.mean()) / (w.std() + 1e-6),
             meta={"gen": "znorm(float64(bursty_speech(256000, 2002)))"})
    # config 3: HuBERT-large width, 20 s
    run_case("c3_20s", synth.bursty_speech(320000, synth.seed_for(3, 0)), keep_mel=False, pool_dim=1024,
             pool_seed=13000, meta={"gen": "bursty_speech(320000, 3000)"})
    # degenerate inputs
    run_case("silence_2s", synth.silence(32000).astype(np.float64), meta={"gen": "zeros(32000) float64"})
    run_case("noise_10s", synth.stationary_noise(160000, 7), keep_mel=False, meta={"gen": "stationary_noise(160000, 7)"})
    # edge lengths
    for i, n in enumerate([100, 201, 1919, 1920, 1999, 2000, 2080, 24000, 24001, 25999, 26000, 48000, 50000]):
        run_case(f"edge_{n}", synth.bursty_speech(n, 9000 + i), keep_mel=n <= 2080, keep_wave=n <= 2080,
                 meta={"gen": f"bursty_speech({n}, {9000 + i})"})
    # min > max configuration really used by the reference (ref:scripts/trainer_train.py:116-122)
    tok_mm = Tok(min_segment_duration_milliseconds=500, max_segment_duration_milliseconds=250)
    run_case("minmax_16s", synth.bursty_speech(256000, synth.seed_for(2, 3)), t=tok_mm, keep_mel=False,
             meta={"gen": "bursty_speech(256000, 2003)", "min_ms": 500, "max_ms": 250})
    # long-form (config 4): one 30-min stream; mel is pinned by sha256 only
    run_case("c4_30min", synth.bursty_speech(28_800_000, synth.seed_for(4, 0)), keep_mel=False,
             meta={"gen": "bursty_speech(28800000, 4000)"})

    # known answers of the merge/split state machine, evaluated by the live reference
    sm_cases = [
        ("sm_25000", tok, 25000, [25000]), ("sm_49000", tok, 49000, [49000]), ("sm_48000", tok, 48000, [48000]),
        ("sm_160000", tok, 160000, [160000]),
        ("sm_merge_a", tok, 10000, [1000, 1500, 5000, 9000, 10000]), ("sm_merge_b", tok, 10000, [5000, 9500, 10000]),
        ("sm_100", tok, 100, [100]), ("sm_1999", tok, 1999, [1999]),
    ]
    tok84 = Tok(min_segment_duration_milliseconds=500, max_segment_duration_milliseconds=250)
    sm_cases += [("sm_mm_20000", tok84, 20000, [20000]), ("sm_mm_12000", tok84, 12000, [3000, 12000]),
                 ("sm_mm_9000", tok84, 9000, [9000]), ("sm_mm_17000", tok84, 17000, [17000])]
    for name, t, n, boarders in sm_cases:
        segs = t.process_segments_boarders(np.zeros(n), boarders)
        out[f"{name}/boarders"] = np.asarray(boarders, dtype=np.int64)
        out[f"{name}/lengths"] = np.asarray([s.shape[-1] for s in segs], dtype=np.int64)
        manifest["cases"][name] = {"n_samples": n, "min_segment_frames": int(t.min_segment_frames),
                                   "max_segment_frames": int(t.max_segment_frames), "state_machine_only": True}

    # constants
    out["const/mel_filters"] = tok.mel_filters
    out["const/window"] = tok.window_fn

    np.savez_compressed(os.path.join(HERE, "golden_v1.npz"), **out)
    with open(os.path.join(HERE, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    size = os.path.getsize(os.path.join(HERE, "golden_v1.npz"))
    print(f"wrote {len(out)} arrays, {size / 1e6:.2f} MB; {len(manifest['cases'])} cases")


if __name__ == "__main__":
    main()
