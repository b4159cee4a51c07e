"""Golden vectors of the training collator's tensors, from the LIVE reference (run in the build container only).

    python tests/golden/make_collate_golden.py

Pins SURVEY.md §8f rows N1 (ragged -> padded scatter), N2 (waveform normalisations) and the n-word cropping of N4 to
outputs of ``aat.training.collate.TokenizedAudioWaveformCollator`` itself — the UNMODIFIED module from
/root/reference/src — instead of to a port.  What has to be worked around to import and call it offline, and how:

* ``aat.training.collate`` imports ``aat.training.trainer`` -> ``aslm.modeling_aslm`` for two enums.  Those modules
  import third-party names that are absent here and play no part in the collator's arithmetic
  (``efficientnet_pytorch``; ``transformers.trainer.ALL_LAYERNORM_LAYERS``, gone in transformers 5): placeholder
  modules / a placeholder symbol are put in ``sys.modules`` before the import.  ``statsmodels`` gets the same empty
  stub as in make_golden.py.
* ``__init__`` downloads ``facebook/hubert-large-ls960-ft``'s processor and lists a data directory, so the object is
  created with ``object.__new__`` and given exactly the attributes ``__init__`` sets
  (ref:src/aat/training/collate.py:62-90).  The audio processor is ``Wav2Vec2FeatureExtractor(feature_size=1,
  sampling_rate=16000, padding_value=0.0, do_normalize=True, return_attention_mask=True)`` — the values of that
  checkpoint's preprocessor_config.json (the hub is unreachable offline; ``__call__`` only uses the feature extractor).
* The text tokenizer is a whitespace stand-in (the text columns are not part of the rows being pinned).
* ``random.randint`` is wrapped to RECORD the collator's draws (n_words, word_start_idx), so the tests can feed the
  same draws to the GPU path.

Writes tests/golden/collate_v1.npz (+ collate_v1.json: recipes, draws, library versions).  Waveforms are regenerated
from their recipes (aat_b200.synth) by the tests; nothing reads /root/reference at test time.
"""
from __future__ import annotations

import hashlib
import json
import os
import random
import sys
import tempfile
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"


class _Placeholder(types.ModuleType):
    """A module whose every attribute is an empty class (for absent third-party imports the collator never calls)."""

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        cls = type(name, (), {})
        setattr(self, name, cls)
        return cls


def _placeholder_module(name):
    parts = name.split(".")
    for i in range(1, len(parts) + 1):
        n = ".".join(parts[:i])
        if n not in sys.modules:
            m = _Placeholder(n)
            m.__path__ = []
            sys.modules[n] = m


def import_reference_collator():
    stub = tempfile.mkdtemp(prefix="aat_stub_")
    os.makedirs(os.path.join(stub, "statsmodels"))
    for name in ("__init__.py", "api.py"):
        open(os.path.join(stub, "statsmodels", name), "w").close()
    sys.path[:0] = [stub, REF_SRC, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
    import torch.nn as nn
    import transformers.trainer as tt

    if not hasattr(tt, "ALL_LAYERNORM_LAYERS"):
        tt.ALL_LAYERNORM_LAYERS = [nn.LayerNorm]
    stubbed = []
    for _ in range(20):
        try:
            import aat.training.collate as collate  # noqa: E402

            break
        except ModuleNotFoundError as e:  # an absent third-party package somewhere in the trainer/model import chain
            if e.name is None or e.name.split(".")[0] in ("aat", "aslm"):
                raise
            _placeholder_module(e.name)
            stubbed.append(e.name)
    else:
        raise RuntimeError("could not import aat.training.collate")
    assert os.path.realpath(collate.__file__).startswith(REF_SRC), collate.__file__
    return collate, stubbed


class WhitespaceTokenizer:
    """Stand-in for the LM tokenizer: the text columns are outside the rows being pinned."""
    bos_token_id, eos_token_id, pad_token_id = 1, 2, 0

    def decode(self, i):
        return {1: "<s>", 2: "</s>"}.get(i, "<unk>")

    def __call__(self, texts, padding=True):
        ids = [[3 + (hash(w) % 1000) for w in t.split()] or [0] for t in texts]
        m = max(len(x) for x in ids)
        return {"input_ids": [x + [0] * (m - len(x)) for x in ids], "attention_mask": [[1] * len(x) + [0] * (m - len(x)) for x in ids]}


def make_collator(collate, ref_tok, audio_encoder_type, segmentation, n_words=None, uniform_frames=None):
    from transformers import Wav2Vec2FeatureExtractor

    c = object.__new__(collate.TokenizedAudioWaveformCollator)
    c.train_config = types.SimpleNamespace(add_prefix=False, sampling_rate=16000)
    c.segmentation = segmentation
    c.audio_encoder_type = audio_encoder_type
    c.uniform_segmentation_frames_per_segment = uniform_frames
    c.n_words = n_words
    c.max_segment_waveform_frames = ref_tok.max_segment_frames
    c.sampling_rate = ref_tok.sampling_rate
    c.noise_augmentation = False
    c.audio_tokenizer = ref_tok
    c.tokenizer = WhitespaceTokenizer()
    c.audio_processor = Wav2Vec2FeatureExtractor(feature_size=1, sampling_rate=16000, padding_value=0.0,
                                                 do_normalize=True, return_attention_mask=True)
    c.melspec_base_path = tempfile.mkdtemp(prefix="aat_mels_")
    c.melspec_files = set()
    return c


def make_items(recipes):
    """recipes: list of (n_samples, seed, dc_offset).  float64 arrays as HF `datasets` delivers them; words every 0.4 s."""
    from aat_b200 import synth

    items = []
    for k, (n, seed, dc) in enumerate(recipes):
        wave = synth.bursty_speech(n, seed).astype(np.float64) + dc
        dur = n / 16000.0
        n_words = max(1, int(dur / 0.4))
        starts = [round(0.05 + 0.4 * i, 3) for i in range(n_words)]
        ends = [round(min(s + 0.33, dur - 0.01), 3) for s in starts]
        items.append({"audio": {"array": wave, "sampling_rate": 16000}, "id": f"utt{k}",
                      "words": [f"w{i}" for i in range(n_words)], "word_start": starts, "word_end": ends})
    return items


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    collate, stubbed = import_reference_collator()
    import scipy
    import torch
    import transformers
    from aat.tokenizer import AdaptiveAudioAmplitudeTokenizer
    from aat.training.trainer import AudioEncoderType
    from aslm.configuration_aslm import SegmentationType

    ref_tok = AdaptiveAudioAmplitudeTokenizer()
    out = {}
    manifest = {"versions": {"numpy": np.__version__, "scipy": scipy.__version__, "torch": torch.__version__,
                             "transformers": transformers.__version__},
                "placeholder_modules": stubbed, "batches": {}}

    draws = []
    real_randint = random.randint

    def recording_randint(a, b):
        v = real_randint(a, b)
        draws.append([a, b, v])
        return v

    batches = {
        # name: (recipes, encoder type, segmentation, n_words, uniform frames, expect_error)
        "adaptive_hubert": ([(48000, 910, 0.0), (80000, 911, 0.02), (32000, 912, 0.0), (68800, 913, -0.5)],
                            AudioEncoderType.hubert.value, SegmentationType.adaptive, None, None, False),
        "adaptive_efficient_net": ([(64000, 920, 0.0), (2080, 921, 0.0), (96000, 922, 0.1)],
                                   AudioEncoderType.efficient_net.value, SegmentationType.adaptive, None, None, False),
        "uniform_hubert": ([(40000, 930, 0.0), (56000, 931, 0.0), (16000, 932, 0.3)],
                           AudioEncoderType.hubert.value, SegmentationType.uniform, None, 16000, False),
        "uniform_efficient_net": ([(40000, 930, 0.0), (50001, 933, 0.0)],
                                  AudioEncoderType.efficient_net.value, SegmentationType.uniform, None, 12000, False),
        "adaptive_nwords": ([(160000, 940, 0.0), (128000, 941, 0.0), (96000, 942, 0.05)],
                            AudioEncoderType.hubert.value, SegmentationType.adaptive, 8, None, False),
        # a segment longer than the tile: the reference fails on the shape mismatch at collate.py:333
        "uniform_too_long": ([(64000, 950, 0.0), (30000, 951, 0.0)],
                             AudioEncoderType.hubert.value, SegmentationType.uniform, None, 30000, True),
    }
    for name, (recipes, enc, seg, n_words, uniform, expect_error) in batches.items():
        items = make_items(recipes)
        c = make_collator(collate, ref_tok, enc, seg, n_words=n_words, uniform_frames=uniform)
        info = {"recipes": recipes, "audio_encoder_type": enc, "segmentation": seg.value, "n_words": n_words,
                "uniform_frames": uniform, "words_per_item": [len(it["words"]) for it in items]}
        draws.clear()
        random.seed(1234)
        random.randint = recording_randint
        try:
            # the intermediate lists first (same seed => same draws as the full call below)
            inter = c._initial_process_segments(items, is_validation=False)
            info["draws_initial"] = [list(d) for d in draws]
            draws.clear()
            random.seed(1234)
            try:
                result = c(items, is_validation=False)
                err = None
            except Exception as e:  # noqa: BLE001 - the error type is the fixture
                result, err = None, type(e).__name__
        finally:
            random.randint = real_randint
        info["draws"] = [list(d) for d in draws]
        assert info["draws"] == info["draws_initial"]
        info["raises"] = err
        assert (err is not None) == expect_error, (name, err)
        for i, (sb, wf, mel) in enumerate(zip(inter["segments_boarders"], inter["audio_segments_waveforms"], inter["items_melspecs"])):
            out[f"{name}/item{i}/segments_boarders"] = np.asarray(sb, dtype=np.int64)
            out[f"{name}/item{i}/melspec"] = np.asarray(mel, dtype=np.float32)
            info.setdefault("waveform_lengths", []).append(int(wf.shape[-1]))
            info.setdefault("waveform_sha256", []).append(sha(np.asarray(wf, dtype=np.float64)))
        out[f"{name}/segments_max_frame_len"] = np.asarray(inter["segments_max_frame_len"], dtype=np.int64)
        if seg == SegmentationType.adaptive:
            # the un-cropped segment lengths (what the cropping starts from), by the reference's own tokenize
            from aat.audio import AudioWaveform

            for i, it in enumerate(items):
                w = it["audio"]["array"]
                normed = (w - w.mean()) / (w.std() + 1e-6)
                segs, _ = ref_tok.tokenize(AudioWaveform(normed, 16000))
                out[f"{name}/item{i}/segment_lengths_full"] = np.asarray([x.waveform.shape[-1] for x in segs], dtype=np.int64)
        if result is not None:
            # what the feature extractor made of the (cropped) waveforms: row N2's wav2vec2 normalisation
            proc = c.audio_processor(inter["audio_segments_waveforms"], padding=True, return_tensors="pt", sampling_rate=16000)
            out[f"{name}/input_values"] = proc.input_values.numpy()
            out[f"{name}/input_attention_mask"] = proc.attention_mask.numpy().astype(np.int64)
            for key in ("segments_boarders_padded", "segments_boarders_attention_mask", "segments_max_frame_len",
                        "batched_segments", "segments_waveforms_mask", "batched_segments_melspectrograms"):
                v = result[key]
                if v is None:
                    info.setdefault("none", []).append(key)
                    continue
                a = v.numpy()
                out[f"{name}/{key}"] = a.astype(np.int64) if a.dtype.kind == "i" else a
            info["segments_count"] = int(result["segments_count"])
        manifest["batches"][name] = info

    np.savez_compressed(os.path.join(HERE, "collate_v1.npz"), **out)
    with open(os.path.join(HERE, "collate_v1.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    size = os.path.getsize(os.path.join(HERE, "collate_v1.npz"))
    print(f"wrote {len(out)} arrays, {size / 1e6:.2f} MB; placeholder modules: {stubbed}")
    for name, info in manifest["batches"].items():
        print(name, "raises" if info["raises"] else "ok", info.get("segments_count"), info["draws"])


if __name__ == "__main__":
    main()
