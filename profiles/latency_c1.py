"""BASELINE config 1 — a single 10 s clip, batch 1, through the reference-compatible numpy API
(tokenize + mean_pool_segments on the list-of-[1, n_i, 768] format), per-call wall-clock latency, with the
oracle port of the reference's CPU path beside it.   python profiles/latency_c1.py > profiles/rN_latency_c1.txt"""
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
import numpy as np
import torch

from aat_b200 import AdaptiveAudioAmplitudeTokenizer, AudioWaveform, mean_pool_segments, synth
from oracle import ref_port


def med(fn, reps):
    ts = []
    for _ in range(reps):
        t0 = time.perf_counter()
        fn()
        ts.append(time.perf_counter() - t0)
    return 1e3 * float(np.median(ts)), 1e3 * float(np.min(ts))


def main():
    wave = synth.bursty_speech(160000, synth.seed_for(1, 0)).astype(np.float64)
    tok, ref = AdaptiveAudioAmplitudeTokenizer(), ref_port.RefTokenizer()
    awf = AudioWaveform(wave, 16000)
    segs, mel = tok.tokenize(awf)
    lengths = [s.waveform.shape[-1] for s in segs]
    frames = synth.hubert_frames(lengths)
    g = torch.Generator().manual_seed(0)
    embs = [torch.randn(1, int(f), 768, generator=g) for f in frames]
    assert lengths == ref.segment_lengths(wave)[0]
    for _ in range(5):
        tok.tokenize(awf), mean_pool_segments(embs)
    rows = [
        ("tokenize (mel + minima + merge/split), numpy in/out", med(lambda: tok.tokenize(awf), 200),
         med(lambda: ref.tokenize(ref_port.AudioWaveform(wave, 16000)), 10)),
        ("get_melspec", med(lambda: tok.get_melspec(wave), 200), med(lambda: ref.get_melspec(wave), 10)),
        ("find_amplitude_minimas", med(lambda: tok.find_amplitude_minimas(mel), 200),
         med(lambda: ref.find_amplitude_minimas(mel), 50)),
        ("mean_pool_segments(list of [1, n_i, 768])", med(lambda: mean_pool_segments(embs), 200),
         med(lambda: ref_port.mean_pool_segments(embs), 50)),
    ]
    print(f"# config 1: one 10 s clip (160000 samples, float64), {len(lengths)} segments, {int(frames.sum())} HuBERT frames")
    print(f"# per-call wall clock, host buffers in and out (H2D + kernels + D2H + sync); torch threads = {torch.get_num_threads()}")
    print(f"{'call':52s} {'B200 median ms':>15s} {'min':>8s} {'CPU port median ms':>19s} {'min':>8s} {'speed-up':>9s}")
    for name, (gm, gmin), (cm, cmin) in rows:
        print(f"{name:52s} {gm:15.3f} {gmin:8.3f} {cm:19.3f} {cmin:8.3f} {cm / gm:9.1f}")
    total_gpu = rows[0][1][0] + rows[3][1][0]
    total_cpu = rows[0][2][0] + rows[3][2][0]
    print(f"# whole path: {total_gpu:.3f} ms vs {total_cpu:.3f} ms -> {10 / 3600 / (total_gpu / 1e3):.2f} vs "
          f"{10 / 3600 / (total_cpu / 1e3):.3f} audio-hours/s at batch 1")


if __name__ == "__main__":
    main()
