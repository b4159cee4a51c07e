"""Bandwidth of the N1/N2 byte movers on a B200 (CUDA events, config-2 shapes).
    python profiles/collate_bw.py > profiles/rN_collate_bw.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
import numpy as np
import torch

from aat_b200 import AdaptiveAudioAmplitudeTokenizer, collate, synth


def timed(fn, reps=20):
    """us per call on the DEVICE: the call is captured into a CUDA graph and the graph is replayed, so that the host
    side of the call (torch.empty, ctypes, two or three launches: 20-30 us, more than some of these kernels take) is
    not what is measured."""
    if os.environ.get("AAT_NO_GRAPH"):  # under ncu: plain launches
        for _ in range(2):
            fn()
        torch.cuda.synchronize()
        return float("nan")
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            fn()
    for _ in range(3):
        g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3


def main():
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    tok = AdaptiveAudioAmplitudeTokenizer()
    B, N = 64, 256000
    waves = [synth.bursty_speech(N, synth.seed_for(2, i)) for i in range(B)]
    batch = tok.plan([N] * B)
    packed = batch.pack([torch.from_numpy(w) for w in waves])
    batch.logmel(packed), batch.boundaries()
    torch.cuda.synchronize()
    s_max = int(batch.seg_count.max().item())
    padded, mask = collate.pad_segment_boarders(batch, s_max)
    n_max = int(padded.max().item())
    wave_padded = torch.zeros(B, n_max, device="cuda")
    wave_padded[:, :N] = packed.view(B, N)
    F = tok.max_segment_frames
    print(f"# config 2 shapes: B={B}, N={N}, S_max={s_max}, max_segment_frames={F}; peak {peak} GB/s (measured copy)")
    print(f"{'kernel':34s} {'algorithmic MB':>15s} {'us':>9s} {'GB/s':>8s} {'frac':>6s}")
    rows = []
    t = timed(lambda: collate.normalize_waveforms(batch, packed, "w2v2"))
    rows.append(("normalize w2v2 f32->f32 (3 kernels)", B * N * 4 * 3, t))  # stats read + apply read + write
    packed64 = packed.double()
    t = timed(lambda: collate.normalize_waveforms(batch, packed64, "zscore"))
    rows.append(("normalize zscore f64->f64 (3 kernels)", B * N * 8 * 3, t))
    t = timed(lambda: collate.normalize_waveforms_padded(batch, packed, "w2v2", n_max=n_max))
    rows.append(("normalize_padded w2v2 + mask", B * N * 4 * 2 + B * n_max * 8, t))
    stats = torch.empty(B, 2, dtype=torch.float64, device="cuda")
    t = timed(lambda: batch.waveform_stats(packed, out=stats))
    rows.append(("waveform_stats f32 (2 kernels)", B * N * 4, t))
    t_plain = timed(lambda: batch.logmel(packed))
    t_fused = timed(lambda: batch.logmel(packed, znorm_stats=stats))
    t_sep = timed(lambda: batch.logmel(collate.normalize_waveforms(batch, packed, "zscore")))
    t = timed(lambda: collate.scatter_segments(batch, wave_padded, padded, F, check=False))
    rows.append(("scatter_segments + mask", B * s_max * F * 4 * 2 + B * N * 4, t))
    t = timed(lambda: collate.scatter_mel_segments(batch, padded, F, check=False))
    rows.append(("scatter_mel_segments", B * s_max * 64 * (1 + F // 160) * 4 + batch.mel.numel() * 4, t))
    for name, nbytes, us in rows:
        print(f"{name:34s} {nbytes / 1e6:15.1f} {us:9.1f} {nbytes / us / 1e3:8.0f} {nbytes / us / 1e3 / peak:6.3f}")
    print("# device time per call (each call captured into a CUDA graph and replayed 20 times)")
    print(f"# z-scored log-mel of the batch: log-mel alone {t_plain:.1f} us; statistics + fused z-score "
          f"{timed(lambda: batch.logmel(packed, znorm_stats=batch.waveform_stats(packed, out=stats))):.1f} us "
          f"(log-mel kernel with the fused z-score {t_fused:.1f} us); separate normalise (float64 copy) + log-mel {t_sep:.1f} us")


if __name__ == "__main__":
    main()
