"""Device time of the log-mel kernel at config-2 shapes for one build of the library (A/B of kernel experiments):

    AAT_B200_LIB=profiles/_build/libaat_exp_XY.so python profiles/logmel_ab.py

Prints one line: the library, us per launch (CUDA-graph replay of 50 launches back to back, best of 3) with float32 input,
with the fused z-score, and a checksum of the mel + amplitude bytes (equal across builds that are bit-identical)."""
import hashlib
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
import torch

from aat_b200 import AdaptiveAudioAmplitudeTokenizer, synth


def timed(fn, reps=50):
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        for _ in range(3):
            fn()
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        with torch.cuda.graph(g, stream=side):
            for _ in range(reps):
                fn()
    g.replay()
    torch.cuda.synchronize()
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g.replay()
        e1.record()
        torch.cuda.synchronize()
        t = e0.elapsed_time(e1) / reps * 1e3
        best = t if best is None else min(best, t)
    return best


def main():
    tok = AdaptiveAudioAmplitudeTokenizer()
    B, N = 64, 256000
    waves = [synth.bursty_speech(N, synth.seed_for(2, i)) for i in range(B)]
    batch = tok.plan([N] * B)
    packed = batch.pack([torch.from_numpy(w) for w in waves])
    stats = torch.empty(B, 2, dtype=torch.float64, device="cuda")
    batch.waveform_stats(packed, out=stats)
    batch.logmel(packed)
    torch.cuda.synchronize()
    h = hashlib.sha256(batch.mel.cpu().numpy().tobytes() + batch.amp.cpu().numpy().tobytes()).hexdigest()[:16]
    t_plain = timed(lambda: batch.logmel(packed))
    t_fused = timed(lambda: batch.logmel(packed, znorm_stats=stats))
    print(f"{os.path.basename(os.environ.get('AAT_B200_LIB', 'libaat_b200.so')):24s} plain {t_plain:7.2f} us   "
          f"fused z-score {t_fused:7.2f} us   sha256(mel+amp) {h}", flush=True)


if __name__ == "__main__":
    main()
