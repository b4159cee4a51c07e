"""Where the boundary kernel's time goes at config-2 size (64 x 16 s): global-timer stamps of thread 0 of every CTA
from the trace build (make -C audio-adaptive-tokenizer_b200/csrc trace).   python profiles/bnd_timeline.py"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
import numpy as np
import torch

from aat_b200 import _cabi

_cabi.LIB_PATH = os.path.join(ROOT, "profiles", "_build", "libaat_b200_trace.so")
from aat_b200 import AdaptiveAudioAmplitudeTokenizer, synth  # noqa: E402


def main():
    B, N = 64, 256_000
    tok = AdaptiveAudioAmplitudeTokenizer(device=0)
    lib = _cabi.lib()
    waves = [synth.bursty_speech(N, synth.seed_for(2, i)) for i in range(B)]
    batch = tok.plan([N] * B)
    wave = torch.from_numpy(np.concatenate(waves)).cuda()
    rel = []
    for i in range(12):
        batch.logmel(wave)
        torch.cuda.synchronize()
        batch.boundaries()
        torch.cuda.synchronize()
        buf = (ctypes.c_ulonglong * (32 * B))()
        assert lib.aat_debug_bnd_trace(buf, B) == 0
        t = np.frombuffer(buf, dtype=np.uint64).reshape(B, 32).astype(np.int64)
        if i >= 4:
            rel.append((t - t[:, 0].min()) / 1e3)
    rel = np.stack(rel)
    T = 1 + N // 160
    n_it = None
    print(f"# boundaries kernel, {B} utterances x {T} frames; stamps of thread 0, us after the first CTA entry (median over 8 runs)")
    print(f"{'stamp':38s} {'min':>8s} {'median':>8s} {'max':>8s}")

    def row(name, k):
        v = rel[:, :, k]
        print(f"{name:38s} {np.median(v.min(1)):8.2f} {np.median(np.median(v, 1)):8.2f} {np.median(v.max(1)):8.2f}")

    row("CTA entry", 0)
    row("prologue done (sizes, first prefetch)", 1)
    for it in range(26):
        if rel[0, 0, 2 + it] <= 0:
            break
        row(f"end of pipeline iteration {it}", 2 + it)
    for c in range(8):
        if rel[0, 0, 10 + 2 * c] <= 0:
            break
        v = rel[:, :, 11 + 2 * c] - rel[:, :, 10 + 2 * c]
        st = rel[:, :, 10 + 2 * c]
        print(f"merge/split thread, chunk {c}: starts {np.median(np.median(st, 1)):6.2f} us, takes "
              f"{np.median(v.min(1)):5.2f} / {np.median(np.median(v, 1)):5.2f} / {np.median(v.max(1)):5.2f} us (min / median / max over CTAs)")
    for name, a, b in (("scan thread, chunk 1 (512 frames)", 24, 25), ("worker warps, test of chunk 1", 26, 27),
                       ("worker warps, test of chunk 1 + load of chunk 3", 26, 9)):
        v = rel[:, :, b] - rel[:, :, a]
        print(f"{name}: {np.median(v.min(1)):5.2f} / {np.median(np.median(v, 1)):5.2f} / {np.median(v.max(1)):5.2f} us (min / median / max over CTAs)")
    row("emit warp: pipeline left", 20)
    row("emit warp: last boarder done", 21)
    row("emit warp: tail segment done", 22)
    row("emit warp: totals published", 23)
    row("look-back done", 28)
    row("CSR slice written, ticket taken", 29)


if __name__ == "__main__":
    main()
