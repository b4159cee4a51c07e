"""Timeline of the pipelined step on one B200: which kernel's CTAs occupy the SMs when, in steady state.

    make -C audio-adaptive-tokenizer_b200/csrc timeline     # profiles/_build/libaat_b200_timeline.so (-DAAT_TIMELINE)
    python profiles/step_timeline.py [c2|c3|c4] [depth] > profiles/rN_step_timeline.txt

The timeline build makes thread 0 of every CTA of the three path kernels record {start, end, SM, CTA} (global timer).
The script runs bench.py's own Workload through TokenizerPipeline, reads the rings back and prints, for a few
steady-state steps: every launch's first start / last start / first end / last end, and how much of the SM x time area
of a step each kernel covers."""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
os.environ.setdefault("AAT_B200_LIB", os.path.join(ROOT, "profiles", "_build", "libaat_b200_timeline.so"))
import numpy as np
import torch

import bench
from aat_b200 import AdaptiveAudioAmplitudeTokenizer, _cabi

RING = 1 << 16


def dump(name):
    lib = _cabi.lib()
    fn = getattr(lib, "aat_debug_timeline_" + name)
    fn.restype = ctypes.c_int
    buf = (ctypes.c_ulonglong * (3 * RING))()
    cnt = ctypes.c_uint(0)
    assert fn(buf, ctypes.byref(cnt)) == 0
    n = min(cnt.value, RING)
    a = np.frombuffer(buf, dtype=np.uint64).reshape(RING, 3)[:n].astype(np.int64)
    return a[np.argsort(a[:, 0])]


def launches(rec, grid):
    """records sorted by start time -> list of launches (chunks of `grid` CTAs: all three kernels fit in one wave)"""
    return [rec[i:i + grid] for i in range(0, len(rec) - grid + 1, grid)]


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "c2"
    depth = int(sys.argv[2]) if len(sys.argv) > 2 else 0
    steps = 132
    tok = AdaptiveAudioAmplitudeTokenizer()
    w = bench.Workload(torch, tok, name, 0, 0, 4, depth)
    pipe = w.pipes[w.depth]
    pipe.fork()
    for i in range(steps):
        w.step(i)
    pipe.join()
    torch.cuda.synchronize()
    rec = {k: dump(k) for k in ("logmel", "boundaries", "pool")}
    # grids: what the last launches used (CTA index range)
    grid = {k: int((rec[k][:, 2] & 0xffffffff).max()) + 1 for k in rec}
    # keep the records of the timed loop only: the last `steps` launches of every kernel
    L = {k: launches(rec[k][-steps * grid[k]:], grid[k]) for k in rec}
    first = steps - 36  # well inside the steady state, well before the drain at the end of the loop
    t0 = L["logmel"][first][:, 0].min()
    print(f"# {name}, {w.depth} batches in flight, grids {grid}; times in us relative to the first CTA start of log-mel launch {first}")
    print("# launch: first start, last start | first end, last end   (duration of the launch = last end - first start)")
    events = []
    for k in L:
        for j, r in enumerate(L[k]):
            events.append((r[:, 0].min(), k, j, r))
    events.sort(key=lambda e: e[0])
    starts = np.array([r[:, 0].min() for r in L["logmel"]])
    t1 = starts[min(first + 2 * w.depth, steps - 1)]  # two rounds of the pipeline
    for begin, k, j, r in events:
        if t0 <= begin < t1:
            us = lambda t: (t - t0) / 1e3
            sms = len(set((r[:, 2] >> 32).tolist()))
            print(f"{k:10s} #{j:3d}  start {us(r[:, 0].min()):8.1f} .. {us(r[:, 0].max()):8.1f} | end {us(r[:, 1].min()):8.1f} .. {us(r[:, 1].max()):8.1f}"
                  f"   ({(r[:, 1].max() - r[:, 0].min()) / 1e3:6.1f} us, CTA lifetime median {np.median(r[:, 1] - r[:, 0]) / 1e3:6.1f} us, on {sms} SMs)")
    # steady-state step = distance between consecutive log-mel launches
    period = np.diff(starts)[steps // 2:-4].mean() / 1e3
    print(f"# steady-state period (first start of a log-mel launch to the next): {period:.1f} us")
    # share of the SM x time area covered by CTAs of each kernel within one period (a CTA slot counts 1/slots-per-SM)
    a, b = starts[first], starts[first + 1]
    slots = {"logmel": 3, "pool": max(1, grid["pool"] // 148), "boundaries": 1}
    for k in rec:
        r = rec[k]
        ov = np.clip(np.minimum(r[:, 1], b) - np.maximum(r[:, 0], a), 0, None).sum()
        print(f"# {k:10s}: CTA-time inside one period = {ov / 1e3:9.1f} us = {ov / ((b - a) * 148 * slots[k]):5.2f} of the period x 148 SMs x {slots[k]} CTA slot(s)")
    # how many log-mel CTAs are resident while pool CTAs are: overlap of the two kernels
    lm, pl = rec["logmel"], rec["pool"]
    pl = pl[(pl[:, 1] > a) & (pl[:, 0] < b)]
    if len(pl):
        pa, pb = pl[:, 0].min(), pl[:, 1].max()
        ov = np.clip(np.minimum(lm[:, 1], pb) - np.maximum(lm[:, 0], pa), 0, None).sum()
        print(f"# while pool CTAs of that period are alive ({(pb - pa) / 1e3:.1f} us), log-mel CTAs resident on average: {ov / (pb - pa):.0f} of 444")


if __name__ == "__main__":
    main()
