"""Pool-kernel sweeps on a B200 (K4 alone):
  * size sweep: duration vs. algorithmic bytes, to separate the fixed cost per launch from the streaming rate, by two
    clocks: an event pair around every launch, and ONE event pair around 32 back-to-back launches on rotating inputs;
  * ragged stress at config-3 size (255 744 x 1024): all-min, all-max, alternating, uniform-random,
    Zipf-heavy and one giant segment, each against the uniform-random mix.
python profiles/pool_sweep.py > profiles/rN_pool_sweep.txt"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
import numpy as np
import torch

from aat_b200 import _cabi
from aat_b200.context import default_context
from aat_b200.pooling import _pool_device
import ctypes


def time_pool(ctx, embs, d_off, S, out, stream, reps=32, plan=None):
    """(us per launch by per-launch event pairs, us per launch by one pair around `reps` back-to-back launches)"""
    for i in range(5):
        _pool_device(ctx, embs[i % len(embs)], d_off, S, None, out, None, stream, plan=plan)
    torch.cuda.synchronize()
    _cabi.profile_enable(ctx.handle, ("pool",))
    for i in range(reps):
        _pool_device(ctx, embs[i % len(embs)], d_off, S, None, out, None, stream, plan=plan)
    torch.cuda.synchronize()
    n, ms = _cabi.profile_summary(ctx.handle)["pool"]
    _cabi.profile_enable(ctx.handle, ())
    best = None
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reps):
            _pool_device(ctx, embs[i % len(embs)], d_off, S, None, out, None, stream, plan=plan)
        e1.record()
        torch.cuda.synchronize()
        us = e0.elapsed_time(e1) * 1e3 / reps
        best = us if best is None else min(best, us)
    return ms / n * 1e3, best


def ragged_mixes(n_rows, rng):
    def from_lens(lens):
        off = np.concatenate([[0], np.cumsum(lens)])
        off = off[off < n_rows]
        return np.concatenate([off, [n_rows]]).astype(np.int64)

    k = n_rows // 6 + 2
    return {
        "uniform 6..74": from_lens(rng.integers(6, 75, size=k)),
        "all-min (6)": from_lens(np.full(k, 6)),
        "all-max (74)": from_lens(np.full(k, 74)),
        "alternating 6/74": from_lens(np.tile([6, 74], k // 2 + 1)),
        "Zipf-heavy": from_lens(np.minimum(6 * rng.zipf(1.6, size=k), n_rows // 3)),
        "one giant": np.asarray([0, n_rows], dtype=np.int64),
        "first half min, second half max": from_lens(np.concatenate([np.full(n_rows // 12, 6), np.full(k, 74)])),
    }


PLAN = None  # a plan only lends its pool scratch (launches without one are serialised by an event per launch)


def main():
    global PLAN
    from aat_b200 import AdaptiveAudioAmplitudeTokenizer

    dev = torch.device("cuda", 0)
    ctx = default_context(0)
    PLAN = AdaptiveAudioAmplitudeTokenizer(device=0).plan([16000])
    assert PLAN.ctx.handle.value == ctx.handle.value
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rng = np.random.default_rng(0)
    print(f"# pool_kernel<float,1,false>: 4 rotating inputs, peak = {peak} GB/s; 'us' = event pair per launch, 'b2b' = one "
          f"event pair around 32 back-to-back launches")
    print(f"{'rows':>9s} {'dim':>5s} {'segments':>9s} {'MB':>9s} {'us':>9s} {'GB/s':>9s} {'frac':>6s} {'b2b us':>9s} {'b2b frac':>8s}")
    # floor of the timing method: a trivial torch fill kernel between two CUDA events
    x = torch.empty(1024, device=dev)
    e0 = [torch.cuda.Event(enable_timing=True) for _ in range(50)]
    e1 = [torch.cuda.Event(enable_timing=True) for _ in range(50)]
    for i in range(50):
        e0[i].record(); x.fill_(1.0); e1[i].record()
    torch.cuda.synchronize()
    print(f"# event-timed trivial kernel (torch fill of 4 KB): {np.median([a.elapsed_time(b) for a, b in zip(e0, e1)]) * 1e3:.2f} us")
    rows_list = [(296, 768), (2960, 768), (6000, 768), (12000, 768), (25000, 768), (49559, 768), (100000, 768), (200000, 768), (400000, 768),
                 (719992, 768), (255744, 1024), (1000000, 1024)]
    pts = []
    for n_rows, dim in rows_list:
        lens = rng.integers(6, 75, size=n_rows // 6 + 2)
        off = np.concatenate([[0], np.cumsum(lens)])
        off = off[off < n_rows]
        off = np.concatenate([off, [n_rows]]).astype(np.int64)
        S = off.size - 1
        d_off = torch.from_numpy(off).to(dev)
        embs = [torch.randn(n_rows, dim, device=dev) for _ in range(4 if n_rows * dim * 4 < 1.5e9 else 2)]
        out = torch.empty(S, dim, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        us, b2b = time_pool(ctx, embs, d_off, S, out, stream, plan=PLAN.handle)
        nbytes = n_rows * dim * 4 + S * dim * 4 + (S + 1) * 8
        gbs = nbytes / us / 1e3
        pts.append((nbytes, us, b2b))
        print(f"{n_rows:9d} {dim:5d} {S:9d} {nbytes / 1e6:9.1f} {us:9.2f} {gbs:9.0f} {gbs / peak:6.3f} {b2b:9.2f} {nbytes / b2b / 1e3 / peak:8.3f}")
        del embs, out
    x = np.array([p[0] for p in pts])
    A = np.vstack([np.ones_like(x), x]).T
    for label, col in (("event pair per launch", 1), ("back to back", 2)):
        t0, slope = np.linalg.lstsq(A, np.array([p[col] for p in pts]), rcond=None)[0]
        print(f"# least squares ({label}): us = {t0:.2f} + bytes / {1 / slope / 1e3:.0f} GB/s")

    # ---- ragged stress at config-3 size
    n_rows, dim = 255744, 1024
    embs = [torch.randn(n_rows, dim, device=dev) for _ in range(4)]
    print(f"\n# ragged stress, {n_rows} x {dim} float32 (config 3 size)")
    print(f"{'mix':34s} {'segments':>9s} {'us':>9s} {'b2b us':>9s} {'b2b frac':>8s} {'vs uniform':>10s}")
    base = None
    for name, off in ragged_mixes(n_rows, rng).items():
        S = off.size - 1
        d_off = torch.from_numpy(off).to(dev)
        out = torch.empty(S, dim, device=dev)
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        us, b2b = time_pool(ctx, embs, d_off, S, out, stream, plan=PLAN.handle)
        nbytes = n_rows * dim * 4 + S * dim * 4 + (S + 1) * 8
        base = b2b if base is None else base
        print(f"{name:34s} {S:9d} {us:9.2f} {b2b:9.2f} {nbytes / b2b / 1e3 / peak:8.3f} {b2b / base:10.3f}")


if __name__ == "__main__":
    main()
