"""Pool-kernel size sweep on a B200: CUDA-event duration of K4 alone vs. algorithmic bytes, to separate the
fixed cost per launch from the streaming rate.   python profiles/pool_sweep.py > profiles/rN_pool_sweep.txt"""
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


def main():
    dev = torch.device("cuda", 0)
    ctx = default_context(0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"] if os.path.exists(
        os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    rng = np.random.default_rng(0)
    print(f"# pool_kernel<float,1,false>: event-timed launches, 4 rotating inputs, peak = {peak} GB/s")
    print(f"{'rows':>9s} {'dim':>5s} {'segments':>9s} {'MB':>9s} {'us':>9s} {'GB/s':>9s} {'frac':>6s}")
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
        for i in range(5):
            _pool_device(ctx, embs[i % len(embs)], d_off, S, None, out, None, stream)
        torch.cuda.synchronize()
        _cabi.profile_enable(ctx.handle, ("pool",))
        reps = 40
        for i in range(reps):
            _pool_device(ctx, embs[i % len(embs)], d_off, S, None, out, None, stream)
        torch.cuda.synchronize()
        n, ms = _cabi.profile_summary(ctx.handle)["pool"]
        _cabi.profile_enable(ctx.handle, ())
        us = ms / n * 1e3
        nbytes = n_rows * dim * 4 + S * dim * 4 + (S + 1) * 8
        gbs = nbytes / us / 1e3
        pts.append((nbytes, us))
        print(f"{n_rows:9d} {dim:5d} {S:9d} {nbytes / 1e6:9.1f} {us:9.2f} {gbs:9.0f} {gbs / peak:6.3f}")
        del embs, out
    x = np.array([p[0] for p in pts]); y = np.array([p[1] for p in pts])
    A = np.vstack([np.ones_like(x), x]).T
    t0, slope = np.linalg.lstsq(A, y, rcond=None)[0]
    print(f"# least squares: us = {t0:.2f} + bytes / {1 / slope / 1e3:.0f} GB/s")


if __name__ == "__main__":
    main()
