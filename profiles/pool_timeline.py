"""Where the pool kernel's time goes at config-2 size: per-CTA global-timer stamps from a trace build of the
library (-DAAT_POOL_TRACE, never the product build).

    make -C audio-adaptive-tokenizer_b200/csrc trace      # profiles/_build/libaat_b200_trace.so
    python profiles/pool_timeline.py > profiles/rN_pool_timeline.txt
"""
import ctypes
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-adaptive-tokenizer_b200")]
import numpy as np
import torch

from aat_b200 import _cabi

_cabi.LIB_PATH = os.path.join(ROOT, "profiles", "_build", "libaat_b200_trace.so")
from aat_b200.context import default_context  # noqa: E402
from aat_b200.pooling import _pool_device  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    ctx = default_context(0)
    lib = _cabi.lib()
    rng = np.random.default_rng(0)
    n_rows, dim = 49559, 768
    lens = rng.integers(6, 75, size=n_rows // 6 + 2)
    off = np.concatenate([[0], np.cumsum(lens)])
    off = off[off < n_rows]
    off = np.concatenate([off, [n_rows]]).astype(np.int64)
    S = off.size - 1
    d_off = torch.from_numpy(off).to(dev)
    embs = [torch.randn(n_rows, dim, device=dev) for _ in range(4)]
    out = torch.empty(S, dim, device=dev)
    colsum = torch.zeros(dim + 1, dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    for variant, cs in (("without colsum", None), ("with colsum", colsum)):
        for i in range(8):
            _pool_device(ctx, embs[i % 4], d_off, S, None, out, cs, stream)
        torch.cuda.synchronize()
        runs = []
        for i in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _pool_device(ctx, embs[i % 4], d_off, S, None, out, cs, stream)
            e1.record()
            torch.cuda.synchronize()
            G = 296
            buf = (ctypes.c_ulonglong * (8 * G))()
            assert lib.aat_debug_pool_trace(buf, G) == 0
            t = np.frombuffer(buf, dtype=np.uint64).reshape(G, 8).astype(np.int64)
            runs.append((e0.elapsed_time(e1) * 1e3, t))
        ev = np.median([r[0] for r in runs])
        # per-run statistics relative to the earliest CTA entry, in microseconds
        rel = []
        for _, t in runs[5:]:
            t0 = t[:, 0].min()
            rel.append((t[:, :6] - t0) / 1e3)
        rel = np.stack(rel)  # [runs, G, 6]
        names = ["CTA entry", "first segment located", "first stage landed", "last stage consumed", "carry collected", "exit"]
        print(f"# pool {n_rows} x {dim}, {S} segments, {variant}: CUDA-event duration median {ev:.2f} us; 15 runs, 296 CTAs")
        print(f"{'stamp (us after the first CTA entry)':40s} {'min':>8s} {'median':>8s} {'p95':>8s} {'max':>8s}")
        for k, nm in enumerate(names):
            v = rel[:, :, k]
            print(f"{nm:40s} {np.median(v.min(1)):8.2f} {np.median(np.median(v, 1)):8.2f} "
                  f"{np.median(np.percentile(v, 95, 1)):8.2f} {np.median(v.max(1)):8.2f}")
        stream_t = rel[:, :, 3] - rel[:, :, 2]
        print(f"{'streaming time per CTA (landed->consumed)':40s} {np.median(stream_t.min(1)):8.2f} {np.median(np.median(stream_t, 1)):8.2f} "
              f"{np.median(np.percentile(stream_t, 95, 1)):8.2f} {np.median(stream_t.max(1)):8.2f}")
        nb = n_rows * dim * 4
        print(f"# bytes / median streaming time = {nb / np.median(np.median(stream_t, 1)) / 1e3:.0f} GB/s; "
              f"bytes / (max exit) = {nb / np.median(rel[:, :, 5].max(1)) / 1e3:.0f} GB/s")
        sm = runs[-1][1][:, 6]
        print(f"# distinct SMs used: {len(set(sm.tolist()))}")
        print()


if __name__ == "__main__":
    main()
