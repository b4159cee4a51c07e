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
    out_cap = torch.empty(4 * S, dim, device=dev)
    d_off_cap = torch.zeros(4 * S + 1, dtype=torch.int64, device=dev)
    d_off_cap[:S + 1] = d_off
    colsum = torch.zeros(dim + 1, dtype=torch.float64, device=dev)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    n_seg_dev = torch.tensor([S], dtype=torch.int64, device=dev)
    for variant, cs, nsd, cap in (("without colsum", None, None, S), ("with colsum", colsum, None, S),
                                  ("with colsum, S read from the device (bench.py's call)", colsum, n_seg_dev, 4 * S)):
        for i in range(8):
            _pool_device(ctx, embs[i % 4], d_off_cap if nsd is not None else d_off, cap, nsd, out_cap if nsd is not None else out, cs, stream)
        torch.cuda.synchronize()
        runs = []
        for i in range(20):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _pool_device(ctx, embs[i % 4], d_off_cap if nsd is not None else d_off, cap, nsd, out_cap if nsd is not None else out, cs, stream)
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
        dump = os.environ.get("AAT_TIMELINE_DUMP")
        if dump:
            np.savez(f"{dump}_{len(variant)}.npz", rel=rel, sm=runs[-1][1][:, 6])
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
        # is the spread a property of the SM (both CTAs of an SM agree, stable across runs) or of the rows (CTA index)?
        done = rel[:, :, 3]                      # [runs, G] last stage consumed
        by_cta = done.mean(0)
        print(f"# spread of 'last stage consumed': std over CTAs of the run-mean = {by_cta.std():.2f} us; "
              f"mean over CTAs of the run-std = {done.std(0).mean():.2f} us (small => systematic per CTA)")
        per_sm = {}
        for c in range(G):
            per_sm.setdefault(int(sm[c]), []).append(by_cta[c])
        pairs = np.array([v for v in per_sm.values() if len(v) == 2])
        if len(pairs):
            print(f"# the two CTAs of an SM: mean |difference| = {np.abs(pairs[:, 0] - pairs[:, 1]).mean():.2f} us; "
                  f"correlation = {np.corrcoef(pairs[:, 0], pairs[:, 1])[0, 1]:.2f}")
        order = np.argsort(by_cta)
        print("# slowest 12 CTAs (index, smid, us):", [(int(c), int(sm[c]), round(float(by_cta[c]), 1)) for c in order[-12:]])
        print("# fastest 12 CTAs (index, smid, us):", [(int(c), int(sm[c]), round(float(by_cta[c]), 1)) for c in order[:12]])
        q = [by_cta[i * G // 8:(i + 1) * G // 8].mean() for i in range(8)]
        print("# mean by octile of the CTA index (position in the row range):", [round(float(x), 1) for x in q])
        smo = np.argsort(sm)
        q = [by_cta[smo][i * G // 8:(i + 1) * G // 8].mean() for i in range(8)]
        print("# mean by octile of the SM id:", [round(float(x), 1) for x in q])
        print()


if __name__ == "__main__":
    main()
