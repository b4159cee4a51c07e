"""What the pool kernel streams when it runs on a FEW SMs only (grid capped by AAT_POOL_GRID, experiments build:
`make -C audio-adaptive-tokenizer_b200/csrc experiments` -> profiles/_build/libaat_b200_exp.so; ring variants with
EXTRA_NVCCFLAGS="-DAAT_POOL_STAGES=S -DAAT_POOL_STAGE_KB=KB -DAAT_POOL_CTAS=C", as in profiles/r2_pool_ring_grid.txt):
the question behind "give the HBM-bound pool 8-32 SMs of its own beside the FP64-bound log-mel kernel".

    AAT_B200_LIB=profiles/_build/libaat_b200_exp.so AAT_POOL_GRID=16 python profiles/pool_grid.py"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "audio-adaptive-tokenizer_b200"), os.path.join(ROOT, "profiles")]
import numpy as np
import torch

from aat_b200.context import default_context
from pool_sweep import time_pool


def main():
    ctx = default_context(0)
    rng = np.random.default_rng(0)
    n_rows, dim = 64 * 799, 768  # config 2
    lens = rng.integers(6, 75, size=n_rows // 6 + 2)
    off = np.concatenate([[0], np.cumsum(lens)])
    off = np.concatenate([off[off < n_rows], [n_rows]]).astype(np.int64)
    S = len(off) - 1
    d_off = torch.from_numpy(off).cuda()
    embs = [torch.randn(n_rows, dim, device="cuda") for _ in range(8)]  # 8 x 157 MB > L2
    out = torch.empty(S, dim, device="cuda")
    stream = torch.cuda.current_stream()
    per_launch, b2b = time_pool(ctx, embs, d_off, S, out, stream)
    nbytes = n_rows * dim * 4 + S * dim * 4 + (S + 1) * 8
    g = os.environ.get("AAT_POOL_GRID", "all")
    print(f"grid {g:>4s}: {b2b:8.1f} us back to back  ({nbytes / b2b / 1e3:7.0f} GB/s"
          + (f", {nbytes / b2b / 1e3 / int(g):6.1f} GB/s per CTA" if g != "all" else "") + ")", flush=True)


if __name__ == "__main__":
    main()
