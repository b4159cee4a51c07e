"""World-size-2 check of the N>1 path on CPU (gloo): utterance sharding + the single allreduce of
(dim + 1) float64 column sums reproduce the single-process dataset mean.  The per-rank column sums
come from the oracle here (no GPU); on the GPU box the same flow runs with K4's epilogue and NCCL."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from aat_b200 import dist as adist
from aat_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _utterance(i, dim):
    rng = np.random.default_rng(100 + i)
    n_seg = int(rng.integers(3, 9))
    lengths = rng.integers(2000, 24000, size=n_seg)
    off = synth.segment_frame_offsets(lengths)
    emb = rng.standard_normal((int(off[-1]), dim), dtype=np.float32)
    return emb, off


def _pooled_sums(indices, dim):
    from oracle import c_oracle

    acc = np.zeros(dim + 1, dtype=np.float64)
    for i in indices:
        emb, off = _utterance(int(i), dim)
        pooled = c_oracle.mean_pool_f32(emb, off)
        acc[:dim] += pooled.astype(np.float64).sum(axis=0)
        acc[dim] += pooled.shape[0]
    return acc


def _worker(rank, world, port, n_utts, dim, out_path):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    n_samples = [16000 * (1 + (i % 5)) for i in range(n_utts)]
    mine = adist.shard_by_duration(n_samples, world)[rank]
    # the product's own accumulator object: DatasetMean.allreduce is what bench.py and the offline job call.  Its
    # device-side pieces (the pool kernel's epilogue, the finalise kernel) need a GPU and are covered by the -m gpu
    # tests; here the running totals are host tensors filled from the oracle, and the collective runs over gloo.
    from aat_b200.pooling import DatasetMean

    dm = DatasetMean.__new__(DatasetMean)
    dm.dim = dim
    dm.acc = torch.from_numpy(_pooled_sums(mine, dim))
    acc = dm.allreduce()
    assert acc is dm.acc
    if rank == 0:
        np.save(out_path, acc.numpy())
    dist.destroy_process_group()


def test_two_rank_dataset_mean_matches_single_process(tmp_path, c_oracle):
    n_utts, dim = 11, 32
    out = str(tmp_path / "acc.npy")
    mp.spawn(_worker, args=(2, _free_port(), n_utts, dim, out), nprocs=2, join=True)
    got = np.load(out)
    want = _pooled_sums(range(n_utts), dim)
    assert got[dim] == want[dim]
    np.testing.assert_allclose(got[:dim] / got[dim], want[:dim] / want[dim], rtol=1e-12, atol=1e-12)
