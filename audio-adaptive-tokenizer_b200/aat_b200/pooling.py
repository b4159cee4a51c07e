"""Ragged per-segment mean-pool of HuBERT frame embeddings (K4) and the dataset-level mean (C1).

``mean_pool_segments`` replaces the idiom of ref:scripts/mean_hubert_embeddings.py:19-20::

    mean_embeddings = [x.mean(dim=1, keepdim=True).to(torch.float32) for x in embedings_list]
    averaged = torch.cat(mean_embeddings, dim=1)        # [1, S, D]

It accepts that very list (tensors of shape [1, n_i, D]) or the packed form ``(emb [T, D],
seg_off [S+1])`` and returns ``[1, S, D]`` float32.  CUDA tensors stay on the device; host inputs go
through ``aat_host_mean_pool``.  ``DatasetMean`` accumulates the column sums the pool kernel emits
and finishes with one SUM allreduce (NCCL via ``torch.distributed``) — the only collective on the path.
"""
from __future__ import annotations

import ctypes
from typing import Optional, Sequence

import numpy as np

from . import _cabi
from .context import Context, default_context


def _torch_dtype_code(dtype) -> int:
    import torch

    code = {torch.float32: _cabi.AAT_F32, torch.float16: _cabi.AAT_F16, torch.bfloat16: _cabi.AAT_BF16}.get(dtype)
    if code is None:
        raise TypeError(f"embedding dtype {dtype} is not supported (float32, float16, bfloat16)")
    return code


def _pool_device(ctx: Context, emb, seg_off, n_seg: int, n_seg_dev, out, colsum, stream, accumulate: bool = False,
                 plan=None, emb_ready: bool = False, rows_from_device: bool = False, share_sms: bool = False):
    """Raw K4 launch on CUDA tensors (no allocation, no sync).  ``plan`` is an ``aat_plan`` handle (its scratch is
    used, so launches on different plans may overlap) or None (the context's scratch; such launches are serialised)."""
    if emb.dim() != 2 or not emb.is_contiguous():
        raise ValueError("emb must be a contiguous [T, D] tensor")
    if out.dtype.is_floating_point is False or out.element_size() != 4 or not out.is_contiguous():
        raise TypeError("out must be a contiguous float32 tensor")
    flags = ((_cabi.AAT_POOL_ACCUMULATE if accumulate else 0) | (_cabi.AAT_POOL_EMB_READY if emb_ready else 0) |
             (_cabi.AAT_POOL_ROWS_FROM_DEVICE if rows_from_device else 0) |
             (_cabi.AAT_POOL_SHARE_SMS if share_sms else 0))
    _cabi.check(_cabi.lib().aat_segment_mean_pool(
        ctx.handle, plan, emb.data_ptr(), _torch_dtype_code(emb.dtype), int(emb.shape[0]), int(emb.shape[1]),
        seg_off.data_ptr(), int(n_seg), n_seg_dev.data_ptr() if n_seg_dev is not None else None, out.data_ptr(),
        colsum.data_ptr() if colsum is not None else None, flags, stream))
    return out


def mean_pool_segments(embeddings, seg_off=None, *, out=None, colsum=None, device=None):
    """Per-segment mean over frames -> ``[1, S, D]`` float32.

    embeddings : list/tuple of tensors ``[1, n_i, D]`` (the reference's on-disk format), or a packed
                 ``[T, D]`` tensor / ndarray when ``seg_off`` is given
    seg_off    : ``[S+1]`` int64 CSR offsets in frames (packed form only)
    out        : optional preallocated ``[S, D]`` float32 CUDA tensor (packed CUDA form)
    colsum     : optional ``[D+1]`` float64 CUDA tensor receiving column sums of the pooled vectors and S
    """
    import torch

    if seg_off is None:
        parts = list(embeddings)
        if not parts:
            raise RuntimeError("torch.cat(): expected a non-empty list of Tensors")
        for x in parts:
            if x.dim() != 3 or x.shape[0] != 1:
                raise ValueError("each embedding must have shape [1, n_i, D]")
        dim = int(parts[0].shape[2])
        if any(int(x.shape[2]) != dim or x.dtype != parts[0].dtype for x in parts):
            raise RuntimeError("Sizes / dtypes of the per-segment tensors must match except in dimension 1")
        lengths = [int(x.shape[1]) for x in parts]
        if all(not x.is_cuda for x in parts):
            # host list (what torch.load gives for one file of the reference): one C-ABI call, every tensor copied once
            # into pinned staging — no torch.cat on the host first
            parts = [x if x.is_contiguous() else x.contiguous() for x in parts]
            n_seg = len(parts)
            ptrs = (ctypes.c_void_p * n_seg)(*[x.data_ptr() for x in parts])
            rows = (ctypes.c_int64 * n_seg)(*lengths)
            res = torch.empty((1, n_seg, dim), dtype=torch.float32)
            cs = np.empty(dim + 1, dtype=np.float64) if colsum is not None else None
            ctx = default_context(device)
            _cabi.check(_cabi.lib().aat_host_mean_pool_list(ctx.handle, ptrs, rows, n_seg, _torch_dtype_code(parts[0].dtype),
                                                            dim, res.data_ptr(), cs.ctypes.data if cs is not None else None))
            if colsum is not None:
                colsum[...] = torch.from_numpy(cs) if isinstance(colsum, torch.Tensor) else cs
            return res
        off = np.zeros(len(parts) + 1, dtype=np.int64)
        np.cumsum(lengths, out=off[1:])
        packed = torch.cat([x.reshape(x.shape[1], x.shape[2]) for x in parts], dim=0)
        return mean_pool_segments(packed, off, colsum=colsum, device=device)

    if isinstance(embeddings, torch.Tensor) and embeddings.is_cuda:
        emb = embeddings.contiguous()
        ctx = default_context(emb.device.index)
        off = torch.as_tensor(seg_off, dtype=torch.int64).to(emb.device).contiguous()
        n_seg = int(off.numel()) - 1
        if out is None:
            out = torch.empty((n_seg, emb.shape[1]), dtype=torch.float32, device=emb.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(emb.device).cuda_stream)
        _pool_device(ctx, emb, off, n_seg, None, out, colsum, stream)
        return out.view(1, n_seg, emb.shape[1])

    # host buffers: one C-ABI call does H2D, the kernel and D2H
    if isinstance(embeddings, torch.Tensor):
        t = embeddings.contiguous()
        code = _torch_dtype_code(t.dtype)
        ptr, n_rows, dim = t.data_ptr(), int(t.shape[0]), int(t.shape[1])
        keep = t
    else:
        a = np.ascontiguousarray(embeddings)
        code = {np.dtype(np.float32): _cabi.AAT_F32, np.dtype(np.float16): _cabi.AAT_F16}.get(a.dtype)
        if code is None:
            raise TypeError(f"embedding dtype {a.dtype} is not supported")
        ptr, n_rows, dim = a.ctypes.data, int(a.shape[0]), int(a.shape[1])
        keep = a
    off = np.ascontiguousarray(np.asarray(seg_off, dtype=np.int64))
    n_seg = int(off.size) - 1
    res = torch.empty((1, n_seg, dim), dtype=torch.float32)
    cs = np.empty(dim + 1, dtype=np.float64) if colsum is not None else None
    ctx = default_context(device)
    _cabi.check(_cabi.lib().aat_host_mean_pool(ctx.handle, ptr, code, n_rows, dim, off.ctypes.data, n_seg,
                                               res.data_ptr(), cs.ctypes.data if cs is not None else None))
    del keep
    if colsum is not None:
        colsum[...] = torch.from_numpy(cs) if isinstance(colsum, torch.Tensor) else cs
    return res


class DatasetMean:
    """Dataset-level mean HuBERT embedding: unweighted mean over every pooled segment vector
    (SURVEY.md §8a row A9).  Column sums come out of the pool kernel's epilogue in float64; batches are
    accumulated on the device and ranks are combined with ONE allreduce(SUM) of ``dim + 1`` doubles."""

    def __init__(self, dim: int, device=None):
        import torch

        self.dim = int(dim)
        self.device = torch.device("cuda", torch.cuda.current_device() if device is None else device)
        self.ctx = default_context(self.device.index)
        self.acc = torch.zeros(self.dim + 1, dtype=torch.float64, device=self.device)
        self.batch = torch.zeros(self.dim + 1, dtype=torch.float64, device=self.device)

    def colsum_buffer(self):
        """Pass this as ``colsum=`` to the pool call, then call :meth:`accumulate`."""
        return self.batch

    def running_buffer(self):
        """Pass this as ``colsum=`` with ``accumulate=True``: the pool call adds the batch's sums straight into
        the running totals (no separate accumulate kernel)."""
        return self.acc

    def accumulate(self):
        import torch

        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(_cabi.lib().aat_colsum_accumulate(self.ctx.handle, self.acc.data_ptr(), self.batch.data_ptr(),
                                                      self.dim, stream))

    def allreduce(self, group=None):
        """SUM over ranks (NCCL on GPU tensors; a no-op when torch.distributed is not initialised)."""
        import torch.distributed as dist

        if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
            dist.all_reduce(self.acc, op=dist.ReduceOp.SUM, group=group)
        return self.acc

    def result(self):
        import torch

        mean = torch.empty(self.dim, dtype=torch.float32, device=self.device)
        stream = ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _cabi.check(_cabi.lib().aat_colsum_finalize(self.ctx.handle, self.acc.data_ptr(), self.dim, mean.data_ptr(),
                                                    stream))
        return mean

    @property
    def count(self) -> int:
        return int(self.acc[self.dim].item())
