"""Batches in flight on several streams: the offline tokenization job as a software pipeline.

Utterance batches are independent, and the three kernels of a step stress different parts of the
chip: the log-mel kernel is bound by the FP64 / shared-memory pipes and leaves HBM idle, the boundary scan is one
latency-bound CTA per utterance (8 to 256 of the 148 x 3 CTA slots) and the pool kernel is bound by HBM and leaves the
arithmetic pipes idle.  Run strictly one after the other (one plan, one stream) every kernel's weak side is exposed; with
``depth`` plans on ``depth`` streams the boundary scan and the pool of batch *i* overlap the log-mel of batch *i + 1*.

Every slot owns a plan (hence its own device-side scheduling state and pool scratch, see include/aat_b200.h), its own
output buffer and its own running column sums, so slots share nothing and the results are exactly what the serial
loop gives (tests/test_gpu_parity.py::test_pipelined_steps_match_serial_steps).
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

from . import _cabi
from .pooling import DatasetMean
from .tokenizer import AdaptiveAudioAmplitudeTokenizer, PackedBatch


class _Slot:
    def __init__(self, torch, tokenizer, n_samples, dim, device, priority):
        self.batch: PackedBatch = tokenizer.plan(n_samples, device=device)
        dev = self.batch.device
        self.stream = torch.cuda.Stream(device=dev, priority=priority)
        self.handle = ctypes.c_void_p(self.stream.cuda_stream)  # passed straight to the C ABI: no current-stream switch
        self.out = torch.empty(self.batch.total_seg_slots, dim, dtype=torch.float32, device=dev)
        self.mean = DatasetMean(dim, device=dev.index)
        self.done = torch.cuda.Event()
        self.stats = None  # [B, 2] waveform statistics of the fused z-score (allocated on first use)
        self.graphs = {}   # (input buffers, options) -> [times seen, captured CUDA graph of the step or None]


class TokenizerPipeline:
    """``depth`` batches of one shape in flight (default: :meth:`default_depth`).

    >>> pipe = TokenizerPipeline(tok, [256000] * 64, dim=768)
    >>> for wave, emb in batches:          # packed CUDA tensors (PackedBatch layout)
    ...     slot = pipe.submit(wave, emb)  # log-mel -> boundaries -> pool (+ column sums) on the slot's stream
    >>> mean = pipe.dataset_mean()         # joins the streams, adds the slots' sums, allreduce, finalise

    ``submit`` returns the slot; ``slot.done`` is recorded behind its kernels, ``slot.batch`` holds the segment
    tables and ``slot.out[:n_seg]`` the pooled vectors — valid until the slot is submitted to again (``depth`` submits
    later), so consume or copy them after ``slot.done.synchronize()`` / ``stream.wait_event(slot.done)``."""

    @staticmethod
    def default_depth(n_samples: Sequence[int], sampling_rate: int = 16000) -> int:
        """Batches in flight when the caller does not say: 6.  The hardware runs the log-mel kernels of a round back to
        back and then the round's pool kernels (profiles/r2_step_timeline.txt); with the pools at one CTA per SM
        (AAT_POOL_SHARE_SMS) longer rounds keep gaining up to five or six batches (measured on a B200,
        profiles/r2_pipeline_ab.txt: 64 x 16 s 0.1450 ms at 4, 0.1415 at 5, 0.1418 at 6, 0.1421 at 8; 256 x 20 s
        0.7164 at 4, 0.7092 at 6; 8 x 30 min 1.983 at 4, 1.889 at 6, 1.917 at 8)."""
        return 6

    def __init__(self, tokenizer: AdaptiveAudioAmplitudeTokenizer, n_samples: Sequence[int], dim: int,
                 depth: Optional[int] = None, device=None, priorities: Optional[Sequence[int]] = None,
                 fused_amp: bool = True, graphs: bool = False, share_sms: Optional[bool] = None):
        import torch

        if depth is None:
            depth = self.default_depth(n_samples, int(tokenizer.sampling_rate))
        if depth < 1:
            raise ValueError("depth must be >= 1")
        self.torch = torch
        self.dim = int(dim)
        pr = list(priorities) if priorities is not None else [0] * depth
        self.slots: List[_Slot] = [_Slot(torch, tokenizer, n_samples, self.dim, device, pr[k]) for k in range(depth)]
        self.device = self.slots[0].batch.device
        # Several batches in flight: the pool kernels launch one CTA per SM (AAT_POOL_SHARE_SMS).  In steady state the
        # hardware runs the `depth` log-mel kernels of a round back to back and then the `depth` pool kernels
        # (profiles/r2_step_timeline.txt); with half grids two of those pools, or a pool and the first CTAs of the next
        # log-mel, share the SMs, and the ramp and tail of every pool launch are covered: -2.2 % (config 2), -2.2 %
        # (config 3), -2.8 % (config 4) per step, although the pool kernel alone is 8 % slower that way.  The flag
        # moves the CTA tile borders and with them the order of the additions: means may differ in the last bit from
        # share_sms=False (deterministic either way).
        self.share_sms = depth > 1 if share_sms is None else bool(share_sms)
        self.submitted = 0
        # True: the log-mel kernel also emits the amplitude curve the boundary scan starts from (its fused epilogue);
        # False: a separate, fully parallel pass over the mel (aat_amplitude) does, between the two kernels.  Bit-identical
        # results (tested).  With several batches in flight the separate pass hides behind the next batch's log-mel,
        # while the epilogue costs the log-mel kernel — the one kernel nothing hides — a CTA barrier and a serial
        # 64-term chain per tile.
        self.fused_amp = bool(fused_amp)
        # A step is four kernel launches: 32 us of host time on a fast host, 80-104 us on the slower hosts of this pool,
        # against 150 us of device time at 64 x 16 s.  With ``graphs=True``, when a slot is handed the same input buffers
        # again (a loader that rotates over a few device buffers, as any double-buffered loader does) the step is captured
        # into a CUDA graph on its second appearance and replayed from then on: 13 us of host time per step.  Off by
        # default because the replayed step takes 3-5 % more DEVICE time (0.1556 vs 0.1507 ms at 64 x 16 s, 2.08 vs 1.99 ms
        # on 30-min streams: the overlap of consecutive kernels by programmatic dependent launch does not survive the
        # capture); switch it on where the host cannot keep the streams fed.
        self.graphs = bool(graphs)
        self._reduced = False  # dataset_mean() has folded (and allreduced) the sums: reset_sums() before the next pass

    def fork(self):
        """Make every slot's stream wait for the caller's current stream (inputs produced there, an event recorded
        there).  ``submit(..., inputs_ready=True)`` then needs no per-step dependency."""
        cur = self.torch.cuda.current_stream(self.device)
        for slot in self.slots:
            slot.stream.wait_stream(cur)

    def submit(self, wave, emb, colsum: bool = True, znorm: bool = False, rows_from_device: bool = False,
               inputs_ready: bool = False, record_done: bool = True) -> _Slot:
        """One step over one batch on the next slot's stream.  ``wave`` / ``emb`` must be ready on the CALLER's current
        stream: the slot's stream is made to wait for it, unless ``inputs_ready`` says they have been for long (resident
        inputs, or after :meth:`fork`) — an event wait between two steps of a slot costs the overlap of the second
        step's first kernel with the first step's last.  ``znorm`` applies the call sites' z-score inside the log-mel
        kernel; ``rows_from_device`` as in :meth:`PackedBatch.pool`.  ``record_done=False`` skips the ``slot.done`` event for
        callers that only consume the dataset mean (or order consumers with ``stream.wait_stream(slot.stream)``)."""
        torch = self.torch
        if colsum and self._reduced:
            raise RuntimeError("the running sums were folded and allreduced by dataset_mean(); call reset_sums() before "
                               "accumulating another pass (adding to them would count the other ranks twice)")
        slot = self.slots[self.submitted % len(self.slots)]
        self.submitted += 1
        if not inputs_ready:
            slot.stream.wait_stream(torch.cuda.current_stream(self.device))
        if self.graphs:
            key = (wave.data_ptr(), emb.data_ptr(), int(emb.shape[0]), bool(colsum), bool(znorm), bool(rows_from_device))
            entry = slot.graphs.get(key)
            if entry is None:
                if len(slot.graphs) >= 8:  # forget the oldest pattern: the cache is for a handful of rotating buffers
                    slot.graphs.pop(next(iter(slot.graphs)))
                entry = slot.graphs[key] = [0, None, 0]
            entry[0] += 1
            if entry[1] is None and entry[0] == 2:  # second appearance (everything lazy is set up by now): capture
                g = torch.cuda.CUDAGraph()
                before = _cabi.launch_count()
                with torch.cuda.graph(g, stream=slot.stream, capture_error_mode="thread_local"):
                    self._enqueue(slot, wave, emb, colsum, znorm, rows_from_device, None)  # None: the capturing stream
                entry[1], entry[2] = g, _cabi.launch_count() - before  # kernels of the library in the graph
                _cabi.lib().aat_kernel_launch_count_add(-entry[2])      # captured, not run: every replay counts them
            if entry[1] is not None:
                with torch.cuda.stream(slot.stream):
                    entry[1].replay()
                _cabi.lib().aat_kernel_launch_count_add(entry[2])
                if record_done:
                    slot.done.record(slot.stream)
                return slot
        self._enqueue(slot, wave, emb, colsum, znorm, rows_from_device, slot.handle)
        if record_done:
            slot.done.record(slot.stream)
        return slot

    def _enqueue(self, slot, wave, emb, colsum, znorm, rows_from_device, st):
        """The launches of one step on stream handle ``st`` (None = torch's current stream, used while capturing)."""
        b = slot.batch
        if self.fused_amp:  # the whole step in one foreign call
            b.step(wave, emb, slot.out, colsum=slot.mean.running_buffer() if colsum else None, accumulate=colsum,
                   znorm=znorm, rows_from_device=rows_from_device, share_sms=self.share_sms, stream=st)
        else:
            if znorm:
                slot.stats = b.waveform_stats(wave, out=slot.stats, stream=st)
                b.logmel(wave, with_amp=False, znorm_stats=slot.stats, stream=st)
            else:
                b.logmel(wave, with_amp=False, stream=st)
            b.amplitude(stream=st)
            b.boundaries(stream=st)
            # the launch in front of the pool kernel is this slot's boundary scan, which does not write embeddings
            b.pool(emb, slot.out, colsum=slot.mean.running_buffer() if colsum else None, accumulate=colsum,
                   emb_ready=not rows_from_device, rows_from_device=rows_from_device, share_sms=self.share_sms, stream=st)

    def join(self):
        """Make the caller's current stream wait for everything submitted so far (no host synchronisation)."""
        cur = self.torch.cuda.current_stream(self.device)
        for slot in self.slots:
            cur.wait_stream(slot.stream)

    def reset_sums(self):
        self.join()
        for slot in self.slots:
            slot.mean.acc.zero_()
        self._reduced = False
        self.fork()  # later submits see the zeroed sums

    def dataset_mean(self, group=None) -> DatasetMean:
        """Joins the streams and returns slot 0's :class:`DatasetMean` holding the sums of ALL slots, allreduced
        (one SUM allreduce of ``dim + 1`` float64).  Call ``.result()`` on it for the mean vector."""
        self.join()
        total = self.slots[0].mean
        for slot in self.slots[1:]:
            total.acc += slot.mean.acc
            slot.mean.acc.zero_()
        total.allreduce(group)
        self._reduced = True
        self.fork()  # later submits are ordered behind the additions above
        return total
