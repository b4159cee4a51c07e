"""``AdaptiveAudioAmplitudeTokenizer`` — the reference's interface, backed by the B200 kernels.

Mirrors ref:src/aat/tokenizer.py:14-200 method for method (same names, argument meaning, return
types and assertion behaviour).  The numpy entry points go through the ``aat_host_*`` C-ABI calls
(host buffers in, host buffers out, copies inside the library); the batched entry points
(:class:`PackedBatch`) take CUDA tensors and never touch the host.  Nothing here computes on the
CPU: without the native library and a CUDA device every compute call raises.
"""
from __future__ import annotations

import ctypes
from typing import List, Optional, Sequence

import numpy as np

from . import _cabi
from .audio import AudioWaveform
from .constants import hann_window_periodic, mel_filter_bank_slaney
from .context import Context, get_context, make_config


def _wave_arg(audio_waveform):
    """Host waveform -> (contiguous array, AAT dtype).  The reference promotes the padded waveform to
    float64 (TF:audio_utils.py:774); float32/float64 are passed as they are, every other real dtype is
    promoted here exactly as ``astype(np.float64)`` would."""
    wf = np.asarray(audio_waveform)
    if wf.ndim != 1:
        raise ValueError(f"Input waveform must have only one dimension, shape is {wf.shape}")
    if np.iscomplexobj(wf):
        raise ValueError("Complex-valued input waveforms are not currently supported")
    if wf.dtype == np.float32:
        return np.ascontiguousarray(wf), _cabi.AAT_F32
    return np.ascontiguousarray(wf, dtype=np.float64), _cabi.AAT_F64


class AdaptiveAudioAmplitudeTokenizer:
    def __init__(self,
                 running_mean_points=12,
                 min_segment_duration_milliseconds=125,
                 max_segment_duration_milliseconds=1500,
                 n_fft=400,
                 hop_length=160,
                 num_mel_filters=64,
                 sampling_rate=16000,
                 max_amplitude_for_minima=15,
                 device: Optional[int] = None):
        self.running_mean_points = running_mean_points
        self.max_amplitude_for_minima = max_amplitude_for_minima

        self.n_fft = n_fft
        self.hop_length = hop_length
        self.num_mel_filters = num_mel_filters
        self.sampling_rate = sampling_rate

        self.min_segment_duration_milliseconds = min_segment_duration_milliseconds
        self.max_segment_duration_milliseconds = max_segment_duration_milliseconds

        self.min_segment_frames = self.milliseconds_to_frames(min_segment_duration_milliseconds)
        self.max_segment_frames = self.milliseconds_to_frames(max_segment_duration_milliseconds)

        self.mel_filters = mel_filter_bank_slaney(
            num_frequency_bins=1 + self.n_fft // 2,
            num_mel_filters=num_mel_filters,
            min_frequency=0.0,
            max_frequency=8000.0,
            sampling_rate=sampling_rate,
        )
        self.window_fn = hann_window_periodic(self.n_fft)

        self.device = device

    # ------------------------------------------------------------------ native context
    def __getstate__(self):
        state = dict(self.__dict__)  # plain attributes only: the native context lives in a per-process cache
        state.pop("_ctx_memo", None)
        return state

    def _ctx(self) -> Context:
        """The native context of this tokenizer's configuration on its device, in this process.  Looked up in the
        per-process cache once and remembered on the instance (the cache key hashes 100 KB of constant tables, which is
        half the latency of a single-clip call); a forked or unpickled copy looks it up again.  Like the reference's
        object, the tokenizer is immutable after construction: change an attribute and build a new one."""
        import os

        memo = self.__dict__.get("_ctx_memo")
        if memo is not None and memo[0] == os.getpid() and memo[1] == self.device and memo[2].handle is not None:
            return memo[2]
        cfg = make_config(self.running_mean_points, self.min_segment_frames, self.max_segment_frames, self.n_fft,
                          self.hop_length, self.num_mel_filters, self.sampling_rate, self.max_amplitude_for_minima)
        ctx = get_context(cfg, self.window_fn, self.mel_filters, self.device)
        self._ctx_memo = (os.getpid(), self.device, ctx)
        return ctx

    def _segment_capacity(self, n_samples: int) -> int:
        cfg = self._ctx().config
        return int(_cabi.lib().aat_segment_capacity(ctypes.byref(cfg), int(n_samples)))

    # ------------------------------------------------------------------ reference interface
    def find_amplitude_minimas(self, melspec: np.ndarray):
        """ref:src/aat/tokenizer.py:55-92 — ``melspec`` (num_mel_filters, seq_len) float32 -> int64 minima."""
        mel = np.asarray(melspec)
        if mel.ndim != 2 or mel.shape[0] != self.num_mel_filters:
            raise ValueError(f"melspec must have shape ({self.num_mel_filters}, seq_len), got {mel.shape}")
        if mel.dtype != np.float32:
            raise TypeError("find_amplitude_minimas is implemented for float32 mel spectrograms (what get_melspec "
                            f"returns); numpy would compute in {mel.dtype}")
        if mel.shape[1] == 0:
            raise ValueError("attempt to get argmax of an empty sequence")
        mel = np.ascontiguousarray(mel)
        n_frames = mel.shape[1]
        minima = np.empty(n_frames, dtype=np.int64)
        count = ctypes.c_int64(0)
        _cabi.check(_cabi.lib().aat_host_find_minimas(self._ctx().handle, mel.ctypes.data, n_frames,
                                                      minima.ctypes.data, ctypes.byref(count)))
        return minima[: count.value].copy()

    def milliseconds_to_frames(self, milliseconds: int) -> int:
        return int(milliseconds * self.sampling_rate / 1000)

    def left_pad_waveform_with_zeros(self, waveform):
        waveform_padded = np.zeros([self.min_segment_frames])
        waveform_padded[-waveform.shape[-1]:] = waveform
        return waveform_padded

    def right_pad_waveform_with_zeros(self, waveform):
        waveform_padded = np.zeros([self.min_segment_frames])
        waveform_padded[:waveform.shape[-1]] = waveform
        return waveform_padded

    def get_melspec(self, audio_waveform: np.ndarray) -> np.ndarray:
        """ref:src/aat/tokenizer.py:107-119 — (n,) waveform -> (num_mel_filters, 1 + n // hop) float32."""
        wf, dt = _wave_arg(audio_waveform)
        if wf.shape[0] == 0:
            raise ValueError("can't extend empty axis 0 using modes other than 'constant' or 'empty'")
        n_frames = 1 + wf.shape[0] // self.hop_length
        mel = np.empty((self.num_mel_filters, n_frames), dtype=np.float32)
        _cabi.check(_cabi.lib().aat_host_logmel(self._ctx().handle, wf.ctypes.data, dt, wf.shape[0], mel.ctypes.data))
        return mel

    def _host_tokenize(self, audio_waveform, melspec, want_mel: bool):
        """One C-ABI call: (mel?) -> minima -> segments for one utterance."""
        n = int(audio_waveform.shape[-1])
        n_frames = 1 + n // self.hop_length
        cap = self._segment_capacity(n)
        starts = np.empty(cap, dtype=np.int64)
        lengths = np.empty(cap, dtype=np.int64)
        minima = np.empty(n_frames, dtype=np.int64)
        n_seg, n_min, tail = ctypes.c_int64(0), ctypes.c_int64(0), ctypes.c_int32(0)
        mel_out = None
        if melspec is None:
            wf, dt = _wave_arg(audio_waveform)
            if n == 0:
                raise ValueError("can't extend empty axis 0 using modes other than 'constant' or 'empty'")
            mel_out = np.empty((self.num_mel_filters, n_frames), dtype=np.float32) if want_mel else None
            _cabi.check(_cabi.lib().aat_host_tokenize(
                self._ctx().handle, wf.ctypes.data, dt, n, None, mel_out.ctypes.data if want_mel else None,
                minima.ctypes.data, ctypes.byref(n_min), starts.ctypes.data, lengths.ctypes.data, cap,
                ctypes.byref(n_seg), ctypes.byref(tail)))
        else:
            mel = np.asarray(melspec)
            if mel.dtype != np.float32 or mel.ndim != 2 or mel.shape[0] != self.num_mel_filters:
                raise TypeError(f"melspec must be a float32 array of shape ({self.num_mel_filters}, seq_len)")
            if mel.shape[1] != n_frames:
                # a cached mel of another length: boundaries follow the mel, the waveform only supplies N
                return self._host_tokenize_mismatched(audio_waveform, mel)
            mel = np.ascontiguousarray(mel)
            _cabi.check(_cabi.lib().aat_host_tokenize(
                self._ctx().handle, None, _cabi.AAT_F32, n, mel.ctypes.data, None, minima.ctypes.data,
                ctypes.byref(n_min), starts.ctypes.data, lengths.ctypes.data, cap, ctypes.byref(n_seg),
                ctypes.byref(tail)))
        return (minima[: n_min.value].copy(), starts[: n_seg.value].copy(), lengths[: n_seg.value].copy(),
                bool(tail.value), mel_out)

    def _host_tokenize_mismatched(self, audio_waveform, mel):
        minima = self.find_amplitude_minimas(mel)
        boarders = (minima * self.hop_length).tolist() + [audio_waveform.shape[-1]]
        starts, lengths, tail = self._process_boarders(int(audio_waveform.shape[-1]), boarders)
        return minima, starts, lengths, tail, None

    def pretokenize(self, audio_waveform: np.ndarray, melspec=None):
        """ref:src/aat/tokenizer.py:121-139 — returns ``(segments_boarders: List[int], melspec)``."""
        if melspec is None:
            minima, _, _, _, melspec = self._host_tokenize(audio_waveform, None, want_mel=True)
        else:
            minima = self.find_amplitude_minimas(melspec)
        item_waveform_minimas = minima * self.hop_length  # move to waveform space
        segments_boarders = item_waveform_minimas.tolist() + [audio_waveform.shape[-1]]
        return segments_boarders, melspec

    def _process_boarders(self, n_samples: int, segments_boarders):
        boarders = np.ascontiguousarray(np.asarray(segments_boarders, dtype=np.int64).reshape(-1))
        reach = max(int(boarders.max()) if boarders.size else 0, n_samples)
        cap = int(boarders.size + reach // self.max_segment_frames + 4)
        starts = np.empty(cap, dtype=np.int64)
        lengths = np.empty(cap, dtype=np.int64)
        n_seg, tail = ctypes.c_int64(0), ctypes.c_int32(0)
        status = _cabi.lib().aat_host_process_boarders(
            self._ctx().handle, n_samples, boarders.ctypes.data, boarders.size, starts.ctypes.data,
            lengths.ctypes.data, cap, ctypes.byref(n_seg), ctypes.byref(tail))
        if status == _cabi.AAT_ERR_TAIL:
            # the reference fails inside right_pad_waveform_with_zeros (ref:src/aat/tokenizer.py:102-105)
            raise ValueError("could not broadcast input array from shape "
                             f"({n_samples - self._last_accepted(boarders)},) into shape ({self.min_segment_frames},)")
        _cabi.check(status)
        return starts[: n_seg.value].copy(), lengths[: n_seg.value].copy(), bool(tail.value)

    def _last_accepted(self, boarders) -> int:
        prev = 0
        for b in boarders.tolist():
            if b - prev >= self.min_segment_frames:
                prev = b
        return prev

    def _materialize(self, audio_waveform, starts, lengths, padded_tail) -> List[np.ndarray]:
        """(start, length) pairs -> the arrays the reference returns: views of the caller's waveform,
        except the zero-padded float64 tail."""
        segments: List[np.ndarray] = []
        last = len(starts) - 1
        for i, (s, l) in enumerate(zip(starts.tolist(), lengths.tolist())):
            if padded_tail and i == last:
                segments.append(self.right_pad_waveform_with_zeros(audio_waveform[s:]))
            else:
                segments.append(audio_waveform[s:s + l])
        return segments

    def process_segments_boarders(self, audio_waveform: np.ndarray, segments_boarders) -> List[np.ndarray]:
        """ref:src/aat/tokenizer.py:141-183 — merge too small segments and split too big segments."""
        starts, lengths, tail = self._process_boarders(int(audio_waveform.shape[-1]), segments_boarders)
        return self._materialize(audio_waveform, starts, lengths, tail)

    def tokenize(self, audio_waveform_sr: AudioWaveform, melspec=None):
        """ref:src/aat/tokenizer.py:185-200 — returns ``(List[AudioWaveform], melspec)``."""
        audio_waveform_sr.assert_sampling_rate(self.sampling_rate)
        audio_waveform = audio_waveform_sr.waveform

        given = melspec
        _, starts, lengths, tail, mel_out = self._host_tokenize(audio_waveform, melspec, want_mel=True)
        melspec = given if given is not None else mel_out
        waveform_segments = self._materialize(audio_waveform, starts, lengths, tail)

        assert len(waveform_segments) < 300
        sum_frames = sum(x.shape[-1] for x in waveform_segments)
        assert sum_frames >= audio_waveform.shape[-1]

        audio_segments_sr = [AudioWaveform(wf, audio_waveform_sr.sampling_rate) for wf in waveform_segments]
        return audio_segments_sr, melspec

    def segment_lengths(self, audio_waveform: np.ndarray, melspec=None):
        """Segment lengths only (what every caller of the reference actually consumes,
        ref:scripts/audio_tokenization.py:37, ref:src/aat/training/collate.py:156-158); no 300-segment
        assertion, so it also serves long-form audio."""
        _, _, lengths, _, _ = self._host_tokenize(audio_waveform, melspec, want_mel=False)
        return lengths

    # ------------------------------------------------------------------ batched device path
    def plan(self, n_samples: Sequence[int], device=None) -> "PackedBatch":
        """Device-resident layout + output buffers for a batch of utterances with these lengths."""
        return PackedBatch(self, n_samples, device)


class PackedBatch:
    """A batch of utterances in the packed layout of ``include/aat_b200.h`` with preallocated outputs.

    All methods enqueue kernels on the current torch CUDA stream and return device tensors; nothing
    synchronises and nothing is allocated per call, so a sequence of calls can be captured into a
    CUDA graph (``torch.cuda.graph``).
    """

    def __init__(self, tokenizer: AdaptiveAudioAmplitudeTokenizer, n_samples: Sequence[int], device=None):
        import torch

        self.tokenizer = tokenizer
        if device is not None:
            tokenizer = _with_device(tokenizer, device)
        self.ctx = tokenizer._ctx()
        self.device = torch.device("cuda", self.ctx.device)
        lib = _cabi.lib()
        ns = np.ascontiguousarray(np.asarray(n_samples, dtype=np.int64).reshape(-1))
        self.n_samples = ns
        self.n_utts = int(ns.size)
        handle = ctypes.c_void_p()
        _cabi.check(lib.aat_plan_create(self.ctx.handle, self.n_utts, ns.ctypes.data, ctypes.byref(handle)))
        self.handle = handle
        self.total_samples = int(lib.aat_plan_total_samples(handle))
        self.total_frames = int(lib.aat_plan_total_frames(handle))
        self.total_seg_slots = int(lib.aat_plan_total_seg_slots(handle))
        self.wave_off = np.empty(self.n_utts + 1, dtype=np.int64)
        self.frame_off = np.empty(self.n_utts + 1, dtype=np.int64)
        self.seg_slot_off = np.empty(self.n_utts + 1, dtype=np.int64)
        _cabi.check(lib.aat_plan_offsets(handle, self.wave_off.ctypes.data, self.frame_off.ctypes.data,
                                         self.seg_slot_off.ctypes.data))
        self.n_mels = int(tokenizer.num_mel_filters)
        dev = self.device
        self.mel = torch.empty(self.n_mels * self.total_frames, dtype=torch.float32, device=dev)
        self.amp = torch.empty(self.total_frames, dtype=torch.float32, device=dev)
        self.seg_start = torch.zeros(self.total_seg_slots, dtype=torch.int64, device=dev)
        self.seg_len = torch.zeros(self.total_seg_slots, dtype=torch.int64, device=dev)
        self.seg_count = torch.zeros(self.n_utts, dtype=torch.int32, device=dev)
        self.status = torch.zeros(self.n_utts, dtype=torch.int32, device=dev)
        self.minima = torch.zeros(self.total_frames, dtype=torch.int64, device=dev)
        self.minima_count = torch.zeros(self.n_utts, dtype=torch.int32, device=dev)
        self.seg_off = torch.zeros(self.total_seg_slots + 1, dtype=torch.int64, device=dev)
        self._csr_totals = torch.zeros(2, dtype=torch.int64, device=dev)  # {segments, HuBERT frames} of the batch
        self.n_seg = self._csr_totals[:1]
        self.n_frames = self._csr_totals[1:]
        self.utt_seg_off = torch.zeros(self.n_utts + 1, dtype=torch.int64, device=dev)

    def close(self):
        if getattr(self, "handle", None):
            _cabi.lib().aat_plan_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self, stream=None):
        """``stream`` as a ``cudaStream_t`` handle: a ``torch.cuda.Stream``, a raw handle (``c_void_p`` / int), or — None —
        the caller's current stream ON THE PLAN'S DEVICE (which need not be the current device).  Passing the stream
        explicitly saves the current-stream switch (``with torch.cuda.stream(...)``) in loops that drive several streams."""
        if stream is None:
            import torch

            return ctypes.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        if isinstance(stream, ctypes.c_void_p):
            return stream
        return ctypes.c_void_p(getattr(stream, "cuda_stream", stream))

    def pack(self, waveforms) -> "torch.Tensor":
        """List of 1-D tensors/arrays (or a [B, N] tensor) -> one packed 1-D CUDA tensor."""
        import torch

        if isinstance(waveforms, torch.Tensor) and waveforms.dim() == 2:
            waveforms = list(waveforms)
        parts = [torch.as_tensor(w) for w in waveforms]
        if len(parts) != self.n_utts or any(int(p.numel()) != int(n) for p, n in zip(parts, self.n_samples)):
            raise ValueError("waveform lengths do not match the plan")
        return torch.cat([p.reshape(-1) for p in parts]).to(self.device)

    def waveform_stats(self, wave, out=None, stream=None):
        """Per-utterance (mean, population variance) of a packed waveform in float64 -> ``[B, 2]`` CUDA tensor: one pass
        over the samples (N2).  Feed it to :meth:`logmel` as ``znorm_stats`` for the fused z-score."""
        import torch

        dt = {torch.float32: _cabi.AAT_F32, torch.float64: _cabi.AAT_F64}.get(wave.dtype)
        if dt is None or not wave.is_cuda or not wave.is_contiguous() or wave.numel() != self.total_samples:
            raise TypeError("wave must be a contiguous packed CUDA tensor (float32 or float64) matching the plan")
        if out is None:
            out = torch.empty((self.n_utts, 2), dtype=torch.float64, device=self.device)
        _cabi.check(_cabi.lib().aat_normalize(self.ctx.handle, self.handle, wave.data_ptr(), dt, _cabi.AAT_NORM_ZSCORE,
                                              None, _cabi.AAT_F64, out.data_ptr(), self._stream(stream)))
        return out

    def logmel(self, wave, with_amp: bool = True, znorm_stats=None, stream=None):
        """K1+K2 on a packed waveform tensor (float32 or float64, CUDA).  Returns the packed mel.

        znorm_stats : optional ``[B, 2]`` float64 CUDA tensor from :meth:`waveform_stats`; the samples are then
                      z-scored ``(x - mean) / (std + 1e-6)`` in float64 as they are staged (what the reference's call
                      sites do before ``get_melspec``), without a normalised copy of the waveform"""
        import torch

        if not (isinstance(wave, torch.Tensor) and wave.is_cuda and wave.is_contiguous()):
            raise TypeError("wave must be a contiguous CUDA tensor (use .pack() for lists)")
        if wave.numel() != self.total_samples:
            raise ValueError(f"packed waveform has {wave.numel()} samples, plan expects {self.total_samples}")
        dt = {torch.float32: _cabi.AAT_F32, torch.float64: _cabi.AAT_F64}.get(wave.dtype)
        if dt is None:
            raise TypeError("waveform dtype must be float32 or float64")
        if znorm_stats is not None and (znorm_stats.dtype != torch.float64 or znorm_stats.numel() != 2 * self.n_utts or
                                        not znorm_stats.is_cuda or not znorm_stats.is_contiguous()):
            raise TypeError("znorm_stats must be a contiguous [B, 2] float64 CUDA tensor (waveform_stats())")
        _cabi.check(_cabi.lib().aat_logmel(self.ctx.handle, self.handle, wave.data_ptr(), dt,
                                           znorm_stats.data_ptr() if znorm_stats is not None else None,
                                           self.mel.data_ptr(), self.amp.data_ptr() if with_amp else None,
                                           self._stream(stream)))
        return self.mel

    def amplitude(self, mel=None, stream=None):
        """``-10 * mel.mean(axis=0)`` of every utterance in numpy's float32 order (ref:src/aat/tokenizer.py:67) into
        ``self.amp``: what :meth:`logmel` writes with ``with_amp=True``, as a separate, fully parallel pass."""
        src = self.mel if mel is None else mel
        if src.numel() != self.n_mels * self.total_frames or not src.is_cuda or not src.is_contiguous():
            raise ValueError("mel must be a contiguous packed CUDA tensor matching the plan")
        _cabi.check(_cabi.lib().aat_amplitude(self.ctx.handle, self.handle, src.data_ptr(), self.amp.data_ptr(),
                                              self._stream(stream)))
        return self.amp

    def boundaries(self, mel=None, use_amp: bool = True, with_minima: bool = True, with_csr: bool = True, stream=None):
        """K3.  ``mel=None`` uses this batch's own mel (and the fused amplitude curve when ``use_amp``);
        pass a packed float32 CUDA tensor to segment somebody else's mel (e.g. the reference's).
        ``with_csr`` also fills ``seg_off`` / ``n_seg`` / ``utt_seg_off`` (what :meth:`frame_csr` computes)
        from the kernel's own tail, saving a launch."""
        mel_ptr, amp_ptr = self.mel.data_ptr(), (self.amp.data_ptr() if use_amp else None)
        if mel is not None:
            if mel.numel() != self.n_mels * self.total_frames or not mel.is_cuda or not mel.is_contiguous():
                raise ValueError("mel must be a contiguous packed CUDA tensor matching the plan")
            mel_ptr, amp_ptr = mel.data_ptr(), None
        _cabi.check(_cabi.lib().aat_boundaries(
            self.ctx.handle, self.handle, mel_ptr, amp_ptr, self.seg_start.data_ptr(), self.seg_len.data_ptr(),
            self.seg_count.data_ptr(), self.minima.data_ptr() if with_minima else None,
            self.minima_count.data_ptr() if with_minima else None, self.status.data_ptr(),
            self.seg_off.data_ptr() if with_csr else None, self.n_seg.data_ptr() if with_csr else None,
            self.utt_seg_off.data_ptr() if with_csr else None, self._stream(stream)))
        return self.seg_len, self.seg_count

    def frame_csr(self):
        """Segment lengths -> packed CSR offsets in HuBERT frames (device), for :meth:`pool`."""
        _cabi.check(_cabi.lib().aat_segment_frame_csr(
            self.ctx.handle, self.handle, self.seg_len.data_ptr(), self.seg_count.data_ptr(), self.seg_off.data_ptr(),
            self.n_seg.data_ptr(), self.utt_seg_off.data_ptr(), self._stream()))
        return self.seg_off, self.n_seg

    def utterance_frame_csr(self, stream=None):
        """Packed CSR of the WHOLE-UTTERANCE encode convention, for :meth:`pool` with ``csr=``: the
        encoder ran once over each utterance, ``emb`` holds ``(N_b - 400) // 320 + 1`` rows per utterance back to back,
        and the segment that starts at sample ``s`` starts at row ``min(s // 320, T_b)`` of its utterance (the collator's
        ``// hop_length``, ref:src/aat/training/collate.py:340).  Needs :meth:`boundaries` (with its CSR) before it on the
        stream.  Returns ``(seg_off, totals)`` — buffers of their own, so the per-segment CSR stays valid beside them;
        ``totals`` is ``{segments, rows}``."""
        import torch

        if getattr(self, "_utt_csr", None) is None:  # allocated on first use (before any graph capture)
            self._utt_csr = (torch.zeros(self.total_seg_slots + 1, dtype=torch.int64, device=self.device),
                             torch.zeros(2, dtype=torch.int64, device=self.device))
        seg_off, totals = self._utt_csr
        _cabi.check(_cabi.lib().aat_utterance_frame_csr(
            self.ctx.handle, self.handle, self.seg_start.data_ptr(), self.seg_count.data_ptr(), self.utt_seg_off.data_ptr(),
            seg_off.data_ptr(), totals.data_ptr(), self._stream(stream)))
        return seg_off, totals

    def pool(self, emb, out, colsum=None, accumulate: bool = False, emb_ready: bool = False,
             rows_from_device: bool = False, share_sms: bool = False, stream=None, csr=None):
        """K4 with the device-resident CSR of :meth:`frame_csr` (or ``csr=(seg_off, totals)`` from
        :meth:`utterance_frame_csr`).  ``out`` is [capacity, D] float32; ``colsum``
        ([D+1] float64) receives the column sums of the pooled vectors, added to its content when ``accumulate``.

        emb_ready        : the previous launch on this stream is this batch's :meth:`boundaries` (or anything else that
                           does not write ``emb``): the kernel may start fetching embeddings before that launch ends
        rows_from_device : ``emb`` is an allocation of at least as many rows as the segments cover; the covered row
                           count is taken from the device (written by :meth:`boundaries`) instead of ``emb.shape[0]``
        share_sms        : other batches are in flight on other streams: one CTA per SM instead of two, so that two pool
                           kernels (or a pool kernel and a log-mel CTA) fit on an SM together (``AAT_POOL_SHARE_SMS``;
                           the order of the additions follows the CTA tiles: means may differ in the last bit from a launch
                           without it)"""
        from .pooling import _pool_device

        seg_off, totals = (self.seg_off, self._csr_totals) if csr is None else csr
        return _pool_device(self.ctx, emb, seg_off, int(out.shape[0]), totals, out, colsum,
                            self._stream(stream), accumulate, plan=self.handle, emb_ready=emb_ready,
                            rows_from_device=rows_from_device, share_sms=share_sms)

    def step(self, wave, emb, out, colsum=None, accumulate: bool = False, znorm: bool = False, emb_ready: bool = True,
             rows_from_device: bool = False, share_sms: bool = False, stream=None):
        """:meth:`logmel` (with the fused z-score when ``znorm``) -> :meth:`boundaries` -> :meth:`pool` in ONE call
        through the C ABI (``aat_tokenize_and_pool``): the same launches, a third of the host time per step, for loops
        that run thousands of steps per second.  Arguments as in the three methods; no validation beyond the C side's."""
        import torch

        bufs = getattr(self, "_step_bufs", None)
        if bufs is None:
            bufs = self._step_bufs = _cabi.AatStepBuffers(
                self.mel.data_ptr(), self.amp.data_ptr(), self.seg_start.data_ptr(), self.seg_len.data_ptr(),
                self.seg_count.data_ptr(), self.minima.data_ptr(), self.minima_count.data_ptr(), self.status.data_ptr(),
                self.seg_off.data_ptr(), self.n_seg.data_ptr(), self.utt_seg_off.data_ptr(), None)
        if znorm and not bufs.znorm_stats:
            self._znorm_stats = torch.empty((self.n_utts, 2), dtype=torch.float64, device=self.device)
            bufs.znorm_stats = self._znorm_stats.data_ptr()
        wdt = _cabi.AAT_F32 if wave.dtype == torch.float32 else _cabi.AAT_F64 if wave.dtype == torch.float64 else None
        edt = {torch.float32: _cabi.AAT_F32, torch.float16: _cabi.AAT_F16, torch.bfloat16: _cabi.AAT_BF16}.get(emb.dtype)
        if wdt is None or edt is None or wave.numel() != self.total_samples or emb.dim() != 2:
            raise TypeError("wave must be the plan's packed float32/float64 tensor, emb a [T, D] float32/16/bfloat16 tensor")
        flags = ((_cabi.AAT_POOL_ACCUMULATE if accumulate else 0) |
                 (_cabi.AAT_POOL_EMB_READY if emb_ready and not rows_from_device else 0) |
                 (_cabi.AAT_POOL_ROWS_FROM_DEVICE if rows_from_device else 0) |
                 (_cabi.AAT_POOL_SHARE_SMS if share_sms else 0))
        _cabi.check(_cabi.lib().aat_tokenize_and_pool(
            self.ctx.handle, self.handle, ctypes.byref(bufs), wave.data_ptr(), wdt, 1 if znorm else 0, emb.data_ptr(), edt,
            int(emb.shape[0]), int(emb.shape[1]), out.data_ptr(), int(out.shape[0]),
            colsum.data_ptr() if colsum is not None else None, flags, self._stream(stream)))
        return out

    # ---- host views (synchronising; for tests and the numpy-facing callers)
    def mel_of(self, b: int):
        o0, o1 = int(self.frame_off[b]), int(self.frame_off[b + 1])
        return self.mel[self.n_mels * o0: self.n_mels * o1].view(self.n_mels, o1 - o0)

    def segments_of(self, b: int):
        """(starts, lengths, padded_tail) of utterance ``b`` as numpy arrays; raises on a device-side error."""
        status = int(self.status[b].item())
        if status < 0:
            raise _cabi.AatError(status, "boundary kernel reported an error for utterance %d" % b)
        c = int(self.seg_count[b].item())
        s0 = int(self.seg_slot_off[b])
        return (self.seg_start[s0:s0 + c].cpu().numpy(), self.seg_len[s0:s0 + c].cpu().numpy(), bool(status & 1))

    def minima_of(self, b: int):
        c = int(self.minima_count[b].item())
        o0 = int(self.frame_off[b])
        return self.minima[o0:o0 + c].cpu().numpy()


def _with_device(tokenizer: AdaptiveAudioAmplitudeTokenizer, device) -> AdaptiveAudioAmplitudeTokenizer:
    import copy

    import torch

    idx = torch.device(device).index if not isinstance(device, int) else device
    if idx is None:
        idx = torch.cuda.current_device()
    if tokenizer.device == idx:
        return tokenizer
    clone = copy.copy(tokenizer)
    clone.device = idx
    return clone
