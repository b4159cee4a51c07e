"""Per-process, per-device handle on the native library's ``aat_ctx``.

The reference tokenizer object is held by the collator and thereby forked or pickled into
DataLoader worker processes (ref:src/aat/training/trainer.py:49).  A CUDA context cannot cross a
fork, so the native context is created lazily, keyed by ``(pid, device, config)``, and is never
pickled.  Use the ``spawn`` start method (or keep CUDA uninitialised in the parent) when workers
call into the tokenizer.
"""
from __future__ import annotations

import ctypes
import os
import threading

import numpy as np

from . import _cabi

_lock = threading.Lock()
_cache: dict = {}


def _default_device() -> int:
    try:
        import torch

        if torch.cuda.is_available():
            return int(torch.cuda.current_device())
    except Exception:
        pass
    return 0


class Context:
    """Owns one ``aat_ctx*``: constant tables on the device + scratch."""

    def __init__(self, device: int, config: _cabi.AatConfig, window: np.ndarray, mel_filters: np.ndarray):
        lib = _cabi.lib()
        self.device = int(device)
        self.config = config
        window = np.ascontiguousarray(window, dtype=np.float64)
        mel_filters = np.ascontiguousarray(mel_filters, dtype=np.float64)
        if window.shape != (config.n_fft,):
            raise ValueError(f"window must have shape ({config.n_fft},), got {window.shape}")
        if mel_filters.shape != (config.n_fft // 2 + 1, config.num_mel_filters):
            raise ValueError(f"mel_filters must have shape ({config.n_fft // 2 + 1}, {config.num_mel_filters})")
        handle = ctypes.c_void_p()
        status = lib.aat_create(self.device, ctypes.byref(config), window.ctypes.data, mel_filters.ctypes.data,
                                ctypes.byref(handle))
        if status == _cabi.AAT_ERR_UNSUPPORTED:
            msg = lib.aat_last_error().decode()
            raise NotImplementedError(msg)
        _cabi.check(status)
        self.handle = handle
        self.pid = os.getpid()

    def close(self):
        if getattr(self, "handle", None) and self.pid == os.getpid():
            _cabi.lib().aat_destroy(self.handle)
        self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def make_config(running_mean_points, min_segment_frames, max_segment_frames, n_fft, hop_length, num_mel_filters,
                sampling_rate, max_amplitude_for_minima) -> _cabi.AatConfig:
    cfg = _cabi.AatConfig()
    cfg.sampling_rate = int(sampling_rate)
    cfg.n_fft = int(n_fft)
    cfg.hop_length = int(hop_length)
    cfg.num_mel_filters = int(num_mel_filters)
    cfg.running_mean_points = int(running_mean_points)
    cfg.min_segment_frames = int(min_segment_frames)
    cfg.max_segment_frames = int(max_segment_frames)
    # numpy compares the float32 running mean with the Python scalar cast to float32 (NEP 50)
    cfg.max_amplitude_for_minima = float(np.float32(max_amplitude_for_minima))
    return cfg


def get_context(config: _cabi.AatConfig, window: np.ndarray, mel_filters: np.ndarray, device=None) -> Context:
    device = _default_device() if device is None else int(device)
    key = (os.getpid(), device, bytes(config), window.tobytes(), mel_filters.tobytes())
    with _lock:
        ctx = _cache.get(key)
        if ctx is None or ctx.handle is None:
            ctx = Context(device, config, window, mel_filters)
            _cache[key] = ctx
        return ctx


_default: dict = {}


def default_context(device=None) -> Context:
    """Context with the reference's default constructor arguments (used by the pooling entry point).  Remembered per
    (process, device): building the constant tables and the cache key costs more than a small pooling call."""
    device = _default_device() if device is None else int(device)
    hit = _default.get((os.getpid(), device))
    if hit is not None and hit.handle is not None:
        return hit
    from .constants import hann_window_periodic, mel_filter_bank_slaney

    cfg = make_config(12, 2000, 24000, 400, 160, 64, 16000, 15)
    ctx = get_context(cfg, hann_window_periodic(400), mel_filter_bank_slaney(201, 64, 0.0, 8000.0, 16000), device)
    _default[(os.getpid(), device)] = ctx
    return ctx
