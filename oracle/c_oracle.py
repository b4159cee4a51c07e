"""ORACLE (test infrastructure, not product code) — ctypes front for oracle/aat_oracle.c.

``build()`` compiles ``libaat_oracle.so`` with the Makefile beside it; the loader
never falls back to anything else.  Nothing under
``audio-adaptive-tokenizer_b200/`` may import this module.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libaat_oracle.so")
_lib = None

_i64 = ctypes.c_int64
_f32p = ctypes.POINTER(ctypes.c_float)
_f64p = ctypes.POINTER(ctypes.c_double)
_i64p = ctypes.POINTER(ctypes.c_int64)
_i32p = ctypes.POINTER(ctypes.c_int32)


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "aat_oracle.c")
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", _HERE, "-B", "libaat_oracle.so"], stdout=subprocess.DEVNULL)
    return _SO


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_SO):
            build()
        _lib = ctypes.CDLL(_SO)
        _lib.orc_find_minimas.restype = _i64
        _lib.orc_find_minimas.argtypes = [_f32p, _i64, _i64, _i64, _i64, ctypes.c_float, _i64p, _f32p, _f32p, _f32p]
        _lib.orc_state_machine.restype = _i64
        _lib.orc_state_machine.argtypes = [_i64, _i64p, _i64, _i64, _i64, _i64p, _i64p, _i64, _i32p]
        _lib.orc_mean_pool_f32.restype = None
        _lib.orc_mean_pool_f32.argtypes = [_f32p, _i64, _i64p, _i64, _f32p]
        _lib.orc_mean_pool_f64.restype = None
        _lib.orc_mean_pool_f64.argtypes = [_f32p, _i64, _i64p, _i64, _f64p]
        _lib.orc_logmel_naive.restype = None
        _lib.orc_logmel_naive.argtypes = [_f64p, _i64, _f64p, _i64, _i64, _f64p, _i64, _f32p]
    return _lib


def _p(a, t):
    return a.ctypes.data_as(t)


def find_minimas(mel, running_mean_points=12, max_amplitude=15.0, intermediates=False):
    mel = np.ascontiguousarray(mel, dtype=np.float32)
    n_mels, T = mel.shape
    minima = np.empty(max(T, 1), dtype=np.int64)
    amp = np.empty(max(T, 1), dtype=np.float32)
    cs = np.empty(max(T, 1), dtype=np.float32)
    rm = np.empty(max(T, 1), dtype=np.float32)
    n = lib().orc_find_minimas(_p(mel, _f32p), n_mels, T, T, running_mean_points, float(max_amplitude),
                               _p(minima, _i64p), _p(amp, _f32p), _p(cs, _f32p), _p(rm, _f32p))
    if intermediates:
        L = max(T - running_mean_points, 0)
        return minima[:n].copy(), amp[:T].copy(), cs[:T].copy(), rm[:L].copy()
    return minima[:n].copy()


def state_machine(n_samples, boarders, min_frames, max_frames):
    b = np.ascontiguousarray(boarders, dtype=np.int64)
    cap = int(n_samples // max(min_frames, 1) + n_samples // max(max_frames, 1) + len(b) + 8)
    starts = np.empty(cap, dtype=np.int64)
    lengths = np.empty(cap, dtype=np.int64)
    padded = ctypes.c_int32(0)
    n = lib().orc_state_machine(int(n_samples), _p(b, _i64p), b.size, int(min_frames), int(max_frames),
                                _p(starts, _i64p), _p(lengths, _i64p), cap, ctypes.byref(padded))
    if n == -2:
        raise ValueError("could not broadcast tail into min_segment_frames buffer")
    assert n >= 0, "oracle capacity"
    return starts[:n].copy(), lengths[:n].copy(), bool(padded.value)


def mean_pool_f32(emb, seg_off):
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    off = np.ascontiguousarray(seg_off, dtype=np.int64)
    out = np.empty((off.size - 1, emb.shape[1]), dtype=np.float32)
    lib().orc_mean_pool_f32(_p(emb, _f32p), emb.shape[1], _p(off, _i64p), off.size - 1, _p(out, _f32p))
    return out


def mean_pool_f64(emb, seg_off):
    emb = np.ascontiguousarray(emb, dtype=np.float32)
    off = np.ascontiguousarray(seg_off, dtype=np.int64)
    out = np.empty((off.size - 1, emb.shape[1]), dtype=np.float64)
    lib().orc_mean_pool_f64(_p(emb, _f32p), emb.shape[1], _p(off, _i64p), off.size - 1, _p(out, _f64p))
    return out


def logmel_naive(wave, window, mel_filters, n_fft=400, hop=160):
    wave = np.ascontiguousarray(wave, dtype=np.float64)
    window = np.ascontiguousarray(window, dtype=np.float64)
    mf = np.ascontiguousarray(mel_filters, dtype=np.float64)
    T = 1 + wave.size // hop
    out = np.empty((mf.shape[1], T), dtype=np.float32)
    lib().orc_logmel_naive(_p(wave, _f64p), wave.size, _p(window, _f64p), n_fft, hop, _p(mf, _f64p), mf.shape[1],
                           _p(out, _f32p))
    return out
