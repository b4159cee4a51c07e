"""ctypes binding of ``libaat_b200.so`` (C ABI declared in ``include/aat_b200.h``).

This is the only place the shared library is loaded.  There is no fallback: if the
library is missing the import of any compute entry point raises, and every compute
call needs a CUDA device.
"""
from __future__ import annotations

import ctypes
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("AAT_B200_LIB") or os.path.join(_HERE, "libaat_b200.so")  # override: profiles/ experiments
CSRC_DIR = os.path.join(os.path.dirname(_HERE), "csrc")

ABI_VERSION = 201  # AAT_B200_VERSION of include/aat_b200.h this binding was written against
AAT_OK = 0
AAT_ERR_INVALID = -1
AAT_ERR_UNSUPPORTED = -2
AAT_ERR_CUDA = -3
AAT_ERR_CAPACITY = -4
AAT_ERR_TAIL = -5

AAT_F32, AAT_F64, AAT_F16, AAT_BF16 = 0, 1, 2, 3
AAT_NORM_ZSCORE, AAT_NORM_W2V2 = 0, 1
AAT_POOL_ACCUMULATE, AAT_POOL_EMB_READY, AAT_POOL_ROWS_FROM_DEVICE, AAT_POOL_SHARE_SMS = 1, 2, 4, 8

c_i32 = ctypes.c_int32
c_i64 = ctypes.c_int64
c_void = ctypes.c_void_p
p_i64 = ctypes.POINTER(ctypes.c_int64)
p_f64 = ctypes.POINTER(ctypes.c_double)
p_f32 = ctypes.POINTER(ctypes.c_float)


class AatConfig(ctypes.Structure):
    """``struct aat_config`` (include/aat_b200.h)."""

    _fields_ = [
        ("sampling_rate", c_i32),
        ("n_fft", c_i32),
        ("hop_length", c_i32),
        ("num_mel_filters", c_i32),
        ("running_mean_points", c_i32),
        ("reserved0", c_i32),
        ("min_segment_frames", c_i64),
        ("max_segment_frames", c_i64),
        ("max_amplitude_for_minima", ctypes.c_float),
        ("reserved1", c_i32),
    ]


class AatStepBuffers(ctypes.Structure):
    """``struct aat_step_buffers`` (include/aat_b200.h)."""

    _fields_ = [(name, ctypes.c_void_p) for name in
                ("mel", "amp", "seg_start", "seg_len", "seg_count", "minima", "minima_count", "status", "seg_off", "n_seg",
                 "utt_seg_off", "znorm_stats")]


class AatError(RuntimeError):
    def __init__(self, status: int, message: str):
        super().__init__(f"aat_b200 status {status}: {message}")
        self.status = status


# every symbol include/aat_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "aat_version": (ctypes.c_int, []),
    "aat_last_error": (ctypes.c_char_p, []),
    "aat_kernel_launch_count": (c_i64, []),
    "aat_kernel_launch_count_add": (None, [c_i64]),
    "aat_create": (ctypes.c_int, [ctypes.c_int, ctypes.POINTER(AatConfig), c_void, c_void, ctypes.POINTER(c_void)]),
    "aat_destroy": (ctypes.c_int, [c_void]),
    "aat_profile_enable": (ctypes.c_int, [c_void, ctypes.c_uint32]),
    "aat_profile_summary": (ctypes.c_int, [c_void, c_void, c_void]),
    "aat_profile_sample_every": (ctypes.c_int, [c_void, c_i32]),
    "aat_get_config": (ctypes.c_int, [c_void, ctypes.POINTER(AatConfig)]),
    "aat_plan_create": (ctypes.c_int, [c_void, c_i32, c_void, ctypes.POINTER(c_void)]),
    "aat_plan_destroy": (ctypes.c_int, [c_void]),
    "aat_plan_total_samples": (c_i64, [c_void]),
    "aat_plan_total_frames": (c_i64, [c_void]),
    "aat_plan_total_seg_slots": (c_i64, [c_void]),
    "aat_plan_offsets": (ctypes.c_int, [c_void, c_void, c_void, c_void]),
    "aat_logmel": (ctypes.c_int, [c_void, c_void, c_void, ctypes.c_int, c_void, c_void, c_void, c_void]),
    "aat_amplitude": (ctypes.c_int, [c_void, c_void, c_void, c_void, c_void]),
    "aat_boundaries": (ctypes.c_int, [c_void] * 14),
    "aat_process_boarders": (ctypes.c_int, [c_void, c_i64, c_void, c_i64, c_void, c_void, c_i64, c_void, c_void, c_void]),
    "aat_segment_frame_csr": (ctypes.c_int, [c_void] * 8),
    "aat_utterance_frame_csr": (ctypes.c_int, [c_void] * 8),
    "aat_segment_mean_pool": (ctypes.c_int, [c_void, c_void, c_void, ctypes.c_int, c_i64, c_i32, c_void, c_i64, c_void,
                                             c_void, c_void, ctypes.c_int, c_void]),
    "aat_tokenize_and_pool": (ctypes.c_int, [c_void, c_void, ctypes.POINTER(AatStepBuffers), c_void, ctypes.c_int, ctypes.c_int,
                                             c_void, ctypes.c_int, c_i64, c_i32, c_void, c_i64, c_void, ctypes.c_int, c_void]),
    "aat_colsum_accumulate": (ctypes.c_int, [c_void, c_void, c_void, c_i32, c_void]),
    "aat_colsum_finalize": (ctypes.c_int, [c_void, c_void, c_i32, c_void, c_void]),
    "aat_normalize": (ctypes.c_int, [c_void, c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, ctypes.c_int, c_void, c_void]),
    "aat_pad_segment_boarders": (ctypes.c_int, [c_void, c_void, c_void, c_void, c_i64, c_void, c_void, c_void, c_void]),
    "aat_scatter_segments": (ctypes.c_int, [c_void, c_void, c_i64, c_i32, c_void, c_i64, c_i64, c_void, c_void, c_void,
                                            c_void]),
    "aat_scatter_mel_segments": (ctypes.c_int, [c_void, c_void, c_void, c_void, c_i64, c_i64, c_void, c_void, c_void]),
    "aat_scatter_mel_tiles": (ctypes.c_int, [c_void, c_i32, c_void, c_void, c_void, c_void, c_void, c_i64, c_i64, c_void, c_void,
                                             c_void]),
    "aat_normalize_padded": (ctypes.c_int, [c_void, c_void, c_void, ctypes.c_int, ctypes.c_int, c_void, c_i64, c_void, c_void, c_void]),
    "aat_masked_mean_pool": (ctypes.c_int, [c_void, c_void, ctypes.c_int, c_i64, c_i64, c_i32, c_void, c_void, c_void, c_void]),
    "aat_synth_workspace_bytes": (c_i64, [c_void]),
    "aat_synth_waveforms": (ctypes.c_int, [c_void, c_void, ctypes.c_uint64, c_i64, c_void, c_void, c_void]),
    "aat_synth_normal": (ctypes.c_int, [c_void, c_void, c_i64, ctypes.c_uint64, c_void]),
    "aat_host_logmel": (ctypes.c_int, [c_void, c_void, ctypes.c_int, c_i64, c_void]),
    "aat_host_find_minimas": (ctypes.c_int, [c_void, c_void, c_i64, c_void, c_void]),
    "aat_host_process_boarders": (ctypes.c_int, [c_void, c_i64, c_void, c_i64, c_void, c_void, c_i64, c_void, c_void]),
    "aat_host_tokenize": (ctypes.c_int, [c_void, c_void, ctypes.c_int, c_i64, c_void, c_void, c_void, c_void, c_void,
                                         c_void, c_i64, c_void, c_void]),
    "aat_host_mean_pool": (ctypes.c_int, [c_void, c_void, ctypes.c_int, c_i64, c_i32, c_void, c_i64, c_void, c_void]),
    "aat_host_mean_pool_list": (ctypes.c_int, [c_void, c_void, c_void, c_i64, ctypes.c_int, c_i32, c_void, c_void]),
    "aat_segment_capacity": (c_i64, [ctypes.POINTER(AatConfig), c_i64]),
    "aat_num_mel_frames": (c_i64, [ctypes.POINTER(AatConfig), c_i64]),
}

_lib = None


def build(verbose: bool = False) -> str:
    """Compile ``libaat_b200.so`` in-tree with nvcc for sm_100a (csrc/Makefile)."""
    cmd = ["make", "-C", CSRC_DIR, "-j4"]
    proc = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if verbose or proc.returncode != 0:
        print(proc.stdout)
    if proc.returncode != 0:
        raise RuntimeError("building libaat_b200.so failed")
    return LIB_PATH


def lib() -> ctypes.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(
                f"{LIB_PATH} is missing: the CUDA extension is the product and there is no CPU fallback. "
                "Build it with `python -c 'import __graft_entry__ as g; g.build()'` or "
                "`make -C audio-adaptive-tokenizer_b200/csrc`."
            )
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        if handle.aat_version() != ABI_VERSION:
            raise ImportError(f"{LIB_PATH} has ABI version {handle.aat_version()}, this binding needs {ABI_VERSION}: "
                              "rebuild it (make -C audio-adaptive-tokenizer_b200/csrc)")
        _lib = handle
    return _lib


def check(status: int) -> None:
    if status != AAT_OK:
        msg = lib().aat_last_error()
        raise AatError(status, msg.decode("utf-8", "replace") if msg else "")


KERNEL_NAMES = ("logmel", "boundaries", "frame_csr", "pool")


def profile_enable(ctx_handle, names=KERNEL_NAMES, every: int = 1) -> None:
    """Record CUDA-event pairs around the named kernels (every ``every``-th launch of each)."""
    check(lib().aat_profile_sample_every(ctx_handle, every))
    mask = 0
    for n in names:
        mask |= 1 << KERNEL_NAMES.index(n)
    check(lib().aat_profile_enable(ctx_handle, mask))


def profile_summary(ctx_handle) -> dict:
    """{kernel name: (launches, total_ms)} of the launches recorded since the last enable."""
    n = (ctypes.c_int64 * len(KERNEL_NAMES))()
    ms = (ctypes.c_double * len(KERNEL_NAMES))()
    check(lib().aat_profile_summary(ctx_handle, n, ms))
    return {name: (int(n[i]), float(ms[i])) for i, name in enumerate(KERNEL_NAMES)}


def launch_count() -> int:
    return int(lib().aat_kernel_launch_count())
