"""ORACLE (test infrastructure, not product code) — self-contained restatement.

``ref_port.py`` calls the reference's third-party dependencies directly.  This
module restates *their* algorithms in plain numpy so the checker does not
silently inherit a bug from (or a version drift of) those libraries, and so the
sequential float32 semantics that decide the boundaries are written out as
explicit loops that the CUDA kernels can be compared against step by step.

Third-party algorithms restated (none is vendored under /root/reference):
  transformers 5.5.0  audio_utils.hertz_to_mel / mel_to_hertz   TF:audio_utils.py:263-332
                      _create_triangular_filter_bank            TF:audio_utils.py:356-375
                      mel_filter_bank (norm=slaney, scale=slaney) TF:audio_utils.py:453-544
                      window_function("hann", periodic)         TF:audio_utils.py:560-620
                      spectrogram(power=2, mel, log10)          TF:audio_utils.py:769-830
  scipy 1.18.1        signal._peak_finding._boolrelextrema      SP:signal/_peak_finding.py:22-79
  numpy 2.3.5         mean(axis=0) / cumsum on float32 (sequential accumulation)

Nothing under ``audio-adaptive-tokenizer_b200/`` may import this module.
"""
from __future__ import annotations

import numpy as np


# --------------------------------------------------------------------------- constants
def hertz_to_mel_slaney(freq):
    freq = np.asarray(freq, dtype=np.float64)
    mels = 3.0 * freq / 200.0
    logstep = 27.0 / np.log(6.4)
    log_region = freq >= 1000.0
    safe = np.where(log_region, freq, 1000.0)
    return np.where(log_region, 15.0 + np.log(safe / 1000.0) * logstep, mels)


def mel_to_hertz_slaney(mels):
    mels = np.asarray(mels, dtype=np.float64)
    freq = 200.0 * mels / 3.0
    logstep = np.log(6.4) / 27.0
    log_region = mels >= 15.0
    return np.where(log_region, 1000.0 * np.exp(logstep * (mels - 15.0)), freq)


def mel_filter_bank_slaney(num_frequency_bins=201, num_mel_filters=64, min_frequency=0.0,
                           max_frequency=8000.0, sampling_rate=16000):
    mel_min = float(hertz_to_mel_slaney(min_frequency))
    mel_max = float(hertz_to_mel_slaney(max_frequency))
    mel_freqs = np.linspace(mel_min, mel_max, num_mel_filters + 2)
    filter_freqs = mel_to_hertz_slaney(mel_freqs)
    fft_freqs = np.linspace(0, sampling_rate // 2, num_frequency_bins)
    filter_diff = np.diff(filter_freqs)
    slopes = np.expand_dims(filter_freqs, 0) - np.expand_dims(fft_freqs, 1)
    down = -slopes[:, :-2] / filter_diff[:-1]
    up = slopes[:, 2:] / filter_diff[1:]
    fb = np.maximum(np.zeros(1), np.minimum(down, up))
    enorm = 2.0 / (filter_freqs[2 : num_mel_filters + 2] - filter_freqs[:num_mel_filters])
    return fb * np.expand_dims(enorm, 0)


def hann_periodic(n=400):
    return np.hanning(n + 1)[:-1]


# --------------------------------------------------------------------------- log-mel
def reflect_index(p, n):
    """Source index for padded position ``p - pad`` under numpy's ``mode='reflect'``."""
    p = np.asarray(p, dtype=np.int64)
    if n == 1:
        return np.zeros_like(p)
    period = 2 * (n - 1)
    m = np.mod(p, period)
    return np.where(m < n, m, period - m)


def logmel(waveform, window=None, mel_filters=None, n_fft=400, hop=160, batched=True):
    """spectrogram(power=2, mel_filters, log_mel='log10') -> (n_mels, T) float32.

    ``batched=True`` frames everything at once and calls rfft on the matrix; the
    survey measured this bit-identical to the per-frame loop (same pocketfft
    plan per row).  ``batched=False`` is the literal per-frame loop.
    """
    window = hann_periodic(n_fft) if window is None else np.asarray(window, dtype=np.float64)
    mel_filters = mel_filter_bank_slaney() if mel_filters is None else mel_filters
    x = np.asarray(waveform)
    n = x.shape[0]
    pad = n_fft // 2
    idx = reflect_index(np.arange(-pad, n + pad), n)
    xp = x[idx].astype(np.float64)
    num_frames = int(1 + np.floor((xp.size - n_fft) / hop))
    if batched:
        fidx = np.arange(num_frames)[:, None] * hop + np.arange(n_fft)[None, :]
        spec = np.fft.rfft(xp[fidx] * window[None, :], axis=-1).astype(np.complex64)
    else:
        spec = np.empty((num_frames, n_fft // 2 + 1), dtype=np.complex64)
        buf = np.zeros(n_fft)
        for f in range(num_frames):
            buf[:] = xp[f * hop : f * hop + n_fft]
            buf *= window
            spec[f] = np.fft.rfft(buf)
    power = np.abs(spec, dtype=np.float64) ** 2.0
    mel = np.maximum(1e-10, np.dot(mel_filters.T, power.T))
    return np.asarray(np.log10(mel), dtype=np.float32)


# --------------------------------------------------------------------------- boundaries
def amplitude_curve_seq(melspec):
    """-10 * mean over mels, written as the sequential float32 loop numpy performs."""
    mel = np.asarray(melspec, dtype=np.float32)
    acc = mel[0].copy()
    for r in range(1, mel.shape[0]):
        acc = (acc + mel[r]).astype(np.float32)
    mean = (acc / np.float32(mel.shape[0])).astype(np.float32)
    return (np.float32(-10.0) * mean).astype(np.float32)


def cumsum_seq(x):
    x = np.asarray(x, dtype=np.float32)
    out = np.empty_like(x)
    s = np.float32(0.0)
    first = True
    for i in range(x.size):
        s = x[i] if first else np.float32(s + x[i])
        first = False
        out[i] = s
    return out


def find_minimas_seq(melspec, running_mean_points=12, max_amplitude_for_minima=15, intermediates=False):
    amp = amplitude_curve_seq(melspec)
    cs = cumsum_seq(amp)
    n = running_mean_points
    if amp.size > n:
        rm = ((cs[n:] - cs[:-n]).astype(np.float32) / np.float32(n)).astype(np.float32)
    else:
        rm = np.zeros(0, dtype=np.float32)
    eps = np.float32(1e-5)
    minimas = []
    for i in range(1, rm.size - 1):
        if rm[i] > np.float32(rm[i + 1] + eps) and rm[i] > np.float32(rm[i - 1] + eps):
            if rm[i] > np.float32(max_amplitude_for_minima):
                minimas.append(i)
    out = np.asarray(minimas, dtype=np.int64)
    if intermediates:
        return out, amp, cs, rm
    return out


def segment_state_machine(n_samples, boarders, min_frames, max_frames):
    """(start, length) of every emitted segment — integer restatement of
    ref:src/aat/tokenizer.py:141-183 including ``np.split``'s clamp semantics
    when the adjusted last cut falls below the previous one (min > max configs).
    Returns (starts, lengths, padded_tail) where padded_tail says the final
    entry is the zero-padded tail (its length is ``min_frames``)."""
    starts, lengths = [], []
    prev = 0
    for b in boarders:
        length = b - prev
        if length < min_frames:
            continue
        if length > max_frames:
            k = length // max_frames
            cuts = [max_frames * (j + 1) for j in range(k)]
            gap = length - cuts[-1]
            if gap == 0:
                cuts = cuts[:-1]
            elif gap < min_frames:
                cuts[-1] = length - min_frames
            lo = 0
            for c in cuts:
                starts.append(prev + min(lo, length))
                lengths.append(max(0, min(c, length) - min(lo, length)))
                lo = c
            starts.append(prev + min(lo, length))
            lengths.append(length - min(lo, length))
        else:
            starts.append(prev)
            lengths.append(length)
        prev = b
    padded_tail = prev != n_samples
    if padded_tail:
        if n_samples - prev > min_frames:
            raise ValueError("could not broadcast tail into min_segment_frames buffer")
        starts.append(prev)
        lengths.append(min_frames)
    return starts, lengths, padded_tail


# --------------------------------------------------------------------------- pooling
def mean_pool_csr_f64(emb, seg_off):
    """High-precision pooled means (fp64 accumulate) for tolerance accounting."""
    emb = np.asarray(emb)
    off = np.asarray(seg_off, dtype=np.int64)
    out = np.empty((off.size - 1, emb.shape[1]), dtype=np.float64)
    for i in range(off.size - 1):
        seg = emb[off[i] : off[i + 1]].astype(np.float64)
        out[i] = seg.sum(axis=0) / seg.shape[0] if seg.shape[0] else np.nan
    return out


def utterance_frame_offsets_loop(segment_lengths, n_samples):
    """Frame CSR of ONE whole-utterance encoding (convention (ii)), as a plain loop: the encoder
    (receptive field 400, stride 320: TF:models/hubert/modeling_hubert.py:675-688) yields T rows for the utterance; a
    segment that starts at sample s starts at row min(s // 320, T), like the collator's `// hop_length`
    (ref:src/aat/training/collate.py:340).  There is no reference code for this convention; this loop DEFINES it for
    the tests (parity unpinned by construction)."""
    n_samples = int(n_samples)
    rows = 0 if n_samples < 400 else (n_samples - 400) // 320 + 1
    off, start = [], 0
    for length in segment_lengths:
        off.append(min(start // 320, rows))
        start += int(length)
    assert start >= n_samples
    off.append(min(start // 320, rows))
    return np.asarray(off, dtype=np.int64)
