#!/usr/bin/env python
"""Benchmark of the tokenization front end: audio-hours/s tokenized (log-mel + segment + pool).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload c2|c3|c4]

A *step* is one pass of the whole hot path over one batch of synthetic input: K1+K2 log-mel,
K3 boundaries (+ frame CSR in its tail), K4 mean-pool (with the column-sum epilogue feeding the dataset mean).
The headline workload is BASELINE.json configs[1] ("c2"): 64 x 16 s utterances, HuBERT-base 768-d embeddings
on one B200; with N GPUs every rank runs that batch on its own shard (weak scaling, no data-path
collective) and the ranks meet once, in the dataset-mean allreduce, at the end of the timed region.

One JSON line on stdout (rank 0).
  value        whole-job audio-hours/s with inputs resident in HBM (generated there by the library's counter-based
               generator, `aat_synth_*`; distinct utterances in every rotating buffer set)
  e2e          the same metric through the public batched API with HOST (pinned) buffers, H2D/D2H inside the timed
               region, plus the plain pinned-copy rate measured in the same run (`copy_peak`) as its own roofline
  roofline     the pool kernel: algorithmic bytes / duration, against MEASURED_PEAKS.json.  Two clocks are printed:
               `us_per_launch` = K back-to-back launches on rotating inputs inside ONE event pair (the kernel's
               sustained rate; an event pair around a single ~30 us kernel measures the pair as much as the kernel),
               and `sampled_in_step` = event pairs around every n-th launch of the timed region
  cpu_baseline the oracle port of the reference's CPU path on a bounded sample (rank 0, N = 1)
  configs      sub-records of the other BASELINE configs measured in the same run: c1 (one 10 s clip through the numpy
               API), c3 (256 x 20 s, D = 1024), c4 (8 x 30 min), and c5: the dataset job — >= 1000 synthetic
               audio-hours of DISTINCT utterances sharded over the N ranks, generated on the device chunk by chunk
               and consumed by the path, ending in the one allreduce of the dataset-mean embedding, whose value is
               checked against an independent reduction
`--impl reference` times the reference's CPU path (oracle port) alone with all host cores.

The kernels of a step are chained by programmatic dependent launch (each one's prologue overlaps its
predecessor's tail).  `kernel_us` comes from an extra, untimed pass with events around every kernel.
"""
from __future__ import annotations

import argparse
import json
import math
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "audio-adaptive-tokenizer_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

METRIC = "audio_hours_per_sec_tokenized"
UNIT = "audio-hours/s"
WORKLOADS = {
    # name: (batch, samples per utterance, embedding dim, BASELINE.json config index)
    "c2": (64, 256_000, 768, 1),
    "c3": (256, 320_000, 1024, 2),
    "c4": (8, 28_800_000, 768, 3),
}
FALLBACK_HBM_GBS = 6650.0  # /opt/skills/guides/B200_PROFILING.md fallback
C5_AUDIO_HOURS = 1000.0    # BASELINE.json configs[4]
DEFAULT_ROTATE = 4


def workload_name(name):
    b, n, d, idx = WORKLOADS[name]
    return f"configs[{idx}]: batch {b} x {n / 16000:g} s synthetic 16 kHz utterances, {d}-d HuBERT-shaped embeddings"


def hubert_rows_upper_bound(n_samples, min_segment_frames=2000):
    """Rows an utterance's segments can cover under the per-segment-encode convention: sum_i ((L_i - 400) // 320 + 1)
    <= sum_i L_i / 320, and the segments add up to at most n_samples + min_segment_frames (zero-padded tail)."""
    return (n_samples + min_segment_frames) // 320 + 2


def workload_config(name, world, rotate):
    """`config` of the JSON line: the workload only, identical in both arms (the driver compares them)."""
    b, n, d, _ = WORKLOADS[name]
    resident_mb = rotate * (b * n * 4 + b * hubert_rows_upper_bound(n) * d * 4) / 1e6
    return {
        "workload": workload_name(name), "batch_per_gpu": b, "samples_per_utterance": n, "dim": d,
        "embedding_dtype": "f32", "segment_convention": "per-segment encode: n_i = (L_i - 400) // 320 + 1 frames",
        "parallelism": f"utterance-sharded dp{world}, one allreduce of {d + 1} f64",
        "l2": f"b200 arm: inputs rotate over {rotate} buffer sets of distinct utterances (~{resident_mb:.0f} MB) "
              f"> 126 MB L2; reference arm: host memory",
    }


# ------------------------------------------------------------------------------------------- clocks
class ClockSampler:
    """`nvidia-smi -lms` in the background; samples are time-stamped so that only those taken inside the
    timed region are summarised (the process is started early because it needs ~0.3 s to produce a line)."""
    FIELDS = ("timestamp,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.samples = []
        self.proc = None
        self.gpu_index = gpu_index

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", f"--id={self.gpu_index}", f"--query-gpu={self.FIELDS}", "--format=csv,noheader,nounits",
                 "-lms", "20"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, bufsize=1)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:
            self.proc = None

    def _read(self):
        import datetime

        for line in self.proc.stdout:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) >= 8:
                try:
                    ts = datetime.datetime.strptime(parts[0], "%Y/%m/%d %H:%M:%S.%f").timestamp()
                except ValueError:
                    ts = time.time()
                self.samples.append((ts, parts[1:]))

    def stop(self, t0=None, t1=None):
        if self.proc is not None:
            self.proc.terminate()
        inside = [p for ts, p in self.samples if t0 is None or (t0 <= ts <= t1)]
        scope = "timed region"
        if not inside:  # region shorter than the sampling period: fall back to the samples nearest to it
            mid = 0.5 * ((t0 or 0) + (t1 or 0))
            inside = [p for _, p in sorted(self.samples, key=lambda s: abs(s[0] - mid))[:3]]
            scope = "nearest samples (region shorter than the sampling period)"
        sm, reasons, sm_max = [], set(), None
        for p in inside:
            try:
                sm.append(float(p[0]))
                sm_max = float(p[1])
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), p[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": sm_max, "reasons": sorted(reasons),
                "samples": len(sm), "scope": scope}


# ------------------------------------------------------------------------------------------- CPU baseline
_LOCAL_ITEMS = []  # per-process inputs of the CPU arm (waveform, per-segment embeddings), built before the timed phase
_TOK = []


def _cpu_tokenizer():
    if not _TOK:
        from oracle import ref_port

        _TOK.append(ref_port.RefTokenizer())
    return _TOK[0]


def _cpu_tokenize_one(item):
    """Reference CPU path for one utterance: log-mel + minima + merge/split + per-segment mean-pool.  The embeddings
    are inputs (the GPU arm's are resident before its timed region too): they were generated in `_prepare_local_items`
    from this utterance's own segmentation."""
    from oracle import ref_port

    wave, embs = item
    _cpu_tokenizer().segment_lengths(wave)
    ref_port.mean_pool_segments(embs)


def _prepare_local_items(name, n_local):
    import torch

    from aat_b200 import synth

    _, n, dim, idx = WORKLOADS[name]
    n = min(n, 4_800_000)  # 30-min streams: time a 5-min slice per item, throughput is linear in length
    items = []
    for i in range(n_local):
        wave = synth.bursty_speech(n, synth.seed_for(idx + 1, 5000 + i)).astype(np.float64)
        lengths, _, _ = _cpu_tokenizer().segment_lengths(wave)  # untimed: fixes the embedding shapes
        g = torch.Generator().manual_seed(i)
        embs = [torch.randn(1, int(f), dim, generator=g) for f in synth.hubert_frames(lengths)]
        items.append((wave, embs))
    _LOCAL_ITEMS[:] = items
    _cpu_tokenize_one(_LOCAL_ITEMS[0])  # warm imports (transformers, scipy)


def _cpu_worker_init(name="c2", n_local=4):
    """One BLAS / torch thread per worker process (the pool already uses every core), and the worker's own
    copy of the synthetic utterances, so that nothing but task indices crosses a pipe in the timed phase."""
    try:
        import threadpoolctl

        _cpu_worker_init.limit = threadpoolctl.threadpool_limits(1)
    except Exception:
        pass
    import torch

    torch.set_num_threads(1)
    _prepare_local_items(name, n_local)


def _cpu_task(i):
    _cpu_tokenize_one(_LOCAL_ITEMS[i % len(_LOCAL_ITEMS)])


def _cpu_ready(_):
    time.sleep(0.05)  # keeps the task on this worker long enough for every worker to take one
    return len(_LOCAL_ITEMS)


def cpu_reference_throughput(name, n_utts, pool=None):
    """Audio-hours/s of the oracle port (reference algorithm, same third-party calls) on host cores: `n_utts`
    utterances of the workload, in this process or spread over a worker pool whose processes already hold
    their inputs (waveforms AND embeddings)."""
    _, n, dim, idx = WORKLOADS[name]
    n = min(n, 4_800_000)
    if pool is None and not _LOCAL_ITEMS:
        _prepare_local_items(name, 8)
    t0 = time.perf_counter()
    if pool is not None:
        pool.map(_cpu_task, range(n_utts), chunksize=1)
    else:
        for i in range(n_utts):
            _cpu_task(i)
    dt = time.perf_counter() - t0
    return (n_utts * n / 16000 / 3600) / dt, dt


def run_reference(args):
    """`--impl reference`: the reference's CPU implementation of the path (oracle port: the reference
    is pure Python and /root/reference does not travel), all host cores, bounded sample per step."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    _, n, dim, _ = WORKLOADS[args.workload]
    import multiprocessing as mp

    with mp.get_context("fork").Pool(cores, initializer=_cpu_worker_init, initargs=(args.workload, 4)) as pool:
        pool.map(_cpu_ready, range(4 * cores), chunksize=1)  # wait for the initialisers (imports, inputs)
        # size the per-step sample so that warm-up + K steps end within ~2.5 minutes whatever K is
        cpu_reference_throughput(args.workload, cores, pool)
        _, t_probe = cpu_reference_throughput(args.workload, cores, pool)  # one utterance per core
        budget_s = 150.0
        per_step = int(budget_s / max(t_probe, 1e-3) / max(args.steps + args.warmup, 1)) * cores
        per_step = max(cores, min(64 * cores, per_step))
        for _ in range(args.warmup):
            cpu_reference_throughput(args.workload, per_step, pool)
        vals, t_total = [], 0.0
        for _ in range(args.steps):
            v, dt = cpu_reference_throughput(args.workload, per_step, pool)
            vals.append(v)
            t_total += dt
    value = float(np.mean(vals))
    sample = (f"{per_step} utterances of {min(n, 4_800_000) / 16000:g} s per step, multiprocessing.Pool({cores}), "
              f"inputs (waveforms and per-segment embeddings) resident in the workers, 1 BLAS thread each")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * t_total / max(args.steps, 1), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64 (log-mel) / f32 (boundaries, pool)", "data": "synthetic",
        "config": workload_config(args.workload, max(world, args.gpus), args.rotate),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample,
                         "host": host_description()},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def host_description():
    """What the CPU figures were measured on (state core count, CPU model and library versions)."""
    info = {"cpu_count": os.cpu_count()}
    try:
        with open("/proc/cpuinfo") as f:
            for line in f:
                if line.startswith("model name"):
                    info["cpu_model"] = line.split(":", 1)[1].strip()
                    break
    except OSError:
        pass
    for mod in ("numpy", "scipy", "transformers", "torch"):
        try:
            info[mod] = __import__(mod).__version__
        except Exception:
            info[mod] = None
    try:
        import torch

        info["torch_threads"] = torch.get_num_threads()
    except Exception:
        pass
    return info


# ------------------------------------------------------------------------------------------- B200 arm
def bind_to_gpu_numa_node(local_rank):
    """Run this rank (and therefore first-touch its pinned host buffers) on the NUMA node its GPU hangs off:
    N ranks reading pinned memory of one node was what capped the round-1 e2e figure at N >= 4."""
    try:
        out = subprocess.run(["nvidia-smi", f"--id={local_rank}", "--query-gpu=pci.bus_id", "--format=csv,noheader"],
                             stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True, timeout=20).stdout.strip()
        bus = out.lower()
        if bus.startswith("00000000:"):
            bus = bus[4:]
        with open(f"/sys/bus/pci/devices/{bus}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            return {"node": None, "note": "the platform reports no NUMA affinity for the GPU"}
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                lo, _, hi = part.partition("-")
                cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
        return {"node": node, "cpus": len(allowed)}
    except Exception as exc:  # best effort: a VM may not expose the topology
        return {"node": None, "note": f"{type(exc).__name__}: {exc}"[:120]}


def hbm_peak():
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        return float(json.load(open(peaks_path))["hbm_gbs"]), "MEASURED_PEAKS.json hbm_gbs (burst copy)"
    return FALLBACK_HBM_GBS, "fallback of B200_PROFILING.md"


class Workload:
    """One BASELINE config resident on one GPU: R rotating buffer sets of distinct, device-generated utterances with
    the embeddings their own segmentation calls for, and the step that runs the path over one set."""

    def __init__(self, torch, tok, name, rank, local_rank, rotate, depth=0, fused_amp=True, graphs=False, share_sms=None):
        from aat_b200 import synth
        from aat_b200.pipeline import TokenizerPipeline

        self.torch, self.name = torch, name
        self.B, self.N, self.D, idx = WORKLOADS[name]
        self.dev = torch.device("cuda", local_rank)
        self.batch = tok.plan([self.N] * self.B)
        B, D = self.B, self.D
        per_set = B * self.N * 4 + B * hubert_rows_upper_bound(self.N) * D * 4
        self.R = max(2, min(rotate, int(24e9 // per_set)))  # bounded HBM footprint for the long-form config
        self.wave_sets, self.emb_sets, self.n_seg, self.n_rows = [], [], [], []
        for s in range(self.R):
            # distinct utterances in every set and on every rank (seed = 1000 * config + utterance index)
            wave = synth.device_bursty_batch(self.batch, 1000 * (idx + 1), (rank * self.R + s) * B)
            self.batch.logmel(wave), self.batch.boundaries()  # one untimed pass fixes the segmentation, hence the embedding shape
            torch.cuda.synchronize()
            assert int(self.batch.status.min().item()) >= 0
            n_seg = int(self.batch.n_seg.item())
            n_rows = int(self.batch.seg_off[n_seg].item())
            assert n_rows == int(self.batch.n_frames.item())
            emb = torch.empty(n_rows, D, device=self.dev)
            synth.device_normal(emb, 1234 + 97 * rank + s)
            self.wave_sets.append(wave), self.emb_sets.append(emb), self.n_seg.append(n_seg), self.n_rows.append(n_rows)
        self.out = torch.empty(self.batch.total_seg_slots, D, device=self.dev)
        # the step runs through the public pipeline object: `depth` plans on `depth` streams, so that consecutive log-mel
        # kernels cover each other's ramp and tail, the boundary scans hide beside them and the pools share the SMs
        # (profiles/r2_step_timeline.txt; depth 1 = strictly serial, kept for the A/B)
        self.depth = depth = depth if depth > 0 else TokenizerPipeline.default_depth([self.N] * B)
        # with --graphs the pipelined schedule replays each slot's step from a CUDA graph (the rotating input buffers come
        # round again); the strictly serial one always launches kernel by kernel, as round 1's step did (and so that
        # kernels can be event-timed)
        self.pipes = {d: TokenizerPipeline(tok, [self.N] * B, D, depth=d, device=local_rank, fused_amp=fused_amp,
                                           graphs=d > 1 and graphs, share_sms=share_sms if d > 1 else None)
                      for d in sorted({1, depth})}
        self.audio_hours_per_step = B * self.N / 16000 / 3600
        self.pool_bytes = float(np.mean([r * D * 4 + s * D * 4 + (s + 1) * 8 for r, s in zip(self.n_rows, self.n_seg)]))

    def step(self, i, colsum=True, depth=None):
        s = i % self.R
        # results stay on the device in this loop (that is what `value` measures; `e2e` reads them back): no per-step event
        self.pipes[self.depth if depth is None else depth].submit(self.wave_sets[s], self.emb_sets[s], colsum=colsum,
                                                                   inputs_ready=True, record_done=False)

    def expected_sums(self, uses, share_sms=False):
        """Independent reduction of what `steps` steps must have accumulated: the pooled vectors of every buffer set
        are summed by torch in float64 and weighted with how often the set was used.  `share_sms`: pool with the grid
        the pipeline under test uses (the flag moves the CTA tile borders, hence the last bit of some means)."""
        torch = self.torch
        acc = torch.zeros(self.D + 1, dtype=torch.float64, device=self.dev)
        for s in range(self.R):
            if uses[s] == 0:
                continue
            self.batch.logmel(self.wave_sets[s]), self.batch.boundaries()
            self.batch.pool(self.emb_sets[s], self.out, share_sms=share_sms)
            torch.cuda.synchronize()
            n_seg = int(self.batch.n_seg.item())
            acc[: self.D] += uses[s] * self.out[:n_seg].double().sum(dim=0)
            acc[self.D] += uses[s] * n_seg
        return acc

    def pool_roofline(self, _cabi, sampled, sample_every, traffic):
        """Roofline of the pool kernel.  Back-to-back clock: 8 x R launches on the rotating sets inside one event pair."""
        torch, b = self.torch, self.batch
        peak, peak_src = hbm_peak()
        reps = 8 * self.R
        # per-set copies of the CSR, so that launch i pools set i % R with that set's own offsets
        csr = []
        for s in range(self.R):
            b.logmel(self.wave_sets[s]), b.boundaries()
            torch.cuda.synchronize()
            csr.append((b.seg_off.clone(), b._csr_totals.clone()))
        from aat_b200.pooling import _pool_device

        stream = b._stream()

        def launch(i):
            s = i % self.R
            seg_off, totals = csr[s]
            _pool_device(b.ctx, self.emb_sets[s], seg_off, int(self.out.shape[0]), totals, self.out, None, stream,
                         plan=b.handle, emb_ready=True)

        for i in range(2 * self.R):
            launch(i)
        torch.cuda.synchronize()
        best = None
        for _ in range(3):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(reps):
                launch(i)
            e1.record()
            torch.cuda.synchronize()
            us = e0.elapsed_time(e1) * 1e3 / reps
            best = us if best is None else min(best, us)
        achieved = self.pool_bytes / (best * 1e-6) / 1e9
        roof = {"kernel": "pool_kernel<float,1,false>" if self.D <= 1024 else "pool_kernel<float,k,false>", "bound": "hbm",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                "traffic_source": "dram__bytes_read.sum + dram__bytes_write.sum of one `ncu --set full` capture of this kernel "
                                  "at this workload (profiles/pool_traffic.json, profiles/r2_prof_pool*.txt); not measured "
                                  "in this run" if traffic is not None else None,
                "bytes_per_launch": self.pool_bytes, "us_per_launch": best, "launches_timed": reps,
                "method": f"{reps} back-to-back launches over the {self.R} rotating input sets inside one CUDA-event "
                          f"pair, best of 3 (no column-sum epilogue: the reduce kernel is a separate launch)",
                "peak_source": peak_src,
                "peak_note": "the peak is a COPY figure (read + write); this kernel is 96 % reads, and a pure read stream "
                             "through the same ring reaches 7.26 TB/s on this chip (profiles/r2_ubench_tma2d.txt), so "
                             "frac can pass 1 at the large configs without the bytes being wrong (traffic = ncu DRAM bytes)"}
        if sampled is not None:
            n, ms = sampled
            if n:
                us_s = ms / n * 1e3
                a = self.pool_bytes / (us_s * 1e-6) / 1e9
                roof["sampled_in_step"] = {"us_per_launch": us_s, "achieved": a, "frac": a / peak, "launches_timed": n,
                                           "method": f"event pair around every {sample_every}th pool launch of the timed "
                                                     f"region (an event in front of a kernel stops it overlapping its "
                                                     f"predecessor, and the pair itself costs several us)"}
        return roof


def measure_workload(torch, dist, w, steps, warmup, world, sample_every, no_colsum=False, sampler=None, depth=None,
                     breakdown=True):
    """Warm-up, then exactly `steps` timed steps (+ the allreduce of the dataset mean) between barriers; returns the
    timing record.  Device time by CUDA events, max over ranks."""
    from aat_b200 import _cabi

    dev = w.dev
    depth = w.depth if depth is None else depth
    pipe = w.pipes[depth]
    ctx_handle = w.batch.ctx.handle

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    if pipe.graphs:  # every (slot, buffer set) pair comes round twice in the warm-up: its graph is captured there
        warmup = max(warmup, 2 * math.lcm(depth, w.R) + 2)
    for i in range(warmup):
        w.step(i, not no_colsum, depth)
    barrier()
    if sampler is not None:
        t_spin = time.time()
        while not sampler.samples and time.time() - t_spin < 2.0:  # keep the GPU busy until nvidia-smi is up
            w.step(0, not no_colsum, depth)
            torch.cuda.synchronize()
    pipe.reset_sums()  # the dataset mean is the mean of the TIMED steps
    _cabi.profile_enable(ctx_handle, ("pool",), every=sample_every)
    launches0 = _cabi.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    wall0 = time.time()
    ev0.record()
    pipe.fork()  # the slots' streams start behind ev0
    host0 = time.perf_counter()
    for i in range(steps):
        w.step(i, not no_colsum, depth)
    host_us = (time.perf_counter() - host0) * 1e6 / steps  # what the host needs to enqueue a step (must stay below ms_per_step)
    dm = pipe.dataset_mean()  # joins the streams, adds the slots' sums, allreduce
    mean_vec = dm.result()
    ev1.record()
    barrier()
    wall1 = time.time()
    elapsed_ms = ev0.elapsed_time(ev1)
    launches = _cabi.launch_count() - launches0
    prof = _cabi.profile_summary(ctx_handle)
    _cabi.profile_enable(ctx_handle, ())
    t = torch.tensor([elapsed_ms, host_us], dtype=torch.float64, device=dev)
    by_rank = None
    if world > 1:  # every rank's own clock (the job's time is the slowest rank's)
        parts = [torch.empty_like(t) for _ in range(world)]
        dist.all_gather(parts, t)
        by_rank = {"ms_per_step": [float(p[0].item()) / steps for p in parts],
                   "host_enqueue_us_per_step": [float(p[1].item()) for p in parts]}
        elapsed_ms = max(float(p[0].item()) for p in parts)

    # ---- the collective's result against an independent reduction (A9): per-set pooled sums by torch, weighted by use
    # count, gathered from every rank and added in rank order (no NCCL reduction on the checking side)
    check = None
    if not no_colsum:
        uses = [len(range(s, steps, w.R)) for s in range(w.R)]
        mine = w.expected_sums(uses, share_sms=pipe.share_sms)
        if world > 1:
            parts = [torch.empty_like(mine) for _ in range(world)]
            dist.all_gather(parts, mine)
            want = torch.stack(parts).sum(dim=0)
        else:
            want = mine
        got = dm.acc
        assert int(got[w.D].item()) == int(want[w.D].item()), (got[w.D].item(), want[w.D].item())
        scale = want[: w.D].abs().max().clamp_min(1e-300)
        rel = float(((got[: w.D] - want[: w.D]).abs().max() / scale).item())
        assert rel <= 1e-12, f"dataset-mean sums differ from the independent reduction by {rel:.3e}"
        want_mean = (want[: w.D] / want[w.D]).float()
        assert bool(torch.isfinite(mean_vec).all())
        mean_err = float((mean_vec - want_mean).abs().max().item())
        assert mean_err <= 1e-6 * float(want_mean.abs().max().clamp_min(1e-30).item()) + 1e-12
        check = {"segments": int(want[w.D].item()), "max_rel_err_sums": rel, "max_abs_err_mean": mean_err,
                 "against": "torch float64 sums of the pooled vectors of every buffer set x use count, all_gather'ed and "
                            "added in rank order"}

    # ---- per-kernel breakdown (untimed extra pass, strictly serial, with every kernel instrumented)
    kernel_us = None
    if breakdown:
        w.pipes[1].reset_sums()
        _cabi.profile_enable(ctx_handle, _cabi.KERNEL_NAMES)
        for i in range(min(steps, 50)):
            w.step(i, not no_colsum, 1)
        torch.cuda.synchronize()
        kernel_us = {k: (ms / n * 1e3 if n else None) for k, (n, ms) in _cabi.profile_summary(ctx_handle).items()}
        _cabi.profile_enable(ctx_handle, ())
    return {"elapsed_ms": elapsed_ms, "value": world * w.audio_hours_per_step * steps / (elapsed_ms / 1e3),
            "ms_per_step": elapsed_ms / steps, "launches": int(launches), "sampled_pool": prof["pool"],
            "kernel_us": kernel_us, "wall": (wall0, wall1), "dataset_mean_check": check, "depth": depth,
            "host_enqueue_us_per_step": host_us, "by_rank": by_rank}


def schedule_note(depth):
    if depth == 1:
        return "one plan, one stream: log-mel -> boundaries -> pool strictly one after the other"
    return (f"aat_b200.pipeline.TokenizerPipeline, {depth} plans on {depth} streams: the log-mel kernels of a round run back "
            f"to back, the boundary scans beside them, the round's pool kernels share the SMs (one CTA per SM each)")


def pool_traffic(name):
    tpath = os.path.join(ROOT, "profiles", "pool_traffic.json")
    if os.path.exists(tpath):
        return json.load(open(tpath)).get(name)
    return None


def sub_record(torch, dist, tok, name, rank, local_rank, world, args):
    """configs[...] sub-record of another BASELINE config, same machinery as the headline at fewer steps."""
    from aat_b200 import _cabi

    w = Workload(torch, tok, name, rank, local_rank, args.rotate, args.depth, graphs=args.graphs)
    steps = {"c3": 200, "c4": 60, "c2": 500}[name]
    m = measure_workload(torch, dist, w, steps, 5, world, args.pool_sample_every)
    serial = measure_workload(torch, dist, w, steps, 5, world, args.pool_sample_every, depth=1, breakdown=False)
    roof = w.pool_roofline(_cabi, serial["sampled_pool"], args.pool_sample_every, pool_traffic(name))
    rec = {"workload": workload_name(name), "value": m["value"], "unit": UNIT, "steps": steps, "ms_per_step": m["ms_per_step"],
           "schedule": schedule_note(m["depth"]),
           "serial": {"value": serial["value"], "ms_per_step": serial["ms_per_step"], "schedule": schedule_note(1)},
           "kernel_us": m["kernel_us"], "roofline": roof, "segments_per_batch": w.n_seg, "hubert_frames_per_batch": w.n_rows,
           "rotating_sets": w.R, "dataset_mean_check": m["dataset_mean_check"]}
    del w
    torch.cuda.empty_cache()
    return rec


def run_c1(torch, tok):
    """BASELINE configs[0]: one 10 s clip through the reference-facing numpy API (host buffers in and out, one C-ABI
    call each), with the oracle port on one host core beside it."""
    from aat_b200 import AudioWaveform, mean_pool_segments, synth
    from oracle import ref_port  # this config's cpu_baseline leg: the port is timed beside the product and checks it

    n, dim = 160_000, 768
    wave = synth.bursty_speech(n, synth.seed_for(1, 0))
    awf = AudioWaveform(wave, 16000)
    segments, mel = tok.tokenize(awf)
    lengths = [s.waveform.shape[-1] for s in segments]
    g = torch.Generator().manual_seed(0)
    embs = [torch.randn(1, int(f), dim, generator=g) for f in synth.hubert_frames(lengths)]
    ref = ref_port.RefTokenizer()
    assert ref.segment_lengths(wave)[0] == lengths  # parity of what is being timed

    def best_ms(fn, reps):
        fn()
        best = float("inf")
        for _ in range(reps):
            t0 = time.perf_counter()
            fn()
            best = min(best, time.perf_counter() - t0)
        return best * 1e3

    gpu_tok = best_ms(lambda: tok.tokenize(awf), 30)
    gpu_pool = best_ms(lambda: mean_pool_segments(embs), 30)
    gpu_min = best_ms(lambda: tok.find_amplitude_minimas(mel), 30)
    cpu_tok = best_ms(lambda: ref.segment_lengths(wave), 3)
    cpu_pool = best_ms(lambda: ref_port.mean_pool_segments(embs), 10)
    cpu_min = best_ms(lambda: ref.find_amplitude_minimas(mel), 10)
    hours = n / 16000 / 3600
    return {"workload": "configs[0]: single 10 s 16 kHz clip, batch 1, 768-d embeddings, numpy API (host in / host out)",
            "value": hours / ((gpu_tok + gpu_pool) * 1e-3), "unit": UNIT, "segments": len(lengths),
            "latency_ms": {"tokenize": gpu_tok, "mean_pool_segments(list)": gpu_pool, "find_amplitude_minimas": gpu_min},
            "cpu_port_latency_ms": {"tokenize": cpu_tok, "mean_pool_segments(list)": cpu_pool, "find_amplitude_minimas": cpu_min},
            "cpu_port_value": hours / ((cpu_tok + cpu_pool) * 1e-3), "method": "best of 30 calls, wall clock, one thread"}


def run_c5(torch, dist, tok, rank, local_rank, world, args):
    """BASELINE configs[4]: >= 1000 synthetic audio-hours of DISTINCT 16 s utterances sharded over the N ranks.  Every
    rank generates its shard on the device chunk by chunk (counter-based generator; generation is outside the timed
    brackets and reported beside them), runs the path over every batch of the chunk (timed, CUDA events), accumulates
    the column sums on the device, and the ranks meet once in the allreduce of the dataset-mean embedding."""
    from aat_b200 import synth
    from aat_b200.pipeline import TokenizerPipeline
    from aat_b200.pooling import DatasetMean

    B, N, D, _ = WORKLOADS["c2"]
    dev = torch.device("cuda", local_rank)
    hours_per_batch = B * N / 16000 / 3600
    total_batches = math.ceil(args.c5_hours / hours_per_batch)
    per_rank = math.ceil(total_batches / world)
    chunk = min(64, per_rank)  # 4096 utterances: 4.2 GB of waveforms + 10 GB of embeddings per chunk
    rows_ub = B * hubert_rows_upper_bound(N)
    batch = tok.plan([N] * B)  # generator layout + the audit's plan
    # fresh buffers every chunk: nothing to replay, the steps are launched kernel by kernel
    pipe = TokenizerPipeline(tok, [N] * B, D, depth=args.depth if args.depth > 0 else None, device=local_rank, graphs=False)
    waves = torch.empty(chunk, batch.total_samples, dtype=torch.float32, device=dev)
    embs = torch.empty(chunk, rows_ub, D, dtype=torch.float32, device=dev)
    out = torch.empty(batch.total_seg_slots, D, device=dev)
    status_min = torch.zeros(1, dtype=torch.int32, device=dev)
    gen_ms = run_ms = 0.0
    first_batch = rank * per_rank  # global batch index of this rank's shard

    def generate(c0, n):
        for k in range(n):
            g = first_batch + c0 + k
            synth.device_bursty_batch(batch, 5000, g * B, out=waves[k])  # seed = 1000 * config + utterance index
            synth.device_normal(embs[k], 7_000_000 + g)

    def consume(n):
        pipe.fork()  # the slots' streams start behind the generation (and behind the timing event)
        for k in range(n):
            # embeddings are an allocation of the upper-bound row count: the rows the segments cover are read on the device
            pipe.submit(waves[k], embs[k], rows_from_device=True, inputs_ready=True, record_done=False)
        pipe.join()

    # warm-up on the first chunk's first batches (untimed), then reset the accumulators
    generate(0, min(chunk, 6))
    consume(min(chunk, 6))
    torch.cuda.synchronize()
    pipe.reset_sums()
    audit = None
    if world > 1:
        dist.barrier()
    done = 0
    while done < per_rank:
        n = min(chunk, per_rank - done)
        e = [torch.cuda.Event(enable_timing=True) for _ in range(3)]
        e[0].record()
        generate(done, n)
        e[1].record()
        consume(n)
        for slot in pipe.slots:
            torch.minimum(status_min, slot.batch.status.min().reshape(1), out=status_min)  # the slots' last batches
        e[2].record()
        torch.cuda.synchronize()
        gen_ms += e[0].elapsed_time(e[1])
        run_ms += e[1].elapsed_time(e[2])
        if done == 0:
            # audit of the first chunk (untimed): the same batches again, one by one, every pooled vector summed by torch
            chk = DatasetMean(D, device=local_rank)
            ref = torch.zeros(D + 1, dtype=torch.float64, device=dev)
            for k in range(min(n, 16)):
                batch.logmel(waves[k]), batch.boundaries()
                batch.pool(embs[k], out, colsum=chk.running_buffer(), accumulate=True, rows_from_device=True)
                torch.cuda.synchronize()
                s = int(batch.n_seg.item())
                ref[:D] += out[:s].double().sum(dim=0)
                ref[D] += s
            rel = float(((chk.acc[:D] - ref[:D]).abs().max() / ref[:D].abs().max()).item())
            assert int(chk.acc[D].item()) == int(ref[D].item()) and rel <= 1e-12, rel
            audit = {"batches": min(n, 16), "segments": int(ref[D].item()), "max_rel_err_sums": rel}
        done += n
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    dm = pipe.dataset_mean()  # adds the slots' sums, ONE allreduce
    mean_vec = dm.result()
    ev1.record()
    torch.cuda.synchronize()
    run_ms += ev0.elapsed_time(ev1)
    assert int(status_min.item()) >= 0
    t = torch.tensor([run_ms, gen_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    run_ms, gen_ms = float(t[0].item()), float(t[1].item())
    hours = world * per_rank * hours_per_batch
    segments = int(dm.acc[D].item())
    utterances = world * per_rank * B
    assert bool(torch.isfinite(mean_vec).all()) and 20 * utterances <= segments <= 60 * utterances
    rec = {"workload": f"configs[4]: {hours:.1f} synthetic audio-hours = {utterances} distinct 16 s utterances sharded over "
                       f"{world} GPU(s), {D}-d embeddings, NCCL allreduce of the dataset-mean embedding",
           "audio_hours_processed": hours, "utterances": utterances, "segments": segments, "value": hours / (run_ms / 1e3),
           "unit": UNIT, "device_seconds_path": run_ms / 1e3, "device_seconds_generation": gen_ms / 1e3,
           "batches_per_rank": per_rank, "chunk_batches": chunk, "scaling": "strong (fixed 1000 audio-hours)",
           "schedule": schedule_note(len(pipe.slots)),
           "timing": "sum over chunks of CUDA-event brackets around the path (generation bracketed separately), + the "
                     "allreduce and finalisation; max over ranks",
           "dataset_mean": {"norm": float(mean_vec.norm().item()), "audit_first_chunk": audit}}
    del waves, embs
    torch.cuda.empty_cache()
    return rec


def run_b200(args):
    import torch
    import torch.distributed as dist

    from aat_b200 import AdaptiveAudioAmplitudeTokenizer, _cabi

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    assert torch.cuda.is_available(), "bench.py needs a CUDA device (there is no CPU fallback)"
    numa = bind_to_gpu_numa_node(local_rank)  # before any pinned allocation
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        # NCCL prints its version banner to stdout when the communicator is created; the contract is ONE
        # JSON line on stdout, so park stdout on stderr until the first collective has run.
        sys.stdout.flush()
        saved = os.dup(1)
        os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            warm = torch.zeros(1, device=dev)
            dist.all_reduce(warm)
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush()
            os.dup2(saved, 1)
            os.close(saved)

    tok = AdaptiveAudioAmplitudeTokenizer(device=local_rank)
    w = Workload(torch, tok, args.workload, rank, local_rank, args.rotate, args.depth, not args.unfused_amp, args.graphs,
                 False if args.no_share_sms else None)
    sampler = ClockSampler(local_rank)
    sampler.start()
    m = measure_workload(torch, dist, w, args.steps, args.warmup, world, args.pool_sample_every, args.no_colsum, sampler)
    clocks = sampler.stop(*m["wall"])
    serial = m
    if w.depth != 1:  # same-run A/B: the strictly serial schedule (round 1's step)
        serial = measure_workload(torch, dist, w, max(50, args.steps // 4), 5, world, args.pool_sample_every, args.no_colsum,
                                  depth=1, breakdown=False)
    roofline = w.pool_roofline(_cabi, serial["sampled_pool"], args.pool_sample_every, pool_traffic(args.workload))

    # ---- e2e: same step through the public batched API with host (pinned) buffers
    e2e = None
    if not args.no_e2e:
        e2e = run_e2e(torch, dev, w, args, world)
        e2e["numa"] = numa

    cfg = workload_config(args.workload, world, args.rotate)
    line = {
        "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": m["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64 (log-mel) / f32 (boundaries, pool)", "data": "synthetic", "config": cfg,
        "workload_detail": {"segments_per_batch": w.n_seg, "hubert_frames_per_batch": w.n_rows, "rotating_sets": w.R,
                            "resident_mb": sum(x.numel() * 4 for x in w.wave_sets + w.emb_sets) / 1e6,
                            "generator": "aat_synth_waveforms / aat_synth_normal (Philox 4x32-10, on the device)"},
        "clocks": clocks, "e2e": e2e, "gpu_launches": m["launches"], "roofline": roofline,
        "kernel_us": m["kernel_us"], "dataset_mean_check": m["dataset_mean_check"],
        "schedule": schedule_note(m["depth"]), "host_enqueue_us_per_step": m["host_enqueue_us_per_step"],
        "by_rank": m["by_rank"],
        "serial": {"value": serial["value"], "ms_per_step": serial["ms_per_step"], "schedule": schedule_note(1)},
    }
    if not args.no_configs and args.workload == "c2":
        configs = {}
        for name, fn in (("c3", lambda: sub_record(torch, dist, tok, "c3", rank, local_rank, world, args)),
                         ("c4", lambda: sub_record(torch, dist, tok, "c4", rank, local_rank, world, args)),
                         ("c5", lambda: run_c5(torch, dist, tok, rank, local_rank, world, args))):
            # a failure is fatal on purpose: a silent partial record would read as coverage
            configs[name] = fn()
        if rank == 0:
            configs["c1"] = run_c1(torch, tok)
        line["configs"] = configs
    if rank == 0:
        if world == 1 and not args.no_cpu_baseline:
            _, t_probe = cpu_reference_throughput(args.workload, 8)
            n_sample = int(max(8, min(4096, 12.0 / max(t_probe / 8, 1e-4))))  # ~12 s of CPU work
            v, dt = cpu_reference_throughput(args.workload, n_sample)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": 1, "kind": "port",
                                    "sample": f"{n_sample} utterances of the workload (8 distinct, cycled), single "
                                              f"process (datasets.map without num_proc), {dt:.1f} s of CPU work; waveforms "
                                              f"and embeddings resident before the clock starts",
                                    "host": host_description()}
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


def measure_copy_peak(torch, dev, h2d_bytes, d2h_bytes, world):
    """Plain pinned-memory copies of the e2e step's byte counts, H2D and D2H on two streams at once, all ranks at the
    same time: what the host side of the box sustains, i.e. the roofline of the e2e figure."""
    import torch.distributed as dist

    src = torch.empty(h2d_bytes, dtype=torch.uint8).pin_memory()
    dst = torch.empty(h2d_bytes, dtype=torch.uint8, device=dev)
    back_d = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8, device=dev)
    back_h = torch.empty(max(d2h_bytes, 1), dtype=torch.uint8).pin_memory()
    s_in, s_out = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)
    reps = 10

    def once():
        with torch.cuda.stream(s_in):
            dst.copy_(src, non_blocking=True)
        with torch.cuda.stream(s_out):
            back_h.copy_(back_d, non_blocking=True)

    for _ in range(3):
        once()
    torch.cuda.synchronize()
    best = None
    for _ in range(3):  # best of three batches: the first copies after an allocation run below the sustained rate
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        for _ in range(reps):
            once()
        torch.cuda.synchronize()
        dt = time.perf_counter() - t0
        best = dt if best is None else min(best, dt)
    t = torch.tensor([best], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item()) / reps  # seconds per step's worth of copies


def run_e2e(torch, dev, w, args, world):
    """Host buffers in, host buffers out: every step copies its waveforms and embeddings from pinned host
    memory, runs the path, and reads segment lengths and pooled vectors back.  Three streams: H2D of step i + 1,
    kernels of step i and D2H of step i - 1 overlap (throughput metric over K steps)."""
    import torch.distributed as dist

    batch, D = w.batch, w.D
    host_wave = w.wave_sets[0].cpu().pin_memory()
    host_emb = w.emb_sets[0].cpu().pin_memory()
    n_seg = w.n_seg[0]
    n_slots = batch.total_seg_slots
    bufs = []
    for _ in range(2):
        bufs.append({
            "wave": torch.empty_like(w.wave_sets[0]), "emb": torch.empty_like(w.emb_sets[0]),
            "out": torch.empty(n_slots, D, device=dev), "seg_len": torch.empty_like(batch.seg_len),
            "seg_count": torch.empty_like(batch.seg_count),
            "pooled_host": torch.empty((n_seg, D), dtype=torch.float32).pin_memory(),
            "len_host": torch.empty(n_slots, dtype=torch.int64).pin_memory(),
            "count_host": torch.empty(batch.n_utts, dtype=torch.int32).pin_memory(),
            "in_free": None, "out_free": None,
        })
    h2d = host_wave.numel() * 4 + host_emb.numel() * 4
    d2h = n_seg * D * 4 + n_slots * 8 + batch.n_utts * 4
    s_in, s_run, s_out = (torch.cuda.Stream(device=dev) for _ in range(3))

    def e2e_step(i):
        b = bufs[i % 2]
        with torch.cuda.stream(s_in):
            if b["in_free"] is not None:
                s_in.wait_event(b["in_free"])  # the kernels that read this input pair two steps ago are done
            b["wave"].copy_(host_wave, non_blocking=True)
            b["emb"].copy_(host_emb, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record()
        with torch.cuda.stream(s_run):
            s_run.wait_event(ready)
            if b["out_free"] is not None:
                s_run.wait_event(b["out_free"])  # the D2H that read this output pair two steps ago is done
            batch.logmel(b["wave"]), batch.boundaries()
            batch.pool(b["emb"], b["out"], emb_ready=True)
            # the plan's segment buffers are rewritten by the next step: hand the D2H stream its own copy (device to
            # device, 0.2 MB) so that result copies never hold the next step's kernels back
            b["seg_len"].copy_(batch.seg_len, non_blocking=True)
            b["seg_count"].copy_(batch.seg_count, non_blocking=True)
            b["in_free"] = torch.cuda.Event()
            b["in_free"].record()
        with torch.cuda.stream(s_out):
            s_out.wait_event(b["in_free"])
            b["pooled_host"].copy_(b["out"][:n_seg], non_blocking=True)
            b["len_host"].copy_(b["seg_len"], non_blocking=True)
            b["count_host"].copy_(b["seg_count"], non_blocking=True)
            b["out_free"] = torch.cuda.Event()
            b["out_free"].record()

    steps = max(4, min(args.steps, 40))
    for i in range(2):
        e2e_step(i)
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    t0 = time.perf_counter()
    for i in range(steps):
        e2e_step(i)
    torch.cuda.synchronize()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    wall = float(t.item())
    assert int(bufs[0]["count_host"].sum()) == n_seg
    # parity of the host-side result with the device-resident path on the same input set
    batch.logmel(w.wave_sets[0]), batch.boundaries()
    batch.pool(w.emb_sets[0], w.out)
    torch.cuda.synchronize()
    assert torch.equal(bufs[0]["pooled_host"], w.out[:n_seg].cpu())
    copy_s = measure_copy_peak(torch, dev, h2d, d2h, world)
    step_s = wall / steps
    return {"value": world * w.audio_hours_per_step * steps / wall, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
            "d2h_bytes_per_step": int(d2h), "steps": steps, "ms_per_step": 1e3 * step_s,
            "copy_peak": {"ms_per_step": 1e3 * copy_s, "h2d_gbs_per_gpu": h2d / copy_s / 1e9,
                          "how": "the step's H2D and D2H byte counts as plain pinned-memory copies on two streams, all "
                                 "ranks at once, same run"},
            "frac_of_copy_peak": copy_s / step_s,
            "api": "PackedBatch.logmel/boundaries/pool on pinned host tensors; H2D, kernels and D2H on three streams "
                   "(2 buffer pairs)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=None)
    ap.add_argument("--warmup", type=int, default=None)
    ap.add_argument("--impl", choices=["b200", "reference"], default="b200")
    ap.add_argument("--workload", choices=sorted(WORKLOADS), default="c2")
    ap.add_argument("--rotate", type=int, default=DEFAULT_ROTATE)
    ap.add_argument("--depth", type=int, default=0,
                    help="batches in flight (plans x streams) of the step's pipeline; 1 = serial, 0 = the pipeline's own default")
    ap.add_argument("--no-share-sms", action="store_true", help="A/B: pool kernels of the pipelined schedule with the full grid (two CTAs per SM) instead of AAT_POOL_SHARE_SMS")
    ap.add_argument("--unfused-amp", action="store_true", help="amplitude curve by the separate pass (aat_amplitude) instead of the log-mel kernel's epilogue")
    ap.add_argument("--graphs", action="store_true",
                    help="replay the pipelined steps from CUDA graphs (14 instead of 32-100 us of host time per step, 3 %% more "
                         "device time: for hosts whose launch path cannot keep up)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true", help="skip the host-buffer leg (profiling runs)")
    ap.add_argument("--no-configs", action="store_true", help="skip the c1/c3/c4/c5 sub-records (profiling runs)")
    ap.add_argument("--no-colsum", action="store_true", help="diagnostic: pool without the column-sum epilogue")
    ap.add_argument("--c5-hours", type=float, default=C5_AUDIO_HOURS, help="audio-hours of the dataset job (configs[4])")
    ap.add_argument("--pool-sample-every", type=int, default=8, help="event-time every n-th pool launch of the timed region")
    args = ap.parse_args()
    if args.impl == "reference":
        args.steps = 3 if args.steps is None else args.steps
        args.warmup = 1 if args.warmup is None else args.warmup
        run_reference(args)
    else:
        args.steps = 2000 if args.steps is None else args.steps
        args.warmup = max(3, 10 if args.warmup is None else args.warmup)
        run_b200(args)


if __name__ == "__main__":
    main()
