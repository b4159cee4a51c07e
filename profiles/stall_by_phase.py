"""Warp-stall samples of the log-mel kernel by PHASE of the tile loop (run in the build container, no GPU needed):

    python profiles/stall_by_phase.py > profiles/r2_logmel_stall_by_phase.txt

Joins `ncu --page source` (per-SASS-instruction stall samples of gpurun_out/prof_logmel_r2.ncu-rep, captured with
--import-source on) with `nvdisasm -g` (source line of every SASS instruction of the built library).  An instruction that
belongs to an inlined helper (dft20, band, fast_log10, cp_async ...) is attributed to the phase of the kernel-body line
that precedes it in the SASS.  AAT_B200_LIB / AAT_LOGMEL_SRC = the library build the report was captured with and its logmel.cu (the line info must
match; r2: `git show f73c775:audio-adaptive-tokenizer_b200/csrc/logmel.cu`)."""
import collections
import csv
import io
import os
import re
import subprocess
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.environ.get("AAT_B200_LIB") or os.path.join(ROOT, "audio-adaptive-tokenizer_b200", "aat_b200", "libaat_b200.so")  # the build the report was captured with
REP = os.path.join(ROOT, "gpurun_out", sys.argv[1] if len(sys.argv) > 1 else "prof_logmel_r2.ncu-rep")
KERNEL = "logmel_kernelIfLb1ELb0E"  # <float, hop 160, no z-score>: the bench's instantiation


def phases_from_source():
    """(first line, name) of every phase of the kernel body, from the section comments of logmel.cu."""
    src = open(os.environ.get("AAT_LOGMEL_SRC") or os.path.join(ROOT, "audio-adaptive-tokenizer_b200", "csrc", "logmel.cu")).read().splitlines()
    marks = []
    for i, line in enumerate(src, 1):
        for pat, name in ((r"__global__ void __launch_bounds__\(kThreads, 3\) logmel_kernel", "kernel entry (shared-memory layout, indices)"),
                          (r"auto fetch_desc = ", "tile fetch: descriptors and raw samples by cp.async (issued behind pass 1's barrier)"),
                          (r"// ---- prologue: descriptors of the first two tiles", "prologue (tables, first descriptors), once per CTA"),
                          (r"while \(tile_id < p\.n_tiles\)", "loop top: wait for samples, CTA barrier, tile counter"),
                          (r"---- pass 1:", "pass 1: window, 20-point transforms over n1, twiddles, exchange stores"),
                          (r"\*s_next = 3 \* G \+ grabbed", "barrier after pass 1, next tile's fetch issued"),
                          (r"---- pass 2 \+ split:", "pass 2: exchange loads, 20-point transforms over n2, mirror exchange"),
                          (r"double pa\[11\], pb\[11\];", "split, float32 rounding (Veltkamp), power, overlay barrier + stores"),
                          (r"---- mel projection, floor, log10", "mel projection (banded taps)"),
                          (r"for \(int i0 = 0; i0 < kMaxPerThread; i0 \+= kBatch\)", "log10 batches + float32 stores"),
                          (r"---- fused amplitude curve", "amplitude epilogue (barrier + 64-term float32 chain)"),
                          (r"cp_async_wait<0>\(\);\s*$", None)):
            if name and re.search(pat, line):
                marks.append((i, name))
    return sorted(set(marks))


def sass_lines():
    tmp = tempfile.mkdtemp()
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(LIB)], cwd=tmp, stdout=subprocess.DEVNULL, stderr=subprocess.DEVNULL)
    cubin = sorted(f for f in os.listdir(tmp) if f.startswith("logmel") and f.endswith(".cubin"))[0]
    dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(tmp, cubin)], stdout=subprocess.PIPE, text=True).stdout
    out, line, inside = {}, None, False
    for l in dis.splitlines():
        if l.startswith(".text.") and l.rstrip().endswith(":"):
            inside = KERNEL in l and "$" not in l
            continue
        if not inside:
            continue
        m = re.search(r'//## File ".*logmel\.cu", line (\d+)', l)
        if m:
            line = int(m.group(1))
            continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
        if m:
            out[int(m.group(1), 16)] = (line, m.group(2).strip())
    return out


def main():
    marks = phases_from_source()
    body_first = marks[0][0]
    lines = sass_lines()
    txt = subprocess.run(["ncu", "-i", REP, "--page", "source", "--csv"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
    rows = list(csv.reader(io.StringIO(txt)))
    hdr = rows[1]
    ix = {h: i for i, h in enumerate(hdr)}
    data = rows[2:]
    base = int(data[0][ix["Address"]], 16) if data[0][ix["Address"]].startswith("0x") else int(data[0][ix["Address"]])
    stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
    per = collections.OrderedDict((name, collections.Counter()) for _, name in marks)
    per["(not attributed)"] = collections.Counter()
    phase, in_loop = marks[0][1], False
    total = 0
    for r in data:
        a = r[ix["Address"]]
        off = (int(a, 16) if a.startswith("0x") else int(a)) - base
        line = lines.get(off, (None, ""))[0]
        if line is not None and line >= body_first:  # a kernel-body line: (re)locate the phase
            located = [name for first, name in marks if first <= line][-1]
            # the entry and prologue lines (shared-memory layout, thread indices, table pointers) are rematerialised
            # all over the loop: they only count as entry / prologue before the loop has been entered
            if not (in_loop and (located == marks[0][1] or located.startswith("prologue"))):
                phase = located
            in_loop = in_loop or located.startswith("loop top")
        n = int(r[ix["# Samples"]] or 0)
        total += n
        c = per[phase]
        if os.environ.get("AAT_STALL_DUMP") and phase.startswith(tuple(os.environ["AAT_STALL_DUMP"].split("|"))) and n >= 8:
            print(f"    {off:6x} line {line} {n:5d} samples  {lines.get(off, (None, ''))[1][:70]}", file=sys.stderr)
        c["samples"] += n
        c["instructions"] += int(r[ix["Instructions Executed"]] or 0)
        for h in stall_cols:
            c[h[6:]] += int(r[ix[h]] or 0)
    tot_inst = sum(c["instructions"] for c in per.values())
    print(f"# {os.path.basename(REP)}: warp-stall samples of logmel_kernel<float, hop 160> by phase of the tile loop ({total} samples);")
    print("# share of samples | share of executed instructions | the three largest stall reasons of the phase (share of its samples)")
    for name, c in per.items():
        if not c["samples"]:
            continue
        top = sorted(((v, k) for k, v in c.items() if k not in ("samples", "instructions")), reverse=True)[:3]
        tops = ", ".join(f"{k} {100 * v / c['samples']:.0f} %" for v, k in top)
        print(f"{100 * c['samples'] / total:5.1f} %  {100 * c['instructions'] / tot_inst:5.1f} %  {name:75s} {tops}")


if __name__ == "__main__":
    main()
