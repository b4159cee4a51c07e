"""SASS evidence of what the kernels use, from the built library (run in the build container, no GPU needed):

    python profiles/sass_evidence.py          # writes profiles/sass_pool.txt, profiles/sass_logmel.txt, profiles/sass_summary.txt

`cuobjdump -sass` of libaat_b200.so, per kernel: the mnemonic histogram and the excerpts that show the mechanisms the
design relies on — UBLKCP (cp.async.bulk, the TMA engine without a tensor map) + SYNCS (mbarrier) in the pool kernel,
the DFMA/DADD/DMUL mix + LDGSTS (cp.async) staging of the log-mel kernel, and the absence of tensor-core (UTC*MMA / HMMA)
and tensor-map (UTMALDG/UTMASTG) instructions anywhere: nothing on this path is a dense contraction."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "audio-adaptive-tokenizer_b200", "aat_b200", "libaat_b200.so")


def kernels():
    txt = subprocess.run(["cuobjdump", "-sass", LIB], stdout=subprocess.PIPE, text=True, check=True).stdout
    arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
    out, name, body = {}, None, []
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            if name:
                out[name] = body
            name, body = m.group(1), []
        elif name:
            body.append(line)
    if name:
        out[name] = body
    return arch, out


def demangle(name):
    try:
        return subprocess.run(["cu++filt", name], stdout=subprocess.PIPE, text=True).stdout.strip() or name
    except OSError:
        return name


def mnemonics(body):
    c = collections.Counter()
    for line in body:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m:
            c[m.group(1).split(".")[0]] += 1
    return c


def excerpt(body, pattern, context=2, limit=3):
    hits = [i for i, l in enumerate(body) if re.search(pattern, l)]
    out = []
    for i in hits[:limit]:
        out += [re.sub(r"\s+/\* 0x[0-9a-f]+ \*/\s*$", "", l) for l in body[max(0, i - context): i + context + 1] if l.strip()] + ["        ..."]
    return out


def main():
    arch, ks = kernels()
    names = {k: demangle(k) for k in ks}
    summary = [f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)}: cubin architectures {arch}",
               f"{'kernel':72s} {'instr':>6s} {'UBLKCP':>7s} {'SYNCS':>6s} {'LDGSTS':>7s} {'DFMA':>6s} {'DADD':>6s} {'DMUL':>6s} {'FADD':>6s} "
               f"{'UTMA*':>6s} {'UTC*MMA':>8s} {'HMMA':>5s}"]
    tot = collections.Counter()
    for k, body in sorted(ks.items(), key=lambda kv: names[kv[0]]):
        c = mnemonics(body)
        tot.update(c)
        short = re.sub(r"\(bool\)", "", re.sub(r"aat::<unnamed>::|aat::\(anonymous namespace\)::|void |\((aat::|int|float|const|unsigned|long).*", "", names[k]))[:72]
        utma = sum(v for m, v in c.items() if m.startswith("UTMA"))
        utc = sum(v for m, v in c.items() if m.startswith("UTC"))
        summary.append(f"{short:72s} {sum(c.values()):6d} {c['UBLKCP']:7d} {c['SYNCS']:6d} {c['LDGSTS']:7d} {c['DFMA']:6d} {c['DADD']:6d} "
                       f"{c['DMUL']:6d} {c['FADD']:6d} {utma:6d} {utc:8d} {c['HMMA']:5d}")
    summary.append(f"# whole library: UBLKCP {tot['UBLKCP']}, SYNCS {tot['SYNCS']}, LDGSTS {tot['LDGSTS']}, DFMA {tot['DFMA']}, "
                   f"UTMALDG/UTMASTG {sum(v for m, v in tot.items() if m.startswith('UTMA'))}, "
                   f"UTC*MMA {sum(v for m, v in tot.items() if m.startswith('UTC'))}, HMMA/IMMA/DMMA "
                   f"{tot['HMMA'] + tot['IMMA'] + tot['DMMA']}")
    open(os.path.join(ROOT, "profiles", "sass_summary.txt"), "w").write("\n".join(summary) + "\n")

    def one(pattern, path, title, excerpts):
        k = next(k for k in ks if re.search(pattern, names[k]))
        c = mnemonics(ks[k])
        lines = [f"# {title}", f"# {names[k]}", "# mnemonic histogram (top 24): " + ", ".join(f"{m} {v}" for m, v in c.most_common(24)), ""]
        for label, pat in excerpts:
            lines.append(f"# --- {label}")
            lines += excerpt(ks[k], pat)
            lines.append("")
        open(os.path.join(ROOT, "profiles", path), "w").write("\n".join(lines) + "\n")

    one(r"pool_kernel<float, *\(?i?n?t?\)?1, *(false|\(bool\)0)>", "sass_pool.txt",
        "pool_kernel<float, 1, false>: the embedding stream moves by 1-D bulk async copies tracked by mbarrier transaction counts",
        [("bulk copy global -> shared with mbarrier completion (cp.async.bulk ... mbarrier::complete_tx::bytes)", r"UBLKCP"),
         ("mbarrier arrive / expect_tx / try_wait", r"SYNCS"),
         ("128-bit shared loads of the consumers (one 16-byte column slab per thread and row)", r"LDS\.128"),
         ("release / acquire of the cross-CTA carry flags", r"\.STRONG\.GPU|ST\.E\.STRONG|LD\.E\.STRONG")])
    one(r"logmel_kernel<float, *(true|\(bool\)1), *(false|\(bool\)0)>", "sass_logmel.txt",
        "logmel_kernel<float, true, false>: FP64 transform, cp.async staging",
        [("cp.async staging of raw samples and tile descriptors", r"LDGSTS"),
         ("FP64 butterflies of the 20-point transforms", r"DFMA"),
         ("float -> double of the raw samples (the one conversion left on the XU path)", r"F2F\.F64\.F32"),
         ("tile scheduler", r"ATOMG|ATOM\.")])
    print("\n".join(summary))


if __name__ == "__main__":
    main()
