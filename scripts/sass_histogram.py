"""Per-kernel SASS opcode histograms of libusv_b200.so -> profiles/<tag>_sass_histogram.{json,md}.

  python scripts/sass_histogram.py [tag]        (default tag: r2)

Evidence that the hand-written paths are what the library contains: UTCIMMA / LDTM / UTCBAR (tcgen05 + TMEM),
IMMA (mma.sync), VABSDIFF4 / IDP.4A (packed-byte ALU), LDGSTS (cp.async), UTMALDG (TMA), SYNCS (mbarrier),
per kernel symbol (cuobjdump -sass on the sm_100a cubin). Runs on the build host: no GPU needed."""
import collections
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "unsynchronized_stereo_vision_proj325_b200", "libusv_b200.so")
KEYS = ["UTCIMMA", "UTCHMMA", "LDTM", "UTCBAR", "IMMA", "VABSDIFF4", "IDP", "LDGSTS", "UTMALDG", "SYNCS", "DFMA", "DMUL", "DADD", "DSETP",
        "IMAD", "VIMNMX", "SHFL", "LDS", "STS", "BAR", "SEL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def main():
    tag = sys.argv[1] if len(sys.argv) > 1 else "r2"
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    kernels, cur = collections.OrderedDict(), None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = m.group(1)
            kernels[cur] = collections.Counter()
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)((?:\.[A-Z0-9_x]+)*)", line)
        if m and cur:
            kernels[cur][m.group(1) + m.group(2)] += 1
    names = demangle(list(kernels))
    rec = {}
    for k, c in kernels.items():
        short = re.sub(r"\(.*", "", names.get(k, k))
        short = re.sub(r"^void ", "", short)
        total = sum(c.values())
        fam = collections.Counter()
        for op, n in c.items():
            base = op.split(".")[0]
            if base in KEYS:
                fam[base] += n
        detail = {op: n for op, n in c.items() if op.split(".")[0] in ("UTCIMMA", "LDTM", "IMMA", "VABSDIFF4", "IDP", "LDGSTS", "UTMALDG", "UTCBAR")}
        rec.setdefault(short, []).append({"symbol": k, "instructions": total, "families": dict(fam), "detail": detail})
    os.makedirs(os.path.join(ROOT, "profiles"), exist_ok=True)
    json.dump(rec, open(os.path.join(ROOT, "profiles", tag + "_sass_histogram.json"), "w"), indent=1, sort_keys=True)
    cols = ["UTCIMMA", "LDTM", "IMMA", "VABSDIFF4", "IDP", "LDGSTS", "UTMALDG", "SYNCS", "DFMA", "IMAD", "VIMNMX", "SHFL", "SEL"]
    lines = ["# SASS opcode histogram of libusv_b200.so (%s), cuobjdump -sass, static instruction counts per kernel" % tag, "",
             "Instantiations of one template are summed; `n` = number of instantiations. Produced by `scripts/sass_histogram.py`.", "",
             "| kernel | n | instr | " + " | ".join(cols) + " |", "|---|---|---|" + "---|" * len(cols)]
    tot = collections.Counter()
    for short in sorted(rec):
        fam = collections.Counter()
        for r in rec[short]:
            fam.update(r["families"])
        tot.update(fam)
        lines.append("| `%s` | %d | %d | %s |" % (short.replace("usv::", "").split("<")[0], len(rec[short]), sum(r["instructions"] for r in rec[short]),
                                               " | ".join(str(fam.get(c, 0)) for c in cols)))
    lines.append("| **library** | %d | %d | %s |" % (sum(len(v) for v in rec.values()), sum(r["instructions"] for v in rec.values() for r in v),
                                                   " | ".join(str(tot.get(c, 0)) for c in cols)))
    open(os.path.join(ROOT, "profiles", tag + "_sass_histogram.md"), "w").write("\n".join(lines) + "\n")
    print("\n".join(lines[-12:]))


if __name__ == "__main__":
    main()
