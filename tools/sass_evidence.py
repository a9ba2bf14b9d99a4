"""Writes profiles/<round>_sass_tcgen05_tma.txt: per kernel of libhowl_b200.so, the count of the SASS mnemonics that prove tcgen05 / TMEM /
TMA use (B200_PROFILING.md).   python tools/sass_evidence.py r02"""
import collections, os, re, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PAT = re.compile(r"\b(UTCHMMA|UTCQMMA|UTCCP|LDTM|STTM|UBLKCP|UBLKPF|UTMALDG|UTMASTG|SYNCS|UTCBAR|UTCATOMSWS|HMMA|FFMA|SHFL|ATOMS|ATOMG|RED)\b")


def main(tag):
    sass = subprocess.run(["cuobjdump", "-sass", os.path.join(ROOT, "howl_b200", "lib", "libhowl_b200.so")], capture_output=True, text=True).stdout
    cur, counts = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1)
        elif cur:
            for op in PAT.findall(line):
                counts[cur][op] += 1
    rows = []
    for fn, c in counts.items():
        if any(k in c for k in ("UTCHMMA", "UTCCP", "LDTM", "UBLKCP", "UTMALDG", "UBLKPF")):
            name = re.sub(r"\(.*", "", subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip())
            rows.append((name, c))
    with open(os.path.join(ROOT, "profiles", f"{tag}_sass_tcgen05_tma.txt"), "w") as f:
        f.write("# cuobjdump -sass howl_b200/lib/libhowl_b200.so (sm_100a); per kernel: count of the SASS mnemonics that prove tcgen05 / TMEM / TMA use\n"
                "# UTCHMMA = tcgen05.mma (bf16), UTCCP = tcgen05.cp (smem -> TMEM), LDTM = tcgen05.ld, UBLKCP = cp.async.bulk (TMA 1-D), UBLKPF = bulk L2\n"
                "# prefetch, SYNCS = mbarrier ops, UTCBAR = tcgen05.commit, UTCATOMSWS = tcgen05.alloc / dealloc; FFMA / SHFL / ATOMS / RED for context.\n"
                "# Written by tools/sass_evidence.py\n")
        for name, c in sorted(rows):
            f.write(f"{name}\n    " + "  ".join(f"{k}={v}" for k, v in sorted(c.items())) + "\n")


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else "r02")
