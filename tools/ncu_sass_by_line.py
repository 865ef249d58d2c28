"""Executed SASS instructions of one kernel in an .ncu-rep, attributed to source lines and filtered by opcode:

    python tools/ncu_sass_by_line.py report.ncu-rep k_front [opcode-regex] [top]

Needs -lineinfo and --import-source on.  A SASS address may be listed under several source lines (inlining);
it is counted once, under the first line that lists it."""
import collections
import csv
import io
import re
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else ".")
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 30
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    fname, cur, seen = "?", None, set()
    hit, alli, srcs, ops = collections.Counter(), collections.Counter(), {}, collections.Counter()
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] in ("Function Name", "Line No"):
            continue
        if r[0].strip().isdigit():
            cur = (fname, int(r[0]))
            srcs[cur] = r[1].strip()
            continue
        if len(r) < 8 or not r[2].startswith("0x") or r[2] in seen:
            continue
        seen.add(r[2])
        try:
            n = int(r[7] or 0)
        except ValueError:
            continue
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[3])
        op = m.group(2) if m else "?"
        alli[cur] += n
        if pat.search(op):
            hit[cur] += n
            ops[op] += n
    tot = sum(alli.values())
    print("matched %d of %d executed warp instructions (%.1f%%)" % (sum(hit.values()), tot, 100.0 * sum(hit.values()) / max(tot, 1)))
    print("  ".join("%s=%d" % kv for kv in ops.most_common(12)))
    for k, v in hit.most_common(top):
        print("%9d / %9d  %s:%d  %s" % (v, alli[k], k[0], k[1], srcs[k][:100]))


if __name__ == "__main__":
    main()
