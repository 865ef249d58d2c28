"""Executed warp instructions of k_front by phase, with the average number of active lanes, from an .ncu-rep captured
with -lineinfo and --import-source on:

    python tools/ncu_phase_split.py report.ncu-rep

Source lines are attributed to a phase by file and line range of fauxgl_b200/csrc (see BUCKETS); a SASS address listed
under several lines (inlining) is counted once, under the first."""
import collections
import csv
import io
import subprocess
import sys


import os
import re


def _marks():
    """Line numbers of the phase boundaries in the current fgl_geom.cu (the FRONT_TICK markers of k_front)."""
    src = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "fauxgl_b200", "csrc", "fgl_geom.cu")
    m = {}
    for i, l in enumerate(open(src), 1):
        if re.match(r"k_front\(const __grid_constant__", l):
            m["kernel"] = i
        for k in ("FRONT_TICK(0)", "FRONT_TICK(2)"):
            if k in l and "define" not in l:
                m.setdefault(k, i)
    return m["kernel"], m["FRONT_TICK(0)"], m["FRONT_TICK(2)"]


L_KERNEL, L_PHASE1_END, L_WALK = _marks()


def bucket(fname, line):
    if fname == "fgl_walk.cuh":
        return "phase 2: row walker (replay, skip-ahead, run loops)"
    if fname == "fgl_math.cuh":
        return "phase 1: vertex transform, go_int, min/max (fgl_math.cuh)"
    if fname == "fgl_block.cuh":
        return "block / warp scans"
    if fname == "fgl_geom.cu":
        if line < L_KERNEL:
            return "phase 1: bbox, per-triangle setup, cull filter (helpers above k_front)"
        if line <= L_PHASE1_END:
            return "phase 1: k_front body (loads, NDC divisions, area, screen transform)"
        if line <= L_WALK:
            return "scans, item ranges, first record lookup, region hand-over"
        return "phase 2: walk loop body (item -> record, SegV stores), publish"
    return "other (" + fname + ")"


def main():
    rep = sys.argv[1]
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:k_front"], capture_output=True, text=True).stdout
    fname, cur, seen = "?", None, set()
    inst, thr = collections.Counter(), collections.Counter()
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
            continue
        if len(r) < 9 or not r[2].startswith("0x") or r[2] in seen or cur is None:
            continue
        seen.add(r[2])
        try:
            n, t = int(r[7] or 0), int(r[8] or 0)
        except ValueError:
            continue
        b = bucket(*cur)
        inst[b] += n
        thr[b] += t
    tot = sum(inst.values())
    print("k_front: %d executed warp instructions" % tot)
    for b, v in inst.most_common():
        print("  %5.1f %%  %10d  lanes %4.1f  %s" % (100.0 * v / tot, v, thr[b] / max(v, 1), b))


if __name__ == "__main__":
    main()
