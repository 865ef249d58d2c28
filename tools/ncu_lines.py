"""Per-source-line warp-stall samples of one kernel in an .ncu-rep (needs -lineinfo and --import-source on):

    python tools/ncu_lines.py report.ncu-rep k_strip [top]

Reads `ncu --page source --print-source cuda,sass --csv`: one block per source file, CUDA rows (with a
line number) carry the samples aggregated over their SASS rows."""
import csv
import io
import subprocess
import sys


def main():
    rep, kern = sys.argv[1], sys.argv[2]
    top = int(sys.argv[3]) if len(sys.argv) > 3 else 40
    raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--print-source", "cuda,sass", "--csv",
                          "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
    fname, hdr, out, total = "?", None, [], 0
    for r in csv.reader(io.StringIO(raw)):
        if not r:
            continue
        if r[0] == "File Path":
            fname = r[1].split("/")[-1]
            continue
        if r[0] == "Function Name":
            continue
        if r[0] == "Line No":
            hdr = r
            continue
        if hdr is None or not r[0].strip().isdigit():
            continue
        col = {h: i for i, h in reversed(list(enumerate(hdr)))}   # first occurrence wins
        def num(name):
            try:
                return int(r[col[name]] or 0)
            except (KeyError, ValueError, IndexError):
                return 0
        n = num("# Samples")
        total += n
        stalls = sorted(((h[6:], num(h)) for h in hdr if h.startswith("stall_") and "Not Issued" not in h), key=lambda kv: -kv[1])[:3]
        out.append((n, num("Instructions Executed"), num("Avg. Threads Executed"), fname, r[0], r[1].strip(), stalls))
    print("total samples", total)
    for n, inst, thr, f, line, src, stalls in sorted(out, key=lambda t: -t[0])[:top]:
        if not n:
            break
        print("%5.1f%% %6d inst %8d thr %2d  %s:%s  %-64.64s  %s" % (100.0 * n / max(total, 1), n, inst, thr, f, line, src,
              " ".join("%s=%d" % kv for kv in stalls if kv[1])))


if __name__ == "__main__":
    main()
