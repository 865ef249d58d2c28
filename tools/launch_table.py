"""Summarise an ncu `--metrics gpu__time_duration.sum --csv` launch list: per-launch
times of the last frame and per-kernel totals/shares."""
import collections
import csv
import sys


def load(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 5]
    hdr, out = None, []
    for r in rows:
        if r[0] == "ID":
            hdr = r
            continue
        if hdr:
            d = dict(zip(hdr, r))
            if d.get("Metric Name") == "gpu__time_duration.sum":
                name = d["Kernel Name"].split("(")[0].replace("fgl::", "").replace("<unnamed>::", "")
                scale = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(d.get("Metric Unit", "ns"), 1e-3)
                out.append((name, float(d["Metric Value"].replace(",", "")) * scale))
    return out


def main():
    path = sys.argv[1]
    frames = int(sys.argv[2]) if len(sys.argv) > 2 else 2
    ls = load(path)
    # last frame = from the last k_clear_depth on
    idx = max(i for i, (n, _) in enumerate(ls) if n.startswith("k_clear_depth"))
    last = ls[idx:]
    tot = sum(t for _, t in last)
    print("last frame: %d launches, %.1f us (serialised, cold cache)" % (len(last), tot))
    agg = collections.OrderedDict()
    for n, t in last:
        a = agg.setdefault(n, [0, 0.0])
        a[0] += 1
        a[1] += t
    for n, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("  %-28s x%-2d %9.1f us  %5.1f %%" % (n, c, t, 100 * t / tot))


if __name__ == "__main__":
    main()
