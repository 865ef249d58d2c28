"""Run bench.py under several values of a tuning environment variable and print the stage times.

    python tools/sweep_env.py FGL_SMALL_ROWS 0 2 4 8 16
"""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def main():
    var, values = sys.argv[1], sys.argv[2:]
    for v in values:
        env = dict(os.environ)
        env[var] = v
        out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--no-cpu", "--steps", "20"], env=env,
                             capture_output=True, text=True)
        line = [l for l in out.stdout.splitlines() if l.startswith("{")]
        if not line:
            print(var, v, "FAILED", out.stderr[-400:])
            continue
        j = json.loads(line[-1])
        st = j["roofline"]["stages_ms"]
        s8 = j["ssaa16"]["stages_ms"] if j.get("ssaa16") else {}
        print("%s=%-3s 1080p %.4f ms  [geo %.3f span %.3f sort %.3f raster %.3f]   8K+resolve %.4f ms [geo %.3f span %.3f sort %.3f raster %.3f]" % (
            var, v, j["ms_per_step"], st["geometry_ms"], st["spans_ms"], st["sort_ms"], st["raster_ms"],
            j["ssaa16"]["ms_per_frame"] if j.get("ssaa16") else 0, s8.get("geometry_ms", 0), s8.get("spans_ms", 0),
            s8.get("sort_ms", 0), s8.get("raster_ms", 0)), flush=True)


if __name__ == "__main__":
    main()
