"""Debug driver: bench.bench_sort_last alone, single process, with progress on stderr."""
import argparse, os, sys, faulthandler
faulthandler.enable()
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench, json, torch
a = argparse.Namespace(steps=5, warmup=3)
def barrier(): torch.cuda.synchronize()
print("start", file=sys.stderr, flush=True)
try:
    r = bench.bench_sort_last(a, 0, 0, 1, barrier, lambda x: x, False)
    print(json.dumps(r, indent=1))
except BaseException as e:
    import traceback; traceback.print_exc()
    print("EXC", type(e), e, file=sys.stderr, flush=True)
print("end", file=sys.stderr, flush=True)
