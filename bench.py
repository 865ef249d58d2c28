#!/usr/bin/env python
"""bench.py -- the headline benchmark of fogleman/fauxgl's DrawMesh path
(README.md:36: 871 306-triangle Phong render at 1920x1080; and 16x SSAA) on
N B200s.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl native|reference]

A "step" is one frame: ClearDepthBuffer + ClearColorBufferWith + DrawMesh of the
synthetic 871 306-triangle mesh (SURVEY.md 8d "M871k"; the dragon itself is not
in the reference repository).  `value` is device-timed with the mesh resident in
HBM; `e2e` goes through the public Context API with host (pinned) mesh arrays
uploaded and the image read back inside the timed region.  With N > 1 every
rank renders its own frames of the batch (frame sharding, no collective: weak
scaling); the sort-last composite is timed separately in `sort_last`.

--impl reference times the reference's own CPU algorithm (goroutine striding +
256 mutexes restated with pthreads, oracle/fauxgl_oracle.c -- the Go toolchain is
absent) on the host cores, same scene, same metric.
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for _p in (ROOT, os.path.join(ROOT, "tests")):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import numpy as np  # noqa: E402

os.environ.setdefault("CUDA_DEVICE_MAX_CONNECTIONS", "32")   # independent streams (clears, copies, peer flags) get queues of their own

T_TRIANGLES = 871306
WORKLOAD = ("M871k: synthetic 871306-triangle bumpy closed surface, smoothed normals, Phong, 1920x1080, README camera; "
            "one frame = ClearDepth + ClearColor + DrawMesh")
W1, H1 = 1920, 1080
SSAA = 4
ALGO_BYTES_C1 = T_TRIANGLES * 144 + W1 * H1 * 12                     # 150 351 264
ALGO_BYTES_C2 = (T_TRIANGLES * 144 + (W1 * SSAA) * (H1 * SSAA) * 12
                 + (W1 * SSAA) * (H1 * SSAA) * 4 + W1 * H1 * 4)      # 664 604 064


def log(msg: str):
    """Progress on stderr (stdout carries the one JSON line)."""
    sys.stderr.write("[bench %s] %s\n" % (os.environ.get("RANK", "0"), msg))
    sys.stderr.flush()


def load_traffic(kernel: str):
    """dram__bytes_read.sum + dram__bytes_write.sum of one launch of `kernel`, from the committed summary of the
    latest `ncu --set full` capture (profiles/ncu_traffic.json; tools/ncu_summary.py writes the numbers)."""
    try:
        with open(os.path.join(ROOT, "profiles", "ncu_traffic.json")) as f:
            e = json.load(f)[kernel]
        return int(e["dram_bytes"]), e.get("source")
    except Exception:
        return None, None


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md)"


def scene_setup():
    """README.md:53-114 'Complete Example'."""
    from fauxgl_b200 import HexColor, LookAt, NewPhongShader, V
    eye, center, up = V(-3, 1, -0.75), V(0, -0.07, 0), V(0, 1, 0)
    matrix = LookAt(eye, center, up).Perspective(30, 1920 / 1080, 1, 10)
    shader = NewPhongShader(matrix, V(-0.75, 1, 0.25).Normalize(), eye)
    shader.ObjectColor = HexColor("#468966")
    return shader, HexColor("#FFF8E3")


class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML polled from a thread
    every ~0.5 ms (nvidia-smi's 100 ms loop is coarser than a whole timed region here)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, index: int):
        self.index, self.sm, self.mask, self.h, self.stop_flag, self.max_mhz = index, [], 0, None, False, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            try:
                import torch
                uuid = str(torch.cuda.get_device_properties(index).uuid)
                self.h = pynvml.nvmlDeviceGetHandleByUUID(("GPU-" + uuid) if not uuid.startswith("GPU-") else uuid)
            except Exception:
                self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.h = None

    def _poll(self):
        nv = self.nv
        while not self.stop_flag:
            try:
                self.sm.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)))
                self.mask |= int(nv.nvmlDeviceGetCurrentClocksEventReasons(self.h))
            except Exception:
                pass
            time.sleep(0.0005)

    def start(self):
        if self.h is None:
            return
        self.stop_flag = False
        self.t = threading.Thread(target=self._poll, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "samples": 0, "reasons": ["nvml unavailable"]}
        self.stop_flag = True
        self.t.join(timeout=1)
        reasons = sorted(name for bit, name in self.REASONS.items() if self.mask & bit)
        return {"sm_mhz": float(np.median(self.sm)) if self.sm else None, "sm_max_mhz": self.max_mhz,
                "samples": len(self.sm), "reasons": reasons}


def cpu_reference_frames(mesh, frames: int, warmup: int, threads: int):
    """The reference's CPU schedule (context.go:413-433: worker wi takes i%wn==wi, 256
    mutexes, unlocked early-Z) restated in oracle/fauxgl_oracle.c, timed per frame."""
    from oracle import pyoracle
    shader, bg = scene_setup()
    octx = pyoracle.OracleContext(W1, H1, threads=threads)
    octx.Shader = shader
    verts = np.ascontiguousarray(mesh.triangle_vertices())  # flattening is mesh prep, not DrawMesh
    times = []
    for i in range(warmup + frames):
        t0 = time.perf_counter()
        octx.ClearDepthBuffer()
        octx.ClearColorBufferWith(bg)
        octx._draw(pyoracle.lib().oracle_draw_triangles, verts)
        dt = time.perf_counter() - t0
        if i >= warmup:
            times.append(dt)
    return times


def run_reference(args, rank):
    if rank != 0:
        return
    from fauxgl_b200 import synth
    mesh = synth.bumpy_surface()
    cores = os.cpu_count() or 1
    times = cpu_reference_frames(mesh, args.steps, min(args.warmup, 5), cores)
    ms = 1e3 * sum(times) / len(times)
    value = T_TRIANGLES / (ms / 1e3) / 1e6
    line = {
        "impl": "reference", "metric": "Mtri/s, 871k-tri Phong 1080p DrawMesh", "value": value, "unit": "Mtri/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": min(args.warmup, 5), "ms_per_step": ms,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": value / (T_TRIANGLES / 0.150 / 1e6), "dtype": "f64", "data": "synthetic",
        "config": {"workload": WORKLOAD, "frames_per_step": 1},
        "cpu_baseline": {"value": value, "unit": "Mtri/s", "cores": cores, "kind": "port",
                         "sample": "%d full frames (clear + DrawMesh) of the same scene, pthread restatement of the reference's "
                                   "goroutine schedule (no Go toolchain)" % len(times)},
        "e2e": {"value": value, "unit": "Mtri/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def bench_sort_last(args, rank, local_rank, world, barrier, max_over_ranks, use_dist):
    """BASELINE config 5 (SURVEY 8d "M10M"): a 10 003 864-triangle unit sphere at 7680x4320, the triangles split over
    the N ranks, every rank drawing into its own full-frame buffers, then the depth composite onto rank 0.  Runs at
    every N including 1 (the anchor of the scaling curve).  Per N it reports, for both composites of the library
    (exact sparse peer-memory kernel; NCCL stripe reduce-scatter + gather on packed keys) and both ways of splitting
    the triangles (contiguous ranges; blocks of 4096 dealt round-robin), the frame time with its breakdown and the
    number of pixels in which rank 0's composited frame differs from a single-GPU render of the whole mesh (which
    rank 0 makes once, untimed) -- and asserts that count against the north star's 0.01 % budget."""
    import torch
    import torch.distributed as dist
    from fauxgl_b200 import multigpu, synth
    from fauxgl_b200.context import Context, DeviceMesh
    K = max(3, min(args.steps, 8))
    log("sort-last: building the 10M-triangle sphere")
    nu = nv = 2237                                   # 2237 * (2*2237 - 2) = 10 003 864 triangles
    big = synth.uv_sphere(nu, nv)
    Tb = big.num_triangles
    Wb, Hb = W1 * SSAA, H1 * SSAA
    npix = Wb * Hb
    ctx = Context(Wb, Hb, local_rank)
    sh_b, bg_b = scene_setup()
    ctx.Shader = sh_b
    ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local_rank))

    def gather_floats(x):
        if not use_dist:
            return [float(x)]
        t = torch.zeros(world, dtype=torch.float64, device="cuda")
        t[rank] = x
        dist.all_reduce(t)
        return [float(v) for v in t.tolist()]

    def timed(frame_fn, steps, finish):
        s_ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        m_ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        e_ev = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        barrier()
        with torch.cuda.stream(ext):
            for i in range(steps):
                s_ev[i].record(ext)
                frame_fn(lambda i=i: m_ev[i].record(ext))
                e_ev[i].record(ext)
        info = finish()
        barrier()
        frame = sum(a.elapsed_time(b) for a, b in zip(s_ev, e_ev)) / steps
        draw = sum(a.elapsed_time(b) for a, b in zip(s_ev, m_ev)) / steps
        return frame, draw, info

    # ---- every rank keeps the whole mesh resident and draws a sub-range of it: re-balancing moves no data
    dm_full = DeviceMesh(ctx, big, ("position", "normal"))
    ctx.ClearDepthBuffer(); ctx.ClearColorBufferWith(bg_b)
    full_info = ctx.DrawTriangles(dm_full)           # synchronous: sizes the work buffers, gives the RasterizeInfo
    ref_img = single = None
    if rank == 0:                                    # the single-GPU render of the whole mesh: the image to compare with
        ref_img = ctx.Image().copy()
        ref_depth = ctx.DepthBuffer

    def stage_ms(st):
        return {"geometry_ms": st.geometry_ms / st.draws, "spans_ms": st.spans_ms / st.draws,
                "sort_ms": st.sort_ms / st.draws, "raster_ms": st.raster_ms / st.draws}
    if world == 1:                                   # ... and the N = 1 anchor of the scaling curve
        def full_frame(mark):
            ctx.ClearDepthBuffer()
            ctx.ClearColorBufferWith(bg_b)
            ctx.DrawMeshAsync(dm_full)
            mark()
        for _ in range(2):
            full_frame(lambda: None)
        ctx.Sync()
        fms, _, _ = timed(full_frame, K, ctx.Sync)
        ctx.SetProfiling(True)
        for _ in range(3):
            full_frame(lambda: None)
        ctx.Sync()
        st = ctx.StageTimes()
        ctx.SetProfiling(False)
        single = {"ms_per_frame": fms, "mtri_s": Tb / (fms / 1e3) / 1e6, "stages_ms": stage_ms(st),
                  "total_pixels": int(full_info.TotalPixels), "updated_pixels": int(full_info.UpdatedPixels)}
    barrier()

    peer = multigpu.PeerGroup(ctx, rank, world)
    nccl = multigpu.NcclComposite(ctx, rank, world)
    runs = {}
    for partition in ("contiguous", "balanced", "interleaved"):
        log("sort-last: " + partition)
        dm, first, count = dm_full, 0, Tb
        bounds = [r * Tb // world for r in range(world + 1)]
        if partition == "interleaved":
            idx = multigpu.triangle_blocks(Tb, rank, world, 4096)
            part = type(big)(big.position[idx], big.normal[idx])
            dm = DeviceMesh(ctx, part, ("position", "normal"))
            first, count = 0, part.num_triangles
            del part
        elif partition == "balanced" and world > 1:
            # feedback: draw, measure every rank's draw time, cut the cumulative cost into equal parts, repeat
            for _round in range(6):
                first, count = bounds[rank], bounds[rank + 1] - bounds[rank]
                ctx.ClearDepthBuffer(); ctx.ClearColorBufferWith(bg_b)
                ctx.DrawTriangles(dm, first, count)              # (sizes the work buffers for the new range)
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                with torch.cuda.stream(ext):
                    e0.record(ext)
                    for _ in range(3):
                        ctx.ClearDepthBuffer(); ctx.ClearColorBufferWith(bg_b); ctx.DrawMeshAsync(dm, first, count)
                    e1.record(ext)
                ctx.Sync()
                bounds = multigpu.rebalance_ranges(bounds, gather_floats(e0.elapsed_time(e1) / 3))
        first, count = (bounds[rank], bounds[rank + 1] - bounds[rank]) if partition != "interleaved" else (first, count)
        ctx.ClearDepthBuffer(); ctx.ClearColorBufferWith(bg_b)
        my_info = ctx.DrawTriangles(dm, first, count)    # synchronous: sizes the work buffers for this share
        for method, comp, finish in (("peer", lambda: peer.composite(0), lambda: (ctx.Sync(), peer.status())[0]),
                                     ("peer_color", lambda: peer.composite(0, color_only=True), lambda: (ctx.Sync(), peer.status())[0]),
                                     ("nccl", lambda: nccl.composite(0), ctx.Sync)):
            def frame(mark, comp=comp):
                ctx.ClearDepthBuffer()
                ctx.ClearColorBufferWith(bg_b)
                ctx.DrawMeshAsync(dm, first, count)
                mark()
                comp()
            for _ in range(2):
                frame(lambda: None)
            finish()
            fms, dms, info = timed(frame, K, finish)
            fms = max_over_ranks(fms)
            draws = gather_floats(dms)
            # correctness of what rank 0 now holds, against its own single-GPU render of the whole mesh
            mism = dmism = None
            if rank == 0:
                img = ctx.Image()
                mism = int((img != ref_img).any(axis=-1).sum())
                if method == "peer":
                    dmism = int((ctx.DepthBuffer.view(np.uint64) != ref_depth.view(np.uint64)).sum())
            # stage breakdown: a second, shorter pass with the library's stage timers on
            ctx.SetProfiling(True)
            for _ in range(3):
                frame(lambda: None)
            finish()
            stages = (nccl if method == "nccl" else peer).stage_times()
            dstages = stage_ms(ctx.StageTimes())
            ctx.SetProfiling(False)
            stages.pop("composites", None)
            runs["%s/%s" % (partition, method)] = {
                "ms_per_frame": fms, "mtri_s": Tb / (fms / 1e3) / 1e6,
                "draw_ms_max": max(draws), "draw_ms_min": min(draws), "draw_ms_per_rank": draws,
                "draw_stages_ms_rank0": dstages, "composite_stages_ms_rank0": stages,
                "mismatch_px": mism, "depth_mismatch_px": dmism,
                "mismatch_frac": None if mism is None else mism / npix}
            if rank == 0:
                budget = 1e-4 * npix
                assert mism <= (budget if method == "nccl" else 0), (partition, method, mism)
                assert dmism in (None, 0), (partition, method, dmism)
        totals = gather_floats(float(my_info.TotalPixels))
        runs[partition + "/total_pixels_per_rank"] = [int(v) for v in totals]
        if partition != "interleaved":
            runs[partition + "/range_bounds"] = bounds
        else:
            del dm
        barrier()
    peer.close()
    nccl.close()
    del dm_full
    ctx.Close()
    best = min((k for k in runs if "ms_per_frame" in (runs[k] if isinstance(runs[k], dict) else {})),
               key=lambda k: runs[k]["ms_per_frame"])
    return {"workload": "M10M: %d-triangle unit sphere, Phong, 7680x4320, triangles split over the N ranks, composite onto "
                        "rank 0 (order-independent state)" % Tb,
            "triangles": Tb, "n_gpus": world, "steps": K, "scaling": "strong",
            "ms_per_frame": runs[best]["ms_per_frame"], "mtri_s": runs[best]["mtri_s"], "best": best,
            "mismatch_px": runs[best]["mismatch_px"], "single_gpu": single, "runs": runs,
            "methods": {"peer": "fgl_peer_composite: exact float64 depth, sparse (dirty strips only), one P2P kernel per "
                                "rank over NVLink, device-side flags; rank 0 receives depth + colour",
                        "peer_color": "the same with FGL_COMPOSITE_COLOR_ONLY: rank 0 receives the colours only (it presents "
                                      "the frame; 4 instead of 12 bytes per pixel through its NVLink ingress)",
                        "nccl": "fgl_composite: packed keys (depth32<<32|rgba8), ncclReduceScatter(min, uint64) by stripe + "
                                "ncclSend/Recv gather to rank 0, inside the library"},
            "partitions": {"contiguous": "rank r draws triangles [rT/N, (r+1)T/N)",
                           "balanced": "contiguous ranges of equal measured draw time (six rounds of feedback: "
                                       "multigpu.rebalance_ranges; every rank keeps the mesh resident, so moving a boundary moves no data)",
                           "interleaved": "blocks of 4096 consecutive triangles dealt round-robin"}}


_REAL_STDOUT = None


def claim_stdout():
    """stdout carries exactly ONE JSON line.  Libraries print there too (NCCL writes its version banner to stdout when
    NCCL_DEBUG is set in the environment), so file descriptor 1 is pointed at stderr for the run and the line goes to a
    private duplicate of the original."""
    global _REAL_STDOUT
    if _REAL_STDOUT is None:
        sys.stdout.flush()
        _REAL_STDOUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)


def emit(line: dict):
    _REAL_STDOUT.write(json.dumps(line) + "\n")
    _REAL_STDOUT.flush()


def main():
    claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="native", choices=["native", "reference"])
    ap.add_argument("--no-ssaa", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    ap.add_argument("--no-batch", action="store_true")
    ap.add_argument("--no-sort-last", action="store_true")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))

    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from fauxgl_b200 import synth
    from fauxgl_b200.context import Context, DeviceMesh

    assert torch.cuda.is_available(), "bench.py needs a CUDA device (no CPU fallback)"
    torch.cuda.set_device(local_rank)
    use_dist = world > 1
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))

    def barrier():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(x: float) -> float:
        if not use_dist:
            return x
        t = torch.tensor([x], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    K, Wm = args.steps, max(args.warmup, 3)
    mesh = synth.bumpy_surface()
    shader, bg = scene_setup()
    hbm_peak, peak_src = load_peaks()

    # ---------------- device-resident throughput (value) ----------------
    def timed_frames(width, height, resolve, steps, warmup, profile, probe=False, graph=False):
        ctx = Context(width, height, local_rank)
        ctx.Shader = shader
        dm = DeviceMesh(ctx, mesh, ("position", "normal"))
        ext = torch.cuda.ExternalStream(ctx.stream_ptr, device=torch.device("cuda", local_rank))
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")

        def frame():
            ctx.ClearDepthBuffer()
            ctx.ClearColorBufferWith(bg)
            ctx.DrawMeshAsync(dm)
            if resolve:
                ctx.ResolveDevice(resolve)
        ctx.DrawMesh(dm)          # one synchronous draw sizes the work buffers (async draws cannot regrow)
        if resolve:
            ctx.ResolveDevice(resolve)   # (allocates the resolve target: nothing may be allocated while recording)
        direct = frame
        recorded = None
        if graph and not profile:
            # the frame recorded once into a CUDA graph (fgl_graph_begin/end) and replayed with one call per step:
            # the same kernels in the same order, without the host's launch gaps between them
            ctx.GraphBegin()
            direct()
            recorded = ctx.GraphEnd()
            frame = recorded.launch
        for _ in range(warmup):
            frame()
        info = ctx.Sync()
        stats = ctx.DrawStats()
        launches_per_frame = 2 + stats.kernel_launches + (1 if resolve else 0)   # clears + the async draw (+ resolve)
        if profile:
            ctx.SetProfiling(True)
        starts = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        ends = [torch.cuda.Event(enable_timing=True) for _ in range(steps)]
        sampler = ClockSampler(local_rank)
        barrier()
        sampler.start()
        with torch.cuda.stream(ext):
            for i in range(steps):
                flush.zero_()                       # L2 flush between timed frames (256 MiB > 126 MB L2)
                starts[i].record(ext)
                frame()
                ends[i].record(ext)
        info = ctx.Sync()
        barrier()
        clocks = sampler.stop()
        ms = [s.elapsed_time(e) for s, e in zip(starts, ends)]
        st = ctx.StageTimes() if profile else None
        if profile:
            ctx.SetProfiling(False)
        if recorded is not None:
            recorded.Close()
        out = {"ms": ms, "info": info, "records": int(stats.records), "pairs": int(stats.pairs),
               "launches_per_frame": int(launches_per_frame), "clocks": clocks, "stage": st}
        if probe:   # fragment-rate bound: 64-bit atomicMin on a depth-buffer-sized array of this box
            out["atomic_ops_per_s"] = ctx.ProbeAtomicRate(1 << 26)
        del dm
        ctx.Close()
        return out

    log("device-resident frames (value)")
    # `value`: no stage events between the kernels (they would serialise the programmatic dependent launches);
    # the per-stage times come from a second, shorter pass with the library's stage timers on.
    r1 = timed_frames(W1, H1, 0, K, Wm, False, probe=(rank == 0), graph=True)
    ms1 = max_over_ranks(sum(r1["ms"]) / K)
    r1d = timed_frames(W1, H1, 0, max(3, min(K, 10)), 3, False)
    ms1_direct = max_over_ranks(sum(r1d["ms"]) / len(r1d["ms"]))
    value = world * T_TRIANGLES / (ms1 / 1e3) / 1e6
    r1["stage"] = timed_frames(W1, H1, 0, max(3, min(K, 10)), 3, True)["stage"]

    def stage_ms(st):
        return {"geometry_ms": st.geometry_ms / st.draws, "spans_ms": st.spans_ms / st.draws,
                "sort_ms": st.sort_ms / st.draws, "raster_ms": st.raster_ms / st.draws}
    stages = stage_ms(r1["stage"])
    # Dominant kernel: k_front, the fused geometry + span kernel (50 % of the frame in the ncu launch list,
    # profiles/r01_launch_table_v7.txt); the `geometry` stage timer brackets exactly that launch (and the
    # 64-byte counter memset before it).  Its algorithmic bytes: the position planes,
    # T * 72 B, read once -- the segments it writes are implementation, not algorithm.
    geo_bytes = T_TRIANGLES * 72
    frame_fragments = int(r1["info"].TotalPixels // K)   # RasterizeInfo.TotalPixels of one frame (Sync() returns the timed frames' sum)
    fragment_bound = None
    if r1.get("atomic_ops_per_s"):
        f_us = frame_fragments / r1["atomic_ops_per_s"] * 1e6
        h_us = ALGO_BYTES_C1 / (hbm_peak * 1e9) * 1e6
        fragment_bound = {"fragments_per_frame": frame_fragments, "atomic_min_u64_per_s": r1["atomic_ops_per_s"],
                          "bound_us": f_us, "hbm_bound_us": h_us, "slower_bound": "hbm" if h_us >= f_us else "fragment",
                          "frac_of_slower_bound": max(f_us, h_us) / (ms1 * 1e3),
                          "primitive": "global 64-bit atomicMin on random words of a 1920x1080 u64 array, measured on this "
                                       "box (fgl_probe_atomic_rate); the path itself resolves fragments in shared memory"}
    roofline = {
        "bound": "hbm", "kernel": "k_front",
        "achieved": geo_bytes / (stages["geometry_ms"] / 1e3) / 1e9, "peak": hbm_peak, "unit": "GB/s",
        "frac": geo_bytes / (stages["geometry_ms"] / 1e3) / 1e9 / hbm_peak,
        "traffic": load_traffic("k_front")[0], "traffic_source": load_traffic("k_front")[1],
        "peak_source": peak_src, "algorithmic_bytes": geo_bytes,
        "frame": {"algorithmic_bytes": ALGO_BYTES_C1, "achieved": ALGO_BYTES_C1 / (ms1 / 1e3) / 1e9,
                  "frac": ALGO_BYTES_C1 / (ms1 / 1e3) / 1e9 / hbm_peak},
        "stages_ms": stages, "fragment_bound": fragment_bound,
        "note": "issue/latency-bound float64 replay of the reference's arithmetic, not bandwidth-bound: "
                "45.5 M warp instructions at 0.49 IPC per scheduler, see profiles/README.md",
    }

    # ---------------- 16x SSAA (7680x4320 + resolve) ----------------
    ssaa = None
    log("8K + resolve")
    if not args.no_ssaa:
        K2 = max(3, min(K, 10))
        r2 = timed_frames(W1 * SSAA, H1 * SSAA, SSAA, K2, 3, False, graph=True)
        ms2 = max_over_ranks(sum(r2["ms"]) / K2)
        s2 = timed_frames(W1 * SSAA, H1 * SSAA, SSAA, 3, 3, True)["stage"]
        ssaa = {"ms_per_frame": ms2, "mtri_s": world * T_TRIANGLES / (ms2 / 1e3) / 1e6, "steps": K2,
                "total_pixels": int(r2["info"].TotalPixels // K2),
                "roofline_frac_frame": ALGO_BYTES_C2 / (ms2 / 1e3) / 1e9 / hbm_peak,
                "stages_ms": stage_ms(s2),
                "workload": "same mesh at 7680x4320 + 4x4 nfnt-bilinear resolve to 1920x1080"}

    # ---------------- animation batch: independent frames in flight on several contexts ----------------
    # 64 frames of an examples/animate.go-style turntable (the camera matrix turns 5 degrees per frame), the
    # frames of this rank dealt round-robin to NCTX contexts (one stream each) that share one device-resident
    # mesh.  Every kernel of a frame is latency-bound and leaves most of the GPU idle (profiles/README.md), so
    # independent frames overlap well.  Device-timed: one start event that every stream waits on, one end event
    # per stream, the longest interval counts.  No L2 flush is possible between concurrent frames; the working
    # set (mesh 125 MB + ~300 MB of segments per context) is several times the 126 MB L2.
    batch = None
    log("animation batch")
    if not args.no_batch:
        from fauxgl_b200 import NewPhongShader, Rotate, V
        NCTX, NFRAMES = 4, 64
        my_frames = [k for k in range(NFRAMES) if k % world == rank]
        ctxs = [Context(W1, H1, local_rank) for _ in range(NCTX)]
        dmb = DeviceMesh(ctxs[0], mesh, ("position", "normal"))
        streams = [torch.cuda.ExternalStream(c.stream_ptr, device=torch.device("cuda", local_rank)) for c in ctxs]
        shaders = []
        for k in range(NFRAMES):
            sh = NewPhongShader(shader.Matrix.Mul(Rotate(V(0, 1, 0), 5.0 * k * 3.141592653589793 / 180.0)), shader.LightDirection,
                                shader.CameraPosition)
            sh.ObjectColor = shader.ObjectColor
            shaders.append(sh)

        def bframe(c, k):
            c.Shader = shaders[k]
            c.ClearDepthBuffer()
            c.ClearColorBufferWith(bg)
            c.DrawMeshAsync(dmb)
        for c in ctxs:
            c.Shader = shaders[0]
            c.DrawMesh(dmb)                      # sizes the work buffers
        for i, k in enumerate(my_frames[:2 * NCTX]):
            bframe(ctxs[i % NCTX], k)
        for c in ctxs:
            c.Sync()
        barrier()
        ev0 = torch.cuda.Event(enable_timing=True)
        evs = [torch.cuda.Event(enable_timing=True) for _ in ctxs]
        ev0.record(streams[0])
        for st_ in streams[1:]:
            st_.wait_event(ev0)
        for i, k in enumerate(my_frames):
            bframe(ctxs[i % NCTX], k)
        for e_, st_ in zip(evs, streams):
            e_.record(st_)
        binfo = [c.Sync() for c in ctxs]
        barrier()
        bms = max_over_ranks(max(ev0.elapsed_time(e_) for e_ in evs))
        batch = {"frames": NFRAMES, "contexts_per_gpu": NCTX, "ms_total": bms, "ms_per_frame": bms / NFRAMES,
                 "mtri_s": T_TRIANGLES * NFRAMES / (bms / 1e3) / 1e6,
                 "total_pixels_this_rank": int(sum(b.TotalPixels for b in binfo)),
                 "workload": "64-frame turntable of the M871k mesh at 1920x1080 (clear + DrawMesh per frame, 5 degrees "
                             "per frame), frames k = rank (mod N) on each GPU, %d frames in flight per GPU" % NCTX}
        del dmb
        for c in ctxs:
            c.Close()

    # ---------------- end to end through the public API with host buffers ----------------
    log("end to end")
    Ke = max(3, min(K, 10))
    ctx = Context(W1, H1, local_rank)
    ctx.Shader = shader
    ctx.upload_attributes = ("position", "normal")
    pin_pos = torch.empty(mesh.position.shape, dtype=torch.float64, pin_memory=True)
    pin_nrm = torch.empty(mesh.normal.shape, dtype=torch.float64, pin_memory=True)
    pin_img = torch.empty((H1, W1, 4), dtype=torch.uint8, pin_memory=True)
    pin_pos.numpy()[...] = mesh.position
    pin_nrm.numpy()[...] = mesh.normal
    hmesh = type(mesh)(pin_pos.numpy(), pin_nrm.numpy())
    img = pin_img.numpy()

    def e2e_frame():
        hmesh.Invalidate()                 # host mesh changed (animate.go:66): positions+normals are re-uploaded
        ctx.ClearDepthBuffer()
        ctx.ClearColorBufferWith(bg)
        info = ctx.DrawMesh(hmesh)         # synchronous, returns RasterizeInfo like the reference
        ctx.Image(out=img)                 # read the frame back
        return info
    for _ in range(3):
        e2e_frame()
    barrier()
    t0 = time.perf_counter()
    for _ in range(Ke):
        einfo = e2e_frame()
    torch.cuda.synchronize()
    serial_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / Ke)
    barrier()
    serial_checksum = int(img.astype(np.uint64).sum())

    # The same frames through the streaming entry points, DEPTH in flight: the upload of frame i+1 (copy
    # stream) overlaps the draw and the read-back of frame i.  Every frame still uploads its mesh from
    # pinned host memory and reads its image and RasterizeInfo back; this is the throughput a loop like
    # examples/animate.go gets from the library, and the headline e2e figure.
    from fauxgl_b200.pipeline import FramePipeline
    DEPTH = 3   # frames in flight (3 vs 2: 0.461 vs 0.476 ms per frame on the B200 box, tools/e2e_probe.py)
    pipe = FramePipeline(ctx, hmesh, depth=DEPTH)

    def pipelined(n):
        last = None
        for _ in range(n):
            if len(pipe) == DEPTH:
                last = pipe.collect()
            pipe.submit(hmesh, bg)
        while len(pipe):
            last = pipe.collect()
        return last
    pipelined(3)
    barrier()
    t0 = time.perf_counter()
    pimg, pinfo = pipelined(Ke)
    torch.cuda.synchronize()
    e2e_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / Ke)
    # (the pipelined flavours below run Kp frames: filling and draining a pipeline DEPTH deep is part of the timed
    # region, and over ten frames that alone is a tenth of the figure)
    Kp = max(Ke, 40)
    barrier()
    assert tuple(pinfo) == tuple(einfo) and int(pimg.astype(np.uint64).sum()) == serial_checksum
    soup = {"ms_per_step": e2e_ms, "mtri_s": world * T_TRIANGLES / (e2e_ms / 1e3) / 1e6,
            "h2d_bytes_per_step": int(pin_pos.numel() * 8 + pin_nrm.numel() * 8), "d2h_bytes_per_step": int(W1 * H1 * 4 + 64),
            "serial_ms_per_step": serial_ms,
            "what": "the mesh as the reference holds it in memory, expanded per triangle corner (position + normal, 144 B "
                    "per triangle), uploaded every frame (fgl_mesh_update_async); serial_ms_per_step = the same frame as "
                    "blocking calls, nothing overlapped"}
    del pipe

    # The headline e2e: the same frames with the mesh kept as shared-vertex tables (Mesh.indexed(): what an OBJ file
    # holds, obj.go:19-79).  The corner indices stay on the device; every frame uploads its v / vn tables from pinned
    # host memory (fgl_mesh_update_indexed_async: 48 B per shared vertex), expands them on the copy stream, clears,
    # draws and reads the image and the RasterizeInfo back -- bit-identical frames, a sixth of the PCIe bytes.
    from fauxgl_b200.pipeline import IndexedFramePipeline
    tv, tvn, corners = mesh.indexed()
    pin_v = torch.empty(tv.shape, dtype=torch.float64, pin_memory=True)
    pin_vn = torch.empty(tvn.shape, dtype=torch.float64, pin_memory=True)
    pin_v.numpy()[...] = tv
    pin_vn.numpy()[...] = tvn
    ipipe = IndexedFramePipeline(ctx, pin_v.numpy(), pin_vn.numpy(), corners, depth=DEPTH)

    def ipipelined(n):
        last = None
        for _ in range(n):
            if len(ipipe) == DEPTH:
                last = ipipe.collect()
            ipipe.submit(pin_v.numpy(), pin_vn.numpy(), bg)
        while len(ipipe):
            last = ipipe.collect()
        return last
    ipipelined(3)
    barrier()
    t0 = time.perf_counter()
    iimg, iinfo = ipipelined(Kp)
    torch.cuda.synchronize()
    idx_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / Kp)
    barrier()
    assert tuple(iinfo) == tuple(einfo) and int(iimg.astype(np.uint64).sum()) == serial_checksum
    del ipipe

    # examples/animate.go as the library lets it be written: the mesh lives on the device, a frame's input is the 4x4
    # matrix of Mesh.Transform (mesh.go:167-175 on the device: fgl_mesh_transform), and the image is read back.
    from fauxgl_b200 import Rotate, V
    from fauxgl_b200.context import Fence, pinned_empty
    dmt = DeviceMesh(ctx, mesh, ("position", "normal"))
    rot = Rotate(V(0, 1, 0), 5.0 * 3.141592653589793 / 180.0)
    slots = [(pinned_empty((H1, W1, 4), np.uint8), Fence(ctx)) for _ in range(DEPTH)]

    def transform_frames(n):
        info = None
        for i in range(n):
            out, fence = slots[i % DEPTH]
            if i >= DEPTH:
                info = fence.wait()
            dmt.Transform(rot)
            ctx.ClearDepthBuffer()
            ctx.ClearColorBufferWith(bg)
            ctx.DrawMeshAsync(dmt)
            ctx.FrameEnd(out, fence)
        for i in range(n, n + DEPTH):
            info = slots[i % DEPTH][1].wait()
        return info
    ctx.DrawMesh(dmt)
    transform_frames(3)
    barrier()
    t0 = time.perf_counter()
    tinfo = transform_frames(Kp)
    torch.cuda.synchronize()
    tr_ms = max_over_ranks((time.perf_counter() - t0) * 1e3 / Kp)
    barrier()
    assert tinfo.TotalPixels > 0
    del dmt
    # What PCIe allows for this frame: the same bytes as bare copies (pinned host memory, both directions at once).
    def pcie_floor():
        dev = torch.device("cuda", local_rank)
        dv = [torch.empty_like(pin_v, device=dev), torch.empty_like(pin_vn, device=dev)]
        dimg = torch.empty(H1 * W1 * 4, dtype=torch.uint8, device=dev)
        himg = torch.empty(H1 * W1 * 4, dtype=torch.uint8, pin_memory=True)
        s_up, s_dn = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

        def both():
            with torch.cuda.stream(s_up):
                dv[0].copy_(pin_v, non_blocking=True)
                dv[1].copy_(pin_vn, non_blocking=True)
            with torch.cuda.stream(s_dn):
                himg.copy_(dimg, non_blocking=True)
        both()
        torch.cuda.synchronize()
        t = time.perf_counter()
        for _ in range(10):
            both()
        torch.cuda.synchronize()
        return (time.perf_counter() - t) * 1e3 / 10
    pcie_ms = max_over_ranks(pcie_floor())
    e2e = {"value": world * T_TRIANGLES / (idx_ms / 1e3) / 1e6, "unit": "Mtri/s", "ms_per_step": idx_ms, "steps": Kp,
           "pcie_bound": {"ms_per_step": pcie_ms, "frac": pcie_ms / idx_ms,
                          "what": "the frame's 20.9 MB up and 8.3 MB down as bare cudaMemcpyAsync from / to pinned memory, both "
                                  "directions at once, measured here: the PCIe roofline of this e2e definition"},
           "h2d_bytes_per_step": int(pin_v.numel() * 8 + pin_vn.numel() * 8),
           "d2h_bytes_per_step": int(W1 * H1 * 4 + 64),
           "frames_in_flight": DEPTH,
           "what": "per frame: this frame's vertex tables (435986 shared vertices x (position + normal), pinned host memory) "
                   "uploaded and expanded on the device by resident corner indices (fgl_mesh_update_indexed_async) + clears "
                   "+ DrawMesh + image and RasterizeInfo read-back through the public API; frames pipelined three deep; image "
                   "and RasterizeInfo asserted identical to the blocking per-triangle-soup path",
           "soup_upload": soup,
           "device_transform": {"ms_per_step": tr_ms, "mtri_s": world * T_TRIANGLES / (tr_ms / 1e3) / 1e6,
                                "h2d_bytes_per_step": 128, "d2h_bytes_per_step": int(W1 * H1 * 4 + 64),
                                "what": "examples/animate.go with the mesh resident: per frame Mesh.Transform(Rotate(up, 5 deg)) "
                                        "on the device (fgl_mesh_transform, input = one 4x4 matrix) + clears + DrawMesh + image "
                                        "and RasterizeInfo read-back, two frames in flight"}}
    checksum = int(img.astype(np.uint64).sum())
    ctx.Close()

    # ---------------- sort-last: 10 M-triangle sphere at 7680x4320 (BASELINE config 5) ----------------
    sort_last = None
    log("sort-last")
    if not args.no_sort_last:
        sort_last = bench_sort_last(args, rank, local_rank, world, barrier, max_over_ranks, use_dist)

    # ---------------- CPU baseline beside it (rank 0, N == 1 only) ----------------
    cpu = None
    log("cpu baseline")
    if rank == 0 and world == 1 and not args.no_cpu:
        cores = os.cpu_count() or 1
        times = cpu_reference_frames(mesh, 5, 1, cores)
        cms = 1e3 * float(np.median(times))
        cpu = {"value": T_TRIANGLES / (cms / 1e3) / 1e6, "unit": "Mtri/s", "cores": cores, "kind": "port",
               "ms_per_frame_median": cms, "ms_per_frame_best": 1e3 * min(times),
               "sample": "5 full frames (clear + DrawMesh) of the same 871306-triangle scene; pthread restatement of "
                         "the reference's goroutine schedule (context.go:413-433), the Go toolchain being absent"}

    if rank == 0:
        line = {
            "metric": "Mtri/s, 871k-tri Phong 1080p DrawMesh", "value": value, "unit": "Mtri/s", "n_gpus": world,
            "steps": K, "warmup": Wm, "ms_per_step": ms1, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": value / (T_TRIANGLES / 0.150 / 1e6),   # README.md:36: ~150 ms/frame on the author's (unspecified) CPU
            "dtype": "f64", "data": "synthetic",
            "config": {"workload": WORKLOAD, "frames_per_step": 1, "sharding": "frames per GPU, no collective" if world > 1 else "single GPU",
                       "l2": "flushed between timed frames (256 MiB memset outside the event pair)",
                       "timing": "CUDA events on the library's stream, per frame, summed; max over ranks",
                       "launch": "each frame (2 clears + DrawMesh) recorded once with fgl_graph_begin/end and replayed with "
                                 "fgl_graph_launch; ms_per_step_direct_launch = the same frame as individual calls"},
            "ms_per_step_direct_launch": ms1_direct,
            "roofline": roofline, "cpu_baseline": cpu, "e2e": e2e,
            "gpu_launches": int(r1["launches_per_frame"] * K),
            "clocks": r1["clocks"], "ssaa16": ssaa, "animation_batch": batch, "sort_last": sort_last,
            "raster_info": {"total_pixels": int(einfo.TotalPixels), "updated_pixels": int(einfo.UpdatedPixels),
                            "records": r1["records"], "pairs": r1["pairs"], "image_checksum": checksum},
        }
        emit(line)
    log("done")
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
