"""World-size-2 tests of the multi-GPU host logic on CPU (gloo backend): triangle-range
partition + packed-key min composite (sort-last), and frame sharding.  The per-rank
renders come from the CPU oracle here; on the GPU box the same flow runs through the
CUDA back end (tests/test_multigpu_gpu.py)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from fauxgl_b200 import multigpu


def test_triangle_ranges_partition_exactly():
    for T in (0, 1, 7, 871306, 10_000_000):
        for N in (1, 2, 3, 4, 8):
            spans = [multigpu.triangle_range(T, r, N) for r in range(N)]
            assert spans[0][0] == 0 and sum(c for _, c in spans) == T
            for (f0, c0), (f1, _c1) in zip(spans, spans[1:]):
                assert f0 + c0 == f1
            assert max(c for _, c in spans) - min(c for _, c in spans) <= 1


def test_frame_shards_cover_the_batch():
    for F in (1, 64, 72):
        for N in (1, 2, 4, 8):
            frames = sorted(k for r in range(N) for k in multigpu.frame_shard(F, r, N))
            assert frames == list(range(F))


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _pack_keys(pyoracle, color, depth):
    flat_c = color.reshape(-1, 4)
    flat_d = depth.reshape(-1)
    keys = np.empty(flat_d.size, dtype=np.uint64)
    for i in range(flat_d.size):
        keys[i] = pyoracle.pack_key(float(flat_d[i]), flat_c[i].tolist())
    return keys.view(np.int64)


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        import scenes
        from oracle import pyoracle
        from fauxgl_b200 import HexColor, LookAt, NewPhongShader, V
        mesh = scenes.bumpy_mesh(41, 41)
        W, H = 160, 96
        eye = V(-3, 1, -0.75)
        matrix = LookAt(eye, V(0, -0.07, 0), V(0, 1, 0)).Perspective(30, W / H, 1, 10)
        shader = NewPhongShader(matrix, V(-0.75, 1, 0.25).Normalize(), eye)
        shader.ObjectColor = HexColor("#468966")
        ctx = pyoracle.OracleContext(W, H)
        ctx.Shader = shader
        ctx.ClearColorBufferWith(HexColor("#FFF8E3"))
        first, count = multigpu.triangle_range(mesh.num_triangles, rank, world)
        info = ctx.DrawTriangles(mesh, first, count)
        keys = torch.from_numpy(_pack_keys(pyoracle, ctx.ColorBuffer, ctx.DepthBuffer).copy())
        multigpu.composite_min(keys)
        tot = torch.tensor([info[0]], dtype=torch.int64)
        dist.all_reduce(tot)
        # frame sharding: every rank renders its own frames, nothing is exchanged
        frames = multigpu.frame_shard(6, rank, world)
        np.save(os.path.join(out_dir, "keys_%d.npy" % rank), keys.numpy())
        np.save(os.path.join(out_dir, "meta_%d.npy" % rank), np.array([int(tot.item()), len(frames)]))
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(300)
def test_sort_last_composite_world2_gloo(tmp_path, oracle_lib):
    import scenes
    from fauxgl_b200 import HexColor, LookAt, NewPhongShader, V
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    k0 = np.load(tmp_path / "keys_0.npy")
    k1 = np.load(tmp_path / "keys_1.npy")
    assert (k0 == k1).all()                       # all-reduce: every rank holds the composite
    m0, m1 = np.load(tmp_path / "meta_0.npy"), np.load(tmp_path / "meta_1.npy")
    assert m0[1] + m1[1] == 6
    # single-process reference render of the whole mesh
    mesh = scenes.bumpy_mesh(41, 41)
    W, H = 160, 96
    eye = V(-3, 1, -0.75)
    matrix = LookAt(eye, V(0, -0.07, 0), V(0, 1, 0)).Perspective(30, W / H, 1, 10)
    shader = NewPhongShader(matrix, V(-0.75, 1, 0.25).Normalize(), eye)
    shader.ObjectColor = HexColor("#468966")
    ctx = oracle_lib.OracleContext(W, H)
    ctx.Shader = shader
    ctx.ClearColorBufferWith(HexColor("#FFF8E3"))
    info = ctx.DrawTriangles(mesh)
    assert int(m0[0]) == info[0]                  # TotalPixels adds up over ranks
    want = _pack_keys(oracle_lib, ctx.ColorBuffer, ctx.DepthBuffer)
    mismatch = int((want != k0).sum())
    print("sort-last composite mismatches vs single render:", mismatch, "of", want.size)
    assert mismatch <= max(1, int(1e-4 * want.size))   # depth ties / 32-bit quantisation only


def test_rebalance_ranges_converges_to_equal_cost():
    """multigpu.rebalance_ranges: feedback on measured per-rank times cuts the triangle index into contiguous ranges of
    equal cost (the cost density along the index is not uniform: ranges that face the camera cover more pixels)."""
    import numpy as np
    from fauxgl_b200.multigpu import rebalance_ranges, triangle_blocks
    dens = np.concatenate([np.full(3000, 1.0), np.full(4000, 3.0), np.full(3000, 0.5)])
    for world in (2, 4, 8):
        bounds = [r * len(dens) // world for r in range(world + 1)]
        for _ in range(8):
            times = [float(dens[bounds[i]:bounds[i + 1]].sum()) for i in range(world)]
            bounds = rebalance_ranges(bounds, times)
            assert bounds[0] == 0 and bounds[-1] == len(dens) and all(a <= b for a, b in zip(bounds, bounds[1:]))
        times = [float(dens[bounds[i]:bounds[i + 1]].sum()) for i in range(world)]
        assert max(times) <= 1.03 * (sum(times) / world), (world, bounds, times)
    # degenerate inputs leave the ranges alone / valid
    assert rebalance_ranges([0, 5, 10], [0.0, 0.0]) == [0, 5, 10]
    assert rebalance_ranges([0, 10], [3.0]) == [0, 10]
    # the round-robin block partition covers every triangle exactly once
    T, world = 100003, 3
    parts = [triangle_blocks(T, r, world, 4096) for r in range(world)]
    assert sorted(np.concatenate(parts).tolist()) == list(range(T))
