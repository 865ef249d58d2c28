"""Sort-last composite tests.

On ANY box (one GPU is enough): the peer-memory composite of the library (fgl_peer_*: device-side flags, dirty-strip
bitmaps, sparse P2P kernel) with the ranks being several contexts of this process on device 0 -- the same kernels and
the same protocol as across GPUs, only the pointers are local -- and the NCCL entry points with a one-rank
communicator.

On a box with >= 2 GPUs: one process per GPU; the in-library NCCL composite (stripe reduce-scatter + gather / all-gather),
the torch.distributed all-reduce baseline, and the peer group over CUDA IPC + NVLink."""
import os
import socket

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import scenes
        from fauxgl_b200 import multigpu
        from fauxgl_b200.context import Context
        mesh = scenes.bumpy_mesh(201, 201)
        sc = scenes.dragon_scene(mesh, 1280, 720)
        ctx = Context(sc.width, sc.height, device=rank)
        keys = torch.empty(sc.width * sc.height, dtype=torch.int64, device="cuda")
        infos = []

        class SortLast:
            def __init__(self, c):
                self.__dict__["c"] = c

            def __getattr__(self, k):
                return getattr(self.c, k)

            def __setattr__(self, k, v):
                setattr(self.c, k, v)

            def DrawMesh(self, m):
                info = multigpu.sort_last_draw(self.c, m, keys, rank, world)
                infos.append(info)
                return info
        sc.run(SortLast(ctx))
        ctx.Sync()
        np.save(os.path.join(out_dir, "img_%d.npy" % rank), ctx.Image())
        np.save(os.path.join(out_dir, "tot_%d.npy" % rank), np.array([infos[0][0]]))
        ctx.Close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_sort_last_two_gpus_nccl(tmp_path, gpu_capi):
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import scenes
    from fauxgl_b200.context import Context
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mesh = scenes.bumpy_mesh(201, 201)
    sc = scenes.dragon_scene(mesh, 1280, 720)
    full = Context(sc.width, sc.height)
    finfo = sc.run(full)
    want = full.Image()
    imgs = [np.load(tmp_path / ("img_%d.npy" % r)) for r in range(world)]
    assert (imgs[0] == imgs[1]).all()
    tot = sum(int(np.load(tmp_path / ("tot_%d.npy" % r))[0]) for r in range(world))
    assert tot == finfo[0][0]
    mism = int((imgs[0] != want).any(axis=-1).sum())
    print("2-GPU sort-last mismatches:", mism)
    assert mism <= 1e-4 * want.shape[0] * want.shape[1]
    full.Close()


def _peer_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    try:
        import scenes
        from fauxgl_b200 import multigpu
        from fauxgl_b200.context import Context
        mesh = scenes.bumpy_mesh(201, 201)
        sc = scenes.dragon_scene(mesh, 1280, 720)
        ctx = Context(sc.width, sc.height, device=rank)
        pc = multigpu.PeerGroup(ctx, rank, world)
        first, count = multigpu.triangle_range(mesh.num_triangles, rank, world)

        class RangeCtx:
            def __init__(self, c):
                self.__dict__["c"] = c

            def __getattr__(self, k):
                return getattr(self.c, k)

            def __setattr__(self, k, v):
                setattr(self.c, k, v)

            def DrawMesh(self, m):
                return self.c.DrawTriangles(m, first, count)
        for _frame in range(3):                      # several frames: the device-side flags must make re-use safe
            sc.run(RangeCtx(ctx))
            pc.composite(-1)                         # every rank receives the frame
        pc.status()
        np.save(os.path.join(out_dir, "pimg_%d.npy" % rank), ctx.Image())
        np.save(os.path.join(out_dir, "pdep_%d.npy" % rank), ctx.DepthBuffer)
        pc.close()
        ctx.Close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_peer_memory_composite_two_gpus_exact(tmp_path, gpu_capi):
    """The sparse fused P2P composite over CUDA IPC + NVLink, synchronised by flags in peer memory: both ranks end
    with the single-GPU frame, bit for bit."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import scenes
    from fauxgl_b200.context import Context
    world = 2
    mp.spawn(_peer_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mesh = scenes.bumpy_mesh(201, 201)
    sc = scenes.dragon_scene(mesh, 1280, 720)
    full = Context(sc.width, sc.height)
    sc.run(full)
    want_c, want_d = full.Image(), full.DepthBuffer
    for r in range(world):
        assert (np.load(tmp_path / ("pimg_%d.npy" % r)) == want_c).all()
        assert (np.load(tmp_path / ("pdep_%d.npy" % r)).view(np.uint64) == want_d.view(np.uint64)).all()
    full.Close()


def _nccl_lib_worker(rank, world, port, out_dir):
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    for p in (root, os.path.join(root, "tests")):
        if p not in sys.path:
            sys.path.insert(0, p)
    import torch
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("gloo", rank=rank, world_size=world)   # only carries the NCCL unique id
    try:
        import scenes
        from fauxgl_b200 import multigpu
        from fauxgl_b200.context import Context
        mesh = scenes.bumpy_mesh(201, 201)
        sc = scenes.dragon_scene(mesh, 1280, 720)
        ctx = Context(sc.width, sc.height, device=rank)
        comm = multigpu.NcclComposite(ctx, rank, world)
        first, count = multigpu.triangle_range(mesh.num_triangles, rank, world)

        class RangeCtx:
            def __init__(self, c):
                self.__dict__["c"] = c

            def __getattr__(self, k):
                return getattr(self.c, k)

            def __setattr__(self, k, v):
                setattr(self.c, k, v)

            def DrawMesh(self, m):
                return self.c.DrawTriangles(m, first, count)
        for root_rank in (0, -1):                    # gather to rank 0, then all-gather
            sc.run(RangeCtx(ctx))
            comm.composite(root_rank)
            ctx.Sync()
            np.save(os.path.join(out_dir, "nimg_%d_%d.npy" % (root_rank + 1, rank)), ctx.Image())
        comm.close()
        ctx.Close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_in_library_nccl_composite_two_gpus(tmp_path, gpu_capi):
    """fgl_comm_init / fgl_composite: reduce-scatter(min) by stripe + gather to the presenting rank (or all-gather),
    NCCL loaded by the library itself."""
    import torch
    import torch.multiprocessing as mp
    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    import scenes
    from fauxgl_b200.context import Context
    world = 2
    mp.spawn(_nccl_lib_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    mesh = scenes.bumpy_mesh(201, 201)
    sc = scenes.dragon_scene(mesh, 1280, 720)
    full = Context(sc.width, sc.height)
    sc.run(full)
    want = full.Image()
    budget = 1e-4 * want.shape[0] * want.shape[1]
    root_img = np.load(tmp_path / "nimg_1_0.npy")                 # root = 0: rank 0 holds the frame
    assert int((root_img != want).any(axis=-1).sum()) <= budget
    for r in range(world):                                          # root < 0: every rank does
        img = np.load(tmp_path / ("nimg_0_%d.npy" % r))
        assert int((img != want).any(axis=-1).sum()) <= budget
    full.Close()


# ---- one GPU is enough ----------------------------------------------------------------------------------------------

def _render_ranges(sc, mesh, ctxs, ranges):
    """Every 'rank' (a context of this process) draws its triangles of the scene."""
    class RangeCtx:
        def __init__(self, c, idx):
            self.__dict__["c"] = c
            self.__dict__["idx"] = idx

        def __getattr__(self, k):
            return getattr(self.c, k)

        def __setattr__(self, k, v):
            setattr(self.c, k, v)

        def DrawMesh(self, m):
            first, count = self.idx
            return self.c.DrawTriangles(m, first, count)
    infos = []
    for c, rg in zip(ctxs, ranges):
        infos.append(sc.run(RangeCtx(c, rg))[0])
    return infos


@pytest.mark.parametrize("world,root", [(2, -1), (3, 0), (4, 2), (8, -1), (8, 7)])
def test_peer_group_in_process_is_exact(world, root, gpu_capi):
    """fgl_peer_group over `world` contexts of this process: after the composite the presenting rank (every rank for
    root < 0) holds the single-context frame bit for bit -- float64 depth and colour --, frame after frame."""
    import scenes
    from fauxgl_b200 import multigpu
    from fauxgl_b200.context import Context
    mesh = scenes.bumpy_mesh(151, 151)
    sc = scenes.dragon_scene(mesh, 1000, 562)      # ragged: 1000 is not a multiple of the strip width
    full = Context(sc.width, sc.height)
    finfo = sc.run(full)
    want_c, want_d = full.Image(), full.DepthBuffer
    ctxs = [Context(sc.width, sc.height) for _ in range(world)]
    groups = multigpu.PeerGroup.local(ctxs)
    ranges = [multigpu.triangle_range(mesh.num_triangles, r, world) for r in range(world)]
    for _frame in range(3):
        infos = _render_ranges(sc, mesh, ctxs, ranges)
        multigpu.PeerGroup.composite_local(groups, root)   # enqueued on every rank's stream; nothing blocks on the host
        for g in groups:
            g.status()
        assert sum(i[0] for i in infos) == finfo[0][0]   # TotalPixels adds up over the ranks
        holders = range(world) if root < 0 else [root]
        for r in holders:
            assert (ctxs[r].Image() == want_c).all(), (world, root, r)
            assert (ctxs[r].DepthBuffer.view(np.uint64) == want_d.view(np.uint64)).all(), (world, root, r)
    # colours only (FGL_COMPOSITE_COLOR_ONLY): the holders get the frame's colours, their depth stays as they drew it
    _render_ranges(sc, mesh, ctxs, ranges)
    own_depth = [c.DepthBuffer for c in ctxs]
    multigpu.PeerGroup.composite_local(groups, root, color_only=True)
    for g in groups:
        g.status()
    for r in (range(world) if root < 0 else [root]):
        assert (ctxs[r].Image() == want_c).all(), (world, root, r)
        assert (ctxs[r].DepthBuffer.view(np.uint64) == own_depth[r].view(np.uint64)).all()
    for g in groups:
        g.close()
    for c in ctxs + [full]:
        c.Close()


def test_peer_group_continues_drawing_after_composite(gpu_capi):
    """The composite leaves consistent buffers AND dirty-strip bitmaps behind: a later draw into the composited frame
    and a second composite still give the single-context result."""
    import scenes
    from fauxgl_b200 import multigpu
    from fauxgl_b200.context import Context
    mesh = scenes.bumpy_mesh(101, 101)
    sc = scenes.dragon_scene(mesh, 640, 360)
    T = mesh.num_triangles
    full = Context(sc.width, sc.height)
    sc.run(full)
    ctxs = [Context(sc.width, sc.height) for _ in range(2)]
    groups = multigpu.PeerGroup.local(ctxs)
    # first half of the mesh split over the two ranks, composite to all; then the second half, composite again
    quarters = [(0, T // 4), (T // 4, T // 2 - T // 4), (T // 2, T // 4), (T // 2 + T // 4, T - T // 2 - T // 4)]
    _render_ranges(sc, mesh, ctxs, quarters[:2])
    multigpu.PeerGroup.composite_local(groups, -1)
    for c, (first, count) in zip(ctxs, quarters[2:]):
        c.DrawTriangles(mesh, first, count)          # no clear: on top of the composited frame
    multigpu.PeerGroup.composite_local(groups, -1)
    for g in groups:
        g.status()
    for c in ctxs:
        assert (c.Image() == full.Image()).all()
        assert (c.DepthBuffer.view(np.uint64) == full.DepthBuffer.view(np.uint64)).all()
    for g in groups:
        g.close()
    for c in ctxs + [full]:
        c.Close()


def test_in_library_nccl_single_rank(gpu_capi):
    """fgl_comm_unique_id / fgl_comm_init / fgl_composite with one rank: NCCL is found and loaded by the library, and
    pack -> (no exchange) -> unpack leaves the colour buffer unchanged and the depth buffer quantised to the key."""
    import scenes
    from fauxgl_b200 import multigpu
    from fauxgl_b200.context import Context
    sc = scenes.bumpy_small()
    ctx = Context(sc.width, sc.height)
    sc.run(ctx)
    before = ctx.Image()
    comm = multigpu.NcclComposite(ctx, 0, 1)
    comm.composite(0)
    comm.composite(-1)
    ctx.Sync()
    assert (ctx.Image() == before).all()
    comm.close()
    ctx.Close()
