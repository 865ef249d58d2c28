"""Two-GPU test of the sort-last path over NCCL: every rank draws its triangle range
on its own B200, the packed keys are min-reduced with torch.distributed (NCCL over
NVLink) and every rank must end up with the single-GPU image (up to depth ties).
Skipped on a one-GPU box; the same flow runs on one GPU in
test_features_gpu.py::test_sort_last_composite_on_one_gpu."""
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
        pc = multigpu.PeerComposite(ctx, rank, world)
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
        for _frame in range(2):                      # twice: the fences must make re-use safe
            sc.run(RangeCtx(ctx))
            pc.composite()
        np.save(os.path.join(out_dir, "pimg_%d.npy" % rank), ctx.Image())
        np.save(os.path.join(out_dir, "pdep_%d.npy" % rank), ctx.DepthBuffer)
        pc.close()
        ctx.Close()
    finally:
        dist.destroy_process_group()


@pytest.mark.timeout(600)
def test_peer_memory_composite_two_gpus_exact(tmp_path, gpu_capi):
    """The fused P2P composite over CUDA IPC + NVLink: both ranks end with the single-GPU frame, bit for bit."""
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
