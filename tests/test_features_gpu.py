"""GPU tests of the rest of the C ABI: the benchmark-size mesh, SSAA resolve, mesh
transform/update on the device, async draws, the composite kernels, a sort-last
composite on one GPU, and the error paths.  Every check is against the CPU
oracle (bit-exact) or a size-independent property."""
import ctypes as C

import numpy as np
import pytest

import scenes
from parity import compare_buffers, run_both

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Context(gpu_capi):
    from fauxgl_b200.context import Context
    return Context


@pytest.fixture(scope="module")
def m871k():
    from fauxgl_b200 import synth
    return synth.bumpy_surface()


def test_benchmark_mesh_1080p_bit_exact(m871k, oracle_lib, Context):
    """BASELINE config: 871 306 triangles, Phong, 1920x1080 -- full-size parity."""
    stats = run_both(scenes.dragon_scene(m871k, 1920, 1080), oracle_lib, Context)
    print(stats)
    assert stats["depth_mismatch"] == 0 and stats["color_mismatch"] == 0
    assert stats["gpu_info"] == stats["oracle_info"]
    assert stats["gpu_info"][0][0] > 500000


def test_benchmark_mesh_ssaa_8k_and_resolve_bit_exact(m871k, oracle_lib, Context):
    """Same mesh at 7680x4320 (16x SSAA) + the 4x resolve."""
    sc = scenes.dragon_scene(m871k, 7680, 4320)
    octx = oracle_lib.OracleContext(sc.width, sc.height)
    oinfo = sc.run(octx)
    gctx = Context(sc.width, sc.height)
    ginfo = sc.run(gctx)
    stats = compare_buffers(octx.ColorBuffer, octx.DepthBuffer, gctx.Image(), gctx.DepthBuffer)
    print(stats, oinfo, ginfo)
    assert stats["depth_mismatch"] == 0 and stats["color_mismatch"] == 0 and ginfo == oinfo
    got = gctx.Resolve(4)
    want = octx.Resolve(4)
    assert got.shape == (1080, 1920, 4) and (got == want).all()
    gctx.Close()


@pytest.mark.timeout(900)
def test_m10m_sphere_8k_bit_exact(oracle_lib, Context):
    """BASELINE config 5 on ONE GPU: the 10 003 864-triangle sphere at 7680x4320 against the oracle.  The oracle runs
    the reference's own threaded schedule (context.go:413-433 restated with pthreads), which in this
    order-independent state (opaque, DepthBias 0, depth test on) gives the sequential depth and colour
    (test_oracle_cpu.py::test_threaded_schedule_matches_sequential_when_order_free); TotalPixels is schedule
    independent, UpdatedPixels is not and is not compared.  Bit-exact float64 depth and NRGBA8 colour."""
    import os
    from fauxgl_b200 import synth
    mesh = synth.uv_sphere(2237, 2237)
    assert mesh.num_triangles == 10003864
    sc = scenes.dragon_scene(mesh, 7680, 4320)
    octx = oracle_lib.OracleContext(sc.width, sc.height, threads=os.cpu_count() or 1)
    oinfo = sc.run(octx)
    gctx = Context(sc.width, sc.height)
    gctx.upload_attributes = ("position", "normal")
    ginfo = sc.run(gctx)
    stats = compare_buffers(octx.ColorBuffer, octx.DepthBuffer, gctx.Image(), gctx.DepthBuffer)
    print(stats, oinfo, ginfo)
    assert stats["depth_mismatch"] == 0 and stats["color_mismatch"] == 0
    assert ginfo[0][0] == oinfo[0][0] and ginfo[0][0] > 5_000_000
    gctx.Close()


@pytest.mark.parametrize("factor,w,h", [(2, 130, 70), (4, 256, 128), (3, 99, 60), (8, 64, 64), (1, 33, 17)])
def test_resolve_matches_oracle(factor, w, h, oracle_lib, Context):
    rng = np.random.RandomState(factor)
    img = rng.randint(0, 256, (h * factor, w * factor, 4)).astype(np.uint8)
    img[: h * factor // 2, :, 3] = 255           # half opaque, half translucent (premultiply path)
    ctx = Context(w * factor, h * factor)
    ctx.UploadColorBuffer(img)
    got = ctx.Resolve(factor)
    want = oracle_lib.resolve(img, w, h)
    assert (got == want).all()
    ctx.Close()


@pytest.mark.parametrize("w,h", [(256, 128), (67, 33)])
def test_resolve_opaque_4x_fast_path_matches_oracle(w, h, oracle_lib, Context):
    """Opaque pixels at factor 4 take the DP4A fast path away from the left/right border (the generic loop
    elsewhere); a few translucent pixels inside force the generic loop in the middle of fast rows."""
    rng = np.random.RandomState(7)
    img = rng.randint(0, 256, (h * 4, w * 4, 4)).astype(np.uint8)
    img[..., 3] = 255
    img[5::37, 3::53, 3] = rng.randint(0, 255, img[5::37, 3::53, 3].shape)
    ctx = Context(w * 4, h * 4)
    ctx.UploadColorBuffer(img)
    got = ctx.Resolve(4)
    want = oracle_lib.resolve(img, w, h)
    assert (got == want).all()
    ctx.Close()


@pytest.mark.parametrize("front", ["fused", "split"])
def test_work_buffers_regrow_under_both_front_ends(front, oracle_lib, Context, monkeypatch):
    """Large triangles after a context sized for nothing: the segment capacity overflows, the synchronous draw
    regrows and re-issues (the fused front end reserves rows x strips per record, the split one counts)."""
    from fauxgl_b200 import NewSolidColorShader, NewTriangleMesh, Orthographic, Color
    monkeypatch.setenv("FGL_FRONT", front)
    ctx = Context(2048, 2048)
    octx = oracle_lib.OracleContext(2048, 2048)
    tri = np.array([[[-1, -1, 0], [1, -1, 0], [0, 1, 0]]] * 60, dtype=np.float64)
    tri[:, :, 2] = np.linspace(-0.5, 0.5, 60)[:, None]
    mesh = NewTriangleMesh(tri)
    for c in (ctx, octx):
        c.Shader = NewSolidColorShader(Orthographic(-1, 1, -1, 1, -1, 1), Color(0.2, 0.4, 0.6, 1))
        c.Cull = 1
    gi, oi = ctx.DrawTriangles(mesh), octx.DrawTriangles(mesh)
    assert tuple(gi) == oi and ctx.DrawStats().retries >= 1
    assert (ctx.DepthBuffer.view(np.uint64) == octx.DepthBuffer.view(np.uint64)).all()
    assert (ctx.Image() == octx.ColorBuffer).all()
    ctx.Close()


def test_mesh_transform_on_device_matches_host(Context):
    """fgl_mesh_transform == Mesh.Transform (mesh.go:167-175), bit for bit."""
    from fauxgl_b200 import Rotate, Radians, V
    mesh = scenes.load_fixture("bowser_mesh")
    mesh.BiUnitCube()
    ctx = Context(64, 64)
    dm = ctx.device_mesh(mesh)
    m = Rotate(V(0, 0, 1), Radians(5)).Translate(V(0.1, -0.2, 0.3))
    for _ in range(3):
        dm.Transform(m)
        mesh.Transform(m)
    pos, nrm, _lp, _ln = dm.read()
    assert (pos.view(np.uint64) == mesh.position.view(np.uint64)).all()
    assert (nrm.view(np.uint64) == mesh.normal.view(np.uint64)).all()
    ctx.Close()


def test_animation_frames_device_transform_vs_host_transform(oracle_lib, Context):
    """examples/animate.go:43-67: rotate 5 degrees per frame.  Device-side transform + draw equals
    host-side transform + re-upload + draw, frame by frame."""
    from fauxgl_b200 import Gray, HexColor, LookAt, NewPhongShader, Radians, Rotate, V
    mesh = scenes.load_fixture("bowser_mesh")
    mesh.BiUnitCube()
    eye, up = V(4, 4, 2), V(0, 0, 1)
    matrix = LookAt(eye, V(0, 0, 0), up).Perspective(30, 1.0, 1, 10)
    shader = NewPhongShader(matrix, V(0.25, 0.5, 1).Normalize(), eye)
    shader.ObjectColor = HexColor("#FEB41C")
    shader.DiffuseColor, shader.SpecularColor, shader.SpecularPower = Gray(0.9), Gray(0.25), 100
    a, b = Context(400, 400), Context(400, 400)
    o = oracle_lib.OracleContext(400, 400)
    a.Shader = b.Shader = o.Shader = shader
    dm = a.device_mesh(mesh.Copy())
    rot = Rotate(up, Radians(5))
    for frame in range(3):
        for c in (a, b, o):
            c.ClearDepthBuffer()
            c.ClearColorBufferWith(HexColor("#24221F"))
        ia = a.DrawMesh(dm)
        ib = b.DrawMesh(mesh)          # host mesh: re-uploaded in place after Transform (generation bump)
        io = o.DrawMesh(mesh)
        assert tuple(ia) == tuple(ib) == io
        assert (a.Image() == o.ColorBuffer).all() and (b.Image() == o.ColorBuffer).all()
        dm.Transform(rot)
        mesh.Transform(rot)
    a.Close(); b.Close()


def test_async_draws_accumulate_info(oracle_lib, Context):
    sc = scenes.shapes_multipass()
    octx = oracle_lib.OracleContext(sc.width, sc.height)
    oinfo = sc.run(octx)
    gctx = Context(sc.width, sc.height)
    sc.run(gctx)                       # sizes the work buffers
    gctx.ClearDepthBuffer()
    gctx.Wireframe, gctx.DepthBias = False, 0.0

    class AsyncCtx:                    # same scene script, async draw calls
        def __init__(self, c):
            self.__dict__["c"] = c

        def __getattr__(self, k):
            return getattr(self.c, k)

        def __setattr__(self, k, v):
            setattr(self.c, k, v)

        def DrawMesh(self, mesh):
            self.c.DrawMeshAsync(mesh)
            return (0, 0)
    sc.run(AsyncCtx(gctx))
    info = gctx.Sync()
    assert tuple(info) == (sum(i[0] for i in oinfo), sum(i[1] for i in oinfo))
    assert (gctx.Image() == octx.ColorBuffer).all()
    gctx.Close()


def test_work_buffers_regrow_transparently(oracle_lib, Context):
    """A tiny first draw sizes small buffers; a big-triangle draw must regrow them (sync path)."""
    from fauxgl_b200 import NewSolidColorShader, NewTriangleMesh, Orthographic, Color
    ctx = Context(2048, 2048)
    octx = oracle_lib.OracleContext(2048, 2048)
    tri = np.array([[[-1, -1, 0], [1, -1, 0], [0, 1, 0]]] * 40, dtype=np.float64)
    tri[:, :, 2] = np.linspace(-0.5, 0.5, 40)[:, None]
    mesh = NewTriangleMesh(tri)
    for c in (ctx, octx):
        c.Shader = NewSolidColorShader(Orthographic(-1, 1, -1, 1, -1, 1), Color(0.2, 0.4, 0.6, 1))
        c.Cull = 1
    gi, oi = ctx.DrawTriangles(mesh), octx.DrawTriangles(mesh)
    assert tuple(gi) == oi and ctx.DrawStats().retries >= 1
    assert (ctx.DepthBuffer.view(np.uint64) == octx.DepthBuffer.view(np.uint64)).all()
    ctx.Close()


def test_composite_kernels_match_oracle_keys(oracle_lib, Context):
    import torch
    sc = scenes.hello()
    ctx = Context(sc.width, sc.height)
    sc.run(ctx)
    keys = torch.empty(sc.width * sc.height, dtype=torch.int64, device="cuda")
    ctx.CompositePack(keys.data_ptr())
    ctx.Sync()
    torch.cuda.synchronize()
    color, depth = ctx.Image(), ctx.DepthBuffer
    got = keys.cpu().numpy().view(np.uint64).reshape(sc.height, sc.width)
    rng = np.random.RandomState(0)
    for _ in range(2000):
        y, x = rng.randint(sc.height), rng.randint(sc.width)
        assert int(got[y, x]) == oracle_lib.pack_key(float(depth[y, x]), color[y, x].tolist())
    # unpack(pack(.)) keeps colour exactly and depth to 32 bits
    ctx.CompositeUnpack(keys.data_ptr())
    assert (ctx.Image() == color).all()
    d2 = ctx.DepthBuffer
    cov = depth < 1e300
    assert (d2[~cov] == depth[~cov]).all() and np.abs(d2[cov] - depth[cov]).max() < 2.4e-10
    ctx.Close()


@pytest.mark.parametrize("parts", [2, 4, 8])
def test_sort_last_composite_on_one_gpu(parts, oracle_lib, Context):
    """N triangle ranges -> N full-frame renders -> packed-key min == one render of the whole mesh
    (up to depth ties / 32-bit depth quantisation; the mismatch count is reported)."""
    import torch
    from fauxgl_b200 import multigpu
    mesh = scenes.bumpy_mesh(201, 201)
    sc = scenes.dragon_scene(mesh, 1280, 720)
    full = Context(sc.width, sc.height)
    finfo = sc.run(full)
    n = sc.width * sc.height
    acc = torch.empty(n, dtype=torch.int64, device="cuda")
    tmp = torch.empty(n, dtype=torch.int64, device="cuda")
    total = 0
    part = Context(sc.width, sc.height)
    for r in range(parts):
        first, count = multigpu.triangle_range(mesh.num_triangles, r, parts)

        class RangeCtx:
            def __init__(self, c):
                self.__dict__["c"] = c

            def __getattr__(self, k):
                return getattr(self.c, k)

            def __setattr__(self, k, v):
                setattr(self.c, k, v)

            def DrawMesh(self, m):
                return self.c.DrawTriangles(m, first, count)
        info = sc.run(RangeCtx(part))
        total += info[0][0]
        part.CompositePack((acc if r == 0 else tmp).data_ptr())
        if r:
            part.CompositeMin(acc.data_ptr(), tmp.data_ptr(), n)
        part.Sync()
    part.CompositeUnpack(acc.data_ptr())
    assert total == finfo[0][0]                                   # TotalPixels adds up
    a, b = part.Image(), full.Image()
    mism = int((a != b).any(axis=-1).sum())
    print("sort-last parts=%d colour mismatches: %d of %d" % (parts, mism, n))
    assert mism <= 1e-4 * n
    full.Close(); part.Close()


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_peer_composite_kernel_is_exact(parts, oracle_lib, Context):
    """fgl_composite_peer (the fused P2P composite) on one GPU: N contexts stand in for N ranks.  It keeps
    the float64 depth and breaks ties towards the later triangle range, so every 'rank' must end up with
    exactly the single-context render: bit-identical depth AND colour, no tolerance."""
    from fauxgl_b200 import multigpu
    mesh = scenes.bumpy_mesh(201, 201)
    sc = scenes.dragon_scene(mesh, 1283, 717)        # not a multiple of the tile size or of `parts`
    full = Context(sc.width, sc.height)
    finfo = sc.run(full)
    ctxs = [Context(sc.width, sc.height) for _ in range(parts)]
    total = 0
    for r, c in enumerate(ctxs):
        first, count = multigpu.triangle_range(mesh.num_triangles, r, parts)

        class RangeCtx:
            def __init__(self, c):
                self.__dict__["c"] = c

            def __getattr__(self, k):
                return getattr(self.c, k)

            def __setattr__(self, k, v):
                setattr(self.c, k, v)

            def DrawMesh(self, m):
                return self.c.DrawTriangles(m, first, count)
        total += sc.run(RangeCtx(c))[0][0]
    for c in ctxs:
        c.Sync()
    colors = [c.color_ptr for c in ctxs]
    depths = [c.depth_ptr for c in ctxs]
    for r, c in enumerate(ctxs):
        c.CompositePeer(r, colors, depths)           # each 'rank' composites its own stripe into all buffers
    for c in ctxs:
        c.Sync()
    assert total == finfo[0][0]
    want_c, want_d = full.Image(), full.DepthBuffer
    for c in ctxs:
        assert (c.DepthBuffer.view(np.uint64) == want_d.view(np.uint64)).all()
        assert (c.Image() == want_c).all()
        c.Close()
    full.Close()


def test_error_paths(gpu_capi, Context):
    from fauxgl_b200 import context, NewTriangleMesh, Identity
    with pytest.raises(context.FauxglError):
        Context(0, 10)
    with pytest.raises(context.FauxglError):
        Context(16, 16, device=99)
    ctx = Context(16, 16)
    mesh = NewTriangleMesh(np.zeros((2, 3, 3)))
    with pytest.raises(context.FauxglError):
        ctx.DrawTriangles(mesh, 1, 5)                  # range outside the mesh
    with pytest.raises(context.FauxglError):
        ctx.Resolve(3)                                 # 16 % 3 != 0

    class Custom:                                      # user-defined Shader: no device equivalent
        def describe(self):
            return {"kind": 7, "matrix": tuple(Identity())}
    ctx.Shader = Custom()
    with pytest.raises(context.FauxglError) as e:
        ctx.DrawMesh(mesh)
    assert e.value.status == -4
    # the C ABI itself rejects an unknown shader kind too
    st, sh = ctx._state(), context._Shader()
    sh.kind = 42
    info = context._Info()
    dm = ctx.device_mesh(mesh)
    rc = gpu_capi.fgl_draw_triangles(ctx._h, C.byref(st), C.byref(sh), dm.handle, 0, 2, C.byref(info))
    assert rc == -4 and b"no CPU fallback" in gpu_capi.fgl_last_error(ctx._h)
    ctx.Close()


def _stacked_scene(kind, state):
    """Hundreds of triangles piled onto the same few strips, so that a handful of strips hold far more than 96
    segments each: the strip kernel hands those to whole CTAs (fgl_raster.cu, heavy strips).  kind "wide":
    every segment spans its strip (a 32-segment chunk has more fragments than can be staged: the in-order
    fallback); kind "narrow": slivers of a few pixels (staged chunks replayed by warp 0).  Depths are drawn
    from a small set, so the <= tie rule decides many pixels."""
    from fauxgl_b200 import (CullNone, Gray, HexColor, NewPhongShader, NewTriangleMesh, Orthographic, V)
    # 640 x 360: 7 200 strips of 32 pixels, so the strip kernel runs its full grid (two CTAs per SM) and the
    # ~100-200 heavy strips below stay under one per CTA -- the condition for the cooperative path
    W, H = 640, 360
    rng = np.random.RandomState(11 if kind == "wide" else 12)
    n = 260
    tri = []
    for k in range(n):
        z = -0.1 * rng.randint(0, 6)                       # few distinct depths: ties
        if kind == "wide":
            x0, x1 = -0.12 + 0.004 * rng.rand(), 0.33 + 0.004 * rng.rand()
            y0 = -0.1 + 0.0003 * k
            tri.append([(x0, y0, z), (x1, y0, z - 0.05 * rng.rand()), (0.5 * (x0 + x1), y0 + 0.2, z)])
        else:
            cx = -0.3 + 0.004 * rng.rand()
            tri.append([(cx, -0.12, z), (cx + 0.012 + 0.006 * rng.rand(), -0.12, z), (cx + 0.005, 0.12, z - 0.03 * rng.rand())])
    mesh = NewTriangleMesh(np.array(tri, dtype=np.float64))

    def run(ctx):
        ctx.ClearColorBufferWith(Gray(0.25))
        shader = NewPhongShader(Orthographic(-1, 1, -1, 1, -1, 1), V(0.3, 0.2, 1).Normalize(), V(0, 0, 5))
        shader.ObjectColor = HexColor("#468966")
        ctx.Shader = shader
        ctx.Cull = CullNone
        if state == "bias":
            ctx.DepthBias = -1e-3
        elif state == "no_read":
            ctx.ReadDepth = False
        elif state == "no_write":
            ctx.WriteDepth = False
        return [ctx.DrawMesh(mesh), ctx.DrawTriangles(mesh, 17, 200)]
    sc = scenes.Scene(W, H, run)
    sc.mesh = mesh
    return sc


@pytest.mark.parametrize("state", ["default", "bias", "no_read", "no_write"])
@pytest.mark.parametrize("kind", ["wide", "narrow"])
def test_heavy_strips_taken_by_whole_ctas(kind, state, oracle_lib, Context):
    """The CTA-cooperative path of the strip kernel (a few strips with hundreds of segments) is bit-exact for
    every render state, for chunks that are staged and for chunks that are not."""
    stats = run_both(_stacked_scene(kind, state), oracle_lib, Context)
    print(stats)
    assert stats["depth_mismatch"] == 0 and stats["color_mismatch"] == 0
    assert stats["gpu_info"] == stats["oracle_info"]
    assert stats["gpu_info"][0][0] > 15000


@pytest.mark.parametrize("kind", ["wide", "narrow"])
def test_heavy_strips_per_primitive_info(kind, oracle_lib, Context):
    """fgl_draw_triangles_each through the same path: UpdatedPixels attributed per triangle."""
    sc = _stacked_scene(kind, "default")
    ctx, o = Context(sc.width, sc.height), oracle_lib.OracleContext(sc.width, sc.height)

    class Capture:
        def __init__(self, c):
            self.__dict__["c"] = c

        def __getattr__(self, k):
            return getattr(self.c, k)

        def __setattr__(self, k, v):
            setattr(self.c, k, v)

        def DrawMesh(self, m):
            return (0, 0)

        def DrawTriangles(self, m, *a):
            return (0, 0)
    sc.run(Capture(ctx)); sc.run(Capture(o))
    got, want = ctx.DrawTrianglesEach(sc.mesh), o.DrawTrianglesEach(sc.mesh)
    assert (got == want).all() and int(got[:, 1].sum()) > 1000
    assert (ctx.Image() == o.ColorBuffer).all()
    assert (ctx.DepthBuffer.view(np.uint64) == o.DepthBuffer.view(np.uint64)).all()
    ctx.Close()


@pytest.mark.parametrize("scene_name,resolve", [("bumpy_small", 2), ("shapes_multipass", 0), ("lines", 0)])
def test_recorded_frame_replays_identically(scene_name, resolve, oracle_lib, Context):
    """fgl_graph_*: a frame recorded into a CUDA graph (clears, async draws -- fused and split front ends, deferred and
    inline shading, lines --, the resolve) and replayed gives the buffers, the RasterizeInfo and the resolved image of
    the calls themselves, replay after replay, also when other work ran on the context in between."""
    sc = scenes.SCENES[scene_name]()

    class Recorder:
        """Runs the scene script with DrawMesh / DrawTriangles / DrawLines turned into their async forms."""
        def __init__(self, c):
            self.__dict__["c"] = c

        def __getattr__(self, k):
            return getattr(self.c, k)

        def __setattr__(self, k, v):
            setattr(self.c, k, v)

        def DrawMesh(self, m):
            self.c.DrawMeshAsync(m)
            return (0, 0)

        def DrawTriangles(self, m, first=0, count=None):
            self.c.DrawMeshAsync(self.c.device_mesh(m), first, self.c.device_mesh(m).num_triangles - first if count is None else count)
            return (0, 0)

        def DrawLines(self, m, first=0, count=None):
            dm = self.c.device_mesh(m)
            st, sh = self.c._state(), self.c._shader()
            from fauxgl_b200.context import _check, capi
            n = dm.num_lines - first if count is None else count
            _check(capi().fgl_draw_lines_async(self.c._h, C.byref(st), C.byref(sh), dm.handle, first, n), self.c._h)
            return (0, 0)
    ref = Context(sc.width, sc.height)
    rinfo = sc.run(ref)
    want_c, want_d = ref.Image(), ref.DepthBuffer
    want_r = ref.Resolve(resolve) if resolve else None
    ctx = Context(sc.width, sc.height)
    sc.run(ctx)                                  # a first, direct frame: uploads the meshes, sizes the work buffers
    ctx2 = ctx                                   # the script mutates Wireframe / DepthBias / Cull ...: back to NewContext's state
    ctx2.ReadDepth = ctx2.WriteDepth = ctx2.WriteColor = ctx2.AlphaBlend = True
    ctx2.Wireframe, ctx2.FrontFace, ctx2.Cull, ctx2.LineWidth, ctx2.DepthBias = False, 2, 3, 2.0, 0.0
    if resolve:
        ctx2.ResolveDevice(resolve)               # its buffer is allocated on first use: not while recording
    ctx2.ClearDepthBuffer()
    ctx2.GraphBegin()
    ctx2.ClearDepthBuffer()
    sc.run(Recorder(ctx2))
    if resolve:
        ctx2.ResolveDevice(resolve)
    graph = ctx2.GraphEnd()
    total = sum(i[0] for i in rinfo), sum(i[1] for i in rinfo)
    for rep in range(3):
        graph.launch()
        info = ctx2.Sync()
        assert tuple(info) == total, (rep, info, total)
        assert (ctx2.Image() == want_c).all() and (ctx2.DepthBuffer.view(np.uint64) == want_d.view(np.uint64)).all(), rep
        if resolve:
            out = np.empty_like(want_r)
            from fauxgl_b200.context import _check, capi
            _check(capi().fgl_read_resolved(ctx2._h, out.ctypes.data), ctx2._h)
            assert (out == want_r).all()
        if rep == 0:                              # unrelated direct work between two replays
            ctx2.ClearColorBufferWith(scenes.HexColor("#123456"))
            ctx2.ClearDepthBufferWith(0.25)
    # what cannot be recorded says so
    ctx2.GraphBegin()
    with pytest.raises(Exception):
        ctx2.Image()
    ctx2.GraphEnd()
    # a recording is refused once the buffers it points at have been reallocated (here: another resolve size)
    if resolve:
        ctx2.ResolveDevice(1)
        with pytest.raises(Exception):
            graph.launch()
    graph.Close()
    ctx.Close(); ref.Close()


def test_branch_free_division_matches_the_operator(gpu_capi):
    """k_front divides with the fast path of ptxas's own div.rn.f64 / rcp.rn.f64 expansions written as straight-line
    code (one reciprocal refinement per denominator, no branch between independent divisions) and falls back to the
    operator when the expansion's range test fails.  Wherever it does not fall back, every bit must equal `a / b` and
    `1 / b`: 2^27 operand pairs per seed -- raw bit patterns, NaN, infinities, subnormals, zeros, rasteriser-sized values."""
    from fauxgl_b200.context import Context
    ctx = Context(64, 64)
    total_fast = 0
    for seed in (1, 2, 0xfa0c61):
        bad, fast = ctx.DivCheck(seed, 1 << 27)
        assert bad == 0, (seed, bad, fast)
        total_fast += fast
    assert total_fast > (3 << 27) // 2   # most pairs take the fast path
    ctx.Close()
