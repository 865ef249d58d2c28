"""GPU tests of the data formats either side of the draw: binary STL straight to
the device layout, bounding box / BiUnitCube on the device, per-primitive RasterizeInfo
(Context.DrawLine / DrawTriangle return values) and Context.DepthImage.  All checks are
bit-exact against the CPU oracle."""
import struct

import numpy as np
import pytest

import scenes

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def Context(gpu_capi):
    from fauxgl_b200.context import Context
    return Context


def stl_bytes(position: np.ndarray) -> bytes:
    """A binary STL file holding the (T,3,3) positions rounded to float32 (normals left zero, as many writers do)."""
    t = len(position)
    rec = np.zeros((t, 50), dtype=np.uint8)
    rec[:, 12:48] = np.ascontiguousarray(position, dtype="<f4").reshape(t, 9).view(np.uint8).reshape(t, 36)
    return b"\0" * 80 + struct.pack("<I", t) + rec.tobytes()


def same_bits(a, b):
    a, b = np.ascontiguousarray(a, dtype=np.float64), np.ascontiguousarray(b, dtype=np.float64)
    return a.shape == b.shape and bool((a.view(np.uint64) == b.view(np.uint64)).all())


def test_binary_stl_to_device_matches_loader(oracle_lib, Context, tmp_path):
    """fgl_mesh_create_stl == loadSTLB (stl.go:86-154): widened float32 positions, face normals."""
    from fauxgl_b200 import mesh as fmesh
    from fauxgl_b200.context import DeviceMesh
    src = scenes.load_fixture("bowser_mesh")
    pos = src.position.copy()
    pos[5] = pos[5][0]            # a degenerate triangle: its normal is NaN in the reference too
    data = stl_bytes(pos)
    want_p, want_n = oracle_lib.stl_triangles(data[84:])
    path = tmp_path / "m.stl"
    path.write_bytes(data)
    host = fmesh.LoadSTL(str(path))
    assert same_bits(host.position, want_p) and same_bits(host.normal, want_n)
    ctx = Context(32, 32)
    dm = DeviceMesh.FromSTL(ctx, str(path))
    assert (dm.num_triangles, dm.num_lines) == (len(pos), 0)
    got_p, got_n, _lp, _ln = dm.read()
    assert same_bits(got_p, want_p)
    assert same_bits(got_n, want_n)
    # odd sizes: the staging loop works on 256-record chunks
    for n in (0, 1, 255, 257):
        d2 = DeviceMesh.FromSTL(ctx, stl_bytes(pos[:n]))
        p2, n2, _a, _b = d2.read()
        assert same_bits(p2, want_p[:n]) and same_bits(n2, want_n[:n])
        d2.Close()
    ctx.Close()


def test_device_bounds_and_biunitcube_match_host(Context):
    """fgl_mesh_bounds == Mesh.BoundingBox; DeviceMesh.BiUnitCube == Mesh.BiUnitCube (mesh.go:122-165)."""
    from fauxgl_b200 import synth
    from fauxgl_b200.context import DeviceMesh
    ctx = Context(32, 32)
    mesh = scenes.load_fixture("bowser_mesh")
    mesh.Add(synth.NewCubeOutline(-3.0, -0.0, -1.0, 0.5, 2.0, 7.0))   # lines count too (mesh.go:159-161)
    dm = DeviceMesh(ctx, mesh)
    bb, hb = dm.BoundingBox(), mesh.BoundingBox()
    assert tuple(bb.Min) == tuple(hb.Min) and tuple(bb.Max) == tuple(hb.Max)
    m_dev = dm.BiUnitCube()
    m_host = mesh.BiUnitCube()
    assert tuple(m_dev) == tuple(m_host)
    pos, nrm, lpos, _ln = dm.read()
    assert same_bits(pos, mesh.position) and same_bits(nrm, mesh.normal) and same_bits(lpos, mesh.lposition)
    empty = DeviceMesh.FromSTL(ctx, stl_bytes(np.zeros((0, 3, 3))))
    eb = empty.BoundingBox()
    assert tuple(eb.Min) == (0.0, 0.0, 0.0) and tuple(eb.Max) == (0.0, 0.0, 0.0)
    ctx.Close()


def test_stl_device_pipeline_renders_like_host_pipeline(oracle_lib, Context):
    """LoadSTL -> BiUnitCube -> DrawMesh with the mesh never materialised on the host equals the
    reference's host pipeline (examples/hello.go shape), pixels and RasterizeInfo."""
    from fauxgl_b200 import Gray, HexColor, LookAt, NewPhongShader, V
    from fauxgl_b200 import mesh as fmesh
    from fauxgl_b200.context import DeviceMesh
    data = stl_bytes(scenes.load_fixture("bowser_mesh").position)
    host = fmesh._load_stl_binary(data, (len(data) - 84) // 50)
    host.BiUnitCube()
    eye = V(3, 1, 0.75)
    matrix = LookAt(eye, V(0, 0, 0), V(0, 0, 1)).Perspective(40, 640 / 480, 1, 10)
    shader = NewPhongShader(matrix, V(0.75, 0.25, 1).Normalize(), eye)
    shader.ObjectColor = HexColor("#468966")
    shader.SpecularColor, shader.SpecularPower = Gray(0.3), 32
    ctx, o = Context(640, 480), oracle_lib.OracleContext(640, 480)
    dm = DeviceMesh.FromSTL(ctx, data)
    dm.BiUnitCube()
    for c in (ctx, o):
        c.ClearColorBufferWith(HexColor("#FFF8E3"))
        c.Shader = shader
    assert tuple(ctx.DrawMesh(dm)) == o.DrawMesh(host)
    assert (ctx.Image() == o.ColorBuffer).all()
    assert same_bits(ctx.DepthBuffer, o.DepthBuffer)
    ctx.Close()


def test_depth_image_matches_oracle(oracle_lib, Context):
    """Context.DepthImage, context.go:87-117."""
    sc = scenes.hello()
    ctx, o = Context(sc.width, sc.height), oracle_lib.OracleContext(sc.width, sc.height)
    # nothing drawn: every pixel is math.MaxFloat64 -> 0xffff
    assert (ctx.DepthImage() == 0xffff).all() and (o.DepthImage() == 0xffff).all()
    sc.run(ctx); sc.run(o)
    got, want = ctx.DepthImage(), o.DepthImage()
    assert got.shape == want.shape == (sc.height, sc.width)
    assert (got == want).all()
    assert got.min() == 0 and len(np.unique(got)) > 1000
    # hand-made buffers: negative depths, a single value (0/0 -> NaN -> 0), a NaN pixel
    rng = np.random.default_rng(5)
    for kind in range(3):
        d = np.full((sc.height, sc.width), np.finfo(np.float64).max)
        if kind == 0:
            d[10:200, 30:900] = rng.uniform(-3.0, 7.0, size=(190, 870))
            d[0, 0] = -0.0
        elif kind == 1:
            d[50:60, :] = 0.25
        else:
            d[10:200, 30:900] = rng.uniform(0.0, 1.0, size=(190, 870))
            d[11, 31] = np.nan
        ctx.UploadDepthBuffer(d)
        assert (ctx.DepthImage() == oracle_lib.depth_image(d)).all(), kind
    ctx.Close()


@pytest.mark.parametrize("front", ["split", "fused"])
def test_per_line_rasterize_info_matches_sequential_drawline(front, oracle_lib, Context, monkeypatch):
    """fgl_draw_lines_each == a loop of Context.DrawLine (examples/silhouette.go:163-166): occluder first,
    then biased lines whose UpdatedPixels/TotalPixels ratio decides visibility.  Under both front ends."""
    monkeypatch.setenv("FGL_FRONT", front)
    from fauxgl_b200 import HexColor, LookAt, NewSolidColorShader, Scale, V, White, Black, synth  # noqa: F401
    from fauxgl_b200.mesh import Mesh
    cube = synth.NewCube()
    cube.Transform(Scale(V(1.6, 1.6, 1.6)))
    lines = Mesh()
    for off in [(-0.8, -0.8, -0.8, 0.8, 0.8, 0.8), (-3, -3, -0.2, 3, 3, 0.2), (-0.2, -6, -1.5, 0.2, 6, 1.5),
                (-0.5, -0.5, -0.5, 0.5, 0.5, 0.5)]:
        lines.Add(synth.NewCubeOutline(*off))
    eye = V(3, 2.5, 1.8)
    matrix = LookAt(eye, V(0, 0, 0), V(0, 0, 1)).Perspective(50, 900 / 500, 1, 20)
    ctx, o = Context(900, 500), oracle_lib.OracleContext(900, 500)
    for c in (ctx, o):
        c.ClearColorBufferWith(White)
        c.Shader = NewSolidColorShader(matrix, HexColor("#7E827A"))
        c.DrawMesh(cube)
        c.Shader = NewSolidColorShader(matrix, Black)
        c.DepthBias = -1e-5
        c.LineWidth = 3
    got = ctx.DrawLinesEach(lines)
    want = o.DrawLinesEach(lines)
    assert got.shape == want.shape == (lines.num_lines, 2)
    assert (got == want).all()
    ratio = got[:, 1] / np.maximum(got[:, 0], 1)
    assert (ratio < 0.666).any() and (ratio >= 0.666).any()   # some hidden, some visible
    assert (got[:, 0] == 0).any()                              # fully clipped lines report zeros
    assert (ctx.Image() == o.ColorBuffer).all()
    # a sub-range, after the buffers already hold the lines
    got2, want2 = ctx.DrawLinesEach(lines, 7, 20), o.DrawLinesEach(lines, 7, 20)
    assert (got2 == want2).all()
    assert ctx.DrawLinesEach(lines, 3, 0).shape == (0, 2)
    ctx.Close()


@pytest.mark.parametrize("front", ["split", "fused"])
@pytest.mark.parametrize("wireframe", [False, True])
def test_per_triangle_rasterize_info_with_clipping(wireframe, front, oracle_lib, Context, monkeypatch):
    """fgl_draw_triangles_each == a loop of Context.DrawTriangle (context.go:370-389), including triangles
    the clipper splits (their fan triangles count towards the source triangle) and wireframe mode."""
    monkeypatch.setenv("FGL_FRONT", front)
    sc = scenes.bowser_close()
    ctx, o = Context(sc.width, sc.height), oracle_lib.OracleContext(sc.width, sc.height)
    mesh = {}

    class Capture:
        def __init__(self, c):
            self.__dict__["c"] = c

        def __getattr__(self, k):
            return getattr(self.c, k)

        def __setattr__(self, k, v):
            setattr(self.c, k, v)

        def DrawMesh(self, m):
            mesh["m"] = m
            self.c.Wireframe = wireframe
            self.c.LineWidth = 1.5
            return (0, 0)
    sc.run(Capture(ctx)); sc.run(Capture(o))
    m = mesh["m"]
    got, want = ctx.DrawTrianglesEach(m), o.DrawTrianglesEach(m)
    assert (got == want).all()
    assert int(got[:, 0].sum()) > 100000
    total = ctx.DrawStats()
    assert total.clip_triangles > 0   # the scene does exercise the clipper
    assert (ctx.Image() == o.ColorBuffer).all()
    ctx.Close()


def test_frame_pipeline_matches_synchronous_frames(oracle_lib, Context):
    """fgl_mesh_update_async + async draws + fgl_frame_end/fgl_fence_wait: frames that overlap upload, draw
    and read-back (examples/animate.go:43-67 shape: the host re-poses the mesh every frame) are identical to
    the same frames rendered one synchronous call at a time, and to the oracle."""
    from fauxgl_b200 import Gray, HexColor, LookAt, NewPhongShader, Radians, Rotate, V
    from fauxgl_b200.context import pinned_empty
    from fauxgl_b200.mesh import Mesh
    from fauxgl_b200.pipeline import FramePipeline
    base = scenes.load_fixture("bowser_mesh")
    base.BiUnitCube()
    eye, up = V(4, 4, 2), V(0, 0, 1)
    matrix = LookAt(eye, V(0, 0, 0), up).Perspective(30, 1.0, 1, 10)
    shader = NewPhongShader(matrix, V(0.25, 0.5, 1).Normalize(), eye)
    shader.ObjectColor = HexColor("#FEB41C")
    shader.DiffuseColor, shader.SpecularColor, shader.SpecularPower = Gray(0.9), Gray(0.25), 100
    bg = HexColor("#24221F")
    nframes, depth = 7, 2
    # host-side poses, each in its own pinned arrays (they must stay unchanged while in flight)
    poses = []
    m = base.Copy()
    for _ in range(nframes):
        pp, pn = pinned_empty(m.position.shape, np.float64), pinned_empty(m.normal.shape, np.float64)
        pp[...] = m.position
        pn[...] = m.normal
        poses.append(Mesh(pp, pn))
        m.Transform(Rotate(up, Radians(25)))
    ctx = Context(500, 400)
    ctx.Shader = shader
    pipe = FramePipeline(ctx, poses[0], depth=depth)
    got = []
    for k in range(nframes):
        if len(pipe) == depth:
            img, info = pipe.collect()
            got.append((img.copy(), tuple(info)))
        pipe.submit(poses[k], bg)
    while len(pipe):
        img, info = pipe.collect()
        got.append((img.copy(), tuple(info)))
    assert len(got) == nframes
    ref = Context(500, 400)
    ref.Shader = shader
    o = oracle_lib.OracleContext(500, 400)
    o.Shader = shader
    for k in range(nframes):
        ref.ClearDepthBuffer()
        ref.ClearColorBufferWith(bg)
        info = ref.DrawMesh(poses[k])
        assert tuple(info) == got[k][1], k
        assert (ref.Image() == got[k][0]).all(), k
        if k in (0, nframes - 1):
            o.ClearDepthBuffer()
            o.ClearColorBufferWith(bg)
            assert o.DrawMesh(poses[k]) == got[k][1]
            assert (o.ColorBuffer == got[k][0]).all()
    assert len({g[1] for g in got}) > 1   # the frames do differ
    ctx.Close(); ref.Close()


def test_device_smooth_normals_match_host_bit_for_bit(Context):
    """fgl_mesh_smooth_normals == Mesh.SmoothNormals (mesh.go:105-120): sums in corner order from the zero vector,
    groups by exact position (+0 == -0, NaN never equal).  The host mirror is itself checked against a naive
    dictionary implementation in test_host_cpu.py."""
    from fauxgl_b200 import mesh as fmesh, synth
    from fauxgl_b200.context import DeviceMesh
    ctx = Context(64, 64)
    cases = []
    # the benchmark surface: ~6 corners per position, two poles with 1320 corners each
    cases.append(synth.bumpy_surface(triangles=60000, nu=201, nv=201, smooth=False))
    # a loaded mesh with flat face normals
    cases.append(scenes.load_fixture("bowser_mesh"))
    # hand-made: -0 / +0 coordinates share a position, a NaN position is never found again (zero normal), a
    # degenerate sum normalises to NaN
    pos = np.array([[[0.0, 0, 0], [1, 0, 0], [0, 1, 0]],
                    [[-0.0, 0, 0], [0, 1, 0], [-1, 0, 0]],
                    [[np.nan, 0, 0], [1, 0, 0], [0, -0.0, 0.0]],
                    [[np.nan, 0, 0], [2, 2, 2], [2, 2, 2]]], dtype=np.float64)
    nrm = np.array([[[0, 0, 1.0], [0, 0, 1], [0, 1, 0]],
                    [[0, 0, -1.0], [0, -1, 0], [1, 0, 0]],
                    [[1, 0, 0.0], [0, 1, 0], [0.5, 0.25, -0.0]],
                    [[0, 1, 0.0], [0, 0, 1], [0, 0, 3]]], dtype=np.float64)
    cases.append(fmesh.NewTriangleMesh(pos, nrm, fix_normals=False))
    for m in cases:
        dm = DeviceMesh(ctx, m, ("position", "normal"))
        dm.SmoothNormals()
        _, dn, _, _ = dm.read()
        host = fmesh.Mesh(m.position.copy(), m.normal.copy())
        host.SmoothNormals()
        assert same_bits(dn, host.normal), "device SmoothNormals differs from the host restatement"
        nanpos = np.isnan(m.position).any(axis=2)
        assert (host.normal[nanpos] == 0).all()   # Go: lookup[NaN key] -> Vector{}
        # SmoothNormalsThreshold (mesh.go:80-103) on the same inputs
        from fauxgl_b200 import Radians
        for deg in (30.0, 75.0):
            dm = DeviceMesh(ctx, m, ("position", "normal"))
            dm.SmoothNormalsThreshold(Radians(deg))
            _, dn, _, _ = dm.read()
            host = fmesh.Mesh(m.position.copy(), m.normal.copy())
            host.SmoothNormalsThreshold(Radians(deg))
            assert same_bits(dn, host.normal), "device SmoothNormalsThreshold differs from the host restatement"
            assert np.isnan(host.normal[nanpos]).all()   # Go: nil list -> Vector{}.Normalize() = NaN
            dm.Close()
        dm.Close()
    ctx.Close()


OBJ_TEXT = """# a box with quads, a pentagon fan, negative indices, missing vt / vn, a zero normal
v -1 -1 -1
v  1 -1 -1
v  1  1 -1
v -1  1 -1
v -1 -1  1
v  1 -1  1
v  1  1  1
v -1  1  1
v  0  2  0.5
vt 0 0
vt 1 0
vt 1 1
vt 0 1
vt 0.5 0.25
vn 0 0 -1
vn 0 0 1
vn 0 -0 0
vn 0.6 0 0.8
f 1/1/1 4/4/1 3/3/1 2/2/1
f 5/1/2 6/2/2 7/3/2 8/4/2
f 1//3 2//3 6//3 5//3
f 2/2 3/3 7/4 6/1
f -6 -2 -1 -3 -7
f 4/4/4 8/1/4 7/5/-1
f 1 5 8 4
"""


def test_obj_indexed_ingest_matches_host_loader(Context, tmp_path):
    """fgl_mesh_create_indexed == the expansion of LoadOBJ (obj.go:58-74) + Triangle.FixNormals
    (triangle.go:46-58): positions, normals and a texture-shaded render are bit-identical."""
    from fauxgl_b200 import (HexColor, LoadOBJ, LookAt, NewPhongShader, V, Radians)
    from fauxgl_b200.context import DeviceMesh, FauxglError
    from fauxgl_b200.shader import NewImageTexture
    path = tmp_path / "box.obj"
    path.write_text(OBJ_TEXT)
    host = LoadOBJ(str(path))
    assert host.num_triangles == 2 + 2 + 2 + 2 + 3 + 1 + 2
    ctx = Context(320, 240)
    dm = DeviceMesh.FromOBJ(ctx, str(path))
    dpos, dnrm, _, _ = dm.read()
    assert same_bits(dpos, host.position) and same_bits(dnrm, host.normal)
    # the texture coordinates: compare through a textured Phong render of both copies
    rng = np.random.default_rng(5)
    texels = rng.integers(0, 256, size=(16, 16, 4), dtype=np.uint8)
    texels[..., 3] = 255
    eye = V(3, 2.5, 4)
    matrix = LookAt(eye, V(0, 0.3, 0), V(0, 1, 0)).Perspective(40, 320 / 240, 1, 20)
    images = []
    for mesh in (dm, DeviceMesh(ctx, host, ("position", "normal", "texture"))):
        sh = NewPhongShader(matrix, V(-0.75, 1, 0.25).Normalize(), eye)
        sh.Texture = NewImageTexture(texels)
        ctx.Shader = sh
        ctx.ClearDepthBuffer()
        ctx.ClearColorBufferWith(HexColor("#102030"))
        info = ctx.DrawMesh(mesh)
        images.append((ctx.Image().copy(), tuple(info)))
    assert images[0][1] == images[1][1] and images[0][1][0] > 0
    assert (images[0][0] == images[1][0]).all()
    # an index outside its table is refused (the reference panics)
    bad = tmp_path / "bad.obj"
    bad.write_text("v 0 0 0\nv 1 0 0\nv 0 1 0\nf 1 2 3\n")
    from fauxgl_b200 import mesh as fmesh
    vs, vts, vns, corners = fmesh.ParseOBJ(str(bad))
    corners = corners.copy(); corners[0, 0, 0] = 9
    from fauxgl_b200.context import IndexedDesc, capi, _P
    import ctypes as C
    d = IndexedDesc()
    d.v, d.vt, d.vn = (C.cast(a.ctypes.data, _P) for a in (vs, vts, vns))
    d.nv, d.nvt, d.nvn, d.corners, d.ntriangles = len(vs), len(vts), len(vns), C.cast(corners.ctypes.data, _P), 1
    h = _P()
    assert capi().fgl_mesh_create_indexed(ctx._h, C.byref(d), C.byref(h)) == -1
    ctx.Close()
