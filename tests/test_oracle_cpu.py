"""CPU tests of the oracle (the C restatement of the reference's DrawMesh path).

The reference ships no tests and no golden images, so these pins are ours:
  * tests/golden/oracle_golden.json -- sha256 of colour/depth buffers and
    RasterizeInfo of every parity scene, produced by tests/golden/make_fixtures.py;
  * the coverage/depth probe of examples/hello.go from SURVEY.md Appendix C,
    computed there by an independent numpy restatement (TotalPixels 216 654,
    UpdatedPixels 211 242, 210 310 covered pixels, depth in [0.62673, 0.72183]);
  * a second, independent pure-Python restatement of the per-triangle path for
    small cases (tests/py_reference.py).
"""
import hashlib
import json
import os

import numpy as np
import pytest

import scenes
from oracle import pyoracle

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "oracle_golden.json")


@pytest.fixture(scope="module")
def golden():
    with open(GOLDEN) as f:
        return json.load(f)


@pytest.mark.parametrize("name", scenes.GOLDEN_SCENES)
def test_oracle_matches_golden(name, golden, oracle_lib):
    sc = scenes.SCENES[name]()
    ctx = pyoracle.OracleContext(sc.width, sc.height)   # the reference's own index rule
    infos = sc.run(ctx)
    g = golden[name]
    assert [list(map(int, i)) for i in infos] == g["info"]
    assert hashlib.sha256(ctx.ColorBuffer.tobytes()).hexdigest() == g["color_sha256"]
    assert hashlib.sha256(ctx.DepthBuffer.tobytes()).hexdigest() == g["depth_sha256"]
    assert int((ctx.DepthBuffer < 1e300).sum()) == g["covered"]


def test_hello_matches_survey_probe(oracle_lib):
    """SURVEY.md Appendix C (independent numpy probe of examples/hello.go)."""
    sc = scenes.hello()
    ctx = pyoracle.OracleContext(sc.width, sc.height, x_guard=False)  # faithful mode
    (info,) = sc.run(ctx)
    assert info == (216654, 211242)
    d = ctx.DepthBuffer
    covered = d < 1e300
    assert int(covered.sum()) == 210310
    assert abs(d[covered].min() - 0.62673) < 1e-5 and abs(d[covered].max() - 0.72183) < 1e-5


def test_threaded_schedule_matches_sequential_when_order_free(oracle_lib):
    """The reference's goroutine schedule (i % wn striding + 256 mutexes) gives the
    sequential result whenever the state is order-independent (opaque, no bias, no ties)."""
    sc = scenes.bumpy_small()
    a = pyoracle.OracleContext(sc.width, sc.height, threads=1)
    b = pyoracle.OracleContext(sc.width, sc.height, threads=4)
    ia, ib = sc.run(a), sc.run(b)
    assert ia[0][0] == ib[0][0]                      # TotalPixels is schedule independent
    assert (a.DepthBuffer == b.DepthBuffer).all()
    assert (a.ColorBuffer == b.ColorBuffer).all()


def test_x_guard_divergence_is_confined_to_offscreen_wrap(oracle_lib):
    """The reference never range-checks x (context.go:223-228), so a fat line leaving the screen
    aliases into the neighbouring row; oracle and device reproduce that by default.  The optional
    x-guard rule (fgl_state.x_guard) drops those fragments: bound what it changes."""
    sc = scenes.lines_scene()
    faithful = pyoracle.OracleContext(sc.width, sc.height, x_guard=False)
    guarded = pyoracle.OracleContext(sc.width, sc.height, x_guard=True)
    fi, gi = sc.run(faithful), sc.run(guarded)
    diff = (faithful.DepthBuffer.view(np.uint64) != guarded.DepthBuffer.view(np.uint64)) | \
           (faithful.ColorBuffer != guarded.ColorBuffer).any(axis=-1)
    ys, xs = np.nonzero(diff)
    print("x-guard divergent pixels:", len(xs), "of", diff.size, "TotalPixels", fi, gi)
    # wrapped fragments land within LineWidth of the left/right border only
    assert all(x < 8 or x >= sc.width - 8 for x in xs)
    # this scene is built to provoke the quirk (lines leaving the screen on purpose): 49 of 450 000 pixels
    assert 0 < len(xs) < 100
    assert all(f[0] >= g[0] for f, g in zip(fi, gi))


def test_pow_matches_known_values(oracle_lib):
    """Go math.Pow restatement: exact for integer exponents on dyadic bases, special cases."""
    L = pyoracle.lib()
    assert L.oracle_pow(0.5, 32.0) == 2.0 ** -32
    assert L.oracle_pow(0.75, 2.0) == 0.5625
    assert L.oracle_pow(2.0, -3.0) == 0.125
    assert L.oracle_pow(9.0, 0.5) == 3.0
    assert L.oracle_pow(0.0, 0.0) == 1.0 and L.oracle_pow(123.0, 0.0) == 1.0
    assert L.oracle_pow(0.0, 5.0) == 0.0 and L.oracle_pow(0.0, -1.0) == float("inf")
    assert np.isnan(L.oracle_pow(-2.0, 0.5))
    x = 0.987654321
    # Go's Pow multiplies mantissas by repeated squaring: a few ulps off the correctly rounded
    # value, and exactly what the independent Python restatement computes
    assert abs(L.oracle_pow(x, 100.0) - x ** 100) <= 16 * np.spacing(x ** 100)
    import py_reference
    rng = np.random.RandomState(2)
    for base in rng.rand(200):
        for power in (2.0, 32.0, 64.0, 100.0, 7.0):
            assert L.oracle_pow(float(base), power) == py_reference.go_pow(float(base), power)


def test_resolve_box_properties(oracle_lib):
    """nfnt-bilinear restatement: a constant image stays constant, weights sum to 1024, and an
    opaque input stays opaque; factor 1 is the identity (resize returns its input)."""
    img = np.empty((64, 96, 4), np.uint8)
    img[...] = (10, 200, 33, 255)
    out = pyoracle.resolve(img, 24, 16)
    assert (out == np.array([10, 200, 33, 255], np.uint8)).all()
    assert (pyoracle.resolve(img, 96, 64) == img).all()
    rng = np.random.RandomState(0)
    img = rng.randint(0, 256, (32, 32, 4)).astype(np.uint8)
    out = pyoracle.resolve(img, 8, 8)
    # transparent texels are premultiplied away: alpha channel is the filtered alpha
    assert out.shape == (8, 8, 4)
    img[..., 3] = 255
    out = pyoracle.resolve(img, 8, 8)
    assert (out[..., 3] == 255).all()
    # 4x: the 8 taps are 32,96,160,224,224,160,96,32 over source columns 4x-2 .. 4x+5
    row = np.zeros((1, 16, 4), np.uint8)
    row[..., 3] = 255
    row[0, 6, 0] = 255  # inside the footprint of output column 1 (cols 2..9) with weight 224 (tap 4)
    img = np.repeat(row, 4, axis=0)
    out = pyoracle.resolve(img, 4, 1)
    assert out[0, 1, 0] == (224 * 255) // 1024


def test_pack_key_is_monotone_in_depth(oracle_lib):
    zs = [0.0, 1e-12, 0.1, 0.5, 0.5 + 1e-9, 0.999999, 1.0, 1.7976931348623157e308]
    keys = [np.int64(np.uint64(pyoracle.pack_key(z, (1, 2, 3, 4)))) for z in zs]
    assert all(a <= b for a, b in zip(keys, keys[1:])), keys
    assert keys[0] < keys[3] < keys[-1]
    # colour breaks ties, most significant byte first
    a = np.int64(np.uint64(pyoracle.pack_key(0.25, (1, 0, 0, 255))))
    b = np.int64(np.uint64(pyoracle.pack_key(0.25, (2, 0, 0, 0))))
    assert a < b


def test_c_oracle_matches_python_restatement(oracle_lib):
    """Two independent restatements of the reference (C and pure Python, both written from the
    Go source) must agree bit for bit: depth, colour and RasterizeInfo."""
    import py_reference
    from fauxgl_b200 import LookAt, NewPhongShader, NewTriangleMesh, V, synth
    rng = np.random.RandomState(11)
    cube = synth.NewCube()
    extra = rng.rand(24, 3, 3) * 1.6 - 0.8
    mesh = NewTriangleMesh(np.concatenate([cube.position, extra], axis=0))
    mesh.color[:, :, :3] = rng.rand(mesh.num_triangles, 3, 3)
    mesh.color[:, :, 3] = np.where(rng.rand(mesh.num_triangles, 3) < 0.5, 1.0, 0.6)
    W, H = 96, 64
    eye = V(1.5, 2.0, 1.2)
    matrix = LookAt(eye, V(0, 0, 0), V(0, 0, 1)).Perspective(45, W / H, 0.5, 10)
    shader = NewPhongShader(matrix, V(0.3, 0.5, 1).Normalize(), eye)   # vertex colours (ObjectColor == Discard)
    for cull_back, bias in ((True, 0.0), (False, -1e-4)):
        octx = pyoracle.OracleContext(W, H)
        octx.Shader = shader
        octx.Cull = 3 if cull_back else 1
        octx.DepthBias = bias
        oinfo = octx.DrawMesh(mesh)
        pctx = py_reference.PyContext(W, H)
        pctx.cull_back, pctx.depth_bias = cull_back, bias
        d = shader.describe()
        tot = upd = 0
        for t in range(mesh.num_triangles):
            a, b = pctx.draw_triangle(d, mesh.position[t].tolist(), mesh.normal[t].tolist(), mesh.color[t].tolist())
            tot, upd = tot + a, upd + b
        assert (tot, upd) == oinfo
        assert tot > 500
        pd = np.array(pctx.depth).reshape(H, W)
        pc = np.array(pctx.color, dtype=np.uint8).reshape(H, W, 4)
        assert (pd.view(np.uint64) == octx.DepthBuffer.view(np.uint64)).all()
        assert (pc == octx.ColorBuffer).all()


def test_depth_image_matches_numpy_restatement(oracle_lib):
    """oracle_depth_image (context.go:87-117) against a direct numpy statement of the same formula."""
    sc = scenes.hello()
    o = oracle_lib.OracleContext(sc.width, sc.height)
    assert (o.DepthImage() == 0xffff).all()          # nothing drawn: every pixel cleared -> t = 1
    sc.run(o)
    d = o.DepthBuffer
    mx = np.finfo(np.float64).max
    drawn = d != mx
    lo, hi = d[drawn].min(), d[drawn].max()
    with np.errstate(over="ignore"):
        t = np.where(drawn, (d - lo) / (hi - lo), 1.0)
    want = np.trunc(t * 65535.0).astype(np.uint16)
    got = o.DepthImage()
    assert got.dtype == np.uint16 and (got == want).all()
    assert got[drawn].min() == 0 and got[drawn].max() == 0xffff and (got[~drawn] == 0xffff).all()
    # one depth value only: (d - lo) / 0 = NaN -> uint16(NaN) is 0 on amd64
    flat = np.full((4, 5), mx)
    flat[1, 2] = 0.5
    g = oracle_lib.depth_image(flat)
    assert g[1, 2] == 0 and (np.delete(g.ravel(), 7) == 0xffff).all()


def test_stl_records_match_python_loader(oracle_lib, tmp_path):
    """oracle_stl_triangles (stl.go:86-154) == the Python mirror's LoadSTL on the same bytes."""
    import struct
    from fauxgl_b200 import mesh as fmesh
    rng = np.random.default_rng(11)
    pos = rng.uniform(-50, 50, size=(300, 3, 3)).astype("<f4")
    pos[17] = pos[17][1]    # degenerate: NaN normal, as in the reference
    rec = np.zeros((300, 50), dtype=np.uint8)
    rec[:, 12:48] = pos.reshape(300, 9).view(np.uint8).reshape(300, 36)
    data = b"x" * 80 + struct.pack("<I", 300) + rec.tobytes()
    path = tmp_path / "t.stl"
    path.write_bytes(data)
    m = fmesh.LoadSTL(str(path))
    p, n = oracle_lib.stl_triangles(data[84:])
    assert (p == pos.astype(np.float64)).all()
    assert (m.position.view(np.uint64) == p.view(np.uint64)).all()
    assert (np.ascontiguousarray(m.normal).view(np.uint64) == n.view(np.uint64)).all()
    assert np.isnan(n[17]).all() and np.allclose(np.linalg.norm(np.delete(n, 17, axis=0)[:, 0], axis=1), 1.0)


def test_draw_each_equals_one_primitive_draws(oracle_lib):
    """oracle_draw_each: per-primitive RasterizeInfo == DrawLines/DrawTriangles over one primitive at a time."""
    sc = scenes.lines_scene()
    a, b = oracle_lib.OracleContext(sc.width, sc.height), oracle_lib.OracleContext(sc.width, sc.height)
    captured = {}

    class Capture:
        def __init__(self, c):
            self.__dict__["c"] = c

        def __getattr__(self, k):
            return getattr(self.c, k)

        def __setattr__(self, k, v):
            setattr(self.c, k, v)

        def DrawLines(self, m, *args):
            captured.setdefault("lines", m)
            return (0, 0)
    sc.run(Capture(a)); sc.run(Capture(b))
    lines = captured["lines"]
    each = a.DrawLinesEach(lines)
    one = np.array([b.DrawLines(lines, i, 1) for i in range(lines.num_lines)], dtype=np.uint64)
    assert (each == one).all()
    assert (a.ColorBuffer == b.ColorBuffer).all()
    assert each[:, 0].sum() > 1000 and (each[:, 1] <= each[:, 0]).all()


def _tightened_box(s, W, H):
    """fgl_geom.cu tighten_box restated (same formulas, same constants): the rows [cy0, cy1] and the last column cx1
    the fused front end still walks for a screen triangle s[3][2], or None when it walks nothing / does not tighten."""
    import math
    mnx, mxx = float(s[:, 0].min()), float(s[:, 0].max())
    mny, mxy = float(s[:, 1].min()), float(s[:, 1].max())
    fx0, fx1, fy0, fy1 = math.floor(mnx), math.ceil(mxx), math.floor(mny), math.ceil(mxy)
    if not (0 <= fx0 < W and 0 <= fx1 < W):
        return "skip"   # boxes that leave the screen sideways are not tightened
    # edge_fn(s0, s1, s2), context.go:147-149,163
    area = (s[1, 0] - s[2, 0]) * (s[0, 1] - s[2, 1]) - (s[1, 1] - s[2, 1]) * (s[0, 0] - s[2, 0])
    with np.errstate(divide="ignore", invalid="ignore"):
        ra = np.float64(1.0) / np.float64(area)
    hx, hy = mxx - mnx, mxy - mny
    B = max(hx, hy) + 4.0
    E = 2.0 ** -44 * B * B * (B + 8.0)
    k = 4.0 * E * abs(float(ra))
    my, mx = k * hy + 1e-6, k * hx + 1e-6
    sure = k < 0.25 and my < 0.25 and mx < 0.25
    drop_first = 1 if (sure and mny - fy0 >= 0.5 + my) else 0
    drop_last = (2 if fy1 - mxy >= 0.5 + my else 1) if sure else 0
    drop_right = (2 if fx1 - mxx >= 0.5 + mx else 1) if sure else 0
    cy0, cy1 = max(max(fy0, 0), fy0 + drop_first), min(min(fy1, H - 1), fy1 - drop_last)
    cx1 = fx1 - drop_right
    walked_before = max(0, min(fy1, H - 1) - max(fy0, 0) + 1)
    if cy0 > cy1 or cx1 < fx0:
        return (None, walked_before)
    return ((cy0, cy1, cx1), walked_before)


def test_tightened_boxes_never_lose_a_fragment(oracle_lib):
    """The fused front end does not walk the rows and columns of a triangle's box that its tighten_box proves empty
    (DESIGN.md section 4).  The proof is checked here against the oracle, without a GPU: thousands of triangles --
    sub-pixel, a few pixels, large, slivers, vertices on pixel centres and pixel edges -- are rasterised one at a time
    by the reference's restated loop, and every pixel it covers must lie inside the tightened box."""
    from fauxgl_b200 import Identity, NewSolidColorShader, NewTriangleMesh, HexColor
    W, H = 64, 32   # powers of two: the NDC -> screen mapping below is exact for dyadic coordinates
    rng = np.random.RandomState(2024)
    octx = pyoracle.OracleContext(W, H)
    octx.Shader = NewSolidColorShader(Identity(), HexColor("#ffffff"))
    octx.Cull = 1          # CullNone
    octx.ReadDepth = False  # every covered pixel writes its depth
    CLEAR = np.finfo(np.float64).max
    kept = dropped = tested = vanished = covered = 0
    for it in range(4000):
        kind = it % 5
        c = np.array([rng.uniform(3, W - 3), rng.uniform(3, H - 3)])
        if kind == 0:   # sub-pixel to two pixels
            s = c + rng.uniform(-1.2, 1.2, (3, 2))
        elif kind == 1:  # a few pixels
            s = c + rng.uniform(-4, 4, (3, 2))
        elif kind == 2:  # large
            s = np.stack([rng.uniform(1, W - 2, 3), rng.uniform(1, H - 2, 3)], axis=1)
        elif kind == 3:  # sliver: two vertices a hair apart
            a = c + rng.uniform(-6, 6, 2)
            s = np.stack([a, a + rng.uniform(-1e-3, 1e-3, 2), c + rng.uniform(-6, 6, 2)])
        else:            # vertices on pixel centres, pixel edges and quarter positions
            s = np.round((c + rng.uniform(-3, 3, (3, 2))) * 4) / 4
        s = np.clip(s, 0.25, [W - 1.25, H - 1.25])
        # positions whose screen transform (matrix.go:119-128 via MulPosition) gives s; recompute s from them exactly
        ndc = np.stack([(s[:, 0] - W / 2) / (W / 2), (H / 2 - s[:, 1]) / (H / 2)], axis=1)
        sx = (W / 2) * ndc[:, 0] + W / 2
        sy = -(H / 2) * ndc[:, 1] + H / 2
        s = np.stack([sx, sy], axis=1)
        pos = np.zeros((1, 3, 3))
        pos[0, :, :2] = ndc
        pos[0, :, 2] = rng.uniform(-0.5, 0.5, 3)
        box = _tightened_box(s, W, H)
        if box == "skip":
            continue
        octx.ClearDepthBuffer()
        octx.DrawTriangles(NewTriangleMesh(pos, normal=np.ones((1, 3, 3))))
        ys, xs = np.nonzero(octx.DepthBuffer != CLEAR)
        tested += 1
        covered += len(ys)
        rows, before = box
        if rows is None:
            assert len(ys) == 0, (it, s, ys, xs)
            vanished += 1
            dropped += before
            continue
        cy0, cy1, cx1 = rows
        assert (ys >= cy0).all() and (ys <= cy1).all() and (xs <= cx1).all(), (it, s.tolist(), rows, ys.tolist(), xs.tolist())
        kept += cy1 - cy0 + 1
        dropped += before - (cy1 - cy0 + 1)
    assert tested > 3500 and vanished > 50 and covered > 20000, (tested, vanished, covered)
    assert dropped > 0.2 * (kept + dropped)   # the test is not vacuous: a good share of the rows is dropped
