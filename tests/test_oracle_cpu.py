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
    ctx = pyoracle.OracleContext(sc.width, sc.height, x_guard=True)
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
    """DESIGN.md 'x-guard': the reference never range-checks x (context.go:223-228), so a fat
    line leaving the screen wraps into the neighbouring row.  The adopted rule drops those
    fragments.  Report and bound the divergence on the scene that provokes it."""
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
    # this scene is built to provoke the quirk (lines leaving the screen on purpose): 49 of
    # 450 000 pixels; the divergent count is reported in DESIGN.md rather than hidden
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
        octx = pyoracle.OracleContext(W, H, x_guard=True)
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
