"""CPU tests of the host side: the Python mirror of the reference's value types,
mesh preparation, loaders, and the C-ABI library's exported surface (no compute
calls -- there is no GPU here and the library has no CPU fallback)."""
import math
import os
import re

import numpy as np
import pytest

import fauxgl_b200 as fgl
from fauxgl_b200 import (HexColor, Identity, LookAt, Mesh, NewTriangleMesh, Orthographic, Perspective, Radians, Rotate,
                         Scale, Translate, V, synth)

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = "/root/reference/examples"


def test_matrix_builders_follow_the_reference_formulas():
    # matrix.go:89-93,69-79
    m = Perspective(30, 16 / 9, 1, 10)
    ymax = 1 * math.tan(30 * math.pi / 360)
    xmax = ymax * (16 / 9)
    assert m[0] == 2 / (xmax - -xmax) and m[5] == 2 / (ymax - -ymax)
    assert m[10] == (-10 - 1) / 9 and m[11] == (-2 * 10) / 9 and m[14] == -1 and m[15] == 0
    # method forms left-multiply (matrix.go:167-169): LookAt(..).Perspective(..) == Perspective . LookAt
    eye, c, up = V(-3, 1, -0.75), V(0, -0.07, 0), V(0, 1, 0)
    a = LookAt(eye, c, up).Perspective(30, 16 / 9, 1, 10)
    b = Perspective(30, 16 / 9, 1, 10).Mul(LookAt(eye, c, up))
    assert tuple(a) == tuple(b)
    # LookAt maps the eye to the origin and looks down -z
    L = LookAt(eye, c, up)
    p = L.MulPosition(eye)
    assert max(abs(x) for x in p) < 1e-15
    assert L.MulPosition(c).Z < 0
    # Rotate is orthonormal
    R = Rotate(V(1, 2, 3), 0.7)
    RtR = np.array(R).reshape(4, 4)[:3, :3]
    assert np.allclose(RtR @ RtR.T, np.eye(3), atol=1e-15)
    assert tuple(Identity().Translate(V(1, 2, 3)).Scale(V(2, 2, 2))) == tuple(Scale(V(2, 2, 2)).Mul(Translate(V(1, 2, 3))))
    o = Orthographic(-1, 1, -1, 1, -1, 1)
    assert o.MulPositionW(V(0.25, -0.5, 0.3)) == (0.25, -0.5, -0.3, 1.0)


def test_hexcolor_and_nrgba_truncation():
    c = HexColor("#468966")
    assert c == (0x46 / 255, 0x89 / 255, 0x66 / 255, 1.0)
    assert HexColor("FFF") == (1.0, 1.0, 1.0, 1.0) and HexColor("#0000007f").A == 127 / 255
    assert fgl.Color(0.5, 1.2, -1, 0.999).NRGBA() == (127, 255, 0, 254)  # color.go:56-63: clamp, truncate


def test_biunitcube_fits_and_centres():
    rng = np.random.RandomState(5)
    m = NewTriangleMesh(rng.rand(50, 3, 3) * np.array([3.0, 1.0, 0.5]) + 7)
    m.BiUnitCube()
    p = m.position.reshape(-1, 3)
    size = p.max(0) - p.min(0)
    assert abs(size.max() - 2.0) < 1e-12                   # largest extent fills [-1,1]
    assert np.allclose((p.max(0) + p.min(0)) / 2, 0, atol=1e-12)  # anchored at the centre (mesh.go:127-151)
    assert np.allclose(np.linalg.norm(m.normal.reshape(-1, 3), axis=1), 1)


def test_smooth_normals_threshold_matches_naive_map_implementation():
    """mesh.go:80-103 restated naively (dict of lists, in corner order) vs the vectorised version."""
    mesh = synth.NewLatLngSphere(30, 30)
    naive = mesh.Copy()
    threshold = math.cos(Radians(40))
    lookup = {}
    P = naive.position.reshape(-1, 3) + 0.0
    N = naive.normal.reshape(-1, 3).copy()
    for i in range(len(P)):
        lookup.setdefault(tuple(P[i]), []).append(N[i])
    out = np.empty_like(N)
    for i in range(len(P)):
        acc = np.zeros(3)
        for x in lookup[tuple(P[i])]:
            if x[0] * N[i][0] + x[1] * N[i][1] + x[2] * N[i][2] >= threshold:
                acc = acc + x
        r = 1 / math.sqrt(acc[0] * acc[0] + acc[1] * acc[1] + acc[2] * acc[2])
        out[i] = acc * r
    mesh.SmoothNormalsThreshold(Radians(40))
    assert (mesh.normal.reshape(-1, 3).view(np.uint64) == out.view(np.uint64)).all()


def test_smooth_normals_groups_shared_positions():
    mesh = synth.NewCube()
    mesh.SmoothNormals()
    n = mesh.normal.reshape(-1, 3)
    p = mesh.position.reshape(-1, 3)
    # every corner of a cube ends up with the normalised diagonal direction
    assert np.allclose(n, p / np.linalg.norm(p, axis=1, keepdims=True), atol=1e-12)


def test_mesh_transform_matches_scalar_formulas():
    rng = np.random.RandomState(1)
    mesh = NewTriangleMesh(rng.rand(7, 3, 3))
    m = Rotate(V(0, 0, 1), Radians(5)).Translate(V(1, 2, 3))
    want_p = [m.MulPosition(V(*p)) for p in mesh.position.reshape(-1, 3)]
    want_n = [m.MulDirection(V(*n)) for n in mesh.normal.reshape(-1, 3)]
    mesh.Transform(m)
    assert (mesh.position.reshape(-1, 3) == np.array(want_p)).all()
    assert (mesh.normal.reshape(-1, 3) == np.array(want_n)).all()


def test_synthetic_benchmark_mesh_is_the_specified_one():
    m = synth.bumpy_surface(triangles=2000, nu=41, nv=41, smooth=True)
    assert m.num_triangles == 2000
    full = synth._grid_surface(lambda u, v: np.ones_like(u), 41, 41)
    assert len(full) == 41 * (2 * 41 - 2)
    # closed and outward facing: signed volume positive, every face normal points away from the centre
    e1, e2 = full[:, 1] - full[:, 0], full[:, 2] - full[:, 0]
    nrm = np.cross(e1, e2)
    assert (np.einsum("ij,ij->i", nrm, full.mean(axis=1)) > 0).all()
    assert synth.M871K_TRIANGLES == 871306 and 661 * (2 * 661 - 2) == 872520


@pytest.mark.skipif(not os.path.isdir(REF), reason="reference assets only exist in the build container")
def test_loaders_reproduce_the_committed_fixtures():
    import scenes
    for name, loader, fn in (("hello_mesh", fgl.LoadSTL, "hello.stl"), ("bowser_mesh", fgl.LoadSTL, "bowser.stl"),
                             ("capsule_mesh", fgl.LoadOBJ, "capsule.obj"), ("cube_mesh", fgl.LoadSTL, "cube.stl")):
        got = loader(os.path.join(REF, fn))
        want = scenes.load_fixture(name)
        assert (got.position == want.position).all() and np.array_equal(got.normal, want.normal, equal_nan=True)
    assert scenes.load_fixture("hello_mesh").num_triangles == 1440
    assert scenes.load_fixture("bowser_mesh").num_triangles == 11594
    assert scenes.load_fixture("capsule_mesh").num_triangles == 10200


def test_stl_roundtrip_ascii_and_binary(tmp_path):
    tri = np.array([[[0, 0, 0], [1, 0, 0], [0, 1, 0]], [[0, 0, 1], [1, 0, 1], [0, 1, 1.5]]], dtype=np.float64)
    a = tmp_path / "a.stl"
    a.write_text("solid x\n" + "".join(
        "facet normal 0 0 0\n outer loop\n" + "".join("  vertex %r %r %r\n" % tuple(float(c) for c in v) for v in t) + " endloop\nendfacet\n"
        for t in tri) + "endsolid x\n")
    m = fgl.LoadSTL(str(a))
    assert (m.position == tri).all() and np.allclose(m.normal[0], [0, 0, 1])
    import struct
    b = tmp_path / "b.stl"
    with open(b, "wb") as f:
        f.write(b"\0" * 80 + struct.pack("<I", 2))
        for t in tri:
            f.write(struct.pack("<12fH", 0, 0, 0, *t.reshape(-1), 0))
    mb = fgl.LoadSTL(str(b))
    assert (mb.position == tri.astype(np.float32).astype(np.float64)).all()


def test_obj_loader_indices_and_fan(tmp_path):
    o = tmp_path / "q.obj"
    o.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 1 1\nvt 0 1\nvn 0 0 1\n"
                 "f 1/1/1 2/2/1 3/3/1 4/4/1\nf -4 -3 -2\n")
    m = fgl.LoadOBJ(str(o))
    assert m.num_triangles == 3                      # quad fan (0,1,2),(0,2,3) + one triangle
    assert (m.position[1] == [[0, 0, 0], [1, 1, 0], [0, 1, 0]]).all()
    assert (m.texture[0, 1] == [1, 0, 0]).all()
    assert (m.normal[2] == [[0, 0, 1]] * 3).all()      # missing vn index -> zero -> FixNormals -> face normal


# ---- the C ABI --------------------------------------------------------------------------------------

def _header_functions():
    text = open(os.path.join(ROOT, "include", "fauxgl_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(fgl_[a-z0-9_]+)\s*\(", text)))


def test_library_exports_every_declared_symbol():
    from fauxgl_b200 import build as fbuild
    from fauxgl_b200 import context
    if not os.path.exists(context.LIB_PATH):
        fbuild.build_library()
    lib = context.capi()
    declared = _header_functions()
    assert len(declared) >= 30
    bound = {name for name, _r, _a in context.ABI}
    for name in declared:
        assert hasattr(lib, name), "libfauxgl_b200.so does not export " + name
        assert name in bound, "python binding misses " + name
    assert lib.fgl_abi_version() == 2


def test_no_cpu_fallback_without_a_device():
    """Without a CUDA device context creation must fail loudly (FGL_E_NO_DEVICE), never render."""
    from fauxgl_b200 import context
    lib = context.capi()
    if lib.fgl_device_count() > 0:
        pytest.skip("a CUDA device is present")
    with pytest.raises(context.FauxglError) as e:
        context.Context(16, 16)
    assert e.value.status == -3 and "no CPU fallback" in str(e.value)


def test_unsupported_shader_is_an_error_not_a_fallback():
    from fauxgl_b200 import context

    class MyShader:  # a user-defined Shader with no device equivalent
        def describe(self):
            return {"kind": 99, "matrix": tuple(Identity())}
    ctx = context.Context.__new__(context.Context)
    ctx.Shader = MyShader()
    ctx._textures, ctx._keep = {}, None
    with pytest.raises(context.FauxglError) as e:
        ctx._shader()
    assert e.value.status == -4


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: nothing under fauxgl_b200/ or include/ may reference it."""
    for base in ("fauxgl_b200", "include", "go"):
        for dirpath, _d, files in os.walk(os.path.join(ROOT, base)):
            for fn in files:
                if fn.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".go")):
                    src = open(os.path.join(dirpath, fn), errors="replace").read()
                    assert "pyoracle" not in src and "fauxgl_oracle" not in src and "import oracle" not in src, \
                        os.path.join(dirpath, fn)


def test_parse_obj_tables_and_fan_triangulation(tmp_path):
    """mesh.ParseOBJ / LoadOBJ (obj.go:19-79): 1-based tables with a zero entry 0, negative indices counted from
    the end, missing vt / vn -> entry 0, polygons as fans from their first vertex, FixNormals on zero normals."""
    from fauxgl_b200 import LoadOBJ
    from fauxgl_b200 import mesh as fmesh
    path = tmp_path / "t.obj"
    path.write_text("v 0 0 0\nv 1 0 0\nv 1 1 0\nv 0 1 0\nv 0.5 2 0\nvt 0.25 0.75\nvn 0 0 1\n"
                    "f 1/1/1 2/1/1 3/1/1 4/1/1\n"      # quad -> 2 triangles
                    "f -5 -4 -3 -2 -1\n"                # pentagon, negative indices, no vt/vn -> 3 triangles
                    "# comment\n\nf 1//1 2//1 5\n")
    vs, vts, vns, corners = fmesh.ParseOBJ(str(path))
    assert vs.shape == (6, 3) and vts.shape == (2, 3) and vns.shape == (2, 3) and (vs[0] == 0).all()
    assert corners.shape == (6, 3, 3)
    assert corners[:2, :, 0].tolist() == [[1, 2, 3], [1, 3, 4]]                     # fan from the first vertex
    assert corners[2:5, :, 0].tolist() == [[1, 2, 3], [1, 3, 4], [1, 4, 5]]         # -5..-1 of 6 entries -> 1..5
    assert (corners[2:5, :, 1:] == 0).all()                                         # missing vt / vn -> entry 0
    assert corners[5].tolist() == [[1, 0, 1], [2, 0, 1], [5, 0, 0]]
    m = LoadOBJ(str(path))
    assert m.num_triangles == 6
    assert (m.texture[0, :, :2] == [0.25, 0.75]).all() and (m.texture[2] == 0).all()
    assert (m.normal[0] == [0, 0, 1]).all()
    # zero normals are replaced by the face normal (triangle.go:46-58); the pentagon lies in z = 0 and its fan
    # triangles wind CCW, CCW, CW
    assert (m.normal[2:4] == [0, 0, 1]).all() and (m.normal[4, :, 2] == -1).all() and (m.normal[4, :, :2] == 0).all()
    assert (m.normal[5, 2] == m.normal[5, 0]).all()                                 # corner 3 had no vn: face normal (0,0,1)


def test_view_volume_reject_agrees_with_sutherland_hodgman():
    """The front ends drop a triangle with a vertex outside the view volume when the rule of fgl_geom.cu
    clips_to_nothing says the clipper would return nothing: planes in sutherlandHodgman's order, as long as all three
    vertices are in front of a plane the polygon passes it unchanged, the first plane with all three behind it
    empties it, a plane that cuts the triangle ends the test.  Restated here next to clipping.go:16-52 (pure Python,
    the same float64 operations in the same order): whenever the rule fires the clipper's output is empty -- over random
    clip-space triangles that hug and touch the planes (the `> 0` test is strict), and it fires often enough to matter."""
    import random
    PLANES = [((1, 0, 0, 1), (-1, 0, 0, 1)), ((-1, 0, 0, 1), (1, 0, 0, 1)), ((0, 1, 0, 1), (0, -1, 0, 1)),
              ((0, -1, 0, 1), (0, 1, 0, 1)), ((0, 0, 1, 1), (0, 0, -1, 1)), ((0, 0, -1, 1), (0, 0, 1, 1))]

    def sub(a, b): return tuple(x - y for x, y in zip(a, b))
    def add(a, b): return tuple(x + y for x, y in zip(a, b))
    def dot(a, b): return a[0] * b[0] + a[1] * b[1] + a[2] * b[2] + a[3] * b[3]
    def in_front(pl, v): return dot(sub(v, pl[0]), pl[1]) > 0

    def intersect(pl, v0, v1):
        u, w = sub(v1, v0), sub(v0, pl[0])
        d, n = dot(pl[1], u), -dot(pl[1], w)
        t = n / d if d != 0 else float("nan") if n == 0 else float("inf") * (1 if n > 0 else -1)
        return add(v0, tuple(x * t for x in u))

    def clipper(points):
        out = list(points)
        for pl in PLANES:
            inp, out = out, []
            if not inp:
                return []
            s = inp[-1]
            for e in inp:
                if in_front(pl, e):
                    if not in_front(pl, s):
                        out.append(intersect(pl, s, e))
                    out.append(e)
                elif in_front(pl, s):
                    out.append(intersect(pl, s, e))
                s = e
        return out

    def rule(o):
        for pl in PLANES:
            f = [in_front(pl, v) for v in o]
            if not any(f):
                return True
            if not all(f):
                return False
        return False

    rnd = random.Random(7)
    fired = 0
    for it in range(60000):
        w = [rnd.choice((1.0, 2.5, rnd.uniform(0.5, 9.0))) for _ in range(3)]
        o = []
        for k in range(3):
            v = []
            for _ in range(3):
                mode = rnd.randrange(5)
                if mode == 0:
                    v.append(rnd.choice((-1.0, 1.0)) * w[k])                          # exactly on a plane
                elif mode == 1:
                    v.append(rnd.choice((-1.0, 1.0)) * w[k] * (1 + rnd.uniform(-1e-15, 1e-15)))  # an ulp or two off it
                elif mode == 2:
                    v.append(rnd.uniform(-1, 1) * w[k])                               # inside
                else:
                    v.append(rnd.uniform(1.0, 3.0) * w[k] * rnd.choice((-1.0, 1.0)))  # outside
            o.append((v[0], v[1], v[2], w[k]))
        if rule(o):
            fired += 1
            assert clipper(o) == [], (it, o)
    assert fired > 3000, fired


def test_select_form_of_go_max_min_equals_the_if_chain():
    """fgl_math.cuh writes Go's math.Max / math.Min as nested selects (no branch on k_front's dependent chain).  The
    select form and the if-chain that follows the Go source are restated here and compared bit for bit over every
    pair of special values (signed zeros, infinities, NaN, subnormals) and random values."""
    import math
    import random
    import struct
    INF, NAN = float("inf"), struct.unpack("<d", struct.pack("<Q", 0x7ff8000000000001))[0]

    def bits(x): return struct.unpack("<Q", struct.pack("<d", x))[0]
    def sign(x): return math.copysign(1.0, x) < 0

    def max_chain(x, y):   # math.Max, with the special cases in the order the device code had them
        if x > y: return x
        if y > x: return y
        if x == y: return y if (x == 0 and sign(x)) else x
        if x == INF or y == INF: return INF
        return NAN

    def max_select(x, y):
        r_eq = y if sign(x) else x
        r_nan = INF if (x == INF or y == INF) else NAN
        return x if x > y else (y if y > x else (r_eq if x == y else r_nan))

    def min_chain(x, y):
        if x < y: return x
        if y < x: return y
        if x == y: return y if (x == 0 and not sign(x)) else x
        if x == -INF or y == -INF: return -INF
        return NAN

    def min_select(x, y):
        r_eq = x if sign(x) else y
        r_nan = -INF if (x == -INF or y == -INF) else NAN
        return x if x < y else (y if y < x else (r_eq if x == y else r_nan))

    rnd = random.Random(3)
    vals = [0.0, -0.0, 1.0, -1.0, INF, -INF, NAN, 5e-324, -5e-324, 1.5, -2.25, 1e308, -1e308]
    vals += [rnd.uniform(-10, 10) for _ in range(40)]
    for x in vals:
        for y in vals:
            a, b = max_chain(x, y), max_select(x, y)
            assert bits(a) == bits(b) or (a != a and b != b), (x, y, a, b)
            a, b = min_chain(x, y), min_select(x, y)
            assert bits(a) == bits(b) or (a != a and b != b), (x, y, a, b)
