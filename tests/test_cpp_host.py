"""The C++ host mirror (include/fauxgl.hpp) over the C ABI: it must compile against
the library on any box (CPU test), and on the GPU box the example program's output
must equal the CPU oracle's render of the same scene bit for bit."""
import os
import struct
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIBDIR = os.path.join(ROOT, "fauxgl_b200")


def _build(tmp_path):
    from fauxgl_b200 import build as fbuild
    from fauxgl_b200 import context
    if not os.path.exists(context.LIB_PATH):
        fbuild.build_library()
    exe = str(tmp_path / "cube")
    cmd = ["g++", "-std=c++17", "-O2", "-ffp-contract=off", "-Wall", "-I" + os.path.join(ROOT, "include"),
           os.path.join(ROOT, "examples", "cube.cpp"), "-L" + LIBDIR, "-lfauxgl_b200", "-Wl,-rpath," + LIBDIR, "-o", exe]
    subprocess.check_call(cmd)
    return exe


def test_cpp_mirror_compiles_and_fails_loudly_without_gpu(tmp_path):
    exe = _build(tmp_path)
    from fauxgl_b200 import context
    if context.capi().fgl_device_count() > 0:
        pytest.skip("a CUDA device is present")
    p = subprocess.run([exe, str(tmp_path / "out.raw")], capture_output=True, text=True)
    assert p.returncode == 1 and "no CPU fallback" in p.stderr      # FGL_E_NO_DEVICE, nothing rendered
    assert not (tmp_path / "out.raw").exists()


@pytest.mark.gpu
def test_cpp_example_matches_oracle(tmp_path, oracle_lib, gpu_capi):
    from fauxgl_b200 import Black, HexColor, Matrix, NewLineMesh, NewPhongShader, NewSolidColorShader, V, synth
    exe = _build(tmp_path)
    out = tmp_path / "cube.raw"
    subprocess.check_call([exe, str(out)])
    raw = out.read_bytes()
    w, h = struct.unpack_from("<ii", raw, 0)
    matrix = Matrix(struct.unpack_from("<16d", raw, 8))
    light = struct.unpack_from("<3d", raw, 8 + 128)
    infos = struct.unpack_from("<4Q", raw, 8 + 128 + 24)
    off = 8 + 128 + 24 + 32
    color = np.frombuffer(raw, np.uint8, w * h * 4, off).reshape(h, w, 4)
    depth = np.frombuffer(raw, np.float64, w * h, off + w * h * 4).reshape(h, w)
    # the same scene on the oracle, with the matrix the C++ host computed
    mesh = synth.NewCube()
    v = np.array([[-1, -1, -1], [-1, -1, 1], [-1, 1, -1], [-1, 1, 1], [1, -1, -1], [1, -1, 1], [1, 1, -1], [1, 1, 1]], float) * 0.5
    edges = [(0, 1), (2, 3), (4, 5), (6, 7), (0, 2), (1, 3), (4, 6), (5, 7), (0, 4), (2, 6), (1, 5), (3, 7)]
    lines = NewLineMesh(np.array([[v[a], v[b]] for a, b in edges]))
    o = oracle_lib.OracleContext(w, h)
    o.ClearColorBufferWith(HexColor("#24221F"))
    sh = NewPhongShader(matrix, V(*light), V(2, 1.5, 1.2))
    sh.ObjectColor = HexColor("#468966")
    o.Shader = sh
    i1 = o.DrawTriangles(mesh)
    o.Shader = NewSolidColorShader(matrix, Black)
    o.LineWidth, o.DepthBias = 3, -1e-4
    i2 = o.DrawLines(lines)
    assert infos == (i1[0], i1[1], i2[0], i2[1])
    assert (depth.view(np.uint64) == o.DepthBuffer.view(np.uint64)).all()
    assert (color == o.ColorBuffer).all()
