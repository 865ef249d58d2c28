"""Randomised parity: small random scenes -- triangles and lines partly outside the view volume (clipping on every
plane), random vertex colours with alpha, random render state (cull / front face / wireframe / line width / depth
bias / blend / read-write flags), every shader kind with 8-bit, premultiplying and 16-bit textures, ragged
framebuffer sizes, several draws on top of each other -- rendered by the oracle and by the device, compared bit for
bit (float64 depth, NRGBA8 colour, RasterizeInfo).  Seeds are fixed: a failure names its seed."""
import numpy as np
import pytest

from fauxgl_b200 import (Color, LookAt, Mesh, NewImageTexture, NewLineMesh, NewPhongShader, NewSolidColorShader,
                         NewTextureShader, NewTriangleMesh, Orthographic, V)
from fauxgl_b200.shader import TEX_NRGBA, TEX_RGBA, TEX_RGBA64

pytestmark = pytest.mark.gpu


def _random_texture(rng):
    h, w = int(rng.randint(2, 40)), int(rng.randint(2, 40))
    kind = rng.randint(3)
    if kind == 2:
        px = rng.randint(0, 65536, (h, w, 4)).astype(np.uint16)
        px[..., :3] = np.minimum(px[..., :3], px[..., 3:4])       # premultiplied, as Color.RGBA() returns
        return NewImageTexture(px, TEX_RGBA64)
    px = rng.randint(0, 256, (h, w, 4)).astype(np.uint8)
    if kind == 0:
        px[..., 3] = 255
    return NewImageTexture(px, TEX_RGBA if kind == 0 else TEX_NRGBA)


def _random_mesh(rng, ntri, nline):
    spread = rng.choice([0.6, 1.2, 2.5])                         # 2.5: most primitives cross the view volume
    centres = (rng.rand(ntri, 1, 3) * 2 - 1) * spread
    size = rng.choice([0.02, 0.2, 1.0])
    pos = centres + (rng.rand(ntri, 3, 3) * 2 - 1) * size
    if ntri:
        pick = rng.rand(ntri)
        pos[pick < 0.05, 1] = pos[pick < 0.05, 0]                      # degenerate: two equal vertices
        snap = (pick > 0.05) & (pick < 0.15)
        pos[snap] = np.round(pos[snap] * 8) / 8                       # vertices on a coarse lattice: shared edges, pixel-centre hits, ties
        far = (pick > 0.15) & (pick < 0.18)
        pos[far, 0] *= 40.0                                           # one vertex far outside: long clipped slivers
    nrm = rng.rand(ntri, 3, 3) * 2 - 1
    nrm[rng.rand(ntri) < 0.2] = 0.0                              # zero normals (FixNormals in the clipper)
    tex = rng.rand(ntri, 3, 3) * rng.choice([1.0, 3.0]) - rng.choice([0.0, 1.0])
    mesh = NewTriangleMesh(pos, normal=nrm, texture=tex) if ntri else Mesh()
    if ntri:
        mesh.color[:, :, :3] = rng.rand(ntri, 3, 3)
        mesh.color[:, :, 3] = np.where(rng.rand(ntri, 3) < 0.5, 1.0, rng.rand(ntri, 3))
    if nline:
        lp = (rng.rand(nline, 2, 3) * 2 - 1) * spread * 1.5
        lines = NewLineMesh(lp)
        lines.lcolor[:, :, :3] = rng.rand(nline, 2, 3)
        lines.lcolor[:, :, 3] = np.where(rng.rand(nline, 2) < 0.5, 1.0, rng.rand(nline, 2))
        lines.ltexture[:, :, :2] = rng.rand(nline, 2, 2)
        mesh.Add(lines)
    return mesh


def _script(seed):
    rng = np.random.RandomState(seed)
    W, H = int(rng.randint(3, 300)), int(rng.randint(3, 200))
    if rng.rand() < 0.5:
        eye = V(*(rng.rand(3) * 4 - 2 + np.array([0, 0, 3.0])))
        matrix = LookAt(eye, V(0, 0, 0), V(0, 1, 0)).Perspective(float(rng.uniform(20, 90)), W / H, float(rng.uniform(0.3, 2)), 20)
    else:
        eye = V(0, 0, 5)
        matrix = Orthographic(-1, 1, -1, 1, -2, 2)
    draws = []
    for _ in range(int(rng.randint(1, 4))):
        mesh = _random_mesh(rng, int(rng.choice([0, 5, 60, 300])), int(rng.choice([0, 0, 8, 40])))
        kind = rng.randint(4)
        tex = _random_texture(rng)
        if kind == 0:
            shader = NewSolidColorShader(matrix, Color(*rng.rand(3), float(rng.choice([1.0, 0.5, 0.0]))))
        elif kind == 1:
            shader = NewTextureShader(matrix, tex)
        else:
            shader = NewPhongShader(matrix, V(*(rng.rand(3) * 2 - 1)), eye)
            mode = rng.randint(3)
            if mode == 0:
                shader.ObjectColor = Color(*rng.rand(3), float(rng.choice([1.0, 0.65])))
            elif mode == 1:
                shader.Texture = tex
            shader.SpecularPower = float(rng.choice([0, 1, 7, 32]))
        state = {"ReadDepth": rng.rand() < 0.85, "WriteDepth": rng.rand() < 0.85, "WriteColor": rng.rand() < 0.9,
                 "AlphaBlend": rng.rand() < 0.7, "Wireframe": rng.rand() < 0.25, "FrontFace": int(rng.choice([1, 2])),
                 "Cull": int(rng.choice([1, 2, 3])), "LineWidth": float(rng.choice([0.5, 1.0, 2.0, 5.5])),
                 "DepthBias": float(rng.choice([0.0, 0.0, -1e-4, 1e-3]))}
        draws.append((mesh, shader, state))
    clear = Color(*rng.rand(4)) if rng.rand() < 0.7 else None
    return W, H, clear, draws, bool(rng.rand() < 0.25)


def _run(ctx, clear, draws):
    infos = []
    if clear is not None:
        ctx.ClearColorBufferWith(clear)
    for mesh, shader, state in draws:
        ctx.Shader = shader
        for k, v in state.items():
            setattr(ctx, k, v)
        infos.append(tuple(ctx.DrawMesh(mesh)))
    return infos


@pytest.mark.parametrize("front", ["fused", "split", "auto64"])
@pytest.mark.parametrize("block", range(16))
def test_random_scenes_match_oracle(block, front, oracle_lib, gpu_capi, monkeypatch):
    from fauxgl_b200.context import Context
    if front == "auto64":
        monkeypatch.setenv("FGL_STRIP_W", "64")                   # the library's own choice of front end, 64-pixel strips
    else:
        monkeypatch.setenv("FGL_FRONT", front)
    for seed in range(block * 8, block * 8 + 8):
        W, H, clear, draws, x_guard = _script(seed)
        octx = oracle_lib.OracleContext(W, H, x_guard=x_guard)
        oinfo = _run(octx, clear, draws)
        gctx = Context(W, H)
        gctx.XGuard = x_guard
        ginfo = _run(gctx, clear, draws)
        gd, gc = gctx.DepthBuffer, gctx.Image()
        dm = int((gd.view(np.uint64) != octx.DepthBuffer.view(np.uint64)).sum())
        cm = int((gc != octx.ColorBuffer).any(axis=-1).sum())
        gctx.Close()
        assert ginfo == oinfo and dm == 0 and cm == 0, (seed, front, W, H, ginfo, oinfo, dm, cm)


def _edge_script(seed):
    """Corner cases on purpose: framebuffers of a few pixels, line widths larger than the screen, vertices on and behind the
    eye plane (w = 0, w < 0), texture coordinates exactly on texel and wrap boundaries, depth cleared to arbitrary values."""
    rng = np.random.RandomState(10_000 + seed)
    W, H = int(rng.choice([1, 2, 3, 5, 8, 33, 64, 65])), int(rng.choice([1, 2, 3, 7, 32, 33]))
    eye = V(0, 0, float(rng.choice([0.0, 0.5, 2.0])))                 # 0.0: the eye sits inside the geometry
    matrix = LookAt(eye, V(0, 0, -1), V(0, 1, 0)).Perspective(float(rng.choice([30, 90, 150])), W / H, float(rng.choice([0.01, 0.5])), 10)
    draws = []
    for _ in range(int(rng.randint(1, 4))):
        ntri, nline = int(rng.choice([1, 20, 150])), int(rng.choice([0, 3, 20]))
        pos = (rng.rand(ntri, 3, 3) * 2 - 1) * np.array([2.0, 2.0, 3.0])
        pos[rng.rand(ntri) < 0.3, :, 2] = eye[2]                       # whole triangles in the eye plane: w = 0
        tex = np.round(rng.rand(ntri, 3, 3) * 8) / 4 - 0.5                # u, v on multiples of 1/4 in [-0.5, 1.5]
        mesh = NewTriangleMesh(pos, texture=tex)
        mesh.color[:, :, :3] = rng.rand(ntri, 3, 3)
        mesh.color[:, :, 3] = rng.choice([0.0, 0.5, 1.0], size=(ntri, 3))  # alpha 0 everywhere on a triangle: Discard if the colour is black
        mesh.color[rng.rand(ntri) < 0.1] = 0.0
        if nline:
            lines = NewLineMesh((rng.rand(nline, 2, 3) * 2 - 1) * 3.0)
            lines.lcolor[:, :, :] = rng.rand(nline, 2, 4)
            mesh.Add(lines)
        kind = rng.randint(3)
        if kind == 0:
            shader = NewSolidColorShader(matrix, Color(*rng.rand(3), float(rng.choice([1.0, 0.25]))))
        elif kind == 1:
            shader = NewTextureShader(matrix, _random_texture(rng))
        else:
            shader = NewPhongShader(matrix, V(0.3, 0.4, 1.0), eye)    # vertex colours (ObjectColor == Discard)
            shader.SpecularPower = float(rng.choice([0, 2.5, 100]))   # 2.5: Go's Pow with a fractional exponent
        state = {"ReadDepth": rng.rand() < 0.8, "WriteDepth": rng.rand() < 0.8, "WriteColor": True,
                 "AlphaBlend": rng.rand() < 0.7, "Wireframe": rng.rand() < 0.3, "FrontFace": 2,
                 "Cull": int(rng.choice([1, 3])), "LineWidth": float(rng.choice([0.0, 1.0, 9.0, 200.0])),
                 "DepthBias": float(rng.choice([0.0, -1e-5]))}
        draws.append((mesh, shader, state, float(rng.choice([np.finfo(np.float64).max, 1.0, 0.5]))))
    return W, H, draws


@pytest.mark.parametrize("block", range(8))
def test_random_corner_cases_match_oracle(block, oracle_lib, gpu_capi):
    """Same comparison, plus the per-primitive RasterizeInfo of every draw (fgl_draw_*_each vs a loop of DrawTriangle /
    DrawLine in the oracle)."""
    from fauxgl_b200.context import Context
    for seed in range(block * 6, block * 6 + 6):
        W, H, draws = _edge_script(seed)
        octx, gctx = oracle_lib.OracleContext(W, H), Context(W, H)
        for di, (mesh, shader, state, cleard) in enumerate(draws):
            for c in (octx, gctx):
                c.Shader = shader
                for k, v in state.items():
                    setattr(c, k, v)
                if di % 2 == 1:
                    c.ClearDepthBufferWith(cleard)
            if di % 2 == 0:
                oi, gi = tuple(octx.DrawMesh(mesh)), tuple(gctx.DrawMesh(mesh))
                assert oi == gi, (seed, di, oi, gi)
            else:
                ot, gt = octx.DrawTrianglesEach(mesh), gctx.DrawTrianglesEach(mesh)
                assert (ot == gt).all(), (seed, di, np.nonzero((ot != gt).any(axis=1))[0][:5])
                if mesh.num_lines:
                    ol, gl = octx.DrawLinesEach(mesh), gctx.DrawLinesEach(mesh)
                    assert (ol == gl).all(), (seed, di, np.nonzero((ol != gl).any(axis=1))[0][:5])
            dm = int((gctx.DepthBuffer.view(np.uint64) != octx.DepthBuffer.view(np.uint64)).sum())
            cm = int((gctx.Image() != octx.ColorBuffer).any(axis=-1).sum())
            assert dm == 0 and cm == 0, (seed, di, W, H, dm, cm)
        gctx.Close()
