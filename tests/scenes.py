"""Parity scenes: each is a script against the reference-shaped Context API
(the same calls the reference's examples make), runnable unchanged on the CPU
oracle (oracle.pyoracle.OracleContext) and on the GPU back end
(fauxgl_b200.Context).  Constants come from the reference's examples
(SURVEY.md Appendix B); citations are into /root/reference.
"""
from __future__ import annotations

import os
from typing import Callable, Dict, List

import numpy as np

from fauxgl_b200 import (Black, Color, CullBack, CullFront, CullNone, FaceCCW, FaceCW, Gray, HexColor,
                         Identity, LookAt, Mesh, NewImageTexture, NewPhongShader, NewSolidColorShader,
                         NewTextureShader, NewTriangleMesh, NewLineMesh, Orthographic, Radians, Rotate,
                         Scale, Translate, V, White)
from fauxgl_b200 import synth
from fauxgl_b200.shader import TEX_NRGBA, TEX_RGBA

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def load_fixture(name: str) -> Mesh:
    d = np.load(os.path.join(GOLDEN, name + ".npz"))
    return Mesh(d["position"], d["normal"], d["texture"] if "texture" in d else None)


def reference_texture() -> "ImageTexture":
    """examples/texture.png of the reference (4096 x 4096 opaque RGB UV-checker, examples/square.go:36), re-encoded
    pixel for pixel by tests/golden/make_fixtures.py.  Go's PNG decoder yields *image.RGBA (alpha 255) for it."""
    from PIL import Image
    rgb = np.asarray(Image.open(os.path.join(GOLDEN, "texture.png")).convert("RGB"))
    px = np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2)
    return NewImageTexture(px, TEX_RGBA)


def checker_texture(n=256, alpha=False) -> "ImageTexture":
    """Small deterministic procedural texture (the NRGBA variant exercises the texel-alpha blend path,
    which the opaque texture.png cannot)."""
    y, x = np.mgrid[0:n, 0:n]
    px = np.zeros((n, n, 4), dtype=np.uint8)
    chk = ((x // 16) + (y // 16)) & 1
    px[..., 0] = np.where(chk, 230, 40) ^ (x & 0x1f)
    px[..., 1] = (x * 255 // (n - 1)).astype(np.uint8)
    px[..., 2] = (y * 255 // (n - 1)).astype(np.uint8)
    px[..., 3] = 255 if not alpha else (64 + ((x + y) * 191 // (2 * n - 2))).astype(np.uint8)
    return NewImageTexture(px, TEX_NRGBA if alpha else TEX_RGBA)


class Scene:
    def __init__(self, width: int, height: int, run: Callable):
        self.width, self.height, self._run = width, height, run

    def run(self, ctx) -> List[tuple]:
        return [tuple(i) for i in self._run(ctx)]


def _camera(eye, center, up, fovy, aspect, near, far):
    return LookAt(eye, center, up).Perspective(fovy, aspect, near, far)


# ---- config 0: examples/hello.go as shipped ------------------------------------------------
def hello():
    mesh = load_fixture("hello_mesh")
    mesh.BiUnitCube()
    mesh.SmoothNormalsThreshold(Radians(30))
    W, H = 2000, 1000
    eye, center, up = V(-1, -2, 2), V(-0.07, 0, 0), V(0, 0, 1)

    def run(ctx):
        ctx.ClearColor = Black
        ctx.ClearColorBuffer()
        matrix = _camera(eye, center, up, 20, W / H, 1, 50)
        shader = NewPhongShader(matrix, V(-2, 0, 1).Normalize(), eye)
        shader.ObjectColor = Color(0.5, 1, 0.65, 1)
        ctx.Shader = shader
        return [ctx.DrawMesh(mesh)]
    return Scene(W, H, run)


# ---- config 1: examples/teapot.go camera/shader on bowser.stl (583 triangles need clipping) ----
def _bowser_scene(W, H):
    mesh = load_fixture("bowser_mesh")
    mesh.BiUnitCube()
    mesh.SmoothNormalsThreshold(Radians(60))
    eye, center, up = V(0, 2.4, 0), V(0, 0, 0), V(0, 0, 1)

    def run(ctx):
        ctx.ClearColorBufferWith(White)
        matrix = _camera(eye, center, up, 30, 1920 / 1080, 1, 10)
        shader = NewPhongShader(matrix, V(0, 1, 1).Normalize(), eye)
        shader.ObjectColor = HexColor("#B9121B")
        shader.SpecularColor = Gray(0.25)
        shader.SpecularPower = 64
        ctx.Shader = shader
        return [ctx.DrawMesh(mesh)]
    return Scene(W, H, run)


def bowser():
    return _bowser_scene(1920, 1080)


def bowser_close():
    """Same mesh with the eye inside the bounding cube and DepthBias/cull variants:
    heavy near/side-plane clipping, CullNone so clipped back faces also draw."""
    mesh = load_fixture("bowser_mesh")
    mesh.BiUnitCube()
    eye, center, up = V(0.2, 1.05, 0.1), V(0, 0, 0), V(0, 0, 1)

    def run(ctx):
        ctx.ClearColorBufferWith(HexColor("#24221F"))
        matrix = _camera(eye, center, up, 60, 800 / 600, 0.5, 10)
        shader = NewPhongShader(matrix, V(0.25, 0.5, 1).Normalize(), eye)
        shader.ObjectColor = HexColor("#FEB41C")
        shader.DiffuseColor = Gray(0.9)
        shader.SpecularColor = Gray(0.25)
        shader.SpecularPower = 100
        ctx.Shader = shader
        ctx.Cull = CullNone
        return [ctx.DrawMesh(mesh)]
    return Scene(800, 600, run)


# ---- config 3: capsule.obj + texture (TextureShader, then Phong+texture), examples/square.go:54, capsule.go:51 ----
def capsule_texture():
    mesh = load_fixture("capsule_mesh")
    mesh.BiUnitCube()
    tex = checker_texture(256)
    eye, center, up = V(-3, 1, 2), V(0, 0, 0), V(0, 0, 1)

    def run(ctx):
        ctx.ClearColorBufferWith(HexColor("#FFF8E3"))
        matrix = _camera(eye, center, up, 40, 1.0, 1, 10)
        ctx.Shader = NewTextureShader(matrix, tex)
        return [ctx.DrawMesh(mesh)]
    return Scene(1024, 1024, run)


def capsule_phong_texture():
    mesh = load_fixture("capsule_mesh")
    mesh.BiUnitCube()
    tex = checker_texture(128, alpha=True)
    eye, center, up = V(-3, 1, 2), V(0, 0, 0), V(0, 0, 1)

    def run(ctx):
        ctx.ClearColorBufferWith(HexColor("#102030"))
        matrix = _camera(eye, center, up, 40, 1.0, 1, 10)
        shader = NewPhongShader(matrix, V(-1, 1, 0.25).Normalize(), eye)
        shader.Texture = tex
        ctx.Shader = shader
        return [ctx.DrawMesh(mesh)]   # texel alpha < 1 -> blend path
    return Scene(768, 768, run)


# ---- config 4 as composed in SURVEY 0.1: capsule.obj + texture.png through TextureShader (examples/square.go:54),
# then the translucent and the wireframe + DepthBias passes of examples/shapes.go:71-78, capsule.go's camera ----
def capsule_composed(scale=1):
    mesh = load_fixture("capsule_mesh")
    mesh.BiUnitCube()
    tex = reference_texture()
    eye, center, up = V(-3, 1, 2), V(0, 0, 0), V(0, 0, 1)
    light = V(-1, 1, 0.25).Normalize()
    W = H = 1024 * scale

    def run(ctx):
        ctx.ClearColorBufferWith(White)
        matrix = _camera(eye, center, up, 40, 1.0, 1, 10)
        ctx.Shader = NewTextureShader(matrix, tex)
        infos = [ctx.DrawMesh(mesh)]
        shader = NewPhongShader(matrix, light, eye)
        shader.ObjectColor = HexColor("FFFF9D").Alpha(0.65)
        shader.SpecularPower = 0
        ctx.Shader = shader
        infos.append(ctx.DrawMesh(mesh))
        ctx.Wireframe = True
        ctx.DepthBias = -0.00001
        infos.append(ctx.DrawMesh(mesh))
        return infos
    return Scene(W, H, run)


def capsule_ycbcr_texture():
    """PhongShader + a texture whose Go image type is neither *image.RGBA nor *image.NRGBA (examples/capsule.go:35
    loads a JPEG -> *image.YCbCr): the host converts with At(x,y).RGBA() and uploads the 16-bit values
    (TEX_RGBA64).  The texels here are a deterministic stand-in run through Go's color.YCbCrToRGB restatement."""
    from fauxgl_b200.shader import TEX_RGBA64
    from fauxgl_b200.color import ycbcr_to_rgba64
    mesh = load_fixture("capsule_mesh")
    mesh.BiUnitCube()
    n = 192
    y, x = np.mgrid[0:n, 0:n]
    Y = (40 + (x * 3 + y * 5) % 200).astype(np.uint8)
    Cb = (128 + 90 * np.sin(x / 17.0)).astype(np.uint8)
    Cr = (128 + 90 * np.cos(y / 23.0)).astype(np.uint8)
    tex = NewImageTexture(ycbcr_to_rgba64(Y, Cb, Cr), TEX_RGBA64)
    eye, center, up = V(-3, 1, 2), V(0, 0, 0), V(0, 0, 1)

    def run(ctx):
        ctx.ClearColorBufferWith(HexColor("#101418"))
        matrix = _camera(eye, center, up, 40, 1.0, 1, 10)
        shader = NewPhongShader(matrix, V(-1, 1, 0.25).Normalize(), eye)
        shader.Texture = tex
        ctx.Shader = shader
        return [ctx.DrawMesh(mesh)]
    return Scene(640, 640, run)


# ---- examples/shapes.go:61-78: opaque pass, translucent pass, wireframe pass with DepthBias ----
def shapes_multipass():
    rng = np.random.RandomState(1234)
    mesh = Mesh()
    for _ in range(150):
        while True:
            p = rng.rand(3) * 2 - 1
            if (p * p).sum() < 1:
                break
        c = synth.NewCube()
        m = Rotate(V(*(rng.rand(3) * 2 - 1)), float(rng.rand() * 6.28)).Scale(V(0.2, 0.2, 0.2)).Translate(V(*(p * 4)))
        c.Transform(m)
        mesh.Add(c)
    sphere = synth.NewLatLngSphere(10, 10)
    sphere.SmoothNormals()
    sphere.Transform(Scale(V(2.5, 2.5, 2.5)))
    eye, center, up = V(12, 12, 6), V(0, 0, 0), V(0, 0, 1)
    light = V(0.75, 0.5, 1).Normalize()

    def run(ctx):
        ctx.ClearColorBufferWith(Black)
        matrix = _camera(eye, center, up, 30, 1.0, 1, 100)
        shader = NewPhongShader(matrix, light, eye)
        shader.ObjectColor = HexColor("#468966")
        ctx.Shader = shader
        infos = [ctx.DrawMesh(mesh)]
        shader = NewPhongShader(matrix, light, eye)
        shader.ObjectColor = HexColor("FFFF9D").Alpha(0.65)
        shader.SpecularPower = 0
        ctx.Shader = shader
        infos.append(ctx.DrawMesh(sphere))
        ctx.Wireframe = True
        ctx.DepthBias = -0.00001
        infos.append(ctx.DrawMesh(sphere))
        return infos
    return Scene(800, 800, run)


# ---- DrawLines: examples/magica.go:77-85 style (triangles, then lines with LineWidth/DepthBias) ----
def lines_scene():
    cube = synth.NewCube()
    cube.Transform(Scale(V(1.6, 1.6, 1.6)))
    # outlines of a lattice of boxes, some reaching beyond the view volume (ClipLine) and the screen edge
    lines = Mesh()
    for i, off in enumerate([(-0.8, -0.8, -0.8, 0.8, 0.8, 0.8), (-3, -3, -0.2, 3, 3, 0.2), (-0.2, -6, -1.5, 0.2, 6, 1.5)]):
        lines.Add(synth.NewCubeOutline(*off))
    lines.lcolor[:, :, :] = np.array([1.0, 0.3, 0.1, 1.0])
    eye, center, up = V(3, 2.5, 1.8), V(0, 0, 0), V(0, 0, 1)

    def run(ctx):
        ctx.ClearColorBufferWith(HexColor("#1D181F"))
        matrix = _camera(eye, center, up, 50, 900 / 500, 1, 20)
        shader = NewPhongShader(matrix, V(-0.75, -0.25, 1).Normalize(), eye)
        shader.ObjectColor = HexColor("#2A2C2B")
        shader.SpecularPower = 0
        ctx.Shader = shader
        infos = [ctx.DrawTriangles(cube)]
        ctx.Shader = NewSolidColorShader(matrix, HexColor("#7E827A"))
        ctx.LineWidth = 4
        ctx.DepthBias = -4e-5
        infos.append(ctx.DrawLines(lines))
        # vertex-coloured lines through Phong with ObjectColor == Discard
        shader = NewPhongShader(matrix, V(-0.75, -0.25, 1).Normalize(), eye)
        ctx.Shader = shader
        ctx.LineWidth = 1.5
        infos.append(ctx.DrawLines(lines))
        return infos
    return Scene(900, 500, run)



# ---- the reference's index rule (context.go:223-228): only i = y*W + x is range-checked -----------------------
def _offscreen_lines(W, H, lw, lw2):
    """Lines and wireframe edges that leave the framebuffer sideways.  ClipLine / ClipTriangle cut them at the view
    volume, i.e. exactly at x = 0 or x = W, and the fat-line quad then reaches LineWidth/2 beyond the border: the
    reference keeps those pixels -- they alias into the neighbouring row (depth and blended colour are written there,
    an opaque colour is dropped by SetNRGBA, all of them count in RasterizeInfo).  Corners: (W, -1) lands in row 0,
    (-1, H) in row H-1, (-1, 0) and (W, H-1) fall off the buffer."""
    quad = NewTriangleMesh(np.array([[(-0.95, -0.9, -0.1), (0.95, -0.9, 0.3), (0.95, 0.9, -0.1)],
                                     [(-0.95, -0.9, -0.1), (0.95, 0.9, -0.1), (-0.95, 0.9, 0.3)]], dtype=np.float64))
    segs = [((-2, 0.3, 0.2), (2, 0.3, 0.2)), ((0, 0, 0.1), (1.5, 1.5, 0.1)), ((0, 0, 0.1), (-1.5, -1.5, 0.1)),
            ((0.2, 0, 0.0), (-1.5, 1.7, 0.0)), ((-0.2, 0, 0.0), (1.5, -1.7, 0.0)), ((1, -0.5, 0.4), (1, 0.5, -0.4)),
            ((-1, -0.6, -0.3), (-1, 0.7, 0.5)), ((-3, -0.55, -0.2), (3, -0.5, 0.25)), ((0.97, -2, 0.15), (1.0, 2, 0.15))]
    lines = NewLineMesh(np.array(segs, dtype=np.float64))
    rng = np.random.RandomState(11)
    lines.lcolor[:, :, :3] = rng.rand(len(segs), 2, 3)
    lines.lcolor[:, :, 3] = 0.35 + 0.5 * rng.rand(len(segs), 2)
    big = NewTriangleMesh(np.array([[(-3, -2.5, 0.05), (3.5, -0.2, 0.05), (-0.3, 3, 0.05)],
                                    [(0.4, -0.8, -0.5), (2.5, 0.1, -0.5), (0.6, 0.7, -0.5)]], dtype=np.float64))
    big.color[:, :, :] = np.array([0.2, 0.9, 0.4, 0.6])

    def run(ctx):
        ctx.ClearColorBufferWith(HexColor("#20242A"))
        matrix = Orthographic(-1, 1, -1, 1, -1, 1)
        ctx.Cull = CullNone
        sh = NewPhongShader(matrix, V(0.3, 0.2, 1).Normalize(), V(0, 0, 5))
        sh.ObjectColor = HexColor("#8090A0")
        ctx.Shader = sh
        infos = [ctx.DrawMesh(quad)]
        ctx.Shader = NewSolidColorShader(matrix, HexColor("#E0C040"))      # opaque: depth-only winners + dropped colour
        ctx.LineWidth = lw
        infos.append(ctx.DrawLines(lines))
        ctx.Shader = NewPhongShader(matrix, V(0, 0, 1), V(0, 0, 5))        # vertex colours, alpha < 1: blend path
        ctx.DepthBias = -1e-3
        ctx.LineWidth = lw2
        infos.append(ctx.DrawLines(lines))
        ctx.Wireframe = True                                              # edges of clipped triangles run along the border
        ctx.LineWidth = lw
        infos.append(ctx.DrawMesh(big))
        ctx.Shader = NewSolidColorShader(matrix, Color(0.9, 0.3, 0.2, 1))
        ctx.DepthBias = -2e-3
        ctx.LineWidth = lw2
        infos.append(ctx.DrawMesh(big))
        return infos
    return Scene(W, H, run)


def offscreen_lines():
    return _offscreen_lines(203, 77, 7.0, 3.0)


def offscreen_lines_tiny():
    """Line width larger than the framebuffer: fragments alias several rows away."""
    return _offscreen_lines(40, 30, 90.0, 17.0)

# ---- render-state matrix on one small mesh -------------------------------------------------------
def _state_scene(apply):
    mesh = load_fixture("hello_mesh")
    mesh.BiUnitCube()
    rng = np.random.RandomState(7)
    mesh.color[:, :, :3] = rng.rand(mesh.num_triangles, 3, 3)
    mesh.color[:, :, 3] = 0.25 + 0.75 * rng.rand(mesh.num_triangles, 3)
    eye, center, up = V(-1, -2, 2), V(-0.07, 0, 0), V(0, 0, 1)

    def run(ctx):
        ctx.ClearColorBufferWith(HexColor("#334455"))
        matrix = _camera(eye, center, up, 20, 640 / 360, 1, 50)
        shader = NewPhongShader(matrix, V(-2, 0, 1).Normalize(), eye)  # ObjectColor == Discard -> vertex colours
        ctx.Shader = shader
        infos = [ctx.DrawMesh(mesh)]   # default state first, so the depth buffer is populated
        apply(ctx, shader)
        infos.append(ctx.DrawMesh(mesh))
        return infos
    return Scene(640, 360, run)


def state_cull_front():
    def apply(ctx, sh):
        ctx.Cull = CullFront
        sh.ObjectColor = Color(1, 0.2, 0.2, 0.5)
        ctx.ReadDepth = False
    return _state_scene(apply)


def state_cull_none_cw():
    def apply(ctx, sh):
        ctx.Cull = CullNone
        ctx.FrontFace = FaceCW
        ctx.DepthBias = -1e-4
    return _state_scene(apply)


def state_no_write_depth():
    def apply(ctx, sh):
        ctx.WriteDepth = False
        ctx.AlphaBlend = False
        ctx.Cull = CullNone
        ctx.DepthBias = -1e-3
    return _state_scene(apply)


def state_no_write_color():
    def apply(ctx, sh):
        ctx.WriteColor = False
        ctx.Cull = CullFront
        ctx.FrontFace = FaceCW
        ctx.ClearDepthBuffer()
    return _state_scene(apply)


def state_wireframe_solid():
    def apply(ctx, sh):
        ctx.Shader = NewSolidColorShader(sh.Matrix, Color(0, 0, 0, 1))
        ctx.Wireframe = True
        ctx.LineWidth = 3
        ctx.DepthBias = -1e-4
    return _state_scene(apply)


# ---- edge cases ----------------------------------------------------------------------------------------
def edge_cases():
    """Ragged framebuffer (not a tile multiple), full-screen and degenerate triangles,
    orthographic matrix, sub-pixel slivers, coincident triangles (<= tie rule)."""
    W, H = 203, 77
    tri = []
    tri.append([(-5, -5, -0.3), (5, -5, -0.3), (0, 7, -0.3)])            # covers the whole screen (needs clipping)
    tri.append([(-0.5, -0.5, 0.0), (0.5, -0.5, 0.0), (0.0, 0.5, 0.0)])
    tri.append([(-0.5, -0.5, 0.0), (0.5, -0.5, 0.0), (0.0, 0.5, 0.0)])  # coincident: later index wins on <=
    tri.append([(0.1, 0.1, -0.2), (0.1, 0.1, -0.2), (0.4, 0.4, -0.2)])  # degenerate (zero area)
    tri.append([(-0.9, 0.2, -0.5), (-0.2, 0.21, -0.5), (-0.9, 0.2005, -0.5)])  # sliver
    tri.append([(0.3, -0.9, -0.6), (0.9, -0.9, -0.6), (0.9, -0.3, 0.9)])     # steep depth gradient
    pos = np.array(tri, dtype=np.float64)
    mesh = NewTriangleMesh(pos)
    rng = np.random.RandomState(3)
    mesh.color[:, :, :3] = rng.rand(len(pos), 3, 3)
    mesh.color[:, :, 3] = 1.0
    mesh.color[2, :, 3] = 0.5

    def run(ctx):
        ctx.ClearColorBufferWith(Gray(0.5))
        matrix = Orthographic(-1, 1, -1, 1, -1, 1)
        shader = NewPhongShader(matrix, V(0, 0, 1), V(0, 0, 5))
        shader.AmbientColor = Gray(0.6)
        ctx.Shader = shader
        ctx.Cull = CullNone
        infos = [ctx.DrawMesh(mesh)]
        infos.append(ctx.DrawTriangles(mesh, 1, 2))      # sub-range
        infos.append(ctx.DrawTriangles(mesh, 0, 0))      # empty range
        infos.append(ctx.DrawLines(Mesh()))              # empty mesh
        return infos
    return Scene(W, H, run)


def tiny_framebuffer():
    mesh = synth.NewCube()

    def run(ctx):
        matrix = _camera(V(2, 2, 2), V(0, 0, 0), V(0, 0, 1), 40, 5 / 3, 1, 10)
        shader = NewPhongShader(matrix, V(1, 0.5, 1).Normalize(), V(2, 2, 2))
        shader.ObjectColor = Color(0.9, 0.5, 0.1, 1)
        ctx.Shader = shader
        return [ctx.DrawMesh(mesh)]     # no clear: NewContext state (transparent, MaxFloat64)
    return Scene(5, 3, run)


def degenerate_clipped():
    """Degenerate triangles (two equal vertices) that need clipping: ClipTriangle's Barycentric divides 0 by 0, every
    attribute of the fan triangles -- Output included -- is NaN, `a <= 0` is false for NaN so they are drawn under
    every Cull mode, the integer bounding box is int(NaN) = -2^63 on all four sides, and the reference's loops visit
    one "pixel" whose index -2^63 * W - 2^63 wraps to 0 when the width is odd: pixel (0, 0) is counted, and with
    ReadDepth off it receives a NaN depth and (blend path) a NaN-derived colour."""
    tri = np.array([[(-0.3, -1.2, 0.1), (-0.3, -1.2, 0.1), (0.2, -0.7, 0.1)],      # v1 == v2, crosses y = -1
                    [(0.9, 0.1, 0.0), (1.4, 0.3, 0.0), (1.4, 0.3, 0.0)],            # v2 == v3, crosses x = 1
                    [(-0.5, -0.5, 0.2), (0.5, -0.5, 0.2), (0.0, 0.6, 0.2)]],        # an ordinary one
                   dtype=np.float64)
    mesh = NewTriangleMesh(tri)
    mesh.color[:, :, :] = np.array([0.8, 0.3, 0.2, 0.6])

    def run(ctx):
        ctx.ClearColorBufferWith(HexColor("#405060"))
        matrix = Orthographic(-1, 1, -1, 1, -1, 1)
        sh = NewPhongShader(matrix, V(0, 0, 1), V(0, 0, 5))
        sh.ObjectColor = HexColor("#C0D0E0")
        ctx.Shader = sh
        infos = [ctx.DrawMesh(mesh)]                 # CullBack: the NaN triangles are drawn all the same
        ctx.Shader = NewPhongShader(matrix, V(0, 0, 1), V(0, 0, 5))   # vertex colours, alpha 0.6: blend path
        ctx.ReadDepth = False
        infos.append(ctx.DrawMesh(mesh))
        return infos
    return Scene(171, 19, run)


# ---- synthetic benchmark meshes at test size ---------------------------------------------------------------
_BUMPY_CACHE: Dict[tuple, Mesh] = {}


def bumpy_mesh(nu, nv, triangles=None) -> Mesh:
    key = (nu, nv, triangles)
    if key not in _BUMPY_CACHE:
        n = nu * (2 * nv - 2)
        _BUMPY_CACHE[key] = synth.bumpy_surface(triangles or n, nu, nv)
    return _BUMPY_CACHE[key]


def dragon_scene(mesh: Mesh, W=1920, H=1080) -> Scene:
    """README.md:53-114 'Complete Example' (the 871k-triangle benchmark scene)."""
    eye, center, up = V(-3, 1, -0.75), V(0, -0.07, 0), V(0, 1, 0)

    def run(ctx):
        ctx.ClearDepthBuffer()
        ctx.ClearColorBufferWith(HexColor("#FFF8E3"))
        matrix = _camera(eye, center, up, 30, 1920 / 1080, 1, 10)
        shader = NewPhongShader(matrix, V(-0.75, 1, 0.25).Normalize(), eye)
        shader.ObjectColor = HexColor("#468966")
        ctx.Shader = shader
        return [ctx.DrawMesh(mesh)]
    return Scene(W, H, run)


def bumpy_small():
    return dragon_scene(bumpy_mesh(101, 101), 960, 540)


SCENES: Dict[str, Callable[[], Scene]] = {
    "hello": hello,
    "bowser": bowser,
    "bowser_close": bowser_close,
    "capsule_texture": capsule_texture,
    "capsule_phong_texture": capsule_phong_texture,
    "capsule_composed": capsule_composed,
    "capsule_ycbcr_texture": capsule_ycbcr_texture,
    "shapes_multipass": shapes_multipass,
    "lines": lines_scene,
    "offscreen_lines": offscreen_lines,
    "offscreen_lines_tiny": offscreen_lines_tiny,
    "state_cull_front": state_cull_front,
    "state_cull_none_cw": state_cull_none_cw,
    "state_no_write_depth": state_no_write_depth,
    "state_no_write_color": state_no_write_color,
    "state_wireframe_solid": state_wireframe_solid,
    "edge_cases": edge_cases,
    "tiny_framebuffer": tiny_framebuffer,
    "degenerate_clipped": degenerate_clipped,
    "bumpy_small": bumpy_small,
}
GOLDEN_SCENES = list(SCENES)
