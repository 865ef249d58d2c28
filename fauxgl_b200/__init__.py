"""fauxgl_b200 -- B200-native rasterisation back end behind fogleman/fauxgl's
Context API (NewContext ... DrawMesh ... Image()).

Host-side mirror of the reference's API in Python (the Go toolchain is absent in
the build image; go/fauxgl holds the cgo shim a Go user would import).  All
rendering goes through the C ABI in include/fauxgl_b200.h to hand-written
sm_100a CUDA kernels; there is no CPU fallback.
"""
from .vector import V, Vector, Radians
from .matrix import (Matrix, Identity, Translate, Scale, Rotate, Frustum, Orthographic,
                     Perspective, LookAt, Screen)
from .color import Color, HexColor, Gray, Discard, Transparent, Black, White
from .mesh import Mesh, NewTriangleMesh, NewLineMesh, LoadSTL, LoadOBJ, LoadMesh
from .shader import (ImageTexture, NewImageTexture, LoadTexture, SolidColorShader, TextureShader,
                     PhongShader, NewSolidColorShader, NewTextureShader, NewPhongShader)

FaceCW, FaceCCW = 1, 2                 # context.go:13-17
CullNone, CullFront, CullBack = 1, 2, 3  # context.go:21-26


def __getattr__(name):
    # Context and friends need the CUDA library; resolve lazily so that host-only
    # helpers import on machines where it has not been built yet.
    if name in ("Context", "NewContext", "RasterizeInfo", "DeviceMesh", "FauxglError", "capi"):
        from . import context as _ctx
        return getattr(_ctx, name)
    raise AttributeError(name)
