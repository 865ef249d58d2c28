"""Host-side shader and texture objects mirroring the reference's ``shader.go``
and ``texture.go``.

Only the three built-in shaders have device implementations (shader.go:11,30,49)
-- the same closed set the Go shim resolves by type switch.  Anything else is an
error at draw time; there is no CPU fallback.
"""
from __future__ import annotations

import numpy as np

from .color import Color, Discard
from .matrix import Matrix
from .vector import Vector

SHADER_SOLID, SHADER_TEXTURE, SHADER_PHONG = 1, 2, 3
TEX_RGBA, TEX_NRGBA, TEX_RGBA64 = 0, 1, 2


class ImageTexture:
    """texture.go:21-30.  ``pixels`` is (H,W,4) uint8.  ``format`` records which
    Go image type the decoder would have produced, because MakeColor
    (color.go:25-29) goes through its RGBA() method: TEX_RGBA for *image.RGBA
    (8-bit RGB PNGs decode to it with A=255), TEX_NRGBA for *image.NRGBA
    (8-bit RGBA PNGs), whose RGBA() premultiplies; TEX_RGBA64 for every other
    image type: ``pixels`` is (H,W,4) uint16 holding what At(x,y).RGBA() returns."""

    def __init__(self, pixels: np.ndarray, format: int = TEX_RGBA):
        pixels = np.ascontiguousarray(pixels, dtype=np.uint16 if int(format) == TEX_RGBA64 else np.uint8)
        assert pixels.ndim == 3 and pixels.shape[2] == 4
        self.pixels = pixels
        self.format = int(format)
        self.Height, self.Width = pixels.shape[0], pixels.shape[1]


def NewImageTexture(pixels: np.ndarray, format: int = TEX_RGBA) -> ImageTexture:
    return ImageTexture(pixels, format)


def LoadTexture(path: str) -> ImageTexture:
    """texture.go:13-19 via PIL (PNG only: Go's JPEG decoder yields *image.YCbCr,
    a different integer colour path -- see SURVEY A.13)."""
    from PIL import Image
    im = Image.open(path)
    if im.mode == "RGBA":
        return ImageTexture(np.array(im), TEX_NRGBA)
    rgb = np.array(im.convert("RGB"))
    px = np.concatenate([rgb, np.full(rgb.shape[:2] + (1,), 255, np.uint8)], axis=2)
    return ImageTexture(px, TEX_RGBA)


class SolidColorShader:
    """shader.go:11-27"""

    def __init__(self, matrix: Matrix, color: Color):
        self.Matrix, self.Color = matrix, color

    def describe(self):
        return {"kind": SHADER_SOLID, "matrix": tuple(self.Matrix), "color": tuple(self.Color)}


class TextureShader:
    """shader.go:30-46"""

    def __init__(self, matrix: Matrix, texture: ImageTexture):
        self.Matrix, self.Texture = matrix, texture

    def describe(self):
        return {"kind": SHADER_TEXTURE, "matrix": tuple(self.Matrix), "texture": self.Texture}


class PhongShader:
    """shader.go:49-96; defaults from NewPhongShader (shader.go:61-68)."""

    def __init__(self, matrix: Matrix, lightDirection: Vector, cameraPosition: Vector):
        self.Matrix = matrix
        self.LightDirection = lightDirection
        self.CameraPosition = cameraPosition
        self.ObjectColor = Discard
        self.AmbientColor = Color(0.2, 0.2, 0.2, 1.0)
        self.DiffuseColor = Color(0.8, 0.8, 0.8, 1.0)
        self.SpecularColor = Color(1.0, 1.0, 1.0, 1.0)
        self.Texture = None
        self.SpecularPower = 32.0

    def describe(self):
        return {
            "kind": SHADER_PHONG, "matrix": tuple(self.Matrix),
            "light": tuple(self.LightDirection), "camera": tuple(self.CameraPosition),
            "object": tuple(self.ObjectColor), "ambient": tuple(self.AmbientColor),
            "diffuse": tuple(self.DiffuseColor), "specular": tuple(self.SpecularColor),
            "specular_power": float(self.SpecularPower), "texture": self.Texture,
        }


def NewSolidColorShader(matrix, color) -> SolidColorShader:
    return SolidColorShader(matrix, color)


def NewTextureShader(matrix, texture) -> TextureShader:
    return TextureShader(matrix, texture)


def NewPhongShader(matrix, lightDirection, cameraPosition) -> PhongShader:
    return PhongShader(matrix, lightDirection, cameraPosition)
