"""Generate the committed parity fixtures from the reference's example assets.

Run HERE (the build container), where /root/reference exists:

    python tests/golden/make_fixtures.py

It loads the reference's own example meshes with this repo's restatement of the
reference's loaders (stl.go / obj.go), and stores the *raw loaded* attribute
arrays (float64, before BiUnitCube / normal smoothing, which the tests apply
themselves) as compressed .npz, because /root/reference does not exist on the
GPU box.  It then renders every scene in tests/scenes.py with the CPU oracle
and records sha256 digests + RasterizeInfo in oracle_golden.json, which pins the
oracle against regressions (the reference ships no golden vectors of its own:
parity is otherwise unpinned).
"""
from __future__ import annotations

import hashlib
import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from fauxgl_b200 import LoadOBJ, LoadSTL  # noqa: E402

REF = "/root/reference/examples"
OUT = os.path.dirname(os.path.abspath(__file__))


def save_mesh(name, mesh, with_texture=False):
    arrays = {"position": mesh.position, "normal": mesh.normal}
    if with_texture:
        arrays["texture"] = mesh.texture
    np.savez_compressed(os.path.join(OUT, name + ".npz"), **arrays)
    print(name, mesh.num_triangles, "triangles")


def main():
    save_mesh("hello_mesh", LoadSTL(os.path.join(REF, "hello.stl")))
    save_mesh("bowser_mesh", LoadSTL(os.path.join(REF, "bowser.stl")))
    save_mesh("capsule_mesh", LoadOBJ(os.path.join(REF, "capsule.obj")), with_texture=True)
    save_mesh("cube_mesh", LoadSTL(os.path.join(REF, "cube.stl")))
    # examples/texture.png (examples/square.go:36): the same pixels, re-encoded without the metadata chunks
    from PIL import Image
    im = Image.open(os.path.join(REF, "texture.png")).convert("RGB")
    im.save(os.path.join(OUT, "texture.png"), optimize=True)
    assert (np.asarray(Image.open(os.path.join(OUT, "texture.png"))) == np.asarray(im)).all()
    print("texture.png", im.size, os.path.getsize(os.path.join(OUT, "texture.png")), "bytes")

    import scenes
    from oracle.pyoracle import OracleContext
    golden = {}
    for name in scenes.GOLDEN_SCENES:
        sc = scenes.SCENES[name]()
        ctx = OracleContext(sc.width, sc.height)   # the reference's own index rule (x_guard off)
        infos = sc.run(ctx)
        golden[name] = {
            "width": sc.width, "height": sc.height,
            "info": [[int(a), int(b)] for a, b in infos],
            "color_sha256": hashlib.sha256(ctx.ColorBuffer.tobytes()).hexdigest(),
            "depth_sha256": hashlib.sha256(ctx.DepthBuffer.tobytes()).hexdigest(),
            "covered": int((ctx.DepthBuffer < 1e300).sum()),
        }
        print(name, golden[name]["info"], golden[name]["covered"])
    with open(os.path.join(OUT, "oracle_golden.json"), "w") as f:
        json.dump(golden, f, indent=1, sort_keys=True)


if __name__ == "__main__":
    main()
