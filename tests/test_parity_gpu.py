"""GPU parity: every scene rendered through the C ABI on the device must equal
the CPU oracle's sequential, triangle-index-order render of the same inputs.

Bar (BASELINE.json north_star): same coverage + depth-test outcome on >= 99.99 %
of pixels, <= 1/255 colour error on matching pixels.  The kernels reproduce the
reference's forward-differencing adds literally, so these tests assert the
stronger property: bit-identical float64 depth, identical NRGBA8 colour and
identical RasterizeInfo -- zero mismatches.
"""
import pytest

import scenes
from parity import run_both

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name", list(scenes.SCENES))
def test_scene_matches_oracle(name, oracle_lib, gpu_capi):
    from fauxgl_b200.context import Context
    stats = run_both(scenes.SCENES[name](), oracle_lib, Context)
    print(name, stats)
    assert stats["depth_mismatch"] == 0, stats
    assert stats["color_mismatch"] == 0, stats
    assert stats["gpu_info"] == stats["oracle_info"], stats
