"""shared helpers for the parity tests"""
from __future__ import annotations

import numpy as np

LAYER_NAMES = ["elevation", "variance", "intensity", "color_r", "color_g", "color_b", "lowest"]


def assert_layers_equal(gpu_map, orc_map, names=LAYER_NAMES, rtol=0.0, what=""):
    """bit-exact by default (the CUDA path and the oracle share one arithmetic definition);
    rtol > 0 switches float layers to the 1e-5 relative tolerance of BASELINE.json."""
    for name in names:
        a = gpu_map.get_layer(name)
        b = orc_map.get_layer(name)
        if a.dtype.kind == "f" and rtol > 0:
            ok = np.isclose(a, b, rtol=rtol, atol=0.0, equal_nan=True)
        elif a.dtype.kind == "f":
            ok = (a.view(np.uint32) == b.view(np.uint32)) | (np.isnan(a) & np.isnan(b))
        else:
            ok = a == b
        if not ok.all():
            bad = np.argwhere(~ok)
            i = tuple(bad[0])
            raise AssertionError(f"{what} layer {name}: {bad.shape[0]} cells differ, first at {i}: "
                                 f"gpu={a[i]!r} oracle={b[i]!r}")


def split_rgb(rgba):
    return tuple(rgba[:, k].astype(np.int32) for k in range(3))
