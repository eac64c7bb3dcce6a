"""CPU tests, row N2: the oracle's FAST-9 restatement (oracle/oracle_fast.cpp) is pinned bit-for-bit against the REAL reference library
compiled from /root/reference/thirdparty/fast (oracle/_ref/libfast_ref.so; built by `make -C oracle` where the reference is mounted, travels
to the GPU box as a prebuilt file). Corner coordinates, scores and non-max survivors are integers: exact equality."""
import numpy as np
import pytest

import oracle_lib as O
from hso_b200 import synth

needs_ref = pytest.mark.skipif(not O.ref_fast_available(), reason="oracle/_ref/libfast_ref.so not built (no /root/reference here)")


def _images():
    rng = np.random.default_rng(17)
    yield "texture640", synth.texture(rng, 640, 480), 20
    yield "texture_lowthr", synth.texture(rng, 320, 240), 7
    yield "noise", rng.integers(0, 256, (120, 188), dtype=np.uint8), 40           # width not a multiple of 16 (EuRoC level 2)
    yield "noise_odd", rng.integers(0, 256, (61, 47), dtype=np.uint8), 25         # odd sizes, tail loop of the SSE2 kernel
    yield "narrow", rng.integers(0, 256, (40, 21), dtype=np.uint8), 30            # img_width < 22 -> plain detector path
    sat = rng.integers(0, 256, (90, 100), dtype=np.uint8)
    sat[sat > 200] = 255
    sat[sat < 50] = 0
    yield "saturated", sat, 30                                                     # p +/- b crossing 0 / 255 (saturating SSE2 arithmetic)
    yield "flat", np.full((64, 64), 128, np.uint8), 10


@needs_ref
@pytest.mark.parametrize("name,img,thr", list(_images()), ids=[n for n, _, _ in _images()])
def test_restatement_equals_real_reference(name, img, thr):
    xy_r, sc_r, nm_r = O.ref_fast9(img, thr)
    xy_o, sc_o = O.fast9_corners(img, thr)
    assert xy_o.shape == xy_r.shape, (name, xy_o.shape, xy_r.shape)
    assert np.array_equal(xy_o, xy_r) and np.array_equal(sc_o, sc_r)
    # non-max survivors (border = 0 disables the 8-px filter of fastDetectST)
    det = O.fast_detect(img, thr, border=0)
    surv_r = [(int(xy_r[i, 0]), int(xy_r[i, 1]), int(sc_r[i])) for i in nm_r]
    assert [(x, y, s) for x, y, s, _ in det] == surv_r
    if name == "texture640":
        assert len(surv_r) > 200


def test_border_filter_and_shi_tomasi():
    rng = np.random.default_rng(3)
    img = synth.texture(rng, 160, 120)
    all_ = O.fast_detect(img, 12, border=0)
    kept = O.fast_detect(img, 12, border=8)
    exp = [c for c in all_ if not (c[0] < 8 or c[0] > 160 - 8 or c[1] < 8 or c[1] > 120 - 8)]   # feature_detection.cpp:515
    assert [(x, y, s) for x, y, s, _ in kept] == [(x, y, s) for x, y, s, _ in exp] and len(kept) > 10
    # Shi-Tomasi: integer sums, so the numpy statement is exact up to the final float expression
    f = img.astype(np.float64)
    for x, y, s, st in kept[:20]:
        dx = f[y - 4:y + 4, x - 3:x + 5] - f[y - 4:y + 4, x - 5:x + 3]
        dy = f[y - 3:y + 5, x - 4:x + 4] - f[y - 5:y + 3, x - 4:x + 4]
        a, b, c = (dx * dx).sum() / 128, (dy * dy).sum() / 128, (dx * dy).sum() / 128
        exp_st = 0.5 * (a + b - np.sqrt((a + b) ** 2 - 4 * (a * b - c * c)))
        assert abs(st - exp_st) <= 1e-4 * max(exp_st, 1.0)
