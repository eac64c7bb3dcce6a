"""GPU parity, row N4 (input side): hso_frame_upload_raw_batch = ImageReader resize + AbstractCamera::undistortImage + Frame construction,
against OpenCV golden vectors directly (cv2 4.13) and against the CPU oracle at full sizes. Integer pipelines: bit-exact."""
import os

import numpy as np
import pytest

from hso_b200 import Context, make_cam, synth

pytestmark = pytest.mark.gpu

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv_golden2.npz"))


def _ctx(c, **kw):
    return Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), **kw)


@pytest.mark.parametrize("tag,dkey", [("remap", "remap_d"), ("remap2", "remap2_d")])
def test_undistort_equals_cv2_golden(tag, dkey):
    K = G["remap_K"]
    W, H = [int(v) for v in G["remap_size"]]
    ctx = _ctx(dict(width=W, height=H, fx=K[0, 0], fy=K[1, 1], cx=K[0, 2], cy=K[1, 2], d=G[dkey], model=0))
    m1, m2 = ctx.undistort_maps()
    assert np.array_equal(m1, G[tag + "_map1"]) and np.array_equal(m2, G[tag + "_map2"])  # cv::initUndistortRectifyMap
    ids, _, _ = ctx.upload_raw_frames([G["remap_src"]], undistort=True)
    assert np.array_equal(ctx.download_level(ids[0], 0), G[tag + "_dst"])                  # cv::remap
    ctx.close()


def test_reader_resize_equals_cv2_golden():
    src = G["reader_src"]
    ctx = _ctx(dict(width=115, height=92, fx=100.0, fy=100.0, cx=57.0, cy=46.0, d=(0, 0, 0, 0, 0), model=0))
    ids, _, _ = ctx.upload_raw_frames([src], undistort=False)
    assert np.array_equal(ctx.download_level(ids[0], 0), G["reader_dst"])
    ctx.close()


@pytest.mark.parametrize("cam,raw", [("euroc", None), ("tum_fov", None), ("tum_fov", (1280, 1024)), ("equi", (1280, 1024)), ("icl", (1000, 750))])
def test_raw_upload_full_size_vs_oracle(oracle, cam, raw):
    if cam == "equi":
        c = dict(width=920, height=736, fx=0.35 * 920, fy=0.44 * 736, cx=460.3, cy=367.2, d=(0.9, 0.4, 0.15, 0.05, 0), model=2)  # maps leave the source at the rim
    else:
        c = synth.CAMS[cam]
    W, H = c["width"], c["height"]
    rw, rh = raw if raw else (W, H)
    rng = np.random.default_rng(9)
    imgs = [synth.texture(rng, rw, rh) for _ in range(3)]
    undist = cam != "icl"
    ctx = _ctx(c, max_frames=8)
    ids, integ, gm = ctx.upload_raw_frames(imgs, undistort=undist)
    if undist:
        m1, m2 = oracle.init_undistort_maps(c)
        g1, g2 = ctx.undistort_maps()
        assert np.array_equal(g1, m1) and np.array_equal(g2, m2)
    for b in range(3):
        exp = imgs[b]
        if raw:
            exp = oracle.resize_linear(exp, W, H)       # ImageReader::readImage
        if undist:
            exp = oracle.remap_linear(exp, m1, m2)      # cam->undistortImage
        assert np.array_equal(ctx.download_level(ids[b], 0), exp), (cam, raw, b)
        lv, _ = oracle.create_pyramid(exp, 5)            # the Frame built from it
        for l in range(1, 5):
            assert np.array_equal(ctx.download_level(ids[b], l), lv[l]), (cam, l)
        oi, og = oracle.frame_stats(exp)
        assert abs(integ[b] - oi) <= 5e-5 * abs(oi) and abs(gm[b] - og) <= 2.5e-4 * abs(og)
    if cam == "equi":
        assert (ctx.download_level(ids[0], 0) == 0).mean() > 0.01  # BORDER_CONSTANT band present
    ctx.close()


def test_raw_upload_rejects_bad_arguments():
    from hso_b200 import HsoError
    c = synth.CAMS["icl"]
    ctx = _ctx(c)
    with pytest.raises(HsoError):
        ctx._chk(ctx.lib.hso_frame_upload_raw_batch(ctx.h, 1, None, 640, 480, 640, 1, None, None, None))
    ctx.close()


def test_raw_upload_with_row_padding_equals_tight(oracle):
    """cv::Mat rows may be padded (step.p[0] > cols): the stride argument is honoured on both the resize and the undistort path."""
    c = synth.CAMS["euroc"]
    W, H = c["width"], c["height"]
    rng = np.random.default_rng(4)
    img = synth.texture(rng, W, H)
    padded = np.zeros((H, W + 48), np.uint8)
    padded[:, :W] = img
    ctx = _ctx(c, max_frames=8)
    a, ia, ga = ctx.upload_raw_frames([img], undistort=True)
    b, ib, gb = ctx.upload_raw_frames([padded[:, :W]], undistort=True)   # a view: strides[0] = W + 48
    for l in range(5):
        assert np.array_equal(ctx.download_level(a[0], l), ctx.download_level(b[0], l)), l
    assert ia[0] == ib[0] and ga[0] == gb[0]
    ctx.close()
