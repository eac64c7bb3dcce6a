"""CPU pins of the row-N4 oracle (oracle/oracle_undistort.cpp) against OpenCV golden vectors (cv2 4.13, tests/golden/cv_golden2.npz):
cv::initUndistortRectifyMap(CV_16SC2), cv::convertMaps, cv::remap(INTER_LINEAR, BORDER_CONSTANT) and the ImageReader resize — all bit-exact."""
import os

import numpy as np

import oracle_lib as O

G = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "cv_golden2.npz"))


def _cam(dkey):
    K = G["remap_K"]
    W, H = [int(v) for v in G["remap_size"]]
    return dict(model=0, width=W, height=H, fx=K[0, 0], fy=K[1, 1], cx=K[0, 2], cy=K[1, 2], d=G[dkey])


def test_init_undistort_rectify_map_bit_exact():
    for tag, dkey in (("remap", "remap_d"), ("remap2", "remap2_d")):
        m1, m2 = O.init_undistort_maps(_cam(dkey))
        assert np.array_equal(m1, G[tag + "_map1"]) and np.array_equal(m2, G[tag + "_map2"]), tag


def test_remap_linear_bit_exact():
    for tag in ("remap", "remap2"):
        assert np.array_equal(O.remap_linear(G["remap_src"], G[tag + "_map1"], G[tag + "_map2"]), G[tag + "_dst"]), tag


def test_convert_maps_and_border_branches_bit_exact():
    m1, m2 = O.convert_maps(G["conv_mapx"], G["conv_mapy"])
    assert np.array_equal(m1, G["conv_map1"]) and np.array_equal(m2, G["conv_map2"])
    dst = O.remap_linear(G["remap_src"], m1, m2)
    assert (G["conv_dst"] == 0).mean() > 0.15  # the maps leave the source on every side: BORDER_CONSTANT branches are exercised
    assert np.array_equal(dst, G["conv_dst"])


def test_image_reader_resize_bit_exact():
    dst = O.resize_linear(G["reader_src"], 115, 92)
    assert np.array_equal(dst, G["reader_dst"])


def test_fov_and_equidistant_maps_are_the_distortion_model():
    """FOVCamera::distortPixelFOV / EquidistantCamera::distortPixelEquidistant (src/camera.cpp:247-265,342-363) against a float64 statement
    of the same models: the fixed-point map (1/32 px) must agree to one step (the reference evaluates in float)."""
    W, H = 184, 148
    for model, d in ((1, (0.9, 0, 0, 0, 0)), (2, (-0.01, 0.02, -0.005, 0.001, 0))):
        cam = dict(model=model, width=W, height=H, fx=0.35 * W, fy=0.43 * H, cx=0.49 * W, cy=0.5 * H, d=d)
        m1, m2 = O.init_undistort_maps(cam)
        u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
        ix, iy = (u - cam["cx"]) / cam["fx"], (v - cam["cy"]) / cam["fy"]
        r = np.hypot(ix, iy)
        if model == 1:
            fac = np.where(r > 0, np.arctan(r * 2 * np.tan(d[0] / 2)) / np.maximum(d[0] * r, 1e-300), 1.0)
        else:
            th = np.arctan(r)
            fac = np.where(r > 1e-8, th * (1 + d[0] * th ** 2 + d[1] * th ** 4 + d[2] * th ** 6 + d[3] * th ** 8) / np.maximum(r, 1e-300), 1.0)
        ox, oy = cam["fx"] * fac * ix + cam["cx"], cam["fy"] * fac * iy + cam["cy"]
        gx = m1[:, :, 0].astype(np.float64) + (m2 & 31) / 32.0
        gy = m1[:, :, 1].astype(np.float64) + ((m2 >> 5) & 31) / 32.0
        assert np.abs(gx - ox).max() <= 1.0 / 32 and np.abs(gy - oy).max() <= 1.0 / 32
