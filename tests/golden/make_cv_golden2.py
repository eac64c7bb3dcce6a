"""Generates tests/golden/cv_golden2.npz with OpenCV (cv2 4.13 in the build container): third-party arithmetic of the NEXT rows
 that is not vendored in /root/reference:
  * cv::undistortPoints as PinholeCamera::cam2world calls it (src/camera.cpp:66-87: one CV_32FC2 point, float cvK_ / cvD_ built at
    camera.cpp:43-45, no R/P, default criteria) — row N1 (warp::getWarpMatrixAffine);
  * cv::initUndistortRectifyMap(CV_16SC2) + cv::remap(INTER_LINEAR) as PinholeCamera::undistortImage does (camera.cpp:47-54,127-131) — row N4.
Run:  python tests/golden/make_cv_golden2.py   (needs cv2; the committed .npz is what the tests read)."""
import os

import cv2
import numpy as np

rng = np.random.default_rng(20261018)
out = {}
# EuRoC intrinsics (test/cameras/euroc.txt), as float32 like cvK_/cvD_
K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1.0]])
d = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])
Kf, df = K.astype(np.float32), d.astype(np.float32).reshape(1, 5)
uv = np.stack([rng.uniform(0, 752, 256), rng.uniform(0, 480, 256)], axis=1)
und = np.zeros((256, 2), np.float32)
for i in range(256):
    src = np.array([[uv[i]]], np.float32)  # cv::Point2f uv(u, v)
    und[i] = cv2.undistortPoints(src, Kf, df).reshape(2)
out["undist_K"], out["undist_d"], out["undist_uv"], out["undist_xy"] = K, d, uv, und
# undistortion remap at a reduced size (keeps the fixture small): same K scaled to 188x120, same distortion
s = 0.25
K4 = np.array([[458.654 * s, 0, 367.215 * s], [0, 457.296 * s, 248.375 * s], [0, 0, 1.0]])
K4f = K4.astype(np.float32)
W, H = 188, 120
m1, m2 = cv2.initUndistortRectifyMap(K4f, df, np.eye(3), K4f, (W, H), cv2.CV_16SC2)
img = np.clip(rng.normal(120, 50, (H, W)) + 50 * np.sin(np.arange(W)[None, :] / 4.0) + 30 * np.cos(np.arange(H)[:, None] / 6.0), 0, 255).astype(np.uint8)
out["remap_K"], out["remap_d"], out["remap_size"] = K4, d, np.array([W, H])
out["remap_map1"], out["remap_map2"] = m1, m2
out["remap_src"] = img
out["remap_dst"] = cv2.remap(img, m1, m2, cv2.INTER_LINEAR)
# a stronger distortion so that a visible band maps outside the source (BORDER_CONSTANT 0)
d2 = np.array([-0.45, 0.18, 0.001, -0.0005, 0.0])
d2f = d2.astype(np.float32).reshape(1, 5)
m1b, m2b = cv2.initUndistortRectifyMap(K4f, d2f, np.eye(3), K4f, (W, H), cv2.CV_16SC2)
out["remap2_d"] = d2
out["remap2_map1"], out["remap2_map2"] = m1b, m2b
out["remap2_dst"] = cv2.remap(img, m1b, m2b, cv2.INTER_LINEAR)
# cv::convertMaps(float, float -> CV_16SC2) as FOVCamera/EquidistantCamera::getRemap call it (camera.cpp:244,339), on maps that leave the
# source on every side, and the remap through them (BORDER_CONSTANT branches)
mx = (np.arange(W, dtype=np.float32)[None, :] * np.float32(1.13) - np.float32(9.3) + rng.normal(0, 0.7, (H, W)).astype(np.float32)).astype(np.float32)
my = (np.arange(H, dtype=np.float32)[:, None] * np.float32(1.17) - np.float32(7.1) + rng.normal(0, 0.7, (H, W)).astype(np.float32)).astype(np.float32)
c1, c2 = cv2.convertMaps(mx, my, cv2.CV_16SC2)
out["conv_mapx"], out["conv_mapy"], out["conv_map1"], out["conv_map2"] = mx, my, c1, c2
out["conv_dst"] = cv2.remap(img, c1, c2, cv2.INTER_LINEAR)
# ImageReader-style resize (src/ImageReader.cpp:80 cv::resize to a target size, INTER_LINEAR default): 1280x1024 -> 920x736 scaled by 1/8
src = np.clip(rng.normal(110, 60, (128, 160)), 0, 255).astype(np.uint8)
out["reader_src"] = src
out["reader_dst"] = cv2.resize(src, (115, 92))
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv_golden2.npz"), **out)
print("wrote cv_golden2.npz", {k: v.shape for k, v in out.items()})
