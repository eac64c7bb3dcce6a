"""Generates tests/golden/cv_golden.npz with OpenCV (cv2 4.13 in the build container): the third-party arithmetic the reference
calls on this path and that is NOT vendored in /root/reference: cv::Sobel(CV_16S, k=5, BORDER_REPLICATE)
(src/frame.cpp:218-219), cv::resize(INTER_LINEAR) on the non-%16 pyramid path (src/frame.cpp:309-311), and the radtan
projection of cv::projectPoints as an independent statement of PinholeCamera::world2cam (src/camera.cpp:99-125).
Run:  python tests/golden/make_cv_golden.py   (needs cv2; the committed .npz is what the tests read)."""
import os

import cv2
import numpy as np

rng = np.random.default_rng(20261017)
out = {}
# small images with structure + noise; odd sizes exercise the borders
for name, (h, w) in {"a": (37, 53), "b": (64, 80)}.items():
    img = np.clip(rng.normal(128, 50, (h, w)) + 40 * np.sin(np.arange(w)[None, :] / 3.0) + 30 * np.cos(np.arange(h)[:, None] / 5.0), 0, 255).astype(np.uint8)
    out[f"sobel_img_{name}"] = img
    out[f"sobel_gx_{name}"] = cv2.Sobel(img, cv2.CV_16S, 1, 0, ksize=5, scale=1, delta=0, borderType=cv2.BORDER_REPLICATE)
    out[f"sobel_gy_{name}"] = cv2.Sobel(img, cv2.CV_16S, 0, 1, ksize=5, scale=1, delta=0, borderType=cv2.BORDER_REPLICATE)
# resize chain of the TUM-sized pyramid (920x736 -> 460x368 -> 230x184 -> 115x92 -> 58x46), scaled down 4x to keep the fixture small:
# 230x184 -> 115x92 (exact 2x) -> 58x46 (non-integer ratio) -> 29x23 ; plus an odd pair
src = np.clip(rng.normal(120, 60, (184, 230)), 0, 255).astype(np.uint8)
out["resize_src"] = src
prev = src
for i, (dw, dh) in enumerate([(115, 92), (58, 46), (29, 23)]):
    prev = cv2.resize(prev, (dw, dh), interpolation=cv2.INTER_LINEAR)
    out[f"resize_l{i + 1}"] = prev
odd = np.clip(rng.normal(100, 70, (45, 61)), 0, 255).astype(np.uint8)
out["resize_odd_src"] = odd
out["resize_odd_dst"] = cv2.resize(odd, (23, 31), interpolation=cv2.INTER_LINEAR)
# radtan projection (EuRoC intrinsics, test/cameras/euroc.txt)
K = np.array([[458.654, 0, 367.215], [0, 457.296, 248.375], [0, 0, 1.0]])
d = np.array([-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0])
P = np.stack([rng.uniform(-2, 2, 64), rng.uniform(-1.5, 1.5, 64), rng.uniform(1.0, 8.0, 64)], axis=1)
px, _ = cv2.projectPoints(P.reshape(-1, 1, 3), np.zeros(3), np.zeros(3), K, d)
out["proj_xyz"], out["proj_px"], out["proj_K"], out["proj_d"] = P, px.reshape(-1, 2), K, d
np.savez_compressed(os.path.join(os.path.dirname(os.path.abspath(__file__)), "cv_golden.npz"), **out)
print("wrote cv_golden.npz", {k: v.shape for k, v in out.items()})
