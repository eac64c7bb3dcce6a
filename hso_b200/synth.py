"""Seeded synthetic inputs for the tracking hot path: textured u8 images, a planar scene seen from two
poses, reference features with depths. Pure numpy; used by tests/ and bench.py (no dataset is available offline)."""
import numpy as np

CAMS = {
    # test/cameras/icl-nuim.txt-like pinhole at 640x480
    "icl": dict(width=640, height=480, fx=481.2, fy=480.0, cx=319.5, cy=239.5, d=(0, 0, 0, 0, 0), model=0),
    # test/cameras/euroc.txt: 752x480 pinhole + radtan
    "euroc": dict(width=752, height=480, fx=458.654, fy=457.296, cx=367.215, cy=248.375,
                  d=(-0.28340811, 0.07395907, 0.00019359, 1.76187114e-05, 0.0), model=0),
    # TUM monoVO wide after the driver's resize (920x736), FOV model
    "tum_fov": dict(width=920, height=736, fx=0.349153 * 920, fy=0.436593 * 736, cx=0.493140 * 920 - 0.5, cy=0.499021 * 736 - 0.5,
                    d=(0.933271, 0, 0, 0, 0), model=1),
}


def texture(rng, W, H, contrast=40.0):
    """Band-limited random texture: white noise low-passed at several scales (FFT), mean 128, clipped to u8."""
    fy = np.fft.fftfreq(H)[:, None]
    fx = np.fft.rfftfreq(W)[None, :]
    r2 = fx * fx + fy * fy
    img = np.zeros((H, W))
    for sigma, amp in ((1.5, 0.6), (3.0, 0.8), (6.0, 1.0), (12.0, 1.0)):
        n = rng.standard_normal((H, W))
        g = np.exp(-2 * (np.pi ** 2) * (sigma ** 2) * r2)
        f = np.fft.irfft2(np.fft.rfft2(n) * g, s=(H, W))
        img += amp * f / f.std()
    img = 128.0 + contrast * img / img.std()
    return np.clip(np.rint(img), 0, 255).astype(np.uint8)


def se3_exp(xi):
    """exp([v, w]) as 4x4 (Rodrigues)."""
    v, w = np.asarray(xi[:3], float), np.asarray(xi[3:], float)
    th = np.linalg.norm(w)
    Wx = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-12:
        R, V = np.eye(3) + Wx, np.eye(3) + 0.5 * Wx
    else:
        A, Bc, Cc = np.sin(th) / th, (1 - np.cos(th)) / th ** 2, (th - np.sin(th)) / th ** 3
        R = np.eye(3) + A * Wx + Bc * Wx @ Wx
        V = np.eye(3) + Bc * Wx + Cc * Wx @ Wx
    T = np.eye(4)
    T[:3, :3], T[:3, 3] = R, V @ v
    return T


def warp_plane(img, K, T_cur_ref, depth, gain=1.0):
    """Render the current view of a fronto-parallel textured plane z=depth (reference frame) by inverse homography + bilinear."""
    H, W = img.shape
    R, t = T_cur_ref[:3, :3], T_cur_ref[:3, 3]
    n = np.array([0, 0, 1.0])
    Hm = K @ (R + np.outer(t, n) / depth) @ np.linalg.inv(K)  # ref px -> cur px
    Hi = np.linalg.inv(Hm)
    ys, xs = np.mgrid[0:H, 0:W]
    p = Hi @ np.stack([xs.ravel(), ys.ravel(), np.ones(H * W)])
    u, v = p[0] / p[2], p[1] / p[2]
    u0, v0 = np.floor(u).astype(int), np.floor(v).astype(int)
    su, sv = u - u0, v - v0
    u0c, v0c = np.clip(u0, 0, W - 1), np.clip(v0, 0, H - 1)
    u1c, v1c = np.clip(u0 + 1, 0, W - 1), np.clip(v0 + 1, 0, H - 1)
    f = img.astype(np.float64)
    val = (1 - su) * (1 - sv) * f[v0c, u0c] + su * (1 - sv) * f[v0c, u1c] + (1 - su) * sv * f[v1c, u0c] + su * sv * f[v1c, u1c]
    return np.clip(np.rint(gain * val), 0, 255).astype(np.uint8).reshape(H, W)


def make_pair(seed, cam="icl", F=500, motion_scale=1.0, gain=1.05, depth=4.0, frac_no_point=0.05, border=8):
    """One (ref, cur) problem. Returns dict(ref_img, cur_img, px (F,2), f (F,3), dist (F,), T_true (4x4), gain, cam)."""
    rng = np.random.default_rng(seed)
    c = CAMS[cam] if isinstance(cam, str) else cam
    W, H = c["width"], c["height"]
    K = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1.0]])
    ref = texture(rng, W, H)
    xi = np.concatenate([rng.normal(0, 0.02, 3), rng.normal(0, 0.004, 3)]) * motion_scale
    T = se3_exp(xi)
    cur = warp_plane(ref, K, T, depth, gain)
    px = np.stack([rng.uniform(border, W - border, F), rng.uniform(border, H - border, F)], axis=1)
    ray = np.stack([(px[:, 0] - c["cx"]) / c["fx"], (px[:, 1] - c["cy"]) / c["fy"], np.ones(F)], axis=1)
    f = ray / np.linalg.norm(ray, axis=1, keepdims=True)
    dist = depth / f[:, 2]
    dist[rng.uniform(size=F) < frac_no_point] = -1.0  # features without a point (Feature::point == NULL)
    return dict(ref_img=ref, cur_img=cur, px=px, f=f, dist=dist, T_true=T, gain=gain, cam=c)
