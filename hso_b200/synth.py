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


def cam2world_plane(c, u, v):
    """Inverse of world2cam() on the unit plane z = 1 for arrays of pixels: pinhole closed form, radtan by fixed-point iteration (to
    convergence, unlike cv::undistortPoints' five steps: this is scene generation), FOV by the tangent formula (src/camera.cpp:169-190)."""
    u, v = np.asarray(u, float), np.asarray(v, float)
    xd, yd = (u - c["cx"]) / c["fx"], (v - c["cy"]) / c["fy"]
    d = c["d"]
    if c.get("model", 0) == 0 and abs(d[0]) > 1e-7:
        x, y = xd.copy(), yd.copy()
        for _ in range(40):
            r2 = x * x + y * y
            cd = 1 + d[0] * r2 + d[1] * r2 * r2 + d[4] * r2 ** 3
            dx = 2 * d[2] * x * y + d[3] * (r2 + 2 * x * x)
            dy = d[2] * (r2 + 2 * y * y) + 2 * d[3] * x * y
            x, y = (xd - dx) / cd, (yd - dy) / cd
        return x, y
    if c.get("model", 0) == 1 and not c.get("undistort", 0):
        r = np.sqrt(xd * xd + yd * yd)
        # the model reaches distorted radii below pi / (2 omega) only (image corners of a wide lens lie outside): clamp there
        ang = np.minimum(r * d[0], 1.5)
        fac = np.where(r > 1e-12, np.tan(ang) / (2 * np.maximum(r, 1e-300) * np.tan(d[0] / 2)), 1.0)
        return fac * xd, fac * yd
    return xd, yd


def warp_plane_cam(img, c, T_cur_ref, depth, gain=1.0):
    """warp_plane() through the camera MODEL (lens distortion on both sides): every current pixel is unprojected with cam2world, intersected
    with the plane z = depth of the reference frame, and projected into the reference image with world2cam — photoconsistent for the radtan /
    FOV cameras, where a pinhole homography is not."""
    H, W = img.shape
    R, t = T_cur_ref[:3, :3], T_cur_ref[:3, 3]
    ys, xs = np.mgrid[0:H, 0:W]
    x, y = cam2world_plane(c, xs.ravel().astype(float), ys.ravel().astype(float))
    ray = np.stack([x, y, np.ones_like(x)])          # current-frame rays
    rr = R.T @ ray                                    # X_ref = R^T (lambda ray - t)
    rt = R.T @ t
    lam = (depth + rt[2]) / rr[2]
    Xr = rr * lam - rt[:, None]
    uv = world2cam(c, Xr.T)
    u, v = uv[:, 0], uv[:, 1]
    u0, v0 = np.floor(u).astype(int), np.floor(v).astype(int)
    su, sv = u - u0, v - v0
    u0c, v0c = np.clip(u0, 0, W - 1), np.clip(v0, 0, H - 1)
    u1c, v1c = np.clip(u0 + 1, 0, W - 1), np.clip(v0 + 1, 0, H - 1)
    f = img.astype(np.float64)
    val = (1 - su) * (1 - sv) * f[v0c, u0c] + su * (1 - sv) * f[v0c, u1c] + (1 - su) * sv * f[v1c, u0c] + su * sv * f[v1c, u1c]
    return np.clip(np.rint(gain * val), 0, 255).astype(np.uint8).reshape(H, W)


def has_lens_model(c):
    return (c.get("model", 0) == 0 and abs(c["d"][0]) > 1e-7) or (c.get("model", 0) == 1 and not c.get("undistort", 0))


def make_pair(seed, cam="icl", F=500, motion_scale=1.0, gain=1.05, depth=4.0, frac_no_point=0.05, border=8):
    """One (ref, cur) problem. Returns dict(ref_img, cur_img, px (F,2), f (F,3), dist (F,), T_true (4x4), gain, cam). For the cameras with a
    lens model the current image and the bearings go through that model, so that T_true is the optimum there too."""
    rng = np.random.default_rng(seed)
    c = CAMS[cam] if isinstance(cam, str) else cam
    W, H = c["width"], c["height"]
    K = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1.0]])
    ref = texture(rng, W, H)
    xi = np.concatenate([rng.normal(0, 0.02, 3), rng.normal(0, 0.004, 3)]) * motion_scale
    T = se3_exp(xi)
    lens = has_lens_model(c)
    cur = warp_plane_cam(ref, c, T, depth, gain) if lens else warp_plane(ref, K, T, depth, gain)
    px = np.stack([rng.uniform(border, W - border, F), rng.uniform(border, H - border, F)], axis=1)
    if lens:
        rx, ry = cam2world_plane(c, px[:, 0], px[:, 1])
        ray = np.stack([rx, ry, np.ones(F)], axis=1)
    else:
        ray = np.stack([(px[:, 0] - c["cx"]) / c["fx"], (px[:, 1] - c["cy"]) / c["fy"], np.ones(F)], axis=1)
    f = ray / np.linalg.norm(ray, axis=1, keepdims=True)
    dist = depth / f[:, 2]
    dist[rng.uniform(size=F) < frac_no_point] = -1.0  # features without a point (Feature::point == NULL)
    if lens and c.get("model", 0) == 1:  # pixels outside the FOV model's domain carry no point
        rd = np.hypot((px[:, 0] - c["cx"]) / c["fx"], (px[:, 1] - c["cy"]) / c["fy"])
        dist[rd * c["d"][0] > 1.45] = -1.0
    return dict(ref_img=ref, cur_img=cur, px=px, f=f, dist=dist, T_true=T, gain=gain, cam=c)


def homography(K, T_cur_ref, depth):
    R, t = T_cur_ref[:3, :3], T_cur_ref[:3, 3]
    return K @ (R + np.outer(t, np.array([0, 0, 1.0])) / depth) @ np.linalg.inv(K)


def make_align_jobs(seed, pair, M=500, frac_edgelet=0.3, noise_px=1.5, depth=4.0):
    """Candidate matches for Matcher::findMatchDirect's inner part on a make_pair() scene: the affine warp is the local Jacobian of
    the plane homography (what warp::getWarpMatrixAffine measures by finite differences, src/matcher.cpp:46-72)."""
    rng = np.random.default_rng(seed)
    c = pair["cam"]
    W, H = c["width"], c["height"]
    Kc = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1.0]])
    Hm = homography(Kc, pair["T_true"], depth)

    def warp(p):
        q = Hm @ np.array([p[0], p[1], 1.0])
        return q[:2] / q[2]

    jobs = []
    for m in range(M):
        lvl = int(rng.integers(0, 3))
        px_ref = np.array([rng.uniform(40, W - 40), rng.uniform(40, H - 40)])
        s = 5.0 * (1 << lvl)
        pc = warp(px_ref)
        A = np.stack([(warp(px_ref + [s, 0]) - pc) / 5.0, (warp(px_ref + [0, s]) - pc) / 5.0], axis=1)
        if m % 7 == 3:
            A = A * 2.3  # exercise search levels > 0
        D, sl = np.linalg.det(A), 0
        while D > 3.0 and sl < 2:
            sl += 1
            D *= 0.25
        ang = rng.uniform(0, 2 * np.pi)
        jobs.append(dict(ref_level=lvl, search_level=sl, type=1 if rng.uniform() < frac_edgelet else 0, scale_patch=int(m % 5 == 0),
                         px_ref=px_ref, A_cur_ref=A, grad=np.array([np.cos(ang), np.sin(ang)]), px_cur=pc + rng.normal(0, noise_px, 2),
                         exposure_rat=float(pair["gain"])))
    return jobs


def make_pose_problem(seed, cam="icl", F=400, K=8, noise_px=0.5, frac_outlier=0.1, frac_edgelet=0.3, pose_err=1.0):
    """Inputs of pose_optimizer::optimizeLevenbergMarquardt3rd: F observed bearings of points hosted in K keyframes."""
    rng = np.random.default_rng(seed)
    c = CAMS[cam] if isinstance(cam, str) else cam
    fbar = abs((c["fx"] + c["fy"]) * 0.5)
    T_fw = se3_exp(np.concatenate([rng.normal(0, 0.2, 3), rng.normal(0, 0.05, 3)]))
    T_hosts = [se3_exp(np.concatenate([rng.normal(0, 0.3, 3), rng.normal(0, 0.08, 3)])) for _ in range(K)]
    uv = np.stack([rng.uniform(-0.6, 0.6, F), rng.uniform(-0.45, 0.45, F)], axis=1)
    z = rng.uniform(1.0, 8.0, F)
    Pf = np.stack([uv[:, 0] * z, uv[:, 1] * z, z], axis=1)
    Pw = (np.linalg.inv(T_fw) @ np.concatenate([Pf, np.ones((F, 1))], axis=1).T).T
    host_idx = rng.integers(0, K, F).astype(np.int32)
    p_host = np.stack([(T_hosts[h] @ Pw[i])[:3] for i, h in enumerate(host_idx)]) if F else np.zeros((0, 3))
    noise = rng.normal(0, noise_px / fbar, (F, 2))
    out = rng.uniform(size=F) < frac_outlier
    noise[out] += rng.normal(0, 10.0 / fbar, (int(out.sum()), 2))
    fobs = np.stack([uv[:, 0] + noise[:, 0], uv[:, 1] + noise[:, 1], np.ones(F)], axis=1)
    fobs /= np.linalg.norm(fobs, axis=1, keepdims=True)
    ang = rng.uniform(0, 2 * np.pi, F)
    T0 = se3_exp(np.concatenate([rng.normal(0, 0.01, 3), rng.normal(0, 0.003, 3)]) * pose_err) @ T_fw
    return dict(f=fobs, p_host=p_host, host_idx=host_idx, T_host_w=np.stack([T[:3] for T in T_hosts]),
                grad=np.stack([np.cos(ang), np.sin(ang)], axis=1), level=rng.integers(0, 3, F).astype(np.int8),
                ftype=(rng.uniform(size=F) < frac_edgelet).astype(np.int8),
                ptype=np.where(rng.uniform(size=F) < 0.1, 1, 4).astype(np.int8), T_f_w=T0[:3], T_true=T_fw[:3], n_fts_total=F + 7,
                err_mult2=fbar)


def world2cam(c, P):
    """AbstractCamera::world2cam of the three models (src/camera.cpp:94-125,199-221,307-315) for an (N,3) array."""
    P = np.asarray(P, float).reshape(-1, 3)
    u, v = P[:, 0] / P[:, 2], P[:, 1] / P[:, 2]
    d = c["d"]
    if c.get("model", 0) == 0 and abs(d[0]) > 1e-7:
        r2 = u * u + v * v
        cd = 1 + d[0] * r2 + d[1] * r2 * r2 + d[4] * r2 ** 3
        a1, a2, a3 = 2 * u * v, r2 + 2 * u * u, r2 + 2 * v * v
        u, v = u * cd + d[2] * a1 + d[3] * a2, v * cd + d[2] * a3 + d[3] * a1
    elif c.get("model", 0) == 1 and not c.get("undistort", 0):
        dist = np.sqrt(u * u + v * v)
        ratio = np.where(dist > 0, np.arctan(2 * dist * np.tan(d[0] / 2)) / np.maximum(dist * d[0], 1e-300), 1.0)
        u, v = ratio * u, ratio * v
    return np.stack([c["fx"] * u + c["cx"], c["fy"] * v + c["cy"]], axis=1)


def make_reproject_scene(seed, cam="icl", M=3000, n_kf=4, max_fts=200, depth=4.0, gain=1.05, frac_edgelet=0.25):
    """Inputs of Reprojector::reprojectMap's data path (row N1): n_kf keyframes and a current frame looking at a textured plane
    z = depth of keyframe 0 (= world), M map points on the plane, each hosted in one keyframe and observed (ref_ftr_) in another.
    Returns dict(cam, kf_imgs, cur_img, T_f_w (n_kf,3,4), T_cur_w (3,4), cands [dict], grid dict, cell_order)."""
    rng = np.random.default_rng(seed)
    c = CAMS[cam] if isinstance(cam, str) else cam
    W, H = c["width"], c["height"]
    K = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1.0]])
    base = texture(rng, W, H)
    T_kf = [np.eye(4)] + [se3_exp(np.concatenate([rng.normal(0, 0.05, 3), rng.normal(0, 0.01, 3)])) for _ in range(n_kf - 1)]
    kf_imgs = [base] + [(warp_plane_cam(base, c, T, depth) if has_lens_model(c) else warp_plane(base, K, T, depth)) for T in T_kf[1:]]
    T_cur = se3_exp(np.concatenate([rng.normal(0, 0.04, 3), rng.normal(0, 0.008, 3)]))
    cur_img = warp_plane_cam(base, c, T_cur, depth, gain) if has_lens_model(c) else warp_plane(base, K, T_cur, depth, gain)
    # points on the plane, spread a little beyond the field of view so that some fall outside the current image
    x = rng.uniform(-1.12, 1.12, M) * depth * (W / 2) / c["fx"]
    y = rng.uniform(-1.12, 1.12, M) * depth * (H / 2) / c["fy"]
    Pw = np.stack([x, y, np.full(M, depth)], axis=1)
    cands = []
    for i in range(M):
        h, r = int(rng.integers(0, n_kf)), int(rng.integers(0, n_kf))
        if i % 3 == 0:
            r = h
        Ph = T_kf[h][:3, :3] @ Pw[i] + T_kf[h][:3, 3]
        Pr = T_kf[r][:3, :3] @ Pw[i] + T_kf[r][:3, 3]
        px_ref = world2cam(c, Pr)[0]
        ang = rng.uniform(0, 2 * np.pi)
        ref_pose = r if rng.uniform() > 0.03 else -1  # getCloseViewObs fails for a few
        cands.append(dict(p_host=Ph, px_ref=px_ref, f_ref=Pr / np.linalg.norm(Pr), grad=np.array([np.cos(ang), np.sin(ang)]),
                          depth_ref=float(np.linalg.norm(Pr)), host_pose=h, ref_pose=ref_pose, ref_frame=r, ref_level=int(rng.integers(0, 3)),
                          ftr_type=(1 if rng.uniform() < frac_edgelet else (2 if rng.uniform() < 0.2 else 0)),
                          pt_type=int(rng.choice([0, 1, 2, 3, 4], p=[0.02, 0.13, 0.25, 0.3, 0.3])), pt_ftr_type=int(rng.integers(0, 3)),
                          scale_patch=int(i % 5 == 0), exposure_rat=float(gain)))
    cell_size = int(np.floor(np.float32(np.sqrt(np.float32(W * H) / max_fts)) * np.float32(0.6)))  # Reprojector::caculateGridSize
    n_cols, n_rows = int(np.ceil(W / cell_size)), int(np.ceil(H / cell_size))
    grid = dict(cell_size=cell_size, n_cols=n_cols, n_rows=n_rows, max_fts=max_fts, align_max_iter=10)
    cell_order = rng.permutation(n_cols * n_rows).astype(np.int32)  # std::random_shuffle(grid_.cell_order)
    return dict(cam=c, kf_imgs=kf_imgs, cur_img=cur_img, T_f_w=np.stack([T[:3] for T in T_kf]), T_cur_w=T_cur[:3], cands=cands, grid=grid,
                cell_order=cell_order)


def make_depth_scene(seed, cam="icl", S=2000, n_kf=3, depth=4.0, gain=1.0, frac_edgelet=0.25, baseline=0.12):
    """Inputs of DepthFilter::observeDepthRow (row N3): n_kf keyframes and an active frame looking at a textured plane z = depth of
    keyframe 0 (= world); S seeds, each a feature of one keyframe with an inverse-depth estimate (mu, sigma2) around the truth.
    Returns dict(cam, kf_imgs, cur_img, T_f_w, T_cur_w, seeds [dict], px_error_angle)."""
    rng = np.random.default_rng(seed)
    c = CAMS[cam] if isinstance(cam, str) else cam
    W, H = c["width"], c["height"]
    K = np.array([[c["fx"], 0, c["cx"]], [0, c["fy"], c["cy"]], [0, 0, 1.0]])
    base = texture(rng, W, H)
    T_kf = [np.eye(4)] + [se3_exp(np.concatenate([rng.normal(0, 0.05, 3), rng.normal(0, 0.01, 3)])) for _ in range(n_kf - 1)]
    kf_imgs = [base] + [(warp_plane_cam(base, c, T, depth) if has_lens_model(c) else warp_plane(base, K, T, depth)) for T in T_kf[1:]]
    T_cur = se3_exp(np.concatenate([rng.normal(0, baseline, 3) * [1, 1, 0.3], rng.normal(0, 0.01, 3)]))
    cur_img = warp_plane_cam(base, c, T_cur, depth, gain) if has_lens_model(c) else warp_plane(base, K, T_cur, depth, gain)
    seeds = []
    for i in range(S):
        r = int(rng.integers(0, n_kf))
        # a pixel of keyframe r, its bearing and the true distance to the plane along it
        px = np.array([rng.uniform(20, W - 20), rng.uniform(20, H - 20)])
        ray = np.array([(px[0] - c["cx"]) / c["fx"], (px[1] - c["cy"]) / c["fy"], 1.0])
        f = ray / np.linalg.norm(ray)
        Tinv = np.linalg.inv(T_kf[r])
        o, d = Tinv[:3, 3], Tinv[:3, :3] @ f          # ray in the world frame
        lam = (depth - o[2]) / d[2]                    # intersection with z = depth
        mu_true = 1.0 / lam
        rel = rng.choice([0.03, 0.1, 0.3, 0.8])        # width of the search interval relative to mu
        sigma = rel * mu_true / 2.0
        mu = mu_true + rng.normal(0, 0.5 * sigma)
        if i % 97 == 0:
            mu = -abs(mu)                               # behind the camera
        ang = rng.uniform(0, 2 * np.pi)
        lvl = int(rng.integers(0, 3))
        seeds.append(dict(px=px, f=f, grad=np.array([np.cos(ang), np.sin(ang)]), ref_frame=r, ref_pose=r, level=lvl,
                          ftr_type=(1 if rng.uniform() < frac_edgelet else (2 if rng.uniform() < 0.2 else 0)),
                          mu=float(np.float32(mu)), sigma2=float(np.float32(sigma * sigma)), exposure_rat=float(gain)))
    focal = abs((c["fx"] + c["fy"]) * 0.5)
    return dict(cam=c, kf_imgs=kf_imgs, cur_img=cur_img, T_f_w=np.stack([T[:3] for T in T_kf]), T_cur_w=T_cur[:3], seeds=seeds,
                px_error_angle=float(np.arctan(1.0 / (2.0 * focal)) * 2.0))


def make_seed_reproject_scene(seed, cam="icl", S=600, max_fts=200, gain=1.0, **kw):
    """Inputs of the seed stage of Reprojector::reprojectMap (row a13b, src/reprojector.cpp:309-328): the keyframes, active frame and seeds of
    make_depth_scene() — depth estimates near the truth with a spread of variances, a few behind the camera — plus the reprojection grid."""
    sc = make_depth_scene(seed, cam, S=S, gain=gain, **kw)
    c = sc["cam"]
    rng = np.random.default_rng(seed + 99991)
    W, H = c["width"], c["height"]
    cell_size = int(np.floor(np.float32(np.sqrt(np.float32(W * H) / max_fts)) * np.float32(0.6)))
    n_cols, n_rows = int(np.ceil(W / cell_size)), int(np.ceil(H / cell_size))
    sc["grid"] = dict(cell_size=cell_size, n_cols=n_cols, n_rows=n_rows, max_fts=max_fts, align_max_iter=10)
    sc["cell_order"] = rng.permutation(n_cols * n_rows).astype(np.int32)
    for i, sd in enumerate(sc["seeds"]):
        if i % 9 == 0:  # ties in sigma2 exercise the stable sort
            sd["sigma2"] = float(np.float32(0.004))
    return sc
