#!/usr/bin/env python
"""Small invocation of every kernel of libhso_b200.so for compute-sanitizer (memcheck / racecheck / synccheck run it under the tool; see
tools/sanitize.sh). Sizes are small because the tools slow kernels down 10-100x, but every launch shape class is visited: tracker single CTA,
clusters of 2 / 8 (DSMEM reductions, list exchange), inverse-compositional dual-image and cached paths, the global-memory path (level 0), the
threshold-selection fallback, the chunk-pipelined entry on several streams; pyramid in its three size classes; align / reproject / seeds /
depth / pose / FAST / raw input."""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from hso_b200 import Context, make_cam, synth  # noqa: E402


def ctx_for(c, **kw):
    return Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"], c.get("model", 0)), **kw)


def main():
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "track"):
        p = synth.make_pair(3, "icl", F=700)
        ctx = ctx_for(p["cam"], max_frames=64)
        ids, integ, _ = ctx.upload_frames([p["ref_img"], p["cur_img"], p["ref_img"]])
        a0 = float(np.float32(integ[1]) / np.float32(integ[0]))
        job = dict(ref=ids[0], cur=ids[1], px=p["px"], f=p["f"], dist=p["dist"], T_cur_ref=np.eye(4)[:3], exposure_rat=a0)
        for ic in (False, True):
            for shape in ((0, 0), (1, 256), (2, 256), (8, 64)):
                ctx.set_cluster(*shape)
                ctx.coarse_track_batch([job] * 3, inverse_comp=ic, trace_cap=64)
            ctx.set_cluster(0, 0)
            ctx.coarse_track_batch([job], inverse_comp=ic, min_level=0, n_iter=15)  # level 0: global-memory path
        ctx._chk(ctx.lib.hso_track_set_ic_dual(ctx.h, 0))
        ctx.coarse_track_batch([job] * 2, inverse_comp=True)
        ctx._chk(ctx.lib.hso_track_set_ic_dual(ctx.h, 1))
        same = dict(job, cur=ids[2], exposure_rat=1.0)  # identical images: every |r| = 0 -> selection fallback
        ctx.coarse_track_batch([same], inverse_comp=False)
        big = synth.make_pair(4, "icl", F=3000)  # the benchmark's patch count: cluster of 2 at level 1, |r| scratch in global memory
        ids2, integ2, _ = ctx.upload_frames([big["ref_img"], big["cur_img"]])
        jb = dict(ref=ids2[0], cur=ids2[1], px=big["px"], f=big["f"], dist=big["dist"], T_cur_ref=np.eye(4)[:3], exposure_rat=1.0)
        ctx.set_level_shape(1, 2, 512)
        ctx.coarse_track_batch([jb] * 2)
        ctx.set_level_shape(1, 0, 0)
        # the benchmark's launch shapes (two problems per SM): 256-thread CTA pairs with the streamed reference-patch cache (mode 3: per-warp TMA
        # ring + mbarriers) at levels 4..2, one 512-thread CTA at level 1 in mode 3 (forward) / mode 4 (inverse-compositional, single-buffered ring)
        for ic in (False, True):
            ctx.coarse_track_batch([jb] * 296, inverse_comp=ic)
            assert ctx.level_shape(4)[:3] == (1, 256, 3) and ctx.level_shape(1)[:3] == (1, 512, 4 if ic else 3), [ctx.level_shape(l) for l in (4, 3, 2, 1)]
        # streamed cache forced at every level, single CTA and clusters (the ring in a cluster launch), both modes
        ctx._chk(ctx.lib.hso_track_set_stream_cache(ctx.h, 1))
        for ic in (False, True):
            for shape in ((0, 0), (1, 256), (2, 256)):
                ctx.set_cluster(*shape)
                ctx.coarse_track_batch([job] * 3, inverse_comp=ic, trace_cap=64)
        ctx.set_cluster(0, 0)
        ctx._chk(ctx.lib.hso_track_set_stream_cache(ctx.h, 0))
        # compact feature layout (xyz + float32 px): the copy-as-it-is path and its branch of k_track_compact, both entry points
        xyz, px32 = Context.compact_features(p["px"], p["f"], p["dist"])
        cj = dict(ref=ids[0], cur=ids[1], xyz=xyz, px32=px32, T_cur_ref=np.eye(4)[:3], exposure_rat=a0)
        ctx.coarse_track_batch([cj] * 3)
        ctx.add_frames_track_batch([p["cur_img"]] * 5, [dict(ref=ids[0], xyz=xyz, px32=px32, T_cur_ref=np.eye(4)[:3])] * 5)
        ctx._chk(ctx.lib.hso_set_pipeline(ctx.h, 4, 3))  # 10 problems -> 3 chunks on 3 streams
        ctx.add_frames_track_batch([p["cur_img"]] * 10, [dict(ref=ids[0], px=p["px"], f=p["f"], dist=p["dist"], T_cur_ref=np.eye(4)[:3])] * 10)
        ctx.close()
        print("track ok")
    if which in ("all", "frame"):
        for cam in ("icl", "euroc", "tum_fov"):
            c = synth.CAMS[cam]
            ctx = ctx_for(c, materialize_sobel=True)
            img = synth.texture(np.random.default_rng(1), c["width"], c["height"])
            ids, _, _ = ctx.upload_frames([img, img])
            for l in range(3):
                ctx.fast_detect(ids[0], l, 20) if hasattr(ctx, "fast_detect") else None
            ctx.close()
        cf = synth.CAMS["tum_fov"]
        ctx = ctx_for(cf, max_frames=8)
        raw = synth.texture(np.random.default_rng(2), 1280, 1024)
        ctx.upload_raw_frames([raw, raw], undistort=True)
        ctx.close()
        print("frame ok")
    if which in ("all", "match"):
        sc = synth.make_reproject_scene(2, "icl", M=500)
        ctx = ctx_for(sc["cam"], materialize_sobel=True)
        kf, _, _ = ctx.upload_frames(sc["kf_imgs"])
        cur = ctx.upload_frames([sc["cur_img"]])[0][0]
        ctx.reproject_match(cur, sc["T_cur_w"], sc["T_f_w"], Context.reproj_cands(sc["cands"], frame_ids=kf), sc["grid"], sc["cell_order"])
        ctx.reproject_match(cur, sc["T_cur_w"], sc["T_f_w"], Context.reproj_cands(sc["cands"][:150], frame_ids=kf), sc["grid"], sc["cell_order"])
        ss = synth.make_seed_reproject_scene(3, "icl", S=400)
        kf2, _, _ = ctx.upload_frames(ss["kf_imgs"])
        cur2 = ctx.upload_frames([ss["cur_img"]])[0][0]
        ctx.reproject_seeds(cur2, ss["T_cur_w"], ss["T_f_w"], Context.seed_obs(ss["seeds"], frame_ids=kf2), ss["grid"], ss["cell_order"], n_matches_in=10)
        ctx.depth_observe(cur2, ss["T_cur_w"], ss["T_f_w"], Context.seed_obs(ss["seeds"], frame_ids=kf2), ss["px_error_angle"])
        probs = [synth.make_pose_problem(5 + i, "icl", F=300 + 200 * i, K=4) for i in range(3)]
        ctx.pose_optimize_batch(probs)
        ctx.close()
        print("match ok")


if __name__ == "__main__":
    main()
