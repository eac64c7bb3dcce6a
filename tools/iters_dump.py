import sys, os, json
sys.path.insert(0, '/root/repo')
import numpy as np
import bench
from hso_b200 import Context, make_cam
B, F = 1184, 3000
probs = bench.build_workload(B, F, "icl", 0x450, 0)
c = probs[0]["cam"]
ctx = Context(make_cam(c["width"], c["height"], c["fx"], c["fy"], c["cx"], c["cy"], c["d"]), max_frames=2*B+8)
ref_ids, ref_int, _ = ctx.upload_frames([p["ref_img"] for p in probs])
cur_ids, cur_int, _ = ctx.upload_frames([p["cur_img"] for p in probs])
jobs = [dict(ref=ref_ids[b], cur=cur_ids[b], px=p["px"], f=p["f"], dist=p["dist"], T_cur_ref=p["T0"], exposure_rat=float(np.float32(cur_int[b])/np.float32(ref_int[b]))) for b, p in enumerate(probs)]
res, _ = ctx.coarse_track_batch(jobs)
it = np.array([r["iters_per_level"][:5] for r in res])
np.save('/root/repo/gpurun_out/iters.npy', it)
print(it[:, 1:5].mean(0), it[:, 1:5].std(0))
