// Exercises the C++ host layer (hso_b200/host/hso_b200_host.hpp) end to end on one synthetic problem read from a binary blob
// written by tests/test_gpu_host_cpp.py; prints the results as JSON for the Python side to compare with the ctypes path.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <vector>

#include "../../hso_b200/host/hso_b200_host.hpp"

using namespace hso::b200;

template <class T>
static void rd(std::ifstream& f, T* p, size_t n) { f.read(reinterpret_cast<char*>(p), sizeof(T) * n); }

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  std::ifstream in(argv[1], std::ios::binary);
  int32_t W, H, F;
  double cam[4];
  rd(in, &W, 1); rd(in, &H, 1); rd(in, &F, 1); rd(in, cam, 4);
  std::vector<uint8_t> ref((size_t)W * H), cur((size_t)W * H);
  rd(in, ref.data(), ref.size()); rd(in, cur.data(), cur.size());
  std::vector<double> px(2 * F), f(3 * F), idist(F);
  std::vector<int32_t> has(F);
  rd(in, px.data(), px.size()); rd(in, f.data(), f.size()); rd(in, idist.data(), idist.size()); rd(in, has.data(), has.size());

  hso_cam c;
  std::memset(&c, 0, sizeof c);
  c.model = 0; c.width = W; c.height = H; c.fx = cam[0]; c.fy = cam[1]; c.cx = cam[2]; c.cy = cam[3];
  try {
    Context ctx(c);
    bool threw = false;
    try { Frame bad(ctx, ref.data(), W - 16, H, W - 16, 0.0); } catch (const std::runtime_error&) { threw = true; }
    FramePtr fr(new Frame(ctx, ref.data(), W, H, W, 0.0)), fc(new Frame(ctx, cur.data(), W, H, W, 1.0));
    // the reference frame hosts its own points: hostFeature_ == the feature itself, so dist = |f/idist|
    std::vector<Point> pts(F);
    fr->fts_.resize(F);
    for (int i = 0; i < F; ++i) {
      Feature& ft = fr->fts_[i];
      ft.frame = fr.get(); ft.px[0] = px[2 * i]; ft.px[1] = px[2 * i + 1];
      ft.f[0] = f[3 * i]; ft.f[1] = f[3 * i + 1]; ft.f[2] = f[3 * i + 2];
      if (has[i]) { pts[i].idist_ = idist[i]; pts[i].hostFeature_ = &ft; ft.point = &pts[i]; }
    }
    CoarseTracker tracker(ctx, false, 4, 1, 50, false);
    const size_t n = tracker.run(fr, fc);
    // pose optimiser on the tracked frame: observe the same points from `cur` with bearings predicted by the tracked pose
    fc->fts_.resize(F);
    for (int i = 0; i < F; ++i) {
      Feature& ft = fc->fts_[i];
      ft.frame = fc.get(); ft.level = i % 3; ft.type = (i % 4 == 0) ? Feature::EDGELET : Feature::CORNER;
      ft.point = has[i] ? &pts[i] : nullptr;
      const double inv = 1.0 / idist[i];
      const double ph[3] = {f[3 * i] * inv, f[3 * i + 1] * inv, f[3 * i + 2] * inv};
      double pt[3];
      fc->T_f_w_.apply(ph, pt);
      const double nrm = std::sqrt(pt[0] * pt[0] + pt[1] * pt[1] + pt[2] * pt[2]);
      ft.f[0] = pt[0] / nrm + 1e-3 * ((i * 37) % 11 - 5) / 5.0; ft.f[1] = pt[1] / nrm; ft.f[2] = pt[2] / nrm;
    }
    double scale = 0, e0 = 0, e1 = 0;
    size_t nobs = 0;
    pose_optimizer::optimizeLevenbergMarquardt3rd(ctx, 2.0, 12, false, fc, scale, e0, e1, nobs);
    std::printf("{\"threw\": %d, \"n_tracked\": %zu, \"integral\": [%.9g, %.9g], \"exposure_time\": %.9g, \"T\": [", threw ? 1 : 0, n,
                fr->integralImage_, fc->integralImage_, fc->m_exposure_time);
    for (int k = 0; k < 12; ++k) std::printf("%.17g%s", fc->T_f_w_.m[k], k < 11 ? ", " : "");
    std::printf("], \"T_track\": [");
    for (int k = 0; k < 12; ++k) std::printf("%.17g%s", tracker.last_result().T_cur_ref[k], k < 11 ? ", " : "");
    std::printf("], \"num_obs\": %zu, \"error_final\": %.9g, \"launches\": %llu", nobs, e1, (unsigned long long)hso_kernel_launches(ctx.get()));
    // ---- FrameHandlerMono::addImage front end for 3 streams in one call: must reproduce the single-frame tracker result -----------------
    {
      std::vector<const uint8_t*> imgs(3, cur.data());
      std::vector<FramePtr> refs(3, fr);
      std::vector<SE3> T0(3);
      std::vector<hso_frame_id> ids;
      std::vector<hso_track_result> res;
      addImagesAndTrack(ctx, imgs, W, H, W, refs, T0, false, 4, 1, 50, ids, res);
      std::printf(", \"T_batch\": [");
      for (int k = 0; k < 12; ++k) std::printf("%.17g%s", res[2].T_cur_ref[k], k < 11 ? ", " : "");
      std::printf("], \"batch_iters\": [%d, %d, %d]", res[0].n_iters, res[1].n_iters, res[2].n_iters);
      for (hso_frame_id id : ids) hso_frame_release(ctx.get(), id);
    }
    // ---- Reprojector::reprojectMap: the reference frame is the only keyframe, every point is observed there ----------------------------------
    {
      FramePtr fn(new Frame(ctx, cur.data(), W, H, W, 2.0));
      std::memcpy(fn->T_f_w_.m, tracker.last_result().T_cur_ref, sizeof fn->T_f_w_.m);  // ref pose is the identity
      std::vector<Point*> plist;
      for (int i = 0; i < F; ++i) {
        if (!has[i]) continue;
        Point& p = pts[i];
        const double inv = 1.0 / idist[i];
        for (int k = 0; k < 3; ++k) p.pos_[k] = f[3 * i + k] * inv;
        p.obs_.assign(1, &fr->fts_[i]);
        p.type_ = 1 + i % 4; p.ftr_type_ = i % 3;
        plist.push_back(&p);
      }
      Reprojector rep(ctx, 200);
      unsigned long long state = 12345;
      auto rng = [&]() { state = state * 6364136223846793005ULL + 1442695040888963407ULL; return (unsigned)(state >> 33); };
      rep.resetGrid(rng);
      std::vector<Frame*> kfs;
      rep.reprojectMap(fn, plist, kfs);
      int n_unknown_failed = 0;
      for (Point* p : plist) n_unknown_failed += p->n_failed_reproj_;
      std::printf(", \"reproj\": {\"n_matches\": %zu, \"n_trials\": %zu, \"n_in_frame\": %d, \"new_features\": %zu, \"failed\": %d, \"cell_size\": %d, "
                  "\"cell_order\": [", rep.n_matches_, rep.n_trials_, rep.nFeatures_, fn->fts_.size(), n_unknown_failed, rep.grid_.cell_size);
      for (size_t k = 0; k < rep.grid_.cell_order.size(); ++k) std::printf("%d%s", rep.grid_.cell_order[k], k + 1 < rep.grid_.cell_order.size() ? ", " : "");
      std::printf("], \"first_px\": [%.12g, %.12g]}", fn->fts_.empty() ? 0.0 : fn->fts_[0].px[0], fn->fts_.empty() ? 0.0 : fn->fts_[0].px[1]);
    }
    std::printf("}\n");
  } catch (const std::exception& e) {
    std::fprintf(stderr, "host_smoke: %s\n", e.what());
    return 1;
  }
  return 0;
}
