// CPU-only check of hso::b200::CoarseTracker::makeDepthRef (the host side of src/CoarseTracker.cpp:210-240): reads a scene — a reference frame,
// K host keyframes with their own poses, F features with (has point, host frame, host bearing, inverse depth) — builds the host data model
// (Frame / Feature / Point, pointer-linked like the reference's) on a device-less context and prints the distances, one per line ("%.17g").
#include <cstdio>
#include <vector>

#include "../../hso_b200/host/hso_b200_host.hpp"

using namespace hso::b200;

int main(int argc, char** argv) {
  if (argc < 2) return 2;
  FILE* fp = std::fopen(argv[1], "rb");
  if (!fp) return 2;
  int F = 0, K = 0;
  double Tref[12];
  if (std::fread(&F, 4, 1, fp) != 1 || std::fread(&K, 4, 1, fp) != 1 || std::fread(Tref, 8, 12, fp) != 12) return 2;
  std::vector<double> Th(12 * K), f_host(3 * F), idist(F);
  std::vector<int> has(F), host_of(F);
  if (std::fread(Th.data(), 8, Th.size(), fp) != Th.size() || std::fread(f_host.data(), 8, f_host.size(), fp) != f_host.size() ||
      std::fread(idist.data(), 8, F, fp) != (size_t)F || std::fread(has.data(), 4, F, fp) != (size_t)F ||
      std::fread(host_of.data(), 4, F, fp) != (size_t)F)
    return 2;
  std::fclose(fp);
  hso_cam cam{};
  cam.width = 640; cam.height = 480; cam.fx = cam.fy = 480; cam.cx = 319.5; cam.cy = 239.5;
  Context ctx(cam, Context::HostOnly{});
  Frame ref(ctx, Context::HostOnly{}, 0.0);
  for (int k = 0; k < 12; ++k) ref.T_f_w_.m[k] = Tref[k];
  std::vector<std::unique_ptr<Frame>> hosts;
  std::vector<std::vector<Feature>> host_fts(K);
  for (int k = 0; k < K; ++k) {
    hosts.emplace_back(new Frame(ctx, Context::HostOnly{}, 1.0 + k));
    for (int q = 0; q < 12; ++q) hosts[k]->T_f_w_.m[q] = Th[12 * k + q];
  }
  std::vector<Point> pts(F);
  std::vector<Feature> hf(F);  // the host observation of every point
  ref.fts_.resize(F);
  for (int i = 0; i < F; ++i) {
    hf[i].frame = hosts[host_of[i]].get();
    for (int q = 0; q < 3; ++q) hf[i].f[q] = f_host[3 * i + q];
    pts[i].hostFeature_ = &hf[i];
    pts[i].idist_ = idist[i];
    ref.fts_[i].frame = &ref;
    ref.fts_[i].point = has[i] ? &pts[i] : nullptr;
  }
  std::vector<double> dist;
  CoarseTracker::makeDepthRef(ref, dist);
  for (int i = 0; i < F; ++i) std::printf("%.17g\n", dist[i]);
  return 0;
}
