// CPU-only check of the host layer's wire/disk formats (row N4): timestamp line parsing and the TUM trajectory line. No GPU calls are made
// (the header's device-calling classes are only declared); links against libhso_b200.so for the symbols the inline classes reference.
#include <cmath>
#include <iostream>
#include <sstream>

#include "../../hso_b200/host/hso_b200_host.hpp"

using namespace hso::b200;

int main() {
  std::string s;
  int fails = 0;
  auto expect = [&](const char* line, bool ok, const char* want) {
    std::string got;
    const bool r = parseTimestampLine(line, got);
    if (r != ok || (ok && got != want)) { std::cerr << "parse failed: '" << line << "' -> " << r << " '" << got << "'\n"; ++fails; }
  };
  expect("1403636579.763555 4.68 -1.78 0.81 0.53 -0.15 -0.82 0.10", true, "1403636579.763555");  // stamp + pose
  expect("00042 1465821313.5274 0.01956", true, "1465821313.5274");                                // id stamp exposure (TUM monoVO)
  expect("17 1403636580.013555", true, "1403636580.013555");                                       // id stamp
  // quirk of the reference's cascade: a bare numeric stamp is consumed by "%d %s" first (id = integer part, stamp = the rest)
  expect("1403636580.263555", true, ".263555");
  expect("frame_000123", true, "frame_000123");                                                    // stamp (does not start with an integer)
  expect("", false, "");
  // trajectory line: T_f_w = [Rz(90 deg) | t] -> T_w_f translation = -R^T t, quaternion of R^T
  SE3 T;
  const double c = 0.0, sn = 1.0;
  const double R[9] = {c, -sn, 0, sn, c, 0, 0, 0, 1};
  for (int i = 0; i < 3; ++i) for (int j = 0; j < 3; ++j) T.m[4 * i + j] = R[3 * i + j];
  T.m[3] = 1; T.m[7] = 2; T.m[11] = 3;
  std::ostringstream os;
  writeTrajectoryLine(os, true, 7, "123.5", T);
  writeTrajectoryLine(os, false, 7, "123.5", T);
  std::istringstream is(os.str());
  std::string stamp;
  double tx, ty, tz, qx, qy, qz, qw;
  is >> stamp >> tx >> ty >> tz >> qx >> qy >> qz >> qw;
  const double h = std::sqrt(0.5);
  if (stamp != "123.5" || std::fabs(tx + 2) > 1e-5 || std::fabs(ty - 1) > 1e-5 || std::fabs(tz + 3) > 1e-5 || std::fabs(qx) > 1e-5 || std::fabs(qy) > 1e-5 ||
      std::fabs(qz + h) > 1e-5 || std::fabs(qw - h) > 1e-5) { std::cerr << "trajectory line wrong: " << os.str(); ++fails; }
  is >> stamp;
  if (stamp != "7") { std::cerr << "id line wrong\n"; ++fails; }
  std::cout << (fails ? "FAIL" : "OK") << std::endl;
  return fails;
}
