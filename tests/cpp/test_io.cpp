// The C++ YAML segment I/O (include/eth_trajectory_generation_b200_io.hpp) against the Python mirror: reads argv[1] (written by
// segment_io.py), writes it back to argv[2] through trajectoryToFile; prints the verdict on a few malformed documents.  No device work.
#include <cstdio>
#include <string>

#include "../../include/eth_trajectory_generation_b200_io.hpp"

using namespace eth_trajectory_generation;

int main(int argc, char** argv) {
  if (argc < 3) return 2;
  Trajectory traj;
  if (!trajectoryFromFile(argv[1], &traj)) return 3;
  if (!trajectoryToFile(argv[2], traj)) return 4;
  Segment::Vector segs;
  const char* bad[] = {
      "foo: 1\n",
      "segments:\n  - N: 10\n    D: 4\n    time: 5\n    coefficients:\n      - [1.0]\n",                              // rows != D
      "segments:\n  - N: 10\n    D: 4\n    coefficients:\n      - [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]\n",                  // no time
      "segments:\n  - N: 10\n    D: 4\n    time: -5\n    coefficients:\n      - [1, 2, 3, 4, 5, 6, 7, 8, 9, 10]\n",     // negative time
      "segments:\n  - N: 10\n    D: 4\n    time: 5\n    coefficients: 3\n",                                             // not a sequence
  };
  int rejected = 0;
  for (const char* b : bad) rejected += segmentsFromYaml(b, &segs) ? 0 : 1;
  const bool empty_ok = segmentsFromYaml("segments: []\n", &segs) && segs.empty();
  // another shape (N = 6, D = 3): read, written back identically, evaluated through the general-shape entry point
  const char* n6 =
      "segments:\n  - N: 6\n    D: 3\n    time: 2000000000\n    coefficients:\n      - [1.0, 2.0, 0.5, 0.0, 0.0, 0.25]\n"
      "      - [0.0, -1.0, 0.0, 0.0, 0.0, 0.0]\n      - [3.0, 0.0, 0.0, 0.0, 0.0, 0.0]\n";
  Segment::Vector s6;
  int n6_ok = segmentsFromYaml(n6, &s6) && s6.size() == 1 && s6[0].N() == 6 && s6[0].D() == 3 && s6[0].coefficients(0)[5] == 0.25 ? 1 : 0;
  if (n6_ok) {
    Segment::Vector again;
    n6_ok = segmentsFromYaml(segmentsToYaml(s6), &again) && again.size() == 1 && again[0].N() == 6 && again[0].coefficients(1)[1] == -1.0 ? 1 : 0;
    Trajectory t6;
    t6.setSegments(s6);
    const Vector p = t6.evaluate(1.0, derivative_order::POSITION);  // 1 + 2 + 0.5 + 0.25, -1, 3
    n6_ok = n6_ok && t6.N() == 6 && t6.D() == 3 && p.size() == 3 && p[0] == 3.75 && p[1] == -1.0 && p[2] == 3.0 ? 1 : 0;
  }
  std::printf("segments %d rejected %d empty_ok %d max_time %.17g n6_ok %d\n", traj.K(), rejected, empty_ok ? 1 : 0, traj.getMaxTime(), n6_ok);
  return 0;
}
