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
  std::printf("segments %d rejected %d empty_ok %d max_time %.17g\n", traj.K(), rejected, empty_ok ? 1 : 0, traj.getMaxTime());
  return 0;
}
