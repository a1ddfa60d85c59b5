// eth_trajectory_generation_b200_io.hpp -- the segment YAML interchange of the reference (host side, no GPU work).
//
// Mirrors eth_trajectory_generation/io.h: segmentsToFile / segmentsFromFile / trajectoryToFile / trajectoryFromFile
// (src/eth_trajectory_generation/io.cpp:125-218) and the text forms behind them (io.cpp:34-122), for the Segment / Trajectory
// classes of eth_trajectory_generation_b200.hpp.  The document has the shape yaml-cpp emits for the reference:
//
//   segments:
//     - N: 10
//       D: 4
//       time: 1500000000
//       coefficients:
//         - [c0, c1, ..., c9]        one flow sequence per dimension, increasing powers
//
// keys io.cpp:27-31; the time is uint64 nanoseconds, static_cast<uint64_t>(1e9 * t) on write and ns * 1e-9 on read
// (segment.h:67-76), so a round trip quantises segment times to 1 ns exactly as the reference does.  Coefficients are written with
// 17 significant digits (yaml-cpp writes max_digits10 too), so they survive the round trip bit for bit.
// yaml-cpp is not a dependency: the writer emits the text directly and the reader accepts the block / flow subset above
// (comments, blank lines, any indentation that is consistent inside a segment); anything else makes the *FromYaml functions
// return false, like the reference's `return false` paths.  The Python mirror is mrs_uav_trajectory_generation_b200/segment_io.py;
// tests/test_segment_io.py checks that each side reads what the other wrote.
#ifndef ETH_TRAJECTORY_GENERATION_B200_IO_HPP_
#define ETH_TRAJECTORY_GENERATION_B200_IO_HPP_

#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <sstream>
#include <string>
#include <vector>

#include "eth_trajectory_generation_b200.hpp"

namespace eth_trajectory_generation {

namespace io_detail {
inline std::string strip(const std::string& s) {
  size_t a = 0, b = s.size();
  while (a < b && (s[a] == ' ' || s[a] == '\t' || s[a] == '\r')) ++a;
  while (b > a && (s[b - 1] == ' ' || s[b - 1] == '\t' || s[b - 1] == '\r')) --b;
  return s.substr(a, b - a);
}
inline std::string uncomment(const std::string& s) {  // no quoted strings in this format
  const size_t h = s.find('#');
  return h == std::string::npos ? s : s.substr(0, h);
}
inline bool parse_double(const std::string& t, double* v) {
  const std::string s = strip(t);
  if (s.empty()) return false;
  char* end = nullptr;
  *v = std::strtod(s.c_str(), &end);
  if (end == s.c_str() || *end != '\0') {
    // YAML spellings of the non-finite values (yaml-cpp writes .inf / -.inf / .nan)
    if (s == ".inf" || s == "+.inf") { *v = 1.0 / 0.0; return true; }
    if (s == "-.inf") { *v = -1.0 / 0.0; return true; }
    if (s == ".nan") { *v = 0.0 / 0.0; return true; }
    return false;
  }
  return true;
}
inline bool parse_uint64(const std::string& t, uint64_t* v) {
  const std::string s = strip(t);
  if (s.empty()) return false;
  for (char c : s)
    if (c < '0' || c > '9') return false;
  *v = std::strtoull(s.c_str(), nullptr, 10);
  return true;
}
inline bool parse_flow_sequence(const std::string& t, std::vector<double>* out) {
  const std::string s = strip(t);
  if (s.size() < 2 || s.front() != '[' || s.back() != ']') return false;
  out->clear();
  const std::string body = s.substr(1, s.size() - 2);
  if (strip(body).empty()) return true;
  std::stringstream ss(body);
  std::string item;
  while (std::getline(ss, item, ',')) {
    double v;
    if (!parse_double(item, &v)) return false;
    out->push_back(v);
  }
  return true;
}
}  // namespace io_detail

// segmentsToYaml / trajectoryToYaml (io.cpp:58-70) as text
inline std::string segmentsToYaml(const Segment::Vector& segments) {
  std::string out = "segments:\n";
  char buf[64];
  for (const Segment& s : segments) {
    out += "  - N: " + std::to_string(s.N()) + "\n";
    out += "    D: " + std::to_string(s.D()) + "\n";
    out += "    time: " + std::to_string(static_cast<uint64_t>(1.0e9 * s.getTime())) + "\n";  // getTimeNSec (segment.h:67-69)
    out += "    coefficients:\n";
    for (int d = 0; d < s.D(); ++d) {
      out += "      - [";
      for (int i = 0; i < s.N(); ++i) {
        std::snprintf(buf, sizeof(buf), "%.17g", s.coefficients(d)[i]);
        out += buf;
        if (std::string(buf).find_first_of(".en") == std::string::npos) out += ".0";  // keep it a YAML float ("-0" would be read as an integer and lose its sign)
        if (i + 1 < s.N()) out += ", ";
      }
      out += "]\n";
    }
  }
  return out;
}
inline std::string trajectoryToYaml(const Trajectory& trajectory) {
  Segment::Vector segments;
  trajectory.getSegments(&segments);
  return segmentsToYaml(segments);
}

// segmentsFromYaml (io.cpp:72-112): false when a key is missing, a coefficient row is not a sequence, the number of rows is not
// D or a row does not hold N numbers.  Segments of any N and D are read (what the device entry points then accept is their business:
// N in {6, 8, 10, 12}, D in 1..4)
inline bool segmentsFromYaml(const std::string& text, Segment::Vector* segments) {
  if (!segments) return false;
  segments->clear();
  std::stringstream ss(text);
  std::string line;
  bool have_root = false, in_coeffs = false;
  struct Pending {
    bool open = false, n = false, d = false, t = false, c = false;
    int N = 0, D = 0;
    uint64_t ns = 0;
    std::vector<std::vector<double>> rows;
  } cur;
  auto close = [&]() -> bool {
    if (!cur.open) return true;
    if (!(cur.n && cur.d && cur.t && cur.c)) return false;
    if (cur.N < 1 || cur.D < 1) return false;
    if ((int)cur.rows.size() != cur.D) return false;
    Segment s(cur.N, cur.D);  // any shape the document declares, as the reference reads it (Segment(N, D), io.cpp:85-88)
    for (int d = 0; d < cur.D; ++d) {
      if ((int)cur.rows[d].size() != cur.N) return false;
      for (int i = 0; i < cur.N; ++i) s.coefficients(d)[i] = cur.rows[d][i];
    }
    s.setTime(static_cast<double>(cur.ns) * 1.0e-9);  // setTimeNSec (segment.h:74-76)
    segments->push_back(s);
    cur = Pending();
    return true;
  };
  while (std::getline(ss, line)) {
    std::string t = io_detail::strip(io_detail::uncomment(line));
    if (t.empty() || t == "---") continue;
    if (!have_root) {
      if (t == "segments:") { have_root = true; continue; }
      if (t == "segments: []") return true;
      return false;
    }
    bool new_item = false;
    if (t.size() >= 2 && t[0] == '-' && t[1] == ' ') {
      const std::string rest = io_detail::strip(t.substr(2));
      if (in_coeffs && !rest.empty() && rest[0] == '[') {  // a coefficient row
        std::vector<double> row;
        if (!io_detail::parse_flow_sequence(rest, &row)) return false;
        cur.rows.push_back(row);
        continue;
      }
      new_item = true;
      t = rest;
    }
    if (new_item) {
      if (!close()) return false;
      cur.open = true;
      in_coeffs = false;
    }
    if (!cur.open) return false;
    const size_t colon = t.find(':');
    if (colon == std::string::npos) return false;
    const std::string key = io_detail::strip(t.substr(0, colon)), val = io_detail::strip(t.substr(colon + 1));
    in_coeffs = false;
    if (key == "N") {
      double v;
      if (!io_detail::parse_double(val, &v)) return false;
      cur.N = (int)v;
      cur.n = true;
    } else if (key == "D") {
      double v;
      if (!io_detail::parse_double(val, &v)) return false;
      cur.D = (int)v;
      cur.d = true;
    } else if (key == "time") {
      if (!io_detail::parse_uint64(val, &cur.ns)) return false;
      cur.t = true;
    } else if (key == "coefficients") {
      cur.c = true;
      if (!val.empty()) return false;  // the rows follow as a block sequence
      in_coeffs = true;
    } else {
      return false;
    }
  }
  if (!have_root) return false;
  return close();
}
inline bool trajectoryFromYaml(const std::string& text, Trajectory* trajectory) {
  if (!trajectory) return false;
  Segment::Vector segments;
  if (!segmentsFromYaml(text, &segments)) return false;
  trajectory->setSegments(segments);
  return true;
}

// segmentsToFile / segmentsFromFile / trajectoryToFile / trajectoryFromFile (io.cpp:125-218)
inline bool segmentsToFile(const std::string& filename, const Segment::Vector& segments) {
  std::ofstream f(filename.c_str());
  if (!f.is_open()) return false;
  f << segmentsToYaml(segments);
  return f.good();
}
inline bool trajectoryToFile(const std::string& filename, const Trajectory& trajectory) {
  Segment::Vector segments;
  trajectory.getSegments(&segments);
  return segmentsToFile(filename, segments);
}
inline bool segmentsFromFile(const std::string& filename, Segment::Vector* segments) {
  std::ifstream f(filename.c_str());
  if (!f.is_open()) return false;
  std::stringstream ss;
  ss << f.rdbuf();
  return segmentsFromYaml(ss.str(), segments);
}
inline bool trajectoryFromFile(const std::string& filename, Trajectory* trajectory) {
  if (!trajectory) return false;
  Segment::Vector segments;
  if (!segmentsFromFile(filename, &segments)) return false;
  trajectory->setSegments(segments);
  return true;
}

}  // namespace eth_trajectory_generation

#endif  // ETH_TRAJECTORY_GENERATION_B200_IO_HPP_
