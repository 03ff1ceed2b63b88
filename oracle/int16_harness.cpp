// Harness around the reference's own float -> int16 loops (TEST INFRASTRUCTURE, see oracle/__init__.py).
//
// The body is NOT in this repository: oracle/build_ref.py cuts /root/reference/src/cpp/piper.cpp:411-431 (from
// "// Get max audio value for scaling" to the line before "// Clean up") into oracle/_ref/int16_body.inc at build time
// and this file gives those lines the names they use inside piper::synthesize: `audio`, `audioCount`, `audioBuffer`,
// `MAX_WAV_VALUE` (piper.cpp:29).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <limits>
#include <vector>

namespace {
const float MAX_WAV_VALUE = 32767.0f;   // piper.cpp:29
using std::abs;                         // the reference calls unqualified abs(float) with <cmath> overloads visible

void run(const float* audio, int64_t audioCount, std::vector<int16_t>& audioBuffer) {
#include "_ref/int16_body.inc"
}
}  // namespace

extern "C" void ref_float_to_int16(const float* audio, int64_t n, int16_t* out) {
  std::vector<int16_t> buf;
  run(audio, n, buf);
  std::copy(buf.begin(), buf.end(), out);
}
