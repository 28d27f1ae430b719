// DCHECK family for the image_compression public headers: fatal under _DEBUG, compiled out otherwise
// (same contract as the reference's base/logging.h:34-71).
#ifndef BASE_LOGGING_H_
#define BASE_LOGGING_H_

#include <cstdlib>
#include <iostream>

namespace icb_logging {
struct Sink {  // swallows a streamed message; aborts on destruction when armed
  explicit Sink(bool fatal) : fatal_(fatal) {}
  ~Sink() {
    if (fatal_) {
      std::cerr << std::endl;
      std::abort();
    }
  }
  template <typename T>
  Sink &operator<<(const T &v) {
    if (fatal_) std::cerr << v;
    return *this;
  }
  bool fatal_;
};
}  // namespace icb_logging

#ifdef _DEBUG
#define DCHECK(c) \
  if (c) {        \
  } else          \
    ::icb_logging::Sink(true) << __FILE__ << ":" << __LINE__ << " DCHECK failed: " #c " "
#else
#define DCHECK(c) \
  if (true) {     \
  } else          \
    ::icb_logging::Sink(false)
#endif
#define DCHECK_EQ(a, b) DCHECK((a) == (b))
#define DCHECK_NE(a, b) DCHECK((a) != (b))
#define DCHECK_LT(a, b) DCHECK((a) < (b))
#define DCHECK_LE(a, b) DCHECK((a) <= (b))
#define DCHECK_GT(a, b) DCHECK((a) > (b))
#define DCHECK_GE(a, b) DCHECK((a) >= (b))

#endif  // BASE_LOGGING_H_
