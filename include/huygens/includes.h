// includes.h — the few globals of the reference's src/includes.h that the bank classes need.
// (The reference header also pulls in PortAudio, RtMidi and Eigen: none of them is needed here.)
#ifndef HUYGENS_B200_INCLUDES_H
#define HUYGENS_B200_INCLUDES_H
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstddef>
#include <cstdio>
#include <iostream>
#include <limits>
#include <stdexcept>
#include <string>
#include <vector>

#include "../huygens_cuda.h"

// reference: src/includes.h:30-32 (PI is deliberately the reference's 12-digit constant)
const double PI = 3.14159265359;
const int SR = 48000;

// reference: src/includes.h:38-48 — smoothing coefficient that relaxes to double epsilon in k seconds
inline double relaxation(double k) {
    if (k == 0) return 0;
    return std::pow(2.0, std::log2(std::numeric_limits<double>::epsilon()) / (std::fmax(0, k) * SR));
}

namespace soundmath {

// thrown by the drop-in classes when the C ABI reports an error (the reference aborts through
// Eigen assertions instead: tests/build.sh:8 has no -DNDEBUG)
struct cuda_error : std::runtime_error {
    int status;
    cuda_error(int s, const char* what) : std::runtime_error(what), status(s) {}
};
inline void hg_check(int status) {
    if (status != HG_OK) throw cuda_error(status, hg_last_error());
}
template <typename T> struct hg_dtype;
template <> struct hg_dtype<float> { static const int value = HG_F32; };
template <> struct hg_dtype<double> { static const int value = HG_F64; };

#ifndef HUYGENS_ARITH
#define HUYGENS_ARITH HG_ARITH_STRICT   // x86-64 g++ build of the reference; -DHUYGENS_ARITH=HG_ARITH_FMA for arm64-like
#endif

// minimal stand-in for the Eigen::Array<T, Dynamic, 1> pointers the reference's banks hand around
template <typename S> class ArrayView {
public:
    ArrayView() {}
    explicit ArrayView(size_t n, S fill = S()) : v(n, fill) {}
    S& operator()(size_t i) { return v[i]; }
    const S& operator()(size_t i) const { return v[i]; }
    size_t size() const { return v.size(); }
    S* data() { return v.data(); }
    const S* data() const { return v.data(); }
    S sum() const { S s = S(); for (const S& x : v) s += x; return s; }
    std::vector<S> v;
};

}  // namespace soundmath
#endif
