// multichannel.h — active-channel bookkeeping shared by the banks (reference: src/multichannel.h).
// The reference keeps a bool mask plus a sorted index list maintained by a recursive binary search
// (:121-157); here the mask lives in the CUDA handle and this base only forwards.
#ifndef HUYGENS_B200_MULTICHANNEL_H
#define HUYGENS_B200_MULTICHANNEL_H
#include "includes.h"

namespace soundmath {

template <typename T, int N> class Multichannel {
public:
    typedef ArrayView<std::complex<T>> ArrayCT;
    typedef ArrayView<T> ArrayT;
    Multichannel() : active(N, false) {}
    virtual ~Multichannel() {}

protected:
    std::vector<bool> active;
    virtual void set_active(const std::vector<int>& indices, bool on) {
        for (int i : indices) if (0 <= i && i < N) active[(size_t)i] = on;
    }
    const std::vector<bool>& activity() { return active; }
    void activate(const std::vector<int>& indices) { set_active(indices, true); }        // :87-92
    void deactivate(const std::vector<int>& indices) { set_active(indices, false); }     // :95-100
    void open() { std::vector<int> v(N); for (int i = 0; i < N; i++) v[i] = i; activate(v); }      // :103-109
    void close() { std::vector<int> v(N); for (int i = 0; i < N; i++) v[i] = i; deactivate(v); }   // :112-118
};

}  // namespace soundmath
#endif
