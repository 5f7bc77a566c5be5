// latchbank.h — drop-in soundmath::Latchbank<T,N> over libhuygens_cuda (reference: src/latchbank.h).
// A hysteresis gate per channel, driven by an RMSbank level: channels arm below thresh*ratio, engage
// above thresh*(1-ratio) once armed, and disengage (and disarm) when they fall below thresh*ratio again.
#ifndef HUYGENS_B200_LATCHBANK_H
#define HUYGENS_B200_LATCHBANK_H
#include "rmsbank.h"

namespace soundmath {

template <typename T, int N> class Latchbank : public Multichannel<T, N> {
public:
    using typename Multichannel<T, N>::ArrayCT;
    using typename Multichannel<T, N>::ArrayT;
    using Multichannel<T, N>::activate;
    using Multichannel<T, N>::deactivate;
    using Multichannel<T, N>::open;
    using Multichannel<T, N>::close;
    using Multichannel<T, N>::activity;

    // reference: src/latchbank.h:34-53
    Latchbank(T thresh, T ratio = 0.2) : thresh(thresh), ratio(ratio), armed(N, false), engaged(N, false), out(N, (T)0), outC(N) {}

    // reference: src/latchbank.h:61-79 — gate a real level array by itself (host-side: N booleans)
    const ArrayT* operator()(const ArrayT* input) {
        step(*input);
        for (int i = 0; i < N; i++) out(i) = (*input)(i) * (T)(engaged[(size_t)i] ? 1 : 0);
        return &out;
    }
    // reference: src/latchbank.h:81-100 — level = (*rms)(signal) on the device, gate logic on the host
    // (the per-sample API is not a timed path; blocks go through process())
    const ArrayCT* operator()(RMSbank<T, N>* rms, const ArrayCT* signal) {
        const ArrayT input = *(*rms)(signal);
        step(input);
        for (int i = 0; i < N; i++) outC(i) = (*signal)(i) * (T)(engaged[(size_t)i] ? 1 : 0);
        return &outC;
    }

private:
    void step(const ArrayT& in) {
        const T lo = thresh * ratio, hi = thresh * (1 - ratio);
        for (int i = 0; i < N; i++) {
            const size_t k = (size_t)i;
            armed[k] = armed[k] || (in(i) < lo);
            const bool t1 = engaged[k] && (in(i) < lo);
            const bool t2 = !engaged[k] && (in(i) > hi) && armed[k];
            engaged[k] = engaged[k] && !t1;
            armed[k] = armed[k] && !t1;
            engaged[k] = engaged[k] || t2;
        }
    }
    T thresh, ratio;
    std::vector<bool> armed, engaged;
    ArrayT out;
    ArrayCT outC;
};

}  // namespace soundmath
#endif
