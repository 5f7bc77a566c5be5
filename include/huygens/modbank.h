// modbank.h — drop-in soundmath::Modbank<T,N> (reference: src/modbank.h): element-wise modulation of a
// multichannel signal by a carrier.  N values per call on the host — glue of the per-sample API, not a hot
// loop; blocks of ticks use the fused form instead (amplitude-weighted mix: `out_amix` of
// Oscbank::process / process_env, which is Mixer(Modbank(amps, osc())) per tick).
#ifndef HUYGENS_B200_MODBANK_H
#define HUYGENS_B200_MODBANK_H
#include "multichannel.h"

namespace soundmath {

template <typename T, int N> class Modbank : public Multichannel<T, N> {
public:
    using typename Multichannel<T, N>::ArrayCT;
    using typename Multichannel<T, N>::ArrayT;
    using Multichannel<T, N>::activate;
    using Multichannel<T, N>::deactivate;
    using Multichannel<T, N>::open;
    using Multichannel<T, N>::close;
    using Multichannel<T, N>::activity;

    Modbank() : out(N) {}                                                            // src/modbank.h:28-33
    // src/modbank.h:41-45: *out = (*signal) * (*carrier)
    const ArrayCT* operator()(const ArrayCT* signal, const ArrayCT* carrier) {
        for (int i = 0; i < N; i++) out(i) = (*signal)(i) * (*carrier)(i);
        return &out;
    }
    // :47-51 / :53-57: one complex or real sample against every carrier
    const ArrayCT* operator()(std::complex<T> signal, const ArrayCT* carrier) {
        for (int i = 0; i < N; i++) out(i) = signal * (*carrier)(i);
        return &out;
    }
    const ArrayCT* operator()(T signal, const ArrayCT* carrier) {
        for (int i = 0; i < N; i++) out(i) = signal * (*carrier)(i);
        return &out;
    }

private:
    ArrayCT out;
};

}  // namespace soundmath
#endif
