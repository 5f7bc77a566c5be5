// mixer.h — drop-in soundmath::Mixer<T,N> (reference: src/mixer.h): sum of the N channels, in index order
// like Eigen's sum().  Host-side glue of the per-sample API; the block kernels mix on the device.
#ifndef HUYGENS_B200_MIXER_H
#define HUYGENS_B200_MIXER_H
#include "multichannel.h"

namespace soundmath {

template <typename T, int N> class Mixer : public Multichannel<T, N> {
public:
    using typename Multichannel<T, N>::ArrayCT;
    using typename Multichannel<T, N>::ArrayT;
    using Multichannel<T, N>::activate;
    using Multichannel<T, N>::deactivate;
    using Multichannel<T, N>::open;
    using Multichannel<T, N>::close;
    using Multichannel<T, N>::activity;

    Mixer() {}
    const T operator()(const ArrayCT* input) { return input->sum().real(); }         // src/mixer.h:32-35
    const T operator()(const ArrayT* input) { return input->sum(); }                 // :37-40
};

}  // namespace soundmath
#endif
