// rmsbank.h — drop-in soundmath::RMSbank<T,N> over libhuygens_cuda (reference: src/rmsbank.h).
#ifndef HUYGENS_B200_RMSBANK_H
#define HUYGENS_B200_RMSBANK_H
#include "multichannel.h"

namespace soundmath {

template <typename T, int N> class RMSbank : public Multichannel<T, N> {
public:
    using typename Multichannel<T, N>::ArrayCT;
    using typename Multichannel<T, N>::ArrayT;
    using Multichannel<T, N>::activate;
    using Multichannel<T, N>::deactivate;
    using Multichannel<T, N>::open;
    using Multichannel<T, N>::close;
    using Multichannel<T, N>::activity;

    ~RMSbank() { hg_rms_destroy(rms); }
    RMSbank(const RMSbank&) = delete;
    RMSbank& operator=(const RMSbank&) = delete;
    // reference: src/rmsbank.h:32-45
    RMSbank(unsigned width = SR / 20) : width((int)width), out(N, (T)0) {
        hg_check(hg_rms_create(&rms, hg_dtype<T>::value, HUYGENS_ARITH, N, (int)width, 0));
    }
    // reference: src/rmsbank.h:54-66.  The first call of a tick runs one device tick; every further
    // call before tick() re-applies sqrt(out / width) in place, as the reference does (:64).
    const ArrayT* operator()(const ArrayCT* input) {
        if (!computed) {
            std::vector<T> re((size_t)N), im((size_t)N);
            for (int i = 0; i < N; i++) { re[(size_t)i] = (*input)(i).real(); im[(size_t)i] = (*input)(i).imag(); }
            hg_check(hg_rms_process(rms, re.data(), im.data(), out.data(), 1, nullptr));
            computed = true;
        } else {
            for (int i = 0; i < N; i++) out(i) = std::sqrt(out(i) / (T)width);
        }
        return &out;
    }
    // reference: src/rmsbank.h:68-74
    void tick() { computed = false; }
    // new: L ticks, channel-major [N][L] input (zim may be null) and output
    void process(const T* zre, const T* zim, T* result, size_t L, void* cuda_stream = nullptr) {
        hg_check(hg_rms_process(rms, zre, zim, result, L, cuda_stream));
    }
    // new: L ticks of `gated = latchbank(&rmsbank, z); rmsbank.tick();` (latchbank.h:81-100) fused on the device
    void set_latch(T thresh, T ratio = 0.2) { hg_check(hg_rms_set_latch(rms, (double)thresh, (double)ratio)); }
    void process_latch(const T* zre, const T* zim, T* gated_re, T* gated_im, size_t L, T* level = nullptr, void* cuda_stream = nullptr) {
        hg_check(hg_rms_process_latch(rms, zre, zim, gated_re, gated_im, level, L, cuda_stream));
    }

private:
    hg_rms* rms = nullptr;
    int width;
    bool computed = false;
    ArrayT out;
};

}  // namespace soundmath
#endif
