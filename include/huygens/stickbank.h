// stickbank.h — drop-in soundmath::Stickbank<T,N> over libhuygens_cuda (reference: src/stickbank.h).
// N complex channels through one all-pole smoother with `order` poles derived from `rad`
// (y = (1+rad)^order x - sum_k back[k] y[t-1-k]); same public members (rmod, operator(), tick, poll)
// plus the block entry point process().  The reference template only compiles for T = double
// (src/stickbank.h:49-50); this one also accepts float.
#ifndef HUYGENS_B200_STICKBANK_H
#define HUYGENS_B200_STICKBANK_H
#include "multichannel.h"

namespace soundmath {

template <typename T, int N> class Stickbank : public Multichannel<T, N> {
public:
    using typename Multichannel<T, N>::ArrayCT;
    using typename Multichannel<T, N>::ArrayT;
    using Multichannel<T, N>::activate;
    using Multichannel<T, N>::deactivate;
    using Multichannel<T, N>::open;
    using Multichannel<T, N>::close;
    using Multichannel<T, N>::activity;

    Stickbank(int order, T rad) : out(N) {                                            // src/stickbank.h:46-67
        hg_check(hg_stick_create(&bank, hg_dtype<T>::value, HUYGENS_ARITH, N, order, (double)rad, 0));
    }
    ~Stickbank() { hg_casc_destroy(bank); }
    Stickbank(const Stickbank&) = delete;
    Stickbank& operator=(const Stickbank&) = delete;

    void rmod(T rad) { hg_check(hg_stick_rmod(bank, (double)rad)); }                  // :69-79

    // :82-98 — one tick per call, memoised until tick() (a 1-tick device block: not a timed path)
    const ArrayCT* operator()(const ArrayCT* input) {
        if (!computed) {
            std::vector<T> re(N), im(N), ore(N), oim(N);
            for (int i = 0; i < N; i++) { re[(size_t)i] = (*input)(i).real(); im[(size_t)i] = (*input)(i).imag(); }
            hg_check(hg_casc_process(bank, re.data(), im.data(), ore.data(), oim.data(), 1, nullptr));
            for (int i = 0; i < N; i++) out(i) = std::complex<T>(ore[(size_t)i], oim[(size_t)i]);
            computed = true;
        }
        return &out;
    }
    const ArrayCT* operator()(const ArrayT* input) {
        ArrayCT casted(N);
        for (int i = 0; i < N; i++) casted(i) = std::complex<T>((*input)(i), 0);
        return (*this)(&casted);
    }
    const ArrayCT* operator()() { return computed ? &out : nullptr; }
    void tick() { computed = false; }                                                 // :101-107 (the ring is device state)

    void poll(bool* voices, T thresh, T proportion = 0.8) {                           // :110-137
        T total = 0;
        for (int i = 0; i < N; i++) total += std::norm(out(i));
        const T square = thresh * thresh;
        for (int i = 0; i < N; i++) voices[i] = total < square ? false : std::norm(out(i)) > total * proportion;
    }

    // new: L ticks at once; channel-major [N][L] real / imaginary planes, host or device pointers
    void process(const T* zre, const T* zim, T* out_re, T* out_im, size_t L, void* cuda_stream = nullptr) {
        hg_check(hg_casc_process(bank, zre, zim, out_re, out_im, L, cuda_stream));
        computed = false;
    }
    hg_casc* handle() { return bank; }

private:
    hg_casc* bank = nullptr;
    ArrayCT out;
    bool computed = false;
};

}  // namespace soundmath
#endif
