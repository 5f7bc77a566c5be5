// slidebank.h — drop-in soundmath::Slidebank<T,N> over libhuygens_cuda (reference: src/slidebank.h).
// N channels, each a cascade of `order` complex one-pole low-passes with its own radius; the reference
// writes the cascade as one sparse matrix applied to [input ; states], so every stage lags the one before
// it by a tick.  Same public members (setup, operator(), tick, poll) plus the block entry point process().
#ifndef HUYGENS_B200_SLIDEBANK_H
#define HUYGENS_B200_SLIDEBANK_H
#include "multichannel.h"

namespace soundmath {

template <typename T, int N> class Slidebank : public Multichannel<T, N> {
public:
    using typename Multichannel<T, N>::ArrayCT;
    using typename Multichannel<T, N>::ArrayT;
    using Multichannel<T, N>::activate;
    using Multichannel<T, N>::deactivate;
    using Multichannel<T, N>::open;
    using Multichannel<T, N>::close;
    using Multichannel<T, N>::activity;

    Slidebank() : out(N) {}                                                          // src/slidebank.h:43
    Slidebank(int order, const std::vector<std::complex<T>>& radii) : out(N) { setup(order, radii); }   // :56-59
    ~Slidebank() { hg_casc_destroy(bank); }
    Slidebank(const Slidebank&) = delete;
    Slidebank& operator=(const Slidebank&) = delete;

    void setup(int order, const std::vector<std::complex<T>>& radii) {               // :61-97
        if ((int)radii.size() < N) throw cuda_error(HG_ERR_INVALID, "Slidebank::setup: need one radius per channel");
        std::vector<double> r(2 * (size_t)N);
        for (int i = 0; i < N; i++) { r[2 * (size_t)i] = (double)radii[(size_t)i].real(); r[2 * (size_t)i + 1] = (double)radii[(size_t)i].imag(); }
        hg_casc_destroy(bank);
        bank = nullptr;
        hg_check(hg_slide_create(&bank, hg_dtype<T>::value, HUYGENS_ARITH, N, order, r.data(), 0));
    }

    // :100-124 — one tick per call, memoised until tick() (a 1-tick device block: not a timed path)
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
    void tick() { computed = false; }                                                 // :126-129

    // :132-154 — which channels carry more than `proportion` of the total energy (host side, N values)
    void poll(bool* voices, T thresh, T proportion = 0.8) {
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
