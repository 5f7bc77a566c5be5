// oscbank.h — drop-in soundmath::Oscbank<T,N> over libhuygens_cuda (reference: src/oscbank.h).
#ifndef HUYGENS_B200_OSCBANK_H
#define HUYGENS_B200_OSCBANK_H
#include "multichannel.h"

namespace soundmath {

template <typename T, int N> class Oscbank : public Multichannel<T, N> {
public:
    using typename Multichannel<T, N>::ArrayCT;
    using Multichannel<T, N>::activate;
    using Multichannel<T, N>::deactivate;
    using Multichannel<T, N>::open;
    using Multichannel<T, N>::close;
    using Multichannel<T, N>::activity;

    ~Oscbank() { hg_osc_destroy(osc); }
    Oscbank(const Oscbank&) = delete;
    Oscbank& operator=(const Oscbank&) = delete;
    // reference: src/oscbank.h:37-47 (`k` only feeds the unused `stiffness` member, :37,96)
    Oscbank(double k = 2.0 / SR) : phases(N, std::complex<T>(1, 0)) {
        (void)k;
        hg_check(hg_osc_create(&osc, hg_dtype<T>::value, HUYGENS_ARITH, N, 0));
    }
    // reference: src/oscbank.h:49-56 — a bad index is silently ignored
    void freqmod(int index, T target) { hg_check(hg_osc_freqmod(osc, index, (double)target)); }
    // reference: src/oscbank.h:59-63
    void tick() {
        std::complex<T> dummy;
        hg_check(hg_osc_process(osc, &dummy, nullptr, nullptr, 1, nullptr));
        fresh = false;
    }
    // reference: src/oscbank.h:65-68 — pointer to the phase array (a host copy of the device state)
    const ArrayCT* operator()() { refresh(); return &phases; }
    // reference: src/oscbank.h:81-90 — complex sum over the active set
    std::complex<T> mixdown() {
        refresh();
        std::complex<T> value = 0;
        for (int i = 0; i < N; i++) if (this->active[(size_t)i]) value += phases(i);
        return value;
    }
    // new: L ticks at once; mix[t] = mixdown() before tick t; amix[t] = sum_n amps[n] * Re(phase_n)
    void process(std::complex<T>* mix, size_t L, const T* amps = nullptr, T* amix = nullptr, void* cuda_stream = nullptr) {
        hg_check(hg_osc_process(osc, mix, amps, amix, L, cuda_stream));
        fresh = false;
    }
    // new: the same with one amplitude vector per `hop` ticks — amp_rows[r * N + n] drives ticks [r*hop, (r+1)*hop)
    // (envelope rows of Filterbank::process_env as Modbank amplitudes, src/modbank.h:41-45)
    void process_env(const T* amp_rows, size_t n_rows, int hop, T* amix, size_t L, std::complex<T>* mix = nullptr,
                     void* cuda_stream = nullptr) {
        hg_check(hg_osc_process_env(osc, mix, amp_rows, n_rows, hop, amix, L, cuda_stream));
        fresh = false;
    }
    hg_osc* handle() { return osc; }

protected:
    void set_active(const std::vector<int>& indices, bool on) override {
        Multichannel<T, N>::set_active(indices, on);
        hg_check(hg_osc_activate(osc, indices.data(), (int)indices.size(), on ? 1 : 0));
    }

private:
    void refresh() {
        if (!fresh) { hg_check(hg_osc_get_phases(osc, phases.data())); fresh = true; }
    }
    hg_osc* osc = nullptr;
    ArrayCT phases;
    bool fresh = true;
};

}  // namespace soundmath
#endif
