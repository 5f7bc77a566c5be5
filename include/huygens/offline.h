// offline.h — offline render harness that stands in for the reference's PortAudio stream
// (src/audio.h) for benchmarking and tests: same `int process(const float* in, float* out)`
// callback convention (audio.h:18,139-148), synthetic input buffers instead of a sound card.
#ifndef HUYGENS_B200_OFFLINE_H
#define HUYGENS_B200_OFFLINE_H
#include <chrono>
#include <cstdint>
#include "includes.h"

namespace soundmath {

// Audio(process, bsize) of the reference, minus the device: render() feeds `frames` mono frames
// through the callback in blocks of `bsize` (the last block is zero-padded, as a sound card would).
class OfflineAudio {
public:
    OfflineAudio(int (*proc)(const float* in, float* out), int bsize) : process(proc), bsize(bsize) {}
    double render(const float* in, float* out, size_t frames) {
        std::vector<float> bi((size_t)bsize), bo((size_t)bsize);
        const auto t0 = std::chrono::steady_clock::now();
        for (size_t f = 0; f < frames; f += (size_t)bsize) {
            const size_t n = std::min<size_t>((size_t)bsize, frames - f);
            if (n == (size_t)bsize) process(in + f, out + f);
            else {
                std::fill(bi.begin(), bi.end(), 0.0f);
                std::copy(in + f, in + f + n, bi.begin());
                process(bi.data(), bo.data());
                std::copy(bo.begin(), bo.begin() + (long)n, out + f);
            }
        }
        return std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    }
    int (*process)(const float* in, float* out);
    int bsize;
};

// counter-based uniform(-1,1): x[t] = 2 * (splitmix64(seed_s + (t+1) * GOLDEN) >> 11) * 2^-53 - 1
// (same definition as huygens_b200/design.py; replaces Noise<T>, src/noise.h:25-28, which is rand())
inline double white_sample(uint64_t seed, uint64_t stream, uint64_t t) {
    uint64_t z = (seed ^ (stream * 0xD1342543DE82EF95ull)) + (t + 1) * 0x9E3779B97F4A7C15ull;
    z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
    z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
    z ^= z >> 31;
    return 2.0 * ((double)(z >> 11) * 0x1.0p-53) - 1.0;
}
template <typename S> std::vector<S> white_noise(size_t L, uint64_t seed = 0x5EED, uint64_t stream = 0) {
    std::vector<S> x(L);
    for (size_t t = 0; t < L; t++) x[t] = (S)white_sample(seed, stream, t);
    return x;
}
// exponential sine sweep f0 -> f1 over L samples (double, then cast)
template <typename S> std::vector<S> chirp(size_t L, double f0 = 20.0, double f1 = 20000.0, double amp = 0.5) {
    std::vector<S> x(L);
    const double k = f1 / f0, T = (double)L / SR, two_pi = 6.283185307179586476925286766559;
    for (size_t t = 0; t < L; t++)
        x[t] = (S)(amp * std::sin(two_pi * f0 * T / std::log(k) * (std::pow(k, (double)t / (double)L) - 1.0)));
    return x;
}

}  // namespace soundmath
#endif
