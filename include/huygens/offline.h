// offline.h — offline render harness that stands in for the reference's PortAudio stream
// (src/audio.h) for benchmarking and tests: same `int process(const float* in, float* out)`
// callback convention (audio.h:18,139-148), synthetic input buffers instead of a sound card.
#ifndef HUYGENS_B200_OFFLINE_H
#define HUYGENS_B200_OFFLINE_H
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <iterator>
#include <stdexcept>
#include <string>
#include <vector>
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

// ---- on-disk audio for the harness (the reference has no file I/O at all: src/audio.h is PortAudio only) ----
// read_audio: mono float samples from a raw little-endian float32 file (any other extension) or a RIFF/WAVE
// file (.wav: PCM 16-bit or IEEE float 32-bit; the first channel of interleaved data).  Returns the sample
// rate through `rate` (raw files: unchanged).
inline std::vector<float> read_audio(const std::string& path, int* rate = nullptr) {
    std::ifstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("read_audio: cannot open " + path);
    std::vector<char> bytes((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
    const bool wav = path.size() > 4 && path.compare(path.size() - 4, 4, ".wav") == 0;
    std::vector<float> x;
    if (!wav) {
        x.resize(bytes.size() / sizeof(float));
        std::memcpy(x.data(), bytes.data(), x.size() * sizeof(float));
        return x;
    }
    auto u16 = [&](size_t o) { return (unsigned)(unsigned char)bytes[o] | ((unsigned)(unsigned char)bytes[o + 1] << 8); };
    auto u32 = [&](size_t o) { return u16(o) | (u16(o + 2) << 16); };
    if (bytes.size() < 12 || std::memcmp(bytes.data(), "RIFF", 4) || std::memcmp(bytes.data() + 8, "WAVE", 4))
        throw std::runtime_error("read_audio: not a RIFF/WAVE file: " + path);
    unsigned fmt = 0, channels = 1, bits = 16;
    for (size_t o = 12; o + 8 <= bytes.size();) {
        const unsigned len = u32(o + 4);
        if (!std::memcmp(bytes.data() + o, "fmt ", 4)) { fmt = u16(o + 8); channels = u16(o + 10); if (rate) *rate = (int)u32(o + 12); bits = u16(o + 22); }
        else if (!std::memcmp(bytes.data() + o, "data", 4)) {
            const size_t n = std::min<size_t>(len, bytes.size() - o - 8), frame = (size_t)channels * bits / 8;
            if (!frame || !((fmt == 1 && bits == 16) || (fmt == 3 && bits == 32)))
                throw std::runtime_error("read_audio: only PCM16 and float32 WAVE data are read");
            x.resize(n / frame);
            for (size_t i = 0; i < x.size(); i++) {
                const size_t p = o + 8 + i * frame;
                if (fmt == 1) x[i] = (float)(short)u16(p) / 32768.0f;
                else { unsigned v = u32(p); std::memcpy(&x[i], &v, 4); }
            }
            return x;
        }
        o += 8 + len + (len & 1);
    }
    throw std::runtime_error("read_audio: no data chunk in " + path);
}
// write_audio: ".wav" => mono IEEE-float32 RIFF/WAVE at `rate`; anything else => raw float32
inline void write_audio(const std::string& path, const std::vector<float>& x, int rate = SR) {
    std::ofstream f(path, std::ios::binary);
    if (!f) throw std::runtime_error("write_audio: cannot open " + path);
    const bool wav = path.size() > 4 && path.compare(path.size() - 4, 4, ".wav") == 0;
    if (wav) {
        auto put32 = [&](unsigned v) { char b[4] = {(char)v, (char)(v >> 8), (char)(v >> 16), (char)(v >> 24)}; f.write(b, 4); };
        auto put16 = [&](unsigned v) { char b[2] = {(char)v, (char)(v >> 8)}; f.write(b, 2); };
        const unsigned data = (unsigned)(x.size() * sizeof(float));
        f.write("RIFF", 4); put32(36 + data); f.write("WAVE", 4);
        f.write("fmt ", 4); put32(16); put16(3); put16(1); put32((unsigned)rate); put32((unsigned)rate * 4); put16(4); put16(32);
        f.write("data", 4); put32(data);
    }
    f.write(reinterpret_cast<const char*>(x.data()), (std::streamsize)(x.size() * sizeof(float)));
}

}  // namespace soundmath
#endif
