// render.cpp — offline render of a resonant filter bank, written the way the reference's programs are
// (tests/resynthesis.cpp, tests/noise.cpp): a global bank, an `int process(const float*, float*)`
// callback, a main() that fills the bank and starts "audio" — here soundmath::OfflineAudio over a
// synthetic buffer instead of PortAudio.  Prints one JSON line; optionally dumps the rendered floats.
//
//   render [--filters N] [--seconds S] [--double] [--input white|chirp] [--bsize B] [--per-sample]
//          [--recipe resynthesis|strings] [--in file.wav|file.f32] [--out file.wav|file.f32]
#include <cstring>
#include <fstream>
#include <memory>
#include "../include/huygens/filterbank.h"
#include "../include/huygens/offline.h"

using namespace soundmath;

static int BSIZE = 1024;
static bool per_sample = false;
static std::unique_ptr<Filterbank<double>> bank_d;
static std::unique_ptr<Filterbank<float>> bank_f;
static int N = 64;

// the bandpass magnitude the reference's programs normalise with (tests/resynthesis.cpp:23-27)
static std::complex<double> transfer(double frequency, double r, double f) {
    std::complex<double> z(std::cos(2 * PI * f / SR), -std::sin(2 * PI * f / SR));
    return (1.0 - z * z) / (1.0 - 2 * r * std::cos(2 * PI * frequency / SR) * z + r * r * z * z);
}

template <typename T> static int run_block(Filterbank<T>& bank, const float* in, float* out) {
    if (per_sample) {
        // the reference's loop shape (tests/resynthesis.cpp:33-42)
        for (int i = 0; i < BSIZE; i++) {
            out[i] = (float)bank((T)in[i]);
            bank.tick();
        }
    } else {
        std::vector<T> x(in, in + BSIZE), y((size_t)BSIZE);
        bank.process(x.data(), y.data(), (size_t)BSIZE, HG_MODE_EXACT);
        for (int i = 0; i < BSIZE; i++) out[i] = (float)y[(size_t)i];
    }
    return 0;
}
static int process(const float* in, float* out) {
    return bank_d ? run_block(*bank_d, in, out) : run_block(*bank_f, in, out);
}

template <typename T> static void fill(Filterbank<T>& bank, const std::string& recipe) {
    const double R = 0.999;
    if (recipe == "resynthesis") {
        // tests/resynthesis.cpp:48-54, with N+1 in the denominator (the original puts a filter on
        // Nyquist) and the boost()/open() calls it forgets
        for (int i = 0; i < N; i++) {
            const double frequency = 0.5 * (i + 1) * SR / (N + 1);
            const double cosine = std::cos(2 * PI * frequency / SR);
            const double gain = std::abs(transfer(frequency, R, frequency));
            bank.coefficients(i, {(T)(1.0 / gain), (T)0, (T)(-1.0 / gain)}, {(T)(-2 * R * cosine), (T)(R * R)});
        }
        bank.boost(std::vector<T>((size_t)N, (T)1));
        bank.open();
    } else {
        // tests/filterbank.cpp:255-271: {1,0,-1}, gains 1/|H| through mix()
        std::vector<T> gains((size_t)N);
        for (int i = 0; i < N; i++) {
            const double partial = N > 1 ? 27.5 * std::pow(20000.0 / 27.5, (double)i / (N - 1)) : 27.5;
            const double angle = partial * (2 * PI / SR);
            gains[(size_t)i] = (T)(1.0 / std::abs(transfer(partial, R, partial)));
            bank.coefficients(i, {(T)1.0, (T)0, (T)-1.0}, {(T)(-2 * R * std::cos(angle)), (T)(R * R)});
        }
        bank.mix(gains);
        bank.boost(std::vector<T>((size_t)N, (T)1));
    }
}

int main(int argc, char** argv) {
    double seconds = 1.0;
    bool dbl = false;
    std::string input = "white", recipe = "resynthesis", out_path, in_path;
    for (int i = 1; i < argc; i++) {
        std::string a = argv[i];
        auto next = [&]() { return std::string(i + 1 < argc ? argv[++i] : ""); };
        if (a == "--filters") N = std::stoi(next());
        else if (a == "--seconds") seconds = std::stod(next());
        else if (a == "--double") dbl = true;
        else if (a == "--input") input = next();
        else if (a == "--bsize") BSIZE = std::stoi(next());
        else if (a == "--per-sample") per_sample = true;
        else if (a == "--recipe") recipe = next();
        else if (a == "--out") out_path = next();
        else if (a == "--in") in_path = next();
        else { std::fprintf(stderr, "unknown flag %s\n", a.c_str()); return 2; }
    }
    try {
        const double k_p = recipe == "resynthesis" ? 0.1 : 0.001, k_g = 1;
        if (dbl) { bank_d.reset(new Filterbank<double>(2, N, k_p, k_g)); fill(*bank_d, recipe); }
        else { bank_f.reset(new Filterbank<float>(2, N, k_p, k_g)); fill(*bank_f, recipe); }
        std::vector<float> in;
        if (!in_path.empty()) in = read_audio(in_path);           // .wav (PCM16 / float32, first channel) or raw float32
        else in = input == "chirp" ? chirp<float>((size_t)(seconds * SR)) : white_noise<float>((size_t)(seconds * SR));
        const size_t frames = in.size();
        std::vector<float> out(frames);
        OfflineAudio A(process, BSIZE);
        const double t = A.render(in.data(), out.data(), frames);
        double peak = 0;
        for (float v : out) peak = std::max(peak, (double)std::fabs(v));
        std::printf("{\"filters\": %d, \"frames\": %zu, \"dtype\": \"%s\", \"bsize\": %d, \"per_sample\": %s, "
                    "\"seconds_wall\": %.6f, \"filter_samples_per_s\": %.4e, \"peak\": %.9g}\n",
                    N, frames, dbl ? "f64" : "f32", BSIZE, per_sample ? "true" : "false", t, (double)N * frames / t, peak);
        if (!out_path.empty()) write_audio(out_path, out);        // .wav => float32 RIFF/WAVE, else raw float32
    } catch (const cuda_error& e) {
        std::fprintf(stderr, "huygens_cuda error %d: %s\n", e.status, e.what());
        return 1;
    } catch (const std::exception& e) {
        std::fprintf(stderr, "%s\n", e.what());
        return 1;
    }
    return 0;
}
