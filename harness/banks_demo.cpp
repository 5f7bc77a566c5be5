// banks_demo.cpp — exercises the drop-in Oscbank / RMSbank / FFilterbank per-sample APIs the way
// tests/rmsbank.cpp and tests/harmbank.cpp chain them (osc() -> rms(...) ; tick()), printing values
// for the pytest parity check.   banks_demo <ticks>
#include "../include/huygens/filterbank.h"
#include "../include/huygens/oscbank.h"
#include "../include/huygens/rmsbank.h"
#include "../include/huygens/latchbank.h"
#include "../include/huygens/slidebank.h"
#include "../include/huygens/stickbank.h"
#include "../include/huygens/modbank.h"
#include "../include/huygens/mixer.h"
#include "../include/huygens/offline.h"

using namespace soundmath;

static double softclip(double x) { return x > 0.125 ? 0.125 : (x < -0.125 ? -0.125 : x); }

int main(int argc, char** argv) {
    const int ticks = argc > 1 ? std::atoi(argv[1]) : 64;
    try {
        Oscbank<double, 8> osc;
        RMSbank<double, 8> rms(16);
        RMSbank<double, 8> rms2(16);
        Latchbank<double, 8> latch(0.9, 0.2);        // tests/rmsbank.cpp:15
        FFilterbank<double, 4, 2> F(0.001, 1);
        // tests/slidebank.cpp:52-56,92-101: heterodyned phasors -> Slidebank low-passes -> Stickbank smoother
        std::vector<std::complex<double>> radii;
        for (int i = 0; i < 8; i++) radii.push_back(std::complex<double>(0.99 - 0.01 * i, 0.002 * i));
        Slidebank<double, 8> slide(2, radii);
        Stickbank<double, 8> smooth(1, -0.9);
        Modbank<double, 8> mod;                          // tests/harmbank.cpp: mixdown(demodulators(synthesis(), ...))
        Mixer<double, 8> mixer;
        Oscbank<double, 8>::ArrayCT amps(8);
        for (int i = 0; i < 8; i++) amps(i) = std::complex<double>((i + 1) / 8.0, 0.0);
        for (int i = 0; i < 8; i++) osc.freqmod(i, 110.0 * (i + 1));
        osc.freqmod(99, 1.0);       // ignored, oscbank.h:51
        osc.open();
        osc.deactivate({3});
        for (int i = 0; i < 4; i++) F.coefficients(i, {0.5, 0, -0.5}, {-1.6 + 0.1 * i, 0.81});
        F.open();
        F.boost(std::vector<double>(4, 1.0));
        for (int t = 0; t < ticks; t++) {
            const auto* ph = osc();
            const auto* env = rms(ph);
            const double first = (*env)(0);
            const double again = (*rms(ph))(0);              // second call in the tick: sqrt applied twice (rmsbank.h:64)
            const std::complex<double> m = osc.mixdown();
            const double y = F(m.real());
            const double yd = F(m.real(), &softclip);
            const auto* gated = latch(&rms2, ph);               // tests/harmbank.cpp:84-86
            double gsum = 0;
            for (int i = 0; i < 8; i++) gsum += (*gated)(i).real();
            const auto* sm = smooth(slide(ph));
            std::complex<double> ssum = 0;
            for (int i = 0; i < 8; i++) ssum += (*sm)(i);
            const double weighted = mixer(mod(&amps, ph));      // Mixer(Modbank(amps, osc())), src/modbank.h:41-45, src/mixer.h:32-35
            std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", m.real(), m.imag(), first, again, y, yd, gsum,
                        ssum.real(), ssum.imag(), weighted);
            osc.tick(); rms.tick(); rms2.tick(); F.tick(); slide.tick(); smooth.tick();
        }
        // block entry points of the drop-in Filterbank: fused envelope followers (replica mode) and a parameter
        // schedule at exact ticks; printed as "# name value" lines for the pytest check
        {
            const int NB = 16, L = 1200, width = 96, hop = 32;
            Filterbank<double> bank(2, NB, 0.01, 0.05);
            for (int i = 0; i < NB; i++) bank.coefficients(i, {1.0, 0.0, -1.0}, {-1.9 + 0.05 * i, 0.98});
            bank.open();
            bank.boost(std::vector<double>(NB, 1.0));
            std::vector<double> x = white_noise<double>(L), y(L), env((size_t)(L / hop) * NB);
            bank.follower_mode(HG_ENV_REPLICA);
            bank.process_env(x.data(), y.data(), env.data(), L, width, hop);
            double es = 0, ys = 0;
            for (double v : env) es += v;
            for (double v : y) ys += v;
            std::printf("# env_sum %.17g\n# env_out_sum %.17g\n", es, ys);
            Filterbank<double> bank2(2, NB, 0.01, 0.05);
            for (int i = 0; i < NB; i++) bank2.coefficients(i, {1.0, 0.0, -1.0}, {-1.9 + 0.05 * i, 0.98});
            bank2.open();
            bank2.boost(std::vector<double>(NB, 1.0));
            const double half = 0.5;
            std::vector<hg_fb_event> ev(2);
            ev[0] = hg_fb_event{300, HG_EV_MIX, 2, 1, &half, 0, nullptr, 0};
            ev[1] = hg_fb_event{700, HG_EV_BOOST, 5, 1, &half, 0, nullptr, 0};
            bank2.process(x.data(), y.data(), L, ev);
            ys = 0;
            for (double v : y) ys += v * v;
            std::printf("# events_energy %.17g\n", ys);
        }
    } catch (const cuda_error& e) {
        std::fprintf(stderr, "huygens_cuda error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
