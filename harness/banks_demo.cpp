// banks_demo.cpp — exercises the drop-in Oscbank / RMSbank / FFilterbank per-sample APIs the way
// tests/rmsbank.cpp and tests/harmbank.cpp chain them (osc() -> rms(...) ; tick()), printing values
// for the pytest parity check.   banks_demo <ticks>
#include "../include/huygens/filterbank.h"
#include "../include/huygens/oscbank.h"
#include "../include/huygens/rmsbank.h"
#include "../include/huygens/latchbank.h"
#include "../include/huygens/slidebank.h"
#include "../include/huygens/stickbank.h"

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
            std::printf("%.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g %.17g\n", m.real(), m.imag(), first, again, y, yd, gsum,
                        ssum.real(), ssum.imag());
            osc.tick(); rms.tick(); rms2.tick(); F.tick(); slide.tick(); smooth.tick();
        }
    } catch (const cuda_error& e) {
        std::fprintf(stderr, "huygens_cuda error %d: %s\n", e.status, e.what());
        return 1;
    }
    return 0;
}
