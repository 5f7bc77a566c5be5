// ref_capi.cpp — C entry points over the UNMODIFIED reference headers, compiled where they lie
// (/root/reference/src/{filterbank,oscbank,rmsbank,multichannel,includes}.h) against oracle/shim
// (stand-ins for PortAudio, RtMidi and the Eigen subset; see shim/Eigen/Dense).  Output:
// oracle/_ref/libhuygens_ref.so (git-ignored, travels to the GPU box).  TEST INFRASTRUCTURE ONLY:
// used by tests/test_oracle_vs_ref.py to pin oracle/huygens_oracle.hpp against the reference source.
// No reference source is copied into this repository.
//
// `private`/`protected` are opened so the tests can read the banks' internal state (per-band
// outputs, smoothed gains) — the reference keeps them private (src/filterbank.h:150-160).
#include <complex>
#include <cstring>
#include <iostream>
#include <sstream>
#include <vector>
#define private public
#define protected public
// included by absolute path: putting the reference's src/ on the include path would make its
// signal.h shadow <signal.h> (src/includes.h:10)
#define HG_STR2(x) #x
#define HG_STR(x) HG_STR2(x)
#include HG_STR(HG_REF_SRC/filterbank.h)
#include HG_STR(HG_REF_SRC/oscbank.h)
#include HG_STR(HG_REF_SRC/rmsbank.h)
#include HG_STR(HG_REF_SRC/latchbank.h)
#include HG_STR(HG_REF_SRC/slidebank.h)
#include HG_STR(HG_REF_SRC/stickbank.h)
#undef private
#undef protected

using namespace soundmath;

namespace {

struct FbBase {
    virtual ~FbBase() {}
    virtual void coefficients(int n, const void* f, int nf, const void* b, int nb) = 0;
    virtual void boost(int n0, int count, const void* v) = 0;
    virtual void mix(int n0, int count, const void* v) = 0;
    virtual void open() = 0;
    virtual void process(const void* in, void* out, void* bands, long L) = 0;
    virtual void get_smoothed(void* p, void* g) = 0;
};

// Bank = Filterbank<T> (pointer members) or FFilterbank<T,N,2> (value members): small accessors
template <typename T> struct DynAcc {
    typedef Filterbank<T> Bank;
    static T band(Bank& b, int n) { return (*b.outputs)(b.origin, n); }
    static T p(Bank& b, int n) { return (*b.preamps_out)(n); }
    static T g(Bank& b, int n) { return (*b.gains_out)(n); }
};
template <typename T, size_t N> struct FixAcc {
    typedef FFilterbank<T, N, 2> Bank;
    static T band(Bank& b, int n) { return b.outputs(b.origin, n); }
    static T p(Bank& b, int n) { return b.preamps_out(n); }
    static T g(Bank& b, int n) { return b.gains_out(n); }
};

template <typename T, typename Acc> struct FbImpl : FbBase {
    typename Acc::Bank* bank;
    int N;
    FbImpl(typename Acc::Bank* b, int N) : bank(b), N(N) {}
    ~FbImpl() { delete bank; }
    void coefficients(int n, const void* f, int nf, const void* b, int nb) override {
        const T* ff = (const T*)f; const T* bb = (const T*)b;
        bank->coefficients(n, std::vector<T>(ff, ff + nf), std::vector<T>(bb, bb + nb));
    }
    void boost(int n0, int count, const void* v) override { for (int i = 0; i < count; i++) bank->boost(n0 + i, ((const T*)v)[i]); }
    void mix(int n0, int count, const void* v) override { for (int i = 0; i < count; i++) bank->mix(n0 + i, ((const T*)v)[i]); }
    void open() override { bank->open(); }
    // the callers' loop: tests/resynthesis.cpp:33-42
    void process(const void* in_, void* out_, void* bands_, long L) override {
        const T* in = (const T*)in_; T* out = (T*)out_; T* bands = (T*)bands_;
        for (long t = 0; t < L; t++) {
            const T v = (*bank)(in[t]);
            if (out) out[t] = v;
            if (bands) for (int n = 0; n < N; n++) bands[(size_t)n * L + t] = Acc::band(*bank, n);
            bank->tick();
        }
    }
    void get_smoothed(void* p, void* g) override {
        for (int n = 0; n < N; n++) { ((T*)p)[n] = Acc::p(*bank, n); ((T*)g)[n] = Acc::g(*bank, n); }
    }
};

struct OscBase {
    virtual ~OscBase() {}
    virtual void freqmod(int i, double f) = 0;
    virtual void open() = 0;
    virtual void close() = 0;
    virtual void activate(const std::vector<int>& v, bool on) = 0;
    virtual int active_count() = 0;
    virtual void process(void* mix, void* phases, long L) = 0;
    virtual void get_phases(void* p) = 0;
};
template <typename T, int N> struct OscImpl : OscBase {
    Oscbank<T, N> osc;
    void freqmod(int i, double f) override { osc.freqmod(i, (T)f); }
    void open() override { osc.open(); }
    void close() override { std::stringstream sink; auto* old = std::cout.rdbuf(sink.rdbuf()); osc.close(); std::cout.rdbuf(old); }
    void activate(const std::vector<int>& v, bool on) override {
        std::stringstream sink; auto* old = std::cout.rdbuf(sink.rdbuf());      // popout() prints (multichannel.h:59)
        if (on) osc.activate(v); else osc.deactivate(v);
        std::cout.rdbuf(old);
    }
    int active_count() override { return (int)osc.where->size(); }
    void process(void* mix_, void* phases_, long L) override {
        T* mix = (T*)mix_; T* ph = (T*)phases_;
        for (long t = 0; t < L; t++) {
            if (mix) { std::complex<T> m = osc.mixdown(); mix[2 * t] = m.real(); mix[2 * t + 1] = m.imag(); }
            if (ph) { const auto* p = osc(); for (int n = 0; n < N; n++) { ph[((size_t)t * N + n) * 2] = (*p)(n).real(); ph[((size_t)t * N + n) * 2 + 1] = (*p)(n).imag(); } }
            osc.tick();
        }
    }
    void get_phases(void* p_) override {
        T* p = (T*)p_; const auto* ph = osc();
        for (int n = 0; n < N; n++) { p[2 * n] = (*ph)(n).real(); p[2 * n + 1] = (*ph)(n).imag(); }
    }
};

struct RmsBase {
    virtual ~RmsBase() {}
    virtual void process(const void* zre, const void* zim, void* out, long L, int calls) = 0;
};
template <typename T, int N> struct RmsImpl : RmsBase {
    RMSbank<T, N> rms;
    explicit RmsImpl(int width) : rms((uint)width) {}
    void process(const void* zre_, const void* zim_, void* out_, long L, int calls) override {
        const T* zre = (const T*)zre_; const T* zim = (const T*)zim_; T* out = (T*)out_;
        Eigen::Array<std::complex<T>, Eigen::Dynamic, 1> z;
        z.resize(N);
        for (long t = 0; t < L; t++) {
            for (int n = 0; n < N; n++) z(n) = std::complex<T>(zre[(size_t)n * L + t], zim ? zim[(size_t)n * L + t] : (T)0);
            const Eigen::Array<T, Eigen::Dynamic, 1>* o = nullptr;
            for (int c = 0; c < calls; c++) o = rms(&z);
            for (int n = 0; n < N; n++) out[(size_t)n * L + t] = (*o)(n);
            rms.tick();
        }
    }
};

struct LatchBase {
    virtual ~LatchBase() {}
    virtual void process(const void* zre, const void* zim, void* ore, void* oim, void* level, long L) = 0;
    virtual void process_real(const void* in, void* out, long L) = 0;
};
template <typename T, int N> struct LatchImpl : LatchBase {
    RMSbank<T, N> rms;
    Latchbank<T, N> latch;
    LatchImpl(int width, double thresh, double ratio) : rms((uint)width), latch((T)thresh, (T)ratio) {}
    // tests/harmbank.cpp:84-99: latchbank(&rmsbank, signal) ... rmsbank.tick()
    void process(const void* zre_, const void* zim_, void* ore_, void* oim_, void* level_, long L) override {
        const T* zre = (const T*)zre_; const T* zim = (const T*)zim_; T* ore = (T*)ore_; T* oim = (T*)oim_; T* lv = (T*)level_;
        Eigen::Array<std::complex<T>, Eigen::Dynamic, 1> z;
        z.resize(N);
        for (long t = 0; t < L; t++) {
            for (int n = 0; n < N; n++) z(n) = std::complex<T>(zre[(size_t)n * L + t], zim ? zim[(size_t)n * L + t] : (T)0);
            const auto* o = latch(&rms, &z);
            for (int n = 0; n < N; n++) {
                ore[(size_t)n * L + t] = (*o)(n).real();
                if (oim) oim[(size_t)n * L + t] = (*o)(n).imag();
                if (lv) lv[(size_t)n * L + t] = (*rms.out)(n);
            }
            rms.tick();
        }
    }
    void process_real(const void* in_, void* out_, long L) override {
        const T* in = (const T*)in_; T* out = (T*)out_;
        Eigen::Array<T, Eigen::Dynamic, 1> a;
        a.resize(N);
        for (long t = 0; t < L; t++) {
            for (int n = 0; n < N; n++) a(n) = in[(size_t)n * L + t];
            const auto* o = latch(&a);
            for (int n = 0; n < N; n++) out[(size_t)n * L + t] = (*o)(n);
        }
    }
};
template <typename T> LatchBase* make_latch(int N, int width, double thresh, double ratio) {
    switch (N) {
        case 1: return new LatchImpl<T, 1>(width, thresh, ratio);
        case 8: return new LatchImpl<T, 8>(width, thresh, ratio);
        case 40: return new LatchImpl<T, 40>(width, thresh, ratio);
    }
    return nullptr;
}

template <typename T> FbBase* make_fb(int fixed, int order, int N, double k_p, double k_g) {
    if (!fixed) return new FbImpl<T, DynAcc<T>>(new Filterbank<T>(order, N, k_p, k_g), N);
    if (order != 2) return nullptr;
    switch (N) {
        case 4: return new FbImpl<T, FixAcc<T, 4>>(new FFilterbank<T, 4, 2>(k_p, k_g), N);
        case 16: return new FbImpl<T, FixAcc<T, 16>>(new FFilterbank<T, 16, 2>(k_p, k_g), N);
        case 96: return new FbImpl<T, FixAcc<T, 96>>(new FFilterbank<T, 96, 2>(k_p, k_g), N);
    }
    return nullptr;
}
template <typename T> OscBase* make_osc(int N) {
    switch (N) {
        case 1: return new OscImpl<T, 1>();
        case 8: return new OscImpl<T, 8>();
        case 96: return new OscImpl<T, 96>();
    }
    return nullptr;
}
template <typename T> RmsBase* make_rms(int N, int width) {
    switch (N) {
        case 1: return new RmsImpl<T, 1>(width);
        case 8: return new RmsImpl<T, 8>(width);
        case 40: return new RmsImpl<T, 40>(width);
    }
    return nullptr;
}

// Slidebank<T,N> / Stickbank<T,N> (src/slidebank.h, src/stickbank.h), driven tick by tick like
// tests/slidebank.cpp:92-101: out = bank(&z); bank.tick().  Complex channel-major blocks [N][L].
struct CascBase {
    virtual ~CascBase() {}
    virtual void process(const void* xre, const void* xim, void* ore, void* oim, long L) = 0;
    virtual void rmod(double) {}
};
template <typename T, int N, typename Bank> struct CascImpl : CascBase {
    Bank* bank = nullptr;
    ~CascImpl() { delete bank; }
    void process(const void* xre_, const void* xim_, void* ore_, void* oim_, long L) override {
        const T* xre = (const T*)xre_; const T* xim = (const T*)xim_; T* ore = (T*)ore_; T* oim = (T*)oim_;
        Eigen::Array<std::complex<T>, Eigen::Dynamic, 1> z;
        z.resize(N);
        for (long t = 0; t < L; t++) {
            for (int n = 0; n < N; n++) z(n) = std::complex<T>(xre[(size_t)n * L + t], xim ? xim[(size_t)n * L + t] : (T)0);
            const auto* o = (*bank)(&z);
            for (int n = 0; n < N; n++) { ore[(size_t)n * L + t] = (*o)(n).real(); oim[(size_t)n * L + t] = (*o)(n).imag(); }
            bank->tick();
        }
    }
};
template <typename T, int N> CascBase* new_slide(int order, const double* rr, const double* ri) {
    std::vector<std::complex<T>> radii(N);
    for (int n = 0; n < N; n++) radii[n] = std::complex<T>((T)rr[n], (T)ri[n]);
    auto* impl = new CascImpl<T, N, Slidebank<T, N>>();
    impl->bank = new Slidebank<T, N>(order, radii);
    return impl;
}
template <typename T> CascBase* make_slide(int N, int order, const double* rr, const double* ri) {
    switch (N) {
        case 1: return new_slide<T, 1>(order, rr, ri);
        case 8: return new_slide<T, 8>(order, rr, ri);
        case 96: return new_slide<T, 96>(order, rr, ri);
    }
    return nullptr;
}
// Stickbank compiles for T = double only (stickbank.h:49-50 hands a std::vector<double> to coefficients(const std::vector<T>&))
template <int N> struct StickImpl : CascImpl<double, N, Stickbank<double, N>> {
    StickImpl(int order, double rad) { this->bank = new Stickbank<double, N>(order, rad); }
    void rmod(double r) override { this->bank->rmod(r); }
};
CascBase* make_stick(int N, int order, double rad) {
    switch (N) {
        case 1: return new StickImpl<1>(order, rad);
        case 8: return new StickImpl<8>(order, rad);
        case 96: return new StickImpl<96>(order, rad);
    }
    return nullptr;
}

}  // namespace

extern "C" {
// reference's relaxation() (src/includes.h:43-48) for the host-side designers
double hgr_relaxation(double k) { return relaxation(k); }

void* hgr_fb_create(int dtype, int fixed, int order, int N, double k_p, double k_g) {
    return dtype == 0 ? (void*)make_fb<float>(fixed, order, N, k_p, k_g) : (void*)make_fb<double>(fixed, order, N, k_p, k_g);
}
void hgr_fb_destroy(void* h) { delete (FbBase*)h; }
void hgr_fb_coefficients(void* h, int n, const void* f, int nf, const void* b, int nb) { ((FbBase*)h)->coefficients(n, f, nf, b, nb); }
void hgr_fb_boost(void* h, int n0, int count, const void* v) { ((FbBase*)h)->boost(n0, count, v); }
void hgr_fb_mix(void* h, int n0, int count, const void* v) { ((FbBase*)h)->mix(n0, count, v); }
void hgr_fb_open(void* h) { ((FbBase*)h)->open(); }
void hgr_fb_process(void* h, const void* in, void* out, void* bands, long L) { ((FbBase*)h)->process(in, out, bands, L); }
void hgr_fb_get_smoothed(void* h, void* p, void* g) { ((FbBase*)h)->get_smoothed(p, g); }

void* hgr_osc_create(int dtype, int N) { return dtype == 0 ? (void*)make_osc<float>(N) : (void*)make_osc<double>(N); }
void hgr_osc_destroy(void* h) { delete (OscBase*)h; }
void hgr_osc_freqmod(void* h, int i, double f) { ((OscBase*)h)->freqmod(i, f); }
void hgr_osc_open(void* h) { ((OscBase*)h)->open(); }
void hgr_osc_close(void* h) { ((OscBase*)h)->close(); }
void hgr_osc_activate(void* h, const int* idx, int count, int on) { ((OscBase*)h)->activate(std::vector<int>(idx, idx + count), on != 0); }
int hgr_osc_active_count(void* h) { return ((OscBase*)h)->active_count(); }
void hgr_osc_process(void* h, void* mix, void* phases, long L) { ((OscBase*)h)->process(mix, phases, L); }
void hgr_osc_get_phases(void* h, void* p) { ((OscBase*)h)->get_phases(p); }

void* hgr_rms_create(int dtype, int N, int width) { return dtype == 0 ? (void*)make_rms<float>(N, width) : (void*)make_rms<double>(N, width); }
void hgr_rms_destroy(void* h) { delete (RmsBase*)h; }
void hgr_rms_process(void* h, const void* zre, const void* zim, void* out, long L, int calls) { ((RmsBase*)h)->process(zre, zim, out, L, calls); }

void* hgr_latch_create(int dtype, int N, int width, double thresh, double ratio) {
    return dtype == 0 ? (void*)make_latch<float>(N, width, thresh, ratio) : (void*)make_latch<double>(N, width, thresh, ratio);
}
void hgr_latch_destroy(void* h) { delete (LatchBase*)h; }
void hgr_latch_process(void* h, const void* zre, const void* zim, void* ore, void* oim, void* level, long L) { ((LatchBase*)h)->process(zre, zim, ore, oim, level, L); }
void hgr_latch_process_real(void* h, const void* in, void* out, long L) { ((LatchBase*)h)->process_real(in, out, L); }

void* hgr_slide_create(int dtype, int N, int order, const double* rr, const double* ri) {
    return dtype == 0 ? (void*)make_slide<float>(N, order, rr, ri) : (void*)make_slide<double>(N, order, rr, ri);
}
void* hgr_stick_create(int dtype, int N, int order, double rad) { return dtype == 1 ? (void*)make_stick(N, order, rad) : nullptr; }
void hgr_casc_destroy(void* h) { delete (CascBase*)h; }
void hgr_casc_process(void* h, const void* xre, const void* xim, void* ore, void* oim, long L) { ((CascBase*)h)->process(xre, xim, ore, oim, L); }
void hgr_stick_rmod(void* h, double rad) { ((CascBase*)h)->rmod(rad); }
}
