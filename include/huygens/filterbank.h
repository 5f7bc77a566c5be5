// filterbank.h — drop-in soundmath::Filterbank<T> / FFilterbank<T,N,order> over libhuygens_cuda.
//
// Same public members as the reference (src/filterbank.h:16-188 and :191-319): construction,
// coefficients / boost / mix / open / print, operator() / tick — plus the block entry points
// process() and process_bands() and a (frequencies, Q, gains) constructor.  All arithmetic runs
// in the CUDA library; this header only keeps a host mirror of the coefficients for print().
#ifndef HUYGENS_B200_FILTERBANK_H
#define HUYGENS_B200_FILTERBANK_H
#include "includes.h"

namespace soundmath {

template <typename T> class Filterbank {
public:
    Filterbank() {}
    ~Filterbank() { hg_fb_destroy(fb); }
    Filterbank(const Filterbank&) = delete;            // the reference double-frees on copy; we refuse
    Filterbank& operator=(const Filterbank&) = delete;
    Filterbank(Filterbank&& o) noexcept { swap(o); }
    Filterbank& operator=(Filterbank&& o) noexcept { swap(o); return *this; }

    // reference: src/filterbank.h:36
    Filterbank(int order, int N = 1, double k_p = 0.1, double k_g = 1) { init(order, N, k_p, k_g, 0, 1); }
    // new: the same bank on CUDA device `device` (one shard of a filter-partitioned bank lives on each GPU)
    Filterbank(int order, int N, double k_p, double k_g, int device) { init(order, N, k_p, k_g, 0, 1, device); }

    // new: N resonant band-passes {g,0,-g} / {-2Q cos w, Q^2} at `frequencies` (Hz), pole radius Q
    // ("Q" is the pole radius throughout the reference, src/filter.h:123-138), unity gain at the
    // centre (tests/filterbank.cpp:152-156,261-265), mixdown gains `gains`, pre-gains 1.
    Filterbank(const std::vector<T>& frequencies, T Q, const std::vector<T>& gains, double k_p = 0.1, double k_g = 1) {
        const int n = (int)frequencies.size();
        init(2, n, k_p, k_g, 0, 1);
        for (int i = 0; i < n; i++) {
            const double w = 2 * PI * (double)frequencies[i] / SR, r = (double)Q;
            const std::complex<double> z(std::cos(w), -std::sin(w));
            const double norm = 1.0 / std::abs((1.0 - z * z) / (1.0 - 2 * r * std::cos(w) * z + r * r * z * z));
            coefficients(i, {(T)norm, (T)0, (T)-norm}, {(T)(-2 * r * std::cos(w)), (T)(r * r)});
        }
        mix(gains);
        boost(std::vector<T>(n, (T)1));
    }

    // reference: src/filterbank.h:73-82
    void coefficients(int n, const std::vector<T>& forward, const std::vector<T>& back) {
        hg_check(hg_fb_set_coefficients(fb, n, 1, forward.data(), (int)forward.size(), back.data(), (int)back.size()));
        for (int i = 0; i < std::min<int>(order + 1, (int)forward.size()); i++) forwards[(size_t)n * (order + 1) + i] = forward[i];
        for (int i = 0; i < std::min<int>(order, (int)back.size()); i++) backs[(size_t)n * order + i] = back[i];
    }
    void boost(int n, T value) { hg_check(hg_fb_set_boost(fb, n, 1, &value)); }                       // :85-88
    void boost(const std::vector<T>& values) {                                                         // :91-96
        hg_check(hg_fb_set_boost(fb, 0, std::min<int>(N, (int)values.size()), values.data()));
    }
    void mix(int n, T value) { hg_check(hg_fb_set_mix(fb, n, 1, &value)); }                           // :99-102
    void mix(const std::vector<T>& values) {                                                           // :105-110
        hg_check(hg_fb_set_mix(fb, 0, std::min<int>(N, (int)values.size()), values.data()));
    }
    void open() { hg_check(hg_fb_open(fb)); }                                                          // :112-116
    void print() {                                                                                     // :118-122
        std::cout << "forwards = \n";
        for (int n = 0; n < N; n++) { for (int i = 0; i <= order; i++) std::cout << forwards[(size_t)n * (order + 1) + i] << (i < order ? " " : "\n"); }
        std::cout << "backs = \n";
        for (int n = 0; n < N; n++) { for (int i = 0; i < order; i++) std::cout << backs[(size_t)n * order + i] << (i + 1 < order ? " " : "\n"); }
        std::cout << std::endl;
    }

    // reference: src/filterbank.h:125-131 — memoised until tick(); a one-sample block on the device
    T operator()(T sample) {
        if (!computed) { compute(sample); }
        return last_sum;
    }
    // reference: src/filterbank.h:133-139 — sum_n distortion(y_n * g_n), host function pointer
    T operator()(T sample, T (*distortion)(T in)) {
        if (!computed) { compute(sample); }
        std::vector<T> g((size_t)N);
        hg_check(hg_fb_get_state(fb, nullptr, nullptr, nullptr, g.data(), nullptr));
        T acc = 0;
        for (int n = 0; n < N; n++) acc += distortion(last_bands[n] * g[n]);
        return acc;
    }
    // reference: src/filterbank.h:142-148 (the device already advanced the ring)
    void tick() { computed = false; }

    // new: block form of `out[t] = bank(in[t]); bank.tick();` — host or device pointers
    void process(const T* in, T* out_sum, size_t L, int mode = HG_MODE_EXACT, void* cuda_stream = nullptr) {
        hg_check(hg_fb_process(fb, in, out_sum, nullptr, L, mode, HG_DIST_NONE, cuda_stream));
    }
    // new: per-band outputs y_n[t], bands[n * L + t] (the reference keeps `outputs` private, :155)
    void process_bands(const T* in, T* bands, size_t L, T* out_sum = nullptr, void* cuda_stream = nullptr) {
        hg_check(hg_fb_process(fb, in, out_sum, bands, L, HG_MODE_EXACT, HG_DIST_NONE, cuda_stream));
    }
    // new: process() + one RMSbank(width) envelope follower per band (src/rmsbank.h:54-74 applied to every
    // output of compute()), fused; env[(s * rows + r) * N + n] is the follower of band n after block sample
    // (r+1)*hop - 1, rows = L / hop.  The follower state persists across calls like the filters' own.
    // new: a block with setter calls at exact ticks (audio-rate modulation); see hg_fb_event in huygens_cuda.h
    void process(const T* in, T* out_sum, size_t L, const std::vector<hg_fb_event>& events, int mode = HG_MODE_EXACT, void* cuda_stream = nullptr) {
        hg_check(hg_fb_process_events(fb, in, out_sum, L, events.data(), events.size(), mode, cuda_stream));
    }
    // new: a block whose parameters change every `hop` ticks (hop = 1: audio-rate modulation); rows[(r * 7 + k) * N + n],
    // k = b0 b1 b2 a1 a2 boost mix, row r taking effect before block tick r * hop
    void process_tv(const T* in, T* out_sum, size_t L, const T* rows, size_t n_rows, int hop = 1, void* cuda_stream = nullptr) {
        hg_check(hg_fb_process_tv(fb, in, out_sum, L, rows, n_rows, hop, cuda_stream));
    }
    // follower_mode(HG_ENV_REPLICA) before the first process_env: no ring, ~2x faster; every tick must then go through process_env
    void follower_mode(int mode) { hg_check(hg_fb_set_follower_mode(fb, mode)); }
    void process_env(const T* in, T* out_sum, T* env, size_t L, int width = SR / 20, int hop = 64, void* cuda_stream = nullptr) {
        hg_check(hg_fb_process_env(fb, in, out_sum, env, L, width, hop, cuda_stream));
    }
    // new: collective render of a filter-partitioned bank — this object holds rank hg_comm_rank(comm)'s consecutive
    // slice of the filters, every rank passes the same block; HG_SHARD_CHAIN keeps operator()'s index-order sum across
    // the GPUs (bit-identical to one bank, result on the last rank unless `root` says otherwise)
    void process_sharded(hg_comm* comm, const T* in, T* out_sum, size_t L, int shard_mode = HG_SHARD_CHAIN, int root = -1, void* cuda_stream = nullptr) {
        hg_check(hg_fb_process_sharded(fb, comm, in, out_sum, L, shard_mode, root, cuda_stream));
    }
    // new: post-processing fused into the store of the bank sum — the last step of the reference's caller loops:
    // HG_EPI_SOFTCLIP: softclip(dry * in + sum, width) (tests/filterbank.cpp:205-207); HG_EPI_LIMITER (src/wave.h:150)
    void epilogue(int kind, double dry = 0, double width = 0.9) { hg_check(hg_fb_set_epilogue(fb, kind, dry, width)); }
    // new: out[t] = epilogue(dry * dry_in[t] + bank(in[t], softclip per band)) — tests/filterbank.cpp:200-215 as a block
    void process_voice(const T* in, const T* dry_in, T* out_sum, size_t L, int distortion = HG_DIST_SOFTCLIP, void* cuda_stream = nullptr) {
        hg_check(hg_fb_process_voice(fb, in, dry_in, out_sum, L, distortion, cuda_stream));
    }
    // HG_SUM_ORDERED (default): the mixdown in the reference's index order, bit-identical; HG_SUM_TREE: per-CTA partial sums
    void sum_mode(int mode) { hg_check(hg_fb_set_sum_mode(fb, mode)); }
    int size() const { return N; }
    hg_fb* handle() { return fb; }

protected:
    void init(int order_, int N_, double k_p, double k_g, int flags, int streams, int device = 0) {
        order = order_; N = N_;
        hg_check(hg_fb_create(&fb, hg_dtype<T>::value, HUYGENS_ARITH, order, N, streams, k_p, k_g, flags, device));
        forwards.assign((size_t)N * (order + 1), 0);
        backs.assign((size_t)N * std::max(order, 1), 0);
        last_bands.assign((size_t)N, 0);
    }
    void compute(T sample) {
        hg_check(hg_fb_process(fb, &sample, &last_sum, last_bands.data(), 1, HG_MODE_EXACT, HG_DIST_NONE, nullptr));
        computed = true;
    }
    void swap(Filterbank& o) {
        std::swap(fb, o.fb); std::swap(N, o.N); std::swap(order, o.order); std::swap(computed, o.computed);
        std::swap(last_sum, o.last_sum); forwards.swap(o.forwards); backs.swap(o.backs); last_bands.swap(o.last_bands);
    }
    hg_fb* fb = nullptr;
    int N = 0, order = 0;
    bool computed = false;
    T last_sum = 0;
    std::vector<T> forwards, backs, last_bands;
};

// reference: src/filterbank.h:191-319 — same math, smoothing constants of type T (:291)
template <typename T, size_t N, size_t order> class FFilterbank : public Filterbank<T> {
public:
    FFilterbank(double k_p = 0.1, double k_g = 1) { this->init((int)order, (int)N, k_p, k_g, HG_FB_FIXED, 1); }
};

}  // namespace soundmath
#endif
