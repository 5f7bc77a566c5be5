// huygens_oracle.hpp — CPU restatement of the huygens filter-ensemble hot path.
//
// TEST INFRASTRUCTURE ONLY.  This file is the parity oracle: a dependency-free, literal
// restatement of the reference's per-sample tick loops.  Only tests/, __graft_entry__.smoke()
// and bench.py's cpu_baseline / --impl reference legs may compile, link or call it.  The
// product path (huygens_b200/csrc, include/) never includes it and has no CPU fallback.
//
// PARITY STATUS: the arithmetic of the reference lives in Eigen (>= 3.4, un-vendored and
// unpinned: tests/build.sh:8 only passes -I /opt/homebrew/include/eigen3) and the reference
// ships no golden vectors, assertions or expected outputs.  The restatement is
// pinned two ways: (1) closed-form known-answer tests (tests/test_oracle_kat.py); (2)
// oracle/_ref — the UNMODIFIED reference headers src/{filterbank,oscbank,rmsbank,latchbank,
// slidebank,stickbank,multichannel}.h compiled where they lie against a minimal Eigen/PortAudio/RtMidi shim
// (oracle/shim/, ours) — must agree with this file bit-for-bit on seeded inputs
// (tests/test_oracle_vs_ref.py).  Because the shim is not Eigen itself, Eigen's internal
// summation orders stay unverified => "parity unpinned" at the last-ulp level of (a) the 3-tap
// GEMV when b1 != 0, (b) the N-way mixdown dot, (c) Stickbank's N x order GEMV for order > 1 and
// (d) the column order of Slidebank's sparse product (two terms per row: same value either way).
// Every caller in the reference uses b = {b0, 0, -b0} and Stickbank order 1, for which (a) and (c)
// are order-independent.  tests/golden/*.npz hold _ref outputs so the pin also holds on the GPU box.
//
// Arithmetic modes (template parameter FMA):
//   strict (FMA=false): every multiply and add rounded separately.  This is what
//       `g++ -O3 -std=c++17` (tests/build.sh:8, no -march) produces on x86-64, i.e. what the
//       reference computes on the GPU box's host.  Compile this TU with -ffp-contract=off.
//   fma    (FMA=true): one fixed contraction pattern with explicit std::fma (what an arm64 /
//       -march=native build would plausibly execute).  Compile with -mfma.
//
// Reference citations are relative to /root/reference.
#pragma once
#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstring>
#include <vector>

namespace hgoracle {

// src/includes.h:30-32
static const double PI = 3.14159265359;
static const int SR = 48000;

// src/includes.h:38-48  relaxation(k) = 2^(log2(eps_double) / (max(0,k) * SR)), k == 0 -> 0
inline double relaxation(double k) {
    static const double eps = std::numeric_limits<double>::epsilon();
    static const double order = std::log2(eps);
    if (k == 0) return 0;
    return std::pow(2.0, order / (std::fmax(0, k) * SR));
}

// tests/resynthesis.cpp:23-27, tests/filterbank.cpp:152-156
inline std::complex<double> transfer(double frequency, double r, double f) {
    std::complex<double> z(std::cos(2 * PI * f / SR), -std::sin(2 * PI * f / SR));
    return (1.0 - z * z) / (1.0 - 2 * r * std::cos(2 * PI * frequency / SR) * z + r * r * z * z);
}

// tests/filterbank.cpp:158-171
inline double softclip(double sample, double width) {
    if (std::abs(sample) < width) return sample;
    int sign = (0.0 < sample) - (sample < 0.0);
    double gap = sample - sign * width;
    return sign * width + (1 - width) * 2.0 / PI * std::atan(PI * gap / (2 * (1 - width)));
}
inline double softclip(double sample) { return softclip(sample, 0.125); }
// src/wave.h:150 (FUNCTIONAL lookup, wave.h:4,65-70)
inline double limiter(double x) { return 2.0 / PI * std::atan(x); }

template <typename T, bool FMA> struct Arith {
    static inline T mul(T a, T b) { return a * b; }
    // a*b + c
    static inline T muladd(T a, T b, T c) {
        if (FMA) return std::fma(a, b, c);
        T m = a * b;
        return m + c;
    }
};

// ---------------------------------------------------------------------------------------------
// Filterbank<T>  (src/filterbank.h:16-188)  and FFilterbank<T,N,order> (src/filterbank.h:191-319)
// `fixed` selects the FFilterbank flavour, whose smoothing constants are of type T (:291)
// instead of double (:165).
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA> class Filterbank {
public:
    // src/filterbank.h:36-70 (FF: :203-216)
    Filterbank(int order, int N, double k_p, double k_g, bool fixed = false)
        : N(N), order(order) {
        double sp = relaxation(k_p), sg = relaxation(k_g);
        if (fixed) {
            // FFilterbank: `T smoothing_p` (:291); `(1 - smoothing_p)` evaluated in T
            s_p = (T)sp; s_g = (T)sg;
            c_p = (T)1 - s_p; c_g = (T)1 - s_g;
        } else {
            // Filterbank: `double smoothing_p` (:165); the double scalars (1 - s) and s are
            // converted to T by Eigen's promote_scalar_arg before the multiply (:172-173)
            s_p = (T)sp; s_g = (T)sg;
            c_p = (T)(1 - sp); c_g = (T)(1 - sg);
        }
        forwards.assign((size_t)N * (order + 1), 0);
        backs.assign((size_t)N * std::max(order, 1), 0);
        input.assign(2 * (order + 1), 0);
        outputs.assign((size_t)2 * (order + 1) * N, 0);
        preamps_in.assign(N, 0); preamps_out.assign(N, 0);
        gains_in.assign(N, 0); gains_out.assign(N, 0);
    }
    // :73-82
    void coefficients(int n, const std::vector<T>& forward, const std::vector<T>& back) {
        int coeffs = std::min<int>(order + 1, (int)forward.size());
        for (int i = 0; i < coeffs; i++) forwards[(size_t)n * (order + 1) + i] = forward[i];
        coeffs = std::min<int>(order, (int)back.size());
        for (int i = 0; i < coeffs; i++) backs[(size_t)n * order + i] = back[i];
    }
    void boost(int n, T value) { preamps_in[n] = value; }                       // :85-88
    void boost(const std::vector<T>& v) {                                       // :91-96
        for (int i = 0; i < std::min<int>(N, (int)v.size()); i++) preamps_in[i] = v[i];
    }
    void mix(int n, T value) { gains_in[n] = value; }                           // :99-102
    void mix(const std::vector<T>& v) {                                         // :105-110
        for (int i = 0; i < std::min<int>(N, (int)v.size()); i++) gains_in[i] = v[i];
    }
    void open() { for (int i = 0; i < N; i++) gains_in[i] = 1; }                // :112-116

    // :125-131   outputs.row(origin) * gains_out, summed in index order
    T operator()(T sample) { return mix_from(sample, (T)0); }
    // :133-139   sum_n distortion(y_n * g_n)
    T operator()(T sample, T (*distortion)(T)) { return mix_from(sample, (T)0, distortion); }
    // The same two loops with the accumulator starting from `init` instead of 0.  Test helper, not in the
    // reference: it lets a bank be evaluated as consecutive filter partitions on several host threads (the
    // partition [a, b) continues the running sum of [0, a)) with exactly the operations of the single loop,
    // so full-size oracle sums take seconds instead of a minute.  init = 0 is the reference's operator().
    T mix_from(T sample, T init) {
        if (!computed) compute(sample);
        T acc = init;
        const T* row = &outputs[(size_t)origin * N];
        for (int n = 0; n < N; n++) acc = Arith<T, FMA>::muladd(row[n], gains_out[n], acc);
        return acc;
    }
    T mix_from(T sample, T init, T (*distortion)(T)) {
        if (!computed) compute(sample);
        T acc = init;
        const T* row = &outputs[(size_t)origin * N];
        for (int n = 0; n < N; n++) acc += distortion(row[n] * gains_out[n]);
        return acc;
    }
    // :142-148
    void tick() {
        origin--;
        if (origin < 0) origin += order + 1;
        computed = false;
    }
    // test accessor (the reference keeps `outputs` private, :155): y_n[t] of the current tick
    const T* bands() const { return &outputs[(size_t)origin * N]; }
    const T* gains() const { return gains_out.data(); }
    const T* preamps() const { return preamps_out.data(); }
    int size() const { return N; }

private:
    int N, order;
    T s_p, s_g, c_p, c_g;
    std::vector<T> forwards, backs;        // N x (order+1), N x order   (row n = filter n)
    std::vector<T> input;                  // 2(order+1) doubled ring (:51)
    std::vector<T> outputs;                // 2(order+1) x N doubled ring, row-major by ring slot (:52)
    std::vector<T> preamps_in, preamps_out, gains_in, gains_out;
    int origin = 0;
    bool computed = false;

    // :170-187 (FF :303-318)
    void compute(T sample) {
        typedef Arith<T, FMA> A;
        const int o1 = order + 1;
        for (int n = 0; n < N; n++) {   // :172-173  p = (1-s)*pin + s*p
            preamps_out[n] = A::muladd(s_p, preamps_out[n], c_p * preamps_in[n]);
            gains_out[n] = A::muladd(s_g, gains_out[n], c_g * gains_in[n]);
        }
        input[origin] = sample;                                     // :175
        input[origin + o1] = sample;                                // :176
        T* y_new = &outputs[(size_t)origin * N];
        for (int n = 0; n < N; n++) {
            // :178 GEMV forwards[N x (o+1)] * x[t-i], accumulated over i = 0..order
            const T* f = &forwards[(size_t)n * o1];
            T fw = f[0] * input[origin];
            for (int i = 1; i < o1; i++) fw = A::muladd(f[i], input[origin + i], fw);
            // :179 (backs * outputs.block(origin+1, 0, order, N)).diagonal()
            T fb = 0;
            if (order > 0) {
                const T* b = &backs[(size_t)n * order];
                fb = b[0] * outputs[(size_t)(origin + 1) * N + n];
                for (int i = 1; i < order; i++)
                    fb = A::muladd(b[i], outputs[(size_t)(origin + 1 + i) * N + n], fb);
            }
            // :178-179  temp = fw .* preamps_out - fb
            T y;
            if (FMA) y = std::fma(fw, preamps_out[n], -fb);
            else { T u = fw * preamps_out[n]; y = u - fb; }
            y_new[n] = y;
        }
        std::memcpy(&outputs[(size_t)(origin + o1) * N], y_new, sizeof(T) * N);   // :183-184
        computed = true;
    }
};

// ---------------------------------------------------------------------------------------------
// Multichannel<T,N>  (src/multichannel.h:16-159): sorted active-index list + bool mask.
// The recursive binary-search insert/erase (:121-157) is a sorted-set insert/erase; restated
// as such.  All channels start inactive (:19-24).
// ---------------------------------------------------------------------------------------------
class Multichannel {
public:
    explicit Multichannel(int N) : N(N), active(N, 0) {}
    void activate(const std::vector<int>& idx) {                               // :87-92
        for (int i : idx)
            if (0 <= i && i < N && !active[i]) {
                where.insert(std::lower_bound(where.begin(), where.end(), i), i);
                active[i] = 1;
            }
    }
    void deactivate(const std::vector<int>& idx) {                             // :95-100
        for (int i : idx)
            if (0 <= i && i < N && active[i]) {
                where.erase(std::lower_bound(where.begin(), where.end(), i));
                active[i] = 0;
            }
    }
    void open() { std::vector<int> v(N); for (int i = 0; i < N; i++) v[i] = i; activate(v); }    // :103-109
    void close() { std::vector<int> v(N); for (int i = 0; i < N; i++) v[i] = i; deactivate(v); } // :112-118
    int N;
    std::vector<char> active;
    std::vector<int> where;
};

// ---------------------------------------------------------------------------------------------
// Oscbank<T,N>  (src/oscbank.h:16-97)
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA> class Oscbank : public Multichannel {
public:
    explicit Oscbank(int N) : Multichannel(N), re(N, 1), im(N, 0), fr(N, 1), fi(N, 0) {}   // :37-47
    // :49-56
    void freqmod(int index, T target) {
        if (0 <= index && index < N) {
            fr[index] = (T)std::cos(2 * PI * target / SR);
            fi[index] = (T)std::sin(2 * PI * target / SR);
        }
    }
    // :59-63   p *= f ;  p /= (1 + |p|^2) / 2     over the active set
    void tick() {
        for (int i : where) {
            T a = re[i], b = im[i], c = fr[i], d = fi[i];
            T nr, ni, q;
            if (FMA) {
                nr = std::fma(-b, d, a * c);
                ni = std::fma(b, c, a * d);
                q = std::fma(ni, ni, nr * nr);
            } else {
                T ac = a * c, bd = b * d, ad = a * d, bc = b * c;
                nr = ac - bd;
                ni = ad + bc;
                T r2 = nr * nr, i2 = ni * ni;
                q = r2 + i2;
            }
            T m = ((T)1 + q) / (T)2;
            re[i] = nr / m;
            im[i] = ni / m;
        }
    }
    // :81-90
    std::complex<T> mixdown() const {
        T sr = 0, si = 0;
        for (int i : where) { sr += re[i]; si += im[i]; }
        return std::complex<T>(sr, si);
    }
    std::vector<T> re, im, fr, fi;
};

// ---------------------------------------------------------------------------------------------
// RMSbank<T,N>  (src/rmsbank.h:16-85)
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA> class RMSbank : public Multichannel {
public:
    RMSbank(int N, int width) : Multichannel(N), width(width), origin(0),
        inputs((size_t)N * 2 * (width + 1), 0), out(N, 0), lastout(N, 0) {}     // :32-45
    // :54-66   input: N complex samples (re[], im[]).  Returns out (in place sqrt, every call)
    const T* operator()(const T* zre, const T* zim) {
        const size_t W2 = 2 * (size_t)(width + 1);
        if (!computed) {
            for (int n = 0; n < N; n++) {
                T q;
                if (FMA) q = std::fma(zim[n], zim[n], zre[n] * zre[n]);
                else { T a = zre[n] * zre[n], b = zim[n] * zim[n]; q = a + b; }
                T* ring = &inputs[(size_t)n * W2];
                ring[origin] = q;                                                // :58
                ring[origin + (width + 1)] = q;                                  // :59
                T d = ring[origin] - ring[origin + width];                       // :60
                out[n] = d + lastout[n];
                lastout[n] = out[n];                                             // :61
            }
            computed = true;
        }
        for (int n = 0; n < N; n++) out[n] = std::sqrt(out[n] / (T)width);       // :64
        return out.data();
    }
    // :68-74
    void tick() {
        origin--;
        if (origin < 0) origin += width + 1;
        computed = false;
    }
    int width, origin;
    bool computed = false;
    std::vector<T> inputs, out, lastout;
};

// ---------------------------------------------------------------------------------------------
// Latchbank<T,N>  (src/latchbank.h:16-105): hysteresis gate driven by a level (an RMSbank output).
// Per channel: armed |= in < thresh*ratio; a channel that is engaged and falls below thresh*ratio
// disengages and disarms; a channel that is not engaged, is armed and rises above thresh*(1-ratio)
// engages.  Output = input (or the complex signal) times the 0/1 `engaged` flag.
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA> class Latchbank : public Multichannel {
public:
    Latchbank(int N, T thresh, T ratio) : Multichannel(N), thresh(thresh), ratio(ratio), armed(N, 0), engaged(N, 0), out(N, 0) {}   // :34-53
    // :61-79
    const T* operator()(const T* input) {
        step(input);
        for (int n = 0; n < N; n++) out[n] = input[n] * (T)engaged[n];
        return out.data();
    }
    // :81-100  level = rms(signal) (first evaluation of the tick, in-place sqrt included); gated complex out
    void operator()(RMSbank<T, FMA>& rms, const T* zre, const T* zim, T* ore, T* oim) {
        const T* level = rms(zre, zim);
        step(level);
        for (int n = 0; n < N; n++) { ore[n] = zre[n] * (T)engaged[n]; oim[n] = zim[n] * (T)engaged[n]; }
    }
    std::vector<char> armed, engaged;
    T thresh, ratio;

private:
    std::vector<T> out;
    void step(const T* in) {
        const T lo = thresh * ratio, hi = thresh * ((T)1 - ratio);
        for (int n = 0; n < N; n++) {
            armed[n] = armed[n] || (in[n] < lo);                                   // :63 / :85
            const bool t1 = engaged[n] && (in[n] < lo);                            // :65 / :87
            const bool t2 = !engaged[n] && (in[n] > hi) && armed[n];               // :66 / :88
            engaged[n] = engaged[n] && !t1;                                        // :69 / :91
            armed[n] = armed[n] && !t1;                                            // :70 / :92
            engaged[n] = engaged[n] || t2;                                         // :73 / :95
        }
    }
};

// ---------------------------------------------------------------------------------------------
// Complex products as g++ -O3 evaluates std::complex<T> operator* on finite values:
// (a+ib)(c+id) = (ac - bd) + i(ad + bc); the fma mode contracts like Oscbank::tick above.
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA> inline void cmul(T a, T b, T c, T d, T& re, T& im) {
    if (FMA) { re = std::fma(-b, d, a * c); im = std::fma(b, c, a * d); }
    else { T ac = a * c, bd = b * d, ad = a * d, bc = b * c; re = ac - bd; im = ad + bc; }
}

// ---------------------------------------------------------------------------------------------
// Slidebank<T,N>  (src/slidebank.h:19-186): per channel a cascade of `order` complex one-pole
// low-passes sharing one radius r, written in the reference as ONE sparse matrix `back` applied to
// the stacked vector [input ; states]: every stage reads the PREVIOUS tick's value of the stage
// before it (one tick of delay per stage).  Row (i + j*order) holds (1-r_j) at the column of its
// source and r_j at the column of its own state (:71-83); Eigen's column-major sparse x dense
// product visits columns in ascending order and does  res += value * (alpha * rhs_j), alpha = 1:
//   s_i' = (0 + (1-r) * (1*src)) + r * (1*s_i),   src = input (i = 0) or s_{i-1};   out = s_{order-1}'
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA> class Slidebank : public Multichannel {
public:
    Slidebank(int N, int order, const T* rad_re, const T* rad_im)                       // :53-97 setup()
        : Multichannel(N), order(std::max(1, order)), rr(rad_re, rad_re + N), ri(rad_im, rad_im + N),
          qr(N), qi(N), sr((size_t)N * this->order, 0), si((size_t)N * this->order, 0), nr(this->order), ni(this->order) {
        for (int j = 0; j < N; j++) { qr[j] = (T)1 - rr[j]; qi[j] = (T)0 - ri[j]; }      // complex(1,0) - radii[j]
    }
    // :100-106 + :166-175   one tick: input x (complex), output = last stage, all channels
    void operator()(const T* xre, const T* xim, T* ore, T* oim) {
        for (int j = 0; j < N; j++) {
            T* s_r = &sr[(size_t)j * order]; T* s_i = &si[(size_t)j * order];
            for (int i = 0; i < order; i++) {
                const T ar = i == 0 ? xre[j] : s_r[i - 1], ai = i == 0 ? (xim ? xim[j] : (T)0) : s_i[i - 1];
                T ur, ui, vr, vi, pr, pi, wr, wi;
                cmul<T, FMA>((T)1, (T)0, ar, ai, ur, ui);               // alpha * rhs(source column)
                cmul<T, FMA>(qr[j], qi[j], ur, ui, pr, pi);             // (1 - r) * ...
                T acc_r = (T)0 + pr, acc_i = (T)0 + pi;
                cmul<T, FMA>((T)1, (T)0, s_r[i], s_i[i], vr, vi);       // alpha * rhs(own column)
                cmul<T, FMA>(rr[j], ri[j], vr, vi, wr, wi);             // r * ...
                nr[i] = acc_r + wr; ni[i] = acc_i + wi;
            }
            for (int i = 0; i < order; i++) { s_r[i] = nr[i]; s_i[i] = ni[i]; }
            ore[j] = s_r[order - 1]; oim[j] = s_i[order - 1];
        }
    }
    int order;
    std::vector<T> rr, ri, qr, qi, sr, si;
private:
    std::vector<T> nr, ni;
};

// ---------------------------------------------------------------------------------------------
// Stickbank<T,N>  (src/stickbank.h:15-221): N channels through the same all-pole filter whose
// denominator has `order` roots at `rad`:  y[t] = (1+rad)^order x[t] - sum_k back[k] y[t-1-k]
// (:181-195); back = the low `order` coefficients of prod (z + rad) from the recursive
// coefficients() (:198-218).  The N x order complex GEMV adds its columns one after the other from
// zero (our reading of Eigen's order for order > 1; every caller in the reference uses order 1).
// ---------------------------------------------------------------------------------------------
template <typename T, bool FMA> class Stickbank : public Multichannel {
public:
    Stickbank(int N, int order, T rad) : Multichannel(N), order(std::max(1, order)), yr((size_t)N * this->order, 0),
                                         yi((size_t)N * this->order, 0) { rmod(rad); }
    void rmod(T rad_) {                                                                  // :47-58, :69-79
        rad = rad_;
        std::vector<T> zeros(order, rad);
        std::vector<T> c = coefficients(zeros, 0);
        back.assign(c.begin(), c.begin() + order);
        scale = (T)std::pow((double)((T)1 + rad), order);
    }
    void operator()(const T* xre, const T* xim, T* ore, T* oim) {                         // :82-89 + :181-195
        for (int j = 0; j < N; j++) {
            T* h_r = &yr[(size_t)j * order]; T* h_i = &yi[(size_t)j * order];            // h[k] = y[t-1-k]
            T acc_r = 0, acc_i = 0;
            for (int k = 0; k < order; k++) {
                T pr, pi;
                cmul<T, FMA>(h_r[k], h_i[k], back[k], (T)0, pr, pi);
                acc_r = acc_r + pr; acc_i = acc_i + pi;
            }
            const T xi = xim ? xim[j] : (T)0;
            const T tr = scale * xre[j] - acc_r, ti = scale * xi - acc_i;
            for (int k = order - 1; k > 0; k--) { h_r[k] = h_r[k - 1]; h_i[k] = h_i[k - 1]; }
            h_r[0] = tr; h_i[0] = ti;
            ore[j] = tr; oim[j] = ti;
        }
    }
    // :198-218  coefficients of the monic polynomial with the given roots' negatives, lowest degree first
    static std::vector<T> coefficients(const std::vector<T>& zeros, int start) {
        const int degree = (int)zeros.size() - start;
        if (degree == 0) return {0};
        if (degree == 1) return {zeros[start], 1};
        std::vector<T> total(degree + 1, 0);
        const T first = zeros[start];
        std::vector<T> shifted(coefficients(zeros, start + 1));
        for (int i = 0; i < degree; i++) { total[i] += first * shifted[i]; total[i + 1] += shifted[i]; }
        return total;
    }
    int order;
    T rad, scale;
    std::vector<T> back, yr, yi;
};

}  // namespace hgoracle
