// cascade.cu — Slidebank<T,N> and Stickbank<T,N> on sm_100a.
//
// Replaces /root/reference/src/slidebank.h:100-124,166-175 (operator(), tick, compute: per channel a cascade
// of `order` complex one-pole low-passes, written there as one sparse matrix applied to [input ; states]) and
// /root/reference/src/stickbank.h:82-107,181-195 (N channels through one all-pole smoother) for blocks of
// ticks.  Both are tiny, strictly serial recurrences (the reference runs them with N = 96 channels): one
// lane per channel, state in shared memory so any order fits, 32 x 32 tiles through shared memory so that
// global accesses are 128-byte rows along time.  Every product and sum is rounded as in the oracle —
// including the  alpha * rhs  (alpha = 1+0i) of Eigen's sparse x dense product, which matters for the sign
// of zeros only.
#include <cmath>
#include <vector>
#include "hg_common.cuh"

struct hg_casc {
    int kind;            // 0 = Slidebank, 1 = Stickbank
    int dtype, arith, N, order, device;
    void* d_coef = nullptr;      // Slidebank: [4][N] r.re, r.im, (1-r).re, (1-r).im;  Stickbank: [order] back + [1] scale
    void* d_state = nullptr;     // [2][order][N] re, im (Slidebank: stage values; Stickbank: y[t-1-k])
    hg::DevBuf re_stage, im_stage, ore_stage, oim_stage;
};

namespace hg {

constexpr int CASC_MAX_ORDER = 16;

template <typename T> struct CascParams {
    const T* zre;
    const T* zim;       // may be null (real input)
    T* ore;
    T* oim;
    const T* coef;
    T* state;
    int N, order;
    long L;
};

template <typename T, bool FMA>
__device__ __forceinline__ void cmul(T a, T b, T c, T d, T& re, T& im) {
    typedef Ar<T, FMA> A;
    re = A::muladd(-b, d, A::mul(a, c));      // ac - bd   (strict: two rounded products, one sum)
    im = A::muladd(b, c, A::mul(a, d));       // ad + bc
}

template <typename T, bool FMA, int KIND>
__global__ void __launch_bounds__(32) casc_kernel(CascParams<T> P) {
    typedef Ar<T, FMA> A;
    __shared__ T xr[32][33], xi[32][33], yr[32][33], yi[32][33];
    __shared__ T sr[CASC_MAX_ORDER][32], si[CASC_MAX_ORDER][32];
    const int lane = threadIdx.x;
    const int ch0 = blockIdx.x * 32, ch = ch0 + lane;
    const int N = P.N, order = P.order;
    const long L = P.L;
    const bool ok = ch < N;
    const int rows = min(32, N - ch0);
    for (int i = 0; i < order; i++) {
        sr[i][lane] = ok ? P.state[(size_t)i * N + ch] : (T)0;
        si[i][lane] = ok ? P.state[(size_t)(order + i) * N + ch] : (T)0;
    }
    T rr = 0, ri = 0, qr = 0, qi = 0, scale = 0;
    if (KIND == 0 && ok) { rr = P.coef[ch]; ri = P.coef[(size_t)N + ch]; qr = P.coef[(size_t)2 * N + ch]; qi = P.coef[(size_t)3 * N + ch]; }
    if (KIND == 1) scale = P.coef[order];
    for (long t0 = 0; t0 < L; t0 += 32) {
        const int tn = (int)min(32L, L - t0);
        for (int r = 0; r < rows; r++) {
            if (lane < tn) {
                const size_t g = (size_t)(ch0 + r) * L + t0 + lane;
                xr[r][lane] = P.zre[g];
                xi[r][lane] = P.zim ? P.zim[g] : (T)0;
            }
        }
        __syncwarp();
        if (ok) {
            for (int k = 0; k < tn; k++) {
                const T ar = xr[lane][k], ai = xi[lane][k];
                if (KIND == 0) {
                    // slidebank.h:166-175: every stage from the PREVIOUS tick's values (top stage first, in place)
                    for (int i = order - 1; i >= 0; i--) {
                        const T cr = i == 0 ? ar : sr[i - 1][lane], ci = i == 0 ? ai : si[i - 1][lane];
                        T ur, ui, pr, pi, vr, vi, wr, wi;
                        cmul<T, FMA>((T)1, (T)0, cr, ci, ur, ui);             // alpha * rhs(source column)
                        cmul<T, FMA>(qr, qi, ur, ui, pr, pi);                 // (1 - r) * ...
                        cmul<T, FMA>((T)1, (T)0, sr[i][lane], si[i][lane], vr, vi);
                        cmul<T, FMA>(rr, ri, vr, vi, wr, wi);                 // r * own state
                        sr[i][lane] = A::add(A::add((T)0, pr), wr);
                        si[i][lane] = A::add(A::add((T)0, pi), wi);
                    }
                    yr[lane][k] = sr[order - 1][lane];
                    yi[lane][k] = si[order - 1][lane];
                } else {
                    // stickbank.h:181-195: y = (1+rad)^order x - sum_k back[k] y[t-1-k], columns added in order from zero
                    T acc_r = (T)0, acc_i = (T)0;
                    for (int i = 0; i < order; i++) {
                        T pr, pi;
                        cmul<T, FMA>(sr[i][lane], si[i][lane], P.coef[i], (T)0, pr, pi);
                        acc_r = A::add(acc_r, pr);
                        acc_i = A::add(acc_i, pi);
                    }
                    const T tr = A::sub(A::mul(scale, ar), acc_r), ti = A::sub(A::mul(scale, ai), acc_i);
                    for (int i = order - 1; i > 0; i--) { sr[i][lane] = sr[i - 1][lane]; si[i][lane] = si[i - 1][lane]; }
                    sr[0][lane] = tr; si[0][lane] = ti;
                    yr[lane][k] = tr;
                    yi[lane][k] = ti;
                }
            }
        }
        __syncwarp();
        for (int r = 0; r < rows; r++) {
            if (lane < tn) {
                const size_t g = (size_t)(ch0 + r) * L + t0 + lane;
                P.ore[g] = yr[r][lane];
                P.oim[g] = yi[r][lane];
            }
        }
        __syncwarp();
    }
    if (ok)
        for (int i = 0; i < order; i++) {
            P.state[(size_t)i * N + ch] = sr[i][lane];
            P.state[(size_t)(order + i) * N + ch] = si[i][lane];
        }
}

template <typename T>
static hg_status casc_process_T(hg_casc* h, const T* zre, const T* zim, T* ore, T* oim, size_t L, cudaStream_t st) {
    CascParams<T> P;
    P.zre = zre; P.zim = zim; P.ore = ore; P.oim = oim;
    P.coef = (const T*)h->d_coef; P.state = (T*)h->d_state;
    P.N = h->N; P.order = h->order; P.L = (long)L;
    const int grid = (h->N + 31) / 32;
    const bool fma = h->arith == HG_ARITH_FMA;
    if (h->kind == 0) {
        if (fma) casc_kernel<T, true, 0><<<grid, 32, 0, st>>>(P);
        else casc_kernel<T, false, 0><<<grid, 32, 0, st>>>(P);
    } else {
        if (fma) casc_kernel<T, true, 1><<<grid, 32, 0, st>>>(P);
        else casc_kernel<T, false, 1><<<grid, 32, 0, st>>>(P);
    }
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

}  // namespace hg

namespace {

hg_status casc_alloc(hg_casc** out, int kind, int dtype, int arith, int N, int order, int device, const char* who) {
    if (!out) return hg::fail(HG_ERR_INVALID, "%s: null out", who);
    *out = nullptr;
    if (dtype != HG_F32 && dtype != HG_F64) return hg::fail(HG_ERR_INVALID, "%s: bad dtype %d", who, dtype);
    if (arith != HG_ARITH_STRICT && arith != HG_ARITH_FMA) return hg::fail(HG_ERR_INVALID, "%s: bad arith %d", who, arith);
    if (N < 1) return hg::fail(HG_ERR_INVALID, "%s: need n_channels >= 1", who);
    if (order < 1) order = 1;                       // std::max(1, order): slidebank.h:60, stickbank.h:49
    if (order > hg::CASC_MAX_ORDER) return hg::fail(HG_ERR_UNSUPPORTED, "%s: order %d > %d is not built", who, order, hg::CASC_MAX_ORDER);
    if (hg_device_count() <= device || device < 0)
        return hg::fail(HG_ERR_CUDA, "%s: CUDA device %d not available; there is no CPU fallback", who, device);
    HG_CUDA(cudaSetDevice(device));
    hg_casc* h = new (std::nothrow) hg_casc();
    if (!h) return hg::fail(HG_ERR_NOMEM, "%s: host allocation failed", who);
    h->kind = kind; h->dtype = dtype; h->arith = arith; h->N = N; h->order = order; h->device = device;
    const size_t es = hg::dtype_size(dtype);
    const size_t coef_n = kind == 0 ? (size_t)4 * N : (size_t)order + 1;
    cudaError_t ce;
    if ((ce = cudaMalloc(&h->d_coef, coef_n * es)) != cudaSuccess ||
        (ce = cudaMalloc(&h->d_state, (size_t)2 * order * N * es)) != cudaSuccess) {
        hg_casc_destroy(h);
        return hg::fail(HG_ERR_NOMEM, "%s: allocation failed: %s", who, cudaGetErrorString(ce));
    }
    HG_CUDA(cudaMemset(h->d_state, 0, (size_t)2 * order * N * es));     // inputs->setZero(), outputs->setZero()
    *out = h;
    return HG_OK;
}

// stickbank.h:198-218: coefficients of prod (z + zero_k), lowest degree first, by the reference's recursion
template <typename T> std::vector<T> monic_coefficients(const std::vector<T>& zeros, int start) {
    const int degree = (int)zeros.size() - start;
    if (degree == 0) return {0};
    if (degree == 1) return {zeros[start], 1};
    std::vector<T> total(degree + 1, 0);
    const T first = zeros[start];
    std::vector<T> shifted(monic_coefficients(zeros, start + 1));
    for (int i = 0; i < degree; i++) { total[i] += first * shifted[i]; total[i + 1] += shifted[i]; }
    return total;
}

template <typename T> hg_status stick_upload(hg_casc* h, double rad_d) {
    const T rad = (T)rad_d;
    std::vector<T> zeros(h->order, rad);
    std::vector<T> c = monic_coefficients(zeros, 0);
    std::vector<T> coef(c.begin(), c.begin() + h->order);
    coef.push_back((T)std::pow((double)((T)1 + rad), h->order));        // stickbank.h:187
    HG_CUDA(cudaMemcpy(h->d_coef, coef.data(), coef.size() * sizeof(T), cudaMemcpyHostToDevice));
    return HG_OK;
}

}  // namespace

extern "C" {

hg_status hg_slide_create(hg_casc** out, int dtype, int arith, int n_channels, int order, const double* radii, int device) {
    if (!radii) return hg::fail(HG_ERR_INVALID, "hg_slide_create: null radii");
    HG_TRY(casc_alloc(out, 0, dtype, arith, n_channels, order, device, "hg_slide_create"));
    hg_casc* h = *out;
    const size_t N = n_channels;
    hg_status rc = HG_OK;
    if (dtype == HG_F32) {
        std::vector<float> c(4 * N);
        for (size_t j = 0; j < N; j++) {
            const float rr = (float)radii[2 * j], ri = (float)radii[2 * j + 1];
            c[j] = rr; c[N + j] = ri; c[2 * N + j] = 1.0f - rr; c[3 * N + j] = 0.0f - ri;     // complex(1,0) - radii[j]  (:73)
        }
        if (cudaMemcpy(h->d_coef, c.data(), c.size() * 4, cudaMemcpyHostToDevice) != cudaSuccess) rc = HG_ERR_CUDA;
    } else {
        std::vector<double> c(4 * N);
        for (size_t j = 0; j < N; j++) {
            const double rr = radii[2 * j], ri = radii[2 * j + 1];
            c[j] = rr; c[N + j] = ri; c[2 * N + j] = 1.0 - rr; c[3 * N + j] = 0.0 - ri;
        }
        if (cudaMemcpy(h->d_coef, c.data(), c.size() * 8, cudaMemcpyHostToDevice) != cudaSuccess) rc = HG_ERR_CUDA;
    }
    if (rc != HG_OK) { hg_casc_destroy(h); *out = nullptr; return hg::fail(rc, "hg_slide_create: coefficient upload failed"); }
    return HG_OK;
}

hg_status hg_stick_create(hg_casc** out, int dtype, int arith, int n_channels, int order, double rad, int device) {
    HG_TRY(casc_alloc(out, 1, dtype, arith, n_channels, order, device, "hg_stick_create"));
    hg_status rc = hg_stick_rmod(*out, rad);
    if (rc != HG_OK) { hg_casc_destroy(*out); *out = nullptr; }
    return rc;
}

hg_status hg_stick_rmod(hg_casc* h, double rad) {
    if (!h || h->kind != 1) return hg::fail(HG_ERR_INVALID, "hg_stick_rmod: not a Stickbank handle");
    HG_CUDA(cudaSetDevice(h->device));
    HG_CUDA(cudaDeviceSynchronize());
    return h->dtype == HG_F32 ? stick_upload<float>(h, rad) : stick_upload<double>(h, rad);
}

hg_status hg_casc_destroy(hg_casc* h) {
    if (!h) return HG_OK;
    cudaSetDevice(h->device);
    cudaDeviceSynchronize();
    if (h->d_coef) cudaFree(h->d_coef);
    if (h->d_state) cudaFree(h->d_state);
    h->re_stage.release(); h->im_stage.release(); h->ore_stage.release(); h->oim_stage.release();
    delete h;
    return HG_OK;
}

hg_status hg_casc_process(hg_casc* h, const void* zre, const void* zim, void* out_re, void* out_im, size_t L, void* cuda_stream) {
    if (!h) return hg::fail(HG_ERR_INVALID, "hg_casc_process: null handle");
    if (L == 0) return HG_OK;
    if (!zre || !out_re || !out_im) return hg::fail(HG_ERR_INVALID, "hg_casc_process: null zre/out_re/out_im");
    HG_CUDA(cudaSetDevice(h->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t es = hg::dtype_size(h->dtype), bytes = (size_t)h->N * L * es;
    const void* d_re = zre; const void* d_im = zim;
    void* d_ore = out_re; void* d_oim = out_im;
    if (!hg::is_device_ptr(zre)) { HG_TRY(h->re_stage.reserve(bytes)); HG_CUDA(cudaMemcpyAsync(h->re_stage.p, zre, bytes, cudaMemcpyHostToDevice, st)); d_re = h->re_stage.p; }
    if (zim && !hg::is_device_ptr(zim)) { HG_TRY(h->im_stage.reserve(bytes)); HG_CUDA(cudaMemcpyAsync(h->im_stage.p, zim, bytes, cudaMemcpyHostToDevice, st)); d_im = h->im_stage.p; }
    const bool ore_dev = hg::is_device_ptr(out_re), oim_dev = hg::is_device_ptr(out_im);
    if (!ore_dev) { HG_TRY(h->ore_stage.reserve(bytes)); d_ore = h->ore_stage.p; }
    if (!oim_dev) { HG_TRY(h->oim_stage.reserve(bytes)); d_oim = h->oim_stage.p; }
    hg_status rc = h->dtype == HG_F32
        ? hg::casc_process_T<float>(h, (const float*)d_re, (const float*)d_im, (float*)d_ore, (float*)d_oim, L, st)
        : hg::casc_process_T<double>(h, (const double*)d_re, (const double*)d_im, (double*)d_ore, (double*)d_oim, L, st);
    HG_TRY(rc);
    if (!ore_dev) HG_CUDA(cudaMemcpyAsync(out_re, d_ore, bytes, cudaMemcpyDeviceToHost, st));
    if (!oim_dev) HG_CUDA(cudaMemcpyAsync(out_im, d_oim, bytes, cudaMemcpyDeviceToHost, st));
    if (!ore_dev || !oim_dev || !hg::is_device_ptr(zre) || (zim && !hg::is_device_ptr(zim))) HG_CUDA(cudaStreamSynchronize(st));
    return HG_OK;
}
}
