// filterbank_capi.cu — hg_fb_* entry points (see include/huygens_cuda.h for the reference
// members each one replaces).
#include <cmath>
#include <cstdlib>
#include <limits>
#include "hg_common.cuh"
#include "filterbank_handle.cuh"

namespace {

// /root/reference/src/includes.h:38-48
double relaxation(double k) {
    const double order = std::log2(std::numeric_limits<double>::epsilon());
    if (k == 0) return 0;
    return std::pow(2.0, order / (std::fmax(0, k) * 48000));
}

template <typename T> T* par_row(hg_fb* fb, int k) { return (T*)fb->h_par + (size_t)k * fb->N; }

template <typename T>
hg_status set_rows(hg_fb* fb, int first_row, int max_rows, int n0, int count, const void* src, int n_src) {
    const T* v = (const T*)src;
    const int use = n_src < max_rows ? n_src : max_rows;
    for (int i = 0; i < count; i++)
        for (int k = 0; k < use; k++) par_row<T>(fb, first_row + k)[n0 + i] = v[(size_t)i * n_src + k];
    return HG_OK;
}

hg_status check_range(hg_fb* fb, int n0, int count, const char* what) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "%s: null handle", what);
    if (count < 0 || n0 < 0 || (long)n0 + count > fb->N)
        return hg::fail(HG_ERR_INVALID, "%s: filters [%d, %d) out of range [0, %d)", what, n0, n0 + count, fb->N);
    return HG_OK;
}

}  // namespace

extern "C" {

hg_status hg_fb_create(hg_fb** out, int dtype, int arith, int order, int n_filters, int n_streams,
                       double k_p, double k_g, int flags, int device) {
    if (!out) return hg::fail(HG_ERR_INVALID, "hg_fb_create: null out");
    *out = nullptr;
    if (dtype != HG_F32 && dtype != HG_F64) return hg::fail(HG_ERR_INVALID, "hg_fb_create: bad dtype %d", dtype);
    if (arith != HG_ARITH_STRICT && arith != HG_ARITH_FMA) return hg::fail(HG_ERR_INVALID, "hg_fb_create: bad arith %d", arith);
    if (order < 0) return hg::fail(HG_ERR_INVALID, "hg_fb_create: negative order");
    if (order > 2)
        return hg::fail(HG_ERR_UNSUPPORTED, "hg_fb_create: order %d > 2 is not built for the device (every Filterbank in the reference is order 2)", order);
    if (n_filters < 1 || n_streams < 1) return hg::fail(HG_ERR_INVALID, "hg_fb_create: need n_filters >= 1 and n_streams >= 1");
    if (hg_device_count() <= device || device < 0)
        return hg::fail(HG_ERR_CUDA, "hg_fb_create: CUDA device %d not available (%d visible); there is no CPU fallback", device, hg_device_count());
    HG_CUDA(cudaSetDevice(device));
    hg_fb* fb = new (std::nothrow) hg_fb();
    if (!fb) return hg::fail(HG_ERR_NOMEM, "hg_fb_create: host allocation failed");
    fb->dtype = dtype; fb->arith = arith; fb->order = order; fb->N = n_filters; fb->S = n_streams;
    fb->flags = flags; fb->device = device;
    int sm_count__ = 0;
    {
        const cudaError_t pe = cudaDeviceGetAttribute(&sm_count__, cudaDevAttrMultiProcessorCount, device);
        if (pe != cudaSuccess) { hg_fb_destroy(fb); return hg::fail(HG_ERR_CUDA, "hg_fb_create: cudaDeviceGetAttribute -> %s", cudaGetErrorString(pe)); }
    }
    fb->sm_count = sm_count__;
    const double sp = relaxation(k_p), sg = relaxation(k_g);
    if (dtype == HG_F32) {
        float fsp = (float)sp, fsg = (float)sg;
        fb->s_p = fsp; fb->s_g = fsg;
        if (flags & HG_FB_FIXED) { fb->c_p = 1.0f - fsp; fb->c_g = 1.0f - fsg; }      // filterbank.h:291,305-306
        else { fb->c_p = (float)(1 - sp); fb->c_g = (float)(1 - sg); }                 // filterbank.h:165,172-173
    } else {
        fb->s_p = sp; fb->s_g = sg; fb->c_p = 1 - sp; fb->c_g = 1 - sg;
    }
    const size_t es = hg::dtype_size(dtype);
    const size_t N = fb->N, S = fb->S;
    const char* e = getenv("HG_FB_FORCE_F");
    if (e) fb->force_F = atoi(e);
    e = getenv("HG_FB_WS_NF");
    if (e) fb->force_ws_nf = atoi(e);
    e = getenv("HG_FB_FORCE_CHUNKS");
    if (e) fb->force_chunks = atoi(e);
    e = getenv("HG_FB_PROF");
    if (e && atoi(e)) cudaMalloc((void**)&fb->d_prof, 4096 * 8 * sizeof(long long));
    cudaError_t ce;
    if ((ce = cudaMallocHost(&fb->h_par, 7 * N * es)) != cudaSuccess ||
        (ce = cudaMalloc(&fb->d_par, 7 * N * es)) != cudaSuccess ||
        (ce = cudaMalloc(&fb->d_state, 4 * S * N * es)) != cudaSuccess ||
        (ce = cudaMalloc(&fb->d_xh, 2 * S * 2 * es)) != cudaSuccess ||
        (ce = cudaMalloc((void**)&fb->d_sm_slot, 1024 * sizeof(int))) != cudaSuccess ) {
        hg_fb_destroy(fb);
        return hg::fail(HG_ERR_NOMEM, "hg_fb_create: allocation failed: %s", cudaGetErrorString(ce));
    }
    memset(fb->h_par, 0, 7 * N * es);                         // filterbank.h:60-69 setZero
    HG_CUDA(cudaMemset(fb->d_state, 0, 4 * S * N * es));
    HG_CUDA(cudaMemset(fb->d_xh, 0, 2 * S * 2 * es));
    fb->dirty = true;
    *out = fb;
    return HG_OK;
}

hg_status hg_fb_destroy(hg_fb* fb) {
    if (!fb) return HG_OK;
    cudaSetDevice(fb->device);
    cudaDeviceSynchronize();
    if (fb->h_par) cudaFreeHost(fb->h_par);
    if (fb->d_par) cudaFree(fb->d_par);
    if (fb->d_state) cudaFree(fb->d_state);
    if (fb->d_xh) cudaFree(fb->d_xh);
    if (fb->d_sm_slot) cudaFree(fb->d_sm_slot);
    if (fb->d_chain_ring) cudaFree(fb->d_chain_ring);
    if (fb->d_chain_credit) cudaFree(fb->d_chain_credit);
    if (fb->d_ticket) cudaFree(fb->d_ticket);
    if (fb->d_chain_error) cudaFree(fb->d_chain_error);
    if (fb->d_prof) cudaFree(fb->d_prof);
    fb->dry_stage.release();
    fb->in_stage.release(); fb->out_stage.release(); fb->bands_stage.release(); fb->partial.release();
    fb->scan_a.release(); fb->scan_b.release(); fb->env_stage.release();
    if (fb->d_rms_sum) cudaFree(fb->d_rms_sum);
    if (fb->d_rms_ring) cudaFree(fb->d_rms_ring);
    if (fb->d_lag_state) cudaFree(fb->d_lag_state);
    if (fb->d_lag_par) cudaFree(fb->d_lag_par);
    if (fb->d_xhist) cudaFree(fb->d_xhist);
    fb->xcat.release();
    fb->tv_stage.release();
    delete fb;
    return HG_OK;
}

hg_status hg_fb_set_coefficients(hg_fb* fb, int n0, int count, const void* fwd, int n_fwd, const void* back, int n_back) {
    HG_TRY(check_range(fb, n0, count, "hg_fb_set_coefficients"));
    if ((n_fwd > 0 && !fwd) || (n_back > 0 && !back) || n_fwd < 0 || n_back < 0)
        return hg::fail(HG_ERR_INVALID, "hg_fb_set_coefficients: null/negative coefficient arrays");
    if (fb->dtype == HG_F32) {
        set_rows<float>(fb, 0, fb->order + 1, n0, count, fwd, n_fwd);
        set_rows<float>(fb, 3, fb->order, n0, count, back, n_back);
    } else {
        set_rows<double>(fb, 0, fb->order + 1, n0, count, fwd, n_fwd);
        set_rows<double>(fb, 3, fb->order, n0, count, back, n_back);
    }
    fb->dirty = true;
    return HG_OK;
}

hg_status hg_fb_set_boost(hg_fb* fb, int n0, int count, const void* values) {
    HG_TRY(check_range(fb, n0, count, "hg_fb_set_boost"));
    if (count && !values) return hg::fail(HG_ERR_INVALID, "hg_fb_set_boost: null values");
    if (fb->dtype == HG_F32) set_rows<float>(fb, 5, 1, n0, count, values, 1);
    else set_rows<double>(fb, 5, 1, n0, count, values, 1);
    fb->dirty = true;
    return HG_OK;
}

hg_status hg_fb_set_mix(hg_fb* fb, int n0, int count, const void* values) {
    HG_TRY(check_range(fb, n0, count, "hg_fb_set_mix"));
    if (count && !values) return hg::fail(HG_ERR_INVALID, "hg_fb_set_mix: null values");
    if (fb->dtype == HG_F32) set_rows<float>(fb, 6, 1, n0, count, values, 1);
    else set_rows<double>(fb, 6, 1, n0, count, values, 1);
    fb->dirty = true;
    return HG_OK;
}

hg_status hg_fb_open(hg_fb* fb) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_open: null handle");
    for (int n = 0; n < fb->N; n++) {
        if (fb->dtype == HG_F32) par_row<float>(fb, 6)[n] = 1.0f;
        else par_row<double>(fb, 6)[n] = 1.0;
    }
    fb->dirty = true;
    return HG_OK;
}

hg_status hg_fb_process(hg_fb* fb, const void* in, void* out_sum, void* out_bands, size_t L, int mode,
                        int distortion, void* cuda_stream) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_process: null handle");
    if (mode != HG_MODE_AUTO && mode != HG_MODE_EXACT && mode != HG_MODE_SCAN)
        return hg::fail(HG_ERR_INVALID, "hg_fb_process: bad mode %d", mode);
    if (distortion != HG_DIST_NONE && distortion != HG_DIST_SOFTCLIP)
        return hg::fail(HG_ERR_INVALID, "hg_fb_process: bad distortion %d", distortion);
    if (distortion != HG_DIST_NONE && mode == HG_MODE_SCAN)
        return hg::fail(HG_ERR_UNSUPPORTED, "hg_fb_process: a per-band nonlinearity needs the exact mode");
    fb->cur_dist = distortion;
    fb->last_launches = 0;
    if (L == 0) return HG_OK;
    if (fb->d_rms_sum && !fb->cur_env)
        return hg::fail(HG_ERR_INVALID, "hg_fb_process: the handle's envelope followers are running and must see every tick (their window would skip this block); render through hg_fb_process_env, or call hg_fb_reset_followers first");
    if (!in) return hg::fail(HG_ERR_INVALID, "hg_fb_process: null input");
    if (!out_sum && !out_bands && !fb->cur_peer.outbox) return hg::fail(HG_ERR_INVALID, "hg_fb_process: no output requested");
    HG_CUDA(cudaSetDevice(fb->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;   // NULL = the legacy default stream, as in the CUDA runtime
    const size_t es = hg::dtype_size(fb->dtype);
    const size_t S = fb->S, N = fb->N;

    if (fb->dirty) {
        // structure of the forward taps over the whole bank (the ws kernel skips exact identities)
        bool b1_zero = true, antisym = true;
        for (size_t i = 0; i < N && b1_zero; i++) {
            double b0, b1, b2;
            if (fb->dtype == HG_F32) { b0 = par_row<float>(fb, 0)[i]; b1 = par_row<float>(fb, 1)[i]; b2 = par_row<float>(fb, 2)[i]; }
            else { b0 = par_row<double>(fb, 0)[i]; b1 = par_row<double>(fb, 1)[i]; b2 = par_row<double>(fb, 2)[i]; }
            if (b1 != 0) b1_zero = false;
            if (b2 != -b0) antisym = false;
        }
        fb->taps = b1_zero ? (antisym ? 2 : 1) : 0;
        HG_CUDA(cudaMemcpyAsync(fb->d_par, fb->h_par, 7 * N * es, cudaMemcpyHostToDevice, st));
        // the pinned shadow may be rewritten by a setter right after we return
        HG_CUDA(cudaStreamSynchronize(st));
        fb->dirty = false;
        // replica followers replay the filters `width` ticks late: the new set reaches them that much later
        if (fb->env_mode == HG_ENV_REPLICA && fb->lag_live) {
            hg_fb::LagPending pe;
            pe.tick = fb->rms_ticks + fb->rms_width;
            pe.taps = fb->taps;
            pe.par.assign((const unsigned char*)fb->h_par, (const unsigned char*)fb->h_par + 7 * N * es);
            fb->lag_pending.push_back(std::move(pe));
        }
    }
    if (fb->env_mode == HG_ENV_REPLICA && fb->cur_env && fb->d_lag_par && !fb->lag_live) {
        HG_CUDA(cudaMemcpyAsync(fb->d_lag_par, fb->d_par, 7 * N * es, cudaMemcpyDeviceToDevice, st));
        fb->lag_taps = fb->taps;
        fb->lag_live = true;
    }
    const bool in_dev = hg::is_device_ptr(in);
    const bool out_dev = out_sum ? hg::is_device_ptr(out_sum) : true;
    const bool bands_dev = out_bands ? hg::is_device_ptr(out_bands) : true;
    const void* d_in = in;
    void* d_out = out_sum;
    void* d_bands = out_bands;
    if (!in_dev) {
        HG_TRY(fb->in_stage.reserve(S * L * es));
        HG_CUDA(cudaMemcpyAsync(fb->in_stage.p, in, S * L * es, cudaMemcpyHostToDevice, st));
        d_in = fb->in_stage.p;
    }
    if (out_sum && !out_dev) {
        HG_TRY(fb->out_stage.reserve(S * L * es));
        d_out = fb->out_stage.p;
    }
    if (out_bands && !bands_dev) {
        HG_TRY(fb->bands_stage.reserve(S * N * L * es));
        d_bands = fb->bands_stage.p;
    }

    int use = mode;
    if (distortion != HG_DIST_NONE) use = HG_MODE_EXACT;
    if (use == HG_MODE_AUTO) {
        // scan only pays when the lanes cannot fill the schedulers and the block is long; it is
        // not bit-identical, so AUTO keeps fp32 on the exact path.
        const long lanes = (long)N * S;
        const bool few = lanes < (long)fb->sm_count * 4 * 32;
        use = (few && L >= 65536 && fb->dtype == HG_F64 && !out_bands) ? HG_MODE_SCAN : HG_MODE_EXACT;
    }
    if (use == HG_MODE_SCAN && out_bands)
        return hg::fail(HG_ERR_UNSUPPORTED, "hg_fb_process: scan mode renders the bank sum only");
    if (use == HG_MODE_SCAN && fb->smoothers_custom)
        return hg::fail(HG_ERR_UNSUPPORTED, "hg_fb_process: scan mode assumes the streams share their smoothed-gain state; hg_fb_set_state gave them their own");
    if (use == HG_MODE_SCAN && fb->epi.kind)
        return hg::fail(HG_ERR_UNSUPPORTED, "hg_fb_process: the fused epilogue needs the exact mode");
    fb->last_mode = use;
    hg_status rc;
    if (use == HG_MODE_EXACT) {
        rc = fb->dtype == HG_F32 ? hg::fb_process_exact<float>(fb, (const float*)d_in, (float*)d_out, (float*)d_bands, L, st)
                                 : hg::fb_process_exact<double>(fb, (const double*)d_in, (double*)d_out, (double*)d_bands, L, st);
    } else {
        rc = fb->dtype == HG_F32 ? hg::fb_process_scan<float>(fb, (const float*)d_in, (float*)d_out, L, st)
                                 : hg::fb_process_scan<double>(fb, (const double*)d_in, (double*)d_out, L, st);
    }
    HG_TRY(rc);
    bool need_sync = !in_dev;
    if (out_sum && !out_dev) { HG_CUDA(cudaMemcpyAsync(out_sum, d_out, S * L * es, cudaMemcpyDeviceToHost, st)); need_sync = true; }
    if (out_bands && !bands_dev) { HG_CUDA(cudaMemcpyAsync(out_bands, d_bands, S * N * L * es, cudaMemcpyDeviceToHost, st)); need_sync = true; }
    if (need_sync) {
        HG_CUDA(cudaStreamSynchronize(st));
        HG_TRY(hg_fb_check(fb));
    }
    return HG_OK;
}

hg_status hg_fb_check(hg_fb* fb) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_check: null handle");
    if (!fb->d_chain_error) return HG_OK;
    HG_CUDA(cudaSetDevice(fb->device));
    int flag = 0;
    HG_CUDA(cudaMemcpy(&flag, fb->d_chain_error, sizeof(int), cudaMemcpyDeviceToHost));
    if (flag) {
        cudaMemset(fb->d_chain_error, 0, sizeof(int));
        return hg::fail(HG_ERR_CUDA, "ordered mixdown: a CTA gave up waiting for the running bank-sum row of its predecessor (watchdog); the block's sum is invalid");
    }
    return HG_OK;
}

hg_status hg_fb_debug_profile(hg_fb* fb, long long* out, int n_ctas) {
    if (!fb || !out) return hg::fail(HG_ERR_INVALID, "hg_fb_debug_profile: null argument");
    if (!fb->d_prof) return hg::fail(HG_ERR_INVALID, "hg_fb_debug_profile: create the handle with HG_FB_PROF=1 in the environment");
    if (n_ctas > 4096) n_ctas = 4096;
    HG_CUDA(cudaSetDevice(fb->device));
    HG_CUDA(cudaDeviceSynchronize());
    HG_CUDA(cudaMemcpy(out, fb->d_prof, (size_t)n_ctas * 8 * sizeof(long long), cudaMemcpyDeviceToHost));
    return HG_OK;
}

hg_status hg_fb_process_voice(hg_fb* fb, const void* in, const void* dry_in, void* out_sum, size_t L, int distortion, void* cuda_stream) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_process_voice: null handle");
    if (L == 0) return HG_OK;
    if (!in || !out_sum) return hg::fail(HG_ERR_INVALID, "hg_fb_process_voice: null in/out_sum");
    HG_CUDA(cudaSetDevice(fb->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t bytes = (size_t)fb->S * L * hg::dtype_size(fb->dtype);
    fb->cur_dry_in = nullptr;
    if (dry_in && dry_in != in) {
        if (hg::is_device_ptr(dry_in)) fb->cur_dry_in = dry_in;
        else {
            HG_TRY(fb->dry_stage.reserve(bytes));
            HG_CUDA(cudaMemcpyAsync(fb->dry_stage.p, dry_in, bytes, cudaMemcpyHostToDevice, st));
            fb->cur_dry_in = fb->dry_stage.p;
        }
    }
    hg_status rc = hg_fb_process(fb, in, out_sum, nullptr, L, HG_MODE_EXACT, distortion, cuda_stream);
    if (fb->cur_dry_in == fb->dry_stage.p && rc == HG_OK) rc = cudaStreamSynchronize(st) == cudaSuccess ? HG_OK : hg::fail(HG_ERR_CUDA, "hg_fb_process_voice: stream synchronise failed");
    fb->cur_dry_in = nullptr;
    return rc;
}

hg_status hg_fb_set_epilogue(hg_fb* fb, int kind, double dry, double width) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_set_epilogue: null handle");
    if (kind != HG_EPI_NONE && kind != HG_EPI_SOFTCLIP && kind != HG_EPI_LIMITER) return hg::fail(HG_ERR_INVALID, "hg_fb_set_epilogue: bad kind %d", kind);
    if (kind == HG_EPI_SOFTCLIP && !(width > 0.0 && width < 1.0)) return hg::fail(HG_ERR_INVALID, "hg_fb_set_epilogue: softclip width must be in (0, 1)");
    fb->epi.kind = kind; fb->epi.dry = dry; fb->epi.width = width;
    return HG_OK;
}

hg_status hg_fb_set_sum_mode(hg_fb* fb, int mode) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_set_sum_mode: null handle");
    if (mode != HG_SUM_ORDERED && mode != HG_SUM_TREE) return hg::fail(HG_ERR_INVALID, "hg_fb_set_sum_mode: bad mode %d", mode);
    fb->sum_mode = mode == HG_SUM_ORDERED ? 1 : 0;
    return HG_OK;
}

hg_status hg_fb_process_env(hg_fb* fb, const void* in, void* out_sum, void* out_env, size_t L, int width, int hop,
                            void* cuda_stream) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_process_env: null handle");
    if (width < 64) return hg::fail(HG_ERR_UNSUPPORTED, "hg_fb_process_env: width %d < 64 is not built (the fused follower fetches two 32-sample tiles of delayed values ahead)", width);
    if (hop < 1) return hg::fail(HG_ERR_INVALID, "hg_fb_process_env: hop must be >= 1");
    if (fb->rms_width && fb->rms_width != width)
        return hg::fail(HG_ERR_INVALID, "hg_fb_process_env: width %d differs from the handle's follower width %d", width, fb->rms_width);
    if (L == 0) return HG_OK;
    if (!out_sum) return hg::fail(HG_ERR_INVALID, "hg_fb_process_env: out_sum is required");
    HG_CUDA(cudaSetDevice(fb->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t es = hg::dtype_size(fb->dtype), S = fb->S, N = fb->N;
    if (!fb->d_rms_sum) {
        // RMSbank(width) per (stream, filter): running sums zeroed (rmsbank.h:34-44) and either the ring of the
        // last `width` squares (zeroed) or the lagged replicas (state of the filters now, idle for `width` ticks).
        // Allocate into locals and commit only when everything succeeded: a half-initialised handle would launch
        // the follower kernels on null buffers at the next call.
        void *sum = nullptr, *lag_state = nullptr, *lag_par = nullptr, *xhist = nullptr, *ring = nullptr;
        const size_t H = (size_t)width + 4;
        const size_t ring_bytes = S * (size_t)width * ((N + 31) / 32 * 32) * es;
        cudaError_t ce = cudaMalloc(&sum, S * N * es);
        if (ce == cudaSuccess && fb->env_mode == 2) {
            if ((ce = cudaMalloc(&lag_state, 3 * S * N * es)) == cudaSuccess && (ce = cudaMalloc(&lag_par, 7 * N * es)) == cudaSuccess)
                ce = cudaMalloc(&xhist, S * H * es);
        } else if (ce == cudaSuccess) {
            ce = cudaMalloc(&ring, ring_bytes);
        }
        if (ce != cudaSuccess) {
            cudaFree(sum); cudaFree(lag_state); cudaFree(lag_par); cudaFree(xhist); cudaFree(ring);
            cudaGetLastError();
            return hg::fail(HG_ERR_NOMEM, "hg_fb_process_env: follower state (%s mode, %zu bytes of ring): %s",
                            fb->env_mode == 2 ? "replica" : "ring", fb->env_mode == 2 ? (size_t)0 : ring_bytes, cudaGetErrorString(ce));
        }
        fb->d_rms_sum = sum; fb->d_lag_state = lag_state; fb->d_lag_par = lag_par; fb->d_xhist = xhist; fb->d_rms_ring = ring;
        HG_CUDA(cudaMemsetAsync(fb->d_rms_sum, 0, S * N * es, st));
        if (fb->env_mode == 2) {
            HG_CUDA(cudaMemcpyAsync(fb->d_lag_state, fb->d_state, 3 * S * N * es, cudaMemcpyDeviceToDevice, st));
            HG_CUDA(cudaMemsetAsync(fb->d_xhist, 0, S * H * es, st));
            // the two input samples before the followers start: x[t-1], x[t-2] of every stream
            const char* xh = (const char*)fb->d_xh + (size_t)fb->xh_cur * 2 * S * es;
            HG_CUDA(cudaMemcpy2DAsync((char*)fb->d_xhist + (H - 1) * es, H * es, xh, 2 * es, es, S, cudaMemcpyDeviceToDevice, st));
            HG_CUDA(cudaMemcpy2DAsync((char*)fb->d_xhist + (H - 2) * es, H * es, xh + es, 2 * es, es, S, cudaMemcpyDeviceToDevice, st));
        } else {
            HG_CUDA(cudaMemsetAsync(fb->d_rms_ring, 0, ring_bytes, st));
        }
        fb->rms_width = width;
        fb->rms_ticks = 0;
    }
    const size_t rows = L / (size_t)hop;
    void* d_env = out_env;
    const bool env_dev = out_env ? hg::is_device_ptr(out_env) : true;
    if (out_env && !env_dev) { HG_TRY(fb->env_stage.reserve(S * rows * N * es)); d_env = fb->env_stage.p; }
    fb->cur_env = d_env ? d_env : (void*)1;       // non-null marks "followers on" even without an envelope output
    fb->cur_env_out = d_env;
    fb->cur_hop = hop;
    fb->cur_env_rows = rows;
    hg_status rc = hg_fb_process(fb, in, out_sum, nullptr, L, HG_MODE_EXACT, HG_DIST_NONE, cuda_stream);
    fb->cur_env = nullptr;
    fb->cur_env_out = nullptr;
    HG_TRY(rc);
    fb->rms_ticks += (long long)L;
    if (out_env && !env_dev) {
        HG_CUDA(cudaMemcpyAsync(out_env, d_env, S * rows * N * es, cudaMemcpyDeviceToHost, st));
        HG_CUDA(cudaStreamSynchronize(st));
    }
    return HG_OK;
}

hg_status hg_fb_process_events(hg_fb* fb, const void* in, void* out_sum, size_t L, const hg_fb_event* events,
                               size_t n_events, int mode, void* cuda_stream) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_process_events: null handle");
    if (n_events && !events) return hg::fail(HG_ERR_INVALID, "hg_fb_process_events: null events");
    if (fb->S != 1) return hg::fail(HG_ERR_UNSUPPORTED, "hg_fb_process_events: one stream per handle (the segments are sub-blocks of one row)");
    if (!in || !out_sum) return hg::fail(HG_ERR_INVALID, "hg_fb_process_events: null in/out_sum");
    const size_t es = hg::dtype_size(fb->dtype);
    size_t t = 0;
    int launches = 0;
    for (size_t e = 0; e <= n_events; e++) {
        const size_t upto = e < n_events ? events[e].tick : L;
        if (upto < t || upto > L) return hg::fail(HG_ERR_INVALID, "hg_fb_process_events: event %zu at tick %zu is out of order or past the block", e, upto);
        if (upto > t) {
            HG_TRY(hg_fb_process(fb, (const char*)in + t * es, (char*)out_sum + t * es, nullptr, upto - t, mode, HG_DIST_NONE, cuda_stream));
            launches += fb->last_launches;
            t = upto;
        }
        if (e == n_events) break;
        const hg_fb_event& ev = events[e];
        switch (ev.kind) {
            case HG_EV_COEFFICIENTS: HG_TRY(hg_fb_set_coefficients(fb, ev.n0, ev.count, ev.a, ev.na, ev.b, ev.nb)); break;
            case HG_EV_BOOST: HG_TRY(hg_fb_set_boost(fb, ev.n0, ev.count, ev.a)); break;
            case HG_EV_MIX: HG_TRY(hg_fb_set_mix(fb, ev.n0, ev.count, ev.a)); break;
            case HG_EV_OPEN: HG_TRY(hg_fb_open(fb)); break;
            default: return hg::fail(HG_ERR_INVALID, "hg_fb_process_events: bad event kind %d", ev.kind);
        }
    }
    fb->last_launches = launches;
    return HG_OK;
}

hg_status hg_fb_process_tv(hg_fb* fb, const void* in, void* out_sum, size_t L, const void* rows, size_t n_rows, int hop,
                           void* cuda_stream) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_process_tv: null handle");
    if (L == 0) return HG_OK;
    if (!in || !out_sum || !rows) return hg::fail(HG_ERR_INVALID, "hg_fb_process_tv: null in/out_sum/rows");
    if (hop < 1) return hg::fail(HG_ERR_INVALID, "hg_fb_process_tv: hop must be >= 1");
    if (n_rows * (size_t)hop < L) return hg::fail(HG_ERR_INVALID, "hg_fb_process_tv: %zu rows of %d ticks do not cover %zu ticks", n_rows, hop, L);
    HG_CUDA(cudaSetDevice(fb->device));
    cudaStream_t st = (cudaStream_t)cuda_stream;
    const size_t es = hg::dtype_size(fb->dtype), N = fb->N, row_bytes = 7 * N * es;
    const size_t used = (L + hop - 1) / hop;                  // rows that take effect inside this block
    const void* d_rows = rows;
    if (!hg::is_device_ptr(rows)) {
        HG_TRY(fb->tv_stage.reserve(used * row_bytes));
        HG_CUDA(cudaMemcpyAsync(fb->tv_stage.p, rows, used * row_bytes, cudaMemcpyHostToDevice, st));
        d_rows = fb->tv_stage.p;
    }
    fb->cur_tv = d_rows;
    fb->cur_tv_hop = hop;
    hg_status rc = hg_fb_process(fb, in, out_sum, nullptr, L, HG_MODE_EXACT, HG_DIST_NONE, cuda_stream);
    fb->cur_tv = nullptr;
    HG_TRY(rc);
    // the bank ends the block with the last row's parameters: device copy and host shadow follow
    const char* last = (const char*)d_rows + (used - 1) * row_bytes;
    HG_CUDA(cudaMemcpyAsync(fb->d_par, last, row_bytes, cudaMemcpyDeviceToDevice, st));
    HG_CUDA(cudaMemcpyAsync(fb->h_par, last, row_bytes, cudaMemcpyDeviceToHost, st));
    HG_CUDA(cudaStreamSynchronize(st));
    fb->dirty = true;                                         // re-derive the tap structure at the next plain block
    return HG_OK;
}

hg_status hg_fb_reset_followers(hg_fb* fb) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_reset_followers: null handle");
    HG_CUDA(cudaSetDevice(fb->device));
    HG_CUDA(cudaDeviceSynchronize());
    cudaFree(fb->d_rms_sum); cudaFree(fb->d_rms_ring); cudaFree(fb->d_lag_state); cudaFree(fb->d_lag_par); cudaFree(fb->d_xhist);
    fb->d_rms_sum = fb->d_rms_ring = fb->d_lag_state = fb->d_lag_par = fb->d_xhist = nullptr;
    fb->rms_width = 0; fb->rms_ticks = 0; fb->lag_live = false; fb->lag_pending.clear();
    return HG_OK;
}

hg_status hg_fb_set_follower_mode(hg_fb* fb, int mode) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_set_follower_mode: null handle");
    if (mode != HG_ENV_RING && mode != HG_ENV_REPLICA) return hg::fail(HG_ERR_INVALID, "hg_fb_set_follower_mode: bad mode %d", mode);
    if (fb->d_rms_sum && mode != fb->env_mode) return hg::fail(HG_ERR_INVALID, "hg_fb_set_follower_mode: the followers are already running");
    fb->env_mode = mode;
    return HG_OK;
}

hg_status hg_fb_get_state(hg_fb* fb, void* y1, void* y2, void* p, void* g, void* xhist) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_get_state: null handle");
    HG_CUDA(cudaSetDevice(fb->device));
    HG_CUDA(cudaDeviceSynchronize());
    HG_TRY(hg_fb_check(fb));
    const size_t es = hg::dtype_size(fb->dtype), SN = (size_t)fb->S * fb->N * es;
    void* dst[4] = {y1, y2, p, g};
    for (int k = 0; k < 4; k++)
        if (dst[k]) HG_CUDA(cudaMemcpy(dst[k], (char*)fb->d_state + k * SN, SN, cudaMemcpyDeviceToHost));
    if (xhist) HG_CUDA(cudaMemcpy(xhist, (char*)fb->d_xh + (size_t)fb->xh_cur * 2 * fb->S * es, 2 * fb->S * es, cudaMemcpyDeviceToHost));
    return HG_OK;
}

hg_status hg_fb_set_state(hg_fb* fb, const void* y1, const void* y2, const void* p, const void* g, const void* xhist) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_set_state: null handle");
    if (fb->d_rms_sum)
        return hg::fail(HG_ERR_INVALID, "hg_fb_set_state: envelope followers are running (their window / lagged copies would no longer match the filters); hg_fb_reset_followers first");
    HG_CUDA(cudaSetDevice(fb->device));
    HG_CUDA(cudaDeviceSynchronize());
    const size_t es = hg::dtype_size(fb->dtype), SN = (size_t)fb->S * fb->N * es;
    const void* src[4] = {y1, y2, p, g};
    for (int k = 0; k < 4; k++)
        if (src[k]) HG_CUDA(cudaMemcpy((char*)fb->d_state + k * SN, src[k], SN, cudaMemcpyHostToDevice));
    if ((p || g) && fb->S > 1) fb->smoothers_custom = true;
    if (xhist) HG_CUDA(cudaMemcpy((char*)fb->d_xh + (size_t)fb->xh_cur * 2 * fb->S * es, xhist, 2 * fb->S * es, cudaMemcpyHostToDevice));
    return HG_OK;
}

hg_status hg_fb_last_launch(hg_fb* fb, int* mode_used, int* launches, int* filters_per_thread, int* grid_ctas) {
    if (!fb) return hg::fail(HG_ERR_INVALID, "hg_fb_last_launch: null handle");
    if (mode_used) *mode_used = fb->last_mode;
    if (launches) *launches = fb->last_launches;
    if (filters_per_thread) *filters_per_thread = fb->last_F;
    if (grid_ctas) *grid_ctas = fb->last_ctas;
    return HG_OK;
}
}
