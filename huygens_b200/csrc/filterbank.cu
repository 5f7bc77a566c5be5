// filterbank.cu — Filterbank<T> / FFilterbank<T,N,order> tick loop on sm_100a.
//
// Replaces /root/reference/src/filterbank.h:170-187 (compute), :125-139 (operator()), :142-148
// (tick) for whole blocks of samples.  Two time strategies:
//   K1 "exact": one register-resident lane per (stream, filter); serial in time; every operation
//       of compute() is replayed with the oracle's rounding sequence, so y_n[t] is bit-identical
//       to the reference loop.  Only the N-way mixdown order differs (warp/CTA tree).
//   K2 "scan":  see filterbank_scan.cu.
#include <cuda_pipeline.h>
#include "hg_common.cuh"
#include "filterbank_handle.cuh"

namespace hg {

// ---------------------------------------------------------------------------------------------
// K1 kernel
//   grid  = (filter groups, streams); CTA = NT threads; thread owns F filters
//           n = group*NT*F + k*NT + tid  (k < F) of one stream.
//   smem  = input chunk xs[2][XT] (cp.async double buffer, broadcast reads) +
//           partial sums red[2][TILE][NT+NW] (transposed CTA reduction, conflict-free).
//   The smoothed gains p, g are one-pole recursions towards constant targets; once they reach
//   their floating-point fixed point (p == fl(fl(c*pin) + fl(s*p))) they never change again, so
//   the CTA drops to a loop without them.  That is bit-exact, not an approximation.
// ---------------------------------------------------------------------------------------------
constexpr int K1_TILE = 32;
constexpr int K1_XT = 1024;
constexpr int K1_TVS = 16;     // ticks per staging step of the time-varying parameter rows

template <typename T> struct Vec4 { T v[4]; };

template <typename T> __device__ __forceinline__ void lds4(const T* p, T (&x)[4]);
template <> __device__ __forceinline__ void lds4<float>(const float* p, float (&x)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
template <> __device__ __forceinline__ void lds4<double>(const double* p, double (&x)[4]) {
    double2 a = *reinterpret_cast<const double2*>(p);
    double2 b = *reinterpret_cast<const double2*>(p + 2);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
template <typename T> __device__ __forceinline__ void stg4(T* p, const T (&y)[4]);
template <> __device__ __forceinline__ void stg4<float>(float* p, const float (&y)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(y[0], y[1], y[2], y[3]);
}
template <> __device__ __forceinline__ void stg4<double>(double* p, const double (&y)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(y[0], y[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(y[2], y[3]);
}

template <typename T, int F> struct Lane {
    T b0[F], b1[F], b2[F], a1[F], a2[F];
    T y1[F], y2[F], p[F], g[F], cp[F], cg[F];
};

// one sample for the F filters of this thread.  The mixdown terms go to v[] (and the gains to gm[] in fma mode,
// where the sum warp fuses y*g + acc): v = fl(y*g), or distortion(fl(y*g)) (:133-139); fma without distortion: v = y
template <typename T, bool FMA, int F, bool SMOOTH, bool BANDS>
__device__ __forceinline__ void k1_sample(Lane<T, F>& r, T x0, T x1, T x2, T s_p, T s_g, T (&yb)[F], T (&v)[F], T (&gm)[F], int dist) {
    typedef Ar<T, FMA> A;
#pragma unroll
    for (int f = 0; f < F; f++) {
        if (SMOOTH) {                                    // filterbank.h:172-173
            r.p[f] = A::muladd(s_p, r.p[f], r.cp[f]);
            r.g[f] = A::muladd(s_g, r.g[f], r.cg[f]);
        }
        T fw = A::mul(r.b0[f], x0);                      // :178  GEMV over the 3 taps
        fw = A::muladd(r.b1[f], x1, fw);
        fw = A::muladd(r.b2[f], x2, fw);
        T fb = A::mul(r.a1[f], r.y1[f]);                 // :179  feedback dot
        fb = A::muladd(r.a2[f], r.y2[f], fb);
        T y = A::mulsub(fw, r.p[f], fb);                 // :178-179  fw .* p - fb
        r.y2[f] = r.y1[f];
        r.y1[f] = y;
        if (BANDS) yb[f] = y;
        if (dist) v[f] = softclip_dev<T>(A::mul(y, r.g[f]));
        else if (FMA) { v[f] = y; gm[f] = r.g[f]; }
        else v[f] = A::mul(y, r.g[f]);
    }
}

// CTA = NT compute threads (thread = F filters) + one sum warp.  The compute threads leave the mixdown terms of a
// 32-sample tile in shared memory, vt[t][j] with j = filter index within the CTA; the sum warp (lane = sample) adds
// them in index order onto the running row of the chain (chain.cuh) while the compute threads render the next tile.
template <typename T, bool FMA, int F, int NT, bool BANDS>
__global__ void __launch_bounds__(NT + 32) fb_exact_kernel(FbK1Params<T> P) {
    typedef Ar<T, FMA> A;
    constexpr int NF = NT * F;
    constexpr int RS = NF + 1;           // row stride == 1 (mod 32): the per-sample STS of all threads and the sum warp's column walk are conflict-free
    extern __shared__ __align__(16) unsigned char smem_raw[];
    T* xs = reinterpret_cast<T*>(smem_raw);          // [2][XT]
    T* vt = xs + 2 * K1_XT;                           // [2][TILE][RS]  mixdown terms (fma: y)
    T* gt = vt + 2 * K1_TILE * RS;                    // [2][TILE][RS]  fma: gains
    T* tvs = gt + (FMA ? 2 * K1_TILE * RS : 0);       // [K1_TVS + 1][7][NT * F]: staged parameter rows (time-varying, hop < K1_TVS)

    const int tid = threadIdx.x;
    const int G = gridDim.x;
    const unsigned ticket = chain::take_ticket(P.chain);
    const int grp = (int)(ticket % (unsigned)G), s = (int)(ticket / (unsigned)G);
    const int N = P.N;
    const long L = P.L;
    const int n_base = grp * NF;
    const long nchunks = (L + K1_XT - 1) / K1_XT;
    const int dist = P.dist;

    if (tid >= NT) {
        // ---------------- sum warp ----------------
        const int lane = tid - NT;
        const chain::Params& C = P.chain;
        const bool summing = P.out != nullptr || C.outbox != nullptr;
        const bool ordered = C.ordered != 0;
        const bool head = !ordered || grp == 0, tail = !ordered || grp == G - 1;
        constexpr int WPS = chain::Words<T>::WPS;
        chain::Port<T, 1> pin, pout;
        if (ordered && summing) {
            if (!head) pin.open(C.ring + (((size_t)s * G + grp) << C.ring_log2) * WPS, C.credit + (size_t)s * G + grp, C);
            else if (C.inbox) pin.open(C.inbox + ((size_t)s << C.ring_log2) * WPS, C.inbox_credit + s, C);
            if (!tail) pout.open(C.ring + (((size_t)s * G + grp + 1) << C.ring_log2) * WPS, C.credit + (size_t)s * G + grp + 1, C);
            else if (C.outbox) pout.open(C.outbox + ((size_t)s << C.ring_log2) * WPS, C.outbox_credit + s, C);
        }
        T* out_row = nullptr;
        if (P.out && tail && !pout.on())
            out_row = (ordered || P.direct) ? (P.out + (size_t)s * P.out_stride) : (P.out + ((size_t)grp * P.S + s) * (size_t)L);
        const T* carry = (ordered && head && !pin.on() && P.carry_in) ? P.carry_in + (size_t)s * P.carry_stride : nullptr;
        const int cols = min(NF, N - n_base);                 // filters of this CTA
        typename chain::Port<T, 1>::Raw raw;
        bool issued = false;
        if (pin.on() && lane < L) { pin.issue(C, lane, raw); issued = true; }
        long tile_idx = 0;
        for (long c = 0; c < nchunks; c++) {
            __syncthreads_and(1);
            const int cnt = (int)min((long)K1_XT, L - c * K1_XT);
            for (int t0 = 0; t0 < cnt; t0 += K1_TILE, tile_idx++) {
                __syncthreads_and(1);                         // the compute threads have finished this tile
                const int tn = min(K1_TILE, cnt - t0);
                const long tg = c * K1_XT + t0;
                if (!summing) continue;
                T acc = (T)0;
                if (pin.on()) {
                    T a1v[1] = {(T)0};
                    if (tg + lane < L) pin.recv(C, tg + lane, a1v, raw, issued);
                    acc = a1v[0];
                    issued = false;
                    if (tg + K1_TILE + lane < L) { pin.issue(C, tg + K1_TILE + lane, raw); issued = true; }
                } else if (carry) {
                    if (lane < tn) acc = carry[tg + lane];
                }
                const T* vr = vt + (tile_idx & 1) * (K1_TILE * RS) + lane * RS;
                const T* gr = gt + (tile_idx & 1) * (K1_TILE * RS) + lane * RS;
                if (FMA && !dist) {
#pragma unroll 8
                    for (int j = 0; j < cols; j++) acc = A::muladd(vr[j], gr[j], acc);        // :130
                } else {
#pragma unroll 8
                    for (int j = 0; j < cols; j++) acc = A::add(vr[j], acc);                   // :130 / :137
                }
                if (pout.on()) {
                    pout.reserve(C, tg + K1_TILE);
                    T a1v[1] = {acc};
                    pout.send(C, tg + lane, a1v);
                } else if (out_row) {
                    if (lane < tn) {
                        if (P.epi.kind && ordered) acc = epilogue_dev<T>(P.epi, acc, P.dry_in[(size_t)s * P.in_stride + tg + lane]);
                        out_row[tg + lane] = acc;
                    }
                }
                if (pin.on()) {
                    __syncwarp();
                    if (lane == 0) pin.release(C, tg + K1_TILE);
                }
            }
        }
        return;
    }

    Lane<T, F> r;
    const size_t SN = (size_t)P.S * N;
    bool conv = true;
#pragma unroll
    for (int f = 0; f < F; f++) {
        const int n = n_base + f * NT + tid;
        const bool ok = n < N;
        r.b0[f] = ok ? P.par[0 * (size_t)N + n] : (T)0;
        r.b1[f] = ok ? P.par[1 * (size_t)N + n] : (T)0;
        r.b2[f] = ok ? P.par[2 * (size_t)N + n] : (T)0;
        r.a1[f] = ok ? P.par[3 * (size_t)N + n] : (T)0;
        r.a2[f] = ok ? P.par[4 * (size_t)N + n] : (T)0;
        const T pin = ok ? P.par[5 * (size_t)N + n] : (T)0;
        const T gin = ok ? P.par[6 * (size_t)N + n] : (T)0;
        const size_t si = (size_t)s * N + n;
        r.y1[f] = ok ? P.state[0 * SN + si] : (T)0;
        r.y2[f] = ok ? P.state[1 * SN + si] : (T)0;
        r.p[f] = ok ? P.state[2 * SN + si] : (T)0;
        r.g[f] = ok ? P.state[3 * SN + si] : (T)0;
        r.cp[f] = A::mul(P.c_p, pin);                    // fl((1-s)*pin) is loop invariant
        r.cg[f] = A::mul(P.c_g, gin);
        conv = conv && (A::muladd(P.s_p, r.p[f], r.cp[f]) == r.p[f]) &&
               (A::muladd(P.s_g, r.g[f], r.cg[f]) == r.g[f]);
    }
    T x1 = P.xh[2 * s], x2 = P.xh[2 * s + 1];
    // time-varying parameters: ticks left until the next row takes over, and that row's index
    const bool tv = P.tv_rows != nullptr;
    // the setters of the caller loop, all filters at once: coefficients(), boost(), mix() (filterbank.h:73-110)
    auto tv_load = [&](long rowidx) {
        const T* row = P.tv_rows + (size_t)rowidx * 7 * N;
#pragma unroll
        for (int f = 0; f < F; f++) {
            const int n = n_base + f * NT + tid;
            if (n < N) {
                r.b0[f] = row[0 * (size_t)N + n]; r.b1[f] = row[1 * (size_t)N + n]; r.b2[f] = row[2 * (size_t)N + n];
                r.a1[f] = row[3 * (size_t)N + n]; r.a2[f] = row[4 * (size_t)N + n];
                r.cp[f] = A::mul(P.c_p, row[5 * (size_t)N + n]);
                r.cg[f] = A::mul(P.c_g, row[6 * (size_t)N + n]);
            }
        }
    };
    long tv_row = 0, tv_left = -1;
    if (tv) {
        tv_row = P.tv_t0 / P.tv_hop;
        const long into = P.tv_t0 - tv_row * P.tv_hop;          // ticks of that row already rendered by earlier launches
        if (into > 0) { tv_load(tv_row); tv_row++; tv_left = P.tv_hop - into; }
        else tv_left = 0;                                       // a row starts at this launch's first tick
        conv = false;                                           // targets move: keep the smoothers on the per-sample path
    }

    const T* inp = P.in + (size_t)s * P.in_stride;
    auto stage = [&](long c) {
        const long base = c * K1_XT;
        const int cnt = (int)min((long)K1_XT, L - base);
        T* dst = xs + (c & 1) * K1_XT;
        for (int i = tid; i < cnt; i += NT) __pipeline_memcpy_async(dst + i, inp + base + i, sizeof(T));
        __pipeline_commit();
    };
    if (nchunks > 0) stage(0);

    long tile_idx = 0;
    int all_conv = 0;
    T ybs[F], vv[F], gg[F];
    // one sample: render, leave the mixdown terms in the tile, stream the band rows
    auto emit = [&](T* vrow, T* grow, long tglob) {
#pragma unroll
        for (int f = 0; f < F; f++) {
            vrow[f * NT + tid] = vv[f];
            if (FMA) grow[f * NT + tid] = gg[f];
            if (BANDS) {
                const int n = n_base + f * NT + tid;
                if (n < N) P.bands[((size_t)s * N + n) * P.bands_stride + tglob] = ybs[f];
            }
        }
    };
    for (long c = 0; c < nchunks; c++) {
        if (c + 1 < nchunks) {
            stage(c + 1);
            __pipeline_wait_prior(1);
        } else {
            __pipeline_wait_prior(0);
        }
        all_conv = __syncthreads_and(conv);
        const int cnt = (int)min((long)K1_XT, L - c * K1_XT);
        const T* xc = xs + (c & 1) * K1_XT;
        for (int t0 = 0; t0 < cnt; t0 += K1_TILE, tile_idx++) {
            const int tn = min(K1_TILE, cnt - t0);
            T* vb = vt + (tile_idx & 1) * (K1_TILE * RS);
            T* gb = gt + (tile_idx & 1) * (K1_TILE * RS);
            const long tg = c * K1_XT + t0;          // time index of the tile start in this block
            int tt = 0;
            if (all_conv) {
                for (; tt < tn; tt++) {
                    const T x0 = xc[t0 + tt];
                    k1_sample<T, FMA, F, false, BANDS>(r, x0, x1, x2, P.s_p, P.s_g, ybs, vv, gg, dist);
                    x2 = x1; x1 = x0;
                    emit(vb + tt * RS, gb + tt * RS, tg + tt);
                }
            } else if (tv && P.tv_hop < K1_TVS) {
                // fast parameter modulation: the rows that take effect in the next K1_TVS ticks are copied to shared
                // memory lane by lane (coalesced, all loads in flight at once); lanes then pick their row from
                // there at each boundary instead of waiting for seven dependent global loads per tick
                for (int s0 = 0; s0 < tn; s0 += K1_TVS) {
                    const int sn = min(K1_TVS, tn - s0);
                    const int nrows = tv_left < sn ? 1 + (int)((sn - 1 - tv_left) / P.tv_hop) : 0;
                    // every thread fetches the values of its own filters (7 * F per row): independent loads, all in
                    // flight together, and nobody else reads them back — no barrier needed
                    // (loads of four rows are issued before any of them is stored: an in-order warp would otherwise
                    // wait for each load at the store that follows it)
                    for (int r0 = 0; r0 < nrows; r0 += 4) {
                        T tmp[4][7 * F];
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            const T* src = P.tv_rows + (size_t)(tv_row + r0 + u) * 7 * N;
#pragma unroll
                            for (int k = 0; k < 7; k++)
#pragma unroll
                                for (int f = 0; f < F; f++) {
                                    const int n = n_base + f * NT + tid;
                                    tmp[u][k * F + f] = (r0 + u < nrows && n < N) ? src[(size_t)k * N + n] : (T)0;
                                }
                        }
#pragma unroll
                        for (int u = 0; u < 4; u++) {
                            T* dst = tvs + (size_t)(r0 + u) * 7 * NT * F;
                            if (r0 + u < nrows) {
#pragma unroll
                                for (int k = 0; k < 7; k++)
#pragma unroll
                                    for (int f = 0; f < F; f++) dst[k * NT * F + f * NT + tid] = tmp[u][k * F + f];
                            }
                        }
                    }
                    int lr = 0;
                    for (tt = s0; tt < s0 + sn; tt++) {
                        if (tv_left == 0) {
                            const T* row = tvs + (size_t)lr * 7 * NT * F;
#pragma unroll
                            for (int f = 0; f < F; f++) {
                                const int j = f * NT + tid;
                                r.b0[f] = row[0 * NT * F + j]; r.b1[f] = row[1 * NT * F + j]; r.b2[f] = row[2 * NT * F + j];
                                r.a1[f] = row[3 * NT * F + j]; r.a2[f] = row[4 * NT * F + j];
                                r.cp[f] = A::mul(P.c_p, row[5 * NT * F + j]);
                                r.cg[f] = A::mul(P.c_g, row[6 * NT * F + j]);
                            }
                            lr++;
                            tv_row++;
                            tv_left = P.tv_hop;
                        }
                        tv_left--;
                        const T x0 = xc[t0 + tt];
                        k1_sample<T, FMA, F, true, BANDS>(r, x0, x1, x2, P.s_p, P.s_g, ybs, vv, gg, dist);
                        x2 = x1; x1 = x0;
                        emit(vb + tt * RS, gb + tt * RS, tg + tt);
                    }
                }
                conv = false;
            } else {
                for (; tt < tn; tt++) {
                    if (tv) {
                        if (tv_left == 0) {
                            tv_load(tv_row);
                            tv_row++;
                            tv_left = P.tv_hop;
                        }
                        tv_left--;
                    }
                    const T x0 = xc[t0 + tt];
                    k1_sample<T, FMA, F, true, BANDS>(r, x0, x1, x2, P.s_p, P.s_g, ybs, vv, gg, dist);
                    x2 = x1; x1 = x0;
                    emit(vb + tt * RS, gb + tt * RS, tg + tt);
                }
                conv = !tv;
#pragma unroll
                for (int f = 0; f < F; f++)
                    conv = conv && (A::muladd(P.s_p, r.p[f], r.cp[f]) == r.p[f]) &&
                           (A::muladd(P.s_g, r.g[f], r.cg[f]) == r.g[f]);
            }
            // hand the tile to the sum warp (it adds tile k while tile k+1 is rendered into the other buffer)
            all_conv = __syncthreads_and(conv);
        }
    }
    // state write-back
#pragma unroll
    for (int f = 0; f < F; f++) {
        const int n = n_base + f * NT + tid;
        if (n < N) {
            const size_t si = (size_t)s * N + n;
            P.state[0 * SN + si] = r.y1[f];
            P.state[1 * SN + si] = r.y2[f];
            P.state[2 * SN + si] = r.p[f];
            P.state[3 * SN + si] = r.g[f];
        }
    }
    if (grp == 0 && tid == 0 && L > 0) { P.xh_out[2 * s] = x1; P.xh_out[2 * s + 1] = x2; }
}

// out[s][t] = sum_g partial[g][s][t], fixed order -> deterministic
template <typename T>
__global__ void __launch_bounds__(256) fb_reduce_partials(const T* __restrict__ partial, T* __restrict__ out,
                                                          int groups, int S, long L, long out_stride,
                                                          Epilogue epi, const T* __restrict__ in, long in_stride) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    const int s = blockIdx.y;
    if (i >= L) return;
    T sum = (T)0;
    for (int g = 0; g < groups; g++) sum += partial[((size_t)g * S + s) * (size_t)L + i];
    if (epi.kind) sum = epilogue_dev<T>(epi, sum, in[(size_t)s * in_stride + i]);
    out[(size_t)s * out_stride + i] = sum;
}

// tree mode with a single row (the kernels wrote out[] themselves): the epilogue as a pass of its own
template <typename T>
__global__ void __launch_bounds__(256) fb_epilogue_pass(T* __restrict__ out, long out_stride, long L, Epilogue epi, const T* __restrict__ in, long in_stride) {
    const long i = (long)blockIdx.x * 256 + threadIdx.x;
    const int s = blockIdx.y;
    if (i < L) out[(size_t)s * out_stride + i] = epilogue_dev<T>(epi, out[(size_t)s * out_stride + i], in[(size_t)s * in_stride + i]);
}

template <typename T, bool FMA, int F, int NT, bool BANDS>
static hg_status launch_k1(const FbK1Params<T>& P, int groups, cudaStream_t st) {
    size_t smem = sizeof(T) * (2 * K1_XT + (FMA ? 4 : 2) * K1_TILE * (size_t)(NT * F + 1));
    if (P.tv_rows && P.tv_hop < K1_TVS) smem += sizeof(T) * (size_t)(K1_TVS + 1) * 7 * NT * F;
    auto kern = fb_exact_kernel<T, FMA, F, NT, BANDS>;
    if (smem > 48 * 1024) HG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    if (P.chain.ordered && (P.out || P.chain.outbox)) {
        // every CTA of a stream's chain must be resident at once (see fb_launch_exact_ru)
        int per_sm = 0, dev = 0, sms = 0;
        HG_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, NT + 32, smem));
        HG_CUDA(cudaGetDevice(&dev));
        HG_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
        if ((long)groups > (long)per_sm * sms)
            return fail(HG_ERR_UNSUPPORTED, "ordered mixdown (lane kernel): %d filters of one stream need %d CTAs, this GPU holds %d at once — select HG_SUM_TREE for this bank",
                        P.N, groups, per_sm * sms);
    }
    kern<<<dim3(groups, P.S), NT + 32, smem, st>>>(P);
    HG_CUDA(cudaGetLastError());
    return HG_OK;
}

template <typename T, bool FMA, int NT, bool BANDS>
static hg_status launch_k1_F(int F, const FbK1Params<T>& P, int groups, cudaStream_t st) {
    switch (F) {
        case 1: return launch_k1<T, FMA, 1, NT, BANDS>(P, groups, st);
        case 2: return launch_k1<T, FMA, 2, NT, BANDS>(P, groups, st);
    }
    return fail(HG_ERR_INVALID, "lane kernel: %d filters per thread is not built (1, 2)", F);
}

// Kernel choice.  F_out = 0: warp-specialised kernel (latency-bound regime: every CTA of 32 filters resident at
// once, 1, 2 or 4 per SM; fp64: 1 or 2; opt-in with HG_FB_WS_NF=2, float: CTAs of 64 filters, two chains per
// recurrence lane, 1 or 2 per SM — measured no faster than two CTAs of 32); -2: RU/S kernel
// (throughput regime, fused followers); 1, 2: lane kernel with that many filters per thread (time-varying
// parameter tables, or forced).
static void k1_config(hg_fb* fb, int* F_out, int* groups_out) {
    int nf = 1;
    if (fb->dtype == HG_F32 && fb->force_ws_nf == 2) nf = 2;
    const int per_cta = 32 * nf;
    const long ws_ctas = (long)((fb->N + per_cta - 1) / per_cta) * fb->S;
    const int max_level = nf == 2 ? 1 : (fb->dtype == HG_F32 ? 2 : 1);
    int level = 0;
    while (level < max_level && ws_ctas > ((long)fb->sm_count << level)) level++;
    const bool ws_fits = ws_ctas <= ((long)fb->sm_count << level);
    *groups_out = 0;
    if (fb->cur_env) { *F_out = -2; return; }              // fused envelope followers live in the RU/S kernel
    if (!fb->cur_tv && (fb->force_F == -1 || (fb->force_F == 0 && ws_fits))) {
        fb->ws_level = level;
        fb->ws_nf = nf;
        *F_out = 0;
        *groups_out = (fb->N + per_cta - 1) / per_cta;
        return;
    }
    if (!fb->cur_tv && (fb->force_F == 0 || fb->force_F == -2)) { *F_out = -2; return; }
    int F = 1;
    if (!fb->cur_tv && fb->force_F == 2 && fb->dtype == HG_F32) F = 2;
    *F_out = F;
    *groups_out = (fb->N + K1_NT * F - 1) / (K1_NT * F);
}

}  // namespace hg

extern "C" hg_status hg_fb_debug_ws_config(int dtype, int n_filters, int n_streams, int sm_count, int filters_per_lane, long long out[8]) {
    if (!out || n_filters < 1 || n_streams < 1 || sm_count < 1 || (dtype != HG_F32 && dtype != HG_F64))
        return hg::fail(HG_ERR_INVALID, "hg_fb_debug_ws_config: bad argument");
    hg_fb fb;
    fb.dtype = dtype; fb.arith = HG_ARITH_STRICT; fb.order = 2; fb.N = n_filters; fb.S = n_streams; fb.flags = 0; fb.device = 0;
    fb.sm_count = sm_count;
    fb.force_ws_nf = filters_per_lane;
    int F = 0, groups = 0;
    hg::k1_config(&fb, &F, &groups);
    for (int i = 0; i < 8; i++) out[i] = 0;
    if (F != 0) return HG_OK;                      // the bank does not fit the warp-specialised kernel: out[0] = 0
    int tile = 0, us = 0, ys = 0;
    long long smem = 0;
    const bool built = dtype == HG_F32 ? hg::fb_ws_describe<float>(fb.ws_level, fb.ws_nf, &tile, &us, &ys, &smem)
                                       : hg::fb_ws_describe<double>(fb.ws_level, fb.ws_nf, &tile, &us, &ys, &smem);
    if (!built) return hg::fail(HG_ERR_INVALID, "hg_fb_debug_ws_config: level %d with %d filters per lane is not built", fb.ws_level, fb.ws_nf);
    out[0] = 1; out[1] = fb.ws_level; out[2] = fb.ws_nf; out[3] = groups; out[4] = (long long)groups * n_streams;
    out[5] = tile; out[6] = ys; out[7] = smem;
    return HG_OK;
}

namespace hg {

// rings, credits, ticket counter and watchdog flag of the ordered mixdown for `links` chain links
static hg_status chain_reserve(hg_fb* fb, size_t links, cudaStream_t st) {
    const size_t wps = hg::dtype_size(fb->dtype) / 4;
    int lg = 11;
    while (lg > 8 && ((links << lg) * wps * 8) > ((size_t)256 << 20)) lg--;
    if (!fb->d_ticket) {
        HG_CUDA(cudaMalloc((void**)&fb->d_ticket, sizeof(unsigned)));
        HG_CUDA(cudaMemsetAsync(fb->d_ticket, 0, sizeof(unsigned), st));
        HG_CUDA(cudaMalloc((void**)&fb->d_chain_error, sizeof(int)));
        HG_CUDA(cudaMemsetAsync(fb->d_chain_error, 0, sizeof(int), st));
    }
    if (links > fb->chain_links || lg != fb->chain_ring_log2) {
        // a new geometry: fresh, zeroed rings (tag 0 is never valid), so no stale tag of another ring size can match
        if (fb->d_chain_ring) HG_CUDA(cudaFree(fb->d_chain_ring));
        if (fb->d_chain_credit) HG_CUDA(cudaFree(fb->d_chain_credit));
        fb->d_chain_ring = nullptr; fb->d_chain_credit = nullptr; fb->chain_links = 0;
        const size_t words = (links << lg) * wps;
        cudaError_t ce;
        if ((ce = cudaMalloc(&fb->d_chain_ring, words * 8)) != cudaSuccess || (ce = cudaMalloc((void**)&fb->d_chain_credit, links * (1 + chain::CW) * 8)) != cudaSuccess)
            return fail(HG_ERR_NOMEM, "ordered mixdown: %zu links of %d samples: %s", links, 1 << lg, cudaGetErrorString(ce));
        HG_CUDA(cudaMemsetAsync(fb->d_chain_ring, 0, words * 8, st));
        HG_CUDA(cudaMemsetAsync(fb->d_chain_credit, 0, links * (1 + chain::CW) * 8, st));
        fb->chain_links = links; fb->chain_ring_log2 = lg; fb->chain_ring_words = words;
    }
    return HG_OK;
}

template <typename T>
hg_status fb_process_exact(hg_fb* fb, const T* d_in, T* d_out, T* d_bands, size_t L, cudaStream_t st) {
    int F, groups;
    k1_config(fb, &F, &groups);
    int ru_W = 0, ru_NS = 0, ru_cps = 0;
    if (F == -2) {
        fb_ru_config<T>(fb, d_bands != nullptr, &ru_W, &ru_NS, &ru_cps);
        groups = ru_cps;                     // partial-sum rows per stream: one per CTA
    }
    fb->last_F = F;
    fb->last_ctas = (F == -2 ? ru_cps : groups) * fb->S;
    // ordered mixdown (default): the running row travels through the CTAs, the chain's tail writes out[t] — one
    // row, no scratch.  Tree mode: one partial row per CTA, added up by fb_reduce_partials.
    const bool peer = fb->cur_peer.inbox || fb->cur_peer.outbox;
    const bool summing = d_out != nullptr || fb->cur_peer.outbox != nullptr;
    const bool ordered = fb->sum_mode != 0 || peer || fb->cur_carry_in;
    const int rows = ordered ? 1 : groups;   // partial-sum rows that reach HBM
    HG_TRY(chain_reserve(fb, ordered && summing ? (size_t)groups * fb->S : 1, st));
    if (peer && fb->chain_ring_log2 != 11) return fail(HG_ERR_UNSUPPORTED, "sharded rendering: the bank has too many chain links for the peer ring size");
    // segment the block so the partial-sum scratch stays bounded
    size_t seg = L;
    if (rows > 1 && d_out) {
        const size_t budget = (size_t)1024 << 20;     // partial-sum scratch: 1 GiB of the 180 GB
        size_t per_sample = (size_t)rows * fb->S * sizeof(T);
        size_t max_seg = budget / per_sample;
        max_seg = max_seg < 4096 ? 4096 : (max_seg / 1024) * 1024;
        if (seg > max_seg) seg = max_seg;
        HG_TRY(fb->partial.reserve((size_t)rows * fb->S * seg * sizeof(T)));
    }
    const bool vec_ok = d_bands && (((uintptr_t)d_bands) % 16 == 0) && (L % 4 == 0);
    // replica followers: history + block in one row per stream, so that the lagged input x[t - width] of
    // block tick t is xcat[4 + t] (four leading samples keep x[t-1], x[t-2] of the lagged stream addressable
    // and the rows 16-byte aligned)
    const bool replica = F == -2 && fb->cur_env && fb->env_mode == 2;
    const size_t H = (size_t)fb->rms_width + 4, pitch = H + L;
    if (replica) {
        HG_TRY(fb->xcat.reserve((size_t)fb->S * pitch * sizeof(T)));
        HG_CUDA(cudaMemcpy2DAsync(fb->xcat.p, pitch * sizeof(T), fb->d_xhist, H * sizeof(T), H * sizeof(T), fb->S, cudaMemcpyDeviceToDevice, st));
        HG_CUDA(cudaMemcpy2DAsync((T*)fb->xcat.p + H, pitch * sizeof(T), d_in, L * sizeof(T), L * sizeof(T), fb->S, cudaMemcpyDeviceToDevice, st));
    }
    size_t len = 0;
    for (size_t t0 = 0; t0 < L; t0 += len) {
        len = (L - t0 < seg) ? (L - t0) : seg;
        if (replica) {
            // parameter sets reach the replicas `width` ticks after the filters: hand over the ones that are due,
            // and end this segment where the next one becomes due
            const long long now = fb->rms_ticks + (long long)t0;
            while (!fb->lag_pending.empty() && fb->lag_pending.front().tick <= now) {
                const hg_fb::LagPending& pe = fb->lag_pending.front();
                // pageable source: cudaMemcpyAsync returns once the set has been copied to the driver's staging
                // buffer, so it may be dropped right away — no stream synchronisation on the render path
                HG_CUDA(cudaMemcpyAsync(fb->d_lag_par, pe.par.data(), pe.par.size(), cudaMemcpyHostToDevice, st));
                fb->lag_taps = pe.taps;
                fb->lag_pending.pop_front();
            }
            if (!fb->lag_pending.empty() && fb->lag_pending.front().tick < now + (long long)len)
                len = (size_t)(fb->lag_pending.front().tick - now);
        }
        FbK1Params<T> P;
        P.par = (const T*)fb->d_par;
        P.state = (T*)fb->d_state;
        P.xh = (const T*)fb->d_xh + (size_t)fb->xh_cur * 2 * fb->S;          // ping-pong: late CTAs of this
        P.xh_out = (T*)fb->d_xh + (size_t)(fb->xh_cur ^ 1) * 2 * fb->S;      // launch still read the old history
        fb->xh_cur ^= 1;
        P.in = d_in + t0;
        P.in_stride = (long)L;
        P.direct = (rows == 1) ? 1 : 0;
        P.out = d_out ? (P.direct ? d_out + t0 : (T*)fb->partial.p) : nullptr;
        P.out_stride = (long)L;
        P.bands = d_bands ? d_bands + t0 : nullptr;
        P.bands_stride = (long)L;
        P.bands_vec = (vec_ok && (t0 % 4 == 0)) ? 1 : 0;
        P.taps = fb->taps;
        P.dist = fb->cur_dist;
        P.epi = fb->epi;
        P.dry_in = (fb->cur_dry_in ? (const T*)fb->cur_dry_in : d_in) + t0;
        P.chain = fb->cur_peer;
        P.chain.ring = (unsigned long long*)fb->d_chain_ring;
        P.chain.credit = fb->d_chain_credit;
        P.chain.credit_w = fb->d_chain_credit + fb->chain_links;      // per-warp credit words behind the per-link ones
        P.chain.vbase = fb->chain_vbase;
        P.chain.ticket = fb->d_ticket;
        P.chain.ticket_base = fb->ticket_base;
        P.chain.error = fb->d_chain_error;
        P.chain.ring_log2 = fb->chain_ring_log2;
        P.chain.ordered = ordered ? 1 : 0;
        P.prof = nullptr;
        if (fb->d_prof && fb->last_ctas <= 4096) { HG_CUDA(cudaMemsetAsync(fb->d_prof, 0, 4096 * 8 * sizeof(long long), st)); P.prof = fb->d_prof; }
        P.carry_in = fb->cur_carry_in ? (const T*)fb->cur_carry_in + t0 : nullptr;
        P.carry_stride = (long)L;
        P.N = fb->N;
        P.S = fb->S;
        P.L = (long)len;
        P.c_p = (T)fb->c_p; P.s_p = (T)fb->s_p; P.c_g = (T)fb->c_g; P.s_g = (T)fb->s_g;
        hg_status rc;
        const bool fma = fb->arith == HG_ARITH_FMA;
        P.sm_slot = fb->d_sm_slot;
        P.tv_rows = (const T*)fb->cur_tv; P.tv_hop = fb->cur_tv_hop; P.tv_t0 = (long)t0;
        if (F == 0) {
            if (fb->ws_level > 0) HG_CUDA(cudaMemsetAsync(fb->d_sm_slot, 0, 1024 * sizeof(int), st));
            rc = fb_launch_exact_ws<T>(fma, d_bands != nullptr, fb->ws_level, fb->ws_nf, P, groups, st);
        }
        else if (F == -2) {
            FbEnvArgs<T> E;
            if (fb->cur_env) {
                E.env = (T*)fb->cur_env_out; E.rms_sum = (T*)fb->d_rms_sum; E.rms_ring = (T*)fb->d_rms_ring;
                E.width = fb->rms_width; E.hop = fb->cur_hop; E.env_rows = (long)fb->cur_env_rows;
                E.t_base = (long)t0; E.tick0 = fb->rms_ticks;
                if (replica) {
                    E.in_lag = (const T*)fb->xcat.p + 4 + t0; E.lag_stride = (long)pitch; E.lag_state = (T*)fb->d_lag_state;
                    E.lag_par = (const T*)fb->d_lag_par; E.lag_taps = fb->lag_taps;
                }
            }
            rc = fb_launch_exact_ru<T>(fma, d_bands != nullptr, P, ru_W, ru_NS, ru_cps, st, fb->cur_env ? &E : nullptr);
        }
        else if (d_bands) rc = fma ? launch_k1_F<T, true, K1_NT, true>(F, P, groups, st) : launch_k1_F<T, false, K1_NT, true>(F, P, groups, st);
        else rc = fma ? launch_k1_F<T, true, K1_NT, false>(F, P, groups, st) : launch_k1_F<T, false, K1_NT, false>(F, P, groups, st);
        HG_TRY(rc);
        fb->last_launches++;
        // the next launch's virtual positions start on a fresh lap of the largest ring; its CTAs take tickets after these
        fb->chain_vbase += (((unsigned long long)len + 2047ull) & ~2047ull) + 2048ull;
        fb->ticket_base += (unsigned)fb->last_ctas;
        if (!P.direct && d_out) {
            dim3 grid((unsigned)((len + 255) / 256), fb->S);
            fb_reduce_partials<T><<<grid, 256, 0, st>>>((const T*)fb->partial.p, d_out + t0, rows, fb->S, (long)len, (long)L, P.epi, P.dry_in, (long)L);
            HG_CUDA(cudaGetLastError());
            fb->last_launches++;
        } else if (!ordered && d_out && P.epi.kind) {
            dim3 grid((unsigned)((len + 255) / 256), fb->S);
            fb_epilogue_pass<T><<<grid, 256, 0, st>>>(d_out + t0, (long)L, (long)len, P.epi, P.dry_in, (long)L);
            HG_CUDA(cudaGetLastError());
            fb->last_launches++;
        }
    }
    if (replica)       // the last width + 4 samples become the history of the next block
        HG_CUDA(cudaMemcpy2DAsync(fb->d_xhist, H * sizeof(T), (const T*)fb->xcat.p + L, pitch * sizeof(T), H * sizeof(T), fb->S, cudaMemcpyDeviceToDevice, st));
    return HG_OK;
}

template hg_status fb_process_exact<float>(hg_fb*, const float*, float*, float*, size_t, cudaStream_t);
template hg_status fb_process_exact<double>(hg_fb*, const double*, double*, double*, size_t, cudaStream_t);

}  // namespace hg
