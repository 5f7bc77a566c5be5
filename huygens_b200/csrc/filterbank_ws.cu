// filterbank_ws.cu — K1-ws: warp-specialised exact Filterbank kernel for the latency-bound regime
// (few lanes, long blocks: BASELINE configs[1], 4096 filters x 2.88 M samples).
//
// Per filter and sample the reference (src/filterbank.h:170-187) computes
//     u = fl(fl(fl(b0 x0 + b1 x1) + b2 x2) * p)          feed-forward: independent of y
//     y = fl(u - fl(fl(a1 y1) + fl(a2 y2)))              recurrence: FMUL -> FADD -> FADD chain
//     v = fl(y * g)                                       mixdown term
// Bit-exact replay means the y chain costs 3 dependent FP32 ops = 12 cycles per sample and cannot
// be shortened; everything else can move off the critical warp.  One CTA = 32 filters of one
// stream, four warps, one per SM sub-partition:
//     L (loader)     TMA 1-D bulk copies (cp.async.bulk + mbarrier complete_tx) of the input into a
//                    3-deep ring of 1024-sample chunks; falls back to plain loads when unaligned
//     U (forward)    lane = filter: u tiles [32][TILE] into shared memory, owns the p smoother
//     R (recurrence) lane = filter: LDS.128 u -> chained samples -> STS.128 y, owns y1,y2 and g
//     S (sum)        lane = sample: out[t] = sum over the 32 lanes of y*g, fixed order; also streams
//                    per-band rows in bands mode
// Tiles are handed over with full/empty mbarriers; no __syncthreads in the steady state.
#include <type_traits>
#include "hg_common.cuh"
#include "filterbank_handle.cuh"

namespace hg {

namespace ws {

constexpr int XT = 1024;      // input chunk (samples)
constexpr int XB = 3;         // input ring depth
constexpr int NB = 3;         // u / v tile ring depth
constexpr int CS_MAX = 8;     // thread-block cluster size of the in-cluster partial-sum reduction

template <typename T> struct Cfg;
template <> struct Cfg<float> { static constexpr int TILE = 128, RS = 132; };    // RS/4 odd: conflict-free 128-bit rows
template <> struct Cfg<double> { static constexpr int TILE = 64, RS = 66; };      // RS/2 odd

__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count));
}
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_tx(unsigned long long* b, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(b)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool mbar_test(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
// non-blocking probe (no suspend): used to look ahead from inside the recurrence chain
__device__ __forceinline__ bool mbar_poll(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok) : "r"(smem_u32(b)), "r"(parity));   // no "memory" clobber: must not fence the unrolled chain
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    while (!mbar_test(b, parity)) {}
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                 ::"r"(smem_u32(dst)), "l"(src), "r"(bytes), "r"(smem_u32(bar)) : "memory");
}

// ---- thread-block cluster / distributed shared memory helpers -----------------------------------
__device__ __forceinline__ unsigned cluster_ctarank() {
    unsigned r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ unsigned map_to_cta(const void* local_smem, unsigned rank) {
    unsigned r;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(smem_u32(local_smem)), "r"(rank));
    return r;
}
__device__ __forceinline__ void remote_arrive(unsigned remote_bar) {
    asm volatile("mbarrier.arrive.release.cluster.shared::cluster.b64 _, [%0];" ::"r"(remote_bar) : "memory");
}
__device__ __forceinline__ void mbar_wait_cluster(unsigned long long* b, unsigned parity) {
    unsigned ok = 0;
    while (!ok) {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.acquire.cluster.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    }
}
__device__ __forceinline__ void remote_store_row(unsigned raddr, const float (&v)[4]) {
    asm volatile("st.shared::cluster.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(raddr), "f"(v[0]), "f"(v[1]), "f"(v[2]), "f"(v[3]) : "memory");
}
__device__ __forceinline__ void remote_store_row(unsigned raddr, const double (&v)[2]) {
    asm volatile("st.shared::cluster.v2.f64 [%0], {%1, %2};" ::"r"(raddr), "d"(v[0]), "d"(v[1]) : "memory");
}

template <typename T> __device__ __forceinline__ void ld4(const T* p, T (&x)[4]);
template <> __device__ __forceinline__ void ld4<float>(const float* p, float (&x)[4]) {
    float4 v = *reinterpret_cast<const float4*>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}
template <> __device__ __forceinline__ void ld4<double>(const double* p, double (&x)[4]) {
    double2 a = *reinterpret_cast<const double2*>(p), b = *reinterpret_cast<const double2*>(p + 2);
    x[0] = a.x; x[1] = a.y; x[2] = b.x; x[3] = b.y;
}
template <typename T> __device__ __forceinline__ void st4(T* p, const T (&x)[4]);
template <> __device__ __forceinline__ void st4<float>(float* p, const float (&x)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(x[0], x[1], x[2], x[3]);
}
template <> __device__ __forceinline__ void st4<double>(double* p, const double (&x)[4]) {
    *reinterpret_cast<double2*>(p) = make_double2(x[0], x[1]);
    *reinterpret_cast<double2*>(p + 2) = make_double2(x[2], x[3]);
}

// 16-byte row fragments: 4 floats or 2 doubles
template <typename T> __device__ __forceinline__ void ldrow(const float* p, float (&x)[4]) { ld4<float>(p, x); }
template <typename T> __device__ __forceinline__ void ldrow(const double* p, double (&x)[2]) {
    double2 a = *reinterpret_cast<const double2*>(p);
    x[0] = a.x; x[1] = a.y;
}
template <typename T> __device__ __forceinline__ void strow(float* p, const float (&x)[4]) { st4<float>(p, x); }
template <typename T> __device__ __forceinline__ void strow(double* p, const double (&x)[2]) {
    *reinterpret_cast<double2*>(p) = make_double2(x[0], x[1]);
}

template <typename T> struct Smem {
    static constexpr int TILE = Cfg<T>::TILE, RS = Cfg<T>::RS;
    alignas(16) T xs[XB][XT];
    alignas(16) T u[NB][32 * RS];      // U -> R: feed-forward terms
    alignas(16) T y[NB][32 * RS];      // R -> S: filter outputs
    alignas(16) T v[NB][32 * RS];      // R -> S: per-sample mix gains g[t], only while they still move
    T gval[NB][32];                    // converged mix gains of the tile's 32 filters
    int gtile[NB];                     // 1: gains in v (per sample);  0: gains in gval (constant)
    alignas(16) T cred[2][CS_MAX - 1][TILE];   // cluster leader: partial-sum rows of the other CTAs (via DSMEM)
    unsigned long long x_full[XB], x_empty[XB], u_full[NB], u_empty[NB], v_full[NB], v_empty[NB];
    unsigned long long c_full[2], c_empty[2];  // cluster-scope hand-over of the cred slots
};

}  // namespace ws

// feed-forward term of one sample.  TAPS: 0 = general b0,b1,b2;  1 = b1 == 0 for every filter;
// 2 = b1 == 0 and b2 == -b0 (every bank in the reference: {g,0,-g}).  In modes 1/2 the skipped
// operations are exact identities: fl(b1*x1) = +-0 and x + (+-0) = x, fl((-b0)*x) = -fl(b0*x); only the
// sign of an exact zero can differ, which no later operation can turn into a nonzero difference.
template <typename T, bool FMA, int TAPS>
__device__ __forceinline__ T ws_forward(T b0, T b1, T b2, T x0, T x1, T x2, T& m1, T& m2) {
    typedef Ar<T, FMA> A;
    if (TAPS == 2 && !FMA) {
        const T m0 = A::mul(b0, x0);
        const T fw = A::sub(m0, m2);       // m2 = fl(b0 * x[t-2])
        m2 = m1; m1 = m0;
        return fw;
    }
    T fw = A::mul(b0, x0);                                   // filterbank.h:178
    if (TAPS == 0) fw = A::muladd(b1, x1, fw);
    fw = A::muladd(b2, x2, fw);
    return fw;
}

template <typename T, bool FMA, int TAPS>
__device__ __forceinline__ void ws_u_tile(const T* __restrict__ xc, T* __restrict__ urow, int tn, T b0, T b1, T b2, T p, T& x1, T& x2, T& m1, T& m2) {
    typedef Ar<T, FMA> A;
    constexpr int TILE = ws::Cfg<T>::TILE;
    if (tn == TILE) {
        // straight-line tile, 16 samples per block; the next block's inputs are fetched before this
        // block is computed so the LDS latency hides behind 16 samples of arithmetic
        T xa[16], xb[16];
#pragma unroll
        for (int q = 0; q < 4; q++) ws::ld4<T>(xc + 4 * q, *reinterpret_cast<T(*)[4]>(&xa[4 * q]));
#pragma unroll
        for (int blk = 0; blk < TILE / 16; blk++) {
            T(&cur)[16] = (blk & 1) ? xb : xa;
            T(&nxt)[16] = (blk & 1) ? xa : xb;
            if (blk + 1 < TILE / 16) {
#pragma unroll
                for (int q = 0; q < 4; q++) ws::ld4<T>(xc + 16 * (blk + 1) + 4 * q, *reinterpret_cast<T(*)[4]>(&nxt[4 * q]));
            }
#pragma unroll
            for (int q = 0; q < 4; q++) {
                T uv[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    const T x0 = cur[4 * q + j];
                    const T fw = ws_forward<T, FMA, TAPS>(b0, b1, b2, x0, x1, x2, m1, m2);
                    uv[j] = FMA ? fw : A::mul(fw, p);        // strict: u = fl(fw*p); fma: fused in R
                    x2 = x1; x1 = x0;
                }
                ws::st4<T>(urow + 16 * blk + 4 * q, uv);
            }
        }
    } else {
        for (int tt = 0; tt < tn; tt++) {
            const T x0 = xc[tt];
            const T fw = ws_forward<T, FMA, TAPS>(b0, b1, b2, x0, x1, x2, m1, m2);
            urow[tt] = FMA ? fw : A::mul(fw, p);
            x2 = x1; x1 = x0;
        }
    }
}

// CS > 1: the kernel is launched in thread-block clusters of CS CTAs; their S warps push the partial
// sums of each tile into the leader CTA's shared memory (DSMEM) and the leader writes ONE row per
// cluster, in fixed CTA order (deterministic).  HBM then sees groups/CS partial rows instead of groups.
template <typename T, bool FMA, bool BANDS, int CS>
__global__ void __launch_bounds__(128, 1) fb_exact_ws_kernel(FbK1Params<T> P) {
    using namespace ws;
    typedef Ar<T, FMA> A;
    constexpr int TILE = Cfg<T>::TILE, RS = Cfg<T>::RS;
    constexpr int SPL = TILE / 32;       // samples per S lane: 16 bytes for both dtypes
    typedef Smem<T> SM;
    extern __shared__ __align__(16) unsigned char smem_raw[];
    SM& sm = *reinterpret_cast<SM*>(smem_raw);

    const int lane = threadIdx.x & 31;
    const int grp = blockIdx.x, s = blockIdx.y;
    // rotate roles with the CTA index so that co-resident CTAs put their R warps on different schedulers
    const int role = ((threadIdx.x >> 5) + grp) & 3;     // 0 = R, 1 = U, 2 = S, 3 = L
    const int N = P.N;
    const long L = P.L;
    const int n = grp * 32 + lane;
    const bool ok = n < N;
    const size_t SN = (size_t)P.S * N, si = (size_t)s * N + n;
    // per-tile bookkeeping stays in 32-bit registers: ring slots advance with wrap, no div/mod
    const int ntiles = (int)((L + TILE - 1) / TILE);
    const int nchunks = (int)((L + XT - 1) / XT);
    const int last_tn = (int)(L - (long)(ntiles - 1) * TILE);      // samples in the final tile
    constexpr int TPC = XT / TILE;       // tiles per input chunk

    if (threadIdx.x == 0) {
        for (int i = 0; i < XB; i++) { mbar_init(&sm.x_full[i], 1); mbar_init(&sm.x_empty[i], 1); }
        for (int i = 0; i < NB; i++) {
            mbar_init(&sm.u_full[i], 1); mbar_init(&sm.u_empty[i], 1);
            mbar_init(&sm.v_full[i], 1); mbar_init(&sm.v_empty[i], 1);
        }
        if (CS > 1) {
            mbar_init(&sm.c_full[0], 32 * (CS - 1)); mbar_init(&sm.c_full[1], 32 * (CS - 1));
            mbar_init(&sm.c_empty[0], 1); mbar_init(&sm.c_empty[1], 1);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();
    if (CS > 1) cluster_sync_all();      // every CTA's barriers exist before anyone arrives remotely

    if (role == 3) {
        // ---------------- L: input loader ----------------
        const T* inp = P.in + (size_t)s * P.in_stride;
        int b = 0;
        unsigned xpar = 0;
        for (int c = 0; c < nchunks; c++) {
            mbar_wait(&sm.x_empty[b], xpar ^ 1u);
            const long base = (long)c * XT;
            const int cnt = (int)min((long)XT, L - base);
            const T* src = inp + base;
            const unsigned bytes = (unsigned)(cnt * sizeof(T));
            if ((((uintptr_t)src) & 15) == 0 && (bytes & 15) == 0) {
                if (lane == 0) {
                    mbar_arrive_tx(&sm.x_full[b], bytes);
                    bulk_g2s(sm.xs[b], src, bytes, &sm.x_full[b]);
                }
            } else {
                for (int i = lane; i < cnt; i += 32) sm.xs[b][i] = src[i];
                __syncwarp();
                if (lane == 0) mbar_arrive(&sm.x_full[b]);
            }
            if (++b == XB) { b = 0; xpar ^= 1u; }
        }
    } else if (role == 1) {
        // ---------------- U: feed-forward taps and pre-gain ----------------
        const T b0 = ok ? P.par[0 * (size_t)N + n] : (T)0, b1 = ok ? P.par[1 * (size_t)N + n] : (T)0;
        const T b2 = ok ? P.par[2 * (size_t)N + n] : (T)0;
        const T pin = ok ? P.par[5 * (size_t)N + n] : (T)0;
        T p = ok ? P.state[2 * SN + si] : (T)0;
        const T cp = A::mul(P.c_p, pin), s_p = P.s_p;
        T x1 = P.xh[2 * s], x2 = P.xh[2 * s + 1];
        T m1 = A::mul(b0, x1), m2 = A::mul(b0, x2);          // TAPS == 2: fl(b0*x[t-1]), fl(b0*x[t-2])
        const int taps = P.taps;
        // strict: U owns the p smoother (u = fl(fw*p)); fma: y = fma(fw, p, -fb) needs p in R instead
        bool conv = FMA ? true : __all_sync(0xffffffffu, A::muladd(s_p, p, cp) == p);
        int xb = 0, ub = 0, kc = 0;
        unsigned xpar = 0, upar = 0;
        for (int k = 0; k < ntiles; k++) {
            if (kc == 0) mbar_wait(&sm.x_full[xb], xpar);
            mbar_wait(&sm.u_empty[ub], upar ^ 1u);
            const T* xc = sm.xs[xb] + kc * TILE;
            T* urow = sm.u[ub] + lane * RS;
            const int tn = (k == ntiles - 1) ? last_tn : TILE;
            if (conv) {
                if (taps == 2) ws_u_tile<T, FMA, 2>(xc, urow, tn, b0, b1, b2, p, x1, x2, m1, m2);
                else if (taps == 1) ws_u_tile<T, FMA, 1>(xc, urow, tn, b0, b1, b2, p, x1, x2, m1, m2);
                else ws_u_tile<T, FMA, 0>(xc, urow, tn, b0, b1, b2, p, x1, x2, m1, m2);
            } else {
                for (int tt = 0; tt < tn; tt++) {
                    p = A::muladd(s_p, p, cp);                          // filterbank.h:172
                    const T x0 = xc[tt];
                    const T fw = ws_forward<T, FMA, 0>(b0, b1, b2, x0, x1, x2, m1, m2);
                    urow[tt] = FMA ? fw : A::mul(fw, p);
                    x2 = x1; x1 = x0;
                }
                m1 = A::mul(b0, x1); m2 = A::mul(b0, x2);
                conv = __all_sync(0xffffffffu, A::muladd(s_p, p, cp) == p);
            }
            __syncwarp();
            const bool chunk_done = (kc == TPC - 1) || (k == ntiles - 1);
            if (lane == 0) {
                mbar_arrive(&sm.u_full[ub]);
                if (chunk_done) mbar_arrive(&sm.x_empty[xb]);
            }
            if (++ub == NB) { ub = 0; upar ^= 1u; }
            if (++kc == TPC) { kc = 0; if (++xb == XB) { xb = 0; xpar ^= 1u; } }
        }
        if (ok && !FMA) P.state[2 * SN + si] = p;
        if (grp == 0 && lane == 0 && L > 0) { P.xh_out[2 * s] = x1; P.xh_out[2 * s + 1] = x2; }
    } else if (role == 0) {
        // ---------------- R: the recurrence ----------------
        // Steady state (gains converged): only the 3-op chain, a2*y2 and the tile traffic stay here;
        // the y tile goes to S, which applies g.  While g (or p, fma mode) still moves, R also writes
        // the per-sample gains g[t] and flags the tile.  A full tile is straight-line code; the
        // barriers of the next tile are polled and its first group is fetched inside the chain's
        // stall slots, so the chain restarts without waiting.
        const T a1 = ok ? P.par[3 * (size_t)N + n] : (T)0, a2 = ok ? P.par[4 * (size_t)N + n] : (T)0;
        const T gin = ok ? P.par[6 * (size_t)N + n] : (T)0;
        T y1 = ok ? P.state[0 * SN + si] : (T)0, y2 = ok ? P.state[1 * SN + si] : (T)0;
        T g = ok ? P.state[3 * SN + si] : (T)0;
        T p = ok ? P.state[2 * SN + si] : (T)0;           // fma mode only
        const T pin = ok ? P.par[5 * (size_t)N + n] : (T)0;
        const T cg = A::mul(P.c_g, gin), s_g = P.s_g, cp = A::mul(P.c_p, pin), s_p = P.s_p;
        bool gconv = __all_sync(0xffffffffu, A::muladd(s_g, g, cg) == g);
        bool pconv = FMA ? __all_sync(0xffffffffu, A::muladd(s_p, p, cp) == p) : true;
        bool ready = false, have_u0 = false;     // next tile's barriers already seen complete / first group loaded
        T u0[4] = {(T)0, (T)0, (T)0, (T)0};
        int ub = 0;
        unsigned upar = 0;
        for (int k = 0; k < ntiles; k++) {
            if (!ready) {
                mbar_wait(&sm.u_full[ub], upar);
                mbar_wait(&sm.v_empty[ub], upar ^ 1u);
            }
            const T* urow = sm.u[ub] + lane * RS;
            T* yrow = sm.y[ub] + lane * RS;
            T* vrow = sm.v[ub] + lane * RS;
            const int tn = (k == ntiles - 1) ? last_tn : TILE;
            const bool lean = gconv && pconv;
            const bool more = k + 1 < ntiles;
            const int nb = (ub + 1 == NB) ? 0 : ub + 1;
            const unsigned npar = (ub + 1 == NB) ? (upar ^ 1u) : upar;
            bool next_ready = false, next_u0 = false;
            bool poll_u = false, poll_v = false;
            if (lean && tn == TILE) {
                if (lane == 0) sm.gtile[ub] = 0;
                sm.gval[ub][lane] = g;
                T t2 = A::mul(a2, y2);
                // rotating register buffers, three groups deep: group q+2 is requested while group q runs,
                // so a u load has ~96 cycles to land even when the SM's shared-memory pipe is busy
                // with the U and S warps (measured: one group of look-ahead left ~1 cycle/sample of
                // short-scoreboard stalls on the chain)
                T ub3[3][4];
                if (have_u0) { ub3[0][0] = u0[0]; ub3[0][1] = u0[1]; ub3[0][2] = u0[2]; ub3[0][3] = u0[3]; }
                else ld4<T>(urow, ub3[0]);
                ld4<T>(urow + 4, ub3[1]);
#pragma unroll
                for (int q = 0; q < TILE / 4; q++) {
                    T yy[4];
                    if (q + 2 < TILE / 4) ld4<T>(urow + 4 * (q + 2), ub3[(q + 2) % 3]);
                    const T(&uc)[4] = ub3[q % 3];
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        // filterbank.h:179 fb = fl(fl(a1 y1) + fl(a2 y2)); :178-179 y = fl(u - fb)
                        const T fb = FMA ? A::muladd(a2, y2, A::mul(a1, y1)) : A::add(A::mul(a1, y1), t2);
                        const T y = FMA ? A::mulsub(uc[j], p, fb) : A::sub(uc[j], fb);
                        t2 = A::mul(a2, y1);
                        y2 = y1; y1 = y;
                        yy[j] = y;
                    }
                    st4<T>(yrow + 4 * q, yy);
                    // look ahead: probe the next tile's barriers early, consume the answer late, so the
                    // probe latency and the first LDS of the next tile sit in the chain's stall slots
                    if (q == TILE / 4 - 16 && more) {
                        poll_u = mbar_poll(&sm.u_full[nb], npar);
                        poll_v = mbar_poll(&sm.v_empty[nb], npar ^ 1u);
                    }
                    if (q == TILE / 4 - 5 && more) next_ready = __all_sync(0xffffffffu, poll_u && poll_v);
                    if (q == TILE / 4 - 4 && next_ready) {
                        ld4<T>(sm.u[nb] + lane * RS, u0);
                        next_u0 = true;
                    }
                }
            } else {
                if (lane == 0) sm.gtile[ub] = lean ? 0 : 1;
                sm.gval[ub][lane] = g;
                for (int tt = 0; tt < tn; tt++) {
                    if (!gconv) g = A::muladd(s_g, g, cg);              // filterbank.h:173
                    if (FMA && !pconv) p = A::muladd(s_p, p, cp);       // filterbank.h:172 (fma mode: p lives here)
                    T fb = A::mul(a1, y1);
                    fb = A::muladd(a2, y2, fb);
                    const T y = FMA ? A::mulsub(urow[tt], p, fb) : A::sub(urow[tt], fb);
                    y2 = y1; y1 = y;
                    yrow[tt] = y;
                    if (!lean) vrow[tt] = g;                            // S applies the moving gain g[t]
                }
                if (!gconv) gconv = __all_sync(0xffffffffu, A::muladd(s_g, g, cg) == g);
                if (FMA && !pconv) pconv = __all_sync(0xffffffffu, A::muladd(s_p, p, cp) == p);
            }
            ready = next_ready;
            have_u0 = next_u0;
            __syncwarp();
            if (lane == 0) { mbar_arrive(&sm.u_empty[ub]); mbar_arrive(&sm.v_full[ub]); }
            ub = nb; upar = npar;
        }
        if (ok) {
            P.state[0 * SN + si] = y1; P.state[1 * SN + si] = y2; P.state[3 * SN + si] = g;
            if (FMA) P.state[2 * SN + si] = p;
        }
    } else {
        // ---------------- S: mixdown over the 32 lanes, per-band rows ----------------
        // lane owns SPL consecutive samples (16 bytes) of every row: LDS.128 rows, STG.128 out.
        // out[t] = sum_l fma(y_l[t], g_l[t], .) in lane order l = 0..31 (filterbank.h:130): the same
        // function of (y, g) whether the gains still move (g tile) or not (gval), so the sum does
        // not depend on how the block is split into calls or tiles.
        T* out_row = nullptr;
        if (P.out) out_row = P.direct ? (P.out + (size_t)s * P.out_stride) : (P.out + ((size_t)(grp / CS) * P.S + s) * (size_t)L);
        const bool out_vec = out_row && ((((uintptr_t)out_row) & 15) == 0);
        const unsigned crank = CS > 1 ? cluster_ctarank() : 0u;
        const int rows = min(32, N - grp * 32);
        int ub = 0;
        unsigned upar = 0;
        for (int k = 0; k < ntiles; k++) {
            mbar_wait(&sm.v_full[ub], upar);
            const int tn = (k == ntiles - 1) ? last_tn : TILE;
            const long tg = (long)k * TILE;
            const T* yt = sm.y[ub] + SPL * lane;
            const T* gt = sm.v[ub] + SPL * lane;
            const bool moving = sm.gtile[ub] != 0;
            T acc[SPL];
#pragma unroll
            for (int j = 0; j < SPL; j++) acc[j] = (T)0;
#pragma unroll 8
            for (int l = 0; l < 32; l++) {
                T yv[SPL], gl[SPL];
                ldrow<T>(yt + l * RS, yv);
                if (moving) ldrow<T>(gt + l * RS, gl);
                else {
                    const T gc = sm.gval[ub][l];
#pragma unroll
                    for (int j = 0; j < SPL; j++) gl[j] = gc;
                }
#pragma unroll
                for (int j = 0; j < SPL; j++) acc[j] = A::acc(yv[j], gl[j], acc[j]);
                if (BANDS && l < rows) {
                    T* dst = P.bands + ((size_t)s * N + grp * 32 + l) * P.bands_stride + tg + SPL * lane;
                    if (P.bands_vec && tn == TILE) strow<T>(dst, yv);
                    else {
#pragma unroll
                        for (int j = 0; j < SPL; j++) if (SPL * lane + j < tn) dst[j] = yv[j];
                    }
                }
            }
            if (CS > 1 && out_row) {
                const int slot = k & 1;
                const unsigned use_par = (unsigned)((k >> 1) & 1);
                if (crank != 0) {
                    // hand this CTA's row to the leader: wait until it has consumed the slot's last use
                    mbar_wait_cluster(&sm.c_empty[slot], use_par ^ 1u);
                    remote_store_row(map_to_cta(&sm.cred[slot][crank - 1][SPL * lane], 0), acc);
                    remote_arrive(map_to_cta(&sm.c_full[slot], 0));
                } else {
                    mbar_wait_cluster(&sm.c_full[slot], use_par);
#pragma unroll
                    for (int r = 0; r < CS - 1; r++) {
                        T rv[SPL];
                        ldrow<T>(&sm.cred[slot][r][SPL * lane], rv);
#pragma unroll
                        for (int j = 0; j < SPL; j++) acc[j] += rv[j];
                    }
                    __syncwarp();
                    if (lane < CS - 1) remote_arrive(map_to_cta(&sm.c_empty[slot], (unsigned)lane + 1u));
                }
            }
            if (out_row && crank == 0) {
                T* dst = out_row + tg + SPL * lane;
                if (out_vec && tn == TILE && ((tg * sizeof(T)) & 15) == 0) strow<T>(dst, acc);
                else {
#pragma unroll
                    for (int j = 0; j < SPL; j++) if (SPL * lane + j < tn) dst[j] = acc[j];
                }
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.v_empty[ub]);
            if (++ub == NB) { ub = 0; upar ^= 1u; }
        }
    }
    // no CTA may exit while a peer can still touch its shared memory or barriers
    if (CS > 1) cluster_sync_all();
}

template <typename T, bool FMA, bool BANDS, int CS>
static hg_status launch_ws(const FbK1Params<T>& P, int groups, cudaStream_t st) {
    typedef ws::Smem<T> SM;
    const size_t smem = sizeof(SM);
    auto kern = fb_exact_ws_kernel<T, FMA, BANDS, CS>;
    HG_CUDA(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(groups, P.S);
    cfg.blockDim = dim3(128);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = CS; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = CS > 1 ? 1 : 0;
    HG_CUDA(cudaLaunchKernelEx(&cfg, kern, P));
    return HG_OK;
}

// largest cluster size (8, 4, 2 or 1) that divides the group count and whose clusters are all
// co-resident in one wave (each CTA runs the whole block, so a second wave would double the time)
template <typename T>
int fb_ws_cluster_size(int groups, int streams) {
    static int cached_groups = -1, cached_streams = -1, cached_cs = 1;
    if (groups == cached_groups && streams == cached_streams) return cached_cs;
    int best = 1;
    for (int cs = ws::CS_MAX; cs > 1; cs >>= 1) {
        if (groups % cs) continue;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(groups, streams);
        cfg.blockDim = dim3(128);
        cfg.dynamicSmemBytes = sizeof(ws::Smem<T>);
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = cs; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        int nclusters = 0;
        const void* fn = cs == 8 ? (const void*)fb_exact_ws_kernel<T, false, false, 8> : cs == 4 ? (const void*)fb_exact_ws_kernel<T, false, false, 4>
                                                                                        : (const void*)fb_exact_ws_kernel<T, false, false, 2>;
        cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(ws::Smem<T>));
        if (cudaOccupancyMaxActiveClusters(&nclusters, fn, &cfg) != cudaSuccess) { cudaGetLastError(); continue; }
        if ((long)nclusters * cs >= (long)groups * streams) { best = cs; break; }
    }
    cached_groups = groups; cached_streams = streams; cached_cs = best;
    return best;
}
template int fb_ws_cluster_size<float>(int, int);
template int fb_ws_cluster_size<double>(int, int);

template <typename T, bool FMA, bool BANDS>
static hg_status launch_ws_cs(int cs, const FbK1Params<T>& P, int groups, cudaStream_t st) {
    switch (cs) {
        case 8: return launch_ws<T, FMA, BANDS, 8>(P, groups, st);
        case 4: return launch_ws<T, FMA, BANDS, 4>(P, groups, st);
        case 2: return launch_ws<T, FMA, BANDS, 2>(P, groups, st);
    }
    return launch_ws<T, FMA, BANDS, 1>(P, groups, st);
}

template <typename T>
hg_status fb_launch_exact_ws(bool fma, bool bands, int cs, const FbK1Params<T>& P, int groups, cudaStream_t st) {
    if (bands) return fma ? launch_ws_cs<T, true, true>(cs, P, groups, st) : launch_ws_cs<T, false, true>(cs, P, groups, st);
    return fma ? launch_ws_cs<T, true, false>(cs, P, groups, st) : launch_ws_cs<T, false, false>(cs, P, groups, st);
}
template hg_status fb_launch_exact_ws<float>(bool, bool, int, const FbK1Params<float>&, int, cudaStream_t);
template hg_status fb_launch_exact_ws<double>(bool, bool, int, const FbK1Params<double>&, int, cudaStream_t);

}  // namespace hg
