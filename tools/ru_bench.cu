// ru_bench.cu — how fast does ONE RU warp of K1-ru run its tile loop when it has a scheduler to itself?
// Variants of the source order of the same bit-exact arithmetic (filterbank_ru.cu: ru_sample / ru_tile_v), no barriers,
// shared-memory input and term tile as in the kernel.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o ru_bench ru_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

constexpr int TILE = 32, VS = 36;

__device__ __forceinline__ void ld4(const float* p, float (&x)[4]) {
    const float4 v = *reinterpret_cast<const float4*>(p);
    x[0] = v.x; x[1] = v.y; x[2] = v.z; x[3] = v.w;
}

struct St { float y1, y2, x1, x2, m1, m2, t2; };

// V0: the kernel's order (sample by sample)
__device__ __forceinline__ void tile_v0(const float* xc, float* vcol, float b0, float a1, float a2, float p, float g, St& s) {
    float xq[2][4];
    ld4(xc, xq[0]);
#pragma unroll
    for (int q = 0; q < TILE / 4; q++) {
        if (q + 1 < TILE / 4) ld4(xc + 4 * (q + 1), xq[(q + 1) & 1]);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float x0 = xq[q & 1][j];
            const float m0 = __fmul_rn(b0, x0);
            const float fw = __fsub_rn(m0, s.m2);
            s.m2 = s.m1; s.m1 = m0;
            const float u = __fmul_rn(fw, p);
            const float fb = __fadd_rn(__fmul_rn(a1, s.y1), s.t2);
            const float y = __fsub_rn(u, fb);
            s.t2 = __fmul_rn(a2, s.y1);
            s.y2 = s.y1; s.y1 = y;
            vcol[(4 * q + j) * VS] = __fmul_rn(y, g);
        }
    }
}

// V1: the forward part u[] of the WHOLE tile first (independent of the recurrence), then the recurrence alone
__device__ __forceinline__ void tile_v1(const float* xc, float* vcol, float b0, float a1, float a2, float p, float g, St& s) {
    float u[TILE];
#pragma unroll
    for (int q = 0; q < TILE / 4; q++) {
        float xv[4];
        ld4(xc + 4 * q, xv);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float m0 = __fmul_rn(b0, xv[j]);
            u[4 * q + j] = __fmul_rn(__fsub_rn(m0, s.m2), p);
            s.m2 = s.m1; s.m1 = m0;
        }
    }
#pragma unroll
    for (int t = 0; t < TILE; t++) {
        const float fb = __fadd_rn(__fmul_rn(a1, s.y1), s.t2);
        const float y = __fsub_rn(u[t], fb);
        s.t2 = __fmul_rn(a2, s.y1);
        s.y1 = y;
        vcol[t * VS] = __fmul_rn(y, g);
    }
}

// V2: groups of 4: forward part of group q+1 computed between the recurrence steps of group q (explicit interleave)
__device__ __forceinline__ void tile_v2(const float* xc, float* vcol, float b0, float a1, float a2, float p, float g, St& s) {
    float u[2][4];
    {
        float xv[4];
        ld4(xc, xv);
#pragma unroll
        for (int j = 0; j < 4; j++) { const float m0 = __fmul_rn(b0, xv[j]); u[0][j] = __fmul_rn(__fsub_rn(m0, s.m2), p); s.m2 = s.m1; s.m1 = m0; }
    }
#pragma unroll
    for (int q = 0; q < TILE / 4; q++) {
        float xv[4];
        if (q + 1 < TILE / 4) ld4(xc + 4 * (q + 1), xv);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float fb = __fadd_rn(__fmul_rn(a1, s.y1), s.t2);
            if (q + 1 < TILE / 4) { const float m0 = __fmul_rn(b0, xv[j]); u[(q + 1) & 1][j] = __fmul_rn(__fsub_rn(m0, s.m2), p); s.m2 = s.m1; s.m1 = m0; }
            const float y = __fsub_rn(u[q & 1][j], fb);
            s.t2 = __fmul_rn(a2, s.y1);
            s.y1 = y;
            vcol[(4 * q + j) * VS] = __fmul_rn(y, g);
        }
    }
}

// V3: recurrence only (no forward part, u from shared memory): the floor of the chain
__device__ __forceinline__ void tile_v3(const float* xc, float* vcol, float b0, float a1, float a2, float p, float g, St& s) {
#pragma unroll
    for (int q = 0; q < TILE / 4; q++) {
        float uv[4];
        ld4(xc + 4 * q, uv);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float fb = __fadd_rn(__fmul_rn(a1, s.y1), s.t2);
            const float y = __fsub_rn(uv[j], fb);
            s.t2 = __fmul_rn(a2, s.y1);
            s.y1 = y;
            vcol[(4 * q + j) * VS] = __fmul_rn(y, g);
        }
    }
}

template <int VAR>
__global__ void __launch_bounds__(512) k(float* sink, int ntiles, long long* clk, int active_stride) {
    __shared__ __align__(16) float xs[1024];
    __shared__ __align__(16) float tiles[16][2][TILE * VS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = sinf(0.37f * i) * 0.5f;
    __syncthreads();
    if (warp % active_stride) return;
    const float w = 0.01f + 0.002f * lane;
    const float r = 0.999f;
    const float a1 = -2.f * r * cosf(w), a2 = r * r, b0 = 0.01f, p = 1.0f, g = 0.5f;
    St s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    float* vbase = &tiles[warp][0][0] + lane;
    const long long c0 = clock64();
#pragma unroll 1
    for (int kt = 0; kt < ntiles; kt++) {
        const float* xc = xs + (kt & 31) * TILE;
        float* vcol = vbase + (kt & 1) * TILE * VS;
        if (VAR == 0) tile_v0(xc, vcol, b0, a1, a2, p, g, s);
        else if (VAR == 1) tile_v1(xc, vcol, b0, a1, a2, p, g, s);
        else if (VAR == 2) tile_v2(xc, vcol, b0, a1, a2, p, g, s);
        else tile_v3(xc, vcol, b0, a1, a2, p, g, s);
    }
    const long long c1 = clock64();
    if (s.y1 == 123456.f) sink[0] = s.y1 + s.t2 + s.m1 + s.m2;
    if (lane == 0) clk[warp] = c1 - c0;
}


// ---- the same loop with the kernel's per-tile hand-over: wait for the slot (released by a consumer warp), render, signal ----
__device__ __forceinline__ unsigned smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void mbar_init(unsigned long long* b, int count) { asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(b)), "r"(count)); }
__device__ __forceinline__ void mbar_arrive(unsigned long long* b) { asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(b)) : "memory"); }
__device__ __forceinline__ bool mbar_poll(unsigned long long* b, unsigned parity) {
    unsigned ok;
    asm volatile("{\n\t.reg .pred p;\n\tmbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
    return ok != 0;
}
__device__ __forceinline__ void mbar_wait(unsigned long long* b, unsigned parity) {
    unsigned ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity) : "memory");
}

__device__ __forceinline__ void mbar_wait_hint(unsigned long long* b, unsigned parity, unsigned ns) {
    unsigned ok = 0;
    while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(smem_u32(b)), "r"(parity), "r"(ns) : "memory");
}
// MODE 0: wait / render / arrive as in the kernel.  MODE 1: the next slot's barrier is probed before this tile is rendered.
// The consumer warp (warp 4 + producer index, on the producer's scheduler when SAME, else the next one) frees the slot at once.
template <int MODE>
__global__ void __launch_bounds__(512) kb(float* sink, int ntiles, long long* clk, int producers, int consumer_shift) {
    __shared__ __align__(16) float xs[1024];
    __shared__ __align__(16) float tiles[4][2][TILE * VS];
    __shared__ unsigned long long full[4][2], empty[4][2];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = sinf(0.37f * i) * 0.5f;
    if (threadIdx.x == 0) for (int w = 0; w < 4; w++) for (int b = 0; b < 2; b++) { mbar_init(&full[w][b], 1); mbar_init(&empty[w][b], 1); }
    __syncthreads();
    if (warp < producers) {
        const float w = 0.01f + 0.002f * lane, r = 0.999f;
        const float a1 = -2.f * r * cosf(w), a2 = r * r, b0 = 0.01f, p = 1.0f, g = 0.5f;
        St s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        float* vbase = &tiles[warp][0][0] + lane;
        int yb = 0; unsigned par = 0;
        bool ready = mbar_poll(&empty[warp][0], 1u);
        const long long c0 = clock64();
#pragma unroll 1
        for (int kt = 0; kt < ntiles; kt++) {
            // MODE 2: no wait at all (consumer idle); 3: wait only, never signal; 4: signal without __syncwarp; 5: hand-over every
            // second tile (two tiles per slot pair)
            if (MODE == 0 || MODE == 3 || MODE == 4) mbar_wait(&empty[warp][yb], par ^ 1u);
            else if (MODE == 1 && !ready) mbar_wait(&empty[warp][yb], par ^ 1u);
            else if (MODE == 5 && !(kt & 1)) mbar_wait(&empty[warp][yb], par ^ 1u);
            const int nyb = yb ^ 1; const unsigned npar = par ^ (unsigned)yb;
            if (MODE == 1) ready = mbar_poll(&empty[warp][nyb], npar ^ 1u);
            tile_v0(xs + (kt & 31) * TILE, vbase + (MODE == 5 ? (kt & 1) : yb) * TILE * VS, b0, a1, a2, p, g, s);
            if (MODE != 4 && MODE != 3) __syncwarp();
            if (MODE == 5) { if (kt & 1) { if (lane == 0) mbar_arrive(&full[warp][yb]); yb = nyb; par = npar; } }
            else {
                if (MODE != 3 && lane == 0) mbar_arrive(&full[warp][yb]);
                if (MODE == 3 && lane == 0) mbar_arrive(&empty[warp][yb]);       // keep our own wait satisfied
                yb = nyb; par = npar;
            }
        }
        const long long c1 = clock64();
        if (s.y1 == 123456.f) sink[0] = s.y1 + s.t2;
        if (lane == 0) clk[warp] = c1 - c0;
    } else if (warp >= 4 + consumer_shift && warp < 4 + consumer_shift + producers) {
        const int w = warp - 4 - consumer_shift;
        int yb = 0; unsigned par = 0;
        float acc = 0.f;
        const int nt = (MODE == 5) ? ntiles / 2 : ((MODE == 2 || MODE == 3) ? 0 : ntiles);
        for (int kt = 0; kt < nt; kt++) {
            mbar_wait(&full[w][yb], par);
            acc += tiles[w][yb][lane * VS];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[w][yb]);
            if (++yb == 2) { yb = 0; par ^= 1u; }
        }
        if (acc == 123456.f) sink[0] = acc;
    }
}
template <int MODE>
void runb(const char* name, int producers, int shift) {
    float* sink; long long* clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, 16 * 8);
    const int ntiles = 20000;
    for (int rep = 0; rep < 2; rep++) { kb<MODE><<<1, 12 * 32>>>(sink, ntiles, clk, producers, shift); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    long long h[16];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-60s producers %d: %6.2f cycles per sample  (%s)\n", name, producers, (double)h[0] / (ntiles * 32.0), cudaGetErrorString(e));
    cudaFree(sink); cudaFree(clk);
}

// ring of NBS slots between the producer and its consumer (MODE 0 protocol): how deep must it be to hide the hand-over's round trip?
template <int NBS, int SPIN>
__global__ void __launch_bounds__(512) kd(float* sink, int ntiles, long long* clk) {
    __shared__ __align__(16) float xs[1024];
    __shared__ __align__(16) float tiles[NBS][TILE * VS];
    __shared__ unsigned long long full[NBS], empty[NBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = sinf(0.37f * i) * 0.5f;
    if (threadIdx.x == 0) for (int b = 0; b < NBS; b++) { mbar_init(&full[b], 1); mbar_init(&empty[b], 1); }
    __syncthreads();
    if (warp == 0) {
        const float w = 0.01f + 0.002f * lane, r = 0.999f;
        const float a1 = -2.f * r * cosf(w), a2 = r * r, b0 = 0.01f, p = 1.0f, g = 0.5f;
        St s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int yb = 0; unsigned par = 0;
        const long long c0 = clock64();
#pragma unroll 1
        for (int kt = 0; kt < ntiles; kt++) {
            if (SPIN == 1) { while (!mbar_poll(&empty[yb], par ^ 1u)) {} } else if (SPIN == 3) mbar_wait_hint(&empty[yb], par ^ 1u, 10000000u); else mbar_wait(&empty[yb], par ^ 1u);
            tile_v0(xs + (kt & 31) * TILE, &tiles[yb][0] + lane, b0, a1, a2, p, g, s);
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[yb]);
            if (++yb == NBS) { yb = 0; par ^= 1u; }
        }
        const long long c1 = clock64();
        if (s.y1 == 123456.f) sink[0] = s.y1 + s.t2;
        if (lane == 0) clk[0] = c1 - c0;
    } else if (warp == 5) {
        int yb = 0; unsigned par = 0;
        float acc = 0.f;
        long long tw = 0, tt0 = clock64();
        for (int kt = 0; kt < ntiles; kt++) {
            const long long q0 = clock64();
            if (SPIN == 1) { while (!mbar_poll(&full[yb], par)) {} } else if (SPIN >= 2 && SPIN <= 3) mbar_wait_hint(&full[yb], par, 10000000u); else if (SPIN == 4) { while (!mbar_poll(&full[yb], par)) __nanosleep(200); } else mbar_wait(&full[yb], par);
            tw += clock64() - q0;
            acc += tiles[yb][lane * VS];
            __syncwarp();
            if (lane == 0) mbar_arrive(&empty[yb]);
            if (++yb == NBS) { yb = 0; par ^= 1u; }
        }
        if (acc == 123456.f) sink[0] = acc;
        if (lane == 0) { clk[1] = tw; clk[2] = clock64() - tt0; }
    }
}
template <int NBS, int SPIN>
void rund(const char* name) {
    float* sink; long long* clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, 16 * 8);
    const int ntiles = 20000;
    for (int rep = 0; rep < 2; rep++) { kd<NBS, SPIN><<<1, 8 * 32>>>(sink, ntiles, clk); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    long long h[16];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-60s %6.2f cycles per sample; consumer: %7.1f cycles per tile, of which %7.1f in the wait  (%s)\n", name, (double)h[0] / (ntiles * 32.0), (double)h[2] / ntiles, (double)h[1] / ntiles, cudaGetErrorString(e));
    cudaFree(sink); cudaFree(clk);
}

// the slot wait through a plain shared-memory flag instead of an mbarrier: the consumer stores the number of tiles it has taken
// (st.volatile.shared after its reads), the producer loads it a tile ahead (ld.volatile.shared: an ordinary scoreboarded load)
template <int NBS>
__global__ void __launch_bounds__(512) kf(float* sink, int ntiles, long long* clk) {
    __shared__ __align__(16) float xs[1024];
    __shared__ __align__(16) float tiles[NBS][TILE * VS];
    __shared__ unsigned long long full[NBS];
    __shared__ volatile int taken;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = sinf(0.37f * i) * 0.5f;
    if (threadIdx.x == 0) { for (int b = 0; b < NBS; b++) mbar_init(&full[b], 1); taken = 0; }
    __syncthreads();
    if (warp == 0) {
        const float w = 0.01f + 0.002f * lane, r = 0.999f;
        const float a1 = -2.f * r * cosf(w), a2 = r * r, b0 = 0.01f, p = 1.0f, g = 0.5f;
        St s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int yb = 0;
        int seen = taken;                       // tiles the consumer has taken, as of the last look
        const long long c0 = clock64();
#pragma unroll 1
        for (int kt = 0; kt < ntiles; kt++) {
            while (kt - seen >= NBS) seen = taken;          // slot of tile kt - NBS not yet taken: look again
            const int look = taken;                          // a look ahead for the next tile, lands during this one
            tile_v0(xs + (kt & 31) * TILE, &tiles[yb][0] + lane, b0, a1, a2, p, g, s);
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[yb]);
            seen = look;
            if (++yb == NBS) yb = 0;
        }
        const long long c1 = clock64();
        if (s.y1 == 123456.f) sink[0] = s.y1 + s.t2;
        if (lane == 0) clk[0] = c1 - c0;
    } else if (warp == 5) {
        int yb = 0; unsigned par = 0;
        float acc = 0.f;
        for (int kt = 0; kt < ntiles; kt++) {
            mbar_wait(&full[yb], par);
            const float v = tiles[yb][lane * VS];
            acc += v;
            asm volatile("" ::"f"(v) : "memory");            // the load has returned
            __syncwarp();
            if (lane == 0) taken = kt + 1;
            if (++yb == NBS) { yb = 0; par ^= 1u; }
        }
        if (acc == 123456.f) sink[0] = acc;
    }
}
template <int NBS>
void runf(const char* name) {
    float* sink; long long* clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, 16 * 8);
    const int ntiles = 20000;
    for (int rep = 0; rep < 2; rep++) { kf<NBS><<<1, 8 * 32>>>(sink, ntiles, clk); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    long long h[16];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-60s %6.2f cycles per sample  (%s)\n", name, (double)h[0] / (ntiles * 32.0), cudaGetErrorString(e));
    cudaFree(sink); cudaFree(clk);
}

// which part of the interaction costs the producer?  PW: producer waits on `empty`; CW: consumer waits on `full`;
// CL: consumer loads from the tile; CA: consumer arrives on `empty`
template <bool PW, bool CW, bool CL, bool CA>
__global__ void __launch_bounds__(512) ke(float* sink, int ntiles, long long* clk) {
    constexpr int NBS = 4;
    __shared__ __align__(16) float xs[1024];
    __shared__ __align__(16) float tiles[NBS][TILE * VS];
    __shared__ unsigned long long full[NBS], empty[NBS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = sinf(0.37f * i) * 0.5f;
    if (threadIdx.x == 0) for (int b = 0; b < NBS; b++) { mbar_init(&full[b], 1); mbar_init(&empty[b], 1); }
    __syncthreads();
    if (warp == 0) {
        const float w = 0.01f + 0.002f * lane, r = 0.999f;
        const float a1 = -2.f * r * cosf(w), a2 = r * r, b0 = 0.01f, p = 1.0f, g = 0.5f;
        St s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int yb = 0; unsigned par = 0;
        const long long c0 = clock64();
#pragma unroll 1
        for (int kt = 0; kt < ntiles; kt++) {
            if (PW) mbar_wait(&empty[yb], par ^ 1u);
            tile_v0(xs + (kt & 31) * TILE, &tiles[yb][0] + lane, b0, a1, a2, p, g, s);
            __syncwarp();
            if (lane == 0) mbar_arrive(&full[yb]);
            if (++yb == NBS) { yb = 0; par ^= 1u; }
        }
        const long long c1 = clock64();
        if (s.y1 == 123456.f) sink[0] = s.y1 + s.t2;
        if (lane == 0) clk[0] = c1 - c0;
    } else if (warp == 5) {
        int yb = 0; unsigned par = 0;
        float acc = 0.f;
        for (int kt = 0; kt < ntiles; kt++) {
            if (CW) mbar_wait(&full[yb], par); else __nanosleep(250);
            if (CL) acc += tiles[yb][lane * VS];
            __syncwarp();
            if (CA && lane == 0) mbar_arrive(&empty[yb]);
            if (++yb == NBS) { yb = 0; par ^= 1u; }
        }
        if (acc == 123456.f) sink[0] = acc;
    }
}
template <bool PW, bool CW, bool CL, bool CA>
void rune(const char* name) {
    float* sink; long long* clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, 16 * 8);
    const int ntiles = 20000;
    for (int rep = 0; rep < 2; rep++) { ke<PW, CW, CL, CA><<<1, 8 * 32>>>(sink, ntiles, clk); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    long long h[16];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-76s %6.2f cycles per sample  (%s)\n", name, (double)h[0] / (ntiles * 32.0), cudaGetErrorString(e));
    cudaFree(sink); cudaFree(clk);
}

// what does another warp's shared-memory reading cost the producer?  No barriers; the reader does NL loads of width 4 or 16 bytes
// every `nap` ns from the tile slots (WHERE 0), from the input ring the producer also reads (1), or from a region nobody else touches (2)
template <int NL, bool WIDE, int WHERE>
__global__ void __launch_bounds__(512) kl(float* sink, int ntiles, long long* clk, unsigned nap) {
    __shared__ __align__(16) float xs[1024];
    __shared__ __align__(16) float tiles[4][TILE * VS];
    __shared__ __align__(16) float other[4][TILE * VS];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) xs[i] = sinf(0.37f * i) * 0.5f;
    for (int i = threadIdx.x; i < 4 * TILE * VS; i += blockDim.x) { (&other[0][0])[i] = 1.f; (&tiles[0][0])[i] = 0.f; }
    __shared__ volatile int done;
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    if (warp == 0) {
        const float w = 0.01f + 0.002f * lane, r = 0.999f;
        const float a1 = -2.f * r * cosf(w), a2 = r * r, b0 = 0.01f, p = 1.0f, g = 0.5f;
        St s = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
        int yb = 0;
        const long long c0 = clock64();
#pragma unroll 1
        for (int kt = 0; kt < ntiles; kt++) {
            tile_v0(xs + (kt & 31) * TILE, &tiles[yb][0] + lane, b0, a1, a2, p, g, s);
            if (++yb == 4) yb = 0;
        }
        const long long c1 = clock64();
        if (s.y1 == 123456.f) sink[0] = s.y1 + s.t2;
        if (lane == 0) { clk[0] = c1 - c0; done = 1; }
    } else if (warp == 5) {
        float acc = 0.f;
        int yb = 0;
        long long n = 0;
        while (!done) {
            const float* base = WHERE == 0 ? &tiles[yb][0] : (WHERE == 1 ? xs : &other[yb][0]);
            const float* row = WHERE == 1 ? base + (lane & 7) * 32 : base + lane * VS;
#pragma unroll
            for (int i = 0; i < NL; i++) {
                if (WIDE) { const float4 v = *reinterpret_cast<const float4*>(row + 4 * i); acc += v.x + v.w; }
                else acc += *reinterpret_cast<const volatile float*>(row + i);
            }
            if (++yb == 4) yb = 0;
            n++;
            if (nap) __nanosleep(nap);
        }
        if (acc == 123456.f) sink[0] = acc;
        if (lane == 0) clk[1] = n;
    }
}
template <int NL, bool WIDE, int WHERE>
void runl(const char* name, unsigned nap) {
    float* sink; long long* clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, 16 * 8);
    const int ntiles = 20000;
    for (int rep = 0; rep < 2; rep++) { kl<NL, WIDE, WHERE><<<1, 8 * 32>>>(sink, ntiles, clk, nap); cudaDeviceSynchronize(); }
    cudaError_t e = cudaGetLastError();
    long long h[16];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-70s %6.2f cycles per sample, %5.2f reader passes per tile  (%s)\n", name, (double)h[0] / (ntiles * 32.0), (double)h[1] / ntiles, cudaGetErrorString(e));
    cudaFree(sink); cudaFree(clk);
}

template <int VAR>
void run(const char* name, int warps, int stride) {
    float* sink; long long* clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, 16 * 8);
    const int ntiles = 20000;
    k<VAR><<<1, warps * 32>>>(sink, ntiles, clk, stride);
    cudaDeviceSynchronize();
    k<VAR><<<1, warps * 32>>>(sink, ntiles, clk, stride);
    cudaError_t e = cudaDeviceSynchronize();
    long long h[16];
    cudaMemcpy(h, clk, sizeof(h), cudaMemcpyDeviceToHost);
    printf("%-44s warps %2d (every %d active): %6.2f cycles per sample  (%s)\n", name, warps, stride, (double)h[0] / (ntiles * 32.0), cudaGetErrorString(e));
    cudaFree(sink); cudaFree(clk);
}

int main() {
    run<0>("V0 kernel order", 1, 1);
    run<1>("V1 forward part of the tile first", 1, 1);
    run<2>("V2 forward of group q+1 inside group q", 1, 1);
    run<3>("V3 recurrence only (u given)", 1, 1);
    run<0>("V0, 4 warps = one per scheduler", 4, 1);
    run<2>("V2, 4 warps = one per scheduler", 4, 1);
    run<0>("V0, 8 warps = two per scheduler", 8, 1);
    run<2>("V2, 8 warps = two per scheduler", 8, 1);
    run<0>("V0, 16 warps = four per scheduler", 16, 1);
    run<2>("V2, 16 warps = four per scheduler", 16, 1);
    runb<0>("hand-over per tile, consumer on the same scheduler", 1, 0);
    runb<0>("hand-over per tile, consumer on the next scheduler", 1, 1);
    runb<1>("next slot probed before the tile, consumer same scheduler", 1, 0);
    runb<1>("next slot probed before the tile, consumer next scheduler", 1, 1);
    runb<2>("signal only (consumer idle, never waits)", 1, 1);
    runb<3>("wait only (self-satisfied), no syncwarp", 1, 1);
    runb<4>("hand-over per tile without __syncwarp", 1, 1);
    runb<5>("hand-over every second tile", 1, 1);
    runb<0>("hand-over per tile, 4 producers + 4 consumers", 4, 0);
    runb<1>("next slot probed before the tile, 4 producers + 4 consumers", 4, 0);
    rund<2, 0>("ring of 2 slots, try_wait");
    rund<3, 0>("ring of 3 slots, try_wait");
    rund<4, 0>("ring of 4 slots, try_wait");
    rund<6, 0>("ring of 6 slots, try_wait");
    rund<2, 1>("ring of 2 slots, test_wait spin");
    rund<4, 1>("ring of 4 slots, test_wait spin");
    rund<2, 2>("ring of 2 slots, consumer waits with a 10 ms suspend hint");
    rund<4, 2>("ring of 4 slots, consumer waits with a 10 ms suspend hint");
    rund<2, 3>("ring of 2 slots, both wait with a 10 ms suspend hint");
    rund<4, 3>("ring of 4 slots, both wait with a 10 ms suspend hint");
    rund<4, 4>("ring of 4 slots, consumer polls with 200 ns naps");
    runf<2>("ring of 2 slots, slot wait = shared-memory counter read a tile ahead");
    runf<4>("ring of 4 slots, slot wait = shared-memory counter read a tile ahead");
    rune<false, true, false, false>("consumer only waits on the tile signal (no load, no release); producer never waits");
    rune<false, true, true, false>("consumer waits and loads; producer never waits");
    rune<false, true, true, true>("consumer waits, loads, releases; producer never waits");
    rune<true, true, false, true>("full protocol without the consumer's load");
    rune<true, true, true, true>("full protocol");
    rune<false, false, true, false>("consumer only loads from the tiles every 250 ns (no barriers)");
    runl<1, false, 0>("reader: 1 LDS from the tile slots every 250 ns", 250);
    runl<1, false, 2>("reader: 1 LDS from an untouched region every 250 ns", 250);
    runl<1, false, 1>("reader: 1 LDS from the input ring every 250 ns", 250);
    runl<8, true, 0>("reader: 8 LDS.128 from the tile slots every 250 ns", 250);
    runl<8, true, 2>("reader: 8 LDS.128 from an untouched region every 250 ns", 250);
    runl<8, true, 0>("reader: 8 LDS.128 from the tile slots every 1000 ns", 1000);
    runl<8, true, 0>("reader: 8 LDS.128 from the tile slots, no nap", 0);
    runl<0, false, 2>("reader: no loads, naps of 250 ns only", 250);
    return 0;
}
