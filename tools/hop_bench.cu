// hop_bench.cu — microbenchmark of the chain hand-over (chain.cuh): G CTAs of one warp form a chain; packet k
// (32 lanes x 16 bytes, LL protocol: payload + tag per 8-byte word) travels from CTA 0 to CTA G-1, every CTA
// spending `work` cycles of dependent arithmetic on it.  Reports cycles per packet per stage (the stage cycle that
// bounds the ordered mixdown's throughput) and the one-packet latency per hop, for several load/store flavours
// and prefetch depths.   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o hop_bench hop_bench.cu
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s -> %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

template <int MODE> __device__ __forceinline__ void ld4(const unsigned* p, unsigned& a, unsigned& b, unsigned& c, unsigned& d) {
    if (MODE == 0) asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    else if (MODE == 1) asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
    else asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a), "=r"(b), "=r"(c), "=r"(d) : "l"(p) : "memory");
}
template <int MODE> __device__ __forceinline__ void st4(unsigned* p, unsigned a, unsigned b, unsigned c, unsigned d) {
    if (MODE == 0) asm volatile("st.volatile.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
    else if (MODE == 1) asm volatile("st.relaxed.gpu.global.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
    else asm volatile("st.global.cg.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(p), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

constexpr int RING = 16;      // packets per link

// DEPTH = packets requested ahead of the one being worked on
template <int MODE, int DEPTH>
__global__ void __launch_bounds__(32) chain_kernel(unsigned* ring, int G, int K, int work, long long* t_out, float* sink) {
    const int g = blockIdx.x, lane = threadIdx.x;
    unsigned* in = ring + (size_t)g * RING * 128;            // link g: written by CTA g-1
    unsigned* out = ring + (size_t)(g + 1) * RING * 128;
    unsigned w[DEPTH + 1][4];
    const long long t0 = clock64();
    if (g > 0)
        for (int d = 0; d < DEPTH && d < K; d++) ld4<MODE>(in + (d % RING) * 128 + lane * 4, w[d][0], w[d][1], w[d][2], w[d][3]);
    float acc = 0.f;
#pragma unroll 1
    for (int k = 0; k < K; k++) {
        const unsigned tag = (unsigned)k + 1u;
        unsigned a = k, c = k;
        if (g > 0) {
            const int slot = k % (DEPTH + 1);
            unsigned* p = in + (k % RING) * 128 + lane * 4;
            if (DEPTH == 0) ld4<MODE>(p, w[0][0], w[0][1], w[0][2], w[0][3]);
            // rotate: DEPTH is tiny, select by compile-time unrolled compare
            unsigned x0 = 0, x1 = 0, x2 = 0, x3 = 0;
#pragma unroll
            for (int s = 0; s <= DEPTH; s++) if (s == slot) { x0 = w[s][0]; x1 = w[s][1]; x2 = w[s][2]; x3 = w[s][3]; }
            unsigned spins = 0;
            while (((int)(x1 - tag) < 0 || (int)(x3 - tag) < 0) && ++spins < (1u << 22)) ld4<MODE>(p, x0, x1, x2, x3);   // a lapped packet (newer tag) ends the wait too: no hang
            a = x0; c = x2;
            // request packet k + DEPTH (+1 slot ahead of the ones in flight)
            if (DEPTH > 0 && k + DEPTH < K) {
                const int ns = (k + DEPTH) % (DEPTH + 1);
                unsigned* q = in + ((k + DEPTH) % RING) * 128 + lane * 4;
                unsigned y0, y1, y2, y3;
                ld4<MODE>(q, y0, y1, y2, y3);
#pragma unroll
                for (int s = 0; s <= DEPTH; s++) if (s == ns) { w[s][0] = y0; w[s][1] = y1; w[s][2] = y2; w[s][3] = y3; }
            }
        }
        float v = __uint_as_float(a) * 1e-30f + acc;
        for (int i = 0; i < work; i++) v = __fadd_rn(v, 1.0f);   // `work` dependent adds = 4 cycles each
        acc = v * 1e-30f;
        if (g + 1 < G) st4<MODE>(out + (k % RING) * 128 + lane * 4, a + (unsigned)(acc != 12345.f ? 0 : 1), tag, c, tag);
        // no back-pressure: K <= RING * (something) is not required because tags only grow and the reader is
        // never more than the ring behind in this benchmark (checked on the host by the final packet value)
    }
    const long long t1 = clock64();
    if (lane == 0) { t_out[2 * g] = t0; t_out[2 * g + 1] = t1; }
    if (acc == 999.f) sink[0] = acc;
}

template <int MODE, int DEPTH>
static int run(const char* name, int G, int K, int work) {
    unsigned* ring; long long* t; float* sink;
    CK(cudaMalloc(&ring, (size_t)(G + 1) * RING * 128 * 4));
    CK(cudaMemset(ring, 0, (size_t)(G + 1) * RING * 128 * 4));
    CK(cudaMalloc(&t, sizeof(long long) * 2 * G));
    CK(cudaMalloc(&sink, 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    chain_kernel<MODE, DEPTH><<<G, 32>>>(ring, G, K, work, t, sink);
    cudaEventRecord(e1);
    CK(cudaDeviceSynchronize());
    float ms = 0;
    cudaEventElapsedTime(&ms, e0, e1);
    printf("%-28s G=%3d K=%6d work=%4d  %8.3f ms  -> %8.1f ns per packet (stage cycle), %8.1f ns per hop if K==1\n", name, G, K, work, ms,
           ms * 1e6 / K, ms * 1e6 / G);
    cudaFree(ring); cudaFree(t); cudaFree(sink);
    return 0;
}

int main() {
    // NOTE: without back-pressure the writer may lap the reader when the reader is slower; K is kept moderate and
    // every stage does the same work, so the reader keeps pace.  (RING = 16 packets of slack.)
    const int G = 128;
    for (int work : {0, 100, 400}) {
        run<0, 0>("volatile, no prefetch", G, 20000, work);
        run<0, 1>("volatile, prefetch 1", G, 20000, work);
        run<0, 2>("volatile, prefetch 2", G, 20000, work);
        run<1, 0>("relaxed.gpu, no prefetch", G, 20000, work);
        run<1, 1>("relaxed.gpu, prefetch 1", G, 20000, work);
        run<1, 2>("relaxed.gpu, prefetch 2", G, 20000, work);
        run<2, 1>(".cg, prefetch 1", G, 20000, work);
    }
    run<0, 0>("volatile latency", G, 1, 0);
    run<1, 0>("relaxed.gpu latency", G, 1, 0);
    run<2, 0>(".cg latency", G, 1, 0);
    run<0, 0>("volatile latency G=16", 16, 1, 0);
    return 0;
}
