// common.cu — error text, pointer classification, staging buffers, FMA-peak microbenchmark.
#include "hg_common.cuh"

namespace hg {

char* err_buf() {
    static thread_local char buf[512] = {0};
    return buf;
}

hg_status fail(hg_status code, const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(err_buf(), 512, fmt, ap);
    va_end(ap);
    return code;
}

bool is_device_ptr(const void* p) {
    if (!p) return false;
    cudaPointerAttributes a;
    cudaError_t e = cudaPointerGetAttributes(&a, p);
    if (e != cudaSuccess) {
        cudaGetLastError();
        return false;
    }
    return a.type == cudaMemoryTypeDevice || a.type == cudaMemoryTypeManaged;
}

hg_status DevBuf::reserve(size_t bytes) {
    if (bytes <= cap) return HG_OK;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc(&p, bytes);
    if (e != cudaSuccess) return fail(HG_ERR_NOMEM, "cudaMalloc(%zu) -> %s", bytes, cudaGetErrorString(e));
    cap = bytes;
    return HG_OK;
}

void DevBuf::release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
}

// ---------------------------------------------------------------------------------------------
// FMA issue-rate microbenchmark: 8 independent register chains per thread, 3-register FMAs with
// loop-invariant multiplicands.  Slots = threads * chains * iterations.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(256) fma_peak_kernel(T* sink, int iters, T a, T b, long long* clk) {
    T v0 = (T)threadIdx.x, v1 = v0 + 1, v2 = v0 + 2, v3 = v0 + 3, v4 = v0 + 4, v5 = v0 + 5, v6 = v0 + 6, v7 = v0 + 7;
    long long c0 = clock64();
    unsigned long long g0;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g0));
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 8; k++) {
            v0 = fma(v0, a, b); v1 = fma(v1, a, b); v2 = fma(v2, a, b); v3 = fma(v3, a, b);
            v4 = fma(v4, a, b); v5 = fma(v5, a, b); v6 = fma(v6, a, b); v7 = fma(v7, a, b);
        }
    }
    long long c1 = clock64();
    unsigned long long g1;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(g1));
    T r = ((v0 + v1) + (v2 + v3)) + ((v4 + v5) + (v6 + v7));
    if (r == (T)123456789) sink[0] = r;   // never true; keeps the chains alive
    if (blockIdx.x == 0 && threadIdx.x == 0) { clk[0] = c1 - c0; clk[1] = (long long)(g1 - g0); }
}

// ---------------------------------------------------------------------------------------------
// Dependent-chain microbenchmark: one warp per scheduler runs the biquad recurrence's critical path
//     y <- fl(u - fl(fl(a1*y) + t2))        FMUL -> FADD -> FADD, each consuming the previous result
// with nothing else to issue.  Cycles per iteration is the floor of ANY bit-exact serial evaluation of
// one filter lane (the arithmetic may not be re-associated or fused), measured instead of assumed.
// ---------------------------------------------------------------------------------------------
template <typename T>
__global__ void __launch_bounds__(128) chain_kernel(T* sink, int iters, T a1, T u, T t2, long long* clk) {
    T y = (T)threadIdx.x * (T)1e-3;
    const long long c0 = clock64();
#pragma unroll 1
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int k = 0; k < 16; k++) {
            if (sizeof(T) == 4) y = (T)__fsub_rn((float)u, __fadd_rn(__fmul_rn((float)a1, (float)y), (float)t2));
            else y = (T)__dsub_rn((double)u, __dadd_rn(__dmul_rn((double)a1, (double)y), (double)t2));
        }
    }
    const long long c1 = clock64();
    if (y == (T)123456789) sink[0] = y;
    if (threadIdx.x == 0 && blockIdx.x == 0) clk[0] = c1 - c0;
}

template <typename T>
static hg_status chain_T(int device, double* cycles_per_step) {
    HG_CUDA(cudaSetDevice(device));
    T* sink;
    long long* clk;
    HG_CUDA(cudaMalloc(&sink, sizeof(T)));
    HG_CUDA(cudaMalloc(&clk, sizeof(long long)));
    const int iters = 4096;
    double best = 1e30;
    for (int rep = 0; rep < 4; rep++) {
        chain_kernel<T><<<1, 128>>>(sink, iters, (T)0.999, (T)0.5, (T)0.25, clk);
        HG_CUDA(cudaDeviceSynchronize());
        long long cyc = 0;
        HG_CUDA(cudaMemcpy(&cyc, clk, sizeof(cyc), cudaMemcpyDeviceToHost));
        const double c = (double)cyc / ((double)iters * 16.0);
        if (c < best) best = c;
    }
    *cycles_per_step = best;
    cudaFree(sink);
    cudaFree(clk);
    return HG_OK;
}

template <typename T>
static hg_status fma_peak_T(int device, double* slots, int* sms, double* mhz) {
    HG_CUDA(cudaSetDevice(device));
    cudaDeviceProp prop;
    HG_CUDA(cudaGetDeviceProperties(&prop, device));
    T* sink;
    long long* clk;
    HG_CUDA(cudaMalloc(&sink, sizeof(T)));
    HG_CUDA(cudaMalloc(&clk, 2 * sizeof(long long)));
    const int ctas = prop.multiProcessorCount * 8, nt = 256;
    const int iters = sizeof(T) == 4 ? 4096 : 2048;
    cudaEvent_t e0, e1;
    HG_CUDA(cudaEventCreate(&e0));
    HG_CUDA(cudaEventCreate(&e1));
    double best = 0, best_mhz = 0;
    for (int rep = 0; rep < 6; rep++) {
        HG_CUDA(cudaEventRecord(e0));
        fma_peak_kernel<T><<<ctas, nt>>>(sink, iters, (T)0.999, (T)0.001, clk);
        HG_CUDA(cudaEventRecord(e1));
        HG_CUDA(cudaEventSynchronize(e1));
        float ms = 0;
        HG_CUDA(cudaEventElapsedTime(&ms, e0, e1));
        double s = (double)ctas * nt * 64.0 * iters / (ms * 1e-3);
        long long cyc[2] = {0, 0};
        HG_CUDA(cudaMemcpy(cyc, clk, sizeof(cyc), cudaMemcpyDeviceToHost));
        if (rep >= 1 && s > best) {
            best = s;
            // SM clock seen by block 0: its cycle counter against the nanosecond global timer
            best_mhz = cyc[1] > 0 ? (double)cyc[0] / (double)cyc[1] * 1e3 : 0;
        }
    }
    *slots = best;
    if (sms) *sms = prop.multiProcessorCount;
    if (mhz) *mhz = best_mhz;
    cudaEventDestroy(e0);
    cudaEventDestroy(e1);
    cudaFree(sink);
    cudaFree(clk);
    return HG_OK;
}

}  // namespace hg

extern "C" {

const char* hg_last_error(void) { return hg::err_buf(); }

int hg_device_count(void) {
    int n = 0;
    if (cudaGetDeviceCount(&n) != cudaSuccess) {
        cudaGetLastError();
        return 0;
    }
    return n;
}

hg_status hg_chain_latency(int device, int dtype, double* cycles_per_sample) {
    if (!cycles_per_sample) return hg::fail(HG_ERR_INVALID, "hg_chain_latency: null output");
    if (dtype == HG_F32) return hg::chain_T<float>(device, cycles_per_sample);
    if (dtype == HG_F64) return hg::chain_T<double>(device, cycles_per_sample);
    return hg::fail(HG_ERR_INVALID, "hg_chain_latency: bad dtype %d", dtype);
}

hg_status hg_fma_peak(int device, int dtype, double* slots_per_s, int* sm_count, double* sm_mhz) {
    if (!slots_per_s) return hg::fail(HG_ERR_INVALID, "hg_fma_peak: null output");
    if (dtype == HG_F32) return hg::fma_peak_T<float>(device, slots_per_s, sm_count, sm_mhz);
    if (dtype == HG_F64) return hg::fma_peak_T<double>(device, slots_per_s, sm_count, sm_mhz);
    return hg::fail(HG_ERR_INVALID, "hg_fma_peak: bad dtype %d", dtype);
}
}
