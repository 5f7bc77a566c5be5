// issue_bench.cu — how fast can ONE warp issue FP32 instructions on its scheduler (B200)?
//   single warp, N independent chains of FFMA / FMUL+FADD / FFMA2 (packed f32x2), cycles per warp instruction.
// Also 2 and 4 warps on the same scheduler (CTA of 128*k threads, warps w with w % 4 == 0 active).
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND, int CH>
__global__ void k(float* sink, int iters, float a, float b, long long* clk, int stride) {
    const int warp = threadIdx.x >> 5;
    if (warp % stride) return;                       // only warps on scheduler 0 work
    float v[CH];
    unsigned long long p[CH];
#pragma unroll
    for (int i = 0; i < CH; i++) { v[i] = threadIdx.x + i; p[i] = ((unsigned long long)__float_as_uint(v[i]) << 32) | __float_as_uint(v[i] + 0.5f); }
    unsigned long long aa = ((unsigned long long)__float_as_uint(a) << 32) | __float_as_uint(a);
    unsigned long long bb = ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(b);
    const long long c0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; it++) {
#pragma unroll
        for (int r = 0; r < 8; r++) {
#pragma unroll
            for (int i = 0; i < CH; i++) {
                if (KIND == 0) v[i] = __fmaf_rn(v[i], a, b);
                else if (KIND == 1) v[i] = __fadd_rn(__fmul_rn(v[i], a), b);
                else asm volatile("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(p[i]) : "l"(p[i]), "l"(aa), "l"(bb));
            }
        }
    }
    const long long c1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < CH; i++) s += v[i] + __uint_as_float((unsigned)p[i]);
    if (s == 123456.f) sink[0] = s;
    if (threadIdx.x == 0) clk[0] = c1 - c0;
}

template <int KIND, int CH>
void run(const char* name, int warps_on_sched) {
    float* sink; long long* clk;
    cudaMalloc(&sink, 4); cudaMalloc(&clk, 8);
    const int iters = 2000;
    const int nt = 128 * warps_on_sched;             // warps 0,4,8,.. share scheduler 0
    k<KIND, CH><<<1, nt>>>(sink, iters, 0.999f, 0.001f, clk, 4);
    cudaDeviceSynchronize();
    long long c; cudaMemcpy(&c, clk, 8, cudaMemcpyDeviceToHost);
    const double instr = (double)iters * 8 * CH * (KIND == 1 ? 2 : 1);
    printf("%-22s chains=%d warps/sched=%d: %.2f cycles per warp-instruction of one warp, %.2f warp-instr/cycle on the scheduler\n",
           name, CH, warps_on_sched, c / instr, instr * warps_on_sched / c);
    cudaFree(sink); cudaFree(clk);
}

int main() {
    run<0, 8>("FFMA", 1); run<0, 16>("FFMA", 1); run<0, 8>("FFMA", 2); run<0, 8>("FFMA", 4);
    run<1, 8>("FMUL+FADD", 1); run<1, 8>("FMUL+FADD", 2);
    run<2, 8>("FFMA2 (f32x2)", 1); run<2, 16>("FFMA2 (f32x2)", 1); run<2, 8>("FFMA2 (f32x2)", 2); run<2, 8>("FFMA2 (f32x2)", 4);
    run<0, 1>("FFMA dependent", 1); run<2, 1>("FFMA2 dependent", 1);
    return 0;
}
