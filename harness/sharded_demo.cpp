// sharded_demo.cpp — one host process renders a filter-partitioned Filterbank<float> on all the GPUs of the box
// (BASELINE configs[2]: 65,536 log-spaced resonators over 1/2/4/8 B200).  The shape of a C++ audio host:
// hg_comm_init_all, one soundmath::Filterbank shard per device, device buffers, asynchronous launches — the kernels
// of the ranks run concurrently and hand the running bank-sum row from GPU to GPU in filter order, so the last
// device ends up with the bit-identical sum of the single 65,536-filter bank.
//     sharded_demo [--filters N] [--seconds S] [--gpus G] [--reduce] [--out file.f32]
#include <chrono>
#include <cstring>
#include <cuda_runtime.h>
#include "../include/huygens/filterbank.h"
#include "../include/huygens/offline.h"

using namespace soundmath;

#define CK(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { std::fprintf(stderr, "%s -> %s\n", #x, cudaGetErrorString(e__)); return 1; } } while (0)

int main(int argc, char** argv) {
    int N = 65536, G = 0, reduce = 0;
    double seconds = 10;
    const char* out_path = nullptr;
    for (int i = 1; i < argc; i++) {
        if (!std::strcmp(argv[i], "--filters") && i + 1 < argc) N = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--seconds") && i + 1 < argc) seconds = std::atof(argv[++i]);
        else if (!std::strcmp(argv[i], "--gpus") && i + 1 < argc) G = std::atoi(argv[++i]);
        else if (!std::strcmp(argv[i], "--out") && i + 1 < argc) out_path = argv[++i];
        else if (!std::strcmp(argv[i], "--reduce")) reduce = 1;
        else { std::fprintf(stderr, "unknown argument %s\n", argv[i]); return 2; }
    }
    if (G <= 0) G = hg_device_count();
    if (G < 1) { std::fprintf(stderr, "no CUDA device\n"); return 1; }
    const size_t L = (size_t)(seconds * SR);
    try {
        std::vector<hg_comm*> comms(G, nullptr);
        hg_check(hg_comm_init_all(comms.data(), G, nullptr));
        // the bank of tests/filterbank.cpp:255-271 (forward {1,0,-1}, gains 1/|H(f)| in mix), log-spaced 27.5 Hz .. 20 kHz
        std::vector<float> x = chirp<float>(L);
        std::vector<Filterbank<float>> shards;
        std::vector<float*> d_in(G, nullptr), d_out(G, nullptr);
        std::vector<cudaStream_t> streams(G);
        for (int r = 0; r < G; r++) {
            const int base = N / G, extra = N % G;
            const int lo = r * base + std::min(r, extra), n = base + (r < extra ? 1 : 0);
            CK(cudaSetDevice(r));
            CK(cudaStreamCreate(&streams[r]));
            shards.emplace_back(2, n, 0.001, 1.0, r);
            std::vector<float> gains(n);
            for (int i = 0; i < n; i++) {
                const double f = N == 1 ? 27.5 : 27.5 * std::pow(20000.0 / 27.5, (double)(lo + i) / (N - 1));
                const double w = 2 * PI * f / SR, rad = 0.999;
                const std::complex<double> z(std::cos(w), -std::sin(w));
                gains[i] = (float)(1.0 / std::abs((1.0 - z * z) / (1.0 - 2 * rad * std::cos(w) * z + rad * rad * z * z)));
                shards[r].coefficients(i, {1.0f, 0.0f, -1.0f}, {(float)(-2 * rad * std::cos(w)), (float)(rad * rad)});
            }
            shards[r].mix(gains);
            shards[r].boost(std::vector<float>(n, 1.0f));
            CK(cudaMalloc(&d_in[r], L * sizeof(float)));
            CK(cudaMemcpy(d_in[r], x.data(), L * sizeof(float), cudaMemcpyHostToDevice));
            CK(cudaMalloc(&d_out[r], L * sizeof(float)));
        }
        const int mode = reduce ? HG_SHARD_REDUCE : HG_SHARD_CHAIN, root = reduce ? 0 : G - 1;
        double best_ms = 1e30;
        const int reps = out_path ? 1 : 4;              // --out: the file holds the FIRST block (zero initial state), for the parity test
        for (int rep = 0; rep < reps; rep++) {
            for (int r = 0; r < G; r++) { CK(cudaSetDevice(r)); CK(cudaStreamSynchronize(streams[r])); }
            const auto t0 = std::chrono::steady_clock::now();
            // one thread issues every rank's call: NCCL collectives (reduce mode) must then be grouped
            hg_check(hg_comm_group_start());
            for (int r = 0; r < G; r++) {
                CK(cudaSetDevice(r));
                shards[r].process_sharded(comms[r], d_in[r], r == root ? d_out[r] : nullptr, L, mode, root, streams[r]);
            }
            hg_check(hg_comm_group_end());
            for (int r = 0; r < G; r++) { CK(cudaSetDevice(r)); CK(cudaStreamSynchronize(streams[r])); }
            const double ms = std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now() - t0).count();
            if ((rep > 0 || reps == 1) && ms < best_ms) best_ms = ms;
        }
        std::vector<float> y(L);
        CK(cudaSetDevice(root));
        CK(cudaMemcpy(y.data(), d_out[root], L * sizeof(float), cudaMemcpyDeviceToHost));
        double peak = 0;
        for (size_t t = 0; t < L; t++) peak = std::max(peak, (double)std::fabs(y[t]));
        if (out_path) {
            FILE* f = std::fopen(out_path, "wb");
            if (!f) { std::fprintf(stderr, "cannot write %s\n", out_path); return 1; }
            std::fwrite(y.data(), sizeof(float), L, f);
            std::fclose(f);
        }
        std::printf("{\"filters\": %d, \"gpus\": %d, \"frames\": %zu, \"exchange\": \"%s\", \"ms_per_block\": %.3f, \"filter_samples_per_s\": %.4g, \"peak\": %.6g, \"blocks_rendered\": %d}\n",
                    N, G, L, reduce ? "ncclReduce" : "chain", best_ms, (double)N * L / (best_ms * 1e-3), peak, reps);
        for (int r = 0; r < G; r++) { cudaSetDevice(r); cudaFree(d_in[r]); cudaFree(d_out[r]); }
        shards.clear();
        for (int r = 0; r < G; r++) hg_comm_destroy(comms[r]);
    } catch (const std::exception& e) {
        std::fprintf(stderr, "error: %s\n", e.what());
        return 1;
    }
    return 0;
}
