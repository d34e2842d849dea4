// Microbenchmark: throughput of MATCH.ANY vs the ballot-per-bit equivalent (8-bit keys), full occupancy.
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE, int BITS>
__global__ void k(unsigned *out, int iters) {
    unsigned x = threadIdx.x * 2654435761u + blockIdx.x * 40503u, acc = 0;
    const unsigned lane = threadIdx.x & 31;
    for (int i = 0; i < iters; i++) {
        x = x * 1664525u + 1013904223u;
        unsigned d = (x >> 13) & ((1u << BITS) - 1u);
        unsigned peers;
        if (MODE == 0) peers = __match_any_sync(0xffffffffu, d);
        else {
            peers = 0xffffffffu;
#pragma unroll
            for (int b = 0; b < BITS; b++) { const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u); peers &= ((d >> b) & 1u) ? bal : ~bal; }
        }
        acc += __popc(peers & ((1u << lane) - 1u));
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = acc;
}
template <int MODE, int BITS> void run(const char *name) {
    unsigned *d; cudaMalloc(&d, 148 * 8 * 256 * 4);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int iters = 20000;
    k<MODE, BITS><<<148 * 8, 256>>>(d, 100);
    cudaEventRecord(a); k<MODE, BITS><<<148 * 8, 256>>>(d, iters); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double warp_ops = 148.0 * 8 * 8 * iters;
    printf("%-28s %8.3f ms  %.2f SM-cycles per warp-op (at 1.965 GHz), %.1f G warp-ops/s\n", name, ms, ms * 1e-3 * 1.965e9 * 148 / warp_ops, warp_ops / ms * 1e-6);
    cudaFree(d);
}
int main() {
    run<0, 8>("match_any 8-bit keys");
    run<0, 2>("match_any 2-bit keys");
    run<1, 8>("8 ballots");
    run<1, 7>("7 ballots");
    return 0;
}
