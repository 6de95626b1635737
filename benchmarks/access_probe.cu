// benchmarks/access_probe.cu -- build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o access_probe access_probe.cu
// Result on B200 (round 1): 80-byte bursts 5.86 TB/s, 160 B 5.87, 320 B 5.70, 640 B 7.0, plain streaming read 6.29 TB/s:
// the front end's cooperative row walk is not limited by its memory access pattern.
// access-pattern probe: every warp owns 32 rows of 7680 B (consecutive rows), and walks them in steps of BURST bytes per row,
// loading the 32 x BURST bytes of a step with coalesced 16-byte pieces (piece p = lane + 32 k -> row p / PPR, piece p % PPR).
#include <cuda_runtime.h>
#include <stdio.h>
#include <stdint.h>
template <int BURST, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) k_probe(const uint4 *__restrict__ in, int64_t nrows, unsigned *out) {
    constexpr int PPR = BURST / 16, NP = 32 * PPR / 32;      // pieces per row, pieces per lane per step
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    const int64_t row0 = ((int64_t)blockIdx.x * WARPS + wid) * 32;
    if (row0 >= nrows) return;
    uint32_t off[NP];
#pragma unroll
    for (int k = 0; k < NP; k++) { const int p = lane + 32 * k, r = p / PPR; off[k] = (uint32_t)(r * 480 + (p - PPR * r)); }
    const uint4 *base = in + row0 * 480;
    unsigned acc = 0;
    for (int step = 0; step < 7680 / BURST; step++) {
        uint4 v[NP];
#pragma unroll
        for (int k = 0; k < NP; k++) v[k] = __ldg(base + off[k] + step * PPR);
#pragma unroll
        for (int k = 0; k < NP; k++) acc += v[k].x ^ v[k].y ^ v[k].z ^ v[k].w;
    }
    if (acc == 0x12345678u) out[0] = acc;
}
// plain streaming read for reference
__global__ void k_stream(const uint4 *__restrict__ in, int64_t n16, unsigned *out) {
    unsigned acc = 0;
    for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += (int64_t)gridDim.x * blockDim.x) { uint4 v = __ldg(in + i); acc += v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) out[0] = acc;
}
template <int BURST, int WARPS> float run(const uint4 *d, int64_t nrows, unsigned *o) {
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const unsigned grid = (unsigned)((nrows / 32 + WARPS - 1) / WARPS);
    for (int i = 0; i < 3; i++) k_probe<BURST, WARPS><<<grid, WARPS * 32>>>(d, nrows, o);
    cudaEventRecord(a);
    for (int i = 0; i < 10; i++) k_probe<BURST, WARPS><<<grid, WARPS * 32>>>(d, nrows, o);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    return ms / 10;
}
int main() {
    const int64_t nrows = 256000; const size_t bytes = (size_t)nrows * 7680;
    uint4 *d; unsigned *o; cudaMalloc(&d, bytes); cudaMalloc(&o, 4); cudaMemset(d, 1, bytes);
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    for (int i = 0; i < 3; i++) k_stream<<<148 * 8, 256>>>(d, bytes / 16, o);
    cudaEventRecord(a); for (int i = 0; i < 10; i++) k_stream<<<148 * 8, 256>>>(d, bytes / 16, o); cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b); ms /= 10;
    printf("stream read: %.3f ms %.0f GB/s\n", ms, bytes / ms / 1e6);
#define R(B, W) { float t = run<B, W>(d, nrows, o); printf("burst %4d B, %d warps/CTA: %.3f ms %.0f GB/s\n", B, W, t, bytes / t / 1e6); }
    R(80, 4) R(160, 4) R(240, 4) R(320, 4) R(480, 4) R(640, 4) R(1280, 4) R(80, 8) R(320, 8) R(80, 2) R(80, 1)
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
