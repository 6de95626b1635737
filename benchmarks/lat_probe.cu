// lat_probe.cu -- dependent-issue latency (cycles) of the instruction forms that sit on the serial chains of this library
// (modulator phase scan, timing loop), one warp alone on an SM, one dependency chain, measured with clock64().
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o benchmarks/bin/lat_probe benchmarks/lat_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#define N 4096
template <int OP> __device__ __forceinline__ float step(float a, float b) {
    if (OP == 0) { asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b)); }
    if (OP == 1) { asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a) : "f"(b)); }
    if (OP == 2) { asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b)); }
    if (OP == 3) { asm volatile("cvt.rzi.f32.f32 %0, %0;" : "+f"(a)); }                                   // FRND.TRUNC
    if (OP == 4) { unsigned t; asm volatile("lop3.b32 %0, %1, %2, 0x80000000, 0xf8;" : "=r"(t) : "r"(__float_as_uint(a)), "r"(__float_as_uint(b))); a = __uint_as_float(t); }
    if (OP == 5) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(a)); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(a) : "d"(d)); }   // F2F up + down
    if (OP == 6) { double d; asm volatile("cvt.f64.f32 %0, %1;" : "=d"(d) : "f"(a)); asm volatile("mul.rn.f64 %0, %0, 0d3FF0000000000001;" : "+d"(d)); asm volatile("cvt.rn.f32.f64 %0, %1;" : "=f"(a) : "d"(d)); }
    if (OP == 7) { int t; asm volatile("cvt.rzi.s32.f32 %0, %1;" : "=r"(t) : "f"(a)); asm volatile("cvt.rn.f32.s32 %0, %1;" : "=f"(a) : "r"(t)); }   // F2I + I2F
    if (OP == 8) { unsigned t; asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; selp.f32 %0, %1, %2, p;}" : "=r"(t) : "f"(a), "f"(b)); a = __uint_as_float(t); }
    if (OP == 9) { asm volatile("add.rn.f32 %0, %0, 0f4B400000;" : "+f"(a)); asm volatile("add.rn.f32 %0, %0, 0fCB400000;" : "+f"(a)); }             // magic-number round: 2 FADD
    if (OP == 10) { asm volatile("ex2.approx.f32 %0, %0;" : "+f"(a)); }                                    // MUFU
    if (OP == 11) { int t = __float_as_int(a); asm volatile("add.s32 %0, %0, %1;" : "+r"(t) : "r"(__float_as_int(b))); a = __int_as_float(t); }
    if (OP == 12) { int t = __float_as_int(a); asm volatile("shfl.sync.idx.b32 %0, %0, 0, 31, 0xffffffff;" : "+r"(t)); a = __int_as_float(t); }
    if (OP == 13) { unsigned t; asm volatile("{.reg .pred p; setp.gt.f32 p, %1, %2; vote.sync.ballot.b32 %0, p, 0xffffffff;}" : "=r"(t) : "f"(a), "f"(b)); a = __uint_as_float(t | 0x3f800000u); }
    return a;
}
template <int OP> __global__ void k(float *out, long long *cyc, float b) {
    float a = 1.5f + threadIdx.x * 1e-3f;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) a = step<OP>(a, b);
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (a == 123.456f) out[0] = a;
}
// shared-memory load-to-use latency by pointer chasing
__global__ void k_lds(float *out, long long *cyc) {
    __shared__ int nxt[1024];
    for (int i = threadIdx.x; i < 1024; i += 32) nxt[i] = (i + 33) & 1023;
    __syncwarp();
    int p = threadIdx.x;
    long long t0 = clock64();
#pragma unroll 16
    for (int i = 0; i < N; i++) p = nxt[p];
    long long t1 = clock64();
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
    if (p == -1) out[0] = p;
}
template <int OP> void run(const char *name, int per) {
    float *d; long long *c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
    k<OP><<<1, 32>>>(d, c, 1.0000001f); k<OP><<<1, 32>>>(d, c, 1.0000001f);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %.2f cycles per step (%d instruction%s)\n", name, (double)h / N, per, per > 1 ? "s" : "");
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    printf("# %s: dependent-issue latency, one warp\n", pr.name);
    run<0>("FADD", 1); run<1>("FFMA", 1); run<2>("FMUL", 1); run<3>("FRND.TRUNC", 1); run<4>("LOP3", 1); run<11>("IADD", 1);
    run<5>("F2F.F64.F32 + F2F.F32.F64", 2); run<6>("F2F + DMUL + F2F", 3); run<7>("F2I + I2F", 2); run<8>("FSETP + FSEL", 2);
    run<9>("FADD + FADD (magic round)", 2); run<10>("MUFU.EX2", 1); run<12>("SHFL.IDX", 1); run<13>("FSETP + VOTE.BALLOT", 2);
    float *d; long long *c; cudaMalloc(&d, 4); cudaMalloc(&c, 8);
    k_lds<<<1, 32>>>(d, c); k_lds<<<1, 32>>>(d, c);
    long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-34s %.2f cycles per step\n", "LDS (pointer chase)", (double)h / N);
    return 0;
}
