// fp32_rate.cu -- issue rate of the FP32 instruction forms the RX kernels are made of, per SM sub-partition (scheduler), on this GPU.
// Each thread runs 8 independent dependency chains of one instruction form (so latency is covered with one warp per scheduler);
// the kernel is timed with 1, 2 and 4 warps per scheduler.  Output: warp-instructions per cycle per scheduler.
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o benchmarks/bin/fp32_rate benchmarks/fp32_rate.cu
#include <cstdio>
#include <cuda_runtime.h>
#define CHAINS 8
#define ITERS 2048
typedef unsigned long long u64;
template <int OP> __device__ __forceinline__ void op(float &a, u64 &p, float b, u64 q) {
    if (OP == 0) asm volatile("add.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b));                    // FADD reg, reg
    if (OP == 1) asm volatile("mul.rn.f32 %0, %0, %1;" : "+f"(a) : "f"(b));                    // FMUL reg, reg
    if (OP == 2) asm volatile("mul.rn.f32 %0, %0, 0f3F800001;" : "+f"(a));                     // FMUL reg, imm
    if (OP == 3) asm volatile("fma.rn.f32 %0, %0, %1, %1;" : "+f"(a) : "f"(b));                // FFMA reg, reg, reg
    if (OP == 4) asm volatile("fma.rn.f32 %0, %0, 0f3F800001, 0f33800000;" : "+f"(a));         // FFMA reg, imm, imm -> ptxas may use a register
    if (OP == 5) asm volatile("mul.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));                  // FMUL2 reg, reg
    if (OP == 6) asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(p) : "l"(q));                  // FADD2 reg, reg
    if (OP == 7) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(p) : "l"(q));              // FFMA2 reg, reg, reg
    if (OP == 8) { int t; asm volatile("add.s32 %0, %1, %2;" : "=r"(t) : "r"(__float_as_int(a)), "r"(__float_as_int(b))); a = __int_as_float(t); }   // IADD3
    if (OP == 9) { unsigned t; asm volatile("setp.gt.f32 p9, %1, %2; selp.f32 %0, %1, %2, p9;" : "=r"(t) : "f"(a), "f"(b)); a = __uint_as_float(t); }
}
template <int OP> __global__ void k(float *out, float b0, int iters) {
    float a[CHAINS]; u64 p[CHAINS];
    const float b = b0 + threadIdx.x * 1e-9f;
    u64 q; asm("mov.b64 %0, {%1, %1};" : "=l"(q) : "f"(b));
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { a[c] = 1.0f + c; asm("mov.b64 %0, {%1, %1};" : "=l"(p[c]) : "f"(a[c])); }
    for (int i = 0; i < iters; i++) {
#pragma unroll
        for (int u = 0; u < 8; u++)
#pragma unroll
            for (int c = 0; c < CHAINS; c++) op<OP>(a[c], p[c], b, q);
    }
    float s = 0; 
#pragma unroll
    for (int c = 0; c < CHAINS; c++) { float lo, hi; asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(p[c])); s += a[c] + lo + hi; }
    if (s == 123.456f) out[0] = s;
}
template <int OP> void run(const char *name, int sms, double ghz) {
    float *d; cudaMalloc(&d, 4);
    for (int wps = 1; wps <= 4; wps *= 2) {
        cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
        k<OP><<<sms, 128 * wps>>>(d, 1.0000001f, 16);
        cudaEventRecord(e0);
        k<OP><<<sms, 128 * wps>>>(d, 1.0000001f, ITERS);
        cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        const double instr_per_sched = (double)ITERS * 8 * CHAINS * wps;          // warp-instructions per scheduler
        printf("%-22s warps/scheduler %d: %.3f warp-instr/cycle/scheduler\n", name, wps, instr_per_sched / (ms * 1e-3 * ghz * 1e9));
    }
}
int main() {
    cudaDeviceProp pr; cudaGetDeviceProperties(&pr, 0);
    int khz = 0; cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const double ghz = khz / 1e6;
    printf("# %s, %d SMs, %.3f GHz (attribute clock; rates are relative to it)\n", pr.name, pr.multiProcessorCount, ghz);
    const int sms = pr.multiProcessorCount;
    run<0>("FADD r,r", sms, ghz); run<1>("FMUL r,r", sms, ghz); run<2>("FMUL r,imm", sms, ghz); run<3>("FFMA r,r,r", sms, ghz); run<4>("FFMA r,imm,imm", sms, ghz);
    run<5>("FMUL2 r,r", sms, ghz); run<6>("FADD2 r,r", sms, ghz); run<7>("FFMA2 r,r,r", sms, ghz); run<8>("IADD r,r", sms, ghz);
    return 0;
}
