// eq.cuh -- m17_equalize.cpp: the 5-tap T/2-spaced square-root-Kalman (RLS) equaliser, state and ONE training step.
// Used by the stand-alone batch primitive (k_eq_train, tx.cuh) and by the equaliser option of the live receive chain
// (k_sync_frame<.., EQ = true>, sync.cuh).  The recursion is serial per channel; every operation rounds as the reference's does
// (the library is built with -fmad=false).
#pragma once
#include "afc.cuh"

struct EqState { float c[5], g[5], u[5][5], d[5], E, q, y, fbr, samples[5]; };
// the equaliser option of the live chain also carries the matched filter's output half a symbol before the next symbol instant
// across the block boundary (sync.cuh)
struct RxEqState { EqState e; float mid; float pad[3]; };

// eq_train_known / eq_train_unknown (m17_equalize.cpp:163-213) for one (half-symbol, symbol) pair: eq_update_samples :152-159,
// eq_equalize :122-135, the decision :195-205 (double literals) unless the symbol is known, eq_k_update :105-121 with
// eq_k_calculate :40-100.  Returns the equalised symbol.
__device__ __forceinline__ float eq_step(EqState &e, float in0, float in1, bool known, float train) {
    e.samples[0] = e.samples[2]; e.samples[1] = e.samples[3]; e.samples[2] = e.samples[4];
    e.samples[3] = in0; e.samples[4] = in1;
    float sym = e.samples[0] * e.c[0];
#pragma unroll
    for (int i = 1; i < 5; i++) sym += e.samples[i] * e.c[i];
    float tr;
    if (known) tr = train;
    else if (sym > 0) tr = ((double)sym >= 0.66) ? 1.0f : 0.333f;
    else tr = ((double)sym <= -0.66) ? -1.0f : -0.333f;
    float err = tr - sym;
    float f[5], h[5], a[5];
    const float *x = e.samples;
    f[0] = x[0];
#pragma unroll
    for (int j = 1; j < 5; j++) { f[j] = e.u[0][j] * x[0] + x[j]; for (int i = 1; i < j; i++) f[j] += e.u[i][j] * x[i]; }
#pragma unroll
    for (int j = 0; j < 5; j++) e.g[j] = e.d[j] * f[j];
    a[0] = e.E + e.g[0] * f[0];
#pragma unroll
    for (int j = 1; j < 5; j++) a[j] = a[j - 1] + e.g[j] * f[j];
    const float hq = 1 + e.q, ht = a[4] * e.q;
    e.y = 1.0f / (a[0] + ht);
    e.d[0] = e.d[0] * hq * (e.E + ht) * e.y;
#pragma unroll
    for (int j = 1; j < 5; j++) {
        const float B = a[j - 1] + ht;
        h[j] = -f[j] * e.y;
        e.y = 1.0f / (a[j] + ht);
        e.d[j] = e.d[j] * hq * B * e.y;
        for (int i = 0; i < j; i++) { const float B0 = e.u[i][j]; e.u[i][j] = B0 + h[j] * e.g[i]; e.g[i] += e.g[j] * B0; }
    }
    err *= e.y;
#pragma unroll
    for (int i = 0; i < 5; i++) e.c[i] += err * e.g[i];
    e.fbr = tr;
    return sym;
}
// eq_open :217-224 (full: the statics start at zero, q = 0.08, E = 0.01) / eq_reset :137-141 / eq_restart :142-145 (coffs kept)
__device__ __forceinline__ void eq_init(EqState &e, bool full, bool coffs) {
    if (full) {
        float *w = (float *)&e;
        for (int i = 0; i < (int)(sizeof(EqState) / 4); i++) w[i] = 0.0f;
        e.q = 0.08f; e.E = 0.01f;
    }
    for (int j = 0; j < 5; j++) { for (int i = 0; i < j; i++) e.u[i][j] = 0.0f; e.d[j] = 0.1f; }   // eq_k_reset_ud :25-36
    if (coffs) for (int i = 0; i < 5; i++) e.c[i] = 0.0f;                                          // eq_k_reset_coffs :14-22
}
// The block's n symbols through eq_train_unknown, in place: out[q] <- equalised symbol of the pair (mid[q], out[q]).  Every lane
// runs the same recursion on the same values (broadcast reads); lane 0 writes.  Not inlined: its ~50 live floats stay out of the
// timing-loop kernel's register allocation.
__device__ __noinline__ void eq_block(EqState *gs, float *out, const float *mid, int n, int lane) {
    EqState e = *gs;
    for (int q = 0; q < n; q++) {
        const float s = eq_step(e, mid[q], out[q], false, 0.0f);
        __syncwarp();
        if (lane == 0) out[q] = s;
    }
    __syncwarp();
    if (lane == 0) *gs = e;
}
