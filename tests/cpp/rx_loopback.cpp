// rx_loopback.cpp -- host-side C++ driver above the C ABI (no Python, no torch): synthesises stream-mode channels
// with the library's TX path, runs the batched RX chain, checks that every delivered payload equals what was
// sent, and prints one JSON line.  Mirrors what the reference's only test (m17_test.cpp:42-52, a baseband
// loopback) was meant to do, for N channels at once.  Also the binary used for ncu captures (profiles/).
//   rx_loopback <nchan> <nblocks> <reps> [noise_sigma_lsb] [f0_hz]
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <vector>
#include "m17b200.h"

#define CK(x) do { int rc__ = (x); if (rc__) { fprintf(stderr, "%s -> %d (%s) %s\n", #x, rc__, m17b_error_string(rc__), m17b_last_cuda_error()); return 1; } } while (0)
#define CU(x) do { cudaError_t e__ = (x); if (e__ != cudaSuccess) { fprintf(stderr, "%s -> %s\n", #x, cudaGetErrorString(e__)); return 1; } } while (0)

int main(int argc, char **argv) {
    const int64_t C = argc > 1 ? atoll(argv[1]) : 64, T = argc > 2 ? atoll(argv[2]) : 20;
    const int reps = argc > 3 ? atoi(argv[3]) : 2;
    const float sigma = argc > 4 ? (float)atof(argv[4]) : 300.0f, f0hz = argc > 5 ? (float)atof(argv[5]) : 500.0f;
    if (T < 8) { fprintf(stderr, "need at least 8 blocks\n"); return 2; }
    const int64_t F = T - 6, NS = T * 192;
    m17b_ctx *ctx; m17b_tx *tx; m17b_rx *rx;
    CK(m17b_ctx_create(0, &ctx));
    CK(m17b_tx_create(ctx, C, 10, &tx));
    CK(m17b_rx_create(ctx, C, T, &rx));

    // ---- LSF (dst broadcast, per-channel src, TYPE 0x0005) with CRC computed by the device primitive
    std::vector<uint8_t> lsf(C * 30, 0), payload(C * F * 16);
    srand(12345);
    for (int64_t c = 0; c < C; c++) {
        for (int i = 0; i < 6; i++) lsf[c * 30 + i] = 0xFF;
        uint64_t src = 0x25EA29F + (uint64_t)c;
        for (int i = 0; i < 6; i++) lsf[c * 30 + 6 + i] = (uint8_t)(src >> (40 - 8 * i));
        lsf[c * 30 + 13] = 0x05;
    }
    for (auto &b : payload) b = (uint8_t)(rand() >> 7);
    uint8_t *d_lsf, *d_payload, *d_dib, *d_script; uint16_t *d_crc; int16_t *d_iq; float *d_sigma, *d_f0;
    CU(cudaMalloc(&d_lsf, C * 30)); CU(cudaMalloc(&d_crc, C * 2)); CU(cudaMalloc(&d_payload, C * F * 16));
    CU(cudaMalloc(&d_dib, C * (F + 1) * 192)); CU(cudaMalloc(&d_script, C * NS)); CU(cudaMalloc(&d_iq, C * NS * 10 * 4));
    CU(cudaMalloc(&d_sigma, C * 4)); CU(cudaMalloc(&d_f0, C * 4));
    CU(cudaMemcpy(d_lsf, lsf.data(), C * 30, cudaMemcpyHostToDevice));
    CK(m17b_crc_array_encode(ctx, d_lsf, 30, 28, C, d_crc, nullptr));
    std::vector<uint16_t> crc(C);
    CU(cudaMemcpy(crc.data(), d_crc, C * 2, cudaMemcpyDeviceToHost));
    for (int64_t c = 0; c < C; c++) { lsf[c * 30 + 28] = (uint8_t)(crc[c] >> 8); lsf[c * 30 + 29] = (uint8_t)crc[c]; }
    CU(cudaMemcpy(d_lsf, lsf.data(), C * 30, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_payload, payload.data(), C * F * 16, cudaMemcpyHostToDevice));

    // ---- frames -> symbol script: carrier, preamble x2, LSF, F stream frames, EOT, carrier  (m17_tx_rx.cpp:95-115)
    CK(m17b_tx_set_lsf(tx, d_lsf, nullptr));
    CK(m17b_fmt_link_setup_frame(ctx, d_lsf, C, d_dib, nullptr));                       // [C][192]
    CK(m17b_fmt_stream_frames(tx, d_payload, F, d_dib + C * 192, nullptr));             // [C][F][192]
    uint8_t pre[192], eot[192], car[192];
    m17b_fmt_preamble(pre); m17b_fmt_eot(eot); memset(car, 4, 192);
    for (int64_t c = 0; c < C; c++) {
        uint8_t *s = d_script + c * NS;
        CU(cudaMemcpy(s, car, 192, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s + 192, pre, 192, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s + 384, pre, 192, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s + 576, d_dib + c * 192, 192, cudaMemcpyDeviceToDevice));
        CU(cudaMemcpy(s + 768, d_dib + C * 192 + c * F * 192, F * 192, cudaMemcpyDeviceToDevice));
        CU(cudaMemcpy(s + 768 + F * 192, eot, 192, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(s + 960 + F * 192, car, 192, cudaMemcpyHostToDevice));
    }
    CK(m17b_mod_dibits(tx, d_script, NS, d_iq, nullptr, nullptr));
    std::vector<float> sg(C, sigma), f0(C);
    for (int64_t c = 0; c < C; c++) { f0[c] = (c & 1 ? -f0hz : f0hz) * (float)(1 + c % 3) / 3.0f / 48000.0f; if (c % 4 == 0) sg[c] = 0.0f; }
    CU(cudaMemcpy(d_sigma, sg.data(), C * 4, cudaMemcpyHostToDevice));
    CU(cudaMemcpy(d_f0, f0.data(), C * 4, cudaMemcpyHostToDevice));
    CK(m17b_synth_channel(ctx, d_iq, C, NS * 10, d_sigma, d_f0, 99, nullptr));
    CU(cudaDeviceSynchronize());

    // ---- RX chain, timed with CUDA events on the launching (default) stream
    cudaEvent_t e0, e1; CU(cudaEventCreate(&e0)); CU(cudaEventCreate(&e1));
    float ms = 0;
    for (int r = 0; r < reps + 1; r++) {
        CK(m17b_rx_reset(rx, nullptr));
        if (r == 1) CU(cudaEventRecord(e0));
        CK(m17b_dsp_rx(rx, d_iq, T, nullptr));
    }
    CU(cudaEventRecord(e1)); CU(cudaEventSynchronize(e1));
    if (reps > 0) { CU(cudaEventElapsedTime(&ms, e0, e1)); ms /= reps; }

    // ---- check: every delivered stream payload equals the payload sent with that frame number
    m17b_rx_view v; CK(m17b_rx_get_view(rx, &v));
    std::vector<m17b_frame_rec> fr(C * v.frame_cap); std::vector<int32_t> nf(C);
    CU(cudaMemcpy(fr.data(), v.d_frames, fr.size() * sizeof(m17b_frame_rec), cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(nf.data(), v.d_nframes, C * 4, cudaMemcpyDeviceToHost));
    long delivered = 0, exact = 0, frames = 0;
    for (int64_t c = 0; c < C; c++)
        for (int k = 0; k < nf[c]; k++) {
            const m17b_frame_rec &r = fr[c * v.frame_cap + k];
            frames++;
            if (r.type != M17B_T_STREAM || !(r.flags & M17B_F_DELIVERED)) continue;
            delivered++;
            int fn = (r.data[0] << 8) | r.data[1];
            if (fn < F && !memcmp(r.data + 2, &payload[(c * F + fn) * 16], 16)) exact++;
        }
    printf("{\"channels\": %ld, \"blocks\": %ld, \"ms_per_pass\": %.4f, \"frames_per_s\": %.1f, \"records\": %ld, \"delivered\": %ld, \"payload_exact\": %ld, \"launches\": %d}\n",
           (long)C, (long)T, ms, ms > 0 ? C * T / (ms * 1e-3) : 0.0, frames, delivered, exact, m17b_rx_last_launches(rx));
    m17b_rx_destroy(rx); m17b_tx_destroy(tx); m17b_ctx_destroy(ctx);
    return (delivered > 0 && exact >= delivered * 0.97) ? 0 : 3;
}
