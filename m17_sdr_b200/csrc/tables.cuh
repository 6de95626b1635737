// tables.cuh -- host-side table / filter design and the per-GPU context.
// Replaces the reference's init chain (main.cpp:108-126): m17_prbs9_init, m17_crc_init, m17_init_conv,
// m17_init_de_correlate, m17_golay_init, m17_rx_sync_init, m17_mod_init.
#pragma once
#include "common.cuh"

__constant__ GatherMaps c_maps;
__constant__ TxMaps     c_tx;
__constant__ PunctSteps c_punct;

// M17 randomiser sequence (protocol constant, m17_correlate.cpp:3-7)
static const uint8_t kRandSeq[46] = {
    0xD6, 0xB5, 0xE2, 0x30, 0x82, 0xFF, 0x84, 0x62, 0xBA, 0x4E, 0x96, 0x90, 0xD8, 0x98, 0xDD, 0x5D,
    0x0C, 0xC8, 0x52, 0x43, 0x91, 0x1D, 0xF8, 0x6E, 0x68, 0x2F, 0x35, 0xDA, 0x14, 0xEA, 0xCD, 0x76,
    0x19, 0x8D, 0xD5, 0x80, 0xD1, 0x33, 0x87, 0x13, 0x57, 0x18, 0x2D, 0x29, 0x78, 0xC3};
// Golay(24,12) generator parity rows (protocol constant, m17_golay.cpp:11)
static const uint16_t kGolayRows[12] = {0xC75, 0x63B, 0xF68, 0x7B4, 0x3DA, 0xD99, 0x6CD, 0x367, 0xDC6, 0xA97, 0x93E, 0x8EB};

static thread_local char g_cuda_err[256] = "";
void m17b_set_cuda_error(cudaError_t e, const char *file, int line) {
    snprintf(g_cuda_err, sizeof(g_cuda_err), "%s (%s:%d)", cudaGetErrorString(e), file, line);
}

// ---------------------------------------------------------------- host helpers
static inline bool punct_keeps(int pattern, int i) {
    // P1: period 61, every 4th position starting at 2 dropped; P2: period 12, last dropped; P3: period 8, last dropped
    if (pattern == 1) return ((i % 61) & 3) != 2;
    if (pattern == 2) return (i % 12) != 11;
    return (i % 8) != 7;
}
static inline int qpp_perm(int i) { return (45 * i + 92 * i * i) % 368; }
static inline int rand_bit(int i) { return (kRandSeq[i >> 3] >> (7 - (i & 7))) & 1; }

extern "C" int m17b_build_rrc_filter(float *taps, float rolloff, int ntaps, int sps) {
    // m17_dsp.cpp:295-315.  All arithmetic in double with the host libm, one rounding to float per tap;
    // the first tap time uses C integer division (half-tap offset for even ntaps).
    if (!taps || ntaps <= 0 || sps <= 0) return M17B_E_ARG;
    const double beta = rolloff + 0.0001;
    const double Ts = sps;
    double t = -(ntaps - 1) / 2;
    for (int n = 0; n < ntaps; n++, t = t + 1.0) {
        const double a = 2.0 * beta / (M_PI * sqrt(Ts));
        const double b = cos((1.0 + beta) * M_PI * t / Ts);
        const double c = (t == 0) ? (1.0 - beta) * M_PI / (4 * beta)
                                  : sin((1.0 - beta) * M_PI * t / Ts) / (4.0 * beta * t / Ts);
        const double d = (1.0 - (4.0 * beta * t / Ts) * (4.0 * beta * t / Ts));
        taps[n] = (float)(a * (b + c) / d);
    }
    return M17B_OK;
}
extern "C" int m17b_set_filter_gain(float *taps, float gain, int stride, int ntaps) {
    // m17_dsp.cpp:420-429: float running sum, float quotient, float scaling
    if (!taps) return M17B_E_ARG;
    float acc = 0;
    for (int n = 0; n < ntaps; n++) acc += taps[n * stride];
    gain = gain / acc;
    for (int n = 0; n < ntaps; n++) taps[n * stride] = taps[n * stride] * gain;
    return M17B_OK;
}

static void build_sync_banks(float *mf_out, float *md_out) {
    // m17_rx_sync.cpp:101-123: 1240-tap mother RRC at 80 samples/symbol, cyclic central difference of the
    // UN-normalised mother, polyphase partition bank[p][j] = mother[p + 40 j], each matched branch scaled to sum 1.
    const int N = M17B_NF * M17B_FN;
    float *mother = (float *)malloc(sizeof(float) * N), *deriv = (float *)malloc(sizeof(float) * N);
    m17b_build_rrc_filter(mother, 0.5f, N, M17B_NF * 2);
    for (int i = 0; i < N; i++) deriv[i] = mother[(i + 1) % N] - mother[(i + N - 1) % N];
    for (int p = 0; p < M17B_NF; p++)
        for (int j = 0; j < M17B_FN; j++) {
            mf_out[p * M17B_FN + j] = mother[p + j * M17B_NF];
            md_out[p * M17B_FN + j] = deriv[p + j * M17B_NF];
        }
    for (int p = 0; p < M17B_NF; p++) m17b_set_filter_gain(&mf_out[p * M17B_FN], 1.0f, 1, M17B_FN);
    free(mother);
    free(deriv);
}

static void build_gather_maps(GatherMaps *g, TxMaps *tx) {
    // so[j] (after de-interleave) = +-sb[pi(j)] because pi is an involution (SURVEY 4 KAT); sb[i] is the
    // MSB (i even) or LSB (i odd) soft bit of payload symbol i/2, i.e. frame symbol 8 + i/2.
    auto entry = [](int j) -> uint16_t {
        int i = qpp_perm(j);
        return (uint16_t)((8 + (i >> 1)) | ((i & 1) ? MAP_LSB : 0) | (rand_bit(i) ? MAP_NEG : 0));
    };
    int k = 0;
    for (int p = 0; p < 488; p++) g->p1[p] = punct_keeps(1, p) ? entry(k++) : (uint16_t)MAP_ERASE;
    k = 96;
    for (int p = 0; p < 296; p++) g->p2[p] = punct_keeps(2, p) ? entry(k++) : (uint16_t)MAP_ERASE;
    k = 0;
    for (int p = 0; p < 420; p++) g->p3[p] = punct_keeps(3, p) ? entry(k++) : (uint16_t)MAP_ERASE;
    for (int j = 0; j < 96; j++) g->lich[j] = entry(j);
    k = 0;
    for (int p = 0; p < 402; p++) g->bert[p] = (punct_keeps(2, p) && k < 368) ? entry(k++) : (uint16_t)MAP_ERASE;

    for (int i = 0; i < 368; i++) { tx->qpp[i] = (uint16_t)qpp_perm(i); tx->rnd[i] = (uint8_t)rand_bit(i); }
    for (int pat = 1; pat <= 3; pat++) {
        uint16_t *u = pat == 1 ? tx->unp1 : pat == 2 ? tx->unp2 : tx->unp3;
        int kept = 0;
        for (int p = 0; kept < 368; p++) if (punct_keeps(pat, p)) u[kept++] = (uint16_t)p;
    }
}

__global__ void k_selftest_nofma(const float *in, float *out) {
    // a*b+c must round twice (product, then sum).  With contraction enabled the result differs.
    out[0] = in[0] * in[1] + in[2];
}

template <class T> static int upload(T **dst, const T *src, size_t n) {
    CUDA_TRY(cudaMalloc((void **)dst, n * sizeof(T)));
    CUDA_TRY(cudaMemcpy(*dst, src, n * sizeof(T), cudaMemcpyHostToDevice));
    return M17B_OK;
}

extern "C" int m17b_ctx_create(int device, m17b_ctx **out) {
    if (!out) return M17B_E_ARG;
    *out = nullptr;
    int ndev = 0;
    CUDA_TRY(cudaGetDeviceCount(&ndev));
    if (device < 0 || device >= ndev) return M17B_E_ARG;
    CUDA_TRY(cudaSetDevice(device));
    m17b_ctx *ctx = (m17b_ctx *)calloc(1, sizeof(m17b_ctx));
    if (!ctx) return M17B_E_NOMEM;
    ctx->device = device;
    int rc;

    // CRC-16, polynomial 0x5935, MSB first (m17_crc.cpp:4-24)
    uint16_t crc[256];
    for (int b = 0; b < 256; b++) {
        uint16_t r = (uint16_t)(b << 8);
        for (int k = 0; k < 8; k++) r = (r & 0x8000) ? (uint16_t)((r << 1) ^ 0x5935) : (uint16_t)(r << 1);
        crc[b] = r;
    }
    // The CRC register update is linear over GF(2): step(crc, b) = Z(crc) ^ crc[b] with Z(crc) = (crc << 8) ^ crc[crc >> 8] (a zero
    // byte), so the CRC of 30 bytes is Z^30(0xFFFF) ^ XOR_i Z^(29-i)(crc[b_i]): one table look-up per byte position and an XOR
    // reduction across the lanes of a warp instead of a 30-step dependent chain (k_post).
    uint16_t *crcpos = (uint16_t *)malloc(2 * (30 * 256 + 1));
    {
        auto Z = [&](uint16_t v) { return (uint16_t)((uint16_t)(v << 8) ^ crc[v >> 8]); };
        for (int i = 0; i < 30; i++)
            for (int b = 0; b < 256; b++) {
                uint16_t v = crc[b];
                for (int k = 0; k < 29 - i; k++) v = Z(v);
                crcpos[i * 256 + b] = v;
            }
        uint16_t v = 0xFFFF;
        for (int k = 0; k < 30; k++) v = Z(v);
        crcpos[30 * 256] = v;
    }
    // Golay parity (m17_golay.cpp:31-40) and syndrome -> (weight, data error) tables.  The syndrome table is the
    // reference's brute-force scan: all 24-bit words of weight <= 4 in ascending order, last writer wins, after the
    // 0x400 pre-fill of entries 0..0xFFE (m17_golay.cpp:49-72) -- the order decides which 4-bit pattern
    // an uncorrectable syndrome maps to, so it is reproduced literally.
    uint16_t *genc = (uint16_t *)malloc(2 * 4096), *gerr = (uint16_t *)malloc(2 * 4096);
    for (int d = 0; d < 4096; d++) {
        uint16_t par = 0;
        for (int r = 0; r < 12; r++) if (d & (0x800 >> r)) par ^= kGolayRows[r];
        genc[d] = par;
    }
    for (int s = 0; s < 0xFFF; s++) gerr[s] = 0x400;
    gerr[0xFFF] = 0;
    for (uint32_t w = 0; w < (1u << 24); w++) {
        int wt = __builtin_popcount(w);
        if (wt < 5) gerr[(w & 0xFFF) ^ genc[w >> 12]] = (uint16_t)((wt << 12) | (w >> 12));
    }
    // PRBS9, x^9 + x^5 + 1, seed 1 (m17_prbs9.cpp:16-26)
    uint8_t prbs[511];
    {
        uint16_t sr = 1;
        for (int n = 0; n < 511; n++) { uint8_t b = ((sr >> 8) ^ (sr >> 4)) & 1; sr = ((sr << 1) | b) & 0x1FF; prbs[n] = b; }
    }
    build_sync_banks(ctx->h_mf, ctx->h_md);
    GatherMaps gm; TxMaps tm;
    build_gather_maps(&gm, &tm);
    PunctSteps ps;
    for (int pat = 1; pat <= 3; pat++)
        for (int t = 0; t < 244; t++) ps.keep[pat - 1][t] = (uint8_t)((punct_keeps(pat, 2 * t) ? 1 : 0) | (punct_keeps(pat, 2 * t + 1) ? 2 : 0));
    for (int pat = 1; pat <= 3; pat++) {
        int k = 0;
        for (int t = 0; t <= 16 * VP_CHUNK; t++) {
            if (t % VP_CHUNK == 0) ps.koff[pat - 1][t / VP_CHUNK] = (uint16_t)k;
            if (t < 244) k += (ps.keep[pat - 1][t] & 1) + ((ps.keep[pat - 1][t] >> 1) & 1);
        }
    }

    if ((rc = upload(&ctx->d_crc, crc, 256)) || (rc = upload(&ctx->d_crcpos, crcpos, 30 * 256 + 1)) || (rc = upload(&ctx->d_genc, genc, 4096)) || (rc = upload(&ctx->d_gerr, gerr, 4096)) ||
        (rc = upload(&ctx->d_mf, ctx->h_mf, M17B_NF * M17B_FN)) || (rc = upload(&ctx->d_md, ctx->h_md, M17B_NF * M17B_FN)) ||
        (rc = upload(&ctx->d_prbs, prbs, 511))) { free(genc); free(gerr); free(crcpos); free(ctx); return rc; }
    free(genc); free(gerr); free(crcpos);
    {
        uint16_t smap[STREAM_NIN + 96];
        int k = 0;
        // entry -> index into the frame's 368 soft values in natural order (2 * (symbol - 8) + lsb), bit 15 = negate
        auto compact = [](uint16_t e) { return (uint16_t)((2 * ((e & 0xFF) - 8) + ((e & MAP_LSB) ? 1 : 0)) | ((e & MAP_NEG) ? 0x8000 : 0)); };
        for (int p = 0; p < 296; p++) if (gm.p2[p] != MAP_ERASE) smap[k++] = compact(gm.p2[p]);
        if (k != STREAM_NIN) { free(ctx); return M17B_E_ARG; }
        for (int b = 0; b < 96; b++) smap[STREAM_NIN + b] = compact(gm.lich[b]);
        if ((rc = upload(&ctx->d_smap, smap, (size_t)STREAM_NIN + 96))) { free(ctx); return rc; }
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_maps, &gm, sizeof(gm)));
    CUDA_TRY(cudaMemcpyToSymbol(c_tx, &tm, sizeof(tm)));
    CUDA_TRY(cudaMemcpyToSymbol(c_punct, &ps, sizeof(ps)));

    // contraction self-test: (1+2^-12)*(1+2^-12) - 1 : separate rounding gives 2^-11, an FMA gives 2^-11 + 2^-24
    float h_in[3] = {1.0f + 1.0f / 4096, 1.0f + 1.0f / 4096, -1.0f}, h_out = 0, *d_t;
    CUDA_TRY(cudaMalloc((void **)&d_t, 4 * sizeof(float)));
    CUDA_TRY(cudaMemcpy(d_t, h_in, sizeof(h_in), cudaMemcpyHostToDevice));
    k_selftest_nofma<<<1, 1>>>(d_t, d_t + 3);
    CUDA_TRY(cudaMemcpy(&h_out, d_t + 3, sizeof(float), cudaMemcpyDeviceToHost));
    CUDA_TRY(cudaFree(d_t));
    if (h_out != 1.0f / 2048) {
        snprintf(g_cuda_err, sizeof(g_cuda_err), "libm17b200 was built with FMA contraction enabled; rebuild with -fmad=false");
        free(ctx);
        return M17B_E_CUDA;
    }
    *out = ctx;
    return M17B_OK;
}

extern "C" int m17b_ctx_destroy(m17b_ctx *ctx) {
    if (!ctx) return M17B_E_ARG;
    cudaFree(ctx->d_crc); cudaFree(ctx->d_crcpos); cudaFree(ctx->d_genc); cudaFree(ctx->d_gerr); cudaFree(ctx->d_mf); cudaFree(ctx->d_md); cudaFree(ctx->d_prbs); cudaFree(ctx->d_smap);
    free(ctx);
    return M17B_OK;
}
extern "C" int m17b_get_sync_taps(const m17b_ctx *ctx, float *mf, float *md) {
    if (!ctx || !mf || !md) return M17B_E_ARG;
    memcpy(mf, ctx->h_mf, sizeof(ctx->h_mf));
    memcpy(md, ctx->h_md, sizeof(ctx->h_md));
    return M17B_OK;
}
extern "C" int m17b_version(void) { return M17B_VERSION; }
extern "C" const char *m17b_last_cuda_error(void) { return g_cuda_err; }
extern "C" const char *m17b_error_string(int code) {
    switch (code) {
        case M17B_OK: return "ok";
        case M17B_E_ARG: return "bad argument";
        case M17B_E_CUDA: return "CUDA error";
        case M17B_E_NOMEM: return "out of memory";
        case M17B_E_UNSUPPORTED: return "unsupported";
        case M17B_E_CAPACITY: return "capacity exceeded";
        default: return "unknown";
    }
}
