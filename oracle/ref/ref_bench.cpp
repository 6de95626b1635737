/*
 * ref_bench.cpp -- CPU-baseline timing binary: the UNMODIFIED reference RX chain (m17_dsp_rx,
 * m17_dsp.cpp:461-476), no interposers, non-PIC -O3 objects exactly as the reference makefile
 * builds them (makefile:6), run multi-instance (one process per worker; the reference keeps its
 * state in file statics so a process is a channel).  TEST/BENCH INFRASTRUCTURE ONLY.
 *
 *   m17ref_bench <iq.bin> <C> <T> <nproc> [reps]
 *   m17ref_bench --tx <frames per worker> <nproc>
 * --tx: the reference TX chain, m17_send_stream_frame = m17_fmt_add_stream_frame + m17_mod_dibits (m17_tx_routines.cpp:306-310,
 *       m17_modulate.cpp:79-88) at oversample 10, nproc workers each sending <frames> stream frames after one link setup frame;
 *       radio_transmit_samples discards the IQ.  The clock runs around the send loop only.
 * iq.bin = raw int16 [C][T*1920][2].  nproc workers; each decodes its channels one after another, every channel in a
 * freshly forked child (pristine statics).  The clock runs around the m17_dsp_rx loop only.  Prints one JSON line:
 * frames/s = all frames / slowest worker's summed loop time; delivered = stream payloads passed up in one pass.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <time.h>
#include <unistd.h>
#include <sys/mman.h>
#include <sys/wait.h>
#include <sys/stat.h>
#include <fcntl.h>
#include "m17defines.h"

static long g_delivered = 0, g_aos = 0, g_los = 0;
void gui_aos(void) { g_aos++; }
void gui_los(void) { g_los++; }
void gui_update(void) {}
void gui_save_dest_address(uint48_t) {}
void gui_save_src_address(uint48_t) {}
bool m17_net_new_rx_data(uint16_t, uint8_t *, uint16_t, uint8_t *) { g_delivered++; return true; }
void m17_txrx_spkr_audio(uint8_t *) {}
void radio_afc(float) {}
float radio_get_afc_delta(void) { return 0; }
bool radio_get_afc_status(void) { return false; }
int  radio_get_oversample(void) { return 10; }
int  radio_transmit_samples(scmplx *, uint32_t n) { return (int)n; }
int  udp_send(uint8_t *, int len) { return len; }

static double now_s(void) { struct timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

void m17_send_link_setup_frame(uint48_t dest, uint48_t src, M17Type type, uint8_t *meta);
void m17_send_stream_frame(uint8_t *payload);

static int tx_main(long frames, int nproc) {
    struct Res { double secs; long frames; };
    Res *res = (Res *)mmap(0, sizeof(Res) * nproc, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    m17_prbs9_init(); m17_crc_init(); m17_init_conv(); m17_init_de_correlate(); m17_dsp_init(); m17_fmt_init();
    m17_golay_init(); m17_rx_sync_init(); m17_mod_init();
    for (int w = 0; w < nproc; w++) {
        if (fork() == 0) {
            uint8_t meta[14] = {0}, payload[16];
            M17Type type; memset(&type, 0, sizeof(type));
            m17_send_link_setup_frame(0xFFFFFFFFFFFFull, 0x123456789Aull + w, type, meta);
            unsigned x = 12345u + w;
            for (int i = 0; i < 50; i++) { for (int b = 0; b < 16; b++) { x = x * 1664525u + 1013904223u; payload[b] = x >> 24; } m17_send_stream_frame(payload); }
            double t0 = now_s();
            for (long f = 0; f < frames; f++) {
                for (int b = 0; b < 16; b++) { x = x * 1664525u + 1013904223u; payload[b] = x >> 24; }
                m17_send_stream_frame(payload);
            }
            res[w].secs = now_s() - t0; res[w].frames = frames;
            _exit(0);
        }
    }
    for (int w = 0; w < nproc; w++) { int st; wait(&st); }
    double maxs = 0; long tot = 0;
    for (int w = 0; w < nproc; w++) { if (res[w].secs > maxs) maxs = res[w].secs; tot += res[w].frames; }
    printf("{\"mode\": \"tx\", \"frames\": %ld, \"secs_max_worker\": %.6f, \"frames_per_s\": %.1f, \"nproc\": %d}\n", tot, maxs, tot / maxs, nproc);
    return 0;
}

int main(int argc, char **argv) {
    if (argc >= 4 && !strcmp(argv[1], "--tx")) return tx_main(atol(argv[2]), atoi(argv[3]));
    if (argc < 5) { fprintf(stderr, "usage: %s iq.bin C T nproc [reps]\n", argv[0]); return 2; }
    long C = atol(argv[2]), T = atol(argv[3]); int nproc = atoi(argv[4]); int reps = argc > 5 ? atoi(argv[5]) : 1;
    int fd = open(argv[1], O_RDONLY);
    if (fd < 0) { perror("open"); return 2; }
    size_t bytes = (size_t)C * T * 3840 * sizeof(int16_t);
    const int16_t *iq = (const int16_t *)mmap(0, bytes, PROT_READ, MAP_PRIVATE | MAP_POPULATE, fd, 0);
    if (iq == MAP_FAILED) { perror("mmap"); return 2; }
    struct Res { double secs; long frames, delivered, aos, los; };
    /* one slot per channel (filled by the per-channel child) and one per worker */
    Res *chan = (Res *)mmap(0, sizeof(Res) * C, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    Res *res = (Res *)mmap(0, sizeof(Res) * nproc, PROT_READ | PROT_WRITE, MAP_SHARED | MAP_ANONYMOUS, -1, 0);
    /* main.cpp:108-126 init order */
    m17_prbs9_init(); m17_crc_init(); m17_init_conv(); m17_init_de_correlate(); m17_dsp_init(); m17_fmt_init();
    m17_golay_init(); m17_rx_sync_init(); m17_mod_init(); m17_db_set_chan_type(DRTODN);
    double w0 = now_s();
    for (int w = 0; w < nproc; w++) {
        if (fork() == 0) {
            /* touch my slice once (page-in) before timing */
            volatile long sink = 0;
            for (long c = w; c < C; c += nproc) for (long i = 0; i < T * 3840; i += 2048) sink += iq[c * T * 3840 + i];
            double secs = 0; long frames = 0, deliv = 0;
            for (int r = 0; r < reps; r++)
                for (long c = w; c < C; c += nproc) {
                    /* the reference keeps its state in file statics: a pristine process per channel */
                    pid_t p = fork();
                    if (p == 0) {
                        int16_t blk[3840];
                        const int16_t *src = iq + c * T * 3840;
                        double t0 = now_s();
                        for (long t = 0; t < T; t++) { memcpy(blk, src + t * 3840, sizeof(blk)); m17_dsp_rx((scmplx *)blk, 1920); }
                        chan[c].secs = now_s() - t0; chan[c].frames = T; chan[c].delivered = g_delivered; chan[c].aos = g_aos; chan[c].los = g_los;
                        _exit(0);
                    }
                    int st; waitpid(p, &st, 0);
                    secs += chan[c].secs; frames += chan[c].frames; if (r == 0) deliv += chan[c].delivered;
                }
            res[w].secs = secs; res[w].frames = frames; res[w].delivered = deliv;
            _exit(0);
        }
    }
    for (int w = 0; w < nproc; w++) { int st; wait(&st); }
    double wall = now_s() - w0, maxs = 0; long frames = 0, deliv = 0;
    for (int w = 0; w < nproc; w++) { if (res[w].secs > maxs) maxs = res[w].secs; frames += res[w].frames; deliv += res[w].delivered; }
    printf("{\"frames\": %ld, \"secs_max_worker\": %.6f, \"wall\": %.6f, \"frames_per_s\": %.1f, \"nproc\": %d, \"delivered\": %ld}\n",
           frames, maxs, wall, frames / maxs, nproc, deliv);
    return 0;
}
