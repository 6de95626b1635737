/*
 * m17_records.h -- result-record layout shared by the three checkers/producers
 * (oracle/_ref harness, oracle/ C restatement, and -- by identical layout, declared
 * again in include/m17b200.h -- the CUDA product).  TEST INFRASTRUCTURE ONLY.
 *
 * The reference (m17gismo) reports RX results through synchronous up-calls
 * (m17_rx_frame.cpp:139,149,169 m17_aos/m17_los; m17_rx_parse.cpp:20-32,128,145,153-157).
 * A batched implementation returns one fixed-size record per completed frame instead.
 */
#ifndef M17_RECORDS_H
#define M17_RECORDS_H
#include <stdint.h>

#define M17R_F_SYNC_OK    0x01 /* m17_locked_sync_check() passed            (m17_rx_frame.cpp:144) */
#define M17R_F_PARSED     0x02 /* m17_rx_parse() was called for the frame   (m17_rx_frame.cpp:145,153) */
#define M17R_F_LOS        0x04 /* framer dropped lock on this frame         (m17_rx_frame.cpp:138-150) */
#define M17R_F_DELIVERED  0x08 /* stream payload passed upward              (m17_rx_parse.cpp:148-158) */
#define M17R_F_LSF_EVENT  0x10 /* parse_lsf() ran during this frame         (m17_rx_parse.cpp:82,99) */
#define M17R_F_PKT_EOF    0x20 /* packet frame carried the EOF bit          (m17_rx_parse.cpp:173) */

typedef struct {
    int32_t  sym_off;      /* index of m_f_sym[0] in the channel's emitted symbol stream          */
    uint8_t  type;         /* m17_sync_check() winner 0..5 (m17_rx_frame.cpp:5-12,66-75)          */
    uint8_t  flags;        /* M17R_F_*                                                            */
    uint8_t  golay_err;    /* stream frames: sum of 4 Golay error counts (m17_rx_parse.cpp:124-127)*/
    uint8_t  nbytes;       /* decoded bytes in data[]: 30 LSF / 18 stream / 26 packet / 0 other   */
    uint8_t  lich[6];      /* stream frames: Golay-corrected LICH chunk (m17_rx_parse.cpp:133)    */
    uint8_t  data[30];     /* Viterbi output packed MSB first from bits[1..] (m17_rx_parse.cpp:97,142,171) */
    uint16_t crc;          /* m17_crc_array_encode(data, nbytes) -- our verdict, see SURVEY D3    */
    uint8_t  votes;        /* sync-word sign mismatches (m17_rx_frame.cpp:77-80)                  */
    uint8_t  frame_errors; /* m_frame_errors after this frame (m17_rx_frame.cpp:146-148)          */
    float    variance;     /* find_variance() of the 8 sync symbols (m17_rx_frame.cpp:22-43)      */
    float    cor;          /* demap normaliser 8/sum|sync| (m17_dsp.cpp:88); 0 when not demapped  */
    uint8_t  rsvd[8];
} m17_frame_rec;           /* 64 bytes */

#define M17R_EV_AOS 1
#define M17R_EV_LOS 2
typedef struct {
    int32_t sym_idx;       /* stream index of the symbol that caused the event */
    int32_t kind;          /* M17R_EV_*                                        */
} m17_event_rec;

#endif
