// shim_host_helpers.cpp -- the host-only helpers of include/m17gismo_b200.hpp (m17_encode_call / m17_decode_call /
// m17_pack_type / m17_upack_type, m17_bit_utils.cpp:191-254) as a filter: runs without a GPU (CPU test tier).
//   stdin : one request per line:  "E <9-character call>"  |  "D <hex word>"  |  "T <hex type word>"
//   stdout: the answer per line (hex word | 9-character call in brackets | the six M17Type fields and the re-packed word)
#define M17GISMO_B200_IMPLEMENTATION
#include "m17gismo_b200.hpp"
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
int main() {
    char line[256];
    while (fgets(line, sizeof(line), stdin)) {
        size_t n = strlen(line);
        while (n && (line[n - 1] == '\n' || line[n - 1] == '\r')) line[--n] = 0;
        if (line[0] == 'E' && n >= 2) {
            char call[10];
            memset(call, ' ', 9); call[9] = 0;
            memcpy(call, line + 2, n - 2 > 9 ? 9 : n - 2);
            printf("%llx\n", (unsigned long long)m17_encode_call(call));
        } else if (line[0] == 'D') {
            char call[16];
            printf("[%s]\n", m17_decode_call(strtoull(line + 2, 0, 16), call));
        } else if (line[0] == 'T') {
            M17Type t = m17_upack_type((uint16_t)strtoul(line + 2, 0, 16));
            printf("%u %u %u %u %u %u %x\n", t.p_s, t.dt, t.et, t.est, t.can, t.reserved, m17_pack_type(t));
        }
    }
    return 0;
}
