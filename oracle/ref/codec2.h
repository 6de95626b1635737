/* Stub for the one third-party header m17defines.h includes (m17defines.h:5).  CODEC2 appears only in a
   prototype that is never defined (m17defines.h:358); codec2 itself is off the hot path. */
#ifndef ORACLE_STUB_CODEC2_H
#define ORACLE_STUB_CODEC2_H
struct CODEC2;
#endif
