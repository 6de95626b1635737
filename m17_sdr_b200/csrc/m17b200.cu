// m17b200.cu -- the single translation unit of libm17b200.so (sm_100a only).
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -fmad=false -Xcompiler -fPIC -shared
//        -I include -o m17_sdr_b200/libm17b200.so m17_sdr_b200/csrc/m17b200.cu
// Parts (each includes the previous one): common -> tables -> fec -> decode -> frontend -> afc -> sync -> sync_cta -> sync_g -> framer -> dec -> chan -> rx -> tx (mod) -> app -> net
#include "net.cuh"
