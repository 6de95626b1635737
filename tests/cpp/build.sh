#!/bin/bash
# builds the C++ host-side programs that sit directly on the C ABI (no Python): tests/cpp/bin/*
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"; ROOT="$(cd "$HERE/../.." && pwd)"
mkdir -p "$HERE/bin"
CUDA=${CUDA_HOME:-/usr/local/cuda}
for prog in rx_loopback shim_loopback shim_host_helpers; do
  if [ ! -x "$HERE/bin/$prog" ] || [ "$HERE/$prog.cpp" -nt "$HERE/bin/$prog" ] || [ "$ROOT/m17_sdr_b200/libm17b200.so" -nt "$HERE/bin/$prog" ]; then
    g++ -O2 -std=c++17 -I"$ROOT/include" -I"$CUDA/include" "$HERE/$prog.cpp" -o "$HERE/bin/$prog" \
        -L"$ROOT/m17_sdr_b200" -lm17b200 -L"$CUDA/lib64" -lcudart -Wl,-rpath,'$ORIGIN/../../../m17_sdr_b200' -Wl,-rpath,"$CUDA/lib64"
  fi
done
