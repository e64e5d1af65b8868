#!/bin/sh
# builds tools/trace_fwd.bin (in-kernel timeline of the forward chain; needs -DM2_TRACE, not part of libm2b200.so)
cd "$(dirname "$0")/.." && /usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 --expt-relaxed-constexpr \
  -DM2_TRACE -Im2_mixer_b200/csrc tools/trace_fwd.cu m2_mixer_b200/csrc/chain_ts.cu m2_mixer_b200/csrc/profile.cu -o tools/trace_fwd.bin
