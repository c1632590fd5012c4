// Per-length kernel configurations: points per thread, radix schedule, shared-memory
// pads (chosen with tools/smem_model.py so that every exchange is conflict-free in
// row launches), launch bounds.  X(N, E, R0, R1, R2, R3, pads for complex128, pads for complex64, MAXT, MINB)
#pragma once
#define OFFTB_FFT_CONFIGS(X)                               \
  X(2, 2, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 512, 1)            \
  X(4, 4, 4, 1, 1, 1, 0, 0, 0, 0, 0, 0, 512, 1)            \
  X(8, 8, 8, 1, 1, 1, 0, 0, 0, 0, 0, 0, 512, 1)            \
  X(16, 16, 16, 1, 1, 1, 0, 0, 0, 0, 0, 0, 512, 1)         \
  X(32, 8, 8, 4, 1, 1, 1, 0, 0, 1, 0, 0, 512, 1)           \
  X(64, 8, 8, 8, 1, 1, 1, 0, 0, 1, 0, 0, 512, 1)           \
  X(128, 16, 16, 8, 1, 1, 1, 0, 0, 1, 0, 0, 512, 1)        \
  X(256, 16, 16, 16, 1, 1, 1, 0, 0, 1, 0, 0, 512, 1)       \
  X(512, 8, 8, 8, 8, 1, 0, 1, 0, 8, 2, 0, 512, 1)          \
  X(1024, 16, 16, 16, 4, 1, 4, 2, 0, 4, 4, 0, 512, 1)      \
  X(2048, 16, 16, 16, 8, 1, 0, 1, 0, 8, 2, 0, 512, 1)      \
  X(4096, 16, 16, 16, 16, 1, 0, 1, 0, 0, 1, 0, 512, 1)     \
  X(8192, 16, 16, 16, 16, 2, 0, 2, 4, 0, 2, 8, 512, 1)
