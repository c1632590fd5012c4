// Per-length kernel configurations: points per thread, radix schedule, shared-memory
// pads (chosen with tools/smem_model.py so that every exchange is conflict-free in
// row launches), launch bounds.  X(N, E, R0, R1, R2, R3, pads for complex128, pads for complex64, MAXT, MINB)
#pragma once
// launch bounds of the two lengths that carry BASELINE's configs can be overridden at build time
// (-DOFFTB_MAXT_512=256 -DOFFTB_MINB_512=4): tools/variants.sh builds such variant libraries for A/B runs on the GPU
#ifndef OFFTB_MAXT_512
#define OFFTB_MAXT_512 512
#endif
#ifndef OFFTB_MINB_512
#define OFFTB_MINB_512 1
#endif
#ifndef OFFTB_MAXT_1024
#define OFFTB_MAXT_1024 512
#endif
#ifndef OFFTB_MINB_1024
#define OFFTB_MINB_1024 1
#endif
// the radix schedule of the same two lengths, E, R0..R3, pads (complex128), pads (complex64)
#ifndef OFFTB_CFG_512
#define OFFTB_CFG_512 8, 8, 8, 8, 1, 0, 1, 0, 8, 2, 0
#endif
#ifndef OFFTB_CFG_1024
#define OFFTB_CFG_1024 16, 16, 16, 4, 1, 4, 2, 0, 4, 4, 0
#endif
#define OFFTB_APPLY(X, ...) X(__VA_ARGS__)
#define OFFTB_FFT_CONFIGS(X)                               \
  X(2, 2, 2, 1, 1, 1, 0, 0, 0, 0, 0, 0, 512, 1)            \
  X(4, 4, 4, 1, 1, 1, 0, 0, 0, 0, 0, 0, 512, 1)            \
  X(8, 8, 8, 1, 1, 1, 0, 0, 0, 0, 0, 0, 512, 1)            \
  X(16, 16, 16, 1, 1, 1, 0, 0, 0, 0, 0, 0, 512, 1)         \
  X(32, 8, 8, 4, 1, 1, 1, 0, 0, 1, 0, 0, 512, 1)           \
  X(64, 8, 8, 8, 1, 1, 1, 0, 0, 1, 0, 0, 512, 1)           \
  X(128, 16, 16, 8, 1, 1, 1, 0, 0, 1, 0, 0, 512, 1)        \
  X(256, 16, 16, 16, 1, 1, 1, 0, 0, 1, 0, 0, 512, 1)       \
  OFFTB_APPLY(X, 512, OFFTB_CFG_512, OFFTB_MAXT_512, OFFTB_MINB_512)          \
  OFFTB_APPLY(X, 1024, OFFTB_CFG_1024, OFFTB_MAXT_1024, OFFTB_MINB_1024)      \
  X(2048, 16, 16, 16, 8, 1, 0, 1, 0, 8, 2, 0, 512, 1)      \
  X(4096, 16, 16, 16, 16, 1, 0, 1, 0, 0, 1, 0, 512, 1)     \
  X(8192, 16, 16, 16, 16, 2, 0, 2, 4, 0, 2, 8, 512, 1)
