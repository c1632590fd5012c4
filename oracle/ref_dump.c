/*
 * TEST INFRASTRUCTURE — not product code.
 *
 * Driver for the UNMODIFIED reference library (offt-compute.c / offt-tuning.c
 * compiled from /root/reference into oracle/_ref/, see Makefile): fills the
 * rank-local input box of a seeded global grid, runs the reference's
 * offt_3d_init / offt_3d_execute / offt_3d_fin (offt.h:235-241) and writes each
 * rank's descriptor and raw in-place output array to <prefix>.rank<r>.bin.
 * Used to pin oracle/offt_oracle.c and to generate tests/golden/.
 *
 * usage: OFFT_SHIM_NP=P ref_dump Nx Ny Nz seed prefix is_oned is_equalxy reps [idx=value ...] [r2c=1]
 *        (idx=value overrides custom_params->v[idx], e.g. 0=4 sets _P1_ to 4; r2c=1 runs the real-to-complex
 *        transform of the grid's real parts, input laid out as run-fft.c:53-55 does)
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <mpi.h>
#include "offt.h"

/* counter-based generator shared with oracle/oracle.py and the product's bench: splitmix64 */
static double grid_value(uint64_t seed, uint64_t index) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ULL + index * 0xD1B54A32D192ED03ULL + 0x632BE59BD9B4E019ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  z = z ^ (z >> 31);
  return (double)(z >> 11) * (2.0 / 9007199254740992.0) - 1.0;
}

int main(int argc, char **argv) {
  int p, rank, i;
  MPI_Init(&argc, &argv);
  MPI_Comm_size(MPI_COMM_WORLD, &p);
  MPI_Comm_rank(MPI_COMM_WORLD, &rank);
  if (argc < 9) { if (!rank) fprintf(stderr, "usage: see header\n"); MPI_Finalize(); return 2; }
  int Nx = atoi(argv[1]), Ny = atoi(argv[2]), Nz = atoi(argv[3]);
  uint64_t seed = strtoull(argv[4], NULL, 10);
  const char *prefix = argv[5];
  int is_oned = atoi(argv[6]), is_equalxy = atoi(argv[7]), reps = atoi(argv[8]);
  struct _offt_params *cp = (struct _offt_params *)malloc(sizeof(*cp));
  for (i = 0; i < PARAM_COUNT; i++) cp->v[i] = -1;
  int is_r2c = 0;
  for (i = 9; i < argc; i++) {
    int idx, val;
    if (sscanf(argv[i], "r2c=%d", &val) == 1) { is_r2c = val; continue; }
    if (sscanf(argv[i], "%d=%d", &idx, &val) == 2 && idx >= 0 && idx < PARAM_COUNT) cp->v[idx] = val;
  }
  int p1 = cp->v[_P1_];
  if (p1 < 0) { fprintf(stderr, "ref_dump: P1 (0=value) is required\n"); return 2; }
  int p2 = p / p1;
  /* allocation rule of run-fft.c:294-304, in 64-bit */
  long M1 = (Nx + p1 - 1) / p1, M2 = (Ny + p2 - 1) / p2, M3 = ((is_r2c ? Nz / 2 + 1 : Nz) + p2 - 1) / p2, M4 = (Ny + p1 - 1) / p1;
  long size = (M2 * p2 > M4 * p1) ? M1 * M2 * M3 * p2 : M1 * M3 * M4 * p1;
  double *out = (double *)calloc((size_t)size * 2, sizeof(double));
  struct _offt_plan *po = offt_3d_init(Nx, Ny, Nz, out, out, is_r2c, FFTW_ESTIMATE, is_oned, 0, is_equalxy,
                                        1, 0, 0, 0, 0, 0, cp);
  struct _offt_comm *c = po->comm;
  double tmin = 1e30;
  int r;
  for (r = 0; r < reps; r++) {
    memset(out, 0, (size_t)size * 2 * sizeof(double));
    int x, y, z;
    for (x = 0; x < c->isize[0]; x++)
      for (y = 0; y < c->isize[1]; y++)
        for (z = 0; z < c->isize[2]; z++) {
          uint64_t g = ((uint64_t)(x + c->istart[0]) * (uint64_t)Ny + (uint64_t)(y + c->istart[1])) * (uint64_t)Nz
                       + (uint64_t)(z + c->istart[2]);
          size_t a = (size_t)z * c->istride[2] + (size_t)y * c->istride[1] + (size_t)x * c->istride[0];
          if (is_r2c) {   /* run-fft.c:53-55: real z of row (x, y) at double index z + 2*(row start) */
            out[(size_t)z + 2 * ((size_t)y * c->istride[1] + (size_t)x * c->istride[0])] = grid_value(seed, 2 * g);
            continue;
          }
          out[2 * a] = grid_value(seed, 2 * g);
          out[2 * a + 1] = grid_value(seed, 2 * g + 1);
        }
    MPI_Barrier(MPI_COMM_WORLD);
    double t = -MPI_Wtime();
    offt_3d_execute(po, out, out, 0);
    MPI_Barrier(MPI_COMM_WORLD);
    t += MPI_Wtime();
    if (t < tmin) tmin = t;
    if (!rank) { printf("ref_dump t_rep %d %.6f\n", r, t); fflush(stdout); }
  }
  if (!rank) printf("ref_dump t_min %.6f\n", tmin);
  if (strcmp(prefix, "-") != 0) {
    char name[512];
    snprintf(name, sizeof(name), "%s.rank%d.bin", prefix, rank);
    FILE *f = fopen(name, "wb");
    if (!f) { perror("ref_dump: fopen"); return 3; }
    int64_t hdr[32];
    memset(hdr, 0, sizeof(hdr));
    hdr[0] = p; hdr[1] = rank; hdr[2] = Nx; hdr[3] = Ny; hdr[4] = Nz; hdr[5] = c->p1; hdr[6] = c->p2;
    for (i = 0; i < 3; i++) {
      hdr[7 + i] = c->istart[i]; hdr[10 + i] = c->isize[i]; hdr[13 + i] = c->istride[i];
      hdr[16 + i] = c->ostart[i]; hdr[19 + i] = c->osize[i]; hdr[22 + i] = c->ostride[i];
    }
    hdr[25] = size;
    fwrite(hdr, sizeof(int64_t), 32, f);
    int32_t pv[PARAM_COUNT];
    for (i = 0; i < PARAM_COUNT; i++) pv[i] = po->params->v[i];
    fwrite(pv, sizeof(int32_t), PARAM_COUNT, f);
    fwrite(out, sizeof(double), (size_t)size * 2, f);
    fclose(f);
  }
  offt_3d_fin(po);
  free(out); free(cp);
  MPI_Finalize();
  return 0;
}
