/*
 * TEST INFRASTRUCTURE - not product code.
 *
 * Calls the tunable-parameter hooks of the UNMODIFIED reference (objects compiled from /root/reference by
 * oracle/Makefile) and prints what they return, so that tests/test_hooks_vs_reference.py can compare the
 * product's hooks of the same names with them point by point:
 *   params_range_setup (offt-compute.c:2998), params_set_default (:3127), grid_value_floor/ceil (:3096, :3111),
 *   params_convert (offt-tuning.c:80), is_infeasible_point (offt-tuning.c:144).
 *
 * usage: ref_hooks Nx Ny Nz p is_oned is_W0 is_notest npoints seed [simplex_seed tuning_mode is_r2c]
 *        with the three optional arguments it also prints "simplex <i> <24 indices>" for the 25 vertices the reference's
 *        write_initial_simplex (offt-tuning.c:426-737) draws after srand(simplex_seed)
 * output: "range i n v0 v1 ...", "default v0..v23", then per random index vector
 *         "point <24 indices> -> <24 values> infeasible <ret> <p_i>" and per tunable "floorceil i raw f c fi ci"
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>
#include <mpi.h>
#include "offt.h"
#include "offt-internal.h"

void params_set_default(struct _offt_plan *po);
void params_convert(int is_backward, int *v, long *ahv, struct _offt_plan *po, int **v_list, int *v_list_size);
int is_infeasible_point(struct _offt_plan *po, int *v, int *p_i);
void write_initial_simplex(struct _offt_plan *po, int **v_list, int *v_list_size);

static uint64_t mix(uint64_t seed, uint64_t index) {
  uint64_t z = seed * 0x9E3779B97F4A7C15ULL + index * 0xD1B54A32D192ED03ULL + 0x632BE59BD9B4E019ULL;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ULL;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBULL;
  return z ^ (z >> 31);
}

int main(int argc, char **argv) {
  if (argc < 10) { fprintf(stderr, "usage: see header\n"); return 2; }
  struct _offt_plan *po = (struct _offt_plan *)calloc(1, sizeof(*po));
  po->Nx = atoi(argv[1]); po->Ny = atoi(argv[2]); po->Nz = atoi(argv[3]); po->p = atoi(argv[4]);
  po->is_oned = atoi(argv[5]); po->is_W0 = atoi(argv[6]); po->is_notest = atoi(argv[7]);
  const int npoints = atoi(argv[8]);
  const uint64_t seed = strtoull(argv[9], NULL, 10);
  po->params = (struct _offt_params *)calloc(1, sizeof(struct _offt_params));
  if (argc >= 13) { po->tuning_mode = atoi(argv[11]); po->is_r2c = atoi(argv[12]); }
  int *v_list[PARAM_COUNT];
  int v_list_size[PARAM_COUNT];
  int i, k;
  params_range_setup(po, v_list, v_list_size);
  for (i = 0; i < PARAM_COUNT; i++) {
    printf("range %d %d", i, v_list_size[i]);
    for (k = 0; k < v_list_size[i]; k++) printf(" %d", v_list[i][k]);
    printf("\n");
  }
  params_set_default(po);
  printf("default");
  for (i = 0; i < PARAM_COUNT; i++) printf(" %d", po->params->v[i]);
  printf("\n");
  for (k = 0; k < npoints; k++) {
    long ahv[PARAM_COUNT], back[PARAM_COUNT];
    int v[PARAM_COUNT], bad = -1;
    for (i = 0; i < PARAM_COUNT; i++) ahv[i] = (long)(mix(seed, (uint64_t)k * PARAM_COUNT + i) % (uint64_t)v_list_size[i]);
    params_convert(1, v, ahv, po, v_list, v_list_size);
    const int inf = is_infeasible_point(po, v, &bad);
    printf("point");
    for (i = 0; i < PARAM_COUNT; i++) printf(" %ld", ahv[i]);
    printf(" ->");
    for (i = 0; i < PARAM_COUNT; i++) printf(" %d", v[i]);
    printf(" infeasible %d %d\n", inf, bad);
    (void)back;
  }
  for (i = 0; i < PARAM_COUNT; i++) {
    const int probes[6] = {0, 1, 3, 5, 100, 1000000};
    for (k = 0; k < 6; k++)
      printf("floorceil %d %d %d %d %d %d\n", i, probes[k], grid_value_floor(0, v_list, v_list_size, i, probes[k]),
             grid_value_ceil(0, v_list, v_list_size, i, probes[k]), grid_value_floor(1, v_list, v_list_size, i, probes[k]),
             grid_value_ceil(1, v_list, v_list_size, i, probes[k]));
  }
  if (argc >= 13) {
    char line[4096];
    snprintf(po->user_vertex_file, sizeof(po->user_vertex_file), "/tmp/ref_hooks_uv_%d", (int)getpid());
    srand((unsigned)atoi(argv[10]));
    write_initial_simplex(po, v_list, v_list_size);
    FILE *f = fopen(po->user_vertex_file, "r");
    for (i = 0; f && fgets(line, sizeof(line), f); i++) printf("simplex %d %s", i, line);
    if (f) fclose(f);
    remove(po->user_vertex_file);
  }
  /* the two printers of the public API (offt.h:243-244), on fixed inputs */
  {
    double t[GES];
    for (i = 0; i < GES; i++) t[i] = 0.001 * (i + 1) + 0.0000049 * i;
    printf("print_params: ");
    print_params(po->params->v);
    printf("offt_print_time: ");
    offt_print_time(t);
  }
  return 0;
}
