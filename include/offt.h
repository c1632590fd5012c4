/*
 * offt.h - C API of the B200-native distributed 3-D complex FFT.
 *
 * Source-compatible with the public header of rchyena/offt (reference offt.h:66-257):
 * the same five entry points with the same arity (offt.h:235-244, the NOTEST form),
 * the same three structs with the same field names in the same order
 * (_offt_params offt.h:69-100, _offt_comm offt.h:102-142 in its A2AV form - every
 * target of the reference Makefile passes -DA2AV -, _offt_plan offt.h:144-233), the same
 * parameter and timer index macros, and the min/max inlines the reference driver
 * uses (offt.h:251-257; run-fft.c:66,110,275-276).  run-fft.c compiles against this
 * file unmodified.  Written from scratch for this project; the implementation behind it
 * is CUDA for sm_100a (offt_b200/csrc), not MPI+FFTW.
 *
 * Differences a caller can observe:
 *   - `in`/`out` may be host OR device pointers (complex interleaved, in == out as
 *     in the reference, offt-compute.c:3866); host arrays are staged through HBM.
 *   - the fftw_plan / MPI_Comm typed fields are kept for layout and are NULL;
 *     the engine hangs off the trailing `b200` field.
 *   - extensions (inverse, single precision, device streams, rank emulation) live in
 *     offt_b200.h and never change the five signatures below.
 */
#ifndef OFFT_INCLUDE
#define OFFT_INCLUDE

#ifndef A2AV
#define A2AV /* exact-count bookkeeping fields are always present (offt.h:109-126) */
#endif
#ifndef STRIDE
#define STRIDE /* the _S_ switch is always compiled in (Makefile:27-29) */
#endif
#define NOTEST      /* 17-argument offt_3d_init (offt.h:27, 235-236) */
#define AH_TUNING
#define TUNING_REPS 1
#define SUBTILE_SIZE (8192)
#define BUFFER_SIZE_LIMIT (32 * 1024 * 1024)

#include "fftw3.h"
#include "fftw3-mpi.h"

#ifdef __cplusplus
extern "C" {
#endif

/* tunables, offt.h:69-100 */
struct _offt_params {
  int is_converged;
  int is_infeasible;
  int is_in_database;
#define LOG0 (-1)
#define _P1_ 0   /* process grid p1 x p2, p1 ranks along x */
#define _T1_ 1   /* phase-1 tile thickness (x planes per exchange) */
#define _W1_ 2   /* phase-1 window: tiles in flight = ring depth - 1 */
#define _Px1_ 3  /* CPU cache sub-tile: accepted, shapes CTA tiles at most */
#define _Py1_ 4
#define _Fz_ 5   /* MPI_Test frequencies: accepted and ignored (NCCL progresses by itself) */
#define _FP1_ 6
#define _Ux1_ 7
#define _Uz1_ 8
#define _FU1_ 9
#define _Fy1_ 10
#define _Ry_ 11  /* share (0-10) of the y transforms done before the mid point */
#define _T2_ 12  /* phase-2 tile thickness (z planes per exchange) */
#define _W2_ 13
#define _Pz2_ 14
#define _Px2_ 15
#define _Fy2_ 16
#define _FP2_ 17
#define _Uz2_ 18
#define _Uy2_ 19
#define _FU2_ 20
#define _Fx_ 21
#define _V_ 22   /* bit 1: exact counts in phase 1, bit 0: in phase 2 */
#define _S_ 23   /* 0: output z-y-x (or y-z-x), 1: output x-y-z */
#define PARAM_COUNT 24
  int v[PARAM_COUNT];
};

/* process grid and local boxes, offt.h:102-142 */
struct _offt_comm {
  int p1;
  int p2;
  MPI_Comm *comm1;
  MPI_Comm *comm2;
  MPI_Group *group1;
  MPI_Group *group2;
  int M1, M2, M3, M4; /* ceil(Nx/p1) ceil(Ny/p2) ceil(Nz/p2) ceil(Ny/p1) */
  int F1, F2, F3, F4; /* floors of the same */
  int m1, m2, m3, m4; /* this rank's share */
  int b1, b2, b3, b4; /* ranks holding one extra item */
  int istart[3];
  int isize[3];
  int istride[3];
  int ostart[3];
  int osize[3];
  int ostride[3];
};

struct _offt_plan {
  int p;
  int rank;
  int Nx;
  int Ny;
  int Nz;
  int is_r2c;
  int fftw_flag;
  int ah_strategy;
  int max_loop;
  int tuning_mode;
  int is_W0;
  int extrapolation_window;
  int is_oned;
  int is_a2a;
  int is_equalxy;
  int is_notest;
#define INIT_ALL 0
#define INIT_FFTW 1
#define INIT_AH 2
#define INIT_BUFFER 3
#define T_INIT_COUNT 4
  double t_init[T_INIT_COUNT];
#define ALL 0
#define INIT1 1
#define WAIT1 2
#define TEST1 3
#define INIT2 4
#define WAIT2 5
#define TEST2 6
#define FFTz 7
#define FFTy1 8
#define FFTy2 9
#define FFTx 10
#define TRANSPOSE 11
#define PACK1 12
#define UNPACK1 13
#define PACK2 14
#define UNPACK2 15
#define GES 16
  double t[GES];
#if !defined(SHSONG_HOPPER) && !defined(SHSONG_EDISON)
  int bad_tile_count; /* offt.h:189-190, kept so both flag sets see the same layout they always did */
#endif
  char point_database_file[256];
  char user_vertex_file[256];
  struct _offt_params *params;
  struct _offt_comm *comm;
  void *buffer_chunk;
  void *buffers1;
  void *buffers2;
  fftw_plan pt_transpose;
  fftw_plan *pt_transpose_list;
  int pt_transpose_list_size;
  fftw_plan p1d_x;
  fftw_plan p1d_y;
  fftw_plan p1d_z;
  fftw_plan p1d_x_t;
  fftw_plan p1d_y_t;
  fftw_plan *p1d_x_s_list;
  fftw_plan *p1d_y_s_list;
  int p1d_xy_s_list_size;
  void *b200; /* engine state (streams, device rings, twiddle tables); appended, never read by callers */
};

/* reference offt.h:235-244 */
struct _offt_plan *offt_3d_init(int Nx, int Ny, int Nz, double *in, double *out, int is_r2c, int fftw_flag,
                                int is_oned, int is_a2a, int is_equalxy, int is_notest, int ah_strategy,
                                int max_loop, int tuning_mode, int is_W0, int extrapolation_window,
                                struct _offt_params *custom_params);
void offt_3d_fin(struct _offt_plan *po);
void offt_3d_execute(struct _offt_plan *po, double *in, double *out, int is_tuning);
void print_params(int *v);
void offt_print_time(double *t);

/* reference offt-internal.h:26-38 (the pieces that make sense without MPI/FFTW objects) */
struct _offt_comm *offt_comm_malloc(struct _offt_plan *po);
void offt_comm_free(struct _offt_comm *comm);
int grid_value_floor(int is_index, int **v_list, int *v_list_size, int i, int raw_v);
int grid_value_ceil(int is_index, int **v_list, int *v_list_size, int i, int raw_v);
void params_range_setup(struct _offt_plan *po, int **v_list, int *v_list_size);
int ah_tuning(struct _offt_plan *po, double *in, double *out);
/* reference offt-tuning.c:80, 144 (the index <-> value hook and the feasibility hook of the tuner) */
void params_convert(int is_backward, int *v, long *ahv, struct _offt_plan *po, int **v_list, int *v_list_size);
int is_infeasible_point(struct _offt_plan *po, int *v, int *p_i);
void params_set_default(struct _offt_plan *po); /* reference offt-compute.c:3127 */

#ifdef __cplusplus
}
#endif

#ifndef OFFT_NO_MINMAX /* reference offt.h:251-257 */
#ifndef __GNUC__
#define __inline__ inline
#endif
static __inline__ int max(int a, int b) { return (a > b) ? a : b; }
static __inline__ int min(int a, int b) { return (a < b) ? a : b; }
#endif

#endif /* OFFT_INCLUDE */
