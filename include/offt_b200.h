/*
 * offt_b200.h - entry points the B200 library adds next to the reference's five
 * (offt.h).  Plain C ABI: pointers and integers only.
 *
 * The reference reaches its ranks through MPI_COMM_WORLD implicitly
 * (offt-compute.c:3315-3316).  Here the "world" is process-global state that must be set
 * once before offt_3d_init, by exactly one of
 *   - offtb_world_init        one OS process per GPU (torchrun, offtrun, real MPI),
 *   - offtb_world_init_local  all ranks emulated by this process on one GPU
 *                             (parity tests of the multi-rank schedules on a 1-GPU box),
 *   - the compat MPI_Init of include/compat/mpi.h (does offtb_world_init itself).
 * Every function returns 0 on success and a negative code on failure unless stated;
 * offtb_last_error() holds the message.  Nothing here falls back to the CPU.
 */
#ifndef OFFT_B200_INCLUDE
#define OFFT_B200_INCLUDE

#include <stddef.h>
#include "offt.h"

#ifdef __cplusplus
extern "C" {
#endif

#define OFFTB_UNIQUE_ID_BYTES 128

/* ---- world ---------------------------------------------------------------- */
int offtb_get_unique_id(void *id128);                 /* rank 0: ncclGetUniqueId */
int offtb_world_init(int rank, int size, int device, const void *id128); /* id may be NULL when size == 1 */
int offtb_world_init_local(int size, int device);     /* `size` emulated ranks on one device */
int offtb_world_set_rank(int rank);                   /* local worlds: rank the next offt_3d_init is for */
int offtb_world_size(void);
int offtb_world_rank(void);
int offtb_world_barrier(void);                        /* device-side all-reduce of one int, then stream sync */
void offtb_world_fin(void);
const char *offtb_last_error(void);
/* 1 (default): fatal conditions print and exit(-1) like the reference (offt-compute.c:3440-3443);
 * 0: the void entry points return after recording the error (bindings poll offtb_last_error) */
int offtb_set_exit_on_error(int on);
void offtb_clear_error(void);

/* ---- plan options (set between offt_3d_init and the first execute) -------- */
/* precision for plans created afterwards: 64 (default, the reference's only mode) or 32 */
int offtb_set_default_precision(int bits);
int offtb_plan_precision(const struct _offt_plan *po);
/* 1: every launch runs on the any-length kernel (fft_generic.cu) also where a power-of-two kernel exists - parity tests
 * of that kernel at sizes the fast kernels normally take; 0 (default): only where it is needed (other lengths, uneven splits) */
int offtb_set_force_generic(int on);
/* run on the caller's CUDA stream (cudaStream_t passed as void*); NULL = the plan's own */
int offtb_plan_set_stream(struct _offt_plan *po, void *stream);
/* 1: offt_3d_execute returns right after enqueueing (device pointers only) */
int offtb_plan_set_async(struct _offt_plan *po, int is_async);
/* complex elements the caller's in-place array must hold (run-fft.c:294-304, 64-bit) */
long long offtb_plan_alloc_elems(const struct _offt_plan *po);
long long offtb_alloc_elems(int Nx, int Ny, int Nz, int p, int p1);
long long offtb_alloc_elems_r2c(int Nx, int Ny, int Nz, int p, int p1, int is_r2c);   /* run-fft.c:296-300 with M3 from Nz/2+1 */
/* kernels launched / milliseconds on the device by the last execute of this plan */
int offtb_plan_last_launches(const struct _offt_plan *po);
double offtb_plan_last_ms(const struct _offt_plan *po);
/* per-stage device time of the last execute: fills up to n of
 * {K1 fftz(+pack1), K2 (unpack1+)ffty, K3 ffty(+pack2), K4 (unpack2+)fftx, exchange1, exchange2, h2d, d2h} in ms */
int offtb_plan_stage_ms(const struct _offt_plan *po, double *ms, int n);
int offtb_plan_set_stage_timing(struct _offt_plan *po, int on);

/* ---- execution extensions -------------------------------------------------- */
/* backward transform, unnormalised (FFTW_BACKWARD convention): takes the layout
 * offt_3d_execute produced (ostart/osize/ostride) and returns the input layout. */
int offt_3d_execute_inverse(struct _offt_plan *po, double *in, double *out);
/* local worlds only: run the plans of all emulated ranks together; arrays[r] belongs to plans[r] */
int offtb_execute_group(struct _offt_plan **plans, double **arrays, int n, int inverse);

/* ---- the batched 1-D kernel on its own (kernel-level parity tests and roofline runs) ---- */
/* transforms `howmany` rows of length n in place on the device: element j of row h at
 * data[h*dist + j*stride] (complex elements, precision bits 64/32).  sign -1/+1.
 * Exactly one of stride, dist must be 1.  `repeat` launches back to back; returns ms per launch or <0. */
double offtb_fft_rows(void *device_data, int n, long long stride, long long dist, long long howmany,
                      int sign, int bits, int repeat, void *stream);

/* the same kernel with both address maps spelled out.  A map is 9 integers
 * {off, nlo_count (0: no split), n_hi, n_lo, B0, s0, B1, s1, s2}: element (n, b) lives at
 * off + (n / nlo_count)*n_hi + (n % nlo_count)*n_lo + b0*s0 + b1*s1 + b2*s2 with b = b0 + B0*(b1 + B1*b2).
 * c_log < 0 picks the columns per CTA automatically; ry_level < 0 disables the Ry rule. */
double offtb_fft_launch_raw(const void *in, void *out, int n, int bits, int sign, long long nbatch,
                            const long long *im9, const long long *om9, int c_log, int load_cfast, int store_cfast,
                            int ry_level, int ry_x0, int ry_lo, int ry_hi, int repeat, void *stream);

/* ---- tunables: the reference's parameter space, callable without a GPU ------ */
void offtb_params_default(int Nx, int Ny, int Nz, int p, int is_W0, int is_notest, int *v24);
int offtb_is_infeasible_point(int Nx, int Ny, int Nz, int p, const int *v24, int *bad_index);
void offtb_params_adjust(int Nx, int Ny, int Nz, int p, int is_oned, int *v24);
/* fills lists[i*stride .. ] with the value grid of tunable i, sizes[i] with its length */
void offtb_params_range(int Nx, int Ny, int Nz, int p, int *lists, int stride, int *sizes);
/* rank-local box for (p, p1, rank) without creating a plan */
int offtb_comm_fill(struct _offt_comm *c, int Nx, int Ny, int Nz, int p, int p1, int rank, int S, int is_equalxy);
/* the same for real-to-complex plans (is_r2c: Nz/2+1 complex points per z row, offt-compute.c:63) */
int offtb_comm_fill_r2c(struct _offt_comm *c, int Nx, int Ny, int Nz, int p, int p1, int rank, int S, int is_equalxy, int is_r2c);
/* 0 if the library can run this problem, else a negative code (message in offtb_last_error) */
int offtb_check_supported(int Nx, int Ny, int Nz, int p, int p1);
/* exchange bookkeeping of one tile: blocks per peer in complex elements (phase 1 or 2) */
long long offtb_exchange_block_elems(const struct _offt_plan *po, int phase, int myT);
/* which of a phase's nb tiles the schedule visits i-th (-1: none): ascending, except the backward transform's phase 1,
 * which goes down so that an in-place exchange never overwrites planes a later tile still reads (plan.cu, run_phase) */
int offtb_tile_visited(int nb, int visit, int phase, int inverse);

/* ---- built-in search over the tunables (the fetch / report loop of ah_tuning) ---- */
/* evaluates up to max_loop feasible points on the device and leaves the best in po->params */
int offtb_tune(struct _offt_plan *po, double *in, double *out, int max_loop, int verbose);
/* the same with the candidate source spelled out.  strategy: 0 / 1 Nelder-Mead over all 24 index coordinates from the
 * reference's initial simplex (write_initial_simplex, offt-tuning.c:426-737), 2 random grid points, 3 coordinate descent
 * along the tunables that change the GPU schedule (what offtb_tune does).  search_p1 != 0 also searches what
 * changes the caller's layout - the decomposition _P1_ and the output order _S_; every trial then rebuilds the layout
 * descriptor (offt-tuning.c:929-948), so the caller must read po->comm only afterwards, as run-fft.c does.  Without it
 * only tile sizes, windows and _Ry_ move and istride / ostride stay what they were.  Trials run on an internal zeroed
 * device array. */
int offtb_tune_ex(struct _offt_plan *po, double *in, double *out, int max_loop, int verbose, int strategy, int search_p1);
/* the same loop with the reference's Active Harmony server proposing the points (strategy 0 nm.so, 1 pro.so, 2 random.so,
 * 3 brute.so) where offt_b200/ah/_root was built, the built-in sources otherwise; what ah_tuning / `run-fft -l` use */
int offtb_tune_harmony(struct _offt_plan *po, int max_loop, int verbose, int strategy, int search_p1);
/* offt-tuning.c:426-737: writes the 25 starting vertices (grid indices) of the Nelder-Mead search to po->user_vertex_file */
void write_initial_simplex(struct _offt_plan *po, int **v_list, int *v_list_size);

#ifdef __cplusplus
}
#endif
#endif /* OFFT_B200_INCLUDE */
