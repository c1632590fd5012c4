/*
 * TEST INFRASTRUCTURE - not product code.  Only tests/, __graft_entry__.smoke()
 * and bench.py's CPU-baseline legs may load this library.
 *
 * CPU restatement of the distributed 3-D complex FFT path of rchyena/offt
 * (offt_3d_execute, offt-compute.c:3864-4048) for the flag set the reference's
 * own Makefile uses (-DA2AV -DSTRIDE, /root/reference/Makefile:27-29), C2C only.
 * All `p` ranks are simulated inside one process: every rank owns its in-place
 * array and its a2as/a2ar pair, and the all-to-all is a memcpy between them.
 * The tile loops keep the reference's T1/T2 tiling (tile thickness changes the
 * buffer layout) but not its window W (W only reorders independent work).
 *
 * Pinned (tests/test_oracle.py) against
 *   - outputs of the unmodified reference run in the build container
 *     (oracle/_ref/ref_dump) and the golden fixtures made from them,
 *   - the closed-form DFT of run-fft.c's input ramp (run-fft.c:46-61, 452-503),
 *   - numpy.fft.fftn as an independent arithmetic check.
 * The 1-D arithmetic is oracle_dft.h (FFTW, which the reference calls, is an
 * un-vendored, un-pinned third-party dependency: Makefile:17-24 only mentions
 * a site path to fftw-3.3.2).
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "oracle_dft.h"

#define PARAM_COUNT 24
enum { P_P1, P_T1, P_W1, P_Px1, P_Py1, P_Fz, P_FP1, P_Ux1, P_Uz1, P_FU1, P_Fy1, P_Ry,
       P_T2, P_W2, P_Pz2, P_Px2, P_Fy2, P_FP2, P_Uz2, P_Uy2, P_FU2, P_Fx, P_V, P_S };
#define BUFFER_SIZE_LIMIT (32 * 1024 * 1024) /* offt.h:51 */
#define LOG0 (-1)

static int imin(int a, int b) { return a < b ? a : b; }
static int imax(int a, int b) { return a > b ? a : b; }

/* ------------------------------------------------------------------ layout */

/* struct _offt_comm, offt.h:102-142 (A2AV variant), without the MPI handles */
typedef struct {
  int p1, p2, rank_x, rank_y;
  int M1, M2, M3, M4, F1, F2, F3, F4, m1, m2, m3, m4, b1, b2, b3, b4;
  int istart[3], isize[3], istride[3], ostart[3], osize[3], ostride[3];
} ocomm;

/* offt_comm_malloc, offt-compute.c:57-315 */
void oracle_comm_fill(ocomm *c, int Nx, int Ny, int Nz, int p, int p1, int rank, int S, int is_equalxy) {
  int p2 = p / p1;
  c->p1 = p1; c->p2 = p2;
  int rank_x = c->rank_x = rank / p2;                        /* :74-76 */
  int rank_y = c->rank_y = rank % p2;
  c->M1 = (Nx + p1 - 1) / p1; c->M2 = (Ny + p2 - 1) / p2;   /* :128-139 */
  c->M3 = (Nz + p2 - 1) / p2; c->M4 = (Ny + p1 - 1) / p1;
  c->F1 = Nx / p1; c->F2 = Ny / p2; c->F3 = Nz / p2; c->F4 = Ny / p1;
  c->b1 = Nx % p1; c->b2 = Ny % p2; c->b3 = Nz % p2; c->b4 = Ny % p1;
  c->m1 = (rank_x < p1 - c->b1) ? c->F1 : c->F1 + 1;        /* :141-144 */
  c->m2 = (rank_y < p2 - c->b2) ? c->F2 : c->F2 + 1;
  c->m3 = (rank_y < p2 - c->b3) ? c->F3 : c->F3 + 1;
  c->m4 = (rank_x < p1 - c->b4) ? c->F4 : c->F4 + 1;
  c->istart[0] = (rank_x < p1 - c->b1) ? rank_x * c->F1 : (p1 - c->b1) * c->F1 + (rank_x - (p1 - c->b1)) * (c->F1 + 1); /* :246 */
  c->istart[1] = (rank_y < p2 - c->b2) ? rank_y * c->F2 : (p2 - c->b2) * c->F2 + (rank_y - (p2 - c->b2)) * (c->F2 + 1); /* :247 */
  c->istart[2] = 0;
  c->isize[0] = c->m1; c->isize[1] = c->m2; c->isize[2] = Nz;
  c->istride[0] = (c->M2 * p2 > c->M4 * p1) ? c->M2 * c->M3 * p2 : c->M4 * p1 * c->M3;  /* :260-263 */
  c->istride[1] = c->M3 * p2;
  c->istride[2] = 1;
  c->ostart[0] = 0;                                          /* :266-274 */
  c->ostart[1] = (rank_x < p1 - c->b4) ? rank_x * c->F4 : (p1 - c->b4) * c->F4 + (rank_x - (p1 - c->b4)) * (c->F4 + 1);
  c->ostart[2] = (rank_y < p2 - c->b3) ? rank_y * c->F3 : (p2 - c->b3) * c->F3 + (rank_y - (p2 - c->b3)) * (c->F3 + 1);
  c->osize[0] = Nx; c->osize[1] = c->m4; c->osize[2] = c->m3;
  if (S) {                                                   /* :283-287  x-y-z */
    c->ostride[0] = c->M3 * c->M4; c->ostride[1] = c->M3; c->ostride[2] = 1;
  } else if (is_equalxy && c->M1 == c->M4) {                 /* :289-293  y-z-x */
    c->ostride[0] = 1; c->ostride[1] = c->M1 * p1 * c->M3; c->ostride[2] = c->M1 * p1;
  } else {                                                   /* :295-298  z-y-x */
    c->ostride[0] = 1; c->ostride[1] = c->M1 * p1; c->ostride[2] = c->M1 * p1 * c->M4;
  }
}

/* complex elements of the caller's in-place array, run-fft.c:294-300 (in 64-bit) */
int64_t oracle_alloc_elems(int Nx, int Ny, int Nz, int p, int p1) {
  int p2 = p / p1;
  int64_t M1 = (Nx + p1 - 1) / p1, M2 = (Ny + p2 - 1) / p2, M3 = (Nz + p2 - 1) / p2, M4 = (Ny + p1 - 1) / p1;
  return (M2 * p2 > M4 * p1) ? M1 * M2 * M3 * p2 : M1 * M3 * M4 * p1;
}

/* The block that recurs ~30 times in offt-compute.c (e.g. :1000-1027): which of
 * `g` owners holds index `idx` when the first g-b own F items and the last b own
 * F+1; *off = first index of that owner. */
static void owner_of(int F, int b, int g, int idx, int *a, int *off) {
  if (F * (g - b) <= idx) {
    *a = (idx - F * (g - b)) / (F + 1) + (g - b);
    *off = (g - b) * F + (*a - (g - b)) * (F + 1);
  } else {
    *a = idx / F;
    *off = *a * F;
  }
}

/* ------------------------------------------------------------- rank state */

typedef struct {
  ocomm c;
  odft_cplx *out;   /* the caller's in-place array */
  odft_cplx *a2as, *a2ar;
} orank;

typedef struct {
  int Nx, Ny, Nz, p;
  int v[PARAM_COUNT];
  int is_oned, is_equalxy;
  odft_plan px, py, pz;
  odft_cplx *row;   /* scratch row for in-place 1-D transforms */
  orank *r;
} oworld;

/* fftw_execute_dft on an in-place strided row (what shim_fftw.c does for in == out) */
static void row_fft(oworld *w, const odft_plan *pl, odft_cplx *ptr, int64_t stride) {
  int j;
  odft_exec(pl, w->row, ptr, (int)stride);
  for (j = 0; j < pl->n; j++) ptr[(int64_t)j * stride] = w->row[j];
}

/* ------------------------------------------------ K1  offt-compute.c:905-1206 */
static void fftz_pack1(oworld *w, orank *R, int tile_ind, int myT) {
  ocomm *c = &R->c;
  int T = w->v[P_T1], Nz = w->Nz;
  int F3 = c->F3, b3 = c->b3, m2 = c->m2, p2 = c->p2, M2 = c->M2, M3 = c->M3;
  int is_a2av = w->v[P_V] & 2;                                     /* :920 */
  int from_x = tile_ind * T, to_x = from_x + myT;
  int x, y, z;
  for (x = from_x; x < to_x; x++)
    for (y = 0; y < c->m2; y++)
      row_fft(w, &w->pz, R->out + (int64_t)c->istride[1] * y + (int64_t)c->istride[0] * x, 1);  /* :959-963 */
  for (x = from_x; x < to_x; x++)
    for (y = 0; y < c->m2; y++)
      for (z = 0; z < Nz; z++) {
        int a, z_off;
        int64_t B, dst;
        owner_of(F3, b3, p2, z, &a, &z_off);
        const odft_cplx *src = R->out + z + (int64_t)c->istride[1] * y + (int64_t)c->istride[0] * x;
        if (w->v[P_S]) {                                           /* :989-1035  xyz -> xyzB */
          int Sy, Sx;
          if (is_a2av) {
            int big = (F3 * (p2 - b3) <= z);
            B = big ? (int64_t)(p2 - b3) * ((int64_t)myT * m2 * F3) + (int64_t)(a - (p2 - b3)) * ((int64_t)myT * m2 * (F3 + 1))
                    : (int64_t)a * ((int64_t)myT * m2 * F3);
            Sy = big ? F3 + 1 : F3; Sx = m2 * Sy;
          } else { B = (int64_t)a * ((int64_t)myT * M2 * M3); Sy = M3; Sx = M2 * M3; }
          dst = B + (z - z_off) + (int64_t)y * Sy + (int64_t)(x - from_x) * Sx;          /* :1030 */
        } else {                                                   /* :1059-1118  xyz -> xzyB */
          int Sz, Sx;
          if (is_a2av) {
            int big = (F3 * (p2 - b3) <= z);
            B = big ? (int64_t)(p2 - b3) * ((int64_t)myT * m2 * F3) + (int64_t)(a - (p2 - b3)) * ((int64_t)myT * m2 * (F3 + 1))
                    : (int64_t)a * ((int64_t)myT * m2 * F3);
            Sz = m2; Sx = m2 * (big ? F3 + 1 : F3);
          } else { B = (int64_t)a * ((int64_t)myT * M2 * M3); Sz = M2; Sx = M2 * M3; }
          dst = B + y + (int64_t)(z - z_off) * Sz + (int64_t)(x - from_x) * Sx;          /* :1107 */
        }
        R->a2as[dst] = *src;
      }
}

/* ----------------------------------------------- K2  offt-compute.c:1208-1520 */
static void unpack1_ffty(oworld *w, orank *R, int tile_ind, int myT) {
  ocomm *c = &R->c;
  int T = w->v[P_T1], Ny = w->Ny;
  int F2 = c->F2, b2 = c->b2, M2 = c->M2, M3 = c->M3, M4 = c->M4, m3 = c->m3, p1 = c->p1, p2 = c->p2;
  int is_a2av = w->v[P_V] & 2;
  int from_x = tile_ind * T, to_x = from_x + myT;
  int Ry = w->v[P_Ry];
  int is_ignore_Ry = (w->is_oned && c->p1 == 1);                   /* :1240 */
  int x, y, z;
  for (x = from_x; x < to_x; x++)
    for (z = 0; z < c->m3; z++)
      for (y = 0; y < Ny; y++) {
        int a, y_off;
        int64_t B, src, dst;
        owner_of(F2, b2, p2, y, &a, &y_off);
        int big = (F2 * (p2 - b2) <= y);
        if (w->v[P_S]) {                                           /* :1266-1315  xyzB -> xyz */
          int Sy, Sx;
          if (is_a2av) {
            B = big ? (int64_t)(p2 - b2) * ((int64_t)myT * F2 * m3) + (int64_t)(a - (p2 - b2)) * ((int64_t)myT * (F2 + 1) * m3)
                    : (int64_t)a * ((int64_t)myT * F2 * m3);
            Sy = m3; Sx = (big ? F2 + 1 : F2) * m3;
          } else { B = (int64_t)a * ((int64_t)myT * M2 * M3); Sy = M3; Sx = M2 * M3; }
          src = B + z + (int64_t)(y - y_off) * Sy + (int64_t)(x - from_x) * Sx;          /* :1310 */
          dst = z + (int64_t)M3 * y + (int64_t)M3 * M4 * p1 * x;                          /* :1309 */
        } else {                                                   /* :1339-1397  xzyB -> xzy */
          int Sz, Sx;
          if (is_a2av) {
            B = big ? (int64_t)(p2 - b2) * ((int64_t)myT * F2 * m3) + (int64_t)(a - (p2 - b2)) * ((int64_t)myT * (F2 + 1) * m3)
                    : (int64_t)a * ((int64_t)myT * F2 * m3);
            Sz = big ? F2 + 1 : F2; Sx = Sz * m3;
          } else { B = (int64_t)a * ((int64_t)myT * M2 * M3); Sz = M2; Sx = M2 * M3; }
          src = B + (y - y_off) + (int64_t)z * Sz + (int64_t)(x - from_x) * Sx;          /* :1384 */
          dst = y + (int64_t)M4 * p1 * (z + (int64_t)M3 * x);                             /* :1383 */
        }
        R->out[dst] = R->a2ar[src];
      }
  for (x = from_x; x < to_x; x++)                                  /* :1479-1495 */
    for (z = 0; z < c->m3; z++)
      if (is_ignore_Ry || x % 10 < Ry) {
        if (w->v[P_S]) row_fft(w, &w->py, R->out + z + (int64_t)M4 * p1 * M3 * x, M3);
        else row_fft(w, &w->py, R->out + (int64_t)M4 * p1 * (z + (int64_t)M3 * x), 1);
      }
}

/* ----------------------------------------------- K3  offt-compute.c:1636-2345 */
static void ffty_pack2(oworld *w, orank *R, int tile_ind, int myT) {
  ocomm *c = &R->c;
  int T = w->v[P_T2], Ny = w->Ny;
  int M1 = c->M1, M3 = c->M3, M4 = c->M4, p1 = c->p1, F4 = c->F4, m1 = c->m1, b4 = c->b4;
  int is_a2av = w->v[P_V] & 1;                                     /* :1651 */
  int from_z = tile_ind * T, to_z = from_z + myT;
  int Ry = w->v[P_Ry];
  int is_ignore_Ry = (w->is_oned && c->p1 == w->p);                /* :1690 */
  int eq = (w->is_equalxy && c->M1 == c->M4);
  int x, y, z;
  for (x = 0; x < c->m1; x++)
    for (z = from_z; z < to_z; z++)
      if (is_ignore_Ry || x % 10 >= Ry) {
        if (w->v[P_S]) row_fft(w, &w->py, R->out + z + (int64_t)M4 * p1 * ((int64_t)M3 * x), M3);   /* :1709 */
        else if (eq) row_fft(w, &w->py, R->out + (int64_t)M4 * p1 * (z + (int64_t)M3 * x), 1);       /* :1842 */
        else row_fft(w, &w->py, R->out + (int64_t)M4 * p1 * (x + (int64_t)M1 * z), 1);               /* :1989 */
      }
  for (x = 0; x < c->m1; x++)
    for (y = 0; y < Ny; y++)
      for (z = from_z; z < to_z; z++) {
        int a, y_off;
        int64_t B, src, dst;
        owner_of(F4, b4, p1, y, &a, &y_off);
        int big = (F4 * (p1 - b4) <= y);
        int Fy = big ? F4 + 1 : F4;
        if (is_a2av)
          B = big ? (int64_t)(p1 - b4) * ((int64_t)m1 * F4 * myT) + (int64_t)(a - (p1 - b4)) * ((int64_t)m1 * (F4 + 1) * myT)
                  : (int64_t)a * ((int64_t)m1 * F4 * myT);
        else
          B = (int64_t)a * ((int64_t)M1 * M4 * myT);
        if (w->v[P_S]) {                                           /* :1734-1779  xyz -> xyzB */
          int Sy = myT, Sx = myT * (is_a2av ? Fy : M4);
          dst = B + (z - from_z) + (int64_t)(y - y_off) * Sy + (int64_t)x * Sx;          /* :1774 */
          src = z + (int64_t)M3 * y + (int64_t)M3 * M4 * p1 * x;                          /* :1775 */
        } else if (eq) {                                           /* :1867-1913  xzy -> yzxB */
          int Sz = is_a2av ? m1 : M1, Sy = Sz * myT;
          dst = B + x + (int64_t)(z - from_z) * Sz + (int64_t)(y - y_off) * Sy;          /* :1909 */
          src = y + (int64_t)Ny * z + (int64_t)M4 * p1 * M3 * x;                          /* :1910 */
        } else {                                                   /* :2016-2061  zxy -> zyxB */
          int Sy = is_a2av ? m1 : M1, Sz = Sy * (is_a2av ? Fy : M4);
          dst = B + x + (int64_t)(y - y_off) * Sy + (int64_t)(z - from_z) * Sz;          /* :2056 */
          src = y + (int64_t)M4 * p1 * x + (int64_t)M4 * p1 * M1 * z;                     /* :2057 */
        }
        R->a2as[dst] = R->out[src];
      }
}

/* ----------------------------------------------- K4  offt-compute.c:2347-2993 */
static void unpack2_fftx(oworld *w, orank *R, int tile_ind, int myT) {
  ocomm *c = &R->c;
  int T = w->v[P_T2], Nx = w->Nx;
  int M1 = c->M1, M3 = c->M3, M4 = c->M4, F1 = c->F1, m4 = c->m4, b1 = c->b1, p1 = c->p1;
  int is_a2av = w->v[P_V] & 1;
  int from_z = tile_ind * T, to_z = from_z + myT;
  int eq = (w->is_equalxy && c->M1 == c->M4);
  int x, y, z;
  for (x = 0; x < Nx; x++)
    for (y = 0; y < c->m4; y++)
      for (z = from_z; z < to_z; z++) {
        int a, x_off;
        int64_t B, src, dst;
        owner_of(F1, b1, p1, x, &a, &x_off);
        int big = (F1 * (p1 - b1) <= x);
        int Fx = big ? F1 + 1 : F1;
        if (is_a2av)
          B = big ? (int64_t)(p1 - b1) * ((int64_t)F1 * m4 * myT) + (int64_t)(a - (p1 - b1)) * ((int64_t)(F1 + 1) * m4 * myT)
                  : (int64_t)a * ((int64_t)F1 * m4 * myT);
        else
          B = (int64_t)a * ((int64_t)M1 * M4 * myT);
        if (w->v[P_S]) {                                           /* :2409-2452  xyzB -> xyz */
          int Sy = myT, Sx = myT * (is_a2av ? m4 : M4);
          src = B + (z - from_z) + (int64_t)y * Sy + (int64_t)(x - x_off) * Sx;          /* :2450 */
          dst = z + (int64_t)M3 * y + (int64_t)M3 * M4 * x;                               /* :2449 */
        } else if (eq) {                                           /* :2529-2574  yzxB -> yzx */
          int Sz = is_a2av ? Fx : M1, Sy = Sz * myT;
          src = B + (x - x_off) + (int64_t)(z - from_z) * Sz + (int64_t)y * Sy;          /* :2570 */
          dst = x + (int64_t)M1 * p1 * (z + (int64_t)M3 * y);                             /* :2569 */
        } else {                                                   /* :2645-2690  zyxB -> zyx */
          int Sy = is_a2av ? Fx : M1, Sz = Sy * (is_a2av ? m4 : M4);
          src = B + (x - x_off) + (int64_t)y * Sy + (int64_t)(z - from_z) * Sz;          /* :2688 */
          dst = x + (int64_t)M1 * p1 * (y + (int64_t)M4 * z);                             /* :2687 */
        }
        R->out[dst] = R->a2ar[src];
      }
  for (y = 0; y < c->m4; y++)
    for (z = from_z; z < to_z; z++) {
      if (w->v[P_S]) row_fft(w, &w->px, R->out + z + (int64_t)M3 * y, (int64_t)M3 * M4);             /* :2493-2495 */
      else if (eq) row_fft(w, &w->px, R->out + (int64_t)M1 * p1 * (z + (int64_t)M3 * y), 1);         /* :2612-2613 */
      else row_fft(w, &w->px, R->out + (int64_t)M1 * p1 * (y + (int64_t)M4 * z), 1);                 /* :2729-2730 */
    }
}

/* ------------------------------- setup_transpose + fftw guru copy, :523-653 */
static void local_transpose(oworld *w, orank *R, int64_t alloc) {
  ocomm *c = &R->c;
  int n[3];
  int64_t is[3], os[3];
  int eq = (w->is_equalxy && c->M1 == c->M4);
  if (w->is_oned && c->p1 == 1) {
    n[0] = c->M1; is[0] = (int64_t)c->M4 * c->M3; os[0] = 1;
    n[1] = c->M3; is[1] = c->M4;
    n[2] = c->M4; is[2] = 1;
    if (eq) { os[1] = c->M1; os[2] = (int64_t)c->M1 * c->M3; }            /* :563-573  xzy -> yzx */
    else { os[1] = (int64_t)c->M1 * c->M4; os[2] = c->M1; }                /* :575-584  xzy -> zyx */
  } else if (w->is_oned && c->p1 == w->p) {
    n[0] = c->M1; is[0] = (int64_t)c->M4 * c->p1 * c->M3;
    n[1] = c->M4 * c->p1; is[1] = c->M3; os[1] = 1;
    n[2] = c->M3; is[2] = 1;
    if (eq) { os[0] = (int64_t)c->M4 * c->p1 * c->M3; os[2] = (int64_t)c->M4 * c->p1; }            /* :587-598  xyz -> xzy */
    else { os[0] = (int64_t)c->M4 * c->p1; os[2] = (int64_t)c->M4 * c->p1 * c->M1; }               /* :600-610  xyz -> zxy */
  } else {
    if (eq) return;                                                        /* :613-623  nothing to move */
    n[0] = c->M1; is[0] = (int64_t)c->M4 * c->p1 * c->M3; os[0] = (int64_t)c->M4 * c->p1;          /* :625-634  xzy -> zxy */
    n[1] = c->M3; is[1] = (int64_t)c->M4 * c->p1; os[1] = (int64_t)c->M4 * c->p1 * c->M1;
    n[2] = c->M4 * c->p1; is[2] = 1; os[2] = 1;
  }
  int64_t ext = 1 + (n[0] - 1) * is[0] + (n[1] - 1) * is[1] + (n[2] - 1) * is[2];
  if (ext > alloc) ext = alloc;
  odft_cplx *tmp = (odft_cplx *)malloc(sizeof(odft_cplx) * (size_t)ext);
  memcpy(tmp, R->out, sizeof(odft_cplx) * (size_t)ext);
  int i0, i1, i2;
  for (i0 = 0; i0 < n[0]; i0++)
    for (i1 = 0; i1 < n[1]; i1++)
      for (i2 = 0; i2 < n[2]; i2++)
        R->out[i0 * os[0] + i1 * os[1] + i2 * os[2]] = tmp[i0 * is[0] + i1 * is[1] + i2 * is[2]];
  free(tmp);
}

/* ------------- communicate_a2a / communicate_a2av, :836-881, between simulated ranks */
static void exchange(oworld *w, int phase, const int *members, int g, int myT) {
  int i, j;
  for (i = 0; i < g; i++) {          /* sender i */
    orank *Si = &w->r[members[i]];
    for (j = 0; j < g; j++) {        /* receiver j */
      orank *Rj = &w->r[members[j]];
      int64_t soff, roff, cnt;
      if (phase == 1) {
        if (w->v[P_V] & 2) {         /* :3512-3521 */
          int k;
          soff = 0; roff = 0;
          for (k = 0; k < j; k++) soff += (int64_t)myT * Si->c.m2 * (Si->c.F3 + (k >= g - Si->c.b3));
          for (k = 0; k < i; k++) roff += (int64_t)myT * (Rj->c.F2 + (k >= g - Rj->c.b2)) * Rj->c.m3;
          cnt = (int64_t)myT * Si->c.m2 * (Si->c.F3 + (j >= g - Si->c.b3));
        } else {                     /* :3523 */
          cnt = (int64_t)myT * Si->c.M2 * Si->c.M3; soff = cnt * j; roff = cnt * i;
        }
      } else {
        if (w->v[P_V] & 1) {         /* :3693-3702 */
          int k;
          soff = 0; roff = 0;
          for (k = 0; k < j; k++) soff += (int64_t)myT * Si->c.m1 * (Si->c.F4 + (k >= g - Si->c.b4));
          for (k = 0; k < i; k++) roff += (int64_t)myT * (Rj->c.F1 + (k >= g - Rj->c.b1)) * Rj->c.m4;
          cnt = (int64_t)myT * Si->c.m1 * (Si->c.F4 + (j >= g - Si->c.b4));
        } else {                     /* :3704 */
          cnt = (int64_t)myT * Si->c.M1 * Si->c.M4; soff = cnt * j; roff = cnt * i;
        }
      }
      memcpy(Rj->a2ar + roff, Si->a2as + soff, sizeof(odft_cplx) * (size_t)cnt);
    }
  }
}

/* offt_3d_execute_phase1 / phase2, :3501-3862, for one row / column of the process grid */
static void run_phase(oworld *w, int phase, const int *members, int g) {
  orank *R0 = &w->r[members[0]];
  int myblocks = (phase == 1) ? R0->c.m1 : R0->c.m3;               /* :3529, :3710 */
  int tiling = (phase == 1) ? w->v[P_T1] : w->v[P_T2];
  int blocks = (myblocks + tiling - 1) / tiling;
  int i, k;
  for (i = 0; i < blocks; i++) {
    int myT = (i == blocks - 1) ? myblocks - (blocks - 1) * tiling : tiling;   /* :3551-3552 */
    for (k = 0; k < g; k++) {
      if (phase == 1) fftz_pack1(w, &w->r[members[k]], i, myT);
      else ffty_pack2(w, &w->r[members[k]], i, myT);
    }
    exchange(w, phase, members, g, myT);
    for (k = 0; k < g; k++) {
      if (phase == 1) unpack1_ffty(w, &w->r[members[k]], i, myT);
      else unpack2_fftx(w, &w->r[members[k]], i, myT);
    }
  }
}

static void run_phase_all(oworld *w, int phase) {
  int p1 = w->v[P_P1], p2 = w->p / p1;
  int members[1024];
  int i, j;
  if (phase == 1)
    for (i = 0; i < p1; i++) {       /* comm1: ranks i*p2 .. (i+1)*p2-1, :93-101 */
      for (j = 0; j < p2; j++) members[j] = j + i * p2;
      run_phase(w, 1, members, p2);
    }
  else
    for (j = 0; j < p2; j++) {       /* comm2: ranks j, j+p2, ... */
      for (i = 0; i < p1; i++) members[i] = j + i * p2;
      run_phase(w, 2, members, p1);
    }
}

/*
 * offt_3d_execute, offt-compute.c:3864-4048, forward, in place, for all p ranks.
 * arrays[r] = rank r's in-place array (complex128 interleaved, oracle_alloc_elems long).
 * Returns 0, or -1 for arguments the reference would not survive.
 */
int oracle_execute(int Nx, int Ny, int Nz, int p, const int *v, int is_oned, int is_equalxy, double **arrays) {
  oworld w;
  int r, x, y;
  int p1 = v[P_P1];
  if (p < 1 || p > 1024 || p1 < 1 || p % p1 != 0 || v[P_T1] < 1 || v[P_T2] < 1) return -1;
  memset(&w, 0, sizeof(w));
  w.Nx = Nx; w.Ny = Ny; w.Nz = Nz; w.p = p; w.is_oned = is_oned; w.is_equalxy = is_equalxy;
  memcpy(w.v, v, sizeof(int) * PARAM_COUNT);
  odft_init(&w.px, Nx, -1); odft_init(&w.py, Ny, -1); odft_init(&w.pz, Nz, -1);
  w.row = (odft_cplx *)malloc(sizeof(odft_cplx) * (size_t)imax(imax(Nx, Ny), Nz));
  w.r = (orank *)calloc((size_t)p, sizeof(orank));
  int64_t alloc = oracle_alloc_elems(Nx, Ny, Nz, p, p1);
  for (r = 0; r < p; r++) {
    orank *R = &w.r[r];
    oracle_comm_fill(&R->c, Nx, Ny, Nz, p, p1, r, v[P_S], is_equalxy);
    R->out = (odft_cplx *)arrays[r];
    /* set_buffer_chunk, :697-699 */
    int64_t s1 = (int64_t)v[P_T1] * R->c.M2 * ((int64_t)R->c.M3 * R->c.p2);
    int64_t s2 = (int64_t)R->c.M1 * ((int64_t)R->c.M4 * R->c.p1) * v[P_T2];
    int64_t s = s1 > s2 ? s1 : s2;
    R->a2as = (odft_cplx *)calloc((size_t)s, sizeof(odft_cplx));
    R->a2ar = (odft_cplx *)calloc((size_t)s, sizeof(odft_cplx));
  }
  if (is_oned && p1 == 1) {                                        /* METHOD ONE, :3896-3950 */
    run_phase_all(&w, 1);
    for (r = 0; r < p; r++) {
      orank *R = &w.r[r];
      if (v[P_S]) {                                                /* :3932 with the plan of :400-404 */
        int h, howmany = R->c.M3 * R->c.M4;
        for (h = 0; h < howmany; h++) row_fft(&w, &w.px, R->out + h, (int64_t)R->c.M3 * R->c.M4);
      } else {
        local_transpose(&w, R, alloc);                             /* :3919 */
        int z;
        for (z = 0; z < R->c.m3; z++)                              /* :3935-3939 */
          for (y = 0; y < Ny; y++)
            row_fft(&w, &w.px, R->out + (int64_t)(R->c.M1 * R->c.p1) * (y + (int64_t)R->c.M4 * z), 1);
      }
    }
  } else if (is_oned && p1 == p) {                                 /* METHOD OLD, :3951-3998 */
    for (r = 0; r < p; r++) {
      orank *R = &w.r[r];
      for (x = 0; x < R->c.m1; x++)                                /* :3970-3979 */
        for (y = 0; y < Ny; y++)
          row_fft(&w, &w.pz, R->out + (int64_t)R->c.istride[1] * y + (int64_t)R->c.istride[0] * x, 1);
      if (!v[P_S]) local_transpose(&w, R, alloc);                  /* :3982-3989 */
    }
    run_phase_all(&w, 2);
  } else {                                                         /* p1 x p2, :3999-4037 */
    run_phase_all(&w, 1);
    if (!v[P_S])
      for (r = 0; r < p; r++)
        if (!(is_equalxy && w.r[r].c.M1 == w.r[r].c.M4)) local_transpose(&w, &w.r[r], alloc);     /* :4019-4025 */
    run_phase_all(&w, 2);
  }
  for (r = 0; r < p; r++) { free(w.r[r].a2as); free(w.r[r].a2ar); }
  free(w.r); free(w.row);
  odft_free(&w.px); odft_free(&w.py); odft_free(&w.pz);
  return 0;
}

/* Batched 1-D DFT of `howmany` rows (element stride `stride`, row distance `dist`),
 * the arithmetic check for single GPU kernels. sign = -1 forward, +1 backward. */
void oracle_dft_rows(double *data, int n, int64_t stride, int64_t dist, int64_t howmany, int sign) {
  odft_plan pl;
  odft_cplx *row = (odft_cplx *)malloc(sizeof(odft_cplx) * (size_t)n);
  odft_cplx *d = (odft_cplx *)data;
  int64_t h;
  int j;
  odft_init(&pl, n, sign);
  for (h = 0; h < howmany; h++) {
    odft_exec(&pl, row, d + h * dist, (int)stride);
    for (j = 0; j < n; j++) d[h * dist + (int64_t)j * stride] = row[j];
  }
  odft_free(&pl);
  free(row);
}

/* ------------------------------------------------------------- tunables */

static int inv_log(int n) { return (n == LOG0) ? 0 : (1 << n); }           /* :37-39 */
static int floor_log(int n) {                                              /* :41-52 */
  if (n == 0) return LOG0;
  int count = -1;
  while (n > 0) { count++; n >>= 1; }
  return count;
}

#define MAX_GRID 128
/* params_range_setup, offt-compute.c:2998-3093: the value grid of every tunable.
 * lists[i*MAX_GRID + c] = c-th value of parameter i, sizes[i] = how many. */
void oracle_params_range(int Nx, int Ny, int Nz, int p, int *lists, int *sizes) {
  int i, c;
  for (i = 0; i < PARAM_COUNT; i++) {
    int *L = lists + i * MAX_GRID;
    if (i == P_P1) {
      int p_u = imin(imin(Nx, Ny), p);
      int p_l = imax(imax(p / Nz, p / Ny), 1);
      int pi1;
      c = 0;
      for (pi1 = p_l; pi1 <= p_u; pi1++) if (p % pi1 == 0 && c < MAX_GRID) L[c++] = pi1;
      sizes[i] = c;
    } else if (i == P_W1 || i == P_W2 || i == P_Ry) {
      for (c = 0; c < 11; c++) L[c] = c;
      sizes[i] = 11;
    } else if (i == P_V) {
      for (c = 0; c < 4; c++) L[c] = c;
      sizes[i] = 4;
    } else if (i == P_S) {
      L[0] = 0; L[1] = 1; sizes[i] = 2;
    } else {
      int v_max = -1, zero = 0;
      switch (i) {
        case P_T1: case P_Px1: case P_Ux1: case P_Px2: v_max = Nx; break;
        case P_Py1: case P_Uy2: v_max = Ny; break;
        case P_Uz1: case P_T2: case P_Pz2: case P_Uz2: v_max = Nz; break;
        case P_Fz: case P_FP1: v_max = Nx * Ny; zero = 1; break;
        case P_Fy1: case P_FU1: case P_Fy2: case P_FP2: v_max = Nx * Nz; zero = 1; break;
        case P_FU2: case P_Fx: v_max = Ny * Nz; zero = 1; break;
      }
      int l = floor_log(v_max), cc;
      c = 0;
      if (zero) L[c++] = 0;
      for (cc = 0; cc < l + 1; cc++) L[c++] = inv_log(cc);
      if (inv_log(l) < v_max) L[c++] = v_max;
      sizes[i] = c;
    }
  }
}

static int grid_floor(const int *lists, const int *sizes, int i, int raw) {   /* :3096-3109 */
  int j;
  for (j = sizes[i] - 1; j >= 0; j--) if (lists[i * MAX_GRID + j] <= raw) return lists[i * MAX_GRID + j];
  return raw;
}

static int isqrt_trunc(int n) { return (int)sqrt((double)n); }

/* params_set_default, offt-compute.c:3127-3225 */
void oracle_params_default(int Nx, int Ny, int Nz, int p, int is_W0, int is_notest, int *v) {
  static int lists[PARAM_COUNT * MAX_GRID];
  int sizes[PARAM_COUNT];
  oracle_params_range(Nx, Ny, Nz, p, lists, sizes);
#define GRID(i) v[i] = grid_floor(lists, sizes, i, v[i])
  v[P_P1] = isqrt_trunc(p); GRID(P_P1);
  int p2 = p / v[P_P1];
  int M1 = (Nx + v[P_P1] - 1) / v[P_P1], M2 = (Ny + p2 - 1) / p2, M3 = (Nz + p2 - 1) / p2, M4 = (Ny + v[P_P1] - 1) / v[P_P1];
  v[P_T1] = imax(M1 / 16, 1); GRID(P_T1);
  v[P_W1] = imin(imax(2, 0), (M1 + v[P_T1] - 1) / v[P_T1]); GRID(P_W1);
  int P1_xy = 8192 / Nz;
  v[P_Px1] = imin(imax(isqrt_trunc(P1_xy), 1), v[P_T1]); GRID(P_Px1);
  v[P_Py1] = imin(imax(P1_xy / v[P_Px1], 1), M2); GRID(P_Py1);
  v[P_Fz] = imin(imax(p2 / 2, 0), v[P_T1] * M2); GRID(P_Fz);
  v[P_FP1] = imin(imax(v[P_Fz], 0), v[P_T1] / v[P_Px1] * M2 / v[P_Py1]); GRID(P_FP1);
  int U1_xz = 8192 / Ny;
  v[P_Ux1] = imin(imax(isqrt_trunc(U1_xz), 1), v[P_T1]); GRID(P_Ux1);
  v[P_Uz1] = imin(imax(U1_xz / v[P_Ux1], 1), M3); GRID(P_Uz1);
  v[P_FU1] = imin(imax(v[P_Fz], 0), v[P_T1] / v[P_Ux1] * M3 / v[P_Uz1]); GRID(P_FU1);
  v[P_Fy1] = imin(imax(v[P_Fz], 0), v[P_T1] * M3); GRID(P_Fy1);
  v[P_Ry] = 5;
  v[P_T2] = imax(M3 / 16, 1); GRID(P_T2);
  v[P_W2] = imin(imax(2, 0), (M3 + v[P_T2] - 1) / v[P_T2]); GRID(P_W2);
  int P2_xz = 8192 / Ny;
  v[P_Pz2] = imin(imax(isqrt_trunc(P2_xz), 1), v[P_T2]); GRID(P_Pz2);
  v[P_Px2] = imin(imax(P2_xz / v[P_Pz2], 1), M1); GRID(P_Px2);
  v[P_Fy2] = imin(imax(v[P_P1] / 2, 0), v[P_T2] * M1); GRID(P_Fy2);
  v[P_FP2] = imin(imax(v[P_Fy2], 0), M1 / v[P_Px2] * v[P_T2] / v[P_Pz2]); GRID(P_FP2);
  int U2_yz = 8192 / Nx;
  v[P_Uz2] = imin(imax(isqrt_trunc(U2_yz), 1), v[P_T2]); GRID(P_Uz2);
  v[P_Uy2] = imin(imax(U2_yz / v[P_Uz2], 1), M4); GRID(P_Uy2);
  v[P_FU2] = imin(imax(v[P_FP2], 0), M4 / v[P_Uy2] * v[P_T2] / v[P_Uz2]); GRID(P_FU2);
  v[P_Fx] = imin(imax(v[P_FP2], 0), v[P_T2] * M4); GRID(P_Fx);
#undef GRID
  v[P_V] = 0; v[P_S] = 0;
  if (is_W0) {
    v[P_W1] = v[P_W2] = 0;
    v[P_Fz] = v[P_FP1] = v[P_FU1] = v[P_Fy1] = v[P_Fy2] = v[P_FP2] = v[P_FU2] = v[P_Fx] = 0;
  }
  if (is_notest) v[P_Fz] = v[P_FP1] = v[P_FU1] = v[P_Fy1] = v[P_Fy2] = v[P_FP2] = v[P_FU2] = v[P_Fx] = 0;
}

/* is_infeasible_point, offt-tuning.c:144-226 (AVOID_TILE is off in the hop build).
 * Returns 0 if feasible, else 1 and *bad = index of the offending parameter. */
int oracle_is_infeasible(int Nx, int Ny, int Nz, int p, const int *v, int *bad) {
  *bad = -1;
  int p1 = v[P_P1];
  if (p1 < 1 || p % p1 != 0) { *bad = P_P1; return 1; }
  int p2 = p / p1;
  int M1 = (Nx + p1 - 1) / p1, M2 = (Ny + p2 - 1) / p2, M3 = (Nz + p2 - 1) / p2, M4 = (Ny + p1 - 1) / p1;
#define BAD(i) do { *bad = (i); return 1; } while (0)
  if (p1 > Nx || p1 > Ny || p2 > Ny || p2 > Nz) BAD(P_P1);
  if (v[P_T1] < 1 || M1 < v[P_T1]) BAD(P_T1);
  if ((M1 + v[P_T1] - 1) / v[P_T1] < v[P_W1] || (M1 == v[P_T1] && v[P_W1] > 0) ||
      (v[P_T1] * M2 * (M3 * p2) > BUFFER_SIZE_LIMIT / (v[P_W1] + 1) / 2 / 2)) BAD(P_W1);
  if (v[P_Px1] < 1 || v[P_T1] < v[P_Px1]) BAD(P_Px1);
  if (v[P_Py1] < 1 || M2 < v[P_Py1]) BAD(P_Py1);
  if (v[P_Fz] < 0 || v[P_T1] * M2 < v[P_Fz]) BAD(P_Fz);
  if (v[P_FP1] < 0 || v[P_T1] / v[P_Px1] * M2 / v[P_Py1] < v[P_FP1]) BAD(P_FP1);
  if (v[P_Ux1] < 1 || v[P_T1] < v[P_Ux1]) BAD(P_Ux1);
  if (v[P_Uz1] < 1 || M3 < v[P_Uz1]) BAD(P_Uz1);
  if (v[P_Fy1] < 0 || v[P_T1] * M3 < v[P_Fy1]) BAD(P_Fy1);
  if (v[P_FU1] < 0 || v[P_T1] / v[P_Ux1] * M3 / v[P_Uz1] < v[P_FU1]) BAD(P_FU1);
  if (v[P_T2] < 1 || M3 < v[P_T2]) BAD(P_T2);
  if ((M3 + v[P_T2] - 1) / v[P_T2] < v[P_W2] || (M3 == v[P_T2] && v[P_W2] > 0) ||
      (M1 * (M4 * p1) * v[P_T2] > BUFFER_SIZE_LIMIT / (v[P_W2] + 1) / 2 / 2)) BAD(P_W2);
  if (v[P_Fy2] < 0 || v[P_T2] * M1 < v[P_Fy2]) BAD(P_Fy2);
  if (v[P_Px2] < 1 || M1 < v[P_Px2]) BAD(P_Px2);
  if (v[P_Pz2] < 1 || v[P_T2] < v[P_Pz2]) BAD(P_Pz2);
  if (v[P_FP2] < 0) BAD(P_FP2);
  if (v[P_Uy2] < 1 || M4 < v[P_Uy2]) BAD(P_Uy2);
  if (v[P_Uz2] < 1 || v[P_T2] < v[P_Uz2]) BAD(P_Uz2);
  if (v[P_Fx] < 0 || v[P_T2] * M4 < v[P_Fx]) BAD(P_Fx);
  if (v[P_V] < 0 || v[P_V] > 3) BAD(P_V);
  if (v[P_S] < 0 || v[P_S] > 1) BAD(P_S);
#undef BAD
  return 0;
}

/* the ADJUST_POINT fix-ups of params_convert, offt-tuning.c:90-118 */
void oracle_params_adjust(int Nx, int Ny, int Nz, int p, int is_oned, int *v) {
  if (is_oned && v[P_P1] == 1) {
    v[P_Ry] = 10; v[P_T2] = 1; v[P_W2] = 0;
    v[P_Fy2] = v[P_FP2] = v[P_FU2] = v[P_Fx] = 0;
    v[P_Pz2] = v[P_Px2] = v[P_Uz2] = v[P_Uy2] = 1;
  }
  if (is_oned && v[P_P1] == p) {
    v[P_Ry] = 0; v[P_T1] = 1; v[P_W1] = 0;
    v[P_Fz] = v[P_FP1] = v[P_FU1] = v[P_Fy1] = 0;
    v[P_Px1] = v[P_Py1] = v[P_Ux1] = v[P_Uz1] = 1;
  }
  if (v[P_W1] == 0) v[P_Fz] = v[P_FP1] = v[P_Fy1] = v[P_FU1] = 0;
  if (v[P_W2] == 0) v[P_Fy2] = v[P_FP2] = v[P_Fx] = v[P_FU2] = 0;
  int p1 = v[P_P1], p2 = p / p1;
  if (Ny % p2 == 0 && Nz % p2 == 0) v[P_V] &= 1;
  if (Nx % p1 == 0 && Ny % p1 == 0) v[P_V] &= 2;
}
