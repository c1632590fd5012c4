/*
 * TEST INFRASTRUCTURE — not product code.
 * Link stubs for the FFTW-MPI comparator path of run-fft.c (-a 1), which is
 * out of scope.  Calling them aborts.
 */
#ifndef OFFT_ORACLE_SHIM_FFTW3_MPI_H
#define OFFT_ORACLE_SHIM_FFTW3_MPI_H
#include "fftw3.h"
#include <mpi.h>
#ifdef __cplusplus
extern "C" {
#endif
#define FFTW_MPI_TRANSPOSED_OUT (1U << 30)
void fftw_mpi_init(void);
void fftw_mpi_cleanup(void);
fftw_plan fftw_mpi_plan_dft_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, fftw_complex *in,
                               fftw_complex *out, MPI_Comm comm, int sign, unsigned flags);
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t n0, ptrdiff_t n1, ptrdiff_t n2, double *in,
                                   fftw_complex *out, MPI_Comm comm, unsigned flags);
#ifdef __cplusplus
}
#endif
#endif
