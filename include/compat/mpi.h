/*
 * Stand-in for <mpi.h> on boxes without MPI: the six calls the reference driver makes
 * itself (run-fft.c:158-160, 309, 311, 514) plus the type names offt.h mentions.
 * Ranks are OS processes started by `offt_b200/bin/offtrun -n P <program> ...`, one per
 * GPU; MPI_Init reads the launcher's environment, binds the rank to its GPU and
 * bootstraps NCCL (offt_b200/csrc/mpi_compat.cu).  With a real MPI, use its own mpi.h
 * and call offtb_world_init() after MPI_Init.
 */
#ifndef OFFTB_COMPAT_MPI_H
#define OFFTB_COMPAT_MPI_H
#ifdef __cplusplus
extern "C" {
#endif
typedef int MPI_Comm;
typedef int MPI_Group;
#define MPI_COMM_WORLD 0
#define MPI_SUCCESS 0
int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Barrier(MPI_Comm comm);
double MPI_Wtime(void);
#ifdef __cplusplus
}
#endif
#endif
