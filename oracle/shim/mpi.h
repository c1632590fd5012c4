/*
 * TEST INFRASTRUCTURE — not product code.
 *
 * Minimal single-node stand-in for the MPI subset that the reference
 * (rchyena/offt: offt-compute.c, offt-tuning.c, run-fft.c) calls, so that the
 * reference's own sources compile UNMODIFIED in a container that has no MPI.
 * Ranks are forked processes; collectives run over one shared mmap region
 * (see shim_mpi.c).  Symbol list taken from SURVEY.md section 8(c).
 */
#ifndef OFFT_ORACLE_SHIM_MPI_H
#define OFFT_ORACLE_SHIM_MPI_H

#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;      /* index into the per-process communicator table */
typedef int MPI_Group;     /* index into the per-process group table */
typedef int MPI_Request;   /* index into the per-process request table */
typedef int MPI_Datatype;  /* extent in bytes */
typedef int MPI_Info;
typedef int MPI_Op;
typedef long MPI_Aint;
typedef struct { int MPI_SOURCE, MPI_TAG, MPI_ERROR; } MPI_Status;

#define MPI_COMM_WORLD 0
#define MPI_COMM_NULL (-1)
#define MPI_SUCCESS 0
#define MPI_INFO_NULL 0
#define MPI_DOUBLE 8
#define MPI_INT 4
#define MPI_MAX 1

int MPI_Init(int *argc, char ***argv);
int MPI_Finalize(void);
int MPI_Comm_size(MPI_Comm comm, int *size);
int MPI_Comm_rank(MPI_Comm comm, int *rank);
int MPI_Comm_group(MPI_Comm comm, MPI_Group *group);
int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup);
int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm);
int MPI_Barrier(MPI_Comm comm);
int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm);
int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm);
int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype stype,
                 void *rbuf, int rcount, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype stype,
                  void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rtype, MPI_Comm comm);
int MPI_Ialltoall(const void *sbuf, int scount, MPI_Datatype stype,
                  void *rbuf, int rcount, MPI_Datatype rtype, MPI_Comm comm, MPI_Request *req);
int MPI_Ialltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype stype,
                   void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rtype,
                   MPI_Comm comm, MPI_Request *req);
int MPI_Wait(MPI_Request *req, MPI_Status *status);
int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status);
double MPI_Wtime(void);
int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr);
int MPI_Free_mem(void *base);
int MPI_Type_struct(int count, int *lens, MPI_Aint *disps, MPI_Datatype *types, MPI_Datatype *newtype);
int MPI_Type_commit(MPI_Datatype *type);

#ifdef __cplusplus
}
#endif
#endif
