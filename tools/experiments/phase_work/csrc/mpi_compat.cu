// libofft_b200_mpicompat.so - the handful of MPI and FFTW symbols the reference driver itself
// references (run-fft.c:158-160, 309, 311, 322-340, 421-424, 514), for boxes without MPI and
// FFTW.  Kept out of libofft_b200.so so that a program linked against a real MPI never sees
// two MPI_Init.  Ranks are OS processes, one per GPU, started by offt_b200/bin/offtrun (or by
// torchrun): MPI_Init reads RANK / WORLD_SIZE / LOCAL_RANK, binds the rank to its GPU and
// bootstraps the NCCL world through a rendezvous file.
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <unistd.h>

#define OFFT_NO_MINMAX
#include "mpi.h"
#include "fftw3-mpi.h"
#include "offt_b200.h"

static int env_int(const char *a, const char *b, int dflt) {
  const char *v = getenv(a);
  if (!v && b) v = getenv(b);
  return v ? atoi(v) : dflt;
}

extern "C" {

int MPI_Init(int *argc, char ***argv) {
  (void)argc; (void)argv;
  const int rank = env_int("OFFTB_RANK", "RANK", 0), size = env_int("OFFTB_WORLD_SIZE", "WORLD_SIZE", 1);
  const int local = env_int("OFFTB_LOCAL_RANK", "LOCAL_RANK", rank);
  unsigned char id[OFFTB_UNIQUE_ID_BYTES];
  memset(id, 0, sizeof(id));
  if (size > 1) {
    // rendezvous: rank 0 publishes the NCCL unique id in a file every rank can see
    const char *f = getenv("OFFTB_ID_FILE");
    std::string path = f ? f : std::string("/tmp/offtb_id_") + (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0") + "_" + std::to_string((long)getppid());
    if (rank == 0) {
      if (offtb_get_unique_id(id)) { fprintf(stderr, "MPI_Init (offt_b200 compat): %s\n", offtb_last_error()); exit(-1); }
      const std::string tmp = path + ".tmp";
      FILE *fp = fopen(tmp.c_str(), "wb");
      if (!fp || fwrite(id, 1, sizeof(id), fp) != sizeof(id)) { perror("MPI_Init (offt_b200 compat): rendezvous file"); exit(-1); }
      fclose(fp);
      rename(tmp.c_str(), path.c_str());
    } else {
      FILE *fp = nullptr;
      for (int tries = 0; tries < 6000 && !(fp = fopen(path.c_str(), "rb")); ++tries) std::this_thread::sleep_for(std::chrono::milliseconds(10));
      if (!fp || fread(id, 1, sizeof(id), fp) != sizeof(id)) { fprintf(stderr, "MPI_Init (offt_b200 compat): no rendezvous file %s\n", path.c_str()); exit(-1); }
      fclose(fp);
    }
  }
  if (offtb_world_init(rank, size, local, size > 1 ? id : nullptr)) {
    fprintf(stderr, "MPI_Init (offt_b200 compat): %s\n", offtb_last_error());
    exit(-1);
  }
  if (size > 1) {
    offtb_world_barrier();
    if (rank == 0 && !getenv("OFFTB_ID_FILE")) {
      std::string path = std::string("/tmp/offtb_id_") + (getenv("MASTER_PORT") ? getenv("MASTER_PORT") : "0") + "_" + std::to_string((long)getppid());
      unlink(path.c_str());
    }
  }
  return MPI_SUCCESS;
}

int MPI_Finalize(void) { offtb_world_fin(); return MPI_SUCCESS; }
int MPI_Comm_size(MPI_Comm, int *size) { *size = offtb_world_size(); return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm, int *rank) { *rank = offtb_world_rank(); return MPI_SUCCESS; }
int MPI_Barrier(MPI_Comm) { return offtb_world_barrier() ? 1 : MPI_SUCCESS; }
double MPI_Wtime(void) {
  using namespace std::chrono;
  return duration_cast<duration<double>>(steady_clock::now().time_since_epoch()).count();
}

// run-fft.c's "-a 1" comparator (FFTW-MPI) is not part of this project
static void no_fftw(const char *what) {
  fprintf(stderr, "%s: the FFTW-MPI comparator path (-a 1) is not provided by offt_b200; use -a 0\n", what);
  exit(-1);
}
void fftw_execute(const fftw_plan) { no_fftw("fftw_execute"); }
void fftw_destroy_plan(fftw_plan) {}
void fftw_print_plan(const fftw_plan) {}
void fftw_mpi_init(void) {}
void fftw_mpi_cleanup(void) {}
fftw_plan fftw_mpi_plan_dft_3d(ptrdiff_t, ptrdiff_t, ptrdiff_t, fftw_complex *, fftw_complex *, MPI_Comm, int, unsigned) {
  no_fftw("fftw_mpi_plan_dft_3d");
  return nullptr;
}
fftw_plan fftw_mpi_plan_dft_r2c_3d(ptrdiff_t, ptrdiff_t, ptrdiff_t, double *, fftw_complex *, MPI_Comm, unsigned) {
  no_fftw("fftw_mpi_plan_dft_r2c_3d");
  return nullptr;
}

}  // extern "C"
