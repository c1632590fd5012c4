/*
 * TEST INFRASTRUCTURE — not product code.
 *
 * Single-node MPI stand-in: MPI_Init forks OFFT_SHIM_NP-1 children, every rank
 * continues the caller's main().  One MAP_SHARED region mapped before the fork
 * holds (a) a control block with barriers and per-rank mailboxes and (b) one
 * arena per rank from which MPI_Alloc_mem is served, so that peers can read a
 * rank's send buffer directly at the same virtual address.
 *
 * Non-blocking all-to-all is recorded at post time (count arrays snapshotted)
 * and carried out inside MPI_Wait; every rank of a communicator waits on its
 * requests in the same order in the reference (offt-compute.c:3607-3679,
 * 3789-3861), which keeps the deferred exchange collective.  MPI_Test is a
 * no-op, as with the reference's own NOTEST switch (offt-compute.c:3218-3220).
 */
#define _GNU_SOURCE
#include "mpi.h"

#include <sched.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/mman.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <time.h>
#include <unistd.h>

#define MAX_RANKS 64
#define MAX_COMMS 8192   /* creation slots, reused cyclically (tuning makes 2 per trial) */
#define MAX_REQS 64

typedef struct {
  volatile int count;
  volatile int sense;
} shm_barrier_t;

typedef struct {
  const void *volatile sbuf;
  int scounts[MAX_RANKS];
  int sdispls[MAX_RANKS];
} mailbox_t;

typedef struct {
  int nranks;
  shm_barrier_t barriers[MAX_COMMS][MAX_RANKS]; /* [creation slot][leader world rank] */
  mailbox_t mail[MAX_RANKS];
  volatile int abort_flag;
} control_t;

typedef struct {
  int size;
  int rank;      /* my index within members, -1 if not a member */
  int members[MAX_RANKS];
  int slot;      /* creation slot */
  int leader;    /* smallest world rank */
  int sense;     /* local sense for the barrier */
} comm_t;

typedef struct {
  int n;
  int ranks[MAX_RANKS];
} group_t;

typedef struct {
  int active;
  int is_v;
  MPI_Comm comm;
  const void *sbuf;
  void *rbuf;
  int scount, rcount;
  int esize;
  int scounts[MAX_RANKS], sdispls[MAX_RANKS], rcounts[MAX_RANKS], rdispls[MAX_RANKS];
} req_t;

static control_t *g_ctl;
static char *g_region;
static size_t g_region_size;
static char *g_arena;        /* my arena base */
static size_t g_arena_size;
static int g_rank = 0, g_np = 1;
static pid_t g_children[MAX_RANKS];

static comm_t *g_comms;
static int g_ncomms;
static int g_comm_cap;
static int g_created;        /* number of MPI_Comm_create calls so far */
static group_t *g_groups;
static int g_ngroups, g_group_cap;
static req_t g_reqs[MAX_REQS];

/* ---- arena allocator: first-fit free list, metadata kept process-locally ---- */
typedef struct blk_s { size_t off, size; int used; struct blk_s *next; } blk_t;
static blk_t *g_blks;
static char *g_stage;        /* staging area for user buffers outside the region */
static size_t g_stage_size;

static void *arena_alloc(size_t size) {
  size = (size + 255) & ~(size_t)255;
  if (size == 0) size = 256;
  blk_t *b;
  for (b = g_blks; b; b = b->next) {
    if (!b->used && b->size >= size) {
      if (b->size > size) {
        blk_t *n = (blk_t *)malloc(sizeof(blk_t));
        n->off = b->off + size; n->size = b->size - size; n->used = 0; n->next = b->next;
        b->next = n; b->size = size;
      }
      b->used = 1;
      return g_arena + b->off;
    }
  }
  return NULL;
}

static void arena_free(void *p) {
  size_t off = (size_t)((char *)p - g_arena);
  blk_t *b;
  for (b = g_blks; b; b = b->next) if (b->off == off && b->used) { b->used = 0; break; }
  for (b = g_blks; b && b->next; ) {
    if (!b->used && !b->next->used) {
      blk_t *n = b->next; b->size += n->size; b->next = n->next; free(n);
    } else b = b->next;
  }
}

static int in_region(const void *p) {
  return (const char *)p >= g_region && (const char *)p < g_region + g_region_size;
}

static void die(const char *msg) {
  fprintf(stderr, "[shim-mpi rank %d] %s\n", g_rank, msg);
  if (g_ctl) g_ctl->abort_flag = 1;
  _exit(3);
}

static int new_comm(void) {
  if (g_ncomms == g_comm_cap) {
    g_comm_cap = g_comm_cap ? 2 * g_comm_cap : 64;
    g_comms = (comm_t *)realloc(g_comms, sizeof(comm_t) * g_comm_cap);
  }
  return g_ncomms++;
}

static int new_group(void) {
  if (g_ngroups == g_group_cap) {
    g_group_cap = g_group_cap ? 2 * g_group_cap : 64;
    g_groups = (group_t *)realloc(g_groups, sizeof(group_t) * g_group_cap);
  }
  return g_ngroups++;
}

static void barrier_comm(comm_t *c) {
  if (c->size <= 1) return;
  shm_barrier_t *b = &g_ctl->barriers[c->slot][c->leader];
  int my = c->sense ^ 1;
  c->sense = my;
  if (__sync_add_and_fetch(&b->count, 1) == c->size) {
    b->count = 0;
    __sync_synchronize();
    b->sense = my;
  } else {
    unsigned spins = 0;
    while (b->sense != my) {
      if (g_ctl->abort_flag) _exit(3);
      if (++spins > 200) { sched_yield(); spins = 0; }
    }
  }
  __sync_synchronize();
}

int MPI_Init(int *argc, char ***argv) {
  (void)argc; (void)argv;
  const char *s = getenv("OFFT_SHIM_NP");
  g_np = s ? atoi(s) : 1;
  if (g_np < 1 || g_np > MAX_RANKS) { fprintf(stderr, "shim-mpi: bad OFFT_SHIM_NP\n"); exit(2); }
  const char *g = getenv("OFFT_SHIM_SHM_GB");
  double gb = g ? atof(g) : 0.0;
  if (gb <= 0.0) {
    long pages = sysconf(_SC_PHYS_PAGES), psz = sysconf(_SC_PAGE_SIZE);
    gb = 0.70 * (double)pages * (double)psz / (1024.0 * 1024.0 * 1024.0);
  }
  size_t ctl_size = (sizeof(control_t) + 4095) & ~(size_t)4095;
  size_t per = ((size_t)(gb * 1024.0 * 1024.0 * 1024.0) / (size_t)g_np) & ~(size_t)4095;
  g_region_size = ctl_size + per * (size_t)g_np;
  g_region = (char *)mmap(NULL, g_region_size, PROT_READ | PROT_WRITE,
                          MAP_SHARED | MAP_ANONYMOUS | MAP_NORESERVE, -1, 0);
  if (g_region == MAP_FAILED) { perror("shim-mpi: mmap"); exit(2); }
  g_ctl = (control_t *)g_region;
  g_ctl->nranks = g_np;
  fflush(stdout); fflush(stderr);
  int r;
  for (r = 1; r < g_np; r++) {
    pid_t pid = fork();
    if (pid < 0) { perror("shim-mpi: fork"); exit(2); }
    if (pid == 0) { g_rank = r; break; }
    g_children[r] = pid;
  }
  g_arena = g_region + ctl_size + per * (size_t)g_rank;
  g_arena_size = per;
  g_blks = (blk_t *)malloc(sizeof(blk_t));
  g_blks->off = 0; g_blks->size = per; g_blks->used = 0; g_blks->next = NULL;
  /* world communicator = comm 0, world group = group 0 */
  int w = new_comm();
  comm_t *c = &g_comms[w];
  c->size = g_np; c->rank = g_rank; c->slot = 0; c->leader = 0; c->sense = 0;
  for (r = 0; r < g_np; r++) c->members[r] = r;
  int gi = new_group();
  g_groups[gi].n = g_np;
  for (r = 0; r < g_np; r++) g_groups[gi].ranks[r] = r;
  g_created = 1;
  return MPI_SUCCESS;
}

int MPI_Finalize(void) {
  fflush(stdout); fflush(stderr);
  barrier_comm(&g_comms[0]);
  if (g_rank != 0) { fflush(stdout); _exit(0); }
  int r, status, bad = 0;
  for (r = 1; r < g_np; r++) {
    waitpid(g_children[r], &status, 0);
    if (!WIFEXITED(status) || WEXITSTATUS(status) != 0) bad = 1;
  }
  if (bad) { fprintf(stderr, "shim-mpi: a rank failed\n"); exit(3); }
  return MPI_SUCCESS;
}

int MPI_Comm_size(MPI_Comm comm, int *size) { *size = g_comms[comm].size; return MPI_SUCCESS; }
int MPI_Comm_rank(MPI_Comm comm, int *rank) { *rank = g_comms[comm].rank; return MPI_SUCCESS; }

int MPI_Comm_group(MPI_Comm comm, MPI_Group *group) {
  int gi = new_group();
  g_groups[gi].n = g_comms[comm].size;
  memcpy(g_groups[gi].ranks, g_comms[comm].members, sizeof(int) * g_comms[comm].size);
  *group = gi;
  return MPI_SUCCESS;
}

int MPI_Group_incl(MPI_Group group, int n, const int *ranks, MPI_Group *newgroup) {
  int gi = new_group();
  int i;
  g_groups[gi].n = n;
  for (i = 0; i < n; i++) g_groups[gi].ranks[i] = g_groups[group].ranks[ranks[i]];
  *newgroup = gi;
  return MPI_SUCCESS;
}

int MPI_Comm_create(MPI_Comm comm, MPI_Group group, MPI_Comm *newcomm) {
  /* collective over `comm` (always the world in the reference, offt-compute.c:124-125) */
  (void)comm;
  int slot = g_created % MAX_COMMS;
  g_created++;
  group_t *gr = &g_groups[group];
  int ci = new_comm();
  comm_t *c = &g_comms[ci];
  int i;
  c->size = gr->n; c->rank = -1; c->slot = slot; c->leader = gr->ranks[0];
  for (i = 0; i < gr->n; i++) {
    c->members[i] = gr->ranks[i];
    if (gr->ranks[i] == g_rank) c->rank = i;
    if (gr->ranks[i] < c->leader) c->leader = gr->ranks[i];
  }
  c->sense = g_ctl->barriers[slot][c->leader].sense;
  /* every world rank passes here, so nobody can still be spinning on the slot's old use */
  barrier_comm(&g_comms[0]);
  *newcomm = (c->rank < 0) ? MPI_COMM_NULL : ci;
  return MPI_SUCCESS;
}

int MPI_Barrier(MPI_Comm comm) { barrier_comm(&g_comms[comm]); return MPI_SUCCESS; }

static const void *stage_if_needed(const void *buf, size_t bytes) {
  if (in_region(buf)) return buf;
  if (bytes > g_stage_size) {
    if (g_stage) arena_free(g_stage);
    g_stage = (char *)arena_alloc(bytes);
    if (!g_stage) die("staging allocation failed (raise OFFT_SHIM_SHM_GB)");
    g_stage_size = bytes;
  }
  memcpy(g_stage, buf, bytes);
  return g_stage;
}

int MPI_Bcast(void *buf, int count, MPI_Datatype type, int root, MPI_Comm comm) {
  comm_t *c = &g_comms[comm];
  size_t bytes = (size_t)count * (size_t)type;
  if (c->size <= 1) return MPI_SUCCESS;
  if (c->rank == root) g_ctl->mail[g_rank].sbuf = stage_if_needed(buf, bytes);
  barrier_comm(c);
  if (c->rank != root) memcpy(buf, (const void *)g_ctl->mail[c->members[root]].sbuf, bytes);
  barrier_comm(c);
  return MPI_SUCCESS;
}

int MPI_Reduce(const void *sbuf, void *rbuf, int count, MPI_Datatype type, MPI_Op op, int root, MPI_Comm comm) {
  /* only MPI_DOUBLE / MPI_MAX appears (commented out in run-fft.c:355-357) */
  comm_t *c = &g_comms[comm];
  size_t bytes = (size_t)count * (size_t)type;
  (void)op;
  g_ctl->mail[g_rank].sbuf = stage_if_needed(sbuf, bytes);
  barrier_comm(c);
  if (c->rank == root) {
    int i, j;
    double *out = (double *)rbuf;
    for (j = 0; j < count; j++) out[j] = ((const double *)g_ctl->mail[c->members[0]].sbuf)[j];
    for (i = 1; i < c->size; i++)
      for (j = 0; j < count; j++) {
        double v = ((const double *)g_ctl->mail[c->members[i]].sbuf)[j];
        if (v > out[j]) out[j] = v;
      }
  }
  barrier_comm(c);
  return MPI_SUCCESS;
}

static void do_alltoall(comm_t *c, const void *sbuf, const int *scounts, const int *sdispls,
                        void *rbuf, const int *rcounts, const int *rdispls, int esize) {
  int i;
  size_t total = 0;
  mailbox_t *mine = &g_ctl->mail[g_rank];
  for (i = 0; i < c->size; i++) {
    mine->scounts[i] = scounts[i];
    mine->sdispls[i] = sdispls[i];
    size_t end = ((size_t)sdispls[i] + (size_t)scounts[i]) * (size_t)esize;
    if (end > total) total = end;
  }
  mine->sbuf = stage_if_needed(sbuf, total);
  barrier_comm(c);
  for (i = 0; i < c->size; i++) {
    /* start with myself + 1 to spread the readers over the sources */
    int j = (c->rank + i) % c->size;
    const mailbox_t *src = &g_ctl->mail[c->members[j]];
    size_t n = (size_t)src->scounts[c->rank];
    if ((size_t)rcounts[j] < n) n = (size_t)rcounts[j];
    memcpy((char *)rbuf + (size_t)rdispls[j] * (size_t)esize,
           (const char *)src->sbuf + (size_t)src->sdispls[c->rank] * (size_t)esize,
           n * (size_t)esize);
  }
  barrier_comm(c);
}

int MPI_Alltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype stype,
                  void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rtype, MPI_Comm comm) {
  (void)rtype;
  do_alltoall(&g_comms[comm], sbuf, scounts, sdispls, rbuf, rcounts, rdispls, (int)stype);
  return MPI_SUCCESS;
}

int MPI_Alltoall(const void *sbuf, int scount, MPI_Datatype stype,
                 void *rbuf, int rcount, MPI_Datatype rtype, MPI_Comm comm) {
  comm_t *c = &g_comms[comm];
  int sc[MAX_RANKS], sd[MAX_RANKS], rc[MAX_RANKS], rd[MAX_RANKS], i;
  (void)rtype;
  for (i = 0; i < c->size; i++) { sc[i] = scount; sd[i] = i * scount; rc[i] = rcount; rd[i] = i * rcount; }
  do_alltoall(c, sbuf, sc, sd, rbuf, rc, rd, (int)stype);
  return MPI_SUCCESS;
}

static int new_req(void) {
  int i;
  for (i = 0; i < MAX_REQS; i++) if (!g_reqs[i].active) { g_reqs[i].active = 1; return i; }
  die("too many outstanding requests");
  return -1;
}

int MPI_Ialltoall(const void *sbuf, int scount, MPI_Datatype stype,
                  void *rbuf, int rcount, MPI_Datatype rtype, MPI_Comm comm, MPI_Request *req) {
  int r = new_req();
  req_t *q = &g_reqs[r];
  (void)rtype;
  q->is_v = 0; q->comm = comm; q->sbuf = sbuf; q->rbuf = rbuf;
  q->scount = scount; q->rcount = rcount; q->esize = (int)stype;
  *req = r;
  return MPI_SUCCESS;
}

int MPI_Ialltoallv(const void *sbuf, const int *scounts, const int *sdispls, MPI_Datatype stype,
                   void *rbuf, const int *rcounts, const int *rdispls, MPI_Datatype rtype,
                   MPI_Comm comm, MPI_Request *req) {
  int r = new_req();
  req_t *q = &g_reqs[r];
  int n = g_comms[comm].size;
  (void)rtype;
  q->is_v = 1; q->comm = comm; q->sbuf = sbuf; q->rbuf = rbuf; q->esize = (int)stype;
  memcpy(q->scounts, scounts, sizeof(int) * n);
  memcpy(q->sdispls, sdispls, sizeof(int) * n);
  memcpy(q->rcounts, rcounts, sizeof(int) * n);
  memcpy(q->rdispls, rdispls, sizeof(int) * n);
  *req = r;
  return MPI_SUCCESS;
}

int MPI_Wait(MPI_Request *req, MPI_Status *status) {
  req_t *q = &g_reqs[*req];
  (void)status;
  if (!q->active) return MPI_SUCCESS;
  if (q->is_v)
    MPI_Alltoallv(q->sbuf, q->scounts, q->sdispls, q->esize, q->rbuf, q->rcounts, q->rdispls, q->esize, q->comm);
  else
    MPI_Alltoall(q->sbuf, q->scount, q->esize, q->rbuf, q->rcount, q->esize, q->comm);
  q->active = 0;
  return MPI_SUCCESS;
}

int MPI_Test(MPI_Request *req, int *flag, MPI_Status *status) {
  (void)req; (void)status;
  *flag = 0;
  return MPI_SUCCESS;
}

double MPI_Wtime(void) {
  struct timespec ts;
  clock_gettime(CLOCK_MONOTONIC, &ts);
  return (double)ts.tv_sec + 1e-9 * (double)ts.tv_nsec;
}

int MPI_Alloc_mem(MPI_Aint size, MPI_Info info, void *baseptr) {
  (void)info;
  void *p = arena_alloc((size_t)size);
  if (!p) return 1;
  *(void **)baseptr = p;
  return MPI_SUCCESS;
}

int MPI_Free_mem(void *base) { if (base) arena_free(base); return MPI_SUCCESS; }

int MPI_Type_struct(int count, int *lens, MPI_Aint *disps, MPI_Datatype *types, MPI_Datatype *newtype) {
  long ext = 0;
  int i;
  for (i = 0; i < count; i++) {
    long e = disps[i] + (long)lens[i] * (long)types[i];
    if (e > ext) ext = e;
  }
  *newtype = (MPI_Datatype)ext;
  return MPI_SUCCESS;
}

int MPI_Type_commit(MPI_Datatype *type) { (void)type; return MPI_SUCCESS; }
