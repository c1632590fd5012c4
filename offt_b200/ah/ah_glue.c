/*
 * Drives an Active Harmony tuning session the way the reference's ah_tuning does (offt-tuning.c:773-854, 893-913,
 * 988-1019): a session named "fft" with one integer variable V00..V23 per tunable, each ranging over the INDICES of
 * that tunable's value grid; the strategy plug-in the caller asked for; the initial simplex handed over through the
 * SHSONG_USER_VERTEX_FILE key the reference's patched nm.so reads (strategies/nm.c:369-396); a server started on
 * demand next to this library if none answers; then fetch / report until the caller stops, and the best point.
 *
 * Built by offt_b200/ah/Makefile against the UNMODIFIED Harmony client sources of the reference tree into
 * _root/lib/libofft_ah.so, which libofft_b200.so dlopens when it exists (csrc/tune.cu).  Plain C ABI, one session at a
 * time, rank 0 only - the other ranks follow the decisions through the collective agreement of the search loop.
 */
#include <errno.h>
#include <fcntl.h>
#include <signal.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/types.h>
#include <sys/wait.h>
#include <unistd.h>

#include "hclient.h"
#include "hsession.h"

#define AH_MAX_VARS 64

static hdesc_t *g_desc = NULL;
static long g_var[AH_MAX_VARS];
static int g_nvars = 0;
static pid_t g_server = 0;
static char g_error[512] = "";
#define DBG(...) do { if (getenv("OFFTB_AH_DEBUG")) { fprintf(stderr, "offtb_ah: " __VA_ARGS__); fprintf(stderr, "\n"); } } while (0)

const char *offtb_ah_error(void) { return g_error; }

/* fork + exec of <root>/bin/hserver with its output silenced (the reference's launch_silent, offt-tuning.c:37-75) */
static pid_t start_server(const char *root) {
  char prog[1024];
  snprintf(prog, sizeof(prog), "%s/bin/hserver", root);
  if (access(prog, X_OK) != 0) { snprintf(g_error, sizeof(g_error), "%s is not there", prog); return -1; }
  pid_t pid = fork();
  if (pid == 0) {
    int fd = open("/dev/null", O_WRONLY);
    if (fd >= 0) { dup2(fd, STDOUT_FILENO); dup2(fd, STDERR_FILENO); close(fd); }
    char *argv[2] = {prog, NULL};
    execv(prog, argv);
    _exit(126);
  }
  if (pid < 0) snprintf(g_error, sizeof(g_error), "fork: %s", strerror(errno));
  return pid;
}

/*
 * root: directory holding bin/hserver and libexec/ (offt_b200/ah/_root); sizes[i]: grid points of tunable i;
 * strategy: 0 nm.so, 1 pro.so, 2 random.so, 3 brute.so (offt-tuning.c:787-792); vertex_file: initial simplex or NULL;
 * port: TCP port of the server on localhost (0: the library's default).  Returns 0, or -1 with offtb_ah_error() set.
 */
int offtb_ah_open(const char *root, int nvars, const int *sizes, int strategy, const char *vertex_file, int port) {
  static const char *plugin[4] = {"nm.so", "pro.so", "random.so", "brute.so"};
  hsession_t sess;
  const char *err;
  char name[8], portstr[16];
  int i, tries;

  g_error[0] = 0;
  if (g_desc) { snprintf(g_error, sizeof(g_error), "a session is already open"); return -1; }
  if (nvars < 1 || nvars > AH_MAX_VARS) { snprintf(g_error, sizeof(g_error), "bad variable count %d", nvars); return -1; }
  if (port > 0) { snprintf(portstr, sizeof(portstr), "%d", port); setenv("HARMONY_S_PORT", portstr, 1); }
  setenv("HARMONY_S_HOST", "localhost", 0);

  if (hsession_init(&sess) < 0 || hsession_name(&sess, "fft") < 0) { snprintf(g_error, sizeof(g_error), "could not create the session"); return -1; }
  for (i = 0; i < nvars; i++) {
    snprintf(name, sizeof(name), "V%02d", i);
    if (hsession_int(&sess, name, 0, sizes[i] - 1, 1) < 0) { snprintf(g_error, sizeof(g_error), "could not declare %s", name); return -1; }
  }
  hsession_strategy(&sess, plugin[strategy < 0 || strategy > 3 ? 0 : strategy]);
  if (vertex_file) hsession_cfg(&sess, "SHSONG_USER_VERTEX_FILE", vertex_file);

  DBG("launching the session");
  err = hsession_launch(&sess, NULL, 0);
  DBG("first launch: %s", err ? err : "ok");
  if (err) {                                   /* nobody answers: start our own server and try again */
    g_server = start_server(root);
    if (g_server <= 0) return -1;
    for (tries = 0; tries < 10 && err; tries++) {
      usleep(300000);
      err = hsession_launch(&sess, NULL, 0);
    }
    if (err) {
      snprintf(g_error, sizeof(g_error), "could not launch the tuning session: %s", err);
      kill(g_server, SIGKILL); waitpid(g_server, NULL, 0); g_server = 0;
      return -1;
    }
  }
  DBG("session launched, joining");
  g_desc = harmony_init();
  if (!g_desc) { snprintf(g_error, sizeof(g_error), "harmony_init failed"); return -1; }
  g_nvars = nvars;
  for (i = 0; i < nvars; i++) {
    snprintf(name, sizeof(name), "V%02d", i);
    g_var[i] = 0;
    harmony_bind_int(g_desc, name, &g_var[i]);
  }
  DBG("bound %d variables", nvars);
  if (harmony_join(g_desc, NULL, 0, "fft") < 0) {
    snprintf(g_error, sizeof(g_error), "could not join the session: %s", harmony_error_string(g_desc));
    return -1;
  }
  return 0;
}

/* next candidate as grid indices; returns harmony_fetch's code (1 new point, 0 the best point so far, -1 error) */
int offtb_ah_fetch(long *idx) {
  int i, rc;
  if (!g_desc) return -1;
  rc = harmony_fetch(g_desc);
  for (i = 0; i < g_nvars; i++) idx[i] = g_var[i];
  return rc;
}

int offtb_ah_report(double perf) { return g_desc ? harmony_report(g_desc, perf) : -1; }
int offtb_ah_converged(void) { return g_desc ? harmony_converged(g_desc) : -1; }

int offtb_ah_best(long *idx) {
  int i, rc;
  if (!g_desc) return -1;
  rc = harmony_best(g_desc);
  for (i = 0; i < g_nvars; i++) idx[i] = g_var[i];
  return rc;
}

void offtb_ah_close(void) {
  if (g_desc) { harmony_leave(g_desc); harmony_fini(g_desc); g_desc = NULL; }
  if (g_server > 0) { kill(g_server, SIGKILL); waitpid(g_server, NULL, 0); g_server = 0; }   /* offt-tuning.c:1018 */
}
