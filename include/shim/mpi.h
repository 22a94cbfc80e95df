/* Stand-in for <mpi.h>: serial by default, with an opt-in multi-process mode for one box.
 *
 * This image has no MPI.  The shim lets (a) the C++ host mirror of
 * atrip::Atrip in this repo and the reference's own bench/main.cxx driver, and
 * (b) the oracle build of the reference's unchanged sources
 * (the .cxx files under /root/reference/src/atrip, see oracle/Makefile) compile and run.
 * It is not an MPI implementation.
 *
 *   serial (default)   np = 1: every collective degenerates to a local copy.
 *   ATRIP_SHIM_MPI=1   one process per GPU on one box, started by any launcher that exports
 *                      RANK and WORLD_SIZE (torchrun --no-python, tests/launch helpers):
 *                      MPI_Comm_rank/size report them and MPI_Barrier, MPI_Bcast, MPI_Reduce,
 *                      MPI_Allreduce (SUM / MAX) and MPI_Allgather work across the processes
 *                      through files in ATRIP_SHIM_MPI_DIR (default /dev/shm/atrip_shim_mpi_<MASTER_PORT>):
 *                      collective number s of rank r is the file  s.r ; a rank publishes its
 *                      contribution with an atomic rename, polls for the others' and removes its
 *                      file of collective s-1 (everybody has read it once all files of s exist).
 *                      Every job needs a fresh directory.
 *                      That is all the host mirror of Atrip::run needs from MPI (rank, size, the
 *                      broadcast of the 128-byte NCCL id, barriers) -- slices and the energy sum
 *                      travel over NCCL inside the engine.
 * The point-to-point calls abort in both modes: at np = 1 the reference never reaches them (every
 * slice is SelfSufficient, reference SliceUnion.cxx:138-155), and this repo's host code has none.
 * With a real MPI on the include path this file is simply not used.
 *
 * Convention: a datatype handle IS its size in bytes, so derived types built
 * with MPI_Type_vector(n,1,1,DT) are simply n*DT (the reference uses that for
 * its 56-byte database element, Slice.hpp:223, and its 24-byte tuple,
 * Tuples.cxx:384).  Reductions interpret 8-byte elements as double and 4-byte elements as int.
 */
#ifndef ATRIP_B200_SERIAL_MPI_H
#define ATRIP_B200_SERIAL_MPI_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <unistd.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef int MPI_Comm;
typedef int MPI_Datatype;
typedef int MPI_Request;
typedef int MPI_Op;
typedef long MPI_Aint;
typedef struct {
  int MPI_SOURCE, MPI_TAG, MPI_ERROR;
} MPI_Status;

#define MPI_COMM_WORLD 1
#define MPI_COMM_SELF 2
#define MPI_SUCCESS 0
#define MPI_MAX_PROCESSOR_NAME 256
#define MPI_MAX_ERROR_STRING 256

#define MPI_CHAR 1
#define MPI_BYTE 1
#define MPI_INT 4
#define MPI_DOUBLE 8
#define MPI_UINT64_T 8
#define MPI_DOUBLE_COMPLEX 16

#define MPI_SUM 1
#define MPI_MAX 2

/* process-wide state of the multi-process mode: weak, so that every translation unit and shared
 * object of the process (driver, libatrip.so) shares one collective counter */
struct atrip_shim_mpi_state {
  int init, rank, size;
  long seq;
  char dir[512];
};
__attribute__((weak)) struct atrip_shim_mpi_state atrip_shim_mpi_state_ = {0, 0, 1, 0, {0}};

static inline struct atrip_shim_mpi_state *atrip_shim_mpi(void) {
  struct atrip_shim_mpi_state *s = &atrip_shim_mpi_state_;
  if (!s->init) {
    const char *on = getenv("ATRIP_SHIM_MPI"), *r = getenv("RANK"), *n = getenv("WORLD_SIZE");
    s->init = 1;
    if (on && on[0] == '1' && r && n && atoi(n) > 1) {
      const char *d = getenv("ATRIP_SHIM_MPI_DIR"), *port = getenv("MASTER_PORT");
      s->rank = atoi(r);
      s->size = atoi(n);
      if (d) snprintf(s->dir, sizeof s->dir, "%s", d);
      else snprintf(s->dir, sizeof s->dir, "/dev/shm/atrip_shim_mpi_%s", port ? port : "0");
      char cmd[600];
      snprintf(cmd, sizeof cmd, "mkdir -p '%s'", s->dir);
      if (system(cmd) != 0) {
        fprintf(stderr, "shim mpi.h: cannot create %s\n", s->dir);
        abort();
      }
    }
  }
  return s;
}

/* every rank contributes `bytes` bytes; all[r * bytes ...] receives rank r's (all may be NULL) */
static inline void atrip_shim_mpi_exchange(const void *mine, size_t bytes, void *all) {
  struct atrip_shim_mpi_state *s = atrip_shim_mpi();
  if (s->size == 1) {
    if (all && all != mine) memcpy(all, mine, bytes);
    return;
  }
  char path[640], tmp[660];
  const long q = s->seq++;
  snprintf(path, sizeof path, "%s/%ld.%d", s->dir, q, s->rank);
  snprintf(tmp, sizeof tmp, "%s.tmp", path);
  FILE *f = fopen(tmp, "wb");
  if (!f || (bytes && fwrite(mine, 1, bytes, f) != bytes) || fclose(f) != 0 || rename(tmp, path) != 0) {
    fprintf(stderr, "shim mpi.h: cannot publish %s\n", path);
    abort();
  }
  for (int r = 0; r < s->size; r++) {
    snprintf(path, sizeof path, "%s/%ld.%d", s->dir, q, r);
    long waited = 0;
    while ((f = fopen(path, "rb")) == NULL) {
      usleep(200);
      if (++waited > 3000000) { /* 10 minutes */
        fprintf(stderr, "shim mpi.h: rank %d gave up waiting for rank %d in collective %ld\n", s->rank, r, q);
        abort();
      }
    }
    if (all) {
      if (bytes && fread((char *)all + (size_t)r * bytes, 1, bytes, f) != bytes) {
        fprintf(stderr, "shim mpi.h: short read on %s\n", path);
        abort();
      }
    }
    fclose(f);
  }
  if (q > 0) {
    snprintf(path, sizeof path, "%s/%ld.%d", s->dir, q - 1, s->rank);
    unlink(path);
  }
}

static inline void atrip_shim_mpi_reduce(const void *sbuf, void *rbuf, int n, MPI_Datatype dt, MPI_Op op) {
  struct atrip_shim_mpi_state *s = atrip_shim_mpi();
  const size_t bytes = (size_t)n * (size_t)dt;
  if (s->size == 1) {
    if (rbuf != sbuf) memcpy(rbuf, sbuf, bytes);
    return;
  }
  if (dt != 8 && dt != 4) {
    fprintf(stderr, "shim mpi.h: reductions are implemented for double and int only\n");
    abort();
  }
  char *all = (char *)malloc(bytes * (size_t)s->size);
  atrip_shim_mpi_exchange(sbuf, bytes, all);
  for (int i = 0; i < n; i++) { /* rank order: every rank computes the same bits */
    if (dt == 8) {
      double acc = ((double *)all)[i];
      for (int r = 1; r < s->size; r++) {
        const double v = ((double *)(all + (size_t)r * bytes))[i];
        acc = op == MPI_MAX ? (v > acc ? v : acc) : acc + v;
      }
      ((double *)rbuf)[i] = acc;
    } else {
      int acc = ((int *)all)[i];
      for (int r = 1; r < s->size; r++) {
        const int v = ((int *)(all + (size_t)r * bytes))[i];
        acc = op == MPI_MAX ? (v > acc ? v : acc) : acc + v;
      }
      ((int *)rbuf)[i] = acc;
    }
  }
  free(all);
}

static inline int MPI_Init(int *argc, char ***argv) {
  (void)argc;
  (void)argv;
  atrip_shim_mpi();
  return MPI_SUCCESS;
}
static inline int MPI_Finalize(void) {
  struct atrip_shim_mpi_state *s = atrip_shim_mpi();
  /* a last barrier; its (empty) files stay behind, a peer may still be polling for them: give every
   * job a fresh ATRIP_SHIM_MPI_DIR */
  if (s->size > 1) atrip_shim_mpi_exchange("", 0, NULL);
  return MPI_SUCCESS;
}
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) {
  *rank = c == MPI_COMM_SELF ? 0 : atrip_shim_mpi()->rank;
  return MPI_SUCCESS;
}
static inline int MPI_Comm_size(MPI_Comm c, int *size) {
  *size = c == MPI_COMM_SELF ? 1 : atrip_shim_mpi()->size;
  return MPI_SUCCESS;
}
static inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *out) {
  (void)color;
  (void)key;
  *out = c;
  return MPI_SUCCESS;
}
static inline int MPI_Barrier(MPI_Comm c) {
  if (c != MPI_COMM_SELF) atrip_shim_mpi_exchange("", 0, NULL);
  return MPI_SUCCESS;
}
static inline int MPI_Bcast(void *b, int n, MPI_Datatype dt, int root, MPI_Comm c) {
  struct atrip_shim_mpi_state *s = atrip_shim_mpi();
  if (c == MPI_COMM_SELF || s->size == 1) return MPI_SUCCESS;
  const size_t bytes = (size_t)n * (size_t)dt;
  char *all = (char *)malloc(bytes * (size_t)s->size);
  atrip_shim_mpi_exchange(b, bytes, all);
  memcpy(b, all + (size_t)root * bytes, bytes);
  free(all);
  return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype dt,
                             MPI_Op op, int root, MPI_Comm c) {
  (void)root; /* every rank receives the result */
  if (c == MPI_COMM_SELF) memcpy(r, s, (size_t)n * (size_t)dt);
  else atrip_shim_mpi_reduce(s, r, n, dt, op);
  return MPI_SUCCESS;
}
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype dt,
                                MPI_Op op, MPI_Comm c) {
  if (c == MPI_COMM_SELF) memcpy(r, s, (size_t)n * (size_t)dt);
  else atrip_shim_mpi_reduce(s, r, n, dt, op);
  return MPI_SUCCESS;
}
static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype sdt, void *r,
                                int rn, MPI_Datatype rdt, MPI_Comm c) {
  (void)rn;
  (void)rdt;
  if (c == MPI_COMM_SELF) memcpy(r, s, (size_t)sn * (size_t)sdt);
  else atrip_shim_mpi_exchange(s, (size_t)sn * (size_t)sdt, r);
  return MPI_SUCCESS;
}
static inline int MPI_Allgatherv(const void *s, int sn, MPI_Datatype sdt, void *r,
                                 const int *rn, const int *displs,
                                 MPI_Datatype rdt, MPI_Comm c) {
  (void)rn;
  (void)c;
  memcpy((char *)r + (size_t)displs[0] * (size_t)rdt, s, (size_t)sn * (size_t)sdt);
  return MPI_SUCCESS;
}
static inline int MPI_Scatter(const void *s, int sn, MPI_Datatype sdt, void *r,
                              int rn, MPI_Datatype rdt, int root, MPI_Comm c) {
  (void)rn;
  (void)rdt;
  (void)root;
  (void)c;
  memcpy(r, s, (size_t)sn * (size_t)sdt);
  return MPI_SUCCESS;
}
static inline int MPI_Get_processor_name(char *name, int *len) {
  strcpy(name, "localhost");
  *len = 9;
  return MPI_SUCCESS;
}
static inline int MPI_Type_vector(int n, int bl, int stride, MPI_Datatype dt,
                                  MPI_Datatype *out) {
  (void)bl;
  (void)stride;
  *out = n * dt;
  return MPI_SUCCESS;
}
static inline int MPI_Type_create_struct(int n, const int *lens,
                                         const MPI_Aint *disp,
                                         const MPI_Datatype *types,
                                         MPI_Datatype *out) {
  long end = 0;
  for (int i = 0; i < n; i++) {
    long e = disp[i] + (long)lens[i] * types[i];
    if (e > end) end = e;
  }
  *out = (MPI_Datatype)end;
  return MPI_SUCCESS;
}
static inline int MPI_Type_commit(MPI_Datatype *dt) {
  (void)dt;
  return MPI_SUCCESS;
}
static inline int MPI_Type_free(MPI_Datatype *dt) {
  (void)dt;
  return MPI_SUCCESS;
}
static inline int MPI_Get_address(const void *p, MPI_Aint *a) {
  *a = (MPI_Aint)(intptr_t)p;
  return MPI_SUCCESS;
}
static inline MPI_Aint MPI_Aint_diff(MPI_Aint a, MPI_Aint b) { return a - b; }
static inline int MPI_Error_string(int code, char *s, int *len) {
  *len = snprintf(s, MPI_MAX_ERROR_STRING, "serial-mpi error %d", code);
  return MPI_SUCCESS;
}
static inline int MPI_Wait(MPI_Request *r, MPI_Status *s) {
  (void)r;
  if (s) s->MPI_SOURCE = s->MPI_TAG = s->MPI_ERROR = 0;
  return MPI_SUCCESS;
}
static inline int MPI_Waitall(int n, MPI_Request *r, MPI_Status *s) {
  (void)n;
  (void)r;
  (void)s;
  return MPI_SUCCESS;
}
static inline int MPI_Request_free(MPI_Request *r) {
  (void)r;
  return MPI_SUCCESS;
}
static inline int serial_mpi_no_p2p_(const char *what) {
  fprintf(stderr, "shim mpi.h: %s is not supported (point-to-point calls are not part of the shim)\n", what);
  abort();
  return 1;
}
static inline int MPI_Isend(const void *b, int n, MPI_Datatype dt, int dst, int tag,
                            MPI_Comm c, MPI_Request *r) {
  (void)b;
  (void)n;
  (void)dt;
  (void)dst;
  (void)tag;
  (void)c;
  (void)r;
  return serial_mpi_no_p2p_("MPI_Isend");
}
static inline int MPI_Irecv(void *b, int n, MPI_Datatype dt, int src, int tag,
                            MPI_Comm c, MPI_Request *r) {
  (void)b;
  (void)n;
  (void)dt;
  (void)src;
  (void)tag;
  (void)c;
  (void)r;
  return serial_mpi_no_p2p_("MPI_Irecv");
}

#ifdef __cplusplus
}
#endif
#endif
