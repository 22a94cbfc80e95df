/* Serial, single-process stand-in for <mpi.h>.
 *
 * This image has no MPI.  The shim lets (a) the C++ host mirror of
 * atrip::Atrip in this repo and the reference's own bench/main.cxx driver, and
 * (b) the oracle build of the reference's unchanged sources
 * (/root/reference/src/atrip/*.cxx, see oracle/Makefile) compile and run at
 * np = 1.  It is not an MPI implementation: every collective degenerates to a
 * local copy and the point-to-point calls abort, because at np = 1 the
 * reference never reaches them (every slice is SelfSufficient, reference
 * SliceUnion.cxx:138-155).  With a real MPI on the include path this file is
 * simply not used.
 *
 * Convention: a datatype handle IS its size in bytes, so derived types built
 * with MPI_Type_vector(n,1,1,DT) are simply n*DT (the reference uses that for
 * its 56-byte database element, Slice.hpp:223, and its 24-byte tuple,
 * Tuples.cxx:384).
 */
#ifndef ATRIP_B200_SERIAL_MPI_H
#define ATRIP_B200_SERIAL_MPI_H

#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

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

static inline int MPI_Init(int *argc, char ***argv) {
  (void)argc;
  (void)argv;
  return MPI_SUCCESS;
}
static inline int MPI_Finalize(void) { return MPI_SUCCESS; }
static inline int MPI_Comm_rank(MPI_Comm c, int *rank) {
  (void)c;
  *rank = 0;
  return MPI_SUCCESS;
}
static inline int MPI_Comm_size(MPI_Comm c, int *size) {
  (void)c;
  *size = 1;
  return MPI_SUCCESS;
}
static inline int MPI_Comm_split(MPI_Comm c, int color, int key, MPI_Comm *out) {
  (void)color;
  (void)key;
  *out = c;
  return MPI_SUCCESS;
}
static inline int MPI_Barrier(MPI_Comm c) {
  (void)c;
  return MPI_SUCCESS;
}
static inline int MPI_Bcast(void *b, int n, MPI_Datatype dt, int root, MPI_Comm c) {
  (void)b;
  (void)n;
  (void)dt;
  (void)root;
  (void)c;
  return MPI_SUCCESS;
}
static inline int MPI_Reduce(const void *s, void *r, int n, MPI_Datatype dt,
                             MPI_Op op, int root, MPI_Comm c) {
  (void)op;
  (void)root;
  (void)c;
  memcpy(r, s, (size_t)n * (size_t)dt);
  return MPI_SUCCESS;
}
static inline int MPI_Allreduce(const void *s, void *r, int n, MPI_Datatype dt,
                                MPI_Op op, MPI_Comm c) {
  (void)op;
  (void)c;
  memcpy(r, s, (size_t)n * (size_t)dt);
  return MPI_SUCCESS;
}
static inline int MPI_Allgather(const void *s, int sn, MPI_Datatype sdt, void *r,
                                int rn, MPI_Datatype rdt, MPI_Comm c) {
  (void)rn;
  (void)rdt;
  (void)c;
  memcpy(r, s, (size_t)sn * (size_t)sdt);
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
  fprintf(stderr, "serial mpi.h: %s reached at np=1 -- not supported\n", what);
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
