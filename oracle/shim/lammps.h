/* ORACLE / TEST INFRASTRUCTURE ONLY -- never linked into the product library.
 *
 * Minimal stand-in for the LAMMPS top-level object so that the reference's
 * solver sources (src/main/gfmd_solver.cpp, src/solvers/gfmd_solver_fft.cpp,
 * src/solvers/gfmd_solver_static.cpp) compile UNCHANGED from where they lie
 * under /root/reference.  Only the members those three files touch exist
 * (gfmd_solver.cpp:67-68,95-123,215-251; gfmd_solver_static.cpp:92-135,177).
 * Memory/Error/Domain/Force are the reference's own src/LAMMPS_STUB classes.
 */
#ifndef ORACLE_SHIM_LAMMPS_H
#define ORACLE_SHIM_LAMMPS_H

#include <cstdio>
#include <cstring>
#include "mpi.h"      /* reference src/main/mpi.h serial stubs */
#include "error.h"    /* reference src/LAMMPS_STUB */
#include "memory.h"
#include "domain.h"
#include "force.h"

namespace LAMMPS_NS {

class Comm {
 public:
  int me, nprocs;
  int procgrid[3];
  Comm() : me(0), nprocs(1) { procgrid[0] = procgrid[1] = procgrid[2] = 1; }
};

class LAMMPS {
 public:
  Memory *memory;
  Error *error;
  Comm *comm;
  Domain *domain;
  Force *force;
  FILE *screen;
  FILE *logfile;
  MPI_Comm world;
  int suffix_enable;
  char *suffix;

  LAMMPS() : screen(NULL), logfile(NULL), world(0), suffix_enable(0),
             suffix(NULL) {
    error = new Error();
    memory = new Memory(error);
    comm = new Comm();
    domain = new Domain();
    force = new Force();
  }
  ~LAMMPS() {
    delete force; delete domain; delete comm; delete memory; delete error;
  }
};

}

#endif
