/* ORACLE / TEST INFRASTRUCTURE ONLY.  Stand-in for LAMMPS' Pointers base
 * class: exposes lmp/memory/error/comm/domain/screen/world to the reference
 * solver classes (GFMDSolver : protected Pointers, gfmd_solver.h:34). */
#ifndef ORACLE_SHIM_POINTERS_H
#define ORACLE_SHIM_POINTERS_H

#include "lammps.h"

namespace LAMMPS_NS {

class Pointers {
 public:
  Pointers(LAMMPS *ptr)
      : lmp(ptr), memory(ptr->memory), error(ptr->error), comm(ptr->comm),
        domain(ptr->domain), force(ptr->force), screen(ptr->screen),
        logfile(ptr->logfile), world(ptr->world) {}
  virtual ~Pointers() {}

 protected:
  LAMMPS *lmp;
  Memory *&memory;
  Error *&error;
  Comm *&comm;
  Domain *&domain;
  Force *&force;
  FILE *&screen;
  FILE *&logfile;
  MPI_Comm &world;
};

}

#endif
