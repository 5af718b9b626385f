/* ORACLE / TEST INFRASTRUCTURE ONLY.  Comm lives in the lammps.h shim. */
#ifndef ORACLE_SHIM_COMM_H
#define ORACLE_SHIM_COMM_H
#include "lammps.h"
#endif
