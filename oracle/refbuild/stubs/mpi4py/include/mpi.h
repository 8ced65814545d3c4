/* Serial stand-in for <mpi.h>: the reference only needs MPI_VERSION to be
 * defined (base/PyNucleus_base/mpi-compat.h:13-23); no MPI symbol is called
 * from C.  Test infrastructure only (oracle build), never shipped. */
#ifndef PNB_FAKE_MPI_H
#define PNB_FAKE_MPI_H
#define MPI_VERSION 4
#endif
