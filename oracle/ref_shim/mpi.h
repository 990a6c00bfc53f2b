/* oracle/ref_shim: TEST INFRASTRUCTURE.  Stand-in for <mpi.h> (NaluEnv.h only
 * needs the communicator type to be declared). */
#ifndef NW_REF_SHIM_MPI_H
#define NW_REF_SHIM_MPI_H
typedef int MPI_Comm;
#define MPI_COMM_WORLD 0
#endif
