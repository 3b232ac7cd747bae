// stubs/mpi.h -- DECLARATIONS ONLY of the handful of MPI names the host layer touches, for the compile-only check of the
// -DPARELAGMC_B200_WITH_PARELAG branch in an image without MPI (make -C parelagmc_b200/host parelag-syntax).  Never linked.
#pragma once
typedef int MPI_Comm;
typedef int MPI_Datatype;
#define MPI_COMM_WORLD 0
#define MPI_COMM_SELF 1
#define MPI_BYTE 1
int MPI_Comm_size(MPI_Comm, int *);
int MPI_Comm_rank(MPI_Comm, int *);
int MPI_Bcast(void *, int, MPI_Datatype, int, MPI_Comm);
int MPI_Init(int *, char ***);
int MPI_Finalize();
