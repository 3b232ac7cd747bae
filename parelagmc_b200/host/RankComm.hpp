// RankComm.hpp -- the ranks that share a sample budget, for the host layer.
//
// The reference runs every realisation on every MPI rank of one communicator (src/MLMC_Manager.cpp:103-179, the mesh is
// distributed); here a rank is one GPU with the whole hierarchy, the ranks own disjoint slices of every level's
// realisations, and InitRun ends with ONE all-reduce of the per-level sums through the C ABI (pmc_allreduce_sums: NCCL
// over NVLink).  With MPI (-DPARELAGMC_B200_WITH_PARELAG) the NCCL id travels by MPI_Bcast; without MPI (this image) the
// ranks are plain processes started by any launcher that sets PMC_WORLD_SIZE / PMC_RANK (or torchrun's WORLD_SIZE /
// RANK), and the id travels through a file (PMC_ID_FILE, default /tmp/pmc_nccl_id.<MASTER_PORT>).
#pragma once
#include <cstdint>
#include "../../include/pmc_b200.h"
#include "shim.hpp"

namespace parelagmc {
/// Collective over `comm`: create the library's NCCL communicator on handle h (no-op for one rank).
void InitDeviceComm(MPI_Comm comm, pmc_handle h);
/// Contiguous slice [first, first + count) of n realisations owned by `rank` of `world`.
inline void SplitSamples(int n, int rank, int world, int &first, int &count)
{
    const int base = n / world, rem = n % world;
    count = base + (rank < rem ? 1 : 0);
    first = rank * base + (rank < rem ? rank : rem);
}
}  // namespace parelagmc
