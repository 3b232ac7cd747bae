#include "RankComm.hpp"

#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <stdexcept>
#include <string>
#include <thread>

namespace parelagmc {

void InitDeviceComm(MPI_Comm comm, pmc_handle h)
{
    int size = 1, rank = 0;
    MPI_Comm_size(comm, &size);
    MPI_Comm_rank(comm, &rank);
    if (size <= 1) return;
    unsigned char id[PMC_COMM_ID_BYTES];
#ifdef PARELAGMC_B200_WITH_PARELAG
    if (rank == 0 && pmc_comm_unique_id(id) != PMC_OK) throw std::runtime_error(std::string("pmc_comm_unique_id: ") + pmc_last_error(nullptr));
    MPI_Bcast(id, PMC_COMM_ID_BYTES, MPI_BYTE, 0, comm);
#else
    const char *f = std::getenv("PMC_ID_FILE");
    const char *port = std::getenv("MASTER_PORT");
    const std::string path = f && *f ? f : std::string("/tmp/pmc_nccl_id.") + (port ? port : "0");
    if (rank == 0) {
        if (pmc_comm_unique_id(id) != PMC_OK) throw std::runtime_error(std::string("pmc_comm_unique_id: ") + pmc_last_error(nullptr));
        const std::string tmp = path + ".tmp";
        FILE *fp = std::fopen(tmp.c_str(), "wb");
        if (!fp || std::fwrite(id, 1, sizeof id, fp) != sizeof id) throw std::runtime_error("cannot write " + tmp);
        std::fclose(fp);
        if (std::rename(tmp.c_str(), path.c_str()) != 0) throw std::runtime_error("cannot publish " + path);
    } else {
        bool ok = false;
        for (int t = 0; t < 6000 && !ok; ++t) {   // up to 10 minutes
            if (FILE *fp = std::fopen(path.c_str(), "rb")) {
                ok = std::fread(id, 1, sizeof id, fp) == sizeof id;
                std::fclose(fp);
            }
            if (!ok) std::this_thread::sleep_for(std::chrono::milliseconds(100));
        }
        if (!ok) throw std::runtime_error("rank " + std::to_string(rank) + ": no NCCL id at " + path);
    }
#endif
    if (pmc_comm_init(h, size, rank, id) != PMC_OK) throw std::runtime_error(std::string("pmc_comm_init: ") + pmc_last_error(h));
#ifndef PARELAGMC_B200_WITH_PARELAG
    if (rank == 0) std::remove(path.c_str());   // pmc_comm_init is collective: every rank has read the id by now
#endif
}

}  // namespace parelagmc
