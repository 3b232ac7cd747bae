// MC_Manager.hpp -- single-level Monte Carlo manager; interface of /root/reference/src/MC_Manager.hpp:28-57.
#pragma once
#include <fstream>
#include <iostream>
#include <string>
#include "MLSampler.hpp"
#include "PhysicalMLSolver.hpp"

namespace parelagmc {
class MC_Manager {
public:
    MC_Manager(MPI_Comm comm, PhysicalMLSolver &pSolver, MLSampler &sampler, parelag::ParameterList &master_list);
    ~MC_Manager() = default;
    MC_Manager(MC_Manager const &) = delete;
    MC_Manager &operator=(MC_Manager const &) = delete;

    void Run();
    void InitRun(int nsamples);
    void ShowMe(std::ostream &os = std::cout);
    bool wallTime;

    double Estimate() const { return eQ; }
    double VarQ() const { return varQ; }
    int NumSamples() const { return level_nsamples; }

private:
    enum { Q2 = 0, Q = 1, ABSQ = 2, C = 3, NVAR = 4 };
    void computeNSamplesMSE();
    MPI_Comm comm;
    int rank, pid;
    PhysicalMLSolver &pSolver;
    MLSampler &sampler;
    parelag::ParameterList &prob_list;
    double eps2;
    bool auto_eps2;
    const double ratio;
    const std::string file_name;
    const int init_nsamples;
    double ml_estimator_variance, expected_discretization_error2, actualMSE;
    double sums[NVAR];
    double eQ, eABSQ, eC, varQ, M, time_;
    int level_nsamples, level_nsamples_missing;
    std::ofstream logger;
    bool comm_ready = false;   // NCCL communicator of the ranks (RankComm.hpp) created
};
}  // namespace parelagmc
