// MLMC_Manager.hpp -- multilevel Monte Carlo manager; constructor, Run, InitRun, ShowMe and wallTime as in
// /root/reference/src/MLMC_Manager.hpp:30-61.  When the sampler/solver are the B200 classes of this directory the
// inner sample loop (src/MLMC_Manager.cpp:110-175) is one batched device call per level; with any other
// MLSampler/PhysicalMLSolver the reference's per-sample loop runs unchanged.
#pragma once
#include <fstream>
#include <iostream>
#include <string>
#include <vector>
#include "MLSampler.hpp"
#include "PhysicalMLSolver.hpp"

struct pmc_context_s;

namespace parelagmc {
class MLMC_Manager {
public:
    MLMC_Manager(MPI_Comm comm, const int nlevels, PhysicalMLSolver &pSolver, MLSampler &sampler,
                 parelag::ParameterList &master_list);
    ~MLMC_Manager();
    MLMC_Manager(MLMC_Manager const &) = delete;
    MLMC_Manager &operator=(MLMC_Manager const &) = delete;

    /// Run MLMC simulation
    void Run();
    /// Run level_nsamples_init[ilevel] more samples on every level and update the statistics
    void InitRun(std::vector<int> &level_nsamples_init);
    /// Print MLMC estimators, variances, etc
    void ShowMe(std::ostream &os = std::cout);
    /// If true use wall time as cost, else the number of dofs
    bool wallTime;

    // read access for drivers/tests
    double Estimate() const;
    const std::vector<double> &Sums() const { return sums; }
    const std::vector<int> &NumSamples() const { return level_nsamples; }
    double EstimatorVariance() const { return ml_estimator_variance; }

private:
    enum { Y2 = 0, Y = 1, ABSY = 2, Q2 = 3, Q = 4, ABSQ = 5, C = 6, Y3 = 7, Y4 = 8, NVAR = 9 };
    void computeNSamplesMSE();
    void accumulate(int ilevel, double y, double q, double c);

    MPI_Comm comm;
    int rank, pid;
    const int nlevels;
    PhysicalMLSolver &pSolver;
    MLSampler &sampler;
    parelag::ParameterList &prob_list;
    double eps2;
    bool auto_eps2;
    const double ratio;
    const std::string file_name;
    const int init_nsamples;
    bool use_array_samples;
    std::vector<int> v_init_nsamples;
    double ml_estimator_variance, expected_discretization_error2, actualMSE;
    std::vector<double> sums;  // nlevels x NVAR, row-major
    std::vector<double> eY, eABSY, eQ, eABSQ, eC, varY, varQ, consistency, kurtosis, M, VC, level_time;
    std::vector<size_t> sampler_nnz, physical_nnz;
    double alpha, alphaABS, beta, gamma;
    std::vector<int> level_nsamples, level_nsamples_missing;
    std::ofstream logger;
    std::vector<pmc_context_s *> clones;  // one extra device handle per level > 0 (concurrent level loops)
    bool comm_ready = false;              // NCCL communicator of the ranks (RankComm.hpp) created
};

/// expWRegression (/root/reference/src/Utilities.cpp:257-283)
double expWRegression(const std::vector<double> &y, const std::vector<double> &x, int skip_n_last);
}  // namespace parelagmc
