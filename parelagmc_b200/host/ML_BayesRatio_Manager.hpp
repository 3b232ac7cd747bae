// ML_BayesRatio_Manager.hpp -- multilevel ratio estimator of the posterior expectation E[Q | data] = E[R] / E[Z]
// (R = Q * likelihood, Z = likelihood, independent prior draws for R and Z); constructor, Run, InitRun, ShowMe and
// wallTime as in /root/reference/src/ML_BayesRatio_Manager.hpp:35-130.  With nlevels = 1 it is the reference's
// SL_BayesRatio_Manager.  When the problem's prior and solver are the B200 classes the level loops (hpp:313-424) run as one
// batched device call per level (pmc_bayes_level_batch); otherwise the reference's per-sample loop runs through
// BayesianInverseProblem.
#pragma once
#include <iostream>
#include <vector>
#include "BayesianInverseProblem.hpp"

namespace parelagmc {
class ML_BayesRatio_Manager {
public:
    ML_BayesRatio_Manager(MPI_Comm comm, const int nlevels, BayesianInverseProblem &problem, parelag::ParameterList &params);
    ML_BayesRatio_Manager(ML_BayesRatio_Manager const &) = delete;
    ML_BayesRatio_Manager &operator=(ML_BayesRatio_Manager const &) = delete;
    void Run();
    void InitRun(std::vector<int> &level_nsamples_init);
    void ShowMe(std::ostream &os = std::cout);
    bool wallTime;
    double RatioEstimate() const;
    const std::vector<double> &Sums() const { return sums; }
    const std::vector<int> &NumSamples() const { return level_nsamples; }

private:
    // enum of the reference (hpp:67-70)
    enum { YZ2 = 0, YZ = 1, ABS_YZ = 2, Z2 = 3, Z = 4, ABS_Z = 5, YR2 = 6, YR = 7, ABS_YR = 8, R2 = 9, R = 10, ABS_R = 11,
           C = 18, T = 19, NVAR = 20 };
    void computeNSamplesMSE();
    MPI_Comm comm;
    int rank, pid;
    const int nlevels;
    BayesianInverseProblem &problem;
    parelag::ParameterList &prob_list;
    double eps2;
    bool auto_eps2;
    const double ratio;
    const int init_nsamples;
    std::vector<int> v_init_nsamples;
    double ml_estimator_variance, expected_discretization_error2, actualMSE;
    std::vector<double> sums;   // nlevels x NVAR
    std::vector<double> eR, eABS_R, varR, eYR, eABS_YR, varYR, eZ, eABS_Z, varZ, eYZ, eABS_YZ, varYZ, eC, M, level_time;
    double alphaABS_R = 0, alphaABS_Z = 0;
    std::vector<int> level_nsamples, level_nsamples_missing;
    bool comm_ready = false;
};
}  // namespace parelagmc
