#include "MC_Manager.hpp"

#include <stdexcept>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <iomanip>
#include <limits>
#include <vector>

#include "DarcySolver.hpp"
#include "PDESampler.hpp"
#include "RankComm.hpp"

namespace parelagmc {

MC_Manager::MC_Manager(MPI_Comm comm_, PhysicalMLSolver &pSolver_, MLSampler &sampler_,
                       parelag::ParameterList &master_list)
    : wallTime(true), comm(comm_), rank(1), pid(0), pSolver(pSolver_), sampler(sampler_),
      prob_list(master_list.Sublist("Problem parameters", true)),
      eps2(prob_list.Get("Mean square error", 0.001)), auto_eps2(eps2 < 0),
      ratio(prob_list.Get("MSE splitting ratio", 0.5)),
      file_name(prob_list.Get("Output filename for MC managers", "MLMC.dat")),
      init_nsamples(prob_list.Get("Number of samples", 10)),
      ml_estimator_variance(std::numeric_limits<double>::infinity()),
      expected_discretization_error2(std::numeric_limits<double>::infinity()),
      actualMSE(std::numeric_limits<double>::infinity()), eQ(0), eABSQ(0), eC(0), varQ(0),
      M(pSolver_.GetGlobalNumberOfDofs(0)), time_(0), level_nsamples(0), level_nsamples_missing(0)
{
    MPI_Comm_size(comm, &rank);  // the communicator size, as in the reference (src/MC_Manager.cpp:44)
    MPI_Comm_rank(comm, &pid);
    std::fill(sums, sums + NVAR, 0.);
    if (pid == 0 && !file_name.empty()) logger.open(file_name);
    if (!pid)
        std::cout << '\n' << std::string(50, '*') << '\n'
                  << "*  MC_Manager \n"
                  << "*    MSE: " << eps2 << '\n'
                  << "*    MSE splitting ratio: " << ratio << '\n'
                  << "*    Number of Initial Samples: " << init_nsamples << '\n'
                  << "*    Output filename: " << file_name << '\n' << std::string(50, '*') << '\n';
}

void MC_Manager::InitRun(int nsamples)
{
    PDESampler *bs = dynamic_cast<PDESampler *>(&sampler);
    DarcySolver *bd = dynamic_cast<DarcySolver *>(&pSolver);
    const bool batched = bs && bd && bs->Device().get() == bd->Device().get();
    const auto t0 = std::chrono::steady_clock::now();
    if (batched && nsamples > 0) {
        // the ranks own disjoint slices of the realisations; one all-reduce of {sums, time} per InitRun (RankComm.hpp)
        pmc_handle h = bs->Device()->handle();
        bs->Distribution().Bind();
        if (rank > 1 && !comm_ready) {
            InitDeviceComm(comm, h);
            comm_ready = true;
        }
        int first = 0, mine = 0;
        SplitSamples(nsamples, pid, rank, first, mine);
        const uint64_t pos0 = bs->Distribution().Advance((uint64_t)nsamples * (uint64_t)bs->NoiseSize(0)) +
                              (uint64_t)first * (uint64_t)bs->NoiseSize(0);
        std::vector<double> rows(logger.is_open() ? (size_t)mine * 2 : 0);
        double round[NVAR + 1] = {0.};
        if (mine > 0)
            bs->Device()->check(pmc_mc_level_batch(h, 0, mine, pos0, round, rows.empty() ? nullptr : rows.data(), nullptr),
                                "pmc_mc_level_batch");
        round[NVAR] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
        if (rank > 1) bs->Device()->check(pmc_allreduce_sums(h, round, NVAR + 1), "pmc_allreduce_sums");
        for (int i = 0; i < NVAR; ++i) sums[i] += round[i];
        time_ += round[NVAR];
        if (!pid && logger.is_open())
            for (int j = 0; j < mine; ++j)
                logger << std::setw(14) << rows[2 * j] << std::setw(14) << rows[2 * j + 1] << "\n";
    } else {
        // reference loop (src/MC_Manager.cpp:91-109)
        mfem::Vector xi, sparam;
        for (int isample = 0; isample < nsamples; ++isample) {
            double q = 0, c = 0;
            sampler.Sample(0, xi);
            sampler.Eval(0, xi, sparam);
            pSolver.SolveFwd(0, sparam, q, c);
            sums[Q2] += q * q;
            sums[Q] += q;
            sums[ABSQ] += std::fabs(q);
            sums[C] += c;
            if (!pid && logger.is_open()) logger << std::setw(14) << q << std::setw(14) << c << "\n";
        }
    }
    if (!(batched && nsamples > 0)) time_ += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    level_nsamples += nsamples;
    if (pid == 0 && logger.is_open()) logger << std::flush;
    computeNSamplesMSE();
}

void MC_Manager::Run()
{
    std::fill(sums, sums + NVAR, 0.);
    level_nsamples = level_nsamples_missing = 0;
    time_ = 0;
    int level_nsamples_grain = init_nsamples;
    InitRun(level_nsamples_grain);
    level_nsamples_grain = 0;
    while (ml_estimator_variance > ratio * eps2) {
        level_nsamples_grain = std::min(level_nsamples_missing, init_nsamples + level_nsamples_grain + level_nsamples_missing / 10);
        if (level_nsamples_grain == 0) break;
        InitRun(level_nsamples_grain);
    }
    if (!pid) std::cout << "FINAL SLMC ERRORS" << std::endl;
    ShowMe();
}

void MC_Manager::ShowMe(std::ostream &os)
{
    const int total_width = 79, name_width = 40;
    if (pid) return;
    auto row = [&](const char *name, double v) {
        os << std::setw(name_width + 2) << std::left << name << std::setw(18) << std::left << v << '\n';
    };
    os.precision(8);
    os << std::string(total_width, '=') << std::endl;
    os << "SLMC Manager Errors: " << std::endl << std::string(total_width, '-') << std::endl;
    row("Estimate", eQ);
    row("Target MSE", eps2);
    row("Actual MSE", actualMSE);
    row("SL Estimator Variance", ml_estimator_variance);
    row("Estimator Bias", expected_discretization_error2);
    row("Target Bias Error", std::sqrt(eps2) / std::sqrt(2));
    row("DOFS in Forward Problem", M);
    row("C_l ", eC);
    os << '\n' << std::setw(name_width + 2) << std::left << "NumSamples " << std::setw(2) << std::left << level_nsamples << '\n';
    row("E[Q_l] ", eQ);
    row("E[|Q_l|] ", eABSQ);
    row("Var[Q_l] ", varQ);
    os << std::string(total_width, '=') << std::endl;
}

void MC_Manager::computeNSamplesMSE()
{
    // src/MC_Manager.cpp:194-239
    if (level_nsamples < 2) throw std::runtime_error("MC_Manager: at least 2 samples are needed for the variance estimate");
    const double nl = static_cast<double>(level_nsamples);
    eQ = sums[Q] / nl;
    eABSQ = sums[ABSQ] / nl;
    eC = sums[C] / nl;
    varQ = sums[Q2] / nl;
    varQ -= eQ * eQ;
    varQ *= nl / (nl - 1.);
    expected_discretization_error2 = 0.;
    if (auto_eps2) eps2 = expected_discretization_error2 / (1. - ratio);
    ml_estimator_variance = varQ / nl;
    actualMSE = expected_discretization_error2 + ml_estimator_variance;
    const double cost = wallTime ? time_ / nl : eC;
    const double prop = std::sqrt(varQ * cost) / (ratio * eps2);
    const double missings = prop * std::sqrt(varQ / cost) - nl;
    level_nsamples_missing = std::max(static_cast<int>(std::ceil(missings)), 0);
    ShowMe();
}
}  // namespace parelagmc
