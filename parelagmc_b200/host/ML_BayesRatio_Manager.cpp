#include "ML_BayesRatio_Manager.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <iomanip>
#include <limits>
#include <stdexcept>

#include "MLMC_Manager.hpp"   // expWRegression
#include "PDESampler.hpp"
#include "RankComm.hpp"

namespace parelagmc {

ML_BayesRatio_Manager::ML_BayesRatio_Manager(MPI_Comm comm_, const int nlevels_, BayesianInverseProblem &problem_,
                                             parelag::ParameterList &master_list)
    : wallTime(true), comm(comm_), rank(1), pid(0), nlevels(nlevels_), problem(problem_),
      prob_list(master_list.Sublist("Problem parameters", true)), eps2(prob_list.Get("Mean square error", 0.001)),
      auto_eps2(eps2 < 0), ratio(prob_list.Get("MSE splitting ratio", 0.5)), init_nsamples(prob_list.Get("Number of samples", 10)),
      v_init_nsamples(prob_list.Get("Array number of samples", std::vector<int>())),
      ml_estimator_variance(std::numeric_limits<double>::infinity()),
      expected_discretization_error2(std::numeric_limits<double>::infinity()),
      actualMSE(std::numeric_limits<double>::infinity()), sums(nlevels_ * NVAR, 0.), eR(nlevels_), eABS_R(nlevels_),
      varR(nlevels_), eYR(nlevels_), eABS_YR(nlevels_), varYR(nlevels_), eZ(nlevels_), eABS_Z(nlevels_), varZ(nlevels_),
      eYZ(nlevels_), eABS_YZ(nlevels_), varYZ(nlevels_), eC(nlevels_), M(nlevels_), level_time(nlevels_, 0.),
      level_nsamples(nlevels_, 0), level_nsamples_missing(nlevels_, 0)
{
    MPI_Comm_size(comm, &rank);
    MPI_Comm_rank(comm, &pid);
    for (int i = 0; i < nlevels; ++i) M[i] = problem.GetSolver().GetGlobalNumberOfDofs(i);
    if ((int)v_init_nsamples.size() != nlevels) v_init_nsamples.assign(nlevels, init_nsamples);
}

void ML_BayesRatio_Manager::InitRun(std::vector<int> &level_nsamples_init)
{
    PDESampler *bs = dynamic_cast<PDESampler *>(&problem.GetPrior());
    DarcySolver *bd = dynamic_cast<DarcySolver *>(&problem.GetSolver());
    const bool batched = bs && bd && bs->Device().get() == bd->Device().get();
    std::vector<double> round(nlevels * (NVAR + 1), 0.0);   // this rank's sums of the round | its timings
    double *secs = &round[nlevels * NVAR];
    if (batched) {
        B200Device &dev = *bs->Device();
        pmc_handle h = dev.handle();
        bs->Distribution().Bind();
        if (rank > 1 && !comm_ready) { InitDeviceComm(comm, h); comm_ready = true; }
        for (int ilevel = nlevels - 1; ilevel >= 0; --ilevel) {   // coarsest first (hpp:322,366); two draws per realisation
            const int n = level_nsamples_init[ilevel];
            int first = 0, mine = 0;
            SplitSamples(n, pid, rank, first, mine);
            const uint64_t Ne = (uint64_t)bs->NoiseSize(ilevel);
            const uint64_t pos0 = bs->Distribution().Advance(2ull * (uint64_t)n * Ne) + 2ull * (uint64_t)first * Ne;
            const auto t0 = std::chrono::steady_clock::now();
            if (mine > 0)
                dev.check(pmc_bayes_level_batch(h, ilevel, nlevels, mine, pos0, &round[ilevel * NVAR], nullptr, nullptr),
                          "pmc_bayes_level_batch");
            secs[ilevel] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            level_nsamples[ilevel] += n;
        }
        if (rank > 1) dev.check(pmc_allreduce_sums(h, round.data(), (int)round.size()), "pmc_allreduce_sums");
    } else {
        // the reference's loop through BayesianInverseProblem (hpp:313-424)
        mfem::Vector zxi, xi, zparam, sparam;
        for (int ilevel = nlevels - 1; ilevel >= 0; --ilevel) {
            const bool coarsest = ilevel == nlevels - 1;
            const auto t0 = std::chrono::steady_clock::now();
            for (int isample = 0; isample < level_nsamples_init[ilevel]; ++isample) {
                double z, r, zc = 0, rc = 0, c, c_tot = 0;
                problem.SamplePrior(ilevel, zxi);
                problem.EvalPrior(ilevel, zxi, zparam);
                problem.ComputeLikelihood(ilevel, zparam, z, c); c_tot += c;
                problem.SamplePrior(ilevel, xi);
                problem.EvalPrior(ilevel, xi, sparam);
                problem.ComputeR(ilevel, sparam, r, c); c_tot += c;
                if (!coarsest) {
                    problem.EvalPrior(ilevel + 1, zxi, zparam);
                    problem.ComputeLikelihood(ilevel + 1, zparam, zc, c); c_tot += c;
                    problem.EvalPrior(ilevel + 1, xi, sparam);
                    problem.ComputeR(ilevel + 1, sparam, rc, c); c_tot += c;
                }
                const double y_r = r - rc, y_z = z - zc;
                double *s = &round[ilevel * NVAR];
                s[R] += r; s[ABS_R] += std::fabs(r); s[R2] += r * r;
                s[YR] += y_r; s[ABS_YR] += std::fabs(y_r); s[YR2] += y_r * y_r;
                s[Z] += z; s[ABS_Z] += std::fabs(z); s[Z2] += z * z;
                s[YZ] += y_z; s[ABS_YZ] += std::fabs(y_z); s[YZ2] += y_z * y_z;
                s[C] += c_tot;
            }
            secs[ilevel] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            level_nsamples[ilevel] += level_nsamples_init[ilevel];
        }
    }
    for (int i = 0; i < nlevels * NVAR; ++i) sums[i] += round[i];
    for (int i = 0; i < nlevels; ++i) level_time[i] += secs[i];
    computeNSamplesMSE();
}

void ML_BayesRatio_Manager::Run()
{
    std::fill(sums.begin(), sums.end(), 0.);
    std::fill(level_nsamples.begin(), level_nsamples.end(), 0);
    std::fill(level_time.begin(), level_time.end(), 0.);
    InitRun(v_init_nsamples);
    std::vector<int> grain(nlevels, 0);
    while (ml_estimator_variance > ratio * eps2) {
        int total = 0;
        for (int i = 0; i < nlevels; ++i) {
            grain[i] = std::min(level_nsamples_missing[i], v_init_nsamples[i] + grain[i] + level_nsamples_missing[i] / 10);
            total += grain[i];
        }
        if (total == 0) break;
        InitRun(grain);
    }
    if (!pid) std::cout << "FINAL ML_BayesRatio_Manager ERRORS" << std::endl;
    ShowMe();
}

double ML_BayesRatio_Manager::RatioEstimate() const
{
    double r = 0, z = 0;
    for (int l = 0; l < nlevels; ++l) { r += eYR[l]; z += eYZ[l]; }
    return r / z;
}

void ML_BayesRatio_Manager::computeNSamplesMSE()
{
    // hpp:572-726
    for (int l = 0; l < nlevels; ++l) {
        if (level_nsamples[l] < 2) throw std::runtime_error("ML_BayesRatio_Manager: at least 2 samples per level are needed");
        const double n = level_nsamples[l], unb = n / (n - 1.);
        const double *s = &sums[l * NVAR];
        eR[l] = s[R] / n; eABS_R[l] = s[ABS_R] / n; varR[l] = (s[R2] / n - eR[l] * eR[l]) * unb;
        eYR[l] = s[YR] / n; eABS_YR[l] = s[ABS_YR] / n; varYR[l] = (s[YR2] / n - eYR[l] * eYR[l]) * unb;
        eZ[l] = s[Z] / n; eABS_Z[l] = s[ABS_Z] / n; varZ[l] = (s[Z2] / n - eZ[l] * eZ[l]) * unb;
        eYZ[l] = s[YZ] / n; eABS_YZ[l] = s[ABS_YZ] / n; varYZ[l] = (s[YZ2] / n - eYZ[l] * eYZ[l]) * unb;
        eC[l] = s[C] / n;
    }
    std::vector<double> cost(nlevels);
    for (int i = 0; i < nlevels; ++i) cost[i] = wallTime ? level_time[i] / level_nsamples[i] : eC[i];
    alphaABS_R = expWRegression(eABS_YR, M, 1);
    alphaABS_Z = expWRegression(eABS_YZ, M, 1);
    auto bias2 = [&](const std::vector<double> &eABS, double a) {
        if (nlevels == 1) return 0.;
        const double m = M[0] / M[1];
        if (nlevels > 3) return std::max(std::pow(m, 2. * a) * eABS[1] * eABS[1], eABS[0] * eABS[0]) / std::pow(std::pow(m, -2. * a) - 1., 2);
        if (nlevels == 3) return eABS[0] * eABS[0] / std::pow(std::pow(m, -a) - 1., 2);
        return eABS[0] * eABS[0];
    };
    expected_discretization_error2 = std::max(bias2(eABS_YR, alphaABS_R), bias2(eABS_YZ, alphaABS_Z));
    if (auto_eps2) eps2 = expected_discretization_error2 / (1. - ratio);
    double vz = 0, vr = 0, prop_R = 0, prop_Z = 0;
    for (int l = 0; l < nlevels; ++l) {
        vz += varYZ[l] / level_nsamples[l];
        vr += varYR[l] / level_nsamples[l];
        prop_R += std::sqrt(std::max(varYR[l], 0.) * cost[l]);
        prop_Z += std::sqrt(std::max(varYZ[l], 0.) * cost[l]);
    }
    ml_estimator_variance = std::max(vz, vr);
    actualMSE = expected_discretization_error2 + ml_estimator_variance;
    prop_R /= ratio * eps2;
    prop_Z /= ratio * eps2;
    for (int i = 0; i < nlevels; ++i) {
        const double mR = prop_R * std::sqrt(std::max(varYR[i], 0.) / cost[i]) - level_nsamples[i];
        const double mZ = prop_Z * std::sqrt(std::max(varYZ[i], 0.) / cost[i]) - level_nsamples[i];
        level_nsamples_missing[i] = std::max({static_cast<int>(std::ceil(mR)), static_cast<int>(std::ceil(mZ)), 0});
    }
}

void ML_BayesRatio_Manager::ShowMe(std::ostream &os)
{
    if (pid) return;
    const int total_width = 79, name_width = 40;
    auto row = [&](const char *name, double v) {
        os << std::setw(name_width + 2) << std::left << name << std::setw(18) << std::left << v << '\n';
    };
    auto vec = [&](const char *name, const std::vector<double> &v) {
        os << std::setw(name_width + 2) << std::left << name;
        for (size_t i = 0; i < v.size(); ++i) os << v[i] << (i + 1 < v.size() ? "  " : "");
        os << '\n';
    };
    double r = 0, z = 0;
    for (int l = 0; l < nlevels; ++l) { r += eYR[l]; z += eYZ[l]; }
    os.precision(8);
    os << std::string(total_width, '=') << std::endl << "ML_BayesRatio_Manager Errors: " << std::endl
       << std::string(total_width, '-') << std::endl;
    row("R Estimate", r);
    row("Z Estimate", z);
    row("Ratio Estimate", r / z);
    os << '\n';
    row("Target MSE", eps2);
    row("Actual MSE", actualMSE);
    row("ML Estimator Variance", ml_estimator_variance);
    row("Estimator Bias (Max of R,Z)", expected_discretization_error2);
    vec("DOFS in Forward Problem", M);
    vec("Cost (dofs)", eC);
    os << std::setw(name_width + 2) << std::left << "NumSamples ";
    for (int l = 0; l < nlevels; ++l) os << level_nsamples[l] << (l + 1 < nlevels ? "  " : "");
    os << "\n";
    row("AlphaAbs_R", alphaABS_R);
    vec("E[R] ", eR); vec("E[|R|] ", eABS_R); vec("Var[R] ", varR);
    vec("E[Y_R] ", eYR); vec("E[|Y_R|] ", eABS_YR); vec("Var[Y_R] ", varYR);
    row("AlphaAbs_Z", alphaABS_Z);
    vec("E[Z] ", eZ); vec("E[|Z|] ", eABS_Z); vec("Var[Z] ", varZ);
    vec("E[Y_Z] ", eYZ); vec("E[|Y_Z|] ", eABS_YZ); vec("Var[Y_Z] ", varYZ);
    os << std::string(total_width, '=') << std::endl;
}
}  // namespace parelagmc
