#include "MLMC_Manager.hpp"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <iomanip>
#include <limits>
#include <stdexcept>
#include <thread>

#include "DarcySolver.hpp"
#include "PDESampler.hpp"
#include "RankComm.hpp"

namespace parelagmc {

double expWRegression(const std::vector<double> &y, const std::vector<double> &x, int skip_n_last)
{
    // weighted least squares of log|y_i / y_{i+1}| against log(x_i / x_{i+1}), weights 0.5^i
    const int n = static_cast<int>(y.size()) - 1 - skip_n_last;
    double num = 0., den = 0.;
    for (int i = 0; i < n; ++i) {
        const double logdy = std::log(std::fabs(y[i] / y[i + 1]));
        const double logdx = std::log(x[i] / x[i + 1]);
        const double w = std::pow(.5, i);
        num += logdy * w * logdx;
        den += logdx * w * logdx;
    }
    return num / den;
}

MLMC_Manager::MLMC_Manager(MPI_Comm comm_, const int nlevels_, PhysicalMLSolver &pSolver_, MLSampler &sampler_,
                           parelag::ParameterList &master_list)
    : wallTime(true), comm(comm_), rank(1), pid(0), nlevels(nlevels_), pSolver(pSolver_), sampler(sampler_),
      prob_list(master_list.Sublist("Problem parameters", true)),
      eps2(prob_list.Get("Mean square error", 0.001)), auto_eps2(eps2 < 0),
      ratio(prob_list.Get("MSE splitting ratio", 0.5)),
      file_name(prob_list.Get("Output filename for MC managers", "MLMC.dat")),
      init_nsamples(prob_list.Get("Number of samples", 10)),
      use_array_samples(prob_list.Get("Use array samples", false)),
      v_init_nsamples(prob_list.Get("Array number of samples", std::vector<int>())),
      ml_estimator_variance(std::numeric_limits<double>::infinity()),
      expected_discretization_error2(std::numeric_limits<double>::infinity()),
      actualMSE(std::numeric_limits<double>::infinity()),
      sums(nlevels_ * NVAR, 0.), eY(nlevels_), eABSY(nlevels_), eQ(nlevels_), eABSQ(nlevels_), eC(nlevels_),
      varY(nlevels_), varQ(nlevels_), consistency(nlevels_), kurtosis(nlevels_), M(nlevels_), VC(nlevels_),
      level_time(nlevels_, 0.), sampler_nnz(nlevels_), physical_nnz(nlevels_), alpha(0.), alphaABS(0.), beta(0.),
      gamma(0.), level_nsamples(nlevels_, 0), level_nsamples_missing(nlevels_, 0)
{
    MPI_Comm_size(comm, &rank);  // sic: the reference keeps the communicator size in `rank` (src/MLMC_Manager.cpp:62)
    MPI_Comm_rank(comm, &pid);
    for (int i = 0; i < nlevels; ++i) M[i] = pSolver.GetGlobalNumberOfDofs(i);
    if (pid == 0 && !file_name.empty()) logger.open(file_name);
    if (use_array_samples && static_cast<int>(v_init_nsamples.size()) != nlevels) use_array_samples = false;
    if (!use_array_samples) v_init_nsamples.assign(nlevels, init_nsamples);
    if (!pid) {
        std::cout << '\n' << std::string(50, '*') << '\n'
                  << "*  MLMC_Manager \n"
                  << "*    MSE: " << eps2 << '\n'
                  << "*    MSE splitting ratio: " << ratio << '\n'
                  << "*    Number of Initial Samples: ";
        for (auto i : v_init_nsamples) std::cout << i << " ";
        std::cout << "\n*    Output filename: " << file_name << '\n' << std::string(50, '*') << '\n';
    }
}

MLMC_Manager::~MLMC_Manager()
{
    for (auto h : clones) pmc_destroy(h);
}

void MLMC_Manager::accumulate(int l, double y, double q, double c)
{
    double *s = &sums[l * NVAR];
    s[Y3] += y * y * y;
    s[Y4] += y * y * y * y;
    s[Y2] += y * y;
    s[Y] += y;
    s[ABSY] += std::fabs(y);
    s[Q2] += q * q;
    s[Q] += q;
    s[ABSQ] += std::fabs(q);
    s[C] += c;
}

void MLMC_Manager::InitRun(std::vector<int> &level_nsamples_init)
{
    const int width = 14;
    if (*std::max_element(level_nsamples.begin(), level_nsamples.end()) == 0 && pid == 0 && logger.is_open())
        logger << "%" << std::setw(13) << "level " << std::setw(width) << "Y(xi) " << std::setw(width) << "Q(xi)"
               << std::setw(width) << "Q_c(xi)" << std::setw(width) << "c \n";
    PDESampler *bs = dynamic_cast<PDESampler *>(&sampler);
    DarcySolver *bd = dynamic_cast<DarcySolver *>(&pSolver);
    const bool batched = bs && bd && bs->Device().get() == bd->Device().get();
    if (batched) {
        // The level loops of an InitRun are independent of one another: every level gets its own device handle
        // (pmc_clone: own stream and workspace) and its own host thread, so the levels' kernels share the GPU.  The
        // stream positions are the ones the reference's sequential Sample() calls would have consumed: coarsest level
        // first, then nlevels-2 .. 0 (src/MLMC_Manager.cpp:110,140).
        pmc_handle main_h = bs->Device()->handle();
        bs->Distribution().Bind();
        if (rank > 1 && !comm_ready) {   // the ranks own disjoint slices of every level's realisations (RankComm.hpp)
            InitDeviceComm(comm, main_h);
            comm_ready = true;
        }
        while ((int)clones.size() < nlevels - 1) {
            pmc_handle h = nullptr;
            bs->Device()->check(pmc_clone(main_h, &h), "pmc_clone");
            clones.push_back(h);
        }
        std::vector<uint64_t> pos0(nlevels);
        for (int ilevel = nlevels - 1; ilevel >= 0; --ilevel)
            pos0[ilevel] = bs->Distribution().Advance((uint64_t)level_nsamples_init[ilevel] * (uint64_t)bs->NoiseSize(ilevel));
        std::vector<std::vector<double>> rows(nlevels);
        std::vector<int> rcs(nlevels, 0), mine(nlevels, 0);
        // this round's contribution of this rank: [nlevels x NVAR sums | nlevels times]; reduced over the ranks in ONE
        // collective, so that the sample allocation below sees identical inputs on every rank
        std::vector<double> round(nlevels * (NVAR + 1), 0.0);
        double *secs = &round[nlevels * NVAR];
        std::vector<std::thread> workers;
        for (int ilevel = 0; ilevel < nlevels; ++ilevel) {
            int first = 0;
            SplitSamples(level_nsamples_init[ilevel], pid, rank, first, mine[ilevel]);
            const int nsamples = mine[ilevel];
            if (nsamples <= 0) continue;
            if (logger.is_open()) rows[ilevel].resize((size_t)nsamples * 4);
            pmc_handle h = ilevel == 0 ? main_h : clones[ilevel - 1];
            const uint64_t pos = pos0[ilevel] + (uint64_t)first * (uint64_t)bs->NoiseSize(ilevel);
            workers.emplace_back([&, ilevel, nsamples, h, pos]() {
                const auto t0 = std::chrono::steady_clock::now();
                rcs[ilevel] = pmc_mlmc_level_batch(h, ilevel, nlevels, nsamples, pos, &round[ilevel * NVAR],
                                                   rows[ilevel].empty() ? nullptr : rows[ilevel].data(), nullptr);
                secs[ilevel] = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            });
        }
        for (auto &w : workers) w.join();
        for (int ilevel = 0; ilevel < nlevels; ++ilevel)
            if (rcs[ilevel] != 0)
                throw std::runtime_error(std::string("pmc_mlmc_level_batch: ") +
                                         pmc_last_error(ilevel == 0 ? main_h : clones[ilevel - 1]));
        if (rank > 1) bs->Device()->check(pmc_allreduce_sums(main_h, round.data(), (int)round.size()), "pmc_allreduce_sums");
        for (int i = 0; i < nlevels * NVAR; ++i) sums[i] += round[i];
        for (int ilevel = nlevels - 1; ilevel >= 0; --ilevel) {
            const int nsamples = level_nsamples_init[ilevel];
            const bool coarsest = (ilevel == nlevels - 1);
            if (!pid && logger.is_open())   // with several ranks the log holds rank 0's slice
                for (int j = 0; j < mine[ilevel]; ++j) {
                    const double *r = &rows[ilevel][4 * (size_t)j];
                    logger << std::setw(width) << ilevel << std::setw(width) << r[0] << std::setw(width) << r[1]
                           << std::setw(width);
                    if (coarsest) logger << "0"; else logger << r[2];
                    logger << std::setw(width) << r[3] << "\n";
                }
            level_time[ilevel] += secs[ilevel];
            level_nsamples[ilevel] += nsamples;
        }
    } else {
        // reference loop through the abstract interfaces (src/MLMC_Manager.cpp:113-136 / :144-173)
        for (int ilevel = nlevels - 1; ilevel >= 0; --ilevel) {
            const int nsamples = level_nsamples_init[ilevel];
            const bool coarsest = (ilevel == nlevels - 1);
            const auto t0 = std::chrono::steady_clock::now();
            mfem::Vector xi, sparam, init_s;
            for (int isample = 0; isample < nsamples; ++isample) {
                double q = 0, c = 0, qc = 0, cc = 0, y;
                sampler.Sample(ilevel, xi);
                if (coarsest) {
                    sampler.Eval(ilevel, xi, sparam);
                    pSolver.SolveFwd(ilevel, sparam, q, c);
                    y = q;
                } else {
                    sampler.Eval(ilevel + 1, xi, sparam, init_s, false);
                    pSolver.SolveFwd(ilevel + 1, sparam, qc, cc);
                    sampler.Eval(ilevel, xi, sparam, init_s, true);
                    pSolver.SolveFwd(ilevel, sparam, q, c);
                    y = q - qc;
                    c = c + cc;
                }
                accumulate(ilevel, y, q, c);
                if (!pid && logger.is_open()) {
                    logger << std::setw(width) << ilevel << std::setw(width) << y << std::setw(width) << q
                           << std::setw(width);
                    if (coarsest) logger << "0"; else logger << qc;
                    logger << std::setw(width) << c << "\n";
                }
            }
            level_time[ilevel] += std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
            level_nsamples[ilevel] += nsamples;
        }
    }
    if (pid == 0 && logger.is_open()) logger << std::flush;
    computeNSamplesMSE();
}

void MLMC_Manager::Run()
{
    std::fill(sums.begin(), sums.end(), 0.);
    std::fill(level_nsamples.begin(), level_nsamples.end(), 0);
    std::fill(level_nsamples_missing.begin(), level_nsamples_missing.end(), 0);
    std::fill(level_time.begin(), level_time.end(), 0.);
    InitRun(v_init_nsamples);
    std::vector<int> level_nsamples_grain(nlevels, 0);
    while (ml_estimator_variance > ratio * eps2) {
        int total = 0;
        for (int i = 0; i < nlevels; ++i) {
            level_nsamples_grain[i] = std::min(level_nsamples_missing[i], v_init_nsamples[i] + level_nsamples_grain[i] +
                                                                              level_nsamples_missing[i] / 10);
            total += level_nsamples_grain[i];
        }
        if (total == 0) break;  // nothing left to add (the reference would spin here)
        InitRun(level_nsamples_grain);
    }
    if (!pid) std::cout << "FINAL MLMC ERRORS" << std::endl;
    ShowMe();
}

double MLMC_Manager::Estimate() const
{
    double s = 0;
    for (double v : eY) s += v;
    return s;
}

namespace {
template <typename T>
void print_vec(std::ostream &os, const std::vector<T> &v)
{
    for (size_t i = 0; i < v.size(); ++i) os << v[i] << (i + 1 < v.size() ? "  " : "");
    os << "\n";
}
}  // namespace

void MLMC_Manager::ShowMe(std::ostream &os)
{
    for (int i = 0; i < nlevels; i++) {
        physical_nnz[i] = pSolver.GetNNZ(i);
        sampler_nnz[i] = sampler.GetNNZ(i);
    }
    const int total_width = 79, name_width = 40;
    if (pid) return;
    auto row = [&](const char *name, double v) {
        os << std::setw(name_width + 2) << std::left << name << std::setw(18) << std::left << v << '\n';
    };
    auto vec = [&](const char *name) -> std::ostream & { return os << std::setw(name_width + 2) << std::left << name; };
    os.precision(8);
    os << std::string(total_width, '=') << std::endl;
    os << "MLMC Manager Errors: " << std::endl << std::string(total_width, '-') << std::endl;
    row("Estimate", Estimate());
    row("Target MSE", eps2);
    row("Actual MSE", actualMSE);
    row("ML Estimator Variance", ml_estimator_variance);
    row("Estimator Bias", expected_discretization_error2);
    row("Alpha", alpha);
    row("AlphaAbs", alphaABS);
    row("Beta", beta);
    row("Gamma", gamma);
    os << "\n";
    print_vec(vec("DOFS in Forward Problem"), M);
    print_vec(vec("C_l "), eC);
    os << '\n';
    print_vec(vec("NumSamples "), level_nsamples);
    os << '\n';
    print_vec(vec("E[Y_l] "), eY);
    print_vec(vec("E[|Y_l|] "), eABSY);
    print_vec(vec("Var[Y_l] "), varY);
    print_vec(vec("E[Q_l] "), eQ);
    print_vec(vec("E[|Q_l|] "), eABSQ);
    print_vec(vec("Var[Q_l] "), varQ);
    print_vec(vec("V[Y_l]*C_l "), VC);
    print_vec(vec("Consistency "), consistency);
    print_vec(vec("Kurtosis"), kurtosis);
    print_vec(vec("NNZ-Sampler"), sampler_nnz);
    print_vec(vec("NNZ-ForwardSolve"), physical_nnz);
    os << std::string(total_width, '=') << std::endl;
}

void MLMC_Manager::computeNSamplesMSE()
{
    for (int l = 0; l < nlevels; ++l)   // the reference would go on with inf/NaN (src/MLMC_Manager.cpp:319-321)
        if (level_nsamples[l] < 2)
            throw std::runtime_error("MLMC_Manager: at least 2 samples per level are needed for the variance estimate (level " +
                                     std::to_string(l) + " has " + std::to_string(level_nsamples[l]) + ")");
    for (int l = 0; l < nlevels; ++l) {
        const double n = static_cast<double>(level_nsamples[l]);
        const double *s = &sums[l * NVAR];
        eY[l] = s[Y] / n;
        eABSY[l] = s[ABSY] / n;
        eQ[l] = s[Q] / n;
        eABSQ[l] = s[ABSQ] / n;
        eC[l] = s[C] / n;
        varY[l] = s[Y2] / n;
        varQ[l] = s[Q2] / n;
        kurtosis[l] = s[Y4] / n;
        kurtosis[l] /= varY[l] * varY[l];  // before the mean is subtracted, as in the reference (:318-319)
        varY[l] -= eY[l] * eY[l];
        varY[l] *= n / static_cast<double>(level_nsamples[l] - 1);
        varQ[l] -= eQ[l] * eQ[l];
        varQ[l] *= n / static_cast<double>(level_nsamples[l] - 1);
    }
    for (int l = 0; l < nlevels - 1; ++l)
        consistency[l] = std::abs(eQ[l] - eQ[l + 1] + eY[l]) /
                         (3 * (std::sqrt(varQ[l]) + std::sqrt(varQ[l + 1]) + std::sqrt(varY[l])));
    alpha = expWRegression(eY, M, 1);
    alphaABS = expWRegression(eABSY, M, 1);
    beta = expWRegression(varY, M, 1);
    if (nlevels == 1)
        expected_discretization_error2 = 0.;
    else {
        const double m = M[0] / M[1];
        if (nlevels > 3)
            expected_discretization_error2 = std::max(std::pow(m, 2. * alphaABS) * eABSY[1] * eABSY[1], eABSY[0] * eABSY[0]) /
                                             (std::pow(std::pow(m, -2. * alphaABS) - 1., 2));
        else if (nlevels == 3)
            expected_discretization_error2 = (eABSY[0] * eABSY[0]) / (std::pow(std::pow(m, -alphaABS) - 1., 2));
        else
            expected_discretization_error2 = (eABSY[0] * eABSY[0]);
    }
    if (auto_eps2) eps2 = expected_discretization_error2 / (1. - ratio);
    ml_estimator_variance = 0.;
    for (int l = 0; l < nlevels; ++l) ml_estimator_variance += varY[l] / static_cast<double>(level_nsamples[l]);
    actualMSE = expected_discretization_error2 + ml_estimator_variance;
    std::vector<double> cost(nlevels);
    for (int i = 0; i < nlevels; ++i)  // wall time per sample, or the dof-count cost (:368-382)
        cost[i] = wallTime ? level_time[i] / static_cast<double>(level_nsamples[i]) : eC[i];
    gamma = expWRegression(cost, M, 0);
    double prop = 0.;
    for (int i = 0; i < nlevels; ++i) prop += std::sqrt(varY[i] * cost[i]);
    prop /= ratio * eps2;
    for (int i = 0; i < nlevels; ++i) {
        double missings = prop * std::sqrt(varY[i] / cost[i]);
        missings -= static_cast<double>(level_nsamples[i]);
        level_nsamples_missing[i] = std::max(static_cast<int>(std::ceil(missings)), 0);
        VC[i] = varY[i] * cost[i];
    }
    ShowMe();
}
}  // namespace parelagmc
