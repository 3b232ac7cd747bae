// LikelihoodExample.cpp -- /root/reference/examples/LikelihoodExample.cpp:256-278 and the estimator loop of
// examples/RatioEstimator_MC.cpp:295-345 on the host layer (enlarged-mesh sampler + DarcySolver + BayesianInverseProblem).
// With the reference's default problem in MFEM's element numbering the output is what the ctests
// `BayesianInverseProblem_LikelihoodEvaluation` ("L = 0 : 0.9279...; L = 1 : 0.9578...; L = 2 : 0.9269...") and, with
// --ratio-mc, `BayesianInverseProblem_MC_RatioEstimator` match (examples/CMakeLists.txt:97-115).
//   LikelihoodExample.exe --hierarchy FILE [--ratio-mc] [--nsamples 10]
#include <cstring>
#include <iomanip>
#include <iostream>
#include <memory>

#include "../BayesianInverseProblem.hpp"
#include "../DarcySolver.hpp"
#include "../NormalDistributionSampler.hpp"
#include "../PDESampler.hpp"
#include "driver_common.hpp"

using namespace parelagmc;

int main(int argc, char **argv)
{
    try {
        DriverArgs a = DriverArgs::Parse(argc, argv);
        bool ratio_mc = false;
        for (int i = 1; i < argc; ++i) ratio_mc |= !strcmp(argv[i], "--ratio-mc");
        auto hier = std::make_shared<HierarchyData>(HierarchyData::Load(a.hierarchy));
        int nLevels = hier->nlevels;
        parelag::ParameterList master_list("Default");
        auto &prob = master_list.Sublist("Problem parameters");
        prob.Set("Correlation length", hier->corlen);
        prob.Set("Lognormal", true);
        master_list.Sublist("Bayesian inverse problem parameters").Set("Noise", 0.1);   // CreateBayesianParameterList.hpp:48
        auto dev = std::make_shared<B200Device>(a.device, nLevels);
        dev->check(pmc_set_tolerances(dev->handle(), a.rel_tol, a.abs_tol, a.max_iter), "pmc_set_tolerances");
        DarcySolver solver(hier, dev, master_list);
        solver.BuildHierachySpaces();
        NormalDistributionSampler dist(0, a.variance, dev);
        dist.Split(1, 0);
        L2ProjectionPDESampler sampler(hier, dist, master_list);
        sampler.BuildHierarchy();
        BayesianInverseProblem bayesian_problem(hier, solver, sampler, master_list);
        bayesian_problem.GenerateObservationalData();
        mfem::Vector xi, u;
        if (!ratio_mc) {
            double like, c;
            bayesian_problem.SamplePrior(0, xi);
            std::cout << "Likelihood value:" << std::endl;
            for (int i = 0; i < nLevels; i++) {
                bayesian_problem.EvalPrior(i, xi, u);
                bayesian_problem.ComputeLikelihood(i, u, like, c);
                std::cout << "L = " << i << " : " << like << std::endl;
            }
            return EXIT_SUCCESS;
        }
        // RatioEstimator_MC with "Use independent samples" = false (its default): R and Z from the same realisation
        double c, R, Z, r = 0., r2 = 0., z = 0., z2 = 0., ratio_diff = 0.;
        const double n = static_cast<double>(a.nsamples);
        mfem::Vector zxi, zcoef;
        for (int i = 0; i < a.nsamples; i++) {
            bayesian_problem.SamplePrior(0, zxi);
            bayesian_problem.EvalPrior(0, zxi, zcoef);
            bayesian_problem.ComputeLikelihood(0, zcoef, Z, c);
            bayesian_problem.ComputeR(0, zcoef, R, c);
            r += R; r2 += R * R;
            z += Z; z2 += Z * Z;
            ratio_diff += R / Z;
        }
        r /= n; r2 /= n; z /= n; z2 /= n; ratio_diff /= n;
        const double form_var_r = n * (r2 - r * r) / (n - 1.), form_var_z = n * (z2 - z * z) / (n - 1.);
        std::cout << "L" << std::setw(10) << "E[R] " << std::setw(12) << "Var[R] " << std::setw(12) << "E[Z] " << std::setw(12)
                  << "Var[Z] " << std::setw(12) << "E[Q] " << std::setw(15) << "Splitting E[Q] \n";
        std::cout << 0 << std::setw(10) << r << std::setw(12) << form_var_r << std::setw(12) << z << std::setw(12) << form_var_z
                  << std::setw(12) << r / z << std::setw(15) << ratio_diff << '\n';
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
    }
    return EXIT_SUCCESS;
}
