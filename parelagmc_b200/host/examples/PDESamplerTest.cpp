// PDESamplerTest.cpp -- the statistics loop of /root/reference/examples/PDESamplerTest.cpp:188-275 on the host layer:
// one realisation on every level from fine noise, then per level `nsamples` x (Sample, Eval), the squared L2 errors of the
// sample mean and of the second moment against the exact Gaussian moments (PDESampler::ComputeL2Error,
// src/PDESampler.cpp:614-624: here sum_e |e| (v_e - exact)^2 on the level's own elements, which is what the prolongation
// to the fine grid integrates), printed by OutputRandomFieldErrors' format (src/Utilities.cpp:697-715).  With the
// reference's default problem in MFEM's element numbering the first column is what the ctest `PDESamplerTest` matches
// (examples/CMakeLists.txt:83-87: 1.2593 / 9.3103 / 6.3853).
//   PDESamplerTest.exe --hierarchy FILE [--nsamples 10] [--rel-tol X]
#include <cmath>
#include <iomanip>
#include <iostream>
#include <memory>

#include "../NormalDistributionSampler.hpp"
#include "../PDESampler.hpp"
#include "driver_common.hpp"

using namespace parelagmc;

int main(int argc, char **argv)
{
    try {
        DriverArgs a = DriverArgs::Parse(argc, argv);
        auto hier = std::make_shared<HierarchyData>(HierarchyData::Load(a.hierarchy));
        const int nLevels = hier->nlevels;
        parelag::ParameterList master_list("Default");
        auto &prob = master_list.Sublist("Problem parameters");
        prob.Set("Correlation length", hier->corlen);
        prob.Set("Lognormal", false);   // CreateSamplerParameterList.hpp:33
        auto dev = std::make_shared<B200Device>(a.device, nLevels);
        dev->check(pmc_set_tolerances(dev->handle(), a.rel_tol, a.abs_tol, a.max_iter), "pmc_set_tolerances");
        NormalDistributionSampler dist(0, a.variance, dev);
        dist.Split(1, 0);
        PDESampler sampler(hier, dist, master_list);
        sampler.BuildHierarchy();
        mfem::Vector xi, coef;
        sampler.Sample(0, xi);   // "Realization computation" (:188-194)
        for (int ilevel = 0; ilevel < nLevels; ilevel++) sampler.Eval(ilevel, xi, coef);
        const double exact_expectation = 0.0, exact_variance = a.variance;
        std::vector<double> exp_error(nLevels), var_error(nLevels);
        for (int ilevel = 0; ilevel < nLevels; ++ilevel) {
            const int s_size = sampler.SampleSize(ilevel);
            std::vector<double> expectation(s_size, 0.), marginal_variance(s_size, 0.);
            for (int i = 0; i < a.nsamples; ++i) {
                sampler.Sample(ilevel, xi);
                sampler.Eval(ilevel, xi, coef);
                for (int k = 0; k < s_size; ++k) { expectation[k] += coef(k); marginal_variance[k] += coef(k) * coef(k); }
            }
            const std::vector<double> &vol = hier->sampler[ilevel].Wdiag;
            double e2 = 0., v2 = 0.;
            for (int k = 0; k < s_size; ++k) {
                const double m = expectation[k] / a.nsamples - exact_expectation, v = marginal_variance[k] / a.nsamples - exact_variance;
                e2 += vol[k] * m * m;
                v2 += vol[k] * v * v;
            }
            exp_error[ilevel] = std::sqrt(e2);
            var_error[ilevel] = std::sqrt(v2);
        }
        std::cout << "\nSampler Error: Expected E[u] = " << exact_expectation << ",  Expected V[u] = " << exact_variance << '\n'
                  << "\n L2 Error PDE Sampler \n";
        std::cout << "|| E[u] - Ex ||   || V[u] - Ex ||" << std::endl;
        for (int i = 0; i < nLevels; i++)
            std::cout << std::scientific << std::setprecision(6) << exp_error[i] << "  " << std::setw(17) << var_error[i] << '\n';
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
    }
    return EXIT_SUCCESS;
}
