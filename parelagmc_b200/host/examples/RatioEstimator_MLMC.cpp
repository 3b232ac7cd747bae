// RatioEstimator_MLMC.cpp -- /root/reference/examples/RatioEstimator_MLMC_Manager.cpp on the host layer: enlarged-mesh
// sampler + DarcySolver + BayesianInverseProblem -> ML_BayesRatio_Manager.Run() (BASELINE configs[3]).
//   RatioEstimator_MLMC.exe --hierarchy FILE [--samples n0,n1,..] [--mse X] [--dof-cost] [--per-sample]
// --per-sample runs the reference's per-sample loop through BayesianInverseProblem instead of the batched device call
// (same stream positions, same sums up to round-off).
#include <cstring>
#include <iostream>
#include <memory>

#include "../ML_BayesRatio_Manager.hpp"
#include "../PDESampler.hpp"
#include "driver_common.hpp"

using namespace parelagmc;

namespace {
// hides the device classes behind the abstract interfaces: the manager then takes the reference's per-sample loop
struct OpaqueSampler : MLSampler {
    explicit OpaqueSampler(PDESampler &s) : s_(s) {}
    void Sample(const int l, mfem::Vector &xi) override { s_.Sample(l, xi); }
    void Eval(const int l, const mfem::Vector &xi, mfem::Vector &s) override { s_.Eval(l, xi, s); }
    void Eval(const int l, const mfem::Vector &xi, mfem::Vector &s, mfem::Vector &u, bool i) override { s_.Eval(l, xi, s, u, i); }
    int SampleSize(int l) const override { return s_.SampleSize(l); }
    size_t GetNNZ(int l) const override { return s_.GetNNZ(l); }
    void BuildHierarchy() override {}
    PDESampler &s_;
};
}  // namespace

int main(int argc, char **argv)
{
    try {
        DriverArgs a = DriverArgs::Parse(argc, argv);
        bool per_sample = false;
        for (int i = 1; i < argc; ++i) per_sample |= !strcmp(argv[i], "--per-sample");
        auto hier = std::make_shared<HierarchyData>(HierarchyData::Load(a.hierarchy));
        const int nLevels = hier->nlevels;
        parelag::ParameterList master_list("Default");
        auto &prob = master_list.Sublist("Problem parameters");
        prob.Set("Correlation length", hier->corlen);
        prob.Set("Lognormal", true);
        prob.Set("Mean square error", a.mse);
        prob.Set("Number of samples", a.nsamples);
        if (!a.samples.empty()) prob.Set("Array number of samples", a.samples);
        master_list.Sublist("Bayesian inverse problem parameters").Set("Noise", 0.1);
        auto dev = std::make_shared<B200Device>(a.device, nLevels);
        dev->check(pmc_set_tolerances(dev->handle(), a.rel_tol, a.abs_tol, a.max_iter), "pmc_set_tolerances");
        DarcySolver solver(hier, dev, master_list);
        solver.BuildHierachySpaces();
        NormalDistributionSampler dist(0, a.variance, dev);
        dist.Split(1, 0);
        L2ProjectionPDESampler sampler(hier, dist, master_list);
        sampler.BuildHierarchy();
        OpaqueSampler opaque(sampler);
        MLSampler &prior = per_sample ? static_cast<MLSampler &>(opaque) : static_cast<MLSampler &>(sampler);
        BayesianInverseProblem problem(hier, solver, sampler, master_list);
        problem.GenerateObservationalData();
        problem.UploadObservations(*dev);
        BayesianInverseProblem problem_ps(hier, solver, prior, master_list);   // same data, opaque prior when --per-sample
        ML_BayesRatio_Manager mgr(MPI_COMM_WORLD, nLevels, per_sample ? problem_ps : problem, master_list);
        if (per_sample) problem_ps.SetObservationalData(problem.ObservationalData());
        mgr.wallTime = a.wall_time;
        mgr.Run();
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
    }
    return EXIT_SUCCESS;
}
