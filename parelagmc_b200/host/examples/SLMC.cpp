// SLMC.cpp -- single-level driver (/root/reference/examples/SLMC.cpp) on the B200 host layer with the plain
// PDESampler: hierarchy -> DarcySolver + PDESampler -> MC_Manager.Run().
#include <iostream>
#include <memory>

#include "../DarcySolver.hpp"
#include "../MC_Manager.hpp"
#include "../NormalDistributionSampler.hpp"
#include "../PDESampler.hpp"
#include "driver_common.hpp"

using namespace parelagmc;

int main(int argc, char **argv)
{
    try {
        DriverArgs a = DriverArgs::Parse(argc, argv);
        auto hier = std::make_shared<HierarchyData>(HierarchyData::Load(a.hierarchy));
        parelag::ParameterList master_list("Default");
        auto &prob = master_list.Sublist("Problem parameters");
        prob.Set("Lognormal", true);
        prob.Set("Correlation length", hier->corlen);
        prob.Set("Mean square error", a.mse);
        prob.Set("Number of samples", a.nsamples);
        prob.Set("Output filename for MC managers", a.log);
        auto dev = std::make_shared<B200Device>(a.device, hier->nlevels);
        dev->check(pmc_set_tolerances(dev->handle(), a.rel_tol, a.abs_tol, a.max_iter), "pmc_set_tolerances");
        DarcySolver solver(hier, dev, master_list);
        solver.BuildHierachySpaces();
        NormalDistributionSampler dist(0, a.variance, dev);
        PDESampler sampler(hier, dist, master_list);
        sampler.BuildHierarchy();
        MC_Manager mc(MPI_COMM_WORLD, solver, sampler, master_list);
        mc.wallTime = a.wall_time;
        mc.Run();
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
    }
    return EXIT_SUCCESS;
}
