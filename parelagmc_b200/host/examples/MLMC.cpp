// MLMC.cpp -- the north-star driver (/root/reference/examples/MLMC.cpp:43-283) on the B200 host layer: hierarchy ->
// DarcySolver + NormalDistributionSampler + PDESampler -> MLMC_Manager.Run().  The mesh/topology/DeRham part of the
// reference's main() (lines 163-239) is replaced by loading the already-built hierarchy data.
//   MLMC.exe --hierarchy FILE [--samples n0,n1,..] [--mse X] [--rel-tol X] [--dof-cost] [--log FILE]
#include <iostream>
#include <memory>

#include "../DarcySolver.hpp"
#include "../MLMC_Manager.hpp"
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
        prob.Set("Lognormal", true);
        prob.Set("Correlation length", hier->corlen);
        prob.Set("Mean square error", a.mse);
        prob.Set("Number of samples", a.nsamples);
        prob.Set("Output filename for MC managers", a.log);
        if (!a.samples.empty()) {
            prob.Set("Use array samples", true);
            prob.Set("Array number of samples", a.samples);
        }
        auto dev = std::make_shared<B200Device>(a.device, nLevels);
        dev->check(pmc_set_tolerances(dev->handle(), a.rel_tol, a.abs_tol, a.max_iter), "pmc_set_tolerances");

        DarcySolver solver(hier, dev, master_list);
        solver.BuildHierachySpaces();
        NormalDistributionSampler dist(0, a.variance, dev);
        dist.Split(1, 0);  // dist.Split(num_procs, myid) (reference :242)
        PDESampler sampler(hier, dist, master_list);
        sampler.BuildHierarchy();

        MLMC_Manager mlmc(MPI_COMM_WORLD, nLevels, solver, sampler, master_list);
        mlmc.wallTime = a.wall_time;
        mlmc.Run();
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;  // as the reference: report and return success (:277-282)
    }
    return EXIT_SUCCESS;
}
