// DarcyTest.cpp -- deterministic solve with k == 1 on every level, printing the table the reference's ctest
// `DarcyDeterministicTest` matches (/root/reference/examples/DarcyTest.cpp:232-254, examples/CMakeLists.txt:62-66).
// Also exercises the per-sample interface: Sample / Eval (both overloads) / SolveFwd for one realisation per level.
#include <iomanip>
#include <iostream>
#include <memory>
#include <sstream>

#include "../DarcySolver.hpp"
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
        master_list.Sublist("Problem parameters").Set("Correlation length", hier->corlen);
        auto dev = std::make_shared<B200Device>(a.device, nLevels);
        dev->check(pmc_set_tolerances(dev->handle(), a.rel_tol, a.abs_tol, a.max_iter), "pmc_set_tolerances");
        DarcySolver solver(hier, dev, master_list);
        solver.BuildHierachySpaces();
        std::stringstream msg;
        msg << "\nL  " << std::setw(8) << std::left << "QoI  " << std::setw(8) << std::left << "  Cost (dofs)\n";
        for (int ilevel = 0; ilevel < nLevels; ++ilevel) {
            mfem::Vector ones(solver.GetSizeOfStochasticData(ilevel));
            ones = 1.;
            double Q, C;
            solver.SolveFwd(ilevel, ones, Q, C);
            msg << ilevel << "  " << std::setw(8) << std::left << Q << "  " << std::setw(8) << std::left << C << "\n";
        }
        std::cout << msg.str();
        // one random realisation per level through the per-sample interface (DarcyTest_RandomInput style)
        NormalDistributionSampler dist(0, a.variance, dev);
        PDESampler sampler(hier, dist, master_list);
        sampler.BuildHierarchy();
        std::cout << "\nL  " << std::setw(12) << std::left << "QoI(xi)" << std::setw(12) << std::left << "QoI_c(xi)" << "\n";
        for (int ilevel = nLevels - 1; ilevel >= 0; --ilevel) {
            mfem::Vector xi, s, init_s;
            double q = 0, qc = 0, c;
            sampler.Sample(ilevel, xi);
            if (ilevel == nLevels - 1) {
                sampler.Eval(ilevel, xi, s);
                solver.SolveFwd(ilevel, s, q, c);
            } else {
                sampler.Eval(ilevel + 1, xi, s, init_s, false);
                solver.SolveFwd(ilevel + 1, s, qc, c);
                sampler.Eval(ilevel, xi, s, init_s, true);
                solver.SolveFwd(ilevel, s, q, c);
            }
            std::cout << ilevel << "  " << std::setprecision(10) << std::setw(12) << std::left << q << "  " << std::setw(12)
                      << std::left << qc << "\n";
        }
    } catch (std::exception &e) {
        std::cout << e.what() << std::endl;
    }
    return EXIT_SUCCESS;
}
