// MLMC_reference_api.cpp -- COMPILE-ONLY (make parelag-syntax): the call sequence of the reference's north-star driver
// (/root/reference/examples/MLMC.cpp:163-275) against the host layer built with -DPARELAGMC_B200_WITH_PARELAG, i.e. with
// the reference's own constructor and set-up signatures.  It shows that a driver written for the reference's classes
// compiles against these classes unchanged; it is checked against the declaration-only stubs of host/stubs/ because
// MFEM / ParELAG / MPI are not in this image, and it is never linked or run here.
#include <iostream>
#include <memory>
#include <vector>

#include "../DarcySolver.hpp"
#include "../MLMC_Manager.hpp"
#include "../NormalDistributionSampler.hpp"
#include "../PDESampler.hpp"

using namespace mfem;
using namespace parelag;
using namespace parelagmc;

int reference_style_main(MPI_Comm comm, std::shared_ptr<ParMesh> pmesh,
                         std::vector<std::shared_ptr<AgglomeratedTopology>> &topology, ParameterList &master_list,
                         Array<int> &ess_attr, Array<int> &obs_attr, Array<int> &inflow_attr)
{
    int num_procs, myid;
    MPI_Comm_size(comm, &num_procs);
    MPI_Comm_rank(comm, &myid);
    const int nLevels = (int)topology.size();
    const int nDimensions = pmesh->Dimension();
    const double variance = master_list.Sublist("Problem parameters", true).Get("Variance", 1.0);

    ConstantCoefficient zero(0.), one(1.), minus_one(-1.);
    RestrictedCoefficient obs_coeff(one, obs_attr);
    RestrictedCoefficient pinflow_coeff(minus_one, inflow_attr);
    Vector zeros_nDim(nDimensions);
    zeros_nDim = 0.;
    VectorConstantCoefficient zero_vcoeff(zeros_nDim);

    DarcySolver solver(pmesh, master_list);
    solver.BuildHierachySpaces(topology, make_unique<VectorFEMassIntegrator>(one));
    solver.BuildBdrObservationFunctional(new VectorFEBoundaryFluxLFIntegrator(obs_coeff));
    solver.SetEssBdrConditions(ess_attr, zero_vcoeff);
    solver.BuildForcingTerms(zero_vcoeff, pinflow_coeff, zero);

    NormalDistributionSampler dist(0, variance);
    dist.Split(num_procs, myid);

    PDESampler sampler(pmesh, dist, master_list);
    sampler.SetDeRhamSequence(solver.GetSequence());
    sampler.BuildHierarchy();

    MLMC_Manager mlmc(comm, nLevels, solver, sampler, master_list);
    mlmc.Run();

    // the per-sample interface, as the test drivers use it (/root/reference/examples/PDESamplerTest.cpp:188-232)
    Vector xi, coef;
    double Q, C;
    sampler.Sample(0, xi);
    sampler.Eval(0, xi, coef);
    solver.SolveFwd(0, coef, Q, C);
    const double err = sampler.ComputeL2Error(0, coef, 0.0);
    if (myid == 0) std::cout << Q << " " << C << " " << err << " " << sampler.GetNNZ(0) << " " << solver.GetNNZ(0) << std::endl;
    return 0;
}
