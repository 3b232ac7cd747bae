// BayesianInverseProblem.hpp -- the public interface of /root/reference/src/BayesianInverseProblem.hpp:29-120 on the host
// layer: observational data G_obs = G(k(xi_0)) + eta, the likelihood exp(-|G(k) - G_obs|^2 / (2 noise)) and R = Q *
// likelihood for a prior realisation.  The per-sample methods go through MLSampler / PhysicalMLSolver (one device call
// each); UploadObservations() hands the functionals and G_obs to the device so that the ratio-estimator managers can run
// their level loops batched (pmc_bayes_level_batch).  The observation functionals g_obs_func[i][level] -- which the
// reference builds from the mesh (ChangeMeshAttributes + DomainLFIntegrator, src/BayesianInverseProblem.cpp:46-104) --
// come with the hierarchy data (host-once set-up).
#pragma once
#include <memory>
#include <vector>
#include "DarcySolver.hpp"
#include "HierarchyData.hpp"
#include "MLSampler.hpp"
#include "NormalDistributionSampler.hpp"
#include "PhysicalMLSolver.hpp"

namespace parelagmc {
class BayesianInverseProblem {
public:
    BayesianInverseProblem(std::shared_ptr<const HierarchyData> hier, PhysicalMLSolver &solver, MLSampler &prior,
                           parelag::ParameterList &master_list);
    BayesianInverseProblem(BayesianInverseProblem const &) = delete;
    BayesianInverseProblem &operator=(BayesianInverseProblem const &) = delete;

    /// G_obs = G(u) + N(0, noise) with u the first prior realisation on level 0 (src/BayesianInverseProblem.cpp:130-175)
    void GenerateObservationalData();
    void SamplePrior(int ilevel, mfem::Vector &xi) { prior.Sample(ilevel, xi); }
    void EvalPrior(int ilevel, const mfem::Vector &xi, mfem::Vector &u) { prior.Eval(ilevel, xi, u); }
    /// G[i] = g_obs_func[i][level] . p / sum(g_obs_func[i][level])  (:178-192)
    void ComputeG(int ilevel, mfem::Vector &k_over_k_ref, mfem::Vector &G, double &C, double &Q, bool compute_Q);
    void ComputeLikelihood(int ilevel, mfem::Vector &k_over_k_ref, double &likelihood, double &C);
    void ComputeLikelihoodAndQ(int ilevel, mfem::Vector &k_over_k_ref, double &likelihood, double &C, double &Q);
    void ComputeR(int ilevel, mfem::Vector &k_over_k_ref, double &R, double &C);
    PhysicalMLSolver &GetSolver() { return solver; }
    MLSampler &GetPrior() { return prior; }
    const mfem::Vector &ObservationalData() const { return G_obs; }
    void SetObservationalData(const mfem::Vector &g) { G_obs = g; }
    double Noise() const { return noise; }
    /// Hand functionals, G_obs and the noise variance to the device (needed by the batched ratio-estimator loops).
    void UploadObservations(B200Device &dev);

private:
    std::shared_ptr<const HierarchyData> hier;
    PhysicalMLSolver &solver;
    MLSampler &prior;
    parelag::ParameterList &bayesian_list;
    const double noise;
    const int size_obs_data;
    mfem::Vector G_obs, Gl;
    double c = 0, q = 0;
};
}  // namespace parelagmc
