#include "BayesianInverseProblem.hpp"

#include <cmath>
#include <stdexcept>

#include "PDESampler.hpp"

namespace parelagmc {

BayesianInverseProblem::BayesianInverseProblem(std::shared_ptr<const HierarchyData> hier_, PhysicalMLSolver &solver_,
                                               MLSampler &prior_, parelag::ParameterList &master_list)
    : hier(std::move(hier_)), solver(solver_), prior(prior_),
      bayesian_list(master_list.Sublist("Bayesian inverse problem parameters", true)),
      noise(bayesian_list.Get("Noise", 0.1)), size_obs_data(hier->n_obs)
{
    if (size_obs_data < 1) throw std::runtime_error("BayesianInverseProblem: the hierarchy data carries no observation functionals");
}

void BayesianInverseProblem::ComputeG(int ilevel, mfem::Vector &k_over_k_ref, mfem::Vector &G, double &C, double &Q,
                                      bool compute_Q)
{
    mfem::Vector p;
    solver.SolveFwd_RtnPressure(ilevel, k_over_k_ref, p, C, Q, compute_Q);
    const int Ne = p.Size();
    G.SetSize(size_obs_data);
    for (int i = 0; i < size_obs_data; ++i) {
        const double *g = hier->gobs[ilevel].data() + (size_t)i * Ne;
        double dot = 0., sum = 0.;
        for (int e = 0; e < Ne; ++e) { dot += g[e] * p(e); sum += g[e]; }
        G(i) = dot / sum;
    }
}

void BayesianInverseProblem::GenerateObservationalData()
{
    // "Generate reference observational data": the first prior draw of the run, then eta from a SECOND, default-seeded
    // NormalDistributionSampler(0, noise) (src/BayesianInverseProblem.cpp:158-175)
    mfem::Vector xi, u, eta;
    prior.Sample(0, xi);
    prior.Eval(0, xi, u);
    ComputeG(0, u, G_obs, c, q, false);
    PDESampler *ps = dynamic_cast<PDESampler *>(&prior);
    if (!ps) throw std::runtime_error("BayesianInverseProblem: the prior must be a device sampler");
    eta.SetSize(G_obs.Size());
    {
        NormalDistributionSampler noise_dist(0, noise, ps->Device());
        noise_dist(eta);
        ps->Device()->rng_owner = nullptr;   // noise_dist dies here: the prior re-binds at its next use
    }
    for (int i = 0; i < G_obs.Size(); ++i) G_obs(i) += eta(i);
}

void BayesianInverseProblem::ComputeLikelihood(int ilevel, mfem::Vector &k_over_k_ref, double &likelihood, double &C)
{
    ComputeG(ilevel, k_over_k_ref, Gl, C, q, false);
    double n2 = 0.;
    for (int i = 0; i < Gl.Size(); ++i) n2 += (Gl(i) - G_obs(i)) * (Gl(i) - G_obs(i));
    likelihood = std::exp((-1. / (noise * 2)) * n2);   // :199
}

void BayesianInverseProblem::ComputeLikelihoodAndQ(int ilevel, mfem::Vector &k_over_k_ref, double &likelihood, double &C,
                                                   double &Q)
{
    ComputeG(ilevel, k_over_k_ref, Gl, C, Q, true);
    double n2 = 0.;
    for (int i = 0; i < Gl.Size(); ++i) n2 += (Gl(i) - G_obs(i)) * (Gl(i) - G_obs(i));
    likelihood = std::exp((-1. / (noise * 2)) * n2);
}

void BayesianInverseProblem::ComputeR(int ilevel, mfem::Vector &k_over_k_ref, double &R, double &C)
{
    ComputeLikelihoodAndQ(ilevel, k_over_k_ref, R, C, q);
    R *= q;   // :217
}

void BayesianInverseProblem::UploadObservations(B200Device &dev)
{
    for (int l = 0; l < hier->nlevels; ++l)
        dev.check(pmc_upload_observations(dev.handle(), l, size_obs_data, hier->gobs[l].data(), G_obs.GetData(), noise),
                  "pmc_upload_observations");
}
}  // namespace parelagmc
