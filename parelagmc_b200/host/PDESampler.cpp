#include "PDESampler.hpp"

#include <cmath>
#include <iostream>

namespace parelagmc {

double ComputeScalingCoefficientForSPDE(double corlen, int myDim)
{
    // /root/reference/src/Utilities.hpp:188-200 (tgamma(nu + d), as the code does)
    const double dim = static_cast<double>(myDim);
    const double nu = 2. - dim / 2.;
    const double gnu = std::tgamma(nu), gnudim = std::tgamma(nu + dim);
    const double c = std::pow(16. * std::atan(1.), 0.5 * dim);
    const double k = std::pow(1. / corlen, 2. * nu);
    return std::sqrt(c * gnudim * k / gnu);
}

PDESampler::PDESampler(std::shared_ptr<const HierarchyData> hier, NormalDistributionSampler &dist_sampler,
                       parelag::ParameterList &master_list)
    : hier_(std::move(hier)), dist_sampler_(dist_sampler),
      prob_list_(master_list.Sublist("Problem parameters", true)),
      lognormal_(prob_list_.Get("Lognormal", true)),              // default true (src/PDESampler.cpp:38)
      corlen_(prob_list_.Get("Correlation length", 0.1)),         // :41
      alpha_(1. / (corlen_ * corlen_)),                           // :42
      matern_coeff_(ComputeScalingCoefficientForSPDE(corlen_, hier_->dim))
{
    for (const auto &s : hier_->sampler) { level_size_.push_back(s.Ne); out_size_.push_back(s.Ne); }
    nnz_.assign(hier_->nlevels, 0);
}

void PDESampler::BuildHierarchy()
{
    if (built_) return;
#ifdef PARELAGMC_B200_WITH_PARELAG
    if (!hier_ && mesh_) ExtractFromSequences();   // constructed from a mesh: ParELAG's hierarchy -> plain arrays
#endif
    auto &dev = *Device();
    for (int l = 0; l < hier_->nlevels; ++l) {
        const SamplerLevelData &s = hier_->sampler[l];
        const bool hasP = !s.P.empty();
        dev.check(pmc_upload_sampler_level(dev.handle(), l, s.Ne, s.Nf, s.M.rowptr.data(), s.M.col.data(), s.M.val.data(),
                                           s.B.rowptr.data(), s.B.col.data(), s.B.val.data(), s.Wdiag.data(),
                                           hasP ? s.P.cols : 0, hasP ? s.P.rowptr.data() : nullptr,
                                           hasP ? s.P.col.data() : nullptr, hasP ? s.P.val.data() : nullptr, alpha_,
                                           matern_coeff_, lognormal_ ? 1 : 0),
                  "pmc_upload_sampler_level");
        nnz_[l] = s.M.nnz() + 2 * s.B.nnz() + (size_t)s.Ne;  // pM + pB + pBt + pW (src/PDESampler.cpp:260-265)
        if (!s.T.empty()) {
            // enlarged-domain variants: meshP (src/EmbeddedPDESampler.cpp:63-89) or Gt + 1/diag(W_orig)
            // (src/L2ProjectionPDESampler.cpp:488-514)
            dev.check(pmc_upload_field_transfer(dev.handle(), l, s.T.rows, s.T.rowptr.data(), s.T.col.data(), s.T.val.data(),
                                                s.Tscale.empty() ? nullptr : s.Tscale.data()),
                      "pmc_upload_field_transfer");
            out_size_[l] = s.T.rows;
        }
    }
    built_ = true;
}

int PDESampler::FindLevel(int size) const
{
    for (int l = 0; l < (int)level_size_.size(); ++l)
        if (level_size_[l] == size) return l;
    throw std::runtime_error("PDESampler: vector length does not match any level size");
}

void PDESampler::Sample(const int level, mfem::Vector &xi)
{
    xi.SetSize(level_size_[level]);  // src/PDESampler.cpp:336-340
    dist_sampler_(xi);
}

void PDESampler::Eval(const int level, const mfem::Vector &xi, mfem::Vector &s)
{
    const int xi_level = FindLevel(xi.Size());  // :349
    if (xi_level > level) throw std::runtime_error("PDESampler::Eval: noise coarser than the evaluation level");
    s.SetSize(out_size_[level]);
    auto &dev = *Device();
    dev.check(pmc_sampler_eval_batch(dev.handle(), level, xi_level, 1, xi.GetData(), nullptr, 0, -1, s.GetData(),
                                     nullptr, nullptr),
              "pmc_sampler_eval_batch");
}

void PDESampler::Eval(const int level, const mfem::Vector &xi, mfem::Vector &s, mfem::Vector &embed_s, bool use_init)
{
    const int xi_level = FindLevel(xi.Size());  // :419
    if (xi_level > level) throw std::runtime_error("PDESampler::Eval: noise coarser than the evaluation level");
    int init_level = 0;
    mfem::Vector init;
    if (use_init) {  // embed_s holds the coarser Gaussian field; its level is inferred from its length (:498)
        init_level = FindLevel(embed_s.Size());
        init = embed_s;
    }
    s.SetSize(out_size_[level]);
    embed_s.SetSize(level_size_[level]);
    auto &dev = *Device();
    dev.check(pmc_sampler_eval_batch(dev.handle(), level, xi_level, 1, xi.GetData(), use_init ? init.GetData() : nullptr,
                                     init_level, use_init ? 1 : 0, s.GetData(), embed_s.GetData(), nullptr),
              "pmc_sampler_eval_batch");
}
}  // namespace parelagmc
