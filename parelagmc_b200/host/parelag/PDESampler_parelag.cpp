// The -DPARELAGMC_B200_WITH_PARELAG members of PDESampler: the reference's constructor and host-side helpers
// (/root/reference/src/PDESampler.hpp:52-206), written against ParELAG/MFEM.  Hot path: unchanged (PDESampler.cpp).
#ifdef PARELAGMC_B200_WITH_PARELAG
#include <algorithm>
#include <fstream>
#include <iomanip>
#include <sstream>

#include "../PDESampler.hpp"
#include "ParelagExtract.hpp"

namespace parelagmc {

PDESampler::PDESampler(const std::shared_ptr<mfem::ParMesh> &mesh, NormalDistributionSampler &dist_sampler,
                       parelag::ParameterList &master_list)
    : hier_(), dist_sampler_(dist_sampler), prob_list_(master_list.Sublist("Problem parameters", true)),
      lognormal_(prob_list_.Get("Lognormal", true)), corlen_(prob_list_.Get("Correlation length", 0.1)),
      alpha_(1. / (corlen_ * corlen_)), matern_coeff_(ComputeScalingCoefficientForSPDE(corlen_, mesh->Dimension())),
      mesh_(mesh), uform_(mesh->Dimension() - 1), sform_(mesh->Dimension())
{
}

void PDESampler::BuildDeRhamSequence(std::vector<std::shared_ptr<parelag::AgglomeratedTopology>> &topology)
{
    const int feorder = prob_list_.Get("Finite element order", 0), upscaling = prob_list_.Get("Upscaling order", 0);
    sequence_ = BuildSequences(mesh_, topology, feorder, upscaling, nullptr);
}

void PDESampler::SetDeRhamSequence(std::vector<std::shared_ptr<parelag::DeRhamSequence>> &sequence) { sequence_ = sequence; }

// Called by BuildHierarchy() when the object was constructed from a mesh: sequences -> plain arrays (host, once).
void PDESampler::ExtractFromSequences()
{
    const int nLevels = (int)sequence_.size();
    own_hier_ = std::make_shared<HierarchyData>();
    own_hier_->nlevels = nLevels;
    own_hier_->dim = mesh_->Dimension();
    own_hier_->corlen = corlen_;
    own_hier_->sampler.resize(nLevels);
    const int bdr_size = mesh_->bdr_attributes.Size() ? mesh_->bdr_attributes.Max() : 0;
    Ps_.resize(nLevels > 0 ? nLevels - 1 : 0);
    for (int i = 0; i < nLevels; ++i) {
        ExtractSamplerLevel(*sequence_[i], uform_, sform_, i + 1 < nLevels, bdr_size, own_hier_->sampler[i]);
        if (i + 1 < nLevels) Ps_[i] = sequence_[i]->ComputeTrueP(sform_);
    }
    hier_ = own_hier_;
    level_size_.clear();
    out_size_.clear();
    for (const auto &s : hier_->sampler) { level_size_.push_back(s.Ne); out_size_.push_back(s.Ne); }
    nnz_.assign(nLevels, 0);
}

void PDESampler::prolongate_to_fine_grid(int level, const mfem::Vector &coeff, mfem::Vector &fine) const
{
    fine = coeff;
    for (int lev = level; lev > 0; --lev) {          // piecewise-constant interpolation up to level 0
        mfem::Vector help(Ps_[lev - 1]->Height());
        Ps_[lev - 1]->Mult(fine, help);
        fine = help;
    }
}

double PDESampler::ComputeL2Error(int level, const mfem::Vector &coeff, double exact) const
{
    // /root/reference/src/PDESampler.cpp:614-624: the squared L2 error of the prolongated field on the fine mesh
    mfem::Vector fine;
    prolongate_to_fine_grid(level, coeff, fine);
    mfem::GridFunction x;
    x.MakeRef(sequence_[0]->FemSequence()->GetFeSpace(sform_), fine, 0);
    mfem::ConstantCoefficient exact_soln(exact);
    const double err = x.ComputeL2Error(exact_soln);
    return err * err;
}

double PDESampler::ComputeMaxError(int /*level*/, const mfem::Vector &coeff, double exact) const
{
    return std::max(coeff.Max() - exact, exact - coeff.Min());
}

void PDESampler::SaveMeshGLVis(const std::string prefix) const
{
    std::ostringstream name;
    name << prefix << "." << std::setfill('0') << std::setw(6) << 0;
    std::ofstream ofs(name.str().c_str());
    ofs.precision(8);
    mesh_->Print(ofs);
}

void PDESampler::SaveFieldGLVis(int level, const mfem::Vector &coeff, const std::string prefix) const
{
    mfem::Vector fine;
    prolongate_to_fine_grid(level, coeff, fine);
    mfem::GridFunction x;
    x.MakeRef(sequence_[0]->FemSequence()->GetFeSpace(sform_), fine, 0);
    std::ostringstream name;
    name << prefix << "_L" << std::setfill('0') << std::setw(2) << level << "." << std::setw(6) << 0;
    std::ofstream ofs(name.str().c_str());
    ofs.precision(8);
    x.Save(ofs);
}
}  // namespace parelagmc
#endif
