// The -DPARELAGMC_B200_WITH_PARELAG members of DarcySolver: the reference's constructor and host-once set-up calls
// (/root/reference/src/DarcySolver.hpp:40-95), written against ParELAG/MFEM; every one of them ends in plain arrays
// (HierarchyData) that BuildHierachySpaces() -- the upload of DarcySolver.cpp -- hands to the C ABI before the first solve.
#ifdef PARELAGMC_B200_WITH_PARELAG
#include "../DarcySolver.hpp"
#include "ParelagExtract.hpp"

namespace parelagmc {

DarcySolver::DarcySolver(const std::shared_ptr<mfem::ParMesh> &mesh, parelag::ParameterList & /*prec_params*/)
    : hier_(), dev_(B200Device::Default()), mesh_(mesh), uform_(mesh->Dimension() - 1), pform_(mesh->Dimension())
{
}

void DarcySolver::BuildHierachySpaces(std::vector<std::shared_ptr<parelag::AgglomeratedTopology>> &topos,
                                      std::unique_ptr<mfem::BilinearFormIntegrator> massIntegrator)
{
    sequence_ = BuildSequences(mesh_, topos, 0, 0, std::move(massIntegrator));
    const int nLevels = (int)sequence_.size();
    own_hier_ = std::make_shared<HierarchyData>();
    own_hier_->nlevels = nLevels;
    own_hier_->dim = mesh_->Dimension();
    own_hier_->darcy.resize(nLevels);
    for (int i = 0; i < nLevels; ++i) {
        DarcyLevelData &d = own_hier_->darcy[i];
        ExtractDarcyLevel(*sequence_[i], uform_, pform_, i + 1 < nLevels, d);
        const size_t N = (size_t)d.Nf + d.Ne;
        d.ess_u.assign(d.Nf, 0);
        d.ess_data.assign(N, 0.0);
        d.rhs.assign(N, 0.0);
        d.obs.assign(N, 0.0);
    }
    uspace_ = sequence_[0]->FemSequence()->GetFeSpace(uform_);
    pspace_ = sequence_[0]->FemSequence()->GetFeSpace(pform_);
    hier_ = own_hier_;
    nnz_.assign(nLevels, 0);
}

// level i -> i + 1 of a block vector [u; p]: with the transposed prolongators (functionals, right-hand sides:
// src/DarcySolver.cpp:314-315, :410-411) or with the cochain projectors (essential data: :374-375)
void DarcySolver::carry_down(std::vector<double> DarcyLevelData::*field, bool with_projector)
{
    for (int i = 0; i + 1 < own_hier_->nlevels; ++i) {
        DarcyLevelData &f = own_hier_->darcy[i], &c = own_hier_->darcy[i + 1];
        mfem::Vector fu((f.*field).data(), f.Nf), fp((f.*field).data() + f.Nf, f.Ne);
        mfem::Vector cu((c.*field).data(), c.Nf), cp((c.*field).data() + c.Nf, c.Ne);
        if (with_projector) {
            sequence_[i]->GetPi(uform_)->GetProjectorMatrix().Mult(fu, cu);
            sequence_[i]->GetPi(pform_)->GetProjectorMatrix().Mult(fp, cp);
        } else {
            sequence_[i]->GetP(uform_)->MultTranspose(fu, cu);
            sequence_[i]->GetP(pform_)->MultTranspose(fp, cp);
        }
    }
}

void DarcySolver::BuildVolumeObservationFunctional(mfem::LinearFormIntegrator *observationFunctional_u,
                                                   mfem::LinearFormIntegrator *observationFunctional_p)
{
    DarcyLevelData &d = own_hier_->darcy[0];
    mfem::Vector obs0(d.obs.data(), d.Nf + d.Ne);
    mfem::LinearForm fun_u, fun_p;
    fun_u.AddDomainIntegrator(observationFunctional_u);
    fun_u.Update(uspace_, obs0, 0);
    fun_u.Assemble();
    fun_p.AddDomainIntegrator(observationFunctional_p);
    fun_p.Update(pspace_, obs0, d.Nf);
    fun_p.Assemble();
    carry_down(&DarcyLevelData::obs, false);
    built_ = false;
}

void DarcySolver::BuildBdrObservationFunctional(mfem::LinearFormIntegrator *observationFunctional)
{
    DarcyLevelData &d = own_hier_->darcy[0];
    mfem::Vector obs0(d.obs.data(), d.Nf + d.Ne);
    mfem::LinearForm fun_u;
    fun_u.AddBoundaryIntegrator(observationFunctional);
    fun_u.Update(uspace_, obs0, 0);
    fun_u.Assemble();
    carry_down(&DarcyLevelData::obs, false);
    built_ = false;
}

void DarcySolver::SetEssBdrConditions(mfem::Array<int> &ess_bc, mfem::VectorCoefficient &u_bdr)
{
    DarcyLevelData &d0 = own_hier_->darcy[0];
    mfem::Vector ess0(d0.ess_data.data(), d0.Nf + d0.Ne);
    mfem::GridFunction u;
    u.MakeRef(uspace_, ess0, 0);
    u.ProjectBdrCoefficientNormal(u_bdr, ess_bc);
    carry_down(&DarcyLevelData::ess_data, true);
    for (int i = 0; i < own_hier_->nlevels; ++i) {     // the 0/1 mask of the essential RT dofs of every level
        DarcyLevelData &d = own_hier_->darcy[i];
        mfem::Array<int> marker(d.Nf);
        sequence_[i]->GetDofHandler(uform_)->MarkDofsOnSelectedBndr(ess_bc, marker);
        for (int j = 0; j < d.Nf; ++j) d.ess_u[j] = marker[j] ? 1 : 0;
    }
    built_ = false;
}

void DarcySolver::BuildForcingTerms(mfem::VectorCoefficient &f, mfem::Coefficient &p_bdr, mfem::Coefficient &q)
{
    DarcyLevelData &d = own_hier_->darcy[0];
    mfem::Vector rhs0(d.rhs.data(), d.Nf + d.Ne);
    mfem::LinearForm rhs_u, rhs_p;
    rhs_u.AddDomainIntegrator(new mfem::VectorFEDomainLFIntegrator(f));
    rhs_u.AddBoundaryIntegrator(new mfem::VectorFEBoundaryFluxLFIntegrator(p_bdr));
    rhs_u.Update(uspace_, rhs0, 0);
    rhs_u.Assemble();
    rhs_p.AddDomainIntegrator(new mfem::DomainLFIntegrator(q));
    rhs_p.Update(pspace_, rhs0, d.Nf);
    rhs_p.Assemble();
    carry_down(&DarcyLevelData::rhs, false);
    built_ = false;
}
}  // namespace parelagmc
#endif
