#include "DarcySolver.hpp"

namespace parelagmc {

DarcySolver::DarcySolver(std::shared_ptr<const HierarchyData> hier, std::shared_ptr<B200Device> dev,
                         parelag::ParameterList & /*master_list*/)
    : hier_(std::move(hier)), dev_(std::move(dev))
{
    nnz_.assign(hier_->nlevels, 0);
}

void DarcySolver::BuildHierachySpaces()
{
    if (built_) return;
    for (int l = 0; l < hier_->nlevels; ++l) {
        const DarcyLevelData &d = hier_->darcy[l];
        const bool hasP = !d.Pp.empty();
        dev_->check(pmc_upload_darcy_level(dev_->handle(), l, d.Ne, d.Nf, d.elem_ptr.data(), d.elem_dofs.data(),
                                           d.elem_mat.data(), d.B.rowptr.data(), d.B.col.data(), d.B.val.data(),
                                           d.ess_u.data(), d.ess_data.data(), d.rhs.data(), d.obs.data(),
                                           hasP ? d.Pp.cols : 0, hasP ? d.Pp.rowptr.data() : nullptr,
                                           hasP ? d.Pp.col.data() : nullptr, hasP ? d.Pp.val.data() : nullptr),
                    "pmc_upload_darcy_level");
        // nnz of [[M, B^T], [B, 0]] as the reference counts it after assemble (src/DarcySolver.cpp:510): pattern of
        // M from the element dof lists + 2 nnz(B)
        size_t m = 0;
        for (int e = 0; e < d.Ne; ++e) {
            const size_t n = (size_t)(d.elem_ptr[e + 1] - d.elem_ptr[e]);
            m += n * n;
        }
        nnz_[l] = (int)(m + 2 * d.B.nnz());
    }
    built_ = true;
}

void DarcySolver::SolveFwd(int ilevel, mfem::Vector &k_over_k_ref, double &Q, double &C)
{
    // assemble -> solve -> Q = obs . sol, C = #dofs (src/DarcySolver.cpp:416-437)
    if (!built_) BuildHierachySpaces();   // reference-signature set-up: upload after the last Build* / Set* call
    if (k_over_k_ref.Size() != hier_->darcy[ilevel].Ne)
        throw std::runtime_error("DarcySolver::SolveFwd: coefficient vector has the wrong size");
    dev_->check(pmc_darcy_solve_batch(dev_->handle(), ilevel, 1, k_over_k_ref.GetData(), &Q, &C, nullptr, nullptr),
                "pmc_darcy_solve_batch");
}

void DarcySolver::SolveFwd_RtnPressure(int ilevel, mfem::Vector &k_over_k_ref, mfem::Vector &P, double &C, double &Q,
                                       bool compute_Q)
{
    // src/DarcySolver.cpp:439-470: same solve, also returns the pressure block
    if (!built_) BuildHierachySpaces();
    const DarcyLevelData &d = hier_->darcy[ilevel];
    if (k_over_k_ref.Size() != d.Ne) throw std::runtime_error("DarcySolver::SolveFwd_RtnPressure: wrong size");
    mfem::Vector sol(d.Nf + d.Ne);
    double q = 0.0;
    dev_->check(pmc_darcy_solve_batch(dev_->handle(), ilevel, 1, k_over_k_ref.GetData(), &q, &C, sol.GetData(), nullptr),
                "pmc_darcy_solve_batch");
    P.SetSize(d.Ne);
    for (int i = 0; i < d.Ne; ++i) P(i) = sol(d.Nf + i);
    if (compute_Q) Q = q;
}
}  // namespace parelagmc
