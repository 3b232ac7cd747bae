// DarcySolver.hpp -- mixed Darcy forward solver with the public interface of
// /root/reference/src/DarcySolver.hpp:36-169 for the methods on the hot path.
#pragma once
#include <memory>
#include <vector>
#include "B200Device.hpp"
#include "HierarchyData.hpp"
#include "PhysicalMLSolver.hpp"

namespace parelagmc {
class DarcySolver : public PhysicalMLSolver {
public:
    DarcySolver(std::shared_ptr<const HierarchyData> hier, std::shared_ptr<B200Device> dev,
                parelag::ParameterList &master_list);
#ifdef PARELAGMC_B200_WITH_PARELAG
    /// The reference's constructor and set-up calls, verbatim (/root/reference/src/DarcySolver.hpp:40-42, :55-95): the
    /// hierarchy is built by ParELAG on the host, once; the last set-up call made before the first solve uploads it.
    DarcySolver(const std::shared_ptr<mfem::ParMesh> &mesh, parelag::ParameterList &prec_params);
    void BuildHierachySpaces(std::vector<std::shared_ptr<parelag::AgglomeratedTopology>> &topos,
                             std::unique_ptr<mfem::BilinearFormIntegrator> massIntegrator);
    void BuildVolumeObservationFunctional(mfem::LinearFormIntegrator *observationFunctional_u,
                                          mfem::LinearFormIntegrator *observationFunctional_p);
    void BuildBdrObservationFunctional(mfem::LinearFormIntegrator *observationFunctional);
    void SetEssBdrConditions(mfem::Array<int> &ess_bc, mfem::VectorCoefficient &u_bdr);
    void BuildForcingTerms(mfem::VectorCoefficient &f, mfem::Coefficient &p_bdr, mfem::Coefficient &q);
    std::vector<std::shared_ptr<parelag::DeRhamSequence>> &GetSequence() override { return sequence_; }
    mfem::FiniteElementSpace *GetPressureSpace() const override { return pspace_; }
    mfem::FiniteElementSpace *GetVelocitySpace() const { return uspace_; }
#endif
    virtual ~DarcySolver() = default;
    DarcySolver(DarcySolver const &) = delete;
    DarcySolver &operator=(DarcySolver const &) = delete;

    /// Uploads element blocks, B, boundary data, rhs and the observation functional
    /// (what BuildHierachySpaces .. BuildForcingTerms leave behind, src/DarcySolver.cpp:60-414).
    void BuildHierachySpaces();
    void SolveFwd(int ilevel, mfem::Vector &k_over_k_ref, double &Q, double &C) override;
    void SolveFwd_RtnPressure(int ilevel, mfem::Vector &k_over_k_ref, mfem::Vector &P, double &C, double &Q,
                              bool compute_Q) override;
    int GetNumberOfDofs(int ilevel) const override { return hier_->darcy[ilevel].Ne + hier_->darcy[ilevel].Nf; }
    int GetGlobalNumberOfDofs(int ilevel) const override { return GetNumberOfDofs(ilevel); }
    int GetNNZ(int ilevel) const override { return nnz_[ilevel]; }
    int GetSizeOfStochasticData(int ilevel) const { return hier_->darcy[ilevel].Ne; }
    int GetNumIters() const { return -1; }  // as the reference (src/DarcySolver.hpp:103-107)
    const std::shared_ptr<B200Device> &Device() const { return dev_; }

private:
    std::shared_ptr<const HierarchyData> hier_;
    std::shared_ptr<B200Device> dev_;
    std::vector<int> nnz_;
    bool built_ = false;
#ifdef PARELAGMC_B200_WITH_PARELAG
    void carry_down(std::vector<double> DarcyLevelData::*field, bool with_projector);
    std::shared_ptr<mfem::ParMesh> mesh_;
    std::vector<std::shared_ptr<parelag::DeRhamSequence>> sequence_;
    std::shared_ptr<HierarchyData> own_hier_;
    mfem::FiniteElementSpace *uspace_ = nullptr, *pspace_ = nullptr;
    int uform_ = 0, pform_ = 0;
#endif
};
}  // namespace parelagmc
