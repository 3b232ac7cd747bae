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
};
}  // namespace parelagmc
