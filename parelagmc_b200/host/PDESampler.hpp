// PDESampler.hpp -- SPDE sampler with the public interface of /root/reference/src/PDESampler.hpp:45-206 for the
// methods on the hot path; the saddle solve [M B^T; B -alpha W] runs batched on the GPU through include/pmc_b200.h.
#pragma once
#include <memory>
#include <vector>
#include "HierarchyData.hpp"
#include "MLSampler.hpp"
#include "NormalDistributionSampler.hpp"

namespace parelagmc {
class PDESampler : public MLSampler {
public:
    /// The reference takes the ParMesh (src/PDESampler.hpp:52-55) and builds the hierarchy with ParELAG; here the
    /// already-built hierarchy data is handed over (see INTEGRATION.md for the ParELAG-side extraction).
    PDESampler(std::shared_ptr<const HierarchyData> hier, NormalDistributionSampler &dist_sampler,
               parelag::ParameterList &master_list);
#ifdef PARELAGMC_B200_WITH_PARELAG
    /// The reference's constructor, verbatim (/root/reference/src/PDESampler.hpp:52-55).  BuildDeRhamSequence /
    /// SetDeRhamSequence then BuildHierarchy as in the reference's drivers; the hierarchy data is extracted from the
    /// DeRham sequences (parelag/ParelagExtract.hpp) at BuildHierarchy.
    PDESampler(const std::shared_ptr<mfem::ParMesh> &mesh_, NormalDistributionSampler &dist_sampler,
               parelag::ParameterList &master_list);
    void BuildDeRhamSequence(std::vector<std::shared_ptr<parelag::AgglomeratedTopology>> &topology) override;
    void SetDeRhamSequence(std::vector<std::shared_ptr<parelag::DeRhamSequence>> &sequence) override;
    mfem::HypreParMatrix *GetTrueP(int level) override { return Ps_[level].get(); }
    void SaveMeshGLVis(const std::string prefix) const override;
    void SaveFieldGLVis(int level, const mfem::Vector &coeff, const std::string prefix) const override;
    double ComputeL2Error(int level, const mfem::Vector &coeff, double exact) const override;
    double ComputeMaxError(int level, const mfem::Vector &coeff, double exact) const override;
    int GetGlobalNumberOfDofs(int level) const { return GetNumberOfDofs(level); }
#endif
    virtual ~PDESampler() = default;
    PDESampler(PDESampler const &) = delete;
    PDESampler &operator=(PDESampler const &) = delete;

    /// Uploads the per-level operators (end of src/PDESampler.cpp:177-334).
    void BuildHierarchy() override;
    void Sample(const int level, mfem::Vector &xi) override;
    void Eval(const int level, const mfem::Vector &xi, mfem::Vector &s) override;
    void Eval(const int level, const mfem::Vector &xi, mfem::Vector &s, mfem::Vector &u, bool use_init) override;
    int SampleSize(int level) const override { return level_size_[level]; }
    int GlobalSampleSize(int level) const { return level_size_[level]; }
    size_t GetNNZ(int level) const override { return nnz_[level]; }
    int GetNumberOfDofs(int level) const { return hier_->sampler[level].Ne + hier_->sampler[level].Nf; }
    int GetNumIters() const { return -1; }  // as the reference (src/PDESampler.hpp:141-145)

    NormalDistributionSampler &Distribution() { return dist_sampler_; }
    const std::shared_ptr<B200Device> &Device() const { return dist_sampler_.Device(); }
    bool Lognormal() const { return lognormal_; }
    int NoiseSize(int level) const { return level_size_[level]; }
    int OutputSize(int level) const { return out_size_[level]; }

private:
    int FindLevel(int size) const;  // level_size.Find(xi.Size()) (src/PDESampler.cpp:349,419)
    std::shared_ptr<const HierarchyData> hier_;
    NormalDistributionSampler &dist_sampler_;
    parelag::ParameterList &prob_list_;
    bool lognormal_;
    double corlen_, alpha_, matern_coeff_;
    std::vector<int> level_size_;   // noise / Gaussian-field size per level (the enlarged mesh for the variants below)
    std::vector<int> out_size_;     // size of the field handed to the forward solver
    std::vector<size_t> nnz_;
    bool built_ = false;
#ifdef PARELAGMC_B200_WITH_PARELAG
    void prolongate_to_fine_grid(int level, const mfem::Vector &coeff, mfem::Vector &fine) const;
    void ExtractFromSequences();
    std::shared_ptr<mfem::ParMesh> mesh_;
    std::vector<std::shared_ptr<parelag::DeRhamSequence>> sequence_;
    std::vector<std::unique_ptr<mfem::HypreParMatrix>> Ps_;
    std::shared_ptr<HierarchyData> own_hier_;
    int uform_ = 0, sform_ = 0;
#endif
};

/// EmbeddedPDESampler (/root/reference/src/EmbeddedPDESampler.hpp:46): the SPDE is solved on an enlarged MATCHING mesh
/// and the field is restricted to the original elements with the 0/1 selection meshP.  Same machinery as PDESampler;
/// the transfer comes with the hierarchy data.  As in the reference, Sample draws the enlarged size while SampleSize
/// reports the original-mesh size (src/EmbeddedPDESampler.hpp:126-129).
class EmbeddedPDESampler : public PDESampler {
public:
    using PDESampler::PDESampler;
    int SampleSize(int level) const override { return OutputSize(level); }
};

/// L2ProjectionPDESampler (/root/reference/src/L2ProjectionPDESampler.hpp:48): enlarged NON-matching mesh; the field is
/// brought to the original mesh with s = W^-1 G^T s_bar (src/L2ProjectionPDESampler.cpp:595-611).  G itself (mortar
/// assembly, ParMoonolith) stays a host-once input.
class L2ProjectionPDESampler : public PDESampler {
public:
    using PDESampler::PDESampler;
    int SampleSize(int level) const override { return OutputSize(level); }
};

/// ComputeScalingCoefficientForSPDE (/root/reference/src/Utilities.hpp:188-200)
double ComputeScalingCoefficientForSPDE(double corlen, int myDim);
}  // namespace parelagmc
