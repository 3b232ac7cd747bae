// MLSampler.hpp -- abstract sampler interface; same methods as /root/reference/src/MLSampler.hpp:22-91.
// (Methods whose signatures need ParELAG/MFEM types that have no shim -- BuildDeRhamSequence, SetDeRhamSequence,
// GetTrueP, GLVis output, L2 errors -- exist only when building against real ParELAG.)
#pragma once
#include <cstddef>
#include <memory>
#include <string>
#include <vector>
#include "shim.hpp"

namespace parelagmc {
class MLSampler {
public:
    MLSampler() {}
    virtual ~MLSampler() = default;
    /// Fill vector with random sample using dist_sampler
    virtual void Sample(const int level, mfem::Vector &xi) = 0;
    /// Evaluate random field at level with random sample xi
    virtual void Eval(const int level, const mfem::Vector &xi, mfem::Vector &s) = 0;
    virtual void Eval(const int level, const mfem::Vector &xi, mfem::Vector &s, mfem::Vector &u, bool use_init) = 0;
    virtual int SampleSize(int level) const = 0;
    virtual size_t GetNNZ(int level) const = 0;
    virtual void BuildHierarchy() = 0;
#ifdef PARELAGMC_B200_WITH_PARELAG
    // the rest of /root/reference/src/MLSampler.hpp:62-87, with the reference's signatures
    virtual void BuildDeRhamSequence(std::vector<std::shared_ptr<parelag::AgglomeratedTopology>> & /*topology*/) {}
    virtual void BuildDeRhamSequence(std::vector<std::shared_ptr<parelag::AgglomeratedTopology>> & /*topology*/,
                                     std::vector<std::shared_ptr<parelag::AgglomeratedTopology>> & /*embed_topology*/) {}
    virtual void SetDeRhamSequence(std::vector<std::shared_ptr<parelag::DeRhamSequence>> & /*sequence*/) {}
    virtual mfem::HypreParMatrix *GetTrueP(int level) = 0;
    virtual void SaveMeshGLVis(const std::string prefix) const = 0;
    virtual void SaveFieldGLVis(int level, const mfem::Vector &coeff, const std::string prefix) const = 0;
    virtual double ComputeL2Error(int level, const mfem::Vector &coeff, double exact) const = 0;
    virtual double ComputeMaxError(int level, const mfem::Vector &coeff, double exact) const = 0;
#endif
};
}  // namespace parelagmc
