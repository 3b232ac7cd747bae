// MLSampler.hpp -- abstract sampler interface; same methods as /root/reference/src/MLSampler.hpp:22-91.
// (Methods whose signatures need ParELAG/MFEM types that have no shim -- BuildDeRhamSequence, SetDeRhamSequence,
// GetTrueP, GLVis output, L2 errors -- exist only when building against real ParELAG.)
#pragma once
#include <cstddef>
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
};
}  // namespace parelagmc
