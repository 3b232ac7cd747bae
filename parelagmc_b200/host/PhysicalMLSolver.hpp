// PhysicalMLSolver.hpp -- abstract forward-solver interface; same methods as
// /root/reference/src/PhysicalMLSolver.hpp:22-63 (GetSequence / GetPressureSpace need real ParELAG/MFEM types).
#pragma once
#include <memory>
#include <vector>
#include "shim.hpp"

namespace parelagmc {
class PhysicalMLSolver {
public:
    PhysicalMLSolver() {}
    virtual ~PhysicalMLSolver() = default;
    /// Solve and update quantity of interest Q, cost C
    virtual void SolveFwd(int ilevel, mfem::Vector &k_over_k_ref, double &Q, double &C) = 0;
    virtual void SolveFwd_RtnPressure(int ilevel, mfem::Vector &k_over_k_ref, mfem::Vector &P, double &C, double &Q,
                                      bool compute_Q) = 0;
    virtual int GetNumberOfDofs(int ilevel) const = 0;
    virtual int GetGlobalNumberOfDofs(int ilevel) const = 0;
    virtual int GetNNZ(int ilevel) const = 0;
#ifdef PARELAGMC_B200_WITH_PARELAG
    // the rest of /root/reference/src/PhysicalMLSolver.hpp:49-52
    virtual std::vector<std::shared_ptr<parelag::DeRhamSequence>> &GetSequence() = 0;
    virtual mfem::FiniteElementSpace *GetPressureSpace() const = 0;
#endif
};
}  // namespace parelagmc
