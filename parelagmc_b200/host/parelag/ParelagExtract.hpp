// ParelagExtract.hpp -- (only with -DPARELAGMC_B200_WITH_PARELAG) from ParELAG's DeRhamSequence hierarchy to the plain
// arrays the C ABI uploads (HierarchyData.hpp).  This is the one place that touches ParELAG/MFEM containers; the calls are
// the ones the reference makes at /root/reference/src/PDESampler.cpp:218-258 and src/DarcySolver.cpp:160-244.
// One rank holds the whole hierarchy (rank = GPU): dofs and true dofs coincide.
#pragma once
#ifdef PARELAGMC_B200_WITH_PARELAG
#include <memory>
#include <vector>
#include "../HierarchyData.hpp"
#include "../shim.hpp"

namespace parelagmc {
typedef std::vector<std::shared_ptr<parelag::DeRhamSequence>> SequenceVector;

/// mfem::SparseMatrix -> CSR arrays exactly as stored (bit-exact index structures are part of the contract)
CsrData ToCsr(const mfem::SparseMatrix &A);
/// The DeRham hierarchy of one mesh (what BuildDeRhamSequence / BuildHierachySpaces build): finest-level FE sequence of
/// the given order with constant targets, coarsened over `topology`.  massIntegrator (may be null) replaces the mass
/// integrator of the uform on the elements (src/DarcySolver.cpp:107-109).
SequenceVector BuildSequences(const std::shared_ptr<mfem::ParMesh> &mesh,
                              std::vector<std::shared_ptr<parelag::AgglomeratedTopology>> &topology, int feorder,
                              int upscalingOrder, std::unique_ptr<mfem::BilinearFormIntegrator> massIntegrator);
/// Eliminated M and B = W D, diag(W), P_s of one sampler level (src/PDESampler.cpp:230-258); essential = whole boundary
void ExtractSamplerLevel(parelag::DeRhamSequence &seq, int uform, int sform, bool has_coarser, int bdr_size,
                         SamplerLevelData &out);
/// Un-assembled per-agglomerate mass blocks + their dof lists, B = W D, P_p of one Darcy level
void ExtractDarcyLevel(parelag::DeRhamSequence &seq, int uform, int pform, bool has_coarser, DarcyLevelData &out);
}  // namespace parelagmc
#endif
