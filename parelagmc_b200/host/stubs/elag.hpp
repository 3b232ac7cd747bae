// stubs/elag.hpp -- DECLARATIONS ONLY of the subset of ParELAG 2.0 the -DPARELAGMC_B200_WITH_PARELAG branch of the host
// layer uses (compile-only check, see stubs/mpi.h).  Names follow the reference's own call sites
// (/root/reference/src/PDESampler.cpp:60-334, src/DarcySolver.cpp:60-414); the two accessors of the un-assembled
// per-agglomerate mass blocks (DeRhamSequence::GetM, DofHandler::GetEntityRDofTable / GetrDofDofTable) are recalled from
// ParELAG's headers, which are not in the image [UPSTREAM-UNVERIFIED].
#pragma once
#include <memory>
#include <string>
#include <vector>
#include "mfem.hpp"

namespace parelag {
template <class T, class... A> std::unique_ptr<T> make_unique(A &&... a);
template <class T> std::unique_ptr<T> ToUnique(T *p);

class ParameterList {
public:
    explicit ParameterList(const std::string &name = "Default");
    ParameterList &Sublist(const std::string &name, bool must_exist = false);
    template <class T> T Get(const std::string &key, const T &def) const;
    std::string Get(const std::string &key, const char *def) const;
    template <class T> void Set(const std::string &key, const T &v);
};

class SharingMap {
public:
    int GetTrueLocalSize() const;
};

class AgglomeratedTopology {
public:
    enum Entity { ELEMENT = 0, FACET = 1, RIDGE = 2, PEAK = 3 };
    AgglomeratedTopology(const std::shared_ptr<mfem::ParMesh> &mesh, int codim);
    int GetNumberLocalEntities(Entity e) const;
    std::shared_ptr<AgglomeratedTopology> CoarsenLocalPartitioning(mfem::Array<int> &partitioning, bool check_topology,
                                                                   bool preserve_material_interfaces);
};

class DofHandler {
public:
    virtual ~DofHandler();
    int GetNDofs() const;
    SharingMap &GetDofTrueDof();
    int MarkDofsOnSelectedBndr(const mfem::Array<int> &bndr_attribute_selected, mfem::Array<int> &dof_marker) const;
    const mfem::SparseMatrix &GetEntityDofTable(AgglomeratedTopology::Entity e) const;
    const mfem::SparseMatrix &GetEntityRDofTable(AgglomeratedTopology::Entity e) const;   // [UPSTREAM-UNVERIFIED]
    const mfem::SparseMatrix &GetrDofDofTable(AgglomeratedTopology::Entity e) const;      // [UPSTREAM-UNVERIFIED]
};

class CochainProjector {
public:
    const mfem::SparseMatrix &GetProjectorMatrix();
};

class MultiVector { public: ~MultiVector(); };
class DeRhamSequenceFE;

class DeRhamSequence {
public:
    virtual ~DeRhamSequence();
    int GetNumberOfForms() const;
    int GetNumberOfDofs(int jform) const;
    int GetNumberOfTrueDofs(int jform) const;
    DofHandler *GetDofHandler(int jform);
    std::unique_ptr<mfem::SparseMatrix> ComputeMassOperator(int jform);
    std::unique_ptr<mfem::SparseMatrix> ComputeMassOperator(int jform, mfem::Vector &elemMatrixScaling);
    mfem::SparseMatrix *GetM(AgglomeratedTopology::Entity e, int jform);                  // [UPSTREAM-UNVERIFIED]
    mfem::SparseMatrix *GetDerivativeOperator(int jform);
    mfem::SparseMatrix *GetP(int jform);
    CochainProjector *GetPi(int jform);
    std::unique_ptr<mfem::HypreParMatrix> ComputeTrueP(int jform);
    std::shared_ptr<DeRhamSequence> Coarsen();
    DeRhamSequenceFE *FemSequence();
    void SetjformStart(int j);
    void SetTargets(const mfem::Array<MultiVector *> &targets);
};

class DeRhamSequenceFE : public DeRhamSequence {
public:
    mfem::FiniteElementSpace *GetFeSpace(int jform);
    void ReplaceMassIntegrator(AgglomeratedTopology::Entity e, int jform, std::unique_ptr<mfem::BilinearFormIntegrator> m,
                               bool recompute = true);
    std::unique_ptr<MultiVector> InterpolateScalarTargets(int jform, const mfem::Array<mfem::Coefficient *> &t);
    std::unique_ptr<MultiVector> InterpolateVectorTargets(int jform, const mfem::Array<mfem::VectorCoefficient *> &t);
    void ProjectVectorCoefficient(int jform, mfem::VectorCoefficient &c, mfem::Vector &v);
};
class DeRhamSequence3D_FE : public DeRhamSequenceFE {
public:
    DeRhamSequence3D_FE(const std::shared_ptr<AgglomeratedTopology> &topo, mfem::ParMesh *mesh, int order);
};
class DeRhamSequence2D_Hdiv_FE : public DeRhamSequenceFE {
public:
    DeRhamSequence2D_Hdiv_FE(const std::shared_ptr<AgglomeratedTopology> &topo, mfem::ParMesh *mesh, int order);
};
}  // namespace parelag
