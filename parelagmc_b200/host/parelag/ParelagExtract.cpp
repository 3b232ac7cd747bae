#ifdef PARELAGMC_B200_WITH_PARELAG
#include "ParelagExtract.hpp"

#include <stdexcept>

namespace parelagmc {
using parelag::AgglomeratedTopology;

CsrData ToCsr(const mfem::SparseMatrix &A)
{
    CsrData c;
    c.rows = A.Height();
    c.cols = A.Width();
    const int *I = A.GetI(), *J = A.GetJ();
    const double *V = A.GetData();
    c.rowptr.assign(I, I + c.rows + 1);
    c.col.assign(J, J + I[c.rows]);
    c.val.assign(V, V + I[c.rows]);
    return c;
}

SequenceVector BuildSequences(const std::shared_ptr<mfem::ParMesh> &mesh,
                              std::vector<std::shared_ptr<AgglomeratedTopology>> &topology, int feorder, int upscalingOrder,
                              std::unique_ptr<mfem::BilinearFormIntegrator> massIntegrator)
{
    const int nDim = mesh->Dimension(), nLevels = (int)topology.size();
    const int uform = nDim - 1;
    SequenceVector seq(nLevels);
    if (nDim == 3) seq[0] = std::make_shared<parelag::DeRhamSequence3D_FE>(topology[0], mesh.get(), feorder);
    else seq[0] = std::make_shared<parelag::DeRhamSequence2D_Hdiv_FE>(topology[0], mesh.get(), feorder);
    parelag::DeRhamSequenceFE *fe = seq[0]->FemSequence();
    if (!fe) throw std::runtime_error("BuildSequences: no finite element sequence on the finest level");
    if (massIntegrator) fe->ReplaceMassIntegrator(AgglomeratedTopology::ELEMENT, uform, std::move(massIntegrator), true);
    // lowest-order targets (upscaling order 0: constants); only the forms uform, uform + 1 are built, as the drivers'
    // default "Build entire sequence" = false does (src/PDESampler.cpp:113-117)
    if (upscalingOrder != 0) throw std::runtime_error("BuildSequences: upscaling order 0 only");
    seq[0]->SetjformStart(uform);
    mfem::ConstantCoefficient one(1.0);
    mfem::Vector ex(nDim), ey(nDim), ez(nDim);
    ex = 0.0; ey = 0.0; ez = 0.0;
    ex(0) = 1.0; ey(1) = 1.0;
    if (nDim == 3) ez(2) = 1.0;
    mfem::VectorConstantCoefficient cx(ex), cy(ey), cz(ez);
    mfem::Array<mfem::Coefficient *> l2(1);
    l2[0] = &one;
    mfem::Array<mfem::VectorCoefficient *> hdiv(nDim);
    hdiv[0] = &cx; hdiv[1] = &cy;
    if (nDim == 3) hdiv[2] = &cz;
    std::vector<std::unique_ptr<parelag::MultiVector>> targets(seq[0]->GetNumberOfForms());
    targets[uform] = fe->InterpolateVectorTargets(uform, hdiv);
    targets[uform + 1] = fe->InterpolateScalarTargets(uform + 1, l2);
    mfem::Array<parelag::MultiVector *> tin((int)targets.size());
    for (int i = 0; i < tin.Size(); ++i) tin[i] = targets[i].get();
    seq[0]->SetTargets(tin);
    for (int i = 0; i + 1 < nLevels; ++i) seq[i + 1] = seq[i]->Coarsen();
    return seq;
}

void ExtractSamplerLevel(parelag::DeRhamSequence &seq, int uform, int sform, bool has_coarser, int bdr_size,
                         SamplerLevelData &out)
{
    out.Nf = seq.GetNumberOfDofs(uform);
    out.Ne = seq.GetNumberOfDofs(sform);
    auto M = seq.ComputeMassOperator(uform);
    auto W = seq.ComputeMassOperator(sform);
    mfem::SparseMatrix *D = seq.GetDerivativeOperator(uform);
    mfem::Array<int> ess_bdr(bdr_size), ess_dof(out.Nf);
    ess_bdr = 1;                                                     // u.n = 0 on the whole boundary (:210-214)
    seq.GetDofHandler(uform)->MarkDofsOnSelectedBndr(ess_bdr, ess_dof);
    for (int i = 0; i < ess_dof.Size(); ++i)
        if (ess_dof[i]) M->EliminateRowCol(i);                       // :236-241
    D->EliminateCols(ess_dof);                                       // :243
    std::unique_ptr<mfem::SparseMatrix> B(mfem::Mult(*W, *D));       // :245
    mfem::Vector wd(out.Ne);
    W->GetDiag(wd);
    out.Wdiag.assign(wd.GetData(), wd.GetData() + out.Ne);
    out.M = ToCsr(*M);
    out.B = ToCsr(*B);
    if (has_coarser) out.P = ToCsr(*seq.GetP(sform));                // Ps[i] (:189-193; serial: P = true P)
}

void ExtractDarcyLevel(parelag::DeRhamSequence &seq, int uform, int pform, bool has_coarser, DarcyLevelData &out)
{
    out.Nf = seq.GetNumberOfDofs(uform);
    out.Ne = seq.GetNumberOfDofs(pform);
    // un-assembled mass: block diagonal over the agglomerates in repeated-dof numbering, rdof -> dof map, agglomerate ->
    // rdof ranges.  ComputeMassOperator(uform, k) is Assemble(ELEMENT, diag-scale(M_el, k), dof, dof)
    // (call site /root/reference/src/DarcySolver.cpp:479): exactly sum_e k_e R_e^T M_e R_e of these blocks.
    parelag::DofHandler *dh = seq.GetDofHandler(uform);
    const mfem::SparseMatrix &Mel = *seq.GetM(AgglomeratedTopology::ELEMENT, uform);
    const mfem::SparseMatrix &el_rdof = dh->GetEntityRDofTable(AgglomeratedTopology::ELEMENT);
    const mfem::SparseMatrix &rdof_dof = dh->GetrDofDofTable(AgglomeratedTopology::ELEMENT);
    if (el_rdof.Height() != out.Ne) throw std::runtime_error("ExtractDarcyLevel: one L2 dof per agglomerate expected");
    const int *eI = el_rdof.GetI(), *eJ = el_rdof.GetJ(), *rJ = rdof_dof.GetJ();
    const int *mI = Mel.GetI(), *mJ = Mel.GetJ();
    const double *mV = Mel.GetData();
    out.elem_ptr.assign(1, 0);
    for (int e = 0; e < out.Ne; ++e) {
        const int r0 = eI[e], r1 = eI[e + 1], n = r1 - r0;
        for (int a = r0; a < r1; ++a) out.elem_dofs.push_back(rJ[eJ[a]]);
        const size_t base = out.elem_mat.size();
        out.elem_mat.resize(base + (size_t)n * n, 0.0);
        for (int a = r0; a < r1; ++a) {
            const int row = eJ[a];
            for (int p = mI[row]; p < mI[row + 1]; ++p) {
                int b = -1;                                          // position of column rdof inside the agglomerate
                for (int q = r0; q < r1; ++q)
                    if (eJ[q] == mJ[p]) { b = q - r0; break; }
                if (b < 0) throw std::runtime_error("ExtractDarcyLevel: mass block couples two agglomerates");
                out.elem_mat[base + (size_t)(a - r0) * n + b] = mV[p];
            }
        }
        out.elem_ptr.push_back((int)out.elem_dofs.size());
    }
    auto W = seq.ComputeMassOperator(pform);
    std::unique_ptr<mfem::SparseMatrix> B(mfem::Mult(*W, *seq.GetDerivativeOperator(uform)));   // src/DarcySolver.cpp:203-207
    out.B = ToCsr(*B);
    if (has_coarser) out.Pp = ToCsr(*seq.GetP(pform));
}
}  // namespace parelagmc
#endif
