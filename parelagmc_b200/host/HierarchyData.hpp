// HierarchyData.hpp -- the host-once data ParELAG leaves behind after DarcySolver::BuildHierachySpaces /
// PDESampler::BuildHierarchy (see include/pmc_b200.h for the meaning of every array), as plain arrays.  On a real
// build these are filled from the DeRhamSequence; here they are read from a dump written by
// parelagmc_b200/hierarchy.py (the Cartesian stand-in for ParELAG), so the drivers in examples/ run in this image.
#pragma once
#include <string>
#include <vector>

namespace parelagmc {

struct CsrData {
    int rows = 0, cols = 0;
    std::vector<int> rowptr, col;
    std::vector<double> val;
    bool empty() const { return rowptr.empty(); }
    size_t nnz() const { return col.size(); }
};

struct SamplerLevelData {
    int Ne = 0, Nf = 0;
    CsrData M, B, P;  // eliminated M, B (src/PDESampler.cpp:236-246); P = Ps[level] or empty
    std::vector<double> Wdiag;
    // enlarged-domain samplers only: transfer of the field to the forward problem's mesh, s = Tscale .* (T field)
    // (EmbeddedPDESampler: meshP, no scale; L2ProjectionPDESampler: G^T and 1/diag W_orig)
    CsrData T;
    std::vector<double> Tscale;
};

struct DarcyLevelData {
    int Ne = 0, Nf = 0;
    std::vector<int> elem_ptr, elem_dofs, ess_u;
    std::vector<double> elem_mat, ess_data, rhs, obs;
    CsrData B, Pp;
};

struct HierarchyData {
    int nlevels = 0, dim = 3;
    double corlen = 0.1;
    std::vector<SamplerLevelData> sampler;
    std::vector<DarcyLevelData> darcy;
    // BayesianInverseProblem set-up data (src/BayesianInverseProblem.cpp:46-104), optional: n_obs pressure functionals
    // g_obs_func[i][level], stored per level as [n_obs][Ne(level)] (un-normalised: element volumes of the marked elements
    // on level 0, P^T of the finer one below)
    int n_obs = 0;
    std::vector<std::vector<double>> gobs;
    // Reads the binary dump (magic "PMCH2", or "PMCH3" = PMCH2 + the observation functionals); throws
    // std::runtime_error on malformed input.
    static HierarchyData Load(const std::string &path);
};

}  // namespace parelagmc
