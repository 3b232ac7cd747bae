#include "HierarchyData.hpp"

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <stdexcept>

namespace parelagmc {
namespace {
struct Reader {
    FILE *f;
    explicit Reader(const std::string &p) : f(fopen(p.c_str(), "rb"))
    {
        if (!f) throw std::runtime_error("HierarchyData::Load: cannot open " + p);
    }
    ~Reader() { if (f) fclose(f); }
    void raw(void *dst, size_t bytes)
    {
        if (bytes && fread(dst, 1, bytes, f) != bytes) throw std::runtime_error("HierarchyData::Load: truncated file");
    }
    int32_t i32() { int32_t v; raw(&v, 4); return v; }
    double f64() { double v; raw(&v, 8); return v; }
    std::vector<int> ivec()
    {
        int64_t n; raw(&n, 8);
        if (n < 0 || n > (int64_t)1 << 33) throw std::runtime_error("HierarchyData::Load: bad array length");
        std::vector<int> v((size_t)n); raw(v.data(), (size_t)n * 4); return v;
    }
    std::vector<double> dvec()
    {
        int64_t n; raw(&n, 8);
        if (n < 0 || n > (int64_t)1 << 33) throw std::runtime_error("HierarchyData::Load: bad array length");
        std::vector<double> v((size_t)n); raw(v.data(), (size_t)n * 8); return v;
    }
    CsrData csr()
    {
        CsrData m;
        int present = i32();
        if (!present) return m;
        m.rows = i32(); m.cols = i32();
        m.rowptr = ivec(); m.col = ivec(); m.val = dvec();
        if ((int)m.rowptr.size() != m.rows + 1 || m.col.size() != m.val.size())
            throw std::runtime_error("HierarchyData::Load: inconsistent CSR block");
        return m;
    }
};
}  // namespace

HierarchyData HierarchyData::Load(const std::string &path)
{
    Reader r(path);
    char magic[8];
    r.raw(magic, 8);
    const bool v3 = memcmp(magic, "PMCH3\0\0\0", 8) == 0;
    if (!v3 && memcmp(magic, "PMCH2\0\0\0", 8) != 0) throw std::runtime_error("HierarchyData::Load: bad magic in " + path);
    HierarchyData h;
    h.nlevels = r.i32();
    h.dim = r.i32();
    h.corlen = r.f64();
    if (h.nlevels < 1 || h.nlevels > 64) throw std::runtime_error("HierarchyData::Load: bad level count");
    h.sampler.resize(h.nlevels);
    h.darcy.resize(h.nlevels);
    for (int l = 0; l < h.nlevels; ++l) {
        SamplerLevelData &s = h.sampler[l];
        s.Ne = r.i32(); s.Nf = r.i32();
        s.M = r.csr(); s.B = r.csr(); s.P = r.csr();
        s.Wdiag = r.dvec();
        s.T = r.csr();
        s.Tscale = r.dvec();
        DarcyLevelData &d = h.darcy[l];
        d.Ne = r.i32(); d.Nf = r.i32();
        d.elem_ptr = r.ivec(); d.elem_dofs = r.ivec(); d.elem_mat = r.dvec();
        d.B = r.csr(); d.Pp = r.csr();
        d.ess_u = r.ivec(); d.ess_data = r.dvec(); d.rhs = r.dvec(); d.obs = r.dvec();
    }
    if (v3) {
        h.n_obs = r.i32();
        if (h.n_obs < 0 || h.n_obs > 4096) throw std::runtime_error("HierarchyData::Load: bad number of observations");
        h.gobs.resize(h.nlevels);
        for (int l = 0; l < h.nlevels; ++l) {
            h.gobs[l] = r.dvec();
            if (h.gobs[l].size() != (size_t)h.n_obs * (size_t)h.darcy[l].Ne)
                throw std::runtime_error("HierarchyData::Load: observation functionals of the wrong size");
        }
    }
    return h;
}
}  // namespace parelagmc
