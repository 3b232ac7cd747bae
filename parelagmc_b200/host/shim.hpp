// shim.hpp -- the few third-party types the reference's L3/L4 interfaces mention, so that the host layer builds
// without MFEM / ParELAG / MPI (none is in this image).  With -DPARELAGMC_B200_WITH_PARELAG the real headers are
// used instead and these definitions vanish; the class interfaces in this directory are written against the subset
// both provide (mfem::Vector::{SetSize,Size,GetData,operator()}, parelag::ParameterList::{Get,Sublist}).
#pragma once
#ifdef PARELAGMC_B200_WITH_PARELAG
#include <elag.hpp>
#include <mpi.h>
#else
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

#include <cstdlib>

typedef int MPI_Comm;
#ifndef MPI_COMM_WORLD
#define MPI_COMM_WORLD 0
#endif
// Without MPI the ranks of a multi-GPU run are plain processes started by a launcher that sets PMC_WORLD_SIZE / PMC_RANK
// (or torchrun's WORLD_SIZE / RANK): the two calls of MPI the managers make read those (RankComm.hpp).
inline int pmc_shim_env_int(const char *a, const char *b, int def)
{
    const char *v = std::getenv(a);
    if (!v || !*v) v = std::getenv(b);
    return v && *v ? std::atoi(v) : def;
}
inline int MPI_Comm_size(MPI_Comm, int *n) { *n = pmc_shim_env_int("PMC_WORLD_SIZE", "WORLD_SIZE", 1); return 0; }
inline int MPI_Comm_rank(MPI_Comm, int *r) { *r = pmc_shim_env_int("PMC_RANK", "RANK", 0); return 0; }

namespace mfem {
class Vector {
public:
    Vector() {}
    explicit Vector(int n) : d_(n, 0.0) {}
    void SetSize(int n) { d_.resize(n); }
    int Size() const { return (int)d_.size(); }
    double *GetData() { return d_.data(); }
    const double *GetData() const { return d_.data(); }
    double &operator()(int i) { return d_[i]; }
    const double &operator()(int i) const { return d_[i]; }
    double &operator[](int i) { return d_[i]; }
    const double &operator[](int i) const { return d_[i]; }
    Vector &operator=(double v) { for (auto &x : d_) x = v; return *this; }
    double Sum() const { double s = 0; for (double x : d_) s += x; return s; }
private:
    std::vector<double> d_;
};
}  // namespace mfem

namespace parelag {
// Typed key/value tree with defaults, the subset of parelag::ParameterList the hot path reads.
class ParameterList {
public:
    explicit ParameterList(const std::string &name = "Default") : name_(name) {}
    ParameterList &Sublist(const std::string &n, bool /*must_exist*/ = false)
    {
        auto &p = subs_[n];
        if (!p) p.reset(new ParameterList(n));
        return *p;
    }
    void Set(const std::string &k, double v) { num_[k] = v; }
    void Set(const std::string &k, int v) { num_[k] = v; }
    void Set(const std::string &k, bool v) { num_[k] = v ? 1.0 : 0.0; }
    void Set(const std::string &k, const char *v) { str_[k] = v; }
    void Set(const std::string &k, const std::string &v) { str_[k] = v; }
    void Set(const std::string &k, const std::vector<int> &v) { ivec_[k] = v; }
    double Get(const std::string &k, double def) const { auto it = num_.find(k); return it == num_.end() ? def : it->second; }
    int Get(const std::string &k, int def) const { auto it = num_.find(k); return it == num_.end() ? def : (int)it->second; }
    bool Get(const std::string &k, bool def) const { auto it = num_.find(k); return it == num_.end() ? def : it->second != 0.0; }
    std::string Get(const std::string &k, const char *def) const { auto it = str_.find(k); return it == str_.end() ? std::string(def) : it->second; }
    std::vector<int> Get(const std::string &k, const std::vector<int> &def) const { auto it = ivec_.find(k); return it == ivec_.end() ? def : it->second; }
private:
    std::string name_;
    std::map<std::string, double> num_;
    std::map<std::string, std::string> str_;
    std::map<std::string, std::vector<int>> ivec_;
    std::map<std::string, std::unique_ptr<ParameterList>> subs_;
};
}  // namespace parelag
#endif
