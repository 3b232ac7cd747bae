// NormalDistributionSampler.hpp -- same interface as /root/reference/src/NormalDistributionSampler.hpp:27-66; the
// trng::yarn5 engine + trng::normal_dist pair lives on the GPU (csrc/rng.cuh).  The object keeps the stream
// position the reference's engine would be at, so every call returns exactly the values the sequential engine
// would (draws are addressed by absolute position on the device).
#pragma once
#include <cstdint>
#include <memory>
#include "B200Device.hpp"
#include "shim.hpp"

namespace parelagmc {
class NormalDistributionSampler {
public:
    /// Constructor (mean mu, variance sigma2); `dev` is the device context shared with the sampler/solver.
    NormalDistributionSampler(double mu, double sigma2, std::shared_ptr<B200Device> dev);
    /// The reference's constructor, verbatim (/root/reference/src/NormalDistributionSampler.hpp:31): the engine lives on
    /// the process-wide default device (B200Device::Default).
    NormalDistributionSampler(double mu, double sigma2) : NormalDistributionSampler(mu, sigma2, B200Device::Default()) {}
    ~NormalDistributionSampler() = default;
    NormalDistributionSampler(NormalDistributionSampler const &) = delete;
    NormalDistributionSampler(NormalDistributionSampler &&) = delete;
    NormalDistributionSampler &operator=(NormalDistributionSampler const &) = delete;
    NormalDistributionSampler &operator=(NormalDistributionSampler &&) = delete;

    /// Provides statistically independent of random numbers to each process (trng leapfrog split)
    void Split(int nparts, int mypart);
    /// Get a random number from normal distribution.
    double operator()();
    /// Fill uncorrelated random numbers from normal distribution.
    void operator()(mfem::Vector &v);

    /// Stream position of the next draw / reserve `n` draws for a batched device-side generation.
    /// Make the device generator carry this sampler's distribution (no-op while it already does).
    void Bind();
    uint64_t Position() const { return pos_; }
    uint64_t Advance(uint64_t n) { uint64_t p = pos_; pos_ += n; return p; }
    const std::shared_ptr<B200Device> &Device() const { return dev_; }

private:
    std::shared_ptr<B200Device> dev_;
    double mu_, sigma_;
    int nparts_ = 1, mypart_ = 0;
    uint64_t pos_ = 0;
};
}  // namespace parelagmc
