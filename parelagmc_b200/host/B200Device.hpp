// B200Device.hpp -- one pmc_handle (include/pmc_b200.h) shared by the sampler and the solver of a run, plus the
// error convention of the host layer: C-ABI status codes become std::runtime_error, which the reference drivers
// catch in main (/root/reference/examples/MLMC.cpp:277-282).
#pragma once
#include <cstdlib>
#include <memory>
#include <stdexcept>
#include <string>
#include "../../include/pmc_b200.h"

namespace parelagmc {
class B200Device {
public:
    B200Device(int device, int nlevels) : nlevels_(nlevels)
    {
        if (pmc_create(device, nlevels, &h_) != PMC_OK)
            throw std::runtime_error(std::string("pmc_create: ") + pmc_last_error(nullptr));
    }
    ~B200Device() { pmc_destroy(h_); }
    B200Device(const B200Device &) = delete;
    B200Device &operator=(const B200Device &) = delete;
    /// The device context the reference-signature constructors (which have no device argument) share: created on first
    /// use on GPU PMC_DEVICE / LOCAL_RANK (default 0) with room for `nlevels` levels.
    static std::shared_ptr<B200Device> Default(int nlevels = 16)
    {
        static std::shared_ptr<B200Device> d;
        if (!d) {
            const char *e = std::getenv("PMC_DEVICE");
            if (!e) e = std::getenv("LOCAL_RANK");
            d = std::make_shared<B200Device>(e ? std::atoi(e) : 0, nlevels);
        }
        return d;
    }
    pmc_handle handle() const { return h_; }
    /// The NormalDistributionSampler whose (mu, sigma, split) the handle's generator currently carries: several samplers
    /// can share one device (the prior and the observation noise of BayesianInverseProblem), each re-binds when it is used.
    const void *rng_owner = nullptr;
    int nlevels() const { return nlevels_; }
    void check(int rc, const char *what) const
    {
        if (rc != PMC_OK) throw std::runtime_error(std::string(what) + ": " + pmc_last_error(h_));
    }
private:
    pmc_handle h_ = nullptr;
    int nlevels_;
};
}  // namespace parelagmc
