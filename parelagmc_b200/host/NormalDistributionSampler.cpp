#include "NormalDistributionSampler.hpp"

#include <cmath>

namespace parelagmc {

NormalDistributionSampler::NormalDistributionSampler(double mu, double sigma2, std::shared_ptr<B200Device> dev)
    : dev_(std::move(dev)), mu_(mu), sigma_(std::sqrt(sigma2))  // d(mu, sqrt(sigma2)), reference .cpp:17-19
{
    Bind();
}

void NormalDistributionSampler::Bind()
{
    if (dev_->rng_owner == this) return;
    dev_->check(pmc_rng_init(dev_->handle(), mu_, sigma_, nparts_, mypart_), "pmc_rng_init");
    dev_->rng_owner = this;
}

void NormalDistributionSampler::Split(int nparts, int mypart)
{
    // rng.split(nparts, mypart) (reference .cpp:21-24); the sub-stream restarts at its own position 0
    nparts_ = nparts;
    mypart_ = mypart;
    dev_->rng_owner = nullptr;
    Bind();
    pos_ = 0;
}

double NormalDistributionSampler::operator()()
{
    double v = 0.0;
    Bind();
    dev_->check(pmc_rng_fill(dev_->handle(), pos_, 1, &v), "pmc_rng_fill");
    pos_ += 1;
    return v;
}

void NormalDistributionSampler::operator()(mfem::Vector &x)
{
    // for( ; it != end; ++it) *it = d(rng);   (reference .cpp:31-37): x.Size() consecutive draws, in order
    Bind();
    dev_->check(pmc_rng_fill(dev_->handle(), pos_, x.Size(), x.GetData()), "pmc_rng_fill");
    pos_ += (uint64_t)x.Size();
}
}  // namespace parelagmc
