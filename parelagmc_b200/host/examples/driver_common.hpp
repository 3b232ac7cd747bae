// driver_common.hpp -- tiny command-line helper shared by the example drivers.
#pragma once
#include <cstdlib>
#include <cstring>
#include <sstream>
#include <string>
#include <vector>

struct DriverArgs {
    std::string hierarchy = "hierarchy.pmch";
    std::vector<int> samples;
    double mse = 0.001, rel_tol = 1e-6, abs_tol = 1e-12, variance = 1.0;
    int max_iter = 300, device = 0, nsamples = 10;
    bool wall_time = true;
    std::string log = "MLMC.dat";
    static DriverArgs Parse(int argc, char **argv)
    {
        DriverArgs a;
        // one rank per GPU: the launcher's LOCAL_RANK (torchrun) or PMC_RANK picks the device unless --device is given
        if (const char *lr = getenv("LOCAL_RANK")) a.device = atoi(lr);
        else if (const char *r = getenv("PMC_RANK")) a.device = atoi(r);
        for (int i = 1; i < argc; ++i) {
            auto next = [&]() -> const char * { return i + 1 < argc ? argv[++i] : ""; };
            if (!strcmp(argv[i], "--hierarchy")) a.hierarchy = next();
            else if (!strcmp(argv[i], "--samples")) {
                std::stringstream ss(next());
                std::string t;
                while (std::getline(ss, t, ',')) a.samples.push_back(atoi(t.c_str()));
            } else if (!strcmp(argv[i], "--nsamples")) a.nsamples = atoi(next());
            else if (!strcmp(argv[i], "--mse")) a.mse = atof(next());
            else if (!strcmp(argv[i], "--rel-tol")) a.rel_tol = atof(next());
            else if (!strcmp(argv[i], "--abs-tol")) a.abs_tol = atof(next());
            else if (!strcmp(argv[i], "--max-iter")) a.max_iter = atoi(next());
            else if (!strcmp(argv[i], "--variance")) a.variance = atof(next());
            else if (!strcmp(argv[i], "--device")) a.device = atoi(next());
            else if (!strcmp(argv[i], "--dof-cost")) a.wall_time = false;
            else if (!strcmp(argv[i], "--log")) a.log = next();
        }
        return a;
    }
};
