// TEST INFRASTRUCTURE: C entry point around genomic_b200/csrc/prune.h (a host-only header of the product) so that the CPU suite can
// compare it with the compiled reference's undo_prune path (tests/test_cpu_host.py).
#include <cstdint>
#include <vector>

#include "../../genomic_b200/csrc/prune.h"

extern "C" int prune_host(const double* x, int n, const int* lseg, int nseg, double cutoff, int* out) {
    const std::vector<int> in(lseg, lseg + nseg);
    const std::vector<int> r = cbsg::prune_lengths(x, n, in, cutoff);
    for (size_t i = 0; i < r.size(); ++i) out[i] = r[i];
    return (int)r.size();
}
