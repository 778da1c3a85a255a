// oracle/shim_uniform.h -- TEST INFRASTRUCTURE (CPU oracle build only).
//
// Force-included in front of the UNMODIFIED reference translation unit
// /root/reference/src/models/BranchingProcessPricer.cpp so that its `std::uniform_int_distribution<> pathDist(0, N-1)`
// (BranchingProcessPricer.cpp:86, drawn at :108) can take its path indices from a caller-supplied sequence.  That TU
// is compiled WITHOUT -fopenmp for the oracle, so the `#pragma omp parallel for` at :90-92 is ignored and the
// consumption order is deterministic: for path i, for exercise date e (while e < exerciseTimes.back()), for branch b.
// With no sequence installed the wrapper forwards to the genuine distribution (reference as shipped).
#pragma once
#include <cstddef>
#include <random>
#include <stdexcept>

extern "C" {
extern thread_local const int* orc_index_ptr;
extern thread_local std::size_t orc_index_left;
extern thread_local std::size_t orc_index_used;
}

namespace std {
template <class T = int>
class orc_injected_uniform_int {
public:
    orc_injected_uniform_int(T a, T b) : real_(a, b) {}
    template <class G>
    T operator()(G& g) {
        if (orc_index_ptr) {
            if (orc_index_left == 0) throw std::runtime_error("oracle: injected index sequence exhausted");
            --orc_index_left;
            ++orc_index_used;
            return static_cast<T>(*orc_index_ptr++);
        }
        return real_(g);
    }
private:
    std::uniform_int_distribution<T> real_;
};
}  // namespace std

#define uniform_int_distribution orc_injected_uniform_int
