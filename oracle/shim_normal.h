// oracle/shim_normal.h -- TEST INFRASTRUCTURE (CPU oracle build only).
//
// Force-included (`g++ -include oracle/shim_normal.h`) in front of the UNMODIFIED reference translation
// unit /root/reference/src/models/RoughVolatility.cpp so that its two uses of
// `std::normal_distribution<double>` (RoughVolatility.cpp:241 genComplexGaussians, :255 gaussians) draw
// from a caller-supplied sequence instead of a std::random_device-seeded mt19937 (which makes the
// reference non-reproducible by construction, RoughVolatility.cpp:239-240,253-254).
//
// Draw order consumed per path (RoughVolatility.cpp:346-352):
//   Zre_0, Zim_0, ..., Zre_{n-1}, Zim_{n-1}, W1_0..W1_{n-1}, W2_0..W2_{n-1}      (4n normals)
//
// When no sequence is installed (orc_inject_begin not called on this thread) the wrapper forwards to the
// real std::normal_distribution, i.e. the reference runs exactly as shipped (used for CPU timing).
#pragma once
#include <cstddef>
#include <random>
#include <vector>
#include <complex>
#include <cmath>
#include <algorithm>
#include <numeric>
#include <iostream>
#include <stdexcept>

extern "C" {
// Installed per thread by oracle/ref_api.cpp.
extern thread_local const double* orc_inject_ptr;
extern thread_local std::size_t orc_inject_left;
extern thread_local std::size_t orc_inject_used;
}

namespace std {
template <class T>
class orc_injected_normal {
public:
    orc_injected_normal(T mean, T sd) : real_(mean, sd) {}
    template <class G>
    T operator()(G& g) {
        if (orc_inject_ptr) {
            if (orc_inject_left == 0) throw std::runtime_error("oracle: injected draw sequence exhausted");
            --orc_inject_left;
            ++orc_inject_used;
            return static_cast<T>(*orc_inject_ptr++);
        }
        return real_(g);
    }
private:
    std::normal_distribution<T> real_;  // declared before the macro below: the genuine libstdc++ type
};
}  // namespace std

#define normal_distribution orc_injected_normal
