// ORACLE -- TEST INFRASTRUCTURE ONLY.  oracle/_ref/libref_sobol.so: the REFERENCE'S OWN sobolld code, compiled unmodified from
// /root/reference/include/wt/sampler/sobolld/{integer3,irreducible_gf3,sobolld_sampler}.hpp (two shim headers under oracle/ref_shims/ replace the
// glm / mp-units umbrella headers those files include but do not need).  It pins the restatement in ot_sobol.h -- and, through it, the device
// generator -- bit for bit: tests/test_sobol.py::test_oracle_sobol_equals_the_reference_code.  Built only where /root/reference is mounted.
#include <cstdint>
#include <cstring>
#include <wt/sampler/sobolld/sobolld_sampler.hpp>

using namespace wt::sampler::sobolld;
using sampler_t = sobolls_sampler<47>;      // src/sampler/sobolld.cpp:31

extern "C" {
// table file in the format of data/sobolld/initIrreducibleGF3.dat (parsed by the reference's load_mk); seeds: the 47 values the reference
// draws from its RNG, one per dimension (sobolld_sampler.hpp:69-71); out: n_points x 47 f32, point-major (generate_points' layout)
int ref_sobol_points(const char* dat_path, const uint64_t* seeds, uint32_t n_points, float* out) {
    try {
        const irreducible_gf3_t gf3{ std::filesystem::path(dat_path) };
        const sampler_t s(sampler_t::max_mat_size(), gf3);     // src/sampler/sobolld.cpp:36-38
        std::size_t i = 0;
        const auto pts = s.generate_points<float>([&]() { return seeds[i++]; }, (std::size_t)n_points);
        std::memcpy(out, pts.data(), pts.size() * sizeof(float));
        return (int)(pts.size() / 47);
    } catch (...) { return -1; }
}
// generator matrices as gen_mat builds them: out[dim][row][col], 47 x 11 x 11
int ref_sobol_matrices(const char* dat_path, int32_t* out) {
    try {
        const irreducible_gf3_t gf3{ std::filesystem::path(dat_path) };
        const sampler_t s(sampler_t::max_mat_size(), gf3);
        for (std::size_t d = 0; d < 47; ++d) for (std::size_t r = 0; r < 11; ++r) for (std::size_t c = 0; c < 11; ++c) out[(d * 11 + r) * 11 + c] = (int32_t)s.matrix[d][r][c];
        return 0;
    } catch (...) { return -1; }
}
}
