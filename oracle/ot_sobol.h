// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_sobol.h: literal CPU restatement of the reference's base-3 Owen-scrambled Sobol' sampler ("sobolld"):
//   include/wt/sampler/sobolld/integer3.hpp:20-76          (base-3 integers, mod / fma tables)
//   include/wt/sampler/sobolld/irreducible_gf3.hpp:32-118  (direction numbers m_k over GF(3) from a primitive polynomial)
//   include/wt/sampler/sobolld/sobolld_sampler.hpp:25-208  (generator matrices, incremental point update, nested scrambling)
//   src/sampler/sobolld.cpp:29-60                          (47 dimensions, matrix size 11 -> batches of 3^11 points)
// All of it is integer arithmetic: the device generator (wave_tracer_b200/csrc/dsobol.cuh, which computes a point directly
// from its index instead of incrementally) must reproduce the digit values BIT-EXACTLY (tests/test_sobol.py).
//
// PINNED against the reference's own code: those three headers compile unmodified into oracle/_ref/libref_sobol.so (oracle/ref_sobol.cpp, `make ref`)
// and tests/test_sobol.py::test_oracle_sobol_equals_the_reference_code compares matrices and points bit for bit.
//
// The table data/sobolld/initIrreducibleGF3.dat is a Git-LFS pointer stub in the reference tree: the table is
// an input here (wtgpu_sobol_entry[48]), so the index / digit / scramble math is pinned for any table.
#pragma once
#include <cstdint>
#include <cstddef>
#include <array>
#include <vector>
#include <cassert>
#include "../include/wtgpu.h"

namespace ot { namespace sobol {

static constexpr std::array<std::uint64_t, 21> pow3tab = {
    1,3,9,27,81,243,729,2187,6561,19683,59049,
    177147,531441,1594323,4782969,14348907,43046721,129140163,387420489,1162261467,3486784401
};

// integer3.hpp:29-76 (digit_t = int64, N = 11)
static constexpr std::size_t N = 11;        // irreducible_gf3.hpp:33 sobolld_gfn_seq_length
static constexpr std::size_t ENTRIES = 48;  // irreducible_gf3.hpp:34
static constexpr std::size_t D = 47;        // src/sampler/sobolld.cpp:31
using digit_t = std::int64_t;
using uint_t = std::uint64_t;

struct int3_t {
    std::array<digit_t, N> digits;
    int3_t(std::uint64_t x) { for (std::size_t i = 0; i < N; ++i) digits[i] = (digit_t)((x / pow3tab[i]) % 3); }
    int3_t() : int3_t(0) {}
    std::uint64_t value(std::size_t m = N) const {
        std::uint64_t x = 0;
        for (std::size_t i = 0; i < m; ++i) x += pow3tab[i] * (std::uint64_t)digits[i];
        return x;
    }
    float value_fp(std::size_t m = N) const { return float(value(m)) / float(pow3tab[m]); }
    // integer3.hpp:52-73: x % 3 for x in [0,6) and (a + (b*c)%3)%3 for digits a,b,c (the reference tabulates both)
    static digit_t mod(int x) { assert(x >= 0 && x < 6); return x % 3; }
    static digit_t fma(int a, int b, int c) { assert(a >= 0 && a < 3 && b >= 0 && b < 3 && c >= 0 && c < 3); return (a + (b * c) % 3) % 3; }
};

// irreducible_gf3.hpp:50-118
struct gf3_t {
    digit_t sobol_dj[ENTRIES], sobol_sj[ENTRIES], sobol_aj[ENTRIES], sobol_mk[ENTRIES][32];
    explicit gf3_t(const wtgpu_sobol_entry* e) {           // replaces load_mk (irreducible_gf3.hpp:124-158): entries arrive parsed
        for (std::size_t i = 0; i < ENTRIES; ++i) {
            sobol_dj[i] = e[i].d; sobol_sj[i] = e[i].sj; sobol_aj[i] = e[i].aj;
            for (int k = 0; k < 32; ++k) sobol_mk[i][k] = e[i].mk[k];
        }
    }
    static void to_digit_array(digit_t* digits, digit_t val, int base, int len) { for (int i = 0; i < len; ++i) { digits[i] = val % base; val = val / base; } }
    static digit_t from_digit_array(const digit_t* digits, int base, int len) {
        digit_t pow = 1, res = 0;
        for (int i = 0; i < len; ++i) { res += pow * digits[i]; pow *= base; }
        return res;
    }
    static digit_t multiply_by_factor_in_gfn(digit_t x, digit_t factor, int base) {
        digit_t digits[N];
        to_digit_array(digits, x, base, (int)N);
        for (std::size_t i = 0; i < N; ++i) digits[i] = (digits[i] * factor) % base;
        return from_digit_array(digits, base, (int)N);
    }
    static digit_t bit_xor_gfn(digit_t* data, int base, const digit_t* lst, int len, int polynomial_degree) {
        digit_t digits[N][N];
        for (int i = 0; i <= polynomial_degree; ++i) {
            to_digit_array(data, lst[i], base, len);
            for (int j = 0; j < len; ++j) digits[i][j] = data[j];
        }
        digit_t final_digits[N + 1];
        for (int i = 0; i < len; ++i) {
            final_digits[i] = 0;
            for (int j = 0; j <= polynomial_degree; ++j) final_digits[i] += digits[j][i];
            final_digits[i] %= base;
        }
        return from_digit_array(final_digits, base, len);
    }
    static void generate_mkgf3(digit_t ipolynomial, digit_t polynomial_degree, digit_t* msobol, int base) {
        static constexpr digit_t convert_to_gf3[3] = { 0, 2, 1 };
        digit_t polynomial[N], d[N], lst[N];
        to_digit_array(polynomial, ipolynomial, base, (int)polynomial_degree + 1);
        for (int i = (int)polynomial_degree + 1; i <= (int)N; ++i) {
            lst[0] = msobol[i - polynomial_degree - 1];
            for (int j = 1; j < polynomial_degree + 1; ++j)
                lst[j] = (digit_t)pow3tab[j] * multiply_by_factor_in_gfn(msobol[i - j - 1], convert_to_gf3[polynomial[polynomial_degree - j]], base);
            msobol[i - 1] = bit_xor_gfn(d, base, lst, i, (int)polynomial_degree);
        }
    }
};

// sobolld_sampler.hpp:25-208
struct sobolls_sampler {
    using row_t = std::array<digit_t, N>;
    using matrix_t = std::vector<row_t>;
    std::array<matrix_t, D> matrix;

    sobolls_sampler(std::size_t mat_size, const gf3_t& gf3) {
        for (std::size_t d = 0; d < D; ++d) {
            std::array<digit_t, 32> mk;
            for (int k = 0; k < 32; ++k) mk[k] = gf3.sobol_mk[d + 1][k];
            gf3_t::generate_mkgf3(gf3.sobol_aj[d + 1], gf3.sobol_sj[d + 1], mk.data(), 3);
            matrix[d] = gen_mat(mk, mat_size);
        }
    }
    static matrix_t gen_mat(const std::array<digit_t, 32>& sobol_mk, std::size_t mat_size) {
        matrix_t m; m.resize(mat_size);        // value-initialised (zeros)
        for (std::size_t i = 0; i < mat_size; ++i) {
            const int val = (int)sobol_mk[i];
            const std::size_t len = i + 1;
            digit_t digits[N];
            gf3_t::to_digit_array(digits, val, 3, (int)len);
            for (std::size_t j = 0; j < len; ++j) m[len - j - 1][i] = digits[j];
        }
        return m;
    }
    static int3_t point3_digits(const matrix_t& matrix, const int3_t& i3, int3_t& p3, int3_t& x3) {
        const std::size_t M = matrix.size();
        for (std::size_t k = 0; k < M; ++k) {
            if (p3.digits[k] != i3.digits[k]) {
                digit_t d = digit_t(i3.digits[k]) - digit_t(p3.digits[k]);
                d = int3_t::mod((int)d + 3);
                for (std::size_t j = 0; j < M; ++j)
                    x3.digits[j] = int3_t::fma((int)x3.digits[j], (int)d, (int)matrix[M - 1 - j][k]);
            }
        }
        p3 = i3;
        return x3;
    }
    struct rng_t {
        uint_t n{}; uint_t key{};
        explicit rng_t(uint_t s) : key((s << 1) | 1u) {}
        rng_t& index(uint_t i) { n = i; return *this; }
        uint_t sample() { return hash(++n * key); }
        uint_t sample_range(uint_t range) {
            uint_t divisor = ((-range) / range) + 1;
            if (divisor == 0) return 0;
            while (true) { uint_t x = sample() / divisor; if (x < range) return x; }
        }
        static uint_t hash(uint_t x) { x ^= x >> 16; x *= 0x21f0aaad; x ^= x >> 15; x *= 0xd35a2d97; x ^= x >> 15; return x; }
    };
    static int3_t scramble_base3(const int3_t& a3, uint_t seed, uint_t ndigits) {
        static constexpr std::int8_t scramble[6][3] = { {0,1,2}, {0,2,1}, {1,0,2}, {1,2,0}, {2,0,1}, {2,1,0} };
        rng_t rng(seed);
        int3_t b3;
        uint_t node_index = 0;
        for (uint_t i = 0; i < ndigits; ++i) {
            const uint_t flip = rng.index(node_index).sample_range(6);
            const uint_t digit = (uint_t)a3.digits[ndigits - 1 - i];
            b3.digits[ndigits - 1 - i] = scramble[flip][digit];
            node_index = 3 * node_index + 1 + digit;
        }
        return b3;
    }
    // generate_points (sobolld_sampler.hpp:59-99): `seeds` replaces the D draws of the caller's rng; output point-major [i*D + d].
    // Also returns the integer numerators value(M) so that parity can be checked on integers.
    void generate_points(const uint_t* seeds, std::size_t max_sample_count, std::vector<float>& samples, std::vector<std::uint32_t>* numerators = nullptr) const {
        const std::size_t M = matrix[0].size();
        const std::size_t sample_count = std::min<std::size_t>((std::size_t)pow3tab[M], max_sample_count);
        samples.clear(); samples.reserve(sample_count * D);
        std::array<int3_t, D> x3, p3;
        for (std::size_t d = 0; d < D; ++d) {
            const int3_t y = scramble_base3(x3[d], seeds[d], M);
            samples.push_back(y.value_fp(M));
            if (numerators) numerators->push_back((std::uint32_t)y.value(M));
        }
        for (std::size_t i = 1; i < sample_count; ++i) {
            const int3_t i3{ (uint_t)i };
            for (std::size_t d = 0; d < D; ++d) {
                const int3_t x = point3_digits(matrix[d], i3, p3[d], x3[d]);
                const int3_t y = scramble_base3(x, seeds[d], M);
                samples.push_back(y.value_fp(M));
                if (numerators) numerators->push_back((std::uint32_t)y.value(M));
            }
        }
    }
};

} } // namespace ot::sobol
