// sobol_tables.h -- host code shared by the CUDA library (wtgpu_scene_create uploads the tables to constant memory) and the host library
// (wthost_sobol_tables hands them to tests / tools): plain C++, no CUDA.
#pragma once
#include <cstdint>
#include <cstring>
#include <string>
#include "../../include/wtgpu.h"
namespace wt {
// Generator matrices of the sobolld sampler from the parsed table, as row masks for dsobol.cuh.
// Direction numbers m_1..m_11 of a dimension are kept as base-3 digit vectors v[c][t] (digit t of m_{c+1}); the first s_j come from the
// table, the rest from the recurrence over GF(3) of irreducible_gf3.hpp:103-118 written digit-wise:
//     v[c][t] = v[c-deg][t] + sum_{j=1..deg} g_j * v[c-j][t-j]   (mod 3),   g_j = -a_{deg-j}  (the reference's convert_to_gf3 = {0,2,1}).
// gen_mat (sobolld_sampler.hpp:140-154) then puts digit (c - r) of m_{c+1} at row r, column c (upper triangular).
inline bool sobol_build_tables(const wtgpu_sobol_entry* e, uint16_t (*t_ones)[WTGPU_SOBOL_DIGITS], uint16_t (*t_twos)[WTGPU_SOBOL_DIGITS], std::string& why) {
    const int M = (int)WTGPU_SOBOL_DIGITS;
    memset(t_ones, 0, sizeof(uint16_t) * WTGPU_SOBOL_DIMS * WTGPU_SOBOL_DIGITS); memset(t_twos, 0, sizeof(uint16_t) * WTGPU_SOBOL_DIMS * WTGPU_SOBOL_DIGITS);
    for (int dim = 0; dim < (int)WTGPU_SOBOL_DIMS; ++dim) {
        const wtgpu_sobol_entry& en = e[dim + 1];       // entry 0 is skipped by the reference (sobolld_sampler.hpp:50-52: d+1)
        const int deg = en.sj;
        if (en.d == 0 || deg < 1 || deg + 1 > M) { why = "sobol table: bad entry " + std::to_string(dim + 1); return false; }
        int poly[16] = { 0 };
        { int a = en.aj; for (int i = 0; i <= deg; ++i) { poly[i] = a % 3; a /= 3; } if (a != 0 || poly[deg] == 0) { why = "sobol table: polynomial/degree mismatch in entry " + std::to_string(dim + 1); return false; } }
        int v[WTGPU_SOBOL_DIGITS][WTGPU_SOBOL_DIGITS] = {};
        for (int c = 0; c < deg; ++c) {
            int m = en.mk[c], lim = 1; for (int q = 0; q <= c; ++q) lim *= 3;
            if (m <= 0 || m >= lim) { why = "sobol table: direction number out of range in entry " + std::to_string(dim + 1); return false; }
            for (int q = 0; q <= c; ++q) { v[c][q] = m % 3; m /= 3; }
        }
        for (int c = deg; c < M; ++c)
            for (int q = 0; q <= c; ++q) {
                int acc = v[c - deg][q];
                for (int j = 1; j <= deg; ++j) if (q >= j) acc += ((3 - poly[deg - j]) % 3) * v[c - j][q - j];
                v[c][q] = acc % 3;
            }
        for (int j = 0; j < M; ++j) {
            const int r = M - 1 - j;                    // output digit j reads matrix row M-1-j (sobolld_sampler.hpp:170)
            uint16_t o1 = 0, o2 = 0;
            for (int c = r; c < M; ++c) { const int x = v[c][c - r]; if (x == 1) o1 |= (uint16_t)(1u << c); else if (x == 2) o2 |= (uint16_t)(1u << c); }
            t_ones[dim][j] = o1; t_twos[dim][j] = o2;
        }
    }
    return true;
}

} // namespace wt
