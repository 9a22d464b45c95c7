// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_rng.h: the counter-based sampling contract (ours -- the reference seeds mt19937_64 from
// random_device^tid^time and has no user seed, include/wt/util/seeded_mt19937_64.hpp:31-50) and the
// sampler warps of include/wt/sampler/sampler.hpp:78-306.
//
// Stream definition: draw number d of stream (seed, pixel, sample) is lane (d&3) of
// Philox4x32-10(key = (seed_lo, seed_hi), counter = (d>>2, sample, pixel, stream)); float = (u32>>8)*2^-24.
// stream is 0 for plt_path.  plt_bdpt splits a sample into independent sub-streams so that the two subpath walks and every (s,t)
// connection can run concurrently on the device: 0 = emitter/wavenumber/source sampling, 1 = sensor subpath walk, 2 = emitter
// subpath walk, 3 + 32 t + s = connection (s,t); each sub-stream starts at d = 0.
#pragma once
#include "ot_math.h"
#include "ot_sobol.h"
#include <map>
#include <memory>
#include <mutex>

namespace ot {

inline void philox4x32_10(uint32_t c[4], uint32_t k0, uint32_t k1) {
    for (int r = 0; r < 10; ++r) {
        const uint64_t p0 = (uint64_t)0xD2511F53u * c[0];
        const uint64_t p1 = (uint64_t)0xCD9E8D57u * c[2];
        const uint32_t n0 = (uint32_t)(p1 >> 32) ^ c[1] ^ k0;
        const uint32_t n1 = (uint32_t)p1;
        const uint32_t n2 = (uint32_t)(p0 >> 32) ^ c[3] ^ k1;
        const uint32_t n3 = (uint32_t)p0;
        c[0] = n0; c[1] = n1; c[2] = n2; c[3] = n3;
        k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
    }
}

// The sobolld scene sampler under our seeding contract (include/wtgpu.h "sobolld contract"): whole batches are produced by the
// literal generate_points() restatement (ot_sobol.h) and cached; a draw indexes the flat point-major batch exactly as
// sobolld_t::next_sample does (include/wt/sampler/sobolld.hpp:40-48).
struct sobol_ctx_t {
    sobol::gf3_t gf3;
    sobol::sobolls_sampler gen;
    mutable std::mutex m;
    mutable std::map<uint64_t, std::shared_ptr<std::vector<float>>> batches;
    explicit sobol_ctx_t(const wtgpu_sobol_entry* e) : gf3(e), gen(sobol::N, gf3) {}
    static void seeds_for_batch(uint64_t seed, uint64_t batch, uint64_t* out);
    std::shared_ptr<std::vector<float>> batch(uint64_t seed, uint64_t b) const {
        std::lock_guard<std::mutex> l(m);
        auto it = batches.find(b);
        if (it != batches.end()) return it->second;
        uint64_t seeds[sobol::D]; seeds_for_batch(seed, b, seeds);
        auto v = std::make_shared<std::vector<float>>();
        gen.generate_points(seeds, (size_t)-1, *v);
        if (batches.size() >= 4) batches.erase(batches.begin());
        batches[b] = v;
        return v;
    }
};

struct sampler_t {
    static constexpr uint32_t sobol_flag = 0x80000000u;
    const sobol_ctx_t* sob = nullptr;
    std::shared_ptr<std::vector<float>> sob_batch; uint64_t sob_batch_idx = ~0ull;
    uint64_t seed = 0;
    uint32_t pixel = 0, sample = 0;
    uint32_t d = 0;
    uint32_t stream = 0;
    uint32_t cached_block = 0xffffffffu;
    uint32_t cache[4];

    uint32_t next_u32() {
        const uint32_t block = d >> 2, lane = d & 3;
        if (block != cached_block) {
            cache[0] = block; cache[1] = sample; cache[2] = pixel; cache[3] = stream;
            philox4x32_10(cache, (uint32_t)seed, (uint32_t)(seed >> 32));
            cached_block = block;
        }
        ++d;
        return cache[lane];
    }
    void set_stream(uint32_t s) { stream = s; d = 0; cached_block = 0xffffffffu; }
    // scene sampler = sobolld: flag | spp in `stream` (the device carries it the same way, csrc/dscene.cuh rnd())
    void begin_scene_draws(const sobol_ctx_t* ctx, uint32_t spp) { if (ctx) { sob = ctx; stream = sobol_flag | spp; } }
    void end_scene_draws() { if (stream & sobol_flag) { stream = 0; cached_block = 0xffffffffu; } }     // path sampling continues on the Philox stream at the same d
    f_t sobol_r() {
        const uint64_t g = (uint64_t)pixel * (uint64_t)(stream & ~sobol_flag) + sample + d / sobol::D;
        const uint32_t dim = d % sobol::D; ++d;
        const uint64_t npts = sobol::pow3tab[sobol::N], b = g / npts, i = g % npts;
        if (b != sob_batch_idx) { sob_batch = sob->batch(seed, b); sob_batch_idx = b; }
        return (*sob_batch)[i * sobol::D + dim];
    }
    // test hook: replay a scripted sequence instead of the Philox stream (lets tests compare the ORDER of draws with the reference's own code)
    const float* script = nullptr; uint32_t script_n = 0;
    f_t r() {
        if (script) { const f_t v = d < script_n ? script[d] : .5f; ++d; return v; }
        if (stream & sobol_flag) return sobol_r();
        return (f_t)(next_u32() >> 8) * (1.0f / 16777216.0f);
    }
    v2 r2() { const f_t a = r(); const f_t b = r(); return { a, b }; }
    v3 r3() { const f_t a = r(); const f_t b = r(); const f_t c = r(); return { a, b, c }; }

    // sampler.hpp:106-109
    int uniform_int_interval(int start, int end) { return std::min(end - 1, int(r() * (end - start)) + start); }
};

inline void sobol_ctx_t::seeds_for_batch(uint64_t seed, uint64_t batch, uint64_t* out) {
    for (uint32_t d = 0; d < sobol::D; ++d) {
        uint32_t c[4] = { d, (uint32_t)batch, (uint32_t)(batch >> 32), 0x50B01Du };
        philox4x32_10(c, (uint32_t)seed, (uint32_t)(seed >> 32));
        out[d] = c[0];
    }
}

// sampler.hpp:139-150
inline v3 uniform_sphere(v2 u) {
    const f_t z = 1 - 2 * u.x;
    const f_t rr = std::sqrt(std::max(0.f, 1 - sqr(z)));
    const f_t phi = two_pi * u.y;
    return { rr * lm::cos(phi), rr * lm::sin(phi), z };
}
// sampler.hpp:164-181
inline v2 concentric_disk(v2 u) {
    const v2 offset = 2.f * u - v2{ 1, 1 };
    f_t rr, theta;
    if (offset.x == 0 && offset.y == 0) { rr = 0; theta = 0; }
    else if (std::fabs(offset.x) > std::fabs(offset.y)) { rr = offset.x; theta = pi_4 * (offset.y / offset.x); }
    else { rr = offset.y; theta = pi_2 - pi_4 * (offset.x / offset.y); }
    return rr * v2{ lm::cos(theta), lm::sin(theta) };
}
// sampler.hpp:197-203
inline v3 cosine_hemisphere(v2 u) {
    const v2 d = concentric_disk(u);
    const f_t z = std::sqrt(std::max(0.f, 1 - sqr(d.x) - sqr(d.y)));
    return { d.x, d.y, z };
}
inline f_t cosine_hemisphere_pdf(f_t cosine) { return inv_pi * cosine; }
// sampler.hpp:222-233
inline v3 uniform_cone(f_t solid_angle, v2 u) {
    const f_t cos_theta_max = 1 - inv_two_pi * solid_angle;
    const f_t cos_theta = 1 + u.x * (cos_theta_max - 1);
    const f_t sin_theta = std::sqrt(std::max(0.f, 1 - sqr(cos_theta)));
    const f_t phi = two_pi * u.y;
    return { lm::cos(phi) * sin_theta, lm::sin(phi) * sin_theta, cos_theta };
}
inline f_t uniform_cone_pdf(f_t solid_angle) { return 1.f / solid_angle; }
// sampler.hpp:253-260
inline v2 normal2d(v2 u) {
    const f_t r = std::sqrt(-2 * lm::log(1 - u.x));
    const f_t theta = two_pi * u.y;
    return { r * lm::cos(theta), r * lm::sin(theta) };
}
// sampler.hpp:281-286
inline v2 uniform_triangle(v2 u) { if (u.x + u.y > 1) u = v2{ 1, 1 } - u; return u; }

} // namespace ot
