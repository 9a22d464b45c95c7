// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_ads.h: CPU restatement of the 8-wide BVH traversal (ray / shadow ray / cone) of
// /root/reference/src/ads/bvh8w.cpp and the work/record helpers of include/wt/ads/traversal_common.hpp,
// reading the flattened tables of include/wtgpu.h (the tables are data, shared by oracle and product).
#pragma once
#include <atomic>
#include "ot_math.h"
#include "../include/wtgpu.h"
#include <vector>
#include <set>

namespace ot {

struct ads_t {
    const wtgpu_scene_desc* d;
    explicit ads_t(const wtgpu_scene_desc* desc) : d(desc) {}

    v3 tri_a(uint32_t t) const { const auto& q = d->tris[t]; return { q.ax, q.ay, q.az }; }
    v3 tri_b(uint32_t t) const { const auto& q = d->tris[t]; return { q.bx, q.by, q.bz }; }
    v3 tri_c(uint32_t t) const { const auto& q = d->tris[t]; return { q.cx, q.cy, q.cz }; }
    v3 tri_n(uint32_t t) const { const auto& q = d->tris[t]; return { q.nx, q.ny, q.nz }; }
};

// bvh8w.cpp:44-57
struct stack_node_ptr_t { f_t min_range; int32_t ptr; };
inline void stack_sorter(stack_node_ptr_t* stack, int size) {
    for (int i = 1; i < size; ++i) {
        const auto p = stack[i];
        int j;
        for (j = i - 1; j >= 0 && p.min_range > stack[j].min_range; --j) stack[j + 1] = stack[j];
        stack[j + 1] = p;
    }
}

struct ray_hit_t {
    uint32_t tuid = WTGPU_INVALID_IDX;
    f_t dist = inf;
    v2 bary{ -1, -1 };
    bool front_face = false;
    bool empty() const { return tuid == WTGPU_INVALID_IDX; }
};

struct ads_counters_t { uint64_t nodes = 0, tris = 0, ray_casts = 0, cone_casts = 0, shadow_casts = 0; };

// ---- ray traversal: bvh8w.cpp:394-603
template <bool shadow>
inline bool ray_gather_tris(const ads_t& ads, const ray_t& ray, uint32_t t0, uint32_t count, range_t range, ray_hit_t& rec, ads_counters_t* ctr) {
    bool intersects = false;
    for (uint32_t t = 0; t < count; ++t) {
        const uint32_t tuid = t0 + t;
        const v3 a = ads.tri_a(tuid), b = ads.tri_b(tuid), c = ads.tri_c(tuid);
        if (ctr) ctr->tris++;
        if constexpr (shadow) {
            if (test_ray_tri_w(ray.o, ray.d, a, b, c, range)) { rec.dist = range.min; return true; }
            continue;
        }
        const auto r = intersect_ray_tri_w(ray.o, ray.d, a, b, c, range);
        const bool intrs = r.result != -inf;
        if (intrs && r.result < rec.dist) {
            rec.dist = r.result; rec.bary = { r.baryx, r.baryy }; rec.tuid = tuid;
            rec.front_face = dot(ads.tri_n(tuid), ray.d) <= 0;
            intersects = true;
        }
    }
    return intersects;
}

// intersect_ray_aabb_fast, one lane of the 8-wide form (intersect/ray.hpp:331-351).  The slabs are chosen by the SIGN BIT of 1/d (a blendv on
// the raw bits); vmaxps / vminps return their second operand when the comparison is false -- also when either is NaN (0 * inf: origin on a slab
// plane of an axis the ray is parallel to) -- and the four-argument forms pair up as (x, y), (z, range) (simd/math.hpp:333-356).
struct ray_aabb_fast_t { bool mask; f_t min, max; };
inline ray_aabb_fast_t ray_aabb_fast(v3 ro, v3 invd, v3 mn, v3 mx, range_t range) {
    const bool nx = std::signbit(invd.x), ny = std::signbit(invd.y), nz = std::signbit(invd.z);
    const f_t mnx = nx ? mx.x : mn.x, mxx = nx ? mn.x : mx.x;
    const f_t mny = ny ? mx.y : mn.y, mxy = ny ? mn.y : mx.y;
    const f_t mnz = nz ? mx.z : mn.z, mxz = nz ? mn.z : mx.z;
    const f_t t1x = (mnx - ro.x) * invd.x, t1y = (mny - ro.y) * invd.y, t1z = (mnz - ro.z) * invd.z;
    const f_t t2x = (mxx - ro.x) * invd.x, t2y = (mxy - ro.y) * invd.y, t2z = (mxz - ro.z) * invd.z;
    auto vmax = [](f_t a, f_t b) { return a > b ? a : b; };
    auto vmin = [](f_t a, f_t b) { return a < b ? a : b; };
    const f_t rmin = vmax(vmax(t1x, t1y), vmax(t1z, range.min));
    const f_t rmax = vmin(vmin(t2x, t2y), vmin(t2z, range.max));
    return { rmin <= rmax, rmin, rmax };
}
template <bool shadow>
inline bool ray_traverse(const ads_t& ads, const ray_t& ray, range_t range, ray_hit_t& rec, ads_counters_t* ctr) {
    constexpr int stack_size = 64;
    stack_node_ptr_t stack[stack_size];
    int s = 1;
    stack[0] = { 0, ads.d->root_ptr };
    auto unwind = [&]() { while (s > 0 && stack[s - 1].min_range >= rec.dist) --s; };
    while (s > 0) {
        const int32_t ptr = stack[s - 1].ptr;
        if (ptr < 0) {
            const wtgpu_leaf& leaf = ads.d->leaves[-ptr - 1];
            const bool intr = ray_gather_tris<shadow>(ads, ray, leaf.tris_ptr, leaf.count, range, rec, ctr);
            --s;
            if (intr) { if constexpr (shadow) return true; unwind(); }
        } else {
            const wtgpu_node& n = ads.d->nodes[ptr - 1];
            --s;
            if (ctr) ctr->nodes++;
            if (n.tris_count <= 16) {       // ray_traversal_treat_node_as_leaf_if_triangle_count_lt, bvh8w.cpp:29,512-526
                const bool intr = ray_gather_tris<shadow>(ads, ray, n.tris_start, n.tris_count, range, rec, ctr);
                if (intr) { if constexpr (shadow) return true; unwind(); }
                continue;
            }
            const int begin = s;
            for (int i = 0; i < 8; ++i) {
                const ray_aabb_fast_t hit = ray_aabb_fast(ray.o, ray.invd, { n.minx[i], n.miny[i], n.minz[i] }, { n.maxx[i], n.maxy[i], n.maxz[i] }, { 0.f, rec.dist });
                const f_t rmin = hit.min, rmax = hit.max;
                if (rmin <= rmax && n.child[i] != 0) stack[s++] = { rmin, n.child[i] };
            }
            stack_sorter(&stack[begin], s - begin);
        }
    }
    return rec.dist < inf;
}

// bvh8w.cpp:556-580 + traversal_common.hpp:93-110
inline ray_hit_t intersect_ray(const ads_t& ads, const ray_t& ray, range_t range, ads_counters_t* ctr = nullptr) {
    ray_hit_t rec;
    if (ctr) ctr->ray_casts++;
    ray_traverse<false>(ads, ray, range, rec, ctr);
    if (!std::isfinite(rec.dist) || rec.dist > range.max) return {};
    return rec;
}
// bvh8w.cpp:582-603
inline bool shadow_ray(const ads_t& ads, const ray_t& ray, range_t range, ads_counters_t* ctr = nullptr) {
    ray_hit_t rec;
    if (ctr) ctr->shadow_casts++;
    ray_traverse<true>(ads, ray, range, rec, ctr);
    return rec.dist < inf;
}

// ---- cone traversal: bvh8w.cpp:107-347, traversal_common.hpp:60-149
struct cone_record_t {
    f_t dist = -inf;
    bool front_face = false;
    std::vector<uint32_t> tris;
    std::vector<uint32_t> edges;       // ascending edge id == iteration order of std::set<tuid_t>
    bool empty() const { return tris.empty(); }
};

// diagnostic: histogram of cone-query sizes (log2 buckets): [0..31] queries by triangles TESTED, [32..63] queries by triangles ACCEPTED,
// [64..95] triangles tested summed per tested-bucket.  Filled only while g_cone_hist_on (tools/cone_hist.py).
inline std::atomic<uint64_t> g_cone_hist[96];
inline std::atomic<int> g_cone_hist_on{0};
inline std::atomic<uint64_t> g_cone_late[2];   // [0] queries whose closest distance still improved after 512 accepted triangles, [1] further improvements
inline int cone_hist_bucket(uint64_t n) { int b = 0; while (n) { ++b; n >>= 1; } return b; }
struct cone_work_t {
    std::vector<uint32_t> triangles;
    f_t intr_dist = inf;
    f_t z_search_range_scale = 1;
    bool front_face = false;
    range_t searchrange{ 0, inf };
    int late = 0;
    range_t search_range(const elliptic_cone_t& cone) const {        // traversal_common.hpp:78-84
        const f_t dist = std::max(searchrange.min, intr_dist);
        const f_t z_dist = cone.axes(dist).x * z_search_range_scale;
        return range_t{ searchrange.min, std::min(searchrange.max, dist + z_dist) } & range_t::positive();
    }
};

inline bool cone_gather_tris(const ads_t& ads, const elliptic_cone_t& cone, range_t range, uint32_t t0, uint32_t tcount, cone_work_t& rec, ads_counters_t* ctr) {
    bool found = false;
    for (uint32_t t = 0; t < tcount; ++t) {
        const uint32_t tuid = t0 + t;
        const v3 n = ads.tri_n(tuid);
        const bool front_face = dot(n, -cone.d()) > 0;
        if (ctr) ctr->tris++;
        const auto intrs = intersect_cone_tri(cone, ads.tri_a(tuid), ads.tri_b(tuid), ads.tri_c(tuid), n, range);
        if (intrs) {
            const f_t dist = intrs->dist;
            if (dist > range.max) continue;
            if (dist < rec.intr_dist) { if (g_cone_hist_on.load(std::memory_order_relaxed) && rec.triangles.size() >= 512) g_cone_late[rec.late++ ? 1 : 0]++; rec.intr_dist = dist; rec.front_face = front_face; }
            found = true;
            rec.triangles.push_back(tuid);
        }
    }
    return found;
}

// cone_cluster_intersect, bvh8w.cpp:187-230, one lane: the box enlarged by the cone's radius at the box's farthest depth, slab-tested against the axis
inline bool cone_cluster_lane(v3 ro, v3 rd, v3 rinvd, f_t ta, f_t ix, v3 mn, v3 mx, range_t range, f_t& tmin_out) {
    f_t omnx = mn.x - ro.x, omny = mn.y - ro.y, omnz = mn.z - ro.z;
    f_t omxx = mx.x - ro.x, omxy = mx.y - ro.y, omxz = mx.z - ro.z;
    const bool sx = std::signbit(rinvd.x), sy = std::signbit(rinvd.y), sz = std::signbit(rinvd.z);
    auto vmax = [](f_t a, f_t b) { return a > b ? a : b; };         // vmaxps / vminps: the second operand unless the comparison holds
    auto vmin = [](f_t a, f_t b) { return a < b ? a : b; };
    // b = selectv(max, min, sign(rinvd)): min where negative
    const f_t bx = sx ? omnx : omxx, by = sy ? omny : omxy, bz = sz ? omnz : omxz;
    const f_t dot_d_b = std::fma(rd.z, bz, std::fma(rd.y, by, rd.x * bx));
    const f_t maxz = vmin(vmax(dot_d_b, 0.f), range.max);          // wide clamp, simd/math.hpp:358-367
    const f_t enlr = std::fma(maxz, ta, ix);
    omnx -= enlr; omny -= enlr; omnz -= enlr;
    omxx += enlr; omxy += enlr; omxz += enlr;
    const f_t aex = sx ? omxx : omnx, aey = sy ? omxy : omny, aez = sz ? omxz : omnz;
    const f_t bex = sx ? omnx : omxx, bey = sy ? omny : omxy, bez = sz ? omnz : omxz;
    const f_t dminx = aex * rinvd.x, dminy = aey * rinvd.y, dminz = aez * rinvd.z;
    const f_t dmaxx = bex * rinvd.x, dmaxy = bey * rinvd.y, dmaxz = bez * rinvd.z;
    f_t tmin = 0, tmax = dmaxx;
    tmin = vmax(tmin, dminx);
    tmax = vmin(tmax, dmaxy);
    tmin = vmax(tmin, dminy);
    tmax = vmin(tmax, dmaxz);
    tmin = vmax(tmin, dminz);
    tmin_out = tmin;
    return tmin <= tmax && tmax >= range.min && tmin <= range.max;
}
inline cone_record_t intersect_cone(const ads_t& ads, const elliptic_cone_t& cone, range_t traversal_range, f_t z_scale, bool detect_edges, ads_counters_t* ctr = nullptr) {
    const uint64_t tris_before = ctr ? ctr->tris : 0;
    cone_work_t work; work.searchrange = traversal_range; work.z_search_range_scale = z_scale;
    if (ctr) ctr->cone_casts++;
    range_t range = work.search_range(cone);

    constexpr int stack_size = 128;
    stack_node_ptr_t stack[stack_size];
    int s = 1;
    stack[0] = { 0, ads.d->root_ptr };
    auto unwind = [&]() { range = work.search_range(cone); while (s > 0 && stack[s - 1].min_range >= range.max) --s; };

    const v3 ro = cone.o(), rd = cone.d(), rinvd = cone.r.invd;
    const f_t ta = cone.tan_alpha, ix = cone.x0;
    while (s > 0) {
        const int32_t ptr = stack[s - 1].ptr;
        if (ptr < 0) {
            const wtgpu_leaf& leaf = ads.d->leaves[-ptr - 1];
            const bool intr = cone_gather_tris(ads, cone, range, leaf.tris_ptr, leaf.count, work, ctr);
            --s;
            if (intr) unwind();
        } else {
            const wtgpu_node& n = ads.d->nodes[ptr - 1];
            --s;
            if (ctr) ctr->nodes++;
            const int begin = s;
            for (int i = 0; i < 8; ++i) {
                f_t tmin;
                const bool result = cone_cluster_lane(ro, rd, rinvd, ta, ix, { n.minx[i], n.miny[i], n.minz[i] }, { n.maxx[i], n.maxy[i], n.maxz[i] }, range, tmin);
                if (!result || n.child[i] == 0) continue;
                if (tmin >= range.max) continue;
                stack[s++] = { tmin, n.child[i] };
            }
            stack_sorter(&stack[begin], s - begin);
        }
    }

    if (ctr && g_cone_hist_on.load(std::memory_order_relaxed)) {
        const uint64_t tested = ctr->tris - tris_before;
        g_cone_hist[cone_hist_bucket(tested)]++; g_cone_hist[32 + cone_hist_bucket(work.triangles.size())]++; g_cone_hist[64 + cone_hist_bucket(tested)] += tested;
    }
    // cone_work_to_intersection_record, traversal_common.hpp:116-149 (work tri dist is value-initialised to 0: nothing is culled)
    cone_record_t ret;
    ret.dist = work.intr_dist; ret.front_face = work.front_face;
    std::set<uint32_t> edges;
    for (uint32_t tuid : work.triangles) {
        ret.tris.push_back(tuid);
        if (detect_edges) {
            const auto& m = ads.d->tri_meta[tuid];
            if (m.edge_ab != WTGPU_INVALID_IDX) edges.insert(m.edge_ab);
            if (m.edge_bc != WTGPU_INVALID_IDX) edges.insert(m.edge_bc);
            if (m.edge_ca != WTGPU_INVALID_IDX) edges.insert(m.edge_ca);
        }
    }
    ret.edges.assign(edges.begin(), edges.end());
    if (ret.tris.empty()) ret.dist = -inf;
    return ret;
}

} // namespace ot
