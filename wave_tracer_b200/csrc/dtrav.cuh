// dtrav.cuh -- 8-wide BVH traversal on the device: closest-hit ray, any-hit (shadow) ray, elliptic-cone query.
//
// Follows the traversal order of bvh8w_t (/root/reference/src/ads/bvh8w.cpp:123-347 cone, 394-603 ray): ordered
// stack traversal, children pushed far-to-near after an insertion sort on tmin, range shrinking on hits -- so the
// triangle set a cone query returns is the same set the CPU path returns (hard part 1).
// Layout: nodes are 256-B records (8 AABBs SoA + 8 child pointers) fetched as 16-B vectors; the first
// `top_nodes` nodes (BFS order = top of the tree) may be staged in shared memory by the calling kernel.
#pragma once
#include "dmath.cuh"
#include "../../include/wtgpu.h"

namespace wt {

// Capacities of the per-path variable-length lists.  The reference keeps them in std::vector / std::set (traversal_common.hpp:116-149,
// fraunhofer/free_space_diffraction.cpp:22-129, vertex arenas of plt_bdpt.cpp:59-67); here they are rows of HBM arrays whose length is a RUN-TIME
// property of the scene handle: a render that finds a list longer than its row records how long it had to be (DevCounters::need_*), and
// wtgpu_render re-sizes the rows and renders again -- a two-pass count / fill at the granularity of the render call, so results never depend on a
// capacity (wavefront.cu, "capacity growth").
struct Caps {
    uint32_t tris;          // triangles of a cone query kept in the path's own row (kTriRow); longer lists continue in extents of the spill arena
    uint32_t spill_words;   // size of the spill arena (entries), shared by all paths of an iteration
    uint32_t edges;         // edges around a vertex (rows of hit_edges / the UTD aperture's edge list)
    uint32_t seg;           // segments of a Fraunhofer aperture
    uint32_t ap_walk;       // Fraunhofer apertures per subpath
    uint32_t verts;         // vertices per subpath (plt_bdpt: max_depth + 2)
    uint32_t ap_words, arena_words;     // derived: words per aperture record, per sample slot (dbdpt.cuh)
};
struct DScene {
    const wtgpu_node* nodes; const wtgpu_leaf* leaves; int32_t root_ptr;
    const float4* tris;               // 3 float4 per triangle
    const wtgpu_tri_meta* tri_meta; const wtgpu_tri_shading* tri_shading;
    const wtgpu_edge* edges;
    const wtgpu_shape* shapes; const uint32_t* shape_tri_tuid; const float* shape_tri_cdf;
    const wtgpu_spectrum* spectra; const float* spectrum_data;
    const wtgpu_bsdf* bsdfs; const wtgpu_bsdf_bin* bsdf_bins;
    const wtgpu_emitter* emitters; const float* emitter_cdf; const wtgpu_kdist* emitter_kdist; const float* kdist_data;
    uint32_t n_emitters, n_bsdfs, n_tris, n_nodes, n_edges_total;
    wtgpu_sensor sensor;
    wtgpu_integrator integrator;
    const float* erf_lut;             // 1024-entry erf table (include/wt/math/erf_lut.hpp)
    uint32_t scene_stream;            // Sampler::stream of the scene-sampler draws: 0 (uniform) or kSobolStreamFlag | spp (sobolld); set per render
    float ray_cull_abs;               // absolute slack of the ray-query range culling (RayCull below); +inf disables the culling; set per render
    Caps cap;                         // set per render
    // spill arena of the cone-query triangle lists (TriList below): entries, bump cursor (reset every iteration), extent tables (kTriExt per path)
    uint32_t* spill; unsigned int* spill_head; uint32_t* spill_ext;
};

// ---- triangle lists of cone queries.  The reference returns a std::vector<tuid_t> of any length (traversal_common.hpp:116-149); most hold a few
// triangles, a few hold 10^5 (a beam several millimetres wide over a finely tessellated mesh).  A list lives for one iteration (written by the
// traversal kernel, read by the resolve kernel): its first kTriRow entries are in the path's own row, the rest in extents of doubling size
// (256, 512, 1024, ...) bump-allocated from one arena that is reset every iteration; the path's extent table gives O(1) random access.
constexpr uint32_t kTriRow = 128u, kTriExt0 = 256u, kTriExt = 22u;
struct TriList { const uint32_t* row; const uint32_t* spill; const uint32_t* ext; uint32_t n; };
WT_D void tri_locate(uint32_t i, uint32_t& e, uint32_t& off) {     // entry i >= kTriRow -> extent, offset
    const uint32_t j = i - kTriRow;
    e = 31u - (uint32_t)__clz(j / kTriExt0 + 1u);
    off = j - kTriExt0 * ((1u << e) - 1u);
}
WT_D uint32_t tri_at(const TriList& l, uint32_t i) {
    if (i < kTriRow) return l.row[i];
    uint32_t e, off; tri_locate(i, e, off);
    return l.spill[l.ext[e] + off];
}
// writer side: `alloc_end` entries are backed by storage; reserve() extends it (one thread does the bump allocation), put() stores entry i
struct TriWriter { uint32_t* row; uint32_t* ext; uint32_t n_ext, alloc_end; bool fail; };
WT_D void tw_begin(TriWriter& w) { w.n_ext = 0u; w.alloc_end = kTriRow; w.fail = false; }       // (extents of an abandoned list stay allocated until the iteration ends)
// called by ONE thread: back entries [0, upto) with storage
WT_D void tw_reserve(const DScene& sc, TriWriter& w, uint32_t upto) {
    while (upto > w.alloc_end && !w.fail) {
        if (!w.ext) { w.fail = true; break; }
        const uint32_t size = kTriExt0 << w.n_ext;
        const uint32_t base = atomicAdd(sc.spill_head, size);      // (the cursor keeps counting past the arena's end: it measures the demand)
        if (w.n_ext >= kTriExt || base > sc.cap.spill_words || size > sc.cap.spill_words - base) { w.fail = true; break; }
        w.ext[w.n_ext++] = base; w.alloc_end += size;
    }
}
WT_D void tw_put(const DScene& sc, const TriWriter& w, uint32_t i, uint32_t tuid) {
    if (i < kTriRow) { w.row[i] = tuid; return; }
    if (i >= w.alloc_end) return;
    uint32_t e, off; tri_locate(i, e, off);
    sc.spill[w.ext[e] + off] = tuid;
}

// Range culling of RAY queries.  bvh8w.cpp:469-554 tests nodes against {0, closest hit} only, so a ray cast over a short range (a
// ballistic segment of traverse(), a shadow ray) still walks every node along the infinite ray and rejects the triangles one by one
// (intersect_ray_tri's range check).  A child whose slab interval [rmin, rmax] lies outside the query range cannot contain an accepted
// hit, so it is not pushed: same hits, same order of the remaining stack, far fewer node visits.  The slack covers the rounding of the
// two distance computations (slab vs. watertight ray-triangle: a few ulp of the largest vertex distance): 1e-4 relative + 1e-5 x the
// largest |coordinate| of the scene.
struct RayCull { float mn, mx; };
WT_D RayCull ray_cull(const DScene& sc, Range r) {
    RayCull c; c.mx = r.mx + (1e-4f * r.mx + sc.ray_cull_abs); c.mn = r.mn - (1e-4f * fabsf(r.mn) + sc.ray_cull_abs);
    return c;
}
WT_D bool ray_cull_keep(const RayCull& c, float rmin, float rmax) { return !(rmin > c.mx) && !(rmax < c.mn); }

struct Counters {       // per-thread, flushed with one atomic per counter per warp
    uint32_t nodes, tris, ray_casts, cone_casts, shadow_casts;
    uint32_t stack_drops;   // children a full traversal stack could not take (the reference's stacks are as deep: 64 ray / 128 cone, bvh8w.cpp:305-307 exits in a debug build)
};
WT_D void counters_zero(Counters& c) { c.nodes = c.tris = c.ray_casts = c.cone_casts = c.shadow_casts = c.stack_drops = 0; }

struct Tri3 { V3 a, b, c, n; };
WT_D Tri3 load_tri(const DScene& sc, uint32_t tuid) {
    const float4 q0 = __ldg(sc.tris + 3 * tuid), q1 = __ldg(sc.tris + 3 * tuid + 1), q2 = __ldg(sc.tris + 3 * tuid + 2);
    Tri3 t; t.a = mk3(q0.x, q0.y, q0.z); t.b = mk3(q1.x, q1.y, q1.z); t.c = mk3(q2.x, q2.y, q2.z); t.n = mk3(q0.w, q1.w, q2.w);
    return t;
}

struct StackEnt { float tmin; int32_t ptr; };
WT_D void stack_sort(StackEnt* st, int n) {       // bvh8w.cpp:44-57: descending tmin, stable insertion sort
    for (int i = 1; i < n; ++i) {
        const StackEnt p = st[i];
        int j;
        for (j = i - 1; j >= 0 && p.tmin > st[j].tmin; --j) st[j + 1] = st[j];
        st[j + 1] = p;
    }
}

struct RayHit { uint32_t tuid; float dist; float bx, by; bool front; };

template <bool SHADOW>
WT_D bool ray_gather(const DScene& sc, V3 ro, V3 rd, uint32_t t0, uint32_t cnt, Range range, RayHit& rec, Counters& ctr) {
    bool any = false;
    for (uint32_t t = 0; t < cnt; ++t) {
        const uint32_t tuid = t0 + t;
        const Tri3 tr = load_tri(sc, tuid);
        ctr.tris++;
        if (SHADOW) {
            if (test_ray_tri_w(ro, rd, tr.a, tr.b, tr.c, range)) { rec.dist = range.mn; return true; }
            continue;
        }
        float bx, by;
        const float z = intersect_ray_tri_w(ro, rd, tr.a, tr.b, tr.c, range, bx, by);
        if (z != -WT_INF && z < rec.dist) { rec.dist = z; rec.bx = bx; rec.by = by; rec.tuid = tuid; rec.front = dot(tr.n, rd) <= 0.f; any = true; }
    }
    return any;
}

// Closest-hit (SHADOW=false) / any-hit (SHADOW=true) ray traversal.  bvh8w.cpp:469-554.
template <bool SHADOW>
WT_DN bool ray_traverse(const DScene& sc, V3 ro, V3 rd, Range range, RayHit& rec, Counters& ctr) {
    rec.tuid = WTGPU_INVALID_IDX; rec.dist = WT_INF; rec.bx = rec.by = -1.f; rec.front = false;
    const V3 inv = mk3(1.f / rd.x, 1.f / rd.y, 1.f / rd.z);
    const bool nx = signbit(inv.x), ny = signbit(inv.y), nz = signbit(inv.z);
    const RayCull cull = ray_cull(sc, range);
    StackEnt stack[64];
    int s = 1;
    stack[0].tmin = 0.f; stack[0].ptr = sc.root_ptr;
    while (s > 0) {
        const int32_t ptr = stack[--s].ptr;
        bool hit = false;
        if (ptr < 0) {
            const wtgpu_leaf lf = sc.leaves[-ptr - 1];
            hit = ray_gather<SHADOW>(sc, ro, rd, lf.tris_ptr, lf.count, range, rec, ctr);
        } else {
            const wtgpu_node* __restrict__ n = sc.nodes + (ptr - 1);
            ctr.nodes++;
            const uint2 tr = __ldg(reinterpret_cast<const uint2*>(&n->tris_start));
            if (tr.y <= 16u) {      // ray_traversal_treat_node_as_leaf_if_triangle_count_lt (bvh8w.cpp:29)
                hit = ray_gather<SHADOW>(sc, ro, rd, tr.x, tr.y, range, rec, ctr);
            } else {
                const int begin = s;
                const float4* __restrict__ q = reinterpret_cast<const float4*>(n);
#pragma unroll 1        // two passes over four children instead of eight unrolled slab tests: this function is inlined at every shadow-ray site (etoile-like k_shade -8 %)
                for (int h = 0; h < 2; ++h) {
                    const float4 mnx = __ldg(q + 0 + h), mny = __ldg(q + 2 + h), mnz = __ldg(q + 4 + h);
                    const float4 mxx = __ldg(q + 6 + h), mxy = __ldg(q + 8 + h), mxz = __ldg(q + 10 + h);
                    const int4 ch = __ldg(reinterpret_cast<const int4*>(q + 12 + h));
                    const float amnx[4] = { mnx.x, mnx.y, mnx.z, mnx.w }, amny[4] = { mny.x, mny.y, mny.z, mny.w }, amnz[4] = { mnz.x, mnz.y, mnz.z, mnz.w };
                    const float amxx[4] = { mxx.x, mxx.y, mxx.z, mxx.w }, amxy[4] = { mxy.x, mxy.y, mxy.z, mxy.w }, amxz[4] = { mxz.x, mxz.y, mxz.z, mxz.w };
                    const int ach[4] = { ch.x, ch.y, ch.z, ch.w };
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        // intersect_ray_aabb_fast (intersect/ray.hpp:331-351), range {0, closest hit}
                        const float t1x = ((nx ? amxx[i] : amnx[i]) - ro.x) * inv.x, t2x = ((nx ? amnx[i] : amxx[i]) - ro.x) * inv.x;
                        const float t1y = ((ny ? amxy[i] : amny[i]) - ro.y) * inv.y, t2y = ((ny ? amny[i] : amxy[i]) - ro.y) * inv.y;
                        const float t1z = ((nz ? amxz[i] : amnz[i]) - ro.z) * inv.z, t2z = ((nz ? amnz[i] : amxz[i]) - ro.z) * inv.z;
                        const float rmin = vmaxps(vmaxps(t1x, t1y), vmaxps(t1z, 0.f));        // (x, y), (z, range): simd/math.hpp:333-356 -- decides when a slab gives NaN
                        const float rmax = vminps(vminps(t2x, t2y), vminps(t2z, rec.dist));
                        if (rmin <= rmax && ach[i] != 0 && ray_cull_keep(cull, rmin, rmax)) { if (s < 64) { stack[s].tmin = rmin; stack[s].ptr = ach[i]; ++s; } else ctr.stack_drops++; }
                    }
                }
                stack_sort(stack + begin, s - begin);
            }
        }
        if (hit) {
            if (SHADOW) return true;
            while (s > 0 && stack[s - 1].tmin >= rec.dist) --s;
        }
    }
    return rec.dist < WT_INF;
}

// ads_t::intersect(ray, range) incl. ray_work_to_intersection_record (traversal_common.hpp:93-110)
WT_D bool intersect_ray(const DScene& sc, V3 ro, V3 rd, Range range, RayHit& rec, Counters& ctr) {
    ctr.ray_casts++;
    ray_traverse<false>(sc, ro, rd, range, rec, ctr);
    if (!isfinite(rec.dist) || rec.dist > range.mx) { rec.tuid = WTGPU_INVALID_IDX; rec.dist = WT_INF; return false; }
    return true;
}
WT_D bool shadow_ray(const DScene& sc, V3 ro, V3 rd, Range range, Counters& ctr) {
    ctr.shadow_casts++;
    RayHit rec;
    ray_traverse<true>(sc, ro, rd, range, rec, ctr);
    return rec.dist < WT_INF;
}

// ---- cone query.  Results: closest distance, front-face flag of the closest triangle, the triangle list in traversal order.
struct ConeResult { float dist; bool front; uint32_t n_tris; bool overflow; };      // n_tris counts every accepted triangle, also those beyond the row

WT_D Range cone_search_range(const Cone& cone, Range searchrange, float intr_dist, float z_scale) {    // traversal_common.hpp:78-84
    const float dist = fmaxf(searchrange.mn, intr_dist);
    const float zd = cone_axes(cone, dist).x * z_scale;
    return rand_(mkr(searchrange.mn, fminf(searchrange.mx, dist + zd)), mkr(0.f, WT_INF));
}

WT_DN void cone_traverse(const DScene& sc, const Cone& cone, Range traversal_range, float z_scale, TriWriter& tw, ConeResult& res, Counters& ctr) {
    ctr.cone_casts++;
    res.dist = WT_INF; res.front = false; res.n_tris = 0; res.overflow = false;
    tw_begin(tw);
    Range range = cone_search_range(cone, traversal_range, res.dist, z_scale);
    const Frame frame = cone_frame(cone);
    const V3 ro = cone.o, rd = cone.d;
    const V3 inv = mk3(1.f / rd.x, 1.f / rd.y, 1.f / rd.z);
    const bool sx = signbit(inv.x), sy = signbit(inv.y), sz = signbit(inv.z);
    StackEnt stack[128];
    int s = 1;
    stack[0].tmin = 0.f; stack[0].ptr = sc.root_ptr;
    while (s > 0) {
        const int32_t ptr = stack[--s].ptr;
        if (ptr < 0) {
            const wtgpu_leaf lf = sc.leaves[-ptr - 1];
            bool found = false;
            for (uint32_t t = 0; t < lf.count; ++t) {     // gather_tris (bvh8w.cpp:123-185)
                const uint32_t tuid = lf.tris_ptr + t;
                const Tri3 tr = load_tri(sc, tuid);
                ctr.tris++;
                const float d = intersect_cone_tri(cone, frame, tr.a, tr.b, tr.c, tr.n, range);
                if (d < WT_INF) {
                    if (d > range.mx) continue;
                    if (d < res.dist) { res.dist = d; res.front = dot(tr.n, -rd) > 0.f; }
                    found = true;
                    tw_reserve(sc, tw, res.n_tris + 1u);
                    if (tw.fail) res.overflow = true; else tw_put(sc, tw, res.n_tris, tuid);
                    res.n_tris++;
                }
            }
            if (found) {
                range = cone_search_range(cone, traversal_range, res.dist, z_scale);
                while (s > 0 && stack[s - 1].tmin >= range.mx) --s;
            }
        } else {
            const wtgpu_node* __restrict__ n = sc.nodes + (ptr - 1);
            ctr.nodes++;
            const int begin = s;
            const float4* __restrict__ q = reinterpret_cast<const float4*>(n);
#pragma unroll
            for (int h = 0; h < 2; ++h) {
                const float4 mnx = __ldg(q + 0 + h), mny = __ldg(q + 2 + h), mnz = __ldg(q + 4 + h);
                const float4 mxx = __ldg(q + 6 + h), mxy = __ldg(q + 8 + h), mxz = __ldg(q + 10 + h);
                const int4 ch = __ldg(reinterpret_cast<const int4*>(q + 12 + h));
                const float amnx[4] = { mnx.x, mnx.y, mnx.z, mnx.w }, amny[4] = { mny.x, mny.y, mny.z, mny.w }, amnz[4] = { mnz.x, mnz.y, mnz.z, mnz.w };
                const float amxx[4] = { mxx.x, mxx.y, mxx.z, mxx.w }, amxy[4] = { mxy.x, mxy.y, mxy.z, mxy.w }, amxz[4] = { mxz.x, mxz.y, mxz.z, mxz.w };
                const int ach[4] = { ch.x, ch.y, ch.z, ch.w };
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    // cone_cluster_intersect (bvh8w.cpp:187-230): inflate the AABB by the cone radius at its farthest z, slab test
                    float omnx = amnx[i] - ro.x, omny = amny[i] - ro.y, omnz = amnz[i] - ro.z;
                    float omxx = amxx[i] - ro.x, omxy = amxy[i] - ro.y, omxz = amxz[i] - ro.z;
                    const float bx = sx ? omnx : omxx, by = sy ? omny : omxy, bz = sz ? omnz : omxz;
                    const float ddb = fmaf(rd.z, bz, fmaf(rd.y, by, rd.x * bx));
                    const float maxz = fminf(fmaxf(ddb, 0.f), range.mx);
                    const float enl = fmaf(maxz, cone.ta, cone.x0);
                    omnx -= enl; omny -= enl; omnz -= enl; omxx += enl; omxy += enl; omxz += enl;
                    const float dminx = (sx ? omxx : omnx) * inv.x, dminy = (sy ? omxy : omny) * inv.y, dminz = (sz ? omxz : omnz) * inv.z;
                    const float dmaxx = (sx ? omnx : omxx) * inv.x, dmaxy = (sy ? omny : omxy) * inv.y, dmaxz = (sz ? omnz : omxz) * inv.z;
                    float tmin = 0.f, tmax = dmaxx;
                    tmin = vmaxps(tmin, dminx); tmax = vminps(tmax, dmaxy);
                    tmin = vmaxps(tmin, dminy); tmax = vminps(tmax, dmaxz);
                    tmin = vmaxps(tmin, dminz);
                    const bool ok = tmin <= tmax && tmax >= range.mn && tmin <= range.mx;
                    if (!ok || ach[i] == 0) continue;
                    if (tmin >= range.mx) continue;
                    if (s < 128) { stack[s].tmin = tmin; stack[s].ptr = ach[i]; ++s; } else ctr.stack_drops++;
                }
            }
            stack_sort(stack + begin, s - begin);
        }
    }
}

// edges of a triangle list, deduplicated and sorted ascending (the iteration order of std::set<tuid_t>,
// traversal_common.hpp:124-146)
WT_D uint32_t collect_edges(const DScene& sc, const TriList& tris, uint32_t* edges, uint32_t max_edges, bool& overflow) {
    uint32_t n = 0;
    for (uint32_t i = 0; i < tris.n; ++i) {
        const wtgpu_tri_meta m = sc.tri_meta[tri_at(tris, i)];
#pragma unroll 1
        for (int k = 0; k < 3; ++k) {       // (one copy of the insertion: code size)
            const uint32_t e = k == 0 ? m.edge_ab : k == 1 ? m.edge_bc : m.edge_ca;
            if (e == WTGPU_INVALID_IDX) continue;
            uint32_t pos = 0;
            while (pos < n && edges[pos] < e) ++pos;
            if (pos < n && edges[pos] == e) continue;
            if (n >= max_edges) { overflow = true; continue; }
            for (uint32_t j = n; j > pos; --j) edges[j] = edges[j - 1];
            edges[pos] = e; ++n;
        }
    }
    return n;
}

} // namespace wt
