// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The reference's BVH traversal loops themselves -- src/ads/bvh8w.cpp: the cone traversal (gather_tris :123-185, cone_cluster_intersect :186-230,
// traverse :232-318: stack of 128, nearest child popped first, search range shrinking with every accepted triangle, unwinding), and the ray /
// shadow-ray traversal (8-wide triangle clusters :64-100, gather_tris :394-452, ray_cluster_intersect :454-467, traverse :469-554: stack of 64,
// nodes of <= 16 triangles treated as leaves) -- with the work records and search_range() of include/wt/ads/traversal_common.hpp:21-149, the
// reference's own ads/common.hpp, ads/bvh8w/bvh8w_node.hpp, ads/bvh8w/common.hpp and, beneath them, everything oracle/ref_cone.cpp compiles
// (cone / ray tests, elliptic_cone.hpp, frame.hpp ...), the reference's own ads/intersection_record.hpp, the record conversions of
// traversal_common.hpp:90-149 (distance culling, edge sets) and, on top, the ballistic / diffusive state machine of include/wt/integrator/traversal.hpp:26-248
// (calculate_min_ballistic_distance, max_ballistic_distance, traverse, traverse_shadow) and, cut the same way further down in this file: both integrators' find_closest_triangle (plt_path_detail.hpp:244-276, plt_bdpt_detail.hpp:352-419), the self-intersection offsets of
// src/interaction/intersection.cpp, and the apertures of the two diffraction models (src/interaction/fsd/free_space_diffraction.cpp, fsd/fraunhofer/free_space_diffraction.cpp)
// -> oracle/_ref/libref_traverse.so.  tests/test_oracle_kats.py runs it over the BVH the HOST
// LAYER built for a scene and compares, per query, the accepted-triangle list IN TRAVERSAL ORDER, distances, barycentrics and faces with ot_ads.h.
// bvh8w.cpp as a whole needs tinybvh, the scene tree and the statistics collectors; so the Makefile writes the line ranges named above, as they are,
// to the git-ignored oracle/_ref/bvh8w_traverse_part.hpp / traversal_common_part.hpp at build time (deleted again once the library is linked) and this TU includes those.  What stands in here:
// the tree container (arrays + the five accessors the loops call), ads_t::intersect_opts_t, and the statistics wrappers of ads_stats.hpp reduced to
// their forwarding line.  Wide vectors are the shim's arrays of lanes (WT_SHIM_WIDE_LANES).
#define WT_SHIM_DISTINCT_PQ
#define WT_SHIM_WIDE_LANES
#define WT_SHIM_MM_UNIT
#define RELEASE
#include <wt/util/assert.hpp>
#include <cstdint>
#include <bitset>
#include <vector>
#include <cstring>
#include "/root/reference/include/wt/math/util.hpp"
#include "_ref/cone_scalar_part.hpp"
#include <wt/ads/common.hpp>
#include "/root/reference/include/wt/ads/bvh8w/bvh8w_node.hpp"
#include "/root/reference/include/wt/ads/bvh8w/common.hpp"
#include "../include/wtgpu.h"
#include "/root/reference/include/wt/ads/intersection_record.hpp"
#include "/root/reference/include/wt/util/unreachable.hpp"
#include <set>
namespace wt::ads {
class ads_t {       // ads.hpp: the options (:30-35) and the four queries (:60-100) the integrators call
public:
    struct intersect_opts_t { bool detect_edges = false, accumulate_edges = false, accumulate_triangles = false; f_t z_search_range_scale = 1; };
    virtual intersection_record_t intersect(const ray_t& ray, const pqrange_t<> range) const noexcept = 0;
    virtual intersection_record_t intersect(const elliptic_cone_t& cone, const pqrange_t<> range, const intersect_opts_t& opts) const noexcept = 0;
    virtual bool shadow(const ray_t& ray, const pqrange_t<> range) const noexcept = 0;
    virtual bool shadow(const elliptic_cone_t& cone, const pqrange_t<> range) const noexcept = 0;
    virtual const tri_t& tri(tuid_t t) const noexcept = 0;
    virtual const edge_t& edge(tuid_t e) const noexcept = 0;
};
struct vectorized_tri_data_t { std::vector<f_t> ax, ay, az, bx, by, bz, cx, cy, cz, nx, ny, nz; };
class bvh8w_t : public ads_t {
public:
    std::vector<tri_t> tris; std::vector<bvh8w::node_t> nodes; std::vector<bvh8w::leaf_node_t> leaves; std::int32_t root = 0; vectorized_tri_data_t vt; std::vector<edge_t> edges;
    const tri_t& tri(tuid_t t) const noexcept override { return tris[t.uid]; }
    const edge_t& edge(tuid_t e) const noexcept override { return edges[e.uid]; }
    const bvh8w::node_t& node(idx_t i) const noexcept { return nodes[i]; }
    const bvh8w::leaf_node_t& leaf_node(idx_t i) const noexcept { return leaves[i]; }
    std::int32_t root_ptr() const noexcept { return root; }
    const vectorized_tri_data_t& vectorized_tri_data() const noexcept { return vt; }
    intersection_record_t intersect(const ray_t& ray, const pqrange_t<> range) const noexcept override;
    intersection_record_t intersect(const elliptic_cone_t& cone, const pqrange_t<> range, const intersect_opts_t& opts) const noexcept override;
    bool shadow(const ray_t& ray, const pqrange_t<> range) const noexcept override;
    bool shadow(const elliptic_cone_t& cone, const pqrange_t<> range) const noexcept override;
};
}
namespace wt::ads_stats {       // ads_stats.hpp:148-201 without the counters
static constexpr auto additional_ads_counters = false;
inline void on_ray_aabb_8w_test() noexcept {}
template <typename... Ts> inline auto intersect_ray_tri_8w(Ts&&... ts) noexcept { return intersect::intersect_ray_tri(std::forward<Ts>(ts)...); }
template <typename... Ts> inline auto test_ray_tri_8w(Ts&&... ts) noexcept { return intersect::test_ray_tri(std::forward<Ts>(ts)...); }
template <typename... Ts> inline std::optional<intersect::intersect_cone_tri_ret_t> intersect_cone_tri(Ts&&... ts) noexcept { return intersect::intersect_cone_tri(std::forward<Ts>(ts)...); }
template <typename... Ts> inline bool test_cone_tri(Ts&&... ts) noexcept { return intersect::test_cone_tri(std::forward<Ts>(ts)...); }
}
#include "_ref/traversal_common_part.hpp"
using namespace wt;
using namespace wt::ads;
#include "_ref/bvh8w_traverse_part.hpp"

// the four bvh8w_t members (bvh8w.cpp:320-375, :556-603) without their statistics / timing lines
intersection_record_t bvh8w_t::intersect(const elliptic_cone_t& cone, const pqrange_t<> traversal_range, const intersect_opts_t& opts) const noexcept {
    auto work = intersection_record_vec_work_t{ traversal_range, opts.z_search_range_scale };
    int internal_nodes = 0, leaf_nodes = 0, subtrees = 0;
    ::traverse<false>(this, cone, opts, work, internal_nodes, leaf_nodes, subtrees);
    return cone_work_to_intersection_record(*this, work, cone, opts);
}
bool bvh8w_t::shadow(const elliptic_cone_t& cone, const pqrange_t<> traversal_range) const noexcept {
    if (traversal_range.empty()) return false;
    auto work = intersection_record_vec_work_t{ traversal_range };
    int internal_nodes = 0, leaf_nodes = 0, subtrees = 0;
    ::traverse<true>(this, cone, { .detect_edges = false }, work, internal_nodes, leaf_nodes, subtrees);
    return m::isfinite(work.intr_dist);
}
intersection_record_t bvh8w_t::intersect(const ray_t& ray, const pqrange_t<> traversal_range) const noexcept {
    intersection_record_ray_work_t work{ traversal_range }; int nodes = 0;
    ::traverse<false>(this, ray, work, nodes);
    return ray_work_to_intersection_record(*this, work, traversal_range);
}
bool bvh8w_t::shadow(const ray_t& ray, const pqrange_t<> traversal_range) const noexcept {
    if (traversal_range.empty()) return false;
    intersection_record_ray_work_t work{ traversal_range }; int nodes = 0;
    return ::traverse<true>(this, ray, work, nodes);
}
static bvh8w_t g_tree;
namespace wt::beam { struct beam_generic_t { static inline constexpr f_t major_axis_to_z_scale() noexcept { return 2; } }; }     // beam/beam_generic.hpp:50
namespace wt { template <typename T> concept Wavelength = std::is_floating_point_v<T>; }
#include "_ref/integrator_traversal_part.hpp"
// the primary-triangle pick of plt_path::random_walk (plt_path_detail.hpp:244-276) over the reference's own cone_intersection_tolerance.hpp
#include <optional>
#include "/root/reference/include/wt/math/intersect/cone_intersection_tolerance.hpp"
namespace ref_plt_path { using namespace wt; using namespace wt::ads;
#include "_ref/plt_path_closest_part.hpp"
}
// ... and plt_bdpt's (plt_bdpt_detail.hpp:351-419): the same pick, then -- when no triangle lies under the point -- the beam's Gaussian integrated over the
// front- or back-facing triangles of the list, clipped to the interaction depth and projected onto the cross-section at its centre, summed in list order;
// over the reference's own clip.hpp, gaussian_wavefront.hpp, elliptic_cone_t::project_local and src/math/gaussian2d.cpp (oracle/ref_gaussian2d_lanes.cpp)
#include "/root/reference/include/wt/math/intersect/clip.hpp"
#include "/root/reference/include/wt/beam/gaussian_wavefront.hpp"
namespace ref_plt_bdpt { using namespace wt; using namespace wt::ads;
#include "_ref/plt_bdpt_closest_part.hpp"
}
// the geometry of an edge record -- wedge normals pointing outwards, in-face tangents, opening angle, the 160-degree cut -- as the reference derives it from
// the two triangles sharing the edge: edge_for, include/wt/ads/edge_classification.hpp:31-86.  Pins the edge table the HOST LAYER builds (libwt_host.so).
#include <atomic>
namespace wt::ads::construction {
#include "_ref/edge_for_part.hpp"
}
// the UTD aperture of plt_path's diffusive vertices and its evaluation: the constructor and f() of wt::free_space_diffraction_t, src/interaction/fsd/
// free_space_diffraction.cpp:17-81 and :199-240 (front-face choice per wedge, edges clamped to the interaction region's ellipsoid, Fermat point per
// edge, wedge-side rejection, UTD coefficients), over the reference's own utd.hpp / fsd/common.hpp and intersect_edge_ellipsoid.  The class is declared
// here with the members those two functions touch (free_space_diffraction.hpp:28-75 pulls in sampler/density.hpp, written directly over mp-units); sample()
// and pdf() need its angle-density types and are not part of the pin.  libcerf's cerfc is supplied by the test (ref_traverse_set_cerfc), as for libref_utd.so.
#include "/root/reference/include/wt/interaction/fsd/utd.hpp"
extern "C" { ref_cerfc_fn ref_cerfc_hook = nullptr; void ref_traverse_set_cerfc(ref_cerfc_fn f) { ref_cerfc_hook = f; } }
namespace wt {
class free_space_diffraction_t {
public:
    utd::fsd_aperture_t aperture; pqvec3_t interaction_wp;
    struct diffracting_edge_t { utd::UTD_ret_t utd; ads::tuid_t edge_idx; pqvec3_t p; dir3_t wi, wo; length_t ri, ro; };
    using eval_ret_t = std::vector<diffracting_edge_t>;
    free_space_diffraction_t(const ads::ads_t* ads, const pqvec3_t& interaction_wp, const frame_t& interaction_region_frame, const pqvec3_t& interaction_region_size,
                             const dir3_t& wi, wavenumber_t k, const ads::intersection_record_t::edges_container_t& edges) noexcept;
    eval_ret_t f(const pqvec3_t& src, const pqvec3_t& dst) const noexcept;
};
}
namespace wt {
#include "_ref/utd_fsd_part.hpp"
}
// the Fraunhofer aperture of plt_bdpt's diffusive vertices: the constructor of fraunhofer::free_space_diffraction_t, src/interaction/fsd/fraunhofer/
// free_space_diffraction.cpp:18-129 (silhouette edges of the cone query's edge set, clamped to the beam's 3-sigma ellipse, cut into segments of a third of
// its radius, each weighted by Pj; the 0-th order lobe from eight samples of the aperture's spectrum), over the reference's own fsd.hpp, fsd_sampler.hpp,
// gaussian_wavefront.hpp, intersect_edge_ellipse and is_point_in_ellipsoid.  The class is declared here with the members the constructor touches
// (free_space_diffraction.hpp:30-60 pulls in sampler/density.hpp, written directly over mp-units).
#include "/root/reference/include/wt/interaction/fsd/fraunhofer/fsd.hpp"
#include "/root/reference/include/wt/interaction/fsd/fraunhofer/fsd_sampler.hpp"
namespace wt::fraunhofer {
class free_space_diffraction_t {
public:
    static constexpr auto fsd_unit = f_t(1) * u::mm;
    fsd::fsd_aperture_t aperture; wavenumber_t k; frame_t frame; const fsd_sampler::fsd_sampler_t* fsd_sampler;
    free_space_diffraction_t(const ads::ads_t* ads, const fsd_sampler::fsd_sampler_t* fsd_sampler, const frame_t& frame, wavenumber_t k, f_t totalPower,
                             const elliptic_cone_t& beam, const ads::intersection_record_t::edges_container_t& edges, const beam::gaussian_wavefront_t& wave_function) noexcept;
};
}
namespace wt::fraunhofer {
#include "_ref/ffsd_ctor_part.hpp"
}
// self-intersection offsets: compute_intersection_triangle_fp_errors and intersection_edge_t::offseted_ray_origin (src/interaction/intersection.cpp:149-170, :187-211)
#include "_ref/intersection_offset_part.hpp"


extern "C" {
// the host layer's BVH (include/wtgpu.h: 8-wide nodes, leaves, packed triangles) into the containers the loops read
void ref_traverse_load(const wtgpu_scene_desc* d) {
    bvh8w_t& t = g_tree;
    t.tris.resize(d->n_tris); t.nodes.resize(d->n_nodes); t.leaves.resize(d->n_leaves); t.root = d->root_ptr;
    auto& v = t.vt;
    for (auto* a : { &v.ax, &v.ay, &v.az, &v.bx, &v.by, &v.bz, &v.cx, &v.cy, &v.cz, &v.nx, &v.ny, &v.nz }) a->assign(d->n_tris + 8, 0.f);
    for (uint32_t i = 0; i < d->n_tris; ++i) {
        const wtgpu_tri& s = d->tris[i];
        t.tris[i].a = pqvec3_t{ s.ax, s.ay, s.az }; t.tris[i].b = pqvec3_t{ s.bx, s.by, s.bz }; t.tris[i].c = pqvec3_t{ s.cx, s.cy, s.cz }; t.tris[i].n = dir3_t{ s.nx, s.ny, s.nz };
        if (d->tri_meta) { const auto& mt = d->tri_meta[i]; t.tris[i].edge_ab = tuid_t{ mt.edge_ab }; t.tris[i].edge_bc = tuid_t{ mt.edge_bc }; t.tris[i].edge_ca = tuid_t{ mt.edge_ca }; }
        v.ax[i] = s.ax; v.ay[i] = s.ay; v.az[i] = s.az; v.bx[i] = s.bx; v.by[i] = s.by; v.bz[i] = s.bz; v.cx[i] = s.cx; v.cy[i] = s.cy; v.cz[i] = s.cz; v.nx[i] = s.nx; v.ny[i] = s.ny; v.nz[i] = s.nz;
    }
    for (uint32_t i = 0; i < d->n_nodes; ++i) {
        const wtgpu_node& s = d->nodes[i]; auto& n = t.nodes[i];
        for (int l = 0; l < 8; ++l) {
            n.min.c[0].v[l] = s.minx[l]; n.min.c[1].v[l] = s.miny[l]; n.min.c[2].v[l] = s.minz[l];
            n.max.c[0].v[l] = s.maxx[l]; n.max.c[1].v[l] = s.maxy[l]; n.max.c[2].v[l] = s.maxz[l];
            n.child_ptrs[l] = s.child[l];
        }
        n.tris_start = s.tris_start; n.tris_count = s.tris_count;
    }
    auto& g_edges = t.edges; g_edges.assign(d->n_edges, edge_t{});
    for (uint32_t i = 0; i < d->n_edges; ++i) {
        const wtgpu_edge& s = d->edges[i]; edge_t& e = g_edges[i];
        e.t1 = dir3_t{ s.t1[0], s.t1[1], s.t1[2] }; e.t2 = dir3_t{ s.t2[0], s.t2[1], s.t2[2] };
        e.a = pqvec3_t{ s.a[0], s.a[1], s.a[2] }; e.b = pqvec3_t{ s.b[0], s.b[1], s.b[2] }; e.e = dir3_t{ s.e[0], s.e[1], s.e[2] };
        e.n1 = dir3_t{ s.n1[0], s.n1[1], s.n1[2] }; e.n2 = dir3_t{ s.n2[0], s.n2[1], s.n2[2] }; e.alpha = s.alpha;
        e.tri1 = &t.tris[s.tri1]; e.tri2 = s.tri2 != WTGPU_INVALID_IDX ? &t.tris[s.tri2] : nullptr;
    }
    for (uint32_t i = 0; i < d->n_leaves; ++i) { t.leaves[i].tris_ptr = d->leaves[i].tris_ptr; t.leaves[i].count = d->leaves[i].count; }
}
// per query in: o[3] d[3] x[3] tan_alpha eccentricity x0 tmin tmax z_scale; out: the accepted triangles in traversal order (first `cap`), their number,
// the closest distance, the face flag  (bvh8w_t::intersect(cone), bvh8w.cpp:320-331, before cone_work_to_intersection_record)
void ref_traverse_cones(uint32_t n, const float* q, uint32_t cap, uint32_t* counts, uint32_t* tuids, float* dist, uint32_t* front) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 15 * i;
        const elliptic_cone_t cone{ ray_t{ pqvec3_t{ c[0], c[1], c[2] }, dir3_t{ c[3], c[4], c[5] } }, dir3_t{ c[6], c[7], c[8] }, c[9], c[10], length_t(c[11]) };
        ads_t::intersect_opts_t opts; opts.z_search_range_scale = c[14];
        auto work = intersection_record_vec_work_t{ pqrange_t<>{ c[12], c[13] }, opts.z_search_range_scale };
        int internal_nodes = 0, leaf_nodes = 0, subtrees = 0;
        ::traverse<false>(&g_tree, cone, opts, work, internal_nodes, leaf_nodes, subtrees);
        counts[i] = (uint32_t)work.triangles.size(); dist[i] = work.intr_dist; front[i] = work.front_face ? 1u : 0u;
        for (uint32_t k = 0; k < cap; ++k) tuids[(size_t)i * cap + k] = k < work.triangles.size() ? (uint32_t)work.triangles[k].tuid.uid : 0xffffffffu;
    }
}
// integrator::traverse (traversal.hpp:94-172).  per query in: o[3] d[3] x[3] tan_alpha eccentricity x0 lambda distance force_ray_tracing detect_edges = 16;
// out per query: 12 floats (empty ballistic origin[3] distance depth front_face has_rt bary[2] pad) and the triangle / edge lists (counts + first `cap`)
void ref_integrator_traverse(uint32_t n, const float* q, uint32_t cap, float* out, uint32_t* ntris, uint32_t* tris, uint32_t* nedges, uint32_t* edges) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 16 * i; float* o = out + 12 * i;
        const elliptic_cone_t cone{ ray_t{ pqvec3_t{ c[0], c[1], c[2] }, dir3_t{ c[3], c[4], c[5] } }, dir3_t{ c[6], c[7], c[8] }, c[9], c[10], length_t(c[11]) };
        const auto r = integrator::traverse(g_tree, cone, length_t(c[12]), length_t(c[13]), integrator::traversal_opts_t{ .force_ray_tracing = c[14] != 0, .detect_edges = c[15] != 0 });
        const bool empty = r.record.empty();
        o[0] = empty; o[1] = r.ballistic; o[2] = r.origin.x; o[3] = r.origin.y; o[4] = r.origin.z; o[5] = empty ? 0.f : (float)r.record.distance(); o[6] = r.intersection_region_depth;
        o[7] = !empty && r.record.is_front_face(); o[8] = r.record.has_raytracing_intersection_record();
        o[9] = o[8] ? r.record.get_raytracing_intersection_record().bary.uv.x : 0.f; o[10] = o[8] ? r.record.get_raytracing_intersection_record().bary.uv.y : 0.f; o[11] = 0;
        uint32_t k = 0; for (const auto& t : r.record.triangles()) { if (k < cap) tris[(size_t)i * cap + k] = t.uid; ++k; }
        ntris[i] = k; for (; k < cap; ++k) tris[(size_t)i * cap + k] = 0xffffffffu;
        k = 0; for (const auto& e : r.record.edges()) { if (k < cap) edges[(size_t)i * cap + k] = e.uid; ++k; }
        nedges[i] = k; for (; k < cap; ++k) edges[(size_t)i * cap + k] = 0xffffffffu;
    }
}
// plt_bdpt's find_closest_triangle.  per query in: origin[3] dir[3] zmin zmax first_tuid count | beam frame t[3] b[3] (n = dir) | envelope o[3] x[3] tan_alpha
// eccentricity x0 (d = dir) | wavefront sigma x y | integrate_front_facing = 28; out: tuid (or ~0), dist bary[2] integrated_radiant_flux
void ref_bd_find_closest_triangle(uint32_t n, const float* q, float* out, uint32_t* tuid) {
    std::vector<tuid_t> list;
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 28 * i; float* o = out + 4 * i;
        list.clear(); for (uint32_t k = 0; k < (uint32_t)c[9]; ++k) list.push_back(tuid_t{ (uint32_t)c[8] + k });
        const intersection_record_t::triangles_accessor_t acc{ .s = list.data(), .e = list.data() + list.size() };
        const dir3_t dir{ c[3], c[4], c[5] };
        const frame_t bf{ dir3_t{ c[10], c[11], c[12] }, dir3_t{ c[13], c[14], c[15] }, dir };
        const elliptic_cone_t env{ ray_t{ pqvec3_t{ c[16], c[17], c[18] }, dir }, dir3_t{ c[19], c[20], c[21] }, c[22], c[23], length_t(c[24]) };
        const beam::gaussian_wavefront_t wf{ gaussian2d_t{ vec2_t{ c[25], c[26] } } };
        const auto id = ref_plt_bdpt::find_closest_triangle(acc, g_tree, pqrange_t<>{ c[6], c[7] }, pqvec3_t{ c[0], c[1], c[2] }, dir, bf, env, wf, true, c[27] != 0);
        tuid[i] = id.primary ? (uint32_t)(id.primary - g_tree.tris.data()) : 0xffffffffu;
        o[0] = id.primary ? (float)id.primary_intersection_record.dist : 0.f; o[1] = id.primary ? id.primary_intersection_record.bary.uv.x : 0.f; o[2] = id.primary ? id.primary_intersection_record.bary.uv.y : 0.f;
        o[3] = id.integrated_radiant_flux;
    }
}
// per query in: beam o[3] d[3] x[3] tan_alpha eccentricity x0 | frame t[3] b[3] n[3] | k [1/mm] total_power | wavefront sigma x y | first edge, count = 27
// out: counts[i] = aperture edges; summary[4] = recp_I psi02 P0 P0_pdf; the first `cap` edges x 9 (e[2] v[2] a_b re im iab_2 re im pdf)
void ref_ffsd_aperture(uint32_t n, const float* q, uint32_t cap, uint32_t* counts, float* summary, float* edges) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 27 * i;
        const elliptic_cone_t beam{ ray_t{ pqvec3_t{ c[0], c[1], c[2] }, dir3_t{ c[3], c[4], c[5] } }, dir3_t{ c[6], c[7], c[8] }, c[9], c[10], length_t(c[11]) };
        const frame_t fr{ dir3_t{ c[12], c[13], c[14] }, dir3_t{ c[15], c[16], c[17] }, dir3_t{ c[18], c[19], c[20] } };
        const beam::gaussian_wavefront_t wf{ gaussian2d_t{ vec2_t{ c[23], c[24] } } };
        std::set<tuid_t> es; for (uint32_t k = 0; k < (uint32_t)c[26]; ++k) es.insert(tuid_t{ (uint32_t)c[25] + k });
        const fraunhofer::free_space_diffraction_t f(&g_tree, nullptr, fr, wavenumber_t{ c[21] }, c[22], beam, es, wf);
        const auto& ap = f.aperture;
        counts[i] = (uint32_t)ap.edges.size();
        summary[4 * i] = ap.recp_I; summary[4 * i + 1] = ap.psi02; summary[4 * i + 2] = ap.P0; summary[4 * i + 3] = ap.P0_pdf;
        for (uint32_t k = 0; k < cap; ++k) {
            float* o = edges + ((size_t)i * cap + k) * 9;
            if (k < ap.edges.size()) { const auto& e = ap.edges[k]; o[0] = e.e.x; o[1] = e.e.y; o[2] = e.v.x; o[3] = e.v.y; o[4] = e.a_b.real(); o[5] = e.a_b.imag(); o[6] = e.iab_2.real(); o[7] = e.iab_2.imag(); o[8] = ap.edge_pdfs[k]; }
            else for (int j = 0; j < 9; ++j) o[j] = 0.f;
        }
    }
}
// per query in: interaction_wp[3] | region frame t[3] b[3] n[3] | region size[3] | wi[3] | k [1/mm] | first edge, count | src[3] dst[3] = 28
// out: nap[i] aperture wedges, the first `cap` x 15 (v[3] l nff[3] tff[3] nbf[3] alpha idx); nf[i] diffracting edges of f(src, dst), the first `cap` x 10
// (idx p[3] ri ro Ds re im Dh re im)
void ref_utd_fsd(uint32_t n, const float* q, uint32_t cap, uint32_t* nap, float* ap, uint32_t* nf, float* fo) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 28 * i;
        const frame_t fr{ dir3_t{ c[3], c[4], c[5] }, dir3_t{ c[6], c[7], c[8] }, dir3_t{ c[9], c[10], c[11] } };
        std::set<tuid_t> es; for (uint32_t k = 0; k < (uint32_t)c[21]; ++k) es.insert(tuid_t{ (uint32_t)c[20] + k });
        const wt::free_space_diffraction_t f(&g_tree, pqvec3_t{ c[0], c[1], c[2] }, fr, pqvec3_t{ c[12], c[13], c[14] }, dir3_t{ c[15], c[16], c[17] }, wavenumber_t{ c[18] }, es);
        nap[i] = (uint32_t)f.aperture.edges.size();
        for (uint32_t k = 0; k < cap; ++k) {
            float* o = ap + ((size_t)i * cap + k) * 15; for (int j = 0; j < 15; ++j) o[j] = 0.f;
            if (k >= f.aperture.edges.size()) continue;
            const auto& e = f.aperture.edges[k];
            o[0] = e.v.x; o[1] = e.v.y; o[2] = e.v.z; o[3] = (float)e.l; o[4] = e.nff.x; o[5] = e.nff.y; o[6] = e.nff.z; o[7] = e.tff.x; o[8] = e.tff.y; o[9] = e.tff.z;
            o[10] = e.nbf.x; o[11] = e.nbf.y; o[12] = e.nbf.z; o[13] = (float)e.alpha; o[14] = (float)e.ads_edge_idx.uid;
        }
        const auto r = f.f(pqvec3_t{ c[22], c[23], c[24] }, pqvec3_t{ c[25], c[26], c[27] });
        nf[i] = (uint32_t)r.size();
        for (uint32_t k = 0; k < cap; ++k) {
            float* o = fo + ((size_t)i * cap + k) * 10; for (int j = 0; j < 10; ++j) o[j] = 0.f;
            if (k >= r.size()) continue;
            const auto& d = r[k];
            o[0] = (float)d.edge_idx.uid; o[1] = d.p.x; o[2] = d.p.y; o[3] = d.p.z; o[4] = (float)d.ri; o[5] = (float)d.ro;
            o[6] = d.utd.Ds.real(); o[7] = d.utd.Ds.imag(); o[8] = d.utd.Dh.real(); o[9] = d.utd.Dh.imag();
        }
    }
}
// per item in: tri1 a[3] b[3] c[3] n[3] | has tri2 | tri2 a[3] b[3] c[3] n[3] | edge a[3] b[3] | c1[3] c2[3] = 37
// out: found, e[3] n1[3] t1[3] n2[3] t2[3] alpha, inconsistent-normals flag = 18
void ref_edge_for(uint32_t n, const float* q, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 37 * i; float* o = out + 18 * i;
        auto load = [](const float* t) { tri_t r{}; r.a = pqvec3_t{ t[0], t[1], t[2] }; r.b = pqvec3_t{ t[3], t[4], t[5] }; r.c = pqvec3_t{ t[6], t[7], t[8] }; r.n = dir3_t{ t[9], t[10], t[11] }; return r; };
        const tri_t t1 = load(c), t2 = load(c + 13);
        const bool has2 = c[12] != 0;
        const pqvec3_t a{ c[25], c[26], c[27] }, b{ c[28], c[29], c[30] }, c1{ c[31], c[32], c[33] }, c2{ c[34], c[35], c[36] };
        std::atomic<bool> inconsistent{ false };
        const auto e = construction::edge_for(&t1, has2 ? &t2 : nullptr, tuid_t{ 0 }, tuid_t{ 1 }, a, b, c1, has2 ? &c2 : nullptr, inconsistent);
        for (int j = 0; j < 18; ++j) o[j] = 0.f;
        o[17] = inconsistent ? 1.f : 0.f;
        if (!e) continue;
        o[0] = 1.f; o[1] = e->e.x; o[2] = e->e.y; o[3] = e->e.z; o[4] = e->n1.x; o[5] = e->n1.y; o[6] = e->n1.z; o[7] = e->t1.x; o[8] = e->t1.y; o[9] = e->t1.z;
        o[10] = e->n2.x; o[11] = e->n2.y; o[12] = e->n2.z; o[13] = e->t2.x; o[14] = e->t2.y; o[15] = e->t2.z; o[16] = (float)e->alpha;
    }
}
// per query in: edge index, ray o[3] d[3]; out: the offset origin (intersection.cpp:187-211), then the fp error bound of the edge's first triangle (:149-170)
void ref_edge_offsets(uint32_t n, const float* q, float* out) {
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 7 * i; float* o = out + 6 * i;
        const edge_t& e = g_tree.edges[(uint32_t)c[0]];
        const ray_t ray{ pqvec3_t{ c[1], c[2], c[3] }, dir3_t{ c[4], c[5], c[6] } };
        const auto p = intersection_edge_t{ &e, ray.o }.offseted_ray_origin(ray);
        const auto err = compute_intersection_triangle_fp_errors(e.tri1->a, e.tri1->b, e.tri1->c, ray.o);
        o[0] = p.x; o[1] = p.y; o[2] = p.z; o[3] = err.x; o[4] = err.y; o[5] = err.z;
    }
}
// find_closest_triangle (plt_path_detail.hpp:253-276).  per query in: origin[3] dir[3] zmin zmax first_tuid count; out: tuid (or ~0) dist bary[2]
void ref_find_closest_triangle(uint32_t n, const float* q, float* out, uint32_t* tuid) {
    std::vector<tuid_t> list;
    for (uint32_t i = 0; i < n; ++i) {
        const float* c = q + 10 * i;
        list.clear(); for (uint32_t k = 0; k < (uint32_t)c[9]; ++k) list.push_back(tuid_t{ (uint32_t)c[8] + k });
        const intersection_record_t::triangles_accessor_t acc{ .s = list.data(), .e = list.data() + list.size() };
        const auto id = ref_plt_path::find_closest_triangle(acc, g_tree, pqrange_t<>{ c[6], c[7] }, pqvec3_t{ c[0], c[1], c[2] }, dir3_t{ c[3], c[4], c[5] });
        tuid[i] = id.primary ? (uint32_t)(id.primary - g_tree.tris.data()) : 0xffffffffu;
        out[3 * i] = id.primary ? (float)id.primary_intersection_record.dist : 0.f;
        out[3 * i + 1] = id.primary ? id.primary_intersection_record.bary.uv.x : 0.f; out[3 * i + 2] = id.primary ? id.primary_intersection_record.bary.uv.y : 0.f;
    }
}
// bvh8w_t::intersect(ray) / shadow(ray), bvh8w.cpp:556-603, before ray_work_to_intersection_record
void ref_traverse_rays(uint32_t n, const wtgpu_ray_query* q, wtgpu_ray_hit* out, uint32_t* shadow) {
    for (uint32_t i = 0; i < n; ++i) {
        const ray_t ray{ pqvec3_t{ q[i].o[0], q[i].o[1], q[i].o[2] }, dir3_t{ q[i].d[0], q[i].d[1], q[i].d[2] } };
        const pqrange_t<> range{ q[i].tmin, q[i].tmax };
        { intersection_record_ray_work_t work{ range }; int nodes = 0;
          ::traverse<false>(&g_tree, ray, work, nodes);
          const bool hit = m::isfinite(work.triangle.dist) && !(work.triangle.dist > range.max);
          out[i].tuid = hit ? (uint32_t)work.triangle.tuid.uid : 0xffffffffu; out[i].dist = hit ? (float)work.triangle.dist : std::numeric_limits<float>::infinity();
          out[i].bary[0] = hit ? work.intersection.bary.uv.x : -1.f; out[i].bary[1] = hit ? work.intersection.bary.uv.y : -1.f; out[i].front_face = hit && work.triangle.front_face; }
        { intersection_record_ray_work_t work{ range }; int nodes = 0;
          shadow[i] = ::traverse<true>(&g_tree, ray, work, nodes) ? 1u : 0u; }
    }
}
}
