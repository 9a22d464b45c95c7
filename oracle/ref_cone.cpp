// ORACLE -- TEST INFRASTRUCTURE ONLY.
// The cone tests of the reference's include/wt/math/intersect/cone.hpp -- intersect_cone_edge (cone.hpp:38-128), intersect_cone_plane (:170-258) and
// the cone-triangle drivers test_cone_tri (:479-539) and intersect_cone_tri (:550-626), i.e. what every cone query of the hot path runs per triangle --
// together with the reference's own include/wt/math/shapes/elliptic_cone.hpp, shapes/ray.hpp, intersect/ray.hpp (scalar entry points), intersect/misc.hpp,
// math/util.hpp and math/frame.hpp, compiled from where they lie -> oracle/_ref/libref_cone.so.  tests/test_oracle_kats.py compares them bit for bit with
// ot_math.h.
// The middle of cone.hpp (:272-475, cone-AABB on 8-wide AVX vectors) needs the reference's SIMD layer, which does not compile against the shim; so the
// Makefile's `ref` target writes lines 1-271 and 476-end of the header as they are to the git-ignored oracle/_ref/cone_scalar_part.hpp at build time
// and this TU includes that.  The cut is deleted again once the library is linked: no reference text is committed or left in the tree.  The 4-wide vector the triangle drivers stage the vertices in is the shim's array of lanes
// (WT_SHIM_WIDE_LANES, ref_shims/wt/math/simd/wide_vector.hpp): one IEEE operation per AVX instruction.
#define WT_SHIM_DISTINCT_PQ
#define WT_SHIM_WIDE_LANES
#include <wt/util/assert.hpp>
#include "/root/reference/include/wt/math/util.hpp"
#include "_ref/cone_scalar_part.hpp"
using namespace wt;
namespace {
// per cone: o[3] d[3] x[3] tan_alpha eccentricity x0  (the public constructor, elliptic_cone.hpp:64-78)
inline elliptic_cone_t make_cone(const float* c) {
    return elliptic_cone_t{ ray_t{ pqvec3_t{ c[0], c[1], c[2] }, dir3_t{ c[3], c[4], c[5] } }, dir3_t{ c[6], c[7], c[8] }, c[9], c[10], length_t(c[11]) };
}
template <bool in_local>
inline void edge_one(const float* a, float* o) {
    const auto cone = make_cone(a);
    const auto r = intersect::intersect_cone_edge<in_local>(cone, pqvec3_t{ a[12], a[13], a[14] }, pqvec3_t{ a[15], a[16], a[17] }, pqrange_t<>{ a[18], a[19] });
    for (int k = 0; k < 10; ++k) o[k] = 0.f;
    if (!r) return;
    o[0] = 1.f; o[1] = r->p0.x; o[2] = r->p0.y; o[3] = r->p0.z;
    if (r->pts == 2) { o[4] = r->p1.x; o[5] = r->p1.y; o[6] = r->p1.z; }
    o[7] = r->range.min; o[8] = r->range.max; o[9] = (float)r->pts;
}
template <bool in_local>
inline void plane_one(const float* a, float* o) {
    const auto cone = make_cone(a);
    const auto r = intersect::intersect_cone_plane<in_local>(cone, dir3_t{ a[12], a[13], a[14] }, length_t(a[15]), pqrange_t<>{ a[16], a[17] });
    for (int k = 0; k < 9; ++k) o[k] = 0.f;
    if (r.range.empty()) return;
    o[0] = 1.f; o[1] = r.range.min; o[2] = r.range.max;
    o[3] = r.near.x; o[4] = r.near.y; o[5] = r.near.z; o[6] = r.far.x; o[7] = r.far.y; o[8] = r.far.z;
}
}
// ---- the per-node test of the BVH cone traversal, src/ads/bvh8w.cpp:186-230 (cone_cluster_intersect) with its inputs (:107-121) and the child
// stack's insertion sort (:44-57): the Makefile writes those line ranges of the .cpp, as they are, to oracle/_ref/bvh8w_cone_cluster_part.hpp.
#include <bitset>
#include <vector>
namespace wt::ads { struct bvh8w_t; namespace bvh8w { struct bvh8w_aabbs_t { pqvec3_w_t<8> min, max; }; } }      // include/wt/ads/bvh8w/common.hpp:19-21
using namespace wt::ads;
#include "_ref/bvh8w_cone_cluster_part.hpp"
extern "C" {
// per item in: cone[12] p0[3] p1[3] range[2]; out: found p0[3] p1[3] (zero unless pts == 2) range[2] pts
void ref_cone_edge(unsigned n, int in_local, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) { if (in_local) edge_one<true>(in + 20 * i, out + 10 * i); else edge_one<false>(in + 20 * i, out + 10 * i); }
}
// per item in: cone[12] n[3] d range[2]; out: found (range not empty) range[2] near[3] far[3]
void ref_cone_plane(unsigned n, int in_local, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) { if (in_local) plane_one<true>(in + 18 * i, out + 9 * i); else plane_one<false>(in + 18 * i, out + 9 * i); }
}
// per item in: cone[12] a[3] b[3] c[3] n[3] range[2]; out: found dist p[3] (intersect_cone_tri, cone.hpp:550-626), test_cone_tri (cone.hpp:479-539)
void ref_cone_tri(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 26 * i; float* o = out + 6 * i;
        const auto cone = make_cone(a);
        const pqvec3_t A{ a[12], a[13], a[14] }, B{ a[15], a[16], a[17] }, Cc{ a[18], a[19], a[20] };
        const pqrange_t<> range{ a[24], a[25] };
        const auto h = intersect::intersect_cone_tri(cone, A, B, Cc, dir3_t{ a[21], a[22], a[23] }, range);
        o[0] = h ? 1.f : 0.f; o[1] = h ? (float)h->dist : 0.f; o[2] = h ? h->p.x : 0.f; o[3] = h ? h->p.y : 0.f; o[4] = h ? h->p.z : 0.f;
        o[5] = intersect::test_cone_tri(cone, A, B, Cc, range) ? 1.f : 0.f;
    }
}
// per item in: ro[3] rd[3] a[3] b[3] c[3] range[2] tol; out: found dist bary[2] (scalar intersect_ray_tri, ray.hpp:147-179: the ray degenerate of
// intersect_cone_tri), test_ray_tri with tol = 0 and with tol (ray.hpp:56-76), then intersect_line_plane(a, b, c, rd) found / d (ray.hpp:30-49)
void ref_ray_tri(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 18 * i; float* o = out + 8 * i;
        const ray_t r{ pqvec3_t{ a[0], a[1], a[2] }, dir3_t{ a[3], a[4], a[5] } };
        const pqvec3_t A{ a[6], a[7], a[8] }, B{ a[9], a[10], a[11] }, Cc{ a[12], a[13], a[14] };
        const pqrange_t<> range{ a[15], a[16] };
        const auto h = intersect::intersect_ray_tri(r, A, B, Cc, range);
        o[0] = h ? 1.f : 0.f; o[1] = h ? (float)h->dist : 0.f; o[2] = h ? h->bary.uv.x : 0.f; o[3] = h ? h->bary.uv.y : 0.f;
        o[4] = intersect::test_ray_tri(r, A, B, Cc, range) ? 1.f : 0.f;
        o[5] = intersect::test_ray_tri(r, A, B, Cc, range, a[17]) ? 1.f : 0.f;
        const auto lp = intersect::intersect_line_plane(A, B, Cc, r.d);
        o[6] = lp ? 1.f : 0.f; o[7] = lp ? *lp : 0.f;
    }
}
// the 8-wide entry points of intersect/ray.hpp the BVH ray traversal evaluates (8 items per call, one per lane; n a multiple of 8).
// per item in: ro[3] rd[3] a[3] b[3] c[3] range[2]; out: result (distance or -inf) baryx baryy (intersect_ray_tri<8>, ray.hpp:192-236), test_ray_tri<8> (:93-128)
void ref_ray_tri_w8(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i + 8 <= n; i += 8) {
        const float* a0 = in + 17 * i;
        const pqrange_t<> range{ a0[15], a0[16] };          // (the range of the first lane: the caller gives all 8 the same)
        pqvec3_w_t<8> ro, A, B, Cc; vec3_w_t<8> rd;
        for (int l = 0; l < 8; ++l) for (int k = 0; k < 3; ++k) {
            const float* a = a0 + 17 * l;
            ro.c[k].v[l] = a[k]; rd.c[k].v[l] = a[3 + k]; A.c[k].v[l] = a[6 + k]; B.c[k].v[l] = a[9 + k]; Cc.c[k].v[l] = a[12 + k];
        }
        const auto r = intersect::intersect_ray_tri<8>(ro, rd, A, B, Cc, range);
        const auto t = intersect::test_ray_tri<8>(ro, rd, A, B, Cc, range);
        for (int l = 0; l < 8; ++l) { float* o = out + 4 * (i + l); o[0] = r.result.v[l]; o[1] = r.baryx.v[l]; o[2] = r.baryy.v[l]; o[3] = t.v[l] ? 1.f : 0.f; }
    }
}
// per item in: ro[3] rinvd[3] aabb_min[3] aabb_max[3] range[2]; out: mask min max (intersect_ray_aabb_fast<8>, ray.hpp:331-351)
void ref_ray_aabb_fast_w8(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i + 8 <= n; i += 8) {
        const float* a0 = in + 14 * i;
        const pqrange_t<> range{ a0[12], a0[13] };
        pqvec3_w_t<8> ro, mn, mx; vec3_w_t<8> inv;
        for (int l = 0; l < 8; ++l) for (int k = 0; k < 3; ++k) {
            const float* a = a0 + 14 * l;
            ro.c[k].v[l] = a[k]; inv.c[k].v[l] = a[3 + k]; mn.c[k].v[l] = a[6 + k]; mx.c[k].v[l] = a[9 + k];
        }
        const auto r = intersect::intersect_ray_aabb_fast<8>(ro, inv, mn, mx, range);
        for (int l = 0; l < 8; ++l) { float* o = out + 3 * (i + l); o[0] = r.mask.v[l] ? 1.f : 0.f; o[1] = r.min.v[l]; o[2] = r.max.v[l]; }
    }
}
// 8 boxes per call against ONE cone (that of the first lane's item).  per item in: cone[12] aabb_min[3] aabb_max[3] range[2]; out: hit tmin
void ref_cone_cluster(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i + 8 <= n; i += 8) {
        const float* a0 = in + 20 * i;
        const auto cone = make_cone(a0);
        const pqrange_t<> range{ a0[18], a0[19] };
        const cone_cluster_intersect_data_t data{ cone };
        bvh8w::bvh8w_aabbs_t boxes;
        for (int l = 0; l < 8; ++l) for (int k = 0; k < 3; ++k) { boxes.min.c[k].v[l] = a0[20 * l + 12 + k]; boxes.max.c[k].v[l] = a0[20 * l + 15 + k]; }
        const auto r = cone_cluster_intersect(nullptr, range, data, boxes);
        for (int l = 0; l < 8; ++l) { out[2 * (i + l)] = r.result_mask[l] ? 1.f : 0.f; out[2 * (i + l) + 1] = r.tmins.v[l]; }
    }
}
// sorts runs of `run` (key, id) pairs in place with the reference's stack_sorter; in/out: n pairs of (min_range, ptr as float)
void ref_stack_sorter(unsigned n, unsigned run, float* io) {
    std::vector<stack_node_ptr_t> st(run);
    for (unsigned i = 0; i + run <= n; i += run) {
        for (unsigned k = 0; k < run; ++k) st[k] = { io[2 * (i + k)], (int32_t)io[2 * (i + k) + 1] };
        stack_sorter(st.data(), (int)run);
        for (unsigned k = 0; k < run; ++k) { io[2 * (i + k)] = st[k].min_range; io[2 * (i + k) + 1] = (float)st[k].ptr; }
    }
}
// the beam re-fits of src/math/elliptic_cone.cpp (compiled from where it lies, linked into this library): out for both: x[3] x0 e one_over_e tan_alpha z_apex
static void put_cone(const elliptic_cone_t& c, float* o) {
    o[0] = c.x().x; o[1] = c.x().y; o[2] = c.x().z; o[3] = (float)c.x0(); o[4] = c.get_e(); o[5] = c.get_one_over_e(); o[6] = c.get_tan_alpha(); o[7] = (float)c.get_z_apex();
}
// per item in: x[3] y[3] n[3] ro[3] rd[3] tan_alpha; out: cone[8] self_intersection_distance  (elliptic_cone.cpp:19-84)
void ref_cone_through_ellipse(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 16 * i; float* o = out + 9 * i;
        length_t sid = 0;
        const auto c = elliptic_cone_t::cone_through_ellipse(pqvec3_t{ a[0], a[1], a[2] }, pqvec3_t{ a[3], a[4], a[5] }, dir3_t{ a[6], a[7], a[8] },
                                                             ray_t{ pqvec3_t{ a[9], a[10], a[11] }, dir3_t{ a[12], a[13], a[14] } }, a[15], &sid);
        put_cone(c, o); o[8] = (float)sid;
    }
}
// per item in: axes[3] frame t[3] b[3] n[3] ro[3] rd[3] tan_alpha; out: cone[8]  (elliptic_cone.cpp:86-145)
void ref_cone_through_ellipsoid(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 19 * i;
        const frame_t F{ dir3_t{ a[3], a[4], a[5] }, dir3_t{ a[6], a[7], a[8] }, dir3_t{ a[9], a[10], a[11] } };
        put_cone(elliptic_cone_t::cone_through_ellipsoid(pqvec3_t{ a[0], a[1], a[2] }, F, ray_t{ pqvec3_t{ a[12], a[13], a[14] }, dir3_t{ a[15], a[16], a[17] } }, a[18]), out + 8 * i);
    }
}
// per item in: cone[12] z; out: axes x y, z_apex, e, one_over_e  (elliptic_cone.hpp: axes(), get_z_apex(), the eccentricity constructor)
void ref_cone_basics(unsigned n, const float* in, float* out) {
    for (unsigned i = 0; i < n; ++i) {
        const float* a = in + 13 * i; float* o = out + 5 * i;
        const auto cone = make_cone(a);
        const auto ax = cone.axes(length_t(a[12]));
        o[0] = ax.x; o[1] = ax.y; o[2] = cone.get_z_apex(); o[3] = cone.get_e(); o[4] = cone.get_one_over_e();
    }
}
}
