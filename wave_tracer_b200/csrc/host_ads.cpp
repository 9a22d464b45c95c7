// host_ads.cpp -- host-side ADS construction (product code, CPU).
//
// mesh -> triangles        follows /root/reference/src/mesh/mesh.cpp:31-150 (double-precision transform, face
//                          normal, winding flip, octahedral normal encode/decode, tangent frames)
// triangles -> binary BVH  binned SAH, C_INT=100 / C_TRAV=1 / 128 bins as configured for tinybvh in
//                          /root/reference/src/ads/bvh_constructor.cpp:17-31 (tinybvh itself is not vendored in the
//                          reference tree; any valid SAH BVH is acceptable -- SURVEY.md 8c)
// binary -> 8-wide         3 binary levels per node, /root/reference/src/ads/bvh8w_constructor.cpp:27-103,153-268
// edges                    /root/reference/include/wt/ads/edge_classification.hpp:31-238, with deterministic edge ids
//                          (vertex hashing instead of ball queries; same exact-equality adjacency rule)
#include "../../include/wthost.h"

#include <algorithm>
#include <array>
#include <cmath>
#include <cstring>
#include <deque>
#include <limits>
#include <memory>
#include <numeric>
#include <string>
#include <unordered_map>
#include <vector>

namespace {

struct d3 { double x, y, z; };
inline d3 operator-(d3 a, d3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline d3 cross(d3 a, d3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline double dot(d3 a, d3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }

struct f3 { float x, y, z; };
inline f3 operator-(f3 a, f3 b) { return { a.x - b.x, a.y - b.y, a.z - b.z }; }
inline f3 operator+(f3 a, f3 b) { return { a.x + b.x, a.y + b.y, a.z + b.z }; }
inline f3 operator*(f3 a, float s) { return { a.x * s, a.y * s, a.z * s }; }
inline f3 operator-(f3 a) { return { -a.x, -a.y, -a.z }; }
inline f3 crossf(f3 a, f3 b) { return { a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x }; }
inline float dotf(f3 a, f3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
inline float lengthf(f3 a) { return std::sqrt(dotf(a, a)); }
inline f3 normalizef(f3 a) { const float l = 1.f / lengthf(a); return { a.x * l, a.y * l, a.z * l }; }
inline bool eq(f3 a, f3 b) { return a.x == b.x && a.y == b.y && a.z == b.z; }

// a*b - c*d with one compensation step (reference: include/wt/math/eft/eft.hpp:117-125)
inline float diff_prod(float a, float b, float c, float d) {
    const float cd = c * d;
    const float ret = std::fma(a, b, -cd);
    return ret + std::fma(-c, d, cd);
}

// octahedral normal encoding round trip (reference: include/wt/math/encoded_normal.hpp:20-66)
inline void oct_wrap(float vx, float vy, float& ox, float& oy) {
    ox = (1.f - std::fabs(vy)) * (vx >= 0 ? 1.f : -1.f);
    oy = (1.f - std::fabs(vx)) * (vy >= 0 ? 1.f : -1.f);
}
inline f3 encode_decode_normal(f3 n) {
    const float s = std::fabs(n.x) + std::fabs(n.y) + std::fabs(n.z);
    float nx = n.x / s, ny = n.y / s, nz = n.z / s;
    float ex, ey;
    if (nz >= 0) { ex = nx; ey = ny; } else oct_wrap(nx, ny, ex, ey);
    ex = ex * .5f + .5f; ey = ey * .5f + .5f;
    // decode
    ex = ex * 2.f - 1.f; ey = ey * 2.f - 1.f;
    const float z = 1.f - std::fabs(ex) - std::fabs(ey);
    float dx, dy;
    if (z >= 0) { dx = ex; dy = ey; } else oct_wrap(ex, ey, dx, dy);
    return normalizef({ dx, dy, z });
}

struct mesh_tri_t {
    f3 p[3];
    f3 geo_n;
    f3 n[3];
    float uv[3][2];
    bool has_uv;
    f3 dpdu;
    uint32_t shape_idx, shape_tri_idx;
};

// reference: include/wt/mesh/surface_differentials.hpp (surface_differentials_for_triangle)
inline f3 compute_dpdu(const mesh_tri_t& t) {
    const f3 dp02 = t.p[0] - t.p[2], dp12 = t.p[1] - t.p[2];
    float uv0[2] = { 0, 0 }, uv1[2] = { 0, 0 }, uv2[2] = { 0, 0 };
    if (t.has_uv) { memcpy(uv0, t.uv[0], 8); memcpy(uv1, t.uv[1], 8); memcpy(uv2, t.uv[2], 8); }
    const float duv02[2] = { uv0[0] - uv2[0], uv0[1] - uv2[1] };
    const float duv12[2] = { uv1[0] - uv2[0], uv1[1] - uv2[1] };
    const float det = diff_prod(duv02[0], duv12[1], duv02[1], duv12[0]);

    // face normal of the f32 vertices
    f3 ng = { 0, 0, 1 };
    {
        const f3 n = crossf(t.p[1] - t.p[0], t.p[2] - t.p[0]);
        if (!(n.x == 0 && n.y == 0 && n.z == 0)) ng = normalizef(n);
    }
    if (std::fabs(det) < 1e-10f) {
        if (std::fabs(ng.x) > std::fabs(ng.y)) {
            const float l = std::sqrt(ng.x * ng.x + ng.z * ng.z);
            return { -ng.z / l, 0.f, ng.x / l };
        }
        const float l = std::sqrt(ng.y * ng.y + ng.z * ng.z);
        return { 0.f, ng.z / l, -ng.y / l };
    }
    const float r = 1.f / det;
    return {
        diff_prod(duv12[1], dp02.x, duv02[1], dp12.x) * r,
        diff_prod(duv12[1], dp02.y, duv02[1], dp12.y) * r,
        diff_prod(duv12[1], dp02.z, duv02[1], dp12.z) * r,
    };
}

struct bnode_t {
    double mn[3], mx[3];
    int left = -1, right = -1;      // children (inner) or -1
    uint32_t first = 0, count = 0;  // triangle range (valid for every node after DFS)
    bool leaf() const { return left < 0; }
};

struct builder_t {
    static constexpr int BINS = 128;
    static constexpr double C_INT = 100, C_TRAV = 1;

    const std::vector<mesh_tri_t>& tris;
    std::vector<uint32_t> idx;      // permutation
    std::vector<std::array<double, 3>> cen, tmn, tmx;
    std::vector<bnode_t> nodes;

    explicit builder_t(const std::vector<mesh_tri_t>& t) : tris(t) {
        const size_t n = t.size();
        idx.resize(n); std::iota(idx.begin(), idx.end(), 0u);
        cen.resize(n); tmn.resize(n); tmx.resize(n);
        for (size_t i = 0; i < n; ++i) {
            for (int a = 0; a < 3; ++a) {
                const double v0 = (&t[i].p[0].x)[a], v1 = (&t[i].p[1].x)[a], v2 = (&t[i].p[2].x)[a];
                tmn[i][a] = std::min(v0, std::min(v1, v2));
                tmx[i][a] = std::max(v0, std::max(v1, v2));
                cen[i][a] = (v0 + v1 + v2) * (1.0 / 3.0);
            }
        }
        nodes.reserve(2 * n);
    }

    static double half_area(const double* mn, const double* mx) {
        const double ex = mx[0] - mn[0], ey = mx[1] - mn[1], ez = mx[2] - mn[2];
        return ex * ey + ey * ez + ez * ex;
    }

    static constexpr uint32_t kLeafMaxTris = 8;
    int build(uint32_t first, uint32_t count) {
        const int ni = (int)nodes.size();
        nodes.emplace_back();
        {
            bnode_t& n = nodes[ni];
            for (int a = 0; a < 3; ++a) { n.mn[a] = 1e300; n.mx[a] = -1e300; }
            for (uint32_t i = first; i < first + count; ++i)
                for (int a = 0; a < 3; ++a) {
                    n.mn[a] = std::min(n.mn[a], tmn[idx[i]][a]);
                    n.mx[a] = std::max(n.mx[a], tmx[idx[i]][a]);
                }
            n.first = first; n.count = count;
        }
        // Leaves hold up to 8 triangles: the device tests a leaf's triangles with eight lanes at once (gtrav.cuh), and ray queries
        // already treat subtrees of <= 16 triangles as leaves (bvh8w.cpp:29), so finer leaves would only add node steps.
        if (count <= kLeafMaxTris) return ni;

        // centroid bounds
        double cmn[3] = { 1e300, 1e300, 1e300 }, cmx[3] = { -1e300, -1e300, -1e300 };
        for (uint32_t i = first; i < first + count; ++i)
            for (int a = 0; a < 3; ++a) {
                cmn[a] = std::min(cmn[a], cen[idx[i]][a]);
                cmx[a] = std::max(cmx[a], cen[idx[i]][a]);
            }

        double best_cost = 1e300; int best_axis = -1, best_bin = -1;
        for (int a = 0; a < 3; ++a) {
            const double ext = cmx[a] - cmn[a];
            if (!(ext > 0)) continue;
            const double scale = BINS / ext;
            struct bin_t { double mn[3], mx[3]; uint32_t n; };
            static thread_local std::vector<bin_t> bins(BINS);
            for (auto& b : bins) { b.n = 0; for (int k = 0; k < 3; ++k) { b.mn[k] = 1e300; b.mx[k] = -1e300; } }
            for (uint32_t i = first; i < first + count; ++i) {
                const uint32_t t = idx[i];
                int bi = (int)((cen[t][a] - cmn[a]) * scale);
                bi = std::min(BINS - 1, std::max(0, bi));
                auto& b = bins[bi];
                b.n++;
                for (int k = 0; k < 3; ++k) { b.mn[k] = std::min(b.mn[k], tmn[t][k]); b.mx[k] = std::max(b.mx[k], tmx[t][k]); }
            }
            // sweep
            double lA[BINS], rA[BINS]; uint32_t lN[BINS], rN[BINS];
            double amn[3] = { 1e300, 1e300, 1e300 }, amx[3] = { -1e300, -1e300, -1e300 }; uint32_t cnt = 0;
            for (int b = 0; b < BINS - 1; ++b) {
                cnt += bins[b].n;
                for (int k = 0; k < 3; ++k) { amn[k] = std::min(amn[k], bins[b].mn[k]); amx[k] = std::max(amx[k], bins[b].mx[k]); }
                lN[b] = cnt; lA[b] = cnt ? half_area(amn, amx) : 0;
            }
            for (int k = 0; k < 3; ++k) { amn[k] = 1e300; amx[k] = -1e300; } cnt = 0;
            for (int b = BINS - 1; b > 0; --b) {
                cnt += bins[b].n;
                for (int k = 0; k < 3; ++k) { amn[k] = std::min(amn[k], bins[b].mn[k]); amx[k] = std::max(amx[k], bins[b].mx[k]); }
                rN[b - 1] = cnt; rA[b - 1] = cnt ? half_area(amn, amx) : 0;
            }
            for (int b = 0; b < BINS - 1; ++b) {
                if (!lN[b] || !rN[b]) continue;
                const double c = lA[b] * lN[b] + rA[b] * rN[b];
                if (c < best_cost) { best_cost = c; best_axis = a; best_bin = b; }
            }
        }
        if (best_axis < 0) return ni;       // all centroids coincide: leaf

        const double pa = half_area(nodes[ni].mn, nodes[ni].mx);
        const double split_cost = C_TRAV + C_INT * (pa > 0 ? best_cost / pa : best_cost);
        const double leaf_cost = C_INT * count;
        if (split_cost >= leaf_cost) return ni;

        const double scale = BINS / (cmx[best_axis] - cmn[best_axis]);
        auto mid = std::partition(idx.begin() + first, idx.begin() + first + count, [&](uint32_t t) {
            int bi = (int)((cen[t][best_axis] - cmn[best_axis]) * scale);
            bi = std::min(BINS - 1, std::max(0, bi));
            return bi <= best_bin;
        });
        const uint32_t lc = (uint32_t)(mid - (idx.begin() + first));
        if (lc == 0 || lc == count) return ni;

        const int l = build(first, lc);
        const int r = build(first + lc, count - lc);
        nodes[ni].left = l; nodes[ni].right = r;
        return ni;
    }

    double sah_cost(int n, double root_area) const {
        const bnode_t& nd = nodes[n];
        const double a = half_area(nd.mn, nd.mx) / root_area;
        if (nd.leaf()) return C_INT * a * nd.count;
        return C_TRAV * a + sah_cost(nd.left, root_area) + sah_cost(nd.right, root_area);
    }
};

} // namespace

struct wthost_ads {
    std::vector<wtgpu_node> nodes;
    std::vector<wtgpu_leaf> leaves;
    int32_t root_ptr = 0;
    std::vector<wtgpu_tri> tris;
    std::vector<wtgpu_tri_meta> meta;
    std::vector<wtgpu_tri_shading> shading;
    std::vector<wtgpu_edge> edges;
    std::vector<wtgpu_shape> shapes;
    std::vector<uint32_t> shape_tri_tuid;
    std::vector<float> shape_tri_cdf;
    float world_min[3], world_max[3];
    double sah = 0;
    uint32_t max_depth = 0;
};

namespace {

thread_local std::string g_host_err;

void extract(const builder_t& b, int n, int depth, std::vector<int>& out) {
    const bnode_t& nd = b.nodes[n];
    if (nd.leaf() || depth == 0) { out.push_back(n); return; }
    extract(b, nd.left, depth - 1, out);
    extract(b, nd.right, depth - 1, out);
}

uint32_t depth_of(const wthost_ads& a, int32_t ptr, uint32_t d) {
    if (ptr <= 0) return d;
    uint32_t m = d;
    for (int c = 0; c < 8; ++c) m = std::max(m, depth_of(a, a.nodes[ptr - 1].child[c], d + 1));
    return m;
}

struct vkey { uint32_t x, y, z; bool operator==(const vkey& o) const { return x == o.x && y == o.y && z == o.z; } };
struct vhash { size_t operator()(const vkey& k) const { return ((size_t)k.x * 73856093u) ^ ((size_t)k.y * 19349663u) ^ ((size_t)k.z * 83492791u); } };
inline vkey key_of(f3 p) {
    vkey k; float x = p.x + 0.f, y = p.y + 0.f, z = p.z + 0.f;  // -0 -> +0
    memcpy(&k.x, &x, 4); memcpy(&k.y, &y, 4); memcpy(&k.z, &z, 4);
    return k;
}

struct atri_t { f3 a, b, c, n; uint32_t e_ab = WTGPU_INVALID_IDX, e_bc = WTGPU_INVALID_IDX, e_ca = WTGPU_INVALID_IDX; };

// the reference's vector algebra (include/wt/math/vecmath.hpp:21-66) for edge_for below: dot = product + fused multiply-adds in component order, cross by
// compensated products, normalize = division by the length.  (The CPU test suite compares the table with the reference's own edge_for bit for bit:
// test_host_edge_table_equals_the_reference_code.)
inline float dot_r(f3 a, f3 b) { return std::fma(a.z, b.z, std::fma(a.y, b.y, a.x * b.x)); }
inline f3 cross_r(f3 x, f3 y) { return { diff_prod(x.y, y.z, x.z, y.y), diff_prod(x.z, y.x, x.x, y.z), diff_prod(x.x, y.y, x.y, y.x) }; }
inline f3 normalize_r(f3 a) { const float l = std::sqrt(dot_r(a, a)); return { a.x / l, a.y / l, a.z / l }; }

// reference: edge_classification.hpp:31-86 (edge_for)
bool edge_for(const atri_t* t1, const atri_t* t2, uint32_t tuid1, uint32_t tuid2,
              f3 a, f3 b, f3 c1, const f3* c2, wtgpu_edge& out) {
    f3 n1 = t1->n;
    f3 n2 = t2 ? t2->n : -n1;
    const f3 e = normalize_r(b - a);
    const f3 m = { (a.x + b.x) / 2.f, (a.y + b.y) / 2.f, (a.z + b.z) / 2.f };
    f3 tt1 = { 0, 0, 1 }, tt2 = { 0, 0, 1 };
    if (t2) {
        const bool concave1 = dot_r(n1, *c2 - m) > 0;
        const bool concave2 = dot_r(n2, c1 - m) > 0;
        if (concave1 != concave2) return false;     // inconsistent normals
        if (concave1 && concave2) { n1 = -n1; n2 = -n2; }
        tt2 = cross_r(n2, e);
        if (dot_r(tt2, *c2 - m) < 0) tt2 = -tt2;
    }
    tt1 = cross_r(n1, e);
    if (dot_r(tt1, c1 - m) < 0) tt1 = -tt1;
    if (!t2) tt2 = tt1;

    const float pi = 3.14159265358979323846f;
    const float d = std::min(1.f, std::max(-1.f, dot_r(n1, n2)));
    const float alpha = std::max(0.f, pi - std::acos(d));
    if (alpha > 160.f / 180.f * pi) return false;

    memcpy(out.a, &a, 12); memcpy(out.b, &b, 12); memcpy(out.e, &e, 12);
    memcpy(out.n1, &n1, 12); memcpy(out.t1, &tt1, 12);
    memcpy(out.n2, &n2, 12); memcpy(out.t2, &tt2, 12);
    out.alpha = alpha;
    out.tri1 = tuid1; out.tri2 = t2 ? tuid2 : WTGPU_INVALID_IDX;
    return true;
}

void find_all_edges(std::vector<atri_t>& tris, std::vector<wtgpu_edge>& edges) {
    std::unordered_map<vkey, std::vector<uint32_t>, vhash> vmap;
    vmap.reserve(tris.size() * 2);
    for (uint32_t t = 0; t < tris.size(); ++t) {
        vmap[key_of(tris[t].a)].push_back(t);
        vmap[key_of(tris[t].b)].push_back(t);
        vmap[key_of(tris[t].c)].push_back(t);
    }
    std::vector<uint32_t> cand;
    for (uint32_t tuid = 0; tuid < tris.size(); ++tuid) {
        atri_t* tri = &tris[tuid];
        cand.clear();
        for (const f3& v : { tri->a, tri->b, tri->c }) {
            const auto& l = vmap[key_of(v)];
            cand.insert(cand.end(), l.begin(), l.end());
        }
        std::sort(cand.begin(), cand.end());
        cand.erase(std::unique(cand.begin(), cand.end()), cand.end());

        bool found_ab = false, found_bc = false, found_ca = false;
        auto insert_edge = [&](bool ok, const wtgpu_edge& e, uint32_t* e1, uint32_t* e2) {
            if (!ok) return;
            const uint32_t eid = (uint32_t)edges.size();
            *e1 = eid; if (e2) *e2 = eid;
            edges.push_back(e);
        };

        for (uint32_t t : cand) {
            if (t == tuid) continue;
            atri_t& other = tris[t];
            const bool fa = eq(tri->a, other.a) || eq(tri->a, other.b) || eq(tri->a, other.c);
            const bool fb = eq(tri->b, other.a) || eq(tri->b, other.b) || eq(tri->b, other.c);
            const bool fc = eq(tri->c, other.a) || eq(tri->c, other.b) || eq(tri->c, other.c);
            const bool sa = eq(other.a, tri->a) || eq(other.a, tri->b) || eq(other.a, tri->c);
            const bool sb = eq(other.b, tri->a) || eq(other.b, tri->b) || eq(other.b, tri->c);
            const bool sc = eq(other.c, tri->a) || eq(other.c, tri->b) || eq(other.c, tri->c);
            uint32_t* te2 = sa && sb ? &other.e_ab : sb && sc ? &other.e_bc : &other.e_ca;
            const f3 c2 = sa && sb ? other.c : sb && sc ? other.a : other.b;
            wtgpu_edge e{};
            if (fa && fb) {
                if (found_ab) continue;
                found_ab = true;
                if (t <= tuid) continue;
                insert_edge(edge_for(tri, &other, tuid, t, tri->a, tri->b, tri->c, &c2, e), e, &tri->e_ab, te2);
            }
            if (fb && fc) {
                if (found_bc) continue;
                found_bc = true;
                if (t <= tuid) continue;
                insert_edge(edge_for(tri, &other, tuid, t, tri->b, tri->c, tri->a, &c2, e), e, &tri->e_bc, te2);
            }
            if (fc && fa) {
                if (found_ca) continue;
                found_ca = true;
                if (t <= tuid) continue;
                insert_edge(edge_for(tri, &other, tuid, t, tri->c, tri->a, tri->b, &c2, e), e, &tri->e_ca, te2);
            }
        }
        wtgpu_edge e{};
        if (!found_ab) insert_edge(edge_for(tri, nullptr, tuid, 0, tri->a, tri->b, tri->c, nullptr, e), e, &tri->e_ab, nullptr);
        if (!found_bc) insert_edge(edge_for(tri, nullptr, tuid, 0, tri->b, tri->c, tri->a, nullptr, e), e, &tri->e_bc, nullptr);
        if (!found_ca) insert_edge(edge_for(tri, nullptr, tuid, 0, tri->c, tri->a, tri->b, nullptr, e), e, &tri->e_ca, nullptr);
    }
}

} // namespace

extern "C" {

int wthost_ads_build(uint32_t n_meshes, const wthost_mesh_desc* meshes, wthost_ads** out) {
    if (!out || (!meshes && n_meshes)) return WTGPU_E_INVALID;
    auto ads = std::make_unique<wthost_ads>();

    // ---- triangles per shape (mesh.cpp:31-101)
    std::vector<mesh_tri_t> all;
    std::vector<uint32_t> shape_first(n_meshes + 1, 0);
    for (uint32_t s = 0; s < n_meshes; ++s) {
        const wthost_mesh_desc& m = meshes[s];
        const double* M = m.to_world;
        auto xpoint = [&](const float* p) {
            const double x = p[0], y = p[1], z = p[2];
            return d3{ M[0] * x + M[1] * y + M[2] * z + M[3], M[4] * x + M[5] * y + M[6] * z + M[7], M[8] * x + M[9] * y + M[10] * z + M[11] };
        };
        auto xdir = [&](const float* p) {
            double x = p[0], y = p[1], z = p[2];
            const double l = std::sqrt(x * x + y * y + z * z); x /= l; y /= l; z /= l;
            d3 r{ M[0] * x + M[1] * y + M[2] * z, M[4] * x + M[5] * y + M[6] * z, M[8] * x + M[9] * y + M[10] * z };
            const double rl = std::sqrt(dot(r, r));
            return f3{ (float)(r.x / rl), (float)(r.y / rl), (float)(r.z / rl) };
        };
        uint32_t shape_tri_idx = 0;
        float area = 0;
        std::vector<float> areas;
        for (uint32_t t = 0; t < m.n_tris; ++t) {
            uint32_t i0 = m.indices[3 * t], i1 = m.indices[3 * t + 1], i2 = m.indices[3 * t + 2];
            if (i0 >= m.n_verts || i1 >= m.n_verts || i2 >= m.n_verts) { g_host_err = "mesh index out of range"; return WTGPU_E_INVALID; }
            d3 a = xpoint(m.positions + 3 * i0), b = xpoint(m.positions + 3 * i1), c = xpoint(m.positions + 3 * i2);
            const d3 n = cross(b - a, c - a);
            if (n.x == 0 && n.y == 0 && n.z == 0) continue;     // degenerate
            const double nl = std::sqrt(dot(n, n));
            f3 gn = { (float)(n.x / nl), (float)(n.y / nl), (float)(n.z / nl) };

            mesh_tri_t mt{};
            mt.has_uv = m.uvs != nullptr;
            if (mt.has_uv) {
                memcpy(mt.uv[0], m.uvs + 2 * i0, 8); memcpy(mt.uv[1], m.uvs + 2 * i1, 8); memcpy(mt.uv[2], m.uvs + 2 * i2, 8);
            }
            if (m.normals) {
                f3 n1 = xdir(m.normals + 3 * i0), n2 = xdir(m.normals + 3 * i1), n3 = xdir(m.normals + 3 * i2);
                if (dotf(n1, gn) < 0 && dotf(n2, gn) < 0 && dotf(n3, gn) < 0) {
                    std::swap(a, b);
                    if (mt.has_uv) { std::swap(mt.uv[0][0], mt.uv[1][0]); std::swap(mt.uv[0][1], mt.uv[1][1]); }
                    std::swap(n1, n2);
                    gn = -gn;
                }
                mt.n[0] = encode_decode_normal(n1); mt.n[1] = encode_decode_normal(n2); mt.n[2] = encode_decode_normal(n3);
            } else {
                const f3 en = encode_decode_normal(gn);
                mt.n[0] = mt.n[1] = mt.n[2] = en;
            }
            mt.p[0] = { (float)a.x, (float)a.y, (float)a.z };
            mt.p[1] = { (float)b.x, (float)b.y, (float)b.z };
            mt.p[2] = { (float)c.x, (float)c.y, (float)c.z };
            mt.geo_n = gn;
            mt.dpdu = compute_dpdu(mt);
            mt.shape_idx = s; mt.shape_tri_idx = shape_tri_idx++;
            all.push_back(mt);
            // util::tri_surface_area (include/wt/math/util.hpp:190-196)
            const float ar = .5f * lengthf(crossf(mt.p[2] - mt.p[0], mt.p[1] - mt.p[0]));
            areas.push_back(ar); area += ar;
        }
        shape_first[s + 1] = (uint32_t)all.size();

        wtgpu_shape sh{};
        sh.bsdf = m.bsdf; sh.emitter = m.emitter; sh.surface_area = area;
        sh.tri_first = shape_first[s]; sh.n_tris = shape_tri_idx;
        sh.cdf_first = (uint32_t)ads->shape_tri_cdf.size();
        // discrete_distribution_t (include/wt/math/distribution/discrete_distribution.hpp:45-66)
        {
            std::vector<float> dcdf(areas.size() + 1, 0.f);
            for (size_t i = 0; i < areas.size(); ++i) dcdf[i + 1] = dcdf[i] + std::max(0.f, areas[i]);
            const float sum = dcdf.back();
            if (sum > 0) { const float r = 1.f / sum; for (auto& v : dcdf) v *= r; } else dcdf.back() = 1;
            ads->shape_tri_cdf.insert(ads->shape_tri_cdf.end(), dcdf.begin(), dcdf.end());
        }
        ads->shapes.push_back(sh);
    }
    if (all.empty()) { g_host_err = "(bvh_constructor) no triangles found!"; return WTGPU_E_INVALID; }

    // ---- binary BVH
    builder_t b(all);
    const int root = b.build(0, (uint32_t)all.size());
    {
        const double ra = builder_t::half_area(b.nodes[root].mn, b.nodes[root].mx);
        ads->sah = ra > 0 ? b.sah_cost(root, ra) : 0;
    }

    // triangles in DFS (leaf) order: idx[] already is, because build() partitions in place
    const uint32_t nt = (uint32_t)all.size();
    ads->tris.resize(nt); ads->meta.resize(nt); ads->shading.resize(nt);
    ads->shape_tri_tuid.assign(nt, 0);
    std::vector<atri_t> atris(nt);
    for (uint32_t tuid = 0; tuid < nt; ++tuid) {
        const mesh_tri_t& mt = all[b.idx[tuid]];
        wtgpu_tri& t = ads->tris[tuid];
        t.ax = mt.p[0].x; t.ay = mt.p[0].y; t.az = mt.p[0].z;
        t.bx = mt.p[1].x; t.by = mt.p[1].y; t.bz = mt.p[1].z;
        t.cx = mt.p[2].x; t.cy = mt.p[2].y; t.cz = mt.p[2].z;
        t.nx = mt.geo_n.x; t.ny = mt.geo_n.y; t.nz = mt.geo_n.z;
        wtgpu_tri_shading& sh = ads->shading[tuid];
        memcpy(sh.n0, &mt.n[0], 12); memcpy(sh.n1, &mt.n[1], 12); memcpy(sh.n2, &mt.n[2], 12);
        memcpy(sh.uv0, mt.uv[0], 8); memcpy(sh.uv1, mt.uv[1], 8); memcpy(sh.uv2, mt.uv[2], 8);
        memcpy(sh.dpdu, &mt.dpdu, 12);
        sh.has_uv = mt.has_uv ? 1u : 0u;
        ads->meta[tuid] = wtgpu_tri_meta{ mt.shape_idx, mt.shape_tri_idx, WTGPU_INVALID_IDX, WTGPU_INVALID_IDX, WTGPU_INVALID_IDX, { 0, 0, 0 } };
        ads->shape_tri_tuid[shape_first[mt.shape_idx] + mt.shape_tri_idx] = tuid;
        atris[tuid].a = mt.p[0]; atris[tuid].b = mt.p[1]; atris[tuid].c = mt.p[2]; atris[tuid].n = mt.geo_n;
    }

    // ---- 8-wide encode (bvh8w_constructor.cpp:59-103, 206-243): BFS over work items
    struct work_t { int bnode; uint32_t w8; };
    std::deque<work_t> q;
    ads->nodes.emplace_back(); memset(&ads->nodes[0], 0, sizeof(wtgpu_node));
    q.push_back({ root, 0 });
    while (!q.empty()) {
        const work_t w = q.front(); q.pop_front();
        std::vector<int> ch; ch.reserve(8);
        extract(b, w.bnode, 3, ch);
        wtgpu_node nd; memset(&nd, 0, sizeof(nd));
        nd.tris_start = b.nodes[w.bnode].first; nd.tris_count = b.nodes[w.bnode].count;
        for (size_t c = 0; c < ch.size(); ++c) {
            const bnode_t& bn = b.nodes[ch[c]];
            nd.minx[c] = (float)bn.mn[0]; nd.miny[c] = (float)bn.mn[1]; nd.minz[c] = (float)bn.mn[2];
            nd.maxx[c] = (float)bn.mx[0]; nd.maxy[c] = (float)bn.mx[1]; nd.maxz[c] = (float)bn.mx[2];
            if (bn.leaf()) {
                ads->leaves.push_back({ bn.first, bn.count });
                nd.child[c] = -(int32_t)ads->leaves.size();
            } else {
                const uint32_t cidx = (uint32_t)ads->nodes.size();
                ads->nodes.emplace_back(); memset(&ads->nodes.back(), 0, sizeof(wtgpu_node));
                nd.child[c] = (int32_t)cidx + 1;
                q.push_back({ ch[c], cidx });
            }
        }
        ads->nodes[w.w8] = nd;
    }
    ads->root_ptr = 1;
    for (int a = 0; a < 3; ++a) { ads->world_min[a] = (float)b.nodes[root].mn[a]; ads->world_max[a] = (float)b.nodes[root].mx[a]; }
    ads->max_depth = depth_of(*ads, ads->root_ptr, 0);

    // ---- edges
    find_all_edges(atris, ads->edges);
    for (uint32_t t = 0; t < nt; ++t) {
        ads->meta[t].edge_ab = atris[t].e_ab; ads->meta[t].edge_bc = atris[t].e_bc; ads->meta[t].edge_ca = atris[t].e_ca;
    }

    *out = ads.release();
    return WTGPU_OK;
}

int wthost_ads_fill(const wthost_ads* a, wtgpu_scene_desc* d) {
    if (!a || !d) return WTGPU_E_INVALID;
    d->n_nodes = (uint32_t)a->nodes.size(); d->nodes = a->nodes.data();
    d->n_leaves = (uint32_t)a->leaves.size(); d->leaves = a->leaves.data();
    d->root_ptr = a->root_ptr;
    d->n_tris = (uint32_t)a->tris.size(); d->tris = a->tris.data(); d->tri_meta = a->meta.data(); d->tri_shading = a->shading.data();
    d->n_edges = (uint32_t)a->edges.size(); d->edges = a->edges.data();
    memcpy(d->world_min, a->world_min, 12); memcpy(d->world_max, a->world_max, 12);
    d->n_shapes = (uint32_t)a->shapes.size(); d->shapes = a->shapes.data();
    d->n_shape_tris = (uint32_t)a->shape_tri_tuid.size(); d->shape_tri_tuid = a->shape_tri_tuid.data();
    d->n_shape_cdf = (uint32_t)a->shape_tri_cdf.size(); d->shape_tri_cdf = a->shape_tri_cdf.data();
    return WTGPU_OK;
}

void wthost_ads_destroy(wthost_ads* a) { delete a; }
double wthost_ads_sah_cost(const wthost_ads* a) { return a ? a->sah : 0; }
uint32_t wthost_ads_max_depth(const wthost_ads* a) { return a ? a->max_depth : 0; }

} // extern "C"
