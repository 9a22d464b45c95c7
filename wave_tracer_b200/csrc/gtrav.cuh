// gtrav.cuh -- group-cooperative traversal: integrator::traverse (include/wt/integrator/traversal.hpp:94-172) with the 8-wide BVH
// queries of src/ads/bvh8w.cpp (ray: 469-554, cone: 123-347) executed by EIGHT LANES PER QUERY.  (Included by wavefront.cu.)
//
// Why: one thread per beam (dtrav.cuh) leaves a warp running ~2-4 lanes on cone-heavy scenes -- every lane is in a different place
// of a different query, and a few wide beams take 100x longer than the rest.  Here a group of 8 lanes owns one beam at a time:
//   * node step: lane i tests child i of the 256-B node (the node's SoA rows are read as 32-B segments), the surviving children
//     are ranked with shuffles and pushed in the order the sequential insertion sort would leave them;
//   * leaf step: lane i intersects triangle i of the leaf (ray: of a <=16-triangle subtree); hits are appended with a ballot
//     prefix, the closest is found with a (distance, lane) min-reduction = the first minimal one in sequential order;
//   * the stack lives in shared memory (1 KB per group); accepted triangles go straight to the beam's row of the result array in HBM, whose
//     length is a run-time capacity (dtrav.cuh Caps) -- the count keeps running past it, so a too-short row is known exactly;
//   * groups pull beams from the list with an atomic cursor, so a long beam delays only its own group; a cone query that has tested more than
//     TravTiers::big_tested triangles is handed over, state and stack, to the team traversal (ctrav.cuh);
//   * the four groups of a warp run node steps freely but meet before every leaf step, where the expensive cone-triangle code runs.
// Every decision is taken on the same values, in the same order, as the sequential code: results are bit-identical to dtrav.cuh.
#pragma once
#include <cuda_pipeline.h>

namespace wt {

constexpr int kGW = 8;
constexpr int kGStack = 128;
#ifdef WT_NODE_STAGING
struct alignas(16) GShared { float tmin[kGStack]; int32_t ptr[kGStack]; float key[kGW]; alignas(16) float node[64]; };     // node: staging slot of the 256-B node record in hand
#else
struct alignas(16) GShared { float tmin[kGStack]; int32_t ptr[kGStack]; float key[kGW]; float node[1]; };
#endif

// what traverse() returns, as stored between the traversal kernel and the per-thread resolve kernel
struct alignas(16) TravRec { uint32_t flags, ray_tuid; float ray_dist, bx, by, cone_dist; uint32_t n_tris; float region_depth, ox, oy, oz; uint32_t pad_; };
enum : uint32_t { TR_EMPTY = 1u, TR_BALLISTIC = 2u, TR_CONE_FRONT = 4u, TR_OVERFLOW = 8u, TR_RAY_FRONT = 16u };

struct GLane { unsigned gl, gshift, gmask; };
WT_D unsigned g_ballot(const GLane& g, bool p) { return (__ballot_sync(g.gmask, p) >> g.gshift) & 0xffu; }
template <class T> WT_D T g_shfl(const GLane& g, T v, int src) { return __shfl_sync(g.gmask, v, src, kGW); }
// (value, lane) lexicographic minimum over the lanes with `valid`; returns the winning lane (or -1)
WT_D int g_argmin(const GLane& g, float v, bool valid) {
    float bv = valid ? v : WT_INF; int bl = valid ? (int)g.gl : 64;
#pragma unroll
    for (int o = 1; o < kGW; o <<= 1) {
        const float ov = __shfl_xor_sync(g.gmask, bv, o, kGW); const int ol = __shfl_xor_sync(g.gmask, bl, o, kGW);
        if (ol < 64 && (bl >= 64 || ov < bv || (ov == bv && ol < bl))) { bv = ov; bl = ol; }
    }
    return bl < 64 ? bl : -1;
}

// the same over a whole warp
WT_D int g_argmin32(float v, bool valid, unsigned lane) {
    float bv = valid ? v : WT_INF; int bl = valid ? (int)lane : 64;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float ov = __shfl_xor_sync(0xffffffffu, bv, o); const int ol = __shfl_xor_sync(0xffffffffu, bl, o);
        if (ol < 64 && (bl >= 64 || ov < bv || (ov == bv && ol < bl))) { bv = ov; bl = ol; }
    }
    return bl < 64 ? bl : -1;
}

// One group's traversal machine.  All members hold the same value in the 8 lanes of a group.
struct GTrav {
    // beam
    Cone env; float lambda, dist, bd, min_prog; uint32_t seg; int wstate;       // wstate: 0 ray-only, 1 segment ray, 2 segment cone, 3 edge query around a ballistic hit
    float zs; bool edge_query;  // z_search_range_scale of the current cone query; plt_path: follow a ballistic hit of a finite beam with the edge query
    // query
    int mode, s;                // mode: 1 ray, 2 cone; s: stack size
    V3 inv; bool nx, ny, nz; Frame frame;
    Range qrange, crange;       // ray range / cone traversal range; current cone search range
    RayCull cull;               // range culling bounds of the current ray query (dtrav.cuh)
    RayHit rec; ConeResult res;
    uint32_t qtested;           // triangles the current cone query has tested (hand-over trigger)
    TriWriter tw;               // where the current cone query's triangle list goes (dtrav.cuh TriList): the beam's row + extents of the spill arena
};
constexpr uint32_t kBigQuery = 32u;     // a triangle list longer than this is resolved by the warp-per-list resolve kernels (k_*_resolve_big)
// A beam in the middle of a cone query, as handed from one traversal kernel to the next (ctrav.cuh): the whole machine state and the stack.
struct alignas(16) TravSave { GTrav t; int item; int pad_[3]; float tmin[kGStack]; int32_t ptr[kGStack]; };
// hand-over thresholds (triangles tested by the current cone query): group -> warp team, warp team -> block team
// (TravTiers is declared in wavefront.cu ahead of RenderArgs: { big_tested, huge_tested, restart_div, init_div }) restart_div / init_div: the team batch after a range change / at the start is NL / div leaves

WT_D void g_start_ray(const DScene& sc, const GLane& g, GShared& sh, GTrav& t, Range r, Counters& ctr) {
    t.mode = 1; t.qrange = r; t.cull = ray_cull(sc, r); t.rec.tuid = WTGPU_INVALID_IDX; t.rec.dist = WT_INF; t.rec.bx = t.rec.by = -1.f; t.rec.front = false;
    t.s = 1;
    if (g.gl == 0u) { sh.tmin[0] = 0.f; sh.ptr[0] = sc.root_ptr; ctr.ray_casts++; }
    __syncwarp(g.gmask);
}
WT_D void g_start_cone(const DScene& sc, const GLane& g, GShared& sh, GTrav& t, Range tr, Counters& ctr) {
    t.mode = 2; t.qtested = 0u; t.qrange = tr; t.res.dist = WT_INF; t.res.front = false; t.res.n_tris = 0u; t.res.overflow = false; tw_begin(t.tw);
    t.crange = cone_search_range(t.env, tr, t.res.dist, t.zs);
    t.s = 1;
    if (g.gl == 0u) { sh.tmin[0] = 0.f; sh.ptr[0] = sc.root_ptr; ctr.cone_casts++; }
    __syncwarp(g.gmask);
}
// push the children that passed, in the order insertion sort (descending tmin, stable) leaves them (bvh8w.cpp:44-57)
WT_D void g_push_sorted(const GLane& g, GShared& sh, GTrav& t, bool push, float tmin, int32_t ch, int cap, Counters& ctr) {
    const unsigned m = g_ballot(g, push);
    const int idx = __popc(m & ((1u << g.gl) - 1u));
    const bool keep = push && t.s + idx < cap;
    const unsigned km = g_ballot(g, keep);
    if (km != m && g.gl == 0u) ctr.stack_drops += (uint32_t)(__popc(m) - __popc(km));
    int rank = 0;
    if (__popc(km) > 1) {       // (group-uniform) nothing to order when at most one child passed (etoile-like k_gtraverse -6 %)
    // the eight keys go through shared memory (one store, two 16-B loads) instead of eight shuffles (fewer instructions in the hottest loop: etoile-like k_gtraverse 111 -> 96 ms/step); lanes that do not push publish -inf
    sh.key[g.gl] = keep ? tmin : -WT_INF;
    __syncwarp(g.gmask);
    {
        const float4 k0 = *reinterpret_cast<const float4*>(&sh.key[0]), k1 = *reinterpret_cast<const float4*>(&sh.key[4]);
        const float kk[8] = { k0.x, k0.y, k0.z, k0.w, k1.x, k1.y, k1.z, k1.w };
#pragma unroll
        for (int j = 0; j < kGW; ++j) rank += (kk[j] > tmin || (kk[j] == tmin && j < (int)g.gl)) ? 1 : 0;
    }
    __syncwarp(g.gmask);
    }

    if (keep) { sh.tmin[t.s + rank] = tmin; sh.ptr[t.s + rank] = ch; }
    t.s += __popc(km);
    __syncwarp(g.gmask);
}
// The 256-B node record (eight rows of eight words: min x/y/z, max x/y/z, child pointers, triangle range); lane i gets word i of each row.
// -DWT_NODE_STAGING: the record goes global -> shared memory with two 16-B asynchronous copies per lane (LDGSTS: the group's eight lanes cover
// the record's two 128-B lines in one coalesced request each) and is read back from the group's 64-word `slot`.  MEASURED SLOWER than the plain
// loads on every workload (profiles/r02_ab_node_staging.log: cornell -3.5 %, etoile -1.5 %, double_slits -4 %, sponza -35 %): the copy has to be
// waited for as a whole and costs a shared-memory round trip, whereas the seven sector loads are all in flight at once and feed the slab test
// directly; the node table (4 MB) lives in L2 / L1 either way.  Kept as the A/B switch.
WT_D void g_stage_node(const DScene& sc, const GLane& g, float* slot, int32_t ptr, float& mnx, float& mny, float& mnz, float& mxx, float& mxy, float& mxz, int32_t& ch) {
#ifndef WT_NODE_STAGING         // default: seven 4-B loads per lane straight from global memory (each a 32-B sector per group), consumed as they arrive
    const wtgpu_node* __restrict__ n = sc.nodes + (ptr - 1);
    mnx = __ldg(&n->minx[g.gl]); mny = __ldg(&n->miny[g.gl]); mnz = __ldg(&n->minz[g.gl]);
    mxx = __ldg(&n->maxx[g.gl]); mxy = __ldg(&n->maxy[g.gl]); mxz = __ldg(&n->maxz[g.gl]);
    ch = __ldg(&n->child[g.gl]);
#else
    const float4* __restrict__ src = reinterpret_cast<const float4*>(sc.nodes + (ptr - 1)) + 2u * g.gl;
    float4* dst = reinterpret_cast<float4*>(slot) + 2u * g.gl;
    __pipeline_memcpy_async(dst, src, 16); __pipeline_memcpy_async(dst + 1, src + 1, 16);
    __pipeline_commit(); __pipeline_wait_prior(0);
    __syncwarp(g.gmask);
    mnx = slot[0 * 8 + g.gl]; mny = slot[1 * 8 + g.gl]; mnz = slot[2 * 8 + g.gl];
    mxx = slot[3 * 8 + g.gl]; mxy = slot[4 * 8 + g.gl]; mxz = slot[5 * 8 + g.gl];
    ch = __float_as_int(slot[6 * 8 + g.gl]);
    __syncwarp(g.gmask);
#endif
}
// cone_cluster_intersect (bvh8w.cpp:187-230) for one child box: the AABB inflated by the cone radius at its farthest z, slab test against the
// current search range; true = the child is pushed (with key tmin)
WT_D bool cone_child_test(V3 ro, V3 rd, V3 inv, bool nx, bool ny, bool nz, float ta, float x0, Range cr, float mnx, float mny, float mnz, float mxx, float mxy, float mxz, float& tmin_out) {
    float omnx = mnx - ro.x, omny = mny - ro.y, omnz = mnz - ro.z;
    float omxx = mxx - ro.x, omxy = mxy - ro.y, omxz = mxz - ro.z;
    const float bx = nx ? omnx : omxx, by = ny ? omny : omxy, bz = nz ? omnz : omxz;
    const float ddb = fmaf(rd.z, bz, fmaf(rd.y, by, rd.x * bx));
    const float maxz = fminf(fmaxf(ddb, 0.f), cr.mx);
    const float enl = fmaf(maxz, ta, x0);
    omnx -= enl; omny -= enl; omnz -= enl; omxx += enl; omxy += enl; omxz += enl;
    const float dminx = (nx ? omxx : omnx) * inv.x, dminy = (ny ? omxy : omny) * inv.y, dminz = (nz ? omxz : omnz) * inv.z;
    const float dmaxx = (nx ? omnx : omxx) * inv.x, dmaxy = (ny ? omny : omxy) * inv.y, dmaxz = (nz ? omnz : omxz) * inv.z;
    float tmin = 0.f, tmax = dmaxx;
    tmin = vmaxps(tmin, dminx); tmax = vminps(tmax, dmaxy);
    tmin = vmaxps(tmin, dminy); tmax = vminps(tmax, dmaxz);
    tmin = vmaxps(tmin, dminz);
    tmin_out = tmin;
    const bool ok = tmin <= tmax && tmax >= cr.mn && tmin <= cr.mx;
    return ok && !(tmin >= cr.mx);
}
WT_D void g_node_step(const DScene& sc, const GLane& g, GShared& sh, GTrav& t, int32_t ptr, Counters& ctr) {
    if (g.gl == 0u) ctr.nodes++;
    float mnx, mny, mnz, mxx, mxy, mxz; int32_t ch;
    g_stage_node(sc, g, sh.node, ptr, mnx, mny, mnz, mxx, mxy, mxz, ch);
    const V3 ro = t.env.o, rd = t.env.d;
    bool push; float key; int cap;       // (one copy of the ranked push for both query kinds: code size is what bounds this kernel)
    if (t.mode == 1) {      // intersect_ray_aabb_fast (intersect/ray.hpp:331-351), range {0, closest hit}
        const float t1x = ((t.nx ? mxx : mnx) - ro.x) * t.inv.x, t2x = ((t.nx ? mnx : mxx) - ro.x) * t.inv.x;
        const float t1y = ((t.ny ? mxy : mny) - ro.y) * t.inv.y, t2y = ((t.ny ? mny : mxy) - ro.y) * t.inv.y;
        const float t1z = ((t.nz ? mxz : mnz) - ro.z) * t.inv.z, t2z = ((t.nz ? mnz : mxz) - ro.z) * t.inv.z;
        const float rmin = vmaxps(vmaxps(t1x, t1y), vmaxps(t1z, 0.f));        // (x, y), (z, range): simd/math.hpp:333-356 -- decides when a slab gives NaN
        const float rmax = vminps(vminps(t2x, t2y), vminps(t2z, t.rec.dist));
        push = rmin <= rmax && ch != 0 && ray_cull_keep(t.cull, rmin, rmax); key = rmin; cap = 64;
    } else {                // cone_cluster_intersect (bvh8w.cpp:187-230)
        float tmin;
        push = cone_child_test(ro, rd, t.inv, t.nx, t.ny, t.nz, t.env.ta, t.env.x0, t.crange, mnx, mny, mnz, mxx, mxy, mxz, tmin) && ch != 0; key = tmin; cap = kGStack;
    }
    g_push_sorted(g, sh, t, push, key, ch, cap, ctr);
}
// triangles [t0, t0+cnt) against the current query
WT_D void g_leaf_step(const DScene& sc, const GLane& g, GShared& sh, GTrav& t, uint32_t t0, uint32_t cnt, Counters& ctr) {
    const V3 ro = t.env.o, rd = t.env.d;
    if (t.mode == 1) {      // ray_gather (bvh8w.cpp:394-467)
        bool hit = false;
        for (uint32_t base = 0; base < cnt; base += (uint32_t)kGW) {
            const uint32_t k = base + g.gl; const bool valid = k < cnt; const uint32_t tuid = t0 + k;
            float z = -WT_INF, bx = 0.f, by = 0.f; bool front = false;
            if (valid) { const Tri3 tr = load_tri(sc, tuid); ctr.tris++; z = intersect_ray_tri_w(ro, rd, tr.a, tr.b, tr.c, t.qrange, bx, by); front = dot(tr.n, rd) <= 0.f; }
            const int w = g_argmin(g, z, valid && z != -WT_INF && z < t.rec.dist);
            if (w >= 0) { t.rec.dist = g_shfl(g, z, w); t.rec.bx = g_shfl(g, bx, w); t.rec.by = g_shfl(g, by, w); t.rec.tuid = g_shfl(g, tuid, w); t.rec.front = g_shfl(g, front ? 1 : 0, w) != 0; hit = true; }
        }
        if (hit) while (t.s > 0 && sh.tmin[t.s - 1] >= t.rec.dist) --t.s;
    } else {                // gather_tris (bvh8w.cpp:123-185)
        bool found = false;
        t.qtested += cnt;
        for (uint32_t base = 0; base < cnt; base += (uint32_t)kGW) {
            const uint32_t k = base + g.gl; const bool valid = k < cnt; const uint32_t tuid = t0 + k;
            float d = WT_INF; bool front = false;
            if (valid) { const Tri3 tr = load_tri(sc, tuid); ctr.tris++; d = intersect_cone_tri(t.env, t.frame, tr.a, tr.b, tr.c, tr.n, t.crange); front = dot(tr.n, -rd) > 0.f; }
            const bool acc = d < WT_INF && !(d > t.crange.mx);
            const unsigned m = g_ballot(g, acc);
            if (m) {
                found = true;
                const int w = g_argmin(g, d, acc && d < t.res.dist);
                if (w >= 0) { t.res.dist = g_shfl(g, d, w); t.res.front = g_shfl(g, front ? 1 : 0, w) != 0; }
                const uint32_t upto = t.res.n_tris + (uint32_t)__popc(m);
                if (upto > t.tw.alloc_end && !t.tw.fail) {      // (group-uniform) the leader extends the list's storage
                    if (g.gl == 0u) tw_reserve(sc, t.tw, upto);
                    t.tw.n_ext = g_shfl(g, t.tw.n_ext, 0); t.tw.alloc_end = g_shfl(g, t.tw.alloc_end, 0); t.tw.fail = g_shfl(g, t.tw.fail ? 1 : 0, 0) != 0;
                    __syncwarp(g.gmask);
                }
                if (acc) tw_put(sc, t.tw, t.res.n_tris + (uint32_t)__popc(m & ((1u << g.gl) - 1u)), tuid);
                t.res.n_tris = upto;
                if (t.tw.fail) t.res.overflow = true;
            }
        }
        if (found) {
            t.crange = cone_search_range(t.env, t.qrange, t.res.dist, t.zs);
            while (t.s > 0 && sh.tmin[t.s - 1] >= t.crange.mx) --t.s;
        }
        __syncwarp(g.gmask);
    }
}
// beam set-up: integrator::traverse up to the first query
WT_D void g_begin(const DScene& sc, const GLane& g, GShared& sh, GTrav& t, const Cone& env0, const Geo& prev, float lambda, bool force_rt, bool edge_query, Counters& ctr) {
    t.env = env0; t.env.o = offseted_ray_origin(sc, prev, env0.o, env0.d);
    t.lambda = lambda; t.zs = kMajorToZ; t.edge_query = edge_query;
    const V3 rd = t.env.d;
    t.inv = mk3(1.f / rd.x, 1.f / rd.y, 1.f / rd.z);
    t.nx = signbit(t.inv.x); t.ny = signbit(t.inv.y); t.nz = signbit(t.inv.z);
    t.frame = cone_frame(t.env);
    t.res.n_tris = 0u; t.res.overflow = false; t.res.dist = WT_INF; t.res.front = false;
    if (force_rt || cone_is_ray(t.env)) { t.wstate = 0; g_start_ray(sc, g, sh, t, mkr(0.f, WT_INF), ctr); return; }
    t.dist = 0.f; t.seg = 0u; t.bd = max_ballistic_distance(lambda, 0u, 0.f);      // calculate_min_ballistic_distance == 0: the ray starts at the envelope's origin
    t.wstate = 1;
    g_start_ray(sc, g, sh, t, mkr(t.dist, fminf(WT_INF, t.dist + t.bd * 1.001f)), ctr);
}
// the current query has no work left: next query of this beam, or the beam's result (returns true, fills out)
WT_D bool g_query_done(const DScene& sc, const GLane& g, GShared& sh, GTrav& t, TravRec& out, Counters& ctr) {
    out.ox = t.env.o.x; out.oy = t.env.o.y; out.oz = t.env.o.z; out.pad_ = 0u; out.region_depth = 0.f;
    out.n_tris = 0u; out.cone_dist = WT_INF; out.ray_tuid = WTGPU_INVALID_IDX; out.ray_dist = WT_INF; out.bx = out.by = -1.f; out.flags = 0u;
    if (t.mode == 1) {
        // ads_t::intersect(ray, range) post-processing (traversal_common.hpp:93-110)
        bool hit = true;
        if (!isfinite(t.rec.dist) || t.rec.dist > t.qrange.mx) { t.rec.tuid = WTGPU_INVALID_IDX; t.rec.dist = WT_INF; hit = false; }
        if (t.edge_query && t.wstate == 1 && hit) {
            // plt_path only: a ballistic hit of a finite beam is followed by a cone query around the hit, +-zdist/2, z scale 1, whose triangles
            // are only used to collect edges (plt_path_detail.hpp:656-660).  The ray record is kept in t.rec.
            const float zd = cone_axes(t.env, t.rec.dist).x * kMajorToZ;
            t.wstate = 3; t.zs = 1.f;
            g_start_cone(sc, g, sh, t, mkr(t.rec.dist - zd / 2.f, t.rec.dist + zd / 2.f), ctr);
            return false;
        }
        if (t.wstate == 0 || hit) {
            out.flags = TR_BALLISTIC | (hit ? 0u : TR_EMPTY) | (t.rec.front ? TR_RAY_FRONT : 0u);
            out.ray_tuid = t.rec.tuid; out.ray_dist = t.rec.dist; out.bx = t.rec.bx; out.by = t.rec.by;
            return true;
        }
        t.dist += t.bd;
        if (t.bd == WT_INF || t.dist >= WT_INF) { out.flags = TR_BALLISTIC | TR_EMPTY; return true; }
        t.min_prog = cone_axes(t.env, t.dist).x / 2.f;
        t.wstate = 2;
        g_start_cone(sc, g, sh, t, mkr(t.dist, WT_INF), ctr);
        return false;
    }
    if (t.wstate == 3) {        // edge query done: the ballistic hit, plus the triangles around it
        out.flags = TR_BALLISTIC | (t.rec.front ? TR_RAY_FRONT : 0u) | (t.res.overflow ? TR_OVERFLOW : 0u);
        out.ray_tuid = t.rec.tuid; out.ray_dist = t.rec.dist; out.bx = t.rec.bx; out.by = t.rec.by;
        out.n_tris = t.res.n_tris;
        return true;
    }
    const bool cempty = t.res.n_tris == 0u;
    if (cempty || t.res.dist - t.dist >= t.min_prog) {
        out.flags = (cempty ? TR_EMPTY : 0u) | (t.res.front ? TR_CONE_FRONT : 0u) | (t.res.overflow ? TR_OVERFLOW : 0u);
        out.cone_dist = t.res.dist; out.n_tris = t.res.n_tris;
        out.region_depth = cempty ? 0.f : kMajorToZ * cone_axes(t.env, t.res.dist).x;
        // the ray record of the last (missed) ballistic segment stays as traverse() leaves it
        out.ray_tuid = t.rec.tuid; out.ray_dist = t.rec.dist; out.bx = t.rec.bx; out.by = t.rec.by; if (t.rec.front) out.flags |= TR_RAY_FRONT;
        return true;
    }
    ++t.seg;
    t.bd = max_ballistic_distance(t.lambda, t.seg, 0.f);
    t.wstate = 1;
    g_start_ray(sc, g, sh, t, mkr(t.dist, fminf(WT_INF, t.dist + t.bd * 1.001f)), ctr);
    return false;
}

// The driver loop.  fetch(i, env, prev, lambda, tris_out) loads item i (group-uniformly) and names its triangle-list row; emit(i, rec, g) stores its result.
// (A fully warp-uniform variant -- one pop per group per iteration, node steps of all groups in the same instructions -- was measured 8 %
// SLOWER on the etoile-like scene and on double_slits, and again 6-9 % slower after the code-size work removed the instruction-fetch
// stall (profiles/r01s3_phases.txt, session V): a group that reaches its leaf early idles through the others' node steps.)
template <class Fetch, class Emit>
WT_D void g_traverse_all(const DScene& sc, int n_items, int* cursor, GShared* shm, bool force_rt, bool edge_query, Counters& ctr, TravSave* big_save, int* n_big, uint32_t big_tested, Fetch&& fetch, Emit&& emit) {
    GLane g; g.gl = threadIdx.x & 7u; g.gshift = (threadIdx.x & 31u) & 24u; g.gmask = 0xffu << g.gshift;
    GShared& sh = shm[threadIdx.x / kGW];
    GTrav t; t.mode = 0; t.s = 0;
    bool have = false, done = false; int item = 0;
    for (;;) {
        // phase 1: bookkeeping and node steps, until a leaf is on top of this group's stack
        uint32_t lt0 = 0u, lcnt = 0u; bool leaf = false;
        for (;;) {
            if (!have) {
                if (done) break;
                int i = 0;
                if (g.gl == 0u) i = atomicAdd(cursor, 1);
                i = g_shfl(g, i, 0);
                if (i >= n_items) { done = true; break; }
                item = i;
                Cone env; Geo prev; float lambda;
                fetch(item, env, prev, lambda, t.tw);
                g_begin(sc, g, sh, t, env, prev, lambda, force_rt, edge_query, ctr);
                have = true;
            }
            if (t.s == 0) {
                TravRec out;
                if (g_query_done(sc, g, sh, t, out, ctr)) { emit(item, out, g); have = false; __syncwarp(g.gmask); }
                continue;
            }
            const int32_t top = sh.ptr[t.s - 1];
            if (top < 0) { const wtgpu_leaf lf = sc.leaves[-top - 1]; lt0 = lf.tris_ptr; lcnt = lf.count; leaf = true; --t.s; break; }
            if (t.mode == 1) {      // ray_traversal_treat_node_as_leaf_if_triangle_count_lt (bvh8w.cpp:29)
                const uint2 tr = __ldg(reinterpret_cast<const uint2*>(&sc.nodes[top - 1].tris_start));
                if (tr.y <= 16u) { if (g.gl == 0u) ctr.nodes++; lt0 = tr.x; lcnt = tr.y; leaf = true; --t.s; break; }
            }
            --t.s;
            g_node_step(sc, g, sh, t, top, ctr);
        }
        if (__all_sync(0xffffffffu, done && !have)) break;
        // phase 2: the groups of the warp do their leaf steps together
        if (leaf) {
            g_leaf_step(sc, g, sh, t, lt0, lcnt, ctr);
            // a cone query over many triangles: one group would walk it for milliseconds while the rest of the launch waits -- the beam is
            // handed, state and stack, to the team traversal (ctrav.cuh), which continues it from here (same decisions, same list)
            if (big_save && t.mode == 2 && t.s > 0 && t.qtested > big_tested) {
                int pos = 0;
                if (g.gl == 0u) pos = atomicAdd(n_big, 1);
                pos = g_shfl(g, pos, 0);
                TravSave& sv = big_save[pos];
                if (g.gl == 0u) { sv.t = t; sv.item = item; }
                for (int i = (int)g.gl; i < t.s; i += kGW) { sv.tmin[i] = sh.tmin[i]; sv.ptr[i] = sh.ptr[i]; }
                have = false; t.s = 0; t.mode = 0;
                __syncwarp(g.gmask);
            }
        }
    }
}

} // namespace wt
