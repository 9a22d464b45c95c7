// wavefront.cu -- the B200-native wavefront plt_path integrator and its C-ABI (include/wtgpu.h).
//
// Replaces the per-pixel recursion  integrator_t::integrate -> plt_path::random_walk
//   (/root/reference/src/integrator/plt_path.cpp:40-51, include/wt/integrator/plt_path/plt_path_detail.hpp:542-828)
// driven by scene_renderer_t::render's job loop (/root/reference/src/scene/render.cpp:381-579) with a pool of
// paths advanced one vertex per iteration:
//     k_generate   refill dead slots with new samples (emitter/spectrum/sensor sampling: plt_path_detail.hpp:764-828)
//     k_traverse   traverse() ballistic/diffusive state machine over the 8-wide BVH (integrator/traversal.hpp:94-172),
//                  primary-triangle pick (plt_path_detail.hpp:253-276), edge collection -> compact hit record + sort key
//     k_hist/scan/scatter   counting sort of live paths by material key (branch-coherent shading)
//     k_shade      pending UTD evaluation, NEE / emission / sensing, interaction sampling, beam transform, RR, film splats
// Path state lives in HBM as structure-of-arrays of 16-B chunks (chunk c of slot s at base[c*pool + s]): every warp-level
// state access is a run of fully coalesced 512-B transactions.
#include "dfsd.cuh"
#include "sobol_tables.h"
#include "../../include/wthost.h"      // (types only: wtgpu_debug_sizeof reports the layout of wthost_mesh_desc)

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>
#include <cmath>
#include <limits>

using namespace wt;

// ================================================================================================ state
constexpr unsigned kBlockT = 128u;

enum : uint32_t { F_SAMPLED_FSD = 1u, F_HAS_FSD = 2u, F_DPD_DISC = 4u };

struct alignas(16) PathCore {
    Beam beam;
    Geo prev_geo;
    uint32_t flags;
    float dpd_v, throughput, rspd;
    uint32_t depth, pixel, sample, rng_d;
    float ox, oy;
    float L[4];
};
struct alignas(16) PathFsd {     // what a path keeps of its previous vertex's aperture: the edge ids are in its row of RenderArgs::ap_edges
    Beam prev_beam;
    ApHead ap; uint32_t n;
};
enum : uint32_t { H_EMPTY = 1u, H_BALLISTIC = 2u, H_PRIMARY = 4u, H_FRONT = 8u, H_OVERFLOW = 16u };
struct alignas(16) HitRec {
    uint32_t flags, primary;
    float pdist, bx, by, d2i, region_depth;
    V3 origin;
    uint32_t n_edges;           // edge ids: the path's row of RenderArgs::hit_edges
    float flux;                 // plt_bdpt: Gaussian power over the clipped triangles (no primary hit)
};
static_assert(sizeof(HitRec) == 48, "HitRec is three 16-B chunks");
template <class T> __host__ __device__ constexpr int chunks_of() { return (int)(sizeof(T) / 16); }
template <class T> WT_D void soa_load(T& v, const float4* __restrict__ base, uint32_t pool, uint32_t slot) {
    float4* p = reinterpret_cast<float4*>(&v);
#pragma unroll
    for (int c = 0; c < chunks_of<T>(); ++c) p[c] = base[(size_t)c * pool + slot];
}
template <class T> WT_D void soa_store(const T& v, float4* __restrict__ base, uint32_t pool, uint32_t slot) {
    const float4* p = reinterpret_cast<const float4*>(&v);
#pragma unroll
    for (int c = 0; c < chunks_of<T>(); ++c) base[(size_t)c * pool + slot] = p[c];
}

WT_D void hit_store(const HitRec& v, float4* __restrict__ base, uint32_t pool, uint32_t slot) { soa_store(v, base, pool, slot); }
WT_D void hit_load(HitRec& v, const float4* __restrict__ base, uint32_t pool, uint32_t slot) { soa_load(v, base, pool, slot); }

struct DevCounters {
    unsigned long long samples, segments, ray_casts, cone_casts, shadow_casts, nodes, tris, edges, surface, fsd, null_, splats, overflow, shade_nodes, shade_tris, shaded;
    unsigned int next_sample_lo; unsigned int pad0;
    unsigned long long next_sample;     // samples handed out
    int live;                           // paths alive
    int n_trav;                         // entries of trav_list (paths to traverse this iteration)
    int n_sorted;                       // live paths in `order`
    int n_pairs[5];                     // plt_bdpt: (s,t) strategies queued this iteration, per strategy class
    unsigned long long strategies[5], walker_steps;
    int trav_head;                      // work-fetch cursor of the group traversal
    int n_fsd_list[3], fsd_head;        // plt_bdpt: walkers waiting for a Fraunhofer sample (three rotating lists); work-fetch cursor
    // capacity growth (dtrav.cuh Caps): the longest list a too-short row was asked to hold (0: every list fitted)
    unsigned int need_spill, need_edges, need_seg, need_ap, need_verts;
    unsigned int spill_head;            // bump cursor of the triangle-list arena (dtrav.cuh TriList); reset every iteration
    int n_big, big_head, n_big_res, big_res_head;      // beams handed to the team traversal (ctrav.cuh) / lists handed to the warp-per-list resolve kernels this iteration; their work cursors
    int n_huge, huge_head;                             // beams the warp teams handed on to the block teams
    int n_flux_items, flux_item_head, n_flux_tasks, flux_task_head;   // plt_bdpt: lists queued for the flat Gaussian-power kernels, their 32-entry chunk tasks; work cursors
    unsigned int flux_scratch_head;                    // bump cursor of the piece-value scratch
    int n_closest_tasks, closest_task_head;            // plt_bdpt: 256-entry tasks of the long lists' closest-triangle search; work cursor
    int n_quad_tasks, quad_task_head;                  // plt_bdpt: long quadrature pieces queued by the flat Gaussian-power kernel; work cursor
    unsigned long long stack_drops;
    unsigned long long dbg[32];         // -DWT_TEAM_DEBUG: team-traversal diagnostics (ctrav.cuh), printed by wtgpu_render when WT_DEBUG_TEAM is set
};
WT_D void need_max(unsigned int* p, uint32_t v) { if (v > *reinterpret_cast<volatile unsigned int*>(p)) atomicMax(p, v); }

namespace wt { struct TravRec; struct TravSave; struct TravTiers { uint32_t big_tested, huge_tested, restart_div, init_div; }; }
struct RenderArgs {
    DScene sc;
    wt::TravRec* trav_rec; uint32_t* trav_tris;      // results of traverse(), per slot; trav_tris: the first kTriRow triangle ids of the slot's cone-query list (dtrav.cuh TriList)
    wt::TravSave* big_save; wt::TravSave* huge_save;  // beams handed from the group traversal to the warp teams, and from those to the block teams (gtrav.cuh TravSave)
    wt::TravTiers tiers;                              // the hand-over thresholds (triangles tested by the current cone query) and team batch sizes
    uint32_t* big_res_list;                           // items (indices into trav_list) handed to the warp-per-list resolve kernels
    uint2* closest_tasks; unsigned long long* closest_best;     // plt_bdpt: (list, chunk) tasks of the flat closest-triangle search; its result per walker (key of the winning entry)
    uint2* flux_items; uint2* flux_tasks; float4* flux_scratch; uint32_t flux_cap; float4* quad_tasks; uint32_t quad_cap;     // plt_bdpt: (item, scratch base) per queued list; (list, chunk) tasks; piece values; scratch entries
    uint32_t* edge_bits;                              // scratch bitmaps (one bit per edge of the scene) of the warp-per-beam resolve kernel, one per resident warp
    uint32_t* hit_edges;                              // sc.cap.edges edge ids per slot: the edges around the vertex (HitRec::n_edges of them)
    uint32_t* ap_edges; uint32_t it_parity;           // plt_path: 2 x pool rows of sc.cap.edges: UTD aperture edge lists (this iteration's row set: it_parity)
    float4* core; float4* fsd; float4* hit;
    uint32_t* alive; uint32_t* keys; uint32_t* order; uint32_t* key_count; uint32_t* key_cursor; uint32_t* trav_list;
    DevCounters* ctr;
    float* film_block; float* film_light;
    uint32_t pool, n_keys;
    uint32_t seed_lo, seed_hi;
    uint32_t tile_x0, tile_y0, tile_w, tile_h;
    uint32_t sample_begin, n_samples;
    unsigned long long total;
    uint32_t part, n_parts;     // this sub-pool renders the samples whose linear id is part (mod n_parts)
};

WT_D void flush_counters(DevCounters* g, const Counters& c, bool shade = false) {
    const unsigned m = __activemask();
    const unsigned n = __reduce_add_sync(m, c.nodes), t = __reduce_add_sync(m, c.tris), r = __reduce_add_sync(m, c.ray_casts),
                   cc = __reduce_add_sync(m, c.cone_casts), s = __reduce_add_sync(m, c.shadow_casts);
    if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) {
        if (n) atomicAdd(shade ? &g->shade_nodes : &g->nodes, (unsigned long long)n);
        if (t) atomicAdd(shade ? &g->shade_tris : &g->tris, (unsigned long long)t);
        if (r) atomicAdd(&g->ray_casts, (unsigned long long)r);
        if (cc) atomicAdd(&g->cone_casts, (unsigned long long)cc);
        if (s) atomicAdd(&g->shadow_casts, (unsigned long long)s);
    }
    if (c.stack_drops) atomicAdd(&g->stack_drops, (unsigned long long)c.stack_drops);
}
WT_D void count1(unsigned long long* p, bool pred) {
    const unsigned m = __activemask();
    const unsigned n = __reduce_add_sync(m, pred ? 1u : 0u);
    if (n && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(p, (unsigned long long)n);
}

// warp-aggregated append of `slot` to a list (one atomic per warp)
WT_D void list_append(uint32_t* list, int* counter, bool pred, uint32_t slot) {
    const unsigned m = __ballot_sync(__activemask(), pred);
    if (!pred) return;
    const int lane = threadIdx.x & 31, leader = __ffs(m) - 1;
    int base = 0;
    if (lane == leader) base = atomicAdd(counter, __popc(m));
    base = __shfl_sync(m, base, leader);
    list[base + __popc(m & ((1u << lane) - 1u))] = slot;
}

// ================================================================================================ generate
__global__ void __launch_bounds__(128) k_generate(const RenderArgs a) {
    const uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
    bool gen = false;
    if (slot < a.pool && a.alive[slot] == 0u) {
        const unsigned long long id = atomicAdd(&a.ctr->next_sample, 1ull) * a.n_parts + a.part;
        if (id < a.total) {
            gen = true;
            const DScene& sc = a.sc;
            const uint64_t npix = (uint64_t)a.tile_w * a.tile_h;
            const uint32_t pi = (uint32_t)(id % npix), si = (uint32_t)(id / npix);
            const uint32_t ex = a.tile_x0 + pi % a.tile_w, ey = a.tile_y0 + pi / a.tile_w;
            PathCore pc;
            pc.pixel = ey * sc.sensor.width + ex; pc.sample = a.sample_begin + si;
            Sampler smp; smp.k0 = a.seed_lo; smp.k1 = a.seed_hi; smp.pixel = pc.pixel; smp.sample = pc.sample; smp.d = 0; smp.stream = sc.scene_stream;
            // integrate_backward / integrate_forward preamble (plt_path_detail.hpp:764-828)
            const int32_t em = sample_emitter(sc, smp);
            const KSample ks = sample_wavenumber(sc, em, smp);
            const float k = ks.k;
            if (sc.integrator.direction == WTGPU_DIRECTION_BACKWARD) {
                pc.rspd = ks.wpd.disc ? 1.f / ks.wpd.v : 1.f / sum_spectral_pdf(sc, k);
                const SensorSample ss = sensor_sample(sc, smp, ex, ey, k);
                pc.beam = ss.beam; pc.ox = ss.el.ox; pc.oy = ss.el.oy;
            } else {
                const EmitterSample es = emitter_sample(sc, em, smp, k);
                pc.rspd = 1.f / sum_spectral_pdf(sc, k);
                pc.beam = es.beam; pc.ox = pc.oy = 0.f;
            }
            pc.prev_geo = geo_point(pc.beam.env.o);
            pc.flags = F_DPD_DISC; pc.dpd_v = 0.f; pc.throughput = 1.f; pc.depth = 1u; pc.rng_d = smp.d;
            pc.L[0] = pc.L[1] = pc.L[2] = pc.L[3] = 0.f;
            soa_store(pc, a.core, a.pool, slot);
            a.alive[slot] = 1u;
        }
    }
    list_append(a.trav_list, &a.ctr->n_trav, gen, slot);
    const unsigned m = __activemask();
    const unsigned n = __reduce_add_sync(m, gen ? 1u : 0u);
    if (n && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) { atomicAdd(&a.ctr->live, (int)n); atomicAdd(&a.ctr->samples, (unsigned long long)n); }
}

// ================================================================================================ traverse
// integrator::traverse (integrator/traversal.hpp:28-57, 94-172, 276-286)
struct TravOut { bool empty, ballistic; RayHit ray; ConeResult cone; float region_depth; V3 origin; };

WT_D float min_ballistic_distance(const Cone& env, V3 ro) {
    if (!veq(ro, env.o)) {
        const V3 rl = to_local(cone_frame(env), ro - env.o) * mk3(1.f, env.e, 1.f);
        const float d = (length(mk2(rl.x, rl.y)) - env.x0) / env.ta - rl.z;
        return max3f(0.f, -rl.z, d);
    }
    return 0.f;
}
WT_D float max_ballistic_distance(float lambda, uint32_t seg, float mbd) {
    const unsigned long long B = min(1ull << 16, 8ull << (2u * min(seg, 16u) + 1u));
    return seg >= 16u ? WT_INF : mbd * 1.05f + lambda * (float)B;
}
WT_D void traverse(const DScene& sc, Cone env, const Geo& prev, float lambda, bool force_rt, TriWriter& tw, TravOut& out, Counters& ctr) {
    env.o = offseted_ray_origin(sc, prev, env.o, env.d);
    out.origin = env.o; out.region_depth = 0.f;
    out.cone.n_tris = 0; out.cone.overflow = false; out.cone.dist = WT_INF; out.cone.front = false;
    if (force_rt || cone_is_ray(env)) {
        out.ballistic = true;
        out.empty = !intersect_ray(sc, env.o, env.d, mkr(0.f, WT_INF), out.ray, ctr);
        return;
    }
    const float mbd = min_ballistic_distance(env, env.o);
    float dist = 0.f;
    for (uint32_t seg = 0;; ++seg) {
        const float bd = max_ballistic_distance(lambda, seg, mbd);
        if (intersect_ray(sc, env.o, env.d, mkr(dist, fminf(WT_INF, dist + bd * 1.001f)), out.ray, ctr)) { out.ballistic = true; out.empty = false; return; }
        dist += bd;
        if (bd == WT_INF || dist >= WT_INF) { out.ballistic = true; out.empty = true; return; }
        const float min_prog = cone_axes(env, dist).x / 2.f;
        cone_traverse(sc, env, mkr(dist, WT_INF), kMajorToZ, tw, out.cone, ctr);
        const bool cempty = out.cone.n_tris == 0u;
        if (cempty || out.cone.dist - dist >= min_prog) {
            out.ballistic = false; out.empty = cempty;
            out.region_depth = cempty ? 0.f : kMajorToZ * cone_axes(env, out.cone.dist).x;
            return;
        }
    }
}

WT_D void reset_iteration_lists(DevCounters* c) {        // the triangle-list arena and the hand-over lists live for one iteration
    need_max(&c->need_spill, c->spill_head);
    c->spill_head = 0u; c->n_big = 0; c->big_head = 0; c->n_big_res = 0; c->big_res_head = 0; c->n_huge = 0; c->huge_head = 0; c->n_flux_items = 0; c->flux_item_head = 0; c->n_flux_tasks = 0; c->flux_task_head = 0; c->flux_scratch_head = 0u; c->n_closest_tasks = 0; c->closest_task_head = 0; c->n_quad_tasks = 0; c->quad_task_head = 0;
}
// the triangle-list writer / reader of path `slot` (rows of kTriRow entries + the slot's extent table)
WT_D TriWriter tri_writer(const DScene& sc, uint32_t* trav_tris, uint32_t slot) { TriWriter w; w.row = trav_tris + (size_t)slot * kTriRow; w.ext = sc.spill_ext + (size_t)slot * kTriExt; tw_begin(w); return w; }
WT_D TriList tri_list(const DScene& sc, const uint32_t* trav_tris, uint32_t slot, uint32_t n, bool overflow) {
    TriList l; l.row = trav_tris + (size_t)slot * kTriRow; l.spill = sc.spill; l.ext = sc.spill_ext + (size_t)slot * kTriExt; l.n = overflow ? min(n, kTriRow) : n; return l;
}

// find_closest_triangle (plt_path_detail.hpp:253-276 / plt_bdpt_detail.hpp:362-390): the first triangle of the list, in list order, with the smallest hit distance
WT_D void find_closest(const DScene& sc, const TriList& tl, V3 origin, V3 dir, Range zr, uint32_t& primary, float& pdist, float& bx, float& by) {
    for (uint32_t i = 0; i < tl.n; ++i) {
        const uint32_t tu = tri_at(tl, i);
        const Tri3 t = load_tri(sc, tu);
        const float tol = cone_intersection_tolerance(origin, t.a, t.b, t.c);
        const RayTri rt = intersect_ray_tri(origin, dir, t.a, t.b, t.c, mkr(zr.mn - tol, zr.mx + tol));
        if (rt.hit && rt.dist < pdist) { primary = tu; pdist = rt.dist; bx = rt.bx; by = rt.by; }
    }
}
// the same by a whole warp (all lanes call with the same arguments and get the same result): lane l takes entries l, l + 32, ...; the
// (distance, list index) minimum over the lanes is the sequential loop's pick
WT_D void w_find_closest(const DScene& sc, const TriList& tl, V3 origin, V3 dir, Range zr, uint32_t& primary, float& pdist, float& bx, float& by) {
    const unsigned lane = threadIdx.x & 31u, FULL = 0xffffffffu;
    float bd = WT_INF, bbx = -1.f, bby = -1.f; uint32_t bi = 0xffffffffu, btu = WTGPU_INVALID_IDX;
    for (uint32_t i = lane; i < tl.n; i += 32u) {
        const uint32_t tu = tri_at(tl, i);
        const Tri3 t = load_tri(sc, tu);
        const float tol = cone_intersection_tolerance(origin, t.a, t.b, t.c);
        const RayTri rt = intersect_ray_tri(origin, dir, t.a, t.b, t.c, mkr(zr.mn - tol, zr.mx + tol));
        if (rt.hit && rt.dist < pdist && rt.dist < bd) { bd = rt.dist; bi = i; btu = tu; bbx = rt.bx; bby = rt.by; }
    }
    float rd = bd; uint32_t ri = bi;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const float od = __shfl_xor_sync(FULL, rd, o); const uint32_t oi = __shfl_xor_sync(FULL, ri, o);
        if (oi != 0xffffffffu && (ri == 0xffffffffu || od < rd || (od == rd && oi < ri))) { rd = od; ri = oi; }
    }
    if (ri == 0xffffffffu) return;
    const int src = __ffs(__ballot_sync(FULL, bi == ri)) - 1;
    primary = __shfl_sync(FULL, btu, src); pdist = rd; bx = __shfl_sync(FULL, bbx, src); by = __shfl_sync(FULL, bby, src);
}
// edges of a triangle list, deduplicated and ascending (collect_edges), by a whole warp through a scratch bitmap with one bit per edge of the
// scene (all zero on entry and on return).  Returns the number stored; `need` = the number there are.
WT_D uint32_t w_collect_edges(const DScene& sc, const TriList& tl, uint32_t* edges, uint32_t max_edges, uint32_t* bits, bool& overflow, uint32_t& need) {
    const unsigned lane = threadIdx.x & 31u, FULL = 0xffffffffu;
    uint32_t lo = 0xffffffffu, hi = 0u;
    for (uint32_t i = lane; i < tl.n; i += 32u) {
        const wtgpu_tri_meta m = sc.tri_meta[tri_at(tl, i)];
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const uint32_t e = k == 0 ? m.edge_ab : k == 1 ? m.edge_bc : m.edge_ca;
            if (e == WTGPU_INVALID_IDX) continue;
            atomicOr(bits + (e >> 5), 1u << (e & 31u)); lo = min(lo, e >> 5); hi = max(hi, e >> 5);
        }
    }
    lo = __reduce_min_sync(FULL, lo); hi = __reduce_max_sync(FULL, hi);
    __threadfence_block(); __syncwarp();
    uint32_t count = 0u;
    if (lo != 0xffffffffu)
        for (uint32_t w0 = lo; w0 <= hi; w0 += 32u) {
            const uint32_t w = w0 + lane;
            uint32_t v = w <= hi ? __ldcg(bits + w) : 0u;
            if (v) bits[w] = 0u;
            const uint32_t c = (uint32_t)__popc(v);
            uint32_t incl = c;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += u; }
            uint32_t pos = count + incl - c;
            while (v) { const uint32_t b = (uint32_t)__ffs(v) - 1u; v &= v - 1u; if (pos < max_edges) edges[pos] = w * 32u + b; ++pos; }
            count += __shfl_sync(FULL, incl, 31);
        }
    __syncwarp();
    need = count;
    if (count > max_edges) { overflow = true; return max_edges; }
    return count;
}

#include "gtrav.cuh"
#include "ctrav.cuh"
#include "dbdpt.cuh"

__global__ void __launch_bounds__(128) k_traverse(const RenderArgs a) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    Counters ctr; counters_zero(ctr);
    bool seg = false, ovf = false;
    if (li < (uint32_t)a.ctr->n_trav) {
        const uint32_t slot = a.trav_list[li];
        const DScene& sc = a.sc;
        seg = true;
        PathCore pc; soa_load(pc, a.core, a.pool, slot);
        TriWriter tw = tri_writer(sc, a.trav_tris, slot);
        uint32_t* edges = a.hit_edges + (size_t)slot * sc.cap.edges;
        TravOut tr;
        const bool force_rt = sc.sensor.ray_trace_only != 0u;
        traverse(sc, pc.beam.env, pc.prev_geo, wavenum_to_wavelen(pc.beam.k), force_rt, tw, tr, ctr);
        HitRec h; h.flags = 0u; h.primary = WTGPU_INVALID_IDX; h.pdist = WT_INF; h.bx = h.by = -1.f; h.n_edges = 0u; h.flux = 0.f;
        h.origin = tr.origin; h.region_depth = tr.region_depth; h.d2i = 0.f;
        uint32_t key = a.n_keys - 1u;       // miss
        if (tr.empty) h.flags |= H_EMPTY;
        else {
            const bool is_ballistic = tr.ballistic || cone_is_ray(pc.beam.env);
            Cone env = pc.beam.env; env.o = tr.origin;
            const V3 dir = env.d;
            if (tr.ballistic) {
                h.flags |= H_BALLISTIC | H_PRIMARY | (tr.ray.front ? H_FRONT : 0u);
                h.primary = tr.ray.tuid; h.pdist = tr.ray.dist; h.bx = tr.ray.bx; h.by = tr.ray.by; h.d2i = tr.ray.dist;
            } else {
                h.d2i = tr.cone.dist;
                if (tr.cone.front) h.flags |= H_FRONT;
                if (tr.cone.overflow) { h.flags |= H_OVERFLOW; ovf = true; }
                const TriList tl = tri_list(sc, a.trav_tris, slot, tr.cone.n_tris, tr.cone.overflow);
                // find_closest_triangle (plt_path_detail.hpp:253-276)
                const Range zr = mkr(h.d2i, h.d2i + h.region_depth);
                for (uint32_t i = 0; i < tl.n; ++i) {
                    const uint32_t tu = tri_at(tl, i);
                    const Tri3 t = load_tri(sc, tu);
                    const float tol = cone_intersection_tolerance(tr.origin, t.a, t.b, t.c);
                    const RayTri rt = intersect_ray_tri(tr.origin, dir, t.a, t.b, t.c, mkr(zr.mn - tol, zr.mx + tol));
                    if (rt.hit && rt.dist < h.pdist) { h.primary = tu; h.pdist = rt.dist; h.bx = rt.bx; h.by = rt.by; }
                }
                if (h.primary != WTGPU_INVALID_IDX) h.flags |= H_PRIMARY;
                if (sc.integrator.fsd) { bool eo = false; h.n_edges = collect_edges(sc, tl, edges, sc.cap.edges, eo); if (eo) { h.flags |= H_OVERFLOW; ovf = true; need_max(&a.ctr->need_edges, 3u * tl.n); } }
            }
            // ballistic hit with a finite beam: collect the edges around the hit (plt_path_detail.hpp:656-660)
            if (is_ballistic && !cone_is_ray(pc.beam.env) && !force_rt) {
                const float zd = cone_axes(env, h.d2i).x * kMajorToZ;
                ConeResult cr;
                cone_traverse(sc, env, mkr(h.d2i - zd / 2.f, h.d2i + zd / 2.f), 1.f, tw, cr, ctr);
                const TriList tl = tri_list(sc, a.trav_tris, slot, cr.n_tris, cr.overflow);
                bool eo = false;
                h.n_edges = collect_edges(sc, tl, edges, sc.cap.edges, eo);
                if (eo) need_max(&a.ctr->need_edges, 3u * tl.n);
                if (eo || cr.overflow) { h.flags |= H_OVERFLOW; ovf = true; }
            }
            if (h.flags & H_PRIMARY) key = (uint32_t)sc.shapes[sc.tri_meta[h.primary].shape_idx].bsdf;
            else key = h.n_edges ? a.n_keys - 3u : a.n_keys - 2u;
        }
        hit_store(h, a.hit, a.pool, slot);
        a.keys[slot] = key;
    }
    flush_counters(a.ctr, ctr);
    count1(&a.ctr->segments, seg);
    count1(&a.ctr->overflow, ovf);
}

// traverse() for every path of the list, eight lanes per beam (gtrav.cuh), including the edge query that follows a ballistic hit of a
// finite beam; k_resolve then builds the hit record per thread.  Bit-identical to k_traverse (WTGPU_RENDER_THREAD_TRAVERSE selects that one).
#ifndef WT_GT_MINB
#define WT_GT_MINB 5      // 96 registers: 5 blocks/SM; measured against 4 (119 regs), 6 (80, spills) and 8 (64, spills): profiles/r01s3_variants.txt
#endif
__global__ void __launch_bounds__(128, WT_GT_MINB) k_gtraverse(const RenderArgs a) {
    __shared__ GShared shm[128 / kGW];
    Counters ctr; counters_zero(ctr);
    const DScene& sc = a.sc;
    g_traverse_all(sc, a.ctr->n_trav, &a.ctr->trav_head, shm, sc.sensor.ray_trace_only != 0u, true, ctr, a.big_save, &a.ctr->n_big, a.tiers.big_tested,
        [&](int i, Cone& env, Geo& prev, float& lambda, TriWriter& tw) {
            const uint32_t slot = a.trav_list[i];
            PathCore pc; soa_load(pc, a.core, a.pool, slot);
            env = pc.beam.env; prev = pc.prev_geo; lambda = wavenum_to_wavelen(pc.beam.k);
            tw = tri_writer(sc, a.trav_tris, slot);
        },
        [&](int i, const TravRec& r, const GLane& g) { if (g.gl == 0u) a.trav_rec[a.trav_list[i]] = r; });
    flush_counters(a.ctr, ctr);
}
// the beams k_gtraverse handed over in the middle of a large cone query: one warp team per beam (ctrav.cuh), then one block team for the largest
__global__ void __launch_bounds__(128, WT_WT_MINB) k_wtraverse(const RenderArgs a) {
    __shared__ TShared<32> shm[4];
    Counters ctr; counters_zero(ctr);
    t_traverse_all<32>(a.sc, a.ctr->n_big, a.big_save, &a.ctr->big_head, shm[threadIdx.x >> 5], ctr, a.huge_save, &a.ctr->n_huge, a.tiers, a.ctr->dbg,
        [&](int i, const TravRec& r, const GLane& g) { if (g.gl == 0u) a.trav_rec[a.trav_list[i]] = r; });
    flush_counters(a.ctr, ctr);
}
__global__ void __launch_bounds__(256, WT_CT_MINB) k_ctraverse(const RenderArgs a) {
    __shared__ TShared<256> shm;
    Counters ctr; counters_zero(ctr);
    t_traverse_all<256>(a.sc, a.ctr->n_huge, a.huge_save, &a.ctr->huge_head, shm, ctr, nullptr, nullptr, a.tiers, a.ctr->dbg + 8,
        [&](int i, const TravRec& r, const GLane& g) { if (g.gl == 0u) a.trav_rec[a.trav_list[i]] = r; });
    flush_counters(a.ctr, ctr);
}
// primary-triangle pick (plt_path_detail.hpp:253-276), edge collection, sort key: the per-thread tail of k_traverse.
// WARP = false: one thread per path (lists of more than kBigQuery triangles are deferred to big_res_list); WARP = true: one warp per deferred path.
template <bool WARP> WT_D void path_resolve(const RenderArgs& a, uint32_t li, uint32_t* edge_bits, bool& seg, bool& ovf) {
    const uint32_t slot = a.trav_list[li];
    const DScene& sc = a.sc;
    const unsigned lane = threadIdx.x & 31u;
    const TravRec r = a.trav_rec[slot];
    const bool tr_ovf = (r.flags & TR_OVERFLOW) != 0u;
    const TriList tl = tri_list(sc, a.trav_tris, slot, r.n_tris, tr_ovf);
    if (!WARP && tl.n > kBigQuery && !(r.flags & TR_EMPTY)) { a.big_res_list[atomicAdd(&a.ctr->n_big_res, 1)] = li; return; }
    seg = !WARP || lane == 0u;
    uint32_t* edges = a.hit_edges + (size_t)slot * sc.cap.edges;
    const V3 origin = mk3(r.ox, r.oy, r.oz);
    // the beam's mean direction: floats 3..5 of PathCore (beam.env = {o, d, ...}), i.e. chunk 0 .w and chunk 1 .xy
    static_assert(offsetof(PathCore, beam) == 0 && offsetof(Beam, env) == 0 && offsetof(Cone, d) == 12 && sizeof(V3) == 12, "PathCore layout");
    const float4 c0 = a.core[slot], c1 = a.core[(size_t)a.pool + slot];
    const V3 dir = mk3(c0.w, c1.x, c1.y);
    HitRec h; h.flags = 0u; h.primary = WTGPU_INVALID_IDX; h.pdist = WT_INF; h.bx = h.by = -1.f; h.n_edges = 0u; h.flux = 0.f;
    h.origin = origin; h.region_depth = r.region_depth; h.d2i = 0.f;
    uint32_t key = a.n_keys - 1u;       // miss
    bool o2 = false;
    if (r.flags & TR_EMPTY) h.flags |= H_EMPTY;
    else {
        if (tr_ovf) { h.flags |= H_OVERFLOW; o2 = true; }
        if (r.flags & TR_BALLISTIC) {
            h.flags |= H_BALLISTIC | H_PRIMARY | ((r.flags & TR_RAY_FRONT) ? H_FRONT : 0u);
            h.primary = r.ray_tuid; h.pdist = r.ray_dist; h.bx = r.bx; h.by = r.by; h.d2i = r.ray_dist;
        } else {
            h.d2i = r.cone_dist;
            if (r.flags & TR_CONE_FRONT) h.flags |= H_FRONT;
            const Range zr = mkr(h.d2i, h.d2i + h.region_depth);
            if (WARP) w_find_closest(sc, tl, origin, dir, zr, h.primary, h.pdist, h.bx, h.by);
            else find_closest(sc, tl, origin, dir, zr, h.primary, h.pdist, h.bx, h.by);
            if (h.primary != WTGPU_INVALID_IDX) h.flags |= H_PRIMARY;
        }
        // cone segment: edges of the returned triangles when FSD is on; ballistic hit: edges of the edge query's triangles (always, as k_traverse)
        if (((r.flags & TR_BALLISTIC) || sc.integrator.fsd) && tl.n) {
            bool eo = false; uint32_t need = 0u;
            if (WARP) h.n_edges = w_collect_edges(sc, tl, edges, sc.cap.edges, edge_bits, eo, need);
            else { h.n_edges = collect_edges(sc, tl, edges, sc.cap.edges, eo); need = 3u * tl.n; }
            if (eo) { h.flags |= H_OVERFLOW; o2 = true; if (!WARP || lane == 0u) need_max(&a.ctr->need_edges, need); }
        }
        if (h.flags & H_PRIMARY) key = (uint32_t)sc.shapes[sc.tri_meta[h.primary].shape_idx].bsdf;
        else key = h.n_edges ? a.n_keys - 3u : a.n_keys - 2u;
    }
    if (!WARP || lane == 0u) { hit_store(h, a.hit, a.pool, slot); a.keys[slot] = key; ovf = o2; }
}
__global__ void __launch_bounds__(128) k_resolve(const RenderArgs a) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    bool seg = false, ovf = false;
    if (li < (uint32_t)a.ctr->n_trav) path_resolve<false>(a, li, nullptr, seg, ovf);
    count1(&a.ctr->segments, seg);
    count1(&a.ctr->overflow, ovf);
}
// the paths k_resolve deferred: one warp per path (closest triangle by a warp-wide ordered minimum, edge set through a scratch bitmap)
__global__ void __launch_bounds__(128) k_resolve_big(const RenderArgs a, uint32_t bit_words) {
    const unsigned lane = threadIdx.x & 31u;
    uint32_t* bits = a.edge_bits + (size_t)(blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5)) * bit_words;
    unsigned long long n_seg = 0, n_ovf = 0;
    for (;;) {
        int i = 0;
        if (lane == 0u) i = atomicAdd(&a.ctr->big_res_head, 1);
        i = __shfl_sync(0xffffffffu, i, 0);
        if (i >= a.ctr->n_big_res) break;
        bool seg = false, ovf = false;
        path_resolve<true>(a, a.big_res_list[i], bits, seg, ovf);
        n_seg += seg ? 1u : 0u; n_ovf += ovf ? 1u : 0u;
        __syncwarp();
    }
    if (lane == 0u) { if (n_seg) atomicAdd(&a.ctr->segments, n_seg); if (n_ovf) atomicAdd(&a.ctr->overflow, n_ovf); }
}

// ================================================================================================ material sort (counting sort)
__global__ void k_hist(const RenderArgs a) {
    extern __shared__ uint32_t sh[];
    for (uint32_t i = threadIdx.x; i < a.n_keys; i += blockDim.x) sh[i] = 0u;
    __syncthreads();
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li < (uint32_t)a.ctr->n_trav) atomicAdd(&sh[a.keys[a.trav_list[li]]], 1u);
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < a.n_keys; i += blockDim.x) if (sh[i]) atomicAdd(&a.key_count[i], sh[i]);
}
// exclusive scan of the key histogram by one block (keys = bsdf nodes + 3: a scene with hundreds of materials has thousands)
__global__ void __launch_bounds__(1024) k_scan(const RenderArgs a) {
    __shared__ uint32_t warp_tot[32];
    __shared__ uint32_t carry;
    const uint32_t lane = threadIdx.x & 31u, wrp = threadIdx.x >> 5;
    if (threadIdx.x == 0) carry = 0u;
    __syncthreads();
    for (uint32_t base = 0; base < a.n_keys; base += blockDim.x) {
        const uint32_t i = base + threadIdx.x;
        const uint32_t c = i < a.n_keys ? a.key_count[i] : 0u;
        uint32_t incl = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, incl, o); if ((int)lane >= o) incl += v; }
        if (lane == 31u) warp_tot[wrp] = incl;
        __syncthreads();
        if (wrp == 0u) {
            uint32_t w = lane < (blockDim.x >> 5) ? warp_tot[lane] : 0u, wi = w;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const uint32_t v = __shfl_up_sync(0xffffffffu, wi, o); if ((int)lane >= o) wi += v; }
            warp_tot[lane] = wi - w;        // exclusive prefix of the warp totals
        }
        __syncthreads();
        const uint32_t excl = carry + warp_tot[wrp] + incl - c;
        if (i < a.n_keys) { a.key_cursor[i] = excl; a.key_count[i] = 0u; }
        __syncthreads();
        if (threadIdx.x == blockDim.x - 1u) carry = excl + c;      // total so far
        __syncthreads();
    }
    if (threadIdx.x == 0) a.ctr->n_sorted = (int)carry;
}
__global__ void k_scatter(const RenderArgs a) {
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    const bool act = li < (uint32_t)a.ctr->n_trav;
    uint32_t slot = 0, key = 0;
    if (act) { slot = a.trav_list[li]; key = a.keys[slot]; }
    // one atomic per (warp, key): lanes with equal keys elect a leader
    const unsigned am = __ballot_sync(0xffffffffu, act);
    if (act) {
        const unsigned peers = __match_any_sync(am, key);
        const int lane = threadIdx.x & 31, leader = __ffs(peers) - 1;
        uint32_t base = 0;
        if (lane == leader) base = atomicAdd(&a.key_cursor[key], (uint32_t)__popc(peers));
        base = __shfl_sync(peers, base, leader);
        a.order[base + __popc(peers & ((1u << lane) - 1u))] = slot;
    }
}
__global__ void k_reset_trav(const RenderArgs a) { if (threadIdx.x == 0 && blockIdx.x == 0) { a.ctr->n_trav = 0; a.ctr->trav_head = 0; reset_iteration_lists(a.ctr); } }
__global__ void k_identity_order(const RenderArgs a) {     // WTGPU_RENDER_NO_SORT: live slots in slot order (compaction only)
    const uint32_t li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li < (uint32_t)a.ctr->n_trav) a.order[li] = a.trav_list[li];
    if (li == 0) a.ctr->n_sorted = a.ctr->n_trav;
}

// ================================================================================================ shade
WT_D float MIS(float p1, float p2) { if (p2 == 0.f) return 1.f; return p1 * p1 / (p1 * p1 + p2 * p2); }

// One thread per path; the two do_fsd evaluations of a vertex (pending UTD of the previous vertex, forward NEE through the new aperture) are
// served by the whole warp at convergent points (warp_do_fsd, dfsd.cuh), so the kernel body is written as flat phases.
__global__ void __launch_bounds__(128, 3) k_shade(const RenderArgs a) {
    __shared__ UtdShared utd_sh[4];
    UtdShared& ush = utd_sh[threadIdx.x >> 5];
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    const DScene& sc = a.sc;
    Counters ctr; counters_zero(ctr);
    uint32_t n_splat = 0, n_edges_fetched = 0, my_slot = 0; bool c_surface = false, c_fsd = false, c_null = false, died = false, survive = false;
    const bool act = i < (uint32_t)a.ctr->n_sorted;
    bool proc = false, alive = true;
    uint32_t slot = 0;
    PathCore pc; HitRec h; PathFsd pf; Aperture ap, apo; ap.n = 0u; ap.edges = nullptr; apo.n = 0u; apo.edges = nullptr;
    const uint32_t* hedges = nullptr;
    Sampler smp;
    bool has_new_fsd = false;
    Beam prev_beam_new;     // beam before this vertex's interaction (becomes prev_vert_beam)
    if (act) {
        slot = a.order[i];
        soa_load(pc, a.core, a.pool, slot);
        hit_load(h, a.hit, a.pool, slot);
        hedges = a.hit_edges + (size_t)slot * sc.cap.edges;
        ap.edges = a.ap_edges + ((size_t)a.it_parity * a.pool + slot) * sc.cap.edges;             // the aperture this vertex may build
        apo.edges = a.ap_edges + ((size_t)(a.it_parity ^ 1u) * a.pool + slot) * sc.cap.edges;     // the one the path carries from its previous vertex
        smp.k0 = a.seed_lo; smp.k1 = a.seed_hi; smp.pixel = pc.pixel; smp.sample = pc.sample; smp.d = pc.rng_d; smp.stream = 0u;
        if (h.flags & H_EMPTY) alive = false; else proc = true;
    }
    Beam& beam = pc.beam;
    const bool fwd = act ? beam.fwd : false;
    const uint32_t max_depth = sc.integrator.max_depth;
    const float k = proc ? beam.k : 0.f;
    const float d2i = proc ? h.d2i : 0.f;
    const V3 origin_wp = proc ? h.origin : mk3(0.f, 0.f, 0.f);
    const V3 dir = proc ? beam.env.d : mk3(0.f, 0.f, 1.f);
    const V3 interaction_wp = origin_wp + d2i * dir;

    // ---- evaluate fsd from the previous interaction (plt_path_detail.hpp:591-610)
    const bool need1 = proc && (pc.flags & F_HAS_FSD);
    if (need1) { soa_load(pf, a.fsd, a.pool, slot); apo.wp = pf.ap.wp; apo.fr = pf.ap.fr; apo.size = pf.ap.size; apo.wi = pf.ap.wi; apo.k = pf.ap.k; apo.n = pf.n; }
    const float f1 = warp_do_fsd(sc, ush, need1, pf.prev_beam.env, pc.prev_geo, interaction_wp, apo, k, ctr, n_edges_fetched);
    if (need1) {
        pc.flags &= ~F_HAS_FSD;
        if (pc.flags & F_SAMPLED_FSD) beam_mul(beam, f1);
        else {
            beam_transform_region(pf.prev_beam, origin_wp, length(origin_wp - pf.prev_beam.env.o), dir, f1);
            beam_add(beam, pf.prev_beam);
        }
    }

    // ---- surface record of the primary triangle (plt_path_detail.hpp:634-652)
    const bool has_primary = proc && (h.flags & H_PRIMARY) != 0u;
    const float region_end = has_primary ? h.pdist : d2i;
    Surface surf;
    int32_t bsdf = -1, emitter = -1;
    bool need2 = false; SensorDirect sd;
    if (proc) {
        const Frame beam_frame = cone_frame(beam.env);
        if (has_primary) {
            surf = make_surface(sc, h.primary, mk2(h.bx, h.by), origin_wp + h.pdist * dir);
            surf.fp = surface_footprint_static(beam, surf, d2i);
            const wtgpu_shape shp = sc.shapes[sc.tri_meta[h.primary].shape_idx];
            bsdf = shp.bsdf; emitter = shp.emitter;
        }

        // ---- construct the fsd aperture from the edges (plt_path_detail.hpp:663-679)
        if (h.n_edges) {
            ap.wp = interaction_wp; ap.fr = beam_frame; ap.size = beam_footprint(beam, d2i); ap.wi = -dir; ap.k = k; ap.n = 0u;
            for (uint32_t j = 0; j < h.n_edges; ++j) { Wedge w; ++n_edges_fetched; const uint32_t ed = hedges[j]; if (wedge_build(sc, ap, ed, w)) ap.edges[ap.n++] = ed; }
            has_new_fsd = ap.n > 0u;
            c_fsd = true;
        }

        // ---- NEE (plt_path_detail.hpp:350-424 backward, 468-510 forward)
        if (!fwd) {
            if (pc.depth < max_depth && has_primary && !bsdf_is_delta_only(sc, bsdf, k)) {
                const EmitterDirect ds = scene_sample_emitter_direct(sc, smp, surf.wp, k);
                if (beam_intensity(ds.beam) != 0.f) {
                    const V3 wiw = -dir, wow = -ds.beam.env.d;
                    const V3 wi = to_local(surf.shading, wiw), wo = to_local(surf.shading, wow);
                    const float wig = dot(wiw, surf.geo.n), wog = dot(wow, surf.geo.n);
                    if (!(wi.z * wig <= 0.f || wo.z * wog <= 0.f)) {
                        BsdfQuery q; q.k = k; q.fwd = false; q.lobes = 0xffffffffu;
                        const Mueller f = bsdf_f(sc, bsdf, wi, wo, q);
                        if (f.m[0] != 0.f) {
                            const Geo eg = ds.has_surface ? geo_surface(ds.sp, ds.stuid, true) : geo_point(ds.beam.env.o);
                            if (!shadow_between(sc, geo_surface(surf.wp, surf.tuid, true), eg, ctr)) {
                                Beam nb = beam;
                                beam_transform_surface(nb, surf, wow, f, 1.f);
                                const Stokes sL = integrate_beams(nb, ds.beam);
                                float mis = 1.f;
                                if (!ds.dpd.disc) mis = MIS(ds.dpd.v * ds.emitter_pdf, bsdf_pdf(sc, bsdf, wi, wo, q));
                                _Pragma("unroll") for (int c = 0; c < 4; ++c) pc.L[c] += sL.s[c] * mis;
                            }
                        }
                    }
                }
            }
        } else if (pc.depth < max_depth && has_new_fsd && sc.sensor.type == WTGPU_SENSOR_VIRTUAL_PLANE) {
            sd = sensor_sample_direct(sc, smp, interaction_wp, k);
            need2 = (sd.dpd.disc || sd.dpd.v != 0.f) && beam_intensity(sd.beam) > 0.f;
        }
    }
    const float f2 = warp_do_fsd(sc, ush, need2, beam.env, pc.prev_geo, sd.beam.env.o, ap, k, ctr, n_edges_fetched);
    if (need2 && f2 != 0.f) {
        Beam fb = beam;
        beam_transform_region(fb, interaction_wp, d2i, -sd.beam.env.d, f2);
        const Stokes sL = integrate_beams(sd.beam, fb);
        n_splat += film_splat(sc, a.film_block, a.film_light, true, sd.el, sL.s[0] * pc.rspd, k);
    }

    if (proc) {
        // ---- organic connections: emission (plt_path_detail.hpp:427-465) / sensing (513-540)
        if (!fwd) {
            if (has_primary && emitter >= 0) {
                const Stokes sL = emitter_Li(sc, emitter, beam, surf);
                float mis = 1.f;
                if (!(pc.flags & F_DPD_DISC)) {
                    const wtgpu_emitter E = sc.emitters[emitter];
                    const float ppd = E.type == WTGPU_EMITTER_AREA ? 1.f / sc.shapes[E.shape].surface_area : 0.f;
                    const float dn = dot(-dir, surf.geo.n);
                    const float rdn = dn != 0.f ? 1.f / fabsf(dn) : 0.f;
                    const float pd_nee = ppd * length2(beam.env.o - surf.wp) * rdn;
                    mis = MIS(pc.dpd_v, pd_nee * pdf_emitter(sc, emitter));
                }
                _Pragma("unroll") for (int c = 0; c < 4; ++c) pc.L[c] += sL.s[c] * mis;
            }
        } else {
            const float maxd = region_end - fmaxf(0.f, dot(dir, origin_wp - beam.env.o));
            Beam se; Element el;
            if (sensor_Si(sc, beam, mkr(0.f, maxd), se, el)) {
                const Stokes sL = integrate_beams(se, beam);
                n_splat += film_splat(sc, a.film_block, a.film_light, true, el, sL.s[0] * pc.rspd, k);
            }
        }

        // ---- interactions (plt_path_detail.hpp:156-237, 729-749)
        bool sampled_null = false;
        if (has_primary) {
            BsdfQuery q; q.k = k; q.fwd = fwd; q.lobes = 0xffffffffu;
            const V3 wiw = -dir;
            const V3 wi = to_local(surf.shading, wiw);
            const float wig = dot(wiw, surf.geo.n);
            if (wig * wi.z <= 0.f) alive = false;
            else {
                const BsdfSample bs = bsdf_sample(sc, bsdf, wi, q, smp);
                if (!bs.valid || bs.dpd.v == 0.f) alive = false;
                else {
                    const V3 wow = normalize(to_world(surf.shading, bs.wo));
                    c_surface = true;
                    if (dot(wow, surf.geo.n) * bs.wo.z <= 0.f) alive = false;
                    else {
                        pc.dpd_v = bs.dpd.v; pc.flags = (pc.flags & ~(F_DPD_DISC | F_SAMPLED_FSD)) | (bs.dpd.disc ? F_DPD_DISC : 0u);
                        pc.prev_geo = geo_surface(surf.wp, surf.tuid, true);
                        prev_beam_new = beam;
                        beam_transform_surface(beam, surf, wow, bs.M, 1.f);
                        pc.throughput *= 1.f * bs.M.m[0];
                        if (bs.eta.re != 1.f) pc.throughput /= sqrf(bs.eta.re);
                    }
                }
            }
        } else if (has_new_fsd) {
            V3 wo; float w;
            fsd_sample(sc, ap, pc.prev_geo.p, smp, wo, w);
            pc.dpd_v = 0.f; pc.flags = (pc.flags | F_DPD_DISC | F_SAMPLED_FSD);
            pc.prev_geo = geo_point(interaction_wp);
            prev_beam_new = beam;
            beam_transform_region(beam, interaction_wp, d2i, wo, w);
            pc.throughput *= w;
        } else {
            sampled_null = true; c_null = true;
            beam_transform_restart(beam, interaction_wp, d2i);
        }

        // ---- continue walk (plt_path_detail.hpp:123-142, 755-757)
        if (alive) {
            if (pc.depth >= max_depth) alive = false;
            else if (beam_intensity(beam) == 0.f) alive = false;
            else if (!sampled_null && sc.integrator.russian_roulette) {
                const float r = pc.throughput < 1.f ? fmaxf(pc.throughput, .5f) : 1.f;
                if (rnd(smp) <= r) { const float s = 1.f / r; beam_mul(beam, s); pc.throughput *= s; }
                else alive = false;
            }
            if (alive && !sampled_null) pc.depth++;
        }
    }

    if (act) {
        survive = alive;
        my_slot = slot;
        if (alive) {
            pc.rng_d = smp.d;
            if (has_new_fsd) {
                pc.flags |= F_HAS_FSD;
                pf.prev_beam = prev_beam_new; pf.ap.wp = ap.wp; pf.ap.fr = ap.fr; pf.ap.size = ap.size; pf.ap.wi = ap.wi; pf.ap.k = ap.k; pf.n = ap.n;
                soa_store(pf, a.fsd, a.pool, slot);
            }
            soa_store(pc, a.core, a.pool, slot);
        } else {
            died = true;
            if (!fwd) {     // splat_backward (plt_path_detail.hpp:799-800)
                Element el; el.ex = pc.pixel % sc.sensor.width; el.ey = pc.pixel / sc.sensor.width; el.ox = pc.ox; el.oy = pc.oy;
                n_splat += film_splat(sc, a.film_block, a.film_light, false, el, pc.L[0] * pc.rspd, pc.beam.k);
            }
            a.alive[slot] = 0u;
        }
    }
    list_append(a.trav_list, &a.ctr->n_trav, survive, my_slot);
    flush_counters(a.ctr, ctr, true);
    count1(&a.ctr->shaded, act);
    {
        const unsigned m = __activemask();
        const unsigned ns = __reduce_add_sync(m, n_splat), ne = __reduce_add_sync(m, n_edges_fetched), nd = __reduce_add_sync(m, died ? 1u : 0u);
        if ((threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) {
            if (ns) atomicAdd(&a.ctr->splats, (unsigned long long)ns);
            if (ne) atomicAdd(&a.ctr->edges, (unsigned long long)ne);
            if (nd) atomicAdd(&a.ctr->live, -(int)nd);
        }
    }
    count1(&a.ctr->surface, c_surface);
    count1(&a.ctr->fsd, c_fsd);
    count1(&a.ctr->null_, c_null);
}

// ================================================================================================ debug kernels
__global__ void k_debug_rays(const DScene sc, uint32_t n, const wtgpu_ray_query* q, wtgpu_ray_hit* out, uint32_t* shadow_out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Counters ctr; counters_zero(ctr);
    const V3 o = mk3(q[i].o), d = mk3(q[i].d);
    if (shadow_out) { shadow_out[i] = shadow_ray(sc, o, d, mkr(q[i].tmin, q[i].tmax), ctr) ? 1u : 0u; return; }
    RayHit h; intersect_ray(sc, o, d, mkr(q[i].tmin, q[i].tmax), h, ctr);
    out[i].tuid = h.tuid; out[i].dist = h.dist; out[i].bary[0] = h.bx; out[i].bary[1] = h.by; out[i].front_face = h.front ? 1u : 0u;
}
__global__ void k_debug_cones(const DScene sc, uint32_t n, const wtgpu_cone_query* q, wtgpu_cone_hit* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Counters ctr; counters_zero(ctr);
    const Cone c = mkcone(mk3(q[i].o), mk3(q[i].d), mk3(q[i].x), q[i].x0, q[i].tan_alpha, 1.f / q[i].e, q[i].e);
    static_assert(WTGPU_MAX_CONE_TRIS == kTriRow, "the debug entry point returns the row part of a triangle list");
    uint32_t tris[WTGPU_MAX_CONE_TRIS]; ConeResult r;
    TriWriter tw; tw.row = tris; tw.ext = nullptr;      // (no extents: a longer list only counts on)
    cone_traverse(sc, c, mkr(q[i].tmin, q[i].tmax), q[i].z_scale, tw, r, ctr);
    wtgpu_cone_hit& h = out[i];
    h.dist = r.dist; h.front_face = r.front ? 1u : 0u; h.n_tris = r.n_tris;
    const uint32_t nt = min(r.n_tris, (uint32_t)WTGPU_MAX_CONE_TRIS);
    for (uint32_t j = 0; j < nt; ++j) h.tris[j] = tris[j];
    bool eo = false; uint32_t edges[WTGPU_MAX_CONE_EDGES];
    TriList tl; tl.row = tris; tl.spill = nullptr; tl.ext = nullptr; tl.n = nt;
    h.n_edges = collect_edges(sc, tl, edges, WTGPU_MAX_CONE_EDGES, eo);
    for (uint32_t j = 0; j < h.n_edges; ++j) h.edges[j] = edges[j];
}
// sobolld: dimensions 0..46 of points g0 .. g0+n-1 (one thread per value)
__global__ void k_debug_sobol(uint32_t k0, uint32_t k1, unsigned long long g0, uint32_t n, uint32_t* num, float* val) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n * WTGPU_SOBOL_DIMS) return;
    const uint32_t u = sobol_numerator(k0, k1, g0 + i / WTGPU_SOBOL_DIMS, i % WTGPU_SOBOL_DIMS);
    num[i] = u; val[i] = sobol_value(u);
}
__global__ void k_debug_rng(uint32_t k0, uint32_t k1, uint32_t pixel, uint32_t sample, uint32_t n, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    Sampler s; s.k0 = k0; s.k1 = k1; s.pixel = pixel; s.sample = sample; s.d = i; s.stream = 0u;
    out[i] = rnd(s);
}

// pmath.h on the device, one thread per argument: fn 0 sin 1 cos 2 tan 3 exp 4 log 5 pow(x,y) 6 atan2(x,y) 7 acos 8 hypot(x,y) 9/10 Re/Im UTDF(x)
__global__ void k_debug_pmath(int fn, uint32_t n, const float* x, const float* y, float* out) {
    const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float a = x[i], b = y ? y[i] : 0.f;
    float r = 0.f;
    switch (fn) {
    case 0: r = pm::sinf(a); break; case 1: r = pm::cosf(a); break; case 2: r = pm::tanf(a); break; case 3: r = pm::expf(a); break;
    case 4: r = pm::logf(a); break; case 5: r = pm::powf(a, b); break; case 6: r = pm::atan2f(a, b); break; case 7: r = pm::acosf(a); break;
    case 8: r = pm::hypotf(a, b); break; case 9: r = UTDF(a).re; break; case 10: r = UTDF(a).im; break;
    }
    out[i] = r;
}

// ================================================================================================ host side
static thread_local std::string g_err;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { g_err = std::string(#x) + ": " + cudaGetErrorString(e_); return WTGPU_E_CUDA; } } while (0)

// Device block cache.  Path pools are GBs (a plt_bdpt pool of 2^18 sample slots is ~8 GB); a host that renders through
// scene create -> render -> destroy repeatedly (one call per sensor, per frame, per bench step) must not pay cudaMalloc/cudaFree of
// those every time.  Freed blocks >= 1 MiB are kept per (device, size) and handed back on the next request of the same size;
// wtgpu_trim() releases them.  Blocks carry no state: every user initialises what it reads.
#include <map>
#include <mutex>
#include <unordered_map>
namespace {
struct BlockCache {
    std::mutex m;
    std::multimap<std::pair<int, size_t>, void*> idle;
    std::unordered_map<void*, std::pair<int, size_t>> live;
    size_t idle_bytes = 0;
};
BlockCache g_blocks;
constexpr size_t kCacheMinBlock = 1ull << 20;
// idle device blocks kept for the next scene handle / pool re-size (a fresh cudaMalloc of a 50 GB arena costs seconds): WT_CACHE_GB overrides the 112 GB default
const size_t kCacheMaxIdle = []() { const char* e = getenv("WT_CACHE_GB"); return (size_t)(e ? atoi(e) : 112) << 30; }();
cudaError_t wt_malloc_impl(void** p, size_t bytes) {
    int dev = 0; cudaGetDevice(&dev);
    {
        std::lock_guard<std::mutex> l(g_blocks.m);
        auto it = g_blocks.idle.find({ dev, bytes });
        if (it != g_blocks.idle.end()) { *p = it->second; g_blocks.idle.erase(it); g_blocks.idle_bytes -= bytes; g_blocks.live[*p] = { dev, bytes }; return cudaSuccess; }
    }
    cudaError_t e = cudaMalloc(p, bytes);
    if (e != cudaSuccess) {     // out of memory with idle blocks around: drop them and retry once
        std::vector<void*> drop;
        { std::lock_guard<std::mutex> l(g_blocks.m); for (auto& kv : g_blocks.idle) drop.push_back(kv.second); g_blocks.idle.clear(); g_blocks.idle_bytes = 0; }
        for (void* q : drop) cudaFree(q);
        cudaGetLastError();
        e = cudaMalloc(p, bytes);
    }
    if (e == cudaSuccess) { std::lock_guard<std::mutex> l(g_blocks.m); g_blocks.live[*p] = { dev, bytes }; }
    return e;
}
template <class T> cudaError_t wt_malloc(T** p, size_t bytes) { return wt_malloc_impl(reinterpret_cast<void**>(p), bytes); }
void wt_free(void* p) {
    if (!p) return;
    std::pair<int, size_t> info{ -1, 0 };
    {
        std::lock_guard<std::mutex> l(g_blocks.m);
        auto it = g_blocks.live.find(p);
        if (it != g_blocks.live.end()) { info = it->second; g_blocks.live.erase(it); }
        if (info.first >= 0 && info.second >= kCacheMinBlock && g_blocks.idle_bytes + info.second <= kCacheMaxIdle) {
            g_blocks.idle.insert({ info, p }); g_blocks.idle_bytes += info.second; return;
        }
    }
    cudaFree(p);
}
}

// One sub-pool of paths in flight, with its own stream(s) and counters.  A render runs several side by side (wtgpu_render, "sub-pools"): each
// advances its own wavefront -- one vertex per iteration for its paths -- and while one sub-pool's launch drains its stragglers (a handful of
// beams over 10^5 triangles, a path with thousands of diffracting edges), the kernels of the others keep the SMs busy.
struct Pool {
    uint32_t size = 0;
    std::vector<void*> allocs;
    float4 *core = nullptr, *fsd = nullptr, *hit = nullptr;
    uint32_t *alive = nullptr, *keys = nullptr, *order = nullptr, *key_count = nullptr, *key_cursor = nullptr, *trav_list = nullptr, *trav_tris = nullptr, *hit_edges = nullptr, *ap_edges = nullptr;
    uint32_t *spill = nullptr, *spill_ext = nullptr, *big_res_list = nullptr, *edge_bits = nullptr;
    uint2 *flux_items = nullptr, *flux_tasks = nullptr, *closest_tasks = nullptr; unsigned long long* closest_best = nullptr; float4* flux_scratch = nullptr; uint32_t flux_cap = 0; float4* quad_tasks = nullptr; uint32_t quad_cap = 0;
    wt::TravSave *big_save = nullptr, *huge_save = nullptr;
    TravRec* trav_rec = nullptr;
    DevCounters* ctr = nullptr;
    // plt_bdpt (P sample slots, 2P walkers)
    float* bdpt_arena = nullptr;
    float4 *bd_walkers = nullptr, *bd_headers = nullptr, *bd_fsd_out = nullptr;
    int* bd_pending = nullptr; float* bd_L0 = nullptr; uint32_t *bd_nverts = nullptr, *bd_fsd_list = nullptr; unsigned long long* bd_pairs = nullptr;
    // host side: made once per handle (cudaMallocHost / event creation per render cost 5-150 ms of driver time, profiles/r01s3_phases.txt)
    DevCounters* hctr = nullptr;
    cudaStream_t st = nullptr, st_fsd = nullptr; cudaEvent_t ev_iter = nullptr, ev_shade = nullptr, ev_samp = nullptr, ev_done = nullptr;
    std::vector<cudaEvent_t> evs; size_t n_ev = 0;     // per-kernel timing marks (WTGPU_RENDER_TIME_KERNELS)
    // state of the render in progress
    uint64_t iters = 0; unsigned long long total = 0; bool done = false;
    // iterations are queued ahead of the host's look at the counters (kPipeDepth in flight per sub-pool), so the GPU never waits for the host to
    // react between two iterations; the iterations queued after the last useful one find nothing to do
    static constexpr int kPipeDepth = 4;
    DevCounters* hring[kPipeDepth] = {}; cudaEvent_t ev_ring[kPipeDepth] = {}; uint64_t submitted = 0, completed = 0; bool draining = false;
    void free_buffers() { for (void* p : allocs) wt_free(p); allocs.clear(); size = 0; }
    void destroy() {
        free_buffers();
        if (hring[0]) cudaFreeHost(hring[0]);      // (one pinned block holds the ring and hctr)
        for (cudaEvent_t e : ev_ring) if (e) cudaEventDestroy(e);
        for (cudaEvent_t e : evs) cudaEventDestroy(e);
        for (cudaEvent_t e : { ev_iter, ev_shade, ev_samp, ev_done }) if (e) cudaEventDestroy(e);
        if (st) cudaStreamDestroy(st);
        if (st_fsd) cudaStreamDestroy(st_fsd);
    }
};
struct wtgpu_scene {
    int device = 0;
    DScene d{};
    std::vector<void*> allocs;
    wtgpu_sensor sensor{};
    wtgpu_integrator integ{};
    float ray_cull_abs = 0.f;
    uint32_t n_keys = 0;
    // capacities of the per-path lists (dtrav.cuh Caps): grown by wtgpu_render when a render needed more; kept for the next render
    Caps caps{};
    // the sub-pools: sized lazily for (paths in flight, number of sub-pools, kind of render, caps)
    std::vector<Pool> pools; uint32_t pool = 0, pool_parts = 0; Caps pool_caps{}; uint32_t pool_kind = 0;
    uint32_t bit_words = 0, big_blocks = 0;
    DevCounters total_ctr{};            // the sub-pools' counters of the last pass, summed (needs: maximum)
    cudaEvent_t ev_begin = nullptr, ev_end = nullptr;
    bool has_sobol = false;             // sobolld generator matrices are in constant memory of this device
    wt::FLut lut{};                     // plt_bdpt: Fraunhofer sampling tables
    void free_pool() { for (Pool& p : pools) p.free_buffers(); pool = 0; }
    ~wtgpu_scene() {
        cudaSetDevice(device);
        for (void* p : allocs) wt_free(p);
        if (ev_begin) cudaEventDestroy(ev_begin);
        if (ev_end) cudaEventDestroy(ev_end);
        for (Pool& p : pools) p.destroy();
    }
};
static void caps_derive(Caps& c) { c.ap_words = 16u + 9u * c.seg; c.arena_words = 2u * c.verts * wt::kVertWords + 2u * c.ap_walk * c.ap_words; }
static bool caps_equal(const Caps& a, const Caps& b) { return a.spill_words == b.spill_words && a.edges == b.edges && a.seg == b.seg && a.ap_walk == b.ap_walk && a.verts == b.verts; }

template <class T> static int upload(wtgpu_scene* s, const T* src, size_t n, const T** dst) {
    void* p = nullptr;
    const size_t bytes = std::max<size_t>(1, n) * sizeof(T);
    CK(wt_malloc(&p, bytes));
    s->allocs.push_back(p);
    if (n && src) CK(cudaMemcpy(p, src, n * sizeof(T), cudaMemcpyHostToDevice));
    *dst = reinterpret_cast<const T*>(p);
    return WTGPU_OK;
}

extern "C" {

int wtgpu_device_count(void) { int n = 0; if (cudaGetDeviceCount(&n) != cudaSuccess) return 0; return n; }
const char* wtgpu_last_error(void) { return g_err.c_str(); }

uint64_t wtgpu_debug_sizeof(int which) {
    switch (which) {
    case 0: return sizeof(wtgpu_node); case 1: return sizeof(wtgpu_leaf); case 2: return sizeof(wtgpu_tri); case 3: return sizeof(wtgpu_tri_meta);
    case 4: return sizeof(wtgpu_tri_shading); case 5: return sizeof(wtgpu_edge); case 6: return sizeof(wtgpu_shape); case 7: return sizeof(wtgpu_spectrum);
    case 8: return sizeof(wtgpu_bsdf); case 9: return sizeof(wtgpu_bsdf_bin); case 10: return sizeof(wtgpu_emitter); case 11: return sizeof(wtgpu_kdist);
    case 12: return sizeof(wtgpu_sensor); case 13: return sizeof(wtgpu_integrator); case 14: return sizeof(wtgpu_scene_desc); case 15: return sizeof(wtgpu_render_opts);
    case 16: return sizeof(wtgpu_stats); case 17: return sizeof(wtgpu_ray_query); case 18: return sizeof(wtgpu_ray_hit); case 19: return sizeof(wtgpu_cone_query);
    case 20: return sizeof(wtgpu_cone_hit); case 21: return sizeof(wthost_mesh_desc); case 22: return sizeof(wtgpu_sobol_entry);
    }
    return 0;
}

int wtgpu_scene_create(const wtgpu_scene_desc* desc, int device, wtgpu_scene** out) {
    if (!desc || !out) { g_err = "null argument"; return WTGPU_E_INVALID; }
    if (desc->api_version != WTGPU_API_VERSION) { g_err = "api version mismatch"; return WTGPU_E_INVALID; }
    if (desc->integrator.type != WTGPU_INTEGRATOR_PLT_PATH && desc->integrator.type != WTGPU_INTEGRATOR_PLT_BDPT) { g_err = "unknown integrator type"; return WTGPU_E_UNSUPPORTED; }
    const bool bdpt = desc->integrator.type == WTGPU_INTEGRATOR_PLT_BDPT;
    if (desc->integrator.max_depth == 0u || desc->integrator.max_depth > 4000u) { g_err = "integrator max_depth out of range (1..4000)"; return WTGPU_E_INVALID; }
    for (uint32_t i = 0; i < desc->n_bsdfs; ++i)
        if (desc->bsdfs[i].type > WTGPU_BSDF_SCALE) { g_err = "bsdf type " + std::to_string(desc->bsdfs[i].type) + " (mask / normalmap / bumpmap: need textures) is not implemented on the device"; return WTGPU_E_UNSUPPORTED; }
    // plt_bdpt.cpp:189-194: the Fraunhofer sampling tables are only needed (and only loaded by the reference) when the sensor is not ray-tracing only
    if (bdpt && desc->integrator.fsd && !desc->sensor.ray_trace_only && (!desc->fsd_lut_n || !desc->fsd_lut_m || !desc->fsd_icdf1 || !desc->fsd_icdf2 || !desc->fsd_icdf_theta1 || !desc->fsd_icdf_theta2)) {
        g_err = "plt_bdpt with FSD needs the Fraunhofer sampling tables (fsd_lut_*)"; return WTGPU_E_INVALID; }
    if (desc->n_edges >= (1u << 26)) { g_err = "more than 2^26 edges unsupported"; return WTGPU_E_UNSUPPORTED; }
    if (desc->sensor.rf_radius > 4) { g_err = "reconstruction filter radius > 4 unsupported"; return WTGPU_E_UNSUPPORTED; }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) { g_err = "no CUDA device"; return WTGPU_E_NO_DEVICE; }
    CK(cudaSetDevice(device));
    auto* s = new wtgpu_scene(); s->device = device;
    DScene& d = s->d;
    int rc = WTGPU_OK;
#define UP(field, src, n) if ((rc = upload(s, src, n, &d.field)) != WTGPU_OK) { delete s; return rc; }
    UP(nodes, desc->nodes, desc->n_nodes) UP(leaves, desc->leaves, desc->n_leaves)
    { const wtgpu_tri* t; if ((rc = upload(s, desc->tris, desc->n_tris, &t)) != WTGPU_OK) { delete s; return rc; } d.tris = reinterpret_cast<const float4*>(t); }
    UP(tri_meta, desc->tri_meta, desc->n_tris) UP(tri_shading, desc->tri_shading, desc->n_tris) UP(edges, desc->edges, desc->n_edges)
    UP(shapes, desc->shapes, desc->n_shapes) UP(shape_tri_tuid, desc->shape_tri_tuid, desc->n_shape_tris) UP(shape_tri_cdf, desc->shape_tri_cdf, desc->n_shape_cdf)
    UP(spectra, desc->spectra, desc->n_spectra) UP(spectrum_data, desc->spectrum_data, desc->n_spectrum_data)
    UP(bsdfs, desc->bsdfs, desc->n_bsdfs) UP(bsdf_bins, desc->bsdf_bins, desc->n_bsdf_bins)
    UP(emitters, desc->emitters, desc->n_emitters) UP(emitter_cdf, desc->emitter_cdf, desc->n_emitters + 1) UP(emitter_kdist, desc->emitter_kdist, desc->n_emitters)
    UP(kdist_data, desc->kdist_data, desc->n_kdist_data)
    {   // erf table (include/wt/math/erf_lut.hpp:27-32)
        std::vector<float> lut(1024);
        for (int i = 0; i < 1024; ++i) lut[i] = std::erf((float)i / 1023.f * 3.5f);
        UP(erf_lut, lut.data(), lut.size())
    }
    if (bdpt && desc->fsd_lut_n) {
        s->lut.N = desc->fsd_lut_n; s->lut.M = desc->fsd_lut_m;
        const size_t mm = (size_t)desc->fsd_lut_m * desc->fsd_lut_m;
        if ((rc = upload(s, desc->fsd_icdf_theta1, desc->fsd_lut_n, &s->lut.th1)) != WTGPU_OK || (rc = upload(s, desc->fsd_icdf_theta2, desc->fsd_lut_n, &s->lut.th2)) != WTGPU_OK ||
            (rc = upload(s, desc->fsd_icdf1, mm, &s->lut.c1)) != WTGPU_OK || (rc = upload(s, desc->fsd_icdf2, mm, &s->lut.c2)) != WTGPU_OK) { delete s; return rc; }
    }
#undef UP
    if (desc->sobol_table) {
        wt::SobolTables tb; std::string why;
        if (!wt::sobol_build_tables(desc->sobol_table, tb.ones, tb.twos, why)) { g_err = why; delete s; return WTGPU_E_INVALID; }
        cudaError_t e_ = cudaMemcpyToSymbol(wt::c_sobol, &tb, sizeof(tb));
        if (e_ != cudaSuccess) { g_err = std::string("cudaMemcpyToSymbol(c_sobol): ") + cudaGetErrorString(e_); delete s; return WTGPU_E_CUDA; }
        s->has_sobol = true;
    }
    d.scene_stream = 0u;
    {   // slack of the ray-query range culling (dtrav.cuh RayCull): 1e-5 x the largest |coordinate| of the scene
        float mabs = 0.f;
        for (uint32_t i = 0; i < desc->n_tris; ++i) { const float* v = &desc->tris[i].ax; for (int k = 0; k < 12; ++k) if ((k & 3) != 3) mabs = std::max(mabs, std::fabs(v[k])); }
        s->ray_cull_abs = 1e-5f * mabs;
        d.ray_cull_abs = s->ray_cull_abs;
    }
    d.n_edges_total = desc->n_edges;
    d.root_ptr = desc->root_ptr; d.n_emitters = desc->n_emitters; d.n_bsdfs = desc->n_bsdfs; d.n_tris = desc->n_tris; d.n_nodes = desc->n_nodes;
    d.sensor = desc->sensor; d.integrator = desc->integrator;
    s->sensor = desc->sensor; s->integ = desc->integrator;
    s->n_keys = desc->n_bsdfs + 3u;
    // initial list capacities; a render that needs more grows them (wtgpu_render).  plt_bdpt keeps max_depth + 2 vertices per subpath, but
    // with Russian roulette long subpaths are rare: start at 18 (max_depth 16, the reference scenes' setting) and grow on demand.
    s->caps.tris = wt::kTriRow; s->caps.spill_words = 16u << 20; s->caps.edges = 48u; s->caps.seg = 48u; s->caps.ap_walk = 4u;
    s->caps.verts = bdpt ? std::min(desc->integrator.max_depth + 2u, 18u) : 0u;
    caps_derive(s->caps);
    *out = s;
    return WTGPU_OK;
}

void wtgpu_scene_destroy(wtgpu_scene* s) { delete s; }

void wtgpu_trim(void) {
    std::vector<std::pair<int, void*>> drop;
    { std::lock_guard<std::mutex> l(g_blocks.m); for (auto& kv : g_blocks.idle) drop.push_back({ kv.first.first, kv.second }); g_blocks.idle.clear(); g_blocks.idle_bytes = 0; }
    int cur = 0; cudaGetDevice(&cur);
    for (auto& d : drop) { cudaSetDevice(d.first); cudaFree(d.second); }
    cudaSetDevice(cur);
}

// kinds of pool: what the render at hand needs
enum : uint32_t { POOL_PATH = 1u, POOL_BDPT_WAVE = 2u, POOL_BDPT_MEGA = 3u };
static uint32_t bdpt_max_pairs(uint32_t verts, uint32_t max_depth) {     // strategies per sample: the enumeration of plt_bdpt.cpp:96-110 at full subpath lengths
    uint32_t n_pairs = 0;
    const int n = (int)verts, maxd = (int)max_depth;
    for (int t = 0; t <= n; ++t) for (int q = 0; q <= n; ++q) { const int depth = t + q - 2; if ((t == 1 && q == 1) || depth < 0) continue; if (depth > maxd) break; ++n_pairs; }
    return n_pairs;
}
static size_t pool_bytes(const wtgpu_scene* s, uint32_t kind, uint32_t pool, uint32_t parts, const Caps& c) {
    const size_t P = pool;
    const size_t row = 4ull * (wt::kTriRow + wt::kTriExt) + 8ull + 2ull * sizeof(wt::TravSave), shared = (size_t)parts * (4ull * c.spill_words + 16ull * s->bit_words * s->big_blocks + (kind == POOL_BDPT_WAVE ? 9ull * c.spill_words + (20ull << 20) : 0ull));      // triangle-list row + extent table + hand-over lists; per sub-pool: the arena, scratch bitmaps
    if (kind == POOL_PATH) return shared + P * (16ull * (chunks_of<PathCore>() + chunks_of<PathFsd>() + chunks_of<HitRec>()) + sizeof(TravRec) + row + 12ull * c.edges + 16ull);
    if (kind == POOL_BDPT_MEGA) return shared + P * (4ull * c.arena_words + row + 4ull * c.edges);
    const size_t W2 = 2 * P;
    return shared + P * (4ull * c.arena_words + 16ull * chunks_of<BdHeader>() + 8ull * (bdpt_max_pairs(c.verts, s->integ.max_depth) + 4ull * (c.verts + 1u)) + 16ull) +
           W2 * (16ull * (chunks_of<BdWalker>() + chunks_of<HitRec>()) + sizeof(TravRec) + row + 4ull * c.edges + 12ull + 32ull + 16ull);
}
static void bitmap_geometry(wtgpu_scene* s) {
    int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device);
    s->bit_words = (s->d.n_edges_total + 31u) / 32u + 1u;
    s->big_blocks = (uint32_t)std::max<size_t>(1, std::min<size_t>((size_t)n_sm * 4, (256ull << 20) / (16ull * s->bit_words)));     // 4 warps per block; <= 256 MiB of bitmaps per sub-pool
}
// `pool` paths in flight in all, split over `parts` sub-pools
static int ensure_pool(wtgpu_scene* s, uint32_t kind, uint32_t pool, uint32_t parts) {
    if (s->pool == pool && s->pool_parts == parts && s->pool_kind == kind && caps_equal(s->pool_caps, s->caps)) return WTGPU_OK;
    s->free_pool();
    if (s->pools.size() < parts) s->pools.resize(parts);
    const Caps& c = s->caps;
    const uint32_t psize = ((pool / parts) + 127u) & ~127u;
    for (uint32_t k = 0; k < parts; ++k) {
        Pool& q = s->pools[k];
        int rc = WTGPU_OK;
        auto get = [&](auto** p, size_t bytes) { if (rc != WTGPU_OK) return; cudaError_t e = wt_malloc(p, std::max<size_t>(bytes, 16)); if (e != cudaSuccess) { g_err = std::string("device allocation of the path pool: ") + cudaGetErrorString(e); cudaGetLastError(); rc = WTGPU_E_CUDA; *p = nullptr; } else q.allocs.push_back((void*)*p); };
        const size_t P = psize;
        get(&q.ctr, sizeof(DevCounters));
        get(&q.key_count, 4ull * s->n_keys); get(&q.key_cursor, 4ull * s->n_keys);
        {   // triangle-list arena, extent tables and hand-over lists (one row per path / walker / thread); scratch edge bitmaps of the warp-per-beam resolve
            const size_t rows = kind == POOL_BDPT_WAVE ? 2 * P : P;
            get(&q.spill, 4ull * c.spill_words); get(&q.spill_ext, 4ull * wt::kTriExt * rows); get(&q.big_save, sizeof(wt::TravSave) * rows); get(&q.huge_save, sizeof(wt::TravSave) * rows); get(&q.big_res_list, 4ull * rows);
            if (kind == POOL_BDPT_WAVE) { q.flux_cap = c.spill_words / 2u + (1u << 20); get(&q.flux_items, 8ull * rows); get(&q.flux_tasks, 8ull * ((size_t)q.flux_cap / 32u + rows)); get(&q.flux_scratch, 16ull * q.flux_cap); q.quad_cap = q.flux_cap / 16u + 65536u; get(&q.quad_tasks, 32ull * q.quad_cap);
                                          get(&q.closest_tasks, 8ull * ((size_t)c.spill_words / wt::kClosestChunk + 2 * rows)); get(&q.closest_best, 8ull * rows); }
            get(&q.edge_bits, 16ull * s->bit_words * s->big_blocks);
            if (rc == WTGPU_OK) { cudaError_t e = cudaMemset(q.edge_bits, 0, 16ull * s->bit_words * s->big_blocks); if (e != cudaSuccess) { g_err = "cudaMemset(edge bitmaps)"; rc = WTGPU_E_CUDA; } }
        }
        if (kind == POOL_PATH) {
            get(&q.trav_rec, sizeof(TravRec) * P); get(&q.trav_tris, 4ull * wt::kTriRow * P); get(&q.hit_edges, 4ull * c.edges * P); get(&q.ap_edges, 8ull * c.edges * P);
            get(&q.core, (size_t)chunks_of<PathCore>() * 16 * P); get(&q.fsd, (size_t)chunks_of<PathFsd>() * 16 * P); get(&q.hit, (size_t)chunks_of<HitRec>() * 16 * P);
            get(&q.alive, 4ull * P); get(&q.keys, 4ull * P); get(&q.order, 4ull * P); get(&q.trav_list, 4ull * P);
        } else if (kind == POOL_BDPT_MEGA) {
            get(&q.bdpt_arena, 4ull * c.arena_words * P); get(&q.trav_tris, 4ull * wt::kTriRow * P); get(&q.hit_edges, 4ull * c.edges * P);
        } else {
            const size_t W2 = 2 * P;
            get(&q.bdpt_arena, 4ull * c.arena_words * P);
            get(&q.bd_walkers, (size_t)chunks_of<BdWalker>() * 16 * W2); get(&q.bd_headers, (size_t)chunks_of<BdHeader>() * 16 * P); get(&q.hit, (size_t)chunks_of<HitRec>() * 16 * W2);
            get(&q.bd_pending, 4ull * P); get(&q.bd_L0, 4ull * P); get(&q.bd_nverts, 4ull * W2); get(&q.alive, 4ull * P);
            get(&q.keys, 4ull * W2); get(&q.order, 4ull * W2); get(&q.trav_list, 4ull * W2);
            get(&q.bd_pairs, 8ull * P * (bdpt_max_pairs(c.verts, s->integ.max_depth) + 4ull * (c.verts + 1u)));
            get(&q.trav_rec, sizeof(TravRec) * W2); get(&q.trav_tris, 4ull * wt::kTriRow * W2); get(&q.hit_edges, 4ull * c.edges * W2);
            get(&q.bd_fsd_list, 12ull * W2); get(&q.bd_fsd_out, 32ull * W2);
        }
        if (rc == WTGPU_OK && !q.hctr) {
            if (cudaMallocHost(&q.hring[0], sizeof(DevCounters) * (Pool::kPipeDepth + 1)) != cudaSuccess || cudaStreamCreateWithFlags(&q.st, cudaStreamNonBlocking) != cudaSuccess ||
                cudaStreamCreateWithFlags(&q.st_fsd, cudaStreamNonBlocking) != cudaSuccess || cudaEventCreateWithFlags(&q.ev_iter, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&q.ev_shade, cudaEventDisableTiming) != cudaSuccess || cudaEventCreateWithFlags(&q.ev_samp, cudaEventDisableTiming) != cudaSuccess ||
                cudaEventCreateWithFlags(&q.ev_done, cudaEventDisableTiming) != cudaSuccess) { g_err = "creating the sub-pool's streams / events failed"; rc = WTGPU_E_CUDA; }
            if (rc == WTGPU_OK) {
                q.hctr = q.hring[0] + Pool::kPipeDepth;
                for (int r = 0; r < Pool::kPipeDepth && rc == WTGPU_OK; ++r) {
                    q.hring[r] = q.hring[0] + r;
                    if (cudaEventCreateWithFlags(&q.ev_ring[r], cudaEventDisableTiming) != cudaSuccess) { g_err = "creating the sub-pool's counter ring failed"; rc = WTGPU_E_CUDA; }
                }
            }
        }
        if (rc != WTGPU_OK) { s->free_pool(); return rc; }
        q.size = psize;
    }
    s->pool = pool; s->pool_parts = parts; s->pool_kind = kind; s->pool_caps = s->caps;
    return WTGPU_OK;
}

// film develop (film_storage.hpp:256-291, 354-358): value / weight of the block image + light image / spp
__global__ void k_develop(const float2* __restrict__ block, const float* __restrict__ light, float* __restrict__ out, size_t n, float inv_spp) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        float v = 0.f;
        if (block) { const float2 b = __ldg(block + i); v = b.y > 0.f ? b.x / b.y : 0.f; }
        if (light) v += __ldg(light + i) * inv_spp;
        out[i] = v;
    }
}
__global__ void k_film_add(float* __restrict__ dst, const float* __restrict__ src, size_t n) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) dst[i] += src[i];
}

// One pass over the samples with the scene handle's current capacities, into the device films dblock / dlight: every sub-pool advances its own
// wavefront on its own stream; the host hands a sub-pool its next iteration as soon as the previous one's counters are back.  The summed
// counters are left in s->total_ctr.
static int render_pass(wtgpu_scene* s, const wtgpu_render_opts* o, uint32_t kind, float* dblock, float* dlight, unsigned long long total, uint32_t x1, uint32_t y1,
                       bool use_thread_trav, bool time_phases, uint64_t& launches, uint64_t& iters_total) {
    cudaStream_t user = (cudaStream_t)o->stream;
    const bool bdpt = kind != POOL_PATH;
    const uint32_t parts = s->pool_parts;
    s->d.cap = s->caps;
    // Block size of the one-thread-per-item kernels (generate / resolve / sort / shade / connect).  Their warps run for very different times (a
    // path with thousands of UTD edges next to paths that die at once), and a block's slots are only handed on when its LAST warp retires:
    // with one warp per block a finished warp is replaced immediately.  (The group-traversal and Fraunhofer-sampler kernels keep 128: their
    // shared-memory layout is per 128 threads and they pull work from a cursor anyway.)  WT_BLOCK_T overrides for A/B runs.
    static const unsigned bt = []() { const char* e = getenv("WT_BLOCK_T"); const unsigned v = e ? (unsigned)atoi(e) : kBlockT; return (v == 32u || v == 64u || v == 128u) ? v : kBlockT; }();
    int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device);
    const dim3 blk(128), blkT(bt);
    const bool nosort = (o->flags & WTGPU_RENDER_NO_SORT) != 0;
    const bool has_fsd = kind == POOL_BDPT_WAVE && s->integ.fsd != 0u && !s->sensor.ray_trace_only;
    const dim3 gBig(s->big_blocks);
    // hand-over thresholds of the traversal tiers (gtrav.cuh TravTiers); WT_BIG_TESTED / WT_HUGE_TESTED override for A/B runs
    static const wt::TravTiers tiers = []() { wt::TravTiers t; const char* b = getenv("WT_BIG_TESTED"); const char* h = getenv("WT_HUGE_TESTED");
                                              t.big_tested = b ? (uint32_t)atoi(b) : 192u; t.huge_tested = h ? (uint32_t)atoi(h) : 8000u;
                                              const char* r = getenv("WT_RESTART_DIV"); const char* i = getenv("WT_INIT_DIV"); t.restart_div = r ? (uint32_t)atoi(r) : 16u; t.init_div = i ? (uint32_t)atoi(i) : 4u; return t; }();

    // the sub-pools start after whatever the caller queued on its stream
    CK(cudaEventRecord(s->ev_begin, user));
    std::vector<RenderArgs> args(parts);
    for (uint32_t k = 0; k < parts; ++k) {
        Pool& q = s->pools[k];
        CK(cudaStreamWaitEvent(q.st, s->ev_begin, 0));
        DScene d = s->d; d.spill = q.spill; d.spill_ext = q.spill_ext; d.spill_head = &q.ctr->spill_head;
        RenderArgs& a = args[k];
        a.sc = d; a.core = q.core; a.fsd = q.fsd; a.hit = q.hit; a.alive = q.alive; a.keys = q.keys; a.order = q.order;
        a.trav_rec = q.trav_rec; a.trav_tris = q.trav_tris; a.hit_edges = q.hit_edges; a.ap_edges = q.ap_edges; a.it_parity = 0u;
        a.big_save = q.big_save; a.huge_save = q.huge_save; a.tiers = tiers; a.big_res_list = q.big_res_list; a.closest_tasks = q.closest_tasks; a.closest_best = q.closest_best; a.flux_items = q.flux_items; a.flux_tasks = q.flux_tasks; a.flux_scratch = q.flux_scratch; a.flux_cap = q.flux_cap; a.quad_tasks = q.quad_tasks; a.quad_cap = q.quad_cap; a.edge_bits = q.edge_bits;
        a.key_count = q.key_count; a.key_cursor = q.key_cursor; a.trav_list = q.trav_list; a.ctr = q.ctr; a.film_block = dblock; a.film_light = dlight;
        a.pool = q.size; a.n_keys = s->n_keys; a.seed_lo = (uint32_t)o->seed; a.seed_hi = (uint32_t)(o->seed >> 32);
        a.tile_x0 = o->tile_x0; a.tile_y0 = o->tile_y0; a.tile_w = x1 - o->tile_x0; a.tile_h = y1 - o->tile_y0;
        a.sample_begin = o->sample_begin; a.n_samples = o->sample_end - o->sample_begin; a.total = total; a.part = k; a.n_parts = parts;
        q.total = total > k ? (total - k + parts - 1) / parts : 0ull;       // samples with id = k (mod parts)
        q.iters = 0; q.done = false; q.n_ev = 0; q.submitted = q.completed = 0; q.draining = false;
        if (q.alive) CK(cudaMemsetAsync(q.alive, 0, 4ull * q.size, q.st));
        CK(cudaMemsetAsync(q.key_count, 0, 4ull * s->n_keys, q.st));
        CK(cudaMemsetAsync(q.ctr, 0, sizeof(DevCounters), q.st));
    }
    auto mark = [&](Pool& q) { if (time_phases) { if (q.n_ev == q.evs.size()) { cudaEvent_t e; cudaEventCreate(&e); q.evs.push_back(e); } cudaEventRecord(q.evs[q.n_ev++], q.st); } };

    // one wavefront iteration of sub-pool k, queued on its stream; ends with the counters' read-back and an event
    auto launch_iteration = [&](uint32_t k) -> int {
        Pool& q = s->pools[k];
        RenderArgs& a = args[k];
        cudaStream_t st = q.st;
        const uint32_t pool = q.size;
        const dim3 grd((pool + bt - 1) / bt);
        if (kind == POOL_BDPT_MEGA) {     // one thread per sample, one launch (dbdpt.cuh driver 1)
            BdptArgs b;
            b.sc = a.sc; b.lut = s->lut; b.arena = q.bdpt_arena; b.P = pool; b.trav_tris = q.trav_tris; b.hit_edges = q.hit_edges; b.ctr = q.ctr; b.film_block = dblock; b.film_light = dlight;
            b.seed_lo = a.seed_lo; b.seed_hi = a.seed_hi; b.tile_x0 = a.tile_x0; b.tile_y0 = a.tile_y0; b.tile_w = a.tile_w; b.tile_h = a.tile_h; b.sample_begin = a.sample_begin; b.total = total;
            k_bdpt<<<pool / 128, blk, 0, st>>>(b); ++launches;
        } else if (bdpt) {      // wavefront (dbdpt.cuh driver 2)
            const uint32_t P = pool, W2 = 2u * pool;
            const uint32_t nmaxv = s->caps.verts + 1u;
            BdArgs b;
            b.r = a; b.r.pool = W2;
            b.lut = s->lut; b.arena = q.bdpt_arena; b.P = P; b.walkers = q.bd_walkers; b.headers = q.bd_headers;
            b.pending = q.bd_pending; b.L0 = q.bd_L0; b.nverts = q.bd_nverts; b.pairs = q.bd_pairs; b.fsd_list = q.bd_fsd_list; b.fsd_out = q.bd_fsd_out; b.trav_rec = q.trav_rec; b.trav_tris = q.trav_tris;
            for (int c = 0; c < 5; ++c) b.pair_off[c] = (size_t)P * nmaxv * (size_t)c;      // classes 0-3 hold <= cap.verts + 1 strategies per sample, class 4 the rest
            const dim3 gP((P + bt - 1) / bt), gW((W2 + bt - 1) / bt), gC(n_sm * 8), gCT(n_sm * 8 * (128 / bt));
            // Fraunhofer direction sampling of iteration j runs on the sub-pool's second stream, overlapped with the strategies of j and the walk
            // kernels of j+1; its walkers rejoin at "finish" in iteration j+1.  Three rotating lists keep producer and consumers apart.
            const uint64_t it = q.iters;
            b.fl_cur = (uint32_t)(it % 3ull); b.fl_next = (uint32_t)((it + 1ull) % 3ull); b.fl_fin = (uint32_t)((it + 2ull) % 3ull);
            b.tag = 16.f + (float)(it % 1024ull); b.tag_fin = 16.f + (float)((it + 1023ull) % 1024ull);
            mark(q);
            k_bd_generate<<<gP, blkT, 0, st>>>(b); ++launches; mark(q);
            if (use_thread_trav) { k_bd_traverse<<<gW, blkT, 0, st>>>(b); ++launches; }
            else {
                k_bd_gtraverse<<<gC, blk, 0, st>>>(b); k_bd_wtraverse<<<gC, blk, 0, st>>>(b); k_bd_ctraverse<<<dim3(n_sm * WT_CT_MINB), dim3(256), 0, st>>>(b);
                k_bd_resolve<<<gW, blkT, 0, st>>>(b); k_bd_closest_chunks<<<gC, blk, 0, st>>>(b); k_bd_resolve_big<<<gBig, blk, 0, st>>>(b, s->bit_words); k_bd_flux_chunks<<<gC, blk, 0, st>>>(b); k_bd_quad_tasks<<<gC, blk, 0, st>>>(b); k_bd_flux_finish<<<gBig, blk, 0, st>>>(b, s->bit_words); launches += 9;
            }
            mark(q);
            k_hist<<<gW, blkT, s->n_keys * 4, st>>>(b.r);
            k_scan<<<1, 1024, 0, st>>>(b.r);
            k_scatter<<<gW, blkT, 0, st>>>(b.r); launches += 3; mark(q);
            k_bd_reset<<<1, 32, 0, st>>>(b);
            k_bd_shade<<<gW, blkT, 0, st>>>(b); launches += 2;
            if (has_fsd) {
                CK(cudaEventRecord(q.ev_shade, st));
                if (it > 0) { CK(cudaStreamWaitEvent(st, q.ev_samp, 0)); k_bd_fsd_finish<<<gW, blkT, 0, st>>>(b); ++launches; }
                CK(cudaStreamWaitEvent(q.st_fsd, q.ev_shade, 0));
                CK(cudaMemsetAsync(&q.ctr->fsd_head, 0, sizeof(int), q.st_fsd));
                k_bd_fsd_sample<<<dim3(n_sm * 16), blk, 0, q.st_fsd>>>(b); ++launches;
                CK(cudaEventRecord(q.ev_samp, q.st_fsd));
            }
            mark(q);
            k_bd_connect<0><<<gCT, blkT, 0, st>>>(b); k_bd_connect<1><<<gCT, blkT, 0, st>>>(b); k_bd_connect<2><<<gCT, blkT, 0, st>>>(b);
            k_bd_connect<3><<<gCT, blkT, 0, st>>>(b); k_bd_connect<4><<<gCT, blkT, 0, st>>>(b); launches += 5; mark(q);
        } else {
            a.it_parity = (uint32_t)(q.iters & 1ull);
            mark(q);
            k_generate<<<grd, blkT, 0, st>>>(a); ++launches; mark(q);
            if (use_thread_trav) { k_traverse<<<grd, blkT, 0, st>>>(a); ++launches; }
            else {
                k_gtraverse<<<dim3(n_sm * 8), blk, 0, st>>>(a); k_wtraverse<<<dim3(n_sm * 8), blk, 0, st>>>(a); k_ctraverse<<<dim3(n_sm * WT_CT_MINB), dim3(256), 0, st>>>(a);
                k_resolve<<<grd, blkT, 0, st>>>(a); k_resolve_big<<<gBig, blk, 0, st>>>(a, s->bit_words); launches += 5;
            }
            mark(q);
            if (nosort) {
                k_identity_order<<<grd, blkT, 0, st>>>(a); ++launches;
            } else {
                k_hist<<<grd, blkT, s->n_keys * 4, st>>>(a);
                k_scan<<<1, 1024, 0, st>>>(a);
                k_scatter<<<grd, blkT, 0, st>>>(a); launches += 3;
            }
            mark(q);
            k_reset_trav<<<1, 32, 0, st>>>(a);
            k_shade<<<grd, blkT, 0, st>>>(a); launches += 2; mark(q);
        }
        ++q.iters; ++iters_total;
        const int slot = (int)(q.submitted % (uint64_t)Pool::kPipeDepth);
        CK(cudaMemcpyAsync(q.hring[slot], q.ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, st));
        CK(cudaEventRecord(q.ev_ring[slot], st));
        ++q.submitted;
        return WTGPU_OK;
    };

    static const int depth_env = []() { const char* e = getenv("WT_PIPE_DEPTH"); const int v = e ? atoi(e) : 2; return v >= 1 && v <= Pool::kPipeDepth ? v : 2; }();      // measured: 2 = 4 (+5 % on sponza, double_slits, cornell over 1)
    const uint64_t depth = kind == POOL_BDPT_MEGA ? 1u : (uint64_t)depth_env;      // (the one-launch driver renders everything in its first iteration)
    uint32_t remaining = 0;
    for (uint32_t k = 0; k < parts; ++k) {
        Pool& q = s->pools[k];
        if (q.total == 0ull) { q.done = true; continue; }
        ++remaining;
        while (q.submitted - q.completed < depth) { const int rc = launch_iteration(k); if (rc != WTGPU_OK) return rc; }
    }
    uint32_t rr = 0;
    while (remaining) {
        bool progressed = false;
        for (uint32_t j = 0; j < parts; ++j) {
            const uint32_t k = (rr + j) % parts;
            Pool& q = s->pools[k];
            if (q.done) continue;
            while (q.completed < q.submitted) {
                const int slot = (int)(q.completed % (uint64_t)Pool::kPipeDepth);
                const cudaError_t e = cudaEventQuery(q.ev_ring[slot]);
                if (e == cudaErrorNotReady) break;
                CK(e);
                ++q.completed; progressed = true;
                if (!q.draining && (kind == POOL_BDPT_MEGA || (q.hring[slot]->next_sample >= q.total && q.hring[slot]->live <= 0))) q.draining = true;
            }
            if (q.draining) { if (q.completed == q.submitted) { q.done = true; --remaining; } continue; }
            if (q.iters > 100000000ull) { g_err = "render did not converge"; return WTGPU_E_CUDA; }
            while (q.submitted - q.completed < depth) { const int rc = launch_iteration(k); if (rc != WTGPU_OK) return rc; }
        }
        if (!progressed) {      // nothing ready: sleep on the oldest iteration of the next sub-pool in turn
            for (uint32_t j = 0; j < parts; ++j) {
                const uint32_t k = (rr + j) % parts; Pool& q = s->pools[k];
                if (!q.done && q.completed < q.submitted) { CK(cudaEventSynchronize(q.ev_ring[(int)(q.completed % (uint64_t)Pool::kPipeDepth)])); break; }
            }
        }
        rr = (rr + 1u) % parts;
    }
    // all sub-pools done: final counters, and the caller's stream continues after them
    DevCounters& T = s->total_ctr; memset(&T, 0, sizeof(T));
    for (uint32_t k = 0; k < parts; ++k) {
        Pool& q = s->pools[k];
        if (q.total == 0ull) continue;
        if (has_fsd) CK(cudaStreamSynchronize(q.st_fsd));
        CK(cudaMemcpyAsync(q.hctr, q.ctr, sizeof(DevCounters), cudaMemcpyDeviceToHost, q.st));
        CK(cudaEventRecord(q.ev_done, q.st));
        CK(cudaStreamWaitEvent(user, q.ev_done, 0));
    }
    for (uint32_t k = 0; k < parts; ++k) {
        Pool& q = s->pools[k];
        if (q.total == 0ull) continue;
        CK(cudaStreamSynchronize(q.st));
        const DevCounters& c = *q.hctr;
        T.samples += c.samples; T.segments += c.segments; T.ray_casts += c.ray_casts; T.cone_casts += c.cone_casts; T.shadow_casts += c.shadow_casts; T.nodes += c.nodes; T.tris += c.tris;
        T.edges += c.edges; T.surface += c.surface; T.fsd += c.fsd; T.null_ += c.null_; T.splats += c.splats; T.overflow += c.overflow; T.shade_nodes += c.shade_nodes; T.shade_tris += c.shade_tris;
        T.shaded += c.shaded; T.walker_steps += c.walker_steps; T.stack_drops += c.stack_drops; for (int i = 0; i < 32; ++i) T.dbg[i] += c.dbg[i];
        for (int i = 0; i < 5; ++i) T.strategies[i] += c.strategies[i];
        T.need_spill = std::max(T.need_spill, std::max(c.need_spill, c.spill_head)); T.need_edges = std::max(T.need_edges, c.need_edges); T.need_seg = std::max(T.need_seg, c.need_seg);
        T.need_ap = std::max(T.need_ap, c.need_ap); T.need_verts = std::max(T.need_verts, c.need_verts);
    }
    CK(cudaGetLastError());
    return WTGPU_OK;
}

int wtgpu_render(wtgpu_scene* s, const wtgpu_render_opts* o, float* film_block, float* film_light, wtgpu_stats* stats) {
    if (!s || !o) { g_err = "null argument"; return WTGPU_E_INVALID; }
    CK(cudaSetDevice(s->device));
    const uint32_t W = s->sensor.width, H = s->sensor.height, C = s->sensor.channels;
    const uint32_t x1 = std::min(o->tile_x1, W), y1 = std::min(o->tile_y1, H);
    if (x1 <= o->tile_x0 || y1 <= o->tile_y0 || o->sample_end <= o->sample_begin) { if (stats) memset(stats, 0, sizeof(*stats)); return WTGPU_OK; }
    if (o->sampler == WTGPU_SAMPLER_SOBOLLD) {
        if (!s->has_sobol) { g_err = "sampler = sobolld needs wtgpu_scene_desc::sobol_table"; return WTGPU_E_INVALID; }
        if (o->spp == 0u || (o->spp & kSobolStreamFlag)) { g_err = "sampler = sobolld: spp out of range"; return WTGPU_E_INVALID; }
        s->d.scene_stream = kSobolStreamFlag | o->spp;
    } else if (o->sampler == WTGPU_SAMPLER_UNIFORM) s->d.scene_stream = 0u;
    else { g_err = "unknown sampler"; return WTGPU_E_UNSUPPORTED; }
    cudaStream_t st = (cudaStream_t)o->stream;
    const unsigned long long total = (unsigned long long)(x1 - o->tile_x0) * (y1 - o->tile_y0) * (o->sample_end - o->sample_begin);
    const bool bdpt = s->integ.type == WTGPU_INTEGRATOR_PLT_BDPT;
    const uint32_t kind = !bdpt ? POOL_PATH : (o->flags & WTGPU_RENDER_BDPT_MEGAKERNEL) ? POOL_BDPT_MEGA : POOL_BDPT_WAVE;
    int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, s->device);
    uint32_t pool = o->pool_size ? o->pool_size : (kind == POOL_BDPT_MEGA ? (uint32_t)n_sm * 8u * 128u : kind == POOL_BDPT_WAVE ? (1u << 18) : (1u << 20));
    if (pool > (1u << 22) && bdpt) { g_err = "plt_bdpt: pool_size > 4M sample slots"; return WTGPU_E_INVALID; }
    pool = (uint32_t)std::min<unsigned long long>(pool, std::max<unsigned long long>(total, 1024ull));
    pool = (pool + 127u) & ~127u;
    const bool time_phases = stats != nullptr && (o->flags & WTGPU_RENDER_TIME_KERNELS) != 0;
    // Sub-pools (Pool above): 4 independent wavefronts side by side once there are enough paths to share out.  Per-kernel timing
    // (WTGPU_RENDER_TIME_KERNELS) wants one kernel at a time on the device: a single sub-pool.  WT_SUBPOOLS overrides for A/B runs.
    uint32_t parts = (kind == POOL_BDPT_MEGA || time_phases) ? 1u : (pool >= (1u << 16) ? 4u : pool >= (1u << 14) ? 2u : 1u);
    if (o->flags & WTGPU_RENDER_ONE_SUBPOOL) parts = 1u;
    if (const char* e = getenv("WT_SUBPOOLS")) { const int v = atoi(e); if (v >= 1 && v <= 16 && kind != POOL_BDPT_MEGA) parts = (uint32_t)v; }

    {   // the group-traversal kernels keep their stacks in shared memory: ask for the large shared-memory carve-out so that the register file, not
        // the L1/shared split, bounds the resident blocks (per device: a process may drive several)
        static std::mutex m; static std::vector<int> done;
        std::lock_guard<std::mutex> l(m);
        if (std::find(done.begin(), done.end(), s->device) == done.end()) {
            cudaFuncSetAttribute(k_gtraverse, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            cudaFuncSetAttribute(wt::k_bd_gtraverse, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
            done.push_back(s->device);
        }
    }
    s->d.ray_cull_abs = (o->flags & WTGPU_RENDER_NO_RAY_CULL) ? std::numeric_limits<float>::infinity() : s->ray_cull_abs;
    // traverse(): eight lanes per beam pay off when queries are long (cone queries over real geometry); on a handful of triangles one thread
    // per beam is faster (measured: double_slits plt_path 99 vs 52 Msamples/s; etoile-like 6 vs 16).  Both give bit-identical results.
    const bool use_thread_trav = (o->flags & WTGPU_RENDER_THREAD_TRAVERSE) ? true : (o->flags & WTGPU_RENDER_GROUP_TRAVERSE) ? false : (!bdpt && s->d.n_tris < 128u);
    if (!s->ev_begin) { CK(cudaEventCreate(&s->ev_begin)); CK(cudaEventCreate(&s->ev_end)); }
    if (!s->bit_words) bitmap_geometry(s);

    // The films of this call are accumulated in scratch device buffers and added to the caller's at the end: a pass that finds a list longer
    // than its row (capacity growth, below) is discarded and repeated, and must not have touched the caller's film.
    const size_t nb = (size_t)W * H * C * 2, nl = (size_t)W * H * C;
    const bool on_dev = o->film_on_device != 0;
    float *dblock = nullptr, *dlight = nullptr;
    struct Scratch { float*& a; float*& b; ~Scratch() { if (a) wt_free(a); if (b) wt_free(b); } } scratch{ dblock, dlight };
    CK(wt_malloc(&dblock, nb * 4)); CK(wt_malloc(&dlight, nl * 4));

    uint64_t launches = 0, iters = 0;
    uint32_t passes = 0;
    cudaEvent_t t0 = nullptr, t1 = nullptr;     // the call's device time on the caller's stream
    CK(cudaEventCreate(&t0)); CK(cudaEventCreate(&t1));
    struct Ev { cudaEvent_t& a; cudaEvent_t& b; ~Ev() { if (a) cudaEventDestroy(a); if (b) cudaEventDestroy(b); } } evg{ t0, t1 };
    CK(cudaEventRecord(t0, st));
    const DevCounters* hctr = &s->total_ctr;
    // ---- capacity growth: two-pass count / fill at the granularity of the render.  The reference keeps cone-query results, edge sets,
    // aperture segments and subpath vertices in std::vector / std::set of any length; here they are rows of HBM arrays (and, for the triangle
    // lists, extents of an arena).  A pass records the longest list any too-short row was asked to hold; the rows are re-sized and the pass
    // repeated, so a result never depends on a capacity.  The capacities stay with the scene handle: the next render starts with rows that fitted.
    for (;;) {
        uint32_t use_pool = pool;
        if (!(s->pool == pool && s->pool_parts == parts && s->pool_kind == kind && caps_equal(s->pool_caps, s->caps))) {
            s->free_pool();
            size_t free_b = 0, total_b = 0; cudaMemGetInfo(&free_b, &total_b);
            { std::lock_guard<std::mutex> l(g_blocks.m); free_b += g_blocks.idle_bytes; }       // idle cached blocks are reclaimable (wt_malloc drops them when cudaMalloc fails)
            while (use_pool > 4096u && pool_bytes(s, kind, use_pool, parts, s->caps) > (size_t)(0.85 * (double)free_b)) use_pool = ((use_pool / 2u) + 127u) & ~127u;     // long rows: fewer paths in flight
        }
        int rc = ensure_pool(s, kind, use_pool, parts);
        if (rc != WTGPU_OK) return rc;
        pool = use_pool;
        CK(cudaMemsetAsync(dblock, 0, nb * 4, st)); CK(cudaMemsetAsync(dlight, 0, nl * 4, st));
        rc = render_pass(s, o, kind, dblock, dlight, total, x1, y1, use_thread_trav, time_phases, launches, iters);
        if (rc != WTGPU_OK) return rc;
        ++passes;
        if (hctr->overflow == 0) break;
        Caps nc = s->caps;
        auto grow = [](uint32_t cur, uint32_t need) { return need > cur ? std::max((need + 31u) & ~31u, cur + cur / 2u) : cur; };
        // the triangle-list arena: the bump cursor kept counting past its end, so the largest demand of an iteration is known
        if (hctr->need_spill > nc.spill_words) nc.spill_words = (uint32_t)std::min<unsigned long long>(0xfff00000ull, (unsigned long long)hctr->need_spill + hctr->need_spill / 4u + (1u << 20));
        nc.edges = grow(nc.edges, hctr->need_edges); nc.seg = grow(nc.seg, hctr->need_seg);
        // (apertures are 4 KB each and every subpath reserves ap_walk of them: grown to the measured need, not in steps of 32)
        if (hctr->need_ap > nc.ap_walk) nc.ap_walk = std::min(std::max(hctr->need_ap, nc.ap_walk + 1u), std::max(s->integ.max_depth, 1u));
        if (bdpt) nc.verts = std::min(grow(nc.verts, hctr->need_verts), s->integ.max_depth + 2u);
        caps_derive(nc);
        if (caps_equal(nc, s->caps) || passes >= 12u) {
            g_err = "a per-path list outgrew its capacity and could not be grown further (triangle-list arena " + std::to_string(hctr->need_spill) + ", edges " + std::to_string(hctr->need_edges) +
                    ", segments " + std::to_string(hctr->need_seg) + ", apertures " + std::to_string(hctr->need_ap) + ", vertices " + std::to_string(hctr->need_verts) + ")";
            return WTGPU_E_CAPACITY;
        }
        s->caps = nc;
    }
    CK(cudaEventRecord(t1, st));
    CK(cudaEventSynchronize(t1));
    float ms = 0; cudaEventElapsedTime(&ms, t0, t1);
    double t_trav = 0, t_shade = 0, t_gen = 0, t_sort = 0, t_conn = 0;
    if (time_phases) {
        const Pool& q = s->pools[0];
        const size_t per_it = (kind == POOL_BDPT_WAVE) ? 6 : 5;      // marks per iteration
        for (size_t i = 0; i + per_it - 1 < q.n_ev; i += per_it) {
            float f;
            cudaEventElapsedTime(&f, q.evs[i], q.evs[i + 1]); t_gen += f;
            cudaEventElapsedTime(&f, q.evs[i + 1], q.evs[i + 2]); t_trav += f;
            cudaEventElapsedTime(&f, q.evs[i + 2], q.evs[i + 3]); t_sort += f;
            cudaEventElapsedTime(&f, q.evs[i + 3], q.evs[i + 4]); t_shade += f;
            if (per_it == 6) { cudaEventElapsedTime(&f, q.evs[i + 4], q.evs[i + 5]); t_conn += f; }
        }
    }

    if (on_dev) {
        if (film_block) k_film_add<<<dim3(n_sm * 4), 256, 0, st>>>(film_block, dblock, nb);
        if (film_light) k_film_add<<<dim3(n_sm * 4), 256, 0, st>>>(film_light, dlight, nl);
        CK(cudaStreamSynchronize(st));
    } else {
        std::vector<float> tmp(std::max(nb, nl));
        CK(cudaMemcpy(tmp.data(), dblock, nb * 4, cudaMemcpyDeviceToHost));
        if (film_block) for (size_t i = 0; i < nb; ++i) film_block[i] += tmp[i];
        CK(cudaMemcpy(tmp.data(), dlight, nl * 4, cudaMemcpyDeviceToHost));
        if (film_light) for (size_t i = 0; i < nl; ++i) film_light[i] += tmp[i];
    }
    if (stats) {
        memset(stats, 0, sizeof(*stats));
        stats->samples = hctr->samples; stats->segments = hctr->segments; stats->ray_casts = hctr->ray_casts; stats->cone_casts = hctr->cone_casts;
        stats->shadow_casts = hctr->shadow_casts; stats->nodes_visited = hctr->nodes + hctr->shade_nodes; stats->tris_tested = hctr->tris + hctr->shade_tris;
        stats->traverse_nodes = hctr->nodes; stats->traverse_tris = hctr->tris; stats->shaded_paths = hctr->shaded; stats->edges_fetched = hctr->edges;
        stats->surface_interactions = hctr->surface; stats->fsd_interactions = hctr->fsd; stats->null_interactions = hctr->null_; stats->splats = hctr->splats;
        stats->capacity_overflows = hctr->overflow; stats->kernel_launches = launches; stats->iterations = iters;
        stats->gpu_ms = ms; stats->traverse_ms = t_trav; stats->shade_ms = t_shade; stats->generate_ms = t_gen; stats->sort_ms = t_sort; stats->connect_ms = t_conn;
        for (int c = 0; c < 5; ++c) stats->strategies[c] = hctr->strategies[c];
        stats->walker_steps = hctr->walker_steps;
        if (getenv("WT_DEBUG_TEAM")) { fprintf(stderr, "team dbg:"); for (int i = 0; i < 32; ++i) fprintf(stderr, " %llu", hctr->dbg[i]); fprintf(stderr, "\n"); }
        stats->passes = passes; stats->stack_drops = hctr->stack_drops; stats->pool_used = pool; stats->subpools = parts;
        stats->cap_tris = s->caps.spill_words; stats->cap_edges = s->caps.edges; stats->cap_segments = s->caps.seg; stats->cap_apertures = s->caps.ap_walk; stats->cap_vertices = s->caps.verts;
    }
    if (hctr->stack_drops) {
        g_err = "a BVH traversal stack (64 entries for rays, 128 for cones, as bvh8w.cpp's) was full " + std::to_string((unsigned long long)hctr->stack_drops) + " times and dropped children: the tree is too deep for the reference's traversal";
        return WTGPU_E_CAPACITY;
    }
    return WTGPU_OK;
}

int wtgpu_get_capacities(wtgpu_scene* s, uint32_t out[5]) {
    if (!s || !out) { g_err = "null argument"; return WTGPU_E_INVALID; }
    out[0] = s->caps.spill_words; out[1] = s->caps.edges; out[2] = s->caps.seg; out[3] = s->caps.ap_walk; out[4] = s->caps.verts;
    return WTGPU_OK;
}
int wtgpu_set_capacities(wtgpu_scene* s, const uint32_t in[5]) {
    if (!s || !in) { g_err = "null argument"; return WTGPU_E_INVALID; }
    const bool bdpt = s->integ.type == WTGPU_INTEGRATOR_PLT_BDPT;
    s->caps.tris = wt::kTriRow; s->caps.spill_words = std::max(1024u, in[0]); s->caps.edges = std::max(4u, in[1]); s->caps.seg = std::max(4u, in[2]); s->caps.ap_walk = std::max(1u, in[3]);
    s->caps.verts = bdpt ? std::min(s->integ.max_depth + 2u, std::max(3u, in[4])) : 0u;
    caps_derive(s->caps);
    return WTGPU_OK;
}

int wtgpu_develop(const wtgpu_sensor* sensor, uint32_t spp, const float* film_block, const float* film_light, float* out) {
    if (!sensor || !out) return WTGPU_E_INVALID;
    const size_t n = (size_t)sensor->width * sensor->height * sensor->channels;
    const float sl = spp > 0 ? 1.f / (float)spp : 0.f;      // film_storage.hpp:354-358
    for (size_t i = 0; i < n; ++i) {
        float v = 0.f;
        if (film_block) { const float val = film_block[2 * i], w = film_block[2 * i + 1]; v = w > 0.f ? val / w : 0.f; }
        if (film_light) v += film_light[i] * sl;
        out[i] = v;
    }
    return WTGPU_OK;
}

int wtgpu_develop_device(const wtgpu_sensor* sensor, uint32_t spp, const float* d_block, const float* d_light, float* d_out, void* stream, int device) {
    if (!sensor || !d_out) { g_err = "null argument"; return WTGPU_E_INVALID; }
    CK(cudaSetDevice(device));
    const size_t n = (size_t)sensor->width * sensor->height * sensor->channels;
    int n_sm = 148; cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device);
    k_develop<<<dim3(n_sm * 8), 256, 0, (cudaStream_t)stream>>>(reinterpret_cast<const float2*>(d_block), d_light, d_out, n, spp > 0 ? 1.f / (float)spp : 0.f);
    CK(cudaGetLastError());
    return WTGPU_OK;
}

int wtgpu_debug_intersect_rays(wtgpu_scene* s, uint32_t n, const wtgpu_ray_query* q, wtgpu_ray_hit* out) {
    if (!s) return WTGPU_E_INVALID;
    CK(cudaSetDevice(s->device));
    wtgpu_ray_query* dq; wtgpu_ray_hit* dh;
    CK(wt_malloc(&dq, sizeof(*dq) * n)); CK(wt_malloc(&dh, sizeof(*dh) * n));
    CK(cudaMemcpy(dq, q, sizeof(*dq) * n, cudaMemcpyHostToDevice));
    k_debug_rays<<<(n + 127) / 128, 128>>>(s->d, n, dq, dh, nullptr);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dh, sizeof(*dh) * n, cudaMemcpyDeviceToHost));
    wt_free(dq); wt_free(dh);
    return WTGPU_OK;
}
int wtgpu_debug_shadow_rays(wtgpu_scene* s, uint32_t n, const wtgpu_ray_query* q, uint32_t* out) {
    if (!s) return WTGPU_E_INVALID;
    CK(cudaSetDevice(s->device));
    wtgpu_ray_query* dq; uint32_t* dh;
    CK(wt_malloc(&dq, sizeof(*dq) * n)); CK(wt_malloc(&dh, 4ull * n));
    CK(cudaMemcpy(dq, q, sizeof(*dq) * n, cudaMemcpyHostToDevice));
    k_debug_rays<<<(n + 127) / 128, 128>>>(s->d, n, dq, nullptr, dh);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dh, 4ull * n, cudaMemcpyDeviceToHost));
    wt_free(dq); wt_free(dh);
    return WTGPU_OK;
}
int wtgpu_debug_intersect_cones(wtgpu_scene* s, uint32_t n, const wtgpu_cone_query* q, wtgpu_cone_hit* out) {
    if (!s) return WTGPU_E_INVALID;
    CK(cudaSetDevice(s->device));
    wtgpu_cone_query* dq; wtgpu_cone_hit* dh;
    CK(wt_malloc(&dq, sizeof(*dq) * n)); CK(wt_malloc(&dh, sizeof(*dh) * n));
    CK(cudaMemcpy(dq, q, sizeof(*dq) * n, cudaMemcpyHostToDevice));
    CK(cudaMemset(dh, 0, sizeof(*dh) * n));
    k_debug_cones<<<(n + 63) / 64, 64>>>(s->d, n, dq, dh);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dh, sizeof(*dh) * n, cudaMemcpyDeviceToHost));
    wt_free(dq); wt_free(dh);
    return WTGPU_OK;
}
int wtgpu_debug_sobol(wtgpu_scene* s, uint64_t seed, uint64_t g0, uint32_t n, uint32_t* out_num, float* out_val) {
    if (!s || !out_num || !out_val) { g_err = "null argument"; return WTGPU_E_INVALID; }
    if (!s->has_sobol) { g_err = "scene has no sobol table"; return WTGPU_E_INVALID; }
    CK(cudaSetDevice(s->device));
    const size_t m = (size_t)n * WTGPU_SOBOL_DIMS;
    if (m == 0) return WTGPU_OK;
    uint32_t* dn; float* dv; CK(wt_malloc(&dn, 4 * m)); CK(wt_malloc(&dv, 4 * m));
    k_debug_sobol<<<(unsigned)((m + 127) / 128), 128>>>((uint32_t)seed, (uint32_t)(seed >> 32), g0, n, dn, dv);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out_num, dn, 4 * m, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(out_val, dv, 4 * m, cudaMemcpyDeviceToHost));
    wt_free(dn); wt_free(dv);
    return WTGPU_OK;
}

int wtgpu_debug_pmath(int fn, uint32_t n, const float* x, const float* y, float* out, int device) {
    if (!x || !out || fn < 0 || fn > 10) { g_err = "bad argument"; return WTGPU_E_INVALID; }
    if (n == 0) return WTGPU_OK;
    CK(cudaSetDevice(device));
    float *dx, *dy = nullptr, *dout;
    CK(wt_malloc(&dx, 4ull * n)); CK(wt_malloc(&dout, 4ull * n));
    CK(cudaMemcpy(dx, x, 4ull * n, cudaMemcpyHostToDevice));
    if (y) { CK(wt_malloc(&dy, 4ull * n)); CK(cudaMemcpy(dy, y, 4ull * n, cudaMemcpyHostToDevice)); }
    k_debug_pmath<<<(n + 127) / 128, 128>>>(fn, n, dx, dy, dout);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, dout, 4ull * n, cudaMemcpyDeviceToHost));
    wt_free(dx); wt_free(dout); if (dy) wt_free(dy);
    return WTGPU_OK;
}

int wtgpu_debug_rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out, int device) {
    CK(cudaSetDevice(device));
    float* d; CK(wt_malloc(&d, 4ull * n));
    k_debug_rng<<<(n + 127) / 128, 128>>>((uint32_t)seed, (uint32_t)(seed >> 32), pixel, sample, n, d);
    CK(cudaDeviceSynchronize());
    CK(cudaMemcpy(out, d, 4ull * n, cudaMemcpyDeviceToHost));
    wt_free(d);
    return WTGPU_OK;
}

} // extern "C"
