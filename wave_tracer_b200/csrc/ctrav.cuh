// ctrav.cuh -- TEAM traversal: the cone queries the group traversal (gtrav.cuh) hands over because they run over hundreds to 10^5 triangles
// (a beam a few millimetres wide over a finely tessellated mesh).  (Included by wavefront.cu after gtrav.cuh.)
//
// A team of TEAM threads (one warp, or one 256-thread block) continues ONE beam at a time from the state the previous tier saved (TravSave:
// machine state + stack; nothing is redone).  The team's first warp is the control warp: it runs the same machine as gtrav.cuh (g_begin /
// g_query_done, ray queries, the sequential commit logic); the cone query's triangle work is done by the whole team in batches:
//   L  lookahead: the top NW = TEAM/8 stack entries are looked at together, one 8-lane group each -- an internal node's eight children are tested
//      against the current search range and ranked in pop order, NOT pushed; a leaf entry is itself.  The batch window is the leading run of
//      "simple" entries: leaves, and nodes whose surviving children are all leaves (each <= 8 triangles).  A node on top of the stack that is
//      not simple is expanded for real (that is simply the sequential algorithm's next step).
//   T1 every triangle of the window's leaves (up to 8 TEAM of them) goes through the cheap rejections of the cone-triangle test
//      (cone_tri_maybe: z range, separating axes); the survivors -- about one in seven -- are compacted;
//   T2 the survivors go through the full test (intersect_cone_tri), one per thread: the expensive code runs with full warps;
//   C  commit, in stack order, exactly what the sequential loop does leaf by leaf (bvh8w.cpp:123-185, 245-347): append the accepted triangles,
//      closest hit, search-range update, unwinding of stale stack entries.  Everything in the batch was computed against the range at the start
//      of the batch; if a commit NARROWS the range, the rest of the batch is discarded -- the not-yet-committed children of the entry being
//      committed go onto the stack (that expansion was valid: the range had not changed when the sequential loop would have made it), later
//      entries are still on the stack untouched -- and the next batch tests them again.  When no leaf of the batch improves the closest hit and
//      no window entry is stale (the common case once the hit distance has settled), the whole batch commits with one parallel pass.
// Every decision is taken on the same values, in the same order, as the sequential code: lists, distances and counters are identical.
#pragma once

namespace wt {

template <int TEAM> struct alignas(16) TShared {
    static constexpr int NW = TEAM / 8;
    GShared g;                                  // the beam's stack (+ the ranked-push scratch of the control warp's node steps)
    Cone env; Frame frame; Range crange; V3 inv; int nx, ny, nz;       // the query, published by the control warp for the batch phases
    int cmd, s, wlen, nl, nsurv;
    int wn[NW], wfl[NW];                        // per window entry: children in pop order (1 for a leaf entry); flags: 1 internal, 2 has an internal child, 4 no entry, 8 a leaf of > 8 triangles
    float wc_tmin[NW][8]; int32_t wc_ptr[NW][8]; uint32_t wc_t0[NW][8], wc_cnt[NW][8];
    int lbase[NW];                              // first batch leaf of window entry w
    uint32_t bt0[TEAM], bcnt[TEAM];             // batch leaves: triangle range
    float dres[TEAM * 8]; uint16_t surv[TEAM * 8];
    uint32_t lmask[TEAM], larg[TEAM]; float ldmin[TEAM];    // per batch leaf: accepted slots, first closest slot, its distance
};
template <int TEAM> WT_D void t_sync() { if (TEAM == 32) __syncwarp(); else __syncthreads(); }
enum { TC_DONE = 0, TC_BATCH = 1 };

WT_D void t_prune(GShared& sh, GTrav& t) { while (t.s > 0 && sh.tmin[t.s - 1] >= t.crange.mx) --t.s; }
// lane-0 reservation of list storage up to `upto`, result broadcast over the control warp
WT_D void t_reserve(const DScene& sc, GTrav& t, uint32_t upto, unsigned lane) {
    if (upto > t.tw.alloc_end && !t.tw.fail) {
        if (lane == 0u) tw_reserve(sc, t.tw, upto);
        t.tw.n_ext = __shfl_sync(0xffffffffu, t.tw.n_ext, 0); t.tw.alloc_end = __shfl_sync(0xffffffffu, t.tw.alloc_end, 0); t.tw.fail = __shfl_sync(0xffffffffu, t.tw.fail ? 1 : 0, 0) != 0;
        __syncwarp();
    }
}

template <int TEAM, class Emit>
WT_D void t_traverse_all(const DScene& sc, int n_items, const TravSave* saves, int* cursor, TShared<TEAM>& sh, Counters& ctr,
                         TravSave* huge_save, int* n_huge, uint32_t huge_tested, Emit&& emit) {
    constexpr int NW = TEAM / 8;
    const unsigned FULL = 0xffffffffu;
    const unsigned tid = threadIdx.x % (unsigned)TEAM, lane = threadIdx.x & 31u;
    const bool ctl = tid < 32u;
    GLane gw; gw.gl = lane; gw.gshift = 0u; gw.gmask = FULL;                                    // the control warp as one "group" (set-up / bookkeeping code of gtrav.cuh)
    GLane g8; g8.gl = lane & 7u; g8.gshift = 0u; g8.gmask = 0xffu;                              // its lanes 0-7: real node steps
    GLane gg; gg.gl = lane & 7u; gg.gshift = lane & 24u; gg.gmask = 0xffu << gg.gshift;         // this thread's group of eight (lookahead)
    const int grp = (int)(tid >> 3);
    GTrav t; t.mode = 0; t.s = 0;
    bool have = false; int item = 0;
    for (;;) {
        // ---- control warp: the sequential machine, up to the next cone batch
        if (ctl) {
            int cmd = TC_BATCH;
            for (;;) {
                if (!have) {
                    int i = 0;
                    if (lane == 0u) i = atomicAdd(cursor, 1);
                    i = __shfl_sync(FULL, i, 0);
                    if (i >= n_items) { cmd = TC_DONE; break; }
                    const TravSave& sv = saves[i];
                    t = sv.t; item = sv.item;
                    for (int k = (int)lane; k < t.s; k += 32) { sh.g.tmin[k] = sv.tmin[k]; sh.g.ptr[k] = sv.ptr[k]; }
                    __syncwarp();
                    have = true;
                }
                if (t.s == 0) {
                    TravRec out;
                    if (g_query_done(sc, gw, sh.g, t, out, ctr)) { emit(item, out, gw); have = false; __syncwarp(); }
                    continue;
                }
                const int32_t top = sh.g.ptr[t.s - 1];
                if (t.mode == 1) {      // a ray query (ballistic segment): the whole of it in the control warp
                    uint32_t rt0 = 0u, rcnt = 0u; bool ray_leaf = false;
                    if (top >= 0) {
                        const uint2 tr = __ldg(reinterpret_cast<const uint2*>(&sc.nodes[top - 1].tris_start));
                        if (tr.y <= 16u) { rt0 = tr.x; rcnt = tr.y; ray_leaf = true; if (lane == 0u) ctr.nodes++; }      // ray_traversal_treat_node_as_leaf_if_triangle_count_lt (bvh8w.cpp:29)
                        else {
                            --t.s;
                            if (lane < 8u) g_node_step(sc, g8, sh.g, t, top, ctr);
                            t.s = __shfl_sync(FULL, t.s, 0);
                            __syncwarp();
                            continue;
                        }
                    }
                    if (!ray_leaf) { const wtgpu_leaf lf = sc.leaves[-top - 1]; rt0 = lf.tris_ptr; rcnt = lf.count; }
                    --t.s;
                    const V3 ro = t.env.o, rd = t.env.d;
                    bool hit = false;
                    for (uint32_t base = 0; base < rcnt; base += 32u) {     // ray_gather (bvh8w.cpp:394-467), 32 triangles at a time
                        const uint32_t k = base + lane; const bool valid = k < rcnt; const uint32_t tuid = rt0 + k;
                        float z = -WT_INF, bx = 0.f, by = 0.f; bool front = false;
                        if (valid) { const Tri3 tr = load_tri(sc, tuid); ctr.tris++; z = intersect_ray_tri_w(ro, rd, tr.a, tr.b, tr.c, t.qrange, bx, by); front = dot(tr.n, rd) <= 0.f; }
                        const int w = g_argmin32(z, valid && z != -WT_INF && z < t.rec.dist, lane);
                        if (w >= 0) { t.rec.dist = __shfl_sync(FULL, z, w); t.rec.bx = __shfl_sync(FULL, bx, w); t.rec.by = __shfl_sync(FULL, by, w); t.rec.tuid = __shfl_sync(FULL, tuid, w); t.rec.front = __shfl_sync(FULL, front ? 1 : 0, w) != 0; hit = true; }
                    }
                    if (hit) while (t.s > 0 && sh.g.tmin[t.s - 1] >= t.rec.dist) --t.s;
                    continue;
                }
                // a cone query with work on its stack.  One that has grown very large goes to the next tier (a whole block per beam).
                if (huge_save && t.qtested > huge_tested) {
                    int pos = 0;
                    if (lane == 0u) pos = atomicAdd(n_huge, 1);
                    pos = __shfl_sync(FULL, pos, 0);
                    TravSave& sv = huge_save[pos];
                    if (lane == 0u) { sv.t = t; sv.item = item; }
                    for (int k = (int)lane; k < t.s; k += 32) { sv.tmin[k] = sh.g.tmin[k]; sv.ptr[k] = sh.g.ptr[k]; }
                    have = false; t.s = 0; t.mode = 0;
                    __syncwarp();
                    continue;
                }
                break;
            }
            if (lane == 0u) {
                sh.cmd = cmd; sh.s = t.s; sh.nsurv = 0;
                if (cmd == TC_BATCH) { sh.env = t.env; sh.frame = t.frame; sh.crange = t.crange; sh.inv = t.inv; sh.nx = t.nx ? 1 : 0; sh.ny = t.ny ? 1 : 0; sh.nz = t.nz ? 1 : 0; }
            }
        }
        t_sync<TEAM>();
        if (sh.cmd == TC_DONE) break;
        const int s0 = sh.s;
        const Cone env = sh.env; const Frame frame = sh.frame; const Range cr = sh.crange;

        // ---- L: lookahead over the top NW stack entries, one group of eight lanes each
        {
            const int sidx = s0 - 1 - grp;
            if (sidx >= 0) {
                const int32_t ptr = sh.g.ptr[sidx];
                if (ptr < 0) {
                    if (gg.gl == 0u) {
                        const wtgpu_leaf lf = sc.leaves[-ptr - 1];
                        sh.wn[grp] = 1; sh.wfl[grp] = lf.count > 8u ? 8 : 0;
                        sh.wc_tmin[grp][0] = sh.g.tmin[sidx]; sh.wc_ptr[grp][0] = ptr; sh.wc_t0[grp][0] = lf.tris_ptr; sh.wc_cnt[grp][0] = lf.count;
                    }
                } else {
                    const wtgpu_node* __restrict__ n = sc.nodes + (ptr - 1);
                    const float mnx = __ldg(&n->minx[gg.gl]), mny = __ldg(&n->miny[gg.gl]), mnz = __ldg(&n->minz[gg.gl]);
                    const float mxx = __ldg(&n->maxx[gg.gl]), mxy = __ldg(&n->maxy[gg.gl]), mxz = __ldg(&n->maxz[gg.gl]);
                    const int32_t ch = __ldg(&n->child[gg.gl]);
                    float tmin;
                    const bool push = cone_child_test(env.o, env.d, sh.inv, sh.nx != 0, sh.ny != 0, sh.nz != 0, env.ta, env.x0, cr, mnx, mny, mnz, mxx, mxy, mxz, tmin) && ch != 0;
                    const unsigned m = g_ballot(gg, push);
                    const int np = __popc(m);
                    // rank in the order the insertion sort leaves the children on the stack (descending tmin, stable: bvh8w.cpp:44-57); pop order is the reverse
                    const float key = push ? tmin : -WT_INF;
                    int rank = 0;
#pragma unroll
                    for (int j = 0; j < 8; ++j) { const float kj = g_shfl(gg, key, j); rank += (kj > key || (kj == key && j < (int)gg.gl)) ? 1 : 0; }
                    uint32_t big = 0u;
                    if (push) {
                        const int pi = np - 1 - rank;
                        sh.wc_tmin[grp][pi] = tmin; sh.wc_ptr[grp][pi] = ch;
                        if (ch < 0) { const wtgpu_leaf lf = sc.leaves[-ch - 1]; sh.wc_t0[grp][pi] = lf.tris_ptr; sh.wc_cnt[grp][pi] = lf.count; big = lf.count > 8u ? 1u : 0u; }
                    }
                    const unsigned deep = g_ballot(gg, push && ch > 0), bigm = g_ballot(gg, big != 0u);
                    if (gg.gl == 0u) { sh.wn[grp] = np; sh.wfl[grp] = 1 | (deep ? 2 : 0) | (bigm ? 8 : 0); }
                }
            } else if (gg.gl == 0u) { sh.wn[grp] = 0; sh.wfl[grp] = 4; }
        }
        t_sync<TEAM>();

        // ---- B: the window = the leading run of simple entries; its leaves in pop order
        if (ctl) {
            const int w = (int)lane;
            const int fl = w < NW ? sh.wfl[w] : 4, wn = w < NW ? sh.wn[w] : 0;
            const bool simple = !(fl & (2 | 4 | 8)) && (!(fl & 1) || (s0 - 1 - w) + wn <= kGStack);
            const unsigned sm = __ballot_sync(FULL, simple);
            const int wlen = sm == FULL ? 32 : __ffs(~sm) - 1;
            const int cntw = w < wlen ? wn : 0;
            int incl = cntw;
#pragma unroll
            for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += u; }
            const int nl = __shfl_sync(FULL, incl, 31);
            if (w < wlen) { const int b = incl - cntw; sh.lbase[w] = b; for (int c = 0; c < cntw; ++c) { sh.bt0[b + c] = sh.wc_t0[w][c]; sh.bcnt[b + c] = sh.wc_cnt[w][c]; } }
            if (lane == 0u) { sh.wlen = wlen; sh.nl = nl; }
        }
        t_sync<TEAM>();
        const int wlen = sh.wlen, nl = sh.nl;

        if (wlen == 0) {
            // the top entry is not simple: the sequential algorithm's next step, by the control warp
            if (ctl) {
                const int fl = sh.wfl[0];
                if (fl & 1) {       // an internal node: expand it (the lookahead already holds its children when they fit the stack)
                    const int np = sh.wn[0];
                    if (s0 - 1 + np <= kGStack) {
                        if (lane == 0u) ctr.nodes++;
                        if ((int)lane < np) { const int at = s0 - 1 + (np - 1 - (int)lane); sh.g.tmin[at] = sh.wc_tmin[0][lane]; sh.g.ptr[at] = sh.wc_ptr[0][lane]; }
                        t.s = s0 - 1 + np;
                    } else {
                        const int32_t top = sh.g.ptr[s0 - 1];
                        t.s = s0 - 1;
                        if (lane < 8u) g_node_step(sc, g8, sh.g, t, top, ctr);
                        t.s = __shfl_sync(FULL, t.s, 0);
                    }
                    __syncwarp();
                } else {            // a leaf of more than eight triangles: the general leaf step, 32 triangles at a time
                    const uint32_t lt0 = sh.wc_t0[0][0], lcnt = sh.wc_cnt[0][0];
                    t.s = s0 - 1;
                    t.qtested += lcnt;
                    bool found = false;
                    for (uint32_t base = 0; base < lcnt; base += 32u) {
                        const uint32_t k = base + lane; const bool valid = k < lcnt; const uint32_t tuid = lt0 + k;
                        float d = WT_INF; bool front = false;
                        if (valid) { const Tri3 tr = load_tri(sc, tuid); ctr.tris++; d = intersect_cone_tri(t.env, t.frame, tr.a, tr.b, tr.c, tr.n, t.crange); front = dot(tr.n, -t.env.d) > 0.f; }
                        const bool acc = d < WT_INF && !(d > t.crange.mx);
                        const unsigned m = __ballot_sync(FULL, acc);
                        if (m) {
                            found = true;
                            const int wl = g_argmin32(d, acc && d < t.res.dist, lane);
                            if (wl >= 0) { t.res.dist = __shfl_sync(FULL, d, wl); t.res.front = __shfl_sync(FULL, front ? 1 : 0, wl) != 0; }
                            const uint32_t upto = t.res.n_tris + (uint32_t)__popc(m);
                            t_reserve(sc, t, upto, lane);
                            if (acc) tw_put(sc, t.tw, t.res.n_tris + (uint32_t)__popc(m & ((1u << lane) - 1u)), tuid);
                            t.res.n_tris = upto;
                            if (t.tw.fail) t.res.overflow = true;
                        }
                    }
                    if (found) { t.crange = cone_search_range(t.env, t.qrange, t.res.dist, t.zs); t_prune(sh.g, t); }
                    __syncwarp();
                }
            }
            continue;
        }

        // ---- T1: the cheap rejections for every triangle of the batch; survivors compacted
        {
            const int nslots = nl * 8;
            for (int base = 0; base < nslots; base += TEAM) {
                const int j = base + (int)tid;
                bool maybe = false;
                if (j < nslots) {
                    const int li = j >> 3; const uint32_t k = (uint32_t)j & 7u;
                    if (k < sh.bcnt[li]) { const Tri3 tr = load_tri(sc, sh.bt0[li] + k); maybe = cone_tri_maybe(env, frame, tr.a, tr.b, tr.c, cr); }
                    sh.dres[j] = WT_INF;
                }
                const unsigned m = __ballot_sync(FULL, maybe);
                if (m) {
                    int at = 0;
                    if (lane == (unsigned)(__ffs(m) - 1)) at = atomicAdd(&sh.nsurv, __popc(m));
                    at = __shfl_sync(FULL, at, __ffs(m) - 1);
                    if (maybe) sh.surv[at + __popc(m & ((1u << lane) - 1u))] = (uint16_t)j;
                }
            }
        }
        t_sync<TEAM>();
        // ---- T2: the full test for the survivors
        {
            const int ns = sh.nsurv;
            for (int q = (int)tid; q < ns; q += TEAM) {
                const int j = (int)sh.surv[q];
                const Tri3 tr = load_tri(sc, sh.bt0[j >> 3] + ((uint32_t)j & 7u));
                sh.dres[j] = intersect_cone_tri(env, frame, tr.a, tr.b, tr.c, tr.n, cr);
            }
        }
        t_sync<TEAM>();
        // ---- S: per leaf, the accepted slots and the first closest one (gather_tris' running minimum within the leaf)
        for (int li = (int)tid; li < nl; li += TEAM) {
            uint32_t m = 0u, arg = 0u; float dm = WT_INF;
#pragma unroll
            for (int k = 0; k < 8; ++k) { const float d = sh.dres[li * 8 + k]; if (d < WT_INF && !(d > cr.mx)) { m |= 1u << k; if (d < dm) { dm = d; arg = (uint32_t)k; } } }
            sh.lmask[li] = m; sh.ldmin[li] = dm; sh.larg[li] = arg;
        }
        t_sync<TEAM>();

        // ---- C: commit in stack order (control warp)
        if (ctl) {
            // fast path: no leaf improves the closest hit (so the range cannot change) and no window entry below the top is stale (so the
            // unwinding after a leaf with hits never removes one)
            bool slow = false;
            { const int w = (int)lane; if (__any_sync(FULL, w >= 1 && w < wlen && sh.g.tmin[s0 - 1 - w] >= t.crange.mx)) slow = true; }
            for (int c0 = 0; c0 < nl && !slow; c0 += 32) { const int li = c0 + (int)lane; if (__any_sync(FULL, li < nl && sh.lmask[li] != 0u && sh.ldmin[li] < t.res.dist)) slow = true; }
            if (!slow) {
                uint32_t ntri = 0u, nnode = 0u;
                for (int c0 = 0; c0 < nl; c0 += 32) {
                    const int li = c0 + (int)lane;
                    uint32_t m = li < nl ? sh.lmask[li] : 0u;
                    ntri += li < nl ? sh.bcnt[li] : 0u;
                    const uint32_t c = (uint32_t)__popc(m);
                    uint32_t incl = c;
#pragma unroll
                    for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl += u; }
                    const uint32_t tot = __shfl_sync(FULL, incl, 31);
                    if (tot) {
                        t_reserve(sc, t, t.res.n_tris + tot, lane);
                        uint32_t pos = t.res.n_tris + incl - c;
                        while (m) { const uint32_t k = (uint32_t)__ffs(m) - 1u; m &= m - 1u; tw_put(sc, t.tw, pos++, sh.bt0[li] + k); }
                        t.res.n_tris += tot;
                        if (t.tw.fail) t.res.overflow = true;
                    }
                }
                { const int w = (int)lane; nnode = (w < wlen && (sh.wfl[w] & 1)) ? 1u : 0u; }
                ctr.tris += ntri; ctr.nodes += nnode;
                t.qtested += __reduce_add_sync(FULL, ntri);
                const bool last_hit = sh.wn[wlen - 1] > 0 && sh.lmask[nl - 1] != 0u;
                t.s = s0 - wlen;
                if (last_hit) t_prune(sh.g, t);
            } else {
                // leaf by leaf.  last_hit: the last thing processed was a leaf with hits (the sequential loop unwinds stale entries right after it)
                bool last_hit = false, restarted = false;
                uint32_t ntri = 0u, nnode = 0u;
                for (int w = 0; w < wlen && !restarted; ++w) {
                    const int sidx = s0 - 1 - w;
                    if (w > 0 && last_hit && sh.g.tmin[sidx] >= t.crange.mx) continue;      // unwound, unvisited (last_hit stays: the unwinding goes on)
                    const int np = sh.wn[w];
                    if (sh.wfl[w] & 1) { ++nnode; last_hit = false; }
                    for (int c = 0; c < np; ++c) {
                        const int li = sh.lbase[w] + c;
                        ntri += sh.bcnt[li];
                        const uint32_t m = sh.lmask[li];
                        if (!m) { last_hit = false; continue; }
                        const uint32_t upto = t.res.n_tris + (uint32_t)__popc(m);
                        t_reserve(sc, t, upto, lane);
                        if (lane < 8u && ((m >> lane) & 1u)) tw_put(sc, t.tw, t.res.n_tris + (uint32_t)__popc(m & ((1u << lane) - 1u)), sh.bt0[li] + lane);
                        t.res.n_tris = upto;
                        if (t.tw.fail) t.res.overflow = true;
                        if (sh.ldmin[li] < t.res.dist) {
                            t.res.dist = sh.ldmin[li];
                            const Tri3 tr = load_tri(sc, sh.bt0[li] + sh.larg[li]);
                            t.res.front = dot(tr.n, -t.env.d) > 0.f;
                        }
                        last_hit = true;
                        const Range nr = cone_search_range(t.env, t.qrange, t.res.dist, t.zs);
                        if (nr.mx != t.crange.mx || nr.mn != t.crange.mn) {
                            // the rest of the batch saw a stale range: entries 0..w leave the stack, the uncommitted children of entry w go onto it
                            // (in stack order = reverse pop order), the unwinding runs, and the next batch tests what is left against the new range
                            t.crange = nr;
                            int ns = sidx;
                            __syncwarp();
                            for (int cc = np - 1; cc > c; --cc) { if (lane == 0u) { sh.g.tmin[ns] = sh.wc_tmin[w][cc]; sh.g.ptr[ns] = sh.wc_ptr[w][cc]; } ++ns; }
                            __syncwarp();
                            t.s = ns;
                            t_prune(sh.g, t);
                            restarted = true;
                            break;
                        }
                    }
                }
                if (lane == 0u) { ctr.tris += ntri; ctr.nodes += nnode; }
                t.qtested += ntri;
                if (!restarted) { t.s = s0 - wlen; if (last_hit) t_prune(sh.g, t); }
            }
            __syncwarp();
        }
    }
}

} // namespace wt
