// ctrav.cuh -- TEAM traversal: the cone queries the group traversal (gtrav.cuh) hands over because they run over hundreds to 10^5 triangles
// (a beam a few millimetres wide over a finely tessellated mesh).  (Included by wavefront.cu after gtrav.cuh.)
//
// A team of TEAM threads (one warp, or one 256-thread block) continues ONE beam at a time from the state the previous tier saved (TravSave:
// machine state + stack; nothing is redone).  The team's first warp is the control warp: it runs the same machine as gtrav.cuh (g_begin /
// g_query_done, ray queries, the sequential commit logic); the cone query's triangle work is done by the whole team in batches.
//
// The stack of the cone query lives in shared memory as Q (top = last entry).  A batch is built by LOOKAHEAD EXPANSION: the leading leaves on top
// of Q move, in pop order, to the run R; then the first nodes below them (up to TEAM/8, one 8-lane group each) are expanded where they stand --
// children tested against the current search range, ranked, written in pop order in place of the node, under a MARKER that keeps the node itself
// -- and the step repeats until R holds enough leaves.  Expanding a node before the sequential algorithm reaches it is speculation on one thing
// only: that the search range is still the same when it gets there (children, their order and their tmin depend on nothing else).  Then
//   T1 every triangle of R's leaves (up to 8 TEAM of them) goes through the cheap rejections of the cone-triangle test (cone_tri_maybe: z range,
//      separating axes); the survivors -- about one in seven -- are compacted;
//   T2 the survivors go through the full test (intersect_cone_tri), one per thread: the expensive code runs with full warps;
//   C  commit, in pop order, exactly what the sequential loop does (bvh8w.cpp:123-185, 245-347): a marker = "node visited"; a leaf = append the
//      accepted triangles, closest hit, search-range update, unwinding of stale entries.  If a commit NARROWS the range, everything after it was
//      computed against a stale range: the rest of R goes back on top of Q and every marker still in Q is REVERTED (its children dropped, the
//      node restored), which leaves exactly the sequential algorithm's stack; the next batch starts from there with the new range.  When no leaf of
//      the batch improves the closest hit (the common case once the hit distance has settled) the whole batch commits in one parallel pass.
// Entries that are stale (pushed under an older, wider range: tmin >= the current range's end) are never moved or expanded ahead of their turn:
// whether the sequential loop visits or unwinds them depends on whether the leaf before them had hits, which is only known at commit.
// The number of stack entries the sequential algorithm would hold is tracked per entry (ssz), so its 128-entry limit is honoured.
// Every decision is taken on the same values, in the same order, as the sequential code: lists, distances and counters are identical.
#pragma once

// resident blocks per SM the team kernels are compiled for (register budget): warp teams 4 x 128 threads (128 registers), block teams 2 x 256
#ifndef WT_WT_MINB
#define WT_WT_MINB 4
#endif
#ifndef WT_CT_MINB
#define WT_CT_MINB 2
#endif

namespace wt {

enum : uint32_t { QK_LEAF = 0u, QK_NODE = 1u, QK_MARK = 2u };
WT_D uint32_t q_info(uint32_t kind, uint32_t depth, uint32_t ssz) { return kind | (depth << 2) | (ssz << 8); }
WT_D uint32_t q_kind(uint32_t i) { return i & 3u; }
WT_D uint32_t q_depth(uint32_t i) { return (i >> 2) & 63u; }
WT_D uint32_t q_ssz(uint32_t i) { return i >> 8; }

template <int TEAM> struct alignas(16) TShared {
    static constexpr int NW = TEAM / 8;                     // nodes expanded per round (one 8-lane group each)
    static constexpr int QCAP = TEAM == 32 ? 256 : 1024;    // entries of Q (the sequential stack never exceeds kGStack; the rest is lookahead)
    static constexpr int NL = TEAM;                         // leaves per batch
    static constexpr int RCAP = 2 * TEAM;                   // entries of R (leaves + markers)
    GShared g;                                              // ray queries' stack; scratch of real node steps
    int32_t qptr[QCAP]; float qtmin[QCAP]; uint16_t qinfo[QCAP];
    int32_t rptr[RCAP]; float rtmin[RCAP]; uint16_t rinfo[RCAP]; uint16_t rleaf[NL];       // R in pop order; rleaf: R index of batch leaf li
    Cone env; Frame frame; Range crange; V3 inv; int nx, ny, nz;       // the query, published by the control warp for the batch phases
    int cmd, go, nsel, rn, nl, nsurv, big, ntri;
    unsigned keep[QCAP / 32 + 1];                           // q_revert: keep flags per 32-entry chunk
    int sel[NW];                                            // Q index of the nodes being expanded this round
    int wn[NW]; float wc_tmin[NW][8]; int32_t wc_ptr[NW][8];
#ifdef WT_NODE_STAGING
    alignas(16) float nstage[NW][64];                       // per expanding group: staging slot of its node record (gtrav.cuh g_stage_node)
#endif
    uint32_t bt0[NL], bcnt[NL];                             // batch leaves: triangle range
    float dres[NL * 8]; uint16_t surv[NL * 8]; uint16_t tslot[NL * 8];      // per triangle slot (leaf * 8 + k): test result; the slots that survived T1; the slots that hold a triangle
    uint32_t lmask[NL], larg[NL]; float ldmin[NL];          // per batch leaf: accepted slots, first closest slot, its distance
};
template <int TEAM> WT_D void t_sync() { if (TEAM == 32) __syncwarp(); else __syncthreads(); }
enum { TC_DONE = 0, TC_BATCH = 1 };

// lane-0 reservation of list storage up to `upto`, result broadcast over the control warp
WT_D void t_reserve(const DScene& sc, GTrav& t, uint32_t upto, unsigned lane) {
    if (upto > t.tw.alloc_end && !t.tw.fail) {
        if (lane == 0u) tw_reserve(sc, t.tw, upto);
        t.tw.n_ext = __shfl_sync(0xffffffffu, t.tw.n_ext, 0); t.tw.alloc_end = __shfl_sync(0xffffffffu, t.tw.alloc_end, 0); t.tw.fail = __shfl_sync(0xffffffffu, t.tw.fail ? 1 : 0, 0) != 0;
        __syncwarp();
    }
}
// Control warp (all lanes, same arguments).  Reverts every marker of Q[0, qs): its children are dropped, the node comes back -- Q becomes the
// stack of the sequential algorithm.  An entry is the child of a marker still in Q exactly when some entry ABOVE it is shallower (its own
// ancestors' markers are popped before it; everything else above it belongs to subtrees that precede it at the same or a greater depth), so:
// pass 1, top down, a running minimum of the depths above each entry decides keep / drop; pass 2, bottom up, compacts.  Returns the new size.
template <int TEAM> WT_D int q_revert(TShared<TEAM>& sh, int qs, unsigned lane) {
    const unsigned FULL = 0xffffffffu;
    __syncwarp();
    if (qs <= 0) return 0;
    uint32_t above = 64u;           // minimum depth of the entries above the chunk in hand
    for (int hi = qs; hi > 0; hi -= 32) {
        const int r = hi - 1 - (int)lane;           // lane 0 = the chunk's top entry
        const uint32_t d = r >= 0 ? q_depth(sh.qinfo[r]) : 64u;
        uint32_t incl = d;                          // minimum over this lane and the lanes above it (smaller lane index)
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const uint32_t u = __shfl_up_sync(FULL, incl, o); if ((int)lane >= o) incl = min(incl, u); }
        uint32_t excl = __shfl_up_sync(FULL, incl, 1); if (lane == 0u) excl = 64u;
        const bool keep = r >= 0 && !(min(above, excl) < d);
        const unsigned km = __ballot_sync(FULL, keep);
        if (lane == 0u) sh.keep[(hi - 1) >> 5] = km;          // bit l <-> entry hi - 1 - l
        above = min(above, __shfl_sync(FULL, incl, 31));
    }
    __syncwarp();
    int w = 0;
    // the chunks were cut from the top: [hi - 32, hi) with hi = qs, qs - 32, ...; walk them bottom up
    for (int hi = qs - ((qs - 1) / 32) * 32; hi <= qs; hi += 32) {
        const unsigned km = sh.keep[(hi - 1) >> 5];
        // ascending index within the chunk = descending lane: entry r = hi - 1 - l; position among the kept = number of kept entries with a smaller index = kept lanes above l
        const int r = hi - 1 - (int)lane;
        const bool keep = (km >> lane) & 1u;
        int32_t p = 0; float tm = 0.f; uint32_t inf = 0u;
        if (keep) { p = sh.qptr[r]; tm = sh.qtmin[r]; inf = sh.qinfo[r]; }
        __syncwarp();
        if (keep) {
            const int at = w + __popc(km & ~((2u << lane) - 1u));          // kept lanes with a larger lane index = smaller entry index
            if (q_kind(inf) == QK_MARK) inf = q_info(QK_NODE, q_depth(inf), q_ssz(inf));
            sh.qptr[at] = p; sh.qtmin[at] = tm; sh.qinfo[at] = (uint16_t)inf;
        }
        w += __popc(km);
        __syncwarp();
    }
    return w;
}
// the unwinding after a leaf with hits (bvh8w.cpp:262-267): stale entries leave the top of the stack (a marker stands for a node that was not stale)
template <int TEAM> WT_D void q_prune(const TShared<TEAM>& sh, GTrav& t) { while (t.s > 0 && q_kind(sh.qinfo[t.s - 1]) != QK_MARK && sh.qtmin[t.s - 1] >= t.crange.mx) --t.s; }
// the sequential stack in sh.g (as saved / as the group code leaves it) -> Q
template <int TEAM> WT_D void q_from_stack(TShared<TEAM>& sh, int s, unsigned lane) {
    for (int k = (int)lane; k < s; k += 32) { const int32_t p = sh.g.ptr[k]; sh.qptr[k] = p; sh.qtmin[k] = sh.g.tmin[k]; sh.qinfo[k] = (uint16_t)q_info(p < 0 ? QK_LEAF : QK_NODE, 0u, (uint32_t)k + 1u); }
    __syncwarp();
}

#ifdef WT_TEAM_DEBUG
#define TDBG(i, v) do { if (ctl && lane == 0u) d_[i] += (unsigned long long)(v); } while (0)
#define TCLK(i) do { if (ctl && lane == 0u) { const long long c_ = clock64(); d_[i] += (unsigned long long)(c_ - clk_); clk_ = c_; } } while (0)
#else
#define TDBG(i, v) do { } while (0)
#define TCLK(i) do { } while (0)
#endif
template <int TEAM, class Emit>
WT_D void t_traverse_all(const DScene& sc, int n_items, const TravSave* saves, int* cursor, TShared<TEAM>& sh, Counters& ctr,
                         TravSave* huge_save, int* n_huge, TravTiers tiers, unsigned long long* dbg, Emit&& emit) {
    const uint32_t huge_tested = tiers.huge_tested;
    constexpr int NW = TShared<TEAM>::NW, QCAP = TShared<TEAM>::QCAP, NL = TShared<TEAM>::NL, RCAP = TShared<TEAM>::RCAP;
    const unsigned FULL = 0xffffffffu;
    const unsigned tid = threadIdx.x % (unsigned)TEAM, lane = threadIdx.x & 31u;
    const bool ctl = tid < 32u;
    GLane gw; gw.gl = lane; gw.gshift = 0u; gw.gmask = FULL;                                    // the control warp as one "group" (set-up / bookkeeping code of gtrav.cuh)
    GLane g8; g8.gl = lane & 7u; g8.gshift = 0u; g8.gmask = 0xffu;                              // its lanes 0-7: real node steps
    GLane gg; gg.gl = lane & 7u; gg.gshift = lane & 24u; gg.gmask = 0xffu << gg.gshift;         // this thread's group of eight (lookahead)
    const int grp = (int)(tid >> 3);
    GTrav t; t.mode = 0; t.s = 0;          // in a cone query t.s is the size of Q
#ifdef WT_TEAM_DEBUG
    unsigned long long d_[8] = { 0, 0, 0, 0, 0, 0, 0, 0 }; long long clk_ = clock64();      // 0 batches, 1 leaves, 2 rounds, 3 slow commits | cycles: 4 control, 5 lookahead, 6 tests, 7 commit
#endif
    bool have = false; int item = 0;
    const int nl_init = max(NL / (int)max(tiers.init_div, 1u), 8), nl_restart = max(NL / (int)max(tiers.restart_div, 1u), 8);
    int nl_cap = nl_init;           // leaves per batch: small after the search range has changed (what follows such a leaf is thrown away), doubling while it holds
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
                    q_from_stack(sh, t.s, lane);        // (items arrive in the middle of a cone query)
                    have = true; nl_cap = nl_init;
                }
                if (t.s == 0) {
                    TravRec out;
                    if (g_query_done(sc, gw, sh.g, t, out, ctr)) { emit(item, out, gw); have = false; __syncwarp(); }
                    else if (t.mode == 2) q_from_stack(sh, t.s, lane);
                    continue;
                }
                if (t.mode == 1) {      // a ray query (ballistic segment): the whole of it in the control warp
                    const int32_t top = sh.g.ptr[t.s - 1];
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
                    t.s = q_revert(sh, t.s, lane);
                    int pos = 0;
                    if (lane == 0u) pos = atomicAdd(n_huge, 1);
                    pos = __shfl_sync(FULL, pos, 0);
                    TravSave& sv = huge_save[pos];
                    if (lane == 0u) { sv.t = t; sv.item = item; }
                    for (int k = (int)lane; k < t.s; k += 32) { sv.tmin[k] = sh.qtmin[k]; sv.ptr[k] = sh.qptr[k]; }
                    have = false; t.s = 0; t.mode = 0;
                    __syncwarp();
                    continue;
                }
                break;
            }
            if (lane == 0u) {
                sh.cmd = cmd; sh.nsurv = 0; sh.ntri = 0; sh.rn = 0; sh.nl = 0; sh.big = 0x7fffffff;
                if (cmd == TC_BATCH) { sh.env = t.env; sh.frame = t.frame; sh.crange = t.crange; sh.inv = t.inv; sh.nx = t.nx ? 1 : 0; sh.ny = t.ny ? 1 : 0; sh.nz = t.nz ? 1 : 0; }
            }
        }
        t_sync<TEAM>();
        TCLK(4);
        if (sh.cmd == TC_DONE) break;
        const Cone env = sh.env; const Frame frame = sh.frame; const Range cr = sh.crange;

        // ---- lookahead: leading leaves / markers of Q -> R; the first nodes below them expanded in place; repeat
        int rn = 0, nl = 0;         // (control warp) entries / leaves of R
        int myk = -1;               // (control warp) which of the round's selected nodes this lane's entry is
        for (int round = 0; ; ++round) {
            if (ctl) {
                // (1) the leading run of Q moves to R.  A stale entry moves only when it is the very next thing the sequential loop pops (R empty).
                bool room = true;
                while (room && t.s > 0) {
                    const int qi = t.s - 1 - (int)lane;
                    uint32_t inf = QK_NODE; float tm = 0.f; int32_t p = 0;
                    if (qi >= 0) { inf = sh.qinfo[qi]; tm = sh.qtmin[qi]; p = sh.qptr[qi]; }
                    const uint32_t kd = q_kind(inf);
                    const bool stale = kd != QK_MARK && tm >= cr.mx;
                    const bool mov = qi >= 0 && kd != QK_NODE && !(stale && (lane > 0u || rn > 0));
                    const unsigned mm = __ballot_sync(FULL, mov);
                    int take = mm == FULL ? 32 : __ffs(~mm) - 1;                // leading lanes that move
                    const unsigned lm = __ballot_sync(FULL, mov && kd == QK_LEAF);
                    // capacity: R entries and batch leaves
                    if (take > RCAP - rn) { take = RCAP - rn; room = false; }
                    int nleaf = __popc(lm & (take >= 32 ? FULL : ((1u << take) - 1u)));
                    while (nl + nleaf > nl_cap) { --take; room = false; nleaf = __popc(lm & ((1u << take) - 1u)); }
                    if ((int)lane < take) {
                        sh.rptr[rn + lane] = p; sh.rtmin[rn + lane] = tm; sh.rinfo[rn + lane] = (uint16_t)inf;
                        if (kd == QK_LEAF) sh.rleaf[nl + __popc(lm & ((1u << lane) - 1u))] = (uint16_t)(rn + lane);
                    }
                    rn += take; nl += nleaf; t.s -= take;
                    if (take < 32) break;
                }
                // (2) the nodes to expand: the first NW eligible ones among the top 32 entries, up to the first entry that has to wait its turn
                int nsel = 0; myk = -1;
                const bool want = room && nl < (nl_cap * 3) / 4 && round < 6 && t.s > 0 && t.s + rn + 9 * NW <= QCAP;       // (Q must be able to take R back: t.s + rn never exceeds QCAP)
                if (want) {
                    const int qi = t.s - 1 - (int)lane;
                    uint32_t inf = QK_LEAF; float tm = 0.f;
                    if (qi >= 0) { inf = sh.qinfo[qi]; tm = sh.qtmin[qi]; }
                    const uint32_t kd = q_kind(inf);
                    const bool stale = kd != QK_MARK && tm >= cr.mx && (lane > 0u || rn > 0);
                    const bool isnode = qi >= 0 && kd == QK_NODE;
                    const bool stop = qi < 0 || stale || (isnode && q_ssz(inf) + 7u > (uint32_t)kGStack);
                    const unsigned sm = __ballot_sync(FULL, stop);
                    const unsigned before = sm ? ((1u << (__ffs(sm) - 1)) - 1u) : FULL;
                    const unsigned nm = __ballot_sync(FULL, isnode) & before;
                    const int k = __popc(nm & ((1u << lane) - 1u));
                    if (((nm >> lane) & 1u) && k < NW) { sh.sel[k] = qi; myk = k; }
                    nsel = min(__popc(nm), NW);
                }
                if (lane == 0u) { sh.nsel = nsel; sh.go = nsel > 0 ? 1 : 0; }
            }
            t_sync<TEAM>();
            if (!sh.go) break;
            TDBG(2, 1);
            // (3) one group of eight lanes per selected node: children against the current range, ranked in pop order
            if (grp < sh.nsel) {
                const int32_t ptr = sh.qptr[sh.sel[grp]];
                float mnx, mny, mnz, mxx, mxy, mxz; int32_t ch;
#ifdef WT_NODE_STAGING
                float* const nslot = sh.nstage[grp];
#else
                float* const nslot = nullptr;
#endif
                g_stage_node(sc, gg, nslot, ptr, mnx, mny, mnz, mxx, mxy, mxz, ch);
                float tmin;
                const bool push = cone_child_test(env.o, env.d, sh.inv, sh.nx != 0, sh.ny != 0, sh.nz != 0, env.ta, env.x0, cr, mnx, mny, mnz, mxx, mxy, mxz, tmin) && ch != 0;
                const unsigned m = g_ballot(gg, push);
                const int np = __popc(m);
                // rank in the order the insertion sort leaves the children on the stack (descending tmin, stable: bvh8w.cpp:44-57); pop order is the reverse
                const float key = push ? tmin : -WT_INF;
                int rank = 0;
#pragma unroll
                for (int j = 0; j < 8; ++j) { const float kj = g_shfl(gg, key, j); rank += (kj > key || (kj == key && j < (int)gg.gl)) ? 1 : 0; }
                if (push) { const int pi = np - 1 - rank; sh.wc_tmin[grp][pi] = tmin; sh.wc_ptr[grp][pi] = ch; }
                if (gg.gl == 0u) sh.wn[grp] = np;
            }
            t_sync<TEAM>();
            // (4) rewrite the top of Q: every selected node becomes [children in stack order ..., marker]
            if (ctl) {
                const int nsel = sh.nsel;
                const int lowest = sh.sel[nsel - 1];            // the deepest selected entry: the region is [lowest, t.s)
                const int qi = t.s - 1 - (int)lane;
                const bool inreg = qi >= lowest;
                uint32_t inf = 0u; float tm = 0.f; int32_t p = 0; const int k = inreg ? myk : -1;      // (this lane looked at the same entry when it selected)
                if (inreg) { inf = sh.qinfo[qi]; tm = sh.qtmin[qi]; p = sh.qptr[qi]; }
                const int np = k >= 0 ? sh.wn[k] : 0;
                const int size = inreg ? (k >= 0 ? np + 1 : 1) : 0;
                // entries deeper in the region (larger lane) come first in the new layout: offset = sum of the sizes of the lanes above this one
                int incl = size;
#pragma unroll
                for (int o = 1; o < 32; o <<= 1) { const int u = __shfl_down_sync(FULL, incl, o); if ((int)lane + o < 32) incl += u; }
                const int total = __shfl_sync(FULL, incl, 0);
                const int at = lowest + incl - size;
                __syncwarp();
                if (inreg) {
                    if (k < 0) { sh.qptr[at] = p; sh.qtmin[at] = tm; sh.qinfo[at] = (uint16_t)inf; }
                    else {
                        const uint32_t d = q_depth(inf), ssz = q_ssz(inf);
                        for (int c = 0; c < np; ++c) {      // child with pop index c sits at at + (np - 1 - c); when it is on top the sequential stack holds ssz - 1 + (np - c) entries
                            const int32_t cp = sh.wc_ptr[k][c];
                            sh.qptr[at + np - 1 - c] = cp; sh.qtmin[at + np - 1 - c] = sh.wc_tmin[k][c];
                            sh.qinfo[at + np - 1 - c] = (uint16_t)q_info(cp < 0 ? QK_LEAF : QK_NODE, min(d + 1u, 63u), ssz - 1u + (uint32_t)(np - c));
                        }
                        sh.qptr[at + np] = p; sh.qtmin[at + np] = tm; sh.qinfo[at + np] = (uint16_t)q_info(QK_MARK, d, ssz);
                    }
                }
                t.s = lowest + total;
                __syncwarp();
            }
        }
        if (ctl && lane == 0u) { sh.rn = rn; sh.nl = nl; }
        t_sync<TEAM>();
        rn = sh.rn; nl = sh.nl;
        TCLK(5); TDBG(0, 1); TDBG(1, nl);

        // ---- B: the batch leaves' triangle ranges
        for (int li = (int)tid; li < nl; li += TEAM) {
            const wtgpu_leaf lf = sc.leaves[-sh.rptr[sh.rleaf[li]] - 1];
            sh.bt0[li] = lf.tris_ptr; sh.bcnt[li] = lf.count;
            if (lf.count > 8u) atomicMin(&sh.big, li);
            else {      // the leaf's occupied slots join the batch's triangle list (any order: results are stored per slot)
                const int at = atomicAdd(&sh.ntri, (int)lf.count);
                for (uint32_t k = 0; k < lf.count; ++k) sh.tslot[at + (int)k] = (uint16_t)(li * 8 + (int)k);
            }
        }
        t_sync<TEAM>();
        const int big = sh.big;     // (read by every thread before anything below can lead the control warp back to the top of the loop, where it is reset)
        if (rn == 0 || big == 0) {
            // nothing could be moved to R: the top of Q is a node that cannot be expanded in place (the sequential stack is nearly full, or Q
            // is) -- or the first leaf holds more than eight triangles.  The sequential algorithm's next step, by the control warp, on the real stack.
            if (ctl) {
                // R goes back (nothing of it was committed), the markers are reverted
                for (int e = rn - 1 - (int)lane; e >= 0; e -= 32) { const int at = t.s + (rn - 1 - e); sh.qptr[at] = sh.rptr[e]; sh.qtmin[at] = sh.rtmin[e]; sh.qinfo[at] = sh.rinfo[e]; }
                t.s = q_revert(sh, t.s + rn, lane);
                for (int k = (int)lane; k < t.s; k += 32) { sh.g.tmin[k] = sh.qtmin[k]; sh.g.ptr[k] = sh.qptr[k]; }
                __syncwarp();
                const int32_t top = sh.g.ptr[t.s - 1];
                --t.s;
                if (top >= 0) {
                    if (lane < 8u) g_node_step(sc, g8, sh.g, t, top, ctr);
                    t.s = __shfl_sync(FULL, t.s, 0);
                    __syncwarp();
                } else {            // the general leaf step, 32 triangles at a time
                    const wtgpu_leaf lf = sc.leaves[-top - 1];
                    const uint32_t lt0 = lf.tris_ptr, lcnt = lf.count;
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
                    if (found) { t.crange = cone_search_range(t.env, t.qrange, t.res.dist, t.zs); while (t.s > 0 && sh.g.tmin[t.s - 1] >= t.crange.mx) --t.s; }
                    __syncwarp();
                }
                q_from_stack(sh, t.s, lane);
            }
            t_sync<TEAM>();         // (no thread is still deciding on `big` when the control warp starts the next step)
            continue;
        }
        if (big < nl) {             // a leaf of more than eight triangles further down the run: the batch ends before it
            if (ctl) {
                const int cut = (int)sh.rleaf[big];
                for (int e = rn - 1 - (int)lane; e >= cut; e -= 32) { const int at = t.s + (rn - 1 - e); sh.qptr[at] = sh.rptr[e]; sh.qtmin[at] = sh.rtmin[e]; sh.qinfo[at] = sh.rinfo[e]; }
                t.s += rn - cut;
                __syncwarp();
            }
            nl = big; rn = (int)sh.rleaf[big];
        }

        // ---- T1: the cheap rejections for every triangle of the batch; survivors compacted
        {
            const int ntri = sh.ntri;
            for (int base = 0; base < ntri; base += TEAM) {
                const int tq = base + (int)tid;
                bool maybe = false; int j = 0;
                if (tq < ntri) {
                    j = (int)sh.tslot[tq];
                    const int li = j >> 3;
                    if (li < nl) {      // (a batch cut at a leaf of more than eight triangles: the leaves behind it wait)
                        const Tri3 tr = load_tri(sc, sh.bt0[li] + ((uint32_t)j & 7u)); maybe = cone_tri_maybe(env, frame, tr.a, tr.b, tr.c, cr);
                        sh.dres[j] = WT_INF;
                    }
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
            for (int k = 0; k < 8; ++k) { const float d = k < (int)sh.bcnt[li] ? sh.dres[li * 8 + k] : WT_INF; if (d < WT_INF && !(d > cr.mx)) { m |= 1u << k; if (d < dm) { dm = d; arg = (uint32_t)k; } } }
            sh.lmask[li] = m; sh.ldmin[li] = dm; sh.larg[li] = arg;
        }
        t_sync<TEAM>();

        TCLK(6);
        // ---- C: commit R in pop order (control warp).  Leaves are committed in SEGMENTS that end at the first leaf improving the closest hit: up
        // to there nothing the commit depends on changes, so the segment's accepted triangles are appended in one parallel pass; the improving
        // leaf then moves the hit distance, and if that narrows the search range the rest of the batch is void (it saw the old range).
        if (ctl) {
            int start = 0, e0 = 0;          // first uncommitted leaf / R entry
            bool restarted = false;
            uint32_t ntri = 0u;
            for (;;) {
                int first = nl;
                for (int c0 = start; c0 < nl; c0 += 32) {
                    const int li = c0 + (int)lane;
                    const unsigned im = __ballot_sync(FULL, li < nl && sh.lmask[li] != 0u && sh.ldmin[li] < t.res.dist);
                    if (im) { first = c0 + __ffs(im) - 1; break; }
                }
                const int upto = first < nl ? first + 1 : nl;
                for (int c0 = start; c0 < upto; c0 += 32) {
                    const int li = c0 + (int)lane;
                    uint32_t m = li < upto ? sh.lmask[li] : 0u;
                    ntri += li < upto ? sh.bcnt[li] : 0u;
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
                const int e1 = first < nl ? (int)sh.rleaf[first] + 1 : rn;         // R entries [e0, e1) are done: e1 - e0 - (upto - start) of them are markers (nodes visited)
                if (lane == 0u) ctr.nodes += (uint32_t)((e1 - e0) - (upto - start));
                if (first >= nl) break;
                t.res.dist = sh.ldmin[first];
                { const Tri3 tr = load_tri(sc, sh.bt0[first] + sh.larg[first]); t.res.front = dot(tr.n, -t.env.d) > 0.f; }
                const Range nr = cone_search_range(t.env, t.qrange, t.res.dist, t.zs);
                if (nr.mx != t.crange.mx || nr.mn != t.crange.mn) {
                    // the rest of R goes back on top of Q, the markers are reverted -- Q is the sequential stack again -- and the unwinding
                    // that follows a leaf with hits runs with the new range
                    TDBG(3, 1);
                    t.crange = nr;
                    const int e = e1 - 1, back = rn - e1;
                    for (int x = rn - 1 - (int)lane; x > e; x -= 32) { const int at = t.s + (rn - 1 - x); sh.qptr[at] = sh.rptr[x]; sh.qtmin[at] = sh.rtmin[x]; sh.qinfo[at] = sh.rinfo[x]; }
                    t.s = q_revert(sh, t.s + back, lane);
                    q_prune(sh, t);
                    restarted = true;
                    nl_cap = nl_restart;
                    break;
                }
                start = upto; e0 = e1;
            }
            ctr.tris += ntri;
            t.qtested += __reduce_add_sync(FULL, ntri);
            if (!restarted) {
                const bool last_hit = nl > 0 && (int)sh.rleaf[nl - 1] == rn - 1 && sh.lmask[nl - 1] != 0u;      // the last thing popped was a leaf with hits
                if (last_hit) q_prune(sh, t);
                nl_cap = min(NL, nl_cap * 2);
            }
            __syncwarp();
        }
        TCLK(7);
    }
#ifdef WT_TEAM_DEBUG
    if (ctl && lane == 0u) for (int i = 0; i < 8; ++i) atomicAdd(dbg + i, d_[i]);
#endif
}

} // namespace wt
