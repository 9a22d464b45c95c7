// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_beam.h: beams -- restating include/wt/beam/{beam.hpp,beam_generic.hpp,beam_geometry.hpp},
// include/wt/interaction/common.hpp (footprint) and the surface record of include/wt/interaction/intersection.hpp.
// Lengths in metres, wavenumbers in 1/mm: k*length products carry the factor 1000 that mp-units inserts.
#pragma once
#include "ot_polar.h"
#include "../include/wtgpu.h"

namespace ot {

static constexpr f_t beam_cross_section_envelope = 3;      // gaussian_wavefront.hpp:26
static constexpr f_t major_axis_to_z_scale = 2;            // beam_generic.hpp:50

inline f_t k_times_len(f_t k, f_t len) { return k * len * 1000.f; }
inline f_t wavenum_to_wavelen(f_t k) { return (two_pi / k) * 0.001f; }     // quantity/math.hpp:31-33 (result in metres)

// interaction/common.hpp:18-35
struct footprint_t {
    v2 x{ 1, 0 };
    f_t la = 0, lb = 0;
    v2 a() const { return x * la; }
    v2 y() const { return { -x.y, x.x }; }
    v2 b() const { return y() * lb; }
};

// The parts of intersection_surface_t (interaction/intersection.hpp:34-160) the walk needs.
struct surface_t {
    v3 wp{};
    v2 uv{};
    v2 bary{};
    footprint_t footprint;
    uint32_t tuid = WTGPU_INVALID_IDX;      // (shape, mesh_tri_idx) <-> tuid
    bool has_shape = false;
    frame_t geo{}, shading{};
    const v3& ng() const { return geo.n; }
    const v3& ns() const { return shading.n; }
    // intersection.hpp:117-127
    v3 s_direction(v3 w) const {
        const v3 crs = cross(w, shading.n);
        const f_t l2 = length2(crs);
        const v3 ret = l2 < 1e-14f ? shading.t : crs / std::sqrt(l2);
        return dot(w, shading.n) < 0 ? -ret : ret;
    }
    // intersection.hpp:133-141
    frame_t sp_frame(v3 w) const {
        const v3 s = s_direction(w);
        const v3 p = cross(s, w);
        return { s, dot(w, shading.n) < 0 ? -p : p, w };
    }
};

// beam_geometry.hpp:32-180
struct phase_space_extent_t {
    f_t spatial_extent;     // m^2
    f_t tan_alpha;
    f_t k;
    phase_space_extent_t enlarge(f_t scale) const {
        if (scale == 1) return *this;
        return { spatial_extent * sqr(scale), tan_alpha * scale, k };
    }
    static constexpr f_t mub_sbp = 0.25f;
    static f_t minimum_uncertainty_tan_alpha(f_t spatial_length, f_t k) {     // beam_geometry.hpp:115-128
        return spatial_length > 0 ? std::sqrt(mub_sbp) * sqr(beam_cross_section_envelope) / k_times_len(k, spatial_length) : 0.f;
    }
    static f_t minimum_uncertainty_spatial_extent(f_t tan_alpha, f_t k) {      // beam_geometry.hpp:165-179
        const f_t spatial_length = tan_alpha > 0 ? (std::sqrt(mub_sbp) * sqr(beam_cross_section_envelope) / (k * tan_alpha)) * 0.001f : 0.f;
        return sqr(spatial_length);
    }
};

// beam_geometry.hpp:186-342 (surface-less sourcing only: no caller on the hot path sources from a surface)
struct sourcing_geometry_t {
    v3 x{ 1, 0, 0 };
    v2 initial_spatial_lengths{ 0, 0 };
    f_t tan_alpha = 0;
    f_t k = 0;
    phase_space_extent_t phase_space_extent() const { return { initial_spatial_lengths.x * initial_spatial_lengths.y, tan_alpha, k }; }
    elliptic_cone_t envelope(const ray_t& ray, f_t& sid) const {            // beam_geometry.hpp:209-231
        sid = 0;
        if (initial_spatial_lengths.x != initial_spatial_lengths.y) {
            const f_t ix = std::max(initial_spatial_lengths.x, initial_spatial_lengths.y);
            const f_t e = std::min(initial_spatial_lengths.x, initial_spatial_lengths.y) / ix;
            return elliptic_cone_t::make_ecc(ray, x, tan_alpha, e, ix);
        }
        return elliptic_cone_t::make_iso(ray, tan_alpha, initial_spatial_lengths.x);
    }
    static sourcing_geometry_t source_mub_from_tan_alpha(f_t tan_alpha, f_t k) {   // :236-245
        const f_t l = std::sqrt(phase_space_extent_t::minimum_uncertainty_spatial_extent(tan_alpha, k));
        sourcing_geometry_t g; g.initial_spatial_lengths = { l, l }; g.tan_alpha = tan_alpha; g.k = k; return g;
    }
    static sourcing_geometry_t source_mub_from_length(f_t l, f_t k) {               // :250-258
        sourcing_geometry_t g; g.initial_spatial_lengths = { l, l }; g.tan_alpha = phase_space_extent_t::minimum_uncertainty_tan_alpha(l, k); g.k = k; return g;
    }
    static sourcing_geometry_t source(f_t l, f_t tan_alpha, f_t k) {                // :300-310
        sourcing_geometry_t g; g.initial_spatial_lengths = { l, l }; g.tan_alpha = tan_alpha; g.k = k; return g;
    }
    static sourcing_geometry_t source(const phase_space_extent_t& e) {              // :316-324
        const f_t l = std::sqrt(e.spatial_extent);
        sourcing_geometry_t g; g.initial_spatial_lengths = { l, l }; g.tan_alpha = e.tan_alpha; g.k = e.k; return g;
    }
};

// beam.hpp:255-518 + beam_generic.hpp:38-194; one struct for both transports
struct beam_t {
    elliptic_cone_t envelope;
    f_t self_intersection_distance = 0;
    f_t k = 0;
    bool forward = true;
    stokes_t S;             // forward: Stokes vector in `frame`
    mueller_t M;            // backward: Mueller operator with incident frame `frame`, times `scale`
    f_t scale = 0;
    frame_t frame{};

    static beam_t make_forward(const ray_t& ray, f_t s, f_t k, const sourcing_geometry_t& sg) {   // beam.hpp:298-306
        beam_t b; b.forward = true; b.k = k;
        b.envelope = sg.envelope(ray, b.self_intersection_distance);
        b.S = stokes_t::unpolarized(s); b.frame = b.envelope.frame();
        return b;
    }
    static beam_t make_backward(const ray_t& ray, f_t scale, f_t k, const sourcing_geometry_t& sg) {  // beam.hpp:337-345
        beam_t b; b.forward = false; b.k = k;
        b.envelope = sg.envelope(ray, b.self_intersection_distance);
        b.M = mueller_t::identity(); b.frame = b.envelope.frame(); b.scale = scale;
        return b;
    }
    f_t intensity() const { return forward ? S.intensity() : M.mean_intensity() * scale; }
    const v3& dir() const { return envelope.d(); }
    const v3& origin() const { return envelope.o(); }
    bool is_ray() const { return envelope.is_ray(); }
    void mul(f_t f) { if (forward) S = S * f; else scale *= f; }             // operator*= (beam.hpp:105-108, 210-213)
    void div(f_t f) { if (forward) { for (auto& v : S.S) v /= f; } else scale /= f; }
    // operator+= (beam.hpp:95-98, 200-203, 482-485)
    void add(const beam_t& o) {
        if (forward) S = S + o.S.reorient(o.frame, frame);
        else M = M + change_incident_frame(o.M, o.frame, frame);
    }

    v3 footprint(f_t dist) const { const v2 a = envelope.axes(dist); return { a.x, a.y, major_axis_to_z_scale * a.x }; }   // beam_generic.hpp:114-117

    // beam_generic.hpp:171-193
    footprint_t surface_footprint_static(const surface_t& surface, f_t beam_z_dist) const {
        const v3 ls = footprint(beam_z_dist);
        const v3 x = surface.geo.to_local(envelope.x());
        if (x.x != 0 || x.y != 0) return { normalize(v2{ x.x, x.y }), ls.x, ls.y };
        const f_t avg = (ls.x + ls.y) / 2.f;
        return { { 1, 0 }, avg, avg };
    }

    // beam_radiometric_data_t::apply_bsdf (beam.hpp:53-66 forward, 163-173 backward)
    void apply_bsdf(const mueller_t& op, v3 wo, const surface_t& surface, const frame_t& frame_after) {
        if (forward) {
            const frame_t SPin = surface.sp_frame(frame.n), SPout = surface.sp_frame(wo);
            S = mueller_apply(op, S, frame, SPin, frame_after, SPout);
            frame = frame_after;
        } else {
            const frame_t SPin = surface.sp_frame(wo), SPout = surface.sp_frame(frame.n);
            M = compose(M, op, frame, SPout);
            frame = SPin;
        }
    }
    // beam.hpp:379-397
    void transform_surface_interaction(const surface_t& surface, v3 wo, const mueller_t& bsdfM, f_t weight) {
        const ray_t ray{ surface.wp, wo };
        f_t new_sid;
        // elliptic_cone_t::cone_through_ellipse(surface, ...) (elliptic_cone.hpp:283-301)
        const v3 wa = surface.geo.to_world(surface.footprint.a());
        const v3 wb = surface.geo.to_world(surface.footprint.b());
        envelope = elliptic_cone_t::cone_through_ellipse(wa, wb, surface.geo.n, ray, envelope.tan_alpha, &new_sid);
        apply_bsdf(weight * bsdfM, wo, surface, envelope.frame());
        self_intersection_distance = new_sid;
    }
    // beam.hpp:407-425
    void transform_region_interaction(v3 wp, f_t dist, v3 wo, f_t weight) {
        const v3 axes_local = footprint(dist);
        envelope = elliptic_cone_t::cone_through_ellipsoid(axes_local, envelope.frame(), ray_t{ wp, wo }, envelope.tan_alpha);
        mul(weight);
        frame = envelope.frame();
        self_intersection_distance = 0;
    }
    // beam.hpp:464-471
    void transform_restart(v3 wp, f_t dist) {
        envelope.set_o(wp);
        envelope.set_x0(envelope.x0 + dist * envelope.tan_alpha);
        self_intersection_distance = 0;
    }
};

// beam::integrate_beams (beam.hpp:562-603): S = detection (backward) beam, I = radiation (forward) beam
inline stokes_t integrate_beams(const beam_t& S, const beam_t& I) {
    if (S.intensity() == 0 || I.intensity() == 0) return {};
    return mueller_apply(S.M, I.S, I.frame, S.frame) * S.scale;
}

} // namespace ot
