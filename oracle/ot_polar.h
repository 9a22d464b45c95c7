// ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path.
//
// ot_polar.h: Stokes vectors, Mueller operators, Fresnel -- restating
//   include/wt/interaction/polarimetric/stokes.hpp, mueller.hpp, include/wt/interaction/fresnel.hpp.
// Mueller matrices keep glm's column-major storage (M.m[col][row]) so the reference's constructor /
// transpose idiosyncrasies (mueller.hpp:36-46, 244-258, 294-309, 370-396) carry over literally.
#pragma once
#include "ot_math.h"

namespace ot {

struct stokes_t {
    f_t S[4] = { 0, 0, 0, 0 };
    f_t intensity() const { return S[0]; }
    bool is_unpolarized() const { return S[1] == 0 && S[2] == 0 && S[3] == 0; }
    bool isfinite() const { return std::isfinite(S[0]) && std::isfinite(S[1]) && std::isfinite(S[2]) && std::isfinite(S[3]); }
    static stokes_t unpolarized(f_t I) { return { { I, 0, 0, 0 } }; }
    stokes_t flip_handness() const { return { { S[0], S[1], -S[2], -S[3] } }; }
    // stokes.hpp:152-175
    stokes_t reorient(const frame_t& cur, const frame_t& nw) const {
        const v3 tl = cur.to_local(nw.t), bl = cur.to_local(nw.b);
        const v2 tou{ tl.x, tl.y }, tov{ bl.x, bl.y };
        const mat2 R = rotation_matrix(v2{ 1, 0 }, tou);
        const v2 S12 = R * (R * v2{ S[1], S[2] });
        const stokes_t r{ { S[0], S12.x, S12.y, S[3] } };
        const v2 v = R * v2{ 0, 1 };
        if (dot(v, tov) < 0) return r.flip_handness();
        return r;
    }
};
inline stokes_t operator*(const stokes_t& s, f_t f) { return { { s.S[0] * f, s.S[1] * f, s.S[2] * f, s.S[3] * f } }; }
inline stokes_t operator+(const stokes_t& a, const stokes_t& b) { return { { a.S[0] + b.S[0], a.S[1] + b.S[1], a.S[2] + b.S[2], a.S[3] + b.S[3] } }; }

struct mueller_t {
    f_t m[4][4];    // m[col][row]
    mueller_t() { std::memset(m, 0, sizeof(m)); }
    // 16-scalar glm ctor fills columns (mueller.hpp:36-46)
    mueller_t(f_t a0, f_t a1, f_t a2, f_t a3, f_t b0, f_t b1, f_t b2, f_t b3, f_t c0, f_t c1, f_t c2, f_t c3, f_t d0, f_t d1, f_t d2, f_t d3)
        : m{ { a0, a1, a2, a3 }, { b0, b1, b2, b3 }, { c0, c1, c2, c3 }, { d0, d1, d2, d3 } } {}
    f_t mean_intensity() const { return m[0][0]; }
    static mueller_t identity() { return mueller_t(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1, 0, 0, 0, 0, 1); }
    static mueller_t handness_flip() { return mueller_t(1, 0, 0, 0, 0, 1, 0, 0, 0, 0, -1, 0, 0, 0, 0, -1); }
    static mueller_t perfect_depolarizer() { mueller_t P; P.m[0][0] = 1; return P; }
    mueller_t transposed() const { mueller_t r; for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.m[c][k] = m[k][c]; return r; }
    // mueller.hpp:244-258
    static mueller_t rotation(v2 t1, v2 t2) {
        mat2 R = rotation_matrix(t1, t2);
        R = R * R;
        mueller_t T;
        T.m[0][0] = T.m[3][3] = 1;
        T.m[1][1] = R.m[0][0];
        T.m[2][1] = R.m[1][0];
        T.m[1][2] = R.m[0][1];
        T.m[2][2] = R.m[1][1];
        return T.transposed();
    }
    // mueller.hpp:294-309
    static mueller_t fresnel(c_t fs, c_t fp) {
        const f_t Rs = std::norm(fs), Rp = std::norm(fp);
        const f_t m00 = (Rs + Rp) / 2.f, m01 = (Rs - Rp) / 2.f;
        const c_t pc = fp * std::conj(fs);
        const f_t m22 = pc.real(), m23 = pc.imag();
        return mueller_t(m00, m01, 0, 0, m01, m00, 0, 0, 0, 0, m22, m23, 0, 0, -m23, m22).transposed();
    }
};
inline mueller_t operator*(const mueller_t& A, f_t s) { mueller_t r; for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.m[c][k] = A.m[c][k] * s; return r; }
inline mueller_t operator*(f_t s, const mueller_t& A) { mueller_t r; for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.m[c][k] = s * A.m[c][k]; return r; }
inline mueller_t operator/(const mueller_t& A, f_t s) { mueller_t r; for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.m[c][k] = A.m[c][k] / s; return r; }
inline mueller_t operator+(const mueller_t& A, const mueller_t& B) { mueller_t r; for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k) r.m[c][k] = A.m[c][k] + B.m[c][k]; return r; }
// glm mat4*mat4: result[c][r] = sum_k A[k][r]*B[c][k], accumulated left to right
inline mueller_t operator*(const mueller_t& A, const mueller_t& B) {
    mueller_t r;
    for (int c = 0; c < 4; ++c) for (int k = 0; k < 4; ++k)
        r.m[c][k] = A.m[0][k] * B.m[c][0] + A.m[1][k] * B.m[c][1] + A.m[2][k] * B.m[c][2] + A.m[3][k] * B.m[c][3];
    return r;
}
// mueller.hpp:130-146: M*S = dot(row_i(M), S) with the fma-chain dot
inline stokes_t operator*(const mueller_t& M, const stokes_t& S) {
    stokes_t r;
    for (int i = 0; i < 4; ++i)
        r.S[i] = std::fma(M.m[3][i], S.S[3], std::fma(M.m[2][i], S.S[2], std::fma(M.m[1][i], S.S[1], M.m[0][i] * S.S[0])));
    return r;
}
// mueller.hpp:155-165 (3-arg operator())
inline stokes_t mueller_apply(const mueller_t& M, const stokes_t& S, const frame_t& Sin, const frame_t& Min) {
    if (S.is_unpolarized()) return M * S;
    return M * S.reorient(Sin, Min);
}
// mueller.hpp:175-185 (5-arg operator())
inline stokes_t mueller_apply(const mueller_t& M, const stokes_t& S, const frame_t& Sin, const frame_t& Min, const frame_t& Sout, const frame_t& Mout) {
    const stokes_t r = M * S.reorient(Sin, Min);
    return r.reorient(Mout, Sout);
}
// mueller.hpp:191-203
inline mueller_t change_incident_frame(const mueller_t& M, const frame_t& oldf, const frame_t& newf) {
    const v3 tl = oldf.to_local(newf.t);
    mueller_t R = mueller_t::rotation(v2{ tl.x, tl.y }, v2{ 1, 0 });
    if (oldf.handness() != newf.handness()) R = R * mueller_t::handness_flip();
    return M * R;
}
// mueller.hpp:204-216
inline mueller_t change_exitant_frame(const mueller_t& M, const frame_t& oldf, const frame_t& newf) {
    const v3 tl = oldf.to_local(newf.t);
    mueller_t R = mueller_t::rotation(v2{ tl.x, tl.y }, v2{ 1, 0 });
    if (oldf.handness() != newf.handness()) R = R * mueller_t::handness_flip();
    return R * M;
}
// mueller.hpp:402-415
inline mueller_t compose(const mueller_t& M1, const mueller_t& M2, const frame_t& M1in, const frame_t& M2out) {
    const v3 tl = M1in.to_local(M2out.t);
    mueller_t R = mueller_t::rotation(v2{ tl.x, tl.y }, v2{ 1, 0 });
    if (M1in.handness() != M2out.handness()) R = mueller_t::handness_flip() * R;
    return M1 * R * M2;
}

// ---- Fresnel: include/wt/interaction/fresnel.hpp
inline v3 reflect(v3 w, v3 n = { 0, 0, 1 }) { return 2.f * (dot(w, n) * n) - w; }
struct refract_ret_t { v3 t; f_t cost, eta_12; bool TIR; };
inline refract_ret_t refract(f_t eta_12, v3 w, v3 n = { 0, 0, 1 }) {
    const f_t wn = dot(w, n);
    eta_12 = wn > 0 ? eta_12 : 1.f / eta_12;
    const f_t cost2 = 1 - sqr(eta_12) * (1 - sqr(wn));
    if (cost2 >= 0) {
        const f_t cost = std::sqrt(cost2);
        const v3 t = eta_12 * (wn * n - w) - cost * (wn >= 0 ? n : -n);
        return { normalize(t), cost, eta_12, false };
    }
    return { { 0, 0, 1 }, 0, eta_12, true };
}
struct fresnel_ret_t { v3 t; c_t eta_12; f_t Z; c_t rs, rp, ts, tp; f_t Ts, Tp; bool TIR() const { return Ts == 0 && Tp == 0; } };
inline fresnel_ret_t fresnel(c_t eta_12, v3 w, v3 n = { 0, 0, 1 }) {      // fresnel.hpp:74-117
    if (eta_12 == c_t{ 1, 0 }) return { -w, eta_12, 1, 0, 0, 1, 1, 1, 1 };
    const f_t abs_cosi = std::fabs(dot(w, n));
    const auto refr = refract(eta_12.real(), w, n);
    if (abs_cosi == 0 || refr.TIR) return { { 0, 0, 1 }, refr.eta_12, 1, 1, 1, 0, 0, 0, 0 };
    const f_t cost = refr.cost;
    const f_t eta = refr.eta_12;
    // eta_12 becomes real here (assigned from refr.eta_12): complex arithmetic with zero imaginary part
    const c_t e{ eta, 0 };
    const c_t rs = (e * abs_cosi - cost) / (e * abs_cosi + cost);
    const c_t rp = (abs_cosi - e * cost) / (abs_cosi + e * cost);
    const c_t ts = rs + c_t{ 1, 0 };
    const c_t tp = (rp + c_t{ 1, 0 }) * e;
    const f_t Z = lm::cabs(cost / (e * abs_cosi));
    return { refr.t, e, Z, rs, rp, ts, tp, std::min(1.f, Z * std::norm(ts)), std::min(1.f, Z * std::norm(tp)) };
}
struct fresnel_conductor_ret_t { c_t rs, rp; };
inline fresnel_conductor_ret_t fresnel_reflection(c_t eta_12, v3 w, v3 n = { 0, 0, 1 }) {  // fresnel.hpp:128-144
    const f_t wn = dot(w, n);
    if (eta_12 == c_t{ 1, 0 } || wn < 0) return { 0, 0 };
    const c_t t2 = c_t{ 1, 0 } - (1 - sqr(wn)) * (eta_12 * eta_12);
    const c_t t = lm::csqrt(t2);
    const c_t i{ wn, 0 };
    return { (eta_12 * i - t) / (eta_12 * i + t), (i - eta_12 * t) / (i + eta_12 * t) };
}
// mueller.hpp:318-344
inline mueller_t mueller_fresnel_reflection(c_t eta_12, v3 w, v3 n = { 0, 0, 1 }) { const auto f = fresnel_reflection(eta_12, w, n); return mueller_t::fresnel(f.rs, f.rp); }
inline mueller_t mueller_fresnel_transmission(c_t eta_12, v3 w, v3 n = { 0, 0, 1 }) { const auto f = fresnel(eta_12, w, n); return f.Z * mueller_t::fresnel(f.ts, f.tp); }
inline mueller_t mueller_fresnel(c_t eta_12, bool reflection, v3 w, v3 n = { 0, 0, 1 }) {
    return reflection ? mueller_fresnel_reflection(eta_12, w, n) : mueller_fresnel_transmission(eta_12, w, n);
}

} // namespace ot
