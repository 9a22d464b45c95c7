/* ORACLE -- TEST INFRASTRUCTURE ONLY.  Never linked, imported or executed by the product path. */
#ifndef ORACLE_H
#define ORACLE_H
#include "../include/wtgpu.h"
#ifdef __cplusplus
extern "C" {
#endif
typedef struct oracle_stats {
    uint64_t samples, segments, surface, fsd, null_, splats;
    uint64_t nodes, tris, ray_casts, cone_casts, shadow_casts;
    double seconds; uint32_t threads; uint32_t pad_;
} oracle_stats;
/* film_block: double[h][w][c][2], film_light: double[h][w][c]; accumulated into */
int oracle_render(const wtgpu_scene_desc* desc, const wtgpu_render_opts* opts, double* film_block, double* film_light, uint32_t n_threads, oracle_stats* st);
int oracle_intersect_rays(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_ray_query* q, wtgpu_ray_hit* out);
int oracle_shadow_rays(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_ray_query* q, uint32_t* out);
int oracle_intersect_rays_bruteforce(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_ray_query* q, wtgpu_ray_hit* out);
int oracle_intersect_cones(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_cone_query* q, wtgpu_cone_hit* out);
int oracle_cone_closest_bruteforce(const wtgpu_scene_desc* desc, uint32_t n, const wtgpu_cone_query* q, float* dist);
int oracle_sobol_batch(const wtgpu_sobol_entry* table, uint64_t seed, uint64_t batch, uint32_t n_points, uint32_t* out_numerators, float* out_values);
int oracle_sobol_points_with_seeds(const wtgpu_sobol_entry* table, const uint64_t* seeds, uint32_t n_points, uint32_t* out_numerators, float* out_values);
void oracle_sobol_seeds(uint64_t seed, uint64_t batch, uint64_t* out /* 47 */);
int oracle_sobol_matrices(const wtgpu_sobol_entry* table, int32_t* out);
int oracle_rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out);
void oracle_svd(const float A[4], float out[6]);
void oracle_svd_n(uint32_t n, const float* A, float* out);
void oracle_utdf(float x, float out[2]);
void oracle_utdf_n(uint32_t n, const float* x, float* out);
void oracle_utd(uint32_t n, const float* wedge, const float* q, float* out);
void oracle_utd_diffraction_points(uint32_t n, const float* wedge, const float* pts, int* found, float* out);
void oracle_cerfc_rot45(double s, double out[2]);
void oracle_fresnel(float eta_re, float eta_im, const float w[3], float out[12]);
void oracle_fresnel_full(float eta_re, float eta_im, const float w[3], float out[16]);
void oracle_fresnel_reflection(float eta_re, float eta_im, const float w[3], float out[4]);
void oracle_reflect(const float w[3], float out[3]);
float oracle_mub_sbp(float length, float k);
float oracle_bsdf_albedo(const wtgpu_scene_desc* desc, int32_t bsdf, const float wi[3], float k, uint32_t n, uint64_t seed);
void oracle_profile_eval(const wtgpu_scene_desc* desc, int32_t bsdf, const float wi[3], const float wo[3], float k, float out[3]);
void oracle_profile_check(const wtgpu_scene_desc* desc, int32_t bsdf, const float wi[3], float k, uint32_t n, uint64_t seed, float out[4]);
void oracle_fuzz_cone_quick_reject(uint32_t n, uint64_t seed, uint64_t out[4]);
void oracle_fuzz_ray_cull(uint32_t n, uint64_t seed, uint64_t out[4]);
void oracle_cone_through_ellipsoid(const float axes[3], const float frame[9], const float o[3], const float d[3], float tan_alpha, float out[8]);
float oracle_fraunhofer_asf(uint32_t n, const float* edges, float xix, float xiy);
void oracle_fsd_eval(uint32_t n, const float* edges, float P0v, float psi02, float xix, float xiy, float out[9]);
void oracle_fsd_lut_sample(uint32_t n, uint32_t m, const float* theta, const float* icdf, uint32_t cnt, const float* rand, float* out);
void oracle_fsd_sampler_sample(uint32_t n, uint32_t m, const float* th1, const float* th2, const float* c1, const float* c2, uint32_t n_edges, const float* edges,
                               const float* edge_pdfs, float P0v, float P0_pdf, float psi02, float recp_I, const float* script, uint32_t n_script, uint32_t n_samples, float* out);
void oracle_sampler_warps(float u1, float u2, float solid_angle, float out[13]);
float oracle_erf_lut(float x);
void oracle_pmath(int fn, uint32_t n, const float* x, const float* y, float* out);
float oracle_gaussian_integrate_triangle(float sx, float sy, const float tri[6]);
void oracle_binned_eval(uint32_t n, const float* ys, const float* dcdf, float k0, float dk, float norm, uint32_t m, const float* v, float* icdf, const float* x, float* value, float* pdf);
void oracle_gaussian1d_integrate(float sigma, uint32_t n, const float* mn, const float* mx, float* out);
void oracle_discrete_icdf(uint32_t n, const float* dcdf, uint32_t m, const float* v, int* idx);
void oracle_gaussian_pdf(float sx, float sy, uint32_t n, const float* pts, float* out);
void oracle_clip_triangles(uint32_t n, const float* tri, const float* zr, int* ntris, float* polygon, float* pieces);
void oracle_gaussian_integrate_triangles(float sx, float sy, uint32_t n, const float* tri, float* out);
void oracle_debug_cone_hist(int on, uint64_t out[96]);
void oracle_frame_orthogonal(uint32_t n, const float* nrm, float* out);
void oracle_frame_shading(uint32_t n, const float* nrm, const float* dpdu, float* out);
void oracle_frame_xform(uint32_t n, const float* fr, const float* v, float* out);
void oracle_rotation2(uint32_t n, const float* from, const float* to, float* out);
void oracle_edge_ellipsoid(uint32_t n, const float* in, float* out);
void oracle_edge_ellipse(uint32_t n, const float* in, float* out);
void oracle_edge_plane(uint32_t n, const float* in, float* out);
void oracle_cone_edge(uint32_t n, int in_local, const float* in, float* out);
void oracle_cone_plane(uint32_t n, int in_local, const float* in, float* out);
void oracle_cone_tri(uint32_t n, const float* in, float* out);
void oracle_ray_tri(uint32_t n, const float* in, float* out);
void oracle_ray_tri_w(uint32_t n, const float* in, float* out);
void oracle_ray_aabb_fast(uint32_t n, const float* in, float* out);
void oracle_cone_work_lists(const wtgpu_scene_desc* desc, uint32_t n, const float* q, uint32_t cap, uint32_t* counts, uint32_t* tuids, float* dist, uint32_t* front);
void oracle_cone_through_ellipse_n(uint32_t n, const float* in, float* out);
void oracle_cone_through_ellipsoid_n(uint32_t n, const float* in, float* out);
void oracle_mueller(uint32_t n, const float* in, float* out);
void oracle_integrator_traverse(const wtgpu_scene_desc* desc, uint32_t n, const float* q, uint32_t cap, float* out, uint32_t* ntris, uint32_t* tris, uint32_t* nedges, uint32_t* edges);
void oracle_find_closest_triangle(const wtgpu_scene_desc* desc, uint32_t n, const float* q, float* out, uint32_t* tuid);
void oracle_edge_offsets(const wtgpu_scene_desc* desc, uint32_t n, const float* q, float* out);
void oracle_bd_find_closest_triangle(const wtgpu_scene_desc* desc, uint32_t n, const float* q, float* out, uint32_t* tuid);
void oracle_ffsd_aperture(const wtgpu_scene_desc* desc, uint32_t n, const float* q, uint32_t cap, uint32_t* counts, float* summary, float* edges);
void oracle_utd_fsd(const wtgpu_scene_desc* desc, uint32_t n, const float* q, uint32_t cap, uint32_t* nap, float* ap, uint32_t* nf, float* fo);
void oracle_cone_cluster(uint32_t n, const float* in, float* out);
void oracle_stack_sorter(uint32_t n, uint32_t run, float* io);
void oracle_cone_basics(uint32_t n, const float* in, float* out);
void oracle_point_in_triangle3(uint32_t n, const float* in, float* out);
void oracle_point_in_triangle2(uint32_t n, const float* in, float* out);
#ifdef __cplusplus
}
#endif
#endif
