/*
 * wtgpu.h -- C-ABI drop-in boundary of the B200-native wave_tracer hot path.
 *
 * This header is the only thing a host application needs.  It replaces the reference's
 * per-pixel virtual call
 *     integrator_t::integrate(ctx, block, sensor_element, samples_per_element)
 *         (/root/reference/include/wt/integrator/integrator.hpp:54-57),
 * as driven by scene_renderer_t::render's job loop
 *         (/root/reference/src/scene/render.cpp:99-113, 381-579),
 * with one call that renders a (tile, sample-range) of one sensor on one GPU:  wtgpu_render().
 *
 * The scene is handed over as plain-old-data tables (wtgpu_scene_desc).  Each table cites the reference
 * structure it flattens.  All lengths are metres (f32), wavenumbers are 1/mm (f32), as in the reference
 * (include/wt/math/quantity/defs.hpp:129,197).  No C++/torch types cross this boundary.
 *
 * Ownership: the caller owns every pointer inside wtgpu_scene_desc (copied at wtgpu_scene_create).
 * The library owns device memory.  The film buffers passed to wtgpu_render are caller-owned and may be
 * host or device pointers (wtgpu_render_opts::film_on_device).
 *
 * Error convention: every entry point returns 0 on success, a negative WTGPU_E_* code otherwise and
 * never throws; wtgpu_last_error() returns a thread-local message.  The reference hot path is noexcept
 * and drops invalid samples silently (film.hpp:232-240,271-277): so do we, device side.
 */
#ifndef WTGPU_H
#define WTGPU_H

#include <stdint.h>
#include <stddef.h>

#ifdef __cplusplus
extern "C" {
#endif

#define WTGPU_API_VERSION 1

/* ---- error codes ---- */
#define WTGPU_OK               0
#define WTGPU_E_INVALID       -1   /* bad argument / inconsistent description            */
#define WTGPU_E_CUDA          -2   /* CUDA runtime error (message in wtgpu_last_error)   */
#define WTGPU_E_NO_DEVICE     -3   /* no CUDA device / extension built without a GPU     */
#define WTGPU_E_UNSUPPORTED   -4   /* feature in the description not implemented         */
#define WTGPU_E_CAPACITY      -5   /* a per-path list could not be grown to the length the scene needs (out of device memory), or a BVH
                                      traversal stack of the reference's depth (64 / 128) was full: reported, never silent */

#define WTGPU_INVALID_IDX 0xffffffffu

/* ---------------------------------------------------------------------------------------------
 * Accelerating data structure: 8-wide BVH.
 * Flattens bvh8w::node_t (include/wt/ads/bvh8w/bvh8w_node.hpp:27-41): 8 child AABBs in SoA form,
 * 8 child pointers (>0: inner node idx+1, <0: -(leaf idx+1), 0: empty), and the contiguous triangle
 * range covered by the node.  Padded to 256 B so that one node is two 128-B lines.
 * ------------------------------------------------------------------------------------------- */
typedef struct wtgpu_node {
    float minx[8], miny[8], minz[8];
    float maxx[8], maxy[8], maxz[8];
    int32_t child[8];
    uint32_t tris_start, tris_count;
    uint32_t pad_[6];
} wtgpu_node;                       /* 256 B */

/* bvh8w::leaf_node_t (bvh8w_node.hpp): contiguous triangle range of a leaf. */
typedef struct wtgpu_leaf { uint32_t tris_ptr, count; } wtgpu_leaf;

/* ads::tri_t geometry (include/wt/ads/common.hpp:37-47), packed as three 16-B vectors so a triangle is
 * three LDG.128: (a.xyz, n.x) (b.xyz, n.y) (c.xyz, n.z). */
typedef struct wtgpu_tri {
    float ax, ay, az, nx;
    float bx, by, bz, ny;
    float cx, cy, cz, nz;
} wtgpu_tri;                        /* 48 B */

/* ads::tri_t bookkeeping: owning shape, triangle index in the shape's mesh, edge ids (WTGPU_INVALID_IDX: none). */
typedef struct wtgpu_tri_meta {
    uint32_t shape_idx, shape_tri_idx;
    uint32_t edge_ab, edge_bc, edge_ca;
    uint32_t pad_[3];
} wtgpu_tri_meta;                   /* 32 B */

/* mesh::triangle_t shading data (include/wt/mesh/triangle.hpp): per-vertex shading normals (already
 * decoded from the octahedral encoding, include/wt/math/encoded_normal.hpp), UVs, and dpdu of the
 * tangent frame (include/wt/mesh/surface_differentials.hpp).  Indexed by tuid. */
typedef struct wtgpu_tri_shading {
    float n0[3], n1[3], n2[3];
    float uv0[2], uv1[2], uv2[2];
    float dpdu[3];
    uint32_t has_uv;
} wtgpu_tri_shading;                /* 80 B */

/* ads::edge_t (include/wt/ads/common.hpp:53-72). */
typedef struct wtgpu_edge {
    float a[3], b[3], e[3];
    float n1[3], t1[3];
    float n2[3], t2[3];
    float alpha;                    /* wedge opening angle (rad) */
    uint32_t tri1, tri2;            /* tuids; tri2 == WTGPU_INVALID_IDX for an open edge */
} wtgpu_edge;                       /* 96 B */

/* shape_t (include/wt/scene/shape.hpp): bsdf, optional area emitter, uniform position sampling data. */
typedef struct wtgpu_shape {
    int32_t bsdf;                   /* index into bsdfs */
    int32_t emitter;                /* index into emitters or -1 */
    float surface_area;             /* m^2 */
    uint32_t tri_first, n_tris;     /* range in shape_tri_tuid / shape_tri_cdf (cdf has n_tris+1 entries per shape, at tri_first+shape index) */
    uint32_t cdf_first;
    uint32_t pad_[2];
} wtgpu_shape;

/* ---------------------------------------------------------------------------------------------
 * Spectra.  spectrum_t::value(k) (complex) / spectrum_real_t::f(k) (include/wt/spectrum/spectrum.hpp:37-93)
 * baked by the host either to a constant or to a uniformly tabulated, linearly interpolated table
 * over the sensor's wavenumber range.  Constant textures are baked the same way.
 * ------------------------------------------------------------------------------------------- */
#define WTGPU_SPECTRUM_CONSTANT 0u
#define WTGPU_SPECTRUM_TABLE    1u
typedef struct wtgpu_spectrum {
    uint32_t type;
    float re, im;                   /* constant value */
    float k0, inv_dk;               /* table: sample i is at k0 + i/inv_dk (1/mm) */
    uint32_t n, offset;             /* table: n entries (re,im pairs) starting at spectrum_data[2*offset] */
    uint32_t pad_;
} wtgpu_spectrum;

/* ---------------------------------------------------------------------------------------------
 * BSDFs.  bsdf_t implementations (include/wt/bsdf/<all>.hpp, src/bsdf/<all>.cpp) flattened to a node array;
 * wrappers reference their nested node by index.
 * ------------------------------------------------------------------------------------------- */
#define WTGPU_BSDF_DIFFUSE     0u   /* src/bsdf/diffuse.cpp      spec[0]=reflectance                       */
#define WTGPU_BSDF_DIELECTRIC  1u   /* src/bsdf/dielectric.cpp   spec[0]=extIOR spec[1]=IOR spec[2]=refl scale spec[3]=trans scale */
#define WTGPU_BSDF_SURFACE_SPM 2u   /* src/bsdf/surface_spm.cpp  same spectra + surface profile            */
#define WTGPU_BSDF_TWO_SIDED   3u   /* src/bsdf/two_sided.cpp    child                                     */
#define WTGPU_BSDF_COMPOSITE   4u   /* src/bsdf/composite.cpp    bins: wavenumber range -> child           */
#define WTGPU_BSDF_SCALE       5u   /* src/bsdf/scale.cpp        child, spec[0]=scale                      */
#define WTGPU_BSDF_MASK        6u   /* src/bsdf/mask.cpp         child, spec[0]=mask (constant only)       */

#define WTGPU_PROFILE_DIRAC             0u  /* interaction/surface_profile/dirac.hpp    */
#define WTGPU_PROFILE_GAUSSIAN          1u  /* interaction/surface_profile/gaussian.hpp:80-255, roughness-parametrised: prof_spec[0]=roughness */
#define WTGPU_PROFILE_FRACTAL_ROUGHNESS 2u  /* interaction/surface_profile/fractal.hpp  : prof_spec[0]=roughness */
#define WTGPU_PROFILE_FRACTAL_T         3u  /* fractal.hpp: prof_spec[0]=T (mm^2), prof_spec[1]=sigma_h (1/mm) */
#define WTGPU_PROFILE_GAUSSIAN_SIGMA    4u  /* gaussian.hpp, sigma-parametrised: prof_spec[0]=sigma (1/mm, surface_profile.hpp:41) */
typedef struct wtgpu_bsdf {
    uint32_t type;
    int32_t child;
    int32_t spec[4];                /* spectrum ids, -1: absent (=1) */
    uint32_t profile_type;
    int32_t prof_spec[2];
    float gamma;                    /* fractal profile exponent */
    uint32_t n_bins, bin_first;     /* composite: entries in bsdf_bins */
    uint32_t pad_[3];
} wtgpu_bsdf;
typedef struct wtgpu_bsdf_bin { float kmin, kmax; int32_t child; uint32_t pad_; } wtgpu_bsdf_bin;

/* ---------------------------------------------------------------------------------------------
 * Emitters (include/wt/emitter/{point,spot,directional,area}.hpp, src/emitter/<all>.cpp).
 * ------------------------------------------------------------------------------------------- */
#define WTGPU_EMITTER_POINT       0u
#define WTGPU_EMITTER_SPOT        1u
#define WTGPU_EMITTER_DIRECTIONAL 2u
#define WTGPU_EMITTER_AREA        3u
typedef struct wtgpu_emitter {
    uint32_t type;
    int32_t spectrum;               /* radiant intensity / irradiance / average radiance */
    float scale;                    /* area emitter scale */
    float pse_scale;                /* emitter_phase_space_extent_scale */
    float pos[3];                   /* point/spot position */
    float rot[9], inv_rot[9];       /* to_world linear part (row major) and its inverse */
    float cutoff, falloff;          /* spot angles (rad) */
    float extent;                   /* optional spatial extent (m); <=0: default 10 lambda */
    int32_t shape;                  /* area emitter: owning shape */
    float dir[3];                   /* directional: direction of propagation */
    float tan_alpha;                /* directional: sourced beams' tan(alpha) */
    float world_centre[3];          /* directional: scene bounding sphere */
    float world_radius;             /* directional: target_radius (directional.hpp:56-82) */
    float far_dist;                 /* directional: distance from world centre to the sourcing plane */
    uint32_t pad_[2];
} wtgpu_emitter;

/* scene_sensor_t::emitter_sampling_data_t (include/wt/scene/scene_sensor.hpp:31-148): per-emitter distribution
 * of wavenumbers (product of emission and sensitivity), discrete lines or binned piecewise-linear. */
#define WTGPU_KDIST_DISCRETE 0u
#define WTGPU_KDIST_BINNED   1u
typedef struct wtgpu_kdist {
    uint32_t type;
    uint32_t n, first;              /* kdist_data[first..]: discrete: k[n], y[n], dcdf[n+1] (discrete_distribution.hpp:171-200);
                                       binned: ys[n] knots, dcdf[n] (binned_piecewise_linear_distribution.hpp:37-60)           */
    float k0, dk;                   /* binned: knot i at k0+i*dk (1/mm) */
    float norm;                     /* 1/sum: pdf = y*norm */
    uint32_t pad_[2];
} wtgpu_kdist;

/* ---------------------------------------------------------------------------------------------
 * Sensor + film (include/wt/sensor/sensor/{perspective,virtual_plane_sensor}.hpp, sensor/film/film.hpp).
 * ------------------------------------------------------------------------------------------- */
#define WTGPU_SENSOR_PERSPECTIVE   0u
#define WTGPU_SENSOR_VIRTUAL_PLANE 1u
typedef struct wtgpu_sensor {
    uint32_t type;
    uint32_t width, height, channels;
    float rfilter_stddev;           /* film_t rfilter sigma in elements */
    uint32_t rf_radius;             /* ceil(3 sigma) (film.hpp:130) */
    uint32_t ray_trace_only;        /* sensor_t::ray_trace_only() */
    int32_t response[4];            /* per-channel response spectrum f(channel,k) */
    /* perspective */
    float pos[3];
    float rot[9], inv_rot[9];       /* sensor_transform linear part, row major */
    float s2c[16], c2s[16];         /* sensor_to_camera_trns matrix and inverse, row major */
    float sourcing_tan_alpha;
    float pse_scale;                /* phase_space_extent_scale */
    /* virtual plane */
    float frame_t[3], frame_b[3], frame_n[3];
    float origin[3];
    float extent[2];
    float requested_tan_alpha;      /* <0: none */
    uint32_t pad_[2];
} wtgpu_sensor;

/* ---------------------------------------------------------------------------------------------
 * Integrator options (src/integrator/plt_path.cpp:64-94, plt_bdpt.cpp:161-197).
 * ------------------------------------------------------------------------------------------- */
#define WTGPU_INTEGRATOR_PLT_PATH 0u
#define WTGPU_INTEGRATOR_PLT_BDPT 1u
#define WTGPU_DIRECTION_BACKWARD  0u
#define WTGPU_DIRECTION_FORWARD   1u
typedef struct wtgpu_integrator {
    uint32_t type;
    uint32_t direction;             /* plt_path only */
    uint32_t max_depth;
    uint32_t russian_roulette;
    uint32_t fsd;
    uint32_t mis, sensor_direct, emitter_direct;   /* plt_bdpt only */
} wtgpu_integrator;

/* ---------------------------------------------------------------------------------------------
 * sobolld sampler table: one line "d sj aj mk[0..sj)" of data/sobolld/initIrreducibleGF3.dat
 * (include/wt/sampler/sobolld/irreducible_gf3.hpp:124-158): dimension, degree of the irreducible polynomial over GF(3),
 * the polynomial's coefficients as a base-3 number (leading coefficient included), initial direction numbers.
 * ------------------------------------------------------------------------------------------- */
#define WTGPU_SOBOL_ENTRIES 48u     /* irreducible_gf3.hpp:34 */
#define WTGPU_SOBOL_DIMS    47u     /* src/sampler/sobolld.cpp:31: sobolls_sampler<47> */
#define WTGPU_SOBOL_DIGITS  11u     /* irreducible_gf3.hpp:33: batches of 3^11 = 177147 points */
typedef struct wtgpu_sobol_entry { int32_t d, sj, aj; int32_t mk[32]; int32_t pad_; } wtgpu_sobol_entry;   /* 144 B */

typedef struct wtgpu_scene_desc {
    uint32_t api_version;           /* WTGPU_API_VERSION */

    /* ADS */
    uint32_t n_nodes;    const wtgpu_node* nodes;
    uint32_t n_leaves;   const wtgpu_leaf* leaves;
    int32_t  root_ptr;              /* encoded like a child pointer */
    uint32_t n_tris;     const wtgpu_tri* tris;
                         const wtgpu_tri_meta* tri_meta;
                         const wtgpu_tri_shading* tri_shading;
    uint32_t n_edges;    const wtgpu_edge* edges;
    float world_min[3], world_max[3];

    /* shapes */
    uint32_t n_shapes;   const wtgpu_shape* shapes;
    uint32_t n_shape_tris; const uint32_t* shape_tri_tuid;   /* (shape, shape_tri_idx) -> tuid */
    uint32_t n_shape_cdf;  const float* shape_tri_cdf;       /* per shape: n_tris+1 normalised area cdf values */

    /* materials */
    uint32_t n_spectra;  const wtgpu_spectrum* spectra;
    uint32_t n_spectrum_data; const float* spectrum_data;    /* (re,im) pairs */
    uint32_t n_bsdfs;    const wtgpu_bsdf* bsdfs;
    uint32_t n_bsdf_bins; const wtgpu_bsdf_bin* bsdf_bins;

    /* emitters + spectral sampling */
    uint32_t n_emitters; const wtgpu_emitter* emitters;
                         const float* emitter_cdf;           /* n_emitters+1, normalised */
                         const wtgpu_kdist* emitter_kdist;   /* n_emitters */
    uint32_t n_kdist_data; const float* kdist_data;

    wtgpu_sensor sensor;
    wtgpu_integrator integrator;

    /* Fraunhofer FSD importance-sampling tables (plt_bdpt): fsd_lut_t (include/wt/interaction/fsd/fraunhofer/fsd_lut.hpp:27-69).
     * The reference ships them as data/fsd/iCDFa{1,2}{,theta}.fp64 (Git-LFS stubs here): regenerated by the host
     * (wave_tracer_b200/fsd_lut.py).  icdf_theta*: n entries over u in [0,1] -> theta in [0,pi/2];
     * icdf*: m x m, row = theta bin, column = u -> radius. */
    uint32_t fsd_lut_n, fsd_lut_m;
    const float *fsd_icdf_theta1, *fsd_icdf_theta2, *fsd_icdf1, *fsd_icdf2;

    /* sobolld scene sampler (src/sampler/sobolld.cpp:29-60): WTGPU_SOBOL_ENTRIES parsed lines of data/sobolld/initIrreducibleGF3.dat
     * (irreducible_gf3.hpp:124-158; entry 0 is skipped by the reference, dimensions use entries 1..47), or NULL when the scene's
     * sampler is the uniform one.  Required when wtgpu_render_opts::sampler == WTGPU_SAMPLER_SOBOLLD. */
    const struct wtgpu_sobol_entry* sobol_table;
} wtgpu_scene_desc;

/* ---------------------------------------------------------------------------------------------
 * Render call.
 * Film layout (caller-owned, zero-initialised by the caller or accumulated into):
 *   film_block : float[height][width][channels][2]  (value, weight) -- block splats, film_t::splat (film.hpp:254-288)
 *   film_light : float[height][width][channels]     -- direct splats, film_t::splat_direct (film.hpp:214-252)
 * Developed linear image = value/weight + light/spp (film_storage.hpp:256-291,354-358): wtgpu_develop().
 *
 * RNG contract (the reference has no user seed -- seeded_mt19937_64.hpp:31-50 -- so this is ours):
 * a counter-based Philox4x32-10 stream keyed by `seed`, indexed by (pixel linear index, sample index,
 * draw counter); results are invariant to tiling, sample ranges and the number of GPUs.
 * Draw d of (pixel, sample, stream) is lane d&3 of Philox4x32-10(key = seed, counter = (d>>2, sample, pixel, stream)),
 * float = (u32 >> 8) * 2^-24.  stream is 0 for plt_path.  plt_bdpt splits a sample into sub-streams (each starting at d = 0) so
 * that the two subpath walks and every (s,t) strategy are independent: 0 = emitter / wavenumber / source sampling, 1 = sensor
 * subpath walk, 2 = emitter subpath walk, 3 + 4096 t + s = strategy (s,t).
 *
 * sobolld contract (the reference draws a fresh batch of 3^11 x 47 values per pool thread with random seeds,
 * src/sampler/sobolld.cpp:53-60, and consumes it as one flat stream, sobolld.hpp:40-48): sample `s` of element `p` owns point
 * g = p * spp + s of the global sequence; batch b = g / 3^11 is Owen-scrambled with seeds[dim] = lane 0 of
 * Philox4x32-10(key = seed, counter = (dim, b_lo, b_hi, 0x50B01D)); scene-sampler draw d of that sample is dimension d % 47 of
 * point g + d / 47 (the reference's flat layout, each sample starting at a point boundary).  Digit arithmetic is bit-exact with
 * sobolld_sampler.hpp:59-207 (wtgpu_debug_sobol() exposes the integer numerators).
 * ------------------------------------------------------------------------------------------- */
typedef struct wtgpu_render_opts {
    uint64_t seed;
    uint32_t spp;                   /* the sensor's samples per element (normalises the light image) */
    uint32_t sample_begin, sample_end;  /* render samples [begin,end) of each element */
    uint32_t tile_x0, tile_y0, tile_x1, tile_y1;  /* element rectangle [x0,x1) x [y0,y1) */
    int32_t device;                 /* CUDA device ordinal */
    uint32_t film_on_device;        /* film pointers are device pointers */
    uint32_t pool_size;             /* paths in flight (0: default) */
    uint32_t sampler;               /* WTGPU_SAMPLER_*: the scene sampler (scene_t::sampler(), src/scene/loader/loader.cpp:299-300) */
    uint32_t flags;                 /* WTGPU_RENDER_* */
    void* stream;                   /* cudaStream_t or NULL */
} wtgpu_render_opts;
#define WTGPU_SAMPLER_UNIFORM 0u    /* sampler::uniform_t -> the Philox stream */
#define WTGPU_SAMPLER_SOBOLLD 1u    /* sampler::sobolld_t for the scene-sampler draws (emitter / wavenumber / source beam / sensor element:
                                       plt_path_detail.hpp:772,783; plt_bdpt.cpp:56,70); path sampling stays on the Philox stream, as the
                                       reference's integrators keep their own uniform_t for it (plt_path_detail.hpp:59-60,146; plt_bdpt.cpp:50-52) */
#define WTGPU_RENDER_NO_SORT 1u     /* disable the material sort (for A/B measurement) */
#define WTGPU_RENDER_BDPT_MEGAKERNEL 4u  /* plt_bdpt: run the one-thread-per-sample cross-check kernel instead of the wavefront */
#define WTGPU_RENDER_THREAD_TRAVERSE 8u  /* one thread per beam in traverse() instead of eight lanes per beam (A/B measurement; bit-identical results) */
#define WTGPU_RENDER_GROUP_TRAVERSE 16u  /* force eight lanes per beam (default: chosen by scene size for plt_path, always for plt_bdpt) */
#define WTGPU_RENDER_NO_RAY_CULL 32u /* ray queries walk every node along the infinite ray, as bvh8w.cpp:469-554 does, instead of culling children outside the query range (A/B; same results) */
#define WTGPU_RENDER_TIME_KERNELS 2u /* record CUDA events around every kernel (fills wtgpu_stats::*_ms); runs ONE sub-pool so that kernels do not overlap */
#define WTGPU_RENDER_ONE_SUBPOOL 64u /* one wavefront at a time instead of four side by side (A/B measurement; same results) */

/* Device counters gathered during wtgpu_render (the quantities the reference exposes in a `profile` build:
 * include/wt/ads/ads_stats.hpp:36-95, include/wt/integrator/stats.hpp:27-82); inputs of the roofline byte count. */
typedef struct wtgpu_stats {
    uint64_t samples;
    uint64_t segments;              /* traverse() calls */
    uint64_t ray_casts, cone_casts, shadow_casts;
    uint64_t nodes_visited;         /* 256-B node fetches */
    uint64_t tris_tested;           /* triangle fetches */
    uint64_t edges_fetched;
    uint64_t surface_interactions, fsd_interactions, null_interactions;
    uint64_t splats;                /* film taps written */
    uint64_t capacity_overflows;    /* lists that did not fit their row in the FINAL pass: 0 on success (see `passes`) */
    uint64_t kernel_launches;
    uint64_t iterations;
    uint64_t traverse_nodes, traverse_tris;   /* node / triangle fetches of k_traverse alone (roofline of the dominant kernel) */
    uint64_t shaded_paths;          /* path-vertices processed by k_shade */
    double   gpu_ms;                /* CUDA-event time of the whole call on its stream */
    double   traverse_ms, shade_ms, generate_ms, sort_ms;
    double   connect_ms;            /* plt_bdpt: the (s,t) strategy kernels (shade_ms then covers the vertex step only) */
    uint64_t strategies[5];         /* plt_bdpt: strategies evaluated per class: s=0, t=0, s=1, t=1, vertex-vertex */
    uint64_t walker_steps;          /* plt_bdpt: subpath-walker traverse() calls */
    /* Capacity growth.  Cone-query triangle lists, edge sets, aperture segments, apertures and subpath vertices are std::vector / std::set of any
     * length in the reference (traversal_common.hpp:116-149); here they are rows of device arrays whose lengths belong to the scene handle.  A pass
     * over the samples that finds a longer list is discarded, the rows are re-sized to what it measured, and the pass is repeated (`passes` > 1);
     * the handle keeps the lengths for the next render.  Results therefore never depend on a capacity. */
    uint32_t passes;                /* passes over the samples this call took (1: every list fitted) */
    uint32_t pool_used;             /* paths / sample slots in flight (smaller than asked when long rows would not fit in device memory) */
    uint32_t cap_tris, cap_edges, cap_segments, cap_apertures, cap_vertices;   /* capacities after this call: entries of the shared cone-triangle-list
                                       arena (lists longer than 128 triangles continue there; it is recycled every iteration), then the row lengths */
    uint32_t subpools;              /* independent wavefronts the call ran side by side (1 under WTGPU_RENDER_TIME_KERNELS) */
    uint64_t stack_drops;           /* children a full BVH traversal stack dropped (bvh8w.cpp's stack depths); non-zero => WTGPU_E_CAPACITY */
} wtgpu_stats;

typedef struct wtgpu_scene wtgpu_scene;

int wtgpu_device_count(void);
const char* wtgpu_last_error(void);
int wtgpu_scene_create(const wtgpu_scene_desc* desc, int device, wtgpu_scene** out);
void wtgpu_scene_destroy(wtgpu_scene* scene);
/* Device blocks >= 1 MiB released by a destroyed scene / finished render are kept for reuse by the next one (path pools are GBs);
 * wtgpu_trim() returns them to the driver. */
void wtgpu_trim(void);
int wtgpu_render(wtgpu_scene* scene, const wtgpu_render_opts* opts,
                 float* film_block, float* film_light, wtgpu_stats* stats);
/* capacities of the per-path lists of a scene handle, {cone-triangle-list arena entries, edges, aperture segments, apertures per subpath, vertices per subpath}:
 * read them / preset them (e.g. from a previous run of the same scene, to skip the measuring pass; or tiny, to test the growth) */
int wtgpu_get_capacities(wtgpu_scene* scene, uint32_t out[5]);
int wtgpu_set_capacities(wtgpu_scene* scene, const uint32_t in[5]);
/* out[h][w][c] = block value/weight + light/spp ; host pointers */
int wtgpu_develop(const wtgpu_sensor* sensor, uint32_t spp,
                  const float* film_block, const float* film_light, float* out);
/* the same on the device (film_storage.hpp:256-291, 354-358: film develop of the block image + the light image): DEVICE pointers, queued on
 * `stream` (a cudaStream_t, or null); one pass over the films, HBM-bound (12 B read + 4 B written per element).  What a multi-GPU render
 * calls on rank 0 after the film reduce, so that only the developed image crosses PCIe. */
int wtgpu_develop_device(const wtgpu_sensor* sensor, uint32_t spp,
                         const float* d_film_block, const float* d_film_light, float* d_out, void* stream, int device);

/* ---- unit-level entry points used by the parity tests (same kernels the renderer uses) ---- */
typedef struct wtgpu_ray_query { float o[3], d[3]; float tmin, tmax; } wtgpu_ray_query;
typedef struct wtgpu_ray_hit { uint32_t tuid; float dist; float bary[2]; uint32_t front_face; } wtgpu_ray_hit;
int wtgpu_debug_intersect_rays(wtgpu_scene* scene, uint32_t n, const wtgpu_ray_query* q, wtgpu_ray_hit* out);
int wtgpu_debug_shadow_rays(wtgpu_scene* scene, uint32_t n, const wtgpu_ray_query* q, uint32_t* out);

/* elliptic_cone_t (include/wt/math/shapes/elliptic_cone.hpp:30-56) */
typedef struct wtgpu_cone_query {
    float o[3], d[3], x[3];
    float x0, tan_alpha, e;         /* initial major axis, tan half-angle, major/minor ratio (>=1) */
    float tmin, tmax;
    float z_scale;                  /* intersect_opts_t::z_search_range_scale */
} wtgpu_cone_query;
#define WTGPU_MAX_CONE_TRIS  128
#define WTGPU_MAX_CONE_EDGES 48
typedef struct wtgpu_cone_hit {
    float dist; uint32_t front_face;
    uint32_t n_tris, n_edges;
    uint32_t tris[WTGPU_MAX_CONE_TRIS];
    uint32_t edges[WTGPU_MAX_CONE_EDGES];
} wtgpu_cone_hit;
int wtgpu_debug_intersect_cones(wtgpu_scene* scene, uint32_t n, const wtgpu_cone_query* q, wtgpu_cone_hit* out);

/* counter-based RNG stream: out[i] = i-th draw of stream (seed, pixel, sample) */
int wtgpu_debug_rng(uint64_t seed, uint32_t pixel, uint32_t sample, uint32_t n, float* out, int device);
/* the portable elementary functions of wave_tracer_b200/csrc/pmath.h evaluated on the device (they replace the host libm the reference calls through
 * m::sin ..., include/wt/math/common.hpp, and return the same bits on host and device): fn 0 sin, 1 cos, 2 tan, 3 exp, 4 log, 5 pow(x,y), 6 atan2(x,y),
 * 7 acos, 8 hypot(x,y), 9 / 10 real / imaginary part of the UTD transition function UTDF(x) (interaction/fsd/utd.hpp:36-57); y may be NULL for unary fn */
int wtgpu_debug_pmath(int fn, uint32_t n, const float* x, const float* y, float* out, int device);
/* sobolld: numerators (value * 3^11) and floats of dimensions 0..46 of points g0 .. g0+n-1 of the global sequence; out arrays n*47 */
int wtgpu_debug_sobol(wtgpu_scene* scene, uint64_t seed, uint64_t g0, uint32_t n, uint32_t* out_numerators, float* out_values);
/* sizeof of the i-th ABI struct (order of wave_tracer_b200/_abi.py:ABI_STRUCTS); lets bindings verify their layout */
uint64_t wtgpu_debug_sizeof(int which);

#ifdef __cplusplus
}
#endif
#endif /* WTGPU_H */
