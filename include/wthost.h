/*
 * wthost.h -- host-side (CPU, C++) scene preparation for the B200 wave_tracer hot path, C-ABI.
 *
 * Replaces, for the standalone mode, the reference's ADS constructor plug-in
 *   scene_bootstrap_t<SceneLoader, ADSCtor>  (/root/reference/include/wt/scene/loader/bootstrap.hpp:84-88)
 *   bvh8w_constructor_t                      (/root/reference/src/ads/bvh8w_constructor.cpp:153-268)
 *   bvh_constructor_t                        (/root/reference/src/ads/bvh_constructor.cpp:123-251)
 *   find_edges                               (/root/reference/include/wt/ads/edge_classification.hpp:31-238)
 * and mesh_t's triangle preparation          (/root/reference/src/mesh/mesh.cpp:31-150).
 * When linked into the real application the same tables can instead be filled from the reference's
 * bvh8w_t accessors (bvh8w.hpp:72-84) -- see INTEGRATION.md.
 */
#ifndef WTHOST_H
#define WTHOST_H

#include "wtgpu.h"

#ifdef __cplusplus
extern "C" {
#endif

/* One shape's indexed mesh, as handed to mesh_t's constructor (src/mesh/mesh.cpp:103-118). */
typedef struct wthost_mesh_desc {
    uint32_t n_verts;
    const float* positions;         /* 3*n_verts, metres */
    const float* normals;           /* 3*n_verts or NULL */
    const float* uvs;               /* 2*n_verts or NULL */
    uint32_t n_tris;
    const uint32_t* indices;        /* 3*n_tris */
    double to_world[16];            /* row-major 4x4 applied in double precision */
    int32_t bsdf, emitter;          /* recorded into wtgpu_shape */
} wthost_mesh_desc;

typedef struct wthost_ads wthost_ads;

/* Builds triangles, binary SAH BVH (C_INT=100, C_TRAV=1, 128 bins -- bvh_constructor.cpp:17-31),
 * collapses 3 binary levels into 8-wide nodes, and classifies edges (deterministic edge ids). */
int wthost_ads_build(uint32_t n_meshes, const wthost_mesh_desc* meshes, wthost_ads** out);
/* Points the ADS and shape fields of `desc` at tables owned by `ads` (valid until wthost_ads_destroy). */
int wthost_ads_fill(const wthost_ads* ads, wtgpu_scene_desc* desc);
void wthost_ads_destroy(wthost_ads* ads);
/* diagnostics */
double wthost_ads_sah_cost(const wthost_ads* ads);
uint32_t wthost_ads_max_depth(const wthost_ads* ads);


/* sobolld: generator matrices of the 47 dimensions derived from the parsed table (the reference's generate_mkgf3 + gen_mat,
 * include/wt/sampler/sobolld/irreducible_gf3.hpp:103-118, sobolld_sampler.hpp:140-154) in the packed form the device keeps in
 * constant memory: bit c of ones[dim*11 + j] is set when matrix row 10-j, column c equals 1 (twos: equals 2).  Host only. */
int wthost_sobol_tables(const wtgpu_sobol_entry* table, uint16_t* ones /* 47*11 */, uint16_t* twos /* 47*11 */);

#ifdef __cplusplus
}
#endif
#endif
