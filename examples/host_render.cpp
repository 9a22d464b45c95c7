// host_render.cpp -- a C++ host driving the B200 hot path through the C-ABI alone (include/wtgpu.h, include/wthost.h); no Python, no torch.
//
// This is the shape of the glue INTEGRATION.md describes for wave_tracer's own `scene_renderer_t::render` (src/scene/render.cpp:381-579):
// flatten the scene into plain-old-data tables, create the device scene once, render (tile, sample-range) slices into caller-owned films,
// develop.  The scene is built here by hand -- a slit in a diffuse screen between a spot emitter and a virtual-plane sensor, plt_path forward
// with UTD free-space diffraction -- field for field as wave_tracer_b200/scene.py does it (tests/test_host_example.py renders the same scene
// through the Python layer and compares the films).
//
//   g++ -std=c++17 -O2 -Iinclude examples/host_render.cpp -Lwave_tracer_b200 -lwt_b200 -Wl,-rpath,'$ORIGIN/../wave_tracer_b200' -o examples/host_render
//   examples/host_render [res] [spp] [out.f32]        prints one line: samples, film sum, device ms; optionally writes the developed image
#include <wtgpu.h>
#include <wthost.h>

#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

static const double kTwoPi = 6.283185307179586476925286766559;

struct Rect { std::vector<float> pos, uv; std::vector<uint32_t> idx; };
// rectangle shape (src/mesh/rectangle.cpp:26-79): corners p, p+x, p+x+y, p+y in f32, uv = cell corners, triangles (0,1,2), (2,3,0) -- exactly what
// wave_tracer_b200.scene.rectangle hands to the mesh (the uv set fixes the tangent frame, hence the mapping of random numbers to directions)
static Rect rectangle(const double pd[3], const double xd[3], const double yd[3]) {
    Rect r;
    float p[3], x[3], y[3];
    for (int i = 0; i < 3; ++i) { p[i] = (float)pd[i]; x[i] = (float)xd[i]; y[i] = (float)yd[i]; }
    const float uv[4][2] = { { 0, 0 }, { 1, 0 }, { 1, 1 }, { 0, 1 } };
    for (auto& c : uv) {
        for (int i = 0; i < 3; ++i) r.pos.push_back((p[i] + c[0] * x[i]) + c[1] * y[i]);
        r.uv.push_back(c[0]); r.uv.push_back(c[1]);
    }
    r.idx = { 0, 1, 2, 2, 3, 0 };
    return r;
}
static void identity16(double m[16]) { for (int i = 0; i < 16; ++i) m[i] = (i % 5 == 0) ? 1.0 : 0.0; }

#define CHECK(call) do { const int rc_ = (call); if (rc_ != WTGPU_OK) { std::fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, wtgpu_last_error()); return 2; } } while (0)

int main(int argc, char** argv) {
    const uint32_t res = argc > 1 ? (uint32_t)std::atoi(argv[1]) : 96u, spp = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 4u;
    const char* out_path = argc > 3 ? argv[3] : nullptr;
    const double MM = 1e-3, lam = 0.08 * MM;
    const float k = (float)(kTwoPi / (lam * 1e3));                 // wavenumber in 1/mm (include/wt/math/quantity/math.hpp:25-27)

    // ---- geometry: back wall, two half-screens leaving a 0.9 mm slit
    std::vector<Rect> rects;
    { const double p[3] = { -80 * MM, -15 * MM, 40 * MM }, x[3] = { 160 * MM, 0, 0 }, y[3] = { 0, 30 * MM, 0 }; rects.push_back(rectangle(p, x, y)); }
    { const double p[3] = { -6 * MM, -15 * MM, -12 * MM }, x[3] = { (6 - .45) * MM, 0, 0 }, y[3] = { 0, 30 * MM, 0 }; rects.push_back(rectangle(p, x, y)); }
    { const double p[3] = { .45 * MM, -15 * MM, -12 * MM }, x[3] = { (6 - .45) * MM, 0, 0 }, y[3] = { 0, 30 * MM, 0 }; rects.push_back(rectangle(p, x, y)); }
    std::vector<wthost_mesh_desc> meshes(rects.size());
    for (size_t i = 0; i < rects.size(); ++i) {
        wthost_mesh_desc& m = meshes[i]; std::memset(&m, 0, sizeof(m));
        m.n_verts = 4; m.positions = rects[i].pos.data(); m.uvs = rects[i].uv.data(); m.n_tris = 2; m.indices = rects[i].idx.data();
        identity16(m.to_world);
        m.bsdf = i == 0 ? 0 : 2;            // bsdf nodes below: 0 = twosided(wall), 2 = twosided(screen)
        m.emitter = -1;
    }
    wthost_ads* ads = nullptr;
    CHECK(wthost_ads_build((uint32_t)meshes.size(), meshes.data(), &ads));
    wtgpu_scene_desc d; std::memset(&d, 0, sizeof(d));
    d.api_version = WTGPU_API_VERSION;
    CHECK(wthost_ads_fill(ads, &d));

    // ---- spectra (constants at the sensor's single wavenumber) and bsdf nodes: twosided(diffuse .85), twosided(diffuse .3)
    auto constant = [](float re) { wtgpu_spectrum s; std::memset(&s, 0, sizeof(s)); s.type = WTGPU_SPECTRUM_CONSTANT; s.re = re; return s; };
    const std::vector<wtgpu_spectrum> spectra = { constant(1.f) /*0: sensor response*/, constant(900.f) /*1: radiant intensity*/, constant(.85f) /*2*/, constant(.3f) /*3*/ };
    auto bsdf = [](uint32_t type, int32_t child, int32_t spec0) {
        wtgpu_bsdf b; std::memset(&b, 0, sizeof(b)); b.type = type; b.child = child;
        for (int i = 0; i < 4; ++i) b.spec[i] = -1;
        b.spec[0] = spec0; b.prof_spec[0] = b.prof_spec[1] = -1; b.gamma = 3.f;
        return b;
    };
    const std::vector<wtgpu_bsdf> bsdfs = { bsdf(WTGPU_BSDF_TWO_SIDED, 1, -1), bsdf(WTGPU_BSDF_DIFFUSE, -1, 2), bsdf(WTGPU_BSDF_TWO_SIDED, 3, -1), bsdf(WTGPU_BSDF_DIFFUSE, -1, 3) };
    d.n_spectra = (uint32_t)spectra.size(); d.spectra = spectra.data();
    d.n_bsdfs = (uint32_t)bsdfs.size(); d.bsdfs = bsdfs.data();

    // ---- sensor: virtual plane 200 mm x 66.7 mm at z = 40 mm looking back at the slit, res x res/3 film, monochromatic response
    wtgpu_sensor& s = d.sensor;
    s.type = WTGPU_SENSOR_VIRTUAL_PLANE; s.width = res; s.height = res / 3; s.channels = 1;
    s.rfilter_stddev = .25f * .1f;                                  // beam_source_spatial_stddev * rfilter_scale (film.hpp:417)
    s.rf_radius = (uint32_t)(std::ceil(s.rfilter_stddev * 3.f) + .5f);
    s.ray_trace_only = 0; s.response[0] = 0; s.response[1] = s.response[2] = s.response[3] = -1;
    // lookat(origin (0,0,39.999 mm), target (0,0,2 mm), up (0,-1,0)): d = -z, l = up x d = +x, u = d x l = -y
    const float ft[3] = { 1, 0, 0 }, fb[3] = { 0, -1, 0 }, fn[3] = { 0, 0, -1 };
    const double ex = 200 * MM, ey = 200.0 / 3 * MM, cz = (40 - .001) * MM;
    for (int i = 0; i < 3; ++i) { s.frame_t[i] = ft[i]; s.frame_b[i] = fb[i]; s.frame_n[i] = fn[i]; }
    s.origin[0] = (float)(-ex / 2); s.origin[1] = (float)(ey / 2); s.origin[2] = (float)cz;          // centre - ex/2 t - ey/2 b
    s.extent[0] = (float)ex; s.extent[1] = (float)ey;
    s.requested_tan_alpha = (float)std::tan(.002 * kTwoPi / 360.0);

    // ---- emitter: spot at z = -400 mm pointing at +z (to_world columns l, u, d = (0,1,0), (0,0,... ) built from up = (1,0,0))
    wtgpu_emitter e; std::memset(&e, 0, sizeof(e));
    e.type = WTGPU_EMITTER_SPOT; e.spectrum = 1; e.scale = 1.f; e.pse_scale = 1.f; e.shape = -1; e.extent = -1.f;
    e.pos[0] = 0; e.pos[1] = 0; e.pos[2] = (float)(-400 * MM);
    // lookat(origin, target = 0, up = (1,0,0)): d = (0,0,1), l = up x d = (0,-1,0), u = d x l = (1,0,0); rotation columns (l, u, d), row major
    const float R[9] = { 0, 1, 0, -1, 0, 0, 0, 0, 1 };
    const float Ri[9] = { 0, -1, 0, 1, 0, 0, 0, 0, 1 };             // inverse = transpose
    for (int i = 0; i < 9; ++i) { e.rot[i] = R[i]; e.inv_rot[i] = Ri[i]; }
    e.cutoff = (float)(.3 * kTwoPi / 360.0); e.falloff = (float)(.15 * kTwoPi / 360.0);
    // wavenumber distribution of the (emission x sensitivity) product: one discrete line (scene_sensor.hpp:31-148)
    wtgpu_kdist kd; std::memset(&kd, 0, sizeof(kd));
    const float prod = 900.f * 1.f;
    kd.type = WTGPU_KDIST_DISCRETE; kd.n = 1; kd.first = 0; kd.norm = 1.f / prod;
    const std::vector<float> kdist_data = { k, prod, 0.f, 1.f };    // k[1], y[1], dcdf[2]
    const float emitter_cdf[2] = { 0.f, 1.f };
    d.n_emitters = 1; d.emitters = &e; d.emitter_cdf = emitter_cdf; d.emitter_kdist = &kd;
    d.n_kdist_data = (uint32_t)kdist_data.size(); d.kdist_data = kdist_data.data();

    // ---- integrator: plt_path forward, max_depth 12, RR off, UTD on (src/integrator/plt_path.cpp:64-94)
    d.integrator.type = WTGPU_INTEGRATOR_PLT_PATH; d.integrator.direction = WTGPU_DIRECTION_FORWARD;
    d.integrator.max_depth = 12; d.integrator.russian_roulette = 0; d.integrator.fsd = 1;

    if (wtgpu_device_count() < 1) { std::fprintf(stderr, "no CUDA device: the product path has no CPU fallback\n"); wthost_ads_destroy(ads); return 3; }
    wtgpu_scene* scene = nullptr;
    CHECK(wtgpu_scene_create(&d, 0, &scene));

    // ---- render in two sample-range slices (what an interruptible driver does), host films accumulate
    const size_t n = (size_t)s.width * s.height;
    std::vector<float> block(n * 2, 0.f), light(n, 0.f), image(n, 0.f);
    wtgpu_render_opts o; std::memset(&o, 0, sizeof(o));
    o.seed = 0x5EED; o.spp = spp; o.tile_x1 = s.width; o.tile_y1 = s.height; o.device = 0; o.sampler = WTGPU_SAMPLER_UNIFORM;
    unsigned long long samples = 0; double ms = 0;
    const uint32_t cuts[3] = { 0u, spp / 2u, spp };
    for (int part = 0; part < 2; ++part) {
        o.sample_begin = cuts[part]; o.sample_end = cuts[part + 1];
        if (o.sample_end <= o.sample_begin) continue;
        wtgpu_stats st; std::memset(&st, 0, sizeof(st));
        CHECK(wtgpu_render(scene, &o, block.data(), light.data(), &st));
        samples += st.samples; ms += st.gpu_ms;
    }
    CHECK(wtgpu_develop(&d.sensor, spp, block.data(), light.data(), image.data()));
    double sum = 0; size_t lit = 0;
    for (float v : image) { sum += v; lit += v > 0.f; }
    std::printf("host_render: film %ux%u spp %u samples %llu sum %.9e lit %zu gpu_ms %.2f\n", s.width, s.height, spp, samples, sum, lit, ms);
    if (out_path) { FILE* f = std::fopen(out_path, "wb"); if (f) { std::fwrite(image.data(), sizeof(float), image.size(), f); std::fclose(f); } }
    wtgpu_scene_destroy(scene);
    wthost_ads_destroy(ads);
    return 0;
}
