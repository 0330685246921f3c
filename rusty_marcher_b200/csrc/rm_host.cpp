// rm_host.cpp -- host-side scene builder (include/rm_b200_host.h): the constructors of the
// reference's engine crate restated in f64 so that scenes built here carry exactly the values
// the reference computes (precomputed triangle normals/centres, polygon plane, L-inf normalised
// light colours, index-gradient OBJ colours) before they are flattened for the GPU.
// Compiled with -ffp-contract=off: no FMA may sneak into these f64 precomputations.
#include <cmath>
#include <cstdlib>
#include <cstdio>
#include <cstring>
#include <fstream>
#include <map>
#include <sstream>
#include <string>
#include <vector>

#include "../../include/rm_b200_host.h"
#include "rm_scene.h"

namespace {

struct D3 {
    double x, y, z;
};
inline D3 sub(D3 a, D3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline D3 add(D3 a, D3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }
inline D3 mul(D3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }
inline D3 cross(D3 a, D3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
inline D3 unit(D3 a) {                                       // geometry.rs:104-109
    double n = std::sqrt(a.x * a.x + a.y * a.y + a.z * a.z);
    return n > 0. ? mul(a, 1. / n) : a;
}
inline D3 ld(const double* p) { return {p[0], p[1], p[2]}; }
inline void st(D3 v, double* p) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// triangle.rs:33-47
RmTriangle make_triangle(D3 a, D3 b, D3 c) {
    RmTriangle t;
    st(a, t.vertices);
    st(b, t.vertices + 3);
    st(c, t.vertices + 6);
    st(mul(add(add(a, b), c), 1. / 3.), t.center);
    st(unit(cross(sub(b, a), sub(c, b))), t.normal);
    return t;
}

struct WavefrontModel {
    std::string name;
    std::vector<float> corners;   // 9 f32 per triangle (tobj stores f32 positions)
};

// Minimal reader with the behaviour of the `tobj` crate that the reference calls with
// LoadOptions{single_index, triangulate, ignore_points, ignore_lines} (obj.rs:44-58).
bool read_wavefront(const std::string& path, std::vector<WavefrontModel>& out) {
    std::ifstream file(path);
    if (!file) return false;
    const size_t cut = path.find_last_of('/');
    const std::string folder = cut == std::string::npos ? "" : path.substr(0, cut + 1);

    std::vector<float> xyz;
    std::vector<std::vector<long>> pending;
    std::map<std::string, int> materials;
    std::string current = "unnamed_object";
    int material = -1;

    auto emit = [&]() {
        WavefrontModel m;
        m.name = current;
        for (const auto& face : pending) {
            for (size_t k = 2; k < face.size(); k++) {          // fan: (0,1,2), (0,2,3), ...
                const long tri[3] = {face[0], face[k - 1], face[k]};
                for (long v : tri) m.corners.insert(m.corners.end(), {xyz[3 * v], xyz[3 * v + 1], xyz[3 * v + 2]});
            }
        }
        pending.clear();
        out.push_back(std::move(m));
    };

    for (std::string row; std::getline(file, row);) {
        std::istringstream words(row);
        std::string key;
        if (!(words >> key)) continue;
        if (key == "v") {
            int got = 0;
            for (std::string w; got < 3 && (words >> w); got++) xyz.push_back(std::strtof(w.c_str(), nullptr));
            for (; got < 3; got++) xyz.push_back(0.f);
        } else if (key == "f") {
            std::vector<long> face;
            const long count = (long)xyz.size() / 3;
            for (std::string w; words >> w;) {
                long i = std::strtol(w.c_str(), nullptr, 10);
                i = i < 0 ? count + i : i - 1;
                if (i < 0 || i >= count) return false;
                face.push_back(i);
            }
            pending.push_back(std::move(face));
        } else if (key == "o" || key == "g") {
            if (!pending.empty()) emit();
            std::string rest;
            std::getline(words, rest);
            const size_t b = rest.find_first_not_of(" \t\r"), e = rest.find_last_not_of(" \t\r");
            current = b == std::string::npos ? "unnamed_object" : rest.substr(b, e - b + 1);
        } else if (key == "mtllib") {
            for (std::string lib; words >> lib;) {
                std::ifstream mtl(folder + lib);
                for (std::string mrow; std::getline(mtl, mrow);) {
                    std::istringstream mw(mrow);
                    std::string mkey, mname;
                    if ((mw >> mkey) && mkey == "newmtl" && (mw >> mname)) {
                        int next = (int)materials.size();
                        materials[mname] = next;
                    }
                }
            }
        } else if (key == "usemtl") {
            std::string mname;
            if (words >> mname) {
                auto it = materials.find(mname);
                const int wanted = it == materials.end() ? -1 : it->second;
                if (wanted != material && !pending.empty()) emit();   // a material change splits the model
                material = wanted;
            }
        }
    }
    emit();
    return true;
}

}  // namespace

struct RmSceneBuilder {
    rm::OwnedFlatScene s;
    std::vector<std::string> names;   // per shape
    double camera[3] = {0., 0., 0.};
    RmFlatScene flat{};
    int n_prims = 0;

    int push_shape(int kind, int index, const std::string& name, int prims) {
        s.shapes.push_back({kind, index});
        names.push_back(name);
        n_prims += prims;
        return (int)s.shapes.size() - 1;
    }

    int add_triangles(const std::string& name, const std::vector<RmTriangle>& tris, const double* offset) {
        RmObj o{(int32_t)s.triangles.size(), (int32_t)tris.size()};
        const double n = (double)tris.size();
        for (size_t t = 0; t < tris.size(); t++) {
            RmTriangle tr = tris[t];
            if (offset) {                                       // triangle.rs:19-24: centre and vertices move, normal stays
                for (int k = 0; k < 3; k++) {
                    tr.center[k] += offset[k];
                    for (int v = 0; v < 3; v++) tr.vertices[3 * v + k] += offset[k];
                }
            }
            s.triangles.push_back(tr);
            RmReflectance r;
            rm_reflectance_default(&r);
            const double t_f = (double)t;                       // obj.rs:125-138
            r.diffuse_color[0] = 1. - t_f / n;
            r.diffuse_color[1] = t_f / n;
            r.diffuse_color[2] = 1.;
            s.triangle_reflectances.push_back(r);
        }
        s.objs.push_back(o);
        return push_shape(RM_SHAPE_OBJ, (int)s.objs.size() - 1, name, (int)tris.size());
    }
};

extern "C" {

void rm_reflectance_default(RmReflectance* r) {              // shapes.rs:50-60
    r->diffusion = 1.;
    r->diffuse_color[0] = r->diffuse_color[1] = r->diffuse_color[2] = 1.;
    r->specular = 1.;
    r->specular_exponent = 30.;
    r->is_glass_like = 0;
    r->reflection = 0.95;
    r->refractive_index = 1.;
}

RmSceneBuilder* rm_builder_new(void) { return new RmSceneBuilder(); }
void rm_builder_free(RmSceneBuilder* b) { delete b; }

void rm_builder_set_camera(RmSceneBuilder* b, const double xyz[3]) { std::memcpy(b->camera, xyz, sizeof b->camera); }
void rm_builder_offset_camera(RmSceneBuilder* b, const double xyz[3]) {
    for (int k = 0; k < 3; k++) b->camera[k] += xyz[k];
}
void rm_builder_get_camera(const RmSceneBuilder* b, double xyz[3]) { std::memcpy(xyz, b->camera, sizeof b->camera); }

int rm_builder_add_sphere(RmSceneBuilder* b, const double center[3], double radius, const RmReflectance* r) {
    if (!b || !center) return RM_ERR_INVALID_ARGUMENT;
    RmSphere s;
    std::memcpy(s.center, center, sizeof s.center);
    s.radius_square = radius * radius;                         // sphere.rs:16
    if (r) s.reflectance = *r; else rm_reflectance_default(&s.reflectance);
    b->s.spheres.push_back(s);
    return b->push_shape(RM_SHAPE_SPHERE, (int)b->s.spheres.size() - 1, "", 1);
}

int rm_builder_add_polygon(RmSceneBuilder* b, const double* v, int n, const RmReflectance* r) {
    if (!b || !v) return RM_ERR_INVALID_ARGUMENT;
    if (n < 3) return RM_ERR_SCENE;                            // polygon.rs:18
    RmPolygon p;
    p.first_vertex = (int32_t)(b->s.polygon_vertices.size() / 3);
    p.n_vertices = n;
    D3 mean = {0., 0., 0.};
    for (int i = 0; i < n; i++) mean = add(mean, ld(v + 3 * i));             // polygon.rs:25-28
    mean = mul(mean, 1. / (double)n);                                        // polygon.rs:29
    st(mean, p.plane_point);
    st(unit(cross(sub(ld(v + 3), ld(v)), sub(ld(v + 6), ld(v + 3)))), p.plane_normal);   // polygon.rs:32-38
    if (r) p.reflectance = *r; else rm_reflectance_default(&p.reflectance);
    b->s.polygon_vertices.insert(b->s.polygon_vertices.end(), v, v + 3 * n);
    b->s.polygons.push_back(p);
    return b->push_shape(RM_SHAPE_POLYGON, (int)b->s.polygons.size() - 1, "", 1);
}

int rm_builder_add_mesh(RmSceneBuilder* b, const double* tv, int n_triangles, const double offset[3]) {
    if (!b || (!tv && n_triangles) || n_triangles < 0) return RM_ERR_INVALID_ARGUMENT;
    std::vector<RmTriangle> tris;
    tris.reserve(n_triangles);
    for (int t = 0; t < n_triangles; t++) tris.push_back(make_triangle(ld(tv + 9 * t), ld(tv + 9 * t + 3), ld(tv + 9 * t + 6)));
    return b->add_triangles("mesh", tris, offset);
}

int rm_builder_add_triangles(RmSceneBuilder* b, const RmTriangle* triangles, const RmReflectance* reflectances,
                             int n_triangles, const char* name) {
    if (!b || (!triangles && n_triangles) || n_triangles < 0) return RM_ERR_INVALID_ARGUMENT;
    std::vector<RmTriangle> tris(triangles, triangles + n_triangles);
    const size_t first = b->s.triangle_reflectances.size();
    int shape = b->add_triangles(name ? name : "mesh", tris, nullptr);
    if (reflectances)
        for (int t = 0; t < n_triangles; t++) b->s.triangle_reflectances[first + t] = reflectances[t];
    return shape;
}

int rm_builder_add_obj_file(RmSceneBuilder* b, const char* path, const double offset[3]) {
    if (!b || !path) return RM_ERR_INVALID_ARGUMENT;
    std::vector<WavefrontModel> models;
    if (!read_wavefront(path, models)) return RM_ERR_INVALID_ARGUMENT;      // obj.rs:53-56: "Could not load obj"
    for (const auto& m : models) {
        std::vector<RmTriangle> tris;
        for (size_t t = 0; t + 9 <= m.corners.size(); t += 9) {
            const float* c = &m.corners[t];                                  // obj.rs:102-106: f32 widened to f64
            tris.push_back(make_triangle({(double)c[0], (double)c[1], (double)c[2]}, {(double)c[3], (double)c[4], (double)c[5]},
                                         {(double)c[6], (double)c[7], (double)c[8]}));
        }
        b->add_triangles(m.name, tris, offset);
    }
    return (int)models.size();
}

void rm_builder_add_light(RmSceneBuilder* b, const double position[3], const double color[3], double intensity) {
    RmLight l;
    std::memcpy(l.position, position, sizeof l.position);
    const double m = std::fmax(std::fmax(color[0], color[1]), color[2]);    // geometry.rs:111-116
    for (int k = 0; k < 3; k++) l.color[k] = m > 0. ? color[k] * (1. / m) : color[k];
    l.intensity = intensity;
    b->s.lights.push_back(l);
}

int rm_builder_num_shapes(const RmSceneBuilder* b) { return (int)b->s.shapes.size(); }
int rm_builder_num_prims(const RmSceneBuilder* b) { return b->n_prims; }
const char* rm_builder_shape_name(const RmSceneBuilder* b, int shape) {
    return (shape >= 0 && shape < (int)b->names.size()) ? b->names[shape].c_str() : "";
}

const RmFlatScene* rm_builder_flatten(RmSceneBuilder* b) {
    b->flat = b->s.view();
    return &b->flat;
}

int rm_builder_upload(RmSceneBuilder* b, RmScene* out_handle) {
    if (!b) return RM_ERR_INVALID_ARGUMENT;
    return rm_scene_upload(rm_builder_flatten(b), out_handle);
}

// framebuffer.rs:26-38
int rm_write_ppm(const char* path, int width, int height, const uint8_t* rgb8) {
    if (!path || !rgb8 || width <= 0 || height <= 0) return RM_ERR_INVALID_ARGUMENT;
    std::FILE* f = std::fopen(path, "wb");
    if (!f) return RM_ERR_INVALID_ARGUMENT;
    std::fprintf(f, "P6\n%d %d\n255\n", width, height);       // framebuffer.rs:32
    const size_t n = (size_t)width * (size_t)height * 3;
    const bool ok = std::fwrite(rgb8, 1, n, f) == n;
    return (std::fclose(f) == 0 && ok) ? RM_OK : RM_ERR_INVALID_ARGUMENT;
}

// scene.rs:28-211.  The reference threads ONE mutable Reflectance through the whole function, so
// each shape inherits whatever the previous ones left in it; `r` below is mutated the same way.
RmSceneBuilder* rm_builder_create_default(void) {
    RmSceneBuilder* b = new RmSceneBuilder();
    RmReflectance r;
    rm_reflectance_default(&r);
    auto colour = [&r](double x, double y, double z) { r.diffuse_color[0] = x; r.diffuse_color[1] = y; r.diffuse_color[2] = z; };

    colour(0.8, 0., 0.);
    r.specular_exponent = 100.;
    const RmReflectance red = r;

    colour(0.6, 0., 0.7);
    const RmReflectance tri = r;

    r.diffusion = 1.0;
    r.specular = 1.;
    r.is_glass_like = 1;
    r.refractive_index = 1.5;
    r.reflection = 0.5;
    colour(0.3, 0.9, 0.9);
    const RmReflectance floor_r = r;

    r.specular = 1.0;
    r.diffusion = 0.1;
    colour(0., 0., 0.2);
    r.is_glass_like = 1;
    r.refractive_index = 1.5;
    r.reflection = 0.2;
    const RmReflectance blue = r;

    r.diffusion = 1.;
    r.reflection = 1.;
    r.is_glass_like = 0;
    r.specular = 0.8;
    colour(0., 1., 0.);
    const RmReflectance green = r;

    colour(0.9, 0.9, 0.9);
    const RmReflectance white = r;

    // shape order of scene.rs:201-208: blue, green, red, white spheres, triangle, floor
    const double c_blue[3] = {-0.5, -1.5, -5.}, c_green[3] = {6., -0.5, -18.}, c_red[3] = {-5., 0., -16.},
                 c_white[3] = {-10., 6., -14.};
    rm_builder_add_sphere(b, c_blue, 2., &blue);
    rm_builder_add_sphere(b, c_green, 3., &green);
    rm_builder_add_sphere(b, c_red, 4., &red);
    rm_builder_add_sphere(b, c_white, 4., &white);
    const double v_tri[9] = {7., -4., -8., 15., 0., -9., 6., 3., -8.};
    rm_builder_add_polygon(b, v_tri, 3, &tri);
    const double v_floor[12] = {20., -3., -50., -20., -3., -50., -15., -6., -3., 15., -6., -3.};
    rm_builder_add_polygon(b, v_floor, 4, &floor_r);

    const double origin[3] = {0., 0., 0.}, white_l[3] = {1., 1., 1.};
    const double far_p[3] = {20., 20., 20.}, reddish[3] = {1., 0.5, 0.5};
    rm_builder_add_light(b, origin, white_l, 1.);
    rm_builder_add_light(b, far_p, reddish, 0.8);
    return b;
}

}  // extern "C"
