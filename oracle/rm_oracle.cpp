// rm_oracle.cpp -- CPU ORACLE (test infrastructure, never shipped, never on the product path).
//
// A plain f64 C++17 restatement of rusty-marcher's per-pixel render hot path.  Every function
// cites the reference file:line (relative to /root/reference) whose arithmetic it follows, in
// the same evaluation order (left-to-right dot products, multiply-by-reciprocal normalisation,
// strict comparisons), so that `g++ -O2 -ffp-contract=off` reproduces the reference's f64
// results.  Pinned against the reference's golden engine/out.ppm (see rm_oracle.h).
//
// Besides the image it emits what the parity harness needs and the reference does not produce:
// the primary-ray primitive id, a per-pixel "fragile" (near-tie) mask and the event counters
// that define the algorithmic work.
#include "rm_oracle.h"

#include <atomic>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <map>
#include <memory>
#include <sstream>
#include <string>
#include <thread>
#include <vector>

namespace {

// ---------------------------------------------------------------- geometry.rs:4-182
struct V3 {
    double x, y, z;
};
inline V3 operator+(V3 a, V3 b) { return {a.x + b.x, a.y + b.y, a.z + b.z}; }   // geometry.rs:118-128
inline V3 operator-(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }   // geometry.rs:163-172
inline V3 operator*(V3 a, V3 b) { return {a.x * b.x, a.y * b.y, a.z * b.z}; }   // geometry.rs:152-161
inline V3 operator-(V3 a) { return {-a.x, -a.y, -a.z}; }                        // geometry.rs:130-140
inline V3 scaled(V3 a, double s) { return {a.x * s, a.y * s, a.z * s}; }        // geometry.rs:35-53
inline double dot(V3 a, V3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }     // geometry.rs:180-182
inline double squared_norm(V3 a) { return dot(a, a); }                          // geometry.rs:65-67
inline V3 cross(V3 a, V3 b) {                                                   // geometry.rs:57-63
    return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x};
}
inline V3 normalized(V3 a) {                                                    // geometry.rs:104-109
    double norm = std::sqrt(dot(a, a));
    if (norm > 0.) a = scaled(a, 1. / norm);
    return a;
}
inline V3 normalized_l0(V3 a) {                                                 // geometry.rs:111-116
    double norm = std::fmax(std::fmax(a.x, a.y), a.z);
    if (norm > 0.) a = scaled(a, 1. / norm);
    return a;
}
inline V3 from3(const double* p) { return {p[0], p[1], p[2]}; }
inline void to3(V3 v, double* p) { p[0] = v.x; p[1] = v.y; p[2] = v.z; }

// ---------------------------------------------------------------- shapes.rs:4-61
struct Reflectance {
    double diffusion = 1.;
    V3 diffuse_color{1., 1., 1.};
    double specular = 1.;
    double specular_exponent = 30.;
    bool is_glass_like = false;
    double reflection = 0.95;
    double refractive_index = 1.;
};
struct Hit {
    V3 point{0, 0, 0}, normal{0, 0, 0};
    Reflectance refl;
    int prim = -1;   // oracle-defined flattened primitive index
};

// Near-tie margins: a decision is "fragile" when its margin is within kF32 (32 ulp of binary32)
// of the magnitudes that produced it -- i.e. an FP32 evaluation could legitimately flip it.
constexpr double kF32 = 32. * 1.1920928955078125e-07;

// Instrumentation state of one worker thread.  Null on the timed CPU-baseline path.
struct Probe {
    OrcCounters c{};
    uint8_t fragile = 0;    // accumulated for the current pixel
    bool primary = false;   // currently inside the primary closest-hit query
    bool counting = true;   // false once an any-hit Obj scan has found its first hit
    bool muted = false;     // inside a degenerate-projection primitive: margins are not meaningful
    uint64_t degenerate_hits = 0;
    double scale = 1.;      // largest |coordinate| in the scene
    double best = 0, second = 0;
    int ncand = 0;
    void mark() { if (!muted) fragile |= primary ? 3 : 2; }
    // 32 ulp on the primary query (excuses prim-id mismatches, must be conservative), 8 ulp deeper
    // in the tree (diagnostic for colour differences only)
    void near(double v, double tol) { if (std::fabs(v) < (primary ? tol : 0.25 * tol)) mark(); }
    void cand(double d2) {
        if (ncand == 0) best = d2;
        else if (d2 < best) { second = best; best = d2; }
        else if (ncand == 1 || d2 < second) second = d2;
        ncand++;
    }
    void begin_segment() { ncand = 0; counting = true; }
    void end_closest() { if (ncand > 1 && (second - best) < 2. * kF32 * best) mark(); }
};
#define CNT(field) do { if (T && pr->counting && trail.alive) pr->c.field++; } while (0)
#define CNT0(field) do { if (T && pr->counting) pr->c.field++; } while (0)

// The decisions of one ray/primitive test.  Untracked, a failed decision ends the test (the
// reference's early return).  Tracked, a failed decision whose margin is within FP32 resolution
// lets the evaluation continue in "what-if" mode (nothing is counted any more): if every later
// decision passes or is itself marginal, the primitive's hit/miss outcome hinges on a near-tie
// and the pixel is marked fragile.  A robust failure anywhere makes the primitive a robust miss.
template <bool T> struct Trail {
    Probe* pr;
    bool alive = true, frag = false;
    explicit Trail(Probe* p) : pr(p) {}
    bool step(bool pass, double margin, double tol) {
        if (!T) return pass;
        const bool fr = std::fabs(margin) < (pr->primary ? tol : 0.25 * tol);
        if (fr) frag = true;
        if (!pass) { alive = false; if (!fr) return false; }
        return true;
    }
    bool finish() { if (T && frag) pr->mark(); return alive; }
};

struct Shape {
    int prim_base = 0;
    virtual ~Shape() {}
    virtual int n_prims() const { return 1; }
    virtual bool intersect(const V3& o, const V3& d, Hit& h) const = 0;
    virtual bool intersect_probe(const V3& o, const V3& d, Hit& h, Probe* pr, bool anyhit) const = 0;
    virtual double max_abs() const = 0;
    virtual bool is_obj() const { return false; }
    // materials of the shape's primitives, in primitive order (extension-mode helpers below; not used by the restatement)
    virtual void materials(std::vector<Reflectance*>& out) = 0;
};

inline double max_abs3(V3 v) { return std::fmax(std::fabs(v.x), std::fmax(std::fabs(v.y), std::fabs(v.z))); }

// ---------------------------------------------------------------- sphere.rs:6-61
struct Sphere : Shape {
    V3 center;
    double radius_square;
    Reflectance reflectance;
    void materials(std::vector<Reflectance*>& out) override { out.push_back(&reflectance); }
    template <bool T>
    bool isect(const V3& o, const V3& d, Hit& h, Probe* pr) const {
        Trail<T> trail(pr);
        V3 line = center - o;                                  // sphere.rs:28
        double tca = dot(line, d);                             // sphere.rs:33
        double d2 = dot(line, line) - tca * tca;               // sphere.rs:34
        CNT(sphere_tests);
        if (!trail.step(!(d2 > radius_square), d2 - radius_square, kF32 * std::fmax(radius_square, dot(line, line))))
            return false;                                      // sphere.rs:36-38
        CNT(sphere_disc);
        double thc = std::sqrt(std::fmax(radius_square - d2, 0.));   // sphere.rs:40 (the clamp only matters in what-if mode)
        if (!T) thc = std::sqrt(radius_square - d2);
        double t0 = tca - thc;                                 // sphere.rs:42
        double t1 = tca + thc;                                 // sphere.rs:43
        const double ttol = kF32 * std::sqrt(dot(line, line));
        if (t0 < 0.) {                                         // sphere.rs:45-47
            if (T && std::fabs(t0) < (pr->primary ? ttol : 0.25 * ttol)) trail.frag = true;   // which root is taken is marginal
            t0 = t1;
        }
        if (!trail.step(!(t0 < 0.), t0, ttol)) return false;   // sphere.rs:49-51
        if (T && !trail.alive) { trail.finish(); return false; }
        CNT(sphere_hits);
        V3 p = o + scaled(d, t0);                              // sphere.rs:54
        h.point = p;
        h.normal = normalized(p - center);                     // sphere.rs:58
        h.refl = reflectance;
        h.prim = prim_base;
        trail.finish();
        return true;
    }
    bool intersect(const V3& o, const V3& d, Hit& h) const override { return isect<false>(o, d, h, nullptr); }
    bool intersect_probe(const V3& o, const V3& d, Hit& h, Probe* pr, bool) const override { return isect<true>(o, d, h, pr); }
    double max_abs() const override { return max_abs3(center) + std::sqrt(radius_square); }
};

// A planar primitive whose XY projection has (almost) no area -- normal.z == 0, or 5e-9 as for the
// dodecahedron faces that tobj's f32 vertices tilt slightly -- can never pass the z-only inside test
// in exact arithmetic; what the reference computes for it is rounding noise.  The oracle counts any
// hit on such a primitive (tests assert there are none) and does not let its meaningless margins
// pollute the near-tie mask.
inline bool degenerate_projection(const V3* v, size_t n) {
    double area2 = 0., perimeter = 0.;
    for (size_t i = 0; i < n; i++) {
        size_t j = (i + 1) % n;
        area2 += v[i].x * v[j].y - v[i].y * v[j].x;
        perimeter += std::hypot(v[j].x - v[i].x, v[j].y - v[i].y);
    }
    return std::fabs(area2) <= 1e-6 * perimeter * perimeter;
}

// shared by triangle.rs:13-15 and polygon.rs:54-56: only the z component of the cross product
template <bool T>
inline bool inside(const V3& a, const V3& p1, const V3& p2, Probe* pr, Trail<T>& trail) {
    V3 u = p1 - a, v = p2 - a;
    double cz = cross(u, v).z;
    CNT(edge_tests);
    double tol = 0.;
    if (T) {
        double nu = std::sqrt(squared_norm(u)), nv = std::sqrt(squared_norm(v));
        tol = kF32 * (nu * nv + (max_abs3(a) + max_abs3(p1)) * (nu + nv));
    }
    return trail.step(cz > 0., cz, tol);
}

// ---------------------------------------------------------------- triangle.rs:6-83
struct Triangle {
    V3 v[3];
    V3 normal, center;
    bool degenerate = false;
    static Triangle create(V3 a, V3 b, V3 c) {                 // triangle.rs:33-47
        Triangle t;
        t.v[0] = a; t.v[1] = b; t.v[2] = c;
        t.center = scaled(a + b + c, 1. / 3.);
        V3 edge_1 = b - a;
        V3 edge_2 = c - b;
        t.normal = normalized(cross(edge_1, edge_2));
        t.degenerate = degenerate_projection(t.v, 3);
        return t;
    }
    void offset(V3 off) {                                      // triangle.rs:19-24 (normal untouched)
        center = center + off;
        for (auto& p : v) p = p + off;
    }
    template <bool T>
    bool isect(const V3& o, const V3& d, V3& point, Probe* pr) const {
        if (T) pr->muted = degenerate;
        bool got = isect_impl<T>(o, d, point, pr);
        if (T) { pr->muted = false; if (got && degenerate) pr->degenerate_hits++; }
        return got;
    }
    template <bool T>
    bool isect_impl(const V3& o, const V3& d, V3& point, Probe* pr) const {
        Trail<T> trail(pr);
        double dot_product = dot(d, normal);                   // triangle.rs:56
        CNT(plane_tests);
        if (!trail.step(!(std::fabs(dot_product) < 1e-6), std::fabs(dot_product) - 1e-6, kF32)) return false;   // triangle.rs:57-59
        double dist = dot(center - o, normal) / dot_product;   // triangle.rs:62
        CNT(plane_dist);
        if (!trail.step(!(dist < 0.), dist, kF32 * (max_abs3(center) + max_abs3(o)) / std::fabs(dot_product)))
            return false;                                      // triangle.rs:65-67
        V3 p = o + scaled(d, dist);                            // triangle.rs:69
        CNT(plane_point);
        for (int i = 0; i < 3; i++)                            // triangle.rs:72-76
            if (!inside<T>(p, v[i], v[(i + 1) % 3], pr, trail)) return false;
        point = p;
        return trail.finish();
    }
};

// ---------------------------------------------------------------- polygon.rs:6-98
struct ConvexPolygon : Shape {
    std::vector<V3> vertices;
    Reflectance reflectance;
    void materials(std::vector<Reflectance*>& out) override { out.push_back(&reflectance); }
    V3 plane_normal, plane_point;
    bool degenerate = false;
    static std::unique_ptr<ConvexPolygon> create(std::vector<V3> verts, const Reflectance& r) {  // polygon.rs:16-42
        auto p = std::make_unique<ConvexPolygon>();
        V3 mean{0, 0, 0};
        for (auto& v : verts) mean = mean + v;
        mean = scaled(mean, 1. / (double)verts.size());
        V3 edge_1 = verts[1] - verts[0];
        V3 edge_2 = verts[2] - verts[1];
        p->plane_normal = normalized(cross(edge_1, edge_2));
        p->plane_point = mean;
        p->vertices = std::move(verts);
        p->degenerate = degenerate_projection(p->vertices.data(), p->vertices.size());
        p->reflectance = r;
        return p;
    }
    template <bool T>
    bool isect(const V3& o, const V3& d, Hit& h, Probe* pr) const {
        if (T) pr->muted = degenerate;
        bool got = isect_impl<T>(o, d, h, pr);
        if (T) { pr->muted = false; if (got && degenerate) pr->degenerate_hits++; }
        return got;
    }
    template <bool T>
    bool isect_impl(const V3& o, const V3& d, Hit& h, Probe* pr) const {
        Trail<T> trail(pr);
        double dotprod = dot(d, plane_normal);                 // polygon.rs:65
        CNT(plane_tests);
        if (!trail.step(!(dotprod == 0.), dotprod, kF32)) return false;     // polygon.rs:66-68
        double dist = dot(plane_point - o, plane_normal) / dotprod;   // polygon.rs:71
        CNT(plane_dist);
        if (!trail.step(!(dist < 0.), dist, kF32 * (max_abs3(plane_point) + max_abs3(o)) / std::fmax(std::fabs(dotprod), 1e-300)))
            return false;                                      // polygon.rs:74-76
        V3 p = o + scaled(d, dist);                            // polygon.rs:78
        CNT(plane_point);
        size_t n = vertices.size();
        for (size_t i = 0; i < n; i++)                         // polygon.rs:83-91
            if (!inside<T>(p, vertices[i], vertices[(i + 1) % n], pr, trail)) return false;
        if (!trail.finish()) return false;
        h.point = p;
        h.normal = plane_normal;
        h.refl = reflectance;
        h.prim = prim_base;
        return true;
    }
    bool intersect(const V3& o, const V3& d, Hit& h) const override { return isect<false>(o, d, h, nullptr); }
    bool intersect_probe(const V3& o, const V3& d, Hit& h, Probe* pr, bool) const override { return isect<true>(o, d, h, pr); }
    double max_abs() const override {
        double m = 0;
        for (auto& v : vertices) m = std::fmax(m, max_abs3(v));
        return m;
    }
};

// ---------------------------------------------------------------- obj.rs:13-29,185-221
struct Obj : Shape {
    std::string name;
    std::vector<Triangle> triangles;
    std::vector<Reflectance> reflectances;
    int n_prims() const override { return (int)triangles.size(); }
    bool is_obj() const override { return true; }
    void materials(std::vector<Reflectance*>& out) override { for (auto& r : reflectances) out.push_back(&r); }
    void offset(V3 off) { for (auto& t : triangles) t.offset(off); }   // obj.rs:24-29
    void gradient_colours() {                                          // obj.rs:125-138
        size_t n = triangles.size();
        reflectances.assign(n, Reflectance());
        for (size_t t = 0; t < n; t++) {
            double t_f = (double)t;
            reflectances[t].diffuse_color = {1. - t_f / (double)n, t_f / (double)n, 1.};
        }
    }
    template <bool T>
    bool isect(const V3& o, const V3& d, Hit& h, Probe* pr, bool anyhit) const {
        bool hit_triangle = false;                             // obj.rs:190
        double dist_closest = 0.;
        for (size_t i = 0; i < triangles.size(); i++) {        // obj.rs:194
            V3 p;
            if (triangles[i].isect<T>(o, d, p, pr)) {
                double dist_hit = squared_norm(p - o);         // obj.rs:197
                CNT0(cand_dist);
                if (T && !anyhit) pr->cand(dist_hit);
                if (!hit_triangle || dist_hit < dist_closest) {   // obj.rs:198
                    h.point = p;
                    h.normal = triangles[i].normal;
                    h.refl = reflectances[i];                  // obj.rs:203
                    h.prim = prim_base + (int)i;
                    hit_triangle = true;
                    dist_closest = dist_hit;
                }
                // The reference keeps scanning (obj.rs:194-210); the algorithmic count of an
                // any-hit query stops at the first hitting primitive.
                if (T && anyhit) pr->counting = false;
            }
        }
        return hit_triangle;
    }
    bool intersect(const V3& o, const V3& d, Hit& h) const override { return isect<false>(o, d, h, nullptr, false); }
    bool intersect_probe(const V3& o, const V3& d, Hit& h, Probe* pr, bool anyhit) const override {
        return isect<true>(o, d, h, pr, anyhit);
    }
    double max_abs() const override {
        double m = 0;
        for (auto& t : triangles) for (auto& v : t.v) m = std::fmax(m, max_abs3(v));
        return m;
    }
};

// ---------------------------------------------------------------- lights.rs:4-16
struct Light {
    V3 position, color;
    double intensity;
};
inline Light create_light(V3 position, V3 color, double intensity) {
    return {position, normalized_l0(color), intensity};
}

}  // namespace

// ---------------------------------------------------------------- scene.rs:9-27
struct OrcScene {
    std::vector<Light> lights;
    std::vector<std::unique_ptr<Shape>> shapes;
    V3 camera{0, 0, 0};
    int n_prims = 0;
    void push(std::unique_ptr<Shape> s) {
        s->prim_base = n_prims;
        n_prims += s->n_prims();
        shapes.push_back(std::move(s));
    }
};

namespace {

// ---------------------------------------------------------------- shapes.rs:92-108
template <bool T>
bool intersect_shape_set(const V3& o, const V3& d, const OrcScene& sc, Probe* pr) {
    if (T) { pr->c.anyhit_segments++; pr->begin_segment(); }
    Hit scratch;
    for (auto& shape : sc.shapes) {
        bool hit = T ? shape->intersect_probe(o, d, scratch, pr, true) : shape->intersect(o, d, scratch);
        if (hit) { if (T) pr->counting = true; return true; }
    }
    return false;
}

// ---------------------------------------------------------------- shapes.rs:110-143
template <bool T>
bool find_closest_intersect(const V3& o, const V3& d, const OrcScene& sc, Hit& out, Probe* pr) {
    if (T) { pr->c.closest_segments++; pr->begin_segment(); }
    bool hit = false;
    double dist_closest = 0.;
    Hit test;
    for (auto& shape : sc.shapes) {
        bool got = T ? shape->intersect_probe(o, d, test, pr, false) : shape->intersect(o, d, test);
        if (got) {
            double dist_hit = squared_norm(test.point - o);    // shapes.rs:128
            CNT0(cand_dist);
            if (T && !shape->is_obj()) pr->cand(dist_hit);
            if (!hit || dist_hit < dist_closest) {             // shapes.rs:130
                out = test;
                hit = true;
                dist_closest = dist_hit;
            }
        }
    }
    if (T) pr->end_closest();
    return hit;
}

// ---------------------------------------------------------------- optics.rs:4-6
inline V3 reflect(V3 incident, V3 normal) { return incident - scaled(normal, 2. * dot(incident, normal)); }

// ---------------------------------------------------------------- optics.rs:8-48
template <bool T>
bool reflect_ray(V3 incident, const Hit& hit, double refractive_index, V3& ro, V3& rd, Probe* pr) {
    V3 normal = hit.normal;
    double c = dot(normal, incident);                          // optics.rs:16
    if (T) pr->near(c, kF32);
    double r = (c < 0.) ? refractive_index : 1. / refractive_index;   // optics.rs:19-23
    if (c < 0.) { c = -c; normal = -normal; }                  // optics.rs:25-28
    double cos_theta_2 = 1. - r * r * (1. - c * c);            // optics.rs:30
    if (T) pr->near(cos_theta_2, kF32 * std::fmax(1., r * r));
    if (cos_theta_2 > 0.) return false;                        // optics.rs:33-35
    rd = reflect(incident, normal);                            // optics.rs:38
    double side = dot(rd, hit.normal);
    if (T) pr->near(side, kF32);
    if (side < 0.) ro = hit.point - scaled(hit.normal, 1e-4);  // optics.rs:41-45
    else ro = hit.point + scaled(hit.normal, 1e-4);
    return true;
}

// ---------------------------------------------------------------- optics.rs:50-89
template <bool T>
bool refract_ray(V3 incident, const Hit& hit, double refractive_index, V3& ro, V3& rd, Probe* pr) {
    V3 normal = hit.normal;
    double c = -dot(normal, incident);                         // optics.rs:57
    if (T) pr->near(c, kF32);
    double r = (c < 0.) ? refractive_index : 1. / refractive_index;   // optics.rs:60-64
    if (c < 0.) { c = -c; normal = -normal; }                  // optics.rs:66-69
    double cos_theta_2 = 1. - r * r * (1. - c * c);            // optics.rs:71
    if (T) pr->near(cos_theta_2, kF32 * std::fmax(1., r * r));
    if (cos_theta_2 < 0.) return false;                        // optics.rs:74-76
    rd = normalized(scaled(incident, r) + scaled(normal, r * c - std::sqrt(cos_theta_2)));   // optics.rs:78-79
    double side = dot(rd, normal);
    if (T) pr->near(side, kF32);
    if (side > 0.) ro = hit.point + scaled(normal, 1e-4);      // optics.rs:82-86
    else ro = hit.point - scaled(normal, 1e-4);
    return true;
}

// ---------------------------------------------------------------- renderer.rs:138-193
template <bool T>
V3 direct_lighting(const V3& origin, const Hit& hit, const OrcScene& sc, Probe* pr) {
    V3 light_intensity{0, 0, 0};
    for (auto& light : sc.lights) {
        V3 light_dir = normalized(light.position - hit.point);             // renderer.rs:166
        double side = dot(light_dir, hit.normal);
        if (T) {
            pr->c.light_evals++;
            double len = std::sqrt(squared_norm(light.position - hit.point));
            pr->near(side, kF32 * std::fmax(1., (max_abs3(light.position) + max_abs3(hit.point)) / std::fmax(len, 1e-300)));
        }
        V3 intersect_orig = (side < 0.) ? hit.point - scaled(hit.normal, 1e-3)   // renderer.rs:168-172
                                        : hit.point + scaled(hit.normal, 1e-3);
        if (intersect_shape_set<T>(intersect_orig, light_dir, sc, pr)) continue;  // renderer.rs:174-177
        if (T) pr->c.lit_lights++;
        double diffusion = std::fmax(dot(light_dir, hit.normal), 0.);      // renderer.rs:138-140,180
        light_intensity = light_intensity +
            scaled(scaled(light.color * hit.refl.diffuse_color, diffusion), light.intensity);   // renderer.rs:181-183
        V3 incident = -light_dir;                                          // renderer.rs:144
        V3 reflected = reflect(incident, hit.normal);                      // renderer.rs:145
        V3 dir_to_viewer = normalized(origin - hit.point);                 // renderer.rs:149
        double spec_f = std::fmax(dot(reflected, dir_to_viewer), 0.);      // renderer.rs:150
        double specular = std::pow(spec_f * hit.refl.specular, hit.refl.specular_exponent);   // renderer.rs:186-188
        light_intensity = light_intensity + scaled(light.color, specular); // renderer.rs:189
    }
    return scaled(light_intensity, hit.refl.diffusion);                    // renderer.rs:192
}

// ---------------------------------------------------------------- renderer.rs:195-309
template <bool T>
V3 cast_ray(const V3& orig, V3 dir, const OrcScene& sc, const V3& background, int n_recursion,
            int max_depth, Probe* pr, int* primary_prim) {
    if (n_recursion > max_depth) return background;                        // renderer.rs:262-264 (max_depth = 3)
    Hit hit;
    if (T) pr->primary = (n_recursion == 1);
    bool got = find_closest_intersect<T>(orig, dir, sc, hit, pr);          // renderer.rs:266
    if (T) pr->primary = false;
    if (primary_prim) *primary_prim = got ? hit.prim : -1;
    if (!got) return (n_recursion > 1) ? background : V3{0, 0, 0};         // renderer.rs:300-306
    if (T) pr->c.hits++;
    V3 light_intensity = background;                                       // renderer.rs:272
    light_intensity = light_intensity + direct_lighting<T>(orig, hit, sc, pr);   // renderer.rs:275
    if (hit.refl.is_glass_like) {                                          // renderer.rs:277
        if (T) pr->c.glass_hits++;
        V3 ro, rd;
        if (reflect_ray<T>(dir, hit, hit.refl.refractive_index, ro, rd, pr)) {    // renderer.rs:195-222
            if (T) pr->c.reflections++;
            light_intensity = light_intensity +
                scaled(cast_ray<T>(ro, rd, sc, background, n_recursion + 1, max_depth, pr, nullptr), hit.refl.reflection);
        }
        if (refract_ray<T>(dir, hit, hit.refl.refractive_index, ro, rd, pr)) {    // renderer.rs:225-252
            if (T) pr->c.refractions++;
            light_intensity = light_intensity +
                scaled(cast_ray<T>(ro, rd, sc, background, n_recursion + 1, max_depth, pr, nullptr), 1. - hit.refl.reflection);
        }
    }
    return light_intensity;
}

uint64_t g_last_degenerate_hits = 0;

void add_counters(OrcCounters& a, const OrcCounters& b) {
    uint64_t* pa = reinterpret_cast<uint64_t*>(&a);
    const uint64_t* pb = reinterpret_cast<const uint64_t*>(&b);
    for (size_t i = 0; i < sizeof(OrcCounters) / sizeof(uint64_t); i++) pa[i] += pb[i];
}

// ---------------------------------------------------------------- tobj restatement (obj.rs:44-58)
// tobj is an un-vendored dependency (engine/Cargo.toml:8, "*"; API shape => >= 3.0).  Restated
// behaviour: `v` positions parsed as f32; faces with negative (relative) indices; a model is
// emitted at each `o`/`g` when faces are pending, at a `usemtl` that changes the material id while
// faces are pending, and at end of file; quads -> (a,b,c),(a,c,d); n-gons -> fan; points and lines
// dropped.  PARITY UNPINNED: no reference test asserts triangle count or order.
struct ObjModel {
    std::string name;
    std::vector<float> tri_pos;   // 9 floats per triangle, f32 like tobj::Mesh::positions
};

bool load_mtl_names(const std::string& path, std::map<std::string, int>& mat_map, int& n_materials) {
    std::ifstream in(path);
    if (!in) return false;
    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tok;
        if (!(ls >> tok)) continue;
        if (tok == "newmtl") {
            std::string name;
            ls >> name;
            mat_map[name] = n_materials++;
        }
    }
    return true;
}

bool load_obj(const std::string& path, std::vector<ObjModel>& models) {
    std::ifstream in(path);
    if (!in) return false;
    std::string dir;
    size_t slash = path.find_last_of('/');
    if (slash != std::string::npos) dir = path.substr(0, slash + 1);

    std::vector<float> pos;
    std::vector<std::vector<int>> faces;   // pending faces of the current model (absolute 0-based indices)
    std::string name = "unnamed_object";
    std::map<std::string, int> mat_map;
    int n_materials = 0;
    int mat_id = -1;   // Option<usize>: -1 = None

    auto flush = [&]() {
        ObjModel m;
        m.name = name;
        for (auto& f : faces) {
            if (f.size() < 3) continue;            // ignore_points / ignore_lines
            for (size_t k = 2; k < f.size(); k++) {   // triangle, quad and n-gon fan alike
                int idx[3] = {f[0], f[k - 1], f[k]};
                for (int i : idx) {
                    m.tri_pos.push_back(pos[3 * i]);
                    m.tri_pos.push_back(pos[3 * i + 1]);
                    m.tri_pos.push_back(pos[3 * i + 2]);
                }
            }
        }
        models.push_back(std::move(m));
        faces.clear();
    };

    std::string line;
    while (std::getline(in, line)) {
        std::istringstream ls(line);
        std::string tok;
        if (!(ls >> tok)) continue;
        if (tok == "v") {
            std::string w;
            int n = 0;
            while (n < 3 && (ls >> w)) { pos.push_back(std::strtof(w.c_str(), nullptr)); n++; }
            while (n++ < 3) pos.push_back(0.f);
        } else if (tok == "f") {
            std::vector<int> f;
            std::string w;
            while (ls >> w) {
                long i = std::strtol(w.c_str(), nullptr, 10);   // leading "v" of "v/vt/vn"
                long n_pos = (long)pos.size() / 3;
                long abs_i = (i < 0) ? n_pos + i : i - 1;
                if (abs_i < 0 || abs_i >= n_pos) return false;
                f.push_back((int)abs_i);
            }
            faces.push_back(std::move(f));
        } else if (tok == "o" || tok == "g") {
            if (!faces.empty()) flush();
            std::string rest;
            std::getline(ls, rest);
            size_t b = rest.find_first_not_of(" \t\r"), e = rest.find_last_not_of(" \t\r");
            name = (b == std::string::npos) ? "unnamed_object" : rest.substr(b, e - b + 1);
        } else if (tok == "mtllib") {
            std::string lib;
            while (ls >> lib) load_mtl_names(dir + lib, mat_map, n_materials);
        } else if (tok == "usemtl") {
            std::string mat;
            if (ls >> mat) {
                auto it = mat_map.find(mat);
                int new_mat = (it == mat_map.end()) ? -1 : it->second;
                if (new_mat != mat_id && !faces.empty()) flush();
                mat_id = new_mat;
            }
        }
    }
    flush();   // the last model is pushed unconditionally
    return true;
}

std::unique_ptr<Obj> make_obj(const std::string& name, const double* tri, int n_tri) {
    auto o = std::make_unique<Obj>();
    o->name = name;
    o->triangles.reserve(n_tri);
    for (int t = 0; t < n_tri; t++)
        o->triangles.push_back(Triangle::create(from3(tri + 9 * t), from3(tri + 9 * t + 3), from3(tri + 9 * t + 6)));
    o->gradient_colours();
    return o;
}

Reflectance from_c(const OrcReflectance* r) {
    Reflectance o;
    if (!r) return o;
    o.diffusion = r->diffusion;
    o.diffuse_color = from3(r->diffuse_color);
    o.specular = r->specular;
    o.specular_exponent = r->specular_exponent;
    o.is_glass_like = r->is_glass_like != 0;
    o.reflection = r->reflection;
    o.refractive_index = r->refractive_index;
    return o;
}

template <bool T>
void render_patches(const OrcScene& sc, int W, int H, double fov, int max_depth, const std::vector<int>& patches,
                    std::atomic<int>& next, double* out_rgb, int32_t* prim_id, uint8_t* fragile, Probe* pr) {
    // renderer.rs:25-33
    const double half_fov = std::tan(fov / 2.);
    const double height = (double)H, width = (double)W;
    const double ratio = width / height;
    const V3 background{0.1, 0.1, 0.1};                        // renderer.rs:40-44
    const int patch_size = 32;                                  // renderer.rs:47
    const int n_width = W / patch_size;
    for (;;) {
        int k = next.fetch_add(1);
        if (k >= (int)patches.size()) break;
        int p = patches[k];
        int p_line = p % n_width * patch_size;                  // renderer.rs:69 (x)
        int p_col = p / n_width * patch_size;                   // renderer.rs:70 (y)
        for (int i = p_col; i < p_col + patch_size; i++) {
            for (int j = p_line; j < p_line + patch_size; j++) {
                // backproject(j, i), renderer.rs:128-135
                V3 dir = normalized(V3{2. * ((double)j / width - 0.5) * half_fov * ratio,
                                       -2. * ((double)i / height - 0.5) * half_fov, -1.});
                int prim = -1;
                if (T) { pr->fragile = 0; pr->c.pixels++; }
                V3 c = cast_ray<T>(sc.camera, dir, sc, background, 1, max_depth, pr, &prim);
                size_t px = (size_t)i * W + j;                  // renderer.rs:92-108 reassembly
                out_rgb[3 * px] = c.x; out_rgb[3 * px + 1] = c.y; out_rgb[3 * px + 2] = c.z;
                if (prim_id) prim_id[px] = prim;
                if (T && fragile) fragile[px] = pr->fragile;
            }
        }
    }
}

inline uint8_t quantize(double f) {                             // framebuffer.rs:80-82
    return (uint8_t)(255. * std::fmin(std::fmax(f, 0.), 1.));
}

}  // namespace

extern "C" {

OrcScene* orc_scene_new(void) { return new OrcScene(); }
void orc_scene_free(OrcScene* s) { delete s; }
void orc_scene_set_camera(OrcScene* s, const double xyz[3]) { s->camera = from3(xyz); }
void orc_scene_offset_camera(OrcScene* s, const double xyz[3]) { s->camera = s->camera + from3(xyz); }

void orc_reflectance_default(OrcReflectance* out) {
    Reflectance r;
    out->diffusion = r.diffusion;
    to3(r.diffuse_color, out->diffuse_color);
    out->specular = r.specular;
    out->specular_exponent = r.specular_exponent;
    out->is_glass_like = r.is_glass_like;
    out->reflection = r.reflection;
    out->refractive_index = r.refractive_index;
}

int orc_scene_add_sphere(OrcScene* s, const double center[3], double radius, const OrcReflectance* r) {
    auto sp = std::make_unique<Sphere>();                       // sphere.rs:13-25
    sp->center = from3(center);
    sp->radius_square = radius * radius;
    sp->reflectance = from_c(r);
    s->push(std::move(sp));
    return (int)s->shapes.size() - 1;
}

int orc_scene_add_polygon(OrcScene* s, const double* verts, int n, const OrcReflectance* r) {
    if (n < 3) return -1;                                       // polygon.rs:18
    std::vector<V3> v;
    for (int i = 0; i < n; i++) v.push_back(from3(verts + 3 * i));
    s->push(ConvexPolygon::create(std::move(v), from_c(r)));
    return (int)s->shapes.size() - 1;
}

int orc_scene_add_mesh(OrcScene* s, const double* tri, int n_tri, const double offset[3]) {
    auto o = make_obj("mesh", tri, n_tri);
    if (offset) o->offset(from3(offset));
    s->push(std::move(o));
    return (int)s->shapes.size() - 1;
}

int orc_scene_add_obj_file(OrcScene* s, const char* path, const double offset[3]) {
    std::vector<ObjModel> models;
    if (!load_obj(path, models)) return -1;
    for (auto& m : models) {
        std::vector<double> tri(m.tri_pos.begin(), m.tri_pos.end());   // obj.rs:102-106: f32 widened to f64
        auto o = make_obj(m.name, tri.data(), (int)(tri.size() / 9));
        if (offset) o->offset(from3(offset));                   // main.rs:278-288
        s->push(std::move(o));
    }
    return (int)models.size();
}

void orc_scene_add_light(OrcScene* s, const double pos[3], const double color[3], double intensity) {
    s->lights.push_back(create_light(from3(pos), from3(color), intensity));
}

// ---- EXTENSION MODE (no reference counterpart: the reference has one scalar refractive index per
// material, shapes.rs:21-32, and loads OBJ meshes opaque, obj.rs:125-138).  BASELINE.json configs[3] asks for a
// refractive mesh with per-channel R/G/B indices; it is traced as three passes of the UNCHANGED restatement, one scalar
// index each, and channel c of the frame is channel c of pass c.  These two setters only edit materials between passes.
int orc_scene_make_glass(OrcScene* s, int shape, double reflection, double refractive_index, double diffusion) {
    if (shape < 0 || shape >= (int)s->shapes.size()) return -1;
    std::vector<Reflectance*> m;
    s->shapes[shape]->materials(m);
    for (Reflectance* r : m) {
        r->is_glass_like = true;
        r->reflection = reflection;
        r->refractive_index = refractive_index;
        r->diffusion = diffusion;
    }
    return (int)m.size();
}
int orc_scene_set_glass_index(OrcScene* s, double refractive_index) {
    int n = 0;
    for (auto& shape : s->shapes) {
        std::vector<Reflectance*> m;
        shape->materials(m);
        for (Reflectance* r : m)
            if (r->is_glass_like) { r->refractive_index = refractive_index; n++; }
    }
    return n;
}

int orc_scene_num_shapes(const OrcScene* s) { return (int)s->shapes.size(); }
int orc_scene_num_prims(const OrcScene* s) { return s->n_prims; }

int orc_scene_obj_triangles(const OrcScene* s, int shape, double* out, int max_tris) {
    if (shape < 0 || shape >= (int)s->shapes.size() || !s->shapes[shape]->is_obj()) return 0;
    const Obj* o = static_cast<const Obj*>(s->shapes[shape].get());
    int n = (int)o->triangles.size();
    if (out)
        for (int t = 0; t < n && t < max_tris; t++)
            for (int k = 0; k < 3; k++) to3(o->triangles[t].v[k], out + 9 * t + 3 * k);
    return n;
}

// scene.rs:28-211 -- the demo scene.  The reference mutates ONE Reflectance value between shapes,
// so fields carry over from one shape to the next; the same sequence of mutations is restated here.
OrcScene* orc_scene_create_default(void) {
    OrcScene* s = new OrcScene();
    OrcReflectance r;
    orc_reflectance_default(&r);
    auto set3 = [](double* d, double x, double y, double z) { d[0] = x; d[1] = y; d[2] = z; };
    auto sphere = [](const double* c, double radius, const OrcReflectance& r) {
        auto sp = std::make_unique<Sphere>();
        sp->center = from3(c);
        sp->radius_square = radius * radius;
        sp->reflectance = from_c(&r);
        return sp;
    };
    // red sphere, scene.rs:31-47
    set3(r.diffuse_color, 0.8, 0., 0.);
    r.specular_exponent = 100.;
    const double c_red[3] = {-5., 0., -16.};
    auto sphere_red = sphere(c_red, 4., r);
    // triangle polygon, scene.rs:50-75
    set3(r.diffuse_color, 0.6, 0., 0.7);
    auto triangle = ConvexPolygon::create({{7., -4., -8.}, {15., 0., -9.}, {6., 3., -8.}}, from_c(&r));
    // floor quad, scene.rs:78-113
    r.diffusion = 1.0;
    r.specular = 1.;
    r.is_glass_like = 1;
    r.refractive_index = 1.5;
    r.reflection = 0.5;
    set3(r.diffuse_color, 0.3, 0.9, 0.9);
    auto square = ConvexPolygon::create({{20., -3., -50.}, {-20., -3., -50.}, {-15., -6., -3.}, {15., -6., -3.}}, from_c(&r));
    // blue sphere, scene.rs:116-136
    r.specular = 1.0;
    r.diffusion = 0.1;
    set3(r.diffuse_color, 0., 0., 0.2);
    r.is_glass_like = 1;
    r.refractive_index = 1.5;
    r.reflection = 0.2;
    const double c_blue[3] = {-0.5, -1.5, -5.};
    auto sphere_blue = sphere(c_blue, 2., r);
    // green sphere, scene.rs:139-159
    r.diffusion = 1.;
    r.reflection = 1.;
    r.is_glass_like = 0;
    r.specular = 0.8;
    set3(r.diffuse_color, 0., 1., 0.);
    const double c_green[3] = {6., -0.5, -18.};
    auto sphere_green = sphere(c_green, 3., r);
    // white sphere, scene.rs:162-175
    set3(r.diffuse_color, 0.9, 0.9, 0.9);
    const double c_white[3] = {-10., 6., -14.};
    auto sphere_white = sphere(c_white, 4., r);
    // lights, scene.rs:178-198
    s->lights.push_back(create_light({0., 0., 0.}, {1., 1., 1.}, 1.));
    s->lights.push_back(create_light({20., 20., 20.}, {1., 0.5, 0.5}, 0.8));
    // shape order, scene.rs:201-208
    s->push(std::move(sphere_blue));
    s->push(std::move(sphere_green));
    s->push(std::move(sphere_red));
    s->push(std::move(sphere_white));
    s->push(std::move(triangle));
    s->push(std::move(square));
    return s;
}

int orc_render(const OrcScene* sc, int W, int H, double fov, int max_depth, int n_threads,
               int patch_row_begin, int patch_row_end, int patch_stride,
               double* out_rgb, int32_t* prim_id, uint8_t* fragile, OrcCounters* counters) {
    if (W <= 0 || H <= 0 || W % 32 != 0) return -1;             // renderer.rs:107 would index out of bounds
    const int n_height = H / 32, n_width = W / 32;              // renderer.rs:53-55
    if (patch_row_end < 0 || patch_row_end > n_height) patch_row_end = n_height;
    if (patch_row_begin < 0) patch_row_begin = 0;
    if (patch_stride < 1) patch_stride = 1;
    std::vector<int> patches;
    for (int p = patch_row_begin * n_width; p < patch_row_end * n_width; p += patch_stride) patches.push_back(p);
    if (n_threads < 1) n_threads = 1;
    const bool track = prim_id || fragile || counters;
    double scale = 1.;
    for (auto& s : sc->shapes) scale = std::fmax(scale, s->max_abs());
    std::atomic<int> next{0};
    std::vector<Probe> probes(n_threads);
    for (auto& p : probes) p.scale = scale;
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) {
        Probe* pr = &probes[t];
        if (track || fragile)
            pool.emplace_back([&, pr] { render_patches<true>(*sc, W, H, fov, max_depth, patches, next, out_rgb, prim_id, fragile, pr); });
        else
            pool.emplace_back([&] { render_patches<false>(*sc, W, H, fov, max_depth, patches, next, out_rgb, nullptr, nullptr, nullptr); });
    }
    for (auto& th : pool) th.join();
    if (counters) {
        std::memset(counters, 0, sizeof(*counters));
        for (auto& p : probes) add_counters(*counters, p.c);
    }
    g_last_degenerate_hits = 0;
    for (auto& p : probes) g_last_degenerate_hits += p.degenerate_hits;
    return 0;
}

double orc_normalize(double* rgb, int W, int H) {               // framebuffer.rs:58-77
    double mx = 0., my = 0., mz = 0.;
    size_t n = (size_t)W * H;
    for (size_t i = 0; i < n; i++) {
        mx = std::fmax(mx, rgb[3 * i]);
        my = std::fmax(my, rgb[3 * i + 1]);
        mz = std::fmax(mz, rgb[3 * i + 2]);
    }
    double max_val = std::fmax(std::fmax(mx, my), mz);
    if (max_val > 0.) {
        double inv = 1. / max_val;                              // framebuffer.rs:73: scale(1. / max_val)
        for (size_t i = 0; i < 3 * n; i++) rgb[i] *= inv;
    }
    return max_val;
}

void orc_to_vec(const double* rgb, int W, int H, uint8_t* out) {   // framebuffer.rs:40-55
    size_t n = (size_t)W * H * 3;
    for (size_t i = 0; i < n; i++) out[i] = quantize(rgb[i]);
}

int orc_write_ppm(const char* path, const double* rgb, int W, int H) {   // framebuffer.rs:26-38
    FILE* f = std::fopen(path, "wb");
    if (!f) return -1;
    std::fprintf(f, "P6\n%d %d\n255\n", W, H);
    std::vector<uint8_t> buf((size_t)W * H * 3);
    orc_to_vec(rgb, W, H, buf.data());
    std::fwrite(buf.data(), 1, buf.size(), f);
    std::fclose(f);
    return 0;
}

void orc_vec_normalized(const double v[3], double out[3]) { to3(normalized(from3(v)), out); }
void orc_vec_normalized_l0(const double v[3], double out[3]) { to3(normalized_l0(from3(v)), out); }
double orc_vec_dot(const double a[3], const double b[3]) { return dot(from3(a), from3(b)); }
void orc_vec_cross(const double a[3], const double b[3], double out[3]) { to3(cross(from3(a), from3(b)), out); }
void orc_vec_scaled(const double a[3], double s, double out[3]) { to3(scaled(from3(a), s), out); }
void orc_reflect(const double i[3], const double n[3], double out[3]) { to3(reflect(from3(i), from3(n)), out); }

int orc_reflect_ray(const double incident[3], const double point[3], const double normal[3], double ri,
                    double out_orig[3], double out_dir[3]) {
    Hit h;
    h.point = from3(point);
    h.normal = from3(normal);
    V3 ro, rd;
    if (!reflect_ray<false>(from3(incident), h, ri, ro, rd, nullptr)) return 0;
    to3(ro, out_orig);
    to3(rd, out_dir);
    return 1;
}

int orc_refract_ray(const double incident[3], const double point[3], const double normal[3], double ri,
                    double out_orig[3], double out_dir[3]) {
    Hit h;
    h.point = from3(point);
    h.normal = from3(normal);
    V3 ro, rd;
    if (!refract_ray<false>(from3(incident), h, ri, ro, rd, nullptr)) return 0;
    to3(ro, out_orig);
    to3(rd, out_dir);
    return 1;
}

int orc_triangle_intersect(const double verts[9], const double orig[3], const double dir[3],
                           double out_point[3], double out_normal[3]) {
    Triangle t = Triangle::create(from3(verts), from3(verts + 3), from3(verts + 6));
    V3 p;
    to3(t.normal, out_normal);
    if (!t.isect<false>(from3(orig), from3(dir), p, nullptr)) return 0;
    to3(p, out_point);
    return 1;
}

int orc_sphere_intersect(const double center[3], double radius, const double orig[3], const double dir[3],
                         double out_point[3], double out_normal[3]) {
    Sphere s;
    s.center = from3(center);
    s.radius_square = radius * radius;
    Hit h;
    if (!s.isect<false>(from3(orig), from3(dir), h, nullptr)) return 0;
    to3(h.point, out_point);
    to3(h.normal, out_normal);
    return 1;
}

uint64_t orc_last_degenerate_hits(void) { return g_last_degenerate_hits; }

int orc_hardware_threads(void) {
    unsigned n = std::thread::hardware_concurrency();
    return n ? (int)n : 1;
}

}  // extern "C"
