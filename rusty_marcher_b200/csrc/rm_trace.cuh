// rm_trace.cuh -- per-ray work of the render hot path: closest-hit / any-hit queries over the
// flattened scene, direct lighting with shadow rays, the optics of glass-like materials and the
// depth-capped reflect/refract recursion, unrolled into an iterative per-thread stack.
//
// Each routine cites the reference lines it replaces.  The code is templated on the scalar type
// (see rm_math.cuh) and on a compile-time `S` flag that switches the event counters on (used
// once per workload outside the timed region to obtain the algorithmic work).
#pragma once

#include "rm_math.cuh"

namespace rm {

constexpr int kMaxDepth = 8;   // frames of the iterative stack (reference: 3, renderer.rs:262)

// indices into Counters::c -- same order as RmStats / OrcCounters
enum CounterId {
    C_PIXELS, C_CLOSEST, C_ANYHIT, C_SPH_TEST, C_SPH_DISC, C_SPH_HIT, C_PLN_TEST, C_PLN_DIST, C_PLN_POINT,
    C_EDGE, C_CAND, C_HITS, C_LIGHT_EVAL, C_LIT, C_GLASS, C_REFL, C_REFR, C_COUNT
};

template <bool S> struct Counters {
    RM_HD void clear() {}
    RM_HD void add(int, unsigned = 1) {}
};
template <> struct Counters<true> {
    unsigned c[C_COUNT];
    RM_HD void clear() { for (int i = 0; i < C_COUNT; i++) c[i] = 0; }
    RM_HD void add(int id, unsigned n = 1) { c[id] += n; }
};

// The scene as the kernels see it (pointers into shared or global memory).
//   sph[i]    = {cx, cy, cz, r^2}                       sphere.rs:6-11
//   pln_n[i]  = {nx, ny, nz, thr}   hit needs |d.n| > thr: thr = pred(1e-6) for triangles
//                                   (triangle.rs:57), 0 for polygons (polygon.rs:66)
//   pln_c[i]  = f64: {Cx, Cy, Cz, 0} plane point (triangle.center / polygon.plane_point)
//               f32: {n.C, v0.x, v0.y, 0}
//   pln_v[i]  = {first, count} into vert[]: f64 {x, y} per vertex (only x,y are ever read:
//               triangle.rs:13-15 uses the z component of the cross product); f32 {A, B, C, 0} per edge
//   *_id[i]   = flattened primitive id (index in scene order, Obj expanded per triangle)
//   mat_a[id] = {kd.r, kd.g, kd.b, diffusion}; mat_b[id] = {specular, exponent, reflection, index}
//   mat_f[id] = bit0 is_glass_like                      shapes.rs:21-32
//   lgt_p[l]  = {px, py, pz, intensity}; lgt_c[l] = {r, g, b, 0}      lights.rs:4-8
// `order*` (instrumented variant only) lists the resident primitives in scene order so that the
// counters follow the reference's traversal: order[k] = slot (sphere i -> i, plane i -> n_sph+i),
// order_shape[k] = index of the owning shape, bit 30 set when the shape is an Obj.
template <typename R> struct VertOf { typedef R2<R> type; };
template <> struct VertOf<float> { typedef R4<float> type; };
template <typename R> using VertT = typename VertOf<R>::type;

template <typename R> struct HitRec;
template <typename R> struct SceneView;
template <typename R, bool S>
RM_HD bool find_closest_intersect(const SceneView<R>& sc, const Vec3<R> o, const Vec3<R> d, HitRec<R>& best, Counters<S>& st);
template <typename R, bool S>
RM_HD bool intersect_shape_set(const SceneView<R>& sc, const Vec3<R> o, const Vec3<R> d, Counters<S>& st);
template <typename R> RM_HD void surface_of(const SceneView<R>& sc, HitRec<R>& h, const Vec3<R> o, const Vec3<R> d, Vec3<R>& normal);
template <typename R, bool S, typename SC>
RM_HD Vec3<R> direct_lighting(const SC& sc, const Vec3<R> origin, const Vec3<R> point, const Vec3<R> normal, const R4<R> ma,
                              const R4<R> mb, Counters<S>& st);

template <typename R> struct SceneView {
    // the query interface cast_ray / direct_lighting are written against (rm_fast.cuh has a second implementation)
    template <bool S> RM_HD bool closest(const Vec3<R> o, const Vec3<R> d, int /*level*/, HitRec<R>& h, Counters<S>& st) const {
        return find_closest_intersect<R, S>(*this, o, d, h, st);
    }
    template <bool S> RM_HD bool anyhit(const Vec3<R> o, const Vec3<R> d, Counters<S>& st) const {
        return intersect_shape_set<R, S>(*this, o, d, st);
    }
    RM_HD void surface(HitRec<R>& h, const Vec3<R> o, const Vec3<R> d, Vec3<R>& normal) const { surface_of(*this, h, o, d, normal); }
    template <bool S> RM_HD Vec3<R> direct(const Vec3<R> origin, const Vec3<R> /*dir*/, const Vec3<R> point, const Vec3<R> normal,
                                           const R4<R> ma, const R4<R> mb, Counters<S>& st) const {
        return direct_lighting<R, S, SceneView<R>>(*this, origin, point, normal, ma, mb, st);
    }

    const R4<R>* sph;
    const int* sph_id;
    int n_sph;
    const R4<R>* pln_n;
    const R4<R>* pln_c;
    const I2* pln_v;
    const int* pln_id;
    int n_pln;
    const VertT<R>* vert;
    const R4<R>* mat_a;
    const R4<R>* mat_b;
    const int* mat_f;
    const R4<R>* lgt_p;
    const R4<R>* lgt_c;
    int n_lgt;
    const int* order;
    const int* order_shape;
    int n_order;
};

template <typename R> struct HitRec {
    Vec3<R> p;
    R dist;     // f64: |p - orig|^2 (shapes.rs:128); f32: the ray parameter t (same ordering, d is unit)
    int slot;   // sphere i -> i, plane i -> n_sph + i
    int id;     // flattened primitive id
};

// One candidate hit.  f64 carries the point like the reference; f32 only the ray parameter, the point
// is formed once for the winner.
template <typename R> struct Cand {
    Vec3<R> p;
    R key;
};

// ---------------------------------------------------------------- sphere.rs:27-61
template <bool S>
RM_HD bool sphere_intersect(const R4<double> s, const Vec3<double> o, const Vec3<double> d, Cand<double>& c, Counters<S>& st) {
    Vec3<double> line = {s.x - o.x, s.y - o.y, s.z - o.z};
    double tca = dot(line, d);
    double d2 = dot(line, line) - tca * tca;
    st.add(C_SPH_TEST);
    if (d2 > s.w) return false;
    st.add(C_SPH_DISC);
    double thc = sqrt(s.w - d2);
    double t0 = tca - thc;
    double t1 = tca + thc;
    if (t0 < 0.) t0 = t1;
    if (t0 < 0.) return false;
    st.add(C_SPH_HIT);
    c.p = axpy(o, d, t0);
    c.key = squared_norm(c.p - o);                             // shapes.rs:128
    return true;
}
// f32: the reference's d2 = line.line - tca^2 cancels catastrophically in single precision for a small
// far sphere (error 2^-24*|line|^2 against r^2), so the squared distance of the centre to the ray is
// taken from the perpendicular component itself: perp = line - tca*d, d2 = perp.perp.
template <bool S>
RM_HD bool sphere_intersect(const R4<float> s, const Vec3<float> o, const Vec3<float> d, Cand<float>& c, Counters<S>& st) {
    Vec3<float> line = {s.x - o.x, s.y - o.y, s.z - o.z};
    float tca = dot(line, d);
    Vec3<float> perp = axmy(line, d, tca);
    float d2 = dot(perp, perp);
    st.add(C_SPH_TEST);
    if (d2 > s.w) return false;
    st.add(C_SPH_DISC);
    float thc = sqrtf(s.w - d2);
    float t0 = tca - thc;
    if (t0 < 0.f) t0 = tca + thc;
    if (t0 < 0.f) return false;
    st.add(C_SPH_HIT);
    c.key = t0;
    return true;
}

// Hit point and unit normal of the winning sphere hit (sphere.rs:54-60).
RM_HD void sphere_point_normal(const R4<double> s, const Vec3<double>, const Vec3<double>, const Vec3<double> p_in,
                               Vec3<double>& p, Vec3<double>& n) {
    p = p_in;
    n = normalized(p_in - xyz(s));
}
// f32: p - c = (t - tca)*d - perp = -+thc*d - perp, formed from the two short vectors instead of the
// difference of two long ones (o + t*d) - c.
RM_HD void sphere_point_normal(const R4<float> s, const Vec3<float> o, const Vec3<float> d, const Vec3<float>,
                               Vec3<float>& p, Vec3<float>& n) {
    Vec3<float> line = {s.x - o.x, s.y - o.y, s.z - o.z};
    float tca = dot(line, d);
    Vec3<float> perp = axmy(line, d, tca);
    float thc = sqrtf(fmaxf(s.w - dot(perp, perp), 0.f));
    float w = (tca - thc < 0.f) ? thc : -thc;                  // same root as sphere_intersect chose
    Vec3<float> pc = {fmaf(w, d.x, -perp.x), fmaf(w, d.y, -perp.y), fmaf(w, d.z, -perp.z)};
    p = xyz(s) + pc;
    n = normalized(pc);
}

// ---------------------------------------------------------------- triangle.rs:49-83, polygon.rs:60-98
// f64: the reference's arithmetic, operation for operation.
//   pln_n = {n, thr}, pln_c = {plane point, 0}, vert[] = {x, y} of each vertex
template <bool S>
RM_HD bool plane_intersect(const R4<double> n4, const R4<double> c4, const I2 vr, const R2<double>* __restrict__ vert,
                           const Vec3<double> o, const Vec3<double> d, Cand<double>& c, Counters<S>& st) {
    Vec3<double> n = xyz(n4);
    double dp = dot(d, n);
    st.add(C_PLN_TEST);
    if (!(fabs(dp) > n4.w)) return false;                      // parallel (triangle.rs:57 / polygon.rs:66)
    Vec3<double> co = {c4.x - o.x, c4.y - o.y, c4.z - o.z};
    double dist = dot(co, n) / dp;                             // triangle.rs:62
    st.add(C_PLN_DIST);
    if (dist < 0.) return false;                               // going away
    Vec3<double> q = axpy(o, d, dist);
    st.add(C_PLN_POINT);
    // inside(): ((v_i - q) x (v_i+1 - q)).z > 0 for every edge, early out (triangle.rs:72-76)
    R2<double> v1 = vert[vr.x];
    for (int i = 0; i < vr.y; i++) {
        R2<double> v2 = vert[vr.x + ((i + 1 == vr.y) ? 0 : i + 1)];
        double ux = v1.x - q.x, uy = v1.y - q.y, vx = v2.x - q.x, vy = v2.y - q.y;
        double cz = ux * vy - uy * vx;
        st.add(C_EDGE);
        if (!(cz > 0.)) return false;
        v1 = v2;
    }
    c.p = q;
    c.key = squared_norm(q - o);                               // obj.rs:197 / shapes.rs:128
    return true;
}
// f32: same predicates in a form that is robust in single precision.  The reference evaluates the
// edge term at the hit point p = o + t*d, which loses all significance in FP32 when a grazing ray
// meets the plane far away (|p|^2 * 2^-24 exceeds the triangle's area).  The term is affine in p,
//   e_i(p) = A_i*(p.x - v0.x) + B_i*(p.y - v0.y) + C_i,   A_i = v_i.y - v_i+1.y,  B_i = v_i+1.x - v_i.x,
//   C_i = ((v_i - v0) x (v_i+1 - v0)).z,
// so it is evaluated as e_i(o) + t*(A_i*d.x + B_i*d.y): two numbers of the size of the result.
//   pln_n = {n, thr}, pln_c = {n.C, v0.x, v0.y, 0} (f64 on the host), vert[] = {A_i, B_i, C_i, 0}
template <bool S>
RM_HD bool plane_intersect(const R4<float> n4, const R4<float> k4, const I2 vr, const R4<float>* __restrict__ edge,
                           const Vec3<float> o, const Vec3<float> d, Cand<float>& c, Counters<S>& st) {
    Vec3<float> n = xyz(n4);
    float dp = dot(d, n);
    st.add(C_PLN_TEST);
    if (!(fabsf(dp) > n4.w)) return false;
    float t = (k4.x - dot(o, n)) / dp;                         // ((C - o).n) / (d.n)
    st.add(C_PLN_DIST);
    if (t < 0.f) return false;
    st.add(C_PLN_POINT);
    const float wx = o.x - k4.y, wy = o.y - k4.z;
    for (int i = 0; i < vr.y; i++) {
        const R4<float> e = edge[vr.x + i];
        float g = fmaf(e.y, d.y, e.x * d.x);
        float e0 = fmaf(e.x, wx, fmaf(e.y, wy, e.z));
        st.add(C_EDGE);
        if (!(fmaf(t, g, e0) > 0.f)) return false;
    }
    c.key = t;
    return true;
}

template <typename R>
RM_HD void consider(HitRec<R>& best, bool& hit, const Cand<R>& c, int slot, int id) {
    // strict '<' with the first primitive in scene order winning ties (shapes.rs:130, obj.rs:198);
    // the id comparison keeps that rule although spheres are traversed before planes here
    if (!hit || c.key < best.dist || (c.key == best.dist && id < best.id)) {
        best.p = c.p;
        best.dist = c.key;
        best.slot = slot;
        best.id = id;
        hit = true;
    }
}

// ---------------------------------------------------------------- shapes.rs:110-143 + obj.rs:186-216
template <typename R, bool S>
RM_HD bool find_closest_intersect(const SceneView<R>& sc, const Vec3<R> o, const Vec3<R> d, HitRec<R>& best,
                                  Counters<S>& st) {
    bool hit = false;
    st.add(C_CLOSEST);
    if constexpr (S) {
        // instrumented: scene order, so that per-event counts equal the reference traversal
        int last_shape = -1;
        for (int k = 0; k < sc.n_order; k++) {
            int slot = sc.order[k];
            int shape = sc.order_shape[k];
            Cand<R> c;
            bool got;
            int id;
            if (slot < sc.n_sph) {
                got = sphere_intersect<S>(sc.sph[slot], o, d, c, st);
                id = sc.sph_id[slot];
            } else {
                int i = slot - sc.n_sph;
                got = plane_intersect<S>(sc.pln_n[i], sc.pln_c[i], sc.pln_v[i], sc.vert, o, d, c, st);
                id = sc.pln_id[i];
            }
            if (got) {
                // one distance per hit in the shape loop (shapes.rs:128) plus, inside an Obj, one per
                // hitting triangle (obj.rs:197)
                if (shape & (1 << 30)) st.add(C_CAND, (shape != last_shape) ? 2 : 1);
                else st.add(C_CAND);
                last_shape = shape;
                consider(best, hit, c, slot, id);
            }
        }
        return hit;
    } else {
        for (int i = 0; i < sc.n_sph; i++) {
            Cand<R> c;
            if (sphere_intersect<S>(sc.sph[i], o, d, c, st)) consider(best, hit, c, i, sc.sph_id[i]);
        }
        for (int i = 0; i < sc.n_pln; i++) {
            Cand<R> c;
            if (plane_intersect<S>(sc.pln_n[i], sc.pln_c[i], sc.pln_v[i], sc.vert, o, d, c, st))
                consider(best, hit, c, sc.n_sph + i, sc.pln_id[i]);
        }
        return hit;
    }
}

// ---------------------------------------------------------------- shapes.rs:92-108
// No maximum distance: occluders behind the light still shadow.  The reference keeps scanning the
// triangles of an Obj after its first hit (obj.rs:194-210); the boolean result is the same when we
// stop at the first hitting primitive, and that is how the algorithmic work is counted.
template <typename R, bool S>
RM_HD bool intersect_shape_set(const SceneView<R>& sc, const Vec3<R> o, const Vec3<R> d, Counters<S>& st) {
    st.add(C_ANYHIT);
    Cand<R> c;
    if constexpr (S) {
        for (int k = 0; k < sc.n_order; k++) {
            int slot = sc.order[k];
            bool got;
            if (slot < sc.n_sph) got = sphere_intersect<S>(sc.sph[slot], o, d, c, st);
            else {
                int i = slot - sc.n_sph;
                got = plane_intersect<S>(sc.pln_n[i], sc.pln_c[i], sc.pln_v[i], sc.vert, o, d, c, st);
            }
            if (got) {
                if (sc.order_shape[k] & (1 << 30)) st.add(C_CAND);   // obj.rs:197
                return true;
            }
        }
        return false;
    } else {
        for (int i = 0; i < sc.n_sph; i++)
            if (sphere_intersect<S>(sc.sph[i], o, d, c, st)) return true;
        for (int i = 0; i < sc.n_pln; i++)
            if (plane_intersect<S>(sc.pln_n[i], sc.pln_c[i], sc.pln_v[i], sc.vert, o, d, c, st)) return true;
        return false;
    }
}

// ---------------------------------------------------------------- optics.rs:4-6
template <typename R> RM_HD Vec3<R> reflect(Vec3<R> incident, Vec3<R> normal) {
    return axmy(incident, normal, R(2) * dot(incident, normal));
}

// ---------------------------------------------------------------- optics.rs:8-48
template <typename R, typename N = Exact<R>>
RM_HD bool reflect_ray(Vec3<R> incident, Vec3<R> point, Vec3<R> hit_normal, R refractive_index, Vec3<R>& ro, Vec3<R>& rd) {
    Vec3<R> normal = hit_normal;
    R c = dot(normal, incident);
    R r = (c < R(0)) ? refractive_index : N::rcp_(refractive_index);
    if (c < R(0)) { c = -c; normal = -normal; }
    R cos_theta_2 = R(1) - r * r * (R(1) - c * c);
    if (cos_theta_2 > R(0)) return false;
    rd = reflect(incident, normal);
    if (dot(rd, hit_normal) < R(0)) ro = axmy(point, hit_normal, R(1e-4));
    else ro = axpy(point, hit_normal, R(1e-4));
    return true;
}

// ---------------------------------------------------------------- optics.rs:50-89
template <typename R, typename N = Exact<R>>
RM_HD bool refract_ray(Vec3<R> incident, Vec3<R> point, Vec3<R> hit_normal, R refractive_index, Vec3<R>& ro, Vec3<R>& rd) {
    Vec3<R> normal = hit_normal;
    R c = -dot(normal, incident);
    R r = (c < R(0)) ? refractive_index : N::rcp_(refractive_index);
    if (c < R(0)) { c = -c; normal = -normal; }
    R cos_theta_2 = R(1) - r * r * (R(1) - c * c);
    if (cos_theta_2 < R(0)) return false;
    rd = N::normalized(scaled(incident, r) + scaled(normal, r * c - N::sqrt_(cos_theta_2)));
    if (dot(rd, normal) > R(0)) ro = axpy(point, normal, R(1e-4));
    else ro = axmy(point, normal, R(1e-4));
    return true;
}

// point + unit normal of the winning hit (sphere.rs:54-58; planes carry their precomputed normal)
template <typename R> RM_HD void surface_of(const SceneView<R>& sc, HitRec<R>& h, const Vec3<R> o, const Vec3<R> d, Vec3<R>& normal) {
    if (h.slot < sc.n_sph) {
        sphere_point_normal(sc.sph[h.slot], o, d, h.p, h.p, normal);
    } else {
        if (sizeof(R) == 4) h.p = axpy(o, d, h.dist);          // f32: point of the winner only
        normal = xyz(sc.pln_n[h.slot - sc.n_sph]);
    }
}

// ---------------------------------------------------------------- renderer.rs:138-193
template <typename R, bool S, typename SC>
RM_HD Vec3<R> direct_lighting(const SC& sc, const Vec3<R> origin, const Vec3<R> point, const Vec3<R> normal,
                              const R4<R> ma, const R4<R> mb, Counters<S>& st) {
    Vec3<R> acc = {R(0), R(0), R(0)};
    const Vec3<R> to_viewer = normalized(origin - point);                                         // renderer.rs:149 (the same for every light)
    for (int l = 0; l < sc.n_lgt; l++) {
        R4<R> lp = sc.lgt_p[l];
        R4<R> lc4 = sc.lgt_c[l];
        Vec3<R> lc = xyz(lc4);
        Vec3<R> light_dir = normalized(Vec3<R>{lp.x - point.x, lp.y - point.y, lp.z - point.z});   // renderer.rs:166
        R side = dot(light_dir, normal);
        st.add(C_LIGHT_EVAL);
        Vec3<R> so = (side < R(0)) ? axmy(point, normal, R(1e-3)) : axpy(point, normal, R(1e-3)); // renderer.rs:168-172
        pin(so);
        pin(light_dir);
        if (sc.template anyhit<S>(so, light_dir, st)) continue;                                   // renderer.rs:174-177
        st.add(C_LIT);
        R diffusion = Num<R>::max_(side, R(0));                                                   // renderer.rs:138-140
        Vec3<R> kd = {ma.x, ma.y, ma.z};
        acc = acc + scaled(scaled(lc * kd, diffusion), lp.w);                                     // renderer.rs:181-183
        Vec3<R> reflected = reflect(-light_dir, normal);                                          // renderer.rs:144-145
        R sf = Num<R>::max_(dot(reflected, to_viewer), R(0));                                     // renderer.rs:150
        R specular = Num<R>::pow_(sf * mb.x, mb.y);                                               // renderer.rs:186-188
        acc = acc + scaled(lc, specular);                                                         // renderer.rs:189
    }
    return scaled(acc, ma.w);                                                                     // renderer.rs:192
}

// ---------------------------------------------------------------- renderer.rs:254-309
// cast_ray's recursion (branching factor 2, depth cap max_depth) as an explicit stack.  A frame is
// pushed for a glass hit that spawned at least one secondary ray; it keeps the partial sum and the
// pending refracted ray so that the additions happen in the reference's order:
//   c = bg + direct;  c += cast(reflected) * k;  c += cast(refracted) * (1 - k).
template <typename R, bool S, typename SC>
RM_HD Vec3<R> cast_ray(const SC& sc, Vec3<R> o, Vec3<R> d, R background, int max_depth, int& primary_id,
                       Counters<S>& st) {
    struct Frame {
        Vec3<R> c, ro, rd;
        R k;
        int state;   // bit0: a refracted ray is pending, bit1: the refracted ray is the one in flight
    };
    Frame fr[kMaxDepth];
    int sp = 0;
    const Vec3<R> bg = {background, background, background};
    Vec3<R> v;
    primary_id = -1;
    for (;;) {
        const int level = sp + 1;                               // n_recursion
        if (level > max_depth) {
            v = bg;                                             // renderer.rs:262-264
        } else {
            HitRec<R> h;
            bool got = sc.template closest<S>(o, d, level, h, st);      // renderer.rs:266
            if (level == 1 && got) primary_id = h.id;
            if (!got) {
                v = (level > 1) ? bg : Vec3<R>{R(0), R(0), R(0)};       // renderer.rs:300-306
            } else {
                st.add(C_HITS);
                Vec3<R> normal;
                sc.surface(h, o, d, normal);
                const R4<R> ma = sc.mat_a[h.id];
                const R4<R> mb = sc.mat_b[h.id];
                Vec3<R> c = bg + sc.template direct<S>(o, d, h.p, normal, ma, mb, st);        // renderer.rs:272-275
                bool pushed = false;
                if (sc.mat_f[h.id] & 1) {                       // renderer.rs:277
                    st.add(C_GLASS);
                    Vec3<R> ro1, rd1, ro2, rd2;
                    bool has_refl = reflect_ray<R>(d, h.p, normal, mb.w, ro1, rd1);       // renderer.rs:203-207
                    bool has_refr = refract_ray<R>(d, h.p, normal, mb.w, ro2, rd2);       // renderer.rs:235-239
                    if (has_refl) st.add(C_REFL);
                    if (has_refr) st.add(C_REFR);
                    if (has_refl || has_refr) {
                        Frame& f = fr[sp];
                        f.c = c;
                        f.k = mb.z;
                        if (has_refl) {
                            f.state = has_refr ? 1 : 0;
                            f.ro = ro2;
                            f.rd = rd2;
                            o = ro1;
                            d = rd1;
                        } else {
                            f.state = 2;
                            o = ro2;
                            d = rd2;
                        }
                        sp++;
                        pushed = true;
                    }
                }
                if (pushed) continue;
                v = c;
            }
        }
        // return v to the callers on the stack
        for (;;) {
            if (sp == 0) return v;
            Frame& f = fr[sp - 1];
            if (f.state & 2) {
                f.c = f.c + scaled(v, R(1) - f.k);              // renderer.rs:249
                v = f.c;
                sp--;
            } else {
                f.c = f.c + scaled(v, f.k);                     // renderer.rs:219
                if (f.state & 1) {
                    o = f.ro;
                    d = f.rd;
                    f.state = 2;
                    break;
                }
                v = f.c;
                sp--;
            }
        }
    }
}

// Per-frame constants.  backproject (renderer.rs:128-135):
//   x = 2*(j/W - 0.5)*half_fov*ratio, y = -2*(i/H - 0.5)*half_fov, z = -1, then normalised.
template <typename R> struct FrameParams {
    int width, height;
    int row_begin, row_end;       // pixel-row span of this call
    int row_step;                 // pixel rows from one rendered 32-row band to the next (32 = contiguous; 32*G when the
                                  // bands of a frame are dealt round-robin to G ranks)
    int n_bands;                  // rendered 32-row bands: rows row_begin + b*row_step + [0, 32), b < n_bands
    int buf_row0;                 // pixel row stored at offset 0 of the output buffers
    R width_r, height_r, half_fov, ratio;   // f64 path: the reference's own operands
    R half_w, half_h, sx, sy;     // f32 path: x = (j - W/2)*sx, y = (i - H/2)*sy, folded on the host in f64
    Vec3<R> camera;
    R background;
    int max_depth;
    int accel;                    // RmParams.accel: scene queries through the hierarchy (FP32 production kernel)
    // the reference's own f64 operands of backproject (renderer.rs:128-135) and the camera, whatever R is: the FP32
    // production kernel traces the paths behind a glass-like primary hit with f64 ray geometry (cast_glass, rm_fast.cuh)
    double cam64[3], w64, h64, hf64, ratio64, inv_w64, inv_h64;
};

// pixel row of the l-th rendered row of this call (l in [0, 32 * n_bands))
template <typename R> RM_HD int frame_row(const FrameParams<R>& fp, int l) { return fp.row_begin + (l >> 5) * fp.row_step + (l & 31); }

template <typename R> RM_HD Vec3<R> backproject(const FrameParams<R>& fp, int j, int i) {
    Vec3<R> v = {(R(j) - fp.half_w) * fp.sx, (R(i) - fp.half_h) * fp.sy, R(-1)};
    return normalized(v);
}
template <> RM_HD Vec3<double> backproject<double>(const FrameParams<double>& fp, int j, int i) {
    Vec3<double> v = {2. * ((double)j / fp.width_r - 0.5) * fp.half_fov * fp.ratio,
                      -2. * ((double)i / fp.height_r - 0.5) * fp.half_fov, -1.};
    return normalized(v);
}

}  // namespace rm
