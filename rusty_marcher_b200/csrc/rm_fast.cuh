// rm_fast.cuh -- the FP32 production path of the render kernel.
//
// Same predicates as rm_trace.cuh (and therefore as the reference), arranged for the B200's FP32
// pipes instead of for a literal restatement:
//
//  * Triangles (every OBJ triangle and every 3-vertex polygon) are fixed 64-byte records, four
//    128-bit shared-memory loads each, fully unrolled -- no per-edge loop, no index arithmetic.
//  * PRIMARY rays all start at the camera, so for them the whole ray/triangle test collapses to
//    four affine functions of the UN-normalised pixel direction D = (X, Y, -1):
//        dpD = n.D,   s_i = G_i.D   with  G_i = e_i(cam)*n + K*(A_i, B_i, 0),  K = (C - cam).n
//    (e_i(p) is the reference's edge term ((v_i-p) x (v_i+1-p)).z, affine in p; at the plane point
//    p = cam + t*D, t = K/dpD, one gets e_i(p)*dpD = G_i.D).  Hit <=> t >= 0, |d.n| >= 1e-6 and
//    e_i > 0 for i = 0..2 <=> (after folding sign(K) into the record) dpD > thr*|D| and s_i > 0.
//    Eight FFMAs, two FMNMX and two FSETP per triangle; the division happens only for actual hits.
//    The records are rebuilt per frame in FP64 by prepare_raster() (one thread per triangle).
//  * Secondary and shadow rays use the general form with the edge terms evaluated relative to
//    vertex 0 (see plane_intersect<float> in rm_trace.cuh for why not at the hit point itself).
//  * Spheres and n-gons (n > 3) reuse the routines of rm_trace.cuh.
#pragma once

#include "rm_trace.cuh"

namespace rm {

constexpr int kTriSrcDoubles = 16;   // n[3], dn, v0x, v0y, A0, B0, A1, B1, C1, A2, B2, thr, id, pad

#if defined(__CUDA_ARCH__)
RM_HD float fast_div(float a, float b) { return __fdividef(a, b); }
#else
RM_HD float fast_div(float a, float b) { return a / b; }
#endif

// Camera-specialised record of one triangle (4 x R4<float>), from its FP64 source record.
//   r0 = {n'x, n'y, -n'z, |K|}   r1 = {G0x, G0y, -G0z, thr}   r2 = {G1x, G1y, -G1z, id}   r3 = {G2x, G2y, -G2z, 0}
// with everything multiplied by sign(K) so that a hit needs dpD > 0 and s_i > 0.
RM_HD void prepare_raster(const double* __restrict__ src, const double cam[3], R4<float>* __restrict__ out) {
    const double nx = src[0], ny = src[1], nz = src[2], dn = src[3];
    const double wx = cam[0] - src[4], wy = cam[1] - src[5];
    const double K = dn - (cam[0] * nx + cam[1] * ny + cam[2] * nz);
    const double A[3] = {src[6], src[8], src[11]}, B[3] = {src[7], src[9], src[12]}, Cc[3] = {0., src[10], 0.};
    const double s = (K < 0.) ? -1. : 1.;
    out[0] = {(float)(s * nx), (float)(s * ny), (float)(-s * nz), (float)fabs(K)};
    float last[3] = {(float)src[13], 0.f, 0.f};
    int id = (int)src[14];
#if defined(__CUDA_ARCH__)
    last[1] = __int_as_float(id);
#else
    memcpy(&last[1], &id, 4);
#endif
    for (int i = 0; i < 3; i++) {
        const double E = A[i] * wx + B[i] * wy + Cc[i];
        out[1 + i] = {(float)(s * (E * nx + K * A[i])), (float)(s * (E * ny + K * B[i])), (float)(-s * (E * nz)), last[i]};
    }
}

struct FastView {
    // spheres
    const R4<float>* sph;
    const int* sph_id;
    int n_sph;
    // triangles: tri_g = scene records {n,dn | v0x,v0y,A0,B0 | A1,B1,C1,A2 | B2,thr,id,0}, tri_r = raster records
    const R4<float>* tri_g;
    const R4<float>* tri_r;
    int n_tri;
    // n-gons: indices into the generic plane arrays
    const int* poly_slot;
    int n_poly;
    const R4<float>* pln_n;
    const R4<float>* pln_c;
    const I2* pln_v;
    const int* pln_id;
    const R4<float>* vert;
    // materials, lights
    const R4<float>* mat_a;
    const R4<float>* mat_b;
    const int* mat_f;
    const R4<float>* lgt_p;
    const R4<float>* lgt_c;
    int n_lgt;
    // primary hit of this thread's pixel
    bool prim_got;
    HitRec<float> prim_hit;

    static RM_HD int as_int(float f) {
#if defined(__CUDA_ARCH__)
        return __float_as_int(f);
#else
        int i;
        memcpy(&i, &f, 4);
        return i;
#endif
    }

    // general ray against triangle j
    RM_HD bool tri_hit(int j, const Vec3<float> o, const Vec3<float> d, float& t_out) const {
        const R4<float> a = tri_g[4 * j], l = tri_g[4 * j + 3];
        const Vec3<float> n = xyz(a);
        const float dp = dot(d, n);
        if (!(fabsf(dp) > l.y)) return false;                  // triangle.rs:57
        const float t = fast_div(a.w - dot(o, n), dp);         // triangle.rs:62
        if (t < 0.f) return false;                             // triangle.rs:65
        const R4<float> b = tri_g[4 * j + 1], c = tri_g[4 * j + 2];
        const float qx = fmaf(t, d.x, o.x - b.x), qy = fmaf(t, d.y, o.y - b.y);   // hit point relative to vertex 0
        const float e0 = fmaf(b.z, qx, b.w * qy);
        const float e1 = fmaf(c.x, qx, fmaf(c.y, qy, c.z));
        const float e2 = fmaf(c.w, qx, l.x * qy);
        if (!(fminf(fminf(e0, e1), e2) > 0.f)) return false;   // triangle.rs:72-76
        t_out = t;
        return true;
    }

    RM_HD static void keep(HitRec<float>& best, bool& hit, float t, int slot, int id) {
        if (!hit || t < best.dist || (t == best.dist && id < best.id)) {
            best.dist = t;
            best.slot = slot;
            best.id = id;
            hit = true;
        }
    }

    // closest hit of the primary ray of pixel direction D = (X, Y, -1), |D| = lenD, d = D / lenD
    RM_HD void primary(const Vec3<float> cam, const float X, const float Y, const float lenD, const Vec3<float> d) {
        bool hit = false;
        HitRec<float> best;
        Counters<false> st;
        for (int i = 0; i < n_sph; i++) {
            Cand<float> c;
            if (sphere_intersect<false>(sph[i], cam, d, c, st)) keep(best, hit, c.key, i, sph_id[i]);
        }
        for (int j = 0; j < n_tri; j++) {
            const R4<float> r0 = tri_r[4 * j], r1 = tri_r[4 * j + 1], r2 = tri_r[4 * j + 2], r3 = tri_r[4 * j + 3];
            const float dpD = fmaf(r0.x, X, fmaf(r0.y, Y, r0.z));
            const float s0 = fmaf(r1.x, X, fmaf(r1.y, Y, r1.z));
            const float s1 = fmaf(r2.x, X, fmaf(r2.y, Y, r2.z));
            const float s2 = fmaf(r3.x, X, fmaf(r3.y, Y, r3.z));
            if (fminf(fminf(s0, s1), s2) > 0.f && dpD > r1.w * lenD)
                keep(best, hit, fast_div(r0.w * lenD, dpD), n_sph + j, as_int(r2.w));
        }
        for (int k = 0; k < n_poly; k++) {
            const int i = poly_slot[k];
            Cand<float> c;
            if (plane_intersect<false>(pln_n[i], pln_c[i], pln_v[i], vert, cam, d, c, st))
                keep(best, hit, c.key, n_sph + n_tri + i, pln_id[i]);
        }
        prim_got = hit;
        prim_hit = best;
    }

    template <bool S> RM_HD bool closest(const Vec3<float> o, const Vec3<float> d, int level, HitRec<float>& h, Counters<S>& st) const {
        if (level == 1) {
            h = prim_hit;
            return prim_got;
        }
        bool hit = false;
        for (int i = 0; i < n_sph; i++) {
            Cand<float> c;
            if (sphere_intersect<S>(sph[i], o, d, c, st)) keep(h, hit, c.key, i, sph_id[i]);
        }
        for (int j = 0; j < n_tri; j++) {
            float t;
            if (tri_hit(j, o, d, t)) keep(h, hit, t, n_sph + j, as_int(tri_g[4 * j + 3].z));
        }
        for (int k = 0; k < n_poly; k++) {
            const int i = poly_slot[k];
            Cand<float> c;
            if (plane_intersect<S>(pln_n[i], pln_c[i], pln_v[i], vert, o, d, c, st))
                keep(h, hit, c.key, n_sph + n_tri + i, pln_id[i]);
        }
        return hit;
    }

    template <bool S> RM_HD bool anyhit(const Vec3<float> o, const Vec3<float> d, Counters<S>& st) const {
        Cand<float> c;
        for (int i = 0; i < n_sph; i++)
            if (sphere_intersect<S>(sph[i], o, d, c, st)) return true;
        float t;
        for (int j = 0; j < n_tri; j++)
            if (tri_hit(j, o, d, t)) return true;
        for (int k = 0; k < n_poly; k++) {
            const int i = poly_slot[k];
            if (plane_intersect<S>(pln_n[i], pln_c[i], pln_v[i], vert, o, d, c, st)) return true;
        }
        return false;
    }

    RM_HD void surface(HitRec<float>& h, const Vec3<float> o, const Vec3<float> d, Vec3<float>& normal) const {
        if (h.slot < n_sph) {
            sphere_point_normal(sph[h.slot], o, d, h.p, h.p, normal);
        } else {
            h.p = axpy(o, d, h.dist);
            normal = (h.slot < n_sph + n_tri) ? xyz(tri_g[4 * (h.slot - n_sph)]) : xyz(pln_n[h.slot - n_sph - n_tri]);
        }
    }
};

// One pixel of the FP32 production kernel.
RM_HD Vec3<float> fast_pixel(FastView& fv, const FrameParams<float>& fp, int x, int y, int& primary_id) {
    const float X = (float(x) - fp.half_w) * fp.sx, Y = (float(y) - fp.half_h) * fp.sy;
    const float len2 = fmaf(X, X, fmaf(Y, Y, 1.f));
    const float lenD = sqrtf(len2);
    const float inv = 1.f / lenD;                              // geometry.rs:104-109: scale(1/norm)
    const Vec3<float> d = {X * inv, Y * inv, -inv};
    fv.primary(fp.camera, X, Y, lenD, d);
    Counters<false> st;
    return cast_ray<float, false, FastView>(fv, fp.camera, d, fp.background, fp.max_depth, primary_id, st);
}

}  // namespace rm
