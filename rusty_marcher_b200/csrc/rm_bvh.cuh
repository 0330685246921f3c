// rm_bvh.cuh -- bounding-volume hierarchy over the hittable primitives of a scene.
//
// The reference carries bounding boxes it never uses (engine/src/shapes.rs:34-38,63-86, sphere.rs:18-21,
// obj.rs:88-92) and answers every scene query by brute force (shapes.rs:92-143, obj.rs:186-216).  For
// scenes of BASELINE.json configs[4]'s class (4096 spheres + 100k triangles) that is ~10^5 primitive tests
// per ray segment.  With RmParams.accel = 1 the production kernel walks this hierarchy instead.  Only the
// SET of primitives a ray is tested against changes -- a superset of the ones it can hit; every
// (ray, primitive) test is the very routine the brute-force path runs (rm_fast.cuh / rm_trace.cuh), and the
// winner is chosen by the same (distance, primitive id) order -- so the frame is bit-identical to the
// brute-force FP32 frame (tests/test_kernel_emulation.py on the host, tests/test_gpu_parity.py on the B200).
//
// Why a box test in FP32 may drop primitives at all: a primitive test that reports a hit in FP32 does so for
// a ray that passes within a few ulps of scene-scale coordinates (~2^-20 S) of the primitive; the boxes are
// grown by 2^-14 S on every side (S = largest coordinate magnitude of the scene) and rounded outwards, the
// slab interval is widened by 2^-20 relative, and the closest-hit cut-off by 1e-5 relative.  The margins
// assume ray origins within ~2^6 S of the scene.
//
// Layout: a binary hierarchy, 64-byte nodes that hold the boxes of BOTH children (one visit = four 128-bit
// loads, two box tests):
//   n[0] = {c0.lo.x, c0.hi.x, c0.lo.y, c0.hi.y}   n[1] = {c1.lo.x, c1.hi.x, c1.lo.y, c1.hi.y}
//   n[2] = {c0.lo.z, c0.hi.z, c1.lo.z, c1.hi.z}   n[3] = {child0, child1, -, -} (int bits)
// child >= 0: node index; child < 0: leaf ~child = first << 3 | count into prims[] (count <= kBvhLeafMax, 0 = empty).
// prims[] entry: kind << 30 | index -- sphere i, fast-path triangle j, or n-gon by its index in the plane arrays.
#pragma once

#include "rm_math.cuh"

namespace rm {

constexpr int kBvhStack = 64;      // the builder switches to median splits below depth 40: 40 + log2(n) < 64
constexpr int kBvhLeafMax = 4;
constexpr int kBvhMaxLights = 8;   // lights the hierarchy kernel shades (unrolled pairs, rm_fast.cuh direct())
enum BvhKind { BVH_SPHERE = 0, BVH_TRI = 1, BVH_POLY = 2 };

struct BvhView {
    const R4<float>* nodes = nullptr;
    const int* prims = nullptr;
    int n_nodes = 0;
    // device only: 16 words a walk writes when it gives up (see bvh_walk); word 0 is the flag
    int* status = nullptr;
};

// Walk statistics of the host emulation (tests/emu, -DRM_EMU_STATS): node visits and primitive tests per walk -- the
// algorithmic cost of a hierarchy, comparable across builder settings without a GPU.  Compiles to nothing elsewhere.
#if defined(RM_EMU_STATS) && !defined(__CUDA_ARCH__)
struct BvhStats { unsigned long long walks = 0, nodes = 0, prims = 0; };
inline BvhStats& bvh_stats() { static thread_local BvhStats s; return s; }
#define RM_BVH_STAT(field) (bvh_stats().field++)
#else
#define RM_BVH_STAT(field) ((void)0)
#endif

RM_HD int bvh_int(float f) {
#if defined(__CUDA_ARCH__)
    return __float_as_int(f);
#else
    int i;
    memcpy(&i, &f, 4);
    return i;
#endif
}

struct BvhRay {
    Vec3<float> o, inv;            // inv = 1/d per component (+-inf for a zero component: IEEE slab test)
};
RM_HD BvhRay bvh_ray(const Vec3<float> o, const Vec3<float> d) { return {o, {1.f / d.x, 1.f / d.y, 1.f / d.z}}; }

// Ray/box slab test on [0, tcut].  fminf/fmaxf return the other operand for a NaN (0 * inf: origin exactly on
// a face plane of a box the ray runs parallel to), which degrades that axis to a single plane -- conservative.
RM_HD bool bvh_slab(const float lox, const float hix, const float loy, const float hiy, const float loz, const float hiz,
                    const BvhRay& r, const float tcut, float& tn) {
    const float ax = (lox - r.o.x) * r.inv.x, bx = (hix - r.o.x) * r.inv.x;
    const float ay = (loy - r.o.y) * r.inv.y, by = (hiy - r.o.y) * r.inv.y;
    const float az = (loz - r.o.z) * r.inv.z, bz = (hiz - r.o.z) * r.inv.z;
    const float t0 = fmaxf(fmaxf(fminf(ax, bx), fminf(ay, by)), fmaxf(fminf(az, bz), 0.f));
    const float t1 = fminf(fminf(fmaxf(ax, bx), fmaxf(ay, by)), fmaxf(az, bz));
    tn = t0;
    return (t0 <= t1 * 1.000001f) & (t0 <= tcut);
}

// Work counters of the counting instantiation (RmParams.accel = 2): node visits (one visit = both children's boxes:
// 64 bytes, two slab tests) and leaf entries tested, per thread.
struct BvhCount {
    unsigned nodes = 0, sph = 0, pln = 0;
};

// Depth-first walk, nearer child first.  `leaf(entry)` tests one primitive and returns true to end the walk
// (any-hit); `cut()` is the current closest-hit cut-off (the walk skips boxes entered beyond it).
template <bool kCount = false, class Leaf, class Cut>
RM_HD bool bvh_walk(const BvhView& bv, const Vec3<float> o, const Vec3<float> d, Leaf&& leaf, Cut&& cut, BvhCount* cnt = nullptr) {
    if (bv.n_nodes <= 0) return false;
    const BvhRay r = bvh_ray(o, d);
    int stack[kBvhStack];
    int sp = 0, cur = 0;
    // A walk visits every node at most once.  Corrupt node memory must not be able to hang the GPU: past that bound the
    // walk gives up (reports a miss), raises the scene's status flag and leaves the ray behind for the host to report.
    int budget = 2 * bv.n_nodes + 8;
    RM_BVH_STAT(walks);
    for (;;) {
        while (cur >= 0) {
            if (--budget < 0 || sp >= kBvhStack - 1 || cur >= bv.n_nodes) {
#if defined(__CUDA_ARCH__)
                if (bv.status && atomicExch(bv.status, 1) == 0) {
                    float* f = reinterpret_cast<float*>(bv.status);
                    f[1] = o.x; f[2] = o.y; f[3] = o.z; f[4] = d.x; f[5] = d.y; f[6] = d.z;
                    bv.status[7] = cur; bv.status[8] = sp; bv.status[9] = budget;
                }
#endif
                return false;
            }
            RM_BVH_STAT(nodes);
            if constexpr (kCount) cnt->nodes++;
            const R4<float>* n = bv.nodes + 4 * (size_t)cur;
            const R4<float> a = n[0], b = n[1], z = n[2], c = n[3];
            const float tc = cut();
            float t0, t1;
            const bool h0 = bvh_slab(a.x, a.y, a.z, a.w, z.x, z.y, r, tc, t0);
            const bool h1 = bvh_slab(b.x, b.y, b.z, b.w, z.z, z.w, r, tc, t1);
            const int c0 = bvh_int(c.x), c1 = bvh_int(c.y);
            if (h0 & h1) {
                const bool swap = t1 < t0;
                stack[sp++] = swap ? c0 : c1;
                cur = swap ? c1 : c0;
            } else if (h0) {
                cur = c0;
            } else if (h1) {
                cur = c1;
            } else {
                if (sp == 0) return false;
                cur = stack[--sp];
            }
        }
        const int code = ~cur, first = code >> 3, cnt = code & 7;
        for (int k = 0; k < cnt; k++) {
            RM_BVH_STAT(prims);
            if (leaf(bv.prims[first + k])) return true;
        }
        if (sp == 0) return false;
        cur = stack[--sp];
    }
}

}  // namespace rm
