/*
 * rm_oracle.h -- C API of the CPU ORACLE for rusty-marcher's per-pixel render hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  This is a plain f64 C++ restatement of the reference
 * algorithm (engine/src/renderer.rs, shapes.rs, sphere.rs, triangle.rs, polygon.rs,
 * obj.rs, optics.rs, geometry.rs, lights.rs, scene.rs, framebuffer.rs).  Only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 * The product library (rusty_marcher_b200/csrc) never links, includes or calls it.
 *
 * Parity pin: the 800x600 demo render of this oracle is byte-identical to the
 * reference's own golden image engine/out.ppm (sha256 82d51afa...4797), see
 * tests/test_oracle_golden.py.  The OBJ reader restates the un-vendored `tobj` crate
 * (engine/Cargo.toml:8, version "*") and is "parity unpinned": the reference holds no
 * test that asserts triangle count/order (engine/src/obj.rs:229-233 only checks is_some()).
 */
#ifndef RM_ORACLE_H
#define RM_ORACLE_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* engine/src/shapes.rs:21-32 */
typedef struct OrcReflectance {
    double diffusion;
    double diffuse_color[3];
    double specular;
    double specular_exponent;
    int32_t is_glass_like;
    double reflection;
    double refractive_index;
} OrcReflectance;

/* Event counters: the "algorithmic work" of SURVEY.md 8(d).  Closest-hit queries count
 * every primitive; any-hit queries count up to and including the first hitting primitive. */
typedef struct OrcCounters {
    uint64_t pixels;            /* primary rays generated (renderer.rs:128-135)            */
    uint64_t closest_segments;  /* find_closest_intersect calls (renderer.rs:266)          */
    uint64_t anyhit_segments;   /* intersect_shape_set calls (renderer.rs:174)             */
    uint64_t sphere_tests;      /* sphere.rs:28-38 reject stage                            */
    uint64_t sphere_disc;       /* sphere.rs:40-51 discriminant stage                      */
    uint64_t sphere_hits;       /* sphere.rs:54-60 point+normal                            */
    uint64_t plane_tests;       /* triangle.rs:56 / polygon.rs:65  d.n                     */
    uint64_t plane_dist;        /* triangle.rs:62 / polygon.rs:71                          */
    uint64_t plane_point;       /* triangle.rs:69 / polygon.rs:78                          */
    uint64_t edge_tests;        /* triangle.rs:13-15 / polygon.rs:54-56                    */
    uint64_t cand_dist;         /* shapes.rs:128, obj.rs:197 candidate distance^2          */
    uint64_t hits;              /* closest segments that hit (renderer.rs:272,192)         */
    uint64_t light_evals;       /* renderer.rs:166-172                                     */
    uint64_t lit_lights;        /* renderer.rs:180-189 (one pow each)                      */
    uint64_t glass_hits;        /* optics.rs:15-35,56-76 preambles                         */
    uint64_t reflections;       /* optics.rs:38-47 spawned                                 */
    uint64_t refractions;       /* optics.rs:78-88 spawned                                 */
} OrcCounters;

typedef struct OrcScene OrcScene;

OrcScene* orc_scene_new(void);                      /* scene.rs:16-23 */
OrcScene* orc_scene_create_default(void);           /* scene.rs:28-211 */
void      orc_scene_free(OrcScene*);
void      orc_scene_set_camera(OrcScene*, const double xyz[3]);
void      orc_scene_offset_camera(OrcScene*, const double xyz[3]);   /* scene.rs:25-27 */
void      orc_reflectance_default(OrcReflectance* out);              /* shapes.rs:50-60 */
int       orc_scene_add_sphere(OrcScene*, const double center[3], double radius, const OrcReflectance*);
int       orc_scene_add_polygon(OrcScene*, const double* verts_xyz, int n_vertices, const OrcReflectance*);
/* obj.rs:44-151 + main.rs:278-288: load, one shape per model, offset every model. Returns #models or -1. */
int       orc_scene_add_obj_file(OrcScene*, const char* path, const double offset[3]);
/* one Obj shape from raw triangles (f64 vertices, 9 per triangle), gradient colours as obj.rs:125-138 */
int       orc_scene_add_mesh(OrcScene*, const double* tri_verts, int n_triangles, const double offset[3]);
void      orc_scene_add_light(OrcScene*, const double pos[3], const double color[3], double intensity);
/* Extension mode (no reference counterpart): every primitive of `shape` becomes glass-like with the
 * given reflection / refractive index / diffusion (returns the number of materials changed, -1 for a bad index); and
 * every glass-like material of the scene gets a new refractive index (one pass of the per-channel dispersion). */
int       orc_scene_make_glass(OrcScene*, int shape, double reflection, double refractive_index, double diffusion);
int       orc_scene_set_glass_index(OrcScene*, double refractive_index);
int       orc_scene_num_shapes(const OrcScene*);
int       orc_scene_num_prims(const OrcScene*);     /* flattened primitive count (prim_id space) */
/* triangles of shape `shape` (0 if not an Obj); optionally copies 9 doubles per triangle */
int       orc_scene_obj_triangles(const OrcScene*, int shape, double* out_verts, int max_tris);

/*
 * Renderer::render (renderer.rs:36-126).  Renders patch rows [patch_row_begin, patch_row_end)
 * of 32x32 patches (pass 0,-1 for all floor(H/32) rows) with n_threads worker threads pulling
 * patches from an atomic counter (Rayon analogue).  patch_stride>1 renders every stride-th patch
 * only (bounded CPU-baseline samples).  out_rgb is H*W*3 doubles, rows outside the rendered range
 * are left untouched.  prim_id (H*W int32, optional): flattened primitive index hit by the primary
 * ray, -1 for a miss.  fragile (H*W uint8, optional): bit0 = a primary closest-hit decision had a
 * margin below f32 resolution, bit1 = some decision anywhere in the pixel's ray tree did.
 * Returns 0, or -1 when W is not a multiple of 32 (the reference indexes out of bounds there).
 */
int orc_render(const OrcScene*, int width, int height, double fov, int max_depth,
               int n_threads, int patch_row_begin, int patch_row_end, int patch_stride,
               double* out_rgb, int32_t* prim_id, uint8_t* fragile, OrcCounters* counters);

/* framebuffer.rs:58-77 (global max, multiply by 1/max), :40-55,80-82 (clamp, truncate), :26-38 */
double orc_normalize(double* rgb, int width, int height);
void   orc_to_vec(const double* rgb, int width, int height, uint8_t* out_rgb8);
int    orc_write_ppm(const char* path, const double* rgb, int width, int height);

/* unit-test hooks for the reference's own #[test]s */
void   orc_vec_normalized(const double v[3], double out[3]);          /* geometry.rs:104-109 */
void   orc_vec_normalized_l0(const double v[3], double out[3]);       /* geometry.rs:111-116 */
double orc_vec_dot(const double a[3], const double b[3]);             /* geometry.rs:180-182 */
void   orc_vec_cross(const double a[3], const double b[3], double out[3]); /* geometry.rs:57-63 */
void   orc_vec_scaled(const double a[3], double s, double out[3]);    /* geometry.rs:49-53 */
void   orc_reflect(const double incident[3], const double normal[3], double out[3]);  /* optics.rs:4-6 */
int    orc_reflect_ray(const double incident[3], const double point[3], const double normal[3],
                       double refractive_index, double out_orig[3], double out_dir[3]); /* optics.rs:8-48 */
int    orc_refract_ray(const double incident[3], const double point[3], const double normal[3],
                       double refractive_index, double out_orig[3], double out_dir[3]); /* optics.rs:50-89 */
int    orc_triangle_intersect(const double verts[9], const double orig[3], const double dir[3],
                              double out_point[3], double out_normal[3]);   /* triangle.rs:33-83 */
int    orc_sphere_intersect(const double center[3], double radius, const double orig[3],
                            const double dir[3], double out_point[3], double out_normal[3]); /* sphere.rs:27-61 */
/* hits on degenerate-projection planar primitives during the last instrumented orc_render (see rm_oracle.cpp) */
uint64_t orc_last_degenerate_hits(void);
int    orc_hardware_threads(void);

#ifdef __cplusplus
}
#endif
#endif
