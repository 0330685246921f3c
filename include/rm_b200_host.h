/*
 * rm_b200_host.h -- C ABI of the host-side scene builder that sits above rm_b200.h.
 *
 * It mirrors the constructors of the reference's engine crate so that FFI users without their
 * own geometry code (the Python package, the C++ harness, tests) build scenes exactly the way
 * `cargo run` does, and it produces the RmFlatScene that rm_scene_upload() consumes.  A Rust
 * caller does not need it: its own Scene already holds these values and flattens them itself
 * (`Shape::flatten`).
 *
 *   rm_builder_new              Scene::new                     engine/src/scene.rs:16-23
 *   rm_builder_create_default   Scene::create_default          engine/src/scene.rs:28-211
 *   rm_builder_offset_camera    Scene::offset_camera           engine/src/scene.rs:25-27
 *   rm_builder_add_sphere       sphere::create                 engine/src/sphere.rs:13-25
 *   rm_builder_add_polygon      ConvexPolygon::create          engine/src/polygon.rs:16-42
 *   rm_builder_add_obj_file     obj::load + Obj::offset        engine/src/obj.rs:44-151, main.rs:278-288
 *   rm_builder_add_mesh         Triangle::create per face      engine/src/triangle.rs:33-47
 *   rm_builder_add_light        lights::create_light           engine/src/lights.rs:10-16
 */
#ifndef RM_B200_HOST_H
#define RM_B200_HOST_H

#include "rm_b200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct RmSceneBuilder RmSceneBuilder;

RmSceneBuilder* rm_builder_new(void);
RmSceneBuilder* rm_builder_create_default(void);
void rm_builder_free(RmSceneBuilder* b);

void rm_builder_set_camera(RmSceneBuilder* b, const double xyz[3]);
void rm_builder_offset_camera(RmSceneBuilder* b, const double xyz[3]);
void rm_builder_get_camera(const RmSceneBuilder* b, double xyz[3]);

/* each returns the index of the new shape in Scene.shapes, or a negative RmStatus */
int rm_builder_add_sphere(RmSceneBuilder* b, const double center[3], double radius, const RmReflectance* r);
int rm_builder_add_polygon(RmSceneBuilder* b, const double* vertices_xyz, int n_vertices, const RmReflectance* r);
/* one Obj shape from n_triangles x 9 doubles; colours follow obj.rs:125-138; offset may be NULL */
int rm_builder_add_mesh(RmSceneBuilder* b, const double* triangle_vertices, int n_triangles, const double offset[3]);
/* one Obj shape from already-built triangles (normal/centre as Triangle::create left them, possibly
 * offset afterwards) and their per-triangle reflectances (NULL = the obj.rs:125-138 gradient) */
int rm_builder_add_triangles(RmSceneBuilder* b, const RmTriangle* triangles, const RmReflectance* reflectances,
                             int n_triangles, const char* name);
/* loads a Wavefront OBJ the way the reference does (one shape per model, every model moved by
 * `offset`, main.rs:278-288).  Returns the number of models, or a negative RmStatus. */
int rm_builder_add_obj_file(RmSceneBuilder* b, const char* path, const double offset[3]);
void rm_builder_add_light(RmSceneBuilder* b, const double position[3], const double color[3], double intensity);

int rm_builder_num_shapes(const RmSceneBuilder* b);
int rm_builder_num_prims(const RmSceneBuilder* b);
/* name of the model behind shape `shape` ("" for spheres and polygons) */
const char* rm_builder_shape_name(const RmSceneBuilder* b, int shape);

/* The flattened scene; the pointer stays valid until the builder is changed or freed. */
const RmFlatScene* rm_builder_flatten(RmSceneBuilder* b);
/* rm_scene_upload(rm_builder_flatten(b)) */
int rm_builder_upload(RmSceneBuilder* b, RmScene* out_handle);

/* FrameBuffer::write_ppm (framebuffer.rs:26-38) for the 8-bit frame the kernels deliver (rm_render's out_rgb8, or
 * rm_render_frame's frame on rank 0 after a copy to the host): "P6\n{W} {H}\n255\n" followed by H*W*3 bytes. */
int rm_write_ppm(const char* path, int width, int height, const uint8_t* rgb8);

#ifdef __cplusplus
}
#endif
#endif /* RM_B200_HOST_H */
