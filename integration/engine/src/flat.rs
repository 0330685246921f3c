//! flat.rs -- `Scene` -> `RmFlatScene`.  `Scene.shapes` is a `Vec<Box<dyn Shape + Sync>>` whose concrete types keep
//! their fields private (sphere.rs:6-11, polygon.rs:6-12, obj.rs:14-20), so the `Shape` trait gains one method,
//! `flatten`, with which every shape appends itself to these lists (patches/shapes.rs.patch and friends).
//! Values are copied bit for bit (f64); the library derives plane constants, edge functions, culling classes and
//! its hierarchy itself, in f64, before anything is rounded to f32.
use ffi::*;
use geometry::Vec3f;
use lights::Light;
use scene::Scene;
use shapes::Reflectance;

#[derive(Default)]
pub struct FlatSceneBuilder {
    pub shapes: Vec<RmShapeRef>,
    pub spheres: Vec<RmSphere>,
    pub polygons: Vec<RmPolygon>,
    pub polygon_vertices: Vec<f64>,
    pub objs: Vec<RmObj>,
    pub triangles: Vec<RmTriangle>,
    pub triangle_reflectances: Vec<RmReflectance>,
    pub lights: Vec<RmLight>,
}

pub fn v3(v: &Vec3f) -> [f64; 3] {
    [v.x, v.y, v.z]
}

impl<'a> From<&'a Reflectance> for RmReflectance {
    fn from(r: &'a Reflectance) -> RmReflectance {
        RmReflectance {
            diffusion: r.diffusion,
            diffuse_color: v3(&r.diffuse_color),
            specular: r.specular,
            specular_exponent: r.specular_exponent,
            is_glass_like: r.is_glass_like as i32,
            reflection: r.reflection,
            refractive_index: r.refractive_index,
        }
    }
}

impl<'a> From<&'a Light> for RmLight {
    fn from(l: &'a Light) -> RmLight {
        // create_light has already L-inf normalised the colour (lights.rs:13)
        RmLight { position: v3(&l.position), color: v3(&l.color), intensity: l.intensity }
    }
}

impl FlatSceneBuilder {
    /// scene.shapes in order (the primitive ids the library reports follow this order), then the lights
    pub fn from_scene(scene: &Scene) -> FlatSceneBuilder {
        let mut flat = FlatSceneBuilder::default();
        for s in &scene.shapes {
            s.flatten(&mut flat);
        }
        for l in &scene.lights {
            flat.lights.push(l.into());
        }
        flat
    }

    pub fn n_prims(&self) -> usize {
        self.spheres.len() + self.polygons.len() + self.triangles.len()
    }

    /// Raw-pointer view; valid while `self` is alive and unchanged.
    pub fn as_c(&self) -> RmFlatScene {
        RmFlatScene {
            n_shapes: self.shapes.len() as i32,
            shapes: self.shapes.as_ptr(),
            n_spheres: self.spheres.len() as i32,
            spheres: self.spheres.as_ptr(),
            n_polygons: self.polygons.len() as i32,
            polygons: self.polygons.as_ptr(),
            n_polygon_vertices: (self.polygon_vertices.len() / 3) as i32,
            polygon_vertices: self.polygon_vertices.as_ptr(),
            n_objs: self.objs.len() as i32,
            objs: self.objs.as_ptr(),
            n_triangles: self.triangles.len() as i32,
            triangles: self.triangles.as_ptr(),
            triangle_reflectances: self.triangle_reflectances.as_ptr(),
            n_lights: self.lights.len() as i32,
            lights: self.lights.as_ptr(),
        }
    }
}
