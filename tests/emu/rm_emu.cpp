// rm_emu.cpp -- DEV/TEST TOOL, not part of the product and never loaded by it.
//
// Compiles the product's own kernel code (rusty_marcher_b200/csrc/rm_trace.cuh, the same
// templates the CUDA kernels instantiate) for the host, so that logic and FP32-vs-FP64 parity
// can be studied in this GPU-less container before spending GPU minutes.  It is NOT a fallback:
// librm_b200.so does not contain it and fails loudly without a GPU.  Differences to the device:
// powf/sqrtf/division come from glibc instead of the CUDA math library.
#define RM_EMU_STATS 1
#include <atomic>
#include <cstring>
#include <thread>
#include <vector>

#include "../../rusty_marcher_b200/csrc/rm_scene.cpp"
#include "../../rusty_marcher_b200/csrc/rm_bvh.cpp"
#include "../../rusty_marcher_b200/csrc/rm_host.cpp"

// rm_host.cpp's rm_builder_upload needs this symbol; the emulator has no device.
extern "C" int rm_scene_upload(const RmFlatScene*, RmScene*) { return RM_ERR_NO_DEVICE; }

namespace {

template <typename R>
rm::SceneView<R> view_of(const rm::PackedScene<R>& ps, bool cull) {
    rm::SceneView<R> sc;
    const rm::BlobLayout& L = ps.lay;
    const unsigned char* base = ps.blob_data();
    sc.sph = reinterpret_cast<const rm::R4<R>*>(base + L.off_sph);
    sc.sph_id = reinterpret_cast<const int*>(base + L.off_sph_id);
    sc.n_sph = L.n_sph;
    sc.pln_n = reinterpret_cast<const rm::R4<R>*>(base + L.off_pln_n);
    sc.pln_c = reinterpret_cast<const rm::R4<R>*>(base + L.off_pln_c);
    sc.pln_v = reinterpret_cast<const rm::I2*>(base + L.off_pln_v);
    sc.pln_id = reinterpret_cast<const int*>(base + L.off_pln_id);
    sc.n_pln = rm::plane_count<R>(L, cull);
    sc.vert = reinterpret_cast<const rm::VertT<R>*>(base + L.off_vert);
    sc.mat_a = ps.mat_a.data();
    sc.mat_b = ps.mat_b.data();
    sc.mat_f = ps.mat_f.data();
    sc.lgt_p = reinterpret_cast<const rm::R4<R>*>(base + L.off_lgt_p);
    sc.lgt_c = reinterpret_cast<const rm::R4<R>*>(base + L.off_lgt_c);
    sc.n_lgt = L.n_lgt;
    sc.order = ps.order[cull].data();
    sc.order_shape = ps.order_shape[cull].data();
    sc.n_order = (int)ps.order[cull].size();
    return sc;
}

template <typename R>
int emu_render_impl(const RmFlatScene* fs, const RmParams* p, R* out_rgb, int32_t* prim, RmStats* stats, int n_threads) {
    rm::PackedScene<R> ps;
    std::string err;
    int rc = rm::pack_scene<R>(*fs, ps, err);
    if (rc != RM_OK) return rc;
    if (p->width % 32) return RM_ERR_DIMENSIONS;
    const bool cull = p->cull_backfacing != 0;
    const rm::SceneView<R> sc = view_of(ps, cull);
    const rm::FrameParams<R> fp = rm::make_frame_params<R>(*p);
    std::atomic<int> next{fp.row_begin};
    std::vector<rm::Counters<true>> tc(n_threads);
    std::vector<R> tmax(n_threads, R(0));
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) {
        pool.emplace_back([&, t] {
            rm::Counters<true>& st = tc[t];
            st.clear();
            for (;;) {
                int y = next.fetch_add(1);
                if (y >= fp.row_end) break;
                for (int x = 0; x < fp.width; x++) {
                    st.add(rm::C_PIXELS);
                    int pid;
                    rm::Vec3<R> dir = rm::backproject<R>(fp, x, y);
                    rm::Vec3<R> c = rm::cast_ray<R, true, rm::SceneView<R>>(sc, fp.camera, dir, fp.background, fp.max_depth, pid, st);
                    size_t px = (size_t)y * fp.width + x;
                    out_rgb[3 * px] = c.x; out_rgb[3 * px + 1] = c.y; out_rgb[3 * px + 2] = c.z;
                    if (prim) prim[px] = pid;
                    R m = rm::Num<R>::max_(rm::Num<R>::max_(c.x, c.y), c.z);
                    if (m > tmax[t]) tmax[t] = m;
                }
            }
        });
    }
    for (auto& th : pool) th.join();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        uint64_t* dst = &stats->pixels;
        for (auto& c : tc) for (int i = 0; i < rm::C_COUNT; i++) dst[i] += c.c[i];
        R m = 0;
        for (R v : tmax) m = v > m ? v : m;
        stats->max_value = (double)m;
        stats->resident_prims = sc.n_sph + sc.n_pln;
    }
    return RM_OK;
}

}  // namespace

static std::atomic<unsigned long long> g_walks{0}, g_walk_nodes{0}, g_walk_prims{0};   // hierarchy walks of all renders since the last reset
static std::atomic<float*> g_cost{nullptr};   // optional H x W x 2 floats: per-pixel walk cost of stage A / of shading (tools/tail_model.py)
static std::atomic<int> g_glass_mode{-1}, g_last_glass_mode{-1};   // -1: as launch_fast decides (rm::glass_mode); 1 / 2 force GLASS_F32 / GLASS_F64
static std::atomic<int> g_strip_bound{1};   // 0: walk every triangle in every strip (to prove the bound changes no pixel)

// the FP32 production path (rm_fast.cuh): prepare_raster + fast_pixel, as the CUDA kernels run them
// (kBvh: RmParams.accel -- every query through the hierarchy of rm_bvh.cuh, stage A pixel by pixel)
template <bool kBvh>
int emu_render_fast_impl(const RmFlatScene* fs, const RmParams* p, float* out_rgb, int32_t* prim, RmStats* stats, int n_threads) {
    rm::PackedScene<float> ps;
    std::string err;
    int rc = rm::pack_scene<float>(*fs, ps, err);
    if (rc != RM_OK) return rc;
    if (p->width % 32) return RM_ERR_DIMENSIONS;
    const bool cull = p->cull_backfacing != 0;
    const rm::BlobLayout& L = ps.lay;
    const unsigned char* base = ps.blob_data();
    const int n_tri = rm::tri_count(L, cull);
    std::vector<rm::R4<float>> tri_r((size_t)n_tri * 4 + 4);
    for (int j = 0; j < n_tri; j++) rm::prepare_raster(ps.tri_src.data() + (size_t)j * rm::kTriSrcDoubles, p->camera, tri_r.data() + 4 * j);
    rm::FastViewT<kBvh> fv0;
    fv0.bvh.nodes = ps.bvh_nodes.data();
    fv0.bvh.prims = ps.bvh_prims.data();
    fv0.bvh.n_nodes = (int)(ps.bvh_nodes.size() / 4);
    fv0.sph = reinterpret_cast<const rm::R4<float>*>(base + L.off_sph);
    fv0.sph_id = reinterpret_cast<const int*>(base + L.off_sph_id);
    fv0.n_sph = L.n_sph;
    fv0.tri_g = reinterpret_cast<const rm::R4<float>*>(base + L.off_tri_g);
    fv0.tri_r = tri_r.data();
    fv0.n_tri = n_tri;
    fv0.poly_slot = reinterpret_cast<const int*>(base + L.off_poly_slot);
    fv0.n_poly = rm::poly_count(L, cull);
    fv0.pln_n = reinterpret_cast<const rm::R4<float>*>(base + L.off_pln_n);
    fv0.pln_c = reinterpret_cast<const rm::R4<float>*>(base + L.off_pln_c);
    fv0.pln_v = reinterpret_cast<const rm::I2*>(base + L.off_pln_v);
    fv0.pln_id = reinterpret_cast<const int*>(base + L.off_pln_id);
    fv0.vert = reinterpret_cast<const rm::R4<float>*>(base + L.off_vert);
    fv0.mat_a = ps.mat_a.data();
    fv0.mat_b = ps.mat_b.data();
    fv0.mat_f = ps.mat_f.data();
    fv0.lgt_p = reinterpret_cast<const rm::R4<float>*>(base + L.off_lgt_p);
    fv0.lgt_c = reinterpret_cast<const rm::R4<float>*>(base + L.off_lgt_c);
    fv0.n_lgt = L.n_lgt;
    fv0.sph64 = ps.sph64.data();
    fv0.tri64 = ps.tri_src.data();
    fv0.pln64 = ps.pln64.data();
    const rm::FrameParams<float> fp = rm::make_frame_params<float>(*p);
    const int gm = g_glass_mode.load() >= 0 ? g_glass_mode.load() : rm::glass_mode(L.any_glass != 0, L.n_sph, L.coord_max, L.r_min, p->camera);
    g_last_glass_mode.store(gm);
    std::atomic<int> next{fp.row_begin};
    std::vector<float> tmax(n_threads, 0.f);
    std::vector<std::thread> pool;
    for (int t = 0; t < n_threads; t++) {
        pool.emplace_back([&, t] {
            rm::FastViewT<kBvh> fv = fv0;
            std::vector<int> cand;
            for (;;) {
                // one 4-row band per grab; inside it the kernel's 32x4 warp strips, each with its own exact
                // triangle bound (tri_may_touch) followed by the per-thread walk over the survivors
                int y0 = next.fetch_add(4);
                if (y0 >= fp.row_end) break;
                for (int xs = 0; xs < fp.width; xs += 32) {
                    const float Xa = rm::pixel_X(fp, xs), Xb = rm::pixel_X(fp, xs + 31);
                    const float Ya = rm::pixel_Y(fp, y0), Yb = rm::pixel_Y(fp, y0 + 3);
                    cand.clear();
                    // the classify kernel's tile-level bound (32x32 tile around this strip): provably empty tiles are only zero-filled
                    bool tile_busy = kBvh || fv.n_sph + fv.n_poly > 0 || !g_strip_bound.load();
                    {
                        const int ty0 = fp.row_begin + ((y0 - fp.row_begin) / 32) * 32;
                        const float TYa = rm::pixel_Y(fp, ty0), TYb = rm::pixel_Y(fp, ty0 + 31);
                        for (int j = 0; j < n_tri && !tile_busy; j++)
                            tile_busy = rm::tri_may_touch(tri_r[4 * j], tri_r[4 * j + 1], tri_r[4 * j + 2], tri_r[4 * j + 3], Xa, Xb, TYa, TYb);
                    }
                    if (!tile_busy) {
                        for (int y = y0; y < y0 + 4; y++)
                            for (int x = xs; x < xs + 32; x++) {
                                size_t px = (size_t)y * fp.width + x;
                                out_rgb[3 * px] = out_rgb[3 * px + 1] = out_rgb[3 * px + 2] = 0.f;
                                if (prim) prim[px] = -1;
                            }
                        continue;
                    }
                    for (int j = 0; j < n_tri && !kBvh; j++)
                        if (!g_strip_bound.load() || rm::tri_may_touch(tri_r[4 * j], tri_r[4 * j + 1], tri_r[4 * j + 2], tri_r[4 * j + 3], Xa, Xb, Ya, Yb)) cand.push_back(j);
                    for (int lane = 0; lane < 32; lane++) {
                        const int x = xs + (lane & 7) * 4, y = y0 + (lane >> 3);
                        rm::PrimaryState<4> ps;
                        rm::primary_begin<4>(ps, fp, x, y);
                        float* const cost = g_cost.load();
                        auto walk_cost = [] { const rm::BvhStats& b = rm::bvh_stats(); return (float)b.nodes + 0.7f * (float)b.prims; };
                        if constexpr (kBvh) {
                            const float c0 = walk_cost();
                            rm::primary_bvh<4>(ps, fv, fp);
                            if (cost)
                                for (int k = 0; k < 4; k++) cost[2 * ((size_t)y * fp.width + x + k)] = 0.25f * (walk_cost() - c0);
                        } else {
                            for (int j : cand) rm::primary_tri<4>(ps, tri_r[4 * j], tri_r[4 * j + 1], tri_r[4 * j + 2], tri_r[4 * j + 3], fv.n_sph + j);
                            if (fv.n_sph + fv.n_poly > 0) rm::primary_rest<4>(ps, fv, fp);
                        }
                        for (int k = 0; k < 4; k++) {
                            const float c1 = walk_cost();
                            rm::Vec3<float> c = {0.f, 0.f, 0.f};
                            if (ps.slot[k] >= 0)            // the instantiation launch_fast picks for this scene and camera
                                c = gm == rm::GLASS_NONE  ? rm::fast_shade<rm::GLASS_NONE>(fv, fp, x + k, y, ps.t[k], ps.slot[k], ps.id[k])
                                    : gm == rm::GLASS_F32 ? rm::fast_shade<rm::GLASS_F32>(fv, fp, x + k, y, ps.t[k], ps.slot[k], ps.id[k])
                                                          : rm::fast_shade<rm::GLASS_F64>(fv, fp, x + k, y, ps.t[k], ps.slot[k], ps.id[k]);
                            size_t px = (size_t)y * fp.width + x + k;
                            if (cost && kBvh) cost[2 * px + 1] = walk_cost() - c1;
                            out_rgb[3 * px] = c.x; out_rgb[3 * px + 1] = c.y; out_rgb[3 * px + 2] = c.z;
                            if (prim) prim[px] = ps.id[k];
                            float m = fmaxf(fmaxf(c.x, c.y), c.z);
                            if (m > tmax[t]) tmax[t] = m;
                        }
                    }
                }
            }
            rm::BvhStats& bs = rm::bvh_stats();
            g_walks += bs.walks; g_walk_nodes += bs.nodes; g_walk_prims += bs.prims;
            bs = rm::BvhStats();
        });
    }
    for (auto& th : pool) th.join();
    if (stats) {
        std::memset(stats, 0, sizeof(*stats));
        float m = 0;
        for (float v : tmax) m = v > m ? v : m;
        stats->max_value = (double)m;
        stats->resident_prims = L.n_sph + n_tri + fv0.n_poly;
    }
    return RM_OK;
}

// FNV-1a digests of everything the packer produces for a scene (layout, hot blob, records, materials, scene-order lists,
// hierarchy): out[0] the FP32 pack, out[1] the FP64 pack.  The packer runs on several threads for large scenes; its output
// must not depend on how many.
namespace {
unsigned long long fnv(const void* p, size_t n, unsigned long long h) {
    const unsigned char* b = static_cast<const unsigned char*>(p);
    for (size_t i = 0; i < n; i++) {
        h ^= b[i];
        h *= 1099511628211ull;
    }
    return h;
}
template <typename R> unsigned long long pack_digest(const rm::PackedScene<R>& ps) {
    unsigned long long h = 1469598103934665603ull;
    const rm::BlobLayout L = ps.lay;
    h = fnv(&L.n_sph, sizeof(int) * 6, h);
    h = fnv(&L.off_sph, sizeof(int) * 9, h);
    h = fnv(&L.n_tri, sizeof(int) * 7, h);
    h = fnv(&L.any_glass, 4, h);
    h = fnv(&L.coord_max, 16, h);
    h = fnv(ps.blob.data(), ps.blob_bytes(), h);
    h = fnv(ps.tri_src.data(), ps.tri_src.size() * 8, h);
    h = fnv(ps.sph64.data(), ps.sph64.size() * 8, h);
    h = fnv(ps.pln64.data(), ps.pln64.size() * 8, h);
    h = fnv(ps.mat_a.data(), ps.mat_a.size() * sizeof(ps.mat_a[0]), h);
    h = fnv(ps.mat_b.data(), ps.mat_b.size() * sizeof(ps.mat_b[0]), h);
    h = fnv(ps.mat_f.data(), ps.mat_f.size() * 4, h);
    for (int k = 0; k < 2; k++) {
        h = fnv(ps.order[k].data(), ps.order[k].size() * 4, h);
        h = fnv(ps.order_shape[k].data(), ps.order_shape[k].size() * 4, h);
    }
    h = fnv(&ps.n_prims, 4, h);
    h = fnv(ps.bvh_nodes.data(), ps.bvh_nodes.size() * 16, h);
    h = fnv(ps.bvh_prims.data(), ps.bvh_prims.size() * 4, h);
    h = fnv(&ps.bvh_depth, 4, h);
    return h;
}
}  // namespace
extern "C" {
void emu_set_strip_bound(int on) { g_strip_bound.store(on); }
void emu_set_glass_mode(int mode) { g_glass_mode.store(mode); }
int emu_last_glass_mode(void) { return g_last_glass_mode.load(); }
void emu_set_cost_buffer(float* hw2) { g_cost.store(hw2); }
// {walks, node visits, primitive tests} of the hierarchy walks since the last call (accel renders only)
void emu_walk_stats(unsigned long long out[3]) {
    out[0] = g_walks.exchange(0); out[1] = g_walk_nodes.exchange(0); out[2] = g_walk_prims.exchange(0);
}
int emu_render_fast(const RmFlatScene* fs, const RmParams* p, float* out_rgb, int32_t* prim, RmStats* stats, int n_threads) {
    // (the device keeps scenes with more than kBvhMaxLights lights on the brute-force kernel: launch_fast)
    return p->accel && fs->n_lights <= rm::kBvhMaxLights ? emu_render_fast_impl<true>(fs, p, out_rgb, prim, stats, n_threads)
                    : emu_render_fast_impl<false>(fs, p, out_rgb, prim, stats, n_threads);
}
// the hierarchy of a scene's FP32 pack, for the builder's invariants: nodes (16 floats each) and leaf entries
int emu_bvh(const RmFlatScene* fs, float* nodes, int nodes_cap, int* prims, int prims_cap, int* n_nodes, int* n_prims, int* depth) {
    rm::PackedScene<float> ps;
    std::string err;
    int rc = rm::pack_scene<float>(*fs, ps, err);
    if (rc != RM_OK) return rc;
    *n_nodes = (int)(ps.bvh_nodes.size() / 4);
    *n_prims = (int)ps.bvh_prims.size();
    *depth = ps.bvh_depth;
    if (*n_nodes > nodes_cap || *n_prims > prims_cap) return RM_ERR_INVALID_ARGUMENT;
    std::memcpy(nodes, ps.bvh_nodes.data(), ps.bvh_nodes.size() * sizeof(rm::R4<float>));
    std::memcpy(prims, ps.bvh_prims.data(), ps.bvh_prims.size() * sizeof(int));
    return RM_OK;
}
int emu_pack_digest(const RmFlatScene* fs, unsigned long long out[2]) {
    std::string err;
    rm::PackedScene<float> pf;
    int rc = rm::pack_scene<float>(*fs, pf, err);
    if (rc != RM_OK) return rc;
    rm::PackedScene<double> pd;
    rc = rm::pack_scene<double>(*fs, pd, err);
    if (rc != RM_OK) return rc;
    out[0] = pack_digest(pf);
    out[1] = pack_digest(pd);
    return RM_OK;
}
int emu_render_f32(const RmFlatScene* fs, const RmParams* p, float* out_rgb, int32_t* prim, RmStats* stats, int n_threads) {
    return emu_render_impl<float>(fs, p, out_rgb, prim, stats, n_threads);
}
int emu_render_f64(const RmFlatScene* fs, const RmParams* p, double* out_rgb, int32_t* prim, RmStats* stats, int n_threads) {
    return emu_render_impl<double>(fs, p, out_rgb, prim, stats, n_threads);
}
void rm_params_default(RmParams* p, int width, int height) {
    std::memset(p, 0, sizeof(*p));
    p->width = width; p->height = height; p->fov = 1.5; p->max_depth = 3; p->background = 0.1;
    p->patch_size = 32; p->precision = RM_FP32; p->patch_row_begin = 0; p->patch_row_end = -1; p->cull_backfacing = 1;
}
}
