// rm_kernels.h -- launchers of the sm_100a kernels (rm_kernels.cu), used by the C ABI (rm_api.cu).
#pragma once

#include <cuda_runtime.h>

#include "rm_scene.h"

namespace rm {

// The packed scene in HBM (see rm_scene.h for the blob layout).
template <typename R> struct DeviceScene {
    const unsigned char* blob = nullptr;
    BlobLayout lay;
    const R4<R>* mat_a = nullptr;
    const R4<R>* mat_b = nullptr;
    const int* mat_f = nullptr;
    const int* order[2] = {nullptr, nullptr};        // [0] every primitive, [1] after culling
    const int* order_shape[2] = {nullptr, nullptr};
    int n_order[2] = {0, 0};
    // FP32 fast path: FP64 triangle sources (input of the per-frame prepare kernel) and the
    // camera-specialised raster records it writes (4 x R4<float> per triangle, rebuilt every frame)
    const double* tri_src = nullptr;
    R4<float>* tri_r = nullptr;
    // frame control block of the persistent render kernel, zero between frames (the last CTA to finish resets it):
    //   ctr[0] next busy tile   ctr[1] fully covered tiles   ctr[2] empty tiles   ctr[3] CTAs finished   ctr[5] partially covered tiles
    int* ctr = nullptr;
    // tile schedule written by the classify kernel: busy tiles (some triangle may touch them) first, then the
    // tiles that are provably empty and only need their black pixels stored
    int* tile_order = nullptr;
    int tile_order_cap = 0;
};

// Optional extras of a render launch.
struct RenderExtras {
    // in: the 8-bit frame (indexed like rgb).  When the frame gets a tile schedule the render kernel zeroes the bytes of
    // every pixel it visits, so that launch_tonemap_busy() only has to convert the busy tiles afterwards.
    unsigned char* rgb8_zero = nullptr;
    // in: events recorded on the stream before K0, between K0 and K1, after K1 (profiling; may be null)
    cudaEvent_t ev_begin = nullptr, ev_prepared = nullptr, ev_rendered = nullptr;
    // out: the frame was rendered with a tile schedule (and rgb8_zero, if given, has been zero-filled)
    bool scheduled = false;
};

// K0 + K1.  With counters == null and R = float this is the production path: prepare_raster_kernel
// (one thread per triangle, FP64) followed by render_fast_kernel; otherwise the generic kernel.
// K1: render rows [fp.row_begin, fp.row_end).  rgb: H*W*3 of R; prim_id optional; dmax: scalar R that
// receives max(old, tile max); counters: 17 x u64 (instrumented kernel) or null (production kernel).
template <typename R>
cudaError_t launch_render(const DeviceScene<R>& ds, const FrameParams<R>& fp, bool cull, R* rgb, int* prim_id, R* dmax,
                          unsigned long long* counters, cudaStream_t stream, const double camera[3] = nullptr,
                          int* launches = nullptr, RenderExtras* extras = nullptr);

// K4: FrameBuffer::normalize + to_vec (framebuffer.rs:40-82) over rows [fp.row_begin, fp.row_end).
template <typename R>
cudaError_t launch_tonemap(const FrameParams<R>& fp, const R* rgb, const R* dmax, bool normalise, unsigned char* rgb8,
                           cudaStream_t stream);

// K4 over the busy tiles of the frame `ds` rendered last with a schedule and RenderExtras::rgb8_zero (same fp).
cudaError_t launch_tonemap_busy(const DeviceScene<float>& ds, const FrameParams<float>& fp, const float* rgb, const float* dmax,
                                bool normalise, unsigned char* rgb8, cudaStream_t stream);

// Pure-FFMA probe: `iters` x 16 dependent-chain FFMAs per thread on every SM; returns flop count.
cudaError_t launch_ffma_probe(float* sink, int iters, int blocks, cudaStream_t stream, double* flops);

}  // namespace rm
