// xp_internal.h -- host-side declarations shared by xp_api.cu / xp_build.cu / xp_query.cu.
#pragma once

#include "xp_device.cuh"

namespace xp {

// RAII stage timer: with XP_OPT_TIMING=1 brackets the stage with a pair of CUDA events on the
// scene's stream (no synchronisation); resolve_stage_timers() turns the recorded pairs into
// stats.stage_ms once the stream has been synchronised at the end of the call.
struct StageTimer {
    xp_scene *s; int slot;
    StageTimer(xp_scene *scene, int st);
    ~StageTimer();
};
void resolve_stage_timers(xp_scene *s);

inline unsigned grid_for(uint64_t n, unsigned block) { return (unsigned)((n + block - 1) / block); }

// ---- xp_build.cu ----
// LSD radix sort of (key, value) pairs on `bits` key bits, 10 bits per pass, CUB-free.
// Ping-pongs between (ka,va) and (kb,vb); *in_a tells where the sorted result ended up.
// inv_out (optional): the last pass also writes the inverse permutation, inv_out[value] = position.
// first_hist_done: the first pass's per-tile digit histogram already sits in s->hist (k_morton_tris leaves it there).
cudaError_t radix_sort_pairs(xp_scene *s, uint32_t *ka, uint32_t *kb, uint32_t *va, uint32_t *vb,
                             uint64_t n, int bits, bool *in_a, uint32_t *inv_out = nullptr, bool first_hist_done = false);
cudaError_t build_lbvh(xp_scene *s);
cudaError_t emit_heightmap(xp_scene *s, const float *d_heights, unsigned nx, unsigned nz, float x0, float z0, float spacing);
cudaError_t height_range(xp_scene *s, const float *d_heights, uint64_t n);     // -> s->grid_yrange
cudaError_t height_windows(xp_scene *s, const float *d_heights, unsigned nx, unsigned nz);   // -> s->grid_mm
cudaError_t emit_clipmap(xp_scene *s, const float *d_heights, unsigned size, const int *d_parts, unsigned n_parts,
                         float unit0, uint64_t n_tris);
cudaError_t build_aux(xp_scene *s);      // Node4 / 32-wide planes, on first use by the method that needs them
// order[k] = index of the k-th sphere in Morton order of its centre (scene bounds).
// rays_done (nullable): when given, the ordering may also write the ray records + empty best keys of the sweep that
// follows into s->rays / s->best (reserved by the caller) and reports whether it did; status_for_rays as in k_ray_setup.
cudaError_t sphere_morton_order(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t **order, uint32_t **inv_order,
                                uint32_t *status_for_rays = nullptr, bool *rays_done = nullptr);

// ---- xp_query.cu ----
static constexpr int XP_MAX_PEERS = 16;
struct PeerBufs { xp_contact *p[XP_MAX_PEERS]; int n; int rot; int multicast; };   // multicast: p[0] is an NVSwitch multicast address
struct QueryOut {            // device pointers, any may be null
    xp_collision *col; uint8_t *hit; uint32_t *tri_index; uint32_t *status; xp_contact *contact;
    // fused contact exchange: when peers.n > 0 every contact record is written into each peer buffer at
    // record index slot_first + i (the buffers are mapped peer memory, NVLink P2P stores)
    PeerBufs peers; uint64_t slot_first;
    // asynchronous steps: the result kernel is launched on finalize_stream once ev_np (recorded after the
    // narrowphase and the counter snapshot) has passed, and ev_fin is recorded behind it
    cudaStream_t finalize_stream; cudaEvent_t ev_np, ev_fin; xp::DeviceCounters *ctr_snapshot;
    // asynchronous steps with copy_self >= 0: the kernel writes only peers.p[copy_self] (this rank's buffer)
    // and the copy engines push that span to the other buffers, so no SM waits on NVLink
    int copy_self_plus1;
    xp_hit *compact;         // 8 B per sphere: (time_to, tri_index or 0xFFFFFFFF)
};
// collect = false: leave the statistics on the device (no host synchronisation); the caller ends
// the batch with collect_stats().  After collecting, s->pair_overflowed tells whether the
// optimistic two-stage mode dropped candidates and the sweep must be redone with s->safe_pairs.
cudaError_t detect_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                       const QueryOut &out, bool collect = true);
cudaError_t collect_stats(xp_scene *s);
cudaError_t response_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t max_iter,
                         uint32_t horizon_mode, xp_response *d_out, uint32_t *d_iter,
                         uint32_t *d_status, xp_vec3 *d_normal);
// collision_response for a handful of spheres from and to HOST memory: one copy in, one launch, one copy out.
// *taken = false: not eligible (options, size) or the kernel gave up -- nothing was written, use the batch path.
cudaError_t response_small_host(xp_scene *s, const xp_response *in, uint64_t n, uint32_t max_iter, uint32_t horizon_mode,
                                xp_response *out, uint32_t *iterations, uint32_t *status, xp_vec3 *normal, bool *taken);
cudaError_t slide_dev(xp_scene *s, const xp_vec3 *d_pos, const xp_vec3 *d_vel, uint64_t n, uint32_t max_iter, const float *radii,
                      xp_vec3 *d_out_pos, xp_vec3 *d_out_vel, uint32_t *d_handled);
cudaError_t pairs_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                      uint64_t cap, uint32_t *d_sphere, uint32_t *d_tri, uint64_t *h_count);
cudaError_t single_detect_dev(const xp_response *h_resp, const xp_triangle *h_tri,
                              xp_collision *h_out, int *result);
cudaError_t single_respond_dev(const xp_response *h_resp, const xp_collision *h_col,
                               xp_response *h_out, int *result);
cudaError_t probe_fp32(double *tflops);

// ---- xp_debug.cu ----
cudaError_t check_arith(uint64_t n_random, uint64_t seed, unsigned long long *h_bad /* [8] */);

}  // namespace xp
