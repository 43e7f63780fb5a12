// xp_api.cu -- the C ABI declared in include/xp_physics_cuda.h (scene lifetime, host<->device
// staging, option handling).  No arithmetic of the path happens on the host: every entry point
// that computes launches kernels, and fails with XP_ECUDA when no CUDA device is usable.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cstring>
#include <new>

#include "xp_internal.h"
#include <utility>

namespace xp {

static thread_local char g_cuda_err[256] = "";

static int fail_cuda(cudaError_t e) {
    snprintf(g_cuda_err, sizeof g_cuda_err, "%s: %s", cudaGetErrorName(e), cudaGetErrorString(e));
    cudaGetLastError();   // clear the sticky-free error state
    return e == cudaErrorMemoryAllocation ? XP_ENOMEM : XP_ECUDA;
}

StageTimer::StageTimer(xp_scene *scene, int st) : s(scene), slot(-1) {
    if (!s->timing) return;
    if (!s->ev_created) {
        for (int i = 0; i < 2 * xp_scene::EV_POOL; ++i) cudaEventCreate(&s->ev[i]);
        s->ev_created = true;
    }
    if (s->ev_base + s->ev_count >= xp_scene::EV_POOL) {      // pool exhausted (long response loops): drain it
        cudaStreamSynchronize(s->stream);
        resolve_stage_timers(s);
    }
    slot = s->ev_base + s->ev_count++;
    s->ev_stage[slot] = st;
    cudaEventRecord(s->ev[2 * slot], s->stream);
}
StageTimer::~StageTimer() {
    if (slot >= 0) cudaEventRecord(s->ev[2 * slot + 1], s->stream);
}
void resolve_stage_timers(xp_scene *s) {
    for (int i = s->ev_base; i < s->ev_base + s->ev_count; ++i) {
        float ms = 0.0f;
        if (cudaEventSynchronize(s->ev[2 * i + 1]) == cudaSuccess &&
            cudaEventElapsedTime(&ms, s->ev[2 * i], s->ev[2 * i + 1]) == cudaSuccess)
            s->stats.stage_ms[s->ev_stage[i]] += ms;
    }
    s->ev_count = 0;
}

static void reset_stats(xp_scene *s) { memset(&s->stats, 0, sizeof s->stats); }

}  // namespace xp

using namespace xp;

#define XP_API_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return fail_cuda(_e); } while (0)
#define XP_BIND(s) do { cudaError_t _e = cudaSetDevice((s)->device); if (_e != cudaSuccess) return fail_cuda(_e); } while (0)

extern "C" {

int xp_device_count(int *count) {
    if (!count) return XP_EINVAL;
    int n = 0;
    cudaError_t e = cudaGetDeviceCount(&n);
    if (e != cudaSuccess) { *count = 0; return fail_cuda(e); }
    *count = n;
    return XP_OK;
}

int xp_scene_create(int device, xp_scene **out) {
    if (!out) return XP_EINVAL;
    *out = nullptr;
    int n = 0;
    XP_API_TRY(cudaGetDeviceCount(&n));
    if (device < 0 || device >= n) return XP_EINVAL;
    XP_API_TRY(cudaSetDevice(device));
    XP_API_TRY(cudaFree(0));
    xp_scene *s = new (std::nothrow) xp_scene();
    if (!s) return XP_ENOMEM;
    s->device = device;
    int n_sm = 0;
    if (cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, device) == cudaSuccess && n_sm > 0) s->n_sm = n_sm;
    if (const char *e = getenv("XP_LEAF_RUNS")) s->leaf_runs = atoi(e) ? 1 : 0;     // A/B hook: the option's initial value
    reset_stats(s);
    *out = s;
    return XP_OK;
}

void xp_scene_destroy(xp_scene *s) {
    if (!s) return;
    cudaSetDevice(s->device);
    cudaStreamSynchronize(s->stream);
    if (s->io_ready) { cudaStreamSynchronize(s->copy_in); cudaStreamSynchronize(s->copy_out); }   // batches still in flight
    for (auto &x : s->ex_slot) if (x.busy) cudaEventSynchronize(x.fin);
    DevBuf *bufs[] = {&s->tris_aos, &s->recs, &s->nodes, &s->leaf_box,
                      &s->refit_flags, &s->tree_work, &s->keys_a, &s->keys_b,
                      &s->vals_a, &s->vals_b, &s->hist, &s->bounds, &s->wide, &s->wide_lo, &s->wide_hi, &s->nodes4, &s->degen, &s->grid_h, &s->grid_yrange, &s->grid_mm, &s->recs_grid, &s->rays, &s->best, &s->active_a,
                      &s->active_b, &s->inv_order, &s->pairs, &s->counters, &s->rays2, &s->best2, &s->active_c, &s->io_backup, &s->sort_ka, &s->sort_kb, &s->io_in, &s->io_out,
                      &s->io_hit, &s->io_tri, &s->io_status, &s->io_iter, &s->io_normal, &s->io_col, &s->small_dev};
    for (DevBuf *b : bufs) b->release();
    if (s->small_pin) cudaFreeHost(s->small_pin);
    for (auto &a : s->async_slot) {
        DevBuf *ab[] = {&a.in, &a.col, &a.hit, &a.tri, &a.status, &a.compact};
        for (DevBuf *b : ab) b->release();
        if (a.h2d) cudaEventDestroy(a.h2d);
        if (a.done) cudaEventDestroy(a.done);
        if (a.d2h) cudaEventDestroy(a.d2h);
        if (a.h_ctr) cudaFreeHost(a.h_ctr);
    }
    for (auto &x : s->ex_slot) {
        DevBuf *xb[] = {&x.rays, &x.best, &x.inv_order, &x.status};
        for (DevBuf *b : xb) b->release();
        if (x.np) cudaEventDestroy(x.np);
        if (x.fin) cudaEventDestroy(x.fin);
        if (x.h_ctr) cudaFreeHost(x.h_ctr);
    }
    if (s->ev_created)
        for (int i = 0; i < 2 * xp_scene::EV_POOL; ++i) cudaEventDestroy(s->ev[i]);
    if (s->io_ready) {
        cudaStreamDestroy(s->copy_in);
        cudaStreamDestroy(s->copy_out);
        for (int i = 0; i < 32; ++i) cudaEventDestroy(s->io_ev[i]);
    }
    delete s;
}

int xp_scene_set_stream(xp_scene *s, void *cuda_stream) {
    if (!s) return XP_EINVAL;
    s->stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return XP_OK;
}

int xp_scene_set_option(xp_scene *s, int option, int64_t value) {
    if (!s) return XP_EINVAL;
    switch (option) {
    case XP_OPT_ARITH:
        if (value < XP_ARITH_EXACT || value > XP_ARITH_FAST_LITERAL) return XP_EINVAL;
        s->arith = (int)value; return XP_OK;
    case XP_OPT_METHOD:
        if (value < XP_METHOD_WIDE_FUSED || value > XP_METHOD_Q4_PAIRS) return XP_EINVAL;
        s->method = (int)value; return XP_OK;
    case XP_OPT_TIMING: s->timing = value ? 1 : 0; return XP_OK;
    case XP_OPT_SORT_SPHERES: s->sort_spheres = value ? 1 : 0; return XP_OK;
    case XP_OPT_MAX_PAIRS:
        if (value < 4096) return XP_EINVAL;
        s->max_pairs = (uint64_t)value; return XP_OK;
    case XP_OPT_PRUNE:
        if (value < 0 || value > 2) return XP_EINVAL;
        s->prune = (int)value; s->cand_per_ray = 0.0; return XP_OK;
    case XP_OPT_FAR_EXACT: s->far_exact = value ? 1 : 0; return XP_OK;
    case XP_OPT_GRID_WALK: s->use_grid = value ? 1 : 0; return XP_OK;
    case XP_OPT_LEAF_RUNS: s->leaf_runs = value ? 1 : 0; return XP_OK;
    case XP_OPT_CONTACT_RECORD:
        if (value > 1) return XP_EINVAL;
        s->contact_record = (int)value; return XP_OK;
    default: return XP_EINVAL;
    }
}

int xp_scene_get_stats(const xp_scene *s, xp_stats *out) {
    if (!s || !out) return XP_EINVAL;
    *out = s->stats;
    return XP_OK;
}

uint64_t xp_scene_triangle_count(const xp_scene *s) { return s ? s->n_tris : 0; }

int xp_scene_set_triangles(xp_scene *s, const xp_triangle *aos, uint64_t n) {
    if (!s || (!aos && n)) return XP_EINVAL;
    XP_BIND(s);
    s->built = false;
    s->grid_valid = false;
    s->n_tris = 0;
    if (n) {
        XP_API_TRY(s->tris_aos.reserve(n * sizeof(xp_triangle)));
        XP_API_TRY(cudaMemcpyAsync(s->tris_aos.p, aos, n * sizeof(xp_triangle), cudaMemcpyHostToDevice, s->stream));
        XP_API_TRY(cudaStreamSynchronize(s->stream));
    }
    s->n_tris = n;
    return XP_OK;
}

int xp_scene_set_triangles_dev(xp_scene *s, const xp_triangle *d_aos, uint64_t n) {
    if (!s || (!d_aos && n)) return XP_EINVAL;
    XP_BIND(s);
    s->built = false;
    s->grid_valid = false;
    s->n_tris = 0;
    if (n) {
        XP_API_TRY(s->tris_aos.reserve(n * sizeof(xp_triangle)));
        XP_API_TRY(cudaMemcpyAsync(s->tris_aos.p, d_aos, n * sizeof(xp_triangle), cudaMemcpyDeviceToDevice, s->stream));
    }
    s->n_tris = n;
    return XP_OK;
}

int xp_scene_set_heightmap(xp_scene *s, const float *heights, uint32_t nx, uint32_t nz, float x0, float z0,
                           float spacing) {
    if (!s || !heights || nx == 0 || nz == 0) return XP_EINVAL;
    const uint64_t n = 2ull * nx * nz;
    if (n >= (1ull << 27)) return XP_EINVAL;
    XP_BIND(s);
    s->built = false;
    s->n_tris = 0;
    const uint64_t hcount = ((uint64_t)nx + 1) * ((uint64_t)nz + 1), hbytes = hcount * sizeof(float);
    s->grid_valid = false;
    XP_API_TRY(s->grid_h.reserve(hbytes));               // kept: the height grid doubles as the broadphase structure
    XP_API_TRY(cudaMemcpyAsync(s->grid_h.p, heights, hbytes, cudaMemcpyHostToDevice, s->stream));
    XP_API_TRY(emit_heightmap(s, s->grid_h.as<float>(), nx, nz, x0, z0, spacing));
    XP_API_TRY(height_range(s, s->grid_h.as<float>(), hcount));
    XP_API_TRY(height_windows(s, s->grid_h.as<float>(), nx, nz));
    XP_API_TRY(cudaStreamSynchronize(s->stream));
    s->n_tris = n;
    // The grid walk (xp_gridwalk.cuh) needs a regular grid: finite origin, positive spacing, grid lines that really
    // advance in fp32, and coordinates small enough that the rounding of an enumerated position (~32 ulp of the
    // largest coordinate, the slack `eps`) stays a fraction of a cell.  Anything else simply keeps the LBVH walk.
    const double xmax = fabs((double)x0) + (double)spacing * nx, zmax = fabs((double)z0) + (double)spacing * nz;
    const double big = (xmax > zmax ? xmax : zmax) + 2.0;
    const double err = 32.0 * 1.1920929e-7 * big;
    if (spacing > 0.0f && spacing <= 3.0e38f && x0 == x0 && z0 == z0 && big < 1.0e30 && err < 0.25 * spacing) {
        s->grid_nx = nx; s->grid_nz = nz; s->grid_x0 = x0; s->grid_z0 = z0; s->grid_sp = spacing;
        s->grid_eps = (float)err;
        s->grid_valid = true;
    }
    return XP_OK;
}

int xp_scene_set_clipmap(xp_scene *s, const float *heights, uint32_t levels, uint32_t size, const int32_t *parts,
                         uint32_t n_parts, float unit0) {
    if (!s || !heights || !parts || levels == 0 || levels > 24 || size == 0 || n_parts == 0 || n_parts > 4096) return XP_EINVAL;
    uint64_t n = 0;
    for (uint32_t i = 0; i < n_parts; ++i) {             // rows: level, x0, z0, size_x, size_z, first, count, 0
        const int32_t *r = parts + 8 * (size_t)i;
        if (r[0] < 0 || (uint32_t)r[0] >= levels || r[3] < 2 || r[4] < 2) return XP_EINVAL;
        if ((uint64_t)r[5] != n || (int64_t)r[6] != 2ll * (r[3] - 1) * (r[4] - 1)) return XP_EINVAL;
        n += (uint64_t)r[6];
    }
    if (n == 0 || n >= (1ull << 27)) return XP_EINVAL;
    XP_BIND(s);
    s->built = false;
    s->grid_valid = false;
    s->n_tris = 0;
    const uint64_t hbytes = (uint64_t)levels * size * size * sizeof(float), pbytes = (uint64_t)n_parts * 8 * sizeof(int32_t);
    XP_API_TRY(s->io_in.reserve(hbytes + pbytes));
    char *d = static_cast<char *>(s->io_in.p);
    XP_API_TRY(cudaMemcpyAsync(d, heights, hbytes, cudaMemcpyHostToDevice, s->stream));
    XP_API_TRY(cudaMemcpyAsync(d + hbytes, parts, pbytes, cudaMemcpyHostToDevice, s->stream));
    XP_API_TRY(emit_clipmap(s, reinterpret_cast<const float *>(d), size, reinterpret_cast<const int *>(d + hbytes), n_parts, unit0, n));
    XP_API_TRY(cudaStreamSynchronize(s->stream));
    s->n_tris = n;
    return XP_OK;
}

int xp_scene_set_positions(xp_scene *s, const xp_vec3 *soup, uint64_t n_vec3) {
    if (n_vec3 % 3 != 0) return XP_EINVAL;     // lib.rs:19-24: chunks(3) then vs[2] panics
    // Vec3 triples are exactly the Triangle layout (lib.rs:20-24), so this is one upload
    return xp_scene_set_triangles(s, reinterpret_cast<const xp_triangle *>(soup), n_vec3 / 3);
}

int xp_scene_build(xp_scene *s) {
    if (!s) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    s->n_degen = 0;
    XP_API_TRY(build_lbvh(s));
    uint32_t nd = 0;
    if (s->n_tris > 0) XP_API_TRY(cudaMemcpyAsync(&nd, s->degen.p, sizeof nd, cudaMemcpyDeviceToHost, s->stream));
    XP_API_TRY(cudaStreamSynchronize(s->stream));
    s->n_degen = nd;
    resolve_stage_timers(s);
    return XP_OK;
}

// One sweep in the optimistic (sync-free) pair mode; if a chunk overflowed the pair buffer the
// sweep is redone in the synchronous, chunk-halving mode.  Inputs are read-only, so a rerun is safe.
static cudaError_t detect_retry(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                const QueryOut &q) {
    cudaError_t e = detect_dev(s, d_in, n, horizon_mode, q);
    if (e == cudaSuccess && s->pair_overflowed) {
        s->pair_overflowed = false;
        s->safe_pairs = true;
        const uint32_t launches = s->stats.kernel_launches;
        memset(&s->stats, 0, sizeof s->stats);
        s->stats.kernel_launches = launches;
        e = detect_dev(s, d_in, n, horizon_mode, q);
        s->safe_pairs = false;
        s->pair_overflowed = false;
    }
    return e;
}

static int check_query(xp_scene *s, const void *in, uint64_t n) {
    if (!s || (!in && n)) return XP_EINVAL;
    if (!s->built) return XP_ESTATE;
    return XP_OK;
}

int xp_detect_batch_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                        xp_collision *d_out, uint8_t *d_hit, uint32_t *d_tri_index, uint32_t *d_status) {
    int rc = check_query(s, d_in, n);
    if (rc) return rc;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    QueryOut q = {d_out, d_hit, d_tri_index, d_status, nullptr, {}, 0};
    XP_API_TRY(detect_retry(s, d_in, n, horizon_mode, q));
    return XP_OK;
}

int xp_detect_contacts_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                           xp_contact *d_out) {
    int rc = check_query(s, d_in, n);
    if (rc) return rc;
    if (!d_out && n) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    XP_API_TRY(s->io_status.reserve(n * 4));
    QueryOut q = {nullptr, nullptr, nullptr, s->io_status.as<uint32_t>(), d_out, {}, 0};
    XP_API_TRY(detect_retry(s, d_in, n, horizon_mode, q));
    return XP_OK;
}

// Host-pointer sweep, software-pipelined over chunks of the batch: chunk i+1's H2D copy and chunk
// i-1's D2H copies run on their own streams (both DMA engines) while chunk i computes.  With
// pinned caller buffers the copies are truly asynchronous; with pageable memory the runtime stages
// them and the overlap is partial.
static constexpr int IO_MAX_CHUNKS = 16;

static int ensure_io_streams(xp_scene *s) {
    if (s->io_ready) return XP_OK;
    XP_API_TRY(cudaStreamCreateWithFlags(&s->copy_in, cudaStreamNonBlocking));
    XP_API_TRY(cudaStreamCreateWithFlags(&s->copy_out, cudaStreamNonBlocking));
    for (int i = 0; i < 2 * IO_MAX_CHUNKS; ++i) XP_API_TRY(cudaEventCreateWithFlags(&s->io_ev[i], cudaEventDisableTiming));
    s->io_ready = true;
    return XP_OK;
}

// ---- multi-GPU contact exchange over peer memory -------------------------------------------------
int xp_exchange_alloc(xp_scene *s, uint64_t bytes, void **dev_ptr, uint8_t ipc_handle[64]) {
    if (!s || !dev_ptr || !ipc_handle || bytes == 0) return XP_EINVAL;
    XP_BIND(s);
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "IPC handle size");
    void *p = nullptr;
    XP_API_TRY(cudaMalloc(&p, bytes));
    cudaIpcMemHandle_t h;
    cudaError_t e = cudaIpcGetMemHandle(&h, p);
    if (e != cudaSuccess) { cudaFree(p); return fail_cuda(e); }
    memcpy(ipc_handle, &h, 64);
    *dev_ptr = p;
    return XP_OK;
}

int xp_exchange_free(xp_scene *s, void *dev_ptr) {
    if (!s) return XP_EINVAL;
    XP_BIND(s);
    if (dev_ptr) XP_API_TRY(cudaFree(dev_ptr));
    return XP_OK;
}

int xp_exchange_open(xp_scene *s, const uint8_t ipc_handle[64], void **peer_ptr) {
    if (!s || !ipc_handle || !peer_ptr) return XP_EINVAL;
    XP_BIND(s);
    cudaIpcMemHandle_t h;
    memcpy(&h, ipc_handle, 64);
    XP_API_TRY(cudaIpcOpenMemHandle(peer_ptr, h, cudaIpcMemLazyEnablePeerAccess));
    return XP_OK;
}

int xp_exchange_close(xp_scene *s, void *peer_ptr) {
    if (!s) return XP_EINVAL;
    XP_BIND(s);
    if (peer_ptr) XP_API_TRY(cudaIpcCloseMemHandle(peer_ptr));
    return XP_OK;
}

int xp_detect_contacts_exchange_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                    void *const *bufs, int n_bufs, uint64_t slot_first) {
    int rc = check_query(s, d_in, n);
    if (rc) return rc;
    if (!bufs || n_bufs <= 0 || n_bufs > XP_MAX_PEERS || (slot_first & 1)) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    XP_API_TRY(s->io_status.reserve(n * 4));
    QueryOut q = {nullptr, nullptr, nullptr, s->io_status.as<uint32_t>(), nullptr, {}, slot_first};
    q.peers.n = n_bufs;
    q.peers.rot = (int)((slot_first / (n ? n : 1)) % (uint64_t)n_bufs) + 1;     // ~ this rank's position + 1
    for (int i = 0; i < n_bufs; ++i) {
        if (!bufs[i]) return XP_EINVAL;
        q.peers.p[i] = static_cast<xp_contact *>(bufs[i]);
    }
    XP_API_TRY(detect_retry(s, d_in, n, horizon_mode, q));
    return XP_OK;
}

int xp_detect_contacts_multicast_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                     void *multicast_buf, uint64_t slot_first) {
    int rc = check_query(s, d_in, n);
    if (rc) return rc;
    if (!multicast_buf || (slot_first & 1) || (reinterpret_cast<uintptr_t>(multicast_buf) & 15u)) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    XP_API_TRY(s->io_status.reserve(n * 4));
    QueryOut q = {nullptr, nullptr, nullptr, s->io_status.as<uint32_t>(), nullptr, {}, slot_first};
    q.peers.n = 1;
    q.peers.multicast = 1;
    q.peers.p[0] = static_cast<xp_contact *>(multicast_buf);
    XP_API_TRY(detect_retry(s, d_in, n, horizon_mode, q));
    return XP_OK;
}

int xp_detect_batch(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                    xp_collision *out, uint8_t *hit, uint32_t *tri_index, uint32_t *status) {
    int rc = check_query(s, in, n);
    if (rc) return rc;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    rc = ensure_io_streams(s);
    if (rc) return rc;
    cudaStream_t st = s->stream;
    XP_API_TRY(s->io_in.reserve(n * sizeof(xp_response)));
    XP_API_TRY(s->io_col.reserve(n * sizeof(xp_collision)));
    XP_API_TRY(s->io_hit.reserve(n));
    XP_API_TRY(s->io_tri.reserve(n * 4));
    XP_API_TRY(s->io_status.reserve(n * 4));
    // Measured on config 2: a sweep of 2^20 spheres is 1.45x / 2.0x / 2.9x slower in total when cut into
    // 2 / 4 / 8 chunks (the persistent broadphase needs several runs of 32 spheres per warp to hide
    // latency), which costs more than the overlapped copies save.  So chunks hold >= 2^21 spheres:
    // batches up to that size go through in one piece, larger ones are pipelined.
    uint64_t chunk = 1ull << 21;
    if (const char *e = getenv("XP_IO_CHUNKS")) {      // experiment knob: force a chunk count
        const uint64_t k = (uint64_t)atoi(e) > 0 ? (uint64_t)atoi(e) : 1;
        chunk = (n + k - 1) / k;
    }
    if ((n + chunk - 1) / chunk > IO_MAX_CHUNKS) chunk = (n + IO_MAX_CHUNKS - 1) / IO_MAX_CHUNKS;
    const int n_chunks = (int)((n + chunk - 1) / chunk);
    xp_response *d_in = s->io_in.as<xp_response>();
    // all H2D copies are queued up front on the copy-in stream, one event per chunk
    for (int c = 0; c < n_chunks; ++c) {
        const uint64_t off = (uint64_t)c * chunk, m = (n - off < chunk) ? n - off : chunk;
        XP_API_TRY(cudaMemcpyAsync(d_in + off, in + off, m * sizeof(xp_response), cudaMemcpyHostToDevice, s->copy_in));
        XP_API_TRY(cudaEventRecord(s->io_ev[c], s->copy_in));
    }
    const xp_stats zero_stats = s->stats;
    (void)zero_stats;
    for (int c = 0; c < n_chunks; ++c) {
        const uint64_t off = (uint64_t)c * chunk, m = (n - off < chunk) ? n - off : chunk;
        XP_API_TRY(cudaStreamWaitEvent(st, s->io_ev[c], 0));
        if (out) XP_API_TRY(cudaMemsetAsync(s->io_col.as<xp_collision>() + off, 0, m * sizeof(xp_collision), st));
        QueryOut q = {out ? s->io_col.as<xp_collision>() + off : nullptr, s->io_hit.as<uint8_t>() + off,
                      tri_index ? s->io_tri.as<uint32_t>() + off : nullptr, s->io_status.as<uint32_t>() + off, nullptr, {}, 0};
        XP_API_TRY(detect_dev(s, d_in + off, m, horizon_mode, q, /*collect=*/false));   // no host sync
        XP_API_TRY(cudaEventRecord(s->io_ev[IO_MAX_CHUNKS + c], st));
        XP_API_TRY(cudaStreamWaitEvent(s->copy_out, s->io_ev[IO_MAX_CHUNKS + c], 0));
        cudaStream_t so = s->copy_out;
        if (out) XP_API_TRY(cudaMemcpyAsync(out + off, s->io_col.as<xp_collision>() + off, m * sizeof(xp_collision), cudaMemcpyDeviceToHost, so));
        if (hit) XP_API_TRY(cudaMemcpyAsync(hit + off, s->io_hit.as<uint8_t>() + off, m, cudaMemcpyDeviceToHost, so));
        if (tri_index) XP_API_TRY(cudaMemcpyAsync(tri_index + off, s->io_tri.as<uint32_t>() + off, m * 4, cudaMemcpyDeviceToHost, so));
        if (status) XP_API_TRY(cudaMemcpyAsync(status + off, s->io_status.as<uint32_t>() + off, m * 4, cudaMemcpyDeviceToHost, so));
    }
    XP_API_TRY(cudaStreamSynchronize(s->copy_out));
    XP_API_TRY(collect_stats(s));                   // the one synchronisation of the compute stream
    s->counters_live = false;
    if (s->pair_overflowed) {                       // rare: redo everything in the synchronous pair mode
        s->pair_overflowed = false;
        s->safe_pairs = true;
        if (out) XP_API_TRY(cudaMemsetAsync(s->io_col.p, 0, n * sizeof(xp_collision), st));
        QueryOut q = {out ? s->io_col.as<xp_collision>() : nullptr, s->io_hit.as<uint8_t>(),
                      tri_index ? s->io_tri.as<uint32_t>() : nullptr, s->io_status.as<uint32_t>(), nullptr, {}, 0};
        cudaError_t e = detect_dev(s, d_in, n, horizon_mode, q);
        s->safe_pairs = false;
        s->pair_overflowed = false;
        if (e != cudaSuccess) return fail_cuda(e);
        if (out) XP_API_TRY(cudaMemcpyAsync(out, s->io_col.p, n * sizeof(xp_collision), cudaMemcpyDeviceToHost, st));
        if (hit) XP_API_TRY(cudaMemcpyAsync(hit, s->io_hit.p, n, cudaMemcpyDeviceToHost, st));
        if (tri_index) XP_API_TRY(cudaMemcpyAsync(tri_index, s->io_tri.p, n * 4, cudaMemcpyDeviceToHost, st));
        if (status) XP_API_TRY(cudaMemcpyAsync(status, s->io_status.p, n * 4, cudaMemcpyDeviceToHost, st));
        XP_API_TRY(cudaStreamSynchronize(st));
    }
    return XP_OK;
}


// ---- asynchronous host-pointer sweep ----------------------------------------------------------
// submit() queues H2D copy -> sweep -> D2H copies of one batch on three streams and returns; wait()
// blocks until that batch's results are in the caller's buffers.  With XP_ASYNC_SLOTS batches in
// flight the copies of one batch run under the kernels of another, so a stream of batches moves at
// the pace of the slower of compute and PCIe instead of their sum.  The optimistic pair buffer is
// checked from a counter snapshot copied back with the results; a batch that dropped candidates is
// redone synchronously inside wait() (inputs are read-only, so that is safe).
static int ensure_async(xp_scene *s) {
    int rc = ensure_io_streams(s);
    if (rc) return rc;
    if (s->async_ready) return XP_OK;
    for (auto &a : s->async_slot) {
        XP_API_TRY(cudaEventCreateWithFlags(&a.h2d, cudaEventDisableTiming));
        XP_API_TRY(cudaEventCreateWithFlags(&a.done, cudaEventDisableTiming));
        XP_API_TRY(cudaEventCreateWithFlags(&a.d2h, cudaEventDisableTiming));
        XP_API_TRY(cudaHostAlloc(reinterpret_cast<void **>(&a.h_ctr), sizeof(DeviceCounters), cudaHostAllocDefault));
    }
    s->async_ready = true;
    return XP_OK;
}

static int submit_impl(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode, xp_collision *out,
                       uint8_t *hit, uint32_t *tri_index, uint32_t *status, xp_hit *compact, int *ticket) {
    int rc = check_query(s, in, n);
    if (rc) return rc;
    if (!ticket) return XP_EINVAL;
    XP_BIND(s);
    rc = ensure_async(s);
    if (rc) return rc;
    int slot = -1;
    for (int i = 0; i < xp_scene::ASYNC_SLOTS; ++i)
        if (!s->async_slot[i].busy) { slot = i; break; }
    if (slot < 0) return XP_ESTATE;                  // every slot is in flight: wait for a ticket first
    xp_scene::AsyncSlot &a = s->async_slot[slot];
    a.h_in = in; a.n = n; a.horizon_mode = horizon_mode;
    a.out = out; a.hit_out = hit; a.tri_out = tri_index; a.status_out = status; a.compact_out = compact;
    a.launches = 0;
    *ticket = slot;
    if (n == 0) { memset(a.h_ctr, 0, sizeof(DeviceCounters)); a.busy = true; return XP_OK; }
    cudaStream_t st = s->stream;
    XP_API_TRY(a.in.reserve(n * sizeof(xp_response)));
    if (out) XP_API_TRY(a.col.reserve(n * sizeof(xp_collision)));
    if (hit) XP_API_TRY(a.hit.reserve(n));
    if (tri_index) XP_API_TRY(a.tri.reserve(n * 4));
    if (compact) XP_API_TRY(a.compact.reserve(n * sizeof(xp_hit)));
    XP_API_TRY(a.status.reserve(n * 4));
    XP_API_TRY(cudaMemcpyAsync(a.in.p, in, n * sizeof(xp_response), cudaMemcpyHostToDevice, s->copy_in));
    XP_API_TRY(cudaEventRecord(a.h2d, s->copy_in));
    XP_API_TRY(cudaStreamWaitEvent(st, a.h2d, 0));
    if (out) XP_API_TRY(cudaMemsetAsync(a.col.p, 0, n * sizeof(xp_collision), st));
    QueryOut q = {out ? a.col.as<xp_collision>() : nullptr, hit ? a.hit.as<uint8_t>() : nullptr, tri_index ? a.tri.as<uint32_t>() : nullptr,
                  a.status.as<uint32_t>(), nullptr, {}, 0};
    q.compact = compact ? a.compact.as<xp_hit>() : nullptr;
    const int timing = s->timing;
    s->timing = 0;                                   // stage events are resolved by synchronous calls only
    s->counters_live = false;                        // this batch gets freshly zeroed counters
    const uint32_t before = s->stats.kernel_launches;
    cudaError_t e = detect_dev(s, a.in.as<xp_response>(), n, horizon_mode, q, /*collect=*/false);
    s->timing = timing;
    s->counters_live = false;
    a.launches = s->stats.kernel_launches - before;
    if (e != cudaSuccess) return fail_cuda(e);
    XP_API_TRY(cudaMemcpyAsync(a.h_ctr, s->counters.p, sizeof(DeviceCounters), cudaMemcpyDeviceToHost, st));
    XP_API_TRY(cudaEventRecord(a.done, st));
    cudaStream_t so = s->copy_out;
    XP_API_TRY(cudaStreamWaitEvent(so, a.done, 0));
    if (out) XP_API_TRY(cudaMemcpyAsync(out, a.col.p, n * sizeof(xp_collision), cudaMemcpyDeviceToHost, so));
    if (hit) XP_API_TRY(cudaMemcpyAsync(hit, a.hit.p, n, cudaMemcpyDeviceToHost, so));
    if (tri_index) XP_API_TRY(cudaMemcpyAsync(tri_index, a.tri.p, n * 4, cudaMemcpyDeviceToHost, so));
    if (status) XP_API_TRY(cudaMemcpyAsync(status, a.status.p, n * 4, cudaMemcpyDeviceToHost, so));
    if (compact) XP_API_TRY(cudaMemcpyAsync(compact, a.compact.p, n * sizeof(xp_hit), cudaMemcpyDeviceToHost, so));
    XP_API_TRY(cudaEventRecord(a.d2h, so));
    a.busy = true;                                   // only now: an error return above leaves the slot free
    return XP_OK;
}

int xp_detect_batch_submit(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                           xp_collision *out, uint8_t *hit, uint32_t *tri_index, uint32_t *status, int *ticket) {
    return submit_impl(s, in, n, horizon_mode, out, hit, tri_index, status, nullptr, ticket);
}

int xp_detect_batch_compact_submit(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                                   xp_hit *out, uint32_t *status, int *ticket) {
    if (!out && n) return XP_EINVAL;
    return submit_impl(s, in, n, horizon_mode, nullptr, nullptr, nullptr, status, out, ticket);
}

int xp_detect_batch_compact_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                xp_hit *d_out, uint32_t *d_status) {
    int rc = check_query(s, d_in, n);
    if (rc) return rc;
    if (!d_out && n) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    QueryOut q = {nullptr, nullptr, nullptr, d_status, nullptr, {}, 0};
    q.compact = d_out;
    XP_API_TRY(detect_retry(s, d_in, n, horizon_mode, q));
    return XP_OK;
}

int xp_detect_batch_compact(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode,
                            xp_hit *out, uint32_t *status) {
    int rc = check_query(s, in, n);
    if (rc) return rc;
    if (!out && n) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    cudaStream_t st = s->stream;
    XP_API_TRY(s->io_in.reserve(n * sizeof(xp_response)));
    XP_API_TRY(s->io_col.reserve(n * sizeof(xp_hit)));
    XP_API_TRY(s->io_status.reserve(n * 4));
    XP_API_TRY(cudaMemcpyAsync(s->io_in.p, in, n * sizeof(xp_response), cudaMemcpyHostToDevice, st));
    QueryOut q = {nullptr, nullptr, nullptr, s->io_status.as<uint32_t>(), nullptr, {}, 0};
    q.compact = s->io_col.as<xp_hit>();
    XP_API_TRY(detect_retry(s, s->io_in.as<xp_response>(), n, horizon_mode, q));
    XP_API_TRY(cudaMemcpyAsync(out, s->io_col.p, n * sizeof(xp_hit), cudaMemcpyDeviceToHost, st));
    if (status) XP_API_TRY(cudaMemcpyAsync(status, s->io_status.p, n * 4, cudaMemcpyDeviceToHost, st));
    XP_API_TRY(cudaStreamSynchronize(st));
    return XP_OK;
}

int xp_detect_batch_wait(xp_scene *s, int ticket) {
    if (!s || ticket < 0 || ticket >= xp_scene::ASYNC_SLOTS) return XP_EINVAL;
    xp_scene::AsyncSlot &a = s->async_slot[ticket];
    if (!a.busy) return XP_ESTATE;
    XP_BIND(s);
    a.busy = false;
    reset_stats(s);
    if (a.n == 0) return XP_OK;
    XP_API_TRY(cudaEventSynchronize(a.d2h));
    const DeviceCounters &h = *a.h_ctr;
    if (h.stack_overflow) return fail_cuda(cudaErrorAssert);
    if (h.pair_overflow) {                           // rare: candidates were dropped -- redo this batch synchronously
        if (a.compact_out) return xp_detect_batch_compact(s, a.h_in, a.n, a.horizon_mode, a.compact_out, a.status_out);
        return xp_detect_batch(s, a.h_in, a.n, a.horizon_mode, a.out, a.hit_out, a.tri_out, a.status_out);
    }
    s->stats.narrowphase_tests = h.tests;
    s->stats.broadphase_pairs = h.pairs_total;
    s->stats.node_visits = h.node_visits;
    s->stats.kernel_launches = a.launches;
    return XP_OK;
}

// ---- asynchronous fused exchange: step i's record build + all-gather under step i+1's sweep ----------
int xp_scene_set_exchange_stream(xp_scene *s, void *cuda_stream) {
    if (!s) return XP_EINVAL;
    s->exchange_stream = reinterpret_cast<cudaStream_t>(cuda_stream);
    return XP_OK;
}

static int ensure_exchange_slots(xp_scene *s) {
    if (s->ex_ready) return XP_OK;
    for (auto &x : s->ex_slot) {
        XP_API_TRY(cudaEventCreateWithFlags(&x.np, cudaEventDisableTiming));
        XP_API_TRY(cudaEventCreateWithFlags(&x.fin, cudaEventDisableTiming));
        XP_API_TRY(cudaHostAlloc(reinterpret_cast<void **>(&x.h_ctr), sizeof(DeviceCounters), cudaHostAllocDefault));
    }
    s->ex_ready = true;
    return XP_OK;
}

int xp_detect_contacts_exchange_submit(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t horizon_mode,
                                       void *const *bufs, int n_bufs, int self_index, uint64_t slot_first, int *ticket) {
    int rc = check_query(s, d_in, n);
    if (rc) return rc;
    if (!bufs || !ticket || n_bufs <= 0 || n_bufs > XP_MAX_PEERS || (slot_first & 1)) return XP_EINVAL;
    if (self_index < -1 || self_index >= n_bufs) return XP_EINVAL;
    if (n == 0 || n > (1ull << 24)) return XP_EINVAL;            // one chunk per step: its buffers stay untouched until the exchange has read them
    if (!s->exchange_stream || s->exchange_stream == s->stream) return XP_ESTATE;
    XP_BIND(s);
    rc = ensure_exchange_slots(s);
    if (rc) return rc;
    const int t = s->ex_next;
    xp_scene::ExSlot &x = s->ex_slot[t];
    if (x.busy) return XP_ESTATE;                                // wait for this ticket first
    QueryOut q = {nullptr, nullptr, nullptr, nullptr, nullptr, {}, slot_first, s->exchange_stream, x.np, x.fin, x.h_ctr, self_index + 1};
    q.peers.n = n_bufs;
    q.peers.rot = (int)((slot_first / (n ? n : 1)) % (uint64_t)n_bufs) + 1;
    for (int i = 0; i < n_bufs; ++i) {
        if (!bufs[i]) return XP_EINVAL;
        q.peers.p[i] = static_cast<xp_contact *>(bufs[i]);
    }
    // this step works in its slot's buffers: the previous user of the slot (two steps ago) must have been read
    if (x.used) XP_API_TRY(cudaStreamWaitEvent(s->stream, x.fin, 0));
    struct Swap {
        xp_scene *s; xp_scene::ExSlot &x;
        void flip() { std::swap(s->rays, x.rays); std::swap(s->best, x.best); std::swap(s->inv_order, x.inv_order); std::swap(s->io_status, x.status); }
        Swap(xp_scene *s_, xp_scene::ExSlot &x_) : s(s_), x(x_) { flip(); }
        ~Swap() { flip(); }
    } swap(s, x);
    cudaError_t e = s->io_status.reserve(n * 4);
    if (e != cudaSuccess) return fail_cuda(e);
    q.status = s->io_status.as<uint32_t>();
    reset_stats(s);
    s->ev_base = t * (xp_scene::EV_POOL / 2);
    s->ev_count = 0;
    s->counters_live = false;                                    // freshly zeroed counters for this step
    e = detect_dev(s, d_in, n, horizon_mode, q, /*collect=*/false);
    s->counters_live = false;
    x.ev_count = s->ev_count;
    x.launches = s->stats.kernel_launches;
    s->ev_base = 0;
    s->ev_count = 0;
    if (e != cudaSuccess) return fail_cuda(e);
    x.busy = x.used = true;
    s->ex_next = t ^ 1;
    *ticket = t;
    return XP_OK;
}

int xp_detect_contacts_exchange_wait(xp_scene *s, int ticket) {
    if (!s || ticket < 0 || ticket > 1) return XP_EINVAL;
    xp_scene::ExSlot &x = s->ex_slot[ticket];
    if (!x.busy) return XP_ESTATE;
    XP_BIND(s);
    x.busy = false;
    XP_API_TRY(cudaEventSynchronize(x.fin));
    reset_stats(s);
    const DeviceCounters &h = *x.h_ctr;
    s->stats.narrowphase_tests = h.tests;
    s->stats.broadphase_pairs = h.pairs_total;
    s->stats.node_visits = h.node_visits;
    s->stats.kernel_launches = x.launches;
    s->ev_base = ticket * (xp_scene::EV_POOL / 2);
    s->ev_count = x.ev_count;
    resolve_stage_timers(s);
    s->ev_base = 0;
    if (h.stack_overflow) return fail_cuda(cudaErrorAssert);
    if (h.pair_overflow) return XP_ECAP;                         // candidates were dropped: redo the step with the synchronous call
    return XP_OK;
}

int xp_collision_response_batch_dev(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t max_iter,
                                    uint32_t horizon_mode, xp_response *d_out, uint32_t *d_iterations,
                                    uint32_t *d_status, xp_vec3 *d_last_slide_normal) {
    int rc = check_query(s, d_in, n);
    if (rc) return rc;
    if (!d_out && n) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    XP_API_TRY(response_dev(s, d_in, n, max_iter, horizon_mode, d_out, d_iterations, d_status, d_last_slide_normal));
    return XP_OK;
}

int xp_collision_response_batch(xp_scene *s, const xp_response *in, uint64_t n, uint32_t max_iter,
                                uint32_t horizon_mode, xp_response *out, uint32_t *iterations,
                                uint32_t *status, xp_vec3 *last_slide_normal) {
    int rc = check_query(s, in, n);
    if (rc) return rc;
    if (!out && n) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    {
        bool taken = false;
        XP_API_TRY(response_small_host(s, in, n, max_iter, horizon_mode, out, iterations, status, last_slide_normal, &taken));
        if (taken) return XP_OK;
    }
    cudaStream_t st = s->stream;
    XP_API_TRY(s->io_out.reserve(n * sizeof(xp_response)));
    XP_API_TRY(s->io_iter.reserve(n * 4));
    XP_API_TRY(s->io_status.reserve(n * 4));
    XP_API_TRY(s->io_normal.reserve(n * sizeof(xp_vec3)));
    // the current Response of every sphere lives in io_out and is updated in place
    XP_API_TRY(cudaMemcpyAsync(s->io_out.p, in, n * sizeof(xp_response), cudaMemcpyHostToDevice, st));
    XP_API_TRY(response_dev(s, s->io_out.as<xp_response>(), n, max_iter, horizon_mode, s->io_out.as<xp_response>(),
                            s->io_iter.as<uint32_t>(), s->io_status.as<uint32_t>(),
                            last_slide_normal ? s->io_normal.as<xp_vec3>() : nullptr));
    XP_API_TRY(cudaMemcpyAsync(out, s->io_out.p, n * sizeof(xp_response), cudaMemcpyDeviceToHost, st));
    if (iterations) XP_API_TRY(cudaMemcpyAsync(iterations, s->io_iter.p, n * 4, cudaMemcpyDeviceToHost, st));
    if (status) XP_API_TRY(cudaMemcpyAsync(status, s->io_status.p, n * 4, cudaMemcpyDeviceToHost, st));
    if (last_slide_normal)
        XP_API_TRY(cudaMemcpyAsync(last_slide_normal, s->io_normal.p, n * sizeof(xp_vec3), cudaMemcpyDeviceToHost, st));
    XP_API_TRY(cudaStreamSynchronize(st));
    return XP_OK;
}

static int check_slide(xp_scene *s, const void *a, const void *b, const void *c, const void *d_, uint64_t n, const float *radii) {
    int rc = check_query(s, a, n);
    if (rc) return rc;
    if (n && (!b || !c || !d_)) return XP_EINVAL;
    if (s->method != XP_METHOD_LBVH_PAIRS) return XP_EINVAL;
    if (radii && !(radii[0] > 0.0f && radii[1] > 0.0f && radii[2] > 0.0f)) return XP_EINVAL;
    return XP_OK;
}

int xp_collide_and_slide_batch_dev(xp_scene *s, const xp_vec3 *d_position, const xp_vec3 *d_velocity, uint64_t n,
                                   uint32_t max_iter, const float radii[3], xp_vec3 *d_out_position,
                                   xp_vec3 *d_out_velocity, uint32_t *d_handled) {
    int rc = check_slide(s, d_position, d_velocity, d_out_position, d_out_velocity, n, radii);
    if (rc) return rc;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    XP_API_TRY(slide_dev(s, d_position, d_velocity, n, max_iter, radii, d_out_position, d_out_velocity, d_handled));
    return XP_OK;
}

int xp_collide_and_slide_batch(xp_scene *s, const xp_vec3 *position, const xp_vec3 *velocity, uint64_t n,
                               uint32_t max_iter, const float radii[3], xp_vec3 *out_position,
                               xp_vec3 *out_velocity, uint32_t *handled) {
    int rc = check_slide(s, position, velocity, out_position, out_velocity, n, radii);
    if (rc) return rc;
    XP_BIND(s);
    reset_stats(s);
    if (n == 0) return XP_OK;
    cudaStream_t st = s->stream;
    const size_t vb = n * sizeof(xp_vec3);
    XP_API_TRY(s->io_in.reserve(2 * vb));
    XP_API_TRY(s->io_col.reserve(2 * vb));
    XP_API_TRY(s->io_iter.reserve(n * 4));
    xp_vec3 *d_p = s->io_in.as<xp_vec3>(), *d_v = d_p + n, *o_p = s->io_col.as<xp_vec3>(), *o_v = o_p + n;
    XP_API_TRY(cudaMemcpyAsync(d_p, position, vb, cudaMemcpyHostToDevice, st));
    XP_API_TRY(cudaMemcpyAsync(d_v, velocity, vb, cudaMemcpyHostToDevice, st));
    XP_API_TRY(slide_dev(s, d_p, d_v, n, max_iter, radii, o_p, o_v, s->io_iter.as<uint32_t>()));
    XP_API_TRY(cudaMemcpyAsync(out_position, o_p, vb, cudaMemcpyDeviceToHost, st));
    XP_API_TRY(cudaMemcpyAsync(out_velocity, o_v, vb, cudaMemcpyDeviceToHost, st));
    if (handled) XP_API_TRY(cudaMemcpyAsync(handled, s->io_iter.p, n * 4, cudaMemcpyDeviceToHost, st));
    XP_API_TRY(cudaStreamSynchronize(st));
    return XP_OK;
}

int xp_broadphase_pairs(xp_scene *s, const xp_response *in, uint64_t n, uint32_t horizon_mode, uint64_t cap,
                        uint32_t *sphere_idx, uint32_t *tri_idx, uint64_t *count) {
    int rc = check_query(s, in, n);
    if (rc) return rc;
    if (!count || (cap && (!sphere_idx || !tri_idx))) return XP_EINVAL;
    XP_BIND(s);
    reset_stats(s);
    *count = 0;
    if (n == 0) return XP_OK;
    if (n > (1ull << 24)) return XP_EINVAL;     // introspection call: bounded batch
    cudaStream_t st = s->stream;
    XP_API_TRY(s->io_in.reserve(n * sizeof(xp_response)));
    XP_API_TRY(s->io_tri.reserve((cap ? cap : 1) * 4));
    XP_API_TRY(s->io_iter.reserve((cap ? cap : 1) * 4));
    XP_API_TRY(cudaMemcpyAsync(s->io_in.p, in, n * sizeof(xp_response), cudaMemcpyHostToDevice, st));
    uint64_t total = 0;
    XP_API_TRY(pairs_dev(s, s->io_in.as<xp_response>(), n, horizon_mode, cap, s->io_iter.as<uint32_t>(),
                         s->io_tri.as<uint32_t>(), &total));
    *count = total;
    const uint64_t k = total < cap ? total : cap;
    if (k) {
        XP_API_TRY(cudaMemcpyAsync(sphere_idx, s->io_iter.p, k * 4, cudaMemcpyDeviceToHost, st));
        XP_API_TRY(cudaMemcpyAsync(tri_idx, s->io_tri.p, k * 4, cudaMemcpyDeviceToHost, st));
    }
    XP_API_TRY(cudaStreamSynchronize(st));
    return total > cap ? XP_ECAP : XP_OK;
}

int xp_scene_debug_tree(xp_scene *s, uint32_t *sorted_tri, uint32_t *morton, int32_t *node_child,
                        xp_vec3 *node_lo, xp_vec3 *node_hi) {
    if (!s) return XP_EINVAL;
    if (!s->built) return XP_ESTATE;
    XP_BIND(s);
    const uint64_t n = s->n_tris;
    if (n == 0) return XP_OK;
    XP_API_TRY(cudaStreamSynchronize(s->stream));
    if (sorted_tri) XP_API_TRY(cudaMemcpy(sorted_tri, s->vals_a.p, n * 4, cudaMemcpyDeviceToHost));
    if (morton) XP_API_TRY(cudaMemcpy(morton, s->keys_a.p, n * 4, cudaMemcpyDeviceToHost));
    const uint64_t ni = n > 1 ? n - 1 : 1;
    if (node_child) {
        Node *h = new (std::nothrow) Node[ni];
        if (!h) return XP_ENOMEM;
        cudaError_t e = cudaMemcpy(h, s->nodes.p, ni * sizeof(Node), cudaMemcpyDeviceToHost);
        if (e == cudaSuccess)
            for (uint64_t i = 0; i < ni; ++i) { node_child[2 * i] = h[i].link.x; node_child[2 * i + 1] = h[i].link.y; }
        delete[] h;
        if (e != cudaSuccess) return fail_cuda(e);
    }
    if (node_lo || node_hi) {          // a node's own box = the union of the two child boxes its record holds
        Node *h = new (std::nothrow) Node[ni];
        if (!h) return XP_ENOMEM;
        cudaError_t e = cudaMemcpy(h, s->nodes.p, ni * sizeof(Node), cudaMemcpyDeviceToHost);
        if (e != cudaSuccess) { delete[] h; return fail_cuda(e); }
        for (uint64_t i = 0; i < ni; ++i) {
            const Node &d = h[i];
            if (node_lo) { node_lo[i].x = fminf(d.q0.x, d.q1.z); node_lo[i].y = fminf(d.q0.y, d.q1.w); node_lo[i].z = fminf(d.q0.z, d.q2.x); }
            if (node_hi) { node_hi[i].x = fmaxf(d.q0.w, d.q2.y); node_hi[i].y = fmaxf(d.q1.x, d.q2.z); node_hi[i].z = fmaxf(d.q1.y, d.q2.w); }
        }
        delete[] h;
    }
    return XP_OK;
}

int xp_scene_debug_leaf_runs(xp_scene *s, int32_t *node_run) {
    if (!s || !node_run) return XP_EINVAL;
    if (!s->built) return XP_ESTATE;
    XP_BIND(s);
    const uint64_t n = s->n_tris;
    if (n == 0) return XP_OK;
    XP_API_TRY(cudaStreamSynchronize(s->stream));
    const uint64_t ni = n > 1 ? n - 1 : 1;
    Node *h = new (std::nothrow) Node[ni];
    if (!h) return XP_ENOMEM;
    const cudaError_t e = cudaMemcpy(h, s->nodes.p, ni * sizeof(Node), cudaMemcpyDeviceToHost);
    if (e == cudaSuccess)
        for (uint64_t i = 0; i < ni; ++i) { node_run[2 * i] = h[i].link.z; node_run[2 * i + 1] = h[i].link.w; }
    delete[] h;
    if (e != cudaSuccess) return fail_cuda(e);
    return XP_OK;
}

int xp_sphere_triangle_detect_collision(int device, const xp_response *response, const xp_triangle *triangle,
                                        xp_collision *out) {
    if (!response || !triangle || !out) return XP_EINVAL;
    XP_API_TRY(cudaSetDevice(device));
    int result = 0;
    XP_API_TRY(single_detect_dev(response, triangle, out, &result));
    return result < 0 ? XP_EINVAL : result;
}

int xp_sphere_triangle_calculate_response(int device, const xp_response *response, const xp_collision *collision,
                                          xp_response *out) {
    if (!response || !collision || !out) return XP_EINVAL;
    XP_API_TRY(cudaSetDevice(device));
    int result = 0;
    XP_API_TRY(single_respond_dev(response, collision, out, &result));
    return result;
}

int xp_debug_check_arith(int device, uint64_t n_random, uint64_t seed, uint64_t counters[8]) {
    if (!counters) return XP_EINVAL;
    XP_API_TRY(cudaSetDevice(device));
    static_assert(sizeof(unsigned long long) == sizeof(uint64_t), "counter width");
    XP_API_TRY(check_arith(n_random, seed, reinterpret_cast<unsigned long long *>(counters)));
    return XP_OK;
}

int xp_probe_fp32_peak(int device, double *tflops) {
    if (!tflops) return XP_EINVAL;
    XP_API_TRY(cudaSetDevice(device));
    XP_API_TRY(probe_fp32(tflops));
    return XP_OK;
}

const char *xp_strerror(int code) {
    switch (code) {
    case XP_OK: return "ok";
    case XP_EINVAL: return "invalid argument";
    case XP_ECUDA: return "CUDA error (see xp_last_cuda_error)";
    case XP_ENOMEM: return "out of memory";
    case XP_ESTATE: return "scene not built";
    case XP_ECAP: return "output capacity too small";
    default: return "unknown error";
    }
}

const char *xp_last_cuda_error(void) { return g_cuda_err; }

const char *xp_version(void) { return "xp_physics_cuda 0.1 (sm_100a)"; }

}  // extern "C"
