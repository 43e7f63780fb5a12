// xp_device.cuh -- device data layouts and the scene object shared by the .cu files.
//
// HBM layout (all fp32, 16-byte aligned float4 planes so every access is one LDG.128):
//   TriRec[T]  64 B  leaf j (Morton order): q0=(v0, n.x) q1=(v1, n.y) q2=(v2, n.z)
//                    q3=(plane_constant, |v1-v0|^2, |v2-v1|^2, bits(upload index))
//   Node[T-1]  64 B  internal node i: both child AABBs (already inflated by 1+1e-4) and the two
//                    child links (>= 0 internal, < 0 leaf ~c); node 0 is the root.  link.z / link.w: leaf run of the
//                    left / right child -- l + 1 when that child is an internal node over exactly the leaves l, l + 1,
//                    else 0 (XP_OPT_LEAF_RUNS; written by the tree pass, read by k_broadphase only)
//   RayRec[S]  64 B  per sphere per sweep: q0=(c, |m|) q1=(m, |m|^2) q2=(m/|m|, horizon)
//                    q3=(1/m.x, 1/m.y, 1/m.z, bits(flags))
//   best[S]     8 B  (bits(t) << 32) | ~upload_index, all ones = no hit; atomicMin gives the
//                    reference's min_by with ties going to the later triangle (lib.rs:33-42)
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/xp_physics_cuda.h"

namespace xp {

struct __align__(16) TriRec { float4 q0, q1, q2, q3; };
struct __align__(16) Node { float4 q0, q1, q2; int4 link; };
struct __align__(16) RayRec { float4 q0, q1, q2, q3; };
// Quantised 4-wide node (64 B), built by collapsing two levels of the binary LBVH.  Child boxes are
// 8-bit offsets from the node's own origin in power-of-two cells, rounded OUTWARD by one extra cell,
// so a stage-1 traversal over these nodes emits a superset of P; stage 2 re-checks the exact
// predicate on the triangle's own fp32 box before the narrowphase.
//   w0.xyz = origin (fp32), w0.w = cell exponents (ex | ey << 8 | ez << 16, biased by 127)
//   w1 = (qlo.x[4], qlo.y[4], qlo.z[4], qhi.x[4])   one byte per child
//   w2 = (qhi.y[4], qhi.z[4], link0, link1)         links: >= 0 node, < 0 leaf ~c, Q4_EMPTY = no child
//   w3 = (link2, link3, 0, 0)
struct __align__(16) Node4 { uint4 w0, w1, w2, w3; };
static constexpr int Q4_EMPTY = (int)0x80000000;

// A 64 B record (TriRec, Node, RayRec) or a 32 B box leaves its lane as 256-bit stores (STG.E.256, sm_100+): each writes one
// whole 32 B sector, where the four 16 B stores the compiler makes of a struct assignment write half-sectors that L2 has to
// merge.  p must be 32 B aligned.
#ifndef XP_STG256
#define XP_STG256 1
#endif
#if defined(__CUDACC__)
__device__ __forceinline__ void stg256(void *p, float4 a, float4 b) {
#if XP_STG256
    asm volatile("st.global.v8.f32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};"
                 :: "l"(p), "f"(a.x), "f"(a.y), "f"(a.z), "f"(a.w), "f"(b.x), "f"(b.y), "f"(b.z), "f"(b.w) : "memory");
#else
    reinterpret_cast<float4 *>(p)[0] = a; reinterpret_cast<float4 *>(p)[1] = b;
#endif
}
template <class R> __device__ __forceinline__ void store_rec64(R *dst, const R &r) {     // R = {float4 q0, q1, q2; 16 B}
    static_assert(sizeof(R) == 64, "64 B records only");
    const float4 *q = reinterpret_cast<const float4 *>(&r);
    stg256(dst, q[0], q[1]);
    stg256(reinterpret_cast<char *>(dst) + 32, q[2], q[3]);
}
#endif

static constexpr unsigned long long BEST_NONE = 0xFFFFFFFFFFFFFFFFull;
static constexpr uint32_t RAY_VALID = 1u;
static constexpr uint32_t RAY_GRID_OK = 2u;      // centre and movement finite: the height-grid walk may take this ray
static constexpr uint32_t HARD_ITER_CAP = 65536u;

// Growable device buffer (never shrinks; freed with the scene).
struct DevBuf {
    void *p = nullptr;
    size_t cap = 0;
    cudaError_t reserve(size_t bytes) {
        if (bytes <= cap) return cudaSuccess;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        size_t want = bytes + bytes / 8 + 256;
        cudaError_t e = cudaMalloc(&p, want);
        if (e != cudaSuccess) { e = cudaMalloc(&p, bytes); want = bytes; }
        if (e == cudaSuccess) cap = want;
        return e;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
    template <class T> T *as() const { return reinterpret_cast<T *>(p); }
};

// 32-wide implicit tree over the Morton-sorted triangles: level 0 groups 32 consecutive triangles,
// level L groups 32 consecutive level-(L-1) nodes; node i of level L has children 32i..32i+31.
// One node = 6 planes of 32 floats (lo.x[32] lo.y[32] lo.z[32] hi.x[32] hi.y[32] hi.z[32]) = 768 B,
// so a warp reads a node's 32 child boxes with six fully coalesced 128 B loads.
static constexpr int WIDE_MAX_LEVELS = 6;          // 32^6 > 2^27 triangles
static constexpr int WIDE_NODE_FLOATS = 192;
struct WideTree {
    const float *planes;                           // all levels, level L starts at node off[L]
    int top;                                       // index of the root level (levels = top + 1)
    unsigned off[WIDE_MAX_LEVELS];
};

struct DeviceCounters {           // one small block of device memory, zeroed per call
    unsigned long long tests;
    unsigned long long pairs;
    unsigned long long node_visits;
    unsigned long long path[5];   // tests per code path (XP_PATH_*), only with XP_OPT_TIMING=1
    unsigned long long ray_cursor;  // k_broadphase: next unclaimed run of 32 rays
    unsigned long long pairs_total; // candidates over all chunks of the call
    unsigned long long redone;      // narrowphase tests redone with the literal form (operands outside the guarded range)
    unsigned int next_active;
    unsigned int stack_overflow;
    unsigned int pair_overflow;     // a chunk produced more candidates than the pair buffer holds
    unsigned int irregular_rays;    // rays the height-grid walk left to the LBVH walk (non-finite centre / movement)
    unsigned int active[2];         // response loop: live spheres of the current / next iteration (double-buffered)
    unsigned int iters_done;        // response loop: iterations that had work
    unsigned int pad0;
    unsigned long long rays_swept;  // response loop: sum of the live counts over its iterations
    unsigned long long np_cursor;   // k_narrow_pairs: next unclaimed span of the pair list (the last block to leave resets it)
    unsigned int np_done;           // k_narrow_pairs: blocks that have left
    unsigned int pad1;
};

}  // namespace xp

struct xp_scene {
    int device = 0;
    int n_sm = 1;                      // multiProcessorCount of the device (persistent grids are sized from it)
    cudaStream_t stream = nullptr;
    // options
    int arith = XP_ARITH_EXACT;
    int method = XP_METHOD_LBVH_PAIRS;
    int prune = 2;                     // XP_OPT_PRUNE: auto (distance windows once a sweep saw > 96 candidates per sphere)
    int timing = 0;
    int sort_spheres = 1;
    int contact_record = 0;            // XP_OPT_CONTACT_RECORD: 0 = xp_contact (40 B), 1 = xp_contact16 (16 B)
    int leaf_runs = 1;                 // XP_OPT_LEAF_RUNS: two-leaf children are emitted, not visited (stage 2 filters with P)
    int far_exact = 0;                 // XP_OPT_FAR_EXACT: second pass for far-range grazing noise of the reference
    uint64_t max_pairs = 1ull << 26;
    float cur_horizon = 0.0f;          // horizon of the rays currently in `rays`
    bool safe_pairs = false;           // two-stage methods: read the pair count back after every chunk
    bool pair_overflowed = false;      // optimistic mode dropped candidates: results must be recomputed
    bool counters_live = false;        // device counters hold uncollected statistics
    const unsigned *na_dev = nullptr;  // response loop: device address of the live ray count (null: the host knows it)
    uint64_t swept_rays = 0;           // rays swept since the counters were zeroed
    double cand_per_ray = 0.0;         // candidates per sphere of the last exhaustive sweep (XP_OPT_PRUNE = 2)
    // geometry
    uint64_t n_tris = 0;
    bool built = false;
    bool q4_built = false, wide_built = false;   // optional structures, see build_aux()
    xp::DevBuf tris_aos;      // xp_triangle[T], upload order
    xp::DevBuf recs;          // TriRec[T], Morton order
    xp::DevBuf nodes;         // Node[max(T-1,1)]
    xp::DevBuf leaf_box;                             // float4 (lo, hi) pairs, 32 B per leaf: what the bottom-up tree pass starts from
    xp::DevBuf refit_flags;                          // the tree pass's arrival slots (+ 4 ints: root info, work-list length)
    xp::DevBuf tree_work;                            // the tree pass's work list (subtrees that continue through the global protocol)
    xp::DevBuf keys_a, keys_b, vals_a, vals_b;       // radix sort ping-pong (triangles / spheres)
    xp::DevBuf hist;          // radix histograms + scan scratch
    xp::DevBuf bounds;        // 6 ordered-int floats: scene AABB of triangle centres
    xp::DevBuf wide, wide_lo, wide_hi;               // 32-wide implicit tree planes + per-node union boxes
    xp::DevBuf nodes4;                               // Node4[max(T-1,1)], indexed like the binary nodes
    xp::DevBuf degen;                                // uint32[1 + T]: count, then the leaves of ill-conditioned triangles
    // heightmap scenes (xp_scene_set_heightmap): the height grid itself is the broadphase structure (xp_gridwalk.cuh)
    bool grid_valid = false;                         // the triangles are exactly the soup of this grid
    int use_grid = 1;                                // XP_OPT_GRID_WALK
    unsigned grid_nx = 0, grid_nz = 0;
    float grid_x0 = 0, grid_z0 = 0, grid_sp = 0;
    float grid_eps = 0.0f;                           // slack on enumerated positions (~32 ulp of the largest coordinate)
    xp::DevBuf grid_mm;                              // nx x nz (min, max) of the heights under windows of 8 cells along z
    xp::DevBuf grid_h;                               // (nx+1) x (nz+1) heights
    xp::DevBuf grid_yrange;                          // 2 floats: min / max height (2 ordered ints while being reduced)
    xp::DevBuf recs_grid;                            // TriRec[T] in UPLOAD order (cell-major): what the grid walk's pairs index
    uint64_t n_degen = 0;                            // (swept exhaustively, see k_pack / k_sweep_degenerate)
    xp::WideTree wide_tree = {};
    // per-call workspaces
    xp::DevBuf rays, best, active_a, active_b, inv_order, pairs, counters;
    xp::DevBuf rays2, best2, active_c, io_backup;       // response loop: next iteration's rays / keys / list, and the input kept for a redo
    xp::DevBuf sort_ka, sort_kb;   // sphere Morton keys (ping-pong)
    xp::DevBuf io_in, io_out, io_hit, io_tri, io_status, io_iter, io_normal, io_col;
    xp::DevBuf small_dev;          // the small-batch response path's one block each way: counters | in | out | iterations | status | normals
    void *small_pin = nullptr;     // ... and its pinned mirror
    // stats
    xp_stats stats;
    // XP_OPT_TIMING: a pool of event pairs recorded without synchronising; resolved into
    // stats.stage_ms when the call collects its counters
    static constexpr int EV_POOL = 96;
    cudaEvent_t ev[2 * EV_POOL] = {};
    int ev_stage[EV_POOL] = {};
    int ev_count = 0;
    int ev_base = 0;               // first pool entry of the call being recorded (asynchronous steps use half a pool each)
    bool ev_created = false;
    // host-pointer calls: copy streams + per-chunk events for the H2D / compute / D2H pipeline
    cudaStream_t copy_in = nullptr, copy_out = nullptr;
    cudaEvent_t io_ev[32] = {};
    bool io_ready = false;
    // xp_detect_batch_submit / _wait: batches in flight, each with its own device buffers, so that
    // one batch's copies overlap another batch's kernels
    static constexpr int ASYNC_SLOTS = 2;
    struct AsyncSlot {
        xp::DevBuf in, col, hit, tri, status, compact;
        cudaEvent_t h2d = nullptr, done = nullptr, d2h = nullptr;
        xp::DeviceCounters *h_ctr = nullptr;     // pinned snapshot of the device counters of this batch
        bool busy = false;
        const xp_response *h_in = nullptr;       // the call's arguments (a dropped-candidate batch is redone)
        uint64_t n = 0;
        uint32_t horizon_mode = 0, launches = 0;
        xp_collision *out = nullptr;
        uint8_t *hit_out = nullptr;
        uint32_t *tri_out = nullptr, *status_out = nullptr;
        xp_hit *compact_out = nullptr;           // compact form (xp_detect_batch_compact_submit)
    };
    AsyncSlot async_slot[ASYNC_SLOTS];
    bool async_ready = false;
    // xp_detect_contacts_exchange_submit / _wait: the record build + exchange of step i runs on
    // exchange_stream (caller-provided) under the sweep of step i+1; each step in flight owns the
    // buffers its finalize kernel reads
    cudaStream_t exchange_stream = nullptr;
    struct ExSlot {
        xp::DevBuf rays, best, inv_order, status;
        cudaEvent_t np = nullptr, fin = nullptr;
        xp::DeviceCounters *h_ctr = nullptr;     // pinned snapshot of the counters after the narrowphase
        bool busy = false, used = false;
        int ev_count = 0;
        uint32_t launches = 0;
    };
    ExSlot ex_slot[2];
    int ex_next = 0;
    bool ex_ready = false;
};
