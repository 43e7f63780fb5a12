// xp_build.cu -- device LBVH construction over the uploaded triangle soup (sm_100a).
//
// New relative to the reference, which has no broadphase: collision_response walks every
// triangle on every iteration (/root/reference/xp_physics/src/lib.rs:33-35).  The per-triangle
// quantities cached here are the ones collision_detect.rs:162,171 recomputes on every test
// (Triangle::normal, triangle.rs:44-46; Triangle::plane_constant, triangle.rs:47-50).
//
// Stages (bytes per triangle and measured times in DESIGN.md section 4):
//   K0 bounds   AABB of the triangle-box centres (ordered-int atomics, block-reduced)
//   K2 morton   30-bit Morton code of each centre (+ the per-tile histogram of the sort's first digit)
//   K3 sort     CUB-free stable LSD radix sort, 10 bits x 3 passes (histogram / scan / scatter through the tile's
//               own digit order in shared memory)
//   K1 pack     TriRec + inflated leaf AABB written in Morton order, the ill-conditioned list
//   K4+K5 tree  hierarchy AND child boxes in one bottom-up pass (k_tree_local: a warp per 256 leaves in shared
//               memory; k_tree_global: what crosses segments), writes the 64 B nodes (child boxes, links and the leaf-run
//               words of XP_OPT_LEAF_RUNS); node 0 = root
#include <cstdio>
#include <cstdlib>

#include "xp_internal.h"
#include "xp_narrowphase.cuh"
#include "xp_gridwalk.cuh"
#include "xp_rays.cuh"

namespace xp {

// ------------------------------------------------------------------------------------------
// ordered-int encoding so float min/max can use integer atomics
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned enc_f(float f) {
    unsigned b = __float_as_uint(f);
    return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__device__ __forceinline__ float dec_f(unsigned e) {
    unsigned b = (e & 0x80000000u) ? (e & 0x7FFFFFFFu) : ~e;
    return __uint_as_float(b);
}

__device__ __forceinline__ void load_tri(const float *__restrict__ aos, uint64_t i, V3 &v0, V3 &v1, V3 &v2) {
    const float *p = aos + i * 9;
    v0 = {__ldg(p + 0), __ldg(p + 1), __ldg(p + 2)};
    v1 = {__ldg(p + 3), __ldg(p + 4), __ldg(p + 5)};
    v2 = {__ldg(p + 6), __ldg(p + 7), __ldg(p + 8)};
}

__device__ __forceinline__ V3 box_centre(V3 lo, V3 hi) {
    return {0.5f * (lo.x + hi.x), 0.5f * (lo.y + hi.y), 0.5f * (lo.z + hi.z)};
}

// K0 -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_bounds(const float *__restrict__ aos, uint64_t n,
                                                unsigned *__restrict__ bounds) {
    float mn[3] = {INFINITY, INFINITY, INFINITY}, mx[3] = {-INFINITY, -INFINITY, -INFINITY};
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n;
         i += (uint64_t)gridDim.x * blockDim.x) {
        V3 v0, v1, v2, lo, hi;
        load_tri(aos, i, v0, v1, v2);
        xp_tri_aabb(v0, v1, v2, lo, hi);
        V3 c = box_centre(lo, hi);
        mn[0] = fminf(mn[0], c.x); mn[1] = fminf(mn[1], c.y); mn[2] = fminf(mn[2], c.z);
        mx[0] = fmaxf(mx[0], c.x); mx[1] = fmaxf(mx[1], c.y); mx[2] = fmaxf(mx[2], c.z);
    }
#pragma unroll
    for (int k = 0; k < 3; ++k)
        for (int o = 16; o > 0; o >>= 1) {
            mn[k] = fminf(mn[k], __shfl_xor_sync(0xffffffffu, mn[k], o));
            mx[k] = fmaxf(mx[k], __shfl_xor_sync(0xffffffffu, mx[k], o));
        }
    if ((threadIdx.x & 31) == 0) {
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            atomicMin(&bounds[k], enc_f(mn[k]));
            atomicMax(&bounds[3 + k], enc_f(mx[k]));
        }
    }
}

__global__ void k_bounds_init(unsigned *bounds) {
    if (threadIdx.x < 3) bounds[threadIdx.x] = 0xFFFFFFFFu;
    else if (threadIdx.x < 6) bounds[threadIdx.x] = 0u;
}

// K2 -----------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned expand_bits10(unsigned v) {
    v = (v * 0x00010001u) & 0xFF0000FFu;
    v = (v * 0x00000101u) & 0x0F00F00Fu;
    v = (v * 0x00000011u) & 0xC30C30C3u;
    v = (v * 0x00000005u) & 0x49249249u;
    return v;
}
__device__ __forceinline__ unsigned morton30(V3 c, const unsigned *__restrict__ bounds) {
    // ISOTROPIC quantisation: every axis is scaled by the LARGEST extent, so a cell is a cube.
    // (Per-axis scaling would spend 10 bits on the few units of height of a terrain and destroy
    // the x/z coherence of the curve.)
    float q[3] = {c.x, c.y, c.z};
    float lo[3], ext = 0.0f;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        lo[k] = dec_f(bounds[k]);
        ext = fmaxf(ext, dec_f(bounds[3 + k]) - lo[k]);
    }
    const float scale = (ext > 0.0f) ? 1024.0f / ext : 0.0f;
    unsigned code = 0;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        const float u = fminf(fmaxf((q[k] - lo[k]) * scale, 0.0f), 1023.0f);
        code |= expand_bits10((unsigned)u) << (2 - k);
    }
    return code;
}

// One block = one tile of the radix sort (RS_TILE triangles): besides codes and indices it leaves the tile's histogram
// of the FIRST digit (hist, digit-major like k_rs_hist's; may be null), so the sort's first pass starts at its scan.
static constexpr int MT_TILE = 4096, MT_BINS = 1024;           // == RS_TILE, RS_BINS (asserted where those are defined)
__global__ void __launch_bounds__(256) k_morton_tris(const float *__restrict__ aos, uint64_t n,
                                                     const unsigned *__restrict__ bounds,
                                                     uint32_t *__restrict__ keys,
                                                     uint32_t *__restrict__ vals, uint32_t *__restrict__ hist, unsigned num_tiles) {
    __shared__ unsigned h[MT_BINS];
    if (hist) {
        for (int b = threadIdx.x; b < MT_BINS; b += 256) h[b] = 0;
        __syncthreads();
    }
    const uint64_t base = (uint64_t)blockIdx.x * MT_TILE;
#pragma unroll 4
    for (int c = 0; c < MT_TILE / 256; ++c) {
        const uint64_t i = base + (uint64_t)c * 256 + threadIdx.x;
        if (i < n) {
            V3 v0, v1, v2, lo, hi;
            load_tri(aos, i, v0, v1, v2);
            xp_tri_aabb(v0, v1, v2, lo, hi);
            const uint32_t key = morton30(box_centre(lo, hi), bounds);
            keys[i] = key;
            vals[i] = (uint32_t)i;
            if (hist) atomicAdd(&h[key & (MT_BINS - 1)], 1u);
        }
    }
    if (hist) {
        __syncthreads();
        for (int b = threadIdx.x; b < MT_BINS; b += 256) hist[(uint64_t)b * num_tiles + blockIdx.x] = h[b];
    }
}

__global__ void __launch_bounds__(256) k_morton_spheres(const xp_response *__restrict__ in, uint64_t n,
                                                        const unsigned *__restrict__ bounds,
                                                        uint32_t *__restrict__ keys,
                                                        uint32_t *__restrict__ vals, int mode,
                                                        uint32_t *__restrict__ bins) {
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    unsigned key = 0xFFFFFFFFu;              // lanes past the end: a key no sphere has
    if (i < n) {
        const float *p = reinterpret_cast<const float *>(in) + i * 7;
        V3 c = {__ldg(p + 0), __ldg(p + 1), __ldg(p + 2)};
        const unsigned code = morton30(c, bounds);
        // mode 0: 24-bit Morton code of the centre, 8 bits per axis.  mode 1: direction octant (signs of
        // the movement) on top of a 7-bit-per-axis Morton code, so the lanes of a warp walk the tree in
        // the same order.  mode 2 (default): 16 bits, 5-6 bits per axis.
        key = code >> 6;
        if (mode == 1) {
            const unsigned oct = (__ldg(p + 4) < 0.0f ? 4u : 0u) | (__ldg(p + 5) < 0.0f ? 2u : 0u) | (__ldg(p + 6) < 0.0f ? 1u : 0u);
            key = (oct << 21) | (code >> 9);
        }
        if (mode == 2) key = code >> 14;
        keys[i] = key;
        if (!bins) vals[i] = (uint32_t)i;
    }
    if (bins) {                              // counting sort: one atomic per distinct key per warp
        const unsigned peers = __match_any_sync(0xffffffffu, key);
        if (i < n && (threadIdx.x & 31) == (unsigned)(__ffs(peers) - 1)) atomicAdd(&bins[key], (unsigned)__popc(peers));
    }
}

// Counting sort on 16-bit keys (the default processing order of the spheres): bin sizes come from
// k_morton_spheres, k_bin_offsets turns them into offsets (per group of 1024 bins; the scatter adds the
// prefix of the 64 group totals), and every sphere takes the next free slot of
// its bin.  The order INSIDE a bin depends on the atomics' timing; that is harmless -- the order only
// decides which lanes walk the tree together, every result is written back through the inverse
// permutation and the closest hit is an order-independent atomicMin.
static constexpr int CS_BINS = 1 << 16;
static constexpr int CS_GROUPS = CS_BINS / 1024;             // one block of k_bin_offsets per 1024 bins
__global__ void __launch_bounds__(1024) k_bin_offsets(uint32_t *__restrict__ bins, uint32_t *__restrict__ group_tot) {
    __shared__ unsigned warp_tot[32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    const unsigned v = bins[blockIdx.x * 1024 + threadIdx.x];
    unsigned inc = v;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    const unsigned t = warp_tot[lane];                       // every warp scans the 32 warp totals itself
    unsigned ti = t;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, ti, o); if (lane >= o) ti += y; }
    bins[blockIdx.x * 1024 + threadIdx.x] = __shfl_sync(0xffffffffu, ti - t, w) + (inc - v);   // offset inside the group
    if (threadIdx.x == 1023) group_tot[blockIdx.x] = ti;
}

// RAYS: the scatter also writes the sphere's ray record at its slot -- the ray setup of the sweep that follows, fused
// (one launch and one gather of the Response less); the empty best keys are a memset before the launch.
// XP_SCATTER_COOP: the 64 B record leaves through shared memory, four lanes per record, so one store instruction of a
// warp writes 8 whole records (16 full 32 B sectors) instead of a 16 B quarter of 32 different records.
#ifndef XP_SCATTER_COOP
#define XP_SCATTER_COOP 1
#endif
template <class O, bool RAYS>
__global__ void __launch_bounds__(256) k_bin_scatter(const uint32_t *__restrict__ keys, uint64_t n,
                                                     uint32_t *__restrict__ offs, const uint32_t *__restrict__ group_tot,
                                                     uint32_t *__restrict__ order, uint32_t *__restrict__ inv,
                                                     const xp_response *__restrict__ in, RayRec *__restrict__ rays,
                                                     unsigned long long *__restrict__ best, uint32_t *__restrict__ status,
                                                     int have_tris) {
    __shared__ unsigned group_base[CS_GROUPS];
#if XP_SCATTER_COOP == 1
    __shared__ float4 stage_all[RAYS ? 256 / 32 : 1][RAYS ? 32 * 5 : 1];    // a record every 80 B: the 16 B accesses of a quarter-warp fall into distinct banks
#endif
    const uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    // everything this thread reads is requested first: its key and (RAYS) its Response
    const unsigned key = i < n ? __ldg(keys + i) : 0xFFFFFFFFu;
    float in7[7] = {0.f, 0.f, 0.f, 1.f, 0.f, 0.f, 0.f};
    if (RAYS && i < n) {
        const float *p = reinterpret_cast<const float *>(in) + i * 7;
#pragma unroll
        for (int k = 0; k < 7; ++k) in7[k] = __ldg(p + k);
    }
    if (threadIdx.x < 32) {                                  // exclusive prefix of the 64 group totals
        const unsigned a = group_tot[2 * threadIdx.x], b = group_tot[2 * threadIdx.x + 1];
        unsigned inc = a + b;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, inc, o); if ((int)threadIdx.x >= o) inc += y; }
        group_base[2 * threadIdx.x] = inc - a - b;
        group_base[2 * threadIdx.x + 1] = inc - b;
    }
    __syncthreads();
    const int lane = threadIdx.x & 31;
    const unsigned peers = __match_any_sync(0xffffffffu, key);
    const int leader = __ffs(peers) - 1;
    unsigned base = 0;
    if (i < n && lane == leader) base = group_base[key >> 10] + atomicAdd(&offs[key], (unsigned)__popc(peers));
    // The slot comes back from L2 about a microsecond later (ncu: a fifth of the kernel's stall samples sat on it, and the
    // Response was only read after it).  Nothing below needs it until the stores, so the record is built in the meantime.
    RayRec rr;
    bool valid = true;
    if (RAYS && i < n) {
        valid = in7[3] == 1.0f;
        rr = make_ray_rec<O>(V3{in7[0], in7[1], in7[2]}, V3{in7[4], in7[5], in7[6]}, valid);
    }
    base = __shfl_sync(0xffffffffu, base, leader);
    const unsigned pos = base + __popc(peers & ((1u << lane) - 1u));
#if XP_SCATTER_COOP == 1
    if (RAYS) {
        float4 *stage = stage_all[threadIdx.x >> 5];
        if (i < n) {
            stage[lane * 5 + 0] = rr.q0; stage[lane * 5 + 1] = rr.q1; stage[lane * 5 + 2] = rr.q2; stage[lane * 5 + 3] = rr.q3;
        }
        __syncwarp();
        const uint64_t warp_first = i - lane;                               // the warp's lanes are consecutive spheres
#pragma unroll
        for (int r = 0; r < 4; ++r) {
            const int src = r * 8 + (lane >> 2), part = lane & 3;
            const unsigned dst = __shfl_sync(0xffffffffu, pos, src);
            if (warp_first + src < n) reinterpret_cast<float4 *>(rays + dst)[part] = stage[src * 5 + part];
        }
    }
#endif
    if (i >= n) return;
    order[pos] = (uint32_t)i;
    if (inv) inv[i] = pos;
    if (RAYS) {
        if (!valid && status && have_tris) status[i] |= XP_ST_RADIUS_NE_1;     // collision_detect.rs:161 (see k_ray_setup)
#if XP_SCATTER_COOP == 2
        store_rec64(rays + pos, rr);
#elif !XP_SCATTER_COOP
        store_rec64(rays + pos, rr);
        best[pos] = BEST_NONE;
#endif
    }
}

// K3: radix sort -------------------------------------------------------------------------------
static constexpr int RS_THREADS = 256;
static constexpr int RS_ITEMS = 16;
static constexpr int RS_TILE = RS_THREADS * RS_ITEMS;   // 4096 keys per block
static constexpr int RS_WARPS = RS_THREADS / 32;
static constexpr int RS_SEG = RS_TILE / RS_WARPS;       // 512 consecutive keys per warp
#ifndef XP_RS_BITS
#define XP_RS_BITS 10
#endif
static constexpr int RS_BITS = XP_RS_BITS;              // digit width: 10 bits = three passes over the 30-bit Morton codes
static constexpr int RS_BINS = 1 << RS_BITS;
static constexpr unsigned RS_MASK = RS_BINS - 1;
static_assert(RS_TILE == MT_TILE && RS_BINS == MT_BINS, "k_morton_tris leaves the first pass's histogram in the sort's tile layout");

// per-tile digit histogram, stored digit-major so one flat exclusive scan yields the bases
__global__ void __launch_bounds__(RS_THREADS) k_rs_hist(const uint32_t *__restrict__ keys, uint64_t n,
                                                        int shift, uint32_t *__restrict__ hist,
                                                        unsigned num_tiles) {
    __shared__ unsigned h[RS_BINS];
    for (int b = threadIdx.x; b < RS_BINS; b += RS_THREADS) h[b] = 0;
    __syncthreads();
    const uint64_t base = (uint64_t)blockIdx.x * RS_TILE;
#pragma unroll 4
    for (int c = 0; c < RS_ITEMS; ++c) {
        uint64_t idx = base + (uint64_t)c * RS_THREADS + threadIdx.x;
        if (idx < n) atomicAdd(&h[(keys[idx] >> shift) & RS_MASK], 1u);
    }
    __syncthreads();
    for (int b = threadIdx.x; b < RS_BINS; b += RS_THREADS) hist[(uint64_t)b * num_tiles + blockIdx.x] = h[b];
}

// stable scatter: warp w owns keys [w*512, (w+1)*512) of the tile, processed in chunks of 32.
// inv_out (last pass only, may be null): the inverse permutation, inv_out[value] = final position.
// Lanes holding the same digit are found with one ballot per digit bit (registers only) rather than MATCH.ANY, whose
// result comes back through the shared-memory pipe: ncu showed the kernel waiting on exactly that (short scoreboard 18
// per issue, 14 % of the issue slots used).  The rank of a key inside its warp's segment is fixed in the first pass and
// kept in a register, so the second pass is a table lookup; values are read where they are written (no register copy).
#ifndef XP_RS_MATCH
#define XP_RS_MATCH 1              // 1: ballots per digit bit, 0: __match_any_sync
#endif
__device__ __forceinline__ unsigned rs_peers(unsigned d, unsigned vm) {
#if XP_RS_MATCH
    unsigned peers = vm;
#pragma unroll
    for (int b = 0; b < RS_BITS; ++b) {
        const unsigned bal = __ballot_sync(0xffffffffu, (d >> b) & 1u);
        peers &= ((d >> b) & 1u) ? bal : ~bal;
    }
    return peers;
#else
    return __match_any_sync(vm, d);
#endif
}
// the scan of the tiles x bins table works on blocks of SC_TILE entries (k_scan_local)
static constexpr int SC_THREADS = 256;
static constexpr int SC_ITEMS = 8;
static constexpr int SC_TILE = SC_THREADS * SC_ITEMS;
#ifndef XP_RS_MINB
#define XP_RS_MINB 4
#endif
#ifndef XP_RS_LOCALSORT
#define XP_RS_LOCALSORT 1          // 1: the tile is put in digit order in shared memory first, so that global stores go out in runs
#endif
// With 10-bit digits a tile of 4096 keys holds ~4 keys per digit: written straight from the ranking pass, every lane of
// a store instruction lands in a different 32 B sector -- 2 x 10 M sector requests per pass, and the request rate, not the
// bytes, is what the kernel waits for (ncu: lts 52 %, DRAM 13 %).  So the keys (and the tile-local index of their value)
// are first placed at their position inside the tile's own digit order in shared memory; thread j then stores sorted
// element j, and consecutive elements of a digit go to consecutive addresses: a warp's store touches ~9 sectors, not 32.
__global__ void __launch_bounds__(RS_THREADS, XP_RS_MINB) k_rs_scatter(const uint32_t *__restrict__ keys_in,
                                                           const uint32_t *__restrict__ vals_in,
                                                           uint32_t *__restrict__ keys_out,
                                                           uint32_t *__restrict__ vals_out, uint64_t n,
                                                           int shift, const uint32_t *__restrict__ scanned,
                                                           const uint32_t *__restrict__ block_sums,
                                                           unsigned num_tiles, uint32_t *__restrict__ inv_out) {
#if XP_RS_LOCALSORT
    __shared__ uint16_t cnt[RS_WARPS][RS_BINS];        // per (warp, digit): count, then the tile-local position of the warp's first such key
    __shared__ uint32_t s_key[RS_TILE];                // the tile in digit order
    __shared__ uint16_t s_src[RS_TILE];                // ... and where in the tile each key came from (its value is fetched from there)
    __shared__ uint32_t delta[RS_BINS];                // global position of a digit's first key of this tile - its tile-local position
    __shared__ unsigned warp_tot[RS_WARPS];
#else
    __shared__ unsigned cnt[RS_WARPS][RS_BINS];
#endif
    const int w = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const unsigned lt_mask = (1u << lane) - 1u;
    for (int i = threadIdx.x; i < RS_WARPS * RS_BINS; i += RS_THREADS) (&cnt[0][0])[i] = 0;
    __syncthreads();
    const uint64_t tile_base = (uint64_t)blockIdx.x * RS_TILE;
    const uint64_t base = tile_base + (uint64_t)w * RS_SEG;
    uint32_t k[RS_ITEMS];
    uint16_t rank[RS_ITEMS];               // position among the keys of this warp's segment that share the digit (< 512)
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const uint64_t idx = base + (uint64_t)c * 32 + lane;
        k[c] = idx < n ? keys_in[idx] : 0xFFFFFFFFu;
        // the values are first touched in pass B, a whole ranking pass from here: ask for them now (L2)
        if ((lane & 7) == 0 && idx < n) asm volatile("prefetch.global.L2 [%0];" ::"l"(vals_in + idx));   // one per 32 B sector
    }
    // pass A: per-warp digit counts and each key's rank inside the warp's segment (original order -> stable)
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const uint64_t idx = base + (uint64_t)c * 32 + lane;
        const bool valid = idx < n;
        const unsigned vm = __ballot_sync(0xffffffffu, valid);
        const unsigned d = (k[c] >> shift) & RS_MASK;
        const unsigned peers = rs_peers(d, vm) & vm;
        unsigned before = 0;
        if (valid) before = cnt[w][d];
        __syncwarp();
        if (valid && lane == (31 - __clz(peers))) cnt[w][d] = before + __popc(peers);
        rank[c] = (uint16_t)(before + __popc(peers & lt_mask));
        __syncwarp();
    }
    __syncthreads();
#if XP_RS_LOCALSORT
    {
        // thread t owns digits 4t .. 4t+3: exclusive prefix over the warps, then over all 1024 digits of the tile
        constexpr int DPT = RS_BINS / RS_THREADS;
        unsigned tot[DPT], sum = 0;
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            const unsigned d = threadIdx.x * DPT + q;
            unsigned run = 0;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ++ww) { const unsigned t = cnt[ww][d]; cnt[ww][d] = (uint16_t)run; run += t; }
            tot[q] = run; sum += run;
        }
        unsigned inc = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) { const unsigned y = __shfl_up_sync(0xffffffffu, inc, o); if (lane >= o) inc += y; }
        if (lane == 31) warp_tot[w] = inc;
        __syncthreads();
        unsigned before_warp = 0;
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) before_warp += ww < w ? warp_tot[ww] : 0u;
        unsigned lb = before_warp + inc - sum;           // tile-local position of this thread's first digit
#pragma unroll
        for (int q = 0; q < DPT; ++q) {
            const unsigned d = threadIdx.x * DPT + q;
#pragma unroll
            for (int ww = 0; ww < RS_WARPS; ++ww) cnt[ww][d] = (uint16_t)(cnt[ww][d] + lb);
            const uint64_t at = (uint64_t)d * num_tiles + blockIdx.x;          // the table was scanned per block of SC_TILE entries:
            delta[d] = scanned[at] + block_sums[at / SC_TILE] - lb;             // the block's base is added here, not by a k_scan_add pass
            lb += tot[q];
        }
    }
    __syncthreads();
    // pass B: into the tile's digit order
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const uint64_t idx = base + (uint64_t)c * 32 + lane;
        if (idx < n) {
            const unsigned lpos = cnt[w][(k[c] >> shift) & RS_MASK] + rank[c];
            s_key[lpos] = k[c];
            s_src[lpos] = (uint16_t)(w * RS_SEG + c * 32 + lane);
        }
    }
    __syncthreads();
    // pass C: out, element j of the sorted tile from thread j
    const unsigned nv = (unsigned)((n - tile_base < (uint64_t)RS_TILE) ? (n - tile_base) : (uint64_t)RS_TILE);
#pragma unroll 4
    for (unsigned j = threadIdx.x; j < nv; j += RS_THREADS) {
        const uint32_t key = s_key[j];
        const unsigned pos = j + delta[(key >> shift) & RS_MASK];
        const uint32_t v = vals_in[tile_base + s_src[j]];
        keys_out[pos] = key;
        vals_out[pos] = v;
        if (inv_out) inv_out[v] = pos;
    }
#else
    for (unsigned d = threadIdx.x; d < (unsigned)RS_BINS; d += RS_THREADS) {   // exclusive prefix over warps, seeded with this tile's global base for the digit
        const uint64_t at = (uint64_t)d * num_tiles + blockIdx.x;
        unsigned run = scanned[at] + block_sums[at / SC_TILE];
#pragma unroll
        for (int ww = 0; ww < RS_WARPS; ++ww) {
            unsigned t = cnt[ww][d];
            cnt[ww][d] = run;
            run += t;
        }
    }
    __syncthreads();
    // pass B: a lookup
#pragma unroll
    for (int c = 0; c < RS_ITEMS; ++c) {
        const uint64_t idx = base + (uint64_t)c * 32 + lane;
        if (idx < n) {
            const unsigned pos = cnt[w][(k[c] >> shift) & RS_MASK] + rank[c];
            const uint32_t v = vals_in[idx];
            keys_out[pos] = k[c];
            vals_out[pos] = v;
            if (inv_out) inv_out[v] = pos;
        }
    }
#endif
}

// exclusive scan of a uint32 array (hierarchical: 2048 elements per block) -------------------

__global__ void __launch_bounds__(SC_THREADS) k_scan_local(uint32_t *__restrict__ data, uint64_t n,
                                                           uint32_t *__restrict__ sums) {
    __shared__ unsigned warp_tot[SC_THREADS / 32];
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
    unsigned x[SC_ITEMS], tot = 0;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        x[i] = (base + i < n) ? data[base + i] : 0u;
        tot += x[i];
    }
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    unsigned inc = tot;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        unsigned y = __shfl_up_sync(0xffffffffu, inc, o);
        if (lane >= o) inc += y;
    }
    if (lane == 31) warp_tot[w] = inc;
    __syncthreads();
    if (w == 0) {
        unsigned t = (lane < SC_THREADS / 32) ? warp_tot[lane] : 0u;
        unsigned ti = t;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            unsigned y = __shfl_up_sync(0xffffffffu, ti, o);
            if (lane >= o) ti += y;
        }
        if (lane < SC_THREADS / 32) warp_tot[lane] = ti - t;   // exclusive
        if (lane == SC_THREADS / 32 - 1 && sums) sums[blockIdx.x] = ti;
    }
    __syncthreads();
    unsigned run = warp_tot[w] + (inc - tot);
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i) {
        if (base + i < n) data[base + i] = run;
        run += x[i];
    }
}

__global__ void __launch_bounds__(SC_THREADS) k_scan_add(uint32_t *__restrict__ data, uint64_t n,
                                                         const uint32_t *__restrict__ sums) {
    const unsigned add = sums[blockIdx.x];
    const uint64_t base = (uint64_t)blockIdx.x * SC_TILE + (uint64_t)threadIdx.x * SC_ITEMS;
#pragma unroll
    for (int i = 0; i < SC_ITEMS; ++i)
        if (base + i < n) data[base + i] += add;
}

// scratch must hold ceil(n/2048) + ceil(that/2048) + ... uint32
// histogram table + scan scratch (geometric series of block sums: < hist_n/2047 + 64)
static cudaError_t radix_sort_reserve(xp_scene *s, uint64_t n) {
    const uint64_t hist_n = (uint64_t)grid_for(n, RS_TILE) * RS_BINS;
    return s->hist.reserve((hist_n + hist_n / 1024 + 4096) * sizeof(uint32_t));
}

// defer_add: the top level's k_scan_add is left to the consumer, which adds scratch[i / SC_TILE] (the fully scanned
// block sums) to data[i] itself -- one launch and one pass over the table less per radix pass
static cudaError_t exclusive_scan_u32(uint32_t *data, uint64_t n, uint32_t *scratch, cudaStream_t st,
                                      unsigned *launches, bool defer_add = false) {
    if (n == 0) return cudaSuccess;
    const unsigned blocks = grid_for(n, SC_TILE);
    k_scan_local<<<blocks, SC_THREADS, 0, st>>>(data, n, (blocks > 1 || defer_add) ? scratch : nullptr);
    ++*launches;
    if (blocks > 1) {
        cudaError_t e = exclusive_scan_u32(scratch, blocks, scratch + blocks, st, launches);
        if (e != cudaSuccess) return e;
        if (!defer_add) {
            k_scan_add<<<blocks, SC_THREADS, 0, st>>>(data, n, scratch);
            ++*launches;
        }
    } else if (defer_add) {
        cudaError_t e = cudaMemsetAsync(scratch, 0, sizeof(uint32_t), st);      // a single block: its base is 0
        if (e != cudaSuccess) return e;
    }
    return cudaGetLastError();
}

// first_hist_done: the caller has already left the first pass's per-tile histogram in s->hist (k_morton_tris)
cudaError_t radix_sort_pairs(xp_scene *s, uint32_t *ka, uint32_t *kb, uint32_t *va, uint32_t *vb,
                             uint64_t n, int bits, bool *in_a, uint32_t *inv_out, bool first_hist_done) {
    *in_a = true;
    if (n == 0) return cudaSuccess;
    const unsigned tiles = grid_for(n, RS_TILE);
    const uint64_t hist_n = (uint64_t)tiles * RS_BINS;
    cudaError_t e = radix_sort_reserve(s, n);
    if (e != cudaSuccess) return e;
    uint32_t *hist = s->hist.as<uint32_t>();
    uint32_t *scratch = hist + hist_n;
    uint32_t *ki = ka, *ko = kb, *vi = va, *vo = vb;
    for (int shift = 0; shift < bits; shift += RS_BITS) {
        if (!(shift == 0 && first_hist_done)) {
            k_rs_hist<<<tiles, RS_THREADS, 0, s->stream>>>(ki, n, shift, hist, tiles);
            ++s->stats.kernel_launches;
        }
        e = exclusive_scan_u32(hist, hist_n, scratch, s->stream, &s->stats.kernel_launches, /*defer_add=*/true);
        if (e != cudaSuccess) return e;
        k_rs_scatter<<<tiles, RS_THREADS, 0, s->stream>>>(ki, vi, ko, vo, n, shift, hist, scratch, tiles,
                                                          shift + RS_BITS >= bits ? inv_out : nullptr);
        ++s->stats.kernel_launches;
        uint32_t *t = ki; ki = ko; ko = t;
        t = vi; vi = vo; vo = t;
        *in_a = !*in_a;
    }
    return cudaGetLastError();
}

// K1 -----------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_pack(const float *__restrict__ aos, uint64_t n,
                                              const uint32_t *__restrict__ sorted_idx,
                                              TriRec *__restrict__ recs, float4 *__restrict__ leaf_box,
                                              uint32_t *__restrict__ degen, TriRec *__restrict__ recs_upload) {
    uint64_t j = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (j >= n) return;
    const uint32_t i = sorted_idx[j];
    V3 v0, v1, v2, lo, hi;
    load_tri(aos, i, v0, v1, v2);
    const TriDerived d = xp_tri_derive(v0, v1, v2);
    xp_tri_aabb(v0, v1, v2, lo, hi);
    TriRec r;
    r.q0 = make_float4(v0.x, v0.y, v0.z, d.n.x);
    r.q1 = make_float4(v1.x, v1.y, v1.z, d.n.y);
    r.q2 = make_float4(v2.x, v2.y, v2.z, d.n.z);
    // the squared lengths of edges (v0,v1) and (v1,v2), op for op what the edge test computes per test
    // (+inf instead of the first one marks a triangle outside the narrowphase's regular domain: literal path)
    r.q3 = make_float4(d.pc, xp_tri_regular(v0, v1, v2) ? xp_edge_len2<ExactOps>(v0, v1) : INFINITY,
                       xp_edge_len2<ExactOps>(v1, v2), __uint_as_float(i));
    store_rec64(recs + j, r);
    if (recs_upload) store_rec64(recs_upload + i, r);          // heightmap scenes: a second copy in upload (cell-major) order for the grid walk
    // Ill-conditioned triangles.  When the edges at v0 are (nearly) parallel -- sin^2 of their angle <= 1e-4 --
    // the reference's normal and the denominator of its point-in-triangle test (triangle.rs:11-32) are
    // dominated by rounding, and its face test can report a "collision" arbitrarily far from the triangle;
    // the broadphase predicate P, which is geometric, would never offer that pair.  Such triangles (and any
    // with non-finite terms) are listed here and swept against every sphere instead (k_sweep_degenerate), as
    // the reference sweeps every triangle.  Exactly collinear ones have a NaN normal and only their geometric
    // vertex / edge tests can hit; listing them too keeps the rule simple.  The second, scale-aware rule lists
    // long thin triangles: the point-in-triangle rounding can accept a face contact about
    // 6e-8 * L^2 / w outside a triangle of length L and width w = |cross| / L; beyond L^2 / w ~ 400 that is no
    // longer small against the 1e-4 by which P inflates the box.
    {
        const V3 e1 = {v1.x - v0.x, v1.y - v0.y, v1.z - v0.z}, e2 = {v2.x - v0.x, v2.y - v0.y, v2.z - v0.z};
        const V3 c = {e1.y * e2.z - e1.z * e2.y, e1.z * e2.x - e1.x * e2.z, e1.x * e2.y - e1.y * e2.x};
        const float cc = c.x * c.x + c.y * c.y + c.z * c.z;
        const float l1 = e1.x * e1.x + e1.y * e1.y + e1.z * e1.z, l2 = e2.x * e2.x + e2.y * e2.y + e2.z * e2.z;
        const float m2 = fmaxf(l1, l2);
        const bool ill = !(cc > 1e-4f * (l1 * l2)) || !(cc * 1.7e5f > m2 * m2 * m2);
        const unsigned act = __activemask(), im = __ballot_sync(act, ill);
        if (im) {                                      // one atomic per warp
            const int lane = threadIdx.x & 31, leader = __ffs(im) - 1;
            unsigned at = 0;
            if (lane == leader) at = atomicAdd(&degen[0], (unsigned)__popc(im));
            at = __shfl_sync(act, at, leader);
            if (ill) degen[1 + at + __popc(im & ((1u << lane) - 1u))] = (uint32_t)j;
        }
    }
    stg256(leaf_box + 2 * j, make_float4(lo.x, lo.y, lo.z, 0.0f), make_float4(hi.x, hi.y, hi.z, 0.0f));   // (lo, hi) side by side: one 32 B sector per box for the tree pass
}

// K4 + K5: hierarchy and boxes in ONE bottom-up pass ----------------------------------------------------------
// Round 1 emitted the hierarchy top-down per node (Karras 2012: two binary searches over the keys per internal node,
// 0.23 ms for 10 M triangles) and then refitted bottom-up through parent pointers (0.87 ms: an atomic counter, a
// fence, four box reads and two box writes per node, 2.1 GB of traffic).  The tree over sorted keys is determined by
// the keys alone, so the bottom-up walk can FIND the parent itself (Apetrei 2014): a subtree covering leaves [l, r]
// hangs under the split right of r or left of l, whichever separates less, i.e. has the smaller
// metric(i) = (key[i] ^ key[i+1], i ^ (i+1)) -- the index part is Karras' tie-break for equal codes, so the topology
// is exactly the old one.  Internal node p is the split between leaves p and p + 1.  A thread writes its box and its
// link straight into its parent's 64 B record (left child: floats 0..5 + link.x, right child: floats 6..11 + link.y),
// swaps its range bound into flags[p], and leaves if it came first; the second arrival reads the sibling's half of the
// record, merges, and climbs.  One atomic per visit, no parent arrays, no per-node box array, no second kernel.
// The root is the split whose subtree covers everything; the walks start at node 0, so k_tree_root_to_zero swaps the
// root's record with node 0's (info[1], info[2]: who links to internal node 0, and on which side).
//
// Measured, the plain form (every visit a global atomic behind a __threadfence) took the same 1.1 ms as the two
// kernels it replaced, and so did a version that let 512-thread blocks climb through shared memory first: a climb is a
// dependent chain, the trees over Morton codes of surfaces are skewed (25 iterations per warp where a balanced tree has
// 9), and a warp keeps running for its last climbing lane -- ncu counted 7.7 of 32 lanes active.  So:
//   * k_tree_local: ONE WARP owns a segment of TREE_SEG consecutive leaves and all of its shared state.  Its lanes pull
//     leaves from the segment as they become free (a lane whose subtree arrived first parks it and takes the next leaf;
//     a lane that arrived second merges and climbs), so the lanes stay busy while the long climbers climb.  A child
//     whose parent split p lies strictly inside the segment attaches through shared memory: exchange of the range bound
//     in s_flag[p], the first arrival parks its box and link, the second builds the node's whole 64 B record and writes
//     it out.  ~90 % of all nodes are finished this way, without a global atomic or fence.
//   * What a segment cannot finish alone -- subtrees whose parent split is the segment's boundary or beyond, and parked
//     children whose sibling reaches over the boundary and so never shows up -- goes onto a work list in global memory.
//   * k_tree_global: one thread per work item continues through the global protocol (half-record store, fence,
//     atomicExch(flags[p]); the second arrival merges and climbs), all items in flight at once.
// A node's two arrivals always use the same protocol: both inside one segment's warp, or both global.
#ifndef XP_TREE_SEG
#define XP_TREE_SEG 256                   /* 10 KB of shared memory per warp: 22 warps per SM; 512: 0.58 ms, 256: 0.48, 128: 0.52 (10 M triangles) */
#endif
static constexpr int TREE_SEG = XP_TREE_SEG;  // leaves per warp
static constexpr unsigned long long METRIC_INF = ~0ull;
struct TreeItem { int l, r, cur; float lo[3], hi[3]; };      // a subtree still to be attached: leaves [l, r], its link, its box

// metric(j) = (key[j] ^ key[j+1], j ^ (j+1)); the ends of the array are walls
__device__ __forceinline__ unsigned long long split_metric(const uint32_t *__restrict__ keys, int n, int j) {
    if (j < 0 || j >= n - 1) return METRIC_INF;
    return ((unsigned long long)(__ldg(keys + j) ^ __ldg(keys + j + 1)) << 32) | (unsigned)(j ^ (j + 1));
}

// the global protocol: attach the subtree (l, r, cur, lo, hi) to its parent and climb while second to arrive
__device__ void tree_climb_global(const uint32_t *__restrict__ keys, int n, Node *nodes, int *flags, int *info, int l, int r,
                                  int cur, float4 lo, float4 hi) {
    for (;;) {
        const bool left_child = split_metric(keys, n, r) < split_metric(keys, n, l - 1);
        const int p = left_child ? r : l - 1;
        float *rec = reinterpret_cast<float *>(nodes + p);
        if (left_child) {
            __stcg(reinterpret_cast<float4 *>(rec), make_float4(lo.x, lo.y, lo.z, hi.x));
            __stcg(reinterpret_cast<float2 *>(rec + 4), make_float2(hi.y, hi.z));
            __stcg(reinterpret_cast<int *>(rec + 12), cur);
            __stcg(reinterpret_cast<int *>(rec + 14), r - l == 1 ? l + 1 : 0);     // leaf run (see Node)
        } else {
            __stcg(reinterpret_cast<float2 *>(rec + 6), make_float2(lo.x, lo.y));
            __stcg(reinterpret_cast<float4 *>(rec + 8), make_float4(lo.z, hi.x, hi.y, hi.z));
            __stcg(reinterpret_cast<int *>(rec + 13), cur);
            __stcg(reinterpret_cast<int *>(rec + 15), r - l == 1 ? l + 1 : 0);
        }
        if (cur == 0) { info[1] = p; info[2] = left_child ? 0 : 1; }
        __threadfence();
        const int other = atomicExch(&flags[p], left_child ? l : r);
        if (other == -1) return;                       // first arrival: the sibling finishes this node
        if (left_child) {                              // the sibling is the right child: floats 6..11, its range ends at `other`
            const float2 a = __ldcg(reinterpret_cast<const float2 *>(rec + 6));
            const float4 b = __ldcg(reinterpret_cast<const float4 *>(rec + 8));
            lo = make_float4(fminf(lo.x, a.x), fminf(lo.y, a.y), fminf(lo.z, b.x), 0.0f);
            hi = make_float4(fmaxf(hi.x, b.y), fmaxf(hi.y, b.z), fmaxf(hi.z, b.w), 0.0f);
            r = other;
        } else {
            const float4 a = __ldcg(reinterpret_cast<const float4 *>(rec));
            const float2 b = __ldcg(reinterpret_cast<const float2 *>(rec + 4));
            lo = make_float4(fminf(lo.x, a.x), fminf(lo.y, a.y), fminf(lo.z, a.z), 0.0f);
            hi = make_float4(fmaxf(hi.x, a.w), fmaxf(hi.y, b.x), fmaxf(hi.z, b.y), 0.0f);
            l = other;
        }
        cur = p;
        if (l == 0 && r == n - 1) { info[0] = p; return; }
    }
}

// work[0 .. cap): the list; info[3]: its length.  A full list (never seen: the capacity is a quarter of the leaves,
// real scenes need 3 - 8 %) makes the lane climb globally right here instead -- slow, still correct.
__device__ __forceinline__ void tree_export(TreeItem *__restrict__ work, unsigned cap, unsigned slot, const uint32_t *__restrict__ keys,
                                            int n, Node *nodes, int *flags, int *info, int l, int r, int cur, float4 lo, float4 hi) {
    if (slot < cap) {
        TreeItem it;
        it.l = l; it.r = r; it.cur = cur;
        it.lo[0] = lo.x; it.lo[1] = lo.y; it.lo[2] = lo.z; it.hi[0] = hi.x; it.hi[1] = hi.y; it.hi[2] = hi.z;
        work[slot] = it;
    } else {
        tree_climb_global(keys, n, nodes, flags, info, l, r, cur, lo, hi);
    }
}

__global__ void __launch_bounds__(32) k_tree_local(const uint32_t *__restrict__ keys, int n, const float4 *__restrict__ leaf_box,
                                                   Node *nodes, int *flags, int *info, TreeItem *__restrict__ work, unsigned work_cap) {
    __shared__ unsigned long long s_metric[TREE_SEG + 1];   // metric(j), j = b0 - 1 .. b1 - 1, at index j - b0 + 1
    __shared__ int s_flag[TREE_SEG];                        // per split p = b0 + k: -1 nobody yet, >= 0 the far range bound of the parked child, -2 finished
    __shared__ float s_box[TREE_SEG][6];                    // the parked child's box
    __shared__ int s_link[TREE_SEG];                        // the parked child's link
    const unsigned lane = threadIdx.x, lt_mask = (1u << lane) - 1u;
    const int b0 = blockIdx.x * TREE_SEG, b1 = min(b0 + TREE_SEG, n), n_seg = b1 - b0;
    for (int k = lane; k <= n_seg; k += 32) s_metric[k] = split_metric(keys, n, b0 - 1 + k);
    for (int k = lane; k < TREE_SEG; k += 32) s_flag[k] = -1;
    __syncwarp();
    int next = 0;                          // next leaf of the segment nobody has taken
    bool carry = false;
    int l = 0, r = 0, cur = 0;
    float4 lo = make_float4(0.f, 0.f, 0.f, 0.f), hi = lo;
    // Leaf boxes are read a chunk of 32 ahead, one per lane, and handed to the lane that takes the leaf by shuffle: a
    // load inside the loop would put a trip to HBM on the critical path of every iteration (measured: 2400 cycles per
    // iteration with 11 warps per SM to hide them).  a: leaves chunk .. chunk + 31 of the segment, b: the next 32.
    int chunk = 0;
    float4 alo = lo, ahi = lo, blo = lo, bhi = lo;
    if ((int)lane < n_seg) { alo = __ldg(leaf_box + 2 * (b0 + lane)); ahi = __ldg(leaf_box + 2 * (b0 + lane) + 1); }
    if (32 + (int)lane < n_seg) { blo = __ldg(leaf_box + 2 * (b0 + 32 + lane)); bhi = __ldg(leaf_box + 2 * (b0 + 32 + lane) + 1); }
    for (;;) {
        // ---- free lanes take fresh leaves (from the current chunk only; the next iteration continues in the next chunk)
        const unsigned idle = __ballot_sync(0xffffffffu, !carry);
        if (idle && next < n_seg) {
            const int lim = min(chunk + 32, n_seg);
            const int mine = next + __popc(idle & lt_mask);
            const int src = (mine - chunk) & 31;
            const float lx = __shfl_sync(0xffffffffu, alo.x, src), ly = __shfl_sync(0xffffffffu, alo.y, src), lz = __shfl_sync(0xffffffffu, alo.z, src);
            const float hx = __shfl_sync(0xffffffffu, ahi.x, src), hy = __shfl_sync(0xffffffffu, ahi.y, src), hz = __shfl_sync(0xffffffffu, ahi.z, src);
            if (!carry && mine < lim) {
                l = r = b0 + mine; cur = ~(b0 + mine);
                lo = make_float4(lx, ly, lz, 0.f); hi = make_float4(hx, hy, hz, 0.f);
                carry = true;
            }
            next = min(next + __popc(idle), lim);
            if (next == chunk + 32 && next < n_seg) {         // chunk used up: the one read ahead becomes current, read the one after
                chunk += 32;
                alo = blo; ahi = bhi;
                const int k = chunk + 32 + (int)lane;
                if (k < n_seg) { blo = __ldg(leaf_box + 2 * (b0 + k)); bhi = __ldg(leaf_box + 2 * (b0 + k) + 1); }
            }
        }
        if (!__any_sync(0xffffffffu, carry)) break;
        // ---- one step of every carrying lane: attach to the parent
        bool left_child = false, local = false;
        int p = 0, other = -1;
        if (carry) {
            left_child = s_metric[r - b0 + 1] < s_metric[l - b0];          // metric(r) < metric(l - 1)
            p = left_child ? r : l - 1;
            local = p >= b0 && p + 1 < b1;
            if (cur == 0) { info[1] = p; info[2] = left_child ? 0 : 1; }
            if (local) other = atomicExch(&s_flag[p - b0], left_child ? l : r);
        }
        // subtrees whose parent split is not inside the segment leave for the work list
        const unsigned out = __ballot_sync(0xffffffffu, carry && !local);
        if (out) {
            unsigned base = 0;
            if (lane == 0) base = atomicAdd(reinterpret_cast<unsigned *>(info + 3), (unsigned)__popc(out));
            base = __shfl_sync(0xffffffffu, base, 0);
            if (carry && !local) {
                tree_export(work, work_cap, base + __popc(out & lt_mask), keys, n, nodes, flags, info, l, r, cur, lo, hi);
                carry = false;
            }
        }
        __syncwarp();
        if (carry && other == -1) {                    // first arrival: park box and link, take a fresh leaf next
            float *q = s_box[p - b0];
            q[0] = lo.x; q[1] = lo.y; q[2] = lo.z; q[3] = hi.x; q[4] = hi.y; q[5] = hi.z;
            s_link[p - b0] = cur;
            carry = false;
        }
        __syncwarp();
        if (carry) {                                   // second arrival: the node is complete
            const float *q = s_box[p - b0];
            const float4 slo = make_float4(q[0], q[1], q[2], 0.f), shi = make_float4(q[3], q[4], q[5], 0.f);
            const int sl = s_link[p - b0];
            s_flag[p - b0] = -2;
            Node nd;
            if (left_child) {
                nd.q0 = make_float4(lo.x, lo.y, lo.z, hi.x); nd.q1 = make_float4(hi.y, hi.z, slo.x, slo.y);
                nd.q2 = make_float4(slo.z, shi.x, shi.y, shi.z);
                nd.link = make_int4(cur, sl, p - l == 1 ? l + 1 : 0, other - p == 2 ? p + 2 : 0);     // children [l, p], [p + 1, other]
                r = other;
            } else {
                nd.q0 = make_float4(slo.x, slo.y, slo.z, shi.x); nd.q1 = make_float4(shi.y, shi.z, lo.x, lo.y);
                nd.q2 = make_float4(lo.z, hi.x, hi.y, hi.z);
                nd.link = make_int4(sl, cur, p - other == 1 ? other + 1 : 0, r - p == 2 ? p + 2 : 0);  // children [other, p], [p + 1, r]
                l = other;
            }
            store_rec64(nodes + p, nd);
            lo = make_float4(fminf(lo.x, slo.x), fminf(lo.y, slo.y), fminf(lo.z, slo.z), 0.f);
            hi = make_float4(fmaxf(hi.x, shi.x), fmaxf(hi.y, shi.y), fmaxf(hi.z, shi.z), 0.f);
            cur = p;
            if (l == 0 && r == n - 1) { info[0] = p; carry = false; }
        }
    }
    __syncwarp();
    // ---- parked children whose sibling never came: it reaches outside the segment, so both go global
    for (int k0 = 0; k0 < n_seg; k0 += 32) {
        const int k = k0 + (int)lane;
        const int pend = k < n_seg ? s_flag[k] : -1;
        const unsigned out = __ballot_sync(0xffffffffu, pend >= 0);
        if (!out) continue;
        unsigned base = 0;
        if (lane == 0) base = atomicAdd(reinterpret_cast<unsigned *>(info + 3), (unsigned)__popc(out));
        base = __shfl_sync(0xffffffffu, base, 0);
        if (pend >= 0) {
            const int p = b0 + k;
            const float *q = s_box[k];
            // the parked child is the left one, [pend, p], if its far bound lies left of the split, else the right one, [p + 1, pend]
            tree_export(work, work_cap, base + __popc(out & lt_mask), keys, n, nodes, flags, info, pend <= p ? pend : p + 1,
                        pend <= p ? p : pend, s_link[k], make_float4(q[0], q[1], q[2], 0.f), make_float4(q[3], q[4], q[5], 0.f));
        }
    }
}

__global__ void __launch_bounds__(256) k_tree_global(const uint32_t *__restrict__ keys, int n, Node *nodes, int *flags, int *info,
                                                     const TreeItem *__restrict__ work, unsigned work_cap) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    unsigned count = *reinterpret_cast<const unsigned *>(info + 3);
    if (count > work_cap) count = work_cap;
    if (i >= count) return;
    const TreeItem it = work[i];
    tree_climb_global(keys, n, nodes, flags, info, it.l, it.r, it.cur, make_float4(it.lo[0], it.lo[1], it.lo[2], 0.f),
                      make_float4(it.hi[0], it.hi[1], it.hi[2], 0.f));
}

__global__ void k_tree_root_to_zero(Node *nodes, const int *__restrict__ info) {
    const int root = info[0];
    if (root == 0) return;
    const Node a = nodes[0], b = nodes[root];
    nodes[0] = b;
    nodes[root] = a;
    const int parent = info[1] == root ? 0 : info[1];            // whoever linked to internal node 0 now links to `root`
    int *link = reinterpret_cast<int *>(&nodes[parent].link);
    link[info[2]] = root;
}

// 32-wide implicit tree: one level per launch.  Thread t owns child slot t of the destination
// level: it copies its source box into the node's SoA planes and the warp (= one node) reduces
// the union box that becomes the node's own entry one level up.  Missing children get the empty
// box (+inf, -inf), which the slab test can never hit.
__global__ void __launch_bounds__(256) k_wide_level(const float4 *__restrict__ src_lo,
                                                    const float4 *__restrict__ src_hi, int src_stride, uint64_t n_src,
                                                    float *__restrict__ planes, float4 *__restrict__ dst_lo,
                                                    float4 *__restrict__ dst_hi, uint64_t n_dst) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n_dst * 32) return;                  // whole warps only: n_dst*32 is a multiple of 32
    const int lane = threadIdx.x & 31;
    float4 lo = make_float4(INFINITY, INFINITY, INFINITY, 0.0f);
    float4 hi = make_float4(-INFINITY, -INFINITY, -INFINITY, 0.0f);
    if (t < n_src) { lo = src_lo[t * src_stride]; hi = src_hi[t * src_stride]; }
    float *node = planes + (t >> 5) * WIDE_NODE_FLOATS;
    node[0 * 32 + lane] = lo.x; node[1 * 32 + lane] = lo.y; node[2 * 32 + lane] = lo.z;
    node[3 * 32 + lane] = hi.x; node[4 * 32 + lane] = hi.y; node[5 * 32 + lane] = hi.z;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo.x = fminf(lo.x, __shfl_xor_sync(0xffffffffu, lo.x, o));
        lo.y = fminf(lo.y, __shfl_xor_sync(0xffffffffu, lo.y, o));
        lo.z = fminf(lo.z, __shfl_xor_sync(0xffffffffu, lo.z, o));
        hi.x = fmaxf(hi.x, __shfl_xor_sync(0xffffffffu, hi.x, o));
        hi.y = fmaxf(hi.y, __shfl_xor_sync(0xffffffffu, hi.y, o));
        hi.z = fmaxf(hi.z, __shfl_xor_sync(0xffffffffu, hi.z, o));
    }
    if (lane == 0) { dst_lo[t >> 5] = lo; dst_hi[t >> 5] = hi; }
}

// Quantised 4-wide collapse: Node4[X] holds the (up to four) grandchildren of binary node X -- the
// children of X's internal children, or X's leaf children themselves.  Only the nodes reachable from
// the root in hops of two levels are ever visited; the others are written and never read.
__device__ __forceinline__ void q4_add(float (&lo)[4][3], float (&hi)[4][3], int (&link)[4], int &cnt,
                                       int l, float lx, float ly, float lz, float hx, float hy, float hz) {
    lo[cnt][0] = lx; lo[cnt][1] = ly; lo[cnt][2] = lz;
    hi[cnt][0] = hx; hi[cnt][1] = hy; hi[cnt][2] = hz;
    link[cnt] = l;
    ++cnt;
}

__global__ void __launch_bounds__(256) k_collapse4(const Node *__restrict__ nodes, int n_internal,
                                                   Node4 *__restrict__ out) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x;
    if (x >= n_internal) return;
    float lo[4][3], hi[4][3];
    int link[4] = {Q4_EMPTY, Q4_EMPTY, Q4_EMPTY, Q4_EMPTY};
    int cnt = 0;
    const Node nx = nodes[x];
    const int kid[2] = {nx.link.x, nx.link.y};
#pragma unroll
    for (int k = 0; k < 2; ++k) {
        if (kid[k] < 0) {                                  // leaf child of X (the empty child of a 1-triangle root has an inverted box)
            if (k == 0) q4_add(lo, hi, link, cnt, kid[k], nx.q0.x, nx.q0.y, nx.q0.z, nx.q0.w, nx.q1.x, nx.q1.y);
            else if (nx.q1.z <= nx.q2.y) q4_add(lo, hi, link, cnt, kid[k], nx.q1.z, nx.q1.w, nx.q2.x, nx.q2.y, nx.q2.z, nx.q2.w);
        } else {
            const Node ny = nodes[kid[k]];
            q4_add(lo, hi, link, cnt, ny.link.x, ny.q0.x, ny.q0.y, ny.q0.z, ny.q0.w, ny.q1.x, ny.q1.y);
            q4_add(lo, hi, link, cnt, ny.link.y, ny.q1.z, ny.q1.w, ny.q2.x, ny.q2.y, ny.q2.z, ny.q2.w);
        }
    }
    float org[3], ext[3];
    unsigned ebits = 0;
    float inv_cell[3];
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        float mn = lo[0][a], mx = hi[0][a];
        for (int c = 1; c < cnt; ++c) { mn = fminf(mn, lo[c][a]); mx = fmaxf(mx, hi[c][a]); }
        org[a] = mn; ext[a] = mx - mn;
        // smallest power-of-two cell with ext <= 252 cells (leaves room for the one-cell outward margins)
        int e;
        frexpf(fmaxf(ext[a], 1e-30f) / 252.0f, &e);        // ext/252 = f * 2^e, f in [0.5,1)  ->  cell = 2^e >= ext/252
        e = max(-126, min(126, e));
        ebits |= (unsigned)(e + 127) << (8 * a);
        inv_cell[a] = __uint_as_float((unsigned)(127 - e) << 23);      // 2^-e, exact
    }
    unsigned qlo[3] = {0, 0, 0}, qhi[3] = {0, 0, 0};
    for (int c = 0; c < 4; ++c) {
#pragma unroll
        for (int a = 0; a < 3; ++a) {
            unsigned l = 255u, h = 0u;                      // empty slot: inverted box, never hit
            if (c < cnt) {
                const float fl = floorf((lo[c][a] - org[a]) * inv_cell[a]) - 1.0f;    // one cell outward
                const float fh = ceilf((hi[c][a] - org[a]) * inv_cell[a]) + 1.0f;
                l = (unsigned)fminf(fmaxf(fl, 0.0f), 255.0f);
                h = (unsigned)fminf(fmaxf(fh, 0.0f), 255.0f);
            }
            qlo[a] |= l << (8 * c);
            qhi[a] |= h << (8 * c);
        }
    }
    // qlo is clamped at 0 = the origin itself, which is <= every child's lo, so clamping stays conservative
    Node4 o;
    o.w0 = make_uint4(__float_as_uint(org[0]), __float_as_uint(org[1]), __float_as_uint(org[2]), ebits);
    o.w1 = make_uint4(qlo[0], qlo[1], qlo[2], qhi[0]);
    o.w2 = make_uint4(qhi[1], qhi[2], (unsigned)link[0], (unsigned)link[1]);
    o.w3 = make_uint4((unsigned)link[2], (unsigned)link[3], 0u, 0u);
    out[x] = o;
}

// a scene with one triangle: root = (leaf 0, empty); the empty box (+inf,-inf) can never be hit
__global__ void k_single_root(Node *nodes, const float4 *leaf_box) {
    const float4 lo = leaf_box[0], hi = leaf_box[1];
    const float inf = INFINITY;
    nodes[0].q0 = make_float4(lo.x, lo.y, lo.z, hi.x);
    nodes[0].q1 = make_float4(hi.y, hi.z, inf, inf);
    nodes[0].q2 = make_float4(inf, -inf, -inf, -inf);
    nodes[0].link = make_int4(~0, ~0, 0, 0);
}

// Row f2 of SURVEY.md section 8: heightmap terrain -> triangle soup ON THE DEVICE.  Emits, for every quad
// (ix, iz) of an (nx+1) x (nz+1) height grid, the two triangles (p00,p01,p11) and (p00,p11,p10) with
// p_ab = (x0 + spacing*(ix+a), h[ix+a][iz+b], z0 + spacing*(iz+b)) -- x outer, z inner: the vertex order of
// low_poly_nice_graphics/src/world/heightmap_loader.rs:19-63 and of the clipmap vertex shader
// (game/src/shaders/shader_clipmap.vert:43-65).  4 B per vertex cross PCIe instead of 36 B per triangle.
__global__ void __launch_bounds__(256) k_emit_heightmap(const float *__restrict__ h, unsigned nx, unsigned nz,
                                                        float x0, float z0, float spacing, float *__restrict__ tris) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q >= (uint64_t)nx * nz) return;
    const unsigned ix = (unsigned)(q / nz), iz = (unsigned)(q % nz);
    const float xa = __fadd_rn(x0, __fmul_rn(spacing, (float)ix)), xb = __fadd_rn(x0, __fmul_rn(spacing, (float)(ix + 1)));
    const float za = __fadd_rn(z0, __fmul_rn(spacing, (float)iz)), zb = __fadd_rn(z0, __fmul_rn(spacing, (float)(iz + 1)));
    const uint64_t row = (uint64_t)nz + 1;
    const float h00 = __ldg(h + ix * row + iz), h01 = __ldg(h + ix * row + iz + 1);
    const float h10 = __ldg(h + (ix + 1) * row + iz), h11 = __ldg(h + (ix + 1) * row + iz + 1);
    float *o = tris + q * 18;
    o[0] = xa; o[1] = h00; o[2] = za;   o[3] = xa; o[4] = h01; o[5] = zb;   o[6] = xb; o[7] = h11; o[8] = zb;
    o[9] = xa; o[10] = h00; o[11] = za; o[12] = xb; o[13] = h11; o[14] = zb; o[15] = xb; o[16] = h10; o[17] = za;
}

// min / max of the height grid (the y extent of the scene box the grid walk clips rays against): ordered-int
// atomics, decoded in place by the second kernel
__global__ void __launch_bounds__(256) k_height_range(const float *__restrict__ h, uint64_t n, unsigned *__restrict__ out) {
    float mn = INFINITY, mx = -INFINITY;
    for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i < n; i += (uint64_t)gridDim.x * blockDim.x) {
        const float v = __ldg(h + i);
        mn = fminf(mn, v); mx = fmaxf(mx, v);
    }
    for (int o = 16; o > 0; o >>= 1) {
        mn = fminf(mn, __shfl_xor_sync(0xffffffffu, mn, o));
        mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
    }
    if ((threadIdx.x & 31) == 0) { atomicMin(&out[0], enc_f(mn)); atomicMax(&out[1], enc_f(mx)); }
}
__global__ void k_height_range_decode(unsigned *io) {
    if (threadIdx.x < 2) reinterpret_cast<float *>(io)[threadIdx.x] = dec_f(io[threadIdx.x]);
}

// mm[(ix, iz)] = (min, max) of h[ix .. ix+1][iz .. iz + W], the heights under the W cells (ix, iz .. iz+W-1): the grid
// walk skips such a window when the ray misses the box of its height range
__global__ void __launch_bounds__(256) k_height_windows(const float *__restrict__ h, unsigned nx, unsigned nz, int window,
                                                        float2 *__restrict__ mm) {
    const uint64_t q = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (q >= (uint64_t)nx * nz) return;
    const unsigned ix = (unsigned)(q / nz), iz = (unsigned)(q % nz);
    const uint64_t row = (uint64_t)nz + 1;
    const unsigned last = min(iz + (unsigned)window, nz);
    float mn = INFINITY, mx = -INFINITY;
    for (unsigned j = iz; j <= last; ++j) {
        const float a = __ldg(h + ix * row + j), b = __ldg(h + (ix + 1) * row + j);
        mn = fminf(mn, fminf(a, b)); mx = fmaxf(mx, fmaxf(a, b));
    }
    mm[q] = make_float2(mn, mx);
}

#define XP_TRY(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)

cudaError_t height_windows(xp_scene *s, const float *d_heights, unsigned nx, unsigned nz) {
    const int window = XP_GW_WINDOW;
    const uint64_t cells = (uint64_t)nx * nz;
    XP_TRY(s->grid_mm.reserve(cells * sizeof(float2)));
    k_height_windows<<<grid_for(cells, 256), 256, 0, s->stream>>>(d_heights, nx, nz, window, s->grid_mm.as<float2>());
    ++s->stats.kernel_launches;
    return cudaGetLastError();
}

cudaError_t height_range(xp_scene *s, const float *d_heights, uint64_t n) {
    XP_TRY(s->grid_yrange.reserve(2 * sizeof(float)));
    unsigned *out = s->grid_yrange.as<unsigned>();
    const unsigned init[2] = {0xFFFFFFFFu, 0u};
    XP_TRY(cudaMemcpyAsync(out, init, sizeof init, cudaMemcpyHostToDevice, s->stream));
    unsigned blocks = grid_for(n, 256);
    if (blocks > (unsigned)s->n_sm * 8) blocks = (unsigned)s->n_sm * 8;
    k_height_range<<<blocks, 256, 0, s->stream>>>(d_heights, n, out);
    k_height_range_decode<<<1, 32, 0, s->stream>>>(out);
    s->stats.kernel_launches += 2;
    return cudaGetLastError();
}

cudaError_t emit_heightmap(xp_scene *s, const float *d_heights, unsigned nx, unsigned nz, float x0, float z0,
                           float spacing) {
    const uint64_t quads = (uint64_t)nx * nz;
    XP_TRY(s->tris_aos.reserve(quads * 2 * sizeof(xp_triangle)));
    k_emit_heightmap<<<grid_for(quads, 256), 256, 0, s->stream>>>(d_heights, nx, nz, x0, z0, spacing,
                                                                  s->tris_aos.as<float>());
    ++s->stats.kernel_launches;
    return cudaGetLastError();
}

// Row f2, clipmap half: the triangles the geometry-clipmap renderer draws, emitted from the per-level
// TOROIDAL height textures exactly like game/src/shaders/shader_clipmap.vert:43-65 places its vertices:
// vertex (ix, iz) of level L = (ix * unit_L, heights[L][iz mod S][ix mod S], iz * unit_L).  A part is one
// create_grid (clipmap.rs:752-786): its quads with x fastest, first every [i0, i2, i1] triangle, then
// every [i3, i1, i2]; parts[] rows are {level, x0, z0, size_x, size_z, first_triangle, triangles, 0}.
__global__ void __launch_bounds__(256) k_emit_clipmap(const float *__restrict__ heights, unsigned size,
                                                      const int4 *__restrict__ parts, unsigned n_parts,
                                                      float unit0, uint64_t n_tris, float *__restrict__ tris) {
    const uint64_t t = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x;
    if (t >= n_tris) return;
    unsigned lo = 0, hi = n_parts - 1;                   // last part whose first triangle is <= t
    while (lo < hi) {
        const unsigned mid = (lo + hi + 1) >> 1;
        if ((uint64_t)(unsigned)__ldg(&parts[2 * mid + 1].y) <= t) lo = mid; else hi = mid - 1;
    }
    const int4 a = __ldg(&parts[2 * lo]), b = __ldg(&parts[2 * lo + 1]);    // a = level, x0, z0, size_x;  b = size_z, first, count, 0
    const int level = a.x, x0 = a.y, z0 = a.z, sx = a.w;
    const unsigned quads = (unsigned)b.z >> 1;
    unsigned q = (unsigned)(t - (unsigned)b.y);
    const bool second = q >= quads;
    if (second) q -= quads;
    const int qx = x0 + (int)(q % (unsigned)(sx - 1)), qz = z0 + (int)(q / (unsigned)(sx - 1));
    // [i0, i2, i1] = (x,z) (x,z+1) (x+1,z);   [i3, i1, i2] = (x+1,z+1) (x+1,z) (x,z+1)
    const int vx[3] = {second ? qx + 1 : qx, second ? qx + 1 : qx, second ? qx : qx + 1};
    const int vz[3] = {second ? qz + 1 : qz, second ? qz : qz + 1, second ? qz + 1 : qz};
    const float unit = __fmul_rn((float)(1u << level), unit0);
    const float *tex = heights + (size_t)level * size * size;
    float *o = tris + t * 9;
#pragma unroll
    for (int k = 0; k < 3; ++k) {
        o[3 * k + 0] = __fmul_rn((float)vx[k], unit);
        o[3 * k + 1] = __ldg(tex + (size_t)((unsigned)vz[k] % size) * size + ((unsigned)vx[k] % size));
        o[3 * k + 2] = __fmul_rn((float)vz[k], unit);
    }
}

#define XP_TRY_(x) do { cudaError_t _e = (x); if (_e != cudaSuccess) return _e; } while (0)
cudaError_t emit_clipmap(xp_scene *s, const float *d_heights, unsigned size, const int *d_parts, unsigned n_parts,
                         float unit0, uint64_t n_tris) {
    XP_TRY_(s->tris_aos.reserve(n_tris * sizeof(xp_triangle)));
    k_emit_clipmap<<<grid_for(n_tris, 256), 256, 0, s->stream>>>(d_heights, size, reinterpret_cast<const int4 *>(d_parts),
                                                                n_parts, unit0, n_tris, s->tris_aos.as<float>());
    ++s->stats.kernel_launches;
    return cudaGetLastError();
}
#undef XP_TRY_

cudaError_t build_lbvh(xp_scene *s) {
    const uint64_t n = s->n_tris;
    s->built = false;
    if (n == 0) { s->built = true; return cudaSuccess; }
    if (n >= (1ull << 27)) return cudaErrorInvalidValue;   // leaf index must fit 27 bits (queue packing)
    cudaStream_t st = s->stream;
    const float *aos = s->tris_aos.as<float>();
    XP_TRY(s->bounds.reserve(8 * sizeof(unsigned)));
    XP_TRY(s->keys_a.reserve(n * 4)); XP_TRY(s->keys_b.reserve(n * 4));
    XP_TRY(s->vals_a.reserve(n * 4)); XP_TRY(s->vals_b.reserve(n * 4));
    XP_TRY(s->recs.reserve(n * sizeof(TriRec)));
    XP_TRY(s->nodes.reserve((n > 1 ? n - 1 : 1) * sizeof(Node)));
    XP_TRY(s->leaf_box.reserve(n * 32));
    XP_TRY(s->refit_flags.reserve((n + 4) * 4));
    unsigned *bounds = s->bounds.as<unsigned>();
    {
        StageTimer t(s, XP_STAGE_BUILD_BOUNDS);
        k_bounds_init<<<1, 32, 0, st>>>(bounds);
        unsigned blocks = grid_for(n, 256);
        if (blocks > (unsigned)s->n_sm * 8) blocks = (unsigned)s->n_sm * 8;
        k_bounds<<<blocks, 256, 0, st>>>(aos, n, bounds);
        s->stats.kernel_launches += 2;
    }
    {
        StageTimer t(s, XP_STAGE_BUILD_MORTON);
        XP_TRY(radix_sort_reserve(s, n));
        k_morton_tris<<<grid_for(n, MT_TILE), 256, 0, st>>>(aos, n, bounds, s->keys_a.as<uint32_t>(),
                                                           s->vals_a.as<uint32_t>(), s->hist.as<uint32_t>(), grid_for(n, MT_TILE));
        ++s->stats.kernel_launches;
    }
    bool in_a = true;
    {
        StageTimer t(s, XP_STAGE_BUILD_SORT);
        XP_TRY(radix_sort_pairs(s, s->keys_a.as<uint32_t>(), s->keys_b.as<uint32_t>(),
                                s->vals_a.as<uint32_t>(), s->vals_b.as<uint32_t>(), n, 30, &in_a, nullptr, /*first_hist_done=*/true));   // 30-bit Morton codes
    }
    if (!in_a) {   // keep the sorted arrays in *_a (scene invariant used by debug_tree and queries)
        xp::DevBuf t = s->keys_a; s->keys_a = s->keys_b; s->keys_b = t;
        t = s->vals_a; s->vals_a = s->vals_b; s->vals_b = t;
    }
    const uint32_t *keys = s->keys_a.as<uint32_t>();
    const uint32_t *vals = s->vals_a.as<uint32_t>();
    {
        StageTimer t(s, XP_STAGE_BUILD_PACK);
        XP_TRY(s->degen.reserve((n + 1) * sizeof(uint32_t)));
        XP_TRY(cudaMemsetAsync(s->degen.p, 0, sizeof(uint32_t), st));
        if (s->grid_valid) XP_TRY(s->recs_grid.reserve(n * sizeof(TriRec)));
        k_pack<<<grid_for(n, 256), 256, 0, st>>>(aos, n, vals, s->recs.as<TriRec>(),
                                                s->leaf_box.as<float4>(), s->degen.as<uint32_t>(),
                                                s->grid_valid ? s->recs_grid.as<TriRec>() : nullptr);
        ++s->stats.kernel_launches;
    }
    if (n == 1) {
        k_single_root<<<1, 1, 0, st>>>(s->nodes.as<Node>(), s->leaf_box.as<float4>());
        ++s->stats.kernel_launches;
    } else {
        StageTimer t(s, XP_STAGE_BUILD_TREE);
        int *flags = s->refit_flags.as<int>();                 // n - 1 arrival slots (-1 = nobody yet) + 4 ints: root, who links to node 0 (node, side), work-list length
        // XP_TREE_WORK_CAP (tests): a tiny list forces the overflow path -- lanes that find the list full climb globally in place
        static const long cap_env = getenv("XP_TREE_WORK_CAP") ? atol(getenv("XP_TREE_WORK_CAP")) : -1;
        const unsigned work_cap = cap_env >= 0 ? (unsigned)cap_env + 1u : (unsigned)(n / 4 + 4096);
        XP_TRY(s->tree_work.reserve((size_t)work_cap * sizeof(TreeItem)));
        XP_TRY(cudaMemsetAsync(flags, 0xFF, n * 4, st));
        XP_TRY(cudaMemsetAsync(flags + n, 0, 4 * 4, st));
        k_tree_local<<<grid_for(n, TREE_SEG), 32, 0, st>>>(keys, (int)n, s->leaf_box.as<float4>(), s->nodes.as<Node>(), flags,
                                                          flags + n, s->tree_work.as<TreeItem>(), work_cap);
        k_tree_global<<<grid_for(work_cap, 256), 256, 0, st>>>(keys, (int)n, s->nodes.as<Node>(), flags, flags + n,
                                                              s->tree_work.as<TreeItem>(), work_cap);
        ++s->stats.kernel_launches;
        k_tree_root_to_zero<<<1, 1, 0, st>>>(s->nodes.as<Node>(), flags + n);
        s->stats.kernel_launches += 2;
    }
    s->q4_built = s->wide_built = false;      // optional structures are (re)built on first use by build_aux()
    XP_TRY(cudaGetLastError());
    s->built = true;
    return cudaSuccess;
}

// Structures only some methods walk (Node4 for Q4_PAIRS, the 32-wide planes for WIDE_*), built from the
// binary tree on first use after a (re)build so that the per-frame rebuild of config 3 pays only for
// what the active method needs.
cudaError_t build_aux(xp_scene *s) {
    if (s->n_tris == 0) return cudaSuccess;
    const uint64_t n = s->n_tris;
    cudaStream_t st = s->stream;
    const bool need_q4 = s->method == XP_METHOD_Q4_PAIRS;
    const bool need_wide = s->method == XP_METHOD_WIDE_FUSED || s->method == XP_METHOD_WIDE_PAIRS;
    if (need_q4 && !s->q4_built) {
        StageTimer t(s, XP_STAGE_BUILD_TREE);
        const uint64_t ni = n > 1 ? n - 1 : 1;
        XP_TRY(s->nodes4.reserve(ni * sizeof(Node4)));
        k_collapse4<<<grid_for(ni, 256), 256, 0, st>>>(s->nodes.as<Node>(), (int)ni, s->nodes4.as<Node4>());
        ++s->stats.kernel_launches;
        s->q4_built = true;
    }
    if (need_wide && !s->wide_built) {
        StageTimer t(s, XP_STAGE_BUILD_TREE);
        uint64_t cnt[WIDE_MAX_LEVELS + 1];
        int levels = 0;
        uint64_t total = 0, m = n;
        do {
            m = (m + 31) / 32;
            cnt[levels] = m;
            s->wide_tree.off[levels] = (unsigned)total;
            total += m;
            ++levels;
        } while (m > 1 && levels < WIDE_MAX_LEVELS);
        if (m > 1) return cudaErrorInvalidValue;
        XP_TRY(s->wide.reserve(total * WIDE_NODE_FLOATS * sizeof(float)));
        XP_TRY(s->wide_lo.reserve(total * 16));
        XP_TRY(s->wide_hi.reserve(total * 16));
        const float4 *src_lo = s->leaf_box.as<float4>(), *src_hi = s->leaf_box.as<float4>() + 1;
        int src_stride = 2;
        uint64_t n_src = n;
        for (int L = 0; L < levels; ++L) {
            float *planes = s->wide.as<float>() + (uint64_t)s->wide_tree.off[L] * WIDE_NODE_FLOATS;
            float4 *dlo = s->wide_lo.as<float4>() + s->wide_tree.off[L];
            float4 *dhi = s->wide_hi.as<float4>() + s->wide_tree.off[L];
            k_wide_level<<<grid_for(cnt[L] * 32, 256), 256, 0, st>>>(src_lo, src_hi, src_stride, n_src, planes, dlo, dhi, cnt[L]);
            ++s->stats.kernel_launches;
            src_lo = dlo; src_hi = dhi; src_stride = 1; n_src = cnt[L];
        }
        s->wide_tree.planes = s->wide.as<float>();
        s->wide_tree.top = levels - 1;
        s->wide_built = true;
    }
    return cudaGetLastError();
}

cudaError_t sphere_morton_order(xp_scene *s, const xp_response *d_in, uint64_t n, uint32_t **order, uint32_t **inv_order,
                                uint32_t *status_for_rays, bool *rays_done) {
    if (rays_done) *rays_done = false;
    StageTimer t(s, XP_STAGE_SPHERE_SORT);
    XP_TRY(s->active_a.reserve(n * 4)); XP_TRY(s->active_b.reserve(n * 4));
    XP_TRY(s->sort_ka.reserve(n * 4)); XP_TRY(s->sort_kb.reserve(n * 4));
    uint32_t *ka = s->sort_ka.as<uint32_t>(), *kb = s->sort_kb.as<uint32_t>();
    uint32_t *va = s->active_a.as<uint32_t>(), *vb = s->active_b.as<uint32_t>();
    // default: 16-bit keys (measured on config 2: the walk is as fast as with 24-bit keys, 0.955 vs 0.950 ms)
    // sorted by counting; XP_RAY_KEY=0/1 select the 24-bit modes (three stable radix passes)
    static const int key_mode = getenv("XP_RAY_KEY") ? atoi(getenv("XP_RAY_KEY")) : 2;
    uint32_t *inv = nullptr;
    if (inv_order) { XP_TRY(s->inv_order.reserve(n * 4)); inv = s->inv_order.as<uint32_t>(); *inv_order = inv; }
    if (key_mode == 2) {                    // counting sort: three launches
        static_assert(CS_GROUPS == 64, "k_bin_scatter scans 64 group totals with one warp");
        XP_TRY(s->hist.reserve((CS_BINS + CS_GROUPS) * sizeof(uint32_t)));
        uint32_t *bins = s->hist.as<uint32_t>();
        XP_TRY(cudaMemsetAsync(bins, 0, CS_BINS * sizeof(uint32_t), s->stream));
        k_morton_spheres<<<grid_for(n, 256), 256, 0, s->stream>>>(d_in, n, s->bounds.as<unsigned>(), ka, va, key_mode, bins);
        k_bin_offsets<<<CS_GROUPS, 1024, 0, s->stream>>>(bins, bins + CS_BINS);
        if (rays_done) {             // the caller sweeps right away: write the ray records from here (s->rays / s->best are reserved)
            RayRec *rays = s->rays.as<RayRec>();
            unsigned long long *best = s->best.as<unsigned long long>();
            const bool fast = s->arith == XP_ARITH_FAST || s->arith == XP_ARITH_FAST_LITERAL;
#if XP_SCATTER_COOP
            XP_TRY(cudaMemsetAsync(best, 0xFF, n * sizeof(unsigned long long), s->stream));      // BEST_NONE everywhere
#endif
            if (fast) k_bin_scatter<FastOps, true><<<grid_for(n, 256), 256, 0, s->stream>>>(ka, n, bins, bins + CS_BINS, va, inv, d_in, rays, best, status_for_rays, s->n_tris > 0);
            else k_bin_scatter<ExactOps, true><<<grid_for(n, 256), 256, 0, s->stream>>>(ka, n, bins, bins + CS_BINS, va, inv, d_in, rays, best, status_for_rays, s->n_tris > 0);
            *rays_done = true;
        } else {
            k_bin_scatter<ExactOps, false><<<grid_for(n, 256), 256, 0, s->stream>>>(ka, n, bins, bins + CS_BINS, va, inv, nullptr, nullptr, nullptr, nullptr, 0);
        }
        s->stats.kernel_launches += 3;
        *order = va;
        return cudaGetLastError();
    }
    k_morton_spheres<<<grid_for(n, 256), 256, 0, s->stream>>>(d_in, n, s->bounds.as<unsigned>(), ka, va, key_mode, nullptr);
    ++s->stats.kernel_launches;
    bool in_a = true;
    XP_TRY(radix_sort_pairs(s, ka, kb, va, vb, n, key_mode == 2 ? 16 : 24, &in_a, inv));
    if (!in_a) { xp::DevBuf tmp = s->active_a; s->active_a = s->active_b; s->active_b = tmp; }
    *order = s->active_a.as<uint32_t>();
    return cudaGetLastError();
}

}  // namespace xp
