// warpq.cuh -- "warp queue" traversal (device only).  EXPERIMENT, selectable with TMPT_HIT_KERNEL=5 for A/B runs;
// bit-exact, but slower than bvh::traverse on B200 -- kept as the measured record of the idea.
//
// Why it exists (ncu, profiles/r1_hit_scene_variants_ncu.txt): with one ray per lane and the leaf test
// inline, the wide-node step ran with ~15 of 32 lanes active but the exact Moller-Trumbore
// code -- 46 % of all warp instructions -- ran with THREE, because only a few lanes stand at
// a leaf in any given iteration.  Here the two kinds of work are decoupled inside the warp:
//
//   * every lane walks the inner nodes of ITS ray (octant-addressed slab test, nearest child
//     first, local-memory stack) and never tests a triangle itself;
//   * a lane that reaches a leaf appends (owner lane, triangle slot) pairs to a per-warp ring
//     in shared memory and keeps walking -- speculatively, against a best-t that may still
//     shrink;
//   * when 32 pairs are queued (or nobody has node work left) the whole warp runs ONE exact
//     triangle test per lane on the queued pairs, reading the owner's ray from shared memory
//     and folding accepted hits into the owner's 64-bit key  (t bits << 32 | original index)
//     with a shared-memory atomicMin -- which IS the reference's candidate rule: nearest t,
//     lowest index among bit-equal t.
//
// A finished lane (no node work, all its pairs consumed) writes its result and is handed a new
// ray (kernels.cu), so the warp stays full until the ray list is exhausted.
#pragma once
#include "bvh.cuh"

namespace wq {

constexpr int QCAP = 256;        // ring capacity (pairs); a node step can add at most 32 * 8
constexpr int STACK = 48;
constexpr unsigned FULL = 0xffffffffu;
constexpr int SLOT_BITS = 27;    // pair = owner lane << 27 | triangle slot

struct WarpShared {
    float ox[32], oy[32], oz[32], dx[32], dy[32], dz[32];  // the 32 rays in flight
    unsigned long long best[32];                           // (t bits << 32) | original triangle index
    uint32_t queue[QCAP];
};

// t >= tMin >= 0 on every accepted hit, so the float's bit pattern orders like the value.
__device__ __forceinline__ unsigned long long make_key(float t, uint32_t id) { return ((unsigned long long)__float_as_uint(t) << 32) | id; }
__device__ __forceinline__ float key_t(unsigned long long k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ int key_id(unsigned long long k) { return (int)(uint32_t)k; }  // 0xFFFFFFFF -> -1

struct Lane {
    float idx, idy, idz, ox, oy, oz;  // 1/dir, orig/dir
    uint32_t sx, sy, sz;
    uint32_t cur;                     // next ref to process (inner node or leaf), NONE = pop needed / nothing
    int sp;
    uint32_t lastTail;                // ring position after this lane's latest enqueue
    bool any;
};

__device__ __forceinline__ float f4(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

// Give `lane` a new ray.  tMax enters through the initial key.
__device__ __forceinline__ void lane_start(Lane& L, WarpShared& ws, int lane, ex::V3 o, ex::V3 d, float tMax, bool any, uint32_t rootRef,
                                           uint32_t tail) {
    const float dx = bvh::safe_dir(d.x), dy = bvh::safe_dir(d.y), dz = bvh::safe_dir(d.z);
    L.idx = 1.0f / dx; L.idy = 1.0f / dy; L.idz = 1.0f / dz;
    L.ox = o.x * L.idx; L.oy = o.y * L.idy; L.oz = o.z * L.idz;
    L.sx = dx < 0.0f; L.sy = dy < 0.0f; L.sz = dz < 0.0f;
    L.cur = bvh::ray_has_nan(o, d) ? bvh::NONE : rootRef;
    L.sp = 0;
    L.lastTail = tail;
    L.any = any;
    ws.ox[lane] = o.x; ws.oy[lane] = o.y; ws.oz[lane] = o.z;
    ws.dx[lane] = d.x; ws.dy[lane] = d.y; ws.dz[lane] = d.z;
    ws.best[lane] = make_key(tMax, 0xFFFFFFFFu);
}

// Pop the next stack entry that the current best t has not culled.
__device__ __forceinline__ uint32_t pop(Lane& L, const uint32_t* stackRef, const float* stackT, float bestT) {
    while (L.sp > 0) {
        --L.sp;
        if (stackT[L.sp] <= bestT) return stackRef[L.sp];
    }
    return bvh::NONE;
}

// One wide-node step for a lane whose cur is an inner node.
template <bool STATS>
__device__ __forceinline__ void node_step(Lane& L, const bvh::SceneView& sc, uint32_t* stackRef, float* stackT, float tMin, float bestT,
                                          bvh::TravStats* stats) {
    if (STATS) ++stats->nodes;
    const float4* n = sc.nodes + (size_t)L.cur * bvh::NODE_F4;
    const float4 nx = __ldg(n + L.sx), fx = __ldg(n + (L.sx ^ 1u));
    const float4 ny = __ldg(n + 2 + L.sy), fy = __ldg(n + 2 + (L.sy ^ 1u));
    const float4 nz = __ldg(n + 4 + L.sz), fz = __ldg(n + 4 + (L.sz ^ 1u));
    const float4 rf = __ldg(n + 6);
    float tn[4];
    uint32_t key[4], ref[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const float a = fmaxf(fmaxf(__fmaf_rn(f4(nx, k), L.idx, -L.ox), __fmaf_rn(f4(ny, k), L.idy, -L.oy)),
                              fmaxf(__fmaf_rn(f4(nz, k), L.idz, -L.oz), tMin));
        const float b = fminf(fminf(__fmaf_rn(f4(fx, k), L.idx, -L.ox), __fmaf_rn(f4(fy, k), L.idy, -L.oy)),
                              fminf(__fmaf_rn(f4(fz, k), L.idz, -L.oz), bestT));
        ref[k] = __float_as_uint(f4(rf, k));
        tn[k] = a;
        key[k] = (a <= b && ref[k] != bvh::NONE) ? ((__float_as_uint(a) & ~3u) | (uint32_t)k) : 0xFFFFFFFFu;
    }
    const uint32_t kmin = min(min(key[0], key[1]), min(key[2], key[3]));
    if (kmin == 0xFFFFFFFFu) {
        L.cur = pop(L, stackRef, stackT, bestT);
        return;
    }
    const int ks = (int)(kmin & 3u);
    L.cur = ks == 0 ? ref[0] : ks == 1 ? ref[1] : ks == 2 ? ref[2] : ref[3];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        if (key[k] != 0xFFFFFFFFu && k != ks) {
            if (L.sp < STACK) { stackRef[L.sp] = ref[k]; stackT[L.sp] = tn[k]; ++L.sp; }
            else if (sc.status) *sc.status |= bvh::STACK_OVERFLOW;
        }
    }
}

// One exact-test pass over up to 32 queued pairs.  Warp-synchronous: all 32 lanes call it.
template <bool STATS>
__device__ __forceinline__ void tri_pass(WarpShared& ws, const bvh::SceneView& sc, int lane, uint32_t& head, uint32_t tail, float tMin, float tMax,
                                         unsigned anyMask, bvh::TravStats* stats) {
    const uint32_t count = tail - head;
    if ((uint32_t)lane < count) {
        if (STATS) ++stats->tris;
        const uint32_t p = ws.queue[(head + (uint32_t)lane) & (QCAP - 1)];
        const int owner = (int)(p >> SLOT_BITS);
        const uint32_t slot = p & ((1u << SLOT_BITS) - 1u);
        const float4* tp = sc.tris + (size_t)slot * 3;
        const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
        const ex::V3 o = ex::v3(ws.ox[owner], ws.oy[owner], ws.oz[owner]);
        const ex::V3 d = ex::v3(ws.dx[owner], ws.dy[owner], ws.dz[owner]);
        const float bound = key_t(ws.best[owner]);  // never above tMax; a stale (larger) value only costs a lost atomicMin
        float t, u, v;
        if (bvh::mt_exact(o, d, ex::v3(a.x, a.y, a.z), ex::v3(b.x, b.y, b.z), ex::v3(c.x, c.y, c.z), tMin, bound, t, u, v) && t < tMax) {
            // an any-hit ray is finished by its first accepted triangle: key 0 culls everything that is left
            const unsigned long long k = ((anyMask >> owner) & 1u) ? 0ull : make_key(t, __float_as_uint(a.w));
            atomicMin(&ws.best[owner], k);
        }
    }
    head += min(count, 32u);
    __syncwarp();
}

// Append the triangles of every lane's leaf (has = lane stands at a leaf ref) to the ring,
// draining it first if they would not fit.  Warp-synchronous.
template <bool STATS>
__device__ __forceinline__ void enqueue_leaves(WarpShared& ws, const bvh::SceneView& sc, int lane, bool has, uint32_t leafRef, uint32_t& head,
                                               uint32_t& tail, uint32_t& lastTail, float tMin, float tMax, unsigned anyMask, bvh::TravStats* stats) {
    const uint32_t cnt = has ? (uint32_t)bvh::leaf_count(leafRef) : 0u;
    uint32_t incl = cnt;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
        const uint32_t y = __shfl_up_sync(FULL, incl, o);
        if (lane >= o) incl += y;
    }
    const uint32_t total = __shfl_sync(FULL, incl, 31);
    while (tail - head + total > (uint32_t)QCAP) tri_pass<STATS>(ws, sc, lane, head, tail, tMin, tMax, anyMask, stats);
    if (has) {
        const uint32_t base = tail + incl - cnt, first = bvh::leaf_first(leafRef);
        for (uint32_t k = 0; k < cnt; ++k) ws.queue[(base + k) & (QCAP - 1)] = ((uint32_t)lane << SLOT_BITS) | (first + k);
        lastTail = base + cnt;
    }
    tail += total;
    __syncwarp();
}

}  // namespace wq
