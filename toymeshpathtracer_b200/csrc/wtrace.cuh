// wtrace.cuh -- warp-synchronous traversal with DEFERRED triangle tests (device only).  EXPERIMENTS, selectable
// with TMPT_HIT_KERNEL=10/11 (deferral thresholds) and 20 (cooperative leaf phase) for A/B runs; bit-exact; the
// per-lane "parked leaf" idea that came out of them is what bvh::traverse now does.
//
// Same HitScene contract as bvh::traverse.  All 32 lanes of a warp call it together, each
// with its own ray (or active = false).  What it changes, and why (ncu, profiles/): with the
// leaf test inline, the exact Moller-Trumbore code -- about half of all warp instructions --
// executes with ~3 of 32 lanes, because few lanes stand at a leaf in the same iteration.
// Here a lane that reaches a leaf parks it (triPos..triEnd) and keeps walking inner nodes; the
// warp runs the triangle block -- ONE triangle per parked lane -- only when at least TRI_MIN
// lanes have one parked, or fewer than WALK_MIN lanes can still walk.  The walk is speculative
// (best t may shrink late), which costs a few extra node visits and never changes a result:
// the candidate rule is still the lexicographic minimum of (t, original index).
#pragma once
#include "bvh.cuh"

namespace wt {

constexpr int STACK = 48;
constexpr unsigned FULL = 0xffffffffu;

__device__ __forceinline__ float f4(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

template <bool STATS, int TRI_MIN, int WALK_MIN>
__device__ __forceinline__ bvh::HitRec traverse_warp(const bvh::SceneView& sc, ex::V3 o, ex::V3 d, float tMin, float tMax, bool any, bool active,
                                                     bvh::TravStats* stats) {
    bvh::HitRec best;
    best.id = -1; best.t = tMax; best.u = 0.0f; best.v = 0.0f;
    const float dx = bvh::safe_dir(d.x), dy = bvh::safe_dir(d.y), dz = bvh::safe_dir(d.z);
    const float idx = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;
    const float ox = o.x * idx, oy = o.y * idy, oz = o.z * idz;
    const uint32_t sx = dx < 0.0f, sy = dy < 0.0f, sz = dz < 0.0f;
    uint32_t stackRef[STACK];
    float stackT[STACK];
    int sp = 0;
    uint32_t cur = active && !bvh::ray_has_nan(o, d) ? sc.rootRef : bvh::NONE;
    uint32_t triPos = 0, triEnd = 0;

    auto pop = [&]() -> uint32_t {
        while (sp > 0) {
            --sp;
            if (stackT[sp] <= best.t) return stackRef[sp];
        }
        return bvh::NONE;
    };

    for (;;) {
        // ---- walk: one wide-node step for every lane that stands at an inner node
        if (cur != bvh::NONE && !bvh::ref_is_leaf(cur)) {
            if (STATS) ++stats->nodes;
            const float4* n = sc.nodes + (size_t)cur * bvh::NODE_F4;
            const float4 nx = __ldg(n + sx), fx = __ldg(n + (sx ^ 1u));
            const float4 ny = __ldg(n + 2 + sy), fy = __ldg(n + 2 + (sy ^ 1u));
            const float4 nz = __ldg(n + 4 + sz), fz = __ldg(n + 4 + (sz ^ 1u));
            const float4 rf = __ldg(n + 6);
            float tn[4];
            uint32_t key[4], ref[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a = fmaxf(fmaxf(__fmaf_rn(f4(nx, k), idx, -ox), __fmaf_rn(f4(ny, k), idy, -oy)), fmaxf(__fmaf_rn(f4(nz, k), idz, -oz), tMin));
                const float b = fminf(fminf(__fmaf_rn(f4(fx, k), idx, -ox), __fmaf_rn(f4(fy, k), idy, -oy)), fminf(__fmaf_rn(f4(fz, k), idz, -oz), best.t));
                ref[k] = __float_as_uint(f4(rf, k));
                tn[k] = a;
                key[k] = (a <= b && ref[k] != bvh::NONE) ? ((__float_as_uint(a) & ~3u) | (uint32_t)k) : 0xFFFFFFFFu;
            }
            const uint32_t kmin = min(min(key[0], key[1]), min(key[2], key[3]));
            if (kmin == 0xFFFFFFFFu) {
                cur = pop();
            } else {
                const int ks = (int)(kmin & 3u);
                cur = ks == 0 ? ref[0] : ks == 1 ? ref[1] : ks == 2 ? ref[2] : ref[3];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (key[k] != 0xFFFFFFFFu && k != ks) {
                        if (sp < STACK) { stackRef[sp] = ref[k]; stackT[sp] = tn[k]; ++sp; }
                        else if (sc.status) *sc.status |= bvh::STACK_OVERFLOW;
                    }
                }
            }
        }
        // ---- a lane that stands at a leaf parks it (if its parking slot is free) and moves on
        if (cur != bvh::NONE && bvh::ref_is_leaf(cur) && triPos == triEnd) {
            triPos = bvh::leaf_first(cur);
            triEnd = triPos + (uint32_t)bvh::leaf_count(cur);
            cur = pop();
        }
        const bool parked = triPos < triEnd;
        const bool canWalk = cur != bvh::NONE && !bvh::ref_is_leaf(cur);
        const unsigned pm = __ballot_sync(FULL, parked), wm = __ballot_sync(FULL, canWalk);
        if ((pm | wm) == 0u) break;
        // ---- exact tests: one triangle per parked lane, when enough lanes take part
        if (pm != 0u && (__popc(pm) >= TRI_MIN || __popc(wm) < WALK_MIN)) {
            if (parked) {
                if (STATS) ++stats->tris;
                const float4* tp = sc.tris + (size_t)triPos * 3;
                const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
                ++triPos;
                float t, u, v;
                if (bvh::mt_exact(o, d, ex::v3(a.x, a.y, a.z), ex::v3(b.x, b.y, b.z), ex::v3(c.x, c.y, c.z), tMin, best.t, t, u, v)) {
                    const int id = (int)__float_as_uint(a.w);
                    if (t < best.t || (best.id >= 0 && id < best.id)) {
                        best.t = t; best.id = id; best.u = u; best.v = v;
                        if (any) { cur = bvh::NONE; sp = 0; triPos = triEnd; }
                    }
                }
            }
        }
    }
    return best;
}

// ---- cooperative leaf phase ---------------------------------------------------------------
// Lanes walk inner nodes on their own; when lanes reach leaves, the WARP tests the triangles:
// up to four leaf owners per pass, eight worker lanes each (one triangle per worker, a leaf
// holds at most eight).  Workers get the owner's ray by shuffle, run the exact test, reduce
// the 64-bit key (t bits << 32 | original index) over their group of eight and hand it back.
// One pass replaces "max leaf size" trips through the ~85-instruction triangle code that the
// inline form executes with 3-4 active lanes (ncu: 40 % of all warp instructions).
__device__ __forceinline__ unsigned long long make_key(float t, uint32_t id) { return ((unsigned long long)__float_as_uint(t) << 32) | id; }
__device__ __forceinline__ float key_t(unsigned long long k) { return __uint_as_float((uint32_t)(k >> 32)); }
__device__ __forceinline__ unsigned long long shfl64(unsigned long long v, int src) {
    const uint32_t lo = __shfl_sync(FULL, (uint32_t)v, src), hi = __shfl_sync(FULL, (uint32_t)(v >> 32), src);
    return ((unsigned long long)hi << 32) | lo;
}
__device__ __forceinline__ unsigned long long shfl64_xor(unsigned long long v, int m) {
    const uint32_t lo = __shfl_xor_sync(FULL, (uint32_t)v, m), hi = __shfl_xor_sync(FULL, (uint32_t)(v >> 32), m);
    return ((unsigned long long)hi << 32) | lo;
}

// Leaf phase for the lanes in `m` (ballot of lanes whose `leafRef` is a leaf).  Returns, for an
// owner lane, the best key found in its leaf (~0 if none); other lanes get ~0.  Warp-synchronous.
template <bool STATS>
__device__ __forceinline__ unsigned long long leaf_phase(const bvh::SceneView& sc, unsigned m, uint32_t leafRef, ex::V3 o, ex::V3 d, float tMin,
                                                         float tMaxStrict, float bound, bvh::TravStats* stats) {
    const int lane = threadIdx.x & 31, g = lane >> 3, w = lane & 7;
    const int rank = __popc(m & ((1u << lane) - 1u));  // my position among the owners
    const bool owner = (m >> lane) & 1u;
    unsigned long long mine = ~0ull;
    unsigned rest = m;
    for (int base = 0; rest != 0u; base += 4) {
        // the next four owners: lowest set bits of `rest`
        const int o0 = __ffs(rest) - 1; rest &= rest - 1;
        const int o1 = rest ? __ffs(rest) - 1 : -1; if (rest) rest &= rest - 1;
        const int o2 = rest ? __ffs(rest) - 1 : -1; if (rest) rest &= rest - 1;
        const int o3 = rest ? __ffs(rest) - 1 : -1; if (rest) rest &= rest - 1;
        const int src = g == 0 ? o0 : g == 1 ? o1 : g == 2 ? o2 : o3;
        const int s = src < 0 ? 0 : src;
        const uint32_t ref = __shfl_sync(FULL, leafRef, s);
        const float rox = __shfl_sync(FULL, o.x, s), roy = __shfl_sync(FULL, o.y, s), roz = __shfl_sync(FULL, o.z, s);
        const float rdx = __shfl_sync(FULL, d.x, s), rdy = __shfl_sync(FULL, d.y, s), rdz = __shfl_sync(FULL, d.z, s);
        const float rb = __shfl_sync(FULL, bound, s);
        unsigned long long key = ~0ull;
        if (src >= 0 && w < bvh::leaf_count(ref)) {
            if (STATS) ++stats->tris;
            const float4* tp = sc.tris + (size_t)(bvh::leaf_first(ref) + (uint32_t)w) * 3;
            const float4 a = bvh::ld_row(tp + 0), b = bvh::ld_row(tp + 1), c = bvh::ld_row(tp + 2);
            float t, u, v;
            if (bvh::mt_exact(ex::v3(rox, roy, roz), ex::v3(rdx, rdy, rdz), ex::v3(a.x, a.y, a.z), ex::v3(b.x, b.y, b.z), ex::v3(c.x, c.y, c.z), tMin, rb,
                              t, u, v) && t < tMaxStrict)
                key = make_key(t, __float_as_uint(a.w));
        }
#pragma unroll
        for (int x = 4; x > 0; x >>= 1) {
            const unsigned long long other = shfl64_xor(key, x);
            key = other < key ? other : key;
        }
        const unsigned long long got = shfl64(key, ((rank - base) & 3) * 8);
        if (owner && rank >= base && rank < base + 4) mine = got;
    }
    return mine;
}

// Traversal with the cooperative leaf phase.  All 32 lanes call it together.  Returns the key
// of the nearest hit ((tMax bits << 32) | 0xFFFFFFFF when there is none); the caller re-runs
// the exact test on the winner for (u, v) -- same operations, same bits.
template <bool STATS>
__device__ __forceinline__ unsigned long long traverse_warp8(const bvh::SceneView& sc, ex::V3 o, ex::V3 d, float tMin, float tMax, bool any, bool active,
                                                             bvh::TravStats* stats) {
    unsigned long long best = make_key(tMax, 0xFFFFFFFFu);
    float bestT = tMax;
    const float dx = bvh::safe_dir(d.x), dy = bvh::safe_dir(d.y), dz = bvh::safe_dir(d.z);
    const float idx = 1.0f / dx, idy = 1.0f / dy, idz = 1.0f / dz;
    const float ox = o.x * idx, oy = o.y * idy, oz = o.z * idz;
    const uint32_t sx = dx < 0.0f, sy = dy < 0.0f, sz = dz < 0.0f;
    unsigned long long stack[STACK];
    int sp = 0;
    bool overflow = false;
    uint32_t cur = active && !bvh::ray_has_nan(o, d) ? sc.rootRef : bvh::NONE;

    for (;;) {
        if (cur != bvh::NONE && !bvh::ref_is_leaf(cur)) {
            if (STATS) ++stats->nodes;
            const float4* n = sc.nodes + (size_t)cur * bvh::NODE_F4;
            const float4 nx = bvh::ld_row(n + sx), fx = bvh::ld_row(n + (sx ^ 1u));
            const float4 ny = bvh::ld_row(n + 2 + sy), fy = bvh::ld_row(n + 2 + (sy ^ 1u));
            const float4 nz = bvh::ld_row(n + 4 + sz), fz = bvh::ld_row(n + 4 + (sz ^ 1u));
            const float4 rf = bvh::ld_row(n + 6);
            uint32_t key[4], ref[4];
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const float a = fmaxf(fmaxf(__fmaf_rn(f4(nx, k), idx, -ox), __fmaf_rn(f4(ny, k), idy, -oy)), fmaxf(__fmaf_rn(f4(nz, k), idz, -oz), tMin));
                const float b = fminf(fminf(__fmaf_rn(f4(fx, k), idx, -ox), __fmaf_rn(f4(fy, k), idy, -oy)), fminf(__fmaf_rn(f4(fz, k), idz, -oz), bestT));
                ref[k] = __float_as_uint(f4(rf, k));
                key[k] = (a <= b && ref[k] != bvh::NONE) ? ((__float_as_uint(a) & ~3u) | (uint32_t)k) : 0xFFFFFFFFu;
            }
            const uint32_t kmin = min(min(key[0], key[1]), min(key[2], key[3]));
            if (kmin == 0xFFFFFFFFu) {
                cur = bvh::NONE;
            } else {
                const uint32_t ks = kmin & 3u;
                cur = ks == 0 ? ref[0] : ks == 1 ? ref[1] : ks == 2 ? ref[2] : ref[3];
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    if (key[k] != 0xFFFFFFFFu && (uint32_t)k != ks) {
                        if (sp < STACK) stack[sp++] = ((unsigned long long)key[k] << 32) | ref[k];
                        else overflow = true;
                    }
                }
            }
        }
        const bool atLeaf = cur != bvh::NONE && bvh::ref_is_leaf(cur);
        const unsigned m = __ballot_sync(FULL, atLeaf);
        if (m != 0u) {
            const unsigned long long got = leaf_phase<STATS>(sc, m, cur, o, d, tMin, tMax, bestT, stats);
            if (atLeaf) {
                if (got < best) {
                    best = got; bestT = key_t(got);
                    if (any) { sp = 0; }
                }
                cur = bvh::NONE;
            }
        }
        while (cur == bvh::NONE && sp > 0) {
            const unsigned long long e = stack[--sp];
            if (__uint_as_float((uint32_t)(e >> 32)) <= bestT) cur = (uint32_t)e;
        }
        if (__ballot_sync(FULL, cur != bvh::NONE) == 0u) break;
    }
    if (overflow && sc.status) *sc.status |= bvh::STACK_OVERFLOW;
    return best;
}

}  // namespace wt
