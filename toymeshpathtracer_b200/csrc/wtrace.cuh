// wtrace.cuh -- warp-synchronous traversal with DEFERRED triangle tests (device only).
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
    uint32_t cur = active ? sc.rootRef : bvh::NONE;
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

}  // namespace wt
