// trace.cuh -- the ray-traversal state machine used by the persistent kernels (device only).
//
// Same HitScene contract as bvh::traverse (bvh.cuh), restructured for SIMD efficiency on a
// 32-wide warp where every lane owns a different, incoherent ray:
//   * a lane's whole traversal state lives in a `Lane` struct so that a finished lane can be
//     given a NEW ray while its neighbours keep walking (ray regeneration, kernels.cu);
//   * one loop iteration = at most one wide-node step and ONE triangle test ("if-if" with a
//     leaf cursor): lanes at inner nodes never wait for a neighbour to scan a whole leaf --
//     with whole-leaf steps ncu showed 8-9 of 32 lanes active (profiles/);
//   * the near / far slab planes are picked by the ray's octant through the LOAD ADDRESS
//     (rows lo/hi of the SoA node are adjacent), which removes the six min/max per child;
//     the remaining reductions are 3-input FMNMX3 / VIMNMX3 on sm_100a;
//   * the nearest hit child is entered directly, the others are pushed unsorted with their
//     entry distance and culled against the current best t when popped.
// Exactness is untouched: leaves run bvh::mt_exact and the candidate rule is the
// lexicographic minimum of (t, original index).
#pragma once
#include "bvh.cuh"

namespace trc {

constexpr int STACK = 48;

struct Lane {
    // ray
    ex::V3 o, d;
    float idx, idy, idz, ox, oy, oz;  // 1/dir, orig/dir (slab test: t = plane * idir - ox)
    uint32_t sx, sy, sz;              // 1 if the direction component is negative
    float tMin;
    // best candidate
    float t, u, v;
    int id;
    // walk: an inner node to enter (cur) OR a leaf being scanned one triangle per step
    uint32_t cur;
    uint32_t triPos, triEnd;
    int sp;
    bool any;
};

// Returns false if there is nothing to walk (empty scene).
__device__ __forceinline__ bool lane_start(Lane& L, ex::V3 o, ex::V3 d, float tMin, float tMax, bool any, uint32_t rootRef) {
    L.o = o; L.d = d;
    const float dx = bvh::safe_dir(d.x), dy = bvh::safe_dir(d.y), dz = bvh::safe_dir(d.z);
    L.idx = 1.0f / dx; L.idy = 1.0f / dy; L.idz = 1.0f / dz;
    L.ox = o.x * L.idx; L.oy = o.y * L.idy; L.oz = o.z * L.idz;
    L.sx = dx < 0.0f; L.sy = dy < 0.0f; L.sz = dz < 0.0f;
    L.tMin = tMin;
    L.t = tMax; L.u = 0.0f; L.v = 0.0f; L.id = -1;
    L.cur = rootRef;  // the root is always an inner node (build_logic.cuh: emit_single_leaf_root)
    L.triPos = 0; L.triEnd = 0;
    L.sp = 0;
    L.any = any;
    return rootRef != bvh::NONE;
}

__device__ __forceinline__ float f4(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }

__device__ __forceinline__ void enter(Lane& L, uint32_t ref) {
    if (bvh::ref_is_leaf(ref)) { L.triPos = bvh::leaf_first(ref); L.triEnd = L.triPos + (uint32_t)bvh::leaf_count(ref); }
    else L.cur = ref;
}

// One iteration of the walk: at most one wide-node step AND one triangle test.  Returns false
// when the ray is finished.
template <bool STATS>
__device__ __forceinline__ bool lane_step(Lane& L, const bvh::SceneView& sc, uint32_t* stackRef, float* stackT, bvh::TravStats* stats) {
    bool pop = false;
    if (L.triPos == L.triEnd) {
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
                                  fmaxf(__fmaf_rn(f4(nz, k), L.idz, -L.oz), L.tMin));
            const float b = fminf(fminf(__fmaf_rn(f4(fx, k), L.idx, -L.ox), __fmaf_rn(f4(fy, k), L.idy, -L.oy)),
                                  fminf(__fmaf_rn(f4(fz, k), L.idz, -L.oz), L.t));
            ref[k] = __float_as_uint(f4(rf, k));
            tn[k] = a;
            // entry distance with the child slot in its two low mantissa bits: a cheap arg-min
            key[k] = (a <= b && ref[k] != bvh::NONE) ? ((__float_as_uint(a) & ~3u) | (uint32_t)k) : 0xFFFFFFFFu;
        }
        const uint32_t kmin = min(min(key[0], key[1]), min(key[2], key[3]));
        if (kmin == 0xFFFFFFFFu) {
            pop = true;
        } else {
            const int ks = (int)(kmin & 3u);
            enter(L, ks == 0 ? ref[0] : ks == 1 ? ref[1] : ks == 2 ? ref[2] : ref[3]);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (key[k] != 0xFFFFFFFFu && k != ks) {
                    if (L.sp < STACK) { stackRef[L.sp] = ref[k]; stackT[L.sp] = tn[k]; ++L.sp; }
                    else if (sc.status) *sc.status |= bvh::STACK_OVERFLOW;
                }
            }
        }
    }
    if (L.triPos < L.triEnd) {
        if (STATS) ++stats->tris;
        const float4* tp = sc.tris + (size_t)L.triPos * 3;
        const float4 a = __ldg(tp + 0), b = __ldg(tp + 1), c = __ldg(tp + 2);
        float t, u, v;
        // bound by the running best (never above tMax): accept t < best, or equal t and a lower original index
        if (bvh::mt_exact(L.o, L.d, ex::v3(a.x, a.y, a.z), ex::v3(b.x, b.y, b.z), ex::v3(c.x, c.y, c.z), L.tMin, L.t, t, u, v)) {
            const int id = (int)__float_as_uint(a.w);
            if (t < L.t || (L.id >= 0 && id < L.id)) {
                L.t = t; L.id = id; L.u = u; L.v = v;
                if (L.any) return false;
            }
        }
        ++L.triPos;
        pop = L.triPos == L.triEnd;
    }
    if (pop) {
        while (L.sp > 0) {
            --L.sp;
            if (stackT[L.sp] <= L.t) { enter(L, stackRef[L.sp]); return true; }
        }
        return false;
    }
    return true;
}

}  // namespace trc
