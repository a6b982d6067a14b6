// bvh.cuh -- HBM layout of the scene and the ray traversal (closest-hit and any-hit).
//
// Replaces the reference's pointer octree + recursive walk (scene.cpp:7-19, 21-52, 86-97)
// and its slab test (maths.h:116-134).  Semantics kept exactly (DESIGN.md "HitScene
// contract"): nearest t over ALL triangles with the reference's own Moller-Trumbore
// arithmetic (maths.cpp:339-380), range tMin <= t <= tMax and t < tMax, ties on bit-equal
// t go to the lowest ORIGINAL triangle index, whatever order the tree is walked in.
//
// Layout (all 16-byte aligned, read with 128-bit loads, L2-resident at every config):
//   nodes : 4-wide BVH, one node = 7 x float4 = 112 bytes, children SoA:
//             [0] lo.x[4] [1] hi.x[4] [2] lo.y[4] [3] hi.y[4] [4] lo.z[4] [5] hi.z[4]
//             [6] child refs[4] (bit-cast)
//           The 112-byte stride is deliberate: every lane of a warp loads the SAME row of a
//           DIFFERENT node, and with a 128-byte stride all those 16-byte pieces sit at the same
//           offset of their L1 line -> same data banks -> one L1 wavefront per lane (ncu:
//           l1tex__data_pipe_lsu_wavefronts at 78 % of peak, profiles/).  At 112 bytes the row
//           offset rotates with the node index and the lanes spread over the banks.
//           child ref: bit 31 clear -> index of an inner node
//                      bit 31 set   -> leaf: bits 28..30 = triCount-1, bits 0..27 = first slot
//           an unused child has ref 0xFFFFFFFF and an inverted box (lo = +3e38, hi = -3e38): tNear <= tFar
//           fails for every ray, so it is never entered and the node step does not look at its ref.
//   tris  : leaf-ordered triangle slots, 3 x float4 each, precomputed Moller-Trumbore form:
//             [0] v0.xyz, original index (bit-cast int)   [1] e1 = v1-v0   [2] e2 = v2-v0
//           (e1, e2 are the same correctly rounded differences maths.cpp:343-344 computes)
//   tris9 : the caller's AoS array, untouched, indexed by ORIGINAL id.
//   hitdata : per ORIGINAL id, 3 x float4: (v0, n.x) (v1, n.y) (v2, n.z) with n = normalize(e1 x e2) computed once
//           at build time by the same operations the reference performs per hit (maths.cpp:375) -- read once per
//           ray that hits, to form Hit.pos / Hit.normal (three aligned loads instead of nine scalar ones).
// Child boxes are the union of PADDED triangle boxes (build_logic.cuh: pad_for) so that the
// float slab test below can never cull a triangle the exact test would accept.
#pragma once
#include "exact.cuh"
#include "sungrid.cuh"


#ifdef __CUDA_ARCH__
#define TMPT_LDG4(p) __ldg(p)
#else
#define TMPT_LDG4(p) (*(p))
#endif

namespace bvh {

constexpr uint32_t LEAF_BIT = 0x80000000u;
constexpr int MAX_LEAF_TRIS = 8;
constexpr int WIDTH = 4;
constexpr int NODE_F4 = 7;              // float4 rows per float node (112 bytes)
constexpr int QNODE_ROWS = 4;           // 16-byte rows per quantised node (64 bytes)
#ifndef TMPT_STACK_SIZE
#define TMPT_STACK_SIZE 128
#endif
constexpr int STACK_SIZE = TMPT_STACK_SIZE;  // entries.  A node step pushes at most 3 and descends one level: a tree of depth D needs 3 D + 4.
constexpr int MAX_TREE_DEPTH = (STACK_SIZE - 4) / 3;  // 41 levels; the SAH builder never exceeds it (build_logic.cuh: sah_must_halve)
constexpr uint32_t NONE = 0xFFFFFFFFu;     // empty child / empty stack; no leaf ref reaches it (slots < 2^28 - 1)
constexpr uint32_t STACK_OVERFLOW = 1;  // bit in the scene's device status word

TMPT_HD uint32_t make_leaf_ref(uint32_t firstSlot, int count) { return LEAF_BIT | (uint32_t(count - 1) << 28) | firstSlot; }
TMPT_HD bool ref_is_leaf(uint32_t r) { return (r & LEAF_BIT) != 0; }
TMPT_HD int leaf_count(uint32_t r) { return int((r >> 28) & 7u) + 1; }
TMPT_HD uint32_t leaf_first(uint32_t r) { return r & 0x0FFFFFFFu; }

struct SceneView {
    const float4* nodes;  // NODE_F4 float4 per node
    const float4* tris;   // 3 float4 per slot
    const float* tris9;   // original triangles
    const float4* hitdata;  // 3 float4 per ORIGINAL triangle: vertices + precomputed normal (may be null: host emulation)
    uint32_t rootRef;     // may itself be a leaf ref for tiny scenes
    int triCount;
    uint32_t* status;     // device status word (STACK_OVERFLOW)
    const uint4* qnodes;  // QNODE_ROWS rows per node: the quantised form of `nodes` (only in a -DTMPT_QNODES=1 build)
    float farLimit;       // rays whose origin has a component beyond this are answered by the all-triangle scan (ray_is_far); <= 0: no limit
    sun::View sun;        // the shadow rays' grid (sungrid.cuh); sun.n == 0: none, shadow rays walk the tree
};

struct HitRec {
    int id;     // original triangle index, -1 = miss
    float t;
    float u, v;  // barycentrics as the exact test computed them
};

// maths.cpp:339-380 on a precomputed (v0, e1, e2) slot.  Returns true iff the reference's
// function would return true for [tMin, tMax]; t/u/v get the reference's bits.


TMPT_HD bool mt_exact(ex::V3 o, ex::V3 d, ex::V3 v0, ex::V3 e1, ex::V3 e2, float tMin, float tMax,
                      float& t, float& u, float& v) {
    const float Epsilon = 1e-5f;
    ex::V3 pvec = ex::cross(d, e2);
    float det = ex::dot(e1, pvec);
    // The determinant test (maths.cpp:350-351) and the u test (:358-359) share ONE exit: the same values in the
    // same order decide, but v0 is needed before the first branch, so the compiler cannot sink the triangle's
    // first row below it and turn one memory round trip into two dependent ones.  (A near-zero det -- under
    // 1 % of the tests -- makes invDet huge or infinite and u garbage; the det clause rejects those.)
    float invDet = ex::rcp(det);
    ex::V3 tvec = ex::sub(o, v0);
    u = ex::mul(ex::dot(tvec, pvec), invDet);
    if ((det > -Epsilon && det < Epsilon) || u < 0.0f || u > 1.0f) return false;
    ex::V3 qvec = ex::cross(tvec, e1);
    v = ex::mul(ex::dot(d, qvec), invDet);
    if (v < 0.0f || ex::add(u, v) > 1.0f) return false;
    t = ex::mul(ex::dot(e2, qvec), invDet);
    return t >= tMin && t <= tMax;
}

// Hit.pos / Hit.normal exactly as maths.cpp:374-375 forms them, from the ORIGINAL vertices.
TMPT_HD ex::V3 tri_normal(ex::V3 v0, ex::V3 v1, ex::V3 v2) { return ex::normalize(ex::cross(ex::sub(v1, v0), ex::sub(v2, v0))); }
TMPT_HD float4 ld_row_payload(const float4* p);  // (defined below, with the other row loads)
TMPT_HD void hit_payload(const SceneView& sc, int id, float u, float v, ex::V3& pos, ex::V3& normal) {
    ex::V3 v0, v1, v2;
    if (sc.hitdata) {
        const float4* h = sc.hitdata + (size_t)id * 3;
        const float4 a = ld_row_payload(h), b = ld_row_payload(h + 1), c = ld_row_payload(h + 2);
        v0 = ex::v3(a.x, a.y, a.z); v1 = ex::v3(b.x, b.y, b.z); v2 = ex::v3(c.x, c.y, c.z);
        normal = ex::v3(a.w, b.w, c.w);
    } else {
        const float* p = sc.tris9 + (size_t)id * 9;
        v0 = ex::v3(p[0], p[1], p[2]); v1 = ex::v3(p[3], p[4], p[5]); v2 = ex::v3(p[6], p[7], p[8]);
        normal = tri_normal(v0, v1, v2);
    }
    float w = ex::sub(ex::sub(1.0f, u), v);
    pos = ex::add(ex::add(ex::muls(v0, w), ex::muls(v1, u)), ex::muls(v2, v));
}

// A zero (or denormal) direction component would make idir infinite and b*inf - o*inf a NaN
// or a wrongly signed infinity; it is replaced by +-1e-20, for which the slab interval is
// "everything" when the origin lies between the planes and empty otherwise -- exactly the
// test an axis-parallel ray needs.
TMPT_HD float safe_dir(float d) { return fabsf(d) < 1.0e-20f ? copysignf(1.0e-20f, d) : d; }

// Work counters of an instrumented pass (bench.py's roofline: box and triangle tests per ray; the rest describes how the
// lanes of a warp spend the iterations of the walk -- DESIGN.md 5).
struct TravStats {
    unsigned long long nodes = 0;      // wide nodes visited (4 box tests each)
    unsigned long long tris = 0;       // exact triangle tests
    unsigned long long iters = 0;      // walk iterations of this lane (one node step and / or one triangle test and / or one pop each)
    unsigned long long culledPops = 0; // pops whose entry the current best t had already culled (the lane sits out the next node step)
    unsigned long long leafWaits = 0;  // iterations in which the lane stood at a leaf while its parked leaf was still being tested
    unsigned long long nodeWarps = 0;  // iterations in which at least one lane of the warp made a node step (counted by one lane)
    unsigned long long triWarps = 0;   // ... at least one lane tested a triangle
    unsigned long long warpIters = 0;  // iterations of the warp (counted by one lane): warpIters * 32 = lane slots issued
    unsigned long long overflows = 0;  // rays whose short stack (SmemStack) overflowed and were traced again
    unsigned long long depthOver[5] = {0, 0, 0, 0, 0};  // rays whose stack held more than 4 / 8 / 12 / 16 / 24 entries at some point
    int maxSp = 0;                     // (scratch: deepest stack of the ray being traced)
};

// ---- traversal stack -------------------------------------------------------------------------------------------------
// 64-bit entries, (entry distance bits << 32) | child ref.  LocalStack is a per-thread array (local memory: L1-cached,
// lane-interleaved); SmemStack is the shared-memory short stack.
struct LocalStack {
    static constexpr bool kCanOverflow = false;
    unsigned long long e[STACK_SIZE];
    // store entry i if `push`; returns whether the stack grew (the builder guarantees i < STACK_SIZE, kernels.cu: build_bvh)
    TMPT_HD bool put_if(bool push, int i, unsigned long long v) {
        if (push) e[i] = v;
        return push;
    }
    TMPT_HD unsigned long long get(int i) const { return *(const volatile unsigned long long*)&e[i]; }  // volatile: issue the load where it is written
    TMPT_HD bool overflowed() const { return false; }
    TMPT_HD void reset() {}
};
// The short stack north_star names: S entries per lane in SHARED memory, entry-major and lane-interleaved ([entry][thread]:
// any mix of stack depths in a warp is bank-conflict free; in local memory every distinct depth is another 128-byte line
// through the L1 data stage, and the stack lines compete with the nodes for L1).  There is no spill path in the walk: a
// push beyond S entries is DROPPED and remembered, and the ray is then traced again with a LocalStack (traverse_with) --
// exactness never depends on S, only the share of rays that pay twice does (tmpt_render_stats reports the depth histogram).
template <int S, int STRIDE>
struct SmemStack {
    static constexpr bool kCanOverflow = true;
    uint32_t col;  // shared-window address of this thread's column: entry i at col + i * STRIDE * 8.  Produced by an opaque
                   // asm move (make()), because the compiler otherwise recomputes it from %tid and the window base at every
                   // access -- five instructions per push -- instead of holding one register.
    bool ovf;
#ifdef __CUDA_ARCH__
    static __device__ __forceinline__ SmemStack make(const void* smemBase, uint32_t thread) {
        uint32_t a = (uint32_t)__cvta_generic_to_shared(smemBase) + thread * 8u, held;
        asm volatile("mov.u32 %0, %1;" : "=r"(held) : "r"(a));
        return SmemStack{held, false};
    }
    __device__ __forceinline__ bool put_if(bool push, int i, unsigned long long v) {
        const bool ok = push && i < S;
        if (ok) asm volatile("st.shared.u64 [%0], %1;" ::"r"(col + (uint32_t)i * (uint32_t)(STRIDE * 8)), "l"(v) : "memory");
        ovf = ovf || (push && !ok);
        return ok;
    }
    __device__ __forceinline__ unsigned long long get(int i) const {
        unsigned long long v;
        asm volatile("ld.shared.u64 %0, [%1];" : "=l"(v) : "r"(col + (uint32_t)i * (uint32_t)(STRIDE * 8)) : "memory");
        return v;
    }
#else
    bool put_if(bool, int, unsigned long long) { return false; }
    unsigned long long get(int) const { return 0; }
#endif
    TMPT_HD bool overflowed() const { return ovf; }
    TMPT_HD void reset() { ovf = false; }
};


// 128-bit loads through the read-only path (ld.global.nc), spelled as asm so that every row is one
// LDG.E.128 exactly where it is written as far as the front end is concerned.  (ptxas still schedules
// them freely -- see mt_exact for how the triangle's three rows are kept together.)
TMPT_HD float4 ld_row(const float4* p) {
#ifdef __CUDA_ARCH__
    float4 v;
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p));
    return v;
#else
    return *p;
#endif
}
// the same load with an L1 eviction-priority hint.  Node rows are re-used (the top of the tree by every ray), triangle rows and the
// hit payload are read once per test / per hit: nodes evict_last, the other two evict_first in L1.  TMPT_CACHE_HINTS is a bit mask:
// 1 node rows evict_last, 2 triangle rows evict_first, 4 triangle rows no_allocate, 8 payload rows evict_first, 16 payload rows
// no_allocate, 32 the evict_first loads carry an L2 evict_last policy.  Measured on the headline frame (profiles/r2_tuning_sweeps.txt):
// 0: 5219-5238 Mrays/s, 0.06 GB of DRAM reads per frame | 3: 5273, 0.90 GB | 11: 5277, 2.54 GB | 43 (the default): 5279, 0.06 GB.
#ifndef TMPT_CACHE_HINTS
#define TMPT_CACHE_HINTS 43
#endif
#ifdef __CUDA_ARCH__
#define TMPT_LD_HINT(name, hint)                                                                                                      \
    __device__ __forceinline__ float4 name(const float4* p) {                                                                        \
        float4 v;                                                                                                                     \
        asm volatile("ld.global.nc." hint ".v4.f32 {%0,%1,%2,%3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p)); \
        return v;                                                                                                                     \
    }
TMPT_LD_HINT(ld_row_evict_last, "L1::evict_last")
#if TMPT_CACHE_HINTS & 32
// evict_first in L1 only: without an L2 policy the line is the first to leave L2 as well, and the 3 MB of triangle rows and hit payload are
// re-read from DRAM all frame long (ncu: dram__bytes_read 0.06 -> 2.5 GB per frame).  The L2 cache-hint operand keeps them resident there.
__device__ __forceinline__ float4 ld_row_evict_first(const float4* p) {
    float4 v;
    unsigned long long pol;
    asm("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    asm volatile("ld.global.nc.L1::evict_first.L2::cache_hint.v4.f32 {%0,%1,%2,%3}, [%4], %5;" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "l"(pol));
    return v;
}
#else
TMPT_LD_HINT(ld_row_evict_first, "L1::evict_first")
#endif
TMPT_LD_HINT(ld_row_no_allocate, "L1::no_allocate")
#undef TMPT_LD_HINT
#endif
TMPT_HD float4 ld_row_node(const float4* p) {
#if defined(__CUDA_ARCH__) && (TMPT_CACHE_HINTS & 1)
    return ld_row_evict_last(p);
#else
    return ld_row(p);
#endif
}
TMPT_HD float4 ld_row_tri(const float4* p) {
#if defined(__CUDA_ARCH__) && (TMPT_CACHE_HINTS & 2)
    return ld_row_evict_first(p);
#elif defined(__CUDA_ARCH__) && (TMPT_CACHE_HINTS & 4)
    return ld_row_no_allocate(p);
#else
    return ld_row(p);
#endif
}
TMPT_HD float4 ld_row_payload(const float4* p) {
#if defined(__CUDA_ARCH__) && (TMPT_CACHE_HINTS & 8)
    return ld_row_evict_first(p);
#elif defined(__CUDA_ARCH__) && (TMPT_CACHE_HINTS & 16)
    return ld_row_no_allocate(p);
#else
    return *p;
#endif
}
TMPT_HD float f4c(const float4& v, int k) { return k == 0 ? v.x : k == 1 ? v.y : k == 2 ? v.z : v.w; }
TMPT_HD float fmaf_(float a, float b, float c) {
#ifdef __CUDA_ARCH__
    return __fmaf_rn(a, b, c);
#elif defined(TMPT_EMU_FMA)
    return fmaf(a, b, c);  // host emulation built to round the slab distances exactly as the device does (tests/emu, tag "fma")
#else
    return a * b + c;  // conservativeness does not depend on fusing
#endif
}

// A ray with a NaN component can never be accepted by the exact test (det, u, v or t is NaN and
// the last comparison t >= tMin fails), so it is a miss by definition -- but fminf/fmaxf DROP
// NaN operands, which would make every slab test pass and the walk visit the whole tree
// (measured: one such ray = 37 ms; the reference's own normalize(0) scatter, SURVEY.md 0.7,
// produces them).  Traversal starts from an empty root for these rays.
TMPT_HD bool ray_has_nan(ex::V3 o, ex::V3 d) { return o.x != o.x || o.y != o.y || o.z != o.z || d.x != d.x || d.y != d.y || d.z != d.z; }

// The boxes are padded for rays that start within a few scene sizes of the scene (build_logic.cuh: pad_for): the slab
// arithmetic's rounding error grows with |origin| (t = plane / d - origin / d cancels two numbers of that size), the padding does
// not.  A ray that starts further out than farLimit = 16 x the scene's largest |coordinate| -- 2.5x inside what the padding
// covers -- is therefore answered by the all-triangle scan, which is exact by definition; the camera and every hit point of a
// render are far inside the limit, so only tmpt_hit_scene callers with distant origins ever pay for it.
TMPT_HD bool ray_is_far(const SceneView& sc, ex::V3 o) {
    return sc.farLimit > 0.0f && fmaxf(fmaxf(fabsf(o.x), fabsf(o.y)), fabsf(o.z)) > sc.farLimit;
}

// One traversal, closest (ANY=false) or any-hit (ANY=true).
//
// Node step: the near / far slab planes are picked by the ray's octant through the LOAD
// ADDRESS (rows lo/hi of an axis are adjacent), so a child costs 6 FMA + two 3-input min/max
// (FMNMX3 on sm_100a).  The nearest hit child is found by a 4-way integer min over
// (entry-distance bits with the child slot in the two low mantissa bits) and entered
// directly; the other hit children are pushed unsorted, branch-free, as 64-bit
// (entry distance, ref) pairs and culled against the current best t when popped.
// Culling keeps a child when tNear <= tFar with tFar clipped to the CURRENT best t -- "<=",
// not "<", so that a triangle in another leaf with bit-equal t and a lower index is still
// tested.  The candidate rule is the lexicographic minimum of (t, id).
// Ray constants of the slab test and the walk state of one lane.
struct RayCtx {
    float idx, idy, idz, ox, oy, oz;  // 1 / dir, orig / dir  (t = plane * idir - ox)
    uint32_t sx, sy, sz;              // 1 where the direction component is negative: row offset of the near plane
};
TMPT_HD RayCtx make_ray_ctx(ex::V3 o, ex::V3 d) {
    RayCtx r;
    const float sdx = safe_dir(d.x), sdy = safe_dir(d.y), sdz = safe_dir(d.z);
    r.idx = 1.0f / sdx; r.idy = 1.0f / sdy; r.idz = 1.0f / sdz;
    r.ox = o.x * r.idx; r.oy = o.y * r.idy; r.oz = o.z * r.idz;
    r.sx = sdx < 0.0f ? 1u : 0u; r.sy = sdy < 0.0f ? 1u : 0u; r.sz = sdz < 0.0f ? 1u : 0u;
    return r;
}

// One 4-wide node step: enter the nearest hit child (returned; NONE if no child is hit), push the others.
// The stack cannot overflow: a step pushes at most three entries and descends one level, and scene creation
// refuses trees deeper than (STACK_SIZE - 4) / 3 levels (kernels.cu).  Rows are addressed with 32-bit row indices
// (node * 7 + row < 2^32), which keeps the address arithmetic to one IMAD.WIDE per load.
//
// Child order.  Closest-hit rays enter the NEAREST hit child first (what makes best t shrink early).  Any-hit rays
// only need some occluder, so they enter the child the ray LEAVES LAST first -- the one with the largest exit distance,
// i.e. the far and the big boxes, where an occluder is most likely: on the four test scenes this never costs more and
// saves 12 % of all node + triangle rows on the Sponza stand-in (24 % of the shadow rays' own work); the answer (a
// boolean) cannot depend on the order.
// The part of a node step that follows the four slab tests: a[k] / b[k] = entry / exit distance of child k (already
// clipped to [tMin, bestT]), ref[k] its reference.
template <class Stack>
TMPT_HD uint32_t enter_and_push(const float (&a)[4], const float (&b)[4], const uint32_t (&ref)[4], Stack& stack, int& sp, bool anyRay) {
    uint32_t key[4], okey[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        // stack key = entry distance (for the pop-time cull) with the child slot in its two low mantissa bits; clearing those
        // bits only lowers the distance: still conservative.  (An empty child has an inverted box and can never pass a <= b.)
        key[k] = (a[k] <= b[k]) ? ((ex::f2u(a[k]) & ~3u) | (uint32_t)k) : 0xFFFFFFFFu;
        // order key: smallest wins.  b >= tMin >= 0 for a hit child, so its bit pattern orders like its value.
        okey[k] = !anyRay ? key[k] : (a[k] <= b[k]) ? (((0x7F800000u - ex::f2u(b[k])) & ~3u) | (uint32_t)k) : 0xFFFFFFFFu;
    }
    const uint32_t k01 = okey[0] < okey[1] ? okey[0] : okey[1], k23 = okey[2] < okey[3] ? okey[2] : okey[3];
    const uint32_t kmin = k01 < k23 ? k01 : k23;
    // no early exit for "no child hit": straight-line code keeps the refs row in the same load batch as the boxes
    // (with a branch here the compiler sinks that load below it: a second, dependent round trip per step)
    const uint32_t ks = kmin & 3u;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const bool push = key[k] != 0xFFFFFFFFu && (uint32_t)k != ks;  // (all keys empty -> nothing is pushed)
        sp += stack.put_if(push, sp, ((unsigned long long)key[k] << 32) | ref[k]) ? 1 : 0;
    }
    // a two-level select: the chain "ks == 0 ? .. : ks == 1 ? .." compiled to branches and a reconvergence point (+2.2 %)
    const uint32_t r01 = (ks & 1u) ? ref[1] : ref[0], r23 = (ks & 1u) ? ref[3] : ref[2];
    const uint32_t nearest = (ks & 2u) ? r23 : r01;
    return kmin == 0xFFFFFFFFu ? NONE : nearest;
}

// Blackwell's packed FP32 pair instructions (FFMA2: two fused multiply-adds per lane in ONE issue slot; the scalar operands are
// broadcast by the instruction itself).  The walk is bound by instruction issue, not by the FMA pipe (28 % busy), so pairing the
// 24 slab FMAs of a node step frees 12 issue slots.  Each half is the same correctly rounded fma as the scalar form: the
// distances, and therefore the walk, are bit-identical.
#ifndef TMPT_FMA2
#define TMPT_FMA2 1   // measured: +3.9 % on the headline frame (5043 -> 5240 Mrays/s, same bytes), profiles/r2_tuning_sweeps.txt
#endif
#if defined(__CUDA_ARCH__) && TMPT_FMA2
__device__ __forceinline__ void fma2_bcast(float a0, float a1, float m, float c, float& r0, float& r1) {
    unsigned long long a, mm, cc, r;
    asm("mov.b64 %0, {%1,%2};" : "=l"(a) : "f"(a0), "f"(a1));
    asm("mov.b64 %0, {%1,%1};" : "=l"(mm) : "f"(m));
    asm("mov.b64 %0, {%1,%1};" : "=l"(cc) : "f"(c));
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(mm), "l"(cc));
    asm("mov.b64 {%0,%1}, %2;" : "=f"(r0), "=f"(r1) : "l"(r));
}
#endif

#ifndef TMPT_NODE_ADDR
#define TMPT_NODE_ADDR 1
#endif
#if defined(__CUDA_ARCH__)
// node row at p + OFF bytes (OFF rides in the load instruction), hinted like ld_row_node
template <int OFF>
__device__ __forceinline__ float4 ld_node_at(const float4* p) {
    float4 v;
#if TMPT_CACHE_HINTS & 1
    asm volatile("ld.global.nc.L1::evict_last.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "n"(OFF));
#else
    asm volatile("ld.global.nc.v4.f32 {%0,%1,%2,%3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p), "n"(OFF));
#endif
    return v;
}
#endif

template <class Stack>
TMPT_HD uint32_t wide_node_step(const SceneView& sc, uint32_t node, const RayCtx& r, float tMin, float bestT, Stack& stack, int& sp,
                                bool anyRay) {
#if defined(__CUDA_ARCH__) && TMPT_NODE_ADDR
    // Row addresses with the constant part in the load's immediate offset and the far row at "minus the sign": near row of axis k at
    // row (7 node + s) + 2k, far row at (7 node - s) + 2k + 1.  No "+ 2k" adds and no "s ^ 1": 14 address instructions per node step
    // instead of 22, 3-4 of them on the ALU pipe instead of 8.
    const int i0 = (int)(node * (uint32_t)NODE_F4);
    const float4 nx = ld_node_at<0>(sc.nodes + (i0 + (int)r.sx)), fx = ld_node_at<16>(sc.nodes + (i0 - (int)r.sx));
    const float4 ny = ld_node_at<32>(sc.nodes + (i0 + (int)r.sy)), fy = ld_node_at<48>(sc.nodes + (i0 - (int)r.sy));
    const float4 nz = ld_node_at<64>(sc.nodes + (i0 + (int)r.sz)), fz = ld_node_at<80>(sc.nodes + (i0 - (int)r.sz));
    const float4 rf = ld_node_at<96>(sc.nodes + i0);
#else
    const uint32_t row0 = node * (uint32_t)NODE_F4;
    const float4 nx = ld_row_node(sc.nodes + (row0 + r.sx)), fx = ld_row_node(sc.nodes + (row0 + (r.sx ^ 1u)));
    const float4 ny = ld_row_node(sc.nodes + (row0 + 2u + r.sy)), fy = ld_row_node(sc.nodes + (row0 + 2u + (r.sy ^ 1u)));
    const float4 nz = ld_row_node(sc.nodes + (row0 + 4u + r.sz)), fz = ld_row_node(sc.nodes + (row0 + 4u + (r.sz ^ 1u)));
    const float4 rf = ld_row_node(sc.nodes + (row0 + 6u));
#endif
    float a[4], b[4];
    uint32_t ref[4];
#if defined(__CUDA_ARCH__) && TMPT_FMA2
    float tnx[4], tny[4], tnz[4], tfx[4], tfy[4], tfz[4];
    fma2_bcast(nx.x, nx.y, r.idx, -r.ox, tnx[0], tnx[1]); fma2_bcast(nx.z, nx.w, r.idx, -r.ox, tnx[2], tnx[3]);
    fma2_bcast(ny.x, ny.y, r.idy, -r.oy, tny[0], tny[1]); fma2_bcast(ny.z, ny.w, r.idy, -r.oy, tny[2], tny[3]);
    fma2_bcast(nz.x, nz.y, r.idz, -r.oz, tnz[0], tnz[1]); fma2_bcast(nz.z, nz.w, r.idz, -r.oz, tnz[2], tnz[3]);
    fma2_bcast(fx.x, fx.y, r.idx, -r.ox, tfx[0], tfx[1]); fma2_bcast(fx.z, fx.w, r.idx, -r.ox, tfx[2], tfx[3]);
    fma2_bcast(fy.x, fy.y, r.idy, -r.oy, tfy[0], tfy[1]); fma2_bcast(fy.z, fy.w, r.idy, -r.oy, tfy[2], tfy[3]);
    fma2_bcast(fz.x, fz.y, r.idz, -r.oz, tfz[0], tfz[1]); fma2_bcast(fz.z, fz.w, r.idz, -r.oz, tfz[2], tfz[3]);
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = fmaxf(fmaxf(tnx[k], tny[k]), fmaxf(tnz[k], tMin));
        b[k] = fminf(fminf(tfx[k], tfy[k]), fminf(tfz[k], bestT));
        ref[k] = ex::f2u(f4c(rf, k));
    }
#else
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        a[k] = fmaxf(fmaxf(fmaf_(f4c(nx, k), r.idx, -r.ox), fmaf_(f4c(ny, k), r.idy, -r.oy)), fmaxf(fmaf_(f4c(nz, k), r.idz, -r.oz), tMin));
        b[k] = fminf(fminf(fmaf_(f4c(fx, k), r.idx, -r.ox), fmaf_(f4c(fy, k), r.idy, -r.oy)), fminf(fmaf_(f4c(fz, k), r.idz, -r.oz), bestT));
        ref[k] = ex::f2u(f4c(rf, k));
    }
#endif
    return enter_and_push(a, b, ref, stack, sp, anyRay);
}

// ---- quantised nodes (compile-time -DTMPT_QNODES=1; measured 8-14 % slower than the float nodes on B200, NOT the default,
// DESIGN.md 5): 64 bytes = 4 rows instead of 7 ------------------------------------------------------------------------------
// The walk is bound by the L1 data stage -- 16-byte rows gathered per ray -- so a node is stored the way
// Ylitie, Karras & Laine 2017 ("Efficient incoherent ray traversal on GPUs through compressed wide BVHs") store theirs:
// child planes as 8-bit offsets on a per-node grid, plane = origin + q * 2^e per axis.
//   row 0: origin.x origin.y origin.z | E.x E.y E.z 0   (E = biased float exponent of 2^(e+15), one byte per axis)
//   row 1: child refs[4]              (as in the float node; an empty child refers to triangle slot 0 and has an inverted box)
//   row 2: lo.x[4] hi.x[4] lo.y[4] hi.y[4]   (one byte per child)
//   row 3: lo.z[4] hi.z[4] 0x3F800000 0
// lo bytes are rounded down, hi bytes up, each by an extra 1/64 step (build_logic.cuh: quantize_node), so the decoded box
// contains the float box; the children only grow by up to a 255th of the node's extent per side.
// Decode: one PRMT drops the byte into the mantissa of 1.0f, v = 1 + q / 32768 (exact), and the plane's ray distance is one
// FFMA, t = v * A + C, with per-node  A = 2^(e+15) / d  and  C = (origin - o) / d - A.  Forming C cancels at most 128 node
// extents against each other: an error of 2^-17 of the extent = 0.002 grid steps, inside the 1/64 step margin.
// The rows of node n sit at n*4 + (r ^ (n>>1 & 3)): with a plain 64-byte stride every lane of a warp would read the same
// two 16-byte columns of its L1 line (the bank conflict that made 128-byte float nodes slow, see above).
#ifndef TMPT_QSTRIDE
#define TMPT_QSTRIDE 4
#endif
constexpr int QNODE_STRIDE = TMPT_QSTRIDE;  // rows from one node to the next: 4 (skewed as described) or 5 (80 bytes, one row unused, no skew needed)
TMPT_HD uint32_t qnode_row(uint32_t node, uint32_t r) {
    return QNODE_STRIDE == 4 ? ((node << 2) ^ ((node >> 1) & 3u) ^ r) : node * (uint32_t)QNODE_STRIDE + r;
}
TMPT_HD uint4 ld_rowu(const uint4* p) {
#ifdef __CUDA_ARCH__
    uint4 v;
    asm volatile("ld.global.nc.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p));
    return v;
#else
    return *p;
#endif
}
// `one` = 0x3F800000 held in a REGISTER: PRMT has a single immediate slot, and with the constant in it the selector
// would need a register (and a MOV) per use.
template <int K>
TMPT_HD float qbyte_f(uint32_t w, uint32_t one) {  // 1 + (byte K of w) / 32768
#ifdef __CUDA_ARCH__
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(w), "r"(one), "n"(0x7604 | (K << 4)));
    return __uint_as_float(d);
#else
    return ex::u2f(one | (((w >> (8 * K)) & 0xFFu) << 8));
#endif
}
template <int K>
TMPT_HD void qchild(uint32_t one, uint32_t nxw, uint32_t nyw, uint32_t nzw, uint32_t fxw, uint32_t fyw, uint32_t fzw, float ax, float ay, float az,
                    float cx, float cy, float cz, float tMin, float bestT, float& a, float& b) {
    a = fmaxf(fmaxf(fmaf_(qbyte_f<K>(nxw, one), ax, cx), fmaf_(qbyte_f<K>(nyw, one), ay, cy)), fmaxf(fmaf_(qbyte_f<K>(nzw, one), az, cz), tMin));
    b = fminf(fminf(fmaf_(qbyte_f<K>(fxw, one), ax, cx), fmaf_(qbyte_f<K>(fyw, one), ay, cy)), fminf(fmaf_(qbyte_f<K>(fzw, one), az, cz), bestT));
}
template <class Stack>
TMPT_HD uint32_t qnode_step(const SceneView& sc, uint32_t node, const RayCtx& r, float tMin, float bestT, Stack& stack, int& sp,
                            bool anyRay) {
    uint4 h, rf, qa, qb;
    if (QNODE_STRIDE == 4) {
        const uint32_t base = (node << 2) ^ ((node >> 1) & 3u);
        h = ld_rowu(sc.qnodes + base); rf = ld_rowu(sc.qnodes + (base ^ 1u));
        qa = ld_rowu(sc.qnodes + (base ^ 2u)); qb = ld_rowu(sc.qnodes + (base ^ 3u));
    } else {
        const uint4* p = sc.qnodes + node * (uint32_t)QNODE_STRIDE;
        h = ld_rowu(p); rf = ld_rowu(p + 1); qa = ld_rowu(p + 2); qb = ld_rowu(p + 3);
    }
    const float ax = ex::u2f((h.w << 23) & 0x7F800000u) * r.idx, ay = ex::u2f((h.w << 15) & 0x7F800000u) * r.idy,
                az = ex::u2f((h.w << 7) & 0x7F800000u) * r.idz;
    const float cx = fmaf_(ex::u2f(h.x), r.idx, -r.ox) - ax, cy = fmaf_(ex::u2f(h.y), r.idy, -r.oy) - ay, cz = fmaf_(ex::u2f(h.z), r.idz, -r.oz) - az;
    const uint32_t nxw = r.sx ? qa.y : qa.x, fxw = r.sx ? qa.x : qa.y;
    const uint32_t nyw = r.sy ? qa.w : qa.z, fyw = r.sy ? qa.z : qa.w;
    const uint32_t nzw = r.sz ? qb.y : qb.x, fzw = r.sz ? qb.x : qb.y;
    const uint32_t one = qb.z;  // 0x3F800000, stored in the node so that it arrives in a register (see qbyte_f)
    float a[4], b[4];
#if defined(__CUDA_ARCH__) && TMPT_FMA2
    {  // the 24 plane distances as 12 FFMA2, like wide_node_step
        float tn[3][4], tf[3][4];
#define TMPT_QPAIR(T, W, A, C)                                                           \
        fma2_bcast(qbyte_f<0>(W, one), qbyte_f<1>(W, one), A, C, T[0], T[1]);            \
        fma2_bcast(qbyte_f<2>(W, one), qbyte_f<3>(W, one), A, C, T[2], T[3]);
        TMPT_QPAIR(tn[0], nxw, ax, cx) TMPT_QPAIR(tn[1], nyw, ay, cy) TMPT_QPAIR(tn[2], nzw, az, cz)
        TMPT_QPAIR(tf[0], fxw, ax, cx) TMPT_QPAIR(tf[1], fyw, ay, cy) TMPT_QPAIR(tf[2], fzw, az, cz)
#undef TMPT_QPAIR
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            a[k] = fmaxf(fmaxf(tn[0][k], tn[1][k]), fmaxf(tn[2][k], tMin));
            b[k] = fminf(fminf(tf[0][k], tf[1][k]), fminf(tf[2][k], bestT));
        }
    }
    const uint32_t ref2[4] = {rf.x, rf.y, rf.z, rf.w};
    if (true) return enter_and_push(a, b, ref2, stack, sp, anyRay);
#endif
    qchild<0>(one, nxw, nyw, nzw, fxw, fyw, fzw, ax, ay, az, cx, cy, cz, tMin, bestT, a[0], b[0]);
    qchild<1>(one, nxw, nyw, nzw, fxw, fyw, fzw, ax, ay, az, cx, cy, cz, tMin, bestT, a[1], b[1]);
    qchild<2>(one, nxw, nyw, nzw, fxw, fyw, fzw, ax, ay, az, cx, cy, cz, tMin, bestT, a[2], b[2]);
    qchild<3>(one, nxw, nyw, nzw, fxw, fyw, fzw, ax, ay, az, cx, cy, cz, tMin, bestT, a[3], b[3]);
    const uint32_t ref[4] = {rf.x, rf.y, rf.z, rf.w};
    return enter_and_push(a, b, ref, stack, sp, anyRay);
}

#ifndef TMPT_QNODES
#define TMPT_QNODES 0
#endif

template <class Stack>
TMPT_HD uint32_t node_step(const SceneView& sc, uint32_t node, const RayCtx& r, float tMin, float bestT, Stack& stack, int& sp, bool anyRay) {
#if TMPT_QNODES
    return qnode_step(sc, node, r, tMin, bestT, stack, sp, anyRay);
#else
    return wide_node_step(sc, node, r, tMin, bestT, stack, sp, anyRay);
#endif
}

// One exact test of triangle slot `slot`; returns true when `best` improved.
TMPT_HD bool tri_step(const SceneView& sc, uint32_t slot, ex::V3 o, ex::V3 d, float tMin, float tMax, HitRec& best) {
    const uint32_t row0 = slot * 3u;  // 32-bit row index: slot < 2^27
    const float4 a = ld_row_tri(sc.tris + row0), b = ld_row_tri(sc.tris + (row0 + 1u)), c = ld_row_tri(sc.tris + (row0 + 2u));
    float t, u, v;
    if (mt_exact(o, d, ex::v3(a.x, a.y, a.z), ex::v3(b.x, b.y, b.z), ex::v3(c.x, c.y, c.z), tMin, tMax, t, u, v)) {
        const int id = (int)ex::f2u(a.w);
        if (t < best.t || (t == best.t && best.id >= 0 && id < best.id)) {
            best.t = t; best.id = id; best.u = u; best.v = v;
            return true;
        }
    }
    return false;
}

// Shadow query through the sun grid (sungrid.cuh): is the point p shadowed along the light direction l?  Walks the list of p's cell
// from the sun side down to p's own depth, exact test on every entry.  The next entry is requested before the current one is tested.
template <bool STATS>
TMPT_HD bool sun_occluded(const SceneView& sc, ex::V3 p, ex::V3 l, float tMin, float tMax, TravStats* stats) {
    const sun::View& g = sc.sun;
    const float u = sun::proj_u(g, p), v = sun::proj_v(g, p), w = sun::proj_w(g, p);
    if (!(u == u) || !(v == v) || !(w == w)) return false;  // NaN origin: a miss, as in the tree walk
    const int cx = sun::cell_of(u, g.loU, g.invCell, g.n), cy = sun::cell_of(v, g.loV, g.invCell, g.n);
    const uint32_t c = (uint32_t)cy * (uint32_t)g.n + (uint32_t)cx;
#ifdef __CUDA_ARCH__
    uint32_t e = __ldg(g.cellStart + c);
    const uint32_t end = __ldg(g.cellStart + c + 1);
#else
    uint32_t e = g.cellStart[c];
    const uint32_t end = g.cellStart[c + 1];
#endif
    if (e == end) return false;
#ifdef __CUDA_ARCH__
    uint2 cur = __ldg(g.entries + e);
#else
    uint2 cur = g.entries[e];
#endif
    for (;;) {
        if (ex::u2f(cur.y) < w) return false;  // this and everything after it ends before the origin
        ++e;
        uint2 nxt = cur;
        if (e < end) {
#ifdef __CUDA_ARCH__
            nxt = __ldg(g.entries + e);
#else
            nxt = g.entries[e];
#endif
        }
        if (STATS) ++stats->tris;
        const uint32_t row0 = cur.x * 3u;
        const float4 a = ld_row_tri(sc.tris + row0), b = ld_row_tri(sc.tris + (row0 + 1u)), cc = ld_row_tri(sc.tris + (row0 + 2u));
        float t, bu, bv;
        if (mt_exact(p, l, ex::v3(a.x, a.y, a.z), ex::v3(b.x, b.y, b.z), ex::v3(cc.x, cc.y, cc.z), tMin, tMax, t, bu, bv)) return true;
        if (e >= end) return false;
        cur = nxt;
    }
}

// Walk state of one ray in one lane.
struct WalkState {
    RayCtx r;
    ex::V3 o, d;
    HitRec best;
    uint32_t cur, triPos, triEnd;
    int sp;
    bool any;
};
TMPT_HD void walk_start(WalkState& w, const SceneView& sc, ex::V3 o, ex::V3 d, float tMax, bool any) {
    w.r = make_ray_ctx(o, d);
    w.o = o; w.d = d;
    w.best.id = -1; w.best.t = tMax; w.best.u = 0.0f; w.best.v = 0.0f;
    w.cur = ray_has_nan(o, d) ? NONE : sc.rootRef;
    w.triPos = 0; w.triEnd = 0;
    w.sp = 0;
    w.any = any;
}

// One iteration of the walk; returns true when the ray is finished.  A lane that reaches a
// leaf PARKS it (triPos..triEnd) and tests one of its triangles per iteration while it keeps
// walking inner nodes in the same iteration: the node step and the exact test of one ray
// overlap instead of alternating, and nobody loops over a whole leaf while its neighbours
// wait (+12 % on the frame, profiles/).  The walk runs at most one leaf ahead of the tests, so
// almost nothing is visited that a tighter best t would have culled.
// warp votes of the instrumented pass (the host emulation has no warps: every "warp" is one lane)
TMPT_HD bool stats_any(bool pred) {
#ifdef __CUDA_ARCH__
    return __any_sync(__activemask(), pred);
#else
    return pred;
#endif
}
TMPT_HD bool stats_leader() {
#ifdef __CUDA_ARCH__
    const unsigned m = __activemask();
    unsigned lane;
    asm("mov.u32 %0, %%laneid;" : "=r"(lane));
    return (unsigned)(__ffs((int)m) - 1) == lane;
#else
    return true;
#endif
}

template <bool STATS, class Stack>
TMPT_HD bool walk_step(WalkState& w, const SceneView& sc, float tMin, float tMax, Stack& stack, TravStats* stats) {
    if (STATS) {
        ++stats->iters;
        const bool nodeAny = stats_any(w.cur != NONE && !ref_is_leaf(w.cur));
        if (stats_leader()) { ++stats->warpIters; stats->nodeWarps += nodeAny ? 1 : 0; }
    }
    if (w.cur != NONE && !ref_is_leaf(w.cur)) {
        if (STATS) ++stats->nodes;
        w.cur = node_step(sc, w.cur, w.r, tMin, w.best.t, stack, w.sp, w.any);
        if (STATS && w.sp > stats->maxSp) stats->maxSp = w.sp;
    }
    if (w.cur != NONE && ref_is_leaf(w.cur) && w.triPos == w.triEnd) {  // park the leaf, free the walker
        w.triPos = leaf_first(w.cur);
        w.triEnd = w.triPos + (uint32_t)leaf_count(w.cur);
        w.cur = NONE;
    }
    if (STATS) {
        if (w.cur != NONE && ref_is_leaf(w.cur)) ++stats->leafWaits;
        const bool triAny = stats_any(w.triPos < w.triEnd);
        if (stats_leader()) stats->triWarps += triAny ? 1 : 0;
    }
    // if a pop is coming, request the top entry now: its latency hides behind the triangle test
    // (ncu: the compare after this load was the hottest stall site of the kernel)
    const bool popping = w.cur == NONE && w.sp > 0;
    unsigned long long top = 0;
    if (popping) top = stack.get(w.sp - 1);
    if (w.triPos < w.triEnd) {
        if (STATS) ++stats->tris;
        if (tri_step(sc, w.triPos++, w.o, w.d, tMin, tMax, w.best) && w.any) return true;
    }
    // pop ONE entry per iteration.  If the shrinking best.t has culled it the lane sits out the next node step and pops again
    // behind the next prefetch.  (A loop here that pops until something survives ran in 75 % of the iterations for a single
    // lane, with its stack latency exposed to the whole warp: +3.3 % without it.)
    if (popping) {
        --w.sp;
        // (the child slot in the key's two low bits is masked off: WITH it the key can exceed the entry distance by up to three
        // ulps, and an entry clamped to tMin = 0 -- key = slot, a denormal -- compared greater than a best t of +-0: a ray that starts
        // ON a vertex shared by several triangles lost the lower-index ones.  Found by tools/fuzz_emu.py.)
        if (ex::u2f((uint32_t)(top >> 32) & ~3u) <= w.best.t) w.cur = (uint32_t)top;
        else if (STATS) ++stats->culledPops;
    }
    const bool done = w.cur == NONE && w.triPos == w.triEnd && w.sp == 0;
    if (STATS && done) {
        const int lim[5] = {4, 8, 12, 16, 24};
        for (int k = 0; k < 5; ++k) stats->depthOver[k] += stats->maxSp > lim[k] ? 1 : 0;
        stats->maxSp = 0;
    }
    return done;
}

// Upstream's HitScene: every triangle, no tree.  Same candidate rule, so it must agree with
// traverse<false> bit for bit -- the on-GPU cross-check of box conservativeness at sizes the
// CPU checker cannot reach (TMPT_HIT_BRUTE) -- and the answer for rays that start too far from the scene for the padded
// boxes (ray_is_far).  Out of line on the device: it is never part of a hot loop.
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
HitRec scan_all(const float4* tris, int triCount, ex::V3 o, ex::V3 d, float tMin, float tMax) {
    HitRec best;
    best.id = -1; best.t = tMax; best.u = 0.0f; best.v = 0.0f;
    for (int k = 0; k < triCount; ++k) {
        const float4* tp = tris + (size_t)k * 3;
        const float4 a = TMPT_LDG4(tp + 0), b = TMPT_LDG4(tp + 1), c = TMPT_LDG4(tp + 2);
        float t, u, v;
        if (mt_exact(o, d, ex::v3(a.x, a.y, a.z), ex::v3(b.x, b.y, b.z), ex::v3(c.x, c.y, c.z), tMin, tMax, t, u, v)) {
            const int id = (int)ex::f2u(a.w);
            if (t < best.t || (t == best.t && best.id >= 0 && id < best.id)) { best.t = t; best.id = id; best.u = u; best.v = v; }
        }
    }
    return best;
}
TMPT_HD HitRec brute_force(const SceneView& sc, ex::V3 o, ex::V3 d, float tMin, float tMax) { return scan_all(sc.tris, sc.triCount, o, d, tMin, tMax); }

// The shadow query for a CALLER's origin (tmpt_hit_scene, TMPT_HIT_SUN): the grid's pads, like the boxes', are sized for origins
// near the scene -- the rounding of an origin's projection and the drift of the ray's own projection along its length both grow
// with |origin| (at 1e6 scene sizes a query lands in a neighbouring cell) -- so an origin beyond the far limit is answered by the
// scan, as in traverse().  The integrator's own shadow rays start on a triangle and call sun_occluded directly.
template <bool STATS>
TMPT_HD bool sun_query(const SceneView& sc, ex::V3 o, float tMin, float tMax, TravStats* stats) {
    const ex::V3 l = ex::v3(sc.sun.lx, sc.sun.ly, sc.sun.lz);
    if (ray_is_far(sc, o)) return scan_all(sc.tris, sc.triCount, o, l, tMin, tMax).id >= 0;
    return sun_occluded<STATS>(sc, o, l, tMin, tMax, stats);
}

template <bool ANY, bool STATS = false>
TMPT_HD HitRec traverse(const SceneView& sc, ex::V3 o, ex::V3 d, float tMin, float tMax, TravStats* stats = nullptr) {
    if (ray_is_far(sc, o)) return scan_all(sc.tris, sc.triCount, o, d, tMin, tMax);
    LocalStack stack;
    WalkState w;
    walk_start(w, sc, o, d, tMax, ANY);
    while (!walk_step<STATS>(w, sc, tMin, tMax, stack, stats)) {}
    return w.best;
}
// the second attempt of a ray whose short stack overflowed: out of line, so that it costs the walk above it nothing
template <bool ANY>
#ifdef __CUDA_ARCH__
__device__ __noinline__
#else
inline
#endif
HitRec traverse_again(const SceneView& sc, ex::V3 o, ex::V3 d, float tMin, float tMax) { return traverse<ANY, false>(sc, o, d, tMin, tMax); }

// FAR: check ray_is_far.  The batched HitScene kernels always do (the rays are the caller's); the render kernel only in the
// instantiation launch_render picks when the CAMERA stands beyond the limit -- every other ray of a path starts on a triangle.
template <bool ANY, bool STATS, bool FAR, class Stack>
TMPT_HD HitRec traverse_with(Stack& stack, const SceneView& sc, ex::V3 o, ex::V3 d, float tMin, float tMax, TravStats* stats = nullptr) {
    if (FAR && ray_is_far(sc, o)) return scan_all(sc.tris, sc.triCount, o, d, tMin, tMax);
    WalkState w;
    walk_start(w, sc, o, d, tMax, ANY);
    while (!walk_step<STATS>(w, sc, tMin, tMax, stack, stats)) {}
    if (Stack::kCanOverflow && stack.overflowed()) {
        stack.reset();
        if (STATS) ++stats->overflows;
        // (an any-hit answer "hit" stands whatever was dropped; everything else needs the entries that were lost)
        if (!(ANY && w.best.id >= 0)) return traverse_again<ANY>(sc, o, d, tMin, tMax);
    }
    return w.best;
}

}  // namespace bvh
