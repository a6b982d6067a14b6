// build_logic.cuh -- per-element logic of the on-device BVH build (K1).
//
// Replaces Scene::BuildOctree / OctreeNode::Subdivide / InternalDivide (scene.cpp:75-83,
// 99-160) and the SAT triangle-box test they rely on (maths.cpp:199-298): the octree is NOT
// rebuilt.  Pipeline (kernels.cu drives it, one launch per stage):
//   1 prim_bounds   triangle AABBs + scene bounds (atomic min/max on order-preserving ints)
//   2 morton_keys   63-bit Morton code of the box centre; padded triangle boxes
//   3 radix sort    (key, prim) pairs, 16 passes of 4 bits           [kernels.cu]
//   4 karras_node   binary radix tree over the sorted keys (Karras 2012), index tie-break
//   5 refit_up      bottom-up boxes, subtree counts, SAH cost and the leaf decision
//   6 collapse_node top-down: binary tree -> 4-wide nodes + leaf triangle slots
// Every function here is __host__ __device__ so tests/emu can run the same logic serially
// on the CPU; the library only runs it inside kernels.
#pragma once
#include "bvh.cuh"

namespace bld {

struct Box {
    float lox, loy, loz, hix, hiy, hiz;
};
TMPT_HD Box box_union(const Box& a, const Box& b) {
    return Box{fminf(a.lox, b.lox), fminf(a.loy, b.loy), fminf(a.loz, b.loz),
               fmaxf(a.hix, b.hix), fmaxf(a.hiy, b.hiy), fmaxf(a.hiz, b.hiz)};
}
TMPT_HD float box_half_area(const Box& b) {
    float dx = b.hix - b.lox, dy = b.hiy - b.loy, dz = b.hiz - b.loz;
    return dx * dy + dy * dz + dz * dx;
}

// Order-preserving float <-> uint map for atomicMin / atomicMax on floats.
TMPT_HD uint32_t float_to_ordered(float f) {
    uint32_t u = ex::f2u(f);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
TMPT_HD float ordered_to_float(uint32_t u) {
    return ex::u2f((u & 0x80000000u) ? (u & 0x7FFFFFFFu) : ~u);
}

// How far a triangle's box is grown before it enters the tree (DESIGN.md "Conservative
// boxes").  Two terms: (a) the float slab test and Moller-Trumbore's t each carry a few
// ulps of the largest coordinate in play -> 64 ulps of the scene's largest |coordinate|;
// (b) MT accepts points whose computed barycentrics pass although the exact ones are
// outside by a relative error that grows for grazing rays -> 1e-3 of the triangle's own
// box diagonal.  Growing boxes can only add work, never change a result.
TMPT_HD float pad_for(float triDiag, float sceneMaxAbs) {
    return 1.0e-3f * triDiag + 64.0f * 1.1920929e-7f * sceneMaxAbs;
}

TMPT_HD Box tri_box(const float* t9) {
    Box b;
    b.lox = fminf(fminf(t9[0], t9[3]), t9[6]); b.hix = fmaxf(fmaxf(t9[0], t9[3]), t9[6]);
    b.loy = fminf(fminf(t9[1], t9[4]), t9[7]); b.hiy = fmaxf(fmaxf(t9[1], t9[4]), t9[7]);
    b.loz = fminf(fminf(t9[2], t9[5]), t9[8]); b.hiz = fmaxf(fmaxf(t9[2], t9[5]), t9[8]);
    return b;
}

TMPT_HD uint64_t expand_bits21(uint64_t v) {  // 21 bits -> every third bit of 63
    v &= 0x1FFFFFull;
    v = (v | (v << 32)) & 0x1F00000000FFFFull;
    v = (v | (v << 16)) & 0x1F0000FF0000FFull;
    v = (v | (v << 8)) & 0x100F00F00F00F00Full;
    v = (v | (v << 4)) & 0x10C30C30C30C30C3ull;
    v = (v | (v << 2)) & 0x1249249249249249ull;
    return v;
}
TMPT_HD uint64_t morton63(float cx, float cy, float cz, const Box& scene) {
    float ex_ = scene.hix - scene.lox, ey = scene.hiy - scene.loy, ez = scene.hiz - scene.loz;
    float nx = ex_ > 0.0f ? (cx - scene.lox) / ex_ : 0.0f;
    float ny = ey > 0.0f ? (cy - scene.loy) / ey : 0.0f;
    float nz = ez > 0.0f ? (cz - scene.loz) / ez : 0.0f;
    const float S = 2097151.0f;  // 2^21 - 1
    uint64_t ix = (uint64_t)fminf(fmaxf(nx * S, 0.0f), S);
    uint64_t iy = (uint64_t)fminf(fmaxf(ny * S, 0.0f), S);
    uint64_t iz = (uint64_t)fminf(fmaxf(nz * S, 0.0f), S);
    return (expand_bits21(ix) << 2) | (expand_bits21(iy) << 1) | expand_bits21(iz);
}

TMPT_HD int clz64(uint64_t v) {
#ifdef __CUDA_ARCH__
    return __clzll((long long)v);
#else
    return v ? __builtin_clzll(v) : 64;
#endif
}
TMPT_HD int clz32(uint32_t v) {
#ifdef __CUDA_ARCH__
    return __clz((int)v);
#else
    return v ? __builtin_clz(v) : 32;
#endif
}

// Binary tree over n primitives in a pool of 2n-1 nodes, root = node 0.  Every node covers a
// contiguous range [first, first+count) of the ordered primitive list `prim` (position ->
// original index).  The LBVH builder (Karras) puts inner nodes at [0, n-1) and the single-
// primitive leaves at [n-1, 2n-1) (leaf n-1+j = sorted position j); the SAH builder allocates
// nodes as it goes.  Both hand the same structure to the wide collapse.
struct BinTree {
    int n;
    const uint64_t* keys;  // sorted Morton codes (LBVH only)
    int* left;             // [2n-1]
    int* right;            // [2n-1]
    int* parent;           // [2n-1] (LBVH only)
    float4* lo;            // [2n-1] xyz = box min, w = SAH cost of the subtree
    float4* hi;            // [2n-1] xyz = box max, w = bit-cast int: triangles in the subtree,
                           //         NEGATIVE when the node is (to become) one leaf
    int* first;            // [2n-1] first position of the node's range
    uint32_t* visits;      // [n-1] bottom-up arrival counters (LBVH only)
};

// Karras 2012, "Maximizing parallelism in the construction of BVHs, octrees and k-d trees":
// delta = length of the common prefix of two keys; equal keys fall back to the index bits.
TMPT_HD int karras_delta(const uint64_t* keys, int n, int i, int j) {
    if (j < 0 || j >= n) return -1;
    uint64_t a = keys[i], b = keys[j];
    if (a == b) return 64 + clz32((uint32_t)i ^ (uint32_t)j);
    return clz64(a ^ b);
}

TMPT_HD void karras_node(const BinTree& t, int i) {
    const int n = t.n;
    const int d = (karras_delta(t.keys, n, i, i + 1) - karras_delta(t.keys, n, i, i - 1)) >= 0 ? 1 : -1;
    const int dmin = karras_delta(t.keys, n, i, i - d);
    int lmax = 2;
    while (karras_delta(t.keys, n, i, i + lmax * d) > dmin) lmax *= 2;
    int l = 0;
    for (int s = lmax / 2; s >= 1; s /= 2)
        if (karras_delta(t.keys, n, i, i + (l + s) * d) > dmin) l += s;
    const int j = i + l * d;
    const int dnode = karras_delta(t.keys, n, i, j);
    int s = 0;
    for (int step = (l + 1) / 2;; step = (step + 1) / 2) {
        if (karras_delta(t.keys, n, i, i + (s + step) * d) > dnode) s += step;
        if (step == 1) break;
    }
    const int gamma = i + s * d + (d < 0 ? -1 : 0);
    const int lo = i < j ? i : j, hi = i < j ? j : i;
    const int L = (lo == gamma) ? (n - 1 + gamma) : gamma;
    const int R = (hi == gamma + 1) ? (n - 1 + gamma + 1) : (gamma + 1);
    t.left[i] = L;
    t.right[i] = R;
    t.first[i] = lo;
    t.parent[L] = i;
    t.parent[R] = i;
    if (i == 0) t.parent[0] = -1;
}

// SAH constants (relative): one 4-wide node visit vs one exact triangle test.  The binary
// tree charges C_INNER per binary node; a 4-wide node absorbs up to three of them.
struct SahParams {
    float cInner;  // per BINARY inner node
    float cTri;
    int maxLeaf;
};

// Combine two finished children into inner node `i` (called by whichever child arrives last).
TMPT_HD void refit_node(const BinTree& t, int i, const SahParams& sp) {
    const int L = t.left[i], R = t.right[i];
    const float4 llo = t.lo[L], lhi = t.hi[L], rlo = t.lo[R], rhi = t.hi[R];
    Box b = box_union(Box{llo.x, llo.y, llo.z, lhi.x, lhi.y, lhi.z}, Box{rlo.x, rlo.y, rlo.z, rhi.x, rhi.y, rhi.z});
    int cl = (int)ex::f2u(lhi.w), cr = (int)ex::f2u(rhi.w);
    int cnt = (cl < 0 ? -cl : cl) + (cr < 0 ? -cr : cr);
    float area = box_half_area(b);
    float costSplit = sp.cInner * area + llo.w + rlo.w;
    float costLeaf = sp.cTri * area * (float)cnt;
    bool asLeaf = cnt <= sp.maxLeaf && costLeaf <= costSplit;
    t.lo[i] = make_float4(b.lox, b.loy, b.loz, asLeaf ? costLeaf : costSplit);
    t.hi[i] = make_float4(b.hix, b.hiy, b.hiz, ex::u2f((uint32_t)(asLeaf ? -cnt : cnt)));
}

// ---- binned SAH (top-down builder): per-task logic shared by the kernel and the host emulation ----
#ifndef TMPT_SAH_BINS
#define TMPT_SAH_BINS 16
#endif
constexpr int SAH_BINS = TMPT_SAH_BINS;
struct SahBin {
    Box box;
    int count;
};
TMPT_HD Box empty_box() { return Box{3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f}; }

// bin of a centroid coordinate along an axis whose centroid range is [cmin, cmin + extent]
TMPT_HD int sah_bin_of(float c, float cmin, float extent) {
    if (!(extent > 0.0f)) return 0;
    int b = (int)((c - cmin) * ((float)SAH_BINS / extent));
    return b < 0 ? 0 : (b > SAH_BINS - 1 ? SAH_BINS - 1 : b);
}

// Cost (areaL*nL + areaR*nR) of splitting after bin `split` (left = bins [0, split]) on one axis;
// < 0 when one side would be empty.
TMPT_HD float sah_split_cost(const SahBin* bins, int split, int* outLeftCount) {
    Box l = empty_box(), r = empty_box();
    int nl = 0, nr = 0;
    for (int b = 0; b < SAH_BINS; ++b) {
        if (bins[b].count == 0) continue;
        if (b <= split) { l = box_union(l, bins[b].box); nl += bins[b].count; }
        else { r = box_union(r, bins[b].box); nr += bins[b].count; }
    }
    *outLeftCount = nl;
    if (nl == 0 || nr == 0) return -1.0f;
    return box_half_area(l) * (float)nl + box_half_area(r) * (float)nr;
}

// Depth guarantee.  The traversal stack holds bvh::STACK_SIZE entries, enough for bvh::MAX_TREE_DEPTH levels.  SAH splits can
// be arbitrarily lopsided (a mesh graded over many scales peels off one triangle per level), so the builder watches
// depth + ceil(log2(count)) of every task: once that reaches the limit the task is halved by position instead, which keeps the
// sum constant from there on -- no subtree can end deeper than the limit, whatever the input.  Healthy scenes never get
// near it (the Sponza stand-in: root 17, deepest task 25 of 40).
constexpr int SAH_DEPTH_LIMIT = bvh::MAX_TREE_DEPTH - 1;
TMPT_HD int ceil_log2(int n) { int k = 0; while ((1 << k) < n) ++k; return k; }
TMPT_HD bool sah_must_halve(int depth, int count) { return depth + ceil_log2(count) >= SAH_DEPTH_LIMIT; }

struct SahDecision {
    int axis;       // -1: make a leaf; 3: median split by position (binning found nothing)
    int split;      // last bin of the left side
    int leftCount;
};
// costs[axis * (SAH_BINS-1) + split] as computed by sah_split_cost; nodeArea = half area of the node box
TMPT_HD SahDecision sah_decide(const float* costs, const int* leftCounts, int count, float nodeArea, const SahParams& sp, int depth = 0) {
    if (sah_must_halve(depth, count)) return count <= sp.maxLeaf ? SahDecision{-1, 0, 0} : SahDecision{3, 0, count / 2};
    SahDecision d{-1, 0, 0};
    float best = 3.0e38f;
    for (int k = 0; k < 3 * (SAH_BINS - 1); ++k) {
        if (costs[k] >= 0.0f && costs[k] < best) { best = costs[k]; d.axis = k / (SAH_BINS - 1); d.split = k % (SAH_BINS - 1); d.leftCount = leftCounts[k]; }
    }
    const float leafCost = sp.cTri * (float)count * nodeArea;
    if (d.axis < 0) {  // all centroids coincide
        if (count <= sp.maxLeaf) return SahDecision{-1, 0, 0};
        return SahDecision{3, 0, count / 2};
    }
    const float splitCost = sp.cInner * nodeArea + sp.cTri * best;
    if (count <= sp.maxLeaf && leafCost <= splitCost) return SahDecision{-1, 0, 0};
    return d;
}

// ---- collapse: binary -> 4-wide ----
struct WideOut {
    float4* nodes;         // bvh::NODE_F4 float4 per wide node
    float4* tris;          // 3 float4 per slot
    const float* tris9;    // original triangles
    const uint32_t* prim;  // sorted position -> original index
    uint32_t* counters;    // [0] wide nodes allocated, [1] triangle slots allocated, [2] leaves, [3] max depth
    float* sahAccum;       // [0] sum over inner nodes of half-area, [1] sum over leaves of half-area * count
    uint32_t* parent;      // [wide node] its parent's index (root: itself); may be null.  Kept for tmpt_scene_refit.
};
struct WorkItem {
    int bnode;     // binary inner node to expand
    uint32_t wide; // wide node index it becomes
    int depth;
};

TMPT_HD bool subtree_is_leaf(const BinTree& t, int node) { return (int)ex::f2u(t.hi[node].w) < 0; }
TMPT_HD int subtree_count(const BinTree& t, int node) {
    int c = (int)ex::f2u(t.hi[node].w);
    return c < 0 ? -c : c;
}

// counter ops: atomic on the device, plain in the serial host emulation
TMPT_HD uint32_t counter_add(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    return atomicAdd(p, v);
#else
    uint32_t o = *p; *p += v; return o;
#endif
}
TMPT_HD void counter_max(uint32_t* p, uint32_t v) {
#ifdef __CUDA_ARCH__
    atomicMax(p, v);
#else
    if (v > *p) *p = v;
#endif
}
TMPT_HD void accum_add(float* p, float v) {
#ifdef __CUDA_ARCH__
    atomicAdd(p, v);
#else
    *p += v;
#endif
}

// Write the triangles of a leaf node (its range of the ordered list) into consecutive slots.
TMPT_HD void emit_leaf_tris(const BinTree& t, const WideOut& w, int node, uint32_t firstSlot) {
    const int first = t.first[node], cnt = subtree_count(t, node);
    for (int k = 0; k < cnt; ++k) {
        const uint32_t id = w.prim[first + k];
        const float* p = w.tris9 + (size_t)id * 9;
        ex::V3 v0 = ex::v3(p[0], p[1], p[2]), v1 = ex::v3(p[3], p[4], p[5]), v2 = ex::v3(p[6], p[7], p[8]);
        ex::V3 e1 = ex::sub(v1, v0), e2 = ex::sub(v2, v0);  // maths.cpp:343-344
        float4* o = w.tris + (size_t)(firstSlot + k) * 3;
        o[0] = make_float4(v0.x, v0.y, v0.z, ex::u2f(id));
        o[1] = make_float4(e1.x, e1.y, e1.z, 0.0f);
        o[2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
    }
}

// Expand binary inner node `it.bnode` into wide node `it.wide`: repeatedly open the child
// with the largest surface area until there are four (or only leaves remain).  Children that
// stay inner get consecutive wide indices and are appended to `outQueue`.
TMPT_HD void collapse_node(const BinTree& t, const WideOut& w, const WorkItem& it, WorkItem* outQueue, uint32_t* outCount) {
    int c[4];
    int k = 2;
    c[0] = t.left[it.bnode];
    c[1] = t.right[it.bnode];
    while (k < 4) {
        int pick = -1;
        float bestArea = -1.0f;
        for (int j = 0; j < k; ++j) {
            if (subtree_is_leaf(t, c[j])) continue;
            const float4 lo = t.lo[c[j]], hi = t.hi[c[j]];
            float a = box_half_area(Box{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z});
            if (a > bestArea) { bestArea = a; pick = j; }
        }
        if (pick < 0) break;
        const int open = c[pick];
        c[pick] = t.left[open];
        c[k++] = t.right[open];
    }
    int nInner = 0, nLeafTris = 0;
    for (int j = 0; j < k; ++j) {
        if (subtree_is_leaf(t, c[j])) nLeafTris += subtree_count(t, c[j]);
        else ++nInner;
    }
    uint32_t wideBase = nInner ? counter_add(&w.counters[0], (uint32_t)nInner) : 0u;
    uint32_t qBase = nInner ? counter_add(outCount, (uint32_t)nInner) : 0u;
    // the leaf children of one node get CONSECUTIVE triangle slots: a ray that visits one usually visits its
    // siblings, and a 128-byte line holds 2.7 slots
    uint32_t slotBase = nLeafTris ? counter_add(&w.counters[1], (uint32_t)nLeafTris) : 0u;

    float lox[4], loy[4], loz[4], hix[4], hiy[4], hiz[4];
    uint32_t refs[4];
    float myArea;
    {
        const float4 lo = t.lo[it.bnode], hi = t.hi[it.bnode];
        myArea = box_half_area(Box{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z});
    }
    accum_add(&w.sahAccum[0], myArea);
    int inner = 0;
    for (int j = 0; j < 4; ++j) {
        if (j >= k) {
            // empty child: an INVERTED box (lo = +3e38, hi = -3e38) fails tNear <= tFar for every ray, so the
            // traversal needs no "is this child there" test
            lox[j] = loy[j] = loz[j] = 3.0e38f;
            hix[j] = hiy[j] = hiz[j] = -3.0e38f;
            refs[j] = bvh::NONE;
            continue;
        }
        const float4 lo = t.lo[c[j]], hi = t.hi[c[j]];
        lox[j] = lo.x; loy[j] = lo.y; loz[j] = lo.z; hix[j] = hi.x; hiy[j] = hi.y; hiz[j] = hi.z;
        if (subtree_is_leaf(t, c[j])) {
            const int cnt = subtree_count(t, c[j]);
            const uint32_t first = slotBase;
            slotBase += (uint32_t)cnt;
            counter_add(&w.counters[2], 1u);
            emit_leaf_tris(t, w, c[j], first);
            refs[j] = bvh::make_leaf_ref(first, cnt);
            accum_add(&w.sahAccum[1], box_half_area(Box{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z}) * (float)cnt);
        } else {
            const uint32_t wi = wideBase + (uint32_t)inner;
            refs[j] = wi;
            if (w.parent) w.parent[wi] = it.wide;
            outQueue[qBase + inner] = WorkItem{c[j], wi, it.depth + 1};
            ++inner;
        }
    }
    counter_max(&w.counters[3], (uint32_t)it.depth);
    float4* o = w.nodes + (size_t)it.wide * bvh::NODE_F4;
    o[0] = make_float4(lox[0], lox[1], lox[2], lox[3]);
    o[1] = make_float4(hix[0], hix[1], hix[2], hix[3]);
    o[2] = make_float4(loy[0], loy[1], loy[2], loy[3]);
    o[3] = make_float4(hiy[0], hiy[1], hiy[2], hiy[3]);
    o[4] = make_float4(loz[0], loz[1], loz[2], loz[3]);
    o[5] = make_float4(hiz[0], hiz[1], hiz[2], hiz[3]);
    o[6] = make_float4(ex::u2f(refs[0]), ex::u2f(refs[1]), ex::u2f(refs[2]), ex::u2f(refs[3]));
}

// Root special case: the whole scene is one leaf (n <= maxLeaf and SAH says so, or n == 1).
// Emits wide node 0 with a single leaf child.
TMPT_HD void emit_single_leaf_root(const BinTree& t, const WideOut& w, int rootNode) {
    const int cnt = subtree_count(t, rootNode);
    const uint32_t first = counter_add(&w.counters[1], (uint32_t)cnt);
    counter_add(&w.counters[2], 1u);
    emit_leaf_tris(t, w, rootNode, first);
    const float4 lo = t.lo[rootNode], hi = t.hi[rootNode];
    const float F = 3.0e38f;  // empty children: inverted boxes (see collapse_node)
    float4* o = w.nodes;
    o[0] = make_float4(lo.x, F, F, F); o[1] = make_float4(hi.x, -F, -F, -F);
    o[2] = make_float4(lo.y, F, F, F); o[3] = make_float4(hi.y, -F, -F, -F);
    o[4] = make_float4(lo.z, F, F, F); o[5] = make_float4(hi.z, -F, -F, -F);
    o[6] = make_float4(ex::u2f(bvh::make_leaf_ref(first, cnt)), ex::u2f(bvh::NONE), ex::u2f(bvh::NONE), ex::u2f(bvh::NONE));
    accum_add(&w.sahAccum[1], box_half_area(Box{lo.x, lo.y, lo.z, hi.x, hi.y, hi.z}) * (float)cnt);
}


// ---- refit: same topology, moved vertices (tmpt_scene_refit) ----
// Recompute every child box of wide node `node` from the CURRENT triangle array: a leaf child's box is the union of
// its triangles' padded boxes (and its slots are rewritten with the new v0 / e1 / e2), an inner child's box is the
// union of that child's own four child boxes, which must already be up to date (the kernel walks bottom-up).
// `ldnode` reads a row of another node (on the device: a load that bypasses the non-coherent L1).
template <class LoadRow>
TMPT_HD void refit_wide_node(float4* nodes, float4* tris, const float* tris9, uint32_t node, float sceneMaxAbs, LoadRow ldnode) {
    float4* o = nodes + (size_t)node * bvh::NODE_F4;
    const float4 rf = o[6];
    const uint32_t refs[4] = {ex::f2u(rf.x), ex::f2u(rf.y), ex::f2u(rf.z), ex::f2u(rf.w)};
    float lo[3][4], hi[3][4];
    for (int k = 0; k < 4; ++k) {
        Box b{3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};  // empty child: stays inverted
        if (refs[k] != bvh::NONE && bvh::ref_is_leaf(refs[k])) {
            const uint32_t first = bvh::leaf_first(refs[k]);
            for (int j = 0; j < bvh::leaf_count(refs[k]); ++j) {
                float4* slot = tris + (size_t)(first + j) * 3;
                const uint32_t id = ex::f2u(slot[0].w);
                const float* p = tris9 + (size_t)id * 9;
                const ex::V3 v0 = ex::v3(p[0], p[1], p[2]), v1 = ex::v3(p[3], p[4], p[5]), v2 = ex::v3(p[6], p[7], p[8]);
                const ex::V3 e1 = ex::sub(v1, v0), e2 = ex::sub(v2, v0);  // maths.cpp:343-344
                slot[0] = make_float4(v0.x, v0.y, v0.z, ex::u2f(id));
                slot[1] = make_float4(e1.x, e1.y, e1.z, 0.0f);
                slot[2] = make_float4(e2.x, e2.y, e2.z, 0.0f);
                Box tb = tri_box(p);
                const float dx = tb.hix - tb.lox, dy = tb.hiy - tb.loy, dz = tb.hiz - tb.loz;
                const float pad = pad_for(sqrtf(dx * dx + dy * dy + dz * dz), sceneMaxAbs);
                tb.lox -= pad; tb.loy -= pad; tb.loz -= pad; tb.hix += pad; tb.hiy += pad; tb.hiz += pad;
                b = box_union(b, tb);
            }
        } else if (refs[k] != bvh::NONE) {
            const float4 r0 = ldnode(refs[k], 0), r1 = ldnode(refs[k], 1), r2 = ldnode(refs[k], 2), r3 = ldnode(refs[k], 3), r4 = ldnode(refs[k], 4),
                         r5 = ldnode(refs[k], 5);
            // (a child's empty slots are inverted boxes: min / max ignore them)
            b.lox = fminf(fminf(r0.x, r0.y), fminf(r0.z, r0.w)); b.hix = fmaxf(fmaxf(r1.x, r1.y), fmaxf(r1.z, r1.w));
            b.loy = fminf(fminf(r2.x, r2.y), fminf(r2.z, r2.w)); b.hiy = fmaxf(fmaxf(r3.x, r3.y), fmaxf(r3.z, r3.w));
            b.loz = fminf(fminf(r4.x, r4.y), fminf(r4.z, r4.w)); b.hiz = fmaxf(fmaxf(r5.x, r5.y), fmaxf(r5.z, r5.w));
        }
        lo[0][k] = b.lox; lo[1][k] = b.loy; lo[2][k] = b.loz; hi[0][k] = b.hix; hi[1][k] = b.hiy; hi[2][k] = b.hiz;
    }
    for (int a = 0; a < 3; ++a) {
        o[2 * a] = make_float4(lo[a][0], lo[a][1], lo[a][2], lo[a][3]);
        o[2 * a + 1] = make_float4(hi[a][0], hi[a][1], hi[a][2], hi[a][3]);
    }
}
TMPT_HD int wide_inner_children(const float4* nodes, uint32_t node) {
    const float4 rf = nodes[(size_t)node * bvh::NODE_F4 + 6];
    const uint32_t refs[4] = {ex::f2u(rf.x), ex::f2u(rf.y), ex::f2u(rf.z), ex::f2u(rf.w)};
    int c = 0;
    for (int k = 0; k < 4; ++k) c += (refs[k] != bvh::NONE && !bvh::ref_is_leaf(refs[k])) ? 1 : 0;
    return c;
}

// ---- quantise: float wide node -> 64-byte node (bvh.cuh: qnode_step) ----
// Per axis: grid step 2^e, the smallest power of two that fits the children's span into the 8-bit range with the margins
// below; origin = a float at least 1/64 step under the lowest child plane; lo bytes floor(x - 1/64), hi bytes ceil(x + 1/64)
// in grid units.  The differences are formed in binary64, where they are exact (or wrong by far less than the margin).
// An empty child keeps its inverted box (lo byte 255, hi byte 0) and refers to triangle slot 0: should rounding in the slab
// test ever let a ray into it, it tests a real triangle it is allowed to test anyway.
TMPT_HD float float_below(float x) {  // the next float towards -infinity (finite, non-NaN input)
    uint32_t u = ex::f2u(x);
    if ((u & 0x7FFFFFFFu) == 0u) return ex::u2f(0x80000001u);
    return ex::u2f((u & 0x80000000u) ? u + 1u : u - 1u);
}
TMPT_HD double pow2_d(int e) {  // 2^e as a double, |e| < 1000
    unsigned long long b = (unsigned long long)(e + 1023) << 52;
    double d;
#ifdef __CUDA_ARCH__
    d = __longlong_as_double((long long)b);
#else
    memcpy(&d, &b, 8);
#endif
    return d;
}
TMPT_HD void quantize_node(const float4* nodesF, uint4* qnodes, uint32_t i) {
    const float4* n = nodesF + (size_t)i * bvh::NODE_F4;
    uint32_t refs[4];
    {
        const float4 rf = n[6];
        refs[0] = ex::f2u(rf.x); refs[1] = ex::f2u(rf.y); refs[2] = ex::f2u(rf.z); refs[3] = ex::f2u(rf.w);
    }
    uint32_t originBits[3], expo[3], loW[3], hiW[3];
    for (int a = 0; a < 3; ++a) {
        const float4 lo4 = n[2 * a], hi4 = n[2 * a + 1];
        const float lo[4] = {lo4.x, lo4.y, lo4.z, lo4.w}, hi[4] = {hi4.x, hi4.y, hi4.z, hi4.w};
        float mn = 3.0e38f, mx = -3.0e38f;
        for (int k = 0; k < 4; ++k)
            if (refs[k] != bvh::NONE) { mn = fminf(mn, lo[k]); mx = fmaxf(mx, hi[k]); }
        const double ext = (double)mx - (double)mn;
        // first guess for e: 2^e * 240 >= ext, and not below a quarter ulp of the larger coordinate (the origin is a float)
        int e = -100;
        {
            const float m = fmaxf(fabsf(mn), fabsf(mx));
            const int ulpExp = (int)((ex::f2u(m) >> 23) & 0xFFu) - 127 - 23 - 2;
            if (m > 0.0f && ulpExp > e) e = ulpExp;
            while (e < 100 && pow2_d(e) * 240.0 < ext) ++e;
        }
        for (;; ++e) {
            const double step = pow2_d(e);
            float origin = (float)((double)mn - 0.5 * step);
            while ((double)origin > (double)mn - step * (1.0 / 64.0)) origin = float_below(origin);
            bool fits = true;
            uint32_t lw = 0, hw = 0;
            for (int k = 0; k < 4; ++k) {
                uint32_t ql = 255u, qh = 0u;
                if (refs[k] != bvh::NONE) {
                    const double fl = floor(((double)lo[k] - (double)origin) / step - 1.0 / 64.0);
                    const double ch = ceil(((double)hi[k] - (double)origin) / step + 1.0 / 64.0);
                    if (fl < 0.0 || ch > 255.0) { fits = false; break; }
                    ql = (uint32_t)fl; qh = (uint32_t)ch;
                }
                lw |= ql << (8 * k); hw |= qh << (8 * k);
            }
            if (!fits && e < 120) continue;
            originBits[a] = ex::f2u(origin); expo[a] = (uint32_t)(e + 15 + 127); loW[a] = lw; hiW[a] = hw;
            break;
        }
    }
    for (int k = 0; k < 4; ++k)
        if (refs[k] == bvh::NONE) refs[k] = bvh::make_leaf_ref(0u, 1);
    uint4* o = qnodes;
    o[bvh::qnode_row(i, 0)] = make_uint4(originBits[0], originBits[1], originBits[2], expo[0] | (expo[1] << 8) | (expo[2] << 16));
    o[bvh::qnode_row(i, 1)] = make_uint4(refs[0], refs[1], refs[2], refs[3]);
    o[bvh::qnode_row(i, 2)] = make_uint4(loW[0], hiW[0], loW[1], hiW[1]);
    o[bvh::qnode_row(i, 3)] = make_uint4(loW[2], hiW[2], 0x3F800000u, 0u);
}

}  // namespace bld
