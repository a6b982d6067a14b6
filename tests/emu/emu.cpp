// tests/emu/emu.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// Compiles the product's __host__ __device__ headers (csrc/exact.cuh, bvh.cuh,
// build_logic.cuh, integrator.cuh) for the HOST with g++ and drives them serially, so that
// the CPU test suite (-m "not gpu", no GPU in the build container) can check the per-element
// logic of the BVH build, the traversal and the integrator against the oracle before any
// GPU time is spent.  The shipped library (libtmpt.so) never contains or calls this file;
// it has no CPU path.
#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <vector>

#include "../../toymeshpathtracer_b200/csrc/build_logic.cuh"
#include "../../toymeshpathtracer_b200/csrc/integrator.cuh"

namespace {
struct EmuScene {
    std::vector<float> tris9;
    std::vector<float4> nodes, tris;
    std::vector<uint4> qnodes;
    uint32_t status = 0;
    uint32_t counters[4] = {0, 0, 0, 0};
    float sah[2] = {0, 0};
    bvh::SceneView view;
    int n = 0;
    std::vector<uint32_t> sunStart;
    std::vector<uint2> sunEntries;
};

// kernels.cu: build_sun_grid, serially (count, prefix, fill, sort) with the same per-(triangle, cell) functions
void build_sun_grid(EmuScene* s, int cellsForced) {
    s->view.sun = sun::View{};
    if (s->n == 0 || cellsForced == 0) return;
    sun::View g;
    float bmin[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bmax[3] = {-3.0e38f, -3.0e38f, -3.0e38f};  // kernels.cu: k_prim_bounds
    for (size_t i = 0; i < (size_t)s->n * 9; ++i) {
        bmin[i % 3] = std::min(bmin[i % 3], s->tris9[i]);
        bmax[i % 3] = std::max(bmax[i % 3], s->tris9[i]);
    }
    if (!sun::setup_view(bmin, bmax, ex::normalize(ex::v3(-0.7f, 1.0f, 0.5f)), g)) return;
    sun::set_resolution(g, cellsForced > 0 ? cellsForced : sun::default_cells_per_side(s->n));
    const size_t nCells = (size_t)g.n * g.n;
    s->sunStart.assign(nCells + 1, 0u);
    for (int pass = 0; pass < 2; ++pass) {
        std::vector<uint32_t> cursor(nCells, 0u);
        for (int slot = 0; slot < s->n; ++slot) {
            const int id = (int)ex::f2u(s->tris[(size_t)slot * 3].w);
            const sun::Tri2 t = sun::project_tri(g, s->tris9.data() + (size_t)id * 9);
            int x0, x1, y0, y1;
            sun::cell_range(g, t, x0, x1, y0, y1);
            for (int cy = y0; cy <= y1; ++cy)
                for (int cx = x0; cx <= x1; ++cx) {
                    if (!sun::touches_cell(g, t, cx, cy)) continue;
                    const size_t c = (size_t)cy * g.n + cx;
                    if (pass == 0) ++s->sunStart[c + 1];
                    else s->sunEntries[s->sunStart[c] + cursor[c]++] = make_uint2((uint32_t)slot, ex::f2u(sun::far_depth(g, t, cx, cy)));
                }
        }
        if (pass == 0) {
            for (size_t c = 0; c < nCells; ++c) s->sunStart[c + 1] += s->sunStart[c];
            s->sunEntries.assign(std::max<size_t>(s->sunStart[nCells], 1), make_uint2(0, 0));
        }
    }
    for (size_t c = 0; c < nCells; ++c) {
        if (s->sunStart[c + 1] - s->sunStart[c] > sun::kSortMax) {  // k_sun_sort: long lists are not sorted and have no early exit
            for (uint32_t i = s->sunStart[c]; i < s->sunStart[c + 1]; ++i) s->sunEntries[i].y = ex::f2u(sun::kFarthest);
            continue;
        }
        std::sort(s->sunEntries.begin() + s->sunStart[c], s->sunEntries.begin() + s->sunStart[c + 1], sun::entry_before);
    }
    g.cellStart = s->sunStart.data();
    g.entries = s->sunEntries.data();
    s->view.sun = g;
}
}  // namespace

extern "C" {

// builder: 0 = binned SAH (top-down), 1 = LBVH (Karras)
void* emu_scene_create2(const float* tris9, int n, int builder, float cInner, float cTri, int maxLeaf) {
    EmuScene* s = new EmuScene();
    s->n = n;
    s->tris9.assign(tris9, tris9 + (size_t)n * 9);
    s->nodes.assign((size_t)std::max(n, 1) * bvh::NODE_F4, make_float4(0, 0, 0, 0));
    s->tris.assign((size_t)std::max(n, 1) * 3, make_float4(0, 0, 0, 0));
    s->view = bvh::SceneView{s->nodes.data(), s->tris.data(), s->tris9.data(), nullptr, n ? 0u : bvh::NONE, n, &s->status, nullptr, 0.0f};
    if (n == 0) return s;
    // k_prim_bounds
    bld::Box scene{3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = 0; i < n; ++i) scene = bld::box_union(scene, bld::tri_box(tris9 + (size_t)i * 9));
    const float maxAbs = std::max(std::max(std::max(std::fabs(scene.lox), std::fabs(scene.hix)), std::max(std::fabs(scene.loy), std::fabs(scene.hiy))),
                                  std::max(std::fabs(scene.loz), std::fabs(scene.hiz)));
    s->view.farLimit = 16.0f * maxAbs;  // kernels.cu: bvh_far_limit
    // padded primitive boxes (k_prim_boxes)
    std::vector<bld::Box> pbox(n);
    for (int i = 0; i < n; ++i) {
        bld::Box b = bld::tri_box(tris9 + (size_t)i * 9);
        const float dx = b.hix - b.lox, dy = b.hiy - b.loy, dz = b.hiz - b.loz;
        const float pad = bld::pad_for(sqrtf(dx * dx + dy * dy + dz * dz), maxAbs);
        b.lox -= pad; b.loy -= pad; b.loz -= pad; b.hix += pad; b.hiy += pad; b.hiz += pad;
        pbox[i] = b;
    }
    std::vector<uint32_t> prim(n);
    std::vector<float4> lo(2 * (size_t)n), hi(2 * (size_t)n);
    std::vector<int> left(2 * (size_t)n), right(2 * (size_t)n), parent(2 * (size_t)n, -1), first(2 * (size_t)n);
    std::vector<uint32_t> visits(n, 0);
    std::vector<uint64_t> skeys(n);
    bld::SahParams sp{cInner, cTri, maxLeaf};
    bool rootIsLeaf = n == 1;
    if (builder == 1) {
        // k_morton + k_radix_sort (stable) + k_leaf_boxes
        std::vector<uint64_t> keys(n);
        for (int i = 0; i < n; ++i) {
            const bld::Box b = bld::tri_box(tris9 + (size_t)i * 9);
            keys[i] = bld::morton63(0.5f * (b.lox + b.hix), 0.5f * (b.loy + b.hiy), 0.5f * (b.loz + b.hiz), scene);
            prim[i] = (uint32_t)i;
        }
        std::stable_sort(prim.begin(), prim.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
        for (int j = 0; j < n; ++j) {
            skeys[j] = keys[prim[j]];
            const bld::Box& b = pbox[prim[j]];
            lo[n - 1 + j] = make_float4(b.lox, b.loy, b.loz, cTri * bld::box_half_area(b));
            hi[n - 1 + j] = make_float4(b.hix, b.hiy, b.hiz, ex::u2f((uint32_t)-1));
            first[n - 1 + j] = j;
        }
    }
    bld::BinTree t{n, skeys.data(), left.data(), right.data(), parent.data(), lo.data(), hi.data(), first.data(), visits.data()};
    if (builder == 1) {
        if (n > 1) {
            for (int i = 0; i < n - 1; ++i) bld::karras_node(t, i);  // k_karras
            for (int j = 0; j < n; ++j) {                              // k_refit
                int node = parent[n - 1 + j];
                while (node >= 0) {
                    if (visits[node]++ == 0) break;
                    bld::refit_node(t, node, sp);
                    node = parent[node];
                }
            }
        }
    } else {
        // k_sah_level, serially: task = (node, first, count)
        struct Task { int node, first, count, depth; };
        for (int i = 0; i < n; ++i) prim[i] = (uint32_t)i;
        std::vector<Task> q{{0, 0, n, 0}};
        int nodeCounter = 1;
        std::vector<uint32_t> tmp(n);
        while (!q.empty()) {
            std::vector<Task> next;
            for (const Task& tk : q) {
                bld::Box nb = bld::empty_box(), cb = bld::empty_box();
                for (int i = 0; i < tk.count; ++i) {
                    const bld::Box& b = pbox[prim[tk.first + i]];
                    nb = bld::box_union(nb, b);
                    const float cx = 0.5f * (b.lox + b.hix), cy = 0.5f * (b.loy + b.hiy), cz = 0.5f * (b.loz + b.hiz);
                    cb = bld::box_union(cb, bld::Box{cx, cy, cz, cx, cy, cz});
                }
                const float cmin[3] = {cb.lox, cb.loy, cb.loz}, ext[3] = {cb.hix - cb.lox, cb.hiy - cb.loy, cb.hiz - cb.loz};
                bld::SahBin bins[3][bld::SAH_BINS];
                for (int a = 0; a < 3; ++a) for (int b = 0; b < bld::SAH_BINS; ++b) bins[a][b] = bld::SahBin{bld::empty_box(), 0};
                auto cen = [&](const bld::Box& b, int a) { return a == 0 ? 0.5f * (b.lox + b.hix) : a == 1 ? 0.5f * (b.loy + b.hiy) : 0.5f * (b.loz + b.hiz); };
                for (int i = 0; i < tk.count; ++i) {
                    const bld::Box& b = pbox[prim[tk.first + i]];
                    for (int a = 0; a < 3; ++a) {
                        bld::SahBin& bn = bins[a][bld::sah_bin_of(cen(b, a), cmin[a], ext[a])];
                        bn.box = bld::box_union(bn.box, b);
                        ++bn.count;
                    }
                }
                float costs[3 * (bld::SAH_BINS - 1)];
                int lcs[3 * (bld::SAH_BINS - 1)];
                for (int k = 0; k < 3 * (bld::SAH_BINS - 1); ++k) costs[k] = bld::sah_split_cost(bins[k / (bld::SAH_BINS - 1)], k % (bld::SAH_BINS - 1), &lcs[k]);
                bld::SahDecision d = tk.count == 1 ? bld::SahDecision{-1, 0, 0} : bld::sah_decide(costs, lcs, tk.count, bld::box_half_area(nb), sp, tk.depth);
                lo[tk.node] = make_float4(nb.lox, nb.loy, nb.loz, 0.0f);
                first[tk.node] = tk.first;
                if (d.axis < 0) {
                    hi[tk.node] = make_float4(nb.hix, nb.hiy, nb.hiz, ex::u2f((uint32_t)(-tk.count)));
                    continue;
                }
                hi[tk.node] = make_float4(nb.hix, nb.hiy, nb.hiz, ex::u2f((uint32_t)tk.count));
                int nl = 0, nr = 0;
                for (int i = 0; i < tk.count; ++i) {
                    const uint32_t id = prim[tk.first + i];
                    const bool goLeft = d.axis == 3 ? i < d.leftCount : bld::sah_bin_of(cen(pbox[id], d.axis), cmin[d.axis], ext[d.axis]) <= d.split;
                    if (goLeft) tmp[tk.first + nl++] = id;
                    else tmp[tk.first + tk.count - 1 - nr++] = id;
                }
                for (int i = 0; i < tk.count; ++i) prim[tk.first + i] = tmp[tk.first + i];
                const int base = nodeCounter; nodeCounter += 2;
                left[tk.node] = base; right[tk.node] = base + 1;
                next.push_back({base, tk.first, nl, tk.depth + 1});
                next.push_back({base + 1, tk.first + nl, nr, tk.depth + 1});
            }
            q.swap(next);
        }
    }
    rootIsLeaf = (int)ex::f2u(hi[0].w) < 0 || n == 1;
    if (builder == 1 && n == 1) { /* single leaf n-1+0 == node 0 */ }
    bld::WideOut w{s->nodes.data(), s->tris.data(), s->tris9.data(), prim.data(), s->counters, s->sah};
    s->counters[0] = 1;
    if (rootIsLeaf) {
        bld::emit_single_leaf_root(t, w, 0);
    } else {
        std::vector<bld::WorkItem> qa(n), qb(n);
        qa[0] = bld::WorkItem{0, 0u, 0};
        uint32_t count = 1;
        while (count) {
            uint32_t next = 0;
            for (uint32_t i = 0; i < count; ++i) bld::collapse_node(t, w, qa[i], qb.data(), &next);
            qa.swap(qb);
            count = next;
        }
    }
    // k_quantize_nodes
    s->qnodes.assign((size_t)s->counters[0] * bvh::QNODE_STRIDE, make_uint4(0, 0, 0, 0));
    for (uint32_t i = 0; i < s->counters[0]; ++i) bld::quantize_node(s->nodes.data(), s->qnodes.data(), i);
    s->view.qnodes = s->qnodes.data();
    build_sun_grid(s, -1);
    return s;
}
void* emu_scene_create(const float* tris9, int n) { return emu_scene_create2(tris9, n, 1, 1.0f, 1.0f, bvh::MAX_LEAF_TRIS); }
void emu_scene_destroy(void* h) { delete (EmuScene*)h; }
// [0] wide nodes [1] slots [2] leaves [3] max depth [4] status
void emu_scene_info(void* h, uint32_t out[5]) {
    EmuScene* s = (EmuScene*)h;
    for (int k = 0; k < 4; ++k) out[k] = s->counters[k];
    out[4] = s->status;
}
int emu_nodes(void* h, float* out) {
    EmuScene* s = (EmuScene*)h;
    if (out) memcpy(out, s->nodes.data(), (size_t)s->counters[0] * bvh::NODE_F4 * 16);
    return (int)s->counters[0];
}
// tmpt_scene_refit, serially: children have larger indices than their parents (the collapse allocates level by level), so a
// sweep from the last node to the root is bottom-up
void emu_scene_refit(void* h, const float* tris9) {
    EmuScene* s = (EmuScene*)h;
    if (s->n == 0) return;
    s->tris9.assign(tris9, tris9 + (size_t)s->n * 9);
    s->view.tris9 = s->tris9.data();
    bld::Box scene{3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = 0; i < s->n; ++i) scene = bld::box_union(scene, bld::tri_box(tris9 + (size_t)i * 9));
    const float maxAbs = std::max(std::max(std::max(std::fabs(scene.lox), std::fabs(scene.hix)), std::max(std::fabs(scene.loy), std::fabs(scene.hiy))),
                                  std::max(std::fabs(scene.loz), std::fabs(scene.hiz)));
    s->view.farLimit = 16.0f * maxAbs;
    const float4* nodes = s->nodes.data();
    for (uint32_t i = s->counters[0]; i-- > 0;)
        bld::refit_wide_node(s->nodes.data(), s->tris.data(), s->tris9.data(), i, maxAbs,
                             [nodes](uint32_t n, int row) { return nodes[(size_t)n * bvh::NODE_F4 + row]; });
    for (uint32_t i = 0; i < s->counters[0]; ++i) bld::quantize_node(s->nodes.data(), s->qnodes.data(), i);
    build_sun_grid(s, s->view.sun.n > 0 ? s->view.sun.n : -1);
}
// the quantised nodes, in LOGICAL row order (the bank skew undone): 16 uint32 per node
void emu_qnodes(void* h, uint32_t* out) {
    EmuScene* s = (EmuScene*)h;
    for (uint32_t i = 0; i < s->counters[0]; ++i)
        for (uint32_t r = 0; r < (uint32_t)bvh::QNODE_ROWS; ++r) memcpy(out + ((size_t)i * bvh::QNODE_ROWS + r) * 4, &s->qnodes[bvh::qnode_row(i, r)], 16);
}
void emu_slots(void* h, float* out) {
    EmuScene* s = (EmuScene*)h;
    memcpy(out, s->tris.data(), (size_t)s->n * 48);
}

void emu_hit_scene(void* h, const float* rays6, long n, float tMin, float tMax, int mode, int* outID, float* outT, float* outPos, float* outNormal) {
    EmuScene* s = (EmuScene*)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        const float* r = rays6 + i * 6;
        const ex::V3 o = ex::v3(r[0], r[1], r[2]), d = ex::v3(r[3], r[4], r[5]);
        bvh::HitRec hr = mode == 2 ? bvh::brute_force(s->view, o, d, tMin, tMax)
                         : mode == 1 ? bvh::traverse<true>(s->view, o, d, tMin, tMax)
                                     : bvh::traverse<false>(s->view, o, d, tMin, tMax);
        if (mode == 1) { outID[i] = hr.id < 0 ? -1 : 1; continue; }
        outID[i] = hr.id;
        if (hr.id >= 0) {
            ex::V3 pos, nrm;
            bvh::hit_payload(s->view, hr.id, hr.u, hr.v, pos, nrm);
            if (outT) outT[i] = hr.t;
            if (outPos) { outPos[i * 3] = pos.x; outPos[i * 3 + 1] = pos.y; outPos[i * 3 + 2] = pos.z; }
            if (outNormal) { outNormal[i * 3] = nrm.x; outNormal[i * 3 + 1] = nrm.y; outNormal[i * 3 + 2] = nrm.z; }
        }
    }
}

void emu_render(void* h, const float cam22[22], int w, int hgt, int spp, int row0, int row1, uint8_t* rgba, float* outLinear, long long* rayCount) {
    EmuScene* s = (EmuScene*)h;
    integ::Camera cam;
    memcpy(&cam, cam22, sizeof cam);
    const ex::V3 lightDir = ex::normalize(ex::v3(-0.7f, 1.0f, 0.5f));
    unsigned long long total = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : total)
    for (int y = row0; y < row1; ++y) {
        unsigned long long rays = 0;
        for (int x = 0; x < w; ++x) {
            ex::V3 lin;
            const uchar4 px = integ::render_pixel(s->view, cam, x, y, w, hgt, spp, lightDir, rays, &lin);
            const size_t p = (size_t)y * w + x;
            rgba[p * 4] = px.x; rgba[p * 4 + 1] = px.y; rgba[p * 4 + 2] = px.z; rgba[p * 4 + 3] = px.w;
            if (outLinear) { outLinear[p * 3] = lin.x; outLinear[p * 3 + 1] = lin.y; outLinear[p * 3 + 2] = lin.z; }
        }
        total += rays;
    }
    if (rayCount) *rayCount = (long long)total;
}

// work counters of a render: [0] rays [1] wide-node visits [2] triangle tests
void emu_render_stats(void* h, const float cam22[22], int w, int hgt, int spp, unsigned long long out[3]) {
    EmuScene* s = (EmuScene*)h;
    integ::Camera cam;
    memcpy(&cam, cam22, sizeof cam);
    const ex::V3 lightDir = ex::normalize(ex::v3(-0.7f, 1.0f, 0.5f));
    unsigned long long rays = 0, nodes = 0, tris = 0;
#pragma omp parallel for schedule(dynamic, 1) reduction(+ : rays, nodes, tris)
    for (int y = 0; y < hgt; ++y) {
        unsigned long long r = 0;
        bvh::TravStats ts;
        for (int x = 0; x < w; ++x) integ::render_pixel<true>(s->view, cam, x, y, w, hgt, spp, lightDir, r, nullptr, &ts);
        rays += r; nodes += ts.nodes; tris += ts.tris;
    }
    out[0] = rays; out[1] = nodes; out[2] = tris;
}

// walk iterations of each ray (development: distribution of ray lengths, tools/sim_pool.py)
void emu_ray_iters(void* h, const float* rays6, long n, float tMin, float tMax, int anyHit, int* outIters) {
    EmuScene* s = (EmuScene*)h;
#pragma omp parallel for schedule(dynamic, 256)
    for (long i = 0; i < n; ++i) {
        const float* r = rays6 + i * 6;
        bvh::LocalStack stack;
        bvh::WalkState w;
        bvh::walk_start(w, s->view, ex::v3(r[0], r[1], r[2]), ex::v3(r[3], r[4], r[5]), tMax, anyHit != 0);
        int it = 1;
        while (!bvh::walk_step<false>(w, s->view, tMin, tMax, stack, nullptr)) ++it;
        outIters[i] = it;
    }
}

// node visits per node index (development: how much of the walk happens in the top of the tree, tools/exp_top_levels.py)
void emu_node_visits(void* h, const float* rays6, long n, float tMin, float tMax, int anyHit, unsigned long long* visitsPerNode) {
    EmuScene* s = (EmuScene*)h;
    for (long i = 0; i < n; ++i) {
        const float* r = rays6 + i * 6;
        bvh::LocalStack stack;
        bvh::WalkState w;
        bvh::walk_start(w, s->view, ex::v3(r[0], r[1], r[2]), ex::v3(r[3], r[4], r[5]), tMax, anyHit != 0);
        for (;;) {
            if (w.cur != bvh::NONE && !bvh::ref_is_leaf(w.cur)) ++visitsPerNode[w.cur];
            if (bvh::walk_step<false>(w, s->view, tMin, tMax, stack, nullptr)) break;
        }
    }
}

// the sun grid: rebuild with n cells per side (0 = none: shadow rays walk the tree, -1 = default), counts, and the query itself
void emu_sun_grid(void* h, int cells) { build_sun_grid((EmuScene*)h, cells); }
void emu_sun_info(void* h, long long out[3]) {
    EmuScene* s = (EmuScene*)h;
    out[0] = s->view.sun.n;
    out[1] = s->view.sun.n ? (long long)s->sunStart[(size_t)s->view.sun.n * s->view.sun.n] : 0;
    long long longest = 0;
    if (s->view.sun.n) for (size_t c = 0; c < (size_t)s->view.sun.n * s->view.sun.n; ++c) longest = std::max<long long>(longest, s->sunStart[c + 1] - s->sunStart[c]);
    out[2] = longest;
}
void emu_sun_occluded(void* h, const float* origins3, long n, float tMin, float tMax, int* out, unsigned long long* triTests) {
    EmuScene* s = (EmuScene*)h;
    bvh::TravStats ts;
    for (long i = 0; i < n; ++i)  // as k_hit_scene<TMPT_HIT_SUN> answers it: far origins by the scan
        out[i] = s->view.sun.n > 0 && bvh::sun_query<true>(s->view, ex::v3(origins3[3 * i], origins3[3 * i + 1], origins3[3 * i + 2]), tMin, tMax, &ts) ? 1 : -1;
    if (triTests) *triTests = ts.tris;
}

void emu_sincos(float a, float* s, float* c) { ex::sincos_spec(a, *s, *c); }
int emu_sah_must_halve(int depth, int count) { return bld::sah_must_halve(depth, count) ? 1 : 0; }
int emu_max_tree_depth() { return bvh::MAX_TREE_DEPTH; }
uint32_t emu_pixel_seed(uint32_t i) { return ex::pixel_seed(i); }
uint32_t emu_chunk_seed(uint32_t chunk, uint32_t pixel, uint32_t pixels) { return ex::chunk_seed(chunk, pixel, pixels); }
}
