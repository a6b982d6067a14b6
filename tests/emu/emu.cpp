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
    uint32_t status = 0;
    uint32_t counters[4] = {0, 0, 0, 0};
    float sah[2] = {0, 0};
    bvh::SceneView view;
    int n = 0;
};
}  // namespace

extern "C" {

void* emu_scene_create(const float* tris9, int n) {
    EmuScene* s = new EmuScene();
    s->n = n;
    s->tris9.assign(tris9, tris9 + (size_t)n * 9);
    s->nodes.assign((size_t)std::max(n, 1) * bvh::NODE_F4, make_float4(0, 0, 0, 0));
    s->tris.assign((size_t)std::max(n, 1) * 3, make_float4(0, 0, 0, 0));
    s->view = bvh::SceneView{s->nodes.data(), s->tris.data(), s->tris9.data(), n ? 0u : bvh::NONE, n, &s->status};
    if (n == 0) return s;
    // k_prim_bounds
    bld::Box scene{3.0e38f, 3.0e38f, 3.0e38f, -3.0e38f, -3.0e38f, -3.0e38f};
    for (int i = 0; i < n; ++i) scene = bld::box_union(scene, bld::tri_box(tris9 + (size_t)i * 9));
    // k_morton
    std::vector<uint64_t> keys(n);
    std::vector<uint32_t> prim(n);
    for (int i = 0; i < n; ++i) {
        const bld::Box b = bld::tri_box(tris9 + (size_t)i * 9);
        keys[i] = bld::morton63(0.5f * (b.lox + b.hix), 0.5f * (b.loy + b.hiy), 0.5f * (b.loz + b.hiz), scene);
        prim[i] = (uint32_t)i;
    }
    // k_radix_sort (stable)
    std::stable_sort(prim.begin(), prim.end(), [&](uint32_t a, uint32_t b) { return keys[a] < keys[b]; });
    std::vector<uint64_t> skeys(n);
    for (int j = 0; j < n; ++j) skeys[j] = keys[prim[j]];
    // k_leaf_boxes
    std::vector<float4> lo(2 * (size_t)n), hi(2 * (size_t)n);
    const float maxAbs = std::max(std::max(std::max(std::fabs(scene.lox), std::fabs(scene.hix)), std::max(std::fabs(scene.loy), std::fabs(scene.hiy))),
                                  std::max(std::fabs(scene.loz), std::fabs(scene.hiz)));
    const float cTri = 1.0f, cInner = 1.0f;
    for (int j = 0; j < n; ++j) {
        bld::Box b = bld::tri_box(tris9 + (size_t)prim[j] * 9);
        const float dx = b.hix - b.lox, dy = b.hiy - b.loy, dz = b.hiz - b.loz;
        const float pad = bld::pad_for(sqrtf(dx * dx + dy * dy + dz * dz), maxAbs);
        b.lox -= pad; b.loy -= pad; b.loz -= pad; b.hix += pad; b.hiy += pad; b.hiz += pad;
        lo[n - 1 + j] = make_float4(b.lox, b.loy, b.loz, cTri * bld::box_half_area(b));
        hi[n - 1 + j] = make_float4(b.hix, b.hiy, b.hiz, ex::u2f(1u));
    }
    std::vector<int> left(n), right(n), parent(2 * (size_t)n, -1);
    std::vector<uint32_t> visits(n, 0);
    bld::BinTree t{n, skeys.data(), left.data(), right.data(), parent.data(), lo.data(), hi.data(), visits.data()};
    bld::SahParams sp{cInner, cTri, bvh::MAX_LEAF_TRIS};
    bld::WideOut w{s->nodes.data(), s->tris.data(), s->tris9.data(), prim.data(), s->counters, s->sah};
    bool rootIsLeaf = n == 1;
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
        rootIsLeaf = (int)ex::f2u(hi[0].w) < 0;
    }
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
    return s;
}
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

void emu_sincos(float a, float* s, float* c) { ex::sincos_spec(a, *s, *c); }
uint32_t emu_pixel_seed(uint32_t i) { return ex::pixel_seed(i); }
}
