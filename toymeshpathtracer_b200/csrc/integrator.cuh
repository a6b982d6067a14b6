// integrator.cuh -- the per-sample arithmetic of the Trace() loop: camera rays, sky,
// direct light, diffuse scatter, the back-to-front unwind and the pixel resolve.
//
// Replaces Camera::GetRay (maths.h:93-104), Scatter (main.cpp:44-73), Trace (main.cpp:82-119)
// and the pixel epilogue of TraceImageBody::operator() (main.cpp:209-233).  Every value is
// formed with the reference's operation order and one rounding per operation (exact.cuh), so
// a pixel computed here equals, bit for bit, the same pixel computed by the CPU checker with
// the same RNG stream.  Pure per-lane functions: the kernels in kernels.cu decide how lanes
// are scheduled; __host__ __device__ so tests/emu can run them without a GPU.
#pragma once
#include "bvh.cuh"

namespace integ {

constexpr int kMaxDepth = 10;        // main.cpp:33
constexpr float kMinT = 0.001f;      // main.cpp:30
constexpr float kMaxT = 1.0e7f;      // main.cpp:31

// Camera (maths.h:106-111), 22 floats, same order as tmpt_camera
struct Camera {
    ex::V3 origin, lowerLeftCorner, horizontal, vertical, u, v, w;
    float lensRadius;
};

// maths.h:93-104
TMPT_HD void camera_get_ray(const Camera& c, float s, float t, uint32_t& rng, ex::V3& o, ex::V3& d) {
    float px, py;
    ex::random_in_unit_disk(rng, px, py);
    const float rdx = ex::mul(c.lensRadius, px), rdy = ex::mul(c.lensRadius, py);  // lensRadius * disk
    const ex::V3 offset = ex::add(ex::muls(c.u, rdx), ex::muls(c.v, rdy));
    o = ex::add(c.origin, offset);
    d = ex::normalize(ex::sub(ex::sub(ex::add(ex::add(c.lowerLeftCorner, ex::muls(c.horizontal, s)), ex::muls(c.vertical, t)), c.origin), offset));
}

// main.cpp:214-215 as the reference build evaluates it (v drawn before u), then GetRay
TMPT_HD void primary_ray(const Camera& c, int x, int y, float invW, float invH, uint32_t& rng, ex::V3& o, ex::V3& d) {
    const float fv = ex::mul(ex::add((float)y, ex::random_float01(rng)), invH);
    const float fu = ex::mul(ex::add((float)x, ex::random_float01(rng)), invW);
    camera_get_ray(c, fu, fv, rng, o, d);
}

// main.cpp:106-107: ((1-t)*(1,1,1) + t*(0.5,0.7,1.0)) * 0.5
TMPT_HD ex::V3 sky(ex::V3 dir) {
    const float t = ex::mul(0.5f, ex::add(dir.y, 1.0f));
    const float a = ex::sub(1.0f, t);
    return ex::muls(ex::add(ex::v3(ex::mul(1.0f, a), ex::mul(1.0f, a), ex::mul(1.0f, a)), ex::muls(ex::v3(0.5f, 0.7f, 1.0f), t)), 0.5f);
}

// main.cpp:62-68: the scalar that multiplies albedo*kLightColor when the sun is visible
TMPT_HD float sun_term(ex::V3 normal, ex::V3 rayDir, ex::V3 lightDir) {
    const ex::V3 nl = ex::dot(normal, rayDir) < 0.0f ? normal : ex::neg(normal);
    return fmaxf(0.0f, ex::dot(lightDir, nl));
}

// main.cpp:71-72: normalize((pos + normal + RandomUnitVector) - pos), GEOMETRIC normal
TMPT_HD ex::V3 scatter_dir(ex::V3 pos, ex::V3 normal, uint32_t& rng) {
    const ex::V3 target = ex::add(ex::add(pos, normal), ex::random_unit_vector(rng));
    return ex::normalize(ex::sub(target, pos));
}

// main.cpp:112-116 with light[i] = (albedo*kLightColor)*k_i (+ 0, main.cpp:48, 68) and atten = albedo
TMPT_HD ex::V3 unwind_step(float k, ex::V3 color) {
    const ex::V3 albedo = ex::v3(0.7f, 0.7f, 0.7f);
    const ex::V3 kLightColor = ex::v3(0.7f, 0.6f, 0.5f);
    const ex::V3 light = ex::add(ex::v3(0.0f, 0.0f, 0.0f), ex::muls(ex::mulv(albedo, kLightColor), k));
    return ex::add(light, ex::mulv(albedo, color));
}
// (a shadowed bounce has light = (0,0,0) exactly, which is what k = 0 produces: finite * 0 = +0)

// main.cpp:221-233: mean, sqrt, quantise
TMPT_HD uchar4 resolve_pixel(ex::V3 sum, float sppRecip) {
    const ex::V3 c = ex::muls(sum, sppRecip);
    uchar4 px;
    px.x = ex::quantise(ex::sqrt_rn(c.x));
    px.y = ex::quantise(ex::sqrt_rn(c.y));
    px.z = ex::quantise(ex::sqrt_rn(c.z));
    px.w = 255;
    return px;
}

// One whole pixel, serially: the straightforward form used by the host emulation and by the
// first (non-wavefront) render kernel.  kk[] holds the per-bounce sun term (0 for a
// shadowed bounce).
template <bool STATS = false, bool FAR = true, class Stack, class Scene>
TMPT_HD ex::V3 trace_path(Stack& stack, const Scene& sc, const Camera& cam, ex::V3 o, ex::V3 d, ex::V3 lightDir, uint32_t& rng,
                          unsigned long long& rays, bvh::TravStats* stats = nullptr) {
    float kk[kMaxDepth];
    int depth = 0;
    ex::V3 color = ex::v3(0.0f, 0.0f, 0.0f);
    while (depth < kMaxDepth) {
        ++rays;
        const bvh::HitRec h = bvh::traverse_with<false, STATS, FAR>(stack, sc, o, d, kMinT, kMaxT, stats);
        if (h.id < 0) { color = sky(d); break; }
        ex::V3 pos, normal;
        bvh::hit_payload(sc, h.id, h.u, h.v, pos, normal);
        ++rays;
        // the shadow ray (main.cpp:59): through the sun grid when the scene has one (sungrid.cuh), else through the tree
#if defined(TMPT_SUN_ONLY) && TMPT_SUN_ONLY  // (experiment: what the tree fallback in the same kernel costs)
        const bool shadowed = bvh::sun_occluded<STATS>(sc, pos, lightDir, kMinT, kMaxT, stats);
#else
        const bool shadowed = sc.sun.n > 0 ? bvh::sun_occluded<STATS>(sc, pos, lightDir, kMinT, kMaxT, stats)
                                           : bvh::traverse_with<true, STATS, FAR>(stack, sc, pos, lightDir, kMinT, kMaxT, stats).id >= 0;
#endif
        kk[depth] = shadowed ? 0.0f : sun_term(normal, d, lightDir);
        d = scatter_dir(pos, normal, rng);
        o = pos;
        ++depth;
    }
    for (int i = depth - 1; i >= 0; --i) color = unwind_step(kk[i], color);
    return color;
}

// The unit of parallel work: one CHUNK of a pixel's samples.  A chunk owns an XorShift32 stream
// seeded from (chunk, pixel) and adds its samples in order; a pixel is the in-order sum of its
// chunk sums (DESIGN.md "RNG").  The chunk length depends on spp only -- spp/32 clamped to
// [1, 8] -- so a frame does not depend on how it is split over lanes or GPUs.  Short chunks keep
// the work items short: 2 samples at the headline 64 spp (with 8-sample chunks an item ran for
// 4.4 ms, and on an 8-GPU split, 60 ms per frame, every GPU lost half an item at the end of its
// share: 3.4 % of the frame), one sample below 64 spp (a 640x360x4 frame with four-sample items
// has 1.5 of them per resident warp), 8 samples from 256 spp on, where chunk sums per pixel
// would otherwise cost more memory than they save time.
constexpr int kMaxChunkSamples = 8;
TMPT_HD int chunk_len(int spp) {
    const int c = spp / 32;
    return c < 1 ? 1 : c > kMaxChunkSamples ? kMaxChunkSamples : c;
}
TMPT_HD int chunk_count(int spp) { const int c = chunk_len(spp); return (spp + c - 1) / c; }

// `len` = samples per chunk: chunk_len(spp) for a one-shot frame, kMaxChunkSamples for a progressive pass
template <bool STATS = false, bool FAR = true, class Stack, class Scene>
TMPT_HD ex::V3 render_chunk(Stack& stack, const Scene& sc, const Camera& cam, int x, int y, int chunk, int width, int height, int spp, int len, ex::V3 lightDir,
                            unsigned long long& rays, bvh::TravStats* stats = nullptr) {
    const float invW = ex::divf(1.0f, (float)width), invH = ex::divf(1.0f, (float)height);
    uint32_t rng = ex::chunk_seed((uint32_t)chunk, (uint32_t)y * (uint32_t)width + (uint32_t)x, (uint32_t)width * (uint32_t)height);
    ex::V3 sum = ex::v3(0.0f, 0.0f, 0.0f);
    const int s0 = chunk * len, s1 = s0 + len < spp ? s0 + len : spp;
    for (int s = s0; s < s1; ++s) {
        ex::V3 o, d;
        primary_ray(cam, x, y, invW, invH, rng, o, d);
        sum = ex::add(sum, trace_path<STATS, FAR>(stack, sc, cam, o, d, lightDir, rng, rays, stats));
    }
    return sum;
}

// One whole pixel, serially (host emulation; the kernels spread the chunks over lanes).
template <bool STATS = false, class Scene>
TMPT_HD uchar4 render_pixel(const Scene& sc, const Camera& cam, int x, int y, int width, int height, int spp, ex::V3 lightDir,
                            unsigned long long& rays, ex::V3* outLinear = nullptr, bvh::TravStats* stats = nullptr) {
    const float sppRecip = ex::divf(1.0f, (float)spp);
    ex::V3 sum = ex::v3(0.0f, 0.0f, 0.0f);
    bvh::LocalStack stack;
    for (int c = 0; c < chunk_count(spp); ++c) sum = ex::add(sum, render_chunk<STATS>(stack, sc, cam, x, y, c, width, height, spp, chunk_len(spp), lightDir, rays, stats));
    if (outLinear) *outLinear = ex::muls(sum, sppRecip);
    return resolve_pixel(sum, sppRecip);
}

}  // namespace integ
