// oracle/ref_harness.cpp -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.
//
// A C-ABI window onto the UNMODIFIED reference sources, compiled from where they
// lie under $(REF)/source (see oracle/Makefile; nothing is copied into this repo).
// It exists so that tests/ and oracle/gen_golden.py can
//   * pin the CPU restatement in oracle/oracle.c against the real reference, and
//   * generate the golden vectors under tests/golden/.
// Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl
// reference legs may load the resulting oracle/_ref/libref.so.
//
// How the reference's file-static functions are reached: source/main.cpp is
// #included below with `main` renamed, which brings Scatter (main.cpp:44),
// Trace (:82), LoadScene (:122) and TraceImageBody (:180) into this TU.
// Scene::m_triangles (scene.h:41) is read through a `private -> public` define
// that is active only while scene.h is parsed.
#include <algorithm>
#include <atomic>
#include <cfloat>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <utility>
#include <vector>

#include "tbb/blocked_range.h"
#include "tbb/cache_aligned_allocator.h"
#include "tbb/parallel_for.h"
#include "tbb/task_scheduler_init.h"

#include "maths.h"
#define private public
#include "scene.h"
#undef private

#define main reference_main
#include "main.cpp"
#undef main

namespace {
struct RefScene {
    std::unique_ptr<Scene> scene;
    glm::vec3 mn, mx;
};

inline Ray make_ray(const float* r6) {
    Ray r;  // default ctor: no AssertUnit, same as passing a prebuilt Ray
    r.orig = glm::vec3(r6[0], r6[1], r6[2]);
    r.dir = glm::vec3(r6[3], r6[4], r6[5]);
    return r;
}

inline void store_hit(const Hit& h, long i, float* outT, float* outPos, float* outNormal) {
    if (outT) outT[i] = h.t;
    if (outPos) { outPos[i * 3 + 0] = h.pos.x; outPos[i * 3 + 1] = h.pos.y; outPos[i * 3 + 2] = h.pos.z; }
    if (outNormal) { outNormal[i * 3 + 0] = h.normal.x; outNormal[i * 3 + 1] = h.normal.y; outNormal[i * 3 + 2] = h.normal.z; }
}

inline Camera camera_from22(const float* c22) {
    Camera cam;
    static_assert(sizeof(Camera) == 22 * sizeof(float), "Camera is 7 vec3 + float (maths.h:106-111)");
    std::memcpy(&cam, c22, sizeof(Camera));
    return cam;
}
}  // namespace

extern "C" {

// LoadScene (main.cpp:122-170) followed by BuildOctree exactly as main() does it
// (main.cpp:296-297, 312).  Prints the reference's own "Initialized scene" line.
void* ref_scene_load(const char* objPath, float mn[3], float mx[3], int* triCount) {
    auto* rs = new RefScene();
    rs->scene = LoadScene(objPath, rs->mn, rs->mx);
    if (!rs->scene) { delete rs; return nullptr; }
    glm::vec3 sceneSize = rs->mx - rs->mn;
    glm::vec3 extra = sceneSize * 0.7f;
    rs->scene->BuildOctree(rs->mn - extra, rs->mx + extra);
    for (int k = 0; k < 3; ++k) { mn[k] = rs->mn[k]; mx[k] = rs->mx[k]; }
    *triCount = int(rs->scene->m_triangles.size());
    return rs;
}

// Scene::Scene (scene.cpp:54-57) + BuildOctree (scene.cpp:75-83) from a raw array;
// mn/mx are the MODEL bounds (the octree root gets +-0.7*size like main.cpp:297, 312).
void* ref_scene_from_tris(const float* tris9, int n, const float mn[3], const float mx[3]) {
    static_assert(sizeof(Triangle) == 9 * sizeof(float), "Triangle is 3 vec3 (maths.h:56-59)");
    auto* rs = new RefScene();
    rs->scene = std::make_unique<Scene>(reinterpret_cast<const Triangle*>(tris9), n);
    rs->mn = glm::vec3(mn[0], mn[1], mn[2]);
    rs->mx = glm::vec3(mx[0], mx[1], mx[2]);
    glm::vec3 extra = (rs->mx - rs->mn) * 0.7f;
    rs->scene->BuildOctree(rs->mn - extra, rs->mx + extra);
    return rs;
}

void ref_scene_free(void* h) { delete static_cast<RefScene*>(h); }

int ref_scene_triangles(void* h, float* out9) {
    auto* rs = static_cast<RefScene*>(h);
    const auto& t = rs->scene->m_triangles;
    if (out9) std::memcpy(out9, t.data(), t.size() * sizeof(Triangle));
    return int(t.size());
}

// Scene::HitScene through the octree (scene.cpp:86-97).  outFlag gets the
// reference's own return value: 1 on hit, -1 on miss.  Hit fields are written
// only on a hit (the caller pre-fills them).
void ref_hit_scene(void* h, const float* rays6, long n, float tMin, float tMax,
                   int* outFlag, float* outT, float* outPos, float* outNormal) {
    auto* rs = static_cast<RefScene*>(h);
    for (long i = 0; i < n; ++i) {
        Ray r = make_ray(rays6 + i * 6);
        Hit hit;
        int id = rs->scene->HitScene(r, tMin, tMax, hit);
        outFlag[i] = id;
        if (id != -1) store_hit(hit, i, outT, outPos, outNormal);
    }
}

// ID-carrying brute force over the INPUT triangle order using the reference's own
// RayIntersectTriangleImproved (maths.cpp:340-380) and the octree path's accept
// rule `hit.t < hitMinT`, hitMinT initialised to tMax (scene.cpp:34, 90): first
// tested wins ties, so the index is the lowest among bit-equal nearest t.
void ref_hit_brute(void* h, const float* rays6, long n, float tMin, float tMax,
                   int* outID, float* outT, float* outPos, float* outNormal) {
    auto* rs = static_cast<RefScene*>(h);
    const auto& tris = rs->scene->m_triangles;
    for (long i = 0; i < n; ++i) {
        Ray r = make_ray(rays6 + i * 6);
        int best = -1;
        float bestT = tMax;
        Hit bestHit;
        for (size_t k = 0; k < tris.size(); ++k) {
            Hit hit;
            if (RayIntersectTriangleImproved(r, tris[k], tMin, tMax, hit) && hit.t < bestT) {
                bestT = hit.t;
                best = int(k);
                bestHit = hit;
            }
        }
        outID[i] = best;
        if (best != -1) store_hit(bestHit, i, outT, outPos, outNormal);
    }
}

// RNG known-answer generators (maths.cpp:5-38).  XorShift32 is file-static there,
// so raw states are observed through the state word RandomFloat01 leaves behind.
void ref_rng_states(uint32_t seed, int n, uint32_t* outStates, float* outFloats) {
    uint32_t s = seed;
    for (int i = 0; i < n; ++i) {
        float f = RandomFloat01(s);
        if (outStates) outStates[i] = s;
        if (outFloats) outFloats[i] = f;
    }
}

uint32_t ref_random_unit_vectors(uint32_t seed, int n, float* out3) {
    uint32_t s = seed;
    for (int i = 0; i < n; ++i) {
        glm::vec3 v = RandomUnitVector(s);
        out3[i * 3 + 0] = v.x; out3[i * 3 + 1] = v.y; out3[i * 3 + 2] = v.z;
    }
    return s;
}

uint32_t ref_random_in_unit_disk(uint32_t seed, int n, float* out3) {
    uint32_t s = seed;
    for (int i = 0; i < n; ++i) {
        glm::vec3 v = RandomInUnitDisk(s);
        out3[i * 3 + 0] = v.x; out3[i * 3 + 1] = v.y; out3[i * 3 + 2] = v.z;
    }
    return s;
}

// Camera::Camera (maths.cpp:40-59) -> 22 floats in declaration order (maths.h:106-111).
void ref_camera_make(const float from[3], const float at[3], const float up[3], float vfov,
                     float aspect, float aperture, float focusDist, float out22[22]) {
    Camera cam(glm::vec3(from[0], from[1], from[2]), glm::vec3(at[0], at[1], at[2]),
               glm::vec3(up[0], up[1], up[2]), vfov, aspect, aperture, focusDist);
    std::memcpy(out22, &cam, sizeof(Camera));
}

// Camera placement.  This is INSIDE main() in the reference (main.cpp:296-307), so
// it cannot be called; the lines are replayed here with the same expressions.  The
// full-binary image hashes in tests/golden/ pin this replay.
void ref_camera_for_scene(void* h, const char* objPath, int w, int hgt, float out22[22]) {
    auto* rs = static_cast<RefScene*>(h);
    glm::vec3 sceneMin = rs->mn, sceneMax = rs->mx;
    glm::vec3 sceneSize = sceneMax - sceneMin;
    glm::vec3 sceneCenter = (sceneMin + sceneMax) * 0.5f;
    glm::vec3 lookfrom = sceneCenter + sceneSize * glm::vec3(0.3f, 0.6f, 1.2f);
    if (strstr(objPath, "sponza.obj") != nullptr) lookfrom = glm::vec3(-5.96f, 4.08f, -1.22f);
    glm::vec3 lookat = sceneCenter + sceneSize * glm::vec3(0.0f, -0.1f, 0.0f);
    const float distToFocus = length(lookfrom - lookat);
    const float aperture = 0.03f;
    auto camera = Camera(lookfrom, lookat, glm::vec3(0.0f, 1.0f, 0.0f), 60.0f,
                         float(w) / float(hgt), aperture, distToFocus);
    std::memcpy(out22, &camera, sizeof(Camera));
}

// Camera::GetRay (maths.h:93-104) for n (s,t) pairs drawn from one RNG stream.
uint32_t ref_camera_get_rays(const float cam22[22], const float* st2, int n, uint32_t seed, float* outRays6) {
    Camera cam = camera_from22(cam22);
    uint32_t s = seed;
    for (int i = 0; i < n; ++i) {
        Ray r = cam.GetRay(st2[i * 2 + 0], st2[i * 2 + 1], s);
        outRays6[i * 6 + 0] = r.orig.x; outRays6[i * 6 + 1] = r.orig.y; outRays6[i * 6 + 2] = r.orig.z;
        outRays6[i * 6 + 3] = r.dir.x;  outRays6[i * 6 + 4] = r.dir.y;  outRays6[i * 6 + 5] = r.dir.z;
    }
    return s;
}

// Record the rays a reference render actually shoots, for the fixed hit-ID ray set
// (SURVEY.md 8(d)): for pixels on a stride grid, one camera sample each, follow the
// path with the reference's own Scatter (main.cpp:44-73) and HitScene, and append
// every queried ray (kind 0 primary, 1 bounce, 2 shadow).  Returns rays written.
long ref_record_path_rays(void* h, const float cam22[22], int w, int hgt, int stride,
                          long maxRays, float* outRays6, int* outKind) {
    auto* rs = static_cast<RefScene*>(h);
    const Scene& scene = *rs->scene;
    Camera cam = camera_from22(cam22);
    long n = 0;
    const float invW = 1.0f / w, invH = 1.0f / hgt;
    auto push = [&](const Ray& r, int kind) {
        if (n >= maxRays) return;
        float* o = outRays6 + n * 6;
        o[0] = r.orig.x; o[1] = r.orig.y; o[2] = r.orig.z; o[3] = r.dir.x; o[4] = r.dir.y; o[5] = r.dir.z;
        outKind[n] = kind;
        ++n;
    };
    for (int y = 0; y < hgt; y += stride) {
        uint32_t rng = uint32_t(y) * 9781 + 1;  // main.cpp:204
        for (int x = 0; x < w; x += stride) {
            float fu = (float(x) + RandomFloat01(rng)) * invW;
            float fv = (float(y) + RandomFloat01(rng)) * invH;
            Ray ray = cam.GetRay(fu, fv, rng);
            int kind = 0;
            for (int depth = 0; depth < kMaxDepth; ++depth) {
                push(ray, kind);
                Hit hit;
                if (scene.HitScene(ray, kMinT, kMaxT, hit) == -1) break;
                push(Ray(hit.pos, kLightDir), 2);
                glm::vec3 att, light;
                int rc = 0;
                ray = Scatter(ray, scene, hit, att, light, rng, rc);
                kind = 1;
            }
        }
    }
    return n;
}

// The reference's own row functor (main.cpp:180-246) over all rows, scheduled by the
// TBB shim.  rgba is caller-owned w*h*4, row 0 = bottom (not flipped).
void ref_render(void* h, const float cam22[22], int w, int hgt, int spp, uint8_t* rgba, long long* rayCount) {
    auto* rs = static_cast<RefScene*>(h);
    Camera cam = camera_from22(cam22);
    TraceData data;
    data.screenWidth = w;
    data.screenHeight = hgt;
    data.samplesPerPixel = spp;
    data.image = rgba;
    data.camera = &cam;
    data.rayCount = 0;
    tbb::parallel_for(tbb::blocked_range<int64_t>(0, hgt, 1), TraceImageBody(&data, rs->scene.get()));
    *rayCount = data.rayCount;
}

// kLightDir as the reference's static initialiser computed it (main.cpp:36).
void ref_light_dir(float out3[3]) { out3[0] = kLightDir.x; out3[1] = kLightDir.y; out3[2] = kLightDir.z; }

int ref_threads() { return tbb::task_scheduler_init::default_num_threads(); }

}  // extern "C"
