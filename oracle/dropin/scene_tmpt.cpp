// scene_tmpt.cpp -- the reference-side binding: a replacement for the reference's source/scene.cpp.
//
// The reference has no plugin / FFI layer; its seam is `struct Scene` (source/scene.h:17-43).  This file implements that struct's
// member functions -- the declarations stay the reference's own, scene.h is compiled UNMODIFIED -- on top of the C ABI of
// include/tmpt.h, so that the reference's own main.cpp (argument parsing, LoadScene, camera, Trace / Scatter, TBB row loop, PNG)
// links against libtmpt.so instead of its octree:
//
//     g++ <ref>/source/main.cpp <ref>/source/maths.cpp <ref>/source/external/objparser.cpp oracle/dropin/scene_tmpt.cpp
//         -I<ref>/source -I<repo>/include -L<repo>/toymeshpathtracer_b200 -ltmpt          (oracle/Makefile: `make dropin`)
//
// What it demonstrates (tests/test_zz_gpu_fuzz_regressions.py::test_reference_program_with_the_scene_class_swapped): with every
// HitScene answered by the GPU library -- flag, Hit.pos, Hit.normal, Hit.t bit for bit -- the reference's integrator walks the
// same paths and writes the same output.png, byte for byte, and counts the same rays.  One ray per call is of course the slowest
// possible way to use a GPU (about 40 us per query: a copy in, a launch, a copy out); it is the smallest possible patch, not the
// fast one -- INTEGRATION.md B replaces the row loop by tmpt_render.
//
// TEST INFRASTRUCTURE: built into oracle/_ref/ only, never part of libtmpt.so.
#include <cstdio>
#include <cstdlib>

#include "scene.h"  // the reference's, unmodified
#include "tmpt.h"

static_assert(sizeof(Triangle) == 9 * sizeof(float), "Triangle is three packed vec3 (maths.h:56-59): tmpt_scene_create reads it as float[9]");
static_assert(sizeof(Ray) == 6 * sizeof(float), "Ray is orig + dir (maths.h:31-40): tmpt_hit_scene reads it as float[6]");

// scene.h only forward-declares OctreeNode and keeps a unique_ptr to it: here it holds the library's scene handle.
struct OctreeNode
{
    tmpt_scene* handle = nullptr;
    ~OctreeNode() { tmpt_scene_destroy(handle); }
};

static void die(const char* what)
{
    printf("ERROR: %s: %s\n", what, tmpt_last_error());  // main.cpp's error style: a line on stdout, exit code 1
    exit(1);
}

Scene::Scene(const Triangle* triangles, int triangleCount)  // scene.cpp:54-57
{
    m_triangles.assign(triangles, triangles + triangleCount);
}

Scene::~Scene() = default;

void Scene::Cull(const glm::vec3&) {}  // dead code in the reference (main.cpp:309-310 keeps it commented out)

// scene.cpp:75-83 builds the octree inside [min, max]; the library builds its BVH over the triangles themselves.
void Scene::BuildOctree(const glm::vec3&, const glm::vec3&)
{
    m_octree = std::make_unique<OctreeNode>();
    const char* dev = getenv("TMPT_DEVICE");
    if (tmpt_scene_create(reinterpret_cast<const float*>(m_triangles.data()), int(m_triangles.size()), dev ? atoi(dev) : 0,
                          TMPT_BUILD_DEFAULT, &m_octree->handle) != TMPT_OK)
        die("tmpt_scene_create");
}

// scene.cpp:86-97.  Returns what the fork returns: 1 for a hit, -1 for a miss (scene.cpp:37); outHit is written only for a hit.
int Scene::HitScene(const Ray& ray, float tMin, float tMax, Hit& outHit) const
{
    const float r6[6] = {ray.orig.x, ray.orig.y, ray.orig.z, ray.dir.x, ray.dir.y, ray.dir.z};
    int32_t id = -1;
    float t = 0.0f, pos[3], normal[3];
    if (tmpt_hit_scene(m_octree->handle, r6, 1, tMin, tMax, TMPT_HIT_CLOSEST, TMPT_HOST, &id, &t, pos, normal, nullptr) != TMPT_OK)
        die("tmpt_hit_scene");
    if (id < 0)
        return -1;
    outHit.t = t;
    outHit.pos = glm::vec3(pos[0], pos[1], pos[2]);
    outHit.normal = glm::vec3(normal[0], normal[1], normal[2]);
    return 1;
}
