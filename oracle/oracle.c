/* oracle/oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.  See oracle.h.
 *
 * Plain-C restatement of the reference's Trace() hot path.  Every function names
 * the reference lines it follows (paths relative to /root/reference/source).
 * Arithmetic is IEEE binary32, one rounding per operation, in the operation order
 * glm 0.9.9.5's scalar path fixes (external/glm/detail/func_geometric.inl:48-90);
 * build with -ffp-contract=off (oracle/Makefile does).
 *
 * Where the reference leaves the order of two RNG draws to the compiler
 * (main.cpp:214-215, maths.cpp:25) this file follows what g++ 13.3 does on x86-64
 * -- the LAST argument is evaluated first -- because "the reference image" means
 * that build; tests pin it against oracle/_ref.
 *
 * Nearest-hit search is the ID-carrying brute force over the input array
 * (SURVEY.md 8(c)): same per-triangle test and accept rule as the reference's
 * octree walk, index order, so the winner is the lowest index among bit-equal
 * nearest t.  Rows of a render are spread over pthreads; results do not depend
 * on the thread count.
 */
#include "oracle.h"

#include <math.h>
#include <pthread.h>
#include <stdlib.h>
#include <string.h>

typedef struct { float x, y, z; } v3;

/* ---- glm scalar-path vector ops (one rounding per flop, left to right) ---- */
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 add(v3 a, v3 b) { return V(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mulv(v3 a, v3 b) { return V(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 muls(v3 a, float s) { return V(a.x * s, a.y * s, a.z * s); }
static inline v3 neg(v3 a) { return V(-a.x, -a.y, -a.z); }
/* func_geometric.inl:48-55: tmp = a*b; tmp.x + tmp.y + tmp.z */
static inline float dot(v3 a, v3 b) { float x = a.x * b.x, y = a.y * b.y, z = a.z * b.z; return (x + y) + z; }
/* func_geometric.inl:68-79 */
static inline v3 cross(v3 x, v3 y) {
    return V(x.y * y.z - y.y * x.z, x.z * y.x - y.z * x.x, x.x * y.y - y.x * x.y);
}
/* func_geometric.inl:82-90 with func_exponential.inl:136-139: v * (1 / sqrt(dot(v,v))) */
static inline v3 normalize(v3 v) { float s = 1.0f / sqrtf(dot(v, v)); return muls(v, s); }
static inline float length3(v3 v) { return sqrtf(dot(v, v)); }
/* func_common.inl:17-30 */
static inline float glm_min(float x, float y) { return (y < x) ? y : x; }
static inline float glm_max(float x, float y) { return (x < y) ? y : x; }

/* ---- RNG: maths.cpp:5-38 ---- */
static inline uint32_t xorshift32(uint32_t* state) { /* maths.cpp:5-13, shifts 13/17/15 */
    uint32_t x = *state;
    x ^= x << 13;
    x ^= x >> 17;
    x ^= x << 15;
    *state = x;
    return x;
}
static inline float random_float01(uint32_t* state) { /* maths.cpp:15-18 */
    return (float)(xorshift32(state) & 0xFFFFFF) / 16777216.0f;
}
static v3 random_in_unit_disk(uint32_t* state) { /* maths.cpp:20-28 */
    v3 p;
    do {
        /* vec3(R(), R(), 0): g++ evaluates the second argument first */
        float ry = random_float01(state);
        float rx = random_float01(state);
        p = V(2.0f * rx - 1.0f, 2.0f * ry - 1.0f, 2.0f * 0.0f - 0.0f);
    } while (dot(p, p) >= 1.0f);
    return p;
}

/* "Trig spec" (DESIGN.md): sin and cos of a in [0, 8) computed in binary64 with
 * plain (unfused) multiply/add only, so that a CPU and a GPU produce the same
 * bits.  k = floor(a*2/pi + 1/2); y = (a - k*P1) - k*P2 (Cody-Waite, k*P1 exact
 * for k < 2^20); Taylor polynomials to y^15 / y^16 in Horner form; quadrant
 * fix-up; round once to binary32.  |error| < 1e-15 before the final rounding. */
void orc_sincos_spec(float a, float* s, float* c) {
    static const double TWO_OVER_PI = 0.63661977236758134308;
    static const double P1 = 1.57079632673412561417e+00; /* 0x3FF921FB54400000: pi/2 to 33 bits */
    static const double P2 = 6.07710050650619224932e-11; /* pi/2 - P1 */
    double x = (double)a;
    double kd = floor(x * TWO_OVER_PI + 0.5);
    double y = (x - kd * P1) - kd * P2;
    double z = y * y;
    /* sin y = y + y*z*(S3 + z*(S5 + ... z*S15)) */
    double ps = -1.0 / 1307674368000.0;           /* -1/15! */
    ps = ps * z + 1.0 / 6227020800.0;             /* +1/13! */
    ps = ps * z + -1.0 / 39916800.0;              /* -1/11! */
    ps = ps * z + 1.0 / 362880.0;                 /* +1/9!  */
    ps = ps * z + -1.0 / 5040.0;                  /* -1/7!  */
    ps = ps * z + 1.0 / 120.0;                    /* +1/5!  */
    ps = ps * z + -1.0 / 6.0;                     /* -1/3!  */
    double sy = y + y * (z * ps);
    /* cos y = 1 + z*(C2 + z*(C4 + ... z*C16)) */
    double pc = 1.0 / 20922789888000.0;           /* +1/16! */
    pc = pc * z + -1.0 / 87178291200.0;           /* -1/14! */
    pc = pc * z + 1.0 / 479001600.0;              /* +1/12! */
    pc = pc * z + -1.0 / 3628800.0;               /* -1/10! */
    pc = pc * z + 1.0 / 40320.0;                  /* +1/8!  */
    pc = pc * z + -1.0 / 720.0;                   /* -1/6!  */
    pc = pc * z + 1.0 / 24.0;                     /* +1/4!  */
    pc = pc * z + -0.5;                           /* -1/2!  */
    double cy = 1.0 + z * pc;
    int q = (int)kd & 3;
    double sv = (q == 0) ? sy : (q == 1) ? cy : (q == 2) ? -sy : -cy;
    double cv = (q == 0) ? cy : (q == 1) ? -sy : (q == 2) ? -cy : sy;
    *s = (float)sv;
    *c = (float)cv;
}

static v3 random_unit_vector(uint32_t* state, int trig) { /* maths.cpp:30-38; kPI maths.h:14 */
    const float kPI = 3.1415926f;
    float z = random_float01(state) * 2.0f - 1.0f;
    float a = random_float01(state) * 2.0f * kPI;
    float r = sqrtf(1.0f - z * z);
    float sa, ca;
    if (trig == ORC_TRIG_SPEC) orc_sincos_spec(a, &sa, &ca);
    else { ca = cosf(a); sa = sinf(a); }
    float x = r * ca;
    float y = r * sa;
    return V(x, y, z);
}

/* Seed of the per-pixel stream used by the GPU path (ORC_RNG_PIXEL): the reference's
 * row seed expression (main.cpp:204) applied to the pixel index, scrambled with
 * Wang's 32-bit hash so neighbouring pixels decorrelate, never 0 (xorshift fixed
 * point).  DESIGN.md "RNG". */
uint32_t orc_stream_seed(uint64_t streamIndex) {
    /* The stream index chunk * width * height + pixelIndex needs more than 32 bits for large frames at high spp
     * (10000 x 10000 from chunk 43 on) and for long progressive renders: the high word is folded in before the
     * hash, so that streams beyond 2^32 do not repeat the first 2^32 in order; below 2^32 (hi = 0) this is
     * orc_pixel_seed of the index. */
    const uint32_t lo = (uint32_t)streamIndex, hi = (uint32_t)(streamIndex >> 32);
    uint32_t s = (lo * 9781u + 1u) ^ (hi * 0x9E3779B9u);
    s = (s ^ 61u) ^ (s >> 16);
    s *= 9u;
    s ^= s >> 4;
    s *= 0x27d4eb2du;
    s ^= s >> 15;
    return s ? s : 1u;
}
uint32_t orc_pixel_seed(uint32_t pixelIndex) {
    uint32_t s = pixelIndex * 9781u + 1u;
    s = (s ^ 61u) ^ (s >> 16);
    s *= 9u;
    s ^= s >> 4;
    s *= 0x27d4eb2du;
    s ^= s >> 15;
    return s ? s : 1u;
}

void orc_rng_states(uint32_t seed, int n, uint32_t* outStates, float* outFloats) {
    uint32_t s = seed;
    for (int i = 0; i < n; ++i) {
        float f = random_float01(&s);
        if (outStates) outStates[i] = s;
        if (outFloats) outFloats[i] = f;
    }
}
uint32_t orc_random_in_unit_disk(uint32_t seed, int n, float* out3) {
    uint32_t s = seed;
    for (int i = 0; i < n; ++i) { v3 p = random_in_unit_disk(&s); out3[i * 3] = p.x; out3[i * 3 + 1] = p.y; out3[i * 3 + 2] = p.z; }
    return s;
}
uint32_t orc_random_unit_vectors(uint32_t seed, int n, float* out3, int trig) {
    uint32_t s = seed;
    for (int i = 0; i < n; ++i) { v3 p = random_unit_vector(&s, trig); out3[i * 3] = p.x; out3[i * 3 + 1] = p.y; out3[i * 3 + 2] = p.z; }
    return s;
}

/* ---- camera: maths.cpp:40-59, maths.h:93-111 ---- */
typedef struct {
    v3 origin, lowerLeftCorner, horizontal, vertical, u, v, w;
    float lensRadius;
} camera_t;

static camera_t camera_make(v3 lookFrom, v3 lookAt, v3 vup, float vfov, float aspect, float aperture, float focusDist) {
    const float kPI = 3.1415926f;
    camera_t c;
    c.lensRadius = aperture * 0.5f;
    float theta = vfov * kPI / 180.0f;
    float halfHeight = tanf(theta * 0.5f);
    float halfWidth = aspect * halfHeight;
    c.origin = lookFrom;
    c.w = normalize(sub(lookFrom, lookAt));
    c.u = normalize(cross(vup, c.w));
    c.v = cross(c.w, c.u);
    /* origin - halfWidth*focusDist*u - halfHeight*focusDist*v - focusDist*w  (scalar products first) */
    c.lowerLeftCorner = sub(sub(sub(c.origin, muls(c.u, halfWidth * focusDist)), muls(c.v, halfHeight * focusDist)),
                            muls(c.w, focusDist));
    c.horizontal = muls(c.u, 2.0f * halfWidth * focusDist);
    c.vertical = muls(c.v, 2.0f * halfHeight * focusDist);
    return c;
}

typedef struct { v3 orig, dir; } ray_t;

static ray_t camera_get_ray(const camera_t* c, float s, float t, uint32_t* state) { /* maths.h:93-104 */
    v3 rd = muls(random_in_unit_disk(state), c->lensRadius);
    v3 offset = add(muls(c->u, rd.x), muls(c->v, rd.y));
    ray_t r;
    r.orig = add(c->origin, offset);
    r.dir = normalize(sub(sub(add(add(c->lowerLeftCorner, muls(c->horizontal, s)), muls(c->vertical, t)), c->origin), offset));
    return r;
}

static camera_t camera_from22(const float* f) { camera_t c; memcpy(&c, f, sizeof c); return c; }

void orc_camera_make(const float from[3], const float at[3], const float up[3], float vfov,
                     float aspect, float aperture, float focusDist, float out22[22]) {
    camera_t c = camera_make(V(from[0], from[1], from[2]), V(at[0], at[1], at[2]), V(up[0], up[1], up[2]),
                             vfov, aspect, aperture, focusDist);
    memcpy(out22, &c, sizeof c);
}

void orc_camera_for_scene(const float mn[3], const float mx[3], int isSponza, int w, int h, float out22[22]) {
    /* main.cpp:296-307 */
    v3 sceneMin = V(mn[0], mn[1], mn[2]), sceneMax = V(mx[0], mx[1], mx[2]);
    v3 sceneSize = sub(sceneMax, sceneMin);
    v3 sceneCenter = muls(add(sceneMin, sceneMax), 0.5f);
    v3 lookfrom = add(sceneCenter, mulv(sceneSize, V(0.3f, 0.6f, 1.2f)));
    if (isSponza) lookfrom = V(-5.96f, 4.08f, -1.22f);
    v3 lookat = add(sceneCenter, mulv(sceneSize, V(0.0f, -0.1f, 0.0f)));
    float distToFocus = length3(sub(lookfrom, lookat));
    camera_t c = camera_make(lookfrom, lookat, V(0.0f, 1.0f, 0.0f), 60.0f, (float)w / (float)h, 0.03f, distToFocus);
    memcpy(out22, &c, sizeof c);
}

uint32_t orc_camera_get_rays(const float cam22[22], const float* st2, int n, uint32_t seed, float* out) {
    camera_t c = camera_from22(cam22);
    uint32_t s = seed;
    for (int i = 0; i < n; ++i) {
        ray_t r = camera_get_ray(&c, st2[i * 2], st2[i * 2 + 1], &s);
        out[i * 6 + 0] = r.orig.x; out[i * 6 + 1] = r.orig.y; out[i * 6 + 2] = r.orig.z;
        out[i * 6 + 3] = r.dir.x;  out[i * 6 + 4] = r.dir.y;  out[i * 6 + 5] = r.dir.z;
    }
    return s;
}

/* ---- scene ingest: main.cpp:132-162 ---- */
int orc_add_floor(const float* model9, int n, float* out9, float mn[3], float mx[3]) {
    v3 bmin = V(+1.0e6f, +1.0e6f, +1.0e6f), bmax = V(-1.0e6f, -1.0e6f, -1.0e6f);
    for (int i = 0; i < n * 3; ++i) {
        v3 p = V(model9[i * 3], model9[i * 3 + 1], model9[i * 3 + 2]);
        bmin = V(glm_min(bmin.x, p.x), glm_min(bmin.y, p.y), glm_min(bmin.z, p.z));
        bmax = V(glm_max(bmax.x, p.x), glm_max(bmax.y, p.y), glm_max(bmax.z, p.z));
    }
    memcpy(out9, model9, (size_t)n * 9 * sizeof(float));
    v3 size = sub(bmax, bmin);
    v3 extra = muls(size, 0.7f);
    float* f = out9 + (size_t)n * 9;
    float x0 = bmin.x - extra.x, x1 = bmax.x + extra.x, z0 = bmin.z - extra.z, z1 = bmax.z + extra.z, y = bmin.y;
    float a[18] = {x0, y, z0, x0, y, z1, x1, y, z0,   /* main.cpp:157-159 */
                   x0, y, z1, x1, y, z1, x1, y, z0};  /* main.cpp:160-162 */
    memcpy(f, a, sizeof a);
    mn[0] = bmin.x; mn[1] = bmin.y; mn[2] = bmin.z;
    mx[0] = bmax.x; mx[1] = bmax.y; mx[2] = bmax.z;
    return n + 2;
}

static v3 light_dir(void) { return normalize(V(-0.7f, 1.0f, 0.5f)); } /* main.cpp:36 */
void orc_light_dir(float out3[3]) { v3 l = light_dir(); out3[0] = l.x; out3[1] = l.y; out3[2] = l.z; }

/* ---- ray-triangle: maths.cpp:339-380 ---- */
typedef struct { v3 pos, normal; float t; } hit_t;

static int ray_intersect_triangle_improved(const ray_t* r, const float* tri, float tMin, float tMax, hit_t* out) {
    const float Epsilon = 1e-5f;
    v3 v0 = V(tri[0], tri[1], tri[2]), v1 = V(tri[3], tri[4], tri[5]), v2 = V(tri[6], tri[7], tri[8]);
    v3 edge1 = sub(v1, v0);
    v3 edge2 = sub(v2, v0);
    v3 pvec = cross(r->dir, edge2);
    float det = dot(edge1, pvec);
    if (det > -Epsilon && det < Epsilon) return 0;
    float invDet = 1.0f / det;
    v3 tvec = sub(r->orig, v0);
    float u = dot(tvec, pvec) * invDet;
    if (u < 0.0f || u > 1.0f) return 0;
    v3 qvec = cross(tvec, edge1);
    float v = dot(r->dir, qvec) * invDet;
    if (v < 0.0f || u + v > 1.0f) return 0;
    float t = dot(edge2, qvec) * invDet;
    if (t >= tMin && t <= tMax) {
        out->t = t;
        out->pos = add(add(muls(v0, 1.0f - u - v), muls(v1, u)), muls(v2, v));
        out->normal = normalize(cross(edge1, edge2));
        return 1;
    }
    return 0;
}

/* scene.cpp:86-97 + :29-41 semantics on the flat input array; returns index or -1 */
static int hit_scene(const float* tris9, int n, const ray_t* r, float tMin, float tMax, int anyHit, hit_t* outHit) {
    int id = -1;
    float hitMinT = tMax;
    for (int i = 0; i < n; ++i) {
        hit_t h;
        if (ray_intersect_triangle_improved(r, tris9 + (size_t)i * 9, tMin, tMax, &h)) {
            if (h.t < hitMinT) {
                hitMinT = h.t;
                id = i;
                *outHit = h;
                if (anyHit) break;
            }
        }
    }
    return id;
}

/* ---- integrator: main.cpp:30-37, 44-119 ---- */
#define K_MAX_DEPTH 10
static const float kMinT = 0.001f;
static const float kMaxT = 1.0e7f;

typedef struct {
    const float* tris9;
    int nTris;
    v3 lightDir;
    int trig;
} scene_t;

static ray_t scatter(const scene_t* sc, const ray_t* r, const hit_t* hit, v3* outAtten, v3* outLight,
                     uint32_t* rng, int64_t* rayCount) { /* main.cpp:44-73 */
    const v3 kLightColor = V(0.7f, 0.6f, 0.5f);
    *outLight = V(0.0f, 0.0f, 0.0f);
    v3 albedo = V(0.7f, 0.7f, 0.7f);
    *outAtten = albedo;
    ++*rayCount;
    hit_t lightHit;
    ray_t shadow; shadow.orig = hit->pos; shadow.dir = sc->lightDir;
    int id = hit_scene(sc->tris9, sc->nTris, &shadow, kMinT, kMaxT, 0, &lightHit);
    if (id == -1) {
        v3 nl = dot(hit->normal, r->dir) < 0 ? hit->normal : neg(hit->normal);
        float k = fmaxf(0.0f, dot(sc->lightDir, nl));
        *outLight = add(*outLight, muls(mulv(albedo, kLightColor), k));
    }
    v3 target = add(add(hit->pos, hit->normal), random_unit_vector(rng, sc->trig));
    ray_t out; out.orig = hit->pos; out.dir = normalize(sub(target, hit->pos));
    return out;
}

static v3 trace(const scene_t* sc, ray_t ray, uint32_t* rng, int64_t* rayCount) { /* main.cpp:82-119 */
    v3 light[K_MAX_DEPTH], atten[K_MAX_DEPTH];
    int depth = 0;
    v3 color = V(0.0f, 0.0f, 0.0f);
    while (depth < K_MAX_DEPTH) {
        ++*rayCount;
        hit_t hit;
        memset(&hit, 0, sizeof hit);
        int id = hit_scene(sc->tris9, sc->nTris, &ray, kMinT, kMaxT, 0, &hit);
        if (id != -1) {
            ray = scatter(sc, &ray, &hit, &atten[depth], &light[depth], rng, rayCount);
            ++depth;
        } else {
            float t = 0.5f * (ray.dir.y + 1.0f);
            color = muls(add(muls(V(1.0f, 1.0f, 1.0f), 1.0f - t), muls(V(0.5f, 0.7f, 1.0f), t)), 0.5f);
            break;
        }
    }
    for (int i = depth - 1; i >= 0; --i) color = add(light[i], mulv(atten[i], color));
    return color;
}

/* main.cpp:16-19 + 230-232: uint8_t(saturate(c) * 255.0f).  NaN propagates through
 * glm's clamp; uint8_t(NaN) is UB in C++ and yields 0 with g++/x86-64,
 * which is what this returns. */
static uint8_t quantise(float c) {
    float s = glm_min(glm_max(c, 0.0f), 1.0f) * 255.0f;
    if (s != s) return 0;
    return (uint8_t)s;
}

typedef struct {
    scene_t sc;
    camera_t cam;
    int w, h, spp, rngMode, row1;
    uint8_t* rgba;
    float* linear;
    int next_row; /* guarded by mu */
    int64_t rays;
    pthread_mutex_t mu;
} render_job;

static void render_row(render_job* j, int y, int64_t* rayCount) { /* main.cpp:202-235 */
    const float invWidth = 1.0f / (float)j->w, invHeight = 1.0f / (float)j->h;
    const float sppRecip = 1.0f / (float)j->spp;
    uint32_t rng = (uint32_t)y * 9781u + 1u; /* main.cpp:204 */
    for (int x = 0; x < j->w; ++x) {
        v3 col = V(0.0f, 0.0f, 0.0f);
        /* ORC_RNG_ROW: one chunk holding every sample, the row's stream flowing on (the reference).
         * ORC_RNG_PIXEL: chunks of ORC_CHUNK_SAMPLES(spp) samples, each with its own stream seeded from
         * (chunk, pixel); chunk sums are added in chunk order (DESIGN.md "RNG"). */
        const int chunkLen = j->rngMode == ORC_RNG_PIXEL ? ORC_CHUNK_SAMPLES(j->spp) : j->spp;
        for (int s0 = 0, c = 0; s0 < j->spp; s0 += chunkLen, ++c) {
            if (j->rngMode == ORC_RNG_PIXEL)
                rng = orc_stream_seed((uint64_t)c * ((uint64_t)j->w * (uint64_t)j->h) + (uint64_t)y * (uint64_t)j->w + (uint64_t)x);
            v3 chunk = V(0.0f, 0.0f, 0.0f);
            const int s1 = s0 + chunkLen < j->spp ? s0 + chunkLen : j->spp;
            for (int s = s0; s < s1; ++s) {
                /* GetRay(argU, argV, rng): g++ evaluates argV (second) first */
                float fv = ((float)y + random_float01(&rng)) * invHeight;
                float fu = ((float)x + random_float01(&rng)) * invWidth;
                ray_t ray = camera_get_ray(&j->cam, fu, fv, &rng);
                chunk = add(chunk, trace(&j->sc, ray, &rng, rayCount));
            }
            col = j->rngMode == ORC_RNG_PIXEL ? add(col, chunk) : chunk;
        }
        col = muls(col, sppRecip);
        size_t p = (size_t)y * (size_t)j->w + (size_t)x;
        if (j->linear) { j->linear[p * 3] = col.x; j->linear[p * 3 + 1] = col.y; j->linear[p * 3 + 2] = col.z; }
        col = V(sqrtf(col.x), sqrtf(col.y), sqrtf(col.z));
        j->rgba[p * 4 + 0] = quantise(col.x);
        j->rgba[p * 4 + 1] = quantise(col.y);
        j->rgba[p * 4 + 2] = quantise(col.z);
        j->rgba[p * 4 + 3] = 255;
    }
}

static void* render_worker(void* arg) {
    render_job* j = (render_job*)arg;
    int64_t rays = 0;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        int y = j->next_row++;
        pthread_mutex_unlock(&j->mu);
        if (y >= j->row1) break;
        render_row(j, y, &rays);
    }
    pthread_mutex_lock(&j->mu);
    j->rays += rays;
    pthread_mutex_unlock(&j->mu);
    return NULL;
}

static void run_threads(void* (*fn)(void*), void* arg, int threads) {
    if (threads < 1) threads = 1;
    if (threads > 256) threads = 256;
    pthread_t th[256];
    for (int i = 1; i < threads; ++i) pthread_create(&th[i], NULL, fn, arg);
    fn(arg);
    for (int i = 1; i < threads; ++i) pthread_join(th[i], NULL);
}

void orc_render(const float* tris9, int nTris, const float cam22[22], int w, int h, int spp,
                int rngMode, int trig, int row0, int row1, uint8_t* rgba, float* outLinear,
                int64_t* rayCount, int threads) {
    render_job j;
    j.sc.tris9 = tris9; j.sc.nTris = nTris; j.sc.lightDir = light_dir(); j.sc.trig = trig;
    j.cam = camera_from22(cam22);
    j.w = w; j.h = h; j.spp = spp; j.rngMode = rngMode; j.row1 = row1;
    j.rgba = rgba; j.linear = outLinear; j.next_row = row0; j.rays = 0;
    pthread_mutex_init(&j.mu, NULL);
    run_threads(render_worker, &j, threads);
    pthread_mutex_destroy(&j.mu);
    if (rayCount) *rayCount = j.rays;
}

typedef struct {
    const float* tris9; int nTris; const float* rays6; long nRays; float tMin, tMax; int anyHit;
    int* outID; float* outT; float* outPos; float* outNormal;
    long next; pthread_mutex_t mu;
} hit_job;

static void* hit_worker(void* arg) {
    hit_job* j = (hit_job*)arg;
    const long chunk = 256;
    for (;;) {
        pthread_mutex_lock(&j->mu);
        long b = j->next; j->next += chunk;
        pthread_mutex_unlock(&j->mu);
        if (b >= j->nRays) break;
        long e = b + chunk < j->nRays ? b + chunk : j->nRays;
        for (long i = b; i < e; ++i) {
            ray_t r;
            r.orig = V(j->rays6[i * 6], j->rays6[i * 6 + 1], j->rays6[i * 6 + 2]);
            r.dir = V(j->rays6[i * 6 + 3], j->rays6[i * 6 + 4], j->rays6[i * 6 + 5]);
            hit_t h;
            memset(&h, 0, sizeof h);
            int id = hit_scene(j->tris9, j->nTris, &r, j->tMin, j->tMax, j->anyHit, &h);
            j->outID[i] = id;
            if (id != -1) {
                if (j->outT) j->outT[i] = h.t;
                if (j->outPos) { j->outPos[i * 3] = h.pos.x; j->outPos[i * 3 + 1] = h.pos.y; j->outPos[i * 3 + 2] = h.pos.z; }
                if (j->outNormal) { j->outNormal[i * 3] = h.normal.x; j->outNormal[i * 3 + 1] = h.normal.y; j->outNormal[i * 3 + 2] = h.normal.z; }
            }
        }
    }
    return NULL;
}

void orc_hit_brute(const float* tris9, int nTris, const float* rays6, long nRays, float tMin, float tMax,
                   int anyHit, int* outID, float* outT, float* outPos3, float* outNormal3, int threads) {
    hit_job j = {tris9, nTris, rays6, nRays, tMin, tMax, anyHit, outID, outT, outPos3, outNormal3, 0, PTHREAD_MUTEX_INITIALIZER};
    run_threads(hit_worker, &j, threads);
}
