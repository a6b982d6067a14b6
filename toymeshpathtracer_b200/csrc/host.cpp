// host.cpp -- the C++ host glue that main() wraps around the hot path (no CUDA in this file):
// OBJ ingest, floor + bounds, camera, PNG output and the `<width> <height> <spp> <datafile>`
// command line.  Each function names the reference lines whose behaviour it keeps
// (paths relative to /root/reference/source).  Compiled with -ffp-contract=off: the floats
// produced here (triangle vertices, bounds, camera) must equal the reference's bit for bit.
#include <cctype>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.h"

namespace {

// ---- OBJ text -> floats / ints: grammar of external/objparser.cpp:29-155 -------------------
inline const char* skip_blank(const char* s) {
    while (*s == ' ' || *s == '\t') ++s;
    return s;
}

// objparser.cpp:34-60: optional sign, decimal digits, wraps like unsigned arithmetic
int parse_int(const char* s, const char** end) {
    s = skip_blank(s);
    const bool negative = *s == '-';
    if (*s == '-' || *s == '+') ++s;
    unsigned value = 0;
    while (*s >= '0' && *s <= '9') value = value * 10u + unsigned(*s++ - '0');
    *end = s;
    return negative ? int(0u - value) : int(value);  // (the reference writes -int(result): the same bits, without the overflow at 2^31)
}

// objparser.cpp:62-131: mantissa digits accumulated in a double, decimal exponent applied by
// ONE division or multiplication with an exact power of ten (|exp| <= 22), then rounded to float.
float parse_float(const char* s, const char** end) {
    static const double kPow10[23] = {1e0,  1e1,  1e2,  1e3,  1e4,  1e5,  1e6,  1e7,  1e8,  1e9,  1e10, 1e11,
                                      1e12, 1e13, 1e14, 1e15, 1e16, 1e17, 1e18, 1e19, 1e20, 1e21, 1e22};
    s = skip_blank(s);
    const double sign = *s == '-' ? -1.0 : 1.0;
    if (*s == '-' || *s == '+') ++s;
    double mantissa = 0.0;
    int exp10 = 0;
    while (*s >= '0' && *s <= '9') mantissa = mantissa * 10.0 + double(*s++ - '0');
    if (*s == '.') {
        ++s;
        while (*s >= '0' && *s <= '9') {
            mantissa = mantissa * 10.0 + double(*s++ - '0');
            --exp10;
        }
    }
    if (*s == 'e' || *s == 'E') {
        ++s;
        const int esign = *s == '-' ? -1 : 1;
        if (*s == '-' || *s == '+') ++s;
        unsigned e = 0;  // wraps like the reference's int does in practice (objparser.cpp:113-117), without signed overflow
        while (*s >= '0' && *s <= '9') e = e * 10u + unsigned(*s++ - '0');
        exp10 = int(unsigned(exp10) + unsigned(esign) * e);
    }
    *end = s;
    if (exp10 <= 0 && exp10 >= -22) return float(sign * mantissa / kPow10[-exp10]);
    if (exp10 > 0 && exp10 <= 22) return float(sign * mantissa * kPow10[exp10]);
    return float(sign * mantissa * std::pow(10.0, exp10));
}

// objparser.cpp:133-155: "v", "v/vt", "v//vn", "v/vt/vn" -- only the position index matters here
const char* parse_face_vertex(const char* s, int& vi) {
    vi = parse_int(s, &s);
    if (*s != '/') return s;
    ++s;
    int ignored;
    if (*s != '/') ignored = parse_int(s, &s);
    if (*s != '/') return s;
    ++s;
    ignored = parse_int(s, &s);
    (void)ignored;
    return s;
}

struct ObjMesh {
    std::vector<float> positions;  // xyz
    std::vector<int> corners;      // 3 position indices per triangle, already 0-based
};

// objparser.cpp:185-302 (only the "v " and "f " records reach LoadScene)
void parse_obj_line(ObjMesh& m, const char* line) {
    if (line[0] == 'v' && line[1] == ' ') {
        const char* s = line + 2;
        for (int k = 0; k < 3; ++k) m.positions.push_back(parse_float(s, &s));
    } else if (line[0] == 'f' && line[1] == ' ') {
        const char* s = line + 2;
        const int vcount = int(m.positions.size() / 3);
        int first = 0, prev = 0, seen = 0;
        while (*s) {
            int vi;
            s = parse_face_vertex(s, vi);
            if (vi == 0) break;                                  // objparser.cpp:256-257
            const int idx = vi > 0 ? vi - 1 : vcount + vi;       // objparser.cpp:29-32
            if (seen == 0) first = idx;
            else if (seen == 1) prev = idx;
            else {                                               // fan: (first, prev, current), objparser.cpp:263-280
                m.corners.push_back(first);
                m.corners.push_back(prev);
                m.corners.push_back(idx);
                prev = idx;
            }
            if (seen < 2) ++seen;
        }
    }
}

bool parse_obj_file(const char* path, ObjMesh& m) {
    FILE* f = std::fopen(path, "rb");
    if (!f) return false;
    std::string text;
    char buf[1 << 16];
    size_t got;
    while ((got = std::fread(buf, 1, sizeof buf, f)) > 0) text.append(buf, got);
    std::fclose(f);
    // records end at '\n' (objparser.cpp:320-335); a '\r' stays in the record and ends numbers
    size_t pos = 0;
    while (pos < text.size()) {
        size_t eol = text.find('\n', pos);
        if (eol == std::string::npos) eol = text.size();
        std::string line = text.substr(pos, eol - pos);
        const size_t nul = line.find('\0');  // the reference hands C strings to the line parser
        if (nul != std::string::npos) line.resize(nul);
        parse_obj_line(m, line.c_str());
        pos = eol + 1;
    }
    return true;
}

// glm::min / glm::max (func_common.inl:17-30)
inline float gmin(float x, float y) { return (y < x) ? y : x; }
inline float gmax(float x, float y) { return (x < y) ? y : x; }

struct V3 {
    float x, y, z;
};
inline V3 sub(V3 a, V3 b) { return V3{a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 add(V3 a, V3 b) { return V3{a.x + b.x, a.y + b.y, a.z + b.z}; }
inline V3 mulv(V3 a, V3 b) { return V3{a.x * b.x, a.y * b.y, a.z * b.z}; }
inline V3 muls(V3 a, float s) { return V3{a.x * s, a.y * s, a.z * s}; }
inline float dot(V3 a, V3 b) { const float x = a.x * b.x, y = a.y * b.y, z = a.z * b.z; return (x + y) + z; }  // func_geometric.inl:48-55
inline V3 cross(V3 a, V3 b) { return V3{a.y * b.z - b.y * a.z, a.z * b.x - b.z * a.x, a.x * b.y - b.x * a.y}; }  // :68-79
inline V3 normalize(V3 v) { const float s = 1.0f / std::sqrt(dot(v, v)); return muls(v, s); }                  // :82-90
inline void put(float* dst, V3 v) { dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; }

// ---- PNG: 8-bit RGBA, zlib stream of stored (uncompressed) deflate blocks -------------------
uint32_t crc32_update(uint32_t crc, const uint8_t* p, size_t n) {
    static uint32_t table[256];
    static bool ready = false;
    if (!ready) {
        for (uint32_t i = 0; i < 256; ++i) {
            uint32_t c = i;
            for (int k = 0; k < 8; ++k) c = (c & 1u) ? 0xEDB88320u ^ (c >> 1) : c >> 1;
            table[i] = c;
        }
        ready = true;
    }
    crc = ~crc;
    for (size_t i = 0; i < n; ++i) crc = table[(crc ^ p[i]) & 0xFFu] ^ (crc >> 8);
    return ~crc;
}
void be32(std::vector<uint8_t>& v, uint32_t x) {
    v.push_back(uint8_t(x >> 24)); v.push_back(uint8_t(x >> 16)); v.push_back(uint8_t(x >> 8)); v.push_back(uint8_t(x));
}
void png_chunk(std::vector<uint8_t>& out, const char type[4], const std::vector<uint8_t>& data) {
    be32(out, (uint32_t)data.size());
    const size_t start = out.size();
    out.insert(out.end(), type, type + 4);
    out.insert(out.end(), data.begin(), data.end());
    be32(out, crc32_update(0u, out.data() + start, out.size() - start));
}

}  // namespace

// ------------------------------------------------------------------------------------------
extern "C" void tmpt_free(void* p) { std::free(p); }

// LoadScene (main.cpp:122-170) up to the point where the Scene is constructed.
extern "C" int tmpt_load_obj(const char* path, float** outTris9, int* outTriCount, float boundsMin[3], float boundsMax[3]) {
    if (!path || !outTris9 || !outTriCount || !boundsMin || !boundsMax) return tmpt::fail(TMPT_ERR_ARG, "tmpt_load_obj: NULL argument");
    *outTris9 = nullptr;
    *outTriCount = 0;
    ObjMesh m;
    if (!parse_obj_file(path, m)) return tmpt::fail(TMPT_ERR_IO, "failed to load .obj file '%s'", path);
    const int objTris = int(m.corners.size() / 3);
    const int vcount = int(m.positions.size() / 3);
    for (int idx : m.corners)  // the reference indexes unchecked (UB on a bad file); this fails instead
        if (idx < 0 || idx >= vcount) return tmpt::fail(TMPT_ERR_IO, "'%s': face references vertex %d of %d", path, idx + 1, vcount);
    float* tris = (float*)std::malloc(sizeof(float) * 9 * (size_t)(objTris + 2));
    if (!tris) return tmpt::fail(TMPT_ERR_OOM, "tmpt_load_obj: out of memory");
    V3 mn{+1.0e6f, +1.0e6f, +1.0e6f}, mx{-1.0e6f, -1.0e6f, -1.0e6f};  // main.cpp:132-133
    for (int i = 0; i < objTris; ++i) {
        for (int c = 0; c < 3; ++c) {
            const float* p = &m.positions[(size_t)m.corners[(size_t)i * 3 + c] * 3];
            float* dst = tris + (size_t)i * 9 + c * 3;
            dst[0] = p[0]; dst[1] = p[1]; dst[2] = p[2];
            mn = V3{gmin(mn.x, p[0]), gmin(mn.y, p[1]), gmin(mn.z, p[2])};
            mx = V3{gmax(mx.x, p[0]), gmax(mx.y, p[1]), gmax(mx.z, p[2])};
        }
    }
    // the two floor triangles (main.cpp:153-162)
    const V3 extra = muls(sub(mx, mn), 0.7f);
    const float x0 = mn.x - extra.x, x1 = mx.x + extra.x, z0 = mn.z - extra.z, z1 = mx.z + extra.z, y = mn.y;
    const float floorTris[18] = {x0, y, z0, x0, y, z1, x1, y, z0, x0, y, z1, x1, y, z1, x1, y, z0};
    std::memcpy(tris + (size_t)objTris * 9, floorTris, sizeof floorTris);
    put(boundsMin, mn);
    put(boundsMax, mx);
    *outTris9 = tris;
    *outTriCount = objTris + 2;
    return TMPT_OK;
}

// Camera::Camera (maths.cpp:40-59)
extern "C" void tmpt_camera_make(const float lookFrom[3], const float lookAt[3], const float vup[3], float vfovDeg, float aspect,
                                 float aperture, float focusDist, tmpt_camera* out) {
    const float kPI = 3.1415926f;  // maths.h:14
    const V3 from{lookFrom[0], lookFrom[1], lookFrom[2]}, at{lookAt[0], lookAt[1], lookAt[2]}, up{vup[0], vup[1], vup[2]};
    out->lensRadius = aperture * 0.5f;
    const float theta = vfovDeg * kPI / 180.0f;
    const float halfHeight = std::tan(theta * 0.5f);  // tanf
    const float halfWidth = aspect * halfHeight;
    const V3 w = normalize(sub(from, at));
    const V3 u = normalize(cross(up, w));
    const V3 v = cross(w, u);
    const V3 llc = sub(sub(sub(from, muls(u, halfWidth * focusDist)), muls(v, halfHeight * focusDist)), muls(w, focusDist));
    put(out->origin, from);
    put(out->lowerLeftCorner, llc);
    put(out->horizontal, muls(u, 2.0f * halfWidth * focusDist));
    put(out->vertical, muls(v, 2.0f * halfHeight * focusDist));
    put(out->u, u);
    put(out->v, v);
    put(out->w, w);
}

// Camera placement of main() (main.cpp:296-307)
extern "C" void tmpt_camera_for_scene(const char* objPath, const float boundsMin[3], const float boundsMax[3], int width, int height,
                                      tmpt_camera* out) {
    const V3 mn{boundsMin[0], boundsMin[1], boundsMin[2]}, mx{boundsMax[0], boundsMax[1], boundsMax[2]};
    const V3 size = sub(mx, mn);
    const V3 center = muls(add(mn, mx), 0.5f);
    V3 lookfrom = add(center, mulv(size, V3{0.3f, 0.6f, 1.2f}));
    if (objPath && std::strstr(objPath, "sponza.obj")) lookfrom = V3{-5.96f, 4.08f, -1.22f};  // main.cpp:300-301
    const V3 lookat = add(center, mulv(size, V3{0.0f, -0.1f, 0.0f}));
    const V3 dv = sub(lookfrom, lookat);
    const float distToFocus = std::sqrt(dot(dv, dv));
    const float from[3] = {lookfrom.x, lookfrom.y, lookfrom.z}, at[3] = {lookat.x, lookat.y, lookat.z}, up[3] = {0.0f, 1.0f, 0.0f};
    tmpt_camera_make(from, at, up, 60.0f, float(width) / float(height), 0.03f, distToFocus, out);
}

// stbi_write_png with stbi_flip_vertically_on_write (main.cpp:341-342): same decoded pixels.
extern "C" int tmpt_write_png(const char* path, int width, int height, const uint8_t* rgba, int flipVertically) {
    if (!path || !rgba || width < 1 || height < 1) return tmpt::fail(TMPT_ERR_ARG, "tmpt_write_png: bad arguments");
    const size_t rowBytes = (size_t)width * 4;
    std::vector<uint8_t> raw;
    raw.reserve((rowBytes + 1) * (size_t)height);
    for (int y = 0; y < height; ++y) {
        const uint8_t* row = rgba + rowBytes * (size_t)(flipVertically ? height - 1 - y : y);
        raw.push_back(0);  // filter: none
        raw.insert(raw.end(), row, row + rowBytes);
    }
    std::vector<uint8_t> z;
    z.reserve(raw.size() + raw.size() / 65535 * 5 + 16);
    z.push_back(0x78); z.push_back(0x01);
    uint32_t a = 1, b = 0;  // Adler-32
    size_t pos = 0;
    do {
        const size_t n = std::min<size_t>(65535, raw.size() - pos);
        const bool last = pos + n == raw.size();
        z.push_back(last ? 1 : 0);
        z.push_back(uint8_t(n)); z.push_back(uint8_t(n >> 8)); z.push_back(uint8_t(~n)); z.push_back(uint8_t((~n) >> 8));
        z.insert(z.end(), raw.begin() + pos, raw.begin() + pos + n);
        for (size_t i = 0; i < n; ++i) {
            a += raw[pos + i]; if (a >= 65521u) a -= 65521u;
            b += a; if (b >= 65521u) b -= 65521u;
        }
        pos += n;
    } while (pos < raw.size());
    be32(z, (b << 16) | a);

    std::vector<uint8_t> out = {0x89, 'P', 'N', 'G', 0x0D, 0x0A, 0x1A, 0x0A};
    std::vector<uint8_t> ihdr;
    be32(ihdr, (uint32_t)width);
    be32(ihdr, (uint32_t)height);
    ihdr.push_back(8); ihdr.push_back(6); ihdr.push_back(0); ihdr.push_back(0); ihdr.push_back(0);  // 8-bit RGBA
    png_chunk(out, "IHDR", ihdr);
    png_chunk(out, "IDAT", z);
    png_chunk(out, "IEND", {});
    FILE* f = std::fopen(path, "wb");
    if (!f) return tmpt::fail(TMPT_ERR_IO, "cannot write '%s'", path);
    const bool ok = std::fwrite(out.data(), 1, out.size(), f) == out.size();
    std::fclose(f);
    return ok ? TMPT_OK : tmpt::fail(TMPT_ERR_IO, "short write to '%s'", path);
}

// main() (main.cpp:248-345): same argv grammar, messages, exit codes and report lines.
extern "C" int tmpt_main(int argc, const char** argv) {
    if (argc < 5) {
        std::printf("Usage: TrimeshTracer.exe [width] [height] [samplesPerPixel] [objFile]\n");
        return 1;
    }
    const int width = std::atoi(argv[1]);
    if (width < 1 || width > 10000) { std::printf("ERROR: invalid width argument '%s'\n", argv[1]); return 1; }
    const int height = std::atoi(argv[2]);
    if (height < 1 || height > 10000) { std::printf("ERROR: invalid height argument '%s'\n", argv[2]); return 1; }
    const int spp = std::atoi(argv[3]);
    if (spp < 1 || spp > 1024) { std::printf("ERROR: invalid samplesPerPixel argument '%s'\n", argv[3]); return 1; }

    float* tris = nullptr;
    int triCount = 0;
    float mn[3], mx[3];
    if (tmpt_load_obj(argv[4], &tris, &triCount, mn, mx) != TMPT_OK) {
        std::printf("ERROR: failed to load .obj file\n");
        return 1;
    }
    // TMPT_DEVICE=<first device>, TMPT_GPUS=<n>: n replicas on devices first .. first+n-1, rows dealt out in stripes
    const char* devEnv = std::getenv("TMPT_DEVICE");
    const char* gpusEnv = std::getenv("TMPT_GPUS");
    const int device = devEnv ? std::atoi(devEnv) : 0;
    const int gpus = gpusEnv ? std::atoi(gpusEnv) : 1;
    if (gpus < 1 || gpus > 64) { std::printf("ERROR: invalid TMPT_GPUS '%s'\n", gpusEnv); tmpt_free(tris); return 1; }
    const auto t0 = std::chrono::steady_clock::now();
    std::vector<tmpt_scene*> scenes((size_t)gpus, nullptr);
    auto destroy_all = [&]() { for (tmpt_scene* sc : scenes) tmpt_scene_destroy(sc); };
    for (int i = 0; i < gpus; ++i) {
        if (tmpt_scene_create(tris, triCount, device + i, TMPT_BUILD_DEFAULT, &scenes[(size_t)i]) != TMPT_OK) {
            std::printf("ERROR: %s\n", tmpt_last_error());
            tmpt_free(tris);
            destroy_all();
            return 1;
        }
    }
    tmpt_free(tris);
    const double initSec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    std::printf("Initialized scene '%s' (%i tris) in %.3fs\n", argv[4], triCount, initSec);

    tmpt_camera camera;
    tmpt_camera_for_scene(argv[4], mn, mx, width, height, &camera);
    std::vector<uint8_t> image((size_t)width * height * 4, 0);
    uint64_t rayCount = 0;
    double dt = 0.0;
    if (tmpt_render_multi(scenes.data(), gpus, &camera, width, height, spp, image.data(), &rayCount, &dt) != TMPT_OK) {
        std::printf("ERROR: %s\n", tmpt_last_error());
        destroy_all();
        return 1;
    }
    std::printf("Rendered scene at %ix%i,%ispp in %.3f s\n", width, height, spp, dt);
    std::printf("- %.1f K Rays, %.1f K Rays/s\n", rayCount / 1000.0, rayCount / 1000.0 / dt);
    destroy_all();
    if (tmpt_write_png("output.png", width, height, image.data(), 1) != TMPT_OK) {
        std::printf("ERROR: %s\n", tmpt_last_error());
        return 1;
    }
    return 0;
}
