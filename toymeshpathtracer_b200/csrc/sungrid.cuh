// sungrid.cuh -- the shadow rays' own acceleration structure: a 2-D grid in the sun's projection.
//
// Every shadow ray of the integrator has the same direction, kLightDir (main.cpp:36, 59): seen along that direction a ray is a
// POINT, and the triangles that can occlude it are the ones whose projection covers the point.  So the triangles are projected onto
// the plane perpendicular to the sun, their footprints binned into an n x n grid, and each cell's list is kept sorted by how far
// towards the sun the triangle reaches inside that cell.  A shadow query walks ONE list from the sun side down to the depth of its
// origin and runs the exact Moller-Trumbore test (bvh::mt_exact, the reference's arithmetic) on what it meets: on the headline scene
// 2.5 exact tests per shadow ray and no tree walk, against 10.4 node steps + 2.5 tests through the BVH (tools/exp_sun_grid.py).
//
// The grid only CULLS.  A query's answer is "does any triangle pass the exact test", which does not depend on order or on how many
// triangles are tested, so it equals the reference's HitScene(shadowRay) != miss as long as no triangle the exact test would accept
// is missing from the list that is walked.  That is what the padding guarantees (below); tests/ compare it with the tree and with
// the all-triangle scan ray by ray.
//
//   basis      : (U, V, L) with L = the light direction as the integrator uses it (the same floats), U, V unit vectors perpendicular to it
//   cell (i,j) : [lo + i * cell, lo + (i+1) * cell) in (u, v)
//   cellStart  : n*n + 1 offsets
//   entries    : (triangle slot, far depth of the triangle over this cell as float bits), sorted by depth descending, then slot
//
// Conservative by construction:
//   - a triangle is listed in every cell its projection touches after the cell has been grown by `pad` on every side.  pad =
//     2^-14 of the projected extent: four orders of magnitude above the rounding of the projections (a few ulp of the coordinates),
//     above the drift of a ray's projection along its length (U.L, V.L are ~1e-8, not 0) and above the slack with which the exact
//     test accepts points outside a triangle;
//   - the far depth of an entry is the triangle's plane evaluated at the corners of the grown cell, clamped to the triangle's own
//     depth range, plus zpad and 1 % of the plane's variation over the cell; a triangle seen (nearly) edge-on keeps its own maximum.
//     A hit lies at depth(origin) + t |L|^2 with t >= tMin > 0, so an entry whose far depth is below the origin's depth cannot hit.
#pragma once
#include <math.h>

#include "exact.cuh"

namespace sun {

struct View {
    const uint32_t* cellStart;  // n*n + 1
    const uint2* entries;
    float ux, uy, uz, vx, vy, vz, lx, ly, lz;
    float loU, loV, cell, invCell;
    float pad, zpad;
    float extent;  // projected extent the grid spans (before padding)
    int n;  // cells per side; 0 = no grid (shadow rays walk the BVH)
};

TMPT_HD float proj(float ax, float ay, float az, ex::V3 p) { return fmaf(p.x, ax, fmaf(p.y, ay, p.z * az)); }
TMPT_HD float proj_u(const View& g, ex::V3 p) { return proj(g.ux, g.uy, g.uz, p); }
TMPT_HD float proj_v(const View& g, ex::V3 p) { return proj(g.vx, g.vy, g.vz, p); }
TMPT_HD float proj_w(const View& g, ex::V3 p) { return proj(g.lx, g.ly, g.lz, p); }

// cell coordinate of a projected coordinate, clamped into the grid (NaN -> 0)
TMPT_HD int cell_of(float x, float lo, float invCell, int n) {
    const float f = fminf(fmaxf((x - lo) * invCell, 0.0f), (float)(n - 1));
    return (int)f;
}

// A triangle in projected coordinates.
struct Tri2 {
    float u[3], v[3], w[3];
};
TMPT_HD Tri2 project_tri(const View& g, const float* t9) {
    Tri2 t;
    for (int k = 0; k < 3; ++k) {
        const ex::V3 p = ex::v3(t9[3 * k], t9[3 * k + 1], t9[3 * k + 2]);
        t.u[k] = proj_u(g, p); t.v[k] = proj_v(g, p); t.w[k] = proj_w(g, p);
    }
    return t;
}
TMPT_HD float min3(const float* a) { return fminf(a[0], fminf(a[1], a[2])); }
TMPT_HD float max3(const float* a) { return fmaxf(a[0], fmaxf(a[1], a[2])); }

// cells the padded bounding box of the projection touches
TMPT_HD void cell_range(const View& g, const Tri2& t, int& x0, int& x1, int& y0, int& y1) {
    x0 = cell_of(min3(t.u) - g.pad, g.loU, g.invCell, g.n); x1 = cell_of(max3(t.u) + g.pad, g.loU, g.invCell, g.n);
    y0 = cell_of(min3(t.v) - g.pad, g.loV, g.invCell, g.n); y1 = cell_of(max3(t.v) + g.pad, g.loV, g.invCell, g.n);
}

// Does the projected triangle touch cell (cx, cy) grown by pad?  Separating axes: the square's own axes are the bounding-box test
// (cell_range); the triangle's three edge normals are tested here.  A degenerate projection (a segment or a point) has no interior
// side to orient a normal by: it is kept wherever its bounding box reaches.  (For a sliver the SIGN of `side` can be rounding noise
// -- only for an edge whose line passes within ~1e-7 of the triangle's size of the opposite vertex, and then every point of the
// triangle lies that close to the line: a grown cell that touches the triangle straddles it, and passes whichever way the normal
// points.  tools/fuzz_emu.py, kinds "slivers" / "edge-on" / "grazing-slivers".)
TMPT_HD bool touches_cell(const View& g, const Tri2& t, int cx, int cy) {
    const float c0x = g.loU + (float)cx * g.cell - g.pad, c0y = g.loV + (float)cy * g.cell - g.pad;
    const float c1x = g.loU + (float)(cx + 1) * g.cell + g.pad, c1y = g.loV + (float)(cy + 1) * g.cell + g.pad;
    for (int k = 0; k < 3; ++k) {
        const int b = (k + 1) % 3, o = (k + 2) % 3;
        float nx = -(t.v[b] - t.v[k]), ny = t.u[b] - t.u[k];
        const float side = nx * (t.u[o] - t.u[k]) + ny * (t.v[o] - t.v[k]);
        if (side == 0.0f) return true;  // no area: bounding box only
        if (side < 0.0f) { nx = -nx; ny = -ny; }
        const float px = nx >= 0.0f ? c1x : c0x, py = ny >= 0.0f ? c1y : c0y;  // the square's corner furthest inside
        if (nx * (px - t.u[k]) + ny * (py - t.v[k]) < 0.0f) return false;
    }
    return true;
}

// Upper bound of the triangle's depth (towards the sun) over cell (cx, cy).
TMPT_HD float far_depth(const View& g, const Tri2& t, int cx, int cy) {
    const float zmin = min3(t.w), zmax = max3(t.w);
    const float du1 = t.u[1] - t.u[0], dv1 = t.v[1] - t.v[0], dw1 = t.w[1] - t.w[0];
    const float du2 = t.u[2] - t.u[0], dv2 = t.v[2] - t.v[0], dw2 = t.w[2] - t.w[0];
    const float det = du1 * dv2 - du2 * dv1;
    const float eu = max3(t.u) - min3(t.u), ev = max3(t.v) - min3(t.v);
    float z = zmax;
    if (fabsf(det) > 9.765625e-4f * eu * ev) {  // 2^-10 of the bounding box: the plane's slopes are good to 2^-13
        const float a = (dw1 * dv2 - dw2 * dv1) / det, b = (du1 * dw2 - du2 * dw1) / det;
        const float c0x = g.loU + (float)cx * g.cell - g.pad - t.u[0], c0y = g.loV + (float)cy * g.cell - g.pad - t.v[0];
        const float c1x = g.loU + (float)(cx + 1) * g.cell + g.pad - t.u[0], c1y = g.loV + (float)(cy + 1) * g.cell + g.pad - t.v[0];
        const float zc = t.w[0] + fmaxf(a * c0x, a * c1x) + fmaxf(b * c0y, b * c1y);
        const float slack = 0.01f * (fabsf(a) + fabsf(b)) * (g.cell + 2.0f * g.pad);
        z = fminf(zmax, fmaxf(zc + slack, zmin));
    }
    return z + g.zpad;
}

// Grid resolution for a scene: about 32 cells per triangle, a power of two in [32, 4096].  Headline scene (66 k triangles), Mrays/s
// of the whole frame at 512 / 1024 / 2048 / 4096 cells per side: 7110 / 7575 / 7790 / 7875 (tree: 5255); 2048 = 17 MB of offsets + 30 MB of lists.
TMPT_HD int default_cells_per_side(int triCount) {
    int n = 32;
    while (n < 4096 && 2ll * n * n < 32ll * triCount) n *= 2;  // (nearest power of two, geometrically)
    return n;
}

// Lists longer than this are not sorted (one thread sorts a list by insertion): their entries get the far depth "infinitely far"
// instead, so a query tests all of them -- what a scan would do -- and the early exit never fires on an unsorted list.
constexpr uint32_t kSortMax = 1024;
constexpr float kFarthest = 3.0e38f;

// entry order inside a cell: far depth descending, then slot ascending (a total order: the build is deterministic)
TMPT_HD bool entry_before(uint2 a, uint2 b) {
    const float za = ex::u2f(a.y), zb = ex::u2f(b.y);
    return za > zb || (za == zb && a.x < b.x);
}

// ---- host side of the build (the product's tmpt_scene_create and the tests' host emulation share it) -------------------------
// Basis, projected bounds and pads of a scene with bounding box [bmin, bmax] for light direction l (the eight corners of the box
// bound every vertex's projection).  false: the box is not finite (no grid).
inline bool setup_view(const float bmin[3], const float bmax[3], ex::V3 l, View& g) {
    g = View{};
    const double L[3] = {l.x, l.y, l.z};
    const bool xAxis = (L[0] < 0 ? -L[0] : L[0]) < 0.9;
    const double ax[3] = {xAxis ? 1.0 : 0.0, xAxis ? 0.0 : 1.0, 0.0};
    double U[3] = {L[1] * ax[2] - L[2] * ax[1], L[2] * ax[0] - L[0] * ax[2], L[0] * ax[1] - L[1] * ax[0]};
    const double ul = sqrt(U[0] * U[0] + U[1] * U[1] + U[2] * U[2]), ll = sqrt(L[0] * L[0] + L[1] * L[1] + L[2] * L[2]);
    for (int k = 0; k < 3; ++k) U[k] /= ul;
    const double V[3] = {(L[1] * U[2] - L[2] * U[1]) / ll, (L[2] * U[0] - L[0] * U[2]) / ll, (L[0] * U[1] - L[1] * U[0]) / ll};
    g.ux = (float)U[0]; g.uy = (float)U[1]; g.uz = (float)U[2];
    g.vx = (float)V[0]; g.vy = (float)V[1]; g.vz = (float)V[2];
    g.lx = l.x; g.ly = l.y; g.lz = l.z;
    float lo[3] = {3.0e38f, 3.0e38f, 3.0e38f}, hi[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    for (int corner = 0; corner < 8; ++corner) {
        const ex::V3 p = ex::v3((corner & 1) ? bmax[0] : bmin[0], (corner & 2) ? bmax[1] : bmin[1], (corner & 4) ? bmax[2] : bmin[2]);
        const float c[3] = {proj_u(g, p), proj_v(g, p), proj_w(g, p)};
        for (int k = 0; k < 3; ++k) {
            if (!(c[k] == c[k]) || fabsf(c[k]) > 1.0e30f) return false;
            lo[k] = fminf(lo[k], c[k]); hi[k] = fmaxf(hi[k], c[k]);
        }
    }
    g.extent = fmaxf(fmaxf(hi[0] - lo[0], hi[1] - lo[1]), 1.0e-20f);
    const float mag = fmaxf(fmaxf(fmaxf(fabsf(lo[0]), fabsf(hi[0])), fmaxf(fabsf(lo[1]), fabsf(hi[1]))), fmaxf(fabsf(lo[2]), fabsf(hi[2])));
    // pads: 2^-14 of the extent, and never less than 2^-16 of the largest coordinate (rounding scales with the coordinates)
    g.pad = fmaxf(g.extent * 6.103515625e-5f, mag * 1.52587890625e-5f);
    g.zpad = fmaxf((hi[2] - lo[2]) * 6.103515625e-5f, mag * 1.52587890625e-5f);
    g.loU = lo[0] - g.pad; g.loV = lo[1] - g.pad;
    return true;
}
inline void set_resolution(View& g, int cells) {
    g.n = cells;
    g.cell = (g.extent + 2.0f * g.pad) / (float)cells;
    g.invCell = 1.0f / g.cell;
}

}  // namespace sun
