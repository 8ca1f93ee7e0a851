// mc.cu -- SURVEY §8(f) N2: surface reconstruction from the path's output (MarchingCubeGrid.py:160-209,262-409):
// bin the liquids into the dense marching-cubes grid, evaluate the colour field sum_j m/rho_j W(|x - x_j|) on every
// grid node, polygonise the 0.5 iso-surface.
//
// Differences in structure, none in result:
//  * the reference's per-cell slot table grid[grid_num][maxInGrid] (atomic slot insert) becomes ONE radix sort of the
//    liquids by the 64-bit key (cell << 32 | reference index) + a cell histogram and scan.  A cell's particles are then
//    a contiguous span in reference-index order, i.e. the order a serial run of MarchingCubeGrid.py:166-179 inserts
//    them, and "the first maxInGrid of a cell survive" (:173-177) is a rank test.  Solids are never binned: they only
//    occupy slots behind the liquids of a cell and cal_surface_point skips them (:202).
//  * a node walks 81 (x, y) rows of its 9x9x9 stencil; the 9 z-cells of a row are one span of the sorted array
//    (z is the fastest grid axis, :371-372), so a row costs two scan lookups instead of nine bucket counts.
//  * marching_cube (:262-352) appends triangles through an atomic counter; here count -> exclusive scan -> emit, which
//    yields the triangles in cell order, the order of a serial run.
// Every per-term operation is issued with round-to-nearest intrinsics in the reference's evaluation order (no FMA
// contraction, IEEE sqrt / divide), and the sums run in the serial order, so surface_value and the mesh are
// bit-identical to the CPU restatement (tests/test_mc_gpu.py), not merely within tolerance.
#include "engine.cuh"
#include "mc_table.h"
#include <cub/cub.cuh>

struct McG {
    float minx, miny, minz, gridR, inv, h, h2hi, kmk, kh3, iso, mw0, mass;
    int bx, by, bz, gn, maxInGrid;
};

__constant__ signed char c_tri[256 * 16];
__constant__ unsigned char c_nvert[256];
static bool g_tables_ready[64];

static int mc_upload_tables() {
    int dev = 0;
    CUDA_TRY(cudaGetDevice(&dev));
    if (dev < 64 && g_tables_ready[dev]) return 0;
    signed char tri[256 * 16]; unsigned char nv[256];
    for (int c = 0; c < 256; c++) {
        int n = 0;
        for (int k = 0; k < 16; k++) {
            const char ch = MC_TRITABLE_HEX[16 * c + k];
            const int v = ch == 'f' ? -1 : (ch <= '9' ? ch - '0' : ch - 'a' + 10);
            tri[16 * c + k] = (signed char)v;
            if (v >= 0 && n == k) n = k + 1;
        }
        nv[c] = (unsigned char)n;
    }
    CUDA_TRY(cudaMemcpyToSymbol(c_tri, tri, sizeof(tri)));
    CUDA_TRY(cudaMemcpyToSymbol(c_nvert, nv, sizeof(nv)));
    if (dev < 64) g_tables_ready[dev] = true;
    return 0;
}

static McG mc_consts(const wcsph_mc_grid* m) {
    McG g;
    g.minx = m->min_boundary[0]; g.miny = m->min_boundary[1]; g.minz = m->min_boundary[2];
    g.gridR = (float)m->gridR; g.inv = (float)(1.0 / m->gridR);                     // MarchingCubeGrid.py:22-23
    const double sr = m->gridR * 4.0;                                               // :25
    g.h = (float)sr;
    g.h2hi = (float)(sr * sr * 1.001);              // beyond this |r|^2 the kernel is exactly 0 whatever the rounding of sqrt and divide
    g.kmk = (float)(8.0 / 3.14159265358979323846);                                  // CubicKernel.py:14-15
    g.kh3 = (float)(1.0 / (sr * sr * sr));
    g.iso = m->isolevel;
    g.mass = m->liqiudMass;
    // liqiudMass * Cubic_W_norm(0.0) (:203), evaluated like the kernel does: (P(0) * m_k) * h3
    volatile float w0 = 1.0f * g.kmk; w0 = w0 * g.kh3;
    volatile float mw0 = g.mass * w0;
    g.mw0 = mw0;
    g.bx = m->block[0]; g.by = m->block[1]; g.bz = m->block[2];
    g.gn = g.bx * g.by * g.bz; g.maxInGrid = m->max_in_grid;
    return g;
}

// workspace carve-up (all 256-byte aligned)
struct McWork {
    unsigned long long *keys, *keys_sorted;
    int *vals, *vals_sorted;
    float4* mcpos;          // (x, y, z, m/rho or 0) in sorted order
    float4* mcblend;        // anisotropic branch: 0.05 x + 0.95 pos_avr (MarchingCubeGrid.py:229), same order
    float4* mcG;            // anisotropic branch: the three rows of G_j, same order
    int *cs;                // [gn + 1] cell histogram -> exclusive scan
    int *voff;              // [gn + 1] vertices per cell -> exclusive scan
    unsigned char *rowmask; // [gn] 1 if any binned liquid sits in cells (x, y, z-4..z+4)
    unsigned char *slabmask;// [gn] 1 if any sits in (x, y-4..y+4, z-4..z+4)
    void* cub; size_t cub_bytes;
    size_t total;
};
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static McWork mc_carve(char* base, long long gn, int nl) {
    McWork w; size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += al256(bytes); return p; };
    const size_t n = nl > 0 ? nl : 1;
    w.keys = (unsigned long long*)take(n * 8); w.keys_sorted = (unsigned long long*)take(n * 8);
    w.vals = (int*)take(n * 4); w.vals_sorted = (int*)take(n * 4);
    w.mcpos = (float4*)take(n * 16);
    w.mcblend = (float4*)take(n * 16); w.mcG = (float4*)take(n * 48);
    w.cs = (int*)take((size_t)(gn + 1) * 4); w.voff = (int*)take((size_t)(gn + 1) * 4);
    w.rowmask = (unsigned char*)take((size_t)gn); w.slabmask = (unsigned char*)take((size_t)gn);
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (unsigned long long*)nullptr, (unsigned long long*)nullptr, (int*)nullptr, (int*)nullptr, (int)n, 0, 64);
    cub::DeviceScan::ExclusiveSum(nullptr, t2, (int*)nullptr, (int*)nullptr, (int)(gn + 1));
    w.cub_bytes = al256(t1 > t2 ? t1 : t2);
    w.cub = take(w.cub_bytes);
    w.total = off;
    return w;
}

static int mc_check(wcsph_ctx* c, const wcsph_mc_grid* m, const void* work, size_t work_bytes, const char* fn, McWork* w) {
    if (!c || !m || !work) { wcsph_set_error("%s: null argument", fn); return WCSPH_EINVAL; }
    // z-slab ranks: update_grid / cal_surface_point work on the rank's OWNED liquids and yield its share of the colour field (a cell's
    // particles all live on the rank that owns its layer, so the first-maxInGrid rule stays local); the caller sums the shares over
    // the ranks (MCGrid.cal_surface_point: all_reduce) and runs marching_cube on the total.  The anisotropic pass needs single-GPU state.
    if (c->R > 1 && strstr(fn, "anistropic")) { wcsph_set_error("%s: the anisotropic branch runs on a single-GPU context", fn); return WCSPH_EINVAL; }
    if (!c->uploaded) { wcsph_set_error("%s: no positions uploaded", fn); return WCSPH_EINVAL; }
    const long long gn = (long long)m->block[0] * m->block[1] * m->block[2];
    if (m->block[0] <= 0 || m->block[1] <= 0 || m->block[2] <= 0 || gn >= (1ll << 31) - 2 || !(m->gridR > 0.0) || m->max_in_grid <= 0 || !(m->liqiudMass > 0.0f)) {
        wcsph_set_error("%s: bad grid (block %d x %d x %d, gridR %g)", fn, m->block[0], m->block[1], m->block[2], m->gridR); return WCSPH_EINVAL; }
    *w = mc_carve((char*)work, gn, c->NL);
    if (work_bytes < w->total) { wcsph_set_error("%s: workspace %zu < %zu bytes", fn, work_bytes, w->total); return WCSPH_EINVAL; }
    return 0;
}

extern "C" size_t wcsph_mc_workspace_bytes(const wcsph_mc_grid* m, int liquid_count) {
    if (!m || liquid_count < 0) return 0;
    const long long gn = (long long)m->block[0] * m->block[1] * m->block[2];
    if (gn <= 0 || gn >= (1ll << 31) - 2) return 0;
    return mc_carve(nullptr, gn, liquid_count).total;
}

// ---- update_grid (MarchingCubeGrid.py:160-179) ----
__global__ void k_mc_keys(const float4* __restrict__ pos, const int* __restrict__ sorted_id, int n, McG g,
                          unsigned long long* __restrict__ keys, int* __restrict__ vals, int* __restrict__ cnt) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float4 p = pos[k];
    // :167: cast((pos - min) * invGridR, i32) -- truncation toward zero, no FMA
    const float fx = __fmul_rn(__fsub_rn(p.x, g.minx), g.inv), fy = __fmul_rn(__fsub_rn(p.y, g.miny), g.inv), fz = __fmul_rn(__fsub_rn(p.z, g.minz), g.inv);
    int cell = g.gn;                                             // out of the box (:169): sorted behind every cell
    if (fabsf(fx) < 2.0e9f && fabsf(fy) < 2.0e9f && fabsf(fz) < 2.0e9f) {
        const int x = (int)fx, y = (int)fy, z = (int)fz;
        if (x >= 0 && x < g.bx && y >= 0 && y < g.by && z >= 0 && z < g.bz) { cell = (x * g.by + y) * g.bz + z; atomicAdd(cnt + cell, 1); }
    }
    keys[k] = ((unsigned long long)(unsigned int)cell << 32) | (unsigned int)sorted_id[k];
    vals[k] = k;
}

__global__ void k_mc_gather(const float4* __restrict__ pos, const float* __restrict__ rho, const unsigned long long* __restrict__ keys_sorted,
                            const int* __restrict__ vals_sorted, const int* __restrict__ cs, int n, McG g, float4* __restrict__ mcpos, Scalars* sc) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int slot = vals_sorted[k];
    const int cell = (int)(keys_sorted[k] >> 32);
    float4 p = pos[slot];
    const float r = rho[slot];
    // :203 low-density particles do not contribute; :173-177 only the first maxInGrid of a cell are kept
    float a = r > g.mw0 ? __fdiv_rn(g.mass, r) : 0.0f;
    if (cell < g.gn && k - cs[cell] >= g.maxInGrid) { a = 0.0f; atomicOr(&sc->flags, WCSPH_FLAG_MC_OVERFLOW); }
    if (cell >= g.gn) a = 0.0f;
    p.w = a;
    mcpos[k] = p;
}

// occupancy masks that let cal_surface_point skip empty parts of the 9x9x9 stencil: most nodes of the dense grid are
// far from any liquid, and inside the liquid ~4 of 5 (x, y) rows hold no particle (cell 0.0225 vs spacing 0.05)
__global__ void k_mc_rowmask(const int* __restrict__ cs, McG g, unsigned char* __restrict__ rowmask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.gn) return;
    const int cz = i % g.bz, base = i - cz;
    rowmask[i] = cs[base + min(cz + 4, g.bz - 1) + 1] > cs[base + max(cz - 4, 0)];
}
__global__ void k_mc_slabmask(const unsigned char* __restrict__ rowmask, McG g, unsigned char* __restrict__ slabmask) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.gn) return;
    const int cy = (i % (g.by * g.bz)) / g.bz;
    unsigned int any = 0;
    for (int y = max(cy - 4, 0); y <= min(cy + 4, g.by - 1); y++) any |= rowmask[i + (y - cy) * g.bz];
    slabmask[i] = any != 0;
}

extern "C" int wcsph_mc_update_grid(wcsph_ctx* c, const wcsph_mc_grid* m, void* work_dev, size_t work_bytes) {
    McWork w; TRY(mc_check(c, m, work_dev, work_bytes, __func__, &w));
    const McG g = mc_consts(m);
    cudaStream_t st = c->stream;
    const int n = c->nown;
    CUDA_TRY(cudaMemsetAsync(w.cs, 0, (size_t)(g.gn + 1) * sizeof(int), st));
    prof_begin(c, "k_mc_keys");
    k_mc_keys<<<nblocks(n), WCSPH_BLOCK, 0, st>>>(fown<float4>(c, "pos"), c->sorted_id[c->cur] + c->i0, n, g, w.keys, w.vals, w.cs);
    prof_end(c); LAUNCH_CHECK(c);
    int bits = 1; while ((1ll << bits) <= (long long)g.gn) bits++;
    size_t tb = w.cub_bytes;
    prof_begin(c, "mc_radix_sort");
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(w.cub, tb, w.keys, w.keys_sorted, w.vals, w.vals_sorted, n, 0, 32 + bits, st));
    prof_end(c); c->launches++;
    tb = w.cub_bytes;
    prof_begin(c, "mc_exclusive_scan");
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(w.cub, tb, w.cs, w.cs, g.gn + 1, st));
    prof_end(c); c->launches++;
    prof_begin(c, "k_mc_gather");
    k_mc_gather<<<nblocks(n), WCSPH_BLOCK, 0, st>>>(fown<float4>(c, "pos"), fown<float>(c, "rho"), w.keys_sorted, w.vals_sorted, w.cs, n, g, w.mcpos, c->sc);
    prof_end(c); LAUNCH_CHECK(c);
    prof_begin(c, "k_mc_rowmask"); k_mc_rowmask<<<nblocks(g.gn), WCSPH_BLOCK, 0, st>>>(w.cs, g, w.rowmask); prof_end(c); LAUNCH_CHECK(c);
    prof_begin(c, "k_mc_slabmask"); k_mc_slabmask<<<nblocks(g.gn), WCSPH_BLOCK, 0, st>>>(w.rowmask, g, w.slabmask); prof_end(c); LAUNCH_CHECK(c);
    return 0;
}

// ---- cal_surface_point (MarchingCubeGrid.py:183-209) ----
__device__ __forceinline__ float mc_W(const McG& g, float dx, float dy, float dz) {
    // Cubic_W(r) = Cubic_W_P(|r| / h) * m_k * h3 (CubicKernel.py:36-54), reference evaluation order
    const float r2 = __fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz));
    const float q = __fdiv_rn(__fsqrt_rn(r2), g.h);
    float res = 0.0f;
    if (q <= 1.0f) {
        if (q <= 0.5f) {
            const float qq = __fmul_rn(q, q), qqq = __fmul_rn(qq, q);
            res = __fadd_rn(__fsub_rn(__fmul_rn(6.0f, qqq), __fmul_rn(6.0f, qq)), 1.0f);
        } else {
            const float f = __fsub_rn(1.0f, q);
            res = __fmul_rn(__fmul_rn(__fmul_rn(2.0f, f), f), f);
        }
    }
    return __fmul_rn(__fmul_rn(res, g.kmk), g.kh3);
}

__global__ void __launch_bounds__(WCSPH_BLOCK)
k_mc_surface(const float4* __restrict__ mcpos, const int* __restrict__ cs, const unsigned char* __restrict__ rowmask,
             const unsigned char* __restrict__ slabmask, McG g, float* __restrict__ surface_value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.gn) return;
    const int yz = g.by * g.bz;
    const int cx = i / yz, cy = (i % yz) / g.bz, cz = i % g.bz;
    const float px = __fadd_rn(g.minx, __fmul_rn((float)cx, g.gridR)), py = __fadd_rn(g.miny, __fmul_rn((float)cy, g.gridR)),
                pz = __fadd_rn(g.minz, __fmul_rn((float)cz, g.gridR));
    const int z0 = max(cz - 4, 0), z1 = min(cz + 4, g.bz - 1);
    const int x0 = max(cx - 4, 0), x1 = min(cx + 4, g.bx - 1), y0 = max(cy - 4, 0), y1 = min(cy + 4, g.by - 1);
    float acc = 0.0f;
    for (int x = x0; x <= x1; x++) {
        if (!slabmask[(x * g.by + cy) * g.bz + cz]) continue;
        for (int y = y0; y <= y1; y++) {
            const int base = (x * g.by + y) * g.bz;
            if (!rowmask[base + cz]) continue;
            const int s = cs[base + z0], e = cs[base + z1 + 1];
            for (int k = s; k < e; k++) {
                const float4 pj = mcpos[k];
                if (pj.w == 0.0f) continue;
                const float dx = __fsub_rn(px, pj.x), dy = __fsub_rn(py, pj.y), dz = __fsub_rn(pz, pj.z);
                if (dx * dx + dy * dy + dz * dz > g.h2hi) continue;      // half of the 9^3 cube lies outside the support sphere
                const float W = mc_W(g, dx, dy, dz);
                if (W > 0.0f) acc = __fadd_rn(acc, __fmul_rn(pj.w, W));
            }
        }
    }
    surface_value[i] = acc;
}

extern "C" int wcsph_mc_cal_surface_point(wcsph_ctx* c, const wcsph_mc_grid* m, void* work_dev, size_t work_bytes, float* surface_value_dev) {
    McWork w; TRY(mc_check(c, m, work_dev, work_bytes, __func__, &w));
    if (!surface_value_dev) { wcsph_set_error("%s: null surface_value", __func__); return WCSPH_EINVAL; }
    const McG g = mc_consts(m);
    prof_begin(c, "k_mc_surface");
    k_mc_surface<<<nblocks(g.gn), WCSPH_BLOCK, 0, c->stream>>>(w.mcpos, w.cs, w.rowmask, w.slabmask, g, surface_value_dev);
    prof_end(c); LAUNCH_CHECK(c);
    return 0;
}

// ---- cal_surface_point_anistropic (MarchingCubeGrid.py:215-246) ----
// kernel centre 0.05 x_j + 0.95 pos_avr_j, distance G_j r * 2 (Yu & Turk 2013).  The candidates are the same cells as in the
// isotropic pass (the particle is binned by x_j, :167), but the support is no sphere in r, so no early-out on |r|.
__global__ void k_mc_aniso_gather(const float4* __restrict__ pos, const float4* __restrict__ pos_avr, const float4* __restrict__ G,
                                  const int* __restrict__ vals_sorted, int n, float4* __restrict__ mcblend, float4* __restrict__ mcG) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const int slot = vals_sorted[k];
    const float4 p = pos[slot], a = pos_avr[slot];
    mcblend[k] = make_float4(__fadd_rn(__fmul_rn(0.05f, p.x), __fmul_rn(0.95f, a.x)), __fadd_rn(__fmul_rn(0.05f, p.y), __fmul_rn(0.95f, a.y)),
                             __fadd_rn(__fmul_rn(0.05f, p.z), __fmul_rn(0.95f, a.z)), 0.f);
    mcG[3 * (size_t)k] = G[3 * (size_t)slot]; mcG[3 * (size_t)k + 1] = G[3 * (size_t)slot + 1]; mcG[3 * (size_t)k + 2] = G[3 * (size_t)slot + 2];
}

__global__ void __launch_bounds__(WCSPH_BLOCK)
k_mc_surface_aniso(const float4* __restrict__ mcpos, const float4* __restrict__ mcblend, const float4* __restrict__ mcG,
                   const int* __restrict__ cs, const unsigned char* __restrict__ rowmask, const unsigned char* __restrict__ slabmask,
                   McG g, float* __restrict__ surface_value) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.gn) return;
    const int yz = g.by * g.bz;
    const int cx = i / yz, cy = (i % yz) / g.bz, cz = i % g.bz;
    const float px = __fadd_rn(g.minx, __fmul_rn((float)cx, g.gridR)), py = __fadd_rn(g.miny, __fmul_rn((float)cy, g.gridR)),
                pz = __fadd_rn(g.minz, __fmul_rn((float)cz, g.gridR));
    const int z0 = max(cz - 4, 0), z1 = min(cz + 4, g.bz - 1);
    const int x0 = max(cx - 4, 0), x1 = min(cx + 4, g.bx - 1), y0 = max(cy - 4, 0), y1 = min(cy + 4, g.by - 1);
    float acc = 0.0f;
    for (int x = x0; x <= x1; x++) {
        if (!slabmask[(x * g.by + cy) * g.bz + cz]) continue;
        for (int y = y0; y <= y1; y++) {
            const int base = (x * g.by + y) * g.bz;
            if (!rowmask[base + cz]) continue;
            const int s = cs[base + z0], e = cs[base + z1 + 1];
            for (int k = s; k < e; k++) {
                const float a = mcpos[k].w;
                if (a == 0.0f) continue;
                const float4 pj = mcblend[k];
                const float rx = __fsub_rn(px, pj.x), ry = __fsub_rn(py, pj.y), rz = __fsub_rn(pz, pj.z);
                const float4 g0 = mcG[3 * (size_t)k], g1 = mcG[3 * (size_t)k + 1], g2 = mcG[3 * (size_t)k + 2];
                // Gr = G @ r * 2.0 (:234): matmul accumulates left to right
                const float ux = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(g0.x, rx), __fmul_rn(g0.y, ry)), __fmul_rn(g0.z, rz)), 2.0f);
                const float uy = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(g1.x, rx), __fmul_rn(g1.y, ry)), __fmul_rn(g1.z, rz)), 2.0f);
                const float uz = __fmul_rn(__fadd_rn(__fadd_rn(__fmul_rn(g2.x, rx), __fmul_rn(g2.y, ry)), __fmul_rn(g2.z, rz)), 2.0f);
                const float W = mc_W(g, ux, uy, uz);
                if (W > 0.0f) acc = __fadd_rn(acc, __fmul_rn(a, W));
            }
        }
    }
    surface_value[i] = acc;
}

// pos_avr4 / G12: the slot-ordered device buffers wcsph_pd_cal_anistropic_kernel filled
extern "C" int wcsph_mc_cal_surface_point_anistropic(wcsph_ctx* c, const wcsph_mc_grid* m, void* work_dev, size_t work_bytes,
                                                      const float* pos_avr4_dev, const float* G12_dev, float* surface_value_dev) {
    McWork w; TRY(mc_check(c, m, work_dev, work_bytes, __func__, &w));
    if (!surface_value_dev || !pos_avr4_dev || !G12_dev) { wcsph_set_error("%s: null buffer", __func__); return WCSPH_EINVAL; }
    const McG g = mc_consts(m);
    const int n = c->nown;
    prof_begin(c, "k_mc_aniso_gather");
    k_mc_aniso_gather<<<nblocks(n), WCSPH_BLOCK, 0, c->stream>>>(fown<float4>(c, "pos"), (const float4*)pos_avr4_dev, (const float4*)G12_dev,
                                                                w.vals_sorted, n, w.mcblend, w.mcG);
    prof_end(c); LAUNCH_CHECK(c);
    prof_begin(c, "k_mc_surface_aniso");
    k_mc_surface_aniso<<<nblocks(g.gn), WCSPH_BLOCK, 0, c->stream>>>(w.mcpos, w.mcblend, w.mcG, w.cs, w.rowmask, w.slabmask, g, surface_value_dev);
    prof_end(c); LAUNCH_CHECK(c);
    return 0;
}

// ---- marching_cube (MarchingCubeGrid.py:262-352) ----
__device__ __forceinline__ int mc_cube_index(const float* __restrict__ sv, const McG& g, int i, int cx, int cy, int cz, float* val) {
    const int yz = g.by * g.bz;
    // corners 0..7: (0,0,0) (1,0,0) (1,1,0) (0,1,0) (0,0,1) (1,0,1) (1,1,1) (0,1,1)   (:271-278)
    val[0] = sv[i]; val[1] = sv[i + yz]; val[2] = sv[i + yz + g.bz]; val[3] = sv[i + g.bz];
    val[4] = sv[i + 1]; val[5] = sv[i + yz + 1]; val[6] = sv[i + yz + g.bz + 1]; val[7] = sv[i + g.bz + 1];
    int cube = 0;
#pragma unroll
    for (int k = 0; k < 8; k++) if (val[k] < g.iso) cube |= 1 << k;
    return cube;
}

__global__ void k_mc_count(const float* __restrict__ sv, McG g, int* __restrict__ nvert) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i > g.gn) return;
    int nv = 0;
    if (i < g.gn) {
        const int yz = g.by * g.bz;
        const int cx = i / yz, cy = (i % yz) / g.bz, cz = i % g.bz;
        if (cx + 1 < g.bx && cy + 1 < g.by && cz + 1 < g.bz) {                      // :268
            float val[8];
            nv = c_nvert[mc_cube_index(sv, g, i, cx, cy, cz, val)];
        }
    }
    nvert[i] = nv;
}

// :392-409 -- the later axis overrides the earlier one, equal points give 1
__device__ __forceinline__ int mc_check_pos(const float* p2, const float* p1) {
    int ret = 1;
    if (p2[0] < p1[0]) ret = 1; else if (p2[0] > p1[0]) ret = 0;
    if (p2[1] < p1[1]) ret = 1; else if (p2[1] > p1[1]) ret = 0;
    if (p2[2] < p1[2]) ret = 1; else if (p2[2] > p1[2]) ret = 0;
    return ret;
}

__global__ void k_mc_emit(const float* __restrict__ sv, const int* __restrict__ voff, McG g, float* __restrict__ triangle, int max_vertex) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= g.gn) return;
    const int off = voff[i], nv = voff[i + 1] - off;
    if (nv == 0) return;
    const int yz = g.by * g.bz;
    const int cx = i / yz, cy = (i % yz) / g.bz, cz = i % g.bz;
    float val[8];
    const int cube = mc_cube_index(sv, g, i, cx, cy, cz, val);
    for (int k = 0; k < nv; k += 3) {
        if (off + k >= max_vertex) return;                                           // :344 `old < MAX_VERTEX`, per triangle
        for (int t = 0; t < 3; t++) {
            const int e = c_tri[16 * cube + k + t];
            // edge e joins corners (a, b): 0-1 1-2 2-3 3-0 4-5 5-6 6-7 7-4 0-4 1-5 2-6 3-7   (:307-330)
            const int a = e < 8 ? e : e - 8, b = e < 8 ? ((e & 3) == 3 ? e - 3 : e + 1) : e - 4;
            float p1[3], p2[3];
            const int ax = ((a & 3) == 1 || (a & 3) == 2), ay = (a & 3) >= 2, az = a >> 2;
            const int bx_ = ((b & 3) == 1 || (b & 3) == 2), by_ = (b & 3) >= 2, bz_ = b >> 2;
            p1[0] = __fadd_rn(g.minx, __fmul_rn((float)(cx + ax), g.gridR)); p1[1] = __fadd_rn(g.miny, __fmul_rn((float)(cy + ay), g.gridR)); p1[2] = __fadd_rn(g.minz, __fmul_rn((float)(cz + az), g.gridR));
            p2[0] = __fadd_rn(g.minx, __fmul_rn((float)(cx + bx_), g.gridR)); p2[1] = __fadd_rn(g.miny, __fmul_rn((float)(cy + by_), g.gridR)); p2[2] = __fadd_rn(g.minz, __fmul_rn((float)(cz + bz_), g.gridR));
            float v1 = val[a], v2 = val[b];
            // vertex_interp (:375-389)
            const float *q1 = p1, *q2 = p2;
            if (mc_check_pos(q2, q1) == 1) { const float* tp = q1; q1 = q2; q2 = tp; const float tv = v1; v1 = v2; v2 = tv; }
            float out[3] = {q1[0], q1[1], q1[2]};
            if (fabsf(__fsub_rn(v1, v2)) > 0.00001f) {
                const float den = __fsub_rn(v2, v1), lev = __fsub_rn(g.iso, v1);
#pragma unroll
                for (int d = 0; d < 3; d++) out[d] = __fadd_rn(q1[d], __fmul_rn(__fdiv_rn(__fsub_rn(q2[d], q1[d]), den), lev));
            }
            float* dst = triangle + 3 * (size_t)(off + k + t);
            dst[0] = out[0]; dst[1] = out[1]; dst[2] = out[2];
        }
    }
}

extern "C" int wcsph_mc_marching_cube(wcsph_ctx* c, const wcsph_mc_grid* m, void* work_dev, size_t work_bytes, const float* surface_value_dev,
                                      float* triangle_dev, int max_vertex, int* vertex_count_out) {
    McWork w; TRY(mc_check(c, m, work_dev, work_bytes, __func__, &w));
    if (!surface_value_dev || !triangle_dev || !vertex_count_out || max_vertex < 0 || max_vertex % 3 != 0) {
        wcsph_set_error("%s: null buffer or max_vertex not a multiple of 3", __func__); return WCSPH_EINVAL; }
    TRY(mc_upload_tables());
    const McG g = mc_consts(m);
    cudaStream_t st = c->stream;
    prof_begin(c, "k_mc_count"); k_mc_count<<<nblocks(g.gn + 1), WCSPH_BLOCK, 0, st>>>(surface_value_dev, g, w.voff); prof_end(c); LAUNCH_CHECK(c);
    size_t tb = w.cub_bytes;
    prof_begin(c, "mc_exclusive_scan");
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(w.cub, tb, w.voff, w.voff, g.gn + 1, st));
    prof_end(c); c->launches++;
    prof_begin(c, "k_mc_emit"); k_mc_emit<<<nblocks(g.gn), WCSPH_BLOCK, 0, st>>>(surface_value_dev, w.voff, g, triangle_dev, max_vertex); prof_end(c); LAUNCH_CHECK(c);
    int total = 0;                                              // vertex_count[0]: keeps counting past max_vertex (:343-349)
    CUDA_TRY(cudaMemcpyAsync(&total, w.voff + g.gn, sizeof(int), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    *vertex_count_out = total;
    return 0;
}
