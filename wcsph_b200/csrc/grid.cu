// grid.cu -- HashGrid.update_grid (HashGrid.py:57-106) re-designed for B200.
//
// The reference rebuilds an N x 64 bucket table and an NL x 2048 candidate table every
// step.  Here: liquids are radix-sorted by TRUE cell id (x fastest), a per-row span walk of
// the 5x5 rows of the stencil builds a compact list of IN-RANGE neighbours once per step,
// and every later sweep of the step streams that list.  The two places where the
// reference's hash table is observable are reproduced exactly:
//   (Q1) bucket aliasing inside the 125-cell stencil duplicates neighbours: a static table
//        of near-alias cell pairs drives a fix-up kernel that appends the duplicates;
//   (Q2) neighborCount counts every candidate of the 125 buckets: evaluated from the
//        bucket-occupancy table with a separable 5x5x5 box sum, minus the self visits.
#include "engine.cuh"
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>

// ---- keys ---------------------------------------------------------------------------------
// HashGrid.py:67-76 "insert pos": cell of every particle, bucket occupancy (gridCount)
// The sort key is the cell of the SEARCH grid gs (gs == g, or g refined by F = 2): a reference cell is F^3 search cells, the
// particle's half along each axis is decided on the same f32 product the reference truncates (Q19), so that `search cell / F` is
// the reference's cell for every in-box particle (also for the slightly negative coordinates that truncate into cell 0).
__device__ __forceinline__ int search_key(const GridDims& g, const GridDims& gs, int F, float4 p, int cx, int cy, int cz) {
    if (F == 1) return (cz * g.by + cy) * g.bx + cx;
    const float ax = __fmul_rn(__fsub_rn(p.x, g.minx), g.inv) - (float)cx, ay = __fmul_rn(__fsub_rn(p.y, g.miny), g.inv) - (float)cy,
                az = __fmul_rn(__fsub_rn(p.z, g.minz), g.inv) - (float)cz;
    const int sx = 2 * cx + (ax >= 0.5f), sy = 2 * cy + (ay >= 0.5f), sz = 2 * cz + (az >= 0.5f);
    return (sz * gs.by + sy) * gs.bx + sx;
}
__global__ void k_keys(const float4* __restrict__ pos, int n, GridDims g, GridDims gs, int F, int* __restrict__ keys,
                       int* __restrict__ cell_count, int* __restrict__ occ) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos[i];
    int cx, cy, cz; cell_coords(g, p.x, p.y, p.z, cx, cy, cz);
    int key = gs.ncells;
    if (in_box(g, cx, cy, cz)) {
        key = search_key(g, gs, F, p, cx, cy, cz);
        atomicAdd(&occ[cell_hash(cx, cy, cz, g.n_hash)], 1);
    }
    keys[i] = key;
    atomicAdd(&cell_count[key], 1);
}

__global__ void k_iota(int* a, int n) { int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) a[i] = i; }

// ---- static tables --------------------------------------------------------------------------
__global__ void k_bucket_of_cell(GridDims g, int* __restrict__ boc) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncells) return;
    int cx = c % g.bx, cy = (c / g.bx) % g.by, cz = c / (g.bx * g.by);
    boc[c] = cell_hash(cx, cy, cz, g.n_hash);
}

// m_self(c) = #{in-box o in [-2,2]^3 : bucket(c+o) == bucket(c)}: how often the walk of
// HashGrid.py:82-85 visits the particle's own bucket (each visit skips i, HashGrid.py:98)
__global__ void k_m_self(GridDims g, const int* __restrict__ boc, unsigned char* __restrict__ m_self) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncells) return;
    int cx = c % g.bx, cy = (c / g.bx) % g.by, cz = c / (g.bx * g.by);
    int b = boc[c], m = 0;
    for (int dz = -2; dz <= 2; dz++) for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) {
        int x = cx + dx, y = cy + dy, z = cz + dz;
        if (in_box(g, x, y, z)) m += (boc[(z * g.by + y) * g.bx + x] == b);
    }
    m_self[c] = (unsigned char)(m > 255 ? 255 : m);
}

// solids are static: a cell whose 5x5x5 block holds no solid particle never needs the solid spans
// (per REFERENCE cell; css is indexed by search cell: a reference cell spans F search cells per axis)
__global__ void k_solid_near(GridDims g, GridDims gs, int F, const int* __restrict__ css, unsigned char* __restrict__ near) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncells) return;
    int cx = c % g.bx, cy = (c / g.bx) % g.by, cz = c / (g.bx * g.by);
    int x0 = F * max(cx - 2, 0), x1 = F * (min(cx + 2, g.bx - 1) + 1) - 1;
    int n = 0;
    for (int z = F * max(cz - 2, 0); z < F * (min(cz + 2, g.bz - 1) + 1); z++)
        for (int y = F * max(cy - 2, 0); y < F * (min(cy + 2, g.by - 1) + 1); y++) {
            const int base = (z * gs.by + y) * gs.bx;
            n += css[base + x1 + 1] - css[base + x0];
        }
    near[c] = n > 0;
}

// every unordered pair of distinct in-box cells with equal bucket and Chebyshev distance <= 4
// (both can sit in one 125-cell stencil).  Expected count ~ ncells*364/N ~ 10^3.
__global__ void k_alias_pairs(GridDims g, const int* __restrict__ boc, int* __restrict__ pairs, Scalars* sc) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncells) return;
    int cx = c % g.bx, cy = (c / g.bx) % g.by, cz = c / (g.bx * g.by);
    int b = boc[c];
    for (int dz = 0; dz <= 4; dz++) for (int dy = (dz ? -4 : 0); dy <= 4; dy++)
        for (int dx = ((dz || dy) ? -4 : 1); dx <= 4; dx++) {     // forward half: each pair once
            int x = cx + dx, y = cy + dy, z = cz + dz;
            if (!in_box(g, x, y, z)) continue;
            int c2 = (z * g.by + y) * g.bx + x;
            if (boc[c2] == b) {
                int slot = atomicAdd(&sc->alias_count, 1);
                if (slot < WCSPH_ALIAS_CAP) { pairs[2 * slot] = c; pairs[2 * slot + 1] = c2; }
                else atomicOr(&sc->flags, WCSPH_FLAG_ALIAS_OVERFLOW);
            }
        }
}

// ---- permutation of the persistent fields -------------------------------------------------
struct PermuteArgs { int n4, n1; const float4* src4[8]; float4* dst4[8]; const float* src1[8]; float* dst1[8]; };
// kspan > 0: the sort ran on slab-relative keys (mgpu.cu: slab_sort_key); keys_sorted goes back to global cell ids here
__global__ void k_permute(PermuteArgs a, const int* __restrict__ perm, int n,
                          const int* __restrict__ sid_old, int* __restrict__ sid_new, int* __restrict__ keys_sorted, int kbase, int kspan, int ncells) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    if (kspan > 0) { const int ks = keys_sorted[k]; keys_sorted[k] = ks >= kspan ? ncells + (ks - kspan) : ks + kbase; }
    int o = perm[k];
    sid_new[k] = sid_old[o];
    for (int f = 0; f < a.n4; f++) a.dst4[f][k] = a.src4[f][o];
    for (int f = 0; f < a.n1; f++) a.dst1[f][k] = a.src1[f][o];
}

// ---- reference-exact neighborCount ---------------------------------------------------------
// pass X gathers occ[bucket(cell)] for the 5 x-neighbours; passes Y, Z finish the box sum
// the three passes only cover the cell range a rank needs: x, y over the slab + 2 layers each side,
// z over the slab itself (single GPU: the whole grid)
// (each cell's clipped bucket occupancy is gathered ONCE -- two dependent loads, the second one random into the hash table -- and the
// five taps come from shared memory; the first version gathered it five times per cell)
__global__ void __launch_bounds__(WCSPH_BLOCK) k_box_x(GridDims g, const int* __restrict__ boc, const int* __restrict__ occ, int max_in_grid,
                        int* __restrict__ out, Scalars* sc, int c0, int c1) {
    __shared__ int sv[WCSPH_BLOCK + 4];
    const int b0 = c0 + blockIdx.x * blockDim.x;              // first cell of this block
    const int c = b0 + threadIdx.x;
    // own cell -> sv[t + 2]; threads 0..3 also fetch the two cells left and right of the block (x taps never leave a row, rows are
    // contiguous in c, so only cells inside the grid are ever used)
    int o = 0;
    if (c < g.ncells) {
        o = occ[boc[c]];
        if (o > max_in_grid) { o = max_in_grid; if (c < c1) atomicOr(&sc->flags, WCSPH_FLAG_BUCKET_OVERFLOW); }   // Q4
    }
    sv[threadIdx.x + 2] = o;
    if (threadIdx.x < 4) {
        const int h = threadIdx.x < 2 ? b0 - 2 + threadIdx.x : b0 + blockDim.x + (threadIdx.x - 2);
        int v = 0;
        if (h >= 0 && h < g.ncells) v = min(occ[boc[h]], max_in_grid);
        sv[threadIdx.x < 2 ? threadIdx.x : blockDim.x + threadIdx.x] = v;
    }
    __syncthreads();
    if (c >= c1) return;
    const int cx = c % g.bx;
    int s = 0;
#pragma unroll
    for (int d = -2; d <= 2; d++) {
        const int x = cx + d;
        if (x >= 0 && x < g.bx) s += sv[threadIdx.x + 2 + d];
    }
    out[c] = s;
}
__global__ void k_box_y(GridDims g, const int* __restrict__ in, int* __restrict__ out, int c0, int c1) {
    int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    int cy = (c / g.bx) % g.by;
    int s = 0;
    for (int d = -2; d <= 2; d++) { int y = cy + d; if (y >= 0 && y < g.by) s += in[c + d * g.bx]; }
    out[c] = s;
}
__global__ void k_box_z(GridDims g, const int* __restrict__ in, int* __restrict__ out, int c0, int c1, int zmin, int zmax) {
    int c = c0 + blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= c1) return;
    int cz = c / (g.bx * g.by);
    int s = 0, pl = g.bx * g.by;
    // layers outside [zmin, zmax) were not produced by the x/y passes of this rank; they are outside the box
    for (int d = -2; d <= 2; d++) { int z = cz + d; if (z >= zmin && z < zmax) s += in[c + d * pl]; }
    out[c] = s;
}

// ---- compact in-range lists -----------------------------------------------------------------
// one thread per sorted liquid particle; 25 rows x one contiguous span per row (x is the
// fastest cell axis), liquids then solids.  Lanes of a warp sit in x-adjacent cells, so
// their spans are shifted copies of each other: the float4 loads are near-coalesced L1 hits.
// cell_start lookup.  Single GPU: cs[c] is the slot of the first liquid of cell c.  Z-slab rank:
// the local sorted sequence is [ghost_lo | owned in-box | (owned out-of-box) | ghost_hi]; cs[] is an
// exclusive scan over that sequence without the out-of-box particles, so cells of the upper ghost
// layer (c >= hi_cell0) are shifted by their count.
struct CellStart { const int* cs; int base; int hi_cell0; int n_oob; int c_lo, c_hi; };   // cs[] is valid for cells in [c_lo, c_hi]
__device__ __forceinline__ int cs_at(const CellStart& C, int c) { return C.base + C.cs[c] + (c >= C.hi_cell0 ? C.n_oob : 0); }

__global__ void __launch_bounds__(WCSPH_BLOCK)
k_build_lists(const float4* __restrict__ pos, const int* __restrict__ keys_sorted, int i0, int nown, int SB, GridDims g,
              CellStart CS, const int* __restrict__ css, float cull_r,
              uint32_t* __restrict__ nbr_l, uint32_t* __restrict__ nbr_s, int capL, int capS,
              int* __restrict__ nl_cnt, int* __restrict__ ns_cnt, int* __restrict__ neighborCount,
              const int* __restrict__ boxsum, const unsigned char* __restrict__ m_self, const unsigned char* __restrict__ solid_near,
              int max_neighbour, Scalars* sc) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nown) return;
    const int i = i0 + li;
    int c = keys_sorted[li];
    if (c >= g.ncells) {            // HashGrid.py:81: outside the initial box -> no neighbours
        nl_cnt[li] = 0; ns_cnt[li] = 0; neighborCount[li] = 0; return;
    }
    const float4 pi = pos[i];
    const bool has_solid = solid_near[c] != 0;
    const int cx = c % g.bx, cy = (c / g.bx) % g.by, cz = c / (g.bx * g.by);
    // cells that can hold an in-range particle: [p - r, p + r] widened by 1e-3 cell against the
    // f32 rounding of cell_coords (Q19), clipped to the reference's +-2 stencil and the box
    const float pad = cull_r + 1e-3f * g.cell;
    int x0 = max(max((int)floorf((pi.x - pad - g.minx) * g.inv), cx - 2), 0);
    int x1 = min(min((int)floorf((pi.x + pad - g.minx) * g.inv), cx + 2), g.bx - 1);
    int y0 = max(max((int)floorf((pi.y - pad - g.miny) * g.inv), cy - 2), 0);
    int y1 = min(min((int)floorf((pi.y + pad - g.miny) * g.inv), cy + 2), g.by - 1);
    int z0 = max(max((int)floorf((pi.z - pad - g.minz) * g.inv), cz - 2), 0);
    int z1 = min(min((int)floorf((pi.z + pad - g.minz) * g.inv), cz + 2), g.bz - 1);
    const float r2max = cull_r * cull_r * (1.0f + 1e-5f);
    int nl = 0, ns = 0;
    // running store cursors: entry k of this particle's row sits at row base + (k >> 2) * 128 + (k & 3)
    uint32_t* const dstl = &NBR_AT(nbr_l, capL, li, 0);
    uint32_t* const dsts = &NBR_AT(nbr_s, capS, li, 0);
    int offl = 0, offs = 0;
    // per row (y, z): skip it if the row's cell column is farther than the cull radius in the yz plane,
    // else shrink the x span to the chord of the cull sphere at that distance (padded like above).
    // ~65 candidates are distance-tested instead of the ~125 of the full 5x5x5 block.
    // in cell units: f = position inside the particle's own cell (0..1), R = cull radius (+ pad)
    const float fx = (pi.x - g.minx) * g.inv - (float)cx, fy = (pi.y - g.miny) * g.inv - (float)cy, fz = (pi.z - g.minz) * g.inv - (float)cz;
    const float R = pad * g.inv, R2 = R * R;
    for (int z = z0; z <= z1; z++) {
        const int dzc = z - cz;
        const float az = dzc > 0 ? (float)dzc - fz : (dzc < 0 ? fz - (float)(dzc + 1) : 0.0f);     // distance to that layer
        for (int y = y0; y <= y1; y++) {
            const int dyc = y - cy;
            const float ay = dyc > 0 ? (float)dyc - fy : (dyc < 0 ? fy - (float)(dyc + 1) : 0.0f);
            const float rem = R2 - (ay * ay + az * az);
            if (rem < 0.0f) continue;
            const float xr = sqrtf(rem);
            const int xa = max(x0, cx + (int)floorf(fx - xr));
            const int xb = min(x1, cx + (int)floorf(fx + xr));
            if (xa > xb) continue;
            const int base = (z * g.by + y) * g.bx;
            int s = cs_at(CS, base + xa), e = cs_at(CS, base + xb + 1);
            // a row span never straddles the out-of-box block: rows are within one z layer
            for (int j = s; j < e; j++) {
                const float4 pj = pos[j];
                const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                const float r2 = dx * dx + dy * dy + dz * dz;
                if (r2 <= r2max && j != i) {
                    if (nl < capL) dstl[offl] = (uint32_t)j;
                    nl++;
                    offl += (nl & 3) ? 1 : 125;             // next slot of the uint4-grouped, warp-interleaved row (NBR_AT)
                }
            }
            if (!has_solid) continue;
            s = css[base + xa]; e = css[base + xb + 1];
            for (int j = SB + s; j < SB + e; j++) {
                const float4 pj = pos[j];
                const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                const float r2 = dx * dx + dy * dy + dz * dz;
                if (r2 <= r2max) {
                    if (ns < capS) dsts[offs] = (uint32_t)j;
                    ns++;
                    offs += (ns & 3) ? 1 : 125;
                }
            }
        }
    }
    nl_cnt[li] = nl; ns_cnt[li] = ns;
    int cnt = boxsum[c] - (int)m_self[c];
    neighborCount[li] = cnt;
    unsigned int fl = 0;
    if (nl > capL || ns > capS) fl |= WCSPH_FLAG_LIST_OVERFLOW;
    if (cnt > max_neighbour) fl |= WCSPH_FLAG_NEIGHBOR_OVERFLOW;      // Q3
    if (fl) atomicOr(&sc->flags, fl);
}

// k_build_lists, second version (round 2).  Same lists in the same order as the first one (rows z-major, y, then the row's span in
// slot order), one third of the issued instructions:
//  * the per-row chord of the cull sphere comes from three 5-entry tables of squared axis distances computed once per particle
//    (4 compares per row) instead of floor / sqrt / floor per row;
//  * list entries are collected four at a time in registers and leave as ONE 16-byte store per group (the warp-interleaved
//    layout makes the lane's group slot a whole uint4) instead of four scattered 4-byte stores;
//  * the pad of +-1e-3 cell against the f32 rounding of cell_coords (Q19) is kept, so the candidate set is a superset of the
//    first version's and the distance test decides: identical lists.
__device__ __forceinline__ void axis_d2(float f, float cell, int c, int n, float* a2) {
    // squared distance from the particle (offset f in [0, cell) inside its cell c) to the cell c + d, d = -2..2; cells outside [0, n) get +inf
#pragma unroll
    for (int d = -2; d <= 2; d++) {
        float a = d > 0 ? (float)d * cell - f : (d < 0 ? f - (float)(d + 1) * cell : 0.0f);
        a = fmaxf(a - 1e-3f * cell, 0.0f);
        a2[d + 2] = (c + d < 0 || c + d >= n) ? 3.0e38f : a * a;
    }
}

// one row span [s, e): distance test, accepted indices are shifted into a uint4 (newest in .w) that leaves as one 16-byte store per
// four entries.  SELF: the span can contain the particle itself (centre row only).
template <bool SELF>
__device__ __forceinline__ void build_row(const float4* __restrict__ pos, float4 pi, int i, int s, int e, float r2max,
                                          uint4* __restrict__ row4, int cap, int& n, uint4& grp) {
    const float4* p = pos + s;
#pragma unroll 2
    for (int j = s; j < e; j++, p++) {
        const float4 pj = __ldg(p);
        const float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
        const float r2 = dx * dx + dy * dy + dz * dz;
        if (r2 <= r2max && (!SELF || j != i)) {
            grp.x = grp.y; grp.y = grp.z; grp.z = grp.w; grp.w = (uint32_t)j;
            n++;
            if ((n & 3) == 0 && n <= cap) row4[(size_t)((n >> 2) - 1) * 32] = grp;
        }
    }
}
// the open group: its k = n & 3 entries sit in the LAST k lanes of grp -> rotate them to the front
__device__ __forceinline__ void flush_open_group(uint4* __restrict__ row4, int cap, int n, uint4 grp) {
    const int k = n & 3;
    if (!k || n >= cap) return;
    uint4 o;
    if (k == 1) o = make_uint4(grp.w, 0, 0, 0);
    else if (k == 2) o = make_uint4(grp.z, grp.w, 0, 0);
    else o = make_uint4(grp.y, grp.z, grp.w, 0);
    row4[(size_t)(n >> 2) * 32] = o;
}

// SLAB: z-slab rank (cell starts carry the ghost / out-of-box offsets, CellStart); single GPU reads cs[] directly.
// The walk runs on the SEARCH grid gs (cells of half the hash cell when F = 2) over +-S cells, S = ceil(cull_r / cell) (2, PCISPH
// on the refined grid 3); neighborCount, m_self and solid_near are tables of the REFERENCE grid g, indexed with the reference cell.
template <int S>
__device__ __forceinline__ void axis_d2s(float f, float cell, int c, int n, float* a2) {
#pragma unroll
    for (int d = -S; d <= S; d++) {
        float a = d > 0 ? (float)d * cell - f : (d < 0 ? f - (float)(d + 1) * cell : 0.0f);
        a = fmaxf(a - 1e-3f * cell, 0.0f);
        a2[d + S] = (c + d < 0 || c + d >= n) ? 3.0e38f : a * a;
    }
}
// five resident CTAs of 256 threads (48 registers, 8-32 B of spill) instead of three or four at 57 (S = 2) / 78 (S = 3) registers, measured A/B:
// DFSPH 1M 0.272 -> 0.225 ms, SESPH 1M 0.227 -> 0.198 ms, IISPH 2M 0.418 -> 0.340 ms, PCISPH 4M (S = 3) 1.838 -> 1.541 ms; 6 and 8 CTAs are
// no better.  (The sweeps are the opposite case: k_dfsph_drho at 48 registers without a spill is 6 % slower than at 64.)
#ifndef WCSPH_BL_MINB
#define WCSPH_BL_MINB 5
#endif
template <bool SLAB, int S>
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_BL_MINB)
k_build_lists2(const float4* __restrict__ pos, const int* __restrict__ keys_sorted, int i0, int nown, int SB, GridDims g, GridDims gs, int F,
               CellStart CS, const int* __restrict__ css, float cull_r,
               uint32_t* __restrict__ nbr_l, uint32_t* __restrict__ nbr_s, int capL, int capS,
               int* __restrict__ nl_cnt, int* __restrict__ ns_cnt, int* __restrict__ neighborCount,
               const int* __restrict__ boxsum, const unsigned char* __restrict__ m_self, const unsigned char* __restrict__ solid_near,
               int max_neighbour, Scalars* sc) {
    const int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nown) return;
    const int i = i0 + li;
    const int c = keys_sorted[li];
    if (c >= gs.ncells) {            // HashGrid.py:81: outside the initial box -> no neighbours
        nl_cnt[li] = 0; ns_cnt[li] = 0; neighborCount[li] = 0; return;
    }
    const float4 pi = pos[i];
    const int cx = c % gs.bx, cy = (c / gs.bx) % gs.by, cz = c / (gs.bx * gs.by);
    const int rc = F == 1 ? c : ((cz >> 1) * g.by + (cy >> 1)) * g.bx + (cx >> 1);          // the reference's cell
    const bool has_solid = solid_near[rc] != 0;
    float a2x[2 * S + 1], a2y[2 * S + 1];
    axis_d2s<S>((pi.x - gs.minx) - (float)cx * gs.cell, gs.cell, cx, gs.bx, a2x);
    axis_d2s<S>((pi.y - gs.miny) - (float)cy * gs.cell, gs.cell, cy, gs.by, a2y);
    const float fz = (pi.z - gs.minz) - (float)cz * gs.cell;
    const float r2max = cull_r * cull_r * (1.0f + 1e-5f);
    int nl = 0, ns = 0;
    uint4 gl = make_uint4(0, 0, 0, 0), gsol = gl;
    uint4* const rowl = (uint4*)nbr_l + ((size_t)(li >> 5) * (capL >> 2)) * 32 + (li & 31);
    uint4* const rows = (uint4*)nbr_s + ((size_t)(li >> 5) * (capS >> 2)) * 32 + (li & 31);
    const int* __restrict__ cs = CS.cs;
    // z layers in a loop (code size: unrolling all 25 rows starves the instruction cache -- measured, ncu stall no_instruction),
    // the y rows of a layer unrolled so that their distance table stays in registers
#pragma unroll 1
    for (int dz = -S; dz <= S; dz++) {
        if (cz + dz < 0 || cz + dz >= gs.bz) continue;
        float az = dz > 0 ? (float)dz * gs.cell - fz : (dz < 0 ? fz - (float)(dz + 1) * gs.cell : 0.0f);
        az = fmaxf(az - 1e-3f * gs.cell, 0.0f);
        const float remz = r2max - az * az;
        if (remz < 0.0f) continue;
#pragma unroll
        for (int dy = 0; dy < 2 * S + 1; dy++) {
            const float rem = remz - a2y[dy];
            if (rem >= 0.0f) {
                // chord: skip the x offsets whose whole cell is farther than the remaining budget (a2x falls towards the centre)
                int xa = cx - S, xb = cx + S;
#pragma unroll
                for (int d = 0; d < S; d++) { xa += (a2x[d] > rem); xb -= (a2x[2 * S - d] > rem); }
                const int base = ((cz + dz) * gs.by + (cy + dy - S)) * gs.bx;
                const int s = SLAB ? cs_at(CS, base + xa) : cs[base + xa];
                const int e = SLAB ? cs_at(CS, base + xb + 1) : cs[base + xb + 1];
                build_row<true>(pos, pi, i, s, e, r2max, rowl, capL, nl, gl);
                if (has_solid) build_row<false>(pos, pi, i, SB + css[base + xa], SB + css[base + xb + 1], r2max, rows, capS, ns, gsol);
            }
        }
    }
    flush_open_group(rowl, capL, nl, gl);
    flush_open_group(rows, capS, ns, gsol);
    nl_cnt[li] = nl; ns_cnt[li] = ns;
    const int cnt = boxsum[rc] - (int)m_self[rc];
    neighborCount[li] = cnt;
    unsigned int fl = 0;
    if (nl > capL || ns > capS) fl |= WCSPH_FLAG_LIST_OVERFLOW;
    if (cnt > max_neighbour) fl |= WCSPH_FLAG_NEIGHBOR_OVERFLOW;      // Q3
    if (fl) atomicOr(&sc->flags, fl);
}

// particles of a REFERENCE cell in the search-sorted arrays: F*F row fragments of F consecutive search cells each
__device__ __forceinline__ void ref_cell_frag(const GridDims& g, const GridDims& gs, int F, int rc, int frag, int& lo, int& hi) {
    if (F == 1) { lo = rc; hi = rc + 1; return; }
    const int x = rc % g.bx, y = (rc / g.bx) % g.by, z = rc / (g.bx * g.by);
    lo = ((2 * z + (frag >> 1)) * gs.by + 2 * y + (frag & 1)) * gs.bx + 2 * x;
    hi = lo + 2;
}

// Q1 fix-up: for a near-alias pair (c1,c2) every particle whose stencil holds both cells walks
// their shared bucket twice, i.e. sees the particles of c1 and of c2 one extra time each.
// (cells of the REFERENCE grid g; the particle arrays are sorted on the search grid gs)
__global__ void k_alias_fixup(const float4* __restrict__ pos, int i0, int nown, int SB, GridDims g, GridDims gs, int F, const int* __restrict__ pairs,
                              CellStart CS, const int* __restrict__ css, float cull_r,
                              uint32_t* __restrict__ nbr_l, uint32_t* __restrict__ nbr_s, int capL, int capS,
                              int* __restrict__ nl_cnt, int* __restrict__ ns_cnt, Scalars* sc) {
    int npairs = min(sc->alias_count, WCSPH_ALIAS_CAP);
    const float r2max = cull_r * cull_r * (1.0f + 1e-5f);
    const int nfrag = F * F;
    for (int p = blockIdx.x; p < npairs; p += gridDim.x) {
        int c1 = pairs[2 * p], c2 = pairs[2 * p + 1];
        // a z-slab rank only has cell starts for its own layers (+2 ghost layers): a pair with a cell
        // outside that range cannot sit in the stencil of an owned particle   (slab ranks run with F == 1: search cell == reference cell)
        if (F == 1 && (c1 < CS.c_lo || c1 >= CS.c_hi || c2 < CS.c_lo || c2 >= CS.c_hi)) continue;
        int x1 = c1 % g.bx, y1 = (c1 / g.bx) % g.by, z1 = c1 / (g.bx * g.by);
        int x2 = c2 % g.bx, y2 = (c2 / g.bx) % g.by, z2 = c2 / (g.bx * g.by);
        // nothing to duplicate if both cells are empty (locally: owned + ghost + solid)
        int n12 = 0;
        for (int side = 0; side < 2; side++) for (int f = 0; f < nfrag; f++) {
            int lo, hi; ref_cell_frag(g, gs, F, side ? c2 : c1, f, lo, hi);
            n12 += (cs_at(CS, hi) - cs_at(CS, lo)) + (css[hi] - css[lo]);
        }
        if (n12 == 0) continue;
        int lx = max(max(x1, x2) - 2, 0), hx = min(min(x1, x2) + 2, g.bx - 1);
        int ly = max(max(y1, y2) - 2, 0), hy = min(min(y1, y2) + 2, g.by - 1);
        int lz = max(max(z1, z2) - 2, 0), hz = min(min(z1, z2) + 2, g.bz - 1);
        int wx = hx - lx + 1, wy = hy - ly + 1, wz = hz - lz + 1;
        if (wx <= 0 || wy <= 0 || wz <= 0) continue;
        for (int t = threadIdx.x; t < wx * wy * wz * nfrag; t += blockDim.x) {
            const int tf = t % nfrag, tc = t / nfrag;
            int cc = ((lz + tc / (wx * wy)) * g.by + (ly + (tc / wx) % wy)) * g.bx + (lx + tc % wx);
            int ilo, ihi; ref_cell_frag(g, gs, F, cc, tf, ilo, ihi);
            const int ib = cs_at(CS, ilo), ie = cs_at(CS, ihi);
            for (int i = ib; i < ie; i++) {
                const int li = i - i0;
                if (li < 0 || li >= nown) continue;          // ghost particle: its owner appends
                float4 pi = pos[i];
                for (int side = 0; side < 2; side++) for (int f = 0; f < nfrag; f++) {
                    int jlo, jhi; ref_cell_frag(g, gs, F, side ? c2 : c1, f, jlo, jhi);
                    for (int j = cs_at(CS, jlo); j < cs_at(CS, jhi); j++) {
                        float4 pj = pos[j];
                        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                        if (dx * dx + dy * dy + dz * dz <= r2max && j != i) {
                            int slot = atomicAdd(&nl_cnt[li], 1);
                            if (slot < capL) NBR_AT(nbr_l, capL, li, slot) = (uint32_t)j;
                            else atomicOr(&sc->flags, WCSPH_FLAG_LIST_OVERFLOW);
                        }
                    }
                    for (int j = SB + css[jlo]; j < SB + css[jhi]; j++) {
                        float4 pj = pos[j];
                        float dx = pi.x - pj.x, dy = pi.y - pj.y, dz = pi.z - pj.z;
                        if (dx * dx + dy * dy + dz * dz <= r2max) {
                            int slot = atomicAdd(&ns_cnt[li], 1);
                            if (slot < capS) NBR_AT(nbr_s, capS, li, slot) = (uint32_t)j;
                            else atomicOr(&sc->flags, WCSPH_FLAG_LIST_OVERFLOW);
                        }
                    }
                }
            }
        }
    }
}

// clamp the counts to the list stride and pad each list to a multiple of 4 with the particle's
// own index (a self pair contributes exactly 0 to every gradW-weighted sum, see sweep.cuh)
__global__ void k_finish_lists(int* nl_cnt, int* ns_cnt, uint32_t* nbr_l, uint32_t* nbr_s, int i0, int nown, int capL, int capS) {
    int li = blockIdx.x * blockDim.x + threadIdx.x;
    if (li >= nown) return;
    int nl = min(nl_cnt[li], capL), ns = min(ns_cnt[li], capS);
    nl_cnt[li] = nl; ns_cnt[li] = ns;
    for (int k = nl; k < ((nl + 7) & ~7); k++) NBR_AT(nbr_l, capL, li, k) = (uint32_t)(i0 + li);      // pad to whole 8-groups
    for (int k = ns; k < ((ns + 7) & ~7); k++) NBR_AT(nbr_s, capS, li, k) = (uint32_t)(i0 + li);
}

__global__ void k_pack_pos(const float* __restrict__ xyz, float4* __restrict__ out, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) out[i] = make_float4(xyz[3 * i], xyz[3 * i + 1], xyz[3 * i + 2], 0.f);
}
__global__ void k_gather4(const float4* __restrict__ src, const int* __restrict__ perm, float4* __restrict__ dst, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) dst[i] = src[perm[i]];
}

static int radix_bits(int ncells) { int b = 1; while ((1LL << b) <= (long long)ncells) b++; return b; }

// keeps the liquids whose cell layer lies in [zlo, zhi) (z-slab rank); order is irrelevant (the first
// update_grid sorts), the reference index travels in sid
__global__ void k_select_owned(const float4* __restrict__ all, int NL, GridDims g, int zlo, int zhi, int last_rank,
                               float4* __restrict__ pos_own, int* __restrict__ sid_own, int cap, int* __restrict__ counter, Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    float4 p = all[i];
    int cx, cy, cz; cell_coords(g, p.x, p.y, p.z, cx, cy, cz);
    cz = min(max(cz, 0), g.bz - 1);
    if (cz >= zlo && (cz < zhi || last_rank)) {
        int slot = atomicAdd(counter, 1);
        if (slot < cap) { pos_own[slot] = p; sid_own[slot] = i; }
        else atomicOr(&sc->flags, WCSPH_FLAG_LIST_OVERFLOW);
    }
}

// ParticleData.setup_data_cpu ParticleData.py:180-185 + HashGrid.setup_grid_cpu HashGrid.py:44-54
extern "C" int wcsph_upload_pos(wcsph_ctx* c, const float* host_xyz) {
    if (!c || !host_xyz) return WCSPH_EINVAL;
    const int N = c->N, NL = c->NL, NS = c->NS, SB = c->SB, i0 = c->i0;
    cudaStream_t st = c->stream;
    FieldSlot* fp = wcsph_find_field(c, "pos");
    float4* pos0 = (float4*)fp->buf[0]; float4* pos1 = (float4*)fp->buf[1];
    c->cur = 0; c->host_scalars_valid = 0;
    // host xyz (insertion order) -> float4 scratch.  Single GPU: the scratch is pos buffer 1 itself
    // (liquids at [0,NL), solids behind); z-slab rank: the upper half of the staging area.
    CUDA_TRY(cudaMemcpyAsync(c->stage, host_xyz, (size_t)N * 12, cudaMemcpyHostToDevice, st));
    float4* tmp4 = c->R > 1 ? (float4*)(c->stage + (size_t)4 * N) : pos1;
    k_pack_pos<<<nblocks(N), WCSPH_BLOCK, 0, st>>>(c->stage, tmp4, N); LAUNCH_CHECK(c);
    const GridDims g = c->g;
    if (c->R > 1) {
        CUDA_TRY(cudaMemsetAsync(c->mg_counts, 0, 16 * sizeof(int), st));
        k_select_owned<<<nblocks(NL), WCSPH_BLOCK, 0, st>>>(tmp4, NL, g, c->zlo, c->zhi, c->rank == c->R - 1, pos0 + i0, c->sorted_id[0] + i0,
                                                        c->capOwn, c->mg_counts, c->sc); LAUNCH_CHECK(c);
        CUDA_TRY(cudaMemcpyAsync(c->mg_counts_host, c->mg_counts, sizeof(int), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaStreamSynchronize(st));
        c->nown = c->mg_counts_host[0] < c->capOwn ? c->mg_counts_host[0] : c->capOwn;
    } else {
        CUDA_TRY(cudaMemcpyAsync(pos0, tmp4, (size_t)NL * 16, cudaMemcpyDeviceToDevice, st));
        k_iota<<<nblocks(NL), WCSPH_BLOCK, 0, st>>>(c->sorted_id[0], NL); LAUNCH_CHECK(c);
        c->nown = NL;
    }
    k_iota<<<nblocks(c->CL > NS ? c->CL : NS), WCSPH_BLOCK, 0, st>>>(c->iota, c->CL > NS ? c->CL : NS); LAUNCH_CHECK(c);
    // static tables
    k_bucket_of_cell<<<nblocks(g.ncells), WCSPH_BLOCK, 0, st>>>(g, c->bucket_of_cell); LAUNCH_CHECK(c);
    k_m_self<<<nblocks(g.ncells), WCSPH_BLOCK, 0, st>>>(g, c->bucket_of_cell, c->m_self); LAUNCH_CHECK(c);
    CUDA_TRY(cudaMemsetAsync(&c->sc->alias_count, 0, 4, st));
    k_alias_pairs<<<nblocks(g.ncells), WCSPH_BLOCK, 0, st>>>(g, c->bucket_of_cell, c->alias_pairs, c->sc); LAUNCH_CHECK(c);
    // solids (replicated on every rank): keys -> sort once -> cell_start_s, occ_solid
    const GridDims gs = c->gs;
    CUDA_TRY(cudaMemsetAsync(c->cell_start_s, 0, ((size_t)gs.ncells + 2) * 4, st));
    CUDA_TRY(cudaMemsetAsync(c->occ_solid, 0, (size_t)N * 4, st));
    if (NS > 0) {
        k_keys<<<nblocks(NS), WCSPH_BLOCK, 0, st>>>(tmp4 + NL, NS, g, gs, c->F, c->keys, c->cell_start_s, c->occ_solid); LAUNCH_CHECK(c);
        size_t tb = c->cub_temp_bytes;
        CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->cub_temp, tb, c->keys, c->keys_sorted, c->iota, c->solid_sorted_id, NS, 0,
                                                 radix_bits(gs.ncells), st));
        c->launches += 4;
        k_gather4<<<nblocks(NS), WCSPH_BLOCK, 0, st>>>(tmp4 + NL, c->solid_sorted_id, pos0 + SB, NS); LAUNCH_CHECK(c);
        CUDA_TRY(cudaMemcpyAsync(pos1 + SB, pos0 + SB, (size_t)NS * 16, cudaMemcpyDeviceToDevice, st));
    }
    {
        size_t tb = c->cub_temp_bytes;
        CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->cub_temp, tb, c->cell_start_s, c->cell_start_s, gs.ncells + 1, st));
        c->launches += 2;
    }
    k_solid_near<<<nblocks(g.ncells), WCSPH_BLOCK, 0, st>>>(g, gs, c->F, c->cell_start_s, c->solid_near); LAUNCH_CHECK(c);
    CUDA_TRY(cudaStreamSynchronize(st));
    c->uploaded = 1;
    wcsph_invalidate_graphs(c);
    return 0;
}

int wcsph_mgpu_update_grid(wcsph_ctx* c);     // mgpu.cu

// S(c) = sum over the in-box 5x5x5 of occ[bucket(cell)]: separable box filter over the cell layers this rank needs, issued on
// c->stream (z-slab ranks run it on the side stream behind the occupancy all-reduce, concurrently with migration and sort)
int wcsph_box_filter(wcsph_ctx* c) {
    const GridDims g = c->g;
    cudaStream_t st = c->stream;
    const int plane = g.bx * g.by;
    const int zo0 = c->R > 1 ? max(c->zlo, 0) : 0, zo1 = c->R > 1 ? min(c->zhi, g.bz) : g.bz;        // layers whose S(c) is needed
    const int zh0 = max(zo0 - 2, 0), zh1 = min(zo1 + 2, g.bz);                                          // + the z stencil
    const int h0 = zh0 * plane, h1 = zh1 * plane, o0 = zo0 * plane, o1 = zo1 * plane;
    prof_begin(c, "k_box_x"); k_box_x<<<nblocks(h1 - h0), WCSPH_BLOCK, 0, st>>>(g, c->bucket_of_cell, c->occ, c->desc.max_in_grid > 0 ? c->desc.max_in_grid : 64, c->boxA, c->sc, h0, h1); prof_end(c); LAUNCH_CHECK(c);
    prof_begin(c, "k_box_y"); k_box_y<<<nblocks(h1 - h0), WCSPH_BLOCK, 0, st>>>(g, c->boxA, c->boxB, h0, h1); prof_end(c); LAUNCH_CHECK(c);
    prof_begin(c, "k_box_z"); k_box_z<<<nblocks(o1 - o0), WCSPH_BLOCK, 0, st>>>(g, c->boxB, c->boxA, o0, o1, zh0, zh1); prof_end(c); LAUNCH_CHECK(c);
    return 0;
}

// tail of update_grid shared by the single-GPU and the z-slab path: reference-exact neighborCount
// (S(c) = sum over the in-box 5x5x5 of occ[bucket(cell)]), compact in-range lists, Q1 duplicates
int wcsph_grid_finish(wcsph_ctx* c, CellStartArgs csa) {
    const GridDims g = c->g;
    cudaStream_t st = c->stream;
    FieldSlot* fp = wcsph_find_field(c, "pos");
    const float4* pos = (const float4*)fp->buf[c->cur];
    CellStart CS; CS.cs = c->cell_start_l; CS.base = csa.base; CS.hi_cell0 = csa.hi_cell0; CS.n_oob = csa.n_oob;
    CS.c_lo = csa.c_lo; CS.c_hi = csa.c_hi;
    if (!csa.box_done) TRY(wcsph_box_filter(c));
    prof_begin(c, "k_build_lists");
    const GridDims gs = c->gs;
    const int mxn = c->desc.max_neighbour > 0 ? c->desc.max_neighbour : 2048;
    const int S = (int)ceilf(c->cull_r / gs.cell - 1e-4f);          // stencil radius on the search grid: 2 (PCISPH on the refined grid: 3)
    if (S > 3 || S < 1) { wcsph_set_error("cull radius %g spans %d search cells of %g (supported: <= 3)", c->cull_r, S, gs.cell); return WCSPH_EINVAL; }
#define BL2_ARGS pos, c->keys_sorted, c->i0, c->nown, c->SB, g, gs, c->F, CS, c->cell_start_s, c->cull_r, c->nbr_l, c->nbr_s, c->capL, c->capS, \
                 c->nl_cnt, c->ns_cnt, c->neighborCount, c->boxA, c->m_self, c->solid_near, mxn, c->sc
    if (c->list_build_v1 && c->F == 1 && S <= 2)
        k_build_lists<<<nblocks(c->nown), WCSPH_BLOCK, 0, st>>>(pos, c->keys_sorted, c->i0, c->nown, c->SB, g, CS, c->cell_start_s, c->cull_r,
            c->nbr_l, c->nbr_s, c->capL, c->capS, c->nl_cnt, c->ns_cnt, c->neighborCount, c->boxA, c->m_self, c->solid_near, mxn, c->sc);
    else if (c->R > 1) { if (S <= 2) k_build_lists2<true, 2><<<nblocks(c->nown), WCSPH_BLOCK, 0, st>>>(BL2_ARGS); else k_build_lists2<true, 3><<<nblocks(c->nown), WCSPH_BLOCK, 0, st>>>(BL2_ARGS); }
    else { if (S <= 2) k_build_lists2<false, 2><<<nblocks(c->nown), WCSPH_BLOCK, 0, st>>>(BL2_ARGS); else k_build_lists2<false, 3><<<nblocks(c->nown), WCSPH_BLOCK, 0, st>>>(BL2_ARGS); }
#undef BL2_ARGS
    prof_end(c); LAUNCH_CHECK(c);
    prof_begin(c, "k_alias_fixup"); k_alias_fixup<<<296, 64, 0, st>>>(pos, c->i0, c->nown, c->SB, g, gs, c->F, c->alias_pairs, CS, c->cell_start_s, c->cull_r,
        c->nbr_l, c->nbr_s, c->capL, c->capS, c->nl_cnt, c->ns_cnt, c->sc); prof_end(c); LAUNCH_CHECK(c);
    prof_begin(c, "k_finish_lists"); k_finish_lists<<<nblocks(c->nown), WCSPH_BLOCK, 0, st>>>(c->nl_cnt, c->ns_cnt, c->nbr_l, c->nbr_s, c->i0, c->nown, c->capL, c->capS); prof_end(c); LAUNCH_CHECK(c);
    return 0;
}

// sort + permute of the n liquids at [i0, i0+n) by keys[0..n): shared by both paths
int wcsph_sort_permute(wcsph_ctx* c, int n, int kbase, int kspan) {
    const GridDims g = c->g;
    cudaStream_t st = c->stream;
    const int cur = c->cur, nxt = cur ^ 1;
    size_t tb = c->cub_temp_bytes;
    prof_begin(c, "cub_radix_sort");
    CUDA_TRY(cub::DeviceRadixSort::SortPairs(c->cub_temp, tb, c->keys, c->keys_sorted, c->iota, c->perm, n, 0, radix_bits((kspan > 0 ? kspan : c->gs.ncells) + 4), st));
    prof_end(c);
    c->launches += 4;
    PermuteArgs pa; memset(&pa, 0, sizeof(pa));
    for (int f = 0; f < c->nfields; f++) {
        FieldSlot& F = c->fields[f];
        if (!F.persistent) continue;
        if (F.stride == 4) { pa.src4[pa.n4] = (const float4*)F.buf[cur] + c->i0; pa.dst4[pa.n4] = (float4*)F.buf[nxt] + c->i0; pa.n4++; }
        else if (F.stride == 1) { pa.src1[pa.n1] = (const float*)F.buf[cur] + c->i0; pa.dst1[pa.n1] = (float*)F.buf[nxt] + c->i0; pa.n1++; }
    }
    prof_begin(c, "k_permute"); k_permute<<<nblocks(n), WCSPH_BLOCK, 0, st>>>(pa, c->perm, n, c->sorted_id[cur] + c->i0, c->sorted_id[nxt] + c->i0, c->keys_sorted, kbase, kspan, c->gs.ncells); prof_end(c); LAUNCH_CHECK(c);
    c->cur = nxt; c->inv_id_valid = 0;
    return 0;
}

// HashGrid.update_grid HashGrid.py:57-85
extern "C" int wcsph_hashgrid_update_grid(wcsph_ctx* c) {
    if (!c || !c->uploaded) { wcsph_set_error("update_grid before upload_pos"); return WCSPH_EINVAL; }
    if (c->R > 1) return wcsph_mgpu_update_grid(c);
    const int N = c->N, NL = c->NL;
    const GridDims g = c->g;
    cudaStream_t st = c->stream;
    if (NL == 0) return 0;
    FieldSlot* fp = wcsph_find_field(c, "pos");
    // 1. keys, true-cell histogram, bucket occupancy (solids' share is static)
    prof_begin(c, "grid_clear(memcpy occ + memset cells)");
    CUDA_TRY(cudaMemcpyAsync(c->occ, c->occ_solid, (size_t)N * 4, cudaMemcpyDeviceToDevice, st));
    const GridDims gs = c->gs;
    CUDA_TRY(cudaMemsetAsync(c->cell_start_l, 0, ((size_t)gs.ncells + 2) * 4, st));
    prof_end(c);
    PROF(c, "k_keys", (k_keys<<<nblocks(NL), WCSPH_BLOCK, 0, st>>>((const float4*)fp->buf[c->cur], NL, g, gs, c->F, c->keys, c->cell_start_l, c->occ))); LAUNCH_CHECK(c);
    // 2. stable sort by cell -> permutation of the persistent fields; exclusive scan -> cell starts
    TRY(wcsph_sort_permute(c, NL, 0, 0));
    size_t tb = c->cub_temp_bytes;
    prof_begin(c, "cub_exclusive_scan");
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->cub_temp, tb, c->cell_start_l, c->cell_start_l, gs.ncells + 1, st));
    prof_end(c);
    c->launches += 2;
    // 3. neighborCount, compact in-range lists (+ Q1 duplicates)
    CellStartArgs csa; csa.base = 0; csa.hi_cell0 = 0x7fffffff; csa.n_oob = 0; csa.c_lo = 0; csa.c_hi = gs.ncells; csa.box_done = 0;
    return wcsph_grid_finish(c, csa);
}

// statistics of the compact lists of the last update_grid: in-range (liquid, solid) pairs summed over the owned particles and
// the largest single list -- the pair count is what the FP32 figure of the roofline report is computed from
extern "C" int wcsph_pair_counts(wcsph_ctx* c, long long out[4]) {
    if (!c || !out) return WCSPH_EINVAL;
    out[0] = out[1] = out[2] = out[3] = 0;
    if (c->nown <= 0) return 0;
    std::vector<int> h((size_t)c->nown);
    for (int k = 0; k < 2; k++) {
        CUDA_TRY(cudaMemcpyAsync(h.data(), k ? c->ns_cnt : c->nl_cnt, (size_t)c->nown * 4, cudaMemcpyDeviceToHost, c->stream));
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        long long s = 0; int mx = 0;
        for (int v : h) { s += v; if (v > mx) mx = v; }
        out[k] = s; out[2 + k] = mx;
    }
    return 0;
}

// lazy debug view of HashGrid.neighbor restricted to in-range candidates, reference indices
__global__ void k_neighbors_of(int slot, int NL, int SB, const uint32_t* nbr_l, const uint32_t* nbr_s, int capL, int capS,
                               const int* nl_cnt, const int* ns_cnt, const int* sid, const int* solid_sid, int* out) {
    int nl = nl_cnt[slot], ns = ns_cnt[slot];
    for (int k = threadIdx.x; k < nl + ns; k += blockDim.x) {
        if (k < nl) out[1 + k] = sid[NBR_AT(nbr_l, capL, slot, k)];
        else out[1 + k] = NL + solid_sid[NBR_AT(nbr_s, capS, slot, k - nl) - SB];
    }
    if (threadIdx.x == 0) out[0] = nl + ns;
}
__global__ void k_invert(const int* sid, int* inv, int n) { int k = blockIdx.x * blockDim.x + threadIdx.x; if (k < n) inv[sid[k]] = k; }

extern "C" int wcsph_hashgrid_neighbors_of(wcsph_ctx* c, int ref_index, int* host_out, int cap, int* n_out) {
    if (!c || ref_index < 0 || ref_index >= c->NL || !host_out || !n_out) return WCSPH_EINVAL;
    cudaStream_t st = c->stream;
    if (!c->inv_id_valid) { k_invert<<<nblocks(c->NL), WCSPH_BLOCK, 0, st>>>(c->sorted_id[c->cur], c->inv_id, c->NL); LAUNCH_CHECK(c); c->inv_id_valid = 1; }
    int slot = 0;
    CUDA_TRY(cudaMemcpyAsync(&slot, c->inv_id + ref_index, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int* out = (int*)c->stage;
    k_neighbors_of<<<1, 128, 0, st>>>(slot, c->NL, c->SB, c->nbr_l, c->nbr_s, c->capL, c->capS, c->nl_cnt, c->ns_cnt,
                                      c->sorted_id[c->cur], c->solid_sorted_id, out); LAUNCH_CHECK(c);
    int total = 0;
    CUDA_TRY(cudaMemcpyAsync(&total, out, 4, cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    int ncopy = total < cap ? total : cap;
    if (ncopy > 0) CUDA_TRY(cudaMemcpy(host_out, out + 1, (size_t)ncopy * 4, cudaMemcpyDeviceToHost));
    *n_out = total;
    return 0;
}
