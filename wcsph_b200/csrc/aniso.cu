// aniso.cu -- SURVEY 8(f) N2, second half: the anisotropic-kernel branch of the surface reconstruction (Yu & Turk 2013) that
// the reference keeps next to the active one (MarchingCubeGrid.py:148-149 has the two calls commented out):
//   ParticleData.compute_color_map      ParticleData.py:187-218   colour field + its normalised gradient per liquid particle
//   ParticleData.cal_anistropic_kernel  ParticleData.py:220-285   weighted mean position, covariance, G = R diag(1/sigma') R^T
//   MCGrid.cal_surface_point_anistropic MarchingCubeGrid.py:215-246 (in mc.cu: it shares the cell-sorted particle array)
//
// Which pairs enter the sums.  Both kernels of ParticleData loop over HashGrid.neighbor[i, 0:neighborCount[i]], i.e. over EVERY
// candidate of the 125-bucket walk (HashGrid.py:79-106), alias duplicates included.
//  * compute_color_map weights them with the cubic kernel of support h = 2 hash cells;
//  * cal_anistropic_kernel weights them with 1 - (d / 2R_mc)^3, d < 2R_mc = 0.18 m > h: candidates beyond the compact lists count.
//    For both the candidate multiset is rebuilt from its definition: particle j in cell b appears in i's list once for every in-box
//    cell s of i's 5x5x5 stencil with bucket(s) == bucket(b).  That is (1) the stencil walk itself, over the true cells, plus
//    (2) one extra appearance for every ORDERED pair of distinct in-box cells (s, b) with equal bucket, s in the stencil of i.
//    Since only d < 0.18 m = 3.6 cells matters, b lies within Chebyshev distance 4 of i's cell and the pair within distance 6:
//    a static table (built on first use, like the distance-4 table the list build uses).
// The eigen-decomposition replaces ti.svd (ParticleData.py:270): for the symmetric positive semi-definite covariance the two agree
// and R diag(f(sigma)) R^T does not depend on eigenvector signs (cyclic Jacobi in double, in registers).
#include "engine.cuh"

#define ANISO_PAIR_CAP 262144

struct AnisoWork {
    float* mom;          // [nown][10]: sum w, sum w d (3), sum w d d^T (xx, xy, xz, yy, yz, zz), d = x_j - x_i
    int* pairs;          // [ANISO_PAIR_CAP][2] ordered alias pairs (s, b)
    int* npairs;         // [1]
    int* tag;            // [2]: n_hash and ncells the table was built for (rebuilt when they change)
    size_t total;
};
static size_t al256(size_t x) { return (x + 255) & ~(size_t)255; }
static AnisoWork aniso_carve(char* base, int nown) {
    AnisoWork w; size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += al256(bytes); return p; };
    w.mom = (float*)take((size_t)(nown > 0 ? nown : 1) * 10 * sizeof(float));
    w.pairs = (int*)take((size_t)ANISO_PAIR_CAP * 2 * sizeof(int));
    w.npairs = (int*)take(256);
    w.tag = w.npairs + 8;
    w.total = off;
    return w;
}

extern "C" size_t wcsph_pd_aniso_workspace_bytes(wcsph_ctx* c) {
    if (!c) return 0;
    return aniso_carve(nullptr, c->capOwn).total;
}

// ---- the candidate walk shared by the three sums ------------------------------------------------------------------------------
// MODE 0: color (liquid + solid neighbours, cubic kernel of support h)          ParticleData.py:189-203
// MODE 1: color_grad numerator (liquid neighbours)                             ParticleData.py:205-217
// MODE 2: anisotropy moments (liquid neighbours, weight 1 - (d / 2 R_mc)^3)   ParticleData.py:226-266
// The walk uses the cells of the LAST update_grid (cell_start tables, keys_sorted) with the CURRENT positions, like the
// reference, whose candidate table is the one HashGrid.update_grid left while pos has moved on (update_pos).  The compact in-range
// lists are not used here: they are culled at h with the positions of update_grid time and would miss pairs that came into range.
struct WalkArgs {
    const float4* pos; const float* rho; const float* color;     // slot-indexed (liquids), solids behind SB in pos
    const int* cs; const int* css; int SB;
    KC k; float R2, R2inv;
};
#include "sweep.cuh"

template <int MODE>
__device__ __forceinline__ void pair_term(const WalkArgs& A, float* m, float3 pi, int j, bool solid) {
    const float4 pj4 = A.pos[j];
    const float3 d = f3(pj4.x - pi.x, pj4.y - pi.y, pj4.z - pi.z);             // x_j - x_i
    const float r2 = dot3(d, d);
    if (MODE == 0) {
        const float W = cubic_W2(A.k, r2);
        if (solid) m[1] += W; else m[0] += __fdividef(A.k.mass, A.rho[j]) * W;
    } else if (MODE == 1) {
        // gradW(x_i - x_j) = -s * d
        const float s = cubic_gradW_s(A.k, r2) * __fdividef(A.k.mass, A.rho[j]) * A.color[j];
        m[0] -= s * d.x; m[1] -= s * d.y; m[2] -= s * d.z;
    } else {
        const float dis = sqrtf(r2);
        if (dis < A.R2) {
            const float q = dis * A.R2inv;
            const float w = 1.0f - q * q * q;              // 1 - pow(dis / (2 R_mc), 3.0)  ParticleData.py:296
            m[0] += w; m[1] += w * d.x; m[2] += w * d.y; m[3] += w * d.z;
            m[4] += w * d.x * d.x; m[5] += w * d.x * d.y; m[6] += w * d.x * d.z; m[7] += w * d.y * d.y; m[8] += w * d.y * d.z; m[9] += w * d.z * d.z;
        }
    }
}
template <int MODE> struct NAcc { static const int n = MODE == 0 ? 2 : (MODE == 1 ? 3 : 10); };

// (1) the stencil walk over the true cells: every particle j != i of the in-box 5x5x5 block, 25 row spans (+ the solids' spans)
template <int MODE>
__global__ void __launch_bounds__(WCSPH_BLOCK)
k_walk(WalkArgs A, const int* __restrict__ keys_sorted, int nown, GridDims g, float* __restrict__ mom) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nown) return;
    float m[10];
#pragma unroll
    for (int k = 0; k < 10; k++) m[k] = 0.f;
    const int c = keys_sorted[i];
    if (c < g.ncells) {                                     // outside the box: no candidates (HashGrid.py:81)
        const float3 pi = xyz(A.pos[i]);
        const int cx = c % g.bx, cy = (c / g.bx) % g.by, cz = c / (g.bx * g.by);
        const int x0 = max(cx - 2, 0), x1 = min(cx + 2, g.bx - 1);
        for (int z = max(cz - 2, 0); z <= min(cz + 2, g.bz - 1); z++)
            for (int y = max(cy - 2, 0); y <= min(cy + 2, g.by - 1); y++) {
                const int base = (z * g.by + y) * g.bx;
                for (int j = A.cs[base + x0]; j < A.cs[base + x1 + 1]; j++) if (j != i) pair_term<MODE>(A, m, pi, j, false);
                if (MODE == 0)
                    for (int j = A.SB + A.css[base + x0]; j < A.SB + A.css[base + x1 + 1]; j++) pair_term<MODE>(A, m, pi, j, true);
            }
    }
#pragma unroll
    for (int k = 0; k < NAcc<MODE>::n; k++) mom[(size_t)i * 10 + k] = m[k];
}

// ordered pairs (s, b), s != b, both in the box, bucket(s) == bucket(b), Chebyshev distance <= 6
__global__ void k_aniso_pairs(GridDims g, const int* __restrict__ boc, int* __restrict__ pairs, int* __restrict__ npairs, Scalars* sc) {
    int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= g.ncells) return;
    const int cx = c % g.bx, cy = (c / g.bx) % g.by, cz = c / (g.bx * g.by);
    const int b = boc[c];
    for (int dz = -6; dz <= 6; dz++) for (int dy = -6; dy <= 6; dy++) for (int dx = -6; dx <= 6; dx++) {
        if (!(dx | dy | dz)) continue;
        const int x = cx + dx, y = cy + dy, z = cz + dz;
        if (!in_box(g, x, y, z)) continue;
        const int c2 = (z * g.by + y) * g.bx + x;
        if (boc[c2] == b) {
            const int slot = atomicAdd(npairs, 1);
            if (slot < ANISO_PAIR_CAP) { pairs[2 * slot] = c; pairs[2 * slot + 1] = c2; }
            else atomicOr(&sc->flags, WCSPH_FLAG_ALIAS_OVERFLOW);
        }
    }
}

// (2) alias extras: for the ordered pair (s, b) every particle i whose stencil holds s sees the particles of b once more.
// `reach`: how many cells away b can be from i's cell and still matter (support / cell size, + 1 for the staleness of the cells)
template <int MODE>
__global__ void k_walk_alias(WalkArgs A, GridDims g, const int* __restrict__ pairs, const int* __restrict__ npairs, int reach,
                             float* __restrict__ mom) {
    const int np = min(*npairs, ANISO_PAIR_CAP);
    for (int p = blockIdx.x; p < np; p += gridDim.x) {
        const int s = pairs[2 * p], b = pairs[2 * p + 1];
        const int jb = A.cs[b], je = A.cs[b + 1];
        const int sb = MODE == 0 ? A.css[b] : 0, se = MODE == 0 ? A.css[b + 1] : 0;
        if (jb == je && sb == se) continue;
        const int sx = s % g.bx, sy = (s / g.bx) % g.by, sz = s / (g.bx * g.by);
        const int bx_ = b % g.bx, by_ = (b / g.bx) % g.by, bz_ = b / (g.bx * g.by);
        // cells c with |c - s| <= 2 (s is in c's stencil) and |c - b| <= reach
        const int lx = max(max(sx - 2, bx_ - reach), 0), hx = min(min(sx + 2, bx_ + reach), g.bx - 1);
        const int ly = max(max(sy - 2, by_ - reach), 0), hy = min(min(sy + 2, by_ + reach), g.by - 1);
        const int lz = max(max(sz - 2, bz_ - reach), 0), hz = min(min(sz + 2, bz_ + reach), g.bz - 1);
        const int wx = hx - lx + 1, wy = hy - ly + 1, wz = hz - lz + 1;
        if (wx <= 0 || wy <= 0 || wz <= 0) continue;
        for (int t = threadIdx.x; t < wx * wy * wz; t += blockDim.x) {
            const int cc = ((lz + t / (wx * wy)) * g.by + (ly + (t / wx) % wy)) * g.bx + (lx + t % wx);
            for (int i = A.cs[cc]; i < A.cs[cc + 1]; i++) {
                float m[10];
#pragma unroll
                for (int k = 0; k < 10; k++) m[k] = 0.f;
                const float3 pi = xyz(A.pos[i]);
                for (int j = jb; j < je; j++) if (j != i) pair_term<MODE>(A, m, pi, j, false);
                if (MODE == 0) for (int j = A.SB + sb; j < A.SB + se; j++) pair_term<MODE>(A, m, pi, j, true);
#pragma unroll
                for (int k = 0; k < NAcc<MODE>::n; k++) if (m[k] != 0.f) atomicAdd(&mom[(size_t)i * 10 + k], m[k]);
            }
        }
    }
}

static int aniso_prepare(wcsph_ctx* c, void* work_dev, size_t work_bytes, AnisoWork* w, const char* fn) {
    if (c->R > 1) { wcsph_set_error("%s runs on a single-GPU context", fn); return WCSPH_EINVAL; }
    if (!c->uploaded) { wcsph_set_error("%s before upload_pos", fn); return WCSPH_EINVAL; }
    if (c->F != 1) { wcsph_set_error("%s: needs a context whose particles are sorted on the reference's hash grid (ParticleData(particleRadius), "
                                     "as dfsph.py constructs it); this one sorts on a refined search grid", fn); return WCSPH_EINVAL; }
    *w = aniso_carve((char*)work_dev, c->capOwn);
    if (work_bytes < w->total) { wcsph_set_error("%s: workspace %zu < %zu bytes (wcsph_pd_aniso_workspace_bytes)", fn, work_bytes, w->total); return WCSPH_EINVAL; }
    const GridDims g = c->g;
    cudaStream_t st = c->stream;
    // the pair table is static per (hash modulus, grid); the tag says whether this workspace already holds it
    int tag[2] = {0, 0};
    CUDA_TRY(cudaMemcpyAsync(tag, w->tag, sizeof(tag), cudaMemcpyDeviceToHost, st));
    CUDA_TRY(cudaStreamSynchronize(st));
    if (tag[0] != g.n_hash || tag[1] != g.ncells) {
        CUDA_TRY(cudaMemsetAsync(w->npairs, 0, sizeof(int), st));
        prof_begin(c, "k_aniso_pairs");
        k_aniso_pairs<<<nblocks(g.ncells), WCSPH_BLOCK, 0, st>>>(g, c->bucket_of_cell, w->pairs, w->npairs, c->sc);
        prof_end(c); LAUNCH_CHECK(c);
        tag[0] = g.n_hash; tag[1] = g.ncells;
        CUDA_TRY(cudaMemcpyAsync(w->tag, tag, sizeof(tag), cudaMemcpyHostToDevice, st));
        CUDA_TRY(cudaStreamSynchronize(st));
    }
    return 0;
}
static WalkArgs walk_args(wcsph_ctx* c, const float* color, float R2) {
    WalkArgs A;
    A.pos = fcur<float4>(c, "pos"); A.rho = fcur<float>(c, "rho"); A.color = color;
    A.cs = c->cell_start_l; A.css = c->cell_start_s; A.SB = c->SB;
    A.k = make_kc(c->prm); A.R2 = R2; A.R2inv = R2 > 0.f ? 1.0f / R2 : 0.f;
    return A;
}
template <int MODE>
static int run_walk(wcsph_ctx* c, const AnisoWork& w, const WalkArgs& A, int reach, const char* name) {
    const int n = c->nown;
    prof_begin(c, name);
    k_walk<MODE><<<nblocks(n), WCSPH_BLOCK, 0, c->stream>>>(A, c->keys_sorted, n, c->g, w.mom);
    prof_end(c); LAUNCH_CHECK(c);
    prof_begin(c, "k_walk_alias");
    k_walk_alias<MODE><<<592, 64, 0, c->stream>>>(A, c->g, w.pairs, w.npairs, reach, w.mom);
    prof_end(c); LAUNCH_CHECK(c);
    return 0;
}

// ---- compute_color_map (ParticleData.py:187-218) ------------------------------------------------------------------------------
__global__ void k_color_finish(const float* __restrict__ mom, const float* __restrict__ rho, KC K, int n, float* __restrict__ color) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    color[i] = K.mass / rho[i] * cubic_W(K, 0.f) + mom[(size_t)i * 10] + K.VS0 * mom[(size_t)i * 10 + 1];     // :191-203
}
__global__ void k_color_grad_finish(const float* __restrict__ mom, const float* __restrict__ color, int n, float4* __restrict__ grad) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float inv = 1.0f / color[i];                                                                       // :218
    grad[i] = make_float4(mom[(size_t)i * 10] * inv, mom[(size_t)i * 10 + 1] * inv, mom[(size_t)i * 10 + 2] * inv, 0.f);
}

// color / color_grad: caller-owned device buffers in SLOT order (the cell-sorted order of wcsph_field_device views), CL entries.
// kernel_c of ParticleData is CubicKernel(hash_grid.searchR) (ParticleData.py:31), support = 2 hash cells = the stencil reach.
extern "C" int wcsph_pd_compute_color_map(wcsph_ctx* c, void* work_dev, size_t work_bytes, float* color_dev, float* color_grad4_dev) {
    if (!c || !work_dev || !color_dev || !color_grad4_dev) { wcsph_set_error("compute_color_map: null argument"); return WCSPH_EINVAL; }
    AnisoWork w; TRY(aniso_prepare(c, work_dev, work_bytes, &w, "compute_color_map"));
    WalkArgs A = walk_args(c, color_dev, 0.f);
    // ParticleData's own kernel: support hash_grid.searchR = 2 * gridR whatever the solver module uses for its physics
    const double h = 2.0 * c->desc.hash_gridR;
    A.k.h = (float)h; A.k.inv_h = (float)(1.0 / h);
    A.k.m_k = (float)(8.0 / 3.14159265358979323846 / (h * h * h));
    A.k.m_l = (float)(48.0 / 3.14159265358979323846 / (h * h * h));
    A.k.m_l_h = (float)(48.0 / 3.14159265358979323846 / (h * h * h) / h);
    const int n = c->nown;
    TRY(run_walk<0>(c, w, A, 3, "k_walk<color>"));
    k_color_finish<<<nblocks(n), WCSPH_BLOCK, 0, c->stream>>>(w.mom, fcur<float>(c, "rho"), A.k, n, color_dev); LAUNCH_CHECK(c);
    TRY(run_walk<1>(c, w, A, 3, "k_walk<color_grad>"));
    k_color_grad_finish<<<nblocks(n), WCSPH_BLOCK, 0, c->stream>>>(w.mom, color_dev, n, (float4*)color_grad4_dev); LAUNCH_CHECK(c);
    return 0;
}

// ---- cal_anistropic_kernel (ParticleData.py:220-285) ---------------------------------------------------------------------------
// cyclic Jacobi for a symmetric 3x3, eigenvalues descending, V columns = eigenvectors
__device__ void sym_eig3_dev(double a[3][3], double w[3], double V[3][3]) {
    for (int r = 0; r < 3; r++) for (int cc = 0; cc < 3; cc++) V[r][cc] = r == cc;
    for (int sweep = 0; sweep < 32; sweep++) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1e-300 || off < 1e-18 * (fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]))) break;
        for (int pq = 0; pq < 3; pq++) {
            const int p = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            if (fabs(a[p][q]) < 1e-300) continue;
            const double th = (a[q][q] - a[p][p]) / (2.0 * a[p][q]);
            const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
            const double co = 1.0 / sqrt(t * t + 1.0), si = t * co;
            for (int k = 0; k < 3; k++) { const double x = a[k][p], y = a[k][q]; a[k][p] = co * x - si * y; a[k][q] = si * x + co * y; }
            for (int k = 0; k < 3; k++) { const double x = a[p][k], y = a[q][k]; a[p][k] = co * x - si * y; a[q][k] = si * x + co * y; }
            for (int k = 0; k < 3; k++) { const double x = V[k][p], y = V[k][q]; V[k][p] = co * x - si * y; V[k][q] = si * x + co * y; }
        }
    }
    for (int k = 0; k < 3; k++) w[k] = a[k][k];
    for (int x = 0; x < 2; x++) for (int y = x + 1; y < 3; y++) if (w[y] > w[x]) {
        const double t = w[x]; w[x] = w[y]; w[y] = t;
        for (int k = 0; k < 3; k++) { const double u = V[k][x]; V[k][x] = V[k][y]; V[k][y] = u; }
    }
}

// :238-241 pos_avr, :243-279 G (kr 4, ks 1400, kn 0.5, ne 25)
__global__ void k_aniso_finish(const float4* __restrict__ pos, const float* __restrict__ mom, const int* __restrict__ ncount, int nown,
                               float4* __restrict__ pos_avr, float4* __restrict__ G) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= nown) return;
    const float* m = mom + (size_t)i * 10;
    const float3 pi = xyz(pos[i]);
    const float sw = m[0];
    float3 mean = f3(0, 0, 0);
    if (sw > 0.0f) mean = f3(m[1] / sw, m[2] / sw, m[3] / sw);
    pos_avr[i] = f4(pi + mean);
    const float kr = 4.0f, ks = 1400.0f, kn = 0.5f, ne = 25.0f;
    float g9[9] = {kn, 0, 0, 0, kn, 0, 0, 0, kn};
    if ((float)ncount[i] > ne && sw > 0.0f) {
        // C = sum w (d - mean)(d - mean)^T / sum w = sum w d d^T / sum w - mean mean^T
        double C[3][3], w3[3], R[3][3];
        const double s = (double)sw, mx = (double)m[1] / s, my = (double)m[2] / s, mz = (double)m[3] / s;
        C[0][0] = (double)m[4] / s - mx * mx; C[0][1] = C[1][0] = (double)m[5] / s - mx * my; C[0][2] = C[2][0] = (double)m[6] / s - mx * mz;
        C[1][1] = (double)m[7] / s - my * my; C[1][2] = C[2][1] = (double)m[8] / s - my * mz; C[2][2] = (double)m[9] / s - mz * mz;
        sym_eig3_dev(C, w3, R);
        const float s0 = (float)w3[0], s1 = (float)w3[1], s2 = (float)w3[2];
        if (s0 > 0.0f) {
            const float inv[3] = {1.0f / (ks * s0), 1.0f / (ks * fmaxf(s1, s0 / kr)), 1.0f / (ks * fmaxf(s2, s0 / kr))};
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
                double acc = 0.0;
                for (int k = 0; k < 3; k++) acc += R[a][k] * (double)inv[k] * R[b][k];
                g9[3 * a + b] = (float)acc;
            }
        }
    }
    G[3 * (size_t)i] = make_float4(g9[0], g9[1], g9[2], 0.f);
    G[3 * (size_t)i + 1] = make_float4(g9[3], g9[4], g9[5], 0.f);
    G[3 * (size_t)i + 2] = make_float4(g9[6], g9[7], g9[8], 0.f);
}

// pos_avr: float4 per slot; G: three float4 rows per slot (slot order, like wcsph_field_device views of a 3x3 field)
extern "C" int wcsph_pd_cal_anistropic_kernel(wcsph_ctx* c, float mc_searchR, void* work_dev, size_t work_bytes, float* pos_avr4_dev, float* G12_dev) {
    if (!c || !work_dev || !pos_avr4_dev || !G12_dev || !(mc_searchR > 0.f)) { wcsph_set_error("cal_anistropic_kernel: bad argument"); return WCSPH_EINVAL; }
    AnisoWork w; TRY(aniso_prepare(c, work_dev, work_bytes, &w, "cal_anistropic_kernel"));
    const float R2 = (float)((double)mc_searchR * 2.0);            // self.mc_grid.searchR*2.0, folded in Python (ParticleData.py:295-296)
    WalkArgs A = walk_args(c, nullptr, R2);
    // a particle of cell b can be within R2 of one in cell c only if |c - b| <= ceil(R2 / cell) (+ 1: the cells are one step old)
    const int reach = (int)ceilf(R2 / c->g.cell) + 1;
    if (reach > 4 + 1) { wcsph_set_error("cal_anistropic_kernel: 2 * mc_searchR = %g spans more than 4 hash cells of %g", R2, c->g.cell); return WCSPH_EINVAL; }
    TRY(run_walk<2>(c, w, A, reach > 4 ? 4 : reach, "k_walk<aniso>"));
    const int n = c->nown;
    prof_begin(c, "k_aniso_finish");
    k_aniso_finish<<<nblocks(n), WCSPH_BLOCK, 0, c->stream>>>(A.pos, w.mom, c->neighborCount, n, (float4*)pos_avr4_dev, (float4*)G12_dev);
    prof_end(c); LAUNCH_CHECK(c);
    return 0;
}
