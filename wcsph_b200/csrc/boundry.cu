// boundry.cu -- SURVEY 8(f) N3: the boundary pre-processing tool of the reference (boundry.py): parallel Poisson-disk sampling of a
// triangle mesh (Bowers et al. 2010) whose output, `<mesh>_boundry.obj`, is the solid point cloud the solver scripts load
// (dfsph.py:597, iisph.py:411).  Stages, each a former @ti.kernel:
//   init_point_set      boundry.py:223-247   area-weighted random points on the triangles (ti.random: own counter-based generator
//                                             here, or an injected point set)
//   gpu_bitonic_sort    :208-219,322-336     sort by cell, the SAME network (the order inside a cell decides which point a trial picks)
//   build_hmap          :250-271             first point of every cell -> hash slot; 27 phase groups of cells
//   possion_disk_sample :390-407             per (phase, trial): the cell's trial-th point is accepted if no sample of the 5x5x5
//                                             neighbourhood is closer than particleRadius in the geodesic-corrected distance (:340-373)
// The reference's parallel loops race in two places; their serial meaning (ascending index) is implemented deterministically:
//   * two cells with the same hash slot: the later one keeps it (atomicMax on the start index);
//   * append order of phase_group / possion_sample: prefix sums instead of atomic counters.
// A launch evaluates all its cells against the state BEFORE the launch and then applies the acceptances in order -- equal to the
// serial run because cells of one phase group are >= 3 cells = 2 gridR * sqrt(3)... >= 2 * gridR > particleRadius apart, and a second
// phase-group entry that resolves to the same hash slot (it re-tests the same point) can never be accepted after the first.
// Kept as written: tri_normal[face id] on a per-VERTEX array (:361-362), unverified hash slots (:345-349), the count reset to 4 on
// a sixth sample (:400-402).  Compiled WITHOUT -use_fast_math: the acceptance test compares against particleRadius, IEEE divide /
// sqrt / asinf keep the decisions those of the restatement.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <cub/device/device_radix_sort.cuh>
#include <cub/device/device_scan.cuh>
#include "../../include/wcsph_b200.h"

void wcsph_set_error(const char* fmt, ...);
#define BD_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { wcsph_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return WCSPH_ECUDA; } } while (0)
#define BD_BLOCK 256
static inline int bd_blocks(long long n) { return (int)((n + BD_BLOCK - 1) / BD_BLOCK > 0 ? (n + BD_BLOCK - 1) / BD_BLOCK : 1); }

struct BdWork {
    int4* cell;            // [P] x, y, z, (unused)
    float4* pos;           // [P] x, y, z, id (int bits)
    int* start_index;      // [H]
    int4* hcell;           // [H]
    int* hash_trace;       // [n]
    int* phase_key;        // [n] phase of a head, 27 otherwise
    int* phase_key_sorted; int* head_idx; int* head_idx_sorted;
    int* phase_group_count;// [32]: 27 counts, [27] = heads total
    int* phase_offset;     // [32]
    int4* phase_group;     // [27][V]
    unsigned char* dup;    // [27][V] 1 if an earlier entry of the same phase group resolves to the same hash slot
    int* sample_count;     // [H]
    int* sample;           // [H][cap]
    float* possion_sample; // [n][3]
    int* selected;         // [n]
    int* accept;           // [V] flag, [V] exclusive scan
    int* accept_scan;
    int* cand;             // [V]
    unsigned char* foreign;// [V] the entry's hash slot belongs to ANOTHER cell (collision): its candidate is that cell's point
    int* fin_list;         // [V] scratch of the serial fix-up
    int* counters;         // [0] samples so far, [1] hash_count, [2] phase overflow flag, [3] sample overflow events, [4] launch has an accepted foreign entry
    void* cub; size_t cub_bytes;
    size_t total;
};
static size_t bal(size_t x) { return (x + 255) & ~(size_t)255; }
static BdWork bd_carve(char* base, const wcsph_bd_desc* d) {
    BdWork w; size_t off = 0;
    auto take = [&](size_t bytes) { char* p = base ? base + off : nullptr; off += bal(bytes ? bytes : 1); return p; };
    const size_t P = d->padding, H = d->hash_size, n = d->n, V = d->phase_vec_max > 0 ? d->phase_vec_max : 1;
    w.cell = (int4*)take(P * 16); w.pos = (float4*)take(P * 16);
    w.start_index = (int*)take(H * 4); w.hcell = (int4*)take(H * 16);
    w.hash_trace = (int*)take(n * 4);
    w.phase_key = (int*)take(n * 4); w.phase_key_sorted = (int*)take(n * 4); w.head_idx = (int*)take(n * 4); w.head_idx_sorted = (int*)take(n * 4);
    w.phase_group_count = (int*)take(32 * 4); w.phase_offset = (int*)take(32 * 4);
    w.phase_group = (int4*)take(27 * V * 16); w.dup = (unsigned char*)take(27 * V);
    w.sample_count = (int*)take(H * 4); w.sample = (int*)take(H * (size_t)d->sample_cap * 4);
    w.possion_sample = (float*)take(n * 12); w.selected = (int*)take(n * 4);
    w.accept = (int*)take((V + 1) * 4); w.accept_scan = (int*)take((V + 1) * 4); w.cand = (int*)take(V * 4);
    w.foreign = (unsigned char*)take(V); w.fin_list = (int*)take(2 * V * 4);
    w.counters = (int*)take(64);
    size_t t1 = 0, t2 = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, t1, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int*)nullptr, (int)(n > 0 ? n : 1), 0, 5);
    cub::DeviceScan::ExclusiveSum(nullptr, t2, (int*)nullptr, (int*)nullptr, (int)(V + 1));
    w.cub_bytes = bal(t1 > t2 ? t1 : t2); w.cub = take(w.cub_bytes);
    w.total = off;
    return w;
}
static int bd_check(const wcsph_bd_desc* d, const void* work, size_t bytes, BdWork* w, const char* fn) {
    if (!d || !work) { wcsph_set_error("%s: null argument", fn); return WCSPH_EINVAL; }
    if (d->n < 1 || d->padding < d->n || (d->padding & (d->padding - 1)) || d->hash_size < 1 || d->sample_cap < 1 || d->sample_cap > 16 ||
        !(d->gridR > 0.f) || !(d->radius > 0.f)) { wcsph_set_error("%s: bad descriptor (n %d, padding %d, hash %d)", fn, d->n, d->padding, d->hash_size); return WCSPH_EINVAL; }
    *w = bd_carve((char*)work, d);
    if (bytes < w->total) { wcsph_set_error("%s: workspace %zu < %zu bytes", fn, bytes, w->total); return WCSPH_EINVAL; }
    return 0;
}
extern "C" size_t wcsph_bd_workspace_bytes(const wcsph_bd_desc* d) {
    if (!d || d->n < 1 || d->padding < d->n || d->hash_size < 1 || d->sample_cap < 1) return 0;
    return bd_carve(nullptr, d).total;
}

// ---- shared helpers ---------------------------------------------------------------------------------------------------------------
__host__ __device__ __forceinline__ int bd_hash(int x, int y, int z, int size) {        // :284-290 + the floor-mod of :255
    const int p1 = (int)(73856093u * (unsigned)x), p2 = (int)(19349663u * (unsigned)y), p3 = (int)(83492791u * (unsigned)z);
    int m = (p1 ^ p2 ^ p3) % size;
    if (m < 0) m += size;
    return m;
}
__device__ __forceinline__ int bd_cmp(int4 a, int4 b) {                                   // compare_cell :293-305
    if (a.x != b.x) return a.x > b.x ? 1 : -1;
    if (a.y != b.y) return a.y > b.y ? 1 : -1;
    if (a.z != b.z) return a.z > b.z ? 1 : -1;
    return 0;
}

// ---- init_point_set (:223-247) ----------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t bd_mix(uint32_t x) { x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16; return x; }
__device__ __forceinline__ float bd_rand(uint32_t seed, uint32_t i, uint32_t& ctr) {
    const uint32_t h = bd_mix(bd_mix(seed ^ (i * 0x9e3779b9u)) + (ctr++) * 0x85ebca6bu);
    return (float)(h >> 8) * (1.0f / 16777216.0f);                                        // [0, 1)
}
__global__ void k_bd_init_points(const float* __restrict__ tri_v, const float* __restrict__ tri_area, int face_num, float max_area,
                                 uint32_t seed, wcsph_bd_desc d, float4* __restrict__ pos) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.padding) return;
    if (i >= d.n) { pos[i] = make_float4(0.f, 0.f, 0.f, __int_as_float(0)); return; }
    uint32_t ctr = 0;
    const float rn1 = sqrtf(bd_rand(seed, i, ctr));
    const float bc1 = 1.0f - rn1, bc2 = bd_rand(seed, i, ctr) * rn1, bc3 = 1.0f - bc1 - bc2;
    int f = 0;
    for (int tries = 0; tries < 4096; tries++) {                                          // `while 1` of :232-235, bounded
        f = min((int)((float)face_num * bd_rand(seed, i, ctr)), face_num - 1);
        if (bd_rand(seed, i, ctr) < tri_area[f] / max_area) break;
    }
    const float* a = tri_v + 9 * (size_t)f;
    pos[i] = make_float4(bc1 * a[0] + bc2 * a[3] + bc3 * a[6], bc1 * a[1] + bc2 * a[4] + bc3 * a[7], bc1 * a[2] + bc2 * a[5] + bc3 * a[8],
                         __int_as_float(f));
}
// init_cell = cast((init_pos - min_point) / gridR, i32) + 1 (:245); padding entries sort last (:247)
__global__ void k_bd_cells(const float4* __restrict__ pos, wcsph_bd_desc d, int4* __restrict__ cell) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.padding) return;
    if (i >= d.n) { cell[i] = make_int4(1000000, 1000000, 1000000, 0); return; }
    const float4 p = pos[i];
    cell[i] = make_int4((int)__fdiv_rn(__fsub_rn(p.x, d.min_point[0]), d.gridR) + 1, (int)__fdiv_rn(__fsub_rn(p.y, d.min_point[1]), d.gridR) + 1,
                        (int)__fdiv_rn(__fsub_rn(p.z, d.min_point[2]), d.gridR) + 1, 0);
}

// ---- gpu_bitonic_sort (:208-219) / gpu_merge (:322-336) ----------------------------------------------------------------------------
// one compare-exchange step of the network on global memory
__global__ void k_bd_merge(int4* __restrict__ cell, float4* __restrict__ pos, int P, int j, int k) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= P) return;
    const int ixj = i ^ j;
    if (ixj <= i) return;
    const int4 a = cell[i], b = cell[ixj];
    const int c = bd_cmp(a, b);
    if (((i & k) == 0) ? (c == 1) : (c == -1)) {
        cell[i] = b; cell[ixj] = a;
        const float4 t = pos[i]; pos[i] = pos[ixj]; pos[ixj] = t;
    }
}
// all steps j = j0, j0/2, ..., 1 of stage k whose partners stay inside a tile of 2 * BD_TILE elements, in shared memory
#define BD_TILE 512
__global__ void __launch_bounds__(BD_TILE)
k_bd_merge_tile(int4* __restrict__ cell, float4* __restrict__ pos, int P, int j0, int k) {
    __shared__ int4 sc[2 * BD_TILE];
    __shared__ float4 sp[2 * BD_TILE];
    const int base = blockIdx.x * 2 * BD_TILE;
    for (int t = threadIdx.x; t < 2 * BD_TILE; t += BD_TILE) { sc[t] = cell[base + t]; sp[t] = pos[base + t]; }
    __syncthreads();
    for (int j = j0; j > 0; j >>= 1) {
        // thread t handles the pair (lo, lo ^ j) with lo = the t-th index whose bit j is clear
        const int t = threadIdx.x;
        const int lo = ((t & ~(j - 1)) << 1) | (t & (j - 1));
        const int hi = lo | j;
        const int gi = base + lo;
        const int4 a = sc[lo], b = sc[hi];
        const int c = bd_cmp(a, b);
        if (((gi & k) == 0) ? (c == 1) : (c == -1)) {
            sc[lo] = b; sc[hi] = a;
            const float4 tp = sp[lo]; sp[lo] = sp[hi]; sp[hi] = tp;
        }
        __syncthreads();
    }
    for (int t = threadIdx.x; t < 2 * BD_TILE; t += BD_TILE) { cell[base + t] = sc[t]; pos[base + t] = sp[t]; }
}

// ---- build_hmap (:250-271) ----------------------------------------------------------------------------------------------------------
__global__ void k_bd_heads(const int4* __restrict__ cell, wcsph_bd_desc d, int* __restrict__ start_index, int* __restrict__ hash_trace,
                           int* __restrict__ phase_key, int* __restrict__ head_idx, int* __restrict__ pg_count) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n) return;
    const int4 c = cell[i];
    head_idx[i] = i;
    bool head = i == 0;
    if (!head) { const int4 p = cell[i - 1]; head = (p.x != c.x) | (p.y != c.y) | (p.z != c.z); }
    int key = 27, tr = 0;
    if (head) {
        const int h = bd_hash(c.x, c.y, c.z, d.hash_size);
        atomicMax(&start_index[h], i);                       // serial order: the later head keeps the slot
        tr = h;
        key = c.x % 3 + 3 * (c.y % 3) + 9 * (c.z % 3);
        atomicAdd(&pg_count[key], 1);
        atomicAdd(&pg_count[27], 1);
    }
    hash_trace[i] = tr;
    phase_key[i] = key;
}
__global__ void k_bd_hcell(const int4* __restrict__ cell, wcsph_bd_desc d, const int* __restrict__ start_index, const int* __restrict__ phase_key,
                           int4* __restrict__ hcell) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= d.n || phase_key[i] == 27) return;
    const int4 c = cell[i];
    const int h = bd_hash(c.x, c.y, c.z, d.hash_size);
    if (start_index[h] == i) hcell[h] = make_int4(c.x, c.y, c.z, 0);
}
__global__ void k_bd_phase_offsets(const int* __restrict__ pg_count, int* __restrict__ off, int* __restrict__ counters, int V) {
    if (threadIdx.x || blockIdx.x) return;
    int s = 0, over = 0;
    for (int p = 0; p < 27; p++) { off[p] = s; s += pg_count[p]; over |= pg_count[p] > V; }
    off[27] = s;
    counters[1] = pg_count[27];
    counters[2] = over;                                       // "longer phase_group is needed!" (:270)
}
// heads are stably sorted by phase: entry `rank` of phase group p is the rank-th head of that phase in ascending index
__global__ void k_bd_phase_groups(const int4* __restrict__ cell, wcsph_bd_desc d, const int* __restrict__ key_sorted, const int* __restrict__ idx_sorted,
                                  const int* __restrict__ off, int4* __restrict__ phase_group) {
    const int s = blockIdx.x * blockDim.x + threadIdx.x;
    if (s >= d.n) return;
    const int p = key_sorted[s];
    if (p >= 27) return;
    const int rank = s - off[p];
    if (rank < d.phase_vec_max) { const int4 c = cell[idx_sorted[s]]; phase_group[(size_t)p * d.phase_vec_max + rank] = make_int4(c.x, c.y, c.z, 0); }
}
// an entry whose hash slot already belongs to an EARLIER entry of the same phase group re-tests that entry's point: never accepted
__global__ void k_bd_dups(wcsph_bd_desc d, const int* __restrict__ pg_count, const int4* __restrict__ phase_group, unsigned char* __restrict__ dup) {
    const int p = blockIdx.y;
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    const int cnt = min(pg_count[p], d.phase_vec_max);
    if (t >= cnt) return;
    const int4 c = phase_group[(size_t)p * d.phase_vec_max + t];
    const int h = bd_hash(c.x, c.y, c.z, d.hash_size);
    // colliding cells are rare (expected heads^2 / (2 H)); the scan below only runs over the earlier entries of this phase group
    unsigned char dd = 0;
    for (int u = 0; u < t; u++) {
        const int4 e = phase_group[(size_t)p * d.phase_vec_max + u];
        if (bd_hash(e.x, e.y, e.z, d.hash_size) == h) { dd = 1; break; }
    }
    dup[(size_t)p * d.phase_vec_max + t] = dd;
}

// ---- possion_disk_sample (:390-407) --------------------------------------------------------------------------------------------------
__device__ __forceinline__ int bd_check_cell_distance(const wcsph_bd_desc& d, int nx, int ny, int nz, float4 cur, const float4* __restrict__ pos,
                                                      const float* __restrict__ tri_normal, const int* __restrict__ sample_count,
                                                      const int* __restrict__ sample) {
    const int h = bd_hash(nx, ny, nz, d.hash_size);
    const int cnt = sample_count[h];
    const int cid = __float_as_int(cur.w);
    int ret = 0;
    for (int k = 0; k < cnt && ret == 0; k++) {
        const float4 nb = pos[sample[(size_t)h * d.sample_cap + k]];
        const float dx = __fsub_rn(cur.x, nb.x), dy = __fsub_rn(cur.y, nb.y), dz = __fsub_rn(cur.z, nb.z);
        const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
        float dist = len;
        const int nid = __float_as_int(nb.w);
        if (cid != nid) {
            const float inv = __fdiv_rn(1.0f, len);
            const float vx = __fmul_rn(inv, dx), vy = __fmul_rn(inv, dy), vz = __fmul_rn(inv, dz);
            const float* n1 = tri_normal + 3 * (size_t)cid; const float* n2 = tri_normal + 3 * (size_t)nid;      // face id on a per-vertex array (sic)
            const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(n1[0], vx), __fmul_rn(n1[1], vy)), __fmul_rn(n1[2], vz));
            const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(n2[0], vx), __fmul_rn(n2[1], vy)), __fmul_rn(n2[2], vz));
            if (fabsf(__fsub_rn(c1, c2)) > 0.00001f) dist = __fmul_rn(dist, __fdiv_rn(__fsub_rn(asinf(c1), asinf(c2)), __fsub_rn(c1, c2)));
            else dist = __fdiv_rn(dist, __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(c1, c1))));
        }
        if (dist < d.radius) ret = 1;
    }
    return ret;
}
__global__ void k_bd_sample_eval(wcsph_bd_desc d, int pg, int trial, const int* __restrict__ pg_count, const int4* __restrict__ phase_group,
                                 const unsigned char* __restrict__ dup, const int4* __restrict__ cell, const float4* __restrict__ pos,
                                 const float* __restrict__ tri_normal, const int* __restrict__ start_index, const int* __restrict__ sample_count,
                                 const int* __restrict__ sample, int* __restrict__ accept, int* __restrict__ cand_out,
                                 unsigned char* __restrict__ foreign, int* __restrict__ counters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t > d.phase_vec_max) return;
    const int cnt = min(pg_count[pg], d.phase_vec_max);
    int acc = 0, cand = -1, fgn = 0;
    if (t < cnt && !dup[(size_t)pg * d.phase_vec_max + t]) {
        const int4 c = phase_group[(size_t)pg * d.phase_vec_max + t];
        const int h = bd_hash(c.x, c.y, c.z, d.hash_size);
        const int s0 = start_index[h];
        cand = s0 + trial;
        if (cand < d.n) {
            const int4 cc = cell[cand], c0 = cell[s0];
            if (cc.x == c0.x && cc.y == c0.y && cc.z == c0.z) {
                fgn = (c0.x != c.x) | (c0.y != c.y) | (c0.z != c.z);
                const float4 cur = pos[cand];
                int conflicts = 0;                                                          // check_cell :376-386
                for (int a = -2; a < 3; a++) for (int b = -2; b < 3; b++) for (int k = -2; k < 3; k++)
                    conflicts += bd_check_cell_distance(d, a + cc.x, b + cc.y, k + cc.z, cur, pos, tri_normal, sample_count, sample);
                acc = conflicts == 0;
            }
        }
    }
    if (t < d.phase_vec_max) { cand_out[t] = cand; foreign[t] = (unsigned char)(fgn && acc); }
    accept[t] = t < d.phase_vec_max ? acc : 0;
    if (fgn && acc) atomicOr(&counters[4], 1);          // this launch holds an accepted foreign entry: the serial fix-up has work
}

// pairwise form of check_cell_distance for two candidates of the SAME launch: would `cur` (visiting its 125 neighbour cells) have
// met the sample `a` (stored in hash slot ha) and found it too close?
__device__ int bd_pair_conflict(const wcsph_bd_desc& d, int4 ccell, float4 cur, int ha, float4 nb, const float* __restrict__ tri_normal) {
    bool visible = false;
    for (int a = -2; a < 3 && !visible; a++) for (int b = -2; b < 3 && !visible; b++) for (int k = -2; k < 3; k++)
        if (bd_hash(a + ccell.x, b + ccell.y, k + ccell.z, d.hash_size) == ha) { visible = true; break; }
    if (!visible) return 0;
    const float dx = __fsub_rn(cur.x, nb.x), dy = __fsub_rn(cur.y, nb.y), dz = __fsub_rn(cur.z, nb.z);
    const float len = __fsqrt_rn(__fadd_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)), __fmul_rn(dz, dz)));
    float dist = len;
    const int cid = __float_as_int(cur.w), nid = __float_as_int(nb.w);
    if (cid != nid) {
        const float inv = __fdiv_rn(1.0f, len);
        const float vx = __fmul_rn(inv, dx), vy = __fmul_rn(inv, dy), vz = __fmul_rn(inv, dz);
        const float* n1 = tri_normal + 3 * (size_t)cid; const float* n2 = tri_normal + 3 * (size_t)nid;
        const float c1 = __fadd_rn(__fadd_rn(__fmul_rn(n1[0], vx), __fmul_rn(n1[1], vy)), __fmul_rn(n1[2], vz));
        const float c2 = __fadd_rn(__fadd_rn(__fmul_rn(n2[0], vx), __fmul_rn(n2[1], vy)), __fmul_rn(n2[2], vz));
        if (fabsf(__fsub_rn(c1, c2)) > 0.00001f) dist = __fmul_rn(dist, __fdiv_rn(__fsub_rn(asinf(c1), asinf(c2)), __fsub_rn(c1, c2)));
        else dist = __fdiv_rn(dist, __fsqrt_rn(__fsub_rn(1.0f, __fmul_rn(c1, c1))));
    }
    return dist < d.radius;
}
// Serial fix-up of a launch that accepted a FOREIGN entry (its hash slot was taken over by a colliding cell, so its candidate lives
// in that other cell and is NOT guaranteed to be 3 cells away from the launch's other candidates).  Such an acceptance can conflict
// with another acceptance of the same launch; the reference's serial order lets the earlier one win.  One thread replays the
// accepted entries in order: a foreign entry is checked against every earlier survivor, a normal one against the earlier foreign
// survivors.  Launches without an accepted foreign entry (almost all) return at once.
__global__ void k_bd_sample_fix(wcsph_bd_desc d, int pg, const int* __restrict__ pg_count, const int4* __restrict__ phase_group,
                                const int4* __restrict__ cell, const float4* __restrict__ pos, const float* __restrict__ tri_normal,
                                int* __restrict__ accept, const int* __restrict__ cand, const unsigned char* __restrict__ foreign,
                                int* __restrict__ fin_list, int* __restrict__ counters) {
    if (threadIdx.x || blockIdx.x) return;
    if (!counters[4]) return;
    counters[4] = 0;
    const int cnt = min(pg_count[pg], d.phase_vec_max);
    int* all = fin_list; int* fgn = fin_list + d.phase_vec_max;
    int nall = 0, nfgn = 0;
    for (int t = 0; t < cnt; t++) {
        if (!accept[t]) continue;
        const float4 cur = pos[cand[t]];
        const int4 cc = cell[cand[t]];
        const int* lst = foreign[t] ? all : fgn;
        const int nl = foreign[t] ? nall : nfgn;
        bool rej = false;
        for (int q = 0; q < nl && !rej; q++) {
            const int u = lst[q];
            const int4 ce = phase_group[(size_t)pg * d.phase_vec_max + u];
            rej = bd_pair_conflict(d, cc, cur, bd_hash(ce.x, ce.y, ce.z, d.hash_size), pos[cand[u]], tri_normal) != 0;
        }
        if (rej) { accept[t] = 0; continue; }
        all[nall++] = t;
        if (foreign[t]) fgn[nfgn++] = t;
    }
}
__global__ void k_bd_sample_apply(wcsph_bd_desc d, int pg, const int4* __restrict__ phase_group, const float4* __restrict__ pos, const int* __restrict__ accept,
                                  const int* __restrict__ scan, const int* __restrict__ cand, int* __restrict__ sample_count, int* __restrict__ sample,
                                  float* __restrict__ possion_sample, int* __restrict__ selected, int* __restrict__ counters) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= d.phase_vec_max || !accept[t]) return;
    const int4 c = phase_group[(size_t)pg * d.phase_vec_max + t];
    const int h = bd_hash(c.x, c.y, c.z, d.hash_size);
    const int old = sample_count[h];                           // one accepted entry per slot and launch (duplicates are skipped)
    if (old < d.sample_cap) { sample[(size_t)h * d.sample_cap + old] = cand[t]; sample_count[h] = old + 1; }
    else { sample_count[h] = d.sample_cap - 1; atomicAdd(&counters[3], 1); }              // :400-402 "exceed hash sample!"
    const int dst = counters[0] + scan[t];
    const float4 p = pos[cand[t]];
    possion_sample[3 * (size_t)dst] = p.x; possion_sample[3 * (size_t)dst + 1] = p.y; possion_sample[3 * (size_t)dst + 2] = p.z;
    selected[dst] = cand[t];
}
__global__ void k_bd_bump(int* counters, const int* scan, int V) { if (!threadIdx.x && !blockIdx.x) counters[0] += scan[V]; }

// ---- entry points ---------------------------------------------------------------------------------------------------------------------
extern "C" int wcsph_bd_init_point_set(const wcsph_bd_desc* d, void* work, size_t bytes, const float* tri_vertices_dev, const float* tri_area_dev,
                                       int face_num, float max_area, unsigned int seed, void* stream) {
    BdWork w; int r = bd_check(d, work, bytes, &w, __func__); if (r) return r;
    if (!tri_vertices_dev || !tri_area_dev || face_num < 1 || !(max_area > 0.f)) { wcsph_set_error("%s: bad mesh", __func__); return WCSPH_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    k_bd_init_points<<<bd_blocks(d->padding), BD_BLOCK, 0, st>>>(tri_vertices_dev, tri_area_dev, face_num, max_area, seed, *d, w.pos);
    k_bd_cells<<<bd_blocks(d->padding), BD_BLOCK, 0, st>>>(w.pos, *d, w.cell);
    BD_TRY(cudaGetLastError());
    return 0;
}
// injected initial point set: host init_pos [n][3] f32, init_id [n] i32
extern "C" int wcsph_bd_set_points(const wcsph_bd_desc* d, void* work, size_t bytes, const float* host_pos, const int* host_id, void* stream) {
    BdWork w; int r = bd_check(d, work, bytes, &w, __func__); if (r) return r;
    if (!host_pos || !host_id) { wcsph_set_error("%s: null points", __func__); return WCSPH_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    float4* tmp = (float4*)malloc((size_t)d->padding * 16);
    if (!tmp) return WCSPH_ENOMEM;
    for (int i = 0; i < d->padding; i++) {
        if (i < d->n) { tmp[i].x = host_pos[3 * i]; tmp[i].y = host_pos[3 * i + 1]; tmp[i].z = host_pos[3 * i + 2]; memcpy(&tmp[i].w, &host_id[i], 4); }
        else { tmp[i].x = tmp[i].y = tmp[i].z = tmp[i].w = 0.f; }
    }
    cudaError_t e = cudaMemcpyAsync(w.pos, tmp, (size_t)d->padding * 16, cudaMemcpyHostToDevice, st);
    if (e == cudaSuccess) e = cudaStreamSynchronize(st);
    free(tmp);
    BD_TRY(e);
    k_bd_cells<<<bd_blocks(d->padding), BD_BLOCK, 0, st>>>(w.pos, *d, w.cell);
    BD_TRY(cudaGetLastError());
    return 0;
}

extern "C" int wcsph_bd_bitonic_sort(const wcsph_bd_desc* d, void* work, size_t bytes, void* stream) {
    BdWork w; int r = bd_check(d, work, bytes, &w, __func__); if (r) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const int P = d->padding;
    for (int k = 2; k <= P; k <<= 1) {
        int j = k >> 1;
        if (P >= 2 * BD_TILE) {
            // partners i ^ j with j > BD_TILE leave the 1024-element tile: one global compare-exchange pass each
            for (; j > BD_TILE; j >>= 1) k_bd_merge<<<bd_blocks(P), BD_BLOCK, 0, st>>>(w.cell, w.pos, P, j, k);
            // the remaining passes j, j/2, ..., 1 of this stage stay inside a tile: one launch, shared memory
            k_bd_merge_tile<<<P / (2 * BD_TILE), BD_TILE, 0, st>>>(w.cell, w.pos, P, j, k);
        } else {
            for (; j > 0; j >>= 1) k_bd_merge<<<bd_blocks(P), BD_BLOCK, 0, st>>>(w.cell, w.pos, P, j, k);
        }
    }
    BD_TRY(cudaGetLastError());
    return 0;
}

extern "C" int wcsph_bd_build_hmap(const wcsph_bd_desc* d, void* work, size_t bytes, void* stream) {
    BdWork w; int r = bd_check(d, work, bytes, &w, __func__); if (r) return r;
    cudaStream_t st = (cudaStream_t)stream;
    const size_t H = d->hash_size, V = d->phase_vec_max > 0 ? d->phase_vec_max : 1;
    BD_TRY(cudaMemsetAsync(w.start_index, 0, H * 4, st));
    BD_TRY(cudaMemsetAsync(w.hcell, 0, H * 16, st));
    BD_TRY(cudaMemsetAsync(w.phase_group_count, 0, 32 * 4, st));
    BD_TRY(cudaMemsetAsync(w.phase_group, 0, 27 * V * 16, st));
    BD_TRY(cudaMemsetAsync(w.sample_count, 0, H * 4, st));
    BD_TRY(cudaMemsetAsync(w.sample, 0, H * (size_t)d->sample_cap * 4, st));
    BD_TRY(cudaMemsetAsync(w.counters, 0, 64, st));
    k_bd_heads<<<bd_blocks(d->n), BD_BLOCK, 0, st>>>(w.cell, *d, w.start_index, w.hash_trace, w.phase_key, w.head_idx, w.phase_group_count);
    k_bd_hcell<<<bd_blocks(d->n), BD_BLOCK, 0, st>>>(w.cell, *d, w.start_index, w.phase_key, w.hcell);
    size_t tb = w.cub_bytes;
    BD_TRY(cub::DeviceRadixSort::SortPairs(w.cub, tb, w.phase_key, w.phase_key_sorted, w.head_idx, w.head_idx_sorted, d->n, 0, 5, st));
    k_bd_phase_offsets<<<1, 1, 0, st>>>(w.phase_group_count, w.phase_offset, w.counters, d->phase_vec_max);
    k_bd_phase_groups<<<bd_blocks(d->n), BD_BLOCK, 0, st>>>(w.cell, *d, w.phase_key_sorted, w.head_idx_sorted, w.phase_offset, w.phase_group);
    dim3 g(bd_blocks(V), 27);
    k_bd_dups<<<g, BD_BLOCK, 0, st>>>(*d, w.phase_group_count, w.phase_group, w.dup);
    BD_TRY(cudaGetLastError());
    return 0;
}

// one launch of possion_disk_sample(pg, trial, phase_group_count[pg]) (:390); no host synchronisation
extern "C" int wcsph_bd_sample(const wcsph_bd_desc* d, void* work, size_t bytes, const float* tri_normal_dev, int pg, int trial, void* stream) {
    BdWork w; int r = bd_check(d, work, bytes, &w, __func__); if (r) return r;
    if (!tri_normal_dev || pg < 0 || pg >= 27 || trial < 0) { wcsph_set_error("%s: bad argument", __func__); return WCSPH_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    const int V = d->phase_vec_max;
    if (V < 1) return 0;
    k_bd_sample_eval<<<bd_blocks(V + 1), BD_BLOCK, 0, st>>>(*d, pg, trial, w.phase_group_count, w.phase_group, w.dup, w.cell, w.pos, tri_normal_dev,
                                                         w.start_index, w.sample_count, w.sample, w.accept, w.cand, w.foreign, w.counters);
    k_bd_sample_fix<<<1, 32, 0, st>>>(*d, pg, w.phase_group_count, w.phase_group, w.cell, w.pos, tri_normal_dev, w.accept, w.cand, w.foreign,
                                      w.fin_list, w.counters);
    size_t tb = w.cub_bytes;
    BD_TRY(cub::DeviceScan::ExclusiveSum(w.cub, tb, w.accept, w.accept_scan, V + 1, st));
    k_bd_sample_apply<<<bd_blocks(V), BD_BLOCK, 0, st>>>(*d, pg, w.phase_group, w.pos, w.accept, w.accept_scan, w.cand, w.sample_count, w.sample,
                                                      w.possion_sample, w.selected, w.counters);
    k_bd_bump<<<1, 1, 0, st>>>(w.counters, w.accept_scan, V);
    BD_TRY(cudaGetLastError());
    return 0;
}

// named device arrays -> host (tests, export): "cell" int4[P], "pos" float4[P] (w = face id bits), "start_index" [H], "hcell" int4[H],
// "hash_trace" [n], "phase_group_count" [27], "phase_group" int4[27][V], "sample_count" [H], "sample" [H][cap], "possion_sample" f32[n][3],
// "selected" [n], "counters" [4] = samples, occupied hash slots, phase-group overflow, hash-sample overflows
extern "C" int wcsph_bd_get(const wcsph_bd_desc* d, void* work, size_t bytes, const char* name, void* host_dst, size_t dst_bytes, void* stream) {
    BdWork w; int r = bd_check(d, work, bytes, &w, __func__); if (r) return r;
    if (!name || !host_dst) return WCSPH_EINVAL;
    const size_t P = d->padding, H = d->hash_size, n = d->n, V = d->phase_vec_max > 0 ? d->phase_vec_max : 1;
    const void* src = nullptr; size_t nb = 0;
    if (!strcmp(name, "cell")) { src = w.cell; nb = P * 16; }
    else if (!strcmp(name, "pos")) { src = w.pos; nb = P * 16; }
    else if (!strcmp(name, "start_index")) { src = w.start_index; nb = H * 4; }
    else if (!strcmp(name, "hcell")) { src = w.hcell; nb = H * 16; }
    else if (!strcmp(name, "hash_trace")) { src = w.hash_trace; nb = n * 4; }
    else if (!strcmp(name, "phase_group_count")) { src = w.phase_group_count; nb = 27 * 4; }
    else if (!strcmp(name, "phase_group")) { src = w.phase_group; nb = 27 * V * 16; }
    else if (!strcmp(name, "sample_count")) { src = w.sample_count; nb = H * 4; }
    else if (!strcmp(name, "sample")) { src = w.sample; nb = H * (size_t)d->sample_cap * 4; }
    else if (!strcmp(name, "possion_sample")) { src = w.possion_sample; nb = n * 12; }
    else if (!strcmp(name, "selected")) { src = w.selected; nb = n * 4; }
    else if (!strcmp(name, "counters")) { src = w.counters; nb = 16; }
    else { wcsph_set_error("%s: unknown array '%s'", __func__, name); return WCSPH_ENAME; }
    if (dst_bytes < nb) { wcsph_set_error("%s: buffer too small for '%s': %zu < %zu", __func__, name, dst_bytes, nb); return WCSPH_EINVAL; }
    cudaStream_t st = (cudaStream_t)stream;
    BD_TRY(cudaMemcpyAsync(host_dst, src, nb, cudaMemcpyDeviceToHost, st));
    BD_TRY(cudaStreamSynchronize(st));
    return 0;
}
