// engine.cuh -- shared context, arena layout and device helpers of libwcsph_b200.
//
// Data layout in HBM:
//  * liquids live cell-sorted in slots [0,NL); solids, sorted once, in [NL,N) of `pos`.
//  * vec3 fields are float4 (one LDG.128 per gather); 3x3 is three float4 rows.
//  * fields that survive a step ("persistent") are double buffered and permuted by
//    update_grid; everything else is recomputed before use and single buffered.
//  * neighbours: compact in-range lists, warp-interleaved: entry k of sorted particle i is
//    nbr[(i/32*cap + k)*32 + i%32], so a warp reads 128 contiguous bytes per k.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <vector>
#include <string>
#include <map>
#include "../../include/wcsph_b200.h"

#define WCSPH_MAX_FIELDS 40
#ifndef WCSPH_BLOCK
#define WCSPH_BLOCK 256
#endif
#ifndef WCSPH_MINB
#define WCSPH_MINB 1       // min resident CTAs per SM asked of the sweep kernels (register cap = 65536 / (256 * MINB))
// Kernels that would take 72-96 registers carry their own figure, measured per kernel at 1M (tools/sweep_bench.py A/B builds):
// 72 -> 64 registers (4 CTAs) k_visc_Ad -9 %; 89 -> 80 (3 CTAs) k_visc_minv_residual -17 %, 96 -> 80 k_vorticity_fused -5 % (64 is worse
// for both); the 55-64 register sweeps (k_dfsph_drho, _velcorrect, _head) are SLOWER at 48 registers / 5 CTAs even without a spill.
#endif
#define WCSPH_ALIAS_CAP 65536

struct FieldSlot {
    const char* name;
    int ncomp;       // components the reference exposes (1, 3, 9)
    int stride;      // floats per element on the device (1, 4, 12)
    int n;           // element count (NL, or N for pos)
    int persistent;  // double buffered + permuted by update_grid
    int is_int;
    void* buf[2];
};

// scalars block (device): one float/int per named 1-element field + loop control
struct Scalars {
    float deltaT;
    float avg_density_err;
    float cg_delta, cg_delta_old, cg_delta_zero;
    float rho_err;
    float cg_dAd;
    float vel_max0;
    float dt_prev;              // deltaT before optimize_time_step (the omega update of dfsph.py:330 uses it)
    float red_tmp;              // raw total of a reduction between the ranks' all-reduce and its use
    unsigned int flags;
    int vs_iter, dv_iter, pr_iter;
    int loop_continue;          // device-evaluated predicate of the host loops
    float err_threshold;        // dfsph.py:143 err (device copy)
    unsigned int ticket;        // last-block reduction ticket
    int n_inbox;
    int alias_count;
    unsigned int step_counter;  // steps completed by graph launches (index into the iteration log)
    int pad[4];
};

#define WCSPH_ITER_LOG 4096     // ring of (vs, dv, pr) per graph-launched step

struct GridDims {
    int bx, by, bz;      // HashGrid.blockSize
    int ncells;
    float minx, miny, minz;
    float inv;           // float(1.0 / gridR)  HashGrid.py:16
    float cell;          // float(gridR)
    int n_hash;          // particle_data.count: hash modulus HashGrid.py:114
};

// optional per-kernel CUDA-event timing (bench.py roofline pass); off on the timed path
struct ProfRec { const char* name; cudaEvent_t a, b; };
struct Profiler {
    int enabled = 0;
    std::vector<ProfRec> recs;
    std::vector<cudaEvent_t> pool;
    std::map<std::string, std::pair<double, long long>> acc;   // name -> (ms, launches)
};

#define WCSPH_MAX_RANKS 16
// one per rank, written by the peers: word = (epoch << 32) | payload bits, double-buffered on the epoch's parity (a rank can be at
// most one exchange ahead of a peer: it needs the peer's word of exchange e before it can start e + 1)
struct Mailbox {
    unsigned long long red[2][WCSPH_MAX_RANKS];   // scalar all-reduce: [parity][source rank]
    unsigned long long cnt[2][2];                 // neighbour counts: [parity][0 = from the lower neighbour, 1 = from the upper]
};
struct wcsph_ctx {
    wcsph_desc desc;
    wcsph_params prm;
    cudaStream_t stream;
    int N, NL, NS;               // GLOBAL counts (ParticleData.count / liquid_count / solid_count)
    // slot layout of the liquid arrays: [ghost_lo (right-aligned in [0,G)) | owned [i0, i0+nown) | ghost_hi]
    // single GPU: G = 0, i0 = 0, nown = NL = CL.  Solids start at SB in `pos`.
    int i0, nown, capOwn, G, CL, SB;
    // z-slab decomposition (mgpu.cu); R == 1: none of it is touched
    int R, rank, zlo, zhi;       // this rank owns cell layers z in [zlo, zhi)
    int n_inbox, n_glo, n_ghi, n_send_lo, n_send_hi;
    long long mig_total[4];      // cumulative migrants: sent to lower / upper neighbour, received from lower / upper
    int halo_group_depth;
    long long halo_exchanges;    // cumulative halo exchanges (grouped send/recv sets) issued by this rank
    void* comm;                  // ncclComm_t: main-stream collectives (counts, migration, scalar all-reduces)
    void* comm2;                 // duplicate communicator for everything issued on the side stream
    // peer mailboxes (mgpu.cu): the ranks' one-float all-reduces and the neighbours' count exchange go through 8-byte
    // (epoch, value) words stored straight into the peers' memory over NVLink instead of through NCCL launches
    struct Mailbox* mbox;        // mine (cudaMalloc'ed by the library: an IPC handle needs its own allocation)
    struct Mailbox** mbox_peers; // device array [R]: every rank's mailbox as mapped into this process (own entry = mbox)
    void* mbox_opened[WCSPH_MAX_RANKS];  // host copies of the opened peer mappings (cudaIpcCloseMemHandle)
    unsigned int red_epoch, cnt_epoch;
    int p2p_scalars;             // option / state: 1 when the mailboxes are open and in use
    // halo / sweep overlap: the halo runs on side_stream while the interior particles are swept
    cudaStream_t side_stream, main_saved;
    cudaEvent_t ev_main, ev_halo, ev_occ;
    int halo_overlap;                   // option: overlap the halo with the interior sweep (default on)
    int list_build_v1;                  // option: the round-1 list-build kernel (A/B; same lists)
    int cfl_true_max;                   // option: CFL maximum over every liquid particle instead of the reference's first-P subset (Q15)
    int sub_active, sub_off, sub_n;     // sub-range of the owned particles a sweep launch covers
    int sub_gap_at, sub_gap_len;        // hole inside that sub-range (the interior, when one launch covers both boundary strips)
    int part_off, sweep_parts;          // block-partial offset of that launch / total of the split sweep
    float *mig_send[2], *mig_recv[2];   // packed migration records (lower / upper neighbour), capM = G records each
    int mig_rec;                        // floats per record: 4 per vec field + 1 per scalar field + 1 (reference index)
    int* mg_counts;              // device int[16]: migration / halo counts exchanged with the z neighbours
    int* mg_counts_host;         // pinned mirror
    int nwarps;                  // ceil(capOwn/32)
    int capL, capS;
    float cull_r;                // in-range radius for the compact lists
    GridDims g;                  // the reference's hash grid (HashGrid.blockSize, cell = gridR): neighborCount, aliasing, in-box test
    int F;                       // refinement of the SEARCH grid the particles are sorted on: F = 2 when the hash cell is the full
    GridDims gs;                 // support (sesph / pcisph / iisph construct ParticleData(gridR), Q5), else 1 (gs == g)
    // arena
    char* arena; size_t arena_bytes; size_t arena_used;
    // fields
    FieldSlot fields[WCSPH_MAX_FIELDS]; int nfields;
    int cur;                     // which buffer of persistent fields is current
    // grid / sort tables
    int *keys, *keys_sorted, *perm, *iota;
    int *sorted_id[2];           // sorted slot -> reference liquid index
    int *inv_id;                 // reference liquid index -> sorted slot (rebuilt lazily)
    int  inv_id_valid;
    int *solid_sorted_id;        // solid slot (0..NS) -> reference index - NL
    int *cell_start_l, *cell_start_s;
    int *occ, *occ_solid;        // bucket occupancy (HashGrid.gridCount)
    unsigned short* occ_h;       // z-slab ranks: this rank's liquid share of occ as fp16 bits, all-reduced instead of the int table
    int *bucket_of_cell;         // static: get_cell_hash(cell)
    int *boxA, *boxB;            // separable 5x5x5 box sums of occ[bucket(cell)]
    unsigned char* m_self;       // static: #{o : bucket(c+o) == bucket(c)}
    unsigned char* solid_near;   // static: 1 if any solid particle lives within +-2 cells of the cell
    int *alias_pairs;            // static: near-alias cell pairs (2 ints each)
    int *nl_cnt, *ns_cnt, *neighborCount;
    uint32_t *nbr_l, *nbr_s;
    void* cub_temp; size_t cub_temp_bytes;
    float* partials;             // block partial sums (3 per block)
    Scalars* sc;                 // device
    Scalars* sc_host;            // pinned host mirror
    float* stage; size_t stage_bytes;   // device staging for field get/set (N*4 floats)
    int uploaded;
    int host_scalars_valid;      // sc_host holds the avg_density_err / deltaT the next fused step's first loop test reads (stream-ordered path)
    unsigned int seen_flags;            // device status bits the host has read but the caller has not acknowledged (wcsph_status)
    int vs_iter, dv_iter, pr_iter;      // host copies (host-driven loops)
    long long launches;
    Profiler* prof;
    // whole-step CUDA graphs (one per parity of the double buffers), loops as conditional WHILE nodes
    int use_graph;
    cudaStream_t cap_stream;            // second stream for capturing loop bodies
    cudaGraph_t step_graph[2]; cudaGraphExec_t step_exec[2]; int step_graph_valid[2];
    int g_fixed, g_div_body, g_vs_body, g_pr_body;   // launches per graph: fixed part / per loop iteration
    int* iter_log;                      // device ring [WCSPH_ITER_LOG][3]
    unsigned int log_read;              // steps whose log entry the host has consumed
    int graph_pending;                  // graph-launched steps whose counters the host has not read yet
};

static inline void prof_begin(wcsph_ctx* c, const char* name) {
    Profiler* p = c->prof;
    if (!p || !p->enabled) return;
    ProfRec r; r.name = name;
    for (cudaEvent_t* e : {&r.a, &r.b}) {
        if (!p->pool.empty()) { *e = p->pool.back(); p->pool.pop_back(); }
        else cudaEventCreate(e);
    }
    cudaEventRecord(r.a, c->stream);
    p->recs.push_back(r);
}
static inline void prof_end(wcsph_ctx* c) {
    Profiler* p = c->prof;
    if (!p || !p->enabled) return;
    cudaEventRecord(p->recs.back().b, c->stream);
}

// ---- error plumbing -------------------------------------------------------------------
void wcsph_set_error(const char* fmt, ...);
#define CUDA_TRY(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { \
    wcsph_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, cudaGetErrorString(e_)); return WCSPH_ECUDA; } } while (0)
#define PROF(ctx, name, stmt) do { prof_begin(ctx, name); stmt; prof_end(ctx); } while (0)
#define LAUNCH_CHECK(ctx) do { (ctx)->launches++; cudaError_t e_ = cudaGetLastError(); if (e_ != cudaSuccess) { \
    wcsph_set_error("%s:%d launch -> %s", __FILE__, __LINE__, cudaGetErrorString(e_)); return WCSPH_ECUDA; } } while (0)
#define TRY(x) do { int r_ = (x); if (r_ != 0) return r_; } while (0)

FieldSlot* wcsph_find_field(wcsph_ctx* c, const char* name);
int wcsph_finalize_reduce(wcsph_ctx* c, int nparts, int op, float eps);   // api.cu
int wcsph_drain_iter_log(wcsph_ctx* c);                                    // api.cu
int wcsph_fatal_flags(wcsph_ctx* c);                                       // api.cu: WCSPH_EOVERFLOW if pairs were dropped
void wcsph_invalidate_graphs(wcsph_ctx* c);                               // api.cu
struct CellStartArgs { int base, hi_cell0, n_oob, c_lo, c_hi, box_done; };
int wcsph_box_filter(wcsph_ctx* c);                                        // grid.cu
int wcsph_grid_finish(wcsph_ctx* c, CellStartArgs csa);                    // grid.cu
int wcsph_sort_permute(wcsph_ctx* c, int n, int kbase, int kspan);
int wcsph_p2p_allreduce_apply(wcsph_ctx* c, int op, float eps);                                  // mgpu.cu                               // grid.cu
int wcsph_halo(wcsph_ctx* c, const char* name);                            // mgpu.cu (no-op on one GPU)
int wcsph_allreduce_scalar(wcsph_ctx* c, float* dev, int is_max);          // mgpu.cu
#define HALO(c, name) do { if ((c)->R > 1) TRY(wcsph_halo(c, name)); } while (0)
int wcsph_halo_group(wcsph_ctx* c, int begin);   // mgpu.cu: fuse the halos issued in between into one NCCL group
int wcsph_halo_begin(wcsph_ctx* c);   // mgpu.cu: fork the halo onto the side stream
int wcsph_halo_end(wcsph_ctx* c);
int wcsph_halo_wait(wcsph_ctx* c);
template <class T> static inline T* fcur(wcsph_ctx* c, const char* name) {
    FieldSlot* f = wcsph_find_field(c, name);
    if (!f) return nullptr;
    return (T*)f->buf[f->persistent ? c->cur : 0];
}
// pointer to the first OWNED element of a field (streaming kernels index from 0)
template <class T> static inline T* fown(wcsph_ctx* c, const char* name) {
    FieldSlot* f = wcsph_find_field(c, name);
    if (!f) return nullptr;
    return (T*)((char*)f->buf[f->persistent ? c->cur : 0] + (size_t)c->i0 * f->stride * sizeof(float));
}
static inline int nblocks(int n, int b = WCSPH_BLOCK) { return n > 0 ? (n + b - 1) / b : 1; }

// ---- device helpers -------------------------------------------------------------------
struct KC {                     // kernel constants passed by value
    float h, inv_h, m_k, m_l, m_l_h; int style;
    float rho0, rhoS0, VL0, VS0, mass, eps;
};
static inline KC make_kc(const wcsph_params& p) {
    KC k; k.h = p.searchR; k.inv_h = (float)(1.0 / (double)p.searchR); k.m_k = p.m_k; k.m_l = p.m_l;
    k.m_l_h = (float)((double)p.m_l / (double)p.searchR);
    k.style = p.kernel_style; k.rho0 = p.rho_L0; k.rhoS0 = p.rho_S0; k.VL0 = p.VL0; k.VS0 = p.VS0;
    k.mass = p.liqiudMass; k.eps = p.eps; return k;
}

__device__ __forceinline__ float3 f3(float x, float y, float z) { return make_float3(x, y, z); }
__device__ __forceinline__ float3 xyz(float4 a) { return make_float3(a.x, a.y, a.z); }
__device__ __forceinline__ float4 f4(float3 a, float w = 0.f) { return make_float4(a.x, a.y, a.z, w); }
__device__ __forceinline__ float3 operator+(float3 a, float3 b) { return f3(a.x + b.x, a.y + b.y, a.z + b.z); }
__device__ __forceinline__ float3 operator-(float3 a, float3 b) { return f3(a.x - b.x, a.y - b.y, a.z - b.z); }
__device__ __forceinline__ float3 operator*(float3 a, float s) { return f3(a.x * s, a.y * s, a.z * s); }
__device__ __forceinline__ float3 operator/(float3 a, float s) { return f3(a.x / s, a.y / s, a.z / s); }
__device__ __forceinline__ void operator+=(float3& a, float3 b) { a.x += b.x; a.y += b.y; a.z += b.z; }
__device__ __forceinline__ float dot3(float3 a, float3 b) { return a.x * b.x + a.y * b.y + a.z * b.z; }
__device__ __forceinline__ float3 cross3(float3 a, float3 b) {
    return f3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x);
}

// CubicKernel.py:44-54 / :36-37 (style 0)  |  sesph.py:112-124 (style 1); rl = |r|
// q = rl * (1/h): W and gradW are continuous at q = 0.5 and vanish at q = 1, so the last-ulp
// difference to the reference's rl / h cannot flip a contribution by more than rounding noise.
// Branch-free form of the piecewise cubic: with t = sat(1 - q), u = sat(1 - 2q)
//   6q^3 - 6q^2 + 1 (q <= 1/2), 2(1-q)^3 (q <= 1), 0   ==  2 t^3 - u^3
//   q(3q - 2)       (q <= 1/2), -(1-q)^2 (q <= 1), 0   ==  u^2 - t^2
// (expand (1-2q)^2 - (1-q)^2 and 2(1-q)^3 - (1-2q)^3): two saturating adds replace three compares, two selects and
// both predicated polynomial arms -- 7 issue slots fewer per pair in sweeps that are issue-bound.
// sat(1 - q) and sat(1 - 2q) as ONE instruction each (FADD.SAT / FFMA.SAT); __saturatef(expr) compiles to expr + a second FADD.SAT
__device__ __forceinline__ float sat_1mq(float q) { float t; asm("sub.sat.ftz.f32 %0, %1, %2;" : "=f"(t) : "f"(1.0f), "f"(q)); return t; }
__device__ __forceinline__ float sat_1m2q(float q) { float u; asm("fma.rn.sat.ftz.f32 %0, %1, %2, %3;" : "=f"(u) : "f"(-2.0f), "f"(q), "f"(1.0f)); return u; }
__device__ __forceinline__ float cubic_W(const KC& k, float rl) {
    const float q = rl * k.inv_h;
    const float t = sat_1mq(q), u = sat_1m2q(q);
    return k.m_k * fmaf(2.0f * t, t * t, -(u * u * u));          // m_k = 8/(pi h^3) in both styles (folded on the host in float64)
}
// W from the squared distance (one MUFU.RSQ instead of sqrt)
__device__ __forceinline__ float cubic_W2(const KC& k, float r2) {
    return cubic_W(k, r2 * rsqrtf(fmaxf(r2, 1e-30f)));
}

// CubicKernel.py:21-32 | sesph.py:97-108: gradW = m_l * f(q) * r / (|r| h), 0 if |r| <= 1e-5 or q > 1.
// cubic_gradW_s returns the scalar s with gradW = s * r, so that a sweep that only needs gradW . x or
// gradW * c forms s * (r . x) / r * (s * c) and never materialises the vector (3 multiplies fewer per pair).
// cubic_gradW_u: the same without the constant, gradW = m_l_h * u * r -- sweeps whose sum is linear in gradW apply
// m_l_h once per particle instead of once per pair.
__device__ __forceinline__ float cubic_gradW_u(const KC& k, float r2) {
    // |r| <= 1e-5 -> 0 (CubicKernel.py:25), tested on r2 so that the same select also guards rsqrt(0): with
    // inv_rl = 0 the pair has q = 0, f = 1 - 1 = 0 and contributes exactly nothing (self-index padding included)
    const float inv_rl = r2 > 1.0e-10f ? rsqrtf(r2) : 0.0f;
    const float q = r2 * inv_rl * k.inv_h;
    const float t = sat_1mq(q), u = sat_1m2q(q);
    return fmaf(u, u, -(t * t)) * inv_rl;
}
__device__ __forceinline__ float cubic_gradW_s(const KC& k, float r2) { return cubic_gradW_u(k, r2) * k.m_l_h; }   // m_l_h = m_l / h
__device__ __forceinline__ float3 cubic_gradW(const KC& k, float3 r, float r2) { return r * cubic_gradW_s(k, r2); }

// HashGrid.py:109-114 -- i32 wrap-around products, floor-mod by particle count
__device__ __forceinline__ int cell_hash(int x, int y, int z, int n) {
    int p1 = (int)(73856093u * (unsigned)x);
    int p2 = (int)(19349663u * (unsigned)y);
    int p3 = (int)(83492791u * (unsigned)z);
    int m = (p1 ^ p2 ^ p3) % n;
    if (m < 0) m += n;
    return m;
}

// HashGrid.py:68 -- ti.cast((pos - min) * invGridR, i32): f32 sub, f32 mul (no FMA), trunc
__device__ __forceinline__ void cell_coords(const GridDims& g, float x, float y, float z, int& cx, int& cy, int& cz) {
    cx = (int)__fmul_rn(__fsub_rn(x, g.minx), g.inv);
    cy = (int)__fmul_rn(__fsub_rn(y, g.miny), g.inv);
    cz = (int)__fmul_rn(__fsub_rn(z, g.minz), g.inv);
}
__device__ __forceinline__ bool in_box(const GridDims& g, int cx, int cy, int cz) {
    return !(cx < 0 || cx >= g.bx || cy < 0 || cy >= g.by || cz < 0 || cz >= g.bz);
}

// Deterministic block reduction of up to 3 values -> partials[blockIdx*3+..]; the last block
// to finish (ticket) sums the partials in a fixed order and hands the totals to `fin`.
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Global reductions are two-phase and deterministic: every block leaves one partial per value
// (fixed shuffle / shared-memory tree), and a one-block finalize kernel launched right behind it
// sums the partials in a fixed order.  (A last-block-done scheme needs a gpu-scope fence per block,
// which on this part invalidates the SM's L1 and showed up as ~30 us on every reducing sweep.)
template <int NV, bool IS_MAX>
__device__ __forceinline__ void block_partials(float (&v)[NV], float* partials) {
    __shared__ float sm[NV][WCSPH_BLOCK / 32];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
    for (int a = 0; a < NV; a++) {
        float x = IS_MAX ? warp_max(v[a]) : warp_sum(v[a]);
        if (lane == 0) sm[a][w] = x;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
#pragma unroll
        for (int a = 0; a < NV; a++) {
            float x = sm[a][0];
            for (int i = 1; i < (int)(blockDim.x >> 5); i++) x = IS_MAX ? fmaxf(x, sm[a][i]) : x + sm[a][i];
            partials[(size_t)blockIdx.x * NV + a] = x;
        }
    }
}

enum FinOp { FIN_AVG_ERR = 0, FIN_CG_DELTA0, FIN_CG_DAD, FIN_CG_DELTA, FIN_VEL_MAX, FIN_RHO_ERR };
// what a finished global reduction writes into the scalar block (k_finalize, k_apply_fin, k_p2p_allreduce_apply)
__device__ __forceinline__ void apply_fin(Scalars* sc, int op, float eps, float t) {
    switch (op) {
        case FIN_AVG_ERR:   sc->avg_density_err = t; break;
        case FIN_CG_DELTA0: sc->cg_delta_zero = t; sc->cg_delta = t; break;
        case FIN_CG_DAD:    sc->cg_dAd = eps + t; break;
        case FIN_CG_DELTA:  sc->cg_delta_old = sc->cg_delta; sc->cg_delta = t; break;
        case FIN_VEL_MAX:   sc->vel_max0 = t; break;
        case FIN_RHO_ERR:   sc->rho_err += t; break;
    }
}

// ---- peer mailboxes (struct Mailbox): device side -----------------------------------------------------------------------------
// word = (epoch << 32) | payload: ONE naturally aligned 8-byte store into the peer's memory, so value and "it is there" arrive together
__device__ __forceinline__ void mb_store(unsigned long long* p, unsigned int epoch, unsigned int bits) {
    *(volatile unsigned long long*)p = ((unsigned long long)epoch << 32) | bits;
}
// spins until the word of `epoch` is there; ~30 s without it (6e10 SM cycles: far beyond any imbalance between ranks that execute the same
// step sequence, short enough that a dead peer does not hang the GPU) -> WCSPH_FLAG_COMM_TIMEOUT, fatal at the next check
__device__ __forceinline__ unsigned int mb_wait(const unsigned long long* p, unsigned int epoch, Scalars* sc) {
    const long long t0 = clock64();
    for (;;) {
        const unsigned long long w = *(const volatile unsigned long long*)p;
        if ((unsigned int)(w >> 32) == epoch) return (unsigned int)w;
        if (clock64() - t0 > 60000000000ll) { atomicOr(&sc->flags, WCSPH_FLAG_COMM_TIMEOUT); return 0u; }
        __nanosleep(20);
    }
}
// all-reduce of one float over the ranks' mailboxes, called by threads 0..R-1 of ONE block (the others pass through the barrier):
// every rank stores its value into every mailbox (its own included) and combines the R words it received IN RANK ORDER, so the
// result is bit-identical on every rank.  Returns the result in thread 0.
__device__ __forceinline__ float p2p_allreduce(float mine_val, bool is_max, Mailbox* mine, Mailbox* const* peers, int R, int rank,
                                               unsigned int epoch, Scalars* sc, float* got /* shared, >= R */) {
    const int t = threadIdx.x, par = epoch & 1;
    if (t < R) {
        mb_store(&peers[t]->red[par][rank], epoch, __float_as_uint(mine_val));
        got[t] = __uint_as_float(mb_wait(&mine->red[par][t], epoch, sc));
    }
    __syncthreads();
    float x = got[0];
    if (t == 0) for (int r = 1; r < R; r++) x = is_max ? fmaxf(x, got[r]) : x + got[r];
    return x;
}

// compact neighbour lists: entries k = 4*k4 .. 4*k4+3 of sorted particle i form ONE uint4 at
// ((uint4*)nbr)[((i/32)*(cap/4) + k4)*32 + i%32]: a warp reads 512 contiguous bytes per k4 and
// each lane gets four neighbour indices per LDG.128.
#define NBR_AT(ptr, cap, i, k) ((ptr)[(((size_t)((i) >> 5) * ((cap) >> 2) + ((k) >> 2)) * 32 + ((i) & 31)) * 4 + ((k) & 3)])
#define NBR_ROW4(ptr, cap, i) (((const uint4*)(ptr)) + ((size_t)((i) >> 5) * ((cap) >> 2)) * 32 + ((i) & 31))
