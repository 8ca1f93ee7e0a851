/*
 * wcsph_b200.h -- C ABI of the B200-native SPH per-step hot path.
 *
 * This is the drop-in boundary for the one hot path of lyd405121/wcsph
 * (SURVEY.md section 8): HashGrid.update_grid + every neighbour-sweep /
 * streaming @ti.kernel of sesph.py / pcisph.py / iisph.py / dfsph.py.
 * The reference has no FFI of its own (it is Taichi DSL); each entry point
 * below names the reference kernel it replaces (file:line under
 * /root/reference).  Plain pointers and sizes only -- no torch types.
 *
 * Memory model: the caller owns ONE device arena (wcsph_arena_bytes() bytes,
 * e.g. a torch.uint8 CUDA tensor); the library sub-allocates every table from
 * it and never calls cudaMalloc.  All work is issued on the caller's stream
 * (wcsph_set_stream) -- e.g. torch.cuda.current_stream().cuda_stream.
 *
 * Particle order: liquid particles are kept cell-sorted on the device and
 * re-sorted by every wcsph_hashgrid_update_grid(); solids are sorted once.
 * wcsph_field_get/_set translate to and from the reference's insertion order
 * (liquid indices [0,NL), solids [NL,N) -- dfsph.py:258).
 *
 * Every function returns 0 on success, a negative WCSPH_E* code otherwise;
 * wcsph_last_error() gives the text.  There is no CPU fallback.
 */
#ifndef WCSPH_B200_H
#define WCSPH_B200_H
#include <stddef.h>
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define WCSPH_ABI_VERSION 2

enum { WCSPH_SESPH = 0, WCSPH_PCISPH = 1, WCSPH_IISPH = 2, WCSPH_DFSPH = 3 };
enum { WCSPH_OK = 0, WCSPH_EINVAL = -1, WCSPH_ECUDA = -2, WCSPH_ENOMEM = -3, WCSPH_ENAME = -4,
       WCSPH_EOVERFLOW = -5 /* a capacity of the engine was exceeded on the device: results are incomplete (wcsph_check) */ };

/* device-side status bits (wcsph_status) */
#define WCSPH_FLAG_BUCKET_OVERFLOW   1u  /* HashGrid.py:72-74 "exceed grid": a 64-slot bucket overflowed   */
#define WCSPH_FLAG_NEIGHBOR_OVERFLOW 2u  /* HashGrid.py:101-103 "exceed neighbor": > maxNeighbour candidates */
#define WCSPH_FLAG_LIST_OVERFLOW     4u  /* compact in-range list stride exceeded (raise list_cap_*)        */
#define WCSPH_FLAG_ALIAS_OVERFLOW    8u  /* static alias-pair table exceeded                                */
#define WCSPH_FLAG_NAN               16u /* dfsph.py:645 NaN probe, evaluated on the device                 */
#define WCSPH_FLAG_MC_OVERFLOW       32u /* MarchingCubeGrid.py:173-175 "mc exceed grid": > maxInGrid liquids in a cell */
#define WCSPH_FLAG_MIGRATE_FAR       64u /* z-slab rank received a particle whose cell layer is outside its slab (moved > 1 slab) */
#define WCSPH_FLAG_COMM_TIMEOUT      128u /* a peer's mailbox word did not arrive within ~30 s (a rank died or left the step sequence) */
/* bits that mean "pairs were dropped": the step entry points and wcsph_check() turn them into WCSPH_EOVERFLOW */
#define WCSPH_FLAGS_FATAL (WCSPH_FLAG_BUCKET_OVERFLOW | WCSPH_FLAG_LIST_OVERFLOW | WCSPH_FLAG_ALIAS_OVERFLOW | WCSPH_FLAG_MIGRATE_FAR | WCSPH_FLAG_COMM_TIMEOUT)

/* Constants that the reference bakes into its kernels at JIT time; the host
 * evaluates them in float64 exactly like the reference's Python and narrows
 * once.  Cited per member. */
typedef struct wcsph_params {
    float searchR;        /* physics support h: hash_grid.searchR dfsph.py:76 | sesph.py:41          */
    float m_k;            /* W norm:   8/(pi h^3)   CubicKernel.py:15,37 | sesph.py:44               */
    float m_l;            /* gradW norm: 48/(pi h^3) CubicKernel.py:16,28 | sesph.py:45              */
    float m_k_raw, h3inv; /* CubicKernel.py:14-15: W = P(q)*m_k_raw*h3inv (dfsph evaluation order)   */
    int   kernel_style;   /* 0: CubicKernel class (dfsph)  1: script-inline W (sesph/pcisph/iisph)   */
    float coh_m_k, coh_m_c; /* CohesionKernel.py:15-16 */
    float adh_m_k;          /* AdhesionKernel.py:15    */
    float rho_L0, rho_S0, VL0, VS0, liqiudMass;      /* ParticleData.py:18-22 | sesph.py:35-38       */
    float gravity[3];                                 /* ParticleData.py:61                           */
    float dim_coff, viscosity, viscosity_b, viscosity_err; /* ParticleData.py:62-65                  */
    float tension_coff, tension_coff_b;               /* ParticleData.py:80-81                        */
    float viscosity_omega, vorticity_coff, vorticity_init; /* ParticleData.py:85-87                  */
    float stiffness;      /* sesph.py:58   */
    float pci_coff;       /* pcisph.py:87-115 */
    float omega_relax;    /* iisph.py:78   */
    float eps;            /* dfsph.py:23   */
    float particleRadius; /* dfsph.py:28   */
    float user_max_t, user_min_t; /* dfsph.py:40-41 */
} wcsph_params;

typedef struct wcsph_desc {
    int    abi_version;       /* WCSPH_ABI_VERSION */
    int    solver;            /* WCSPH_* : selects which solver-local fields exist */
    int    count;             /* ParticleData.count        (N)  */
    int    liquid_count;      /* ParticleData.liquid_count (NL) */
    double hash_gridR;        /* HashGrid(gridR, ...) HashGrid.py:10,17 */
    int    max_in_grid;       /* HashGrid maxInGrid    (64)   -- overflow flag only */
    int    max_neighbour;     /* HashGrid maxNeighbour (2048) -- overflow flag only */
    int    list_cap_liquid;   /* stride of the compact in-range liquid list (0 = default: 64, PCISPH 128) */
    int    list_cap_solid;    /* stride of the compact in-range solid list  (0 = default: 64, PCISPH 128) */
    float  cull_scale;        /* in-range test radius = cull_scale * searchR (0 = default: 1.0; PCISPH 1.25 because
                                 pcisph.py:266-268 evaluates gradW against PREDICTED positions) */
    float  min_boundary[3];   /* ParticleData.minboundarynp ParticleData.py:91-96 */
    float  max_boundary[3];
    /* z-slab decomposition over the GPUs of one box; world_size <= 1: single GPU.
     * count / liquid_count stay the GLOBAL numbers on every rank (the hash modulus is global). */
    int    world_size, rank;
    int    z_lo, z_hi;        /* this rank owns the cell layers z in [z_lo, z_hi) of HashGrid.blockSize.z */
    int    cap_own;           /* liquid slots for owned particles (0 = liquid_count) */
    int    cap_ghost;         /* ghost slots per side (2 cell layers of the neighbour slab)   */
    wcsph_params params;
} wcsph_desc;

typedef struct wcsph_ctx wcsph_ctx;

/* ---- lifetime -------------------------------------------------------- */
const char* wcsph_last_error(void);
int    wcsph_abi_version(void);
/* bytes of device arena the scene needs (ParticleData.setup_data_gpu ParticleData.py:142-177
 * + HashGrid.setup_grid_gpu HashGrid.py:34-40, minus the N x 64 / NL x 2048 tables) */
size_t wcsph_arena_bytes(const wcsph_desc* desc);
int    wcsph_create(const wcsph_desc* desc, void* device_arena, size_t arena_bytes,
                    void* cuda_stream, wcsph_ctx** out);
void   wcsph_destroy(wcsph_ctx* ctx);
int    wcsph_set_stream(wcsph_ctx* ctx, void* cuda_stream);
int    wcsph_set_params(wcsph_ctx* ctx, const wcsph_params* params);
/* ParticleData.setup_data_cpu ParticleData.py:180-185: host xyz (N x 3 f32, insertion order)
 * -> device; sorts the static solids, builds the static alias tables, sizes the grid
 * (HashGrid.setup_grid_cpu HashGrid.py:44-54). */
int    wcsph_upload_pos(wcsph_ctx* ctx, const float* host_pos_xyz);
int    wcsph_block_size(wcsph_ctx* ctx, int out_xyz[3]);              /* HashGrid.blockSize */

/* ---- Field API: .to_numpy() / .from_numpy() / field[0] (dfsph.py:98,113,129) ---- */
/* name = reference field name ("pos","vel","rho","alpha_coff","kappa", "neighborCount", ...).
 * Host buffers, reference insertion order, f32 (i32 for neighborCount); components per
 * element reported by wcsph_field_info.  Synchronises the stream. */
int    wcsph_field_info(wcsph_ctx* ctx, const char* name, int* count, int* ncomp, int* is_int);
int    wcsph_field_get(wcsph_ctx* ctx, const char* name, void* host_dst, size_t bytes);
int    wcsph_field_set(wcsph_ctx* ctx, const char* name, const void* host_src, size_t bytes);
/* asynchronous variants for pinned host buffers (the e2e path): no stream sync */
int    wcsph_field_get_async(wcsph_ctx* ctx, const char* name, void* pinned_dst, size_t bytes);
int    wcsph_field_set_async(wcsph_ctx* ctx, const char* name, const void* pinned_src, size_t bytes);
/* zero-copy device view in CURRENT SORTED order (stride in floats; 4 for vec3 fields) */
int    wcsph_field_device(wcsph_ctx* ctx, const char* name, void** dev_ptr, int* count, int* stride);
/* sorted slot k holds reference index sorted_id[k]  (device i32[NL]) */
int    wcsph_sorted_id_device(wcsph_ctx* ctx, void** dev_ptr);
/* 1-element fields: "deltaT","avg_density_err","cg_delta","cg_delta_old","cg_delta_zero","rho_err" */
int    wcsph_scalar_get(wcsph_ctx* ctx, const char* name, float* out);
int    wcsph_scalar_set(wcsph_ctx* ctx, const char* name, float v);
int    wcsph_status(wcsph_ctx* ctx, uint32_t* flags);       /* device status bits, cleared on read (acknowledges them) */
/* synchronises and returns WCSPH_EOVERFLOW (text in wcsph_last_error) if any WCSPH_FLAGS_FATAL bit has been raised since the
 * last wcsph_status(); the reference only prints in that case (HashGrid.py:73,103) and carries on with dropped entries.
 * The fused step entry points (wcsph_*_step), wcsph_iters and wcsph_field_get run the same test on the flags they have
 * already seen, so a run with missing pairs cannot go unnoticed. */
int    wcsph_check(wcsph_ctx* ctx);
int    wcsph_iters(wcsph_ctx* ctx, int out_vs_dv_pr[3]);    /* vs_iter, dv_iter, pr_iter of the last fused step */
int    wcsph_set_iters(wcsph_ctx* ctx, int vs, int dv, int pr);   /* restart: seed the counters dfsph.py:122 reads */
/* (vs, dv, pr) of the last max_steps fused steps, oldest first (the per-step console line dfsph.py:629) */
int    wcsph_iters_log(wcsph_ctx* ctx, int* out_3_per_step, int max_steps, int* n_out);
/* options: "graph" (default 1): run wcsph_dfsph_step as one CUDA graph per step with the host loops
 * of dfsph.py:93-99,141-145,160-163 as device-evaluated conditional WHILE nodes (no host round trip) */
int    wcsph_set_option(wcsph_ctx* ctx, const char* name, int value);
int    wcsph_sync(wcsph_ctx* ctx);
/* number of kernels this library launched since the last call with reset != 0 */
long long wcsph_launch_count(wcsph_ctx* ctx, int reset);

/* per-kernel CUDA-event timing (the reference has no profiler): enable != 0
 * starts recording an event pair around every launch, report drains "name\tlaunches\tms\n" */
int wcsph_profile(wcsph_ctx* ctx, int enable);
int wcsph_profile_report(wcsph_ctx* ctx, char* buf, size_t cap);

/* ---- multi-GPU (one process per GPU; the halo exchange is the only data-path collective) ----
 * rank 0 calls wcsph_comm_unique_id, the caller broadcasts the 128 bytes (torch.distributed), every
 * rank calls wcsph_comm_init.  nccl_path: libnccl.so.2 to dlopen (NULL: the copy torch already loaded). */
int wcsph_comm_unique_id(void* out_128_bytes, const char* nccl_path);
int wcsph_comm_init(wcsph_ctx* ctx, const void* unique_id_128_bytes, const char* nccl_path);
/* Optional, after wcsph_comm_init, ranks of ONE node (peer access over NVLink): peer mailboxes for the latency-bound exchanges
 * of a step -- the one-float all-reduce behind every global sum / maximum (fused into the finalize kernel: each rank stores its
 * partial into every peer's mailbox and adds the R words in rank order, so the result is the same on every rank) and the counts
 * the z neighbours tell each other before migration and halo.  wcsph_comm_mailbox_handle allocates this rank's mailbox (the
 * one cudaMalloc of the library: an IPC handle needs its own allocation) and writes its 64-byte cudaIpcMemHandle_t; the caller
 * all-gathers the handles (rank order) and every rank passes the R x 64 bytes to wcsph_comm_mailbox_open.  Without these calls
 * (or with option "p2p_scalars" = 0) the same exchanges run as NCCL calls. */
/* what NCCL reports for this context's communicator: out = {ncclCommCount, ncclCommUserRank, ncclGetVersion, 1 if the peer
 * mailboxes below are in use}; all 0 on a single-GPU context. */
int wcsph_comm_info(wcsph_ctx* ctx, int out[4]);
int wcsph_comm_mailbox_handle(wcsph_ctx* ctx, void* out_64_bytes);
int wcsph_comm_mailbox_open(wcsph_ctx* ctx, const void* handles_R_x_64_bytes);
int wcsph_owned_count(wcsph_ctx* ctx, int* n_owned, int* n_ghost_lo, int* n_ghost_hi);
/* cumulative since creation: particles migrated to the lower / upper z neighbour, received from the lower / upper one,
 * and [4] the number of halo exchanges this rank issued (observability of SURVEY 8e's exchange steps) */
int wcsph_migration_counts(wcsph_ctx* ctx, long long out_5[5]);

/* ---- HashGrid (HashGrid.py:57-106) ------------------------------------ */
int wcsph_hashgrid_update_grid(wcsph_ctx* ctx);
/* statistics of the compact in-range lists built by the last update_grid, summed over this rank's owned particles:
 * out[0] liquid pairs, out[1] solid pairs, out[2] / out[3] the longest liquid / solid list (vs list_cap_*) */
int wcsph_pair_counts(wcsph_ctx* ctx, long long out_4[4]);
/* lazy debug view of HashGrid.neighbor[i, 0:neighborCount[i]] restricted to in-range
 * candidates, as a multiset in reference indices: writes up to cap ints, returns count */
int wcsph_hashgrid_neighbors_of(wcsph_ctx* ctx, int ref_index, int* host_out, int cap, int* n_out);

/* ---- sesph.py:131-196 -------------------------------------------------- */
int wcsph_sesph_reset_param(wcsph_ctx* ctx);
int wcsph_sesph_update_advection_density(wcsph_ctx* ctx);
int wcsph_sesph_update_pressure(wcsph_ctx* ctx);
int wcsph_sesph_compute_force(wcsph_ctx* ctx);
int wcsph_sesph_integrator_sesph(wcsph_ctx* ctx);
int wcsph_sesph_step(wcsph_ctx* ctx, int nsteps);           /* sesph.py:220-225, fused */

/* ---- dfsph.py:168-580 --------------------------------------------------- */
int wcsph_dfsph_reset_param(wcsph_ctx* ctx);
int wcsph_dfsph_compute_density(wcsph_ctx* ctx);
int wcsph_dfsph_compute_dfsph_coff(wcsph_ctx* ctx);
int wcsph_dfsph_warmstart_divergence_vel(wcsph_ctx* ctx);
int wcsph_dfsph_begin_divergence_iter(wcsph_ctx* ctx);
int wcsph_dfsph_divergence_iter(wcsph_ctx* ctx);
int wcsph_dfsph_end_divergence_iter(wcsph_ctx* ctx);
int wcsph_dfsph_clear_nonpressure(wcsph_ctx* ctx);
int wcsph_dfsph_compute_tension(wcsph_ctx* ctx);
int wcsph_dfsph_init_viscosity_para(wcsph_ctx* ctx);
int wcsph_dfsph_compute_viscosity_force(wcsph_ctx* ctx);
int wcsph_dfsph_end_viscosity(wcsph_ctx* ctx);
int wcsph_dfsph_compute_vorticity(wcsph_ctx* ctx);
int wcsph_dfsph_cfl_max(wcsph_ctx* ctx);                    /* dfsph.py:107-111,556-568: vel_max[0] = true max (Q15) */
int wcsph_dfsph_update_vel(wcsph_ctx* ctx);
int wcsph_dfsph_warmstart_pressure(wcsph_ctx* ctx);
int wcsph_dfsph_begin_pressure_iter(wcsph_ctx* ctx);
int wcsph_dfsph_pressure_iter(wcsph_ctx* ctx);
int wcsph_dfsph_end_pressure_iter(wcsph_ctx* ctx);
int wcsph_dfsph_update_pos(wcsph_ctx* ctx);
/* dfsph.py:606-617 whole step(s) with the host loops (dfsph.py:84-164) evaluated on the
 * device: same iteration counts, no host round trip per iteration */
int wcsph_dfsph_step(wcsph_ctx* ctx, int nsteps);

/* ---- iisph.py:178-396 --------------------------------------------------- */
int wcsph_iisph_reset_param(wcsph_ctx* ctx);
int wcsph_iisph_compute_density(wcsph_ctx* ctx);
int wcsph_iisph_init_viscosity_para(wcsph_ctx* ctx);
int wcsph_iisph_compute_viscosity_force(wcsph_ctx* ctx);
int wcsph_iisph_combine_nonpressure(wcsph_ctx* ctx);
int wcsph_iisph_compute_advection(wcsph_ctx* ctx);
int wcsph_iisph_update_iter_info(wcsph_ctx* ctx);
int wcsph_iisph_update_pressure_force(wcsph_ctx* ctx);
int wcsph_iisph_update_pos(wcsph_ctx* ctx);
int wcsph_iisph_step(wcsph_ctx* ctx, int nsteps);           /* iisph.py:419-427 */

/* ---- pcisph.py:194-285 -------------------------------------------------- */
int wcsph_pcisph_reset_param(wcsph_ctx* ctx);
int wcsph_pcisph_compute_nonpressure_force(wcsph_ctx* ctx);
int wcsph_pcisph_compute_tension(wcsph_ctx* ctx);          /* dfsph.py:265-304 applied to PCISPH (BASELINE configs[2]) */
int wcsph_pcisph_init_iter_info(wcsph_ctx* ctx);
int wcsph_pcisph_update_iter_info(wcsph_ctx* ctx);
int wcsph_pcisph_predict_density(wcsph_ctx* ctx);
int wcsph_pcisph_update_pos(wcsph_ctx* ctx);
int wcsph_pcisph_step(wcsph_ctx* ctx, int nsteps);          /* pcisph.py:307-311 */

/* ---- the canvas the step loops draw into (Canvas.py) ------------------------------------
 * zbuf_dev: caller-owned device buffer of sx*sy 64-bit words (pixel (x,y) at x*sy + y), one packed
 * (depth, colour) key per pixel so that the depth test of Canvas.fill_pixel (Canvas.py:143-148) is an
 * atomicMin.  view16 / proj16: HOST pointers to Canvas.view[0] / Canvas.proj[0], row-major 4x4 f32.
 * style 0 = draw_particle of sesph.py:201-207 / pcisph.py:288-293 / iisph.py:401-406 (outline per liquid,
 * point per solid); style 1 = dfsph.py:585-593 (outline per liquid, then a point for every particle).
 * On a z-slab rank only the owned liquids (+ the replicated solids) are drawn: min-reduce zbuf across ranks. */
int wcsph_canvas_clear(wcsph_ctx* ctx, unsigned long long* zbuf_dev, int sx, int sy);                    /* Canvas.py:205-209 */
int wcsph_canvas_draw_particle(wcsph_ctx* ctx, const float* view16, const float* proj16, int sx, int sy,
                               int style, unsigned long long* zbuf_dev);                                /* Canvas.py:138-203 */
/* img_dev: f32 [sx][sy][3] (Canvas.img), depth_dev: f32 [sx][sy] (Canvas.depth) or NULL; device pointers */
int wcsph_canvas_resolve(wcsph_ctx* ctx, const unsigned long long* zbuf_dev, int sx, int sy,
                         float* img_dev, float* depth_dev);                                             /* dfsph.py:623 */

/* ---- surface reconstruction (MarchingCubeGrid.py) ----------------------------------------
 * The dense grid of MCGrid(particleR, maxInGrid, maxNeighbour, particle_data) (ParticleData.py:29,177):
 * gridR = 0.9 * particleR, searchR = 4 * gridR, node / cell (x,y,z) at index x*by*bz + y*bz + z.
 * work_dev: caller-owned device scratch of wcsph_mc_workspace_bytes(); surface_value_dev: f32[bx*by*bz];
 * triangle_dev: f32[max_vertex][3], max_vertex a multiple of 3 (MAX_VERTEX = 3000000, MarchingCubeGrid.py:8).
 * Uses pos and rho as the last solver step left them: call between a step and the next update_grid. */
typedef struct wcsph_mc_grid {
    double gridR;             /* MarchingCubeGrid.py:22 */
    float  isolevel;          /* :27 (0.5) */
    int    max_in_grid;       /* :17 (4): only the first max_in_grid liquids of a cell contribute (:173-177) */
    float  liqiudMass;        /* particle_data.liqiudMass (:203-204) -- ParticleData's value, which is not the solver
                                 module's own mass constant in sesph.py / pcisph.py */
    float  min_boundary[3];   /* :49, = scene bbox min - searchR (ParticleData.py:177) */
    int    block[3];          /* :61-63 blockSize */
} wcsph_mc_grid;
/* z-slab contexts: update_grid / cal_surface_point work on the rank's owned liquids and yield its SHARE of the node field; the
 * caller sums the shares over the ranks (MCGrid.cal_surface_point: one all_reduce) before marching_cube. */
size_t wcsph_mc_workspace_bytes(const wcsph_mc_grid* grid, int liquid_count);
int wcsph_mc_update_grid(wcsph_ctx* ctx, const wcsph_mc_grid* grid, void* work_dev, size_t work_bytes);           /* :160-179 */
int wcsph_mc_cal_surface_point(wcsph_ctx* ctx, const wcsph_mc_grid* grid, void* work_dev, size_t work_bytes,
                               float* surface_value_dev);                                                        /* :183-209 */
/* vertex_count_out (host) = vertex_count[0]; it keeps counting past max_vertex like the reference (:343-349),
 * triangles come out in cell order (the order of a serial run of the reference's atomic append) */
/* anisotropic branch of the surface reconstruction (ParticleData.py:187-285, MarchingCubeGrid.py:215-246; the reference keeps it
 * but leaves the two calls in export_surface commented out, :148-149).  All buffers are caller-owned DEVICE memory in slot order
 * (the cell-sorted order of wcsph_field_device views; use wcsph_sorted_id_device to map a slot to the reference index):
 * color [CL] f32, color_grad [CL] float4, pos_avr [CL] float4, G [CL] 3 x float4 rows.  Single-GPU contexts. */
size_t wcsph_pd_aniso_workspace_bytes(wcsph_ctx* ctx);
int wcsph_pd_compute_color_map(wcsph_ctx* ctx, void* work_dev, size_t work_bytes,
                               float* color_dev, float* color_grad4_dev);                                       /* ParticleData.py:187-218 */
int wcsph_pd_cal_anistropic_kernel(wcsph_ctx* ctx, float mc_searchR, void* work_dev, size_t work_bytes,
                                   float* pos_avr4_dev, float* G12_dev);                                         /* ParticleData.py:220-285 */
int wcsph_mc_cal_surface_point_anistropic(wcsph_ctx* ctx, const wcsph_mc_grid* grid, void* work_dev, size_t work_bytes,
                                          const float* pos_avr4_dev, const float* G12_dev, float* surface_value_dev); /* MarchingCubeGrid.py:215-246 */
int wcsph_mc_marching_cube(wcsph_ctx* ctx, const wcsph_mc_grid* grid, void* work_dev, size_t work_bytes,
                           const float* surface_value_dev, float* triangle_dev, int max_vertex,
                           int* vertex_count_out);                                                               /* :262-352 */

/* ---- boundary pre-processing, boundry.py (parallel Poisson-disk sampling of a triangle mesh) -----------------
 * Stand-alone (no wcsph_ctx): the caller owns ONE device workspace of wcsph_bd_workspace_bytes() bytes and passes a stream.
 * Order: init_point_set | set_points -> bitonic_sort -> build_hmap -> sample(phase, trial) in the order of the reference's
 * main loop (boundry.py:421-457: trial 0 phases 1..26, trials 1..9 phases 0..26) -> get("possion_sample").
 * tri_vertices: f32 [face_num][3][3]; tri_normal: f32 [3*face_num][3] (one normal per VERTEX, boundry.py:148-150 -- the sampler
 * indexes it with the face id, :361-362, kept); tri_area: f32 [face_num]. */
typedef struct wcsph_bd_desc {
    int   n;               /* numInitialPoints = int(40 * totalArea / (pi R^2))   boundry.py:164 */
    int   padding;         /* padding_num = get_pot_num(n) << 1, a power of two   :165 */
    int   hash_size;       /* hash_map_size = 3 n                                  :167 */
    int   phase_vec_max;   /* n / 8                                                :166 */
    int   sample_cap;      /* hash_sample_size = 5                                 :61  */
    float radius;          /* particleRadius                                       :21  */
    float gridR;           /* particleRadius / sqrt(3)                             :22  */
    float min_point[3];    /* bounding-box minimum of the mesh                     :128-136 */
} wcsph_bd_desc;
size_t wcsph_bd_workspace_bytes(const wcsph_bd_desc* desc);
int wcsph_bd_init_point_set(const wcsph_bd_desc* desc, void* work_dev, size_t work_bytes, const float* tri_vertices_dev, const float* tri_area_dev,
                            int face_num, float max_area, unsigned int seed, void* cuda_stream);                         /* boundry.py:223-247 */
int wcsph_bd_set_points(const wcsph_bd_desc* desc, void* work_dev, size_t work_bytes, const float* host_init_pos, const int* host_init_id,
                        void* cuda_stream);                                                                              /* injected initial point set */
int wcsph_bd_bitonic_sort(const wcsph_bd_desc* desc, void* work_dev, size_t work_bytes, void* cuda_stream);              /* :208-219, 322-336 */
int wcsph_bd_build_hmap(const wcsph_bd_desc* desc, void* work_dev, size_t work_bytes, void* cuda_stream);                /* :250-271 */
int wcsph_bd_sample(const wcsph_bd_desc* desc, void* work_dev, size_t work_bytes, const float* tri_normal_dev, int phase_group, int trial,
                    void* cuda_stream);                                                                                  /* :340-407 */
int wcsph_bd_get(const wcsph_bd_desc* desc, void* work_dev, size_t work_bytes, const char* name, void* host_dst, size_t dst_bytes,
                 void* cuda_stream);

#ifdef __cplusplus
}
#endif
#endif /* WCSPH_B200_H */
