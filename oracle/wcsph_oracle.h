/*
 * wcsph_oracle.h -- CPU restatement of the lyd405121/wcsph per-step hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under wcsph_b200/ may include, link or
 * call this.  Only tests/, __graft_entry__.smoke() and bench.py's
 * cpu_baseline / --impl reference legs use it, and there only as the checker
 * or the CPU baseline -- never as the thing measured or shipped.
 *
 * PARITY PINNED ON THE EXECUTED REFERENCE (round 2).  The reference is Taichi DSL and Taichi is not installable in this image, but
 * its unmodified sources run under the serial Taichi-semantics shim oracle/tishim (tests/golden/make_ref_exec.py); the per-kernel
 * goldens tests/golden/ref_exec_*.npz that run produced are what this restatement must reproduce (tests/test_ref_exec.py,
 * tests/test_boundry_cpu.py): HashGrid tables and ORDERED neighbour rows bit-exact, DFSPH / PCISPH / marching cubes / colour map /
 * canvas / boundry.py bit-exact, every other fp32 field within 2e-6, iteration counts and the adapted time step equal.
 * It follows the cited reference lines statement by statement in fp32 with the reference's own data structures (64-slot hash
 * buckets, table size = particle count, 2048-wide neighbour table, 125-cell stencil, one function per @ti.kernel).  Older anchors
 * that need no execution stay (tests/test_oracle_anchors.py): model/liqiud.obj == dfsph.py:70-73, pcisph.py:87-115 GetPciCoff,
 * the closed-form kernel identities and the t=0 hash statistics.
 *
 * Defined behaviour where the reference is undefined:
 *   Q7  pcisph.py:234   rho_err[i]=0 on a 1-element field -> rho_err[0]=0
 *   Q12 dfsph.py:324    omega[j], vel[j] for solid j (OOB)  -> 0
 *   Q15 dfsph.py:563    stride-doubling max tree reads OOB   -> what the executed reference yields: max over [0,P),
 *                       P = largest power of two < NL (the joining pass never runs); oracle_set_cfl_true_max(1): [0,NL)
 *   Q24 pcisph.py:203   rho reset+read race                  -> two-phase (D-PCI)
 *   Q11 dfsph.py:277-304 tension: order dependent            -> D-TENSION (see .c)
 *   Q3/Q4 overflow of the 2048 / 64 caps                     -> counted in flags, entry dropped
 *   atomic-add ordering                                      -> ascending particle index
 */
#ifndef WCSPH_ORACLE_H
#define WCSPH_ORACLE_H
#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
    /* geometry / discretisation */
    float searchR;      /* physics kernel support h */
    float m_k, m_l;     /* W and gradW normalisation, already divided by h^3 (sesph style) */
    float h3inv;        /* 1/h^3 as CubicKernel.py:14 keeps it (dfsph style: m_k*h3) */
    float m_k_raw, m_l_raw; /* 8/pi, 48/pi (CubicKernel.py:15-16) */
    float coh_m_k, coh_m_c; /* CohesionKernel.py:15-16 */
    float adh_m_k;          /* AdhesionKernel.py:15 */
    /* material */
    float rho_L0, rho_S0, VL0, VS0, liqiudMass;
    float gravity[3];
    float dim_coff, viscosity, viscosity_b, viscosity_err;
    float tension_coff, tension_coff_b;
    float viscosity_omega, vorticity_coff, vorticity_init;
    float stiffness;    /* sesph.py:58 */
    float pci_coff;     /* pcisph.py:87-115 */
    float omega_relax;  /* iisph.py:78 */
    float eps;          /* 1e-5 */
    float particleRadius;
    float user_max_t, user_min_t;
} OracleParams;

typedef struct Oracle Oracle;

Oracle* oracle_create(int count, int liquid_count, const float* pos,
                      double gridR, int maxInGrid, int maxNeighbour,
                      const float* maxboundary, const float* minboundary,
                      const OracleParams* prm);
void    oracle_destroy(Oracle* o);
void    oracle_set_threads(int n);
/* pointer to a named per-particle field (float* or int*), or NULL */
void*   oracle_field(Oracle* o, const char* name);
int     oracle_flag(Oracle* o, const char* name); /* overflow counters, iteration counts */
void    oracle_set_params(Oracle* o, const OracleParams* prm);

/* HashGrid.py:57-106 */
void hashgrid_update_grid(Oracle* o);

/* sesph.py:131-196 */
void sesph_reset_param(Oracle* o);
void sesph_update_advection_density(Oracle* o);
void sesph_update_pressure(Oracle* o);
void sesph_compute_force(Oracle* o);
void sesph_integrator_sesph(Oracle* o);
void sesph_step(Oracle* o);

/* dfsph.py:168-580 */
void dfsph_reset_param(Oracle* o);
void dfsph_compute_density(Oracle* o);
void dfsph_compute_dfsph_coff(Oracle* o);
void dfsph_warmstart_divergence_vel(Oracle* o);
void dfsph_begin_divergence_iter(Oracle* o);
void dfsph_divergence_iter(Oracle* o);
void dfsph_end_divergence_iter(Oracle* o);
void dfsph_solve_vel_divergence(Oracle* o);
void dfsph_clear_nonpressure(Oracle* o);
void dfsph_compute_tension(Oracle* o);
void dfsph_init_viscosity_para(Oracle* o);
void dfsph_compute_viscosity_force(Oracle* o);
void dfsph_end_viscosity(Oracle* o);
void dfsph_compute_vorticity(Oracle* o);
void dfsph_compute_nonpressure_force(Oracle* o);
void dfsph_optimize_time_step(Oracle* o);
void oracle_set_cfl_true_max(int on);
void dfsph_update_vel(Oracle* o);
void dfsph_warmstart_pressure(Oracle* o);
void dfsph_begin_pressure_iter(Oracle* o);
void dfsph_pressure_iter(Oracle* o);
void dfsph_end_pressure_iter(Oracle* o);
void dfsph_solve_pressure(Oracle* o);
void dfsph_update_pos(Oracle* o);
void dfsph_step(Oracle* o);

/* iisph.py:178-396 */
void iisph_reset_param(Oracle* o);
void iisph_compute_density(Oracle* o);
void iisph_init_viscosity_para(Oracle* o);
void iisph_compute_viscosity_force(Oracle* o);
void iisph_combine_nonpressure(Oracle* o);
void iisph_compute_nonpressure_force(Oracle* o);
void iisph_compute_advection(Oracle* o);
void iisph_update_iter_info(Oracle* o);
void iisph_update_pressure_force(Oracle* o);
void iisph_solve_pressure(Oracle* o);
void iisph_update_pos(Oracle* o);
void iisph_step(Oracle* o);

/* pcisph.py:194-285 */
void pcisph_reset_param(Oracle* o);
void pcisph_compute_nonpressure_force(Oracle* o);
void pcisph_compute_tension(Oracle* o);
void pcisph_init_iter_info(Oracle* o);
void pcisph_update_iter_info(Oracle* o);
void pcisph_predict_density(Oracle* o);
void pcisph_sovel_pressure(Oracle* o);
void pcisph_update_pos(Oracle* o);
void pcisph_step(Oracle* o);

/* kernels, exposed for the closed-form identity tests */
float oracle_cubic_W_norm(const OracleParams* p, float r, int sesph_style);
void  oracle_set_iters(Oracle* o, int vs, int dv, int pr);
void  oracle_cubic_gradW(const OracleParams* p, const float* r, float* out, int sesph_style);
float oracle_cohesion_W_norm(const OracleParams* p, float r);
float oracle_adhesion_W_norm(const OracleParams* p, float r);

/* §8(f) N1 -- Canvas.py:138-209 + the scripts' draw_particle (dfsph.py:585-593 style 1, sesph.py:201-207 style 0).
   img[sx][sy][3], depth[sx][sy] f32; view/proj row-major 4x4 f32. */
void oracle_canvas_clear(float* img, float* depth, int sx, int sy);
void oracle_canvas_draw_particle(const float* pos, int count, int liquid_count, const float* view, const float* proj,
                                 int sx, int sy, int style, float* img, float* depth);

/* §8(f) N2 -- MarchingCubeGrid.py:160-209,262-409 (serial; grid index = x*by*bz + y*bz + z) */
int  oracle_mc_update_grid(const float* pos, int count, const float* minb, const int* block, double gridR, int maxInGrid,
                           int* gridCount, int* grid);
void oracle_mc_cal_surface_point(const float* pos, const float* rho, int liquid_count, float liqiudMass,
                                 const float* minb, const int* block, double gridR, int maxInGrid,
                                 const int* gridCount, const int* grid, float* surface_value);
int  oracle_mc_marching_cube(const float* surface_value, const float* minb, const int* block, double gridR,
                             const int* edgetable, const int* tritable, float* triangle, int max_vertex);

/* §8(f) N2, anisotropic branch (restatement only, not built on the GPU yet): ParticleData.py:188-298, MarchingCubeGrid.py:215-243 */
void oracle_pd_compute_color_map(const float* pos, const float* rho, int liquid_count, const int* neighborCount, const int* neighbor,
                                 int maxNeighbour, const OracleParams* p, float* color, float* color_grad);
void oracle_pd_cal_anistropic_kernel(const float* pos, int liquid_count, const int* neighborCount, const int* neighbor,
                                     int maxNeighbour, float mc_searchR, float* pos_avr, float* G);
void oracle_mc_cal_surface_point_anistropic(const float* pos, const float* pos_avr, const float* G, const float* rho, int liquid_count,
                                            float liqiudMass, const float* minb, const int* block, double gridR, int maxInGrid,
                                            const int* gridCount, const int* grid, float* surface_value);

#ifdef __cplusplus
}
#endif
#endif
