// pcisph.cu -- predictive-corrective SPH (pcisph.py:194-285) on the compact in-range lists.
// D-PCI: compute_nonpressure_force is two-phase -- density sweep, then the
// viscosity sweep reads complete rho -- in both this file and the oracle.
#include "sweep.cuh"
#include "tension.cuh"

#define NEED(c, S) do { if (!(c) || (c)->desc.solver != (S)) { wcsph_set_error("%s: wrong solver / null ctx", __func__); return WCSPH_EINVAL; } } while (0)
#define STREAM_LAUNCH(c, kern, ...) do { prof_begin(c, #kern); kern<<<nblocks((c)->nown), WCSPH_BLOCK, 0, (c)->stream>>>(__VA_ARGS__); prof_end(c); LAUNCH_CHECK(c); } while (0)

// pcisph.py:194-197
__global__ void k_pci_reset(float4* vel, int NL, Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) sc->deltaT = 0.001f;
    if (i < NL) vel[i] = make_float4(0, 0, 0, 0);
}

// pcisph.py:203,212,215 density part
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_pci_density(SweepArgs A, float* __restrict__ rho) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    float wl = 0.f, ws = 0.f;
    FOR_LIQUID_EXACT(A, i, pi, { wl += cubic_W2(K, r2); })
    FOR_SOLID_EXACT(A, i, pi, { ws += cubic_W2(K, r2); })
    const float d = (K.VL0 * (cubic_W(K, 0.f) + wl) + K.VS0 * ws) * K.rho0;
    rho[i] = d;
    ((float*)A.pos)[4 * (size_t)i + 3] = d;              // pos.w carries rho_j
}

struct PciC { float c_l, c_s, h2c, gx, gy, gz; };
// pcisph.py:202,213,216 viscosity part
__global__ void __launch_bounds__(WCSPH_BLOCK, 4)
k_pci_visc(SweepArgs A, PciC C, const float* __restrict__ rho, const float4* __restrict__ vel, float4* __restrict__ d_vel) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float3 vi = xyz(vel[i]);
    const float rho_i = rho[i];
    float3 al = f3(0, 0, 0), as = f3(0, 0, 0);
    FOR_LIQUID(A, i, pi, { al += cubic_gradW(K, r, r2) * __fdividef(dot3(vi - xyz(vel[j]), r), pj4.w * (r2 + C.h2c)); })
    FOR_SOLID(A, i, pi, { as += cubic_gradW(K, r, r2) * __fdividef(dot3(vi, r), r2 + C.h2c); })
    d_vel[i] = f4(f3(C.gx, C.gy, C.gz) + al * C.c_l + as * (C.c_s * (rho_i / K.rho0)));
}

// init_iter_info pcisph.py:221-226
__global__ void k_pci_init_iter(const float4* __restrict__ pos, const float4* __restrict__ vel, float4* __restrict__ pos_star,
                                float4* __restrict__ vel_star, float* __restrict__ pressure, float4* __restrict__ d_vel_pre, int NL) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    vel_star[i] = vel[i]; pos_star[i] = pos[i]; pressure[i] = 0.f; d_vel_pre[i] = make_float4(0, 0, 0, 0);
}
// update_iter_info pcisph.py:229-235 (Q7: rho_err[0] = 0)
__global__ void k_pci_update_iter(const float4* __restrict__ pos, const float4* __restrict__ vel, const float4* __restrict__ d_vel,
                                  const float4* __restrict__ d_vel_pre, float4* __restrict__ pos_star, float4* __restrict__ vel_star,
                                  float* __restrict__ pressure, int NL, Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) sc->rho_err = 0.f;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    float4 v = vel[i], a = d_vel[i], ap = d_vel_pre[i], p = pos[i];
    float3 vs = f3(v.x + (a.x + ap.x) * dt, v.y + (a.y + ap.y) * dt, v.z + (a.z + ap.z) * dt);
    vel_star[i] = f4(vs);
    pos_star[i] = make_float4(p.x + vs.x * dt, p.y + vs.y * dt, p.z + vs.z * dt, 0.f);
    pressure[i] = 0.f;
}

// predict_density loop 1 pcisph.py:239-256 (Q8: positions, not predicted positions)
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_pci_predict(SweepArgs A, float* __restrict__ adv_rho, float* __restrict__ pressure, float4* __restrict__ pos_star, float pci_coff) {
    SWEEP_PROLOGUE(A)
    float v[1] = {0.f};
    if (live) {
        const float dt = A.sc->deltaT;
        float wl = 0.f, ws = 0.f;
        FOR_LIQUID_EXACT(A, i, pi, { wl += cubic_W2(K, r2); })
        FOR_SOLID_EXACT(A, i, pi, { ws += cubic_W2(K, r2); })
        float a = fmaxf(K.VL0 * (cubic_W(K, 0.f) + wl) + K.VS0 * ws, 1.0f);
        adv_rho[i] = a;
        const float pr = pressure[i] + pci_coff * (a - 1.0f) / (dt * dt);
        pressure[i] = pr; pos_star[i].w = pr;              // pos_star.w carries pressure_j for the next sweep
        v[0] = a - 1.0f;
    }
    block_partials<1, false>(v, A.partials);
}

// predict_density loop 2 pcisph.py:258-278: gradW(pos_i - pos_star_j)
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_pci_paccel(SweepArgs A, const float4* __restrict__ pos_star, const float* __restrict__ pressure, float4* __restrict__ d_vel_pre) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float dpi = pressure[i];
    float3 al = f3(0, 0, 0), as = f3(0, 0, 0);
    {   // liquid neighbours use the PREDICTED position of j (pcisph.py:266-267); pos_star.w = pressure_j
        const float4* A_POS_ = pos_star;
        FOR_NBRS_EXACT_(NBR_ROW4(A.nbr_l, A.capL, i - A.i0 + A.l0), A.nl_cnt[i - A.i0 + A.l0], pi, { al += cubic_gradW(K, r, r2) * (dpi + pj4.w); })
    }
    FOR_SOLID(A, i, pi, { as += cubic_gradW(K, r, r2); })
    d_vel_pre[i] = f4(al * (-K.VL0) + as * (-K.VS0 * dpi));
}

// update_pos pcisph.py:282-285
__global__ void k_pci_update_pos(float4* __restrict__ pos, float4* __restrict__ vel, const float4* __restrict__ d_vel,
                                 const float4* __restrict__ d_vel_pre, int NL, const Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    float4 v = vel[i], a = d_vel[i], ap = d_vel_pre[i], p = pos[i];
    v.x += (a.x + ap.x) * dt; v.y += (a.y + ap.y) * dt; v.z += (a.z + ap.z) * dt;
    p.x += v.x * dt; p.y += v.y * dt; p.z += v.z * dt;
    vel[i] = v; pos[i] = p;
}

static PciC pci_consts(const wcsph_params& p) {
    PciC C;
    C.c_l = (float)((double)p.dim_coff * (double)p.viscosity * (double)p.liqiudMass);
    C.c_s = (float)((double)p.dim_coff * (double)p.viscosity_b * (double)p.VS0);
    C.h2c = (float)(0.01 * (double)p.searchR * (double)p.searchR);
    C.gx = p.gravity[0]; C.gy = p.gravity[1]; C.gz = p.gravity[2];
    return C;
}

extern "C" int wcsph_pcisph_reset_param(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    STREAM_LAUNCH(c, k_pci_reset, fown<float4>(c, "vel"), c->nown, c->sc);
    return 0;
}
extern "C" int wcsph_pcisph_compute_nonpressure_force(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    LAUNCH_SWEEP(c, k_pci_density, make_sweep(c), fcur<float>(c, "rho"));
    // z-slab ranks: rho_j (pos.w) and v_j of the ghosts
    LAUNCH_SWEEP_HALO(c, { HALO(c, "pos"); HALO(c, "vel"); }, k_pci_visc, make_sweep(c), pci_consts(c->prm), fcur<float>(c, "rho"), fcur<float4>(c, "vel"), fcur<float4>(c, "d_vel"));
    return 0;
}
// BASELINE configs[2]: PCISPH + Akinci surface tension.  The reference's pcisph.py has no tension term;
// this adds dfsph.py's compute_tension (D-TENSION) to the PCISPH non-pressure acceleration.
extern "C" int wcsph_pcisph_compute_tension(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    return tension_compute(c);
}
extern "C" int wcsph_pcisph_init_iter_info(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    STREAM_LAUNCH(c, k_pci_init_iter, fown<float4>(c, "pos"), fown<float4>(c, "vel"), fown<float4>(c, "pos_star"), fown<float4>(c, "vel_star"),
                  fown<float>(c, "pressure"), fown<float4>(c, "d_vel_pre"), c->nown);
    return 0;
}
extern "C" int wcsph_pcisph_update_iter_info(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    STREAM_LAUNCH(c, k_pci_update_iter, fown<float4>(c, "pos"), fown<float4>(c, "vel"), fown<float4>(c, "d_vel"), fown<float4>(c, "d_vel_pre"),
                  fown<float4>(c, "pos_star"), fown<float4>(c, "vel_star"), fown<float>(c, "pressure"), c->nown, c->sc);
    return 0;
}
extern "C" int wcsph_pcisph_predict_density(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    LAUNCH_SWEEP_REDUCE(c, FIN_RHO_ERR, 0.f, k_pci_predict, make_sweep(c), fcur<float>(c, "adv_rho"), fcur<float>(c, "pressure"), fcur<float4>(c, "pos_star"), c->prm.pci_coff);
    LAUNCH_SWEEP_HALO(c, HALO(c, "pos_star"), k_pci_paccel, make_sweep(c), fcur<float4>(c, "pos_star"), fcur<float>(c, "pressure"), fcur<float4>(c, "d_vel_pre"));
    return 0;
}
extern "C" int wcsph_pcisph_update_pos(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    STREAM_LAUNCH(c, k_pci_update_pos, fown<float4>(c, "pos"), fown<float4>(c, "vel"), fown<float4>(c, "d_vel"), fown<float4>(c, "d_vel_pre"), c->nown, c->sc);
    return 0;
}

// pcisph.py:307-311 with sovel_pressure pcisph.py:147-157 (host-driven loop)
extern "C" int wcsph_pcisph_step(wcsph_ctx* c, int nsteps) {
    NEED(c, WCSPH_PCISPH);
    TRY(wcsph_fatal_flags(c));          // overflow seen by an earlier call: do not keep stepping on dropped pairs
    const double NLd = (double)c->NL;   // GLOBAL liquid count (thresholds of dfsph.py:143,163)
    for (int s = 0; s < nsteps; s++) {
        TRY(wcsph_hashgrid_update_grid(c));
        TRY(wcsph_pcisph_compute_nonpressure_force(c));
        if (c->prm.tension_coff != 0.0f || c->prm.tension_coff_b != 0.0f) TRY(wcsph_pcisph_compute_tension(c));
        c->pr_iter = 0;
        double err = 0.0;
        TRY(wcsph_pcisph_init_iter_info(c));
        while ((err > 0.01 || c->pr_iter < 3) && c->pr_iter < 50) {
            TRY(wcsph_pcisph_update_iter_info(c));
            TRY(wcsph_pcisph_predict_density(c));
            c->pr_iter++;
            if (c->pr_iter >= 3) {
                CUDA_TRY(cudaMemcpyAsync(c->sc_host, c->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
                CUDA_TRY(cudaStreamSynchronize(c->stream));
                c->seen_flags |= c->sc_host->flags;
                err = (double)c->sc_host->rho_err / NLd;
            }
        }
        TRY(wcsph_pcisph_update_pos(c));
    }
    return 0;
}
