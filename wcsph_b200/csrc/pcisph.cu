// pcisph.cu -- predictive-corrective SPH (pcisph.py:194-285) on the compact in-range lists.
// D-PCI: compute_nonpressure_force is two-phase -- density sweep, then the
// viscosity sweep reads complete rho -- in both this file and the oracle.
#include "sweep.cuh"

#define NEED(c, S) do { if (!(c) || (c)->desc.solver != (S)) { wcsph_set_error("%s: wrong solver / null ctx", __func__); return WCSPH_EINVAL; } } while (0)
#define STREAM_LAUNCH(c, kern, ...) do { prof_begin(c, #kern); kern<<<nblocks((c)->NL), WCSPH_BLOCK, 0, (c)->stream>>>(__VA_ARGS__); prof_end(c); LAUNCH_CHECK(c); } while (0)

// pcisph.py:194-197
__global__ void k_pci_reset(float4* vel, int NL, Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) sc->deltaT = 0.001f;
    if (i < NL) vel[i] = make_float4(0, 0, 0, 0);
}

// pcisph.py:203,212,215 density part
__global__ void __launch_bounds__(WCSPH_BLOCK)
k_pci_density(SweepArgs A, float* __restrict__ rho) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    float d = K.VL0 * cubic_W(K, 0.f) * K.rho0;
    FOR_LIQUID(A, i, pi, { d += K.VL0 * cubic_W(K, sqrtf(r2)) * K.rho0; })
    FOR_SOLID(A, i, pi, { d += K.VS0 * cubic_W(K, sqrtf(r2)) * K.rho0; })
    rho[i] = d;
}

struct PciC { float c_l, c_s, h2c, gx, gy, gz; };
// pcisph.py:202,213,216 viscosity part
__global__ void __launch_bounds__(WCSPH_BLOCK)
k_pci_visc(SweepArgs A, PciC C, const float* __restrict__ rho, const float4* __restrict__ vel, float4* __restrict__ d_vel) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float3 vi = xyz(vel[i]);
    const float rho_i = rho[i];
    float3 a = f3(C.gx, C.gy, C.gz);
    FOR_LIQUID(A, i, pi, {
        float s = C.c_l / rho[j] * dot3(vi - xyz(vel[j]), r) / (r2 + C.h2c);
        a += cubic_gradW(K, r, r2) * s;
    })
    FOR_SOLID(A, i, pi, {
        float s = C.c_s * (rho_i / K.rho0) * dot3(vi, r) / (r2 + C.h2c);
        a += cubic_gradW(K, r, r2) * s;
    })
    d_vel[i] = f4(a);
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
__global__ void __launch_bounds__(WCSPH_BLOCK)
k_pci_predict(SweepArgs A, float* __restrict__ adv_rho, float* __restrict__ pressure, float pci_coff) {
    SWEEP_PROLOGUE(A)
    float v[1] = {0.f};
    if (live) {
        const float dt = A.sc->deltaT;
        float a = K.VL0 * cubic_W(K, 0.f);
        FOR_LIQUID(A, i, pi, { a += K.VL0 * cubic_W(K, sqrtf(r2)); })
        FOR_SOLID(A, i, pi, { a += K.VS0 * cubic_W(K, sqrtf(r2)); })
        a = fmaxf(a, 1.0f);
        adv_rho[i] = a;
        pressure[i] += pci_coff * (a - 1.0f) / (dt * dt);
        v[0] = a - 1.0f;
    }
    Scalars* sc = A.sc;
    grid_reduce<1, false>(v, A.partials, &sc->ticket, [sc](float* t) { sc->rho_err += t[0]; });
}

// predict_density loop 2 pcisph.py:258-278: gradW(pos_i - pos_star_j)
__global__ void __launch_bounds__(WCSPH_BLOCK)
k_pci_paccel(SweepArgs A, const float4* __restrict__ pos_star, const float* __restrict__ pressure, float4* __restrict__ d_vel_pre) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float dpi = pressure[i];
    float3 a = f3(0, 0, 0);
    {   // liquid neighbours use the PREDICTED position of j (pcisph.py:266-267)
        const uint32_t* row_ = NBR_ROW(A.nbr_l, A.capL, i);
        const int n_ = A.nl_cnt[i];
        for (int k_ = 0; k_ < n_; k_++) {
            const int j = (int)row_[(size_t)k_ * 32];
            const float4 pj = pos_star[j];
            const float3 r = f3(pi.x - pj.x, pi.y - pj.y, pi.z - pj.z);
            a += cubic_gradW(K, r, dot3(r, r)) * (-K.VL0 * (dpi + pressure[j]));
        }
    }
    FOR_SOLID(A, i, pi, { a += cubic_gradW(K, r, r2) * (-K.VS0 * dpi); })
    d_vel_pre[i] = f4(a);
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
    STREAM_LAUNCH(c, k_pci_reset, fcur<float4>(c, "vel"), c->NL, c->sc);
    return 0;
}
extern "C" int wcsph_pcisph_compute_nonpressure_force(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    LAUNCH_SWEEP(c, k_pci_density, make_sweep(c), fcur<float>(c, "rho"));
    LAUNCH_SWEEP(c, k_pci_visc, make_sweep(c), pci_consts(c->prm), fcur<float>(c, "rho"), fcur<float4>(c, "vel"), fcur<float4>(c, "d_vel"));
    return 0;
}
extern "C" int wcsph_pcisph_init_iter_info(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    STREAM_LAUNCH(c, k_pci_init_iter, fcur<float4>(c, "pos"), fcur<float4>(c, "vel"), fcur<float4>(c, "pos_star"), fcur<float4>(c, "vel_star"),
                  fcur<float>(c, "pressure"), fcur<float4>(c, "d_vel_pre"), c->NL);
    return 0;
}
extern "C" int wcsph_pcisph_update_iter_info(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    STREAM_LAUNCH(c, k_pci_update_iter, fcur<float4>(c, "pos"), fcur<float4>(c, "vel"), fcur<float4>(c, "d_vel"), fcur<float4>(c, "d_vel_pre"),
                  fcur<float4>(c, "pos_star"), fcur<float4>(c, "vel_star"), fcur<float>(c, "pressure"), c->NL, c->sc);
    return 0;
}
extern "C" int wcsph_pcisph_predict_density(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    LAUNCH_SWEEP(c, k_pci_predict, make_sweep(c), fcur<float>(c, "adv_rho"), fcur<float>(c, "pressure"), c->prm.pci_coff);
    LAUNCH_SWEEP(c, k_pci_paccel, make_sweep(c), fcur<float4>(c, "pos_star"), fcur<float>(c, "pressure"), fcur<float4>(c, "d_vel_pre"));
    return 0;
}
extern "C" int wcsph_pcisph_update_pos(wcsph_ctx* c) {
    NEED(c, WCSPH_PCISPH);
    STREAM_LAUNCH(c, k_pci_update_pos, fcur<float4>(c, "pos"), fcur<float4>(c, "vel"), fcur<float4>(c, "d_vel"), fcur<float4>(c, "d_vel_pre"), c->NL, c->sc);
    return 0;
}

// pcisph.py:307-311 with sovel_pressure pcisph.py:147-157 (host-driven loop)
extern "C" int wcsph_pcisph_step(wcsph_ctx* c, int nsteps) {
    NEED(c, WCSPH_PCISPH);
    const double NLd = (double)c->NL;
    for (int s = 0; s < nsteps; s++) {
        TRY(wcsph_hashgrid_update_grid(c));
        TRY(wcsph_pcisph_compute_nonpressure_force(c));
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
                err = (double)c->sc_host->rho_err / NLd;
            }
        }
        TRY(wcsph_pcisph_update_pos(c));
    }
    return 0;
}
