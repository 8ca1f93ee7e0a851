// sesph.cu -- state-equation SPH (sesph.py:131-196) on the compact in-range lists.
// Multi-GPU (z-slab ranks): update_grid brings the ghost positions, the force sweep exchanges pos.w / vel before it runs.
#include "sweep.cuh"

// sesph.py:131-136
__global__ void k_sesph_reset(float4* vel, float* pressure, int NL, Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) sc->deltaT = 0.001f;
    if (i >= NL) return;
    vel[i] = make_float4(0.f, 0.f, 0.f, 0.f); pressure[i] = 0.f;
}

// sesph.py:139-155 (+ :159-166 when FUSE_EOS)
template <bool FUSE_EOS>
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_sesph_density(SweepArgs A, float* __restrict__ rho, float* __restrict__ pressure, float4* __restrict__ vel, float stiffness) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    float wl = 0.f, ws = 0.f;
    FOR_LIQUID_EXACT(A, i, pi, { wl += cubic_W2(K, r2); })
    FOR_SOLID_EXACT(A, i, pi, { ws += cubic_W2(K, r2); })
    float d = (K.VL0 * (cubic_W(K, 0.f) + wl) + K.VS0 * ws) * K.rho0;
    if (FUSE_EOS) {
        d = fmaxf(d, K.rho0);
        float q = d / K.rho0, qq = q * q, qqqq = qq * qq;
        const float pr = stiffness * (qqqq * qq * q - 1.0f);
        pressure[i] = pr; vel[i].w = pr;                    // vel.w carries pressure_j for compute_force
    }
    rho[i] = d;
    ((float*)A.pos)[4 * (size_t)i + 3] = d;              // pos.w carries rho_j for compute_force
}

// sesph.py:159-166
__global__ void k_sesph_pressure(float* __restrict__ rho, float* __restrict__ pressure, float4* __restrict__ pos, float4* __restrict__ vel, int NL, float rho0, float stiffness) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    float d = fmaxf(rho[i], rho0);
    rho[i] = d; pos[i].w = d;
    float q = d / rho0, qq = q * q, qqqq = qq * qq;
    const float pr = stiffness * (qqqq * qq * q - 1.0f);
    pressure[i] = pr; vel[i].w = pr;
}

struct SesphForceC { float c_l, c_s, h2c, pl, ps, r00; float gx, gy, gz; };

// sesph.py:169-189 (+ :192-196 when FUSE_INTEGRATE: legal because pos/vel of j are read
// from the pre-step buffers and written to the other pair)
__global__ void __launch_bounds__(WCSPH_BLOCK, 4)
k_sesph_force(SweepArgs A, const float4* __restrict__ vel, const float* __restrict__ rho,
              const float* __restrict__ pressure, float4* __restrict__ d_vel, SesphForceC C) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float3 vi = xyz(vel[i]);
    const float rho_i = rho[i], p_i = pressure[i];
    const float pi_term = p_i / (rho_i * rho_i);
    float3 a = f3(C.gx, C.gy, C.gz);
    const float cs = C.c_s * (rho_i / K.rho0);
    const float prs = C.ps * (pi_term + p_i / C.r00);     // Q22
    FOR_LIQUID(A, i, pi, {
        const float3 g = cubic_gradW(K, r, r2);
        const float rho_j = pj4.w;
        const float4 vj = vel[j];                          // vel.w carries pressure_j
        float s = C.c_l * __fdividef(dot3(vi - xyz(vj), r), rho_j * (r2 + C.h2c));
        float pr = C.pl * (pi_term + __fdividef(vj.w, rho_j * rho_j));
        a += g * (s + pr);
    })
    FOR_SOLID(A, i, pi, {
        const float3 g = cubic_gradW(K, r, r2);
        float s = cs * __fdividef(dot3(vi, r), r2 + C.h2c);
        a += g * (s + prs);
    })
    d_vel[i] = f4(a);
}

// sesph.py:192-196
__global__ void k_sesph_integrate(float4* __restrict__ pos, float4* __restrict__ vel, const float4* __restrict__ d_vel,
                                  int NL, const Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    float4 v = vel[i], a = d_vel[i], p = pos[i];
    v.x += a.x * dt; v.y += a.y * dt; v.z += a.z * dt;
    p.x += v.x * dt; p.y += v.y * dt; p.z += v.z * dt;
    vel[i] = v; pos[i] = p;
}

static SesphForceC force_consts(const wcsph_params& p) {
    SesphForceC C;
    C.c_l = (float)((double)p.dim_coff * (double)p.viscosity * (double)p.liqiudMass);
    C.c_s = (float)((double)p.dim_coff * (double)p.viscosity_b * (double)p.VS0);
    C.h2c = (float)(0.01 * (double)p.searchR * (double)p.searchR);
    C.pl = (float)(-(double)p.rho_L0 * (double)p.VL0);
    C.ps = (float)(-(double)p.rho_L0 * (double)p.VS0);
    C.r00 = (float)((double)p.rho_L0 * (double)p.rho_L0);
    C.gx = p.gravity[0]; C.gy = p.gravity[1]; C.gz = p.gravity[2];
    return C;
}

#define NEED(c, S) do { if (!(c) || (c)->desc.solver != (S)) { wcsph_set_error("%s: wrong solver / null ctx", __func__); return WCSPH_EINVAL; } } while (0)

extern "C" int wcsph_sesph_reset_param(wcsph_ctx* c) {
    NEED(c, WCSPH_SESPH);
    k_sesph_reset<<<nblocks(c->nown), WCSPH_BLOCK, 0, c->stream>>>(fown<float4>(c, "vel"), fown<float>(c, "pressure"), c->nown, c->sc);
    LAUNCH_CHECK(c); return 0;
}
extern "C" int wcsph_sesph_update_advection_density(wcsph_ctx* c) {
    NEED(c, WCSPH_SESPH);
    LAUNCH_SWEEP(c, k_sesph_density<false>, make_sweep(c), fcur<float>(c, "rho"), fcur<float>(c, "pressure"), fcur<float4>(c, "vel"), c->prm.stiffness);
    return 0;
}
extern "C" int wcsph_sesph_update_pressure(wcsph_ctx* c) {
    NEED(c, WCSPH_SESPH);
    k_sesph_pressure<<<nblocks(c->nown), WCSPH_BLOCK, 0, c->stream>>>(fown<float>(c, "rho"), fown<float>(c, "pressure"), fown<float4>(c, "pos"), fown<float4>(c, "vel"), c->nown, c->prm.rho_L0, c->prm.stiffness);
    LAUNCH_CHECK(c); return 0;
}
extern "C" int wcsph_sesph_compute_force(wcsph_ctx* c) {
    NEED(c, WCSPH_SESPH);
    // z-slab ranks: the force sweep gathers rho_j (pos.w) and v_j, p_j (vel.xyz, vel.w) of ghost particles
    LAUNCH_SWEEP_HALO(c, { HALO(c, "pos"); HALO(c, "vel"); }, k_sesph_force, make_sweep(c), fcur<float4>(c, "vel"), fcur<float>(c, "rho"),
                      fcur<float>(c, "pressure"), fcur<float4>(c, "d_vel"), force_consts(c->prm));
    return 0;
}
extern "C" int wcsph_sesph_integrator_sesph(wcsph_ctx* c) {
    NEED(c, WCSPH_SESPH);
    k_sesph_integrate<<<nblocks(c->nown), WCSPH_BLOCK, 0, c->stream>>>(fown<float4>(c, "pos"), fown<float4>(c, "vel"), fown<float4>(c, "d_vel"), c->nown, c->sc);
    LAUNCH_CHECK(c); return 0;
}

// sesph.py:220-225; density+EOS fused (update_pressure only touches particle i)
extern "C" int wcsph_sesph_step(wcsph_ctx* c, int nsteps) {
    NEED(c, WCSPH_SESPH);
    TRY(wcsph_fatal_flags(c));          // overflow seen by an earlier call: do not keep stepping on dropped pairs
    for (int s = 0; s < nsteps; s++) {
        TRY(wcsph_hashgrid_update_grid(c));
        LAUNCH_SWEEP(c, k_sesph_density<true>, make_sweep(c), fcur<float>(c, "rho"), fcur<float>(c, "pressure"), fcur<float4>(c, "vel"), c->prm.stiffness);
        TRY(wcsph_sesph_compute_force(c));
        TRY(wcsph_sesph_integrator_sesph(c));
    }
    return 0;
}
