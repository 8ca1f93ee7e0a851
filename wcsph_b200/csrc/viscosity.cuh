// viscosity.cuh -- Weiler-2018 implicit viscosity, block-Jacobi preconditioned CG
// (dfsph.py:182-246 == iisph.py:185-252 up to the kernel evaluation style).
#pragma once
#include "sweep.cuh"

struct ViscC { float c_l, c_s, h2c, VS0, eps; };
static inline ViscC visc_consts(const wcsph_params& p) {
    ViscC C;
    C.c_l = (float)((double)p.dim_coff * (double)p.viscosity * (double)p.liqiudMass);
    C.c_s = (float)((double)p.dim_coff * (double)p.viscosity_b * (double)p.rho_S0);
    C.h2c = (float)(0.01 * (double)p.searchR * (double)p.searchR);
    C.VS0 = p.VS0; C.eps = p.eps;
    return C;
}

// get_viscosity_Ax dfsph.py:182-195: returns x_i - A x.  Per pair the reference evaluates
// c / rho_j * (x_ij . r) / (r^2 + 0.01 h^2) * gradW / rho_i * dt; the i-only factors are
// applied once after the loop and rho_j comes with the position gather (pos.w).
__device__ __forceinline__ float3 visc_Ax(const SweepArgs& A, const ViscC& C, int i, float3 pi,
                                          const float4* __restrict__ x, const float* __restrict__ rho, float dt) {
    const KC& K = A.k;
    const float3 xi = xyz(x[i]);
    const float rho_i = rho[i];
    float3 al = f3(0.f, 0.f, 0.f), as = f3(0.f, 0.f, 0.f);
    FOR_LIQUID(A, i, pi, {
        const float s = __fdividef(dot3(xi - xyz(x[j]), r), pj4.w * (r2 + C.h2c));
        al += r * (cubic_gradW_s(K, r2) * s);
    })
    FOR_SOLID(A, i, pi, {
        const float s = __fdividef(dot3(xi, r), r2 + C.h2c);
        as += r * (cubic_gradW_s(K, r2) * s);
    })
    const float f = dt / rho_i;
    return xi - (al * (C.c_l * f) + as * (C.c_s / rho_i * C.VS0 * f));
}

__device__ __forceinline__ void inv3x3(const float* m, float* o) {
    float a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], k = m[8];
    float A = e * k - f * h, B = -(d * k - f * g), Cc = d * h - e * g;
    float det = a * A + b * B + c * Cc;
    float id = 1.0f / det;
    o[0] = A * id;  o[1] = -(b * k - c * h) * id; o[2] = (b * f - c * e) * id;
    o[3] = B * id;  o[4] = (a * k - c * g) * id;  o[5] = -(a * f - c * d) * id;
    o[6] = Cc * id; o[7] = -(a * h - b * g) * id; o[8] = (a * e - b * d) * id;
}
__device__ __forceinline__ float3 matvec3(const float4* __restrict__ M, int i, float3 v) {
    float4 r0 = M[3 * (size_t)i], r1 = M[3 * (size_t)i + 1], r2 = M[3 * (size_t)i + 2];
    return f3(r0.x * v.x + r0.y * v.y + r0.z * v.z, r1.x * v.x + r1.y * v.y + r1.z * v.z, r2.x * v.x + r2.y * v.y + r2.z * v.z);
}

// init_viscosity_para loop 1 (dfsph.py:199-200): vel_guess += vel
static __global__ void k_visc_guess(float4* __restrict__ vel_guess, const float4* __restrict__ vel, int NL) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    float4 g = vel_guess[i], v = vel[i];
    vel_guess[i] = make_float4(g.x + v.x, g.y + v.y, g.z + v.z, 0.f);
}

// init_viscosity_para loop 2 (dfsph.py:202-215): cg_Minv = (I - dt/rho_i sum s gradW (x) r)^-1
static __global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_visc_minv(SweepArgs A, ViscC C, const float* __restrict__ rho, float4* __restrict__ Minv) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float rho_i = rho[i];
    float m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
    const float cs = C.c_s / rho_i * C.VS0;
    FOR_LIQUID(A, i, pi, {
        const float3 g = cubic_gradW(K, r, r2) * __fdividef(C.c_l, pj4.w * (r2 + C.h2c));
        m[0] += g.x * r.x; m[1] += g.x * r.y; m[2] += g.x * r.z;
        m[3] += g.y * r.x; m[4] += g.y * r.y; m[5] += g.y * r.z;
        m[6] += g.z * r.x; m[7] += g.z * r.y; m[8] += g.z * r.z;
    })
    FOR_SOLID(A, i, pi, {
        const float3 g = cubic_gradW(K, r, r2) * __fdividef(cs, r2 + C.h2c);
        m[0] += g.x * r.x; m[1] += g.x * r.y; m[2] += g.x * r.z;
        m[3] += g.y * r.x; m[4] += g.y * r.y; m[5] += g.y * r.z;
        m[6] += g.z * r.x; m[7] += g.z * r.y; m[8] += g.z * r.z;
    })
    const float f = A.sc->deltaT / rho_i;
    float a[9], o[9];
#pragma unroll
    for (int t = 0; t < 9; t++) a[t] = ((t == 0 || t == 4 || t == 8) ? 1.0f : 0.0f) - m[t] * f;
    inv3x3(a, o);
    Minv[3 * (size_t)i] = make_float4(o[0], o[1], o[2], 0.f);
    Minv[3 * (size_t)i + 1] = make_float4(o[3], o[4], o[5], 0.f);
    Minv[3 * (size_t)i + 2] = make_float4(o[6], o[7], o[8], 0.f);
}

// init_viscosity_para loop 3 (dfsph.py:217-223): r = v - A(vel_guess); dir = Minv r; delta0 = sum r.dir
static __global__ void __launch_bounds__(WCSPH_BLOCK, 3)
k_visc_residual(SweepArgs A, ViscC C, const float* __restrict__ rho, const float4* __restrict__ vel,
                const float4* __restrict__ vel_guess, const float4* __restrict__ Minv,
                float4* __restrict__ cg_r, float4* __restrict__ cg_dir) {
    SWEEP_PROLOGUE(A)
    float v[1] = {0.f};
    if (live) {
        float3 r = xyz(vel[i]) - visc_Ax(A, C, i, pi, vel_guess, rho, A.sc->deltaT);
        float3 d = matvec3(Minv, i, r);
        cg_r[i] = f4(r); cg_dir[i] = f4(d);
        v[0] = dot3(r, d);
    }
    block_partials<1, false>(v, A.partials);
}

// init_viscosity_para loops 2+3 in ONE sweep (fused step path): the preconditioner block and
// A(vel_guess) walk the same pairs; Minv_i is only needed by particle i itself
static __global__ void __launch_bounds__(WCSPH_BLOCK, 3)
k_visc_minv_residual(SweepArgs A, ViscC C, const float* __restrict__ rho, const float4* __restrict__ vel,
                     const float4* __restrict__ vel_guess, float4* __restrict__ Minv,
                     float4* __restrict__ cg_r, float4* __restrict__ cg_dir) {
    SWEEP_PROLOGUE(A)
    float v[1] = {0.f};
    if (live) {
        const float rho_i = rho[i];
        const float dt = A.sc->deltaT;
        const float cs = C.c_s / rho_i * C.VS0;
        const float3 xi = xyz(vel_guess[i]);
        float m[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
        float3 al = f3(0, 0, 0), as = f3(0, 0, 0);
        FOR_LIQUID(A, i, pi, {
            const float inv = __fdividef(1.0f, pj4.w * (r2 + C.h2c));
            const float3 g0 = cubic_gradW(K, r, r2);
            const float3 g = g0 * (C.c_l * inv);
            m[0] += g.x * r.x; m[1] += g.x * r.y; m[2] += g.x * r.z;
            m[3] += g.y * r.x; m[4] += g.y * r.y; m[5] += g.y * r.z;
            m[6] += g.z * r.x; m[7] += g.z * r.y; m[8] += g.z * r.z;
            al += g0 * (dot3(xi - xyz(vel_guess[j]), r) * inv);
        })
        FOR_SOLID(A, i, pi, {
            const float inv = __fdividef(1.0f, r2 + C.h2c);
            const float3 g0 = cubic_gradW(K, r, r2);
            const float3 g = g0 * (cs * inv);
            m[0] += g.x * r.x; m[1] += g.x * r.y; m[2] += g.x * r.z;
            m[3] += g.y * r.x; m[4] += g.y * r.y; m[5] += g.y * r.z;
            m[6] += g.z * r.x; m[7] += g.z * r.y; m[8] += g.z * r.z;
            as += g0 * (dot3(xi, r) * inv);
        })
        const float f = dt / rho_i;
        float a[9], o[9];
#pragma unroll
        for (int t = 0; t < 9; t++) a[t] = ((t == 0 || t == 4 || t == 8) ? 1.0f : 0.0f) - m[t] * f;
        inv3x3(a, o);
        Minv[3 * (size_t)i] = make_float4(o[0], o[1], o[2], 0.f);
        Minv[3 * (size_t)i + 1] = make_float4(o[3], o[4], o[5], 0.f);
        Minv[3 * (size_t)i + 2] = make_float4(o[6], o[7], o[8], 0.f);
        const float3 ax = xi - (al * (C.c_l * f) + as * (cs * f));
        const float3 rr = xyz(vel[i]) - ax;
        const float3 d = f3(o[0] * rr.x + o[1] * rr.y + o[2] * rr.z, o[3] * rr.x + o[4] * rr.y + o[5] * rr.z, o[6] * rr.x + o[7] * rr.y + o[8] * rr.z);
        cg_r[i] = f4(rr); cg_dir[i] = f4(d);
        v[0] = dot3(rr, d);
    }
    block_partials<1, false>(v, A.partials);
}

// compute_viscosity_force loop 1 (dfsph.py:228-230): Ad = A dir; dAd = eps + sum dir.Ad
static __global__ void __launch_bounds__(WCSPH_BLOCK, 4)
k_visc_Ad(SweepArgs A, ViscC C, const float* __restrict__ rho, const float4* __restrict__ cg_dir, float4* __restrict__ cg_Ad) {
    SWEEP_PROLOGUE(A)
    float v[1] = {0.f};
    if (live) {
        float3 ad = visc_Ax(A, C, i, pi, cg_dir, rho, A.sc->deltaT);
        cg_Ad[i] = f4(ad);
        v[0] = dot3(xyz(cg_dir[i]), ad);
    }
    block_partials<1, false>(v, A.partials);
}

// compute_viscosity_force loop 2 (dfsph.py:233-240)
static __global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_visc_update(int NL, Scalars* sc, float* partials, float4* __restrict__ vel_guess, float4* __restrict__ cg_r,
              const float4* __restrict__ cg_dir, const float4* __restrict__ cg_Ad, const float4* __restrict__ Minv,
              float4* __restrict__ cg_s) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const float delta = sc->cg_delta;
    const float alpha = delta / sc->cg_dAd;
    float v[1] = {0.f};
    if (i < NL) {
        float3 d = xyz(cg_dir[i]);
        float3 g = xyz(vel_guess[i]) + d * alpha;
        float3 r = xyz(cg_r[i]) - xyz(cg_Ad[i]) * alpha;
        float3 s = matvec3(Minv, i, r);
        vel_guess[i] = f4(g); cg_r[i] = f4(r); cg_s[i] = f4(s);
        v[0] = dot3(r, s);
    }
    block_partials<1, false>(v, partials);
}

// compute_viscosity_force loop 3 (dfsph.py:243-246)
static __global__ void k_visc_dir(int NL, const Scalars* sc, const float4* __restrict__ cg_s, float4* __restrict__ cg_dir) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float beta = sc->cg_delta / sc->cg_delta_old;
    float4 s = cg_s[i], d = cg_dir[i];
    cg_dir[i] = make_float4(s.x + beta * d.x, s.y + beta * d.y, s.z + beta * d.z, 0.f);
}

// host side: one call per reference kernel, shared by the dfsph_ / iisph_ entry points
static inline int visc_init_viscosity_para(wcsph_ctx* c) {
    SweepArgs A = make_sweep(c); ViscC C = visc_consts(c->prm);
    k_visc_guess<<<nblocks(c->nown), WCSPH_BLOCK, 0, c->stream>>>(fown<float4>(c, "vel_guess"), fown<float4>(c, "vel"), c->nown); LAUNCH_CHECK(c);
    LAUNCH_SWEEP_HALO(c, { HALO(c, "pos"); HALO(c, "vel_guess"); }, k_visc_minv, make_sweep(c), C, fcur<float>(c, "rho"), fcur<float4>(c, "cg_Minv"));
    LAUNCH_SWEEP_REDUCE(c, FIN_CG_DELTA0, 0.f, k_visc_residual, A, C, fcur<float>(c, "rho"), fcur<float4>(c, "vel"), fcur<float4>(c, "vel_guess"),
                 fcur<float4>(c, "cg_Minv"), fcur<float4>(c, "cg_r"), fcur<float4>(c, "cg_dir"));
    return 0;
}
static inline int visc_compute_viscosity_force(wcsph_ctx* c) {
    SweepArgs A = make_sweep(c); ViscC C = visc_consts(c->prm);
    LAUNCH_SWEEP_HALO_REDUCE(c, HALO(c, "cg_dir"), FIN_CG_DAD, C.eps, k_visc_Ad, make_sweep(c), C, fcur<float>(c, "rho"), fcur<float4>(c, "cg_dir"), fcur<float4>(c, "cg_Ad"));
    LAUNCH_SWEEP_REDUCE(c, FIN_CG_DELTA, 0.f, k_visc_update, c->nown, c->sc, c->partials, fown<float4>(c, "vel_guess"), fown<float4>(c, "cg_r"),
        fown<float4>(c, "cg_dir"), fown<float4>(c, "cg_Ad"), fown<float4>(c, "cg_Minv"), fown<float4>(c, "cg_s"));
    k_visc_dir<<<nblocks(c->nown), WCSPH_BLOCK, 0, c->stream>>>(c->nown, c->sc, fown<float4>(c, "cg_s"), fown<float4>(c, "cg_dir")); LAUNCH_CHECK(c);
    return 0;
}

static inline int visc_init_fused(wcsph_ctx* c) {      // vel_guess += vel must already have run
    SweepArgs A = make_sweep(c); ViscC C = visc_consts(c->prm);
    // (pos too: the ghosts' pos.w is rho_j again after a DFSPH correction loop used it for kfac)
    LAUNCH_SWEEP_HALO_REDUCE(c, { HALO(c, "pos"); HALO(c, "vel_guess"); }, FIN_CG_DELTA0, 0.f, k_visc_minv_residual, make_sweep(c), C, fcur<float>(c, "rho"), fcur<float4>(c, "vel"), fcur<float4>(c, "vel_guess"),
                        fcur<float4>(c, "cg_Minv"), fcur<float4>(c, "cg_r"), fcur<float4>(c, "cg_dir"));
    return 0;
}

// fetch a few scalars for a host-driven loop test (one stream sync)
static inline int fetch_scalars(wcsph_ctx* c) {
    CUDA_TRY(cudaMemcpyAsync(c->sc_host, c->sc, sizeof(Scalars), cudaMemcpyDeviceToHost, c->stream));
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    c->seen_flags |= c->sc_host->flags;
    return 0;
}

// the CG loop of dfsph.py:93-99 / iisph.py:114-125 (host-driven)
static inline int visc_cg_loop(wcsph_ctx* c, bool fused_init = false) {
    if (fused_init) TRY(visc_init_fused(c)); else TRY(visc_init_viscosity_para(c));
    c->vs_iter = 0;
    while (c->vs_iter < 100) {
        TRY(visc_compute_viscosity_force(c));
        c->vs_iter++;
        TRY(fetch_scalars(c));
        if ((double)c->sc_host->cg_delta <= (double)c->prm.viscosity_err * (double)c->sc_host->cg_delta_zero ||
            (double)c->sc_host->cg_delta_zero < (double)c->prm.eps)
            break;
    }
    return 0;
}
