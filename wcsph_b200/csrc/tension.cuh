// tension.cuh -- Akinci-2013 surface tension (dfsph.py:265-304) under the D-TENSION definition
// (SURVEY Q11): normals n_i = h sum_{j liquid} m/rho_j gradW_ij, then cohesion +
// curvature over the liquid neighbours and adhesion over the solid ones, sums over |r| <= h.  Shared by
// dfsph (compute_tension) and by pcisph for BASELINE configs[2] (PCISPH + Akinci tension).
#pragma once
#include "sweep.cuh"

static __global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_tension_normal(SweepArgs A, const float* __restrict__ rho, float4* __restrict__ normal) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    float3 n = f3(0, 0, 0);
    FOR_LIQUID(A, i, pi, { n += cubic_gradW(K, r, r2) * __fdividef(K.mass, pj4.w); })
    normal[i] = f4(n * K.h);
}
struct TensionC { float g, gb, sb, coh_m_k, coh_m_c, adh_m_k; };
__device__ __forceinline__ float coh_W(const TensionC& T, float h, float r) {     // CohesionKernel.py:18-29
    float res = 0.f, r2 = r * r;
    if (r2 <= h * h) {
        float r3 = r2 * r;
        const float d = h - r, d3 = d * d * d;          // pow(h - r, 3.0): exact for any sign (r may exceed h by an ulp)
        if (r > 0.5f * h) res = T.coh_m_k * d3 * r3;
        else res = T.coh_m_k * 2.0f * d3 * r3 - T.coh_m_c;
    }
    return res;
}
__device__ __forceinline__ float adh_W(const TensionC& T, float h, float r) {     // AdhesionKernel.py:21-29
    float res = 0.f, r2 = r * r;
    // the radicand vanishes at r = h and can round to -1e-9 there (lattice pairs at exactly 2 spacings): clamp
    if (r2 <= h * h && r > 0.5f * h) res = T.adh_m_k * powf(fmaxf(-4.0f * r2 / h + 6.0f * r - 2.0f * h, 0.0f), 0.25f);
    return res;
}
static __global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_tension_force(SweepArgs A, TensionC T, const float* __restrict__ rho, const float4* __restrict__ normal, float4* __restrict__ d_vel) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    float3 a = xyz(d_vel[i]);
    const float3 ni = xyz(normal[i]);
    const float rho_i = rho[i];
    FOR_LIQUID(A, i, pi, {
        const float len = sqrtf(r2);
        if (len / K.h <= 1.0f) {                 // the list may hold pairs a hair beyond h
            float k_ij = 2.0f * K.rho0 / (rho_i + pj4.w);
            float3 accel = (ni - xyz(normal[j])) * (-T.g);
            if (r2 > K.eps) accel += (r / len) * (-T.g * K.mass * coh_W(T, K.h, len));
            a += accel * k_ij;
        }
    })
    FOR_SOLID(A, i, pi, {
        const float len = sqrtf(r2);
        if (len / K.h <= 1.0f && r2 > K.eps) a += (r / len) * (-T.gb * T.sb * adh_W(T, K.h, len));
    })
    d_vel[i] = f4(a);
}


static inline int tension_compute(wcsph_ctx* c) {
    const wcsph_params& p = c->prm;
    LAUNCH_SWEEP_HALO(c, HALO(c, "pos") /* pos.w = rho_j */, k_tension_normal, make_sweep(c), fcur<float>(c, "rho"), fcur<float4>(c, "normal"));
    if (p.tension_coff == 0.0f && p.tension_coff_b == 0.0f) return 0;
    HALO(c, "normal");
    TensionC T; T.g = p.tension_coff; T.gb = p.tension_coff_b; T.sb = (float)((double)p.rho_S0 * (double)p.VS0);
    T.coh_m_k = p.coh_m_k; T.coh_m_c = p.coh_m_c; T.adh_m_k = p.adh_m_k;
    LAUNCH_SWEEP(c, k_tension_force, make_sweep(c), T, fcur<float>(c, "rho"), fcur<float4>(c, "normal"), fcur<float4>(c, "d_vel"));
    return 0;
}
