// dfsph.cu -- divergence-free SPH (dfsph.py:168-580) on the compact in-range lists.
#include "viscosity.cuh"
#include "tension.cuh"

#define NEED(c, S) do { if (!(c) || (c)->desc.solver != (S)) { wcsph_set_error("%s: wrong solver / null ctx", __func__); return WCSPH_EINVAL; } } while (0)
#define STREAM_LAUNCH(c, kern, ...) do { prof_begin(c, #kern); kern<<<nblocks((c)->nown), WCSPH_BLOCK, 0, (c)->stream>>>(__VA_ARGS__); prof_end(c); LAUNCH_CHECK(c); } while (0)

// dfsph.py:168-178
__global__ void k_dfsph_reset(float4* vel, float4* omega, float* pressure, float* kappa, float* kappa_v, int NL, Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i == 0) sc->deltaT = 0.001f;
    if (i >= NL) return;
    vel[i] = make_float4(0, 0, 0, 0); omega[i] = make_float4(0, 0, 0, 0);
    pressure[i] = 0.f; kappa[i] = 0.f; kappa_v[i] = 0.f;
}

// compute_density dfsph.py:249-262 and compute_dfsph_coff dfsph.py:346-372: both read only
// pos, so one sweep can serve both (DO_RHO / DO_ALPHA select the outputs)
template <bool DO_RHO, bool DO_ALPHA>
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_dfsph_density_alpha(SweepArgs A, float* __restrict__ rho, float* __restrict__ alpha) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    // sums are factored: rho = VL0 rho0 (W0 + sum_l W) + VS0 rhoS0 sum_s W, and likewise for the
    // gradient sums of alpha -- same terms as dfsph.py:251-262 / :354-366, one multiply at the end
    float wl = 0.f, ws = 0.f, g2 = 0.f;
    float3 gl = f3(0, 0, 0), gs = f3(0, 0, 0);
    FOR_LIQUID_EXACT(A, i, pi, {
        if (DO_RHO) wl += cubic_W2(K, r2);
        if (DO_ALPHA) { float3 g = cubic_gradW(K, r, r2); g2 += dot3(g, g); gl += g; }
    })
    FOR_SOLID_EXACT(A, i, pi, {
        if (DO_RHO) ws += cubic_W2(K, r2);
        if (DO_ALPHA) gs += cubic_gradW(K, r, r2);
    })
    if (DO_RHO) {
        float d = K.VL0 * K.rho0 * (cubic_W(K, 0.f) + wl) + K.VS0 * K.rhoS0 * ws;
        rho[i] = d;
        ((float*)A.pos)[4 * (size_t)i + 3] = d;          // pos.w carries rho_j for the later gathers
    }
    if (DO_ALPHA) {
        float3 sg = gl * K.VL0 + gs * K.VS0;
        float sgs = K.VL0 * K.VL0 * g2 + dot3(sg, sg);
        alpha[i] = (sgs > K.eps) ? -1.0f / sgs : 0.0f;
    }
}

// update_drho_divergence dfsph.py:375-392 (MODE 0) / update_drho_pressure dfsph.py:395-412 (MODE 1)
// PRE   : warmstart_divergence_vel loop 1 (dfsph.py:418-420): kappa_v = 0.5*max(kappa_v/dt, -0.5 rho0^2)
// BEGIN : begin_*_iter (dfsph.py:442-446, :512-516): alpha /= dt (/dt); kappa(_v) = 0
// REDUCE: the avg_density_err sum of dfsph.py:475-477 / :545-547
// always leaves kfac = alpha * b for the next velocity sweep (and in pos.w when `pack`, single-GPU contexts)
template <int MODE, bool PRE, bool BEGIN, bool REDUCE>
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_dfsph_drho(SweepArgs A, const float4* __restrict__ vel, const float* __restrict__ rho, float* __restrict__ adv_rho,
             float* __restrict__ alpha, float* __restrict__ kap, float* __restrict__ kfac, float lim, int pack) {
    SWEEP_PROLOGUE(A)
    float v[1] = {0.f};
    if (live) {
        const float dt = A.sc->deltaT;
        if (PRE) kap[i] = 0.5f * fmaxf(kap[i] / dt, lim);
        const float3 vi = xyz(vel[i]);
        float sl = 0.f;
        float3 gs = f3(0, 0, 0);
        FOR_LIQUID(A, i, pi, { sl += cubic_gradW_u(K, r2) * dot3(vi - xyz(vel[j]), r); })
        FOR_SOLID(A, i, pi, { gs += r * cubic_gradW_u(K, r2); })
        float s = K.m_l_h * (K.VL0 * sl + (MODE == 0 ? K.VS0 : K.VL0) * dot3(vi, gs));               // Q14
        float b;
        if (MODE == 0) {
            s = fmaxf(s, 0.0f);
            if (A.ncount[i - A.i0 + A.l0] < 20) s = 0.0f;
            adv_rho[i] = s; b = s;
        } else {
            s = fmaxf(1.0f, rho[i] / K.rho0 + dt * s);
            adv_rho[i] = s; b = s - 1.0f;
        }
        float al = alpha[i];
        if (BEGIN) { al = (MODE == 0) ? al / dt : al / dt / dt; alpha[i] = al; kap[i] = 0.0f; }
        kfac[i] = b * al;
        // pack: pos.w carries kfac_j while a correction loop runs, so that the velocity sweep gathers ONE float4 per
        // pair (the readers of pos_j in this launch ignore .w; rho returns to pos.w when the loop ends)
        if (pack) ((float*)A.pos)[4 * (size_t)i + 3] = b * al;
        v[0] = b;
    }
    if (REDUCE) block_partials<1, false>(v, A.partials);
}

// the velocity-correction sweep: MODE 0 warmstart_divergence_vel loop 2 (dfsph.py:422-438),
// 1 divergence_iter loop 1 (:451-473), 2 warmstart_pressure loop 2 (:492-508), 3 pressure_iter loop 1 (:520-543)
// PACK (MODE 1 / 3 on one GPU): kfac_j rides in pos_j.w, one gather per pair instead of two
template <int MODE, bool PACK = false>
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_dfsph_velcorrect(SweepArgs A, float4* __restrict__ vel, const float* __restrict__ adv_rho, float* __restrict__ kap,
                   const float* __restrict__ kap_v, const float* __restrict__ kfac) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float dt = A.sc->deltaT;
    float ki, ks;
    const float* kj_arr;
    if (MODE == 0)      { if (!(adv_rho[i] > 0.0f)) return;   ki = kap[i]; ks = ki; kj_arr = kap; }
    else if (MODE == 2) { if (!(adv_rho[i] > K.rho0)) return; ki = kap[i]; ks = kap_v[i]; kj_arr = kap; }   // Q13
    else                { ki = kfac[i]; kap[i] += ki; ks = ki; kj_arr = kfac; }
    float3 al = f3(0, 0, 0), as = f3(0, 0, 0);
    FOR_LIQUID(A, i, pi, {
        float sum = ki + (PACK ? pj4.w : kj_arr[j]);
        sum = (fabsf(sum) > K.eps) ? sum : 0.0f;
        al += r * (cubic_gradW_u(K, r2) * sum);
    })
    if (fabsf(ki) > K.eps) {
        FOR_SOLID(A, i, pi, { as += r * cubic_gradW_u(K, r2); })
    }
    float3 v = xyz(vel[i]) + al * (dt * K.VL0 * K.m_l_h) + as * (dt * ks * K.VS0 * K.m_l_h);
    vel[i] = f4(v);
}

__global__ void k_kfac(const float* __restrict__ alpha, const float* __restrict__ adv_rho, float* __restrict__ kfac, int NL, float sub,
                       float4* __restrict__ pos, int pack) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float k = (adv_rho[i] - sub) * alpha[i];
    kfac[i] = k;
    if (pack) pos[i].w = k;
}

// end_divergence_iter dfsph.py:481-484
__global__ void k_end_div(float* kappa_v, float* alpha, int NL, const Scalars* sc, float4* pos, const float* rho, int pack) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    kappa_v[i] *= dt; alpha[i] *= dt;
    if (pack) pos[i].w = rho[i];                // pos.w = rho_j again for the viscosity / vorticity / tension gathers
}
// warmstart_pressure loop 1 dfsph.py:489-490
__global__ void k_warm_pressure_kappa(float* kappa, int NL, const Scalars* sc, float lim) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    kappa[i] = fmaxf(kappa[i] / dt / dt, lim);
}
// end_pressure_iter dfsph.py:550-553
__global__ void k_end_pressure(float* kappa, int NL, const Scalars* sc, float4* pos, const float* rho, int pack) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    kappa[i] *= dt * dt;
    if (pack) pos[i].w = rho[i];
}
// clear_nonpressure dfsph.py:334-337
__global__ void k_clear_nonpressure(float4* d_vel, int NL, float gx, float gy, float gz) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NL) d_vel[i] = make_float4(gx, gy, gz, 0.f);
}
// end_viscosity dfsph.py:340-343
__global__ void k_end_viscosity(float4* __restrict__ d_vel, float4* __restrict__ vel_guess, const float4* __restrict__ vel, int NL, const Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    float4 g = vel_guess[i], v = vel[i], a = d_vel[i];
    float3 d = f3(g.x - v.x, g.y - v.y, g.z - v.z);
    d_vel[i] = make_float4(a.x + d.x / dt, a.y + d.y / dt, a.z + d.z / dt, 0.f);
    vel_guess[i] = f4(d);
}

// compute_vorticity dfsph.py:308-331 (Q12: solid omega = vel = 0; per-candidate damping uses
// the reference-exact neighborCount)
struct VortC { float init, visc_omega, coff, c_dmp; const int* sid; int cfl_limit; };
// Q15, as the EXECUTED reference behaves (tests/golden/ref_exec_dfsph*.npz): optimize_time_step launches cfl_time_step(size) for
// size = 1, 2, 4, ... while size < NL (dfsph.py:107-111); the pass with `index` merges pairs at distance index/2, so the last
// merge that would join the two halves (index = smallest power of two >= NL) never runs and vel_max[0] ends up as the maximum
// over the reference indices [0, P), P = largest power of two < NL -- the out-of-bounds reads of the upper blocks never reach
// slot 0.  The engine evaluates exactly that maximum in one reduction (option "cfl_true_max" = 1: over every liquid particle).
static inline int cfl_limit(const wcsph_ctx* c) {
    if (c->cfl_true_max) return 0x7fffffff;
    int P = 1; while (2 * P < c->NL) P *= 2;
    return c->NL >= 2 ? P : 0;
}
__global__ void __launch_bounds__(WCSPH_BLOCK, 3)
k_vorticity(SweepArgs A, VortC V, const float* __restrict__ rho, const float4* __restrict__ vel, const float4* __restrict__ omega,
            float4* __restrict__ d_vel, float4* __restrict__ d_omega) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float dt = A.sc->deltaT;
    const float3 wi = xyz(omega[i]), vi = xyz(vel[i]);
    const float rho_i = rho[i];
    // factored sums of dfsph.py:317-326
    float3 sw = f3(0, 0, 0);        // sum_l (w_i - w_j) W / rho_j
    float3 cwl = f3(0, 0, 0);       // sum_l (w_i - w_j) x gradW
    float3 cvl = f3(0, 0, 0);       // sum_l (v_i - v_j) x gradW
    float3 gs = f3(0, 0, 0);        // sum_s gradW
    FOR_LIQUID(A, i, pi, {
        const float3 g = cubic_gradW(K, r, r2);
        const float3 wij = wi - xyz(omega[j]);
        sw += wij * __fdividef(cubic_W2(K, r2), pj4.w);
        cwl += cross3(wij, g);
        cvl += cross3(vi - xyz(vel[j]), g);
    })
    FOR_SOLID(A, i, pi, { gs += cubic_gradW(K, r, r2); })
    const float c = V.coff / rho_i;
    float3 dw = sw * (-1.0f / dt * V.init * V.visc_omega * K.mass)
              + cvl * (c * V.init * K.mass)
              + cross3(vi, gs) * (c * V.init * K.rho0 * K.VL0)
              + wi * (V.c_dmp * (float)A.ncount[i - A.i0 + A.l0]);      // dfsph.py:326, once per candidate
    float3 dv = xyz(d_vel[i]) + cwl * (c * K.mass) + cross3(wi, gs) * (c * K.rho0 * K.VS0);
    d_omega[i] = f4(dw); d_vel[i] = f4(dv);
}
__global__ void k_omega_update(float4* __restrict__ omega, const float4* __restrict__ d_omega, int NL, const Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    float4 w = omega[i], d = d_omega[i];
    omega[i] = make_float4(w.x + d.x * dt, w.y + d.y * dt, w.z + d.z * dt, 0.f);
}

// cfl_time_step dfsph.py:556-568 as one max reduction (Q15)
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_cfl_max(int NL, Scalars* sc, float* partials, const float4* __restrict__ vel, const float4* __restrict__ d_vel, float* __restrict__ vel_max,
          const int* __restrict__ sid, int limit) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    float v[1] = {-3.4e38f};
    if (i < NL) {
        const float dt = sc->deltaT;
        float4 a = d_vel[i], u = vel[i];
        float x = u.x + a.x * dt, y = u.y + a.y * dt, z = u.z + a.z * dt;
        float m = fmaxf(x * x + y * y + z * z, 0.1f);
        vel_max[i] = m;
        if (sid[i] < limit) v[0] = m;
    }
    block_partials<1, true>(v, partials);
}
__global__ void k_vel_max_slot0(float* vel_max, const int* sid, int NL, const Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < NL && sid[i] == 0) vel_max[i] = sc->vel_max0;
}

// optimize_time_step dfsph.py:113-129 evaluated on the device (same float64 host arithmetic)
__global__ void k_optimize_dt(Scalars* sc, float eps, float radius, float tmax, float tmin) {
    if (threadIdx.x || blockIdx.x) return;
    sc->dt_prev = sc->deltaT;
    double vmax = (double)sc->vel_max0;
    if (vmax > (double)eps) {
        double ts = 0.5 * 0.4 * (double)radius * 2.0 / sqrt(vmax);
        ts = fmin(ts, (double)tmax); ts = fmax(ts, (double)tmin);
        int a = max(sc->pr_iter, sc->vs_iter);
        int it = max(sc->vs_iter, a);                       // Q17
        float d = sc->deltaT;
        sc->dt_prev = d;
        if (it > 10) d = (float)((double)d * 0.9);
        else if (it < 5) d = (float)((double)d * 1.1);
        if ((double)d > ts) d = (float)ts;
        sc->deltaT = d;
    }
}

// update_vel dfsph.py:573-575, update_pos dfsph.py:578-580
__global__ void k_axpy4(float4* __restrict__ y, const float4* __restrict__ x, int NL, const Scalars* sc) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    float4 a = y[i], b = x[i];
    y[i] = make_float4(a.x + b.x * dt, a.y + b.y * dt, a.z + b.z * dt, a.w);
}
__global__ void k_set_iters(Scalars* sc, int vs, int dv, int pr) {
    if (threadIdx.x || blockIdx.x) return;
    if (vs >= 0) sc->vs_iter = vs; if (dv >= 0) sc->dv_iter = dv; if (pr >= 0) sc->pr_iter = pr;
}

// ---- fused-step kernels (wcsph_dfsph_step): same arithmetic per pair, fewer passes -------------
// compute_density + compute_dfsph_coff + warmstart_divergence_vel loop 1 in one sweep
// (dfsph.py:249-262, :346-372, :418-420 + :375-392): all three read only pos / vel of the neighbours
__global__ void __launch_bounds__(WCSPH_BLOCK, WCSPH_MINB)
k_dfsph_head(SweepArgs A, const float4* __restrict__ vel, float* __restrict__ rho, float* __restrict__ alpha,
             float* __restrict__ adv_rho, float* __restrict__ kappa_v, float lim) {
    SWEEP_PROLOGUE(A)
    if (!live) return;
    const float dt = A.sc->deltaT;
    kappa_v[i] = 0.5f * fmaxf(kappa_v[i] / dt, lim);
    const float3 vi = xyz(vel[i]);
    float wl = 0.f, ws = 0.f, g2 = 0.f, sl = 0.f;
    float3 gl = f3(0, 0, 0), gs = f3(0, 0, 0);
    FOR_LIQUID_EXACT(A, i, pi, {
        wl += cubic_W2(K, r2);
        const float gsc = cubic_gradW_s(K, r2);
        g2 += gsc * gsc * r2; gl += r * gsc;
        sl += gsc * dot3(vi - xyz(vel[j]), r);
    })
    FOR_SOLID_EXACT(A, i, pi, {
        ws += cubic_W2(K, r2);
        gs += cubic_gradW(K, r, r2);
    })
    const float d = K.VL0 * K.rho0 * (cubic_W(K, 0.f) + wl) + K.VS0 * K.rhoS0 * ws;
    rho[i] = d;
    ((float*)A.pos)[4 * (size_t)i + 3] = d;
    const float3 sg = gl * K.VL0 + gs * K.VS0;
    const float sgs = K.VL0 * K.VL0 * g2 + dot3(sg, sg);
    alpha[i] = (sgs > K.eps) ? -1.0f / sgs : 0.0f;
    float s = fmaxf(K.VL0 * sl + K.VS0 * dot3(vi, gs), 0.0f);
    if (A.ncount[i - A.i0 + A.l0] < 20) s = 0.0f;
    adv_rho[i] = s;
}

// end_divergence_iter + clear_nonpressure + init_viscosity_para loop 1 (dfsph.py:481-484, :334-337, :199-200)
__global__ void k_post_div(float* __restrict__ kappa_v, float* __restrict__ alpha, float4* __restrict__ d_vel,
                           float4* __restrict__ vel_guess, const float4* __restrict__ vel, int NL, const Scalars* sc,
                           float gx, float gy, float gz, float4* __restrict__ pos, const float* __restrict__ rho, int pack) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    kappa_v[i] *= dt; alpha[i] *= dt;
    if (pack) pos[i].w = rho[i];
    d_vel[i] = make_float4(gx, gy, gz, 0.f);
    float4 g = vel_guess[i], v = vel[i];
    vel_guess[i] = make_float4(g.x + v.x, g.y + v.y, g.z + v.z, 0.f);
}

// end_viscosity + compute_vorticity loop 1 + the cfl maximum (dfsph.py:340-343, :309-327, :556-559)
__global__ void __launch_bounds__(WCSPH_BLOCK, 3)
k_vorticity_fused(SweepArgs A, VortC V, const float* __restrict__ rho, const float4* __restrict__ vel, const float4* __restrict__ omega,
                  float4* __restrict__ vel_guess, float4* __restrict__ d_vel, float4* __restrict__ d_omega, float* __restrict__ vel_max) {
    SWEEP_PROLOGUE(A)
    float vm[1] = {-3.4e38f};
    if (live) {
        const float dt = A.sc->deltaT;
        const float3 wi = xyz(omega[i]), vi = xyz(vel[i]);
        const float rho_i = rho[i];
        const float3 dg = xyz(vel_guess[i]) - vi;                 // end_viscosity
        vel_guess[i] = f4(dg);
        float3 sw = f3(0, 0, 0), cwl = f3(0, 0, 0), cvl = f3(0, 0, 0), gs = f3(0, 0, 0);
        FOR_LIQUID(A, i, pi, {
            const float3 g = cubic_gradW(K, r, r2);
            const float3 wij = wi - xyz(omega[j]);
            sw += wij * __fdividef(cubic_W2(K, r2), pj4.w);
            cwl += cross3(wij, g);
            cvl += cross3(vi - xyz(vel[j]), g);
        })
        FOR_SOLID(A, i, pi, { gs += cubic_gradW(K, r, r2); })
        const float c = V.coff / rho_i;
        const float3 dw = sw * (-1.0f / dt * V.init * V.visc_omega * K.mass)
                        + cvl * (c * V.init * K.mass)
                        + cross3(vi, gs) * (c * V.init * K.rho0 * K.VL0)
                        + wi * (V.c_dmp * (float)A.ncount[i - A.i0 + A.l0]);
        const float3 dv = (xyz(d_vel[i]) + dg / dt) + cwl * (c * K.mass) + cross3(wi, gs) * (c * K.rho0 * K.VS0);
        d_omega[i] = f4(dw); d_vel[i] = f4(dv);
        const float3 u = vi + dv * dt;                            // cfl_time_step(1)
        const float m = fmaxf(dot3(u, u), 0.1f);
        vel_max[i] = m;
        if (V.sid[i] < V.cfl_limit) vm[0] = m;
    }
    block_partials<1, true>(vm, A.partials);
}

// compute_vorticity loop 2 (old dt) + update_vel + warmstart_pressure loop 1 (new dt)
// (dfsph.py:329-330, :573-575, :489-490)
__global__ void k_pre_pressure(float4* __restrict__ omega, const float4* __restrict__ d_omega, float4* __restrict__ vel,
                               const float4* __restrict__ d_vel, float* __restrict__ kappa, int NL, const Scalars* sc, float lim) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT, dt0 = sc->dt_prev;
    float4 w = omega[i], dw = d_omega[i];
    omega[i] = make_float4(w.x + dw.x * dt0, w.y + dw.y * dt0, w.z + dw.z * dt0, 0.f);
    float4 v = vel[i], a = d_vel[i];
    vel[i] = make_float4(v.x + a.x * dt, v.y + a.y * dt, v.z + a.z * dt, v.w);
    kappa[i] = fmaxf(kappa[i] / dt / dt, lim);
}

// end_pressure_iter + update_pos (dfsph.py:550-553, :578-580)
__global__ void k_post_pressure(float* __restrict__ kappa, float4* __restrict__ pos, const float4* __restrict__ vel, int NL, Scalars* sc,
                                const float* __restrict__ rho, int pack) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= NL) return;
    const float dt = sc->deltaT;
    kappa[i] *= dt * dt;
    float4 p = pos[i], v = vel[i];
    p = make_float4(p.x + v.x * dt, p.y + v.y * dt, p.z + v.z * dt, pack ? rho[i] : p.w);
    pos[i] = p;
    if (!(isfinite(p.x) && isfinite(p.y) && isfinite(p.z))) atomicOr((unsigned int*)&sc->flags, WCSPH_FLAG_NAN);    // dfsph.py:645 NaN probe, every particle
}

// ------------------------------------------------------------------------------------------
static float kappa_lim(const wcsph_params& p) { return (float)(-0.5 * (double)p.rho_L0 * (double)p.rho_L0); }

extern "C" int wcsph_dfsph_reset_param(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    c->host_scalars_valid = 0;
    STREAM_LAUNCH(c, k_dfsph_reset, fown<float4>(c, "vel"), fown<float4>(c, "omega"), fown<float>(c, "pressure"),
                  fown<float>(c, "kappa"), fown<float>(c, "kappa_v"), c->nown, c->sc);
    return 0;
}
extern "C" int wcsph_dfsph_compute_density(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    LAUNCH_SWEEP(c, (k_dfsph_density_alpha<true, false>), make_sweep(c), fcur<float>(c, "rho"), fcur<float>(c, "alpha_coff"));
    return 0;
}
extern "C" int wcsph_dfsph_compute_dfsph_coff(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    LAUNCH_SWEEP(c, (k_dfsph_density_alpha<false, true>), make_sweep(c), fcur<float>(c, "rho"), fcur<float>(c, "alpha_coff"));
    return 0;
}
// pos.w carries kfac while a correction loop runs (one gather per pair in the velocity sweep).  z-slab ranks: the ghosts' pos (16 B, w = the
// owner's kfac) travels instead of a 4-byte kfac halo; ghost pos.w is rho again with the next pos halo (viscosity / tension / vorticity /
// the grid build all start with one).
#define PACK_KFAC(c) 1
#define DRHO_ARGS(c, kapname) make_sweep(c), fcur<float4>(c, "vel"), fcur<float>(c, "rho"), fcur<float>(c, "adv_rho"), \
    fcur<float>(c, "alpha_coff"), fcur<float>(c, kapname), fcur<float>(c, "kfac"), kappa_lim((c)->prm), PACK_KFAC(c)
#define VC_ARGS(c, kapname) make_sweep(c), fcur<float4>(c, "vel"), fcur<float>(c, "adv_rho"), fcur<float>(c, kapname), \
    fcur<float>(c, "kappa_v"), fcur<float>(c, "kfac")

extern "C" int wcsph_dfsph_warmstart_divergence_vel(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    LAUNCH_SWEEP_HALO(c, HALO(c, "vel"), (k_dfsph_drho<0, true, false, false>), DRHO_ARGS(c, "kappa_v"));
    LAUNCH_SWEEP_HALO(c, HALO(c, "kappa_v"), k_dfsph_velcorrect<0>, VC_ARGS(c, "kappa_v"));
    return 0;
}
extern "C" int wcsph_dfsph_begin_divergence_iter(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    LAUNCH_SWEEP_HALO(c, HALO(c, "vel"), (k_dfsph_drho<0, false, true, false>), DRHO_ARGS(c, "kappa_v"));
    return 0;
}
static int div_iter(wcsph_ctx* c, bool refresh_kfac) {
    if (refresh_kfac) STREAM_LAUNCH(c, k_kfac, fown<float>(c, "alpha_coff"), fown<float>(c, "adv_rho"), fown<float>(c, "kfac"), c->nown, 0.0f, fown<float4>(c, "pos"), PACK_KFAC(c));
    LAUNCH_SWEEP_HALO(c, HALO(c, "pos"), (k_dfsph_velcorrect<1, true>), VC_ARGS(c, "kappa_v"));
    LAUNCH_SWEEP_HALO_REDUCE(c, HALO(c, "vel"), FIN_AVG_ERR, 0.f, (k_dfsph_drho<0, false, false, true>), DRHO_ARGS(c, "kappa_v"));
    return 0;
}
extern "C" int wcsph_dfsph_divergence_iter(wcsph_ctx* c) { NEED(c, WCSPH_DFSPH); c->host_scalars_valid = 0; return div_iter(c, true); }
extern "C" int wcsph_dfsph_end_divergence_iter(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    STREAM_LAUNCH(c, k_end_div, fown<float>(c, "kappa_v"), fown<float>(c, "alpha_coff"), c->nown, c->sc, fown<float4>(c, "pos"), fown<float>(c, "rho"), PACK_KFAC(c));
    return 0;
}
extern "C" int wcsph_dfsph_clear_nonpressure(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    STREAM_LAUNCH(c, k_clear_nonpressure, fown<float4>(c, "d_vel"), c->nown, c->prm.gravity[0], c->prm.gravity[1], c->prm.gravity[2]);
    return 0;
}
extern "C" int wcsph_dfsph_compute_tension(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    return tension_compute(c);
}
extern "C" int wcsph_dfsph_init_viscosity_para(wcsph_ctx* c) { NEED(c, WCSPH_DFSPH); return visc_init_viscosity_para(c); }
extern "C" int wcsph_dfsph_compute_viscosity_force(wcsph_ctx* c) { NEED(c, WCSPH_DFSPH); return visc_compute_viscosity_force(c); }
extern "C" int wcsph_dfsph_end_viscosity(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    STREAM_LAUNCH(c, k_end_viscosity, fown<float4>(c, "d_vel"), fown<float4>(c, "vel_guess"), fown<float4>(c, "vel"), c->nown, c->sc);
    return 0;
}
extern "C" int wcsph_dfsph_compute_vorticity(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    const wcsph_params& p = c->prm;
    VortC V; V.init = p.vorticity_init; V.visc_omega = p.viscosity_omega; V.coff = p.vorticity_coff;
    V.c_dmp = (float)(-2.0 * (double)p.vorticity_init * (double)p.vorticity_coff);
    V.sid = nullptr; V.cfl_limit = 0;
    LAUNCH_SWEEP_HALO(c, { HALO(c, "pos"); HALO(c, "omega"); HALO(c, "vel"); }, k_vorticity, make_sweep(c), V, fcur<float>(c, "rho"), fcur<float4>(c, "vel"), fcur<float4>(c, "omega"),
                 fcur<float4>(c, "d_vel"), fcur<float4>(c, "d_omega"));
    STREAM_LAUNCH(c, k_omega_update, fown<float4>(c, "omega"), fown<float4>(c, "d_omega"), c->nown, c->sc);
    return 0;
}
extern "C" int wcsph_dfsph_cfl_max(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    STREAM_LAUNCH(c, k_cfl_max, c->nown, c->sc, c->partials, fown<float4>(c, "vel"), fown<float4>(c, "d_vel"), fown<float>(c, "vel_max"),
                  c->sorted_id[c->cur] + c->i0, cfl_limit(c));
    TRY(wcsph_finalize_reduce(c, nblocks(c->nown), FIN_VEL_MAX, 0.f));
    STREAM_LAUNCH(c, k_vel_max_slot0, fown<float>(c, "vel_max"), (c->sorted_id[c->cur] + c->i0), c->nown, c->sc);
    return 0;
}
extern "C" int wcsph_dfsph_update_vel(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    STREAM_LAUNCH(c, k_axpy4, fown<float4>(c, "vel"), fown<float4>(c, "d_vel"), c->nown, c->sc);
    return 0;
}
extern "C" int wcsph_dfsph_warmstart_pressure(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    STREAM_LAUNCH(c, k_warm_pressure_kappa, fown<float>(c, "kappa"), c->nown, c->sc, kappa_lim(c->prm));
    LAUNCH_SWEEP_HALO(c, HALO(c, "kappa"), k_dfsph_velcorrect<2>, VC_ARGS(c, "kappa"));
    return 0;
}
extern "C" int wcsph_dfsph_begin_pressure_iter(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    LAUNCH_SWEEP_HALO(c, HALO(c, "vel"), (k_dfsph_drho<1, false, true, false>), DRHO_ARGS(c, "kappa"));
    return 0;
}
static int pres_iter(wcsph_ctx* c, bool refresh_kfac) {
    if (refresh_kfac) STREAM_LAUNCH(c, k_kfac, fown<float>(c, "alpha_coff"), fown<float>(c, "adv_rho"), fown<float>(c, "kfac"), c->nown, 1.0f, fown<float4>(c, "pos"), PACK_KFAC(c));
    LAUNCH_SWEEP_HALO(c, HALO(c, "pos"), (k_dfsph_velcorrect<3, true>), VC_ARGS(c, "kappa"));
    LAUNCH_SWEEP_HALO_REDUCE(c, HALO(c, "vel"), FIN_AVG_ERR, 0.f, (k_dfsph_drho<1, false, false, true>), DRHO_ARGS(c, "kappa"));
    return 0;
}
extern "C" int wcsph_dfsph_pressure_iter(wcsph_ctx* c) { NEED(c, WCSPH_DFSPH); c->host_scalars_valid = 0; return pres_iter(c, true); }
extern "C" int wcsph_dfsph_end_pressure_iter(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    STREAM_LAUNCH(c, k_end_pressure, fown<float>(c, "kappa"), c->nown, c->sc, fown<float4>(c, "pos"), fown<float>(c, "rho"), PACK_KFAC(c));
    return 0;
}
extern "C" int wcsph_dfsph_update_pos(wcsph_ctx* c) {
    NEED(c, WCSPH_DFSPH);
    STREAM_LAUNCH(c, k_axpy4, fown<float4>(c, "pos"), fown<float4>(c, "vel"), c->nown, c->sc);
    return 0;
}

// ---- loop control on the device (graph mode): the tests of dfsph.py:141, :98, :160 ----------------
// each kernel is one thread; it updates the iteration counter in the scalar block and sets the
// condition of the enclosing WHILE node, with the same float64 arithmetic as the reference's host code
__global__ void k_loop_div_init(Scalars* sc, cudaGraphConditionalHandle h) {
    sc->dv_iter = 0;
    cudaGraphSetConditional(h, ((double)sc->avg_density_err > -0.1) ? 1u : 0u);            // Q16: stale value, err = -0.1
}
__global__ void k_loop_div_test(Scalars* sc, cudaGraphConditionalHandle h, double NLd) {
    const int it = ++sc->dv_iter;
    const double err = 0.001 * NLd / (double)sc->deltaT;
    cudaGraphSetConditional(h, ((double)sc->avg_density_err > err && it < 10) ? 1u : 0u);
}
__global__ void k_loop_vs_init(Scalars* sc, cudaGraphConditionalHandle h) { sc->vs_iter = 0; cudaGraphSetConditional(h, 1u); }
__global__ void k_loop_vs_test(Scalars* sc, cudaGraphConditionalHandle h, double visc_err, double eps) {
    const int it = ++sc->vs_iter;
    const bool stop = ((double)sc->cg_delta <= visc_err * (double)sc->cg_delta_zero) || ((double)sc->cg_delta_zero < eps);
    cudaGraphSetConditional(h, (!stop && it < 100) ? 1u : 0u);
}
__global__ void k_loop_pr_init(Scalars* sc, cudaGraphConditionalHandle h) { sc->pr_iter = 0; cudaGraphSetConditional(h, 1u); }
__global__ void k_loop_pr_test(Scalars* sc, cudaGraphConditionalHandle h, double NLd) {
    const int it = ++sc->pr_iter;
    const double err = (double)sc->avg_density_err / NLd;
    cudaGraphSetConditional(h, ((err > 0.001 || it < 2) && it < 100) ? 1u : 0u);
}
__global__ void k_log_iters(Scalars* sc, int* log, int from_graph) {
    const unsigned int s = sc->step_counter;
    int* e = log + 4 * (s % WCSPH_ITER_LOG);
    e[0] = sc->vs_iter; e[1] = sc->dv_iter; e[2] = sc->pr_iter; e[3] = from_graph;
    sc->step_counter = s + 1;
}

// adds a WHILE node behind the work captured so far on c->stream and redirects capture into its body;
// returns the handle the body's last kernel must set
struct WhileScope { cudaStream_t outer; };
static int while_begin(wcsph_ctx* c, cudaGraphConditionalHandle* h, WhileScope* ws) {
    cudaStreamCaptureStatus st; cudaGraph_t g; const cudaGraphNode_t* deps; size_t ndeps;
    CUDA_TRY(cudaStreamGetCaptureInfo(c->stream, &st, nullptr, &g, &deps, &ndeps));
    if (st != cudaStreamCaptureStatusActive) { wcsph_set_error("while_begin outside capture"); return WCSPH_EINVAL; }
    cudaGraphNodeParams p = { cudaGraphNodeTypeConditional };
    p.conditional.handle = *h; p.conditional.type = cudaGraphCondTypeWhile; p.conditional.size = 1;
    cudaGraphNode_t node;
    CUDA_TRY(cudaGraphAddNode(&node, g, deps, ndeps, &p));
    CUDA_TRY(cudaStreamUpdateCaptureDependencies(c->stream, &node, 1, cudaStreamSetCaptureDependencies));
    ws->outer = c->stream;
    CUDA_TRY(cudaStreamBeginCaptureToGraph(c->cap_stream, p.conditional.phGraph_out[0], nullptr, nullptr, 0, cudaStreamCaptureModeRelaxed));
    c->stream = c->cap_stream;
    return 0;
}
static int while_end(wcsph_ctx* c, WhileScope* ws) {
    CUDA_TRY(cudaStreamEndCapture(c->cap_stream, nullptr));
    c->stream = ws->outer;
    return 0;
}

// the sequence of dfsph.py:606-617.  graph == false: host-driven loops (one pinned read per test);
// graph == true: called under stream capture, loops become conditional WHILE nodes.
static int dfsph_step_sequence(wcsph_ctx* c, bool graph) {
    const double NLd = (double)c->NL;   // GLOBAL liquid count (thresholds of dfsph.py:143,163)
    const wcsph_params& p = c->prm;
    const bool tension = (p.tension_coff != 0.0f || p.tension_coff_b != 0.0f);
    VortC V; V.init = p.vorticity_init; V.visc_omega = p.viscosity_omega; V.coff = p.vorticity_coff;
    V.c_dmp = (float)(-2.0 * (double)p.vorticity_init * (double)p.vorticity_coff);
    V.cfl_limit = cfl_limit(c);
    cudaGraphConditionalHandle hdiv = 0, hvs = 0, hpr = 0;
    cudaGraph_t g = nullptr;
    long long l0 = 0;
    if (graph) {
        cudaStreamCaptureStatus st;
        CUDA_TRY(cudaStreamGetCaptureInfo(c->stream, &st, nullptr, &g, nullptr, nullptr));
        CUDA_TRY(cudaGraphConditionalHandleCreate(&hdiv, g, 0, 0));
        CUDA_TRY(cudaGraphConditionalHandleCreate(&hvs, g, 0, 0));
        CUDA_TRY(cudaGraphConditionalHandleCreate(&hpr, g, 0, 0));
    }
    TRY(wcsph_hashgrid_update_grid(c));
    // compute_density, compute_dfsph_coff, solve_vel_divergence dfsph.py:131-146
    LAUNCH_SWEEP_HALO(c, HALO(c, "vel"), k_dfsph_head, make_sweep(c), fcur<float4>(c, "vel"), fcur<float>(c, "rho"), fcur<float>(c, "alpha_coff"),
                 fcur<float>(c, "adv_rho"), fcur<float>(c, "kappa_v"), kappa_lim(p));
    // (pos.w = rho_j reaches the ghosts with the pos halo of the viscosity / tension sweeps, after the divergence loop has used pos.w for kfac)
    LAUNCH_SWEEP_HALO(c, HALO(c, "kappa_v"), k_dfsph_velcorrect<0>, VC_ARGS(c, "kappa_v"));
    TRY(wcsph_dfsph_begin_divergence_iter(c));
    if (graph) {
        k_loop_div_init<<<1, 1, 0, c->stream>>>(c->sc, hdiv); LAUNCH_CHECK(c);
        WhileScope ws; TRY(while_begin(c, &hdiv, &ws));
        l0 = c->launches;
        TRY(div_iter(c, false));
        k_loop_div_test<<<1, 1, 0, c->stream>>>(c->sc, hdiv, NLd); LAUNCH_CHECK(c);
        c->g_div_body = (int)(c->launches - l0); c->launches = l0;
        TRY(while_end(c, &ws));
    } else {
        c->dv_iter = 0;
        // the first test reads the STALE avg_density_err of the previous pressure solve (Q16) and the current deltaT: the host copy
        // fetched by the last pressure-loop test of the previous fused step still holds both -> no device round trip here
        if (!c->host_scalars_valid) TRY(fetch_scalars(c));
        c->host_scalars_valid = 0;
        double err = -0.1;
        const double dt_np = (double)c->sc_host->deltaT;
        while ((double)c->sc_host->avg_density_err > err && c->dv_iter < 10) {      // Q16: stale first test
            TRY(div_iter(c, false));
            err = 0.001 * NLd / dt_np;
            c->dv_iter++;
            TRY(fetch_scalars(c));
        }
    }
    // end_divergence_iter; compute_nonpressure_force dfsph.py:84-103
    STREAM_LAUNCH(c, k_post_div, fown<float>(c, "kappa_v"), fown<float>(c, "alpha_coff"), fown<float4>(c, "d_vel"),
                  fown<float4>(c, "vel_guess"), fown<float4>(c, "vel"), c->nown, c->sc, p.gravity[0], p.gravity[1], p.gravity[2],
                  fown<float4>(c, "pos"), fown<float>(c, "rho"), PACK_KFAC(c));
    if (tension) TRY(wcsph_dfsph_compute_tension(c));
    if (graph) {
        TRY(visc_init_fused(c));
        k_loop_vs_init<<<1, 1, 0, c->stream>>>(c->sc, hvs); LAUNCH_CHECK(c);
        WhileScope ws; TRY(while_begin(c, &hvs, &ws));
        l0 = c->launches;
        TRY(visc_compute_viscosity_force(c));
        k_loop_vs_test<<<1, 1, 0, c->stream>>>(c->sc, hvs, (double)p.viscosity_err, (double)p.eps); LAUNCH_CHECK(c);
        c->g_vs_body = (int)(c->launches - l0); c->launches = l0;
        TRY(while_end(c, &ws));
    } else {
        TRY(visc_cg_loop(c, true));
    }
    // (vel ghosts are current since the last Drho/Dt sweep)
    V.sid = c->sorted_id[c->cur];             // slot -> reference index: the cfl maximum covers reference indices [0, P), see cfl_limit
    LAUNCH_SWEEP_HALO(c, HALO(c, "omega"), k_vorticity_fused, make_sweep(c), V, fcur<float>(c, "rho"), fcur<float4>(c, "vel"), fcur<float4>(c, "omega"),
                 fcur<float4>(c, "vel_guess"), fcur<float4>(c, "d_vel"), fcur<float4>(c, "d_omega"), fcur<float>(c, "vel_max"));
    TRY(wcsph_finalize_reduce(c, c->sweep_parts, FIN_VEL_MAX, 0.f));
    // optimize_time_step dfsph.py:107-129 (pr_iter is the previous step's, Q17)
    if (!graph) { k_set_iters<<<1, 1, 0, c->stream>>>(c->sc, c->vs_iter, c->dv_iter, c->pr_iter); LAUNCH_CHECK(c); }
    k_optimize_dt<<<1, 1, 0, c->stream>>>(c->sc, p.eps, p.particleRadius, p.user_max_t, p.user_min_t); LAUNCH_CHECK(c);
    // omega update (old dt), update_vel, solve_pressure dfsph.py:150-164
    STREAM_LAUNCH(c, k_pre_pressure, fown<float4>(c, "omega"), fown<float4>(c, "d_omega"), fown<float4>(c, "vel"), fown<float4>(c, "d_vel"),
                  fown<float>(c, "kappa"), c->nown, c->sc, kappa_lim(p));
    LAUNCH_SWEEP_HALO(c, HALO(c, "kappa"), k_dfsph_velcorrect<2>, VC_ARGS(c, "kappa"));
    TRY(wcsph_dfsph_begin_pressure_iter(c));
    if (graph) {
        k_loop_pr_init<<<1, 1, 0, c->stream>>>(c->sc, hpr); LAUNCH_CHECK(c);
        WhileScope ws; TRY(while_begin(c, &hpr, &ws));
        l0 = c->launches;
        TRY(pres_iter(c, false));
        k_loop_pr_test<<<1, 1, 0, c->stream>>>(c->sc, hpr, NLd); LAUNCH_CHECK(c);
        c->g_pr_body = (int)(c->launches - l0); c->launches = l0;
        TRY(while_end(c, &ws));
    } else {
        c->pr_iter = 0;
        double err = 0.0;
        while ((err > 0.001 || c->pr_iter < 2) && c->pr_iter < 100) {
            TRY(pres_iter(c, false));
            c->pr_iter++;
            if (c->pr_iter >= 2) { TRY(fetch_scalars(c)); err = (double)c->sc_host->avg_density_err / NLd; }
        }
        c->host_scalars_valid = 1;          // avg_density_err (final) and deltaT (set by k_optimize_dt above) are what the next step's first test needs
    }
    STREAM_LAUNCH(c, k_post_pressure, fown<float>(c, "kappa"), fown<float4>(c, "pos"), fown<float4>(c, "vel"), c->nown, c->sc, fown<float>(c, "rho"), PACK_KFAC(c));
    if (!graph) { k_set_iters<<<1, 1, 0, c->stream>>>(c->sc, c->vs_iter, c->dv_iter, c->pr_iter); LAUNCH_CHECK(c); }
    k_log_iters<<<1, 1, 0, c->stream>>>(c->sc, c->iter_log, graph ? 1 : 0); LAUNCH_CHECK(c);
    return 0;
}

static int dfsph_build_graph(wcsph_ctx* c, int parity) {
    if (!c->cap_stream) CUDA_TRY(cudaStreamCreateWithFlags(&c->cap_stream, cudaStreamNonBlocking));
    // seed the device copies of the host-side counters (a host-driven step may have run before)
    TRY(wcsph_drain_iter_log(c));
    k_set_iters<<<1, 1, 0, c->stream>>>(c->sc, c->vs_iter, c->dv_iter, c->pr_iter); LAUNCH_CHECK(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    const int cur0 = c->cur;
    const long long l0 = c->launches;
    cudaStream_t user = c->stream;
    c->stream = c->cap_stream;                      // capture on our own stream, launch on the caller's
    // the loop bodies are captured on a second private stream
    cudaStream_t body = nullptr;
    CUDA_TRY(cudaStreamCreateWithFlags(&body, cudaStreamNonBlocking));
    cudaStream_t outer_cap = c->cap_stream;
    c->cap_stream = body;
    cudaError_t e = cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeRelaxed);
    int rc = 0;
    if (e != cudaSuccess) { wcsph_set_error("begin capture: %s", cudaGetErrorString(e)); rc = WCSPH_ECUDA; }
    if (!rc) rc = dfsph_step_sequence(c, true);
    cudaGraph_t g = nullptr;
    e = cudaStreamEndCapture(outer_cap, &g);
    c->stream = user; c->cap_stream = outer_cap;
    cudaStreamDestroy(body);
    c->cur = cur0; c->inv_id_valid = 0;
    c->g_fixed = (int)(c->launches - l0); c->launches = l0;
    if (rc) { if (g) cudaGraphDestroy(g); return rc; }
    if (e != cudaSuccess) { wcsph_set_error("end capture: %s", cudaGetErrorString(e)); return WCSPH_ECUDA; }
    e = cudaGraphInstantiate(&c->step_exec[parity], g, 0);
    if (e != cudaSuccess) { wcsph_set_error("graph instantiate: %s", cudaGetErrorString(e)); cudaGraphDestroy(g); return WCSPH_ECUDA; }
    c->step_graph[parity] = g; c->step_graph_valid[parity] = 1;
    return 0;
}

// dfsph.py:606-617: whole step(s).  Default: one CUDA graph launch per step, the host loops of
// dfsph.py:93-99,141-145,160-163 evaluated on the device with their exact semantics (Q16, Q17), no
// host round trip inside a step.  With option "graph" = 0 or while profiling: the same sequence
// stream-ordered with host-driven loops (one pinned 128-byte read per loop test).
extern "C" int wcsph_dfsph_step(wcsph_ctx* c, int nsteps) {
    NEED(c, WCSPH_DFSPH);
    TRY(wcsph_fatal_flags(c));          // overflow seen by an earlier call: do not keep stepping on dropped pairs
    const bool graph = c->use_graph && c->R == 1 && !(c->prof && c->prof->enabled);
    for (int s = 0; s < nsteps; s++) {
        if (!graph) {
            if (c->graph_pending) TRY(wcsph_drain_iter_log(c));     // pick up counters from earlier graph steps
            TRY(dfsph_step_sequence(c, false));
            continue;
        }
        const int par = c->cur;
        if (!c->step_graph_valid[par]) TRY(dfsph_build_graph(c, par));
        CUDA_TRY(cudaGraphLaunch(c->step_exec[par], c->stream));
        c->cur ^= 1; c->inv_id_valid = 0; c->graph_pending++;
        c->launches += c->g_fixed;
    }
    return 0;
}
