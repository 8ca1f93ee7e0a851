/*
 * wcsph_oracle.c -- CPU restatement of the lyd405121/wcsph per-step hot path.
 * TEST INFRASTRUCTURE ONLY; pinned on the executed reference (see wcsph_oracle.h for both statements).
 *
 * Conventions used to restate Taichi semantics:
 *  - default_fp=f32: all kernel arithmetic is float; a run of Python-scope
 *    constants at the head of an expression is folded in double and narrowed
 *    once (that is what Taichi sees at trace time), then evaluation proceeds
 *    left to right in float.  Compile with -ffp-contract=off.
 *  - "x[i] += ..." on a field accumulates straight into the float field.
 *  - every top-level `for` of a kernel is one parallel loop with a barrier
 *    after it; OpenMP `parallel for` here (1 thread when used as the checker).
 *  - scalar atomics (avg_density_err, cg dots) are summed in ascending i.
 */
#include "wcsph_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>
#ifdef _OPENMP
#include <omp.h>
#endif

struct Oracle {
    int count, liquid_count, solid_count;
    OracleParams p;
    /* HashGrid.py:10-54 */
    double gridR_d;
    float gridR, invGridR, hash_searchR;
    int maxInGrid, maxNeighbour;
    int blockSize[3];
    float min_boundary[3], max_boundary[3];
    int *gridCount, *grid, *neighborCount, *neighbor;
    int exceed_grid, exceed_neighbor;
    /* ParticleData.py:33-74 (+ solver-local fields) */
    float *pos;                                   /* N x 3 */
    float *vel, *vel_guess, *omega, *d_vel, *d_omega, *normal; /* NL x 3 */
    float *vel_max, *pressure, *rho, *adv_rho;     /* NL */
    float *cg_Minv;                                /* NL x 9 */
    float *cg_r, *cg_dir, *cg_Ad, *cg_s;           /* NL x 3 */
    float avg_density_err, cg_delta, cg_delta_old, cg_delta_zero;
    float *alpha_coff, *kappa, *kappa_v;           /* dfsph.py:46-48 */
    float *a_ii, *d_ii, *dij_pj, *pressure_pre;    /* iisph.py:56-59 */
    float *pos_star, *vel_star, *d_vel_pre;        /* pcisph.py:51-53 */
    float rho_err;                                 /* pcisph.py:57 */
    float deltaT;
    int vs_iter, dv_iter, pr_iter;
    int style; /* 0: CubicKernel class (dfsph), 1: script-inline W (sesph/pcisph/iisph) */
};

static int g_threads = 1;
void oracle_set_threads(int n) {
    g_threads = n > 0 ? n : 1;
#ifdef _OPENMP
    omp_set_num_threads(g_threads);
#endif
}

#define PARFOR _Pragma("omp parallel for schedule(static)")

/* ------------------------------------------------------------------ */
/* smoothing kernels                                                   */
/* ------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;
static inline v3 V(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v3 ld3(const float* a, int i) { return V(a[3*i], a[3*i+1], a[3*i+2]); }
static inline void st3(float* a, int i, v3 v) { a[3*i] = v.x; a[3*i+1] = v.y; a[3*i+2] = v.z; }
static inline v3 add(v3 a, v3 b) { return V(a.x+b.x, a.y+b.y, a.z+b.z); }
static inline v3 sub(v3 a, v3 b) { return V(a.x-b.x, a.y-b.y, a.z-b.z); }
static inline v3 mul(v3 a, float s) { return V(a.x*s, a.y*s, a.z*s); }
static inline v3 divs(v3 a, float s) { return V(a.x/s, a.y/s, a.z/s); }
static inline float dot(v3 a, v3 b) { return a.x*b.x + a.y*b.y + a.z*b.z; }
static inline float nsq(v3 a) { return a.x*a.x + a.y*a.y + a.z*a.z; }
static inline float norm(v3 a) { return sqrtf(nsq(a)); }
static inline v3 cross(v3 a, v3 b) {
    return V(a.y*b.z - a.z*b.y, a.z*b.x - a.x*b.z, a.x*b.y - a.y*b.x);
}

/* CubicKernel.py:44-54 */
static inline float Cubic_W_P(float q) {
    float res = 0.0f;
    if (q <= 1.0f) {
        if (q <= 0.5f) {
            float qq = q*q, qqq = qq*q;
            res = 6.0f*qqq - 6.0f*qq + 1.0f;
        } else {
            float factor = 1.0f - q;
            res = 2.0f*factor*factor*factor;
        }
    }
    return res;
}

/* CubicKernel.py:36-37 (style 0)  |  sesph.py:112-124 (style 1) */
static inline float W_norm_p(const OracleParams* p, float v, int style) {
    if (style == 0) {
        return Cubic_W_P(v / p->searchR) * p->m_k_raw * p->h3inv;
    } else {
        float res = 0.0f;
        float q = v / p->searchR;
        if (q <= 1.0f) {
            if (q <= 0.5f) {
                float qq = q*q, qqq = qq*q;
                res = p->m_k * (6.0f*qqq - 6.0f*qq + 1.0f);
            } else {
                float factor = 1.0f - q;
                res = p->m_k * 2.0f * factor * factor * factor;
            }
        }
        return res;
    }
}

/* CubicKernel.py:21-32 (style 0: constant m_l*h3 folded)  |  sesph.py:97-108 (style 1) */
static inline v3 gradW_p(const OracleParams* p, v3 r, int style) {
    v3 res = V(0.0f, 0.0f, 0.0f);
    float rl = norm(r);
    float q = rl / p->searchR;
    if (rl > 1.0e-5f && q <= 1.0f) {
        v3 gradq = divs(r, rl * p->searchR);
        /* style 0 folds m_l*h3 in double (both Python floats); style 1's m_l
           already contains 1/h^3.  Both arrive here as the float p->m_l. */
        (void)style;
        if (q <= 0.5f) {
            res = mul(gradq, p->m_l * q * (3.0f*q - 2.0f));
        } else {
            float factor = 1.0f - q;
            res = mul(gradq, -p->m_l * (factor*factor));
        }
    }
    return res;
}

/* CohesionKernel.py:18-29 */
static inline float coh_W_norm(const OracleParams* p, float r) {
    float res = 0.0f;
    float radius2 = p->searchR * p->searchR;
    float r2 = r*r;
    if (r2 <= radius2) {
        float r3 = r2*r;
        if (r > 0.5f * p->searchR)
            res = p->coh_m_k * powf(p->searchR - r, 3.0f) * r3;
        else
            res = p->coh_m_k * 2.0f * powf(p->searchR - r, 3.0f) * r3 - p->coh_m_c;
    }
    return res;
}

/* AdhesionKernel.py:21-29 */
static inline float adh_W_norm(const OracleParams* p, float r) {
    float res = 0.0f;
    float radius2 = p->searchR * p->searchR;
    float r2 = r*r;
    if (r2 <= radius2) {
        if (r > 0.5f * p->searchR)
            res = p->adh_m_k * powf(fmaxf(-4.0f*r2 / p->searchR + 6.0f*r - 2.0f*p->searchR, 0.0f), 0.25f);   /* D-TENSION: radicand clamped at 0 (it vanishes at r = h) */
    }
    return res;
}

float oracle_cubic_W_norm(const OracleParams* p, float r, int s) { return W_norm_p(p, r, s); }
void  oracle_cubic_gradW(const OracleParams* p, const float* r, float* out, int s) {
    v3 g = gradW_p(p, V(r[0], r[1], r[2]), s); out[0] = g.x; out[1] = g.y; out[2] = g.z;
}
float oracle_cohesion_W_norm(const OracleParams* p, float r) { return coh_W_norm(p, r); }
float oracle_adhesion_W_norm(const OracleParams* p, float r) { return adh_W_norm(p, r); }

#define Wn(v)    W_norm_p(&o->p, (v), o->style)
#define Wv(r)    W_norm_p(&o->p, norm(r), o->style)
#define GW(r)    gradW_p(&o->p, (r), o->style)

/* ------------------------------------------------------------------ */
/* allocation                                                          */
/* ------------------------------------------------------------------ */
static float* fz(size_t n) { return (float*)calloc(n ? n : 1, sizeof(float)); }

Oracle* oracle_create(int count, int liquid_count, const float* pos,
                      double gridR, int maxInGrid, int maxNeighbour,
                      const float* maxb, const float* minb, const OracleParams* prm)
{
    Oracle* o = (Oracle*)calloc(1, sizeof(Oracle));
    size_t N = (size_t)count, NL = (size_t)liquid_count;
    o->count = count; o->liquid_count = liquid_count; o->solid_count = count - liquid_count;
    o->p = *prm;
    /* HashGrid.py:10-18 */
    o->gridR_d = gridR;
    o->gridR = (float)gridR;
    o->invGridR = (float)(1.0 / gridR);
    o->hash_searchR = (float)(gridR * 2.0);
    o->maxInGrid = maxInGrid; o->maxNeighbour = maxNeighbour;
    /* HashGrid.py:44-52: int((max-min)/gridR + 1); numpy f32 operands, Python float gridR */
    for (int k = 0; k < 3; k++) {
        o->max_boundary[k] = maxb[k]; o->min_boundary[k] = minb[k];
        float d = maxb[k] - minb[k];                /* np.float32 - np.float32 */
        o->blockSize[k] = (int)((double)d / gridR + 1.0);
    }
    o->gridCount = (int*)calloc(N, sizeof(int));
    o->grid = (int*)malloc(N * (size_t)maxInGrid * sizeof(int));
    o->neighborCount = (int*)calloc(NL ? NL : 1, sizeof(int));
    o->neighbor = (int*)malloc((NL ? NL : 1) * (size_t)maxNeighbour * sizeof(int));
    o->pos = (float*)malloc(N * 3 * sizeof(float));
    memcpy(o->pos, pos, N * 3 * sizeof(float));
    o->vel = fz(NL*3); o->vel_guess = fz(NL*3); o->omega = fz(NL*3); o->d_vel = fz(NL*3);
    o->d_omega = fz(NL*3); o->normal = fz(NL*3);
    o->vel_max = fz(NL); o->pressure = fz(NL); o->rho = fz(NL); o->adv_rho = fz(NL);
    o->cg_Minv = fz(NL*9); o->cg_r = fz(NL*3); o->cg_dir = fz(NL*3); o->cg_Ad = fz(NL*3); o->cg_s = fz(NL*3);
    o->alpha_coff = fz(NL); o->kappa = fz(NL); o->kappa_v = fz(NL);
    o->a_ii = fz(NL); o->d_ii = fz(NL*3); o->dij_pj = fz(NL*3); o->pressure_pre = fz(NL);
    o->pos_star = fz(NL*3); o->vel_star = fz(NL*3); o->d_vel_pre = fz(NL*3);
    o->deltaT = 0.001f;
    return o;
}

void oracle_destroy(Oracle* o) {
    if (!o) return;
    free(o->gridCount); free(o->grid); free(o->neighborCount); free(o->neighbor); free(o->pos);
    free(o->vel); free(o->vel_guess); free(o->omega); free(o->d_vel); free(o->d_omega); free(o->normal);
    free(o->vel_max); free(o->pressure); free(o->rho); free(o->adv_rho);
    free(o->cg_Minv); free(o->cg_r); free(o->cg_dir); free(o->cg_Ad); free(o->cg_s);
    free(o->alpha_coff); free(o->kappa); free(o->kappa_v);
    free(o->a_ii); free(o->d_ii); free(o->dij_pj); free(o->pressure_pre);
    free(o->pos_star); free(o->vel_star); free(o->d_vel_pre);
    free(o);
}

void oracle_set_params(Oracle* o, const OracleParams* prm) { o->p = *prm; }

void* oracle_field(Oracle* o, const char* n) {
#define F(x) if (!strcmp(n, #x)) return (void*)o->x;
    F(pos) F(vel) F(vel_guess) F(omega) F(d_vel) F(d_omega) F(normal) F(vel_max) F(pressure) F(rho)
    F(adv_rho) F(cg_Minv) F(cg_r) F(cg_dir) F(cg_Ad) F(cg_s) F(alpha_coff) F(kappa) F(kappa_v)
    F(a_ii) F(d_ii) F(dij_pj) F(pressure_pre) F(pos_star) F(vel_star) F(d_vel_pre)
    F(gridCount) F(grid) F(neighborCount) F(neighbor)
#undef F
#define S(x) if (!strcmp(n, #x)) return (void*)&o->x;
    S(avg_density_err) S(cg_delta) S(cg_delta_old) S(cg_delta_zero) S(rho_err) S(deltaT)
    S(blockSize) S(min_boundary) S(max_boundary) S(style)
#undef S
    return NULL;
}

int oracle_flag(Oracle* o, const char* n) {
    if (!strcmp(n, "exceed_grid")) return o->exceed_grid;
    if (!strcmp(n, "exceed_neighbor")) return o->exceed_neighbor;
    if (!strcmp(n, "vs_iter")) return o->vs_iter;
    if (!strcmp(n, "dv_iter")) return o->dv_iter;
    if (!strcmp(n, "pr_iter")) return o->pr_iter;
    return -1;
}

/* ------------------------------------------------------------------ */
/* HashGrid.py                                                         */
/* ------------------------------------------------------------------ */
/* HashGrid.py:109-114 -- i32 products wrap (two's complement), % is floor-mod */
static inline int get_cell_hash(const Oracle* o, int ax, int ay, int az) {
    int p1 = (int)(73856093u * (unsigned)ax);
    int p2 = (int)(19349663u * (unsigned)ay);
    int p3 = (int)(83492791u * (unsigned)az);
    int n = o->count;
    int m = (p1 ^ p2 ^ p3) % n;          /* C: sign of dividend */
    if (m < 0) m += n;                   /* -> Python/Taichi floor-mod */
    return ((m + n) % n);
}

/* HashGrid.py:118-124 */
static inline int check_in_box(const Oracle* o, int x, int y, int z) {
    return !((x < 0) || (x >= o->blockSize[0]) || (y < 0) || (y >= o->blockSize[1]) ||
             (z < 0) || (z >= o->blockSize[2]));
}

/* HashGrid.py:68,80 -- ti.cast((pos - min) * invGridR, i32): f32 sub, f32 mul, trunc toward 0 */
static inline void cell_of(const Oracle* o, int i, int* c) {
    for (int k = 0; k < 3; k++) {
        float d = o->pos[3*i+k] - o->min_boundary[k];
        float s = d * o->invGridR;
        c[k] = (int)s;
    }
}

/* HashGrid.py:89-106 */
static void insert_neighbor(Oracle* o, int i, int cx, int cy, int cz, int* exceed) {
    if (cx >= 0 && cx < o->blockSize[0] && cy >= 0 && cy < o->blockSize[1] &&
        cz >= 0 && cz < o->blockSize[2]) {
        int h = get_cell_hash(o, cx, cy, cz);
        int k = 0;
        while (k < o->gridCount[h]) {
            int j = o->grid[(size_t)h * o->maxInGrid + k];
            if (j >= 0 && i != j) {
                int old = o->neighborCount[i]++;
                if (old > o->maxNeighbour - 1) (*exceed)++;      /* Q3: counted, entry dropped */
                else o->neighbor[(size_t)i * o->maxNeighbour + old] = j;
            }
            k++;
        }
    }
}

/* HashGrid.py:57-85 */
void hashgrid_update_grid(Oracle* o) {
    int N = o->count, NL = o->liquid_count;
    /* :58-60 */
    PARFOR
    for (int i = 0; i < N; i++) {
        for (int j = 0; j < o->maxInGrid; j++) o->grid[(size_t)i * o->maxInGrid + j] = -1;
        o->gridCount[i] = 0;
    }
    /* :62-64 */
    PARFOR
    for (int i = 0; i < NL; i++) {
        for (int j = 0; j < o->maxNeighbour; j++) o->neighbor[(size_t)i * o->maxNeighbour + j] = -1;
        o->neighborCount[i] = 0;
    }
    /* :67-76 insert pos -- serial, ascending i (one legal atomic order) */
    for (int i = 0; i < N; i++) {
        int c[3]; cell_of(o, i, c);
        if (check_in_box(o, c[0], c[1], c[2]) == 1) {
            int h = get_cell_hash(o, c[0], c[1], c[2]);
            int old = o->gridCount[h]++;
            if (old > o->maxInGrid - 1) { o->exceed_grid++; o->gridCount[h] = o->maxInGrid; } /* Q4 */
            else o->grid[(size_t)h * o->maxInGrid + old] = i;
        }
    }
    /* :79-85 find neighbour */
    int exceed = 0;
    #pragma omp parallel for schedule(static) reduction(+:exceed)
    for (int i = 0; i < NL; i++) {
        int c[3]; cell_of(o, i, c);
        if (check_in_box(o, c[0], c[1], c[2]) == 1) {
            for (int m = -2; m < 3; m++)
                for (int n = -2; n < 3; n++)
                    for (int q = -2; q < 3; q++)
                        insert_neighbor(o, i, c[0]+m, c[1]+n, c[2]+q, &exceed);
        }
    }
    o->exceed_neighbor += exceed;
}

/* neighbour-loop helper: visits exactly the slots a Taichi `while k < cur_neighbor` would,
   skipping slots past the table width (Q3: defined as dropped) */
#define NB_BEGIN(i) { int cur_neighbor = o->neighborCount[i]; \
    int kmax_ = cur_neighbor < o->maxNeighbour ? cur_neighbor : o->maxNeighbour; \
    const int* nb_ = o->neighbor + (size_t)(i) * o->maxNeighbour; \
    for (int k = 0; k < kmax_; k++) { int j = nb_[k];
#define NB_END }}

/* ------------------------------------------------------------------ */
/* sesph.py                                                            */
/* ------------------------------------------------------------------ */
void sesph_reset_param(Oracle* o) {               /* sesph.py:131-136 */
    o->style = 1;
    memset(o->vel, 0, sizeof(float) * 3 * (size_t)o->liquid_count);
    memset(o->pressure, 0, sizeof(float) * (size_t)o->liquid_count);
    o->deltaT = 0.001f;
}

void sesph_update_advection_density(Oracle* o) {  /* sesph.py:139-155 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    PARFOR
    for (int i = 0; i < NL; i++) {
        o->rho[i] = p->VL0 * Wn(0.0f);
        v3 pi = ld3(o->pos, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            float d_den = Wv(r);
            if (j < NL) o->rho[i] += p->VL0 * d_den;
            else        o->rho[i] += p->VS0 * d_den;
        NB_END
        o->rho[i] *= p->rho_L0;
    }
}

void sesph_update_pressure(Oracle* o) {           /* sesph.py:159-166 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    PARFOR
    for (int i = 0; i < NL; i++) {
        o->rho[i] = fmaxf(o->rho[i], p->rho_L0);
        float q = o->rho[i] / p->rho_L0;
        float qq = q*q, qqqq = qq*qq;
        o->pressure[i] = p->stiffness * (qqqq*qq*q - 1.0f);
    }
}

void sesph_compute_force(Oracle* o) {             /* sesph.py:169-189 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    const float c_l = (float)((double)p->dim_coff * (double)p->viscosity * (double)p->liqiudMass);
    const float c_s = (float)((double)p->dim_coff * (double)p->viscosity_b * (double)p->VS0);
    const float h2c = (float)(0.01 * (double)p->searchR * (double)p->searchR);
    const float pl  = (float)(-(double)p->rho_L0 * (double)p->VL0);
    const float ps  = (float)(-(double)p->rho_L0 * (double)p->VS0);
    const float r00 = (float)((double)p->rho_L0 * (double)p->rho_L0);
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 a = V(p->gravity[0], p->gravity[1], p->gravity[2]);
        v3 pi = ld3(o->pos, i), vi = ld3(o->vel, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            v3 grad = GW(r);
            if (j < NL) {
                float s = c_l / o->rho[j] * dot(sub(vi, ld3(o->vel, j)), r) / (nsq(r) + h2c);
                a = add(a, mul(grad, s));
                float pr = pl * (o->pressure[i] / (o->rho[i]*o->rho[i]) + o->pressure[j] / (o->rho[j]*o->rho[j]));
                a = add(a, mul(grad, pr));
            } else {
                float s = c_s * (o->rho[i] / p->rho_L0) * dot(vi, r) / (nsq(r) + h2c);
                a = add(a, mul(grad, s));
                float pr = ps * (o->pressure[i] / (o->rho[i]*o->rho[i]) + o->pressure[i] / r00);   /* Q22 */
                a = add(a, mul(grad, pr));
            }
        NB_END
        st3(o->d_vel, i, a);
    }
}

void sesph_integrator_sesph(Oracle* o) {          /* sesph.py:192-196 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 v = add(ld3(o->vel, i), mul(ld3(o->d_vel, i), dt));
        st3(o->vel, i, v);
        st3(o->pos, i, add(ld3(o->pos, i), mul(v, dt)));
    }
}

void sesph_step(Oracle* o) {                      /* sesph.py:220-225 */
    hashgrid_update_grid(o);
    sesph_update_advection_density(o);
    sesph_update_pressure(o);
    sesph_compute_force(o);
    sesph_integrator_sesph(o);
}

/* ------------------------------------------------------------------ */
/* Weiler-2018 implicit viscosity, shared by dfsph.py:182-246 / iisph.py:185-252 */
/* ------------------------------------------------------------------ */
static inline v3 get_viscosity_Ax(const Oracle* o, const float* x, int i) {
    const OracleParams* p = &o->p; const int NL = o->liquid_count;
    const float c_l = (float)((double)p->dim_coff * (double)p->viscosity * (double)p->liqiudMass);
    const float c_s = (float)((double)p->dim_coff * (double)p->viscosity_b * (double)p->rho_S0);
    const float h2c = (float)(0.01 * (double)p->searchR * (double)p->searchR);
    v3 ret = V(0, 0, 0);
    v3 pi = ld3(o->pos, i), xi = ld3(x, i);
    NB_BEGIN(i)
        v3 r = sub(pi, ld3(o->pos, j));
        if (j < NL) {
            float s = c_l / o->rho[j] * dot(sub(xi, ld3(x, j)), r) / (nsq(r) + h2c);
            ret = add(ret, mul(divs(mul(GW(r), s), o->rho[i]), o->deltaT));
        } else {
            float s = c_s / o->rho[i] * p->VS0 * dot(xi, r) / (nsq(r) + h2c);
            ret = add(ret, mul(divs(mul(GW(r), s), o->rho[i]), o->deltaT));
        }
    NB_END
    return sub(xi, ret);
}

static void inv3(const float* m, float* out) {   /* closed-form adjugate / determinant */
    float a = m[0], b = m[1], c = m[2], d = m[3], e = m[4], f = m[5], g = m[6], h = m[7], k = m[8];
    float A = e*k - f*h, B = -(d*k - f*g), C = d*h - e*g;
    float det = a*A + b*B + c*C;
    float id = 1.0f / det;
    out[0] = A*id;            out[1] = -(b*k - c*h)*id; out[2] = (b*f - c*e)*id;
    out[3] = B*id;            out[4] = (a*k - c*g)*id;  out[5] = -(a*f - c*d)*id;
    out[6] = C*id;            out[7] = -(a*h - b*g)*id; out[8] = (a*e - b*d)*id;
}

static inline v3 matvec(const float* m, v3 v) {
    return V(m[0]*v.x + m[1]*v.y + m[2]*v.z, m[3]*v.x + m[4]*v.y + m[5]*v.z, m[6]*v.x + m[7]*v.y + m[8]*v.z);
}

static void init_viscosity_para(Oracle* o) {     /* dfsph.py:198-223, iisph.py:201-229 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    const float c_l = (float)((double)p->dim_coff * (double)p->viscosity * (double)p->liqiudMass);
    const float c_s = (float)((double)p->dim_coff * (double)p->viscosity_b * (double)p->rho_S0);
    const float h2c = (float)(0.01 * (double)p->searchR * (double)p->searchR);
    PARFOR
    for (int i = 0; i < NL; i++) st3(o->vel_guess, i, add(ld3(o->vel_guess, i), ld3(o->vel, i)));
    PARFOR
    for (int i = 0; i < NL; i++) {
        float m[9] = {0,0,0,0,0,0,0,0,0};
        v3 pi = ld3(o->pos, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            v3 g = GW(r);
            float s;
            if (j < NL) s = c_l / o->rho[j] / (nsq(r) + h2c);
            else        s = c_s / o->rho[i] * p->VS0 / (nsq(r) + h2c);
            m[0] += s*(g.x*r.x); m[1] += s*(g.x*r.y); m[2] += s*(g.x*r.z);
            m[3] += s*(g.y*r.x); m[4] += s*(g.y*r.y); m[5] += s*(g.y*r.z);
            m[6] += s*(g.z*r.x); m[7] += s*(g.z*r.y); m[8] += s*(g.z*r.z);
        NB_END
        float f = o->deltaT / o->rho[i];
        float a[9];
        for (int t = 0; t < 9; t++) a[t] = ((t % 4 == 0) ? 1.0f : 0.0f) - m[t] * f;
        inv3(a, o->cg_Minv + 9*(size_t)i);
    }
    o->cg_delta_zero = 0.0f;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 r = sub(ld3(o->vel, i), get_viscosity_Ax(o, o->vel_guess, i));
        st3(o->cg_r, i, r);
        st3(o->cg_dir, i, matvec(o->cg_Minv + 9*(size_t)i, r));
    }
    float s = 0.0f;
    for (int i = 0; i < NL; i++) s += dot(ld3(o->cg_r, i), ld3(o->cg_dir, i));
    o->cg_delta_zero = s;
    o->cg_delta = o->cg_delta_zero;
}

static void compute_viscosity_force(Oracle* o) { /* dfsph.py:226-246, iisph.py:232-252 */
    const int NL = o->liquid_count;
    float cg_dAd = o->p.eps;
    PARFOR
    for (int i = 0; i < NL; i++) st3(o->cg_Ad, i, get_viscosity_Ax(o, o->cg_dir, i));
    for (int i = 0; i < NL; i++) cg_dAd += dot(ld3(o->cg_dir, i), ld3(o->cg_Ad, i));
    float alpha = o->cg_delta / cg_dAd;
    o->cg_delta_old = o->cg_delta;
    o->cg_delta = 0.0f;
    PARFOR
    for (int i = 0; i < NL; i++) {
        st3(o->vel_guess, i, add(ld3(o->vel_guess, i), mul(ld3(o->cg_dir, i), alpha)));
        v3 r = sub(ld3(o->cg_r, i), mul(ld3(o->cg_Ad, i), alpha));
        st3(o->cg_r, i, r);
        st3(o->cg_s, i, matvec(o->cg_Minv + 9*(size_t)i, r));
    }
    float s = 0.0f;
    for (int i = 0; i < NL; i++) s += dot(ld3(o->cg_r, i), ld3(o->cg_s, i));
    o->cg_delta = s;
    float beta = o->cg_delta / o->cg_delta_old;
    PARFOR
    for (int i = 0; i < NL; i++)
        st3(o->cg_dir, i, add(ld3(o->cg_s, i), mul(ld3(o->cg_dir, i), beta)));
}

static void viscosity_cg_loop(Oracle* o) {       /* dfsph.py:93-99, iisph.py:114-125 */
    init_viscosity_para(o);
    o->vs_iter = 0;
    while (o->vs_iter < 100) {
        compute_viscosity_force(o);
        o->vs_iter++;
        /* Python: f32 field reads -> float64 compare against python floats */
        if ((double)o->cg_delta <= (double)o->p.viscosity_err * (double)o->cg_delta_zero ||
            (double)o->cg_delta_zero < (double)o->p.eps)
            break;
    }
}

/* ------------------------------------------------------------------ */
/* dfsph.py                                                            */
/* ------------------------------------------------------------------ */
void dfsph_reset_param(Oracle* o) {               /* dfsph.py:168-178 */
    size_t NL = (size_t)o->liquid_count;
    o->style = 0;
    memset(o->vel, 0, 12*NL); memset(o->omega, 0, 12*NL);
    memset(o->pressure, 0, 4*NL); memset(o->kappa_v, 0, 4*NL); memset(o->kappa, 0, 4*NL);
    o->deltaT = 0.001f;
}

void dfsph_compute_density(Oracle* o) {           /* dfsph.py:249-262 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    PARFOR
    for (int i = 0; i < NL; i++) {
        o->rho[i] = p->VL0 * Wn(0.0f) * p->rho_L0;
        v3 pi = ld3(o->pos, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            if (j < NL) o->rho[i] += p->VL0 * Wv(r) * p->rho_L0;
            else        o->rho[i] += p->VS0 * Wv(r) * p->rho_S0;
        NB_END
    }
}

/*
 * D-TENSION (deviation).  dfsph.py:265-304 as written scales the
 * normal by searchR once per candidate, overwrites the cohesion term with the
 * curvature term, and restricts adhesion to solids within 0.26 of (0,0.5,0),
 * so its result depends on candidate order.  With tension_coff == 0 (as
 * shipped) it contributes exactly 0 and this restatement matches.  For
 * tension_coff != 0 both this oracle and the CUDA path implement Akinci 2013
 * as the formulas intend:
 *   n_i   = h * sum_{j liquid} m/rho_j gradW_ij
 *   a_i  += sum_{j liquid, |r|<=h} k_ij * ( -g*m*(r/|r|)*C(|r|) - g*(n_i - n_j) ),  k_ij = 2 rho0/(rho_i+rho_j)
 *   a_i  += sum_{j solid, |r|<=h}  -g_b * rho_S0*VS0 * (r/|r|) * A(|r|)
 */
void dfsph_compute_tension(Oracle* o) {
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 n = V(0, 0, 0);
        v3 pi = ld3(o->pos, i);
        NB_BEGIN(i)
            if (j < NL) {
                v3 r = sub(pi, ld3(o->pos, j));
                n = add(n, mul(GW(r), p->liqiudMass / o->rho[j]));
            }
        NB_END
        st3(o->normal, i, mul(n, p->searchR));
    }
    if (p->tension_coff == 0.0f && p->tension_coff_b == 0.0f) return;
    const float sb = (float)((double)p->rho_S0 * (double)p->VS0);
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 a = ld3(o->d_vel, i);
        v3 pi = ld3(o->pos, i), ni = ld3(o->normal, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            float len2 = nsq(r);
            float len = sqrtf(len2);
            if (len / p->searchR > 1.0f) continue;          /* sums run over the support only */
            if (j < NL) {
                float k_ij = 2.0f * p->rho_L0 / (o->rho[i] + o->rho[j]);
                v3 accel = mul(sub(ni, ld3(o->normal, j)), -p->tension_coff);
                if (len2 > p->eps) {
                    v3 xixj = divs(r, len);
                    accel = add(accel, mul(xixj, -p->tension_coff * p->liqiudMass * coh_W_norm(p, len)));
                }
                a = add(a, mul(accel, k_ij));
            } else if (len2 > p->eps) {
                v3 xixj = divs(r, len);
                a = add(a, mul(xixj, -p->tension_coff_b * sb * adh_W_norm(p, len)));
            }
        NB_END
        st3(o->d_vel, i, a);
    }
}

void dfsph_compute_vorticity(Oracle* o) {         /* dfsph.py:308-331 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    const float dt = o->deltaT;
    const float c_dw  = (float)(-1.0);                                   /* -1.0 / deltaT[0] * ... */
    const float c_dmp = (float)(-2.0 * (double)p->vorticity_init * (double)p->vorticity_coff);
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 dw = V(0, 0, 0);
        v3 dv = ld3(o->d_vel, i);
        v3 pi = ld3(o->pos, i), wi = ld3(o->omega, i), vi = ld3(o->vel, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            v3 gradV = GW(r);
            if (j < NL) {
                v3 wij = sub(wi, ld3(o->omega, j));
                float s = c_dw / dt * p->vorticity_init * p->viscosity_omega * (p->liqiudMass / o->rho[j]);
                dw = add(dw, mul(mul(wij, s), Wv(r)));
                dv = add(dv, mul(cross(wij, gradV), p->vorticity_coff / o->rho[i] * p->liqiudMass));
                dw = add(dw, mul(cross(sub(vi, ld3(o->vel, j)), gradV),
                                 p->vorticity_coff / o->rho[i] * p->vorticity_init * p->liqiudMass));
            } else {                                 /* Q12: solid omega = vel = 0 */
                dv = add(dv, mul(cross(wi, gradV), p->vorticity_coff / o->rho[i] * p->rho_L0 * p->VS0));
                dw = add(dw, mul(cross(vi, gradV), p->vorticity_coff / o->rho[i] * p->vorticity_init * p->rho_L0 * p->VL0));
            }
            dw = add(dw, mul(wi, c_dmp));            /* once per candidate (dfsph.py:326) */
        NB_END
        st3(o->d_omega, i, dw);
        st3(o->d_vel, i, dv);
    }
    PARFOR
    for (int i = 0; i < NL; i++)
        st3(o->omega, i, add(ld3(o->omega, i), mul(ld3(o->d_omega, i), dt)));
}

void dfsph_clear_nonpressure(Oracle* o) {         /* dfsph.py:334-337 */
    const int NL = o->liquid_count;
    PARFOR
    for (int i = 0; i < NL; i++) st3(o->d_vel, i, V(o->p.gravity[0], o->p.gravity[1], o->p.gravity[2]));
}

void dfsph_init_viscosity_para(Oracle* o) { init_viscosity_para(o); }
void dfsph_compute_viscosity_force(Oracle* o) { compute_viscosity_force(o); }

void dfsph_end_viscosity(Oracle* o) {             /* dfsph.py:340-343 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 d = sub(ld3(o->vel_guess, i), ld3(o->vel, i));
        st3(o->d_vel, i, add(ld3(o->d_vel, i), divs(d, dt)));
        st3(o->vel_guess, i, d);
    }
}

void dfsph_compute_dfsph_coff(Oracle* o) {        /* dfsph.py:346-372 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 sum_grad = V(0, 0, 0); float sum_grad_square = 0.0f;
        v3 pi = ld3(o->pos, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            v3 gradV = GW(r);
            if (j < NL) {
                v3 temp = mul(gradV, p->VL0);
                sum_grad_square += nsq(temp);
                sum_grad = add(sum_grad, temp);
            } else {
                sum_grad = add(sum_grad, mul(gradV, p->VS0));
            }
        NB_END
        sum_grad_square += nsq(sum_grad);
        o->alpha_coff[i] = (sum_grad_square > p->eps) ? -1.0f / sum_grad_square : 0.0f;
    }
}

static inline void update_drho_divergence(Oracle* o, int i) {  /* dfsph.py:375-392 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    float a = 0.0f; int cnt;
    v3 pi = ld3(o->pos, i), vi = ld3(o->vel, i);
    NB_BEGIN(i)
        v3 r = sub(pi, ld3(o->pos, j));
        v3 gradV = GW(r);
        if (j < NL) a += p->VL0 * dot(sub(vi, ld3(o->vel, j)), gradV);
        else        a += p->VS0 * dot(vi, gradV);
    NB_END
    cnt = o->neighborCount[i];
    a = fmaxf(a, 0.0f);
    if (cnt < 20) a = 0.0f;
    o->adv_rho[i] = a;
}

static inline void update_drho_pressure(Oracle* o, int i) {    /* dfsph.py:395-412 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    float temp = 0.0f;
    v3 pi = ld3(o->pos, i), vi = ld3(o->vel, i);
    NB_BEGIN(i)
        v3 r = sub(pi, ld3(o->pos, j));
        v3 gradV = GW(r);
        if (j < NL) temp += p->VL0 * dot(sub(vi, ld3(o->vel, j)), gradV);
        else        temp += p->VL0 * dot(vi, gradV);                 /* Q14 */
    NB_END
    float a = o->rho[i] / p->rho_L0 + o->deltaT * temp;
    o->adv_rho[i] = fmaxf(1.0f, a);
}

/* the velocity-correction sweep shared by dfsph.py:426-438, :463-473, :497-508, :532-543 */
#define VEL_CORRECT(i, KI, KJ_EXPR, KS)                                        \
    {   v3 v = ld3(o->vel, i); v3 pi = ld3(o->pos, i);                         \
        NB_BEGIN(i)                                                            \
            v3 r = sub(pi, ld3(o->pos, j));                                    \
            v3 gradV = GW(r);                                                  \
            if (j < NL) {                                                      \
                float sum = (KI) + (KJ_EXPR);                                  \
                if (fabsf(sum) > p->eps) v = add(v, mul(gradV, dt * sum * p->VL0)); \
            } else if (fabsf(KI) > p->eps) {                                   \
                v = add(v, mul(gradV, dt * (KS) * p->VS0));                    \
            }                                                                  \
        NB_END                                                                 \
        st3(o->vel, i, v); }

void dfsph_warmstart_divergence_vel(Oracle* o) {  /* dfsph.py:416-438 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p; const float dt = o->deltaT;
    const float lim = (float)(-0.5 * (double)p->rho_L0 * (double)p->rho_L0);
    PARFOR
    for (int i = 0; i < NL; i++) {
        o->kappa_v[i] = 0.5f * fmaxf(o->kappa_v[i] / dt, lim);
        update_drho_divergence(o, i);
    }
    PARFOR
    for (int i = 0; i < NL; i++) {
        if (o->adv_rho[i] > 0.0f) {
            float ki = o->kappa_v[i];
            VEL_CORRECT(i, ki, o->kappa_v[j], ki)
        }
    }
}

void dfsph_begin_divergence_iter(Oracle* o) {     /* dfsph.py:442-446 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        update_drho_divergence(o, i);
        o->alpha_coff[i] = o->alpha_coff[i] / dt;
        o->kappa_v[i] = 0.0f;
    }
}

void dfsph_divergence_iter(Oracle* o) {           /* dfsph.py:450-477 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p; const float dt = o->deltaT;
    o->avg_density_err = 0.0f;
    PARFOR
    for (int i = 0; i < NL; i++) {
        float bi = o->adv_rho[i];
        float ki = bi * o->alpha_coff[i];
        o->kappa_v[i] += ki;
        VEL_CORRECT(i, ki, o->alpha_coff[j] * o->adv_rho[j], ki)
    }
    PARFOR
    for (int i = 0; i < NL; i++) update_drho_divergence(o, i);
    float s = 0.0f;
    for (int i = 0; i < NL; i++) s += o->adv_rho[i];
    o->avg_density_err = s;
}

void dfsph_end_divergence_iter(Oracle* o) {       /* dfsph.py:481-484 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) { o->kappa_v[i] *= dt; o->alpha_coff[i] *= dt; }
}

void dfsph_solve_vel_divergence(Oracle* o) {      /* dfsph.py:131-146 */
    o->dv_iter = 0;
    dfsph_warmstart_divergence_vel(o);
    double err = -0.1;
    dfsph_begin_divergence_iter(o);
    double dt_np = (double)o->deltaT;
    while ((double)o->avg_density_err > err && o->dv_iter < 10) {   /* Q16: stale first test */
        dfsph_divergence_iter(o);
        err = 0.001 * (double)o->liquid_count / dt_np;
        o->dv_iter++;
    }
    dfsph_end_divergence_iter(o);
}

void dfsph_warmstart_pressure(Oracle* o) {        /* dfsph.py:488-508 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p; const float dt = o->deltaT;
    const float lim = (float)(-0.5 * (double)p->rho_L0 * (double)p->rho_L0);
    PARFOR
    for (int i = 0; i < NL; i++) o->kappa[i] = fmaxf(o->kappa[i] / dt / dt, lim);
    PARFOR
    for (int i = 0; i < NL; i++) {
        if (o->adv_rho[i] > p->rho_L0) {              /* Q13: practically never true */
            float ki = o->kappa[i];
            VEL_CORRECT(i, ki, o->kappa[j], o->kappa_v[i])
        }
    }
}

void dfsph_begin_pressure_iter(Oracle* o) {       /* dfsph.py:512-516 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        update_drho_pressure(o, i);
        o->alpha_coff[i] = o->alpha_coff[i] / dt / dt;
        o->kappa[i] = 0.0f;
    }
}

void dfsph_pressure_iter(Oracle* o) {             /* dfsph.py:519-547 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p; const float dt = o->deltaT;
    o->avg_density_err = 0.0f;
    PARFOR
    for (int i = 0; i < NL; i++) {
        float bi = o->adv_rho[i] - 1.0f;
        float ki = bi * o->alpha_coff[i];
        o->kappa[i] += ki;
        VEL_CORRECT(i, ki, (o->adv_rho[j] - 1.0f) * o->alpha_coff[j], ki)
    }
    PARFOR
    for (int i = 0; i < NL; i++) update_drho_pressure(o, i);
    float s = 0.0f;
    for (int i = 0; i < NL; i++) s += o->adv_rho[i] - 1.0f;
    o->avg_density_err = s;
}

void dfsph_end_pressure_iter(Oracle* o) {         /* dfsph.py:550-553 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) o->kappa[i] *= dt * dt;
}

void dfsph_solve_pressure(Oracle* o) {            /* dfsph.py:150-164 */
    dfsph_warmstart_pressure(o);
    o->pr_iter = 0;
    double err = 0.0;
    dfsph_begin_pressure_iter(o);
    while ((err > 0.001 || o->pr_iter < 2) && o->pr_iter < 100) {
        dfsph_pressure_iter(o);
        err = (double)o->avg_density_err / (double)o->liquid_count;
        o->pr_iter++;
    }
    dfsph_end_pressure_iter(o);
}

void dfsph_compute_nonpressure_force(Oracle* o) { /* dfsph.py:84-103 */
    dfsph_clear_nonpressure(o);
    dfsph_compute_tension(o);
    viscosity_cg_loop(o);
    dfsph_end_viscosity(o);
    dfsph_compute_vorticity(o);
}

/* Q15 as the EXECUTED reference behaves (tests/golden/ref_exec_dfsph*.npz): the host loop dfsph.py:107-111 launches
   cfl_time_step(size) for size = 1, 2, 4, ... while size < NL; the pass `index` merges slots at distance index/2, so the merge
   that would join the two halves (index = smallest power of two >= NL) never runs: vel_max[0] = max over [0, P), P = largest
   power of two < NL.  The out-of-bounds reads of dfsph.py:563 only touch the blocks above P.  g_cfl_true_max = 1: all of [0, NL). */
static int g_cfl_true_max = 0;
void oracle_set_cfl_true_max(int on) { g_cfl_true_max = on; }
void dfsph_optimize_time_step(Oracle* o) {        /* dfsph.py:107-129, :556-568 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    int P = 1; while (2 * P < NL) P *= 2;
    if (NL < 2) P = 0;
    if (g_cfl_true_max) P = NL;
    float vmax = 0.0f;
    for (int i = 0; i < NL; i++) {
        float m = fmaxf(nsq(add(ld3(o->vel, i), mul(ld3(o->d_vel, i), dt))), 0.1f);
        o->vel_max[i] = m;
        if (i < P && (i == 0 || m > vmax)) vmax = m;
    }
    if (NL > 0 && P > 0) o->vel_max[0] = vmax;
    if ((double)vmax > (double)o->p.eps) {
        double cfl_factor = 0.5;
        double time_step = cfl_factor * 0.4 * (double)o->p.particleRadius * 2.0 / sqrt((double)vmax);
        if (time_step > (double)o->p.user_max_t) time_step = (double)o->p.user_max_t;
        if (time_step < (double)o->p.user_min_t) time_step = (double)o->p.user_min_t;
        int a = o->pr_iter > o->vs_iter ? o->pr_iter : o->vs_iter;          /* Q17 */
        int iter = o->vs_iter > a ? o->vs_iter : a;
        float d = o->deltaT;
        if (iter > 10)      d = (float)((double)d * 0.9);
        else if (iter < 5)  d = (float)((double)d * 1.1);
        if ((double)d > time_step) d = (float)time_step;
        o->deltaT = d;
    }
}

void dfsph_update_vel(Oracle* o) {                /* dfsph.py:573-575 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) st3(o->vel, i, add(ld3(o->vel, i), mul(ld3(o->d_vel, i), dt)));
}

void dfsph_update_pos(Oracle* o) {                /* dfsph.py:578-580 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) st3(o->pos, i, add(ld3(o->pos, i), mul(ld3(o->vel, i), dt)));
}

void dfsph_step(Oracle* o) {                      /* dfsph.py:606-617 */
    hashgrid_update_grid(o);
    dfsph_compute_density(o);
    dfsph_compute_dfsph_coff(o);
    dfsph_solve_vel_divergence(o);
    dfsph_compute_nonpressure_force(o);
    dfsph_optimize_time_step(o);
    dfsph_update_vel(o);
    dfsph_solve_pressure(o);
    dfsph_update_pos(o);
}

/* ------------------------------------------------------------------ */
/* iisph.py                                                            */
/* ------------------------------------------------------------------ */
void iisph_reset_param(Oracle* o) {               /* iisph.py:178-182 */
    size_t NL = (size_t)o->liquid_count;
    o->style = 1;
    memset(o->vel, 0, 12*NL); memset(o->pressure, 0, 4*NL);
    o->deltaT = 0.001f;
}

void iisph_compute_density(Oracle* o) { dfsph_compute_density(o); }   /* iisph.py:255-268: same statements, inline W */
void iisph_init_viscosity_para(Oracle* o) { init_viscosity_para(o); }
void iisph_compute_viscosity_force(Oracle* o) { compute_viscosity_force(o); }

void iisph_combine_nonpressure(Oracle* o) {       /* iisph.py:271-274 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 d = sub(ld3(o->vel_guess, i), ld3(o->vel, i));
        st3(o->d_vel, i, add(V(o->p.gravity[0], o->p.gravity[1], o->p.gravity[2]), divs(d, dt)));
        st3(o->vel_guess, i, d);
    }
}

void iisph_compute_nonpressure_force(Oracle* o) { /* iisph.py:114-126 */
    viscosity_cg_loop(o);
    iisph_combine_nonpressure(o);
}

void iisph_compute_advection(Oracle* o) {         /* iisph.py:277-316 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 d = V(0, 0, 0);
        st3(o->vel, i, add(ld3(o->vel, i), mul(ld3(o->d_vel, i), dt)));
        v3 pi = ld3(o->pos, i);
        float inv_den = p->rho_L0 / o->rho[i];
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            v3 gradV = GW(r);
            d = add(d, mul(gradV, -p->VL0 * inv_den * inv_den));      /* Q10: VL0 for solids too */
        NB_END
        st3(o->d_ii, i, d);
    }
    PARFOR
    for (int i = 0; i < NL; i++) {
        float aii = 0.0f;
        float density = o->rho[i] / p->rho_L0;
        float adv = density;
        o->pressure_pre[i] = 0.5f * o->pressure[i];
        v3 pi = ld3(o->pos, i), vi = ld3(o->vel, i), dii = ld3(o->d_ii, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            v3 gradV = GW(r);
            if (j < NL) adv += dt * p->VL0 * dot(sub(vi, ld3(o->vel, j)), gradV);
            else        adv += dt * p->VS0 * dot(vi, gradV);
            v3 d_ji = mul(gradV, p->VL0 / (density*density));
            aii += p->VL0 * dot(sub(dii, d_ji), gradV);
        NB_END
        o->a_ii[i] = aii; o->adv_rho[i] = adv;
    }
}

void iisph_update_iter_info(Oracle* o) {          /* iisph.py:319-334 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    o->avg_density_err = 0.0f;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 d = V(0, 0, 0);
        v3 pi = ld3(o->pos, i);
        NB_BEGIN(i)
            if (j < NL) {
                v3 r = sub(pi, ld3(o->pos, j));
                v3 gradV = GW(r);
                float densityj = o->rho[j] / p->rho_L0;
                d = add(d, mul(gradV, -p->VL0 / (densityj*densityj) * o->pressure_pre[j]));
            }
        NB_END
        st3(o->dij_pj, i, d);
    }
}

void iisph_update_pressure_force(Oracle* o) {     /* iisph.py:337-370 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p; const float dt = o->deltaT;
    float* contrib = (float*)calloc((size_t)(NL ? NL : 1), sizeof(float));
    PARFOR
    for (int i = 0; i < NL; i++) {
        float sum = 0.0f;
        v3 pi = ld3(o->pos, i), dpi = ld3(o->dij_pj, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            v3 gradV = GW(r);
            if (j < NL) {
                float density = o->rho[i] / p->rho_L0;
                v3 dji = mul(gradV, p->VL0 / (density*density));
                v3 d_ji_pi = mul(dji, o->pressure_pre[i]);
                v3 d_jk_pk = ld3(o->dij_pj, j);
                v3 t = sub(sub(dpi, mul(ld3(o->d_ii, j), o->pressure_pre[j])), sub(d_jk_pk, d_ji_pi));
                sum += p->VL0 * dot(t, gradV);
            } else {
                sum += p->VS0 * dot(dpi, gradV);
            }
        NB_END
        float b = 1.0f - o->adv_rho[i];
        float h2 = dt * dt;
        float denom = o->a_ii[i] * h2;
        if (fabsf(denom) > p->eps)
            o->pressure[i] = fmaxf((1.0f - p->omega_relax) * o->pressure_pre[i] + p->omega_relax / denom * (b - h2*sum), 0.0f);
        else
            o->pressure[i] = 0.0f;
        if (o->pressure[i] != 0.0f) contrib[i] = (o->a_ii[i]*o->pressure[i] + sum)*h2 - b;
    }
    float s = 0.0f;
    for (int i = 0; i < NL; i++) s += contrib[i];
    o->avg_density_err = s;
    free(contrib);
}

void iisph_solve_pressure(Oracle* o) {            /* iisph.py:130-139 */
    o->pr_iter = 0;
    double err = 0.0;
    while ((err > 0.001 || o->pr_iter < 2) && o->pr_iter < 100) {
        iisph_update_iter_info(o);
        iisph_update_pressure_force(o);
        err = (double)o->avg_density_err / (double)o->liquid_count;
        o->pr_iter++;
    }
}

void iisph_update_pos(Oracle* o) {                /* iisph.py:373-396 */
    const int NL = o->liquid_count; const OracleParams* p = &o->p; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 a = V(0, 0, 0);
        v3 pi = ld3(o->pos, i);
        float density_i = o->rho[i] / p->rho_L0;
        float dpi = o->pressure[i] / (density_i*density_i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            v3 gradV = GW(r);
            if (j < NL) {
                float density_j = o->rho[j] / p->rho_L0;
                float dpj = o->pressure[j] / (density_j*density_j);
                a = add(a, mul(gradV, -p->VL0 * (dpi + dpj)));
            } else {
                a = add(a, mul(gradV, -p->VS0 * dpi));
            }
        NB_END
        st3(o->d_vel, i, a);
    }
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 v = add(ld3(o->vel, i), mul(ld3(o->d_vel, i), dt));
        st3(o->vel, i, v);
        st3(o->pos, i, add(ld3(o->pos, i), mul(v, dt)));
    }
}

void iisph_step(Oracle* o) {                      /* iisph.py:419-427 */
    hashgrid_update_grid(o);
    iisph_compute_density(o);
    iisph_compute_nonpressure_force(o);
    iisph_compute_advection(o);
    iisph_solve_pressure(o);
    iisph_update_pos(o);
}

/* ------------------------------------------------------------------ */
/* pcisph.py                                                           */
/* ------------------------------------------------------------------ */
void pcisph_reset_param(Oracle* o) {              /* pcisph.py:194-197 */
    o->style = 1;
    memset(o->vel, 0, 12*(size_t)o->liquid_count);
    o->deltaT = 0.001f;
}

/* pcisph.py:200-218 with D-PCI (Q24): density sweep first, then the viscosity sweep reads complete rho */
void pcisph_compute_nonpressure_force(Oracle* o) {
    const int NL = o->liquid_count; const OracleParams* p = &o->p;
    const float c_l = (float)((double)p->dim_coff * (double)p->viscosity * (double)p->liqiudMass);
    const float c_s = (float)((double)p->dim_coff * (double)p->viscosity_b * (double)p->VS0);
    const float h2c = (float)(0.01 * (double)p->searchR * (double)p->searchR);
    PARFOR
    for (int i = 0; i < NL; i++) {
        o->rho[i] = p->VL0 * Wn(0.0f) * p->rho_L0;
        v3 pi = ld3(o->pos, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            if (j < NL) o->rho[i] += p->VL0 * Wv(r) * p->rho_L0;
            else        o->rho[i] += p->VS0 * Wv(r) * p->rho_L0;
        NB_END
    }
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 a = V(p->gravity[0], p->gravity[1], p->gravity[2]);
        v3 pi = ld3(o->pos, i), vi = ld3(o->vel, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            if (j < NL) {
                float s = c_l / o->rho[j] * dot(sub(vi, ld3(o->vel, j)), r) / (nsq(r) + h2c);
                a = add(a, mul(GW(r), s));
            } else {
                float s = c_s * (o->rho[i] / p->rho_L0) * dot(vi, r) / (nsq(r) + h2c);
                a = add(a, mul(GW(r), s));
            }
        NB_END
        st3(o->d_vel, i, a);
    }
}

/* BASELINE configs[2]: Akinci tension (D-TENSION, dfsph.py:265-304) on the PCISPH non-pressure acceleration */
void pcisph_compute_tension(Oracle* o) { dfsph_compute_tension(o); }

void pcisph_init_iter_info(Oracle* o) {           /* pcisph.py:221-226 */
    const int NL = o->liquid_count;
    PARFOR
    for (int i = 0; i < NL; i++) {
        st3(o->vel_star, i, ld3(o->vel, i));
        st3(o->pos_star, i, ld3(o->pos, i));
        o->pressure[i] = 0.0f;
        st3(o->d_vel_pre, i, V(0, 0, 0));
    }
}

void pcisph_update_iter_info(Oracle* o) {         /* pcisph.py:229-235 (Q7) */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 vs = add(ld3(o->vel, i), mul(add(ld3(o->d_vel, i), ld3(o->d_vel_pre, i)), dt));
        st3(o->vel_star, i, vs);
        st3(o->pos_star, i, add(ld3(o->pos, i), mul(vs, dt)));
        o->pressure[i] = 0.0f;
    }
    o->rho_err = 0.0f;
}

void pcisph_predict_density(Oracle* o) {          /* pcisph.py:238-278 (Q8: uses pos, not pos_star) */
    const int NL = o->liquid_count; const OracleParams* p = &o->p; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        float a = p->VL0 * Wn(0.0f);
        v3 pi = ld3(o->pos, i);
        NB_BEGIN(i)
            v3 r = sub(pi, ld3(o->pos, j));
            float WW = Wv(r);
            if (j < NL) a += p->VL0 * WW;
            else        a += p->VS0 * WW;
        NB_END
        a = fmaxf(a, 1.0f);
        o->adv_rho[i] = a;
        o->pressure[i] += p->pci_coff * (a - 1.0f) / (dt * dt);
    }
    float s = o->rho_err;
    for (int i = 0; i < NL; i++) s += o->adv_rho[i] - 1.0f;
    o->rho_err = s;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 a = V(0, 0, 0);
        v3 pi = ld3(o->pos, i);
        float dpi = o->pressure[i];
        NB_BEGIN(i)
            v3 pj = (j < NL) ? ld3(o->pos_star, j) : ld3(o->pos, j);
            v3 gradV = GW(sub(pi, pj));
            if (j < NL) a = add(a, mul(gradV, -p->VL0 * (dpi + o->pressure[j])));
            else        a = add(a, mul(gradV, -p->VS0 * dpi));
        NB_END
        st3(o->d_vel_pre, i, a);
    }
}

void pcisph_sovel_pressure(Oracle* o) {           /* pcisph.py:147-157 */
    o->pr_iter = 0;
    double err = 0.0;
    pcisph_init_iter_info(o);
    while ((err > 0.01 || o->pr_iter < 3) && o->pr_iter < 50) {
        pcisph_update_iter_info(o);
        pcisph_predict_density(o);
        err = (double)o->rho_err / (double)o->liquid_count;
        o->pr_iter++;
    }
}

void pcisph_update_pos(Oracle* o) {               /* pcisph.py:282-285 */
    const int NL = o->liquid_count; const float dt = o->deltaT;
    PARFOR
    for (int i = 0; i < NL; i++) {
        v3 v = add(ld3(o->vel, i), mul(add(ld3(o->d_vel, i), ld3(o->d_vel_pre, i)), dt));
        st3(o->vel, i, v);
        st3(o->pos, i, add(ld3(o->pos, i), mul(v, dt)));
    }
}

void pcisph_step(Oracle* o) {                     /* pcisph.py:307-311 */
    hashgrid_update_grid(o);
    pcisph_compute_nonpressure_force(o);
    if (o->p.tension_coff != 0.0f || o->p.tension_coff_b != 0.0f) pcisph_compute_tension(o);
    pcisph_sovel_pressure(o);
    pcisph_update_pos(o);
}

/* test hook: lets a test that drives the host loops itself hand the iteration counters to
   optimize_time_step (dfsph.py:122 reads the module globals) */
void oracle_set_iters(Oracle* o, int vs, int dv, int pr) {
    if (vs >= 0) o->vs_iter = vs;
    if (dv >= 0) o->dv_iter = dv;
    if (pr >= 0) o->pr_iter = pr;
}

/* ---- §8(f) N1: the particle-splat canvas the step loops draw into (Canvas.py:138-209, dfsph.py:585-593,
   sesph.py:201-207).  Serial restatement: the reference's kernel is a parallel loop whose depth test is not
   atomic, so its result is only defined up to that race; the serial order (liquid loop first, then the
   point loop, i ascending, strict '>' depth test) is the semantics the engine is held to. ---- */
static void canvas_matmul4(const float* a, const float* b, float* out) {      /* proj[0] @ view[0], Canvas.py:139 */
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            float acc = a[4 * i] * b[j];
            for (int k = 1; k < 4; k++) acc = acc + a[4 * i + k] * b[4 * k + j];
            out[4 * i + j] = acc;
        }
}
static int canvas_transform(const float* M, const float* v, int sx, int sy, float* out) {   /* Canvas.py:138-141 */
    float s[4];
    for (int i = 0; i < 4; i++) {
        float acc = M[4 * i] * v[0];
        acc = acc + M[4 * i + 1] * v[1];
        acc = acc + M[4 * i + 2] * v[2];
        acc = acc + M[4 * i + 3] * 1.0f;
        s[i] = acc;
    }
    const float x = s[0] / s[3], y = s[1] / s[3], z = s[2] / s[3];
    out[0] = (x + 1.0f) * 0.5f * (float)sx;
    out[1] = (y + 1.0f) * 0.5f * (float)sy;
    out[2] = z;
    /* the i32 cast of a non-finite or huge coordinate is undefined: such particles are not drawn */
    return isfinite(out[0]) && isfinite(out[1]) && fabsf(out[0]) < 1.0e9f && fabsf(out[1]) < 1.0e9f;
}
static void canvas_fill_pixel(float* img, float* depth, int sx, int sy, int px, int py, float z, float c) {   /* Canvas.py:143-148 */
    if (px >= 0 && px < sx && py >= 0 && py < sy) {
        const size_t at = (size_t)px * sy + py;
        if (depth[at] > z) { img[3 * at] = c; img[3 * at + 1] = c; img[3 * at + 2] = c; depth[at] = z; }
    }
}
void oracle_canvas_clear(float* img, float* depth, int sx, int sy) {          /* Canvas.py:205-209 */
    for (size_t at = 0; at < (size_t)sx * sy; at++) { img[3 * at] = img[3 * at + 1] = img[3 * at + 2] = 0.0f; depth[at] = 1.0f; }
}
static void canvas_draw_sphere(float* img, float* depth, int sx, int sy, const float* M, const float* v) {   /* Canvas.py:150-179 */
    float s[3];
    if (!canvas_transform(M, v, sx, sy, s)) return;
    const int xc = (int)s[0], yc = (int)s[1];
    int r = 3, x = 0, y = r, d = 3 - 2 * r;
    while (x <= y) {
        canvas_fill_pixel(img, depth, sx, sy, xc + x, yc + y, s[2], 1.0f);
        canvas_fill_pixel(img, depth, sx, sy, xc - x, yc + y, s[2], 1.0f);
        canvas_fill_pixel(img, depth, sx, sy, xc + x, yc - y, s[2], 1.0f);
        canvas_fill_pixel(img, depth, sx, sy, xc - x, yc - y, s[2], 1.0f);
        canvas_fill_pixel(img, depth, sx, sy, xc + y, yc + x, s[2], 1.0f);
        canvas_fill_pixel(img, depth, sx, sy, xc - y, yc + x, s[2], 1.0f);
        canvas_fill_pixel(img, depth, sx, sy, xc + y, yc - x, s[2], 1.0f);
        canvas_fill_pixel(img, depth, sx, sy, xc - y, yc - x, s[2], 1.0f);
        if (d < 0) d = d + 4 * x + 6;
        else { d = d + 4 * (x - y) + 10; y = y - 1; }
        x += 1;
    }
}
static void canvas_draw_point(float* img, float* depth, int sx, int sy, const float* M, const float* v) {    /* Canvas.py:197-201 */
    float s[3];
    if (!canvas_transform(M, v, sx, sy, s)) return;
    canvas_fill_pixel(img, depth, sx, sy, (int)s[0], (int)s[1], s[2], 0.3f);
}
/* style 0: sesph.py:201-207 (= pcisph.py:288-293, iisph.py:401-406): sphere outline per liquid, point per solid.
   style 1: dfsph.py:585-593: sphere outline per liquid, then a point for EVERY particle. */
void oracle_canvas_draw_particle(const float* pos, int count, int liquid_count, const float* view, const float* proj,
                                 int sx, int sy, int style, float* img, float* depth) {
    float M[16];
    canvas_matmul4(proj, view, M);
    if (style == 0) {
        for (int i = 0; i < count; i++) {
            if (i < liquid_count) canvas_draw_sphere(img, depth, sx, sy, M, pos + 3 * i);
            else canvas_draw_point(img, depth, sx, sy, M, pos + 3 * i);
        }
    } else {
        for (int i = 0; i < liquid_count && i < count; i++) canvas_draw_sphere(img, depth, sx, sy, M, pos + 3 * i);
        for (int i = 0; i < count; i++) canvas_draw_point(img, depth, sx, sy, M, pos + 3 * i);
    }
}

/* ---- §8(f) N2: surface reconstruction on the dense marching-cubes grid (MarchingCubeGrid.py:160-209,262-409).
   Serial restatement with the reference's own structures (gridCount[grid_num], grid[grid_num][maxInGrid]); the
   reference's parallel slot insert and triangle append are atomic, so a serial run (i ascending) is one of its
   legal executions.  Index of node/cell (x,y,z) = x*by*bz + y*bz + z (MarchingCubeGrid.py:371-372). ---- */
typedef struct { float minb[3]; int b[3]; float gridR, invGridR, searchR, isolevel, m_k, h3; int maxInGrid; } McGrid;

static McGrid mc_make(const float* minb, const int* block, double gridR, int maxInGrid) {
    McGrid g;
    for (int k = 0; k < 3; k++) { g.minb[k] = minb[k]; g.b[k] = block[k]; }
    g.gridR = (float)gridR; g.invGridR = (float)(1.0 / gridR);                  /* MarchingCubeGrid.py:22-23 */
    const double sr = gridR * 4.0;                                              /* :25 */
    g.searchR = (float)sr; g.isolevel = 0.5f;                                   /* :27 */
    g.h3 = (float)(1.0 / (sr * sr * sr)); g.m_k = (float)(8.0 / M_PI);          /* CubicKernel.py:14-15 */
    g.maxInGrid = maxInGrid;
    return g;
}
static int mc_in_box(const McGrid* g, int x, int y, int z) {                    /* :355-361 */
    return !(x < 0 || x >= g->b[0] || y < 0 || y >= g->b[1] || z < 0 || z >= g->b[2]);
}
static void mc_node_pos(const McGrid* g, int index, float* p) {                 /* :364-380 */
    const int yz = g->b[1] * g->b[2];
    const int x = index / yz, y = (index % yz) / g->b[2], z = index % g->b[2];
    p[0] = g->minb[0] + (float)x * g->gridR; p[1] = g->minb[1] + (float)y * g->gridR; p[2] = g->minb[2] + (float)z * g->gridR;
}

/* MarchingCubeGrid.py:160-179; returns how often the "mc exceed grid" branch ran */
int oracle_mc_update_grid(const float* pos, int count, const float* minb, const int* block, double gridR, int maxInGrid,
                          int* gridCount, int* grid) {
    const McGrid g = mc_make(minb, block, gridR, maxInGrid);
    const size_t gn = (size_t)g.b[0] * g.b[1] * g.b[2];
    int exceeded = 0;
    for (size_t c = 0; c < gn; c++) { gridCount[c] = 0; for (int k = 0; k < maxInGrid; k++) grid[c * maxInGrid + k] = -1; }
    for (int i = 0; i < count; i++) {
        const int x = (int)((pos[3 * i] - g.minb[0]) * g.invGridR), y = (int)((pos[3 * i + 1] - g.minb[1]) * g.invGridR),
                  z = (int)((pos[3 * i + 2] - g.minb[2]) * g.invGridR);
        if (!mc_in_box(&g, x, y, z)) continue;
        const size_t index = (size_t)x * g.b[1] * g.b[2] + (size_t)y * g.b[2] + z;
        const int old = gridCount[index]++;
        if (old > maxInGrid - 1) { exceeded++; gridCount[index] = maxInGrid; }
        else grid[index * maxInGrid + old] = i;
    }
    return exceeded;
}

/* MarchingCubeGrid.py:183-209 */
void oracle_mc_cal_surface_point(const float* pos, const float* rho, int liquid_count, float liqiudMass,
                                 const float* minb, const int* block, double gridR, int maxInGrid,
                                 const int* gridCount, const int* grid, float* surface_value) {
    const McGrid g = mc_make(minb, block, gridR, maxInGrid);
    const int gn = g.b[0] * g.b[1] * g.b[2], yz = g.b[1] * g.b[2];
    const float w0 = Cubic_W_P(0.0f / g.searchR) * g.m_k * g.h3;                /* Cubic_W_norm(0.0) */
    PARFOR
    for (int i = 0; i < gn; i++) {
        float acc = 0.0f, pi[3];
        mc_node_pos(&g, i, pi);
        const int cx = i / yz, cy = (i % yz) / g.b[2], cz = i % g.b[2];
        for (int m = -4; m < 5; m++) for (int n = -4; n < 5; n++) for (int q = -4; q < 5; q++) {
            if (!mc_in_box(&g, cx + m, cy + n, cz + q)) continue;
            const size_t nei = (size_t)(cx + m) * yz + (size_t)(cy + n) * g.b[2] + (cz + q);
            for (int k = 0; k < gridCount[nei]; k++) {
                const int j = grid[nei * maxInGrid + k];
                const float rx = pi[0] - pos[3 * j], ry = pi[1] - pos[3 * j + 1], rz = pi[2] - pos[3 * j + 2];
                const float W = Cubic_W_P(sqrtf(rx * rx + ry * ry + rz * rz) / g.searchR) * g.m_k * g.h3;
                if (j < liquid_count && W > 0.0f && rho[j] > liqiudMass * w0) acc += liqiudMass / rho[j] * W;
            }
        }
        surface_value[i] = acc;
    }
}

static int mc_check_pos(const float* p2, const float* p1) {                     /* :392-409 */
    int ret = 1;
    if (p2[0] < p1[0]) ret = 1; else if (p2[0] > p1[0]) ret = 0;
    if (p2[1] < p1[1]) ret = 1; else if (p2[1] > p1[1]) ret = 0;
    if (p2[2] < p1[2]) ret = 1; else if (p2[2] > p1[2]) ret = 0;
    return ret;
}
static void mc_vertex_interp(const McGrid* g, const float* a, const float* b, float va, float vb, float* out) {   /* :375-389 */
    const float *p1 = a, *p2 = b;
    if (mc_check_pos(p2, p1) == 1) { const float* t = p1; p1 = p2; p2 = t; const float tv = va; va = vb; vb = tv; }
    for (int k = 0; k < 3; k++) out[k] = p1[k];
    if (fabsf(va - vb) > 0.00001f)
        for (int k = 0; k < 3; k++) out[k] = p1[k] + (p2[k] - p1[k]) / (vb - va) * (g->isolevel - va);
}

/* MarchingCubeGrid.py:262-352.  tritable: i32[256][16], edgetable: i32[256] (MCData.txt).  Returns vertex_count[0]
   (it keeps counting past max_vertex like the reference, :343-349); triangle: f32[max_vertex][3]. */
int oracle_mc_marching_cube(const float* surface_value, const float* minb, const int* block, double gridR,
                            const int* edgetable, const int* tritable, float* triangle, int max_vertex) {
    const McGrid g = mc_make(minb, block, gridR, 4);
    const int gn = g.b[0] * g.b[1] * g.b[2], yz = g.b[1] * g.b[2];
    static const int corner[8][3] = {{0,0,0},{1,0,0},{1,1,0},{0,1,0},{0,0,1},{1,0,1},{1,1,1},{0,1,1}};   /* :271-278 */
    static const int ends[12][2] = {{0,1},{1,2},{2,3},{3,0},{4,5},{5,6},{6,7},{7,4},{0,4},{1,5},{2,6},{3,7}};   /* :307-330 */
    int vertex_count = 0;
    for (int i = 0; i < gn; i++) {
        const int cx = i / yz, cy = (i % yz) / g.b[2], cz = i % g.b[2];
        if (!mc_in_box(&g, cx + 1, cy + 1, cz + 1)) continue;
        int idx[8]; float val[8], p[8][3], vert[12][3];
        int cubeindex = 0;
        for (int c = 0; c < 8; c++) {
            idx[c] = (cx + corner[c][0]) * yz + (cy + corner[c][1]) * g.b[2] + (cz + corner[c][2]);
            val[c] = surface_value[idx[c]];
            mc_node_pos(&g, idx[c], p[c]);
            if (val[c] < g.isolevel) cubeindex |= 1 << c;
        }
        for (int e = 0; e < 12; e++) for (int k = 0; k < 3; k++) vert[e][k] = p[0][k];
        if (edgetable[cubeindex] != 0)
            for (int e = 0; e < 12; e++)
                if (edgetable[cubeindex] & (1 << e)) mc_vertex_interp(&g, p[ends[e][0]], p[ends[e][1]], val[ends[e][0]], val[ends[e][1]], vert[e]);
        for (int k = 0; tritable[cubeindex * 16 + k] != -1; k += 3) {
            const int old = vertex_count; vertex_count += 3;
            if (old < max_vertex)
                for (int t = 0; t < 3; t++) for (int d = 0; d < 3; d++)
                    triangle[3 * (size_t)(old + t) + d] = vert[tritable[cubeindex * 16 + k + t]][d];
        }
    }
    return vertex_count;
}

/* ---- §8(f) N2, second half: the anisotropic-kernel branch (Yu & Turk 2013) the reference keeps but leaves commented out of
   export_surface (MarchingCubeGrid.py:148-149).  RESTATEMENT ONLY: the CUDA engine does not build this branch yet; these
   functions exist so that the next round has its oracle.  They read the HashGrid neighbour table of an Oracle instance
   (oracle_field "neighbor" / "neighborCount"), i.e. every candidate of the 125-bucket walk with its alias duplicates. ---- */

/* ParticleData.py:188-218.  kernel_c = CubicKernel(hash_grid.searchR) (ParticleData.py:31): style 0 of W_norm_p / gradW_p */
void oracle_pd_compute_color_map(const float* pos, const float* rho, int liquid_count, const int* neighborCount, const int* neighbor,
                                 int maxNeighbour, const OracleParams* p, float* color, float* color_grad) {
    PARFOR
    for (int i = 0; i < liquid_count; i++) {
        float c = p->liqiudMass / rho[i] * W_norm_p(p, 0.0f, 0);
        const v3 pi = ld3(pos, i);
        for (int k = 0; k < neighborCount[i]; k++) {
            const int j = neighbor[(size_t)i * maxNeighbour + k];
            const float Wr = W_norm_p(p, norm(sub(pi, ld3(pos, j))), 0);
            if (j < liquid_count) c += p->liqiudMass / rho[j] * Wr;
            else c += p->VS0 * Wr;
        }
        color[i] = c;
    }
    PARFOR
    for (int i = 0; i < liquid_count; i++) {
        v3 g = V(0.0f, 0.0f, 0.0f);
        const v3 pi = ld3(pos, i);
        for (int k = 0; k < neighborCount[i]; k++) {
            const int j = neighbor[(size_t)i * maxNeighbour + k];
            if (j < liquid_count) g = add(g, mul(gradW_p(p, sub(pi, ld3(pos, j)), 0), p->liqiudMass / rho[j] * color[j]));
        }
        st3(color_grad, i, mul(g, 1.0f / color[i]));
    }
}

static float aniso_weight(v3 xi, v3 xj, float mc_searchR) {                     /* ParticleData.py:291-298 */
    const float dis = norm(sub(xi, xj));
    return dis < mc_searchR * 2.0f ? 1.0f - powf(dis / (mc_searchR * 2.0f), 3.0f) : 0.0f;
}

/* eigen-decomposition of a symmetric 3x3 (cyclic Jacobi, double), eigenvalues descending: for the positive semi-definite
   covariance below this IS its SVD (ti.svd, ParticleData.py:270), and R diag(f(sigma)) R^T does not depend on the sign or,
   in a degenerate subspace, the choice of the eigenvectors */
static void sym_eig3(const double A[3][3], double w[3], double Vm[3][3]) {
    double a[3][3];
    for (int r = 0; r < 3; r++) for (int c = 0; c < 3; c++) { a[r][c] = A[r][c]; Vm[r][c] = r == c; }
    for (int sweep = 0; sweep < 64; sweep++) {
        const double off = fabs(a[0][1]) + fabs(a[0][2]) + fabs(a[1][2]);
        if (off < 1e-300 || off < 1e-18 * (fabs(a[0][0]) + fabs(a[1][1]) + fabs(a[2][2]))) break;
        for (int pq = 0; pq < 3; pq++) {
            const int pi_ = pq == 2 ? 1 : 0, q = pq == 0 ? 1 : 2;
            if (fabs(a[pi_][q]) < 1e-300) continue;
            const double th = (a[q][q] - a[pi_][pi_]) / (2.0 * a[pi_][q]);
            const double t = (th >= 0 ? 1.0 : -1.0) / (fabs(th) + sqrt(th * th + 1.0));
            const double c = 1.0 / sqrt(t * t + 1.0), s = t * c;
            for (int k = 0; k < 3; k++) { const double x = a[k][pi_], y = a[k][q]; a[k][pi_] = c * x - s * y; a[k][q] = s * x + c * y; }
            for (int k = 0; k < 3; k++) { const double x = a[pi_][k], y = a[q][k]; a[pi_][k] = c * x - s * y; a[q][k] = s * x + c * y; }
            for (int k = 0; k < 3; k++) { const double x = Vm[k][pi_], y = Vm[k][q]; Vm[k][pi_] = c * x - s * y; Vm[k][q] = s * x + c * y; }
        }
    }
    for (int k = 0; k < 3; k++) w[k] = a[k][k];
    for (int x = 0; x < 2; x++) for (int y = x + 1; y < 3; y++) if (w[y] > w[x]) {
        const double t = w[x]; w[x] = w[y]; w[y] = t;
        for (int k = 0; k < 3; k++) { const double u = Vm[k][x]; Vm[k][x] = Vm[k][y]; Vm[k][y] = u; }
    }
}

/* ParticleData.py:223-285; G: f32[NL][9] row-major */
void oracle_pd_cal_anistropic_kernel(const float* pos, int liquid_count, const int* neighborCount, const int* neighbor,
                                     int maxNeighbour, float mc_searchR, float* pos_avr, float* G) {
    PARFOR
    for (int i = 0; i < liquid_count; i++) {
        float sw = 0.0f; v3 sx = V(0.0f, 0.0f, 0.0f);
        const v3 pi = ld3(pos, i);
        for (int k = 0; k < neighborCount[i]; k++) {
            const int j = neighbor[(size_t)i * maxNeighbour + k];
            if (j < liquid_count) { const float w = aniso_weight(pi, ld3(pos, j), mc_searchR); sw += w; sx = add(sx, mul(ld3(pos, j), w)); }
        }
        st3(pos_avr, i, sw > 0.0f ? mul(sx, 1.0f / sw) : pi);                        /* :240-241 (sum_xj / sum_wij) */
    }
    PARFOR
    for (int i = 0; i < liquid_count; i++) {
        const float kr = 4.0f, ks = 1400.0f, kn = 0.5f, ne = 25.0f;
        float* g = G + 9 * (size_t)i;
        for (int k = 0; k < 9; k++) g[k] = (k % 4 == 0) ? kn : 0.0f;
        if (!((float)neighborCount[i] > ne)) continue;
        float sw = 0.0f, ci[3][3] = {{0}};
        const v3 pi = ld3(pos, i), pa = ld3(pos_avr, i);
        for (int k = 0; k < neighborCount[i]; k++) {
            const int j = neighbor[(size_t)i * maxNeighbour + k];
            if (j >= liquid_count) continue;
            const float w = aniso_weight(pi, ld3(pos, j), mc_searchR);
            const v3 r = sub(ld3(pos, j), pa);
            const float rv[3] = {r.x, r.y, r.z};
            sw += w;
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) ci[a][b] += w * (rv[a] * rv[b]);
        }
        double C[3][3], w3[3], R[3][3];
        for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) C[a][b] = (double)(ci[a][b] / sw);
        sym_eig3(C, w3, R);
        const float s0 = (float)w3[0], s1 = (float)w3[1], s2 = (float)w3[2];
        if (s0 > 0.0f) {                                                             /* :272-279 */
            const float inv[3] = {1.0f / (ks * s0), 1.0f / (ks * fmaxf(s1, s0 / kr)), 1.0f / (ks * fmaxf(s2, s0 / kr))};
            for (int a = 0; a < 3; a++) for (int b = 0; b < 3; b++) {
                double acc = 0.0;
                for (int k = 0; k < 3; k++) acc += R[a][k] * (double)inv[k] * R[b][k];
                g[3 * a + b] = (float)acc;
            }
        }
    }
}

/* MarchingCubeGrid.py:215-243.  The reference reads pos_avr[j] / G[j] before its `j < liquid_count` test, out of bounds for a
   solid j; the restatement tests first. */
void oracle_mc_cal_surface_point_anistropic(const float* pos, const float* pos_avr, const float* G, const float* rho, int liquid_count,
                                            float liqiudMass, const float* minb, const int* block, double gridR, int maxInGrid,
                                            const int* gridCount, const int* grid, float* surface_value) {
    const McGrid g = mc_make(minb, block, gridR, maxInGrid);
    const int gn = g.b[0] * g.b[1] * g.b[2], yz = g.b[1] * g.b[2];
    const float w0 = Cubic_W_P(0.0f / g.searchR) * g.m_k * g.h3;
    PARFOR
    for (int i = 0; i < gn; i++) {
        float acc = 0.0f, pi[3];
        mc_node_pos(&g, i, pi);
        const int cx = i / yz, cy = (i % yz) / g.b[2], cz = i % g.b[2];
        for (int m = -4; m < 5; m++) for (int n = -4; n < 5; n++) for (int q = -4; q < 5; q++) {
            if (!mc_in_box(&g, cx + m, cy + n, cz + q)) continue;
            const size_t nei = (size_t)(cx + m) * yz + (size_t)(cy + n) * g.b[2] + (cz + q);
            for (int k = 0; k < gridCount[nei]; k++) {
                const int j = grid[nei * maxInGrid + k];
                if (j >= liquid_count) continue;
                float r[3], gr[3];
                for (int d = 0; d < 3; d++) r[d] = pi[d] - (0.05f * pos[3 * j + d] + 0.95f * pos_avr[3 * j + d]);
                const float* Gj = G + 9 * (size_t)j;
                for (int a = 0; a < 3; a++) gr[a] = (Gj[3 * a] * r[0] + Gj[3 * a + 1] * r[1] + Gj[3 * a + 2] * r[2]) * 2.0f;
                const float W = Cubic_W_P(sqrtf(gr[0] * gr[0] + gr[1] * gr[1] + gr[2] * gr[2]) / g.searchR) * g.m_k * g.h3;
                if (W > 0.0f && rho[j] > liqiudMass * w0) acc += liqiudMass / rho[j] * W;
            }
        }
        surface_value[i] = acc;
    }
}
