// sweep.cuh -- the neighbour-sweep skeleton shared by every solver.
//
// One thread per cell-sorted liquid particle.  The warp-interleaved compact lists make
// the index loads one 128-byte line per warp per k; the float4 gathers hit L1/L2 (pos and
// vel of 1M particles are 18 MB each, L2 is 126 MB).  Liquid neighbours first, then solid
// neighbours, so the reference's `if j < particleLiquidNum` branch (dfsph.py:258) is two
// loops, not a per-pair branch.
#pragma once
#include "engine.cuh"

struct SweepArgs {
    const float4* pos;
    const uint32_t *nbr_l, *nbr_s;
    const int *nl_cnt, *ns_cnt, *ncount;
    int capL, capS, NL;
    KC k;
    Scalars* sc;
    float* partials;
};

static inline SweepArgs make_sweep(wcsph_ctx* c) {
    SweepArgs a;
    a.pos = fcur<float4>(c, "pos");
    a.nbr_l = c->nbr_l; a.nbr_s = c->nbr_s; a.nl_cnt = c->nl_cnt; a.ns_cnt = c->ns_cnt; a.ncount = c->neighborCount;
    a.capL = c->capL; a.capS = c->capS; a.NL = c->NL;
    a.k = make_kc(c->prm);
    a.sc = c->sc; a.partials = c->partials;
    return a;
}

// for (j in liquid neighbours of i) { r = pos_i - pos_j; r2 = |r|^2; BODY }
#define FOR_LIQUID(A, i, pi, BODY)                                                        \
    {   const uint32_t* row_ = NBR_ROW((A).nbr_l, (A).capL, i);                           \
        const int n_ = (A).nl_cnt[i];                                                     \
        for (int k_ = 0; k_ < n_; k_++) {                                                 \
            const int j = (int)row_[(size_t)k_ * 32];                                     \
            const float4 pj_ = (A).pos[j];                                                \
            const float3 r = f3((pi).x - pj_.x, (pi).y - pj_.y, (pi).z - pj_.z);          \
            const float r2 = dot3(r, r);                                                  \
            BODY                                                                          \
        } }
#define FOR_SOLID(A, i, pi, BODY)                                                         \
    {   const uint32_t* row_ = NBR_ROW((A).nbr_s, (A).capS, i);                           \
        const int n_ = (A).ns_cnt[i];                                                     \
        for (int k_ = 0; k_ < n_; k_++) {                                                 \
            const int j = (int)row_[(size_t)k_ * 32];                                     \
            const float4 pj_ = (A).pos[j];                                                \
            const float3 r = f3((pi).x - pj_.x, (pi).y - pj_.y, (pi).z - pj_.z);          \
            const float r2 = dot3(r, r);                                                  \
            BODY                                                                          \
        } }

#define SWEEP_PROLOGUE(A)                                                                 \
    const int i = blockIdx.x * blockDim.x + threadIdx.x;                                  \
    const bool live = i < (A).NL;                                                         \
    const int ii = live ? i : 0;                                                          \
    const float4 pi4 = (A).pos[ii];                                                       \
    const float3 pi = xyz(pi4);                                                           \
    const KC& K = (A).k;                                                                  \
    (void)K; (void)pi;

#define LAUNCH_SWEEP(c, kern, ...) do { prof_begin(c, #kern); kern<<<nblocks((c)->NL), WCSPH_BLOCK, 0, (c)->stream>>>(__VA_ARGS__); prof_end(c); LAUNCH_CHECK(c); } while (0)
