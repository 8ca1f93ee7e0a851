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
    int capL, capS, NL;      // NL = number of OWNED (list-carrying) particles of this rank
    int i0;                  // first owned slot: owned = [i0, i0+NL); ghosts of a z-slab sit around it
    KC k;
    Scalars* sc;
    float* partials;
};

static inline SweepArgs make_sweep(wcsph_ctx* c) {
    SweepArgs a;
    a.pos = fcur<float4>(c, "pos");
    a.nbr_l = c->nbr_l; a.nbr_s = c->nbr_s; a.nl_cnt = c->nl_cnt; a.ns_cnt = c->ns_cnt; a.ncount = c->neighborCount;
    a.capL = c->capL; a.capS = c->capS; a.NL = c->nown; a.i0 = c->i0;
    a.k = make_kc(c->prm);
    a.sc = c->sc; a.partials = c->partials;
    return a;
}

// Neighbour loops.  The list of particle i is padded to a multiple of 4 with the index i itself
// (k_finish_lists): a self pair has r = 0, so gradW = 0 and every gradW-weighted body adds exactly
// nothing -- FOR_LIQUID / FOR_SOLID therefore run whole uint4 groups without a tail predicate.
// Bodies that are NOT gradW-weighted (the W sums of the density kernels) use the _EXACT forms.
//   inside BODY: j (index), pj4 = pos[j] (xyz, w = rho_j), r = pos_i - pos_j, r2 = |r|^2
#define NBR_PAIR_(jj, pi, BODY)                                                           \
    {   const int j = (int)(jj);                                                          \
        const float4 pj4 = (A_POS_)[j];                                                   \
        const float3 r = f3((pi).x - pj4.x, (pi).y - pj4.y, (pi).z - pj4.z);              \
        const float r2 = dot3(r, r);                                                      \
        (void)pj4; BODY }
// the uint4 of group k+1 is requested before group k is processed: the index stream comes from
// HBM (the lists do not fit L2), its latency then overlaps the gathers + math of the current group
#define FOR_NBRS_(ROW4, CNT, pi, BODY)                                                    \
    {   const uint4* row_ = (ROW4);                                                       \
        const int n4_ = ((CNT) + 3) >> 2;                                                 \
        uint4 Jn_ = make_uint4(0u, 0u, 0u, 0u);                                           \
        if (n4_ > 0) Jn_ = __ldg(row_);                                                   \
        for (int k_ = 0; k_ < n4_; k_++) {                                                \
            const uint4 J_ = Jn_;                                                         \
            if (k_ + 1 < n4_) Jn_ = __ldg(row_ + (size_t)(k_ + 1) * 32);                  \
            NBR_PAIR_(J_.x, pi, BODY) NBR_PAIR_(J_.y, pi, BODY)                           \
            NBR_PAIR_(J_.z, pi, BODY) NBR_PAIR_(J_.w, pi, BODY)                           \
        } }
#define FOR_NBRS_EXACT_(ROW4, CNT, pi, BODY)                                              \
    {   const uint4* row_ = (ROW4);                                                       \
        const int n_ = (CNT);                                                             \
        uint4 Jn_ = make_uint4(0u, 0u, 0u, 0u);                                           \
        if (n_ > 0) Jn_ = __ldg(row_);                                                    \
        for (int k_ = 0; k_ < n_; k_ += 4) {                                              \
            const uint4 J_ = Jn_;                                                         \
            if (k_ + 4 < n_) Jn_ = __ldg(row_ + (size_t)((k_ >> 2) + 1) * 32);            \
            NBR_PAIR_(J_.x, pi, BODY)                                                     \
            if (k_ + 1 < n_) NBR_PAIR_(J_.y, pi, BODY)                                    \
            if (k_ + 2 < n_) NBR_PAIR_(J_.z, pi, BODY)                                    \
            if (k_ + 3 < n_) NBR_PAIR_(J_.w, pi, BODY)                                    \
        } }
#define FOR_LIQUID(A, i, pi, BODY) { const float4* A_POS_ = (A).pos; FOR_NBRS_(NBR_ROW4((A).nbr_l, (A).capL, (i) - (A).i0), (A).nl_cnt[(i) - (A).i0], pi, BODY) }
#define FOR_SOLID(A, i, pi, BODY)  { const float4* A_POS_ = (A).pos; FOR_NBRS_(NBR_ROW4((A).nbr_s, (A).capS, (i) - (A).i0), (A).ns_cnt[(i) - (A).i0], pi, BODY) }
#define FOR_LIQUID_EXACT(A, i, pi, BODY) { const float4* A_POS_ = (A).pos; FOR_NBRS_EXACT_(NBR_ROW4((A).nbr_l, (A).capL, (i) - (A).i0), (A).nl_cnt[(i) - (A).i0], pi, BODY) }
#define FOR_SOLID_EXACT(A, i, pi, BODY)  { const float4* A_POS_ = (A).pos; FOR_NBRS_EXACT_(NBR_ROW4((A).nbr_s, (A).capS, (i) - (A).i0), (A).ns_cnt[(i) - (A).i0], pi, BODY) }

#define SWEEP_PROLOGUE(A)                                                                 \
    const int li_ = blockIdx.x * blockDim.x + threadIdx.x;                                \
    const bool live = li_ < (A).NL;                                                       \
    const int i = (A).i0 + (live ? li_ : 0);                                              \
    const float4 pi4 = (A).pos[i];                                                        \
    const float3 pi = xyz(pi4);                                                           \
    const KC& K = (A).k;                                                                  \
    (void)K; (void)pi;

#define LAUNCH_SWEEP(c, kern, ...) do { prof_begin(c, #kern); kern<<<nblocks((c)->nown), WCSPH_BLOCK, 0, (c)->stream>>>(__VA_ARGS__); prof_end(c); LAUNCH_CHECK(c); } while (0)

// a sweep that ends in a global reduction: launch + one-block finalize
#define LAUNCH_SWEEP_REDUCE(c, op, eps, kern, ...) do { LAUNCH_SWEEP(c, kern, __VA_ARGS__); TRY(wcsph_finalize_reduce(c, nblocks((c)->nown), op, eps)); } while (0)
