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
    int i0;                  // first slot this launch covers: [i0, i0+NL); ghosts of a z-slab sit around the owned range
    int l0;                  // owned ordinal of slot i0 (lists / counts are indexed by owned ordinal)
    int gap_at, gap_len;     // the launch covers [0, gap_at) and [gap_at + gap_len, ...) of its range: both boundary strips of a slab in ONE launch
    KC k;
    Scalars* sc;
    float* partials;
};

static inline SweepArgs make_sweep(wcsph_ctx* c) {
    SweepArgs a;
    a.pos = fcur<float4>(c, "pos");
    a.nbr_l = c->nbr_l; a.nbr_s = c->nbr_s; a.nl_cnt = c->nl_cnt; a.ns_cnt = c->ns_cnt; a.ncount = c->neighborCount;
    a.capL = c->capL; a.capS = c->capS;
    a.l0 = c->sub_active ? c->sub_off : 0;
    a.i0 = c->i0 + a.l0;
    a.NL = c->sub_active ? c->sub_n : c->nown;
    a.gap_at = c->sub_active ? c->sub_gap_at : 0x7fffffff; a.gap_len = c->sub_active ? c->sub_gap_len : 0;
    a.k = make_kc(c->prm);
    a.sc = c->sc; a.partials = c->partials + c->part_off;
    return a;
}

// Neighbour loops.  The list of particle i is padded to a multiple of 4 with the index i itself
// (k_finish_lists): a self pair has r = 0, so gradW = 0 and every gradW-weighted body adds exactly
// nothing -- FOR_LIQUID / FOR_SOLID therefore run whole uint4 groups without a tail predicate.
// Bodies that are NOT gradW-weighted (the W sums of the density kernels) use the _EXACT forms.
//   inside BODY: j (index), pj4 = pos[j] (xyz, w = rho_j), r = pos_i - pos_j, r2 = |r|^2
#define NBR_PAIR_(jj, pi, BODY)                                                           \
    {   const unsigned int j = (jj);                                                      \
        const float4 pj4 = (A_POS_)[j];                                                   \
        const float3 r = f3((pi).x - pj4.x, (pi).y - pj4.y, (pi).z - pj4.z);              \
        const float r2 = dot3(r, r);                                                      \
        (void)pj4; BODY }
// the uint4 of group k+1 is requested before group k is processed: the index stream comes from
// HBM (the lists do not fit L2), its latency then overlaps the gathers + math of the current group
#ifndef WCSPH_GROUP8
#define WCSPH_GROUP8 0
#endif
#if WCSPH_GROUP8
// eight neighbours (two uint4) per iteration: eight independent gathers in flight per thread -- the sweeps
// are bound by the latency of those gathers (ncu: long_scoreboard), not by issue slots.  Lists are padded
// to a multiple of 8 with the particle's own index (k_finish_lists).
#define FOR_NBRS_(ROW4, CNT, pi, BODY)                                                    \
    {   const uint4* row_ = (ROW4);                                                       \
        const int n8_ = ((CNT) + 7) >> 3;                                                 \
        uint4 Ja_ = make_uint4(0u, 0u, 0u, 0u), Jb_ = Ja_;                                \
        if (n8_ > 0) { Ja_ = __ldcs(row_); Jb_ = __ldcs(row_ + 32); }                     \
        for (int k_ = 0; k_ < n8_; k_++) {                                                \
            const uint4 J_ = Ja_, K_ = Jb_;                                               \
            if (k_ + 1 < n8_) { Ja_ = __ldcs(row_ + (size_t)(2 * k_ + 2) * 32); Jb_ = __ldcs(row_ + (size_t)(2 * k_ + 3) * 32); } \
            NBR_PAIR_(J_.x, pi, BODY) NBR_PAIR_(J_.y, pi, BODY)                           \
            NBR_PAIR_(J_.z, pi, BODY) NBR_PAIR_(J_.w, pi, BODY)                           \
            NBR_PAIR_(K_.x, pi, BODY) NBR_PAIR_(K_.y, pi, BODY)                           \
            NBR_PAIR_(K_.z, pi, BODY) NBR_PAIR_(K_.w, pi, BODY)                           \
        } }
#else
#define FOR_NBRS_(ROW4, CNT, pi, BODY)                                                    \
    {   const uint4* row_ = (ROW4);                                                       \
        const int n4_ = ((CNT) + 3) >> 2;                                                 \
        uint4 Jn_ = make_uint4(0u, 0u, 0u, 0u);                                           \
        if (n4_ > 0) Jn_ = __ldcs(row_);                                                   \
        for (int k_ = 0; k_ < n4_; k_++) {                                                \
            const uint4 J_ = Jn_;                                                         \
            if (k_ + 1 < n4_) Jn_ = __ldcs(row_ + (size_t)(k_ + 1) * 32);                  \
            NBR_PAIR_(J_.x, pi, BODY) NBR_PAIR_(J_.y, pi, BODY)                           \
            NBR_PAIR_(J_.z, pi, BODY) NBR_PAIR_(J_.w, pi, BODY)                           \
        } }
#endif
#define FOR_NBRS_EXACT_(ROW4, CNT, pi, BODY)                                              \
    {   const uint4* row_ = (ROW4);                                                       \
        const int n_ = (CNT);                                                             \
        uint4 Jn_ = make_uint4(0u, 0u, 0u, 0u);                                           \
        if (n_ > 0) Jn_ = __ldcs(row_);                                                    \
        for (int k_ = 0; k_ < n_; k_ += 4) {                                              \
            const uint4 J_ = Jn_;                                                         \
            if (k_ + 4 < n_) Jn_ = __ldcs(row_ + (size_t)((k_ >> 2) + 1) * 32);            \
            NBR_PAIR_(J_.x, pi, BODY)                                                     \
            if (k_ + 1 < n_) NBR_PAIR_(J_.y, pi, BODY)                                    \
            if (k_ + 2 < n_) NBR_PAIR_(J_.z, pi, BODY)                                    \
            if (k_ + 3 < n_) NBR_PAIR_(J_.w, pi, BODY)                                    \
        } }
#define FOR_LIQUID(A, i, pi, BODY) { const float4* A_POS_ = (A).pos; FOR_NBRS_(NBR_ROW4((A).nbr_l, (A).capL, (i) - (A).i0 + (A).l0), (A).nl_cnt[(i) - (A).i0 + (A).l0], pi, BODY) }
#define FOR_SOLID(A, i, pi, BODY)  { const float4* A_POS_ = (A).pos; FOR_NBRS_(NBR_ROW4((A).nbr_s, (A).capS, (i) - (A).i0 + (A).l0), (A).ns_cnt[(i) - (A).i0 + (A).l0], pi, BODY) }
#define FOR_LIQUID_EXACT(A, i, pi, BODY) { const float4* A_POS_ = (A).pos; FOR_NBRS_EXACT_(NBR_ROW4((A).nbr_l, (A).capL, (i) - (A).i0 + (A).l0), (A).nl_cnt[(i) - (A).i0 + (A).l0], pi, BODY) }
#define FOR_SOLID_EXACT(A, i, pi, BODY)  { const float4* A_POS_ = (A).pos; FOR_NBRS_EXACT_(NBR_ROW4((A).nbr_s, (A).capS, (i) - (A).i0 + (A).l0), (A).ns_cnt[(i) - (A).i0 + (A).l0], pi, BODY) }

#define SWEEP_PROLOGUE(A)                                                                 \
    const int li_ = blockIdx.x * blockDim.x + threadIdx.x;                                \
    const bool live = li_ < (A).NL;                                                       \
    const int i = (A).i0 + (live ? li_ + (li_ >= (A).gap_at ? (A).gap_len : 0) : 0);      \
    const float4 pi4 = (A).pos[i];                                                        \
    const float3 pi = xyz(pi4);                                                           \
    const KC& K = (A).k;                                                                  \
    (void)K; (void)pi;

#define SWEEP_N(c) ((c)->sub_active ? (c)->sub_n : (c)->nown)
#define LAUNCH_SWEEP(c, kern, ...) do { prof_begin(c, #kern); kern<<<nblocks(SWEEP_N(c)), WCSPH_BLOCK, 0, (c)->stream>>>(__VA_ARGS__); prof_end(c); LAUNCH_CHECK(c); } while (0)

// A sweep that gathers halo'd fields on a z-slab rank.  Owned particles are z-sorted, so the ones whose
// stencil reaches a ghost layer are a prefix (lowest two layers) and a suffix (highest two, + the
// out-of-box tail) of the owned range.  The halo exchange (HALOS) runs on the high-priority side stream and
// the two boundary strips follow it THERE, in one launch, while the main stream sweeps the interior range:
// the strips (a few per cent of the particles, one thin wave of CTAs) fill in between the interior's CTAs
// instead of running as a tail of their own after it.  One GPU: plain launch.
#define LAUNCH_SWEEP_HALO(c, HALOS, kern, ...) do {                                                    \
    const int nlo_ = (c)->n_send_lo, nmid_ = (c)->n_inbox - (c)->n_send_hi - (c)->n_send_lo;              \
    if ((c)->R <= 1) { LAUNCH_SWEEP(c, kern, __VA_ARGS__); (c)->sweep_parts = nblocks((c)->nown); }        \
    else if (nmid_ <= 0 || !(c)->halo_overlap) { TRY(wcsph_halo_group(c, 1)); HALOS; TRY(wcsph_halo_group(c, 0)); LAUNCH_SWEEP(c, kern, __VA_ARGS__); (c)->sweep_parts = nblocks((c)->nown); } \
    else {                                                                                             \
        TRY(wcsph_halo_begin(c)); TRY(wcsph_halo_group(c, 1)); HALOS; TRY(wcsph_halo_group(c, 0));      \
        /* side stream, behind the halo: both boundary strips ([0, nlo) and [nlo + nmid, nown)) in one launch */ \
        (c)->sub_active = 1; (c)->part_off = nblocks(nmid_);                                           \
        (c)->sub_off = 0; (c)->sub_n = (c)->nown - nmid_; (c)->sub_gap_at = nlo_; (c)->sub_gap_len = nmid_; \
        if ((c)->sub_n > 0) { LAUNCH_SWEEP(c, kern, __VA_ARGS__); }                                     \
        (c)->sweep_parts = (c)->part_off + ((c)->sub_n > 0 ? nblocks((c)->sub_n) : 0);                  \
        TRY(wcsph_halo_end(c));                                                                        \
        /* main stream: the interior range, concurrently */                                            \
        (c)->part_off = 0; (c)->sub_gap_at = 0x7fffffff; (c)->sub_gap_len = 0;                          \
        (c)->sub_off = nlo_; (c)->sub_n = nmid_;                                                       \
        LAUNCH_SWEEP(c, kern, __VA_ARGS__);                                                            \
        TRY(wcsph_halo_wait(c));                                                                       \
        (c)->sub_active = 0; (c)->part_off = 0;                                                        \
    } } while (0)

// a sweep that ends in a global reduction: launch + one-block finalize
#define LAUNCH_SWEEP_REDUCE(c, op, eps, kern, ...) do { LAUNCH_SWEEP(c, kern, __VA_ARGS__); TRY(wcsph_finalize_reduce(c, nblocks((c)->nown), op, eps)); } while (0)
#define LAUNCH_SWEEP_HALO_REDUCE(c, HALOS, op, eps, kern, ...) do { LAUNCH_SWEEP_HALO(c, HALOS, kern, __VA_ARGS__); TRY(wcsph_finalize_reduce(c, (c)->sweep_parts, op, eps)); } while (0)
