// mgpu.cu -- z-slab domain decomposition over the GPUs of one box.
//
// One process per GPU.  Rank r owns the liquids whose cell layer z lies in [zlo, zhi); solids are
// replicated.  The sort key is z-major, so inside a rank
//     [ghost_lo | owned, cell-sorted | ghost_hi]
// is ONE sorted sequence: the two boundary layers a neighbour needs are a prefix and a suffix of the
// owned range, and the ghosts land right before / right behind it -- the halo exchange is four
// contiguous ncclSend/ncclRecv per field with no packing kernel.  Collectives on the data path:
//   * per neighbour pass: the halo of exactly the field(s) that pass gathers from j (wcsph_halo);
//   * per convergence test: a 1-float all-reduce of the scalar (wcsph_finalize_reduce);
//   * per step: migration of the particles that changed slab (full persistent state) and the
//     all-reduce of the bucket-occupancy table that makes neighborCount exact (HashGrid.py:100 counts
//     candidates of the GLOBAL hash table).
// NCCL is dlopen'ed (the copy torch already loaded), so the single-GPU build has no link dependency.
#include "engine.cuh"
#include <dlfcn.h>
#include <cub/device/device_scan.cuh>

// ---- minimal NCCL ABI (nccl.h 2.x) -------------------------------------------------------------
typedef struct ncclComm* ncclComm_t;
typedef struct { char internal[128]; } ncclUniqueId;
typedef int ncclResult_t;
enum { ncclInt32 = 2, ncclFloat16 = 6, ncclFloat32 = 7 };
enum { ncclSum = 0, ncclMax = 2 };
struct NcclApi {
    void* h;
    ncclResult_t (*GetUniqueId)(ncclUniqueId*);
    ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
    ncclResult_t (*CommDestroy)(ncclComm_t);
    ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*);
    ncclResult_t (*GroupStart)();
    ncclResult_t (*GroupEnd)();
    ncclResult_t (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t);
    ncclResult_t (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t);
    const char* (*GetErrorString)(ncclResult_t);
    ncclResult_t (*CommCount)(const ncclComm_t, int*);
    ncclResult_t (*CommUserRank)(const ncclComm_t, int*);
    ncclResult_t (*GetVersion)(int*);
};
static NcclApi g_nccl = {nullptr};

static int nccl_load(const char* path) {
    if (g_nccl.h) return 0;
    void* h = dlopen(path && path[0] ? path : "libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { wcsph_set_error("dlopen libnccl: %s", dlerror()); return WCSPH_EINVAL; }
#define SYM(field, name) *(void**)(&g_nccl.field) = dlsym(h, name); if (!g_nccl.field) { wcsph_set_error("dlsym %s failed", name); return WCSPH_EINVAL; }
    SYM(GetUniqueId, "ncclGetUniqueId") SYM(CommInitRank, "ncclCommInitRank") SYM(CommDestroy, "ncclCommDestroy") SYM(CommSplit, "ncclCommSplit")
    SYM(GroupStart, "ncclGroupStart") SYM(GroupEnd, "ncclGroupEnd") SYM(Send, "ncclSend") SYM(Recv, "ncclRecv")
    SYM(AllReduce, "ncclAllReduce") SYM(GetErrorString, "ncclGetErrorString")
    SYM(CommCount, "ncclCommCount") SYM(CommUserRank, "ncclCommUserRank") SYM(GetVersion, "ncclGetVersion")
#undef SYM
    g_nccl.h = h;
    return 0;
}
#define NCCL_TRY(x) do { ncclResult_t r_ = (x); if (r_ != 0) { wcsph_set_error("%s:%d %s -> %s", __FILE__, __LINE__, #x, g_nccl.GetErrorString(r_)); return WCSPH_ECUDA; } } while (0)

extern "C" int wcsph_comm_unique_id(void* out, const char* nccl_path) {
    if (!out) return WCSPH_EINVAL;
    TRY(nccl_load(nccl_path));
    ncclUniqueId id;
    NCCL_TRY(g_nccl.GetUniqueId(&id));
    memcpy(out, &id, sizeof(id));
    return 0;
}

extern "C" int wcsph_comm_init(wcsph_ctx* c, const void* id_bytes, const char* nccl_path) {
    if (!c || !id_bytes) return WCSPH_EINVAL;
    if (c->R <= 1) { wcsph_set_error("comm_init on a single-GPU context"); return WCSPH_EINVAL; }
    TRY(nccl_load(nccl_path));
    ncclUniqueId id; memcpy(&id, id_bytes, sizeof(id));
    ncclComm_t comm;
    NCCL_TRY(g_nccl.CommInitRank(&comm, c->R, id, c->rank));
    c->comm = comm;
    ncclComm_t comm2;
    NCCL_TRY(g_nccl.CommSplit(comm, 0, c->rank, &comm2, nullptr));      // same ranks, independent ordering domain
    c->comm2 = comm2;
    int prio_lo = 0, prio_hi = 0;
    CUDA_TRY(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    // highest priority: the few CTAs of a halo kernel must not queue behind the thousands of the interior sweep
    CUDA_TRY(cudaStreamCreateWithPriority(&c->side_stream, cudaStreamNonBlocking, prio_hi));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_main, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_halo, cudaEventDisableTiming));
    CUDA_TRY(cudaEventCreateWithFlags(&c->ev_occ, cudaEventDisableTiming));
    c->use_graph = 0;            // the z-slab step is stream-ordered (host-driven loops + collectives)
    return 0;
}

// what NCCL itself reports for the communicator of this context: {ncclCommCount, ncclCommUserRank, ncclGetVersion, peer mailboxes in use}
extern "C" int wcsph_comm_info(wcsph_ctx* c, int out[4]) {
    if (!c || !out) return WCSPH_EINVAL;
    out[0] = out[1] = out[2] = out[3] = 0;
    if (c->R <= 1 || !c->comm) return 0;
    NCCL_TRY(g_nccl.CommCount((ncclComm_t)c->comm, &out[0]));
    NCCL_TRY(g_nccl.CommUserRank((ncclComm_t)c->comm, &out[1]));
    NCCL_TRY(g_nccl.GetVersion(&out[2]));
    out[3] = c->p2p_scalars;
    return 0;
}

extern "C" int wcsph_migration_counts(wcsph_ctx* c, long long out[5]) {
    if (!c || !out) return WCSPH_EINVAL;
    for (int k = 0; k < 4; k++) out[k] = c->mig_total[k];
    out[4] = c->halo_exchanges;
    return 0;
}

void wcsph_comm_destroy(wcsph_ctx* c) {
    if (c->comm2 && g_nccl.h) { g_nccl.CommDestroy((ncclComm_t)c->comm2); c->comm2 = nullptr; }
    if (c->comm && g_nccl.h) { g_nccl.CommDestroy((ncclComm_t)c->comm); c->comm = nullptr; }
    for (int r = 0; r < WCSPH_MAX_RANKS; r++) if (c->mbox_opened[r]) { cudaIpcCloseMemHandle(c->mbox_opened[r]); c->mbox_opened[r] = nullptr; }
    if (c->mbox_peers) { cudaFree(c->mbox_peers); c->mbox_peers = nullptr; }
    if (c->mbox) { cudaFree(c->mbox); c->mbox = nullptr; }
    c->p2p_scalars = 0;
    if (c->side_stream) {
        cudaStreamDestroy(c->side_stream); cudaEventDestroy(c->ev_main); cudaEventDestroy(c->ev_halo); cudaEventDestroy(c->ev_occ);
        c->side_stream = nullptr;
    }
}

// ---- halo exchange of one field ------------------------------------------------------------------
// sends my two boundary layers (prefix / suffix of the in-box owned range) to the z neighbours and
// receives theirs into the ghost ranges around my owned range
int wcsph_halo_ptr(wcsph_ctx* c, void* base, int stride_floats) {
    if (c->R <= 1) return 0;
    ncclComm_t comm = (ncclComm_t)(c->stream == c->side_stream ? c->comm2 : c->comm);
    float* p = (float*)base;
    const size_t s = (size_t)stride_floats;
    const bool own = c->halo_group_depth == 0;          // inside wcsph_halo_group the enclosing group is timed / counted
    if (own) { prof_begin(c, "nccl_halo"); c->halo_exchanges++; }
    NCCL_TRY(g_nccl.GroupStart());
    if (c->rank > 0) {
        if (c->n_send_lo) NCCL_TRY(g_nccl.Send(p + (size_t)c->i0 * s, (size_t)c->n_send_lo * s, ncclFloat32, c->rank - 1, comm, c->stream));
        if (c->n_glo) NCCL_TRY(g_nccl.Recv(p + (size_t)(c->i0 - c->n_glo) * s, (size_t)c->n_glo * s, ncclFloat32, c->rank - 1, comm, c->stream));
    }
    if (c->rank < c->R - 1) {
        if (c->n_send_hi) NCCL_TRY(g_nccl.Send(p + (size_t)(c->i0 + c->n_inbox - c->n_send_hi) * s, (size_t)c->n_send_hi * s, ncclFloat32, c->rank + 1, comm, c->stream));
        if (c->n_ghi) NCCL_TRY(g_nccl.Recv(p + (size_t)(c->i0 + c->nown) * s, (size_t)c->n_ghi * s, ncclFloat32, c->rank + 1, comm, c->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    if (own) prof_end(c);
    return 0;
}
int wcsph_halo(wcsph_ctx* c, const char* name) {
    if (c->R <= 1) return 0;
    FieldSlot* f = wcsph_find_field(c, name);
    if (!f) { wcsph_set_error("halo: unknown field '%s'", name); return WCSPH_ENAME; }
    return wcsph_halo_ptr(c, f->buf[f->persistent ? c->cur : 0], f->stride);
}

// the halos of ONE sweep (e.g. kappa_v + pos) travel as one NCCL group: nested ncclGroupStart/End fuse into a single launch
int wcsph_halo_group(wcsph_ctx* c, int begin) {
    if (c->R <= 1) return 0;
    if (begin) { prof_begin(c, "nccl_halo"); c->halo_exchanges++; c->halo_group_depth++; NCCL_TRY(g_nccl.GroupStart()); }
    else { NCCL_TRY(g_nccl.GroupEnd()); c->halo_group_depth--; prof_end(c); }
    return 0;
}

// fork / join of the halo onto the side stream (LAUNCH_SWEEP_HALO)
int wcsph_halo_begin(wcsph_ctx* c) {
    CUDA_TRY(cudaEventRecord(c->ev_main, c->stream));               // producers of the halo'd fields are done
    CUDA_TRY(cudaStreamWaitEvent(c->side_stream, c->ev_main, 0));
    c->main_saved = c->stream; c->stream = c->side_stream;
    return 0;
}
int wcsph_halo_end(wcsph_ctx* c) {
    CUDA_TRY(cudaEventRecord(c->ev_halo, c->side_stream));
    c->stream = c->main_saved;
    return 0;
}
int wcsph_halo_wait(wcsph_ctx* c) { CUDA_TRY(cudaStreamWaitEvent(c->stream, c->ev_halo, 0)); return 0; }

// ---- peer mailboxes: latency-bound exchanges without NCCL launches ---------------------------------------------------------------
// the z neighbours' counts: counts[send_lo] -> lower neighbour's cnt[1] (it receives from its UPPER side), counts[send_hi] -> upper's cnt[0]
__global__ void k_p2p_counts(int* counts, Mailbox* mine, Mailbox* const* peers, int R, int rank, unsigned int epoch,
                             int send_lo, int send_hi, int recv_lo, int recv_hi, Scalars* sc) {
    const int t = threadIdx.x;            // 0: lower neighbour, 1: upper neighbour
    const int par = epoch & 1;
    const int nb = t == 0 ? rank - 1 : rank + 1;
    if (t > 1 || nb < 0 || nb >= R) return;
    mb_store(&peers[nb]->cnt[par][1 - t], epoch, (unsigned int)counts[t == 0 ? send_lo : send_hi]);
    counts[t == 0 ? recv_lo : recv_hi] = (int)mb_wait(&mine->cnt[par][t], epoch, sc);
}

extern "C" int wcsph_comm_mailbox_handle(wcsph_ctx* c, void* out64) {
    if (!c || !out64) return WCSPH_EINVAL;
    if (c->R <= 1 || c->R > WCSPH_MAX_RANKS) { wcsph_set_error("mailboxes need 2..%d ranks", WCSPH_MAX_RANKS); return WCSPH_EINVAL; }
    if (!c->mbox) {
        CUDA_TRY(cudaMalloc((void**)&c->mbox, sizeof(Mailbox)));
        CUDA_TRY(cudaMemset(c->mbox, 0, sizeof(Mailbox)));
        CUDA_TRY(cudaDeviceSynchronize());
    }
    cudaIpcMemHandle_t h;
    CUDA_TRY(cudaIpcGetMemHandle(&h, c->mbox));
    static_assert(sizeof(h) == 64, "cudaIpcMemHandle_t is 64 bytes");
    memcpy(out64, &h, 64);
    return 0;
}
extern "C" int wcsph_comm_mailbox_open(wcsph_ctx* c, const void* handles) {
    if (!c || !handles || !c->mbox) { wcsph_set_error("mailbox_open before mailbox_handle"); return WCSPH_EINVAL; }
    Mailbox* host_ptrs[WCSPH_MAX_RANKS];
    for (int r = 0; r < c->R; r++) {
        if (r == c->rank) { host_ptrs[r] = c->mbox; continue; }
        cudaIpcMemHandle_t h; memcpy(&h, (const char*)handles + 64 * (size_t)r, 64);
        void* p = nullptr;
        cudaError_t e = cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess);
        if (e != cudaSuccess) {
            cudaGetLastError();
            for (int q = 0; q < r; q++) if (c->mbox_opened[q]) { cudaIpcCloseMemHandle(c->mbox_opened[q]); c->mbox_opened[q] = nullptr; }
            wcsph_set_error("cudaIpcOpenMemHandle(rank %d): %s -- staying on the NCCL path", r, cudaGetErrorString(e));
            return WCSPH_ECUDA;
        }
        c->mbox_opened[r] = p; host_ptrs[r] = (Mailbox*)p;
    }
    if (!c->mbox_peers) CUDA_TRY(cudaMalloc((void**)&c->mbox_peers, sizeof(Mailbox*) * WCSPH_MAX_RANKS));
    CUDA_TRY(cudaMemcpy(c->mbox_peers, host_ptrs, sizeof(Mailbox*) * c->R, cudaMemcpyHostToDevice));
    c->red_epoch = c->cnt_epoch = 0;
    c->p2p_scalars = 1;
    return 0;
}

// all-reduce of one device float (sum or max) across the ranks, in place
int wcsph_allreduce_scalar(wcsph_ctx* c, float* dev, int is_max) {
    if (c->R <= 1) return 0;
    NCCL_TRY(g_nccl.AllReduce(dev, dev, 1, ncclFloat32, is_max ? ncclMax : ncclSum, (ncclComm_t)c->comm, c->stream));
    return 0;
}

// ---- per-step grid build on a z-slab rank ------------------------------------------------------------
struct MigFields { int n4, n1; float4* f4[8]; float* f1[8]; int* sid; };     // persistent fields, pointers at the owned start
static MigFields mig_fields(wcsph_ctx* c) {
    MigFields F; memset(&F, 0, sizeof(F));
    for (int f = 0; f < c->nfields; f++) {
        FieldSlot& S = c->fields[f];
        if (!S.persistent) continue;
        if (S.stride == 4) F.f4[F.n4++] = (float4*)S.buf[c->cur] + c->i0;
        else F.f1[F.n1++] = (float*)S.buf[c->cur] + c->i0;
    }
    F.sid = c->sorted_id[c->cur] + c->i0;
    return F;
}
// The radix sort of a slab rank runs on keys RELATIVE to the first cell the rank can see: a rank of an 8-way split of 20 M cells sorts
// 22-bit keys (3 digit passes) instead of 25-bit ones (4); the special keys (left the box, dead slot) follow the slab's span.
// k_permute turns keys_sorted back into global cell ids, so every consumer keeps reading those.
__device__ __forceinline__ int slab_sort_key(int key, int ncells, int kbase, int kspan) { return key >= ncells ? kspan + (key - ncells) : key - kbase; }
// keys of the owned particles + migration in one pass: cell id if the particle stays (in box and in slab),
// ncells if it left the box (stays with its owner, HashGrid.py:81), ncells+3 (dead slot, sorts last) if it
// moved to a neighbour slab -- its full persistent state is packed as one record into the send staging of
// that neighbour.  Every in-box particle counts once into the bucket occupancy; stayers into the cell histogram.
__global__ void k_keys_migrate_pack(MigFields F, int n, GridDims g, int zlo, int zhi, int has_lo, int has_hi,
                                    int* __restrict__ keys, int* __restrict__ occ, int* __restrict__ cell_count, int* __restrict__ counts,
                                    float* __restrict__ send_lo, float* __restrict__ send_hi, int cap, int rec, int kbase, int kspan) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = F.f4[0][i];                    // field 0 is pos
    int cx, cy, cz; cell_coords(g, p.x, p.y, p.z, cx, cy, cz);
    int key = g.ncells;
    if (in_box(g, cx, cy, cz)) {
        atomicAdd(&occ[cell_hash(cx, cy, cz, g.n_hash)], 1);
        const int dir = (cz < zlo && has_lo) ? 0 : ((cz >= zhi && has_hi) ? 1 : -1);
        if (dir >= 0) {
            key = g.ncells + 3;
            const int slot = atomicAdd(&counts[dir], 1);
            if (slot < cap) {
                float* r = (dir ? send_hi : send_lo) + (size_t)slot * rec;
                for (int f = 0; f < F.n4; f++) { float4 v = F.f4[f][i]; r[0] = v.x; r[1] = v.y; r[2] = v.z; r[3] = v.w; r += 4; }
                for (int f = 0; f < F.n1; f++) *r++ = F.f1[f][i];
                *r = __int_as_float(F.sid[i]);
            }
        } else {
            key = (cz * g.by + cy) * g.bx + cx;
            atomicAdd(&cell_count[key], 1);
        }
    }
    keys[i] = slab_sort_key(key, g.ncells, kbase, kspan);
}
// arrivals: records -> field slots behind the owned range, + their keys and cell histogram
// A particle whose cell layer lies outside this rank's slab (it crossed more than one slab in a step, or state was
// restored without re-partitioning) cannot be filed: the cell histogram only spans the slab's layers.  It is parked
// like an out-of-box particle (no neighbours) and WCSPH_FLAG_MIGRATE_FAR is raised -- a hard error at the next check.
__global__ void k_unpack_arrivals(MigFields F, const float* __restrict__ recv, int n, int dst0, int rec, GridDims g,
                                  int* __restrict__ keys, int* __restrict__ cell_count, int zlo, int zhi, int has_lo, int has_hi, Scalars* sc,
                                  int kbase, int kspan) {
    int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= n) return;
    const float* r = recv + (size_t)k * rec;
    const int i = dst0 + k;
    float4 p = make_float4(r[0], r[1], r[2], r[3]);
    for (int f = 0; f < F.n4; f++) { F.f4[f][i] = make_float4(r[0], r[1], r[2], r[3]); r += 4; }
    for (int f = 0; f < F.n1; f++) F.f1[f][i] = *r++;
    F.sid[i] = __float_as_int(*r);
    int cx, cy, cz; cell_coords(g, p.x, p.y, p.z, cx, cy, cz);
    int key = g.ncells;
    if (in_box(g, cx, cy, cz)) {
        if ((cz < zlo && has_lo) || (cz >= zhi && has_hi)) atomicOr(&sc->flags, WCSPH_FLAG_MIGRATE_FAR);
        else { key = (cz * g.by + cy) * g.bx + cx; atomicAdd(&cell_count[key], 1); }
    }
    keys[i] = slab_sort_key(key, g.ncells, kbase, kspan);
}
__global__ void k_keys_ghost(const float4* __restrict__ pos, int n, GridDims g, int* __restrict__ cell_count) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    float4 p = pos[i];
    int cx, cy, cz; cell_coords(g, p.x, p.y, p.z, cx, cy, cz);
    if (in_box(g, cx, cy, cz)) atomicAdd(&cell_count[(cz * g.by + cy) * g.bx + cx], 1);
}
__device__ int lower_bound_dev(const int* a, int n, int v) {
    int lo = 0, hi = n;
    while (lo < hi) { int m = (lo + hi) >> 1; if (a[m] < v) lo = m + 1; else hi = m; }
    return lo;
}
// counts[4] = n_inbox, counts[5] = n_send_lo (my lowest two layers), counts[6] = n_send_hi (my highest two)
__global__ void k_halo_counts(const int* __restrict__ keys_sorted, int n, GridDims g, int zlo, int zhi, int* __restrict__ counts) {
    if (threadIdx.x || blockIdx.x) return;
    const int plane = g.bx * g.by;
    const int n_in = lower_bound_dev(keys_sorted, n, g.ncells);
    counts[4] = n_in;
    counts[5] = lower_bound_dev(keys_sorted, n_in, min(zlo + 2, g.bz) * plane);
    counts[6] = n_in - lower_bound_dev(keys_sorted, n_in, max(min(zhi, g.bz) - 2, 0) * plane);
}
// The bucket-occupancy table travels as fp16: a rank's share is saturated at 255 per bucket (maxInGrid is 64: anything above is
// the overflow case and still reads as > 64 after the sum), so every partial sum of <= 8 ranks is an integer <= 2040 < 2048 and
// therefore exact in fp16 whatever order NCCL adds in.  Half the bytes of the int32 all-reduce of round 1.
#include <cuda_fp16.h>
__global__ void k_occ_pack(const int* __restrict__ occ, __half* __restrict__ h, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) h[i] = __int2half_rn(min(occ[i], 255));
}
__global__ void k_occ_unpack_add(const __half* __restrict__ h, const int* __restrict__ solid, int* __restrict__ occ, int n) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) occ[i] = __half2int_rn(h[i]) + solid[i];
}

static int exchange_counts(wcsph_ctx* c, int send_lo_idx, int send_hi_idx, int recv_lo_idx, int recv_hi_idx) {
    ncclComm_t comm = (ncclComm_t)c->comm;
    prof_begin(c, "nccl_counts(+host sync)");
    if (c->p2p_scalars) {
        c->cnt_epoch++;
        k_p2p_counts<<<1, 32, 0, c->stream>>>(c->mg_counts, c->mbox, c->mbox_peers, c->R, c->rank, c->cnt_epoch, send_lo_idx, send_hi_idx, recv_lo_idx, recv_hi_idx, c->sc); LAUNCH_CHECK(c);
        CUDA_TRY(cudaMemcpyAsync(c->mg_counts_host, c->mg_counts, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
        prof_end(c);
        CUDA_TRY(cudaStreamSynchronize(c->stream));
        return 0;
    }
    NCCL_TRY(g_nccl.GroupStart());
    if (c->rank > 0) {
        NCCL_TRY(g_nccl.Send(c->mg_counts + send_lo_idx, 1, ncclInt32, c->rank - 1, comm, c->stream));
        NCCL_TRY(g_nccl.Recv(c->mg_counts + recv_lo_idx, 1, ncclInt32, c->rank - 1, comm, c->stream));
    }
    if (c->rank < c->R - 1) {
        NCCL_TRY(g_nccl.Send(c->mg_counts + send_hi_idx, 1, ncclInt32, c->rank + 1, comm, c->stream));
        NCCL_TRY(g_nccl.Recv(c->mg_counts + recv_hi_idx, 1, ncclInt32, c->rank + 1, comm, c->stream));
    }
    NCCL_TRY(g_nccl.GroupEnd());
    CUDA_TRY(cudaMemcpyAsync(c->mg_counts_host, c->mg_counts, 16 * sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    prof_end(c);
    CUDA_TRY(cudaStreamSynchronize(c->stream));
    return 0;
}

int wcsph_mgpu_update_grid(wcsph_ctx* c) {
    if (!c->comm) { wcsph_set_error("z-slab context without a communicator (call wcsph_comm_init)"); return WCSPH_EINVAL; }
    const GridDims g = c->g;
    cudaStream_t st = c->stream;
    ncclComm_t comm = (ncclComm_t)c->comm;
    const int i0 = c->i0;
    const int has_lo = c->rank > 0, has_hi = c->rank < c->R - 1;
    FieldSlot* fp = wcsph_find_field(c, "pos");
    // A. keys + bucket occupancy of the owned liquids; leavers are packed for their new owner in the same pass
    // (the cell histogram / scan only spans the layers this rank can see: slab + 2 ghost layers per side)
    const int plane = g.bx * g.by;
    const int cz0 = max(c->zlo - 2, 0) * plane, cz1 = min(min(c->zhi, g.bz) + 2, g.bz) * plane;
    const int rec = c->mig_rec;
    MigFields F = mig_fields(c);
    CUDA_TRY(cudaMemsetAsync(c->mg_counts, 0, 16 * sizeof(int), st));
    CUDA_TRY(cudaMemsetAsync(c->occ, 0, (size_t)c->N * 4, st));
    CUDA_TRY(cudaMemsetAsync(c->cell_start_l + cz0, 0, ((size_t)(cz1 - cz0) + 2) * 4, st));
    if (c->nown > 0) {
        prof_begin(c, "k_keys_migrate_pack");
        k_keys_migrate_pack<<<nblocks(c->nown), WCSPH_BLOCK, 0, st>>>(F, c->nown, g, c->zlo, c->zhi, has_lo, has_hi, c->keys, c->occ, c->cell_start_l,
                                                                  c->mg_counts, c->mig_send[0], c->mig_send[1], c->G, rec, cz0, cz1 - cz0);
        prof_end(c); LAUNCH_CHECK(c);
    }
    // I (early). global bucket occupancy = sum of the ranks' liquid shares (+ the replicated solid share, added
    // below): all-reduced on the side stream / second communicator, hidden behind the migration and the sort
    CUDA_TRY(cudaEventRecord(c->ev_main, st));
    CUDA_TRY(cudaStreamWaitEvent(c->side_stream, c->ev_main, 0));
    {   // side stream: pack -> all-reduce (fp16) -> unpack + solids -> the 5x5x5 box filter, all hidden behind migration and sort
        cudaStream_t main_st = c->stream;
        c->stream = c->side_stream;
        k_occ_pack<<<nblocks(c->N), WCSPH_BLOCK, 0, c->side_stream>>>(c->occ, (__half*)c->occ_h, c->N); LAUNCH_CHECK(c);
        prof_begin(c, "nccl_allreduce_occ(fp16)");
        ncclResult_t r_ = g_nccl.AllReduce(c->occ_h, c->occ_h, (size_t)c->N, ncclFloat16, ncclSum, (ncclComm_t)c->comm2, c->side_stream);
        prof_end(c);
        if (r_ != 0) { c->stream = main_st; wcsph_set_error("ncclAllReduce(occ) -> %s", g_nccl.GetErrorString(r_)); return WCSPH_ECUDA; }
        k_occ_unpack_add<<<nblocks(c->N), WCSPH_BLOCK, 0, c->side_stream>>>((const __half*)c->occ_h, c->occ_solid, c->occ, c->N); LAUNCH_CHECK(c);
        int rb = wcsph_box_filter(c);
        c->stream = main_st;
        if (rb) return rb;
    }
    CUDA_TRY(cudaEventRecord(c->ev_occ, c->side_stream));
    // C. how many cross each face
    TRY(exchange_counts(c, 0, 1, 2, 3));
    const int n_lo = c->mg_counts_host[0], n_up = c->mg_counts_host[1];
    const int n_from_lo = has_lo ? c->mg_counts_host[2] : 0, n_from_up = has_hi ? c->mg_counts_host[3] : 0;
    const int n_all = c->nown + n_from_lo + n_from_up;
    c->mig_total[0] += has_lo ? n_lo : 0; c->mig_total[1] += has_hi ? n_up : 0; c->mig_total[2] += n_from_lo; c->mig_total[3] += n_from_up;
    if (n_lo > c->G || n_up > c->G || n_from_lo > c->G || n_from_up > c->G) { wcsph_set_error("rank %d: %d/%d migrants exceed the staging capacity %d", c->rank, n_lo, n_up, c->G); return WCSPH_ENOMEM; }
    if (n_all > c->capOwn) { wcsph_set_error("rank %d: %d owned particles exceed cap_own %d", c->rank, n_all, c->capOwn); return WCSPH_ENOMEM; }
    // D. one packed message per face; arrivals are unpacked behind the owned range
    if (n_lo + n_up + n_from_lo + n_from_up > 0) {
        prof_begin(c, "nccl_migrate");
        NCCL_TRY(g_nccl.GroupStart());
        if (has_lo) {
            if (n_lo) NCCL_TRY(g_nccl.Send(c->mig_send[0], (size_t)n_lo * rec, ncclFloat32, c->rank - 1, comm, st));
            if (n_from_lo) NCCL_TRY(g_nccl.Recv(c->mig_recv[0], (size_t)n_from_lo * rec, ncclFloat32, c->rank - 1, comm, st));
        }
        if (has_hi) {
            if (n_up) NCCL_TRY(g_nccl.Send(c->mig_send[1], (size_t)n_up * rec, ncclFloat32, c->rank + 1, comm, st));
            if (n_from_up) NCCL_TRY(g_nccl.Recv(c->mig_recv[1], (size_t)n_from_up * rec, ncclFloat32, c->rank + 1, comm, st));
        }
        NCCL_TRY(g_nccl.GroupEnd());
        prof_end(c);
        if (n_from_lo) { k_unpack_arrivals<<<nblocks(n_from_lo), WCSPH_BLOCK, 0, st>>>(F, c->mig_recv[0], n_from_lo, c->nown, rec, g, c->keys, c->cell_start_l, c->zlo, c->zhi, has_lo, has_hi, c->sc, cz0, cz1 - cz0); LAUNCH_CHECK(c); }
        if (n_from_up) { k_unpack_arrivals<<<nblocks(n_from_up), WCSPH_BLOCK, 0, st>>>(F, c->mig_recv[1], n_from_up, c->nown + n_from_lo, rec, g, c->keys, c->cell_start_l, c->zlo, c->zhi, has_lo, has_hi, c->sc, cz0, cz1 - cz0); LAUNCH_CHECK(c); }
    }
    // E. ONE sort: [in box, cell-sorted | left the box | dead slots of the leavers]
    if (n_all > 0) TRY(wcsph_sort_permute(c, n_all, cz0, cz1 - cz0));
    c->nown = n_all - n_lo - n_up;
    // F. sizes of the boundary layers, mine and the neighbours'
    k_halo_counts<<<1, 1, 0, st>>>(c->keys_sorted, c->nown, g, c->zlo, c->zhi, c->mg_counts); LAUNCH_CHECK(c);
    TRY(exchange_counts(c, 5, 6, 7, 8));
    c->n_inbox = c->mg_counts_host[4];
    c->n_send_lo = has_lo ? c->mg_counts_host[5] : 0; c->n_send_hi = has_hi ? c->mg_counts_host[6] : 0;
    c->n_glo = has_lo ? c->mg_counts_host[7] : 0; c->n_ghi = has_hi ? c->mg_counts_host[8] : 0;
    if (c->n_glo > c->G || c->n_ghi > c->G) { wcsph_set_error("rank %d: ghost layer (%d / %d) exceeds cap_ghost %d", c->rank, c->n_glo, c->n_ghi, c->G); return WCSPH_ENOMEM; }
    // G. ghost positions; H. one cell_start table over [ghost_lo | owned in box | ghost_hi]
    TRY(wcsph_halo(c, "pos"));
    const float4* pos = (const float4*)fp->buf[c->cur];
    if (c->n_glo) { k_keys_ghost<<<nblocks(c->n_glo), WCSPH_BLOCK, 0, st>>>(pos + i0 - c->n_glo, c->n_glo, g, c->cell_start_l); LAUNCH_CHECK(c); }
    if (c->n_ghi) { k_keys_ghost<<<nblocks(c->n_ghi), WCSPH_BLOCK, 0, st>>>(pos + i0 + c->nown, c->n_ghi, g, c->cell_start_l); LAUNCH_CHECK(c); }
    size_t tb = c->cub_temp_bytes;
    CUDA_TRY(cub::DeviceScan::ExclusiveSum(c->cub_temp, tb, c->cell_start_l + cz0, c->cell_start_l + cz0, cz1 - cz0 + 1, st));
    c->launches += 2;
    // I. join the side stream: global occupancy (+ solids) and its box sums are ready
    CUDA_TRY(cudaStreamWaitEvent(st, c->ev_occ, 0));
    // J. neighborCount + lists for the owned particles
    CellStartArgs csa;
    csa.base = i0 - c->n_glo;
    csa.hi_cell0 = has_hi ? min(c->zhi, g.bz) * g.bx * g.by : 0x7fffffff;
    csa.n_oob = c->nown - c->n_inbox;
    csa.c_lo = cz0; csa.c_hi = cz1; csa.box_done = 1;
    return wcsph_grid_finish(c, csa);
}
