// tile_proto.cu -- A/B of the north star's "shared-memory staging of the neighbour-cell tiles" against the engine's compact-list
// gather sweep, on the same pairs.  Stand-alone (nvcc -arch=sm_100a tile_proto.cu -o tile_proto; ./tile_proto [n] [reps]).
//
// Scene: n^3 jittered lattice, spacing d = 0.05 = hash cell, support h = 0.1 (the DFSPH configuration: ~1 particle per cell, ~30
// in-range neighbours).  Kernel body = k_dfsph_drho's: s_i = sum_j gradW_u(|r|^2) * (v_i - v_j) . r   (two float4 gathers per pair).
//   A  row-major cell sort (x fastest), uint32 absolute neighbour indices, warp-interleaved uint4 groups, LDG.128 gathers  (the engine)
//   B  block-tiled sort (8 x 8 x 4 cells per CTA), the halo'd tile (12 x 12 x 8 cells) staged in shared memory, uint16 tile-local
//      neighbour indices (8 per uint4), LDS.128 gathers
// Both variants sum the SAME neighbour sets (checked); B's staging loads are part of its time.
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <math.h>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s:%d %s\n", __FILE__, __LINE__, cudaGetErrorString(e)); exit(1); } } while (0)
#define CAP 64
__device__ __forceinline__ float sat_1mq(float q) { float t; asm("sub.sat.ftz.f32 %0, %1, %2;" : "=f"(t) : "f"(1.0f), "f"(q)); return t; }
__device__ __forceinline__ float sat_1m2q(float q) { float u; asm("fma.rn.sat.ftz.f32 %0, %1, %2, %3;" : "=f"(u) : "f"(-2.0f), "f"(q), "f"(1.0f)); return u; }
__device__ __forceinline__ float gradW_u(float r2, float inv_h) {
    const float inv_rl = r2 > 1.0e-10f ? rsqrtf(r2) : 0.0f;
    const float q = r2 * inv_rl * inv_h;
    const float t = sat_1mq(q), u = sat_1m2q(q);
    return fmaf(u, u, -(t * t)) * inv_rl;
}
#define PAIR(PJ, VJ) { const float4 pj = (PJ); const float4 vj = (VJ); const float rx = pi.x - pj.x, ry = pi.y - pj.y, rz = pi.z - pj.z; \
    const float r2 = rx * rx + ry * ry + rz * rz; s += gradW_u(r2, inv_h) * ((vi.x - vj.x) * rx + (vi.y - vj.y) * ry + (vi.z - vj.z) * rz); }

// ---- A: the engine's sweep ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) k_sweep_A(const float4* __restrict__ pos, const float4* __restrict__ vel, const uint4* __restrict__ nbr,
                                                 const int* __restrict__ cnt, int n, float inv_h, float* __restrict__ out) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n) return;
    const float4 pi = pos[i], vi = vel[i];
    const uint4* row = nbr + (size_t)(i >> 5) * (CAP / 4) * 32 + (i & 31);
    const int n4 = (cnt[i] + 3) >> 2;
    float s = 0.f;
    uint4 Jn = make_uint4(0, 0, 0, 0);
    if (n4 > 0) Jn = __ldcs(row);
    for (int k = 0; k < n4; k++) {
        const uint4 J = Jn;
        if (k + 1 < n4) Jn = __ldcs(row + (size_t)(k + 1) * 32);
        PAIR(pos[J.x], vel[J.x]) PAIR(pos[J.y], vel[J.y]) PAIR(pos[J.z], vel[J.z]) PAIR(pos[J.w], vel[J.w])
    }
    out[i] = s;
}

// ---- B: staged tiles -----------------------------------------------------------------------------------------------------------------
#define TX 8
#define TY 8
#define TZ 4
#define HX (TX + 4)
#define HY (TY + 4)
#define HZ (TZ + 4)
#define HALO_CELLS (HX * HY * HZ)           // 1152
// lattice with exactly one particle per cell: slot of cell (x, y, z) in the block-tiled order
__host__ __device__ __forceinline__ int tiled_slot(int x, int y, int z, int ntx, int nty) {
    const int tx = x / TX, ty = y / TY, tz = z / TZ;
    return (((tz * nty + ty) * ntx + tx) * (TX * TY * TZ)) + ((z % TZ) * TY + (y % TY)) * TX + (x % TX);
}
__global__ void __launch_bounds__(256) k_sweep_B(const float4* __restrict__ pos, const float4* __restrict__ vel, const uint4* __restrict__ nbr16,
                                                 const int* __restrict__ cnt, int ncell, int ntx, int nty, float inv_h, float* __restrict__ out) {
    __shared__ float4 sp[HALO_CELLS], sv[HALO_CELLS];
    const int tile = blockIdx.x;
    const int tx = tile % ntx, ty = (tile / ntx) % nty, tz = tile / (ntx * nty);
    // stage the halo'd tile: local index l = ((lz * HY) + ly) * HX + lx  <->  cell (tx*TX - 2 + lx, ...); cells outside the domain stay far away
    for (int l = threadIdx.x; l < HALO_CELLS; l += 256) {
        const int lx = l % HX, ly = (l / HX) % HY, lz = l / (HX * HY);
        const int x = tx * TX - 2 + lx, y = ty * TY - 2 + ly, z = tz * TZ - 2 + lz;
        float4 p = make_float4(1e9f, 1e9f, 1e9f, 0.f), v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (x >= 0 && y >= 0 && z >= 0 && x < ncell && y < ncell && z < ncell) { const int s = tiled_slot(x, y, z, ntx, nty); p = pos[s]; v = vel[s]; }
        sp[l] = p; sv[l] = v;
    }
    __syncthreads();
    const int t = threadIdx.x;
    const int cx = t % TX, cy = (t / TX) % TY, cz = t / (TX * TY);
    if (tx * TX + cx >= ncell || ty * TY + cy >= ncell || tz * TZ + cz >= ncell) return;
    const int i = tile * 256 + t;
    const int li = ((cz + 2) * HY + (cy + 2)) * HX + (cx + 2);
    const float4 pi = sp[li], vi = sv[li];
    const uint4* row = nbr16 + (size_t)(i >> 5) * (CAP / 8) * 32 + (i & 31);
    const int n8 = (cnt[i] + 7) >> 3;
    float s = 0.f;
    uint4 Jn = make_uint4(0, 0, 0, 0);
    if (n8 > 0) Jn = __ldcs(row);
    for (int k = 0; k < n8; k++) {
        const uint4 J = Jn;
        if (k + 1 < n8) Jn = __ldcs(row + (size_t)(k + 1) * 32);
        PAIR(sp[J.x & 0xffff], sv[J.x & 0xffff]) PAIR(sp[J.x >> 16], sv[J.x >> 16]) PAIR(sp[J.y & 0xffff], sv[J.y & 0xffff]) PAIR(sp[J.y >> 16], sv[J.y >> 16])
        PAIR(sp[J.z & 0xffff], sv[J.z & 0xffff]) PAIR(sp[J.z >> 16], sv[J.z >> 16]) PAIR(sp[J.w & 0xffff], sv[J.w & 0xffff]) PAIR(sp[J.w >> 16], sv[J.w >> 16])
    }
    out[i] = s;
}

int main(int argc, char** argv) {
    const int nc = argc > 1 ? atoi(argv[1]) : 100;
    const int reps = argc > 2 ? atoi(argv[2]) : 20;
    const float d = 0.05f, h = 0.1f;
    const int n = nc * nc * nc;
    const int ntx = (nc + TX - 1) / TX, nty = (nc + TY - 1) / TY, ntz = (nc + TZ - 1) / TZ;
    const int ntile = ntx * nty * ntz, nB = ntile * 256;
    printf("lattice %d^3 = %d particles, tiles %d x %d x %d (%d slots for B)\n", nc, n, ntx, nty, ntz, nB);
    std::vector<float4> P(n), V(n);
    srand(7);
    auto rnd = []() { return (float)rand() / RAND_MAX - 0.5f; };
    for (int z = 0; z < nc; z++) for (int y = 0; y < nc; y++) for (int x = 0; x < nc; x++) {
        const int c = (z * nc + y) * nc + x;
        P[c] = make_float4((x + 0.5f + 0.4f * rnd()) * d, (y + 0.5f + 0.4f * rnd()) * d, (z + 0.5f + 0.4f * rnd()) * d, 0.f);   // stays inside its cell
        V[c] = make_float4(rnd(), rnd(), rnd(), 0.f);
    }
    // neighbour sets (cell coordinates), shared by both layouts
    std::vector<int> cnt(n);
    std::vector<std::vector<int>> nb(n);
    long long pairs = 0;
    for (int z = 0; z < nc; z++) for (int y = 0; y < nc; y++) for (int x = 0; x < nc; x++) {
        const int c = (z * nc + y) * nc + x;
        for (int dz = -2; dz <= 2; dz++) for (int dy = -2; dy <= 2; dy++) for (int dx = -2; dx <= 2; dx++) {
            const int X = x + dx, Y = y + dy, Z = z + dz;
            if (X < 0 || Y < 0 || Z < 0 || X >= nc || Y >= nc || Z >= nc || (!dx && !dy && !dz)) continue;
            const int c2 = (Z * nc + Y) * nc + X;
            const float rx = P[c].x - P[c2].x, ry = P[c].y - P[c2].y, rz = P[c].z - P[c2].z;
            if (rx * rx + ry * ry + rz * rz <= h * h) nb[c].push_back(((dz + 2) * 5 + (dy + 2)) * 5 + (dx + 2));
        }
        cnt[c] = (int)nb[c].size(); pairs += cnt[c];
        if (cnt[c] > CAP) { printf("cap\n"); return 1; }
    }
    printf("pairs per particle %.2f\n", (double)pairs / n);
    // A: row-major slots = c; lists of absolute indices, padded with self to a multiple of 4
    std::vector<uint32_t> LA((size_t)((n + 31) / 32) * CAP * 32, 0);
    for (int c = 0; c < n; c++) {
        const int x = c % nc, y = (c / nc) % nc, z = c / (nc * nc);
        for (int k = 0; k < ((cnt[c] + 3) & ~3); k++) {
            int j = c;
            if (k < cnt[c]) { const int o = nb[c][k]; j = ((z + o / 25 - 2) * nc + (y + (o / 5) % 5 - 2)) * nc + (x + o % 5 - 2); }
            LA[(((size_t)(c >> 5) * (CAP / 4) + (k >> 2)) * 32 + (c & 31)) * 4 + (k & 3)] = (uint32_t)j;
        }
    }
    // B: tiled slots; lists of tile-local halo indices (uint16), padded with self to a multiple of 8
    std::vector<float4> PB(nB, make_float4(1e9f, 1e9f, 1e9f, 0.f)), VB(nB, make_float4(0, 0, 0, 0));
    std::vector<int> cntB(nB, 0);
    std::vector<uint16_t> LB((size_t)((nB + 31) / 32) * CAP * 32, 0);
    std::vector<int> slotOf(n);
    for (int c = 0; c < n; c++) {
        const int x = c % nc, y = (c / nc) % nc, z = c / (nc * nc);
        const int s = tiled_slot(x, y, z, ntx, nty);
        slotOf[c] = s; PB[s] = P[c]; VB[s] = V[c]; cntB[s] = cnt[c];
        const int lx = x % TX + 2, ly = y % TY + 2, lz = z % TZ + 2;
        for (int k = 0; k < ((cnt[c] + 7) & ~7); k++) {
            int l = (lz * HY + ly) * HX + lx;
            if (k < cnt[c]) { const int o = nb[c][k]; l = ((lz + o / 25 - 2) * HY + (ly + (o / 5) % 5 - 2)) * HX + (lx + o % 5 - 2); }
            LB[(((size_t)(s >> 5) * (CAP / 8) + (k >> 3)) * 32 + (s & 31)) * 8 + (k & 7)] = (uint16_t)l;
        }
    }
    float4 *dP, *dV, *dPB, *dVB; uint4 *dLA, *dLB; int *dC, *dCB; float *dO, *dOB;
    CK(cudaMalloc(&dP, n * 16)); CK(cudaMalloc(&dV, n * 16)); CK(cudaMalloc(&dPB, (size_t)nB * 16)); CK(cudaMalloc(&dVB, (size_t)nB * 16));
    CK(cudaMalloc(&dLA, LA.size() * 4)); CK(cudaMalloc(&dLB, LB.size() * 2)); CK(cudaMalloc(&dC, n * 4)); CK(cudaMalloc(&dCB, (size_t)nB * 4));
    CK(cudaMalloc(&dO, n * 4)); CK(cudaMalloc(&dOB, (size_t)nB * 4));
    CK(cudaMemcpy(dP, P.data(), n * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dV, V.data(), n * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dPB, PB.data(), (size_t)nB * 16, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dVB, VB.data(), (size_t)nB * 16, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dLA, LA.data(), LA.size() * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dLB, LB.data(), LB.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dC, cnt.data(), n * 4, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dCB, cntB.data(), (size_t)nB * 4, cudaMemcpyHostToDevice));
    CK(cudaMemset(dOB, 0, (size_t)nB * 4));
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    float msA = 0, msB = 0;
    for (int w = 0; w < 2; w++) {
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) k_sweep_A<<<(n + 255) / 256, 256>>>(dP, dV, dLA, dC, n, 1.0f / h, dO);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&msA, e0, e1);
        cudaEventRecord(e0);
        for (int r = 0; r < reps; r++) k_sweep_B<<<ntile, 256>>>(dPB, dVB, dLB, dCB, nc, ntx, nty, 1.0f / h, dOB);
        cudaEventRecord(e1); CK(cudaEventSynchronize(e1)); cudaEventElapsedTime(&msB, e0, e1);
    }
    CK(cudaGetLastError());
    std::vector<float> OA(n), OB(nB);
    CK(cudaMemcpy(OA.data(), dO, n * 4, cudaMemcpyDeviceToHost)); CK(cudaMemcpy(OB.data(), dOB, (size_t)nB * 4, cudaMemcpyDeviceToHost));
    double maxd = 0, maxv = 0;
    for (int c = 0; c < n; c++) { maxd = fmax(maxd, fabs(OA[c] - OB[slotOf[c]])); maxv = fmax(maxv, fabs(OA[c])); }
    printf("A (compact lists, LDG gathers)      : %.4f ms per sweep, list bytes %.1f MB\n", msA / reps, (double)pairs / n * 4 * n / 1e6);
    printf("B (staged tiles, LDS gathers, u16)  : %.4f ms per sweep, list bytes %.1f MB, smem %d B per CTA\n", msB / reps, (double)pairs / n * 2 * n / 1e6, (int)(2 * HALO_CELLS * 16));
    printf("B / A = %.3f   max |A - B| = %.3g (max |A| %.3g)\n", msB / msA, maxd, maxv);
    return 0;
}
