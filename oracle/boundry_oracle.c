/*
 * boundry_oracle.c -- CPU restatement of /root/reference/boundry.py: parallel Poisson-disk sampling of a
 * triangle mesh after Bowers et al. 2010, as the reference wrote it.  TEST INFRASTRUCTURE ONLY (see wcsph_oracle.h).
 *
 * Pinned on tests/golden/ref_exec_boundry.npz, which the UNMODIFIED boundry.py produced under oracle/tishim.  Everything after
 * the random initial point set (init_point_set, boundry.py:223-247, ti.random) is deterministic once the two racing spots of
 * the reference's parallel loops are given their serial meaning (ascending index):
 *   build_hmap :250-271   two different cells with the same hash slot -> the later one keeps it; phase_group append order
 *   possion_disk_sample :390-407   append order of possion_sample
 * Faithfully kept: tri_normal is indexed with the FACE id although it holds one normal per VERTEX (:361-362 -> normal of face
 * id/3); the hash map is read without checking that the slot belongs to the queried cell (:345-349); a sixth sample of a cell
 * resets its count to 4 (:400-402); trial 0 never visits phase group 0 (:421-457).
 */
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct {
    int n, padding, hash_size, phase_vec_max;
    float radius, gridR;
    float minp[3];
} BdParams;

/* :243-246 init_cell = cast((pos - min_point) / gridR, i32) + 1; padding entries (i >= n) sort last */
void oracle_bd_cells(const BdParams* p, const float* pos, int* cell) {
    for (int i = 0; i < p->padding; i++)
        for (int d = 0; d < 3; d++)
            cell[3 * i + d] = i < p->n ? (int)((pos[3 * i + d] - p->minp[d]) / p->gridR) + 1 : 1000000;
}

static int compare_cell(const int* cell, int i, int j) {                       /* :293-305 */
    const int* a = cell + 3 * i; const int* b = cell + 3 * j;
    if (a[0] > b[0]) return 1;
    if (a[0] == b[0] && a[1] > b[1]) return 1;
    if (a[0] == b[0] && a[1] == b[1] && a[2] > b[2]) return 1;
    if (a[0] == b[0] && a[1] == b[1] && a[2] == b[2]) return 0;
    return -1;
}
static void swap_cell(int* cell, float* pos, int* id, int i, int j) {           /* :307-319 */
    for (int d = 0; d < 3; d++) { int t = cell[3 * i + d]; cell[3 * i + d] = cell[3 * j + d]; cell[3 * j + d] = t; }
    for (int d = 0; d < 3; d++) { float t = pos[3 * i + d]; pos[3 * i + d] = pos[3 * j + d]; pos[3 * j + d] = t; }
    int t = id[i]; id[i] = id[j]; id[j] = t;
}
/* gpu_bitonic_sort :208-219 + gpu_merge :322-336 */
void oracle_bd_bitonic_sort(const BdParams* p, int* cell, float* pos, int* id) {
    for (int k = 2; k <= p->padding; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1)
            for (int i = 0; i < p->padding; i++) {
                const int ixj = i ^ j;
                if (ixj > i) {
                    if ((i & k) == 0) { if (compare_cell(cell, i, ixj) == 1) swap_cell(cell, pos, id, i, ixj); }
                    else { if (compare_cell(cell, i, ixj) == -1) swap_cell(cell, pos, id, i, ixj); }
                }
            }
}

static int cell_hash(const int* a, int size) {                                  /* :284-290, i32 wrap-around */
    const int p1 = (int)(73856093u * (unsigned)a[0]), p2 = (int)(19349663u * (unsigned)a[1]), p3 = (int)(83492791u * (unsigned)a[2]);
    int m = (p1 ^ p2 ^ p3) % size;
    if (m < 0) m += size;
    return m;
}

/* build_hmap :250-271.  start_index / hcell: [hash_size]; phase_group: [27][phase_vec_max][3]; returns the occupied-slot count */
int oracle_bd_build_hmap(const BdParams* p, const int* cell, int* start_index, int* hcell, int* hash_trace,
                         int* phase_group_count, int* phase_group) {
    int hash_count = 0;
    for (int i = 0; i < p->n; i++) {
        hash_trace[i] = 0;
        if (i == 0 || compare_cell(cell, i, i - 1) != 0) {
            const int h = cell_hash(cell + 3 * i, p->hash_size);
            start_index[h] = i;
            for (int d = 0; d < 3; d++) hcell[3 * h + d] = cell[3 * i + d];
            hash_trace[i] = h;
            hash_count++;
            const int ph = cell[3 * i] % 3 + 3 * (cell[3 * i + 1] % 3) + 9 * (cell[3 * i + 2] % 3);
            const int old = phase_group_count[ph]++;
            if (old < p->phase_vec_max) for (int d = 0; d < 3; d++) phase_group[((size_t)ph * p->phase_vec_max + old) * 3 + d] = cell[3 * i + d];
        }
    }
    return hash_count;
}

static float norm3(const float* v) { return sqrtf(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }

/* check_cell_distance :340-373 */
static int check_cell_distance(const BdParams* p, const int* ncell, int cur, const float* pos, const int* id, const float* tri_normal,
                               const int* sample_count, const int* sample, int sample_cap) {
    int count = 0, ret = 0;
    const int h = cell_hash(ncell, p->hash_size);
    while (count < sample_count[h] && ret == 0) {
        const int nb = sample[(size_t)h * sample_cap + count];
        float d[3] = {pos[3 * cur] - pos[3 * nb], pos[3 * cur + 1] - pos[3 * nb + 1], pos[3 * cur + 2] - pos[3 * nb + 2]};
        float dist = norm3(d);
        const int cid = id[cur], nid = id[nb];
        if (cid != nid) {
            const float invlen = 1.0f / norm3(d);                                /* normalized(): invlen * v */
            const float v[3] = {invlen * d[0], invlen * d[1], invlen * d[2]};
            const float* n1 = tri_normal + 3 * cid; const float* n2 = tri_normal + 3 * nid;   /* vertex-indexed array, face id (sic) */
            const float c1 = n1[0] * v[0] + n1[1] * v[1] + n1[2] * v[2];
            const float c2 = n2[0] * v[0] + n2[1] * v[1] + n2[2] * v[2];
            if (fabsf(c1 - c2) > 0.00001f) dist *= (asinf(c1) - asinf(c2)) / (c1 - c2);
            else dist /= sqrtf(1.0f - c1 * c1);
        }
        if (dist < p->radius) ret = 1;
        count++;
    }
    return ret;
}

/* possion_disk_sample :390-407, one launch = (phase group pg, trial); returns the new possion_sample count */
int oracle_bd_sample_launch(const BdParams* p, int pg, int trial, int pg_count, const int* phase_group, const int* cell, const float* pos,
                            const int* id, const float* tri_normal, const int* start_index, int* sample_count, int* sample, int sample_cap,
                            float* possion_sample, int* selected, int n_sample) {
    for (int t = 0; t < pg_count; t++) {
        const int* c = phase_group + ((size_t)pg * p->phase_vec_max + t) * 3;
        const int h = cell_hash(c, p->hash_size);
        const int cand = start_index[h] + trial;
        if (cand >= p->n) continue;
        if (compare_cell(cell, cand, start_index[h]) != 0) continue;
        int conflicts = 0;                                                        /* check_cell :376-386 */
        for (int a = -2; a < 3; a++) for (int b = -2; b < 3; b++) for (int k = -2; k < 3; k++) {
            const int nc[3] = {a + cell[3 * cand], b + cell[3 * cand + 1], k + cell[3 * cand + 2]};
            conflicts += check_cell_distance(p, nc, cand, pos, id, tri_normal, sample_count, sample, sample_cap);
        }
        if (conflicts != 0) continue;
        const int old = sample_count[h]++;
        if (old < sample_cap) sample[(size_t)h * sample_cap + old] = cand;
        else sample_count[h] = sample_cap - 1;
        for (int d = 0; d < 3; d++) possion_sample[3 * (size_t)n_sample + d] = pos[3 * cand + d];
        if (selected) selected[n_sample] = cand;
        n_sample++;
    }
    return n_sample;
}
