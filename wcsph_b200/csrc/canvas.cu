// canvas.cu -- SURVEY §8(f) N1: the particle-splat canvas the reference's step loops draw into
// (Canvas.py:138-209 + the scripts' draw_particle kernels dfsph.py:585-593, sesph.py:201-207),
// rendered straight from the cell-sorted device positions.
//
// The reference's depth test (`if depth[v] > z: img[v] = c; depth[v] = z`, Canvas.py:143-148) runs inside a
// parallel loop without atomics; its defined meaning is the serial one: per pixel the fragment with the
// smallest z wins, the first in loop order on ties.  Here a pixel is one 64-bit word
//     key = (ordered_bits(z) << 2) | colour_code        (0 = liquid white, 1 = point grey, 3 = background)
// resolved with atomicMin: liquids' outlines precede every grey point in both loop orders of the reference, so
// "smaller colour code on equal z" is exactly "first in loop order", and the order particles are visited in
// (cell-sorted here) does not matter.  The transform is evaluated without FMA contraction and with IEEE
// division so that the pixel a particle lands in is bit-identical to the CPU restatement's.
#include "engine.cuh"

__device__ __forceinline__ unsigned int ordered_bits(float z) {
    const unsigned int u = __float_as_uint(z);
    return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}
__device__ __forceinline__ float from_ordered_bits(unsigned int o) {
    return __uint_as_float((o & 0x80000000u) ? (o & 0x7fffffffu) : ~o);
}
#define CANVAS_BACKGROUND(one_bits) ((((unsigned long long)(one_bits)) << 2) | 3ull)

struct Mat4 { float m[16]; };

// Canvas.py:205-209
__global__ void k_canvas_clear(unsigned long long* __restrict__ zbuf, int npix) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p < npix) zbuf[p] = CANVAS_BACKGROUND(ordered_bits(1.0f));
}

// Canvas.py:138-141; returns false for coordinates whose i32 cast is undefined (not drawn, like the restatement)
__device__ __forceinline__ bool canvas_transform(const Mat4& M, float x, float y, float z, int sx, int sy, float& ox, float& oy, float& oz) {
    float s[4];
#pragma unroll
    for (int i = 0; i < 4; i++) {
        float acc = __fmul_rn(M.m[4 * i], x);
        acc = __fadd_rn(acc, __fmul_rn(M.m[4 * i + 1], y));
        acc = __fadd_rn(acc, __fmul_rn(M.m[4 * i + 2], z));
        acc = __fadd_rn(acc, M.m[4 * i + 3]);
        s[i] = acc;
    }
    ox = __fmul_rn(__fmul_rn(__fadd_rn(__fdiv_rn(s[0], s[3]), 1.0f), 0.5f), (float)sx);
    oy = __fmul_rn(__fmul_rn(__fadd_rn(__fdiv_rn(s[1], s[3]), 1.0f), 0.5f), (float)sy);
    oz = __fdiv_rn(s[2], s[3]);
    return isfinite(ox) && isfinite(oy) && fabsf(ox) < 1.0e9f && fabsf(oy) < 1.0e9f;
}

// Canvas.py:143-148
__device__ __forceinline__ void canvas_fill_pixel(unsigned long long* zbuf, int sx, int sy, int px, int py, float z, unsigned long long key) {
    if (px >= 0 && px < sx && py >= 0 && py < sy && z < 1.0f) {
        unsigned long long* at = zbuf + (size_t)px * sy + py;
        if (key < *at) atomicMin(at, key);            // the plain read filters most fragments
    }
}

// one thread per drawn particle: slots [a0, a0+na) are liquids (outline, Canvas.py:150-179, and -- style 1 --
// a grey centre point, dfsph.py:591-593), slots [b0, b0+nb) solids (grey point, Canvas.py:197-201)
__global__ void __launch_bounds__(WCSPH_BLOCK)
k_canvas_draw(const float4* __restrict__ pos, int a0, int na, int b0, int nb, Mat4 M, int sx, int sy, int style,
              unsigned long long* __restrict__ zbuf) {
    const int t = blockIdx.x * blockDim.x + threadIdx.x;
    if (t >= na + nb) return;
    const bool liquid = t < na;
    const float4 p = pos[liquid ? a0 + t : b0 + (t - na)];
    float fx, fy, fz;
    if (!canvas_transform(M, p.x, p.y, p.z, sx, sy, fx, fy, fz)) return;
    const int xc = (int)fx, yc = (int)fy;
    const unsigned long long zkey = ((unsigned long long)ordered_bits(fz)) << 2;
    if (liquid) {
        int x = 0, y = 3, d = 3 - 2 * 3;
        while (x <= y) {
            // the reference issues all eight octant writes; those that repeat a pixel (x == 0, x == y) are no-ops there
            const bool mx = x != 0, sw = x != y;
            canvas_fill_pixel(zbuf, sx, sy, xc + x, yc + y, fz, zkey);
            if (mx) canvas_fill_pixel(zbuf, sx, sy, xc - x, yc + y, fz, zkey);
            canvas_fill_pixel(zbuf, sx, sy, xc + x, yc - y, fz, zkey);
            if (mx) canvas_fill_pixel(zbuf, sx, sy, xc - x, yc - y, fz, zkey);
            if (sw) {
                canvas_fill_pixel(zbuf, sx, sy, xc + y, yc + x, fz, zkey);
                canvas_fill_pixel(zbuf, sx, sy, xc - y, yc + x, fz, zkey);
                if (mx) {
                    canvas_fill_pixel(zbuf, sx, sy, xc + y, yc - x, fz, zkey);
                    canvas_fill_pixel(zbuf, sx, sy, xc - y, yc - x, fz, zkey);
                }
            }
            if (d < 0) d = d + 4 * x + 6;
            else { d = d + 4 * (x - y) + 10; y = y - 1; }
            x += 1;
        }
    }
    if (!liquid || style == 1) canvas_fill_pixel(zbuf, sx, sy, xc, yc, fz, zkey | 1ull);
}

// gui.set_image(sph_canvas.img.to_numpy()) (dfsph.py:623): key -> img[sx][sy][3], depth[sx][sy]
__global__ void k_canvas_resolve(const unsigned long long* __restrict__ zbuf, int npix, float* __restrict__ img, float* __restrict__ depth) {
    const int p = blockIdx.x * blockDim.x + threadIdx.x;
    if (p >= npix) return;
    const unsigned long long key = zbuf[p];
    const unsigned int code = (unsigned int)(key & 3ull);
    const float c = code == 0u ? 1.0f : (code == 1u ? 0.3f : 0.0f);
    img[3 * (size_t)p] = c; img[3 * (size_t)p + 1] = c; img[3 * (size_t)p + 2] = c;
    if (depth) depth[p] = from_ordered_bits((unsigned int)(key >> 2));
}

static int canvas_args(wcsph_ctx* c, const void* zbuf, int sx, int sy, const char* fn) {
    if (!c || !zbuf || sx <= 0 || sy <= 0 || (long long)sx * sy > (1ll << 30)) { wcsph_set_error("%s: null ctx / buffer or bad size", fn); return WCSPH_EINVAL; }
    return 0;
}

extern "C" int wcsph_canvas_clear(wcsph_ctx* c, unsigned long long* zbuf_dev, int sx, int sy) {
    TRY(canvas_args(c, zbuf_dev, sx, sy, __func__));
    prof_begin(c, "k_canvas_clear"); k_canvas_clear<<<nblocks(sx * sy), WCSPH_BLOCK, 0, c->stream>>>(zbuf_dev, sx * sy); prof_end(c); LAUNCH_CHECK(c);
    return 0;
}

extern "C" int wcsph_canvas_draw_particle(wcsph_ctx* c, const float* view16, const float* proj16, int sx, int sy, int style,
                                          unsigned long long* zbuf_dev) {
    TRY(canvas_args(c, zbuf_dev, sx, sy, __func__));
    if (!view16 || !proj16 || (style != 0 && style != 1)) { wcsph_set_error("%s: null matrix or style not in {0,1}", __func__); return WCSPH_EINVAL; }
    if (!c->uploaded) { wcsph_set_error("%s: no positions uploaded", __func__); return WCSPH_EINVAL; }
    Mat4 M;                                              // proj[0] @ view[0] (Canvas.py:139), k ascending, fp32
    for (int i = 0; i < 4; i++)
        for (int j = 0; j < 4; j++) {
            volatile float acc = proj16[4 * i] * view16[j];
            for (int k = 1; k < 4; k++) { volatile float t = proj16[4 * i + k] * view16[4 * k + j]; acc = acc + t; }
            M.m[4 * i + j] = acc;
        }
    const int n = c->nown + c->NS;                       // a slab rank draws its own liquids; solids are replicated
    prof_begin(c, "k_canvas_draw");
    k_canvas_draw<<<nblocks(n), WCSPH_BLOCK, 0, c->stream>>>(fcur<float4>(c, "pos"), c->i0, c->nown, c->SB, c->NS, M, sx, sy, style, zbuf_dev);
    prof_end(c); LAUNCH_CHECK(c);
    return 0;
}

extern "C" int wcsph_canvas_resolve(wcsph_ctx* c, const unsigned long long* zbuf_dev, int sx, int sy, float* img_dev, float* depth_dev) {
    TRY(canvas_args(c, zbuf_dev, sx, sy, __func__));
    if (!img_dev) { wcsph_set_error("%s: null image", __func__); return WCSPH_EINVAL; }
    prof_begin(c, "k_canvas_resolve"); k_canvas_resolve<<<nblocks(sx * sy), WCSPH_BLOCK, 0, c->stream>>>(zbuf_dev, sx * sy, img_dev, depth_dev); prof_end(c); LAUNCH_CHECK(c);
    return 0;
}
