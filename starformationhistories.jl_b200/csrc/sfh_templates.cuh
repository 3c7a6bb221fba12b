// sfh_templates.cuh -- building the template stack ON the device.
//
// The reference builds every template with bin_cmd_smooth (src/StarFormationHistories.jl:574-621): one addstar!
// (:348-364 pixel-space kernel, :366-408 real-space kernel) per isochrone point, each adding the pixel-integrated
//   cov_mult == 0  : GaussianPSFAsymmetric (:223-266), exact integral gaussian_int_general (:198-205)
//   cov_mult == +-1: GaussianPSFCovariant (:272-338), 3-point Gauss-Legendre in y x erf in x (:303-333)
// to the pixels of its cut-out; the host then hcat's the T Hess diagrams (stack_models) and they would be uploaded.
// Here the T ragged point lists are scattered by ONE launch straight into device memory in the stack's own layout.
//
// Decomposition: one WARP owns (template, band of Hess rows) exclusively and walks that template's points in their
// given order; lanes take the pixels of the cut-out that fall inside the band.  No atomics: every pixel receives its
// contributions in point order -- the association of the reference's sequential loop (:585, :611) -- so the result
// is deterministic and independent of the launch geometry.  The separable kernel evaluates its 2(w + h) erf
// differences once per point into per-warp shared memory instead of 4 erf per pixel; the covariant kernel hoists its
// three per-row exponentials (Gauss-Legendre nodes in y) the same way and evaluates erf at pixel EDGES, shared between
// neighbouring pixels by shuffle (3 (w+1)/w instead of 6 per pixel).
#pragma once
#include <cstdint>
#include <cuda_runtime.h>

#include "sfh_small.cuh"

namespace sfh {

struct ScatterParams {
    int64_t nx, ny;                        // Hess bins (image is nx x ny column-major: bin = ix + nx * iy)
    double xfirst, xstep, yfirst, ystep;   // left edge of the first bin and the bin width, per axis
    const int64_t *offsets;                // [T+1]: template t owns points [offsets[t], offsets[t+1])
    const double *x, *y, *sx, *sy, *w;     // colours, magnitudes, colour errors, magnitude errors, weights
    const int32_t *cov;                    // [T] cov_mult per template (-1, 0, 1)
    int64_t t0, tc;                        // this launch covers templates [t0, t0 + tc)
    int32_t nbands, band_h;                // rows per band
    double *scratch;                       // [tc][nx * ny] FP64 images, zero-initialised
};

constexpr int kScatterWarps = 4;
#ifndef SFH_SCATTER_MINB
#define SFH_SCATTER_MINB 1
#endif

__device__ __forceinline__ int64_t clamp_to_i64(double v) {   // ceil/rint results of absurd widths stay defined
    return (v < 4.0e9) ? ((v > -4.0e9) ? (int64_t)v : (int64_t)-4000000000LL) : (int64_t)4000000000LL;
}

__global__ void __launch_bounds__(kScatterWarps * 32, SFH_SCATTER_MINB) sfh_templates_scatter_kernel(const ScatterParams p) {
    extern __shared__ double sm_f[];
    const int lane = threadIdx.x & 31, wl = threadIdx.x >> 5;
    const int64_t wid = (int64_t)blockIdx.x * kScatterWarps + wl;
    if (wid >= p.tc * p.nbands) return;
    const int64_t tl = wid / p.nbands;
    const int band = (int)(wid % p.nbands);
    double *fx = sm_f + (size_t)wl * (size_t)(p.nx + 6 * p.band_h);   // [nx] x-factors of the current point
    double *fy = fx + p.nx;                                            // [band_h] (separable kernel) / [band_h][3][2] (covariant)
    const int64_t jb0 = (int64_t)band * p.band_h + 1;              // 1-based rows of this band
    const int64_t jb1 = min(p.ny, (int64_t)(band + 1) * p.band_h);
    double *img = p.scratch + (size_t)tl * (size_t)(p.nx * p.ny);
    const int cov = p.cov[p.t0 + tl];
    const int64_t p0 = p.offsets[p.t0 + tl], p1 = p.offsets[p.t0 + tl + 1];
    const double s2 = sqrt(2.0);
    for (int64_t q = p0; q < p1; ++q) {
        const double xr = p.x[q], yr = p.y[q], ex = p.sx[q], ey = p.sy[q], A = p.w[q];
        if (cov == 0) {
            const double x0 = (xr - p.xfirst) / p.xstep + 1, y0 = (yr - p.yfirst) / p.ystep + 1;   // histogram_pix (:503)
            const double sx = ex / p.xstep, sy = ey / p.ystep;
            const int64_t xo = max((int64_t)1, clamp_to_i64(ceil(sx * 10)) / 2);                  // size(obj) (:247), :351-352
            const int64_t yo = max((int64_t)1, clamp_to_i64(ceil(sy * 10)) / 2);
            const int64_t xc = clamp_to_i64(rint(x0)), yc = clamp_to_i64(rint(y0));
            const int64_t xa = max((int64_t)1, xc - xo), xb = min(p.nx, xc + xo);
            const int64_t ya = max((int64_t)1, yc - yo), yb = min(p.ny, yc + yo);
            if (!(xb - xa + 1 > 1 && yb - ya + 1 > 1)) continue;                                  // :358
            const int64_t ja = max(ya, jb0), jb = min(yb, jb1);
            if (ja > jb) continue;
            const int64_t w = xb - xa + 1, h = jb - ja + 1;
            for (int64_t k = lane; k < w; k += 32) {
                const double dx = (double)(xa + k) + 0.5 - x0;
                fx[k] = erf((dx - 0.5) / (s2 * sx)) - erf((dx + 0.5) / (s2 * sx));
            }
            for (int64_t k = lane; k < h; k += 32) {
                const double dy = (double)(ja + k) + 0.5 - y0;
                fy[k] = erf((dy - 0.5) / (s2 * sy)) - erf((dy + 0.5) / (s2 * sy));
            }
            __syncwarp();
            const double a4 = A / 4;
            const uint32_t wu = (uint32_t)w, npx = (uint32_t)(w * h);   // 32-bit pixel arithmetic: a cut-out never exceeds nx * band_h
            for (uint32_t e = lane; e < npx; e += 32) {
                const uint32_t m = e / wu, k = e - m * wu;
                img[(xa + k - 1) + p.nx * (ja + m - 1)] += a4 * fx[k] * fy[m];
            }
            __syncwarp();
        } else {
            const int64_t xc = clamp_to_i64(rint((xr - p.xfirst) / p.xstep + 1)), yc = clamp_to_i64(rint((yr - p.yfirst) / p.ystep + 1));
            const int64_t xo = max((int64_t)1, clamp_to_i64(rint(15 * ex / p.xstep / 2)));        // size(obj) = (15 sx, 10 sy) (:296)
            const int64_t yo = max((int64_t)1, clamp_to_i64(rint(10 * ey / p.ystep / 2)));
            const int64_t xa = max((int64_t)1, xc - xo), xb = min(p.nx, xc + xo);
            const int64_t ya = max((int64_t)1, yc - yo), yb = min(p.ny, yc + yo);
            if (!(xb - xa + 1 > 1 && yb - ya + 1 > 1)) continue;                                  // :399
            const int64_t ja = max(ya, jb0), jb = min(yb, jb1);
            if (ja > jb) continue;
            const int64_t w = xb - xa + 1, h = jb - ja + 1;
            const double hx = p.xstep / 2, hy = p.ystep / 2;
            const double prefac = A / 2 / sqrt(2.0 * 3.14159265358979323846) / ey;
            const double cm = (double)cov;
            // per (row, Gauss-Legendre node): weight * exp(-(Dy/sy)^2 / 2) and the x shift Dy * cov_mult -- they do not
            // depend on the column, so they are evaluated once per row instead of once per pixel (:321-329)
            for (uint32_t k = lane; k < 3u * (uint32_t)h; k += 32) {
                const uint32_t m = k / 3u;
                const int g = (int)(k - 3u * m);
                const double gx = (g == 0) ? -0.7745966692414834 : (g == 1 ? 0.0 : 0.7745966692414834);
                const double gw = (g == 1) ? 0.8888888888888888 : 0.5555555555555556;
                const double yv = ((double)(ja + m) - 0.5) * p.ystep + p.yfirst;                  // histogram_data(j + 1/2) (:525)
                const double Dy = (gx * hy + yv) - yr;
                const double t = Dy / ey;
                fy[2 * k] = (gw * hy) * exp(-(t * t) / 2);
                fy[2 * k + 1] = Dy * cm;
            }
            __syncwarp();
            // erf(+(Dx + hx)/a) + erf((-Dx + hx)/a) of :326-327 is erf(right edge) - erf(left edge) of the pixel, and the right
            // edge of pixel k is the left edge of pixel k + 1: lanes evaluate EDGES (3 erf each, one per Gauss-Legendre node)
            // and take the neighbour's by shuffle -- 3 (w + 1) / w erf per pixel instead of 6.  Edges are flattened over the
            // rows of the cut-out; an iteration advances by 31 so that lane 31's edge is lane 0's of the next one.
            const uint32_t wu = (uint32_t)w, we = wu + 1, nedge = we * (uint32_t)h;
            const double inv_a = 1.0 / (s2 * ex);
            for (uint32_t base = 0; base < nedge; base += 31) {
                const uint32_t e = base + lane;
                const bool valid = e < nedge;
                const uint32_t m = valid ? e / we : 0u, j = valid ? e - m * we : 0u;
                const double xe = (((double)(xa + j) - 1.0) * p.xstep + p.xfirst) - xr;             // left edge of pixel xa + j
                double r = 0.0;
#pragma unroll
                for (int g = 0; g < 3; ++g) {
                    const double E = erf((xe + fy[2 * (3 * m + g) + 1]) * inv_a);
                    const double En = __shfl_down_sync(0xffffffffu, E, 1);
                    r += fy[2 * (3 * m + g)] * (En - E);
                }
                if (valid && lane < 31 && j < wu) img[(xa + j - 1) + p.nx * (ja + m - 1)] += r * prefac;
            }
            __syncwarp();
        }
    }
}

// scratch images (FP64, host layout) -> the stack's storage type and device layout, rows [row_begin, row_begin + rows)
template <typename S>
__global__ void sfh_templates_store_kernel(const double *__restrict__ scratch, S *__restrict__ M, const StackLayout lay, int64_t rows,
                                           int64_t row_begin, int64_t nb_total, int64_t t0, int64_t tc) {
    const int64_t n = rows * tc;
    for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
        const int64_t i = e % rows, tl = e / rows;
        M[lay.off(i, t0 + tl)] = (S)scratch[(size_t)tl * (size_t)nb_total + (size_t)(row_begin + i)];
    }
}

}  // namespace sfh
