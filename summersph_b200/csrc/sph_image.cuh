// sph_image.cuh — column-density image of the resident state.
//
// The reference's post-processing script Density_Image.py samples rho on a 120^3 grid with a fixed h = 1.25
// (Density_Image.py:105-141) and sums the grid along z (:145).  Here the line-of-sight integral is done
// analytically per particle: Sigma(a, b) = sum_j m_j F(|d_ab| / h_j) / (pi h_j^2), F(q_b) = integral of the M4
// shape w (SUMMER_SPH.f90:66,70) along the line of sight at impact parameter q_b, with the engine's own h_j
// (`smoothing` in fixed-h mode).  The script's image is this Sigma divided by its z spacing, in the limit of a
// fine grid.  Particles narrower than a pixel are widened to h = pixel/2 so that they always reach a pixel centre.
// One warp per particle scatters its footprint with FP64 atomics: an output kernel, not a step kernel.
#pragma once
#include "sph_common.cuh"
#include "sph_integrate.cuh"

#define IMG_TABLE 1024             // samples of F on q_b in [0, 2]
#define IMG_THREADS 256

__global__ void __launch_bounds__(IMG_THREADS)
k_column_density(int n, StateArrays s, int variable_h, double h_fixed, int axis, double u0, double v0, double du, double dv,
                 int nu, int nv, const double* __restrict__ F, double* __restrict__ img) {
  const int lane = threadIdx.x & 31;
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int nwarp = (gridDim.x * blockDim.x) >> 5;
  const double* A = axis == 0 ? s.y : (axis == 1 ? s.z : s.x);      // image abscissa / ordinate: x -> (y,z), y -> (z,x), z -> (x,y)
  const double* B = axis == 0 ? s.z : (axis == 1 ? s.x : s.y);
  const double hmin = 0.5 * fmax(du, dv);
  const double inv_dq = IMG_TABLE / 2.0;
  for (int i = warp; i < n; i += nwarp) {
    const double a = A[i], b = B[i];
    double h = variable_h ? s.h[i] : h_fixed;
    if (!(h > hmin)) h = hmin;
    const double R = 2.0 * h;
    // pixel centres u0 + (k + 1/2) du inside [a - R, a + R]
    const double fu0 = ceil((a - R - u0) / du - 0.5), fu1 = floor((a + R - u0) / du - 0.5);
    const double fv0 = ceil((b - R - v0) / dv - 0.5), fv1 = floor((b + R - v0) / dv - 0.5);
    if (!(fu1 >= 0.0 && fv1 >= 0.0 && fu0 <= nu - 1.0 && fv0 <= nv - 1.0)) continue;    // outside the frame (or NaN)
    const int iu0 = (int)fmax(fu0, 0.0), iu1 = (int)fmin(fu1, nu - 1.0);
    const int iv0 = (int)fmax(fv0, 0.0), iv1 = (int)fmin(fv1, nv - 1.0);
    const int w = iu1 - iu0 + 1, t = iv1 - iv0 + 1;
    if (w <= 0 || t <= 0) continue;
    const double inv_h = 1.0 / h;
    const double amp = s.m[i] * inv_h * inv_h * (1.0 / 3.14159265358979323846);
    const long long total = (long long)w * t;
    for (long long k = lane; k < total; k += 32) {
      const int iu = iu0 + (int)(k % w), iv = iv0 + (int)(k / w);
      const double da = u0 + (iu + 0.5) * du - a, db = v0 + (iv + 0.5) * dv - b;
      const double q = sqrt(da * da + db * db) * inv_h;
      if (q < 2.0) {
        const double x = q * inv_dq;
        int j = (int)x; if (j > IMG_TABLE - 1) j = IMG_TABLE - 1;
        const double f = x - j;
        atomicAdd(&img[(size_t)iv * nu + iu], amp * ((1.0 - f) * F[j] + f * F[j + 1]));
      }
    }
  }
}
