// sph_ics.cuh — seeded initial conditions generated on the device.
//
// Replaces what /root/reference/Disc_ICs.py:1-41 sketches (a host script that is broken as shipped: missing imports,
// undefined `r`): a uniform-surface-density Keplerian disc in the reference's units and file semantics
// (x y z vx vy vz u m alpha h; G = the real(4) literal of SUMMER_SPH.f90:7).  Counter-based: row i depends only on
// (seed, i), so any rank generates any slice of the rows without the others - a 16M-particle start no longer
// costs every rank a 16M-row numpy pass on the host.
#pragma once
#include "sph_common.cuh"

__device__ __forceinline__ double ics_uniform(uint64_t seed, uint64_t i, uint64_t stream) {       // (0, 1)
  const uint64_t z = mix64(mix64(seed ^ (stream * 0xD1B54A32D192ED03ull)) + i * 0x9E3779B97F4A7C15ull);
  return ((double)(z >> 11) + 0.5) * (1.0 / 9007199254740992.0);
}
struct IcsDisc { double r_in, r_out, aspect, m_star, m_disc, u, alpha, eta, G; };
__global__ void k_ics_disc(long long n_global, long long first, int n_local, uint64_t seed, IcsDisc P,
                           double* __restrict__ x, double* __restrict__ y, double* __restrict__ z,
                           double* __restrict__ vx, double* __restrict__ vy, double* __restrict__ vz,
                           double* __restrict__ u, double* __restrict__ m, double* __restrict__ alpha, double* __restrict__ h) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= n_local) return;
  const uint64_t i = (uint64_t)(first + t);
  const double two_pi = 6.283185307179586;
  const double r = sqrt(P.r_in * P.r_in + (P.r_out * P.r_out - P.r_in * P.r_in) * ics_uniform(seed, i, 1));      // uniform surface density
  const double phi = two_pi * ics_uniform(seed, i, 2);
  double zeta = sqrt(-2.0 * log(ics_uniform(seed, i, 3))) * cos(two_pi * ics_uniform(seed, i, 4));               // N(0, 1), clipped at 3 sigma
  zeta = fmin(fmax(zeta, -3.0), 3.0);
  const double H = P.aspect * r, s = sin(phi), c = cos(phi);
  const double vphi = sqrt(P.G * P.m_star / r);
  const double mi = P.m_disc / (double)n_global;
  const double sigma = P.m_disc / (3.141592653589793 * (P.r_out * P.r_out - P.r_in * P.r_in));
  const double rho = sigma / (sqrt(two_pi) * H) * exp(-0.5 * zeta * zeta);
  x[t] = r * c; y[t] = r * s; z[t] = zeta * H;
  vx[t] = -vphi * s; vy[t] = vphi * c; vz[t] = 0.0;
  u[t] = P.u; m[t] = mi; alpha[t] = P.alpha; h[t] = P.eta * cbrt(mi / rho);
}
