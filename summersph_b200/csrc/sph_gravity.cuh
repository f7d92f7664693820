// sph_gravity.cuh — Barnes-Hut gravity walk + direct sink gravity.
//
// Replaces particle_gravforces / particle_gravforce_one (SUMMER_SPH.f90:249-290 | Variable.f90:270-311)
// and sink_gravforces (F:559-591 | V:691-726).
//
// The reference's opening decision is per particle: `size/dist < theta .or. no children` with
// dist = sqrt(|x - COM|^2 + 0.001*smoothing) (F:275-278).  To keep every accepted-node set identical, the
// decision stays per lane; what changes is the traversal: one warp owns 32 Morton-adjacent particles and
// walks the depth-first-preorder node array in lock step.  A lane that accepted node v sets its private
// skip index to next[v] and sleeps until the walk leaves v's subtree; the warp descends (v+1) while any
// lane still wants to open, else jumps to next[v].  Node records are warp-uniform 48-byte loads.
// The visit order per lane is exactly the reference's recursion order (children 1..8), so the
// accumulation order of the gravity terms matches the reference.
#pragma once
#include "sph_common.cuh"
#include "sph_walk.cuh"

struct SinkArrays { double *x, *y, *z, *vx, *vy, *vz, *m, *radius, *ax, *ay, *az; };

struct NodeRegs { double cx, cy, cz, m, size; int next, flags; };
__device__ __forceinline__ NodeRegs load_node(const GNode* __restrict__ nodes, int v) {
  const double2* p = reinterpret_cast<const double2*>(nodes + v);
  const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
  NodeRegs r; r.cx = a.x; r.cy = a.y; r.cz = b.x; r.m = b.y; r.size = c.x;
  r.next = __double2loint(c.y); r.flags = __double2hiint(c.y);
  return r;
}

// layout of dynamic smem: grav table (nq+1 doubles)
__global__ void __launch_bounds__(256, 4)
k_gravity(int p_begin, int p_end, DevParams P, const GNode* __restrict__ nodes, int n_nodes,
          const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
          const double* __restrict__ h, const double* __restrict__ m, const double* __restrict__ g_gt,
          double* __restrict__ ax, double* __restrict__ ay, double* __restrict__ az,
          int do_grav, int n_sink, SinkArrays S, double* __restrict__ sink_partial, WalkCounters* ctr) {
  extern __shared__ double gt[];
  for (int i = threadIdx.x; i <= P.nq; i += blockDim.x) gt[i] = g_gt[i];
  __syncthreads();
  const int i = p_begin + blockIdx.x * blockDim.x + threadIdx.x;
  const int lane = threadIdx.x & 31;
  const bool live = i < p_end;
  const double xi = live ? x[i] : 0.0, yi = live ? y[i] : 0.0, zi = live ? z[i] : 0.0;
  const double hi = live ? (P.variable_h ? h[i] : P.h_fixed) : 1.0;
  const double inv_h = 1.0 / hi;
  const double soft = P.soft_hi ? 0.001 * hi : 0.001 * P.h_fixed;          // F:275 | V:296 | T:298
  const double theta = P.theta, theta2 = theta * theta;
  double gx = 0.0, gy = 0.0, gz = 0.0;
  unsigned opened = 0, accepted = 0;
  if (do_grav) {
    int cur = 0;
    int skip = live ? 0 : 0x7fffffff;
    while (cur < n_nodes) {
      const NodeRegs N = load_node(nodes, cur);
      bool open = false;
      if (cur >= skip) {
        const double dx = xi - N.cx, dy = yi - N.cy, dz = zi - N.cz;          // F:274
        const double d2 = dx * dx + dy * dy + dz * dz + soft;
        bool accept;
        if (N.flags & 1) accept = true;                                       // .not. allocated(children)
        else {
          const double s2 = N.size * N.size, t2 = theta2 * d2;
          if (s2 < t2 * (1.0 - 1e-12)) accept = true;
          else if (s2 > t2 * (1.0 + 1e-12)) accept = false;
          else {   // borderline: redo the reference's arithmetic exactly (no contraction)       F:275-278
            double e2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)), soft);
            accept = __ddiv_rn(N.size, __dsqrt_rn(e2)) < theta;
          }
        }
        if (accept) {
          ++accepted;
          skip = N.next;
          if (N.m > 0.0 && d2 > 0.0) {                                        // F:279
            double dist, rs; fast_sqrt_rsqrt(d2, dist, rs);
            const double q = dist * inv_h;
            double W = 1.0;
            if (q <= 2.0) W = table_lerp1(gt, P.nq, P.dq, P.inv_dq, q);       // F:129-146
            const double f = (P.G * N.m * W) * (rs * rs * rs);                // F:281
            gx -= f * dx; gy -= f * dy; gz -= f * dz;
          }
        } else { open = true; ++opened; }
      }
      cur = __any_sync(FULL_MASK, open) ? cur + 1 : N.next;
    }
  }
  // direct sink <-> gas (unsoftened) F:567-576; per-warp partial sums of the sink side (no block barrier)
  const int gwarp = (i - p_begin) >> 5;
  for (int s = 0; s < n_sink; ++s) {
    const double vx_ = xi - S.x[s], vy_ = yi - S.y[s], vz_ = zi - S.z[s];
    const double dr = sqrt(vx_ * vx_ + vy_ * vy_ + vz_ * vz_);
    const double d3 = dr * dr * dr;
    double wx = P.G * vx_ / d3, wy = P.G * vy_ / d3, wz = P.G * vz_ / d3;    // F:572
    const double ms = S.m[s];
    double px = 0.0, py = 0.0, pz = 0.0;
    if (live) {
      gx -= ms * wx; gy -= ms * wy; gz -= ms * wz;                            // F:574
      const double mi = m[i];
      px = mi * wx; py = mi * wy; pz = mi * wz;                               // F:573
    }
    px = warp_sum(px); py = warp_sum(py); pz = warp_sum(pz);
    if (lane == 0) {
      double* o = sink_partial + ((size_t)gwarp * n_sink + s) * 3;
      o[0] = px; o[1] = py; o[2] = pz;
    }
  }
  if (live) { ax[i] = gx; ay[i] = gy; az[i] = gz; }
  opened = (unsigned)warp_sum_ll(opened); accepted = (unsigned)warp_sum_ll(accepted);
  if (lane == 0 && do_grav) { atomicAdd(&ctr->grav_opened, (unsigned long long)opened); atomicAdd(&ctr->grav_accepted, (unsigned long long)accepted); }
}

// fold per-warp sink partials in a fixed order (deterministic). One block of 256.
__global__ void k_sink_reduce(int n_parts, int n_sink, const double* __restrict__ partial, SinkArrays S, int do_sinks) {
  __shared__ double red[256];
  const int t = threadIdx.x;
  for (int s = 0; s < n_sink; ++s)
    for (int k = 0; k < 3; ++k) {
      double v = 0.0;
      if (do_sinks) for (int b = t; b < n_parts; b += 256) v += partial[((size_t)b * n_sink + s) * 3 + k];
      red[t] = v;
      __syncthreads();
      for (int o = 128; o > 0; o >>= 1) { if (t < o) red[t] += red[t + o]; __syncthreads(); }
      if (t == 0) { double* a = (k == 0) ? S.ax : (k == 1) ? S.ay : S.az; a[s] = red[0]; }
      __syncthreads();
    }
}
// sink-sink pairs F:578-590 (after the cross-rank sum of the gas contributions)
__global__ void k_sink_pairs(int n_sink, SinkArrays S, double G, int do_sinks) {
  if (threadIdx.x != 0 || !do_sinks || n_sink < 2) return;
  for (int i = 0; i < n_sink; ++i)
    for (int j = 0; j < i; ++j) {
      double v[3] = {S.x[j] - S.x[i], S.y[j] - S.y[i], S.z[j] - S.z[i]};
      double dr = sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
      double d3 = dr * dr * dr;
      double* A[3] = {S.ax, S.ay, S.az};
      for (int k = 0; k < 3; ++k) {
        double w = G * v[k] / d3;
        A[k][i] += S.m[j] * w;
        A[k][j] -= S.m[i] * w;
      }
    }
}
