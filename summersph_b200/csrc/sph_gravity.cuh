// sph_gravity.cuh — Barnes-Hut gravity walk + direct sink gravity.
//
// Replaces particle_gravforces / particle_gravforce_one (SUMMER_SPH.f90:249-290 | Variable.f90:270-311)
// and sink_gravforces (F:559-591 | V:691-726).
//
// The reference's opening decision is per particle: `size/dist < theta .or. no children` with
// dist = sqrt(|x - COM|^2 + 0.001*smoothing) (F:275-278), and every accepted-node set stays identical to it.
// What changes is who decides.  One warp owns a run of 32 Morton-consecutive particles (lane = particle; the runs
// restart at every 64th walk group, see k_seg_chunks) and keeps a stack of (node, lane mask) entries in shared
// memory.  Each trip pops up to 32 entries and classifies them lane-parallel against the run's bounding box
// (lane = node, one coalesced 48-byte load each, child blocks are contiguous in the walk layout) - in FP32 on
// box-relative coordinates with 3e-5 margins, since anything uncertain is simply "mixed":
//   * every particle of the run accepts (even the closest point of the box passes the test, or the node is
//     childless)                                   -> interaction list, mask unchanged;
//   * every particle opens (even the farthest point fails it)  -> push the child block, mask unchanged;
//   * otherwise the node is mixed: each particle in the mask runs the reference's own test (FP32 screen, then
//     the FP64 form with a 2e-12 guard band, then the exact uncontracted arithmetic inside the band); the node's
//     lane collects the two ballots: the accepting lanes become the list entry's mask, the opening lanes the
//     mask of the pushed children.  One list insert and one child push per trip serve all three classes.
// The interaction list (COM, G*M, mask) is evaluated in batches with lane = particle (18 FP64 instructions per
// entry, two entries in flight per lane), node data read once per run instead of once per particle.  The
// accumulation order differs from the reference's recursion order (rounding-level only).
#pragma once
#include "sph_common.cuh"
#include "sph_walk.cuh"

struct SinkArrays { double *x, *y, *z, *vx, *vy, *vz, *m, *radius, *ax, *ay, *az; };

#ifndef GW_WARPS
#define GW_WARPS 20         // warps per block
#endif
#ifndef GW_ILP
#define GW_ILP 2            // interaction-list entries in flight per lane
#endif
#ifndef GW_SUBLISTS
#define GW_SUBLISTS 1       // experiment (scripts/build_variant.sh sub4 "-DGW_SUBLISTS=4"): the interaction list is read through
                            // GW_SUBLISTS index lists, one per group of 32 / GW_SUBLISTS consecutive lanes, each holding only the
                            // entries whose mask touches that group, so a sparse entry costs evaluation slots only in the lane
                            // groups it belongs to.  Every lane still adds its own terms in list order.  1 = one list for the warp.
#endif
#define GW_STACK 384        // (node, mask) entries per warp in shared memory
#ifndef GW_LIST
#define GW_LIST  64         // interaction-list entries per warp (evaluated when fewer than 32 slots are left); <= 256 with GW_SUBLISTS
#endif
#define GW_SPILL 8192       // per-warp overflow entries in global memory (never reached in practice; loud if it is)

#ifdef GW_DEBUG
#ifndef GW_DBG_T
#define GW_DBG_T 16
#endif
#ifndef GW_DBG_RC
#define GW_DBG_RC 64
#endif
__device__ unsigned long long gw_dbg[16];
__device__ unsigned long long gw_hist[33];
#define GWD(i, v) do { if (lane == 0) atomicAdd(&gw_dbg[i], (unsigned long long)(v)); } while (0)
__global__ void k_gw_dbg_print() {
  printf("GWDBG trips %llu popped %llu A %llu O %llu M %llu evals %llu entries %llu lanework %llu spills %llu maxsn %llu mixacc %llu mixopen %llu\n",
         gw_dbg[0], gw_dbg[1], gw_dbg[2], gw_dbg[3], gw_dbg[4], gw_dbg[5], gw_dbg[6], gw_dbg[7], gw_dbg[8], gw_dbg[9], gw_dbg[10], gw_dbg[11]);
  for (int i = 0; i < 16; ++i) gw_dbg[i] = 0;
}
#else
#define GWD(i, v)
#endif

struct GravWarpSmem {
  int2     stack[GW_STACK];
  double2  lxy[GW_LIST], lzg[GW_LIST];     // (cx, cy), (cz, G*M)
  unsigned lmask[GW_LIST];
  double   mcx[32], mcy[32], mcz[32], msize[32], mlo[32], mhi[32];   // mixed nodes of the current trip: COM, size, d2 band of the cheap FP64 test
  float4   mf[32];                                                   // the same nodes for the FP32 screen: COM relative to the run's origin, size^2 / theta^2
  unsigned mmask[32];
#if GW_SUBLISTS > 1
  unsigned char sub[GW_SUBLISTS][GW_LIST];   // per lane group: positions (in lxy / lzg / lmask) of the entries that touch the group
#endif
};

// 1/sqrt(x): MUFU seed (~2^-20) + one Halley step (cubic: ~2^-58), x > 0
__device__ __forceinline__ double fast_rsqrt(double x) {
  double y;
  asm("rsqrt.approx.ftz.f64 %0, %1;" : "=d"(y) : "d"(x));
  const double e = fma(-(x * y), y, 1.0);
  const double t = fma(0.375, e, 0.5);
  return fma(y * e, t, y);
}

// one list entry against one particle: a -= G M g(dist/h) dir / dist^3                  F:279-281 | F:129-146
// h2x4 = 4 h^2: dist/h <= 2 is decided on the squares (g(2) = 1 and the table is continuous there, so which side
// of the branch a borderline pair takes changes the term by rounding only); W = 1 beyond it needs no multiply.
__device__ __forceinline__ void grav_term(const double2 a, const double2 b, const bool on, const double xi, const double yi,
                                          const double zi, const double inv_h, const double h2x4, const double soft, const double* __restrict__ gt,
                                          const int nq, const double dq, const double inv_dq, double& gx, double& gy, double& gz) {
  const double dx = xi - a.x, dy = yi - a.y, dz = zi - b.x;                  // F:274
  const double d2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, soft)));
  const double rs = fast_rsqrt(d2);                                          // F:279: M > 0 is checked when the entry is listed; d2 >= soft > 0
  double gm = b.y;
  if (d2 <= h2x4) gm *= table_lerp1(gt, nq, dq, inv_dq, (d2 * rs) * inv_h);
  const double f = gm * (rs * rs * rs);
  if (on) { gx = fma(-f, dx, gx); gy = fma(-f, dy, gy); gz = fma(-f, dz, gz); }
}

#ifdef GW_FAR_REUSE
// Experiment (scripts/r2_gravity_variants.sh): evaluation A of a loop body sees the positions of the previous
// body's evaluation B (F:894 after F:905-912: the second kick moves no particle), so every accepted (node, particle) pair
// and its distance are the same; only h changed (calc_smoothing, V:1152), and h enters a gravity term only through
// g(dist / h) for dist < 2h (F:138-141: W = 1 beyond).  Pass 0 (a full walk) therefore adds the terms with
// dist^2 <= 4 hcut^2 (hcut = GW_HCUT * h = 1.1 h, kept per particle) and all the others apart and stores the far sum together with
// the sink terms; pass 1 (the next evaluation on the same tree, same sinks, every h <= hcut) walks only what can hold a
// near term - subtrees whose cell is farther than 2 max(hcut) from the run's box are dropped - re-evaluates the near terms
// with the new h and adds the stored far sum.  Same terms as the full walk, summed in another order (rounding level).
#ifndef GW_HCUT
#define GW_HCUT 1.1
#endif
#ifndef GW_NEAR_CAP
#define GW_NEAR_CAP 1024    // (node, lane mask) pairs kept per run for pass 1; a run that needs more falls back to the near walk; 0: always walk
#endif
// near / near_cnt: pass 0 also records, per run, every listed entry that can be near for some particle of the run; pass 1 then
// evaluates that list instead of walking (the walk has to visit nearly every node of the full walk just to drop it)
struct GravFar { int pass; double *fx, *fy, *fz, *hcut; int2* near; int* near_cnt; int near_cap; };
#define GW_FAR_PARAM , GravFar FR
__device__ __forceinline__ void grav_term_split(const double2 a, const double2 b, const bool on, const double xi, const double yi,
                                                const double zi, const double inv_h, const double h2x4, const double hc2x4, const int pass,
                                                const double soft, const double* __restrict__ gt, const int nq, const double dq,
                                                const double inv_dq, double& gx, double& gy, double& gz, double& qx, double& qy, double& qz) {
  const double dx = xi - a.x, dy = yi - a.y, dz = zi - b.x;
  const double d2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, soft)));
  if (d2 <= hc2x4) {                           // near class: the only terms that can depend on h
    const double rs = fast_rsqrt(d2);
    double gm = b.y;
    if (d2 <= h2x4) gm *= table_lerp1(gt, nq, dq, inv_dq, (d2 * rs) * inv_h);
    const double f = gm * (rs * rs * rs);
    if (on) { qx = fma(-f, dx, qx); qy = fma(-f, dy, qy); qz = fma(-f, dz, qz); }
  } else if (pass == 0) {
    const double rs = fast_rsqrt(d2);
    const double f = b.y * (rs * rs * rs);
    if (on) { gx = fma(-f, dx, gx); gy = fma(-f, dy, gy); gz = fma(-f, dz, gz); }
  }
}
// any h of this rank's slice above its cutoff voids the stored far sums (the near class would miss terms)
__global__ void k_check_hcut(int p0, int p1, const double* __restrict__ h, const double* __restrict__ hcut, int* flag) {
  const int i = p0 + blockIdx.x * blockDim.x + threadIdx.x;
  if (i < p1 && !(h[i] <= hcut[i])) *flag = 1;
}
#else
#define GW_FAR_PARAM
#endif

// dynamic smem: grav table (nq+1 doubles, padded to even) then one GravWarpSmem per warp
__global__ void __launch_bounds__(GW_WARPS * 32, 1)
k_gravity(int g_begin, int g_end, const int2* __restrict__ groups, const BvhBox* __restrict__ gbox, DevParams P,
          const WNode* __restrict__ wn, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
          const double* __restrict__ h, const double* __restrict__ m, const double* __restrict__ g_gt,
          double* __restrict__ ax, double* __restrict__ ay, double* __restrict__ az,
          int do_grav, int n_sink, SinkArrays S, double* __restrict__ sink_partial, WalkCounters* ctr, int* work,
          int2* __restrict__ spill, int* err_flag GW_FAR_PARAM) {
  extern __shared__ __align__(16) double gsm[];
  double* gt = gsm;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  GravWarpSmem& W = reinterpret_cast<GravWarpSmem*>(gsm + (P.nq + 1) + ((P.nq + 1) & 1))[warp];
  for (int i = threadIdx.x; i <= P.nq; i += blockDim.x) gt[i] = g_gt[i];
  __syncthreads();
  int2* myspill = spill + (size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * GW_SPILL;
  const unsigned lt_mask = (1u << lane) - 1u;
  const double theta = P.theta, theta2 = theta * theta, inv_theta2 = 1.0 / theta2;
  unsigned long long n_open = 0, n_acc = 0;

  for (;;) {
    int chunk = 0;
    if (lane == 0) chunk = atomicAdd(work, 1);
    chunk = __shfl_sync(FULL_MASK, chunk, 0);
    if (chunk >= g_end) break;
    const int2 tg = groups[chunk];
    const int i = tg.x + lane;
    const bool live = lane < tg.y;
    const double xi = live ? x[i] : 0.0, yi = live ? y[i] : 0.0, zi = live ? z[i] : 0.0;
    const double hi = live ? (P.variable_h ? h[i] : P.h_fixed) : 1.0;
    const double inv_h = 1.0 / hi, h2x4 = 4.0 * hi * hi;
    const double soft = P.soft_hi ? 0.001 * hi : 0.001 * P.h_fixed;          // F:275 | V:296 | T:298
    double gx[GW_ILP], gy[GW_ILP], gz[GW_ILP];
#pragma unroll
    for (int u = 0; u < GW_ILP; ++u) gx[u] = gy[u] = gz[u] = 0.0;

#ifdef GW_FAR_REUSE
#if GW_SUBLISTS > 1
#error "GW_FAR_REUSE and GW_SUBLISTS are separate experiments"
#endif
    const double hc = live ? (FR.pass == 0 ? GW_HCUT * hi : FR.hcut[i]) : 1.0;
    const double hc2x4 = 4.0 * hc * hc;
    double qx[GW_ILP], qy[GW_ILP], qz[GW_ILP];      // near terms
#pragma unroll
    for (int u = 0; u < GW_ILP; ++u) qx[u] = qy[u] = qz[u] = 0.0;
#endif
#if GW_SUBLISTS > 1
    int sl[GW_SUBLISTS];                       // entries in each lane group's index list (warp-uniform)
#pragma unroll
    for (int q = 0; q < GW_SUBLISTS; ++q) sl[q] = 0;
    const int myq = lane / (32 / GW_SUBLISTS);
    auto evaluate_list = [&](int) {
      int mylen = 0, maxlen = 0;
#pragma unroll
      for (int q = 0; q < GW_SUBLISTS; ++q) { if (q == myq) mylen = sl[q]; maxlen = sl[q] > maxlen ? sl[q] : maxlen; sl[q] = 0; }
      const unsigned char* mysub = W.sub[myq];
      int k = 0;
      for (; k + GW_ILP <= maxlen; k += GW_ILP) {
#pragma unroll
        for (int u = 0; u < GW_ILP; ++u) {
          const bool have = k + u < mylen;
          const int idx = have ? (int)mysub[k + u] : 0;
          grav_term(W.lxy[idx], W.lzg[idx], have && ((W.lmask[idx] >> lane) & 1u), xi, yi, zi, inv_h, h2x4, soft, gt, P.nq, P.dq, P.inv_dq, gx[u], gy[u], gz[u]);
        }
      }
      for (; k < maxlen; ++k) {
        const bool have = k < mylen;
        const int idx = have ? (int)mysub[k] : 0;
        grav_term(W.lxy[idx], W.lzg[idx], have && ((W.lmask[idx] >> lane) & 1u), xi, yi, zi, inv_h, h2x4, soft, gt, P.nq, P.dq, P.inv_dq, gx[0], gy[0], gz[0]);
      }
    };
#elif defined(GW_FAR_REUSE)
    auto evaluate_list = [&](int cnt) {
      int k = 0;
      for (; k + GW_ILP <= cnt; k += GW_ILP) {
#pragma unroll
        for (int u = 0; u < GW_ILP; ++u)
          grav_term_split(W.lxy[k + u], W.lzg[k + u], (W.lmask[k + u] >> lane) & 1u, xi, yi, zi, inv_h, h2x4, hc2x4, FR.pass, soft, gt,
                          P.nq, P.dq, P.inv_dq, gx[u], gy[u], gz[u], qx[u], qy[u], qz[u]);
      }
      for (; k < cnt; ++k)
        grav_term_split(W.lxy[k], W.lzg[k], (W.lmask[k] >> lane) & 1u, xi, yi, zi, inv_h, h2x4, hc2x4, FR.pass, soft, gt,
                        P.nq, P.dq, P.inv_dq, gx[0], gy[0], gz[0], qx[0], qy[0], qz[0]);
    };
#else
    auto evaluate_list = [&](int cnt) {
      int k = 0;
#ifdef GW_DEBUG
      GWD(5, 1); GWD(6, cnt);
      { unsigned long long w = 0; for (int q = 0; q < cnt; ++q) { int pc = __popc(W.lmask[q]); w += pc; if (lane == 0) { atomicAdd(&gw_hist[pc], 1ull); } } GWD(7, w); }
#endif
      for (; k + GW_ILP <= cnt; k += GW_ILP) {          // GW_ILP independent chains per trip
#pragma unroll
        for (int u = 0; u < GW_ILP; ++u)
          grav_term(W.lxy[k + u], W.lzg[k + u], (W.lmask[k + u] >> lane) & 1u, xi, yi, zi, inv_h, h2x4, soft, gt, P.nq, P.dq, P.inv_dq, gx[u], gy[u], gz[u]);
      }
      for (; k < cnt; ++k) grav_term(W.lxy[k], W.lzg[k], (W.lmask[k] >> lane) & 1u, xi, yi, zi, inv_h, h2x4, soft, gt, P.nq, P.dq, P.inv_dq, gx[0], gy[0], gz[0]);
    };
#endif

    if (do_grav && tg.y > 0) {               // tg.y == 0: an unused tail entry of the run table
      const BvhBox gb = gbox[chunk];
      const double lo0 = gb.plo[0], lo1 = gb.plo[1], lo2 = gb.plo[2], hi0 = gb.phi[0], hi1 = gb.phi[1], hi2 = gb.phi[2];
      const double soft_min = warp_min(live ? soft : INFINITY), soft_max = warp_max(live ? soft : 0.0);
      const unsigned livemask = __ballot_sync(FULL_MASK, live);
      // FP32 screens (classification against the run's box, and the per-particle test on mixed nodes): coordinates
      // relative to the box centre, 3e-5 margins (float rounding of d^2 stays below ~1e-6 relative because
      // d^2 >= soft); whatever falls inside a margin is decided by the FP64 tests below, so no decision changes.
      const double g0x = 0.5 * (lo0 + hi0), g0y = 0.5 * (lo1 + hi1), g0z = 0.5 * (lo2 + hi2);
      const float hxf = __double2float_ru(0.5 * (hi0 - lo0)) * 1.00001f, hyf = __double2float_ru(0.5 * (hi1 - lo1)) * 1.00001f, hzf = __double2float_ru(0.5 * (hi2 - lo2)) * 1.00001f;
      const float xif = (float)(xi - g0x), yif = (float)(yi - g0y), zif = (float)(zi - g0z);
      const float softf = (float)soft, softf_min = __double2float_rd(soft_min), softf_max = __double2float_ru(soft_max);
      const float th2f = (float)theta2, inv_th2f = (float)inv_theta2;
#ifdef GW_FAR_REUSE
      // distance (float, rounded up) beyond which no particle of the run can have a near term
      const float rcf = __double2float_ru(2.0 * warp_max(live ? hc : 0.0)) * 1.0001f;
      int ncnt = 0;                          // pass 0: entries recorded in this run's near list (may exceed the capacity: overflow)
#endif
      int sn = 1, gsp = 0, ln = 0;
#ifdef GW_FAR_REUSE
      if (FR.pass && FR.near_cap > 0) {      // pass 1 from the recorded list: same entries, same masks, same order as the walk would list
        const int cnt_near = FR.near_cnt[chunk];
        if (cnt_near <= FR.near_cap) {
          for (int base = 0; base < cnt_near; base += 32) {
            const int k = base + lane;
            if (k < cnt_near) {
              const int2 en = FR.near[(size_t)chunk * FR.near_cap + k];
              const double2* p = reinterpret_cast<const double2*>(wn + en.x);
              const double2 a = __ldg(p), b = __ldg(p + 1);
              W.lxy[lane] = a; W.lzg[lane] = make_double2(b.x, P.G * b.y); W.lmask[lane] = (unsigned)en.y;
            }
            __syncwarp();
            evaluate_list(cnt_near - base < 32 ? cnt_near - base : 32);
            __syncwarp();
          }
          sn = 0;                            // nothing to walk
        }
      }
#endif
#ifdef GW_DEBUG
      int dq_len = 0, dq_ring = 0;
      auto dq_flush = [&]() { int mx = dq_len, sm = dq_len; for (int o = 16; o > 0; o >>= 1) { mx = max(mx, __shfl_xor_sync(FULL_MASK, mx, o)); sm += __shfl_xor_sync(FULL_MASK, sm, o); }
                              GWD(12, mx); GWD(13, sm); dq_len = 0; dq_ring = 0; };
      auto dq_account = [&](unsigned m) { const int pc = __popc(m); if (pc < GW_DBG_T) { dq_len += (m >> lane) & 1u; if (++dq_ring == GW_DBG_RC) dq_flush(); } else { GWD(14, 1); GWD(15, pc); } };
#endif
      if (lane == 0) W.stack[0] = make_int2(0, (int)livemask);
      __syncwarp();

      auto make_room = [&](int need) {       // keep the pushes inside the shared-memory stack
        if (sn + need <= GW_STACK) return;
        GWD(8, 1);
        __syncwarp();
        if (gsp + sn > GW_SPILL) { if (lane == 0) atomicExch(err_flag, 2); sn = 0; }     // loud: the host returns an error
        for (int e = lane; e < sn; e += 32) myspill[gsp + e] = W.stack[e];
        gsp += sn; sn = 0;
        __syncwarp();
      };

      for (;;) {
        if (sn == 0) {
          if (gsp == 0) break;
          const int take = gsp < GW_STACK / 2 ? gsp : GW_STACK / 2;
          for (int e = lane; e < take; e += 32) W.stack[e] = myspill[gsp - take + e];
          gsp -= take; sn = take;
          __syncwarp();
        }
#ifdef GW_DEBUG
        if (lane == 0) atomicMax(&gw_dbg[9], (unsigned long long)sn);
#endif
        const int npop = sn < 32 ? sn : 32;
        const bool valid = lane < npop;
        const int2 e = valid ? W.stack[sn - 1 - lane] : make_int2(0, 0);
        __syncwarp();
        sn -= npop;
        // ---- lane = node: classify against the group box
        int cls = 0;                           // 1 all accept, 2 all open, 3 mixed
        double ncx = 0.0, ncy = 0.0, ncz = 0.0, nm = 0.0, nsize = 0.0; int nchild = 0, nnch = 0;
        float rxf = 0.f, ryf = 0.f, rzf = 0.f, s2f = 0.f;
#ifdef GW_FAR_REUSE
        bool far_entry = false, near_poss = false;
#endif
        if (valid) {
          const double2* p = reinterpret_cast<const double2*>(wn + e.x);
          const double2 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
          ncx = a.x; ncy = a.y; ncz = b.x; nm = b.y; nsize = c.x;
          nchild = __double2loint(c.y); nnch = __double2hiint(c.y);
          rxf = (float)(ncx - g0x); ryf = (float)(ncy - g0y); rzf = (float)(ncz - g0z);
          const float axf = fabsf(rxf), ayf = fabsf(ryf), azf = fabsf(rzf);
          const float nx = fmaxf(axf - hxf, 0.f), ny = fmaxf(ayf - hyf, 0.f), nz = fmaxf(azf - hzf, 0.f);
          const float fx = axf + hxf, fy = ayf + hyf, fz = azf + hzf;
          const float dmin2 = fmaf(nx, nx, fmaf(ny, ny, fmaf(nz, nz, softf_min))), dmax2 = fmaf(fx, fx, fmaf(fy, fy, fmaf(fz, fz, softf_max)));
          const float szf = (float)nsize;
          s2f = szf * szf;
          if (nnch == 0 || s2f < th2f * dmin2 * (1.f - 3e-5f)) cls = 1;
          else if (s2f > th2f * dmax2 * (1.f + 3e-5f)) cls = 2;
          else cls = 3;
#ifdef GW_FAR_REUSE
          {                  // the node's particles lie inside its cell, i.e. within sqrt(3) size of its centre of mass
            const float dminf = sqrtf(fmaf(nx, nx, fmaf(ny, ny, nz * nz))) * 0.9999f;
            if (FR.pass) {
              if (dminf - 1.7321f * szf > rcf) cls = 0;          // nothing at or below this node is near any particle of the run
              else if (dminf > rcf) far_entry = true;            // its own term is far for every particle; its children may not be
            } else near_poss = !(dminf > rcf);                   // pass 0: worth recording for pass 1
          }
#endif
        }
        const unsigned emask = (unsigned)e.y;
        const unsigned balM = __ballot_sync(FULL_MASK, cls == 3);
        const int nmix = __popc(balM);
        GWD(0, 1); GWD(1, npop); GWD(4, nmix);
        if (cls == 3) {
          const int pos = __popc(balM & lt_mask);
          W.mcx[pos] = ncx; W.mcy[pos] = ncy; W.mcz[pos] = ncz; W.msize[pos] = nsize; W.mmask[pos] = emask;
          // size^2 < theta^2 d2 (1 - 1e-12)  <=>  d2 > size^2 / (theta^2 (1 - 1e-12)); the band in between gets the exact test
          const double s2t = nsize * nsize * inv_theta2;
          W.mhi[pos] = s2t * (1.0 + 2e-12); W.mlo[pos] = s2t * (1.0 - 2e-12);
          W.mf[pos] = make_float4(rxf, ryf, rzf, s2f * inv_th2f);
        }
        __syncwarp();
        // ---- lane = particle: the reference's own test on the mixed nodes; the node's lane keeps the two ballots
        unsigned acc_mask = (cls == 1) ? emask : 0u, open_mask = (cls == 2) ? emask : 0u;
        unsigned bm = balM;
#pragma unroll 2
        for (int q = 0; q < nmix; ++q) {
          const int L = __ffs(bm) - 1; bm &= bm - 1;
          const unsigned mm = W.mmask[q];
          const bool in = (mm >> lane) & 1u;
          const float4 mq = W.mf[q];
          const float dxf = xif - mq.x, dyf = yif - mq.y, dzf = zif - mq.z;
          const float d2f = fmaf(dxf, dxf, fmaf(dyf, dyf, fmaf(dzf, dzf, softf)));
          bool accept = d2f > mq.w * (1.f + 3e-5f);
          if (!accept && !(d2f < mq.w * (1.f - 3e-5f))) {       // inside the float margin: the FP64 test
            const double dx = xi - W.mcx[q], dy = yi - W.mcy[q], dz = zi - W.mcz[q];          // F:274
            const double d2 = fma(dx, dx, fma(dy, dy, fma(dz, dz, soft)));
            accept = d2 > W.mhi[q];
            if (!accept && !(d2 < W.mlo[q])) {   // borderline: redo the reference's arithmetic exactly (no contraction)       F:275-278
              const double size = W.msize[q];
              const double e2 = __dadd_rn(__dadd_rn(__dadd_rn(__dmul_rn(dx, dx), __dmul_rn(dy, dy)), __dmul_rn(dz, dz)), soft);
              accept = __ddiv_rn(size, __dsqrt_rn(e2)) < theta;
            }
          }
          const unsigned balAcc = __ballot_sync(FULL_MASK, in && accept), balOpen = __ballot_sync(FULL_MASK, in && !accept);
          if (lane == L) { acc_mask = balAcc; open_mask = balOpen; }
        }
        // ---- lane = node again: accepted (node, lane set) pairs join the interaction list, opened ones push their child block
        n_acc += __popc(acc_mask); n_open += __popc(open_mask);
#ifdef GW_FAR_REUSE
        const bool ins = acc_mask != 0u && nm > 0.0 && !far_entry;
#else
        const bool ins = acc_mask != 0u && nm > 0.0;                                // F:279: massless nodes add nothing
#endif
        const unsigned balL = __ballot_sync(FULL_MASK, ins);
#if GW_SUBLISTS > 1
        const int lpos = ln + __popc(balL & lt_mask);
        if (ins) {
          W.lxy[lpos] = make_double2(ncx, ncy); W.lzg[lpos] = make_double2(ncz, P.G * nm); W.lmask[lpos] = acc_mask;
        }
#pragma unroll
        for (int q = 0; q < GW_SUBLISTS; ++q) {
          const unsigned qm = ((1u << (32 / GW_SUBLISTS)) - 1u) << (q * (32 / GW_SUBLISTS));
          const bool hit = ins && (acc_mask & qm) != 0u;
          const unsigned bq = __ballot_sync(FULL_MASK, hit);
          if (hit) W.sub[q][sl[q] + __popc(bq & lt_mask)] = (unsigned char)lpos;
          sl[q] += __popc(bq);
        }
#else
        if (ins) {
          const int pos = ln + __popc(balL & lt_mask);
          W.lxy[pos] = make_double2(ncx, ncy); W.lzg[pos] = make_double2(ncz, P.G * nm); W.lmask[pos] = acc_mask;
        }
#endif
#ifdef GW_FAR_REUSE
        if (FR.pass == 0 && FR.near_cap > 0) {
          const bool rec = ins && near_poss;
          const unsigned balN = __ballot_sync(FULL_MASK, rec);
          if (rec) {
            const int q = ncnt + __popc(balN & lt_mask);
            if (q < FR.near_cap) FR.near[(size_t)chunk * FR.near_cap + q] = make_int2(e.x, (int)acc_mask);
          }
          ncnt += __popc(balN);
        }
#endif
#ifdef GW_DEBUG
        __syncwarp();
        for (int q = 0; q < __popc(balL); ++q) dq_account(W.lmask[ln + q]);
#endif
        ln += __popc(balL);
        const int nch = open_mask ? nnch : 0;
        int incl = nch;
        for (int o = 1; o < 32; o <<= 1) { const int t = __shfl_up_sync(FULL_MASK, incl, o); if (lane >= o) incl += t; }
        const int total = __shfl_sync(FULL_MASK, incl, 31);
        make_room(total);
        for (int k = 0; k < nch; ++k) W.stack[sn + incl - nch + k] = make_int2(nchild + k, (int)open_mask);
        sn += total;
        __syncwarp();
        if (ln > GW_LIST - 32) { evaluate_list(ln); ln = 0; __syncwarp(); }
      }
      if (ln > 0) evaluate_list(ln);
      __syncwarp();
#ifdef GW_FAR_REUSE
      if (FR.pass == 0 && FR.near_cap > 0 && lane == 0) FR.near_cnt[chunk] = ncnt;
#endif
#ifdef GW_DEBUG
      if (dq_ring) dq_flush();
#endif
    }
#pragma unroll
    for (int u = 1; u < GW_ILP; ++u) { gx[0] += gx[u]; gy[0] += gy[u]; gz[0] += gz[u]; }
    // direct sink <-> gas (unsoftened) F:567-576; per-group partial sums of the sink side (no block barrier)
    for (int s = 0; s < n_sink; ++s) {
      const double vx_ = xi - S.x[s], vy_ = yi - S.y[s], vz_ = zi - S.z[s];
      const double dr = sqrt(vx_ * vx_ + vy_ * vy_ + vz_ * vz_);
      const double d3 = dr * dr * dr;
      double wx = P.G * vx_ / d3, wy = P.G * vy_ / d3, wz = P.G * vz_ / d3;    // F:572
      const double ms = S.m[s];
      double px = 0.0, py = 0.0, pz = 0.0;
      if (live) {
        gx[0] -= ms * wx; gy[0] -= ms * wy; gz[0] -= ms * wz;                   // F:574
        const double mi = m[i];
        px = mi * wx; py = mi * wy; pz = mi * wz;                               // F:573
      }
      px = warp_sum(px); py = warp_sum(py); pz = warp_sum(pz);
      if (lane == 0) {
        double* o = sink_partial + ((size_t)(chunk - g_begin) * n_sink + s) * 3;
        o[0] = px; o[1] = py; o[2] = pz;
      }
    }
#ifdef GW_FAR_REUSE
#pragma unroll
    for (int u = 1; u < GW_ILP; ++u) { qx[0] += qx[u]; qy[0] += qy[u]; qz[0] += qz[u]; }
    if (live) {
      if (FR.pass == 0) { FR.fx[i] = gx[0]; FR.fy[i] = gy[0]; FR.fz[i] = gz[0]; FR.hcut[i] = hc; }   // far terms + sinks (the loop above)
      else { gx[0] = FR.fx[i]; gy[0] = FR.fy[i]; gz[0] = FR.fz[i]; }                                 // pass 1 is launched with n_sink = 0
      gx[0] += qx[0]; gy[0] += qy[0]; gz[0] += qz[0];
    }
#endif
    if (live) { ax[i] = gx[0]; ay[i] = gy[0]; az[i] = gz[0]; }
  }
  n_open = (unsigned long long)warp_sum_ll((long long)n_open); n_acc = (unsigned long long)warp_sum_ll((long long)n_acc);
  if (lane == 0 && do_grav) { atomicAdd(&ctr->grav_opened, n_open); atomicAdd(&ctr->grav_accepted, n_acc); }
}

// Gravity walk groups: runs of GRAV_CHUNK_WIDTH Morton-consecutive particles.  The gravity walk tolerates a run that
// straddles a coarse cell boundary (more mixed nodes, decided per particle), and full warps matter more to it than
// tight boxes.  Runs restart at every GRAV_SEG-th SPH walk group, and rank slices are cut only at those boundaries,
// so the set of runs - hence every particle's accumulation order - does not depend on the number of ranks.
#define GRAV_SEG 64
// Run formation inside one segment (SPH groups [s * GRAV_SEG, (s + 1) * GRAV_SEG), particles [pa, pb)): walk groups
// are appended to the current run; when the next group does not fit, the run is closed at that (cell) boundary if it
// already holds GRAV_MINFILL particles, else it is filled up with the head of the group.  One thread per segment.
#ifndef GRAV_MINFILL
#define GRAV_MINFILL 33          // > width: never close early (fixed runs)
#endif
template <bool WRITE>
__device__ __forceinline__ int seg_runs(int ga, int gb, int width, const int2* __restrict__ sg, int2* __restrict__ out) {
  int n_runs = 0, first = sg[ga].x, cur = 0;
  for (int g = ga; g < gb; ++g) {
    int sz = sg[g].y;
    while (sz > 0) {
      const int space = width - cur;
      if (sz <= space) { cur += sz; sz = 0; }
      else if (cur >= GRAV_MINFILL) { if (WRITE) out[n_runs] = make_int2(first, cur); ++n_runs; first += cur; cur = 0; continue; }
      else { cur += space; sz -= space; }
      if (cur == width) { if (WRITE) out[n_runs] = make_int2(first, cur); ++n_runs; first += cur; cur = 0; }
    }
  }
  if (cur > 0) { if (WRITE) out[n_runs] = make_int2(first, cur); ++n_runs; }
  return n_runs;
}
__global__ void k_seg_count(int seg0, int nseg, int n_groups, int width, const int2* __restrict__ sg, int* __restrict__ cnt) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s > nseg) return;
  if (s == nseg) { cnt[s] = 0; return; }
  const int ga = (seg0 + s) * GRAV_SEG, gb = min(ga + GRAV_SEG, n_groups);
  cnt[s] = seg_runs<false>(ga, gb, width, sg, nullptr);
}
__global__ void k_seg_chunks(int seg0, int nseg, int n_groups, int width, const int2* __restrict__ sg, const int* __restrict__ off,
                             int2* __restrict__ groups) {
  const int s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= nseg) return;
  const int ga = (seg0 + s) * GRAV_SEG, gb = min(ga + GRAV_SEG, n_groups);
  seg_runs<true>(ga, gb, width, sg, groups + off[s]);
}
// one warp per run: its position box
__global__ void k_grav_boxes(int n_chunks, const int2* __restrict__ groups, const double* __restrict__ x, const double* __restrict__ y,
                             const double* __restrict__ z, BvhBox* __restrict__ box) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_chunks) return;
  const int2 g = groups[warp];
  float plo[3] = {INFINITY, INFINITY, INFINITY}, phi[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (lane < g.y) {
    const int i = g.x + lane;
    const double p[3] = {x[i], y[i], z[i]};
    for (int k = 0; k < 3; ++k) { plo[k] = __double2float_rd(p[k]); phi[k] = __double2float_ru(p[k]); }
  }
  for (int k = 0; k < 3; ++k) { plo[k] = warp_minf(plo[k]); phi[k] = warp_maxf(phi[k]); }
  if (lane == 0) {
    BvhBox b;
    for (int k = 0; k < 3; ++k) { b.plo[k] = plo[k]; b.phi[k] = phi[k]; b.rlo[k] = 0.f; b.rhi[k] = 0.f; }
    box[warp] = b;
  }
}

// Sink-side sums of the gas terms, in an order that does not depend on the number of ranks: the per-run partials
// (one warp sum per 32-particle run, written by k_gravity) are folded serially per GRAV_SEG segment - segments and
// their runs are the same for any rank count, rank slices are cut at segment boundaries - into the global segment
// table seg[global segment][sink][3]; the ranks exchange their rows; k_sink_reduce then folds ALL segments of the
// tree with one fixed pattern on every rank.  (Per-rank totals added by an all-reduce would group the FP64 sum by
// rank: the sinks start at v = 0, so an ulp there reaches every gas acceleration within a step.)
__global__ void k_sink_seg_fold(int nseg, int seg0, int n_sink, const int* __restrict__ seg_off, const double* __restrict__ partial,
                                double* __restrict__ seg) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t >= nseg * n_sink * 3) return;
  const int s = t / (n_sink * 3), r = t % (n_sink * 3);
  double v = 0.0;
  for (int b = seg_off[s]; b < seg_off[s + 1]; ++b) v += partial[(size_t)b * n_sink * 3 + r];
  seg[(size_t)(seg0 + s) * n_sink * 3 + r] = v;
}
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
