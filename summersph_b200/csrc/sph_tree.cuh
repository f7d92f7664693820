// sph_tree.cuh — tree construction kernels.
//
// Replaces create_tree (SUMMER_SPH.f90:795-816) + build_tree (SUMMER_SPH.f90:149-246 | Variable.f90:163-267).
// The reference builds a pointer octree top-down, one particle per leaf, copying particle records at every
// level.  Here: root cube from a min/max reduction -> FP64 centre-descent keys (bit-exact replay of
// `center +- 0.25*size` with the strict `>` octant rule, F:190-214) -> radix sort -> physical re-order ->
// per-particle leaf cell (level from the sorted neighbours' common prefix, centre by re-descent) ->
//   (a) an implicit 8-ary BVH over 32-particle chunks for the neighbour walks, and
//   (b) the compressed Barnes-Hut octree in depth-first preorder with skip pointers for the gravity walk.
#pragma once
#include "sph_common.cuh"

// ------------------------------------------------------------------------------------------------------
// root cube: F:803-808 (gas particles only)
// ------------------------------------------------------------------------------------------------------
__global__ void k_bbox_partial(int n, const double* __restrict__ x, const double* __restrict__ y,
                               const double* __restrict__ z, double* __restrict__ partial) {
  double mn0 = INFINITY, mn1 = INFINITY, mn2 = INFINITY, mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) {
    double a = x[i], b = y[i], c = z[i];
    mn0 = fmin(mn0, a); mx0 = fmax(mx0, a);
    mn1 = fmin(mn1, b); mx1 = fmax(mx1, b);
    mn2 = fmin(mn2, c); mx2 = fmax(mx2, c);
  }
  mn0 = warp_min(mn0); mn1 = warp_min(mn1); mn2 = warp_min(mn2);
  mx0 = warp_max(mx0); mx1 = warp_max(mx1); mx2 = warp_max(mx2);
  __shared__ double s[6][32];
  int w = threadIdx.x >> 5, l = threadIdx.x & 31, nw = blockDim.x >> 5;
  if (l == 0) { s[0][w] = mn0; s[1][w] = mn1; s[2][w] = mn2; s[3][w] = mx0; s[4][w] = mx1; s[5][w] = mx2; }
  __syncthreads();
  if (w == 0) {
    double v[6];
    for (int k = 0; k < 6; ++k) v[k] = (l < nw) ? s[k][l] : (k < 3 ? INFINITY : -INFINITY);
    for (int k = 0; k < 3; ++k) v[k] = warp_min(v[k]);
    for (int k = 3; k < 6; ++k) v[k] = warp_max(v[k]);
    if (l == 0) for (int k = 0; k < 6; ++k) partial[blockIdx.x * 6 + k] = v[k];
  }
}

// single warp: fold the partials (optionally merged with other ranks' boxes already in `partial`)
__global__ void k_bbox_final(int nblocks, const double* __restrict__ partial, RootBox* rb) {
  int l = threadIdx.x;
  double v[6];
  for (int k = 0; k < 6; ++k) v[k] = (k < 3 ? INFINITY : -INFINITY);
  for (int b = l; b < nblocks; b += 32)
    for (int k = 0; k < 6; ++k) v[k] = (k < 3) ? fmin(v[k], partial[b * 6 + k]) : fmax(v[k], partial[b * 6 + k]);
  for (int k = 0; k < 3; ++k) v[k] = warp_min(v[k]);
  for (int k = 3; k < 6; ++k) v[k] = warp_max(v[k]);
  if (l == 0) {
    for (int k = 0; k < 3; ++k) { rb->mn[k] = v[k]; rb->mx[k] = v[3 + k]; }
    rb->cx = __ddiv_rn(__dadd_rn(v[3], v[0]), 2.0);                     // F:803
    rb->cy = __ddiv_rn(__dadd_rn(v[4], v[1]), 2.0);
    rb->cz = __ddiv_rn(__dadd_rn(v[5], v[2]), 2.0);
    double ex = __dsub_rn(v[3], v[0]), ey = __dsub_rn(v[4], v[1]), ez = __dsub_rn(v[5], v[2]);
    rb->size = fmax(fmax(ex, ey), ez);                                  // F:806-808
  }
}

// domain decomposition: the ranks' boxes merge with ONE min all-reduce over [mn, -mx]
__global__ void k_dd_box_pack(const RootBox* __restrict__ rb, double* __restrict__ v) {
  if (threadIdx.x < 3) { v[threadIdx.x] = rb->mn[threadIdx.x]; v[3 + threadIdx.x] = -rb->mx[threadIdx.x]; }
}
__global__ void k_dd_box_unpack(double* __restrict__ v) { if (threadIdx.x < 3) v[3 + threadIdx.x] = -v[3 + threadIdx.x]; }

// ------------------------------------------------------------------------------------------------------
// descent keys: replay of F:190-214 for lmax levels. Digit = [x>cx] + 2[y>cy] + 4[z>cz] (strict >).
// ------------------------------------------------------------------------------------------------------
__global__ void k_keys(int n, const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                       const RootBox* __restrict__ rb, int lmax, uint64_t* __restrict__ key, uint64_t* __restrict__ key_lo,
                       int* __restrict__ idx) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  double px = x[i], py = y[i], pz = z[i];
  double cx = rb->cx, cy = rb->cy, cz = rb->cz, s = rb->size;
  uint64_t k = 0, k2 = 0;
  for (int l = 0; l < lmax; ++l) {
    int bx = px > cx, by = py > cy, bz = pz > cz;
    const uint64_t dg = (uint64_t)(bx | (by << 1) | (bz << 2));
    if (l < SPH_KEY_LEVELS) k = (k << 3) | dg; else k2 = (k2 << 3) | dg;
    double q = 0.25 * s;                                                 // F:195 (exact scaling)
    cx = __dadd_rn(cx, bx ? q : -q);                                     // F:199
    cy = __dadd_rn(cy, by ? q : -q);
    cz = __dadd_rn(cz, bz ? q : -q);
    s = s * 0.5;                                                         // F:191
  }
  const int l1 = lmax < SPH_KEY_LEVELS ? lmax : SPH_KEY_LEVELS;
  key[i] = k << (3 * (SPH_KEY_LEVELS - l1));
  if (key_lo) key_lo[i] = k2 << (3 * (SPH_KEY_LEVELS2 - (lmax > SPH_KEY_LEVELS ? lmax : SPH_KEY_LEVELS)));
  idx[i] = i;
}

__global__ void k_gather_u64(int n, const int* __restrict__ perm, const uint64_t* __restrict__ src, uint64_t* __restrict__ dst) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) dst[i] = src[perm[i]];
}

// physical re-order of the state into sorted order
struct PermuteArgs { const double* src[10]; double* dst[10]; const int* id_src; int* id_dst; };
// Four particles per thread, field-major: the four gathers of one field are in flight together and a block touches
// two streams (one source, one destination array) at a time instead of twenty-two.
__global__ void k_permute(int n, const int* __restrict__ perm, PermuteArgs a) {
  const int base = blockIdx.x * (blockDim.x * 4) + threadIdx.x;
  int p[4]; bool ok[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) { const int i = base + k * blockDim.x; ok[k] = i < n; p[k] = ok[k] ? perm[i] : 0; }
#pragma unroll
  for (int f = 0; f < 10; ++f) {
    if (a.src[f] == nullptr) continue;          // a field that is permuted by a later launch (sph_step_host: still on its way from the host)
    double v[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) v[k] = a.src[f][p[k]];
#pragma unroll
    for (int k = 0; k < 4; ++k) if (ok[k]) a.dst[f][base + k * blockDim.x] = v[k];
  }
  if (a.id_src == nullptr) return;
  int w[4];
#pragma unroll
  for (int k = 0; k < 4; ++k) w[k] = a.id_src[p[k]];
#pragma unroll
  for (int k = 0; k < 4; ++k) if (ok[k]) a.id_dst[base + k * blockDim.x] = w[k];
}

// ------------------------------------------------------------------------------------------------------
// leaf cells. level_i = 1 + max(lcp(i-1,i), lcp(i,i+1)) = the first cell in which i is alone (F:182);
// centre by replaying the descent along the key digits; reach R = 2h + size/2 (F:443 | V:479 right-hand side).
// A particle that shares its full key with a neighbour sits in a depth-limited multi-particle childless
// node, which the density / SPH / accretion walks skip (neither branch of F:431,443 fires): R = -1.
// ------------------------------------------------------------------------------------------------------
__global__ void k_leaf(int n, const uint64_t* __restrict__ key, const uint64_t* __restrict__ key_lo, const double* __restrict__ h,
                       const RootBox* __restrict__ rb, DevParams P, int* __restrict__ level, double* __restrict__ lcx,
                       double* __restrict__ lcy, double* __restrict__ lcz, double* __restrict__ reach, int* __restrict__ err_flag,
                       int has_prev = 0, uint64_t key_prev = 0, int has_next = 0, uint64_t key_next = 0) {
  // has_prev / has_next (domain decomposition, single-word keys): the last key of the preceding domain / the first key
  // of the following one stand in for the neighbours a rank does not hold, so the leaf cells equal the global tree's
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = key[i], k2 = key_lo ? key_lo[i] : 0;
  int d = (i > 0) ? lcp_levels2(key[i - 1], key_lo ? key_lo[i - 1] : 0, k, k2, P.lmax) : (has_prev ? lcp_levels2(key_prev, 0, k, k2, P.lmax) : -1);
  int e = (i < n - 1) ? lcp_levels2(k, k2, key[i + 1], key_lo ? key_lo[i + 1] : 0, P.lmax) : (has_next ? lcp_levels2(k, k2, key_next, 0, P.lmax) : -1);
  int m = d > e ? d : e;
  bool multi = (m >= P.lmax);
  int lev = multi ? P.lmax : m + 1;
  if (multi && P.depth_unbounded) atomicExch(err_flag, 1);
  double cx = rb->cx, cy = rb->cy, cz = rb->cz, s = rb->size;
  for (int l = 0; l < lev; ++l) {
    const int dg = key_digit(k, k2, l);
    double q = 0.25 * s;
    cx = __dadd_rn(cx, (dg & 1) ? q : -q);
    cy = __dadd_rn(cy, (dg & 2) ? q : -q);
    cz = __dadd_rn(cz, (dg & 4) ? q : -q);
    s = s * 0.5;
  }
  level[i] = lev;
  lcx[i] = cx; lcy[i] = cy; lcz[i] = cz;
  double hh = P.variable_h ? h[i] : P.h_fixed;
  reach[i] = multi ? -1.0 : __dadd_rn(2.0 * hh, s / 2.0);                // V:479 `2*max_len + size/2`
}

// The positions did not move since the last build (evaluation A of a step follows evaluation B of the previous
// one, F:894 after F:905, unless particles were removed): keys, order, leaf cells, octree, masses and walk
// groups are what a rebuild would produce again; only h changed (calc_smoothing, V:1152), so only the reach
// R = 2h + size/2 of every leaf (and the BVH boxes built from it) is refreshed.
__global__ void k_refresh_reach(int n, const double* __restrict__ h, const int* __restrict__ level, const RootBox* __restrict__ rb,
                                DevParams P, double* __restrict__ reach) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  if (!(reach[i] > 0.0)) return;                                         // depth-limited multi-particle node: stays -1
  double s = rb->size; const int lev = level[i];
  for (int l = 0; l < lev; ++l) s = s * 0.5;
  const double hh = P.variable_h ? h[i] : P.h_fixed;
  reach[i] = __dadd_rn(2.0 * hh, s / 2.0);                               // V:479 `2*max_len + size/2`
}

// ------------------------------------------------------------------------------------------------------
// implicit 8-ary BVH over walk groups (neighbour walks only need a conservative superset; the exact
// per-particle leaf-box test decides membership).
// ------------------------------------------------------------------------------------------------------
// Walk groups: octree-cell-aligned runs of <= SPH_CHUNK Morton-consecutive particles.  A node with
// count <= SPH_CHUNK whose parent holds more is a bucket; consecutive sibling buckets are packed greedily
// up to SPH_CHUNK.  Every group therefore lies inside one octree cell that holds > SPH_CHUNK particles
// only through its siblings, which bounds its box (fixed 32-particle chunks that straddle a coarse cell
// boundary have boxes spanning a large part of the domain and swallow millions of candidates).
// gsize[i] = size of the group starting at sorted particle i, else 0.
__global__ void k_group_mark(int n_nodes, const GNode* __restrict__ nodes, const int* __restrict__ node_part,
                             const int* __restrict__ node_count, int* __restrict__ gsize) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  const int cnt = node_count[v];
  if (v == 0 && cnt <= SPH_CHUNK) { gsize[0] = cnt; return; }
  if (node_part[v] >= 0 || cnt <= SPH_CHUNK) return;
  const int end = nodes[v].next;
  int c = v + 1, run_first = -1, run_n = 0;
  while (c < end) {
    const int cc = node_count[c];
    const int pc = node_part[c];
    const int first = pc >= 0 ? pc : -1 - pc;
    if (cc <= SPH_CHUNK) {
      if (run_n > 0 && run_n + cc > SPH_CHUNK) { gsize[run_first] = run_n; run_n = 0; }
      if (run_n == 0) run_first = first;
      run_n += cc;
    } else if (run_n > 0) { gsize[run_first] = run_n; run_n = 0; }
    c = nodes[c].next;
  }
  if (run_n > 0) gsize[run_first] = run_n;
}

// compacted (first, size) per group -> packed int2
__global__ void k_group_pack(int n_groups, const int* __restrict__ gfirst, const int* __restrict__ gsize, int2* __restrict__ groups) {
  int g = blockIdx.x * blockDim.x + threadIdx.x;
  if (g >= n_groups) return;
  const int f = gfirst[g];
  groups[g] = make_int2(f, gsize[f]);
}

__global__ void k_bvh_leaf(int n_groups, const int2* __restrict__ groups,
                           const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                           const double* __restrict__ lcx, const double* __restrict__ lcy, const double* __restrict__ lcz,
                           const double* __restrict__ reach, BvhBox* __restrict__ box) {
  int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (warp >= n_groups) return;
  const int2 g = groups[warp];
  int i = g.x + lane;
  float plo[3] = {INFINITY, INFINITY, INFINITY}, phi[3] = {-INFINITY, -INFINITY, -INFINITY};
  float rlo[3] = {INFINITY, INFINITY, INFINITY}, rhi[3] = {-INFINITY, -INFINITY, -INFINITY};
  if (lane < g.y) {
    double p[3] = {x[i], y[i], z[i]}, c[3] = {lcx[i], lcy[i], lcz[i]}, R = reach[i];
    for (int k = 0; k < 3; ++k) {
      plo[k] = __double2float_rd(p[k]); phi[k] = __double2float_ru(p[k]);
      if (R > 0.0) { rlo[k] = __double2float_rd(c[k] - R); rhi[k] = __double2float_ru(c[k] + R); }
    }
  }
  for (int k = 0; k < 3; ++k) {
    plo[k] = warp_minf(plo[k]); phi[k] = warp_maxf(phi[k]);
    rlo[k] = warp_minf(rlo[k]); rhi[k] = warp_maxf(rhi[k]);
  }
  if (lane == 0) {
    BvhBox b;
    for (int k = 0; k < 3; ++k) { b.plo[k] = plo[k]; b.phi[k] = phi[k]; b.rlo[k] = rlo[k]; b.rhi[k] = rhi[k]; }
    box[warp] = b;
  }
}

__global__ void k_bvh_up(int n_child, const BvhBox* __restrict__ child, BvhBox* __restrict__ parent) {
  int p = blockIdx.x * blockDim.x + threadIdx.x;
  int n_parent = (n_child + SPH_BVH_FAN - 1) / SPH_BVH_FAN;
  if (p >= n_parent) return;
  BvhBox b;
  for (int k = 0; k < 3; ++k) { b.plo[k] = INFINITY; b.phi[k] = -INFINITY; b.rlo[k] = INFINITY; b.rhi[k] = -INFINITY; }
  for (int c = p * SPH_BVH_FAN; c < n_child && c < (p + 1) * SPH_BVH_FAN; ++c) {
    BvhBox q = child[c];
    for (int k = 0; k < 3; ++k) {
      b.plo[k] = fminf(b.plo[k], q.plo[k]); b.phi[k] = fmaxf(b.phi[k], q.phi[k]);
      b.rlo[k] = fminf(b.rlo[k], q.rlo[k]); b.rhi[k] = fmaxf(b.rhi[k], q.rhi[k]);
    }
  }
  parent[p] = b;
}

// ------------------------------------------------------------------------------------------------------
// compressed Barnes-Hut octree in DFS preorder.
// A branching node at level L starts at sorted index i iff i is the first particle of its L-cell
// (lcp(i-1,i) < L) and the cell's first and last particle differ in digit L+1 (lcp(first,last) == L).
// Single-child chains of the reference collapse (same M, COM; MAC decided by the smallest cell: see
// SURVEY.md Appendix B).  cnt[i] = number of branching nodes starting at i; preorder slot of the k-th
// (ascending level) = i + off[i] + k; leaf i sits at i + off[i] + cnt[i].
// ------------------------------------------------------------------------------------------------------
__device__ __forceinline__ bool key_le(const uint64_t* __restrict__ key, const uint64_t* __restrict__ key_lo, int r,
                                       uint64_t kmax, uint64_t kmax_lo) {
  const uint64_t a = key[r];
  if (a != kmax) return a < kmax;
  return key_lo ? key_lo[r] <= kmax_lo : true;
}
__device__ __forceinline__ int cell_last(const uint64_t* __restrict__ key, const uint64_t* __restrict__ key_lo, int n, int lo,
                                         uint64_t kmax, uint64_t kmax_lo) {
  // largest r >= lo with key[r] <= kmax (two-word compare), given key[lo] <= kmax; gallop then bisect
  int step = 1, hi = lo;
  while (true) {
    int probe = lo + step;
    if (probe >= n) { hi = n - 1; break; }
    if (key_le(key, key_lo, probe, kmax, kmax_lo)) { lo = probe; step <<= 1; } else { hi = probe - 1; break; }
  }
  while (lo < hi) {
    int mid = (lo + hi + 1) >> 1;
    if (key_le(key, key_lo, mid, kmax, kmax_lo)) lo = mid; else hi = mid - 1;
  }
  return lo;
}

template <bool EMIT>
__global__ void k_oct_nodes(int n, const uint64_t* __restrict__ key, const uint64_t* __restrict__ key_lo, int lmax,
                            const RootBox* __restrict__ rb, int* __restrict__ cnt, const int* __restrict__ off, int n_nodes,
                            GNode* __restrict__ nodes, int* __restrict__ node_part, int* __restrict__ node_count) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint64_t k = key[i], k2 = key_lo ? key_lo[i] : 0;
  int d = (i > 0) ? lcp_levels2(key[i - 1], key_lo ? key_lo[i - 1] : 0, k, k2, lmax) : -1;
  int e = (i < n - 1) ? lcp_levels2(k, k2, key[i + 1], key_lo ? key_lo[i + 1] : 0, lmax) : -1;
  int count = 0;
  int total = EMIT ? cnt[i] : 0;
  int base = EMIT ? i + off[i] : 0;
  int r = i;
  double root_size = EMIT ? rb->size : 0.0;
  for (int L = e; L > d; --L) {            // smallest cell first; cells nest so r only grows
    uint64_t kmax, kmax_lo;
    if (L <= SPH_KEY_LEVELS) {
      const int shift = 3 * (SPH_KEY_LEVELS - L);
      kmax = (shift >= 63) ? ~0ull : (k | ((1ull << shift) - 1ull)); kmax_lo = ~0ull;
    } else {
      const int shift = 3 * (SPH_KEY_LEVELS2 - L);
      kmax = k; kmax_lo = k2 | ((1ull << shift) - 1ull);
    }
    r = cell_last(key, key_lo, n, r, kmax, kmax_lo);
    if (lcp_levels2(k, k2, key[r], key_lo ? key_lo[r] : 0, lmax) == L) {
      if (EMIT) {
        int slot = base + (total - 1 - count);
        double s = root_size;
        for (int q = 0; q < L; ++q) s = s * 0.5;
        GNode g; g.cx = g.cy = g.cz = 0.0; g.m = 0.0; g.size = s;
        g.next = (r + 1 < n) ? (r + 1) + off[r + 1] : n_nodes;
        g.flags = (L >= lmax) ? 1 : 0;     // depth-limited multi-particle node is childless (F:182)
        nodes[slot] = g;
        node_part[slot] = -1 - i;           // internal: encodes first particle
        node_count[slot] = r - i + 1;
      }
      ++count;
    }
  }
  if (!EMIT) { cnt[i] = count; return; }
  int slot = base + total;
  GNode g; g.cx = g.cy = g.cz = 0.0; g.m = 0.0; g.size = 0.0; g.next = slot + 1; g.flags = 1;
  nodes[slot] = g;
  node_part[slot] = i;
  node_count[slot] = 1;
}

// parent / child-count links: one thread per preorder slot; internal nodes walk their child chain.
// wcount = number of children the gravity walk may descend into (0 for leaves and for the depth-limited
// childless multi-particle nodes, F:182): sizes the child blocks of the walk layout below.
__global__ void k_oct_link(int n_nodes, const GNode* __restrict__ nodes, const int* __restrict__ node_part,
                           int* __restrict__ parent, int* __restrict__ nchild, int* __restrict__ wcount) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  if (v == 0) parent[0] = -1;
  if (node_part[v] >= 0) { nchild[v] = 0; wcount[v] = 0; return; }
  int end = nodes[v].next, c = v + 1, k = 0;
  while (c < end) { parent[c] = v; ++k; c = nodes[c].next; }
  nchild[v] = k;
  wcount[v] = (nodes[v].flags & 1) ? 0 : k;
}

// Walk layout of the octree: the children of a node are contiguous (block start = 1 + exclusive scan of
// wcount in preorder), so one warp can classify up to 32 nodes per trip with lane-parallel loads and push
// whole child blocks.  widx[v] = slot of preorder node v (-1: below a childless node, never visited).
__global__ void k_oct_widx(int n_nodes, const GNode* __restrict__ nodes, const int* __restrict__ wcount,
                           const int* __restrict__ wstart, int* __restrict__ widx) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  if (v == 0) widx[0] = 0;
  if (wcount[v] == 0) return;
  const int base = 1 + wstart[v];
  int end = nodes[v].next, c = v + 1, k = 0;
  while (c < end) { widx[c] = base + k; ++k; c = nodes[c].next; }
}

// bottom-up mass / first-moment sums: one thread per particle leaf, last arriver folds the parent
// (children summed in DFS order => deterministic). cx,cy,cz temporarily hold sum(m*x).
__global__ void k_oct_up(int n, const int* __restrict__ off, const int* __restrict__ cnt,
                         const double* __restrict__ x, const double* __restrict__ y, const double* __restrict__ z,
                         const double* __restrict__ m, const double* __restrict__ h, const int* __restrict__ level,
                         const RootBox* __restrict__ rb, GNode* nodes, const int* __restrict__ parent,
                         const int* __restrict__ nchild, int* arrive) {
  int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  int v = i + off[i] + cnt[i];
  double mi = m[i];
  double s = rb->size; int lev = level[i];
  for (int q = 0; q < lev; ++q) s = s * 0.5;
  nodes[v].cx = __dmul_rn(mi, x[i]); nodes[v].cy = __dmul_rn(mi, y[i]); nodes[v].cz = __dmul_rn(mi, z[i]);   // F:170
  nodes[v].m = mi; nodes[v].size = s;
  int p = parent[v];
  while (p >= 0) {
    __threadfence();
    int old = atomicAdd(&arrive[p], 1);
    if (old != nchild[p] - 1) return;
    __threadfence();
    double M = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;
    int end = nodes[p].next, c = p + 1;
    while (c < end) {
      const double2 a0 = __ldcg(reinterpret_cast<const double2*>(&nodes[c]));
      const double2 a1 = __ldcg(reinterpret_cast<const double2*>(&nodes[c]) + 1);
      M += a1.y; sx += a0.x; sy += a0.y; sz += a1.x;
      c = __ldcg(&nodes[c].next);
    }
    nodes[p].cx = sx; nodes[p].cy = sy; nodes[p].cz = sz; nodes[p].m = M;
    p = parent[p];
  }
}

// wbase: first walk-layout slot of this tree (0; DD_TOP_CAP under the domain decomposition, where the top tree comes first)
__global__ void k_oct_finalize(int n_nodes, GNode* nodes, const int* __restrict__ wcount, const int* __restrict__ wstart,
                               const int* __restrict__ widx, WNode* __restrict__ wnodes, int wbase = 0) {
  int v = blockIdx.x * blockDim.x + threadIdx.x;
  if (v >= n_nodes) return;
  GNode g = nodes[v];
  if (g.m > 0.0) {                                                         // F:173-177
    g.cx = __ddiv_rn(g.cx, g.m); g.cy = __ddiv_rn(g.cy, g.m); g.cz = __ddiv_rn(g.cz, g.m);
    nodes[v].cx = g.cx; nodes[v].cy = g.cy; nodes[v].cz = g.cz;
  }
  const int w = widx[v];
  if (w >= 0) {
    WNode o; o.cx = g.cx; o.cy = g.cy; o.cz = g.cz; o.m = g.m; o.size = g.size;
    o.child = wbase + 1 + wstart[v]; o.nchild = wcount[v];
    wnodes[wbase + w] = o;
  }
}
