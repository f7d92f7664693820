/*
 * sph_oracle.cpp — CPU restatement of the reference's per-step hot path.  TEST INFRASTRUCTURE.
 *
 * This file is the parity checker for the CUDA engine.  Only tests/, __graft_entry__.smoke()
 * and bench.py's cpu_baseline / --impl reference legs may build, load or call it.  The product
 * (summersph_b200/, host/) never links or imports anything under oracle/.
 *
 * PARITY UNPINNED: the reference ships no tests, golden vectors or sample inputs, and no Fortran
 * compiler exists in the build container, so this restatement could not be
 * checked against the reference's own executable.  What pins it instead: the formula-derived
 * known-answer values of SURVEY.md §4 (tests/test_oracle_kat.py), an independent O(N^2) numpy
 * brute force for the geometry-independent fixed-h definitions (tests/test_oracle_bruteforce.py),
 * an independent derivation of leaf cells from sorted descent keys, and a second, literal Python
 * reading of both Fortran programs (oracle/pyref.py: AoS records, pointer octree of particle copies,
 * recursive walks) that this file must match bit for bit over whole loop bodies
 * (tests/test_oracle_pyref.py).
 *
 * Citations: F = /root/reference/SUMMER_SPH.f90, V = "/root/reference/SUMMER_SPH - Variable.f90",
 * T = "/root/reference/SUMMER_SPH - Variable (test new)).f90".
 *
 * Arithmetic rules kept from the reference (SURVEY.md §8(a')): FP64 throughout, no FMA
 * contraction (compile with -ffp-contract=off), single-precision literals where the source has
 * them, integer powers expanded like libgcc's __powidf2, left-to-right 3-vector sums, serial
 * loops in the reference's order (OpenMP is an opt-in used for timing only).
 */
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <algorithm>
#include <limits>
#ifdef _OPENMP
#include <omp.h>
#endif

#include "../include/sph_b200.h"

namespace {

// ---- literals exactly as the compilers see them (SURVEY §8(a')) ---------------------------------
const double G_REF   = (double)39.47841760435743f;   // F:7, V:7  real(4) literal -> 39.478416442871094
const double PI_F    = 3.14159265359;                // F:125
const double PI_V    = (double)3.1415926535897932f;  // V:7       real(4) literal -> 3.1415927410125732
const double LIT_001 = (double)0.01f;                // F:373, V:405, V:528
const double LIT_015 = (double)0.15f;                // F:317, V:346
const double LIT_01  = (double)0.1f;                 // F:855
const double LIT_1EM4= (double)0.0001f;              // F:857

// x**n the way libgcc's __powidf2 does it (gfortran -O0; README.md:31 build line has no -O).
inline double powi(double x, int n) {
  double y = (n % 2) ? x : 1.0;
  while (n >>= 1) { x = x * x; if (n % 2) y *= x; }
  return y;
}

struct Particle {               // F:14-27 | V:14-29
  int    number;
  double mass, density, internal_energy, pressure, sound_speed, internal_energy_rate;
  double alpha, alpha_rate, s_length, omega;
  double position[3], velocity[3], acceleration[3];
};
struct Sink {                   // F:30-37
  double mass, radius, position[3], velocity[3], acceleration[3];
  double spin[3];                     // F:33 `spin`: declared and zeroed by the reference, never updated (SPH_FLAG_SINK_MERGE_SPIN fills it)
};
struct Node {                   // F:40-49 | V:42-52 ; particle "copies" are index ranges + snapshots
  double center[3], size;
  int    n_particles, first;    // members = order[first .. first+n)
  double mass_total, max_len, mass_center[3];
  int    child[8];              // -1: no particles in that octant
  bool   has_children;
  int    level;
};

struct Oracle {
  sph_params p;
  bool variable_h, soft_hi;
  bool sink_extras;               // SPH_FLAG_SINK_MERGE_SPIN: spin bookkeeping + the merger the reference leaves as a stub (V:1067-1073)
  int nq; double dq;
  std::vector<double> w_table, dw_table, grav_table;
  std::vector<Particle> bodies;
  std::vector<Sink> sinks;
  // tree
  std::vector<Node> nodes;
  std::vector<int> order;               // particle indices; after build = DFS (Morton) leaf order
  std::vector<int> scratch, which;
  std::vector<double> h_tree;           // snapshot of s_length at build time (node%particles(:)%s_length)
  std::vector<double> x_tree;           // snapshot of positions at build time (3N)
  std::vector<double> m_tree;
  std::vector<int> leaf_of;             // node index of each particle's leaf (or depth-limited node)
  // diagnostics
  std::vector<double> a_grav, a_gs;     // acceleration after gravity, after gravity+sinks (3N)
  sph_counts cnt;
  bool record_ngb; std::vector<std::vector<int>> ngb;
  int threads;
};

// ---- L1 kernel tables: F:55-101 | V:69-115 ------------------------------------------------------
void init_tables(Oracle& o) {
  const int nq = o.nq; const double dq = o.dq;
  o.w_table.assign(nq + 1, 0.0); o.dw_table.assign(nq + 1, 0.0); o.grav_table.assign(nq + 1, 0.0);
  for (int i = 0; i <= nq; ++i) {
    double q = i * dq;                                                    // F:64
    if (q >= 0.0 && q <= 1.0) {
      o.w_table[i]  = 1.0 - 1.5 * powi(q, 2) + 0.75 * powi(q, 3);         // F:66
      o.dw_table[i] = -3.0 * q + 2.25 * powi(q, 2);                       // F:67
      o.grav_table[i] = ((40.0 * powi(q, 3)) - (36.0 * powi(q, 5)) + (15.0 * powi(q, 6))) / 30.0;  // F:91
    } else if (q > 1.0 && q <= 2.0) {
      o.w_table[i]  = 0.25 * powi(2.0 - q, 3);                            // F:70
      o.dw_table[i] = -0.75 * powi(2.0 - q, 2);                           // F:71
      o.grav_table[i] = ((80.0 * powi(q, 3)) - (90.0 * powi(q, 4)) + (36.0 * powi(q, 5))
                         - (5.0 * powi(q, 6)) - 2.0) / 30.0;              // F:94
    } else {
      o.w_table[i] = 0.0; o.dw_table[i] = 0.0; o.grav_table[i] = 1.0;     // F:75-76,98
    }
  }
}

// F:105-127 | V:119-141
inline void lookup_kernel(const Oracle& o, double r, double hi, double& Wi, double& dWi) {
  const double dq = o.dq;
  double qi = r / hi;
  if (qi >= 0.0 && qi <= 2.0) {
    int i = std::min((int)(qi / dq), o.nq - 1);
    double alpha = (qi - i * dq) / dq;
    Wi  = (1.0 - alpha) * o.w_table[i]  + alpha * o.w_table[i + 1];
    dWi = (1.0 - alpha) * o.dw_table[i] + alpha * o.dw_table[i + 1];
  } else { Wi = 0.0; dWi = 0.0; }
  if (o.variable_h) {            // V:139-140
    Wi  = Wi  / (PI_V * powi(hi, 3));
    dWi = dWi / (PI_V * powi(hi, 4));
  } else {                       // F:125-126 normalises with the global `smoothing`
    Wi  = Wi  / (PI_F * powi(o.p.h_fixed, 3));
    dWi = dWi / (PI_F * powi(o.p.h_fixed, 4));
  }
}

// F:129-146
inline double lookup_grav_kernel(const Oracle& o, double r, double hi) {
  const double dq = o.dq;
  double qi = r / hi;
  if (qi >= 0.0 && qi <= 2.0) {
    int i = std::min((int)(qi / dq), o.nq - 1);
    double alpha = (qi - i * dq) / dq;
    return (1.0 - alpha) * o.grav_table[i] + alpha * o.grav_table[i + 1];
  }
  return 1.0;
}

// ---- L2 tree: create_tree F:795-816 | V:999-1020, build_tree F:149-246 | V:163-267 -------------
void build_tree(Oracle& o, int ni, int depth) {
  // mass, centre of mass, max_len over members in node order (ascending number: stable partition)
  {
    Node& nd = o.nodes[ni];
    double M = 0.0, mc[3] = {0.0, 0.0, 0.0}, ml = -std::numeric_limits<double>::infinity();
    for (int k = 0; k < nd.n_particles; ++k) {
      const Particle& b = o.bodies[o.order[nd.first + k]];
      M = M + b.mass;                                                      // F:169
      for (int d = 0; d < 3; ++d) mc[d] = mc[d] + b.mass * b.position[d];   // F:170
      if (b.s_length > ml) ml = b.s_length;                                // V:190-192
    }
    nd.mass_total = M; nd.max_len = ml;
    if (M > 0.0) { for (int d = 0; d < 3; ++d) nd.mass_center[d] = mc[d] / M; }      // F:173-174
    else         { for (int d = 0; d < 3; ++d) nd.mass_center[d] = nd.center[d]; }   // F:176
    nd.has_children = false;
    for (int c = 0; c < 8; ++c) nd.child[c] = -1;
    if (nd.n_particles <= 1 || depth == 0) {                               // F:182
      for (int k = 0; k < nd.n_particles; ++k) o.leaf_of[o.order[nd.first + k]] = ni;
      return;
    }
  }
  const int first = o.nodes[ni].first, np = o.nodes[ni].n_particles;
  const int level = o.nodes[ni].level;
  double ctr[3] = {o.nodes[ni].center[0], o.nodes[ni].center[1], o.nodes[ni].center[2]};
  const double size = o.nodes[ni].size;
  int counts[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int k = 0; k < np; ++k) {                                           // F:208-217 (strict >)
    const Particle& b = o.bodies[o.order[first + k]];
    int ci = 0;
    for (int j = 0; j < 3; ++j) if (b.position[j] > ctr[j]) ci |= (1 << j);
    o.which[first + k] = ci; counts[ci]++;
  }
  int start[8], wp[8]; int acc = 0;
  for (int c = 0; c < 8; ++c) { start[c] = acc; wp[c] = acc; acc += counts[c]; }
  for (int k = 0; k < np; ++k) o.scratch[first + wp[o.which[first + k]]++] = o.order[first + k];   // F:229-233
  std::copy(o.scratch.begin() + first, o.scratch.begin() + first + np, o.order.begin() + first);
  int kids[8];
  for (int c = 0; c < 8; ++c) {
    kids[c] = -1;
    if (counts[c] == 0) continue;
    Node ch;
    ch.size = size * 0.5;                                                  // F:191
    for (int j = 0; j < 3; ++j) {
      double off = ((c >> j) & 1) ? 0.25 * size : -0.25 * size;            // F:195-197
      ch.center[j] = ctr[j] + off;                                         // F:199
    }
    ch.n_particles = counts[c]; ch.first = first + start[c]; ch.level = level + 1;
    ch.mass_total = 0; ch.max_len = 0; ch.has_children = false;
    for (int d = 0; d < 3; ++d) ch.mass_center[d] = 0;
    for (int q = 0; q < 8; ++q) ch.child[q] = -1;
    kids[c] = (int)o.nodes.size();
    o.nodes.push_back(ch);
  }
  o.nodes[ni].has_children = true;
  for (int c = 0; c < 8; ++c) o.nodes[ni].child[c] = kids[c];
  for (int c = 0; c < 8; ++c) if (kids[c] >= 0) build_tree(o, kids[c], depth - 1);   // F:240-244
}

void create_tree(Oracle& o) {
  const int n = (int)o.bodies.size();
  o.nodes.clear(); o.nodes.reserve((size_t)n * 2 + 16);
  o.order.resize(n); o.scratch.resize(n); o.which.resize(n); o.leaf_of.assign(n, -1);
  o.h_tree.resize(n); o.x_tree.resize((size_t)3 * n); o.m_tree.resize(n);
  double mn[3], mx[3];
  for (int d = 0; d < 3; ++d) { mn[d] = std::numeric_limits<double>::infinity(); mx[d] = -mn[d]; }
  for (int i = 0; i < n; ++i) {
    o.order[i] = i;
    o.h_tree[i] = o.bodies[i].s_length; o.m_tree[i] = o.bodies[i].mass;
    for (int d = 0; d < 3; ++d) {
      double v = o.bodies[i].position[d];
      o.x_tree[3 * (size_t)i + d] = v;
      if (v < mn[d]) mn[d] = v;
      if (v > mx[d]) mx[d] = v;
    }
  }
  Node root;
  for (int d = 0; d < 3; ++d) root.center[d] = (mx[d] + mn[d]) / 2.0;      // F:803-805
  root.size = std::max(std::max(mx[0] - mn[0], mx[1] - mn[1]), mx[2] - mn[2]);   // F:806-808
  root.n_particles = n; root.first = 0; root.level = 0;
  root.mass_total = 0; root.max_len = 0; root.has_children = false;
  for (int d = 0; d < 3; ++d) root.mass_center[d] = 0;
  for (int q = 0; q < 8; ++q) root.child[q] = -1;
  o.nodes.push_back(root);
  build_tree(o, 0, o.p.max_depth);                                          // F:815
}

// radius used in the walk box tests: F:431 `2*smoothing` | V:471 `2*node%max_len`
inline double search_reach(const Oracle& o, const Node& nd) {
  return o.variable_h ? 2.0 * nd.max_len : 2.0 * o.p.h_fixed;
}
inline bool box_test(const double* pos, const Node& nd, double reach) {
  const double lim = reach + nd.size / 2.0;
  for (int d = 0; d < 3; ++d) if (!(std::fabs(pos[d] - nd.center[d]) < lim)) return false;
  return true;
}

// ---- density: F:416-457 | V:459-496 --------------------------------------------------------------
struct DensAcc { double rho, omega; int64_t cand, contrib; std::vector<int>* list; };
void density_tree_search(const Oracle& o, int ni, const double* pos, double h_body, DensAcc& acc) {
  const Node& nd = o.nodes[ni];
  const double reach = search_reach(o, nd);
  if (nd.n_particles > 1 && box_test(pos, nd, reach) && nd.has_children) {
    for (int c = 0; c < 8; ++c) if (nd.child[c] >= 0) density_tree_search(o, nd.child[c], pos, h_body, acc);
  } else if (nd.n_particles == 1 && box_test(pos, nd, reach)) {
    const int j = o.order[nd.first];
    double nr[3]; for (int d = 0; d < 3; ++d) nr[d] = pos[d] - o.x_tree[3 * (size_t)j + d];
    double dr = std::sqrt(((nr[0] * nr[0]) + nr[1] * nr[1]) + nr[2] * nr[2]);
    double Wj, dWj; lookup_kernel(o, dr, h_body, Wj, dWj);
    acc.rho = acc.rho + o.m_tree[j] * Wj;                                  // F:454
    if (o.variable_h) {
      double W_h = -(dr * dWj - 3.0 * Wj) / h_body;                         // V:487
      acc.omega = acc.omega + o.m_tree[j] * W_h;                           // V:493
    }
    acc.cand++; if (dr / h_body <= 2.0) acc.contrib++;
    if (acc.list) acc.list->push_back(j);
  }
}

void get_density(Oracle& o) {                                              // F:398-413 | V:440-457
  const int n = (int)o.bodies.size();
  int64_t cand = 0, contrib = 0;
  if (o.record_ngb) o.ngb.assign(n, std::vector<int>());
  #pragma omp parallel for schedule(guided) reduction(+:cand,contrib) if (o.threads > 1)
  for (int i = 0; i < n; ++i) {
    Particle& b = o.bodies[i];
    DensAcc acc{0.0, 0.0, 0, 0, o.record_ngb ? &o.ngb[i] : nullptr};
    density_tree_search(o, 0, b.position, b.s_length, acc);
    b.density = acc.rho;
    if (o.variable_h) b.omega = 1.0 + (b.s_length / (3.0 * b.density)) * acc.omega;   // V:455
    cand += acc.cand; contrib += acc.contrib;
  }
  o.cnt.density_candidates = cand; o.cnt.density_contributing = contrib;
}

void get_pressure_and_sound_speed(Oracle& o) {                             // F:459-468 | V:502-512
  const double gm1 = o.variable_h ? (o.p.gamma - 1.0) : 0.4;               // F:465 literal 0.4_dp
  const double gam = o.variable_h ? o.p.gamma : 1.4;                       // F:466 literal 1.4_dp
  for (auto& b : o.bodies) {
    b.pressure = gm1 * b.internal_energy * b.density;
    b.sound_speed = std::sqrt(gam * b.pressure / b.density);
  }
}

// ---- gravity: F:264-290 | V:285-311 | T:287-313 ----------------------------------------------------
void particle_gravforce_one(const Oracle& o, int ni, Particle& p, double theta, int64_t& opened, int64_t& accepted) {
  const Node& nd = o.nodes[ni];
  double dir[3]; for (int d = 0; d < 3; ++d) dir[d] = p.position[d] - nd.mass_center[d];
  const double soft = o.soft_hi ? 0.001 * p.s_length : 0.001 * o.p.h_fixed;  // F:275, V:296 | T:298
  double d2 = (((dir[0] * dir[0]) + dir[1] * dir[1]) + dir[2] * dir[2]) + soft;
  double dist = std::sqrt(d2);
  if ((nd.size / dist) < theta || !nd.has_children) {
    accepted++;
    if (nd.mass_total > 0.0 && dist > 0.0) {
      double W = lookup_grav_kernel(o, dist, o.variable_h ? p.s_length : o.p.h_fixed);   // F:280 | V:301
      double d3 = powi(dist, 3);
      for (int d = 0; d < 3; ++d)
        p.acceleration[d] = p.acceleration[d] - (G_REF * nd.mass_total * W * dir[d] / d3);   // F:281
    }
  } else {
    opened++;
    for (int c = 0; c < 8; ++c) if (nd.child[c] >= 0) particle_gravforce_one(o, nd.child[c], p, theta, opened, accepted);
  }
}

// ---- sinks: F:559-591 | V:691-726 ----------------------------------------------------------------
void sink_gravforces(Oracle& o) {
  for (auto& s : o.sinks) {
    for (auto& b : o.bodies) {
      double v[3]; for (int d = 0; d < 3; ++d) v[d] = b.position[d] - s.position[d];
      double dr = std::sqrt(((v[0] * v[0]) + v[1] * v[1]) + v[2] * v[2]);
      double dd = dr * dr * dr;
      for (int d = 0; d < 3; ++d) {
        double w = G_REF * v[d] / dd;                                      // F:572
        s.acceleration[d] = s.acceleration[d] + (b.mass * w);
        b.acceleration[d] = b.acceleration[d] - (s.mass * w);
      }
    }
  }
  if (o.sinks.size() < 2) return;
  for (size_t i = 0; i < o.sinks.size(); ++i)
    for (size_t j = 0; j < i; ++j) {
      Sink& si = o.sinks[i]; Sink& sj = o.sinks[j];
      double v[3]; for (int d = 0; d < 3; ++d) v[d] = sj.position[d] - si.position[d];
      double dr = std::sqrt(((v[0] * v[0]) + v[1] * v[1]) + v[2] * v[2]);
      double dd = dr * dr * dr;
      for (int d = 0; d < 3; ++d) {
        double w = G_REF * v[d] / dd;
        si.acceleration[d] = si.acceleration[d] + (sj.mass * w);
        sj.acceleration[d] = sj.acceleration[d] - (si.mass * w);
      }
    }
}

// One pair as the reference evaluates it with `body` = the higher-numbered particle that found `nb` in its walk
// (F:356-387 | V:385-421): acc_contrib (F:381-382 | V:413-414), vdotgradW (F:370 | V:401) and the two du/dt terms.
// hn = the leaf copy's s_length (V:396).  Shared by the full pair loop and by the sampled evaluation below.
inline void pair_terms(const Oracle& o, const Particle& body, const Particle& nb, double hn_tree,
                       double acc_contrib[3], double& vdotgradW, double& ub, double& un) {
    double nr[3], vij[3];
    for (int d = 0; d < 3; ++d) nr[d] = body.position[d] - nb.position[d];
    double dr = std::sqrt(((nr[0] * nr[0]) + nr[1] * nr[1]) + nr[2] * nr[2]);
    for (int d = 0; d < 3; ++d) vij[d] = body.velocity[d] - nb.velocity[d];
    double vdotr = ((vij[0] * nr[0]) + vij[1] * nr[1]) + vij[2] * nr[2];
    if (vdotr >= 0) vdotr = 0.0;                                           // F:361
    for (int d = 0; d < 3; ++d) nr[d] = nr[d] / dr;                        // F:363
    double viscous_cont, Pi_term, Pj_term;
    if (!o.variable_h) {
      const double h = o.p.h_fixed;
      double Wj, dWm; lookup_kernel(o, dr, h, Wj, dWm);                    // F:366
      double dWj[3]; for (int d = 0; d < 3; ++d) dWj[d] = nr[d] * dWm;
      vdotgradW = ((dWj[0] * vij[0]) + dWj[1] * vij[1]) + dWj[2] * vij[2];  // F:370
      double vis_nu = (h * vdotr) / (dr * dr + LIT_001 * h * h);            // F:373
      double avg_c = 0.5 * (body.sound_speed + nb.sound_speed);
      double avg_a = 0.5 * (body.alpha + nb.alpha);
      viscous_cont = (-avg_a * avg_c * vis_nu + 2 * avg_a * vis_nu * vis_nu) / (0.5 * (body.density + nb.density));   // F:378
      Pi_term = body.pressure / (body.density * body.density);
      Pj_term = nb.pressure / (nb.density * nb.density);
      for (int d = 0; d < 3; ++d) acc_contrib[d] = ((Pi_term + Pj_term) + viscous_cont) * dWj[d];   // F:381-382
    } else {
      const double hb = body.s_length, hn = hn_tree;                 // V:395-396 (leaf copy's s_length)
      double Wj, dWjm, Wi, dWim;
      lookup_kernel(o, dr, hb, Wj, dWjm);
      lookup_kernel(o, dr, hn, Wi, dWim);
      double dWj[3], dWi[3];
      for (int d = 0; d < 3; ++d) { dWj[d] = nr[d] * dWjm; dWi[d] = nr[d] * dWim; }
      double dj = ((dWj[0] * vij[0]) + dWj[1] * vij[1]) + dWj[2] * vij[2];
      double di = ((dWi[0] * vij[0]) + dWi[1] * vij[1]) + dWi[2] * vij[2];
      vdotgradW = (dj + di) / 2;                                           // V:401
      double avg_len = (hb + hn) / 2;                                      // V:402
      double vis_nu = (avg_len * vdotr) / (dr * dr + LIT_001 * avg_len * avg_len);   // V:405
      double avg_c = 0.5 * (body.sound_speed + nb.sound_speed);
      double avg_a = 0.5 * (body.alpha + nb.alpha);
      viscous_cont = (-avg_a * avg_c * vis_nu + 2 * avg_a * vis_nu * vis_nu) / (0.5 * (body.density + nb.density));   // V:410
      Pi_term = body.pressure / (body.omega * body.density * body.density);
      Pj_term = nb.pressure / (nb.omega * nb.density * nb.density);
      for (int d = 0; d < 3; ++d)
        acc_contrib[d] = ((Pi_term * dWj[d]) + (Pj_term * dWi[d])) + viscous_cont * (dWi[d] + dWj[d]) / 2;   // V:413-414
    }
    ub = nb.mass * vdotgradW * (Pi_term + 0.5 * viscous_cont);  // F:387 | V:419-421
    un = body.mass * vdotgradW * (Pj_term + 0.5 * viscous_cont);
}

// ---- SPH pair walk: F:323-395 | V:352-432 ----------------------------------------------------------
// `atomic_updates`: timing-only OpenMP variant makes the neighbour-side update atomic (the
// reference's OMP loop is racy, F:302-313); the serial parity path is the literal order.
template <bool ATOMIC>
void SPH_tree_search(Oracle& o, int ni, Particle& body, int64_t& pairs) {
  const Node& nd = o.nodes[ni];
  const double reach = search_reach(o, nd);
  if (nd.n_particles > 1 && box_test(body.position, nd, reach) && nd.has_children) {
    for (int c = 0; c < 8; ++c) if (nd.child[c] >= 0) SPH_tree_search<ATOMIC>(o, nd.child[c], body, pairs);
  } else if (nd.n_particles == 1 && box_test(body.position, nd, reach)) {
    const int num = o.order[nd.first];
    if (o.bodies[num].number >= body.number) return;                       // F:354
    Particle& nb = o.bodies[num];
    pairs++;
    double acc_contrib[3], vdotgradW, ub, un;
    pair_terms(o, body, nb, o.variable_h ? o.h_tree[num] : o.p.h_fixed, acc_contrib, vdotgradW, ub, un);
    for (int d = 0; d < 3; ++d) body.acceleration[d] = body.acceleration[d] - nb.mass * acc_contrib[d];
    body.internal_energy_rate = body.internal_energy_rate + ub;
    body.alpha_rate = body.alpha_rate + nb.mass * vdotgradW;
    if (ATOMIC) {
      for (int d = 0; d < 3; ++d) { double v = body.mass * acc_contrib[d];
        #pragma omp atomic
        nb.acceleration[d] += v; }
      #pragma omp atomic
      nb.internal_energy_rate += un;
      double v2 = body.mass * vdotgradW;
      #pragma omp atomic
      nb.alpha_rate += v2;
    } else {
      for (int d = 0; d < 3; ++d) nb.acceleration[d] = nb.acceleration[d] + body.mass * acc_contrib[d];
      nb.internal_energy_rate = nb.internal_energy_rate + un;
      nb.alpha_rate = nb.alpha_rate + body.mass * vdotgradW;
    }
  }
}

void get_SPH(Oracle& o) {                                                  // F:295-319 | V:324-348
  const int n = (int)o.bodies.size();
  int64_t pairs = 0;
  if (o.nodes[0].has_children) {
    if (o.threads > 1) {
      // timing-only variant: each thread needs a private view of body(i)'s own accumulators,
      // which other threads may be updating -> own-side updates are atomic too.
      #pragma omp parallel for schedule(guided) reduction(+:pairs)
      for (int i = 0; i < n; ++i) {
        Particle tmp = o.bodies[i];
        for (int d = 0; d < 3; ++d) tmp.acceleration[d] = 0.0;
        tmp.internal_energy_rate = 0.0; tmp.alpha_rate = 0.0;
        for (int c = 0; c < 8; ++c) if (o.nodes[0].child[c] >= 0) SPH_tree_search<true>(o, o.nodes[0].child[c], tmp, pairs);
        for (int d = 0; d < 3; ++d) {
          #pragma omp atomic
          o.bodies[i].acceleration[d] += tmp.acceleration[d];
        }
        #pragma omp atomic
        o.bodies[i].internal_energy_rate += tmp.internal_energy_rate;
        #pragma omp atomic
        o.bodies[i].alpha_rate += tmp.alpha_rate;
      }
    } else {
      for (int i = 0; i < n; ++i)
        for (int c = 0; c < 8; ++c)                                        // F:308: size(root%children)=8; empty ones fail both tests
          if (o.nodes[0].child[c] >= 0) SPH_tree_search<false>(o, o.nodes[0].child[c], o.bodies[i], pairs);
    }
  }
  o.cnt.sph_pairs = pairs;
  for (auto& b : o.bodies) {                                               // F:316-318 | V:345-347
    const double h = o.variable_h ? b.s_length : o.p.h_fixed;
    b.alpha_rate = std::max(b.alpha_rate / b.density, 0.0) + LIT_015 * ((0.1 - b.alpha) * b.sound_speed / h);
  }
}

void zero_rates(Oracle& o) {                                               // F:779-793
  for (auto& b : o.bodies) { b.acceleration[0] = b.acceleration[1] = b.acceleration[2] = 0.0; b.internal_energy_rate = 0.0; b.alpha_rate = 0.0; }
  for (auto& s : o.sinks) s.acceleration[0] = s.acceleration[1] = s.acceleration[2] = 0.0;
}

void find_forces(Oracle& o, int mask) {                                    // F:818-829
  const int n = (int)o.bodies.size();
  zero_rates(o);
  const double theta = o.p.theta_override ? o.p.theta : 0.5;               // F:825 literal
  int64_t opened = 0, accepted = 0;
  if (mask & SPH_EVAL_GRAVITY) {
    #pragma omp parallel for schedule(static) reduction(+:opened,accepted) if (o.threads > 1)
    for (int i = 0; i < n; ++i) particle_gravforce_one(o, 0, o.bodies[i], theta, opened, accepted);
  }
  o.cnt.grav_opened = opened; o.cnt.grav_accepted = accepted;
  o.a_grav.resize((size_t)3 * n);
  for (int i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) o.a_grav[3 * (size_t)i + d] = o.bodies[i].acceleration[d];
  if (mask & SPH_EVAL_SINKS) sink_gravforces(o);
  o.a_gs.resize((size_t)3 * n);
  for (int i = 0; i < n; ++i) for (int d = 0; d < 3; ++d) o.a_gs[3 * (size_t)i + d] = o.bodies[i].acceleration[d];
  if (mask & SPH_EVAL_SPH) get_SPH(o);
}

void evaluate(Oracle& o, int mask) {
  const int n = (int)o.bodies.size();
  for (int i = 0; i < n; ++i) o.bodies[i].number = i + 1;                  // F:886-888
  if (mask & SPH_EVAL_TREE) create_tree(o);
  o.cnt.n_gas = n; o.cnt.n_nodes = (int64_t)o.nodes.size();
  if (mask & SPH_EVAL_DENSITY) { get_density(o); get_pressure_and_sound_speed(o); }
  find_forces(o, mask);
}

// ---- sampled evaluation: what ONE full evaluation (F:894-898 | V:1128-1132) gives to a few target particles ------------
// For checking the engine at sizes where the full serial pair loop would take hours (16M particles): the tree is
// built over the whole set, then each target's density, gravity, sink and SPH sums are evaluated on their own.  The
// pair loop is taken in gather form: target i meets every j < i its own walk finds
// (F:351-354 | V:380-383: x_i inside Box(j)) and every j > i whose walk finds it (x_j inside Box(i), all ancestor
// tests of V:368 replayed along the root-to-leaf path of i); each term is pair_terms() with the reference's roles
// (`body` = higher number), so only the order of accumulation differs from the full loop (rounding level).
struct SampleCache { std::vector<int> idx; std::vector<double> rho, omega; };
static void sample_density(const Oracle& o, int i, double& rho, double& omega, std::vector<int>* list) {
  const Particle& b = o.bodies[i];
  DensAcc acc{0.0, 0.0, 0, 0, list};
  density_tree_search(o, 0, b.position, b.s_length, acc);
  rho = acc.rho;
  omega = o.variable_h ? 1.0 + (b.s_length / (3.0 * rho)) * acc.omega : 1.0;          // V:455
}
// all j whose position lies inside Box(i) = every node of i's root-to-leaf path passes the walk's test for x_j
static void sample_box_query(const Oracle& o, const std::vector<int>& path, int ni, const double* lo, const double* hi, std::vector<int>& out) {
  const Node& nd = o.nodes[ni];
  for (int d = 0; d < 3; ++d) if (nd.center[d] + nd.size / 2.0 < lo[d] || nd.center[d] - nd.size / 2.0 > hi[d]) return;
  if (nd.has_children) { for (int c = 0; c < 8; ++c) if (nd.child[c] >= 0) sample_box_query(o, path, nd.child[c], lo, hi, out); return; }
  for (int k = 0; k < nd.n_particles; ++k) {
    const int j = o.order[nd.first + k];
    const double* pos = &o.x_tree[3 * (size_t)j];
    bool ok = true;
    for (size_t q = 0; q < path.size() && ok; ++q) {
      const Node& pn = o.nodes[path[q]];
      ok = box_test(pos, pn, search_reach(o, pn)) && (q + 1 == path.size() ? pn.n_particles == 1 : (pn.n_particles > 1 && pn.has_children));
    }
    if (ok) out.push_back(j);
  }
}
void sample_eval(Oracle& o, int nt, const int* targets, double* rho, double* omega, double* P, double* cs,
                 double* ax, double* ay, double* az, double* udot, double* adot, int32_t* ncount, uint64_t* nhash, int64_t* npairs) {
  const double gm1 = o.variable_h ? (o.p.gamma - 1.0) : 0.4, gam = o.variable_h ? o.p.gamma : 1.4;
  const double theta = o.p.theta_override ? o.p.theta : 0.5;
  auto mix = [](uint64_t z) { z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31); };
  #pragma omp parallel for schedule(dynamic, 4) if (o.threads > 1)
  for (int t = 0; t < nt; ++t) {
    const int i = targets[t];
    auto fill = [&](Particle& p, int idx, std::vector<int>* list) {       // density + EOS of one particle (F:398-468)
      p = o.bodies[idx];
      double r, om; sample_density(o, idx, r, om, list);
      p.density = r; p.omega = om;
      p.pressure = gm1 * p.internal_energy * p.density;
      p.sound_speed = std::sqrt(gam * p.pressure / p.density);
    };
    Particle bi; std::vector<int> found;
    fill(bi, i, &found);
    { std::vector<int> v = found; std::sort(v.begin(), v.end()); uint64_t hs = 0; for (int j : v) hs += mix((uint64_t)j); ncount[t] = (int32_t)v.size(); nhash[t] = hs; }
    for (int d = 0; d < 3; ++d) bi.acceleration[d] = 0.0;
    bi.internal_energy_rate = 0.0; bi.alpha_rate = 0.0;
    // gravity (F:264-290), then sinks (F:567-576), then the pair terms: the reference's order of the three blocks
    int64_t op = 0, ac = 0;
    particle_gravforce_one(o, 0, bi, theta, op, ac);
    for (const Sink& s : o.sinks) {
      double v[3]; for (int d = 0; d < 3; ++d) v[d] = bi.position[d] - s.position[d];
      const double dr = std::sqrt(((v[0] * v[0]) + v[1] * v[1]) + v[2] * v[2]), dd = dr * dr * dr;
      for (int d = 0; d < 3; ++d) { const double w = G_REF * v[d] / dd; bi.acceleration[d] = bi.acceleration[d] - (s.mass * w); }
    }
    int64_t pairs = 0;
    for (int j : found) {                                                  // j < i: i is the `body`
      if (o.bodies[j].number >= bi.number) continue;                       // F:354
      Particle nb; fill(nb, j, nullptr);
      double acc[3], vdg, ub, un;
      pair_terms(o, bi, nb, o.variable_h ? o.h_tree[j] : o.p.h_fixed, acc, vdg, ub, un);
      for (int d = 0; d < 3; ++d) bi.acceleration[d] = bi.acceleration[d] - nb.mass * acc[d];
      bi.internal_energy_rate = bi.internal_energy_rate + ub;
      bi.alpha_rate = bi.alpha_rate + nb.mass * vdg;
      ++pairs;
    }
    const Node& leaf = o.nodes[o.leaf_of[i]];
    if (leaf.n_particles == 1) {                                           // j > i: j is the `body` whose walk reaches leaf i
      std::vector<int> path; int ni = 0;
      for (;;) {
        path.push_back(ni);
        if (ni == o.leaf_of[i]) break;
        const Node& nd = o.nodes[ni]; int ci = 0;
        for (int d = 0; d < 3; ++d) { if (o.x_tree[3 * (size_t)i + d] > nd.center[d]) ci |= (1 << d); }   // F:208-214
        ni = nd.child[ci];
      }
      const double lim = search_reach(o, leaf) + leaf.size / 2.0;
      double lo[3], hi[3]; for (int d = 0; d < 3; ++d) { lo[d] = leaf.center[d] - lim; hi[d] = leaf.center[d] + lim; }
      std::vector<int> in; sample_box_query(o, path, 0, lo, hi, in);
      for (int j : in) {
        if (o.bodies[j].number <= bi.number) continue;
        Particle bj; fill(bj, j, nullptr);
        double acc[3], vdg, ub, un;
        pair_terms(o, bj, bi, o.variable_h ? o.h_tree[i] : o.p.h_fixed, acc, vdg, ub, un);
        for (int d = 0; d < 3; ++d) bi.acceleration[d] = bi.acceleration[d] + bj.mass * acc[d];      // F:384
        bi.internal_energy_rate = bi.internal_energy_rate + un;                                   // F:388
        bi.alpha_rate = bi.alpha_rate + bj.mass * vdg;                                            // F:391
        ++pairs;
      }
    }
    const double h = o.variable_h ? bi.s_length : o.p.h_fixed;
    bi.alpha_rate = std::max(bi.alpha_rate / bi.density, 0.0) + LIT_015 * ((0.1 - bi.alpha) * bi.sound_speed / h);   // F:316-318
    rho[t] = bi.density; omega[t] = bi.omega; P[t] = bi.pressure; cs[t] = bi.sound_speed;
    ax[t] = bi.acceleration[0]; ay[t] = bi.acceleration[1]; az[t] = bi.acceleration[2];
    udot[t] = bi.internal_energy_rate; adot[t] = bi.alpha_rate; npairs[t] = pairs;
  }
}

// ---- integrator: F:742-776 -------------------------------------------------------------------------
void kick(Oracle& o, double dt) {
  for (auto& b : o.bodies) {
    for (int d = 0; d < 3; ++d) b.velocity[d] = b.velocity[d] + 0.5 * b.acceleration[d] * dt;
    b.internal_energy = b.internal_energy + 0.5 * b.internal_energy_rate * dt;
    b.alpha = b.alpha + b.alpha_rate * dt * 0.5;
  }
  for (auto& s : o.sinks) for (int d = 0; d < 3; ++d) s.velocity[d] = s.velocity[d] + 0.5 * s.acceleration[d] * dt;
}
void drift(Oracle& o, double dt) {
  for (auto& b : o.bodies) for (int d = 0; d < 3; ++d) b.position[d] = b.position[d] + b.velocity[d] * dt;
  for (auto& s : o.sinks)  for (int d = 0; d < 3; ++d) s.position[d] = s.position[d] + s.velocity[d] * dt;
}

// F:831-860 | V:1035-1065.  gfortran's MINVAL ignores NaNs unless every element is NaN.
void get_next_timestep(Oracle& o, double& dt) {
  double mn = std::numeric_limits<double>::infinity(); bool any = false, have = false;
  auto upd = [&](double v) { have = true; if (v == v) { any = true; if (v < mn) mn = v; } };
  for (auto& b : o.bodies) {
    const double h = o.variable_h ? b.s_length : o.p.h_fixed;
    double vv = ((b.velocity[0] * b.velocity[0]) + b.velocity[1] * b.velocity[1]) + b.velocity[2] * b.velocity[2];
    double aa = ((b.acceleration[0] * b.acceleration[0]) + b.acceleration[1] * b.acceleration[1]) + b.acceleration[2] * b.acceleration[2];
    upd(std::sqrt(vv / aa));
    upd(b.internal_energy / std::fabs(b.internal_energy_rate));
    upd(h / std::sqrt(vv));
    upd(h / (b.sound_speed + 1.2 * b.sound_speed));
  }
  if (have && !any) mn = std::numeric_limits<double>::quiet_NaN();
  const double scale = o.variable_h ? o.p.timestep_scale : 0.25;           // F:851 | V:1056
  double cand = mn * scale;
  if (cand > 2 * dt && 1.5 * dt < LIT_01) dt = 1.5 * dt;                   // F:855-856
  else if (cand < 0.5 * dt && dt * 0.5 > LIT_1EM4) dt = 0.5 * dt;          // F:857-858
}

// ---- V:515-546 h Newton-Raphson ---------------------------------------------------------------------
void calc_smoothing(Oracle& o) {
  const int n = (int)o.bodies.size();
  const double eta = o.p.eta, conv = o.p.convergence_criteria, max_length = o.p.max_length;
  int64_t iters = 0;
  #pragma omp parallel for schedule(guided) reduction(+:iters) if (o.threads > 1)
  for (int i = 0; i < n; ++i) {
    Particle& b = o.bodies[i];
    double old_len = b.s_length;
    b.s_length = b.s_length * (1 + ((b.mass * (powi(eta / b.s_length, 3)) / b.density) - 1) / (3 * b.omega));   // V:527
    if (b.s_length < max_length && b.s_length > LIT_001) {
      while (((b.s_length - old_len) / old_len) > conv && (b.s_length < 10.0)) {       // V:529
        old_len = b.s_length;
        DensAcc acc{0.0, 0.0, 0, 0, nullptr};
        density_tree_search(o, 0, b.position, b.s_length, acc);           // tree holds pre-update h (V:533)
        b.density = acc.rho;
        b.omega = 1.0 + (b.s_length / (3 * b.density)) * acc.omega;        // V:535
        b.s_length = b.s_length * (1 + ((b.mass * (powi(eta / b.s_length, 3))) / b.density - 1) / (3 * b.omega));   // V:538
        iters++;
      }
    } else {
      b.s_length = old_len;                                                // V:541
    }
  }
  o.cnt.h_iterations = iters;
}

// ---- V:549-597 ------------------------------------------------------------------------------------
void check_sink_creation(Oracle& o) {
  const double eta = o.p.eta;
  for (auto& b : o.bodies) {
    if (b.mass * (powi(eta / b.s_length, 3)) > 0.5) {
      for (auto& s : o.sinks) {
        double d[3]; for (int k = 0; k < 3; ++k) d[k] = s.position[k] - b.position[k];
        double dr = std::sqrt(((d[0] * d[0]) + d[1] * d[1]) + d[2] * d[2]);
        if (dr < s.radius + 2 * b.s_length) return;                        // V:563-565
      }
      Sink ns;
      for (int k = 0; k < 3; ++k) { ns.position[k] = b.position[k]; ns.velocity[k] = b.velocity[k]; ns.acceleration[k] = 0.0; ns.spin[k] = 0.0; }   // V:576-580
      ns.mass = 0.00000000001; ns.radius = 2 * b.s_length;                 // V:581-582
      o.sinks.push_back(ns);
      return;
    }
  }
}

// ---- accretion: F:484-556 | V:616-688 ---------------------------------------------------------------
void sink2gasdists(const Oracle& o, const Sink& s, int ni, std::vector<char>& keep) {
  const Node& nd = o.nodes[ni];
  double over[3]; for (int d = 0; d < 3; ++d) over[d] = nd.center[d] - s.position[d];
  auto all_lt = [&](double lim) { for (int d = 0; d < 3; ++d) if (!(std::fabs(over[d]) < lim)) return false; return true; };
  if (nd.n_particles > 1 && all_lt(s.radius + nd.size / 2.0) && nd.has_children) {    // F:529
    for (int c = 0; c < 8; ++c) if (nd.child[c] >= 0) sink2gasdists(o, s, nd.child[c], keep);
    return;
  }
  if (!o.variable_h) {
    if (nd.n_particles == 1 && all_lt(2 * s.radius + nd.size / 2.0)) {                // F:536
      double dr = 0.0;
      for (int d = 0; d < 3; ++d) dr = dr + std::sqrt(nd.center[d] * nd.center[d] - s.position[d] * s.position[d]);   // F:537
      if (dr < s.radius) keep[o.order[nd.first]] = 0;
    }
  } else {
    if (nd.n_particles == 1 && all_lt(s.radius + nd.size / 2.0)) {                    // V:668
      const int j = o.order[nd.first];
      double dr = 0.0;
      for (int d = 0; d < 3; ++d) { double q = o.x_tree[3 * (size_t)j + d] - s.position[d]; dr = dr + std::sqrt(q * q); }   // V:669
      if (dr < s.radius) keep[j] = 0;
    }
  }
}

// acc += m * (x cross v)
inline void cross_add(double* acc, double m, const double* x, const double* v) {
  acc[0] = acc[0] + m * (x[1] * v[2] - x[2] * v[1]);
  acc[1] = acc[1] + m * (x[2] * v[0] - x[0] * v[2]);
  acc[2] = acc[2] + m * (x[0] * v[1] - x[1] * v[0]);
}

template <class T, class M> void pack_vec(std::vector<T>& v, const M& keep) {
  size_t w = 0;
  for (size_t i = 0; i < v.size(); ++i) if (keep[i]) { if (w != i) v[w] = v[i]; ++w; }
  v.resize(w);
}

void initiate_sink_accretion(Oracle& o) {
  const size_t n = o.bodies.size();
  std::vector<char> any_keep(n, 1), keep(n);
  for (auto& s : o.sinks) {
    std::fill(keep.begin(), keep.end(), 1);
    sink2gasdists(o, s, 0, keep);
    double sm = 0.0, sp[3] = {0, 0, 0}, sv[3] = {0, 0, 0}, la[3] = {0, 0, 0}; size_t n_acc = 0;
    for (size_t j = 0; j < n; ++j) if (!keep[j]) {
      const Particle& b = o.bodies[j];
      sm = sm + b.mass;
      for (int d = 0; d < 3; ++d) { sp[d] = sp[d] + b.mass * b.position[d]; sv[d] = sv[d] + b.mass * b.velocity[d]; }
      cross_add(la, b.mass, b.position, b.velocity);
      any_keep[j] = 0; ++n_acc;
    }
    double l_before[3] = {la[0], la[1], la[2]};
    cross_add(l_before, s.mass, s.position, s.velocity);
    double new_mass = s.mass + sm;                                         // F:497
    for (int d = 0; d < 3; ++d) s.position[d] = (s.mass * s.position[d] + sp[d]) / new_mass;   // F:498-501
    for (int d = 0; d < 3; ++d) s.velocity[d] = (s.mass * s.velocity[d] + sv[d]) / new_mass;   // F:503-506
    s.mass = s.mass + sm;                                                  // F:508
    if (o.sink_extras && n_acc > 0) {     // what the orbit lost goes into the spin: sum of L over sink + accreted is unchanged
      double l_after[3] = {0, 0, 0};
      cross_add(l_after, s.mass, s.position, s.velocity);
      for (int d = 0; d < 3; ++d) s.spin[d] = s.spin[d] + (l_before[d] - l_after[d]);
    }
  }
  pack_vec(o.bodies, any_keep);                                            // F:546-556
}

void check_bounds(Oracle& o) {                                             // F:471-482 | V:599-614
  const double B = o.p.bounding_size;
  std::vector<char> keep(o.bodies.size());
  for (size_t i = 0; i < o.bodies.size(); ++i) {
    const double* x = o.bodies[i].position;
    keep[i] = (std::fabs(x[0]) <= B && std::fabs(x[1]) <= B && std::fabs(x[2]) <= B);
  }
  pack_vec(o.bodies, keep);
  if (o.variable_h) {
    std::vector<char> ks(o.sinks.size());
    for (size_t i = 0; i < o.sinks.size(); ++i) {
      const double* x = o.sinks[i].position;
      ks[i] = (std::fabs(x[0]) <= B && std::fabs(x[1]) <= B && std::fabs(x[2]) <= B);
    }
    pack_vec(o.sinks, ks);
  }
}

// ---- sink merger: NOT in the reference (check_sink_merger is an empty stub, V:1067-1073, its call commented
// out at V:1159).  Opt-in (SPH_FLAG_SINK_MERGE_SPIN), run where that call sits: after check_bounds.
// Two sinks with mass merge when one centre lies inside the other's accretion radius: |x_a - x_b| < max(R_a, R_b).
// The lower index survives with the summed mass, the mass-weighted position, velocity and acceleration, the larger
// radius and spin = S_a + S_b + (orbital L of the two about the origin - orbital L of the merged sink); the higher
// index is removed (order preserving).  Pairs are scanned (a, b > a) ascending and the scan restarts after a merge.
void check_sink_merger(Oracle& o) {
  bool merged = true;
  while (merged) {
    merged = false;
    for (size_t a = 0; a < o.sinks.size() && !merged; ++a)
      for (size_t b = a + 1; b < o.sinks.size() && !merged; ++b) {
        Sink& A = o.sinks[a]; Sink& B = o.sinks[b];
        if (!(A.mass > 0.0 && B.mass > 0.0)) continue;
        double d[3]; for (int k = 0; k < 3; ++k) d[k] = A.position[k] - B.position[k];
        const double dr = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (!(dr < std::max(A.radius, B.radius))) continue;
        double l_before[3] = {0, 0, 0}, l_after[3] = {0, 0, 0};
        cross_add(l_before, A.mass, A.position, A.velocity);
        cross_add(l_before, B.mass, B.position, B.velocity);
        const double M = A.mass + B.mass;
        for (int k = 0; k < 3; ++k) {
          A.position[k] = (A.mass * A.position[k] + B.mass * B.position[k]) / M;
          A.velocity[k] = (A.mass * A.velocity[k] + B.mass * B.velocity[k]) / M;
          A.acceleration[k] = (A.mass * A.acceleration[k] + B.mass * B.acceleration[k]) / M;
        }
        A.mass = M; A.radius = std::max(A.radius, B.radius);
        cross_add(l_after, A.mass, A.position, A.velocity);
        for (int k = 0; k < 3; ++k) A.spin[k] = A.spin[k] + B.spin[k] + (l_before[k] - l_after[k]);
        o.sinks.erase(o.sinks.begin() + (long)b);
        merged = true;
      }
  }
}

// one body of simulate's loop: F:886-928 | V:1120-1162
void step(Oracle& o, double& dt, double& t) {
  evaluate(o, SPH_EVAL_ALL);
  kick(o, dt);
  drift(o, dt);
  evaluate(o, SPH_EVAL_ALL);
  kick(o, dt);
  t = t + dt;
  get_next_timestep(o, dt);
  if (o.variable_h) { calc_smoothing(o); check_sink_creation(o); }
  bool any_mass = false; for (auto& s : o.sinks) if (s.mass > 0.0) any_mass = true;
  if (any_mass) initiate_sink_accretion(o);                                // F:919
  check_bounds(o);
  if (o.sink_extras) check_sink_merger(o);                                 // V:1159 (commented out in the reference)
}

// ---- conserved sums: NOT in the reference (it keeps no energy / momentum bookkeeping) ----------------
// The checker's statement of include/sph_b200.h's sph_conserved(): serial sums in ascending `number`;
// the gas-gas potential runs over the node set particle_gravforce_one accepts (F:273-279 | V:294-300).
inline double soft_potential(double q) {   // d phi/dq = g(q)/q^2, g = the polynomials of F:91,94; -1/q beyond 2
  if (q < 1.0) { const double q2 = q * q; return -1.4 + q2 * (2.0 / 3.0 + q2 * (-0.3 + 0.1 * q)); }
  if (q < 2.0) { const double q2 = q * q; return -1.6 + 1.0 / (15.0 * q) + q2 * (4.0 / 3.0 - q + 0.3 * q2 - (1.0 / 30.0) * q2 * q); }
  return -1.0 / q;
}
void potential_one(const Oracle& o, int ni, int self, const Particle& p, double theta, double& phi) {
  const Node& nd = o.nodes[ni];
  double dir[3]; for (int d = 0; d < 3; ++d) dir[d] = p.position[d] - nd.mass_center[d];
  const double soft = o.soft_hi ? 0.001 * p.s_length : 0.001 * o.p.h_fixed;
  const double d2 = (((dir[0] * dir[0]) + dir[1] * dir[1]) + dir[2] * dir[2]) + soft;
  const double dist = std::sqrt(d2);
  if ((nd.size / dist) < theta || !nd.has_children) {
    const bool own_leaf = nd.n_particles == 1 && o.order[nd.first] == self;
    if (nd.mass_total > 0.0 && dist > 0.0 && !own_leaf) {
      const double h = o.variable_h ? p.s_length : o.p.h_fixed;
      phi += nd.mass_total * (soft_potential(dist * (1.0 / h)) * (1.0 / h));
    }
  } else {
    for (int c = 0; c < 8; ++c) if (nd.child[c] >= 0) potential_one(o, nd.child[c], self, p, theta, phi);
  }
}
void conserved(Oracle& o, double* out) {
  const int n = (int)o.bodies.size();
  create_tree(o);
  const double theta = o.p.theta_override ? o.p.theta : 0.5;
  std::vector<double> phi(n, 0.0);
  #pragma omp parallel for schedule(guided) if (o.threads > 1)
  for (int i = 0; i < n; ++i) potential_one(o, 0, i, o.bodies[i], theta, phi[i]);
  double ekin = 0, eint = 0, P[3] = {0, 0, 0}, L[3] = {0, 0, 0}, mass = 0, egg = 0, es = 0;
  auto add = [&](double m, const double* x, const double* v) {
    ekin += 0.5 * m * (v[0] * v[0] + v[1] * v[1] + v[2] * v[2]);
    for (int d = 0; d < 3; ++d) P[d] += m * v[d];
    L[0] += m * (x[1] * v[2] - x[2] * v[1]); L[1] += m * (x[2] * v[0] - x[0] * v[2]); L[2] += m * (x[0] * v[1] - x[1] * v[0]);
    mass += m;
  };
  for (int i = 0; i < n; ++i) {
    const Particle& b = o.bodies[i];
    add(b.mass, b.position, b.velocity);
    eint += b.mass * b.internal_energy;
    egg += 0.5 * b.mass * (G_REF * phi[i]);
    double e = 0.0;
    for (const auto& s : o.sinks) {
      if (!(s.mass > 0.0)) continue;
      const double a = b.position[0] - s.position[0], c = b.position[1] - s.position[1], d = b.position[2] - s.position[2];
      e -= s.mass / std::sqrt(a * a + c * c + d * d);
    }
    es += G_REF * b.mass * e;
  }
  for (size_t a = 0; a < o.sinks.size(); ++a) {
    const Sink& s = o.sinks[a];
    if (!(s.mass > 0.0)) continue;
    add(s.mass, s.position, s.velocity);
    if (o.sink_extras) for (int d = 0; d < 3; ++d) L[d] += s.spin[d];
    for (size_t b = 0; b < a; ++b) {
      const Sink& t = o.sinks[b];
      if (!(t.mass > 0.0)) continue;
      const double dx = s.position[0] - t.position[0], dy = s.position[1] - t.position[1], dz = s.position[2] - t.position[2];
      es -= G_REF * s.mass * t.mass / std::sqrt(dx * dx + dy * dy + dz * dz);
    }
  }
  out[0] = ekin; out[1] = eint; out[2] = egg + es;
  for (int d = 0; d < 3; ++d) { out[3 + d] = P[d]; out[6 + d] = L[d]; }
  out[9] = mass; out[10] = egg; out[11] = es;
}

}  // namespace

// =====================================================================================================
extern "C" {

struct orc_ctx { Oracle o; };

orc_ctx* orc_create(const sph_params* p, int threads) {
  orc_ctx* c = new orc_ctx();
  Oracle& o = c->o;
  o.p = *p;
  o.variable_h = (p->mode & SPH_MODE_VARIABLE_H) != 0;
  o.soft_hi = (p->mode & SPH_FLAG_SOFT_USES_HI) != 0;
  o.sink_extras = (p->mode & SPH_FLAG_SINK_MERGE_SPIN) != 0;
  o.nq = p->nq; o.dq = 2.0 / p->nq;                                        // F:10
  o.record_ngb = false; o.threads = threads < 1 ? 1 : threads;
#ifdef _OPENMP
  omp_set_num_threads(o.threads);
#else
  o.threads = 1;
#endif
  std::memset(&o.cnt, 0, sizeof(o.cnt));
  init_tables(o);
  return c;
}
void orc_destroy(orc_ctx* c) { delete c; }
int orc_threads(orc_ctx* c) { return c->o.threads; }

void orc_tables(orc_ctx* c, double* w, double* dw, double* g) {
  Oracle& o = c->o;
  for (int i = 0; i <= o.nq; ++i) { w[i] = o.w_table[i]; dw[i] = o.dw_table[i]; g[i] = o.grav_table[i]; }
}
void orc_lookup_kernel(orc_ctx* c, double r, double h, double* W, double* dW) { lookup_kernel(c->o, r, h, *W, *dW); }
double orc_lookup_grav_kernel(orc_ctx* c, double r, double h) { return lookup_grav_kernel(c->o, r, h); }
double orc_G(void) { return G_REF; }

void orc_upload(orc_ctx* c, int64_t n, const double* x, const double* y, const double* z,
                const double* vx, const double* vy, const double* vz, const double* u, const double* m,
                const double* alpha, const double* h, int32_t ns,
                const double* sx, const double* sy, const double* sz, const double* svx, const double* svy,
                const double* svz, const double* sm, const double* srad) {
  Oracle& o = c->o;
  o.bodies.resize(n);
  for (int64_t i = 0; i < n; ++i) {
    Particle& b = o.bodies[i];
    std::memset(&b, 0, sizeof(b));
    b.number = (int)i + 1;
    b.position[0] = x[i]; b.position[1] = y[i]; b.position[2] = z[i];
    b.velocity[0] = vx[i]; b.velocity[1] = vy[i]; b.velocity[2] = vz[i];
    b.internal_energy = u[i]; b.mass = m[i];
    b.alpha = alpha ? alpha[i] : 0.0;                                      // F:681
    b.s_length = (h && o.variable_h) ? h[i] : o.p.h_fixed;   // fixed-h mode ignores column 10
    b.omega = 1.0;
  }
  if (ns > 0) {
    o.sinks.resize(ns);
    for (int i = 0; i < ns; ++i) {
      Sink& s = o.sinks[i]; std::memset(&s, 0, sizeof(s));
      s.position[0] = sx[i]; s.position[1] = sy[i]; s.position[2] = sz[i];
      s.velocity[0] = svx[i]; s.velocity[1] = svy[i]; s.velocity[2] = svz[i];
      s.mass = sm[i]; s.radius = srad ? srad[i] : o.p.sink_radius;
    }
  } else {                                                                 // F:698-707 dummy sink
    o.sinks.resize(1); std::memset(&o.sinks[0], 0, sizeof(Sink));
  }
}

void orc_record_neighbours(orc_ctx* c, int on) { c->o.record_ngb = on != 0; }
void orc_evaluate(orc_ctx* c, int mask) { evaluate(c->o, mask); }
void orc_step(orc_ctx* c, double* dt, double* t) { step(c->o, *dt, *t); }
void orc_calc_smoothing(orc_ctx* c) { calc_smoothing(c->o); }
void orc_kick(orc_ctx* c, double dt) { kick(c->o, dt); }
void orc_drift(orc_ctx* c, double dt) { drift(c->o, dt); }
void orc_next_timestep(orc_ctx* c, double* dt) { get_next_timestep(c->o, *dt); }
void orc_accrete(orc_ctx* c) {
  Oracle& o = c->o; bool any = false; for (auto& s : o.sinks) if (s.mass > 0.0) any = true;
  if (any) initiate_sink_accretion(o);
}
void orc_check_bounds(orc_ctx* c) { check_bounds(c->o); }
void orc_check_sink_creation(orc_ctx* c) { check_sink_creation(c->o); }

void orc_sizes(orc_ctx* c, int64_t* n, int32_t* ns) { *n = (int64_t)c->o.bodies.size(); *ns = (int32_t)c->o.sinks.size(); }

#define PUT(arr, expr) if (arr) { for (size_t i = 0; i < n; ++i) arr[i] = (expr); }
void orc_download(orc_ctx* c, double* x, double* y, double* z, double* vx, double* vy, double* vz,
                  double* u, double* m, double* alpha, double* h,
                  double* sx, double* sy, double* sz, double* svx, double* svy, double* svz, double* sm, double* srad) {
  Oracle& o = c->o; size_t n = o.bodies.size();
  PUT(x, o.bodies[i].position[0]) PUT(y, o.bodies[i].position[1]) PUT(z, o.bodies[i].position[2])
  PUT(vx, o.bodies[i].velocity[0]) PUT(vy, o.bodies[i].velocity[1]) PUT(vz, o.bodies[i].velocity[2])
  PUT(u, o.bodies[i].internal_energy) PUT(m, o.bodies[i].mass) PUT(alpha, o.bodies[i].alpha) PUT(h, o.bodies[i].s_length)
  n = o.sinks.size();
  PUT(sx, o.sinks[i].position[0]) PUT(sy, o.sinks[i].position[1]) PUT(sz, o.sinks[i].position[2])
  PUT(svx, o.sinks[i].velocity[0]) PUT(svy, o.sinks[i].velocity[1]) PUT(svz, o.sinks[i].velocity[2])
  PUT(sm, o.sinks[i].mass) PUT(srad, o.sinks[i].radius)
}
void orc_download_diag(orc_ctx* c, double* rho, double* omega, double* P, double* cs,
                       double* ax, double* ay, double* az, double* udot, double* adot,
                       double* sax, double* say, double* saz) {
  Oracle& o = c->o; size_t n = o.bodies.size();
  PUT(rho, o.bodies[i].density) PUT(omega, o.bodies[i].omega) PUT(P, o.bodies[i].pressure) PUT(cs, o.bodies[i].sound_speed)
  PUT(ax, o.bodies[i].acceleration[0]) PUT(ay, o.bodies[i].acceleration[1]) PUT(az, o.bodies[i].acceleration[2])
  PUT(udot, o.bodies[i].internal_energy_rate) PUT(adot, o.bodies[i].alpha_rate)
  n = o.sinks.size();
  PUT(sax, o.sinks[i].acceleration[0]) PUT(say, o.sinks[i].acceleration[1]) PUT(saz, o.sinks[i].acceleration[2])
}
/* acceleration snapshots: after tree gravity, and after gravity + sinks (3 arrays each) */
void orc_download_accel_parts(orc_ctx* c, double* gx, double* gy, double* gz, double* gsx, double* gsy, double* gsz) {
  Oracle& o = c->o; size_t n = o.bodies.size();
  PUT(gx, o.a_grav[3 * i]) PUT(gy, o.a_grav[3 * i + 1]) PUT(gz, o.a_grav[3 * i + 2])
  PUT(gsx, o.a_gs[3 * i]) PUT(gsy, o.a_gs[3 * i + 1]) PUT(gsz, o.a_gs[3 * i + 2])
}
/* order[k] = 0-based number of the k-th leaf in DFS order; per particle: leaf level, centre, size, and
 * n_in_leaf (1 normally; >1 for a depth-limited childless node). */
void orc_download_tree(orc_ctx* c, int32_t* order, int32_t* level, double* cx, double* cy, double* cz,
                       double* size, int32_t* n_in_leaf) {
  Oracle& o = c->o; size_t n = o.bodies.size();
  PUT(order, o.order[i])
  PUT(level, o.nodes[o.leaf_of[i]].level)
  PUT(cx, o.nodes[o.leaf_of[i]].center[0]) PUT(cy, o.nodes[o.leaf_of[i]].center[1]) PUT(cz, o.nodes[o.leaf_of[i]].center[2])
  PUT(size, o.nodes[o.leaf_of[i]].size) PUT(n_in_leaf, o.nodes[o.leaf_of[i]].n_particles)
}
#undef PUT
void orc_root(orc_ctx* c, double* center3, double* size) {
  Oracle& o = c->o; for (int d = 0; d < 3; ++d) center3[d] = o.nodes[0].center[d]; *size = o.nodes[0].size;
}
static inline uint64_t mix64(uint64_t z) {   /* splitmix64 finaliser; same function in the engine */
  z += 0x9E3779B97F4A7C15ull; z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull; z = (z ^ (z >> 27)) * 0x94D049BB133111EBull; return z ^ (z >> 31);
}
/* requires orc_record_neighbours(1) before the evaluation */
int64_t orc_download_neighbours(orc_ctx* c, int32_t* count, uint64_t* hash, int64_t* offsets, int32_t* list, int64_t cap) {
  Oracle& o = c->o; size_t n = o.bodies.size(); int64_t tot = 0;
  for (size_t i = 0; i < n; ++i) {
    std::vector<int> v = o.ngb[i]; std::sort(v.begin(), v.end());
    if (count) count[i] = (int32_t)v.size();
    if (hash) { uint64_t hsh = 0; for (int j : v) hsh += mix64((uint64_t)j); hash[i] = hsh; }
    if (offsets) offsets[i] = tot;
    if (list) for (size_t k = 0; k < v.size(); ++k) if (tot + (int64_t)k < cap) list[tot + k] = v[k];
    tot += (int64_t)v.size();
  }
  if (offsets) offsets[n] = tot;
  return tot;
}
/* sampled evaluation (see sample_eval): builds the tree over the whole set, then evaluates the `nt` targets
 * (0-based numbers) on their own; outputs have nt entries */
void orc_sample_eval(orc_ctx* c, int32_t nt, const int32_t* targets, double* rho, double* omega, double* P, double* cs,
                     double* ax, double* ay, double* az, double* udot, double* adot, int32_t* ncount, uint64_t* nhash, int64_t* npairs) {
  Oracle& o = c->o;
  for (size_t i = 0; i < o.bodies.size(); ++i) o.bodies[i].number = (int)i + 1;   // F:886-888
  create_tree(o);
  sample_eval(o, nt, targets, rho, omega, P, cs, ax, ay, az, udot, adot, ncount, nhash, npairs);
}
void orc_counters(orc_ctx* c, sph_counts* out) { *out = c->o.cnt; }
void orc_conserved(orc_ctx* c, double* out12) { conserved(c->o, out12); }
void orc_check_sink_merger(orc_ctx* c) { check_sink_merger(c->o); }
void orc_download_sink_spin(orc_ctx* c, double* sx, double* sy, double* sz) {
  for (size_t i = 0; i < c->o.sinks.size(); ++i) { sx[i] = c->o.sinks[i].spin[0]; sy[i] = c->o.sinks[i].spin[1]; sz[i] = c->o.sinks[i].spin[2]; }
}

}  // extern "C"
