// sph_domain_host.inl — host orchestration of the Morton-domain decomposition (included by sph_engine.cu inside its
// anonymous namespace).  See sph_domain.cuh for the scheme; DESIGN.md §4 for the protocol and what stays identical
// to the single-rank run (keys, order, leaf cells, neighbour sets, interaction counts: bit-exact; FP64 sums: the same
// terms, grouped differently only where a walk group or gravity run meets a domain boundary).

// ---- exported arrays (peer-mapped on every rank) ------------------------------------------------------------------------
enum DDSlot { DS_ST = 0, DS_ID = 20, DS_KEY = 22, DS_PERM = 24, DS_LCX = 26, DS_LCY, DS_LCZ, DS_REACH, DS_RHO, DS_CS, DS_POR2, DS_GROUPS, DS_BVH, DS_WN,
              DS_ACCKEY, DS_ACCREC, DS_LEVEL, DS_OMEGA, DS_PRS, DS_AX, DS_AY, DS_AZ, DS_UDOT, DS_ADOT, DS_STAGE1, DS_STAGE2, DS_COUNT };

std::vector<void*> dd_exported(sph_ctx* c) {
  std::vector<void*> a(DS_COUNT, nullptr);
  for (int b = 0; b < 2; ++b) { for (int f = 0; f < 10; ++f) a[DS_ST + b * 10 + f] = c->st[b][f]; a[DS_ID + b] = c->id[b]; a[DS_KEY + b] = c->key_alloc[b]; a[DS_PERM + b] = c->perm_alloc[b]; }
  a[DS_LCX] = c->lcx; a[DS_LCY] = c->lcy; a[DS_LCZ] = c->lcz; a[DS_REACH] = c->reach; a[DS_RHO] = c->rho; a[DS_CS] = c->cs; a[DS_POR2] = c->por2;
  a[DS_GROUPS] = c->groups; a[DS_BVH] = c->bvh; a[DS_WN] = c->wnodes; a[DS_ACCKEY] = c->dd_acc_key; a[DS_ACCREC] = c->dd_acc_rec; a[DS_LEVEL] = c->level;
  a[DS_STAGE1] = c->stage_d; a[DS_STAGE2] = c->stage_d2;
  a[DS_OMEGA] = c->omega; a[DS_PRS] = c->prs; a[DS_AX] = c->ax; a[DS_AY] = c->ay; a[DS_AZ] = c->az; a[DS_UDOT] = c->udot; a[DS_ADOT] = c->adot;
  return a;
}
template <class T> T* dd_peer(sph_ctx* c, int slot, int r) { return reinterpret_cast<T*>(c->peer[slot][r]); }


// all-gather of a small host struct (through the device staging buffer for NCCL, directly for the host communicator)
int dd_allgather_host(sph_ctx* c, const void* mine, size_t bytes, void* all) {
  if (c->hc) {
    if (!c->hc->allgather(mine, bytes, all)) { c->err = "host communicator: all-gather failed"; return SPH_ERR_COMM; }
    return SPH_OK;
  }
  const int R = c->n_ranks;
  if (c->blob_cap < bytes * R) { if (c->d_blob) cudaFree(c->d_blob); CK(cudaMalloc(&c->d_blob, bytes * R)); c->blob_cap = bytes * R; }
  CK(cudaMemcpyAsync((char*)c->d_blob + bytes * c->rank, mine, bytes, cudaMemcpyHostToDevice, c->stream));
  { int r_ = coll_allgather(c, c->stream, (char*)c->d_blob + bytes * c->rank, c->d_blob, bytes); if (r_) return r_; }
  CK(cudaMemcpyAsync(all, c->d_blob, bytes * R, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return SPH_OK;
}
// cross-rank barrier in stream order: every rank's earlier work on its stream is complete before any rank's later work starts
int dd_barrier(sph_ctx* c) {
  if (!c->d_flag) DA(c->d_flag, 4);
  return coll_allreduce(c, c->stream, c->d_flag + 3, 1, NC_INT32, NC_SUM);
}

// Collective error agreement: a capacity failure on one rank must stop every rank at the same point (a rank that
// returned alone would leave its peers waiting in the next collective).  Returns SPH_OK only when no rank reported `code`.
int dd_agree(sph_ctx* c, int code, const char* what) {
  int mine = code, all[DD_MAX_RANKS] = {0};
  { int r_ = dd_allgather_host(c, &mine, sizeof(int), all); if (r_) return r_; }
  int worst = 0, who = -1;
  for (int q = 0; q < c->n_ranks; ++q) if (all[q] && !worst) { worst = all[q]; who = q; }
  if (!worst) return SPH_OK;
  if (!code) c->err = std::string("domain decomposition: rank ") + std::to_string(who) + " stopped in " + what + " (see its error); every rank returns";
  return worst;
}

int dd_alloc(sph_ctx* c) {               // private scratch of the decomposition (sizes that do not depend on the particle count)
  if (c->dd_samples) return SPH_OK;
  const int R = c->n_ranks;
  DA(c->dd_samples, (size_t)R * DD_SAMPLES); DA(c->dd_counts, R); DA(c->dd_split, R + 1); DA(c->dd_sendoff, R + 1); DA(c->dd_segkeys, 2 * DD_MAX_RANKS);
  DA(c->dd_cells, DD_MAX_CELLS); DA(c->dd_contrib, (size_t)DD_MAX_CELLS * 8);
  DA(c->dd_let_ctl, 64); DA(c->dd_create8, 8); DA(c->dd_cand, 2);
  if (!c->d_flag) DA(c->d_flag, 4);
  CK(cudaMemset(c->d_flag, 0, 4 * sizeof(int)));
  CK(cudaFuncSetAttribute(k_dd_splitters, cudaFuncAttributeMaxDynamicSharedMemorySize, DD_MAX_RANKS * DD_SAMPLES * 8));
  return SPH_OK;
}

// ---- top tree assembly (host, identical on every rank) -------------------------------------------------------------------
struct DDTopNode { int cell; int nchild; int kids[8]; bool kid_is_top[8]; int kid_owner[8]; double m, sx, sy, sz; int block; };

// key prefix of `key` at `level` digits (left-aligned like the keys themselves)
inline unsigned long long dd_prefix(unsigned long long key, int level) {
  if (level <= 0) return 0ull;
  const int shift = 3 * (SPH_KEY_LEVELS - level);
  return (key >> shift) << shift;
}
inline int dd_lcp(unsigned long long a, unsigned long long b, int lmax) {
  const unsigned long long x = a ^ b;
  if (x == 0) return lmax;
  const int l = (__builtin_clzll(x) - 1) / 3;
  return l < lmax ? l : lmax;
}

// The straddling cells: every common ancestor cell of the last key of one non-empty rank and the first key of the next.
int dd_make_cells(sph_ctx* c, const std::vector<DDInfo>& info, std::vector<DDCell>& cells) {
  cells.clear();
  std::vector<int> ne;
  for (int r = 0; r < c->n_ranks; ++r) if (info[r].n_own > 0) ne.push_back(r);
  std::vector<std::pair<int, unsigned long long>> set;                     // (level, prefix)
  for (size_t b = 0; b + 1 < ne.size(); ++b) {
    const unsigned long long L = info[ne[b]].last_key, F = info[ne[b + 1]].first_key;
    const int lam = dd_lcp(L, F, c->dp.lmax);
    if (lam >= c->dp.lmax) { c->err = "domain decomposition: equal descent keys on both sides of a domain boundary"; return SPH_ERR_STATE; }
    for (int l = 0; l <= lam; ++l) set.push_back({l, dd_prefix(L, l)});
  }
  std::sort(set.begin(), set.end(), [](const std::pair<int, unsigned long long>& a, const std::pair<int, unsigned long long>& b) {
    return a.second != b.second ? a.second < b.second : a.first < b.first; });           // preorder: by prefix, then level
  set.erase(std::unique(set.begin(), set.end()), set.end());
  if ((int)set.size() > DD_MAX_CELLS) { c->err = "domain decomposition: too many straddling cells"; return SPH_ERR_STATE; }
  for (auto& s : set) {
    DDCell C; C.level = s.first; C.prefix = s.second; C.child_top = 0;
    if (C.level < SPH_KEY_LEVELS)
      for (int o = 0; o < 8; ++o) {
        const unsigned long long cp = C.prefix | ((unsigned long long)o << (3 * (SPH_KEY_LEVELS - 1 - C.level)));
        for (auto& t : set) if (t.first == C.level + 1 && t.second == cp) C.child_top |= (1 << o);
      }
    cells.push_back(C);
  }
  return SPH_OK;
}

// ---- halo ------------------------------------------------------------------------------------------------------------------
int dd_fill_peer_groups(sph_ctx* c, DDPeerGroups& pg) {
  std::memset(&pg, 0, sizeof(pg));
  int acc = 0;
  for (int q = 0; q < DD_MAX_RANKS; ++q) {
    pg.goff[q] = acc;
    if (q < c->n_ranks && q != c->rank) {
      pg.ng[q] = c->dd_info[q].ng_own; pg.groups[q] = dd_peer<const int2>(c, DS_GROUPS, q); pg.box[q] = dd_peer<const BvhBox>(c, DS_BVH, q);
      acc += pg.ng[q];
    }
  }
  pg.goff[DD_MAX_RANKS] = acc;
  return acc;
}

// Select the peer groups this domain needs, place their particles behind the own ones, pull what the density pass and
// the pair loop read of them (everything but rho c P/(Omega rho^2), which follow after the density pass), and build the
// group table + BVH over [own | halo].  Collective (one barrier before: the peers' boxes and fields are final; one after).
int dd_halo(sph_ctx* c) {
  const int T = 256, n_own = (int)c->n, ng_own = c->ng_own;
  StateArrays s = state_of(c, c->cur);
  // own level-0 boxes (from the current reach) and the domain boxes derived from them
  LAUNCH(k_bvh_leaf, cdiv((int64_t)std::max(ng_own, 1) * 32, T), T, 0, ng_own, c->groups, s.x, s.y, s.z, c->lcx, c->lcy, c->lcz, c->reach, c->bvh);
  {   // own-only BVH above them (private): what the halo selection and the LET criterion walk
    if ((size_t)ng_own / 7 + 64 > c->dd_obvh_cap) { c->dd_obvh_cap = (size_t)c->cap / 7 + 64; DA(c->dd_obvh, c->dd_obvh_cap); }
    DDBvh& ob = c->dd_ob;
    int cntl = ng_own, offu = 0, l = 0;
    ob.off[0] = 0; ob.cnt[0] = cntl;
    const BvhBox* src = c->bvh;
    while (cntl > 32) {
      const int np = cdiv(cntl, SPH_BVH_FAN);
      LAUNCH(k_bvh_up, cdiv(np, T), T, 0, cntl, src, c->dd_obvh + offu);
      ob.off[l + 1] = offu; ob.cnt[l + 1] = np;
      src = c->dd_obvh + offu; offu += np; cntl = np; ++l;
    }
    ob.nlev = l + 1;
  }
  { int r_ = dd_barrier(c); if (r_) return r_; }
  DDPeerGroups pg; const int ngp = dd_fill_peer_groups(c, pg);
  c->dd_pg = pg;
  if ((size_t)ngp + 1 > c->dd_halo_cap) {
    c->dd_halo_cap = (size_t)ngp * 5 / 4 + 1024;
    DA(c->dd_halo_flag, c->dd_halo_cap); DA(c->dd_halo_list, c->dd_halo_cap); DA(c->dd_halo_size, c->dd_halo_cap); DA(c->dd_halo_poff, c->dd_halo_cap);
    size_t b1 = 0, b2 = 0;
    cub::DeviceSelect::Flagged(nullptr, b1, cub::CountingInputIterator<int>(0), c->dd_halo_flag, c->dd_halo_list, c->d_nsel, (int)c->dd_halo_cap, c->stream);
    cub::DeviceScan::ExclusiveSum(nullptr, b2, c->dd_halo_size, c->dd_halo_poff, (int)c->dd_halo_cap, c->stream);
    if (std::max(b1, b2) + 256 > c->cub_bytes) { c->cub_bytes = std::max(b1, b2) + 256; if (c->cub_tmp) cudaFree(c->cub_tmp); c->cub_tmp = nullptr; if (cudaMalloc(&c->cub_tmp, c->cub_bytes) != cudaSuccess) { c->err = "cudaMalloc(cub temp)"; return SPH_ERR_OOM; } }
  }
  int nh = 0, n_halo = 0;
  {   // one read-back: selected groups, their particles, and whether ANY rank ran out of room (every rank stops together)
    int h3[3] = {0, 0, 0};
    CK(cudaMemsetAsync(c->dd_let_ctl + 60, 0, 3 * sizeof(int), c->stream));
    if (ngp > 0) {
      LAUNCH(k_dd_halo_mark, cdiv(ngp + 1, T), T, 0, pg, c->bvh, c->dd_obvh, c->dd_ob, c->dd_halo_flag, c->dd_halo_size);
      size_t bytes = c->cub_bytes;
      CK(cub::DeviceSelect::Flagged(c->cub_tmp, bytes, cub::CountingInputIterator<int>(0), c->dd_halo_flag, c->dd_halo_list, c->d_nsel, ngp, c->stream));
      bytes = c->cub_bytes;
      CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp, bytes, c->dd_halo_size, c->dd_halo_poff, ngp + 1, c->stream));
      LAUNCH(k_dd_halo_check, 1, 1, 0, c->d_nsel, c->dd_halo_poff + ngp, n_own, ng_own, (long long)c->cap, c->dd_let_ctl + 60);
    }
    { int r_ = coll_allreduce(c, c->stream, c->dd_let_ctl + 62, 1, NC_INT32, NC_MAX); if (r_) return r_; }
    CK(cudaMemcpyAsync(h3, c->dd_let_ctl + 60, sizeof(h3), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    nh = h3[0]; n_halo = h3[1];
    if (h3[2]) {
      if ((int64_t)n_own + n_halo > c->cap || (int64_t)ng_own + nh > c->cap)
        c->err = "domain decomposition: own + halo particles (" + std::to_string(n_own) + " + " + std::to_string(n_halo) + ") exceed the rank's capacity " +
                 std::to_string((long long)c->cap) + " (raise SPH_B200_DOMAIN_SLACK)";
      else c->err = "domain decomposition: a peer rank's own + halo particles exceed its capacity (see its error); every rank returns";
      return SPH_ERR_OOM;
    }
  }
  c->n_halo = n_halo; c->ng_halo = nh; c->n_groups = ng_own + nh;
  if (nh > 0) {
    DDPull a; std::memset(&a, 0, sizeof(a));
    a.nf = 0; a.with_id = 1; a.write_groups = 1; a.ng_own = ng_own; a.n_own = n_own;
    auto add_state = [&](int f) { for (int q = 0; q < c->n_ranks; ++q) if (q != c->rank) a.src[q][a.nf] = dd_peer<const double>(c, DS_ST + c->dd_info[q].cur * 10 + f, q); a.dst[a.nf] = c->st[c->cur][f]; ++a.nf; };
    auto add_slot = [&](int slot, double* dst) { for (int q = 0; q < c->n_ranks; ++q) if (q != c->rank) a.src[q][a.nf] = dd_peer<const double>(c, slot, q); a.dst[a.nf] = dst; ++a.nf; };
    for (int f = 0; f < 10; ++f) if (f != 6) add_state(f);                 // x y z vx vy vz m alpha h  (u is never read of a source)
    add_slot(DS_LCX, c->lcx); add_slot(DS_LCY, c->lcy); add_slot(DS_LCZ, c->lcz); add_slot(DS_REACH, c->reach);
    for (int q = 0; q < c->n_ranks; ++q) if (q != c->rank) a.src_id[q] = dd_peer<const int>(c, DS_ID + c->dd_info[q].cur, q);
    a.dst_id = c->id[c->cur];
    LAUNCH(k_dd_pull, cdiv((int64_t)nh * 32, T), T, 0, nh, c->dd_halo_list, c->dd_halo_poff, pg, a, c->groups);
  }
  // BVH over [own | halo]
  {
    const int ng = c->n_groups;
    BvhInfo& bi = c->bi;
    int cntl = ng, offl = 0, l = 0;
    bi.off[0] = 0; bi.cnt[0] = cntl;
    if (nh > 0) LAUNCH(k_bvh_leaf, cdiv((int64_t)nh * 32, T), T, 0, nh, c->groups + ng_own, s.x, s.y, s.z, c->lcx, c->lcy, c->lcz, c->reach, c->bvh + ng_own);
    while (cntl > 32) {
      int np = cdiv(cntl, SPH_BVH_FAN);
      bi.off[l + 1] = offl + cntl; bi.cnt[l + 1] = np;
      LAUNCH(k_bvh_up, cdiv(np, T), T, 0, cntl, c->bvh + offl, c->bvh + offl + cntl);
      offl += cntl; cntl = np; ++l;
    }
    bi.nlev = l + 1;
  }
  // no trailing barrier: what the peers pull (own-region state, leaf cells, group table, level-0 boxes) is next written by
  // the kick / the next build, and collectives lie in between (the density pass's "lists void" all-reduce is the first)
  return SPH_OK;
}

// after the density pass: the halo particles' rho, c and P/(Omega rho^2) (what the pair loop reads of its sources)
int dd_pull_density_fields(sph_ctx* c) {      // called right behind the evaluation's all-reduce: every rank's density pass has finished
  if (c->ng_halo > 0) {
    DDPull a; std::memset(&a, 0, sizeof(a));
    a.nf = 0; a.with_id = 0; a.write_groups = 0; a.ng_own = c->ng_own; a.n_own = (int)c->n;
    auto add_slot = [&](int slot, double* dst) { for (int q = 0; q < c->n_ranks; ++q) if (q != c->rank) a.src[q][a.nf] = dd_peer<const double>(c, slot, q); a.dst[a.nf] = dst; ++a.nf; };
    add_slot(DS_RHO, c->rho); add_slot(DS_CS, c->cs); add_slot(DS_POR2, c->por2);
    LAUNCH(k_dd_pull, cdiv((int64_t)c->ng_halo * 32, 256), 256, 0, c->ng_halo, c->dd_halo_list, c->dd_halo_poff, c->dd_pg, a, c->groups);
  }
  return SPH_OK;      // rho c P/(Omega rho^2) of the own region are next written by calc_smoothing / the next density pass: the sink all-reduce of run_gravity lies in between
}

// ---- locally essential tree ----------------------------------------------------------------------------------------------
// Runs on the exchange stream, under the density pass (only the gravity walk reads it: run_gravity waits for let_done).
// Nothing is read back here: an overflow raises `let_flag`, which travels with the sink all-reduce of run_gravity and
// becomes the sticky device error 5 on EVERY rank, seen at the step's (or sph_evaluate's) own read-back.
__global__ void k_dd_let_flag(const int* __restrict__ ctl, double* __restrict__ flag, int force) { *flag = (ctl[1] || force) ? 1.0 : 0.0; }
// flag[0] = LET overflow (set by k_dd_let_flag), flag[1] = this rank's "candidate lists void" (nl_ctl[1]); summed over the ranks
__global__ void k_dd_flags_pack(const int* __restrict__ nl_ctl, double* __restrict__ flag) { flag[1] = nl_ctl[1] ? 1.0 : 0.0; }
__global__ void k_dd_flags_apply(const double* __restrict__ flag, int* err, int* __restrict__ nl_ctl) {
  if (flag[0] > 0.0) atomicExch(err, 5);
  if (flag[1] > 0.0) nl_ctl[1] = 1;
}
#define LAUNCH_X(kern, grid, block, ...) do { kern<<<(grid), (block), 0, c->xstream>>>(__VA_ARGS__); ++c->launches; } while (0)
int dd_build_let(sph_ctx* c, const std::vector<DDLetEntry>& cand) {
  const int T = 256;
  const size_t fcap = (size_t)c->dd_let_fcap;
  if (!c->let_done) CK(cudaEventCreateWithFlags(&c->let_done, cudaEventDisableTiming));
  CK(cudaEventRecord(c->x_ready, c->stream));
  CK(cudaStreamWaitEvent(c->xstream, c->x_ready, 0));
  CK(cudaMemsetAsync(c->dd_let_ctl, 0, 60 * sizeof(int), c->xstream));
  CK(cudaMemcpyAsync(c->dd_let_ctl, &c->dd_let_begin, sizeof(int), cudaMemcpyHostToDevice, c->xstream));        // cursor
  const bool too_many = cand.size() > fcap;
  if (!cand.empty() && !too_many) {
    c->dd_cand_host = cand;                              // stays alive while the copy is in flight
    DDLetEntry* seed = c->dd_let_f[1];
    CK(cudaMemcpyAsync(seed, c->dd_cand_host.data(), cand.size() * sizeof(DDLetEntry), cudaMemcpyHostToDevice, c->xstream));
    const double theta2 = c->dp.theta * c->dp.theta;
    LAUNCH_X(k_dd_let_seed, cdiv((int)cand.size(), T), T, (int)cand.size(), seed, c->dd_let_f[0], c->dd_let_ctl + 2, c->wnodes, c->bvh, c->dd_obvh, c->dd_ob, theta2);
    DDPeerNodes pn; std::memset(&pn, 0, sizeof(pn));
    for (int q = 0; q < c->n_ranks; ++q) if (q != c->rank) pn.wn[q] = dd_peer<const WNode>(c, DS_WN, q);
    const int levels = SPH_KEY_LEVELS + 3;              // a compressed tree is at most lmax <= 21 branching levels deep
    for (int l = 0; l < levels; ++l)
      LAUNCH_X(k_dd_let_level, c->n_sm * 4, T, c->dd_let_f[l & 1], c->dd_let_ctl + 2 + l, c->dd_let_f[(l + 1) & 1], c->dd_let_ctl + 3 + l, (int)fcap,
               c->dd_let_ctl, c->dd_let_end, c->dd_let_ctl + 1, pn, c->wnodes, c->bvh, c->dd_obvh, c->dd_ob, theta2);
  }
  LAUNCH_X(k_dd_let_flag, 1, 1, c->dd_let_ctl, c->let_flag, too_many ? 1 : 0);
  CK(cudaEventRecord(c->let_done, c->xstream));
  c->let_pending = true;
  return SPH_OK;
}

// ---- the tree build ---------------------------------------------------------------------------------------------------------
int dd_build_tree(sph_ctx* c) {
  const int T = 256, R = c->n_ranks;
  int n = (int)c->n;
  c->nl_valid = false; c->grav_groups_valid = false;
  if (c->two_word) { c->err = "domain decomposition supports single-word (21-level) descent keys only"; return SPH_ERR_DEPTH; }
  { int r_ = dd_alloc(c); if (r_) return r_; }
  if (c->p2p_stale) { int r_ = p2p_setup(c); if (r_) return r_; if (!c->p2p_ok) { c->err = "domain decomposition needs peer-mapped device memory between the ranks (NVLink / PCIe peer access, CUDA IPC)"; return SPH_ERR_COMM; } }
  // ---- root cube of ALL gas particles: F:803-808 over the all-reduced extrema (min / max are exact: the same cube as one rank's)
  stage_begin(c, ST_KEYS);
  {
    StateArrays s = state_of(c, c->cur);
    int nb = std::max(1, std::min(cdiv(n, T), c->n_partial));
    LAUNCH(k_bbox_partial, nb, T, 0, n, s.x, s.y, s.z, c->partial);
    LAUNCH(k_bbox_final, 1, 32, 0, nb, c->partial, c->root);
    LAUNCH(k_dd_box_pack, 1, 32, 0, c->root, c->partial);            // [mn, -mx] so that one MIN all-reduce merges both
    { int r_ = coll_allreduce(c, c->stream, c->partial, 6, NC_FLOAT64, NC_MIN); if (r_) return r_; }
    LAUNCH(k_dd_box_unpack, 1, 32, 0, c->partial);
    LAUNCH(k_bbox_final, 1, 32, 0, 1, c->partial, c->root);
    LAUNCH(k_keys, cdiv(std::max(n, 1), T), T, 0, n, s.x, s.y, s.z, c->root, c->dp.lmax, c->key[0], nullptr, c->perm[0]);
  }
  stage_end(c);
  // ---- local order, splitters, migration
  RootBox rb; std::vector<DDInfo>& info = c->dd_info;
  stage_begin(c, ST_SORT);
  {
    if (n > 0) {
      size_t bytes = c->cub_bytes;
      cub::DoubleBuffer<uint64_t> dk(c->key[0], c->key[1]); cub::DoubleBuffer<int> dv(c->perm[0], c->perm[1]);
      CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp, bytes, dk, dv, n, 0, 63, c->stream));
      if (dk.Current() != c->key[0]) std::swap(c->key[0], c->key[1]);
      if (dv.Current() != c->perm[0]) std::swap(c->perm[0], c->perm[1]);
    }
    stage_end(c); stage_begin(c, ST_MIGRATE);
    LAUNCH(k_dd_samples, cdiv(DD_SAMPLES, T), T, 0, n, c->key[0], DD_SAMPLES, c->dd_samples + (size_t)c->rank * DD_SAMPLES);
    { int r_ = coll_allgather(c, c->stream, c->dd_samples + (size_t)c->rank * DD_SAMPLES, c->dd_samples, (size_t)DD_SAMPLES * 8); if (r_) return r_; }
    long long cnt_mine = n;
    CK(cudaMemcpyAsync(c->dd_counts + c->rank, &cnt_mine, 8, cudaMemcpyHostToDevice, c->stream));
    { int r_ = coll_allgather(c, c->stream, c->dd_counts + c->rank, c->dd_counts, 8); if (r_) return r_; }
    if (R > 1) LAUNCH(k_dd_splitters, R - 1, 256, (size_t)R * DD_SAMPLES * 8, R, DD_SAMPLES, c->dd_samples, c->dd_counts, c->dd_split);
    LAUNCH(k_dd_segments, 1, 32, 0, R, n, c->key[0], c->dd_split, c->dd_sendoff, c->dd_segkeys);
    // every rank's segment table (offsets + boundary keys) + which of its buffers hold the (unsorted) state, the sorted keys and the permutation
    struct Seg { int off[DD_MAX_RANKS + 1]; int cur, key_slot, perm_slot; long long cap; unsigned long long first[DD_MAX_RANKS], last[DD_MAX_RANKS]; } mine, all[DD_MAX_RANKS];
    std::memset(&mine, 0, sizeof(mine));
    CK(cudaMemcpyAsync(mine.off, c->dd_sendoff, sizeof(int) * (R + 1), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(mine.first, c->dd_segkeys, sizeof(unsigned long long) * 2 * DD_MAX_RANKS, cudaMemcpyDeviceToHost, c->stream));      // first[] and last[] are adjacent
    CK(cudaMemcpyAsync(&rb, c->root, sizeof(rb), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    mine.cap = c->cap; mine.cur = c->cur; mine.key_slot = c->key[0] == c->key_alloc[0] ? 0 : 1; mine.perm_slot = c->perm[0] == c->perm_alloc[0] ? 0 : 1;
    { int r_ = dd_allgather_host(c, &mine, sizeof(Seg), all); if (r_) return r_; }
    // what every domain looks like after the migration: size, boundary keys, which state buffer is current
    info.assign(R, DDInfo());
    c->n_global = 0;
    for (int q = 0; q < R; ++q) {
      DDInfo& I = info[q]; std::memset(&I, 0, sizeof(I));
      bool q_moved = false; bool any = false;
      for (int pq = 0; pq < R; ++pq) {
        const int cnt = all[pq].off[q + 1] - all[pq].off[q];
        if (cnt <= 0) continue;
        if (pq != q) q_moved = true;
        I.n_own += cnt;
        if (!any || all[pq].first[q] < I.first_key) I.first_key = all[pq].first[q];
        if (!any || all[pq].last[q] > I.last_key) I.last_key = all[pq].last[q];
        any = true;
      }
      I.cur = all[q].cur ^ 1; if (q_moved && I.n_own > 0) I.cur ^= 1;
      c->n_global += I.n_own;
    }
    DDMigrate m; std::memset(&m, 0, sizeof(m));
    m.R = R;
    // destination order: the rows this rank keeps first (one sorted run), then the rows arriving from rank 0, 1, ...
    int acc = 0; bool moved = false; int n_kept = 0;
    std::vector<int> order; order.push_back(c->rank); for (int q = 0; q < R; ++q) if (q != c->rank) order.push_back(q);
    for (int slot = 0; slot < R; ++slot) {
      const int q = order[slot];
      m.dst_off[slot] = acc; m.src_off[slot] = all[q].off[c->rank];
      const int cnt = all[q].off[c->rank + 1] - all[q].off[c->rank];
      if (q != c->rank && cnt > 0) moved = true;
      if (q == c->rank) n_kept = cnt;
      acc += cnt;
      for (int f = 0; f < 10; ++f) m.st[slot][f] = q == c->rank ? c->st[c->cur][f] : dd_peer<const double>(c, DS_ST + all[q].cur * 10 + f, q);
      m.id[slot] = q == c->rank ? c->id[c->cur] : dd_peer<const int>(c, DS_ID + all[q].cur, q);
      m.key[slot] = q == c->rank ? c->key[0] : dd_peer<const uint64_t>(c, DS_KEY + all[q].key_slot, q);
      m.perm[slot] = q == c->rank ? c->perm[0] : dd_peer<const int>(c, DS_PERM + all[q].perm_slot, q);
    }
    m.dst_off[R] = acc;
    for (int q = 0; q < R; ++q) {      // every rank holds every segment table: all ranks find the same overfull rank and stop together
      long long in = 0; for (int pq = 0; pq < R; ++pq) in += all[pq].off[q + 1] - all[pq].off[q];
      if (in > all[q].cap) { c->err = "domain decomposition: " + std::to_string(in) + " particles migrate to rank " + std::to_string(q) + ", capacity " + std::to_string(all[q].cap) + " (raise SPH_B200_DOMAIN_SLACK)"; return SPH_ERR_OOM; }
    }
    for (int f = 0; f < 10; ++f) m.dst[f] = c->st[c->cur ^ 1][f];
    m.dst_id = c->id[c->cur ^ 1]; m.dst_key = c->key[1];
    if (acc > 0) LAUNCH(k_dd_migrate, cdiv(acc, T), T, 0, m);
    { int r_ = dd_barrier(c); if (r_) return r_; }      // every pull has finished: the source buffers may be reused
    c->cur ^= 1; std::swap(c->key[0], c->key[1]);
    c->n = n = acc;
    stage_end(c); stage_begin(c, ST_SORT);
    if (moved && n > 0) {      // [kept rows: sorted | arrivals: a few sorted runs]: sort the arrivals, MERGE the two, move the state once more
      const int nf = n - n_kept;
      LAUNCH(k_iota, cdiv(n, T), T, 0, n, c->perm[0]);
      cub::DoubleBuffer<uint64_t> dk(c->key[0] + n_kept, c->key[1] + n_kept); cub::DoubleBuffer<int> dv(c->perm[0] + n_kept, c->perm[1] + n_kept);
      size_t bytes = c->cub_bytes;
      CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp, bytes, dk, dv, nf, 0, 63, c->stream));
      uint64_t* mkey = reinterpret_cast<uint64_t*>(c->acc_key[0]); int* mperm = c->acc_val[0];      // scratch of the cull stage
      size_t mb = 0;
      CK(cub::DeviceMerge::MergePairs(nullptr, mb, c->key[0], c->perm[0], n_kept, dk.Current(), dv.Current(), nf, mkey, mperm, cuda::std::less<uint64_t>{}, c->stream));
      if (mb + 256 > c->cub_bytes) { c->cub_bytes = mb + 256; if (c->cub_tmp) cudaFree(c->cub_tmp); c->cub_tmp = nullptr; if (cudaMalloc(&c->cub_tmp, c->cub_bytes) != cudaSuccess) { c->err = "cudaMalloc(cub temp)"; return SPH_ERR_OOM; } }
      mb = c->cub_bytes;
      CK(cub::DeviceMerge::MergePairs(c->cub_tmp, mb, c->key[0], c->perm[0], n_kept, dk.Current(), dv.Current(), nf, mkey, mperm, cuda::std::less<uint64_t>{}, c->stream));
      CK(cudaMemcpyAsync(c->key[0], mkey, (size_t)n * 8, cudaMemcpyDeviceToDevice, c->stream));
      PermuteArgs pa;
      for (int f = 0; f < 10; ++f) { pa.src[f] = c->st[c->cur][f]; pa.dst[f] = c->st[c->cur ^ 1][f]; }
      pa.id_src = c->id[c->cur]; pa.id_dst = c->id[c->cur ^ 1];
      LAUNCH(k_permute, cdiv(n, 4 * T), T, 0, n, mperm, pa);
      c->cur ^= 1;
    }
  }
  stage_end(c);
  for (int q = 0; q < R; ++q) if (info[q].n_own < 2) { c->err = "domain decomposition: rank " + std::to_string(q) + " holds fewer than 2 particles"; return SPH_ERR_STATE; }      // the same verdict on every rank
  // ---- local octree (leaf levels see the neighbouring domains' boundary keys)
  stage_begin(c, ST_TREE);
  StateArrays s = state_of(c, c->cur);
  unsigned long long kprev = 0, knext = 0; int has_prev = 0, has_next = 0;
  for (int q = c->rank - 1; q >= 0; --q) if (info[q].n_own > 0) { kprev = info[q].last_key; has_prev = 1; break; }
  for (int q = c->rank + 1; q < R; ++q) if (info[q].n_own > 0) { knext = info[q].first_key; has_next = 1; break; }
  LAUNCH(k_leaf, cdiv(n, T), T, 0, n, c->key[0], nullptr, s.h, c->root, c->dp, c->level, c->lcx, c->lcy, c->lcz, c->reach, &c->sc->err, has_prev, kprev, has_next, knext);
  LAUNCH(k_oct_nodes<false>, cdiv(n, T), T, 0, n, c->key[0], nullptr, c->dp.lmax, c->root, c->cnt, c->off, 0, c->nodes, c->node_part, c->node_count);
  CK(cudaMemsetAsync(c->cnt + n, 0, sizeof(int), c->stream));
  size_t bytes = c->cub_bytes;
  CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp, bytes, c->cnt, c->off, n + 1, c->stream));
  int n_int = 0, key_err = 0;
  CK(cudaMemcpyAsync(&n_int, c->off + n, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaMemcpyAsync(&key_err, &c->sc->err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  if (key_err) { c->err = "domain decomposition: particles share a full 63-bit descent key while max_depth is deeper (the two-word path is single-rank only)"; return SPH_ERR_DEPTH; }
  const int nn = n + n_int;
  c->counts.n_nodes = nn;
  LAUNCH(k_oct_nodes<true>, cdiv(n, T), T, 0, n, c->key[0], nullptr, c->dp.lmax, c->root, c->cnt, c->off, nn, c->nodes, c->node_part, c->node_count);
  LAUNCH(k_oct_link, cdiv(nn, T), T, 0, nn, c->nodes, c->node_part, c->parent, c->nchild, c->wcount);
  bytes = c->cub_bytes;
  CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp, bytes, c->wcount, c->wstart, nn, c->stream));
  CK(cudaMemsetAsync(c->widx, 0xff, sizeof(int) * (size_t)nn, c->stream));
  LAUNCH(k_oct_widx, cdiv(nn, T), T, 0, nn, c->nodes, c->wcount, c->wstart, c->widx);
  CK(cudaMemsetAsync(c->arrive, 0, sizeof(int) * (size_t)nn, c->stream));
  LAUNCH(k_oct_up, cdiv(n, T), T, 0, n, c->off, c->cnt, s.x, s.y, s.z, s.m, s.h, c->level, c->root, c->nodes, c->parent, c->nchild, c->arrive);
  // ---- walk groups of the own particles (marked here so that their count travels with the top-tree records)
  CK(cudaMemsetAsync(c->gsize, 0, sizeof(int) * (size_t)n, c->stream));
  LAUNCH(k_group_mark, cdiv(nn, T), T, 0, nn, c->nodes, c->node_part, c->node_count, c->gsize);
  bytes = c->cub_bytes;
  CK(cub::DeviceSelect::Flagged(c->cub_tmp, bytes, cub::CountingInputIterator<int>(0), c->gsize, c->gfirst, c->d_nsel, n, c->stream));
  // ---- top tree: straddling cells, their locally complete children from every rank, the single-rank summation order
  std::vector<DDCell> cells;
  { int r_ = dd_make_cells(c, info, cells); if (r_) return r_; }
  const int ncell = (int)cells.size();
  if (ncell == 0) { c->err = "domain decomposition: all particles on one rank"; return SPH_ERR_STATE; }
  std::vector<DDContrib> contrib_all((size_t)R * ncell * 8);
  int ng = 0;
  {
    CK(cudaMemcpyAsync(c->dd_cells, cells.data(), sizeof(DDCell) * ncell, cudaMemcpyHostToDevice, c->stream));
    LAUNCH(k_dd_top_contrib, cdiv(ncell * 8, 64), 64, 0, ncell, c->dd_cells, n, c->key[0], c->off, c->cnt, c->node_count, c->nodes, c->wcount, c->wstart, c->dd_contrib, &c->sc->err);
    // one read-back and one all-gather carry the records AND the own group count (slot 0's count field of an extra record)
    const size_t nrec = (size_t)ncell * 8 + 1;
    std::vector<DDContrib> mine(nrec), all(nrec * R);
    CK(cudaMemcpyAsync(mine.data(), c->dd_contrib, sizeof(DDContrib) * ncell * 8, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaMemcpyAsync(&ng, c->d_nsel, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    std::memset(&mine[nrec - 1], 0, sizeof(DDContrib)); mine[nrec - 1].count = ng;
    { int r_ = dd_allgather_host(c, mine.data(), sizeof(DDContrib) * nrec, all.data()); if (r_) return r_; }
    for (int q = 0; q < R; ++q) {
      std::memcpy(&contrib_all[(size_t)q * ncell * 8], &all[(size_t)q * nrec], sizeof(DDContrib) * ncell * 8);
      info[q].ng_own = all[(size_t)q * nrec + nrec - 1].count;
    }
  }
  c->ng_own = ng; c->n_groups = ng;
  LAUNCH(k_group_pack, cdiv(ng, T), T, 0, ng, c->gfirst, c->gsize, c->groups);
  c->rank_g.assign(R + 1, 0); c->rank_p.assign(R + 1, 0);
  c->g0 = 0; c->g1 = ng; c->p0 = 0; c->p1 = n;
  // local nodes into the walk layout, behind the top region
  LAUNCH(k_oct_finalize, cdiv(nn, T), T, 0, nn, c->nodes, c->wcount, c->wstart, c->widx, c->wnodes, DD_TOP_CAP);
  std::vector<DDLetEntry> cand;
  {
    // resolve(cell index) -> a reference to the node that stands for the cell in the compressed tree
    struct Ref { int kind; int cell; int owner; DDContrib rec; };      // kind 0: nothing, 1: top node (cell), 2: a rank's complete node
    std::vector<DDTopNode> tn(ncell);
    std::vector<int> is_node(ncell, 0);
    auto cell_index = [&](int level, unsigned long long prefix) { for (int i = 0; i < ncell; ++i) if (cells[i].level == level && cells[i].prefix == prefix) return i; return -1; };
    // children of a cell in octant order: top cells recurse, others come from the one rank that reported particles there
    std::function<Ref(int)> resolve = [&](int ci) -> Ref {
      const DDCell& C = cells[ci];
      std::vector<Ref> kids;
      for (int o = 0; o < 8; ++o) {
        if ((C.child_top >> o) & 1) {
          const unsigned long long cp = C.prefix | ((unsigned long long)o << (3 * (SPH_KEY_LEVELS - 1 - C.level)));
          Ref r = resolve(cell_index(C.level + 1, cp));
          if (r.kind) kids.push_back(r);
        } else {
          for (int q = 0; q < R; ++q) {
            const DDContrib& rec = contrib_all[((size_t)q * ncell + ci) * 8 + o];
            if (rec.count > 0) { Ref r; r.kind = 2; r.cell = -1; r.owner = q; r.rec = rec; kids.push_back(r); break; }
          }
        }
      }
      if (kids.empty()) { Ref r; r.kind = 0; r.cell = -1; r.owner = -1; return r; }
      if (kids.size() == 1) return kids[0];                              // single-child chain: collapses
      DDTopNode& t = tn[ci];
      t.cell = ci; t.nchild = (int)kids.size();
      double M = 0.0, sx = 0.0, sy = 0.0, sz = 0.0;                      // k_oct_up: children added in depth-first order, starting from 0
      for (size_t k = 0; k < kids.size(); ++k) {
        const Ref& r = kids[k];
        if (r.kind == 1) { t.kid_is_top[k] = true; t.kids[k] = r.cell; t.kid_owner[k] = -1; M += tn[r.cell].m; sx += tn[r.cell].sx; sy += tn[r.cell].sy; sz += tn[r.cell].sz; }
        else { t.kid_is_top[k] = false; t.kid_owner[k] = r.owner; t.kids[k] = (int)c->dd_top_recs.size(); c->dd_top_recs.push_back(r.rec); M += r.rec.m; sx += r.rec.sx; sy += r.rec.sy; sz += r.rec.sz; }
      }
      t.m = M; t.sx = sx; t.sy = sy; t.sz = sz;
      is_node[ci] = 1;
      Ref r; r.kind = 1; r.cell = ci; r.owner = -1; return r;
    };
    c->dd_top_recs.clear();
    std::vector<WNode>& top = c->dd_top_host; top.clear();      // a member: stays alive while the copy below is in flight
    auto make_wnode = [&](double m, double sx, double sy, double sz, double size) {
      WNode w; w.m = m; w.size = size; w.child = 0; w.nchild = 0;
      if (m > 0.0) { w.cx = sx / m; w.cy = sy / m; w.cz = sz / m; } else { w.cx = sx; w.cy = sy; w.cz = sz; }          // k_oct_finalize, F:173-177
      return w;
    };
    auto cell_size = [&](int level) { double sz = rb.size; for (int q = 0; q < level; ++q) sz = sz * 0.5; return sz; };
    Ref root = resolve(0);                                              // cells[0] = (level 0, prefix 0): the root cube
    if (root.kind != 1) { c->err = "domain decomposition: the root cell does not branch across ranks"; return SPH_ERR_STATE; }
    // walk layout of the top nodes: root at slot 0, every top node's children in one contiguous block
    top.push_back(make_wnode(tn[root.cell].m, tn[root.cell].sx, tn[root.cell].sy, tn[root.cell].sz, cell_size(cells[root.cell].level)));
    std::vector<std::pair<int, int>> queue;  queue.push_back({root.cell, 0});           // (top cell, slot of its WNode)
    for (size_t qi = 0; qi < queue.size(); ++qi) {
      const DDTopNode& t = tn[queue[qi].first];
      const int slot = queue[qi].second, block = (int)top.size();
      top[slot].child = block; top[slot].nchild = t.nchild;
      top.resize(top.size() + t.nchild);
      for (int k = 0; k < t.nchild; ++k) {
        if (t.kid_is_top[k]) {
          const DDTopNode& u = tn[t.kids[k]];
          top[block + k] = make_wnode(u.m, u.sx, u.sy, u.sz, cell_size(cells[u.cell].level));
          queue.push_back({u.cell, block + k});
        } else {
          const DDContrib& rec = c->dd_top_recs[t.kids[k]];
          WNode w = make_wnode(rec.m, rec.sx, rec.sy, rec.sz, rec.size);
          w.nchild = rec.nchild; w.child = DD_TOP_CAP + rec.child;         // index in the OWNER's array (valid here when the owner is this rank)
          top[block + k] = w;
          if (t.kid_owner[k] != c->rank && rec.nchild > 0) cand.push_back(DDLetEntry{block + k, t.kid_owner[k], DD_TOP_CAP + rec.child, rec.nchild});
        }
      }
    }
    if ((int)top.size() > DD_TOP_CAP) { c->err = "domain decomposition: top tree exceeds DD_TOP_CAP"; return SPH_ERR_STATE; }
    c->dd_top_n = (int)top.size();
    CK(cudaMemcpyAsync(c->wnodes, top.data(), sizeof(WNode) * top.size(), cudaMemcpyHostToDevice, c->stream));
  }
  {
    const int nst = cdiv(ng, GRAV_SEG), nsk = std::max(c->n_sink, 1);
    if ((size_t)nst * nsk * 3 > c->sink_seg_cap) { c->sink_seg_cap = (size_t)(nst + nst / 8 + 64) * (nsk + 1) * 3; DA(c->sink_seg, c->sink_seg_cap); }
  }
  stage_end(c);
  // ---- halo (also computes this domain's boxes and crosses a barrier: the peers' node arrays are final), then the LET
  stage_begin(c, ST_HALO);
  { int r_ = dd_halo(c); if (r_) return r_; }
  stage_end(c); stage_begin(c, ST_LET);
  { int r_ = dd_build_let(c, cand); if (r_) return r_; }
  // (the peers read the local-tree region of this rank's node array only; it is next written by the next build, collectives lie in between)
  stage_end(c);
  c->tree_valid = true; c->pos_moved = false;
  return SPH_OK;
}

// positions unchanged since the last build (evaluation A after evaluation B, no removals anywhere): only h moved, so
// only the reach of the own leaves, the boxes and the halo (whose members and fields follow the reach / h v alpha) change
int dd_refresh_tree(sph_ctx* c) {
  const int n = (int)c->n, T = 256;
  c->nl_valid = false;
  stage_begin(c, ST_TREE);
  if (c->dp.variable_h) {
    StateArrays s = state_of(c, c->cur);
    LAUNCH(k_refresh_reach, cdiv(n, T), T, 0, n, s.h, c->level, c->root, c->dp, c->reach);
  }
  stage_end(c);
  stage_begin(c, ST_HALO);
  { int r_ = dd_halo(c); if (r_) return r_; }
  stage_end(c);
  return SPH_OK;
}

// ---- gathers for the host-facing downloads (tests, saves): every rank ends up with all ranks' own values, in rank order --
int dd_gather(sph_ctx* c, int slot_base, int per_cur /* 0: fixed slot, 1: + the owner's current buffer index (ids), 2: + the owner's sorted-key buffer, 3: + 10 x the owner's current buffer (state fields) */, size_t elem, void* dst_dev) {
  { int r_ = dd_barrier(c); if (r_) return r_; }
  size_t off = 0;
  for (int q = 0; q < c->n_ranks; ++q) {
    const size_t cnt = (size_t)c->dd_info[q].n_own;
    const int slot = slot_base + (per_cur == 1 ? c->dd_info[q].cur : per_cur == 2 ? c->dd_info[q].key_slot : per_cur == 3 ? 10 * c->dd_info[q].cur : 0);
    const void* src = q == c->rank ? dd_exported(c)[slot] : c->peer[slot][q];
    if (cnt) CK(cudaMemcpyAsync((char*)dst_dev + off * elem, src, cnt * elem, cudaMemcpyDefault, c->stream));
    off += cnt;
  }
  return dd_barrier(c);
}

// ---- accretion: every rank applies ALL ranks' accreted (sink, number) records in ascending number (F:497-508) ------------
int dd_accrete(sph_ctx* c, int n_acc, int* n_removed_global) {
  const int T = 256, R = c->n_ranks;
  StateArrays s = state_of(c, c->cur);
  if (n_acc > 0) LAUNCH(k_dd_acc_records, cdiv(n_acc, T), T, 0, n_acc, c->acc_key[0], c->acc_val[0], s.m, s.x, s.y, s.z, s.vx, s.vy, s.vz, c->dd_acc_key, c->dd_acc_rec);
  struct Cnt { int n_acc, n_removed; } mine{n_acc, c->h_sc->n_removed}, all[DD_MAX_RANKS];
  { int r_ = dd_allgather_host(c, &mine, sizeof(Cnt), all); if (r_) return r_; }
  int tot = 0, rem = 0;
  for (int q = 0; q < R; ++q) { tot += all[q].n_acc; rem += all[q].n_removed; }
  *n_removed_global = rem;
  if ((size_t)tot > c->dd_accg_cap) {
    c->dd_accg_cap = (size_t)tot * 2 + 1024;
    for (int b = 0; b < 2; ++b) { DA(c->dd_accg_key[b], c->dd_accg_cap); DA(c->dd_accg_idx[b], c->dd_accg_cap); }
    DA(c->dd_accg_rec, c->dd_accg_cap);
    size_t b1 = 0; cub::DoubleBuffer<unsigned long long> dk(c->dd_accg_key[0], c->dd_accg_key[1]); cub::DoubleBuffer<int> dv(c->dd_accg_idx[0], c->dd_accg_idx[1]);
    cub::DeviceRadixSort::SortPairs(nullptr, b1, dk, dv, (int)c->dd_accg_cap, 0, 64, c->stream);
    if (b1 + 256 > c->cub_bytes) { c->cub_bytes = b1 + 256; if (c->cub_tmp) cudaFree(c->cub_tmp); c->cub_tmp = nullptr; if (cudaMalloc(&c->cub_tmp, c->cub_bytes) != cudaSuccess) { c->err = "cudaMalloc(cub temp)"; return SPH_ERR_OOM; } }
  }
  if (tot > 0) {
    { int r_ = dd_barrier(c); if (r_) return r_; }
    int off = 0;
    for (int q = 0; q < R; ++q) {
      const int cnt = all[q].n_acc;
      if (cnt > 0) {
        const void* sk = q == c->rank ? (const void*)c->dd_acc_key : (const void*)c->peer[DS_ACCKEY][q];
        const void* sr = q == c->rank ? (const void*)c->dd_acc_rec : (const void*)c->peer[DS_ACCREC][q];
        CK(cudaMemcpyAsync(c->dd_accg_key[0] + off, sk, (size_t)cnt * 8, cudaMemcpyDefault, c->stream));
        CK(cudaMemcpyAsync(c->dd_accg_rec + off, sr, (size_t)cnt * sizeof(DDAccRec), cudaMemcpyDefault, c->stream));
      }
      off += cnt;
    }
    { int r_ = dd_barrier(c); if (r_) return r_; }
    LAUNCH(k_iota, cdiv(tot, T), T, 0, tot, c->dd_accg_idx[0]);
    if (tot > 1) {
      cub::DoubleBuffer<unsigned long long> dk(c->dd_accg_key[0], c->dd_accg_key[1]); cub::DoubleBuffer<int> dv(c->dd_accg_idx[0], c->dd_accg_idx[1]);
      size_t bytes = c->cub_bytes;
      CK(cub::DeviceRadixSort::SortPairs(c->cub_tmp, bytes, dk, dv, tot, 0, 64, c->stream));
      if (dk.Current() != c->dd_accg_key[0]) std::swap(c->dd_accg_key[0], c->dd_accg_key[1]);
      if (dv.Current() != c->dd_accg_idx[0]) std::swap(c->dd_accg_idx[0], c->dd_accg_idx[1]);
    }
  }
  LAUNCH(k_dd_accrete_apply, 1, SPH_MAX_SINKS, 0, tot, c->dd_accg_key[0], c->dd_accg_idx[0], c->dd_accg_rec, c->S, c->sc, c->sink_spin);
  return SPH_OK;
}

// ---- host-facing downloads under the decomposition: all ranks' rows, ascending number -------------------------------------
__global__ void k_rank_of_ids(int n, const int* __restrict__ id, const int* __restrict__ rank, int* __restrict__ pos) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) pos[i] = rank[id[i]];
}
__global__ void k_mark_ids(int n, const int* __restrict__ id, int* __restrict__ present) {
  int i = blockIdx.x * blockDim.x + threadIdx.x; if (i < n) present[id[i]] = 1;
}
int dd_prepare_download(sph_ctx* c) {
  const int T = 256, R = c->n_ranks;
  DDInfo mine; std::memset(&mine, 0, sizeof(mine));
  if ((int)c->dd_info.size() == R) mine = c->dd_info[c->rank];
  mine.n_own = c->n; mine.cur = c->cur; mine.key_slot = c->key[0] == c->key_alloc[0] ? 0 : 1;
  c->dd_info.resize(R);
  std::vector<DDInfo> all(R);
  { int r_ = dd_allgather_host(c, &mine, sizeof(DDInfo), all.data()); if (r_) return r_; }
  c->dd_info = all;
  long long ng = 0; for (int q = 0; q < R; ++q) ng += all[q].n_own;
  c->n_global = ng;
  if (c->p2p_stale) { int r_ = p2p_setup(c); if (r_) return r_; if (!c->p2p_ok) { c->err = "domain decomposition needs peer-mapped device memory"; return SPH_ERR_COMM; } }
  if ((size_t)std::max<long long>(ng, c->n_upload) + 1 > c->dd_g_cap) {
    c->dd_g_cap = (size_t)std::max<long long>(ng, c->n_upload) + 1024;
    DA(c->dd_gid, c->dd_g_cap); DA(c->dd_gpos, c->dd_g_cap); DA(c->dd_gstage, 2 * c->dd_g_cap); DA(c->dd_gcnt, c->dd_g_cap + 1); DA(c->dd_goff, c->dd_g_cap + 1);
    size_t b1 = 0; cub::DeviceScan::ExclusiveSum(nullptr, b1, c->dd_gcnt, c->dd_goff, (int)c->dd_g_cap + 1, c->stream);
    if (b1 + 256 > c->cub_bytes) { c->cub_bytes = b1 + 256; if (c->cub_tmp) cudaFree(c->cub_tmp); c->cub_tmp = nullptr; if (cudaMalloc(&c->cub_tmp, c->cub_bytes) != cudaSuccess) { c->err = "cudaMalloc(cub temp)"; return SPH_ERR_OOM; } }
  }
  { int r_ = dd_gather(c, DS_ID, 1, sizeof(int), c->dd_gid); if (r_) return r_; }
  const int nu = (int)c->n_upload, n = (int)ng;
  CK(cudaMemsetAsync(c->dd_gcnt, 0, sizeof(int) * (size_t)(nu + 1), c->stream));
  LAUNCH(k_mark_ids, cdiv(n, T), T, 0, n, c->dd_gid, c->dd_gcnt);
  size_t bytes = c->cub_bytes;
  CK(cub::DeviceScan::ExclusiveSum(c->cub_tmp, bytes, c->dd_gcnt, c->dd_goff, nu + 1, c->stream));
  LAUNCH(k_rank_of_ids, cdiv(n, T), T, 0, n, c->dd_gid, c->dd_goff, c->dd_gpos);
  // the same for the particles this rank holds (own + halo): their ascending-number positions (neighbour diagnostics)
  if (!c->tree_valid) { c->n_halo = 0; c->ng_halo = 0; }      // a compaction voided the halo region
  const int nl = (int)c->n + c->n_halo;
  LAUNCH(k_rank_of_ids, cdiv(nl, T), T, 0, nl, c->id[c->cur], c->dd_goff, c->pos);
  return SPH_OK;
}
// gather one field of every rank and scatter it into ascending-number order, then to the host
int dd_fetch_ordered(sph_ctx* c, int slot, int per_cur, double* dst_host) {
  if (!dst_host) return SPH_OK;
  const int n = (int)c->n_global, T = 256;
  { int r_ = dd_gather(c, slot, per_cur, 8, c->dd_gstage); if (r_) return r_; }
  LAUNCH(k_scatter_d, cdiv(n, T), T, 0, n, c->dd_gpos, c->dd_gstage, c->dd_gstage + c->dd_g_cap);
  CK(cudaMemcpyAsync(dst_host, c->dd_gstage + c->dd_g_cap, (size_t)n * 8, cudaMemcpyDeviceToHost, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  return SPH_OK;
}
