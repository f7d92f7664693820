// sph_hostcomm.h — a tiny host-side communicator over POSIX shared memory.
//
// The production collective backend is NCCL (sph_comm_init).  NCCL refuses two ranks on one device, so the
// multi-rank logic could only ever run where >= 2 GPUs are visible.  This backend carries the same few small
// collectives (all-reduce of a handful of scalars, all-gather of a few KB, barrier) through a shared-memory
// segment on the host, for ranks that are threads of one process or processes on one node - on ANY number of
// devices, including several ranks sharing one GPU ("virtual ranks": the multi-rank tests then run on a 1-GPU
// box).  Bulk data never goes through it: slices, halos and tree nodes move over peer-mapped device memory
// exactly as with NCCL.  Every collective here synchronises the calling stream first: it is a test / bring-up
// path, not a fast one.
#pragma once
#include <atomic>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <sys/stat.h>
#include <unistd.h>
#include <time.h>

#define SPH_HC_SLOT (4u << 20)          // bytes per rank of exchange space
#define SPH_HC_MAGIC 0x53504842u        // "SPHB"

struct HostCommHeader {
  std::atomic<uint32_t> magic;
  std::atomic<int> arrived;
  std::atomic<int> generation;
  std::atomic<int> attached;
  int n_ranks;
  int pad[11];
};

struct HostComm {
  int rank = 0, n = 1;
  void* base = nullptr; size_t bytes = 0; std::string name;
  HostCommHeader* hdr() const { return reinterpret_cast<HostCommHeader*>(base); }
  char* slot(int r) const { return reinterpret_cast<char*>(base) + 4096 + (size_t)r * SPH_HC_SLOT; }

  static double now() { timespec ts; clock_gettime(CLOCK_MONOTONIC, &ts); return ts.tv_sec + 1e-9 * ts.tv_nsec; }

  // rank 0 creates the segment, the others attach (retrying until it exists); returns "" on success
  std::string open(const char* nm, int rank_, int n_, double timeout_s = 120.0) {
    rank = rank_; n = n_; name = nm;
    bytes = 4096 + (size_t)n * SPH_HC_SLOT;
    int fd = -1;
    const double t0 = now();
    if (rank == 0) {
      shm_unlink(nm);
      fd = shm_open(nm, O_CREAT | O_EXCL | O_RDWR, 0600);
      if (fd < 0) return std::string("shm_open(create) failed for ") + nm;
      if (ftruncate(fd, (off_t)bytes) != 0) { ::close(fd); return "ftruncate failed"; }
    } else {
      for (;;) {
        fd = shm_open(nm, O_RDWR, 0600);
        if (fd >= 0) { struct stat st; if (fstat(fd, &st) == 0 && (size_t)st.st_size >= bytes) break; ::close(fd); fd = -1; }
        if (now() - t0 > timeout_s) return std::string("timed out waiting for shared segment ") + nm;
        usleep(1000);
      }
    }
    base = mmap(nullptr, bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
    ::close(fd);
    if (base == MAP_FAILED) { base = nullptr; return "mmap failed"; }
    HostCommHeader* h = hdr();
    if (rank == 0) {
      h->arrived.store(0); h->generation.store(0); h->attached.store(0); h->n_ranks = n;
      h->magic.store(SPH_HC_MAGIC, std::memory_order_release);
    } else {
      while (h->magic.load(std::memory_order_acquire) != SPH_HC_MAGIC) {
        if (now() - t0 > timeout_s) return "timed out waiting for the segment header";
        usleep(200);
      }
      if (h->n_ranks != n) return "rank count mismatch in shared segment";
    }
    h->attached.fetch_add(1);
    while (h->attached.load() < n) { if (now() - t0 > timeout_s) return "timed out waiting for all ranks to attach"; usleep(200); }
    if (!barrier(timeout_s)) return "barrier timed out";
    if (rank == 0) shm_unlink(nm);            // everyone holds a mapping: the name is no longer needed
    return "";
  }
  void close() { if (base) { munmap(base, bytes); base = nullptr; } }

  // generation barrier; false on timeout (a peer died): callers turn that into SPH_ERR_COMM instead of hanging
  bool barrier(double timeout_s = 300.0) {
    HostCommHeader* h = hdr();
    const int gen = h->generation.load(std::memory_order_acquire);
    if (h->arrived.fetch_add(1, std::memory_order_acq_rel) == n - 1) {
      h->arrived.store(0, std::memory_order_relaxed);
      h->generation.store(gen + 1, std::memory_order_release);
      return true;
    }
    const double t0 = now();
    int spins = 0;
    while (h->generation.load(std::memory_order_acquire) == gen) {
      if (++spins > 2000) { sched_yield(); if ((spins & 1023) == 0 && now() - t0 > timeout_s) return false; }
    }
    return true;
  }
  // all[r * nbytes ...] = rank r's `mine`
  bool allgather(const void* mine, size_t nbytes, void* all) {
    if (nbytes > SPH_HC_SLOT) return false;
    std::memcpy(slot(rank), mine, nbytes);
    if (!barrier()) return false;
    for (int r = 0; r < n; ++r) std::memcpy((char*)all + (size_t)r * nbytes, slot(r), nbytes);
    return barrier();                          // nobody overwrites a slot before everyone has read it
  }
};
