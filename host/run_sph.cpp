// run_sph.cpp — C++ twin of the reference's `program run_sph` + `simulate` shell
// (SUMMER_SPH.f90:863-955 | "SUMMER_SPH - Variable.f90":1076-1191), driving the CUDA engine through the
// C-ABI of include/sph_b200.h.  It exists because this image has no Fortran compiler; the Fortran host
// host/run_sph_b200.f90 binds the same entry points with ISO_C_BINDING.
//
//   run_sph [--variable] [--params parameters.txt] [--end-time T] [--max-steps N] [--save-dir DIR] [--drift] ics.txt
//
// --drift adds what the reference never prints: the engine's conserved sums (sph_conserved) before the first and
// after the last step and their relative drift (energy, momentum, angular momentum, mass).
//
// Surface kept from the reference: header + whitespace rows, 8 columns read in fixed-h mode (alpha := 0),
// 10 in variable-h mode, u == 0 marks a sink, dummy sink if none, save<k>.txt cadence t > k*end_time/1000,
// per-step "SPH Particles: N dt : dt time : t" line, loop ends at the first step with t >= end_time.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <cmath>
#include <algorithm>
#include <string>
#include <vector>
#include <fstream>
#include <sstream>

#include "../include/sph_b200.h"
#include "../include/sph_textio.h"

struct Table { std::vector<double> c[10]; std::vector<double> s[8]; };

static bool read_data_from_file(const std::string& fn, const sph_params& p, Table& t) {    // F:594-716
  // host-parallel parser (include/sph_textio.h): header skipped, 8 | 10 columns, u == 0 rows are sinks, dummy sink if none
  sph_ics* ics = nullptr;
  if (sph_ics_open(fn.c_str(), (p.mode & SPH_MODE_VARIABLE_H) ? 1 : 0, p.h_fixed, p.sink_radius, 0, &ics)) {
    std::fprintf(stderr, " %s\n", sph_textio_last_error());
    return false;
  }
  int64_t n = 0; int32_t ns = 0;
  sph_ics_sizes(ics, &n, &ns);
  for (auto& v : t.c) v.resize((size_t)n);
  for (auto& v : t.s) v.resize((size_t)ns);
  sph_ics_fetch(ics, t.c[0].data(), t.c[1].data(), t.c[2].data(), t.c[3].data(), t.c[4].data(), t.c[5].data(), t.c[6].data(), t.c[7].data(),
                t.c[8].data(), t.c[9].data(), t.s[0].data(), t.s[1].data(), t.s[2].data(), t.s[3].data(), t.s[4].data(), t.s[5].data(),
                t.s[6].data(), t.s[7].data());
  sph_ics_close(ics);
  std::printf(" Successfully read %lld bodies and %d sinks from %s.\n", (long long)n, (int)ns, fn.c_str());   // F:714
  return true;
}

static bool read_params_from_file(const std::string& fn, sph_params& p) {                  // V:854-919
  std::ifstream f(fn);
  if (!f) { std::fprintf(stderr, " Error opening file: %s\n", fn.c_str()); return false; }
  std::string line; std::getline(f, line);
  bool any = false;
  while (std::getline(f, line)) {
    std::istringstream is(line);
    double b, th, g, e, cv, ml, ts, et; double md;
    if (!(is >> b >> md >> th >> g >> e >> cv >> ml >> ts >> et)) break;
    p.bounding_size = b; p.max_depth = (int)md; p.theta = th; p.gamma = g; p.eta = e;
    p.convergence_criteria = cv; p.max_length = ml; p.timestep_scale = ts; p.end_time = et;
    any = true;
  }
  if (any) std::printf(" Successfully read parameters from%s.\n", fn.c_str());
  return any;
}

static bool make_save(sph_ctx* ctx, const sph_params& p, int number, const std::string& dir) {   // F:719-738
  int64_t n; int32_t ns; sph_sizes(ctx, &n, &ns);
  std::vector<double> c[10], s[8];
  for (auto& v : c) v.resize(n);
  for (auto& v : s) v.resize(ns);
  if (sph_download(ctx, c[0].data(), c[1].data(), c[2].data(), c[3].data(), c[4].data(), c[5].data(), c[6].data(), c[7].data(),
                   c[8].data(), c[9].data(), s[0].data(), s[1].data(), s[2].data(), s[3].data(), s[4].data(), s[5].data(),
                   s[6].data(), s[7].data())) return false;
  const std::string fn = dir + "/save" + std::to_string(number) + ".txt";
  // an existing file is an error like status="new" (F:728); rows formatted and written by all host cores
  if (sph_save_write(fn.c_str(), (p.mode & SPH_MODE_VARIABLE_H) ? 1 : 0, n, c[0].data(), c[1].data(), c[2].data(), c[3].data(), c[4].data(),
                     c[5].data(), c[6].data(), c[7].data(), c[8].data(), c[9].data(), ns, s[0].data(), s[1].data(), s[2].data(), s[3].data(),
                     s[4].data(), s[5].data(), s[6].data(), 0)) {
    std::fprintf(stderr, "%s\n", sph_textio_last_error());
    return false;
  }
  return true;
}

// relative drifts between two sph_conserved() results (same scales as summersph_b200/_abi.py:drift_report)
static void print_drift(const double* a, const double* b, long steps) {
  const double e0 = a[0] + a[1] + a[2], e1 = b[0] + b[1] + b[2];
  auto nz = [](double v) { return v != 0.0 ? v : 1.0; };
  const double dp = std::sqrt((b[3] - a[3]) * (b[3] - a[3]) + (b[4] - a[4]) * (b[4] - a[4]) + (b[5] - a[5]) * (b[5] - a[5]));
  const double dl = std::sqrt((b[6] - a[6]) * (b[6] - a[6]) + (b[7] - a[7]) * (b[7] - a[7]) + (b[8] - a[8]) * (b[8] - a[8]));
  const double l0 = std::max(std::sqrt(a[6] * a[6] + a[7] * a[7] + a[8] * a[8]), std::sqrt(b[6] * b[6] + b[7] * b[7] + b[8] * b[8]));
  std::printf(" Conserved sums: E = %.17g -> %.17g  (kin %.17g int %.17g pot %.17g)\n", e0, e1, b[0], b[1], b[2]);
  std::printf(" Drift over %ld steps: dE/|E0| = %.3e  |dP|/sqrt(2 E_kin M) = %.3e  |dL|/|L| = %.3e  dM/M0 = %.3e\n", steps,
              (e1 - e0) / nz(std::fabs(e0)), dp / nz(std::sqrt(2.0 * std::max(std::fabs(a[0]), std::fabs(b[0])) * std::max(a[9], b[9]))), dl / nz(l0), (b[9] - a[9]) / nz(a[9]));
}

int main(int argc, char** argv) {
  bool drift = false;
  int mode = SPH_MODE_FIXED_H; std::string params_file, save_dir, ics; double end_override = -1; long max_steps = -1;
  for (int i = 1; i < argc; ++i) {
    std::string a = argv[i];
    if (a == "--variable") mode = SPH_MODE_VARIABLE_H;
    else if (a == "--params" && i + 1 < argc) { params_file = argv[++i]; mode = SPH_MODE_VARIABLE_H; }
    else if (a == "--end-time" && i + 1 < argc) end_override = std::atof(argv[++i]);
    else if (a == "--max-steps" && i + 1 < argc) max_steps = std::atol(argv[++i]);
    else if (a == "--save-dir" && i + 1 < argc) save_dir = argv[++i];
    else if (a == "--drift") drift = true;
    else ics = a;
  }
  if (ics.empty()) { std::fprintf(stderr, "usage: run_sph [--variable] [--params parameters.txt] [--end-time T] [--max-steps N] [--save-dir DIR] [--drift] ics.txt\n"); return 2; }
  sph_params p; sph_default_params(mode, &p);
  if (!params_file.empty() && !read_params_from_file(params_file, p)) return 1;
  if (end_override >= 0) p.end_time = end_override;
  Table t;
  if (!read_data_from_file(ics, p, t)) return 1;
  sph_ctx* ctx = nullptr;
  if (sph_create(&p, 0, &ctx)) { std::fprintf(stderr, "sph_create: %s\n", sph_last_error(nullptr)); return 1; }
  if (sph_upload(ctx, (int64_t)t.c[0].size(), t.c[0].data(), t.c[1].data(), t.c[2].data(), t.c[3].data(), t.c[4].data(), t.c[5].data(),
                 t.c[6].data(), t.c[7].data(), t.c[8].data(), t.c[9].data(), (int32_t)t.s[0].size(), t.s[0].data(), t.s[1].data(),
                 t.s[2].data(), t.s[3].data(), t.s[4].data(), t.s[5].data(), t.s[6].data(), t.s[7].data())) {
    std::fprintf(stderr, "sph_upload: %s\n", sph_last_error(ctx)); return 1;
  }
  double cons0[SPH_CONSERVED_COUNT] = {}, cons1[SPH_CONSERVED_COUNT] = {};
  if (drift && sph_conserved(ctx, cons0, SPH_CONSERVED_COUNT)) { std::fprintf(stderr, "sph_conserved: %s\n", sph_last_error(ctx)); return 1; }
  double tt = 0.0, dt = 1.0e-2;                                                            // F:872,875
  int t_test = 0; long steps = 0; int64_t n = (int64_t)t.c[0].size(); int32_t ns = 0;
  while (tt < p.end_time) {                                                                // F:879
    if (!save_dir.empty() && tt > t_test * p.end_time / 1000.0) {                          // F:881
      if (!make_save(ctx, p, t_test, save_dir)) return 1;
      ++t_test;
    }
    std::printf(" SPH Particles: %lld dt :   %.17g time :    %.17g\n", (long long)n, dt, tt);   // F:891
    std::fflush(stdout);
    if (sph_step(ctx, &dt, &tt, &n, &ns)) { std::fprintf(stderr, "sph_step: %s\n", sph_last_error(ctx)); return 1; }
    ++steps;
    if (max_steps >= 0 && steps >= max_steps) break;
  }
  if (drift) {
    if (sph_conserved(ctx, cons1, SPH_CONSERVED_COUNT)) { std::fprintf(stderr, "sph_conserved: %s\n", sph_last_error(ctx)); return 1; }
    print_drift(cons0, cons1, steps);
  }
  sph_destroy(ctx);
  return 0;
}
