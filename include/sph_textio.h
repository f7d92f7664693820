/*
 * sph_textio.h — C-ABI of the host-parallel text reader / writer for the reference's file formats
 * (at 16M-64M particles the IC / save text files are 3-13 GB and the
 * reference's list-directed Fortran I/O, or a scalar parser, dominates the wall time of a run).
 *
 * Formats:
 *   IC file   : line 1 is a header and is skipped (SUMMER_SPH.f90:617,645); then one body per row,
 *               whitespace (or comma) separated reals.  Fixed h reads the first 8 values
 *               `x y z vx vy vz u m`, ignores the rest and sets alpha := 0 (F:647,681); variable h reads
 *               10: `... alpha h` ("SUMMER_SPH - Variable.f90":782).  u == 0.0 exactly marks a sink row
 *               (F:658-659; x, v, m taken, radius := sink_radius F:694 | V:830).  No sink row => one dummy
 *               sink of zeros (F:698-707).  Blank rows are skipped (list-directed reads skip empty records).
 *   save file : `save<k>.txt` (F:719-738 | V:921-942): header, gas rows of 9 (fixed h) or 10 (variable h)
 *               reals, then sink rows `x y z vx vy vz 0.0 m`; 17 significant digits (lossless), opened
 *               like status="new": an existing file is an error (F:728).
 *
 * Plain pointers and sizes, no exceptions across the boundary; every function returns 0 or SPH_TEXTIO_ERR_*;
 * sph_textio_last_error() gives the text (thread-local).  Pure host code: no CUDA, no GPU needed.
 */
#ifndef SPH_TEXTIO_H
#define SPH_TEXTIO_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SPH_TEXTIO_OK          0
#define SPH_TEXTIO_ERR_ARG    -1
#define SPH_TEXTIO_ERR_OPEN   -2   /* "Error opening file" F:612-615                       */
#define SPH_TEXTIO_ERR_EMPTY  -3   /* "No data found in file" F:625-628                    */
#define SPH_TEXTIO_ERR_PARSE  -4   /* "Error reading line N" F:648-651 (N = data row)      */
#define SPH_TEXTIO_ERR_EXISTS -5   /* save file already exists (status="new", F:728)       */
#define SPH_TEXTIO_ERR_WRITE  -6

typedef struct sph_ics sph_ics;    /* a parsed IC file held in host memory */

const char* sph_textio_last_error(void);

/* read_data_from_file (F:594-716 | V:729-852): parse `path` with `threads` host threads (<= 0: all cores).
 * variable_h != 0 reads 10 columns per gas row, else 8 (alpha := 0, h := h_fixed). */
int sph_ics_open(const char* path, int32_t variable_h, double h_fixed, double sink_radius, int32_t threads, sph_ics** out);
/* n_sink >= 1 (the dummy sink counts, F:698-707) */
int sph_ics_sizes(const sph_ics* ics, int64_t* n_gas, int32_t* n_sink);
/* copy the columns out: gas = x y z vx vy vz u m alpha h (file order of the gas rows = `number` order, F:684);
 * sinks = x y z vx vy vz m radius.  Any pointer may be NULL. */
int sph_ics_fetch(const sph_ics* ics,
                  double* x, double* y, double* z, double* vx, double* vy, double* vz, double* u, double* m, double* alpha, double* h,
                  double* sx, double* sy, double* sz, double* svx, double* svy, double* svz, double* sm, double* sradius);
int sph_ics_close(sph_ics* ics);

/* make_save (F:719-738 | V:921-942): fixed-width rows (25 characters per value + one blank, 17 significant digits) formatted by
 * `threads` host threads and written at their offsets.  n_sink rows are written as given (pass 0 to write none). */
int sph_save_write(const char* path, int32_t variable_h, int64_t n_gas,
                   const double* x, const double* y, const double* z, const double* vx, const double* vy, const double* vz,
                   const double* u, const double* m, const double* alpha, const double* h,
                   int32_t n_sink, const double* sx, const double* sy, const double* sz, const double* svx, const double* svy,
                   const double* svz, const double* sm, int32_t threads);

#ifdef __cplusplus
}
#endif
#endif
