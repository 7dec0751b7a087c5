/*
 * smrf_b200.h -- C ABI of libsmrf_b200.so: the B200 (sm_100a) implementation of
 * the SMRF ground-classification hot path of thomaspingel/neilpy.
 *
 * The reference has no FFI of its own: its boundary for this path is the Python
 * call surface re-exported at neilpy/__init__.py:1
 *     smrf                     neilpy/neilpy.py:1685-1808
 *     create_dem               neilpy/neilpy.py:1110-1166
 *     inpaint_nans_by_springs  neilpy/neilpy.py:1227-1271
 *     progressive_filter       neilpy/neilpy.py:1659-1680
 * Each entry point below replaces the arithmetic of the reference lines it
 * cites; the Python mirror (neilpy_b200/api.py) keeps the reference signatures
 * and binds these symbols through ctypes (see INTEGRATION.md for the stub).
 *
 * Conventions
 *   - plain pointers and sizes only; every pointer is a DEVICE pointer unless the
 *     parameter name ends in _host.  The caller owns every buffer (inputs,
 *     outputs, workspaces); the library never allocates, frees or retains.
 *   - `stream` is a cudaStream_t passed as void* (0 = legacy default stream).
 *     Calls are asynchronous on that stream except where stated.
 *   - return value: 0 = ok, < 0 = argument error (SMRF_E_*), > 0 = cudaError_t.
 *     smrf_last_error() returns a per-thread description of the last failure.
 *   - `dtype` selects the grid element type: SMRF_F32 (throughput mode) or
 *     SMRF_F64 (parity mode, the reference's own float64).
 *   - grids are row-major [ny][nx], row 0 = north (as the reference's I[r, c]).
 *   - masks are uint8 (0/1), one byte per cell or point (numpy bool layout).
 */
#ifndef SMRF_B200_H
#define SMRF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SMRF_F32 0
#define SMRF_F64 1

/* point stream layouts */
#define SMRF_PTS_SOA_F64 0 /* three separate double arrays x, y, z (the reference's own) */
#define SMRF_PTS_XYZW_F32 1 /* one interleaved float4 array (x, y, z, unused); `x` is its base, y = z = NULL */
#define SMRF_PTS_SOA_F32 2 /* three separate float arrays */

#define SMRF_BIN_MIN 0
#define SMRF_BIN_MAX 1

#define SMRF_E_ARG (-1)      /* null pointer / bad enum / bad size */
#define SMRF_E_WORKSPACE (-2) /* workspace too small */
#define SMRF_E_UNSUPPORTED (-3)

int smrf_abi_version(void);
const char* smrf_last_error(void);
/* number of CUDA kernels this process has launched through the library (bench: gpu_launches) */
unsigned long long smrf_launch_count(void);
/* name of the opening implementation a call with these parameters would use
 * ("march_f32_w<=18", "tile_generic", ...) -- for logs and the bench JSON. */
const char* smrf_open_variant(int dtype, int window);

/* ---- create_dem: extent -------------------------------------- neilpy.py:1117-1124
 * min/max of x and y (np.min / np.max).  out4 = {min x, max x, min y, max y} as
 * doubles, nonfinite[0] = number of non-finite x or y (the reference would fail on
 * them in np.arange / np.ravel_multi_index).  `scratch` holds 4 int64 keys. */
int smrf_extent(const void* x, const void* y, int64_t n, int point_fmt,
                double* out4, int64_t* nonfinite, int64_t* scratch4, void* stream);

/* ---- create_dem: binning ------------------------------------- neilpy.py:1136-1161
 * Replaces ~t*(x,y) -> floor -> ravel_multi_index -> groupby().min()/max() -> scatter.
 * inv6 = {ra, rb, rc, rd, re, rf}: the inverse affine exactly as the `affine`
 * package computes it on the host (col = x*ra + y*rb + rc, row = x*rd + y*re + rf,
 * evaluated left to right in float64 without FMA contraction).
 *   smrf_bin_init       : fill `grid` (ny*nx elements of `dtype`) with the empty key
 *   smrf_bin_accumulate : atomic min/max of order-preserving keys; may be called
 *                         repeatedly (chunks / several point sources).  NaN z is
 *                         skipped (pandas skips NaN).  Points that fall outside the
 *                         grid are counted in out_of_range[0] (the reference raises
 *                         ValueError from np.ravel_multi_index).
 *   smrf_bin_finalize   : keys -> values in place, untouched cells -> NaN,
 *                         empty[i] = 1 for them (neilpy.py:1742 is_empty_cell). */
int smrf_bin_init(void* grid, int64_t ny, int64_t nx, int dtype, int bin_type, void* stream);
int smrf_bin_accumulate(const void* x, const void* y, const void* z, int64_t n, int point_fmt,
                        const double* inv6_host, void* grid, int64_t ny, int64_t nx, int dtype,
                        int bin_type, int64_t* out_of_range, void* stream);
int smrf_bin_finalize(void* grid, uint8_t* empty, int64_t ny, int64_t nx, int dtype, int bin_type,
                      void* stream);
/* Row-band sharding: every rank bins its own arbitrary slice of the points
 * into a full-grid replica; smrf_bin_finalize_partial decodes the keys leaving +inf (min) /
 * -inf (max) in untouched cells, so that an elementwise MIN / MAX reduce-scatter over the
 * ranks (NCCL) yields each rank's row band of the global binning; smrf_bin_mark_empty then
 * turns the cells that are still +-inf into NaN and writes the empty mask. */
int smrf_bin_finalize_partial(void* grid, int64_t ny, int64_t nx, int dtype, int bin_type, void* stream);
int smrf_bin_mark_empty(void* grid, uint8_t* empty, int64_t ny, int64_t nx, int dtype, int bin_type,
                        void* stream);

/* ---- row-band sharding: every point travels once to the rank that owns its row band.
 * With the points of a band resident on its rank, binning (neilpy.py:1142-1161) and interpolation +
 * classification (neilpy.py:1772-1795) are band-local.  The owner of a point is the band of floor(row), the
 * row computed with the binning's own arithmetic, bands being `rows_per_band` rows each.
 *   smrf_route_plan   : dest[i] = owner band (world = not in the grid / non-finite), counts[0..world] += 1
 *                       (device int64[world + 1], zeroed here)
 *   smrf_route_pack   : writes the points grouped by destination into the send buffer -- one float4 stream
 *                       (out_xyzw; float32 inputs) or three float64 columns -- and perm[i] = slot of point i;
 *                       cursors[d] (device int64[world + 1]) must hold the first slot of every destination
 *                       and is advanced
 *   smrf_route_unpack : out[i] = back[perm[i]] (the classification that came back, in the caller's order)
 *   smrf_bin_accumulate_band / smrf_classify_band : the band forms of smrf_bin_accumulate / smrf_classify:
 *                       `band` / `coef_band` hold rows [row0, row0 + rows) of the ny x nx grid; a point of
 *                       another band counts as out of range (binning) / must not occur (classify: the 4 x 4
 *                       taps of a point in cell row r span rows r-2 .. r+2 of the interleaved coefficients). */
int smrf_route_plan(const void* x, const void* y, int64_t n, int point_fmt, const double* inv6_host, int64_t ny,
                    int64_t nx, int64_t rows_per_band, int world, uint8_t* dest, int64_t* counts, void* stream);
int smrf_route_pack(const void* x, const void* y, const void* z, int64_t n, int point_fmt, int world,
                    const uint8_t* dest, int64_t* cursors, void* out_xyzw, double* out_x, double* out_y,
                    double* out_z, int64_t* perm, void* stream);
int smrf_route_unpack(const uint8_t* back, const int64_t* perm, int64_t n, uint8_t* out, void* stream);
int smrf_bin_accumulate_band(const void* x, const void* y, const void* z, int64_t n, int point_fmt,
                             const double* inv6_host, void* band, int64_t ny, int64_t nx, int64_t row0,
                             int64_t rows, int dtype, int bin_type, int64_t* out_of_range, void* stream);
int smrf_classify_band(const void* x, const void* y, const void* z, int64_t n, int point_fmt,
                       const double* inv6_host, const void* coef_band, int64_t ny, int64_t nx, int64_t row0,
                       int64_t rows, int dtype, double elevation_threshold, double elevation_scaler,
                       uint8_t* is_object, void* stream);

/* ---- inpaint_nans_by_springs --------------------------------- neilpy.py:1227-1271
 * Discrete harmonic fill of the NaN cells of `grid` (in place): for every NaN cell
 * deg*u - sum(NaN nbrs u) = sum(known nbrs a), deg = number of in-grid 4-neighbours
 * (the normal equations of the reference's spring system).  Solved in float64 by
 * multigrid-preconditioned conjugate gradients entirely in HBM; stops when the
 * max-norm of the residual is <= tol (metres) or after max_iter iterations.
 * A grid with no NaN is returned unchanged; an all-NaN grid becomes zeros (the
 * minimum-norm answer LSQR gives).  `unknown` (optional, may be NULL) receives the
 * NaN mask.  `guess` (optional, ny*nx elements of `dtype`) seeds the NaN cells; without it
 * they start from the mean of the known cells (the answer does not depend on it, the
 * iteration count does: smrf() passes the last opened surface).  info_host[0] = iterations, info_host[1] = final residual max-norm,
 * info_host[2] = number of unknown cells.  Synchronises `stream`. */
size_t smrf_inpaint_workspace_bytes(int64_t ny, int64_t nx);
int smrf_inpaint(void* grid, int64_t ny, int64_t nx, int dtype, uint8_t* unknown, const void* guess,
                 void* workspace, size_t workspace_bytes, double tol, int max_iter,
                 double* info_host, void* stream);

/* ---- inpaint_nans_by_fda ------------------------------------- neilpy.py:1171-1216
 * The "finite difference approximation" fill: one equation per cell whose stencil touches a NaN,
 *   V (u[up] + u[down] - 2u) + H (u[left] + u[right] - 2u) = 0   (V = 1 on rows 1..ny-2, H = 1 on columns 1..nx-2),
 * solved for the NaN cells in the least-squares sense (the reference calls LSQR), every equation weighted by the number
 * of NaN cells its stencil touches (the reference's row selection repeats it that often).  Here: CGLS in float64, started
 * from zero like LSQR; stops when max |A^T r| <= tol or after max_iter iterations.  NaN cells of `grid` are
 * overwritten.  info_host = {iterations, final max |A^T r|, number of NaN cells}.  Synchronises `stream`. */
size_t smrf_inpaint_fda_workspace_bytes(int64_t ny, int64_t nx);
int smrf_inpaint_fda(void* grid, int64_t ny, int64_t nx, int dtype, void* workspace, size_t workspace_bytes,
                     double tol, int max_iter, double* info_host, void* stream);

/* The same solver, one phase at a time, for a row band whose neighbours live on other
 * ranks (the caller all-reduces the dot-product slots and exchanges one boundary row of u
 * (once) and of p (every iteration) between the calls -- neilpy_b200/distributed.py):
 *   setup  : NaN mask + multigrid hierarchy of the band; statistics of the known cells
 *   start  : phase 0: u = known value, else guess_grid (if given), else `guess`;  phase 1: r = b - A u with the
 *            neighbours' boundary rows of u (u_above / u_below, nx doubles each)
 *   step   : phase 0: z = M^-1 r (band-local V-cycle), rz[k] += r.z   (phase 20: rz only, z = z_ext)
 *            phase 1: p = z + (rz[k]/rz[k-1]) p
 *            phase 2: q = A p with the neighbours' boundary rows of p and of the NaN mask,
 *                     pq[k] += p.q
 *            phase 3: u += alpha p, r -= alpha q, rmax[k+1] = max |r|
 *   finish : write the solution into the NaN cells of `grid`
 * has_above / has_below say whether a neighbouring band exists (its cells count in the
 * degree; inside the preconditioner they carry no correction: block Jacobi over bands).
 * smrf_inpaint_layout reports byte offsets into the workspace: {u plane, p plane, NaN
 * mask, rz[], pq[], rmax[] (bit patterns of non-negative doubles), statistics block
 * {sum_known f64, n_known u64, n_unknown u64}, slots per array}. */
int smrf_inpaint_layout(int64_t ny, int64_t nx, int64_t* out8_host);
int smrf_inpaint_setup(const void* grid, int64_t ny, int64_t nx, int dtype, void* workspace,
                       size_t workspace_bytes, int has_above, int has_below, void* stream);
int smrf_inpaint_start(const void* grid, int64_t ny, int64_t nx, int dtype, void* workspace,
                       size_t workspace_bytes, int has_above, int has_below, double guess,
                       const void* guess_grid, int phase, const double* u_above, const double* u_below,
                       void* stream);
int smrf_inpaint_step(int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, int has_above,
                      int has_below, int k, int phase, const float* z_ext, const double* p_above,
                      const double* p_below, const uint8_t* m_above, const uint8_t* m_below, void* stream);
/* The row-band solver with COMPACT CG vectors (u, r, p, q hold the band's NaN cells only, as smrf_inpaint does on
 * one GPU); replaces smrf_inpaint_start / _step phases 1-3 / _finish after smrf_inpaint_setup.  op:
 *   0  index maps + zeroed float32 residual plane        1  starting guess (`guess`, `guess_grid`)
 *   2  out_first / out_last = boundary rows of u (known elevations included) for the neighbours
 *   3  r = b - A u with row_above / row_below (the neighbours' rows of u), rmax[0]
 *   4  p = z + (rz[k]/rz[k-1]) p, then out_first / out_last = boundary rows of p (0 on known cells)
 *   5  q = A p with row_above / row_below (the neighbours' rows of p), pq[k] += p.q
 *   6  u += alpha p, r -= alpha q, float32 residual plane (level-0 right-hand side of the cycle), rmax[k+1]
 *   7  write the solution into the NaN cells of `grid`
 * n_unknown = this band's NaN cells (statistics block of smrf_inpaint_setup, before any all-reduce).
 * r_plane (ops 0, 3, 6; may be NULL) = where the float32 residual plane lives instead of the workspace's own
 * level-0 right-hand side: ny x nx floats, e.g. the owned rows of the ghost-extended band smrf_mg_cycle_part reads. */
int smrf_inpaint_compact(int op, const void* grid, int64_t ny, int64_t nx, int dtype, void* workspace,
                         size_t workspace_bytes, int has_above, int has_below, int64_t n_unknown, int k, double guess,
                         const void* guess_grid, const float* z, float* r_plane, const double* row_above,
                         const double* row_below, double* out_first, double* out_last, void* stream);
/* Preconditioning a row band with the GLOBAL V-cycle.  A band-local cycle (any closure at the band
 * edge) mistreats every error mode that is smooth across the edge and costs ~60 % more CG
 * iterations; instead the caller (neilpy_b200/distributed.py)
 *   - keeps a second hierarchy for the band extended by G ghost rows per side (masks exchanged
 *     once; G >= the cycle's dependency radius of the band-local levels, 35 rows for 3 levels),
 *   - per iteration exchanges G rows of the level-0 right-hand side, runs the down legs of levels
 *     [0, split) on the extended band (smrf_mg_cycle_part, part 0), all-gathers the owned rows of
 *     the level-`split` right-hand side into a hierarchy of the global coarse grid that every
 *     rank holds (smrf_mg_setup_mask once, smrf_mg_vcycle per iteration), copies the rows of the
 *     global coarse correction that its extended band covers back, and runs the up legs (part 2);
 *   - hands the owned rows of the result to step phase 20 (rz) and phase 1 (p) as z_ext.
 * On the owned rows this equals the single-GPU V-cycle, so the iteration count does not depend
 * on the number of bands.  part 1 = the cycle from level `split` to the coarsest and back.
 * smrf_mg_level_layout: {ny_l, nx_l, byte offsets of mask, x, y (result), b (rhs)} of a level. */
int smrf_mg_level_layout(int64_t ny, int64_t nx, int level, int64_t* out6_host);
int smrf_mg_setup_mask(const uint8_t* mask, int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes,
                       void* stream);
int smrf_mg_cycle_part(int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, int has_above,
                       int has_below, int split, int part, void* stream);
int smrf_mg_vcycle(int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, void* stream);
/* part 2 (the up legs of levels split-1 .. 0) with the preconditioned product formed on the way: the level-0 leg
 * adds sum(b * z) over rows [row_lo, row_hi) of the extended band (the rows this rank owns) to *rz_slot (device
 * double, e.g. the rz[k] slot of the band's own workspace), which replaces step phase 20. */
int smrf_mg_cycle_up_rz(int64_t ny, int64_t nx, void* workspace, size_t workspace_bytes, int has_above,
                        int has_below, int split, double* rz_slot, int64_t row_lo, int64_t row_hi, void* stream);
int smrf_inpaint_finish(void* grid, int64_t ny, int64_t nx, int dtype, void* workspace,
                        size_t workspace_bytes, void* stream);

/* ---- progressive_filter -------------------------------------- neilpy.py:1659-1680
 * For each radius windows_host[i] (in order): this = opening(last, disk(w));
 * new = (last - this) > thresholds_host[i] (float64 compare); mask |= new;
 * when_dropped[new] = i (optional); last = this.  disk(w) = {dx^2+dy^2 <= w^2};
 * borders ignore out-of-image samples (== scipy.ndimage 'reflect' for a disk).
 *   surface  : ny*nx elements, the input Z; never written (the reference works on a copy)
 *   workspace: smrf_open_workspace_bytes() bytes (ping-pong surfaces + scratch)
 *   mask     : uint8 ny*nx, OR-accumulated (caller zeroes it for a fresh filter)
 *   negate   : open -Z instead of Z (the low-outlier pass, neilpy.py:1744); one window only
 *   last_out : optional ny*nx elements receiving the last window's opening (NULL to skip)
 * thresholds are computed by the caller exactly as the reference does
 * (slope_threshold * (windows * cellsize), neilpy.py:1661). */
size_t smrf_open_workspace_bytes(int64_t ny, int64_t nx, int dtype, int max_window);
int smrf_progressive_open(const void* surface, void* workspace, size_t workspace_bytes, uint8_t* mask,
                          uint8_t* when_dropped, int64_t ny, int64_t nx, int dtype,
                          const int32_t* windows_host, const double* thresholds_host, int n_windows,
                          int negate, void* last_out, void* stream);
/* one window, out of place: `out` = opening(in, disk(window)); mask/when_dropped as above
 * (either may be NULL); `tmp` = ny*nx elements.  Only rows [row_lo,row_hi) of out/mask are
 * written: with row-band sharding the halo rows outside are inputs only (pass 0, ny
 * otherwise).  `pitch` is the row stride of in / out / tmp in elements (>= nx; mask and
 * when_dropped are always nx wide): rows padded to 16 bytes let the marching kernels use
 * 16-byte copies and stores for any nx (smrf_progressive_open pads its ping-pong surfaces
 * itself).  Building block of smrf_progressive_open and of the multi-GPU driver. */
int smrf_open_window(const void* in, void* out, void* tmp, uint8_t* mask, uint8_t* when_dropped,
                     int64_t ny, int64_t nx, int64_t pitch, int dtype, int window, double threshold,
                     int window_index, int negate, int64_t row_lo, int64_t row_hi, void* stream);
/* brute-force disk opening (|disk| loads per cell); the in-library cross-check of the fast kernels */
int smrf_open_window_bruteforce(const void* in, void* out, void* tmp, int64_t ny, int64_t nx, int dtype,
                                int window, void* stream);

/* ---- mask merge + punch -------------------------------------- neilpy.py:1762-1763
 * object = empty | low | obj (written to `object_cells`); grid[object] = NaN. */
int smrf_merge_punch(void* grid, const uint8_t* empty, const uint8_t* low, const uint8_t* obj,
                     uint8_t* object_cells, int64_t ny, int64_t nx, int dtype, void* stream);

/* ---- slope ---------------------------------------------------- neilpy.py:1785-1786
 * S = sqrt(gy^2 + gx^2), gy, gx = np.gradient(Z, cellsize) (central differences,
 * one-sided at the edges).  float64 arithmetic, stored as `dtype`. */
int smrf_slope(const void* grid, void* slope, int64_t ny, int64_t nx, int dtype, double cellsize,
               void* stream);

/* ---- RectBivariateSpline(kx=ky=3, s=0) ------------------------ neilpy.py:1768-1774,1788-1790
 * B-spline coefficients of the interpolating (not-a-knot) bicubic spline through the
 * grid values at the cell centres 0.5, 1.5, ...  The banded collocation system of each
 * axis depends only on its length; the caller factors it once on the host
 * (neilpy_b200.spline.notaknot_factors) and passes, per axis, a DEVICE array of 5*n
 * doubles {l1, l2, 1/pivot, u1, u2} (banded LU without pivoting; row_factors for the
 * ny-long axis, col_factors for the nx-long one).  Intermediates are float64 in
 * `workspace`.  Coefficient (i, j) is written to coef[(i*nx + j)*coef_stride + coef_offset]:
 * stride 1 / offset 0 is a plain grid (and may alias `grid`); stride 2 with offsets 0 and 1
 * interleaves the DTM's and the slope raster's coefficients so that smrf_classify fetches both
 * splines' taps from the same sectors.  ny, nx >= 4 (FITPACK raises otherwise). */
size_t smrf_spline_workspace_bytes(int64_t ny, int64_t nx);
int smrf_spline_prefilter(const void* grid, void* coef, int64_t coef_stride, int64_t coef_offset,
                          int64_t ny, int64_t nx, int dtype, const double* row_factors,
                          const double* col_factors, void* workspace, size_t workspace_bytes, void* stream);

/* ---- interpolate + classify ----------------------------------- neilpy.py:1772-1795
 * For every point: (c, r) = ~t*(x, y); elevation = spline(Zpro).ev(r, c);
 * slope = spline(S).ev(r, c) (arguments clamped to the centre range as FITPACK's bispeu
 * does); is_object = |elevation - z| > elevation_threshold + elevation_scaler*slope.
 * coef_s == NULL: coef_z holds interleaved (DTM, slope) coefficient pairs, [ny][nx][2].
 * Optional outputs (may be NULL): elevation, slope_out (float64 per point),
 * when_dropped_pt = drop_raster[round(r), round(c)] if drop_raster != NULL. */
int smrf_classify(const void* x, const void* y, const void* z, int64_t n, int point_fmt,
                  const double* inv6_host, const void* coef_z, const void* coef_s, int64_t ny,
                  int64_t nx, int dtype, double elevation_threshold, double elevation_scaler,
                  uint8_t* is_object, double* elevation, double* slope_out,
                  const uint8_t* drop_raster, uint8_t* when_dropped_pt, void* stream);

/* ---- LAS ingest / egress (the step before and after the path) -- neilpy.py:903-1087
 * `records` is the point-data block of a LAS file as it lies in the file (packed records of
 * `record_length` bytes, little-endian, X Y Z int32 first), copied to the device at a
 * 16-byte aligned address.  smrf_las_decode writes x = X*scale[0] + offset[0] etc. as float64
 * (product rounded, then sum rounded: neilpy.py:1056-1059) and, if `classification` is not
 * NULL, the raw byte at `class_offset` of every record (the reference's 'class' column:
 * offset 15 in point formats 0-5, 16 in formats 6-10).  scale3/offset3 are HOST arrays of
 * three doubles.  record_length 12..200. */
int smrf_las_decode(const uint8_t* records, int64_t n, int record_length, const double* scale3_host,
                    const double* offset3_host, double* x, double* y, double* z, uint8_t* classification,
                    int class_offset, void* stream);
/* classification byte of record i = (old & keep_mask) | (is_object_point[i] ? object_code :
 * ground_code), in place.  The reference's laspy notebook writes 2*(1-is_object_point)
 * (examples/smrf/SMRF Classification using laspy to read and write.ipynb, cell 5): ground 2,
 * object 0; keep_mask 0xE0 preserves the flag bits that share the byte in formats 0-5, 0 in
 * formats 6-10. */
int smrf_las_write_class(uint8_t* records, int64_t n, int record_length, int class_offset, int keep_mask,
                         const uint8_t* is_object_point, int ground_code, int object_code, void* stream);

/* ---- raster products of the DTM (the step after the path) ------ neilpy.py:456-484,814-867
 * One pass over `grid` (ny x nx, `dtype`), float64 arithmetic in the reference's order:
 *   SMRF_TERRAIN_SLOPE      slope(Z, cellsize, z_factor, return_as): gradient spacing
 *                           `spacing` = cellsize / z_factor; return_as 0 percent, 1 radians,
 *                           2 degrees -> out_f64                       neilpy.py:456-467
 *   SMRF_TERRAIN_ASPECT     aspect(Z, return_as, flat_as): unit-spacing gradient, compass
 *                           bearing, return_as 1 / 2; flat cells = NaN (flat_is_nan) or
 *                           flat_value -> out_f64                      neilpy.py:471-484
 *   SMRF_TERRAIN_HILLSHADE  hillshade(Z, cellsize, z_factor, zenith, azimuth): the caller passes
 *                           cos / sin of the zenith and the azimuth in radians (np.deg2rad);
 *                           exactly one of out_u8 (return_uint8=True) / out_f64
 *                                                                      neilpy.py:814-824
 *   SMRF_TERRAIN_PSSM       pssm(Z, cellsize, ve): index round(255*deg(atan(ve*S))/90) -> out_u8
 *                           (may be NULL) and / or its colour lut_rgba[index] -> rgba
 *                           ([ny][nx][4] float64; lut_rgba = 256 x 4 float64 on the device,
 *                           matplotlib's bone_r or bone)                neilpy.py:846-867
 * Unused parameters are ignored. */
enum { SMRF_TERRAIN_SLOPE = 0, SMRF_TERRAIN_ASPECT = 1, SMRF_TERRAIN_HILLSHADE = 2, SMRF_TERRAIN_PSSM = 3 };
int smrf_terrain(const void* grid, int64_t ny, int64_t nx, int dtype, int mode, int return_as, double spacing,
                 int flat_is_nan, double flat_value, double cos_zenith, double sin_zenith, double azimuth_rad,
                 double ve, double* out_f64, uint8_t* out_u8, double* rgba, const double* lut_rgba, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SMRF_B200_H */
