/*
 * ips.h -- C ABI of the B200-native hot path of the image-processing suite.
 *
 * The reference (Saguaro-Biosciences/image-processing-suite) has no FFI: its boundary is
 * "Python function inside a CLI script".  Each entry point below replaces the arithmetic
 * of one such function; the reference site is cited as file:line relative to the
 * reference repository.  INTEGRATION.md shows the ctypes binding a maintainer of the
 * reference would add at each site.
 *
 * Conventions (all entry points):
 *   - extern "C", plain pointers and sizes, no exceptions across the boundary;
 *   - return 0 (IPS_OK) or a negative IPS_ERR_* code; ips_last_error() returns a
 *     thread-local human-readable message for the last failure on the calling thread;
 *   - device pointers unless the name says _host; row-major, contiguous; base pointers
 *     16-byte aligned; never allocate or free caller-visible memory (scratch comes in
 *     through (ws, ws_bytes), sized by the matching *_workspace_bytes query);
 *   - asynchronous on the cudaStream_t passed as `stream` (typed void* so that C callers
 *     do not need cuda_runtime.h); re-entrant and thread-safe across distinct streams.
 *   - there is no CPU fallback: without a CUDA device every compute entry point fails
 *     with IPS_ERR_CUDA.
 */
#ifndef IPS_H_
#define IPS_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define IPS_ABI_VERSION 1

enum {
  IPS_OK = 0,
  IPS_ERR_BAD_SHAPE = -1,
  IPS_ERR_BAD_DTYPE = -2,
  IPS_ERR_BAD_ALIGN = -3,
  IPS_ERR_CUDA = -4,
  IPS_ERR_NCCL = -5,
  IPS_ERR_NOMEM = -6,
  IPS_ERR_BAD_ARG = -7
};

typedef void* ips_stream_t; /* cudaStream_t */

/* ---- library ------------------------------------------------------------------- */
int ips_abi_version(void);
const char* ips_last_error(void);
/* Number of kernel launches this library has issued from the calling process so far
 * (bench.py reports the difference across its timed region as "gpu_launches"). */
uint64_t ips_launch_count(void);
/* sm_count, compute capability and L2 size of the current device. */
int ips_device_info(int* sm_count, int* cc_major, int* cc_minor, size_t* l2_bytes);

/* ---- K1: fused z-max -> illumination divide -> b x b sum-bin ---------------------
 * Replaces  np.maximum.reduce(images)                 MaxProjection.py:45
 *           img.astype(float) / illum_cache[i]        Illumination_QC_mult.py:145-150,
 *                                                      Cellpose_GPU_s3fs.py:72
 *           calculate_saturation_cp_exact (optional)  Illumination_QC_mult.py:73-95
 *           and north_star's 2x2 / 4x4 sum re-binning (no reference site).
 *
 * raw        [F][C][Z][H][W] uint16
 * illum      [C][H][W] float32, or NULL (no correction)
 * maxproj    [F][C][H][W] uint16, or NULL
 * corrected  [F][C][H][W] float32 = maxproj / illum, or NULL (requires illum)
 * binned     [F][C][H/bin][W/bin]: uint32 sums of maxproj when illum == NULL, float32
 *            sums of corrected otherwise; or NULL.  bin in {1, 2, 4}; H, W % bin == 0.
 * pct_maximal[F][C] float64 = 100 * #(x == max x) / (H*W) evaluated on the float64
 *            quotient (or on the integer max projection when illum == NULL), or NULL.
 *            Needs ws of ips_preprocess_workspace_bytes(); ws may be NULL otherwise.
 */
size_t ips_preprocess_workspace_bytes(int F, int C, int H, int W, int bin);
int ips_preprocess_fused(const uint16_t* raw, const float* illum, uint16_t* maxproj,
                         float* corrected, void* binned, int bin, double* pct_maximal,
                         void* ws, size_t ws_bytes, int F, int C, int Z, int H, int W,
                         ips_stream_t stream);

/* ---- K3: per-object statistics over a label mask -----------------------------------
 * Replaces the CellProfiler 4.2.8 MeasureObjectSizeShape / MeasureObjectIntensity
 * subprocess launched at Feature_extraction_opt.py:166-167 (schema consumers
 * Normalize_CP_ami.py:47-127, Pycyto_pertime.py:46-75) and the regionprops loop of
 * Cellpose_GPU_s3fs.py:149-170 (label / bbox / centroid conventions).
 *
 * labels     [F][H][W] int32, 0 = background, objects 1..Nmax
 * maxproj    [F][C][H][W] uint16;  illum [C][H][W] float32 or NULL
 * intensity value of a pixel = maxproj / illum * intensity_scale
 * n_objects  [F] int32: number of rows written for the field, or -1 if a label > Nmax
 *            was met (rows of that field are then unspecified)
 * ints       [F][Nmax][6] int32: label, area, y0, x0, y1, x1 (half-open box)
 * flts       [F][Nmax][2+5C] float32: centroid y, x; per channel sum, mean,
 *            population std, min, max.  Rows are in ascending label order, absent
 *            labels are skipped (regionprops convention).  1 <= C <= 8.
 */
size_t ips_object_stats_workspace_bytes(int F, int C, int Nmax);
int ips_object_stats(const int32_t* labels, const uint16_t* maxproj, const float* illum,
                     float intensity_scale, int32_t* n_objects, int32_t* ints, float* flts,
                     int Nmax, void* ws, size_t ws_bytes, int F, int C, int H, int W,
                     ips_stream_t stream);

/* ---- K1 + K3 in one pass over the field -------------------------------------------------
 * Same results as ips_preprocess_fused (maxproj, binned) followed by ips_object_stats on
 * that max projection, without writing and re-reading it: the raw z-stack, the
 * illumination function and the label mask are each read once.
 * illum may be NULL (integer mode: uint32 bin sums, exact integer intensity sums).
 * maxproj and binned may be NULL (outputs skipped).
 */
size_t ips_field_fused_workspace_bytes(int F, int C, int H, int W, int bin, int Nmax);
int ips_field_fused(const uint16_t* raw, const float* illum, const int32_t* labels,
                    uint16_t* maxproj, void* binned, int bin, float intensity_scale,
                    int32_t* n_objects, int32_t* ints, float* flts, int Nmax, void* ws,
                    size_t ws_bytes, int F, int C, int Z, int H, int W, ips_stream_t stream);

/* The same pass with the inputs as they really arrive: label masks as uint16 (label_bytes 2, the
 * dtype Cellpose writes below 65536 objects, Cellpose_GPU_s3fs.py:143) or int32 (4), and the
 * illumination function optionally as its reciprocal (illum_is_reciprocal != 0: the per-pixel
 * divide becomes a multiply; ips_illum_reciprocal computes 1 / illum once per plate with IEEE
 * division).  uint16 masks and reciprocal functions need W % 8 == 0, 16-byte aligned buffers and
 * Nmax <= 65535.  Float features are not bit-reproducible from run to run (float64 atomics in
 * arrival order); integer features are. */
int ips_field_fused_ex(const uint16_t* raw, const float* illum, int illum_is_reciprocal,
                       const void* labels, int label_bytes, uint16_t* maxproj, void* binned, int bin,
                       float intensity_scale, int32_t* n_objects, int32_t* ints, float* flts,
                       int Nmax, void* ws, size_t ws_bytes, int F, int C, int Z, int H, int W,
                       ips_stream_t stream);
int ips_illum_reciprocal(const float* illum, float* rcp_out, int64_t n, ips_stream_t stream);

/* ---- K2: per-plate illumination-function estimation ----------------------------------
 * Produces the {ch}_illum.npy functions that Illumination_QC_mult.py:186-193 and
 * Cellpose_GPU_s3fs.py:56 load (the reference computes them outside the repository).
 * accumulate: acc[c][y][x] += sum_f maxproj[f][c][y][x]   (uint32, exact up to 65537 fields)
 * finalize:   mean = acc / n_fields; edge-normalised separable Gaussian (sigma, truncate
 *             4.0); divide by the robust minimum (sorted positive values at position
 *             int(n * robust_frac)) and clamp below at 1.
 */
int ips_illum_accumulate(const uint16_t* maxproj, uint32_t* acc, int F, int C, int H, int W,
                         ips_stream_t stream);
size_t ips_illum_finalize_workspace_bytes(int C, int H, int W);
int ips_illum_finalize(const uint32_t* acc, uint64_t n_fields, double sigma, double robust_frac,
                       float* illum_out, void* ws, size_t ws_bytes, int C, int H, int W,
                       ips_stream_t stream);
/* Per-pixel median over a device-resident stack (median mode): stack [N][C][H][W] uint16
 * -> raw_out [C][H][W] float32 (mean of the two middle values for even N). */
int ips_illum_median(const uint16_t* stack, float* raw_out, int N, int C, int H, int W,
                     ips_stream_t stream);
/* Gaussian + robust rescale on a float32 raw function (shared by mean and median mode). */
/* (ws of ips_illum_finalize_workspace_bytes is large enough for this call too.) */
int ips_illum_smooth_rescale(const float* raw, double sigma, double robust_frac, float* illum_out,
                             void* ws, size_t ws_bytes, int C, int H, int W, ips_stream_t stream);

/* ---- K5: Pillow-exact LANCZOS resize of 16-bit planes -------------------------------
 * Replaces  img.resize(target_size, resample=LANCZOS)   Image_re-binning.py:18
 * in  [C][H][W] uint16 -> out [C][outH][outW] uint16, bit-exact with Pillow 12.2.0's
 * I;16 path (horizontal pass, round, vertical pass; double accumulation; per-byte clip).
 */
size_t ips_lanczos_workspace_bytes(int C, int H, int W, int outH, int outW);
int ips_lanczos_resize_u16(const uint16_t* in, uint16_t* out, int C, int H, int W, int outH,
                           int outW, void* ws, size_t ws_bytes, ips_stream_t stream);

/* ---- K6: radial power spectrum ring sums ----------------------------------------------
 * Replaces the ring binning of rps()   Illumination_QC_mult.py:39-43, :61-68
 * spec [F][H][W] complex64/complex128-free interface: re/im planes are passed as an
 * interleaved float64 array [F][H][W][2] (the unshifted 2-D DFT of the mean-removed
 * image).  mag_out / pow_out [F][n_rings] float64 for ring labels 2 .. n_rings+1.
 */
int ips_ring_sums(const double* spec_interleaved, double* mag_out, double* pow_out, int n_rings,
                  int F, int H, int W, ips_stream_t stream);

/* The same metric without host round trips (one image per call; see csrc/qc.cu):
 * ips_rps_prepare     exactly one of raw_u16 / img_f64 [n]; illum_f64 [n] or NULL (the float64 divide of
 *                     Illumination_QC_mult.py:145-150).  corrected_out [n] float64 receives the (divided)
 *                     image when one is computed here (uint16 input or a division; may alias fft_in_out).
 *                     fft_in_out [n] = img / median(|img - mean|) - its mean (:52-57; img - mean for a
 *                     constant image): the input of the real FFT.  Median: exact radix select.
 * ips_ring_sums_half  as ips_ring_sums on the Hermitian half [F][H][W/2+1][2] of a real FFT.
 * ips_loglog_slope    powersum [F][n_rings] -> slope_out [F]: least squares of log(power) over
 *                     log(label), labels 2.., rings with power > 0; 0.0 with fewer than three (:108-114). */
size_t ips_rps_prepare_workspace_bytes(int64_t n);
int ips_rps_prepare(const uint16_t* raw_u16, const double* img_f64, const double* illum_f64,
                    double* corrected_out, double* fft_in_out, int64_t n, void* ws, size_t ws_bytes,
                    ips_stream_t stream);
int ips_ring_sums_half(const double* half_spec_interleaved, double* mag_out, double* pow_out,
                       int n_rings, int F, int H, int W, ips_stream_t stream);
int ips_loglog_slope(const double* powersum, double* slope_out, int n_rings, int F, ips_stream_t stream);

/* ---- K4: replicate-group cosine similarity, strict upper triangle ---------------------
 * Replaces  cosine_similarity(features) -> triu(k=1) -> mean
 *           Feature_select_cosine_ami.py:145-149, Pycyto_pertime.py:132-140
 * X [N][D] float32 (NaN already replaced by 0).  Rows are L2-normalised (zero rows stay
 * zero).  group [N] int32, ids ascending from 0 with the rows of a group contiguous, or NULL
 * (one group).  sum_out[g] float64 = sum over i<j in group g of cos(i, j); npairs_out[g] =
 * #pairs.  Groups are evaluated with exact fp32 products on CUDA cores; a single group of at
 * least 1024 rows (group == NULL) runs on the tensor cores -- tcgen05.mma on a bf16 hi/lo
 * split of the normalised rows, fp32 accumulators in TMEM, |error| ~ 1e-7 per cosine.
 */
size_t ips_cosine_workspace_bytes(int N, int D);
int ips_cosine_triu(const float* X, const int32_t* group, int n_groups, double* sum_out,
                    uint64_t* npairs_out, int N, int D, void* ws, size_t ws_bytes,
                    ips_stream_t stream);

/* One large group sharded over ranks (BASELINE configs[4]: 10^6 x 3000).  The
 * normalised rows live as two bf16 planes (hi, lo) of ips_cosine_planes_bytes(N, D) bytes, 1024-byte
 * aligned: plane p starts at p * ips_cosine_plane_stride_bytes(N, D), a row is
 * ips_cosine_plane_row_bytes(D) bytes.  Every rank normalises and splits its own rows into its slice
 * (ips_cosine_split_rows), the slices are all-gathered once per plane (ips_allgather_blocks: rows as
 * blocks), and every rank runs the tensor-core pass over its share of the upper-triangular tile
 * schedule -- part `part` of `n_parts`, equal tile counts = equal triangle areas
 * (ips_cosine_triu_part; sum_out[0] = that share's sum over i < j of cos(i, j)).  The partial sums are
 * added in rank order by the caller.  N a multiple of n_parts is not required; pad with zero rows. */
size_t ips_cosine_planes_bytes(int N, int D);
size_t ips_cosine_plane_row_bytes(int D);
size_t ips_cosine_plane_stride_bytes(int N, int D);
int ips_cosine_split_rows(const float* X_rows, int n_rows, int row0, int N, int D, void* planes,
                          ips_stream_t stream);
int ips_cosine_triu_part(void* planes, double* sum_out, int N, int D, int part, int n_parts,
                         ips_stream_t stream);

/* As ips_cosine_triu (exact fp32 path) plus every pair's similarity: pairs_out holds, group
 * after group, the row-major strict upper triangle of the group's similarity matrix -- the
 * vector Pycyto_pertime.py:150-155 keeps as `cosine_similarities`; pair_offsets_out
 * [n_groups + 1] are the block boundaries (device).  The caller sizes pairs_out for
 * sum_g n_g (n_g - 1) / 2 values (pairs_capacity is checked against the one-group bound). */
size_t ips_cosine_pairs_workspace_bytes(int N, int D);
int ips_cosine_triu_pairs(const float* X, const int32_t* group, int n_groups, double* sum_out,
                          uint64_t* npairs_out, double* pairs_out, uint64_t* pair_offsets_out,
                          uint64_t pairs_capacity, int N, int D, void* ws, size_t ws_bytes,
                          ips_stream_t stream);

/* ---- well-level aggregation (consumer of the all-gather) -------------------------------
 * Replaces  df.groupby("Metadata_Well").agg("mean")    Normalize_CP_ami.py:126,
 *                                                       Pycyto_pertime.py:69-72
 * rows [N][D] float32, well [N] int32 in [0, n_wells) -> mean_out [n_wells][D] float64,
 * count_out [n_wells] int32 = rows of the well (wells without rows get NaN means, as an absent
 * group).  NaN values are skipped per column, as pandas does: every (well, column) divides by
 * its own count of non-NaN values (all-NaN -> NaN).  Any D.  Float32 rows are accumulated exactly:
 * one 64-bit INTEGER sum per (well, column, group of 8 binades) in units of that group's smallest
 * bit, added with integer atomics, so the result cannot depend on the order in which threads,
 * chunks or ranks deliver the rows (up to 2^32 values per well, column and group); only the final
 * float64 sum over the 32 groups rounds.  +-inf give a +-inf mean (NaN when both occur).
 */
size_t ips_well_mean_workspace_bytes(int n_wells, int D);
int ips_well_mean(const float* rows, const int32_t* well, double* mean_out, int32_t* count_out,
                  int N, int D, int n_wells, void* ws, size_t ws_bytes, ips_stream_t stream);
/* Streaming form of the same aggregation (same workspace): reset once, add row blocks as they
 * arrive -- e.g. one all-gather chunk at a time -- finalize once. */
int ips_well_sums_reset(void* ws, size_t ws_bytes, int D, int n_wells, ips_stream_t stream);
int ips_well_sums_add(const float* rows, const int32_t* well, int64_t N, void* ws, size_t ws_bytes,
                      int D, int n_wells, ips_stream_t stream);
int ips_well_sums_finalize(const void* ws, size_t ws_bytes, double* mean_out, int32_t* count_out,
                           int D, int n_wells, ips_stream_t stream);
/* Add a table of header-led blocks [n_blocks][block_rows][D] as ips_pack_rows_block writes and
 * ips_allgather_blocks gathers them: the well id of a row is its column 0, the valid rows of a
 * block are the first <header count> behind its header row.  No id array, no host-side counts. */
int ips_well_sums_add_blocks(const float* table, int64_t n_blocks, int64_t block_rows, void* ws,
                             size_t ws_bytes, int D, int n_wells, ips_stream_t stream);
/* The same aggregation on float64 rows (the CellProfiler tables the scripts read; ordinary
 * float64 atomics), and  df.groupby(...).agg("median")  (--well_agg_func median,
 * Normalize_CP_ami.py:126,163): perm [N] int64 = row indices grouped by well, offsets
 * [n_wells + 1] int64 = group boundaries in perm; exact (radix select), NaN skipped, mean of the
 * two middle values for even counts. */
int ips_well_mean_f64(const double* rows, const int32_t* well, double* mean_out, int32_t* count_out,
                      int64_t N, int D, int n_wells, void* ws, size_t ws_bytes, ips_stream_t stream);
int ips_well_median_f64(const double* rows, const int64_t* perm, const int64_t* offsets,
                        double* median_out, int32_t* count_out, int64_t N, int D, int n_wells,
                        ips_stream_t stream);

/* ---- centroid-centred masked cell crops, scaled to 8 bit -----------------------------------
 * Replaces the per-cell loop of Cellpose_GPU_s3fs.py:149-182 and scale_to_8bit (:34-43).
 * corrected [F][C][H][W] float32 (the illumination-corrected planes, K1's `corrected`),
 * labels [F][H][W] int32, ints / n_objects: the rows of ips_object_stats / ips_field_fused.
 * For every object in label order: centroid truncated to integers (exact), dropped when the
 * box x box window leaves the image, every channel of the window masked by `label == id`,
 * min-max scaled to [0, 255] in float32 and truncated (a constant window gives zeros).
 * crops [F][max_crops][C][box][box] uint8; kept [F][max_crops][3] int32 = label, yc, xc;
 * n_kept [F] = number of objects that passed the edge test (only the first max_crops are
 * written).  box even, box <= H, W.
 */
size_t ips_cell_crops_workspace_bytes(int F, int Nmax);
int ips_cell_crops(const float* corrected, const int32_t* labels, const int32_t* ints,
                   const int32_t* n_objects, int box, int max_crops, uint8_t* crops, int32_t* n_kept,
                   int32_t* kept, void* ws, size_t ws_bytes, int Nmax, int F, int C, int H, int W,
                   ips_stream_t stream);

/* ---- K7: TIFF strip codec (the file edge of the re-binning and max-projection scripts) ------
 * Replaces  img.save(buf, format='TIFF', compression='tiff_lzw')   Image_re-binning.py:19-21
 *      and  Image.open(...) / imageio.imread(...) of LZW strips     Image_re-binning.py:17,
 *                                                                   MaxProjection.py:39.
 * The codec is libtiff's (inside Pillow): TIFF 6.0 LZW, MSB-first 9..12-bit codes, ClearCode
 * first, reset at code 4094 or when the ratio check (every 10000 bytes) fails.  The encoder
 * follows that policy exactly: the files are byte-identical to Pillow 12.2 / libtiff 4.7.
 *
 * ips_tiff_rows_per_strip   Pillow's strip height: min(65536 / (2 W), H), at least 1.
 * ips_tiff_lzw_bound        largest LZW stream of a strip of that many bytes.
 * ips_tiff_file_bound       largest file of one H x W uint16 plane (multiple of 16).
 * ips_tiff_lzw_encode_u16   planes [P][H][W] uint16 -> files [P][file_cap] bytes, each a complete
 *                           little-endian single-IFD TIFF of file_bytes[p] bytes (0 = did not
 *                           fit; cannot happen with file_cap >= ips_tiff_file_bound).
 * ips_tiff_lzw_decode       n_strips LZW streams src[src_off[s] .. +src_bytes[s]) -> exactly
 *                           dst_bytes[s] bytes at dst[dst_off[s]]; status[s] = 0, or 1 truncated,
 *                           2 corrupt (or a stream of 512 MiB or more), 3 pre-6.0 bit order (the
 *                           remainder is zero-filled).
 *                           The five descriptor arrays are device pointers.
 * ips_tiff_fix_u16          in place on rows x W samples: byte swap (big-endian files), then
 *                           running sum modulo 2^16 along the row when predictor == 2.
 */
int ips_tiff_rows_per_strip(int H, int W);
size_t ips_tiff_lzw_bound(size_t strip_bytes);
size_t ips_tiff_file_bound(int H, int W, int rows_per_strip);
size_t ips_tiff_encode_workspace_bytes(int P, int H, int W, int rows_per_strip);
int ips_tiff_lzw_encode_u16(const uint16_t* planes, int P, int H, int W, int rows_per_strip,
                            uint8_t* files, size_t file_cap, uint64_t* file_bytes, void* ws,
                            size_t ws_bytes, ips_stream_t stream);
int ips_tiff_lzw_decode(const uint8_t* src, const uint64_t* src_off, const uint32_t* src_bytes,
                        uint8_t* dst, const uint64_t* dst_off, const uint32_t* dst_bytes,
                        int n_strips, int32_t* status, ips_stream_t stream);
int ips_tiff_fix_u16(uint16_t* img, int64_t rows, int W, int predictor, int byteswap,
                     ips_stream_t stream);

/* ---- robust-z normalisation of well profiles and the double sigmoid -------------------------
 * Replaces  pycytominer normalize(method="mad_robustize", samples=<DMSO wells>)
 *           Normalize_CP_ami.py:137-142, Pycyto_pertime.py:84-89:
 *           out = (x - median_ctrl) / (1.4826 * MAD_ctrl + 1e-18) per feature column, the
 *           statistics over the rows flagged in is_control (NaN ignored);
 *      and  double_sigmoid(x, k, alpha).abs()   Feature_select_cosine_ami.py:22-27, :117-118.
 * profiles / out [W][D] float64 (W <= 2048 wells), is_control [W] uint8.
 */
int ips_mad_robustize(const double* profiles, const uint8_t* is_control, double* out, int W, int D,
                      ips_stream_t stream);
int ips_double_sigmoid_abs(const double* x, double* y, int64_t n, int k, double alpha,
                           ips_stream_t stream);

/* ---- dense object rows (what the all-gather moves and the well aggregation reads) ---------
 * Compacts the padded per-field outputs of ips_object_stats / ips_field_fused into one
 * float32 row per object, the table the reference's consumers read from Nuclei.csv /
 * Cells.csv (Normalize_CP_ami.py:57-64):
 *   [well, field, label, area, y0, x0, y1, x1, cy, cx, C x (sum, mean, std, min, max)]
 * (D = 10 + 5C columns; integers are exact in float32).  field_well [F] int32 is the well id
 * of each field, field = field_base + index.  Fields with n_objects < 0 contribute no rows.
 * total_out (device int64) receives the number of rows.  F <= 65535.
 */
size_t ips_pack_rows_workspace_bytes(int F);
int ips_pack_rows(const int32_t* ints, const float* flts, const int32_t* n_objects,
                  const int32_t* field_well, int field_base, float* rows_out, int64_t* total_out,
                  int Nmax, int C, int F, void* ws, size_t ws_bytes, ips_stream_t stream);
/* The same rows behind a header row, the unit ips_allgather_blocks moves: block_out
 * [block_rows][D], row 0 = header (row count as two 32-bit words, rest zero), rows 1..count = the
 * objects.  block_rows - 1 >= F * Nmax.  ips_block_counts reads the headers of a table back. */
int ips_pack_rows_block(const int32_t* ints, const float* flts, const int32_t* n_objects,
                        const int32_t* field_well, int field_base, float* block_out,
                        int64_t block_rows, int Nmax, int C, int F, void* ws, size_t ws_bytes,
                        ips_stream_t stream);
int ips_block_counts(const float* table, int64_t* counts_out, int64_t n_blocks, int64_t block_rows,
                     int D, ips_stream_t stream);
/* Column 0 of a gathered [world][cap_per_rank][D] row table -> int32 well id per row, -1 for
 * the padding behind each rank's count (ips_well_mean drops ids outside [0, n_wells)). */
int ips_rows_well_ids(const float* rows, const int64_t* counts_dev, int32_t* well_out,
                      int64_t cap_per_rank, int world, int D, ips_stream_t stream);

/* ---- the one collective: all-gather of per-object rows ----------------------------------
 * north_star's "one NCCL all-gather of per-object feature rows for well-level aggregation
 * and normalisation" (the reference has no collective; parity target is the pandas groupby
 * of Normalize_CP_ami.py:126 over all rows).  NCCL is bound at run time (libnccl.so.2).
 * The unique id is created on one rank and distributed by the caller's control plane.
 * all_rows [world][cap_per_rank][row_bytes]: rank r's rows land in block r, the first
 * counts_dev[r] of them valid; counts_dev [world] int64 (device).
 */
int ips_comm_unique_id_bytes(void);
int ips_comm_unique_id(void* out, int bytes);
int ips_comm_create(void** comm_out, const void* unique_id, int bytes, int rank, int world);
int ips_comm_destroy(void* comm);
int ips_allgather_rows(void* comm, const void* local_rows, int64_t n_local, int row_bytes,
                       void* all_rows, int64_t* counts_dev, int64_t cap_per_rank,
                       ips_stream_t stream);
/* The plate pipeline's form: table [world][block_rows][row_bytes] with rank r's header-led block
 * (ips_pack_rows_block) already in place; ONE ncclAllGather, the row counts travel in the
 * headers, the launching thread never waits for the device. */
int ips_allgather_blocks(void* comm, void* table, int64_t block_rows, int row_bytes,
                         ips_stream_t stream);
int ips_comm_rank(void* comm, int* rank, int* world);
/* The bulk of the same gather over NVSwitch peer memory, without SMs: every rank's table (same size,
 * same layout, one cudaMalloc each) is exported with CUDA IPC (ips_ipc_export; the handles travel over
 * the caller's control plane), ips_peer_table_open maps the peers' tables, and ips_peer_push stores
 * a byte range of the local table into the same range of every peer's table with world - 1
 * copy-engine transfers on `stream`.  Blocks are pushed as they become final; ONE ips_allgather_blocks
 * at the end of the plate (on the per-rank totals) is the barrier that makes every push visible. */
int ips_device_alloc(void** out, size_t bytes);   /* cudaMalloc of its own, exportable */
int ips_device_free(void* p);
int ips_ipc_handle_bytes(void);
int ips_ipc_export(void* dev_ptr, void* handle_out, int bytes);
int ips_peer_table_open(void** table_out, const void* handles, int rank, int world, void* own_table);
int ips_peer_table_close(void* table);
int ips_peer_push(void* table, size_t offset, size_t bytes, ips_stream_t stream);

/* ---- host-buffer pipeline (the end-to-end call a script makes) -------------------------
 * One call = H2D of a batch of raw fields + label masks, K1, K3, D2H of max projections,
 * binned planes and object rows, on three streams with `depth` device slots so that
 * copies and kernels of consecutive batches overlap.  Host buffers must be page-locked
 * (ips_host_alloc) for the copies to be asynchronous.
 */
typedef struct ips_pipeline ips_pipeline_t;
int ips_host_alloc(void** out, size_t bytes);
/* flags: 1 = portable, 2 = write-combined (host writes / device reads only). */
int ips_host_alloc_flags(void** out, size_t bytes, unsigned flags);
int ips_host_free(void* p);
/* label_bytes: 2 = uint16 label masks (Cellpose's own dtype below 65536 objects,
 * Cellpose_GPU_s3fs.py:143; widened on the device), 4 = int32. */
int ips_pipeline_create(ips_pipeline_t** out, int fields_per_batch, int C, int Z, int H, int W,
                        int bin, int Nmax, int depth, int label_bytes,
                        const float* illum_host /* or NULL */, float intensity_scale);
/* Enqueue one batch; returns a ticket (>= 0) or a negative error.  Output pointers may be
 * NULL to skip that device->host copy. */
int64_t ips_pipeline_submit(ips_pipeline_t* p, const uint16_t* raw_host,
                            const void* labels_host, uint16_t* maxproj_host,
                            void* binned_host, int32_t* n_objects_host, int32_t* ints_host,
                            float* flts_host);
int ips_pipeline_wait(ips_pipeline_t* p, int64_t ticket); /* blocks until that batch's outputs landed */
int ips_pipeline_destroy(ips_pipeline_t* p);

#ifdef __cplusplus
}
#endif
#endif /* IPS_H_ */
