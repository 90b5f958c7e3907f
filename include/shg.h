/*
 * shg.h -- C ABI of libshg.so: the B200 (sm_100a) kernels behind the
 * Solex_ser_recon frame-stack reconstruction path.
 *
 * The reference (thelondonsmiths/Solex_ser_recon_EN) is pure Python and has no
 * FFI of its own: the drop-in boundary is its Python callables.
 * The Python modules in solex_ser_recon_en_b200/ keep those names and
 * signatures and bind the entry points below with ctypes.
 * Each entry point cites the reference code whose work it replaces.
 *
 * Conventions
 *  - every function returns 0 on success, non-zero on failure; the message is
 *    available from shg_last_error() (thread-local);
 *  - pointers named d_* are DEVICE pointers, h_* are HOST pointers; sizes are
 *    in elements unless the name says bytes;
 *  - `stream` is a cudaStream_t passed as void* (NULL = default stream); all
 *    kernels are asynchronous on it, nothing here synchronises unless stated;
 *  - a raw frame is H rows x W columns exactly as stored in the SER/AVI file;
 *    "rotated" (W > H) means the reference image is np.rot90(raw):
 *    img[i][j] = raw[j][W-1-i], ih = W slit positions, iw = H dispersion pixels
 *    (reference video_reader.py:84-91,117-120);
 *  - "disk" images are produced FRAME-MAJOR: disk[s][k][i] (shift, frame, slit
 *    position), i.e. the transpose of the reference's (ih, N) arrays, so that
 *    every kernel reads and writes contiguous runs; shg_transpose_u16 gives the
 *    reference layout where it is handed back to Python.
 */
#ifndef SHG_H
#define SHG_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SHG_VERSION 1

/* ---- status / device ---------------------------------------------------- */
const char* shg_last_error(void);
int shg_version(void);
/* out[0]=sm count, out[1]=cc major, out[2]=cc minor, out[3]=total HBM bytes,
 * out[4]=L2 bytes, out[5]=max opt-in shared memory per block */
int shg_device_info(int device, int64_t* out6);

/* ---- a3: mean / max frame (reference solex_util.py:174-188) ------------- */
/* Add frames [0,n_frames) of d_frames (contiguous, frame_px pixels each,
 * bytes_per_px 1 or 2, raw file units) into d_sum[frame_px] (uint64) and
 * d_max[frame_px] (uint32).  Integer atomics: the result does not depend on
 * launch geometry; partial results of frame ranges / ranks add exactly. */
int shg_accumulate(const void* d_frames, int bytes_per_px, int64_t n_frames, int64_t frame_px,
                   uint64_t* d_sum, uint32_t* d_max, void* stream);

/* d_out[k] = sum of the raw pixels of frame k (file units): all_video_reader.means[k] = scale * d_out[k] / frame_px
 * exactly as np.mean(frame) gives it (reference video_reader.py:143-147; integer sums below 2^53 are exact). */
int shg_frame_sums(const void* d_frames, int bytes_per_px, int64_t n_frames, int64_t frame_px,
                   uint64_t* d_out, void* stream);

/* mean = floor(scale*sum / n_total), max = scale*max (scale = 256 for 8-bit
 * input: reference video_reader.py:121-122), written in IMAGE orientation
 * (ih x iw, rotated when W > H) as uint16. */
int shg_finalize_mean_max(const uint64_t* d_sum, const uint32_t* d_max, int64_t n_total,
                          int W, int H, int eight_bit,
                          uint16_t* d_mean_img, uint16_t* d_max_img, void* stream);

/* ---- a4/a5: line detection (reference solex_util.py:165-172,223-231,242) */
/* cv2.blur(img, (kw, kh)) on uint16 as OpenCV 4.13 computes it: exact box sum,
 * BORDER_REFLECT_101, anchor k/2, rint(float32(S)*float32(1/(kw*kh))) for
 * columns < cols - cols%8 and rint(S*(1.0/(kw*kh))) in double for the rest.
 * d_tmp: scratch of rows*cols uint32. */
int shg_box_blur_u16(const uint16_t* d_img, int rows, int cols, int kw, int kh,
                     uint16_t* d_out, uint32_t* d_tmp, void* stream);
/* per-row integer sums (for np.mean(blur, axis=1)) */
int shg_row_sums_u16(const uint16_t* d_img, int rows, int cols, uint64_t* d_out, void* stream);
/* per-row FIRST argmin over columns [c0, c1) -> absolute column index */
int shg_row_argmin_u16(const uint16_t* d_img, int rows, int cols, int c0, int c1,
                       int32_t* d_out, void* stream);

/* ---- a6: cubic least-squares fit (reference solex_util.py:233-259) ------ */
/* Masked cubic fit of y[i] against x = x0 + i, i in [0,n): fp64 moment sums
 * reduced with warp shuffles, 4x4 solve on the device.  d_mask may be NULL.
 * d_coef[4]: ascending coefficients in raw x (as np.flip(np.polyfit(...))).
 * Optionally (d_resid != NULL) writes resid[i] = polyval(x0+i) - y_resid[i]. */
int shg_polyfit3(const int32_t* d_y, const uint8_t* d_mask, int x0, int n,
                 double* d_coef, const int32_t* d_y_resid, double* d_resid, void* stream);
/* keep[i] = |resid[i] / std(resid)| < nsigma   (np.std, ddof 0) */
int shg_sigma_mask(const double* d_resid, int n, double nsigma, uint8_t* d_keep, void* stream);
/* good[i] = |resid[i] - centre| < tol */
int shg_window_mask(const double* d_resid, int n, double centre, double tol, uint8_t* d_good, void* stream);
/* fit table rows [floor(c), c-floor(c), y, c], c = polyval(y), y in [0,ih) */
int shg_fit_table(const double* d_coef, int ih, double* d_fit /* ih x 4 */, void* stream);

/* ---- a7: per-frame reconstruction (reference solex_util.py:93-144) ------ */
/* For every frame k, shift s, slit position i:
 *   il = clamp((int)fit[i][0] + shift[s], 0, iw-2)
 *   disk[s][k][i] = trunc(L*lw + R*rw), L/R = img_k[i][il], img_k[i][il+1]
 *   lw = 1 - fit[i][1], rw = 1 - lw   (fp64, separate mul/mul/add)
 * 8-bit input is scaled by 256.  h_fit is the HOST (ih x 4) fit table exactly
 * as read_video_improved receives it.  Disks are frame-major; element (s,k,i)
 * lives at image_s[(k0_out+k)*ih + i], where image_s = d_disk + s*shift_stride,
 * or, when h_out_ptrs (HOST array of n_shifts device addresses) is given,
 * image_s = h_out_ptrs[s]: the images may then live on PEER GPUs (addresses
 * from shg_ipc_open), so each rank writes its frame rows straight into the
 * owner's image over NVLink -- reconstruction and the row exchange are one
 * kernel.  d_work: device scratch of at least shg_recon_workspace_bytes(ih,
 * n_shifts) bytes.  impl: bits 0-7: 0 = auto, 1 = generic direct-load kernel, 2 = TMA
 * band kernel; bits 8-15: use at most that many SMs (0 = all), for callers that run
 * latency-critical small kernels beside this one.
 * d_min (optional, device uint32[n_shifts], in the order of h_shifts): the kernel
 * folds min(d_min[s], minimum of the pixels it writes for shift s) into it, so the
 * caller pre-fills it with 65535 (or the partial minimum of other frame ranges);
 * the circularisation clips to that minimum (ellipse_to_circle.py:112-118) and
 * does not have to re-read the images for it.  *h_min_done (optional, HOST) is
 * set to 1 when the kernel variant that ran maintains d_min, else 0 (the caller
 * then computes the minimum with shg_minmax_u16). */
int64_t shg_recon_workspace_bytes(int ih, int n_shifts);
int shg_recon(const void* d_frames, int bytes_per_px, int64_t n_frames, int W, int H,
              const double* h_fit, const int32_t* h_shifts, int n_shifts,
              uint16_t* d_disk, int64_t shift_stride, const uint64_t* h_out_ptrs, int64_t k0_out, int impl,
              void* d_work, int64_t work_bytes, uint32_t* d_min, int* h_min_done, void* stream);

/* ---- multi-GPU row exchange: device buffers other ranks of the box can write
 * (cudaMalloc + CUDA IPC; handles travel through torch.distributed) -------- */
#define SHG_IPC_HANDLE_BYTES 64
int shg_ipc_alloc(int64_t bytes, void** d_ptr, unsigned char* handle64);
int shg_ipc_free(void* d_ptr);
int shg_ipc_open(const unsigned char* handle64, void** d_ptr);   /* enables peer access lazily */
int shg_ipc_close(void* d_ptr);

/* ---- layout helpers ------------------------------------------------------ */
/* out[c][r'] = in[r][c] for in of shape (rows, cols); flip != 0 reverses the
 * output's fastest axis (r' = rows-1-r): frame-major disk -> reference (ih, N)
 * layout, with the reference's flip_x (Solex_recon.py:75-76) fused. */
int shg_transpose_u16(const uint16_t* d_in, int64_t rows, int64_t cols, uint16_t* d_out, int flip, void* stream);
/* Batched min / max: image j (j < n_imgs) starts at d_in + (d_sel ? d_sel[j] : j)*img_stride
 * and has n uint16 values; d_out[2j] = min, d_out[2j+1] = max (initialised here). */
int shg_minmax_u16(const uint16_t* d_in, int64_t n, int64_t img_stride, const int32_t* d_sel, int n_imgs,
                   uint32_t* d_out, void* stream);

/* Position-sensitive 64-bit checksum of n uint16 values: *d_out = sum_i (v[i] + 1) * (mix(i) | 1) mod 2^64 with
 * mix(i) = t ^ (t >> 29), t = (i + 1) * 0x9E3779B97F4A7C15.  Integer, order-independent.  bench.py folds the
 * checksums of the final images (reference Solex_recon.py:136-152 output, one per requested shift) into the
 * `outputs_crc` it prints, so that runs on 1, 2, 4 and 8 GPUs can be compared bit for bit. */
int shg_checksum_u16(const uint16_t* d_in, int64_t n, uint64_t* d_out, void* stream);

/* ---- a10: circularisation warp (reference ellipse_to_circle.py:94-118) -- */
/* Per-row 1-D linear resample of frame-major disks (n_frames x ih each), batched
 * over n_imgs images that share the geometry (every shift of one scan):
 *   x = (m00*c + m01*r) + m02; out[r][c] = trunc(clip((1-d)*in[floor x][r] + d*in[ceil x][r]))
 * taps outside [0,n_frames) (or rows >= ih) read cval = image[0][0]; the clip
 * range [lo,hi] = the image's min / max as skimage.transform.warp uses it, read
 * from d_minmax (shg_minmax_u16 output, stays on the device).  Image j is
 * d_disk + (d_sel ? d_sel[j] : j)*disk_stride; output j is row-major
 * (out_rows x out_cols) at d_out + j*out_stride.  flip != 0 reads the disks with
 * their frame axis reversed (the reference's flip_x). */
int shg_warp_rows(const uint16_t* d_disk, int64_t disk_stride, const int32_t* d_sel, int n_imgs,
                  int64_t n_frames, int ih, int flip, double m00, double m01, double m02,
                  const uint32_t* d_minmax, uint16_t* d_out, int64_t out_stride, int out_rows, int out_cols,
                  void* stream);
/* The same for a scan whose frames are spread over several GPUs: d_disk is a base pointer such that
 * d_disk + k*ih (+ sel*disk_stride) is frame k of the WHOLE scan, of which this rank holds [own_lo, own_hi]
 * (logical frame order, i.e. after flip; INT_MIN / INT_MAX = open end).  The call produces exactly the
 * output pixels whose left tap floor(x) lies in [own_lo, own_hi) -- its own frames plus ONE frame of the
 * next rank are all it reads, whatever the tilt -- so the ranks' calls tile the image.  d_cval[i]
 * (optional) replaces the read of image i's pixel [0][0] (the constant for taps outside the image; it
 * lives on one rank only), and d_out_ptrs[i] (optional, DEVICE array of addresses) replaces
 * d_out + i*out_stride: the images may live on PEER GPUs (shg_ipc_open), so each rank stores its pixels of
 * the circularised image straight into the owner's buffer over NVLink. */
int shg_warp_rows_window(const uint16_t* d_disk, int64_t disk_stride, const int32_t* d_sel, int n_imgs,
                         int64_t n_frames, int ih, int flip, double m00, double m01, double m02,
                         const uint32_t* d_minmax, uint16_t* d_out, int64_t out_stride, int out_rows,
                         int out_cols, const uint32_t* d_cval, int own_lo, int own_hi,
                         const uint64_t* d_out_ptrs, void* stream);

/* The TMA formulation of the same warp (csrc/warp.cu, warp_tma_kernel): the input tile is staged by
 * cp.async.bulk.tensor boxes, a lane owns a slit row and walks output columns, eight pixels leave as one 16-byte
 * store.  Identical results.  The disks are described as they lie in memory: d_disk points at frame
 * `frame_origin` (PHYSICAL frame index) of image 0, n_local_frames frames per image are present, images are
 * disk_stride elements apart and there are n_disk_images of them (d_sel picks among them).  Complete images:
 * frame_origin 0, n_local_frames == n_frames, own_lo / own_hi = INT_MIN / INT_MAX.  Frame-sharded scans: the
 * rank's frames plus its halo, [own_lo, own_hi) in logical order as for shg_warp_rows_window, d_cval required.
 * Needs ih % 8 == 0, disk_stride % 8 == 0 and 16-byte aligned buffers: shg_warp_rows_tma_ok() says whether a
 * geometry qualifies (else use shg_warp_rows / shg_warp_rows_window). */
int shg_warp_rows_tma_ok(const uint16_t* d_disk, int64_t disk_stride, int ih, const uint16_t* d_out, int64_t out_stride,
                         const uint64_t* d_out_ptrs);
int shg_warp_rows_tma(const uint16_t* d_disk, int64_t disk_stride, int n_disk_images, int64_t n_local_frames,
                      int64_t frame_origin, const int32_t* d_sel, int n_imgs, int64_t n_frames, int ih, int flip,
                      double m00, double m01, double m02, const uint32_t* d_minmax, uint16_t* d_out, int64_t out_stride,
                      int out_rows, int out_cols, const uint32_t* d_cval, int own_lo, int own_hi,
                      const uint64_t* d_out_ptrs, void* stream);
/* Exchange step of the frame-sharded circularisation: copy, for every image and row, the columns this
 * rank produced (left tap in [own_lo, own_hi): one interval per row) from its local full-width images
 * (d_local + i*local_stride) into the image's owner (d_out_ptrs[i], local or PEER memory) with 16-byte stores in
 * 512-byte contiguous runs per warp.  Rows whose destination is the source are skipped. */
int shg_exchange_rows(const uint16_t* d_local, int64_t local_stride, int n_imgs, int out_rows, int out_cols, double m00,
                      double m01, double m02, int own_lo, int own_hi, const uint64_t* d_out_ptrs, void* stream);

/* ---- a11 helper: 4x4 block sums (reference ellipse_to_circle.py:301) ---- */
/* downscale_local_mean numerator: out[ri][ci] = sum of the 4x4 block of the
 * (ih, n_frames) image (zero padded), from a frame-major disk. */
int shg_downscale4_sum(const uint16_t* d_disk, int64_t n_frames, int ih, int flip,
                       uint32_t* d_out, int out_rows, int out_cols, void* stream);

/* ---- a11: limb detection for the ellipse fit (reference ellipse_to_circle.py:148-250)
 * Works on the integer 4x4 block sums S of shg_downscale4_sum (the reference's
 * downscaled image/65536 is S * 2^-20 exactly).  See csrc/limb.cu for how each
 * step mirrors cv2.blur / np.percentile / np.histogram / scipy.ndimage /
 * skimage.feature.canny operation order. */
/* exact box sums, BORDER_REFLECT_101, anchor k/2 (cv2.blur on CV_64F before its scale) */
int shg_box_sum_u32(const uint32_t* d_in, int rows, int cols, int kw, int kh, uint32_t* d_out,
                    uint32_t* d_tmp, void* stream);
int shg_sum_u32(const uint32_t* d_in, int64_t n, uint64_t* d_out, void* stream);
/* exact order statistics: h_out[q] = the h_ranks[q]-th smallest (0-based) of d_vals (byte-wise radix
 * select; ranks that share a prefix share a pass).  Blocking: reads a 256-bin histogram back per pass
 * (keeping the state on the device and queueing all passes was measured SLOWER when the limb search
 * runs underneath the reconstruction kernel).  d_work: at least 256 uint32 of device scratch. */
int shg_select_u32(const uint32_t* d_vals, int64_t n, const int64_t* h_ranks, int n_ranks,
                   uint32_t* h_out, uint32_t* d_work, void* stream);
/* blurred = (B * 2^-20) * scale.  d_out2 = {min B, max B} over pixels with blurred < ceiling */
int shg_blur_range(const uint32_t* d_box, int64_t n, double scale, double ceiling, uint32_t* d_out2, void* stream);
/* np.histogram(blurred[blurred < ceiling], bins=n_bins) given its n_bins+1 edges; d_counts: 32 uint64 */
int shg_blur_hist(const uint32_t* d_box, int64_t n, double scale, double ceiling, const double* h_edges,
                  int n_bins, uint64_t* d_counts, void* stream);
/* flood = blurred < level ? 0 : 65000;  smoothed = gaussian(flood) / (gaussian(ones) + eps),
 * scipy.ndimage.gaussian_filter order (axis 0 then 1, mode constant); h_weights[k] = weight at
 * offset k (0..radius).  d_tmp2: two rows*cols double images. */
int shg_flood_smooth(const uint32_t* d_box, int rows, int cols, double scale, double level,
                     const double* h_weights, int radius, double eps, double* d_smoothed,
                     double* d_tmp2, void* stream);
/* scipy.ndimage.sobel along axis 0 (d_gi) and axis 1 (d_gj), mode reflect, and sqrt(gi*gi + gj*gj) */
int shg_sobel_mag(const double* d_smoothed, int rows, int cols, double* d_gi, double* d_gj, double* d_mag,
                  void* stream);
/* skimage canny's interpolated non-maximum suppression: pixels with mag >= low that are local maxima
 * along the gradient are appended (unordered) to d_list_idx (flat index) / d_list_mag; *d_count may
 * exceed cap, in which case the list is truncated and the caller retries with a larger cap. */
int shg_nms_candidates(const double* d_gi, const double* d_gj, const double* d_mag, int rows, int cols,
                       double low, uint32_t* d_count, uint32_t cap, uint32_t* d_list_idx,
                       double* d_list_mag, void* stream);

/* The whole threshold search of get_flood_image + the median behind canny's thresholds (reference
 * ellipse_to_circle.py:148-175, 243-244) as ONE queue of kernels with every data-dependent scalar kept on the
 * device, and one blocking read-back at the end (the step-by-step entry points above cost ~15 round trips):
 *   box  = bw x bw box sums of S, box5 = 5 x 5 box sums (d_box, d_box5, d_tmp: rows*cols uint32 each);
 *   order statistics h_ranks4 = {two ranks in box (np.percentile's bracket), two ranks in box5 (np.median's)};
 *   ceiling = np.percentile(blurred, q) from its bracket with interpolation weight `gamma`;
 *   edges = np.linspace(min, max of blurred < ceiling, n_bins + 1); counts = np.histogram on those edges.
 * d_state: shg_limb_state_bytes() bytes of device scratch.  h_out (>= 8 + 33 + 32 doubles), all exact:
 *   [0] sum of S, [1..4] the four order statistics (box sums), [5] ceiling, [6] [7] min / max box sum below
 *   the ceiling, [8 .. 8+n_bins] edges, [41 .. 41+n_bins) counts.  Synchronises the stream. */
int64_t shg_limb_state_bytes(void);
int shg_limb_front(const uint32_t* d_sums, int rows, int cols, int bw, const int64_t* h_ranks4, double gamma,
                   int n_bins, uint32_t* d_box, uint32_t* d_box5, uint32_t* d_tmp, void* d_state, double* h_out,
                   void* stream);
/* shg_flood_smooth + shg_sobel_mag + shg_nms_candidates queued back to back, then the candidate count and
 * the first `first_chunk` list entries copied to the host in one round trip (the rest, if any, in a second).
 * d_buf6: six rows*cols double images.  *h_count may exceed cap (list truncated: retry with a larger cap).
 * h_list_*: host buffers of `cap` entries (pinned for speed).  Synchronises the stream. */
int shg_limb_canny(const uint32_t* d_box, int rows, int cols, double scale, double level, const double* h_weights,
                   int radius, double eps, double low, double* d_buf6, uint32_t* d_count, uint32_t cap,
                   uint32_t* d_list_idx, double* d_list_mag, uint32_t first_chunk, uint32_t* h_count,
                   uint32_t* h_list_idx, double* h_list_mag, void* stream);

/* HOST helper (no GPU): indices of the strict convex-hull vertices of n integer points (x0,y0,x1,y1,...),
 * in hull order; collinear points on a hull edge are not vertices (the set scipy.spatial.ConvexHull(...).vertices
 * reports; reference ellipse_to_circle.py:263-269 only tests which regions own a hull vertex). */
int shg_hull_vertices(const int64_t* h_xy, int64_t n, int64_t* h_vertex_index, int64_t* h_n_vertices);

/* HOST helper (no GPU): S[6][6] = sum of d d^T over n points (x0,y0,x1,y1,...), d = [x^2, xy, y^2, x, y, 1]: the
 * scatter matrix of the direct least-squares ellipse fit (lsq-ellipse, reference ellipse_to_circle.py:57-59). */
int shg_conic_scatter(const double* h_xy, int64_t n, double* h_s36);

/* HOST helper (no GPU): 8-connected components of a sparse pixel list (flat = row*cols + col,
 * strictly ascending), labelled 1.. in raster order of each component's first pixel, i.e. what
 * scipy.ndimage.label(edges, ones((3,3))) gives on those pixels (reference ellipse_to_circle.py:252). */
int shg_label_points(const int64_t* h_flat, int64_t n, int64_t cols, int32_t* h_labels, int32_t* h_n_labels);

/* ---- a12: transversalium (reference solex_util.py:76-86,383-395,489-516) */
/* Pixels are uint16, so the reference's log(img[y]/img[y-1]) is taken as L(a) - L(b), where L(v) is
 * log(v) to ~2e-15 absolute (L(0) = -inf) computed in-kernel.  This entry point tabulates that same
 * L for v in [0, 65536) (tests, diagnostics); the row statistics do not read a table. */
int shg_log_table(double* d_tab65536, void* stream);
/* For each listed row y (rows[j]) of each of n_imgs images (image i at
 * d_img + i*img_stride), over columns [xa[j], xb[j]):
 *   rat = log(img[y][x] / img[y-1][x]);  out[i*n_list + j] = mean(rat[|rat-med|/MAD < 2])
 * (median / MAD as np.median; MAD == 0 keeps all; an empty chord or a nan
 * gives nan as in the reference).  One CTA per (row, image), exact order
 * statistics: a register-resident kernel takes every ordinary row (window selects
 * whose counts are taken from the data; the MAD window comes from the median's
 * histogram), and hands the rest (zero pixels, chords <= 256 px, heavy ties) to
 * the classic kernel (counting select, radix select as fall-back) through a todo
 * list in d_work.  max_len = max(xb-xa).  d_work: shg_transv_workspace_bytes
 * bytes (the todo list; plus scratch rows when a chord exceeds shared memory);
 * without it only the classic kernel runs. */
int64_t shg_transv_workspace_bytes(int n_list, int max_len, int n_imgs);
int shg_transv_row_stats(const uint16_t* d_img, int rows, int cols, int n_imgs, int64_t img_stride,
                         const int32_t* d_rows, const int32_t* d_xa, const int32_t* d_xb, int n_list,
                         int max_len, double* d_out, void* d_work, int64_t work_bytes, void* stream);
/* Gain vectors from the row statistics (reference solex_util.py:400-404,456-479):
 * ratios = [0, stats...] (n values); trend = savgol_filter(ratios, window, 3) with
 * d_coeffs = scipy.signal.savgol_coeffs(window, 3) and cubic end fits; detrended
 * minus its mean; correction = exp(-cumsum); gain[y1:y1+n] = 1 + (correction-1)*taper,
 * 1 elsewhere.  d_stats: n_imgs x (n-1); d_gains: n_imgs x n_rows.  window odd, 5 <= window <= n. */
int shg_transv_gain(const double* d_stats, int n_imgs, int n, int window, const double* d_coeffs,
                    const double* d_taper, int y1, int n_rows, double* d_gains, void* stream);
/* out[i][r][c] = trunc(min(img[i][r][c] * gain[i][r], 65535)), images at stride img_stride */
int shg_row_scale_u16(const uint16_t* d_img, int rows, int cols, int n_imgs, int64_t img_stride,
                      const double* d_gain, uint16_t* d_out, void* stream);

/* ---- a1/a2: ingest (reference video_reader.py:94-123) ------------------- */
/* ---- f2: the image_process tail (reference solex_util.py:527-546) --------------------------
 * CLAHE on uint16 exactly as cv2.createCLAHE(clipLimit, tileGridSize).apply computes it (OpenCV
 * modules/imgproc/src/clahe.cpp: 65536-bin tile histograms of the image extended by BORDER_REFLECT_101 to a
 * multiple of the tile grid, clip + redistribution, float lut, float bilinear blend of the four tile luts) and
 * rescale_brightness (:519-525).  Histograms are uint32[65536] per tile (tile-major); the percentiles the
 * reference takes (np.percentile(frame, 99.9999), np.percentile(cl1, 10), np.max(cl1)) are read off the
 * whole-image histograms d_full_hist / d_out_hist by the host. */
int shg_tile_hist_u16(const uint16_t* d_img, int rows, int cols, int tiles_x, int tiles_y, uint32_t* d_tile_hist,
                      uint32_t* d_full_hist, void* stream);
/* pixels per tile of the (possibly extended) image, as OpenCV sizes its tiles */
int64_t shg_clahe_tile_area(int rows, int cols, int tiles_x, int tiles_y);
/* clip / redistribute the tile histograms in place and build the tile luts (uint16[n_tiles][65536]) */
int shg_clahe_lut(uint32_t* d_tile_hist, int n_tiles, int64_t tile_area, double clip_limit, uint16_t* d_lut, void* stream);
int shg_clahe_apply(const uint16_t* d_img, int rows, int cols, int tiles_x, int tiles_y, const uint16_t* d_lut,
                    uint16_t* d_out, uint32_t* d_out_hist, void* stream);
/* out = trunc(clip(65535.0 * (img - lo) / (hi - lo), 0, 65535)) in double, left to right */
int shg_rescale_u16(const uint16_t* d_img, int64_t n, double lo, double hi, uint16_t* d_out, void* stream);

/* Streams frames of a file into a device-resident stack through a ring of
 * pinned host buffers filled by reader threads (pread), with the H2D copies on
 * a private copy stream and, when d_sum/d_max are given, the mean/max
 * accumulation of each chunk overlapped on a private compute stream.
 * frame_stride_bytes > frame_bytes skips per-frame container headers (AVI). */
typedef struct shg_ingest shg_ingest;
int shg_ingest_create(int device, int64_t slot_bytes, int n_slots, int n_threads, shg_ingest** out);
int shg_ingest_destroy(shg_ingest* ing);
/* Blocking.  Reads frames [frame0, frame0+n_frames) of `path`.  h_stats[0..3]:
 * seconds total, seconds waiting on the file, bytes copied, chunks. */
int shg_ingest_file(shg_ingest* ing, const char* path, int64_t payload_offset,
                    int64_t frame_bytes, int64_t frame_stride_bytes, int64_t frame0, int64_t n_frames,
                    void* d_stack, int bytes_per_px, uint64_t* d_sum, uint32_t* d_max, double* h_stats4);
/* Same pipeline from a host memory image of the payload (e.g. an mmap). */
int shg_ingest_memory(shg_ingest* ing, const void* h_payload, int64_t frame_bytes,
                      int64_t frame_stride_bytes, int64_t n_frames, void* d_stack, int bytes_per_px,
                      uint64_t* d_sum, uint32_t* d_max, double* h_stats4);

/* Pinned host memory of an exact size (cudaHostAlloc, portable) for payloads
 * and result images, and a stream-ordered copy in either direction
 * (kind: 1 = host->device, 2 = device->host, 3 = device->device). */
int shg_host_alloc(int64_t bytes, void** out);
int shg_host_free(void* p);
int shg_memcpy_async(void* dst, const void* src, int64_t bytes, int kind, void* stream);

/* ---- synthetic scans on the device (bench / full-size property tests) --- */
/* Fills frames [k0, k0+n) of a synthetic scan of n_total frames (same recipe
 * family as solex_ser_recon_en_b200/synth.py, hash noise). */
int shg_synth_fill(void* d_frames, int bytes_per_px, int64_t k0, int64_t n, int64_t n_total,
                   int W, int H, uint64_t seed, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SHG_H */
