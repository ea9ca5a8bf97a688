/* rocketfft_b200 -- C ABI of the B200-native transform library (librocketfft_b200.so).
 *
 * Two families of entry points:
 *
 * (1) The ten `numba_*` symbols.  They replace, one for one and with identical
 *     signatures, the externs the reference exports from
 *     rocket_fft/_pocketfft_numba.cpp (good_size :25-29, c2c :31-49, dct :51-69,
 *     dst :71-89, r2c :91-109, c2c_sym :111-143, c2r :145-163, r2r_fftpack :165-183,
 *     r2r_separable_hartley :185-203, r2r_genuine_hartley :205-223) and that Numba-
 *     compiled code calls by name (rocket_fft/pocketfft.py:33-128).  Arrays arrive as
 *     pointers to Numba's array record (numba `_arraystruct.h`), see rfb200_array_record.
 *     `data` may be a host pointer (the record of a NumPy array: the shim stages
 *     H2D -> kernels -> D2H on the library's stream and returns when the result is in
 *     host memory) or a device pointer (detected with cudaPointerGetAttributes: no
 *     staging, the call is asynchronous on the current library stream).
 *     As in the reference they return void; failures are recorded and can be read
 *     with rfb200_last_error().
 *
 * (2) `rfb200_*` device entry points for arrays that are already resident in HBM
 *     (`__cuda_array_interface__`): plain shape / byte-stride / axes arrays, raw device
 *     pointers and a CUDA stream.  They return 0 on success, nonzero on error.
 *
 * There is no CPU fallback behind any of these: every transform runs as sm_100a CUDA
 * kernels; only good_size is host integer arithmetic.
 */
#ifndef ROCKETFFT_B200_H
#define ROCKETFFT_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Numba's array record (numba/_arraystruct.h), as read by the reference at
 * _pocketfft_numba.cpp:34-37: shape = shape_and_strides[0..ndim),
 * strides (bytes, signed) = shape_and_strides[ndim..2*ndim). */
typedef struct rfb200_array_record {
    void *meminfo;
    void *parent;
    intptr_t nitems;
    intptr_t itemsize;
    void *data;
    intptr_t shape_and_strides[1]; /* really 2*ndim entries */
} rfb200_array_record;

/* ---- (1) drop-in symbols (reference: _pocketfft_numba.cpp:25-223) ------------- */
uint64_t numba_good_size(uint64_t target, bool real);
void numba_c2c(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
               rfb200_array_record *axes, bool forward, double fct, uint64_t nthreads);
void numba_r2c(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
               rfb200_array_record *axes, bool forward, double fct, uint64_t nthreads);
void numba_c2r(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
               rfb200_array_record *axes, bool forward, double fct, uint64_t nthreads);
void numba_c2c_sym(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                   rfb200_array_record *axes, bool forward, double fct, uint64_t nthreads);
void numba_dct(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
               rfb200_array_record *axes, uint64_t type, double fct, bool ortho, uint64_t nthreads);
void numba_dst(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
               rfb200_array_record *axes, uint64_t type, double fct, bool ortho, uint64_t nthreads);
void numba_r2r_fftpack(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                       rfb200_array_record *axes, bool real2hermitian, bool forward, double fct,
                       uint64_t nthreads);
void numba_r2r_separable_hartley(uint64_t ndim, const rfb200_array_record *ain,
                                 rfb200_array_record *aout, rfb200_array_record *axes, double fct,
                                 uint64_t nthreads);
void numba_r2r_genuine_hartley(uint64_t ndim, const rfb200_array_record *ain,
                               rfb200_array_record *aout, rfb200_array_record *axes, double fct,
                               uint64_t nthreads);

/* ---- (2) device-resident entry points ------------------------------------------- */
/* precision: 0 = single (float32 / complex64), 1 = double (float64 / complex128).
 * `shape` is the shape the reference reads: the input's for every op except c2r,
 * where it is the (real) output's.  Strides are signed bytes.  `stream` is a
 * cudaStream_t passed as void* (NULL = legacy default stream). */
typedef enum { RFB200_F32 = 0, RFB200_F64 = 1 } rfb200_precision;

int rfb200_c2c(int precision, size_t ndim, const int64_t *shape, const int64_t *stride_in,
               const int64_t *stride_out, size_t naxes, const uint64_t *axes, int forward,
               double fct, const void *d_in, void *d_out, void *stream);
int rfb200_r2c(int precision, size_t ndim, const int64_t *shape_in, const int64_t *stride_in,
               const int64_t *stride_out, size_t naxes, const uint64_t *axes, int forward,
               double fct, const void *d_in, void *d_out, void *stream);
int rfb200_c2r(int precision, size_t ndim, const int64_t *shape_out, const int64_t *stride_in,
               const int64_t *stride_out, size_t naxes, const uint64_t *axes, int forward,
               double fct, const void *d_in, void *d_out, void *stream);
int rfb200_c2c_sym(int precision, size_t ndim, const int64_t *shape, const int64_t *stride_in,
                   const int64_t *stride_out, size_t naxes, const uint64_t *axes, int forward,
                   double fct, const void *d_in, void *d_out, void *stream);
int rfb200_dct(int precision, size_t ndim, const int64_t *shape, const int64_t *stride_in,
               const int64_t *stride_out, size_t naxes, const uint64_t *axes, int type,
               double fct, int ortho, const void *d_in, void *d_out, void *stream);
int rfb200_dst(int precision, size_t ndim, const int64_t *shape, const int64_t *stride_in,
               const int64_t *stride_out, size_t naxes, const uint64_t *axes, int type,
               double fct, int ortho, const void *d_in, void *d_out, void *stream);
int rfb200_r2r_fftpack(int precision, size_t ndim, const int64_t *shape,
                       const int64_t *stride_in, const int64_t *stride_out, size_t naxes,
                       const uint64_t *axes, int real2hermitian, int forward, double fct,
                       const void *d_in, void *d_out, void *stream);
int rfb200_r2r_separable_hartley(int precision, size_t ndim, const int64_t *shape,
                                 const int64_t *stride_in, const int64_t *stride_out,
                                 size_t naxes, const uint64_t *axes, double fct,
                                 const void *d_in, void *d_out, void *stream);
int rfb200_r2r_genuine_hartley(int precision, size_t ndim, const int64_t *shape,
                               const int64_t *stride_in, const int64_t *stride_out,
                               size_t naxes, const uint64_t *axes, double fct, const void *d_in,
                               void *d_out, void *stream);

/* c2c along ONE axis with the output axis scattered: output index k along `axis` (extent n) is
 * cut into `nparts` equal blocks and block h is written to d_out_parts[h] -- an array shaped like
 * `shape` with the axis extent n/nparts, addressed with stride_out -- which may live on a PEER GPU
 * (NVLink-mapped pointer): the local transform and the all-to-all push of a slab-decomposed N-D
 * transform happen in one kernel.  Needs a power-of-two axis length (16..16384).  No reference
 * counterpart (the reference is single-process). */
int rfb200_c2c_scatter(int precision, size_t ndim, const int64_t *shape, const int64_t *stride_in,
                       const int64_t *stride_out, size_t axis, int forward, double fct,
                       const void *d_in, size_t nparts, void *const *d_out_parts, void *stream);

/* ---- (3) the step either side of the path: fused zero-padding / cropping, index rotation ------
 * The reference's numpy/scipy layer materialises a zero-padded (or cropped) copy of the input before
 * calling the transform (rocket_fft/overloads.py:575-609, `n` / `s` arguments).  These entry points
 * take the array as it is (`shape_in`) plus the shape that is to be transformed (`shape`): along
 * transformed axes the input is cropped to, or zero-extended to, the transform extent while its lines
 * are loaded; lines that would consist of padding only are never read.  Extents may differ only
 * along axes listed in `axes`.
 *   rfb200_c2c_pad: shape_in / shape complex; output has `shape`.
 *   rfb200_r2c_pad: shape_in / shape real; output has `shape` with axes[naxes-1] -> n/2+1.
 *   rfb200_c2r_pad: shape_in complex (the bins present), shape = real OUTPUT shape; bins beyond
 *                   shape_in are zero, bins beyond n/2 along axes[naxes-1] are ignored. */
int rfb200_c2c_pad(int precision, size_t ndim, const int64_t *shape_in, const int64_t *shape,
                   const int64_t *stride_in, const int64_t *stride_out, size_t naxes,
                   const uint64_t *axes, int forward, double fct, const void *d_in, void *d_out,
                   void *stream);
int rfb200_r2c_pad(int precision, size_t ndim, const int64_t *shape_in, const int64_t *shape,
                   const int64_t *stride_in, const int64_t *stride_out, size_t naxes,
                   const uint64_t *axes, int forward, double fct, const void *d_in, void *d_out,
                   void *stream);
int rfb200_c2r_pad(int precision, size_t ndim, const int64_t *shape_in, const int64_t *shape,
                   const int64_t *stride_in, const int64_t *stride_out, size_t naxes,
                   const uint64_t *axes, int forward, double fct, const void *d_in, void *d_out,
                   void *stream);
/* out[(i + shift[d]) mod shape[d]] = in[i] along every dim d (np.roll; fftshift: shift = n/2,
 * ifftshift: shift = -(n/2); reference: rocket_fft/overloads.py:752-855, 1221-1310).
 * itemsize 4, 8 or 16 bytes; in and out must not overlap. */
int rfb200_roll(int itemsize, size_t ndim, const int64_t *shape, const int64_t *stride_in,
                const int64_t *stride_out, const int64_t *shift, const void *d_in, void *d_out,
                void *stream);

/* d_data[l][j] *= d_table[j] for l < nlines, j < n; lines contiguous, items real (complex_items = 0)
 * or complex (1), precision as above.  The coefficient multiply between the rfft and the irfft of the
 * fast Hankel transform and its bias factors (reference: rocket_fft/overloads.py:880-901, 1768-1780). */
int rfb200_scale_lines(int precision, int complex_items, uint64_t nlines, uint64_t n,
                       const void *d_table, void *d_data, void *stream);

/* ---- housekeeping ----------------------------------------------------------------- */
/* Last error message of the calling thread ("" if none); cleared by rfb200_clear_error. */
const char *rfb200_last_error(void);
void rfb200_clear_error(void);
/* Drop all cached plans (device twiddle/chirp tables) of the current device. */
void rfb200_plan_cache_clear(void);
/* Stream used by the numba_* shims for the calling thread (cudaStream_t as void*; NULL is
 * the legacy default stream).  Default: a library-owned non-blocking stream per device,
 * restored by rfb200_use_library_stream(). */
void rfb200_set_stream(void *stream);
void rfb200_use_library_stream(void);
/* Number of kernel launches issued by this library since the last reset (all threads). */
uint64_t rfb200_launch_count(void);
void rfb200_launch_count_reset(void);
/* Launch trace: rfb200_launch_trace(1) starts (and clears) a record of the names of the kernels launched, (0) stops it;
 * rfb200_launch_trace_get returns them as one ';'-separated string (valid until the calling thread's next call).
 * Lets a benchmark name the kernels it timed from what actually ran. */
void rfb200_launch_trace(int enable);
const char *rfb200_launch_trace_get(void);
/* DST-II/III with ortho=true: 1 (default) reproduces the reference, which scales element
 * 0 (H:3033-3039, README.md:61-65); 0 scales element N-1 as SciPy does. */
void rfb200_set_dst_ortho_quirk(int enabled);
/* Number of numba_* calls that failed since the library was loaded.  Those entry points return void (as the reference's
 * do, _pocketfft_numba.cpp:31-223); a failed call prints its reason to stderr, leaves it readable through
 * rfb200_last_error, fills a HOST output array with NaN (element by element), and increments this counter -- it never
 * returns silently with an untouched output. */
uint64_t rfb200_failure_count(void);
/* Entries / bytes of the device table ("plan") cache.  Bounded: RFB200_PLAN_CACHE_ENTRIES (default 256) tables,
 * RFB200_PLAN_CACHE_MB (default 1024) MiB; least recently used tables are released beyond that (reference: the 16-entry
 * LRU plan cache, _pocketfft_hdronly.h:3169-3223). */
void rfb200_plan_cache_stats(uint64_t *entries, uint64_t *bytes);
/* Page-locks (cudaHostRegister) / releases a range of host memory the caller owns, e.g. the buffer of a NumPy array that
 * numba_* calls will use many times: such calls then move their data at the rate of pinned memory (no staging copy).  The
 * caller keeps the range alive and unpins it before freeing it -- which is why the library does not do this on its own
 * behind a cache: it cannot see free() / munmap().  0 on success, non-zero + rfb200_last_error otherwise. */
int rfb200_host_pin(void *ptr, uint64_t bytes);
int rfb200_host_unpin(void *ptr);
/* numba_dst with an explicit choice of the DST-II/III scaling under ortho (quirk: 1 = the reference's, which scales
 * element 0, _pocketfft_hdronly.h:3033-3039; 0 = SciPy's, element N-1), whatever rfb200_set_dst_ortho_quirk says.
 * rfb200_dst takes the same choice in its `ortho` argument: 0 off, 1 on (process-wide choice), 2 on/SciPy, 3 on/reference. */
void rfb200_host_dst(uint64_t ndim, const rfb200_array_record *ain, rfb200_array_record *aout,
                     rfb200_array_record *axes, uint64_t type, double fct, int ortho, int quirk);
/* Host-side view of the tile order of the fused four-step kernel (pow2_fused4_kernel.cuh; test aid, no GPU needed):
 * unit `unit` of 2*nstrips -> (step << 32) | strip with step 0 = A, 1 = B; -1 past the end or if lag > nstrips. */
int64_t rfb200_debug_fuse4_unit(uint32_t unit, uint32_t nstrips, uint32_t lag);
/* Measurement aid (csrc/probe.cu, tools/probe_strided_copy.py): a pure copy with the access pattern of the four-step column
 * passes over a rows x cols complex64 array of row pitch `pitch_bytes` (rows a multiple of 128): mode 0 strided -> strided,
 * 1 strided -> dense (out needs ceil(cols/32)*rows*256 bytes), 2 dense -> strided.  Returns 0 on success. */
int rfb200_debug_tile_copy(const void *in, void *out, uint64_t rows, uint64_t cols, int64_t pitch_bytes, int mode,
                           void *stream);
/* The same aid in general form: tiles0 x tiles1 CTAs of `threads` threads, each copying nrows segments of seg_bytes (in place
 * coordinates: segment r of tile (a, b) at a*seg_bytes + b*outer_stride + r*row_stride); at most 16 items of 8 bytes per thread;
 * smem_bytes of (unused) dynamic shared memory per CTA reproduce a transform kernel's occupancy. */
int rfb200_debug_seg_copy(const void *in, void *out, uint32_t nrows, uint32_t seg_bytes, int64_t row_stride, uint32_t tiles0,
                          uint32_t tiles1, int64_t outer_stride, uint32_t threads, uint32_t smem_bytes, void *stream);
/* Measurement aids of round 2 (csrc/probe.cu, tools/probe_l2_dsmem.py): sweep an L2-resident or DRAM-sized buffer (mode 0 read,
 * 1 copy) `reps` times from `ctas` CTAs; exchange `kb` KiB with every other CTA of a thread-block cluster through distributed
 * shared memory (mode 0 write, 1 read), cycles per CTA returned in `cycles`. */
int rfb200_debug_l2_sweep(const void *in, void *out, uint64_t bytes, int reps, int mode, uint32_t ctas, void *sink, void *stream);
int rfb200_debug_dsmem(uint32_t cluster, uint32_t nclusters, uint32_t kb, int reps, int mode, void *cycles, void *sink, void *stream);
/* Library version string. */
const char *rfb200_version(void);

#ifdef __cplusplus
}
#endif
#endif /* ROCKETFFT_B200_H */
