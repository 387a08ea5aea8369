/* hbfill.h -- C ABI of libhbfill.so, the B200-native fill engine behind histbook's Hist/Book API.
 *
 * The reference (scikit-hep/histbook 1.2.5) is pure Python + numpy and has no FFI of its own; the only
 * alternate-backend precedent is the Spark dispatch in Hist.fill / Book.fill (histbook/hist.py:355-359,
 * histbook/book.py:557-565).  The entry points below are what a binding for the fill hot path needs, and
 * each one names the reference code it stands in for:
 *
 *   hb_program_compile / hb_program_load   Fillable.fields (histbook/fill.py:41-59): build the instruction
 *                                          program once per Hist/Book -- here: NVRTC-compile one fused
 *                                          sm_100a kernel generated from the goal trees
 *   hb_arena_create / _clear / _clone      Hist._prefill (histbook/hist.py:381-390), Hist.clear (:93-95),
 *                                          Hist.copy / copyonfill (:80-91): float64 bin contents
 *   hb_fill                                Fillable._fill (histbook/fill.py:85-162) + Hist._postfill
 *                                          (histbook/hist.py:392-504): evaluate, index, scatter-accumulate
 *   hb_arena_read / hb_arena_write         every reader / writer of Hist._content (histbook/proj.py:498-640,
 *                                          histbook/hist.py:731,741): host copies of device-resident bins
 *   hb_arena_axpy / hb_arena_scale         Hist.__add__/__iadd__/__mul__/__imul__ (histbook/hist.py:506-607)
 *   hb_keyset_*                            numpy.unique in histbook_groupby / histbook_groupbin
 *                                          (histbook/calc/__init__.py:149-201)
 *
 * Conventions: plain C types only; every function returns 0 on success and a non-zero status otherwise,
 * with a thread-local message available from hb_last_error().  The library never keeps a column pointer
 * after hb_fill returns and never frees caller memory.  Handles are owned by the caller.
 * There is no CPU fallback: without a CUDA device every compute entry point fails with HB_ERR_CUDA.
 */
#ifndef HBFILL_H
#define HBFILL_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define HB_OK 0
#define HB_ERR_CUDA 1       /* CUDA runtime / driver error, or no device */
#define HB_ERR_NVRTC 2      /* kernel compilation failed (log in hb_last_error) */
#define HB_ERR_ARG 3        /* invalid argument */
#define HB_ERR_OVERFLOW 4   /* a device-side table is full (key sets) */

#define HB_MAX_COLS 32
#define HB_MAX_HISTS 128
#define HB_MAX_AUX 64
#define HB_MAX_WIN 128     /* run-time integer parameters of a launch (window geometry, fixed-point scales) */

/* column dtypes (numpy names) */
enum hb_dtype {
    HB_F64 = 0, HB_F32 = 1, HB_I64 = 2, HB_I32 = 3, HB_I16 = 4, HB_I8 = 5,
    HB_U64 = 6, HB_U32 = 7, HB_U16 = 8, HB_U8 = 9, HB_BOOL = 10
};

typedef struct hb_program hb_program;   /* a loaded fill kernel */
typedef struct hb_arena hb_arena;       /* float64 bin contents on the device */

typedef struct {
    const void* ptr;        /* first element (host or device memory); ignored for scalars */
    int32_t dtype;          /* enum hb_dtype */
    int32_t location;       /* 0 = host memory, 1 = device memory of the current device */
    int32_t is_scalar;      /* 1 = broadcast scalar, value in scalar_bits */
    int32_t reserved;
    uint64_t scalar_bits;   /* raw little-endian bits of the scalar, in `dtype` */
} hb_column;

typedef struct {
    int64_t events;         /* events processed */
    int64_t h2d_bytes;      /* bytes copied host -> device for this call */
    int32_t launches;       /* fill-kernel launches */
    int32_t grid;           /* blocks per launch */
    int32_t blocks_per_sm;  /* resident blocks per SM the grid was sized for */
    int32_t reserved;
    double ms_kernel;       /* device time of the launches (only with HB_FILL_TIMED) */
} hb_stats;

#define HB_FILL_TIMED 1     /* bracket the launches with CUDA events and synchronise (fills hb_stats.ms_kernel) */
#define HB_FILL_CONCURRENT 4   /* hb_fill_multi: one stream per part instead of one launch after the other */

const char* hb_last_error(void);
int hb_version(void);
int hb_nvrtc_version(void);                               /* e.g. 12090; -1 if NVRTC cannot be loaded */

/* device ------------------------------------------------------------------------------------------------ */
int hb_init(int device);                                 /* select the device (idempotent), load the driver */
int hb_device_info(int* sm_count, int* max_smem_optin, int64_t* free_bytes, int64_t* total_bytes);
int hb_set_stream(void* cuda_stream);                    /* compute stream for all later calls (default: stream 0) */
int hb_sync(void);
/* Stream order: work queued on `signaler` so far completes before work queued on `waiter` from now on (NULL = the
 * engine's compute stream).  Brackets a fill whose device columns were produced on another stream. */
int hb_stream_wait(void* waiter, void* signaler);
int hb_set_chunk_events(int64_t events);                 /* staging chunk for host-resident columns */

/* programs ---------------------------------------------------------------------------------------------- */
/* NVRTC only -- needs no GPU.  `header` is the text of hb_template.cuh, made available as
 * #include "hb_template.cuh".  On success *cubin is malloc'ed (release with hb_free). */
int hb_program_compile(const char* source, const char* header, int lineinfo,
                       void** cubin, size_t* cubin_size);
int hb_program_load(const void* cubin, size_t cubin_size, const char* kernel_name,
                    int threads, int smem_bytes, hb_program** out);
int hb_program_destroy(hb_program* prog);
/* events one block handles per round of the kernel's grid-stride loop (default 4 x threads; the sorted-book kernels
 * of codegen.py use HB_EPT x threads): hb_fill sizes the grid with it */
int hb_program_set_tile(hb_program* prog, int tile_events);
/* size the grid for at most `blocks` resident blocks per SM (0 = as many as fit): the parts of a book that was split
 * into several programs share the SMs (hb_fill_multi) */
int hb_program_set_max_blocks_per_sm(hb_program* prog, int blocks);
int hb_program_info(const hb_program* prog, int* regs, int* static_smem, int* blocks_per_sm);
void hb_free(void* p);

/* arenas ------------------------------------------------------------------------------------------------ */
int hb_arena_create(int64_t ndoubles, hb_arena** out);   /* zero-filled */
int hb_arena_destroy(hb_arena* a);
int hb_arena_clear(hb_arena* a);
int hb_arena_clone(const hb_arena* src, hb_arena** out);
int hb_arena_resize(hb_arena* a, int64_t ndoubles);      /* keeps the common prefix, zero-fills growth */
int hb_arena_read(const hb_arena* a, int64_t offset, int64_t count, double* host_out);
int hb_arena_write(hb_arena* a, int64_t offset, int64_t count, const double* host_in);
int hb_arena_axpy(hb_arena* dst, const hb_arena* src, double alpha);   /* dst += alpha * src */
/* dst[dst_off, dst_off + count) += alpha * src[src_off, src_off + count): slab slots of group-axis histograms */
int hb_arena_axpy_at(hb_arena* dst, int64_t dst_off, const hb_arena* src, int64_t src_off, int64_t count, double alpha);
int hb_arena_scale(hb_arena* dst, double alpha);
/* ONE launch copying nseg ranges src[i][src_off[i] .. + count[i]) to dst + dst_off[i] (doubles; dst = device memory): the
 * snapshot of all histograms of a book -- group-axis slabs slot by slot, in the merged key order -- that the rank merge
 * reduces (Hist.__add__ across ranks, histbook/hist.py:506-542) */
int hb_arena_snapshot(int nseg, hb_arena* const* src, const int64_t* src_off, const int64_t* dst_off, const int64_t* count, void* dst);                       /* dst *= alpha */
/* Hist.project (histbook/proj.py:227-296) on the device: sum the content over the fixed axes with keep[k] == 0.
 * `shape` = bins per fixed axis (ndim <= 8), ncol = size of the last (column) dimension, which is always kept.
 * *out is a new arena of prod(kept shape) * ncol doubles. */
int hb_arena_project(const hb_arena* src, int ndim, const int64_t* shape, int ncol, const int32_t* keep, hb_arena** out);
/* Hist.select on the device (histbook/proj.py:470-496: `content[slc].copy()` with a slice on ONE axis, the only array work
 * of `_selectaxis`): the arena is viewed as [outer][len][inner]; *out is a new arena [outer][count][inner] whose element k
 * along the cut axis is element start + k * step of the source. */
int hb_arena_slice(const hb_arena* src, int64_t outer, int64_t len, int64_t inner, int64_t start, int64_t step, int64_t count, hb_arena** out);
/* Hist.table on the device (histbook/proj.py:577-627 `handlearray`): *out is a new arena of nrows x ncolumns doubles with, per
 * bin, [count, err(count)] (flags & 1, & 2), [effcount] (flags & 4) and per profile j [mean, err(mean)] -- the columns in the
 * reference's order, computed with the same IEEE operations (bit-identical).  sumw / sumw2 / sumwx[j] / sumwx2[j] are column
 * numbers inside a row of ncol doubles; sumw2 = -1 for an unweighted histogram.  `normalized` tables stay on the host. */
int hb_arena_table(const hb_arena* src, int64_t nrows, int ncol, int sumw, int sumw2, int flags, int nprof,
                   const int32_t* sumwx, const int32_t* sumwx2, hb_arena** out);
void* hb_arena_ptr(const hb_arena* a);                   /* device pointer (for NCCL / DLPack wrapping) */
int64_t hb_arena_size(const hb_arena* a);

/* fill -------------------------------------------------------------------------------------------------- */
/* One pass over n events: every histogram of the program is updated in place.  Device columns are used
 * where they lie; host columns are streamed through double-buffered staging.  Returns after the work is
 * enqueued (device columns) or after the last host chunk has been copied (host columns). */
int hb_fill(hb_program* prog, int ncols, const hb_column* cols, int64_t n,
            int nhists, hb_arena* const* hists, int naux, const void* const* aux,
            int nwin, const int32_t* win, int flags, hb_stats* stats);

/* A book split into several programs ("parts": more histograms than one kernel's parameter block or shared memory
 * holds, or -- big books -- two half-books whose kernels share every SM so that one's arithmetic overlaps the other's
 * shared-memory traffic): ONE pass over the host columns (every chunk is staged once), every part's kernel launched
 * on its own stream per chunk, all joined on the compute stream.  Stands in for the single loop over all histograms
 * of Book.fill (histbook/book.py:566-586).  `cols` is the union of the parts' columns; colmap[i] = index in `cols`
 * of the part's column slot i (NULL: identity). */
typedef struct {
    hb_program* prog;
    int32_t ncols;
    const int32_t* colmap;
    int32_t nhists;
    hb_arena* const* hists;
    hb_arena* const* exact;       /* deterministic mode (see hb_fill_exact), or NULL */
    int32_t naux;
    const void* const* aux;
    int32_t nwin;
    const int32_t* win;
} hb_launch;
int hb_fill_multi(int nlaunch, const hb_launch* launches, int ncols, const hb_column* cols, int64_t n, int flags, hb_stats* stats);
/* worker threads that copy pageable host columns into the page-locked bounce ring (default: 12, or 3/4 of the cores on small hosts;
 * environment HB_COPY_THREADS) */
int hb_set_copy_threads(int n);

/* Deterministic mode.  hb_fill_exact is hb_fill with, per histogram, a second arena (`exact[h]`, may be NULL) that
 * holds two 192-bit fixed-point accumulators (upper and lower exponent window, 2 x 3 64-bit words) per (row, float64 slot): kernels generated with the
 * "deterministic" option add every float64 term there with integer atomics instead of float64 REDs.
 * hb_fx_finalize then rounds each accumulator to float64 once, adds it to the columns named by dest_mask[slot]
 * (bit c = column c) of `content`, and zeroes it.  k[slot] = 1075 + binary exponent of the accumulator's least
 * significant bit (the same run-time integer the fill kernel received).  Stands in for the per-column
 * numpy.add.at of Hist._postfill (histbook/hist.py:439-456), whose sequential float64 sum it replaces by the
 * correctly rounded exact sum. */
int hb_fill_exact(hb_program* prog, int ncols, const hb_column* cols, int64_t n,
                  int nhists, hb_arena* const* hists, hb_arena* const* exact, int naux, const void* const* aux,
                  int nwin, const int32_t* win, int flags, hb_stats* stats);
int hb_fx_finalize(hb_arena* content, hb_arena* exact, int64_t nrows, int ncol, int nslots,
                   const int32_t* k, const uint64_t* dest_mask);

/* key sets (group axes) --------------------------------------------------------------------------------- */
/* A device-side set of 64-bit keys (raw bits of float64 / int64 group keys), capacity a power of two. */
typedef struct hb_keyset hb_keyset;
int hb_keyset_create(int capacity, hb_keyset** out);
int hb_keyset_destroy(hb_keyset* ks);
int hb_keyset_clear(hb_keyset* ks);
void* hb_keyset_ptr(const hb_keyset* ks);                /* device table, passed to discovery kernels via aux[] */
int hb_keyset_read(const hb_keyset* ks, uint64_t* keys_out, int max_keys, int* nkeys, int* overflowed);

/* utilities ---------------------------------------------------------------------------------------------- */
int hb_flush_l2(void);                                   /* overwrite a buffer larger than L2, then stream another one through it (benchmark hygiene) */
int hb_memcpy_h2d(void* dst_device, const void* src_host, size_t bytes);
int hb_memcpy_d2h(void* dst_host, const void* src_device, size_t bytes);
int hb_device_malloc(size_t bytes, void** out);
int hb_device_free(void* p);

#ifdef __cplusplus
}
#endif
#endif /* HBFILL_H */
