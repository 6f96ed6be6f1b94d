/*
 * minidiff_b200 -- C ABI of the B200-native array backend for minidiff.
 *
 * This is the drop-in boundary: plain pointers, sizes and POD descriptors, no torch / C++ types.
 * The reference (ahoynodnarb/minidiff) has NO native layer; its boundary is the table of NumPy
 * functions a backend plugin exports (minidiff/backend/numpy.py:14-206, copied into the
 * `minidiff.backend` namespace by minidiff/backend/__init__.py:80-85) plus a handful of raw-array
 * protocol methods (`.astype`, `+=` family, `__setitem__`; minidiff/tensor.py:105,269-379).
 * Each entry point below names the reference interface it stands behind.  The Python shim that
 * binds these with ctypes is minidiff_b200/backend/_lib.py; INTEGRATION.md shows the stub.
 *
 * Conventions
 *   - every function returns 0 on success, non-zero on failure; mdb_last_error() gives the text
 *     (the Python shim raises ValueError for MDB_EINVAL-class errors, RuntimeError otherwise,
 *     matching the exception types NumPy raises through the reference: SURVEY 8b "Errors").
 *   - all work is enqueued on ONE compute stream owned by the library (mdb_stream()); calls
 *     return immediately, only mdb_sync / mdb_d2h / mdb_item wait (reference semantics are
 *     synchronous eager; only as_numpy/item/repr observe values: SURVEY 8b "Threading").
 *   - strides are in ELEMENTS, may be 0 (broadcast view) or negative (flip view).
 */
#ifndef MINIDIFF_B200_H
#define MINIDIFF_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define MDB_MAX_DIMS 8
#define MDB_ABI_VERSION 1

/* status codes */
enum { MDB_OK = 0, MDB_EINVAL = 1, MDB_ECUDA = 2, MDB_ENOMEM = 3, MDB_ENOTSUP = 4, MDB_ECOMM = 5,
       MDB_EINDEX = 6 /* data-dependent index out of range: the shim raises IndexError */ };

/* element types (mirrors the dtype objects a backend exports: backend/numpy.py:188-200) */
typedef enum {
  MDB_BOOL = 0, MDB_U8 = 1, MDB_I8 = 2, MDB_I16 = 3, MDB_I32 = 4, MDB_I64 = 5,
  MDB_F32 = 6, MDB_F64 = 7, MDB_U16 = 8, MDB_U32 = 9, MDB_U64 = 10, MDB_F16 = 11
} mdb_dtype;

/* strided view of device memory == what `backend.tensor_class` instances carry
 * (shape/strides semantics of numpy.ndarray, the reference's tensor_class: backend/numpy.py:15-16) */
typedef struct {
  void*   ptr;                    /* address of element [0,...,0] (NULL for an immediate) */
  int32_t dtype;                  /* mdb_dtype */
  int32_t ndim;                   /* 0..MDB_MAX_DIMS */
  int64_t shape[MDB_MAX_DIMS];
  int64_t strides[MDB_MAX_DIMS];  /* elements */
  double  imm;                    /* value when ptr == NULL: an un-promoted Python scalar operand
                                     (reference passes them through unwrapped: wrapping.py:121-124) */
  int64_t imm_i;                  /* integer image of imm (exact for int64 immediates) */
} mdb_array;

/* ---- elementwise op ids: one per backend function (backend/numpy.py:19-95) ------------------ */
typedef enum {
  /* unary */
  MDB_OP_COPY = 0, MDB_OP_NEG, MDB_OP_ABS, MDB_OP_SIGN, MDB_OP_CEIL, MDB_OP_FLOOR,
  MDB_OP_SIN, MDB_OP_COS, MDB_OP_TAN, MDB_OP_SINH, MDB_OP_COSH, MDB_OP_TANH,
  MDB_OP_EXP, MDB_OP_LOG, MDB_OP_SQRT, MDB_OP_RECIP, MDB_OP_SQUARE, MDB_OP_LOGICAL_NOT,
  MDB_OP_INVERT, MDB_OP_ISNAN,
  MDB_OP_RELU,       /* where(x > 0, x, 0) in one pass: the forward of a user-level fused relu op */
  /* binary */
  MDB_OP_ADD = 32, MDB_OP_SUB, MDB_OP_MUL, MDB_OP_DIV, MDB_OP_POW, MDB_OP_MOD, MDB_OP_FLOORDIV,
  MDB_OP_MAXIMUM, MDB_OP_MINIMUM,
  MDB_OP_EQ, MDB_OP_NE, MDB_OP_GT, MDB_OP_GE, MDB_OP_LT, MDB_OP_LE,
  MDB_OP_AND, MDB_OP_OR, MDB_OP_XOR,
  /* ternary */
  MDB_OP_WHERE = 64, MDB_OP_CLIP,
  MDB_OP_FMA,        /* in0 + in1*in2, rounded as separate mul then add (== two backend calls) */
  /* fused backward forms (replace the call chains of ops/definitions.py grad lambdas; each is
     rounded step by step exactly like the unfused chain so results are bit-identical) */
  MDB_OP_SIN_BWD = 96,   /* g * cos(x)            definitions.py:378  */
  MDB_OP_COS_BWD,        /* g * (-sin(x))         definitions.py:313  */
  MDB_OP_EXP_BWD,        /* g * exp(x)            definitions.py:321  */
  MDB_OP_LOG_BWD,        /* g / x                 definitions.py:342  */
  MDB_OP_TANH_BWD,       /* g * (1/cosh(x)**2)    definitions.py:414  */
  MDB_OP_POW_BWD,        /* (g*p) * x**(p-1), p = immediate in2   definitions.py:509 */
  MDB_OP_DIV_BWD_Y,      /* g * (-x / y**2)       definitions.py:531  */
  MDB_OP_RELU_MASK_BWD,  /* g * (x > 0)           where-grad_y definitions.py:557 with greater */
  MDB_OP_POW_BWD_LIN,    /* (g*p) * x : POW_BWD for p == 2 (x**1 == x exactly), chosen by the library */
} mdb_op;

/* reductions (backend sum/mean/max/min/prod/any/all/argmax/argmin: backend/numpy.py:20-57) */
typedef enum {
  MDB_RED_SUM = 0, MDB_RED_MEAN, MDB_RED_MAX, MDB_RED_MIN, MDB_RED_PROD, MDB_RED_ANY, MDB_RED_ALL,
  MDB_RED_ARGMAX, MDB_RED_ARGMIN
} mdb_red;

/* ---- runtime ------------------------------------------------------------------------------- */
int         mdb_abi_version(void);
const char* mdb_last_error(void);
int         mdb_device_count(int* count);
int         mdb_init(int device);                 /* idempotent; creates the compute stream      */
int         mdb_shutdown(void);
int         mdb_device_info(int* sm_count, size_t* total_bytes, int* cc_major, int* cc_minor);
void*       mdb_stream(void);                     /* cudaStream_t of the compute stream          */
int         mdb_sync(void);

/* caching allocator (size-class free lists; stream-ordered reuse on the compute stream).
 * Stands behind every array a backend function returns + DeviceArray finalizers (SURVEY 8b
 * "Ownership"). */
int mdb_alloc(size_t bytes, void** out);
int mdb_free(void* ptr);
int mdb_empty_cache(void);
int mdb_mem_stats(size_t* in_use, size_t* cached, size_t* peak_in_use, uint64_t* n_device_allocs);
int mdb_host_alloc(size_t bytes, void** out);     /* pinned host memory for H2D/D2H staging      */
int mdb_host_free(void* ptr);

/* tensor_constructor / as_numpy / tensor_item (backend/numpy.py:15,161-163,204-206) */
int mdb_h2d(void* dst, const void* src, size_t bytes);         /* async if src is pinned          */
int mdb_d2h(void* dst, const void* src, size_t bytes);         /* waits for completion            */
int mdb_d2d(void* dst, const void* src, size_t bytes);
/* input pipeline: upload the NEXT batch from pinned host memory on a copy stream while the current
 * step computes.  The copy is ordered after all compute work enqueued before the call (the last
 * reader of dst); mdb_prefetch_wait() orders later compute after the most recent prefetch. */
int mdb_prefetch_h2d(void* dst, const void* src, size_t bytes);
int mdb_prefetch_wait(void);

/* timing on the compute stream (bench.py: CUDA events on the launching stream) */
int mdb_event_create(void** ev);
int mdb_event_record(void* ev);
int mdb_event_elapsed_ms(void* start, void* stop, float* ms);  /* syncs on `stop`                 */
int mdb_event_destroy(void* ev);
uint64_t mdb_launch_count(void);                  /* kernels launched by this library so far     */
/* CUDA graphs -- the device-side counterpart of the reference's caching.reuse_graph
 * (caching.py:15-65, topology.py:152-162: "this graph repeats, do the bookkeeping once").  Everything
 * launched on the compute stream between begin and end is captured; mdb_graph_launch replays it with
 * one launch.  Memory handed out while capturing is pinned to the graph (private pool) until
 * mdb_graph_destroy, so replays always find their buffers.  A capture must not synchronise, read back
 * (mdb_d2h) or upload (mdb_h2d); the profiler must be off. */
int mdb_graph_begin(void);
int mdb_graph_end(void** graph);
int mdb_graph_launch(void* graph);
int mdb_graph_info(void* graph, uint64_t* kernel_launches, size_t* pinned_bytes);
int mdb_graph_destroy(void* graph);
/* per-class device time of the library's own launches, measured with CUDA event pairs on the
 * compute stream around each public compute call while enabled.  cls: 0 elementwise (incl. copy /
 * fill), 1 reductions (incl. the fused un-broadcast form), 2 GEMM, 3 other.  `work` is the summed
 * ALGORITHMIC work of those calls: bytes (each distinct input element once at its un-broadcast
 * size + each output element once) for classes 0/1/3, flops (2*M*N*K) for class 2. */
int mdb_prof_enable(int on);                      /* turning on clears earlier records           */
int mdb_prof_read(int cls, double* total_ms, uint64_t* calls, double* work);

/* ---- compute ------------------------------------------------------------------------------- */
/* ones_like/zeros_like/full/full_like (backend/numpy.py:98-103) */
int mdb_fill(const mdb_array* out, double value);
/* copy / astype / ascontiguous / broadcast materialisation / `a[key] = v` on basic keys
 * (backend/numpy.py:29,63-65; tensor.py:376-379): out[...] = cast(in[...]) with broadcasting */
int mdb_copy(const mdb_array* out, const mdb_array* in);
/* every elementwise backend function; `out` may alias in[0] exactly (the `+=` family of
 * tensor.py:269-362).  Inputs broadcast against out's shape (stride 0 where stretched). */
int mdb_elementwise(int op, const mdb_array* out, int n_in, const mdb_array* in);
/* the same op with the operands passed as pointers to descriptors (what the shim's arrays cache) and
 * an output that is ALLOCATED by the call when out->ptr is NULL (C-contiguous; the address is written
 * back into out->ptr): one ABI crossing per backend function instead of alloc + marshal + launch */
int mdb_elementwise_new(int op, mdb_array* out, int n_in, const mdb_array* in0, const mdb_array* in1,
                        const mdb_array* in2);
/* sum/mean/max/min/prod/any/all/argmax/argmin over the axes whose bit is set in axis_mask;
 * `out` is given in keepdims form (same ndim as `in`, reduced extents 1). */
int mdb_reduce(int red, const mdb_array* out, const mdb_array* in, uint32_t axis_mask);
/* fused "un-broadcast": out (+)= sum over the stretched axes of op(in...) where out's extents
 * are 1 on the reduced axes (replaces grad-lambda -> md.unbroadcast -> `grad + new` of
 * topology.py:93-104 + definitions.py:157-183 with one pass). accumulate: 0 store, 1 add. */
int mdb_elementwise_reduce(int op, const mdb_array* out, int n_in, const mdb_array* in,
                           int accumulate);
/* matmul (backend/numpy.py:84) for 2-D operands of any row/column-major mix, fp32, 3xTF32 on
 * tcgen05 tensor cores when shapes allow, else an fp32 CUDA-core kernel.
 * C = A@B (accumulate=0) or C += A@B (accumulate=1, the in-place form of topology.py:101-104). */
int mdb_gemm(const mdb_array* c, const mdb_array* a, const mdb_array* b, int accumulate);
/* np.matmul for stacked and/or float64 operands: c[..., M, N] = a[..., M, K] @ b[..., K, N], all three
 * given with the SAME rank (leading batch axes of a / b may have extent 1 = broadcast), float32 or
 * float64, any strides; every matrix of the batch in ONE launch of a CUDA-core kernel (O(M*N) memory). */
int mdb_gemm_batched(const mdb_array* c, const mdb_array* a, const mdb_array* b);
int mdb_gemm_tune(int flags);                     /* kernel tuning switches for A/B measurements   */
int mdb_gemm_config(int force_path);              /* 0 auto, 1 CUDA-core kernel only, 2 tensor-core
                                                     kernel only (tests) */
/* which kernel the GEMM launches took since the last reset: counts[MDB_GEMM_NPATHS], indexed by the
 * enum below (parity tests at BASELINE dims assert that every GEMM of the step ran on the CTA-pair
 * tcgen05 kernel the benchmark measures) */
enum { MDB_GEMM_PATH_SIMT = 0,        /* fp32 CUDA-core kernel (small / odd shapes)                 */
       MDB_GEMM_PATH_TC_SINGLE = 1,   /* tcgen05, one CTA per 128x128 tile, TMA reads operands in place */
       MDB_GEMM_PATH_TC_PRESPLIT = 2, /* tcgen05 single-CTA after a hi/lo gather pre-pass           */
       MDB_GEMM_PATH_TC_PAIR = 3,     /* tcgen05 cta_group::2, 256x256 tiles, whole tiles per pair  */
       MDB_GEMM_PATH_TC_PAIR_STREAMK = 4, /* same kernel, k-range split across pairs (stream-K)     */
       MDB_GEMM_PATH_SIMT_BATCHED = 5,    /* batched / float64 CUDA-core kernel (mdb_gemm_batched)    */
       MDB_GEMM_NPATHS = 8 };
int mdb_gemm_stats(uint64_t* counts, int reset);
/* measurement knobs of the CTA-pair kernel's planner (-1 = automatic): tile order, L2 eviction hints
 * (0 none, 1 evict_first, 2 evict_last), stream-K (0 never, 1 whenever legal) */
enum { MDB_GEMM_KNOB_RASTER = 0, MDB_GEMM_KNOB_GROUP = 1, MDB_GEMM_KNOB_HINT_A = 2, MDB_GEMM_KNOB_HINT_B = 3,
       MDB_GEMM_KNOB_HINT_C = 4, MDB_GEMM_KNOB_STREAMK = 5, MDB_GEMM_KNOB_L2_BUDGET_MB = 6,
       MDB_GEMM_KNOB_MAX_CLUSTERS = 7 /* cap on co-resident CTA pairs: leaves SMs to a concurrent NCCL kernel */,
       MDB_GEMM_KNOB_SPLIT = 8 /* CTA-pair kernel: 0 and -1 (default) = 3xTF32 (three TF32 MMAs per product); 1 = "fast" split: one TF32 MMA + the two
                                   cross terms as BF16 MMAs (8 instead of 12 tensor-core instructions per k-block; rms error 1.4e-6 instead
                                   of 0.5e-6 of the result's rms, maximum ~1e-5: at the edge of the GEMM tolerance, hence opt-in) */,
       MDB_GEMM_KNOB_CHUNK = 9 /* CTA-pair kernel: k-blocks (32 K each) accumulated in tensor memory before the partial sum is promoted to
                                   fp32 registers (the tensor core truncates on accumulate); -1 = default */,
       MDB_GEMM_KNOB_RZ_GAIN = 10 /* CTA-pair kernel: compensation of that truncation's bias per MMA instruction, in units of 1e-10
                                     (0 = off, -1 = default) */ };
int mdb_gemm_knob(int knob, int value);
/* plan of the most recent CTA-pair launch: clusters, raster, group, dp_tiles, sk_clusters, sk_share,
 * hints (100*A + 10*B + C), tiles */
int mdb_gemm_last_plan(int* out8);
/* GEMM with a fused epilogue, the device op behind a user-defined stateful op (the reference's
 * create_stateful_op_func / OpClass, ops/wrapping.py:47-76,181-217 -- e.g. linear_relu):
 *   C (+)= mask( relu?( A@B + bias? ) )      bias: fp32 [N] or NULL;  relu: where(v > 0, v, 0);
 *   mask_src: fp32 [M, N] row-major or NULL, multiplies by (mask_src > 0) -- the ReLU backward
 *   `grad * (y > 0)` applied where the gradient GEMM produces it.
 * Each step rounds like the separate backend call it replaces (add / where / multiply), so the
 * result is bit-identical to matmul -> add -> where (resp. matmul -> multiply by the mask).
 * Returns MDB_ENOTSUP when the problem cannot run on the CTA-pair tensor-core kernel (callers
 * then issue the unfused chain). */
int mdb_gemm_fused(const mdb_array* c, const mdb_array* a, const mdb_array* b, int accumulate,
                   const mdb_array* bias, int relu, const mdb_array* mask_src);

/* integer-array indexing (getitem / `a[key] = v` / index_add with array keys:
 * backend/numpy.py:73-75,105; tensor.py:376-379) over the leading axis:
 *   gather : out[i, ...] = src[idx[i], ...]
 *   scatter: dst[idx[i], ...] = src[i, ...]  (add=0)   or   += (add=1, duplicates accumulate,
 *            the np.add.at semantics getitem_grad relies on: ops/definitions.py:186-189)
 * idx is a contiguous int64 vector; negative entries wrap.  src of scatter may broadcast
 * (stride 0). */
/* When the indexed array's shape[0] is MDB_ROWS_ARE_OFFSETS, idx holds ELEMENT OFFSETS that
 * mdb_index_offsets has already validated (signed: views with negative strides work); they are used
 * as they are, without wrapping or clamping. */
#define MDB_ROWS_ARE_OFFSETS ((int64_t)1 << 62)
int mdb_gather_rows(const mdb_array* out, const mdb_array* src, const mdb_array* idx);
int mdb_scatter_rows(const mdb_array* dst, const mdb_array* src, const mdb_array* idx, int add);

/* off (+)= wrap(idx) * stride for an integer index array `idx` (any integer dtype, given broadcast to
 * off's shape) over an axis of `extent` elements: negative indices count from the end, anything outside
 * [-extent, extent) makes the call return MDB_EINDEX with NumPy's message ("index N is out of bounds
 * for axis with size E": backend/numpy.py:73-75 raises IndexError there).  The check reads one flag
 * back (stream sync); inside a CUDA-graph capture it cannot, and offending indices are clamped. */
int mdb_index_offsets(const mdb_array* off, const mdb_array* idx, int64_t extent, int64_t stride, int accumulate);
/* stream compaction: flat positions of the non-zero entries of a contiguous bool mask, in order, into
 * out_indices (int64, room for every element); *count = how many (argwhere, ops/definitions.py:279-290;
 * boolean-mask getitem / setitem, backend/numpy.py:73-75).  Synchronises (the size is data-dependent). */
int mdb_nonzero(const mdb_array* mask, const mdb_array* out_indices, int64_t* count);
/* unravel_index (tensor.py:509-515): out is int64 [ndim, n]; out-of-range -> MDB_EINVAL like NumPy's ValueError */
int mdb_unravel_index(const mdb_array* out, const mdb_array* indices, int ndim, const int64_t* dims);
/* isin (tensor.py:503-507): out[i] = (elements[i] in test) != invert; integer pairs compare as int64,
 * anything else as float64 */
int mdb_isin(const mdb_array* out, const mdb_array* elements, const mdb_array* test, int invert);

/* counter-based RNG on device (rand / randn / randint / binomial / permutation / choice:
 * backend/numpy.py:131-138, tensor.py:608-659); Philox4x32-10, (seed, offset) select the stream */
/* offset == MDB_RNG_DEVICE_OFFSET: use (and advance) the library's DEVICE-RESIDENT stream position, so that
 * a captured CUDA graph draws fresh numbers on every replay; mdb_random_reset sets that position (seed()) */
#define MDB_RNG_DEVICE_OFFSET UINT64_MAX
int mdb_random_reset(uint64_t position);
int mdb_random(const mdb_array* out, int normal, uint64_t seed, uint64_t offset);
int mdb_random_bits(const mdb_array* out, uint64_t seed, uint64_t offset);            /* raw 32-bit words */
int mdb_randint(const mdb_array* out, int64_t low, int64_t high, uint64_t seed, uint64_t offset);
int mdb_binomial(const mdb_array* out, int64_t trials, const mdb_array* p, uint64_t seed, uint64_t offset);
/* out = uniformly random permutation of 0..n-1 (device bitonic sort of (random word, i) keys) */
int mdb_permutation(const mdb_array* out, const mdb_array* bits);
/* arange (backend/numpy.py:127): out[i] = start + i*step; integral != 0 computes in int64 exactly */
int mdb_arange(const mdb_array* out, double start, double step, int64_t istart, int64_t istep, int integral);
/* weighted choice: inclusive float64 scan of the weights, then out[i] = searchsorted(cdf / cdf[-1], u[i], "right") */
int mdb_cumsum_f64(const mdb_array* out, const mdb_array* in);
int mdb_searchsorted_cdf(const mdb_array* out, const mdb_array* cdf, const mdb_array* u);

/* ---- data-parallel exchange (config 4; no counterpart in the reference: SURVEY 2.1) --------- */
int mdb_comm_unique_id(void* id128, const char* nccl_lib_path);          /* rank 0               */
int mdb_comm_init(int rank, int world, const void* id128, const char* nccl_lib_path);
int mdb_comm_allreduce_f32(void* ptr, size_t count, int average);        /* on the comm stream,
                                                                            ordered after compute */
/* n buffers averaged in ONE NCCL launch (ncclGroupStart/End): small gradients ride with large ones */
int mdb_comm_allreduce_multi_f32(void* const* ptrs, const size_t* counts, int n, int average);
int mdb_comm_wait(void);                          /* compute stream waits for the comm stream    */
/* per-exchange completion: every mdb_comm_allreduce_f32 gets the next sequence number; the compute
 * stream can wait for ONE of them (update that parameter while later gradients are still in flight) */
uint64_t mdb_comm_last_seq(void);
int mdb_comm_wait_seq(uint64_t seq);
int mdb_comm_destroy(void);

#ifdef __cplusplus
}
#endif
#endif /* MINIDIFF_B200_H */
