/* A NULL DEVICE for profiling the HOST side (Python + ctypes) of minidiff_b200 without a GPU.
 * Every entry point of include/minidiff_b200.h exists and returns success immediately; nothing is
 * computed, allocations are fake addresses, read-backs return zeros.  It is NOT a backend and is
 * never loaded by the package: scripts/host_profile.py builds it into a scratch copy of the
 * package under /tmp to measure microseconds of host work per op. */
#include <stddef.h>
#include <stdint.h>
#include <string.h>
static uintptr_t next_ptr = 0x100000000ull;
static uint64_t launches = 0;
int mdb_abi_version(void) { return 1; }
const char* mdb_last_error(void) { return ""; }
int mdb_device_count(int* c) { *c = 1; return 0; }
int mdb_init(int d) { (void)d; return 0; }
int mdb_shutdown(void) { return 0; }
int mdb_device_info(int* sm, size_t* tot, int* a, int* b) { *sm = 148; *tot = 180ull << 30; *a = 10; *b = 0; return 0; }
void* mdb_stream(void) { return 0; }
int mdb_sync(void) { return 0; }
int mdb_alloc(size_t bytes, void** out) { *out = (void*)next_ptr; next_ptr += (bytes + 511) & ~(size_t)511; return 0; }
int mdb_free(void* p) { (void)p; return 0; }
int mdb_empty_cache(void) { return 0; }
int mdb_mem_stats(size_t* a, size_t* b, size_t* c, uint64_t* d) { *a = *b = *c = 0; *d = 0; return 0; }
int mdb_host_alloc(size_t bytes, void** out) { (void)bytes; *out = 0; return 0; }
int mdb_host_free(void* p) { (void)p; return 0; }
int mdb_h2d(void* d, const void* s, size_t n) { (void)d; (void)s; (void)n; return 0; }
int mdb_d2h(void* d, const void* s, size_t n) { (void)s; memset(d, 0, n); return 0; }
int mdb_d2d(void* d, const void* s, size_t n) { (void)d; (void)s; (void)n; return 0; }
int mdb_prefetch_h2d(void* d, const void* s, size_t n) { (void)d; (void)s; (void)n; return 0; }
int mdb_prefetch_wait(void) { return 0; }
int mdb_event_create(void** e) { *e = (void*)1; return 0; }
int mdb_event_record(void* e) { (void)e; return 0; }
int mdb_event_elapsed_ms(void* a, void* b, float* ms) { (void)a; (void)b; *ms = 0.f; return 0; }
int mdb_event_destroy(void* e) { (void)e; return 0; }
uint64_t mdb_launch_count(void) { return launches; }
int mdb_graph_begin(void) { return 0; }
int mdb_graph_end(void** g) { *g = (void*)1; return 0; }
int mdb_graph_launch(void* g) { (void)g; return 0; }
int mdb_graph_info(void* g, uint64_t* n, size_t* b) { (void)g; *n = 0; *b = 0; return 0; }
int mdb_graph_destroy(void* g) { (void)g; return 0; }
int mdb_prof_enable(int on) { (void)on; return 0; }
int mdb_prof_read(int c, double* ms, uint64_t* n, double* w) { (void)c; *ms = 0; *n = 0; *w = 0; return 0; }
#define LAUNCH(name, ...) int name(__VA_ARGS__) { ++launches; return 0; }
LAUNCH(mdb_fill, const void* o, double v)
LAUNCH(mdb_copy, const void* o, const void* i)
LAUNCH(mdb_elementwise, int op, const void* o, int n, const void* i)
typedef struct { void* ptr; int dtype, ndim; long long shape[8], strides[8]; double imm; long long imm_i; } arr_t;
int mdb_elementwise_new(int op, arr_t* o, int n, const void* a, const void* b, const void* c) {
  (void)op; (void)n; (void)a; (void)b; (void)c; ++launches;
  if (!o->ptr) { o->ptr = (void*)next_ptr; next_ptr += 512; }
  return 0;
}
LAUNCH(mdb_reduce, int r, const void* o, const void* i, uint32_t m)
LAUNCH(mdb_elementwise_reduce, int op, const void* o, int n, const void* i, int a)
LAUNCH(mdb_gemm, const void* c, const void* a, const void* b, int acc)
LAUNCH(mdb_gemm_batched, const void* c, const void* a, const void* b)
int mdb_random_reset(uint64_t p) { (void)p; return 0; }
LAUNCH(mdb_random_bits, const void* o, uint64_t s, uint64_t off)
LAUNCH(mdb_randint, const void* o, int64_t lo, int64_t hi, uint64_t s, uint64_t off)
LAUNCH(mdb_binomial, const void* o, int64_t n, const void* p, uint64_t s, uint64_t off)
LAUNCH(mdb_permutation, const void* o, const void* b)
LAUNCH(mdb_arange, const void* o, double a, double b, int64_t c, int64_t d, int e)
LAUNCH(mdb_cumsum_f64, const void* o, const void* i)
LAUNCH(mdb_searchsorted_cdf, const void* o, const void* c, const void* u)
LAUNCH(mdb_index_offsets, const void* o, const void* i, int64_t e, int64_t s, int a)
LAUNCH(mdb_unravel_index, const void* o, const void* i, int nd, const int64_t* d)
LAUNCH(mdb_isin, const void* o, const void* e, const void* t, int inv)
int mdb_nonzero(const void* m, const void* o, int64_t* c) { (void)m; (void)o; *c = 0; return 0; }
LAUNCH(mdb_gemm_fused, const void* c, const void* a, const void* b, int acc, const void* bias, int relu, const void* m)
LAUNCH(mdb_gather_rows, const void* o, const void* s, const void* i)
LAUNCH(mdb_scatter_rows, const void* d, const void* s, const void* i, int add)
LAUNCH(mdb_random, const void* o, int n, uint64_t s, uint64_t off)
int mdb_gemm_tune(int f) { (void)f; return 0; }
int mdb_gemm_config(int f) { (void)f; return 0; }
int mdb_gemm_stats(uint64_t* c, int r) { (void)r; if (c) memset(c, 0, 64); return 0; }
int mdb_gemm_knob(int k, int v) { (void)k; (void)v; return 0; }
int mdb_gemm_last_plan(int* o) { memset(o, 0, 32); return 0; }
int mdb_comm_unique_id(void* id, const char* p) { (void)id; (void)p; return 0; }
int mdb_comm_init(int r, int w, const void* id, const char* p) { (void)r; (void)w; (void)id; (void)p; return 0; }
int mdb_comm_allreduce_f32(void* p, size_t n, int avg) { (void)p; (void)n; (void)avg; return 0; }
int mdb_comm_allreduce_multi_f32(void* const* p, const size_t* c, int n, int avg) { (void)p; (void)c; (void)n; (void)avg; return 0; }
int mdb_comm_wait(void) { return 0; }
uint64_t mdb_comm_last_seq(void) { return 0; }
int mdb_comm_wait_seq(uint64_t s) { (void)s; return 0; }
int mdb_comm_destroy(void) { return 0; }
