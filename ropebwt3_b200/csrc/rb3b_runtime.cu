/*
 * rb3b_runtime.cu -- device selection, stream, memory pool, error text, counters.
 * There is no CPU fallback anywhere in this library: if no CUDA device can be
 * initialised every entry point fails with RB3B_ENODEV.
 */
#include <stdarg.h>
#include <string.h>
#include <map>
#include <string>
#include "rb3b_internal.cuh"

cudaStream_t rb3b_stream = 0;
int64_t rb3b_seg_len = 0;         /* target LF-walk segment length ("seg_len") */
int64_t rb3b_rank_variant = 0;    /* 0: LDG.128 per lane, 1: cp.async.bulk (TMA) staged */

static int g_inited = 0, g_device = 0, g_own_stream = 0;
static cudaStream_t g_my_stream = 0;
static thread_local char g_err[1024] = "";
static std::map<std::string, int64_t> g_stats;
static std::map<std::string, int64_t> g_params;

/* ---- scratch arena ---- */
#include <vector>
struct Chunk { char *p; size_t cap; };
static std::vector<Chunk> g_chunks;
static size_t g_chunk_i = 0, g_chunk_off = 0, g_used = 0, g_high = 0;
static int g_depth = 0;

void *rb3b_arena_alloc(size_t bytes)
{
	bytes = (bytes + 511) & ~(size_t)511;
	for (;;) {
		if (g_chunk_i < g_chunks.size() && g_chunk_off + bytes <= g_chunks[g_chunk_i].cap) {
			void *p = g_chunks[g_chunk_i].p + g_chunk_off;
			g_chunk_off += bytes; g_used += bytes;
			if (g_used > g_high) g_high = g_used;
			return p;
		}
		if (g_chunk_i + 1 < g_chunks.size()) { ++g_chunk_i; g_chunk_off = 0; continue; }
		Chunk c;
		c.cap = bytes > ((size_t)64 << 20) ? bytes + (bytes >> 2) : (size_t)64 << 20;
		if (cudaMalloc((void**)&c.p, c.cap) != cudaSuccess) { cudaGetLastError(); return 0; }
		g_chunks.push_back(c);
		g_chunk_i = g_chunks.size() - 1; g_chunk_off = 0;
	}
}

void rb3b_arena_enter(void) { ++g_depth; }

void rb3b_arena_leave(void)
{
	if (--g_depth > 0) return;
	if (g_chunks.size() > 1) { /* fold the chunks into one that fits everything the last call needed */
		cudaStreamSynchronize(rb3b_stream);
		for (size_t i = 0; i < g_chunks.size(); ++i) cudaFree(g_chunks[i].p);
		g_chunks.clear();
		Chunk c;
		c.cap = g_high + (g_high >> 2) + ((size_t)16 << 20);
		if (cudaMalloc((void**)&c.p, c.cap) == cudaSuccess) g_chunks.push_back(c); else cudaGetLastError();
	}
	g_chunk_i = 0; g_chunk_off = 0; g_used = 0;
}

int rb3b_fail(int code, const char *fmt, ...)
{
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(g_err, sizeof(g_err), fmt, ap);
	va_end(ap);
	return code;
}

int64_t rb3b_n_launch = 0;
static cudaEvent_t g_ev[T_COUNT][2];
static int g_ev_ok = 0, g_ev_pending[T_COUNT];
static const char *g_ev_name[T_COUNT] = { "us_prep", "us_walk_first", "us_walk_fix", "us_merge", "us_scatter", "us_bwt" };

void rb3b_tic(int id)
{
	if (!g_ev_ok) {
		for (int i = 0; i < T_COUNT; ++i) { cudaEventCreate(&g_ev[i][0]); cudaEventCreate(&g_ev[i][1]); g_ev_pending[i] = 0; }
		g_ev_ok = 1;
	}
	if (g_ev_pending[id]) rb3b_tflush();
	cudaEventRecord(g_ev[id][0], rb3b_stream);
}

void rb3b_toc(int id) { cudaEventRecord(g_ev[id][1], rb3b_stream); g_ev_pending[id] = 1; }

void rb3b_tflush(void)
{
	for (int i = 0; i < T_COUNT; ++i)
		if (g_ev_ok && g_ev_pending[i]) {
			float ms = 0;
			cudaEventSynchronize(g_ev[i][1]);
			if (cudaEventElapsedTime(&ms, g_ev[i][0], g_ev[i][1]) == cudaSuccess) rb3b_stat_add(g_ev_name[i], (int64_t)(ms * 1000.0f + 0.5f));
			g_ev_pending[i] = 0;
		}
}

void rb3b_stat_set(const char *key, int64_t v) { g_stats[key] = v; }
void rb3b_stat_add(const char *key, int64_t v) { g_stats[key] += v; }

extern "C" const char *rb3b_last_error(void) { return g_err; }
extern "C" const char *rb3b_version(void) { return "rb3b200-0.1 (ropebwt3 3.10-r281 merge path, sm_100a)"; }

extern "C" int rb3b_init(int device)
{
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0)
		return rb3b_fail(RB3B_ENODEV, "no CUDA device (%s); this library has no CPU fallback", e != cudaSuccess ? cudaGetErrorString(e) : "count=0");
	if (device < 0 || device >= n) return rb3b_fail(RB3B_EINVAL, "device %d out of range [0,%d)", device, n);
	CK(cudaSetDevice(device));
	g_device = device;
	if (!g_my_stream) CK(cudaStreamCreateWithFlags(&g_my_stream, cudaStreamNonBlocking));
	if (!g_own_stream) rb3b_stream = g_my_stream;
	cudaMemPool_t pool;
	CK(cudaDeviceGetDefaultMemPool(&pool, device));
	uint64_t thr = UINT64_MAX; /* keep freed scratch in the pool: merges reuse it */
	CK(cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &thr));
	g_inited = 1;
	return RB3B_OK;
}

int rb3b_ensure_init(void)
{
	if (g_inited) { cudaSetDevice(g_device); return RB3B_OK; }
	return rb3b_init(0);
}

extern "C" int rb3b_set_stream(void *s)
{
	TRY(rb3b_ensure_init());
	if (s) { rb3b_stream = (cudaStream_t)s; g_own_stream = 1; }
	else { rb3b_stream = g_my_stream; g_own_stream = 0; }
	return RB3B_OK;
}

extern "C" int rb3b_sync(void)
{
	TRY(rb3b_ensure_init());
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}

extern "C" int rb3b_set_param(const char *key, int64_t value)
{
	if (!strcmp(key, "seg_len")) { if (value != 0 && value < 16) return rb3b_fail(RB3B_EINVAL, "seg_len must be >= 16 (0 = automatic)"); rb3b_seg_len = value; }
	else if (!strcmp(key, "rank_variant")) rb3b_rank_variant = value;
	else g_params[key] = value;
	return RB3B_OK;
}

int64_t rb3b_get_param(const char *key, int64_t dflt)
{
	std::map<std::string, int64_t>::iterator it = g_params.find(key);
	return it == g_params.end() ? dflt : it->second;
}

extern "C" int64_t rb3b_get_stat(const char *key)
{
	if (!strcmp(key, "kernel_launches")) return rb3b_n_launch;
	if (!strcmp(key, "seg_len")) return rb3b_seg_len;
	if (!strcmp(key, "reset")) { g_stats.clear(); rb3b_n_launch = 0; return 0; }
	std::map<std::string, int64_t>::iterator it = g_stats.find(key);
	return it == g_stats.end() ? -1 : it->second;
}

/* Largest batch (symbols) the device can take next to an index that will grow by it: the suffix sorter needs ~44 bytes
 * per symbol; the rank phase of a batch beyond 2^29 rows 8 (LF) + 8 (rows) + 1 (symbols) + 8 (kseq) + 8 (ka) + 1 (list
 * nodes) + 16 (bucketed scatter) + sort scratch ~ 60; the scratch arena keeps its high-water mark; both ping-pong halves
 * of the index ~2 x 1 byte per symbol; 32-bit suffix array. */
extern "C" int64_t rb3b_max_batch_symbols(int64_t index_symbols)
{
	size_t fr = 0, tot = 0;
	if (rb3b_ensure_init() != RB3B_OK) return -1;
	if (cudaMemGetInfo(&fr, &tot) != cudaSuccess) return -1;
	int64_t avail = (int64_t)tot - (int64_t)(tot >> 4) - 2 * index_symbols; /* what this process already holds counts as available */
	int64_t n = avail / 72;
	const int64_t cap = (1LL << 32) - 4096;
	if (n > cap) n = cap;
	return n > 0 ? n : 0;
}

extern "C" void *rb3b_dev_alloc(int64_t bytes)
{
	void *p = 0;
	if (rb3b_ensure_init() != RB3B_OK) return 0;
	if (cudaMalloc(&p, bytes > 0 ? bytes : 1) != cudaSuccess) { rb3b_fail(RB3B_ENOMEM, "cudaMalloc(%lld) failed", (long long)bytes); return 0; }
	return p;
}

extern "C" void rb3b_dev_free(void *p) { if (p) cudaFree(p); }

extern "C" int rb3b_h2d(void *dst, const void *src, int64_t bytes)
{
	TRY(rb3b_ensure_init());
	CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}

extern "C" int rb3b_d2h(void *dst, const void *src, int64_t bytes)
{
	TRY(rb3b_ensure_init());
	CK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}
