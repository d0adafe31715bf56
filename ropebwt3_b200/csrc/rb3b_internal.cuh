/*
 * rb3b_internal.cuh -- shared definitions of the B200-native BWT-merge engine.
 *
 * Device index layout (replaces the B+-tree rope of rope.c/rle.c and the
 * frame+header walk of rld0.c:371-408; see DESIGN.md "Data layout in HBM"):
 *
 *   blocks[b]  128 B, 128-B aligned = 8 x uint4 (one per lane of an 8-lane group)
 *       quad 0   counts of $,A,C  before the block, 3 x 42 bit packed in 128 bit
 *       quad 1   counts of G,T,N  before the block, same packing
 *       quad 2-7 48 run entries of 16 bit:  sym:3 | big:1 | val:12
 *                length = val (big=0) or val<<12 (big=1); val==0 -> padding
 *   bstart[b]  absolute position of the first symbol of block b (bstart[nb] = n)
 *   dir[j]     index of the block containing position j << dir_shift
 *
 * One rank query therefore touches one dir sector, (sometimes) one bstart
 * sector and exactly one 128-B block, which an 8-lane group reads with a single
 * coalesced 16-B-per-lane load and reduces with shuffles.
 */
#ifndef RB3B_INTERNAL_CUH
#define RB3B_INTERNAL_CUH

#include <stdint.h>
#include <stdio.h>
#include <cuda_runtime.h>
#include "../../include/rb3_b200.h"

#define RB3B_ENT_PER_BLK 48
#define RB3B_BIG_SHIFT   12
#define RB3B_VAL_MASK    0xfffu
#define RB3B_BIG_BIT     0x1000u
#define RB3B_GROUP       8          /* lanes cooperating on one query */
#define RB3B_M42         ((1ULL << 42) - 1)

struct rb3b_index_s {
	int64_t n;                    /* #symbols */
	int64_t tot[RB3B_ASIZE];      /* marginal counts */
	int64_t acc[RB3B_ASIZE + 1];  /* C[] */
	int64_t n_blocks, n_entries;
	uint4 *blocks;                /* n_blocks * 8 quads */
	uint64_t *bstart;             /* n_blocks + 1 */
	uint32_t *dir;                /* n_dir */
	int64_t n_dir;
	int dir_shift;
	size_t bytes;
};

/* by-value kernel argument */
struct DevIndex {
	const uint4 *blocks;
	const uint64_t *bstart;
	const uint32_t *dir;
	int64_t n, n_blocks;
	int dir_shift;
	int64_t tot[RB3B_ASIZE];
	int64_t acc[RB3B_ASIZE + 1];
};

static inline DevIndex rb3b_dev_view(const rb3b_index_s *x)
{
	DevIndex d;
	d.blocks = x->blocks; d.bstart = x->bstart; d.dir = x->dir;
	d.n = x->n; d.n_blocks = x->n_blocks; d.dir_shift = x->dir_shift;
	for (int c = 0; c < RB3B_ASIZE; ++c) d.tot[c] = x->tot[c];
	for (int c = 0; c <= RB3B_ASIZE; ++c) d.acc[c] = x->acc[c];
	return d;
}

/* ---- runtime (rb3b_runtime.cu) ---- */
extern cudaStream_t rb3b_stream;
extern int64_t rb3b_seg_len, rb3b_rank_variant;
int  rb3b_fail(int code, const char *fmt, ...);
int  rb3b_ensure_init(void);
void rb3b_stat_set(const char *key, int64_t v);
void rb3b_stat_add(const char *key, int64_t v);
extern int64_t rb3b_n_launch;   /* kernels of this library launched so far (CUB internals not counted) */
/* device-side timing of the main kernels: tic/toc record events on the stream, tflush (after a sync) adds "us_<name>" stats */
enum { T_PREP, T_WALK1, T_WALKFIX, T_MERGE, T_FINAL, T_BWT, T_COUNT };
void rb3b_tic(int id);
void rb3b_toc(int id);
void rb3b_tflush(void);

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return rb3b_fail(RB3B_ENODEV, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CKK() do { ++rb3b_n_launch; CK(cudaGetLastError()); } while (0)   /* after every launch of one of OUR kernels */
#define TRY(call) do { int r_ = (call); if (r_ != RB3B_OK) return r_; } while (0)

/* stream-ordered scratch buffer */
template<typename T> struct DBuf {
	T *p; size_t n;
	DBuf() : p(0), n(0) {}
	~DBuf() { release(); }
	int alloc(size_t n_) {
		release();
		n = n_;
		cudaError_t e = cudaMallocAsync((void**)&p, (n ? n : 1) * sizeof(T), rb3b_stream);
		if (e != cudaSuccess) { p = 0; return rb3b_fail(RB3B_ENOMEM, "cudaMallocAsync(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e)); }
		return RB3B_OK;
	}
	void release() { if (p) { cudaFreeAsync(p, rb3b_stream); p = 0; } }
	T *take() { T *q = p; p = 0; return q; }
private:
	DBuf(const DBuf&); DBuf &operator=(const DBuf&);
};

/* ---- primitives (rb3b_index.cu) ---- */
int rb3b_scan_excl_i64(const int64_t *d_in, int64_t *d_out, int64_t n);              /* exclusive prefix sum */
int rb3b_index_free_dev(rb3b_index_s *x);
int rb3b_index_from_runs_dev(rb3b_index_s *x, int64_t n_runs, const uint8_t *d_sym, const int64_t *d_len);
int rb3b_index_finalize(rb3b_index_s *x);     /* blocks hold entries; fills headers, bstart, dir, totals */
int rb3b_export_runs_dev(const rb3b_index_s *x, DBuf<uint8_t> &sym, DBuf<int64_t> &len, int64_t *n_runs);

/* ---- device helpers ---- */
#ifdef __CUDACC__

__device__ __forceinline__ unsigned rb3b_gmask()
{ /* mask of the 8-lane group this thread belongs to */
	return 0xffu << (threadIdx.x & 24);
}

__device__ __forceinline__ void rb3b_hdr_unpack(const uint4 v, uint64_t &c0, uint64_t &c1, uint64_t &c2)
{
	uint64_t lo = (uint64_t)v.x | (uint64_t)v.y << 32, hi = (uint64_t)v.z | (uint64_t)v.w << 32;
	c0 = lo & RB3B_M42;
	c1 = (lo >> 42 | hi << 22) & RB3B_M42;
	c2 = (hi >> 20) & RB3B_M42;
}

__device__ __forceinline__ uint4 rb3b_hdr_pack(uint64_t c0, uint64_t c1, uint64_t c2)
{
	uint64_t lo = c0 | c1 << 42, hi = c1 >> 22 | c2 << 20;
	return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

__device__ __forceinline__ uint32_t rb3b_ent_len(uint32_t e)
{
	uint32_t l = e & RB3B_VAL_MASK;
	return (e & RB3B_BIG_BIT) ? l << RB3B_BIG_SHIFT : l;
}

/* number of 16-bit entries needed by a run of length l */
__host__ __device__ __forceinline__ int64_t rb3b_nent(int64_t l)
{
	int64_t q = l >> RB3B_BIG_SHIFT;
	return (q + 4094) / 4095 + ((l & RB3B_VAL_MASK) ? 1 : 0);
}

/* pointer to entry e of the entry space embedded in the blocks */
__device__ __forceinline__ uint16_t *rb3b_ent_ptr(uint4 *blocks, int64_t e)
{
	int64_t b = e / RB3B_ENT_PER_BLK;
	int s = (int)(e - b * RB3B_ENT_PER_BLK);
	return (uint16_t*)(blocks + b * 8 + 2) + s;
}

/* write the entries of run (c,l) starting at entry index e; returns next e */
__device__ __forceinline__ int64_t rb3b_emit_run(uint4 *blocks, int64_t e, int c, int64_t l)
{
	int64_t q = l >> RB3B_BIG_SHIFT;
	while (q > 0) {
		int64_t t = q < 4095 ? q : 4095;
		*rb3b_ent_ptr(blocks, e++) = (uint16_t)(c << 13 | RB3B_BIG_BIT | (uint32_t)t);
		q -= t;
	}
	if (l & RB3B_VAL_MASK) *rb3b_ent_ptr(blocks, e++) = (uint16_t)(c << 13 | (uint32_t)(l & RB3B_VAL_MASK));
	return e;
}

/* block containing position k (0 <= k < n); executed by all lanes of a group with the same k */
__device__ __forceinline__ int64_t rb3b_locate(const DevIndex &x, int64_t k, int gl, unsigned gmask)
{
	int64_t j = k >> x.dir_shift;
	uint32_t b0 = __ldg(x.dir + j), b1 = __ldg(x.dir + j + 1);
	while (b1 - b0 > RB3B_GROUP) {
		uint32_t mid = b0 + (b1 - b0 + 1) / 2;
		if (__ldg(x.bstart + mid) <= (uint64_t)k) b0 = mid; else b1 = mid - 1;
	}
	if (b1 > b0) {
		uint32_t cand = b0 + 1 + gl;
		bool le = cand <= b1 && __ldg(x.bstart + cand) <= (uint64_t)k;
		b0 += __popc(__ballot_sync(gmask, le) & gmask);
	}
	return b0;
}

/* Decoded view of one block in the registers of an 8-lane group */
struct BlkLane {
	uint32_t len[8];
	uint32_t sym[8];
	uint32_t tot;      /* symbols held by this lane */
	uint32_t pre;      /* symbols held by lower lanes of the group */
	uint64_t start;    /* absolute position of the block */
	uint64_t c0, c1, c2; /* header counts of this lane (lanes 0,1 only) */
};

__device__ __forceinline__ void rb3b_decode(const uint4 v, int gl, unsigned gmask, BlkLane &B)
{
	uint64_t hs = 0;
	B.c0 = B.c1 = B.c2 = 0; B.tot = 0;
	if (gl < 2) {
		rb3b_hdr_unpack(v, B.c0, B.c1, B.c2);
		hs = B.c0 + B.c1 + B.c2;
#pragma unroll
		for (int j = 0; j < 8; ++j) B.len[j] = 0, B.sym[j] = 7;
	} else {
		const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			uint32_t e = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu;
			B.len[j] = rb3b_ent_len(e);
			B.sym[j] = e >> 13;
			B.tot += B.len[j];
		}
	}
	int gbase = threadIdx.x & 24;
	B.start = __shfl_sync(gmask, hs, gbase) + __shfl_sync(gmask, hs, gbase + 1);
	uint32_t inc = B.tot;
#pragma unroll
	for (int d = 1; d < RB3B_GROUP; d <<= 1) {
		uint32_t t = __shfl_up_sync(gmask, inc, d, RB3B_GROUP);
		if (gl >= d) inc += t;
	}
	B.pre = inc - B.tot;
}

/* #c in [0,k) of the indexed BWT; all 8 lanes of the group call with identical (k,c) */
__device__ __forceinline__ int64_t rb3b_rank_c(const DevIndex &x, int64_t k, int c)
{
	if (k >= x.n) return x.tot[c];
	if (k <= 0) return 0;
	const unsigned gmask = rb3b_gmask();
	const int gl = threadIdx.x & 7, gbase = threadIdx.x & 24;
	int64_t b = rb3b_locate(x, k, gl, gmask);
	uint4 v = __ldg(x.blocks + b * 8 + gl);
	BlkLane B;
	rb3b_decode(v, gl, gmask, B);
	int cc = c - 3 * gl;
	uint64_t hc = (gl < 2 && cc >= 0 && cc < 3) ? (cc == 0 ? B.c0 : cc == 1 ? B.c1 : B.c2) : 0;
	uint64_t base = __shfl_sync(gmask, hc, gbase + (c >= 3));
	uint32_t off = (uint32_t)((uint64_t)k - B.start);
	uint32_t rem = off > B.pre ? min(off - B.pre, B.tot) : 0;
	uint32_t contrib = 0;
#pragma unroll
	for (int j = 0; j < 8; ++j) {
		uint32_t take = min(B.len[j], rem);
		contrib += B.sym[j] == (uint32_t)c ? take : 0;
		rem -= take;
	}
#pragma unroll
	for (int d = RB3B_GROUP / 2; d > 0; d >>= 1) contrib += __shfl_xor_sync(gmask, contrib, d, RB3B_GROUP);
	return (int64_t)(base + contrib);
}

#endif /* __CUDACC__ */
#endif
