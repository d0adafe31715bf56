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
 *   dir[j]     8-B cell for positions [j << dir_shift, (j+1) << dir_shift):
 *                bits 0-31  b0 = block containing the first position of the cell
 *                bits 32-62 off = offset in the cell at which block b0+1 starts
 *                           (0x7fffffff: no block starts inside the cell)
 *                bit 63     more than one block starts inside: search bstart[]
 *
 * One rank query therefore touches one dir cell and exactly one 128-B block
 * (plus, rarely, a few bstart entries), which a group of G = 2, 4 or 8 lanes
 * reads with 16-B-per-lane loads and reduces with shuffles.
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
	uint4 *spare;                 /* the other half of the ping-pong pair: the next merge writes here */
	int64_t cap_blocks, cap_spare, cap_bstart, cap_dir; /* capacities (elements) of the persistent buffers */
	uint64_t *bstart;             /* n_blocks + 1 */
	uint64_t *dir;                /* n_dir cells: b0 | off << 32 | multi << 63 */
	int64_t n_dir;
	int dir_shift;
	size_t bytes;
};

/* by-value kernel argument */
struct DevIndex {
	const uint4 *blocks;
	const uint64_t *bstart;
	const uint64_t *dir;
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

static inline size_t rb3b_round_cap(size_t bytes)
{
	size_t c = 1 << 16;
	while (c < bytes) c += (c >> 2) & ~(size_t)255;
	return c;
}

/* stream-ordered scratch buffer */
template<typename T> struct DBuf {
	T *p; size_t n;
	DBuf() : p(0), n(0) {}
	~DBuf() { release(); }
	int alloc(size_t n_) {
		release();
		n = n_;
		/* sizes are quantised (x1.25 steps) so that the slightly larger buffers of the next merge reuse pool blocks */
		cudaError_t e = cudaMallocAsync((void**)&p, rb3b_round_cap((n ? n : 1) * sizeof(T)), rb3b_stream);
		if (e != cudaSuccess) { p = 0; return rb3b_fail(RB3B_ENOMEM, "cudaMallocAsync(%zu bytes): %s", n * sizeof(T), cudaGetErrorString(e)); }
		return RB3B_OK;
	}
	void release() { if (p) { cudaFreeAsync(p, rb3b_stream); p = 0; } }
	T *take() { T *q = p; p = 0; return q; }
private:
	DBuf(const DBuf&); DBuf &operator=(const DBuf&);
};

/* ---- primitives (rb3b_index.cu) ---- */
/* grow-only persistent buffer: reallocates (x1.5) only when `need` exceeds the capacity; contents are not kept */
int rb3b_reserve(void **p, int64_t *cap, int64_t need, size_t elt);
int rb3b_scan_excl_i64(const int64_t *d_in, int64_t *d_out, int64_t n);              /* exclusive prefix sum */
int rb3b_index_free_dev(rb3b_index_s *x);
int rb3b_index_from_runs_dev(rb3b_index_s *x, int64_t n_runs, const uint8_t *d_sym, const int64_t *d_len);
int rb3b_index_finalize(rb3b_index_s *x);     /* blocks hold entries; fills headers, bstart, dir, totals */
int rb3b_export_runs_dev(const rb3b_index_s *x, DBuf<uint8_t> &sym, DBuf<int64_t> &len, int64_t *n_runs);

/* ---- device helpers ---- */
#ifdef __CUDACC__

#define RB3B_DIR_NONE 0x7fffffffu

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

/*
 * Rank machinery for a group of G lanes (G = 2, 4 or 8) working on one query.
 * Lane gl holds quads [gl*NQ, (gl+1)*NQ) of the 128-B block, NQ = 8/G.
 */
template<int G> struct Grp {
	static const int NQ = 8 / G;
	__device__ __forceinline__ static int lane() { return threadIdx.x & (G - 1); }
	__device__ __forceinline__ static int base() { return threadIdx.x & 31 & ~(G - 1); }
	__device__ __forceinline__ static unsigned mask() { return ((1u << G) - 1u) << base(); }

	/* block containing position k (0 <= k < n); all lanes of the group pass the same k */
	__device__ __forceinline__ static int64_t locate(const DevIndex &x, int64_t k)
	{
		int64_t j = k >> x.dir_shift;
		uint64_t cell = __ldg(x.dir + j);
		uint32_t b0 = (uint32_t)cell, off = (uint32_t)(cell >> 32) & RB3B_DIR_NONE;
		uint64_t in_cell = (uint64_t)k - ((uint64_t)j << x.dir_shift);
		if ((int64_t)cell < 0) { /* rare: several block starts inside the cell */
			uint32_t b1 = (uint32_t)__ldg(x.dir + j + 1);
			while (b0 < b1) { /* last block in [b0,b1] whose start is <= k; same trip count for the whole group */
				uint32_t mid = b0 + (b1 - b0 + 1) / 2;
				if (__ldg(x.bstart + mid) <= (uint64_t)k) b0 = mid; else b1 = mid - 1;
			}
			return b0;
		}
		return (int64_t)b0 + (in_cell >= off ? 1 : 0);
	}

	__device__ __forceinline__ static void load(const DevIndex &x, int64_t b, uint4 (&v)[NQ])
	{
		const uint4 *p = x.blocks + b * 8 + lane() * NQ;
#pragma unroll
		for (int j = 0; j < NQ; ++j) v[j] = __ldg(p + j);
	}

	/* #c among the first (k - start of block) symbols of the block held in v, plus the header count of c */
	__device__ __forceinline__ static int64_t count(const uint4 (&v)[NQ], int64_t k, int c)
	{
		const int gl = lane(), gb = base();
		const unsigned gm = mask();
		uint64_t hs = 0, hc = 0; /* sum of the header counts held by this lane; header count of symbol c */
		if (G == 8) {
			if (gl < 2) {
				uint64_t a0, a1, a2;
				rb3b_hdr_unpack(v[0], a0, a1, a2);
				hs = a0 + a1 + a2;
				int cc = c - 3 * gl;
				hc = cc == 0 ? a0 : cc == 1 ? a1 : cc == 2 ? a2 : 0;
			}
		} else if (gl == 0) {
			uint64_t a[6];
			rb3b_hdr_unpack(v[0], a[0], a[1], a[2]);
			rb3b_hdr_unpack(v[NQ > 1 ? 1 : 0], a[3], a[4], a[5]);
			hs = a[0] + a[1] + a[2] + a[3] + a[4] + a[5];
			hc = c == 0 ? a[0] : c == 1 ? a[1] : c == 2 ? a[2] : c == 3 ? a[3] : c == 4 ? a[4] : a[5];
		}
		uint64_t start = __shfl_sync(gm, hs, gb);
		if (G == 8) start += __shfl_sync(gm, hs, gb + 1);
		uint64_t basec = __shfl_sync(gm, hc, gb + ((G == 8 && c >= 3) ? 1 : 0));
		/* entries */
		uint32_t len[NQ * 8], tot = 0;
		uint32_t isc = 0; /* bit i: entry i has symbol c */
#pragma unroll
		for (int j = 0; j < NQ; ++j) {
			const bool ent = gl * NQ + j >= 2; /* quads 0,1 are the header */
			const uint32_t w[4] = { v[j].x, v[j].y, v[j].z, v[j].w };
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				uint32_t e = (w[i >> 1] >> (16 * (i & 1))) & 0xffffu;
				uint32_t l = ent ? rb3b_ent_len(e) : 0;
				len[j * 8 + i] = l;
				isc |= ((e >> 13) == (uint32_t)c ? 1u : 0u) << (j * 8 + i);
				tot += l;
			}
		}
		uint32_t inc = tot;
#pragma unroll
		for (int d = 1; d < G; d <<= 1) {
			uint32_t t = __shfl_up_sync(gm, inc, d, G);
			if (gl >= d) inc += t;
		}
		uint32_t pre = inc - tot, off = (uint32_t)((uint64_t)k - start);
		uint32_t rem = off > pre ? min(off - pre, tot) : 0, contrib = 0;
#pragma unroll
		for (int i = 0; i < NQ * 8; ++i) {
			uint32_t take = min(len[i], rem);
			contrib += (isc >> i & 1) ? take : 0;
			rem -= take;
		}
#pragma unroll
		for (int d = G / 2; d > 0; d >>= 1) contrib += __shfl_xor_sync(gm, contrib, d, G);
		return (int64_t)(basec + contrib);
	}

	/* #c in [0,k) */
	__device__ __forceinline__ static int64_t rank(const DevIndex &x, int64_t k, int c)
	{
		int64_t kk = k < x.n ? (k < 0 ? 0 : k) : x.n - 1;
		uint4 v[NQ];
		load(x, locate(x, kk), v);
		int64_t r = count(v, kk, c);
		return k < x.n ? r : x.tot[c];
	}

	/* two positions at once, same symbol: the two block fetches overlap and the instruction stream is the same
	 * whether or not k1 == k2, which keeps the groups of a warp in lockstep during the LF walk */
	__device__ __forceinline__ static void rank2(const DevIndex &x, int64_t k1, int64_t k2, int c, int64_t &r1, int64_t &r2)
	{
		int64_t q1 = k1 < x.n ? k1 : x.n - 1, q2 = k2 < x.n ? k2 : x.n - 1;
		int64_t b1 = locate(x, q1), b2 = locate(x, q2);
		uint4 v1[NQ], v2[NQ];
		load(x, b1, v1); load(x, b2, v2);
		int64_t t = x.tot[c];
		r1 = count(v1, q1, c); r2 = count(v2, q2, c);
		r1 = k1 < x.n ? r1 : t; r2 = k2 < x.n ? r2 : t;
	}
};

#endif /* __CUDACC__ */
#endif
