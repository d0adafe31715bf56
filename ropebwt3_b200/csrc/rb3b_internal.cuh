/*
 * rb3b_internal.cuh -- shared definitions of the B200-native BWT-merge engine.
 *
 * Device index layout (replaces the B+-tree rope of rope.c/rle.c and the
 * frame+header walk of rld0.c:371-408; see DESIGN.md "Data layout in HBM"):
 *
 *   The BWT is cut into CELLS of a fixed span of 2^shift positions
 *   (5 <= shift <= 16).  cells[j] is one 128-B line = 8 x uint4 describing
 *   positions [j << shift, (j+1) << shift):
 *       quad 0   counts of $,A,C before the cell, 3 x 42 bit; bit 127 = OVERFLOW
 *       quad 1   counts of G,T,N before the cell, same packing
 *       quad 2-7 inline cell: up to 48 run entries of 16 bit, sym:3 | len:13,
 *                len == 0 = padding.  A run never crosses a cell boundary.
 *   A cell whose span holds more than 48 runs is an overflow cell:
 *       quad 2   u32 first overflow block, u32 number of blocks
 *       quad 3-7 40 x u16: offset in the cell at which blocks 1..40 start
 *   ovf[b]     128-B overflow block: quad 0 = 6 x u16 counts of the symbols of
 *              this cell that precede the block (+2 spare), quads 1-7 = 56 entries.
 *
 * The cell of position k is cells[k >> shift]: a rank query is ONE 128-B access
 * at an arithmetic address (plus one overflow block for the few dense cells),
 * read by a group of G = 2, 4 or 8 lanes with 16-B-per-lane loads and reduced
 * with shuffles.  shift is chosen per index so that an average cell is ~40% full.
 *
 * Second cell kind, BITMAP (kind = 1), used while the index fits 1 byte per
 * symbol in HBM: span 128, two self-contained 64-B halves
 *       quad 0   counts of $,A,C before the cell (3 x 42 bit)
 *       quad 1-3 one-hot bit planes of $, A, C: bit p set <=> position p holds it
 *       quad 4   counts of G,T,N            quad 5-7  planes of G, T, N
 * rank(c, k) = count + popcount(plane below k): two 16-B loads inside one 64-B
 * DRAM atom, ~20 instructions, one thread per query, no decoding at all.
 */
#ifndef RB3B_INTERNAL_CUH
#define RB3B_INTERNAL_CUH

#include <stdint.h>
#include <stdio.h>
#include <map>
#include <string>
#include <vector>
#include <cuda_runtime.h>
#include "../../include/rb3_b200.h"

#define RB3B_ENT_PER_CELL 48
#define RB3B_ENT_PER_OVF  56
#define RB3B_MAX_OVF      41
#define RB3B_MIN_SHIFT    5
#define RB3B_MAX_SHIFT    16         /* cells of up to 65536 positions: a run longer than 8191 takes several entries */
#define RB3B_LEN_MASK     0x1fffu
#define RB3B_M42          ((1ULL << 42) - 1)
#define RB3B_KIND_RLE     0
#define RB3B_KIND_BM      1
#define RB3B_BM_SHIFT     7
#define RB3B_GROUP        8          /* lanes per query of the API rank kernels and of the LF walk */

struct rb3b_index_s {
	int64_t n;                    /* #symbols */
	int64_t tot[RB3B_ASIZE];      /* marginal counts */
	int64_t acc[RB3B_ASIZE + 1];  /* C[] */
	int shift;                    /* log2 of the cell span */
	int kind;                     /* RB3B_KIND_RLE or RB3B_KIND_BM */
	int so;                       /* order of the sentinels / of equal suffixes: 0 input order, 1 RLO, 2 RCLO (mrope.h:6-8) */
	int64_t n_cells, n_ovf, n_entries;
	uint4 *cells, *ovf;           /* n_cells * 8 and n_ovf * 8 quads */
	uint4 *cells2, *ovf2;         /* the other half of the ping-pong pair: the next merge writes here */
	int64_t cap_cells, cap_ovf, cap_cells2, cap_ovf2; /* capacities in quads */
	size_t bytes;
	/* asynchronous merge (bitmap -> bitmap): the merge kernels run on the context's second stream while the caller goes on
	 * (typically into the next batch's LF table and walk-order rewrite, which do not need the index).  The host-side fields
	 * above already describe the merged index (its totals are known beforehand: old totals + batch totals). */
	cudaEvent_t ready_ev;         /* recorded behind the last kernel that writes the current cells */
	int has_ev;
	int pending;                  /* a merge is in flight; its deferred checks are in pend_host once ready_ev has fired */
	int broken;                   /* a deferred check failed: the cells are not the index the host fields describe */
	int64_t *pend_host;           /* pinned: [0..5] symbol totals the merge kernels counted, [6] "positions not monotone" flag */
	int64_t pend_expect[RB3B_ASIZE];
	int ms_flip;                  /* which of the two batch copies the NEXT merge uses (the merge in flight reads the other) */
	int64_t ms_rows;              /* the scratch is laid out for batches of up to this many rows (grow-only, so that the regions of consecutive merges coincide) */
	char *ms; size_t ms_cap, ms_used; /* scratch that must outlive the call: interleave positions and batch copy (ms_used bytes), then the merge's tables */
};

/* a batch prepared for merging (rb3b_batch_prepare*): its partial BWT and the batch in walk order, all in device memory
 * owned by the object.  Replaces what step 0 of the reference's pipeline hands to step 1 (build.c:55-83). */
struct rb3b_batch_s {
	int64_t len, n_seq;
	int64_t acc[RB3B_ASIZE + 1];  /* C[] of the batch */
	int device;
	uint8_t *bwt;                 /* len */
	uint8_t *wsym;                /* len + 64, or NULL: only the BWT was prepared (batches of 2^29 symbols or more) */
	uint32_t *wrow;               /* len + 8 */
	int64_t *c_base, *c_len;      /* n_seq each */
};

/* by-value kernel argument */
struct DevIndex {
	const uint4 *cells, *ovf;
	int64_t n, n_cells;
	int shift, kind;
	int64_t tot[RB3B_ASIZE];
	int64_t acc[RB3B_ASIZE + 1];
};

static inline DevIndex rb3b_dev_view(const rb3b_index_s *x)
{
	DevIndex d;
	d.cells = x->cells; d.ovf = x->ovf;
	d.n = x->n; d.n_cells = x->n_cells; d.shift = x->shift; d.kind = x->kind;
	for (int c = 0; c < RB3B_ASIZE; ++c) d.tot[c] = x->tot[c];
	for (int c = 0; c <= RB3B_ASIZE; ++c) d.acc[c] = x->acc[c];
	return d;
}

/* ---- runtime (rb3b_runtime.cu) ---- */
/* device-side timing of the main kernels: tic/toc record events on the stream, tflush (after a sync) adds "us_<name>" stats */
enum { T_PREP, T_WALK1, T_WALKFIX, T_MERGE, T_SCATTER, T_BWT, T_COMM, T_COUNT };

struct Rb3bChunk { char *p; size_t cap; };
/* execution context: everything a call needs besides its arguments.  One per host thread by default (created on first
 * use), or explicit (rb3b_ctx_create) to drive several devices / streams from one process. */
struct rb3b_ctx_s {
	int device;
	cudaStream_t stream, my_stream;   /* the stream calls run on; the context's own stream */
	int own_stream;                   /* stream was supplied by the caller (rb3b_set_stream) */
	std::vector<Rb3bChunk> chunks;    /* scratch arena */
	size_t chunk_i, chunk_off, used, high;
	int depth;                        /* nesting of API calls: the arena is reset when the outermost returns */
	std::map<std::string, int64_t> stats;
	int64_t n_launch;                 /* kernels of this library launched so far (CUB internals not counted) */
	cudaEvent_t ev[T_COUNT][2];
	int ev_ok, ev_pending[T_COUNT];
	void *comm;                       /* ncclComm_t when this context is a rank of a multi-device group (rb3b_dist.cu) */
	void *comm2;                      /* a second communicator of the same ranks for collectives queued on stream2 (0: none) */
	int rank, world;
	cudaStream_t stream2;             /* asynchronous merges run here */
	cudaEvent_t ev_hand;              /* hand-over between the two streams */
	char *bump; size_t bump_off, bump_cap; /* when set, scratch comes from this region (an index's merge scratch) instead of the arena */
	/* batches copied ahead of the call that consumes them (rb3b_prefetch_batch): two staging buffers, a copy stream */
	struct { const void *host; int64_t len; uint8_t *dev; size_t cap; cudaEvent_t ev; int valid; } pf[2];
	int pf_next;
	cudaStream_t stream_copy;
};
uint8_t *rb3b_prefetched(const void *host, int64_t len); /* the device copy of a prefetched batch (the current stream waits for it), or NULL */
rb3b_ctx_s *rb3b_cur(void);
#define rb3b_stream   (rb3b_cur()->stream)
#define rb3b_n_launch (rb3b_cur()->n_launch)
extern int64_t rb3b_seg_len, rb3b_rank_variant;
int  rb3b_fail(int code, const char *fmt, ...);
int  rb3b_ensure_init(void);
void rb3b_stat_set(const char *key, int64_t v);
void rb3b_stat_add(const char *key, int64_t v);
int64_t rb3b_get_param(const char *key, int64_t dflt);
void rb3b_tic(int id);
void rb3b_toc(int id);
void rb3b_tflush(void);

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
	return rb3b_fail(RB3B_ENODEV, "%s:%d: %s: %s", __FILE__, __LINE__, #call, cudaGetErrorString(e_)); } while (0)
#define CKK() do { ++rb3b_n_launch; CK(cudaGetLastError()); } while (0)   /* after every launch of one of OUR kernels */
#define TRY(call) do { int r_ = (call); if (r_ != RB3B_OK) return r_; } while (0)

/*
 * Scratch memory.  Every API call bump-allocates its temporaries from a device arena that is reset when the
 * outermost call returns, so steady-state merges perform no cudaMalloc/cudaFree at all (allocation growth used
 * to cost more than the kernels).  The arena grows by adding chunks; the next reset folds them into one.
 */
void *rb3b_arena_alloc(size_t bytes);
void rb3b_arena_enter(void);
void rb3b_arena_leave(void);
struct ApiScope { ApiScope() { rb3b_arena_enter(); } ~ApiScope() { rb3b_arena_leave(); } };

template<typename T> struct DBuf {
	T *p; size_t n;
	DBuf() : p(0), n(0) {}
	int alloc(size_t n_) {
		n = n_;
		p = (T*)rb3b_arena_alloc((n ? n : 1) * sizeof(T));
		if (p == 0) return rb3b_fail(RB3B_ENOMEM, "device arena: cannot allocate %zu bytes", n * sizeof(T));
		return RB3B_OK;
	}
private:
	DBuf(const DBuf&); DBuf &operator=(const DBuf&);
};

/* ---- primitives (rb3b_index.cu) ---- */
/* grow-only persistent buffer: reallocates (x1.5) only when `need` exceeds the capacity; contents are not kept */
int rb3b_reserve(void **p, int64_t *cap, int64_t need, size_t elt);
int rb3b_scan_excl_i64(const int64_t *d_in, int64_t *d_out, int64_t n);              /* exclusive prefix sum */
int rb3b_index_free_dev(rb3b_index_s *x);
int rb3b_index_wait_i(rb3b_index_s *x);          /* host: wait for the in-flight merge of x and run its deferred checks */
int rb3b_index_use(const rb3b_index_s *x);       /* device: kernels launched on the current stream from now on see the merged cells */
int rb3b_index_from_runs_dev(rb3b_index_s *x, int64_t n_runs, const uint8_t *d_sym, const int64_t *d_len);
int rb3b_export_runs_dev(const rb3b_index_s *x, DBuf<uint8_t> &sym, DBuf<int64_t> &len, int64_t *n_runs);
int rb3b_index_to_plain_dev(const rb3b_index_s *x, DBuf<uint8_t> &plain);   /* rb3b_bwt.cu */
int rb3b_pick_shift(int64_t n, int64_t n_entries_est);
int rb3b_want_bitmap(int64_t n_symbols);

/* ---- device helpers ---- */
#ifdef __CUDACC__

__device__ __forceinline__ void rb3b_hdr_unpack(const uint4 v, uint64_t &c0, uint64_t &c1, uint64_t &c2)
{
	uint64_t lo = (uint64_t)v.x | (uint64_t)v.y << 32, hi = (uint64_t)v.z | (uint64_t)v.w << 32;
	c0 = lo & RB3B_M42;
	c1 = (lo >> 42 | hi << 22) & RB3B_M42;
	c2 = (hi >> 20) & RB3B_M42;
}

__device__ __forceinline__ uint4 rb3b_hdr_pack(uint64_t c0, uint64_t c1, uint64_t c2, bool flag)
{
	uint64_t lo = c0 | c1 << 42, hi = c1 >> 22 | c2 << 20 | (flag ? 1ULL << 63 : 0);
	return make_uint4((uint32_t)lo, (uint32_t)(lo >> 32), (uint32_t)hi, (uint32_t)(hi >> 32));
}

__device__ __forceinline__ bool rb3b_is_ovf(const uint4 q0) { return (q0.w >> 31) != 0; }

/* #c among the first `off` symbols of an overflow cell (excluding the cell header count).  Slow path, executed
 * identically by every lane that calls it. */
static __device__ __noinline__ uint32_t rb3b_ovf_count(const uint4 *__restrict__ cell, const uint4 *__restrict__ ovf, uint32_t off, int c)
{
	uint4 q2 = __ldg(cell + 2);
	uint32_t first = q2.x, nblk = q2.y, t = 0, start = 0;
	const uint16_t *st = (const uint16_t*)(cell + 3);
	for (uint32_t i = 0; i + 1 < nblk; ++i) {
		uint32_t s = st[i];
		if (s <= off) { t = i + 1; start = s; } else break;
	}
	const uint4 *blk = ovf + (int64_t)(first + t) * 8;
	uint32_t cnt = ((const uint16_t*)blk)[c], rem = off - start;
	const uint16_t *e = (const uint16_t*)(blk + 1);
	for (int i = 0; i < RB3B_ENT_PER_OVF && rem > 0; ++i) {
		uint32_t x = e[i], l = x & RB3B_LEN_MASK, take = l < rem ? l : rem;
		if ((x >> 13) == (uint32_t)c) cnt += take;
		rem -= take;
	}
	return cnt;
}

/*
 * Rank machinery for a group of G lanes (G = 2, 4 or 8) working on one query.
 * Lane gl holds quads [gl*NQ, (gl+1)*NQ) of the 128-B cell, NQ = 8/G.
 */
template<int G_> struct Grp {
	static const int G = G_;
	static const bool ALWAYS2 = false;
	static const int NQ = 8 / G;
	__device__ __forceinline__ static int lane() { return threadIdx.x & (G - 1); }
	__device__ __forceinline__ static int base() { return threadIdx.x & 31 & ~(G - 1); }
	__device__ __forceinline__ static unsigned mask() { return ((1u << G) - 1u) << base(); }

	__device__ __forceinline__ static void load(const DevIndex &x, int64_t j, uint4 (&v)[NQ])
	{
		const uint4 *p = x.cells + j * 8 + lane() * NQ;
#pragma unroll
		for (int i = 0; i < NQ; ++i) v[i] = __ldg(p + i);
	}

	/* #c in [0, k) given the cell of k in v; 0 <= k < n.  All lanes of the group pass the same (k, c). */
	__device__ __forceinline__ static int64_t count(const DevIndex &x, const uint4 (&v)[NQ], int64_t k, int c)
	{
		const int gl = lane(), gb = base();
		const unsigned gm = mask();
		uint64_t hc = 0; /* header count of symbol c, held by one lane */
		uint32_t flag = 0;
		if (G == 8) {
			if (gl < 2) {
				uint64_t a0, a1, a2;
				rb3b_hdr_unpack(v[0], a0, a1, a2);
				int cc = c - 3 * gl;
				hc = cc == 0 ? a0 : cc == 1 ? a1 : cc == 2 ? a2 : 0;
				flag = gl == 0 ? v[0].w >> 31 : 0;
			}
		} else if (gl == 0) {
			uint64_t a[6];
			rb3b_hdr_unpack(v[0], a[0], a[1], a[2]);
			rb3b_hdr_unpack(v[NQ > 1 ? 1 : 0], a[3], a[4], a[5]);
			hc = c == 0 ? a[0] : c == 1 ? a[1] : c == 2 ? a[2] : c == 3 ? a[3] : c == 4 ? a[4] : a[5];
			flag = v[0].w >> 31;
		}
		uint64_t basec = __shfl_sync(gm, hc, gb + ((G == 8 && c >= 3) ? 1 : 0));
		flag = __shfl_sync(gm, flag, gb);
		const uint32_t off = (uint32_t)k & ((1u << x.shift) - 1u);
		if (flag) return (int64_t)(basec + rb3b_ovf_count(x.cells + (k >> x.shift) * 8, x.ovf, off, c)); /* rare, group-uniform */
		uint32_t len[NQ * 8], tot = 0, isc = 0; /* isc bit i: entry i has symbol c */
#pragma unroll
		for (int j = 0; j < NQ; ++j) {
			const bool ent = gl * NQ + j >= 2; /* quads 0,1 are the header */
			const uint32_t w[4] = { v[j].x, v[j].y, v[j].z, v[j].w };
#pragma unroll
			for (int i = 0; i < 8; ++i) {
				uint32_t e = (w[i >> 1] >> (16 * (i & 1))) & 0xffffu;
				uint32_t l = ent ? (e & RB3B_LEN_MASK) : 0;
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
		uint32_t pre = inc - tot;
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
		load(x, kk >> x.shift, v);
		int64_t r = count(x, v, kk, c);
		return k < x.n ? r : x.tot[c];
	}

	/* two positions at once, same symbol: the two cell fetches overlap and the instruction stream is the same
	 * whether or not k1 == k2, which keeps the groups of a warp in lockstep during the LF walk */
	__device__ __forceinline__ static void rank2(const DevIndex &x, int64_t k1, int64_t k2, int c, int64_t &r1, int64_t &r2)
	{
		int64_t q1 = k1 < x.n ? k1 : x.n - 1, q2 = k2 < x.n ? k2 : x.n - 1;
		uint4 v1[NQ], v2[NQ];
		load(x, q1 >> x.shift, v1); load(x, q2 >> x.shift, v2);
		int64_t t = x.tot[c];
		r1 = count(x, v1, q1, c); r2 = count(x, v2, q2, c);
		r1 = k1 < x.n ? r1 : t; r2 = k2 < x.n ? r2 : t;
	}
};

/* Run-length cells, ONE thread per rank (the arithmetic of k_lf_t1, rb3b_index.cu): the header quad of the symbol's half,
 * then the entry quads one after the other until the offset is reached.  As a walk ranker (G = 1) it costs ~10x fewer warp
 * instructions per row than the 8-lane groups: ncu showed the Grp<8> walk bound by instruction issue (135 warp instructions
 * per row, instruction-cache misses), not by memory. */
struct RleT1 {
	static const int G = 1;
	static const bool ALWAYS2 = false;
	__device__ __forceinline__ static int lane() { return 0; }
	__device__ __forceinline__ static int base() { return threadIdx.x & 31; }
	__device__ __forceinline__ static unsigned mask() { return 1u << (threadIdx.x & 31); }
	__device__ __forceinline__ static int64_t count(const DevIndex &x, int64_t k, int c)
	{ /* 0 <= k < n */
		const uint4 *cell = x.cells + (k >> x.shift) * 8;
		const int h = c >= 3, cc = c - 3 * h;
		const uint4 hq = __ldg(cell + h);
		uint4 e0 = __ldg(cell + 2), e1 = __ldg(cell + 3); /* in flight together with the header */
		uint64_t a0, a1, a2;
		rb3b_hdr_unpack(hq, a0, a1, a2);
		const uint64_t basec = cc == 0 ? a0 : cc == 1 ? a1 : a2;
		const uint32_t off = (uint32_t)k & ((1u << x.shift) - 1u);
		uint32_t cnt = 0;
		if (hq.w >> 31) cnt = rb3b_ovf_count(cell, x.ovf, off, c); /* the overflow flag is repeated in both header quads */
		else {
			uint32_t rem = off;
#pragma unroll 1
			for (int qd = 2; qd < 8 && rem > 0; qd += 2) {
				if (qd > 2) { e0 = __ldg(cell + qd); e1 = __ldg(cell + qd + 1); }
				const uint32_t w[8] = { e0.x, e0.y, e0.z, e0.w, e1.x, e1.y, e1.z, e1.w };
#pragma unroll
				for (int i = 0; i < 16; ++i) {
					const uint32_t e = (w[i >> 1] >> (16 * (i & 1))) & 0xffffu, l = e & RB3B_LEN_MASK, take = min(l, rem);
					cnt += (e >> 13) == (uint32_t)c ? take : 0u;
					rem -= take;
				}
			}
		}
		return (int64_t)(basec + cnt);
	}
	__device__ __forceinline__ static int64_t rank(const DevIndex &x, int64_t k, int c)
	{
		const int64_t kk = k < x.n ? (k < 0 ? 0 : k) : x.n - 1;
		const int64_t r = count(x, kk, c);
		return k < x.n ? r : x.tot[c];
	}
	__device__ __forceinline__ static void rank2(const DevIndex &x, int64_t k1, int64_t k2, int c, int64_t &r1, int64_t &r2)
	{
		r1 = rank(x, k1, c); r2 = rank(x, k2, c);
	}
};

/* Sequential reader of the symbols of an index from a given position on (used by the merge and the export) */
struct CellReader {
	const uint4 *cells, *ovf;
	int64_t n, j;        /* current cell */
	int shift;
	uint32_t left;       /* positions of the current cell not yet consumed */
	uint32_t eidx;       /* next entry of the current cell */
	uint32_t first;      /* first overflow block, or 0xffffffff for an inline cell */
	int cur; uint32_t rem; /* current run piece */

	__device__ __forceinline__ uint32_t entry(uint32_t i) const
	{
		if (first == 0xffffffffu) return ((const uint16_t*)(cells + j * 8 + 2))[i];
		return ((const uint16_t*)(ovf + (int64_t)(first + i / RB3B_ENT_PER_OVF) * 8 + 1))[i % RB3B_ENT_PER_OVF];
	}
	__device__ __forceinline__ void open_cell()
	{
		uint4 q0 = cells[j * 8];
		first = rb3b_is_ovf(q0) ? cells[j * 8 + 2].x : 0xffffffffu;
		int64_t p0 = j << shift, span = n - p0 < (1LL << shift) ? n - p0 : (1LL << shift);
		left = (uint32_t)span; eidx = 0;
	}
	__device__ __forceinline__ void next_piece()
	{ /* precondition: rem == 0 */
		if (left == 0) {
			if (((j + 1) << shift) >= n) { cur = -1; return; }
			++j; open_cell();
		}
		uint32_t e = entry(eidx++);
		cur = (int)(e >> 13); rem = e & RB3B_LEN_MASK;
		left -= rem;
	}
	__device__ __forceinline__ void seek(int64_t pos)
	{ /* 0 <= pos < n */
		j = pos >> shift;
		open_cell();
		uint32_t skip = (uint32_t)(pos - (j << shift));
		rem = 0;
		for (;;) {
			next_piece();
			if (skip < rem) { rem -= skip; return; }
			skip -= rem; rem = 0;
		}
	}
	__device__ __forceinline__ void advance(uint32_t l) { rem -= l; if (rem == 0) next_piece(); }
};


/* ---- bitmap cells ---- */

__device__ __forceinline__ int rb3b_bm_plane_quad(int s) { return s < 3 ? 1 + s : 2 + s; }

/* set bits of the 128-bit plane p among positions [0, off), 0 <= off <= 128 */
__device__ __forceinline__ uint32_t rb3b_bm_popc_below(const uint4 p, uint32_t off)
{
	const uint32_t w[4] = { p.x, p.y, p.z, p.w };
	uint32_t r = 0;
#pragma unroll
	for (int i = 0; i < 4; ++i) {
		int t = (int)off - 32 * i;
		uint32_t m = t >= 32 ? 0xffffffffu : t <= 0 ? 0u : (1u << t) - 1u;
		r += __popc(w[i] & m);
	}
	return r;
}

/* one thread per query */
struct BmRank {
	static const int G = 1;
	static const bool ALWAYS2 = false;
	__device__ __forceinline__ static int lane() { return 0; }
	__device__ __forceinline__ static int base() { return threadIdx.x & 31; }
	__device__ __forceinline__ static unsigned mask() { return 1u << (threadIdx.x & 31); }
	__device__ __forceinline__ static int64_t count(const DevIndex &x, int64_t k, int c)
	{ /* 0 <= k < n.  Kept short on purpose: in the LF walk this sits on a dependent chain, one call per row. */
		const int h = c >= 3, cc = c - 3 * h;
		const uint4 *half = x.cells + ((k >> RB3B_BM_SHIFT) * 8 + 4 * h);
		const uint4 cq = __ldg(half), pq = __ldg(half + 1 + cc);
		const uint64_t lo = (uint64_t)cq.x | (uint64_t)cq.y << 32, hi = (uint64_t)cq.z | (uint64_t)cq.w << 32;
		const uint64_t cnt = (cc == 0 ? lo : cc == 1 ? (lo >> 42 | hi << 22) : hi >> 20) & RB3B_M42;
		const uint64_t p0 = (uint64_t)pq.x | (uint64_t)pq.y << 32, p1 = (uint64_t)pq.z | (uint64_t)pq.w << 32;
		const uint32_t o = (uint32_t)k & 127u;
		const uint64_t m0 = o >= 64u ? ~0ULL : (1ULL << o) - 1ULL, m1 = o <= 64u ? 0ULL : (1ULL << (o - 64u)) - 1ULL;
		return (int64_t)(cnt + (uint32_t)(__popcll(p0 & m0) + __popcll(p1 & m1)));
	}
	__device__ __forceinline__ static int64_t rank(const DevIndex &x, int64_t k, int c)
	{
		int64_t kk = k < x.n ? (k < 0 ? 0 : k) : x.n - 1;
		int64_t r = count(x, kk, c);
		return k < x.n ? r : x.tot[c];
	}
	__device__ __forceinline__ static void rank2(const DevIndex &x, int64_t k1, int64_t k2, int c, int64_t &r1, int64_t &r2)
	{
		int64_t q1 = k1 < x.n ? k1 : x.n - 1, q2 = k2 < x.n ? k2 : x.n - 1, t = x.tot[c];
		r1 = count(x, q1, c); r2 = count(x, q2, c);
		r1 = k1 < x.n ? r1 : t; r2 = k2 < x.n ? r2 : t;
	}
};

/* two lanes per walk: while a walk carries a bracket [lo,hi] the even lane ranks lo and the odd lane hi, so that each
 * lane executes the instruction stream of ONE rank per step (the walk is bound by the instruction latency of the few
 * resident warps, not by bandwidth); for an exact walk both lanes load the same cell, which the LSU merges */
struct BmPair {
	static const int G = 2;
	static const bool ALWAYS2 = true;
	__device__ __forceinline__ static int lane() { return threadIdx.x & 1; }
	__device__ __forceinline__ static int base() { return threadIdx.x & 30; }
	__device__ __forceinline__ static unsigned mask() { return 3u << (threadIdx.x & 30); }
	__device__ __forceinline__ static int64_t rank(const DevIndex &x, int64_t k, int c) { return BmRank::rank(x, k, c); }
	__device__ __forceinline__ static void rank2(const DevIndex &x, int64_t k1, int64_t k2, int c, int64_t &r1, int64_t &r2)
	{
		const int odd = threadIdx.x & 1;
		const int64_t mine = BmRank::rank(x, odd ? k2 : k1, c);
		const int64_t other = __shfl_xor_sync(mask(), mine, 1);
		r1 = odd ? other : mine; r2 = odd ? mine : other;
	}
};

/* Sequential reader of a bitmap index as run pieces (same interface as CellReader) */
struct BmReader {
	const uint4 *cells;
	int64_t n, pos; /* pos = position of the current piece */
	int cur; uint32_t rem;
	__device__ __forceinline__ uint32_t word(int64_t j, int s, int w) const { return __ldg((const uint32_t*)(cells + j * 8 + rb3b_bm_plane_quad(s)) + w); }
	__device__ __forceinline__ void load_piece()
	{ /* the maximal run starting at pos that stays inside pos's cell */
		if (pos >= n) { cur = -1; rem = 0; return; }
		const int64_t j = pos >> RB3B_BM_SHIFT;
		uint32_t p = (uint32_t)pos & 127u, w = p >> 5, b = p & 31u;
		int s = 0;
#pragma unroll
		for (int a = 1; a < RB3B_ASIZE; ++a) if (word(j, a, w) >> b & 1u) s = a;
		cur = s;
		uint32_t len = 0;
		for (;;) {
			uint32_t inv = ~(word(j, s, w) >> b); /* zero bits above the shifted-in zeros count as run ends */
			uint32_t ones = inv ? (uint32_t)__ffs(inv) - 1u : 32u;
			if (ones > 32u - b) ones = 32u - b;
			len += ones;
			if (ones < 32u - b || w == 3) break;
			++w; b = 0;
		}
		int64_t cap = n - pos;
		rem = (int64_t)len < cap ? len : (uint32_t)cap;
	}
	__device__ __forceinline__ void seek(int64_t p) { pos = p; load_piece(); }
	__device__ __forceinline__ void advance(uint32_t l) { rem -= l; pos += l; if (rem == 0) load_piece(); }
};

#endif /* __CUDACC__ */
#endif
