/*
 * rb3b_merge.cu -- the merge hot path on the device.
 *
 * Replaces rb3_fmi_merge_plain (fm-index.c:279-303):
 *   phase A  rb3_mg_rank_plain + rb3_mg_rank1_plain (fm-index.c:160-225): for
 *            every row i of the batch BWT B, ka[i] = #suffixes of the indexed
 *            collection A that are smaller than suffix i of B.
 *   phase B  worker_mgins (fm-index.c:237-249) / rope_insert_run (rope.c:114) /
 *            rle_insert_cached (rle.c:10): interleave B into A.
 *
 * The reference walks one dependent LF chain per new sequence.  Here every chain
 * is cut into segments at "marked" rows of B.  A segment that does not start at a
 * sentinel does not know its ka yet, so it starts with the bracket [lo,hi] = SA
 * interval (in A) of the empty string restricted to its first symbol and narrows
 * it by backward search with the symbols it walks over; once lo == hi the value
 * is exact (the walked string no longer occurs in A) and independent of anything
 * to its right.  Rows walked before the collapse stay unresolved and are filled
 * in a later round by re-walking them from the exact value with which the
 * segment to the right arrived at the mark.  Rounds repeat until nothing is
 * unresolved (a batch sequence that is an exact substring of A degenerates to
 * the sequential chain, still correct).
 *
 * Phase B is a streaming merge: because ka[] is non-decreasing, block b of A
 * and the rows with bstart[b] <= ka < bstart[b+1] form an independent tile whose
 * merged runs are counted, prefix-summed and written into a fresh block array.
 */
#include <string.h>
#include <cub/cub.cuh>
#include "rb3b_internal.cuh"
#include "rb3b_emit.cuh"

#define TPB 256
#define PREP_PER_THREAD 16
#define PREP_TILE (TPB * PREP_PER_THREAD)

static inline unsigned nblk(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

static int n_sm(void)
{
	static int n = 0;
	if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
	return n;
}

/* ------------------------------------------------------------------ */
/* LF mapping of the batch (fm-index.c:207-216)                         */
/* lfb[i] = LF_B(i) << 5 | coarse mark << 4 | fine mark << 3 | B[i]     */
/* ------------------------------------------------------------------ */

#define LFB_SHIFT 5
#define LFB_FINE 8u
#define LFB_COARSE 16u

__global__ void __launch_bounds__(TPB) k_prep_count(int64_t len, const uint8_t *__restrict__ bwt, int64_t nt, int64_t *__restrict__ tcnt, int *__restrict__ bad)
{
	__shared__ unsigned int sh[RB3B_ASIZE];
	if (threadIdx.x < RB3B_ASIZE) sh[threadIdx.x] = 0;
	__syncthreads();
	int64_t i0 = (int64_t)blockIdx.x * PREP_TILE + (int64_t)threadIdx.x * PREP_PER_THREAD;
	unsigned int c[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
	for (int j = 0; j < PREP_PER_THREAD; ++j) {
		int64_t i = i0 + j;
		if (i < len) {
			int a = bwt[i];
			if (a >= RB3B_ASIZE) { *bad = 1; a = 5; }
#pragma unroll
			for (int b = 0; b < RB3B_ASIZE; ++b) c[b] += a == b;
		}
	}
#pragma unroll
	for (int b = 0; b < RB3B_ASIZE; ++b) {
		unsigned int v = c[b];
		for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
		if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[b], v);
	}
	__syncthreads();
	if (threadIdx.x < RB3B_ASIZE) tcnt[(int64_t)threadIdx.x * (nt + 1) + blockIdx.x] = sh[threadIdx.x];
	if (blockIdx.x == 0 && threadIdx.x < RB3B_ASIZE) tcnt[(int64_t)threadIdx.x * (nt + 1) + nt] = 0;
}

struct Acc7 { int64_t v[RB3B_ASIZE + 1]; };

__global__ void __launch_bounds__(TPB) k_prep_lf(int64_t len, const uint8_t *__restrict__ bwt, int64_t nt, const int64_t *__restrict__ tex,
                                                  Acc7 accB, int64_t fine_len, uint64_t *__restrict__ lfb)
{
	typedef cub::BlockScan<uint32_t, TPB> Scan;
	__shared__ typename Scan::TempStorage tmp[3];
	int64_t i0 = (int64_t)blockIdx.x * PREP_TILE + (int64_t)threadIdx.x * PREP_PER_THREAD;
	uint8_t s[PREP_PER_THREAD];
	uint32_t p[3] = {0, 0, 0}, ex[3]; /* p[w] holds counts of symbols 2w (low half) and 2w+1 (high half) */
	for (int j = 0; j < PREP_PER_THREAD; ++j) {
		int64_t i = i0 + j;
		s[j] = i < len ? bwt[i] : 7;
		if (s[j] < RB3B_ASIZE) p[s[j] >> 1] += 1u << (16 * (s[j] & 1));
	}
	Scan(tmp[0]).ExclusiveSum(p[0], ex[0]);
	Scan(tmp[1]).ExclusiveSum(p[1], ex[1]);
	Scan(tmp[2]).ExclusiveSum(p[2], ex[2]);
	int64_t base[RB3B_ASIZE];
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a)
		base[a] = accB.v[a] + (tex[(int64_t)a * (nt + 1) + blockIdx.x] - tex[(int64_t)a * (nt + 1)]) + ((ex[a >> 1] >> (16 * (a & 1))) & 0xffffu);
	for (int j = 0; j < PREP_PER_THREAD; ++j) {
		int64_t i = i0 + j;
		if (i >= len) break;
		int a = s[j];
		int64_t lf = 0;
#pragma unroll
		for (int b = 0; b < RB3B_ASIZE; ++b) if (a == b) lf = base[b]++;
		uint64_t mark = (i < accB.v[1] || i % fine_len == 0) ? LFB_FINE : 0u;
		lfb[i] = (uint64_t)lf << LFB_SHIFT | mark | (uint64_t)a;
	}
}

/* ------------------------------------------------------------------ */
/* balanced segmentation of the chains                                  */
/* ------------------------------------------------------------------ */

/* Fine marks: the sentinel rows (fine node = row) and every fine_len-th row.  They cut the chains into short
 * pieces of random length; walking them (LF_B only, no rank) and list-ranking the pieces gives every fine
 * mark its distance to the start of its sequence, from which evenly spaced coarse marks are chosen. */
struct Fine {
	int64_t n_fine, n_seq, fine_len, m0;
	__host__ __device__ int64_t row(int64_t f) const { return f < n_seq ? f : (m0 + (f - n_seq)) * fine_len; }
	__host__ __device__ int64_t of_row(int64_t r) const { return r < n_seq ? r : n_seq + (r / fine_len - m0); }
};

__global__ void k_fine_walk(Fine F, const uint64_t *__restrict__ lfb, int64_t *__restrict__ succ, int64_t *__restrict__ dist)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= F.n_fine) return;
	int64_t kb = F.row(f), n = 0, nx = -1;
	uint64_t x = __ldg(lfb + kb);
	for (;;) {
		++n;
		if ((x & 7) == 0) break;
		kb = (int64_t)(x >> LFB_SHIFT);
		x = __ldg(lfb + kb);
		if (x & LFB_FINE) { nx = F.of_row(kb); break; }
	}
	succ[f] = nx; dist[f] = n;
}

/* Wyllie pointer jumping: after ceil(log2 n) rounds dist[f] = #rows from fine mark f to the start of its sequence */
__global__ void k_list_rank(int64_t n, const int64_t *__restrict__ succ_in, const int64_t *__restrict__ dist_in, const int64_t *__restrict__ term_in,
                            int64_t *__restrict__ succ_out, int64_t *__restrict__ dist_out, int64_t *__restrict__ term_out)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n) return;
	int64_t s = succ_in[f], d = dist_in[f], t = term_in ? term_in[f] : f; /* t: last fine mark of f's chain seen so far */
	if (s >= 0) { d += dist_in[s]; t = term_in ? term_in[s] : s; s = succ_in[s]; }
	succ_out[f] = s; dist_out[f] = d; term_out[f] = t;
}

/* the sentinel fine mark p starts a chain: record the chain's length at its terminal mark */
__global__ void k_chain_len(int64_t n_seq, const int64_t *__restrict__ to_end, const int64_t *__restrict__ term, int64_t *__restrict__ chain_len)
{
	int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n_seq) chain_len[term[p]] = to_end[p];
}

/* Which segments does device `part` of `n_parts` walk?  Long chains are cut into n_parts contiguous stretches (by the
 * distance walked from the sentinel); a part also walks the `halo` rows before its stretch speculatively so that its
 * first segment receives an exact value without any exchange.  Short chains go to one part as a whole.
 * role: 0 not mine, 1 mine (counts for completeness), 2 halo (walked, result not required). */
__global__ void k_seg_role(Fine F, const int64_t *__restrict__ flag, const int64_t *__restrict__ sid, const int64_t *__restrict__ to_end,
                           const int64_t *__restrict__ term, const int64_t *__restrict__ chain_len, int part, int n_parts, int64_t halo, int64_t min_stretch,
                           uint8_t *__restrict__ role)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= F.n_fine || !flag[f]) return;
	int r = 1;
	if (n_parts > 1) {
		int64_t L = chain_len[term[f]], ds = L - to_end[f]; /* rows walked from the sentinel before this segment */
		if (L <= 0 || ds < 0) r = 1; /* not a valid BWT: let the completeness check report it */
		else if (L < min_stretch * n_parts) r = (int)(term[f] % n_parts) == part;
		else {
			int64_t q = ds * n_parts / L;
			if (q >= n_parts) q = n_parts - 1;
			if (q == part) r = 1;
			else {
				int64_t b = ((int64_t)part * L + n_parts - 1) / n_parts; /* first ds of my stretch */
				r = (part > 0 && ds < b && ds + halo >= b) ? 2 : 0;
			}
		}
	}
	role[sid[f]] = (uint8_t)r;
}

/* a fine mark becomes a coarse mark when the walk crosses a multiple of seg_len on the way to it */
__global__ void k_coarse_flag(Fine F, int64_t seg_len, const int64_t *__restrict__ succ, const int64_t *__restrict__ piece, const int64_t *__restrict__ to_end, int64_t *__restrict__ flag)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= F.n_fine) return;
	if (f < F.n_seq) flag[f] = 1;
	int64_t s = succ[f];
	if (s >= 0) {
		int64_t r = to_end[f], rs = r - piece[f];
		if (r / seg_len != rs / seg_len) flag[s] = 1;
	}
}

__global__ void k_coarse_fill(Fine F, const int64_t *__restrict__ flag, const int64_t *__restrict__ sid, int64_t *__restrict__ seg_row, int32_t *__restrict__ cmap, uint64_t *__restrict__ lfb)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= F.n_fine) return;
	if (flag[f]) {
		int64_t r = F.row(f);
		seg_row[sid[f]] = r;
		cmap[f] = (int32_t)sid[f];
		lfb[r] |= LFB_COARSE;
	} else cmap[f] = -1;
}

/* length of every segment = rows between its coarse mark and the next one (or the start of the sequence) */
__global__ void k_seg_len(Fine F, const int64_t *__restrict__ flag, const int64_t *__restrict__ sid, const int64_t *__restrict__ succ, const int64_t *__restrict__ piece, int64_t *__restrict__ seglen)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= F.n_fine || !flag[f]) return;
	int64_t n = 0, g = f;
	do { n += piece[g]; g = succ[g]; } while (g >= 0 && !flag[g]);
	seglen[sid[f]] = n;
}

/* ------------------------------------------------------------------ */
/* segmented LF walk                                                    */
/* ------------------------------------------------------------------ */

struct Segs {
	int64_t n_seg, n_seq;
	const int64_t *row;   /* first row of the segment (a coarse mark) */
	const int32_t *cmap;  /* fine node -> segment, -1 if the fine mark is not a coarse mark */
	int64_t *d;       /* #rows at the start of the segment that are still unresolved */
	int64_t *len;     /* #rows of the segment */
	int64_t *succ;    /* segment entered after the last row, or -1 at the start of a sequence */
	int64_t *arr_lo, *arr_hi; /* bracket with which the walk arrived at succ's first row */
	/* fix-up log (bitmap cells): while a walk carries a bracket it records, in walk order, the row, the bracket's low
	 * end, the row's symbol and whether the bracket is at most 128 wide.  The later rounds then stream this log
	 * instead of chasing LF_B, and for narrow brackets the two cells that can hold the exact position are known in
	 * advance, so their loads no longer sit on the dependent chain. */
	const uint8_t *role;    /* 0: another device walks this segment, 1: mine, 2: halo */
	const int64_t *logbase; /* per segment: first log slot; NULL = no log */
	longlong2 *log;         /* x = row, y = low end | symbol << 42 | narrow flag */
};

#define LOG_C_SHIFT 42
#define LOG_NARROW (1LL << 45)
#define LOG_NCELL 3              /* a "narrow" bracket spans at most LOG_NCELL consecutive cells */
#define LOG_WIDTH ((LOG_NCELL - 1) * 128)

/* The walk kernels are written for a "ranker" WG: Grp<8> (RLE cells, 8 lanes per walk) or BmRank (bitmap cells,
 * one thread per walk). */

/* round 1: every segment walks from its mark to the next mark */
template<class WG>
__global__ void __launch_bounds__(TPB) k_walk_first(DevIndex A, Acc7 accB, Segs S, Fine F, const uint64_t *__restrict__ lfb, int64_t *__restrict__ ka, int64_t *next_seg)
{
	const int gl = WG::lane(), gbase = WG::base();
	const unsigned gmask = WG::mask();
	for (;;) {
		int64_t s = 0;
		if (gl == 0) s = (int64_t)atomicAdd((unsigned long long*)next_seg, 1ULL);
		s = __shfl_sync(gmask, s, gbase);
		if (s >= S.n_seg) break;
		if (S.role[s] == 0) continue; /* group-uniform */
		int64_t kb = S.row[s], lo, hi, d = 0, len = 0, succ = -1;
		if (kb < S.n_seq) lo = hi = A.acc[1]; /* new sentinels sort after all old ones, fm-index.c:164 */
		else {
			int c0 = 1;
			while (c0 < RB3B_ASIZE - 1 && kb >= accB.v[c0 + 1]) ++c0;
			lo = A.acc[c0]; hi = A.acc[c0 + 1];
		}
		uint64_t x = __ldg(lfb + kb);
		for (;;) {
			int c = (int)(x & 7);
			if (lo == hi) { if (gl == 0) ka[kb] = lo; }
			else {
				if (S.logbase && gl == 0) {
					int64_t o = S.logbase[s] + d;
					S.log[o] = make_longlong2(kb, lo | (int64_t)c << LOG_C_SHIFT | (hi - lo <= LOG_WIDTH ? LOG_NARROW : 0));
				}
				++d;
			}
			++len;
			if (c == 0) break; /* reached the first symbol of the sequence, fm-index.c:170 */
			kb = (int64_t)(x >> LFB_SHIFT);
			x = __ldg(lfb + kb); /* the B chain does not depend on A: fetch one step ahead */
			int64_t r1, r2;
			/* the groups that currently run together take the two-position path only while one of them still
			 * carries a bracket; either path is correct for an exact group, so this is purely a cost choice */
			/* (a private per-thread branch was measured slower for one-thread walks too: the two paths serialise) */
			if (__any_sync(__activemask(), lo != hi)) WG::rank2(A, lo, hi, c, r1, r2);
			else r1 = r2 = WG::rank(A, lo, c);
			lo = A.acc[c] + r1; hi = A.acc[c] + r2;
			if (x & LFB_COARSE) { succ = S.cmap[F.of_row(kb)]; break; }
		}
		if (gl == 0) { S.d[s] = d; S.len[s] = len; S.succ[s] = succ; S.arr_lo[s] = lo; S.arr_hi[s] = hi; }
	}
}

/* after round 1: segments whose predecessor arrived with an exact value and that have unresolved rows */
__global__ void k_collect_first(Segs S, int64_t *__restrict__ wl_seg, int64_t *__restrict__ wl_val, unsigned long long *wl_n)
{
	int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= S.n_seg || S.role[s] == 0) return;
	int64_t t = S.succ[s];
	if (t >= 0 && S.arr_lo[s] == S.arr_hi[s] && S.d[t] > 0) {
		unsigned long long o = atomicAdd(wl_n, 1ULL);
		wl_seg[o] = t; wl_val[o] = S.arr_lo[s];
	}
}

/* one logged row held by one lane; for a narrow bracket everything that does not depend on the exact value is
 * precomputed: T[w] = #c between the window start and 32-bit word w of the window, W[w] = that word of plane c */
struct LogRow {
	int64_t kb, pos0, base; /* row; first position of the window; C[c] + #c before the window */
	int c, narrow;
	uint32_t T[4 * LOG_NCELL], W[4 * LOG_NCELL];
	__device__ __forceinline__ void load(const DevIndex &A, const Segs &S, int64_t slot, bool valid)
	{
		kb = 0; pos0 = 0; base = 0; c = 0; narrow = 0;
		if (!valid) return;
		const longlong2 rec = S.log[slot];
		kb = rec.x;
		int64_t w = rec.y, lo = w & (int64_t)RB3B_M42;
		c = (int)(w >> LOG_C_SHIFT) & 7; narrow = (w & LOG_NARROW) != 0;
		if (!narrow) return;
		const int h = c >= 3, cc = c - 3 * h;
		const int64_t j = (lo < A.n ? lo : A.n - 1) >> RB3B_BM_SHIFT;
		uint4 cq[LOG_NCELL], pq[LOG_NCELL];
#pragma unroll
		for (int i = 0; i < LOG_NCELL; ++i) { /* loads independent of the chain */
			const int64_t ji = j + i < A.n_cells ? j + i : A.n_cells - 1;
			cq[i] = __ldg(A.cells + ji * 8 + 4 * h); pq[i] = __ldg(A.cells + ji * 8 + 4 * h + 1 + cc);
		}
		uint64_t h0 = 0;
		uint32_t run = 0;
#pragma unroll
		for (int i = 0; i < LOG_NCELL; ++i) {
			uint64_t a0, a1, a2;
			rb3b_hdr_unpack(cq[i], a0, a1, a2);
			uint64_t hi = cc == 0 ? a0 : cc == 1 ? a1 : a2;
			if (i == 0) h0 = hi;
			const bool real = j + i < A.n_cells; /* past the end of the index: no symbols, the count stays */
			if (real) run = (uint32_t)(hi - h0);
			const uint32_t ww[4] = { pq[i].x, pq[i].y, pq[i].z, pq[i].w };
#pragma unroll
			for (int k = 0; k < 4; ++k) { T[4 * i + k] = run; W[4 * i + k] = real ? ww[k] : 0u; run += real ? __popc(ww[k]) : 0; }
		}
		pos0 = j << RB3B_BM_SHIFT;
		base = A.acc[c] + (int64_t)h0;
	}
	/* C[c] + rank(c, v) for the exact value v, which lies inside the window */
	__device__ __forceinline__ int64_t step(const DevIndex &A, int64_t v) const
	{
		if (v >= A.n) return A.acc[c] + A.tot[c];
		if (!narrow) return A.acc[c] + BmRank::rank(A, v, c);
		uint32_t off = (uint32_t)(v - pos0), w = off >> 5;
		return base + T[w] + __popc(W[w] & ((1u << (off & 31u)) - 1u));
	}
};

/* shared-memory image of the 32 rows a warp works on */
struct FixRows {
	uint2 TW[32][4 * LOG_NCELL];   /* per row: x = word of plane c, y = #c between the window start and that word */
	int64_t kb[32], pos0[32], base[32];
	int32_t K[32];                 /* base - first position of the NEXT row's window: keeps the chain in 32 bits */
	uint32_t off[32];              /* result: exact value of the row relative to its window */
	int32_t c[32];
};

/* round >= 2 with the fix-up log (bitmap cells): one WARP per listed segment.  Each lane fetches one logged row
 * (coalesced) and the cells it may need, one iteration (32 rows) ahead, and publishes the row's count/word tables in
 * shared memory.  Lane 0 then runs the dependent chain.  For a row with a narrow bracket the exact value is carried as a
 * 32-bit offset into the row's cell window: off' = K + T[off >> 5] + popc(W[off >> 5] below off), one 8-byte
 * shared-memory read and ~10 instructions per row (the kernel is bound by the instruction count of this serial
 * chain: handing the value from lane to lane with shuffles cost 56 warp instructions per row).  Afterwards all
 * lanes write the interleave positions of their rows.  Rows whose bracket was wider than the window (the first ~10 of
 * a segment) take the general path with a random cell access. */
__global__ void __launch_bounds__(128) k_walk_fix_log(DevIndex A, Segs S, int64_t *__restrict__ ka, int64_t n_items,
                                                       const int64_t *__restrict__ wl_seg, const int64_t *__restrict__ wl_val,
                                                       int64_t *__restrict__ nx_seg, int64_t *__restrict__ nx_val, unsigned long long *nx_n)
{
	__shared__ FixRows rows[4]; /* four warps per block */
	const int lane = threadIdx.x & 31;
	FixRows &R = rows[threadIdx.x >> 5];
	int64_t it = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	if (it >= n_items) return; /* warp-uniform */
	int64_t t = wl_seg[it], v = wl_val[it];
	unsigned long long n_rows = 0, n_wide = 0; /* statistics only */
	for (;;) { /* a segment that never collapsed hands its exact arrival straight to its successor: same warp, no new launch */
		const int64_t d = S.d[t], len = S.len[t], base = S.logbase[t];
		int ended = 0;
		LogRow nxt;
		nxt.load(A, S, base + lane, lane < d);
		n_rows += (unsigned long long)d;
		for (int64_t i0 = 0; i0 < d && !ended; i0 += 32) {
			const int cnt = d - i0 < 32 ? (int)(d - i0) : 32;
			const bool live = lane < cnt;
			/* this lane's row: fast = its step can be taken from the tables */
			const bool fast = live && nxt.narrow && nxt.c != 0;
			const int64_t my_kb = nxt.kb, my_pos0 = nxt.pos0, my_base = nxt.base;
			const int64_t next_pos0 = __shfl_down_sync(0xffffffffu, nxt.pos0, 1);
			const unsigned fastmask = __ballot_sync(0xffffffffu, fast);
			/* link: the next row of this iteration is fast too, so the value can stay relative */
			const bool link = fast && lane + 1 < cnt && (fastmask >> (lane + 1) & 1u);
			const unsigned linkmask = __ballot_sync(0xffffffffu, link);
			n_wide += __popc(__ballot_sync(0xffffffffu, live && !nxt.narrow));
#pragma unroll
			for (int k = 0; k < 4 * LOG_NCELL; ++k) R.TW[lane][k] = make_uint2(nxt.W[k], nxt.T[k]);
			R.kb[lane] = my_kb; R.pos0[lane] = my_pos0; R.base[lane] = my_base; R.c[lane] = nxt.c;
			R.K[lane] = link ? (int32_t)(my_base - next_pos0) : 0;
			R.off[lane] = 0xffffffffu;
			nxt.load(A, S, base + i0 + 32 + lane, i0 + 32 + lane < d); /* fetch the next iteration's row meanwhile */
			__syncwarp();
			if (lane == 0) {
				int u = 0;
				while (u < cnt) {
					if (fastmask >> u & 1u) { /* a run of table rows: 32-bit relative chain */
						uint32_t off = (uint32_t)(v - R.pos0[u]);
						for (;;) {
							R.off[u] = off;
							const uint2 tw = R.TW[u][off >> 5];
							const uint32_t r = tw.y + __popc(tw.x & ((1u << (off & 31u)) - 1u));
							if (!(linkmask >> u & 1u)) { v = R.base[u] + r; ++u; break; }
							off = (uint32_t)(R.K[u] + (int32_t)r);
							++u;
						}
					} else { /* general row */
						ka[R.kb[u]] = v;
						const int c = R.c[u];
						if (c == 0) { ended = 1; break; }
						v = A.acc[c] + BmRank::rank(A, v, c);
						++u;
					}
				}
			}
			__syncwarp();
			if (R.off[lane] != 0xffffffffu) ka[my_kb] = my_pos0 + R.off[lane]; /* rows resolved through the tables */
			v = __shfl_sync(0xffffffffu, v, 0);
			ended = __shfl_sync(0xffffffffu, ended, 0);
			__syncwarp(); /* the rows are consumed before they are overwritten */
		}
		const int64_t u2 = S.succ[t];
		if (lane == 0) {
			S.d[t] = 0;
			if (d == len && u2 >= 0) S.arr_lo[t] = S.arr_hi[t] = v;
		}
		if (!(d == len && u2 >= 0 && S.d[u2] > 0)) break; /* u2's only predecessor is t: nobody else touches it */
		t = u2;
	}
	if (lane == 0) { /* nx_n doubles as a statistics block: [1] rows, [2] rows with a wide bracket, [3] longest chain */
		atomicAdd(nx_n + 1, n_rows); atomicAdd(nx_n + 2, n_wide); atomicMax(nx_n + 3, n_rows);
	}
	(void)nx_seg; (void)nx_val;
}

/* round >= 2: re-walk the unresolved prefix of each listed segment from its now exact start */
template<class WG>
__global__ void __launch_bounds__(TPB) k_walk_fix(DevIndex A, Segs S, const uint64_t *__restrict__ lfb, int64_t *__restrict__ ka, int64_t n_items,
                                                   const int64_t *__restrict__ wl_seg, const int64_t *__restrict__ wl_val, int64_t *next_item,
                                                   int64_t *__restrict__ nx_seg, int64_t *__restrict__ nx_val, unsigned long long *nx_n)
{
	const int gl = WG::lane(), gbase = WG::base();
	const unsigned gmask = WG::mask();
	for (;;) {
		int64_t it = 0;
		if (gl == 0) it = (int64_t)atomicAdd((unsigned long long*)next_item, 1ULL);
		it = __shfl_sync(gmask, it, gbase);
		if (it >= n_items) break;
		int64_t t = wl_seg[it], v = wl_val[it], kb = S.row[t], d = S.d[t], len = S.len[t];
		uint64_t x = __ldg(lfb + kb);
		for (int64_t i = 0; i < d; ++i) {
			int c = (int)(x & 7);
			if (gl == 0) ka[kb] = v;
			if (c == 0) break;
			kb = (int64_t)(x >> LFB_SHIFT);
			x = __ldg(lfb + kb);
			v = A.acc[c] + WG::rank(A, v, c);
		}
		if (gl == 0) {
			S.d[t] = 0;
			int64_t u = S.succ[t];
			if (d == len && u >= 0) { /* never collapsed: only now is the arrival value known */
				S.arr_lo[t] = S.arr_hi[t] = v;
				if (S.d[u] > 0) { /* u is processed by nobody else in this round: its only predecessor is t */
					unsigned long long o = atomicAdd(nx_n, 1ULL);
					nx_seg[o] = u; nx_val[o] = v;
				}
			}
		}
	}
}

__global__ void k_seg_check(Segs S, unsigned long long *sums)
{
	int64_t s = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= S.n_seg) return;
	if (S.role[s] != 1) return;
	if (S.d[s]) atomicAdd(&sums[0], (unsigned long long)S.d[s]);
	atomicAdd(&sums[1], (unsigned long long)S.len[s]);
	atomicAdd(&sums[2], 1ULL);
}

/* rb[i] = (ka+i)<<6 | B[i]<<3 | first symbol of suffix i (fm-index.c:168) */
__global__ void k_pack_rb(int64_t len, const uint8_t *__restrict__ bwt, const int64_t *__restrict__ ka, Acc7 accB, int64_t *__restrict__ rb)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= len) return;
	int c0 = 0;
	while (c0 < RB3B_ASIZE - 1 && i >= accB.v[c0 + 1]) ++c0;
	rb[i] = (ka[i] + i) << 6 | (int64_t)bwt[i] << 3 | c0;
}

__global__ void k_check_monotone(int64_t len, const int64_t *__restrict__ ka, int64_t nA, int *bad)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= len) return;
	int64_t v = ka[i];
	if (v < 0 || v > nA || (i > 0 && ka[i - 1] > v)) *bad = 1;
}

int64_t rb3b_get_param(const char *key, int64_t dflt); /* rb3b_runtime.cu */

/* interleave positions of the batch in device memory: ka[len], accB */
/* part/n_parts: see k_seg_role.  ka_out != NULL: write there instead of allocating.  *incomplete is set when a part could
 * not resolve all of its own rows locally (only possible with n_parts > 1). */
static int rank_phase(const rb3b_index_s *A, int64_t len, const uint8_t *d_bwt, DBuf<int64_t> &ka, int64_t accB[RB3B_ASIZE + 1],
                      int part = 0, int n_parts = 1, int64_t *ka_out = 0, int *incomplete = 0)
{
	int64_t nt = (len + PREP_TILE - 1) / PREP_TILE;
	DBuf<int64_t> tcnt, tex;
	DBuf<int> bad;
	DBuf<uint64_t> lfb;
	int hbad = 0;
	if (len <= 0) return rb3b_fail(RB3B_EINVAL, "empty batch");
	if (A->n_cells == 0) return rb3b_fail(RB3B_EINVAL, "rank phase on an empty index");
	/* batch LF mapping */
	TRY(tcnt.alloc((nt + 1) * RB3B_ASIZE)); TRY(tex.alloc((nt + 1) * RB3B_ASIZE)); TRY(bad.alloc(1)); TRY(lfb.alloc(len));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	rb3b_tic(T_PREP);
	k_prep_count<<<(unsigned)nt, TPB, 0, rb3b_stream>>>(len, d_bwt, nt, tcnt.p, bad.p); CKK();
	TRY(rb3b_scan_excl_i64(tcnt.p, tex.p, (nt + 1) * RB3B_ASIZE));
	int64_t tot[RB3B_ASIZE], base[RB3B_ASIZE];
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		CK(cudaMemcpyAsync(&tot[a], tex.p + a * (nt + 1) + nt, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaMemcpyAsync(&base[a], tex.p + a * (nt + 1), 8, cudaMemcpyDeviceToHost, rb3b_stream));
	}
	CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (hbad) return rb3b_fail(RB3B_EINVAL, "batch BWT holds a symbol >= %d", RB3B_ASIZE);
	Acc7 acc;
	acc.v[0] = 0;
	for (int a = 0; a < RB3B_ASIZE; ++a) acc.v[a + 1] = acc.v[a] + (tot[a] - base[a]);
	memcpy(accB, acc.v, sizeof(acc.v));
	if (acc.v[1] <= 0) return rb3b_fail(RB3B_EINVAL, "batch BWT holds no sentinel");
	/* fine marks, their list ranks, coarse marks */
	int64_t seg_len = rb3b_seg_len;
	Fine F;
	F.n_seq = acc.v[1];
	F.fine_len = rb3b_get_param("fine_len", 32);
	if (F.fine_len > seg_len) F.fine_len = seg_len;
	if (F.fine_len < 1) F.fine_len = 1;
	F.m0 = (F.n_seq + F.fine_len - 1) / F.fine_len;
	int64_t n_samp = (len - 1) / F.fine_len - F.m0 + 1;
	F.n_fine = F.n_seq + (n_samp > 0 ? n_samp : 0);
	k_prep_lf<<<(unsigned)nt, TPB, 0, rb3b_stream>>>(len, d_bwt, nt, tex.p, acc, F.fine_len, lfb.p); CKK();
	DBuf<int64_t> fn; /* succ, piece, 2 x (succ, dist, term) ping-pong, flag, sid, chain_len */
	DBuf<int32_t> cmap;
	TRY(fn.alloc(F.n_fine * 11)); TRY(cmap.alloc(F.n_fine));
	int64_t *f_succ = fn.p, *f_piece = fn.p + F.n_fine;
	int64_t *pp[2][3] = { { fn.p + 2 * F.n_fine, fn.p + 3 * F.n_fine, fn.p + 4 * F.n_fine }, { fn.p + 5 * F.n_fine, fn.p + 6 * F.n_fine, fn.p + 7 * F.n_fine } };
	int64_t *f_flag = fn.p + 8 * F.n_fine, *f_sid = fn.p + 9 * F.n_fine, *f_clen = fn.p + 10 * F.n_fine;
	k_fine_walk<<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, lfb.p, f_succ, f_piece); CKK();
	const int64_t *cs = f_succ, *cd = f_piece, *ct = 0;
	int cur = 0;
	for (int64_t span = 1; span < F.n_fine || ct == 0; span <<= 1) { /* at least once: it also initialises the terminal marks */
		k_list_rank<<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F.n_fine, cs, cd, ct, pp[cur][0], pp[cur][1], pp[cur][2]); CKK();
		cs = pp[cur][0]; cd = pp[cur][1]; ct = pp[cur][2]; cur ^= 1;
	}
	CK(cudaMemsetAsync(f_flag, 0, F.n_fine * 8, rb3b_stream));
	k_coarse_flag<<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, seg_len, f_succ, f_piece, cd, f_flag); CKK();
	TRY(rb3b_scan_excl_i64(f_flag, f_sid, F.n_fine));
	int64_t last[2];
	CK(cudaMemcpyAsync(&last[0], f_sid + F.n_fine - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&last[1], f_flag + F.n_fine - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	/* segments */
	Segs S;
	S.n_seq = F.n_seq;
	S.n_seg = last[0] + last[1];
	DBuf<int64_t> seg, wl, ctr;
	DBuf<uint8_t> role;
	TRY(seg.alloc(S.n_seg * 6)); TRY(wl.alloc(S.n_seg * 4)); TRY(ctr.alloc(8)); TRY(role.alloc(S.n_seg));
	if (ka_out) ka.p = ka_out; else TRY(ka.alloc(len));
	CK(cudaMemsetAsync(seg.p, 0, S.n_seg * 5 * 8, rb3b_stream)); /* d = 0 for the segments other devices walk */
	k_chain_len<<<nblk(F.n_seq, TPB), TPB, 0, rb3b_stream>>>(F.n_seq, cd, ct, f_clen); CKK();
	k_seg_role<<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, f_flag, f_sid, cd, ct, f_clen, part, n_parts,
		rb3b_get_param("halo_segments", 8) * seg_len, 4 * seg_len, role.p); CKK();
	S.role = role.p;
	S.d = seg.p; S.len = seg.p + S.n_seg; S.succ = seg.p + 2 * S.n_seg; S.arr_lo = seg.p + 3 * S.n_seg; S.arr_hi = seg.p + 4 * S.n_seg;
	S.row = seg.p + 5 * S.n_seg; S.cmap = cmap.p;
	k_coarse_fill<<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, f_flag, f_sid, seg.p + 5 * S.n_seg, cmap.p, lfb.p); CKK();
	const bool bm = A->kind == RB3B_KIND_BM;
	DBuf<int64_t> seglen, logbase, logbuf;
	S.logbase = 0; S.log = 0;
	if (bm && rb3b_get_param("fix_log", 1) && len * 16 <= rb3b_get_param("fix_log_max_bytes", 16LL << 30)) {
		TRY(seglen.alloc(S.n_seg)); TRY(logbase.alloc(S.n_seg)); TRY(logbuf.alloc(2 * len));
		k_seg_len<<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, f_flag, f_sid, f_succ, f_piece, seglen.p); CKK();
		TRY(rb3b_scan_excl_i64(seglen.p, logbase.p, S.n_seg));
		S.logbase = logbase.p; S.log = (longlong2*)logbuf.p;
	}
	rb3b_toc(T_PREP);
	CK(cudaMemsetAsync(ctr.p, 0, 8 * 8, rb3b_stream));
	CK(cudaMemsetAsync(ka.p, 0xff, len * 8, rb3b_stream));
	DevIndex dA = rb3b_dev_view(A);
	/* the batch LF table is hit once per row at random: keep it resident in L2 while the index cells stream through */
	const bool pin = rb3b_get_param("pin_lfb", 0) != 0; /* measured slower on B200 (set-aside shrinks the L2 left for cells, ka and the log): off by default */
	if (pin) rb3b_l2_pin(lfb.p, (size_t)len * 8);
	/* bitmap walks are single threads: small CTAs spread the few thousand walks over all SMs */
	const int wg = bm ? 1 : 8, wtpb = bm ? 32 : TPB;
	int64_t want = (S.n_seg * wg + wtpb - 1) / wtpb, cap = (int64_t)n_sm() * (bm ? 32 : 8);
	rb3b_tic(T_WALK1);
	if (bm) k_walk_first<BmRank><<<(unsigned)(want < cap ? want : cap), wtpb, 0, rb3b_stream>>>(dA, acc, S, F, lfb.p, ka.p, ctr.p);
	else k_walk_first<Grp<8> ><<<(unsigned)(want < cap ? want : cap), wtpb, 0, rb3b_stream>>>(dA, acc, S, F, lfb.p, ka.p, ctr.p);
	CKK();
	rb3b_toc(T_WALK1);
	int64_t *wl_seg[2] = { wl.p, wl.p + 2 * S.n_seg }, *wl_val[2] = { wl.p + S.n_seg, wl.p + 3 * S.n_seg };
	k_collect_first<<<nblk(S.n_seg, TPB), TPB, 0, rb3b_stream>>>(S, wl_seg[0], wl_val[0], (unsigned long long*)(ctr.p + 1)); CKK();
	int64_t n_items = 0, rounds = 1, fix_rows = 0;
	CK(cudaMemcpyAsync(&n_items, ctr.p + 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	cur = 0;
	while (n_items > 0) {
		/* ctr[2] = item cursor, ctr[3] = size of the next list */
		CK(cudaMemsetAsync(ctr.p + 2, 0, 40, rb3b_stream));
		want = (n_items * wg + wtpb - 1) / wtpb;
		rb3b_tic(T_WALKFIX);
		if (bm && S.logbase) k_walk_fix_log<<<nblk(n_items * 32, 128), 128, 0, rb3b_stream>>>(dA, S, ka.p, n_items, wl_seg[cur], wl_val[cur],
			wl_seg[cur ^ 1], wl_val[cur ^ 1], (unsigned long long*)(ctr.p + 3));
		else if (bm) k_walk_fix<BmRank><<<(unsigned)(want < cap ? want : cap), wtpb, 0, rb3b_stream>>>(dA, S, lfb.p, ka.p, n_items, wl_seg[cur], wl_val[cur], ctr.p + 2,
			wl_seg[cur ^ 1], wl_val[cur ^ 1], (unsigned long long*)(ctr.p + 3));
		else k_walk_fix<Grp<8> ><<<(unsigned)(want < cap ? want : cap), wtpb, 0, rb3b_stream>>>(dA, S, lfb.p, ka.p, n_items, wl_seg[cur], wl_val[cur], ctr.p + 2,
			wl_seg[cur ^ 1], wl_val[cur ^ 1], (unsigned long long*)(ctr.p + 3));
		CKK();
		rb3b_toc(T_WALKFIX);
		fix_rows += n_items;
		int64_t fst[4] = {0, 0, 0, 0};
		CK(cudaMemcpyAsync(fst, ctr.p + 3, 32, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		n_items = fst[0];
		if (bm && S.logbase) { rb3b_stat_set("fix_rows", fst[1]); rb3b_stat_set("fix_wide_rows", fst[2]); rb3b_stat_set("fix_longest_chain", fst[3]); }
		rb3b_tflush();
		cur ^= 1; ++rounds;
	}
	if (pin) rb3b_l2_pin(0, 0);
	unsigned long long sums[3];
	CK(cudaMemsetAsync(ctr.p + 4, 0, 24, rb3b_stream));
	k_seg_check<<<nblk(S.n_seg, TPB), TPB, 0, rb3b_stream>>>(S, (unsigned long long*)(ctr.p + 4)); CKK();
	CK(cudaMemcpyAsync(sums, ctr.p + 4, 24, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	rb3b_tflush();
	rb3b_stat_set("n_segments", S.n_seg);
	rb3b_stat_set("n_fine", F.n_fine);
	rb3b_stat_set("fix_rounds", rounds - 1);
	rb3b_stat_add("fix_rounds_total", rounds - 1);
	rb3b_stat_set("fix_segments", fix_rows);
	rb3b_stat_set("unresolved_rows", (int64_t)sums[0]);
	rb3b_stat_set("own_segments", (int64_t)sums[2]);
	rb3b_stat_set("own_rows", (int64_t)sums[1]);
	if (n_parts > 1) { /* completeness of the whole batch is checked after the exchange */
		if (incomplete) *incomplete = sums[0] != 0;
		return RB3B_OK;
	}
	if (sums[0] != 0 || (int64_t)sums[1] != len)
		return rb3b_fail(RB3B_EINVAL, "batch is not the BWT of a sentinel-terminated string set (%lld of %lld rows reachable, %lld unresolved)",
		                 (long long)sums[1], (long long)len, (long long)sums[0]);
	return RB3B_OK;
}

/* ------------------------------------------------------------------ */
/* streaming merge (rb3b_emit.cuh)                                      */
/* ------------------------------------------------------------------ */

static int merge_phase(rb3b_index_s *A, int64_t len, const uint8_t *d_bwt, const int64_t *d_ka)
{
	DBuf<int> bad;
	int hbad = 0;
	TRY(bad.alloc(1));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	rb3b_tic(T_MERGE);
	k_check_monotone<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, d_ka, A->n, bad.p); CKK();
	CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (hbad) return rb3b_fail(RB3B_EINVAL, "interleave positions are not monotone: the batch is not a valid BWT");
	int rc;
	if (A->kind == RB3B_KIND_BM) {
		BmSrc src;
		src.R.cells = A->cells; src.R.n = A->n; src.R.pos = 0; src.R.cur = -1; src.R.rem = 0;
		src.n = A->n; src.cur = -1;
		if (rb3b_want_bitmap(A->n + len)) rc = rb3b_emit_build_bm(A, src, A->n, len, d_ka, d_bwt);
		else rc = rb3b_emit_build(A, src, A->n, len, d_ka, d_bwt, (A->n + len) / 16); /* outgrew 1 B/symbol: switch to RLE cells */
	} else {
		CellSrc src;
		src.R.cells = A->cells; src.R.ovf = A->ovf; src.R.n = A->n; src.R.shift = A->shift;
		src.R.j = 0; src.R.left = 0; src.R.eidx = 0; src.R.first = 0xffffffffu; src.R.cur = -1; src.R.rem = 0;
		src.n = A->n; src.cur = -1;
		/* every inserted row adds at most two entries; most extend or split a run of a shared column */
		rc = rb3b_emit_build(A, src, A->n, len, d_ka, d_bwt, A->n_entries + len / 2);
	}
	rb3b_toc(T_MERGE);
	cudaStreamSynchronize(rb3b_stream);
	rb3b_tflush();
	return rc;
}

/* ------------------------------------------------------------------ */
/* C ABI                                                                */
/* ------------------------------------------------------------------ */

extern "C" int rb3b_mg_rank_plain_dev(const rb3b_index_t *x, int64_t len, const uint8_t *d_bwt, int64_t *d_rb, int64_t acc[RB3B_ASIZE + 1])
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	DBuf<int64_t> ka;
	Acc7 a;
	TRY(rank_phase(x, len, d_bwt, ka, a.v));
	k_pack_rb<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, d_bwt, ka.p, a, d_rb); CKK();
	if (acc) memcpy(acc, a.v, sizeof(a.v));
	return RB3B_OK;
}

extern "C" int rb3b_mg_rank_plain(const rb3b_index_t *x, int64_t len, const uint8_t *bwt, int64_t *rb, int64_t acc[RB3B_ASIZE + 1])
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> d; DBuf<int64_t> drb;
	if (len <= 0) return rb3b_fail(RB3B_EINVAL, "empty batch");
	TRY(d.alloc(len)); TRY(drb.alloc(len));
	CK(cudaMemcpyAsync(d.p, bwt, len, cudaMemcpyHostToDevice, rb3b_stream));
	TRY(rb3b_mg_rank_plain_dev(x, len, d.p, drb.p, acc));
	CK(cudaMemcpyAsync(rb, drb.p, len * 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}

/* ---- multi-device building blocks: index replicated, rows of the batch split among the devices ---- */

extern "C" int rb3b_mg_rank_part(const rb3b_index_t *x, int64_t len, const uint8_t *d_bwt, int part, int n_parts, int64_t *d_ka)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (n_parts < 1 || part < 0 || part >= n_parts) return rb3b_fail(RB3B_EINVAL, "part %d of %d", part, n_parts);
	DBuf<int64_t> ka;
	int64_t accB[RB3B_ASIZE + 1];
	int incomplete = 0;
	TRY(rank_phase(x, len, d_bwt, ka, accB, part, n_parts, d_ka, &incomplete));
	return incomplete ? 1 : RB3B_OK;
}

extern "C" int rb3b_merge_with_ka(rb3b_index_t *x, int64_t len, const uint8_t *d_bwt, const int64_t *d_ka)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (len <= 0) return RB3B_OK;
	if (x->n_cells == 0) return rb3b_fail(RB3B_EINVAL, "merge_with_ka on an empty index");
	return merge_phase(x, len, d_bwt, d_ka); /* rejects arrays with holes (-1) or out of order */
}

extern "C" int rb3b_merge_plain_dev(rb3b_index_t *x, int64_t len, const uint8_t *d_bwt)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (len <= 0) return RB3B_OK;
	if (x->n_cells == 0) return rb3b_index_from_plain_dev(x, len, d_bwt);
	DBuf<int64_t> ka;
	int64_t accB[RB3B_ASIZE + 1];
	TRY(rank_phase(x, len, d_bwt, ka, accB));
	return merge_phase(x, len, d_bwt, ka.p);
}

extern "C" int rb3b_merge_plain(rb3b_index_t *x, int64_t len, const uint8_t *bwt)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> d;
	if (len <= 0) return RB3B_OK;
	TRY(d.alloc(len));
	CK(cudaMemcpyAsync(d.p, bwt, len, cudaMemcpyHostToDevice, rb3b_stream));
	TRY(rb3b_merge_plain_dev(x, len, d.p));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}
