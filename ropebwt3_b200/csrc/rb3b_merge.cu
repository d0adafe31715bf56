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
 * The reference walks one dependent LF chain per new sequence, and every step of
 * it chases two things at once: LF_B (where is the next row of B) and rank_A
 * (where does it go in A).  Here the two are separated:
 *
 *   1. B only: the chains of LF_B are cut at "fine marks" (the sentinel rows and
 *      every fine_len-th row), the pieces are walked and LIST-RANKED (Wyllie
 *      pointer jumping), which gives every piece its offset in WALK ORDER.  A
 *      second walk of the pieces writes the batch in walk order: wsym[p] = the
 *      p-th symbol met by the reference's loop (all sequences one after the
 *      other), wrow[p] = the row it was met at.  This is the text of the batch
 *      read backwards and its inverse suffix array, recovered from the BWT
 *      alone; it touches only the batch's own LF table.
 *   2. A only: walk order is cut into equal SLICES of seg_len positions; one
 *      walk per slice streams its symbols (sequential, no dependent load) and
 *      performs the rank chain on A: one random 64-B access per row.  A slice
 *      that does not start at a sentinel does not know its ka yet, so it starts
 *      with the bracket [lo,hi] = SA interval (in A) of its first symbol and
 *      narrows it by backward search with the symbols it walks over; once
 *      lo == hi the value is exact (the walked string no longer occurs in A)
 *      and independent of anything before it.  Rows walked before the collapse
 *      stay unresolved (flagged in kseq[]) and are filled in by the fix-up pass
 *      from the exact value with which the previous slice arrived.  A batch
 *      sequence that is an exact substring of A degenerates to the sequential
 *      chain, still correct.
 *   3. ka[wrow[p]] = kseq[p]: one scatter back to row order.
 *
 * Phase B is a streaming merge: ka[] is non-decreasing, so every output cell is
 * produced independently from a slice of A and a slice of the batch.
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
/* lf[i] = LF_B(i) << 3 | B[i]; 32-bit entries while len < 2^29         */
/* ------------------------------------------------------------------ */

#define LF_SHIFT 3
#define LF32_MAX_LEN (1LL << 29)

/* the 16 symbols of one thread: one 16-byte load when the batch is 16-byte aligned (PREP_PER_THREAD == 16), 7 = past the end */
__device__ __forceinline__ void prep_load16(const uint8_t *__restrict__ bwt, int64_t len, int64_t i0, uint8_t (&s)[PREP_PER_THREAD])
{
	if (i0 + PREP_PER_THREAD <= len && (((uintptr_t)bwt) & 15) == 0) {
		const uint4 v = __ldg((const uint4*)(bwt + i0));
		const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
		for (int j = 0; j < PREP_PER_THREAD; ++j) s[j] = (uint8_t)(w[j >> 2] >> (8 * (j & 3)));
	} else {
#pragma unroll
		for (int j = 0; j < PREP_PER_THREAD; ++j) s[j] = i0 + j < len ? bwt[i0 + j] : 7;
	}
}

__global__ void __launch_bounds__(TPB) k_prep_count(int64_t len, const uint8_t *__restrict__ bwt, int64_t nt, int64_t *__restrict__ tcnt, int *__restrict__ bad)
{
	__shared__ unsigned int sh[RB3B_ASIZE];
	if (threadIdx.x < RB3B_ASIZE) sh[threadIdx.x] = 0;
	__syncthreads();
	int64_t i0 = (int64_t)blockIdx.x * PREP_TILE + (int64_t)threadIdx.x * PREP_PER_THREAD;
	unsigned int c[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
	uint8_t s16[PREP_PER_THREAD];
	prep_load16(bwt, len, i0, s16);
#pragma unroll
	for (int j = 0; j < PREP_PER_THREAD; ++j) {
		if (i0 + j < len) {
			int a = s16[j];
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

template<typename LfT>
__global__ void __launch_bounds__(TPB) k_prep_lf(int64_t len, const uint8_t *__restrict__ bwt, int64_t nt, const int64_t *__restrict__ tex,
                                                  Acc7 accB, LfT *__restrict__ lf)
{
	typedef cub::BlockScan<uint32_t, TPB> Scan;
	__shared__ typename Scan::TempStorage tmp[3];
	int64_t i0 = (int64_t)blockIdx.x * PREP_TILE + (int64_t)threadIdx.x * PREP_PER_THREAD;
	uint8_t s[PREP_PER_THREAD];
	uint32_t p[3] = {0, 0, 0}, ex[3]; /* p[w] holds counts of symbols 2w (low half) and 2w+1 (high half) */
	prep_load16(bwt, len, i0, s);
#pragma unroll
	for (int j = 0; j < PREP_PER_THREAD; ++j) {
		if (i0 + j >= len) s[j] = 7;
		if (s[j] < RB3B_ASIZE) p[s[j] >> 1] += 1u << (16 * (s[j] & 1));
	}
	Scan(tmp[0]).ExclusiveSum(p[0], ex[0]);
	Scan(tmp[1]).ExclusiveSum(p[1], ex[1]);
	Scan(tmp[2]).ExclusiveSum(p[2], ex[2]);
	int64_t base[RB3B_ASIZE];
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a)
		base[a] = accB.v[a] + (tex[(int64_t)a * (nt + 1) + blockIdx.x] - tex[(int64_t)a * (nt + 1)]) + ((ex[a >> 1] >> (16 * (a & 1))) & 0xffffu);
	LfT out[PREP_PER_THREAD];
#pragma unroll
	for (int j = 0; j < PREP_PER_THREAD; ++j) {
		const int a = s[j];
		int64_t v = 0;
#pragma unroll
		for (int b = 0; b < RB3B_ASIZE; ++b) if (a == b) v = base[b]++;
		out[j] = (LfT)((uint64_t)v << LF_SHIFT | (uint64_t)(a & 7));
	}
	if (i0 + PREP_PER_THREAD <= len) { /* i0 is a multiple of 16 elements: whole 16-byte stores */
		uint4 *o4 = (uint4*)(lf + i0);
		const uint4 *s4 = (const uint4*)out;
#pragma unroll
		for (int k = 0; k < (int)(PREP_PER_THREAD * sizeof(LfT) / 16); ++k) o4[k] = s4[k];
	} else {
		for (int j = 0; j < PREP_PER_THREAD; ++j) if (i0 + j < len) lf[i0 + j] = out[j];
	}
}

/* ------------------------------------------------------------------ */
/* the batch in walk order                                              */
/* ------------------------------------------------------------------ */

/* Fine marks: the sentinel rows (fine node = row) and every fine_len-th row (fine_len a power of two).  They cut the
 * chains into short pieces of random length. */
struct Fine {
	int64_t n_fine, n_seq, m0;
	int fshift;
	__host__ __device__ int64_t row(int64_t f) const { return f < n_seq ? f : (m0 + (f - n_seq)) << fshift; }
	__host__ __device__ int64_t of_row(int64_t r) const { return r < n_seq ? r : n_seq + ((r >> fshift) - m0); }
	__host__ __device__ bool is_mark(int64_t r) const { return r < n_seq || (r & ((1LL << fshift) - 1)) == 0; }
};

/* list-ranking node of a fine mark: x = #rows from the mark to the end of what has been linked so far,
 * y = next mark (low 32 bits, -1 = none) | last mark of the chain seen so far (high 32 bits) */
typedef longlong2 FNode;
__device__ __forceinline__ FNode fnode(int64_t dist, int32_t succ, int32_t term) { return make_longlong2(dist, (int64_t)((uint64_t)(uint32_t)succ | (uint64_t)(uint32_t)term << 32)); }
__device__ __forceinline__ int32_t fnode_succ(const FNode &n) { return (int32_t)(uint32_t)n.y; }
__device__ __forceinline__ int32_t fnode_term(const FNode &n) { return (int32_t)(uint32_t)((uint64_t)n.y >> 32); }

/* walk every piece once (LF_B only, no rank): its length and the piece that follows it */
template<typename LfT>
__global__ void k_fine_walk(Fine F, const LfT *__restrict__ lf, FNode *__restrict__ node, int32_t *__restrict__ piece_len, int64_t f_lo, int64_t f_hi)
{
	int64_t f = f_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= f_hi) return;
	int64_t kb = F.row(f), n = 0, nx = -1;
	for (;;) {
		const uint64_t x = __ldg(lf + kb);
		++n;
		if ((x & 7) == 0) break; /* first symbol of the sequence: the chain ends here, fm-index.c:170 */
		kb = (int64_t)(x >> LF_SHIFT);
		if (F.is_mark(kb)) { nx = F.of_row(kb); break; }
	}
	node[f] = fnode(n, (int32_t)nx, (int32_t)f);
	piece_len[f] = (int32_t)n;
}

/* multi-device: the pieces were walked by their owners and the nodes exchanged; every device needs all piece lengths */
__global__ void k_piece_len(int64_t n, const FNode *__restrict__ node, int32_t *__restrict__ piece_len)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f < n) piece_len[f] = (int32_t)node[f].x;
}

/* Wyllie pointer jumping: after ceil(log2 n) rounds x = #rows from fine mark f to the start of its sequence.  One 16-byte
 * gather per mark and round. */
__global__ void k_list_rank(int64_t n, const FNode *__restrict__ in, FNode *__restrict__ out)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= n) return;
	FNode a = in[f];
	const int32_t s = fnode_succ(a);
	if (s >= 0) {
		const FNode b = in[s];
		a = fnode(a.x + b.x, fnode_succ(b), fnode_term(b));
	}
	out[f] = a;
}

/* the sentinel fine mark p starts chain p: its length, and which chain the terminal mark belongs to */
__global__ void k_chain_len(int64_t n_seq, const FNode *__restrict__ node, int64_t *__restrict__ chain_len, int64_t *__restrict__ chain_of)
{
	int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p < n_seq) { const FNode a = node[p]; chain_len[p] = a.x; chain_of[fnode_term(a)] = p; }
}

/* second walk of the pieces: write the batch in walk order.  Chain p occupies positions [chain_base[p], +chain_len[p]);
 * the piece of fine mark f starts node[f].x positions before the end of its chain.  Every thread writes one contiguous
 * stretch; it collects 8 symbols / 16 bytes of rows in registers and stores them as one aligned word (the unaligned
 * head and tail go out element by element: neighbouring pieces own the rest of those words). */
template<typename RowT> struct RowVec;
template<> struct RowVec<uint32_t> {
	static const int N = 4;
	uint32_t r[4];
	__device__ __forceinline__ void push(uint32_t v) { r[0] = r[1]; r[1] = r[2]; r[2] = r[3]; r[3] = v; }
	__device__ __forceinline__ void store(uint32_t *dst) const { *(uint4*)dst = make_uint4(r[0], r[1], r[2], r[3]); }
	__device__ __forceinline__ uint32_t last(int k) const { return k == 0 ? r[3] : k == 1 ? r[2] : k == 2 ? r[1] : r[0]; } /* k-th newest */
};
template<> struct RowVec<int64_t> {
	static const int N = 2;
	int64_t r[2];
	__device__ __forceinline__ void push(int64_t v) { r[0] = r[1]; r[1] = v; }
	__device__ __forceinline__ void store(int64_t *dst) const { *(longlong2*)dst = make_longlong2(r[0], r[1]); }
	__device__ __forceinline__ int64_t last(int k) const { return k == 0 ? r[1] : r[0]; }
};

template<typename LfT, typename RowT>
__global__ void k_write_walk(Fine F, int64_t len, const LfT *__restrict__ lf, const FNode *__restrict__ node,
                             const int64_t *__restrict__ chain_of, const int64_t *__restrict__ chain_base, const int64_t *__restrict__ chain_len,
                             const int32_t *__restrict__ piece_len, int64_t p_lo, int64_t p_hi, RowT *__restrict__ wrow, uint8_t *__restrict__ wsym)
{
	int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= F.n_fine) return;
	const FNode me = node[f];
	const int64_t p = chain_of[fnode_term(me)];
	if (p < 0) return; /* not on a chain that starts at a sentinel: not a valid BWT, reported by the caller */
	int64_t pos = chain_base[p] + chain_len[p] - me.x, kb = F.row(f);
	if (pos < 0 || pos + me.x > len) return; /* cannot happen for a valid BWT */
	if (pos >= p_hi || pos + piece_len[f] <= p_lo) return; /* another device walks these positions */
	uint64_t sw = 0;
	int ns = 0, nr = 0; /* symbols / rows collected */
	RowVec<RowT> rv;
	rv.r[0] = rv.r[1] = 0; if (RowVec<RowT>::N == 4) { rv.r[RowVec<RowT>::N - 2] = 0; rv.r[RowVec<RowT>::N - 1] = 0; }
	for (;;) {
		const uint64_t x = __ldg(lf + kb);
		const uint64_t c = x & 7;
		if (ns == 0 && (pos & 7) != 0) wsym[pos] = (uint8_t)c; /* unaligned head */
		else {
			sw = sw >> 8 | c << 56;
			if (++ns == 8) { *(uint64_t*)(wsym + pos - 7) = sw; ns = 0; }
		}
		if (nr == 0 && (pos & (RowVec<RowT>::N - 1)) != 0) wrow[pos] = (RowT)kb;
		else {
			rv.push((RowT)kb);
			if (++nr == RowVec<RowT>::N) { rv.store(wrow + pos - (RowVec<RowT>::N - 1)); nr = 0; }
		}
		++pos;
		if (c == 0) break;
		kb = (int64_t)(x >> LF_SHIFT);
		if (F.is_mark(kb)) break;
	}
	/* tails: the ns newest symbols are the top bytes of sw, the nr newest rows the last entries of rv */
	for (int k = 0; k < ns; ++k) wsym[pos - 1 - k] = (uint8_t)(sw >> (56 - 8 * k));
	for (int k = 0; k < nr; ++k) wrow[pos - 1 - k] = rv.last(k);
}

/* ---- one chase instead of two (large batches: the LF table is far bigger than L2, every chase step is a DRAM access) ----
 * k_fine_walk_buf is k_fine_walk that also KEEPS what it meets: the first `cap` (row, symbol) pairs of every piece go to a
 * piece-major buffer (cap entries per piece, written as whole 16-byte / 8-byte words), and the row of step `cap` is
 * remembered for the few pieces that are longer.  k_copy_walk then moves the buffered pairs to their walk-order positions
 * -- streaming reads -- and chases only the tails beyond `cap` (cap = 2 x the mean piece length: 13.5% of the rows). */
template<typename LfT, typename RowT>
__global__ void k_fine_walk_buf(Fine F, const LfT *__restrict__ lf, FNode *__restrict__ node, int32_t *__restrict__ piece_len,
                                RowT *__restrict__ rowbuf, uint8_t *__restrict__ symbuf, int cap, RowT *__restrict__ tail_row)
{
	const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= F.n_fine) return;
	int64_t kb = F.row(f), n = 0, nx = -1;
	RowT *rb = rowbuf + f * cap;
	uint8_t *sb = symbuf + f * cap;
	RowVec<RowT> rv;
	uint64_t sw = 0;
	for (;;) {
		const uint64_t x = __ldg(lf + kb);
		const uint64_t c = x & 7;
		if (n < cap) {
			rv.push((RowT)kb);
			sw = sw >> 8 | c << 56;
			if (((n + 1) & (RowVec<RowT>::N - 1)) == 0) rv.store(rb + n + 1 - RowVec<RowT>::N);
			if (((n + 1) & 7) == 0) *(uint64_t*)(sb + n - 7) = sw;
		}
		++n;
		if (c == 0) break; /* first symbol of the sequence: the chain ends here, fm-index.c:170 */
		kb = (int64_t)(x >> LF_SHIFT);
		if (n == cap) tail_row[f] = (RowT)kb; /* where the part that is not buffered starts */
		if (F.is_mark(kb)) { nx = F.of_row(kb); break; }
	}
	const int64_t nbuf = n < cap ? n : cap; /* partial last words */
	for (int k = 0; k < (int)(nbuf & (RowVec<RowT>::N - 1)); ++k) rb[nbuf - 1 - k] = rv.last(k);
	for (int k = 0; k < (int)(nbuf & 7); ++k) sb[nbuf - 1 - k] = (uint8_t)(sw >> (56 - 8 * k));
	node[f] = fnode(n, (int32_t)nx, (int32_t)f);
	piece_len[f] = (int32_t)n;
}

template<typename LfT, typename RowT>
__global__ void k_copy_walk(Fine F, int64_t len, const LfT *__restrict__ lf, const FNode *__restrict__ node,
                            const int64_t *__restrict__ chain_of, const int64_t *__restrict__ chain_base, const int64_t *__restrict__ chain_len,
                            const int32_t *__restrict__ piece_len, const RowT *__restrict__ rowbuf, const uint8_t *__restrict__ symbuf, int cap,
                            const RowT *__restrict__ tail_row, RowT *__restrict__ wrow, uint8_t *__restrict__ wsym)
{
	const int64_t f = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (f >= F.n_fine) return;
	const FNode me = node[f];
	const int64_t p = chain_of[fnode_term(me)];
	if (p < 0) return; /* not on a chain that starts at a sentinel: not a valid BWT, reported by the caller */
	int64_t pos = chain_base[p] + chain_len[p] - me.x;
	const int plen = piece_len[f];
	if (pos < 0 || pos + me.x > len) return; /* cannot happen for a valid BWT */
	const RowT *rb = rowbuf + f * cap;
	const uint8_t *sb = symbuf + f * cap;
	uint64_t sw = 0;
	int ns = 0, nr = 0; /* symbols / rows collected */
	RowVec<RowT> rv;
	rv.r[0] = rv.r[1] = 0; if (RowVec<RowT>::N == 4) { rv.r[RowVec<RowT>::N - 2] = 0; rv.r[RowVec<RowT>::N - 1] = 0; }
	int64_t kb = 0;
	for (int k = 0; k < plen; ++k) {
		uint64_t c;
		if (k < cap) { kb = (int64_t)rb[k]; c = sb[k]; }
		else {
			if (k == cap) kb = (int64_t)tail_row[f];
			const uint64_t x = __ldg(lf + kb);
			c = x & 7;
			if (ns == 0 && (pos & 7) != 0) wsym[pos] = (uint8_t)c;
			else { sw = sw >> 8 | c << 56; if (++ns == 8) { *(uint64_t*)(wsym + pos - 7) = sw; ns = 0; } }
			if (nr == 0 && (pos & (RowVec<RowT>::N - 1)) != 0) wrow[pos] = (RowT)kb;
			else { rv.push((RowT)kb); if (++nr == RowVec<RowT>::N) { rv.store(wrow + pos - (RowVec<RowT>::N - 1)); nr = 0; } }
			++pos;
			kb = (int64_t)(x >> LF_SHIFT);
			continue;
		}
		if (ns == 0 && (pos & 7) != 0) wsym[pos] = (uint8_t)c; /* unaligned head */
		else { sw = sw >> 8 | c << 56; if (++ns == 8) { *(uint64_t*)(wsym + pos - 7) = sw; ns = 0; } }
		if (nr == 0 && (pos & (RowVec<RowT>::N - 1)) != 0) wrow[pos] = (RowT)kb;
		else { rv.push((RowT)kb); if (++nr == RowVec<RowT>::N) { rv.store(wrow + pos - (RowVec<RowT>::N - 1)); nr = 0; } }
		++pos;
	}
	for (int k = 0; k < ns; ++k) wsym[pos - 1 - k] = (uint8_t)(sw >> (56 - 8 * k));
	for (int k = 0; k < nr; ++k) wrow[pos - 1 - k] = rv.last(k);
}

/* ------------------------------------------------------------------ */
/* sliced LF walk over A                                                */
/* ------------------------------------------------------------------ */

struct Slices {
	int64_t n_seg, len, seg_len;  /* slice s = walk-order positions [s * seg_len, min(len, (s + 1) * seg_len)); seg_len % 8 == 0 */
	int64_t own_lo, own_hi;       /* slices this device is responsible for */
	int64_t walk_lo;              /* first slice this device walks (own_lo minus the halo) */
	int64_t *d;                   /* #rows at the start of the slice that are still unresolved */
	int64_t *arr_lo, *arr_hi;     /* bracket with which the walk arrived at the first row of the next slice */
	__host__ __device__ int64_t slice_len(int64_t s) const { int64_t r = len - s * seg_len; return r < seg_len ? r : seg_len; }
};

/* kseq[p]: interleave position of walk-order position p, or, while unresolved, the low end of its bracket plus flags */
#define KS_UNRES  (1LL << 62)
#define KS_NARROW (1LL << 61)    /* k_walk_pair only: the row's transfer mask is in wmask[] (KS_TIGHT) */

/* The walk kernels are written for a "ranker" WG: Grp<8> (RLE cells, 8 lanes per walk) or BmRank (bitmap cells,
 * one thread per walk). */

/* ---- RLO / RCLO collections (build -s / -r; mr_insert_multi_aux, mrope.c:226-275) ----
 * In a sorted collection the sentinel of a new string does not go after all old ones: rows whose suffix runs to the end
 * of the string ("S_t $") are ordered by the symbols that PRECEDE them, '$' (start of string) first, then A,C,G,T (RLO)
 * or T,G,C,A (RCLO), then N.  Walking a new string from its sentinel, [l,u) = rows of A with exactly the same suffix
 * (the reference's triple64_t interval, mrope.c:216-220); l and u both advance by LF; the row goes to
 *     ka_t = l_t + less_t + (ka_{t+1} - l_{t+1}),   less_t = #rows of A[l_t,u_t) preceded by a smaller symbol
 * (mrope.c:247-265: the new run is placed before the existing equal symbols), i.e. ka_t = l_t + sum_{s>=t} less_s; once
 * the interval is empty the position is exact and the ordinary walk takes over.  These "heads" are short (log_4 of the
 * number of strings for reads) and are resolved by one walk per new string BEFORE the sliced walk, which then skips
 * rows flagged KS_HEAD and restarts from the exact value flagged KS_SEED. */
#define KS_HEAD (1LL << 60)
#define KS_SEED (1LL << 59)

__device__ __forceinline__ int so_rank(int so, int c) { return (so == 2 && c >= 1 && c <= 4) ? 5 - c : c; } /* rope_comp6, mrope.c:224 */

template<class WG>
__global__ void __launch_bounds__(TPB) k_so_heads(DevIndex A, int so, int64_t n_seq, const int64_t *__restrict__ chain_base, const int64_t *__restrict__ chain_len,
                                                   const uint8_t *__restrict__ wsym, int64_t *__restrict__ kseq)
{
	const int gl = WG::lane();
	const int64_t p = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / WG::G;
	if (p >= n_seq) return; /* group-uniform */
	const int64_t pos0 = chain_base[p], L = chain_len[p];
	int64_t l = 0, u = A.acc[1], P = 0, t = 0;
	bool ended = false;
	while (t < L && l < u) {
		const int c = (int)wsym[pos0 + t];
		int64_t tl[RB3B_ASIZE], tu[RB3B_ASIZE], less = 0;
#pragma unroll
		for (int b = 0; b < RB3B_ASIZE; ++b) { tl[b] = WG::rank(A, l, b); tu[b] = WG::rank(A, u, b); }
#pragma unroll
		for (int b = 0; b < RB3B_ASIZE; ++b) if (so_rank(so, b) < so_rank(so, c)) less += tu[b] - tl[b];
		if (gl == 0) kseq[pos0 + t] = l - P; /* + the final P = l + sum_{s >= t} less_s */
		P += less;
		++t;
		if (c == 0) { ended = true; break; }
		int64_t nl = 0, nu = 0;
#pragma unroll
		for (int b = 0; b < RB3B_ASIZE; ++b) if (b == c) { nl = tl[b]; nu = tu[b]; }
		l = A.acc[c] + nl; u = A.acc[c] + nu;
	}
	if (gl == 0) {
		for (int64_t s = 0; s < t; ++s) kseq[pos0 + s] = (kseq[pos0 + s] + P) | KS_HEAD;
		if (!ended && t < L) kseq[pos0 + t] = l | KS_SEED;
	}
}

/* round 1: every slice is walked from its first position to its last */
template<class WG, bool SO>
__global__ void __launch_bounds__(TPB) k_walk_first(DevIndex A, Slices S, const uint8_t *__restrict__ wsym, int64_t *__restrict__ kseq, int64_t *next_seg)
{
	const int gl = WG::lane(), gbase = WG::base();
	const unsigned gmask = WG::mask();
	for (;;) {
		int64_t s = 0;
		if (gl == 0) s = S.walk_lo + (int64_t)atomicAdd((unsigned long long*)next_seg, 1ULL);
		s = __shfl_sync(gmask, s, gbase);
		if (s >= S.own_hi) break;
		const int64_t p0 = s * S.seg_len, n = S.slice_len(s);
		const int c0 = s == 0 ? 0 : (int)wsym[p0 - 1]; /* the symbol that led here = first symbol of this row's suffix */
		int64_t lo, hi, d = 0;
		if (c0 == 0) lo = hi = SO ? 0 : A.acc[1]; /* a sentinel row: new sentinels sort after all old ones, fm-index.c:164 (SO: flagged rows follow) */
		else { lo = A.acc[c0]; hi = A.acc[c0 + 1]; }
		uint64_t w = __ldg((const uint64_t*)(wsym + p0)); /* wsym is padded: whole words can always be read */
		for (int64_t j = 0; j < n; j += 8) {
			const uint64_t wn = j + 8 < n ? __ldg((const uint64_t*)(wsym + p0 + j + 8)) : 0; /* the symbols do not depend on A: fetch ahead */
			int64_t vb[8];
			if (SO) { /* rows resolved or seeded by k_so_heads */
#pragma unroll
				for (int jj = 0; jj < 8; ++jj) vb[jj] = j + jj < n ? kseq[p0 + j + jj] : 0;
			}
#pragma unroll
			for (int jj = 0; jj < 8; ++jj) {
				const int c = (int)(w >> (8 * jj)) & 7;
				const bool live = j + jj < n;
				bool head = false;
				if (SO) {
					const int64_t f = vb[jj];
					if (f & KS_HEAD) { head = true; vb[jj] = f & ~KS_HEAD; lo = hi = 0; }
					else if (f & KS_SEED) lo = hi = f & (int64_t)RB3B_M42;
				}
				if (head) continue;
				if (lo == hi) vb[jj] = lo;
				else { vb[jj] = lo | KS_UNRES; d += live; }
				if (live) {
					if (c == 0) lo = hi = SO ? 0 : A.acc[1]; /* first symbol of the sequence (fm-index.c:170): the next position is a sentinel row */
					else {
						int64_t r1, r2;
						/* the walks that currently run together take the two-position path only while one of them still
						 * carries a bracket; either path is correct for an exact walk, so this is purely a cost choice
						 * (a private per-thread branch was measured slower: the two paths serialise) */
						if (WG::ALWAYS2 || __any_sync(gmask, lo != hi)) WG::rank2(A, lo, hi, c, r1, r2); /* vote over the whole group only: lo, hi are group-uniform */
						else r1 = r2 = WG::rank(A, lo, c);
						lo = A.acc[c] + r1; hi = A.acc[c] + r2;
					}
				}
			}
			if (gl == 0) {
				if (j + 8 <= n) { /* full 64-B line of results */
					longlong2 *o = (longlong2*)(kseq + p0 + j);
					o[0] = make_longlong2(vb[0], vb[1]); o[1] = make_longlong2(vb[2], vb[3]);
					o[2] = make_longlong2(vb[4], vb[5]); o[3] = make_longlong2(vb[6], vb[7]);
				} else {
#pragma unroll
					for (int jj = 0; jj < 8; ++jj) if (j + jj < n) kseq[p0 + j + jj] = vb[jj];
				}
			}
			w = wn;
		}
		if (gl == 0) { S.d[s] = d; S.arr_lo[s] = lo; S.arr_hi[s] = hi; }
	}
}

/* after round 1: slices whose predecessor arrived with an exact value and that have unresolved rows */
/* nc_of != 0: also number the slices that never collapsed and hand an inexact bracket on (nc_of[s] = 0, 1, ...; -1 for all
 * others): the ones that get a transfer table (k_fix_tables) */
__global__ void k_collect_first(Slices S, int64_t *__restrict__ wl_seg, int64_t *__restrict__ wl_val, unsigned long long *wl_n,
                                int32_t *__restrict__ nc_of = 0, int32_t *__restrict__ nc_list = 0, unsigned long long *nc_n = 0)
{
	int64_t s = S.walk_lo + (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (s >= S.own_hi) return;
	if (nc_of) {
		int32_t id = -1;
		if (S.d[s] == S.slice_len(s) && S.arr_lo[s] != S.arr_hi[s]) { id = (int32_t)atomicAdd(nc_n, 1ULL); nc_list[id] = (int32_t)(s - S.walk_lo); }
		nc_of[s - S.walk_lo] = id;
	}
	if (s + 1 >= S.own_hi) return;
	if (S.arr_lo[s] == S.arr_hi[s] && S.d[s + 1] > 0) {
		unsigned long long o = atomicAdd(wl_n, 1ULL);
		wl_seg[o] = s + 1; wl_val[o] = S.arr_lo[s];
	}
}

/* ---- bitmap cells: lane-pair walk that leaves a transfer function behind ----
 * Same walk as k_walk_first<BmPair>, plus:
 *  - WARM-UP: the walk of slice s starts `warm` positions before the slice (rows of the previous slice, results
 *    discarded), so that its bracket is already a few dozen rows wide when the first own row is reached;
 *  - MASKS: for an unresolved row whose bracket is at most 127 wide the walk stores wmask[p] = the 128 bits of plane c
 *    (c = the row's symbol) from position lo on.  With v = lo + x the exact position of the row (0 <= x <= hi - lo),
 *    the next row's exact position is lo' + popc(mask below x): the fix-up never touches the index again.
 * The even lane ranks lo, the odd lane hi; each holds the plane of its own cell, the odd lane's travels by shuffle. */
#define KS_TIGHT KS_NARROW   /* the row's mask is in wmask[] */
#define TIGHT_WIDTH 127
#define KS_WIDTH_SHIFT 43    /* tight rows: the width of the bracket (7 bits) above the 42-bit low end */

template<bool SO>
__global__ void __launch_bounds__(32) k_walk_pair(DevIndex A, Slices S, const uint8_t *__restrict__ wsym, int64_t *__restrict__ kseq, uint4 *__restrict__ wmask,
                                                  int warm, int64_t *next_seg)
{
	const int odd = threadIdx.x & 1, gbase = threadIdx.x & 30;
	const unsigned gmask = 3u << gbase;
	for (;;) {
		int64_t s = 0;
		if (!odd) s = S.walk_lo + (int64_t)atomicAdd((unsigned long long*)next_seg, 1ULL);
		s = __shfl_sync(gmask, s, gbase);
		if (s >= S.own_hi) break;
		const int64_t p0 = s * S.seg_len, n = S.slice_len(s);
		const int64_t q0 = (SO || s == 0) ? p0 : p0 - warm; /* warm % 8 == 0, warm <= seg_len */
		const int c0 = q0 == 0 ? 0 : (int)wsym[q0 - 1];
		int64_t lo, hi, d = 0;
		if (c0 == 0) lo = hi = SO ? 0 : A.acc[1];
		else { lo = A.acc[c0]; hi = A.acc[c0 + 1]; }
		uint64_t w = __ldg((const uint64_t*)(wsym + q0));
		for (int64_t j = q0 - p0; j < n; j += 8) {
			const uint64_t wn = j + 8 < n ? __ldg((const uint64_t*)(wsym + p0 + j + 8)) : 0;
			const bool own = j >= 0;
			int64_t vb[8];
			if (SO) {
#pragma unroll
				for (int jj = 0; jj < 8; ++jj) vb[jj] = j + jj < n ? kseq[p0 + j + jj] : 0;
			}
#pragma unroll
			for (int jj = 0; jj < 8; ++jj) {
				const int c = (int)(w >> (8 * jj)) & 7;
				const bool live = j + jj < n;
				bool head = false;
				if (SO) {
					const int64_t f = vb[jj];
					if (f & KS_HEAD) { head = true; vb[jj] = f & ~KS_HEAD; lo = hi = 0; }
					else if (f & KS_SEED) lo = hi = f & (int64_t)RB3B_M42;
				}
				if (head) continue;
				const bool unres = lo != hi;
				const bool tight = unres && wmask != 0 && hi - lo <= TIGHT_WIDTH;
				if (!unres) vb[jj] = lo;
				else { vb[jj] = lo | KS_UNRES | (tight ? KS_TIGHT | (hi - lo) << KS_WIDTH_SHIFT : 0); d += live && own; }
				if (live) {
					if (c == 0) lo = hi = SO ? 0 : A.acc[1];
					else {
						const int64_t k = odd ? hi : lo, kk = k < A.n ? k : A.n - 1;
						const int h = c >= 3, cc = c - 3 * h;
						const uint4 *half = A.cells + ((kk >> RB3B_BM_SHIFT) * 8 + 4 * h);
						const uint4 cq = __ldg(half), pq = __ldg(half + 1 + cc);
						const uint64_t l64 = (uint64_t)cq.x | (uint64_t)cq.y << 32, h64 = (uint64_t)cq.z | (uint64_t)cq.w << 32;
						const uint64_t cnt = (cc == 0 ? l64 : cc == 1 ? (l64 >> 42 | h64 << 22) : h64 >> 20) & RB3B_M42;
						const uint64_t b0 = (uint64_t)pq.x | (uint64_t)pq.y << 32, b1 = (uint64_t)pq.z | (uint64_t)pq.w << 32;
						const uint32_t o = (uint32_t)kk & 127u;
						const uint64_t m0 = o >= 64u ? ~0ULL : (1ULL << o) - 1ULL, m1 = o <= 64u ? 0ULL : (1ULL << (o - 64u)) - 1ULL;
						int64_t r = (int64_t)(cnt + (uint32_t)(__popcll(b0 & m0) + __popcll(b1 & m1)));
						r = k < A.n ? r : A.tot[c];
						const int64_t other = __shfl_xor_sync(gmask, r, 1);
						if (tight && own) { /* pair-uniform: both lanes carry the same lo, hi */
							const uint32_t hx = __shfl_xor_sync(gmask, pq.x, 1), hy = __shfl_xor_sync(gmask, pq.y, 1);
							const uint32_t hz = __shfl_xor_sync(gmask, pq.z, 1), hw = __shfl_xor_sync(gmask, pq.w, 1);
							const int64_t jo = __shfl_xor_sync(gmask, kk >> RB3B_BM_SHIFT, 1);
							if (!odd) {
								const bool nxt = jo != (kk >> RB3B_BM_SHIFT); /* hi lies in the next cell */
								uint32_t ww[9] = { pq.x, pq.y, pq.z, pq.w, nxt ? hx : 0u, nxt ? hy : 0u, nxt ? hz : 0u, nxt ? hw : 0u, 0u };
								const uint32_t q = o >> 5, rr = o & 31u;
								if (q & 1u) {
#pragma unroll
									for (int i = 0; i < 8; ++i) ww[i] = ww[i + 1];
								}
								if (q & 2u) {
#pragma unroll
									for (int i = 0; i < 7; ++i) ww[i] = ww[i + 2];
								}
								wmask[p0 + j + jj] = make_uint4(__funnelshift_r(ww[0], ww[1], rr), __funnelshift_r(ww[1], ww[2], rr),
								                                __funnelshift_r(ww[2], ww[3], rr), __funnelshift_r(ww[3], ww[4], rr));
							}
						}
						lo = A.acc[c] + (odd ? other : r); hi = A.acc[c] + (odd ? r : other);
					}
				}
			}
			if (!odd && own) {
				if (j + 8 <= n) {
					longlong2 *o2 = (longlong2*)(kseq + p0 + j);
					o2[0] = make_longlong2(vb[0], vb[1]); o2[1] = make_longlong2(vb[2], vb[3]);
					o2[2] = make_longlong2(vb[4], vb[5]); o2[3] = make_longlong2(vb[6], vb[7]);
				} else {
#pragma unroll
					for (int jj = 0; jj < 8; ++jj) if (j + jj < n) kseq[p0 + j + jj] = vb[jj];
				}
			}
			w = wn;
		}
		if (!odd) { S.d[s] = d; S.arr_lo[s] = lo; S.arr_hi[s] = hi; }
	}
}

/* mask of the low min(max(t, 0), 32) bits */
__device__ __forceinline__ uint32_t below32(int t)
{
	uint32_t r;
	const uint32_t tt = (uint32_t)max(t, 0);
	asm("bmsk.clamp.b32 %0, %1, %2;" : "=r"(r) : "r"(0u), "r"(tt));
	return r;
}

/* #set bits of the 128-bit mask m among bit positions [0, x), 0 <= x <= 128 */
__device__ __forceinline__ uint32_t mask_rank(const uint4 m, uint32_t x)
{
	return __popc(m.x & below32((int)x)) + __popc(m.y & below32((int)x - 32)) + __popc(m.z & below32((int)x - 64)) + __popc(m.w & below32((int)x - 96));
}

__device__ __forceinline__ void cp_async16(void *smem, const void *gmem)
{ asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory"); }
__device__ __forceinline__ void cp_async8(void *smem, const void *gmem)
{ asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory"); }
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template<int N> __device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" :: "n"(N) : "memory"); }

#define FIX_TPB 32      /* default threads per CTA of the fix-up ("fix_tpb": 16 or 32) */
#define FIX_STAGES 4    /* default depth of its cp.async ring ("fix_stages": 2 or 4) */
/* shared-memory ring of a fix-up CTA: per thread and stage eight consecutive rows of its slice (thread-interleaved so that
 * the threads of a warp hit different banks).  T threads x ST stages x 208 bytes: the ring, not the registers, bounds how
 * many of these one-warp CTAs an SM holds (32 x 4: 8 CTAs; 16 x 2: 32 CTAs) */
template<int T, int ST>
struct FixRing {
	uint4 m[ST][8][T];        /* transfer masks */
	longlong2 ks[ST][4][T];   /* kseq words (low end of the bracket + flags) */
	uint64_t sym[ST][T];
	int64_t lo_next[ST][T];   /* kseq word of the row after the eight */
};

/* fix-up for bitmap cells: one THREAD per listed slice.  The rows of a slice are consecutive in kseq / wsym / wmask / wrow,
 * so every thread streams its own rows in blocks of eight through a private cp.async ring in shared memory (FIX_STAGES
 * blocks in flight: nothing it fetches depends on the value it carries) and the dependent chain per row is
 * x' = popc(mask below x), 32-bit, no access to the index.  Rows without a mask (bracket wider than 127: huge indexes, or
 * no warm-up) take the general step with one random cell access.  A slice that never collapsed hands its exact arrival
 * straight to the next slice in the same thread, so one launch resolves every cascade.
 * stats: [0] rows, [1] rows that took the general step, [2] longest chain.
 * (Tried and dropped: scattering the exact rows to ka[wrow[p]] from inside the walk and this kernel instead of a separate
 * pass -- the random stores compete with the walk's own dependent accesses, 0.31 -> 0.67 ms for the walk; and G = 4..32 lanes
 * per slice with the mask travelling by shuffle -- 0.53-0.61 ms against 0.27 ms for one thread per slice.) */
/* the general step of a row without a mask: one random cell access (kept out of line: the chain loop stays short) */
__device__ __noinline__ int64_t fix_wide_step(const uint4 *cells, int64_t n, int64_t tot_c, int64_t acc_c, int64_t pos, int c, int64_t lo_n)
{
	DevIndex A;
	A.cells = cells; A.n = n; A.tot[c] = tot_c;
	return acc_c + BmRank::rank(A, pos, c) - lo_n;
}

template<int T, int ST>
__global__ void __launch_bounds__(T) k_fix_chain(DevIndex A, Slices S, const uint8_t *__restrict__ wsym, int64_t *__restrict__ kseq, const uint4 *__restrict__ wmask,
                                                  int64_t n_items, const int64_t *__restrict__ wl_seg, const int64_t *__restrict__ wl_val, unsigned long long *stats,
                                                  int cascade, const unsigned long long *__restrict__ n_items_dev)
{
	__shared__ FixRing<T, ST> R;
	const int tx = threadIdx.x;
	const int64_t it = (int64_t)blockIdx.x * blockDim.x + tx;
	if (n_items_dev) n_items = (int64_t)*n_items_dev; /* the list was counted on the device: the grid covers an upper bound */
	unsigned long long n_rows = 0, n_wide = 0;
	if (it < n_items) {
		int64_t t = wl_seg[it], v = wl_val[it];
		for (;;) {
			const int64_t d = S.d[t], len = S.slice_len(t), base = t * S.seg_len;
			/* a slice that collapsed exactly on its last row already had an exact arrival in round 1: k_collect_first
			 * queued its successor as an item of its own, so it must not be entered from here as well */
			const bool was_exact = S.arr_lo[t] == S.arr_hi[t];
			const int64_t arr = S.arr_lo[t];
			const int nb = (int)((d + 7) >> 3), nd = (int)d, nl = (int)len;
			bool ended = false;
			n_rows += (unsigned long long)d;
			/* blocks are fetched whole (8 rows, aligned): the buffers are padded past the last slice */
			const int64_t *ksp = kseq + base;
			const uint8_t *syp = wsym + base;
			const uint4 *mkp = wmask ? wmask + base : 0;
#define FIX_ISSUE(b) do { if ((b) < nb) { const int st_ = (b) & (ST - 1); const int p_ = 8 * (b); \
				for (int i_ = 0; i_ < 4; ++i_) cp_async16(&R.ks[st_][i_][tx], ksp + p_ + 2 * i_); \
				cp_async8(&R.sym[st_][tx], syp + p_); \
				if (mkp) for (int i_ = 0; i_ < 8; ++i_) cp_async16(&R.m[st_][i_][tx], mkp + p_ + i_); \
				if (p_ + 8 < nl) cp_async8(&R.lo_next[st_][tx], ksp + p_ + 8); } \
				cp_async_commit(); } while (0)
			for (int b = 0; b < ST - 1; ++b) FIX_ISSUE(b);
			int64_t x = v - (kseq[base] & (int64_t)RB3B_M42); /* the exact value relative to the current row's low end */
			for (int b = 0; b < nb && !ended; ++b) {
				FIX_ISSUE(b + ST - 1);
				cp_async_wait<ST - 1>();
				const int st = b & (ST - 1);
				const int i0 = 8 * b;
				int64_t ksv[8];
#pragma unroll
				for (int i = 0; i < 4; ++i) { const longlong2 q = R.ks[st][i][tx]; ksv[2 * i] = q.x; ksv[2 * i + 1] = q.y; }
				const uint64_t sym = R.sym[st][tx] & 0x0707070707070707ULL;
				/* rows of this block still to resolve; a sentinel symbol among them ends the chain (the next position is a
				 * sentinel row, exact by itself): its row is the last one written */
				int du = nd - i0 < 8 ? nd - i0 : 8;
				const uint64_t z = (sym - 0x0101010101010101ULL) & ~sym & 0x8080808080808080ULL; /* lowest flagged byte = first zero byte */
				const int zi = z ? (__ffsll((long long)z) - 1) >> 3 : 8;
				if (zi < du) { du = zi + 1; ended = true; }
				const int n_upd = ended ? du - 1 : du; /* rows after which x moves on */
				uint4 mk[8]; /* all eight masks up front: the shared-memory latency stays off the chain */
#pragma unroll
				for (int u = 0; u < 8; ++u) mk[u] = R.m[st][u][tx];
				const int64_t all_tight = ksv[0] & ksv[1] & ksv[2] & ksv[3] & ksv[4] & ksv[5] & ksv[6] & ksv[7] & KS_TIGHT;
				if (n_upd == 8 && all_tight) {
					/* the common block -- eight rows, all with a mask, no sentinel: straight-line code.  Only the eight
					 * popc steps depend on each other; the positions are added up beside them */
					uint32_t xs[9];
					xs[0] = (uint32_t)x;
					const int64_t any_w = ksv[0] | ksv[1] | ksv[2] | ksv[3] | ksv[4] | ksv[5] | ksv[6] | ksv[7];
					if (((any_w >> (KS_WIDTH_SHIFT + 5)) & 3) == 0) { /* every bracket narrower than 32: x stays below 32, one word of the mask decides */
#pragma unroll
						for (int u = 0; u < 8; ++u) xs[u + 1] = __popc(mk[u].x & below32((int)xs[u]));
					} else {
#pragma unroll
						for (int u = 0; u < 8; ++u) xs[u + 1] = mask_rank(mk[u], xs[u]);
					}
#pragma unroll
					for (int u = 0; u < 8; ++u) ksv[u] = (ksv[u] & (int64_t)RB3B_M42) + (int64_t)xs[u];
					x = (int64_t)xs[8];
				} else
#pragma unroll
				for (int u = 0; u < 8; ++u) {
					if (u < du) {
						const int64_t ks = ksv[u];
						ksv[u] = (ks & (int64_t)RB3B_M42) + x;
						if (u < n_upd) {
							if (ks & KS_TIGHT) x = (int64_t)mask_rank(mk[u], (uint32_t)x);
							else {
								const int c = (int)(sym >> (8 * u)) & 7;
								const int64_t nx = u < 7 ? ksv[u + 1] : (i0 + 8 < nl ? R.lo_next[st][tx] : arr);
								x = fix_wide_step(A.cells, A.n, A.tot[c], A.acc[c], ksv[u], c, nx & (int64_t)RB3B_M42);
								++n_wide;
							}
						}
					}
				}
				if (i0 + 8 <= nl) {
					longlong2 *o2 = (longlong2*)(kseq + base + i0);
					o2[0] = make_longlong2(ksv[0], ksv[1]); o2[1] = make_longlong2(ksv[2], ksv[3]);
					o2[2] = make_longlong2(ksv[4], ksv[5]); o2[3] = make_longlong2(ksv[6], ksv[7]);
				} else {
#pragma unroll
					for (int u = 0; u < 8; ++u) if (i0 + u < nl) kseq[base + i0 + u] = ksv[u];
				}
			}
			cp_async_wait<0>(); /* the ring is restarted for the next slice */
#undef FIX_ISSUE
			v = arr + x; /* meaningful only when the slice never collapsed */
			const bool go_on = cascade && d == len && !ended && !was_exact && t + 1 < S.own_hi && S.d[t + 1] > 0; /* t + 1's only predecessor is t and t + 1 is on nobody's list */
			S.d[t] = 0;
			if (d == len && !ended) S.arr_lo[t] = S.arr_hi[t] = v;
			if (!go_on) break;
			++t;
		}
	}
	/* statistics (all T threads of the CTA arrive here together) */
	unsigned long long tot = n_rows, mx = n_rows;
	const unsigned lanes = T == 32 ? 0xffffffffu : (1u << T) - 1u;
	for (int o = T / 2; o > 0; o >>= 1) {
		n_wide += __shfl_xor_sync(lanes, n_wide, o);
		tot += __shfl_xor_sync(lanes, tot, o);
		const unsigned long long y = __shfl_xor_sync(lanes, mx, o);
		mx = y > mx ? y : mx;
	}
	if (tx == 0 && tot) { atomicAdd(stats, tot); atomicAdd(stats + 1, n_wide); atomicMax(stats + 2, mx); }
}

/* fix_tpb x fix_stages: 32 x 4 (eight one-warp CTAs per SM) or fewer threads / stages per CTA, i.e. more warps per scheduler
 * to cover the latency of the dependent chain */
static int launch_fix_chain(const DevIndex &dA, const Slices &S, const uint8_t *wsym, int64_t *kseq, const uint4 *wmask, int64_t n_items,
                            const int64_t *wl_seg, const int64_t *wl_val, unsigned long long *stats, int cascade, const unsigned long long *n_items_dev = 0)
{
	const int tpb = rb3b_get_param("fix_tpb", FIX_TPB) == 16 ? 16 : 32, st = rb3b_get_param("fix_stages", FIX_STAGES) == 2 ? 2 : 4;
	const unsigned grid = (unsigned)((n_items + tpb - 1) / tpb);
	if (tpb == 32 && st == 4) k_fix_chain<32, 4><<<grid, 32, 0, rb3b_stream>>>(dA, S, wsym, kseq, wmask, n_items, wl_seg, wl_val, stats, cascade, n_items_dev);
	else if (tpb == 32) k_fix_chain<32, 2><<<grid, 32, 0, rb3b_stream>>>(dA, S, wsym, kseq, wmask, n_items, wl_seg, wl_val, stats, cascade, n_items_dev);
	else if (st == 4) k_fix_chain<16, 4><<<grid, 16, 0, rb3b_stream>>>(dA, S, wsym, kseq, wmask, n_items, wl_seg, wl_val, stats, cascade, n_items_dev);
	else k_fix_chain<16, 2><<<grid, 16, 0, rb3b_stream>>>(dA, S, wsym, kseq, wmask, n_items, wl_seg, wl_val, stats, cascade, n_items_dev);
	CKK();
	return RB3B_OK;
}

/* ---- cascades without the chain: transfer tables of the slices that never collapsed ----
 * A slice whose bracket never collapsed maps the offset x_in with which the exact value enters it (0 <= x_in <= w0, the
 * width of its first bracket) to the offset with which it leaves: a monotone function of at most 128 arguments, independent
 * of everything outside the slice.  k_fix_tables evaluates it for ALL arguments at once -- one lane per argument, every lane
 * running the same popc(mask below x) chain over the slice's rows -- for every such slice in parallel.  k_fix_hops then
 * follows every cascade through the tables (one lookup per slice instead of one chain step per row) and lists every slice
 * with the exact value it starts from; k_fix_chain resolves all listed slices at once, no slice waiting for another.  The
 * longest dependent chain of the fix-up drops from the longest cascade (thousands of rows) to one slice. */
__global__ void __launch_bounds__(128) k_fix_tables(Slices S, const int64_t *__restrict__ kseq, const uint4 *__restrict__ wmask, int64_t n_nc, const int32_t *__restrict__ nc_list,
                                                      uint8_t *__restrict__ tab, uint8_t *__restrict__ tab_ok)
{
	/* one warp per (listed slice, block of 32 arguments); the rows' masks are fetched 32 rows at a time (one row per lane,
	 * coalesced) and handed round by shuffle, so the chain never waits for memory */
	const int64_t W = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
	const int lane = threadIdx.x & 31, cb = (int)(W & 3);
	if ((W >> 2) >= n_nc) return; /* warp-uniform */
	const int64_t id = W >> 2, s = S.walk_lo + nc_list[id];
	const int64_t len = S.slice_len(s), base = s * S.seg_len;
	const int64_t ks0 = __ldg(kseq + base);
	if (!(ks0 & KS_TIGHT)) { if (cb == 0 && lane == 0) tab_ok[id] = 0; return; }
	const int w0 = (int)(ks0 >> KS_WIDTH_SHIFT) & 127;
	if (cb * 32 > w0) return;
	uint32_t x = (uint32_t)(cb * 32 + lane);
	bool ok = true;
	for (int64_t u0 = 0; u0 < len && ok; u0 += 32) {
		const int64_t r = u0 + lane;
		const bool have = r < len;
		const int64_t ks = have ? __ldg(kseq + base + r) : (int64_t)KS_TIGHT;
		const uint4 m = have ? __ldg(wmask + base + r) : make_uint4(0, 0, 0, 0);
		const unsigned loose = __ballot_sync(0xffffffffu, !(ks & KS_TIGHT));
		int n = len - u0 < 32 ? (int)(len - u0) : 32;
		if (loose) { ok = false; n = 0; } /* a row without a mask: the general fix-up handles this cascade */
		for (int j = 0; j < n; ++j) {
			uint4 mj;
			mj.x = __shfl_sync(0xffffffffu, m.x, j); mj.y = __shfl_sync(0xffffffffu, m.y, j);
			mj.z = __shfl_sync(0xffffffffu, m.z, j); mj.w = __shfl_sync(0xffffffffu, m.w, j);
			x = mask_rank(mj, x);
		}
	}
	if (ok) tab[id * 128 + cb * 32 + lane] = (uint8_t)x;
	if (cb == 0 && lane == 0) tab_ok[id] = ok ? 1 : 0;
}

/* follow every cascade through the tables: one lookup per slice.  Every slice met is listed with the exact value it starts
 * from.  n_stop counts cascades that ran into a slice without a table (left to the general fix-up). */
__global__ void k_fix_hops(Slices S, const int64_t *__restrict__ kseq, const int32_t *__restrict__ nc_of, const uint8_t *__restrict__ tab, const uint8_t *__restrict__ tab_ok,
                           int64_t n_items, const int64_t *__restrict__ wl_seg, const int64_t *__restrict__ wl_val,
                           int64_t *__restrict__ out_seg, int64_t *__restrict__ out_val, unsigned long long *out_n, unsigned long long *n_stop)
{
	const int64_t it = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (it >= n_items) return;
	int64_t t = wl_seg[it], v = wl_val[it];
	for (;;) {
		const unsigned long long o = atomicAdd(out_n, 1ULL);
		out_seg[o] = t; out_val[o] = v;
		const int32_t id = nc_of[t - S.walk_lo];
		if (id < 0) break; /* collapses inside, or hands an exact value on: its successor is an item of its own */
		if (!tab_ok[id]) { atomicAdd(n_stop, 1ULL); break; } /* a row without a mask: the general fix-up carries on from here afterwards */
		const int64_t x = v - (kseq[t * S.seg_len] & (int64_t)RB3B_M42);
		if (x < 0 || x > TIGHT_WIDTH) { atomicAdd(n_stop, 1ULL); break; } /* cannot happen for a valid batch */
		v = S.arr_lo[t] + (int64_t)tab[(int64_t)id * 128 + x];
		++t;
		if (t >= S.own_hi || S.d[t] == 0) break;
	}
}

/* generic fix-up (RLE cells, or bitmap cells without the tables): re-walk the unresolved prefix of each listed slice
 * from its now exact start; slices that never collapsed put their successor on the next round's list */
template<class WG>
__global__ void __launch_bounds__(TPB) k_walk_fix(DevIndex A, Slices S, const uint8_t *__restrict__ wsym, int64_t *__restrict__ kseq, int64_t n_items,
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
		const int64_t t = wl_seg[it], d = S.d[t], len = S.slice_len(t), p0 = t * S.seg_len;
		const bool was_exact = S.arr_lo[t] == S.arr_hi[t]; /* then t + 1 was already listed by whoever made t's arrival exact */
		int64_t v = wl_val[it];
		int c = 1;
		for (int64_t i = 0; i < d; ++i) {
			c = (int)wsym[p0 + i];
			if (gl == 0) kseq[p0 + i] = v;
			if (c == 0) break;
			v = A.acc[c] + WG::rank(A, v, c);
		}
		if (gl == 0) {
			S.d[t] = 0;
			if (d == len && c != 0) { /* never collapsed: only now is the arrival value known */
				S.arr_lo[t] = S.arr_hi[t] = v;
				if (!was_exact && t + 1 < S.own_hi && S.d[t + 1] > 0) { /* t + 1 is processed by nobody else */
					unsigned long long o = atomicAdd(nx_n, 1ULL);
					nx_seg[o] = t + 1; nx_val[o] = v;
				}
			}
		}
	}
}

/* back to row order: ka[rows[i]] = vals[i] */
template<typename RowT, typename KaT>
__global__ void k_scatter_ka(int64_t n, const RowT *__restrict__ rows, const int64_t *__restrict__ vals, KaT *__restrict__ ka, unsigned long long *n_unres)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned int bad = 0;
	if (i < n) {
		const int64_t v = vals[i];
		if (v & KS_UNRES) bad = 1;
		else ka[(int64_t)rows[i]] = (KaT)v;
	}
	bad = __popc(__ballot_sync(0xffffffffu, bad));
	if (bad && (threadIdx.x & 31) == 0) atomicAdd(n_unres, (unsigned long long)bad);
}

/* rows still flagged unresolved (multi-device exchange: nothing else looks at every value before it leaves) */
__global__ void k_count_flagged(int64_t n, const int64_t *__restrict__ vals, unsigned long long *n_unres)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	unsigned int bad = i < n && (vals[i] & KS_UNRES) ? 1u : 0u;
	bad = __popc(__ballot_sync(0xffffffffu, bad));
	if (bad && (threadIdx.x & 31) == 0) atomicAdd(n_unres, (unsigned long long)bad);
}

/* ---- exchange between devices: (row, position) pairs are routed to the device that owns the row's range ---- */
#define MAX_RANKS 64

__global__ void __launch_bounds__(TPB) k_dest_count(int64_t n, const uint32_t *__restrict__ rows, int64_t chunk, int n_ranks, unsigned long long *__restrict__ cnt)
{
	__shared__ unsigned int sh[MAX_RANKS];
	if (threadIdx.x < MAX_RANKS) sh[threadIdx.x] = 0;
	__syncthreads();
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) atomicAdd(&sh[rows[i] / (uint32_t)chunk], 1u);
	__syncthreads();
	if (threadIdx.x < n_ranks && sh[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

/* base[d] = where destination d's pairs start in the packed arrays; cursor[d] = pairs of d placed so far */
__global__ void __launch_bounds__(TPB) k_dest_pack(int64_t n, const uint32_t *__restrict__ rows, const int64_t *__restrict__ vals, int64_t chunk, int n_ranks,
                                                   const int64_t *__restrict__ base, unsigned long long *__restrict__ cursor,
                                                   uint32_t *__restrict__ prow, int64_t *__restrict__ pval)
{
	__shared__ unsigned int sh[MAX_RANKS];
	__shared__ unsigned long long at[MAX_RANKS];
	if (threadIdx.x < MAX_RANKS) sh[threadIdx.x] = 0;
	__syncthreads();
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	uint32_t r = 0, d = 0, li = 0;
	if (i < n) { r = rows[i]; d = r / (uint32_t)chunk; li = atomicAdd(&sh[d], 1u); }
	__syncthreads();
	if (threadIdx.x < n_ranks && sh[threadIdx.x]) at[threadIdx.x] = atomicAdd(&cursor[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
	__syncthreads();
	if (i < n) {
		const int64_t o = base[d] + (int64_t)at[d] + li;
		prow[o] = r; pval[o] = vals[i];
	}
}

/* the pairs this device received: into its dense range of the interleave array */
__global__ void k_fill_dense(int64_t n, const uint32_t *__restrict__ rows, const int64_t *__restrict__ vals, int64_t row0, int64_t *__restrict__ dense)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) dense[(int64_t)rows[i] - row0] = vals[i];
}

/* A scatter of 8-byte values over a target much larger than L2 costs a DRAM read-modify-write of a sector per value.
 * For large batches the (row, value) pairs are therefore first partitioned by the high bits of the row (one or two
 * radix passes, streaming), so that the scatter proper works on windows of 2^19 rows (4 MB) that stay in L2. */
template<typename RowT, typename KaT>
static int scatter_to_rows(int64_t n, int64_t len, const RowT *rows, const int64_t *vals, KaT *ka, unsigned long long *n_unres)
{
	if (n <= 0) return RB3B_OK;
	const int win_bits = (int)rb3b_get_param("scatter_win_bits", 19);
	int bits = 1;
	while ((1LL << bits) < len) ++bits;
	if (n >= rb3b_get_param("scatter_bucket_min", 24LL << 20) && bits > win_bits) {
		DBuf<RowT> r2; DBuf<int64_t> v2; DBuf<uint8_t> tmp;
		size_t tb = 0;
		TRY(r2.alloc(n)); TRY(v2.alloc(n));
		CK(cub::DeviceRadixSort::SortPairs(0, tb, rows, r2.p, vals, v2.p, n, win_bits, bits, rb3b_stream));
		TRY(tmp.alloc(tb));
		CK(cub::DeviceRadixSort::SortPairs(tmp.p, tb, rows, r2.p, vals, v2.p, n, win_bits, bits, rb3b_stream));
		k_scatter_ka<RowT, KaT><<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, r2.p, v2.p, ka, n_unres); CKK();
	} else {
		k_scatter_ka<RowT, KaT><<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, rows, vals, ka, n_unres); CKK();
	}
	return RB3B_OK;
}

__global__ void k_widen_ka(int64_t n, const uint32_t *__restrict__ in, int64_t *__restrict__ out)
{
	const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i < n) out[i] = (int64_t)in[i];
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


/* everything of the rank phase that depends on the width of the batch's LF table */
int rb3b_all_gather(const void *send, void *recv, size_t bytes_per_rank); /* rb3b_dist.cu */

/* part / n_parts > 1: the ranks of the communicator share the first walk of the pieces (each walks a contiguous range of
 * fine marks, the 16-byte nodes are all-gathered over NVLink); LF table and list ranking are replicated */
template<typename LfT, typename RowT>
static int walk_order(int64_t len, const uint8_t *d_bwt, int64_t nt, const int64_t *tex, const Acc7 &acc, Fine &F,
                      int64_t p_lo, int64_t p_hi, DBuf<uint8_t> &wsym, void **wrow_out, const int64_t **chain_base_out, const int64_t **chain_len_out, const void **lf_out = 0,
                      int part = 0, int n_parts = 1)
{
	DBuf<LfT> lf;
	DBuf<RowT> wrow;
	TRY(lf.alloc(len)); TRY(wrow.alloc(len + 8)); /* padded: read in whole groups of eight */
	k_prep_lf<LfT><<<(unsigned)nt, TPB, 0, rb3b_stream>>>(len, d_bwt, nt, tex, acc, lf.p); CKK();
	DBuf<FNode> nd; /* two buffers of list-ranking nodes (ping-pong) */
	DBuf<int64_t> fc, ch; /* chain_of; chain_len, chain_base */
	DBuf<int32_t> pl;
	if (F.n_fine >= (1LL << 31)) return rb3b_fail(RB3B_EINVAL, "batch too large: %lld fine marks", (long long)F.n_fine);
	const bool share = n_parts > 1 && rb3b_cur()->world == n_parts && rb3b_cur()->comm != 0 && rb3b_get_param("share_fine_walk", 1) != 0;
	const int64_t f_chunk = share ? (F.n_fine + n_parts - 1) / n_parts : F.n_fine, f_pad = share ? f_chunk * n_parts : F.n_fine;
	TRY(nd.alloc(f_pad * 2)); TRY(fc.alloc(F.n_fine)); TRY(ch.alloc(F.n_seq * 2)); TRY(pl.alloc(F.n_fine));
	FNode *pp[2] = { nd.p, nd.p + f_pad };
	int64_t *f_cof = fc.p, *c_len = ch.p, *c_base = ch.p + F.n_seq;
	/* one chase instead of two (k_fine_walk_buf / k_copy_walk) where a chase step is a DRAM access: batches from 2^25 rows on
	 * ("piece_buf": 0 never, 1 always, -1 by size); single device only (with several, the walks are split by different keys) */
	const int64_t pbuf = rb3b_get_param("piece_buf", -1);
	const bool use_buf = !share && n_parts == 1 && (pbuf > 0 || (pbuf < 0 && len >= (32LL << 20) && sizeof(RowT) == 4)); /* by size: 32-bit rows only (what was measured) */
	const int cap = F.fshift >= 2 ? (int)(2LL << F.fshift) : 8; /* twice the mean piece length; a multiple of 8 (whole-word stores) */
	DBuf<RowT> rowbuf, tail_row;
	DBuf<uint8_t> symbuf;
	if (share) {
		const int64_t f_lo = f_chunk * part, f_hi = f_lo + f_chunk < F.n_fine ? f_lo + f_chunk : F.n_fine;
		if (f_hi > f_lo) { k_fine_walk<LfT><<<nblk(f_hi - f_lo, TPB), TPB, 0, rb3b_stream>>>(F, lf.p, pp[0], pl.p, f_lo, f_hi); CKK(); }
		TRY(rb3b_all_gather(pp[0] + f_lo, pp[0], (size_t)f_chunk * sizeof(FNode)));
		k_piece_len<<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F.n_fine, pp[0], pl.p); CKK();
	} else if (use_buf) {
		TRY(rowbuf.alloc((size_t)F.n_fine * cap)); TRY(symbuf.alloc((size_t)F.n_fine * cap)); TRY(tail_row.alloc(F.n_fine));
		k_fine_walk_buf<LfT, RowT><<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, lf.p, pp[0], pl.p, rowbuf.p, symbuf.p, cap, tail_row.p); CKK();
	} else { k_fine_walk<LfT><<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, lf.p, pp[0], pl.p, 0, F.n_fine); CKK(); }
	int cur = 0, n_rounds = 0;
	for (int64_t span = 1; span < F.n_fine; span <<= 1) ++n_rounds;
	/* (a cooperative single-launch version with grid-wide barriers was measured: 0.518 vs 0.506 ms of prep -- the rounds are
	 * bound by their dependent L2 gathers, not by launches) */
	for (int r = 0; r < n_rounds; ++r) {
		k_list_rank<<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F.n_fine, pp[cur], pp[cur ^ 1]); CKK();
		cur ^= 1;
	}
	const FNode *node = pp[cur];
	CK(cudaMemsetAsync(f_cof, 0xff, F.n_fine * 8, rb3b_stream));
	k_chain_len<<<nblk(F.n_seq, TPB), TPB, 0, rb3b_stream>>>(F.n_seq, node, c_len, f_cof); CKK();
	TRY(rb3b_scan_excl_i64(c_len, c_base, F.n_seq));
	int64_t last[2];
	CK(cudaMemcpyAsync(&last[0], c_base + F.n_seq - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&last[1], c_len + F.n_seq - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (last[0] + last[1] != len) /* LF_B is a permutation, so the chains are disjoint: they cover the batch iff their lengths add up */
		return rb3b_fail(RB3B_EINVAL, "batch is not the BWT of a sentinel-terminated string set (%lld of %lld rows reachable from the sentinels)",
		                 (long long)(last[0] + last[1]), (long long)len);
	/* (Tried: the pieces sorted by length with one 8-bit radix pass, so that the lanes of a warp run equally long and the
	 * longest chases start first -- ncu shows 8.4 of 32 lanes busy here.  Slower: 0.57 vs 0.51 ms of prep with one genome
	 * per batch, 6.5 vs 5.6 ms with ten; neighbouring pieces start at neighbouring rows, the sorted order gives that up.) */
	if (use_buf) { k_copy_walk<LfT, RowT><<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, len, lf.p, node, f_cof, c_base, c_len, pl.p, rowbuf.p, symbuf.p, cap, tail_row.p, wrow.p, wsym.p); CKK(); }
	else { k_write_walk<LfT, RowT><<<nblk(F.n_fine, TPB), TPB, 0, rb3b_stream>>>(F, len, lf.p, node, f_cof, c_base, c_len, pl.p, p_lo, p_hi, wrow.p, wsym.p); CKK(); }
	*wrow_out = wrow.p; *chain_base_out = c_base; *chain_len_out = c_len; /* arena memory: lives until the API call returns */
	if (lf_out) *lf_out = lf.p;
	return RB3B_OK;
}

/* interleave positions of the batch in device memory: ka[len], accB */
/* part/n_parts: device `part` resolves the slices [part, part+1) * n_seg / n_parts and walks a halo of slices before them
 * speculatively so that its first slice receives an exact value without any exchange.  ka_out != NULL: write there
 * instead of allocating (rows of other parts are set to -1).  *incomplete is set when a part could not resolve all of
 * its own rows locally (only possible with n_parts > 1). */
/* what a device resolved itself, still in walk order: (row, position) pairs for the exchange between devices */
struct OwnPairs { const uint32_t *rows; const int64_t *vals; int64_t n; };

/* pairs != 0 (32-bit rows only): nothing is scattered, ka stays untouched, the device's own rows are handed back.
 * pre != 0: the batch comes with its walk order (rb3b_batch_prepare: straight from the suffix sort) -- no LF table, no
 * chase, no list ranking; d_bwt is not read. */
static int rank_phase(const rb3b_index_s *A, int64_t len, const uint8_t *d_bwt, DBuf<int64_t> &ka, int64_t accB[RB3B_ASIZE + 1],
                      int part = 0, int n_parts = 1, int64_t *ka_out = 0, int *incomplete = 0, int so = 0, OwnPairs *pairs = 0,
                      const rb3b_batch_s *pre = 0, uint32_t *ka32 = 0 /* multi-device: 32-bit partial array (0 = nobody's), every position fits */,
                      const unsigned long long **defer_unres = 0 /* single device: leave the count of unresolved rows on the device (arena memory)
                                                                    for the merge to check with its own read-back, instead of waiting for it here */)
{
	int64_t nt = (len + PREP_TILE - 1) / PREP_TILE;
	DBuf<int64_t> tcnt, tex;
	DBuf<int> bad;
	int hbad = 0;
	if (len <= 0) return rb3b_fail(RB3B_EINVAL, "empty batch");
	if (A->n_cells == 0) return rb3b_fail(RB3B_EINVAL, "rank phase on an empty index");
	Acc7 acc;
	rb3b_tic(T_PREP);
	if (pre) memcpy(acc.v, pre->acc, sizeof(acc.v));
	else {
	/* batch LF mapping */
	TRY(tcnt.alloc((nt + 1) * RB3B_ASIZE)); TRY(tex.alloc((nt + 1) * RB3B_ASIZE)); TRY(bad.alloc(1));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	k_prep_count<<<(unsigned)nt, TPB, 0, rb3b_stream>>>(len, d_bwt, nt, tcnt.p, bad.p); CKK();
	TRY(rb3b_scan_excl_i64(tcnt.p, tex.p, (nt + 1) * RB3B_ASIZE));
	int64_t tot[RB3B_ASIZE + 1], base[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
	DBuf<int64_t> gt;
	TRY(gt.alloc(RB3B_ASIZE + 1));
	k_gather_tot<<<1, 32, 0, rb3b_stream>>>(tex.p, nt + 1, bad.p, gt.p); CKK();
	CK(cudaMemcpyAsync(tot, gt.p, sizeof(tot), cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	hbad = (int)tot[RB3B_ASIZE];
	if (hbad) return rb3b_fail(RB3B_EINVAL, "batch BWT holds a symbol >= %d", RB3B_ASIZE);
	acc.v[0] = 0;
	for (int a = 0; a < RB3B_ASIZE; ++a) acc.v[a + 1] = acc.v[a] + (tot[a] - base[a]);
	}
	memcpy(accB, acc.v, sizeof(acc.v));
	if (acc.v[1] <= 0) return rb3b_fail(RB3B_EINVAL, "batch BWT holds no sentinel");
	/* the batch in walk order */
	/* few slices (one genome per batch): the walks are latency bound, shorter slices and pieces give more of them; many
	 * slices: DRAM-access bound, longer slices leave fewer rows to the fix-up */
	const bool big = len >= (32LL << 20);
	/* the slice length follows the rows THIS device walks */
	int64_t seg_len = ((rb3b_seg_len > 0 ? rb3b_seg_len : len / n_parts >= (32LL << 20) ? 512 : A->kind == RB3B_KIND_BM ? 192 : 384) + 7) / 8 * 8;
	Fine F;
	F.n_seq = acc.v[1];
	int64_t fine_len = rb3b_get_param("fine_len", 0);
	if (fine_len <= 0) fine_len = big ? 32 : 16;
	if (fine_len > seg_len) fine_len = seg_len;
	F.fshift = 0;
	while ((2LL << F.fshift) <= fine_len) ++F.fshift; /* largest power of two <= fine_len */
	F.m0 = (F.n_seq + (1LL << F.fshift) - 1) >> F.fshift;
	int64_t n_samp = ((len - 1) >> F.fshift) - F.m0 + 1;
	F.n_fine = F.n_seq + (n_samp > 0 ? n_samp : 0);
	/* slices */
	Slices S;
	S.len = len; S.seg_len = seg_len;
	S.n_seg = (len + seg_len - 1) / seg_len;
	S.own_lo = 0; S.own_hi = S.n_seg; S.walk_lo = 0;
	if (n_parts > 1) {
		S.own_lo = S.n_seg * part / n_parts; S.own_hi = S.n_seg * (part + 1) / n_parts;
		S.walk_lo = S.own_lo - rb3b_get_param("halo_segments", 8);
		if (S.walk_lo < 0) S.walk_lo = 0;
	}
	DBuf<uint8_t> wsym;
	void *wrow = 0;
	if (pre) wsym.p = pre->wsym;
	else TRY(wsym.alloc(len + 64)); /* padded: the walks read whole 8-byte words, the fix-up whole chunks */
	const bool narrow_lf = pre ? true : len < LF32_MAX_LEN && !rb3b_get_param("wide_lf", 0);
	/* walk-order positions this device reads (everything for a sorted collection: the heads are resolved on every device) */
	/* rows walked before every slice to narrow its bracket (k_walk_pair) */
	int warm = (int)rb3b_get_param("warm_rows", 16);
	warm = warm < 0 ? 0 : (warm + 7) / 8 * 8;
	if (warm > seg_len) warm = (int)seg_len;
	const int64_t p_lo = so ? -1 : S.walk_lo * seg_len - warm - 1, p_hi = so ? len : S.own_hi * seg_len;
	const int64_t *c_base = 0, *c_len = 0;
	if (pre) { wrow = pre->wrow; c_base = pre->c_base; c_len = pre->c_len; F.n_fine = 0; }
	else if (narrow_lf) TRY((walk_order<uint32_t, uint32_t>(len, d_bwt, nt, tex.p, acc, F, p_lo, p_hi, wsym, &wrow, &c_base, &c_len, 0, part, n_parts)));
	else TRY((walk_order<uint64_t, int64_t>(len, d_bwt, nt, tex.p, acc, F, p_lo, p_hi, wsym, &wrow, &c_base, &c_len, 0, part, n_parts)));
	const int64_t n_walk = S.own_hi - S.walk_lo;
	DBuf<int64_t> seg, wl, ctr, kseq;
	TRY(seg.alloc(S.n_seg * 3)); TRY(wl.alloc(S.n_seg * 4)); TRY(ctr.alloc(16)); TRY(kseq.alloc(len + 64));
	if (pairs && !narrow_lf) return rb3b_fail(RB3B_EINVAL, "internal error: pair output needs 32-bit rows");
	if (pairs || ka32) ka.p = 0;
	else if (ka_out) ka.p = ka_out; else TRY(ka.alloc(len));
	S.d = seg.p; S.arr_lo = seg.p + S.n_seg; S.arr_hi = seg.p + 2 * S.n_seg;
	rb3b_toc(T_PREP);
	/* everything up to here needed the batch only: it may have run while the previous merge into A was still writing the
	 * cells (asynchronous merge); from here on the kernels read A, and ka may be the scratch that merge was reading */
	if (A->broken) return rb3b_fail(RB3B_EINVAL, "the index was left unusable by an earlier failed merge");
	TRY(rb3b_index_use(A));
	CK(cudaMemsetAsync(ctr.p, 0, 16 * 8, rb3b_stream));
	if (ka32) CK(cudaMemsetAsync(ka32, 0, len * 4, rb3b_stream));
	else if (n_parts > 1 && !pairs) CK(cudaMemsetAsync(ka.p, 0xff, len * 8, rb3b_stream));
	DevIndex dA = rb3b_dev_view(A);
	const bool bm = A->kind == RB3B_KIND_BM;
	/* bitmap walks are lane pairs (single threads in the fix-up): small CTAs spread the walks over all SMs */
	const bool pair = bm && rb3b_get_param("walk_pair", 1) != 0;
	/* run-length cells: one thread per walk as well ("rle_t1"; 0: groups of 8 lanes, ~10x the instructions per row) */
	const bool t1 = !bm && rb3b_get_param("rle_t1", 1) != 0;
	const int wg = bm ? (pair ? 2 : 1) : t1 ? 1 : 8, wtpb = (bm || t1) ? 32 : TPB;
	int64_t want = (n_walk * wg + wtpb - 1) / wtpb, cap = (int64_t)n_sm() * ((bm || t1) ? 32 : 8);
	if (want < 1) want = 1;
	rb3b_tic(T_WALK1);
	const unsigned wgrid = (unsigned)(want < cap ? want : cap);
	/* transfer masks of the unresolved rows (16 bytes per row of the batch): only for batches that can afford them */
	DBuf<uint4> wmask;
	const bool use_mask = pair && len <= rb3b_get_param("mask_max_rows", 1LL << 32); /* 16 bytes per row: every batch the device can hold (rb3b_max_batch_symbols counts them) */
	if (use_mask) TRY(wmask.alloc(len + 64));
	const bool use_log = bm && rb3b_get_param("fix_log", 1) != 0;
	if (so) {
		CK(cudaMemsetAsync(kseq.p, 0, (len + 64) * 8, rb3b_stream));
		if (bm) k_so_heads<BmRank><<<nblk(F.n_seq, TPB), TPB, 0, rb3b_stream>>>(dA, so, F.n_seq, c_base, c_len, wsym.p, kseq.p);
		else k_so_heads<Grp<8> ><<<nblk(F.n_seq * 8, TPB), TPB, 0, rb3b_stream>>>(dA, so, F.n_seq, c_base, c_len, wsym.p, kseq.p);
		CKK();
		if (pair) k_walk_pair<true><<<wgrid, wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, use_mask ? wmask.p : 0, warm, ctr.p);
		else if (bm) k_walk_first<BmRank, true><<<wgrid, wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, ctr.p);
		else if (t1) k_walk_first<RleT1, true><<<wgrid, wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, ctr.p);
		else k_walk_first<Grp<8>, true><<<wgrid, wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, ctr.p);
	} else {
		if (pair) k_walk_pair<false><<<wgrid, wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, use_mask ? wmask.p : 0, warm, ctr.p);
		else if (bm) k_walk_first<BmRank, false><<<wgrid, wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, ctr.p);
		else if (t1) k_walk_first<RleT1, false><<<wgrid, wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, ctr.p);
		else k_walk_first<Grp<8>, false><<<wgrid, wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, ctr.p);
	}
	CKK();
	rb3b_toc(T_WALK1);
	int64_t *wl_seg[2] = { wl.p, wl.p + 2 * S.n_seg }, *wl_val[2] = { wl.p + S.n_seg, wl.p + 3 * S.n_seg };
	const bool use_tab = use_log && use_mask && !so && rb3b_get_param("fix_tables", 0) != 0;
	DBuf<int32_t> nc; /* nc_of[n_walk], nc_list[n_walk] */
	if (use_tab) TRY(nc.alloc((size_t)(n_walk > 0 ? n_walk : 1) * 2));
	k_collect_first<<<nblk(n_walk > 0 ? n_walk : 1, TPB), TPB, 0, rb3b_stream>>>(S, wl_seg[0], wl_val[0], (unsigned long long*)(ctr.p + 1),
		use_tab ? nc.p : 0, use_tab ? nc.p + n_walk : 0, (unsigned long long*)(ctr.p + 10)); CKK();
	int64_t n_items = 0, rounds = 1, fix_items = 0, cnt2[2] = {0, 0};
	int cur = 0;
	if (use_log && !use_tab && n_walk > 0) {
		/* the chain fix-up resolves every cascade in ONE launch, so it needs no list size on the host: the grid covers the
		 * upper bound (every slice on the list) and the kernel reads the count the collection left on the device.  One
		 * read-back per rank phase less. */
		rb3b_tic(T_WALKFIX);
		TRY(launch_fix_chain(dA, S, wsym.p, kseq.p, use_mask ? wmask.p : 0, n_walk, wl_seg[0], wl_val[0], (unsigned long long*)(ctr.p + 4), 1, (const unsigned long long*)(ctr.p + 1)));
		rb3b_toc(T_WALKFIX);
		int64_t fst[4] = {0, 0, 0, 0};
		CK(cudaMemcpyAsync(&n_items, ctr.p + 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaMemcpyAsync(fst + 1, ctr.p + 4, 24, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		rb3b_stat_set("fix_rows", fst[1]); rb3b_stat_set("fix_wide_rows", fst[2]); rb3b_stat_set("fix_longest_chain", fst[3]);
		rb3b_tflush();
		fix_items += n_items;
		if (n_items > 0) ++rounds;
		n_items = 0; /* nothing is ever left for a second round */
	} else {
		CK(cudaMemcpyAsync(&n_items, ctr.p + 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaMemcpyAsync(cnt2, ctr.p + 10, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
	}
	const int64_t n_nc = cnt2[0];
	if (n_items > 0 && use_tab && n_nc > 0) {
		/* cascades through transfer tables (above k_fix_tables): every slice with unresolved rows gets its exact start, then
		 * all are resolved at once -- the longest dependent chain is one slice instead of the longest cascade.  Whatever is
		 * left (a cascade through a row without a mask) goes through the cascading loop below. */
		DBuf<uint8_t> tab, tab_ok;
		TRY(tab.alloc((size_t)n_nc * 128)); TRY(tab_ok.alloc(n_nc));
		CK(cudaMemsetAsync(ctr.p + 2, 0, 40, rb3b_stream));
		CK(cudaMemsetAsync(ctr.p + 9, 0, 8, rb3b_stream));
		CK(cudaMemsetAsync(ctr.p + 11, 0, 8, rb3b_stream));
		rb3b_tic(T_WALKFIX);
		k_fix_tables<<<nblk(n_nc * 4 * 32, 128), 128, 0, rb3b_stream>>>(S, kseq.p, wmask.p, n_nc, nc.p + n_walk, tab.p, tab_ok.p); CKK();
		k_fix_hops<<<nblk(n_items, TPB), TPB, 0, rb3b_stream>>>(S, kseq.p, nc.p, tab.p, tab_ok.p, n_items, wl_seg[0], wl_val[0], wl_seg[1], wl_val[1],
			(unsigned long long*)(ctr.p + 9), (unsigned long long*)(ctr.p + 11)); CKK();
		int64_t n2 = 0, n_stop = 0;
		CK(cudaMemcpyAsync(&n2, ctr.p + 9, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaMemcpyAsync(&n_stop, ctr.p + 11, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		if (n2 > 0) TRY(launch_fix_chain(dA, S, wsym.p, kseq.p, wmask.p, n2, wl_seg[1], wl_val[1], (unsigned long long*)(ctr.p + 4), 0));
		rb3b_toc(T_WALKFIX);
		fix_items += n2;
		n_items = 0;
		int64_t fst[4] = {0, 0, 0, 0};
		if (n_stop > 0) { /* leftovers: slices behind a cascade that stopped at a slice without a table */
			CK(cudaMemsetAsync(ctr.p + 1, 0, 8, rb3b_stream));
			k_collect_first<<<nblk(n_walk > 0 ? n_walk : 1, TPB), TPB, 0, rb3b_stream>>>(S, wl_seg[0], wl_val[0], (unsigned long long*)(ctr.p + 1)); CKK();
			CK(cudaMemcpyAsync(&n_items, ctr.p + 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		}
		CK(cudaMemcpyAsync(fst + 1, ctr.p + 4, 24, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		rb3b_stat_set("fix_rows", fst[1]); rb3b_stat_set("fix_wide_rows", fst[2]); rb3b_stat_set("fix_longest_chain", fst[3]);
		rb3b_stat_set("fix_table_slices", n_nc); rb3b_stat_set("fix_leftover_items", n_items);
		rb3b_tflush();
		++rounds;
	}
	while (n_items > 0) {
		/* ctr[2] = item cursor, ctr[3] = size of the next list, ctr[4..6] = statistics */
		CK(cudaMemsetAsync(ctr.p + 2, 0, 40, rb3b_stream));
		want = (n_items * wg + wtpb - 1) / wtpb;
		rb3b_tic(T_WALKFIX);
		if (use_log) { TRY(launch_fix_chain(dA, S, wsym.p, kseq.p, use_mask ? wmask.p : 0, n_items, wl_seg[cur], wl_val[cur], (unsigned long long*)(ctr.p + 4), 1)); --rb3b_n_launch; }
		else if (bm) k_walk_fix<BmRank><<<(unsigned)(want < cap ? want : cap), wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, n_items, wl_seg[cur], wl_val[cur], ctr.p + 2,
			wl_seg[cur ^ 1], wl_val[cur ^ 1], (unsigned long long*)(ctr.p + 3));
		else if (t1) k_walk_fix<RleT1><<<(unsigned)(want < cap ? want : cap), wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, n_items, wl_seg[cur], wl_val[cur], ctr.p + 2,
			wl_seg[cur ^ 1], wl_val[cur ^ 1], (unsigned long long*)(ctr.p + 3));
		else k_walk_fix<Grp<8> ><<<(unsigned)(want < cap ? want : cap), wtpb, 0, rb3b_stream>>>(dA, S, wsym.p, kseq.p, n_items, wl_seg[cur], wl_val[cur], ctr.p + 2,
			wl_seg[cur ^ 1], wl_val[cur ^ 1], (unsigned long long*)(ctr.p + 3));
		CKK();
		rb3b_toc(T_WALKFIX);
		fix_items += n_items;
		int64_t fst[4] = {0, 0, 0, 0};
		CK(cudaMemcpyAsync(fst, ctr.p + 3, 32, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		n_items = fst[0];
		if (use_log) { rb3b_stat_set("fix_rows", fst[1]); rb3b_stat_set("fix_wide_rows", fst[2]); rb3b_stat_set("fix_longest_chain", fst[3]); }
		rb3b_tflush();
		cur ^= 1; ++rounds;
	}
	/* back to row order */
	unsigned long long unres = 0;
	const int64_t own_rows = (S.own_hi * seg_len < len ? S.own_hi * seg_len : len) - S.own_lo * seg_len;
	CK(cudaMemsetAsync(ctr.p + 8, 0, 8, rb3b_stream));
	rb3b_tic(T_SCATTER);
	const int64_t own_p0 = S.own_lo * seg_len;
	if (pairs) {
		pairs->rows = (const uint32_t*)wrow + own_p0; pairs->vals = kseq.p + own_p0; pairs->n = own_rows;
		k_count_flagged<<<nblk(own_rows > 0 ? own_rows : 1, TPB), TPB, 0, rb3b_stream>>>(own_rows, pairs->vals, (unsigned long long*)(ctr.p + 8)); CKK();
	} else if (ka32 && narrow_lf) TRY((scatter_to_rows<uint32_t, uint32_t>(own_rows, len, (const uint32_t*)wrow + own_p0, kseq.p + own_p0, ka32, (unsigned long long*)(ctr.p + 8))));
	else if (ka32) TRY((scatter_to_rows<int64_t, uint32_t>(own_rows, len, (const int64_t*)wrow + own_p0, kseq.p + own_p0, ka32, (unsigned long long*)(ctr.p + 8))));
	else if (narrow_lf) TRY((scatter_to_rows<uint32_t, int64_t>(own_rows, len, (const uint32_t*)wrow + own_p0, kseq.p + own_p0, ka.p, (unsigned long long*)(ctr.p + 8))));
	else TRY((scatter_to_rows<int64_t, int64_t>(own_rows, len, (const int64_t*)wrow + own_p0, kseq.p + own_p0, ka.p, (unsigned long long*)(ctr.p + 8))));
	rb3b_toc(T_SCATTER);
	const bool deferred = defer_unres != 0 && n_parts == 1;
	if (deferred) *defer_unres = (const unsigned long long*)(ctr.p + 8);
	else {
		CK(cudaMemcpyAsync(&unres, ctr.p + 8, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		rb3b_tflush();
	}
	rb3b_stat_set("n_segments", S.n_seg);
	rb3b_stat_set("seg_len_used", seg_len);
	rb3b_stat_set("n_fine", F.n_fine);
	rb3b_stat_set("fix_rounds", rounds - 1);
	rb3b_stat_add("fix_rounds_total", rounds - 1);
	rb3b_stat_set("fix_segments", fix_items);
	rb3b_stat_set("unresolved_rows", deferred ? -1 : (int64_t)unres);
	rb3b_stat_set("own_segments", S.own_hi - S.own_lo);
	rb3b_stat_set("own_rows", own_rows);
	if (n_parts > 1) { /* completeness of the whole batch is checked after the exchange */
		if (incomplete) *incomplete = unres != 0;
		return RB3B_OK;
	}
	if (unres != 0)
		return rb3b_fail(RB3B_EINVAL, "batch is not the BWT of a sentinel-terminated string set (%lld rows unresolved)", (long long)unres);
	return RB3B_OK;
}


/* ------------------------------------------------------------------ */
/* sampled suffix array (`ropebwt3 ssa`)                     */
/* ------------------------------------------------------------------ */
/*
 * rb3_ssa_gen / ssa_gen1 (ssa.c:17-81) walk every string from its sentinel with one rank1a per symbol and remember, for
 * the rows at multiples of 2^ss behind the sentinel block, how far from the start of the string they are.  That walk is
 * the walk order of the collection itself: the index's own BWT goes through the same list ranking as a batch
 * (walk_order), after which position p of chain k0 IS "l = p - base steps from sentinel k0", so
 *     ssa[(row - C[1]) >> ss] = (L - 1 - l) << ms | k0        (ssa.c:28-38)
 *     r2i[LF(last row of the chain)] = k0                      (ssa.c:36)
 * are one data-parallel pass.  No dependent chain of length n/m is left.
 */
template<typename LfT, typename RowT>
__global__ void k_ssa_emit(int64_t len, int64_t n_seq, int64_t acc1, int ss, int ms, const RowT *__restrict__ wrow, const LfT *__restrict__ lf,
                           const int64_t *__restrict__ chain_base, const int64_t *__restrict__ chain_len, uint64_t *__restrict__ ssa, uint64_t *__restrict__ r2i)
{
	const int64_t p = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (p >= len) return;
	int64_t lo = 0, hi = n_seq; /* last chain with base <= p */
	while (hi - lo > 1) {
		const int64_t mid = (lo + hi) >> 1;
		if (chain_base[mid] <= p) lo = mid; else hi = mid;
	}
	const int64_t l = p - chain_base[lo], L = chain_len[lo], row = (int64_t)wrow[p];
	if (l > 0 && row >= acc1 && ((row - acc1) & ((1LL << ss) - 1)) == 0)
		ssa[(row - acc1) >> ss] = (uint64_t)(L - 1 - l) << ms | (uint64_t)lo;
	if (l == L - 1) r2i[(int64_t)((uint64_t)lf[row] >> LF_SHIFT)] = (uint64_t)lo;
}

template<typename LfT, typename RowT>
static int ssa_gen_t(int64_t len, const uint8_t *d_bwt, int64_t nt, const int64_t *tex, const Acc7 &acc, Fine &F, int ss, int ms, uint64_t *d_ssa, uint64_t *d_r2i)
{
	DBuf<uint8_t> wsym;
	void *wrow = 0;
	const void *lf = 0;
	const int64_t *c_base = 0, *c_len = 0;
	TRY(wsym.alloc(len + 16));
	TRY((walk_order<LfT, RowT>(len, d_bwt, nt, tex, acc, F, -1, len, wsym, &wrow, &c_base, &c_len, &lf)));
	k_ssa_emit<LfT, RowT><<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, F.n_seq, acc.v[1], ss, ms, (const RowT*)wrow, (const LfT*)lf, c_base, c_len, d_ssa, d_r2i); CKK();
	return RB3B_OK;
}

/* device arrays: d_r2i[m], d_ssa[n_ssa] (sizes from rb3b_ssa_sizes) */
extern "C" int rb3b_ssa_sizes(const rb3b_index_t *x, int ssa_shift, int64_t *m, int64_t *n_ssa, int *ms)
{ /* rb3_ssa_gen, ssa.c:62-66 */
	if (ssa_shift < 0 || ssa_shift > 30) return rb3b_fail(RB3B_EINVAL, "bad SSA sample shift %d", ssa_shift);
	*m = x->acc[1];
	int b = 1;
	while ((1LL << b) < *m) ++b;
	*ms = b;
	*n_ssa = (x->n - x->acc[1] + (1LL << ssa_shift) - 1) >> ssa_shift;
	return RB3B_OK;
}

extern "C" int rb3b_ssa_gen_dev(const rb3b_index_t *x, int ssa_shift, uint64_t *d_r2i, uint64_t *d_ssa)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	int64_t m, n_ssa;
	int ms;
	TRY(rb3b_ssa_sizes(x, ssa_shift, &m, &n_ssa, &ms));
	if (x->n == 0 || m == 0) return rb3b_fail(RB3B_EINVAL, "SSA of an empty index");
	const int64_t len = x->n, nt = (len + PREP_TILE - 1) / PREP_TILE;
	DBuf<uint8_t> plain;
	DBuf<int64_t> tcnt, tex;
	DBuf<int> bad;
	TRY(rb3b_index_to_plain_dev(x, plain));
	TRY(tcnt.alloc((nt + 1) * RB3B_ASIZE)); TRY(tex.alloc((nt + 1) * RB3B_ASIZE)); TRY(bad.alloc(1));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	k_prep_count<<<(unsigned)nt, TPB, 0, rb3b_stream>>>(len, plain.p, nt, tcnt.p, bad.p); CKK();
	TRY(rb3b_scan_excl_i64(tcnt.p, tex.p, (nt + 1) * RB3B_ASIZE));
	Acc7 acc;
	for (int a = 0; a <= RB3B_ASIZE; ++a) acc.v[a] = x->acc[a]; /* the batch IS the index: its C[] is known */
	Fine F;
	F.n_seq = m;
	F.fshift = len >= (32LL << 20) ? 5 : 4;
	F.m0 = (F.n_seq + (1LL << F.fshift) - 1) >> F.fshift;
	int64_t n_samp = ((len - 1) >> F.fshift) - F.m0 + 1;
	F.n_fine = F.n_seq + (n_samp > 0 ? n_samp : 0);
	CK(cudaMemsetAsync(d_ssa, 0, n_ssa * 8, rb3b_stream));
	CK(cudaMemsetAsync(d_r2i, 0, m * 8, rb3b_stream));
	if (len < LF32_MAX_LEN && !rb3b_get_param("wide_lf", 0)) TRY((ssa_gen_t<uint32_t, uint32_t>(len, plain.p, nt, tex.p, acc, F, ssa_shift, ms, d_ssa, d_r2i)));
	else TRY((ssa_gen_t<uint64_t, int64_t>(len, plain.p, nt, tex.p, acc, F, ssa_shift, ms, d_ssa, d_r2i)));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}

/* rb3_ssa_gen + rb3_ssa_dump (ssa.c:55-81,198-213): the .ssa file of `ropebwt3 ssa -s ssa_shift`; fn "-" = stdout */
extern "C" int rb3b_ssa_dump(const rb3b_index_t *x, int ssa_shift, const char *fn)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	int64_t m, n_ssa;
	int ms;
	TRY(rb3b_ssa_sizes(x, ssa_shift, &m, &n_ssa, &ms));
	DBuf<uint64_t> d;
	TRY(d.alloc(m + n_ssa));
	TRY(rb3b_ssa_gen_dev(x, ssa_shift, d.p, d.p + m));
	uint64_t *h = (uint64_t*)malloc((size_t)(m + n_ssa) * 8);
	if (h == 0) return rb3b_fail(RB3B_ENOMEM, "out of host memory");
	cudaError_t e = cudaMemcpyAsync(h, d.p, (size_t)(m + n_ssa) * 8, cudaMemcpyDeviceToHost, rb3b_stream);
	if (e == cudaSuccess) e = cudaStreamSynchronize(rb3b_stream);
	if (e != cudaSuccess) { free(h); return rb3b_fail(RB3B_ENODEV, "copying the SSA to the host: %s", cudaGetErrorString(e)); }
	FILE *fp = strcmp(fn, "-") ? fopen(fn, "wb") : stdout;
	if (!fp) { free(h); return rb3b_fail(RB3B_EIO, "failed to open '%s' for writing", fn); }
	uint32_t y;
	fwrite("SSA\1", 1, 4, fp);
	y = (uint32_t)ssa_shift; fwrite(&y, 4, 1, fp);
	y = (uint32_t)ms; fwrite(&y, 4, 1, fp);
	fwrite(&m, 8, 1, fp);
	fwrite(&n_ssa, 8, 1, fp);
	size_t w = fwrite(h, 8, (size_t)(m + n_ssa), fp);
	free(h);
	if (fp != stdout) fclose(fp); else fflush(fp);
	return w == (size_t)(m + n_ssa) ? RB3B_OK : rb3b_fail(RB3B_EIO, "short write to '%s'", fn);
}

/* ------------------------------------------------------------------ */
/* streaming merge (rb3b_emit.cuh)                                      */
/* ------------------------------------------------------------------ */

/* grow-only scratch of an index that outlives API calls (asynchronous merge) */
static int ms_reserve(rb3b_index_s *x, size_t need)
{
	if (x->ms_cap >= need) return RB3B_OK;
	TRY(rb3b_index_wait_i(x)); /* the merge in flight reads the old region */
	if (x->ms) CK(cudaFree(x->ms));
	x->ms = 0; x->ms_cap = 0;
	const size_t want = need + need / 4;
	if (cudaMalloc((void**)&x->ms, want) != cudaSuccess) { cudaGetLastError(); return rb3b_fail(RB3B_ENOMEM, "cannot allocate %zu bytes of merge scratch", want); }
	x->ms_cap = want;
	return RB3B_OK;
}

static inline size_t al512(size_t b) { return (b + 511) & ~(size_t)511; }

/* accB != 0 (the batch's C[], so that the merged totals are known beforehand) and a bitmap -> bitmap merge: the merge is
 * only QUEUED, on the context's second stream; d_ka and d_bwt must then lie in A->ms (first A->ms_used bytes) and the
 * merge's own tables are taken from the rest of A->ms.  The caller's next call overlaps with it up to the point where it
 * needs the merged cells (rb3b_index_use); errors the device finds are reported by rb3b_index_wait_i. */
__global__ void k_flag_if(const unsigned long long *__restrict__ cnt, int *__restrict__ flag) { if (*cnt) *flag = 1; }

/* d_unres != 0: the rank phase left its count of unresolved rows on the device; non-zero = the batch is not a valid BWT,
 * reported with the merge's own validation */
static int merge_phase(rb3b_index_s *A, int64_t len, const uint8_t *d_bwt, const int64_t *d_ka, const int64_t *accB = 0, const unsigned long long *d_unres = 0)
{
	if (accB != 0 && A->kind == RB3B_KIND_BM && rb3b_want_bitmap(A->n + len) && (const char*)d_ka == A->ms) {
		rb3b_ctx_s *ctx = rb3b_cur();
		for (int a = 0; a < RB3B_ASIZE; ++a) A->pend_expect[a] = A->tot[a] + (accB[a + 1] - accB[a]);
		if (A->pend_host == 0) CK(cudaMallocHost((void**)&A->pend_host, 8 * sizeof(int64_t)));
		CK(cudaEventRecord(ctx->ev_hand, ctx->stream)); /* stream2 goes on where the rank phase stops */
		CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_hand, 0));
		if (A->has_ev) CK(cudaStreamWaitEvent(ctx->stream2, A->ready_ev, 0));
		else { CK(cudaEventCreateWithFlags(&A->ready_ev, cudaEventDisableTiming)); A->has_ev = 1; }
		cudaStream_t s1 = ctx->stream;
		ctx->stream = ctx->stream2;
		ctx->bump = A->ms + A->ms_used; ctx->bump_off = 0; ctx->bump_cap = A->ms_cap - A->ms_used;
		int rc = RB3B_OK;
		do {
			DBuf<int> bad;
			if ((rc = bad.alloc(1)) != RB3B_OK) break;
			if (cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream) != cudaSuccess) { rc = rb3b_fail(RB3B_ENODEV, "cudaMemsetAsync failed"); break; }
			rb3b_tic(T_MERGE);
			if (d_unres) { k_flag_if<<<1, 1, 0, rb3b_stream>>>(d_unres, bad.p); ++rb3b_n_launch; }
			BmSrc src;
			src.R.cells = A->cells; src.R.n = A->n; src.R.pos = 0; src.R.cur = -1; src.R.rem = 0;
			src.n = A->n; src.cur = -1;
			rc = rb3b_emit_build_bm(A, src, A->n, len, d_ka, d_bwt, bad.p, true); /* validates the positions while it merges */
			rb3b_toc(T_MERGE);
			cudaEventRecord(A->ready_ev, rb3b_stream);
		} while (0);
		ctx->stream = s1; ctx->bump = 0;
		if (rc == RB3B_OK) A->pending = 1;
		else { cudaStreamSynchronize(ctx->stream2); A->broken = 1; }
		return rc;
	}
	TRY(rb3b_index_wait_i(A)); TRY(rb3b_index_use(A));
	DBuf<int> bad;
	int hbad = 0;
	TRY(bad.alloc(1));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	rb3b_tic(T_MERGE);
	const bool fused_check = A->kind == RB3B_KIND_BM && rb3b_want_bitmap(A->n + len); /* the bitmap -> bitmap kernel validates the positions itself */
	if (d_unres) { k_flag_if<<<1, 1, 0, rb3b_stream>>>(d_unres, bad.p); CKK(); }
	if (!fused_check) {
		k_check_monotone<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, d_ka, A->n, bad.p); CKK();
		CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		if (hbad) return rb3b_fail(RB3B_EINVAL, "interleave positions are not monotone or not all resolved: the batch is not a valid BWT");
	}
	int rc;
	if (A->kind == RB3B_KIND_BM) {
		BmSrc src;
		src.R.cells = A->cells; src.R.n = A->n; src.R.pos = 0; src.R.cur = -1; src.R.rem = 0;
		src.n = A->n; src.cur = -1;
		if (rb3b_want_bitmap(A->n + len)) rc = rb3b_emit_build_bm(A, src, A->n, len, d_ka, d_bwt, bad.p);
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

static int async_buffers(rb3b_index_s *x, int64_t len, const uint8_t *d_bwt, int64_t **ka, uint8_t **bcopy, uint32_t **ka32 = 0);

/* collectives of rb3b_dist.cu (NCCL on the current context's stream) */
int rb3b_all_gather(const void *send, void *recv, size_t bytes_per_rank);
int rb3b_all_reduce_max_i64(void *buf, size_t n);
int rb3b_all_reduce_sum_u32(void *buf, size_t n);
int rb3b_all_to_all_v(const void *send, const int64_t *soff, const int64_t *scnt, void *recv, const int64_t *roff, const int64_t *rcnt);
int rb3b_dist_second_comm(void);

/* rb3_fmi_merge_plain on the ranks of the current communicator (rb3b_dist_init): every rank holds a replica of the index
 * and calls this with the same batch; rank r resolves the slices [r, r+1) * n_slices / world of walk order (plus a
 * speculative halo before them, so that its first own slice needs no value from its neighbour), the interleave positions
 * are combined over NVLink, and every rank applies the same streaming merge to its replica.  If some rank could not
 * resolve all of its rows locally (an exact match longer than the halo across a boundary) all ranks recompute the whole
 * array -- still bit-exact.  Returns 0, or 1 when that fallback was taken. */
extern "C" int rb3b_merge_plain_dist_dev(rb3b_index_t *x, int64_t len, const uint8_t *d_bwt)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	rb3b_ctx_s *ctx = rb3b_cur();
	if (ctx->world <= 1) return rb3b_merge_plain_dev(x, len, d_bwt);
	if (len <= 0) return RB3B_OK;
	if (x->n_cells == 0) return rb3b_index_from_plain_dev(x, len, d_bwt);
	DBuf<int64_t> ka, flag;
	int64_t accB[RB3B_ASIZE + 1], hflag = 0, *aka;
	uint8_t *bcopy;
	int incomplete = 0, fell_back = 0;
	const int W = ctx->world;
	/* the exchange: every device owns a range of `chunk` rows of the interleave array; (row, position) pairs go to the
	 * owner of the row (one all-to-all, 12 bytes per row), the dense ranges are all-gathered (8 bytes per row).  Batches
	 * with 64-bit rows, or more ranks than MAX_RANKS, use the all-reduce(MAX) of partial arrays instead. */
	const bool by_pairs = len < LF32_MAX_LEN && !rb3b_get_param("wide_lf", 0) && W <= MAX_RANKS && rb3b_get_param("dist_pairs", 0) != 0; /* off by default: measured slower than the NVLS all-reduce (N=4: 1.40 vs 0.91 + 0.40 ms scatter) */
	const int64_t chunk = ((len + W - 1) / W + 63) / 64 * 64;
	uint32_t *aka32 = 0;
	if (rb3b_get_param("dist_async", 0) != 0) TRY(rb3b_dist_second_comm()); /* collective, first time only */
	TRY(async_buffers(x, len, d_bwt, &aka, &bcopy, &aka32));
	if (aka && (!by_pairs || chunk * W <= x->ms_rows)) { ka.p = aka; d_bwt = bcopy; } else { aka = 0; TRY(ka.alloc(by_pairs ? chunk * W : len)); }
	TRY(flag.alloc(2 + 2 * MAX_RANKS + (size_t)W * W));
	OwnPairs own;
	int64_t *const ka_full = ka.p;
	/* positions below 2^32 (every BASELINE config 1 index): the partial arrays are 32-bit with 0 for "not mine" and are
	 * combined by a SUM all-reduce -- half the bytes of the 64-bit MAX all-reduce */
	const bool narrow_ka = !by_pairs && x->n + len < (1LL << 32) && rb3b_get_param("dist_ka32", 1) != 0;
	DBuf<uint32_t> ka32;
	if (narrow_ka) { if (aka) ka32.p = aka32; else TRY(ka32.alloc(len)); }
	TRY(rank_phase(x, len, d_bwt, ka, accB, ctx->rank, ctx->world, ka.p, &incomplete, 0, by_pairs ? &own : 0, 0, narrow_ka ? ka32.p : 0));
	ka.p = ka_full; /* with pair / 32-bit output the rank phase leaves the array alone */
	hflag = incomplete;
	/* asynchronous exchange + merge: the partial arrays live in the index's own scratch, so the all-reduce and the merge are
	 * only QUEUED on the second stream (second communicator) and the caller's next batch is prepared meanwhile */
	const bool comm_on_2 = aka != 0 && !by_pairs && ctx->comm2 != 0;
	if (!comm_on_2) rb3b_tic(T_COMM);
	unsigned long long *d_cnt = (unsigned long long*)(flag.p + 2), *d_cur = d_cnt + MAX_RANKS;
	int64_t *d_all = flag.p + 2 + 2 * MAX_RANKS;
	CK(cudaMemcpyAsync(flag.p, &hflag, 8, cudaMemcpyHostToDevice, rb3b_stream));
	if (by_pairs) {
		CK(cudaMemsetAsync(d_cnt, 0, 2 * MAX_RANKS * 8, rb3b_stream));
		if (own.n > 0) { k_dest_count<<<nblk(own.n, TPB), TPB, 0, rb3b_stream>>>(own.n, own.rows, chunk, W, d_cnt); CKK(); }
		TRY(rb3b_all_gather(d_cnt, d_all, (size_t)W * 8)); /* row r of d_all: how many pairs rank r has for every destination */
	}
	TRY(rb3b_all_reduce_max_i64(flag.p, 1));
	std::vector<int64_t> hall((size_t)W * W + 1, 0);
	CK(cudaMemcpyAsync(&hflag, flag.p, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	if (by_pairs) CK(cudaMemcpyAsync(hall.data(), d_all, (size_t)W * W * 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (hflag) { /* rare: every rank computes everything */
		DBuf<int64_t> ka2;
		TRY(rank_phase(x, len, d_bwt, ka2, accB, 0, 1, ka.p));
		fell_back = 1;
	} else if (by_pairs) {
		int64_t soff[MAX_RANKS], scnt[MAX_RANKS], roff[MAX_RANKS], rcnt[MAX_RANKS], n_recv = 0, n_send = 0, b4[MAX_RANKS], c4[MAX_RANKS], b8[MAX_RANKS], c8[MAX_RANKS];
		for (int p = 0; p < W; ++p) { scnt[p] = hall[(size_t)ctx->rank * W + p]; soff[p] = n_send; n_send += scnt[p]; rcnt[p] = hall[(size_t)p * W + ctx->rank]; roff[p] = n_recv; n_recv += rcnt[p]; }
		DBuf<int64_t> d_base, pval, rval;
		DBuf<uint32_t> prow, rrow;
		TRY(d_base.alloc(MAX_RANKS)); TRY(pval.alloc(n_send)); TRY(prow.alloc(n_send)); TRY(rval.alloc(n_recv)); TRY(rrow.alloc(n_recv));
		CK(cudaMemcpyAsync(d_base.p, soff, (size_t)W * 8, cudaMemcpyHostToDevice, rb3b_stream));
		if (own.n > 0) { k_dest_pack<<<nblk(own.n, TPB), TPB, 0, rb3b_stream>>>(own.n, own.rows, own.vals, chunk, W, d_base.p, d_cur, prow.p, pval.p); CKK(); }
		for (int p = 0; p < W; ++p) { b4[p] = soff[p] * 4; c4[p] = scnt[p] * 4; b8[p] = soff[p] * 8; c8[p] = scnt[p] * 8; }
		int64_t rb4[MAX_RANKS], rc4[MAX_RANKS], rb8[MAX_RANKS], rc8[MAX_RANKS];
		for (int p = 0; p < W; ++p) { rb4[p] = roff[p] * 4; rc4[p] = rcnt[p] * 4; rb8[p] = roff[p] * 8; rc8[p] = rcnt[p] * 8; }
		TRY(rb3b_all_to_all_v(prow.p, b4, c4, rrow.p, rb4, rc4));
		TRY(rb3b_all_to_all_v(pval.p, b8, c8, rval.p, rb8, rc8));
		const int64_t row0 = chunk * ctx->rank, mine = len - row0 < chunk ? (len - row0 > 0 ? len - row0 : 0) : chunk;
		if (n_recv != mine) return rb3b_fail(RB3B_EINVAL, "internal error: device %d received %lld pairs for a range of %lld rows", ctx->rank, (long long)n_recv, (long long)mine);
		if (n_recv > 0) { k_fill_dense<<<nblk(n_recv, TPB), TPB, 0, rb3b_stream>>>(n_recv, rrow.p, rval.p, row0, ka.p + row0); CKK(); }
		TRY(rb3b_all_gather(ka.p + row0, ka.p, (size_t)chunk * 8));
		CK(cudaStreamSynchronize(rb3b_stream)); /* the packed buffers are scratch of this call */
	} else {
		cudaStream_t s1 = ctx->stream;
		int rc = RB3B_OK;
		if (comm_on_2) {
			CK(cudaEventRecord(ctx->ev_hand, s1)); /* the second stream goes on where the rank phase stops */
			CK(cudaStreamWaitEvent(ctx->stream2, ctx->ev_hand, 0));
			ctx->stream = ctx->stream2;
			rb3b_tic(T_COMM);
		}
		if (narrow_ka) {
			rc = rb3b_all_reduce_sum_u32(ka32.p, (size_t)len);
			if (rc == RB3B_OK) { k_widen_ka<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, ka32.p, ka.p); ++rb3b_n_launch; }
		} else rc = rb3b_all_reduce_max_i64(ka.p, (size_t)len); /* rows another rank resolved are -1 here */
		if (comm_on_2) { rb3b_toc(T_COMM); ctx->stream = s1; }
		if (rc != RB3B_OK) return rc;
		if (cudaGetLastError() != cudaSuccess) return rb3b_fail(RB3B_ENODEV, "launching the exchange failed");
	}
	if (!comm_on_2) rb3b_toc(T_COMM);
	if (aka) TRY(merge_phase(x, len, bcopy, aka, accB));
	else TRY(merge_phase(x, len, d_bwt, ka.p));
	rb3b_stat_add("dist_fallbacks", fell_back);
	return fell_back;
}

/* same with the batch in host memory: every rank copies only its 1/world share over PCIe and the shares are exchanged over
 * NVLink (bwt: the whole batch on every rank, e.g. one pinned buffer shared by the ranks of a process) */
extern "C" int rb3b_merge_plain_dist(rb3b_index_t *x, int64_t len, const uint8_t *bwt)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	rb3b_ctx_s *ctx = rb3b_cur();
	if (ctx->world <= 1) return rb3b_merge_plain(x, len, bwt);
	if (len <= 0) return RB3B_OK;
	const int64_t chunk = ((len + ctx->world - 1) / ctx->world + 15) / 16 * 16, lo = chunk * ctx->rank;
	DBuf<uint8_t> d;
	TRY(d.alloc((size_t)chunk * ctx->world));
	if (lo < len) CK(cudaMemcpyAsync(d.p + lo, bwt + lo, (size_t)(len - lo < chunk ? len - lo : chunk), cudaMemcpyHostToDevice, rb3b_stream));
	TRY(rb3b_all_gather(d.p + lo, d.p, (size_t)chunk));
	int rc = rb3b_merge_plain_dist_dev(x, len, d.p);
	if (rc < 0) return rc;
	CK(cudaStreamSynchronize(rb3b_stream));
	return rc;
}

extern "C" int rb3b_merge_with_ka(rb3b_index_t *x, int64_t len, const uint8_t *d_bwt, const int64_t *d_ka)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (len <= 0) return RB3B_OK;
	if (x->n_cells == 0) return rb3b_fail(RB3B_EINVAL, "merge_with_ka on an empty index");
	return merge_phase(x, len, d_bwt, d_ka); /* rejects arrays with holes (-1) or out of order */
}

/* interleave positions and a copy of the batch in the index's own scratch, so that the merge can outlive the call (and the
 * caller may reuse its batch buffer as soon as the call returns).  Two batch copies alternate: the merge in flight reads
 * the other one, so this one can be filled right away; the interleave array is written only after the caller's rank phase
 * has waited for that merge (rb3b_index_use). */
/* ka32 != 0 (multi-device, "dist_async"): also a 32-bit partial array -- the all-reduce of the partial arrays then runs on
 * the second stream as well, see rb3b_merge_plain_dist_dev.  Off by default as well: at N = 2 the exchange + merge of batch i
 * do overlap the preparation of batch i + 1, but both phases slow down by what they overlap (prep 0.80 -> 1.17 ms, merge
 * 0.51 -> 0.83 ms, step 2.50 -> 2.53 ms; gpurun_out/r2_dist_async*_n2.json): the list ranking and the streaming merge
 * compete for the same memory system. */
static int async_buffers(rb3b_index_s *x, int64_t len, const uint8_t *d_bwt, int64_t **ka, uint8_t **bcopy, uint32_t **ka32)
{
	*ka = 0; *bcopy = 0;
	if (ka32) *ka32 = 0;
	const bool on = ka32 ? rb3b_get_param("dist_async", 0) != 0 : rb3b_get_param("async_merge", 0) != 0; /* single device: off by default, measured gain 3 % (1.573 vs 1.615 ms per merge, tools/async_probe.py) */
	if (x->kind != RB3B_KIND_BM || !rb3b_want_bitmap(x->n + len) || !on) return RB3B_OK;
	const int64_t n_cells = (x->n + len + 127) >> RB3B_BM_SHIFT;
	if (len > x->ms_rows) { /* a new layout: nothing may be in flight in the old one */
		TRY(rb3b_index_wait_i(x));
		x->ms_rows = len + len / 4;
	}
	const size_t ka_b = al512((size_t)x->ms_rows * 8) + al512((size_t)x->ms_rows * 4) /* one layout for both uses */, bw_b = al512((size_t)x->ms_rows);
	const size_t tables = (size_t)(n_cells + 1) * 8 + (size_t)n_cells * 16 + 4 * (size_t)(n_cells / EMIT_TPB + 2) * RB3B_ASIZE * 8 + ((size_t)8 << 20);
	TRY(ms_reserve(x, ka_b + 2 * bw_b + tables));
	x->ms_used = ka_b + 2 * bw_b;
	*ka = (int64_t*)x->ms; *bcopy = (uint8_t*)(x->ms + ka_b + (x->ms_flip ? bw_b : 0));
	if (ka32) *ka32 = (uint32_t*)(x->ms + al512((size_t)x->ms_rows * 8));
	x->ms_flip ^= 1;
	CK(cudaMemcpyAsync(*bcopy, d_bwt, (size_t)len, cudaMemcpyDeviceToDevice, rb3b_stream)); /* done when the rank phase returns: it waits for the device */
	return RB3B_OK;
}

extern "C" int rb3b_merge_plain_dev(rb3b_index_t *x, int64_t len, const uint8_t *d_bwt)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (len <= 0) return RB3B_OK;
	if (x->n_cells == 0) return rb3b_index_from_plain_dev(x, len, d_bwt);
	DBuf<int64_t> ka;
	int64_t accB[RB3B_ASIZE + 1], *aka;
	uint8_t *bcopy;
	TRY(async_buffers(x, len, d_bwt, &aka, &bcopy));
	const unsigned long long *d_unres = 0;
	if (aka) {
		TRY(rank_phase(x, len, bcopy, ka, accB, 0, 1, aka, 0, 0, 0, 0, 0, &d_unres));
		return merge_phase(x, len, bcopy, aka, accB, d_unres);
	}
	TRY(rank_phase(x, len, d_bwt, ka, accB, 0, 1, 0, 0, 0, 0, 0, 0, &d_unres));
	return merge_phase(x, len, d_bwt, ka.p, 0, d_unres);
}

/* rb3_fmi_merge_plain / rb3_enc_plain2fmr on a batch prepared by rb3b_batch_prepare* (rb3b_bwt.cu): the walk order comes with
 * the batch, so only the walk over the index, the fix-up, the scatter and the merge are left */
extern "C" int rb3b_merge_prepared(rb3b_index_t *x, const rb3b_batch_t *b)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (b == 0 || b->len <= 0) return rb3b_fail(RB3B_EINVAL, "empty batch");
	if (x->n_cells == 0) return rb3b_index_from_plain_dev(x, b->len, b->bwt);
	if (b->wsym == 0) return rb3b_merge_plain_dev(x, b->len, b->bwt); /* very large batch: only its BWT was prepared */
	DBuf<int64_t> ka;
	int64_t accB[RB3B_ASIZE + 1];
	const unsigned long long *d_unres = 0;
	TRY(rank_phase(x, b->len, b->bwt, ka, accB, 0, 1, 0, 0, 0, 0, b, 0, &d_unres));
	return merge_phase(x, b->len, b->bwt, ka.p, 0, d_unres);
}

/* mr_insert_multi (mrope.c:300-385; build -2/-s/-r, build.c:214-218): insert the strings of `text` (concatenated,
 * 0-terminated, as read -- the reference reverses them first, rb3_reverse_all) in the index's sorting order.  The BCR
 * rounds are replaced by one device suffix sort of the batch in that order and one merge whose sentinel rows are placed
 * by k_so_heads.  In input order (so == 0) this is exactly build + merge_plain. */
extern "C" int rb3b_insert_multi_dev(rb3b_index_t *x, int64_t len, const uint8_t *d_text)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (len <= 0) return RB3B_OK;
	DBuf<uint8_t> bwt;
	TRY(bwt.alloc(len));
	TRY(rb3b_build_bwt_so_dev(len, d_text, x->so, bwt.p));
	if (x->n_cells == 0) return rb3b_index_from_plain_dev(x, len, bwt.p);
	DBuf<int64_t> ka;
	int64_t accB[RB3B_ASIZE + 1];
	TRY(rank_phase(x, len, bwt.p, ka, accB, 0, 1, 0, 0, x->so));
	return merge_phase(x, len, bwt.p, ka.p);
}

extern "C" int rb3b_insert_multi(rb3b_index_t *x, int64_t len, const uint8_t *text)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> d;
	if (len <= 0) return RB3B_OK;
	TRY(d.alloc(len));
	CK(cudaMemcpyAsync(d.p, text, len, cudaMemcpyHostToDevice, rb3b_stream));
	TRY(rb3b_insert_multi_dev(x, len, d.p));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}

extern "C" int rb3b_merge_plain(rb3b_index_t *x, int64_t len, const uint8_t *bwt)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> d;
	if (len <= 0) return RB3B_OK;
	if ((d.p = rb3b_prefetched(bwt, len)) == 0) { /* not copied ahead (rb3b_prefetch_batch) */
		TRY(d.alloc(len));
		CK(cudaMemcpyAsync(d.p, bwt, len, cudaMemcpyHostToDevice, rb3b_stream));
	}
	TRY(rb3b_merge_plain_dev(x, len, d.p));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}
