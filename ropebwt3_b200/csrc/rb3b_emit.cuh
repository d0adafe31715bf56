/*
 * rb3b_emit.cuh -- the cell writer shared by index construction (runs -> cells)
 * and by the streaming merge (cells + batch rows -> cells).
 *
 * One thread produces one OUTPUT cell = 2^shift consecutive positions of the
 * result.  Output position P holds batch row i when ka[i] + i == P (the rows'
 * merged positions are strictly increasing), otherwise the next symbol of the
 * source stream.  Because every tile has the same span the work is balanced
 * and every cell lands at an arithmetic address: no global entry offsets.
 *
 *   pass 1 (WRITE = false)  count the entries of every cell -> nent[], per-chunk
 *                           symbol totals and overflow-block totals
 *   (host)                  one small exclusive scan over the chunk totals
 *   pass 2 (WRITE = true)   redo the merge, write entries (inline or into the
 *                           cell's overflow blocks), then the headers from a
 *                           CTA-level scan + the chunk bases
 *
 * Replaces worker_mgins / rope_insert_run / rle_insert_cached (fm-index.c:237-249,
 * rope.c:114-148, rle.c:10-89) and worker_p2fmr (fm-index.c:97-112).
 */
#ifndef RB3B_EMIT_CUH
#define RB3B_EMIT_CUH

#include <cub/cub.cuh>
#include "rb3b_internal.cuh"

#define EMIT_TPB 128

struct EmitOut {
	uint4 *cells, *ovf;
	int shift;
	int64_t n_out, n_cells, n_chunks;
};

/* source stream given as a run list with start positions (start[] = exclusive scan of the lengths) */
struct RunSrc {
	const uint8_t *sym; const int64_t *start;
	int64_t n_runs, n, r, rem64;
	int cur;
	__device__ __forceinline__ int64_t run_end(int64_t i) const { return i + 1 < n_runs ? start[i + 1] : n; }
	__device__ __forceinline__ void seek(int64_t pos)
	{
		int64_t lo = 0, hi = n_runs; /* last run whose start is <= pos; with equal starts that is the non-empty one */
		while (hi - lo > 1) {
			int64_t mid = (lo + hi) >> 1;
			if (start[mid] <= pos) lo = mid; else hi = mid;
		}
		r = lo; cur = sym[r]; rem64 = run_end(r) - pos;
	}
	__device__ __forceinline__ uint32_t avail(int64_t lim) const { return (uint32_t)(rem64 < lim ? rem64 : lim); }
	__device__ __forceinline__ void advance(uint32_t l)
	{
		rem64 -= l;
		while (rem64 == 0 && r + 1 < n_runs) { ++r; cur = sym[r]; rem64 = run_end(r) - start[r]; }
	}
};

/* source stream given as an index */
struct CellSrc {
	CellReader R;
	int64_t n;
	int cur;
	__device__ __forceinline__ void seek(int64_t pos) { R.seek(pos); cur = R.cur; }
	__device__ __forceinline__ uint32_t avail(int64_t lim) const { return (uint32_t)((int64_t)R.rem < lim ? (int64_t)R.rem : lim); }
	__device__ __forceinline__ void advance(uint32_t l) { R.advance(l); cur = R.cur; }
};

/* source stream given as a bitmap index */
struct BmSrc {
	BmReader R;
	int64_t n;
	int cur;
	__device__ __forceinline__ void seek(int64_t pos) { R.seek(pos); cur = R.cur; }
	__device__ __forceinline__ uint32_t avail(int64_t lim) const { return (uint32_t)((int64_t)R.rem < lim ? (int64_t)R.rem : lim); }
	__device__ __forceinline__ void advance(uint32_t l) { R.advance(l); cur = R.cur; }
};

template<bool WRITE> struct CellEmitter {
	int sym; uint32_t len, ne, pos;
	uint32_t cc[RB3B_ASIZE];
	bool is_ovf;
	uint4 *cell, *blk; /* this cell; its first overflow block */
	__device__ __forceinline__ void init(uint4 *cell_, uint4 *blk_, bool is_ovf_)
	{
		sym = -1; len = ne = pos = 0; cell = cell_; blk = blk_; is_ovf = is_ovf_;
#pragma unroll
		for (int a = 0; a < RB3B_ASIZE; ++a) cc[a] = 0;
	}
	__device__ __forceinline__ void flush()
	{
		while (len > RB3B_LEN_MASK) { /* a 13-bit length field: long runs (wide cells only) take several entries */
			const uint32_t rest = len - RB3B_LEN_MASK;
			len = RB3B_LEN_MASK;
			flush1();
			len = rest;
		}
		flush1();
	}
	__device__ __forceinline__ void flush1()
	{
		if (len == 0) return;
		if (WRITE) {
			uint16_t e = (uint16_t)((uint32_t)sym << 13 | len);
			if (!is_ovf) ((uint16_t*)(cell + 2))[ne] = e;
			else {
				uint32_t t = ne / RB3B_ENT_PER_OVF, s = ne % RB3B_ENT_PER_OVF;
				if (s == 0) { /* a new overflow block: its header and its slot in the cell's directory */
					uint16_t *h = (uint16_t*)(blk + (int64_t)t * 8);
#pragma unroll
					for (int a = 0; a < RB3B_ASIZE; ++a) h[a] = (uint16_t)cc[a];
					h[6] = h[7] = 0;
					if (t > 0) ((uint16_t*)(cell + 3))[t - 1] = (uint16_t)pos;
				}
				((uint16_t*)(blk + (int64_t)t * 8 + 1))[s] = e;
			}
		}
		++ne; pos += len;
#pragma unroll
		for (int a = 0; a < RB3B_ASIZE; ++a) cc[a] += sym == a ? len : 0;
		len = 0;
	}
	__device__ __forceinline__ void put(int c, uint32_t l)
	{
		if (c == sym) len += l;
		else { flush(); sym = c; len = l; }
	}
};

__host__ __device__ __forceinline__ uint32_t rb3b_ovf_blocks(uint32_t nent)
{ return nent > RB3B_ENT_PER_CELL ? (nent + RB3B_ENT_PER_OVF - 1) / RB3B_ENT_PER_OVF : 0; }

/* ctot / cex layout: [7][n_chunks + 1]; rows 0-5 symbol totals of a chunk of EMIT_TPB cells, row 6 overflow blocks */
template<class Src, bool WRITE>
__global__ void __launch_bounds__(EMIT_TPB) k_emit(Src src, EmitOut O, int64_t lenB, const int64_t *__restrict__ ka, const uint8_t *__restrict__ bwt,
                                                    const int64_t *__restrict__ ilo, uint16_t *__restrict__ nent, int64_t *__restrict__ ctot,
                                                    const int64_t *__restrict__ cex, unsigned long long *__restrict__ stats)
{
	typedef cub::BlockReduce<int64_t, EMIT_TPB> Red;
	typedef cub::BlockScan<int64_t, EMIT_TPB> Scan;
	__shared__ union { typename Red::TempStorage r; typename Scan::TempStorage s; } tmp[RB3B_ASIZE + 1];
	const int64_t j = (int64_t)blockIdx.x * EMIT_TPB + threadIdx.x;
	const bool live = j < O.n_cells;
	CellEmitter<WRITE> E;
	int64_t novf = 0, ovf_base = 0;
	if (WRITE) {
		novf = live ? rb3b_ovf_blocks(nent[j]) : 0;
		Scan(tmp[RB3B_ASIZE].s).ExclusiveSum(novf, ovf_base);
		ovf_base += cex[(int64_t)RB3B_ASIZE * (O.n_chunks + 1) + blockIdx.x] - cex[(int64_t)RB3B_ASIZE * (O.n_chunks + 1)];
	}
	E.init(O.cells + j * 8, O.ovf + ovf_base * 8, novf > 0);
	if (live) {
		const int64_t P0 = j << O.shift, P1 = min((long long)O.n_out, (long long)(P0 + (1LL << O.shift)));
		int64_t i = ilo ? ilo[j] : 0, iend = ilo ? ilo[j + 1] : 0, P = P0;
		int64_t nextB = i < iend ? ka[i] + i : INT64_MAX;
		if (P0 - i < src.n) src.seek(P0 - i);
		while (P < P1) {
			if (nextB == P) {
				E.put(bwt[i], 1);
				++i; ++P;
				nextB = i < iend ? ka[i] + i : INT64_MAX;
				continue;
			}
			uint32_t t = src.avail((nextB < P1 ? nextB : P1) - P);
			E.put(src.cur, t);
			src.advance(t);
			P += t;
		}
		E.flush();
	}
	if (!WRITE) {
		if (live) nent[j] = (uint16_t)E.ne;
		int64_t ov = live ? rb3b_ovf_blocks(E.ne) : 0;
#pragma unroll
		for (int a = 0; a <= RB3B_ASIZE; ++a) {
			int64_t t = Red(tmp[a].r).Sum(a < RB3B_ASIZE ? (int64_t)E.cc[a < RB3B_ASIZE ? a : 0] : ov);
			if (threadIdx.x == 0) ctot[(int64_t)a * (O.n_chunks + 1) + blockIdx.x] = t;
		}
		if (blockIdx.x == 0 && threadIdx.x <= RB3B_ASIZE) ctot[(int64_t)threadIdx.x * (O.n_chunks + 1) + O.n_chunks] = 0;
		if (live) {
			atomicAdd(&stats[0], (unsigned long long)E.ne);
			if (ov) atomicAdd(&stats[1], 1ULL);
			if (E.ne > RB3B_ENT_PER_CELL) atomicMax(&stats[2], (unsigned long long)E.ne);
		}
	} else {
		uint64_t h[RB3B_ASIZE];
#pragma unroll
		for (int a = 0; a < RB3B_ASIZE; ++a) {
			int64_t ex;
			Scan(tmp[a].s).ExclusiveSum((int64_t)E.cc[a], ex);
			h[a] = (uint64_t)(ex + cex[(int64_t)a * (O.n_chunks + 1) + blockIdx.x] - cex[(int64_t)a * (O.n_chunks + 1)]);
		}
		if (live) {
			uint4 *cell = O.cells + j * 8;
			cell[0] = rb3b_hdr_pack(h[0], h[1], h[2], novf > 0);
			cell[1] = rb3b_hdr_pack(h[3], h[4], h[5], novf > 0); /* the flag is repeated here: a reader of G,T,N needs only this quad */
			if (novf > 0) {
				cell[2] = make_uint4((uint32_t)ovf_base, (uint32_t)novf, 0u, 0u);
				for (uint32_t t = (uint32_t)novf; t < RB3B_MAX_OVF; ++t) ((uint16_t*)(cell + 3))[t - 1] = 0xffffu;
			} else
				for (uint32_t e = E.ne; e < RB3B_ENT_PER_CELL; ++e) ((uint16_t*)(cell + 2))[e] = 0;
		}
	}
}

/* ilo[j] = first batch row whose merged position ka[i] + i is >= j << shift (j = 0..n_cells).  The rows of a batch are
 * spread almost evenly over the merged sequence, so the search gallops out from an interpolated guess. */
static __global__ void k_tile_bounds(int64_t n_cells, int shift, int64_t len, int64_t n_out, const int64_t *__restrict__ ka, int64_t *__restrict__ ilo)
{
	int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j > n_cells) return;
	const int64_t key = j << shift;
	int64_t g = (int64_t)((double)key * (double)len / (double)(n_out > 0 ? n_out : 1));
	g = g < 0 ? 0 : g > len ? len : g;
	int64_t lo, hi, step = 16; /* invariant: every row < lo is below the key, every row >= hi is not */
	if (g < len && ka[g] + g < key) {
		lo = g + 1; hi = len;
		while (lo + step < len) { if (ka[lo + step - 1] + lo + step - 1 < key) { lo += step; step <<= 1; } else { hi = lo + step - 1; break; } }
	} else {
		hi = g; lo = 0;
		while (hi - step > 0) { if (ka[hi - step] + hi - step < key) { lo = hi - step + 1; break; } else { hi -= step; step <<= 1; } }
	}
	while (lo < hi) {
		int64_t mid = (lo + hi) >> 1;
		if (ka[mid] + mid < key) lo = mid + 1; else hi = mid;
	}
	ilo[j] = lo;
}

/*
 * Host driver: write the stream `src` (n_src symbols) interleaved with the batch rows (lenB, ka, bwt; lenB may be 0)
 * into the spare buffers of x, then make them current.  est_entries steers the cell span.
 */
template<class Src>
static int rb3b_emit_build(rb3b_index_s *x, Src src, int64_t n_src, int64_t lenB, const int64_t *d_ka, const uint8_t *d_bwt, int64_t est_entries)
{
	const int64_t n_out = n_src + lenB;
	int shift = rb3b_pick_shift(n_out, est_entries);
	DBuf<int64_t> ilo, ctot, cex;
	DBuf<uint16_t> nent;
	DBuf<unsigned long long> stats;
	unsigned long long hstats[3] = {0, 0, 0};
	EmitOut O;
	TRY(stats.alloc(3));
	for (;;) {
		O.shift = shift; O.n_out = n_out;
		O.n_cells = (n_out + (1LL << shift) - 1) >> shift;
		O.n_chunks = (O.n_cells + EMIT_TPB - 1) / EMIT_TPB;
		O.cells = 0; O.ovf = 0;
		if (O.n_cells >= (1LL << 40)) return rb3b_fail(RB3B_EINVAL, "index too large");
		TRY(nent.alloc(O.n_cells)); TRY(ctot.alloc((O.n_chunks + 1) * (RB3B_ASIZE + 1))); TRY(cex.alloc((O.n_chunks + 1) * (RB3B_ASIZE + 1)));
		if (lenB > 0) {
			TRY(ilo.alloc(O.n_cells + 1));
			k_tile_bounds<<<(unsigned)((O.n_cells + 1 + 255) / 256), 256, 0, rb3b_stream>>>(O.n_cells, shift, lenB, n_out, d_ka, ilo.p); CKK();
		}
		CK(cudaMemsetAsync(stats.p, 0, 24, rb3b_stream));
		k_emit<Src, false><<<(unsigned)O.n_chunks, EMIT_TPB, 0, rb3b_stream>>>(src, O, lenB, d_ka, d_bwt, lenB > 0 ? ilo.p : 0, nent.p, ctot.p, 0, stats.p); CKK();
		CK(cudaMemcpyAsync(hstats, stats.p, 24, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		/* too many overflow cells: halve the span and count again (cheap, rare) */
		if (shift > RB3B_MIN_SHIFT && (double)hstats[1] > 0.03 * (double)O.n_cells) { --shift; continue; }
		/* the densest cell must fit its overflow blocks (an in-cell directory of RB3B_MAX_OVF block starts): a locally dense
		 * stretch forces the span of the WHOLE index down -- the price of arithmetic cell addresses */
		if (shift > RB3B_MIN_SHIFT && hstats[2] > (unsigned long long)RB3B_MAX_OVF * RB3B_ENT_PER_OVF) { --shift; continue; }
		break;
	}
	TRY(rb3b_scan_excl_i64(ctot.p, cex.p, (O.n_chunks + 1) * (RB3B_ASIZE + 1)));
	int64_t tot[RB3B_ASIZE + 1], base[RB3B_ASIZE + 1];
	for (int a = 0; a <= RB3B_ASIZE; ++a) {
		CK(cudaMemcpyAsync(&tot[a], cex.p + a * (O.n_chunks + 1) + O.n_chunks, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaMemcpyAsync(&base[a], cex.p + a * (O.n_chunks + 1), 8, cudaMemcpyDeviceToHost, rb3b_stream));
	}
	CK(cudaStreamSynchronize(rb3b_stream));
	int64_t n_ovf = tot[RB3B_ASIZE] - base[RB3B_ASIZE];
	if (n_ovf >= (1LL << 32)) return rb3b_fail(RB3B_EINVAL, "too many overflow blocks");
	TRY(rb3b_reserve((void**)&x->cells2, &x->cap_cells2, O.n_cells * 8, sizeof(uint4)));
	TRY(rb3b_reserve((void**)&x->ovf2, &x->cap_ovf2, n_ovf * 8 + 8, sizeof(uint4)));
	O.cells = x->cells2; O.ovf = x->ovf2;
	k_emit<Src, true><<<(unsigned)O.n_chunks, EMIT_TPB, 0, rb3b_stream>>>(src, O, lenB, d_ka, d_bwt, lenB > 0 ? ilo.p : 0, nent.p, 0, cex.p, stats.p); CKK();
	/* swap the ping-pong halves (the kernels that read the old half are already enqueued on the same stream) */
	{ uint4 *t = x->cells; x->cells = x->cells2; x->cells2 = t; int64_t c = x->cap_cells; x->cap_cells = x->cap_cells2; x->cap_cells2 = c; }
	{ uint4 *t = x->ovf; x->ovf = x->ovf2; x->ovf2 = t; int64_t c = x->cap_ovf; x->cap_ovf = x->cap_ovf2; x->cap_ovf2 = c; }
	x->kind = RB3B_KIND_RLE; x->shift = shift; x->n_cells = O.n_cells; x->n_ovf = n_ovf; x->n_entries = (int64_t)hstats[0];
	rb3b_stat_set("index_kind", RB3B_KIND_RLE);
	x->acc[0] = 0;
	for (int a = 0; a < RB3B_ASIZE; ++a) { x->tot[a] = tot[a] - base[a]; x->acc[a + 1] = x->acc[a] + x->tot[a]; }
	x->n = x->acc[RB3B_ASIZE];
	x->bytes = (size_t)(O.n_cells + n_ovf) * 128;
	rb3b_stat_set("n_cells", O.n_cells); rb3b_stat_set("n_ovf_blocks", n_ovf); rb3b_stat_set("n_ovf_cells", (int64_t)hstats[1]);
	rb3b_stat_set("cell_shift", shift); rb3b_stat_set("n_entries", (int64_t)hstats[0]);
	if (x->n != n_out) return rb3b_fail(RB3B_EINVAL, "internal error: wrote %lld symbols, expected %lld", (long long)x->n, (long long)n_out);
	return RB3B_OK;
}


/* ------------------------------------------------------------------ */
/* bitmap output                                                        */
/* ------------------------------------------------------------------ */

/* One thread per output cell of 128 positions.  The planes are assembled in registers-backed local memory and
 * written once; quad 0 temporarily receives the cell's own six symbol counts (6 x u16), which k_bm_fin_* turn into
 * the absolute header counts with a two-level scan. */
/* common tail of the bitmap emit kernels: write the planes, the compact per-cell counts and the chunk totals */
template<bool STORE = true>
__device__ __forceinline__ void rb3b_bm_finish_cell(const EmitOut &O, int64_t j, bool live, uint32_t (&pl)[RB3B_ASIZE][4], uint4 *__restrict__ lcnt, int64_t *__restrict__ ctot)
{
	typedef cub::BlockReduce<uint32_t, EMIT_TPB> Red;
	__shared__ typename Red::TempStorage tmp[RB3B_ASIZE];
	uint32_t c[RB3B_ASIZE];
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		c[a] = live ? __popc(pl[a][0]) + __popc(pl[a][1]) + __popc(pl[a][2]) + __popc(pl[a][3]) : 0u;
		if (STORE && live) O.cells[j * 8 + rb3b_bm_plane_quad(a)] = make_uint4(pl[a][0], pl[a][1], pl[a][2], pl[a][3]);
	}
	if (live) lcnt[j] = make_uint4(c[0] | c[1] << 16, c[2] | c[3] << 16, c[4] | c[5] << 16, 0u);
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		uint32_t t = Red(tmp[a]).Sum(c[a]);
		if (threadIdx.x == 0) ctot[(int64_t)a * (O.n_chunks + 1) + blockIdx.x] = t;
	}
	if (blockIdx.x == 0 && threadIdx.x < RB3B_ASIZE) ctot[(int64_t)threadIdx.x * (O.n_chunks + 1) + O.n_chunks] = 0;
}

template<class Src>
__global__ void __launch_bounds__(EMIT_TPB) k_emit_bm(Src src, EmitOut O, int64_t lenB, const int64_t *__restrict__ ka, const uint8_t *__restrict__ bwt,
                                                       const int64_t *__restrict__ ilo, uint4 *__restrict__ lcnt, int64_t *__restrict__ ctot)
{
	const int64_t j = (int64_t)blockIdx.x * EMIT_TPB + threadIdx.x;
	const bool live = j < O.n_cells;
	uint32_t pl[RB3B_ASIZE][4];
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a)
#pragma unroll
		for (int w = 0; w < 4; ++w) pl[a][w] = 0;
	const int64_t P0 = j << RB3B_BM_SHIFT, P1 = !live ? P0 : P0 + 128 < O.n_out ? P0 + 128 : O.n_out;
	int64_t i = (ilo && live) ? ilo[j] : 0, iend = (ilo && live) ? ilo[j + 1] : 0, P = P0;
	int64_t nextB = i < iend ? ka[i] + i : INT64_MAX;
	if (live && P0 - i < src.n) src.seek(P0 - i);
	while (P < P1) {
		int sym; uint32_t t;
		if (nextB == P) {
			sym = bwt[i]; t = 1;
			++i;
			nextB = i < iend ? ka[i] + i : INT64_MAX;
		} else {
			t = src.avail((nextB < P1 ? nextB : P1) - P);
			sym = src.cur;
			src.advance(t);
		}
		/* set bits [p, p+t) of plane sym */
		uint32_t p = (uint32_t)(P - P0), e = p + t;
#pragma unroll
		for (int w = 0; w < 4; ++w) {
			int lo = (int)p - 32 * w, hi = (int)e - 32 * w;
			lo = lo < 0 ? 0 : lo; hi = hi > 32 ? 32 : hi;
			if (lo < hi) {
				uint32_t m = (hi == 32 ? 0xffffffffu : (1u << hi) - 1u) & ~((1u << lo) - 1u);
#pragma unroll
				for (int a = 0; a < RB3B_ASIZE; ++a) pl[a][w] |= a == sym ? m : 0u;
			}
		}
		P += t;
	}
	rb3b_bm_finish_cell(O, j, live, pl, lcnt, ctot);
}


/* Bitmap -> bitmap merge without decoding runs.  The A symbols of an output cell are 128 consecutive bits (per plane)
 * of at most two A cells: X = (A planes >> offset).  The batch rows landing in the cell are bits that have to be INSERTED
 * into that stream (a software PDEP): the cell is produced one 32-bit word at a time; a word starts as the next 32 source
 * bits of every plane, every row landing in it (ascending offsets) opens a gap at its bit and sets the bit in its symbol's
 * plane -- six planes x six instructions per row -- and the source stream advances by 32 minus the rows of the word. */
#define EMIT_SRC_CELLS (EMIT_TPB + 4) /* source cells a CTA can touch: its 128 output cells start inside at most 129 consecutive source cells, plus one neighbour */
__device__ __forceinline__ void emit_cp_async16(void *smem, const void *gmem)
{ asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" :: "r"((uint32_t)__cvta_generic_to_shared(smem)), "l"(gmem) : "memory"); }

/* STAGED: the CTA's source cells come in through shared memory with fully coalesced 16-byte cp.async (quad q of cell l is kept
 * at l * 8 + (q ^ (l & 7)), so that the per-thread reads at a 128-byte stride are free of bank conflicts), and the finished
 * cells go out the same way as whole 128-byte lines (header quads zero: k_bm_fin_write fills them in). */
template<bool STAGED>
static __global__ void __launch_bounds__(EMIT_TPB) k_emit_bm_fast(const uint4 *__restrict__ A, int64_t nA, int64_t nA_cells, EmitOut O,
                                                                   const int64_t *__restrict__ ka, const uint8_t *__restrict__ bwt, const int64_t *__restrict__ ilo,
                                                                   uint4 *__restrict__ lcnt, int64_t *__restrict__ ctot, int64_t lenB, int *__restrict__ bad)
{
	const int64_t j = (int64_t)blockIdx.x * EMIT_TPB + threadIdx.x;
	const bool live = j < O.n_cells;
	const int64_t P0 = j << RB3B_BM_SHIFT;
	const int n_out = !live ? 0 : (int)(O.n_out - P0 < 128 ? O.n_out - P0 : 128);
	int64_t i0 = live ? ilo[j] : 0, i1 = live ? ilo[j + 1] : 0, a0 = P0 - i0;
	/* The interleave positions are validated HERE (bad != 0) instead of by a pass of their own: they are valid iff the
	 * cells' row ranges tile [0, lenB) in order and every cell finds its rows at strictly increasing offsets inside its
	 * own 128 positions.  Anything else sets *bad (the caller then keeps the old cells) and is made harmless. */
	bool wrong = live && (i1 < i0 || a0 < 0 || a0 > nA || (j == 0 && i0 != 0) || (j == O.n_cells - 1 && i1 != lenB));
	if (wrong) { i0 = i1 = 0; a0 = 0; }
	int64_t jA = a0 >> RB3B_BM_SHIFT;
	__shared__ __align__(16) uint4 tile[STAGED ? EMIT_SRC_CELLS * 8 : 1];
	__shared__ int64_t s_first;
	int lA = 0; /* STAGED: jA relative to the first source cell of the CTA */
	if (STAGED) {
		if (threadIdx.x == 0) s_first = jA; /* the first cell of a CTA is always live */
		__syncthreads();
		const int64_t first = s_first;
		int64_t cnt = nA_cells - first;
		cnt = cnt < 0 ? 0 : cnt > EMIT_SRC_CELLS ? EMIT_SRC_CELLS : cnt;
		for (int e = threadIdx.x; e < (int)cnt * 8; e += EMIT_TPB) {
			const int l = e >> 3, qq = e & 7;
			emit_cp_async16(&tile[l * 8 + (qq ^ (l & 7))], A + first * 8 + e);
		}
		asm volatile("cp.async.commit_group;" ::: "memory");
		if (live && (jA < first || jA - first > EMIT_SRC_CELLS - 2)) { wrong = true; i0 = i1 = 0; a0 = first << RB3B_BM_SHIFT; jA = first; } /* only with invalid positions */
		lA = (int)(jA - first);
		asm volatile("cp.async.wait_group 0;" ::: "memory");
		__syncthreads();
	}
	const uint32_t q = ((uint32_t)a0 & 127u) >> 5, r = (uint32_t)a0 & 31u;
	uint32_t X[RB3B_ASIZE][4], OUT[RB3B_ASIZE][4];
#pragma unroll
	for (int s = 0; s < RB3B_ASIZE; ++s) {
		uint4 c0 = make_uint4(0, 0, 0, 0), c1 = c0;
		if (STAGED) {
			if (live && jA < nA_cells) c0 = tile[lA * 8 + (rb3b_bm_plane_quad(s) ^ (lA & 7))];
			if (live && jA + 1 < nA_cells) c1 = tile[(lA + 1) * 8 + (rb3b_bm_plane_quad(s) ^ ((lA + 1) & 7))];
		} else {
			if (live && jA < nA_cells) c0 = __ldg(A + jA * 8 + rb3b_bm_plane_quad(s));
			if (live && jA + 1 < nA_cells) c1 = __ldg(A + (jA + 1) * 8 + rb3b_bm_plane_quad(s));
		}
		uint32_t w[9] = { c0.x, c0.y, c0.z, c0.w, c1.x, c1.y, c1.z, c1.w, 0u };
		if (q & 1u) {
#pragma unroll
			for (int i = 0; i < 8; ++i) w[i] = w[i + 1];
		}
		if (q & 2u) {
#pragma unroll
			for (int i = 0; i < 7; ++i) w[i] = i + 2 < 9 ? w[i + 2] : 0u;
		}
#pragma unroll
		for (int i = 0; i < 4; ++i) X[s][i] = __funnelshift_r(w[i], w[i + 1], r);
	}
	/* the rows of this cell, in ascending output offset */
	int64_t t = i0;
	int b = 1 << 30, sym = 0, last_b = -1; /* offset and symbol of the next row (1 << 30: none left) */
	if (t < i1) { const int64_t bq = ka[t] + t - P0; b = bq < 0 || bq > 127 ? 128 : (int)bq; sym = bwt[t]; }
#pragma unroll
	for (int w = 0; w < 4; ++w) {
		uint32_t o[RB3B_ASIZE];
#pragma unroll
		for (int s = 0; s < RB3B_ASIZE; ++s) o[s] = X[s][0];
		int n_ins = 0;
		while (b < 32 * (w + 1)) { /* rows landing in this word */
			const int bb = b & 31;
			const uint32_t lowm = (1u << bb) - 1u, bit = 1u << bb;
#pragma unroll
			for (int s = 0; s < RB3B_ASIZE; ++s) o[s] = (o[s] & lowm) | ((o[s] & ~lowm) << 1) | (s == sym ? bit : 0u);
			++n_ins; ++t;
			wrong |= b <= last_b;
			last_b = b;
			if (t < i1) { const int64_t bq = ka[t] + t - P0; b = bq < 0 || bq > 127 ? 128 : (int)bq; sym = bwt[t]; } else b = 1 << 30;
		}
		/* positions past the end of the index hold nothing */
		const int valid = n_out - 32 * w;
		const uint32_t vm = valid >= 32 ? 0xffffffffu : valid <= 0 ? 0u : (1u << valid) - 1u;
#pragma unroll
		for (int s = 0; s < RB3B_ASIZE; ++s) OUT[s][w] = o[s] & vm;
		/* the source stream advances by the A symbols this word consumed */
		const int adv = 32 - n_ins;
		if (adv >= 32) {
#pragma unroll
			for (int s = 0; s < RB3B_ASIZE; ++s) { X[s][0] = X[s][1]; X[s][1] = X[s][2]; X[s][2] = X[s][3]; X[s][3] = 0u; }
		} else {
#pragma unroll
			for (int s = 0; s < RB3B_ASIZE; ++s) {
				X[s][0] = __funnelshift_r(X[s][0], X[s][1], adv);
				X[s][1] = __funnelshift_r(X[s][1], X[s][2], adv);
				X[s][2] = __funnelshift_r(X[s][2], X[s][3], adv);
				X[s][3] >>= adv;
			}
		}
	}
	wrong |= t < i1; /* a row that does not land in this cell */
	if (wrong && bad) *bad = 1;
	if (STAGED) {
		__syncthreads(); /* everybody has read the source tile: it becomes the output tile */
		const int tl = threadIdx.x;
		tile[tl * 8 + (0 ^ (tl & 7))] = make_uint4(0, 0, 0, 0);
		tile[tl * 8 + (4 ^ (tl & 7))] = make_uint4(0, 0, 0, 0);
#pragma unroll
		for (int s = 0; s < RB3B_ASIZE; ++s) tile[tl * 8 + (rb3b_bm_plane_quad(s) ^ (tl & 7))] = make_uint4(OUT[s][0], OUT[s][1], OUT[s][2], OUT[s][3]);
		__syncthreads();
		const int64_t j0 = (int64_t)blockIdx.x * EMIT_TPB;
		const int64_t left = O.n_cells - j0;
		const int n_here = (int)(left < EMIT_TPB ? left : EMIT_TPB);
		for (int e = threadIdx.x; e < n_here * 8; e += EMIT_TPB) {
			const int l = e >> 3, qq = e & 7;
			O.cells[j0 * 8 + e] = tile[l * 8 + (qq ^ (l & 7))];
		}
		rb3b_bm_finish_cell<false>(O, j, live, OUT, lcnt, ctot);
	} else rb3b_bm_finish_cell<true>(O, j, live, OUT, lcnt, ctot);
}

/* *validated = 1: the kernel checked the interleave positions itself (d_bad is set when they are not monotone) */
template<class Src> static inline bool rb3b_launch_emit_bm_fast(const Src &, const EmitOut &, int64_t, const int64_t *, const uint8_t *, const int64_t *, uint4 *, int64_t *, int *) { return false; }
static inline bool rb3b_launch_emit_bm_fast(const BmSrc &src, const EmitOut &O, int64_t lenB, const int64_t *d_ka, const uint8_t *d_bwt, const int64_t *d_ilo, uint4 *lcnt, int64_t *ctot, int *d_bad)
{
	if (lenB <= 0 || d_ilo == 0) return false;
	if (rb3b_get_param("emit_staged", 1) != 0)
		k_emit_bm_fast<true><<<(unsigned)O.n_chunks, EMIT_TPB, 0, rb3b_stream>>>(src.R.cells, src.n, (src.n + 127) >> RB3B_BM_SHIFT, O, d_ka, d_bwt, d_ilo, lcnt, ctot, lenB, d_bad);
	else
		k_emit_bm_fast<false><<<(unsigned)O.n_chunks, EMIT_TPB, 0, rb3b_stream>>>(src.R.cells, src.n, (src.n + 127) >> RB3B_BM_SHIFT, O, d_ka, d_bwt, d_ilo, lcnt, ctot, lenB, d_bad);
	return true;
}

/* absolute header counts = scan of the per-cell counts inside the chunk + the chunk bases */
static __global__ void __launch_bounds__(EMIT_TPB) k_bm_fin_write(EmitOut O, const uint4 *__restrict__ lcnt, const int64_t *__restrict__ cex)
{
	typedef cub::BlockScan<int64_t, EMIT_TPB> Scan;
	__shared__ typename Scan::TempStorage tmp[RB3B_ASIZE];
	const int64_t j = (int64_t)blockIdx.x * EMIT_TPB + threadIdx.x;
	int64_t c[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
	if (j < O.n_cells) {
		uint4 q = lcnt[j];
		c[0] = q.x & 0xffffu; c[1] = q.x >> 16; c[2] = q.y & 0xffffu; c[3] = q.y >> 16; c[4] = q.z & 0xffffu; c[5] = q.z >> 16;
	}
	uint64_t h[RB3B_ASIZE];
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		int64_t ex;
		Scan(tmp[a]).ExclusiveSum(c[a], ex);
		h[a] = (uint64_t)(ex + cex[(int64_t)a * (O.n_chunks + 1) + blockIdx.x] - cex[(int64_t)a * (O.n_chunks + 1)]);
	}
	if (j < O.n_cells) {
		O.cells[j * 8] = rb3b_hdr_pack(h[0], h[1], h[2], false);
		O.cells[j * 8 + 4] = rb3b_hdr_pack(h[3], h[4], h[5], false);
	}
}

/* out[a] = last - first of row a of an exclusive scan laid out as RB3B_ASIZE rows of `stride` entries: the six totals in
 * one 48-byte read-back (a dozen 8-byte copies cost ~10 us each in stream order) */
static __global__ void k_gather_tot(const int64_t *__restrict__ ex, int64_t stride, const int *__restrict__ flag, int64_t *__restrict__ out)
{
	const int a = threadIdx.x;
	if (a < RB3B_ASIZE) out[a] = ex[a * stride + stride - 1] - ex[a * stride];
	if (a == RB3B_ASIZE) out[a] = flag ? *flag : 0;
}

/* d_flag: device word, zero on entry, set when the interleave positions turn out not to be monotone (the bitmap -> bitmap
 * kernel validates them while it merges; every other source is validated by the caller beforehand).
 * async: asynchronous merge -- nothing here waits for the device; the totals the kernels counted and *d_flag go to
 * x->pend_host and are checked against x->pend_expect by rb3b_index_wait_i */
template<class Src>
static int rb3b_emit_build_bm(rb3b_index_s *x, Src src, int64_t n_src, int64_t lenB, const int64_t *d_ka, const uint8_t *d_bwt, int *d_flag, bool async = false)
{
	const int *d_async_flag = async ? d_flag : 0;
	const int64_t n_out = n_src + lenB;
	DBuf<int64_t> ilo, ctot, cex;
	DBuf<uint4> lcnt;
	EmitOut O;
	O.shift = RB3B_BM_SHIFT; O.n_out = n_out;
	O.n_cells = (n_out + 127) >> RB3B_BM_SHIFT;
	O.n_chunks = (O.n_cells + EMIT_TPB - 1) / EMIT_TPB;
	TRY(ctot.alloc((O.n_chunks + 1) * RB3B_ASIZE)); TRY(cex.alloc((O.n_chunks + 1) * RB3B_ASIZE)); TRY(lcnt.alloc(O.n_cells));
	if (lenB > 0) {
		TRY(ilo.alloc(O.n_cells + 1));
		k_tile_bounds<<<(unsigned)((O.n_cells + 1 + 255) / 256), 256, 0, rb3b_stream>>>(O.n_cells, O.shift, lenB, n_out, d_ka, ilo.p); CKK();
	}
	TRY(rb3b_reserve((void**)&x->cells2, &x->cap_cells2, O.n_cells * 8, sizeof(uint4)));
	O.cells = x->cells2; O.ovf = 0;
	if (!rb3b_launch_emit_bm_fast(src, O, lenB, d_ka, d_bwt, lenB > 0 ? ilo.p : 0, lcnt.p, ctot.p, d_flag))
		k_emit_bm<Src><<<(unsigned)O.n_chunks, EMIT_TPB, 0, rb3b_stream>>>(src, O, lenB, d_ka, d_bwt, lenB > 0 ? ilo.p : 0, lcnt.p, ctot.p);
	CKK();
	TRY(rb3b_scan_excl_i64(ctot.p, cex.p, (O.n_chunks + 1) * RB3B_ASIZE));
	k_bm_fin_write<<<(unsigned)O.n_chunks, EMIT_TPB, 0, rb3b_stream>>>(O, lcnt.p, cex.p); CKK();
	int64_t tot[RB3B_ASIZE + 1], base[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
	DBuf<int64_t> gt;
	TRY(gt.alloc(RB3B_ASIZE + 1));
	k_gather_tot<<<1, 32, 0, rb3b_stream>>>(cex.p, O.n_chunks + 1, d_flag, gt.p); CKK();
	if (d_async_flag) {
		CK(cudaMemcpyAsync(x->pend_host, gt.p, sizeof(tot), cudaMemcpyDeviceToHost, rb3b_stream));
		for (int a = 0; a < RB3B_ASIZE; ++a) tot[a] = x->pend_expect[a];
	} else {
		CK(cudaMemcpyAsync(tot, gt.p, sizeof(tot), cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		if (tot[RB3B_ASIZE]) return rb3b_fail(RB3B_EINVAL, "interleave positions are not monotone or not all resolved: the batch is not a valid BWT"); /* the old cells stay current */
	}
	{ uint4 *t = x->cells; x->cells = x->cells2; x->cells2 = t; int64_t c = x->cap_cells; x->cap_cells = x->cap_cells2; x->cap_cells2 = c; }
	x->kind = RB3B_KIND_BM; x->shift = RB3B_BM_SHIFT; x->n_cells = O.n_cells; x->n_ovf = 0; x->n_entries = 0;
	x->acc[0] = 0;
	for (int a = 0; a < RB3B_ASIZE; ++a) { x->tot[a] = tot[a] - base[a]; x->acc[a + 1] = x->acc[a] + x->tot[a]; }
	x->n = x->acc[RB3B_ASIZE];
	x->bytes = (size_t)O.n_cells * 128;
	rb3b_stat_set("n_cells", O.n_cells); rb3b_stat_set("n_ovf_blocks", 0); rb3b_stat_set("n_ovf_cells", 0);
	rb3b_stat_set("cell_shift", RB3B_BM_SHIFT); rb3b_stat_set("index_kind", RB3B_KIND_BM);
	if (x->n != n_out) return rb3b_fail(RB3B_EINVAL, "internal error: wrote %lld symbols, expected %lld", (long long)x->n, (long long)n_out);
	return RB3B_OK;
}

#endif
