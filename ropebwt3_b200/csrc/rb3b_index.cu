/*
 * rb3b_index.cu -- the static device index: construction from a plain BWT or a
 * run list, header/directory finalisation, run export, and the batched rank
 * kernels.  Replaces, for the merge path, rb3_enc_plain2fmr (fm-index.c:114-137),
 * rb3_enc_fmd2fmr (fm-index.c:56-85), mr_rank1a/rope_rank2a/rle_rank2a
 * (mrope.c:71-121, rope.c:150-206, rle.c:134-199) and rld_rank1a
 * (rld0.c:416-437).  The data structure is not the reference's: see
 * rb3b_internal.cuh.
 */
#include <string.h>
#include <vector>
#include <cub/cub.cuh>
#include "rb3b_internal.cuh"

#define TPB 256

static inline unsigned nblk(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

/* ------------------------------------------------------------------ */
/* primitives                                                          */
/* ------------------------------------------------------------------ */

int rb3b_scan_excl_i64(const int64_t *d_in, int64_t *d_out, int64_t n)
{
	size_t tmp = 0;
	if (n <= 0) return RB3B_OK;
	CK(cub::DeviceScan::ExclusiveSum((void*)0, tmp, d_in, d_out, n, rb3b_stream));
	DBuf<uint8_t> t;
	TRY(t.alloc(tmp));
	CK(cub::DeviceScan::ExclusiveSum((void*)t.p, tmp, d_in, d_out, n, rb3b_stream));
	return RB3B_OK;
}

static int scan_max_u32(uint32_t *d, int64_t n)
{
	size_t tmp = 0;
	if (n <= 0) return RB3B_OK;
	CK(cub::DeviceScan::InclusiveScan((void*)0, tmp, d, d, cub::Max(), n, rb3b_stream));
	DBuf<uint8_t> t;
	TRY(t.alloc(tmp));
	CK(cub::DeviceScan::InclusiveScan((void*)t.p, tmp, d, d, cub::Max(), n, rb3b_stream));
	return RB3B_OK;
}

/* ------------------------------------------------------------------ */
/* plain BWT -> runs                                                    */
/* ------------------------------------------------------------------ */

#define CHUNK 16

__global__ void k_heads_count(int64_t len, const uint8_t *__restrict__ bwt, int64_t n_chunks, int64_t *__restrict__ cnt, int *__restrict__ bad)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_chunks) return;
	int64_t i0 = t * CHUNK, i1 = min(i0 + CHUNK, len);
	int prev = i0 ? bwt[i0 - 1] : -1, n = 0;
	for (int64_t i = i0; i < i1; ++i) {
		int c = bwt[i];
		if (c >= RB3B_ASIZE) *bad = 1;
		n += c != prev; prev = c;
	}
	cnt[t] = n;
}

__global__ void k_heads_write(int64_t len, const uint8_t *__restrict__ bwt, int64_t n_chunks, const int64_t *__restrict__ off,
                              int64_t *__restrict__ run_start, uint8_t *__restrict__ run_sym)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= n_chunks) return;
	int64_t i0 = t * CHUNK, i1 = min(i0 + CHUNK, len), o = off[t];
	int prev = i0 ? bwt[i0 - 1] : -1;
	for (int64_t i = i0; i < i1; ++i) {
		int c = bwt[i];
		if (c != prev) { run_start[o] = i; run_sym[o] = (uint8_t)c; ++o; }
		prev = c;
	}
}

__global__ void k_starts_to_len(int64_t n_runs, int64_t total, const int64_t *__restrict__ start, int64_t *__restrict__ len)
{
	int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_runs) return;
	len[r] = (r + 1 < n_runs ? start[r + 1] : total) - start[r];
}

static int plain_to_runs_dev(int64_t len, const uint8_t *d_bwt, DBuf<uint8_t> &sym, DBuf<int64_t> &rlen, int64_t *n_runs)
{
	int64_t n_chunks = (len + CHUNK - 1) / CHUNK, last[2];
	DBuf<int64_t> cnt, off, start;
	DBuf<int> bad;
	int hbad = 0;
	*n_runs = 0;
	if (len == 0) return RB3B_OK;
	TRY(cnt.alloc(n_chunks)); TRY(off.alloc(n_chunks)); TRY(bad.alloc(1));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	k_heads_count<<<nblk(n_chunks, TPB), TPB, 0, rb3b_stream>>>(len, d_bwt, n_chunks, cnt.p, bad.p); CKK();
	TRY(rb3b_scan_excl_i64(cnt.p, off.p, n_chunks));
	CK(cudaMemcpyAsync(&last[0], off.p + n_chunks - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&last[1], cnt.p + n_chunks - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (hbad) return rb3b_fail(RB3B_EINVAL, "BWT holds a symbol >= %d (fm-index.c:125 asserts the same)", RB3B_ASIZE);
	*n_runs = last[0] + last[1];
	TRY(start.alloc(*n_runs)); TRY(sym.alloc(*n_runs)); TRY(rlen.alloc(*n_runs));
	k_heads_write<<<nblk(n_chunks, TPB), TPB, 0, rb3b_stream>>>(len, d_bwt, n_chunks, off.p, start.p, sym.p); CKK();
	k_starts_to_len<<<nblk(*n_runs, TPB), TPB, 0, rb3b_stream>>>(*n_runs, len, start.p, rlen.p); CKK();
	return RB3B_OK;
}

/* ------------------------------------------------------------------ */
/* runs -> blocks                                                       */
/* ------------------------------------------------------------------ */

__global__ void k_run_nent(int64_t n_runs, const uint8_t *__restrict__ sym, const int64_t *__restrict__ len, int64_t *__restrict__ nent, int *__restrict__ bad)
{
	int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_runs) return;
	int64_t l = len[r];
	if (sym[r] >= RB3B_ASIZE || l < 0) { *bad = 1; l = 0; }
	nent[r] = rb3b_nent(l);
}

__global__ void k_run_emit(int64_t n_runs, const uint8_t *__restrict__ sym, const int64_t *__restrict__ len, const int64_t *__restrict__ eoff, uint4 *blocks)
{
	int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_runs) return;
	if (len[r] > 0) rb3b_emit_run(blocks, eoff[r], sym[r], len[r]);
}

int rb3b_reserve(void **p, int64_t *cap, int64_t need, size_t elt)
{
	if (*p && *cap >= need) return RB3B_OK;
	if (*p) cudaFreeAsync(*p, rb3b_stream);
	*p = 0; *cap = 0;
	int64_t want = need + need / 2 + 1024;
	cudaError_t e = cudaMallocAsync(p, (size_t)want * elt, rb3b_stream);
	if (e != cudaSuccess) { *p = 0; return rb3b_fail(RB3B_ENOMEM, "cudaMallocAsync(%lld bytes): %s", (long long)(want * (int64_t)elt), cudaGetErrorString(e)); }
	*cap = want;
	return RB3B_OK;
}

int rb3b_index_free_dev(rb3b_index_s *x)
{
	if (x->blocks) cudaFreeAsync(x->blocks, rb3b_stream);
	if (x->spare) cudaFreeAsync(x->spare, rb3b_stream);
	if (x->bstart) cudaFreeAsync(x->bstart, rb3b_stream);
	if (x->dir) cudaFreeAsync(x->dir, rb3b_stream);
	memset(x, 0, sizeof(*x));
	return RB3B_OK;
}

int rb3b_index_from_runs_dev(rb3b_index_s *x, int64_t n_runs, const uint8_t *d_sym, const int64_t *d_len)
{
	DBuf<int64_t> nent, eoff;
	DBuf<int> bad;
	int64_t last[2] = {0, 0};
	int hbad = 0;
	rb3b_index_free_dev(x);
	if (n_runs > 0) {
		TRY(nent.alloc(n_runs)); TRY(eoff.alloc(n_runs)); TRY(bad.alloc(1));
		CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
		k_run_nent<<<nblk(n_runs, TPB), TPB, 0, rb3b_stream>>>(n_runs, d_sym, d_len, nent.p, bad.p); CKK();
		TRY(rb3b_scan_excl_i64(nent.p, eoff.p, n_runs));
		CK(cudaMemcpyAsync(&last[0], eoff.p + n_runs - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaMemcpyAsync(&last[1], nent.p + n_runs - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		if (hbad) return rb3b_fail(RB3B_EINVAL, "run list holds a symbol >= %d or a negative length", RB3B_ASIZE);
	}
	x->n_entries = last[0] + last[1];
	x->n_blocks = (x->n_entries + RB3B_ENT_PER_BLK - 1) / RB3B_ENT_PER_BLK;
	if (x->n_blocks >= (1LL << 32) - 16) return rb3b_fail(RB3B_EINVAL, "index too large for 32-bit block ids");
	if (x->n_blocks == 0) return rb3b_index_finalize(x);
	TRY(rb3b_reserve((void**)&x->blocks, &x->cap_blocks, x->n_blocks * 8, sizeof(uint4)));
	CK(cudaMemsetAsync(x->blocks + (x->n_blocks - 1) * 8, 0, 128, rb3b_stream)); /* padding of the last block */
	k_run_emit<<<nblk(n_runs, TPB), TPB, 0, rb3b_stream>>>(n_runs, d_sym, d_len, eoff.p, x->blocks); CKK();
	return rb3b_index_finalize(x);
}

/* ------------------------------------------------------------------ */
/* finalize: headers, bstart, dir                                       */
/* ------------------------------------------------------------------ */

struct Cnt6 {
	int64_t v[RB3B_ASIZE];
	__host__ __device__ Cnt6 operator+(const Cnt6 &o) const { Cnt6 r; for (int a = 0; a < RB3B_ASIZE; ++a) r.v[a] = v[a] + o.v[a]; return r; }
};

__device__ __forceinline__ Cnt6 blk_counts(const uint4 *__restrict__ blocks, int64_t b)
{
	Cnt6 c;
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a) c.v[a] = 0;
	for (int q = 2; q < 8; ++q) {
		uint4 v = blocks[b * 8 + q];
		const uint32_t w[4] = { v.x, v.y, v.z, v.w };
#pragma unroll
		for (int j = 0; j < 8; ++j) {
			uint32_t e = (w[j >> 1] >> (16 * (j & 1))) & 0xffffu, l = rb3b_ent_len(e), s = e >> 13;
#pragma unroll
			for (int a = 0; a < RB3B_ASIZE; ++a) c.v[a] += s == (uint32_t)a ? l : 0;
		}
	}
	return c;
}

#define FIN_TPB 128

/* pass 1: per-symbol totals of every chunk of FIN_TPB blocks, laid out [6][n_chunks+1] for one flat scan */
__global__ void __launch_bounds__(FIN_TPB) k_fin_count(int64_t nb, const uint4 *__restrict__ blocks, int64_t n_chunks, int64_t *__restrict__ ctot)
{
	typedef cub::BlockReduce<int64_t, FIN_TPB> Red;
	__shared__ typename Red::TempStorage tmp[RB3B_ASIZE];
	int64_t b = (int64_t)blockIdx.x * FIN_TPB + threadIdx.x;
	Cnt6 c;
	if (b < nb) c = blk_counts(blocks, b);
	else for (int a = 0; a < RB3B_ASIZE; ++a) c.v[a] = 0;
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		int64_t t = Red(tmp[a]).Sum(c.v[a]);
		if (threadIdx.x == 0) ctot[(int64_t)a * (n_chunks + 1) + blockIdx.x] = t;
	}
	if (blockIdx.x == 0 && threadIdx.x < RB3B_ASIZE) ctot[(int64_t)threadIdx.x * (n_chunks + 1) + n_chunks] = 0;
}

/* pass 2: recount, scan inside the chunk, add the chunk base: block headers and bstart */
__global__ void __launch_bounds__(FIN_TPB) k_fin_write(int64_t nb, uint4 *blocks, int64_t n_chunks, const int64_t *__restrict__ cex, uint64_t *__restrict__ bstart)
{
	typedef cub::BlockScan<int64_t, FIN_TPB> Scan;
	__shared__ typename Scan::TempStorage tmp[RB3B_ASIZE];
	int64_t b = (int64_t)blockIdx.x * FIN_TPB + threadIdx.x;
	Cnt6 c;
	if (b < nb) c = blk_counts(blocks, b);
	else for (int a = 0; a < RB3B_ASIZE; ++a) c.v[a] = 0;
	uint64_t h[RB3B_ASIZE], s = 0;
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		int64_t ex;
		Scan(tmp[a]).ExclusiveSum(c.v[a], ex);
		h[a] = (uint64_t)(ex + cex[(int64_t)a * (n_chunks + 1) + blockIdx.x] - cex[(int64_t)a * (n_chunks + 1)]);
		s += h[a];
	}
	if (b <= nb) bstart[b] = s; /* thread b == nb sees zero counts of its own: s is the grand total */
	if (b < nb) {
		blocks[b * 8 + 0] = rb3b_hdr_pack(h[0], h[1], h[2]);
		blocks[b * 8 + 1] = rb3b_hdr_pack(h[3], h[4], h[5]);
	}
}

__global__ void k_dir_scatter(int64_t nb, const uint64_t *__restrict__ bstart, int shift, uint32_t *__restrict__ dir, int64_t n_dir)
{
	int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	uint64_t s = bstart[b], j = (s + ((1ULL << shift) - 1)) >> shift;
	if ((j << shift) < bstart[b + 1]) dir[j] = (uint32_t)b;
	if (b == nb - 1) /* cells at or past the end must name the last block: they bound the search from above */
		for (j = (bstart[nb] + ((1ULL << shift) - 1)) >> shift; j < (uint64_t)n_dir; ++j) dir[j] = (uint32_t)b;
}

/* tmp[j] = block containing position j << shift; build the 8-B cells described in rb3b_internal.cuh */
__global__ void k_dir_pack(int64_t n_dir, const uint32_t *__restrict__ tmp, const uint64_t *__restrict__ bstart, int shift, uint64_t *__restrict__ dir)
{
	int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_dir) return;
	uint32_t b0 = tmp[j], b1 = j + 1 < n_dir ? tmp[j + 1] : b0;
	uint64_t lo = (uint64_t)j << shift, hi = (uint64_t)(j + 1) << shift;
	uint32_t inside = b1 - b0; /* block starts in (lo, hi) */
	if (inside && bstart[b1] == hi) --inside;
	uint64_t cell = b0;
	if (inside == 0) cell |= (uint64_t)RB3B_DIR_NONE << 32;
	else {
		uint64_t off = bstart[b0 + 1] - lo;
		if (inside == 1 && off < RB3B_DIR_NONE) cell |= off << 32;
		else cell |= 1ULL << 63;
	}
	dir[j] = cell;
}

int rb3b_index_finalize(rb3b_index_s *x)
{
	int64_t nb = x->n_blocks;
	x->n = 0;
	memset(x->tot, 0, sizeof(x->tot)); memset(x->acc, 0, sizeof(x->acc));
	x->bytes = 0;
	if (nb == 0) { x->n_dir = 0; x->dir_shift = 0; return RB3B_OK; }
	DBuf<int64_t> ctot, cex;
	DBuf<uint32_t> tmp;
	int64_t n_chunks = (nb + 1 + FIN_TPB - 1) / FIN_TPB; /* covers index nb too: that thread writes bstart[nb] */
	int64_t m = (n_chunks + 1) * RB3B_ASIZE, tot[RB3B_ASIZE], base[RB3B_ASIZE];
	TRY(ctot.alloc(m)); TRY(cex.alloc(m));
	TRY(rb3b_reserve((void**)&x->bstart, &x->cap_bstart, nb + 1, 8));
	k_fin_count<<<(unsigned)n_chunks, FIN_TPB, 0, rb3b_stream>>>(nb, x->blocks, n_chunks, ctot.p); CKK();
	TRY(rb3b_scan_excl_i64(ctot.p, cex.p, m));
	k_fin_write<<<(unsigned)n_chunks, FIN_TPB, 0, rb3b_stream>>>(nb, x->blocks, n_chunks, cex.p, x->bstart); CKK();
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		CK(cudaMemcpyAsync(&tot[a], cex.p + a * (n_chunks + 1) + n_chunks, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaMemcpyAsync(&base[a], cex.p + a * (n_chunks + 1), 8, cudaMemcpyDeviceToHost, rb3b_stream));
	}
	CK(cudaStreamSynchronize(rb3b_stream));
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		x->tot[a] = tot[a] - base[a];
		x->acc[a + 1] = x->acc[a] + x->tot[a];
	}
	x->n = x->acc[RB3B_ASIZE];
	if (x->n >= (1LL << 42)) return rb3b_fail(RB3B_EINVAL, "index longer than 2^42 symbols is not supported by the 42-bit block headers");
	/* directory: two to four cells per block */
	int shift = 0;
	while (shift < 40 && (x->n >> (shift + 1)) >= nb * 2) ++shift;
	x->dir_shift = shift;
	x->n_dir = (x->n >> shift) + 2;
	TRY(tmp.alloc(x->n_dir));
	TRY(rb3b_reserve((void**)&x->dir, &x->cap_dir, x->n_dir, 8));
	CK(cudaMemsetAsync(tmp.p, 0, x->n_dir * 4, rb3b_stream));
	k_dir_scatter<<<nblk(nb, TPB), TPB, 0, rb3b_stream>>>(nb, x->bstart, shift, tmp.p, x->n_dir); CKK();
	TRY(scan_max_u32(tmp.p, x->n_dir));
	k_dir_pack<<<nblk(x->n_dir, TPB), TPB, 0, rb3b_stream>>>(x->n_dir, tmp.p, x->bstart, shift, x->dir); CKK();
	x->bytes = (size_t)nb * 128 + (size_t)(nb + 1) * 8 + (size_t)x->n_dir * 8;
	rb3b_stat_set("n_blocks", nb);
	rb3b_stat_set("dir_shift", shift);
	return RB3B_OK;
}

/* ------------------------------------------------------------------ */
/* export: blocks -> canonical (coalesced) run list                     */
/* ------------------------------------------------------------------ */

__device__ __forceinline__ uint32_t blk_entry(const uint4 *blocks, int64_t b, int j)
{
	return ((const uint16_t*)(blocks + b * 8 + 2))[j];
}

template<bool WRITE>
__global__ void k_export(int64_t nb, const uint4 *__restrict__ blocks, const uint64_t *__restrict__ bstart,
                         int64_t *__restrict__ cnt, const int64_t *__restrict__ off, uint8_t *__restrict__ rsym, int64_t *__restrict__ rpos)
{
	int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	int prev = -1, n = 0;
	if (b > 0) prev = blk_entry(blocks, b - 1, RB3B_ENT_PER_BLK - 1) >> 13; /* only the last block may be padded */
	int64_t pos = bstart[b], o = WRITE ? off[b] : 0;
	for (int j = 0; j < RB3B_ENT_PER_BLK; ++j) {
		uint32_t e = blk_entry(blocks, b, j), l = rb3b_ent_len(e);
		int s = e >> 13;
		if (l == 0) continue;
		if (s != prev) {
			if (WRITE) { rsym[o + n] = (uint8_t)s; rpos[o + n] = pos; }
			++n; prev = s;
		}
		pos += l;
	}
	if (!WRITE) cnt[b] = n;
}

int rb3b_export_runs_dev(const rb3b_index_s *x, DBuf<uint8_t> &sym, DBuf<int64_t> &len, int64_t *n_runs)
{
	int64_t nb = x->n_blocks, last[2];
	*n_runs = 0;
	if (nb == 0) return RB3B_OK;
	DBuf<int64_t> cnt, off, pos;
	TRY(cnt.alloc(nb)); TRY(off.alloc(nb));
	k_export<false><<<nblk(nb, 128), 128, 0, rb3b_stream>>>(nb, x->blocks, x->bstart, cnt.p, 0, 0, 0); CKK();
	TRY(rb3b_scan_excl_i64(cnt.p, off.p, nb));
	CK(cudaMemcpyAsync(&last[0], off.p + nb - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&last[1], cnt.p + nb - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	*n_runs = last[0] + last[1];
	TRY(sym.alloc(*n_runs)); TRY(pos.alloc(*n_runs)); TRY(len.alloc(*n_runs));
	k_export<true><<<nblk(nb, 128), 128, 0, rb3b_stream>>>(nb, x->blocks, x->bstart, 0, off.p, sym.p, pos.p); CKK();
	k_starts_to_len<<<nblk(*n_runs, TPB), TPB, 0, rb3b_stream>>>(*n_runs, x->n, pos.p, len.p); CKK();
	return RB3B_OK;
}

/* ------------------------------------------------------------------ */
/* batched rank kernels                                                 */
/* ------------------------------------------------------------------ */

/* rank1a (mr_rank1a / rld_rank1a contract): all six counts and the symbol at k; one 8-lane group per query */
__global__ void __launch_bounds__(TPB) k_rank1a(DevIndex x, int64_t nq, const int64_t *__restrict__ k_, int64_t *__restrict__ ok, int8_t *__restrict__ sym)
{
	typedef Grp<8> G8;
	const int gl = G8::lane();
	int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3, ng = ((int64_t)gridDim.x * blockDim.x) >> 3;
	for (int64_t q = g; q < nq; q += ng) {
		int64_t k = k_[q];
		if (k >= x.n || k < 0) {
			if (gl < RB3B_ASIZE) ok[q * RB3B_ASIZE + gl] = k < 0 ? 0 : x.tot[gl];
			if (gl == 0) sym[q] = -1;
			continue;
		}
		uint4 v[1];
		G8::load(x, G8::locate(x, k), v);
		/* B[k] = the symbol whose count grows between k and k+1 */
		int64_t mine = 0, next = 0;
#pragma unroll
		for (int a = 0; a < RB3B_ASIZE; ++a) {
			int64_t r0 = G8::count(v, k, a), r1 = G8::count(v, k + 1, a);
			if (gl == a) { mine = r0; next = r1; }
		}
		unsigned grew = __ballot_sync(G8::mask(), gl < RB3B_ASIZE && next != mine) >> G8::base();
		if (gl < RB3B_ASIZE) ok[q * RB3B_ASIZE + gl] = mine;
		if (gl == 0) sym[q] = (int8_t)(__ffs(grew) - 1);
	}
}

/* LF flavour used by the merge: out = C[c] + rank(c,k); 144 algorithmic bytes per query */
template<int G>
__global__ void __launch_bounds__(TPB) k_lf(DevIndex x, int64_t nq, const int64_t *__restrict__ k_, const uint8_t *__restrict__ c_, int64_t *__restrict__ out)
{
	const int gl = Grp<G>::lane();
	int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G, ng = ((int64_t)gridDim.x * blockDim.x) / G;
	for (int64_t q = g; q < nq; q += ng) {
		int64_t k = k_[q];
		int c = c_[q];
		int64_t r = x.acc[c] + Grp<G>::rank(x, k, c);
		if (gl == 0) out[q] = r;
	}
}

template<int G> static int launch_lf(const rb3b_index_s *x, int64_t nq, const int64_t *d_k, const uint8_t *d_c, int64_t *d_out, int n_sm)
{
	int occ = 4;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lf<G>, TPB, 0);
	int64_t want = (nq * G + TPB - 1) / TPB, cap = (int64_t)n_sm * (occ > 0 ? occ : 4);
	k_lf<G><<<(unsigned)(want < cap ? want : cap), TPB, 0, rb3b_stream>>>(rb3b_dev_view(x), nq, d_k, d_c, d_out); CKK();
	return RB3B_OK;
}

/* ------------------------------------------------------------------ */
/* C ABI                                                                */
/* ------------------------------------------------------------------ */

static int n_sm(void)
{
	static int n = 0;
	if (!n) { int dev = 0; cudaGetDevice(&dev); cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev); if (n <= 0) n = 148; }
	return n;
}

extern "C" rb3b_index_t *rb3b_index_create(void)
{
	if (rb3b_ensure_init() != RB3B_OK) return 0;
	rb3b_index_s *x = new rb3b_index_s;
	memset(x, 0, sizeof(*x));
	return x;
}

extern "C" void rb3b_index_destroy(rb3b_index_t *x)
{
	if (!x) return;
	rb3b_index_free_dev(x);
	delete x;
}

extern "C" int rb3b_index_from_runs(rb3b_index_t *x, int64_t n_runs, const uint8_t *sym, const int64_t *len)
{
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> ds; DBuf<int64_t> dl;
	TRY(ds.alloc(n_runs)); TRY(dl.alloc(n_runs));
	if (n_runs > 0) {
		CK(cudaMemcpyAsync(ds.p, sym, n_runs, cudaMemcpyHostToDevice, rb3b_stream));
		CK(cudaMemcpyAsync(dl.p, len, n_runs * 8, cudaMemcpyHostToDevice, rb3b_stream));
	}
	return rb3b_index_from_runs_dev(x, n_runs, ds.p, dl.p);
}

extern "C" int rb3b_index_from_runs_device(rb3b_index_t *x, int64_t n_runs, const uint8_t *d_sym, const int64_t *d_len)
{
	TRY(rb3b_ensure_init());
	return rb3b_index_from_runs_dev(x, n_runs, d_sym, d_len);
}

extern "C" int rb3b_index_from_plain_dev(rb3b_index_t *x, int64_t len, const uint8_t *d_bwt)
{
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> sym; DBuf<int64_t> rlen;
	int64_t n_runs;
	if (len < 0) return rb3b_fail(RB3B_EINVAL, "negative length");
	TRY(plain_to_runs_dev(len, d_bwt, sym, rlen, &n_runs));
	return rb3b_index_from_runs_dev(x, n_runs, sym.p, rlen.p);
}

extern "C" int rb3b_index_from_plain(rb3b_index_t *x, int64_t len, const uint8_t *bwt)
{
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> d;
	if (len < 0) return rb3b_fail(RB3B_EINVAL, "negative length");
	TRY(d.alloc(len));
	if (len) CK(cudaMemcpyAsync(d.p, bwt, len, cudaMemcpyHostToDevice, rb3b_stream));
	return rb3b_index_from_plain_dev(x, len, d.p);
}

extern "C" int rb3b_rank1a_dev(const rb3b_index_t *x, int64_t nq, const int64_t *d_k, int64_t *d_ok, int8_t *d_sym)
{
	TRY(rb3b_ensure_init());
	if (nq <= 0) return RB3B_OK;
	int64_t want = (nq * RB3B_GROUP + TPB - 1) / TPB, cap = (int64_t)n_sm() * 8 * 4;
	if (x->n_blocks == 0) { /* empty index: every count is zero (mrope.c:89-93 with zero totals) */
		CK(cudaMemsetAsync(d_ok, 0, nq * RB3B_ASIZE * 8, rb3b_stream));
		CK(cudaMemsetAsync(d_sym, 0xff, nq, rb3b_stream));
		return RB3B_OK;
	}
	k_rank1a<<<(unsigned)(want < cap ? want : cap), TPB, 0, rb3b_stream>>>(rb3b_dev_view(x), nq, d_k, d_ok, d_sym); CKK();
	return RB3B_OK;
}

extern "C" int rb3b_rank1a(const rb3b_index_t *x, int64_t nq, const int64_t *k, int64_t *ok, int8_t *sym)
{
	TRY(rb3b_ensure_init());
	if (nq <= 0) return RB3B_OK;
	DBuf<int64_t> dk, dok; DBuf<int8_t> ds;
	TRY(dk.alloc(nq)); TRY(dok.alloc(nq * RB3B_ASIZE)); TRY(ds.alloc(nq));
	CK(cudaMemcpyAsync(dk.p, k, nq * 8, cudaMemcpyHostToDevice, rb3b_stream));
	TRY(rb3b_rank1a_dev(x, nq, dk.p, dok.p, ds.p));
	CK(cudaMemcpyAsync(ok, dok.p, nq * RB3B_ASIZE * 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(sym, ds.p, nq, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}

int rb3b_lf_tma_launch(const rb3b_index_s *x, int64_t nq, const int64_t *d_k, const uint8_t *d_c, int64_t *d_out); /* rb3b_rank_tma.cu */

extern "C" int rb3b_lf_dev(const rb3b_index_t *x, int64_t nq, const int64_t *d_k, const uint8_t *d_c, int64_t *d_out, int variant)
{ /* variant: 0 default, 1 cp.async.bulk staged, 2/4/8 lanes per query */
	TRY(rb3b_ensure_init());
	if (nq <= 0) return RB3B_OK;
	if (x->n_blocks == 0) return rb3b_fail(RB3B_EINVAL, "empty index");
	if (variant == 1) return rb3b_lf_tma_launch(x, nq, d_k, d_c, d_out);
	if (variant == 0) variant = (int)rb3b_rank_variant ? (int)rb3b_rank_variant : 4;
	if (variant == 2) return launch_lf<2>(x, nq, d_k, d_c, d_out, n_sm());
	if (variant == 4) return launch_lf<4>(x, nq, d_k, d_c, d_out, n_sm());
	if (variant == 8) return launch_lf<8>(x, nq, d_k, d_c, d_out, n_sm());
	return rb3b_fail(RB3B_EINVAL, "unknown rank variant %d", variant);
}

extern "C" int64_t rb3b_get_acc(const rb3b_index_t *x, int64_t acc[RB3B_ASIZE + 1])
{
	for (int a = 0; a <= RB3B_ASIZE; ++a) acc[a] = x->acc[a];
	return x->n;
}

extern "C" int64_t rb3b_index_bytes(const rb3b_index_t *x) { return (int64_t)x->bytes; }

extern "C" int64_t rb3b_export_runs(const rb3b_index_t *x, uint8_t *sym, int64_t *len, int64_t cap)
{
	if (rb3b_ensure_init() != RB3B_OK) return RB3B_ENODEV;
	DBuf<uint8_t> ds; DBuf<int64_t> dl;
	int64_t n_runs;
	int rc = rb3b_export_runs_dev(x, ds, dl, &n_runs);
	if (rc != RB3B_OK) return rc;
	if (sym == 0 || len == 0) { cudaStreamSynchronize(rb3b_stream); return n_runs; }
	if (cap < n_runs) return rb3b_fail(RB3B_EINVAL, "export_runs: capacity %lld < %lld runs", (long long)cap, (long long)n_runs);
	if (n_runs) {
		if (cudaMemcpyAsync(sym, ds.p, n_runs, cudaMemcpyDeviceToHost, rb3b_stream) != cudaSuccess ||
		    cudaMemcpyAsync(len, dl.p, n_runs * 8, cudaMemcpyDeviceToHost, rb3b_stream) != cudaSuccess ||
		    cudaStreamSynchronize(rb3b_stream) != cudaSuccess)
			return rb3b_fail(RB3B_ENODEV, "export_runs: device to host copy failed");
	}
	return n_runs;
}
