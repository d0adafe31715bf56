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
#include "rb3b_emit.cuh"

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
/* runs -> cells                                                        */
/* ------------------------------------------------------------------ */

__global__ void k_run_check(int64_t n_runs, const uint8_t *__restrict__ sym, const int64_t *__restrict__ len, int *__restrict__ bad)
{
	int64_t r = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (r >= n_runs) return;
	if (sym[r] >= RB3B_ASIZE || len[r] < 0) *bad = 1;
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

/* cell span: an average cell about 40% full, never below 32 positions (<= 32 runs always fit inline) */
int rb3b_pick_shift(int64_t n, int64_t n_entries_est)
{
	int shift = RB3B_MIN_SHIFT;
	if (n_entries_est < 1) n_entries_est = 1;
	while (shift < RB3B_MAX_SHIFT && (double)n / (double)n_entries_est * 19.0 >= (double)(2LL << shift)) ++shift;
	return shift;
}



/* bitmap cells cost 1 byte per symbol (twice that with the ping-pong half): use them while that is affordable */
int rb3b_want_bitmap(int64_t n_symbols)
{
	int64_t kind = rb3b_get_param("index_kind", 0); /* 0 auto, 1 RLE cells, 2 bitmap cells */
	if (kind == 1) return 0;
	if (kind == 2) return 1;
	return n_symbols <= rb3b_get_param("bitmap_max_symbols", 24000000000LL);
}

/* host side of an asynchronous merge: wait for it and run the checks that were deferred (rb3b_merge.cu, merge_phase) */
int rb3b_index_wait_i(rb3b_index_s *x)
{
	if (x->pending) {
		CK(cudaEventSynchronize(x->ready_ev));
		x->pending = 0;
		if (x->pend_host[RB3B_ASIZE]) { x->broken = 1; return rb3b_fail(RB3B_EINVAL, "interleave positions are not monotone: the batch is not a valid BWT (the index is unusable now)"); }
		for (int a = 0; a < RB3B_ASIZE; ++a)
			if (x->pend_host[a] != x->pend_expect[a]) {
				x->broken = 1;
				return rb3b_fail(RB3B_EINVAL, "internal error: the merge wrote %lld symbols of code %d, expected %lld", (long long)x->pend_host[a], a, (long long)x->pend_expect[a]);
			}
	}
	if (x->broken) return rb3b_fail(RB3B_EINVAL, "the index was left unusable by an earlier failed merge");
	return RB3B_OK;
}

/* device side: whatever is launched on the current stream from now on runs after the in-flight merge of x */
int rb3b_index_use(const rb3b_index_s *x)
{
	if (x->has_ev) CK(cudaStreamWaitEvent(rb3b_stream, x->ready_ev, 0));
	return RB3B_OK;
}

int rb3b_index_free_dev(rb3b_index_s *x)
{
	if (x->pending) { cudaEventSynchronize(x->ready_ev); x->pending = 0; }
	if (x->has_ev) { cudaStreamWaitEvent(rb3b_stream, x->ready_ev, 0); cudaEventDestroy(x->ready_ev); }
	if (x->ms) cudaFreeAsync(x->ms, rb3b_stream);
	if (x->pend_host) cudaFreeHost(x->pend_host);
	if (x->cells) cudaFreeAsync(x->cells, rb3b_stream);
	if (x->ovf) cudaFreeAsync(x->ovf, rb3b_stream);
	if (x->cells2) cudaFreeAsync(x->cells2, rb3b_stream);
	if (x->ovf2) cudaFreeAsync(x->ovf2, rb3b_stream);
	const int so = x->so; /* a property of the collection, not of the device buffers */
	memset(x, 0, sizeof(*x));
	x->so = so;
	return RB3B_OK;
}

int rb3b_index_from_runs_dev(rb3b_index_s *x, int64_t n_runs, const uint8_t *d_sym, const int64_t *d_len)
{
	DBuf<int64_t> start;
	DBuf<int> bad;
	int64_t last[2] = {0, 0};
	int hbad = 0;
	rb3b_index_free_dev(x);
	if (n_runs <= 0) return RB3B_OK;
	TRY(start.alloc(n_runs)); TRY(bad.alloc(1));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	k_run_check<<<nblk(n_runs, TPB), TPB, 0, rb3b_stream>>>(n_runs, d_sym, d_len, bad.p); CKK();
	TRY(rb3b_scan_excl_i64(d_len, start.p, n_runs));
	CK(cudaMemcpyAsync(&last[0], start.p + n_runs - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&last[1], d_len + n_runs - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (hbad) return rb3b_fail(RB3B_EINVAL, "run list holds a symbol >= %d or a negative length", RB3B_ASIZE);
	int64_t n = last[0] + last[1];
	if (n == 0) return RB3B_OK;
	if (n >= (1LL << 42)) return rb3b_fail(RB3B_EINVAL, "index longer than 2^42 symbols is not supported by the 42-bit cell headers");
	RunSrc src;
	src.sym = d_sym; src.start = start.p; src.n_runs = n_runs; src.n = n; src.r = 0; src.rem64 = 0; src.cur = -1;
	if (rb3b_want_bitmap(n)) return rb3b_emit_build_bm(x, src, n, 0, 0, 0, 0);
	return rb3b_emit_build(x, src, n, 0, 0, 0, n_runs);
}

/* ------------------------------------------------------------------ */
/* export: cells -> canonical (coalesced) run list                      */
/* ------------------------------------------------------------------ */

/* one thread per cell: a run starts wherever the symbol differs from the symbol just before it */
template<class Reader, bool WRITE>
__global__ void k_export(Reader R, int shift, int64_t n, int64_t n_cells, int64_t *__restrict__ cnt, const int64_t *__restrict__ off, uint8_t *__restrict__ rsym, int64_t *__restrict__ rpos)
{
	int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n_cells) return;
	int prev = -1, m = 0;
	if (j > 0) { R.seek((j << shift) - 1); prev = R.cur; }
	int64_t pos = j << shift, end = pos + (1LL << shift) < n ? pos + (1LL << shift) : n, o = WRITE ? off[j] : 0;
	R.seek(pos);
	while (pos < end) {
		if (R.cur != prev) {
			if (WRITE) { rsym[o + m] = (uint8_t)R.cur; rpos[o + m] = pos; }
			++m; prev = R.cur;
		}
		uint32_t l = R.rem;
		if ((int64_t)l > end - pos) l = (uint32_t)(end - pos);
		pos += l;
		if (pos < end) R.advance(l);
	}
	if (!WRITE) cnt[j] = m;
}

template<class Reader>
static int export_with(Reader R, const rb3b_index_s *x, DBuf<uint8_t> &sym, DBuf<int64_t> &len, int64_t *n_runs)
{
	int64_t nc = x->n_cells, last[2];
	DBuf<int64_t> cnt, off, pos;
	TRY(cnt.alloc(nc)); TRY(off.alloc(nc));
	k_export<Reader, false><<<nblk(nc, 128), 128, 0, rb3b_stream>>>(R, x->shift, x->n, nc, cnt.p, 0, 0, 0); CKK();
	TRY(rb3b_scan_excl_i64(cnt.p, off.p, nc));
	CK(cudaMemcpyAsync(&last[0], off.p + nc - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&last[1], cnt.p + nc - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	*n_runs = last[0] + last[1];
	TRY(sym.alloc(*n_runs)); TRY(pos.alloc(*n_runs)); TRY(len.alloc(*n_runs));
	k_export<Reader, true><<<nblk(nc, 128), 128, 0, rb3b_stream>>>(R, x->shift, x->n, nc, 0, off.p, sym.p, pos.p); CKK();
	k_starts_to_len<<<nblk(*n_runs, TPB), TPB, 0, rb3b_stream>>>(*n_runs, x->n, pos.p, len.p); CKK();
	return RB3B_OK;
}

int rb3b_export_runs_dev(const rb3b_index_s *x, DBuf<uint8_t> &sym, DBuf<int64_t> &len, int64_t *n_runs)
{
	*n_runs = 0;
	TRY(rb3b_index_wait_i((rb3b_index_s*)x)); TRY(rb3b_index_use(x));
	if (x->n_cells == 0) return RB3B_OK;
	if (x->kind == RB3B_KIND_BM) {
		BmReader R;
		R.cells = x->cells; R.n = x->n; R.pos = 0; R.cur = -1; R.rem = 0;
		return export_with(R, x, sym, len, n_runs);
	}
	CellReader R;
	R.cells = x->cells; R.ovf = x->ovf; R.n = x->n; R.shift = x->shift;
	R.j = 0; R.left = 0; R.eidx = 0; R.first = 0xffffffffu; R.cur = -1; R.rem = 0;
	return export_with(R, x, sym, len, n_runs);
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
		/* B[k] = the symbol whose count grows between k and k+1 */
		int64_t mine = 0, next = 0;
#pragma unroll
		for (int a = 0; a < RB3B_ASIZE; ++a) {
			int64_t r0 = G8::rank(x, k, a), r1 = G8::rank(x, k + 1, a);
			if (gl == a) { mine = r0; next = r1; }
		}
		unsigned grew = __ballot_sync(G8::mask(), gl < RB3B_ASIZE && next != mine) >> G8::base();
		if (gl < RB3B_ASIZE) ok[q * RB3B_ASIZE + gl] = mine;
		if (gl == 0) sym[q] = (int8_t)(__ffs(grew) - 1);
	}
}

/* bitmap cells: one thread per query */
__global__ void __launch_bounds__(TPB) k_rank1a_bm(DevIndex x, int64_t nq, const int64_t *__restrict__ k_, int64_t *__restrict__ ok, int8_t *__restrict__ sym)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
	for (int64_t q = t; q < nq; q += nt) {
		int64_t k = k_[q];
		int s = -1;
		if (k >= x.n || k < 0) {
			for (int a = 0; a < RB3B_ASIZE; ++a) ok[q * RB3B_ASIZE + a] = k < 0 ? 0 : x.tot[a];
		} else {
			for (int a = 0; a < RB3B_ASIZE; ++a) {
				ok[q * RB3B_ASIZE + a] = BmRank::count(x, k, a);
				uint32_t w = __ldg((const uint32_t*)(x.cells + (k >> RB3B_BM_SHIFT) * 8 + rb3b_bm_plane_quad(a)) + ((k & 127) >> 5));
				if (w >> (k & 31) & 1u) s = a;
			}
		}
		sym[q] = (int8_t)s;
	}
}

__global__ void __launch_bounds__(TPB) k_lf_bm(DevIndex x, int64_t nq, const int64_t *__restrict__ k_, const uint8_t *__restrict__ c_, int64_t *__restrict__ out)
{
	int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
	for (int64_t q = t; q < nq; q += nt) {
		int c = c_[q];
		out[q] = x.acc[c] + BmRank::rank(x, k_[q], c);
	}
}

/* LF flavour used by the merge: out = C[c] + rank(c,k); 144 algorithmic bytes per query.
 * Every group of G lanes works on U queries at a time so that U cell fetches are in flight per group. */
template<int G, int U>
__global__ void __launch_bounds__(TPB) k_lf(DevIndex x, int64_t nq, const int64_t *__restrict__ k_, const uint8_t *__restrict__ c_, int64_t *__restrict__ out)
{
	typedef Grp<G> GG;
	const int gl = GG::lane();
	int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / G, ng = ((int64_t)gridDim.x * blockDim.x) / G;
	for (int64_t q0 = g; q0 < nq; q0 += ng * U) {
		int64_t k[U], kk[U];
		int c[U];
		uint4 v[U][GG::NQ];
#pragma unroll
		for (int u = 0; u < U; ++u) {
			int64_t q = q0 + (int64_t)u * ng;
			k[u] = q < nq ? k_[q] : 0; c[u] = q < nq ? c_[q] : 0;
			kk[u] = k[u] < x.n ? (k[u] < 0 ? 0 : k[u]) : x.n - 1;
			GG::load(x, kk[u] >> x.shift, v[u]);
		}
#pragma unroll
		for (int u = 0; u < U; ++u) {
			int64_t q = q0 + (int64_t)u * ng;
			int64_t r = GG::count(x, v[u], kk[u], c[u]);
			r = k[u] < x.n ? r : x.tot[c[u]];
			if (gl == 0 && q < nq) out[q] = x.acc[c[u]] + r;
		}
	}
}

/* run-length cells, ONE thread per query: the header quad of the symbol's half, then the entry quads one after the other
 * until the offset is reached (an average cell is ~40 % full and the offset uniform in it: mostly one or two of them).
 * 4-5x fewer instructions per query than the lane-group kernels, every thread its own query in flight. */
__global__ void __launch_bounds__(TPB) k_lf_t1(DevIndex x, int64_t nq, const int64_t *__restrict__ k_, const uint8_t *__restrict__ c_, int64_t *__restrict__ out)
{
	const int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x, nt = (int64_t)gridDim.x * blockDim.x;
	for (int64_t q = t; q < nq; q += nt) {
		const int64_t k = k_[q];
		const int c = c_[q];
		const int64_t kk = k < x.n ? (k < 0 ? 0 : k) : x.n - 1;
		const uint4 *cell = x.cells + (kk >> x.shift) * 8;
		const int h = c >= 3, cc = c - 3 * h;
		const uint4 hq = __ldg(cell + h);
		uint4 e0 = __ldg(cell + 2), e1 = __ldg(cell + 3); /* in flight together with the header */
		uint64_t a0, a1, a2;
		rb3b_hdr_unpack(hq, a0, a1, a2);
		const uint64_t basec = cc == 0 ? a0 : cc == 1 ? a1 : a2;
		const uint32_t off = (uint32_t)kk & ((1u << x.shift) - 1u);
		uint32_t cnt = 0;
		if (hq.w >> 31) cnt = rb3b_ovf_count(cell, x.ovf, off, c);
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
		out[q] = x.acc[c] + (k < x.n ? (int64_t)(basec + cnt) : x.tot[c]);
	}
}

template<int G, int U> static int launch_lf(const rb3b_index_s *x, int64_t nq, const int64_t *d_k, const uint8_t *d_c, int64_t *d_out, int n_sm)
{
	int occ = 4;
	cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lf<G, U>, TPB, 0);
	int64_t want = (nq * G + TPB - 1) / TPB, cap = (int64_t)n_sm * (occ > 0 ? occ : 4);
	k_lf<G, U><<<(unsigned)(want < cap ? want : cap), TPB, 0, rb3b_stream>>>(rb3b_dev_view(x), nq, d_k, d_c, d_out); CKK();
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

extern "C" int rb3b_index_set_order(rb3b_index_t *x, int so)
{ /* mr_init(max_nodes, block_len, sorting_order), mrope.c:15 */
	if (so < 0 || so > 2) return rb3b_fail(RB3B_EINVAL, "sorting order must be 0 (input), 1 (RLO) or 2 (RCLO)");
	if (x->n > 0 && so != x->so) return rb3b_fail(RB3B_EINVAL, "the sorting order of a non-empty index cannot be changed");
	x->so = so;
	return RB3B_OK;
}

extern "C" int rb3b_index_get_order(const rb3b_index_t *x) { return x->so; }

extern "C" int rb3b_index_reserve(rb3b_index_t *x, int64_t n_symbols)
{ /* like vector::reserve: size both ping-pong halves for an index of n_symbols so that merges never reallocate */
	TRY(rb3b_ensure_init());
	TRY(rb3b_index_wait_i(x)); TRY(rb3b_index_use(x));
	if (!rb3b_want_bitmap(n_symbols)) return RB3B_OK; /* RLE cells: size depends on the data; grown on demand */
	int64_t quads = ((n_symbols + 127) >> RB3B_BM_SHIFT) * 8;
	if (x->cap_cells2 < quads) {
		if (x->cells2) cudaFreeAsync(x->cells2, rb3b_stream);
		x->cells2 = 0; x->cap_cells2 = 0;
		CK(cudaMallocAsync((void**)&x->cells2, (size_t)quads * sizeof(uint4), rb3b_stream));
		x->cap_cells2 = quads;
	}
	if (x->cap_cells < quads) { /* the current half holds live data: move it */
		uint4 *p = 0;
		CK(cudaMallocAsync((void**)&p, (size_t)quads * sizeof(uint4), rb3b_stream));
		if (x->cells) {
			CK(cudaMemcpyAsync(p, x->cells, (size_t)x->n_cells * 128, cudaMemcpyDeviceToDevice, rb3b_stream));
			cudaFreeAsync(x->cells, rb3b_stream);
		}
		x->cells = p; x->cap_cells = quads;
	}
	return RB3B_OK;
}

extern "C" void rb3b_index_destroy(rb3b_index_t *x)
{
	if (!x) return;
	rb3b_index_free_dev(x);
	delete x;
}

extern "C" int rb3b_index_from_runs(rb3b_index_t *x, int64_t n_runs, const uint8_t *sym, const int64_t *len)
{
	ApiScope scope_;
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
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	return rb3b_index_from_runs_dev(x, n_runs, d_sym, d_len);
}

extern "C" int rb3b_index_from_plain_dev(rb3b_index_t *x, int64_t len, const uint8_t *d_bwt)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> sym; DBuf<int64_t> rlen;
	int64_t n_runs;
	if (len < 0) return rb3b_fail(RB3B_EINVAL, "negative length");
	TRY(plain_to_runs_dev(len, d_bwt, sym, rlen, &n_runs));
	return rb3b_index_from_runs_dev(x, n_runs, sym.p, rlen.p);
}

extern "C" int rb3b_index_from_plain(rb3b_index_t *x, int64_t len, const uint8_t *bwt)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	DBuf<uint8_t> d;
	if (len < 0) return rb3b_fail(RB3B_EINVAL, "negative length");
	TRY(d.alloc(len));
	if (len) CK(cudaMemcpyAsync(d.p, bwt, len, cudaMemcpyHostToDevice, rb3b_stream));
	return rb3b_index_from_plain_dev(x, len, d.p);
}

extern "C" int rb3b_rank1a_dev(const rb3b_index_t *x, int64_t nq, const int64_t *d_k, int64_t *d_ok, int8_t *d_sym)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (nq <= 0) return RB3B_OK;
	TRY(rb3b_index_wait_i((rb3b_index_s*)x)); TRY(rb3b_index_use(x));
	int64_t want = (nq * RB3B_GROUP + TPB - 1) / TPB, cap = (int64_t)n_sm() * 8 * 4;
	if (x->n_cells == 0) { /* empty index: every count is zero (mrope.c:89-93 with zero totals) */
		CK(cudaMemsetAsync(d_ok, 0, nq * RB3B_ASIZE * 8, rb3b_stream));
		CK(cudaMemsetAsync(d_sym, 0xff, nq, rb3b_stream));
		return RB3B_OK;
	}
	if (x->kind == RB3B_KIND_BM) {
		want = (nq + TPB - 1) / TPB;
		k_rank1a_bm<<<(unsigned)(want < cap ? want : cap), TPB, 0, rb3b_stream>>>(rb3b_dev_view(x), nq, d_k, d_ok, d_sym); CKK();
		return RB3B_OK;
	}
	k_rank1a<<<(unsigned)(want < cap ? want : cap), TPB, 0, rb3b_stream>>>(rb3b_dev_view(x), nq, d_k, d_ok, d_sym); CKK();
	return RB3B_OK;
}

extern "C" int rb3b_rank1a(const rb3b_index_t *x, int64_t nq, const int64_t *k, int64_t *ok, int8_t *sym)
{
	ApiScope scope_;
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
{ /* variant: 0 default, 1 cp.async.bulk staged, G (2/4/8) lanes per query, GU = G lanes with U queries in flight */
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (nq <= 0) return RB3B_OK;
	TRY(rb3b_index_wait_i((rb3b_index_s*)x)); TRY(rb3b_index_use(x));
	if (x->n_cells == 0) return rb3b_fail(RB3B_EINVAL, "empty index");
	if (x->kind == RB3B_KIND_BM) { /* bitmap cells have a single kernel: two 16-B loads and a popcount per query */
		int64_t want = (nq + TPB - 1) / TPB, cap = (int64_t)n_sm() * 8 * 4;
		k_lf_bm<<<(unsigned)(want < cap ? want : cap), TPB, 0, rb3b_stream>>>(rb3b_dev_view(x), nq, d_k, d_c, d_out); CKK();
		return RB3B_OK;
	}
	if (variant == 1) return rb3b_lf_tma_launch(x, nq, d_k, d_c, d_out);
	if (variant == 0) variant = (int)rb3b_rank_variant ? (int)rb3b_rank_variant : 11; /* one thread per query: 0.47-0.58 of the HBM peak against 0.37 for the best lane-group kernel */
	if (variant == 11) {
		int occ = 4;
		cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, k_lf_t1, TPB, 0);
		int64_t want = (nq + TPB - 1) / TPB, cap = (int64_t)n_sm() * (occ > 0 ? occ : 4) * 4;
		k_lf_t1<<<(unsigned)(want < cap ? want : cap), TPB, 0, rb3b_stream>>>(rb3b_dev_view(x), nq, d_k, d_c, d_out); CKK();
		return RB3B_OK;
	}
	if (variant == 2) return launch_lf<2, 1>(x, nq, d_k, d_c, d_out, n_sm());
	if (variant == 4) return launch_lf<4, 1>(x, nq, d_k, d_c, d_out, n_sm());
	if (variant == 8) return launch_lf<8, 1>(x, nq, d_k, d_c, d_out, n_sm());
	if (variant == 22) return launch_lf<2, 2>(x, nq, d_k, d_c, d_out, n_sm()); /* G lanes, 2 queries in flight per group */
	if (variant == 42) return launch_lf<4, 2>(x, nq, d_k, d_c, d_out, n_sm());
	if (variant == 82) return launch_lf<8, 2>(x, nq, d_k, d_c, d_out, n_sm());
	if (variant == 44) return launch_lf<4, 4>(x, nq, d_k, d_c, d_out, n_sm());
	if (variant == 84) return launch_lf<8, 4>(x, nq, d_k, d_c, d_out, n_sm());
	return rb3b_fail(RB3B_EINVAL, "unknown rank variant %d", variant);
}

extern "C" int64_t rb3b_get_acc(const rb3b_index_t *x, int64_t acc[RB3B_ASIZE + 1])
{
	for (int a = 0; a <= RB3B_ASIZE; ++a) acc[a] = x->acc[a];
	return x->n;
}

extern "C" int64_t rb3b_index_bytes(const rb3b_index_t *x) { return (int64_t)x->bytes; }

/* block until an asynchronous merge into idx has finished (its batch buffer may then be reused) and report its deferred checks */
extern "C" int rb3b_index_wait(rb3b_index_t *x)
{
	TRY(rb3b_ensure_init());
	return rb3b_index_wait_i(x);
}

extern "C" int64_t rb3b_export_runs(const rb3b_index_t *x, uint8_t *sym, int64_t *len, int64_t cap)
{
	ApiScope scope_;
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
