/*
 * rb3b_fmd_dev.cu -- the .fmd encoder on the device: rld_enc / rld_enc1 / enc_next_block /
 * rld_enc_finish / rld_rank_index / rld_dump (rld0.c:107-243), byte-identical to the reference's file for the same run
 * sequence.  The canonical run list never leaves the device; only the finished image is copied to the host.
 *
 * rld_enc1 is a greedy bit packer with data-dependent block breaks: a run's code goes into the open 64-byte block iff the
 * payload bits used so far plus its width stay BELOW the block's payload capacity, and the capacity depends on the header
 * width, i.e. on the number of symbols in the block before (16/32/64-bit headers).  So the block that starts at run j
 * with header type t ends at a run that follows from two prefix sums (code widths, run lengths): the block starts form a
 * chain in a successor graph over (run, type) states.  The chain is found in three passes:
 *
 *   1. the run list is cut into segments of SEG runs.  For every segment and EVERY state in which the chain can enter it
 *      (offset of the first block start inside the segment < 97, header type 0/1) one thread follows the successor function
 *      to the end of the segment and records the state in which the chain leaves it and how many blocks it opened
 *      (k_fmd_segment_maps) -- all segments at once, no knowledge of the real chain needed;
 *   2. one thread block composes those maps from segment 0 on (k_fmd_compose; the maps of the next 32 segments are staged in
 *      shared memory while one thread walks them), which gives every segment its real entry state and first block number.
 *      Every 2^20-th block is the last of a 2^23-word chunk and one word shorter (rld0.h:81): a segment that holds such a
 *      block is walked block by block instead of through its map;
 *   3. every segment replays its own stretch of the chain and writes block starts and header types (k_fmd_replay).
 *
 * Headers + payload are then one thread per block, the rank index one thread per frame (binary search over the blocks'
 * start positions).  A block of 2^30 symbols or more (64-bit header) sends the caller back to the host writer.
 */
#include <string.h>
#include <cub/cub.cuh>
#include "rb3b_internal.cuh"

#define TPB 256
#define FMD_SSIZE 8
#define FMD_CHUNK_BLOCKS (1LL << 20)     /* 2^23 words per chunk / 8 words per block */
#define SEG 8192                         /* runs per segment */
#define OMAX 97                          /* a block holds at most 384 / 4 = 96 runs: the first block start of a segment is at offset < 97 */
#define NCAND (2 * OMAX)

static inline unsigned nblk(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

__device__ __forceinline__ int ilog2_dev(uint64_t v) { return 63 - __clzll((long long)v); }
__device__ __forceinline__ int fmd_width(int64_t len) { const int y = ilog2_dev((uint64_t)len), z = 31 - __clz(y + 1); return 2 * z + 1 + y + 3; }
__device__ __forceinline__ int fmd_hdr_words(int t) { return t == 0 ? 2 : t == 1 ? 4 : 7; }

__global__ void k_fmd_widths(int64_t R, const uint8_t *__restrict__ sym, const int64_t *__restrict__ len, int64_t *__restrict__ wid, int *__restrict__ bad)
{
	const int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= R) return;
	if (len[j] <= 0 || sym[j] >= RB3B_ASIZE || (j > 0 && sym[j] == sym[j - 1])) { *bad = 1; wid[j] = 4; return; }
	wid[j] = fmd_width(len[j]);
}

/* the block that starts at run j with header type t and capacity `cap` bits ends before run j' = the first run whose code
 * does not fit (W = exclusive prefix sums of the widths, W[R] = total); at least one run always fits */
__device__ __forceinline__ int64_t fmd_next(const int64_t *__restrict__ W, int64_t R, int64_t j, int64_t cap)
{
	const int64_t lim = W[j] + cap; /* runs j..k-1 fit iff W[k] < lim */
	int64_t lo = j + 1, hi = j + OMAX < R ? j + OMAX : R; /* the answer lies in [lo, hi] */
	while (lo < hi) {
		const int64_t mid = (lo + hi + 1) >> 1;
		if (W[mid] < lim) lo = mid; else hi = mid - 1;
	}
	return lo;
}

__device__ __forceinline__ int fmd_type_of(int64_t nsym) { return nsym < 0x4000 ? 0 : nsym < 0x40000000 ? 1 : 2; }
__device__ __forceinline__ int64_t fmd_cap(int t, bool chunk_end) { return (int64_t)(FMD_SSIZE - fmd_hdr_words(t) - (chunk_end ? 1 : 0)) * 64; }

/* pass 1: E[g][c] for candidate c = o * 2 + t: exit offset (7 bits) | exit type (1 bit) << 7 | blocks opened << 8; 0xffffffff = not computed (past the end) */
__global__ void k_fmd_segment_maps(int64_t R, int64_t n_seg, const int64_t *__restrict__ W, const int64_t *__restrict__ L, uint32_t *__restrict__ E, int *__restrict__ wide)
{
	const int64_t id = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (id >= n_seg * NCAND) return;
	const int64_t g = id / NCAND, end = (g + 1) * SEG < R ? (g + 1) * SEG : R;
	const int c = (int)(id % NCAND);
	int64_t j = g * SEG + (c >> 1);
	int t = c & 1;
	uint32_t nb = 0;
	if (j >= R) { E[id] = 0xffffffffu; return; }
	while (j < end) {
		const int64_t j2 = fmd_next(W, R, j, fmd_cap(t, false));
		t = fmd_type_of(L[j2] - L[j]);
		if (t == 2) { *wide = 1; t = 1; }
		j = j2; ++nb;
	}
	E[id] = (uint32_t)(j - end) | (uint32_t)t << 7 | nb << 8;
}

/* pass 2: entry state of every segment: ent[g] = offset | type << 7, blk0[g] = number of the first block opened in g */
__global__ void __launch_bounds__(NCAND) k_fmd_compose(int64_t R, int64_t n_seg, const int64_t *__restrict__ W, const int64_t *__restrict__ L, const uint32_t *__restrict__ E,
                                                       uint32_t *__restrict__ ent, int64_t *__restrict__ blk0, int64_t *__restrict__ out /* [0] data blocks, [1] type of the trailing block */)
{
	__shared__ uint32_t sh[32][NCAND];
	__shared__ int s_o, s_t;
	__shared__ long long s_b;
	if (threadIdx.x == 0) { s_o = 0; s_t = 0; s_b = 0; }
	for (int64_t g0 = 0; g0 < n_seg; g0 += 32) {
		__syncthreads();
		for (int i = 0; i < 32; ++i) if (g0 + i < n_seg) sh[i][threadIdx.x] = E[(g0 + i) * NCAND + threadIdx.x];
		__syncthreads();
		if (threadIdx.x == 0) {
			int o = s_o, t = s_t;
			int64_t b = s_b;
			for (int i = 0; i < 32 && g0 + i < n_seg; ++i) {
				const int64_t g = g0 + i, end = (g + 1) * SEG < R ? (g + 1) * SEG : R;
				ent[g] = (uint32_t)o | (uint32_t)t << 7;
				blk0[g] = b;
				const uint32_t e = sh[i][o * 2 + t];
				const int64_t nb = e >> 8;
				/* does one of the blocks b .. b+nb-1 end a chunk?  then the map does not apply */
				const int64_t next_end = (b / FMD_CHUNK_BLOCKS + 1) * FMD_CHUNK_BLOCKS - 1;
				if (e != 0xffffffffu && next_end >= b + nb) { o = (int)(e & 127u); t = (int)(e >> 7 & 1u); b += nb; continue; }
				int64_t j = g * SEG + o;
				while (j < end) {
					const int64_t j2 = fmd_next(W, R, j, fmd_cap(t, (b + 1) % FMD_CHUNK_BLOCKS == 0));
					t = fmd_type_of(L[j2] - L[j]);
					if (t == 2) t = 1; /* reported by pass 1 or 3 */
					j = j2; ++b;
				}
				o = (int)(j - end);
			}
			s_o = o; s_t = t; s_b = b;
		}
	}
	__syncthreads();
	if (threadIdx.x == 0) { out[0] = s_b; out[1] = s_t; }
}

/* pass 3: block starts and header types */
__global__ void k_fmd_replay(int64_t R, int64_t n_seg, const int64_t *__restrict__ W, const int64_t *__restrict__ L, const uint32_t *__restrict__ ent, const int64_t *__restrict__ blk0,
                             int64_t *__restrict__ bstart, uint8_t *__restrict__ btype, int *__restrict__ wide)
{
	const int64_t g = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (g >= n_seg) return;
	const int64_t end = (g + 1) * SEG < R ? (g + 1) * SEG : R;
	int64_t j = g * SEG + (ent[g] & 127u), b = blk0[g];
	int t = (int)(ent[g] >> 7 & 1u);
	while (j < end) {
		bstart[b] = j; btype[b] = (uint8_t)t;
		const int64_t j2 = fmd_next(W, R, j, fmd_cap(t, (b + 1) % FMD_CHUNK_BLOCKS == 0));
		t = fmd_type_of(L[j2] - L[j]);
		if (t == 2) { *wide = 1; t = 1; }
		j = j2; ++b;
	}
}

/* headers (counts of the previous block, enc_next_block rld0.c:107-135) and payload (rld_delta_enc1 + rld_enc1) of block b;
 * blocks 0 .. n_data-1 hold runs, block n_data is the trailing header-only block.  cnt6[b*6+a] = symbols a in block b. */
__global__ void k_fmd_blocks(int64_t n_data, const int64_t *__restrict__ bstart, const uint8_t *__restrict__ btype, int trail_type, int64_t R,
                             const uint8_t *__restrict__ sym, const int64_t *__restrict__ len, uint64_t *__restrict__ words, int64_t *__restrict__ cnt6)
{
	const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b > n_data) return;
	uint64_t blk[FMD_SSIZE] = {0, 0, 0, 0, 0, 0, 0, 0};
	const int type = b < n_data ? btype[b] : trail_type;
	if (b > 0) {
		uint64_t delta[RB3B_ASIZE + 1] = {0, 0, 0, 0, 0, 0, 0};
		const int64_t j0 = bstart[b - 1], j1 = b < n_data ? bstart[b] : R;
		for (int64_t j = j0; j < j1; ++j) { const uint64_t l = (uint64_t)len[j]; delta[0] += l; delta[sym[j] + 1] += l; }
		if (type == 0) {
			blk[0] = delta[0] | delta[1] << 16 | delta[2] << 32 | delta[3] << 48;
			blk[1] = delta[4] | delta[5] << 16 | delta[6] << 32;
		} else {
			blk[0] = delta[0] | delta[1] << 32; blk[1] = delta[2] | delta[3] << 32;
			blk[2] = delta[4] | delta[5] << 32; blk[3] = delta[6];
		}
		blk[0] |= (uint64_t)type << 62;
	}
	if (b < n_data) {
		int64_t c6[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
		int cur = fmd_hdr_words(type), fr = 64;
		const int64_t j0 = bstart[b], j1 = b + 1 < n_data ? bstart[b + 1] : R;
		for (int64_t j = j0; j < j1; ++j) {
			const int64_t l = len[j];
			const int s = sym[j], y = ilog2_dev((uint64_t)l), width = fmd_width(l);
			const uint64_t bits = ((((uint64_t)l ^ (1ULL << y)) | (uint64_t)(y + 1) << y) << 3) | (uint64_t)s;
#pragma unroll
			for (int a = 0; a < RB3B_ASIZE; ++a) c6[a] += a == s ? l : 0;
			uint64_t hi_part, lo_part = 0;
			int adv = 0;
			if (width > fr) { const int spill = width - fr; hi_part = bits >> spill; lo_part = bits << (64 - spill); adv = 1; fr = 64 - spill; }
			else { fr -= width; hi_part = bits << fr; }
#pragma unroll
			for (int w = 0; w < FMD_SSIZE; ++w) { if (w == cur) blk[w] |= hi_part; if (adv && w == cur + 1) blk[w] |= lo_part; }
			cur += adv;
		}
#pragma unroll
		for (int a = 0; a < RB3B_ASIZE; ++a) cnt6[b * RB3B_ASIZE + a] = c6[a];
	} else {
#pragma unroll
		for (int a = 0; a < RB3B_ASIZE; ++a) cnt6[b * RB3B_ASIZE + a] = 0;
	}
	uint4 *o = (uint4*)(words + b * FMD_SSIZE);
	o[0] = make_uint4((uint32_t)blk[0], (uint32_t)(blk[0] >> 32), (uint32_t)blk[1], (uint32_t)(blk[1] >> 32));
	o[1] = make_uint4((uint32_t)blk[2], (uint32_t)(blk[2] >> 32), (uint32_t)blk[3], (uint32_t)(blk[3] >> 32));
	o[2] = make_uint4((uint32_t)blk[4], (uint32_t)(blk[4] >> 32), (uint32_t)blk[5], (uint32_t)(blk[5] >> 32));
	o[3] = make_uint4((uint32_t)blk[6], (uint32_t)(blk[6] >> 32), (uint32_t)blk[7], (uint32_t)(blk[7] >> 32));
}

/* transpose the per-block counts into six arrays (for six scans) and the total per block */
__global__ void k_fmd_cnt_split(int64_t nb, const int64_t *__restrict__ cnt6, int64_t *__restrict__ planes /* [7][nb]: six symbols, then all */)
{
	const int64_t b = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (b >= nb) return;
	int64_t tot = 0;
#pragma unroll
	for (int a = 0; a < RB3B_ASIZE; ++a) { const int64_t v = cnt6[b * RB3B_ASIZE + a]; planes[a * nb + b] = v; tot += v; }
	planes[RB3B_ASIZE * nb + b] = tot;
}

/* rld_rank_index (rld0.c:163-204): frame k = word offset and cumulative counts of the last block (number >= 1) whose start
 * position is below k << ibits; all-zero when there is none.  ex[a*nb + b] = symbols a before block b, ex[6*nb + b] = all. */
__global__ void k_fmd_frames(int64_t n_frames, int ibits, int64_t nb, const int64_t *__restrict__ ex, uint64_t *__restrict__ frames)
{
	const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (k >= n_frames) return;
	uint64_t f[RB3B_ASIZE + 1] = {0, 0, 0, 0, 0, 0, 0};
	if (k > 0) {
		const int64_t key = k << ibits;
		const int64_t *pos = ex + RB3B_ASIZE * nb;
		int64_t lo = 0, hi = nb - 1; /* last block with pos < key; block 0 has pos 0 < key */
		while (lo < hi) {
			const int64_t mid = (lo + hi + 1) >> 1;
			if (pos[mid] < key) lo = mid; else hi = mid - 1;
		}
		if (lo >= 1) {
			f[0] = (uint64_t)lo * FMD_SSIZE;
#pragma unroll
			for (int a = 0; a < RB3B_ASIZE; ++a) f[a + 1] = (uint64_t)ex[a * nb + lo];
		}
	}
#pragma unroll
	for (int a = 0; a <= RB3B_ASIZE; ++a) frames[k * (RB3B_ASIZE + 1) + a] = f[a];
}

static int scan_i64(const int64_t *in, int64_t *out, int64_t n)
{
	size_t tmp = 0;
	if (n <= 0) return RB3B_OK;
	CK(cub::DeviceScan::ExclusiveSum((void*)0, tmp, in, out, n, rb3b_stream));
	DBuf<uint8_t> t;
	TRY(t.alloc(tmp));
	CK(cub::DeviceScan::ExclusiveSum((void*)t.p, tmp, in, out, n, rb3b_stream));
	return RB3B_OK;
}

/* Run list in device memory -> .fmd image in host memory (malloc'd, *out).  Returns the image size, 0 when the list needs
 * the host writer (a 64-bit block header, a list that is not canonical, fewer than 2 runs), < 0 on error. */
int64_t rb3b_fmd_image_dev(int64_t R, const uint8_t *d_sym, const int64_t *d_len, const int64_t tot_sym[RB3B_ASIZE], uint8_t **out)
{
	*out = 0;
	if (R < 2) return 0;
	DBuf<int64_t> wid, W, L, blk0, res, bstart, cnt6, planes, ex;
	DBuf<uint32_t> E, ent;
	DBuf<uint8_t> btype;
	DBuf<int> flags;
	const int64_t n_seg = (R + SEG - 1) / SEG;
	TRY(wid.alloc(R + 1)); TRY(W.alloc(R + 1)); TRY(L.alloc(R + 1)); TRY(flags.alloc(2));
	CK(cudaMemsetAsync(flags.p, 0, 8, rb3b_stream));
	CK(cudaMemsetAsync(wid.p + R, 0, 8, rb3b_stream));
	k_fmd_widths<<<nblk(R, TPB), TPB, 0, rb3b_stream>>>(R, d_sym, d_len, wid.p, flags.p); CKK();
	TRY(scan_i64(wid.p, W.p, R + 1));   /* W[R] = all code bits */
	{ /* L: exclusive prefix sums of the run lengths, L[R] = total (the scan of R + 1 items reads one item past the list: use a padded copy) */
		DBuf<int64_t> lpad;
		TRY(lpad.alloc(R + 1));
		CK(cudaMemcpyAsync(lpad.p, d_len, (size_t)R * 8, cudaMemcpyDeviceToDevice, rb3b_stream));
		CK(cudaMemsetAsync(lpad.p + R, 0, 8, rb3b_stream));
		TRY(scan_i64(lpad.p, L.p, R + 1));
	}
	TRY(E.alloc((size_t)n_seg * NCAND)); TRY(ent.alloc(n_seg)); TRY(blk0.alloc(n_seg)); TRY(res.alloc(2));
	k_fmd_segment_maps<<<nblk(n_seg * NCAND, TPB), TPB, 0, rb3b_stream>>>(R, n_seg, W.p, L.p, E.p, flags.p + 1); CKK();
	k_fmd_compose<<<1, NCAND, 0, rb3b_stream>>>(R, n_seg, W.p, L.p, E.p, ent.p, blk0.p, res.p); CKK();
	int64_t hres[2];
	int hflags[2];
	CK(cudaMemcpyAsync(hres, res.p, 16, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(hflags, flags.p, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (hflags[0] || hflags[1]) return 0; /* not canonical / 64-bit header somewhere: the host writer handles both */
	const int64_t n_data = hres[0], nb = n_data + 1;
	const int trail_type = (int)hres[1];
	TRY(bstart.alloc(nb)); TRY(btype.alloc(nb)); TRY(cnt6.alloc(nb * RB3B_ASIZE)); TRY(planes.alloc(nb * (RB3B_ASIZE + 1))); TRY(ex.alloc(nb * (RB3B_ASIZE + 1)));
	k_fmd_replay<<<nblk(n_seg, 128), 128, 0, rb3b_stream>>>(R, n_seg, W.p, L.p, ent.p, blk0.p, bstart.p, btype.p, flags.p + 1); CKK();
	/* sizes (rld_enc_finish rld0.c:206-216, rld_rank_index rld0.c:163-176) */
	uint64_t total = 0;
	for (int a = 0; a < RB3B_ASIZE; ++a) total += (uint64_t)tot_sym[a];
	const uint64_t n_words = (uint64_t)n_data * FMD_SSIZE + (trail_type == 0 ? 2 : 4), n_bytes = n_words * 8;
	const uint64_t n_blks = n_words / FMD_SSIZE + 1;
	int ibits = 0;
	{ uint64_t q = total / n_blks; ibits = (q ? 63 - __builtin_clzll(q) : -1) + 4; }
	const uint64_t n_frames = ((total + (1ULL << ibits) - 1) >> ibits) + 1;
	DBuf<uint64_t> body, frames;
	TRY(body.alloc((size_t)nb * FMD_SSIZE)); TRY(frames.alloc((size_t)n_frames * (RB3B_ASIZE + 1)));
	k_fmd_blocks<<<nblk(nb, 128), 128, 0, rb3b_stream>>>(n_data, bstart.p, btype.p, trail_type, R, d_sym, d_len, body.p, cnt6.p); CKK();
	k_fmd_cnt_split<<<nblk(nb, TPB), TPB, 0, rb3b_stream>>>(nb, cnt6.p, planes.p); CKK();
	for (int a = 0; a <= RB3B_ASIZE; ++a) TRY(scan_i64(planes.p + a * nb, ex.p + a * nb, nb));
	k_fmd_frames<<<nblk((int64_t)n_frames, TPB), TPB, 0, rb3b_stream>>>((int64_t)n_frames, ibits, nb, ex.p, frames.p); CKK();
	CK(cudaMemcpyAsync(hflags, flags.p, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	const size_t img = 80 + (size_t)n_bytes + (size_t)n_frames * (RB3B_ASIZE + 1) * 8;
	uint8_t *o = (uint8_t*)malloc(img);
	if (o == 0) return rb3b_fail(RB3B_ENOMEM, "out of host memory for the .fmd image (%zu bytes)", img);
	cudaError_t e1 = cudaMemcpyAsync(o + 80, body.p, (size_t)n_bytes, cudaMemcpyDeviceToHost, rb3b_stream);
	cudaError_t e2 = cudaMemcpyAsync(o + 80 + n_bytes, frames.p, (size_t)n_frames * (RB3B_ASIZE + 1) * 8, cudaMemcpyDeviceToHost, rb3b_stream);
	cudaError_t e3 = cudaStreamSynchronize(rb3b_stream);
	if (e1 != cudaSuccess || e2 != cudaSuccess || e3 != cudaSuccess) { free(o); return rb3b_fail(RB3B_ENODEV, "copying the .fmd image to the host failed"); }
	if (hflags[1]) { free(o); return 0; }
	const char magic[4] = { 'R', 'L', 'D', 3 };
	const uint32_t geom = RB3B_ASIZE << 16 | 3;
	const uint64_t zero = 0;
	memcpy(o, magic, 4); memcpy(o + 4, &geom, 4); memcpy(o + 8, &zero, 8); memcpy(o + 16, &n_bytes, 8); memcpy(o + 24, &n_frames, 8);
	memcpy(o + 32, tot_sym, RB3B_ASIZE * 8);
	rb3b_stat_set("fmd_device_blocks", nb);
	*out = o;
	return (int64_t)img;
}
