/*
 * rb3b_io.cu -- host-side readers/writers of the on-disk formats that the merge
 * path's callers exchange: .fmd (rld0.c:107-243), .fmr
 * (mrope.c:152-177, rope.c:265-330, rle.h:39-75) and the plain-text dump
 * (mrope.c:195-210).  The .fmd written here is byte-identical to the
 * reference's for the same run sequence; the .fmr is a legal tree the reference
 * can load and keep inserting into (its bytes depend on batching even in the
 * reference, README.md:170-171).
 *
 * Everything works on the canonical run list exported from the device index.
 */
#include <string.h>
#include <time.h>
#include <stdlib.h>
#include <string>
#include <vector>
#include <thread>
#include <algorithm>
#include <functional>
#include <memory>
#include "rb3b_internal.cuh"

typedef std::vector<uint8_t> bytes_t;



static inline int ilog2_u64(uint64_t v) { return v ? 63 - __builtin_clzll(v) : -1; }

/* ------------------------------------------------------------------ */
/* FMD writer                                                           */
/* ------------------------------------------------------------------ */

namespace {

const int FMD_SSIZE = 8;                    /* words per small block (sbits = 3, fm-index.c:18) */
const int64_t FMD_LSIZE = 1LL << 23;        /* words per chunk (rld0.h:11-13) */
const int FMD_HDR_WORDS[3] = { 2, 4, 7 };   /* rld0.c:71-73 */

class FmdWriter {
public:
	FmdWriter() : head_(0), cur_(2), free_(64), pend_sym_(-1), pend_len_(0)
	{
		words_.assign(2 * FMD_SSIZE, 0);
		tail_ = tail_of(0);
		memset(run_, 0, sizeof(run_)); memset(mark_, 0, sizeof(mark_));
	}
	/* rld_enc (rld0.c:153-161): neighbours with the same symbol are fused before coding */
	void put(int sym, int64_t len)
	{
		if (len <= 0) return;
		if (sym == pend_sym_) { pend_len_ += len; return; }
		if (pend_len_) code(pend_len_, pend_sym_);
		pend_sym_ = sym; pend_len_ = len;
	}
	/* rld_enc_finish + rld_rank_index + rld_dump (rld0.c:163-243) */
	void finish(bytes_t &out)
	{
		if (pend_len_) code(pend_len_, pend_sym_);
		pend_len_ = 0;
		open_block(); /* trailing header-only block carries the last data block's counts */
		const uint64_t n_words = (uint64_t)cur_, n_bytes = n_words * 8, total = mark_[0];
		std::vector<uint64_t> frames;
		uint64_t n_frames = build_frames(n_words, total, frames);
		out.clear();
		out.reserve(80 + n_bytes + frames.size() * 8);
		const char magic[4] = { 'R', 'L', 'D', 3 };
		append(out, magic, 4);
		uint32_t geom = RB3B_ASIZE << 16 | 3;
		append(out, &geom, 4);
		uint64_t zero = 0;
		append(out, &zero, 8); append(out, &n_bytes, 8); append(out, &n_frames, 8);
		append(out, mark_ + 1, RB3B_ASIZE * 8);
		append(out, words_.data(), n_bytes);
		append(out, frames.data(), frames.size() * 8);
	}
private:
	std::vector<uint64_t> words_;
	int64_t head_, cur_, tail_;  /* first word of the open block, word being filled, last usable word */
	int free_;                   /* unused low bits of words_[cur_] */
	uint64_t run_[RB3B_ASIZE + 1], mark_[RB3B_ASIZE + 1]; /* totals now / at the start of the open block; [0] = all */
	int pend_sym_; int64_t pend_len_;

	static void append(bytes_t &o, const void *p, size_t n) { const uint8_t *q = (const uint8_t*)p; o.insert(o.end(), q, q + n); }
	static int64_t tail_of(int64_t head)
	{ /* the last block of a 2^23-word chunk gives up one more word (rld0.h:81) */
		return head + FMD_SSIZE - (((head + FMD_SSIZE) & (FMD_LSIZE - 1)) == 0 ? 2 : 1);
	}
	void open_block()
	{ /* enc_next_block, rld0.c:107-135 */
		uint64_t delta[RB3B_ASIZE + 1];
		head_ += FMD_SSIZE;
		if ((size_t)(head_ + 2 * FMD_SSIZE) > words_.size()) words_.resize(words_.size() * 2 > (size_t)(head_ + 2 * FMD_SSIZE) ? words_.size() * 2 : head_ + 2 * FMD_SSIZE, 0);
		for (int i = 0; i <= RB3B_ASIZE; ++i) delta[i] = run_[i] - mark_[i];
		int type = delta[0] < 0x4000 ? 0 : delta[0] < 0x40000000 ? 1 : 2;
		uint8_t *dst = (uint8_t*)&words_[head_];
		for (int i = 0; i <= RB3B_ASIZE; ++i) {
			if (type == 0) { uint16_t v = (uint16_t)delta[i]; memcpy(dst + 2 * i, &v, 2); }
			else if (type == 1) { uint32_t v = (uint32_t)delta[i]; memcpy(dst + 4 * i, &v, 4); }
			else memcpy(dst + 8 * i, &delta[i], 8);
		}
		words_[head_] |= (uint64_t)type << 62;
		cur_ = head_ + FMD_HDR_WORDS[type]; free_ = 64; tail_ = tail_of(head_);
		memcpy(mark_, run_, sizeof(run_));
	}
	void code(int64_t len, int sym)
	{ /* rld_delta_enc1 + rld_enc1, rld0.c:45-51,137-151 */
		int y = ilog2_u64((uint64_t)len), z = ilog2_u64((uint64_t)y + 1);
		int width = 2 * z + 1 + y + 3;
		uint64_t bits = ((((uint64_t)len ^ (1ULL << y)) | (uint64_t)(y + 1) << y) << 3) | (uint64_t)sym;
		if (width >= free_ && cur_ == tail_) open_block();
		if (width > free_) {
			int spill = width - free_;
			words_[cur_++] |= bits >> spill;
			free_ = 64 - spill;
			words_[cur_] = bits << free_;
		} else {
			free_ -= width;
			words_[cur_] |= bits << free_;
		}
		run_[0] += len; run_[sym + 1] += len;
	}
	uint64_t build_frames(uint64_t n_words, uint64_t total, std::vector<uint64_t> &fr)
	{ /* rld_rank_index, rld0.c:163-204 */
		const int W = RB3B_ASIZE + 1;
		uint64_t n_blks = n_words / FMD_SSIZE + 1, last = n_words / FMD_SSIZE * FMD_SSIZE;
		int ibits = ilog2_u64(total / n_blks) + 4;
		uint64_t n_frames = ((total + (1ULL << ibits) - 1) >> ibits) + 1, k = 1, sofar[RB3B_ASIZE];
		fr.assign(n_frames * W, 0);
		memset(sofar, 0, sizeof(sofar));
		for (uint64_t i = FMD_SSIZE; i <= last; i += FMD_SSIZE) {
			const uint8_t *src = (const uint8_t*)&words_[i];
			int type = (int)(words_[i] >> 62);
			uint64_t sum = 0;
			for (int j = 1; j <= RB3B_ASIZE; ++j) {
				if (type == 0) { uint16_t v; memcpy(&v, src + 2 * j, 2); sofar[j - 1] += v; }
				else if (type == 1) { uint32_t v; memcpy(&v, src + 4 * j, 4); sofar[j - 1] += v & 0x3fffffffu; }
				else { uint64_t v; memcpy(&v, src + 8 * j, 8); sofar[j - 1] += v; }
				sum += sofar[j - 1];
			}
			while (sum >= k << ibits) ++k;
			if (k < n_frames) {
				fr[k * W] = i;
				memcpy(&fr[k * W + 1], sofar, sizeof(sofar));
			}
		}
		for (k = 1; k < n_frames; ++k)
			if (fr[k * W] == 0) memcpy(&fr[k * W], &fr[(k - 1) * W], W * 8);
		return n_frames;
	}
};


/* ------------------------------------------------------------------ */
/* parallel FMD writer                                                  */
/* ------------------------------------------------------------------ */
/*
 * The writer above is the reference's greedy bit packer (rld_enc1, rld0.c:137-151): a run's code goes into the open
 * block iff the payload bits used so far plus its width stay BELOW the block's payload capacity (the check only fires in
 * the last usable word, and a code is < 64 bits wide, so this is equivalent), and the header width of a block depends on
 * the number of symbols in the block before it.  Both only need prefix sums of code widths and run lengths, so:
 *   1. widths and the two prefix sums                       -- all threads
 *   2. block boundaries, one bounded binary search per block -- one thread (blocks are ~60 runs, so this is ~2 % of the work)
 *   3. headers and payload bits of disjoint block ranges     -- all threads
 *   4. the rank index (frames) from the headers              -- one thread, as before
 * The image is byte-identical to FmdWriter's (tests/test_host.py).  A block with >= 2^30 symbols (64-bit headers) makes
 * the whole call fall back to the sequential writer: that corner is exercised too rarely to deserve a second restatement.
 */
static inline int fmd_code_width(int64_t len) { int y = ilog2_u64((uint64_t)len), z = ilog2_u64((uint64_t)y + 1); return 2 * z + 1 + y + 3; }

static void par_for(int n_threads, int64_t n, const std::function<void(int, int64_t, int64_t)> &f)
{
	if (n_threads <= 1 || n < 2) { f(0, 0, n); return; }
	std::vector<std::thread> th;
	for (int t = 0; t < n_threads; ++t) {
		int64_t a = n * t / n_threads, b = n * (t + 1) / n_threads;
		th.emplace_back([=, &f]() { f(t, a, b); });
	}
	for (auto &x : th) x.join();
}

/* sym/len: run list.  Returns false when the sequential writer must be used (list not canonical, 64-bit headers). */
static bool fmd_encode_parallel(const uint8_t *sym, const int64_t *len, int64_t R, int n_threads, bytes_t &out)
{
	if (R < 2) return false;
	struct timespec ts_; clock_gettime(CLOCK_MONOTONIC, &ts_); double t0_ = ts_.tv_sec + 1e-9 * ts_.tv_nsec;
#define FMD_T(name) do { clock_gettime(CLOCK_MONOTONIC, &ts_); double t1_ = ts_.tv_sec + 1e-9 * ts_.tv_nsec; rb3b_stat_set("us_fmd_" name, (int64_t)((t1_ - t0_) * 1e6)); t0_ = t1_; } while (0)
	/* 1. code widths (one byte per run), symbol totals, and the check that the list is canonical */
	std::unique_ptr<uint8_t[]> wid(new uint8_t[(size_t)R]);
	std::vector<uint64_t> tsym((size_t)n_threads * RB3B_ASIZE, 0);
	std::vector<int> tbad(n_threads, 0);
	par_for(n_threads, R, [&](int t, int64_t a, int64_t b) {
		uint64_t c[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
		int bad = 0;
		for (int64_t j = a; j < b; ++j) {
			if (len[j] <= 0 || sym[j] >= RB3B_ASIZE || (j > 0 && sym[j] == sym[j - 1])) { bad = 1; break; }
			wid[j] = (uint8_t)fmd_code_width(len[j]);
			c[sym[j]] += (uint64_t)len[j];
		}
		tbad[t] = bad;
		for (int i = 0; i < RB3B_ASIZE; ++i) tsym[(size_t)t * RB3B_ASIZE + i] = c[i];
	});
	uint64_t tot_sym[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0}, total = 0;
	for (int t = 0; t < n_threads; ++t) {
		if (tbad[t]) return false;
		for (int i = 0; i < RB3B_ASIZE; ++i) { tot_sym[i] += tsym[(size_t)t * RB3B_ASIZE + i]; total += tsym[(size_t)t * RB3B_ASIZE + i]; }
	}
	FMD_T("widths");
	/* 2. block boundaries: one sequential scan over 9 bytes per run */
	std::vector<int64_t> bstart;           /* first run of block b */
	std::vector<uint8_t> btype;            /* header type of block b */
	bstart.reserve((size_t)(R / 40 + 16)); btype.reserve(bstart.capacity());
	{
		int64_t j = 0, head = 0;
		int type = 0;
		while (j < R) {
			const int64_t tail = head + FMD_SSIZE - (((head + FMD_SSIZE) & (FMD_LSIZE - 1)) == 0 ? 2 : 1);
			const int64_t cap = (tail + 1 - (head + FMD_HDR_WORDS[type])) * 64;
			if (type == 2 || cap <= 0) return false;
			bstart.push_back(j); btype.push_back((uint8_t)type);
			int64_t used = 0;
			uint64_t nsym = 0;
			while (j < R && used + wid[j] < cap) { used += wid[j]; nsym += (uint64_t)len[j]; ++j; } /* the first run always fits: a code is < 64 bits */
			type = nsym < 0x4000 ? 0 : nsym < 0x40000000 ? 1 : 2;
			head += FMD_SSIZE;
		}
		bstart.push_back(R); btype.push_back((uint8_t)type); /* trailing header-only block (rld_enc_finish, rld0.c:206-216) */
		if (type == 2) return false;
	}
	FMD_T("bounds");
	const int64_t n_blk = (int64_t)bstart.size(); /* including the trailing one */
	const uint64_t n_words = (uint64_t)(n_blk - 1) * FMD_SSIZE + FMD_HDR_WORDS[btype[n_blk - 1]];
	const uint64_t n_bytes = n_words * 8;
	const uint64_t n_blks = n_words / FMD_SSIZE + 1, last = n_words / FMD_SSIZE * FMD_SSIZE;
	const int ibits = ilog2_u64(total / n_blks) + 4;
	const uint64_t n_frames = ((total + (1ULL << ibits) - 1) >> ibits) + 1;
	const int W = RB3B_ASIZE + 1;
	/* the image is assembled in place: header, then whole blocks (zero-filled, the threads OR their bits in), frames last */
	bytes_t().swap(out);
	out.reserve(80 + (size_t)n_blk * FMD_SSIZE * 8 + (size_t)n_frames * W * 8);
	out.resize(80 + (size_t)n_blk * FMD_SSIZE * 8);
	uint64_t *words = (uint64_t*)(out.data() + 80); /* 80 is a multiple of 8 and vector storage is suitably aligned */
	/* 3. headers + payload */
	par_for(n_threads, n_blk, [&](int, int64_t b0, int64_t b1) {
		for (int64_t b = b0; b < b1; ++b) {
			uint64_t *blk = words + (size_t)b * FMD_SSIZE;
			const int type = btype[b];
			if (b > 0) { /* header = counts of the previous block (enc_next_block, rld0.c:107-135) */
				uint64_t delta[RB3B_ASIZE + 1] = {0, 0, 0, 0, 0, 0, 0};
				for (int64_t j = bstart[b - 1]; j < bstart[b]; ++j) { delta[0] += (uint64_t)len[j]; delta[sym[j] + 1] += (uint64_t)len[j]; }
				uint8_t *dst = (uint8_t*)blk;
				for (int i = 0; i <= RB3B_ASIZE; ++i) {
					if (type == 0) { uint16_t v = (uint16_t)delta[i]; memcpy(dst + 2 * i, &v, 2); }
					else { uint32_t v = (uint32_t)delta[i]; memcpy(dst + 4 * i, &v, 4); }
				}
				blk[0] |= (uint64_t)type << 62;
			}
			if (b == n_blk - 1) break; /* trailing block: header only */
			int cur = FMD_HDR_WORDS[type], fr = 64;
			for (int64_t j = bstart[b]; j < bstart[b + 1]; ++j) { /* rld_delta_enc1 + rld_enc1 */
				const int y = ilog2_u64((uint64_t)len[j]), width = wid[j];
				const uint64_t bits = ((((uint64_t)len[j] ^ (1ULL << y)) | (uint64_t)(y + 1) << y) << 3) | (uint64_t)sym[j];
				if (width > fr) {
					const int spill = width - fr;
					blk[cur++] |= bits >> spill;
					fr = 64 - spill;
					blk[cur] = bits << fr;
				} else {
					fr -= width;
					blk[cur] |= bits << fr;
				}
			}
		}
	});
	FMD_T("encode");
	/* 4. frames (rld_rank_index, rld0.c:163-204) and the header of the file */
	std::vector<uint64_t> fr((size_t)n_frames * W, 0);
	{
		uint64_t k = 1, sofar[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
		for (uint64_t i = FMD_SSIZE; i <= last; i += FMD_SSIZE) {
			const uint8_t *src = (const uint8_t*)&words[i];
			const int type = (int)(words[i] >> 62);
			uint64_t sum = 0;
			for (int j = 1; j <= RB3B_ASIZE; ++j) {
				if (type == 0) { uint16_t v; memcpy(&v, src + 2 * j, 2); sofar[j - 1] += v; }
				else { uint32_t v; memcpy(&v, src + 4 * j, 4); sofar[j - 1] += v & 0x3fffffffu; }
				sum += sofar[j - 1];
			}
			while (sum >= k << ibits) ++k;
			if (k < n_frames) { fr[k * W] = i; memcpy(&fr[k * W + 1], sofar, sizeof(sofar)); }
		}
		for (k = 1; k < n_frames; ++k)
			if (fr[k * W] == 0) memcpy(&fr[k * W], &fr[(k - 1) * W], W * 8);
	}
	out.resize(80 + n_bytes); /* drops the unused tail of the trailing block */
	uint8_t *o = out.data();
	const char magic[4] = { 'R', 'L', 'D', 3 };
	const uint32_t geom = RB3B_ASIZE << 16 | 3;
	const uint64_t zero = 0;
	memcpy(o, magic, 4); memcpy(o + 4, &geom, 4); memcpy(o + 8, &zero, 8); memcpy(o + 16, &n_bytes, 8); memcpy(o + 24, &n_frames, 8);
	memcpy(o + 32, tot_sym, RB3B_ASIZE * 8);
	out.insert(out.end(), (const uint8_t*)fr.data(), (const uint8_t*)fr.data() + fr.size() * 8);
	FMD_T("frames_image");
#undef FMD_T
	return true;
}

/* run list (not necessarily canonical) -> .fmd image; large lists go through the parallel writer */
static void fmd_encode_any(const uint8_t *sym, const int64_t *len, int64_t n_runs, bytes_t &img)
{
	int n_threads = (int)rb3b_get_param("fmd_threads", 0);
	if (n_threads <= 0) { n_threads = (int)std::thread::hardware_concurrency(); if (n_threads > 32) n_threads = 32; if (n_threads < 1) n_threads = 1; }
	if (n_runs >= rb3b_get_param("fmd_parallel_min_runs", 1 << 20) && fmd_encode_parallel(sym, len, n_runs, n_threads, img)) return;
	FmdWriter w;
	for (int64_t i = 0; i < n_runs; ++i) w.put(sym[i], len[i]);
	w.finish(img);
}

/* ------------------------------------------------------------------ */
/* FMD reader (rld0.h:85-125 restated on a flat image)                  */
/* ------------------------------------------------------------------ */

struct RunList {
	std::vector<uint8_t> sym; std::vector<int64_t> len;
	void add(int c, int64_t l)
	{
		if (l <= 0) return;
		if (!sym.empty() && sym.back() == c) len.back() += l;
		else { sym.push_back((uint8_t)c); len.push_back(l); }
	}
};

/* 64 bits of the block's payload starting at bit offset `bit`; zero past the last usable word */
static uint64_t payload_bits(const uint64_t *w, int64_t first, int64_t tail, int64_t bit)
{
	int64_t wi = first + (bit >> 6);
	int sh = (int)(bit & 63);
	if (wi > tail) return 0;
	uint64_t x = w[wi] << sh;
	if (sh && wi < tail) x |= w[wi + 1] >> (64 - sh);
	return x;
}

static int host_threads(void)
{
	int n = (int)rb3b_get_param("fmd_threads", 0);
	if (n <= 0) { n = (int)std::thread::hardware_concurrency(); if (n > 32) n = 32; if (n < 1) n = 1; }
	return n;
}

/* append b to a, fusing the junction */
static void runs_append(RunList &a, const RunList &b)
{
	size_t i = 0;
	if (!a.sym.empty() && !b.sym.empty() && a.sym.back() == b.sym[0]) { a.len.back() += b.len[0]; i = 1; }
	a.sym.insert(a.sym.end(), b.sym.begin() + i, b.sym.end());
	a.len.insert(a.len.end(), b.len.begin() + i, b.len.end());
}

/* the runs of blocks [h0, h1) (word offsets, multiples of FMD_SSIZE) */
static void fmd_decode_blocks(const uint64_t *w, int64_t h0, int64_t h1, RunList &runs)
{
	runs.sym.reserve((size_t)(h1 - h0) / FMD_SSIZE * 40); runs.len.reserve((size_t)(h1 - h0) / FMD_SSIZE * 40);
	int last = -1; /* symbol of the last run added, to fuse equal neighbours (blocks cut runs nowhere, but stay safe) */
	if (!runs.sym.empty()) last = runs.sym.back();
	for (int64_t h = h0; h < h1; h += FMD_SSIZE) {
		int64_t first = h + FMD_HDR_WORDS[w[h] >> 62 > 2 ? 2 : w[h] >> 62], tail = h + FMD_SSIZE - (((h + FMD_SSIZE) & (FMD_LSIZE - 1)) == 0 ? 2 : 1), bit = 0;
		for (;;) {
			/* a code is at most 59 bits wide (rld_delta_enc1, rld0.c:45-51; runs < 2^43 plus 3 symbol bits): one 64-bit window holds it */
			const uint64_t x = payload_bits(w, first, tail, bit);
			if (x >> 58 == 0) break; /* six zero bits cannot start a code */
			const int z = __builtin_clzll(x);
			const int y = (int)(x << z >> (63 - z)) - 1;
			const int width = 2 * z + 1 + y + 3;
			uint64_t l; int c;
			if (width <= 64) {
				l = (y ? (x << (2 * z + 1)) >> (64 - y) : 0) | 1ULL << y;
				c = (int)((x << (2 * z + 1 + y)) >> 61);
			} else { /* wider than any code the encoders write: take the slow, general path */
				l = (y ? payload_bits(w, first, tail, bit + 2 * z + 1) >> (64 - y) : 0) | 1ULL << y;
				c = (int)(payload_bits(w, first, tail, bit + 2 * z + 1 + y) >> 61);
			}
			if (c >= RB3B_ASIZE) break;
			if (c == last) runs.len.back() += (int64_t)l;
			else { runs.sym.push_back((uint8_t)c); runs.len.push_back((int64_t)l); last = c; }
			bit += width;
		}
	}
}

static int fmd_parse(const bytes_t &img, RunList &runs)
{
	if (img.size() < 80 || memcmp(img.data(), "RLD\3", 4) != 0) return RB3B_EFORMAT;
	uint32_t geom; uint64_t n_bytes, n_frames;
	memcpy(&geom, &img[4], 4); memcpy(&n_bytes, &img[16], 8); memcpy(&n_frames, &img[24], 8);
	if (geom != (RB3B_ASIZE << 16 | 3)) return rb3b_fail(RB3B_EFORMAT, "FMD with asize/sbits 0x%x is not supported (only 6/3)", geom);
	if (n_bytes > img.size() - 80 || (n_bytes & 7) != 0) return rb3b_fail(RB3B_EFORMAT, "truncated or corrupt FMD (n_bytes = %llu, file holds %zu)", (unsigned long long)n_bytes, img.size() - 80);
	std::vector<uint64_t> w(n_bytes / 8 + FMD_SSIZE + 1, 0);
	memcpy(w.data(), &img[80], n_bytes);
	const int64_t n_words = n_bytes / 8, n_blk = n_words / FMD_SSIZE;
	/* the blocks are self-contained (rld0.h:85-125): every thread decodes a range of them, the pieces are joined in order */
	const int T = n_blk >= 4096 ? host_threads() : 1;
	std::vector<RunList> part(T);
	par_for(T, n_blk, [&](int t, int64_t b0, int64_t b1) { fmd_decode_blocks(w.data(), b0 * FMD_SSIZE, b1 * FMD_SSIZE, part[t]); });
	for (int t = 0; t < T; ++t) {
		if (t == 0) { runs.sym.swap(part[0].sym); runs.len.swap(part[0].len); }
		else runs_append(runs, part[t]);
	}
	/* the decoded symbol totals must be the marginal counts of the header (rld0.c:230-236): catches corrupt payloads */
	uint64_t got[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0}, want[RB3B_ASIZE];
	memcpy(want, &img[32], sizeof(want));
	for (size_t i = 0; i < runs.sym.size(); ++i) got[runs.sym[i]] += (uint64_t)runs.len[i];
	for (int a = 0; a < RB3B_ASIZE; ++a)
		if (got[a] != want[a]) return rb3b_fail(RB3B_EFORMAT, "corrupt FMD: decoded %llu symbols of code %d, the header says %llu", (unsigned long long)got[a], a, (unsigned long long)want[a]);
	return RB3B_OK;
}

/* ------------------------------------------------------------------ */
/* FMR codec                                                            */
/* ------------------------------------------------------------------ */

static int rle_put(uint8_t *p, int c, int64_t l)
{ /* rle.h:53-75 */
	if (l < 16) { p[0] = (uint8_t)(l << 3 | c); return 1; }
	if (l < 256) { p[0] = (uint8_t)(0xC0 | (l >> 6) << 3 | c); p[1] = (uint8_t)(0x80 | (l & 0x3f)); return 2; }
	int n = l < (1LL << 19) ? 4 : 8;
	p[0] = (uint8_t)((n == 4 ? 0xE0 : 0xF0) | (l >> (6 * (n - 1))) << 3 | c);
	for (int i = 1; i < n; ++i) p[i] = (uint8_t)(0x80 | ((l >> (6 * (n - 1 - i))) & 0x3f));
	return n;
}

static const int64_t RLE_MAX_RUN = (1LL << 43) - 1; /* rle.h:68-73 */

/* the leaves of one rope: run codes back to back in `code`, one LeafMeta per leaf */
struct LeafMeta { int64_t cnt[RB3B_ASIZE]; size_t off; uint16_t nb; };
struct RopeImg { std::vector<uint8_t> code; std::vector<LeafMeta> leaf; bytes_t out; };

static void fmr_node(bytes_t &o, const RopeImg &R, size_t lo, size_t hi, int height, int fan)
{ /* pre-order node record, rope.c:265-280 */
	size_t per = 1;
	for (int i = 0; i < height; ++i) per *= fan;
	size_t n_child = (hi - lo + per - 1) / per;
	uint8_t is_bottom = height == 0;
	int16_t n = (int16_t)n_child;
	o.push_back(is_bottom);
	o.insert(o.end(), (uint8_t*)&n, (uint8_t*)&n + 2);
	for (size_t i = 0; i < n_child; ++i) {
		/* spread the leaves evenly over the children */
		size_t a = lo + (hi - lo) * i / n_child, b = lo + (hi - lo) * (i + 1) / n_child;
		if (is_bottom) {
			const LeafMeta &L = R.leaf[a];
			o.insert(o.end(), (const uint8_t*)L.cnt, (const uint8_t*)L.cnt + 48);
			o.insert(o.end(), (const uint8_t*)&L.nb, (const uint8_t*)&L.nb + 2);
			o.insert(o.end(), R.code.begin() + L.off, R.code.begin() + L.off + L.nb);
		} else fmr_node(o, R, a, b, height - 1, fan);
	}
}

/* rope a = rows [lo, hi) of the BWT (mrope.c:157, fm-index.c:72-81), starting `used` symbols into run ri */
static void fmr_rope(const uint8_t *sym, const int64_t *len, int64_t n_runs, int64_t ri, int64_t used, int64_t lo, int64_t hi, int max_nodes, int block_len, RopeImg &R)
{
	const int leaf_cap = block_len - 18 - 16; /* keep clear of the split trigger nbytes + 18 > block_len (rope.c:143) */
	const int fan = max_nodes > 4 ? max_nodes / 2 : 2; /* half-full nodes: the CPU code splits full ones on the way down */
	int psym = -1;
	int64_t plen = 0, pos = lo;
	{ /* a guess of the code size: ~1.3 bytes per run, a share of the runs proportional to the rope's share of the rows is too crude, so cap it */
		int64_t guess = (hi - lo) / 4 + 64, cap = n_runs / 2 + 64;
		R.code.resize((size_t)(guess < cap ? guess : cap));
	}
	size_t n_code = 0; /* bytes of R.code in use; the vector is grown in large steps and trimmed at the end */
	LeafMeta *cur = 0;
	auto emit = [&](int c, int64_t left) { /* one (fused) run of the rope, cut into codable pieces */
		while (left > 0) {
			const int64_t l = left < RLE_MAX_RUN ? left : RLE_MAX_RUN;
			if (n_code + 8 > R.code.size()) R.code.resize(R.code.size() + R.code.size() / 2 + 4096);
			const int nb = rle_put(&R.code[n_code], c, l);
			if (cur == 0 || (int)cur->nb + nb > leaf_cap) {
				LeafMeta m;
				memset(m.cnt, 0, sizeof(m.cnt)); m.off = n_code; m.nb = 0;
				R.leaf.push_back(m);
				cur = &R.leaf.back();
			}
			n_code += nb;
			cur->nb = (uint16_t)(cur->nb + nb);
			cur->cnt[c] += l;
			left -= l;
		}
	};
	while (pos < hi) {
		int64_t t = len[ri] - used;
		if (t > hi - pos) t = hi - pos;
		if (t > 0) { /* neighbours with the same symbol are fused, empty runs dropped (the list need not be canonical) */
			if (sym[ri] == psym) plen += t;
			else { if (plen) emit(psym, plen); psym = sym[ri]; plen = t; }
		}
		used += t; pos += t;
		if (used == len[ri]) ++ri, used = 0;
	}
	if (plen) emit(psym, plen);
	R.code.resize(n_code);
	if (R.leaf.empty()) { LeafMeta m; memset(m.cnt, 0, sizeof(m.cnt)); m.off = 0; m.nb = 0; R.leaf.push_back(m); } /* empty rope: one empty leaf (rope.c:64-67) */
	int32_t mn = max_nodes, bl = block_len;
	R.out.reserve(R.code.size() + R.leaf.size() * 52 + 64);
	R.out.insert(R.out.end(), (uint8_t*)&mn, (uint8_t*)&mn + 4);
	R.out.insert(R.out.end(), (uint8_t*)&bl, (uint8_t*)&bl + 4);
	int height = 0;
	for (size_t cap = fan; cap < R.leaf.size(); cap *= fan) ++height;
	fmr_node(R.out, R, 0, R.leaf.size(), height, fan);
}

static void fmr_encode(const uint8_t *sym, const int64_t *len, int64_t n_runs, int max_nodes, int block_len, bytes_t &o)
{
	int64_t acc[RB3B_ASIZE + 1] = {0, 0, 0, 0, 0, 0, 0};
	for (int64_t i = 0; i < n_runs; ++i) acc[sym[i] + 1] += len[i] > 0 ? len[i] : 0;
	for (int a = 0; a < RB3B_ASIZE; ++a) acc[a + 1] += acc[a];
	/* where every rope starts in the run list */
	int64_t start_ri[RB3B_ASIZE], start_used[RB3B_ASIZE];
	{
		int64_t pos = 0, ri = 0;
		for (int a = 0; a < RB3B_ASIZE; ++a) {
			while (ri < n_runs && pos + (len[ri] > 0 ? len[ri] : 0) <= acc[a] && !(pos == acc[a] && len[ri] > 0)) { pos += len[ri] > 0 ? len[ri] : 0; ++ri; }
			start_ri[a] = ri; start_used[a] = acc[a] - pos;
		}
	}
	RopeImg R[RB3B_ASIZE];
	int n_threads = (int)rb3b_get_param("fmd_threads", 0);
	auto work = [&](int a) { if (acc[a + 1] > acc[a]) fmr_rope(sym, len, n_runs, start_ri[a], start_used[a], acc[a], acc[a + 1], max_nodes, block_len, R[a]);
	                         else fmr_rope(sym, len, n_runs, 0, 0, 0, 0, max_nodes, block_len, R[a]); };
	if (n_threads == 1 || n_runs < (1 << 16)) { for (int a = 0; a < RB3B_ASIZE; ++a) work(a); }
	else { /* the six ropes are independent */
		std::vector<std::thread> th;
		for (int a = 0; a < RB3B_ASIZE; ++a) th.emplace_back(work, a);
		for (auto &t : th) t.join();
	}
	const uint8_t hdr[4] = { 'R', 'B', 2, 0 }; /* sorting order 0: input order (rb3b_dump_fmr patches it) */
	size_t tot = 4;
	for (int a = 0; a < RB3B_ASIZE; ++a) tot += R[a].out.size();
	o.clear(); o.reserve(tot);
	o.insert(o.end(), hdr, hdr + 4);
	for (int a = 0; a < RB3B_ASIZE; ++a) o.insert(o.end(), R[a].out.begin(), R[a].out.end());
}

struct FmrCursor { const uint8_t *p, *end; };
struct LeafRef { const uint8_t *p; uint16_t nb; };

/* walk the node records (rope.c:289-317) and list the leaves in order; nothing is decoded here */
static int fmr_collect(FmrCursor &c, std::vector<LeafRef> &leaves, int depth)
{
	if (c.p + 3 > c.end || depth > 64) return RB3B_EFORMAT;
	uint8_t is_bottom = c.p[0];
	int16_t n; memcpy(&n, c.p + 1, 2);
	c.p += 3;
	for (int i = 0; i < n; ++i) {
		if (!is_bottom) { int rc = fmr_collect(c, leaves, depth + 1); if (rc) return rc; continue; }
		if (c.p + 50 > c.end) return RB3B_EFORMAT;
		uint16_t nb; memcpy(&nb, c.p + 48, 2);
		c.p += 50;
		if (c.p + nb > c.end) return RB3B_EFORMAT;
		LeafRef r = { c.p, nb };
		leaves.push_back(r);
		c.p += nb;
	}
	return RB3B_OK;
}

static int fmr_decode_leaves(const LeafRef *lv, int64_t n, RunList &runs)
{
	for (int64_t i = 0; i < n; ++i)
		for (const uint8_t *q = lv[i].p, *e = lv[i].p + lv[i].nb; q < e;) { /* rle_dec1, rle.h:39-51 */
			const int sym = q[0] & 7, n_byte = (q[0] & 0x80) == 0 ? 1 : q[0] >> 5 == 6 ? 2 : (q[0] & 0x10) ? 8 : 4;
			int64_t l;
			if (sym >= RB3B_ASIZE || q + n_byte > e) return RB3B_EFORMAT; /* before the continuation bytes are read */
			if (n_byte == 1) l = q[0] >> 3;
			else if (n_byte == 2) l = (int64_t)(q[0] & 0x18) << 3 | (q[1] & 0x3f);
			else {
				l = q[0] >> 3 & 1;
				for (int j = 1; j < n_byte; ++j) l = l << 6 | (q[j] & 0x3f);
			}
			runs.add(sym, l);
			q += n_byte;
		}
	return RB3B_OK;
}

static int fmr_parse(const bytes_t &img, RunList &runs)
{
	if (img.size() < 4 || memcmp(img.data(), "RB\2", 3) != 0) return RB3B_EFORMAT;
	FmrCursor c = { img.data() + 4, img.data() + img.size() };
	std::vector<LeafRef> leaves;
	for (int a = 0; a < RB3B_ASIZE; ++a) {
		if (c.p + 8 > c.end) return rb3b_fail(RB3B_EFORMAT, "truncated FMR");
		c.p += 8; /* max_nodes, block_len */
		if (fmr_collect(c, leaves, 0) != RB3B_OK) return rb3b_fail(RB3B_EFORMAT, "corrupt FMR node record");
	}
	const int64_t n = (int64_t)leaves.size();
	const int T = n >= 4096 ? host_threads() : 1;
	std::vector<RunList> part(T);
	std::vector<int> rc(T, RB3B_OK);
	par_for(T, n, [&](int t, int64_t a, int64_t b) { rc[t] = fmr_decode_leaves(leaves.data() + a, b - a, part[t]); });
	for (int t = 0; t < T; ++t) {
		if (rc[t] != RB3B_OK) return rb3b_fail(RB3B_EFORMAT, "corrupt FMR leaf");
		if (t == 0) { runs.sym.swap(part[0].sym); runs.len.swap(part[0].len); }
		else runs_append(runs, part[t]);
	}
	return RB3B_OK;
}

/* ------------------------------------------------------------------ */

static int read_file(const char *fn, bytes_t &out)
{
	FILE *fp = strcmp(fn, "-") ? fopen(fn, "rb") : stdin;
	if (!fp) return rb3b_fail(RB3B_EIO, "failed to open '%s' for reading", fn);
	uint8_t buf[1 << 16];
	size_t n;
	out.clear();
	while ((n = fread(buf, 1, sizeof(buf), fp)) > 0) out.insert(out.end(), buf, buf + n);
	if (fp != stdin) fclose(fp);
	return RB3B_OK;
}

static int write_file(const char *fn, const void *p, size_t n)
{
	FILE *fp = strcmp(fn, "-") ? fopen(fn, "wb") : stdout;
	if (!fp) return rb3b_fail(RB3B_EIO, "failed to open '%s' for writing", fn);
	size_t w = fwrite(p, 1, n, fp);
	if (fp != stdout) fclose(fp); else fflush(fp);
	return w == n ? RB3B_OK : rb3b_fail(RB3B_EIO, "short write to '%s'", fn);
}

static int fetch_runs(const rb3b_index_t *x, std::vector<uint8_t> &sym, std::vector<int64_t> &len)
{
	int64_t n = rb3b_export_runs(x, 0, 0, 0);
	if (n < 0) return (int)n;
	sym.resize(n); len.resize(n);
	if (n == 0) return RB3B_OK;
	int64_t m = rb3b_export_runs(x, sym.data(), len.data(), n);
	return m < 0 ? (int)m : RB3B_OK;
}

} /* namespace */

/* in-memory variants used by the tests and the CLI */
extern "C" int64_t rb3b_fmd_image(int64_t n_runs, const uint8_t *sym, const int64_t *len, uint8_t **out)
{
	bytes_t img;
	fmd_encode_any(sym, len, n_runs, img);
	*out = (uint8_t*)malloc(img.size() ? img.size() : 1);
	memcpy(*out, img.data(), img.size());
	return (int64_t)img.size();
}

extern "C" int64_t rb3b_fmr_image(int64_t n_runs, const uint8_t *sym, const int64_t *len, int max_nodes, int block_len, uint8_t **out)
{
	bytes_t img;
	fmr_encode(sym, len, n_runs, max_nodes > 0 ? max_nodes : 64, block_len > 0 ? block_len : 512, img);
	*out = (uint8_t*)malloc(img.size());
	memcpy(*out, img.data(), img.size());
	return (int64_t)img.size();
}

/* host-only reader behind rb3b_restore: .fmd or .fmr image -> canonical run list (malloc'd), returns the number of runs */
/* no C++ exception may cross the C ABI: allocation failures of the host codecs become error codes */
#define GUARD_BEGIN try {
#define GUARD_END } catch (const std::bad_alloc &) { return rb3b_fail(RB3B_ENOMEM, "out of host memory"); } \
	catch (const std::exception &e_) { return rb3b_fail(RB3B_EFORMAT, "malformed input (%s)", e_.what()); }

extern "C" int64_t rb3b_runs_from_image(const uint8_t *image, int64_t n_bytes, uint8_t **sym, int64_t **len, int *sorting_order)
{
	GUARD_BEGIN
	bytes_t img(image, image + (n_bytes > 0 ? n_bytes : 0));
	RunList runs;
	int rc, so = 0;
	if (img.size() >= 4 && memcmp(img.data(), "RLD\3", 4) == 0) rc = fmd_parse(img, runs);
	else if (img.size() >= 4 && memcmp(img.data(), "RB\2", 3) == 0) { rc = fmr_parse(img, runs); so = img[3] <= 2 ? img[3] : 0; }
	else return rb3b_fail(RB3B_EFORMAT, "neither FMD nor FMR");
	if (rc != RB3B_OK) return rc;
	const size_t n = runs.sym.size();
	*sym = (uint8_t*)malloc(n ? n : 1); *len = (int64_t*)malloc((n ? n : 1) * 8);
	if (*sym == 0 || *len == 0) return rb3b_fail(RB3B_ENOMEM, "out of host memory");
	if (n) { memcpy(*sym, runs.sym.data(), n); memcpy(*len, runs.len.data(), n * 8); }
	if (sorting_order) *sorting_order = so;
	return (int64_t)n;
	GUARD_END
}

extern "C" void rb3b_host_free(void *p) { free(p); }

int64_t rb3b_fmd_image_dev(int64_t R, const uint8_t *d_sym, const int64_t *d_len, const int64_t tot_sym[RB3B_ASIZE], uint8_t **out); /* rb3b_fmd_dev.cu */

extern "C" int rb3b_dump_fmd(const rb3b_index_t *x, const char *fn)
{
	GUARD_BEGIN
	if (rb3b_get_param("fmd_device", 1) != 0 && x->n > 0) { /* encode on the device: only the finished image crosses PCIe */
		ApiScope scope_;
		TRY(rb3b_ensure_init());
		DBuf<uint8_t> ds; DBuf<int64_t> dl;
		int64_t n_runs = 0;
		TRY(rb3b_export_runs_dev(x, ds, dl, &n_runs));
		if (n_runs >= rb3b_get_param("fmd_device_min_runs", 4096)) {
			uint8_t *img = 0;
			const int64_t sz = rb3b_fmd_image_dev(n_runs, ds.p, dl.p, x->tot, &img);
			if (sz < 0) return (int)sz;
			if (sz > 0) {
				const int rc = write_file(fn, img, (size_t)sz);
				free(img);
				rb3b_stat_set("fmd_encoded_on_device", 1);
				return rc;
			}
		}
	}
	rb3b_stat_set("fmd_encoded_on_device", 0);
	std::vector<uint8_t> sym; std::vector<int64_t> len;
	TRY(fetch_runs(x, sym, len));
	bytes_t img;
	fmd_encode_any(sym.data(), len.data(), (int64_t)sym.size(), img);
	return write_file(fn, img.data(), img.size());
	GUARD_END
}

extern "C" int rb3b_dump_fmr(const rb3b_index_t *x, const char *fn, int max_nodes, int block_len)
{
	GUARD_BEGIN
	std::vector<uint8_t> sym; std::vector<int64_t> len;
	TRY(fetch_runs(x, sym, len));
	if (max_nodes <= 0) max_nodes = 64;
	if (block_len <= 0) block_len = 512;
	if (block_len < 64 || (block_len & 7)) return rb3b_fail(RB3B_EINVAL, "block_len must be a multiple of 8 and >= 64 (rope.c:59-61)");
	bytes_t img;
	fmr_encode(sym.data(), len.data(), (int64_t)sym.size(), max_nodes, block_len, img);
	if (img.size() >= 4) img[3] = (uint8_t)x->so; /* mr_dump writes the sorting order after the magic, mrope.c:155-156 */
	return write_file(fn, img.data(), img.size());
	GUARD_END
}

extern "C" int rb3b_dump_plain(const rb3b_index_t *x, const char *fn)
{ /* mr_print_bwt, mrope.c:195-210 */
	std::vector<uint8_t> sym; std::vector<int64_t> len;
	TRY(fetch_runs(x, sym, len));
	FILE *fp = strcmp(fn, "-") ? fopen(fn, "wb") : stdout;
	if (!fp) return rb3b_fail(RB3B_EIO, "failed to open '%s' for writing", fn);
	std::string buf;
	for (size_t i = 0; i < sym.size(); ++i) {
		buf.assign((size_t)(len[i] < (1 << 20) ? len[i] : (1 << 20)), "$ACGTN"[sym[i]]);
		for (int64_t left = len[i]; left > 0; left -= (int64_t)buf.size())
			fwrite(buf.data(), 1, (size_t)(left < (int64_t)buf.size() ? left : (int64_t)buf.size()), fp);
	}
	fputc('\n', fp);
	if (fp != stdout) fclose(fp); else fflush(fp);
	return RB3B_OK;
}

extern "C" int rb3b_restore(rb3b_index_t *x, const char *fn)
{ /* rb3_fmi_restore (fm-index.h:123-133): FMD magic first, then FMR */
	GUARD_BEGIN
	bytes_t img;
	RunList runs;
	TRY(read_file(fn, img));
	int rc;
	if (img.size() >= 4 && memcmp(img.data(), "RLD\3", 4) == 0) rc = fmd_parse(img, runs);
	else if (img.size() >= 4 && memcmp(img.data(), "RB\2", 3) == 0) rc = fmr_parse(img, runs);
	else return rb3b_fail(RB3B_EFORMAT, "'%s' is neither FMD nor FMR", fn);
	if (rc != RB3B_OK) return rc;
	TRY(rb3b_index_from_runs(x, (int64_t)runs.sym.size(), runs.sym.data(), runs.len.data()));
	if (memcmp(img.data(), "RB\2", 3) == 0 && img[3] <= 2) x->so = img[3]; /* mr_restore, mrope.c:166-168; an FMD does not record the order */
	return RB3B_OK;
	GUARD_END
}
