/*
 * rb3_oracle.c -- TEST INFRASTRUCTURE ONLY.
 *
 * A plain-C, single-threaded restatement of the ropebwt3 BWT-merge path
 * (reference lh3/ropebwt3 @ v3.10-r281).  It exists so that the CUDA product
 * path in ropebwt3_b200/ can be checked bit-for-bit against an independent
 * CPU implementation of the same mathematics.  Nothing in the product may
 * include, link or call this file; only tests/, __graft_entry__.smoke() and
 * bench.py's cpu_baseline / --impl reference legs may load librb3oracle.so.
 *
 * Parity pinning: every function below is itself checked (tests/test_oracle*.py)
 * against (i) the literal known answers of SURVEY.md section 4.4, (ii) golden
 * fixtures under tests/golden/ generated with the unmodified reference built
 * into oracle/_ref/ (script: tests/golden/make_golden.py) and (iii), when
 * oracle/_ref/librb3ref.so is present, live differential runs.
 *
 * The data model is deliberately NOT the reference's (no B+-tree, no
 * iterators): a BWT is a flat run list (sym[i], len[i]); rank is a binary
 * search over prefix sums.  Each function cites the reference lines whose
 * observable behaviour it restates.
 */
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#include <stdio.h>

#define ASZ 6

/* ------------------------------------------------------------------ */
/* run lists                                                           */
/* ------------------------------------------------------------------ */

typedef struct {
	int64_t n, m;
	uint8_t *sym;
	int64_t *len;
} runs_t;

static void runs_push(runs_t *r, int c, int64_t l)
{ /* appends and coalesces, like rld_enc (rld0.c:153-161) */
	if (l <= 0) return;
	if (r->n > 0 && r->sym[r->n - 1] == c) { r->len[r->n - 1] += l; return; }
	if (r->n == r->m) {
		r->m = r->m ? r->m * 2 : 1024;
		r->sym = (uint8_t*)realloc(r->sym, r->m);
		r->len = (int64_t*)realloc(r->len, r->m * 8);
	}
	r->sym[r->n] = (uint8_t)c; r->len[r->n] = l; ++r->n;
}

/* plain BWT -> coalesced runs; returns #runs (rb3_enc_plain2rld, fm-index.c:13-29) */
int64_t ora_plain2runs(int64_t len, const uint8_t *bwt, uint8_t *osym, int64_t *olen)
{
	int64_t i, i0, n = 0;
	for (i0 = 0, i = 1; i <= len; ++i)
		if (i == len || bwt[i0] != bwt[i]) {
			if (osym) osym[n] = bwt[i0], olen[n] = i - i0;
			++n; i0 = i;
		}
	return len ? n : 0;
}

/* ------------------------------------------------------------------ */
/* rank over a run list                                                */
/* ------------------------------------------------------------------ */

typedef struct {
	int64_t n_runs, n_sym;
	const uint8_t *sym;
	const int64_t *len;
	int64_t *start;      /* start[i] = #symbols before run i; start[n_runs] = n_sym */
	int64_t (*cnt)[ASZ]; /* cnt[i][c] = #c before run i */
	int64_t tot[ASZ], acc[ASZ + 1];
} ridx_t;

static ridx_t *ridx_build(int64_t n, const uint8_t *sym, const int64_t *len)
{
	ridx_t *x = (ridx_t*)calloc(1, sizeof(ridx_t));
	int64_t i; int c;
	x->n_runs = n; x->sym = sym; x->len = len;
	x->start = (int64_t*)malloc((n + 1) * 8);
	x->cnt = (int64_t(*)[ASZ])malloc((n + 1) * sizeof(int64_t[ASZ]));
	x->start[0] = 0; memset(x->cnt[0], 0, sizeof(int64_t[ASZ]));
	for (i = 0; i < n; ++i) {
		memcpy(x->cnt[i + 1], x->cnt[i], sizeof(int64_t[ASZ]));
		x->cnt[i + 1][sym[i]] += len[i];
		x->start[i + 1] = x->start[i] + len[i];
	}
	x->n_sym = x->start[n];
	memcpy(x->tot, x->cnt[n], sizeof(x->tot));
	for (c = 0, x->acc[0] = 0; c < ASZ; ++c) x->acc[c + 1] = x->acc[c] + x->tot[c];
	return x;
}

static void ridx_free(ridx_t *x) { free(x->start); free(x->cnt); free(x); }

/* ok[c] = |{i<k : B[i]=c}|, returns B[k]; k>=n -> totals, -1.
 * Same contract as mr_rank1a (mrope.c:71-121, k>=tot at :89-93) and
 * rld_rank1a (rld0.c:416-437, k>=n at :421-424). */
static int ridx_rank1a(const ridx_t *x, int64_t k, int64_t ok[ASZ])
{
	int64_t lo = 0, hi = x->n_runs;
	if (k >= x->n_sym) { memcpy(ok, x->tot, sizeof(x->tot)); return -1; }
	while (hi - lo > 1) { /* last run with start <= k */
		int64_t mid = (lo + hi) >> 1;
		if (x->start[mid] <= k) lo = mid; else hi = mid;
	}
	memcpy(ok, x->cnt[lo], sizeof(int64_t[ASZ]));
	ok[x->sym[lo]] += k - x->start[lo];
	return x->sym[lo];
}

int ora_rank1a(int64_t n_runs, const uint8_t *sym, const int64_t *len, int64_t nq, const int64_t *k, int64_t *ok, int8_t *ret)
{
	ridx_t *x = ridx_build(n_runs, sym, len);
	int64_t i;
	for (i = 0; i < nq; ++i) ret[i] = (int8_t)ridx_rank1a(x, k[i], ok + i * ASZ);
	ridx_free(x);
	return 0;
}

/* ------------------------------------------------------------------ */
/* partial BWT of a batch (sais-ss.c:10-56 semantics, naive algorithm)  */
/* ------------------------------------------------------------------ */

static const uint8_t *g_T;
static int64_t g_len;

static int suf_cmp(const void *pa, const void *pb)
{ /* generalized suffix order: a suffix ends at its first 0; equal up to and
     including the 0 => the earlier sentinel is smaller (libsais GSA). */
	int64_t a = *(const int64_t*)pa, b = *(const int64_t*)pb, i = a, j = b;
	for (;;) {
		uint8_t x = g_T[i], y = g_T[j];
		if (x != y) return x < y ? -1 : 1;
		if (x == 0) return a < b ? -1 : a > b ? 1 : 0;
		++i, ++j;
	}
}

/* text: concatenated 0-terminated nt6 strings; replaced by its BWT in place:
 * BWT[i] = T[SA[i]-1], and T[len-1] for SA[i]==0 (sais-ss.c:23-26). */
int ora_build_bwt(int64_t len, uint8_t *T, int64_t *sa_out)
{
	int64_t i, *SA;
	uint8_t *B;
	if (len <= 0 || T[len - 1] != 0) return -1;
	SA = (int64_t*)malloc(len * 8); B = (uint8_t*)malloc(len);
	for (i = 0; i < len; ++i) SA[i] = i;
	g_T = T; g_len = len;
	qsort(SA, len, 8, suf_cmp);
	for (i = 0; i < len; ++i) B[i] = T[SA[i] == 0 ? len - 1 : SA[i] - 1];
	memcpy(T, B, len);
	if (sa_out) memcpy(sa_out, SA, len * 8);
	free(SA); free(B);
	return 0;
}

/* ------------------------------------------------------------------ */
/* interleave array: rb3_mg_rank_plain + rb3_mg_rank1_plain             */
/* (fm-index.c:160-175, 202-225)                                        */
/* ------------------------------------------------------------------ */

/* rb[i] on return = (ka+i)<<6 | B[i]<<3 | first symbol of suffix i, where ka =
 * #suffixes of A smaller than suffix i of the batch; acc[7] = C[] of the batch. */
int ora_mg_rank_plain(int64_t n_runs, const uint8_t *asym, const int64_t *alen,
                      int64_t len, const uint8_t *seq, int64_t *rb, int64_t acc[ASZ + 1])
{
	ridx_t *A = ridx_build(n_runs, asym, alen);
	int64_t i, p, c[ASZ];
	int a;
	memset(c, 0, sizeof(c));
	for (i = 0; i < len; ++i) { if (seq[i] >= ASZ) { ridx_free(A); return -1; } ++c[seq[i]]; }
	for (acc[0] = 0, a = 0; a < ASZ; ++a) acc[a + 1] = acc[a] + c[a];
	memset(c, 0, sizeof(c));
	for (i = 0; i < len; ++i) { /* LF of the batch and its symbol, fm-index.c:212-216 */
		a = seq[i];
		rb[i] = (acc[a] + c[a]) << 3 | a;
		++c[a];
	}
	for (p = 0; p < acc[1]; ++p) { /* one dependent chain per new sequence, fm-index.c:160-175 */
		int64_t ka = A->acc[1], kb = p, ok[ASZ];
		int last = 0;
		for (;;) {
			int64_t nxt = rb[kb] >> 3;
			int s = rb[kb] & 7;
			rb[kb] = (ka + kb) << 6 | s << 3 | last;
			last = s;
			if (s == 0) break;
			kb = nxt;
			ridx_rank1a(A, ka, ok);
			ka = A->acc[s] + ok[s];
		}
	}
	ridx_free(A);
	return 0;
}

/* ------------------------------------------------------------------ */
/* merge: worker_mgins + rb3_fmi_merge_plain (fm-index.c:237-249,279-303)*/
/* ------------------------------------------------------------------ */

/* Inserting row i of the batch at merged position rb[i]>>6, in increasing i
 * inside every bucket, is the interleave merged[ka_i+i] = B[i] with A's
 * symbols filling the remaining slots in order.  Output: coalesced runs.
 * Returns #runs (call with osym==NULL to size). */
int64_t ora_merge_runs(int64_t n_runs, const uint8_t *asym, const int64_t *alen,
                       int64_t len, const int64_t *rb, uint8_t *osym, int64_t *olen)
{
	runs_t out = {0, 0, 0, 0};
	int64_t i, ia = 0, used = 0, apos = 0; /* used: consumed part of run ia; apos: A symbols emitted */
	for (i = 0; i < len; ++i) {
		int64_t ka = (rb[i] >> 6) - i;
		while (apos < ka) {
			int64_t t = alen[ia] - used;
			if (t > ka - apos) t = ka - apos;
			runs_push(&out, asym[ia], t);
			used += t; apos += t;
			if (used == alen[ia]) ++ia, used = 0;
		}
		runs_push(&out, rb[i] >> 3 & 7, 1);
	}
	for (; ia < n_runs; ++ia, used = 0) runs_push(&out, asym[ia], alen[ia] - used);
	if (osym) { memcpy(osym, out.sym, out.n); memcpy(olen, out.len, out.n * 8); }
	free(out.sym); free(out.len);
	return out.n;
}

/* ------------------------------------------------------------------ */
/* FMD writer (rld0.c:45-51 delta code, :107-135 block headers,         */
/* :137-161 bit packing, :163-204 rank frames, :206-243 finish+dump)    */
/* ------------------------------------------------------------------ */

#define LBITS 23
#define LSIZE (1LL << LBITS)
#define SSIZE 8

static int ilog2_64(uint64_t v) { int l = -1; while (v) ++l, v >>= 1; return l; } /* ilog2(0) = -1 as in rld0.c:23-40 */

typedef struct {
	uint64_t *w; int64_t m;   /* flat word array (the reference's 2^23-word chunks laid end to end) */
	int64_t h, p; int r;      /* block head, current word, free bits in it */
	int64_t stail;
	uint64_t cnt[ASZ + 1], mcnt[ASZ + 1];
	int pc; int64_t pl;       /* pending run */
} fmdw_t;

static void fmdw_need(fmdw_t *e, int64_t idx)
{
	if (idx >= e->m) {
		int64_t m = e->m ? e->m : 1 << 16;
		while (idx >= m) m <<= 1;
		e->w = (uint64_t*)realloc(e->w, m * 8);
		memset(e->w + e->m, 0, (m - e->m) * 8);
		e->m = m;
	}
}

static int64_t fmdw_stail(int64_t h) { return h + SSIZE - (((h + SSIZE) & (LSIZE - 1)) == 0 ? 2 : 1); } /* rld0.h:81 */

static void fmdw_next_block(fmdw_t *e)
{
	uint64_t d[ASZ + 1];
	int i, type;
	static const int off0[3] = { 2, 4, 7 }; /* rld0.c:71-73 */
	e->h += SSIZE;
	fmdw_need(e, e->h + 2 * SSIZE);
	for (i = 0; i <= ASZ; ++i) d[i] = e->cnt[i] - e->mcnt[i];
	if (d[0] < 0x4000) { uint16_t *q = (uint16_t*)(e->w + e->h); for (i = 0; i <= ASZ; ++i) q[i] = (uint16_t)d[i]; type = 0; }
	else if (d[0] < 0x40000000) { uint32_t *q = (uint32_t*)(e->w + e->h); for (i = 0; i <= ASZ; ++i) q[i] = (uint32_t)d[i]; type = 1; }
	else { uint64_t *q = e->w + e->h; for (i = 0; i <= ASZ; ++i) q[i] = d[i]; type = 2; }
	e->w[e->h] |= (uint64_t)type << 62;
	e->p = e->h + off0[type]; e->r = 64; e->stail = fmdw_stail(e->h);
	memcpy(e->mcnt, e->cnt, sizeof(e->cnt));
}

static void fmdw_enc1(fmdw_t *e, int64_t l, int c)
{
	int y = ilog2_64((uint64_t)l), z = ilog2_64((uint64_t)y + 1);
	int w = 2 * z + 1 + y + 3;
	uint64_t x = ((((uint64_t)l ^ (1ULL << y)) | (uint64_t)(y + 1) << y) << 3) | (uint64_t)c;
	if (w >= e->r && e->p == e->stail) fmdw_next_block(e);
	if (w > e->r) {
		w -= e->r;
		e->w[e->p++] |= x >> w;
		e->r = 64 - w;
		e->w[e->p] = x << e->r;
	} else { e->r -= w; e->w[e->p] |= x << e->r; }
	e->cnt[0] += l; e->cnt[c + 1] += l;
}

/* Encode runs (need not be coalesced) to a malloc'd .fmd image. */
int64_t ora_fmd_encode(int64_t n_runs, const uint8_t *sym, const int64_t *len, uint8_t **out)
{
	fmdw_t e;
	int64_t i, n_words, n_blks, n_frames, last, k, total;
	uint64_t *frame, cnt[ASZ], n_bytes;
	int ibits, j;
	uint8_t *buf, *q;
	memset(&e, 0, sizeof(e));
	fmdw_need(&e, 4 * SSIZE);
	e.h = 0; e.p = 2; e.r = 64; e.stail = fmdw_stail(0); e.pc = -1; e.pl = 0; /* block 0: all-zero type-0 header */
	for (i = 0; i < n_runs; ++i) {
		if (len[i] == 0) continue;
		if (e.pc != sym[i]) { if (e.pl) fmdw_enc1(&e, e.pl, e.pc); e.pl = len[i]; e.pc = sym[i]; }
		else e.pl += len[i];
	}
	if (e.pl) fmdw_enc1(&e, e.pl, e.pc);
	fmdw_next_block(&e);
	n_words = e.p; n_bytes = (uint64_t)n_words * 8;
	total = (int64_t)e.mcnt[0];
	/* rank frames */
	n_blks = n_words / SSIZE + 1;
	last = n_words / SSIZE * SSIZE;
	ibits = ilog2_64((uint64_t)(total / n_blks)) + 4;
	n_frames = ((total + (1LL << ibits) - 1) >> ibits) + 1;
	frame = (uint64_t*)calloc(n_frames * (ASZ + 1), 8);
	memset(cnt, 0, sizeof(cnt));
	for (i = SSIZE, k = 1; i <= last; i += SSIZE) {
		uint64_t sum, *p = e.w + i;
		int type = (int)(*p >> 62);
		if (type == 0) { uint16_t *t = (uint16_t*)p; for (j = 1; j <= ASZ; ++j) cnt[j - 1] += t[j]; }
		else if (type == 1) { uint32_t *t = (uint32_t*)p; for (j = 1; j <= ASZ; ++j) cnt[j - 1] += t[j] & 0x3fffffff; }
		else { for (j = 1; j <= ASZ; ++j) cnt[j - 1] += p[j]; }
		for (j = 0, sum = 0; j < ASZ; ++j) sum += cnt[j];
		while (sum >= (uint64_t)k << ibits) ++k;
		if (k < n_frames) {
			frame[k * (ASZ + 1)] = (uint64_t)i;
			for (j = 0; j < ASZ; ++j) frame[k * (ASZ + 1) + 1 + j] = cnt[j];
		}
	}
	for (k = 1; k < n_frames; ++k)
		if (frame[k * (ASZ + 1)] == 0)
			memcpy(frame + k * (ASZ + 1), frame + (k - 1) * (ASZ + 1), (ASZ + 1) * 8);
	/* file image: rld_dump, rld0.c:222-243 */
	buf = q = (uint8_t*)malloc(80 + n_bytes + n_frames * (ASZ + 1) * 8);
	memcpy(q, "RLD\3", 4); q += 4;
	{ uint32_t a = ASZ << 16 | 3; memcpy(q, &a, 4); q += 4; }
	{ uint64_t z = 0; memcpy(q, &z, 8); q += 8; }
	memcpy(q, &n_bytes, 8); q += 8;
	{ uint64_t f = (uint64_t)n_frames; memcpy(q, &f, 8); q += 8; }
	memcpy(q, e.mcnt + 1, ASZ * 8); q += ASZ * 8;
	memcpy(q, e.w, n_bytes); q += n_bytes;
	memcpy(q, frame, n_frames * (ASZ + 1) * 8); q += n_frames * (ASZ + 1) * 8;
	free(frame); free(e.w);
	*out = buf;
	return q - buf;
}

void ora_free(void *p) { free(p); }

/* FMD reader: image -> runs as stored (one per code).  Returns #runs, or -1.
 * Restates rld_dec0/rld_dec (rld0.h:85-125) on a flat image. */
static uint64_t fmd_peek(const uint64_t *w, int64_t p, int64_t stail, int64_t bit)
{ /* 64 bits starting `bit` bits after word p, zero padded past word stail */
	int64_t wi = p + (bit >> 6); int sh = bit & 63; uint64_t x;
	if (wi > stail) return 0;
	x = w[wi] << sh;
	if (sh && wi < stail) x |= w[wi + 1] >> (64 - sh);
	return x;
}

int64_t ora_fmd_decode(int64_t n_file, const uint8_t *img, uint8_t *osym, int64_t *olen, int64_t mc[ASZ])
{
	uint64_t n_bytes;
	const uint64_t *w;
	int64_t n_words, last, h, n = 0;
	static const int off0[3] = { 2, 4, 7 };
	if (n_file < 80 || memcmp(img, "RLD\3", 4) != 0) return -1;
	memcpy(&n_bytes, img + 16, 8);
	if (mc) memcpy(mc, img + 32, ASZ * 8);
	w = (const uint64_t*)(img + 80);
	n_words = n_bytes / 8; last = n_words / SSIZE * SSIZE;
	for (h = 0; h < last; h += SSIZE) {
		int64_t p = h + off0[w[h] >> 62], stail = fmdw_stail(h), bit = 0;
		for (;;) {
			uint64_t x = fmd_peek(w, p, stail, bit), l;
			int z = 0, y, c;
			if (x >> 58 == 0) break;              /* >= 6 leading zeros: end of block (rld0.h:92) */
			while (!(x << z >> 63)) ++z;          /* gamma prefix: z zeros, then y+1 in z+1 bits */
			y = (int)(x << z >> (63 - z)) - 1;
			l = (y ? fmd_peek(w, p, stail, bit + 2 * z + 1) >> (64 - y) : 0) | 1ULL << y;
			c = (int)(fmd_peek(w, p, stail, bit + 2 * z + 1 + y) >> 61);
			if (c > ASZ) break;
			if (osym) osym[n] = (uint8_t)c, olen[n] = (int64_t)l;
			++n; bit += 2 * z + 1 + y + 3;
		}
	}
	return n;
}

/* ------------------------------------------------------------------ */
/* FMR codec (rle.h:39-75 run code; mrope.c:152-177, rope.c:265-330)     */
/* ------------------------------------------------------------------ */

static int fmr_enc1(uint8_t *p, int c, int64_t l)
{ /* 1/2/4/8-byte UTF-8-like code, rle.h:53-75 */
	if (l < 16) { p[0] = (uint8_t)(l << 3 | c); return 1; }
	if (l < 256) { p[0] = (uint8_t)(0xC0 | (l >> 6) << 3 | c); p[1] = (uint8_t)(0x80 | (l & 0x3f)); return 2; }
	if (l < (1LL << 19)) {
		int i; p[0] = (uint8_t)(0xE0 | (l >> 18) << 3 | c);
		for (i = 1; i < 4; ++i) p[i] = (uint8_t)(0x80 | (l >> (6 * (3 - i)) & 0x3f));
		return 4;
	} else {
		int i; p[0] = (uint8_t)(0xF0 | (l >> 42) << 3 | c);
		for (i = 1; i < 8; ++i) p[i] = (uint8_t)(0x80 | (l >> (6 * (7 - i)) & 0x3f));
		return 8;
	}
}

static const uint8_t *fmr_dec1(const uint8_t *p, int *c, int64_t *l)
{ /* rle.h:39-51 */
	*c = p[0] & 7;
	if ((p[0] & 0x80) == 0) { *l = p[0] >> 3; return p + 1; }
	if (p[0] >> 5 == 6) { *l = (int64_t)(p[0] & 0x18) << 3 | (p[1] & 0x3f); return p + 2; }
	{
		int n = (p[0] & 0x10) ? 8 : 4, i;
		int64_t v = p[0] >> 3 & 1;
		for (i = 1; i < n; ++i) v = v << 6 | (p[i] & 0x3f);
		*l = v; return p + n;
	}
}

typedef struct { const uint8_t *p, *end; int err; runs_t *out; int64_t c[ASZ]; int rope; uint8_t *rope_of_run; } fmr_rd_t;

static void fmr_read_node(fmr_rd_t *r)
{ /* pre-order node record, rope.c:289-317 */
	uint8_t is_bottom; int16_t n; int i;
	if (r->err || r->p + 3 > r->end) { r->err = 1; return; }
	is_bottom = r->p[0]; memcpy(&n, r->p + 1, 2); r->p += 3;
	for (i = 0; i < n; ++i) {
		if (is_bottom) {
			uint16_t nb; const uint8_t *q, *e; int64_t leafc[ASZ], got[ASZ]; int a;
			if (r->p + 50 > r->end) { r->err = 1; return; }
			memcpy(leafc, r->p, 48); memcpy(&nb, r->p + 48, 2); r->p += 50;
			if (r->p + nb > r->end) { r->err = 1; return; }
			memset(got, 0, sizeof(got));
			for (q = r->p, e = r->p + nb; q < e;) {
				int c; int64_t l;
				q = fmr_dec1(q, &c, &l);
				if (c >= ASZ) { r->err = 2; return; }
				/* runs are NOT coalesced across leaves here; keep them as stored */
				if (r->out->n == r->out->m) {
					r->out->m = r->out->m ? r->out->m * 2 : 1024;
					r->out->sym = (uint8_t*)realloc(r->out->sym, r->out->m);
					r->out->len = (int64_t*)realloc(r->out->len, r->out->m * 8);
				}
				r->out->sym[r->out->n] = (uint8_t)c; r->out->len[r->out->n++] = l;
				got[c] += l; r->c[c] += l;
			}
			for (a = 0; a < ASZ; ++a) if (got[a] != leafc[a]) r->err = 3; /* stored margins must match */
			r->p += nb;
		} else fmr_read_node(r);
		if (r->err) return;
	}
}

/* FMR image -> runs as stored.  rope_cnt[a*6+b] = #b in rope a.  Returns #runs or <0. */
int64_t ora_fmr_decode(int64_t n_file, const uint8_t *img, uint8_t *osym, int64_t *olen, int64_t rope_cnt[36], int32_t geom[3])
{
	fmr_rd_t r; runs_t out = {0, 0, 0, 0};
	int a; int64_t n;
	if (n_file < 4 || memcmp(img, "RB\2", 3) != 0) return -1;
	memset(&r, 0, sizeof(r));
	r.p = img + 4; r.end = img + n_file; r.out = &out;
	if (geom) geom[0] = img[3];
	for (a = 0; a < ASZ && !r.err; ++a) {
		int32_t mn, bl;
		if (r.p + 8 > r.end) { r.err = 1; break; }
		memcpy(&mn, r.p, 4); memcpy(&bl, r.p + 4, 4); r.p += 8;
		if (geom) geom[1] = mn, geom[2] = bl;
		memset(r.c, 0, sizeof(r.c));
		fmr_read_node(&r);
		if (rope_cnt) memcpy(rope_cnt + a * ASZ, r.c, sizeof(r.c));
	}
	n = r.err ? -10 - r.err : out.n;
	if (n >= 0 && osym) { memcpy(osym, out.sym, out.n); memcpy(olen, out.len, out.n * 8); }
	free(out.sym); free(out.len);
	return n;
}

/* Runs -> a legal FMR image (balanced tree, leaves <= block_len-18 payload bytes,
 * fan-out <= max_nodes; SURVEY A.2 invariants; rope a = rows [acc[a],acc[a+1])). */
typedef struct { uint8_t *b; int64_t n, m; } obuf_t;
static void ob_put(obuf_t *o, const void *p, int64_t n)
{
	if (o->n + n > o->m) { o->m = (o->n + n) * 2 + 4096; o->b = (uint8_t*)realloc(o->b, o->m); }
	memcpy(o->b + o->n, p, n); o->n += n;
}

typedef struct { int64_t c[ASZ]; uint16_t nb; uint8_t *bytes; } leaf_t;

static void fmr_write_level(obuf_t *o, const leaf_t *lv, int64_t n_leaf, int max_nodes, int64_t lo, int64_t hi, int64_t span)
{ /* node covering leaves [lo,hi); span = leaves per child subtree at this level */
	uint8_t is_bottom = span == 1;
	int64_t i, nchild = (hi - lo + span - 1) / span;
	int16_t n = (int16_t)nchild;
	ob_put(o, &is_bottom, 1); ob_put(o, &n, 2);
	for (i = 0; i < nchild; ++i) {
		int64_t a = lo + i * span, b = a + span < hi ? a + span : hi;
		if (is_bottom) { ob_put(o, lv[a].c, 48); ob_put(o, &lv[a].nb, 2); ob_put(o, lv[a].bytes, lv[a].nb); }
		else fmr_write_level(o, lv, n_leaf, max_nodes, a, b, span / max_nodes);
	}
}

int64_t ora_fmr_encode(int64_t n_runs, const uint8_t *sym, const int64_t *len, int max_nodes, int block_len, int so, uint8_t **outp)
{
	obuf_t o = {0, 0, 0};
	int64_t tot[ASZ], acc[ASZ + 1], i, pos = 0, ri = 0, used = 0;
	int a, cap = block_len - 18 - 8; /* payload cap leaves room for one 8-byte code */
	uint8_t hdr[4] = { 'R', 'B', 2, (uint8_t)so };
	memset(tot, 0, sizeof(tot));
	for (i = 0; i < n_runs; ++i) tot[sym[i]] += len[i];
	for (a = 0, acc[0] = 0; a < ASZ; ++a) acc[a + 1] = acc[a] + tot[a];
	ob_put(&o, hdr, 4);
	for (a = 0; a < ASZ; ++a) {
		leaf_t *lv = 0; int64_t nl = 0, ml = 0, span; int32_t mn = max_nodes, bl = block_len;
		int last_c = -1; int64_t last_l = 0; /* pending run inside this rope */
		/* cut the flat run list at the rope boundary and fill leaves greedily */
#define NEW_LEAF() do { if (nl == ml) { ml = ml ? ml * 2 : 64; lv = (leaf_t*)realloc(lv, ml * sizeof(leaf_t)); } \
		memset(&lv[nl], 0, sizeof(leaf_t)); lv[nl].bytes = (uint8_t*)malloc(block_len); ++nl; } while (0)
#define FLUSH() do { if (last_l) { if (nl == 0 || lv[nl-1].nb > cap) NEW_LEAF(); \
		lv[nl-1].nb += fmr_enc1(lv[nl-1].bytes + lv[nl-1].nb, last_c, last_l); lv[nl-1].c[last_c] += last_l; last_l = 0; } } while (0)
		while (pos < acc[a + 1]) {
			int64_t t = len[ri] - used;
			if (t > acc[a + 1] - pos) t = acc[a + 1] - pos;
			if (sym[ri] == last_c) last_l += t; else { FLUSH(); last_c = sym[ri]; last_l = t; }
			used += t; pos += t;
			if (used == len[ri]) ++ri, used = 0;
		}
		FLUSH();
		if (nl == 0) NEW_LEAF(); /* empty rope = one bottom node with one empty leaf, rope.c:64-67 */
		ob_put(&o, &mn, 4); ob_put(&o, &bl, 4);
		for (span = 1; span * max_nodes < nl; span *= max_nodes) {}
		fmr_write_level(&o, lv, nl, max_nodes, 0, nl, span);
		for (i = 0; i < nl; ++i) free(lv[i].bytes);
		free(lv);
	}
	*outp = o.b;
	return o.n;
}
