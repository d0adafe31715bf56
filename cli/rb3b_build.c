/*
 * rb3b_build.c -- `ropebwt3 build` on the B200 engine (C host side).
 *
 * Keeps the reference CLI for its default construction path (libsais + merge):
 * option string "l:n:m:t:2sri:LFRo:dbTS:p:e" with argument permutation
 * (build.c:146), defaults of rb3_bopt_init (build.c:31-41), -i incremental
 * append (build.c:172-184), per-file batching by -m (io.c:104-125: a batch
 * closes after the record that makes it longer than -m), forward strand then
 * reverse complement per record (io.c:84-102), -S checkpoint after every input
 * file (build.c:232-238), output formats -d/-b/plain (build.c:245-251), exit
 * codes (build.c:175-178,243).  Everything on the hot path -- partial BWT of
 * a batch, interleave array, merge -- runs on the device through librb3b200.so;
 * this file only parses options and sequence files.
 *
 * -2/-s/-r go through rb3b_insert_multi (mr_insert_multi, build.c:214-218).
 *
 * Like the reference's pipeline mode (build.c:55-83,186-201: kt_pipeline of read+SAIS and merge) everything overlaps,
 * here always and for any number of input files: a reader thread parses and encodes batch i+2, a second thread copies
 * batch i+1 to the device and builds its partial BWT there (rb3_build_sais's place in the reference's step 0), and the
 * main thread merges batch i (step 1).  The two device-side threads run in their own library contexts, i.e. on their own
 * CUDA streams, so the suffix sort of one batch and the merge of the previous one share the GPU.
 *
 * `merge` (main.c:84-133): rb3_fmi_merge of further indexes into the first one, both sides on the device.
 *
 * Not covered (documented in DESIGN.md): -T (tree dump: there is no tree), -e (BRE).  -t and -p are accepted and
 * ignored (parallelism comes from the device); -l/-n only shape the .fmr dump.
 */
#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <sys/stat.h>
#include <stdint.h>
#include <unistd.h>
#include <time.h>
#include <sys/resource.h>
#include <zlib.h>
#include <pthread.h>
#include "../include/rb3_b200.h"

static const unsigned char nt6_table[128] = { /* io.c:12-21: $ACGTN = 0..5 */
	0, 1, 2, 3,  4, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,
	5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,
	5, 1, 5, 2,  5, 5, 5, 3,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  4, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,
	5, 1, 5, 2,  5, 5, 5, 3,  5, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5,  4, 5, 5, 5,  5, 5, 5, 5,  5, 5, 5, 5
};

typedef struct { char *s; size_t l, m; } str_t;

static void str_reserve(str_t *b, size_t need)
{
	if (need > b->m) {
		b->m = need + (need >> 1) + 64;
		b->s = (char*)realloc(b->s, b->m);
		if (b->s == 0) { fprintf(stderr, "ERROR: out of host memory\n"); exit(1); }
	}
}

static double t_real0;
static double realtime(void) { struct timespec ts; clock_gettime(CLOCK_REALTIME, &ts); return ts.tv_sec + ts.tv_nsec * 1e-9; }
static double cputime(void) { struct rusage r; getrusage(RUSAGE_SELF, &r); return r.ru_utime.tv_sec + r.ru_stime.tv_sec + 1e-6 * (r.ru_utime.tv_usec + r.ru_stime.tv_usec); }
#define LOG(...) do { fprintf(stderr, "[M::main_build::%.3f*%.2f] ", realtime() - t_real0, cputime() / (realtime() - t_real0 + 1e-9)); fprintf(stderr, __VA_ARGS__); fputc('\n', stderr); } while (0)

static int64_t parse_num(const char *str)
{ /* misc.c:7-16 */
	char *p;
	double x = strtod(str, &p);
	if (*p == 'G' || *p == 'g') x *= 1e9;
	else if (*p == 'M' || *p == 'm') x *= 1e6;
	else if (*p == 'K' || *p == 'k') x *= 1e3;
	return (int64_t)(x + .499);
}

/* ---- sequence reader: FASTA/FASTQ (multi-line) or one sequence per line; gzip transparent ---- */

typedef struct {
	gzFile fp;
	int is_line, last; /* last: a character read ahead (-2 none) */
	unsigned char buf[1 << 16];
	int n, p;
} reader_t;

static int rd_getc(reader_t *r)
{
	if (r->p >= r->n) {
		r->n = gzread(r->fp, r->buf, sizeof(r->buf));
		r->p = 0;
		if (r->n <= 0) return -1;
	}
	return r->buf[r->p++];
}

static int rd_line(reader_t *r, str_t *out, int append)
{ /* one line without its terminator; -1 at EOF with nothing read.  Works on whole buffer spans (memchr + memcpy). */
	int got = 0;
	if (!append) out->l = 0;
	for (;;) {
		if (r->p >= r->n) {
			r->n = gzread(r->fp, r->buf, sizeof(r->buf));
			r->p = 0;
			if (r->n <= 0) break;
		}
		got = 1;
		unsigned char *s = r->buf + r->p, *e = (unsigned char*)memchr(s, '\n', r->n - r->p);
		size_t l = e ? (size_t)(e - s) : (size_t)(r->n - r->p);
		str_reserve(out, out->l + l + 2);
		memcpy(out->s + out->l, s, l);
		out->l += l;
		r->p += (int)l + (e ? 1 : 0);
		if (e) break;
	}
	if (!got) return -1;
	if (out->l && out->s[out->l - 1] == '\r') --out->l;
	str_reserve(out, out->l + 1);
	out->s[out->l] = 0;
	return 0;
}

/* next record's residues into rec; returns -1 at EOF */
static int rd_record(reader_t *r, str_t *rec, str_t *tmp)
{
	int c;
	if (r->is_line) return rd_line(r, rec, 0);
	if (r->last == -2) { /* skip to the first header */
		while ((c = rd_getc(r)) >= 0 && c != '>' && c != '@') {}
		if (c < 0) return -1;
		r->last = c;
	}
	if (r->last < 0) return -1;
	if (rd_line(r, tmp, 0) < 0) { r->last = -1; return -1; } /* rest of the header line */
	rec->l = 0;
	for (;;) {
		c = rd_getc(r);
		if (c < 0) { r->last = -1; break; }
		if (c == '>' || c == '@') { r->last = c; break; }
		if (c == '+') { /* FASTQ: skip the '+' line and as many quality characters as residues */
			size_t q = 0;
			rd_line(r, tmp, 0);
			while (q < rec->l && rd_line(r, tmp, 0) == 0) q += tmp->l;
			while ((c = rd_getc(r)) >= 0 && c != '>' && c != '@') {}
			r->last = c < 0 ? -1 : c;
			break;
		}
		if (c == '\n' || c == '\r') continue;
		str_reserve(rec, rec->l + 2);
		rec->s[rec->l++] = (char)c;
		rd_line(r, rec, 1); /* the rest of this sequence line */
	}
	str_reserve(rec, rec->l + 1);
	rec->s[rec->l] = 0;
	return 0;
}

#if defined(__SSE2__) && !defined(RB3B_NO_SIMD)
#include <emmintrin.h>
/* nt6_table on 16 characters at once: a/c/g/t in either case -> 1..4, the raw codes 0..4 -> themselves, all else -> 5 */
static inline __m128i nt6_map16(__m128i c)
{
	const __m128i l = _mm_or_si128(c, _mm_set1_epi8(0x20));
	const __m128i isa = _mm_cmpeq_epi8(l, _mm_set1_epi8('a')), isc = _mm_cmpeq_epi8(l, _mm_set1_epi8('c'));
	const __m128i isg = _mm_cmpeq_epi8(l, _mm_set1_epi8('g')), ist = _mm_cmpeq_epi8(l, _mm_set1_epi8('t'));
	const __m128i raw = _mm_cmpeq_epi8(_mm_min_epu8(c, _mm_set1_epi8(4)), c); /* c <= 4 (unsigned) */
	__m128i v = _mm_set1_epi8(5);
	v = _mm_sub_epi8(v, _mm_and_si128(isa, _mm_set1_epi8(4)));
	v = _mm_sub_epi8(v, _mm_and_si128(isc, _mm_set1_epi8(3)));
	v = _mm_sub_epi8(v, _mm_and_si128(isg, _mm_set1_epi8(2)));
	v = _mm_sub_epi8(v, _mm_and_si128(ist, _mm_set1_epi8(1)));
	return _mm_or_si128(_mm_and_si128(raw, c), _mm_andnot_si128(raw, v));
}
/* codes 1..4 -> 5 - code (reverse complement), 0 and 5 stay */
static inline __m128i nt6_comp16(__m128i v)
{
	const __m128i keep = _mm_or_si128(_mm_cmpeq_epi8(v, _mm_setzero_si128()), _mm_cmpeq_epi8(v, _mm_set1_epi8(5)));
	return _mm_or_si128(_mm_and_si128(keep, v), _mm_andnot_si128(keep, _mm_sub_epi8(_mm_set1_epi8(5), v)));
}
static inline __m128i rev16(__m128i x)
{ /* the 16 bytes in reverse order */
	x = _mm_or_si128(_mm_slli_epi16(x, 8), _mm_srli_epi16(x, 8));
	x = _mm_shufflelo_epi16(x, 0x1B);
	x = _mm_shufflehi_epi16(x, 0x1B);
	return _mm_shuffle_epi32(x, 0x4E);
}
#endif

static void seq_add(str_t *seq, const str_t *rec, int is_for, int is_rev, int64_t *n_seq)
{ /* rb3_seq_add, io.c:84-102 */
	size_t i, l = rec->l;
	if (is_for) {
		str_reserve(seq, seq->l + l + 1);
		i = 0;
#if defined(__SSE2__) && !defined(RB3B_NO_SIMD)
		for (; i + 16 <= l; i += 16)
			_mm_storeu_si128((__m128i*)(seq->s + seq->l + i), nt6_map16(_mm_loadu_si128((const __m128i*)(rec->s + i))));
#endif
		for (; i < l; ++i) { unsigned char c = (unsigned char)rec->s[i]; seq->s[seq->l + i] = c < 128 ? nt6_table[c] : 5; }
		seq->s[seq->l + l] = 0;
		seq->l += l + 1; ++*n_seq;
	}
	if (is_rev) {
		str_reserve(seq, seq->l + l + 1);
		i = 0;
#if defined(__SSE2__) && !defined(RB3B_NO_SIMD)
		for (; i + 16 <= l; i += 16) /* output positions i..i+15 come from input l-16-i .. l-1-i, reversed */
			_mm_storeu_si128((__m128i*)(seq->s + seq->l + i), rev16(nt6_comp16(nt6_map16(_mm_loadu_si128((const __m128i*)(rec->s + l - 16 - i))))));
#endif
		for (; i < l; ++i) {
			unsigned char c = (unsigned char)rec->s[l - 1 - i];
			int x = c < 128 ? nt6_table[c] : 5;
			seq->s[seq->l + i] = (x >= 1 && x <= 4) ? 5 - x : x;
		}
		seq->s[seq->l + l] = 0;
		seq->l += l + 1; ++*n_seq;
	}
}

/* ---- reader thread: produces batches in input order ---- */

typedef struct {
	str_t seq;          /* encoded batch */
	int64_t n_seq;
	int file;           /* index into argv of the file the batch came from */
	int last_of_file;   /* the file ends with this batch (-S checkpoint point); n_seq == 0: nothing left in the file */
	int open_failed;
} batch_t;

typedef struct {
	int argc, first; char **argv;
	int is_line, no_for, no_rev;
	int64_t batch;      /* -m: soft limit, a batch closes after the record that crosses it (io.c:114,119) */
	int64_t hard;       /* what the device can sort and merge at once: a batch closes BEFORE the record that would cross it */
	batch_t slot[2];
	int n_full, head, tail, done; /* ring of two slots */
	pthread_mutex_t mu; pthread_cond_t cv;
} pipe_t;

static batch_t *pipe_claim(pipe_t *p)
{ /* reader: wait for a free slot */
	pthread_mutex_lock(&p->mu);
	while (p->n_full == 2) pthread_cond_wait(&p->cv, &p->mu);
	pthread_mutex_unlock(&p->mu);
	return &p->slot[p->tail];
}

static void pipe_publish(pipe_t *p)
{
	pthread_mutex_lock(&p->mu);
	p->tail ^= 1; ++p->n_full;
	pthread_cond_broadcast(&p->cv);
	pthread_mutex_unlock(&p->mu);
}

static batch_t *pipe_next(pipe_t *p)
{ /* consumer: next batch in order, NULL when the reader is finished */
	batch_t *b = 0;
	pthread_mutex_lock(&p->mu);
	while (p->n_full == 0 && !p->done) pthread_cond_wait(&p->cv, &p->mu);
	if (p->n_full) b = &p->slot[p->head];
	pthread_mutex_unlock(&p->mu);
	return b;
}

static void pipe_release(pipe_t *p)
{
	pthread_mutex_lock(&p->mu);
	p->head ^= 1; --p->n_full;
	pthread_cond_broadcast(&p->cv);
	pthread_mutex_unlock(&p->mu);
}

static void *reader_main(void *arg)
{
	pipe_t *p = (pipe_t*)arg;
	str_t rec = {0, 0, 0}, tmp = {0, 0, 0};
	int i;
	for (i = p->first; i < p->argc; ++i) {
		reader_t *r = (reader_t*)calloc(1, sizeof(reader_t));
		int eof = 0;
		batch_t *b;
		r->fp = strcmp(p->argv[i], "-") ? gzopen(p->argv[i], "r") : gzdopen(0, "r");
		if (r->fp == 0) { /* build.c:207-210: report and go on */
			b = pipe_claim(p);
			b->seq.l = 0; b->n_seq = 0; b->file = i; b->last_of_file = 0; b->open_failed = 1;
			pipe_publish(p);
			free(r);
			continue;
		}
		r->is_line = p->is_line; r->last = -2;
		{
			int pending = 0; /* rec holds a record that did not fit the previous batch */
			while (!eof) {
				int closed = 0;
				b = pipe_claim(p);
				b->seq.l = 0; b->n_seq = 0; b->file = i; b->open_failed = 0;
				while (pending || rd_record(r, &rec, &tmp) == 0) {
					const int64_t need = (int64_t)(rec.l + 1) * ((p->no_for ? 0 : 1) + (p->no_rev ? 0 : 1));
					if (!pending && p->hard > 0 && b->n_seq > 0 && (int64_t)b->seq.l + need > p->hard) { pending = 1; closed = 1; break; }
					pending = 0;
					seq_add(&b->seq, &rec, !p->no_for, !p->no_rev, &b->n_seq);
					if (p->batch > 0 && (int64_t)b->seq.l > p->batch) { closed = 1; break; } /* io.c:114,119 */
				}
				if (!closed) eof = 1;
				b->last_of_file = eof;
				pipe_publish(p);
			}
		}
		gzclose(r->fp);
		free(r);
	}
	pthread_mutex_lock(&p->mu);
	p->done = 1;
	pthread_cond_broadcast(&p->cv);
	pthread_mutex_unlock(&p->mu);
	free(rec.s); free(tmp.s);
	return 0;
}

/* ---- second stage: batches resident on the device, partial BWT built (two device buffers) ---- */

typedef struct {
	void *d; int64_t cap, len, n_seq;
	rb3b_batch_t *pb;   /* the batch prepared for merging: partial BWT + walk order (NULL for -2/-s/-r) */
	int file, last_of_file, open_failed;
} dbatch_t;

typedef struct {
	pipe_t *in;
	int use_rb2;        /* -2/-s/-r: the batch stays text, rb3b_insert_multi_dev sorts it itself */
	dbatch_t slot[2];
	int n_full, head, tail, done, failed;
	pthread_mutex_t mu; pthread_cond_t cv;
} dpipe_t;

static void *bwt_main(void *arg)
{ /* runs in its own library context (own stream): rb3_build_sais of batch i+1 overlaps the merge of batch i */
	dpipe_t *q = (dpipe_t*)arg;
	batch_t *b;
	while ((b = pipe_next(q->in)) != 0) {
		dbatch_t *o;
		pthread_mutex_lock(&q->mu);
		while (q->n_full == 2) pthread_cond_wait(&q->cv, &q->mu);
		pthread_mutex_unlock(&q->mu);
		o = &q->slot[q->tail];
		o->pb = 0;
		o->len = (int64_t)b->seq.l; o->n_seq = b->n_seq; o->file = b->file; o->last_of_file = b->last_of_file; o->open_failed = b->open_failed;
		if (!b->open_failed && b->n_seq > 0) {
			LOG("read %ld symbols from file '%s'", (long)b->seq.l, q->in->argv[b->file]);
			if (o->len > o->cap) {
				rb3b_dev_free(o->d);
				o->cap = o->len + (o->len >> 2);
				o->d = rb3b_dev_alloc(o->cap);
			}
			if (o->d == 0 || rb3b_h2d(o->d, b->seq.s, o->len) < 0) { fprintf(stderr, "ERROR: host to device copy: %s\n", rb3b_last_error()); q->failed = 1; }
			else if (!q->use_rb2) {
				o->pb = rb3b_batch_prepare_dev(o->len, (const uint8_t*)o->d); /* rb3_build_sais + the batch-only half of the rank phase */
				if (o->pb == 0) { fprintf(stderr, "ERROR: partial BWT: %s\n", rb3b_last_error()); q->failed = 1; }
				else LOG("constructed partial BWT for %ld symbols", (long)o->len);
			}
		}
		pipe_release(q->in);
		pthread_mutex_lock(&q->mu);
		q->tail ^= 1; ++q->n_full;
		pthread_cond_broadcast(&q->cv);
		pthread_mutex_unlock(&q->mu);
		if (q->failed) break;
	}
	pthread_mutex_lock(&q->mu);
	q->done = 1;
	pthread_cond_broadcast(&q->cv);
	pthread_mutex_unlock(&q->mu);
	return 0;
}

static dbatch_t *dpipe_next(dpipe_t *q)
{
	dbatch_t *b = 0;
	pthread_mutex_lock(&q->mu);
	while (q->n_full == 0 && !q->done) pthread_cond_wait(&q->cv, &q->mu);
	if (q->n_full) b = &q->slot[q->head];
	pthread_mutex_unlock(&q->mu);
	return b;
}

static void dpipe_release(dpipe_t *q)
{
	pthread_mutex_lock(&q->mu);
	q->head ^= 1; --q->n_full;
	pthread_cond_broadcast(&q->cv);
	pthread_mutex_unlock(&q->mu);
}

static int usage(FILE *fp)
{
	fprintf(fp, "Usage: ropebwt3-b200 build [options] <in.fa> [...]\n       ropebwt3-b200 merge [-o FILE] [-S FILE] <base.fmr> <other1.fmr> [...]\n       ropebwt3-b200 ssa [-s INT] [-o FILE] <in.fmd>\n");
	fprintf(fp, "Options (those of `ropebwt3 build`):\n");
	fprintf(fp, "  -m NUM   batch size [7G]             -i FILE  append to an existing .fmr/.fmd index\n");
	fprintf(fp, "  -L       one sequence per line       -F/-R    skip forward / reverse strand\n");
	fprintf(fp, "  -o FILE  output file [stdout]        -d/-b    write .fmd / .fmr (default: plain text BWT)\n");
	fprintf(fp, "  -S FILE  save the index (.fmr) after every input file\n");
	fprintf(fp, "  -l INT   leaf block length of the .fmr dump [512]   -n INT  max children per node [64]\n");
	fprintf(fp, "  -t INT / -p INT  accepted for compatibility; the device provides the parallelism\n");
	fprintf(fp, "  -2 / -s / -r   ropebwt2 insertion: input order / RLO / RCLO (mr_insert_multi)\n");
	fprintf(fp, "  -T -e    not supported by this engine (tree dump: there is no tree; BRE)\n");
	fprintf(fp, "Environment: RB3B_DEVICE=<cuda device>   RB3B_DEVICES=<d0,d1,...>  several devices: one replica of the index each,\n             the rank phase of every merge split among them (rb3b_merge_plain_dist over NCCL)\n");
	return fp == stdout ? 0 : 1;
}

#define DIE_IF(rc, what) do { if ((rc) < 0) { fprintf(stderr, "ERROR: %s: %s\n", what, rb3b_last_error()); exit(1); } } while (0)

/* ---- several devices (RB3B_DEVICES=d0,d1,...): one replica of the index per device, one host thread per device ----
 * Every thread is a rank of the library's NCCL communicator and takes part in every merge with rb3b_merge_plain_dist:
 * the rows of the batch are split among the devices for the rank phase, the interleave positions are combined over
 * NVLink, every replica applies the same merge.  The main thread is rank 0 (it also reads, sorts and writes); the others
 * only merge.  A failure on any rank ends the process: the other ranks would wait in a collective for ever. */
typedef struct {
	int world;
	unsigned char uid[128];
	const char *fn_in;            /* -i: every replica loads the index itself */
	int64_t est_symbols;
	pthread_mutex_t mu; pthread_cond_t cv;
	int64_t gen;                  /* batches posted so far */
	const uint8_t *bwt; int64_t len; /* the batch of generation `gen` (host memory, valid until everybody is done with it) */
	int quit, n_done, n_ready;
} mdev_t;

typedef struct { mdev_t *m; int rank, device; pthread_t tid; } mdev_worker_t;

#define MDEV_DIE(cond, what) do { if (cond) { fprintf(stderr, "ERROR: device %d: %s: %s\n", w->device, what, rb3b_last_error()); exit(1); } } while (0)

static void *mdev_main(void *arg)
{
	mdev_worker_t *w = (mdev_worker_t*)arg;
	mdev_t *m = w->m;
	rb3b_ctx_t *ctx = rb3b_ctx_create(w->device);
	rb3b_index_t *idx;
	int64_t seen = 0;
	int reserved = 0;
	MDEV_DIE(ctx == 0, "no context");
	MDEV_DIE(rb3b_ctx_make_current(ctx) < 0, "context");
	MDEV_DIE(rb3b_dist_init(w->rank, m->world, m->uid) < 0, "joining the communicator");
	idx = rb3b_index_create();
	if (m->fn_in) {
		int64_t acc0[7];
		MDEV_DIE(rb3b_restore(idx, m->fn_in) < 0, "loading the index");
		rb3b_index_reserve(idx, rb3b_get_acc(idx, acc0) + m->est_symbols); reserved = 1;
	}
	pthread_mutex_lock(&m->mu); ++m->n_ready; pthread_cond_broadcast(&m->cv); pthread_mutex_unlock(&m->mu);
	for (;;) {
		const uint8_t *bwt; int64_t len;
		pthread_mutex_lock(&m->mu);
		while (m->gen == seen && !m->quit) pthread_cond_wait(&m->cv, &m->mu);
		if (m->gen == seen) { pthread_mutex_unlock(&m->mu); break; } /* quit */
		bwt = m->bwt; len = m->len; seen = m->gen;
		pthread_mutex_unlock(&m->mu);
		MDEV_DIE(rb3b_merge_plain_dist(idx, len, bwt) < 0, "merging the partial BWT");
		if (!reserved) { rb3b_index_reserve(idx, m->est_symbols); reserved = 1; }
		pthread_mutex_lock(&m->mu); ++m->n_done; pthread_cond_broadcast(&m->cv); pthread_mutex_unlock(&m->mu);
	}
	rb3b_index_destroy(idx);
	rb3b_dist_finalize();
	rb3b_ctx_make_current(0);
	rb3b_ctx_destroy(ctx);
	return 0;
}

/* rank 0's side of one merge: hand the batch to the other ranks, take part, wait for them */
static int mdev_merge(mdev_t *m, rb3b_index_t *idx, int64_t len, const uint8_t *bwt)
{
	int rc;
	pthread_mutex_lock(&m->mu);
	m->bwt = bwt; m->len = len; m->n_done = 0; ++m->gen;
	pthread_cond_broadcast(&m->cv);
	pthread_mutex_unlock(&m->mu);
	rc = rb3b_merge_plain_dist(idx, len, bwt);
	pthread_mutex_lock(&m->mu);
	while (m->n_done < m->world - 1) pthread_cond_wait(&m->cv, &m->mu);
	pthread_mutex_unlock(&m->mu);
	return rc;
}

/* RB3B_DEVICES -> device ordinals; returns how many (0 or 1: single-device operation) */
static int mdev_parse(int *dev, int max)
{
	const char *e = getenv("RB3B_DEVICES");
	int n = 0;
	if (e == 0) return 0;
	while (*e && n < max) {
		char *q;
		long v = strtol(e, &q, 10);
		if (q == e) break;
		dev[n++] = (int)v;
		e = *q == ',' ? q + 1 : q;
	}
	return n;
}


int main(int argc, char *argv[])
{
	int c, i, is_line = 0, no_for = 0, no_rev = 0, fmt = 0 /* 0 plain, 1 fmd, 2 fmr */, block_len = 512, max_nodes = 64;
	int use_rb2 = 0, sort_order = 0; /* build.c:153-155 */
	int64_t batch = 7000000000LL, est_symbols = 0;
	int reserved = 0;
	const char *fn_in = 0, *fn_tmp = 0;
	int n_dev = 0, devs[16];       /* RB3B_DEVICES */
	mdev_t M;
	mdev_worker_t W[16];
	uint8_t *hb = 0; int64_t hb_cap = 0; /* several devices: the batch's BWT in pinned host memory */

	rb3b_index_t *idx = 0;

	t_real0 = realtime();
	if (argc >= 2 && strcmp(argv[1], "version") == 0) { puts(rb3b_version()); return 0; }
	if (argc >= 2 && strcmp(argv[1], "batches") == 0) { /* host-only: what the reader thread hands to the device, one line per batch */
		pipe_t P;
		pthread_t tid;
		batch_t *b;
		memset(&P, 0, sizeof(P));
		--argc; ++argv;
		int quiet = 0; /* -q: sizes and a checksum only (for timing the reader) */
		while ((c = getopt(argc, argv, "m:LFRq")) >= 0) {
			if (c == 'm') batch = parse_num(optarg);
			else if (c == 'q') quiet = 1;
			else if (c == 'L') is_line = 1;
			else if (c == 'F') no_for = 1;
			else if (c == 'R') no_rev = 1;
		}
		P.argc = argc; P.argv = argv; P.first = optind; P.is_line = is_line; P.no_for = no_for; P.no_rev = no_rev; P.batch = batch;
		pthread_mutex_init(&P.mu, 0); pthread_cond_init(&P.cv, 0);
		if (pthread_create(&tid, 0, reader_main, &P) != 0) return 1;
		while ((b = pipe_next(&P)) != 0) {
			size_t k;
			if (b->open_failed) printf("%d\t-1\t!\n", b->file - optind);
			else if (quiet) {
				uint64_t h = 0, w;
				for (k = 0; k + 8 <= b->seq.l; k += 8) { memcpy(&w, b->seq.s + k, 8); h = (h << 1 | h >> 63) ^ w; } /* cheap: the reader must stay the bottleneck */
				for (; k < b->seq.l; ++k) h = (h << 1 | h >> 63) ^ (unsigned char)b->seq.s[k];
				printf("%d\t%ld\t%ld symbols, checksum %016llx\t%d\n", b->file - optind, (long)b->n_seq, (long)b->seq.l, (unsigned long long)h, b->last_of_file);
			} else {
				printf("%d\t%ld\t", b->file - optind, (long)b->n_seq);
				for (k = 0; k < b->seq.l; ++k) putchar("$ACGTN"[(unsigned char)b->seq.s[k] < 6 ? (unsigned char)b->seq.s[k] : 5]);
				printf("\t%d\n", b->last_of_file);
			}
			pipe_release(&P);
		}
		pthread_join(tid, 0);
		return 0;
	}
	if (argc >= 2 && strcmp(argv[1], "ssa") == 0) { /* main_ssa, ssa.c:247-279 */
		int ss = 8;
		const char *fn_out = "-";
		--argc; ++argv;
		while ((c = getopt(argc, argv, "t:s:o:")) >= 0) {
			if (c == 's') ss = atoi(optarg);
			else if (c == 'o') fn_out = optarg;
		}
		if (argc == optind) {
			fprintf(stderr, "Usage: ropebwt3-b200 ssa [options] <in.fmd>\nOptions:\n  -t INT     accepted for compatibility\n");
			fprintf(stderr, "  -s INT     sample rate one SA per 2**INT bases [8]\n  -o FILE    output to file [stdout]\n");
			return 1;
		}
		DIE_IF(rb3b_init(getenv("RB3B_DEVICE") ? atoi(getenv("RB3B_DEVICE")) : 0), "no usable CUDA device");
		idx = rb3b_index_create();
		if (rb3b_restore(idx, argv[optind]) < 0) { fprintf(stderr, "[E::main_ssa] failed to load the FM-index\n"); return 1; }
		DIE_IF(rb3b_ssa_dump(idx, ss, fn_out), "sampled suffix array");
		rb3b_index_destroy(idx);
		return 0;
	}
	if (argc >= 2 && strcmp(argv[1], "merge") == 0) { /* main_merge, main.c:84-133 */
		rb3b_index_t *other;
		--argc; ++argv;
		while ((c = getopt(argc, argv, "t:o:S:")) >= 0) {
			if (c == 'o') { if (freopen(optarg, "wb", stdout) == 0) { fprintf(stderr, "ERROR: failed to open '%s' for writing\n", optarg); return 1; } }
			else if (c == 'S') fn_tmp = optarg;
		}
		if (argc - optind < 2) {
			fprintf(stdout, "Usage: ropebwt3-b200 merge [options] <base.fmr> <other1.fmr> [...]\nOptions:\n  -t INT     accepted for compatibility\n");
			fprintf(stdout, "  -o FILE    output FMR to FILE [stdout]\n  -S FILE    save the current index to FILE after each input file []\n");
			return 1;
		}
		DIE_IF(rb3b_init(getenv("RB3B_DEVICE") ? atoi(getenv("RB3B_DEVICE")) : 0), "no usable CUDA device");
		idx = rb3b_index_create();
		if (rb3b_restore(idx, argv[optind]) < 0) { fprintf(stderr, "ERROR: failed to load FMR file '%s'\n", argv[optind]); return 1; }
		for (i = optind + 1; i < argc; ++i) {
			other = rb3b_index_create();
			if (rb3b_restore(other, argv[i]) < 0) { fprintf(stderr, "ERROR: failed to load FMR/FMD file '%s'\n", argv[i]); rb3b_index_destroy(other); break; }
			DIE_IF(rb3b_merge_index(idx, other), "merging"); /* rb3_fmi_merge(r, &fb, n_threads, 1) */
			rb3b_index_destroy(other);
			LOG("merged '%s'", argv[i]);
			if (fn_tmp) { DIE_IF(rb3b_dump_fmr(idx, fn_tmp, max_nodes, block_len), "saving the index"); LOG("saved the current index to '%s'", fn_tmp); }
		}
		DIE_IF(rb3b_dump_fmr(idx, "-", max_nodes, block_len), "writing .fmr"); /* mr_dump(r, stdout) */
		rb3b_index_destroy(idx);
		return 0;
	}
	if (argc < 2 || strcmp(argv[1], "build") != 0) return usage(stderr);
	--argc; ++argv;
	while ((c = getopt(argc, argv, "l:n:m:t:2sri:LFRo:dbTS:p:e")) >= 0) {
		if (c == 'm') batch = parse_num(optarg);
		else if (c == 't' || c == 'p') {}
		else if (c == 'l') block_len = atoi(optarg);
		else if (c == 'n') max_nodes = atoi(optarg);
		else if (c == 'i') fn_in = optarg;
		else if (c == 'L') is_line = 1;
		else if (c == 'F') no_for = 1;
		else if (c == 'R') no_rev = 1;
		else if (c == 'o') { if (freopen(optarg, "wb", stdout) == 0) { fprintf(stderr, "ERROR: failed to open '%s' for writing\n", optarg); return 1; } }
		else if (c == 'd') fmt = 1;
		else if (c == 'b') fmt = 2;
		else if (c == 'S') fn_tmp = optarg;
		else if (c == '2') use_rb2 = 1;
		else if (c == 's') use_rb2 = 1, sort_order = 1;
		else if (c == 'r') use_rb2 = 1, sort_order = 2;
		else if (c == 'T' || c == 'e') {
			fprintf(stderr, "ERROR: option -%c (tree dump / BRE) is outside this engine's scope; use the CPU ropebwt3 for it\n", c);
			return 1;
		} else return usage(stderr);
	}
	if (argc == optind && fn_in == 0) return usage(stderr);
	/* size hint for the device index: about two symbols (both strands) per input byte, more for compressed input; the
	 * ping-pong buffers are then allocated once instead of being regrown by the merges (rb3b_index_reserve) */
	for (i = optind; i < argc; ++i) {
		struct stat st;
		size_t l = strlen(argv[i]);
		if (stat(argv[i], &st) == 0) est_symbols += (int64_t)st.st_size * ((l > 3 && strcmp(argv[i] + l - 3, ".gz") == 0) ? 8 : 2);
	}
	if (est_symbols > 40000000000LL) est_symbols = 40000000000LL;
	if (no_for && no_rev) { fprintf(stderr, "ERROR: -F and -R together leave nothing to index\n"); return 1; }

	n_dev = use_rb2 ? 0 : mdev_parse(devs, 16); /* -2/-s/-r: one device */
	DIE_IF(rb3b_init(n_dev > 1 ? devs[0] : getenv("RB3B_DEVICE") ? atoi(getenv("RB3B_DEVICE")) : 0), "no usable CUDA device");
	if (n_dev > 1) {
		/* NCCL writes its version banner (NCCL_DEBUG=VERSION) to stdout, where the index goes: keep it out */
		const char *nd = getenv("NCCL_DEBUG");
		if (nd == 0 || strcmp(nd, "VERSION") == 0 || strcmp(nd, "version") == 0) setenv("NCCL_DEBUG", "WARN", 1);
		setenv("NCCL_DEBUG_FILE", "/dev/stderr", 0);
		memset(&M, 0, sizeof(M));
		M.world = n_dev; M.fn_in = fn_in; M.est_symbols = est_symbols;
		pthread_mutex_init(&M.mu, 0); pthread_cond_init(&M.cv, 0);
		DIE_IF(rb3b_dist_unique_id(M.uid), "NCCL");
		for (i = 1; i < n_dev; ++i) {
			W[i].m = &M; W[i].rank = i; W[i].device = devs[i];
			if (pthread_create(&W[i].tid, 0, mdev_main, &W[i]) != 0) { fprintf(stderr, "ERROR: failed to start the thread of device %d\n", devs[i]); return 1; }
		}
		DIE_IF(rb3b_dist_init(0, n_dev, M.uid), "joining the communicator");
		LOG("%d devices, one replica of the index each", n_dev);
	}
	if (fn_in) {
		idx = rb3b_index_create();
		if (rb3b_restore(idx, fn_in) < 0) { /* build.c:175-178 */
			fprintf(stderr, "ERROR: failed to open index file '%s' (%s)\n", fn_in, rb3b_last_error());
			return 1;
		}
		LOG("loaded the index from file '%s'", fn_in);
		{ int64_t acc0[7]; rb3b_index_reserve(idx, rb3b_get_acc(idx, acc0) + est_symbols); reserved = 1; }
	}
	{
		pipe_t P;
		dpipe_t Q;
		pthread_t tid, tid2;
		dbatch_t *db;
		int64_t acc[7], fit;
		memset(&P, 0, sizeof(P));
		P.argc = argc; P.argv = argv; P.first = optind; P.is_line = is_line; P.no_for = no_for; P.no_rev = no_rev;
		/* clamp the batch to what the device can sort and merge next to the finished index; the output does not depend
		 * on the batching */
		fit = rb3b_max_batch_symbols((idx ? rb3b_get_acc(idx, acc) : 0) + est_symbols);
		if (use_rb2 && (sort_order || (idx && rb3b_index_get_order(idx)))) fit /= 2; /* RLO/RCLO batches are sorted as augmented strings of twice the length */
		if (fit > 0 && (batch <= 0 || batch > fit)) {
			LOG("batch size limited to %ld symbols by device memory", (long)fit);
			batch = fit;
		}
		P.batch = batch; P.hard = fit;
		pthread_mutex_init(&P.mu, 0); pthread_cond_init(&P.cv, 0);
		if (pthread_create(&tid, 0, reader_main, &P) != 0) { fprintf(stderr, "ERROR: failed to start the reader thread\n"); return 1; }
		memset(&Q, 0, sizeof(Q));
		Q.in = &P; Q.use_rb2 = use_rb2;
		pthread_mutex_init(&Q.mu, 0); pthread_cond_init(&Q.cv, 0);
		if (pthread_create(&tid2, 0, bwt_main, &Q) != 0) { fprintf(stderr, "ERROR: failed to start the BWT thread\n"); return 1; }
		while ((db = dpipe_next(&Q)) != 0) {
			const char *fn = argv[db->file];
			if (Q.failed) return 1;
			if (db->open_failed) fprintf(stderr, "ERROR: failed to open file '%s'\n", fn);
			else if (db->n_seq > 0) {
				if (use_rb2) { /* build.c:214-218; an index loaded with -i keeps its own order, like mr->so */
					if (idx == 0) {
						idx = rb3b_index_create();
						DIE_IF(rb3b_index_set_order(idx, sort_order), "sorting order");
					}
					DIE_IF(rb3b_insert_multi_dev(idx, db->len, (const uint8_t*)db->d), "inserting the batch");
					LOG("inserted %ld symbols", (long)db->len);
				} else if (n_dev > 1) { /* every device takes part: the batch's BWT goes through host memory to all of them */
					if (db->len > hb_cap) {
						rb3b_host_free_pinned(hb);
						hb_cap = db->len + (db->len >> 2);
						hb = (uint8_t*)rb3b_host_alloc_pinned(hb_cap);
						if (hb == 0) { fprintf(stderr, "ERROR: pinned host memory: %s\n", rb3b_last_error()); return 1; }
					}
					DIE_IF(rb3b_d2h(hb, rb3b_batch_bwt_dev(db->pb), db->len), "device to host copy");
					if (idx == 0) idx = rb3b_index_create();
					DIE_IF(mdev_merge(&M, idx, db->len, hb), "merging the partial BWT"); /* rb3_fmi_merge_plain (rb3_enc_plain2fmr on an empty index) */
					LOG("merged the partial BWT for %ld symbols on %d devices", (long)db->len, n_dev);
				} else if (idx == 0) {
					idx = rb3b_index_create();
					DIE_IF(rb3b_merge_prepared(idx, db->pb), "encoding the partial BWT"); /* rb3_enc_plain2fmr: the index is empty */
					LOG("encoded the partial BWT for %ld symbols", (long)db->len);
				} else {
					DIE_IF(rb3b_merge_prepared(idx, db->pb), "merging the partial BWT"); /* rb3_fmi_merge_plain */
					LOG("merged the partial BWT for %ld symbols", (long)db->len);
				}
				if (db->pb) { rb3b_batch_destroy(db->pb); db->pb = 0; }
				/* after the first batch: building the first index resets the buffers */
				if (!reserved) { rb3b_index_reserve(idx, est_symbols); reserved = 1; }
			}
			if (!db->open_failed && db->last_of_file && fn_tmp && idx) { /* build.c:232-238 */
				DIE_IF(rb3b_dump_fmr(idx, fn_tmp, max_nodes, block_len), "saving the index");
				LOG("saved the current index to '%s'", fn_tmp);
			}
			dpipe_release(&Q);
		}
		LOG("all batches merged");
		if (n_dev > 1) {
			pthread_mutex_lock(&M.mu); M.quit = 1; pthread_cond_broadcast(&M.cv); pthread_mutex_unlock(&M.mu);
			for (i = 1; i < n_dev; ++i) pthread_join(W[i].tid, 0);
			rb3b_host_free_pinned(hb);
		}
		pthread_join(tid2, 0);
		if (Q.failed) return 1;
		rb3b_dev_free(Q.slot[0].d); rb3b_dev_free(Q.slot[1].d);
		pthread_join(tid, 0);
		free(P.slot[0].seq.s); free(P.slot[1].seq.s);
	}
	if (idx == 0) return 1; /* build.c:243 */
	if (fmt == 2) DIE_IF(rb3b_dump_fmr(idx, "-", max_nodes, block_len), "writing .fmr");
	else if (fmt == 1) DIE_IF(rb3b_dump_fmd(idx, "-"), "writing .fmd");
	else DIE_IF(rb3b_dump_plain(idx, "-"), "writing the BWT");
	LOG("wrote the index");
	rb3b_index_destroy(idx);
	fprintf(stderr, "[M::main] Real time: %.3f sec; CPU: %.3f sec\n", realtime() - t_real0, cputime());
	return 0;
}
