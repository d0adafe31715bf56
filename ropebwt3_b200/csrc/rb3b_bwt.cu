/*
 * rb3b_bwt.cu -- partial BWT of one batch on the device, the producer of the
 * merge path's input.  Same contract as rb3_build_sais (sais-ss.c:10-56): the
 * text is a concatenation of 0-terminated nt6 strings; suffixes are compared up
 * to and including their first sentinel, an earlier sentinel being smaller
 * (libsais GSA order); BWT[i] = T[SA[i]-1], or T[len-1] when SA[i] == 0.
 *
 * The algorithm is not the reference's (libsais is induced sorting on the
 * host): this is prefix doubling on top of a device radix sort.  Round 0 sorts
 * all suffixes by their first 21 symbols packed 3 bits each (cut after a
 * sentinel; the stable sort orders equal sentinel-terminated prefixes by
 * position, which is exactly the sentinel order); every later round doubles
 * the compared length by sorting the still ambiguous suffixes, group by
 * group, by rank[i+h] (a segmented sort over a compact list of them).
 *
 * Also here: rb3b_merge_index, the BWT-vs-BWT flavour (rb3_fmi_merge,
 * fm-index.c:251-277), which expands the other index and reuses merge_plain.
 */
#include <string.h>
#include <cub/cub.cuh>
#include "rb3b_internal.cuh"

#define TPB 256
#define KMER 21

static inline unsigned nblk(int64_t n, int per) { return (unsigned)((n + per - 1) / per); }

__global__ void k_kmer_keys(uint32_t n, const uint8_t *__restrict__ T, int n_sym, uint64_t *__restrict__ key, uint32_t *__restrict__ idx, int *__restrict__ bad)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t k = 0;
	int ended = 0;
	for (int j = 0; j < KMER; ++j) {
		uint32_t p = i + j;
		int c = (!ended && p < n) ? T[p] : 0;
		if (c >= n_sym) { *bad = 1; c = 5; }
		k = k << 3 | (uint64_t)c;
		if (c == 0) ended = 1;
	}
	key[i] = k << 1 | (uint64_t)ended; /* bit 0: the prefix holds a sentinel, i.e. the suffix is fully compared */
	idx[i] = i;
}

/* head[j] = 1 when sorted element j starts a new group; a finished suffix is always its own group */
__global__ void k_group_heads(uint32_t n, const uint64_t *__restrict__ key, int first_round, uint32_t *__restrict__ head)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	bool h = j == 0 || key[j] != key[j - 1] || (first_round && (key[j] & 1));
	head[j] = h ? j : 0;
}

/* ---- round 0 of small batches (< 2^24 symbols: one or two bacterial genomes): keys-only radix sort ----
 * The suffix position rides in the low 24 bits of the key, so no value array travels through the sort (16 instead of 24
 * bytes per suffix and pass); only the 40 bits above it are sorted -- 5 passes instead of 8.  Those 40 bits hold the first
 * 15 symbols as a base-6 number (6^15 < 2^39; same order as the 3-bit packing) and the "fully compared" flag.  The stable
 * sort leaves equal prefixes in position order, as before. */
#define KMER6 15
#define POS_BITS 24

__global__ void k_kmer_keys6(uint32_t n, const uint8_t *__restrict__ T, int n_sym, uint64_t *__restrict__ key, int *__restrict__ bad)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t k = 0;
	int ended = 0;
	for (int j = 0; j < KMER6; ++j) {
		uint32_t p = i + j;
		int c = (!ended && p < n) ? T[p] : 0;
		if (c >= n_sym) { *bad = 1; c = 5; }
		k = k * 6 + (uint64_t)c;
		if (c == 0) ended = 1;
	}
	key[i] = (k << 1 | (uint64_t)ended) << POS_BITS | (uint64_t)i;
}

/* the same keys, eight consecutive positions per thread from three aligned 8-byte loads (T must be 8-byte aligned): the
 * byte loads of k_kmer_keys6 -- fifteen per suffix -- were most of its time */
__global__ void k_kmer_keys6x8(uint32_t n, const uint8_t *__restrict__ T, int n_sym, uint64_t *__restrict__ key, int *__restrict__ bad)
{
	const uint32_t i0 = (blockIdx.x * blockDim.x + threadIdx.x) * 8u;
	if (i0 >= n) return;
	uint64_t w[3];
#pragma unroll
	for (int k = 0; k < 3; ++k) {
		const uint32_t p = i0 + 8u * k;
		if (p + 8u <= n) w[k] = __ldg((const uint64_t*)(T + p));
		else { /* the last words of the text: byte by byte, nothing past the end */
			w[k] = 0;
			for (int b = 0; b < 8; ++b) if (p + b < n) w[k] |= (uint64_t)T[p + b] << (8 * b);
		}
	}
	int any_bad = 0;
#pragma unroll
	for (int r = 0; r < 8; ++r) {
		uint64_t k = 0;
		int ended = 0;
#pragma unroll
		for (int j = 0; j < KMER6; ++j) {
			const int q = r + j; /* byte q of the 24 loaded ones */
			int c = ended ? 0 : (int)((w[q >> 3] >> (8 * (q & 7))) & 0xff);
			if (c >= n_sym) { any_bad = 1; c = 5; }
			k = k * 6 + (uint64_t)c;
			if (c == 0) ended = 1;
		}
		if (i0 + r < n) key[i0 + r] = (k << 1 | (uint64_t)ended) << POS_BITS | (uint64_t)(i0 + r);
	}
	if (any_bad) *bad = 1;
}

/* group heads of the keys-only round + the suffix array stretch it implies */
__global__ void k_group_heads6(uint32_t n, const uint64_t *__restrict__ key, uint32_t *__restrict__ head, uint32_t *__restrict__ sa)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const uint64_t k = key[j] >> POS_BITS;
	bool h = j == 0 || k != key[j - 1] >> POS_BITS || (k & 1);
	head[j] = h ? j : 0;
	sa[j] = (uint32_t)(key[j] & ((1u << POS_BITS) - 1u));
}

/* after the max-scan head[j] is the index of the first element of j's group = the new rank */
__global__ void k_assign_rank(uint32_t n, const uint32_t *__restrict__ idx, const uint32_t *__restrict__ head, uint32_t *__restrict__ rank, unsigned long long *n_amb)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	rank[idx[j]] = head[j];
	bool amb = head[j] != j || (j + 1 < n && head[j + 1] == head[j]);
	unsigned m = __ballot_sync(__activemask(), amb);
	if (amb && (threadIdx.x & 31) == (unsigned)(__ffs(m) - 1)) atomicAdd(n_amb, (unsigned long long)__popc(m));
}

__global__ void k_pair_keys(uint32_t n, uint32_t h, const uint32_t *__restrict__ rank, uint64_t *__restrict__ key, uint32_t *__restrict__ idx)
{
	uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	uint64_t r2 = (uint64_t)i + h < n ? rank[i + h] : 0;
	key[i] = (uint64_t)rank[i] << 32 | r2;
	idx[i] = i;
}

__global__ void k_sa_to_bwt(uint32_t n, const uint8_t *__restrict__ T, const uint32_t *__restrict__ sa, uint8_t *__restrict__ bwt)
{
	uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	uint32_t p = sa[j];
	bwt[j] = T[p == 0 ? n - 1 : p - 1];
}

static int sort_pairs(uint64_t *k_in, uint64_t *k_out, uint32_t *v_in, uint32_t *v_out, uint32_t n, int end_bit)
{
	size_t tmp = 0;
	CK(cub::DeviceRadixSort::SortPairs((void*)0, tmp, k_in, k_out, v_in, v_out, (int64_t)n, 0, end_bit, rb3b_stream));
	DBuf<uint8_t> t;
	TRY(t.alloc(tmp));
	CK(cub::DeviceRadixSort::SortPairs((void*)t.p, tmp, k_in, k_out, v_in, v_out, (int64_t)n, 0, end_bit, rb3b_stream));
	return RB3B_OK;
}

static int scan_max_u32(uint32_t *d, uint32_t n)
{
	size_t tmp = 0;
	CK(cub::DeviceScan::InclusiveScan((void*)0, tmp, d, d, cub::Max(), (int64_t)n, rb3b_stream));
	DBuf<uint8_t> t;
	TRY(t.alloc(tmp));
	CK(cub::DeviceScan::InclusiveScan((void*)t.p, tmp, d, d, cub::Max(), (int64_t)n, rb3b_stream));
	return RB3B_OK;
}

/* ---- refinement rounds on the still ambiguous suffixes only ----
 * After round 0 the sorted order is final except inside groups of suffixes with equal 21-symbol prefixes.  The ambiguous
 * suffixes are kept as a compact list in sorted order: suf[t] (the suffix), grp[t] (its group = the sorted position of the
 * group's first member = its current rank).  A round sorts every group by rank[suf + h] (one segmented sort: the groups
 * are the segments), gives every sub-group its new rank, writes the now final stretch of the suffix array back and keeps
 * only the members of sub-groups that are still larger than one. */

struct HeadIsAmbiguous { /* sorted position j belongs to a group of more than one suffix */
	const uint32_t *head; uint32_t n;
	__host__ __device__ bool operator()(const uint32_t &j) const { return head[j] != j || (j + 1 < n && head[j + 1] == head[j]); }
};

__global__ void k_amb_gather(uint32_t m, const uint32_t *__restrict__ at, const uint32_t *__restrict__ sa, const uint32_t *__restrict__ head, uint32_t *__restrict__ suf, uint32_t *__restrict__ grp)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= m) return;
	const uint32_t j = at[t];
	suf[t] = sa[j]; grp[t] = head[j];
}

/* second key of every listed suffix, and the segment starts (first list index of every group) */
__global__ void k_list_keys(uint32_t m, uint32_t n, uint32_t h, const uint32_t *__restrict__ suf, const uint32_t *__restrict__ grp, const uint32_t *__restrict__ rank,
                            uint32_t *__restrict__ key, uint8_t *__restrict__ segflag)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= m) return;
	const uint64_t q = (uint64_t)suf[t] + h;
	key[t] = q < n ? rank[q] : 0;
	segflag[t] = t == 0 || grp[t] != grp[t - 1];
}

/* after the segmented sort: sub-group heads.  pos[t] = grp[t] + (t - first list index of the group) is the sorted position
 * of list element t; sub[t] = pos of the first element of its sub-group (max-scan of the heads' positions) */
__global__ void k_sub_heads(uint32_t m, const uint32_t *__restrict__ grp, const uint32_t *__restrict__ key, const uint32_t *__restrict__ segfirst, uint32_t *__restrict__ sub)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= m) return;
	const bool headp = t == 0 || grp[t] != grp[t - 1] || key[t] != key[t - 1];
	sub[t] = headp ? grp[t] + (t - segfirst[t]) : 0;
}

__global__ void k_seg_first(uint32_t m, const uint32_t *__restrict__ grp, uint32_t *__restrict__ segfirst)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= m) return;
	segfirst[t] = (t == 0 || grp[t] != grp[t - 1]) ? t : 0;
}

/* write back: the suffix array stretch, the new ranks, and who stays on the list */
__global__ void k_list_apply(uint32_t m, const uint32_t *__restrict__ suf, const uint32_t *__restrict__ grp, const uint32_t *__restrict__ segfirst, const uint32_t *__restrict__ sub,
                             uint32_t *__restrict__ sa, uint32_t *__restrict__ rank, uint8_t *__restrict__ keep)
{
	uint32_t t = blockIdx.x * blockDim.x + threadIdx.x;
	if (t >= m) return;
	const uint32_t pos = grp[t] + (t - segfirst[t]), s = sub[t], i = suf[t];
	sa[pos] = i;
	if (s != grp[t]) rank[i] = s; /* the first sub-group keeps the group's rank */
	const bool next_same = t + 1 < m && grp[t + 1] == grp[t] && sub[t + 1] == s;
	keep[t] = s != pos || next_same;
}

template<typename T> static int select_flagged(const T *in, const uint8_t *flag, T *out, uint32_t n, unsigned long long *d_count)
{
	size_t tmp = 0;
	CK(cub::DeviceSelect::Flagged((void*)0, tmp, in, flag, out, d_count, (int64_t)n, rb3b_stream));
	DBuf<uint8_t> t;
	TRY(t.alloc(tmp));
	CK(cub::DeviceSelect::Flagged((void*)t.p, tmp, in, flag, out, d_count, (int64_t)n, rb3b_stream));
	return RB3B_OK;
}

/* suffix array of the batch (generalised, sentinel order by position) into sa[len] (device, uint32) */
/* rank_keep != 0: the final rank of every suffix (the inverse suffix array) is left there */
static int suffix_sort(int64_t len, const uint8_t *d_text, DBuf<uint32_t> &sa, int n_sym = RB3B_ASIZE, DBuf<uint32_t> *rank_keep = 0)
{
	if (len >= (1LL << 32) - 2) return rb3b_fail(RB3B_EINVAL, "batches of 2^32 symbols or more are not supported by the device suffix sorter yet");
	uint32_t n = (uint32_t)len;
	uint8_t last = 1;
	CK(cudaMemcpyAsync(&last, d_text + len - 1, 1, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (last != 0) return rb3b_fail(RB3B_EINVAL, "the batch text must end with a sentinel (mrope.c:310 asserts the same)");
	DBuf<uint64_t> key0, key1;
	DBuf<uint32_t> idx0, rank_own, head;
	DBuf<uint32_t> &rank = rank_keep ? *rank_keep : rank_own;
	DBuf<int> bad;
	DBuf<unsigned long long> amb;
	DBuf<uint8_t> flag;
	TRY(key0.alloc(n)); TRY(key1.alloc(n)); TRY(idx0.alloc(n)); TRY(sa.alloc(n)); TRY(rank.alloc(n)); TRY(head.alloc(n)); TRY(bad.alloc(1)); TRY(amb.alloc(1)); TRY(flag.alloc(n));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	/* round 0: all suffixes by their first 21 symbols (15 for small batches, keys only) */
	const bool keys_only = n < (1u << POS_BITS) && n_sym <= 6 && rb3b_get_param("sa_keys_only", 1) != 0; /* base 6: not for the augmented alphabet of the sorted orders */
	const uint64_t kmer = keys_only ? KMER6 : KMER;
	if (keys_only) {
		if (((uintptr_t)d_text & 7) == 0 && rb3b_get_param("sa_keys_x8", 1) != 0) { k_kmer_keys6x8<<<nblk((n + 7) / 8, TPB), TPB, 0, rb3b_stream>>>(n, d_text, n_sym, key0.p, bad.p); CKK(); }
		else { k_kmer_keys6<<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, d_text, n_sym, key0.p, bad.p); CKK(); }
		size_t tb = 0;
		CK(cub::DeviceRadixSort::SortKeys((void*)0, tb, key0.p, key1.p, (int64_t)n, POS_BITS, 64, rb3b_stream));
		DBuf<uint8_t> t;
		TRY(t.alloc(tb));
		CK(cub::DeviceRadixSort::SortKeys((void*)t.p, tb, key0.p, key1.p, (int64_t)n, POS_BITS, 64, rb3b_stream));
		k_group_heads6<<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, key1.p, head.p, sa.p); CKK();
	} else {
		k_kmer_keys<<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, d_text, n_sym, key0.p, idx0.p, bad.p); CKK();
		TRY(sort_pairs(key0.p, key1.p, idx0.p, sa.p, n, 64));
		k_group_heads<<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, key1.p, 1, head.p); CKK();
	}
	TRY(scan_max_u32(head.p, n));
	CK(cudaMemsetAsync(amb.p, 0, 8, rb3b_stream));
	k_assign_rank<<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, sa.p, head.p, rank.p, amb.p); CKK();
	unsigned long long n_amb = 0;
	int hbad = 0, rounds = 1;
	CK(cudaMemcpyAsync(&n_amb, amb.p, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (hbad) return rb3b_fail(RB3B_EINVAL, "batch text holds a symbol >= %d", RB3B_ASIZE);
	rb3b_stat_set("sa_ambiguous_after_round0", (int64_t)n_amb);
	if (n_amb > 0 && n_amb < (1ULL << 31) && rb3b_get_param("sa_discard", 1) != 0) {
		/* the list of ambiguous suffixes: reuse the round-0 buffers (the 64-bit key arrays hold four 32-bit arrays) */
		uint32_t m = (uint32_t)n_amb;
		uint32_t *suf = (uint32_t*)key0.p, *grp = suf + n, *key = (uint32_t*)key1.p, *ksorted = key + n; /* each has room for n */
		DBuf<uint32_t> vsorted, segfirst, sub, off, tmp32;
		TRY(vsorted.alloc(m)); TRY(segfirst.alloc(m)); TRY(sub.alloc(m)); TRY(off.alloc((size_t)m + 1)); TRY(tmp32.alloc(m));
		{ /* sorted positions of the ambiguous suffixes (one select), then their suffixes and groups (one small gather) */
			cub::CountingInputIterator<uint32_t> pos(0);
			HeadIsAmbiguous is_amb; is_amb.head = head.p; is_amb.n = n;
			size_t tb = 0;
			CK(cub::DeviceSelect::If((void*)0, tb, pos, vsorted.p, amb.p, (int64_t)n, is_amb, rb3b_stream));
			DBuf<uint8_t> t;
			TRY(t.alloc(tb));
			CK(cub::DeviceSelect::If((void*)t.p, tb, pos, vsorted.p, amb.p, (int64_t)n, is_amb, rb3b_stream));
			k_amb_gather<<<nblk(m, TPB), TPB, 0, rb3b_stream>>>(m, vsorted.p, sa.p, head.p, suf, grp); CKK();
		}
		for (uint64_t h = kmer; m > 0 && h < n; h <<= 1) {
			/* segments = groups */
			k_list_keys<<<nblk(m, TPB), TPB, 0, rb3b_stream>>>(m, n, (uint32_t)h, suf, grp, rank.p, key, flag.p); CKK();
			k_seg_first<<<nblk(m, TPB), TPB, 0, rb3b_stream>>>(m, grp, segfirst.p); CKK();
			TRY(scan_max_u32(segfirst.p, m));
			/* segment offsets: the list indices where a group starts, then m */
			{
				size_t tb = 0;
				cub::CountingInputIterator<uint32_t> cnt(0);
				CK(cub::DeviceSelect::Flagged((void*)0, tb, cnt, flag.p, off.p, amb.p, (int64_t)m, rb3b_stream));
				DBuf<uint8_t> t;
				TRY(t.alloc(tb));
				CK(cub::DeviceSelect::Flagged((void*)t.p, tb, cnt, flag.p, off.p, amb.p, (int64_t)m, rb3b_stream));
			}
			unsigned long long n_seg = 0;
			CK(cudaMemcpyAsync(&n_seg, amb.p, 8, cudaMemcpyDeviceToHost, rb3b_stream));
			CK(cudaStreamSynchronize(rb3b_stream));
			CK(cudaMemcpyAsync(off.p + n_seg, &m, 4, cudaMemcpyHostToDevice, rb3b_stream));
			{
				size_t tb = 0;
				CK(cub::DeviceSegmentedSort::SortPairs((void*)0, tb, key, ksorted, suf, vsorted.p, (int)m, (int)n_seg, off.p, off.p + 1, rb3b_stream));
				DBuf<uint8_t> t;
				TRY(t.alloc(tb));
				CK(cub::DeviceSegmentedSort::SortPairs((void*)t.p, tb, key, ksorted, suf, vsorted.p, (int)m, (int)n_seg, off.p, off.p + 1, rb3b_stream));
			}
			k_sub_heads<<<nblk(m, TPB), TPB, 0, rb3b_stream>>>(m, grp, ksorted, segfirst.p, sub.p); CKK();
			TRY(scan_max_u32(sub.p, m));
			k_list_apply<<<nblk(m, TPB), TPB, 0, rb3b_stream>>>(m, vsorted.p, grp, segfirst.p, sub.p, sa.p, rank.p, flag.p); CKK();
			/* the survivors, with their sub-group as the new group */
			TRY(select_flagged<uint32_t>(vsorted.p, flag.p, suf, m, amb.p));
			TRY(select_flagged<uint32_t>(sub.p, flag.p, tmp32.p, m, amb.p));
			unsigned long long m2 = 0;
			CK(cudaMemcpyAsync(&m2, amb.p, 8, cudaMemcpyDeviceToHost, rb3b_stream));
			CK(cudaStreamSynchronize(rb3b_stream));
			if (m2) CK(cudaMemcpyAsync(grp, tmp32.p, (size_t)m2 * 4, cudaMemcpyDeviceToDevice, rb3b_stream));
			m = (uint32_t)m2;
			++rounds;
		}
		rb3b_stat_set("sa_rounds", rounds);
		return RB3B_OK;
	}
	/* every round re-sorts all suffixes (lists of 2^31 ambiguous suffixes or more; "sa_discard" = 0) */
	for (uint64_t h = kmer; n_amb > 0 && h < n; h <<= 1) {
		k_pair_keys<<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, (uint32_t)h, rank.p, key0.p, idx0.p); CKK();
		TRY(sort_pairs(key0.p, key1.p, idx0.p, sa.p, n, 64));
		k_group_heads<<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, key1.p, 0, head.p); CKK();
		TRY(scan_max_u32(head.p, n));
		CK(cudaMemsetAsync(amb.p, 0, 8, rb3b_stream));
		k_assign_rank<<<nblk(n, TPB), TPB, 0, rb3b_stream>>>(n, sa.p, head.p, rank.p, amb.p); CKK();
		CK(cudaMemcpyAsync(&n_amb, amb.p, 8, cudaMemcpyDeviceToHost, rb3b_stream));
		CK(cudaStreamSynchronize(rb3b_stream));
		++rounds;
	}
	rb3b_stat_set("sa_rounds", rounds);
	return RB3B_OK;
}

extern "C" int rb3b_build_bwt_dev(int64_t len, const uint8_t *d_text, uint8_t *d_bwt_out)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (len <= 0) return rb3b_fail(RB3B_EINVAL, "empty batch");
	DBuf<uint32_t> sa;
	DBuf<uint8_t> tmp;
	rb3b_tic(T_BWT);
	TRY(suffix_sort(len, d_text, sa));
	uint8_t *dst = d_bwt_out;
	if (d_bwt_out == d_text) { TRY(tmp.alloc(len)); dst = tmp.p; } /* in place, like the reference */
	k_sa_to_bwt<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>((uint32_t)len, d_text, sa.p, dst); CKK();
	if (dst != d_bwt_out) CK(cudaMemcpyAsync(d_bwt_out, dst, len, cudaMemcpyDeviceToDevice, rb3b_stream));
	rb3b_toc(T_BWT);
	CK(cudaStreamSynchronize(rb3b_stream));
	rb3b_tflush();
	return RB3B_OK;
}

/* ---- prepared batches: BWT + walk order straight from the suffix sort ---- */
/*
 * The reference's merge recovers the order in which it meets the rows of the batch by chasing LF over the batch BWT
 * (rb3_mg_rank1_plain, fm-index.c:160-175); our seam call does the same with a list ranking (rb3b_merge.cu).  When the
 * batch is sorted here anyway, that order is known: walk-order position st + l of a string occupying text positions
 * [st, e] (e = its sentinel) is the suffix starting at e - l, i.e. row ISA[e - l], and its BWT symbol is T[e - l - 1].
 * So the batch in walk order is the text read backwards string by string, with the inverse suffix array alongside.
 */
__global__ void k_zero_flags(int64_t len, const uint8_t *__restrict__ T, int64_t *__restrict__ flag);
__global__ void k_zero_pos(int64_t len, const uint8_t *__restrict__ T, const int64_t *__restrict__ sid, int64_t *__restrict__ Z);

struct TextIsZero { /* predicate of the sentinel-position select */
	const uint8_t *T;
	__host__ __device__ bool operator()(const int64_t &i) const { return T[i] == 0; }
};

/* Z[0..n_seq): the sentinel positions, ascending; the string of position q is the first s with Z[s] >= q */
__global__ void k_text_walk_order(int64_t len, const uint8_t *__restrict__ T, const uint32_t *__restrict__ isa, int64_t n_seq, const int64_t *__restrict__ Z,
                                  uint8_t *__restrict__ wsym, uint32_t *__restrict__ wrow, int64_t *__restrict__ c_base, int64_t *__restrict__ c_len)
{
	const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= len) return;
	int64_t lo = 0, hi = n_seq - 1; /* the last symbol is a sentinel: Z[n_seq - 1] = len - 1 >= q */
	while (lo < hi) {
		const int64_t mid = (lo + hi) >> 1;
		if (__ldg(Z + mid) >= q) hi = mid; else lo = mid + 1;
	}
	const int64_t s = lo, st = s ? __ldg(Z + s - 1) + 1 : 0, e = __ldg(Z + s), src = st + e - q;
	wrow[q] = isa[src];
	wsym[q] = src > 0 ? T[src - 1] : 0; /* the symbol before the first string is the last sentinel */
	if (q == e) { c_base[s] = st; c_len[s] = e - st + 1; }
}

__global__ void k_count_zero(int64_t len, const uint8_t *__restrict__ T, unsigned long long *__restrict__ cnt)
{
	unsigned int c = 0;
	for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16, k = 0; k < 16 && i + k < len; ++k) c += T[i + k] == 0;
	for (int d = 16; d > 0; d >>= 1) c += __shfl_xor_sync(0xffffffffu, c, d);
	if ((threadIdx.x & 31) == 0 && c) atomicAdd(cnt, (unsigned long long)c);
}

__global__ void k_batch_count(int64_t len, const uint8_t *__restrict__ bwt, unsigned long long *__restrict__ cnt)
{
	__shared__ unsigned int sh[8];
	if (threadIdx.x < 8) sh[threadIdx.x] = 0;
	__syncthreads();
	unsigned int c[RB3B_ASIZE] = {0, 0, 0, 0, 0, 0};
	for (int64_t i = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) * 16, k = 0; k < 16 && i + k < len; ++k) {
		const int a = bwt[i + k];
#pragma unroll
		for (int b = 0; b < RB3B_ASIZE; ++b) c[b] += a == b;
	}
#pragma unroll
	for (int b = 0; b < RB3B_ASIZE; ++b) {
		unsigned int v = c[b];
		for (int d = 16; d > 0; d >>= 1) v += __shfl_xor_sync(0xffffffffu, v, d);
		if ((threadIdx.x & 31) == 0 && v) atomicAdd(&sh[b], v);
	}
	__syncthreads();
	if (threadIdx.x < RB3B_ASIZE && sh[threadIdx.x]) atomicAdd(&cnt[threadIdx.x], (unsigned long long)sh[threadIdx.x]);
}

extern "C" void rb3b_batch_destroy(rb3b_batch_t *b)
{
	if (b == 0) return;
	cudaSetDevice(b->device);
	/* one stream-ordered block (the pool keeps it for the next batch: no cudaMalloc / cudaFree per batch) */
	if (b->bwt) cudaFreeAsync(b->bwt, rb3b_stream);
	delete b;
}

extern "C" int64_t rb3b_batch_len(const rb3b_batch_t *b) { return b ? b->len : 0; }
extern "C" const uint8_t *rb3b_batch_bwt_dev(const rb3b_batch_t *b) { return b ? b->bwt : 0; }

static int batch_prepare_i(rb3b_batch_s *B, int64_t len, const uint8_t *d_text)
{
	DBuf<uint32_t> sa, isa;
	DBuf<int64_t> Z;
	DBuf<unsigned long long> cnt;
	TRY(isa.alloc(len));
	rb3b_tic(T_BWT);
	TRY(suffix_sort(len, d_text, sa, RB3B_ASIZE, &isa));
	/* layout of the one block: bwt | wsym | wrow | c_base, c_len (the number of strings is not known yet: sized for the
	 * worst case only when small, else allocated after counting) */
	const bool want_order = len < (1LL << 29) && rb3b_get_param("prepare_walk_order", 1) != 0;
	const size_t o_sym = ((size_t)len + 511) & ~(size_t)511, o_row = o_sym + (((size_t)len + 64 + 511) & ~(size_t)511), o_chain = o_row + ((((size_t)len + 8) * 4 + 511) & ~(size_t)511);
	DBuf<unsigned long long> cnt0;
	TRY(cnt0.alloc(8));
	CK(cudaMemsetAsync(cnt0.p, 0, 64, rb3b_stream));
	k_count_zero<<<nblk((len + 15) / 16, TPB), TPB, 0, rb3b_stream>>>(len, d_text, cnt0.p); CKK();
	unsigned long long n_zero = 0;
	CK(cudaMemcpyAsync(&n_zero, cnt0.p, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	const size_t total = want_order ? o_chain + (size_t)n_zero * 16 + 512 : (size_t)len;
	if (cudaMallocAsync((void**)&B->bwt, total, rb3b_stream) != cudaSuccess) { cudaGetLastError(); B->bwt = 0; return rb3b_fail(RB3B_ENOMEM, "out of device memory for the batch (%zu bytes)", total); }
	k_sa_to_bwt<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>((uint32_t)len, d_text, sa.p, B->bwt); CKK();
	TRY(cnt.alloc(8));
	CK(cudaMemsetAsync(cnt.p, 0, 64, rb3b_stream));
	k_batch_count<<<nblk((len + 15) / 16, TPB), TPB, 0, rb3b_stream>>>(len, B->bwt, cnt.p); CKK();
	unsigned long long hc[8];
	CK(cudaMemcpyAsync(hc, cnt.p, 64, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	B->acc[0] = 0;
	for (int a = 0; a < RB3B_ASIZE; ++a) B->acc[a + 1] = B->acc[a] + (int64_t)hc[a];
	B->n_seq = (int64_t)hc[0];
	if (!want_order) { rb3b_toc(T_BWT); CK(cudaStreamSynchronize(rb3b_stream)); return RB3B_OK; } /* 64-bit rows: the seam call rebuilds the order */
	if ((unsigned long long)B->n_seq != n_zero) return rb3b_fail(RB3B_EINVAL, "internal error: %lld sentinels in the BWT, %llu in the text", (long long)B->n_seq, n_zero);
	B->wsym = B->bwt + o_sym; B->wrow = (uint32_t*)(B->bwt + o_row); B->c_base = (int64_t*)(B->bwt + o_chain);
	B->c_len = B->c_base + B->n_seq;
	TRY(Z.alloc(B->n_seq));
	{ /* the sentinel positions: one select over the text (no per-position string ids, no scan) */
		cub::CountingInputIterator<int64_t> pos(0);
		TextIsZero is_zero; is_zero.T = d_text;
		size_t tb = 0;
		CK(cub::DeviceSelect::If((void*)0, tb, pos, Z.p, cnt0.p + 1, len, is_zero, rb3b_stream));
		DBuf<uint8_t> t;
		TRY(t.alloc(tb));
		CK(cub::DeviceSelect::If((void*)t.p, tb, pos, Z.p, cnt0.p + 1, len, is_zero, rb3b_stream));
	}
	CK(cudaMemsetAsync(B->wsym + len, 0, 64, rb3b_stream));
	CK(cudaMemsetAsync(B->wrow + len, 0, 32, rb3b_stream));
	k_text_walk_order<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, d_text, isa.p, B->n_seq, Z.p, B->wsym, B->wrow, B->c_base, B->c_len); CKK();
	rb3b_toc(T_BWT);
	CK(cudaStreamSynchronize(rb3b_stream));
	rb3b_tflush();
	return RB3B_OK;
}

/* rb3_build_sais (sais-ss.c:50-56) + the batch-only half of rb3_mg_rank_plain: text in device memory -> prepared batch */
extern "C" rb3b_batch_t *rb3b_batch_prepare_dev(int64_t len, const uint8_t *d_text)
{
	ApiScope scope_;
	if (rb3b_ensure_init() != RB3B_OK) return 0;
	if (len <= 0) { rb3b_fail(RB3B_EINVAL, "empty batch"); return 0; }
	rb3b_batch_s *B = new rb3b_batch_s;
	memset(B, 0, sizeof(*B));
	B->len = len; B->device = rb3b_cur()->device;
	if (batch_prepare_i(B, len, d_text) != RB3B_OK) { rb3b_batch_destroy(B); return 0; }
	return B;
}

extern "C" rb3b_batch_t *rb3b_batch_prepare(int64_t len, const uint8_t *text)
{
	ApiScope scope_;
	if (rb3b_ensure_init() != RB3B_OK) return 0;
	if (len <= 0) { rb3b_fail(RB3B_EINVAL, "empty batch"); return 0; }
	DBuf<uint8_t> t;
	if ((t.p = rb3b_prefetched(text, len)) == 0) { /* not copied ahead (rb3b_prefetch_batch) */
		if (t.alloc(len) != RB3B_OK) return 0;
		if (cudaMemcpyAsync(t.p, text, (size_t)len, cudaMemcpyHostToDevice, rb3b_stream) != cudaSuccess) { rb3b_fail(RB3B_ENODEV, "host to device copy failed"); return 0; }
	}
	return rb3b_batch_prepare_dev(len, t.p);
}

extern "C" int rb3b_build_bwt(int64_t len, const uint8_t *text, uint8_t *bwt_out)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (len <= 0) return rb3b_fail(RB3B_EINVAL, "empty batch");
	DBuf<uint8_t> t, b;
	TRY(t.alloc(len)); TRY(b.alloc(len));
	CK(cudaMemcpyAsync(t.p, text, len, cudaMemcpyHostToDevice, rb3b_stream));
	TRY(rb3b_build_bwt_dev(len, t.p, b.p));
	CK(cudaMemcpyAsync(bwt_out, b.p, len, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}

/* ---- BWT of a batch in RLO / RCLO order (build -s / -r; mr_insert_multi, mrope.c:300-385) ---- */
/*
 * The reference inserts the reads of a batch column by column (BCR); strings that share the suffix inserted so far
 * are kept as one SA interval and ordered by the symbol inserted next -- '$' first, then A,C,G,T (RLO) or T,G,C,A
 * (RCLO), then N (mrope.c:244-265).  The resulting order is canonical: suffix t of string S sorts by
 *     S[t..) '$'  followed by  S[t-1], S[t-2], ..., S[0], <start of string>
 * with the second half compared under the RLO/RCLO symbol order and the start of the string smallest.  Equal suffixes
 * have equal length, so this is the plain suffix order of the augmented string  S '$' rev(S) 0  (symbols shifted up
 * by one, 0 = terminator): one device suffix sort of a text twice as long replaces the max-read-length BCR rounds.
 */
__global__ void k_zero_flags(int64_t len, const uint8_t *__restrict__ T, int64_t *__restrict__ flag)
{
	int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q < len) flag[q] = T[q] == 0;
}

__global__ void k_zero_pos(int64_t len, const uint8_t *__restrict__ T, const int64_t *__restrict__ sid, int64_t *__restrict__ Z)
{
	int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q < len && T[q] == 0) Z[sid[q]] = q;
}

/* U = augmented text (2 * len symbols), emit[p] = BWT symbol of the suffix of U starting at p when that suffix stands
 * for a suffix of the batch (first half of an augmented string), 255 otherwise */
__global__ void k_augment(int64_t len, const uint8_t *__restrict__ T, const int64_t *__restrict__ sid, const int64_t *__restrict__ Z, int so,
                          uint8_t *__restrict__ U, uint8_t *__restrict__ emit, int *__restrict__ bad)
{
	int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (q >= len) return;
	const int64_t i = sid[q], s = i ? Z[i - 1] + 1 : 0, e = Z[i], L = e - s, base = 2 * s, o = q - s;
	if (q < e) {
		int c = T[q];
		if (c >= RB3B_ASIZE) { *bad = 1; c = 5; }
		const int pc = (so == 2 && c >= 1 && c <= 4) ? 5 - c : c; /* rope_comp6, mrope.c:224 */
		U[base + o] = (uint8_t)(c + 1); emit[base + o] = o == 0 ? 0 : T[q - 1];
		const int64_t r = base + L + 1 + (e - 1 - q);
		U[r] = (uint8_t)(pc + 1); emit[r] = 255;
	} else {
		U[base + L] = 1; emit[base + L] = L > 0 ? T[e - 1] : 0;
		U[base + 2 * L + 1] = 0; emit[base + 2 * L + 1] = 255;
	}
}

__global__ void k_emit_flags(int64_t n, const uint32_t *__restrict__ sa, const uint8_t *__restrict__ emit, int64_t *__restrict__ flag)
{
	int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j < n) flag[j] = emit[sa[j]] != 255;
}

__global__ void k_emit_bwt(int64_t n, const uint32_t *__restrict__ sa, const uint8_t *__restrict__ emit, const int64_t *__restrict__ dst, uint8_t *__restrict__ bwt)
{
	int64_t j = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (j >= n) return;
	const uint8_t c = emit[sa[j]];
	if (c != 255) bwt[dst[j]] = c;
}

extern "C" int rb3b_build_bwt_so_dev(int64_t len, const uint8_t *d_text, int so, uint8_t *d_bwt_out)
{ /* declared in include/rb3_b200.h */
	if (so == 0) return rb3b_build_bwt_dev(len, d_text, d_bwt_out);
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (so < 0 || so > 2) return rb3b_fail(RB3B_EINVAL, "sorting order must be 0, 1 (RLO) or 2 (RCLO)");
	if (len <= 0) return rb3b_fail(RB3B_EINVAL, "empty batch");
	uint8_t last = 1;
	CK(cudaMemcpyAsync(&last, d_text + len - 1, 1, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (last != 0) return rb3b_fail(RB3B_EINVAL, "the batch text must end with a sentinel (mrope.c:310 asserts the same)");
	DBuf<int64_t> flag, sid, Z;
	DBuf<uint8_t> U, emit;
	DBuf<uint32_t> sa;
	DBuf<int> bad;
	int hbad = 0;
	TRY(flag.alloc(2 * len)); TRY(sid.alloc(2 * len)); TRY(U.alloc(2 * len)); TRY(emit.alloc(2 * len)); TRY(bad.alloc(1));
	CK(cudaMemsetAsync(bad.p, 0, sizeof(int), rb3b_stream));
	rb3b_tic(T_BWT);
	k_zero_flags<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, d_text, flag.p); CKK();
	TRY(rb3b_scan_excl_i64(flag.p, sid.p, len));
	int64_t n_seq = 0;
	CK(cudaMemcpyAsync(&n_seq, sid.p + len - 1, 8, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	++n_seq; /* the last symbol is a sentinel */
	TRY(Z.alloc(n_seq));
	k_zero_pos<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, d_text, sid.p, Z.p); CKK();
	k_augment<<<nblk(len, TPB), TPB, 0, rb3b_stream>>>(len, d_text, sid.p, Z.p, so, U.p, emit.p, bad.p); CKK();
	CK(cudaMemcpyAsync(&hbad, bad.p, sizeof(int), cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	if (hbad) return rb3b_fail(RB3B_EINVAL, "batch text holds a symbol >= %d", RB3B_ASIZE);
	TRY(suffix_sort(2 * len, U.p, sa, RB3B_ASIZE + 1));
	k_emit_flags<<<nblk(2 * len, TPB), TPB, 0, rb3b_stream>>>(2 * len, sa.p, emit.p, flag.p); CKK();
	TRY(rb3b_scan_excl_i64(flag.p, sid.p, 2 * len));
	DBuf<uint8_t> tmp;
	uint8_t *dst = d_bwt_out;
	if (d_bwt_out == d_text) { TRY(tmp.alloc(len)); dst = tmp.p; }
	k_emit_bwt<<<nblk(2 * len, TPB), TPB, 0, rb3b_stream>>>(2 * len, sa.p, emit.p, sid.p, dst); CKK();
	if (dst != d_bwt_out) CK(cudaMemcpyAsync(d_bwt_out, dst, len, cudaMemcpyDeviceToDevice, rb3b_stream));
	rb3b_toc(T_BWT);
	CK(cudaStreamSynchronize(rb3b_stream));
	rb3b_tflush();
	return RB3B_OK;
}

extern "C" int rb3b_build_bwt_so(int64_t len, const uint8_t *text, int so, uint8_t *bwt_out)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (len <= 0) return rb3b_fail(RB3B_EINVAL, "empty batch");
	DBuf<uint8_t> t, b;
	TRY(t.alloc(len)); TRY(b.alloc(len));
	CK(cudaMemcpyAsync(t.p, text, len, cudaMemcpyHostToDevice, rb3b_stream));
	TRY(rb3b_build_bwt_so_dev(len, t.p, so, b.p));
	CK(cudaMemcpyAsync(bwt_out, b.p, len, cudaMemcpyDeviceToHost, rb3b_stream));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}

/* ---- BWT-vs-BWT merge ---- */

__global__ void k_runs_to_plain(int64_t n_runs, int64_t n, const uint8_t *__restrict__ sym, const int64_t *__restrict__ start, uint8_t *__restrict__ out)
{
	int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
	if (i >= n) return;
	int64_t lo = 0, hi = n_runs; /* last run with start <= i */
	while (hi - lo > 1) {
		int64_t mid = (lo + hi) >> 1;
		if (start[mid] <= i) lo = mid; else hi = mid;
	}
	out[i] = sym[lo];
}

/* the BWT of an index as one byte per symbol (arena memory) */
int rb3b_index_to_plain_dev(const rb3b_index_s *x, DBuf<uint8_t> &plain)
{
	DBuf<uint8_t> sym;
	DBuf<int64_t> len, start;
	int64_t n_runs;
	TRY(rb3b_export_runs_dev(x, sym, len, &n_runs));
	TRY(start.alloc(n_runs)); TRY(plain.alloc(x->n));
	TRY(rb3b_scan_excl_i64(len.p, start.p, n_runs));
	k_runs_to_plain<<<nblk(x->n, TPB), TPB, 0, rb3b_stream>>>(n_runs, x->n, sym.p, start.p, plain.p); CKK();
	return RB3B_OK;
}

extern "C" int rb3b_merge_index(rb3b_index_t *x, const rb3b_index_t *other)
{
	ApiScope scope_;
	TRY(rb3b_ensure_init());
	if (other->n == 0) return RB3B_OK;
	DBuf<uint8_t> plain;
	TRY(rb3b_index_to_plain_dev(other, plain));
	TRY(rb3b_merge_plain_dev(x, other->n, plain.p));
	CK(cudaStreamSynchronize(rb3b_stream));
	return RB3B_OK;
}
