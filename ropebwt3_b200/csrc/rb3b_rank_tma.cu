/*
 * rb3b_rank_tma.cu -- LF/rank kernel, bulk-copy (TMA engine) variant.
 *
 * Same arithmetic as k_lf (rb3b_index.cu) but the 128-B index cell of every
 * query is fetched with cp.async.bulk (SASS UBLKCP) into a per-group ring of
 * shared-memory stages and its arrival is tracked with an mbarrier, so each
 * 8-lane group keeps NST block fetches in flight instead of one and no
 * registers are tied up while the data travels.
 */
#include "rb3b_internal.cuh"

#define TPB 256
#define NGRP (TPB / RB3B_GROUP)
#define NST 4

__device__ __forceinline__ uint32_t smem_u32(const void *p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count)
{ asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" :: "r"(bar), "r"(count) : "memory"); }

__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes)
{ asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" :: "r"(bar), "r"(bytes) : "memory"); }

__device__ __forceinline__ void bulk_g2s(uint32_t dst, const void *src, uint32_t bytes, uint32_t bar)
{
	asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
	             :: "r"(dst), "l"(src), "r"(bytes), "r"(bar) : "memory");
}

__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity)
{
	uint32_t done;
	do {
		asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
		             : "=r"(done) : "r"(bar), "r"(parity) : "memory");
	} while (!done);
}

__global__ void __launch_bounds__(TPB) k_lf_tma(DevIndex x, int64_t nq, const int64_t *__restrict__ k_, const uint8_t *__restrict__ c_, int64_t *__restrict__ out)
{
	__shared__ __align__(128) uint4 stage[NGRP][NST][8];
	__shared__ __align__(8) uint64_t bars[NGRP][NST];
	typedef Grp<8> G8;
	const int gl = G8::lane(), gi = threadIdx.x >> 3;
	const unsigned gmask = G8::mask();
	const int64_t g = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 3, ng = ((int64_t)gridDim.x * blockDim.x) >> 3;
	if (gl == 0) {
#pragma unroll
		for (int s = 0; s < NST; ++s) mbar_init(smem_u32(&bars[gi][s]), 1);
		asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
	}
	__syncthreads();
	int64_t kq[NST];
	int cq[NST], kind[NST]; /* 0: nothing, 1: block in flight, 2: k >= n */
	uint32_t parity = 0;   /* bit s: phase of stage s */
	/* one stage = one query of this group */
#define ISSUE(s, q) do { \
		kind[s] = 0; \
		if ((q) < nq) { \
			kq[s] = k_[q]; cq[s] = c_[q]; if (kq[s] < 0) kq[s] = 0; \
			if (kq[s] >= x.n) kind[s] = 2; \
			else { \
				if (gl == 0) { \
					uint32_t bar_ = smem_u32(&bars[gi][s]); \
					mbar_expect_tx(bar_, 128); \
					bulk_g2s(smem_u32(&stage[gi][s][0]), x.cells + (kq[s] >> x.shift) * 8, 128, bar_); \
				} \
				kind[s] = 1; \
			} \
		} \
	} while (0)
#pragma unroll
	for (int s = 0; s < NST; ++s) ISSUE(s, g + (int64_t)s * ng);
	for (int64_t q0 = g; q0 < nq; q0 += (int64_t)NST * ng) {
#pragma unroll
		for (int s = 0; s < NST; ++s) {
			int64_t q = q0 + (int64_t)s * ng;
			if (kind[s] == 2) { if (gl == 0) out[q] = x.acc[cq[s]] + x.tot[cq[s]]; }
			else if (kind[s] == 1) {
				mbar_wait(smem_u32(&bars[gi][s]), (parity >> s) & 1);
				parity ^= 1u << s;
				uint4 v[1];
				v[0] = stage[gi][s][gl];
				const int c = cq[s];
				int64_t r = G8::count(x, v, kq[s], c);
				if (gl == 0) out[q] = x.acc[c] + r;
			}
			__syncwarp(gmask); /* every lane has read the stage before it is refilled */
			ISSUE(s, q + (int64_t)NST * ng);
		}
	}
#undef ISSUE
}

int rb3b_lf_tma_launch(const rb3b_index_s *x, int64_t nq, const int64_t *d_k, const uint8_t *d_c, int64_t *d_out)
{
	int dev = 0, sm = 148;
	cudaGetDevice(&dev);
	cudaDeviceGetAttribute(&sm, cudaDevAttrMultiProcessorCount, dev);
	int64_t want = (nq * RB3B_GROUP + TPB - 1) / TPB, cap = (int64_t)sm * 8;
	k_lf_tma<<<(unsigned)(want < cap ? want : cap), TPB, 0, rb3b_stream>>>(rb3b_dev_view(x), nq, d_k, d_c, d_out);
	CKK();
	return RB3B_OK;
}
