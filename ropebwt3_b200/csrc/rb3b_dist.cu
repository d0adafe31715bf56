/*
 * rb3b_dist.cu -- multi-device plumbing: one execution context = one rank of an NCCL communicator.
 *
 * NCCL is bound at run time (dlopen of libnccl.so.2) so that the library loads on a single-GPU box without it and,
 * inside a process that already carries an NCCL (PyTorch's), uses that very copy instead of a second one.  Only the
 * types of <nccl.h> are used at compile time.
 *
 * The reference has no multi-device path; what is parallelised here is what it parallelises over host threads: the
 * rank phase over the sequences of the batch (kt_for in rb3_mg_rank_plain, fm-index.c:220).  See rb3b_merge.cu for
 * the sharded rank phase itself (rb3b_merge_plain_dist_dev).
 */
#include <dlfcn.h>
#include <string.h>
#include <mutex>
#include <nccl.h>
#include "rb3b_internal.cuh"

namespace {

struct Nccl {
	void *h;
	ncclResult_t (*GetUniqueId)(ncclUniqueId*);
	ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int);
	ncclResult_t (*CommDestroy)(ncclComm_t);
	ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t);
	ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t);
	ncclResult_t (*GroupStart)(void);
	ncclResult_t (*GroupEnd)(void);
	const char *(*GetErrorString)(ncclResult_t);
	ncclResult_t (*GetVersion)(int*);
	ncclResult_t (*CommSplit)(ncclComm_t, int, int, ncclComm_t*, void*); /* optional (NCCL >= 2.18) */
};

Nccl g_nccl;
std::mutex g_nccl_mu;

int nccl_load(void)
{
	std::lock_guard<std::mutex> lk(g_nccl_mu);
	if (g_nccl.h) return RB3B_OK;
	void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
	if (h == 0) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
	if (h == 0) return rb3b_fail(RB3B_ENODEV, "multi-device calls need NCCL: %s", dlerror());
#define SYM(name) do { *(void**)&g_nccl.name = dlsym(h, "nccl" #name); if (g_nccl.name == 0) return rb3b_fail(RB3B_ENODEV, "libnccl lacks nccl" #name); } while (0)
	SYM(GetUniqueId); SYM(CommInitRank); SYM(CommDestroy); SYM(AllGather); SYM(AllReduce); SYM(Send); SYM(Recv);
	SYM(GroupStart); SYM(GroupEnd); SYM(GetErrorString); SYM(GetVersion);
#undef SYM
	*(void**)&g_nccl.CommSplit = dlsym(h, "ncclCommSplit");
	g_nccl.h = h;
	return RB3B_OK;
}

} /* namespace */

#define NCK(call) do { ncclResult_t r_ = (call); if (r_ != ncclSuccess) \
	return rb3b_fail(RB3B_ENODEV, "%s:%d: %s: %s", __FILE__, __LINE__, #call, g_nccl.GetErrorString(r_)); } while (0)

void rb3b_dist_release(rb3b_ctx_s *c)
{
	if (c->comm2 && g_nccl.h) { cudaSetDevice(c->device); cudaStreamSynchronize(c->stream2); g_nccl.CommDestroy((ncclComm_t)c->comm2); }
	if (c->comm && g_nccl.h) { cudaSetDevice(c->device); g_nccl.CommDestroy((ncclComm_t)c->comm); }
	c->comm = 0; c->comm2 = 0; c->rank = 0; c->world = 1;
}

extern "C" int rb3b_dist_unique_id(void *id128)
{
	TRY(nccl_load());
	ncclUniqueId id;
	NCK(g_nccl.GetUniqueId(&id));
	memcpy(id128, &id, sizeof(id));
	return RB3B_OK;
}

extern "C" int rb3b_dist_init(int rank, int world, const void *id128)
{
	TRY(rb3b_ensure_init());
	TRY(nccl_load());
	rb3b_ctx_s *c = rb3b_cur();
	if (world < 1 || rank < 0 || rank >= world) return rb3b_fail(RB3B_EINVAL, "rank %d of %d", rank, world);
	if (c->comm) rb3b_dist_release(c);
	if (world == 1) { c->rank = 0; c->world = 1; return RB3B_OK; }
	ncclUniqueId id;
	memcpy(&id, id128, sizeof(id));
	ncclComm_t comm;
	NCK(g_nccl.CommInitRank(&comm, world, id, rank));
	c->comm = comm; c->rank = rank; c->world = world;
	return RB3B_OK;
}

/* a second communicator for the context's second stream ("dist_async"): the exchange of the interleave positions and the
 * merge of batch i run there while the first stream already prepares batch i + 1, whose shared first walk all-gathers on
 * the first communicator; collectives of one communicator must not overlap, those of two may.  Collective: every rank
 * calls it at the same point (the first asynchronous multi-device merge).  Failure just leaves the synchronous path. */
int rb3b_dist_second_comm(void)
{
	rb3b_ctx_s *c = rb3b_cur();
	if (c->comm2 || c->comm == 0 || g_nccl.CommSplit == 0) return RB3B_OK;
	ncclComm_t comm2 = 0;
	if (g_nccl.CommSplit((ncclComm_t)c->comm, 0, c->rank, &comm2, 0) == ncclSuccess) c->comm2 = comm2;
	return RB3B_OK;
}

extern "C" int rb3b_dist_finalize(void) { rb3b_dist_release(rb3b_cur()); return RB3B_OK; }
extern "C" int rb3b_dist_rank(void) { return rb3b_cur()->rank; }
extern "C" int rb3b_dist_world(void) { return rb3b_cur()->world; }
extern "C" int rb3b_dist_nccl_version(void) { int v = 0; if (nccl_load() != RB3B_OK) return -1; g_nccl.GetVersion(&v); return v; }

/* ---- collectives on the current context's stream (used by rb3b_merge.cu) ---- */

/* work queued on the second stream goes through the second communicator */
static inline ncclComm_t cur_comm(rb3b_ctx_s *c) { return (ncclComm_t)(c->stream == c->stream2 && c->comm2 ? c->comm2 : c->comm); }

int rb3b_all_gather(const void *send, void *recv, size_t bytes_per_rank)
{
	rb3b_ctx_s *c = rb3b_cur();
	if (c->world == 1) { if (send != recv) CK(cudaMemcpyAsync(recv, send, bytes_per_rank, cudaMemcpyDeviceToDevice, c->stream)); return RB3B_OK; }
	NCK(g_nccl.AllGather(send, recv, bytes_per_rank, ncclUint8, (ncclComm_t)c->comm, c->stream));
	return RB3B_OK;
}

int rb3b_all_reduce_sum_u32(void *buf, size_t n)
{
	rb3b_ctx_s *c = rb3b_cur();
	if (c->world == 1) return RB3B_OK;
	NCK(g_nccl.AllReduce(buf, buf, n, ncclUint32, ncclSum, cur_comm(c), c->stream));
	return RB3B_OK;
}

int rb3b_all_reduce_max_i64(void *buf, size_t n)
{
	rb3b_ctx_s *c = rb3b_cur();
	if (c->world == 1) return RB3B_OK;
	NCK(g_nccl.AllReduce(buf, buf, n, ncclInt64, ncclMax, cur_comm(c), c->stream));
	return RB3B_OK;
}

/* personalised exchange: rank r sends send + soff[p] (scnt[p] bytes) to every peer p and receives rcnt[p] bytes from it
 * at recv + roff[p]; all counts are known on the host */
int rb3b_all_to_all_v(const void *send, const int64_t *soff, const int64_t *scnt, void *recv, const int64_t *roff, const int64_t *rcnt)
{
	rb3b_ctx_s *c = rb3b_cur();
	if (c->world == 1) { if (scnt[0]) CK(cudaMemcpyAsync((char*)recv + roff[0], (const char*)send + soff[0], scnt[0], cudaMemcpyDeviceToDevice, c->stream)); return RB3B_OK; }
	NCK(g_nccl.GroupStart());
	for (int p = 0; p < c->world; ++p) {
		if (scnt[p]) NCK(g_nccl.Send((const char*)send + soff[p], scnt[p], ncclUint8, p, (ncclComm_t)c->comm, c->stream));
		if (rcnt[p]) NCK(g_nccl.Recv((char*)recv + roff[p], rcnt[p], ncclUint8, p, (ncclComm_t)c->comm, c->stream));
	}
	NCK(g_nccl.GroupEnd());
	return RB3B_OK;
}
