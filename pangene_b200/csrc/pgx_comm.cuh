// pgx_comm.cuh -- collectives for the sharded path (one process per GPU, genomes partitioned across ranks).
// NCCL is bound at run time with dlopen so that a single-GPU build/run has no NCCL dependency and a torchrun
// process reuses the libnccl.so.2 that torch already loaded.  All data-path reductions are integer sums/maxima,
// so the sharded result is bit-identical to the single-GPU one for any genome->rank assignment.
#pragma once
#include <dlfcn.h>
#include "pgx_util.cuh"

namespace pgx {

struct Comm {
	int rank = 0, nranks = 1;
	int32_t n_genome_total = -1, genome_begin = 0;
	void *lib = nullptr, *comm = nullptr;
	// minimal NCCL surface (nccl.h): result 0 == ncclSuccess
	typedef struct { char internal[128]; } UniqueId;
	int (*GetUniqueId)(UniqueId*) = nullptr;
	int (*CommInitRank)(void**, int, UniqueId, int) = nullptr;
	int (*CommDestroy)(void*) = nullptr;
	int (*AllReduce)(const void*, void*, size_t, int, int, void*, cudaStream_t) = nullptr;
	int (*AllGather)(const void*, void*, size_t, int, void*, cudaStream_t) = nullptr;
	int (*GroupStart)() = nullptr;
	int (*GroupEnd)() = nullptr;
	const char *(*GetErrorString)(int) = nullptr;
	enum { T_INT8 = 0, T_UINT8 = 1, T_INT32 = 2, T_UINT32 = 3, T_INT64 = 4, T_UINT64 = 5 };   // ncclDataType_t
	enum { OP_SUM = 0, OP_MAX = 2 };                                              // ncclRedOp_t

	bool load()
	{
		if (lib) return true;
		const char *names[] = { "libnccl.so.2", "libnccl.so", nullptr };
		for (int i = 0; names[i] && !lib; ++i) lib = dlopen(names[i], RTLD_NOW | RTLD_GLOBAL);
		if (!lib) return false;
		GetUniqueId = (decltype(GetUniqueId))dlsym(lib, "ncclGetUniqueId");
		CommInitRank = (decltype(CommInitRank))dlsym(lib, "ncclCommInitRank");
		CommDestroy = (decltype(CommDestroy))dlsym(lib, "ncclCommDestroy");
		AllReduce = (decltype(AllReduce))dlsym(lib, "ncclAllReduce");
		AllGather = (decltype(AllGather))dlsym(lib, "ncclAllGather");
		GetErrorString = (decltype(GetErrorString))dlsym(lib, "ncclGetErrorString");
		GroupStart = (decltype(GroupStart))dlsym(lib, "ncclGroupStart");
		GroupEnd = (decltype(GroupEnd))dlsym(lib, "ncclGroupEnd");
		return GetUniqueId && CommInitRank && CommDestroy && AllReduce && AllGather && GroupStart && GroupEnd;
	}
	void check(int r, const char *what) { if (r != 0) PGX_FATAL("NCCL %s failed: %s", what, GetErrorString? GetErrorString(r) : "?"); }
	bool active() const { return nranks > 1; }
	// several collectives as one launch (ncclGroupStart/End)
	void group_begin() { if (active()) check(GroupStart(), "group start"); }
	void group_end() { if (active()) check(GroupEnd(), "group end"); }
	void allreduce(void *buf, size_t count, int type, int op, cudaStream_t s) { if (active() && count) check(AllReduce(buf, buf, count, type, op, comm, s), "allreduce"); }
	// every rank contributes `bytes` bytes; out holds nranks * bytes
	void allgather(const void *in, void *out, size_t bytes, cudaStream_t s)
	{
		if (!active()) { if (in != out && bytes) PGX_CHECK(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s)); return; }
		if (bytes) check(AllGather(in, out, bytes, T_INT8, comm, s), "allgather");
	}
};

extern Comm g_comm;

} // namespace pgx
