// pgx_comm.cuh -- collectives for the sharded path (one process per GPU, genomes partitioned across ranks).
// NCCL is bound at run time with dlopen so that a single-GPU build/run has no NCCL dependency and a torchrun
// process reuses the libnccl.so.2 that torch already loaded.  All data-path reductions are integer sums/maxima,
// so the sharded result is bit-identical to the single-GPU one for any genome->rank assignment.
//
// Second transport, for TESTING the sharded path where the ranks cannot have a GPU each (the CI box has one): "loopback" -- the ranks are
// processes that may share one device; a collective drains the stream, copies the operand to a POSIX shared-memory segment, meets the other
// ranks at a barrier in that segment, reduces / gathers on the host and copies the result back.  Same call sites, same integer arithmetic,
// so the partition / merge / fallback logic of a multi-rank run is exercised end to end; it is slow by construction and never used by bench.py.
#pragma once
#include <dlfcn.h>
#include <fcntl.h>
#include <sched.h>
#include <sys/mman.h>
#include <unistd.h>
#include <atomic>
#include <cstring>
#include <vector>
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

	// ---- loopback transport (see the header comment) --------------------------------------------------------------------------
	struct ShmHeader { std::atomic<uint32_t> arrived, generation; uint32_t nranks, pad; };
	char *shm = nullptr; size_t shm_bytes = 0, shm_slot = 0;
	bool loopback() const { return shm != nullptr; }
	void shm_open_segment(const char *name, int rank_, int nranks_, size_t slot_bytes)
	{
		shm_slot = slot_bytes; shm_bytes = 4096 + slot_bytes * (size_t)nranks_;
		int fd = -1;
		if (rank_ == 0) {
			::shm_unlink(name);
			fd = ::shm_open(name, O_CREAT | O_EXCL | O_RDWR, 0600);
			if (fd < 0 || ftruncate(fd, (off_t)shm_bytes) != 0) PGX_FATAL("loopback transport: cannot create shared memory segment %s", name);
		} else {
			for (int tries = 0; tries < 20000 && fd < 0; ++tries) { fd = ::shm_open(name, O_RDWR, 0600); if (fd < 0) usleep(1000); }
			if (fd < 0) PGX_FATAL("loopback transport: cannot open shared memory segment %s", name);
			for (int tries = 0; tries < 20000; ++tries) { off_t sz = lseek(fd, 0, SEEK_END); if ((size_t)sz >= shm_bytes) break; usleep(1000); }
		}
		shm = (char*)mmap(nullptr, shm_bytes, PROT_READ | PROT_WRITE, MAP_SHARED, fd, 0);
		::close(fd);
		if (shm == MAP_FAILED) { shm = nullptr; PGX_FATAL("loopback transport: mmap failed"); }
		ShmHeader *h = (ShmHeader*)shm;
		if (rank_ == 0) { h->arrived.store(0); h->generation.store(0); h->nranks = (uint32_t)nranks_; std::atomic_thread_fence(std::memory_order_seq_cst); h->pad = 0x50475821u; }
		else while (((volatile ShmHeader*)h)->pad != 0x50475821u) usleep(100);
		shm_barrier();
		if (rank_ == 0) ::shm_unlink(name);      // everybody has it mapped
	}
	void shm_close() { if (shm) munmap(shm, shm_bytes); shm = nullptr; shm_bytes = 0; }
	void shm_barrier()
	{
		ShmHeader *h = (ShmHeader*)shm;
		const uint32_t gen = h->generation.load();
		if (h->arrived.fetch_add(1) + 1 == (uint32_t)nranks) { h->arrived.store(0); h->generation.fetch_add(1); }
		else while (h->generation.load() == gen) sched_yield();
	}
	char *slot(int r) const { return shm + 4096 + shm_slot * (size_t)r; }
	static size_t type_size(int type) { return type == T_INT8 || type == T_UINT8? 1 : (type == T_INT32 || type == T_UINT32? 4 : 8); }
	template<class T> static void host_reduce(T *acc, const T *x, size_t n, int op) { for (size_t i = 0; i < n; ++i) acc[i] = op == OP_SUM? (T)(acc[i] + x[i]) : (x[i] > acc[i]? x[i] : acc[i]); }
	void shm_allreduce(void *buf, size_t count, int type, int op, cudaStream_t s)
	{
		const size_t bytes = count * type_size(type);
		if (bytes > shm_slot) PGX_FATAL("loopback transport: %zu-byte operand exceeds the %zu-byte slot", bytes, shm_slot);
		PGX_CHECK(cudaMemcpyAsync(slot(rank), buf, bytes, cudaMemcpyDeviceToHost, s));
		PGX_CHECK(cudaStreamSynchronize(s));
		shm_barrier();
		std::vector<char> acc(slot(0), slot(0) + bytes);
		for (int r = 1; r < nranks; ++r) {
			if (type == T_INT32) host_reduce((int32_t*)acc.data(), (const int32_t*)slot(r), count, op);
			else if (type == T_UINT32) host_reduce((uint32_t*)acc.data(), (const uint32_t*)slot(r), count, op);
			else if (type == T_UINT8) host_reduce((uint8_t*)acc.data(), (const uint8_t*)slot(r), count, op);
			else if (type == T_UINT64) host_reduce((uint64_t*)acc.data(), (const uint64_t*)slot(r), count, op);
			else if (type == T_INT64) host_reduce((int64_t*)acc.data(), (const int64_t*)slot(r), count, op);
			else PGX_FATAL("loopback transport: unsupported element type %d", type);
		}
		shm_barrier();                                   // everybody has read every slot
		PGX_CHECK(cudaMemcpyAsync(buf, acc.data(), bytes, cudaMemcpyHostToDevice, s));
		PGX_CHECK(cudaStreamSynchronize(s));
	}
	void shm_allgather(const void *in, void *out, size_t bytes, cudaStream_t s)
	{
		if (bytes > shm_slot) PGX_FATAL("loopback transport: %zu-byte operand exceeds the %zu-byte slot", bytes, shm_slot);
		PGX_CHECK(cudaMemcpyAsync(slot(rank), in, bytes, cudaMemcpyDeviceToHost, s));
		PGX_CHECK(cudaStreamSynchronize(s));
		shm_barrier();
		for (int r = 0; r < nranks; ++r) PGX_CHECK(cudaMemcpyAsync((char*)out + bytes * (size_t)r, slot(r), bytes, cudaMemcpyHostToDevice, s));
		PGX_CHECK(cudaStreamSynchronize(s));
		shm_barrier();
	}

	// several collectives as one launch (ncclGroupStart/End)
	void group_begin() { if (active() && !loopback()) check(GroupStart(), "group start"); }
	void group_end() { if (active() && !loopback()) check(GroupEnd(), "group end"); }
	void allreduce(void *buf, size_t count, int type, int op, cudaStream_t s)
	{
		if (!active() || !count) return;
		if (loopback()) shm_allreduce(buf, count, type, op, s); else check(AllReduce(buf, buf, count, type, op, comm, s), "allreduce");
	}
	// every rank contributes `bytes` bytes; out holds nranks * bytes
	void allgather(const void *in, void *out, size_t bytes, cudaStream_t s)
	{
		if (!active()) { if (in != out && bytes) PGX_CHECK(cudaMemcpyAsync(out, in, bytes, cudaMemcpyDeviceToDevice, s)); return; }
		if (!bytes) return;
		if (loopback()) shm_allgather(in, out, bytes, s); else check(AllGather(in, out, bytes, T_INT8, comm, s), "allgather");
	}
};

extern Comm g_comm;

} // namespace pgx
