"""Launcher-side plumbing for the sharded path: one process per GPU, genomes partitioned across ranks.

libpgx does its own NCCL collectives (include/pgx.h: pgx_comm_*); what is left for the launcher is (1) deciding
which genomes a rank owns, (2) carrying the 128-byte NCCL id from rank 0 to everybody over whatever process group
the launcher already has (torch.distributed with nccl or gloo), and (3) putting the per-rank W lines back into
genome order.  The graph (S and L lines) comes back replicated on every rank.
"""
import ctypes as C


def partition(n_genome, world):
    """contiguous, near-equal genome ranges; rank r owns [begin, end)"""
    return [(n_genome * r // world, n_genome * (r + 1) // world) for r in range(world)]


def broadcast_unique_id(make_id, rank, group=None):
    """rank 0 calls make_id() -> 128 bytes; everybody returns the same 128 bytes"""
    import torch
    import torch.distributed as dist
    dev = "cuda" if dist.get_backend(group) == "nccl" else "cpu"
    t = torch.zeros(128, dtype=torch.uint8, device=dev)
    if rank == 0:
        raw = make_id()
        assert len(raw) == 128
        t = torch.frombuffer(bytearray(raw), dtype=torch.uint8).clone().to(dev)
    dist.broadcast(t, 0, group=group)
    return bytes(t.cpu().numpy().tobytes())


def init_comm(lib, rank, world, n_genome_total, genome_begin, group=None):
    """collective: NCCL communicator inside libpgx for this process group"""
    def make_id():
        buf = (C.c_char * 128)()
        lib.pgx_comm_unique_id(buf)
        return buf.raw
    raw = broadcast_unique_id(make_id, rank, group)
    lib.pgx_comm_init(rank, world, raw, n_genome_total, genome_begin)


def init_comm_loopback(lib, rank, world, n_genome_total, genome_begin, tag="0"):
    """collective: the loopback transport of libpgx (collectives staged through shared memory, ranks may share one GPU) -- for testing the
    sharded path on a single-GPU box; the segment name must be the same on every rank and unique per run"""
    import os
    name = "/pgx_loop_%s_%s" % (os.environ.get("MASTER_PORT", "0"), tag)
    lib.pgx_comm_init_loopback(rank, world, name.encode(), n_genome_total, genome_begin)


def gather_in_rank_order(obj, world, group=None):
    """list of every rank's object, rank order == genome order for contiguous partitions"""
    import torch.distributed as dist
    out = [None] * world
    dist.all_gather_object(out, obj, group=group)
    return out
