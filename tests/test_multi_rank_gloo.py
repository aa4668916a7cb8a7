"""N > 1 host logic on CPU (gloo, world_size 2): the tile partition and the replica-sync protocol.

The pool here is a host-memory stand-in that speaks the same staging format as sync.cu
(8-word header [n_ranges][payload_words][root][flags][colour x 4] | {offset,count,payload_offset} x n | payload): the test checks that
ReplicaSync drives it correctly under a real process group (ONE eager broadcast when the payload fits the eager
chunk, one more for the remainder when it does not, staging growth on both sides, root last)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from vkhashdag_b200 import replica  # noqa: E402


class HostPool:
    """Host-memory pool with the dirty-range contract of the C ABI (append-only buckets)."""

    def __init__(self, n_buckets=16, bucket_shift=6):
        self.shift = bucket_shift
        self.words = np.zeros(n_buckets << bucket_shift, np.uint32)
        self.bw = np.zeros(n_buckets, np.uint32)
        self.synced = np.zeros(n_buckets, np.uint32)
        self.root = 0xFFFFFFFF
        self.applied = 0

    def append(self, bucket, values):
        off = (bucket << self.shift) + int(self.bw[bucket])
        self.words[off:off + len(values)] = values
        self.bw[bucket] += len(values)
        return off

    def _ranges(self):
        return [((b << self.shift) + int(self.synced[b]), int(self.bw[b] - self.synced[b]))
                for b in range(len(self.bw)) if self.bw[b] > self.synced[b]]

    def DirtyCount(self):
        r = self._ranges()
        return len(r), 4 * (8 + 3 * len(r) + sum(c for _, c in r))

    def DirtyPack(self, ptr, capacity):
        r = self._ranges()
        n_words = 8 + 3 * len(r) + sum(c for _, c in r)
        if 4 * n_words > capacity:   # same contract as hd_dirty_pack_dev: report the size, change nothing
            raise replica.StagingTooSmall(4 * n_words)
        buf = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint32 * n_words).from_address(ptr))
        buf[:8] = [len(r), sum(c for _, c in r), self.root, 0, 0, 0, 0xC0000000, 0]
        poff = 0
        payload0 = 8 + 3 * len(r)
        for i, (off, cnt) in enumerate(r):
            buf[8 + 3 * i:11 + 3 * i] = [off, cnt, poff]
            buf[payload0 + poff:payload0 + poff + cnt] = self.words[off:off + cnt]
            poff += cnt
        return 4 * n_words

    def DirtyApply(self, ptr, nbytes):
        buf = np.ctypeslib.as_array((np.ctypeslib.ctypes.c_uint32 * (nbytes // 4)).from_address(ptr))
        n = int(buf[0])
        assert nbytes == replica.packed_bytes(buf)
        payload0 = 8 + 3 * n
        for i in range(n):
            off, cnt, poff = (int(v) for v in buf[8 + 3 * i:11 + 3 * i])
            self.words[off:off + cnt] = buf[payload0 + poff:payload0 + poff + cnt]
            b = off >> self.shift
            self.bw[b] = self.synced[b] = (off & ((1 << self.shift) - 1)) + cnt
        self.root = int(buf[2])   # root last
        self.applied += 1

    def DirtyReset(self):
        self.synced[:] = self.bw


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    pool = HostPool()
    # rounds 0-2: tiny staging + tiny eager chunk -> grow path and the two-collective path;
    # rounds 3-5: roomy eager chunk -> exactly one collective per publish
    sync = replica.ReplicaSync(pool, dist, device="cpu", capacity_bytes=64, eager_bytes=32)
    rng = np.random.default_rng(42)
    for round_ in range(6):
        if round_ == 3:
            assert sync.collectives == 6   # every tiny round needed the remainder broadcast
            sync = replica.ReplicaSync(pool, dist, device="cpu", capacity_bytes=1 << 16, eager_bytes=1 << 12)
        if rank == 0:   # the editing rank appends nodes and moves the root
            for _ in range(5):
                b = int(rng.integers(0, 16))
                n = int(rng.integers(2, 10))
                if pool.bw[b] + n <= 64:
                    pool.root = pool.append(b, rng.integers(1, 2 ** 32, n, dtype=np.uint64).astype(np.uint32))
        nbytes = sync.publish(src=0)
        assert nbytes >= 32
    assert sync.collectives == 3
    # tile partition: each rank owns t % world == rank; together they cover the frame exactly once
    W, H, T = 200, 130, 64
    mine = replica.local_tiles(W, H, T, T, rank, world)
    part = np.zeros(len(mine) * T * T, np.uint32)
    for lt, tx, ty in mine:
        blk = part[lt * T * T:(lt + 1) * T * T].reshape(T, T)
        ys, xs = np.mgrid[0:T, 0:T]
        blk[:] = ((ty * T + ys) << 16) | (tx * T + xs)
    q.put((rank, pool.words.copy(), pool.bw.copy(), pool.root, pool.applied, part))
    dist.barrier()
    dist.destroy_process_group()


def test_replica_sync_and_tile_partition_world2():
    world = 2
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted((q.get(timeout=120) for _ in range(world)), key=lambda t: t[0])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (_, w0, bw0, root0, ap0, part0), (_, w1, bw1, root1, ap1, part1) = res
    assert np.array_equal(w0, w1) and np.array_equal(bw0, bw1) and root0 == root1 != 0xFFFFFFFF
    assert bw0.sum() > 0 and ap0 == 0 and ap1 == 6          # replica applied once per publish, editor never
    W, H, T = 200, 130, 64
    frame = replica.assemble_frame([part0, part1], W, H, T, T, world)
    ys, xs = np.mgrid[0:H, 0:W]
    assert np.array_equal(frame, ((ys << 16) | xs).astype(np.uint32))


def test_tile_partition_matches_c_abi():
    """replica.local_tiles must agree with hd_tile_shard_pixels (host-only C function) for every rank count."""
    import ctypes as C

    import vkhashdag_b200 as v
    from vkhashdag_b200 import abi, api
    L = v.lib()
    P = abi.HdTraceParams()
    for (W, H) in ((3840, 2160), (7680, 4320), (333, 77), (64, 64)):
        P.width, P.height = W, H
        for world in (1, 2, 4, 8):
            owned = set()
            for rank in range(world):
                tiles = replica.local_tiles(W, H, 64, 64, rank, world)
                shard = api.HdTileShard(64, 64, rank, world)
                assert len(tiles) * 64 * 64 == L.hd_tile_shard_pixels(C.byref(P), C.byref(shard))
                tx, ty = C.c_uint32(), C.c_uint32()
                for lt, ex, ey in tiles:    # same map as the kernel's (hd_tile_shard_locate shares tile_of with it)
                    assert L.hd_tile_shard_locate(C.byref(P), C.byref(shard), lt, C.byref(tx), C.byref(ty)) == 0
                    assert (tx.value, ty.value) == (ex, ey)
                    assert replica.tile_owner(ex, ey, -(-W // 64), world) == rank
                assert L.hd_tile_shard_locate(C.byref(P), C.byref(shard), len(tiles), C.byref(tx), C.byref(ty)) != 0
                assert not owned & {(x, y) for _, x, y in tiles}
                owned |= {(x, y) for _, x, y in tiles}
            assert len(owned) == (-(-W // 64)) * (-(-H // 64))
    # a row holding a whole number of rounds must not give a rank the same columns in every row
    cols = {tx for _, tx, ty in replica.local_tiles(7680, 4320, 64, 64, 0, 8)}
    assert len(cols) == 120


def test_bench_frame_dims_weak_scaling():
    sys.path.insert(0, ROOT)
    import bench
    for n in (1, 2, 4, 8):
        w, h = bench.frame_dims(n)
        assert w * h == n * 3840 * 2160 and w % 64 == 0
