"""vkhashdag_b200/replica.py — multi-GPU plumbing: one process per GPU, pool replicated, frame sharded by screen tile.

Trace needs no collective (every rank owns the tiles t % world == rank of hd_tile_shard).  Edits run on ONE rank; its
dirty ranges reach the replicas through ONE broadcast of the packed staging buffer (hd_dirty_pack_dev /
hd_dirty_apply_dev), which replaces DAGNodePool::Flush (src/DAGNodePool.cpp:56-85) for the replicated case.  The
root is published last on every replica (ordering contract of src/main.cpp:281-296).

`dist` is torch.distributed (backend "nccl" on GPUs; "gloo" in the CPU tests, which drive the same protocol with a
host-memory stand-in for the pool).
"""
import torch


def tile_owner(tx, ty, tiles_x, world):
    """Rank that owns screen tile (tx, ty) — the hd_tile_shard map of include/hashdag_b200.h: row-major round robin,
    or (tx + ty) % world when a row holds a whole number of rounds (so that ranks do not end up with fixed columns)."""
    return (ty * tiles_x + tx) % world if tiles_x % world else (tx + ty) % world


def local_tiles(width, height, tile_w, tile_h, rank, world):
    """[(local_index, tile_x, tile_y)] of the tiles a rank owns, in its output order."""
    tiles_x, tiles_y = -(-width // tile_w), -(-height // tile_h)
    if tiles_x % world:
        return [(lt, t % tiles_x, t // tiles_x) for lt, t in enumerate(range(rank, tiles_x * tiles_y, world))]
    per_row = tiles_x // world
    return [(ty * per_row + k, k * world + (rank - ty) % world, ty) for ty in range(tiles_y) for k in range(per_row)]


def assemble_frame(parts, width, height, tile_w, tile_h, world, dtype=None):
    """Stitch per-rank tile-major planes (index = rank) back into a row-major frame (host-side, for checks/display)."""
    import numpy as np
    frame = np.zeros((height, width), dtype=dtype or parts[0].dtype)
    for rank, part in enumerate(parts):
        for lt, tx, ty in local_tiles(width, height, tile_w, tile_h, rank, world):
            x0, y0 = tx * tile_w, ty * tile_h
            w, h = min(tile_w, width - x0), min(tile_h, height - y0)
            blk = part[lt * tile_w * tile_h:(lt + 1) * tile_w * tile_h].reshape(tile_h, tile_w)
            frame[y0:y0 + h, x0:x0 + w] = blk[:h, :w]
    return frame


HEADER_BYTES = 32


def packed_bytes(head):
    """Size in bytes of a packed blob, from its 8-word header (the layout is documented in sync.cu)."""
    h = [int(x) for x in head[:8]]
    words = 8 + 3 * h[0] + h[1]
    if h[3] & 2:   # colour section: two totals, its triples, its payload
        words += 2 + 3 * h[4] + h[5]
    return 4 * words


class StagingTooSmall(Exception):
    """DirtyPack could not fit the packed dirty ranges; `.need` is the size in bytes that would."""

    def __init__(self, need):
        super().__init__(f"staging buffer too small: need {need} bytes")
        self.need = int(need)


class ReplicaSync:
    """Broadcast the editing rank's dirty ranges to every replica.

    Latency matters here (the interactive loop publishes every frame), so the protocol spends ONE collective on the
    common case: every publish broadcasts a fixed-size first chunk of the staging buffer (`eager_bytes`, the same on
    every rank) that starts with the 16-byte header [n_ranges][payload_words][root][clear_first].  A brush edit's
    dirty ranges (tens of bytes to tens of KB) fit in that chunk; only a larger payload (the initial replica copy, a
    big batch) needs a second broadcast for the remainder, whose size the replicas learn from the header."""

    def __init__(self, pool, dist, device, capacity_bytes=64 << 20, eager_bytes=256 << 10):
        self.pool, self.dist, self.device = pool, dist, device
        self.eager = max(HEADER_BYTES, min(int(eager_bytes), int(capacity_bytes)) & ~3)
        self.staging = torch.empty(max(int(capacity_bytes), self.eager), dtype=torch.uint8, device=device)
        self.collectives = 0   # payload collectives issued so far (tests / benches)

    def _grow(self, need, keep=0):
        if need > self.staging.numel():
            old = self.staging
            self.staging = torch.empty(int(need * 1.5), dtype=torch.uint8, device=self.device)
            if keep:
                self.staging[:keep].copy_(old[:keep])

    def publish(self, src=0):
        """Call on EVERY rank after rank `src` edited (and set its root).  Returns the packed size in bytes.

        On GPUs the collectives are issued with the pool's stream current, so NCCL orders itself after the pack
        kernels and the scatter kernel after the broadcast without host synchronisation in between."""
        if torch.device(self.device).type == "cuda":
            stream = torch.cuda.ExternalStream(self.pool.stream, device=torch.device(self.device))
            with torch.cuda.stream(stream):
                return self._publish(src)
        return self._publish(src)

    def _publish(self, src):
        rank = self.dist.get_rank()
        n = 0
        if rank == src:
            try:
                n = self.pool.DirtyPack(self.staging.data_ptr(), self.staging.numel())
            except StagingTooSmall as e:       # nothing changed in the pool: grow and pack again
                self._grow(e.need)
                n = self.pool.DirtyPack(self.staging.data_ptr(), self.staging.numel())
        self.dist.broadcast(self.staging[:self.eager], src)   # header + (normally) the whole payload
        self.collectives += 1
        if rank != src:
            n = packed_bytes(self.staging[:HEADER_BYTES].view(torch.int32).cpu().numpy().view("uint32"))
        if n > self.eager:                                     # large payload: one more broadcast for the remainder
            self._grow(n, keep=self.eager)
            self.dist.broadcast(self.staging[self.eager:n], src)
            self.collectives += 1
        if rank == src:
            self.pool.DirtyReset()
        else:
            self.pool.DirtyApply(self.staging.data_ptr(), n)   # scatter kernel, bucket_words, root last
        return n
