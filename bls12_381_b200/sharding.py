"""Multi-GPU host logic (one process per GPU, torch.distributed).

MSM shards by scalar window (north_star / SURVEY §8e): every rank holds all points and scalars, computes the
partial group element  sum_{w = rank mod world} 2^(c*w) * S_w  with `b200_g{1,2}_msm_shard_dev`, the ranks
exchange their 144-byte (G1) / 288-byte (G2) partials with ONE all_gather, and each rank adds the partials
locally with complete projective additions (`b200_g{1,2}_sum_dev`).  NCCL has no elliptic-curve reduction
operator, so "allreduce of partial sums" = all_gather + local add.
Pairing batches / scalar-mul batches shard by index with no collective (`index_range`).
"""
import torch


def index_range(n, rank, world):
    """contiguous share [lo, hi) of n independent items for `rank` (pairs shard by index)"""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def windows_of(rank, world, c):
    """global window indices handled by `rank` (stride sharding keeps the per-rank Horner chains balanced)"""
    nwin = (256 + c - 1) // c
    return list(range(rank, nwin, world))


class ShardedMSM:
    """k = 1 (G1) or 2 (G2).  `engine` needs msm_dev(k, xy, inf, s, n, out, shard=, n_shards=) and
    sum_dev(k, parts, n, out); `stream` (optional) is the CUDA stream the engine works on, so the collective
    is enqueued behind the engine's kernels without a host sync."""

    def __init__(self, engine, k, dist=None, stream=None, mode="window"):
        self.eng, self.k, self.dist, self.stream, self.mode = engine, k, dist, stream, mode
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def _gather(self, parts, out):
        if self.stream is not None:
            with torch.cuda.stream(self.stream):
                self.dist.all_gather_into_tensor(parts, out)
        else:
            self.dist.all_gather_into_tensor(parts, out)

    def msm(self, xy, inf, scalars, n, out, parts):
        """out (1, 18k) <- sum_i points[i]*scalars[i] on every rank; parts = (world, 18k) scratch"""
        if self.world == 1:
            self.eng.msm_dev(self.k, xy, inf, scalars, n, out)
            return out
        if self.mode == "window":
            self.eng.msm_dev(self.k, xy, inf, scalars, n, out, shard=self.rank, n_shards=self.world)
        else:                                    # point-range sharding: all windows of my slice of the points
            lo, hi = index_range(n, self.rank, self.world)
            self.eng.msm_dev(self.k, xy[lo:hi], None if inf is None else inf[lo:hi], scalars[lo:hi], hi - lo, out)
        self._gather(parts, out)
        self.eng.sum_dev(self.k, parts, self.world, out)
        return out
