"""Multi-GPU host logic (one process per GPU, torch.distributed).

Since round 2 the product path for MSM is INSIDE the library (csrc/capi_multi.cu: b200_ctx_comm_init +
b200_g{1,2}_msm_sharded_dev, or the single-process b200_multi_*); this module keeps the same call sequence on top of
torch.distributed for hosts that bring their own collective, and is what the gloo CPU tests exercise.

MSM shards by scalar window (north_star / SURVEY §8e): every rank holds all points and scalars, computes the
partial group element  sum_{w = rank mod world} 2^(c*w) * S_w  with `b200_g{1,2}_msm_shard_dev`, the ranks
exchange their 144-byte (G1) / 288-byte (G2) partials with ONE all_gather, and each rank adds the partials
locally with complete projective additions (`b200_g{1,2}_sum_dev`).  NCCL has no elliptic-curve reduction
operator, so "allreduce of partial sums" = all_gather + local add.
Pairing batches / scalar-mul batches shard by index with no collective (`index_range`).
The pairing PRODUCT (multi_miller_loop over n terms, SURVEY §8e "mode ii") shards by term index: every rank multiplies
the Miller-loop values of its terms into one Fp12, the ranks exchange those 576-byte partials with one all_gather,
multiply them, and one final exponentiation follows (`ShardedPairingProduct`).
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
        if getattr(self.eng, "comm_world", 1) == self.world:
            # the library owns the NCCL communicator (b200_ctx_comm_init): shard + ncclAllGather + combine on one stream,
            # no host synchronisation in between (round 2; the torch.distributed route below stays for hosts without it)
            self.eng.msm_sharded_dev(self.k, xy, inf, scalars, n, out, mode=self.mode)
            return out
        if self.mode == "window":
            self.eng.msm_dev(self.k, xy, inf, scalars, n, out, shard=self.rank, n_shards=self.world)
        else:                                    # point-range sharding: all windows of my slice of the points
            lo, hi = index_range(n, self.rank, self.world)
            self.eng.msm_dev(self.k, xy[lo:hi], None if inf is None else inf[lo:hi], scalars[lo:hi], hi - lo, out)
        self._gather(parts, out)
        self.eng.sum_dev(self.k, parts, self.world, out)
        return out


# Fp12::one() as 72 little-endian u64 limbs viewed as int64 (c0.c0.c0 = R = 2^384 mod p, src/fp.rs:83-90)
_FP12_ONE = [0x760900000002fffd, 0xebf4000bc40c0002, 0x5f48985753c758ba, 0x77ce585370525745, 0x5c071a97a256ec6d,
             0x15f65ec3fa80e493] + [0] * 66


def _to_i64(v):
    return v - (1 << 64) if v >= (1 << 63) else v


class ShardedPairingProduct:
    """multi_miller_loop(&[(p_i, q_i)]) (src/pairings.rs:554-603) over n terms held by every rank, sharded by term index.
    `engine` needs miller_loop_batch_dev(p, pinf, q, qinf, n, out), fp12_product_dev(f, n, out) and
    final_exponentiation_batch_dev(f, n, out); `stream` = the engine's CUDA stream (torch.cuda.ExternalStream), so the
    collective is ordered with the engine's kernels, as in ShardedMSM.  The product of per-term Miller values equals the reference's shared-
    squaring loop value (f <- f^2 * prod l_i), and identity terms contribute one(), so the MillerLoopResult — and the
    Gt after final_exponentiation — are limb-identical to the single-GPU / reference result."""

    def __init__(self, engine, dist=None, stream=None):
        self.eng, self.dist, self.stream = engine, dist, stream
        self.rank = dist.get_rank() if dist is not None else 0
        self.world = dist.get_world_size() if dist is not None else 1

    def multi_miller_loop(self, p, pinf, q, qinf, n, out, parts, scratch, final_exp=False):
        """out (1, 72) <- product over all n terms on every rank.  p (n,12), q (n,24) affine limbs, pinf/qinf (n,) uint8
        or None; parts = (world, 72) and scratch = (ceil(n / world), 72) int64 work tensors on the engine's device."""
        lo, hi = index_range(n, self.rank, self.world)
        m = hi - lo
        local = parts[self.rank:self.rank + 1] if self.world == 1 else out
        if m == 0:                              # more ranks than terms: this rank contributes Fp12::one()
            one = torch.tensor([_to_i64(v) for v in _FP12_ONE], dtype=torch.int64).reshape(1, 72)
            if self.stream is not None:         # same stream as the collective below
                with torch.cuda.stream(self.stream):
                    local.copy_(one)
            else:
                local.copy_(one)
        else:
            sl = lambda t: None if t is None else t[lo:hi]
            self.eng.miller_loop_batch_dev(p[lo:hi], sl(pinf), q[lo:hi], sl(qinf), m, scratch)
            self.eng.fp12_product_dev(scratch, m, local)
        if self.world > 1:
            if self.stream is not None:
                with torch.cuda.stream(self.stream):
                    self.dist.all_gather_into_tensor(parts, local)
            else:
                self.dist.all_gather_into_tensor(parts, local)
        self.eng.fp12_product_dev(parts, self.world, out)
        if final_exp:
            self.eng.final_exponentiation_batch_dev(out, 1, out)
        return out
