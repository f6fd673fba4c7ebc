"""Host-side logic of the multi-GPU path: individuals are sharded over ranks (one process per GPU); what crosses ranks on the
host is small -- the per-marker class counts (once, at load time), three scalars of the phenotype standardisation, two scalars
per sweep (sum(eps) for mu, ||eps||^2 for sigmaE) and, once, the IPC handles of the exchange mailboxes.  The per-block exchange
of the sweep itself happens inside the kernel over NVLink peer memory (csrc/sweep_dense.cuh), not here.

Works over any torch.distributed backend: NCCL on the GPUs, gloo in the CPU tests (tests/test_sharding_gloo.py)."""
import numpy as np


def shard_bounds(n_inds, world):
    """Row ranges [lo, hi) of the ranks: equal slices rounded up to 64 rows (code words, MMA k-steps and TMA granules stay aligned;
    lo is a multiple of 4, so a shard's PLINK payload starts at a byte boundary).  Trailing ranks may be empty for tiny N."""
    per = ((n_inds + world - 1) // world + 63) // 64 * 64
    return [(min(n_inds, r * per), min(n_inds, (r + 1) * per)) for r in range(world)]


def shard_bed(bed, lo, hi):
    """Rows lo:hi of a PLINK SNP-major payload [M, ceil(N/4)] (lo a multiple of 4)."""
    assert lo % 4 == 0
    return np.ascontiguousarray(bed[:, lo // 4:(hi + 3) // 4])


def class_counts_host(bed, n_inds):
    """Per-marker genotype class counts (n0, n1, n2, n_missing) of a PLINK payload: the host twin of count_classes_kernel
    (PLINK 00 -> 2 copies, 10 -> 1, 11 -> 0, 01 -> missing: reference src/preprocessgraph.cpp:44-53)."""
    M = bed.shape[0]
    out = np.zeros((M, 4), dtype=np.uint32)
    cls_of_plink = np.array([2, 3, 1, 0])
    for k in range(4):
        n_k = len(range(k, n_inds, 4))
        codes = (bed[:, :n_k] >> (2 * k)) & 3
        cls = cls_of_plink[codes]
        for c in range(4):
            out[:, c] += (cls == c).sum(axis=1).astype(np.uint32)
    return out


class ShardGroup:
    """The collectives the sharded chain needs, over torch.distributed (world == 1: no-ops)."""

    def __init__(self, rank=0, world=1, device=None):
        self.rank, self.world = rank, world
        self.device = device
        if world > 1:
            import torch.distributed as dist
            assert dist.is_initialized(), "init_process_group first"
            self.dist = dist
            if device is None:
                self.device = "cuda" if dist.get_backend() == "nccl" else "cpu"

    def _allreduce(self, arr, dtype):
        import torch
        t = torch.from_numpy(np.ascontiguousarray(arr).astype(dtype)).to(self.device)
        self.dist.all_reduce(t)
        return t.cpu().numpy()

    def allreduce_counts(self, counts):
        """Global class counts from the shards' (exact: integers)."""
        if self.world == 1:
            return counts
        return self._allreduce(counts, np.int64).astype(np.uint32)

    def allsum(self, x):
        """Sum of one double over the ranks (every rank gets the bit-identical value: one allreduce)."""
        if self.world == 1:
            return float(x)
        return float(self._allreduce(np.array([x], dtype=np.float64), np.float64)[0])

    def standardise(self, y_local, n_total):
        """The reference's phenotype standardisation (src/BayesRBase.cpp:85-86) on a sharded vector: y passes through float32
        (VectorXf y, src/data.hpp:92), is centred by the global mean and scaled to unit sample variance."""
        y = np.asarray(y_local, dtype=np.float32).astype(np.float64)
        y = y - self.allsum(y.sum()) / n_total
        return y / np.sqrt(self.allsum((y ** 2).sum()) / (n_total - 1))

    def gather(self, obj):
        """All ranks' objects, in rank order (IPC handles of the exchange mailboxes; the phenotype for the CPU baseline)."""
        if self.world == 1:
            return [obj]
        out = [None] * self.world
        self.dist.all_gather_object(out, obj)
        return out
