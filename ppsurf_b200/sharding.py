"""Multi-GPU predict: one process per GPU over ``torch.distributed``.

The reference predicts on a single device (``--trainer.devices 1``, full_run_pps.py:24-25); both loops of its ``predict_step``
are independent across their units, which is what is sharded here:
  * latent loop (source/poco_model.py:203-236): the encoder passes are dealt to the ranks (pass p goes to rank p mod G), every rank
    accumulates its passes into ``latent_sum [N,C]`` / ``counts [N]`` and ONE all-reduce(sum) of the two precedes the division
    (option B of §8e; 102 MB + 0.4 MB for a 100k-point cloud);
  * occupancy queries (source/poco_utils.py:218, 146-168): every sweep's query list is cut into G contiguous slices, each rank decodes
    its slice, one all-gather returns all values to every rank, which then updates its own copy of the volume  --  the region-growing
    bookkeeping is replicated, not communicated.
``Shard(world=1)`` is the single-GPU case: no collective is issued.
"""
import typing

import torch
import torch.distributed as dist


class Shard:

    def __init__(self, world: int = 1, rank: int = 0, group=None):
        self.world, self.rank, self.group = int(world), int(rank), group

    @classmethod
    def from_env(cls) -> 'Shard':
        """the default process group when ``torch.distributed`` is initialised, else the single-process shard"""
        if dist.is_available() and dist.is_initialized():
            return cls(dist.get_world_size(), dist.get_rank(), None)
        return cls()

    def common_seed(self, device) -> int:
        """one random 31-bit seed all ranks agree on (rank 0 draws it): every rank must walk the SAME schedule"""
        seed = torch.randint(0, 2 ** 31 - 1, (1,), dtype=torch.int64)
        if self.world > 1:
            seed = seed.to(device)
            dist.broadcast(seed, src=0, group=self.group)
        return int(seed.item())

    # ---- latent loop ---------------------------------------------------------------------------------------------------------
    def my_passes(self, passes: typing.Iterable) -> typing.Iterator:
        """pass p of the schedule belongs to rank p mod world (every rank walks the same seeded schedule)"""
        for p, item in enumerate(passes):
            if p % self.world == self.rank:
                yield item

    def reduce_latents(self, latent_sum: torch.Tensor, counts: torch.Tensor):
        """sum of the per-rank partial accumulations, in place"""
        if self.world > 1:
            dist.all_reduce(latent_sum, op=dist.ReduceOp.SUM, group=self.group)
            dist.all_reduce(counts, op=dist.ReduceOp.SUM, group=self.group)

    # ---- occupancy queries -----------------------------------------------------------------------------------------------------
    def slice_of(self, n: int) -> typing.Tuple[int, int]:
        """``[first, first + count)`` of this rank in a list of ``n`` queries: contiguous, balanced within one"""
        first = n * self.rank // self.world
        return first, n * (self.rank + 1) // self.world - first

    def evaluate(self, fn: typing.Callable[[torch.Tensor], torch.Tensor], queries: torch.Tensor) -> torch.Tensor:
        """``fn(queries)`` computed slice-wise across the ranks: every rank returns the values of ALL queries"""
        n = queries.shape[0]
        if self.world == 1 or n == 0:
            return fn(queries)
        first, count = self.slice_of(n)
        width = (n + self.world - 1) // self.world  # slices are padded to one width for the all-gather
        local = torch.zeros((width,), dtype=torch.float32, device=queries.device)
        if count > 0:
            local[:count] = fn(queries[first:first + count].contiguous())
        gathered = torch.empty((self.world * width,), dtype=torch.float32, device=queries.device)
        dist.all_gather_into_tensor(gathered, local, group=self.group)
        out = torch.empty((n,), dtype=torch.float32, device=queries.device)
        for r in range(self.world):
            f = n * r // self.world
            c = n * (r + 1) // self.world - f
            out[f:f + c] = gathered[r * width:r * width + c]
        return out
