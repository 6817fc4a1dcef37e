"""Worker for tests/test_host_logic.py::test_replica_sharding_gloo_world2 (CPU, gloo)."""
import os

import torch
import torch.distributed as dist

from mmada_parallel_b200.parallel import max_over_ranks, shard_prompts, sum_over_ranks

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dist.init_process_group("gloo", rank=rank, world_size=world)
shard = shard_prompts(list(range(5)), rank, world)
mx = max_over_ranks(10.0 * (rank + 1), device="cpu")
tot = sum_over_ranks(len(shard), device="cpu")
print(f"OK rank{rank} shard={shard} max_ms={mx} total={int(tot)}")
dist.destroy_process_group()
