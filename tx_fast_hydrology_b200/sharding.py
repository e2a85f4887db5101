"""Multi-GPU sharding helpers (one process per GPU, torch.distributed for the plumbing).

Two natural shardings of the routing path:
  * by ensemble member: every rank holds the whole network and a slice of the member columns;
    routing needs no communication, the assimilation combines its statistics with collectives
    (all-reduce of the row sums, all-gather of the gauge rows / anomalies);
  * by independent basin: disjoint trees share nothing -- whole basins are bin-packed onto ranks
    and routed with no communication at all.
"""
import numpy as np


def combine_statistics(rowsum, HX, members_total, group=None):
    """rowsum [n] and HX [m][Mloc] of this rank -> (ensemble mean [n], HX of all ranks [m][Mtot])."""
    import torch
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(rowsum, group=group)
        parts = [torch.empty_like(HX) for _ in range(dist.get_world_size(group))]
        dist.all_gather(parts, HX.contiguous(), group=group)
        HX = torch.cat(parts, dim=1).contiguous()
    return rowsum / float(members_total), HX


def shard_basins(basin, world):
    """Greedy bin-packing of whole basins (largest first) into `world` shards balanced by reach count.
    Returns a list of basin-id lists."""
    basin = np.asarray(basin)
    sizes = np.bincount(basin)
    order = np.argsort(-sizes, kind="stable")
    load = [0] * world
    parts = [[] for _ in range(world)]
    for b in order:
        if sizes[b] == 0:
            continue
        r = int(np.argmin(load))
        parts[r].append(int(b)); load[r] += int(sizes[b])
    return parts


def extract_shard(endnodes, basin, basin_ids):
    """Sub-network of the given basins: (local endnodes int64[k], global reach index of each local reach)."""
    endnodes = np.asarray(endnodes, dtype=np.int64)
    idx = np.flatnonzero(np.isin(basin, basin_ids))
    local = np.full(endnodes.size, -1, dtype=np.int64)
    local[idx] = np.arange(idx.size)
    sub = local[endnodes[idx]]
    assert (sub >= 0).all(), "a basin drains outside its shard"
    return sub, idx
