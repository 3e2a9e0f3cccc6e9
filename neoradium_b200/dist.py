"""Multi-GPU plumbing: one process per GPU (torchrun), code blocks sharded at transport-block granularity, no
collective in the data path; only the link-level counters are summed -- one all-reduce of an int64[8] vector per SNR
point (NCCL over NVLink on the GPU box, gloo in the CPU tests).  SURVEY.md 8e.
"""
import os

import torch
import torch.distributed as dist

COUNTER_NAMES = ("codeBlocks", "cbCrcFail", "txBlocks", "tbCrcFail", "bitErrors", "sumIterations", "rsv0", "rsv1")


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(num_tb, rank, world):
    """Contiguous slice [lo, hi) of the transport blocks owned by `rank`; a transport block (its C code blocks and its
    HARQ soft buffer) never straddles two GPUs."""
    base, rem = divmod(int(num_tb), int(world))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_counters(counters, group=None):
    """In-place SUM all-reduce of the int64[8] counter vector; a no-op without an initialised process group."""
    assert counters.dtype == torch.int64 and counters.numel() == len(COUNTER_NAMES)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(counters, op=dist.ReduceOp.SUM, group=group)
    return counters


def counters_dict(counters):
    v = counters.detach().cpu().tolist()
    d = dict(zip(COUNTER_NAMES, v))
    d["bler"] = d["tbCrcFail"] / d["txBlocks"] if d["txBlocks"] else None
    d["cbler"] = d["cbCrcFail"] / d["codeBlocks"] if d["codeBlocks"] else None
    d["meanIterations"] = d["sumIterations"] / d["codeBlocks"] if d["codeBlocks"] else None
    return d


def bler_point(codec, num_tb_total, snr_db, num_iter, seed, batch_tbs=64, group=None):
    """One SNR point of a BLER sweep, sharded over the ranks of `group`: payload -> TX chain -> fused QAM + AWGN + max-log
    LLR kernel (`nrldpc_awgn_llr`) -> fused RX chain -> counters, all on the device; returns the globally reduced
    counters as a dict (same on every rank).  Payload bits and the noise of a symbol depend on (seed, global
    transport-block index, position) only, so the counters do not change with the number of GPUs or the batch size."""
    from .modulation import awgn_llr
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    lo, hi = shard_range(num_tb_total, rank, world)
    dev = codec.device
    counters = torch.zeros(len(COUNTER_NAMES), dtype=torch.int64, device=dev)
    sym_per_tb = codec.sumE // codec.qm
    done = lo
    while done < hi:
        n = min(batch_tbs, hi - done)
        payload = codec.random_payload(n, seed=int(seed) ^ 0x5bd1e995, firstTb=done)   # function of the global block index only
        llr = awgn_llr(codec.encode(payload), codec.qm, snr_db=snr_db, seed=seed, offset=done * sym_per_tb)
        out = codec.decode(llr, num_iter)
        codec.accumulate(out, counters, refPayload=payload)
        done += n
    reduce_counters(counters, group)
    return counters_dict(counters)
