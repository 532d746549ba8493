"""Multi-GPU plumbing: the path shards by contig — one process per GPU, contigs
assigned longest-first, no data-path collective — and ends with ONE all-gather that concatenates
the per-rank candidate buffers before VCF emission.  torch.distributed is used for the
collective only (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import numpy as np


def lpt_assign(weights, n_ranks):
    """contig -> rank, longest-processing-time first.  The reference's unit of parallelism is the contig
    (sniffles:313-358); clusters never cross a task, so any contig partition gives identical calls."""
    load = [0] * n_ranks
    owner = [0] * len(weights)
    for c in sorted(range(len(weights)), key=lambda k: (-weights[k], k)):
        r = min(range(n_ranks), key=lambda k: (load[k], k))
        owner[c] = r
        load[r] += weights[c]
    return owner


def subset_block(block, task_ids):
    """The records of the given tasks as a block of their own (arenas are shared, offsets stay valid)."""
    keep = np.isin(block.rec["task"], np.asarray(sorted(task_ids), dtype=block.rec["task"].dtype))
    return type(block)(rec=np.ascontiguousarray(block.rec[keep]), cigar=block.cigar, var=block.var, seq=block.seq, task=block.task,
                       contig=block.contig, tr=block.tr, contig_names=block.contig_names)


def allgather_bytes(local, group=None):
    """All-gather of variable-length uint8 tensors: sizes first, then max-padded payloads.  Returns the list of
    per-rank tensors (trimmed).  `local` lives on the device of the backend (cuda for nccl, cpu for gloo)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    n = torch.tensor([local.numel()], dtype=torch.int64, device=local.device)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n, group=group)
    sizes = [int(s.item()) for s in sizes]
    mx = max(max(sizes), 1)
    pad = torch.zeros(mx, dtype=torch.uint8, device=local.device)
    pad[:local.numel()] = local
    outs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(outs, pad, group=group)
    return [o[:s] for o, s in zip(outs, sizes)]


def gather_struct_arrays(arr, group=None, device="cpu"):
    """All-gather a numpy structured array (e.g. snfb_cand records); returns the per-rank arrays."""
    import torch
    raw = np.ascontiguousarray(arr).view(np.uint8).reshape(-1)
    t = torch.from_numpy(raw.copy()).to(device)
    parts = allgather_bytes(t, group)
    return [np.frombuffer(p.cpu().numpy().tobytes(), dtype=arr.dtype) for p in parts]


class DeviceBytes:
    """zero-copy view of a library-owned device buffer for torch (``torch.as_tensor(DeviceBytes(p, n), device='cuda')``)"""

    def __init__(self, ptr, nbytes):
        self.__cuda_array_interface__ = {"shape": (int(nbytes),), "typestr": "|u1", "data": (int(ptr), False), "version": 3}


def merge_rank_candidates(parts):
    """Concatenate per-rank candidate arrays in reference emission order: task id, then each rank's own order
    (sniffles:544-547 sorts finished tasks by id)."""
    allc = np.concatenate(parts) if parts else parts
    return allc[np.argsort(allc["task"], kind="stable")]
