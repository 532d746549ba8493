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
    """The records of the given tasks as a block of their own (arenas are shared, offsets stay valid; the parent's owner,
    reference N mask and CIGAR16 twin are carried along)."""
    keep = np.isin(block.rec["task"], np.asarray(sorted(task_ids), dtype=block.rec["task"].dtype))
    sub = type(block)(rec=np.ascontiguousarray(block.rec[keep]), cigar=block.cigar, var=block.var, seq=block.seq, task=block.task,
                      contig=block.contig, tr=block.tr, contig_names=block.contig_names, sites=block.sites, _owner=block._owner,
                      mask=block.mask, mask_task_off=block.mask_task_off)
    if block.cigar16 is not None:
        sub.rec16, sub.cigar16 = np.ascontiguousarray(block.rec16[keep]), block.cigar16
    return sub


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


def merge_results(parts):
    """Host restatement of the library's k_gather_merge (csrc/api.cu): concatenate per-rank results in rank order and rebase
    alt_off / lead_off / long_off and the rnames offsets into the merged arenas.  parts: objects with cand, alt, rnames, rn_off,
    cand_leads (binding.Result / OracleResult).  Returns a binding.GatheredResult."""
    from .binding import GatheredResult
    from . import abi
    g = GatheredResult()
    cands, alts, rns, offs, leads = [], [], [], [], []
    b_alt = b_rn = b_leads = 0
    for p in parts:
        c = np.array(p.cand, copy=True)
        c["alt_off"] = np.where(c["alt_off"] >= 0, c["alt_off"] + b_alt, c["alt_off"])
        c["lead_off"] += b_leads
        c["long_off"] += b_leads
        cands.append(c)
        alts.append(np.asarray(p.alt, dtype="u1"))
        rns.append(np.asarray(p.rnames, dtype="<u8"))
        offs.append(np.asarray(p.rn_off[:len(c)], dtype="<u4") + np.uint32(b_rn))
        leads.append(np.asarray(p.cand_leads))
        b_alt += len(p.alt)
        b_rn += len(p.rnames)
        b_leads += len(p.cand_leads)
    g.cand = np.concatenate(cands) if cands else np.zeros(0, abi.CAND_DTYPE)
    g.alt = np.concatenate(alts) if alts else np.zeros(0, "u1")
    g.rnames = np.concatenate(rns) if rns else np.zeros(0, "<u8")
    g.rn_off = np.concatenate(offs + [np.asarray([b_rn], dtype="<u4")])
    g.cand_leads = np.concatenate(leads) if leads else np.zeros(0, abi.LEAD_DTYPE)
    g.n_cand, g.n_alt_bytes, g.n_rnames, g.n_cand_leads = len(g.cand), len(g.alt), len(g.rnames), len(g.cand_leads)
    return g


class _Part:
    pass


def gather_results(res, group=None, device="cpu"):
    """All-gather a rank's whole result (candidate records, ALT arena, read names + offsets, candidate leads) through
    torch.distributed and merge it (gloo on CPU: the test of the N > 1 host logic; the GPU path uses the library's own
    NCCL all-gather, snfb_allgather_candidates)."""
    from . import abi
    parts = None
    for name, dt in (("cand", abi.CAND_DTYPE), ("alt", np.dtype("u1")), ("rnames", np.dtype("<u8")), ("rn_off", np.dtype("<u4")), ("cand_leads", abi.LEAD_DTYPE)):
        arr = np.ascontiguousarray(getattr(res, name))
        got = gather_struct_arrays(arr.view(dt) if arr.dtype != dt else arr, group, device)
        if parts is None:
            parts = [_Part() for _ in got]
        for p, a in zip(parts, got):
            setattr(p, name, a)
    return merge_results(parts)


def merge_rank_candidates(parts):
    """Concatenate per-rank candidate arrays in reference emission order: task id, then each rank's own order
    (sniffles:544-547 sorts finished tasks by id).  Records only: use merge_results for the arenas."""
    allc = np.concatenate(parts) if parts else parts
    return allc[np.argsort(allc["task"], kind="stable")]
