#!/usr/bin/env python
"""bench.py — aligned long-read Gbp/s through lead -> cluster -> consensus (BASELINE.json metric).

    python bench.py --gpus 1 --steps K --warmup W                  # this repo's CUDA path
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...   # one rank per GPU
    python bench.py --impl reference --gpus N --steps K --warmup W  # CPU arm (oracle port, host cores)

A "step" is one pass of the hot path over the whole synthetic workload (BASELINE config 2:
30x ONT whole genome, 24 GRCh38-length contigs, ~6M reads at --scale 1).  With N ranks the
contigs are LPT-sharded over the ranks (strong scaling: the genome is fixed), every rank runs
the three stages on its contigs and the per-rank candidate buffers are concatenated with one
NCCL all-gather.  `value` is timed with the inputs resident in HBM (CUDA events on the
library's stream, max over ranks); `e2e` goes through the same C-ABI call with HOST buffers:
pinned host arenas -> H2D -> kernels -> D2H of the candidate SoA and ALT bytes, every step.
"""
import argparse
import ctypes as C
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)


def log(*a):
    print(*a, file=sys.stderr, flush=True)


def aligned_bp_passing(blk, cfg):
    """sum of query_alignment_length over the records that pass the A2 filters (leadprov.py:494-503)"""
    rec = blk.rec
    if len(rec) == 0:
        return 0
    first = blk.cigar[rec["cigar_off"]]
    last = blk.cigar[rec["cigar_off"] + rec["n_cigar"] - 1]
    lead = np.where((first & 15) == 4, first >> 4, 0).astype(np.int64)
    trail = np.where(((last & 15) == 4) & (rec["n_cigar"] > 1), last >> 4, 0).astype(np.int64)
    alen = rec["l_seq"].astype(np.int64) - lead - trail
    t = blk.task[rec["task"]]
    ok = (rec["mapq"] >= cfg.mapq) & ((rec["flag"] & 256) == 0) & (alen >= cfg.min_alignment_length) & (rec["pos"] >= t["start"]) & (rec["pos"] < t["end"])
    return int(alen[ok].sum())


def algorithmic_bytes_stage_a(blk, n_leads, n_pass):
    """Algorithmic bytes of one launch of the streaming kernel k_scan: per record one 16-byte scan descriptor and
    its CIGAR16 words (2 bytes each, the record's span rounded up to 16 bytes), read once; 12 bytes written per passing read
    (reference end, lead count, NM correction).  The 32-byte event-slice entries (at most one per lead) are left out: a lower bound."""
    w = blk.rec16["n_cigar"].astype(np.int64)
    rd = int(16 * len(blk.rec16) + 2 * (((w + 7) // 8) * 8).sum())
    return rd + 12 * int(n_pass), rd


class ClockSampler(threading.Thread):
    Q = "clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self._halt = index, [], threading.Event()

    def run(self):
        while not self._halt.is_set():
            try:
                out = subprocess.run(["nvidia-smi", f"--id={self.index}", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            self._halt.wait(0.05)

    def stop(self):
        self._halt.set()
        self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(float(r[0]) for r in self.rows if r[0].replace(".", "").isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": float(self.rows[0][1]) if self.rows[0][1].replace(".", "").isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


def measured_peak():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def ncu_traffic():
    """DRAM bytes per launch of the lead kernel from the committed ncu capture, if any."""
    p = os.path.join(ROOT, "profiles", "k_scan_dram.json")
    if os.path.exists(p):
        with open(p) as f:
            return json.load(f)
    return None


def bind_near_gpu(index):
    """Run on the CPUs of the GPU's NUMA node while the block is generated, packed and pinned: first-touch then places the host
    arenas next to the GPU's PCIe root, so the e2e host->device copy does not cross the socket interconnect.  Best effort."""
    try:
        import torch
        pr = torch.cuda.get_device_properties(index)
        dev = f"{pr.pci_domain_id:04x}:{pr.pci_bus_id:02x}:{pr.pci_device_id:02x}.0"
        with open(f"/sys/bus/pci/devices/{dev}/numa_node") as f:
            node = int(f.read().strip())
        if node < 0:
            log(f"[bench] GPU {index} ({dev}): no NUMA node reported")
            return None
        with open(f"/sys/devices/system/node/node{node}/cpulist") as f:
            cpus = set()
            for part in f.read().strip().split(","):
                a, _, b = part.partition("-")
                cpus.update(range(int(a), int(b or a) + 1))
        old = os.sched_getaffinity(0)
        new = cpus & old
        if not new or new == old:
            log(f"[bench] GPU {index} ({dev}) is on NUMA node {node}; affinity unchanged ({len(old)} CPUs)")
            return None
        os.sched_setaffinity(0, new)
        log(f"[bench] GPU {index} ({dev}) is on NUMA node {node}: building the block on its {len(new)} of {len(old)} CPUs")
        return old
    except Exception as e:          # no sysfs entry, old torch, container without the node files
        log(f"[bench] NUMA binding skipped: {e}")
        return None


def workload(args, mask, threads):
    from sniffles_b200 import synth
    t0 = time.time()
    blk = synth.config_block(args.config, args.scale, threads=threads, contig_mask=mask)
    log(f"[bench] generated config {args.config} scale {args.scale}: {len(blk.rec)} records, {blk.cigar.nbytes / 1e9:.2f} GB CIGAR, {blk.seq.nbytes / 1e9:.2f} GB seq in {time.time() - t0:.1f}s")
    return blk


def cpu_sample(blk, cfg, ccfg, threads, target_bp=6e9):
    """Bounded sample for the CPU arm: the trailing tasks of the block, up to ~target_bp aligned bases."""
    import oracle.oracle as orc
    from sniffles_b200 import abi
    rec = blk.rec
    tasks, counts = np.unique(rec["task"], return_counts=True)
    # the reference's unit of CPU parallelism is the contig: take the smallest contigs first so that the bounded sample
    # still keeps many host threads busy
    chosen, bp = [], 0
    for t in tasks[np.argsort(counts)]:
        chosen.append(int(t))
        bp += int(rec["l_seq"][rec["task"] == t].sum())
        if bp >= target_bp:
            break
    keep = np.isin(rec["task"], chosen)
    sub = type(blk)(rec=np.ascontiguousarray(rec[keep]), cigar=blk.cigar, var=blk.var, seq=blk.seq, task=blk.task, contig=blk.contig, tr=blk.tr, contig_names=blk.contig_names)
    abp = aligned_bp_passing(sub, cfg)
    nthr = max(1, min(threads, len(chosen)))
    t0 = time.perf_counter()
    res = orc.run(sub, ccfg, 3, nthr)
    dt = time.perf_counter() - t0
    cpu_sample.last_result = res
    return dict(value=abp / dt / 1e9, unit="Gbp/s", cores=nthr, kind="port",
                sample=f"oracle/snf_oracle.c (C port of the pure-Python reference) on tasks {sorted(chosen)} = {abp / 1e9:.2f} Gbp aligned, {len(res.cand)} candidates, {dt:.2f}s; "
                       "the reference itself measured 0.0546 Gbp/s/core (BASELINE.md §2)"), abp, dt


PARITY_FIELDS = ["task", "svtype", "pos", "end", "svlen", "support", "qual", "precise", "fwd", "rev", "cov_upstream", "cov_start", "cov_center", "cov_end",
                 "cov_downstream", "sa_count", "bnd_mate_contig", "bnd_mate_pos", "lead_n", "alt_off", "alt_len", "stdev_pos", "stdev_len"]


def same_candidates(dev, ora):
    """bit-for-bit comparison of the device run with the oracle pass the cpu_baseline leg just timed on the same block (checker only)"""
    if len(dev.cand) != len(ora.cand) or len(dev.alt) != len(ora.alt):
        return False
    for f in PARITY_FIELDS:
        a, b = dev.cand[f], ora.cand[f]
        if not (((a == b) | ((a != a) & (b != b))).all()):
            return False
    return bool((np.asarray(dev.alt) == np.asarray(ora.alt)).all())


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from sniffles_b200 import abi, config as sconfig
    cfg = sconfig.default_config()
    ccfg = abi.Config.from_sniffles(cfg)
    ncores = os.cpu_count() or 1
    # generate only the contigs the bounded sample will use (trailing contigs up to ~cpu_sample_gbp aligned bases)
    from sniffles_b200 import synth
    lens = [max(200000, int(x * args.scale)) for x in synth.GRCH38]
    mask, bp = [False] * len(lens), 0.0
    for c in sorted(range(len(lens)), key=lambda k: lens[k]):
        mask[c] = True
        bp += 30.0 * lens[c]
        if bp >= args.cpu_sample_gbp * 1e9:
            break
    blk = workload(args, mask, ncores)
    per_step = []
    info = None
    for i in range(args.warmup + args.steps):
        info, abp, dt = cpu_sample(blk, cfg, ccfg, ncores, target_bp=args.cpu_sample_gbp * 1e9)
        if i >= args.warmup:
            per_step.append(dt)
        if i == 0 and dt * (args.warmup + args.steps) > 240:     # keep the whole run within minutes
            args.steps = max(1, min(args.steps, int(240 / dt) - args.warmup))
    ms = 1e3 * sum(per_step) / len(per_step)
    val = abp / (ms / 1e3) / 1e9
    info["value"] = val
    print(json.dumps({"impl": "reference", "metric": "aligned long-read Gbp/s through lead->cluster->consensus", "value": val, "unit": "Gbp/s",
                      "n_gpus": args.gpus, "steps": len(per_step), "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "strong",
                      "vs_baseline": None, "dtype": "int32", "data": "synthetic",
                      "config": {"workload": f"BASELINE config {args.config}: synthetic 30x ONT WGS, scale {args.scale}; CPU arm on a bounded sample ({info['sample']})"},
                      "cpu_baseline": info, "e2e": {"value": val, "unit": "Gbp/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}))


def run_b200(args):
    import torch
    from sniffles_b200 import abi, binding, config as sconfig, synth, dist as sdist
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torch.distributed.run --nproc-per-node {args.gpus}")
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device: the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist.barrier()          # rank 0 has finished build() before anybody loads the libraries
    cfg = sconfig.default_config()
    ccfg = abi.Config.from_sniffles(cfg)
    lens = [max(200000, int(x * args.scale)) for x in synth.GRCH38]
    owner = sdist.lpt_assign(lens, world)
    mask = [o == rank for o in owner] if world > 1 else None
    ncores = os.cpu_count() or 1
    old_affinity = bind_near_gpu(local)
    blk = workload(args, mask, max(1, min(len(os.sched_getaffinity(0)), ncores // world if world > 1 else ncores)))
    abp_local = aligned_bp_passing(blk, cfg)
    L = binding.lib()
    t0 = time.time()
    blk.pack16()            # BAM CIGAR words -> CIGAR16, once per block (part of packing the block, like dropping the base qualities)
    log(f"[bench] packed CIGAR16: {blk.cigar.nbytes / 1e9:.2f} GB -> {blk.cigar16.nbytes / 1e9:.2f} GB in {time.time() - t0:.1f}s")
    pinned = []
    if not args.no_pin:
        t0 = time.time()
        for a in (blk.rec16, blk.cigar16, blk.var, blk.seq):
            if a.nbytes and L.snfb_pin_host(C.c_void_p(a.ctypes.data), a.nbytes) == 0:
                pinned.append(a)
        log(f"[bench] pinned {sum(a.nbytes for a in pinned) / 1e9:.2f} GB of host arenas in {time.time() - t0:.1f}s")
    if old_affinity:            # every thread (the OpenMP pool was created under the narrow mask) gets all CPUs back
        for tid in os.listdir("/proc/self/task"):
            try:
                os.sched_setaffinity(int(tid), old_affinity)
            except OSError:
                pass
    ctx = binding.Context(local)
    ctx.set_config(ccfg)
    ctx.load(blk)

    def gather(res):
        """one NCCL all-gather of the per-rank candidate buffers before VCF emission"""
        if world == 1:
            return len(res.cand)
        dptr, n = ctx.device_candidates()
        nbytes = n * abi.CAND_DTYPE.itemsize
        local = torch.as_tensor(sdist.DeviceBytes(dptr, nbytes), device="cuda") if nbytes else torch.zeros(0, dtype=torch.uint8, device="cuda")
        parts = sdist.allgather_bytes(local)
        return sum(p.numel() for p in parts) // abi.CAND_DTYPE.itemsize

    def step():
        res = ctx.run(want_leads=False, want_cands=True, want_seqs=True, copy=False)
        n_all = gather(res)
        return res, n_all

    def barrier():
        torch.cuda.synchronize()
        if dist is not None:
            dist.barrier()

    for _ in range(args.warmup):
        res, n_all = step()
    barrier()
    launches0 = ctx.launch_count()
    sampler = ClockSampler(local)
    sampler.start()
    dev_ms, kern = 0.0, {}
    t0 = time.perf_counter()
    for _ in range(args.steps):
        res, n_all = step()
        for name, ms, by in ctx.timings():
            if name == "h2d_records":
                continue
            if name == "total":
                dev_ms += ms
            else:
                kern.setdefault(name, [0.0, by])[0] += ms
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop()
    launches = ctx.launch_count() - launches0
    # max over ranks
    tt = torch.tensor([dev_ms, wall * 1e3, float(abp_local)], device="cuda", dtype=torch.float64)
    if dist is not None:
        mx = tt.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        sm = tt.clone()
        dist.all_reduce(sm, op=dist.ReduceOp.SUM)
        dev_ms_max, wall_ms_max, abp_total = float(mx[0]), float(mx[1]), float(sm[2])
    else:
        dev_ms_max, wall_ms_max, abp_total = dev_ms, wall * 1e3, float(abp_local)
    ms_per_step = wall_ms_max / args.steps
    value = abp_total / (ms_per_step / 1e3) / 1e9

    # ---- end to end through the C ABI with host buffers (H2D + kernels + D2H every step) ----
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    full_bytes = blk.rec16.nbytes + blk.cigar16.nbytes + blk.var.nbytes + blk.seq.nbytes
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        ctx.load(blk, seq_on_demand=not args.e2e_full_seq)      # host arenas; the 4-bit seq arena is fetched on demand (slices only)
        res, n_all = step()
    slice_bytes = sum(by for name, ms, by in ctx.timings() if name == "h2d_seq_slices")
    h2d = blk.rec16.nbytes + blk.cigar16.nbytes + blk.var.nbytes + (blk.seq.nbytes if args.e2e_full_seq else slice_bytes)
    barrier()
    e2e_wall = (time.perf_counter() - t0) / e2e_steps
    d2h = res.cand.nbytes + res.cand_leads.nbytes + res.rnames.nbytes + res.alt.nbytes
    et = torch.tensor([e2e_wall], device="cuda", dtype=torch.float64)
    if dist is not None:
        dist.all_reduce(et, op=dist.ReduceOp.MAX)
    e2e_val = abp_total / float(et[0]) / 1e9

    # ---- roofline of the dominant kernel (stage A lead extraction) ----
    ctx.load(blk)
    full = ctx.run(want_leads=True, want_cands=True, want_seqs=True, copy=False)       # also the run the full-size parity check compares with the oracle
    alg, alg_read = algorithmic_bytes_stage_a(blk, len(full.leads), full.n_pass)
    k_ms = kern.get("k_scan", [0.0, 0])[0] / args.steps
    peak, peak_src = measured_peak()
    achieved = alg / (k_ms / 1e3) / 1e9 if k_ms > 0 else 0.0
    traffic = ncu_traffic() if (args.config == 2 and args.scale == 1.0) else None      # the capture is of this workload at full size
    roof = {"bound": "hbm", "kernel": "extract::k_scan", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak if peak else None,
            "peak_source": peak_src, "algorithmic_bytes_per_launch": alg, "kernel_ms": k_ms,
            "traffic": traffic.get("dram_bytes_per_launch") if traffic else None}
    if rank == 0:
        out = {"metric": "aligned long-read Gbp/s through lead->cluster->consensus", "value": value, "unit": "Gbp/s", "n_gpus": world, "steps": args.steps,
               "warmup": args.warmup, "ms_per_step": ms_per_step, "device_ms_per_step": dev_ms_max / args.steps, "higher_is_better": True, "scaling": "strong",
               "vs_baseline": None, "dtype": "int32", "data": "synthetic",
               "config": {"workload": f"BASELINE config {args.config}: synthetic 30x ONT WGS, 24 GRCh38-length contigs x scale {args.scale}, {abp_total / 1e9:.2f} Gbp aligned, germline",
                          "records_rank0": int(len(blk.rec)), "candidates_total": int(n_all), "parallelism": f"contig LPT over {world} GPU(s), one NCCL all-gather of candidates",
                          "l2": f"inputs {full_bytes / 1e9:.2f} GB per rank >> 126 MB L2, no flush needed",
                          "cigar": f"CIGAR16 ({blk.cigar16.nbytes / 1e9:.2f} GB; the BAM words are {blk.cigar.nbytes / 1e9:.2f} GB)"},
               "clocks": clocks, "gpu_launches": int(launches),
               "e2e": {"value": e2e_val, "unit": "Gbp/s", "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h), "steps": e2e_steps, "ms_per_step": float(et[0]) * 1e3,
                       "pinned": bool(pinned), "seq": "full arena" if args.e2e_full_seq else "on demand (slices requested by the device, gathered on the host)"},
               "roofline": roof, "stage_ms": {k: v[0] / args.steps for k, v in kern.items()}}
        if world == 1 and not args.no_cpu:
            info, _, _ = cpu_sample(blk, cfg, ccfg, ncores, target_bp=args.cpu_sample_gbp * 1e9)
            out["cpu_baseline"] = info
            if args.cpu_sample_gbp * 1e9 >= abp_total:        # the oracle saw the whole block: compare it with the device run, candidate by candidate
                out["parity_full_size"] = {"candidates": int(len(full.cand)), "alt_bytes": int(len(full.alt)), "identical_to_oracle": same_candidates(full, cpu_sample.last_result)}
        print(json.dumps(out))
    for a in pinned:
        L.snfb_unpin_host(C.c_void_p(a.ctypes.data))
    ctx.close()
    if dist is not None:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", type=int, default=2, help="BASELINE.json config index (2 = 30x ONT WGS)")
    ap.add_argument("--scale", type=float, default=float(os.environ.get("SNFB_BENCH_SCALE", "1.0")), help="contig length multiplier (1.0 = full GRCh38 lengths)")
    ap.add_argument("--e2e-steps", type=int, default=2)
    ap.add_argument("--cpu-sample-gbp", type=float, default=1000.0, help="aligned Gbp of the CPU arm's sample (smallest contigs first); the default takes every contig: one host thread per contig, the reference's own grain")
    ap.add_argument("--no-pin", action="store_true")
    ap.add_argument("--e2e-full-seq", action="store_true", help="e2e: copy the whole seq arena every step instead of the on-demand slices")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "b200":
        log("[bench] note: timing rules ask for >= 3 warm-up steps")
    import __graft_entry__ as g
    if int(os.environ.get("RANK", "0")) == 0:
        g.build()
    elif args.impl == "b200":
        time.sleep(2.0)         # let rank 0 check/refresh the in-tree libraries first
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
