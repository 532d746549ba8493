"""Wider check of the combine path against the unmodified reference (build container only): random multi-sample shapes and arguments,
the reference's CombineTask vs sniffles_b200.combine (plan + oracle/combine.py grouping + call_group).  python oracle/pyref/fuzz_combine.py [rounds]"""
import json
import os
import random
import sys
import tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness                                     # noqa: E402
from oracle import combine as ocombine             # noqa: E402
from sniffles_b200 import combine, config as sconfig, snf, synth   # noqa: E402

ARGS = [[], ["--combine-pctseq", "0.985"], ["--combine-pctseq", "0.99", "--combine-separate-intra"], ["--combine-pctseq", "0"], ["--combine-separate-intra"], ["--combine-match", "60"], ["--combine-match-max", "300"], ["--combine-low-confidence", "0.5", "--combine-low-confidence-abs", "3"],
        ["--combine-output-filtered"], ["--combine-support-threshold", "6"], ["--combine-null-min-coverage", "30"], ["--combine-pair-relabel"], ["--combine-high-confidence", "0.5"],
        ["--dev-combine-medians"], ["--cluster-merge-bnd", "300"]]


def main(rounds):
    rng = random.Random(77)
    bad = 0
    for r in range(rounds):
        nsamp = rng.choice([2, 3, 5, 8])
        contigs = [rng.randrange(250_000, 900_000) for _ in range(rng.choice([1, 2, 3]))]
        spacing = rng.choice([4000.0, 9000.0, 20000.0])
        keep = rng.choice([0.4, 0.7, 1.0])
        seed = rng.randrange(1 << 30)
        args = rng.choice(ARGS)
        with tempfile.TemporaryDirectory() as d:
            paths = []
            for k in range(nsamp):
                blk = synth.generate(seed, contigs, rng.choice([12.0, 25.0]), len_model=1, len_mean=12000.0, len_sd=500.0, len_min=1000, len_max=80000, tech="ont",
                                     sv_spacing=spacing, sample=k + 1, site_keep=keep, threads=4)
                p = os.path.join(d, f"s{k}.snf")
                harness.write_reference_snf(blk, ["--sample-id", f"S{k}"], p)
                paths.append(p)
            names = [f"ctg{i + 1}" for i in range(len(contigs))]
            _, ref_calls, _ = harness.reference_combine(paths, list(zip(names, contigs)), args)
            cfg = sconfig.default_config(*args)
            cfg.mode = "combine"
            cfg.snf_input_info = [{"internal_id": k, "sample_id": f"S{k}", "filename": p} for k, p in enumerate(paths)]
            cfg.sample_ids_vcf = [(k, f"S{k}") for k in range(nsamp)]
            readers = {k: snf.SNFReader(p) for k, p in enumerate(paths)}
            ok, total = True, 0
            for tid, (name, length) in enumerate(zip(names, contigs)):
                task = combine.CombineTask(tid, name, 0, length - 1, cfg)
                plan = combine.Plan()
                task.plan(readers, plan)
                out = ocombine.combine_groups(combine.plan_arrays(plan, cfg), cfg)
                mine = combine.CombineTask.emit([task], plan, out)[0]
                a = [json.loads(json.dumps(harness.combine_call_dict(c))) for c in mine]
                b = [json.loads(json.dumps(harness.combine_call_dict(c))) for c in ref_calls[name]]
                total += len(b)
                if a != b:
                    ok = False
                    print("MISMATCH", r, name, len(a), len(b))
                    for x, y in zip(a, b):
                        if x != y:
                            print({k: (x[k], y[k]) for k in x if x[k] != y[k]})
                            break
            for rd in readers.values():
                rd.close()
            print(f"round {r}: samples {nsamp} contigs {len(contigs)} spacing {spacing} keep {keep} args {args}: {total} calls {'ok' if ok else 'DIFF'}", flush=True)
            bad += 0 if ok else 1
    print("mismatching rounds:", bad)
    return bad


if __name__ == "__main__":
    sys.exit(1 if main(int(sys.argv[1]) if len(sys.argv) > 1 else 8) else 0)
