"""Generates tests/golden/combine/: four synthetic samples that share planted SV sites, each written as an SNF by the UNMODIFIED reference
(oracle/pyref/harness.write_reference_snf), and the reference's own multi-sample combine over them (harness.reference_combine:
CombineTask.execute per contig, with oracle/pyref/stubs/edlib standing in for edlib's edit distance) as call dictionaries and VCF records.  Run in the build container (needs /root/reference):
    python tests/golden/make_combine_golden.py"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref"))
import harness                                     # noqa: E402
from sniffles_b200 import synth                    # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden", "combine")
CONTIGS = [350_000, 260_000]
N_SAMPLES = 4


def sample_block(k):
    return synth.generate(1004, CONTIGS, 24.0, len_model=1, len_mean=12000.0, len_sd=500.0, len_min=1000, len_max=80000, tech="ont",
                          sv_spacing=9000.0, sample=k + 1, site_keep=0.65, threads=4)


def main():
    os.makedirs(OUT, exist_ok=True)
    paths = []
    for k in range(N_SAMPLES):
        blk = sample_block(k)
        p = os.path.join(OUT, f"sample{k + 1}.snf")
        harness.write_reference_snf(blk, ["--sample-id", f"S{k + 1}"], p)
        for f in os.listdir(OUT):
            if ".snf.tmp_" in f:
                os.unlink(os.path.join(OUT, f))
        paths.append(p)
    names = [f"ctg{i + 1}" for i in range(len(CONTIGS))]
    cases = {}
    for label, args in (("default", []), ("no_alignment", ["--combine-pctseq", "0"]), ("strict_alignment", ["--combine-pctseq", "0.985", "--combine-separate-intra"]),
                        ("loose", ["--combine-match", "100", "--combine-low-confidence", "0.6", "--combine-output-filtered"])):
        config, calls, lines = harness.reference_combine(paths, list(zip(names, CONTIGS)), args)
        cases[label] = dict(args=args, calls={c: [harness.combine_call_dict(x) for x in v] for c, v in calls.items()}, vcf=lines)
        print(label, {c: len(v) for c, v in calls.items()}, len(lines), "VCF records")
    json.dump(dict(contigs=list(zip(names, CONTIGS)), samples=[os.path.basename(p) for p in paths], cases=cases), open(os.path.join(OUT, "expected.json"), "w"), indent=0)


if __name__ == "__main__":
    main()
