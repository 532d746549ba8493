"""Multi-sample combine: plan + grouping + SVGroup.call against the reference's own CombineTask over four SNF files
(tests/golden/combine/, produced by tests/golden/make_combine_golden.py from the unmodified reference).
CPU tests use the grouping restatement in oracle/combine.py; the GPU test runs csrc/combine.cuh through the C ABI and must return the same arrays."""
import io
import json
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref"))
from oracle import combine as ocombine             # noqa: E402
from sniffles_b200 import combine, config as sconfig, snf, vcf   # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden", "combine")
EXPECTED = json.load(open(os.path.join(GOLD, "expected.json")))


def call_dict(c):
    gts = {int(k): [v[0], v[1], v[2], v[3], v[4], list(v[5]) if v[5] else None] + ([v[6]] if len(v) > 6 else []) for k, v in sorted(c.genotypes.items())}
    return dict(contig=c.contig, svtype=c.svtype, pos=c.pos, end=c.end, svlen=c.svlen, id=c.id, alt=c.alt, qual=c.qual, filter=c.filter,
                precise=bool(c.precise), support=c.support, fwd=c.fwd, rev=c.rev, nm=c.nm,
                cov=[c.coverage_upstream, c.coverage_start, c.coverage_center, c.coverage_end, c.coverage_downstream],
                info={k: c.info[k] for k in sorted(c.info)}, genotypes=gts, n_rnames=None if c.rnames is None else len(c.rnames))


def make_config(args):
    cfg = sconfig.default_config(*args)
    cfg.mode = "combine"
    cfg.snf_input_info, cfg.sample_ids_vcf = [], []
    for k, name in enumerate(EXPECTED["samples"]):
        path = os.path.join(GOLD, name)
        r = snf.SNFReader(path)
        sid = r.header["config"].get("sample_id") or os.path.splitext(name)[0]
        r.close()
        cfg.snf_input_info.append({"internal_id": k, "sample_id": sid, "filename": path})
        cfg.sample_ids_vcf.append((k, sid))
    return cfg


def run_case(case, grouper):
    cfg = make_config(case["args"])
    readers = {s["internal_id"]: snf.SNFReader(s["filename"]) for s in cfg.snf_input_info}
    calls, stats = {}, {}
    try:
        for tid, (name, length) in enumerate(EXPECTED["contigs"]):
            task = combine.CombineTask(tid, name, 0, length - 1, cfg)
            plan = combine.Plan()
            task.plan(readers, plan)
            out = grouper(plan, cfg)
            calls[name] = combine.CombineTask.emit([task], plan, out)[0]
            stats[name] = (plan, out)
    finally:
        for r in readers.values():
            r.close()
    return cfg, calls, stats


def cpu_grouper(plan, cfg):
    return ocombine.combine_groups(combine.plan_arrays(plan, cfg), cfg)


def norm(x):
    return json.loads(json.dumps(x))


@pytest.mark.parametrize("label", sorted(EXPECTED["cases"]))
def test_combine_calls_equal_the_reference(label):
    case = EXPECTED["cases"][label]
    cfg, calls, stats = run_case(case, cpu_grouper)
    for name, _ in EXPECTED["contigs"]:
        got = [norm(call_dict(c)) for c in calls[name]]
        want = case["calls"][name]
        assert len(got) == len(want), (label, name)
        for g, w in zip(got, want):
            assert g == w, (label, name, {k: (g[k], w[k]) for k in g if g[k] != w[k]})
    # the grouping is not trivial: some groups hold several samples, some groups were carried across chunks
    plan, out = stats[EXPECTED["contigs"][0][0]]
    sizes = np.bincount(out[0][:len(plan.cands)])
    assert sizes.max() >= 3 and len(plan.chunks) > len(plan.chains)


@pytest.mark.parametrize("label", sorted(EXPECTED["cases"]))
def test_combine_vcf_records_equal_the_reference(label):
    from harness import FakeFasta
    case = EXPECTED["cases"][label]
    cfg, calls, _ = run_case(case, cpu_grouper)
    buf = io.StringIO()
    w = vcf.VCFWriter(cfg, buf, reference=FakeFasta())
    for name, _ in EXPECTED["contigs"]:
        for c in calls[name]:
            w.write_call(c)
    assert buf.getvalue().splitlines() == case["vcf"]


@pytest.mark.gpu
@pytest.mark.parametrize("label", sorted(EXPECTED["cases"]))
def test_device_grouping_equals_the_restatement_and_the_reference(label):
    from sniffles_b200 import binding
    ctx = binding.Context(0)
    case = EXPECTED["cases"][label]
    cfg, calls, stats = run_case(case, lambda plan, c: ctx.combine_groups(plan, c))
    for name, _ in EXPECTED["contigs"]:
        plan, out = stats[name]
        ref = cpu_grouper(plan, cfg)
        n = len(plan.cands)
        assert np.array_equal(out[0][:n], ref[0][:n]) and np.array_equal(out[1][:n], ref[1][:n])
        used = ref[1][:n] >= 0
        assert np.array_equal(out[2][:n][used], ref[2][:n][used]) and np.array_equal(out[3][:n][used], ref[3][:n][used])
        assert [norm(call_dict(c)) for c in calls[name]] == case["calls"][name]
    assert ctx.launch_count() > 0


def test_edit_distance_restatements_agree():
    """oracle/combine.levenshtein (numpy rows) and the edlib stand-in the reference runs with (bit vectors on Python integers)"""
    import random
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref", "stubs"))
    import edlib
    r = random.Random(5)
    for _ in range(120):
        a = bytes(r.choice(b"ACGT") for _ in range(r.randrange(0, 200)))
        b = bytearray(a) if r.random() < 0.6 else bytearray(r.choice(b"ACGTN") for _ in range(r.randrange(0, 200)))
        for _ in range(r.randrange(0, 25)):
            if b and r.random() < 0.5:
                del b[r.randrange(len(b))]
            else:
                b.insert(r.randrange(len(b) + 1), r.choice(b"ACGT"))
        assert ocombine.levenshtein(a, bytes(b)) == edlib.levenshtein(a, bytes(b))
    assert ocombine.levenshtein(b"kitten", b"sitting") == 3 and ocombine.levenshtein(b"", b"abc") == 3 and ocombine.levenshtein(b"<DEL>", b"<DEL>") == 0


@pytest.mark.gpu
def test_device_edit_distance():
    """the kernel behind group.align_call against the restatement: block boundaries (63/64/65 rows), several passes (> 2048 rows), other alphabets"""
    import random
    from sniffles_b200 import binding
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pyref", "stubs"))
    import edlib
    r = random.Random(9)
    pairs = [(b"", b""), (b"", b"ACGT"), (b"A", b"A"), (b"A", b"C"), (b"<DEL>", b"<DEL>"), (b"<DEL>", b"<DUP>"), (b"kitten", b"sitting")]
    for la in (1, 2, 31, 63, 64, 65, 127, 128, 129, 500, 2047, 2048, 2049, 4100, 6000):
        a = bytes(r.choice(b"ACGT") for _ in range(la))
        for mode in range(3):
            if mode == 0:
                b = bytearray(a)
                for _ in range(max(1, la // 20)):
                    k = r.randrange(3)
                    if k == 0 and b:
                        del b[r.randrange(len(b))]
                    elif k == 1:
                        b.insert(r.randrange(len(b) + 1), r.choice(b"ACGTN"))
                    elif b:
                        b[r.randrange(len(b))] = r.choice(b"ACGTRYK")
            elif mode == 1:
                b = bytearray(r.choice(b"ACGT") for _ in range(r.randrange(1, 2 * la + 2)))
            else:
                b = bytearray(a[la // 3:]) + bytearray(r.choice(b"acgtn") for _ in range(la // 5))
            pairs.append((a, bytes(b)))
    ctx = binding.Context(0)
    got = ctx.edit_distances(pairs)
    want = [edlib.levenshtein(a, b) for a, b in pairs]
    assert got.tolist() == want
