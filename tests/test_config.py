"""The host-side config mirror reproduces the reference's defaults and derived constants
(tests/golden/config_defaults.json was dumped from the reference's SnifflesConfig)."""
import json
import os

from sniffles_b200 import config as sconfig, abi

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "config_defaults.json")
OUT_OF_SCOPE = ("combine_", "dev_cache", "dev_debug", "dev_monitor", "dev_population", "dev_progress", "dev_skip", "dev_dump", "dev_merge",
                "dev_combine", "dev_disable", "dev_profile", "dev_split", "low_memory", "re_qc", "reqc", "consensus_low", "coverage_shift_bins_min")


def test_defaults_and_derived_constants():
    with open(GOLDEN) as f:
        ref = json.load(f)
    for args, want in ref.items():
        mine = vars(sconfig.default_config(*args.split()))
        for k, v in want.items():
            if k.startswith(OUT_OF_SCOPE):
                continue
            assert k in mine, f"{k} missing for args {args!r}"
            assert mine[k] == v, f"{k}: {mine[k]!r} != reference {v!r} for args {args!r}"


def test_flattening():
    c = abi.Config.from_sniffles(sconfig.default_config("--mosaic"))
    assert (c.mapq, c.minsvlen, c.minsvlen_screen, c.cluster_binsize, c.consensus_kmer_len) == (20, 50, 45, 100, 6)
    assert c.cluster_merge_len == 0.27 and c.qc_nm_measure == 1 and c.dev_min_leads_cluster == 2
