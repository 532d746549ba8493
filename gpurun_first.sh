set -x
cd $GRAFT_REPO_ROOT
nvidia-smi --query-gpu=name,memory.total --format=csv
python - <<'PY'
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, 'tests')
from sniffles_b200 import abi, binding, synth, config as sconfig
import oracle.oracle as orc, devcheck
blk = synth.config_block(1)
cfg = abi.Config.from_sniffles(sconfig.default_config())
ctx = binding.Context(0); ctx.set_config(cfg); ctx.load(blk)
want = orc.run(blk, cfg, 3, 1)
try:
    got = ctx.extract_leads()
    print("extract ok", len(got.leads), len(want.leads))
    devcheck.assert_leads_same(want.leads, got.leads)
    print("LEADS MATCH")
    got2 = ctx.cluster_call()
    print("cluster ok", len(got2.cand), len(want.cand))
    got3 = ctx.consensus()
    got2.alt = got3.alt
    print(ctx.timings())
except Exception as e:
    import traceback; traceback.print_exc()
PY
