#!/bin/bash
# round-2 GPU call AN: conv kernel without the (removed) CTA-pair multicast variant: full suite + smoke + bench
cd "$GRAFT_REPO_ROOT" 2>/dev/null || cd /root/repo
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -p no:cacheprovider > gpurun_out/r2an_pytest.log 2>&1
echo "rc=$?" >> gpurun_out/r2an_pytest.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/r2an_smoke.log 2>&1
timeout 400 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2an_bench.json 2> gpurun_out/r2an_bench.err
grep -E "passed|failed" gpurun_out/r2an_pytest.log | tail -1; grep -E "^FAILED" gpurun_out/r2an_pytest.log | head; tail -1 gpurun_out/r2an_smoke.log
python - <<'PY'
import json
for l in open('gpurun_out/r2an_bench.json'):
    if l.startswith('{'):
        d=json.loads(l); print(round(d['value'],1), round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],1), d['clocks'])
PY
