# where the end-to-end leg spends its time: ingest timings + variant-scan stage trace of one short bench run
MC_DEBUG=1 MC_VC_TRACE=1 python bench.py --steps 2 --warmup 1 --no-cpu --no-sam > gpurun_out/e2e_trace.json 2> gpurun_out/e2e_trace.err
grep -E "ingest|mc_variant_scan" gpurun_out/e2e_trace.err | tail -40
python -c "
import json; d=json.load(open('gpurun_out/e2e_trace.json')); print(d['value'], d['ms_per_step'], 'e2e', d['e2e']['value'], d['e2e']['ms_per_step'], d['stages_ms_per_step'])"
