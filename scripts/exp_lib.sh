# development aid, sourced: run "<env assignments>" <bench args...> appends a summary line to gpurun_out/exp.txt
run() {
  local envs="$1"; shift
  env $envs timeout 300 python bench.py --no-cpu "$@" > gpurun_out/sw.log 2> gpurun_out/sw.err
  python - "$envs | $*" <<PY >> gpurun_out/exp.txt
import json,sys
try:
    d=json.loads(open("gpurun_out/sw.log").read().strip().splitlines()[-1])
    p=d["phases_ms_per_step"]; a=d["phases_alone_ms"]
    print(sys.argv[1], "| value %.0f e2e %.0f busy %.0f total %.0f | em %.0f estep %.0f scan %.0f tensor %.0f table %.0f rng %.0f prep %.0f | alone: estep %.0f scan %.0f | roof %.3f" % (d["value"], d["e2e"]["value"], p["device_busy_ms"], p["total_ms"], p["em_ms"], p["estep_ms"], p["scan_ms"], p["tensor_ms"], p["table_ms"], p["host_rng_ms"], p["host_prep_ms"], a["estep_ms"], a["scan_ms"], d["roofline"]["frac"]))
except Exception as e:
    print(sys.argv[1], "FAILED", e, open("gpurun_out/sw.err").read()[-400:])
PY
}
