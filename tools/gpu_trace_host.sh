B="--steps 30 --warmup 5 --no-closed-loop --no-multi-agent --no-configs --no-cpu-baseline --no-e2e-variants --e2e-steps 60"
# NOTE: kept for the record -- the RD_HOST_PIPE / RD_HOST_TRACE / RD_HOST_ACT_COPY switches these runs used were removed
# together with the rejected schedules (profiles/r2p_host_pipeline_trace.txt); the script no longer runs as is.
for sh in 8 4; do
echo "== shards $sh"; RD_HOST_TRACE=40 python bench.py $B --e2e-shards $sh 2>&1 >/dev/null | grep trace
done
echo "== f16"; RD_HOST_TRACE=40 python - <<'PY' 2>&1 | grep -E "trace|ms"
import sys, time, numpy as np
sys.path.insert(0,'.')
import bench, torch
from racing_dreamer_b200.host import HostSteppedEnv
wl=bench.workload_of(2)
for kw in (dict(), dict(lidar_dtype="float16")):
    h=HostSteppedEnv(bench.env_config(wl,4096,0,**kw),device="cuda:0",n_shards=8)
    a=bench.scripted_actions(wl,4096,0); h.reset()
    for k in range(60): h.step(a[k%50])
    t=time.perf_counter()
    for k in range(200): h.step(a[k%50])
    print(kw, (time.perf_counter()-t)/200*1e3, "ms")
    h.close()
PY
