"""e2e chunk-plan sweep (run on the GPU box): python tools_chunk_plan_sweep.py  -- QLDPC_CHUNK_PLAN = "first,max" in waves"""
import json, os, subprocess, sys
for plan in ("2,17", "2,34", "1,34", "2,68", "4,34", "1,17", "2,24", "3,24"):
    env = dict(os.environ, QLDPC_CHUNK_PLAN=plan)
    p = subprocess.run([sys.executable, "bench.py", "--steps", "4", "--warmup", "3", "--no-cpu"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().split("\n")[-1])
        print("plan", plan, "device", round(d["value"]), "e2e", round(d["e2e"]["value"]), round(d["e2e"]["ms_per_step"], 2), flush=True)
    except Exception as e:
        print(plan, "failed", p.stderr[-300:])
