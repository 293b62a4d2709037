"""e2e time vs batch size (run on the GPU box)"""
import subprocess, sys, json
for F in (8192, 16384, 32768, 65536, 131072):
    p = subprocess.run([sys.executable, "bench.py", "--steps", "4", "--warmup", "2", "--no-cpu", "--frames", str(F)], capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().split("\n")[-1])
        print(F, "device ms", round(d["ms_per_step"], 3), "e2e_bits ms", round(d["e2e"]["ms_per_step"], 3), "e2e_llr ms", round(d["e2e_llr_api"]["ms_per_step"], 3), flush=True)
    except Exception as e:
        print(F, "failed", p.stderr[-300:])
