"""e2e chunk-size sweep (run on the GPU box): python tools_e2e_sweep.py"""
import importlib, os, subprocess, sys, json
for ch in (1480, 2960, 4440, 5920, 8880, 11840, 16280, 32560):
    env = dict(os.environ, QLDPC_CHUNK_FRAMES=str(ch))
    p = subprocess.run([sys.executable, "bench.py", "--steps", "4", "--warmup", "2", "--no-cpu"], env=env, capture_output=True, text=True)
    try:
        d = json.loads(p.stdout.strip().split("\n")[-1])
        print(ch, "device", round(d["value"]), "e2e_bits", round(d["e2e"]["value"]), "e2e_llr", round(d["e2e_llr_api"]["value"]), flush=True)
    except Exception as e:
        print(ch, "failed", p.stderr[-300:])
